"""The reference's OWN parity and timing programs (compare_results.c, time_results.c), built
unmodified by oracle/Makefile against its ann.c + our libann_b200 (GPU side) + its C path
(CPU side).  compare_results runs precomp on both backends from the same srandom() seed and
counts every differing graph entry, bucket-table entry and ULP of bases/row_means
(compare_results.c:123-171): the bar is 0.  Skipped where oracle/_ref/bin was not built."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def run(prog, *args):
    path = os.path.join(BIN, prog)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/bin not built (needs the reference checkout at build time)")
    out = subprocess.run([path, *args], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


@pytest.mark.parametrize("prog,args", [
    ("compare_results_f32", ["-n", "4096", "-d", "64", "-k", "16", "-t", "8", "-o", "3"]),
    ("compare_results_f64", ["-n", "4096", "-d", "32", "-k", "16", "-t", "8", "-o", "2"]),
    ("compare_results_f32", ["-o", "5"]),                       # the program's own defaults
    ("compare_results_f32", ["-n", "3000", "-d", "20", "-k", "10", "-t", "7", "-b", "3", "-s", "4", "-a", "2", "-r", "2", "-o", "2"]),
])
def test_reference_compare_results_reports_zero_diffs(prog, args):
    out = run(prog, *args)
    m = re.search(r"Average diffs for comp: (\S+)", out)
    assert m, out
    assert float(m.group(1)) == 0.0, out


@pytest.mark.parametrize("prog", ["compare_results_f32", "compare_results_f64"])
def test_reference_compare_results_query_mode(prog):
    out = run(prog, "-n", "4096", "-d", "32", "-k", "16", "-t", "8", "-y", "200", "-o", "3")
    m = re.search(r"Average diffs for query: (\S+)", out)
    assert m, out
    assert float(m.group(1)) == 0.0, out


def test_reference_time_results_runs_on_gpu_backend():
    out = run("time_results_f32", "-n", "16384", "-d", "16", "-k", "10", "-o", "3")
    assert re.search(r"\d", out), out
