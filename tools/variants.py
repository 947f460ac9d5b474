"""Build-time variants of one kernel file, built here and timed side by side in ONE gpurun call.

    python tools/variants.py build  name1:"-DSCR_DEPTH=4" name2:"-DSCR_TABLES4=1" ...
    python tools/variants.py run    [env:NAME=VALUE ...] [bench args]   # on the GPU box: one bench line per variant

`build` recompiles annb_leaf.cu and annb_finish.cu with the extra flags and links them with the
other objects of the regular float build into
approximatenn_b200/build/variants/libann_b200_f32_<name>.so; `run` points bench.py at each of
them through ANN_B200_LIB_F32 (approximatenn_b200/api.py) and prints stage times and the
parity sample of every line.  Experiment tooling: the product library is the one build.py makes.
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from approximatenn_b200 import build as B  # noqa: E402

VDIR = os.path.join(B.HERE, "build", "variants")


def build(specs, sources=("annb_leaf.cu", "annb_finish.cu")):
    B.build()
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition(":")
        mine = {}
        for source in sources:
            obj = os.path.join(VDIR, f"{os.path.splitext(source)[0]}_{name}.o")
            subprocess.run([B.NVCC, *B.ARCH, "-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-DUSE_FLOAT",
                            *flags.split(), "-I", B.INCLUDE, "-I", B.CSRC, "-Xcompiler", "-fPIC", "-c",
                            os.path.join(B.CSRC, source), "-o", obj], check=True)
            mine[source] = obj
        objs = []
        for src in B.CU_SOURCES + B.C_SOURCES:
            o = os.path.join(B.HERE, "build", f"{os.path.splitext(src)[0]}_f32.o")
            objs.append(mine.get(src, o))
        out = os.path.join(VDIR, f"libann_b200_f32_{name}.so")
        subprocess.run([B.NVCC, *B.ARCH, "-shared", "-Xlinker", "-Bsymbolic", "-o", out, *objs,
                        "-lcudart", "-lm", "-lpthread", "-ldl"], check=True)
        print("built", out, flush=True)


def run(args):
    """`env:NAME=VALUE` arguments add runs of the regular library under that environment variable."""
    libs = [("default", None, None)] + [(os.path.basename(p)[len("libann_b200_f32_"):-3], p, None)
                                        for p in sorted(glob.glob(os.path.join(VDIR, "libann_b200_f32_*.so")))]
    libs += [(a[4:], None, a[4:]) for a in args if a.startswith("env:")]
    args = [a for a in args if not a.startswith("env:")]
    base = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--pair-queries", "0",
            "--recall-sample", "0", "--extra-config", "none"] + (args or ["--steps", "10", "--warmup", "3"])
    for rep in range(2):
        for name, path, setenv in libs:
            env = dict(os.environ)
            if path:
                env["ANN_B200_LIB_F32"] = path
            if setenv:
                env[setenv.split("=", 1)[0]] = setenv.split("=", 1)[1]
            r = subprocess.run(base, env=env, capture_output=True, text=True)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                st = d.get("stage_ms", {})
                print(f"{name:28s} rep{rep} device {d['ms_per_step']:.3f} ms  leaf {st.get('leaf', 0):.3f}  "
                      f"buckets {st.get('buckets', 0):.3f}  supercharge {st.get('supercharge', 0):.3f}  "
                      f"e2e {d['e2e']['ms_per_step']:.2f}  mismatch {d.get('parity_sample', {}).get('mismatch')}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(name, "FAILED", e, r.stderr[-400:], flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "build":
        build(sys.argv[2:])
    elif len(sys.argv) >= 2 and sys.argv[1] == "run":
        run(sys.argv[2:])
    else:
        print(__doc__)
