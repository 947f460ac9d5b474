"""No CPU fallback: without a CUDA device the product library refuses to compute, loudly
(message on stderr, exit status 1 — the reference's own error behaviour, gpu_comp.c:15-19)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from approximatenn_b200.api import gpu_backend
pts = np.random.default_rng(0).standard_normal((256, 16)).astype(np.float32)
gpu_backend(np.float32).precomp(pts, 4, tries=2, seed=1)
print("COMPUTED")
""" % ROOT


def test_precomp_gpu_without_a_device_exits_with_a_message():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", SNIPPET], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 1
    assert "COMPUTED" not in out.stdout
    assert "approximatenn_b200" in out.stderr and "no CPU fallback" in out.stderr


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "approximatenn_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "ann_oracle" not in text and "liboracle" not in text and "libannref" not in text, f
