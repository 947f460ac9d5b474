"""Screened-supercharge statistics on a BASELINE config: candidates bracketed vs measured exactly."""
import ctypes, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from bench import CONFIGS, synth_points
from approximatenn_b200.api import gpu_backend, stage_times
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
n, d, k, tries, dtype = CONFIGS[name]
g = gpu_backend(dtype)
g.lib.gpu_init(); g.lib.annh_set_timing(1)
pts = synth_points(n, d, dtype)
out = (ctypes.c_ulonglong * 2)()
for it in range(2):
    g.lib.annb_supercharge_screen_stats(out, 1)
    r = g.precomp(pts, k, tries, seed=1001)
    g.lib.annb_supercharge_screen_stats(out, 0)
    print(name, "bracketed/row %.1f exact/row %.1f" % (out[0] / n, out[1] / n), stage_times(g), flush=True)
