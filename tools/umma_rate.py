"""Development aid: cycles per tcgen05.mma (issue vs execution) for several shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from code_intelligence_b200 import _lib
lib = _lib.load()
out = np.zeros(2, dtype=np.int64)
iters = 2000
for grid in (1, 148):
    for mode, n in [(0, 16), (0, 32), (0, 80), (0, 128), (0, 160), (0, 240), (0, 256), (1, 64), (1, 160), (1, 256)]:
        if mode == 1 and grid == 148:
            g = 74
        else:
            g = grid
        rc = lib.ie_debug_umma_rate(mode, n, iters, 0, g, out.ctypes.data)
        m = 128 if mode == 0 else 256
        ideal = (128 * n / 256) if mode == 0 else (256 * n / 512)
        print(f"grid={g:3d} M={m} N={n:3d}: rc={rc} issue {out[0]/(iters*4):7.1f} cyc/MMA  exec {out[1]/(iters*4):7.1f} cyc/MMA  (tensor floor {ideal:.0f})")
for ce in (1, 2, 4):
    rc = lib.ie_debug_umma_rate(0, 80, iters, ce, 1, out.ctypes.data)
    print(f"M=128 N=80 commit+wait every {ce} groups: {out[1]/(iters*4):7.1f} cyc/MMA")
