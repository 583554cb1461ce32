"""Development aid: cycles per tcgen05.mma (issue vs execution) for several shapes; ntiles>1 = streaming operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from code_intelligence_b200 import _lib
lib = _lib.load()
out = np.zeros(2, dtype=np.int64)
iters = 3000
for ntiles in (1, 4):
    for mode, n in [(0, 16), (0, 80), (0, 160), (0, 240), (0, 256), (1, 64), (1, 160), (1, 256)]:
        rc = lib.ie_debug_umma_rate(mode, n, iters, 0, 1, ntiles, out.ctypes.data)
        m = 128 if mode == 0 else 256
        ideal = (128 * n / 256) if mode == 0 else (256 * n / 512)
        per_sm = 4096 + (n if mode == 0 else n // 2) * 32
        print(f"ntiles={ntiles} M={m} N={n:3d}: rc={rc} exec {out[1]/(iters*4):7.1f} cyc/MMA (tensor floor {ideal:.0f}); local smem operand bytes/MMA {per_sm} -> {per_sm/(out[1]/(iters*4)):.1f} B/clk")
