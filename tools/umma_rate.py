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

print("(bit4 = two interleaved accumulators) pair M=256 N=160, loop mimicking the recurrent kernel's issue thread (bit0: 2 barrier waits, bit1: tcgen05 fence, bit2: 1 commit, bit3: 2nd commit per 4 MMAs)")
for mimic in (0, 16, 15, 31):
    rc = lib.ie_debug_umma_rate(1, 160, iters, mimic, 1, 4, out.ctypes.data)
    print(f"  mimic={mimic:2d}: rc={rc} issue {out[0]/(iters*4):7.1f}  exec {out[1]/(iters*4):7.1f} cyc/MMA")
for mimic in (0, 16):
    rc = lib.ie_debug_umma_rate(1, 160, iters, mimic, 60, 4, out.ctypes.data)
    print(f"  60 pairs mimic={mimic:2d}: rc={rc} exec {out[1]/(iters*4):7.1f} cyc/MMA")

for n in (160, 240, 256):
    for mimic in (0, 16):
        rc = lib.ie_debug_umma_rate(1, n, iters, mimic, 1, 4, out.ctypes.data)
        print(f"  pair N={n} mimic={mimic:2d}: exec {out[1]/(iters*4):7.1f} cyc/MMA")

