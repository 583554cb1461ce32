"""Profiling driver (run under ncu): R4 encoder, `--iters` device-resident encodes of a (B, T) batch after
`--warm` warm-up encodes.  Prints the CUDA-event phase times of the last encode (never a bench value under ncu)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=256)
ap.add_argument("--T", type=int, default=512)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--layers", type=int, default=4)
a = ap.parse_args()

import numpy as np
import torch
from code_intelligence_b200 import IssueEncoder

# weights: any finite values do for profiling; avoid the oracle import here
g = torch.Generator().manual_seed(1234)
dims = [((800 if l == 0 else 2400), (2400 if l != a.layers - 1 else 800)) for l in range(a.layers)]
emb = (torch.rand(60000, 800, generator=g) * 0.2 - 0.1).numpy()
layers = []
for i, o in dims:
    k = 1.0 / np.sqrt(o)
    u = lambda *s: ((torch.rand(*s, generator=g) * 2 - 1) * k).numpy()
    layers.append(dict(w_ih=u(4 * o, i), w_hh=u(4 * o, o), b_ih=u(4 * o), b_hh=u(4 * o)))
enc = IssueEncoder(a.layers).load_weights(emb, layers)
ids = torch.randint(2, 60000, (a.B, a.T), generator=g, dtype=torch.int64).cuda()
lengths = torch.full((a.B,), a.T, dtype=torch.int32, device="cuda")
out = torch.empty((a.B, 2400), device="cuda")
for _ in range(a.warm):
    enc.encode_ids_device(ids, lengths, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    enc.encode_ids_device(ids, lengths, out)
e1.record()
torch.cuda.synchronize()
print(json.dumps(dict(B=a.B, T=a.T, ms_per_encode=e0.elapsed_time(e1) / a.iters, phases=enc.last_phase_ms(),
                      finite=bool(torch.isfinite(out).all()))))
