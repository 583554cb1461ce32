"""Development aid: per-step critical-path timeline of the persistent recurrent kernel (SM clocks -> us at 1.9 GHz)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from code_intelligence_b200 import IssueEncoder

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=128)
ap.add_argument("--layer", type=int, default=1)
ap.add_argument("--ghz", type=float, default=1.9)
ap.add_argument("--B", type=int, default=256)
a = ap.parse_args()
g = torch.Generator().manual_seed(1)
dims = [((800 if l == 0 else 2400), (2400 if l != 3 else 800)) for l in range(4)]
emb = (torch.rand(60000, 800, generator=g) * 0.2 - 0.1).numpy()
layers = []
for i, o in dims:
    k = 1.0 / np.sqrt(o)
    u = lambda *s: ((torch.rand(*s, generator=g) * 2 - 1) * k).numpy()
    layers.append(dict(w_ih=u(4 * o, i), w_hh=u(4 * o, o), b_ih=u(4 * o), b_hh=u(4 * o)))
enc = IssueEncoder().load_weights(emb, layers)
ids = torch.randint(2, 60000, (a.B, a.T), generator=g, dtype=torch.int64).numpy()
enc.encode_ids(ids)   # warm
enc._lib.ie_debug_seq_trace(enc._h, a.layer, None, 0)
enc.encode_ids(ids)
n_cta = (120 if a.layer < 3 else 100) if a.B <= 512 else (148 if a.layer < 3 else 78)
buf = np.zeros((n_cta, a.T, 12), dtype=np.int64)
n = enc._lib.ie_debug_seq_trace(enc._h, -1, buf.ctypes.data, buf.size)
print("records", n)
us = 1e-3   # trace is in ns (%globaltimer)
tr_all = buf.astype(np.float64)
sel = slice(8, a.T - 1)
pub = tr_all[:, :, 6]          # publish time of every CTA, every step
pas = tr_all[:, :, 0]          # barrier-passed time
lead = pub[::2]
print("global: step period us", np.diff(pas.min(0))[sel].mean() * us)
print("global: publish spread across CTAs (max-min) us: mean", ((pub.max(0) - pub.min(0))[sel]).mean() * us)
print("global: last publish(t) -> first barrier pass(t+1) us:", ((pas.min(0)[1:] - pub.max(0)[:-1])[sel]).mean() * us,
      " -> last pass:", ((pas.max(0)[1:] - pub.max(0)[:-1])[sel]).mean() * us)
late = (pub - pub.min(0, keepdims=True))[:, sel].mean(1)
print("global: CTAs publishing latest (cta, mean lag us):", [(int(i), round(late[i] * us, 2)) for i in np.argsort(-late)[:6]])
ep = (tr_all[:, :, 5] - tr_all[:, :, 4])[:, sel].mean(1)
print("global: epilogue duration us: mean", ep.mean() * us, "max", ep.max() * us)
cyc = 1.0 / (a.ghz * 1e3)
print("global: MMA thread per step (us @%.2f GHz): waiting for h stages %.2f, waiting for W stages %.2f, first-stage->all issued %.2f" % (
    a.ghz, tr_all[::2, sel, 8].mean() * cyc, tr_all[::2, sel, 9].mean() * cyc, tr_all[::2, sel, 10].mean() * cyc))
mm = (tr_all[::2, :, 3] - tr_all[::2, :, 2])[:, sel].mean(1)
print("global: MMA phase (first A landed -> last commit) us: mean", mm.mean() * us, "max", mm.max() * us)
names = ["0 barrier passed", "1 last A issued", "2 first A landed(MMA)", "3 MMAs issued+commit", "4 tfull seen", "5 epilogue stores done", "6 fenced+bar", "7 gx loads issued"]
for cta in (0, 2, 4, 80, n_cta - 2):
    tr = buf[cta].astype(np.float64)
    # per step, relative to slot 0 (barrier passed) of the same step; leader CTAs have MMA slots
    rel = (tr - tr[:, 0:1]) * us
    step = np.diff(tr[:, 0]) * us
    sel = slice(8, a.T - 1)
    print(f"cta {cta}: step period us mean={step[sel].mean():.2f} min={step[sel].min():.2f} max={step[sel].max():.2f}")
    for k in range(8):
        v = rel[sel, k]
        if np.all(tr[sel, k] == 0):
            continue
        print(f"    {names[k]:28s} +{v.mean():7.2f} us (min {v.min():7.2f} max {v.max():7.2f})")
    # time from publish (6) of step t to barrier passed (0) of step t+1
    gap = (tr[1:, 0] - tr[:-1, 6]) * us
    print(f"    publish(t) -> barrier passed(t+1): {gap[sel].mean():.2f} us")
