"""Development aid: per-item timeline of the persistent recurrent kernel (csrc/lstm_layer.cu).

Slots per (cta, item k): 0 step counter seen by the h producer, 1 last h tile issued, 2 first h tile landed (MMA
thread), 3 MMAs issued + commit, 4 accumulator seen by the epilogue, 5 epilogue stores done, 6 published, 7 c loaded;
8/9/10/11 = SM cycles the MMA thread waited for h stages / W stages / spent from first stage to last issue / waited for
the TMEM slot.  Times are %globaltimer ns.
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from code_intelligence_b200 import IssueEncoder

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=128)
ap.add_argument("--layer", type=int, default=1)
ap.add_argument("--ghz", type=float, default=1.9)
ap.add_argument("--B", type=int, default=1280)
ap.add_argument("--preroll", type=float, default=0.0, help="seconds of back-to-back encodes before the traced call (sustained clocks)")
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
if a.preroll > 0:
    import time
    ids_d = torch.from_numpy(ids).cuda()
    len_d = torch.full((a.B,), a.T, dtype=torch.int32, device="cuda")
    out_d = torch.empty((a.B, 2400), device="cuda")
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < a.preroll:
        for _ in range(4):
            enc.encode_ids_device(ids_d, len_d, out_d)
        torch.cuda.synchronize()
enc._lib.ie_debug_seq_trace(enc._h, a.layer, None, 0)
if a.preroll > 0:
    for _ in range(3):      # the trace of the LAST call is read back; the ones before keep the device busy
        enc.encode_ids_device(ids_d, len_d, out_d)
    torch.cuda.synchronize()
    print("phase ms", enc.last_phase_ms(), "mhz", enc.last_phase_mhz())
else:
    enc.encode_ids(ids)
ng = (a.B + 255) // 256
tiles = 38 if a.layer < 3 else 13
pairs = min(74, a.T * ng * tiles)
items = (a.T * ng * tiles + pairs - 1) // pairs
buf = np.zeros((2 * pairs, items, 12), dtype=np.int64)
n = enc._lib.ie_debug_seq_trace(enc._h, -1, buf.ctypes.data, buf.size)
print("records", n, "pairs", pairs, "items/pair", items, "ng", ng, "tiles", tiles)
tr = buf.astype(np.float64)
lead = tr[0::2]
sel = slice(8, items - 3)
us = 1e-3
cyc = 1.0 / (a.ghz * 1e3)
span = (lead[:, :, 6].max() - lead[:, 0, 0].min()) * us
print(f"layer span {span / 1e3:.3f} ms = {span / a.T:.2f} us per timestep = {span / a.T / ng:.2f} us per batch-step")
period = np.diff(lead[:, :, 3], axis=1)[:, sel] * us
print(f"item period per pair (commit to commit): mean {period.mean():.2f} us  p10 {np.percentile(period, 10):.2f}  p90 {np.percentile(period, 90):.2f}")
mma = (lead[:, :, 3] - lead[:, :, 2])[:, sel] * us
print(f"MMA phase (first h landed -> commit): mean {mma.mean():.2f} us  max {mma.max():.2f}")
idle = (lead[:, 1:, 2] - lead[:, :-1, 3])[:, sel] * us
print(f"issuer idle between items (commit k -> first h of k+1 landed): mean {idle.mean():.2f} us  p90 {np.percentile(idle, 90):.2f}")
flag_late = (lead[:, 1:, 0] - lead[:, :-1, 3])[:, sel] * us
print(f"step counter of item k+1 seen relative to commit of item k: mean {flag_late.mean():.2f} us  p10 {np.percentile(flag_late, 10):.2f}  p90 {np.percentile(flag_late, 90):.2f}  (negative = early)")
print(f"first h tile landed after counter seen: mean {((lead[:, :, 2] - lead[:, :, 0])[:, sel] * us).mean():.2f} us")
print("MMA thread per item (us @%.2f GHz): waiting h stages %.2f, waiting W stages %.2f, first stage -> last issue %.2f, waiting TMEM slot %.2f" % (
    a.ghz, lead[:, sel, 8].mean() * cyc, lead[:, sel, 9].mean() * cyc, lead[:, sel, 10].mean() * cyc, lead[:, sel, 11].mean() * cyc))
ep = (tr[:, :, 5] - tr[:, :, 4])[:, sel] * us
print(f"epilogue (accumulator seen -> stores done): mean {ep.mean():.2f} us max {ep.max():.2f}")
print(f"epilogue waits for accumulator after c loaded: mean {((tr[:, :, 4] - tr[:, :, 7])[:, sel] * us).mean():.2f} us")
pub = (tr[:, :, 6] - tr[:, :, 5])[:, sel] * us
print(f"stores done -> published: mean {pub.mean():.2f} us")
# dependency slack: item n needs every item of (t-1, g); when was the LAST of them published relative to our counter-seen time
C = ng * tiles
pub_t = np.full((a.T, ng), 0.0)
for p in range(pairs):
    for k in range(items):
        nidx = p + k * pairs
        if nidx >= a.T * C:
            break
        t, c = divmod(nidx, C)
        gidx = c // tiles
        pub_t[t, gidx] = max(pub_t[t, gidx], tr[2 * p, k, 6], tr[2 * p + 1, k, 6])
lat = []
for p in range(0, pairs, 7):
    for k in range(8, items - 3):
        nidx = p + k * pairs
        if nidx >= a.T * C:
            break
        t, c = divmod(nidx, C)
        if t == 0:
            continue
        lat.append(lead[p, k, 0] - pub_t[t - 1, c // tiles])
lat = np.array(lat) * us
print(f"counter seen after the last publish of (t-1, g): mean {lat.mean():.2f} us  p10 {np.percentile(lat, 10):.2f}  p90 {np.percentile(lat, 90):.2f}")
for p in (0, 1, 36, pairs - 1):
    r = lead[p]
    rel = (r - r[:, 2:3]) * us
    print(f"pair {p}: slots relative to first-h-landed (mean over items): " +
          " ".join(f"{s}:{rel[sel, s].mean():+.2f}" for s in range(8)))
