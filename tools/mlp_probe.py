"""Development aid: device-resident timing of the MLP head (BASELINE.json configs[4] shape, D_in -> 600 -> 600 -> 256) at
n = 2^20 rows under the development knob IE_MLP_CHUNK (rows per pass), each
checked against a torch f32 reference on a sample of rows.

    python tools/mlp_probe.py [--what "default,IE_MLP_CHUNK=65536"] [--n 1048576]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from code_intelligence_b200.mlp import MLPHead

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="default,IE_MLP_CHUNK=65536,default")
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(7)
rng = np.random.default_rng(0)
for d_in in (1600, 2400):
    dims = [d_in, 600, 600, 256]
    coefs = [(rng.standard_normal((dims[i], dims[i + 1])) / np.sqrt(dims[i])).astype(np.float32) for i in range(3)]
    ints = [(rng.standard_normal(dims[i + 1]) * 0.1).astype(np.float32) for i in range(3)]
    X = torch.randn((a.n, d_in), generator=g).mul_(0.1).to(dev)
    P = torch.empty((a.n, 256), dtype=torch.float32, device=dev)
    sel = torch.randint(0, a.n, (2048,), generator=g).to(dev)
    h = X[sel]
    for i in range(3):
        h = h @ torch.from_numpy(coefs[i]).to(dev) + torch.from_numpy(ints[i]).to(dev)
        h = torch.relu(h) if i < 2 else torch.sigmoid(h)
    for what in a.what.split(","):
        env = dict(kv.split("=") for kv in what.split("+") if "=" in kv)
        os.environ.update(env)
        head = MLPHead(coefs, ints, device=0)
        for k in env:
            os.environ.pop(k)
        head.predict_proba_device(X, P)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            head.predict_proba_device(X, P)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        err = float((P[sel] - h).abs().max())
        flop = 2.0 * a.n * (d_in * 600 + 600 * 600 + 600 * 256)
        print(json.dumps(dict(d_in=d_in, what=what, n=a.n, ms=round(ms, 3), rows_per_s=round(a.n / ms * 1e3), tflops=round(flop / ms / 1e9, 1),
                              max_abs_vs_torch_f32=err)), flush=True)
        assert err < 5e-3, err
        head.close()
