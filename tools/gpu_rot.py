"""Development aid (not product, not a test): check the experimental rotating-schedule recurrent kernel (IE_ROT,
csrc/lstm_rot.cu) against the default kernels of the same library -- the two must agree bit for bit, because every
(row, unit) sees the same MMA tile shapes in the same K order -- and time both on the R4 encoder.

    python tools/gpu_rot.py [--T 512] [--iters 3] [--skip-small] [--proj] [--poolraw] [--rotvar] [--log gpurun_out/rot.jsonl]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rand_weights(n_layers, emb_sz, n_hid, vocab, seed=1234):
    rng = np.random.default_rng(seed)
    emb = rng.uniform(-0.1, 0.1, (vocab, emb_sz)).astype(np.float32)
    layers = []
    for l in range(n_layers):
        n_in = emb_sz if l == 0 else n_hid
        n_out = emb_sz if l == n_layers - 1 else n_hid
        k = 1.0 / np.sqrt(n_out)
        layers.append(dict(w_ih=rng.uniform(-k, k, (4 * n_out, n_in)).astype(np.float32),
                           w_hh=rng.uniform(-k, k, (4 * n_out, n_out)).astype(np.float32),
                           b_ih=rng.uniform(-k, k, 4 * n_out).astype(np.float32),
                           b_hh=rng.uniform(-k, k, 4 * n_out).astype(np.float32)))
    return emb, layers


KNOBS = ("IE_ROT", "IE_ROT_BATCHES", "IE_ROT_VARIANT", "IE_EMB_PROJ", "IE_POOL_RAW")


def make(cfg, weights, env=None):
    """env: development knobs read at handle creation, e.g. {"IE_ROT": 2, "IE_EMB_PROJ": 1}."""
    from code_intelligence_b200 import IssueEncoder
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    enc = IssueEncoder(*cfg, 1, 0).load_weights(*weights)
    for k in KNOBS:
        os.environ.pop(k, None)
    return enc


def variants(a):
    v = {"rot": {"IE_ROT": 2}}
    if a.proj:     # layer 0 from the per-token table, on the wide kernel (513..768 rows) and on the rotating kernel
        v["wide+proj"] = {"IE_EMB_PROJ": 1}
        v["rot+proj"] = {"IE_ROT": 2, "IE_EMB_PROJ": 1}
    if a.poolraw:  # last layer pooled from its f32 hidden states by a separate kernel
        v["wide+poolraw"] = {"IE_POOL_RAW": 1}
        v["rot+poolraw"] = {"IE_ROT": 2, "IE_POOL_RAW": 1}
    if a.proj and a.poolraw:
        v["rot+proj+poolraw"] = {"IE_ROT": 2, "IE_EMB_PROJ": 1, "IE_POOL_RAW": 1}
    if a.rotvar:   # kernel variants (IE_ROT_VARIANT: 1 = proxy fence in the watcher, 2 = 4-stage h ring, 3 = both), 6 batches
        for var in (1, 2, 3):
            v[f"rot.v{var}"] = {"IE_ROT": 2, "IE_ROT_VARIANT": var}
        v["rot.b6"] = {"IE_ROT": 2, "IE_ROT_BATCHES": 6}
        v["rot.b6.v1"] = {"IE_ROT": 2, "IE_ROT_BATCHES": 6, "IE_ROT_VARIANT": 1}
    return v


def ids_lengths(B, T, vocab, seed, ragged=True):
    rng = np.random.default_rng(seed)
    lengths = rng.integers(max(1, T // 3), T + 1, B).astype(np.int32) if ragged else np.full(B, T, np.int32)
    lengths[0] = T
    ids = rng.integers(2, vocab, (B, T)).astype(np.int64)
    for b in range(B):
        ids[b, lengths[b]:] = 1
    return ids, lengths


def compare(name, cfg, weights, cases, log, var):
    base = make(cfg, weights)
    others = {tag: make(cfg, weights, env) for tag, env in var.items()}
    assert others["rot"].max_batch == 1280 and base.max_batch == 768, (others["rot"].max_batch, base.max_batch)
    for (B, T) in cases:
        ids, lengths = ids_lengths(B, T, cfg[3], seed=B * 131 + T)
        want = base.encode_ids(ids, lengths)
        for tag, enc in others.items():
            if B > enc.max_batch:
                continue
            t0 = time.time()
            got = enc.encode_ids(ids, lengths)
            rec = dict(check=name, path=tag, B=B, T=T, equal=bool(np.array_equal(got, want)),
                       finite=bool(np.isfinite(got).all()), max_abs=float(np.abs(got - want).max()),
                       nbad_rows=int((np.abs(got - want).max(axis=1) > 0).sum()), sec=round(time.time() - t0, 2))
            if B <= 768 and T <= 64:
                rec["raw_equal"] = bool(np.array_equal(base.raw_features(ids), enc.raw_features(ids)))
            print(json.dumps(rec), flush=True)
            log.write(json.dumps(rec) + "\n")
            log.flush()
    return base, others


def timeit(enc, B, T, vocab, iters):
    import torch
    ids, lengths = ids_lengths(B, T, vocab, seed=7, ragged=False)
    ids_d = torch.from_numpy(ids).cuda()
    len_d = torch.from_numpy(lengths).cuda()
    out = torch.empty((B, enc.out_dim), dtype=torch.float32, device="cuda")
    for _ in range(2):
        enc.encode_ids_device(ids_d, len_d, out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        enc.encode_ids_device(ids_d, len_d, out)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / iters
    ph = enc.last_phase_ms()
    return dict(B=B, T=T, ms=round(ms, 3), issues_per_s=round(B / ms * 1e3, 1), ms_per_256=round(ms * 256 / B, 3),
                gemm=[round(x, 2) for x in ph["gemm"]], steps=[round(x, 2) for x in ph["steps"]])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=512)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--skip-small", action="store_true")
    ap.add_argument("--only-small", action="store_true")
    ap.add_argument("--proj", action="store_true", help="also check / time IE_EMB_PROJ=1 (layer 0 from the per-token table)")
    ap.add_argument("--poolraw", action="store_true", help="also check / time IE_POOL_RAW=1 (pooling by a separate kernel)")
    ap.add_argument("--rotvar", action="store_true", help="also check / time IE_ROT_VARIANT=1..3 and IE_ROT_BATCHES=6")
    ap.add_argument("--log", default=os.path.join(ROOT, "gpurun_out", "rot.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.log), exist_ok=True)
    log = open(a.log, "a")
    if not a.skip_small:
        cfg = (3, 96, 200, 500)
        compare("small", cfg, rand_weights(*cfg), [(300, 19), (700, 23), (1100, 17), (1280, 9)], log, variants(a))
    if a.only_small:
        return
    cfg = (4, 800, 2400, 60000)
    base, others = compare("r4", cfg, rand_weights(*cfg), [(768, 24), (1280, 40)], log, variants(a))
    runs = [(base, 768, "wide"), (others["rot"], 1280, "rot5"), (others["rot"], 768, "rot3")]
    for tag, enc in others.items():
        if tag != "rot":
            rows = enc.max_batch if tag.startswith("rot") else 768
            runs.append((enc, rows, f"{tag} ({rows} rows)"))
    for i, (enc, B, tag) in enumerate(runs):
        rec = timeit(enc, B, a.T, cfg[3], a.iters)
        rec["path"] = tag
        print(json.dumps(rec), flush=True)
        log.write(json.dumps(rec) + "\n")
        log.flush()
        if all(e is not enc for e, _, _ in runs[i + 1:]):
            enc.close()        # full-size workspaces are 20-45 GB each: free them as we go


if __name__ == "__main__":
    main()
