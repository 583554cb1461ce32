"""Development aid (not product, not a test): run a list of GPU checks, each in its own subprocess under a timeout,
and append the results to gpurun_out/<name>.log.  Usage on the GPU box:

    python tools/gpu_check.py [--only gemm,tiny,...] [--log gpurun_out/check.log]

Each check compares the CUDA path with the CPU oracle (oracle/) -- the oracle is the checker here, never the
thing measured.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check_gemm():
    import ctypes as C
    import numpy as np
    import torch
    from code_intelligence_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    shapes = [(128, 16, 64, 0), (128, 80, 128, 0), (128, 240, 64, 0), (256, 480, 192, 0), (300, 600, 1600, 1),
              (1000, 250, 600, 2), (4096, 9600, 832, 0), (2048, 3200, 2432, 0)]
    res = []
    for (M, N, K, act) in shapes:
        a = rng.standard_normal((M, K), dtype=np.float32)
        b = (rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
        bias = rng.standard_normal(N, dtype=np.float32)
        d = np.zeros((M, N), dtype=np.float32)
        rc = lib.ie_debug_gemm(a.ctypes.data, b.ctypes.data, bias.ctypes.data, M, N, K, act, d.ctypes.data, 0)
        if rc != 0:
            res.append(dict(shape=(M, N, K, act), rc=rc, err=lib.ie_last_error().decode()))
            continue
        ab = torch.from_numpy(a).bfloat16().double()
        bb = torch.from_numpy(b).bfloat16().double()
        ref = ab @ bb.T + torch.from_numpy(bias).double()
        if act == 1:
            ref = ref.clamp_min(0)
        if act == 2:
            ref = torch.sigmoid(ref)
        err = np.abs(d - ref.numpy())
        res.append(dict(shape=(M, N, K, act), max_abs=float(err.max()), mean_abs=float(err.mean()),
                        ref_scale=float(ref.abs().mean()),
                        worst=[int(x) for x in np.unravel_index(err.argmax(), err.shape)]))
    return res


def _enc_parity(n_layers, emb_sz, n_hid, vocab, B, T, min_len=None, seed=1234, scale=1.0):
    import numpy as np
    import torch
    from code_intelligence_b200 import IssueEncoder
    from oracle import awd_lstm_ref as R
    t0 = time.time()
    ref = R.make_encoder(seed, vocab, emb_sz, n_hid, n_layers, scale=scale)
    emb, layers = ref.export_weights()
    enc = IssueEncoder(n_layers, emb_sz, n_hid, vocab, 1, 0).load_weights(emb, layers)
    t_load = time.time() - t0
    docs = R.synthetic_ids(B, T, seed=seed + 1, vocab_sz=vocab, min_len=min_len)
    lengths = np.array([len(d) for d in docs], dtype=np.int32)
    ids = np.full((B, T), 1, dtype=np.int64)
    for i, d in enumerate(docs):
        ids[i, :len(d)] = d
    t0 = time.time()
    got = enc.encode_ids(ids, lengths)
    t_gpu = time.time() - t0
    t0 = time.time()
    torch.set_num_threads(os.cpu_count())
    want = R.encode_padded(ref, ids, lengths)
    t_cpu = time.time() - t0
    m = R.parity_metrics(got, want)
    # negative control: oracle on permuted ids must NOT match
    perm = ids.copy()
    perm[:, 1:] = np.roll(perm[:, 1:], 1, axis=0) if B > 1 else (perm[:, 1:] + 7) % vocab
    neg = R.parity_metrics(R.encode_padded(ref, perm, lengths), want)
    m.update(cfg=(n_layers, emb_sz, n_hid, vocab, B, T, min_len, scale), t_load=t_load, t_gpu_first=t_gpu, t_cpu=t_cpu,
             neg_rel_l2=neg['rel_l2'], finite=bool(np.isfinite(got).all()), launches=enc.launch_count)
    return m


def check_tiny():
    return [_enc_parity(2, 64, 128, 1000, 3, 7, min_len=2), _enc_parity(3, 96, 200, 500, 130, 19, min_len=1)]


def check_r4_small():
    return [_enc_parity(4, 800, 2400, 60000, 4, 16, min_len=3), _enc_parity(4, 800, 2400, 60000, 200, 48, min_len=8)]


def check_n3():
    return [_enc_parity(3, 800, 2400, 60000, 32, 64, min_len=16, scale=3.0)]


def check_n3b():
    # B=256 so that the persistent kernel path is exercised with "trained-like" weights, T=96
    return [_enc_parity(3, 800, 2400, 60000, 256, 96, min_len=16, scale=3.0),
            _enc_parity(4, 800, 2400, 60000, 256, 96, min_len=16, scale=1.0)]


def check_dual():
    # B > 256: two 256-row batches ride one persistent launch (ping-pong)
    return [_enc_parity(3, 96, 200, 500, 300, 19, min_len=1), _enc_parity(4, 800, 2400, 60000, 512, 64, min_len=8)]


def check_wide():
    # 512 < B <= 768: three batches per launch, N = 256 pair tiles (lstm_wide.cu)
    return [_enc_parity(3, 96, 200, 500, 700, 19, min_len=1), _enc_parity(4, 800, 2400, 60000, 768, 64, min_len=8),
            _enc_parity(3, 800, 2400, 60000, 600, 40, min_len=4, scale=3.0)]


def check_speed():
    """B=256, T=512 R4 with device-resident inputs: per-encode CUDA-event time."""
    import numpy as np
    import torch
    from code_intelligence_b200 import IssueEncoder
    from oracle import awd_lstm_ref as R
    ref = R.make_encoder(1234)
    emb, layers = ref.export_weights()
    enc = IssueEncoder().load_weights(emb, layers)
    out = []
    for (B, T) in [(256, 128), (256, 512)]:
        ids = torch.from_numpy(np.stack(R.synthetic_ids(B, T, seed=7))).cuda()
        lengths = torch.full((B,), T, dtype=torch.int32, device='cuda')
        o = torch.empty((B, 2400), device='cuda')
        for _ in range(2):
            enc.encode_ids_device(ids, lengths, o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            enc.encode_ids_device(ids, lengths, o)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flop = 266.24e6 * B * T
        out.append(dict(B=B, T=T, ms=ms, issues_per_s=B / ms * 1e3, tflops=flop / ms / 1e9,
                        us_per_step_layer=ms * 1e3 / (T * 4)))
    return out


CHECKS = dict(gemm=check_gemm, tiny=check_tiny, r4_small=check_r4_small, n3=check_n3, n3b=check_n3b, dual=check_dual, wide=check_wide, speed=check_speed)

if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=','.join(CHECKS))
    ap.add_argument('--log', default=os.path.join(ROOT, 'gpurun_out', 'check.log'))
    ap.add_argument('--child', default=None)
    ap.add_argument('--timeout', type=int, default=600)
    a = ap.parse_args()
    if a.child:
        r = CHECKS[a.child]()
        print('RESULT ' + json.dumps(r))
        sys.exit(0)
    os.makedirs(os.path.dirname(a.log), exist_ok=True)
    with open(a.log, 'a') as f:
        for name in a.only.split(','):
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', name], capture_output=True,
                                   text=True, timeout=a.timeout, cwd=ROOT)
                tail = (p.stdout[-6000:] + '\n' + p.stderr[-3000:]).strip()
                status = p.returncode
            except subprocess.TimeoutExpired as e:
                tail, status = f'TIMEOUT after {a.timeout}s\n{(e.stdout or b"")[-2000:]}\n{(e.stderr or b"")[-2000:]}', 'timeout'
            msg = f'=== {name} status={status} wall={time.time() - t0:.1f}s\n{tail}\n'
            print(msg)
            f.write(msg)
            f.flush()
