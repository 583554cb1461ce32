"""BASELINE.json configs[1] / configs[3] at (a fraction of) their literal size: N synthetic issues of fixed seq_len 512
through the public bulk API on HOST token-id lists (bulk.encode_bulk_distributed: global length sort -> issue j to rank
j mod G -> pinned staging / H2D under the previous batch's kernels -> ie_encoder_encode -> one NCCL all-gather -> un-sort ->
D2H on rank 0), timed end to end with perf_counter after one small warm-up call.  Rank 0 prints one JSON line.

    python tools/bulk_scale.py --issues 1000000                     # 1 GPU
    torchrun --nproc-per-node 8 tools/bulk_scale.py --issues 1000000  # 8 GPUs
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--issues", dest="n", type=int, default=1000000)
ap.add_argument("--T", type=int, default=512)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from code_intelligence_b200 import IssueEncoder, bulk
# seeded random-init weights of the reference shape (torch-default ranges; no oracle import in tools that check nothing
# against it)
g = torch.Generator().manual_seed(1234)
dims = [((800 if l == 0 else 2400), (2400 if l != 3 else 800)) for l in range(4)]
emb = (torch.rand(60000, 800, generator=g) * 0.2 - 0.1).numpy()
layers = []
for i_, o_ in dims:
    k_ = 1.0 / np.sqrt(o_)
    u = lambda *sh: ((torch.rand(*sh, generator=g) * 2 - 1) * k_).numpy()
    layers.append(dict(w_ih=u(4 * o_, i_), w_hh=u(4 * o_, o_), b_ih=u(4 * o_), b_hh=u(4 * o_)))
enc = IssueEncoder(device=local).load_weights(emb, layers)
rng = np.random.default_rng(4321)              # the same global list on every rank, as the API expects
t0 = time.perf_counter()
ids = rng.integers(0, 60000, size=(a.n, a.T), dtype=np.int64)
ids[ids == 1] = 0
ids[:, 0] = 2
docs = list(ids)                               # N views of (T,) int64
gen_s = time.perf_counter() - t0
fn = lambda d: bulk.encode_sorted_batches_device(d, enc, min_batches_rule=False, to_host=False)
bulk.encode_bulk_distributed(docs[:world * 2560], fn, device=dev, to_host="rank0")      # warm-up (buffers, NCCL)
if world > 1:
    dist.barrier()
torch.cuda.synchronize(dev)
t0 = time.perf_counter()
res = bulk.encode_bulk_distributed(docs, fn, device=dev, to_host="rank0")
if rank != 0:
    torch.cuda.synchronize(dev)
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    assert isinstance(res, np.ndarray) and res.shape == (a.n, 2400) and np.isfinite(res[::1009]).all()
    sel = rng.choice(a.n, size=64, replace=False)
    direct = enc.encode_ids(ids[sel])
    print(json.dumps({"issues": a.n, "seq_len": a.T, "n_gpus": world, "seconds": float(t.item()), "issues_per_s": a.n / float(t.item()),
                      "result_bytes": int(res.nbytes), "ids_bytes": int(ids.nbytes), "host_list_generation_s": gen_s,
                      "rows_bit_equal_to_a_direct_encode": bool(np.array_equal(res[sel], direct)),
                      "api": "bulk.encode_bulk_distributed(host id lists) -> np.ndarray (N, 2400) on rank 0"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
