#!/usr/bin/env python
"""Benchmark of the Issue_Embeddings encoder hot path (BASELINE.json: issues/sec to 2400-d @ seq_len 512 batch 256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of 256 synthetic issues x 512 tokens (BASELINE.json configs[1]
shape; reference-deployed R4 encoder: L=4, E=800, H=2400, V=60000, random-init seed 1234): embedding gather, 4 hoisted
input-projection GEMMs, 4 x 512 recurrent LSTM steps, masked [mean|max|last] pool -> (256, 2400) f32.

Three consecutive steps (three batches of 256) ride one launch of the persistent recurrent kernels (ie_encoder_encode with
768 rows = IE_MAX_BATCH): the (batch, column-tile) MMA chains of the three batches are dealt over all 74 CTA pairs, and
while one batch is in its epilogue / step barrier the tensor pipe works on another.  Each step is still one batch of
256 issues with its own result rows; `single_batch` in the JSON line is the same measurement with one batch per launch.

* `value`      : whole-job issues/s with the token ids already resident in HBM (CUDA events on the launching stream,
                 barrier + synchronize on both sides, max over ranks; under torchrun each rank encodes its own batches
                 -- weak scaling, no data-path collective -- and the timed region ends with the ONE all-gather of the
                 2400-d outputs).
* `e2e`        : the same metric through the public host-buffer API (IssueEncoder.encode_ids == C-ABI ie_encoder_encode
                 with pinned host ids/lengths/out): H2D of ids+lengths and D2H of the (256,2400) result inside the
                 timed region, every step.
* `roofline`   : dominant kernel = lstm_step_kernel (the recurrent h_{t-1} W_hh^T + gates step of the 2400-wide
                 layers).  achieved = algorithmic FLOPs per launch (2*256*2400*9600 = 11.8 GFLOP) / average launch
                 duration, the latter from CUDA events recorded inside ie_encoder_encode around the 512 launches of each
                 layer (ie_encoder_last_phase_ms).  peak = MEASURED_PEAKS.json bf16_tflops_sustained.
* `cpu_baseline`: the CPU oracle (oracle/awd_lstm_ref.py, torch nn.LSTM fp32 == the modules the reference's fastai
                 model wraps) timed on this box's host cores on a bounded sample.
* `--impl reference`: times that CPU path alone (the reference's own encoder is not installable: fastai/spaCy absent,
                 no network -- see DESIGN.md); each step is a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, T = 256, 512
N_LAYERS, EMB, HID, VOCAB = 4, 800, 2400, 60000
FLOP_PER_TOKEN = 2 * sum(4 * o * (i + o) for i, o in [(800, 2400), (2400, 2400), (2400, 2400), (2400, 800)])  # 266.24e6
STEP_FLOP_2400 = 2.0 * B * 2400 * 9600   # one recurrent step of one 2400-wide layer, one batch of 256


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        time.sleep(0.05)
        return self.summarise(self.rows)

    @staticmethod
    def summarise(rows):
        """Median SM clock and board power over the samples taken UNDER LOAD (power >= 60 % of the highest sample: the
        sampler also sees the idle gaps between the arms, where the clock sits at its maximum)."""
        samples, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                clk, cmax = float(r[1]), float(r[2])
            except Exception:
                continue
            try:
                pw = float(r[3])
            except Exception:
                pw = None
            samples.append((clk, pw))
            mx.append(cmax)
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        powers = [pw for _, pw in samples if pw is not None]
        if powers:
            thr = 0.6 * max(powers)
            loaded = [(c, pw) for c, pw in samples if pw is not None and pw >= thr]
        else:
            loaded = samples
        clks = sorted(c for c, _ in loaded)
        pws = sorted(pw for _, pw in loaded if pw is not None)
        return {"sm_mhz": (clks[len(clks) // 2] if clks else None), "sm_max_mhz": (max(mx) if mx else None),
                "reasons": sorted(reasons), "samples": len(samples), "samples_under_load": len(loaded),
                "power_w": (pws[len(pws) // 2] if pws else None), "power_w_max": (max(powers) if powers else None)}


def usable_cpus():
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p) + 0.5)))
    except Exception:
        pass
    return max(1, n)


def cpu_oracle_setup():
    """Build the CPU oracle encoder and pick the torch thread count that maximises its throughput on this box
    (more threads than usable cores, or than the small per-step GEMMs can feed, makes it slower)."""
    import numpy as np
    import torch
    from oracle import awd_lstm_ref as R
    enc = R.make_encoder(1234, VOCAB, EMB, HID, N_LAYERS)
    cores = usable_cpus()
    probe = np.stack(R.synthetic_ids(8, 16, seed=1))
    best = (0.0, 1)
    cands = sorted({c for c in (cores, cores // 2, 32, 16, 8) if 1 <= c <= cores}, reverse=True)
    for th in cands:
        torch.set_num_threads(th)
        R.encode_padded(enc, probe, [16] * 8)
        t0 = time.perf_counter()
        R.encode_padded(enc, probe, [16] * 8)
        rate = 8 * 16 / (time.perf_counter() - t0)
        if rate > best[0]:
            best = (rate, th)
    torch.set_num_threads(best[1])
    return enc, best[1], best[0], cores


def cpu_oracle_rate(budget_s=20.0):
    """issues/s of the CPU oracle on a bounded sample (about `budget_s` seconds) of the step's workload."""
    import numpy as np
    from oracle import awd_lstm_ref as R
    enc, threads, tok_rate, cores = cpu_oracle_setup()
    # per-step cost on the CPU is dominated by streaming the weights, so tokens/s grows with batch: size the
    # sample from the probe conservatively and cap it at one full step
    sb = int(max(1, min(B, tok_rate * budget_s / T)))
    ids = np.stack(R.synthetic_ids(sb, T, seed=2))
    t0 = time.perf_counter()
    out = R.encode_padded(enc, ids, [T] * sb)
    dt = time.perf_counter() - t0
    assert out.shape == (sb, 3 * EMB)
    return sb / dt, dt, threads, cores, sb


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on this box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import awd_lstm_ref as R
    enc, threads, tok_rate, cores = cpu_oracle_setup()
    budget = 120.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    sb = int(max(1, min(B, tok_rate * budget / T)))
    ids = np.stack(R.synthetic_ids(sb, T, seed=3))
    for _ in range(args.warmup):
        R.encode_padded(enc, ids, [T] * sb)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        R.encode_padded(enc, ids, [T] * sb)
    dt = time.perf_counter() - t0
    val = sb * args.steps / dt
    sample = (f"{sb} of the {B} issues of a step (seq_len {T}), torch fp32 nn.LSTM oracle, {threads} threads "
              f"(best of a thread-count probe; {cores} usable cores)")
    print(json.dumps({
        "impl": "reference", "metric": "issues/sec to 2400-d @ seq_len 512 batch 256", "value": val, "unit": "issues/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: R4 encoder, seq_len 512, batch 256 (CPU arm: bounded sample per step)",
                   "sample_issues_per_step": sb},
        "cpu_baseline": {"value": val, "unit": "issues/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "issues/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from code_intelligence_b200 import IssueEncoder
    from oracle import awd_lstm_ref as R   # weights + synthetic ids generator + cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W = args.steps, max(args.warmup, 3)

    ref = R.make_encoder(1234, VOCAB, EMB, HID, N_LAYERS)
    emb, layers = ref.export_weights()
    enc = IssueEncoder(N_LAYERS, EMB, HID, VOCAB, 1, local).load_weights(emb, layers)
    del ref

    # distinct synthetic ids per step and per rank, resident in HBM for the `value` arm, pinned host for `e2e`
    g = torch.Generator().manual_seed(1234 + rank)
    ids_all = torch.randint(0, VOCAB, (K + W, B, T), generator=g, dtype=torch.int64)
    ids_all[ids_all == 1] = 0
    ids_all[:, :, 0] = 2
    ids_dev = ids_all.to(dev)
    ids_pin = ids_all.pin_memory()
    len_dev = torch.full((B,), T, dtype=torch.int32, device=dev)
    len_host = np.full(B, T, dtype=np.int32)
    out_dev = torch.empty((K * B, 3 * EMB), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * K * B, 3 * EMB), dtype=torch.float32, device=dev) if world > 1 else None
    out_pin = torch.empty((B, 3 * EMB), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- launch plan: steps are submitted kPerLaunch at a time (3 x 256 rows per ie_encoder_encode call) ----------
    kPerLaunch = max(1, enc.max_batch // B)   # 3; 5 with the experimental IE_ROT=1 kernel (csrc/lstm_rot.cu)
    def plan(first, count, per_launch):
        # a remainder launch (count % per_launch steps) goes first, so that the LAST launch of a region -- whose
        # phase events feed the roofline -- is a full one
        out, i = [], first
        r = count % per_launch
        if r:
            out.append((i, r))
            i += r
        while i < first + count:
            out.append((i, per_launch))
            i += per_launch
        return out

    ids_flat_dev = ids_dev.view((K + W) * B, T)
    ids_flat_np = ids_pin.view((K + W) * B, T).numpy()
    len_dev2 = torch.full((kPerLaunch * B,), T, dtype=torch.int32, device=dev)
    len_host2 = np.full(kPerLaunch * B, T, dtype=np.int32)
    out_pin2 = torch.empty((kPerLaunch * B, 3 * EMB), dtype=torch.float32).pin_memory()
    out_np2 = out_pin2.numpy()

    def run_device(first, count, per_launch):
        for (i, n) in plan(first, count, per_launch):
            enc.encode_ids_device(ids_flat_dev[i * B:(i + n) * B], len_dev2[:n * B],
                                  out_dev[(i - W) * B:(i - W + n) * B] if i >= W else out_dev[:n * B], stream)

    def device_arm(per_launch):
        run_device(0, W, per_launch)
        barrier()
        l0 = enc.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        run_device(W, K, per_launch)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_dev)          # the single collective of the bulk path
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        return float(t_ms.item()), enc.launch_count - l0, enc.last_phase_ms()

    # ---- device-resident arm ------------------------------------------------------------------------
    sampler = ClockSampler(local)
    ms_single, _, _ = device_arm(1)                  # one batch per launch (reported as `single_batch`)
    if rank == 0:
        sampler.start()
    ms_max, launches, phases = device_arm(kPerLaunch)  # three batches per launch: the bulk-encode mode
    value = world * B * K / (ms_max * 1e-3)
    single_value = world * B * K / (ms_single * 1e-3)

    # ---- end-to-end arm: host buffers through the public API ----------------------------------------------
    lib, h = enc._lib, enc._h
    def run_host(first, count):
        chk = 0.0
        for (i, n) in plan(first, count, kPerLaunch):
            rc = lib.ie_encoder_encode(h, ids_flat_np[i * B:(i + n) * B].ctypes.data, len_host2.ctypes.data, n * B, T,
                                       out_np2.ctypes.data, 0, None)
            assert rc == 0, lib.ie_last_error()
            chk += float(out_np2[0, 0])
        return chk
    run_host(0, W)
    barrier()
    t0 = time.perf_counter()
    checksum = run_host(W, K)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0        # host-blocking API: wall time == device time + copies
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / float(t_e2e.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peaks = measured_peaks()
        peak = (peaks or {}).get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
        # dominant kernel: the persistent recurrent kernel of the 2400-wide layers (one launch = all T steps of one
        # layer for the batches riding the launch): avg launch duration from the CUDA events recorded around it inside
        # ie_encoder_encode, last timed launch
        batches = kPerLaunch if K >= kPerLaunch else K   # batches riding the LAST timed launch (see plan())
        step_ms = phases["steps"][:N_LAYERS - 1]
        avg_launch_ms = sum(step_ms) / len(step_ms)
        flop_per_launch = STEP_FLOP_2400 * T * batches
        achieved = flop_per_launch / (avg_launch_ms * 1e-3) / 1e12
        avg_launch_us = avg_launch_ms * 1e3
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "lstm_seq_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "issues/sec to 2400-d @ seq_len 512 batch 256", "value": value, "unit": "issues/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: 1M-issue bulk encode shape, fixed seq_len 512, batch 256 per step, "
                                   "R4 encoder (L=4,E=800,H=2400,V=60000) random-init seed 1234",
                       "batch": B, "seq_len": T, "batches_per_launch": kPerLaunch, "parallelism": f"dp{world} (issues sharded, one all-gather of outputs)",
                       "l2": "inputs larger than L2: each step streams ~6.5 GB of workspace (Gx 5 GB f32) and new ids",
                       "operands": "bf16 weights/activations, f32 accumulate, f32 cell state and pooling"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "issues/s", "h2d_bytes_per_step": B * T * 8 + B * 4,
                    "d2h_bytes_per_step": B * 3 * EMB * 4 + 4},
            "single_batch": {"value": single_value, "unit": "issues/s", "ms_per_step": ms_single / K,
                             "note": "same measurement with one batch of 256 per launch"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "%s (persistent recurrent kernel, 2400-wide layers, "
                                   "%d batches in the last timed launch)" % (
                                       "lstm_rot_kernel" if batches > 3 else "lstm_wide_kernel", batches),
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "avg_launch_us": avg_launch_us,
                         "flop_per_launch": flop_per_launch,
                         "whole_step_tflops": FLOP_PER_TOKEN * B * T / (ms_max / K * 1e-3) / 1e12,
                         "phase_ms_last_step": phases},
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, dt, threads, cores, sb = cpu_oracle_rate()
            line["cpu_baseline"] = {"value": rate, "unit": "issues/s", "cores": threads, "kind": "port",
                                    "sample": f"{sb} issues x seq_len {T} ({sb}/{B} of a step), torch fp32 nn.LSTM "
                                              f"oracle, {threads} threads of {cores} usable cores, {dt:.1f} s"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
