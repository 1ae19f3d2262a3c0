#!/usr/bin/env python
"""Headline benchmark: TransformerLM_scaled training step (BASELINE.json config #4).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

One process per GPU (torchrun for N > 1, NCCL).  A "step" = forward + backward + gradient
all-reduce + AdamW on one synthetic batch of 64 x 256 tokens per GPU (weak scaling), replayed
from a CUDA graph.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

S = dict(vocab_size=80, embedding_dim=384, context_length=256, num_heads=6, num_layers=6, dropout=0.2,
         batch_size=64, base_lr=3e-4)
FLOP_PER_TOKEN = 3 * (6 * (24 * 384 ** 2 + 2 * (256 + 1) * 384) + 2 * 384 * 80)  # 67 438 080
METRIC = "train_tokens_per_sec_TransformerLM_scaled"
WORKLOAD = ("TransformerLM_scaled train step: V=80 C=384 T=256 NH=6 L=6 dropout=0.2, "
            "AdamW lr 3e-4 betas (0.9,0.95) wd 0.01; 64x256 tokens per GPU per step")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm": p["hbm_gbs"], "src": "measured"}
    except Exception:
        return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batches(n, B, T, V, seed):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randint(0, V, (B, T), generator=g), torch.randint(0, V, (B, T), generator=g)) for _ in range(n)]


# --------------------------------------------------------------------------- #
# CPU arm: the oracle port of the reference path on the host cores
# --------------------------------------------------------------------------- #
def cpu_train_tokens_per_sec(B, steps, warmup, threads=None):
    from oracle import drake_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = {k: S[k] for k in ("vocab_size", "embedding_dim", "context_length", "num_heads", "num_layers")}
    sd = O.synthetic_state_dict("TransformerLM", seed=42, **cfg)
    batches = synthetic_batches(steps + warmup, B, S["context_length"], S["vocab_size"], 42)
    opt = O.AdamW(sd, S["base_lr"])
    times = []
    torch.manual_seed(0)
    for i, (x, y) in enumerate(batches):
        t0 = time.perf_counter()
        _, _, grads = O.loss_and_grads("TransformerLM", sd, x, y, dropout=S["dropout"], training=True)
        opt.step(grads)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times)
    return B * S["context_length"] * len(times) / dt, dt / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 8
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    tps, spt, threads = cpu_train_tokens_per_sec(B, steps, warmup)
    sample = f"{steps} train steps (fwd+bwd+AdamW) of {B}x256 tokens, fp32 torch CPU, {threads} threads"
    line = {"impl": "reference", "metric": METRIC, "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": spt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": S["batch_size"], "seq_len": S["context_length"],
                       "parallelism": f"cpu{threads}", "sample": f"each timed step is {B}x256 tokens of that workload"},
            "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- #
# GPU arm
# --------------------------------------------------------------------------- #
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the roofline kernel (ncu --set full capture,
# profiles/r1_gemm_ffn1_metrics.txt; 13.80 MB read + 1.69 MB written: inputs from HBM, the 50 MB bf16 output is still L2-resident when the kernel ends)
ROOFLINE_TRAFFIC_BYTES = 15.49e6


def time_gemm_roofline(dev, pk):
    """Dominant kernel: the tcgen05 GEMM at the FFN1 shape (M=16384, N=1536, K=384) exactly as the training step
    launches it (bias + ReLU + ReLU bit mask).  R = 6 independent operand/output sets (together 400 MB > the 126 MB
    L2, so every launch reads its operands from HBM), 4 rounds over them captured in one CUDA graph, replays timed
    with CUDA events on the launching stream: device time per launch without host launch latency."""
    from drakegpt_b200 import ops
    M, N, K, R = 16384, 1536, 384, 6
    sets = []
    for _ in range(R):
        sets.append((torch.randn(M, K, device=dev).bfloat16(), torch.randn(N, K, device=dev).bfloat16(),
                     torch.zeros(N, device=dev), torch.empty(M, N, device=dev, dtype=torch.bfloat16),
                     torch.zeros((N // 32) * M, device=dev, dtype=torch.int32)))

    def launch_all():
        for a, w, bias, out, mask in sets:
            ops.raw_gemm(a, w, out, bias=bias, relu=True, relu_mask_out=mask)

    launch_all()
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            launch_all()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        g.replay()
    e1.record()
    e1.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / (reps * 4 * R)
    tf = 2.0 * M * N * K / t / 1e12
    return {"bound": "tensor", "achieved": tf, "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": tf / pk["bf16_burst"],
            "traffic": ROOFLINE_TRAFFIC_BYTES, "algorithmic_bytes": 2.0 * (M * K + N * K + M * N) + 4.0 * (N // 32) * M,
            "algorithmic_flop": 2.0 * M * N * K,
            "kernel": "gemm_tc_kernel<BN=256, bias|relu|mask_out> FFN1 16384x1536x384, operands from HBM (6 rotating sets)",
            "peak_source": pk["src"] + " burst", "us_per_launch": t * 1e6, "launches_timed": reps * 4 * R}


def run_ours(args):
    from drakegpt_b200 import _lib, ops
    from drakegpt_b200 import model as M
    from drakegpt_b200.graph import GraphedTrainStep
    from drakegpt_b200.parallel import init_from_env
    import torch.distributed as dist
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_gpu()
    pk = peaks()
    B, T, V = S["batch_size"], S["context_length"], S["vocab_size"]
    torch.manual_seed(42)
    model = M.TransformerLM(V, S["embedding_dim"], T, S["num_heads"], S["num_layers"], S["dropout"],
                            precision=args.precision).to(dev).train()
    r = model.runner()
    r.base_seed = 1000 + rank
    r.configure_optimizer(lr=S["base_lr"], betas=(0.9, 0.95))
    reducer = r.make_reducer() if world > 1 else None
    pool_host = [(x.pin_memory(), y.pin_memory()) for x, y in synthetic_batches(16, B, T, V, 42 + rank)]
    pool_dev = [(x.to(dev), y.to(dev)) for x, y in pool_host]
    n0 = ops.launch_count()
    step = GraphedTrainStep(r, B, T, reducer) if not args.no_graph else None
    launches_per_step = (ops.launch_count() - n0) // 3 if step is not None else None

    def one(i, pool):
        x, y = pool[i % len(pool)]
        if step is not None:
            return step.step(x, y)
        return r.train_step(x.to(dev, non_blocking=True), y.to(dev, non_blocking=True), reducer)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(pool, sync_each):
        for i in range(args.warmup):
            one(i, pool)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        last = 0.0
        for i in range(args.steps):
            loss = one(i, pool)
            if sync_each:
                last = loss.item()  # device -> host read of the step's loss
        e1.record()
        barrier()
        dt_dev = e0.elapsed_time(e1) * 1e-3
        dt = max(dt_dev, time.perf_counter() - t0) if sync_each else dt_dev
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), float(loss.item() if not sync_each else last)

    n1 = ops.launch_count()
    with ClockSampler(local) as cs:
        dt, loss = timed(pool_dev, sync_each=False)
    eager_launches = ops.launch_count() - n1
    clocks = cs.summary()
    dt_e2e, _ = timed(pool_host, sync_each=True)
    tokens = world * B * T * args.steps
    tps, tps_e2e = tokens / dt, tokens / dt_e2e
    if rank != 0:
        return
    roof = time_gemm_roofline(dev, pk)
    step_tf = tps / world * FLOP_PER_TOKEN / 1e12
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ctps, cspt, threads = cpu_train_tokens_per_sec(8, 2, 1)
        cpu = {"value": ctps, "unit": "tokens/s", "cores": threads, "kind": "port",
               "sample": f"2 train steps of 8x256 tokens (oracle port of the reference, fp32 torch CPU, {threads} threads, "
                         f"{cspt:.2f} s/step)"}
    gl = (launches_per_step * args.steps) if launches_per_step else eager_launches
    line = {"metric": METRIC, "value": tps, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if r.mode == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "global_batch": world * B, "seq_len": T, "parallelism": f"dp{world}",
                       "cuda_graph": step is not None,
                       "l2": "no explicit flush: each step streams > 1 GB of activations through a 126 MB L2"},
            "e2e": {"value": tps_e2e, "unit": "tokens/s", "h2d_bytes_per_step": 2 * B * T * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": gl, "clocks": clocks, "roofline": roof,
            "step_tensor_frac": {"achieved_tflops_per_gpu": step_tf, "of_burst": step_tf / pk["bf16_burst"],
                                 "of_sustained": step_tf / pk["bf16_sustained"], "flop_per_token": FLOP_PER_TOKEN,
                                 "peak_source": pk["src"]},
            "cpu_baseline": cpu, "final_loss": loss}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
