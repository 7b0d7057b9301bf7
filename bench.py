#!/usr/bin/env python
"""bench.py -- decode tokens/s at batch 1 on B200 (BASELINE.json metric), plus roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload nano-168m-q80|qwen3-0.6b-q80|qwen3-0.6b-q4k|...]
    python bench.py --impl reference ...        # the reference's own OpenMP CPU engine on the host cores

A *step* is one pass of the hot path over one batch of synthetic input: a 16-token prompt is pushed through
the token-at-a-time path (infer.c:1258-1260), then greedy decode (temperature 0, repetition penalty 1.0) runs to
`seq`.  tokens/s is counted over the decode segment only, like the reference's own TPS.

  value  : device-resident loop (nb200_decode_greedy: token fed back on the GPU, inputs already in HBM),
           decode segments timed with CUDA events on the launching stream, max over ranks.
  e2e    : the same metric through the per-token C-ABI call a reference host makes (nb200_next_greedy ==
           generate_next_token): every token does a pinned H2D of the step descriptor + token id and a D2H of
           the resulting id; wall-clock over the decode segment.
  roofline : dominant kernel (W1|W3 matvec + SwiGLU) -- algorithmic bytes per launch / mean launch duration
           measured live with CUDA events around every launch of one profiling pass (graph/PDL off).
  cpu_baseline : the unmodified reference (oracle/_ref, Makefile flags) or, if absent, the oracle port, on the
           host cores over a bounded sample of the same workload.

N > 1 (torchrun): the path is batch-1 decode; ranks run independent replicas (one session per GPU, no data-path
collective) => "scaling": "weak".  Tensor-parallel decode of ONE session is a separate mode (--mode tp).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nano_b200 import modelfile as mf  # noqa: E402

WORKLOADS = {
    # name: (preset, quant, gs, seq)   -- BASELINE.json configs[1..3]
    "nano-168m-q80": ("nano-168m", mf.QUANT_Q80, 128, 512),
    "qwen3-0.6b-q80": ("qwen3-0.6b", mf.QUANT_Q80, 128, 2048),
    "qwen3-0.6b-q4k": ("qwen3-0.6b", mf.QUANT_Q4K, 0, 2048),
    "nano-168m-f32": ("nano-168m", mf.QUANT_F32, 0, 128),
    "nano-168m-q4k": ("nano-168m", mf.QUANT_Q4K, 0, 512),
    "qwen3-1.7b-q80": ("qwen3-1.7b", mf.QUANT_Q80, 128, 2048),
    "qwen3-4b-q80": ("qwen3-4b", mf.QUANT_Q80, 128, 4096),
    "toy-qwen3-q80": ("toy-qwen3", mf.QUANT_Q80, 64, 128),
}
FAST_FILE = {"qwen3-1.7b-q80": True, "qwen3-4b-q80": True}      # multi-GB files: synthesise codes/scales directly (seconds, not minutes)
PROMPT = 16


def _metric_name():
    """The headline metric exactly as BASELINE.json names it (tokens/s is `value`; the GB/s-vs-roofline half is `roofline`)."""
    try:
        return json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    except Exception:
        return "decode tokens/sec at batch=1; achieved HBM GB/s vs roofline"


METRIC = _metric_name()
CLASS_NAMES = ["embed", "qkv", "attention", "o_proj", "w13_swiglu", "w2", "classifier"]


def prompt_ids(spec, seq):
    ids = np.zeros(seq + 1, np.uint32)
    ids[:PROMPT] = [(17 + i % 10) if spec.arch == mf.ARCH_NANO else 1000 + i for i in range(PROMPT)]
    return ids


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU baseline (runs in a subprocess so OMP_* take effect before libgomp loads)
# --------------------------------------------------------------------------------------------------
CPU_CHILD = r"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, {root!r})
from oracle import bindings as ob
path, seq, sample_end, P, kind, flavour, arch = {path!r}, {seq}, {sample_end}, {P}, {kind!r}, {flavour!r}, {arch}
ids = np.zeros(seq + 1, np.uint32)
ids[:P] = [(17 + i % 10) if arch == 0 else 1000 + i for i in range(P)]
if kind == "reference":
    eng = ob.RefEngine(path, seq, flavour, penalty=1.0, temperature=0.0)
    step = lambda pos, pre: eng.next(ids, pos, pre)
else:
    eng = ob.NanoOracle(path, seq)
    ob.NanoOracle.lib().nor_set_threads(int(os.environ.get("OMP_NUM_THREADS", "1")))
    step = lambda pos, pre: eng.next_greedy(ids, pos, pre, 1.0)
for pos in range(P - 1):
    ids[pos + 1] = step(pos, 1)
t0 = time.perf_counter()
n = 0
for pos in range(P - 1, sample_end - 1):
    ids[pos + 1] = step(pos, 0); n += 1
dt = time.perf_counter() - t0
print(json.dumps({{"tokens": n, "seconds": dt}}))
"""


def cpu_baseline(workload, path, spec, seq, budget_s=25.0, quick=False):
    """Time the reference CPU engine on a bounded sample: prompt + the first decode positions of the same
    workload, sweeping OMP thread counts (README.md:73: N 'must be found by experiment')."""
    from oracle import bindings as ob
    flavour = ob.best_fast_flavour()
    kind = "reference" if flavour else "port"
    ncpu = os.cpu_count() or 1
    sample_end = min(seq, PROMPT + (32 if spec.n_embd >= 1024 else 64))
    cands = sorted({t for t in (8, 16, 32, 64, ncpu // 2, ncpu) if 1 <= t <= ncpu}) or [1]
    if quick:
        cands = [min(16, ncpu)]
    best = None
    t_start = time.time()
    for th in cands:
        if time.time() - t_start > budget_s:
            break
        env = dict(os.environ, OMP_NUM_THREADS=str(th), OMP_PROC_BIND="true", OMP_WAIT_POLICY="active")
        code = CPU_CHILD.format(root=ROOT, path=path, seq=seq, sample_end=sample_end, P=PROMPT, kind=kind,
                                flavour=flavour or "", arch=spec.arch)
        try:
            out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
            r = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception:
            continue
        tps = r["tokens"] / r["seconds"]
        if best is None or tps > best["value"]:
            best = {"value": tps, "cores": th}
    if best is None:
        return {"value": None, "unit": "tokens/s", "cores": 0, "kind": kind, "sample": "failed"}
    best.update({"unit": "tokens/s", "kind": kind,
                 "sample": f"{workload}: prompt {PROMPT} + decode positions {PROMPT - 1}..{sample_end - 2} "
                           f"({sample_end - PROMPT} tokens) of the seq-{seq} run; best of OMP_NUM_THREADS in {cands} "
                           f"with OMP_PROC_BIND=true OMP_WAIT_POLICY=active; host has {ncpu} logical cores; "
                           f"build {flavour or 'oracle port -O2 strict'}"})
    return best


# --------------------------------------------------------------------------------------------------
def dist_setup(n, backend=None):
    """One process per GPU (torchrun); NCCL when CUDA is present, gloo otherwise (CPU tests of the host logic)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
        return rank, world, local, dist
    return rank, world, local, None


def barrier_max(dist, local, value):
    """barrier + max over ranks of a float (the timing rule: a multi-GPU time is the max over ranks)."""
    if dist is None:
        return value
    import torch
    dev = f"cuda:{local}" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_tokens_per_s(world, steps, tokens_per_step, ms_per_rank_max):
    """Whole-job throughput of `world` independent batch-1 replicas: all ranks' tokens over the slowest rank's time."""
    return world * steps * tokens_per_step / (ms_per_rank_max * 1e-3)


def run_reference_arm(args, spec, quant, gs, seq, path):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(args.workload, path, spec, seq, budget_s=20.0, quick=(i > 0 or args.steps + args.warmup > 2))
        if i >= args.warmup and cb["value"]:
            vals.append(cb["value"])
    v = float(np.mean(vals)) if vals else None
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": (time.time() - t0) * 1e3 / max(1, args.steps + args.warmup),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8xint8->int32 + f32" if quant == mf.QUANT_Q80 else "f32",
            "data": "synthetic", "config": {"workload": f"{args.workload} greedy decode, seq={seq}, prompt={PROMPT}, max_seq_len={seq}", "mode": "reference CPU engine (oracle/_ref, Makefile flags; oracle port if absent)"},
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="nano-168m-q80", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="replicas", choices=["replicas", "tp"],
                    help="replicas: one independent batch-1 session per GPU (weak scaling, the default the driver runs); "
                         "tp: ONE session sharded over the GPUs through NVLink peer memory (strong scaling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exact", action="store_true", help="run the engine in exact (reference-order) mode")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-stream", action="store_true", help="forbid the streaming kernel")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    preset, quant, gs, seq = WORKLOADS[args.workload]
    spec = mf.PRESETS[preset]
    rank_env = int(os.environ.get("RANK", "0"))
    local_env = int(os.environ.get("LOCAL_RANK", "0"))
    if rank_env == 0 or int(os.environ.get("LOCAL_WORLD_SIZE", "1")) == 1:
        path = mf.cached_model(spec, quant, gs or 128, fast=FAST_FILE.get(args.workload, False))
    if args.impl == "reference":
        path = mf.cached_model(spec, quant, gs or 128, fast=FAST_FILE.get(args.workload, False))
        run_reference_arm(args, spec, quant, gs, seq, path)
        return

    from nano_b200 import engine as E
    rank, world, local, dist = dist_setup(args.gpus)
    if dist is not None:
        dist.barrier()
    path = mf.cached_model(spec, quant, gs or 128, fast=FAST_FILE.get(args.workload, False))           # every rank finds the file rank 0 wrote
    flags = (E.FLAG_EXACT if args.exact else 0) | (E.FLAG_NO_PDL if args.no_pdl else 0) | (E.FLAG_NO_GRAPH if args.no_graph else 0) | (E.FLAG_NO_STREAM if args.no_stream else 0)
    tp = args.mode == "tp" and world > 1
    if tp:
        # one rank per process: exchange the CUDA IPC handles of the exchange blocks through torch.distributed
        eng = E.Engine(path, seq, device=local, flags=flags, tp=(rank, world))
        handles = [None] * world
        dist.all_gather_object(handles, eng.tp_export())
        eng.tp_attach_ipc(handles)
        dist.barrier()
    else:
        eng = E.Engine(path, seq, device=local, flags=flags)
    n_dec = seq - PROMPT
    jobs = 1 if tp else world                                # sessions decoded concurrently

    # ---- warm-up (also brings clocks up) ----
    for _ in range(args.warmup):
        ids = prompt_ids(spec, seq)
        eng.decode_greedy(ids, PROMPT, seq)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier_max(dist, local, 0.0)
    launches0 = eng.launches
    t0 = time.perf_counter()
    dec_ms = 0.0
    for _ in range(args.steps):
        ids = prompt_ids(spec, seq)
        _pre, dec = eng.decode_greedy(ids, PROMPT, seq)
        dec_ms += dec
    wall = time.perf_counter() - t0
    launches = eng.launches - launches0
    dec_ms = barrier_max(dist, local, dec_ms)
    wall = barrier_max(dist, local, wall)
    clocks = sampler.stop() if rank == 0 else None
    value = aggregate_tokens_per_s(jobs, args.steps, n_dec, dec_ms)

    # ---- e2e: per-token C-ABI calls with host buffers ----
    e2e_steps = max(1, min(args.steps, 2))
    t_e2e = 0.0
    for _ in range(e2e_steps):
        ids = prompt_ids(spec, seq)
        for pos in range(PROMPT - 1):
            ids[pos + 1] = eng.next_greedy(ids, pos, 1)
        t1 = time.perf_counter()
        for pos in range(PROMPT - 1, seq - 1):
            ids[pos + 1] = eng.next_greedy(ids, pos, 0)
        t_e2e += time.perf_counter() - t1
    t_e2e = barrier_max(dist, local, t_e2e)
    e2e_value = jobs * e2e_steps * n_dec / t_e2e

    if rank != 0 and not tp:
        eng.close()
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event pass) ----
    peak, peak_src = peaks()
    ids = prompt_ids(spec, seq)
    eng.decode_greedy(ids, PROMPT, seq)                      # fill ids + KV with a real run
    nprof = min(64, seq - 1 - PROMPT)
    start = max(PROMPT, (seq // 2) - nprof // 2)
    eng.profile_tokens(ids, start, 4)                        # warm
    ms, cnt = eng.profile_tokens(ids, start, nprof)
    if rank != 0:                                            # tensor-parallel ranks had to take part in the profiling pass
        eng.close()
        dist.barrier(); dist.destroy_process_group()
        return
    E_, F_, Q_, K_ = spec.n_embd, spec.n_hidden, spec.q_dim, spec.kv_dim
    bpw = {mf.QUANT_F32: 4.0, mf.QUANT_Q80: 1.0 + 4.0 / max(gs, 1), mf.QUANT_Q4K: 148.0 / 256.0}[quant]
    mid = start + nprof / 2.0
    T_ = world if tp else 1                                  # a tensor-parallel rank streams 1/T of every matrix and KV head
    alg = {   # algorithmic bytes per launch (weights once + gains + KV rows; activations not counted)
        "qkv": (Q_ + 2 * K_) * E_ * bpw / T_ + 4 * E_ + 4 * K_ / T_,
        "attention": (8 * K_ * (mid + 1) + 4 * K_) / T_ + (8 * spec.hd if spec.arch == mf.ARCH_QWEN3 else 0),
        "o_proj": E_ * Q_ * bpw / T_,
        "w13_swiglu": 2 * F_ * E_ * bpw / T_ + 4 * E_,
        "w2": E_ * F_ * bpw / T_,
        "classifier": spec.vocab * E_ * bpw / T_ + 4 * E_,
        "embed": 4 * E_,
    }
    per_class = {}
    tot_ms = float(ms.sum())
    for i, name in enumerate(CLASS_NAMES):
        if cnt[i]:
            dur = float(ms[i]) / int(cnt[i]) * 1e-3
            per_class[name] = {"launches": int(cnt[i]), "mean_us": dur * 1e6, "share": float(ms[i]) / tot_ms,
                               "alg_bytes": alg[name], "gbs": alg[name] / dur / 1e9}
    dom = "w13_swiglu"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get(args.workload)
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": f"k_matvec<{'Q80' if quant == mf.QUANT_Q80 else 'Q4K' if quant == mf.QUANT_Q4K else 'F32'},SWIGLU> (W1|W3 + SwiGLU)",
            "achieved": per_class[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": per_class[dom]["gbs"] / peak,
            "traffic": traffic, "traffic_source": "ncu --set full capture, profiles/r1_ncu_full_multikernel.md" if traffic else None, "peak_source": peak_src, "alg_bytes_per_launch": alg[dom],
            "mean_launch_us": per_class[dom]["mean_us"], "share_of_step": per_class[dom]["share"],
            "per_kernel": per_class}
    avg_pos = (PROMPT + seq - 1) / 2.0
    bytes_tok = spec.bytes_per_token(quant, gs, avg_pos)
    per_gpu_gbs = bytes_tok * (value / world) / 1e9          # replicas: each GPU streams a whole model per token; tp: 1/T of it
    token_roof = {"alg_bytes_per_token": bytes_tok, "achieved_gbs_per_gpu": per_gpu_gbs,
                  "frac_of_peak": per_gpu_gbs / peak, "roofline_tok_s_per_session": peak * 1e9 / bytes_tok * (world if tp else 1)}
    persistent = eng.path.startswith("streaming")
    if persistent:
        # the step IS one kernel: a launch decodes n_dec tokens, so the dominant kernel's roofline is the token roofline.
        # per_kernel keeps the phase-by-phase profile of the same device code run as separate launches.
        kname = "k_decode_stream"
        launch_us = dec_ms / args.steps * 1e3
        ptraffic = None
        try:
            ptraffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))).get(args.workload + ":stream")
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": f"{kname} (one launch decodes {n_dec} tokens: all layers + classifier + argmax)",
                "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s", "frac": per_gpu_gbs / peak,
                "traffic": ptraffic * n_dec if ptraffic else None,
                "traffic_source": f"ncu --set full capture of one {kname} launch, per-token bytes x {n_dec} tokens, profiles/r2_ncu_full_stream.md" if ptraffic else None,
                "peak_source": peak_src, "alg_bytes_per_launch": bytes_tok * n_dec,
                "mean_launch_us": launch_us, "share_of_step": 1.0,
                "phase_profile_note": "per_kernel = the same phase code launched as separate kernels (graph/PDL off), CUDA events per launch",
                "per_kernel": per_class}

    cb = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_baseline(args.workload, path, spec, seq)

    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall * 1e3 / args.steps, "higher_is_better": True, "scaling": "strong" if tp else "weak",
        "vs_baseline": None, "dtype": {mf.QUANT_Q80: "int8xint8->int32 + f32", mf.QUANT_Q4K: "u4xu4->int32 + f32", mf.QUANT_F32: "f32"}[quant],
        "data": "synthetic",
        "config": {"workload": f"{args.workload} greedy decode, seq={seq}, prompt={PROMPT}, max_seq_len={seq}",
                   "parallelism": (f"tp{world}: one batch-1 session, row-sharded over {world} GPUs, activations exchanged through NVLink peer memory"
                                   if tp else f"{world} independent batch-1 replica(s)"), "mode": "exact" if args.exact else "fast",
                   "engine": eng.path,
                   "l2": "inputs larger than L2: %.0f MB of weights (+KV) streamed per token vs 126 MB L2" % (eng.weight_bytes / 1e6),
                   "timing": "CUDA events around the decode segment of each step on the launching stream; max over ranks"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": 48 * n_dec, "d2h_bytes_per_step": 4 * n_dec,
                "api": "nb200_next_greedy per token (pinned H2D of the 48 B step descriptor incl. token id, D2H of the next id)"},
        "gpu_launches": int(launches),
        "launches_per_token": eng.launches_per_token,
        "roofline": roof, "token_roofline": token_roof, "cpu_baseline": cb,
    }
    print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
