#!/usr/bin/env python
"""bench.py -- the BASELINE.json metrics of the RSSM hot path: imagined latent states/s and world-model train steps/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c2x16|c3|c4|c5] [--impl ours|reference|reference-cuda]

Workloads (BASELINE.json `configs`; a "step" is one pass of the hot path over one batch of synthetic input):
    c2     (default) imagination rollout, 1024 start states x horizon 15 per GPU, car_racer_config sizes   -> imagined states/s, weak scaling
    c2x16  the same with 16 384 start states per GPU
    c3     WorldModel.training_step, batch 16 x seq 64 of 64x64x3 frames per GPU (data parallel: flat-bucket gradient all-reduce
           + packed-scalar all-reduce over NCCL)                                                          -> train steps/s, weak scaling
    c4     16 384 start states x horizon 15 with GRU deter 4096, split over the N GPUs                   -> imagined states/s, STRONG scaling
    c5     one whole synthetic training iteration (replay gather -> 2 world-model steps -> warm start -> rollout -> 2 agent
           steps) with the gradient all-reduces at N GPUs                                                -> iterations/s, weak scaling

Arms: `ours` = this repository on the GPU; `reference` = the UNMODIFIED reference (oracle/_ref, a byte-for-byte copy of its Python
modules made by oracle/make_ref.py) on the host cores, bounded sample; `reference-cuda` = the same unmodified code through stock
PyTorch on the B200 (its real deployment mode) -- context, not the target.

One JSON line is printed by rank 0.  `value` = device-timed whole-job throughput with inputs resident in HBM; `e2e` = the same
through the public host-buffer call (pinned-host inputs copied H2D and results read back D2H inside the timed region);
`roofline` = the dominant kernel against the measured bf16 peak (MEASURED_PEAKS.json); `cpu_baseline` = the reference on this
box's host cores (N = 1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: (kind, B per GPU (c4: total), H / T, config overrides, description)
    "c2": ("rollout", 1024, 15, {}, "imagination rollout 1024 start states x horizon 15, car_racer_config sizes (BASELINE configs[1])"),
    "c2x16": ("rollout", 16384, 15, {}, "imagination rollout 16384 start states x horizon 15, car_racer_config sizes"),
    "c4": ("rollout", 16384, 15, {"hidden_state_dims": 4096}, "imagination rollout 16384 start states x horizon 15, GRU deter 4096, start states split over the GPUs (BASELINE configs[3])"),
    "c3": ("wm", 16, 64, {}, "WorldModel.training_step, batch 16 x seq 64 of 64x64x3 frames per GPU (BASELINE configs[2])"),
    "c5": ("iter", 50, 50, {}, "one training iteration at car_racer_config.yaml (batch 50 x seq 50 per GPU, horizon 30, 2 world-model + 2 actor-critic epochs), gradient all-reduce over the GPUs (BASELINE configs[4])"),
}
WM_FLOPS_PER_BT = 3 * 119.7e6      # 119.7 MFLOP forward per (b, t), x3 for forward + backward


def flops_per_state(cfg):
    """Dense forward FLOPs per imagined state as the reference computes them."""
    D = cfg["hidden_state_dims"]; Z = cfg["latent_state_dims"][0] * cfg["latent_state_dims"][1]; A = cfg["action_dims"]
    NB = cfg["critic_reward_buckets"]
    hp1, hp2 = cfg["dyn_pred_hidden_num_nodes_1"], cfg["dyn_pred_hidden_num_nodes_2"]
    gru = 2 * (3 * D * (Z + A + D))
    actor = 2 * ((D + Z) * cfg["hidden_layer_actor_1_size"] + cfg["hidden_layer_actor_1_size"] * cfg["hidden_layer_actor_2_size"] + cfg["hidden_layer_actor_2_size"] * 2 * A)
    prior = 2 * (D * hp1 + hp1 * hp2 + hp2 * Z)
    rew = 2 * ((D + Z) * cfg["rew_pred_hidden_num_nodes_1"] + cfg["rew_pred_hidden_num_nodes_1"] * cfg["rew_pred_hidden_num_nodes_2"] + cfg["rew_pred_hidden_num_nodes_2"] * NB)
    con = 2 * ((D + Z) * cfg["cont_pred_hidden_num_nodes_1"] + cfg["cont_pred_hidden_num_nodes_1"] * cfg["cont_pred_hidden_num_nodes_2"] + cfg["cont_pred_hidden_num_nodes_2"] * 1)
    return dict(gru=gru, total=gru + actor + prior + rew + con)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="MEASURED_PEAKS.json")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm), reasons=sorted(reasons))


def make_problem(workload, world=1):
    from dreamer_b200 import synthetic as W   # synthetic weights / inputs (deterministic numpy streams)
    kind, B, H, over, desc = WORKLOADS[workload]
    if workload == "c4":
        B = max(1, B // world)                # strong scaling: the 16 384 start states are split over the ranks
    cfg = dict(W.REF_CONFIG, **over)
    if kind == "rollout":
        cfg["horizon"] = H
    elif kind == "wm":
        cfg.update(horizon=H, sequence_length=H, batch_size=B)
    sd = W.make_state_dict(cfg, seed=0, actor_mu_zero=True)
    return kind, cfg, sd, B, H, desc


# ---------------------------------------------------------------------------------------------------------------------------
# the reference arms: the UNMODIFIED reference from oracle/_ref (CPU, or CUDA through stock PyTorch)
# ---------------------------------------------------------------------------------------------------------------------------
def _load_reference(cfg, sd, device):
    """-> (Dreamer instance, kind).  oracle/_ref holds byte-for-byte copies of the reference's modules (oracle/make_ref.py)."""
    from oracle import make_ref                                   # the reference arm may execute oracle/ (tier rule 4)
    if not make_ref.available():
        return None, "port"
    return make_ref.load_dreamer(cfg, sd, device), "reference"


def _fill_reference_buffer(d, cfg, n, seed=1):
    import numpy as np
    rng = np.random.default_rng(seed)
    for i in range(n):
        d.buffer.add_to_buffer(rng.integers(0, 256, size=(3, 64, 64)).astype(np.uint8), rng.uniform(-1, 1, 3).astype(np.float32),
                               float(rng.standard_normal()), float(rng.random() > 0.02))


def run_reference(args, rank, world, cuda):
    """--impl reference / reference-cuda.  Rank 0 only; every step is a bounded sample of the workload."""
    if rank != 0:
        return
    from dreamer_b200 import synthetic as W
    kind, cfg, sd, B, H, desc = make_problem(args.workload, world)
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(dev)
        torch.set_float32_matmul_precision("high")               # train_car_racer.py:13
    else:
        torch.set_num_threads(os.cpu_count() or 1)
    metric = {"rollout": "imagined latent states/sec", "wm": "world-model train steps/sec", "iter": "training iterations/sec"}[kind]
    unit = {"rollout": "states/s", "wm": "steps/s", "iter": "iterations/s"}[kind]
    rows = B if cuda else min(B, args.cpu_rows)
    note = ""
    if kind == "rollout":
        d, rkind = _load_reference(cfg, sd, dev)
        z0, h0, _, n = W.rollout_inputs(cfg, rows, H, seed=1234)
        if d is None:                                              # oracle/_ref missing: the oracle's port with the stock sampler
            from oracle import rssm as O
            fn = lambda: O.dream_episodes(sd, z0, h0, None, n)
        else:
            d.horizon = H
            z0, h0 = z0.to(dev), h0.to(dev)
            fn = lambda: d.dream_episodes(z0, h0)                  # Dreamer.py:143-175, unmodified, under no_grad as evaluate/rollout do
        units = rows * H
        sample = f"{rows} of {B} start states x horizon {H}, stock torch sampler, {'TF32 matmuls on the GPU' if cuda else 'fp32'}, no_grad"
        ctx = torch.no_grad()
    elif kind == "wm":
        Bs, Ts = (B, H) if cuda else (min(B, args.cpu_wm_batch), min(H, args.cpu_wm_seq))
        cfg_s = dict(cfg, batch_size=Bs, horizon=Ts, sequence_length=Ts)
        d, rkind = _load_reference(cfg_s, sd, dev)
        if d is None:
            print(json.dumps(dict(impl="reference", unavailable="oracle/_ref missing: run python -m oracle.make_ref in the build container")), flush=True)
            return
        obs, act, rew, cont, _ = (x.to(dev) for x in W.sequence_inputs(cfg_s, Bs, Ts, seed=4321))
        fn = lambda: d.world_model.training_step(obs, act, rew, cont)     # WorldModel.py:148-202 as shipped (fp16 autocast + GradScaler)
        units = 1.0 * (Bs * Ts) / (B * H)                           # fraction of a full-size step one sample step does
        sample = f"batch {Bs} x seq {Ts} of the batch {B} x seq {H} step, as shipped (fp16 autocast{'' if cuda else ': pathologically slow backward on CPU, SURVEY 8a a9'}); value scaled by the sample's share of (b, t) pairs"
        ctx = torch.enable_grad()
    else:
        bs, sl, hz = (cfg["batch_size"], cfg["sequence_length"], cfg["horizon"]) if cuda else (args.cpu_wm_batch, 16, 8)
        cfg_s = dict(cfg, batch_size=bs, sequence_length=sl, horizon=hz, buffer_size=2048)
        d, rkind = _load_reference(cfg_s, sd, dev)
        if d is None:
            print(json.dumps(dict(impl="reference", unavailable="oracle/_ref missing: run python -m oracle.make_ref in the build container")), flush=True)
            return
        _fill_reference_buffer(d, cfg_s, 512)

        def fn():
            d.train_world_model()                                   # Dreamer.py:228-242
            d.train_Agent()                                         # Dreamer.py:264-287
        units = 1.0 * (bs * sl) / (cfg["batch_size"] * cfg["sequence_length"])
        sample = f"batch {bs} x seq {sl}, horizon {hz} of the batch 50 x seq 50, horizon 30 iteration, as shipped; value scaled by the sample's share of (b, t) pairs"
        ctx = torch.enable_grad()
    times = []
    with ctx:
        for i in range(args.warmup + args.steps):
            if cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            if cuda:
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    total = sum(times)
    val = units * len(times) / total
    cores = torch.get_num_threads()
    line = dict(impl="reference-cuda" if cuda else "reference", metric=metric, value=val, unit=unit, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling="strong" if args.workload == "c4" else "weak",
                vs_baseline=None, dtype="tf32" if cuda else "f32", data="synthetic", config=dict(workload=desc, l2="n/a"),
                cpu_baseline=None if cuda else dict(value=val, unit=unit, cores=cores, kind=rkind, sample=sample),
                sample=sample, reference_kind=rkind,
                e2e=dict(value=val, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def cpu_baseline(args, kind, cfg, sd, B, H):
    """The reference (unmodified, oracle/_ref; the oracle's port if that copy is missing) on this box's host cores, bounded sample."""
    from dreamer_b200 import synthetic as W
    torch.set_num_threads(os.cpu_count() or 1)
    if kind != "rollout":
        Bs, Ts = min(B, args.cpu_wm_batch), min(H, args.cpu_wm_seq)
        cfg_s = dict(cfg, batch_size=Bs, horizon=Ts, sequence_length=Ts)
        d, rkind = _load_reference(cfg_s, sd, "cpu")
        if d is None:
            return None
        obs, act, rew, cont, _ = W.sequence_inputs(cfg_s, Bs, Ts, seed=4321)
        times = []
        for i in range(1 + max(1, args.cpu_reps // 2)):
            t0 = time.perf_counter()
            d.world_model.training_step(obs, act, rew, cont)
            if i:
                times.append(time.perf_counter() - t0)
        share = (Bs * Ts) / (B * H)
        return dict(value=share / statistics.median(times), unit="steps/s", cores=torch.get_num_threads(), kind=rkind,
                    sample=f"WorldModel.training_step as shipped (fp16 autocast) at batch {Bs} x seq {Ts}, scaled by its share of the batch {B} x seq {H} step's (b, t) pairs")
    Bs = min(B, args.cpu_rows)
    z0, h0, _, n = W.rollout_inputs(cfg, Bs, H, seed=1234)
    d, rkind = _load_reference(cfg, sd, "cpu")
    if d is None:
        from oracle import rssm as O
        fn = lambda: O.dream_episodes(sd, z0, h0, None, n)
    else:
        d.horizon = H
        fn = lambda: d.dream_episodes(z0, h0)
    times = []
    with torch.no_grad():
        for i in range(1 + args.cpu_reps):
            t0 = time.perf_counter()
            fn()
            if i:
                times.append(time.perf_counter() - t0)
    return dict(value=Bs * H / statistics.median(times), unit="states/s", cores=torch.get_num_threads(), kind=rkind,
                sample=f"{Bs} of {B} start states x horizon {H}, median of {len(times)} rollouts after 1 warm-up, Dreamer.dream_episodes unmodified, fp32, no_grad")


# ---------------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------------
class Harness:
    def __init__(self, args):
        import torch.distributed as dist
        self.args, self.dist = args, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed + K timed calls, L2 flushed (untimed) before each, CUDA events, max over ranks -> total ms."""
        for _ in range(warmup):
            self.flush.zero_(); fn()
        self.barrier()
        evs = []
        for _ in range(steps):
            self.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        self.barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def finish(self, holders=()):
        if self.world > 1:
            import gc
            for h in holders:                       # captured graphs hold NCCL work: release them before the process group
                h.__dict__.pop("_graphs", None)
            gc.collect()
            self.barrier()
            self.dist.destroy_process_group()


def time_gradient_all_reduce(hs, n_params, dev, world):
    """The one collective of the design's training path, timed alone: all-reduce (SUM) of a flat fp32
    gradient bucket of `n_params` elements over NCCL; device time, max over ranks; bus GB/s = 2 (n - 1) / n x bytes / time."""
    import torch
    buf = torch.zeros(n_params, device=dev)
    for _ in range(5):
        hs.dist.all_reduce(buf)
    hs.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    a.record()
    for _ in range(reps):
        hs.dist.all_reduce(buf)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    hs.dist.all_reduce(t, op=hs.dist.ReduceOp.MAX)
    ms = float(t.item())
    by = n_params * 4
    return dict(bucket_bytes=by, all_reduce_ms=ms, bus_gbs=2 * (world - 1) / world * by / (ms * 1e-3) / 1e9)


def bench_rollout(hs, args, cfg, sd, B, H, desc):
    import ctypes as C
    from dreamer_b200 import _lib as L, ops, synthetic as W
    from dreamer_b200.rollout import dream_episodes_host
    lib = L.load()
    dev, world, rank = hs.dev, hs.world, hs.rank
    seed = 1234 + (rank if args.workload != "c4" else 0)
    z0, h0, u, n = W.rollout_inputs(cfg, B * (world if args.workload == "c4" else 1), H, seed=seed)
    if args.workload == "c4":                                            # this rank's shard of the global batch
        sl = slice(rank * B, (rank + 1) * B)
        z0, h0, u, n = z0[sl], h0[sl], u[:, sl].contiguous(), n[:, sl].contiguous()
    model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in sd.items()})
    ro = ops.Rollout(model, B, H)
    z0d, h0d, ud, nd = (t.to(dev) for t in (z0, h0, u, n))
    info = ro.info()
    launches0 = lib.drm_launch_count()
    ro.run(z0d, h0d, ud, nd, want_idx=False)
    launches_per_rollout = lib.drm_launch_count() - launches0
    step = lambda: ro.run_graphed(z0d, h0d, ud, nd, want_idx=False)
    sampler = ClockSampler(hs.local)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    sampler.start()
    total_ms = hs.timed(step, args.steps, args.warmup)
    clocks = sampler.stop()
    states = B * H * world
    value = states * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public host-buffer call: start states from pinned host memory (the latent as its 32 class
    # indices per state -- it is a one-hot -- and h0 as fp32), rewards / continues read back into pinned host memory
    z0_idx = z0.reshape(B, -1, 32).argmax(-1).to(torch.uint8).contiguous().pin_memory()
    h0_pin = h0.contiguous().pin_memory()
    h2d = z0_idx.numel() + h0_pin.numel() * 4
    res = {}

    def e2e_step():
        out = dream_episodes_host(ro, z0_idx, h0_pin)
        res["d2h"] = sum(t.numel() * t.element_size() for t in out["host"])
    sync_ms = hs.timed(e2e_step, args.steps, args.warmup)                 # one synchronous call at a time
    # ... and as a user streams rollouts: a two-deep submit / result queue (rollout.HostRolloutQueue) -- call i + 1's start states cross
    # the host link on a copy stream while call i computes; every call still copies its own inputs from pinned host memory and reads
    # its own rewards / continues back.  Timed as ONE region around all K calls (device events, max over ranks); the L2 flushes between
    # the rollouts run inside that region and their own (event-bracketed) durations are subtracted.
    from dreamer_b200.rollout import HostRolloutQueue
    queue = HostRolloutQueue(ro)

    def pipelined(k):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flushes, tickets = [], []
        t0.record()
        for i in range(k):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); hs.flush.zero_(); b.record()
            flushes.append((a, b))
            tickets.append(queue.submit(z0_idx, h0_pin))
            if i >= 1:
                res["last"] = float(queue.result(tickets[i - 1])[0][0, 0])      # (touch the host copy of the previous call's result)
        res["last"] = float(queue.result(tickets[-1])[0][0, 0])
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) - sum(a.elapsed_time(b) for a, b in flushes)
    pipelined(max(args.warmup, 3))
    hs.barrier()
    t = torch.tensor([pipelined(args.steps)], dtype=torch.float64, device=dev)
    if world > 1:
        hs.dist.all_reduce(t, op=hs.dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_val = states * args.steps / (e2e_ms * 1e-3)

    # ---- the dominant kernel, timed live with CUDA events on its stream (the library brackets it while profiling is on) ----
    fl = flops_per_state(cfg)
    peaks = measured_peaks()
    lib.drm_profile_enable(1)
    prof_steps = max(3, min(args.steps, 10))
    for _ in range(prof_steps):
        hs.flush.zero_(); ro.run(z0d, h0d, ud, nd, want_idx=False)
    torch.cuda.synchronize()
    lib.drm_profile_enable(0)
    names = ["gru", "prior_l1", "prior_l2", "prior_cat", "heads_l1", "heads_l2", "heads_out", "other", "rollout_persist"]
    stages = {}
    for i, nm in enumerate(names):
        ms, cnt = C.c_double(), C.c_int64()
        lib.drm_profile_read(i, C.byref(ms), C.byref(cnt))
        if cnt.value:
            stages[nm] = dict(ms_per_step=ms.value / prof_steps, launches_per_step=cnt.value / prof_steps, us_per_launch=1e3 * ms.value / cnt.value)
    whole_tf = fl["total"] * B * H * args.steps / (total_ms * 1e-3) / 1e12   # per GPU (B is per rank)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload + ("_persist" if info["persistent"] else ""))
        traffic = tj and tj["dram_bytes_per_launch"]        # dram__bytes_read + write of one launch, ncu --set full (profiles/)
    except Exception:
        pass
    if info["persistent"]:
        k_us = stages["rollout_persist"]["us_per_launch"]
        k_tf = fl["total"] * B * H / (k_us * 1e-6) / 1e12
        roofline = dict(bound="tensor", kernel="rollout_persist_kernel (whole horizon: GRU + prior + sample + actor + reward/continue heads, tcgen05)",
                        achieved=k_tf, peak=peaks["bf16_sustained"], unit="TFLOP/s", frac=k_tf / peaks["bf16_sustained"], traffic=traffic,
                        peak_source=peaks["source"] + " bf16 sustained (the kernel spans the whole step)",
                        flops_per_launch=fl["total"] * B * H, us_per_launch=k_us, share_of_step=(k_us * 1e-3) / (total_ms / args.steps),
                        gru_flops_share=fl["gru"] / fl["total"],
                        note="latency-bound by construction: 15 steps x (6 dependent layers of one m-tile's MLP chain + the GRU epilogue); "
                             "per-tile timeline in profiles/persist_trace_r2.txt",
                        ctas=info["ctas"], gru_tile=info["gru_tile"])
    else:
        gru_us = stages["gru"]["us_per_launch"]
        gru_tf = fl["gru"] * B / (gru_us * 1e-6) / 1e12
        share = stages["gru"]["ms_per_step"] / sum(v["ms_per_step"] for v in stages.values())
        roofline = dict(bound="tensor", kernel="gru_pair_kernel / fused_gemm_kernel<EpiGru> (GRU gates, tcgen05)", achieved=gru_tf, peak=peaks["bf16_burst"],
                        unit="TFLOP/s", frac=gru_tf / peaks["bf16_burst"], traffic=traffic, peak_source=peaks["source"] + " bf16 burst",
                        share_of_step=share, share_note="share of the eager per-stage event times (PDL overlap is off while profiling)",
                        flops_per_launch=fl["gru"] * B, us_per_launch=gru_us,
                        whole_rollout=dict(achieved=whole_tf, peak=peaks["bf16_sustained"], frac=whole_tf / peaks["bf16_sustained"]), stages=stages)
    line = dict(metric="imagined latent states/sec", value=value, unit="states/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="strong" if args.workload == "c4" else "weak", vs_baseline=None, dtype="bf16",
                data="synthetic",
                config=dict(workload=desc, start_states_per_gpu=B, horizon=H, l2="flushed (256 MiB write) between timed iterations",
                            launch=("one CUDA graph per rollout: flag reset + latent zero-fill + 2 pack kernels + ONE persistent kernel for the whole horizon"
                                    if info["persistent"] else "one CUDA graph replay per rollout (7 launches per imagined step)"),
                            parallelism=f"start states sharded over {world} rank(s), no data-path collective"),
                e2e=dict(value=e2e_val, unit="states/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=res.get("d2h", 0), ms_per_step=e2e_ms / args.steps,
                         api="rollout.HostRolloutQueue (two-deep submit / result: the next call's host -> device copy overlaps this call's rollout)",
                         synchronous=dict(value=states * args.steps / (sync_ms * 1e-3), ms_per_step=sync_ms / args.steps,
                                          api="rollout.dream_episodes_host (one call at a time)")),
                gpu_launches=int(launches_per_rollout * args.steps), clocks=clocks, roofline=roofline,
                whole_rollout=dict(achieved_tflops=whole_tf, frac_of_bf16_sustained=whole_tf / peaks["bf16_sustained"], flops_per_state=fl["total"]))
    if args.workload == "c2":
        # the same rollout in the TF32 mode (fp32 operands rounded to TF32, tcgen05.mma kind::tf32: the precision class the reference's own
        # GPU runs use, train_car_racer.py:13) -- launch-per-stage kernels, replayed as one CUDA graph, device-resident inputs
        model32 = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in sd.items()}, precision="tf32")
        ro32 = ops.Rollout(model32, B, H)
        step32 = lambda: ro32.run_graphed(z0d, h0d, ud, nd, want_idx=False)
        for _ in range(4):
            step32()
        ms32 = hs.timed(step32, args.steps, 1)
        line["tf32_mode"] = dict(value=states * args.steps / (ms32 * 1e-3), unit="states/s", ms_per_step=ms32 / args.steps,
                                 note="DRM_PRECISION_TF32 handle; parity bound 1.25e-3 of the tensor's scale (profiles/parity_r2.md: worst 7.8e-4), "
                                      "7 launches per imagined step (the persistent kernel is bf16-only)")
    if world > 1:
        # the rollout itself shards with no exchange; what the same ranks exchange when they TRAIN on these rollouts is one flat gradient
        # bucket per optimiser group -- timed here (outside the timed region) so that every multi-GPU line carries the collective's cost
        coll = time_gradient_all_reduce(hs, 7757035, dev, world)
        coll["note"] = ("not part of this workload: the world model's 31 MB gradient bucket all-reduced alone over NCCL (what --workload c5 "
                        "runs inside every training iteration)")
        line["collective"] = coll
    return line, [ro] + ([ro32] if args.workload == "c2" else [])


def bench_wm(hs, args, cfg, sd, B, T, desc):
    """WorldModel.training_step (WorldModel.py:148-202): kernel forward + hand-scheduled BPTT + fused clip/AdamW, one CUDA graph."""
    from dreamer_b200 import _lib as L, synthetic as W
    lib = L.load()
    dev, world, rank = hs.dev, hs.world, hs.rank
    wm, _ = W.build_learners(cfg, sd, dev)
    obs, act, rew, cont, uu = (x.to(dev) for x in W.sequence_inputs(cfg, B, T, seed=4321 + rank))
    launches0 = lib.drm_launch_count()
    wm.training_step(obs, act, rew, cont, uniforms=uu)
    launches_per_step = lib.drm_launch_count() - launches0
    wm.enable_cuda_graphs(warmup=2)
    step = lambda: wm.training_step(obs, act, rew, cont, uniforms=uu)
    sampler = ClockSampler(hs.local)
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    sampler.start()
    total_ms = hs.timed(step, args.steps, args.warmup)
    clocks = sampler.stop()
    value = args.steps / (total_ms * 1e-3)
    # end to end: u8 frames + actions / rewards / continues from pinned host memory, the loss read back
    obs_u8 = obs.to(torch.uint8).cpu().pin_memory()
    host = [t.cpu().pin_memory() for t in (act, rew, cont)]
    h2d = obs_u8.numel() + sum(t.numel() * 4 for t in host)
    st_obs = torch.empty_like(obs_u8, device=dev)
    st = [torch.empty_like(t, device=dev) for t in host]
    loss_host = torch.zeros(1).pin_memory()

    def e2e_step():
        st_obs.copy_(obs_u8, non_blocking=True)
        for d_, h_ in zip(st, host):
            d_.copy_(h_, non_blocking=True)
        loss = wm.training_step(st_obs.float(), st[0], st[1], st[2], uniforms=uu)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e2e_ms = hs.timed(e2e_step, args.steps, args.warmup)
    peaks = measured_peaks()
    flops = WM_FLOPS_PER_BT * B * T
    tf = flops * args.steps / (total_ms * 1e-3) / 1e12
    roofline = dict(bound="tensor", kernel="whole training step (kernel forward + BPTT + optimiser) as one CUDA graph", achieved=tf,
                    peak=peaks["bf16_sustained"], unit="TFLOP/s", frac=tf / peaks["bf16_sustained"], traffic=None,
                    peak_source=peaks["source"] + " bf16 sustained", flops_per_launch=flops,
                    note="119.7 MFLOP forward per (b, t) x 3 (SURVEY.md section 8d); the hand-scheduled backward's contractions run on drm_gemm_tf32, the conv stacks' backward is cuDNN through torch autograd (DESIGN.md sections 4e, 6)")
    line = dict(metric="world-model train steps/sec", value=value, unit="steps/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=desc, batch_per_gpu=B, seq=T, global_batch=B * world, l2="flushed (256 MiB write) between timed iterations",
                            launch="one CUDA graph replay per step (captured after 2 eager steps)",
                            parallelism=f"data parallel over {world} rank(s): flat 31 MB gradient bucket + packed loss scalars all-reduced over NCCL"),
                samples_per_s=value * B * world,
                e2e=dict(value=args.steps / (e2e_ms * 1e-3), unit="steps/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4, ms_per_step=e2e_ms / args.steps),
                gpu_launches=int(launches_per_step * args.steps), gpu_launches_note="kernels of this library per step (the conv stacks' autograd graph adds library kernels)",
                clocks=clocks, roofline=roofline)
    return line, [wm]


def bench_iteration(hs, args, cfg, desc):
    """One training iteration of Dreamer.train (Dreamer.py:340-341): train_world_model() + train_Agent() on a synthetic replay ring."""
    import numpy as np
    from dreamer_b200 import _lib as L
    from dreamer_b200.hotpath import HotPath
    lib = L.load()
    dev, world, rank = hs.dev, hs.world, hs.rank
    cfg_it = dict(cfg, buffer_size=8192)
    hp = HotPath(cfg_it, dev)
    rng = np.random.default_rng(1 + rank)
    n_it = 4096
    hp.buffer.add_batch(rng.integers(0, 256, size=(n_it, 3, 64, 64)).astype(np.uint8), rng.uniform(-1, 1, (n_it, 3)).astype(np.float32),
                        rng.standard_normal(n_it).astype(np.float32), (rng.random(n_it) > 0.02).astype(np.float32))

    def iteration():
        hp.train_world_model()
        hp.train_Agent()
    launches0 = lib.drm_launch_count()
    iteration()
    launches_per_it = lib.drm_launch_count() - launches0
    iteration()
    hp.world_model.enable_cuda_graphs(warmup=1)
    hp.agent.enable_cuda_graphs(warmup=1)
    hp.cuda_graphs = True
    sampler = ClockSampler(hs.local)
    for _ in range(3):
        iteration()
    torch.cuda.synchronize()
    sampler.start()
    total_ms = hs.timed(iteration, args.steps, args.warmup)
    clocks = sampler.stop()
    value = args.steps / (total_ms * 1e-3)
    # the collective inside it, timed alone: the world model's flat gradient bucket (all-reduce SUM)
    coll = None
    if world > 1:
        n_wm = sum(p.numel() for p in hp.world_model.parameters())
        coll = time_gradient_all_reduce(hs, n_wm, dev, world)
        coll.update(per_iteration=f"{cfg['WM_epochs']} x this bucket + {cfg['AC_epochs']} x (1.47 MB actor + 1.67 MB critic) + packed scalars",
                    share_of_iteration=cfg["WM_epochs"] * coll["all_reduce_ms"] / (total_ms / args.steps))
    line = dict(metric="training iterations/sec", value=value, unit="iterations/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=desc, l2="flushed (256 MiB write) between timed iterations",
                            launch="training steps and rollouts replayed as CUDA graphs", parallelism=f"data parallel over {world} rank(s); environment stepping excluded"),
                e2e=dict(value=value, unit="iterations/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0,
                         note="the replay ring is HBM-resident (Buffer mirror): an iteration moves no host data"),
                gpu_launches=int(launches_per_it * args.steps), clocks=clocks, collective=coll,
                roofline=dict(bound="tensor", kernel="whole iteration", achieved=None, peak=measured_peaks()["bf16_sustained"], unit="TFLOP/s", frac=None, traffic=None))
    return line, [hp.world_model, hp.agent]


def secondary_metrics(dev, peaks, flush):
    """HBM-bound kernels against the measured HBM peak and the large-batch rollout points -- extra keys, N = 1 only."""
    import ctypes as C
    import numpy as np
    from dreamer_b200 import _lib as L, ops, synthetic as W
    lib = L.load()
    out = {}

    def dev_time(fn, reps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts) * 1e-3

    cap, B, Ls = 20000, 1024, 50
    ring = [torch.randint(0, 256, (cap, 3, 64, 64), dtype=torch.uint8, device=dev), torch.rand(cap, 3, device=dev),
            torch.rand(cap, 1, device=dev), torch.ones(cap, 1, device=dev)]
    starts = torch.from_numpy(np.random.RandomState(1).randint(0, cap - Ls, size=B))
    t = dev_time(lambda: ops.replay_gather(*ring, starts, Ls), reps=5)
    by = 61480 * B * Ls
    out["replay_gather"] = dict(workload="1024 x 50 windows of 64x64x3 u8 frames -> fp32", bytes_per_launch=by, us=t * 1e6, achieved_gbs=by / t / 1e9,
                                peak_gbs=peaks["hbm"], frac=by / t / 1e9 / peaks["hbm"])
    del ring
    n = 1 << 22
    lg = torch.randn(n, 32, device=dev); u = torch.rand(n, device=dev)
    t = dev_time(lambda: ops.categorical32(lg, u), reps=5)
    by = n * (128 + 4 + 128 + 1)
    out["categorical32"] = dict(workload="4 Mi rows x 32 classes", bytes_per_launch=by, us=t * 1e6, achieved_gbs=by / t / 1e9, peak_gbs=peaks["hbm"],
                                frac=by / t / 1e9 / peaks["hbm"])
    del lg, u
    # the launch-per-stage chain on the default workload, for comparison with the persistent kernel
    cfg2 = dict(W.REF_CONFIG, horizon=15)
    model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in W.make_state_dict(cfg2, seed=0, actor_mu_zero=True).items()})
    ro = ops.Rollout(model, 1024, 15)
    z0, h0, uu, nn_ = (x.to(dev) for x in W.rollout_inputs(cfg2, 1024, 15, seed=1234))
    lib.drm_set_option(b"persist", 0)
    try:
        t_chain = dev_time(lambda: ro.run_graphed(z0, h0, uu, nn_, want_idx=False), reps=10, warm=4)
    finally:
        lib.drm_set_option(b"persist", 1)
    out["rollout_c2_launch_per_stage"] = dict(workload="1024 x 15 with option persist = 0 (7 launches per imagined step, one graph replay)",
                                               ms_per_rollout=t_chain * 1e3, states_per_s=1024 * 15 / t_chain)
    del ro, model
    # the north star's large-batch points on this one GPU (GRU stage vs the measured bf16 burst peak)
    for wl in ("c2x16", "c4"):
        _, Bn, Hn, over, desc = WORKLOADS[wl]
        cfgn = dict(W.REF_CONFIG, horizon=Hn, **over)
        model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in W.make_state_dict(cfgn, seed=0, actor_mu_zero=True).items()})
        ro = ops.Rollout(model, Bn, Hn)
        z0, h0, uu, nn_ = (x.to(dev) for x in W.rollout_inputs(cfgn, Bn, Hn, seed=1234))
        t = dev_time(lambda: ro.run(z0, h0, uu, nn_, want_idx=False), reps=3, warm=2)
        lib.drm_profile_enable(1)
        ro.run(z0, h0, uu, nn_, want_idx=False)
        torch.cuda.synchronize()
        lib.drm_profile_enable(0)
        ms, cnt = C.c_double(), C.c_int64()
        lib.drm_profile_read(0, C.byref(ms), C.byref(cnt))
        for i in range(1, 9):
            lib.drm_profile_read(i, C.byref(C.c_double()), C.byref(C.c_int64()))
        fl = flops_per_state(cfgn)
        gru_tf = fl["gru"] * Bn / (ms.value / cnt.value * 1e-3) / 1e12
        out["rollout_" + wl] = dict(workload=desc, states_per_s=Bn * Hn / t, ms_per_rollout=t * 1e3, gru_stage_us=1e3 * ms.value / cnt.value,
                                    gru_stage_tflops=gru_tf, gru_frac_of_bf16_burst_peak=gru_tf / peaks["bf16_burst"],
                                    rollout_tflops=fl["total"] * Bn * Hn / t / 1e12,
                                    rollout_frac_of_bf16_sustained_peak=fl["total"] * Bn * Hn / t / 1e12 / peaks["bf16_sustained"])
        del ro, model, z0, h0, uu, nn_
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--cpu-rows", type=int, default=1024, help="start states in the CPU sample")
    ap.add_argument("--cpu-reps", type=int, default=5)
    ap.add_argument("--cpu-wm-batch", type=int, default=2, help="batch of the CPU world-model sample")
    ap.add_argument("--cpu-wm-seq", type=int, default=8, help="sequence length of the CPU world-model sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--secondary", action="store_true", help="add the auxiliary HBM-kernel / large-batch measurements (N = 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl != "ours":
        run_reference(args, rank, world, cuda=args.impl == "reference-cuda")
        return

    from dreamer_b200 import _lib as L
    hs = Harness(args)
    L.check(L.load().drm_device_check(), "device_check")
    kind, cfg, sd, B, H, desc = make_problem(args.workload, world)
    if kind == "rollout":
        line, holders = bench_rollout(hs, args, cfg, sd, B, H, desc)
    elif kind == "wm":
        line, holders = bench_wm(hs, args, cfg, sd, B, H, desc)
    else:
        line, holders = bench_iteration(hs, args, cfg, desc)
    if rank == 0 and world == 1 and args.secondary:
        try:
            line["secondary"] = secondary_metrics(hs.dev, measured_peaks(), hs.flush)
        except Exception as e:     # never lose the headline line to an auxiliary measurement
            line["secondary"] = dict(error=repr(e)[:200])
    if rank == 0 and world == 1 and not args.no_cpu_baseline and kind != "iter":
        line["cpu_baseline"] = cpu_baseline(args, kind, cfg, sd, B, H)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    hs.finish(holders)


if __name__ == "__main__":
    main()
