#!/usr/bin/env python
"""bench.py -- imagined latent states/s of the RSSM imagination rollout (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c2x16|c4] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: Dreamer.dream_episodes (Dreamer.py:143-175) for
B start states x H imagined steps with random-init weights of the reference architecture and
synthetic inputs.  Default workload = BASELINE.json configs[1]: 1024 start states
x horizon 15 at car_racer_config.yaml sizes, per GPU (weak scaling: every rank owns its own 1024
start states; the rollout needs no collective).

One JSON line is printed by rank 0.  `value` = device-timed whole-job states/s with inputs resident
in HBM; `e2e` = the same through the public host-buffer API (pinned-host inputs copied H2D and the
rewards/continues read back D2H inside the timed region); `roofline` = the dominant kernel (the fused
tcgen05 GRU stage) against the measured bf16 peak; `cpu_baseline` = the CPU port of the reference
(oracle/, stock sampler) on this box's host cores, bounded sample.
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
    # name: (B per GPU, H, config overrides, description)
    "c2": (1024, 15, {}, "imagination rollout 1024 start states x horizon 15, car_racer_config sizes (BASELINE configs[1])"),
    "c2x16": (16384, 15, {}, "imagination rollout 16384 start states x horizon 15, car_racer_config sizes"),
    "c4": (16384, 15, {"hidden_state_dims": 4096}, "imagination rollout 16384 start states x horizon 15, GRU deter 4096 (BASELINE configs[3])"),
}


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


def make_problem(workload):
    from dreamer_b200 import synthetic as W   # synthetic weights / inputs (deterministic numpy streams)
    B, H, over, desc = WORKLOADS[workload]
    cfg = dict(W.REF_CONFIG, horizon=H, **over)
    sd = W.make_state_dict(cfg, seed=0, actor_mu_zero=True)
    return cfg, sd, B, H, desc


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (the oracle port with the stock sampler),
    all host threads, on a bounded sample of the workload.  Rank 0 only."""
    if rank != 0:
        return
    from oracle import rssm as O          # the one other place bench.py may execute oracle/: the reference arm
    from dreamer_b200 import synthetic as W
    cfg, sd, B, H, desc = make_problem(args.workload)
    Bs = min(B, args.cpu_rows)
    torch.set_num_threads(os.cpu_count() or 1)
    z0, h0, _, n = W.rollout_inputs(cfg, Bs, H, seed=1234)
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            O.dream_episodes(sd, z0, h0, None, n)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    total = sum(times)
    val = Bs * H * len(times) / total
    sample = f"{Bs} of {B} start states x horizon {H}, {len(times)} timed rollouts, stock torch sampler, fp32, no_grad"
    line = dict(impl="reference", metric="imagined latent states/sec", value=val, unit="states/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", config=dict(workload=desc, l2="n/a (CPU)"),
                cpu_baseline=dict(value=val, unit="states/s", cores=torch.get_num_threads(), kind="port", sample=sample),
                e2e=dict(value=val, unit="states/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def cpu_baseline(args, cfg, sd, B, H):
    from oracle import rssm as O          # cpu_baseline leg: the oracle port timed as the reference's CPU path
    from dreamer_b200 import synthetic as W
    Bs = min(B, args.cpu_rows)
    torch.set_num_threads(os.cpu_count() or 1)
    z0, h0, _, n = W.rollout_inputs(cfg, Bs, H, seed=1234)
    times = []
    with torch.no_grad():
        for i in range(1 + args.cpu_reps):
            t0 = time.perf_counter()
            O.dream_episodes(sd, z0, h0, None, n)
            if i:
                times.append(time.perf_counter() - t0)
    val = Bs * H / statistics.median(times)
    return dict(value=val, unit="states/s", cores=torch.get_num_threads(), kind="port",
                sample=f"{Bs} of {B} start states x horizon {H}, median of {len(times)} rollouts after 1 warm-up, stock torch sampler, fp32, no_grad")


def secondary_metrics(dev, peaks, flush):
    """HBM-bound kernels against the measured HBM peak and the world-model (config 3) rates -- extra keys, N = 1 only."""
    import numpy as np
    from dreamer_b200 import ops
    from dreamer_b200 import synthetic as W
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

    # replay gather (Buffer.sample_sequences, Buffer.py:49-61) at the reference's batch 50 x sequence 50
    cap, B, L = 20000, 50, 50
    ring = [torch.randint(0, 256, (cap, 3, 64, 64), dtype=torch.uint8, device=dev), torch.rand(cap, 3, device=dev),
            torch.rand(cap, 1, device=dev), torch.ones(cap, 1, device=dev)]
    starts = torch.from_numpy(np.random.RandomState(0).randint(0, cap - L, size=B))
    t = dev_time(lambda: ops.replay_gather(*ring, starts, L))
    by = 61480 * B * L
    out["replay_gather"] = dict(workload="50 x 50 windows of 64x64x3 u8 frames -> fp32", bytes_per_launch=by, us=t * 1e6, achieved_gbs=by / t / 1e9,
                                peak_gbs=peaks["hbm"], frac=by / t / 1e9 / peaks["hbm"], note="153.7 MB per launch: short kernel, launch ramp included")
    B2 = 1024
    starts2 = torch.from_numpy(np.random.RandomState(1).randint(0, cap - L, size=B2))
    t = dev_time(lambda: ops.replay_gather(*ring, starts2, L), reps=5)
    by = 61480 * B2 * L
    out["replay_gather_large"] = dict(workload="1024 x 50 windows", bytes_per_launch=by, us=t * 1e6, achieved_gbs=by / t / 1e9, peak_gbs=peaks["hbm"],
                                      frac=by / t / 1e9 / peaks["hbm"])
    del ring
    # stand-alone categorical (softmax / unimix / sample / one-hot / ST), 4M rows of 32 classes
    n = 1 << 22
    lg = torch.randn(n, 32, device=dev); u = torch.rand(n, device=dev)
    t = dev_time(lambda: ops.categorical32(lg, u), reps=5)
    by = n * (128 + 4 + 128 + 1)
    out["categorical32"] = dict(workload="4 Mi rows x 32 classes", bytes_per_launch=by, us=t * 1e6, achieved_gbs=by / t / 1e9, peak_gbs=peaks["hbm"],
                                frac=by / t / 1e9 / peaks["hbm"])
    del lg, u
    # world model, BASELINE configs[2]: batch 16 x seq 64 of 64x64x3 frames: loss forward on the kernels, and the full training step
    cfg = dict(W.REF_CONFIG, horizon=64, sequence_length=64, batch_size=16)
    wm, _ = W.build_learners(cfg, W.make_state_dict(cfg, seed=0), dev)
    obs, act, rew, cont, uu = (x.to(dev) for x in W.sequence_inputs(cfg, 16, 64, seed=4321))
    t_f = dev_time(lambda: wm.loss_forward(obs, act, rew, cont, uniforms=uu), reps=5, warm=2)
    t_s = dev_time(lambda: wm.training_step(obs, act, rew, cont, uniforms=uu), reps=5, warm=2)
    wm.enable_cuda_graphs(warmup=1)          # the same step replayed as ONE CUDA graph (dreamer_b200/graphs.py)
    t_g = dev_time(lambda: wm.training_step(obs, act, rew, cont, uniforms=uu), reps=10, warm=3)
    out["world_model_c3"] = dict(workload="batch 16 x seq 64, 64x64x3 frames (BASELINE configs[2])", loss_forward_steps_per_s=1.0 / t_f,
                                 loss_forward_ms=t_f * 1e3, train_steps_per_s=1.0 / t_g, train_step_ms=t_g * 1e3,
                                 eager_train_step_ms=t_s * 1e3,
                                 note="training step = kernel forward + hand-scheduled BPTT (bptt.py) + fused clip/AdamW on the flat bucket, "
                                      "replayed as one CUDA graph (DESIGN.md section 6); eager_train_step_ms is the same step issued launch by launch")
    del wm
    # actor-critic update on a config-2 rollout (1024 x 15): Agent.train_step, eager and as one CUDA graph
    cfg2 = dict(W.REF_CONFIG, horizon=15)
    wm2, ag = W.build_learners(cfg2, W.make_state_dict(cfg2, seed=0), dev)
    ag.attach_world_model(wm2)               # actor gradient through the imagined states (bptt.actor_backward), as the reference's autograd
    zz = torch.nn.functional.one_hot(torch.randint(0, 32, (1024, 16, 32), device=dev), 32).float()
    hh = torch.tanh(torch.randn(1024, 16, cfg2["hidden_state_dims"], device=dev))
    rr, cc = torch.randn(1024, 15, 1, device=dev), torch.ones(1024, 15, 1, device=dev)
    mu_, sg_ = torch.randn(1024, 15, 3, device=dev) * 0.3, torch.rand(1024, 15, 3, device=dev) * 0.5 + 0.1
    aa = torch.tanh(mu_ + sg_ * torch.randn_like(mu_))
    t_a = dev_time(lambda: ag.train_step(zz, hh, rr, cc, aa, mu_, sg_), reps=5, warm=2)
    ag.enable_cuda_graphs(warmup=1)
    t_ag = dev_time(lambda: ag.train_step(zz, hh, rr, cc, aa, mu_, sg_), reps=10, warm=3)
    out["agent_step_c2"] = dict(workload="Agent.train_step on 1024 x 15 imagined states", train_step_ms=t_ag * 1e3, eager_train_step_ms=t_a * 1e3,
                                states_per_s=1024 * 15 / t_ag)
    del ag, wm2, zz, hh
    # fused optimiser tail on a flat bucket (drm_adamw_step: norm pass + update pass), 64 Mi parameters
    from dreamer_b200 import _lib as L_
    lib_ = L_.load()
    n_p = 1 << 26
    bufs = [torch.randn(n_p, device=dev) * 0.01 for _ in range(2)] + [torch.zeros(n_p, device=dev) for _ in range(2)]
    st_ = torch.zeros(8, device=dev)
    scr_ = torch.zeros(int(lib_.drm_adamw_scratch_bytes()) // 8, dtype=torch.float64, device=dev)
    t = dev_time(lambda: L_.check(lib_.drm_adamw_step(L_.ptr(bufs[0]), L_.ptr(bufs[1]), L_.ptr(bufs[2]), L_.ptr(bufs[3]), n_p, L_.ptr(st_), L_.ptr(scr_),
                                                       1e-4, 0.9, 0.999, 1e-8, 1e-6, 100.0, None, 0.0, 0, L_.stream()), "adamw"), reps=5)
    by = n_p * (4 + 16 + 12)
    out["adamw_flat"] = dict(workload="clip + AdamW on a flat bucket of 64 Mi fp32 parameters (3 launches)", bytes_per_launch=by, us=t * 1e6,
                             achieved_gbs=by / t / 1e9, peak_gbs=peaks["hbm"], frac=by / t / 1e9 / peaks["hbm"],
                             note="algorithmic bytes: 4 B/param norm pass + 16 B read + 12 B written in the update pass")
    del bufs
    # B = 1 acting path (Dreamer.rollout_policy inner loop): record -> observe_step -> act per environment step, host frame in, action out
    import numpy as np
    import time as _time
    from dreamer_b200.acting import ActingPath
    from dreamer_b200.modules import Buffer
    cfg1 = dict(W.REF_CONFIG)
    wm1, ag1 = W.build_learners(cfg1, W.make_state_dict(cfg1, seed=0), dev)
    ring1 = Buffer(4096, 50, cfg1["action_dims"], tuple(cfg1["observation_dims"]), device=dev)
    frames = np.random.default_rng(0).integers(0, 256, size=(64, 3, 64, 64)).astype(np.uint8)
    rates = {}
    for mode, use_graphs in (("graph", True), ("eager", False)):
        ap = ActingPath(wm1, ag1, ring1, use_graphs=use_graphs)
        ap.reset(frames[0]); ap.act()
        for i in range(8):
            ap.step(frames[i % 64], 0.1, 1.0)
        torch.cuda.synchronize()
        n_steps = 300
        t0 = _time.perf_counter()
        for i in range(n_steps):
            ap.step(frames[i % 64], 0.1, 1.0)
        torch.cuda.synchronize()
        rates[mode] = n_steps / (_time.perf_counter() - t0)
    out["acting_b1"] = dict(workload="one environment: pinned u8 frame in -> ring insert + observe_step + act -> action out, per step",
                            env_steps_per_s=rates["graph"], us_per_step=1e6 / rates["graph"], eager_env_steps_per_s=rates["eager"],
                            note="wall clock over 300 steps including the host read-back every step (the environment needs the action)")
    del wm1, ag1, ring1
    # one whole training iteration at the reference's own configuration (car_racer_config.yaml: batch 50 x sequence 50, horizon 30,
    # WM_epochs = AC_epochs = 2): replay sample -> world-model step (x2), replay sample -> warm start -> imagination -> agent step (x2)
    from dreamer_b200.hotpath import HotPath
    cfg_it = dict(W.REF_CONFIG, buffer_size=8192)
    hp_it = HotPath(cfg_it, dev)
    rng_it = np.random.default_rng(1)
    n_it = 4096
    hp_it.buffer.add_batch(rng_it.integers(0, 256, size=(n_it, 3, 64, 64)).astype(np.uint8), rng_it.uniform(-1, 1, (n_it, 3)).astype(np.float32),
                           rng_it.standard_normal(n_it).astype(np.float32), (rng_it.random(n_it) > 0.02).astype(np.float32))

    def iteration():
        hp_it.train_world_model()
        hp_it.train_Agent()
    t_e = dev_time(iteration, reps=3, warm=2)
    hp_it.world_model.enable_cuda_graphs(warmup=1)
    hp_it.agent.enable_cuda_graphs(warmup=1)
    t_gi = dev_time(iteration, reps=5, warm=3)
    out["training_iteration_ref"] = dict(workload="car_racer_config.yaml: batch 50 x seq 50, horizon 30, 2 world-model + 2 actor-critic epochs per iteration "
                                                  "(Dreamer.py:228-287), synthetic replay", ms_per_iteration=t_gi * 1e3, iterations_per_s=1.0 / t_gi,
                                         eager_ms_per_iteration=t_e * 1e3,
                                         note="environment stepping excluded; training steps replayed as CUDA graphs (eager = launch by launch)")
    del hp_it
    # the north star's large-batch points: 16 384 start states x horizon 15 on this one GPU (GRU stage vs the measured bf16 peak)
    import ctypes as C
    from dreamer_b200 import _lib as L
    lib = L.load()
    for wl in ("c2x16", "c4"):
        Bn, Hn, over, desc = WORKLOADS[wl]
        cfgn = dict(W.REF_CONFIG, horizon=Hn, **over)
        model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in W.make_state_dict(cfgn, seed=0, actor_mu_zero=True).items()})
        ro = ops.Rollout(model, Bn, Hn)
        z0, h0, uu, nn_ = (t.to(dev) for t in W.rollout_inputs(cfgn, Bn, Hn, seed=1234))
        t = dev_time(lambda: ro.run(z0, h0, uu, nn_, want_idx=False), reps=3, warm=2)
        lib.drm_profile_enable(1)
        ro.run(z0, h0, uu, nn_, want_idx=False)
        torch.cuda.synchronize()
        lib.drm_profile_enable(0)
        ms, cnt = C.c_double(), C.c_int64()
        lib.drm_profile_read(0, C.byref(ms), C.byref(cnt))
        for i in range(1, 8):
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rows", type=int, default=1024, help="start states in the CPU sample")
    ap.add_argument("--cpu-reps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the auxiliary HBM-kernel / world-model measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from dreamer_b200 import _lib as L
    from dreamer_b200 import ops
    from dreamer_b200.rollout import dream_episodes_host

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L.check(L.load().drm_device_check(), "device_check")

    cfg, sd, B, H, desc = make_problem(args.workload)
    from dreamer_b200 import synthetic as W
    z0, h0, u, n = W.rollout_inputs(cfg, B, H, seed=1234 + rank)      # every rank owns different start states
    model = ops.PackedRssm.from_state_dict({k: v.to(dev) for k, v in sd.items()})
    ro = ops.Rollout(model, B, H)
    z0d, h0d, ud, nd = (t.to(dev) for t in (z0, h0, u, n))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            flush.zero_(); fn()
        barrier()
        evs = []
        for _ in range(steps):
            flush.zero_()                                              # L2 flush between timed iterations (untimed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                   # max over ranks
        return float(t.item())

    lib = L.load()
    # ---- device-resident throughput ------------------------------------------------------------
    # the rollout's launch sequence is replayed as one CUDA graph (Rollout.run_graphed: two eager calls, then capture); the kernels
    # launched per rollout are counted on an eager call, since replays do not pass through the launch counter
    launches0 = lib.drm_launch_count()
    ro.run(z0d, h0d, ud, nd, want_idx=False)
    launches_per_rollout = lib.drm_launch_count() - launches0
    step = lambda: ro.run_graphed(z0d, h0d, ud, nd, want_idx=False)
    sampler = ClockSampler(local)     # every rank samples its own GPU; rank 0's record is reported
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    sampler.start()
    total_ms = timed(step, args.steps, args.warmup)
    clocks = sampler.stop()
    states = B * H * world
    value = states * args.steps / (total_ms * 1e-3)
    gpu_launches = launches_per_rollout * args.steps

    # ---- end to end through the public host-buffer API -------------------------------------------
    # the public call takes the start states (host buffers); the per-step randomness is drawn on the device inside the call,
    # as Dreamer.dream_episodes does
    pin = [t.contiguous().pin_memory() for t in (z0, h0)]
    h2d = sum(t.numel() * t.element_size() for t in pin)
    res = {}

    def e2e_step():
        out = dream_episodes_host(ro, *pin)
        res["d2h"] = sum(t.numel() * t.element_size() for t in out["host"])
    e2e_ms = timed(e2e_step, args.steps, args.warmup)
    e2e_val = states * args.steps / (e2e_ms * 1e-3)

    # ---- per-stage device times (separate profiled pass over the same workload) -----------------
    import ctypes as C
    lib.drm_profile_enable(1)
    prof_steps = max(3, min(args.steps, 10))
    for _ in range(prof_steps):
        flush.zero_(); ro.run(z0d, h0d, ud, nd, want_idx=False)      # eager: the stage events are recorded by the launch code
    torch.cuda.synchronize()
    lib.drm_profile_enable(0)
    names = ["gru", "prior_l1", "prior_l2", "prior_cat", "heads_l1", "heads_l2", "heads_out", "other"]
    stages = {}
    for i, nm in enumerate(names):
        ms, cnt = C.c_double(), C.c_int64()
        lib.drm_profile_read(i, C.byref(ms), C.byref(cnt))
        if cnt.value:
            stages[nm] = dict(ms_per_step=ms.value / prof_steps, launches_per_step=cnt.value / prof_steps, us_per_launch=1e3 * ms.value / cnt.value)
    fl = flops_per_state(cfg)
    peaks = measured_peaks()
    gru_us = stages["gru"]["us_per_launch"]
    gru_tf = fl["gru"] * B / (gru_us * 1e-6) / 1e12
    whole_tf = fl["total"] * B * H * args.steps / (total_ms * 1e-3) / 1e12   # per GPU (B is per rank)
    traffic, gru_kernel = None, "fused_gemm_kernel<EpiGru> (GRU gates, tcgen05)"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        traffic = tj and tj["dram_bytes_per_launch"]        # dram__bytes_read + write of one launch, ncu --set full (profiles/)
        gru_kernel = (tj and tj.get("kernel")) or gru_kernel   # the GRU kernel the launch code picks at this grid size
    except Exception:
        pass
    share = stages["gru"]["ms_per_step"] / sum(v["ms_per_step"] for v in stages.values())
    roofline = dict(bound="tensor", kernel=gru_kernel, achieved=gru_tf, peak=peaks["bf16_burst"],
                    unit="TFLOP/s", frac=gru_tf / peaks["bf16_burst"], traffic=traffic, peak_source=peaks["source"] + " bf16 burst",
                    share_of_step=share,
                    flops_per_launch=fl["gru"] * B, us_per_launch=gru_us,
                    whole_rollout=dict(achieved=whole_tf, peak=peaks["bf16_sustained"], frac=whole_tf / peaks["bf16_sustained"],
                                       flops_per_state=fl["total"]),
                    stages=stages)

    line = dict(metric="imagined latent states/sec", value=value, unit="states/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic", config=dict(workload=desc, start_states_per_gpu=B, horizon=H, l2="flushed (256 MiB write) between timed iterations", launch="one CUDA graph replay per rollout (captured after 2 eager calls)",
                                              parallelism=f"start states sharded over {world} rank(s), no data-path collective"),
                e2e=dict(value=e2e_val, unit="states/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=res.get("d2h", 0), ms_per_step=e2e_ms / args.steps),
                gpu_launches=int(gpu_launches), clocks=clocks, roofline=roofline)
    if rank == 0 and world == 1 and not args.no_secondary:
        try:
            line["secondary"] = secondary_metrics(dev, peaks, flush)
        except Exception as e:     # never lose the headline line to an auxiliary measurement
            line["secondary"] = dict(error=repr(e)[:200])
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, cfg, sd, B, H)
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured graphs are released before the process group (a live graph at NCCL teardown has hung the exit before)
        import gc
        ro.__dict__.pop("_graphs", None)
        gc.collect()
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
