"""Gradients and post-step weights of the training steps against the REFERENCE's own autograd (tests/golden/wm_grads_small.npz,
agent_grads_small.npz: made by oracle/make_golden.py from /root/reference with autocast off, gradients snapshotted on entry to
clip_grad_norm_, i.e. the raw autograd result of WorldModel.training_step / Agent.train_step).

The fixtures hold, per parameter tensor, its L2 norm and a fixed sample of <= 1024 elements (all of them for small tensors).
Bounds: the forward runs on bf16-operand GEMMs and the backward's library GEMMs in TF32 / bf16 convs (the reference trains under
fp16 autocast), so the gradient is compared as a relative L2 error over all sampled elements and per-tensor norm ratios.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rssm as O
from oracle import weights as W

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _digest(g, prefix):
    keys = sorted({k.split("::")[1] for k in g.files if k.startswith(prefix + "::")})
    return {k: (g[f"{prefix}::{k}::idx"].astype(np.int64), g[f"{prefix}::{k}::val"].astype(np.float64), float(g[f"{prefix}::{k}::norm"])) for k in keys}


def _compare(ref, got, what, rel_bound, norm_bound):
    """ref: digest dict; got: {key: tensor}.  Relative L2 error over every sampled element + per-tensor norm ratios."""
    num = den = 0.0
    worst = (0.0, None)
    for k, (ix, val, nrm) in ref.items():
        assert k in got, (what, k, sorted(got)[:5])
        t = got[k].detach().reshape(-1).double().cpu().numpy()
        num += float(((t[ix] - val) ** 2).sum()); den += float((val ** 2).sum())
        if nrm > 1e-6 * max(r[2] for r in ref.values()):          # tensors that carry gradient at all
            ratio = float(np.linalg.norm(t)) / nrm
            if abs(ratio - 1) > worst[0]:
                worst = (abs(ratio - 1), k)
    rel = (num / max(den, 1e-300)) ** 0.5
    assert rel <= rel_bound, f"{what}: relative L2 error over the sampled elements {rel:.3g} > {rel_bound}"
    assert worst[0] <= norm_bound, f"{what}: |norm ratio - 1| = {worst[0]:.3g} for {worst[1]} > {norm_bound}"
    return rel, worst


def test_world_model_gradient_and_step_match_the_reference(golden_dir):
    from dreamer_b200 import bptt
    g = np.load(os.path.join(golden_dir, "wm_grads_small.npz"))
    cfg = json.loads(str(g["cfg"]))
    B, T, seed = int(g["B"]), int(g["T"]), int(g["seed"])
    sd = W.make_state_dict(cfg, seed=seed)
    obs, act, rew, cont, _ = (x.to(DEV) for x in W.sequence_inputs(cfg, B, T, seed=seed + 2))
    used = torch.from_numpy(g["uniforms_used"]).to(DEV)
    wm, _ = W.build_learners(cfg, sd, DEV)
    total, parts = wm.loss_forward(obs, act, rew, cont, used)
    idx, hidden_k = wm.last["scan"]["idx"], wm.last["scan"]["hidden"]
    assert np.array_equal(idx.cpu().numpy(), g["idx"])                                  # the reference's trajectory, class for class
    assert abs(float(total) - float(g["total_loss"])) <= 1e-2 * abs(float(g["total_loss"]))
    wm.optimiser.zero_grad()
    bptt.world_model_backward(wm, parts["obs_norm"], act, rew, cont, idx, hidden_k, parts)
    got = {k: p.grad for k, p in wm.named_parameters() if p.grad is not None}
    rel, worst = _compare(_digest(g, "grad"), got, "world-model gradient", rel_bound=1e-2, norm_bound=2e-2)     # measured: 3.3e-3 / 2.5e-3
    print(f"world-model gradient vs reference autograd: relative L2 {rel:.3g}, worst norm ratio off by {worst[0]:.3g} ({worst[1]})")
    # the whole step (clip(100) + AdamW, first step): the update of every element whose (clipped) gradient is clearly non-zero is
    # -lr * g / (|g| + eps) ~ -lr * sign(g)
    wm2, _ = W.build_learners(cfg, sd, DEV)
    before = {k: p.detach().clone() for k, p in wm2.named_parameters()}
    wm2.training_step(obs, act, rew, cont, uniforms=used)
    lr = float(g["lr"])
    dref, gref = _digest(g, "dw"), _digest(g, "grad")
    n_clear = n_bad = 0
    for k, (ix, val, _) in dref.items():
        if k not in gref:
            continue
        gv = gref[k][1]
        clear = np.abs(gv) > 1e-2 * np.abs(gv).max() if gv.size else np.zeros(0, bool)
        dw = (dict(wm2.named_parameters())[k].detach() - before[k]).reshape(-1).double().cpu().numpy()[ix]
        n_clear += int(clear.sum())
        n_bad += int((np.abs(dw[clear] - val[clear]) > 0.1 * lr).sum())
    assert n_clear > 1000 and n_bad <= 0.01 * n_clear, (n_bad, n_clear)


def test_agent_gradients_match_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "agent_grads_small.npz"))
    cfg = json.loads(str(g["cfg"]))
    B, H, seed = int(g["B"]), int(g["H"]), int(g["seed"])
    sd = W.make_state_dict(cfg, seed=seed)
    z0, h0, _, n = W.rollout_inputs(cfg, B, H, seed=seed + 1)
    with torch.no_grad():
        out = O.dream_episodes(sd, z0, h0, torch.from_numpy(g["uniforms_used"]), n)      # == the reference's rollout (make_golden: <= 1e-6)
    z, h, act, rew, con, mu, sg = (t.to(DEV) for t in out[:7])
    wm, ag = W.build_learners(cfg, sd, DEV)
    ag.attach_world_model(wm)                       # the reference's actor gradient flows through the imagined states (Agent.py:110)
    f = ag.losses_forward(z, h, rew, con, act, mu, sg)
    assert abs(float(f["loss_actor"]) - float(g["loss_actor"])) <= 1e-2 * max(1.0, abs(float(g["loss_actor"])))
    assert abs(float(f["loss_critic"]) - float(g["loss_critic"])) <= 1e-2 * abs(float(g["loss_critic"]))
    for opt in (ag.critic_optimiser, ag.actor_optimiser):      # gradients only: leave the parameters where they are
        opt.step = lambda *a, **k: None
    ag.soft_update_target = lambda *a, **k: None
    ag._backward_and_step(f, z, h, act, mu, sg)
    got = {"actor." + k: p.grad for k, p in ag.actor.named_parameters()}
    got.update({"critic." + k: p.grad for k, p in ag.critic.named_parameters()})
    ref = _digest(g, "grad")
    rel_a, worst_a = _compare({k: v for k, v in ref.items() if k.startswith("actor.")}, got, "actor gradient", rel_bound=1e-2, norm_bound=2e-2)    # measured 1.6e-3
    rel_c, worst_c = _compare({k: v for k, v in ref.items() if k.startswith("critic.")}, got, "critic gradient", rel_bound=1e-2, norm_bound=2e-2)  # measured 1.3e-3
    print(f"actor gradient vs reference autograd: relative L2 {rel_a:.3g}; critic: {rel_c:.3g}")
