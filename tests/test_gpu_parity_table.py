"""Achieved parity, measured: every floating-point quantity of the hot path against the REFERENCE fixtures (tests/golden, made by
oracle/make_golden.py from the reference's own modules) and the oracle, per configuration, plus the index-mismatch rates of
the free-running draws.  The table is written to gpurun_out/parity_table.json / .md (copied under profiles/ per round) and
every row is asserted against the north star's bound:

    bf16-operand quantities:  max |got - ref|  <=  1e-2 * max |ref|      (relative to the tensor's scale)
    TF32 mode (DRM_PRECISION_TF32): the same quantities, bound 1.25e-3 (three more mantissa bits: an eighth of the bf16 bound)
    fp32 kernels:             max |got - ref|  <=  1e-4 * max(1, max |ref|)

Rows whose bound is looser say why in `note`.
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
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS = []


def _row(config, quantity, got, ref, bound=1e-2, kind="bf16", note=""):
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    assert got.shape == ref.shape, (config, quantity, got.shape, ref.shape)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    denom = scale if kind in ("bf16", "tf32") else max(1.0, scale)
    rel = err / max(denom, 1e-30)
    ROWS.append(dict(config=config, quantity=quantity, kind=kind, max_abs_err=err, ref_scale=scale, rel_to_scale=rel, bound=bound, note=note))
    assert rel <= bound, f"{config} / {quantity}: max |err| {err:.4g} = {rel:.3g} of the tensor's scale {scale:.3g} (bound {bound})"


def _rate(config, quantity, mismatches, total, bound, note=""):
    rate = mismatches / max(total, 1)
    ROWS.append(dict(config=config, quantity=quantity, kind="index", mismatches=int(mismatches), draws=int(total), rate=rate, bound=bound, note=note))
    assert rate <= bound, f"{config} / {quantity}: {mismatches} of {total} draws differ ({rate:.3g} > {bound})"


@pytest.fixture(scope="module")
def ops():
    from dreamer_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module", autouse=True)
def _write_table():
    yield
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_table.json"), "w") as f:
            json.dump(ROWS, f, indent=1)
        with open(os.path.join(out, "parity_table.md"), "w") as f:
            f.write("| config | quantity | kind | max abs err | ref scale | err / scale (or mismatch rate) | bound | note |\n|---|---|---|---|---|---|---|---|\n")
            for r in ROWS:
                if r["kind"] == "index":
                    f.write(f"| {r['config']} | {r['quantity']} | index | {r['mismatches']} of {r['draws']} | | {r['rate']:.2e} | {r['bound']:.0e} | {r['note']} |\n")
                else:
                    f.write(f"| {r['config']} | {r['quantity']} | {r['kind']} | {r['max_abs_err']:.3e} | {r['ref_scale']:.3g} | {r['rel_to_scale']:.2e} | {r['bound']:.0e} | {r['note']} |\n")
    except OSError:
        pass


TF32_BOUND = 1.25e-3   # TF32 keeps 10 mantissa bits against bf16's 7: an eighth of the bf16 bound


def _model(ops, cfg, seed, precision="bf16"):
    sd = W.make_state_dict(cfg, seed=seed)
    dsd = {k: v.to(DEV) for k, v in sd.items()}
    return sd, dsd, ops.PackedRssm.from_state_dict(dsd, precision=precision)


@pytest.mark.parametrize("fixture,name", [("rollout_small.npz", "small 6x5 (reference fixture)"), ("rollout_ref_digest.npz", "ref sizes 32x15 (reference fixture)")])
@pytest.mark.parametrize("persist", [1, 0, "tf32"])
def test_rollout_vs_reference_fixture(ops, golden_dir, fixture, name, persist):
    from dreamer_b200 import _lib as L
    g = np.load(os.path.join(golden_dir, fixture))
    cfg = json.loads(str(g["cfg"]))
    B, H, seed = int(g["B"]), int(g["H"]), int(g["seed"])
    tf32 = persist == "tf32"
    persist = 0 if tf32 else persist
    bound, kind = (TF32_BOUND, "tf32") if tf32 else (1e-2, "bf16")
    sd, _, model = _model(ops, cfg, seed, "tf32" if tf32 else "bf16")
    z0, h0, _, n = W.rollout_inputs(cfg, B, H, seed=seed + 1)
    ro = ops.Rollout(model, B, H)
    lib = L.load()
    try:
        L.check(lib.drm_set_option(b"persist", persist), "opt")
        out = ro.run(z0.to(DEV), h0.to(DEV), torch.from_numpy(g["uniforms_used"]).to(DEV), n.to(DEV))
    finally:
        lib.drm_set_option(b"persist", 1)
    cfgname = f"{name}, {'TF32 mode (launch-per-stage)' if tf32 else 'persistent kernel' if persist else 'launch-per-stage'}"
    _rate(cfgname, "sampled classes, whole trajectory", int((out[7].cpu().numpy() != g["idx"]).sum()), g["idx"].size, 0.0,
          "uniforms in the middle half of the reference's CDF bin")
    for key, i in (("actions", 2), ("rewards", 3), ("continues", 4), ("mu", 5), ("sigma", 6)):
        _row(cfgname, key, out[i], torch.from_numpy(g[key]), bound, kind)
    if "hidden" in g.files:
        _row(cfgname, "hidden", out[1], torch.from_numpy(g["hidden"]), bound, kind)
    else:
        _row(cfgname, "hidden (last step)", out[1][:, -1], torch.from_numpy(g["hidden_last"]), bound, kind)


@pytest.mark.parametrize("persist", [1, 0, "tf32"])
def test_rollout_c2_teacher_forced(ops, persist):
    """BASELINE config 2 (1024 x 15): every third step re-derived by the oracle from the kernel's own previous state."""
    from dreamer_b200 import _lib as L
    cfg = dict(W.REF_CONFIG, horizon=15)
    B, H = 1024, 15
    tf32 = persist == "tf32"
    persist = 0 if tf32 else persist
    bound, kind = (TF32_BOUND, "tf32") if tf32 else (1e-2, "bf16")
    sd, _, model = _model(ops, cfg, 0, "tf32" if tf32 else "bf16")
    z0, h0, u, n = W.rollout_inputs(cfg, B, H, seed=1234)
    ro = ops.Rollout(model, B, H)
    lib = L.load()
    try:
        L.check(lib.drm_set_option(b"persist", persist), "opt")
        out = [t.cpu() for t in ro.run(z0.to(DEV), h0.to(DEV), u.to(DEV), n.to(DEV))]
    finally:
        lib.drm_set_option(b"persist", 1)
    lat, hid, act, rew, con, mu, sg, idx = out
    name = f"C2 1024x15 teacher-forced, {'TF32 mode (launch-per-stage)' if tf32 else 'persistent kernel' if persist else 'launch-per-stage'}"
    steps = list(range(0, H, 3))
    refs = {k: [] for k in ("action", "mu", "sigma", "hidden", "reward", "continue")}
    gots = {k: [] for k in refs}
    mismatch = 0
    for t in steps:
        a, m_, s_ = O.actor_act(sd, hid[:, t], lat[:, t], n[t])
        h2, z2, r, c, _, i2, _ = O.imagine_step(sd, hid[:, t], lat[:, t], act[:, t], u[t])
        mismatch += (i2 != idx[:, t].long()).sum().item()
        r_k = O.reward_predict(sd, hid[:, t + 1], lat[:, t + 1]); c_k = torch.sigmoid(O.continue_logit(sd, hid[:, t + 1], lat[:, t + 1]))
        for k, g_, r_ in (("action", act[:, t], a), ("mu", mu[:, t], m_), ("sigma", sg[:, t], s_), ("hidden", hid[:, t + 1], h2),
                          ("reward", rew[:, t], r_k), ("continue", con[:, t], c_k)):
            gots[k].append(g_); refs[k].append(r_)
    for k in refs:
        _row(name, k, torch.stack(gots[k]), torch.stack(refs[k]), bound, kind)
    _rate(name, "free-running prior draws vs the oracle's draw on the same state", mismatch, len(steps) * B * 32, 1e-3 if tf32 else 5e-3,
          "raw uniforms: a draw flips when bf16 / TF32 logits move a CDF edge across the uniform")
    oh = lat[:, 1:].sum(-1)
    assert torch.allclose(oh, torch.ones_like(oh), atol=1e-6)


def test_rollout_c4_sizes_teacher_forced(ops):
    """BASELINE config 4 sizes (GRU deter 4096), 300 rows x 3 steps."""
    cfg = dict(W.REF_CONFIG, horizon=3, hidden_state_dims=4096)
    B, H = 300, 3
    sd, _, model = _model(ops, cfg, 2)
    z0, h0, u, n = W.rollout_inputs(cfg, B, H, seed=99)
    out = [t.cpu() for t in ops.Rollout(model, B, H).run(z0.to(DEV), h0.to(DEV), u.to(DEV), n.to(DEV))]
    lat, hid, act, rew, con, mu, sg, idx = out
    name = "C4 sizes (D = 4096) 300x3 teacher-forced"
    gots, refs, mismatch = {"action": [], "hidden": []}, {"action": [], "hidden": []}, 0
    for t in range(H):
        a, _, _ = O.actor_act(sd, hid[:, t], lat[:, t], n[t])
        h2, _, _, _, _, i2, _ = O.imagine_step(sd, hid[:, t], lat[:, t], act[:, t], u[t])
        mismatch += (i2 != idx[:, t].long()).sum().item()
        gots["action"].append(act[:, t]); refs["action"].append(a); gots["hidden"].append(hid[:, t + 1]); refs["hidden"].append(h2)
    for k in gots:
        _row(name, k, torch.stack(gots[k]), torch.stack(refs[k]))
    _rate(name, "free-running prior draws", mismatch, B * H * 32, 5e-3, "raw uniforms")


@pytest.mark.parametrize("fixture,name", [("observe_small.npz", "small 3x6 observe (reference fixture)"), ("observe_ref_digest.npz", "ref sizes 2x4 observe (reference fixture)")])
def test_observe_vs_reference_fixture(ops, golden_dir, fixture, name):
    g = np.load(os.path.join(golden_dir, fixture))
    cfg = json.loads(str(g["cfg"]))
    B, T, seed = int(g["B"]), int(g["T"]), int(g["seed"])
    sd, dsd, model = _model(ops, cfg, seed)
    vae = ops.PackedVae.from_state_dict(model, dsd, tuple(cfg["observation_dims"]))
    obs, act, rew, cont, _ = W.sequence_inputs(cfg, B, T, seed=seed + 2)
    obs_n = obs / 255.0 - 0.5
    ws = ops.Observe(vae, B, T)
    sc = ws.scan(obs_n.to(DEV), act.to(DEV), torch.from_numpy(g["uniforms_used"]).to(DEV))
    _rate(name, "posterior classes, whole trajectory", int((sc["idx"].cpu().numpy() != g["idx"]).sum()), g["idx"].size, 0.0,
          "uniforms in the middle half of the reference's CDF bin")
    _row(name, "hidden", sc["hidden"], torch.from_numpy(g["hidden"]))
    _row(name, "posterior logits", sc["logits"][:, 1:], torch.from_numpy(g["post_logits"]))
    hd = ws.heads()
    _row(name, "prior logits", hd["prior_logits"][:, 1:], torch.from_numpy(g["prior_logits"]))
    _row(name, "observation log-likelihood (-SSE)", ops.neg_sse_rows(hd["dec_mu"], obs_n.to(DEV))[:, 1:], torch.from_numpy(g["obs_ll"]))
    buckets = sd["world_model.reward_predictor.buckets_rew"].to(DEV)
    _row(name, "reward two-hot log-likelihood", ops.twohot_ce(hd["reward_logits"], rew[:, :T - 1].to(DEV), buckets), torch.from_numpy(g["rew_ll"]))
    bce = torch.nn.functional.binary_cross_entropy_with_logits(hd["cont_logit"].cpu(), cont[:, :T - 1], reduction="none")
    _row(name, "continue BCE", bce, torch.from_numpy(g["cont_bce"]))
    kl = ops.categorical32_kl(sc["logits"][:, 1:], hd["prior_logits"][:, 1:]).cpu()
    _row(name, "KL (masked mean)", (kl * cont[:, :T - 1, 0]).mean().reshape(1), torch.tensor([float(g["kl_mean"])]))


def test_observe_c3_teacher_forced(ops):
    """BASELINE config 3 (batch 16 x seq 64)."""
    cfg = dict(W.REF_CONFIG, horizon=64, sequence_length=64, batch_size=16)
    B, T = 16, 64
    sd, dsd, model = _model(ops, cfg, 0)
    vae = ops.PackedVae.from_state_dict(model, dsd, tuple(cfg["observation_dims"]))
    obs, act, rew, cont, u = W.sequence_inputs(cfg, B, T, seed=4321)
    obs_n = obs / 255.0 - 0.5
    ws = ops.Observe(vae, B, T)
    sc = {k: (v.cpu() if v is not None else None) for k, v in ws.scan(obs_n.to(DEV), act.to(DEV), u.to(DEV)).items()}
    name = "C3 16x64 observe teacher-forced"
    gh, rh, gl, rl, mism = [], [], [], [], 0
    steps = (0, 1, 17, 40, 63)
    for t in steps:
        zp = sc["latent"][:, t - 1] if t > 0 else torch.zeros(B, 32, 32)
        hp = sc["hidden"][:, t - 1] if t > 0 else torch.zeros(B, cfg["hidden_state_dims"])
        ap = act[:, t - 1] if t > 0 else torch.zeros(B, 3)
        z2, h2, lg, idx, _ = O.observe_step(sd, zp, hp, ap, obs_n[:, t], u[t])
        gh.append(sc["hidden"][:, t]); rh.append(h2); gl.append(sc["logits"][:, t]); rl.append(lg)
        mism += (idx != sc["idx"][:, t].long()).sum().item()
    _row(name, "hidden", torch.stack(gh), torch.stack(rh))
    _row(name, "posterior logits", torch.stack(gl), torch.stack(rl))
    _rate(name, "free-running posterior draws", mism, len(steps) * B * 32, 1e-2, "raw uniforms")
    hd = ws.heads(reward=False, cont=False)
    t = 21
    _row(name, "decoder mean", hd["dec_mu"][:, t], O.decoder_forward(sd, sc["hidden"][:, t], sc["latent"][:, t], (64, 64)))
    _row(name, "prior logits", hd["prior_logits"][:, t], O.prior_logits(sd, sc["hidden"][:, t]))


def test_fp32_kernels(ops, golden_dir):
    """lambda-returns, two-hot CE and the bucket read-out are fp32 end to end: 1e-4."""
    g = np.load(os.path.join(golden_dir, "agent_small.npz"))
    cfg = json.loads(str(g["cfg"]))
    sd = W.make_state_dict(cfg, seed=int(g["seed"]))
    buckets = sd["agent.critic.buckets_crit"].to(DEV)
    vals = torch.from_numpy(g["twohot_vals"]).to(DEV)
    lg = torch.randn(vals.shape[0], buckets.numel(), generator=torch.Generator().manual_seed(0)).to(DEV)
    ref = (torch.from_numpy(g["twohot"]) * torch.log_softmax(lg.cpu(), -1)).sum(-1, keepdim=True)
    _row("agent fixture", "two-hot CE against the reference's to_twohot", ops.twohot_ce(lg, vals, buckets, apply_symlog=False), ref, bound=1e-4, kind="fp32")
    gen = torch.Generator().manual_seed(1)
    B, H = 257, 15
    rew, cont, val = torch.randn(B, H, generator=gen), torch.rand(B, H, generator=gen), torch.randn(B, H + 1, generator=gen)
    _row("random 257x15", "lambda-returns", ops.lambda_return(rew.to(DEV), cont.to(DEV), val.to(DEV), 0.99, 0.95),
         O.lambda_returns(rew.unsqueeze(-1), cont.unsqueeze(-1), val.unsqueeze(-1), 0.99, 0.95).squeeze(-1), bound=1e-4, kind="fp32")
