"""GPU parity of the mirrored module classes (the drop-in surface) against the oracle and the REFERENCE fixtures."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import rssm as O
from oracle import weights as W
from oracle.replay import ReplayOracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(got, ref, rel=1e-2, what="", **_):
    """North-star bound for bf16-operand quantities: max |got - ref| <= 1e-2 of the reference tensor's scale (profiles/parity_r2.md)."""
    got, ref = torch.as_tensor(got).detach().cpu().float(), torch.as_tensor(ref).detach().cpu().float()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= rel * max(scale, 1e-6), f"{what}: max abs err {err:.4g} = {err / max(scale, 1e-6):.3g} of the tensor's scale {scale:.3g}"


def _dreamer(cfg, sd):
    """WorldModel + Agent mirrors built exactly as Dreamer.__init__ does (Dreamer.py:71-118), weights from `sd`."""
    return W.build_learners(cfg, sd, DEV)


def test_module_forward_surface_matches_oracle():
    cfg = W.small_config()
    sd = W.make_state_dict(cfg, seed=13)
    wm, ag = _dreamer(cfg, sd)
    B = 7
    z0, h0, u, n = W.rollout_inputs(cfg, B, 1, seed=14)
    zd, hd = z0.to(DEV), h0.to(DEV)
    a = torch.tanh(n[0]).unsqueeze(1)
    # SequenceModel.forward
    _close(wm.sequence_model(zd, hd, a.to(DEV)), O.gru_step(sd, z0[:, 0], h0[:, 0], a[:, 0]).unsqueeze(1), what="SequenceModel")
    # DynamicsPredictor.forward / predict
    lg = wm.dynamics_predictor(hd)
    _close(lg, O.prior_logits(sd, h0[:, 0]).unsqueeze(1), what="DynamicsPredictor.forward")
    uu = O.interior_uniforms(O.unimix_probs(lg.cpu()[:, 0]), u[0], 0.0, 1e-5)
    lat, lg2 = wm.dynamics_predictor.predict(hd, uu.to(DEV))
    assert torch.equal(lat.cpu()[:, 0].argmax(-1), O.categorical_st(lg.cpu()[:, 0], uu)[1])
    # Reward / Continue
    _close(wm.reward_predictor(hd, zd), O.reward_logits(sd, h0[:, 0], z0[:, 0]).unsqueeze(1), what="RewardPredictor.forward")
    _close(wm.reward_predictor.predict(hd, zd), O.reward_predict(sd, h0[:, 0], z0[:, 0]).unsqueeze(1), what="RewardPredictor.predict")
    p, l = wm.continue_predictor(hd, zd)
    _close(l, O.continue_logit(sd, h0[:, 0], z0[:, 0]).unsqueeze(1), what="ContinuePredictor logit")
    _close(wm.continue_predictor.predict(hd, zd), torch.sigmoid(O.continue_logit(sd, h0[:, 0], z0[:, 0])).unsqueeze(1), what="ContinuePredictor.predict")
    # Actor / Critic
    act, mu, sg = ag.actor.act(hd, zd, normals=n[0].to(DEV))
    ra, rmu, rsg = O.actor_act(sd, h0[:, 0], z0[:, 0], n[0])
    _close(act, ra.unsqueeze(1), what="Actor.act"); _close(mu, rmu.unsqueeze(1), what="mu"); _close(sg, rsg.unsqueeze(1), what="sigma")
    det, _, _ = ag.actor.act(hd, zd, deterministic=True)
    _close(det, torch.tanh(rmu).unsqueeze(1), what="Actor.act deterministic")
    _close(ag.critic(hd, zd), O.critic_logits(sd, h0[:, 0], z0[:, 0]).unsqueeze(1), what="Critic.forward")
    _close(ag.critic.value(hd, zd), O.critic_value(sd, h0[:, 0], z0[:, 0]).unsqueeze(1), what="Critic.value")
    # Encoder / Decoder
    obs = (W.sequence_inputs(cfg, B, 1, seed=15)[0] / 255.0 - 0.5)
    _close(wm.encoder(hd, obs.to(DEV)), O.encoder_logits(sd, h0[:, 0], obs[:, 0]).reshape(B, 1, -1), what="Encoder.forward")
    _close(wm.decoder(hd, zd), O.decoder_forward(sd, h0[:, 0], z0[:, 0], (64, 64)).unsqueeze(1), what="Decoder.forward")
    # imagine_step / observe_step return shapes and values
    h2, z2, r, c = wm.imagine_step(hd, zd, a.to(DEV), uniforms=u[0].to(DEV))
    assert h2.shape == (B, 1, cfg["hidden_state_dims"]) and z2.shape == (B, 1, 32, 32) and r.shape == (B, 1, 1) and c.shape == (B, 1, 1)
    _close(h2, O.gru_step(sd, z0[:, 0], h0[:, 0], a[:, 0]).unsqueeze(1), what="imagine_step hidden")
    zl, hh, ll = wm.observe_step(zd, hd, a.to(DEV), obs.to(DEV), uniforms=u[0].to(DEV))
    assert zl.shape == (B, 1, 32, 32) and hh.shape == (B, 1, cfg["hidden_state_dims"]) and ll.shape == (B, 1, 32, 32)


def test_world_model_loss_matches_reference_fixture_and_training_step_updates(golden_dir):
    g = np.load(os.path.join(golden_dir, "observe_small.npz"))
    cfg = json.loads(str(g["cfg"]))
    B, T, seed = int(g["B"]), int(g["T"]), int(g["seed"])
    sd = W.make_state_dict(cfg, seed=seed)
    wm, _ = _dreamer(cfg, sd)
    obs, act, rew, cont, _ = (t.to(DEV) for t in W.sequence_inputs(cfg, B, T, seed=seed + 2))
    u = torch.from_numpy(g["uniforms_used"]).to(DEV)
    total, parts = wm.loss_forward(obs, act, rew, cont, uniforms=u)
    ref = float(g["total_loss"])
    assert abs(total.item() - ref) <= 1e-2 * abs(ref), (total.item(), ref)       # WorldModel.training_step's loss, bf16 tolerance
    assert np.array_equal(wm.last["scan"]["idx"].cpu().numpy(), g["idx"])
    before = {k: v.detach().clone() for k, v in wm.state_dict().items()}
    out = wm.training_step(obs, act, rew, cont, uniforms=u)
    assert abs(out.item() - ref) <= 1e-2 * abs(ref)
    assert abs(wm.last["tail_loss"].item() - ref) <= 2e-2 * abs(ref)             # the gradient graph evaluates the same loss
    changed = sum(int(not torch.equal(before[k], v)) for k, v in wm.state_dict().items() if v.dtype.is_floating_point and "buckets" not in k)
    assert changed >= 60                                                           # AdamW touched (nearly) every parameter tensor
    total2, _ = wm.loss_forward(obs, act, rew, cont, uniforms=u)                   # packed weights are refreshed after the step
    assert total2.item() != total.item()


def test_agent_losses_match_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "agent_small.npz"))
    cfg = json.loads(str(g["cfg"]))
    B, H, seed = int(g["B"]), int(g["H"]), int(g["seed"])
    sd = W.make_state_dict(cfg, seed=seed)
    wm, ag = _dreamer(cfg, sd)
    z0, h0, _, n = W.rollout_inputs(cfg, B, H, seed=seed + 1)
    with torch.no_grad():
        ref = O.dream_episodes(sd, z0, h0, torch.from_numpy(g["uniforms_used"]), n)
    dv = [t.to(DEV) for t in ref[:7]]
    R = ag.compute_batched_R_lambda_returns(dv[1], dv[0], dv[3], dv[4], H)
    _close(R, g["returns"], what="lambda returns")
    la, lc = ag.train_step(dv[0], dv[1], dv[3], dv[4], dv[2], dv[5], dv[6])
    assert abs(la.item() - float(g["loss_actor"])) <= 2e-2 * max(1.0, abs(float(g["loss_actor"])))
    assert abs(lc.item() - float(g["loss_critic"])) <= 1e-2 * abs(float(g["loss_critic"]))
    assert abs(float(ag.S) - float(g["S"])) <= 1e-3
    _close(ag.last["values"], g["values"], what="critic values")


def test_dream_episodes_modules_matches_reference_fixture(golden_dir):
    from dreamer_b200 import rollout
    g = np.load(os.path.join(golden_dir, "rollout_small.npz"))
    cfg = json.loads(str(g["cfg"]))
    B, H, seed = int(g["B"]), int(g["H"]), int(g["seed"])
    sd = W.make_state_dict(cfg, seed=seed)
    wm, ag = _dreamer(cfg, sd)
    z0, h0, _, n = W.rollout_inputs(cfg, B, H, seed=seed + 1)
    out = rollout.dream_episodes_modules(wm, ag, z0.to(DEV), h0.to(DEV), H, uniforms=torch.from_numpy(g["uniforms_used"]).to(DEV), normals=n.to(DEV))
    assert len(out) == 7 and out[0].shape == (B, H + 1, 32, 32) and out[3].shape == (B, H, 1)
    assert np.array_equal(out[0][:, 1:].argmax(-1).cpu().numpy().astype(np.uint8), g["idx"])
    _close(out[1], g["hidden"], what="hidden"); _close(out[3], g["rewards"], what="rewards"); _close(out[2], g["actions"], what="actions")


def test_buffer_matches_reference_ring():
    from dreamer_b200.modules import Buffer
    cap, L, B = 37, 8, 16
    buf = Buffer(cap, L, 3, (64, 64), device=DEV)
    ro = ReplayOracle(cap, L, 3, (64, 64))
    rng = np.random.Generator(np.random.PCG64(7))
    with pytest.raises(ValueError):
        buf.sample_sequences(2)
    for i in range(cap + 11):
        o = rng.integers(0, 256, size=(3, 64, 64)).astype(np.uint8)
        a = rng.uniform(-1, 1, 3).astype(np.float32)
        r = float(rng.standard_normal() * 5)
        c = float(i % 9 != 8)
        buf.add_to_buffer(o, a, r, c); ro.add(o, a, r, c)
    assert (buf.size, buf.next_idx, buf.capacity) == (ro.size, ro.next_idx, ro.capacity)
    np.random.seed(123)
    o, a, r, c, Lr = buf.sample_sequences(B)

    class Legacy:
        rs = np.random.RandomState(123)
        def randint(self, lo, hi, size=None):
            return self.rs.randint(lo, hi, size=size)
    oo, ao, rwo, co, _ = ro.gather(ro.draw_starts(B, rng=Legacy()))
    assert Lr == L and torch.equal(o.cpu(), torch.from_numpy(oo)) and torch.equal(a.cpu(), torch.from_numpy(ao))
    assert torch.allclose(r.cpu(), torch.from_numpy(rwo), atol=1e-6) and torch.equal(c.cpu(), torch.from_numpy(co))
