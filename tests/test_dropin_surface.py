"""The mirrored classes expose the reference's surface: with dreamer_b200.dropin installed, the reference's own
Dreamer.py builds them and its 97-key state_dict interchanges with ours (CPU; needs /root/reference, i.e. the build container)."""
import os
import sys

import pytest
import torch

REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "Dreamer.py")), reason="reference tree not present (GPU box)")
def test_reference_dreamer_builds_on_dropin_modules_and_state_dict_interchanges():
    from dreamer_b200 import dropin
    from oracle import weights as W
    cfg = W.small_config()
    sd = W.make_state_dict(cfg, seed=1)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in ("Dreamer", "WorldModel", "Agent", "Buffer", "SequenceModel",
                                                                          "DynamicsPredictors", "VariationalAutoEncoder", "DreamerUtils")}
    dropin.install()
    sys.path.insert(0, REF)
    try:
        import Dreamer as D   # the reference's unchanged orchestrator (Dreamer.py:13)
        d = D.Dreamer(dict(cfg), torch.device("cpu"))
        assert type(d.world_model).__module__.startswith("dreamer_b200")
        assert type(d.agent.actor).__module__.startswith("dreamer_b200")
        ours = d.state_dict()
        assert list(ours.keys()) == list(sd.keys())                          # the 97 keys, same order
        assert all(tuple(ours[k].shape) == tuple(sd[k].shape) for k in sd)
        d.load_state_dict(sd, strict=True)
        for name in ("imagine_step", "observe_step", "unroll_model", "training_step", "encoder", "sequence_model", "dynamics_predictor",
                     "reward_predictor", "continue_predictor", "decoder", "optimiser", "scalar", "horizon"):
            assert hasattr(d.world_model, name), name
        for name in ("actor", "critic", "target_critic", "S", "train_step", "compute_batched_R_lambda_returns", "update_S", "soft_update_target"):
            assert hasattr(d.agent, name), name
        assert d.buffer.capacity == cfg["buffer_size"] and d.buffer.size == 0 and d.buffer.next_idx == 0
        # patch_dreamer swaps the four hot call sites without changing their signatures
        import inspect
        names = ("dream_episodes", "warm_start_generator", "rollout_policy", "evaluate_agent")
        sigs = {n: list(inspect.signature(getattr(D.Dreamer, n)).parameters) for n in names}
        originals = {n: getattr(D.Dreamer, n) for n in names}
        try:
            dropin.patch_dreamer(D.Dreamer)
            for n in names:
                assert getattr(D.Dreamer, n) is not originals[n]
                assert list(inspect.signature(getattr(D.Dreamer, n)).parameters) == sigs[n], n
        finally:
            for n in names:
                setattr(D.Dreamer, n, originals[n])
        # no CPU path: a data call on CPU tensors must raise, not silently compute
        with pytest.raises(RuntimeError):
            d.world_model.sequence_model(torch.zeros(2, 1, 32, 32), torch.zeros(2, 1, cfg["hidden_state_dims"]), torch.zeros(2, 1, 3))
    finally:
        dropin.uninstall()
        sys.path.remove(REF)
        sys.modules.pop("Dreamer", None)
        sys.modules.update(saved)


def test_mirrors_match_reference_default_init_shapes():
    """Constructor signatures (positional order) as listed in SURVEY.md section 8b."""
    from dreamer_b200 import learners, modules
    sm = modules.SequenceModel(32, 32, 96, 3, num_layers=1, device="cpu")
    assert sm.GRU.weight_ih.shape == (288, 1027)
    a = modules.Actor(3, 32, 32, 96, 72, 72, device="cpu")
    assert float(a.mu_head.weight.abs().sum()) == 0.0                       # Agent.py:188-189 zero-initialised mu head
    c = modules.Critic(32, 32, 96, 72, 72, 255, device="cpu")
    assert c.buckets_crit.shape == (255,) and float(c.buckets_crit[127]) != 0.0
    ag = learners.Agent(3, (32, 32), 96, 72, 72, 72, 72, 255, 8e-5, (0.9, 0.999), 1e-5, 1e-4, (0.9, 0.999), 1e-5, 3e-4, 0.95, 0.99, device="cpu")
    assert not any(p.requires_grad for p in ag.target_critic.parameters()) and ag.S == 1.0


def test_hotpath_checkpoint_is_the_reference_pth_format(tmp_path):
    """HotPath.save_trained_Dreamer writes what Dreamer.save_trained_Dreamer writes (Dreamer.py:289-293): a state_dict with the
    reference's keys in its order; load_pretrained_dreamer reads it back.  When the reference tree is present, the file is
    loaded by the reference's OWN unmodified classes with strict=True, and a file the reference wrote loads here."""
    from dreamer_b200.hotpath import HotPath
    from oracle import weights as W
    cfg = W.small_config()
    hp = HotPath(cfg, "cpu")
    sd = W.make_state_dict(cfg, seed=3)
    hp.load_state_dict(sd)
    path = str(tmp_path / "agent.pth")
    hp.save_trained_Dreamer(path)
    on_disk = torch.load(path, weights_only=True)
    assert list(on_disk.keys()) == list(sd.keys())
    assert all(torch.equal(on_disk[k], sd[k]) for k in sd)
    hp2 = HotPath(cfg, "cpu")
    hp2.load_pretrained_dreamer(path)
    assert all(torch.equal(a, b) for a, b in zip(hp.state_dict().values(), hp2.state_dict().values()))
    with pytest.raises(RuntimeError):
        hp2.load_state_dict({**sd, "bogus.key": torch.zeros(1)})
    if not os.path.exists(os.path.join(REF, "Dreamer.py")):
        return
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in ("Dreamer", "WorldModel", "Agent", "Buffer", "SequenceModel",
                                                                          "DynamicsPredictors", "VariationalAutoEncoder", "DreamerUtils")}
    sys.path.insert(0, REF)
    try:
        import Dreamer as D                                  # the genuine reference classes, no drop-in
        d = D.Dreamer(dict(cfg), torch.device("cpu"))
        assert not type(d.world_model).__module__.startswith("dreamer_b200")
        d.load_pretrained_dreamer(path)                      # our file -> reference (strict)
        ref_path = str(tmp_path / "ref.pth")
        d.save_trained_Dreamer(ref_path)                     # reference file -> us
        hp3 = HotPath(cfg, "cpu")
        hp3.load_pretrained_dreamer(ref_path)
        assert all(torch.equal(a, b) for a, b in zip(hp.state_dict().values(), hp3.state_dict().values()))
    finally:
        sys.path.remove(REF)
        for k in ("Dreamer", "WorldModel", "Agent", "Buffer", "SequenceModel", "DynamicsPredictors", "VariationalAutoEncoder", "DreamerUtils"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)


def test_training_logs_npz_has_the_reference_keys(tmp_path):
    import numpy as np
    from dreamer_b200.hotpath import HotPath
    path = str(tmp_path / "training_logs.npz")
    HotPath.save_training_logs(path, [[torch.tensor(1.5), torch.tensor(2.5)], [torch.tensor(2.0), 3.0]], [torch.tensor(0.1)], [0.2, torch.tensor(0.3)], [10.0])
    z = np.load(path, allow_pickle=True)
    assert sorted(z.files) == ["actor_loss", "critic_loss", "rewards", "world_model_loss"]       # Dreamer.py:358-363
    assert float(z["actor_loss"][0]) == pytest.approx(0.1) and float(z["rewards"][0]) == 10.0 and z["world_model_loss"].shape == (2, 2)
