"""Fused optimiser tail (drm_adamw_step / FlatAdamW) against torch's clip_grad_norm_ + AdamW and against the numpy oracle."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


def _net(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.LayerNorm(53), torch.nn.SiLU(), torch.nn.Linear(53, 7)).to(DEV)


@pytest.mark.parametrize("scale", [1.0, 400.0])   # 400: the global-norm clip at 100 is active
def test_flat_adamw_matches_torch_adamw_with_clipping(scale):
    from dreamer_b200.optim import FlatAdamW
    ref, mine = _net(0), _net(0)
    hp = dict(lr=4e-3, betas=(0.9, 0.95), eps=1e-7, weight_decay=1e-2)
    opt_ref = torch.optim.AdamW(ref.parameters(), **hp)
    opt = FlatAdamW(mine.parameters(), max_norm=100.0, **hp)
    g = torch.Generator(device="cuda").manual_seed(1)
    for it in range(12):
        x = torch.randn(64, 37, device=DEV, generator=g)
        for net, o in ((ref, opt_ref), (mine, opt)):
            o.zero_grad()
            (net(x).square().sum() * scale).backward()
        n_ref = torch.nn.utils.clip_grad_norm_(ref.parameters(), 100.0)
        opt_ref.step()
        opt.step()
        assert torch.allclose(opt.last_grad_norm, n_ref, rtol=1e-5)
        assert float(opt.grad.abs().max()) == 0.0          # gradients are zeroed by the same pass
        for a, b in zip(ref.parameters(), mine.parameters()):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), it
    assert float(opt.opt_state[0]) == 12
    assert (float(opt.opt_state[2]) < 1.0) == (scale > 1.0)
    # optimiser state in torch.optim.AdamW's layout, both directions
    sd = opt.state_dict()
    sd_ref = opt_ref.state_dict()
    for k in sd_ref["state"]:
        assert torch.allclose(sd["state"][k]["exp_avg"], sd_ref["state"][k]["exp_avg"], rtol=2e-4, atol=1e-7)
        assert torch.allclose(sd["state"][k]["exp_avg_sq"], sd_ref["state"][k]["exp_avg_sq"], rtol=2e-4, atol=1e-9)
        assert float(sd["state"][k]["step"]) == float(sd_ref["state"][k]["step"])
    fresh = FlatAdamW(_net(0).parameters(), **hp)
    fresh.load_state_dict(sd_ref)
    assert torch.allclose(fresh.exp_avg, opt.exp_avg, rtol=2e-4, atol=1e-7) and float(fresh.opt_state[0]) == 12


def test_flat_adamw_ema_and_nonfinite_skip():
    from dreamer_b200.optim import FlatAdamW
    critic, target = _net(3), _net(4)
    tgt0 = [p.detach().clone() for p in target.parameters()]
    opt = FlatAdamW(critic.parameters(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6, ema_params=target.parameters(), tau=0.02)
    x = torch.randn(16, 37, device=DEV)
    critic(x).sum().backward()
    opt.step()
    for t, t0, c in zip(target.parameters(), tgt0, critic.parameters()):
        assert torch.allclose(t, 0.98 * t0 + 0.02 * c, rtol=1e-6, atol=1e-7)       # Agent.py:90-94 on the UPDATED critic
    before = [p.detach().clone() for p in list(critic.parameters()) + list(target.parameters())]
    m0, step0 = opt.exp_avg.clone(), float(opt.opt_state[0])
    (critic(x).sum() * float("nan")).backward()
    opt.step()
    assert bool(opt.last_step_skipped) and float(opt.opt_state[0]) == step0
    for a, b in zip(before, list(critic.parameters()) + list(target.parameters())):
        assert torch.equal(a, b)
    assert torch.equal(m0, opt.exp_avg) and float(opt.grad.abs().max()) == 0.0


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 148 * 256 * 4 + 5, 3_000_001])
def test_adamw_step_c_abi_against_numpy_oracle(n):
    from dreamer_b200 import _lib as L
    from oracle import optim as O
    lib = L.load()
    rng = np.random.default_rng(n)
    p, g = rng.standard_normal(n).astype(np.float32), (rng.standard_normal(n) * 3).astype(np.float32)
    m, v, e = np.zeros(n, np.float32), np.zeros(n, np.float32), rng.standard_normal(n).astype(np.float32)
    dp, dg, dm, dv, de = (torch.from_numpy(a.copy()).to(DEV) for a in (p, g, m, v, e))
    state = torch.zeros(8, device=DEV)
    scratch = torch.zeros(int(lib.drm_adamw_scratch_bytes()) // 8, dtype=torch.float64, device=DEV)
    step = 0
    for it in range(3):
        L.check(lib.drm_adamw_step(L.ptr(dp), L.ptr(dg), L.ptr(dm), L.ptr(dv), n, L.ptr(state), L.ptr(scratch), 1e-3, 0.9, 0.999, 1e-8,
                                   1e-2, 100.0, L.ptr(de), 0.02, 0, L.stream()), "adamw_step")
        step, total = O.adamw_step([p], [g], [m], [v], step, 1e-3, (0.9, 0.999), 1e-8, 1e-2, 100.0)
        O.soft_update([e], [p], 0.02)
        assert abs(float(state[1]) - float(total)) <= 1e-5 * float(total)
    for d, h in ((dp, p), (dm, m), (dv, v), (de, e)):
        assert np.allclose(d.cpu().numpy(), h, rtol=3e-5, atol=3e-6)
    norm = torch.zeros(1, device=DEV)
    L.check(lib.drm_grad_norm(L.ptr(dg), n, L.ptr(scratch), L.ptr(norm), L.stream()), "grad_norm")
    assert abs(float(norm) - float(np.sqrt(np.sum(g.astype(np.float64) ** 2)))) <= 1e-5 * float(norm)
    assert lib.drm_adamw_step(ctypes.c_void_p(dp.data_ptr() + 4), L.ptr(dg), L.ptr(dm), L.ptr(dv), n, L.ptr(state),
                              L.ptr(scratch), 1e-3, 0.9, 0.999, 1e-8, 0.0, 100.0, None, 0.0, 0, L.stream()) == -2     # misaligned


def test_flat_adamw_is_graph_capturable():
    """No host synchronisation inside: forward + backward + clip + AdamW replayed as one CUDA graph matches eager steps."""
    from dreamer_b200.optim import FlatAdamW
    eager, graphed = _net(7), _net(7)
    hp = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6)
    oe, og = FlatAdamW(eager.parameters(), **hp), FlatAdamW(graphed.parameters(), **hp)
    xs = torch.randn(5, 32, 37, device=DEV)
    x_static = xs[0].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):                     # warm-up outside capture (allocator, cuBLAS handles)
        graphed(x_static).square().mean().backward()
        og.grad.zero_()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = graphed(x_static).square().mean()
        loss.backward()
        og.step()
    for i in range(5):
        x_static.copy_(xs[i])
        graph.replay()
        eager(xs[i]).square().mean().backward()
        oe.step()
    torch.cuda.synchronize()
    assert float(og.opt_state[0]) == 5
    for a, b in zip(eager.parameters(), graphed.parameters()):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_flat_adamw_follows_parameters_whose_storage_was_replaced():
    """FlatAdamW updates its flat bucket through raw pointers: when a parameter's storage is replaced after construction (e.g.
    load_state_dict(assign=True)) the next step must re-point it at its slice instead of silently training an orphan."""
    import torch
    from dreamer_b200.optim import FlatAdamW
    lin = torch.nn.Linear(8, 4, device="cuda")
    opt = FlatAdamW(lin.parameters(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=0.0)
    lin.weight.data = lin.weight.data.clone()            # storage replaced: no longer a view of the flat bucket
    w0 = lin.weight.detach().clone()
    lin(torch.randn(3, 8, device="cuda")).sum().backward()
    opt.step()
    assert not torch.equal(lin.weight.detach(), w0)      # the live parameter moved
    assert lin.weight.data_ptr() == opt.flat.data_ptr() + 4 * opt._offs[0]
