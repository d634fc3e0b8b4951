"""Hand-scheduled BPTT (dreamer_b200/bptt.py and its elementwise kernels) against torch autograd."""
import pytest
import torch
import torch.nn.functional as F

from dreamer_b200 import synthetic as W

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_ln_silu_bwd_matches_autograd():
    from dreamer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    for rows, n in ((1, 7), (16, 200), (77, 256), (5, 1024)):
        a = (torch.randn(rows, n, device=DEV, generator=g) * 2 + 0.3).requires_grad_(True)
        gamma = (torch.randn(n, device=DEV, generator=g) * 0.5 + 1).requires_grad_(True)
        beta = (torch.randn(n, device=DEV, generator=g) * 0.2).requires_grad_(True)
        dy = torch.randn(rows, n, device=DEV, generator=g)
        y = F.silu(F.layer_norm(a, (n,), gamma, beta, 1e-5))
        y.backward(dy)
        da, dln = ops.ln_silu_bwd(dy, a.detach(), gamma.detach(), beta.detach(), 1e-5, want_dln=True)
        assert torch.allclose(da, a.grad, rtol=2e-4, atol=2e-5), (rows, n)
        xhat = F.layer_norm(a.detach(), (n,), None, None, 1e-5)
        assert torch.allclose((dln * xhat).sum(0), gamma.grad, rtol=2e-4, atol=2e-4)
        assert torch.allclose(dln.sum(0), beta.grad, rtol=2e-4, atol=2e-4)


def test_gru_bwd_matches_autograd_grucell():
    from dreamer_b200 import ops
    torch.manual_seed(1)
    B, Din, D = 9, 21, 40
    cell = torch.nn.GRUCell(Din, D, device=DEV)
    x = torch.randn(B, Din, device=DEV, requires_grad=True)
    h = torch.randn(B, D, device=DEV, requires_grad=True)
    dh = torch.randn(B, D, device=DEV)
    cell(x, h).backward(dh)
    with torch.no_grad():
        gi = torch.addmm(cell.bias_ih, x, cell.weight_ih.t())
        gh = torch.addmm(cell.bias_hh, h, cell.weight_hh.t())
        dgi, dgh = torch.empty_like(gi), torch.empty_like(gh)
        dhp = torch.full_like(h, 0.5)
        ops.gru_bwd(dh, gi, gh, h.detach(), dgi, dgh, dhp, accumulate=True)
        dhp.addmm_(dgh, cell.weight_hh)
        assert torch.allclose(dhp - 0.5, h.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(dgi @ cell.weight_ih, x.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(dgi.t() @ x, cell.weight_ih.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(dgh.t() @ h, cell.weight_hh.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(dgi.sum(0), cell.bias_ih.grad, rtol=1e-4, atol=1e-5)
        assert torch.allclose(dgh.sum(0), cell.bias_hh.grad, rtol=1e-4, atol=1e-5)


def _grads(wm):
    return {k: p.grad.detach().clone() for k, p in wm.named_parameters()}


@pytest.mark.parametrize("B,T", [(3, 6), (5, 9)])
def test_world_model_bptt_gradient_matches_autograd_tail(B, T):
    """Same loss, same teacher-forced classes: the hand-scheduled BPTT and the autograd tail must give the same parameter
    gradients when both walk the SAME hidden-state trajectory (the tail's own fp32 one), for every one of the 7.8 M-layout
    parameter tensors; on the scan kernels' bf16-GEMM trajectory the gradients stay close."""
    from dreamer_b200 import bptt, learners
    bptt.MATMUL_TF32 = False                    # exact comparison first; the default TF32 GEMMs are checked at the end
    cfg = W.small_config(batch_size=B, sequence_length=T, horizon=T)
    wm, _ = W.build_learners(cfg, W.make_state_dict(cfg, seed=11), DEV)
    wm.conv_grad_dtype = torch.float32          # exact comparison first; the default bf16 conv graph is checked at the end
    obs, act, rew, cont, u = (x.to(DEV) for x in W.sequence_inputs(cfg, B, T, seed=12))
    total, parts = wm.loss_forward(obs, act, rew, cont, u)
    idx, hidden_k = wm.last["scan"]["idx"], wm.last["scan"]["hidden"]
    # reference gradient: autograd over the Python scan
    wm.optimiser.zero_grad()
    tail = learners._tail_world_model(wm, parts["obs_norm"], act, rew, cont, idx, parts)
    tail.backward()
    g_ref = _grads(wm)
    # the tail's own fp32 trajectory (recomputed here without autograd)
    with torch.no_grad():
        R, C, Dh = wm.latent_num_rows, wm.latent_num_columns, wm.hidden_dims
        feats = wm.encoder.feature_extractor(parts["obs_norm"].reshape(B * T, 3, 64, 64)).flatten(1).view(B, T, -1)
        h, z, hs = torch.zeros(B, Dh, device=DEV), torch.zeros(B, R * C, device=DEV), []
        for t in range(T):
            a = act[:, t - 1] if t > 0 else torch.zeros(B, wm.action_dims, device=DEV)
            h = wm.sequence_model.GRU(torch.cat([z, a], -1), h)
            z = F.one_hot(idx[:, t].long(), C).float().reshape(B, R * C)
            hs.append(h)
        hidden_fp32 = torch.stack(hs, 1)
    assert torch.allclose(hidden_fp32, hidden_k, atol=3e-2, rtol=3e-2)          # the scan kernels' h_t line up with the tail's
    wm.optimiser.zero_grad()
    val = bptt.world_model_backward(wm, parts["obs_norm"], act, rew, cont, idx, hidden_fp32, parts)
    g_new = _grads(wm)
    assert torch.allclose(val, tail.detach(), rtol=1e-4, atol=1e-4)
    for k in g_ref:
        scale = float(g_ref[k].abs().max()) + 1e-8
        err = float((g_ref[k] - g_new[k]).abs().max()) / scale
        assert err < 2e-3, (k, err, scale)
    # the same with the batched heads / decoder-MLP / KL part as one torch autograd graph (the switch HEADS_BACKWARD = "autograd")
    assert bptt.HEADS_BACKWARD == "drm"
    bptt.HEADS_BACKWARD = "autograd"
    try:
        wm.optimiser.zero_grad()
        val_a = bptt.world_model_backward(wm, parts["obs_norm"], act, rew, cont, idx, hidden_fp32, parts)
    finally:
        bptt.HEADS_BACKWARD = "drm"
    g_a = _grads(wm)
    assert torch.allclose(val_a, val, rtol=1e-5, atol=1e-5)
    for k in g_ref:
        scale = float(g_ref[k].abs().max()) + 1e-8
        assert float((g_ref[k] - g_a[k]).abs().max()) / scale < 2e-3, k
    # every contraction on this library's TF32 GEMM (drm_gemm_tf32; convs still fp32): 5e-3 of each tensor's largest gradient
    # (measured worst 2.6e-3 -- TF32 keeps 10 mantissa bits, operands rounded to nearest)
    bptt.MATMUL_TF32 = True
    assert bptt.GEMM_BATCHED == "drm" and bptt.GEMM_STEP == "drm"
    wm.optimiser.zero_grad()
    bptt.world_model_backward(wm, parts["obs_norm"], act, rew, cont, idx, hidden_fp32, parts)
    g_t = _grads(wm)
    for k in g_ref:
        scale = float(g_ref[k].abs().max()) + 1e-8
        err = float((g_ref[k] - g_t[k]).abs().max()) / scale
        assert err < 5e-3, (k, err, scale)
    bptt.MATMUL_TF32 = False
    # on the kernels' trajectory: same gradient up to the bf16 state rounding
    wm.optimiser.zero_grad()
    bptt.world_model_backward(wm, parts["obs_norm"], act, rew, cont, idx, hidden_k, parts)
    g_k = _grads(wm)
    num = sum(float(((g_ref[k] - g_k[k]) ** 2).sum()) for k in g_ref)
    den = sum(float((g_ref[k] ** 2).sum()) for k in g_ref)
    assert (num / den) ** 0.5 < 5e-2
    # default: TF32 GEMMs (this library's) and the batched encoder / decoder conv graphs under bf16 autocast (the reference runs
    # the whole step under fp16 autocast)
    bptt.MATMUL_TF32 = True
    wm.conv_grad_dtype = torch.bfloat16
    wm.optimiser.zero_grad()
    bptt.world_model_backward(wm, parts["obs_norm"], act, rew, cont, idx, hidden_fp32, parts)
    g_b = _grads(wm)
    num = sum(float(((g_ref[k] - g_b[k]) ** 2).sum()) for k in g_ref)
    assert (num / den) ** 0.5 < 3e-2


@pytest.mark.parametrize("B", [24, 160])          # 24 rows: swapped skinny step GEMMs; 160 rows: 128-row tiles (bptt.GEMM_STEP_LARGE)
def test_actor_gradient_through_the_imagined_states_matches_autograd(B):
    """bptt.actor_backward against torch autograd over a pure-torch imagination (Dreamer.py:158-164 + Agent.py:110-126) on the
    same classes and draws: the reference's actor gradient, INCLUDING the path through the imagined states.  Also shows that the
    detached-state gradient (no world model attached) differs from it, i.e. that the through-the-world-model term is real."""
    from dreamer_b200 import bptt
    bptt.MATMUL_TF32 = False
    cfg = W.small_config(horizon=6)
    sd = W.make_state_dict(cfg, seed=31, actor_mu_zero=False)
    wm, ag = W.build_learners(cfg, sd, DEV)
    H, D, A = cfg["horizon"], cfg["hidden_state_dims"], cfg["action_dims"]
    assert bptt.GEMM_STEP == "drm" and bptt.GEMM_STEP_LARGE == "drm"
    g = torch.Generator(device="cuda").manual_seed(7)
    z = F.one_hot(torch.randint(0, 32, (B, 32), device=DEV, generator=g), 32).float().reshape(B, 1024)
    h = torch.tanh(torch.randn(B, D, device=DEV, generator=g))
    idx = torch.randint(0, 32, (H, B, 32), device=DEV, generator=g)
    eps = torch.randn(H, B, A, device=DEV, generator=g)
    coef = torch.randn(B, H, device=DEV, generator=g) / (B * H)
    actor, gru, prior = ag.actor, wm.sequence_model.GRU, wm.dynamics_predictor.logit_net
    zs, hs, acts, mus, sgs = [z], [h], [], [], []
    for t in range(H):
        base = actor.base_net(torch.cat([h, z], -1))
        mu = actor.mu_head(base)
        sg = F.softplus(torch.clamp(actor.log_sig_head(base), -5.0, 2.0)) + 1e-3
        a = torch.tanh(mu + sg * eps[t])
        h = gru(torch.cat([z, a], -1), h)
        lg = prior(h).view(B, 32, 32)
        p = 0.99 * torch.softmax(lg, -1) + 0.01 / 32
        z = (F.one_hot(idx[t], 32).float() + p - p.detach()).reshape(B, 1024)
        zs.append(z); hs.append(h); acts.append(a); mus.append(mu); sgs.append(sg)
    Z, Hh = torch.stack(zs, 1), torch.stack(hs, 1)
    Aa, MU, SG = torch.stack(acts, 1), torch.stack(mus, 1), torch.stack(sgs, 1)
    logp = bptt._tanh_normal_log_prob(Aa.detach(), MU, SG)
    for p_ in ag.actor.parameters():
        p_.grad = None
    (coef * logp).sum().backward()
    g_ref = {k: p_.grad.detach().clone() for k, p_ in ag.actor.named_parameters()}
    ag.actor_optimiser.zero_grad()
    bptt.actor_backward(ag, wm, Z.detach().view(B, H + 1, 32, 32), Hh.detach(), Aa.detach(), MU.detach(), SG.detach(), coef)
    g_new = {k: p_.grad.detach().clone() for k, p_ in ag.actor.named_parameters()}
    for k in g_ref:
        scale = float(g_ref[k].abs().max()) + 1e-12
        assert float((g_ref[k] - g_new[k]).abs().max()) / scale < 2e-3, k
    bptt.MATMUL_TF32 = True                     # default: TF32 tensor-core GEMMs
    ag.actor_optimiser.zero_grad()
    bptt.actor_backward(ag, wm, Z.detach().view(B, H + 1, 32, 32), Hh.detach(), Aa.detach(), MU.detach(), SG.detach(), coef)
    num = sum(float(((g_ref[k] - p_.grad) ** 2).sum()) for k, p_ in ag.actor.named_parameters())
    den = sum(float((g_ref[k] ** 2).sum()) for k in g_ref)
    err_tf32 = (num / den) ** 0.5
    assert err_tf32 < 3e-3
    # detached-state gradient: same objective, states treated as constants
    ag.actor_optimiser.zero_grad()
    hz = torch.cat([Hh.detach(), Z.detach()], -1)[:, :-1]
    base = actor.base_net(hz)
    lp = bptt._tanh_normal_log_prob(Aa.detach(), actor.mu_head(base), F.softplus(torch.clamp(actor.log_sig_head(base), -5.0, 2.0)) + 1e-3)
    (coef * lp).sum().backward()
    num = sum(float(((g_ref[k] - p_.grad) ** 2).sum()) for k, p_ in ag.actor.named_parameters())
    den = sum(float((g_ref[k] ** 2).sum()) for k in g_ref)
    err_detached = (num / den) ** 0.5
    assert err_detached > 3e-3 and err_detached > 4 * err_tf32      # the omitted term is real, and well above the TF32 rounding


def test_twohot_ce_bwd_matches_autograd():
    from dreamer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    N, NB = 777, 255
    logits = torch.randn(N, NB, device=DEV, generator=g) * 2
    value = torch.randn(N, 1, device=DEV, generator=g) * 30
    value[:5, 0] = torch.tensor([-1e9, 1e9, 0.0, 20.0, -20.0], device=DEV)          # outside / on the ends of the bucket range
    coef = torch.rand(N, 1, device=DEV, generator=g)
    buckets = torch.linspace(-20, 20, NB, device=DEV)
    sd = torch.tensor(0.37, device=DEV)
    for sym in (False, True):
        lg = logits.clone().requires_grad_(True)
        (ops.twohot_ce(lg.detach(), value, buckets, apply_symlog=sym) * 0).sum()      # forward kernel runs on the same inputs
        v = torch.sign(value) * torch.log1p(value.abs()) if sym else value
        v = torch.clamp(v, buckets[0], buckets[-1])
        lo = torch.clamp(torch.searchsorted(buckets, v.contiguous(), right=True) - 1, max=NB - 2)
        w = (v - buckets[lo]) / (buckets[lo + 1] - buckets[lo] + 1e-8)
        lsm = F.log_softmax(lg, -1)
        ll = (1 - w) * lsm.gather(-1, lo) + w * lsm.gather(-1, lo + 1)
        (-(coef * ll).sum() * sd * 2.0).backward()
        got = ops.twohot_ce_bwd(logits, value, buckets, coef=coef, scale_dev=sd, scale=-2.0, apply_symlog=sym)
        assert torch.allclose(got, lg.grad, rtol=1e-4, atol=1e-6)
    got = ops.twohot_ce_bwd(logits, value, buckets)
    assert torch.allclose(got.sum(-1), torch.zeros(N, device=DEV), atol=1e-5)          # twohot and softmax both sum to one


def test_critic_backward_matches_autograd():
    """bptt.critic_backward (two-hot CE backward + LayerNorm-SiLU backward + drm_gemm_tf32, no autograd graph) against torch autograd
    on the batched critic MLP (Agent.py:128-134): fp32 library GEMMs first (2e-3 per tensor), then the default TF32 kernel."""
    from dreamer_b200 import bptt
    cfg = W.small_config(horizon=5)
    wm, ag = W.build_learners(cfg, W.make_state_dict(cfg, seed=21, actor_mu_zero=False), DEV)
    B, H1, D = 40, cfg["horizon"] + 1, cfg["hidden_state_dims"]
    g = torch.Generator(device="cuda").manual_seed(9)
    z = F.one_hot(torch.randint(0, 32, (B, H1, 32), device=DEV, generator=g), 32).float()
    h = torch.tanh(torch.randn(B, H1, D, device=DEV, generator=g))
    R = torch.randn(B, H1 - 1, 1, device=DEV, generator=g) * 5
    n_glob = torch.tensor(float(B * (H1 - 1)), device=DEV)
    for p in ag.critic.parameters():
        p.grad = None
    x = torch.cat([h[:, :-1], z.reshape(B, H1, -1)[:, :-1]], -1)
    b = ag.critic.buckets_crit
    tv = torch.clamp(torch.sign(R) * torch.log1p(R.abs()), b[0], b[-1])
    lo = torch.clamp(torch.searchsorted(b, tv.contiguous(), right=True) - 1, max=len(b) - 2)
    w = (tv - b[lo]) / (b[lo + 1] - b[lo] + 1e-8)
    lsm = F.log_softmax(ag.critic.value_net(x), -1)
    ((-((1 - w) * lsm.gather(-1, lo) + w * lsm.gather(-1, lo + 1))).sum() / n_glob).backward()
    g_ref = {k: p.grad.detach().clone() for k, p in ag.critic.named_parameters()}
    den = sum(float((v ** 2).sum()) for v in g_ref.values())
    for tf32, bound in ((False, 2e-3), (True, 5e-3)):
        bptt.MATMUL_TF32 = tf32
        for p in ag.critic.parameters():
            p.grad = None
        bptt.critic_backward(ag, z, h, R, n_glob)
        num = 0.0
        for k, p in ag.critic.named_parameters():
            scale = float(g_ref[k].abs().max()) + 1e-12
            assert float((g_ref[k] - p.grad).abs().max()) / scale < bound, (k, tf32)
            num += float(((g_ref[k] - p.grad) ** 2).sum())
        assert (num / den) ** 0.5 < bound
    bptt.MATMUL_TF32 = True


def test_tanh_normal_logp_matches_torch():
    from dreamer_b200 import ops, learners
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.tanh(torch.randn(64, 15, 3, device=DEV, generator=g) * 2)
    a[0, 0] = torch.tensor([1.0, -1.0, 0.0], device=DEV)                      # clamped ends
    mu = torch.randn(64, 15, 3, device=DEV, generator=g)
    sg = torch.rand(64, 15, 3, device=DEV, generator=g) * 2 + 1e-2
    coef = torch.randn(64, 15, device=DEV, generator=g)
    mu_l, sg_l = mu.clone().requires_grad_(True), sg.clone().requires_grad_(True)
    ref = learners._tanh_normal_log_prob(a, mu_l, sg_l)
    (coef * ref).sum().backward()
    logp, gm, gs = ops.tanh_normal_logp(a, mu, sg, coef=coef, want_grad=True)
    assert torch.allclose(logp, ref.detach(), rtol=1e-4, atol=1e-3)          # atanh near +-1: ~7.25 with fp32 ulps of the clamp
    assert torch.allclose(gm, mu_l.grad, rtol=1e-4, atol=1e-4 * float(mu_l.grad.abs().max()))
    assert torch.allclose(gs, sg_l.grad, rtol=1e-4, atol=1e-4 * float(sg_l.grad.abs().max()))


def test_colsum_and_ln_affine_outputs():
    from dreamer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    for rows, n in ((1, 5), (7, 33), (1024, 1800), (2049, 200), (15360, 200), (15360, 3), (100000, 40), (1000, 3)):
        big = torch.randn(rows, n + 5, device=DEV, generator=g)
        x = big[:, 2:2 + n]                                                   # a column slice: row pitch n + 5
        ref = x.double().sum(0)
        got = ops.colsum(x)
        assert torch.allclose(got.double(), ref, rtol=1e-5, atol=1e-6 * rows + 1e-4)
        acc = torch.full((n,), 3.0, device=DEV)
        ops.colsum(x, out=acc, accumulate=True)
        assert torch.allclose(acc.double(), ref + 3.0, rtol=1e-5, atol=1e-6 * rows + 1e-4)
        assert torch.equal(ops.colsum(x), got)                                # deterministic
    # bf16 rows (conv bias gradients: grad_output viewed as [N * H * W, C]), incl. a channels-last view and a long matrix
    for rows, n in ((1, 16), (65, 32), (5000, 64), (300000, 24)):
        xb = torch.randn(rows, n, device=DEV, generator=g).to(torch.bfloat16)
        ref = xb.double().sum(0)
        got = ops.colsum(xb)
        assert torch.allclose(got.double(), ref, rtol=1e-5, atol=2e-6 * rows + 1e-4), (rows, n)
        assert torch.equal(ops.colsum(xb), got)
    g4 = torch.randn(6, 32, 5, 7, device=DEV, generator=g).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    rows_view = g4.permute(0, 2, 3, 1).reshape(-1, 32)
    assert rows_view.data_ptr() == g4.data_ptr()
    assert torch.allclose(ops.colsum(rows_view).double(), g4.double().sum(dim=(0, 2, 3)), rtol=1e-5, atol=1e-3)
    rows, n = 300, 200
    a = torch.randn(rows, n, device=DEV, generator=g) * 2
    dy = torch.randn(rows, n, device=DEV, generator=g)
    gamma, beta = torch.randn(n, device=DEV, generator=g), torch.randn(n, device=DEV, generator=g)
    da, dln, dlnx = ops.ln_silu_bwd(dy, a, gamma, beta, 1e-5, want_dln=True, want_dlnx=True)
    da0, dln0 = ops.ln_silu_bwd(dy, a, gamma, beta, 1e-5, want_dln=True)
    assert torch.equal(da, da0) and torch.equal(dln, dln0)
    xhat = F.layer_norm(a, (n,), None, None, 1e-5)
    assert torch.allclose(dlnx, dln * xhat, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("Ci,Co,N,H", [(32, 3, 5, 32), (8, 3, 3, 8), (16, 1, 2, 5)])
def test_image_layer_forward_and_backward_match_torch(Ci, Co, N, H):
    """drm_convt_image_fwd (bptt._ImageLayer) against torch's conv_transpose2d + tanh on the same bf16-rounded operands"""
    from dreamer_b200 import bptt, ops
    g = torch.Generator(device="cuda").manual_seed(Ci + N)
    x = (torch.randn(N, Ci, H, H + 1, device=DEV, generator=g)).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Ci, Co, 4, 4, device=DEV, generator=g) * 0.1
    b = torch.randn(Co, device=DEV, generator=g) * 0.1
    ref = torch.tanh(F.conv_transpose2d(x.float(), w.to(torch.bfloat16).float(), b, stride=2, padding=1))
    got = ops.convt_image_fwd(x, w, b)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, rtol=1e-4, atol=2e-5), float((got - ref).abs().max())
    # through the autograd function: gradients against torch autograd of the same expression
    xr = x.float().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    coef = torch.randn(ref.shape, device=DEV, generator=g)
    (torch.tanh(F.conv_transpose2d(xr, wr, br, stride=2, padding=1)) * coef).sum().backward()
    xl = x.clone().requires_grad_(True)
    wl, bl = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    (bptt._ImageLayer.apply(xl, wl, bl) * coef).sum().backward()
    for a_, r_ in ((xl.grad.float(), xr.grad), (wl.grad, wr.grad), (bl.grad, br.grad)):
        scale = float(r_.abs().max()) + 1e-8
        assert float((a_ - r_).abs().max()) / scale < 2e-2            # bf16 operands / bf16 gradient maps in the backward
