"""Hand-scheduled back-propagation through time for WorldModel.training_step (first stage).

The loss of WorldModel.py:156-188 has ONE recurrent dependency: h_t = GRU([z_{t-1}, a_{t-1}], h_{t-1})
(SequenceModel.py:19-24); everything else -- encoder convs, posterior MLP, prior MLP, decoder, reward / continue heads, KL --
is a function of (h_t, z_t, obs_t) at one time step.  A torch autograd graph over the Python scan spends ~70 launches per time
step (3 850 per step at batch 16 x sequence 64) on 16-row tensors.  Here instead:

  1. everything non-recurrent is evaluated ONCE, batched over all B*T rows, on the teacher-forced trajectory the scan kernels
     produced (hidden states and sampled classes): prior MLP, reward / continue heads, decoder MLP and the KL terms are re-evaluated
     and differentiated by hand on this library's kernels (_heads_manual: drm_gemm_tf32, drm_twohot_ce_bwd, drm_ln_silu_bwd,
     drm_colsum); only the two conv stacks remain a torch autograd graph (cuDNN), closed from / into the hand-written part;
  2. the recurrence is walked backwards with 7 launches per step: the straight-through backward (drm_categorical32_bwd), the
     posterior MLP's input gradient (2 x drm_gemm_tf32 + drm_ln_silu_bwd), the GRU cell backward (drm_gru_bwd) and 2 x drm_gemm_tf32
     (skinny: gradient rows rounded in-kernel, K split over a cluster), chained by programmatic dependent launches;
  3. every weight gradient is a batched drm_gemm_tf32 over all B*T rows after the walk, every bias / LayerNorm-affine gradient a
     drm_colsum.

The same pieces give Agent.train_step its gradients: actor_backward (through the imagined states) and critic_backward.

Gradients are accumulated into the parameters' ``.grad`` (the flat bucket of optim.FlatAdamW).  The autograd tail
(learners._tail_world_model) stays as the reference this is tested against (tests/test_gpu_bptt.py).
"""
from __future__ import annotations

import functools

import torch
import torch.nn.functional as F

from . import dist as D
from . import ops


# The backward's GEMMs run on the tensor cores in TF32 (10-bit mantissa, what the reference's fp16 autocast keeps,
# WorldModel.py:162); False = fp32 library GEMMs whatever the backend below (the exact mode the gradient tests start from).
MATMUL_TF32 = True


class _matmul_precision:
    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(MATMUL_TF32)

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


_SIDE = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=dev)
    return _SIDE[key]


# Which GEMM runs the contractions of the two backward passes: "drm" = this library's TMA / tcgen05 kind::tf32 kernel
# (drm_gemm_tf32, csrc/gemm_tf32.cu), "torch" = the library GEMM behind torch.mm (TF32 when MATMUL_TF32).  GEMM_BATCHED covers the
# GEMMs over all B*T rows (re-evaluated pre-activations, every weight gradient, batched input gradients), GEMM_STEP the per-time-step
# GEMMs inside the two recurrence walks (16 - 1024 rows).
GEMM_BATCHED = "drm"
GEMM_STEP = "drm"
# Step GEMMs with more than 64 gradient rows (the 1024-row agent step of c5) cannot round their gradient rows inside the kernel (that is
# the swapped skinny path): they run on drm_gemm_tf32's 128-row tiles with the gradient operand read in place (truncated by the tensor
# core; the weights are pre-rounded: < 0.1 % over a 15-step walk, tests/test_gpu_bptt.py at 160 rows).  One 128 x 128 tile per SM is
# bound by the ~110 GB/s an SM ingests from L2, so these cost more than the library GEMM (13.6 us against 9.9 us at 1024 x 600 x 1800;
# agent step 1024 x 15: 4.56 ms against 4.18 ms, profiles/gemm_backend_ab.py agent); GEMM_STEP_LARGE = "torch" selects the library GEMM
# for them.
GEMM_STEP_MAX_ROWS = 64
GEMM_STEP_LARGE = "drm"
# The batched part of the world-model backward that reads (h_t, z_t) at one step -- prior MLP, reward / continue heads, decoder MLP,
# KL terms: "drm" = differentiated by hand on this library's kernels (only the conv stacks stay torch autograd / cuDNN),
# "autograd" = one torch autograd graph (the implementation the hand-written one is tested against).
HEADS_BACKWARD = "drm"


def _mm(a, b, out=None, accumulate=False, step=False):
    """out (+)= a [M, K] @ b [K, N].  Operand policy of the "drm" backend: a step GEMM multiplies gradient rows (rounded to nearest
    TF32 on the way in -- the walks compound any bias over their steps) by a weight pre-rounded with _rounded() and read in place;
    a batched GEMM reads both operands in place when they are aligned (tensor-core truncation, as a library TF32 GEMM)."""
    if (GEMM_STEP if step else GEMM_BATCHED) == "drm" and MATMUL_TF32:
        large = step and a.shape[0] > GEMM_STEP_MAX_ROWS
        if not (large and GEMM_STEP_LARGE != "drm"):
            return ops.mm(a, b, out=out, accumulate=accumulate, a_direct=large or not step, b_direct=True)
    if out is None:
        return torch.mm(a, b)
    return out.addmm_(a, b) if accumulate else torch.mm(a, b, out=out)


def _rounded(w):
    """weight operand of the step GEMMs: a TF32-rounded, aligned copy made once per backward (drm backend), else w itself"""
    return ops.pack_tf32(w) if GEMM_STEP == "drm" and MATMUL_TF32 else (w if w.is_contiguous() else w.contiguous())


def _linear(x, weight, bias):
    """x [rows, in] @ weight [out, in]^T + bias"""
    if GEMM_BATCHED == "drm" and MATMUL_TF32:
        return ops.mm_nt(x, weight, bias, a_direct=True, b_direct=True)
    return torch.addmm(bias, x, weight.t())


def _acc(p: torch.nn.Parameter, g: torch.Tensor):
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    p.grad.add_(g.view_as(p))


def _acc_colsum(p: torch.nn.Parameter, x: torch.Tensor):
    """p.grad += x.sum(0)   (bias / LayerNorm-affine gradients: drm_colsum accumulates in place, one launch)"""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    if GEMM_BATCHED == "drm" and x.is_cuda and p.grad.is_contiguous():
        ops.colsum(x, out=p.grad.view(-1), accumulate=True)
    else:
        p.grad.add_(x.sum(0).view_as(p))


def _ln_bwd(dy, a, ln):
    """Backward of SiLU(LayerNorm(a)) for a batched block: returns d(loss)/da and accumulates the affine gradients
    (d gamma = colsum(dln * xhat), d beta = colsum(dln))."""
    da, dln, dlnx = ops.ln_silu_bwd(dy, a, ln.weight, ln.bias, ln.eps, want_dln=True, want_dlnx=True)
    _acc_colsum(ln.weight, dlnx)
    _acc_colsum(ln.bias, dln)
    return da


def _conv_bias_hook(bias, g):
    """grad_output of a conv layer -> bias.grad += sum over N, H, W (the rows of the channels-last matrix [N * H * W, C])"""
    with torch.no_grad():
        C = g.shape[1]
        if g.dim() == 4 and g.is_contiguous(memory_format=torch.channels_last):
            _acc_colsum(bias, g.permute(0, 2, 3, 1).reshape(-1, C))
        else:
            if bias.grad is None:
                bias.grad = torch.zeros_like(bias)
            bias.grad.add_(g.float().sum(dim=(0, 2, 3)))
    return None


class _ImageLayer(torch.autograd.Function):
    """y = tanh(ConvTranspose2d(k 4, s 2, p 1)(x) + b) to <= 3 image channels, forward on this library's direct kernel
    (drm_convt_image_fwd: ~30 us at 1024 frames, the library conv ~340 us -- with 3 output channels a GEMM formulation is all
    overhead); backward: tanh' from the saved output, then the library's convolution_backward for d/dx, d/dweight (cuDNN, as before)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = ops.convt_image_fwd(x, weight, bias)
        ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        da = (dy * (1.0 - y * y)).to(x.dtype).contiguous(memory_format=torch.channels_last)
        dx, dw, _ = torch.ops.aten.convolution_backward(da, x, weight.to(x.dtype), None, [2, 2], [1, 1], [1, 1], True, [0, 0], 1,
                                                        [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        db = da.float().sum(dim=(0, 2, 3)) if ctx.needs_input_grad[2] else None
        return dx, (dw.float() if dw is not None else None), db


def _image_layer_ok(conv, nxt, x):
    return (isinstance(conv, torch.nn.ConvTranspose2d) and isinstance(nxt, torch.nn.Tanh) and conv.out_channels <= 3
            and conv.in_channels in (8, 16, 32, 64) and tuple(conv.kernel_size) == (4, 4) and tuple(conv.stride) == (2, 2)
            and tuple(conv.padding) == (1, 1) and tuple(conv.output_padding) == (0, 0) and conv.groups == 1 and conv.bias is not None
            and x.dtype == torch.bfloat16 and x.is_contiguous(memory_format=torch.channels_last))


def _conv_stack(seq, x):
    """Run a conv stack (nn.Sequential of Conv2d / ConvTranspose2d / SiLU) under autograd with the BIAS gradients taken out of the
    graph: each conv node is recorded while its bias does not require grad (the fused bias add of the forward stays), and a hook on
    the conv's output accumulates the column sums of grad_output into bias.grad (drm_colsum / drm_colsum_bf16) -- autograd's own
    conv-bias reduction is a ~75 us bf16 reduce kernel per layer at these sizes, 0.45 ms per world-model step."""
    if GEMM_BATCHED != "drm":
        return seq(x)
    toggled = []
    mods = list(seq)
    try:
        skip = False
        for k, m in enumerate(mods):
            if skip:                                                        # the Tanh that _ImageLayer applied
                skip = False
                continue
            if k + 1 < len(mods) and _image_layer_ok(m, mods[k + 1], x):
                x = _ImageLayer.apply(x, m.weight, m.bias)                  # the decoder's image layer + tanh: own forward kernel
                skip = True
                continue
            hook = (isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)) and m.bias is not None and m.bias.requires_grad
                    and m.out_channels % 2 == 0 and m.out_channels >= 16)
            if hook:
                m.bias.requires_grad_(False)
                toggled.append(m.bias)
            x = m(x)
            if hook and x.requires_grad:
                x.register_hook(functools.partial(_conv_bias_hook, m.bias))
    finally:
        for b in toggled:
            b.requires_grad_(True)
    return x


def _acc_mm(p: torch.nn.Parameter, a_t: torch.Tensor, b: torch.Tensor):
    """p.grad += a_t^T @ b   (a_t [rows, out], b [rows, in])"""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    if GEMM_BATCHED == "drm" and MATMUL_TF32 and p.grad.dim() == 2 and p.grad.is_contiguous():
        ops.mm_nt(a_t.t(), b.t(), out=p.grad, accumulate=True, a_direct=True, b_direct=True)
    else:
        p.grad.addmm_(a_t.t(), b)


def world_model_backward(wm, obs, act, rew, cont, idx, hidden, parts=None, marks=None, on_loss=None):
    # the batched conv graph has fixed shapes: let cuDNN pick its fastest (TF32 tensor-core) algorithms during the eager warm-up
    with torch.backends.cudnn.flags(enabled=True, benchmark=True, deterministic=False, allow_tf32=True), _matmul_precision():
        return _world_model_backward(wm, obs, act, rew, cont, idx, hidden, parts, marks, on_loss)


def _heads_autograd(wm, obs, rew, cont, Hk, z_oh, LG, parts, conv_dtype, on_loss, mark):
    """(1b) everything that reads (h_t, z_t) at one step as ONE batched torch autograd graph on leaves (HEADS_BACKWARD = "autograd":
    the implementation _heads_manual is tested against).  Returns (ok, total, loss, gH, gZ, gLG), time-major gradients."""
    B, T = obs.shape[:2]
    R, C = wm.latent_num_rows, wm.latent_num_columns
    Z = R * C
    dev = obs.device
    # ---- (1b) everything that reads (h_t, z_t) at one step: batched autograd on leaves --------------------------------------
    Hl = Hk.clone().requires_grad_(True)                                        # (B,T,D)
    Zl = z_oh.clone().requires_grad_(True)                                      # (B,T,Z)
    LGl = LG.transpose(0, 1).reshape(B, T, R, C).clone().requires_grad_(True)   # posterior logits as a leaf (KL terms)
    prior = wm.dynamics_predictor.logit_net(Hl).view(B, T, R, C)
    hz = torch.cat([Hl, Zl], -1)
    x = wm.decoder.upscaler(hz.reshape(B * T, -1)).view(B * T, wm.decoder.num_filters_start, wm.decoder.start_height, wm.decoder.start_width)
    with torch.autocast("cuda", dtype=conv_dtype, enabled=conv_dtype != torch.float32):
        dec = wm.decoder.image_builder(x.contiguous(memory_format=torch.channels_last))
    dec = dec.float().view(obs.shape)
    rl = wm.reward_predictor.logit_net(hz[:, 1:])
    cl = wm.continue_predictor.logit_generator(hz[:, 1:])
    mask = cont[:, :T - 1]
    m1 = mask.squeeze(-1)
    obs_ll = -((dec.float() - obs) ** 2).sum(dim=[-3, -2, -1])[:, 1:] * m1
    b = wm.reward_predictor.buckets_rew
    v = torch.maximum(torch.minimum(rew[:, :T - 1], b[-1]), b[0])
    lo = torch.clamp(torch.searchsorted(b, v.contiguous(), right=True) - 1, max=len(b) - 2)
    w = (v - b[lo]) / (b[lo + 1] - b[lo] + 1e-8)
    lsm = F.log_softmax(rl, -1)
    rew_ll = ((1 - w) * lsm.gather(-1, lo) + w * lsm.gather(-1, lo + 1)) * mask
    cont_ll = F.binary_cross_entropy_with_logits(cl, mask, reduction='none') * mask

    def cat_kl(lp_logits, lq_logits):
        lp = F.log_softmax(lp_logits, -1)
        lq = F.log_softmax(lq_logits, -1)
        return (lp.exp() * (lp - lq)).sum(-1).sum(-1)

    dyn_sum = (cat_kl(LGl[:, 1:].detach(), prior[:, 1:]) * m1).sum()
    rep_sum = (cat_kl(LGl[:, 1:], prior[:, 1:].detach()) * m1).sum()
    total = None
    if isinstance(parts, str):
        if parts != "reduce":
            raise ValueError("world_model_backward: parts must be a dict, None or 'reduce'")
        # forward values of this rank's sums -> one packed all-reduce -> global denominators and the global loss on every rank
        local = torch.stack([obs_ll.sum(), rew_ll.sum(), cont_ll.sum(), mask.sum(), dyn_sum, torch.full((), float(m1.numel()), device=dev)]).detach()
        total, parts = D.world_model_loss_from_sums(local, (wm.beta_pred, wm.beta_dyn, wm.beta_rep))
    if parts is None:
        denom, n_el = mask.sum() + 1e-5, float(m1.numel())
        loss_pred = (-obs_ll.sum() - rew_ll.sum() + cont_ll.sum()) / denom
        one = torch.ones((), device=dev)
        loss = wm.beta_pred * loss_pred + wm.beta_dyn * torch.maximum(one, dyn_sum / n_el) + wm.beta_rep * torch.maximum(one, rep_sum / n_el)
    else:
        denom, n_el, kl_mean = parts["denom"], parts["n_elements"], parts["kl_mean"]
        loss_pred = (-obs_ll.sum() - rew_ll.sum() + cont_ll.sum()) / denom
        live = (kl_mean > 1.0).to(loss_pred.dtype)
        const = (1.0 - live) * ((wm.beta_dyn + wm.beta_rep) / D.world())
        loss = wm.beta_pred * loss_pred + live * (wm.beta_dyn * (dyn_sum / n_el) + wm.beta_rep * (rep_sum / n_el)) + const
    if total is not None and on_loss is not None and not on_loss(total):
        return False, total
    mark("batched heads / decoder forward")
    loss.backward()          # parameter gradients of prior / decoder / heads; d/dh, d/dz, d/dlogits at every step
    with torch.no_grad():
        gH = Hl.grad.transpose(0, 1).contiguous()                               # (T,B,D)  accumulates the total d/dh_t
        gZ = Zl.grad.transpose(0, 1).contiguous()                               # (T,B,Z)  direct d/dz_t (decoder, reward, continue)
        gLG = LGl.grad.view(B, T, Z).transpose(0, 1).contiguous()               # (T,B,Z)  KL terms on the posterior logits
    return True, total, loss, gH, gZ, gLG



def _heads_manual(wm, obs, rew, cont, Hk, z_oh, LG, parts, conv_dtype, on_loss, mark):
    """(1b) on this library's kernels: prior MLP, reward / continue heads and the decoder's MLP are re-evaluated batched over all
    B*T rows and differentiated by hand (drm_twohot_ce_bwd, closed-form KL / BCE gradients, drm_ln_silu_bwd, drm_gemm_tf32); only the
    decoder's transposed-conv stack stays a torch autograd graph (cuDNN), closed from its input.  Same contract as _heads_autograd."""
    B, T = obs.shape[:2]
    R, C, Dh = wm.latent_num_rows, wm.latent_num_columns, wm.hidden_dims
    Z = R * C
    dev = obs.device
    dec_mod = wm.decoder
    u1, un, _, u2, _ = dec_mod.upscaler
    mask = cont[:, :T - 1]
    m1 = mask.squeeze(-1)
    with torch.no_grad():
        hz = torch.cat([Hk, z_oh], -1)                                                  # (B,T,D+Z)
        rows_all = hz.view(B * T, Dh + Z)
        rows_1 = hz[:, 1:].reshape(B * (T - 1), Dh + Z)
        pr = _mlp3_forward(wm.dynamics_predictor.logit_net, Hk.reshape(B * T, Dh))
        rw = _mlp3_forward(wm.reward_predictor.logit_net, rows_1)
        cn = _mlp3_forward(wm.continue_predictor.logit_generator, rows_1)
        ua1 = _linear(rows_all, u1.weight, u1.bias)
        uy1 = F.silu(F.layer_norm(ua1, (ua1.shape[-1],), un.weight, un.bias, un.eps))
        ua2 = _linear(uy1, u2.weight, u2.bias)
    # the transposed-conv stack: autograd (cuDNN) from its pre-SiLU input
    a2l = ua2.requires_grad_(True)
    x = F.silu(a2l).view(B * T, dec_mod.num_filters_start, dec_mod.start_height, dec_mod.start_width)
    with torch.autocast("cuda", dtype=conv_dtype, enabled=conv_dtype != torch.float32):
        dec = _conv_stack(dec_mod.image_builder, x.contiguous(memory_format=torch.channels_last))
    dec = dec.float().view(obs.shape)
    obs_ll = -((dec - obs) ** 2).sum(dim=[-3, -2, -1])[:, 1:] * m1
    with torch.no_grad():
        b = wm.reward_predictor.buckets_rew
        rew_rows = rew[:, :T - 1].reshape(-1, 1)
        mask_rows = mask.reshape(-1, 1)
        rew_ll = ops.twohot_ce(rw["logits"], rew_rows, b) * mask_rows
        cl = cn["logits"]
        cont_ll = F.binary_cross_entropy_with_logits(cl, mask_rows, reduction='none') * mask_rows
        lp = F.log_softmax(LG.transpose(0, 1).reshape(B, T, R, C), -1)                  # posterior (the scan's logits), batch-major
        lq = F.log_softmax(pr["logits"].view(B, T, R, C), -1)
        p_post = lp.exp()
        diff = lp - lq
        kl_row = (p_post * diff).sum(-1, keepdim=True)                                  # (B,T,R,1)
        kl = kl_row.sum(dim=[-2, -1])                                                   # (B,T)
        kl_sum = (kl[:, 1:] * m1).sum()
        total = None
        if isinstance(parts, str):
            if parts != "reduce":
                raise ValueError("world_model_backward: parts must be a dict, None or 'reduce'")
            local = torch.stack([obs_ll.detach().sum(), rew_ll.sum(), cont_ll.sum(), mask.sum(), kl_sum, torch.full((), float(m1.numel()), device=dev)])
            total, parts = D.world_model_loss_from_sums(local, (wm.beta_pred, wm.beta_dyn, wm.beta_rep))
        one = torch.ones((), device=dev)
        if parts is None:
            denom, n_el = mask.sum() + 1e-5, float(m1.numel())
            kl_mean = kl_sum / n_el
            loss_pred = (-obs_ll.detach().sum() - rew_ll.sum() + cont_ll.sum()) / denom
            loss = wm.beta_pred * loss_pred + (wm.beta_dyn + wm.beta_rep) * torch.maximum(one, kl_mean)
            live = (kl_mean > 1.0).to(torch.float32)
        else:
            denom, n_el, kl_mean = parts["denom"], parts["n_elements"], parts["kl_mean"]
            loss_pred = (-obs_ll.detach().sum() - rew_ll.sum() + cont_ll.sum()) / denom
            live = (kl_mean > 1.0).to(torch.float32)
            const = (1.0 - live) * ((wm.beta_dyn + wm.beta_rep) / D.world())
            loss = wm.beta_pred * loss_pred + live * ((wm.beta_dyn + wm.beta_rep) * (kl_sum / n_el)) + const
        c_pred = (wm.beta_pred / denom).to(torch.float32).reshape(())
        c_dyn = (live * (wm.beta_dyn / n_el)).to(torch.float32)
        c_rep = (live * (wm.beta_rep / n_el)).to(torch.float32)
    if total is not None and on_loss is not None and not on_loss(total):
        return False, total
    mark("batched heads / decoder forward")
    (c_pred * (-obs_ll.sum())).backward()                                               # conv stack parameters + d/d(its input)
    with torch.no_grad():
        # decoder MLP: Linear-LN-SiLU-Linear(-SiLU, differentiated above)
        da2 = a2l.grad
        _acc_mm(u2.weight, da2, uy1)
        _acc_colsum(u2.bias, da2)
        da1 = _ln_bwd(_mm(da2, u2.weight), ua1, un)
        _acc_mm(u1.weight, da1, rows_all)
        _acc_colsum(u1.bias, da1)
        gHZ = _mm(da1, u1.weight).view(B, T, Dh + Z)
        # reward head: - c_pred * sum(mask * twohot_ll);   continue head: + c_pred * sum(mask * BCE(cl, mask))
        d_rw = ops.twohot_ce_bwd(rw["logits"], rew_rows, b, coef=mask_rows, scale_dev=c_pred, scale=-1.0)
        d_cn = (c_pred * mask_rows) * (torch.sigmoid(cl) - mask_rows)
        gHZ[:, 1:] += (_mlp3_backward(rw, d_rw, want_dx=True) + _mlp3_backward(cn, d_cn, want_dx=True)).view(B, T - 1, Dh + Z)
        # KL terms (WorldModel.py:175-181): dyn = KL(sg(post) || prior) -> prior logits; rep = KL(post || sg(prior)) -> posterior logits
        mk = torch.zeros(B, T, 1, 1, device=dev)
        mk[:, 1:, 0, 0] = m1
        d_prior = (c_dyn * mk) * (lq.exp() - p_post)
        d_post = (c_rep * mk) * (p_post * (diff - kl_row))
        gH = gHZ[..., :Dh] + _mlp3_backward(pr, d_prior.reshape(B * T, Z), want_dx=True).view(B, T, Dh)
        gH = gH.transpose(0, 1).contiguous()
        gZ = gHZ[..., Dh:].transpose(0, 1).contiguous()
        gLG = d_post.reshape(B, T, Z).transpose(0, 1).contiguous()
    return True, total, loss, gH, gZ, gLG


def _world_model_backward(wm, obs, act, rew, cont, idx, hidden, parts=None, marks=None, on_loss=None):
    """Accumulate d(loss)/d(parameters) of WorldModel.training_step into ``.grad``; returns the (detached) loss value of the
    batched fp32 re-evaluation.

    obs (B,T,3,H,W) normalised, act (B,T,A), rew / cont (B,T,1), idx (B,T,R) classes the scan sampled, hidden (B,T,D) its h_t.
    `parts` as in learners._tail_world_model (globally reduced denominators for data-parallel shares), or the string
    "reduce": the per-rank sums of this batched re-evaluation are all-reduced here (dist.world_model_loss_from_sums) and the
    GLOBAL loss is returned -- the training step then needs no separate loss forward (the decoder / heads are evaluated once,
    in the graph that is differentiated).  `on_loss(total) -> bool` (optional) is called before any gradient work; returning
    False aborts (the reference's NaN / Inf early return, WorldModel.py:191) and the function returns (total, False).
    `marks` (a list) receives (name, cuda event) pairs at the phase boundaries (profiles/wm_step_time.py)."""

    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    B, T = obs.shape[:2]
    R, C, Dh = wm.latent_num_rows, wm.latent_num_columns, wm.hidden_dims
    if C != 32:
        raise RuntimeError("world_model_backward: 32-class latents are required (drm_categorical32_bwd)")
    dev = obs.device
    Z = R * C
    gru = wm.sequence_model.GRU
    lin1, ln1, _, lin2 = wm.encoder.latent_mapper
    n_feat = lin1.weight.shape[1] - Dh

    mark("start")
    # ---- (1a) teacher-forced trajectory and the recurrent pre-activations, batched (no autograd) ----------------------------
    with torch.no_grad():
        Hk = hidden.detach().to(torch.float32)
        z_oh = F.one_hot(idx.long(), C).to(torch.float32).view(B, T, Z)
        H_tm = Hk.transpose(0, 1).contiguous()                                  # (T,B,D)   h_t
        Hprev = torch.zeros_like(H_tm)
        Hprev[1:] = H_tm[:-1]                                                   # h_{t-1}, h_{-1} = 0 (WorldModel.py:92-95)
        X = torch.zeros(T, B, Z + act.shape[-1], device=dev)
        X[1:, :, :Z] = z_oh.transpose(0, 1)[:-1]
        X[1:, :, Z:] = act.transpose(0, 1)[:-1]                                 # x_t = [z_{t-1}, a_{t-1}], zeros at t = 0
        GI = _linear(X.view(T * B, -1), gru.weight_ih, gru.bias_ih).view(T, B, 3 * Dh)
        GH = _linear(Hprev.view(T * B, -1), gru.weight_hh, gru.bias_hh).view(T, B, 3 * Dh)
    # encoder convs: the one autograd graph that is closed later with d(loss)/d(features)
    conv_dtype = wm.__dict__.get("conv_grad_dtype", torch.bfloat16)     # the reference trains these convs under fp16 autocast
    with torch.autocast("cuda", dtype=conv_dtype, enabled=conv_dtype != torch.float32):
        feats = _conv_stack(wm.encoder.feature_extractor, obs.reshape(B * T, *obs.shape[2:]).contiguous(memory_format=torch.channels_last))
    feats = feats.float().flatten(1)                                                               # (B*T, n_feat)
    with torch.no_grad():
        X1 = torch.cat([feats.detach().view(B, T, n_feat).transpose(0, 1), H_tm], -1).contiguous()   # (T,B,n_feat+D)
        A1 = _linear(X1.view(T * B, -1), lin1.weight, lin1.bias)
        Y1 = F.silu(F.layer_norm(A1, (A1.shape[-1],), ln1.weight, ln1.bias, ln1.eps))
        LG = _linear(Y1, lin2.weight, lin2.bias).view(T, B, Z)                   # posterior logits, time-major

    mark("recurrent pre-activations + posterior MLP forward (batched)")
    if HEADS_BACKWARD == "drm":
        res = _heads_manual(wm, obs, rew, cont, Hk, z_oh, LG, parts, conv_dtype, on_loss, mark)
        if res[0] is False:
            return res[1], False
        total, loss, gH, gZ, gLG = res[1:]
    else:
        res = _heads_autograd(wm, obs, rew, cont, Hk, z_oh, LG, parts, conv_dtype, on_loss, mark)
        if res[0] is False:
            return res[1], False
        total, loss, gH, gZ, gLG = res[1:]

    mark("batched heads / decoder backward")
    with torch.no_grad():
        # ---- (2) the recurrence, backwards: 7 launches per step ---------------------------------------------------------
        dLG = torch.empty(T, B, Z, device=dev)
        dGI = torch.empty(T, B, 3 * Dh, device=dev)
        dGH = torch.empty(T, B, 3 * Dh, device=dev)
        W2 = _rounded(lin2.weight)                                              # (Z, Hn)
        W1h = _rounded(lin1.weight[:, n_feat:])                                 # (Hn, D)
        Wih_z = _rounded(gru.weight_ih[:, :Z])                                  # (3D, Z)
        Whh = _rounded(gru.weight_hh)                                           # (3D, D)
        A1_tm = A1.view(T, B, -1)
        dz_carry = None
        # The recurrent term dgh_t W_hh is only needed by the GRU cell backward of step t - 1, three kernels further down the chain
        # (straight-through, posterior MLP): it runs on a side stream into its own buffer and joins there (drm_gru_bwd_add).  In a
        # captured training step the fork / join become graph edges.
        dHrec = torch.empty(T, B, Dh, device=dev)
        main, side = torch.cuda.current_stream(), _side_stream(dev)
        joined = None
        for t in range(T - 1, -1, -1):
            ops.categorical32_bwd(LG[t], gZ[t], dz_carry, gLG[t], out=dLG[t])   # through the ST sample, + the KL term
            dA1 = ops.ln_silu_bwd(_mm(dLG[t], W2, step=True), A1_tm[t], ln1.weight, ln1.bias, ln1.eps)
            _mm(dA1, W1h, out=gH[t], accumulate=True, step=True)                                             # d/dh_t is complete up to the recurrent term
            if joined is not None:
                main.wait_event(joined)
            ops.gru_bwd(gH[t], GI[t], GH[t], Hprev[t], dGI[t], dGH[t], gH[t - 1] if t > 0 else None, accumulate=True,
                        dh_add=dHrec[t] if joined is not None else None)
            if t > 0:
                fork = torch.cuda.Event()
                fork.record(main)
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    _mm(dGH[t], Whh, out=dHrec[t - 1], step=True)
                    joined = torch.cuda.Event()
                    joined.record(side)
                dz_carry = _mm(dGI[t], Wih_z, step=True)
        if joined is not None:
            main.wait_event(joined)     # (every side-stream node is an ancestor of the capture's end)
        mark("recurrence backward (T steps)")
        # ---- (3) weight gradients: batched GEMMs over all B*T rows ----------------------------------------------------
        dGI2, dGH2, dLG2 = dGI.view(T * B, -1), dGH.view(T * B, -1), dLG.view(T * B, -1)
        _acc_mm(gru.weight_ih, dGI2, X.view(T * B, -1))
        _acc_colsum(gru.bias_ih, dGI2)
        _acc_mm(gru.weight_hh, dGH2, Hprev.view(T * B, -1))
        _acc_colsum(gru.bias_hh, dGH2)
        _acc_mm(lin2.weight, dLG2, Y1)
        _acc_colsum(lin2.bias, dLG2)
        dA1_all = _ln_bwd(_mm(dLG2, W2), A1, ln1)
        _acc_mm(lin1.weight, dA1_all, X1.view(T * B, -1))
        _acc_colsum(lin1.bias, dA1_all)
        dfeat = _mm(dA1_all, lin1.weight[:, :n_feat]).view(T, B, n_feat).transpose(0, 1).reshape(B * T, n_feat)
    mark("batched weight-gradient GEMMs")
    feats.backward(dfeat)    # encoder convs
    mark("encoder conv backward")
    return (total, True) if total is not None else loss.detach()


def _tanh_normal_log_prob(a, mu, sigma):
    """Agent.py:110-115 (same expression as learners._tanh_normal_log_prob; duplicated to keep this module import-free)."""
    a = torch.clamp(a, -1.0 + 1e-6, 1.0 - 1e-6)
    y = torch.atanh(a)
    base = -((y - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - 0.9189385332046727
    return (base - 2.0 * (0.6931471805599453 - y - F.softplus(-2.0 * y))).sum(-1)


def _mlp_fwd(x, lin_a, ln_a, lin_b, ln_b):
    """Linear-LN-SiLU x2 keeping the pre-LayerNorm activations (the backward kernels recompute the statistics from them)."""
    a1 = _linear(x, lin_a.weight, lin_a.bias)
    y1 = F.silu(F.layer_norm(a1, (a1.shape[-1],), ln_a.weight, ln_a.bias, ln_a.eps))
    a2 = _linear(y1, lin_b.weight, lin_b.bias)
    y2 = F.silu(F.layer_norm(a2, (a2.shape[-1],), ln_b.weight, ln_b.bias, ln_b.eps))
    return a1, y1, a2, y2


def actor_backward(agent, wm, z, h, act, mu, sigma, coef):
    with _matmul_precision():
        return _actor_backward(agent, wm, z, h, act, mu, sigma, coef)


def _actor_backward(agent, wm, z, h, act, mu, sigma, coef):
    """Accumulate into the ACTOR parameters' ``.grad`` the gradient of   sum_{b,t} coef[b,t] * log pi(a_t | h_t, z_t)
    (Agent.py:110-126 with coef = (-advantage / max(S,1) + nu) / N) as the reference's autograd computes it: mu_t, sigma_t
    depend on the actor parameters directly AND through the imagined state, because (h_t, z_t) were produced from the
    reparameterised earlier actions a_s = tanh(mu_s + sigma_s eps_s) by the world model (Dreamer.py:158-164; the states are
    not detached in Agent.py:110).  World-model parameters receive nothing (their gradients from this loss are discarded
    by the reference's next WorldModel.zero_grad).

    z (B,H+1,R,C) straight-through one-hot latents, h (B,H+1,D), act / mu / sigma (B,H,A) of the rollout, coef (B,H).
    Structure: batched fp32 re-evaluation of every pre-activation on the rollout's trajectory, a backward walk over the H
    steps (prior ST + MLP, GRU cell, actor MLP input gradients: ~27 launches per step on B rows), batched weight gradients."""
    B, H1 = h.shape[:2]
    H = H1 - 1
    actor = agent.actor
    l1, n1, _, l2, n2, _ = actor.base_net
    gru = wm.sequence_model.GRU
    p1, q1, _, p2, q2, _, p3 = wm.dynamics_predictor.logit_net
    Dh, A = h.shape[-1], act.shape[-1]
    if z.shape[-1] != 32:
        raise RuntimeError("actor_backward: 32-class latents are required (drm_categorical32_bwd)")
    with torch.no_grad():
        Htm = h.detach().transpose(0, 1).contiguous()                                   # (H+1,B,D)
        Ztm = z.detach().reshape(B, H1, -1).transpose(0, 1).contiguous()                # (H+1,B,Z)
        Atm = act.detach().transpose(0, 1).contiguous()                                 # (H,B,A)
        Z = Ztm.shape[-1]
        Xa = torch.cat([Htm[:H], Ztm[:H]], -1).view(H * B, Dh + Z)                      # actor input [h, z] (Agent.py:196-198)
        A1, Y1, A2, Y2 = _mlp_fwd(Xa, l1, n1, l2, n2)
        LS = _linear(Y2, actor.log_sig_head.weight, actor.log_sig_head.bias)
        mu_k = mu.detach().transpose(0, 1).reshape(H * B, A)
        sg_k = sigma.detach().transpose(0, 1).reshape(H * B, A)
        a_flat = Atm.view(H * B, A)
        EPS = ((torch.atanh(torch.clamp(a_flat, -1.0 + 1e-6, 1.0 - 1e-6)) - mu_k) / sg_k).view(H, B, A)   # the draw behind a_t
    with torch.no_grad():
        # direct d/d(mu, sigma) of the objective: closed form, one kernel (drm_tanh_normal_logp)
        gMU, gSG = ops.tanh_normal_logp(a_flat, mu_k, sg_k, coef=coef.detach().transpose(0, 1).reshape(-1), want_logp=False, want_grad=True)
        gMU, gSG = gMU.view(H, B, A), gSG.view(H, B, A)
        # world-model pre-activations of the transitions s -> s+1, s = 0..H-1 (needed for s <= H-2)
        X = torch.cat([Ztm[:H], Atm], -1).view(H * B, Z + A)
        GI = _linear(X, gru.weight_ih, gru.bias_ih).view(H, B, 3 * Dh)
        GH = _linear(Htm[:H].reshape(H * B, Dh), gru.weight_hh, gru.bias_hh).view(H, B, 3 * Dh)
        P1, _, P2, PY2 = _mlp_fwd(Htm[1:].reshape(H * B, Dh), p1, q1, p2, q2)
        LG = _linear(PY2, p3.weight, p3.bias).view(H, B, Z)                               # prior logits of states 1..H
        P1, P2 = P1.view(H, B, -1), P2.view(H, B, -1)
        A1s, A2s, LSs = A1.view(H, B, -1), A2.view(H, B, -1), LS.view(H, B, A)
        Wih_z, Wih_a, Whh = _rounded(gru.weight_ih[:, :Z]), _rounded(gru.weight_ih[:, Z:]), _rounded(gru.weight_hh)
        W1_h, W1_z, W2a = _rounded(l1.weight[:, :Dh]), _rounded(l1.weight[:, Dh:]), _rounded(l2.weight)
        Wp1, Wp2, Wp3 = _rounded(p1.weight), _rounded(p2.weight), _rounded(p3.weight)
        Whead = _rounded(torch.cat([actor.mu_head.weight, actor.log_sig_head.weight], 0))   # (2A, h2)
        dHEAD = torch.empty(H, B, 2 * A, device=h.device)                               # d/d[mu | log-sigma pre-activation]
        dY2s, dY1s = torch.empty_like(A2s), torch.empty_like(A1s)
        dgi, dgh = torch.empty(B, 3 * Dh, device=h.device), torch.empty(B, 3 * Dh, device=h.device)
        Gh = Gz = None                                                                  # gradient w.r.t. state s + 1 (zero for s + 1 = H)
        for s in range(H - 1, -1, -1):
            da = None
            ch = cz = None                                                              # carried into state s from the transition s -> s + 1
            if Gh is not None:
                dlg = ops.categorical32_bwd(LG[s], Gz)                                  # z_{s+1} = ST(prior(h_{s+1}))
                dP2 = ops.ln_silu_bwd(_mm(dlg, Wp3, step=True), P2[s], q2.weight, q2.bias, q2.eps)
                dP1 = ops.ln_silu_bwd(_mm(dP2, Wp2, step=True), P1[s], q1.weight, q1.bias, q1.eps)
                _mm(dP1, Wp1, out=Gh, accumulate=True, step=True)                                              # total d/dh_{s+1}
                ch = torch.empty_like(Gh)
                ops.gru_bwd(Gh, GI[s], GH[s], Htm[s], dgi, dgh, ch, accumulate=False)   # h_{s+1} = GRU([z_s, a_s], h_s)
                _mm(dgh, Whh, out=ch, accumulate=True, step=True)
                cz = _mm(dgi, Wih_z, step=True)
                da = _mm(dgi, Wih_a, step=True)
            # actor at state s: a_s = tanh(mu_s + sigma_s eps_s), sigma = softplus(clamp(ls, -5, 2)) + 1e-3 (Agent.py:199-209)
            ops.actor_head_bwd(gMU[s], gSG[s], da, Atm[s], EPS[s], LSs[s], dHEAD[s])
            _mm(dHEAD[s], Whead, out=dY2s[s], step=True)
            dA2 = ops.ln_silu_bwd(dY2s[s], A2s[s], n2.weight, n2.bias, n2.eps)
            _mm(dA2, W2a, out=dY1s[s], step=True)
            if s > 0:                                                                   # state 0 is an input: nothing upstream
                dA1 = ops.ln_silu_bwd(dY1s[s], A1s[s], n1.weight, n1.bias, n1.eps)
                Gh = _mm(dA1, W1_h, step=True) if ch is None else _mm(dA1, W1_h, out=ch, accumulate=True, step=True)
                Gz = _mm(dA1, W1_z, step=True) if cz is None else _mm(dA1, W1_z, out=cz, accumulate=True, step=True)
        # ---- actor weight gradients: batched GEMMs over all B*H rows ----------------------------------------------------
        dH2 = dHEAD.view(H * B, 2 * A)
        _acc_mm(actor.mu_head.weight, dH2[:, :A], Y2)
        _acc_colsum(actor.mu_head.bias, dH2[:, :A])
        _acc_mm(actor.log_sig_head.weight, dH2[:, A:], Y2)
        _acc_colsum(actor.log_sig_head.bias, dH2[:, A:])
        dA2_all = _ln_bwd(dY2s.view(H * B, -1), A2, n2)
        _acc_mm(l2.weight, dA2_all, Y1)
        _acc_colsum(l2.bias, dA2_all)
        dA1_all = _ln_bwd(dY1s.view(H * B, -1), A1, n1)
        _acc_mm(l1.weight, dA1_all, Xa)
        _acc_colsum(l1.bias, dA1_all)


def _mlp3_forward(net, x):
    """Linear-LN-SiLU-Linear-LN-SiLU-Linear (modules._mlp) on the batched rows x [rows, in], fp32 re-evaluation on this library's
    GEMM; keeps what the backward needs (pre-LayerNorm activations, layer inputs)."""
    l1, n1, _, l2, n2, _, l3 = net
    a1, y1, a2, y2 = _mlp_fwd(x, l1, n1, l2, n2)
    return dict(net=net, x=x, a1=a1, y1=y1, a2=a2, y2=y2, logits=_linear(y2, l3.weight, l3.bias))


def _mlp3_backward(c, d3, want_dx=False):
    """Parameter gradients (accumulated into ``.grad``) of the stack from d3 = d(loss)/d(logits) by the chain rule, layer by layer
    (drm_ln_silu_bwd, drm_gemm_tf32); returns d(loss)/dx when asked."""
    l1, n1, _, l2, n2, _, l3 = c["net"]
    _acc_mm(l3.weight, d3, c["y2"])
    _acc_colsum(l3.bias, d3)
    da2 = _ln_bwd(_mm(d3, l3.weight), c["a2"], n2)
    _acc_mm(l2.weight, da2, c["y1"])
    _acc_colsum(l2.bias, da2)
    da1 = _ln_bwd(_mm(da2, l2.weight), c["a1"], n1)
    _acc_mm(l1.weight, da1, c["x"])
    _acc_colsum(l1.bias, da1)
    return _mm(da1, l1.weight) if want_dx else None


def critic_backward(agent, z, h, returns, n_global):
    """Accumulate into the CRITIC parameters' ``.grad`` the gradient of  -sum_{b,t} twohot(symlog(R_bt)) . log_softmax(critic(h_bt, z_bt)) / N
    (Agent.py:128-134 as autograd differentiates it): the states and the lambda-returns are constants, so the whole pass is one batched
    MLP backward -- two-hot cross-entropy backward (drm_twohot_ce_bwd), LayerNorm-SiLU backward (drm_ln_silu_bwd) and this library's
    GEMMs; no autograd graph.  z (B,H+1,R,C), h (B,H+1,D), returns (B,H,1), n_global a 0-d device tensor (global element count)."""
    with _matmul_precision(), torch.no_grad():
        B, H1 = h.shape[:2]
        x = torch.cat([h.detach()[:, :-1], z.detach().reshape(B, H1, -1)[:, :-1]], -1).reshape(B * (H1 - 1), -1)
        inv_n = 1.0 / n_global
        c = _mlp3_forward(agent.critic.value_net, x)
        _mlp3_backward(c, ops.twohot_ce_bwd(c["logits"], returns.reshape(-1, 1), agent.critic.buckets_crit, scale_dev=inv_n, scale=-1.0,
                                            apply_symlog=True))
