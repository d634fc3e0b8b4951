"""Multi-process logic of the data-parallel path on CPU (gloo, world_size 2): shard bounds, the flat gradient bucket,
and exact loss parity of the packed scalar all-reduce against the full-batch oracle loss."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dreamer_b200 import dist as D
from oracle import rssm as O
from oracle import weights as W


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(fn, world=2):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn), nprocs=world, join=True)


def _entry(rank, world, port, fn):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_partition_everything():
    for n in (1, 7, 16, 1024, 1025):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _bucket_job(rank, world):
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    frozen = torch.nn.Linear(2, 2)
    for p in frozen.parameters():
        p.requires_grad = False
    x = torch.arange(20, dtype=torch.float32).reshape(4, 5) / 10.0
    xs = D.shard(x, 0)
    (net(xs) ** 2).sum().backward()                       # per-rank share of a SUM loss
    bucket = D.FlatBucket(list(net.parameters()) + list(frozen.parameters()))
    assert bucket.numel == sum(p.numel() for p in net.parameters())
    bucket.all_reduce()
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    ref.load_state_dict(net.state_dict())
    (ref(x) ** 2).sum().backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-5), (rank, (p.grad - q.grad).abs().max())
    assert D.any_rank_flag(rank == 1, torch.device("cpu")) is True
    assert D.any_rank_flag(False, torch.device("cpu")) is False
    g = D.all_gather_cat(torch.full((3,), float(rank)))
    assert g.tolist() == [0.0] * 3 + [1.0] * 3


def test_flat_bucket_sums_gradients_gloo():
    _run(_bucket_job)


def _loss_job(rank, world):
    """Each rank evaluates the oracle's per-element loss terms on ITS sequences; the packed all-reduce must reproduce the
    full-batch WorldModel.training_step loss exactly (including max(1, global KL mean))."""
    cfg = W.small_config()
    sd = W.make_state_dict(cfg, seed=21)
    B, T = 4, 5
    obs, act, rew, cont, u = W.sequence_inputs(cfg, B, T, seed=22)
    with torch.no_grad():
        full, parts, _ = O.world_model_loss(sd, obs, act, rew, cont, u, T)
        lo, hi = D.shard_bounds(B)
        sl = slice(lo, hi)
        (prior, post, obs_ll, rew_ll, cont_bce), _ = O.unroll_model(sd, obs[sl] / 255.0 - 0.5, act[sl], rew[sl], cont[sl], u[:, sl])
    mask = cont[sl, :T - 1]
    m1 = mask.squeeze(-1)
    kl = O.categorical_kl_terms(post, prior)
    local = torch.stack([(obs_ll * m1).sum(), (rew_ll * mask).sum(), (cont_bce * mask).sum(), mask.sum(), (kl * m1).sum(),
                         torch.tensor(float(kl.numel()))])
    total, g = D.world_model_loss_from_sums(local, (1.0, 0.5, 0.1))
    assert abs(total.item() - full.item()) <= 1e-5 * max(1.0, abs(full.item())), (rank, total.item(), full.item())
    assert abs(g["kl_mean"].item() - parts["kl_mean"].item()) <= 1e-5


def test_world_model_loss_parity_across_ranks_gloo():
    _run(_loss_job)


def _rollout_job(rank, world):
    cfg = W.small_config()
    sd = W.make_state_dict(cfg, seed=5)
    z0, h0, u, n = W.rollout_inputs(cfg, 6, 3, seed=6)
    with torch.no_grad():
        full = O.dream_episodes(sd, z0, h0, u, n)
        mine = O.dream_episodes(sd, D.shard(z0), D.shard(h0), D.shard(u, 1), D.shard(n, 1))
    idx = D.all_gather_cat(mine[7])
    assert torch.equal(idx, full[7])                       # shards concatenate to the single-process result, no data-path collective needed


def test_sharded_rollout_concatenates_gloo():
    _run(_rollout_job)
