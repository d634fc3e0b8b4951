"""2-GPU NCCL check of the data-parallel training path (run under torchrun on a multi-GPU box; not a pytest file):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_nccl_check.py

Each rank owns half of the sequences / start states.  Checks: (1) the sharded WorldModel loss equals the full-batch loss
evaluated by the same rank with collectives disabled; (2) after one DP training step the parameters are bit-identical on
all ranks and match a single-process full-batch step; (3) rollout shards concatenate to the full-batch rollout; (4) the same
training steps replayed as CUDA graphs (NCCL inside the capture) keep the ranks' parameters identical."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist

from dreamer_b200 import dist as D
from dreamer_b200 import rollout
from oracle import weights as W
import test_gpu_modules as tm
from test_gpu_modules import _dreamer


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = W.small_config(batch_size=8)
    sd = W.make_state_dict(cfg, seed=77)
    B, T = 8, cfg["horizon"]
    obs, act, rew, cont, u = (t.to(dev) for t in W.sequence_inputs(cfg, B, T, seed=78))
    # (1)+(2) world model
    tm.DEV = f"cuda:{local}"
    wm_dp, ag_dp = _dreamer(cfg, sd)
    wm_full, ag_full = _dreamer(cfg, sd)
    sh = lambda t, d=0: D.shard(t, d).contiguous()
    loss_dp = wm_dp.training_step(sh(obs), sh(act), sh(rew), sh(cont), uniforms=sh(u, 1))
    D._FORCE_SINGLE = True
    loss_full = wm_full.training_step(obs, act, rew, cont, uniforms=u)
    D._FORCE_SINGLE = False
    rel = abs(loss_dp.item() - loss_full.item()) / abs(loss_full.item())
    assert rel < 1e-4, ("loss parity", loss_dp.item(), loss_full.item())
    worst = 0.0
    for (k, a), (_, b) in zip(wm_dp.state_dict().items(), wm_full.state_dict().items()):
        g = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(g, a.contiguous())
        assert all(torch.equal(g[0], x) for x in g), f"{k} differs across ranks"
        worst = max(worst, (a - b).abs().max().item())
    assert worst < 2e-4, ("DP step vs full-batch step", worst)     # AdamW's first step moves every weight by ~lr = 1e-4
    # (3) rollout shards
    H = 5
    z0, h0, uu, nn_ = (t.to(dev) for t in W.rollout_inputs(cfg, B, H, seed=79))
    full = rollout.dream_episodes_modules(wm_full, ag_full, z0, h0, H, uniforms=uu, normals=nn_)
    mine = rollout.dream_episodes_modules(wm_full, ag_full, sh(z0), sh(h0), H, uniforms=sh(uu, 1), normals=sh(nn_, 1))
    lo, hi = D.shard_bounds(B)
    for a, b in zip(mine, full):
        assert torch.equal(a, b[lo:hi])
    # agent step: global means
    la, lc = ag_dp.train_step(*(sh(t) for t in (full[0], full[1], full[3], full[4], full[2], full[5], full[6])))
    D._FORCE_SINGLE = True
    la_f, lc_f = ag_full.train_step(full[0], full[1], full[3], full[4], full[2], full[5], full[6])
    D._FORCE_SINGLE = False
    assert abs(lc.item() - lc_f.item()) < 1e-4 * abs(lc_f.item()) and abs(la.item() - la_f.item()) < 1e-4 * max(1.0, abs(la_f.item()))
    # (4) the data-parallel training steps as captured CUDA graphs (NCCL all-reduces inside the capture) and the actor gradient
    # through the imagined states: replayed steps keep every rank's parameters identical
    ag_dp.attach_world_model(wm_dp)
    wm_dp.enable_cuda_graphs(warmup=1)
    ag_dp.enable_cuda_graphs(warmup=1)
    roll = (full[0], full[1], full[3], full[4], full[2], full[5], full[6])
    for it in range(4):
        wm_dp.training_step(sh(obs), sh(act), sh(rew), sh(cont), uniforms=sh(u, 1))
        ag_dp.train_step(*(sh(t) for t in roll))
    torch.cuda.synchronize()
    assert wm_dp._graphs.captured(sh(obs), sh(act), sh(rew), sh(cont), sh(u, 1))
    for mod in (wm_dp, ag_dp):
        for k, a in mod.state_dict().items():
            g = [torch.empty_like(a) for _ in range(world)]
            dist.all_gather(g, a.contiguous())
            assert all(torch.equal(g[0], x) for x in g), f"{k} differs across ranks after graph replays"
            assert torch.isfinite(a.float()).all(), k
    if rank == 0:
        print(f"dist_nccl_check ok: world={world} wm loss dp={loss_dp.item():.6f} full={loss_full.item():.6f} max |dW|={worst:.2e} "
              f"critic loss dp={lc.item():.6f} full={lc_f.item():.6f}")
    # captured graphs hold NCCL work: release them before tearing the communicator down
    wm_dp.__dict__["_graphs"] = None
    ag_dp.__dict__["_graphs"] = None
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
