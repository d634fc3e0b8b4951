"""A/B of the posterior scan: persistent kernel vs three launches per step.  python profiles/scan_ab.py [B] [T]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dreamer_b200 import _lib as L, ops, synthetic as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = dict(W.REF_CONFIG, horizon=T, sequence_length=T, batch_size=B)
dev = "cuda"
lib = L.load()
sd = {k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=0).items()}
model = ops.PackedRssm.from_state_dict(sd)
vae = ops.PackedVae.from_state_dict(model, sd, (64, 64))
ws = ops.Observe(vae, B, T)
obs, act, rew, cont, u = (x.to(dev) for x in W.sequence_inputs(cfg, B, T, seed=4321))
obs = obs / 255.0 - 0.5
res = {}
for mode in (0, 1):
    for flag in (1, 0):
        L.check(lib.drm_set_option(b"persist", flag), "opt")
        out = ws.scan(obs, act, u, warm_start=bool(mode))
        torch.cuda.synchronize()
        res[(mode, flag)] = {k: v.clone() for k, v in out.items() if torch.is_tensor(v)}
        ts = []
        for i in range(13):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ws.scan(obs, act, u, warm_start=bool(mode)); b.record(); torch.cuda.synchronize()
            if i >= 3: ts.append(a.elapsed_time(b))
        print(f"mode {mode} persist={flag}: scan {B} x {T}: {statistics.median(ts):.3f} ms")
    a, b = res[(mode, 1)], res[(mode, 0)]
    for k in a:
        if a[k].dtype == torch.uint8:
            print(f"  {k}: mismatching classes {(a[k] != b[k]).float().mean().item():.2e}")
        else:
            print(f"  {k}: max |diff| {(a[k].float() - b[k].float()).abs().max().item():.3e} (scale {b[k].float().abs().max().item():.3g})")
L.check(lib.drm_set_option(b"persist", 1), "opt")
