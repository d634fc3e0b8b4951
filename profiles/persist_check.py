"""Bring-up / A-B driver of the persistent rollout kernel: python profiles/persist_check.py [case ...]

Each case runs in its own process (a trapped launch kills the CUDA context): the rollout is run with option persist = 1 and 0,
compared, timed (CUDA events, L2 flushed between iterations), and on a failed launch the kernel's timeout records are printed.
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    "small8": ("small", 8, 4), "small200": ("small", 200, 4), "small300": ("small", 300, 5), "ref130": ("ref", 130, 3),
    "ref512": ("ref", 512, 15), "ref1024": ("ref", 1024, 15), "ref32": ("ref", 32, 15),
}


def child(name):
    import torch
    from dreamer_b200 import _lib as L, ops, synthetic as W
    kind, B, H = CASES[name]
    cfg = W.small_config() if kind == "small" else dict(W.REF_CONFIG, horizon=H)
    lib = L.load()
    dev = "cuda"
    sd = {k: v.to(dev) for k, v in W.make_state_dict(cfg, seed=5).items()}
    model = ops.PackedRssm.from_state_dict(sd)
    ro = ops.Rollout(model, B, H)
    z0, h0, u, n = (t.to(dev) for t in W.rollout_inputs(cfg, B, H, seed=6))
    res = {"case": name, "B": B, "H": H}
    L.check(lib.drm_set_option(b"persist", 0), "opt")
    base = ro.run(z0, h0, u, n)
    torch.cuda.synchronize()
    L.check(lib.drm_set_option(b"persist", 1), "opt")
    res["info"] = ro.info()
    try:
        alt = ro.run(z0, h0, u, n)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        res["error"] = str(e)[:300]
        res["info_after"] = ro.info()
        print(json.dumps(res)); sys.stdout.flush()
        return 1
    res["idx_mismatch_frac"] = (base[7] != alt[7]).float().mean().item()
    same = (base[7] == alt[7]).all(dim=-1).all(dim=-1)
    names = ["latent", "hidden", "actions", "rewards", "continues", "mu", "sigma"]
    res["max_abs_diff_same_traj"] = {nm: (a[same] - b[same]).abs().max().item() for nm, a, b in zip(names, base, alt)}
    alt2 = ro.run(z0, h0, u, n)
    res["deterministic"] = all(torch.equal(a, b) for a, b in zip(alt, alt2))
    # timing: eager calls, L2 flushed between iterations
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for opt in (1, 0):
        L.check(lib.drm_set_option(b"persist", opt), "opt")
        for _ in range(3):
            ro.run(z0, h0, u, n, want_idx=False)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ro.run(z0, h0, u, n, want_idx=False); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res["ms_persist" if opt else "ms_chain"] = ts[len(ts) // 2]
    L.check(lib.drm_set_option(b"persist", 1), "opt")
    print(json.dumps(res)); sys.stdout.flush()
    return 0


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        sys.exit(child(sys.argv[2]))
    cases = sys.argv[1:] or list(CASES)
    for c in cases:
        try:
            p = subprocess.run([sys.executable, __file__, "--child", c], capture_output=True, text=True, timeout=180)
            print(p.stdout.strip() or ("no output; stderr: " + p.stderr[-600:]))
            if p.returncode:
                print("  rc", p.returncode, p.stderr[-400:].replace("\n", " | "))
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": c, "error": "timeout"}))
        sys.stdout.flush()
