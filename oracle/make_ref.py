"""Recipe for oracle/_ref/: a byte-for-byte copy of the reference's importable Python modules.

TEST / MEASUREMENT INFRASTRUCTURE -- nothing under dreamer_b200/ imports it.

The reference (youngers2006/Dreamer) is pure Python: there is nothing to compile, so "building" its CPU path means putting its
unmodified sources where the GPU box can import them (/root/reference does not exist there).  oracle/_ref/ is git-ignored (the
sources never enter this repository's history) but travels with the snapshot, like a built .so.  Consumers:

    bench.py --impl reference        the reference's own Dreamer.dream_episodes on the host cores       (cpu_baseline.kind "reference")
    bench.py --impl reference-cuda   the same unmodified code through stock PyTorch on the B200 (its real deployment mode)

Run in the build container:  python -m oracle.make_ref     (also called by __graft_entry__.build() when /root/reference exists)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# the modules Dreamer.py imports (Dreamer.py:1-11) + the shipped configuration and licence; Adaptors.py / train_car_racer.py need
# gymnasium / PyFlyt / matplotlib, which this image does not have, and are not on the hot path
FILES = ["Dreamer.py", "WorldModel.py", "Agent.py", "Buffer.py", "DreamerUtils.py", "DynamicsPredictors.py", "SequenceModel.py",
         "VariationalAutoEncoder.py", "car_racer_config.yaml", "LICENSE"]


def make(src: str = REF, dst: str = DST) -> dict:
    if not os.path.isdir(src):
        raise FileNotFoundError(f"{src} not found: oracle/_ref can only be (re)made in the build container")
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for name in FILES:
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
        with open(os.path.join(dst, name), "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(dict(source="youngers2006/Dreamer (unmodified copies, sha256 per file)", files=manifest), f, indent=1)
    return manifest


def available(dst: str = DST) -> bool:
    return all(os.path.exists(os.path.join(dst, n)) for n in FILES[:8])


def load_dreamer(cfg: dict, sd: dict, device):
    """The reference's unmodified Dreamer (Dreamer.py:13) on `device` with the weights `sd` (97 keys, strict)."""
    import torch
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from Dreamer import Dreamer
    d = Dreamer(dict(cfg, device=str(device)), torch.device(device))
    d.load_state_dict(sd, strict=True)
    return d


if __name__ == "__main__":
    m = make()
    print(f"oracle/_ref: {len(m)} files copied from {REF}")
