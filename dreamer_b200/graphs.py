"""CUDA-graph replay of whole training steps (streams and graphs instead of a tracing compiler).

A training step of the hot path is thousands of small launches (the teacher-forced gradient tail walks T GRU steps forward and
backward); on a B200 the CPU cannot issue them as fast as the GPU retires them.  ``StepGraph`` runs the first few calls of a
step eagerly -- real steps on real data, which also creates every workspace, cuDNN plan and packed-weight buffer -- then
captures ONE call (the kernels of this package, the torch autograd tail, the flat-bucket all-reduce and the fused optimiser
are all sync-free) and replays it for every later call with the same input shapes.  Inputs are copied into the captured
static buffers; outputs are the captured static tensors (overwritten by the next call).

Under torch.distributed the NCCL all-reduces are part of the capture (tests/dist_nccl_check.py, 2 GPUs): the capture runs in
thread-local error mode because the NCCL watchdog thread polls CUDA events meanwhile, and the graphs must be dropped
(``module.__dict__["_graphs"] = None``) before ``destroy_process_group()`` -- they hold NCCL work.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class StepGraph:
    def __init__(self, body: Callable, warmup: int = 3, before_capture: Optional[Callable] = None,
                 after_replay: Optional[Callable] = None):
        self.body, self.warmup = body, max(1, int(warmup))
        self.before_capture, self.after_replay = before_capture, after_replay
        self.cache = {}

    @staticmethod
    def _key(args):
        return tuple(None if a is None else (tuple(a.shape), a.dtype, str(a.device)) for a in args)

    def captured(self, *args) -> bool:
        return "graph" in self.cache.get(self._key(args), {})

    def __call__(self, *args):
        for a in args:
            if a is not None and not (isinstance(a, torch.Tensor) and a.is_cuda):
                raise RuntimeError("StepGraph: arguments must be CUDA tensors or None")
        ent = self.cache.setdefault(self._key(args), {"calls": 0})
        if "graph" not in ent:
            if ent["calls"] < self.warmup:
                ent["calls"] += 1
                return self.body(*args)
            static = [None if a is None else a.clone() for a in args]
            if self.before_capture is not None:
                self.before_capture()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            # thread-local capture mode: under torch.distributed the NCCL watchdog thread polls events while this thread captures
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out = self.body(*static)
            ent.update(graph=graph, static=static, out=out)
        with torch.no_grad():
            for s, a in zip(ent["static"], args):
                if s is not None and s.data_ptr() != a.data_ptr():
                    s.copy_(a)
        ent["graph"].replay()
        if self.after_replay is not None:
            self.after_replay()
        return ent["out"]
