"""One CUDA graph per denoise step.

A denoise step of the story pipeline is a fixed sequence of launches per attention layer — projections, gather (or
peer exchange), attention, output projection — plus, once per step, the re-sampling and compaction of the masks
(StoryDiffusion/Comic_Generation.py:119-125).  Issued call by call from Python it costs ~130 us of CPU per layer,
which is what bounds a step once the per-GPU work shrinks (8 GPUs: 4.8 ms of issue time against 1.9 ms of GPU work,
round-1 verdict).  Everything a step touches is pointer-stable:

  * the masks are re-sampled IN PLACE (``CompactMask.resample_``), so the index lists, the sampled list and its
    runs keep their addresses and every kernel reads its geometry (counts, ranges) from device memory;
  * the peer exchange numbers its calls relative to a device-resident epoch base that a one-thread kernel advances
    at the end of the step (``PeerExchange.end_step``), so a replay publishes and awaits fresh flags;
  * activations, Q/K/V/O and the exchange views come from the graph's private memory pool;

so the whole step can be captured once and replayed: ``StepGraph(step_fn).replay()`` is one ``cudaGraphLaunch``.

What a captured step FIXES is its host-side control flow: which branch every call took (the ``random.random()`` gate
of Comic_Generation.py:98-103, the ``cur_step < 5`` cut), ``write`` and the ``cur_step`` slot of the id_bank it reads
or writes.  A replay re-executes exactly those launches on whatever the input buffers hold now — the host's counters
do not advance.  Drivers that want the reference's per-call gate draws call the processors eagerly (always valid);
``StepGraph`` is for fixed-pattern steps: benchmarking, the forced-consistent steady state, or one graph per distinct
pattern kept in a small cache by the caller.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import native


class StepGraph:
    """Capture ``step_fn()`` (no arguments; reads its inputs from tensors that stay allocated) and replay it."""

    def __init__(self, step_fn: Callable[[], object], device: Optional[torch.device] = None, warmup: int = 2):
        self.step_fn = step_fn
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.warmup = warmup
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.output = None
        self.stream = torch.cuda.Stream(self.device)
        self.replays = 0

    def capture(self, before_capture: Optional[Callable[[], None]] = None) -> "StepGraph":
        """Warm up on the capture stream, call ``before_capture`` (e.g. switch instrumentation on), capture."""
        native.flush_batch()
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            # eager runs on the capture stream first: lazily created state (cuBLASLt heuristics, the attention
            # workspace of this stream, projection plans, persistent mask buffers) must exist before the capture
            for _ in range(self.warmup):
                self.step_fn()
        cur.wait_stream(self.stream)
        torch.cuda.synchronize(self.device)
        if before_capture is not None:
            before_capture()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.output = self.step_fn()
            native.flush_batch()
        return self

    def replay(self):
        if self.graph is None:
            self.capture()
        self.graph.replay()
        self.replays += 1
        return self.output
