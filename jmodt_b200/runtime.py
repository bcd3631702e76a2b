"""Stream / CUDA-graph runtime around the hot path (no reference counterpart: the reference launches ≈150 kernels
per frame one by one on the legacy default stream, `tools/eval.py:55-107`).

`CapturedPath` records ONE invocation of a step function — every kernel of this library, the torch glue between them
and the fork/join of the coordinate side stream (detector.PointNet2MSG._geometry) — into a CUDA graph over static
buffers, and replays it with a single launch.  The step of BASELINE config 3 is ≈350 device activities; replaying
it as a graph takes the Python / ctypes enqueue cost (≈10 ms of host time per step) off the critical path.

`TimedRecord` / `EventLog` are the CUDA-event bookkeeping bench.py uses to time individual launches; inside a
captured graph the events are *external* event-record nodes, so they are re-recorded by every replay.
"""
from __future__ import annotations

import os
from typing import Callable, Dict

import torch

from . import _lib


class TimedRecord:
    __slots__ = ("name", "flops", "kind", "desc", "e0", "e1", "ms")

    def __init__(self, name, flops, kind, desc, e0, e1):
        self.name, self.flops, self.kind, self.desc, self.e0, self.e1, self.ms = name, flops, kind, desc, e0, e1, []

    def collect(self):
        """Read the pair's elapsed time (call after a synchronize; once per replay for graph-resident pairs)."""
        self.ms.append(self.e0.elapsed_time(self.e1))


class EventLog:
    """CUDA-event pairs on the launching stream.  `external=True` while a graph is being captured."""

    def __init__(self):
        self.enabled = False
        self.external = False
        self.records = []
        self.launches = 0

    def reset(self):
        self.records, self.launches = [], 0

    def timed(self, name, fn, flops=0.0, kind="", desc=""):
        self.launches += 1
        if not self.enabled:
            return fn()
        e0 = torch.cuda.Event(enable_timing=True, external=self.external)
        e1 = torch.cuda.Event(enable_timing=True, external=self.external)
        e0.record()
        out = fn()
        e1.record()
        self.records.append(TimedRecord(name, flops, kind, desc, e0, e1))
        return out

    def collect(self):
        for r in self.records:
            r.collect()

    def collect_pending(self):
        """Eager mode: every record is one launch; read those not read yet."""
        for r in self.records:
            if not r.ms:
                r.collect()

    def summary(self, kind=None, name=None):
        recs = [r for r in self.records if (kind is None or r.kind == kind) and (name is None or r.name == name)]
        return {"flops": float(sum(r.flops * len(r.ms) for r in recs)), "ms": float(sum(sum(r.ms) for r in recs)),
                "launches": int(sum(len(r.ms) for r in recs))}


class CapturedPath:
    """fn(inputs) -> dict of tensors, captured once and replayed as one CUDA graph.

    `inputs` is a dict of device tensors that become the graph's static input buffers: `load()` copies new data into
    them (from pinned host memory or device tensors) in stream order, `replay()` launches the graph and returns the
    static output tensors (overwritten by the next replay).  fn must be free of host synchronisation, and must have
    been called at least once before (lazy initialisation — weight packing, cudaFuncSetAttribute, stream creation —
    cannot happen during capture)."""

    def __init__(self, fn: Callable[[Dict[str, torch.Tensor]], Dict[str, torch.Tensor]],
                 inputs: Dict[str, torch.Tensor], warmup: int = 2):
        assert all(t.is_cuda for t in inputs.values()), "static inputs live on the device"
        self.inputs = inputs
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count
        # kernel nodes inherit the priority of the stream they were captured on: the step is captured on a high-priority
        # stream (JMB_MAIN_PRIORITY, default -1; its forked branch and background streams stay at the default priority), so
        # the chain of small launches on the critical path is scheduled ahead of the wide background launches (the image
        # decoder's 1 280 CTAs): 7.53 ms per step against 7.62
        prio = int(os.environ.get("JMB_MAIN_PRIORITY", "-1"))
        cap_stream = torch.cuda.Stream(priority=prio) if prio else None
        with torch.cuda.graph(self.graph, stream=cap_stream, capture_error_mode="thread_local"):
            self.outputs = fn(inputs)
        self.launches_per_replay = _lib.launch_count - n0     # this library's kernels inside one replay
        torch.cuda.synchronize()

    def load(self, src: Dict[str, torch.Tensor], non_blocking: bool = True):
        """Copy new data into the static input buffers (keys of `src` that are inputs of the captured step)."""
        for k, t in self.inputs.items():
            if k in src:
                t.copy_(src[k], non_blocking=non_blocking)

    def replay(self):
        self.graph.replay()
        return self.outputs


class StreamedPath:
    """Host-to-host streaming around a CapturedPath: the inputs of step i + 1 cross PCIe while step i computes, and the
    results of step i are read back while step i + 1 computes.

        sp = StreamedPath(path, host_inputs, out_keys)
        for batch in batches:
            results = sp.step(batch)        # pinned host tensors of the step submitted ONE call earlier (None at first)
        last = sp.drain()

    Every step still moves its own inputs host -> device (pinned memory -> staging buffers on a copy stream -> the graph's
    static inputs with a device-side copy in front of the replay) and its own results device -> host (static outputs ->
    staging on the compute stream, -> pinned host buffers on the copy stream); only the waiting is gone."""

    def __init__(self, path: CapturedPath, host_inputs: Dict[str, torch.Tensor], out_keys, extra_outputs=None):
        self.path, self.out_keys = path, list(out_keys)
        self.extra = extra_outputs          # optional fn(outputs) -> list of extra device tensors to read back
        dev = next(iter(path.inputs.values())).device
        self.copy = torch.cuda.Stream(device=dev)
        self.stage_in = {k: torch.empty_like(path.inputs[k]) for k in host_inputs if k in path.inputs}
        self.in_landed, self.in_consumed = torch.cuda.Event(), torch.cuda.Event()
        self.in_consumed.record()
        self.slots, self.pending, self.n = [None, None], [None, None], 0

    def _outputs(self, out):
        ts = [out[k] for k in self.out_keys]
        if self.extra is not None:
            ts += list(self.extra(out))
        return ts

    def step(self, host_inputs: Dict[str, torch.Tensor]):
        main = torch.cuda.current_stream()
        slot = self.n & 1
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(self.in_consumed)              # the previous step's device-side load has read the staging
            for k, t in self.stage_in.items():
                t.copy_(host_inputs[k], non_blocking=True)
            self.in_landed.record(self.copy)
        main.wait_event(self.in_landed)
        self.path.load(self.stage_in)                            # device -> device, in front of the replay
        self.in_consumed.record(main)
        out = self.path.replay()
        ts = self._outputs(out)
        if self.slots[slot] is None:
            self.slots[slot] = ([torch.empty_like(t) for t in ts],
                                [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in ts],
                                torch.cuda.Event(), torch.cuda.Event())
        dev_stage, host_out, staged, landed = self.slots[slot]
        if self.pending[slot] is not None:
            landed.synchronize()                                 # the host is done waiting for the results of step n - 2
        torch._foreach_copy_(dev_stage, ts)                      # free the static outputs for the next replay
        staged.record(main)
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(staged)
            for h, d in zip(host_out, dev_stage):
                h.copy_(d, non_blocking=True)
            landed.record(self.copy)
        self.pending[slot] = host_out
        self.n += 1
        prev = self.n & 1                                        # the slot of the step submitted one call earlier
        if self.pending[prev] is not None and self.n >= 2:
            self.slots[prev][3].synchronize()
            return self.pending[prev]
        return None

    def drain(self):
        """Results of the last submitted step (waits for its read-back)."""
        last = (self.n - 1) & 1
        self.slots[last][3].synchronize()
        return self.pending[last]


class GeometryAhead:
    """Two batches in flight: while the feature stages of batch k run on the caller's stream, the coordinate-only stage
    of batch k + 1 (`geometry_fn`: FPS, ball queries, three_nn — a 4 ms dependent chain that occupies 32 of the 148
    SMs, pure latency) runs on a side stream, and its result is handed over at the end of the step.

        ahead = GeometryAhead(model.geometry, first_batch_points)       # eager: geometry of batch 0
        out_k = ahead.step(lambda plan: model(batch_k, geometry=plan), points_of_batch_k_plus_1)

    Every batch's geometry is computed exactly once, one step early; results are bit-identical to the unpipelined
    forward.  `plan` lives in static buffers (the side stream's result is copied into them after the join), so a
    step can be captured into a CUDA graph (CapturedPath) and replayed."""

    def __init__(self, geometry_fn, first_points: torch.Tensor):
        self.geometry_fn = geometry_fn
        self.plan = geometry_fn(first_points)
        self.side = torch.cuda.Stream(device=first_points.device, priority=int(os.environ.get("JMB_GEO_PRIORITY", "-1")))

    def step(self, main_fn, next_points: torch.Tensor):
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            nxt = self.geometry_fn(next_points)
        out = main_fn(self.plan)
        main.wait_stream(self.side)
        for t in nxt.tensors():
            t.record_stream(main)
        self.plan.copy_from(nxt)
        return out


# ---- independent branches on forked streams -------------------------------------------------------------------------
branch_parallel = os.environ.get("JMB_BRANCH_PARALLEL", "1") != "0"      # False: branches run one after the other on the caller's stream
_branch_streams: dict = {}


def parallel(*fns):
    """Run independent closures concurrently: fns[0] on the caller's stream, the others on forked side streams that
    join before returning (inside a CUDA-graph capture the fork/join becomes graph edges).  The hot path is a chain
    of small launches that each leave most of the 148 SMs idle or sit in their prologue / tail for a third of their
    run time — the two scales of a multi-scale set-abstraction level, the three input projections of the LI-Fusion
    attention, the cls / reg heads — so sibling launches overlap.  Tensors produced on a side stream are handed to
    the caller's stream with record_stream()."""
    if len(fns) == 1 or not branch_parallel or not torch.cuda.is_available():
        return [f() for f in fns]
    main = torch.cuda.current_stream()
    dev = main.device
    # one pool per CALLING stream: two callers on different streams (the step and the coordinate stage of the next batch)
    # must not queue their branches on the same side stream, or one caller's branch waits behind the other's whole chain
    pool = _branch_streams.setdefault((dev, main.cuda_stream), [])
    while len(pool) < len(fns) - 1:
        pool.append(torch.cuda.Stream(device=dev))
    fork = torch.cuda.Event()
    fork.record(main)
    outs = [None] * len(fns)
    joins = []
    for i, f in enumerate(fns[1:], 1):
        s = pool[i - 1]
        s.wait_event(fork)
        with torch.cuda.stream(s):
            outs[i] = f()
            ev = torch.cuda.Event()
            ev.record(s)
        joins.append(ev)
    outs[0] = fns[0]()
    for ev in joins:
        main.wait_event(ev)

    def hand_over(o):
        if isinstance(o, torch.Tensor):
            o.record_stream(main)
        elif isinstance(o, (list, tuple)):
            for x in o:
                hand_over(x)
    for o in outs[1:]:
        hand_over(o)
    return outs


_background_streams: dict = {}


class Spawned:
    """Result of spawn(): join() orders the caller's stream behind the background work and returns its output."""

    def __init__(self, out, done):
        self._out, self._done = out, done

    def join(self):
        if self._done is not None:
            main = torch.cuda.current_stream()
            main.wait_event(self._done)

            def hand_over(o):
                if isinstance(o, torch.Tensor):
                    o.record_stream(main)
                elif isinstance(o, (list, tuple)):
                    for x in o:
                        hand_over(x)
            hand_over(self._out)
            self._done = None
        return self._out


def spawn(fn) -> Spawned:
    """Start `fn` on a background stream of its own (not one of parallel()'s sibling streams, which the caller keeps
    using meanwhile) behind everything queued so far; the caller continues and join()s where it needs the result.  For a
    long launch whose inputs are ready early and whose output is needed late — the image decoder next to the chain of
    small set-abstraction / feature-propagation launches.  Inside a CUDA-graph capture the fork / join become graph edges."""
    if not branch_parallel or not torch.cuda.is_available():
        return Spawned(fn(), None)
    main = torch.cuda.current_stream()
    dev = main.device
    s = _background_streams.get(dev)
    if s is None:
        s = _background_streams[dev] = torch.cuda.Stream(device=dev)
    fork = torch.cuda.Event()
    fork.record(main)
    s.wait_event(fork)
    with torch.cuda.stream(s):
        out = fn()
        done = torch.cuda.Event()
        done.record(s)
    return Spawned(out, done)
