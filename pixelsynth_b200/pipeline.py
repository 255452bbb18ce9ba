"""Pipelined serving of ZbufferModelPts.forward: `depth` batches in flight on one GPU.

The reference runs one batch at a time (models/z_buffermodel.py:291-419 is one synchronous pass).  On a B200 that
leaves most of the device idle for a third of the step: the lmconv sampler is a chain of dependent levels that keeps a
few dozen SMs busy, while the refinement decoder that follows wants all of them.  ViewPipeline
keeps two batches in flight so that batch k+1's sampler runs beside batch k's decoder:

  * the device's SMs are split into two disjoint CUDA green contexts (ps_sm_partition_create): a small partition for
    every sampler launch, the rest for everything else; persistent kernels size their grids to their partition;
  * batch k runs on stream k % depth of the large partition; its sampler launch is moved to the small partition's
    stream behind events (ZbufferModelPts.sampler_stream); its front end (depth net, splat, VQ encoder: 3 ms of work
    the host has to wait for before it can build the generation order) runs on a high-priority stream of the large
    partition (ZbufferModelPts.front_stream), so it is not queued behind the previous batch's decoder and the sampler's
    partition never waits for the host;
  * the slots are shallow copies of one model: weights, packed plans and the sampler's activation cache are shared
    (sampler launches are serialised on their one stream), per-batch state (pinned mask buffer, `last`) is per slot.

Results are bit-identical to model.forward on the same inputs (tests/test_pipeline_gpu.py): nothing about the
arithmetic depends on the grid size or on what runs beside it.
"""
import copy
import ctypes
import os

import torch

from . import _lib
from ._lib import check


class ViewPipeline:
    def __init__(self, model, depth=2, sampler_sms=24, partition=True):
        """model: a ZbufferModelPts.  sampler_sms: size of the sampler's partition (24 of 148 balances the two sides
        at 64 views per batch).  partition=False keeps plain streams on the whole device;
        "auto" does so only when the driver cannot create green contexts (`partition_error` says why)."""
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.model, self.depth = model, depth
        self.device = torch.device(model.device if not isinstance(model.device, int) else f"cuda:{model.device}")
        if self.device.type != "cuda":
            raise RuntimeError("ViewPipeline needs a CUDA device")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.sm_counts, self.partition_error = None, None
        self.front_streams = [None] * depth
        if partition == "auto" and any(k in os.environ for k in ("CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR")):
            # Nsight Compute cannot profile kernels launched into a green context ("Failed to prepare kernel for profiling")
            partition, self.partition_error = False, "a profiler is attached (CUDA injection): plain streams"
        if partition and depth > 1:
            try:
                self.sampler_stream, self.streams, self.front_streams, self.sm_counts = _partition(index, int(sampler_sms), depth)
            except _lib.PixelSynthB200Error as e:
                if partition != "auto":
                    raise
                self.partition_error = str(e)     # a driver without green contexts: same schedule on plain streams
        if self.sm_counts is None:
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
            self.sampler_stream = torch.cuda.Stream(device=self.device) if depth > 1 else None
            if depth > 1:
                self.front_streams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(depth)]
        self.slots = []
        for i in range(depth):
            m = copy.copy(model)          # shares the sub-networks (weights, plans, sampler cache)
            m._bg_pin = None
            m.sampler_stream = self.sampler_stream
            m.front_stream = self.front_streams[i]
            self.slots.append(m)
        self._events = [None] * depth
        self._pending = {}
        self._n = 0

    # ------------------------------------------------------------------------------------------------
    def submit(self, batch, then=None, **kw):
        """Queues model.forward(batch, **kw) on the next slot and returns a ticket.  `then(loss, outputs)` (optional) is
        called inside the slot's stream context right after forward -- e.g. the rescale and the copy to pinned host
        memory -- and its return value replaces (loss, outputs) as the ticket's result.  Blocks only when the slot's
        previous batch has not finished (at most `depth` batches are in flight)."""
        slot = self._n % self.depth
        if self._events[slot] is not None:
            self._events[slot].synchronize()
        st = self.streams[slot]
        st.wait_stream(torch.cuda.current_stream(self.device))  # the caller's inputs
        with torch.cuda.stream(st):
            res = self.slots[slot].forward(batch, **kw)
            if then is not None:
                res = then(*res)
            ev = torch.cuda.Event()
            ev.record(st)
        self._events[slot] = ev
        ticket = self._n
        self._pending[ticket] = (ev, res, st)
        self._n += 1
        return ticket

    def result(self, ticket, host_wait=True):
        """The ticket's (loss, outputs) (or `then`'s value).  host_wait=True blocks the host until the batch is done;
        False only orders the caller's current stream behind it."""
        ev, res, st = self._pending.pop(ticket)
        cur = torch.cuda.current_stream(self.device)
        if host_wait:
            ev.synchronize()
            _lib.check_wedge("ViewPipeline")
        else:
            cur.wait_event(ev)
        for t in _tensors(res):
            if t.is_cuda:
                t.record_stream(cur)
        return res

    def last(self, ticket_or_slot=0):
        return self.slots[ticket_or_slot % self.depth].last

    def map(self, batches, **kw):
        """Generator: forward over an iterable of batches, `depth` in flight, results in order."""
        tickets = []
        for b in batches:
            tickets.append(self.submit(b, **kw))
            if len(tickets) >= self.depth:
                yield self.result(tickets.pop(0))
        for t in tickets:
            yield self.result(t)

    def close(self):
        for ev in self._events:
            if ev is not None:
                ev.synchronize()
        self._pending.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# Partitions live as long as the process: torch's caching allocator keeps per-stream pools for every stream it has
# seen, so the streams of a green context must not be destroyed under it (ps_sm_partition_destroy is for callers that
# own their allocator).  One partition per (device, sampler SMs, depth), shared by every ViewPipeline that asks for it.
_PARTITIONS = {}


def _partition(index, sampler_sms, depth):
    """-> (sampler stream on the small partition, `depth` streams and `depth` high-priority front-end streams on the
    large one, (small, large) SM counts)"""
    key = (index, sampler_sms, depth)
    if key not in _PARTITIONS:
        small = (ctypes.c_void_p * 1)()
        big = (ctypes.c_void_p * depth)()
        ns, nb, handle = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_void_p()
        dev = torch.device("cuda", index)
        L = _lib.lib()
        with torch.cuda.device(index):
            torch.zeros(1, device=dev)  # torch's context first
            check(L.ps_sm_partition_create(index, sampler_sms, 1, small, depth, big, ctypes.byref(ns), ctypes.byref(nb),
                                           ctypes.byref(handle)), "ps_sm_partition_create")
            front = []
            for _ in range(depth):
                st = ctypes.c_void_p()
                check(L.ps_sm_partition_stream(handle, 1, 1, ctypes.byref(st)), "ps_sm_partition_stream")
                front.append(torch.cuda.ExternalStream(int(st.value), device=dev))
        _PARTITIONS[key] = (torch.cuda.ExternalStream(int(small[0]), device=dev),
                            [torch.cuda.ExternalStream(int(big[i]), device=dev) for i in range(depth)], front,
                            (ns.value, nb.value))
    return _PARTITIONS[key]


def _tensors(x):
    if isinstance(x, torch.Tensor):
        yield x
    elif isinstance(x, dict):
        for v in x.values():
            yield from _tensors(v)
    elif isinstance(x, (list, tuple)):
        for v in x:
            yield from _tensors(v)
