"""Data-parallel training and sharded inference (new capability; the reference is single-process).

One process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch) for the plumbing.  Training
shards the batch across ranks; the only exchange step is the gradient all-reduce.  Gradients are
staged by the backward kernels in ONE flat fp32 buffer laid out in parameter order, so a "bucket" is
a contiguous range of it: the two decoders and two halves of the encoder, reduced on a side stream as
soon as the backward pass has finished writing them (reverse layer order) and overlapped with the
remaining backward kernels.  The 1/world_size average is folded into the single un-packing launch
that produces `.grad`.  BatchNorm statistics stay local per rank by default (DDP semantics): parity
oracle = the reference run per shard with gradients averaged; attach(sync_bn=True)
synchronises them (parity oracle = the reference on the concatenated global batch).

Inference shards images contiguously across ranks with no communication.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [begin, end) slice of n_items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def segment_ranges(param_names, offsets, total, n_decoders):
    """Element ranges of the flat buffer for the reduction buckets, keyed by the tags the engine's
    backward emits: ("dec", d) when decoder d is done, ("enc", 3) after down4+down3, ("enc", 0) at the end."""
    def start_of(prefix):
        for k in param_names:
            if k.startswith(prefix):
                return offsets[k]
        raise KeyError(prefix)

    firsts = [k for k in param_names if k.startswith("up1")]
    dec_starts = []
    seen = set()
    for k in firsts:
        mod = k.split(".")[0]
        if mod not in seen:
            seen.add(mod)
            dec_starts.append(offsets[k])
    assert len(dec_starts) == n_decoders, (dec_starts, n_decoders)
    bounds = dec_starts + [total]
    seg = {("dec", d): (bounds[d], bounds[d + 1]) for d in range(n_decoders)}
    s_down3 = start_of("down3.")
    seg[("enc", 3)] = (s_down3, dec_starts[0])
    seg[("enc", 0)] = (0, s_down3)
    return seg


class GradReducer:
    """Attached to an Engine (engine.ddp = GradReducer(...)); driven by Engine.backward."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.stream = None
        self.works = []
        self.segments = None
        self.bytes_reduced = 0
        self.tail = None          # event behind the launches segment_ready(then=...) put on the reducer's stream

    def _segments(self, eng):
        if self.segments is None or self._total != eng.total:
            self.segments = segment_ranges(eng.param_names, eng.offsets, eng.total, len(eng.decoders))
            self._total = eng.total
        return self.segments

    def segment_ready(self, eng, tag, also_after=None, then=None):
        """Reduce the bucket `tag` once the current stream (and the events `also_after`, recorded on the
        engine's side streams) have finished writing it.  then(): launches that consume the reduced bucket
        (Engine.train_step's fused parameter update), enqueued on the reducer's stream right behind the all-reduce."""
        if self.world == 1:
            if then is not None:
                cur = torch.cuda.current_stream(eng.gstage.device)
                for e in (also_after or ()):
                    cur.wait_event(e)
                then()
            return
        a, b = self._segments(eng)[tag]
        view = eng.gstage[a:b]
        self.bytes_reduced += (b - a) * 4
        if view.is_cuda:
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=view.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(view.device))
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ev)
                for e in (also_after or ()):
                    self.stream.wait_event(e)
                work = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                if then is not None:
                    work.wait()             # orders the reducer's STREAM behind the all-reduce (the host does not block)
                    then()
                    self.tail = torch.cuda.Event()
                    self.tail.record(self.stream)
                else:
                    self.works.append(work)
        else:
            self.works.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            if then is not None:
                self.works.pop().wait()
                then()

    def finish(self, eng) -> float:
        """Make the current stream wait for all outstanding reductions; returns the averaging factor."""
        for w in self.works:
            w.wait()
        self.works = []
        if self.tail is not None:      # launches enqueued behind the all-reduces (segment_ready(then=...))
            torch.cuda.current_stream(eng.gstage.device).wait_event(self.tail)
            self.tail = None
        return 1.0 / self.world


def attach(model, group=None, sync_bn=False):
    """Enable gradient averaging across the process group for a UNet / AnomalyUNet of this package.
    sync_bn=True also synchronises the BatchNorm batch statistics (two small all-reduces per BatchNorm
    layer and step): training then equals the single-process reference on the concatenated global batch
    instead of per-shard BatchNorm (the default, = DistributedDataParallel semantics)."""
    eng = model._engine()
    eng.ddp = GradReducer(group)
    eng.sync_bn = (group, dist.get_world_size(group)) if sync_bn and dist.get_world_size(group) > 1 else None
    return eng.ddp


def broadcast_parameters(model, src=0, group=None):
    """Make every rank start from rank `src`'s weights and buffers (one flat broadcast + buffers)."""
    eng = model._engine()
    dev = next(model.parameters()).device
    eng.ensure(dev)
    dist.broadcast(eng.flat_p, src=src, group=group)
    for b in model.buffers():
        dist.broadcast(b, src=src, group=group)
    eng.mark_weights_dirty()
