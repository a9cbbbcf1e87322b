"""World-size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding, gradient bucket
ranges over the flat staging buffer, and the bucketed all-reduce + averaging driven the way
Engine.backward drives it.  The CUDA kernels are not involved: a stand-in object exposes the same
fields the reducer reads from an Engine."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_shard_range_partitions_exactly():
    from tiaozhanbei_unet_b200.distributed import shard_range
    for n in (64, 63, 7, 1, 0):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("arch", ["unet", "anomaly_unet"])
def test_segments_cover_flat_buffer_in_backward_order(arch):
    from tiaozhanbei_unet_b200.distributed import segment_ranges
    from tiaozhanbei_unet_b200.engine import param_layout
    from tiaozhanbei_unet_b200.model import AnomalyUNet, UNet
    m = UNet(3, 1) if arch == "unet" else AnomalyUNet(3)
    named = list(m.named_parameters())
    offs, total = param_layout(named)
    n_dec = 1 if arch == "unet" else 2
    seg = segment_ranges([k for k, _ in named], offs, total, n_dec)
    spans = sorted(seg.values())
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a < b           # disjoint, contiguous, non-empty
    # every parameter lies inside exactly the bucket of its module group
    for k, p in named:
        if k.startswith(("inc.", "down1.", "down2.")):
            tag = ("enc", 0)
        elif k.startswith(("down3.", "down4.")):
            tag = ("enc", 3)
        else:
            tag = ("dec", 1 if k.split(".")[0].endswith("_seg") else 0)
        a, b = seg[tag]
        assert a <= offs[k] and offs[k] + p.numel() <= b, (k, tag)


class _FakeEngine:
    def __init__(self, names, offsets, total, n_dec, rank):
        self.param_names, self.offsets, self.total = names, offsets, total
        self.decoders = [None] * n_dec
        g = torch.Generator().manual_seed(100 + rank)
        self.gstage = torch.randn(total, generator=g)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tiaozhanbei_unet_b200.distributed import GradReducer
        names = ["inc.a", "down1.a", "down2.a", "down3.a", "down4.a", "up1_recon.up.weight", "up4_recon.x",
                 "outc_recon.w", "up1_seg.up.weight", "outc_seg.w"]
        offsets, total = {}, 0
        for i, k in enumerate(names):
            offsets[k] = total
            total += 64 * (i + 1)
        eng = _FakeEngine(names, offsets, total, 2, rank)
        mine = eng.gstage.clone()
        red = GradReducer()
        # same order as Engine.backward: decoders in reverse, then the two encoder halves
        seen = {}
        for tag in (("dec", 1), ("dec", 0), ("enc", 3), ("enc", 0)):
            # then(): what consumes the reduced bucket (the fused parameter update) runs behind its all-reduce
            a, b = red._segments(eng)[tag]
            red.segment_ready(eng, tag, then=(lambda tag=tag, a=a, b=b: seen.__setitem__(tag, eng.gstage[a:b].clone()))
                              if tag[0] == "enc" else None)
        scale = red.finish(eng)
        others = [torch.randn(total, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        expect = sum(others) / world
        for tag, got in seen.items():
            a, b = red._segments(eng)[tag]
            assert torch.allclose(got * scale, expect[a:b], atol=1e-6), tag       # saw the REDUCED bucket
        assert set(seen) == {("enc", 3), ("enc", 0)}
        ok = torch.allclose(eng.gstage * scale, expect, atol=1e-6) and abs(scale - 1.0 / world) < 1e-12
        ok = ok and red.bytes_reduced == total * 4 and not torch.equal(mine, eng.gstage)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
