"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol of
include/unetk.h, the drop-in modules have the reference's parameter layout / state_dict keys /
initialisation stream, the packing recipes are the right permutations, and the product path fails
loudly (no CPU fallback)."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402


def test_library_exports_every_declared_symbol():
    from tiaozhanbei_unet_b200 import _lib
    header = open(os.path.join(ROOT, "include", "unetk.h")).read()
    declared = set(re.findall(r"\b(unetk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = _lib.get()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/unetk.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes signature table and header disagree"
    assert lib.unetk_version() == 100


@pytest.mark.parametrize("arch", ["unet", "anomaly_unet"])
def test_state_dict_layout_matches_reference(arch):
    from tiaozhanbei_unet_b200.model import AnomalyUNet, UNet
    m = UNet(3, 1) if arch == "unet" else AnomalyUNet(3)
    sd = m.state_dict()
    spec = O.state_dict_spec(arch)   # pinned against the reference by tests/test_oracle.py
    assert list(sd.keys()) == list(spec.keys())
    for k, (shape, dtype) in spec.items():
        assert tuple(sd[k].shape) == tuple(shape) and sd[k].dtype == dtype, k
    n = sum(p.numel() for p in m.parameters())
    assert n == (31037633 if arch == "unet" else 43228228)
    # strict load of a reference-shaped state_dict
    m.load_state_dict(O.synthetic_state_dict(arch, seed=1), strict=True)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="live reference only in the build container")
def test_initialisation_stream_matches_reference():
    """Same seed -> same initial weights as the reference constructors (src/train.py:106 seeds 42)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import import_reference
    ref = import_reference()
    from tiaozhanbei_unet_b200.model import AnomalyUNet
    torch.manual_seed(42)
    a = ref["model"].AnomalyUNet(3).state_dict()
    torch.manual_seed(42)
    b = AnomalyUNet(3).state_dict()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_packing_recipes_are_the_documented_permutations():
    from tiaozhanbei_unet_b200 import packing

    def apply(recipe, src):
        (d0, d1, d2), (s0, s1, s2), flip = recipe
        out = np.empty(d0 * d1 * d2, dtype=src.dtype)
        for i0 in range(d0):
            for i1 in range(d1):
                j1 = d1 - 1 - i1 if flip else i1
                for i2 in range(d2):
                    out[(i0 * d1 + i1) * d2 + i2] = src[i0 * s0 + j1 * s1 + i2 * s2]
        return out

    co, ci = 4, 3
    w = np.arange(co * ci * 9, dtype=np.float32).reshape(co, ci, 3, 3)
    wf = apply(packing.conv_fwd_pack(co, ci), w.ravel()).reshape(co, 9, ci)
    assert np.array_equal(wf, w.reshape(co, ci, 9).transpose(0, 2, 1))
    wd = apply(packing.conv_dgrad_pack(co, ci), w.ravel()).reshape(ci, 9, co)
    assert np.array_equal(wd, w[:, :, ::-1, ::-1].reshape(co, ci, 9).transpose(1, 2, 0))
    back = apply(packing.conv_grad_unpack(co, ci), wf.ravel()).reshape(co, ci, 3, 3)
    assert np.array_equal(back, w)
    wt = np.arange(ci * co * 4, dtype=np.float32).reshape(ci, co, 2, 2)
    tf = apply(packing.convt_fwd_pack(ci, co), wt.ravel()).reshape(4, co, ci)
    assert np.array_equal(tf, wt.reshape(ci, co, 4).transpose(2, 1, 0))
    td = apply(packing.convt_dgrad_pack(ci, co), wt.ravel()).reshape(ci, 4, co)
    assert np.array_equal(td, wt.reshape(ci, co, 4).transpose(0, 2, 1))
    gk = wt.reshape(ci, co, 4).transpose(1, 2, 0).copy()      # [co][tap][ci]
    assert np.array_equal(apply(packing.convt_grad_unpack(ci, co), gk.ravel()).reshape(ci, co, 2, 2), wt)
    tab = packing.PermuteTable(2048)
    tab.add(0, 0, *packing.conv_fwd_pack(64, 64))
    tab.add(100, 200, *packing.identity(5000))
    arr, chunks = tab.numpy()
    assert arr.dtype.itemsize == 64 and chunks == 18 + 3 and arr[1]["chunk_begin"] == 18


def test_update_table_layout():
    """Descriptor table of unetk_fused_update (packing.UpdateTable): 64-byte records, block ranges per tensor."""
    from tiaozhanbei_unet_b200 import packing
    tab = packing.UpdateTable(2048)
    tab.add_plain(0, 5000)                         # 3 chunks of 2048
    tab.add_conv(5056, 0, 64 * 64 * 9, 64, 64)     # 2 x 2 tiles of 32 x 32
    tab.add_convt(50000, 100000, 200000, 128, 64)  # [cin 128][cout 64][2][2]: 4 x 2 tiles
    arr, blocks = tab.numpy()
    assert arr.dtype.itemsize == 64 and blocks == 3 + 4 + 8
    assert list(arr["block_begin"]) == [0, 3, 7] and list(arr["kind"]) == [0, 1, 2]
    assert (arr["cout"][0], arr["cout"][1], arr["cin"][1], arr["cout"][2], arr["cin"][2]) == (5000, 64, 64, 64, 128)
    assert list(arr["wf_off"]) == [-1, 0, 100000] and list(arr["wd_off"]) == [-1, 64 * 64 * 9, 200000]
    with pytest.raises(ValueError):
        tab.add_conv(0, 0, 0, 48, 64)


def test_no_cpu_fallback():
    from tiaozhanbei_unet_b200 import train_utils
    from tiaozhanbei_unet_b200.model import UNet, Up
    m = UNet(3, 1)
    with pytest.raises(Exception) as e:
        m(torch.randn(1, 3, 32, 32))
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)
    with pytest.raises(RuntimeError):
        train_utils.CombinedLoss()(torch.rand(1, 3, 8, 8), torch.rand(1, 1, 8, 8), torch.rand(1, 3, 8, 8), torch.rand(1, 1, 8, 8))
    assert isinstance(Up(128, 64, bilinear=True).up, torch.nn.Upsample)      # parameter-free up-sampling branch


def test_fused_update_args_host_logic():
    """optim.FusedOptimizer.fused_update_args (what Engine.train_step(update=...) is handed): the nine hyper-parameter
    floats are the ones unetk_adam_step derives (float arguments, double bias corrections), the step count advances
    like step() and rolls back, the moment buffers are the flat ones behind a torch-shaped state_dict; an optimizer
    that does not own exactly the engine's parameters gets None.  Host logic only: a stand-in engine on the CPU."""
    import math
    import types
    from tiaozhanbei_unet_b200.optim import FusedOptimizer
    flat = torch.zeros(192)
    names = ["a.weight", "a.bias"]
    offsets = {"a.weight": 0, "a.bias": 128}
    params = []
    for k, shape in (("a.weight", (8, 12)), ("a.bias", (8,))):
        p = torch.nn.Parameter(torch.zeros(shape))
        n = p.numel()
        p.data = flat[offsets[k]: offsets[k] + n].view(shape)
        params.append(p)
    eng = types.SimpleNamespace(flat_p=flat, params=params, param_names=names, offsets=offsets, total=192)
    opt = FusedOptimizer(params, "adam", lr=3e-3, weight_decay=1e-2, engine=eng)
    u = opt.fused_update_args(eng)
    assert u is not None and u.mode == 0 and u.m.numel() == 192 and u.v.numel() == 192
    f32 = lambda x: float(np.float32(x))
    b1, b2 = f32(0.9), f32(0.999)
    want = [f32(3e-3), b1, b2, f32(1e-8), f32(1e-2), f32(1.0 - math.pow(b1, 1)), f32(math.sqrt(1.0 - math.pow(b2, 1))), 1.0, 1.0]
    assert u.values == want
    sd = opt.state_dict()
    assert float(sd["state"][0]["step"]) == 1.0 and sd["state"][1]["exp_avg"].shape == (8,)
    assert opt.state[params[0]]["exp_avg"].data_ptr() == u.m.data_ptr()          # views of the flat buffer
    u2 = opt.fused_update_args(eng)
    assert u2.values[7] == 0.0 and u2.values[5] == f32(1.0 - math.pow(b1, 2)) and float(opt.state_dict()["state"][0]["step"]) == 2.0
    u2.rollback()
    assert float(opt.state_dict()["state"][0]["step"]) == 1.0
    opt.param_groups[0]["lr"] = 1e-4
    assert opt.fused_update_args(eng).values[0] == f32(1e-4)
    # SGD: momentum in the beta1 slot, first-step flag from the missing momentum buffer
    sgd = FusedOptimizer(params, "sgd", lr=0.1, weight_decay=0.0, engine=eng)
    us = sgd.fused_update_args(eng)
    assert us.mode == 2 and us.v is None and us.values[1] == f32(0.9) and us.values[7] == 1.0
    assert sgd.fused_update_args(eng).values[7] == 0.0
    # not this engine's parameters / another engine / two parameter groups: no fused update
    assert FusedOptimizer(params[:1], "adam", 1e-3, 0.0, engine=eng).fused_update_args(eng) is None
    assert opt.fused_update_args(types.SimpleNamespace(flat_p=flat)) is None
    two = FusedOptimizer([{"params": params[:1]}, {"params": params[1:]}], "adam", 1e-3, 0.0, engine=eng)
    assert two.fused_update_args(eng) is None


def test_optimizer_state_dict_is_torch_shaped():
    from tiaozhanbei_unet_b200 import train_utils
    lin = torch.nn.Linear(4, 3)
    ours = train_utils.get_optimizer(lin, "adam", 1e-3, 1e-4)
    ref = torch.optim.Adam(lin.parameters(), lr=1e-3, weight_decay=1e-4)
    a, b = ours.state_dict(), ref.state_dict()
    assert a["param_groups"][0].keys() == b["param_groups"][0].keys()
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
        assert a["param_groups"][0][k] == b["param_groups"][0][k]
    ours.load_state_dict(b)          # a torch-Adam checkpoint loads
    ref.load_state_dict(a)
    with pytest.raises(ValueError):
        train_utils.get_optimizer(lin, "rmsprop")
    sch = train_utils.get_scheduler(ours, "cosine", 10)
    assert isinstance(sch, torch.optim.lr_scheduler.CosineAnnealingLR)
    assert train_utils.get_scheduler(ours, "none") is None


def test_graph_state_protocol(monkeypatch):
    """engine._GraphState (CUDA-graph capture of a static launch sequence): call 1 runs eagerly, call 2
    captures and replays, later calls only replay; a failed capture falls back to eager launches for good.
    torch.cuda's graph objects are replaced by recorders (no GPU here)."""
    from tiaozhanbei_unet_b200 import engine

    log = []

    class FakeGraph:
        def replay(self):
            log.append("replay")

    class FakeCapture:
        def __init__(self, g, capture_error_mode=None):
            assert capture_error_mode == "thread_local"

        def __enter__(self):
            log.append("capture-begin")

        def __exit__(self, *exc):
            log.append("capture-end")
            return False

    monkeypatch.setattr(engine.torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(engine.torch.cuda, "graph", FakeCapture)
    monkeypatch.setattr(engine.torch.cuda, "synchronize", lambda: log.append("sync"))
    calls = []

    def body():
        calls.append(len(calls))
        return ["out%d" % len(calls)]

    st = engine._GraphState()
    assert st.run(body) == (["out1"], False) and log == []                      # warm-up: eager
    res, replayed = st.run(body)                                                # capture, then one replay
    assert replayed and res == ["out2"] and log == ["capture-begin", "capture-end", "replay"]
    res, replayed = st.run(body)                                                # replay only: body not called again
    assert replayed and res == ["out2"] and len(calls) == 2 and log[-1] == "replay" and log.count("replay") == 2

    def bad():
        if log and log[-1] == "capture-begin":
            raise RuntimeError("operation not permitted when stream is capturing")
        return ["eager"]

    st2 = engine._GraphState()
    st2.run(bad)
    with pytest.warns(UserWarning):
        res, replayed = st2.run(bad)
    assert res == ["eager"] and not replayed and st2.failed and st2.graph is None
    assert st2.run(bad) == (["eager"], False)                                   # stays eager


def test_train_epoch_host_logic_on_cpu_stand_ins():
    """The loop of train_epoch itself (src/train_utils.py:107-153) with CPU stand-ins for model / criterion /
    optimizer: every batch is stepped once and in order, the three averages are weighted by the batch size
    (ragged last batch), and a step's losses are accounted even though they are read one step late."""
    import torch
    from tiaozhanbei_unet_b200 import train_utils as T

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))
            self.seen = []

        def forward(self, x):
            self.seen.append(float(x.mean()))
            return x * self.w, x[:, :1] * self.w

    class Criterion:
        def __call__(self, recon, amap, image, mask):
            r, s = (recon ** 2).mean(), (amap - mask).abs().mean()
            return {"total_loss": r + s, "recon_loss": r.detach(), "seg_loss": s.detach()}

    sizes = (4, 4, 3)
    batches = [{"image": torch.full((n, 3, 4, 4), float(i + 1)), "mask": torch.zeros(n, 1, 4, 4)} for i, n in enumerate(sizes)]
    model = Model()
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    out = T.train_epoch(model, batches, Criterion(), opt, torch.device("cpu"), 3)
    assert model.seen == [1.0, 2.0, 3.0] and model.training
    n = sum(sizes)
    want_r = sum(s * float(i + 1) ** 2 for i, s in enumerate(sizes)) / n
    want_s = sum(s * float(i + 1) for i, s in enumerate(sizes)) / n
    assert abs(out["recon_loss"] - want_r) < 1e-6 and abs(out["seg_loss"] - want_s) < 1e-6
    assert abs(out["total_loss"] - (want_r + want_s)) < 1e-6
    assert set(out) == {"total_loss", "recon_loss", "seg_loss"}
    # an empty loader returns the reference's zero averages
    assert T.train_epoch(model, [], Criterion(), opt, torch.device("cpu"), 4) == {"total_loss": 0, "recon_loss": 0, "seg_loss": 0}
