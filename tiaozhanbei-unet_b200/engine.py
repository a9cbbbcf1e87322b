"""Static execution plan of one U-Net / AnomalyUNet: flat parameter storage, packed bf16 weights,
pre-allocated NHWC bf16 activations and the forward / backward kernel sequences.

What the reference does with nn.Module.__call__ + autograd (src/model.py:97-108, :188-210 and the
`total_loss.backward()` of src/train_utils.py:132), this file does with explicit launches of the
kernels declared in include/unetk.h, in a fixed order on torch's current stream:

  forward (train)   conv -> [BN stats in the conv epilogue] -> bn_finalize -> bnrelu(+pool+idx)
  forward (eval)    conv with folded BN + ReLU in the epilogue -> pool
  backward          head_bwd -> per conv: bn_bwd_reduce / bn_bwd_apply -> wgrad + dgrad;
                    skip connections accumulate their 2-3 gradients in place (dgrad epilogue /
                    pool backward read-modify-write); weight gradients are staged in
                    [Cout][tap][Cin] and re-laid-out to the reference layouts by one launch.

  train_step        forward, CombinedLoss, backward and (with this package's optimizer) the optimizer step as
                    ONE captured launch sequence: the heads compute their loss term and their own backward in one
                    kernel, and every parameter segment is un-packed, updated and re-packed by one kernel as soon
                    as its gradients are final (unetk_fused_update), so that two consecutive steps have nothing
                    between them.

Parameters stay `nn.Parameter`s under the reference's attribute paths (state_dict identical);
their storage is re-pointed to slices of ONE flat fp32 buffer so the optimizer step, the weight
re-packing, the gradient un-packing and the data-parallel all-reduce are each one launch / one
contiguous range.
"""
from __future__ import annotations

import contextlib
import os

import torch

from . import _lib, ops, packing

# Which producers of an activation gradient also run the first pass (per-channel sums) of the BatchNorm
# backward of the layer below them: "head" (default: the 1x1 head backward already holds the activation in
# registers, +0.01 ms instead of a 0.07 ms pass), "all" (+ conv / transposed-conv dgrad epilogues and the
# pool backward: measured slower at B=16 @256^2, the extra global loads sit on the epilogue's critical
# path), "0" (stand-alone reductions everywhere).
FUSE_BN_REDUCE = os.environ.get("UNETK_FUSE_BN_REDUCE", "head")
DEEP_PIXELS = int(os.environ.get("UNETK_FUSE_DEEP_PIXELS", "65536"))
FUSE_CHECK = os.environ.get("UNETK_FUSE_CHECK", "0") == "1"
FIRST_WGRAD_TC = os.environ.get("UNETK_FIRST_WGRAD_TC", "1") != "0"
# The first conv (Cin = 3) as the 64 -> 64 tcgen05 layer on the channel-padded bf16 image (0.19 -> ~0.12 ms
# at 16 x 256^2, and eval mode gets BatchNorm+ReLU in the epilogue); "0" = CUDA-core kernel on the fp32 image.
FIRST_FWD_TC = os.environ.get("UNETK_FIRST_FWD_TC", "1") != "0"
# Training forward: the BatchNorm finalize (batch sums -> scale/shift, running statistics) is folded into the
# BN+ReLU(+pool) kernel instead of a separate dependent launch per layer.
FUSE_BN_FINALIZE = os.environ.get("UNETK_FUSE_BN_FINALIZE", "1") != "0"
SPLIT_PACK = os.environ.get("UNETK_SPLIT_PACK", "1") != "0"
# Training forward: the BatchNorm + ReLU of every decoder's LAST conv is applied inside the 1x1 head kernel (and
# recomputed from the raw conv output in the head's backward), so that layer's activated tensor is never
# written or re-read: one full-resolution pass less per decoder, at the un-overlapped end of the forward pass.
HEAD_FUSE_BN = os.environ.get("UNETK_HEAD_FUSE_BN", "1") != "0"
# CUDA graphs: the plan is static (same kernels, same buffers every step), so after one eager warm-up call
# the whole forward (and backward) launch sequence, side streams included, is captured once per input shape
# and replayed with ONE call: Python enqueue drops from 6.5-10 ms per step (host dependent) to 2.8 ms (loss,
# optimizer, copies in and out of the static buffers).  "0" = always eager.  Under data parallelism only the
# forward pass is captured: UNETK_GRAPH_DDP=1 also captures the backward with its NCCL all-reduces, which
# ran correctly on 2 B200s (tests/checks/ddp_check.py) but hung in destroy_process_group() at exit.
# Profilers that inject into the process (Nsight Compute / Systems set CUDA_INJECTION64_PATH) serialise and
# replay kernel launches and cannot handle launches recorded during stream capture: stay eager under them.
USE_GRAPHS = os.environ.get("UNETK_GRAPHS", "1") != "0" and not os.environ.get("CUDA_INJECTION64_PATH")
GRAPH_DDP = os.environ.get("UNETK_GRAPH_DDP", "0") == "1"
# Weight gradients are off the critical path of backward (nothing downstream reads them before the
# optimizer): they run on a side stream, each one released only after the data gradient of the same layer,
# so the tensor-bound wgrad kernel shares the SMs with the memory-bound BatchNorm-backward passes of the
# next layer instead of queueing in front of them.
OVERLAP_WGRAD = os.environ.get("UNETK_OVERLAP_WGRAD", "1") != "0"
PARALLEL_DECODERS = os.environ.get("UNETK_PARALLEL_DECODERS", "1") != "0"
# The tensor-core kernels are persistent (one CTA per SM for the whole launch), so kernels of different streams
# that are ready together run one after the other, in the order the hardware picks them.  The chains whose
# kernels other work waits for (forward, data gradients) run on high-priority streams owned by the engine;
# weight gradients, weight packing and gradient un-packing stay at the default (lowest) priority and fill
# the holes.  Off by default: measured on a B200 (B=16 @256^2, 8 runs each way) the step takes 10.0-10.2 ms either way -- the
# backward pass is bound by the sum of its tensor-core kernels, not by their order -- and the default keeps
# everything on the caller's stream.
PRIORITY_STREAMS = os.environ.get("UNETK_PRIORITY_STREAMS", "0") == "1"
# Release a weight-gradient launch as soon as its inputs exist (with the data gradient of the same layer,
# not after it); only sensible together with PRIORITY_STREAMS, which lets the data gradient go first.
WGRAD_EARLY = os.environ.get("UNETK_WGRAD_EARLY", "0") == "1"
# Slots of the d_y / d_up scratch per backward chain: a chain may run this many layers ahead of the (lower
# priority, hence later) weight gradients that still read them before it has to wait for one.
GRAD_SLOTS = max(2, int(os.environ.get("UNETK_GRAD_SLOTS", "2")))
# Plans (activations, scratch, CUDA graphs of one input shape) kept alive at once, least recently used evicted:
# train batch, ragged last train batch, validation batch, ragged last validation batch.
MAX_PLANS = max(1, int(os.environ.get("UNETK_MAX_PLANS", "4")))
# Grid bound of a fused parameter update that runs beside the backward pass (blocks of 256 threads walking the work
# units; 0 = one block per unit).  Measured (B=16 @256^2, 100 steps, two runs each): 74 -> 10.42-10.45 ms, 148 -> 10.29-10.30,
# 296 -> 10.17-10.26, 0 -> 10.14-10.18 (update off: 10.18-10.24): a shorter-lived launch disturbs the chain less than a
# bounded one that lingers.
UPDATE_SIDE_BLOCKS = int(os.environ.get("UNETK_UPDATE_SIDE_BLOCKS", "0"))
ALIGN = 64  # elements; keeps every slice 256-byte aligned


def _align(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


def _note(text):
    """Label the following C-ABI calls in a tools/timeline.py recording (no-op otherwise)."""
    if _lib.timeline is not None:
        _lib.timeline.note = text


def param_layout(named_params):
    """Offsets (in elements) of every parameter inside the flat buffers, in named_parameters() order."""
    offs, total = {}, 0
    for k, p in named_params:
        offs[k] = total
        total += _align(p.numel())
    return offs, total


class _ConvBN:
    """One conv3x3 + BatchNorm record (tensor views are filled in by Engine._build_views)."""

    def __init__(self, name, conv, bn, cin1, cin2, cout, first=False):
        self.name, self.conv, self.bn = name, conv, bn
        self.cin1, self.cin2, self.cout, self.first = cin1, cin2, cout, first


class _ConvT:
    def __init__(self, name, mod, cin, cout):
        self.name, self.mod, self.cin, self.cout = name, mod, cin, cout


class _Head:
    def __init__(self, name, mod, n_out, sigmoid):
        self.name, self.mod, self.n_out, self.sigmoid = name, mod, n_out, sigmoid


class Plan:
    """Buffers for one input shape."""

    def __init__(self, eng, n, h, w, device):
        self.n, self.h, self.w = n, h, w
        bf = torch.bfloat16

        def act(lvl, c):
            return torch.empty(n, h >> lvl, w >> lvl, c, dtype=bf, device=device)

        self.enc = []
        self.enc_c = [c2.cout for _, c2 in eng.encoder]       # 64..1024 (512 at the bottom for bilinear decoders)
        for l in range(5):
            c = self.enc_c[l]
            d = {"y1": act(l, c), "a1": act(l, c), "y2": act(l, c), "a2": act(l, c)}
            if l < 4:
                d["p"] = act(l + 1, c)
                d["idx"] = torch.empty(n, h >> (l + 1), w >> (l + 1), c, dtype=torch.uint8, device=device)
            self.enc.append(d)
        self.dec = []
        for dec_stages, _ in eng.decoders:
            stages = []
            for k in range(1, 5):
                lv = 4 - k
                ct, c1, c2 = dec_stages[k - 1]
                # the transposed conv doubles the level below: up to one row / column short of the skip when
                # the input size is not a multiple of 16 (the reference zero-pads it, src/model.py:57-61: the
                # consumer kernels read it as zero beyond its extent instead)
                up = torch.empty(n, 2 * (h >> (lv + 1)), 2 * (w >> (lv + 1)), c1.cin2, dtype=bf, device=device)
                stages.append({"up": up, "y1": act(lv, c1.cout), "a1": act(lv, c1.cout), "y2": act(lv, c2.cout),
                               "a2": act(lv, c2.cout)})
            self.dec.append(stages)
        self._act = act
        self.all_even = all((h >> l) % 2 == 0 and (w >> l) % 2 == 0 for l in range(4))
        self.n_dec = len(eng.decoders)
        # the image as NHWC bf16 zero-padded to 64 channels (first conv on the tensor cores); the padding
        # channels are written once, here
        self.x64 = ops.zero_(act(0, 64)) if eng.n_channels <= 8 else None
        self.grads_ready = False
        self.bnctx = {}
        self.fused = set()
        self.drop_mask = None
        self.head_fused = False   # the last training forward applied the final BatchNorm inside the head kernels
        self.chains_flow = False  # train_step(): the decoder streams were not joined between forward and backward
        self.fused_loss = None    # train_step(): {"mask", "cfg", "work", "out3", "want_outs"} while its launches are enqueued
        self.m_in = None          # static copy of the mask (train_step graph replays read it)
        self.out3 = None          # static (total, recon, seg) of train_step
        self.static_outs = None   # outputs owned by the forward graph (backward reads them)
        self.graphs = {}          # "train" / "eval" / "bwd" -> _GraphState
        self.x_in = None          # static copy of the input image (graph replays read it)
        self.g_in = None          # static copies of the output gradients

    def alloc_grads(self):
        if self.grads_ready:
            return
        act = self._act
        self.d_skip = [act(l, self.enc_c[l]) for l in range(5)]   # d(a2) of every encoder level
        # skip gradients of the second, third... decoder when the decoders run on parallel streams
        self.d_skipx = [[act(l, self.enc_c[l]) for l in range(4)] for _ in range(self.n_dec - 1)]
        # scratch per backward chain (one per decoder; the encoder continues on chain 0), all of the
        # largest (full resolution, 64 channel) size: d_t = running decoder gradient d(a2) of the current
        # stage, d_a1 = gradient wrt a DoubleConv's inner activation, d_y = gradient wrt a raw conv output,
        # d_up = gradient wrt an up-sampled tensor.  d_y / d_up have GRAD_SLOTS slots each: the weight gradient
        # that reads slot i runs on a side stream while the chain already fills slot i^1.
        self.scratch = [{"d_t": act(0, 64), "d_a1": act(0, 64), "d_y": [act(0, 64) for _ in range(GRAD_SLOTS)],
                         "d_up": [act(0, 64) for _ in range(GRAD_SLOTS)]} for _ in range(self.n_dec)]
        self.d_p = act(1, 64)       # gradient wrt a pooled tensor (max size)
        self.fc_scratch = torch.empty(64 * 9 * 64, dtype=torch.float32, device=self.d_p.device)   # first conv dW[co][tap][64]
        self.grads_ready = True

    @staticmethod
    def view(buf, like):
        return buf.view(-1)[: like.numel()].view(like.shape)


class _GraphState:
    """Capture state of one launch sequence of a plan: call 1 runs eagerly (warm-up), call 2 captures,
    later calls replay."""

    def __init__(self):
        self.calls = 0
        self.graph = None
        self.result = None
        self.failed = False
        self.n_calls = self.n_kernels = 0      # C-ABI calls / kernel launches one replay stands for

    def run(self, body, after=None):
        """body() enqueues the sequence on the current stream (and side streams that rejoin it) and
        returns tensors that live as long as the graph.  Returns (result, replayed)."""
        self.calls += 1
        if self.graph is not None:
            self.graph.replay()
            _lib.add_replayed(self.n_calls, self.n_kernels)
            return self.result, True
        if self.calls >= 2 and not self.failed:
            g = torch.cuda.CUDAGraph()
            c0, k0 = _lib.launch_count(), _lib.kernel_launch_count()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    result = body()
            except Exception as e:      # stay on the eager path: same kernels, launched one by one
                self.failed = True
                torch.cuda.synchronize()
                import warnings
                warnings.warn(f"unetk: CUDA graph capture failed ({type(e).__name__}: {e}); running eagerly")
                return body(), False
            self.graph, self.result = g, result
            self.n_calls, self.n_kernels = _lib.launch_count() - c0, _lib.kernel_launch_count() - k0
            g.replay()                  # capture does not execute (the calls above were counted once already)
            return result, True
        return body(), False


class _Chain:
    """One backward chain (a decoder, then for chain 0 the encoder): its stream, the side stream of its
    weight gradients and its scratch tensors."""

    def __init__(self, stream, wstream, scratch):
        self.stream, self.wstream = stream, wstream
        self.d_t, self.d_a1 = scratch["d_t"], scratch["d_a1"]
        self.d_y, self.d_up = scratch["d_y"], scratch["d_up"]
        self.free_y, self.free_up = [None] * len(self.d_y), [None] * len(self.d_up)   # events after which a slot may be overwritten
        self.slot_y = self.slot_up = 0
        self.last_w = None                                       # last event recorded on the side stream

    def mark(self):
        """Event on the chain's stream at this point (an earlier release point for side()), or None."""
        if self.wstream is None or not WGRAD_EARLY:
            return None
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return ev

    def side(self, fn, after=None):
        """Run the weight-gradient launches of `fn` on the side stream, after everything enqueued on the
        chain's stream so far (the data gradient of the same layer included) or after the event `after`."""
        if self.wstream is None:
            fn()
            return None
        ev = after
        if ev is None:
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.wstream.wait_event(ev)
        with torch.cuda.stream(self.wstream):
            fn()
            done = torch.cuda.Event()
            done.record(self.wstream)
        self.last_w = done
        return done

    def _take(self, bufs, free, slot, like):
        if free[slot] is not None:
            self.stream.wait_event(free[slot])
            free[slot] = None
        return slot, Plan.view(bufs[slot], like)

    def take_y(self, like):
        self.slot_y = (self.slot_y + 1) % len(self.d_y)
        return self._take(self.d_y, self.free_y, self.slot_y, like)

    def take_up(self, like):
        self.slot_up = (self.slot_up + 1) % len(self.d_up)
        return self._take(self.d_up, self.free_up, self.slot_up, like)


class Engine:
    def __init__(self, module, arch, n_channels, head_specs):
        """head_specs: list of (suffix, n_out, sigmoid) — one decoder each."""
        self.module = module
        self.arch = arch
        self.n_channels = n_channels
        self.device = None
        self.plans = {}
        self.flat_p = None
        self._sig = None
        self.ddp = None           # set by distributed.attach()
        self.fuse_check = []      # (layer, relative error) records of UNETK_FUSE_CHECK=1
        self.wstreams = []        # side streams of the weight gradients, one per backward chain (OVERLAP_WGRAD)
        self.dstreams = []        # side streams of the second, third... decoder (PARALLEL_DECODERS)
        self.hstream = None       # high-priority stream of the critical chain (PRIORITY_STREAMS)
        self.pstream = None       # side stream packing the decoders' weights (SPLIT_PACK)
        self.ustream = None       # side stream un-packing the decoders' gradients during the encoder's backward
        self.sync_bn = None       # process group for synchronised BatchNorm statistics (distributed.attach(sync_bn=True))
        self.bottleneck_dropout = 0.0   # SegmentationUNet: nn.Dropout2d(p) on the bottleneck (src/model.py:130,146)
        self.pack_done0 = self.pack_done4 = self.pack_done = self.zero_done = None
        self._pack_pending = None
        self.weights_dirty = True
        self.fold_dirty = True
        # bumped by EVERY forward (any shape, train or eval): the BatchNorm scale / shift / mean / invstd, the
        # statistics and the packed weights are per engine, not per plan, so any later forward invalidates the
        # saved state a pending backward needs
        self.generation = 0
        self.staging_clean = False     # gradient staging buffer zeroed by the last training forward
        self.gstage_zero = False       # ... or by the fused parameter update of the last train_step (which zeroes what it reads)
        self._upd = None               # train_step(update=...): the optimizer state its backward pass updates in place
        m = module
        self.convbn = []
        self.encoder = []
        dcs = [("inc", m.inc)] + [(f"down{k}.maxpool_conv.1", getattr(m, f"down{k}").maxpool_conv[1]) for k in range(1, 5)]
        for l, (name, dc) in enumerate(dcs):
            k1, k2 = dc.double_conv[0], dc.double_conv[3]
            c1 = _ConvBN(f"{name}.double_conv.0", k1, dc.double_conv[1], k1.in_channels, 0, k1.out_channels, first=(l == 0))
            c2 = _ConvBN(f"{name}.double_conv.3", k2, dc.double_conv[4], k2.in_channels, 0, k2.out_channels)
            self.encoder.append((c1, c2))
            self.convbn += [c1, c2]
        self.decoders = []
        for suffix, n_out, sigmoid in head_specs:
            stages = []
            t_c = self.encoder[4][1].cout            # channels of the tensor entering the stage
            for k in range(1, 5):
                up = getattr(m, f"up{k}{suffix}")
                skip_c = self.encoder[4 - k][1].cout
                k1, k2 = up.conv.double_conv[0], up.conv.double_conv[3]
                if getattr(up, "bilinear", False):
                    # nn.Upsample(x2, bilinear, align_corners=True) keeps the channels; DoubleConv(in, out, in // 2)
                    ct, up_c = None, t_c
                else:
                    ct, up_c = _ConvT(f"up{k}{suffix}.up", up.up, t_c, t_c // 2), t_c // 2
                if k1.in_channels != skip_c + up_c:
                    raise _lib.UnetkError(f"up{k}{suffix}: conv expects {k1.in_channels} channels, the graph provides {skip_c} + {up_c}")
                c1 = _ConvBN(f"up{k}{suffix}.conv.double_conv.0", k1, up.conv.double_conv[1], skip_c, up_c, k1.out_channels)
                c2 = _ConvBN(f"up{k}{suffix}.conv.double_conv.3", k2, up.conv.double_conv[4], k2.in_channels, 0, k2.out_channels)
                stages.append((ct, c1, c2))
                self.convbn += [c1, c2]
                t_c = k2.out_channels
            head = _Head(f"outc{suffix}.conv", getattr(m, f"outc{suffix}").conv, n_out, sigmoid)
            self.decoders.append((stages, head))

    # ------------------------------------------------------------------------------------------
    # parameter storage
    # ------------------------------------------------------------------------------------------
    def _signature(self):
        ps = list(self.module.parameters())
        return (ps[0].data_ptr(), ps[-1].data_ptr(), ps[0].device)

    def ensure(self, device):
        """(Re)flatten parameters when the module moved (`.to()`, `.cuda()`) or on first use."""
        if self.flat_p is not None and self._sig == self._signature() and self.device == device:
            return
        if device.type != "cuda":
            raise _lib.UnetkError("this implementation has no CPU path: move the model and inputs to a CUDA (sm_100a) device")
        with torch.cuda.device(device):
            rc = _lib.get().unetk_check_device()
        if rc != 0:
            raise _lib.UnetkError(_lib.last_error())
        self.device = device
        named = list(self.module.named_parameters())
        self.param_names = [k for k, _ in named]
        self.params = [p for _, p in named]
        offs, total = param_layout(named)
        self.offsets, self.total = offs, total
        flat = torch.zeros(total, dtype=torch.float32, device=device)
        for k, p in named:
            v = flat[offs[k]: offs[k] + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
        self.flat_p = flat
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=device)
        self.gstage = torch.zeros(total, dtype=torch.float32, device=device)
        self._build_views(device)
        self.plans = {}
        self._sig = self._signature()
        self.weights_dirty = True
        self.fold_dirty = True

    def _slice(self, buf, key, shape=None):
        p = dict(zip(self.param_names, self.params))[key]
        v = buf[self.offsets[key]: self.offsets[key] + p.numel()]
        return v.view(shape if shape is not None else p.shape)

    def _build_views(self, device):
        chunk = _lib.get().unetk_permute_chunk_elems()
        pack = packing.PermuteTable(chunk)
        unpack_enc, unpack_dec = packing.PermuteTable(chunk), packing.PermuteTable(chunk)
        # conv3x3 weight packing / gradient un-packing in three groups: 0 = encoder levels 0-3 (a quarter of the
        # encoder's weights, needed first: packed on the caller's stream), 1 = the bottleneck level (three quarters)
        # and 2 = the decoders, both packed on a side stream under the first levels of the forward pass; the
        # un-packing mirrors it (groups 2 and 1 under the encoder's backward, only group 0 after its end)
        n_enc_layers = 2 * len(self.encoder)
        n_early = n_enc_layers - 2
        conv_pack_g = [packing.ConvTable(lambda cout, cin: (cout // 32) * (cin // 32)) for _ in range(3)]
        conv_unpack_g = [packing.ConvTable(lambda cout, cin: cout) for _ in range(3)]
        # the fused parameter update (un-pack + optimizer + re-pack in one pass, train_step) works on the segments the
        # backward pass finishes -- and the data-parallel reducer all-reduces -- one after the other: each decoder,
        # encoder levels 3-4 (nine tenths of the encoder's weights), the rest of the encoder
        upd_g = {tag: packing.UpdateTable(_lib.get().unetk_update_plain_chunk_elems())
                 for tag in [("dec", d) for d in range(len(self.decoders))] + [("enc", 3), ("enc", 0)]}

        def upd_of(li):
            if li >= n_enc_layers:
                return upd_g[("dec", (li - n_enc_layers) // 8)]
            return upd_g[("enc", 3) if li >= 6 else ("enc", 0)]
        pdict = dict(zip(self.param_names, self.params))
        wsize = 0

        def walloc(n):
            nonlocal wsize
            o = wsize
            wsize += _align(n)
            return o

        # small per-BN buffers: [stats 2C | sums 2C] regions are contiguous over layers (one memset each)
        tot_c = sum(L.cout for L in self.convbn)
        self.bn_stats_all = torch.zeros(2 * tot_c, dtype=torch.float32, device=device)
        self.bn_sums_all = torch.zeros(2 * tot_c, dtype=torch.float32, device=device)
        self.bn_aux = torch.zeros(4 * tot_c, dtype=torch.float32, device=device)
        co = 0
        wplan = []
        handled = set()
        for li, L in enumerate(self.convbn):
            c = L.cout
            grp = 0 if li < n_early else (1 if li < n_enc_layers else 2)
            conv_pack, conv_unpack = conv_pack_g[grp], conv_unpack_g[grp]
            unpack = unpack_enc if li < n_enc_layers else unpack_dec
            L.stats = self.bn_stats_all[2 * co: 2 * co + 2 * c]
            L.sums = self.bn_sums_all[2 * co: 2 * co + 2 * c]
            L.scale = self.bn_aux[4 * co: 4 * co + c]
            L.shift = self.bn_aux[4 * co + c: 4 * co + 2 * c]
            L.mean = self.bn_aux[4 * co + 2 * c: 4 * co + 3 * c]
            L.invstd = self.bn_aux[4 * co + 3 * c: 4 * co + 4 * c]
            co += c
            wk, gk, bk = L.name + ".weight", None, None
            bnname = L.name.rsplit(".", 1)[0] + "." + str(int(L.name.rsplit(".", 1)[1]) + 1)
            gk, bk = bnname + ".weight", bnname + ".bias"
            L.w = self._slice(self.flat_p, wk)
            L.gamma, L.beta = self._slice(self.flat_p, gk), self._slice(self.flat_p, bk)
            L.g_gamma, L.g_beta = self._slice(self.gstage, gk), self._slice(self.gstage, bk)
            L.rm, L.rv, L.nbt = L.bn.running_mean, L.bn.running_var, L.bn.num_batches_tracked
            cin = L.cin1 + L.cin2
            n = pdict[wk].numel()
            if L.first:
                L.g_w = self._slice(self.gstage, wk)            # written in OIHW directly
                unpack.add(self.offsets[wk], self.offsets[wk], *packing.identity(n))
                upd_of(li).add_plain(self.offsets[wk], n)       # (its packed copy: first_conv_pack after the update)
                wplan.append((L, "wf0", walloc(64 * 9 * 64), 64 * 9 * 64))   # Wf[64][9][64], zero beyond Cin
            else:
                L.g_w = self._slice(self.gstage, wk, (n,))      # [Cout][tap][Cin]
                of, od = walloc(n), walloc(n)
                conv_unpack.add(self.offsets[wk], self.offsets[wk], 0, L.cout, cin)
                conv_pack.add(self.offsets[wk], of, od, L.cout, cin)
                upd_of(li).add_conv(self.offsets[wk], of, od, L.cout, cin)
                wplan.append((L, "wf", of, n))
                wplan.append((L, "wd", od, n))
            for k in (gk, bk):
                unpack.add(self.offsets[k], self.offsets[k], *packing.identity(pdict[k].numel()))
                upd_of(li).add_plain(self.offsets[k], pdict[k].numel())
            handled.update((wk, gk, bk))
        unpack = unpack_dec
        for d, (stages, head) in enumerate(self.decoders):
            upd = upd_g[("dec", d)]
            for ct, _, _ in stages:
                if ct is None:                # bilinear stage: no parameters
                    continue
                wk, bk = ct.name + ".weight", ct.name + ".bias"
                n = pdict[wk].numel()
                ct.bias = self._slice(self.flat_p, bk)
                ct.g_w = self._slice(self.gstage, wk, (n,))
                ct.g_bias = self._slice(self.gstage, bk)
                of, od = walloc(n), walloc(n)
                pack.add(self.offsets[wk], of, *packing.convt_fwd_pack(ct.cin, ct.cout))
                pack.add(self.offsets[wk], od, *packing.convt_dgrad_pack(ct.cin, ct.cout))
                wplan.append((ct, "wf", of, n))
                wplan.append((ct, "wd", od, n))
                unpack.add(self.offsets[wk], self.offsets[wk], *packing.convt_grad_unpack(ct.cin, ct.cout))
                unpack.add(self.offsets[bk], self.offsets[bk], *packing.identity(ct.cout))
                upd.add_convt(self.offsets[wk], of, od, ct.cin, ct.cout)
                upd.add_plain(self.offsets[bk], ct.cout)
                handled.update((wk, bk))
            wk, bk = head.name + ".weight", head.name + ".bias"
            head.w = self._slice(self.flat_p, wk, (head.n_out, 64))
            head.b = self._slice(self.flat_p, bk)
            head.g_w = self._slice(self.gstage, wk, (head.n_out, 64))
            head.g_b = self._slice(self.gstage, bk)
            unpack.add(self.offsets[wk], self.offsets[wk], *packing.identity(head.n_out * 64))
            unpack.add(self.offsets[bk], self.offsets[bk], *packing.identity(head.n_out))
            upd.add_plain(self.offsets[wk], head.n_out * 64)
            upd.add_plain(self.offsets[bk], head.n_out)
            handled.update((wk, bk))
        missing = set(self.param_names) - handled
        if missing:
            raise _lib.UnetkError(f"engine does not cover parameters {sorted(missing)}")
        self.wpack = torch.zeros(max(wsize, ALIGN), dtype=torch.bfloat16, device=device)
        for obj, attr, off, n in wplan:
            setattr(obj, attr, self.wpack[off: off + n])
        self.pack_table = pack.finalize(device)
        self.unpack_tables = (unpack_enc.finalize(device), None, unpack_dec.finalize(device))
        self.conv_pack_tables = tuple(t.finalize(device) for t in conv_pack_g)
        self.conv_unpack_tables = tuple(t.finalize(device) for t in conv_unpack_g)
        self.update_tables = {tag: t.finalize(device) for tag, t in upd_g.items()}
        # a segment's update runs right behind the segment's all-reduce: it must stay inside the reduced range
        from .distributed import segment_ranges
        ranges = segment_ranges(self.param_names, self.offsets, self.total, len(self.decoders))
        for tag, t in upd_g.items():
            lo, hi = ranges[tag]
            if any(not (lo <= e[0] and e[0] + e[7] <= hi) for e in t.entries):
                raise _lib.UnetkError(f"update segment {tag} does not match the parameter range {lo}..{hi} of the same tag")
        self.hyper = torch.zeros(16, dtype=torch.float32, device=device)     # unetk_fused_update's hyper-parameters

    def mark_weights_dirty(self):
        self.weights_dirty = True
        self.fold_dirty = True

    @contextlib.contextmanager
    def _critical(self):
        """Run the enclosed launches on the engine's high-priority stream, forked from / joined to the caller's
        current stream (see PRIORITY_STREAMS)."""
        if not PRIORITY_STREAMS or ops.timing_enabled():
            yield
            return
        user = torch.cuda.current_stream(self.device)
        if self.hstream is None:
            self.hstream = torch.cuda.Stream(device=self.device, priority=-1)
        self.hstream.wait_stream(user)
        with torch.cuda.stream(self.hstream):
            yield
        user.wait_stream(self.hstream)

    def _version_sig(self):
        """Sum of autograd version counters: changes whenever torch code (load_state_dict, a stock
        optimizer) writes parameters or buffers in place.  Our own optimizer calls
        mark_weights_dirty() instead (its kernels write through raw pointers)."""
        v = 0
        for p in self.params:
            v += p._version
        for L in self.convbn:
            v += L.rm._version + L.rv._version
        return v

    def _pack_weights(self, split=True, zero_staging=False):
        """fp32 master weights -> bf16 kernel layouts.  Only the first conv's (tiny) weights on the caller's
        stream; the rest on a side stream, staged the way the forward pass needs it: encoder levels 0-3 at once
        (self.pack_done0, waited for by the second conv); the bottleneck level (self.pack_done4) and the decoders
        (self.pack_done) only after encoder level 1 (_pack_rest) -- packing is memory bound, and under the
        full-resolution BatchNorm passes of the first two levels it doubled their time, while the deeper levels are
        tensor bound and hide it.
        zero_staging (training forward): the side stream finally zeroes the gradient staging buffer and the
        BatchNorm-backward sums for the backward pass that follows (self.zero_done; 173 MB: 26 us that used to sit
        between the loss and the first backward kernel)."""
        self.pack_done0 = self.pack_done4 = self.pack_done = self.zero_done = None
        self._pack_pending = None
        if zero_staging:
            # the BatchNorm batch sums of the forward pass that follows: zeroed here, before the side stream is
            # forked (60 KB; queued behind the packing kernels' thousands of blocks it took 30 us instead of 3)
            ops.zero_(self.bn_stats_all)
        side = split and SPLIT_PACK and not ops.timing_enabled()
        if side and self.pstream is None:
            self.pstream = torch.cuda.Stream(device=self.device)
        dirty = self.weights_dirty
        if dirty:
            first = self.encoder[0][0]
            if side:
                self._fork_pack_stream()            # (the optimizer's parameter update is on the caller's stream)
            if self._first_tc():
                ops.first_conv_pack(first.w, first.wf0)
            with torch.cuda.stream(self.pstream) if side else contextlib.nullcontext():
                ops.pack_conv_weights(self.conv_pack_tables[0], self.flat_p, self.wpack)     # encoder levels 0-3 (tiled)
                if side:
                    self.pack_done0 = self._record(self.pstream)
        self._pack_pending = (side, dirty, zero_staging)
        if not side:
            self._pack_rest()
        self.weights_dirty = False

    def _fork_pack_stream(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.pstream.wait_event(ev)

    @staticmethod
    def _record(stream):
        e = torch.cuda.Event()
        e.record(stream)
        return e

    def _pack_rest(self):
        """Second half of _pack_weights: bottleneck level, decoders, transposed convs, staging memsets."""
        if self._pack_pending is None:
            return
        side, dirty, zero_staging = self._pack_pending
        self._pack_pending = None
        if not (dirty or zero_staging):
            return
        if side:
            self._fork_pack_stream()
        with torch.cuda.stream(self.pstream) if side else contextlib.nullcontext():
            if dirty:
                ops.pack_conv_weights(self.conv_pack_tables[1], self.flat_p, self.wpack)     # bottleneck level
                if side:
                    self.pack_done4 = self._record(self.pstream)
                ops.pack_conv_weights(self.conv_pack_tables[2], self.flat_p, self.wpack)     # decoders
                ops.permute_batched(self.pack_table, self.flat_p, self.wpack)                # transposed convs
                if side:
                    self.pack_done = self._record(self.pstream)
            if zero_staging:
                ops.zero_(self.gstage)
                ops.zero_(self.bn_sums_all)
                self.staging_clean = True
                if side:
                    self.zero_done = self._record(self.pstream)

    def _begin_update_step(self):
        """What is left of _pack_weights(zero_staging=True) when the previous train_step ended in fused updates:
        the packed weights are current and the staging buffer is zero; only the two small BatchNorm sum regions
        (forward statistics, backward sums) are cleared."""
        self.pack_done0 = self.pack_done4 = self.pack_done = self.zero_done = None
        self._pack_pending = None
        ops.zero_(self.bn_stats_all)
        ops.zero_(self.bn_sums_all)
        self.staging_clean = True

    def _wait_pack(self, which):
        """Make the current stream wait for a stage of the side-stream weight packing (see _pack_weights)."""
        ev = getattr(self, which)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            setattr(self, which, None)

    def _first_tc(self):
        return FIRST_FWD_TC and self.n_channels <= 8 and self.encoder[0][0].cout == 64

    def grad_views(self):
        """Fresh views of the flat gradient buffer, one per parameter (autograd takes ownership)."""
        out = []
        for k, p in zip(self.param_names, self.params):
            o = self.offsets[k]
            out.append(self.flat_g[o: o + p.numel()].view(p.shape))
        return out

    def plan_for(self, n, h, w):
        key = (n, h, w)
        pl = self.plans.pop(key, None)
        if pl is None:
            if h < 16 or w < 16:
                raise ValueError(f"input size {h}x{w} is too small: four 2x2 poolings need at least 16x16")
            while len(self.plans) >= MAX_PLANS:      # least recently used first (dicts keep insertion order)
                old = self.plans.pop(next(iter(self.plans)))
                old.graphs.clear()                   # drop the captured graphs (and their memory pools) with it
            pl = Plan(self, n, h, w, self.device)
        self.plans[key] = pl                         # (re)insert at the most-recently-used end
        return pl

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def _check_input(self, x):
        if x.dim() != 4 or x.shape[1] != self.n_channels:
            raise ValueError(f"expected input [N, {self.n_channels}, H, W], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise _lib.UnetkError("input must be a CUDA tensor (no CPU path)")
        if x.dtype != torch.float32:
            x = x.float()
        return x.contiguous()

    def _graphs_ok(self):
        return (USE_GRAPHS and not ops.timing_enabled() and self.bottleneck_dropout == 0.0 and not FUSE_CHECK
                and self.sync_bn is None)        # synchronised BatchNorm puts NCCL calls into the forward pass

    def forward(self, x, training):
        x = self._check_input(x)
        with torch.cuda.device(x.device):      # kernels go to the tensors' device, whatever the caller's current one
            return self._forward(x, training)

    def _forward(self, x, training):
        self.ensure(x.device)
        n, _, h, w = x.shape
        pl = self.plan_for(n, h, w)
        self.generation += 1
        sig = self._version_sig()
        if training or sig != getattr(self, "_vsig", None):
            self.mark_weights_dirty()
            self._vsig = sig
        if self._graphs_ok():
            outs = self._forward_graphed(pl, x, training)
        else:
            with self._critical() if training else contextlib.nullcontext():
                self._pack_weights(zero_staging=training)
                outs = self._forward_train(pl, x) if training else self._forward_eval(pl, x)
            user = torch.cuda.current_stream(self.device)
            for o in outs:
                o.record_stream(user)       # allocated on the engine's stream, consumed on the caller's
            pl.x = x
            pl.static_outs = None
        if training:
            self.fold_dirty = True   # running statistics moved
        return outs, pl, self.generation

    def _forward_graphed(self, pl, x, training):
        """Forward through a captured graph (see USE_GRAPHS): the image is copied into the plan's static
        input, the outputs are clones of the graph's static outputs."""
        if pl.x_in is None:
            pl.x_in = torch.empty_like(x)
        pl.x_in.copy_(x)
        pl.x = pl.x_in
        key = "train" if training else "eval"
        st = pl.graphs.get(key)
        if st is None:
            st = pl.graphs[key] = _GraphState()
        if training:
            def body():
                self.weights_dirty = True       # the optimizer moved the weights: re-pack inside the graph
                with self._critical():
                    self._pack_weights(zero_staging=True)
                    return self._forward_train(pl, pl.x_in)
        else:
            self._pack_weights(split=False)     # only when something changed the weights; not part of the graph
            self._fold()

            def body():
                return self._forward_eval(pl, pl.x_in, fold=False)
        outs, replayed = st.run(body)
        self.weights_dirty = False
        if training:
            self.staging_clean = True       # (a replay ran the captured memsets)
        pl.static_outs = outs if replayed else None
        if replayed:
            return [o.clone() for o in outs]
        user = torch.cuda.current_stream(self.device)
        for o in outs:
            o.record_stream(user)           # eager call: allocated on the engine's stream, consumed on the caller's
        return outs

    def _bn_count(self, L, count):
        """Pixel count behind L.stats; with synchronised BatchNorm the statistics are first summed over the
        ranks (global batch = exactly the single-process reference on the concatenated batch)."""
        if L.bn.momentum is None:
            raise NotImplementedError("BatchNorm momentum=None (cumulative moving average) is not supported: the "
                                      "reference uses nn.BatchNorm2d's default momentum 0.1 (src/model.py:15,18)")
        if self.sync_bn is not None:
            import torch.distributed as dist
            dist.all_reduce(L.stats, op=dist.ReduceOp.SUM, group=self.sync_bn[0])
            count = count * self.sync_bn[1]
        return count

    def _fin(self, L, count):
        """unetk_bn_finalize_args of layer L (cached: the pointers and the count are fixed per plan)."""
        mom = L.bn.momentum
        key = (count, L.bn.eps, mom, L.rm.data_ptr(), L.rv.data_ptr(), L.stats.data_ptr(), L.gamma.data_ptr())
        fin = L.__dict__.get("_fin")
        if fin is None or fin[0] != key:
            fin = (key, ops.bn_finalize_args(L.stats, L.gamma, L.beta, L.rm, L.rv, L.nbt, L.scale, L.shift, L.mean,
                                             L.invstd, count, L.bn.eps, mom))
            L._fin = fin
        return fin[1]

    def _bn_train(self, L, y, a, count, pool=None):
        count = self._bn_count(L, count)
        mom = L.bn.momentum
        if FUSE_BN_FINALIZE:
            fin = self._fin(L, count)
            if pool is None:
                ops.bnrelu_apply_fin(y, fin, a)
            elif y.shape[1] % 2 or y.shape[2] % 2:    # floor pooling leaves the last row / column outside every window
                ops.bnrelu_apply_fin(y, fin, a)
                ops.bnrelu_pool(a, None, None, None, pool[0], pool[1], apply=False)
            else:
                ops.bnrelu_pool_fin(y, fin, a, pool[0], pool[1])
            return
        ops.bn_finalize(L.stats, L.gamma, L.beta, L.rm, L.rv, L.nbt, L.scale, L.shift, L.mean, L.invstd,
                        count, L.bn.eps, mom)
        if pool is None:
            ops.bnrelu_apply(y, L.scale, L.shift, a)
        elif y.shape[1] % 2 or y.shape[2] % 2:
            ops.bnrelu_apply(y, L.scale, L.shift, a)
            ops.bnrelu_pool(a, None, None, None, pool[0], pool[1], apply=False)
        else:
            ops.bnrelu_pool(y, L.scale, L.shift, a, pool[0], pool[1], apply=True)

    def _forward_train(self, pl, x):
        # (self.bn_stats_all was zeroed by _pack_weights(zero_staging=True))
        n = pl.n
        for l, (c1, c2) in enumerate(self.encoder):
            e = pl.enc[l]
            _note(f"fwd enc{l}")
            if l == 4:
                self._wait_pack("pack_done4")
            count = n * (pl.h >> l) * (pl.w >> l)
            if l == 0 and self._first_tc():
                ops.image_to_nhwc64(x, pl.x64)
                ops.conv3x3_fwd(pl.x64, None, c1.wf0, e["y1"], stats=c1.stats, algo_cin=self.n_channels)
            elif l == 0:
                ops.first_conv_fwd(x, c1.w, e["y1"], c1.stats)
            else:
                ops.conv3x3_fwd(pl.enc[l - 1]["p"], None, c1.wf, e["y1"], stats=c1.stats)
            self._bn_train(c1, e["y1"], e["a1"], count)
            if l == 0:
                self._wait_pack("pack_done0")
            ops.conv3x3_fwd(e["a1"], None, c2.wf, e["y2"], stats=c2.stats)
            self._bn_train(c2, e["y2"], e["a2"], count, pool=(e["p"], e["idx"]) if l < 4 else None)
            if l == 1:
                self._pack_rest()
        pl.drop_mask = None
        if self.bottleneck_dropout > 0.0:
            # F.dropout2d(x5, p, training=True): one Bernoulli(1-p)/(1-p) factor per (image, channel), drawn from
            # torch's generator exactly as ATen's feature_dropout does (same stream as the reference on this
            # device); applied in place -- the bottleneck's ReLU mask in backward comes from y2, not from a2
            x5 = pl.enc[4]["a2"]
            keep = 1.0 - self.bottleneck_dropout
            pl.drop_mask = torch.empty(x5.shape[0], x5.shape[3], 1, 1, dtype=torch.float32, device=x5.device).bernoulli_(keep).div_(keep)
            ops.channel_scale(x5, pl.drop_mask)

        def stage(d, k):
            ct, c1, c2 = self.decoders[d][0][k]
            lv = 3 - k
            s = pl.dec[d][k]
            _note(f"fwd dec{d}.up{k + 1}")
            t = pl.dec[d][k - 1]["a2"] if k > 0 else pl.enc[4]["a2"]
            count = n * (pl.h >> lv) * (pl.w >> lv)
            if ct is None:
                ops.upsample2x_fwd(t, s["up"])
            else:
                ops.convt2x2_fwd(t, ct.wf, ct.bias, s["up"])
            ops.conv3x3_fwd(pl.enc[lv]["a2"], s["up"], c1.wf, s["y1"], stats=c1.stats)
            self._bn_train(c1, s["y1"], s["a1"], count)
            ops.conv3x3_fwd(s["a1"], None, c2.wf, s["y2"], stats=c2.stats)
            if k < 3 or not pl.head_fused:
                self._bn_train(c2, s["y2"], s["a2"], count)

        fl = pl.fused_loss          # set by train_step(): the heads also compute CombinedLoss and their own backward

        def head(d, out):
            hd = self.decoders[d][1]
            if not pl.head_fused:
                return ops.head_fwd(pl.dec[d][3]["a2"], hd.w, hd.b, out, hd.sigmoid)
            c2 = self.decoders[d][0][3][2]
            count = self._bn_count(c2, n * pl.h * pl.w)
            fin = self._fin(c2, count)
            y2 = pl.dec[d][3]["y2"]
            if fl is None:
                return ops.head_fwd_bn(y2, fin, hd.w, hd.b, out, hd.sigmoid)
            rw, sw, alpha, gamma = fl["cfg"]
            ops.head_loss_bwd(y2, fin, hd.w, hd.b, x if d == 0 else fl["mask"], out, Plan.view(pl.scratch[d]["d_t"], y2),
                              hd.g_w, hd.g_b, c2.sums, fl["work"], fl["out3"], d, rw, sw, alpha, gamma)

        pl.head_fused = fl is not None or (HEAD_FUSE_BN and FUSE_BN_FINALIZE
                                           and all(st[3][2].cout == 64 for st, _ in self.decoders))
        return self._run_decoders(pl, x, stage, head=head, want_outs=fl is None or fl["want_outs"])

    def _run_decoders(self, pl, x, stage, parallel=True, head=None, want_outs=True):
        """Decoder stages + heads.  The decoders of an AnomalyUNet are independent given the encoder, so in
        training they run on parallel streams, enqueued stage by stage: the memory-bound BatchNorm/ReLU
        passes of one decoder share the SMs with the tensor-bound convolutions of the other (12.0 -> 11.0 ms
        per step together with the backward chains).  Eval mode folds BatchNorm into the conv epilogue, is
        tensor bound throughout and measured slower on two streams (2.90 -> 3.05 ms): one stream."""
        n_dec = len(self.decoders)
        outs = [torch.empty(pl.n, hd.n_out, pl.h, pl.w, dtype=torch.float32, device=x.device) if want_outs else None
                for _, hd in self.decoders]
        streams = self._decoder_streams(n_dec if parallel else 1) * (1 if parallel else n_dec)
        s0 = streams[0]
        for which in ("pack_done0", "pack_done4", "pack_done"):     # decoder weights are packed on a side stream
            ev = getattr(self, which)
            if ev is not None:
                s0.wait_event(ev)
                setattr(self, which, None)
        if n_dec > 1 and streams[1] is not s0:
            ev = torch.cuda.Event()
            ev.record(s0)
            for st in streams[1:]:
                st.wait_event(ev)
        for k in range(5):
            if k == 4 and self.zero_done is not None:      # the heads of train_step() add into the zeroed staging buffers
                for st in {id(st): st for st in streams}.values():
                    st.wait_event(self.zero_done)
                self.zero_done = None
            for d in range(n_dec):
                with torch.cuda.stream(streams[d]):
                    if k < 4:
                        stage(d, k)
                    elif head is not None:
                        head(d, outs[d])
                    else:
                        hd = self.decoders[d][1]
                        ops.head_fwd(pl.dec[d][3]["a2"], hd.w, hd.b, outs[d], hd.sigmoid)
        if pl.fused_loss is not None and pl.fused_loss["flow"] and not want_outs and self._same_backward_streams(pl, streams):
            # train_step(): every decoder's backward chain continues on the stream its forward ran on -- no join
            # here, no fork there: the decoder that is ahead runs on into its backward pass instead of idling until
            # the other one's head is done (225 us at 16 x 256^2)
            pl.chains_flow = True
            return outs
        pl.chains_flow = False
        for st in streams[1:]:
            if st is not s0:
                ev = torch.cuda.Event()
                ev.record(st)
                s0.wait_event(ev)
        return outs

    def _backward_streams(self, pl):
        n_dec = len(self.decoders)
        # (odd sizes: the pool backward cannot add the second decoder's skip gradient to the last row / column)
        return (self._decoder_streams(n_dec) if n_dec == 2 and pl.all_even
                else [torch.cuda.current_stream(self.device)] * n_dec)

    def _same_backward_streams(self, pl, streams):
        return [s.cuda_stream for s in self._backward_streams(pl)] == [s.cuda_stream for s in streams]

    def _decoder_streams(self, n_dec):
        """[current stream, side streams...]; all the current stream when parallel decoders are off or the
        per-kernel timing instrumentation is on."""
        s0 = torch.cuda.current_stream(self.device)
        if n_dec == 1 or not PARALLEL_DECODERS or ops.timing_enabled():
            return [s0] * n_dec
        while len(self.dstreams) < n_dec - 1:
            self.dstreams.append(torch.cuda.Stream(device=self.device, priority=-1 if PRIORITY_STREAMS else 0))
        return [s0] + self.dstreams[: n_dec - 1]

    def _fold(self):
        if self.fold_dirty:
            for L in self.convbn:
                ops.bn_fold(L.gamma, L.beta, L.rm, L.rv, L.scale, L.shift, L.bn.eps)
            self.fold_dirty = False

    def _forward_eval(self, pl, x, fold=True):
        pl.head_fused = False
        if fold:
            self._fold()
        n = pl.n
        for l, (c1, c2) in enumerate(self.encoder):
            e = pl.enc[l]
            if l == 4:
                self._wait_pack("pack_done4")
            if l == 0 and self._first_tc():
                ops.image_to_nhwc64(x, pl.x64)
                ops.conv3x3_fwd(pl.x64, None, c1.wf0, e["a1"], scale=c1.scale, shift=c1.shift, relu=True,
                                algo_cin=self.n_channels)
            elif l == 0:
                ops.first_conv_fwd(x, c1.w, e["y1"], None)
                ops.bnrelu_apply(e["y1"], c1.scale, c1.shift, e["a1"])
            else:
                ops.conv3x3_fwd(pl.enc[l - 1]["p"], None, c1.wf, e["a1"], scale=c1.scale, shift=c1.shift, relu=True)
            if l == 0:
                self._wait_pack("pack_done0")
            ops.conv3x3_fwd(e["a1"], None, c2.wf, e["a2"], scale=c2.scale, shift=c2.shift, relu=True)
            if l < 4:
                ops.bnrelu_pool(e["a2"], None, None, None, e["p"], None, apply=False)
            if l == 1:
                self._pack_rest()

        def stage(d, k):
            ct, c1, c2 = self.decoders[d][0][k]
            lv = 3 - k
            s = pl.dec[d][k]
            t = pl.dec[d][k - 1]["a2"] if k > 0 else pl.enc[4]["a2"]
            if ct is None:
                ops.upsample2x_fwd(t, s["up"])
            else:
                ops.convt2x2_fwd(t, ct.wf, ct.bias, s["up"])
            ops.conv3x3_fwd(pl.enc[lv]["a2"], s["up"], c1.wf, s["a1"], scale=c1.scale, shift=c1.shift, relu=True)
            ops.conv3x3_fwd(s["a1"], None, c2.wf, s["a2"], scale=c2.scale, shift=c2.shift, relu=True)

        return self._run_decoders(pl, x, stage, parallel=False)

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def _bn_apply_bwd(self, pl, L, d_a, y, d_y):
        """BatchNorm+ReLU backward of layer L: per-channel sums (first pass) unless the kernel that
        finished d_a already produced them (see FUSE_BN_REDUCE), then dy / dgamma / dbeta."""
        if L.name in pl.fused:
            pl.fused.discard(L.name)
            if FUSE_CHECK:      # debug: the fused sums against the stand-alone reduction on the same d_a
                ref = ops.zero_(torch.empty_like(L.sums))
                ops.bn_bwd_reduce(d_a, y, L.scale, L.shift, L.mean, L.invstd, ref)
                err = ((L.sums - ref).abs().max() / (ref.abs().max() + 1e-20)).item()
                self.fuse_check.append((L.name, err))
        else:
            ops.bn_bwd_reduce(d_a, y, L.scale, L.shift, L.mean, L.invstd, L.sums)
        local = None
        if self.sync_bn is not None:
            # like torch's SyncBatchNorm: dy uses the sums over all ranks (AVG of the per-rank sums with the
            # per-rank pixel count = global sums / global count); dgamma / dbeta stay this rank's own sums and
            # are averaged with every other gradient
            import torch.distributed as dist
            local = L.sums.clone()
            dist.all_reduce(L.sums, op=dist.ReduceOp.AVG, group=self.sync_bn[0])
        ops.bn_bwd_apply(d_a, y, L.scale, L.shift, L.mean, L.invstd, L.sums, d_y, L.g_gamma, L.g_beta, param_sums=local)

    @staticmethod
    def _ctx(pl, L, y, kind="igemm"):
        """unetk_bn_ctx for the producer of layer L's activation gradient, or None when that producer does
        not fuse the reduction."""
        npix = y.numel() // y.shape[-1]
        if not (FUSE_BN_REDUCE == "all" or kind == "head" and FUSE_BN_REDUCE in ("head", "conv", "deep")
                or kind == "igemm" and (FUSE_BN_REDUCE == "conv" or FUSE_BN_REDUCE == "deep" and npix <= DEEP_PIXELS)):
            return None
        pl.fused.add(L.name)
        key = (L.name, y.data_ptr())
        c = pl.bnctx.get(key)
        if c is None:
            c = pl.bnctx[key] = ops.bn_ctx(y, L.scale, L.shift, L.mean, L.invstd, L.sums)
        return c

    def backward(self, pl, generation, grad_outs, outs):
        if generation != self.generation:
            raise RuntimeError("backward() after another forward() through the same model (any input shape, training "
                               "or eval): the saved activations / BatchNorm statistics of this graph were overwritten "
                               "(only one live graph per model is supported)")
        with torch.cuda.device(self.device):
            # Gradient accumulation (a second backward before the optimizer step, zero_grad(set_to_none=False)):
            # `.grad` is then already a view of flat_g, which the kernels below overwrite.  Keep the old sums,
            # hand autograd the new gradients in fresh storage and let its AccumulateGrad add them up.
            fs = self.flat_g.untyped_storage().data_ptr()
            accumulate = any(p.grad is not None and p.grad.untyped_storage().data_ptr() == fs for p in self.params)
            if accumulate:
                old = self.flat_g.clone()
            grads = self._backward(pl, grad_outs, outs)
            if accumulate:
                new = self.flat_g.clone()
                self.flat_g.copy_(old)
                grads = [new[self.offsets[k]: self.offsets[k] + p.numel()].view(p.shape)
                         for k, p in zip(self.param_names, self.params)]
            return grads

    # ------------------------------------------------------------------------------------------
    # fused training step (AnomalyUNet + CombinedLoss)
    # ------------------------------------------------------------------------------------------
    def can_fuse_step(self):
        """train_step() covers the two-decoder AnomalyUNet graph with 64-channel last layers, no dropout, local
        BatchNorm statistics and every parameter trainable."""
        return (self.arch == "anomaly_unet" and len(self.decoders) == 2 and self.bottleneck_dropout == 0.0
                and self.sync_bn is None and FUSE_BN_FINALIZE and not FUSE_CHECK
                and all(st[3][2].cout == 64 for st, _ in self.decoders)
                and self.decoders[0][1].sigmoid and self.decoders[1][1].sigmoid and self.decoders[1][1].n_out == 1
                and self.decoders[0][1].n_out in (1, 3)
                and all(p.requires_grad for p in self.module.parameters()))

    def train_step(self, x, mask, cfg, want_outs=False, update=None):
        """Forward, CombinedLoss (cfg = recon_weight, seg_weight, focal_alpha, focal_gamma) and backward of one
        batch as ONE launch sequence (one CUDA graph per input shape): what src/train_utils.py:122-132 does with
        `model(images)`, `criterion(...)` and `total_loss.backward()`, with each decoder's head, its loss term and
        the head's backward in a single kernel (csrc/head_fused.cu).  Sets every parameter's `.grad` (a view of
        the flat gradient buffer, replacing what was there) and returns (out3, outs): the device tensor
        (total, recon, seg) and, if asked for, the two sigmoid outputs.

        update (optim.FusedOptimizer.fused_update_args): the `optimizer.step()` of src/train_utils.py:133 becomes
        part of the same launch sequence -- as soon as the gradients of a parameter group are final (decoders, then
        the bottleneck level, then the rest of the encoder) ONE kernel writes the group's `.grad`, applies the
        optimizer, re-packs the new weights for the next step's kernels and zeroes the gradient staging range
        (unetk_fused_update), under the rest of the backward pass.  The next step then starts with its first
        convolution: no packing pass, no memset."""
        x = self._check_input(x)
        if not mask.is_cuda or mask.dtype != torch.float32:
            raise _lib.UnetkError("train_step: the mask must be a float32 CUDA tensor")
        mask = mask.contiguous()
        with torch.cuda.device(x.device):
            self.ensure(x.device)
            n, _, h, w = x.shape
            if tuple(mask.shape) != (n, 1, h, w):
                raise ValueError(f"train_step: mask must be {(n, 1, h, w)}, got {tuple(mask.shape)}")
            pl = self.plan_for(n, h, w)
            self.generation += 1
            sig = self._version_sig()
            if update is None or sig != getattr(self, "_vsig", None):
                self.mark_weights_dirty()   # (a fused update leaves freshly packed weights behind; torch code that
            self._vsig = sig                #  wrote a parameter in place since then shows in the version counters)
            pl.alloc_grads()
            if pl.out3 is None:
                pl.out3 = torch.zeros(3, dtype=torch.float32, device=x.device)
                pl.loss_work = ops.head_loss_workspace(x.device)
            cfg = tuple(float(v) for v in cfg)
            graphs = self._graphs_ok()
            if update is not None:
                # outside the captured sequence: packed weights only if something else changed the parameters, a
                # memset only if the previous backward pass was not followed by a fused update, this step's
                # hyper-parameters (step count, learning rate, 1 / world size)
                if self.weights_dirty:
                    self._pack_weights(split=False)
                if not self.gstage_zero:
                    ops.zero_(self.gstage)
                vals = list(update.values)
                vals[8] = 1.0 / self.ddp.world if self.ddp is not None else 1.0
                ops.set_floats(self.hyper, vals)
            self._upd = update
            upd_key = None if update is None else (update.mode, update.m.data_ptr(),
                                                   0 if update.v is None else update.v.data_ptr())

            whole = self.ddp is None or GRAPH_DDP       # NCCL all-reduces stay outside the capture by default
            # forward and backward enqueued back to back (both eager, or one capture): the decoder chains need no join
            flow = whole or not graphs

            def phases(xs, ms):
                pl.fused_loss = {"mask": ms, "cfg": cfg, "work": pl.loss_work, "out3": pl.out3, "want_outs": want_outs,
                                 "flow": flow}
                try:
                    with self._critical():
                        if update is None:
                            self._pack_weights(zero_staging=True)
                        else:
                            self._begin_update_step()
                        outs = self._forward_train(pl, xs)
                    yield outs
                    self.staging_clean = False
                    self._backward_body(pl, None, None)
                    yield None
                finally:
                    pl.fused_loss = None

            if not graphs:
                ph = phases(x, mask)
                outs = next(ph)
                pl.x = x
                next(ph)
                ph.close()
            else:
                if pl.x_in is None:
                    pl.x_in = torch.empty_like(x)
                if pl.m_in is None:
                    pl.m_in = torch.empty_like(mask)
                pl.x_in.copy_(x)
                pl.m_in.copy_(mask)
                pl.x = pl.x_in
                key = ("step" if whole else "step_fwd", cfg, want_outs, upd_key)
                st = pl.graphs.get(key)
                if st is None:
                    st = pl.graphs[key] = _GraphState()
                keep = {}

                def body():
                    ph = phases(pl.x_in, pl.m_in)
                    outs = next(ph)
                    if whole:
                        next(ph)
                        ph.close()
                    else:
                        keep["ph"] = ph
                    return outs
                outs, replayed = st.run(body)
                if not whole:
                    ph = keep.get("ph")
                    if ph is None:                          # replayed: only the backward half runs eagerly
                        pl.fused_loss = {"mask": pl.m_in, "cfg": cfg, "work": pl.loss_work, "out3": pl.out3, "want_outs": want_outs,
                                         "flow": False}
                        pl.chains_flow = False
                        try:
                            self.staging_clean = False
                            self._backward_body(pl, None, None)
                        finally:
                            pl.fused_loss = None
                    else:
                        next(ph)
                        ph.close()
                if want_outs and replayed:
                    outs = [o.clone() for o in outs]
            self._upd = None
            self.weights_dirty = False
            self.fold_dirty = True
            self.staging_clean = False
            self.gstage_zero = update is not None
            for p, v in zip(self.params, self.grad_views()):
                p.grad = v
            return pl.out3.clone(), (outs if want_outs else None)

    def _backward(self, pl, grad_outs, outs):
        if not self.staging_clean:          # not preceded by a training forward of this engine (cannot happen
            ops.zero_(self.gstage)          # through autograd; kept for direct callers)
            ops.zero_(self.bn_sums_all)
        self.staging_clean = False
        self.gstage_zero = False
        self._upd = None                    # (left set if a train_step(update=...) raised half way)
        pl.alloc_grads()
        graphed = (self._graphs_ok() and pl.static_outs is not None and (self.ddp is None or GRAPH_DDP))
        if not graphed:
            grads = []
            for d in range(len(self.decoders)):
                g = grad_outs[d]
                if g is None:
                    g = ops.zero_(torch.empty_like(outs[d]))
                grads.append(g.contiguous().float())
            return self._backward_body(pl, grads, outs)
        # graph replay: output gradients go through static buffers, the heads read the forward graph's outputs
        if pl.g_in is None:
            pl.g_in = [torch.empty_like(o) for o in pl.static_outs]
        for d, g in enumerate(grad_outs):
            if g is None:
                ops.zero_(pl.g_in[d])
            else:
                pl.g_in[d].copy_(g)
        st = pl.graphs.get("bwd")
        if st is None:
            st = pl.graphs["bwd"] = _GraphState()
        static_outs = pl.static_outs
        st.run(lambda: self._backward_body(pl, pl.g_in, static_outs))
        return self.grad_views()

    def _backward_body(self, pl, grads, outs):
        with self._critical():
            return self._backward_chains(pl, grads, outs)

    def _backward_chains(self, pl, grads, outs):
        V = Plan.view
        ctx = self._ctx
        n_dec = len(self.decoders)
        streams = self._backward_streams(pl)
        s0 = streams[0]
        parallel = n_dec == 2 and streams[1] is not s0
        overlap = OVERLAP_WGRAD and not ops.timing_enabled()
        while overlap and len(self.wstreams) < n_dec:
            self.wstreams.append(torch.cuda.Stream(device=self.device))
        chains = [_Chain(streams[d], self.wstreams[d if parallel else 0] if overlap else None, pl.scratch[d])
                  for d in range(n_dec)]
        if parallel and not (pl.fused_loss is not None and pl.chains_flow):
            ev = torch.cuda.Event()
            ev.record(s0)
            streams[1].wait_event(ev)
        enc4 = self.encoder[4][1]
        bottleneck_done = None          # event: the previous decoder's contribution to d_skip[4] is stored
        skip_written = [False] * 5

        def decoder_stage(d, k, ch):
            nonlocal bottleneck_done
            stages, head = self.decoders[d]
            ct, c1, c2 = stages[k]
            lv = 3 - k
            s = pl.dec[d][k]
            _note(f"bwd dec{d}.up{k + 1}")
            if k == 3 and pl.fused_loss is not None:
                pl.fused.add(c2.name)       # head_loss_bwd left dA in the chain's d_t and the BatchNorm sums in c2.sums
            elif k == 3:
                bnc = ctx(pl, c2, s["y2"], kind="head")
                if bnc is None and pl.head_fused:       # no fused reduction, but the activation comes from y2
                    bnc = pl.bnctx.get((c2.name, "y-only"))
                    if bnc is None:
                        bnc = pl.bnctx[(c2.name, "y-only")] = ops.bn_ctx(s["y2"], c2.scale, c2.shift, c2.mean, c2.invstd, None)
                ops.head_bwd(None if pl.head_fused else s["a2"], head.w, grads[d], outs[d], V(ch.d_t, s["a2"]), head.g_w,
                             head.g_b, head.sigmoid, bn=bnc)
            d_t = V(ch.d_t, s["a2"])
            d_a1 = V(ch.d_a1, s["a1"])
            skip = pl.enc[lv]["a2"]
            # second conv of the DoubleConv
            iy, d_y = ch.take_y(s["y2"])
            self._bn_apply_bwd(pl, c2, d_t, s["y2"], d_y)
            mk = ch.mark()
            ops.conv3x3_dgrad(d_y, c2.wd, d_a1, bn=ctx(pl, c1, s["y1"]))
            ch.free_y[iy] = ch.side(lambda: ops.conv3x3_wgrad(d_y, s["a1"], None, c2.g_w), mk)
            # first conv: two sources (skip, up)
            iy, d_y1 = ch.take_y(s["y1"])
            iu, d_up = ch.take_up(s["up"])
            self._bn_apply_bwd(pl, c1, d_a1, s["y1"], d_y1)
            mk = ch.mark()
            bias_grad = ct.g_bias if ct is not None else None     # transposed-conv bias gradient = column sums of d_up
            if parallel and d > 0:      # own skip-gradient buffer: summed by the encoder's pool backward
                ops.conv3x3_dgrad(d_y1, c1.wd, pl.d_skipx[d - 1][lv], acc1=False, dx2=d_up, acc2=False, stats=bias_grad)
            else:
                ops.conv3x3_dgrad(d_y1, c1.wd, pl.d_skip[lv], acc1=skip_written[lv], dx2=d_up, acc2=False, stats=bias_grad)
                skip_written[lv] = True
            ch.free_y[iy] = ch.side(lambda: ops.conv3x3_wgrad(d_y1, skip, s["up"], c1.g_w), mk)
            # transposed conv
            t_in = pl.dec[d][k - 1]["a2"] if k > 0 else pl.enc[4]["a2"]
            mk = ch.mark()
            if ct is None:              # bilinear up-sampling: no parameters, gather backward
                if k > 0:
                    ops.upsample2x_bwd(d_up, V(ch.d_t, t_in), accumulate=False)
                else:
                    if bottleneck_done is not None:
                        ch.stream.wait_event(bottleneck_done)
                    ops.upsample2x_bwd(d_up, pl.d_skip[4], accumulate=skip_written[4])
                    skip_written[4] = True
                    if parallel:
                        bottleneck_done = torch.cuda.Event()
                        bottleneck_done.record(ch.stream)
                return
            if k > 0:
                prev = stages[k - 1][2]
                ops.convt2x2_dgrad(d_up, ct.wd, V(ch.d_t, t_in), acc=False, bn=ctx(pl, prev, pl.dec[d][k - 1]["y2"]))
            else:
                # the bottleneck gradient is the sum over the decoders, complete after decoder 0 (processed last)
                if bottleneck_done is not None:
                    ch.stream.wait_event(bottleneck_done)
                ops.convt2x2_dgrad(d_up, ct.wd, pl.d_skip[4], acc=skip_written[4],
                                   bn=ctx(pl, enc4, pl.enc[4]["y2"]) if d == 0 and pl.drop_mask is None else None)
                skip_written[4] = True
                if parallel:
                    bottleneck_done = torch.cuda.Event()
                    bottleneck_done.record(ch.stream)
            ch.free_up[iu] = ch.side(lambda: ops.convt2x2_wgrad(t_in, d_up, ct.g_w), mk)

        # decoders: enqueued stage by stage so that parallel chains advance together
        for k in reversed(range(4)):
            for d in reversed(range(n_dec)):
                with torch.cuda.stream(chains[d].stream):
                    decoder_stage(d, k, chains[d])
        upd = self._upd is not None
        side_updates = upd and self.ddp is None and overlap     # (with a reducer the updates follow the all-reduces)
        if side_updates and self.ustream is None:
            self.ustream = torch.cuda.Stream(device=self.device)
        for d in reversed(range(n_dec)):
            ch = chains[d]
            done = None
            if ch.stream is not s0:
                done = torch.cuda.Event()
                done.record(ch.stream)
                s0.wait_event(done)
            if self.ddp is not None:
                self.ddp.segment_ready(self, ("dec", d), [e for e in (done, ch.last_w) if e is not None],
                                       then=(lambda d=d: self._update_segment(("dec", d), side=True)) if upd else None)
            elif side_updates:
                # this decoder's gradients are final: its parameters are updated (and re-packed) on a side stream
                # while the other decoder and the encoder are still in their backward pass
                self.ustream.wait_event(done if done is not None else self._record(s0))
                if ch.last_w is not None:
                    self.ustream.wait_event(ch.last_w)
                with torch.cuda.stream(self.ustream):
                    self._update_segment(("dec", d), side=True)
            elif upd:
                self._update_segment(("dec", d))
        # the decoders' gradients are final: re-lay them out on a side stream under the encoder's backward
        # (without data parallelism; with it the buckets are still being all-reduced)
        early_unpack = None
        early_bottleneck = False
        if self.ddp is None and overlap and not upd:
            if self.ustream is None:
                self.ustream = torch.cuda.Stream(device=self.device)
            ev = torch.cuda.Event()
            ev.record(s0)
            self.ustream.wait_event(ev)
            for c in chains:
                if c.last_w is not None:
                    self.ustream.wait_event(c.last_w)
            with torch.cuda.stream(self.ustream):
                self._unpack(2, 1.0)
                early_unpack = torch.cuda.Event()
                early_unpack.record(self.ustream)
        ch = chains[0]
        if pl.drop_mask is not None:            # Dropout2d backward: the same per-(image, channel) factors
            ops.channel_scale(pl.d_skip[4], pl.drop_mask)
        for l in reversed(range(5)):
            c1, c2 = self.encoder[l]
            e = pl.enc[l]
            _note(f"bwd enc{l}")
            d_a1 = V(ch.d_a1, e["a1"])
            iy, d_y = ch.take_y(e["y2"])
            self._bn_apply_bwd(pl, c2, pl.d_skip[l], e["y2"], d_y)
            mk = ch.mark()
            ops.conv3x3_dgrad(d_y, c2.wd, d_a1, bn=ctx(pl, c1, e["y1"]))
            ch.free_y[iy] = ch.side(lambda d_y=d_y, e=e, c2=c2: ops.conv3x3_wgrad(d_y, e["a1"], None, c2.g_w), mk)
            iy, d_y = ch.take_y(e["y1"])
            self._bn_apply_bwd(pl, c1, d_a1, e["y1"], d_y)
            mk = ch.mark()
            if l == 0:
                def first_wgrad(d_y=d_y, c1=c1):
                    if self.n_channels <= 8 and FIRST_WGRAD_TC:
                        if not self._first_tc():                 # otherwise the forward pass already wrote x64
                            ops.image_to_nhwc64(pl.x, pl.x64)
                        ops.first_conv_wgrad_tc(pl.x64, d_y, pl.fc_scratch, c1.g_w)
                    else:
                        ops.first_conv_wgrad(pl.x, d_y, c1.g_w)
                ch.free_y[iy] = ch.side(first_wgrad, mk)
            else:
                src = pl.enc[l - 1]["p"]
                d_p = V(pl.d_p, src)
                ops.conv3x3_dgrad(d_y, c1.wd, d_p)
                # the skip gradient of level l-1 is complete after this accumulation
                ops.maxpool_bwd(d_p, pl.enc[l - 1]["idx"], pl.d_skip[l - 1], accumulate=skip_written[l - 1],
                                extra=pl.d_skipx[0][l - 1] if parallel else None,
                                bn=ctx(pl, self.encoder[l - 1][1], pl.enc[l - 1]["y2"], kind="pool"))
                skip_written[l - 1] = True
                ch.free_y[iy] = ch.side(lambda d_y=d_y, src=src, c1=c1: ops.conv3x3_wgrad(d_y, src, None, c1.g_w), mk)
            if self.ddp is not None and l in (3, 0):
                self.ddp.segment_ready(self, ("enc", l), [ch.last_w] if ch.last_w is not None else [],
                                       then=(lambda l=l: self._update_segment(("enc", l), side=l == 3)) if upd else None)
            elif upd and l == 3:
                # levels 3-4 hold nine tenths of the encoder's weights: updated under the rest of the encoder's backward
                if side_updates:
                    self.ustream.wait_event(self._record(ch.stream))      # (the data gradients that read the packed weights)
                    if ch.last_w is not None:
                        self.ustream.wait_event(ch.last_w)
                    with torch.cuda.stream(self.ustream):
                        self._update_segment(("enc", 3), side=True)
                else:
                    if ch.last_w is not None:
                        ch.stream.wait_event(ch.last_w)
                    self._update_segment(("enc", 3))
            if l == 4 and early_unpack is not None and ch.last_w is not None:
                # the bottleneck level holds three quarters of the encoder's weights: un-pack its gradients under
                # the rest of the encoder's backward too
                self.ustream.wait_event(ch.last_w)
                with torch.cuda.stream(self.ustream):
                    self._unpack(1, 1.0)
                    early_unpack = torch.cuda.Event()
                    early_unpack.record(self.ustream)
                early_bottleneck = True
        for c in chains:
            if c.last_w is not None:
                s0.wait_event(c.last_w)       # every weight gradient is in the staging buffer
        _note("bwd unpack")
        scale = 1.0
        if self.ddp is not None:
            scale = self.ddp.finish(self)     # (with fused updates: the last segment's update ran behind its all-reduce)
        if upd:
            if side_updates:
                s0.wait_event(self._record(self.ustream))
            if self.ddp is None:
                self._update_segment(("enc", 0))
            return self.grad_views()
        if early_unpack is not None:
            s0.wait_event(early_unpack)       # the decoders' gradients were re-laid-out under the encoder's backward
        else:
            self._unpack(2, scale)
        if not early_bottleneck:
            self._unpack(1, scale)
        self._unpack(0, scale)
        return self.grad_views()

    def _update_segment(self, tag, side=False):
        """train_step(update=...): the gradients of segment `tag` are final (and all-reduced): ONE kernel writes its
        `.grad`, applies the optimizer, re-packs the new weights and zeroes the staging range (unetk_fused_update).
        side: the launch runs beside the rest of the backward pass (optional grid bound, UPDATE_SIDE_BLOCKS)."""
        u = self._upd
        ops.fused_update(self.update_tables[tag], self.flat_p, self.flat_g, u.m, u.v, self.gstage, self.wpack, self.hyper,
                         u.mode, max_blocks=UPDATE_SIDE_BLOCKS if side else 0)
        if tag == ("enc", 0) and self._first_tc():
            first = self.encoder[0][0]
            ops.first_conv_pack(first.w, first.wf0)

    def _unpack(self, part, scale):
        """Staged gradients ([Cout][tap][Cin] conv weights, ...) -> `.grad` in the reference layouts, for the
        parameters of encoder levels 0-3 (part 0), of the bottleneck level's convolutions (part 1) or of the
        decoders (part 2)."""
        if self.conv_unpack_tables[part].n_desc:
            ops.unpack_conv_grads(self.conv_unpack_tables[part], self.gstage, self.flat_g, scale)
        if self.unpack_tables[part] is not None:
            ops.permute_batched(self.unpack_tables[part], self.gstage, self.flat_g, scale)
