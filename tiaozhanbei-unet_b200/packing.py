"""Descriptor tables for unetk_permute_batched: one kernel launch re-lays-out many tensors.

Used for (a) fp32 reference-layout weights -> bf16 kernel layouts and (b) staged fp32 gradients
(conv weights in [Cout][tap][Cin]) -> the reference layouts that `.grad` must have.

Descriptor (64 bytes, mirrors `PermuteDesc` in csrc/optim.cu):
    int64 in_off, out_off, s0, s1, s2; int32 d0, d1, d2, flip1, chunk_begin, pad
    out[out_off + ((i0*d1 + i1)*d2 + i2)] = scale * in[in_off + i0*s0 + f(i1)*s1 + i2*s2]
"""
from __future__ import annotations

import numpy as np
import torch

DESC_DTYPE = np.dtype([("in_off", "<i8"), ("out_off", "<i8"), ("s0", "<i8"), ("s1", "<i8"), ("s2", "<i8"),
                       ("d0", "<i4"), ("d1", "<i4"), ("d2", "<i4"), ("flip1", "<i4"),
                       ("chunk_begin", "<i4"), ("pad", "<i4")])
assert DESC_DTYPE.itemsize == 64


class PermuteTable:
    def __init__(self, chunk_elems: int):
        self.chunk_elems = chunk_elems
        self.entries = []
        self.device_descs = None
        self.n_desc = 0
        self.n_chunks = 0

    def add(self, in_off, out_off, dims, strides, flip1=False):
        d0, d1, d2 = dims
        s0, s1, s2 = strides
        self.entries.append((in_off, out_off, s0, s1, s2, d0, d1, d2, int(flip1)))

    def numpy(self):
        arr = np.zeros(len(self.entries), dtype=DESC_DTYPE)
        chunk = 0
        for i, (io, oo, s0, s1, s2, d0, d1, d2, fl) in enumerate(self.entries):
            arr[i] = (io, oo, s0, s1, s2, d0, d1, d2, fl, chunk, 0)
            chunk += (d0 * d1 * d2 + self.chunk_elems - 1) // self.chunk_elems
        return arr, chunk

    def finalize(self, device):
        arr, chunks = self.numpy()
        self.n_desc = len(self.entries)
        self.n_chunks = chunks
        self.device_descs = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
        return self


# ---- layout recipes (reference layout -> kernel layout), all as (dims, strides, flip) ----------
def conv_fwd_pack(cout, cin):
    """W[co][ci][ky][kx] (OIHW) -> Wf[co][tap][ci]"""
    return (cout, 9, cin), (cin * 9, 1, 9), False


def conv_dgrad_pack(cout, cin):
    """W[co][ci][tap] -> Wd[ci][tap'][co] with tap' = 8 - tap (180-degree rotated filter)"""
    return (cin, 9, cout), (9, 1, cin * 9), True


def convt_fwd_pack(cin, cout):
    """W[ci][co][ky][kx] -> Wt[tap][co][ci]  (rows tap*Cout+co, K = ci)"""
    return (4, cout, cin), (1, 4, cout * 4), False


def convt_dgrad_pack(cin, cout):
    """W[ci][co][tap] -> Wtd[ci][tap][co]  (rows ci, K = tap*Cout+co)"""
    return (cin, 4, cout), (cout * 4, 1, 4), False


def conv_grad_unpack(cout, cin):
    """dW[co][tap][ci] -> OIHW [co][ci][tap]"""
    return (cout, cin, 9), (9 * cin, 1, cin), False


def convt_grad_unpack(cin, cout):
    """dW[co][tap][ci] -> [ci][co][tap]"""
    return (cin, cout, 4), (1, 4 * cin, cin), False


def identity(n):
    return (1, 1, n), (0, 0, 1), False


# ---- tiled conv3x3 fast path (unetk_pack_conv_weights / unetk_unpack_conv_grads) -----------------
CONV_DESC_DTYPE = np.dtype([("src_off", "<i8"), ("dst_off", "<i8"), ("dst2_off", "<i8"), ("cout", "<i4"), ("cin", "<i4"),
                            ("block_begin", "<i4"), ("pad", "<i4", (7,))])
assert CONV_DESC_DTYPE.itemsize == 64


class ConvTable:
    """blocks_per_layer(cout, cin): 32x32 tiles for packing, one block per output channel for un-packing."""

    def __init__(self, blocks_per_layer):
        self.blocks_per_layer = blocks_per_layer
        self.entries = []
        self.device_descs = None
        self.n_desc = self.n_blocks = self.max_cin = 0

    def add(self, src_off, dst_off, dst2_off, cout, cin):
        if cout % 32 or cin % 32:
            raise ValueError("tiled conv packing needs channel counts that are multiples of 32")
        self.entries.append((src_off, dst_off, dst2_off, cout, cin))

    def finalize(self, device):
        arr = np.zeros(len(self.entries), dtype=CONV_DESC_DTYPE)
        blk = 0
        for i, (so, do, d2, cout, cin) in enumerate(self.entries):
            arr[i]["src_off"], arr[i]["dst_off"], arr[i]["dst2_off"] = so, do, d2
            arr[i]["cout"], arr[i]["cin"], arr[i]["block_begin"] = cout, cin, blk
            blk += self.blocks_per_layer(cout, cin)
            self.max_cin = max(self.max_cin, cin)
        self.n_desc, self.n_blocks = len(self.entries), blk
        self.device_descs = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
        return self


# ---- fused parameter update (unetk_fused_update) ------------------------------------------------------
UPDATE_DESC_DTYPE = np.dtype([("off", "<i8"), ("wf_off", "<i8"), ("wd_off", "<i8"), ("kind", "<i4"), ("cout", "<i4"),
                              ("cin", "<i4"), ("block_begin", "<i4"), ("pad", "<i4", (6,))])
assert UPDATE_DESC_DTYPE.itemsize == 64
KIND_PLAIN, KIND_CONV3X3, KIND_CONVT = 0, 1, 2


class UpdateTable:
    """Tensors one unetk_fused_update launch covers: `off` = the tensor's offset in the flat parameter / gradient /
    moment / staging buffers, `wf_off` / `wd_off` = where its two packed bf16 copies live in the weight pack."""

    def __init__(self, plain_chunk: int):
        self.plain_chunk = plain_chunk
        self.entries = []
        self.device_descs = None
        self.n_desc = self.n_blocks = self.max_end = 0

    def add_plain(self, off, n):
        self.entries.append((off, -1, -1, KIND_PLAIN, n, 0, (n + self.plain_chunk - 1) // self.plain_chunk, n))

    def add_conv(self, off, wf_off, wd_off, cout, cin):
        if cout % 32 or cin % 32:
            raise ValueError("fused update of a conv3x3 weight needs channel counts that are multiples of 32")
        self.entries.append((off, wf_off, wd_off, KIND_CONV3X3, cout, cin, (cout // 32) * (cin // 32), cout * cin * 9))

    def add_convt(self, off, wf_off, wd_off, cin, cout):
        if cout % 32 or cin % 32:
            raise ValueError("fused update of a transposed-conv weight needs channel counts that are multiples of 32")
        self.entries.append((off, wf_off, wd_off, KIND_CONVT, cout, cin, (cout // 32) * (cin // 32), cout * cin * 4))

    def numpy(self):
        arr = np.zeros(len(self.entries), dtype=UPDATE_DESC_DTYPE)
        blk = 0
        for i, (off, wf, wd, kind, cout, cin, blocks, _) in enumerate(self.entries):
            arr[i]["off"], arr[i]["wf_off"], arr[i]["wd_off"] = off, wf, wd
            arr[i]["kind"], arr[i]["cout"], arr[i]["cin"], arr[i]["block_begin"] = kind, cout, cin, blk
            blk += blocks
        return arr, blk

    def finalize(self, device):
        arr, blocks = self.numpy()
        self.n_desc, self.n_blocks = len(self.entries), blocks
        self.max_end = max((e[0] + e[7] for e in self.entries), default=0)
        self.device_descs = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
        return self
