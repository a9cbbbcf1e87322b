"""Optimizers returned by train_utils.get_optimizer: torch.optim.Optimizer subclasses whose
state_dict layout is the stock one (`state[i] = {step, exp_avg, exp_avg_sq}` for Adam/AdamW,
`{momentum_buffer}` for SGD, same param_groups keys), so optimizer checkpoints written by the
reference (src/utils.py:37-46, torch.optim.Adam built at src/train_utils.py:266) load here and
vice versa — but whose step is ONE launch of unetk_adam_step over the model's flat parameter buffer.

If the parameters are not (or no longer) views of an engine's flat buffer, or some gradients are
missing, the same kernel is launched once per tensor instead.  There is no torch-op fallback.

Inside train_utils.train_step (AnomalyUNet + CombinedLoss) step() is not called at all: fused_update_args()
hands the engine the flat moment buffers and this step's hyper-parameters, and the update runs inside the
backward pass, segment by segment (Engine.train_step(update=...), unetk_fused_update).
"""
from __future__ import annotations

import torch

from . import ops

_MODE = {"adam": 0, "adamw": 1, "sgd": 2}


class _FusedUpdate:
    """m / v: flat moment buffers laid out like the engine's flat parameter buffer; values: the nine floats of
    unetk_fused_update's hyper[] (index 8, the gradient scale, is filled in by the engine); rollback(): undo the
    step count this update was handed out with (the launch sequence it was meant for raised)."""
    __slots__ = ("m", "v", "mode", "values", "rollback")

    def __init__(self, m, v, mode, values, rollback=lambda: None):
        self.m, self.v, self.mode, self.values, self.rollback = m, v, mode, values, rollback


class FusedOptimizer(torch.optim.Optimizer):
    def __init__(self, params, kind, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8, momentum=0.9, engine=None):
        if kind not in _MODE:
            raise ValueError(f"Unknown optimizer: {kind}")
        self.kind = kind
        self._engine = engine
        self._flat = None     # (m, v) flat state buffers matching engine.flat_p
        if kind == "sgd":
            defaults = dict(lr=lr, momentum=momentum, dampening=0, weight_decay=weight_decay, nesterov=False,
                            maximize=False, foreach=None, differentiable=False, fused=None)
        else:
            defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                            foreach=None, capturable=False, differentiable=False, fused=None)
            if kind == "adam":
                defaults["decoupled_weight_decay"] = False
        super().__init__(params, defaults)

    # -- state helpers ---------------------------------------------------------------------------
    def _names(self):
        return ("momentum_buffer",) if self.kind == "sgd" else ("exp_avg", "exp_avg_sq")

    def _step_of(self, st):
        s = st.get("step", 0)
        return int(s.item()) if torch.is_tensor(s) else int(s)

    def _flat_ok(self, eng, params):
        if eng is None or eng.flat_p is None or len(self.param_groups) != 1:
            return False
        if len(params) != len(eng.params):
            return False
        base_p, base_g = eng.flat_p.data_ptr(), eng.flat_g.data_ptr()
        for p, q, k in zip(params, eng.params, eng.param_names):
            if p is not q or p.grad is None:
                return False
            off = eng.offsets[k] * 4
            if p.data_ptr() != base_p + off or p.grad.data_ptr() != base_g + off or p.grad.dtype != torch.float32:
                return False
        return True

    def _ensure_flat_state(self, eng):
        names = self._names()
        if self._flat is None or self._flat[0].numel() != eng.total or self._flat[0].device != eng.flat_p.device:
            self._flat = tuple(torch.zeros(eng.total, dtype=torch.float32, device=eng.flat_p.device) for _ in names)
        for p, k in zip(eng.params, eng.param_names):
            st = self.state[p]
            off = eng.offsets[k]
            for name, buf in zip(names, self._flat):
                view = buf[off: off + p.numel()].view(p.shape)
                cur = st.get(name)
                if cur is None:
                    st[name] = view
                elif cur.data_ptr() != view.data_ptr():     # e.g. after load_state_dict
                    view.copy_(cur)
                    st[name] = view
            if "step" not in st and self.kind != "sgd":
                st["step"] = torch.tensor(0.0, dtype=torch.float32)

    # -- step as part of Engine.train_step ---------------------------------------------------------
    def fused_update_args(self, eng):
        """Everything Engine.train_step(update=...) needs to run this optimizer's step inside the backward pass
        (unetk_fused_update: one kernel per parameter group un-packs the gradients, updates the parameters and
        re-packs the new weights), or None if this optimizer does not own exactly the engine's flat parameter
        buffer.  Advances the step count like step() does: call it once per training step, then do NOT call step()."""
        import math
        import numpy as np
        if eng is None or eng is not self._engine or eng.flat_p is None or len(self.param_groups) != 1:
            return None
        group = self.param_groups[0]
        params = group["params"]
        key = (eng.flat_p.data_ptr(), len(params))
        if getattr(self, "_fused_ok_key", None) != key:
            ok = len(params) == len(eng.params)
            base = eng.flat_p.data_ptr()
            for p, q, k in zip(params, eng.params, eng.param_names) if ok else ():
                if p is not q or p.data_ptr() != base + eng.offsets[k] * 4 or not p.requires_grad:
                    ok = False
                    break
            self._fused_ok_key, self._fused_ok = key, ok
        if not self._fused_ok or group.get("maximize") or group.get("amsgrad") or group.get("nesterov") or group.get("dampening"):
            return None
        first = self._names()[0] not in self.state[eng.params[0]]
        self._ensure_flat_state(eng)
        f32 = lambda x: float(np.float32(x))
        lr, wd = f32(group["lr"]), f32(group["weight_decay"])
        if self.kind == "sgd":
            mom = f32(group["momentum"])
            step = 1 if first else 2
            values = [lr, mom, 0.0, 0.0, wd, f32(1.0 - math.pow(mom, step)), 1.0, float(first), 1.0]
            return _FusedUpdate(self._flat[0], None, 2, values)
        states = [self.state[p] for p in eng.params]
        old_steps = [st.get("step") for st in states]
        step = self._step_of(states[0]) + 1
        t = torch.tensor(float(step), dtype=torch.float32)
        for st in states:
            st["step"] = t

        def rollback():
            for st, old in zip(states, old_steps):
                st["step"] = old
        b1, b2 = (f32(b) for b in group["betas"])
        # the same double-precision bias corrections unetk_adam_step derives from its float arguments
        bc1 = f32(1.0 - math.pow(b1, step))
        bc2_sqrt = f32(math.sqrt(1.0 - math.pow(b2, step)))
        values = [lr, b1, b2, f32(group["eps"]), wd, bc1, bc2_sqrt, float(step == 1), 1.0]
        return _FusedUpdate(self._flat[0], self._flat[1], _MODE[self.kind], values, rollback)

    # -- step ------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        eng = self._engine
        group0 = self.param_groups[0]
        if self._flat_ok(eng, group0["params"]):
            first = self._names()[0] not in self.state[eng.params[0]]
            self._ensure_flat_state(eng)
            self._launch(group0, eng.flat_p, eng.flat_g, self._flat, [self.state[p] for p in eng.params], first)
        else:
            for group in self.param_groups:
                for p in group["params"]:
                    if p.grad is None:
                        continue
                    if not p.is_cuda:
                        raise RuntimeError("FusedOptimizer has no CPU path: parameters must live on a CUDA device")
                    st = self.state[p]
                    first = self._names()[0] not in st
                    bufs = []
                    for name in self._names():
                        if name not in st:
                            st[name] = torch.zeros_like(p, memory_format=torch.preserve_format)
                        bufs.append(st[name].view(-1))
                    if "step" not in st and self.kind != "sgd":
                        st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    g = p.grad.contiguous().view(-1)
                    self._launch(group, p.data.view(-1), g, tuple(bufs), [st], first)
        if eng is not None:
            eng.mark_weights_dirty()
        return loss

    def _launch(self, group, p, g, bufs, states, first):
        lr, wd = float(group["lr"]), float(group["weight_decay"])
        if self.kind == "sgd":
            ops.adam_step(p, g, bufs[0], None, lr, float(group["momentum"]), 0.0, 0.0, wd, 1 if first else 2, 1.0, 2)
            return
        step = self._step_of(states[0]) + 1
        t = torch.tensor(float(step), dtype=torch.float32)
        for st in states:
            st["step"] = t
        b1, b2 = group["betas"]
        ops.adam_step(p, g, bufs[0], bufs[1], lr, float(b1), float(b2), float(group["eps"]), wd, step, 1.0,
                      _MODE[self.kind])
