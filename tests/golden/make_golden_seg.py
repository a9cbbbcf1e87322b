"""Golden vectors for the SegmentationUNet row from the UNMODIFIED reference
(/root/reference/src/model.py SegmentationUNet, src/metrics.py losses) on CPU, fp32.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_seg.py
Writes tests/golden/seg_golden.pt: seeded-input recipes plus the reference's loss values, logit
gradients, eval-mode logits and argmax maps.
"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402

for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def load(name):
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", f"/root/reference/src/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def loss_inputs(seed, n=2, c=4, h=24, w=20, ignore=None):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, c, h, w, generator=g) * 2.0
    target = torch.randint(0, c, (n, h, w), generator=g)
    if ignore is not None:
        target[torch.rand(n, h, w, generator=g) < 0.1] = ignore
    return logits, target


def main():
    ref_metrics, ref_model = load("metrics"), load("model")
    out = {"torch": torch.__version__, "losses": [], "dice": None, "focal": None, "model": None}
    cases = [dict(ce_weight=1.0, dice_weight=1.0, focal_weight=0.0, ignore_index=None, class_weights=None),
             dict(ce_weight=1.0, dice_weight=0.5, focal_weight=0.0, ignore_index=None, class_weights=[1.0, 2.0, 3.0, 0.5]),
             dict(ce_weight=0.7, dice_weight=0.0, focal_weight=1.5, ignore_index=255, class_weights=None),
             dict(ce_weight=0.0, dice_weight=1.0, focal_weight=0.0, ignore_index=None, class_weights=None),
             dict(ce_weight=1.0, dice_weight=1.0, focal_weight=0.5, ignore_index=255, class_weights=[2.0, 1.0, 1.0, 1.0])]
    for i, kw in enumerate(cases):
        # dice's one_hot rejects labels outside [0, C): ignored pixels only where dice is off
        use_ignore = kw["ignore_index"] if kw["dice_weight"] == 0 else None
        logits, target = loss_inputs(100 + i, ignore=use_ignore)
        logits.requires_grad_(True)
        val = ref_metrics.CombinedSegmentationLoss(**kw)(logits, target)
        val.backward()
        out["losses"].append({"seed": 100 + i, "ignore_in_target": use_ignore, "kwargs": kw, "value": float(val),
                              "grad": logits.grad.clone()})
    logits, target = loss_inputs(200)
    probs = torch.softmax(logits, dim=1).requires_grad_(True)
    d = ref_metrics.dice_loss(probs, target)
    d.backward()
    out["dice"] = {"seed": 200, "value": float(d), "grad": probs.grad.clone()}
    logits, target = loss_inputs(201, ignore=255)
    logits.requires_grad_(True)
    f = ref_metrics.focal_loss(logits, target, alpha=0.75, gamma=2, ignore_index=255)
    f.backward()
    out["focal"] = {"seed": 201, "alpha": 0.75, "gamma": 2, "ignore_index": 255, "value": float(f), "grad": logits.grad.clone()}
    # SegmentationUNet: eval-mode logits and argmax map, synthetic weights by seeded recipe
    m = ref_model.SegmentationUNet(3, 4)
    sd = O.synthetic_state_dict("unet", seed=11, n_classes=4)
    m.load_state_dict(sd)
    m.eval()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 3, 32, 48, generator=g)
    with torch.no_grad():
        z = m(x)
    n_params = sum(p.numel() for p in m.parameters())
    out["model"] = {"weights_seed": 11, "input_seed": 7, "shape": (2, 3, 32, 48), "logits": z.clone(),
                    "argmax": torch.argmax(z, dim=1).clone(), "n_params": n_params,
                    "keys": list(m.state_dict().keys())}
    path = os.path.join(ROOT, "tests", "golden", "seg_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; params", n_params)


if __name__ == "__main__":
    main()
