"""Generate golden vectors from the UNMODIFIED reference (/root/reference/src) on CPU, fp32.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/unet_golden.pt (a few hundred KB): seeded-input recipes plus the reference's
outputs, losses, gradient summaries, BN buffers and one Adam step, for UNet and AnomalyUNet.
Weights come from oracle.unet_oracle.synthetic_state_dict(seed) — a seeded recipe, so the 170 MB
state_dict itself is not stored.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"


def import_reference():
    """src/utils.py imports matplotlib/seaborn at top level; they are absent here -> empty stubs."""
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import importlib.util
    mods = {}
    for name in ("model", "utils", "train_utils"):
        spec = importlib.util.spec_from_file_location(f"_ref_{name}", os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        # train_utils does `from utils import ...` (flat import): expose the reference's utils under that name
        if name == "train_utils":
            saved = sys.modules.get("utils")
            sys.modules["utils"] = mods["utils"]
            try:
                spec.loader.exec_module(mod)
            finally:
                if saved is None:
                    del sys.modules["utils"]
                else:
                    sys.modules["utils"] = saved
        else:
            spec.loader.exec_module(mod)
        mods[name] = mod
    return mods


def summarize(t, k=6):
    t = t.detach().float().flatten()
    idx = torch.linspace(0, t.numel() - 1, k).long()
    return {"norm": float(t.norm()), "sum": float(t.sum()), "samples": t[idx].clone(), "numel": t.numel()}


def make_inputs(seed, n, h, w, c=3):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(n, c, h, w, generator=g)
    mask = (torch.rand(n, 1, h, w, generator=g) > 0.95).float()
    return image, mask


def main():
    from oracle import unet_oracle as O
    ref = import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out = {"torch_version": torch.__version__, "cases": {}}

    for arch, cls_name in (("anomaly_unet", "AnomalyUNet"), ("unet", "UNet")):
        for (n, h, w) in ((2, 32, 32), (1, 48, 64)):
            case = {"arch": arch, "weight_seed": 7, "input_seed": 11, "shape": (n, h, w)}
            sd = O.synthetic_state_dict(arch, seed=7)
            model = getattr(ref["model"], cls_name)(3) if arch == "anomaly_unet" else ref["model"].UNet(3, 1)
            model.load_state_dict(sd, strict=True)
            image, mask = make_inputs(11, n, h, w)
            # ---- eval forward
            model.eval()
            with torch.no_grad():
                o = model(image)
            case["eval_out"] = [t.clone() for t in (o if isinstance(o, tuple) else (o,))]
            # ---- train forward + loss + backward
            model.train()
            crit = ref["train_utils"].CombinedLoss()
            opt = ref["train_utils"].get_optimizer(model)          # Adam lr 1e-3 wd 1e-4
            o = model(image)
            if arch == "anomaly_unet":
                ld = crit(o[0], o[1], image, mask)
                total = ld["total_loss"]
                case["losses"] = {k: float(v) for k, v in ld.items()}
                case["train_out"] = [t.detach().clone() for t in o]
            else:
                p = torch.sigmoid(o)
                rl = torch.nn.functional.mse_loss(p, mask)
                sl = crit.focal_loss(p, mask)
                total = rl + sl
                case["losses"] = {"total_loss": float(total), "recon_loss": float(rl), "seg_loss": float(sl)}
                case["train_out"] = [o.detach().clone()]
            opt.zero_grad()
            total.backward()
            case["grads"] = {k: summarize(p.grad) for k, p in model.named_parameters()}
            case["buffers_after"] = {k: summarize(v) for k, v in model.state_dict().items() if O.is_buffer(k)}
            opt.step()
            case["params_after_adam"] = {k: summarize(p) for k, p in model.named_parameters()}
            out["cases"][f"{arch}_{n}x{h}x{w}"] = case
            print("done", arch, n, h, w, case["losses"])

    # ---- losses and scoring on their own
    g = torch.Generator().manual_seed(5)
    a = torch.rand(2, 3, 40, 56, generator=g)
    b = (a + 0.1 * torch.randn(2, 3, 40, 56, generator=g)).clamp(0, 1)
    a.requires_grad_(True)
    ssim = ref["train_utils"].SSIMLoss()
    val = ssim(a, b)
    val.backward()
    out["ssim"] = {"seed": 5, "shape": (2, 3, 40, 56), "value": float(val), "grad": a.grad.clone(),
                   "per_image": ref["train_utils"].SSIMLoss(size_average=False)(a.detach(), b).clone(),
                   "self": float(ssim(b, b))}
    g = torch.Generator().manual_seed(6)
    recon = torch.rand(2, 3, 24, 24, generator=g).requires_grad_(True)
    amap = torch.rand(2, 1, 24, 24, generator=g).clamp(1e-4, 1 - 1e-4).requires_grad_(True)
    img = torch.randn(2, 3, 24, 24, generator=g)
    msk = (torch.rand(2, 1, 24, 24, generator=g) > 0.9).float()
    crit = ref["train_utils"].CombinedLoss(recon_weight=0.7, seg_weight=1.3)
    ld = crit(recon, amap, img, msk)
    ld["total_loss"].backward()
    out["combined_loss"] = {"seed": 6, "weights": (0.7, 1.3), "values": {k: float(v) for k, v in ld.items()},
                            "d_recon": recon.grad.clone(), "d_amap": amap.grad.clone()}
    out["anomaly_score"] = {"mse": ref["utils"].compute_anomaly_score(recon.detach(), img).clone(),
                            "l1": ref["utils"].compute_anomaly_score(recon.detach(), img, "l1").clone()}
    # ---- known-answer facts of SURVEY.md §4
    out["facts"] = {
        "unet_params": sum(p.numel() for p in ref["model"].UNet(3, 1).parameters()),
        "anomaly_params": sum(p.numel() for p in ref["model"].AnomalyUNet(3).parameters()),
        "unet_keys": list(ref["model"].UNet(3, 1).state_dict().keys()),
        "anomaly_keys": list(ref["model"].AnomalyUNet(3).state_dict().keys()),
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "unet_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
