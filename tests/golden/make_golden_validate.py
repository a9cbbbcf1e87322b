"""Golden vectors for validate_epoch's reductions from the UNMODIFIED reference
(/root/reference/src/train_utils.py validate_epoch, CombinedLoss; src/utils.py) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_validate.py
Writes tests/golden/validate_golden.pt.  A stand-in model replays seeded (reconstruction, anomaly_map) batches,
so the recorded numbers depend on the reference's validation code alone: the three loss averages, image_metrics,
pixel_metrics, the shapes of `predictions`, and the exception the mixed-label case ends in.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def batches(seed, labels, batch=3, h=24, w=20):
    """Seeded validation set: images, sparse masks (none on normal images), replayed model outputs."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(0, len(labels), batch):
        lab = torch.tensor(labels[i:i + batch])
        n = len(lab)
        image = torch.rand(n, 3, h, w, generator=g)
        mask = (torch.rand(n, 1, h, w, generator=g) > 0.9).float() * lab.view(n, 1, 1, 1).float()
        recon = (image + 0.1 * torch.randn(n, 3, h, w, generator=g)).clamp(0, 1)
        amap = (0.35 * mask + 0.65 * torch.rand(n, 1, h, w, generator=g)).clamp(1e-4, 1 - 1e-4)
        out.append({"image": image, "mask": mask, "label": lab, "recon": recon, "amap": amap})
    return out


class Replay(torch.nn.Module):
    def __init__(self, data):
        super().__init__()
        self.data, self.i = data, 0

    def forward(self, images):
        b = self.data[self.i]
        self.i += 1
        return b["recon"], b["amap"]


CASES = {"anomalous": [1, 1, 1, 1, 1, 1, 1], "normal": [0, 0, 0, 0, 0], "mixed": [0, 1, 0, 1, 1, 0]}


def main():
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, "/root/reference/src")      # train_utils does `from utils import ...`
    import train_utils as ref
    out = {"torch": torch.__version__, "numpy": np.__version__, "cases": {}}
    for ci, (name, labels) in enumerate(CASES.items()):
        data = batches(300 + ci, labels)
        rec = {"seed": 300 + ci, "labels": labels}
        try:
            res = ref.validate_epoch(Replay(data), data, ref.CombinedLoss(), torch.device("cpu"))
            rec.update(total_loss=res["total_loss"], recon_loss=res["recon_loss"], seg_loss=res["seg_loss"],
                       image_metrics={k: float(v) for k, v in res["image_metrics"].items()},
                       pixel_metrics={k: float(v) for k, v in res["pixel_metrics"].items()},
                       shapes={k: tuple(v.shape) for k, v in res["predictions"].items()},
                       dtypes={k: str(v.dtype) for k, v in res["predictions"].items()},
                       percentile95=float(np.percentile(res["predictions"]["scores"], 95)),
                       scores_sum=float(res["predictions"]["scores"].astype(np.float64).sum()))
        except Exception as e:            # the mixed-label case: sklearn rejects [N] against [N*H*W]
            rec.update(error_type=type(e).__name__, error=str(e))
            scores = np.array([s for b in data for s in ((b["recon"] - b["image"]) ** 2).mean(1).numpy()])
            rec["percentile95"] = float(np.percentile(scores, 95))
        out["cases"][name] = rec
        print(name, {k: v for k, v in rec.items() if k not in ("labels",)})
    torch.save(out, os.path.join(HERE, "validate_golden.pt"))


if __name__ == "__main__":
    main()
