"""Golden vectors for input sizes that are not multiples of 16 (the F.pad path, src/model.py:57-61) from the
UNMODIFIED reference on CPU, fp32:  python tests/golden/make_golden_odd.py  ->  tests/golden/odd_golden.pt"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("_ref_model", "/root/reference/src/model.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {"torch": torch.__version__, "cases": []}
    for arch, shape, wseed in (("unet", (1, 3, 100, 100), 31), ("anomaly_unet", (2, 3, 52, 76), 32), ("unet", (1, 3, 37, 90), 33)):
        m = ref.UNet(3, 1) if arch == "unet" else ref.AnomalyUNet(3)
        m.load_state_dict(O.synthetic_state_dict(arch, seed=wseed))
        m.eval()
        g = torch.Generator().manual_seed(wseed + 100)
        x = torch.randn(*shape, generator=g)
        with torch.no_grad():
            y = m(x)
        outs = [y] if arch == "unet" else list(y)
        assert all(o.shape[2:] == x.shape[2:] for o in outs)
        out["cases"].append({"arch": arch, "shape": shape, "weights_seed": wseed, "input_seed": wseed + 100,
                             "outputs": [o.clone() for o in outs]})
    path = os.path.join(ROOT, "tests", "golden", "odd_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
