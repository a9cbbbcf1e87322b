"""Golden vectors for the bilinear=True decoders (src/model.py:47-49) from the UNMODIFIED reference on CPU,
fp32:  python tests/golden/make_golden_bilinear.py  ->  tests/golden/bilinear_golden.pt"""
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
    for arch, shape, wseed in (("unet", (2, 3, 48, 64), 51), ("anomaly_unet", (1, 3, 32, 32), 52), ("unet", (1, 3, 50, 36), 53)):
        m = ref.UNet(3, 1, bilinear=True) if arch == "unet" else ref.AnomalyUNet(3, bilinear=True)
        sd = O.synthetic_state_dict(arch, seed=wseed, bilinear=True)
        assert list(sd.keys()) == list(m.state_dict().keys())
        m.load_state_dict(sd)
        m.eval()
        g = torch.Generator().manual_seed(wseed + 100)
        x = torch.randn(*shape, generator=g)
        with torch.no_grad():
            y = m(x)
        outs = [y] if arch == "unet" else list(y)
        out["cases"].append({"arch": arch, "shape": shape, "weights_seed": wseed, "input_seed": wseed + 100,
                             "outputs": [o.clone() for o in outs], "n_params": sum(p.numel() for p in m.parameters()),
                             "keys": list(m.state_dict().keys())})
    path = os.path.join(ROOT, "tests", "golden", "bilinear_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes;", [c["n_params"] for c in out["cases"]])


if __name__ == "__main__":
    main()
