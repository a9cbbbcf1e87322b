"""Run an UNMODIFIED reference script on top of this package.

    python -m tiaozhanbei_unet_b200.dropin.launch /path/to/reference/src/train.py --category bottle ...
    python -m tiaozhanbei_unet_b200.dropin.launch /path/to/reference/train.py --data_root datasets/Gear ...

What it does: (a) registers empty stand-ins for matplotlib / seaborn when they are
not installed (src/utils.py:6-8 imports them at top level; they only draw plots), (b) registers
this package's `model` and `train_utils` under the flat names the reference imports, (c) runs the
script with runpy as __main__ from its own directory, so its `utils` and `dataset` stay its own.

The repository-root scripts of the segmentation pipelines (train.py, test.py, train_kolektorsdd.py ...)
import `src.model` / `src.metrics` instead (train.py:20-23): for them `src.model` is this package's
model module and `src.metrics` is the reference's own file (reports, plots) with its three loss entry points
(src/metrics.py:233-335) replaced by this package's kernels and `SegmentationMetrics` re-based on the
device-side confusion counts (its `compute_*` / `print_metrics` / `plot_confusion_matrix` stay the reference's); `src.utils` and the
dataset modules stay the reference's.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types


class _Anything:
    """Absorbs any attribute access / call (plotting stubs)."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter((_Anything(), _Anything()))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _stub(name):
    mod = types.ModuleType(name)

    def absorb(attr):
        if attr.startswith("__"):          # no fake __file__ / __path__ / __spec__: tools that scan sys.modules read them
            raise AttributeError(attr)
        return _Anything()
    mod.__getattr__ = absorb
    sys.modules[name] = mod
    return mod


def install(stub_plotting: bool = True):
    if stub_plotting:
        for name in ("matplotlib", "seaborn"):
            try:
                importlib.import_module(name)
            except ImportError:
                m = _stub(name)
                if name == "matplotlib":
                    for sub in ("pyplot", "patches", "widgets", "colors", "cm"):
                        setattr(m, sub, _stub(f"matplotlib.{sub}"))
    from tiaozhanbei_unet_b200.dropin import model as _model, train_utils as _train_utils
    sys.modules["model"] = _model
    sys.modules["train_utils"] = _train_utils


def install_src_package(root: str):
    """`src.model` / `src.metrics` for the scripts that sit next to the reference's src/ directory."""
    import importlib.util
    from tiaozhanbei_unet_b200 import metrics as _metrics, model as _model
    sys.modules["src.model"] = _model
    spec = importlib.util.spec_from_file_location("src.metrics", os.path.join(root, "src", "metrics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)       # the reference's SegmentationMetrics etc., unmodified
    for name in ("dice_loss", "focal_loss", "CombinedSegmentationLoss"):
        setattr(mod, name, getattr(_metrics, name))
    # SegmentationMetrics: the reference's own class (compute_* / print / plot unmodified) over the device-side
    # counts: `update` runs in every training step (train.py:146)
    reference_cls = mod.SegmentationMetrics
    mod.SegmentationMetrics = type("SegmentationMetrics", (_metrics.DeviceConfusion, reference_cls),
                                   {"__doc__": reference_cls.__doc__, "__module__": "src.metrics"})
    sys.modules["src.metrics"] = mod


def run(script: str, argv):
    script = os.path.abspath(script)
    install()
    here = os.path.dirname(script)
    if os.path.isfile(os.path.join(here, "src", "metrics.py")) and os.path.isfile(os.path.join(here, "src", "model.py")):
        install_src_package(here)
    sys.path.insert(0, here)
    sys.argv = [script] + list(argv)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m tiaozhanbei_unet_b200.dropin.launch <reference script> [args...]")
    run(sys.argv[1], sys.argv[2:])
