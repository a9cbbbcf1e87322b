"""Run an UNMODIFIED reference script (src/train.py, src/test.py) on top of this package.

    python -m tiaozhanbei_unet_b200.dropin.launch /path/to/reference/src/train.py --category bottle ...

What it does: (a) registers empty stand-ins for matplotlib / seaborn when they are
not installed (src/utils.py:6-8 imports them at top level; they only draw plots), (b) registers
this package's `model` and `train_utils` under the flat names the reference imports, (c) runs the
script with runpy as __main__ from its own directory, so its `utils` and `dataset` stay its own.
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
    mod.__getattr__ = lambda attr: _Anything()
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


def run(script: str, argv):
    script = os.path.abspath(script)
    install()
    sys.path.insert(0, os.path.dirname(script))
    sys.argv = [script] + list(argv)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m tiaozhanbei_unet_b200.dropin.launch <reference script> [args...]")
    run(sys.argv[1], sys.argv[2:])
