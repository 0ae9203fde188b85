"""TEST INFRASTRUCTURE -- import the UNMODIFIED reference modules in the build container.

`load_reference()` is used by `oracle/make_golden.py` and the `not gpu` pinning
tests only; it needs `/root/reference`, which does not exist on the GPU box.
`make_args()` (a plain Namespace with the reference's default flags, no reference
file involved) is also used by `bench.py`, `smoke()` and the GPU tests to
construct the modules.  Nothing in the product package imports this module.

Shims, none of which touches a reference file:
  1. Py2 integer division: `model.py:45-47,52-54,91-93` compute `hidden_size/2`;
     `args.hidden_size` is passed as an int subclass whose `/` floors.
  2. offline weights: `model.py:30-31` calls `models.resnet101(pretrained=True)`;
     `torchvision.models.resnet101` is patched to build with `weights=None`.
  3. import plumbing: Py2 implicit-relative imports -> sys.path order
     [src, src/modules, src/utils]; `munkres` stubbed so `src/test.py` imports.
"""
from __future__ import annotations

import os
import sys
import types
from argparse import Namespace

REFERENCE_ROOT = os.environ.get("RSIS_REFERENCE_ROOT", "/root/reference")


class FloorInt(int):
    """int whose true division floors (Python-2 semantics for `hidden_size/2`)."""

    def __truediv__(self, other):
        return FloorInt(int(self) // int(other))

    def __floordiv__(self, other):
        return FloorInt(int(self) // int(other))


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "modules", "model.py"))


_cached = None


def load_reference():
    """Returns a namespace with FeatureExtractor, RSIS, ConvLSTMCell, test (the reference's own objects)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # never drop __pycache__ into the read-only mount
    src = os.path.join(REFERENCE_ROOT, "src")
    for p in (os.path.join(src, "utils"), os.path.join(src, "modules"), src):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if "munkres" not in sys.modules:
        stub = types.ModuleType("munkres")

        class Munkres:  # pragma: no cover - never called on the inference path
            def compute(self, cost):
                from scipy.optimize import linear_sum_assignment
                import numpy as np
                r, c = linear_sum_assignment(np.asarray(cost))
                return list(zip(r.tolist(), c.tolist()))

        stub.Munkres = Munkres
        sys.modules["munkres"] = stub
    import torchvision.models as tvm

    if not getattr(tvm.resnet101, "_rsis_offline", False):
        _orig = tvm.resnet101

        def resnet101_offline(pretrained=False, **kw):
            kw.pop("weights", None)
            return _orig(weights=None, **kw)

        resnet101_offline._rsis_offline = True
        tvm.resnet101 = resnet101_offline
    import modules.model as ref_model  # noqa: E402  (reference file, unmodified)
    import modules.clstm as ref_clstm  # noqa: E402
    import test as ref_test  # noqa: E402  (src/test.py: the inference loop)

    _cached = Namespace(FeatureExtractor=ref_model.FeatureExtractor, RSIS=ref_model.RSIS,
                        ConvLSTMCell=ref_clstm.ConvLSTMCell, test=ref_test.test)
    return _cached


def make_args(hidden_size=128, kernel_size=3, num_classes=21, maxseqlen=10, **kw) -> Namespace:
    """The subset of `src/args.py` fields the hot path reads (args.py:30,103-123), CPU mode."""
    a = Namespace(base_model="resnet101", hidden_size=FloorInt(hidden_size), kernel_size=kernel_size,
                  num_classes=num_classes, skip_mode="concat", dropout=0.0, dropout_stop=0.0,
                  dropout_cls=0.0, use_gpu=False, maxseqlen=maxseqlen, imsize=256, ngpus=1)
    for k, v in kw.items():
        setattr(a, k, v)
    return a
