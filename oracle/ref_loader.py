"""Import the UNMODIFIED reference Aggregator from /root/reference.  TEST INFRASTRUCTURE ONLY.

The reference file ``cat_seg/modeling/transformer/model.py`` imports ``timm.layers`` (model.py:14),
which is not installed; it only uses ``Mlp`` (:159), ``DropPath`` (:157, never hit with the default
drop_path=0) and ``to_2tuple`` (:69,154).  A minimal stand-in is registered in ``sys.modules``
before loading the file by path.  ``timm.layers.Mlp`` (timm 0.8.3.dev0,
requirements.txt:7) is Linear -> act -> Dropout -> Linear -> Dropout with attribute names
fc1/act/drop1/fc2/drop2.

/root/reference does not exist on the GPU box: ``build_ref()`` (run by ``__graft_entry__.build()`` in the build
container) leaves an unmodified copy of model.py in the git-ignored ``oracle/_ref/``, which travels with the
snapshot; callers must still check ``reference_available()``.
"""
from __future__ import annotations

import collections.abc
import importlib.util
import os
import sys
import types
from itertools import repeat

import torch.nn as nn

_SRC = "/root/reference/cat_seg/modeling/transformer/model.py"
_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# oracle/_ref/model.py: an UNMODIFIED copy made at build time by build_ref() (git-ignored, travels to the GPU box with
# gpurun like the built .so): bench.py --impl reference times the reference's own file there.
_COPY = os.path.join(_REF_DIR, "model.py")


def build_ref() -> bool:
    """Build step (``__graft_entry__.build``): copy the reference's model.py, unmodified, into oracle/_ref/ when
    /root/reference is present.  Returns True if oracle/_ref/model.py exists afterwards."""
    import shutil
    if os.path.isfile(_SRC):
        os.makedirs(_REF_DIR, exist_ok=True)
        shutil.copyfile(_SRC, _COPY)
    return os.path.isfile(_COPY)


def _model_path() -> str:
    env = os.environ.get("CATSEG_REFERENCE_MODEL")
    if env:
        return env
    return _SRC if os.path.isfile(_SRC) else _COPY


REF_MODEL_PATH = _model_path()


def reference_available() -> bool:
    return os.path.isfile(_model_path())


def _ntuple(n):
    def parse(x):
        if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
            return tuple(x)
        return tuple(repeat(x, n))
    return parse


class _Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        hidden_features = hidden_features or in_features
        out_features = out_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


_CACHED = None


def load_reference_module():
    """Returns the reference ``model`` module (with Aggregator etc.)."""
    global _CACHED
    if _CACHED is not None:
        return _CACHED
    path = _model_path()
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    if "timm.layers" not in sys.modules:
        layers = types.ModuleType("timm.layers")
        layers.Mlp, layers.DropPath, layers.PatchEmbed = _Mlp, nn.Identity, object
        layers.to_2tuple, layers.to_ntuple = _ntuple(2), _ntuple
        layers.trunc_normal_, layers._assert = nn.init.trunc_normal_, (lambda c, m: None)
        timm = types.ModuleType("timm")
        timm.layers = layers
        sys.modules["timm"], sys.modules["timm.layers"] = timm, layers
    spec = importlib.util.spec_from_file_location("catseg_reference_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _CACHED = mod
    return mod


def build_reference_aggregator(kwargs: dict, state_dict=None):
    """kwargs = the Aggregator ctor kwargs (cat_seg_predictor.py:97-113)."""
    mod = load_reference_module()
    agg = mod.Aggregator(**kwargs).eval()
    if state_dict is not None:
        missing, unexpected = agg.load_state_dict(state_dict, strict=False)
        # block_2.attn_mask is a buffer the reference builds itself (model.py:161-183)
        missing = [m for m in missing if not m.endswith("attn_mask")]
        if missing or unexpected:
            raise RuntimeError(f"state_dict mismatch: missing={missing} unexpected={unexpected}")
    return agg
