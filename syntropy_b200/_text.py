"""Loader of the thin torch C++ extension (``csrc_torch/ksg_torch.cpp`` -> ``lib/ksg_torch_ext.so``).

The extension is the binding BASELINE.json names for the hot path: tensors in, tensors out, the
current torch stream, every kernel reached through the C ABI of ``include/ksg_b200.h``.  It is built
in-tree next to ``libksg_b200.so`` (it links it through an ``$ORIGIN`` rpath).  ``KSG_B200_BINDING=ctypes``
keeps the headline step on the ctypes view instead (A/B of the host time per step)."""
from __future__ import annotations

import importlib.util
import os

from . import _build, _cabi

_ext = None


def enabled() -> bool:
    return os.environ.get("KSG_B200_BINDING", "torch") != "ctypes"


def load():
    """Import (building first if the source is newer) the extension module; raises if it cannot."""
    global _ext
    if _ext is not None:
        return _ext
    _cabi.load()  # libksg_b200.so first: the extension's DT_NEEDED resolves to the already-loaded library
    if _build.torch_ext_is_stale():
        try:
            _build.build_torch_ext()
        except Exception as exc:
            if not os.path.exists(_build.TORCH_EXT):
                raise ImportError("ksg_torch_ext.so is missing and could not be built (%s)" % exc) from exc
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    spec = importlib.util.spec_from_file_location("ksg_torch_ext", _build.TORCH_EXT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.abi_version() != _cabi.load().ksg_version():
        raise ImportError("ksg_torch_ext.so and libksg_b200.so disagree about the ABI version")
    _ext = mod
    return mod
