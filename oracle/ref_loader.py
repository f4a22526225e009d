"""Import the UNMODIFIED reference classes from /root/reference (dev container only — the GPU box has no
/root/reference, so nothing that runs there may call this).  Used by tests/golden/make_golden.py to generate the
committed fixtures and by tests that cross-check the oracle port when the reference tree is present.

The reference module imports gradio / diffusers / spaces / cog at module scope
(StoryDiffusion/Comic_Generation.py:3,21-28; StoryDiffusion/utils/pipeline.py:13-19); none is installed and none
is used by the hot path, so they are replaced by MagicMock modules through a meta-path finder.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
from unittest import mock

REFERENCE_ROOT = os.environ.get("CSA_REFERENCE_ROOT", "/root/reference")
_STUBBED = ("gradio", "diffusers", "spaces", "cog")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__ = []
        m.__spec__ = spec
        m.__name__ = spec.name
        return m

    def exec_module(self, module):
        return None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "StoryDiffusion", "Comic_Generation.py"))


_cached = None


def load_reference():
    """Returns the reference module ``StoryDiffusion.Comic_Generation`` (its globals are the control surface)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _StubFinder())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _cached = importlib.import_module("StoryDiffusion.Comic_Generation")
    return _cached
