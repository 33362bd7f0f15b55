"""Alias package: ``import sac_eo.<anything>`` resolves to ``sac_expert_b200.sac_eo.<anything>`` (the SAME module
objects), so that code written against the reference's package name - in particular the reference's unmodified
``train.py`` - runs on the CUDA path.  Put ``<repo>/shims`` and ``<repo>`` on PYTHONPATH."""
import importlib
import importlib.abc
import importlib.machinery
import sys

_REAL = "sac_expert_b200.sac_eo"


class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith("sac_eo."):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len("sac_eo"):])

    def exec_module(self, module):
        pass


sys.meta_path.insert(0, _Alias())
_real = importlib.import_module(_REAL)
__path__ = list(_real.__path__)
