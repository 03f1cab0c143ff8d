"""Import alias: the package directory is named `classifying-vae-lstm_b200` (not a valid Python
identifier), so `import clvae_b200[.sub.module]` resolves to the SAME module objects."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_ALIAS, _REAL = "clvae_b200", "classifying-vae-lstm_b200"
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[__name__] = importlib.import_module(_REAL)
