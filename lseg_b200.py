"""Import alias: ``import lseg_b200`` -> the package in ./lang-seg_b200/ (hyphenated directory).

``lseg_b200.X`` and ``lang-seg_b200.X`` are the SAME module objects: a meta-path finder maps every ``lseg_b200.*`` import
onto the real package, so module-level state (the loaded library handle, the tokenizer's stand-in switch) exists once."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_REAL = "lang-seg_b200"
_ALIAS = "lseg_b200"
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(_ALIAS + "."):
            return None
        real = _REAL + fullname[len(_ALIAS):]
        try:
            importlib.import_module(real)
        except ModuleNotFoundError as e:
            if e.name == real:
                return None
            raise
        return importlib.util.spec_from_loader(fullname, self, origin=real)

    def create_module(self, spec):
        return sys.modules[spec.origin]  # hand back the real module object: no second copy is executed

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
sys.modules[__name__] = _pkg
