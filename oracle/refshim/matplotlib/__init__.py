"""placeholder (test infrastructure): any attribute / submodule resolves to an inert object."""
import importlib.abc
import importlib.machinery
import sys
import types


class _Inert(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert(f"{self.__name__}.{name}")

    def __call__(self, *a, **k):
        return self


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.startswith("matplotlib."):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return _Inert(spec.name)

    def exec_module(self, module):
        module.__path__ = []


sys.meta_path.append(_Finder())
__path__ = []
