"""TEST INFRASTRUCTURE: import the reference's Warp kernel modules from /root/reference with `warp` replaced by the pure-Python
stand-in (oracle/warp_shim) and every other absent third-party package (trimesh, yourdfpy, ...) replaced by inert stubs.
Only usable in the authoring container (needs /root/reference); the golden fixtures it produces are what travels."""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFERENCE = "/root/reference"
STUB_PACKAGES = ("trimesh", "yourdfpy", "usd", "pxr", "nvtx", "cuda", "scipy_missing")


class _Inert:
    def __init__(self, *a, **k): pass
    def __call__(self, *a, **k): return _Inert()
    def __getattr__(self, n): return _Inert()
    def __mro_entries__(self, bases): return (object,)
    def __iter__(self): return iter(())
    def __bool__(self): return False


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = types.ModuleType(spec.name)
        m.__path__ = []
        m.__getattr__ = lambda n: _Inert()
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in STUB_PACKAGES:
            return importlib.machinery.ModuleSpec(name, _StubLoader(), is_package=True)
        return None


_ready = False


def prepare(root=None):
    """`root`: where the `curobo` package is imported from -- the reference tree (default; authoring container only) or
    oracle/_ref/pyref, the byte-code build of the reference's call sites that travels to the GPU box (oracle/build_pyref.py)."""
    global _ready
    if _ready:
        return
    root = REFERENCE if root is None else root
    if not os.path.isdir(root):
        raise RuntimeError(f"{root} is not present: golden fixtures can only be regenerated in the authoring container")
    import numpy  # noqa: F401
    import torch  # noqa: F401  (before the reference: its modules import torch lazily in odd orders)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "warp_shim"))
    sys.path.insert(0, root)
    sys.meta_path.append(_StubFinder())   # after the real finders: only packages that are really absent get stubbed
    _ready = True


def ref(module):
    prepare()
    return importlib.import_module(module)
