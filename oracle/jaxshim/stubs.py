"""Import hook that stubs the GUI / rendering / robot dependencies of the reference's env modules (cv2, gym,
pyrender, trimesh, open3d, pxr, the reference's own pyrender/usdrender packages, brax, flax ...) so that
`daxbench.core.envs.*` can be imported under the jax shim for golden generation.  TEST INFRASTRUCTURE ONLY."""
import importlib.abc
import importlib.machinery
import sys
import types
from unittest.mock import MagicMock

PREFIXES = ("cv2", "gym", "pyrender", "trimesh", "open3d", "pxr", "imageio", "matplotlib", "tensorboardX", "wandb",
            "absl", "brax", "flax", "daxbench.core.engine.pyrender", "daxbench.core.engine.usdrender",
            "daxbench.core.envs.others")


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if any(fullname == p or fullname.startswith(p + ".") for p in PREFIXES):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
