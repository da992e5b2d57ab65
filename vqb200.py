"""Import shim: `import vqb200` loads the package directory
`bridging-the-gap-of-robot-learning-via-distribution-reinforcement-learning-vq-vae_b200/`
(whose name is not a valid Python identifier) under the module name `vqb200`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "bridging-the-gap-of-robot-learning-via-distribution-reinforcement-learning-vq-vae_b200")


def _load():
    spec = importlib.util.spec_from_file_location(
        "vqb200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vqb200"] = mod          # replaces this shim so that `vqb200.x` resolves into the package
    spec.loader.exec_module(mod)
    return mod


_load()
