"""Import the UNMODIFIED reference modules in-process (dev container only; TEST INFRASTRUCTURE).

The reference needs two pip packages that are not installed (``fcn``, ``gdown``); neither is used on
the hot path (models.py:1-2,206; utils.py:1), so empty stand-ins are put on ``sys.modules`` first.
Returns ``(models, utils)`` or raises ``FileNotFoundError`` when no reference checkout is present
(the GPU box has none: nothing that runs there may call this).
"""
import importlib
import os
import sys
import types

# the dev container's checkout, then a driver-installed copy inside the repo (git-ignored).  bench.py passes the
# second one explicitly: nothing that runs on the GPU box may read /root/reference.
REF_DIRS = ("/root/reference", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref"))


def reference_root():
    for d in REF_DIRS:
        if os.path.isfile(os.path.join(d, "models.py")):
            return d
    return None


def load_reference(root=None):
    root = root or reference_root()
    if root is None or not os.path.isfile(os.path.join(root, "models.py")):
        raise FileNotFoundError("reference checkout not present")
    if "fcn" not in sys.modules:
        fcn = types.ModuleType("fcn")
        fcn.data = types.SimpleNamespace(cached_download=_no_network)
        fcn.utils = types.SimpleNamespace()
        sys.modules["fcn"] = fcn
    if "gdown" not in sys.modules:
        sys.modules["gdown"] = types.ModuleType("gdown")
    if root not in sys.path:
        sys.path.append(root)
    # the reference's module names are generic ("models", "utils"): import them under private names
    mods = []
    for name in ("models", "utils"):
        spec = importlib.util.spec_from_file_location("szn_reference_" + name, os.path.join(root, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods.append(m)
    return tuple(mods)


def _no_network(**kw):
    raise RuntimeError("no network in this environment")
