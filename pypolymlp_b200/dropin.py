"""Makes the compiled drop-in (`pypolymlp_b200/lib/libmlpcpp.*.so`) importable under the reference's own module
name `pypolymlp.cxx.lib.libmlpcpp`, so that an UNMODIFIED checkout of the reference's Python package
(`src/pypolymlp`, whose `cxx/lib/` directory is a build product and absent from a source tree) runs on top of it.

The reference imports the extension as `from pypolymlp.cxx.lib import libmlpcpp`
(src/pypolymlp/mlp_dev/core/features.py:10, src/pypolymlp/cxx/wrapper/api_gtinv_list.py:3,
src/pypolymlp/calculator/properties_single.py).  INTEGRATION.md describes the permanent variant (copy the two
shared objects into `src/pypolymlp/cxx/lib/`); `install()` is the in-process variant used by the tests."""

import glob
import importlib.util
import os
import sys
import types

_LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")


def load_extension():
    """Imports the compiled pybind11 module itself (not the ctypes mirror `pypolymlp_b200.libmlpcpp`)."""
    hits = sorted(glob.glob(os.path.join(_LIB_DIR, "libmlpcpp.*.so")))
    if not hits:
        raise ImportError("pypolymlp_b200/lib/libmlpcpp.*.so is missing: run `python __graft_entry__.py` (build())")
    spec = importlib.util.spec_from_file_location("libmlpcpp", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def install(reference_src=None):
    """Registers the extension as `pypolymlp.cxx.lib.libmlpcpp`.  `reference_src` (the directory holding the
    reference's `pypolymlp/` package) is prepended to sys.path when given.  Returns the extension module."""
    if reference_src is not None and reference_src not in sys.path:
        sys.path.insert(0, reference_src)
    name = "pypolymlp.cxx.lib.libmlpcpp"
    if name in sys.modules:
        return sys.modules[name]
    ext = load_extension()
    import pypolymlp.cxx  # the reference's package (namespace or regular); fails loudly when it is not importable

    pkg = types.ModuleType("pypolymlp.cxx.lib")
    pkg.__path__ = [_LIB_DIR]
    pkg.libmlpcpp = ext
    sys.modules["pypolymlp.cxx.lib"] = pkg
    sys.modules[name] = ext
    pypolymlp.cxx.lib = pkg
    return ext
