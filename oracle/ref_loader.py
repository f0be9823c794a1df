"""TEST INFRASTRUCTURE.  Imports the reference's own source files (read-only, /root/reference) with
`oracle/jaxshim` supplying the `jax` / `optax` / `jaxopt` / `pynapple` import names (JAX cannot be installed
in this image).  The package `__init__` of the reference (plotting / analysis imports) is bypassed: a bare
namespace module named `poor_man_gplvm` points at the reference directory, so `poor_man_gplvm.core`,
`.decoder`, `.fit_tuning_helper`, `.gp_kernel` load from the unmodified files.
Only `tests/golden/make_golden.py` (run in the build container) uses this; nothing on the GPU box does."""
from __future__ import annotations

import importlib
import os
import sys
import types

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jaxshim")


def available(ref_root="/root/reference"):
    return os.path.isdir(os.path.join(ref_root, "poor_man_gplvm"))


def load_reference(ref_root="/root/reference"):
    """-> namespace with .core, .decoder, .fth, .gpk of the reference."""
    try:
        import jax  # noqa: F401
        if getattr(jax, "__version__", "").endswith("shim") is False:
            raise RuntimeError("a real jax is importable: run the reference on it instead of the shim")
    except ImportError:
        pass
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    pkg = types.ModuleType("poor_man_gplvm")
    pkg.__path__ = [os.path.join(ref_root, "poor_man_gplvm")]
    sys.modules["poor_man_gplvm"] = pkg
    ns = types.SimpleNamespace()
    ns.gpk = importlib.import_module("poor_man_gplvm.gp_kernel")
    ns.fth = importlib.import_module("poor_man_gplvm.fit_tuning_helper")
    ns.decoder = importlib.import_module("poor_man_gplvm.decoder")
    ns.core = importlib.import_module("poor_man_gplvm.core")
    return ns
