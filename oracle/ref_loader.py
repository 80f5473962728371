"""Load the REFERENCE's own hot-path functions when /root/reference is present -- TEST / BASELINE INFRASTRUCTURE ONLY.

CAVIAR-1.0.py imports mdtraj, matplotlib, pyemma and seaborn at module level; none is installed here and none is
touched by the functions on the path (compute_distances_parallel :110-122, build_residue_components :183-199,
_trim_ca_indices :889-894), so empty stand-in modules are registered before the file is executed.
The reference does not travel to the GPU box: there ``load()`` returns None and callers fall back to the port
(oracle/caviar_oracle.py), saying so (``kind: "port"``).
"""
import importlib.util
import os
import sys
import types

REF_DIR = os.environ.get("CAVIAR_REFERENCE", "/root/reference")


def load():
    """-> the reference module (CAVIAR-1.0.py executed unmodified), or None when it is not on this machine."""
    path = os.path.join(REF_DIR, "CAVIAR-1.0.py")
    if not os.path.exists(path):
        return None
    for name in ("mdtraj", "matplotlib", "matplotlib.pyplot", "pyemma", "seaborn"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules["matplotlib"], "use"):
        sys.modules["matplotlib"].use = lambda *a, **k: None
    spec = importlib.util.spec_from_file_location("caviar_reference_main", path)
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except Exception:
        return None
    return mod
