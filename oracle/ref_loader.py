"""Import the reference's OWN importable modules from ``/root/reference`` --
build-container only (the GPU box has no ``/root/reference``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Used by ``oracle/gen_golden.py``
and by CPU tests that are skipped when the checkout is absent.  Nothing under
``-m gpu``, ``smoke()`` or ``bench.py`` calls this.

Importable as-is: ``path_planner/utils/reeds_shepp.py``
(needs a ``matplotlib.pyplot`` stub: imported at ``:3`` but unused),
``path_planner/utils/a_star_utils.py``, ``path_planner/utils/path_utils.py``.
Everything else needs shapely / heapdict / dubins, which are not installable.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HL_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "path_planner", "utils", "reeds_shepp.py"))


def _stub_matplotlib():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
            return
        except Exception:
            pass
        m = types.ModuleType("matplotlib")
        p = types.ModuleType("matplotlib.pyplot")
        m.pyplot = p
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = p


def load(name):
    """Load ``path_planner/utils/<name>.py`` under a private module name."""
    _stub_matplotlib()
    path = os.path.join(REFERENCE_ROOT, "path_planner", "utils", name + ".py")
    spec = importlib.util.spec_from_file_location("_hl_reference_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
