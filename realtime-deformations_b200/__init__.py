"""realtime-deformations_b200 — B200-native MPM substep behind the reference's LagrangeEulerView stage interface.

The directory name is not a Python identifier; import it with
    importlib.import_module("realtime-deformations_b200")
(or `import mpm_b200` from the repo root, which does exactly that).
"""
from . import build, capi, scenes  # noqa: F401
