"""fluid_sims_b200 — B200-native per-timestep update path for the explicit solvers of
seanwevans/fluid-sims (2-D/3-D hypersonic, Gray-Scott, SPH).

The package is a thin host-side mirror of the reference's per-solver interfaces over the C-ABI in
include/tau_b200.h (libtau_b200.so, hand-written sm_100a CUDA).  No CPU fallback exists.
"""
from ._lib import LIB_PATH, TauError, device_count  # noqa: F401
