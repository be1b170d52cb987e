"""recursive_control_variates_b200 — B200-native hot path of "Recursive Control Variates
for Inverse Rendering": correlated two-state volumetric render, recursive CV combination,
path-replay adjoint.  CUDA kernels + C-ABI in csrc/ (librcv_b200.so); the Python modules
mirror the reference's plugin / optimisation-loop interface on top of it."""
__version__ = "0.1.0"
