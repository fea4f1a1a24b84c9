"""vlapy_b200 -- B200-native drop-in for VlaPy's per-timestep phase-space update.

Select it with ``all_params["backend"]["core"] = "b200"``; the factory functions in
``vlapy_b200.core.*`` and ``vlapy_b200.outer_loop`` mirror ``vlapy.core.*`` / ``vlapy.outer_loop``
(same names, arguments and error behaviour) and run hand-written sm_100a CUDA kernels through the
C ABI of ``libvpfp_b200.so`` (include/vpfp_b200.h).  There is no CPU fallback.
"""
BACKEND_NAME = "b200"

from . import _lib  # noqa: E402,F401


def build(force=False, verbose=False):
    return _lib.build(force=force, verbose=verbose)
