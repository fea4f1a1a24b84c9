"""ctypes binding of libvpfp_b200.so (C ABI: include/vpfp_b200.h).

There is no CPU fallback: if the CUDA library is missing or a call fails this module raises.
"""
import ctypes
import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
# VPFP_B200_LIB: another build of the same CUDA library (A/B timing of two revisions in one GPU session)
SO = os.environ.get("VPFP_B200_LIB") or os.path.join(PKG, "lib", "libvpfp_b200.so")
SRC = os.path.join(PKG, "csrc", "vpfp_cuda.cu")
HEADERS = [os.path.join(PKG, "csrc", n) for n in ("vpfp_common.h", "advect.h", "rowops.h", "butterflies.h", "rowfft.cuh", "midfft.cuh", "tinyfft.cuh", "tridiag.h", "spline.h",
                                                    "advect_fast.cuh", "fp_fast.cuh", "fp_reg.cuh")] + [
    os.path.join(ROOT, "include", "vpfp_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
ABI_VERSION = 1


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    deps = [SRC] + HEADERS
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC]
    subprocess.check_call(cmd)
    return SO


_lib = None

_c = ctypes
_P, _L, _I, _D = _c.c_void_p, _c.c_long, _c.c_int, _c.c_double
_SIGS = {
    "vpfp_abi_version": ([], _I),
    "vpfp_last_error": ([], _c.c_char_p),
    "vpfp_shutdown": ([], _I),
    "vpfp_launch_count": ([_I], _L),
    "vpfp_scratch_generation": ([], _c.c_ulong),
    "vpfp_profile_enable": ([_I], _I),
    "vpfp_profile_report": ([_c.c_char_p, _I], _I),
    "vpfp_edfdv_exp": ([_P, _L, _P, _L, _P, _P, _D, _I, _I, _I, _P], _I),
    "vpfp_vdfdx_exp": ([_P, _L, _P, _L, _P, _P, _D, _I, _I, _I, _I, _P], _I),
    "vpfp_vdfdx_exp_density": ([_P, _L, _P, _L, _P, _P, _D, _I, _I, _I, _I, _P, _D, _I, _P], _I),
    "vpfp_edfdv_cd2": ([_P, _L, _P, _L, _P, _D, _D, _I, _I, _P], _I),
    "vpfp_vdfdx_sl": ([_P, _L, _P, _L, _P, _P, _D, _D, _I, _I, _P], _I),
    "vpfp_edfdv_sl": ([_P, _L, _P, _L, _P, _P, _D, _D, _I, _I, _P], _I),
    "vpfp_moments": ([_P, _L, _P, _D, _P, _L, _I, _I, _I, _I, _P], _I),
    "vpfp_poisson": ([_P, _P, _P, _P, _I, _I, _P], _I),
    "vpfp_fp_step": ([_P, _L, _P, _L, _P, _D, _D, _D, _I, _P, _L, _I, _I, _P], _I),
    "vpfp_fp_step_linspace": ([_P, _L, _P, _L, _D, _D, _D, _D, _D, _D, _I, _P, _L, _I, _I, _P], _I),
    "vpfp_fp_diagonals": ([_P, _L, _P, _D, _D, _D, _I, _P, _L, _P, _L, _P, _L, _I, _I, _P], _I),
    "vpfp_tridiag_solve": ([_P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _P], _I),
    "vpfp_xmodes": ([_P, _L, _P, _I, _I, _I, _I, _P], _I),
    "vpfp_xmodes_partial": ([_P, _L, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "vpfp_driver": ([_P, _D, _P, _I, _P, _I, _P], _I),
    "vpfp_driver_dev": ([_P, _P, _P, _I, _P, _I, _P, _I, _P], _I),
    "vpfp_ipc_alloc": ([_c.c_size_t, _c.POINTER(_P), _c.c_char_p], _I),
    "vpfp_ipc_open": ([_c.c_char_p, _c.POINTER(_P)], _I),
    "vpfp_ipc_close": ([_P], _I),
    "vpfp_ipc_free": ([_P], _I),
    "vpfp_edfdv_exp_scatter": ([_P, _L, _P, _L, _P, _P, _D, _I, _I, _I, _c.POINTER(_P), _I, _I, _P], _I),
    "vpfp_vdfdx_exp_scatter": ([_P, _L, _P, _L, _P, _P, _D, _I, _I, _I, _P, _D, _I, _c.POINTER(_P), _I, _I, _P], _I),
    "vpfp_series_batch": ([_P, _L, _P, _P, _P, _I, _I, _P], _I),
    "vpfp_driver_batch": ([_P, _D, _P, _P, _I, _P, _I, _P, _I, _I, _P], _I),
    "vpfp_series": ([_P, _L, _P, _P, _P, _I, _P], _I),
}


def lib():
    """Load the library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise RuntimeError(
            "vlapy_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the b200 backend." % SO)
    handle = ctypes.CDLL(SO)
    for name, (argtypes, restype) in _SIGS.items():
        fn = getattr(handle, name)  # AttributeError here = ABI mismatch, fail loudly
        fn.argtypes = argtypes
        fn.restype = restype
    if handle.vpfp_abi_version() != ABI_VERSION:
        raise RuntimeError("vlapy_b200: libvpfp_b200.so ABI %d != expected %d" %
                           (handle.vpfp_abi_version(), ABI_VERSION))
    _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc):
    if rc == OK:
        return
    msg = lib().vpfp_last_error().decode()
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError(msg)
