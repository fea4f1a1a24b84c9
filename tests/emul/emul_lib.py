"""ctypes loader for the HOST EMULATION of the kernel programs (tests only; see vpfp_emul.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libvpfp_emul.so")
SRC = os.path.join(HERE, "vpfp_emul.cpp")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "vlapy_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, n) for n in ("advect.h", "rowops.h", "tridiag.h", "spline.h", "midfft.cuh", "tinyfft.cuh", "advect_fast.cuh", "vpfp_common.h", "butterflies.h", "rowfft.cuh")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
                               "-o", SO, SRC])
    return SO


SIMT_SO = os.path.join(HERE, "libvpfp_simt.so")
SIMT_SRC = os.path.join(HERE, "fp_simt.cpp")


def build_simt(force=False):
    """fp_fast.cuh / fp_reg.cuh compiled for the host with one OS thread per CUDA thread (simt.h)."""
    deps = [SIMT_SRC, os.path.join(HERE, "simt.h")] + [os.path.join(CSRC, n) for n in
                                                        ("fp_fast.cuh", "fp_reg.cuh", "vpfp_common.h")]
    if force or not os.path.exists(SIMT_SO) or any(os.path.getmtime(d) > os.path.getmtime(SIMT_SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-pthread", "-shared", "-fPIC",
                               "-o", SIMT_SO, SIMT_SRC])
    return SIMT_SO


_simt = None


def fp_simt(which, f, v, nu, dt, dv, op, grid=2):
    """Fokker-Planck kernels run thread by thread on the host: which = 0 fp_fast.cuh, 1 fp_reg.cuh.
    Returns (f_new, moments[8, rows])."""
    global _simt
    if _simt is None:
        _simt = ctypes.CDLL(build_simt())
    f = np.ascontiguousarray(f)
    rows, nv = f.shape
    out = np.empty_like(f)
    mom = np.zeros((8, rows))
    step = (v[-1] - v[0]) / (nv - 1)          # vlapy_b200.ops.linspace_params
    rc = _simt.emul_fp_simt(c_int(which), _p(f), c_long(nv), _p(out), c_long(nv), c_double(v[0]), c_double(step),
                            c_double(v[-1]), c_double(nu), c_double(dt), c_double(dv),
                            c_int(0 if op == "lb" else 1), _p(mom), c_long(rows), c_int(rows), c_int(nv), c_int(grid))
    assert rc == 0
    return out, mom


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


c_long, c_int, c_double = ctypes.c_long, ctypes.c_int, ctypes.c_double


def edfdv_exp(f, e, kv, dt, max_single=8192):
    f = np.ascontiguousarray(f); out = np.empty_like(f)
    rows, nv = f.shape
    lib().emul_edfdv_exp(_p(f), c_long(nv), _p(out), c_long(nv), _p(np.ascontiguousarray(e)),
                         _p(np.ascontiguousarray(kv)), c_double(dt), c_int(rows), c_int(nv), c_int(max_single))
    return out


def edfdv_rowfft(f, e, kv, dt):
    """single-pass row kernel (rowfft.cuh), nv in {4096, 8192, 16384}"""
    f = np.ascontiguousarray(f); out = np.empty_like(f)
    rows, nv = f.shape
    rc = lib().emul_edfdv_rowfft(_p(f), c_long(nv), _p(out), c_long(nv), _p(np.ascontiguousarray(e)),
                                 _p(np.ascontiguousarray(kv)), c_double(dt), c_int(rows), c_int(nv))
    assert rc == 0
    return out


def vdfdx_exp(f, kx, v, dt, batch=1, max_single=2048):
    f = np.ascontiguousarray(f); out = np.empty_like(f)
    ncols = f.shape[-1]
    nx = f.shape[-2]
    lib().emul_vdfdx_exp(_p(f), c_long(ncols), _p(out), c_long(ncols), _p(np.ascontiguousarray(kx)),
                         _p(np.ascontiguousarray(v)), c_double(dt), c_int(batch), c_int(nx), c_int(ncols),
                         c_int(max_single))
    return out


def poisson(n, ook, driver, max_single=8192):
    n = np.ascontiguousarray(np.atleast_2d(n)); batch, nx = n.shape
    ook = np.ascontiguousarray(np.broadcast_to(ook, n.shape))
    drv = None if driver is None else np.ascontiguousarray(np.broadcast_to(driver, n.shape))
    e = np.empty_like(n)
    lib().emul_poisson(_p(n), _p(ook), _p(drv), _p(e), c_int(batch), c_int(nx), c_int(max_single))
    return e


def poisson_rowfft(n, ook, driver):
    """single-launch Poisson solve (rowfft.cuh in Poisson mode), nx in {4096, 8192, 16384}"""
    n = np.ascontiguousarray(np.atleast_2d(n)); batch, nx = n.shape
    ook = np.ascontiguousarray(np.broadcast_to(ook, n.shape))
    drv = None if driver is None else np.ascontiguousarray(np.broadcast_to(driver, n.shape))
    e = np.empty_like(n)
    rc = lib().emul_poisson_rowfft(_p(n), _p(ook), _p(drv), _p(e), c_int(batch), c_int(nx))
    assert rc == 0
    return e


def edfdv_cd2(f, e, dt, dv):
    f = np.ascontiguousarray(f); out = np.empty_like(f); rows, nv = f.shape
    lib().emul_edfdv_cd2(_p(f), c_long(nv), _p(out), c_long(nv), _p(np.ascontiguousarray(e)), c_double(dt),
                         c_double(dv), c_int(rows), c_int(nv))
    return out


def moments(f, v, dv, nmom=8, edge_flags=3):
    f = np.ascontiguousarray(f); rows, ncols = f.shape
    out = np.zeros((nmom, rows))
    lib().emul_moments(_p(f), c_long(ncols), _p(np.ascontiguousarray(v)), c_double(dv), _p(out), c_long(rows),
                       c_int(nmom), c_int(rows), c_int(ncols), c_int(edge_flags))
    return out


def fp_step(f, v, nu, dt, dv, op, want_moments=False, m=0):
    f = np.ascontiguousarray(f); out = np.empty_like(f); rows, nv = f.shape
    mom = np.zeros((8, rows)) if want_moments else None
    lib().emul_fp_step(_p(f), c_long(nv), _p(out), c_long(nv), _p(np.ascontiguousarray(v)), c_double(nu),
                       c_double(dt), c_double(dv), c_int(0 if op == "lb" else 1), _p(mom), c_long(rows),
                       c_int(rows), c_int(nv), c_int(m))
    return (out, mom) if want_moments else out


def fp_diagonals(f, v, nu, dt, dv, op):
    f = np.ascontiguousarray(f); rows, nv = f.shape
    a, b, c = np.empty((rows, nv - 1)), np.empty((rows, nv)), np.empty((rows, nv - 1))
    lib().emul_fp_diagonals(_p(f), c_long(nv), _p(np.ascontiguousarray(v)), c_double(nu), c_double(dt), c_double(dv),
                            c_int(0 if op == "lb" else 1), _p(a), _p(b), _p(c), c_int(rows), c_int(nv))
    return a, b, c


def tridiag_solve(a, b, c, d, m=0):
    a, b, c, d = (np.ascontiguousarray(t, dtype=np.float64) for t in (a, b, c, d))
    rows, nv = d.shape
    x = np.empty_like(d)
    lib().emul_tridiag_solve(_p(a), _p(b), _p(c), _p(d), _p(x), c_int(rows), c_int(nv), c_int(m))
    return x


def vdfdx_sl(f, x, v, dt):
    f = np.ascontiguousarray(f); out = np.empty_like(f); nx, nv = f.shape
    rc = lib().emul_vdfdx_sl(_p(f), _p(out), _p(np.ascontiguousarray(x)), _p(np.ascontiguousarray(v)), c_double(dt),
                             c_double(x[2] - x[1]), c_int(nx), c_int(nv))
    assert rc == 0
    return out


def edfdv_sl(f, e, v, dt):
    f = np.ascontiguousarray(f); out = np.empty_like(f); nx, nv = f.shape
    rc = lib().emul_edfdv_sl(_p(f), _p(out), _p(np.ascontiguousarray(e)), _p(np.ascontiguousarray(v)), c_double(dt),
                             c_double(v[2] - v[1]), c_int(nx), c_int(nv))
    assert rc == 0
    return out


def midfft_rows(f, e, kv, dt):
    """single-pass mid-size e df/dv (midfft.cuh), nv in {256, 512, 1024, 2048}"""
    f = np.ascontiguousarray(f); out = np.empty_like(f); rows, nv = f.shape
    rc = lib().emul_midfft(c_int(1), _p(f), c_long(nv), _p(out), c_long(nv), _p(np.ascontiguousarray(kv)),
                           _p(np.ascontiguousarray(e)), c_double(dt), c_int(1), c_int(rows), c_int(nv))
    assert rc == 0
    return out


def midfft_cols(f, kx, v, dt, batch=1):
    """single-pass mid-size v df/dx (midfft.cuh), nx in {256, 512, 1024, 2048}; f (batch, nx, ncols)"""
    f = np.ascontiguousarray(f); out = np.empty_like(f)
    nx, ncols = f.shape[-2], f.shape[-1]
    rc = lib().emul_midfft(c_int(0), _p(f), c_long(ncols), _p(out), c_long(ncols), _p(np.ascontiguousarray(kx)),
                           _p(np.ascontiguousarray(v)), c_double(dt), c_int(batch), c_int(nx), c_int(ncols))
    assert rc == 0
    return out


def tiny_cols(f, kx, v, dt, batch=1):
    """v df/dx for nx = 16 / 32 with the transform in the registers of one thread (tinyfft.cuh); f (batch, nx, ncols)"""
    f = np.ascontiguousarray(f); out = np.empty_like(f)
    nx, ncols = f.shape[-2], f.shape[-1]
    rc = lib().emul_tiny_cols(_p(f), c_long(ncols), _p(out), c_long(ncols), _p(np.ascontiguousarray(kx)),
                              _p(np.ascontiguousarray(v)), c_double(dt), c_int(batch), c_int(nx), c_int(ncols))
    assert rc == 0
    return out


def tiny_poisson(n, ook, driver=None):
    """spectral Poisson solve for nx = 16 / 32 (tinyfft.cuh); n, ook (batch, nx), driver (batch, nx) or None"""
    n = np.ascontiguousarray(np.atleast_2d(n)); ook = np.ascontiguousarray(np.atleast_2d(ook))
    batch, nx = n.shape
    e = np.empty_like(n)
    d = np.ascontiguousarray(np.atleast_2d(driver)) if driver is not None else None
    rc = lib().emul_tiny_poisson(_p(n), _p(ook), _p(d) if d is not None else None, _p(e), c_int(batch), c_int(nx))
    assert rc == 0
    return e


def midfft_poisson(n, ook, driver=None):
    """spectral Poisson solve through midfft.cuh in Poisson mode; n, ook (batch, nx), driver (batch, nx) or None"""
    n = np.ascontiguousarray(np.atleast_2d(n)); ook = np.ascontiguousarray(np.atleast_2d(ook))
    batch, nx = n.shape
    e = np.empty_like(n)
    d = np.ascontiguousarray(np.atleast_2d(driver)) if driver is not None else None
    rc = lib().emul_midfft_poisson(_p(n), _p(ook), _p(d) if d is not None else None, _p(e), c_int(batch), c_int(nx))
    assert rc == 0
    return e


def midfft_cols_density(f, kx, v, dt, dv, edge_flags=3, batch=1):
    """the same with the charge density fused into the store phase; returns (f_out, n (batch, nx))"""
    f = np.ascontiguousarray(f); out = np.empty_like(f)
    nx, ncols = f.shape[-2], f.shape[-1]
    n = np.empty((batch, nx))
    rc = lib().emul_midfft_cols_density(_p(f), c_long(ncols), _p(out), c_long(ncols), _p(np.ascontiguousarray(kx)),
                                        _p(np.ascontiguousarray(v)), c_double(dt), c_int(batch), c_int(nx), c_int(ncols),
                                        c_double(dv), c_int(edge_flags), _p(n))
    assert rc == 0
    return out, n


def xmodes(f, nmodes=2):
    f = np.ascontiguousarray(f); nx, ncols = f.shape
    out = np.zeros((nmodes, ncols, 2))
    lib().emul_xmodes(_p(f), c_long(ncols), _p(out), c_int(nmodes), c_int(1), c_int(nx), c_int(ncols))
    return out[..., 0] + 1j * out[..., 1]


def driver(x, t, pulses):
    arr = np.array([[p["k0"], p["w0"], p["a0"], p["t_L"], p["t_R"], p["t_wL"], p["t_wR"]]
                    for p in pulses.values()], dtype=np.float64)
    out = np.empty_like(x)
    lib().emul_driver(_p(np.ascontiguousarray(x)), c_double(t), _p(arr), c_int(arr.shape[0]), _p(out), c_int(x.size))
    return out


def series(mom, e, de):
    out = np.zeros(7)
    mom = np.ascontiguousarray(mom)
    lib().emul_series(_p(mom), c_long(mom.shape[1]), _p(np.ascontiguousarray(e)), _p(np.ascontiguousarray(de)),
                      _p(out), c_int(e.size))
    return out


def butterfly(x, radix, direction):
    """register butterflies of the fast kernels: natural order in/out, direction -1 fwd / +1 inv"""
    xy = np.ascontiguousarray(np.stack([x.real, x.imag], axis=1).astype(np.float64))
    rc = lib().emul_butterfly(_p(xy), c_int(radix), c_int(direction))
    assert rc == 0
    return xy[:, 0] + 1j * xy[:, 1]
