"""Device-tensor level operators: one function per C-ABI entry point.

Inputs are CUDA fp64 torch tensors (torch is used for device memory and streams only); every call
enqueues hand-written sm_100a kernels on torch's current stream and returns without synchronising.
"""
import ctypes

import numpy as np
import torch

from . import _lib

FP_OPS = {"lb": 0, "dg": 1}
PHASE_EXACT, PHASE_TABLE, FORCE_GENERIC, FORCE_THREE_PASS = 0, 1, 2, 4


def launch_count(reset=False):
    """kernels this library has enqueued since the last reset, counted at the launch sites inside the
    library (vpfp_launch_count); kernels replayed by a CUDA graph are counted once, at capture"""
    return int(_lib.lib().vpfp_launch_count(1 if reset else 0))


def scratch_generation():
    """number of reallocations of the library's reduction scratch so far (vpfp_scratch_generation)"""
    return int(_lib.lib().vpfp_scratch_generation())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_f(f, name="f"):
    if not (isinstance(f, torch.Tensor) and f.is_cuda and f.dtype == torch.float64):
        raise TypeError("%s must be a CUDA float64 tensor" % name)
    if f.dim() < 2 or f.stride(-1) != 1:
        raise ValueError("%s must have a contiguous last (v) axis" % name)
    rows = int(np.prod(f.shape[:-1]))
    ld = f.stride(-2)
    # leading dims must be collapsible to a uniform row pitch
    exp = ld
    for d in range(f.dim() - 2, -1, -1):
        if f.shape[d] != 1 and f.stride(d) != exp:
            raise ValueError("%s must have a uniform row pitch" % name)
        exp *= f.shape[d]
    return rows, ld


def _vec(x, n, name):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()):
        raise TypeError("%s must be a contiguous CUDA float64 tensor" % name)
    if x.numel() != n:
        raise ValueError("%s has %d elements, expected %d" % (name, x.numel(), n))
    return x


def _is_pow2(n):
    return n > 0 and (n & (n - 1)) == 0


def edfdv_exp(f, e, kv, dt, out=None, flags=PHASE_EXACT):
    """vlapy/core/vlasov.py:123-138 on device. f: (..., nx, nv); e: (..., nx); kv: (nv)."""
    rows, ld = _chk_f(f)
    nv = f.shape[-1]
    if out is None:
        out = torch.empty(f.shape, dtype=f.dtype, device=f.device)
    _, ldo = _chk_f(out, "out")
    _vec(e, rows, "e"); _vec(kv, nv, "kv")
    _lib.check(_lib.lib().vpfp_edfdv_exp(f.data_ptr(), ld, out.data_ptr(), ldo, e.data_ptr(), kv.data_ptr(),
                                         float(dt), rows, nv, flags, _stream()))
    return out


def rowfft_serves(rows, nv, flags):
    """whether e df/dv runs as the single-pass row kernel (csrc/rowfft.cuh; mirrors rowfft_eligible)"""
    return bool(flags & PHASE_TABLE) and not (flags & (FORCE_GENERIC | FORCE_THREE_PASS)) \
        and nv in (4096, 8192, 16384) and rows * nv >= (1 << 22)


def vdfdx_exp(f, kx, v, dt, out=None, flags=PHASE_EXACT, density_out=None, dv=None, edge_flags=3):
    """vlapy/core/vlasov.py:94-108 on device. f: (batch, nx, ncols) or (nx, ncols); kx: (batch, nx).
    density_out (batch*nx doubles) + dv: also return trapz_v of the result (fused epilogue)."""
    _, ld = _chk_f(f)
    nx, ncols = f.shape[-2], f.shape[-1]
    batch = int(np.prod(f.shape[:-2])) if f.dim() > 2 else 1
    if out is None:
        out = torch.empty(f.shape, dtype=f.dtype, device=f.device)
    _, ldo = _chk_f(out, "out")
    _vec(kx, batch * nx, "kx"); _vec(v, ncols, "v")
    if density_out is not None:
        _vec(density_out, batch * nx, "density_out")
        _lib.check(_lib.lib().vpfp_vdfdx_exp_density(f.data_ptr(), ld, out.data_ptr(), ldo, kx.data_ptr(),
                                                     v.data_ptr(), float(dt), batch, nx, ncols, flags,
                                                     density_out.data_ptr(), float(dv), edge_flags, _stream()))
        return out
    _lib.check(_lib.lib().vpfp_vdfdx_exp(f.data_ptr(), ld, out.data_ptr(), ldo, kx.data_ptr(), v.data_ptr(),
                                         float(dt), batch, nx, ncols, flags, _stream()))
    return out


class PeerBuffer:
    """A device buffer that peer GPUs (other processes on the same NVLink domain) can address:
    cudaMalloc + CUDA IPC handle (vpfp_ipc_alloc).  ``tensor(shape)`` views it as a torch tensor
    (torch does not own the memory; keep this object alive while the view is used)."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        ptr = ctypes.c_void_p()
        hbuf = ctypes.create_string_buffer(64)
        _lib.check(_lib.lib().vpfp_ipc_alloc(self.nbytes, ctypes.byref(ptr), hbuf))
        self.ptr, self.handle = ptr.value, hbuf.raw
        self.opened = []

    def tensor(self, shape, device):
        class _View:
            pass
        v = _View()
        v.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": "<f8",
                                      "data": (self.ptr, False), "version": 3, "strides": None}
        t = torch.as_tensor(v, device=device)
        t._vpfp_owner = self
        return t

    def open_peer(self, handle):
        """map a peer's buffer (its 64-byte handle) into this process; returns the device pointer"""
        ptr = ctypes.c_void_p()
        _lib.check(_lib.lib().vpfp_ipc_open(handle, ctypes.byref(ptr)))
        self.opened.append(ptr.value)
        return ptr.value

    def close(self):
        for p in self.opened:
            _lib.lib().vpfp_ipc_close(ctypes.c_void_p(p))
        self.opened = []
        if self.ptr:
            _lib.lib().vpfp_ipc_free(ctypes.c_void_p(self.ptr))
            self.ptr = None


def _ptr_array(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(p) for p in ptrs])


def edfdv_exp_scatter(f, e, kv, dt, scratch, peer_ptrs, my_rank, flags=PHASE_EXACT):
    """e df/dv of an x-shard (rows, nv) whose result lands v-sharded on the peers: column block q is
    stored over NVLink into peer_ptrs[q] (rank q's (rows*P, nv/P) shard) at rows
    [my_rank*rows, (my_rank+1)*rows).  scratch: (rows, nv) for the intermediate passes."""
    rows, ld = _chk_f(f)
    nv = f.shape[-1]
    _, lds = _chk_f(scratch, "scratch")
    _vec(e, rows, "e"); _vec(kv, nv, "kv")
    _lib.check(_lib.lib().vpfp_edfdv_exp_scatter(f.data_ptr(), ld, scratch.data_ptr(), lds, e.data_ptr(),
                                                 kv.data_ptr(), float(dt), rows, nv, flags, _ptr_array(peer_ptrs),
                                                 len(peer_ptrs), int(my_rank), _stream()))


def vdfdx_exp_scatter(f, kx, v, dt, scratch, peer_ptrs, my_rank, flags=PHASE_EXACT, density_out=None, dv=0.0,
                      edge_flags=3):
    """v df/dx of a v-shard (nx, ncols) whose result lands x-sharded on the peers: row block q is
    stored into peer_ptrs[q] (rank q's (nx/P, ncols*P) shard) at columns
    [my_rank*ncols, (my_rank+1)*ncols); density_out (nx) receives the partial trapz_v of the result."""
    _, ld = _chk_f(f)
    nx, ncols = f.shape[-2], f.shape[-1]
    _, lds = _chk_f(scratch, "scratch")
    _vec(kx, nx, "kx"); _vec(v, ncols, "v")
    if density_out is not None:
        _vec(density_out, nx, "density_out")
    _lib.check(_lib.lib().vpfp_vdfdx_exp_scatter(f.data_ptr(), ld, scratch.data_ptr(), lds, kx.data_ptr(),
                                                 v.data_ptr(), float(dt), nx, ncols, flags,
                                                 density_out.data_ptr() if density_out is not None else None,
                                                 float(dv), edge_flags, _ptr_array(peer_ptrs), len(peer_ptrs),
                                                 int(my_rank), _stream()))


def edfdv_cd2(f, e, dt, dv, out=None):
    """vlapy/core/vlasov.py:153-163 on device."""
    rows, ld = _chk_f(f)
    nv = f.shape[-1]
    if out is None:
        out = torch.empty(f.shape, dtype=f.dtype, device=f.device)
    _, ldo = _chk_f(out, "out")
    _vec(e, rows, "e")
    _lib.check(_lib.lib().vpfp_edfdv_cd2(f.data_ptr(), ld, out.data_ptr(), ldo, e.data_ptr(), float(dt),
                                         float(dv), rows, nv, _stream()))
    return out


def vdfdx_sl(f, x, v, dt, dx, out=None):
    """vlapy/core/vlasov.py:42-80 on device: semi-Lagrangian x advection of f (nx, nv); x, v device axes."""
    _, ld = _chk_f(f)
    if f.dim() != 2:
        raise ValueError("the semi-Lagrangian operators take one (nx, nv) grid")
    nx, nv = f.shape
    if out is None:
        out = torch.empty(f.shape, dtype=f.dtype, device=f.device)
    _, ldo = _chk_f(out, "out")
    _vec(x, nx, "x"); _vec(v, nv, "v")
    _lib.check(_lib.lib().vpfp_vdfdx_sl(f.data_ptr(), ld, out.data_ptr(), ldo, x.data_ptr(), v.data_ptr(), float(dt),
                                        float(dx), nx, nv, _stream()))
    return out


def edfdv_sl(f, e, v, dt, dv, out=None):
    """vlapy/core/vlasov.py:168-210 on device: semi-Lagrangian v advection of f (nx, nv); e (nx), v (nv)."""
    _, ld = _chk_f(f)
    if f.dim() != 2:
        raise ValueError("the semi-Lagrangian operators take one (nx, nv) grid")
    nx, nv = f.shape
    if out is None:
        out = torch.empty(f.shape, dtype=f.dtype, device=f.device)
    _, ldo = _chk_f(out, "out")
    _vec(e, nx, "e"); _vec(v, nv, "v")
    _lib.check(_lib.lib().vpfp_edfdv_sl(f.data_ptr(), ld, out.data_ptr(), ldo, e.data_ptr(), v.data_ptr(), float(dt),
                                        float(dv), nx, nv, _stream()))
    return out


def moments(f, v, dv, nmom=8, out=None, edge_flags=3):
    """Rows of v-moments (nmom, rows): n, j, T, q, fv4, vN, int f^2, int f ln f."""
    rows, ld = _chk_f(f)
    ncols = f.shape[-1]
    _vec(v, ncols, "v")
    if out is None:
        out = torch.empty((nmom, rows), dtype=f.dtype, device=f.device)
    _lib.check(_lib.lib().vpfp_moments(f.data_ptr(), ld, v.data_ptr(), float(dv), out.data_ptr(),
                                       out.stride(0), nmom, rows, ncols, edge_flags, _stream()))
    return out


def poisson(n, one_over_kx, driver=None, out=None):
    """vlapy/core/field.py:39-88 on device: e = driver + Re ifft(i one_over_kx fft(1 - n))."""
    nx = n.shape[-1]
    batch = n.numel() // nx
    _vec(n, batch * nx, "n"); _vec(one_over_kx, batch * nx, "one_over_kx")
    if driver is not None:
        _vec(driver, batch * nx, "driver")
    if out is None:
        out = torch.empty_like(n)
    _lib.check(_lib.lib().vpfp_poisson(n.data_ptr(), one_over_kx.data_ptr(),
                                       driver.data_ptr() if driver is not None else None,
                                       out.data_ptr(), batch, nx, _stream()))
    return out


def linspace_params(v):
    """(v0, step, vlast) when the host array v is bit-for-bit np.linspace(v[0], v[-1], len(v)) --
    how vlapy/initializers.py:66 builds the velocity grid -- else None."""
    v = np.asarray(v, dtype=np.float64)
    n = v.size
    if n < 2 or not np.array_equal(v, np.linspace(v[0], v[-1], n)):
        return None
    return float(v[0]), float((v[-1] - v[0]) / (n - 1)), float(v[-1])


FAST_FP_SIZES = (128, 256, 512, 1024, 2048, 4096, 8192, 16384)


def fp_step(f, v, nu, dt, dv, op="lb", out=None, moments_out=None, vgrid=None):
    """Implicit LB / Dougherty step (vlapy/core/collisions.py via step.py:102-108) on device.
    vgrid = linspace_params(v) selects the kernel specialised for np.linspace velocity grids."""
    rows, ld = _chk_f(f)
    nv = f.shape[-1]
    if op not in FP_OPS:
        raise NotImplementedError("Collision Operator: <" + str(op) + "> has not yet been implemented on the b200 backend")
    _vec(v, nv, "v")
    if out is None:
        out = torch.empty(f.shape, dtype=f.dtype, device=f.device)
    _, ldo = _chk_f(out, "out")
    mp, mld = (None, 0) if moments_out is None else (moments_out.data_ptr(), moments_out.stride(0))
    if vgrid is not None and nv in FAST_FP_SIZES:
        _lib.check(_lib.lib().vpfp_fp_step_linspace(f.data_ptr(), ld, out.data_ptr(), ldo, vgrid[0], vgrid[1],
                                                    vgrid[2], float(nu), float(dt), float(dv), FP_OPS[op], mp, mld,
                                                    rows, nv, _stream()))
        return out
    _lib.check(_lib.lib().vpfp_fp_step(f.data_ptr(), ld, out.data_ptr(), ldo, v.data_ptr(), float(nu), float(dt),
                                       float(dv), FP_OPS[op], mp, mld, rows, nv, _stream()))
    return out


def fp_diagonals(f, v, nu, dt, dv, op="lb"):
    """vlapy/core/collisions.py:44-81 / 104-158 on device: f (rows, nv) -> a (rows, nv-1), b (rows, nv), c (rows, nv-1)."""
    rows, ld = _chk_f(f)
    nv = f.shape[-1]
    if op not in FP_OPS:
        raise NotImplementedError("Collision Operator: <" + str(op) + "> has not yet been implemented on the b200 backend")
    _vec(v, nv, "v")
    a = torch.empty((rows, nv - 1), dtype=f.dtype, device=f.device)
    b = torch.empty((rows, nv), dtype=f.dtype, device=f.device)
    c = torch.empty((rows, nv - 1), dtype=f.dtype, device=f.device)
    _lib.check(_lib.lib().vpfp_fp_diagonals(f.data_ptr(), ld, v.data_ptr(), float(nu), float(dt), float(dv), FP_OPS[op],
                                            a.data_ptr(), nv - 1, b.data_ptr(), nv, c.data_ptr(), nv - 1, rows, nv,
                                            _stream()))
    return a, b, c


def tridiag_solve(a, b, c, d, out=None):
    """vlapy/core/collisions.py:232-263 on device: one tridiagonal system per row, general diagonals.
    a, c: (rows, nv-1); b, d: (rows, nv); returns a new (rows, nv) tensor."""
    rows, ldd = _chk_f(d, "d")
    nv = d.shape[-1]
    for t, n, name in ((a, nv - 1, "a"), (b, nv, "b"), (c, nv - 1, "c")):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.dim() == 2
                and t.shape == (rows, n) and t.stride(1) == 1):
            raise ValueError("%s must be a CUDA float64 tensor of shape (%d, %d) with a contiguous last axis" % (name, rows, n))
    if out is None:
        out = torch.empty((rows, nv), dtype=d.dtype, device=d.device)
    _, ldx = _chk_f(out, "out")
    _lib.check(_lib.lib().vpfp_tridiag_solve(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), c.data_ptr(),
                                             c.stride(0), d.data_ptr(), ldd, out.data_ptr(), ldx, rows, nv, _stream()))
    return out


def xmodes(f, nmodes=2, out=None, x_offset=0, nx_total=None):
    """fft_x(f)[:nmodes] per v as (batch, nmodes, ncols) complex128 (vlapy/core/step.py:130-135).
    With x_offset / nx_total: the partial sums of an x-slab of a larger grid."""
    _, ld = _chk_f(f)
    nx, ncols = f.shape[-2], f.shape[-1]
    batch = int(np.prod(f.shape[:-2])) if f.dim() > 2 else 1
    if out is None:
        out = torch.empty((batch, nmodes, ncols, 2), dtype=f.dtype, device=f.device)
    _lib.check(_lib.lib().vpfp_xmodes_partial(f.data_ptr(), ld, out.data_ptr(), nmodes, batch, nx, ncols,
                                              int(x_offset), int(nx_total or nx), _stream()))
    return torch.view_as_complex(out)


def driver(x, t, pulses, out=None):
    """vlapy/field_driver.py:24-50 on device. pulses: host float64 array (npulse, 7)."""
    pulses = np.ascontiguousarray(pulses, dtype=np.float64).reshape(-1, 7)
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.lib().vpfp_driver(x.data_ptr(), float(t), pulses.ctypes.data_as(ctypes.c_void_p),
                                      pulses.shape[0], out.data_ptr(), x.numel(), _stream()))
    return out


class DevTime:
    """A time whose base lives in a device scalar (so that a captured CUDA graph can be replayed
    with a new time); ``+ float`` appends an increment, applied one by one on the device in the
    order the schedule wrote them (vlapy/core/vlasov_poisson.py:116-148 sums left to right)."""
    __slots__ = ("base", "incs")

    def __init__(self, base, incs=()):
        self.base, self.incs = base, tuple(incs)

    def __add__(self, other):
        return DevTime(self.base, self.incs + (float(other),))

    __radd__ = __add__


def driver_dev(x, t, pulses, out=None):
    """driver(x, t) with t a DevTime"""
    pulses = np.ascontiguousarray(pulses, dtype=np.float64).reshape(-1, 7)
    incs = np.asarray(t.incs, dtype=np.float64)
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.lib().vpfp_driver_dev(x.data_ptr(), t.base.data_ptr(), incs.ctypes.data_as(ctypes.c_void_p),
                                          incs.size, pulses.ctypes.data_as(ctypes.c_void_p), pulses.shape[0],
                                          out.data_ptr(), x.numel(), _stream()))
    return out


def series(mom, e, de, out=None):
    """vlapy/core/step.py:202-224 on device: 7 x-means of one stored step."""
    if out is None:
        out = torch.empty(7, dtype=mom.dtype, device=mom.device)
    _lib.check(_lib.lib().vpfp_series(mom.data_ptr(), mom.stride(0), e.data_ptr(),
                                      de.data_ptr() if de is not None else None, out.data_ptr(),
                                      e.numel(), _stream()))
    return out


def pulses_to_array(pulse_dictionary):
    """{name: {k0, w0, a0, t_L, t_R, t_wL, t_wR, ...}} -> (npulse, 7) float64 in ABI order."""
    return np.array([[p["k0"], p["w0"], p["a0"], p["t_L"], p["t_R"], p["t_wL"], p["t_wR"]]
                     for p in pulse_dictionary.values()], dtype=np.float64).reshape(-1, 7)


def profile_enable(on=True):
    """bracket every kernel launch of the library with CUDA events (bench.py roofline block)"""
    _lib.check(_lib.lib().vpfp_profile_enable(1 if on else 0))


def profile_report():
    """{label: (launches, total_ms)} since profiling was enabled; synchronises the device"""
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(_lib.lib().vpfp_profile_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        label, n, ms = line.split()
        out[label] = (int(n), float(ms))
    return out
