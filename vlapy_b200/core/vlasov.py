"""Mirror of vlapy/core/vlasov.py for the b200 backend: x- and v-advection operator factories.

Same factory names, closure signatures and string dispatch as the reference
(vlapy/core/vlasov.py:83-140, 143-165, 213-260).  Closures are functional (a new array is
returned, the argument is never modified) and accept numpy arrays or CUDA tensors, returning the
same kind.
"""
import os

import numpy as np

from .. import ops
from .._util import attach_density, back, const, to_dev


def _check_wavenumbers(k, what):
    """The kernels evaluate the phase for bins 0..N/2 and use exp(-i k[N-m] c) = conj(exp(-i k[m] c)),
    which holds for any np.fft.fftfreq-style grid (k[N-m] == -k[m], as the reference builds it in
    vlapy/initializers.py:67,85)."""
    k = np.asarray(k, dtype=np.float64)
    n = k.shape[-1]
    if n > 2 and not np.array_equal(k[..., 1:(n + 1) // 2], -k[..., :n // 2:-1]):
        raise NotImplementedError(
            what + ": wavenumber grids that are not antisymmetric (np.fft.fftfreq layout) "
            "have not yet been implemented on the b200 backend")
    return k


def _phase_flags(k):
    """Geometric phase tables need a uniform grid k[m] = m k[1] (np.fft.fftfreq); anything else, or
    VLAPY_B200_PHASE=exact, selects the per-bin sincos of theta = (k dt) c."""
    mode = os.environ.get("VLAPY_B200_PHASE", "table")
    if mode == "exact":
        return ops.PHASE_EXACT
    k = np.asarray(k, dtype=np.float64)
    n = k.shape[-1]
    m = np.arange(n // 2 + 1)
    lin = m * k[..., 1:2]
    half = k[..., : n // 2 + 1].copy()
    half[..., n // 2] = -half[..., n // 2] if n % 2 == 0 else half[..., n // 2]   # fftfreq stores the Nyquist bin as -N/2
    # signed comparison: the table kernels assume k[m] = m k[1] (the antisymmetric upper half follows from
    # _check_wavenumbers)
    if n >= 4 and np.all(np.abs(half - lin) <= 8 * np.finfo(float).eps * np.abs(lin)):
        return ops.PHASE_TABLE
    return ops.PHASE_EXACT


def get_vdfdx_exponential(kx, v, dv=None):
    """vlapy/core/vlasov.py:83-110 -- v df/dx exponential integrator.

    kx may be (nx,) or (batch, nx) for ensembles of simulations with their own box length.
    With ``dv`` the charge density of the result is reduced in the kernel's epilogue and attached
    to the returned device tensor (``_util.attach_density``) for the field solve that always follows
    (vlapy/core/vlasov_poisson.py:54-55); an in-place change of the tensor in between invalidates it
    (the tensor's version is recorded) and the field solve integrates f itself."""
    kx_d = const(_check_wavenumbers(kx, "v df/dx"))
    v_d = const(v)
    flags = _phase_flags(kx)

    def step_vdfdx_exponential(f, dt):
        f_d, host = to_dev(f)
        if dv is None or host:
            return back(ops.vdfdx_exp(f_d.contiguous(), kx_d, v_d, dt, flags=flags), host)
        n = f_d.new_empty(f_d.shape[:-1])
        out = ops.vdfdx_exp(f_d.contiguous(), kx_d, v_d, dt, flags=flags, density_out=n, dv=dv)
        return attach_density(out, n)

    return step_vdfdx_exponential


def get_edfdv_exponential(kv):
    """vlapy/core/vlasov.py:113-140 -- e df/dv exponential integrator."""
    kv_d = const(_check_wavenumbers(kv, "e df/dv"))
    flags = _phase_flags(kv)

    def step_edfdv_exponential(f, e, dt):
        f_d, host = to_dev(f)
        e_d, _ = to_dev(e)
        return back(ops.edfdv_exp(f_d.contiguous(), e_d.contiguous(), kv_d, dt, flags=flags), host)

    return step_edfdv_exponential


def get_edfdv_center_differenced(dv):
    """vlapy/core/vlasov.py:143-165 -- f - e * gradient_v(f) * dt, 2nd-order edges."""

    def step_edfdv_center_difference(f, e, dt):
        f_d, host = to_dev(f)
        e_d, _ = to_dev(e)
        return back(ops.edfdv_cd2(f_d.contiguous(), e_d.contiguous(), dt, dv), host)

    return step_edfdv_center_difference


def _axis_spacing(ax):
    """the spacing the reference pads with (vlapy/core/vlasov.py:35-37): ax[2] - ax[1]"""
    ax = np.asarray(ax, dtype=np.float64)
    if ax.ndim != 1 or ax.size < 4:
        raise NotImplementedError("<sl> needs a one-dimensional axis of at least four points")
    return float(ax[2] - ax[1])


def get_vdfdx_sl(x, v):
    """vlapy/core/vlasov.py:42-80 -- backward semi-Lagrangian v df/dx: cubic spline along x through f padded with
    one periodic ghost row on either side, evaluated at x - v dt (csrc/spline.h)."""
    x_d, v_d = const(x), const(v)
    dx = _axis_spacing(x)

    def update_spatial_adv_sl(f, dt):
        f_d, host = to_dev(f)
        return back(ops.vdfdx_sl(f_d.contiguous(), x_d, v_d, dt, dx), host)

    return update_spatial_adv_sl


def get_edfdv_sl(x, v):
    """vlapy/core/vlasov.py:168-210 -- backward semi-Lagrangian e df/dv: cubic spline along v, evaluated at
    v - e dt (the reference's cubic interp1d of e is evaluated at its own nodes, i.e. it returns e)."""
    v_d = const(v)
    dv = _axis_spacing(v)
    _axis_spacing(x)

    def update_velocity_adv_sl(f, e, dt):
        f_d, host = to_dev(f)
        e_d, _ = to_dev(e)
        return back(ops.edfdv_sl(f_d.contiguous(), e_d.contiguous(), v_d, dt, dv), host)

    return update_velocity_adv_sl


def get_vdfdx(stuff_for_time_loop, vdfdx_implementation="exponential"):
    """vlapy/core/vlasov.py:213-235."""
    if vdfdx_implementation == "exponential":
        vdfdx = get_vdfdx_exponential(kx=stuff_for_time_loop["kx"], v=stuff_for_time_loop["v"],
                                      dv=stuff_for_time_loop.get("dv"))
    elif vdfdx_implementation == "sl":
        vdfdx = get_vdfdx_sl(x=stuff_for_time_loop["x"], v=stuff_for_time_loop["v"])
    else:
        raise NotImplementedError(
            "v df/dx: <" + vdfdx_implementation + "> has not yet been implemented on the b200 backend")
    return vdfdx


def get_edfdv(stuff_for_time_loop, edfdv_implementation="exponential"):
    """vlapy/core/vlasov.py:238-260."""
    if edfdv_implementation == "exponential":
        edfdv = get_edfdv_exponential(kv=stuff_for_time_loop["kv"])
    elif edfdv_implementation == "cd2":
        edfdv = get_edfdv_center_differenced(dv=stuff_for_time_loop["dv"])
    elif edfdv_implementation == "sl":
        edfdv = get_edfdv_sl(v=stuff_for_time_loop["v"], x=stuff_for_time_loop["x"])
    else:
        raise NotImplementedError(
            "e df/dv: <" + edfdv_implementation + "> has not yet been implemented on the b200 backend")
    return edfdv
