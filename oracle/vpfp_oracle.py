"""CPU oracle for the VPFP phase-space update  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/scipy restatement of the reference algorithm for the hot path named by
BASELINE.json (spectral x/v advection, pseudospectral Poisson, implicit LB/Dougherty
Fokker-Planck step, the splitting schedules and the per-step moments).  Every function cites
the reference file:line (paths relative to the VlaPy source tree) whose arithmetic it follows,
including evaluation order where that matters at the 1e-12 level.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  Nothing under ``vlapy_b200/`` imports it; the product
path fails loudly when the CUDA library is missing.

Parity pinning: ``tests/golden/make_golden.py`` runs the *reference itself* (imported from its
source tree in the build container) on seeded inputs and commits the outputs as fixtures;
``tests/test_oracle_vs_golden.py`` checks this oracle against them bit-for-tolerance (1e-13)
together with the known values of SURVEY.md Appendix B.  The arithmetic itself lives in
un-vendored, un-pinned third-party code (numpy ufuncs, scipy.fft/pocketfft; reference
setup.py:25-35); fixtures were generated with numpy 2.3.5 / scipy 1.18.1.
"""
from __future__ import annotations

import numpy as np
from scipy import fft as _sfft

# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------


def trapz_last(y, dx):
    """np.trapz(y, dx=dx, axis=-1) restated (numpy's own formula d*(y[1:]+y[:-1])/2 summed).

    Reference call sites: vlapy/core/field.py:36, collisions.py:57-65,117-137, step.py:166-171.
    """
    y = np.asarray(y)
    return (dx * (y[..., 1:] + y[..., :-1]) / 2.0).sum(axis=-1)


# --------------------------------------------------------------------------------------------
# grids and initial state  (vlapy/initializers.py)
# --------------------------------------------------------------------------------------------


def velocity_grid(vmax, nv):
    """initializers.py:56-69 -- cell-centred v grid, dv, and kv = 2*pi*fftfreq."""
    dv = 2 * vmax / nv
    v = np.linspace(-vmax + dv / 2.0, vmax - dv / 2.0, nv)
    kv = np.fft.fftfreq(v.size, d=dv) * 2.0 * np.pi
    return dv, v, kv


def spatial_grid(xmin, xmax, nx):
    """initializers.py:72-89 -- cell-centred x grid, kx and 1/kx with the k=0 entry zeroed."""
    dx = (xmax - xmin) / nx
    x = np.linspace(xmin + dx / 2.0, xmax - dx / 2.0, nx)
    kx = np.fft.fftfreq(x.size, d=dx) * 2.0 * np.pi
    one_over_kx = np.zeros_like(kx)
    one_over_kx[1:] = 1.0 / kx[1:]
    return dx, x, kx, one_over_kx


def maxwellian(nx, nv, vmax=6.0):
    """initializers.py:29-53 -- unit Maxwellian on every x row, trapz-normalised."""
    dv = 2.0 * vmax / nv
    vax = np.linspace(-vmax + dv / 2.0, vmax - dv / 2.0, nv)
    f = np.zeros([nx, nv], dtype=np.float64)
    f[:, :] = np.exp(-(vax ** 2.0) / 2.0)
    return f / trapz_last(f, dv)[:, None]


def shifted_maxwellian(nx, v, v0, vshift):
    """tests/helpers.py:26-34 -- the collision-test initial condition."""
    dv = v[2] - v[1]
    f = np.zeros((nx, v.size))
    f[:, :] = np.exp(-((v - vshift) ** 2.0) / 2.0 / v0)
    return f / trapz_last(f, dv)[:, None]


def epw_root(k0, wp_e=1.0, vth_e=1.0, conv=2.0):
    """diagnostics/z_function.py:26-50 -- complex EPW root of the kinetic dispersion relation.

    Z(x) = i*sqrt(pi)*wofz(x); Z'(x) = -2(1 + x Z(x)); Newton (secant) from the Bohm-Gross guess.
    """
    from scipy import optimize, special

    def zprime(x):
        return -2.0 * (1.0 + x * (1j * np.sqrt(np.pi) * special.wofz(x)))

    chi_e = np.power((wp_e / (vth_e * k0)), 2.0) / conv
    guess = np.sqrt(wp_e ** 2.0 + 3 * (k0 * vth_e) ** 2.0)
    root = optimize.newton(lambda x: 1.0 - chi_e * zprime(x), guess)
    return root * k0 * vth_e * np.sqrt(conv)


def epw_params(k0):
    """initializers.py:170-190 -- w_epw, nu_ld, xmax, v_ph for wavenumber k0."""
    r = epw_root(k0)
    return {
        "w_epw": float(np.real(r)),
        "nu_ld": float(np.imag(r)),
        "xmax": 2.0 * np.pi / k0,
        "v_ph": float(np.real(r)) / k0,
    }


def make_driver_function(x, pulses):
    """field_driver.py:24-50 -- ponderomotive driver E_d(x,t), sum over pulses of a tanh
    envelope times k*a0*sin(kx - wt)."""

    def driver(t):
        total = np.zeros(x.size)
        for p in pulses.values():
            kk, ww = p["k0"], p["w0"]
            env = 0.5 * (
                np.tanh((t - p["t_L"]) / p["t_wL"]) - np.tanh((t - p["t_R"]) / p["t_wR"])
            )
            total += env * kk * p["a0"] * np.sin(kk * x - ww * t)
        return total

    return driver


# --------------------------------------------------------------------------------------------
# A1/A2: exponential advection operators  (vlapy/core/vlasov.py)
# --------------------------------------------------------------------------------------------


def vdfdx_exponential(f, dt, kx, v):
    """vlasov.py:94-108 -- Re ifft_x( exp(((-1j*kx)*dt)*v) * fft_x f ), full complex spectrum,
    imaginary part discarded (Nyquist quirk)."""
    phase = np.exp(-1j * kx[:, None] * dt * v)
    return np.real(_sfft.ifft(phase * _sfft.fft(f, axis=0), axis=0))


def edfdv_exponential(f, e, dt, kv):
    """vlasov.py:123-138 -- Re ifft_v( exp(((-1j*kv)*dt)*e[x]) * fft_v f )."""
    phase = np.exp(-1j * kv * dt * e[:, None])
    return np.real(_sfft.ifft(phase * _sfft.fft(f, axis=1), axis=1))


def edfdv_cd2(f, e, dt, dv):
    """vlasov.py:153-163 -- f - e*df/dv*dt with np.gradient(edge_order=2) restated:
    interior (f[j+1]-f[j-1])/(2dv); edges -(3f0-4f1+f2)/(2dv) and (3fn-4fn1+fn2)/(2dv)."""
    g = np.empty_like(f)
    g[:, 1:-1] = (f[:, 2:] - f[:, :-2]) / (2.0 * dv)
    g[:, 0] = -(3.0 * f[:, 0] - 4.0 * f[:, 1] + f[:, 2]) / (2.0 * dv)
    g[:, -1] = (3.0 * f[:, -1] - 4.0 * f[:, -2] + f[:, -3]) / (2.0 * dv)
    return f - e[:, None] * g * dt


# --------------------------------------------------------------------------------------------
# N4: semi-Lagrangian advection operators  (vlapy/core/vlasov.py:27-80, 168-210)
# --------------------------------------------------------------------------------------------


def padded_grid(ax):
    """vlasov.py:27-39 -- the axis with one periodic ghost point on either side (spacing ax[2] - ax[1])."""
    ax_pad = np.zeros(ax.size + 2)
    ax_pad[1:-1] = ax
    ax_pad[0] = ax[0] - (ax[2] - ax[1])
    ax_pad[-1] = ax[-1] + (ax[2] - ax[1])
    return ax_pad


def vdfdx_sl(f, dt, x, v):
    """vlasov.py:42-80 -- backward semi-Lagrangian x advection: bicubic RectBivariateSpline of f padded with ONE
    periodic ghost row on either side, evaluated at (x - v dt, v); FITPACK clamps points outside the padded
    axis to its ends (bispeu -> fpbisp), so shifts beyond one cell are not periodic."""
    from scipy import interpolate
    xm, vm = np.meshgrid(x, v, indexing="ij")
    xm, vm = xm.flatten(), vm.flatten()
    f_pad = np.zeros((x.size + 2, v.size))
    f_pad[1:-1, :] = f
    f_pad[0, :] = f[-1, :]
    f_pad[-1, :] = f[0, :]
    interp = interpolate.RectBivariateSpline(padded_grid(x), v, f_pad)
    return interp(xm - vm * dt, vm, grid=False).reshape((x.size, v.size))


def edfdv_sl(f, e, dt, x, v):
    """vlasov.py:168-210 -- backward semi-Lagrangian v advection: the field is passed through a cubic interp1d
    evaluated at its own nodes, f is padded with one periodic ghost column on either side and the bicubic spline
    is evaluated at (x, v - e dt)."""
    from scipy import interpolate
    xm, vm = np.meshgrid(x, v, indexing="ij")
    xm, vm = xm.flatten(), vm.flatten()
    f_pad = np.zeros((x.size, v.size + 2))
    f_pad[:, 1:-1] = f
    f_pad[:, 0] = f[:, -1]
    f_pad[:, -1] = f[:, 0]
    em = interpolate.interp1d(x, e, kind="cubic")(xm)
    interp = interpolate.RectBivariateSpline(x, padded_grid(v), f_pad)
    return interp(xm, vm - em * dt, grid=False).reshape((x.size, v.size))


def nak_spline_shift(y, ax, shift):
    """What the two operators above reduce to, line by line (the algorithm of csrc/spline.h): the bicubic
    interpolating spline restricted to a grid line of the OTHER axis is the 1-D not-a-knot cubic spline of that line
    (tensor-product interpolation), so every line y (last axis, already padded with its two ghost points, uniform
    spacing h = ax[2] - ax[1]) is interpolated on its own and evaluated at ax[1:-1] - shift, clamped to
    [ax[0], ax[-1]].  Not-a-knot on a uniform grid: M_0 - 2 M_1 + M_2 = 0 turns the first interior equation into
    6 M_1 = r_1 (likewise at the other end), the rest is the (1, 4, 1) system in the second derivatives M.
    y: (..., n + 2); shift: broadcastable to (..., n)."""
    n2 = y.shape[-1]
    h = ax[2] - ax[1]
    r = 6.0 * (y[..., 2:] - 2.0 * y[..., 1:-1] + y[..., :-2]) / (h * h)          # r_1 .. r_{n2-2}
    M = np.zeros_like(y)
    M[..., 1] = r[..., 0] / 6.0
    M[..., n2 - 2] = r[..., -1] / 6.0
    m = n2 - 4                                                                    # unknowns M_2 .. M_{n2-3}
    if m > 0:
        rhs = r[..., 1:-1].copy()
        rhs[..., 0] -= M[..., 1]
        rhs[..., -1] -= M[..., n2 - 2]
        cp = np.zeros(m)
        dp = np.zeros(rhs.shape)
        cp[0] = 1.0 / 4.0
        dp[..., 0] = rhs[..., 0] / 4.0
        for k in range(1, m):
            den = 4.0 - cp[k - 1]
            cp[k] = 1.0 / den
            dp[..., k] = (rhs[..., k] - dp[..., k - 1]) / den
        M[..., n2 - 3] = dp[..., m - 1]
        for k in range(m - 2, -1, -1):
            M[..., k + 2] = dp[..., k] - cp[k] * M[..., k + 3]
    M[..., 0] = 2.0 * M[..., 1] - M[..., 2]
    M[..., n2 - 1] = 2.0 * M[..., n2 - 2] - M[..., n2 - 3]
    xq = np.clip(ax[1:-1] - shift, ax[0], ax[-1])
    c = np.clip(np.floor((xq - ax[0]) / h).astype(np.int64), 0, n2 - 2)
    s = (xq - (ax[0] + c * h)) / h
    yc, yn = np.take_along_axis(y, c, -1), np.take_along_axis(y, c + 1, -1)
    Mc, Mn = np.take_along_axis(M, c, -1), np.take_along_axis(M, c + 1, -1)
    u = 1.0 - s
    return u * yc + s * yn + (h * h / 6.0) * ((u * u * u - u) * Mc + (s * s * s - s) * Mn)


def vdfdx_sl_lines(f, dt, x, v):
    """vdfdx_sl through nak_spline_shift (one spline per v column along the padded x axis)."""
    fp = np.concatenate([f[-1:, :], f, f[:1, :]], axis=0).T                      # (nv, nx + 2)
    shift = np.broadcast_to((v * dt)[:, None], (v.size, x.size))
    return nak_spline_shift(np.ascontiguousarray(fp), padded_grid(x), shift).T


def edfdv_sl_lines(f, e, dt, x, v):
    """edfdv_sl through nak_spline_shift (one spline per x row along the padded v axis)."""
    fp = np.concatenate([f[:, -1:], f, f[:, :1]], axis=1)                        # (nx, nv + 2)
    shift = np.broadcast_to((e * dt)[:, None], (x.size, v.size))
    return nak_spline_shift(fp, padded_grid(v), shift)


# --------------------------------------------------------------------------------------------
# A3/A4: charge density and spectral Poisson  (vlapy/core/field.py)
# --------------------------------------------------------------------------------------------


def compute_charges(f, dv):
    """field.py:27-36."""
    return trapz_last(f, dv)


def solve_for_field(charge_density, one_over_kx):
    """field.py:39-63 -- E = Re ifft( 1j*one_over_kx * fft(1 - n) )."""
    net = 1.0 - charge_density
    return np.real(_sfft.ifft(1j * one_over_kx * _sfft.fft(net)))


def field_solve(driver_field, f, dv, one_over_kx):
    """field.py:66-88 -- total field = driver + self-consistent field."""
    return driver_field + solve_for_field(compute_charges(f, dv), one_over_kx)


# --------------------------------------------------------------------------------------------
# A5-A8: implicit Fokker-Planck step  (vlapy/core/collisions.py, step.py:70-113)
# --------------------------------------------------------------------------------------------


def lb_diagonals(f, v, nu, dt, dv):
    """collisions.py:44-81 -- Lenard-Bernstein sub/main/super diagonals (moment not
    density-normalised; sub uses v[:-1], super uses v[1:])."""
    nx, nv = f.shape
    v0t_sq = trapz_last(f * v[None, :] ** 2.0, dv)
    a = nu * dt * np.ones((nx, nv - 1)) * (-v0t_sq[:, None] / dv ** 2.0 + v[None, :-1] / 2 / dv)
    b = 1.0 + nu * dt * np.ones((nx, nv)) * (2 * v0t_sq[:, None] / dv ** 2.0)
    c = nu * dt * np.ones((nx, nv - 1)) * (-v0t_sq[:, None] / dv ** 2.0 - v[None, 1:] / 2 / dv)
    return a, b, c


def dg_diagonals(f, v, nu, dt, dv):
    """collisions.py:104-158 -- Dougherty diagonals, drift (v - vbar) and thermal spread about vbar."""
    nx, nv = f.shape
    vbar = trapz_last(f * v[None, :], dv)
    v0t_sq = trapz_last(f * (v[None, :] - vbar[:, None]) ** 2.0, dv)
    a = (
        nu * dt * np.ones((nx, nv - 1))
        * (-v0t_sq[:, None] / dv ** 2.0 + (v[None, :-1] - vbar[:, None]) / 2.0 / dv)
    )
    b = 1.0 + nu * dt * np.ones((nx, nv)) * (2.0 * v0t_sq[:, None] / dv ** 2.0)
    c = (
        nu * dt * np.ones((nx, nv - 1))
        * (-v0t_sq[:, None] / dv ** 2.0 - (v[None, 1:] - vbar[:, None]) / 2.0 / dv)
    )
    return a, b, c


def thomas_batched(a, b, c, d):
    """collisions.py:232-263 -- Thomas algorithm vectorised over x, no pivoting, new array out."""
    nv = b.shape[1]
    ac, bc, cc, dc = a.copy(), b.copy(), c.copy(), d.copy()
    for it in range(1, nv):
        mc = ac[:, it - 1] / bc[:, it - 1]
        bc[:, it] = bc[:, it] - mc * cc[:, it - 1]
        dc[:, it] = dc[:, it] - mc * dc[:, it - 1]
    xc = bc
    xc[:, -1] = dc[:, -1] / bc[:, -1]
    for il in range(nv - 2, -1, -1):
        xc[:, il] = (dc[:, il] - cc[:, il] * xc[:, il + 1]) / bc[:, il]
    return xc


def collision_step(f, v, nu, dt, dv, operator="lb"):
    """step.py:70-113 -- identity for nu == 0; else diagonals + tridiagonal solve."""
    if nu == 0.0:
        return f
    if nu < 0.0:
        raise NotImplementedError
    if operator == "lb":
        a, b, c = lb_diagonals(f, v, nu, dt, dv)
    elif operator == "dg":
        a, b, c = dg_diagonals(f, v, nu, dt, dv)
    else:
        raise NotImplementedError(operator)
    return thomas_batched(a, b, c, f)


# --------------------------------------------------------------------------------------------
# A10: splitting schedules  (vlapy/core/vlasov_poisson.py)
# --------------------------------------------------------------------------------------------

PEFRL_XSI = 0.1786178958448091
PEFRL_LAMBDA = -0.2123418310626054
PEFRL_CHI = -0.6626458266981849e-1

H6 = dict(
    a1=0.168735950563437422448196, a2=0.377851589220928303880766, a3=-0.093175079568731452657924,
    b1=0.049086460976116245491441, b2=0.264177609888976700200146, b3=0.186735929134907054308413,
    c1=-0.000069728715055305084099, c2=-0.000625704827430047189169, c3=-0.002213085124045325561636,
    d2=-2.916600457689847816445691e-6, d3=3.048480261700038788680723e-5,
    e3=4.985549387875068121593988e-7,
)


def schedule(name, dt):
    """The ordered sub-steps of one Vlasov-Poisson step as a list of
    ("v", dt_v) | ("x", dt_x, driver_time_increments) tuples; the driver is evaluated at
    ((t + inc[0]) + inc[1]) + ... summed left to right exactly as the reference writes it.

    leapfrog: vlasov_poisson.py:53-56 (v-half, x-full, field at t+dt, v-half);
    pefrl: vlasov_poisson.py:100-148 (cumulative driver times);
    h-sixth: vlasov_poisson.py:199-228 (driver times t + a_i*dt, NOT cumulative).
    """
    if name == "leapfrog":
        return [("v", 0.5 * dt), ("x", dt, (dt,)), ("v", 0.5 * dt)]
    if name == "pefrl":
        xsi, lambd, chi = PEFRL_XSI, PEFRL_LAMBDA, PEFRL_CHI
        dt1 = xsi * dt
        dt2 = chi * dt
        dt3 = (1.0 - 2.0 * (chi + xsi)) * dt
        dt4, dt5 = dt2, dt1
        vdt1 = 0.5 * (1.0 - 2.0 * lambd) * dt
        vdt2 = lambd * dt
        vdt3, vdt4 = vdt2, vdt1
        return [
            ("x", dt1, (dt1,)), ("v", vdt1),
            ("x", dt2, (dt1, dt2)), ("v", vdt2),
            ("x", dt3, (dt1, dt2, dt3)), ("v", vdt3),
            ("x", dt4, (dt1, dt2, dt3, dt4)), ("v", vdt4),
            ("x", dt5, (dt1, dt2, dt3, dt4, dt5)),
        ]
    if name == "h-sixth":
        c = H6
        D1 = c["b1"] + 2.0 * c["c1"] * dt ** 2.0
        D2 = c["b2"] + 2.0 * c["c2"] * dt ** 2.0 + 4.0 * c["d2"] * dt ** 4.0
        D3 = c["b3"] + 2.0 * c["c3"] * dt ** 2.0 + 4.0 * c["d3"] * dt ** 4.0 - 8.0 * c["e3"] * dt ** 6.0
        a1, a2, a3 = c["a1"], c["a2"], c["a3"]
        return [
            ("v", D1 * dt), ("x", a1 * dt, (a1 * dt,)),
            ("v", D2 * dt), ("x", a2 * dt, (a2 * dt,)),
            ("v", D3 * dt), ("x", a3 * dt, (a3 * dt,)),
            ("v", D3 * dt), ("x", a2 * dt, (a2 * dt,)),
            ("v", D2 * dt), ("x", a1 * dt, (a1 * dt,)),
            ("v", D1 * dt),
        ]
    raise NotImplementedError("df/dt : <" + name + "> has not yet been implemented")


def vp_step(e, f, t, *, integrator, dt, kx, kv, v, dv, one_over_kx, driver_function,
            edfdv="exponential", vdfdx="exponential", x=None):
    """One Vlasov-Poisson step (e, f, t) -> (e, f).  Every x sub-step is followed by a field
    solve at the scheduled driver time (vlasov_poisson.py:54-55, 115-116, 205-206)."""
    for sub in schedule(integrator, dt):
        if sub[0] == "v":
            if edfdv == "exponential":
                f = edfdv_exponential(f, e, sub[1], kv)
            elif edfdv == "cd2":
                f = edfdv_cd2(f, e, sub[1], dv)
            elif edfdv == "sl":
                f = edfdv_sl(f, e, sub[1], x, v)
            else:
                raise NotImplementedError(edfdv)
        else:
            if vdfdx == "exponential":
                f = vdfdx_exponential(f, sub[1], kx, v)
            elif vdfdx == "sl":
                f = vdfdx_sl(f, sub[1], x, v)
            else:
                raise NotImplementedError(vdfdx)
            td = t
            for inc in sub[2]:
                td = td + inc
            e = field_solve(driver_function(td), f, dv, one_over_kx)
    return e, f


# --------------------------------------------------------------------------------------------
# A9: per-step stored quantities  (vlapy/core/step.py:116-283)
# --------------------------------------------------------------------------------------------

FIELD_KEYS = ("e", "driver", "n", "j", "T", "q", "fv4", "vN")
SERIES_KEYS = ("mean_n", "mean_j", "mean_T", "mean_e2", "mean_de2", "mean_f2", "mean_flogf")


def field_moments(f, v, dv):
    """step.py:164-171 -- trapz_v(f * v**p) for p = 0..5, stacked (6, nx)."""
    out = [trapz_last(f, dv), trapz_last(f * v, dv)]
    for p in (2, 3, 4, 5):
        out.append(trapz_last(f * v ** p, dv))
    return np.stack(out)


def series_moments(f, e, de, mom, dv):
    """step.py:189-226 -- x-means of n, j, T, e^2, de^2, int f^2 dv, int f ln f dv."""
    with np.errstate(invalid="ignore", divide="ignore"):
        flogf = np.mean(trapz_last(f * np.log(f), dv), axis=0)
    return np.array([
        np.mean(mom[0]), np.mean(mom[1]), np.mean(mom[2]),
        np.mean(e ** 2.0), np.mean(de ** 2.0),
        np.mean(trapz_last(np.real(f) ** 2.0, dv)), flogf,
    ])


def stored_f_modes(f, nmodes=2):
    """step.py:130-135 -- the first ``nmodes`` x-Fourier modes of f per v (stored as complex64,
    outer_loop.py:172-184)."""
    return _sfft.fft(f, axis=0)[:nmodes]


def timestep(e, f, t, de, *, nu, fp_operator, **vp_kw):
    """step.py:302-326 -- vp_step, fp_step, then the stored quantities of the new state."""
    e, f = vp_step(e, f, t, **vp_kw)
    f = collision_step(f, vp_kw["v"], nu, vp_kw["dt"], vp_kw["dv"], fp_operator)
    mom = field_moments(f, vp_kw["v"], vp_kw["dv"])
    ser = series_moments(f, e, de, mom, vp_kw["dv"])
    return e, f, mom, ser


# --------------------------------------------------------------------------------------------
# configuration builders shared by tests and bench (SURVEY 8d synthetic inputs)
# --------------------------------------------------------------------------------------------


def steps_in_loop_like_manager(nx, nv, nt, max_gb=1, nmodes=2):
    """manager.py:61-83 -- the inner-loop length heuristic (runs n_loops*steps >= nt steps)."""
    mem_f_store = 2 * nmodes * nv
    mem_field_store = nx * 8
    steps = int(1e9 * max_gb / (6 * (mem_f_store + mem_field_store) * 8))
    if steps > nt:
        steps = int(nt / 1.25)
    n_loops = nt // steps + 1
    return steps, n_loops


def make_config(nx, nv, k0, *, tmax, nt, a0, t_R, vmax=6.4, log_nu_over_nu_ld=None,
                w_epw=None, nu_ld=None):
    """Grids, dt, driver and initial state exactly as outer_loop.py:98-144 builds them."""
    if w_epw is None:
        p = epw_params(k0)
        w_epw, nu_ld = p["w_epw"], p["nu_ld"]
    xmax = 2.0 * np.pi / k0
    dx, x, kx, one_over_kx = spatial_grid(0.0, xmax, nx)
    dv, v, kv = velocity_grid(vmax, nv)
    t_dummy = np.linspace(0, tmax, nt)
    dt = t_dummy[1] - t_dummy[0]
    pulses = {"first pulse": {"start_time": 0, "t_L": 6, "t_wL": 2.5, "t_R": t_R, "t_wR": 2.5,
                              "w0": w_epw, "a0": a0, "k0": k0}}
    nu = 0.0 if log_nu_over_nu_ld is None else abs(nu_ld) * 10 ** log_nu_over_nu_ld
    return dict(nx=nx, nv=nv, k0=k0, x=x, dx=dx, kx=kx, one_over_kx=one_over_kx, v=v, dv=dv,
                kv=kv, dt=dt, nu=nu, pulses=pulses, w_epw=w_epw, nu_ld=nu_ld,
                driver_function=make_driver_function(x, pulses),
                f0=maxwellian(nx, nv, vmax), e0=np.zeros(nx), tmax=tmax, nt=nt, vmax=vmax)


# Dispersion roots quoted in SURVEY.md Appendix B (so configs can be built without scipy.special).
EPW_KNOWN = {
    0.3: (1.1598464805919155, -0.012620368421117013),
    0.35: (1.220953506161683, -0.03431805085829906),
}


def landau_config(nx=32, nv=512, k0=0.3):
    """C1: tests/test_landau_damping.py:52-79 with the default params of initializers.py:118-149."""
    w, g = EPW_KNOWN.get(k0, (None, None))
    return make_config(nx, nv, k0, tmax=80, nt=500, a0=1e-7, t_R=20, w_epw=w, nu_ld=g)


def nlepw_config(nx=256, nv=2048, k0=0.35, log_nu=-4):
    """C2/C3/C5: run_nlepw.py:29-57 with k0 fixed (the script draws it at random)."""
    w, g = EPW_KNOWN.get(k0, (None, None))
    return make_config(nx, nv, k0, tmax=1000, nt=4000, a0=4e-2, t_R=25,
                       log_nu_over_nu_ld=log_nu, w_epw=w, nu_ld=g)


def run_steps(cfg, nsteps, integrator="leapfrog", fp_operator="lb", edfdv="exponential",
              collect=False, vdfdx="exponential"):
    """Drive ``nsteps`` full timesteps from the config's initial state the way
    outer_loop.py:265-272 + step.py:302-326 do (time of step i is i*dt, driver row is the
    driver at that time)."""
    e, f = cfg["e0"].copy(), cfg["f0"].copy()
    kw = dict(integrator=integrator, dt=cfg["dt"], kx=cfg["kx"], kv=cfg["kv"], v=cfg["v"],
              dv=cfg["dv"], one_over_kx=cfg["one_over_kx"], driver_function=cfg["driver_function"],
              edfdv=edfdv, vdfdx=vdfdx, x=cfg["x"])
    hist = {"e": [], "mom": [], "series": []}
    for i in range(nsteps):
        t = cfg["dt"] * i
        de = cfg["driver_function"](t)
        e, f, mom, ser = timestep(e, f, t, de, nu=cfg["nu"], fp_operator=fp_operator, **kw)
        if collect:
            hist["e"].append(e.copy()); hist["mom"].append(mom); hist["series"].append(ser)
    if collect:
        return e, f, {k: np.array(val) for k, val in hist.items()}
    return e, f


def damping_rate(e_hist, tax):
    """diagnostics/low_level_helpers.py:69-87 -- mean d/dt log|E_k1| over the last 75 % of steps."""
    t_ind = tax.size // 4
    ek = np.fft.fft(e_hist, axis=1)
    ek_mag = np.abs(ek[:, 1])[t_ind:]
    return float(np.mean(np.gradient(np.log(ek_mag), tax[2] - tax[1])))
