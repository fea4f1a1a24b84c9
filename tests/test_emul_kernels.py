"""CPU check of the kernel *programs* (vlapy_b200/csrc/advect.h, rowops.h): the same sources that
nvcc compiles for sm_100a are compiled with g++ and run thread-by-thread (tests/emul/), then
compared with the golden fixtures / oracle at the parity tolerance of BASELINE.json (1e-12).
This validates index arithmetic and numerics where no GPU exists; the -m gpu tests repeat the
comparison on the real kernels."""
import os
import sys

import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import vpfp_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emul"))
import emul_lib as E  # noqa: E402

TOL = 1e-12


@pytest.mark.parametrize("name", ["ops_small", "ops_c1", "ops_white"])
@pytest.mark.parametrize("max_single", [8192, 8])
def test_advection_programs(name, max_single):
    g = golden(name)
    f, e, dt = g["f"], g["e"], float(g["dt"])
    assert rel_err(E.vdfdx_exp(f, g["kx"], g["v"], dt, max_single=max_single), g["vdfdx"]) < TOL
    assert rel_err(E.vdfdx_exp(f, g["kx"], g["v"], -0.066 * dt, max_single=max_single), g["vdfdx_neg"]) < TOL
    assert rel_err(E.edfdv_exp(f, e, g["kv"], 0.5 * dt, max_single=max_single), g["edfdv"]) < TOL
    assert rel_err(E.edfdv_exp(f, e, g["kv"], -0.21 * dt, max_single=max_single), g["edfdv_neg"]) < TOL


@pytest.mark.parametrize("shape", [(2, 4), (4, 8), (6, 16), (3, 64), (64, 4), (7, 128), (1, 32), (5, 256)])
@pytest.mark.parametrize("max_single", [8192, 4, 16])
def test_edfdv_shapes(shape, max_single):
    """rows: any count (odd counts leave the last packed channel empty); nv: powers of two."""
    nx, nv = shape
    rng = np.random.default_rng(nx * 1000 + nv)
    f = rng.standard_normal((nx, nv))
    e = rng.standard_normal(nx)
    dv, v, kv = O.velocity_grid(6.4, nv)
    assert rel_err(E.edfdv_exp(f, e, kv, 0.37, max_single=max_single), O.edfdv_exponential(f, e, 0.37, kv)) < TOL


@pytest.mark.parametrize("shape", [(2, 4), (4, 2), (8, 6), (16, 18), (64, 2), (128, 6), (32, 34), (256, 4)])
@pytest.mark.parametrize("max_single", [8192, 4, 16])
def test_vdfdx_shapes(shape, max_single):
    """nx: powers of two; column count: any even number (ragged last tile)."""
    nx, nv = shape
    rng = np.random.default_rng(nx * 1000 + nv)
    f = rng.standard_normal((nx, nv))
    v = np.linspace(-6.0, 6.0, nv)
    dx, x, kx, ook = O.spatial_grid(0.0, 20.0, nx)
    assert rel_err(E.vdfdx_exp(f, kx, v, 0.37, max_single=max_single), O.vdfdx_exponential(f, 0.37, kx, v)) < TOL


def test_vdfdx_batched_sims_have_their_own_kx():
    rng = np.random.default_rng(5)
    batch, nx, nv = 3, 16, 8
    f = rng.standard_normal((batch, nx, nv))
    dv, v, kv = O.velocity_grid(6.4, nv)
    kxs = np.stack([O.spatial_grid(0.0, 2 * np.pi / k0, nx)[2] for k0 in (0.25, 0.3, 0.45)])
    out = E.vdfdx_exp(f, kxs, v, 0.2, batch=batch)
    for b in range(batch):
        assert rel_err(out[b], O.vdfdx_exponential(f[b], 0.2, kxs[b], v)) < TOL


@pytest.mark.parametrize("max_single", [8192, 8])
def test_poisson_program(max_single):
    for name in ("ops_small", "ops_c1", "ops_white"):
        g = golden(name)
        e = E.poisson(g["charges"], g["one_over_kx"], g["drv"], max_single=max_single)[0]
        assert rel_err(e, g["efield"]) < TOL
    # several rows (odd count) with different one_over_kx rows
    rng = np.random.default_rng(3)
    n = 1.0 + 0.1 * rng.standard_normal((5, 64))
    ooks = np.stack([O.spatial_grid(0.0, 2 * np.pi / k0, 64)[3] for k0 in (0.25, 0.3, 0.35, 0.4, 0.45)])
    drv = 0.01 * rng.standard_normal((5, 64))
    e = E.poisson(n, ooks, drv, max_single=max_single)
    for i in range(5):
        assert rel_err(e[i], drv[i] + O.solve_for_field(n[i], ooks[i])) < TOL


def test_cd2_moments_modes_series_driver():
    for name in ("ops_small", "ops_c1"):
        g = golden(name)
        dv, dt = float(g["dv"]), float(g["dt"])
        assert rel_err(E.edfdv_cd2(g["f"], g["e"], 0.5 * dt, dv), g["cd2"]) < TOL
        mom = E.moments(g["fpos"], g["v"], dv)
        assert rel_err(mom[:6], g["moments"]) < TOL
        assert rel_err(E.moments(g["f"], g["v"], dv, nmom=1)[0], g["charges"]) < TOL
        ser = E.series(mom, g["e"], g["drv"])
        np.testing.assert_allclose(ser, g["series"], rtol=1e-12)
        assert rel_err(E.xmodes(g["fpos"]), g["modes"]) < 1e-12
    cfg = O.landau_config()
    for t in (0.0, 3.3, 7.7, 21.0):
        np.testing.assert_allclose(E.driver(cfg["x"], t, cfg["pulses"]), cfg["driver_function"](t),
                                   rtol=1e-13, atol=1e-22)


def test_moments_nan_semantics_of_flogf():
    g = golden("ops_small")
    mom = E.moments(g["f"], g["v"], float(g["dv"]))     # f has negative cells -> log() is NaN
    ref = O.series_moments(g["f"], g["e"], g["drv"], O.field_moments(g["f"], g["v"], float(g["dv"])), float(g["dv"]))
    assert np.isnan(ref[6]) and np.isnan(mom[7]).any()


@pytest.mark.parametrize("op", ["lb", "dg"])
def test_fp_program_vs_reference(op):
    for name in ("ops_small", "ops_c1", "ops_white"):
        g = golden(name)
        out, mom = E.fp_step(g["fpos"], g["v"], float(g["nu"]), float(g["dt"]), float(g["dv"]), op, want_moments=True)
        assert rel_err(out, g[op + "_solve"]) < TOL
        assert rel_err(mom[:6], O.field_moments(g[op + "_solve"], g["v"], float(g["dv"]))) < TOL


def fp_longdouble(f, v, nu, dt, dv, op):
    """The reference algorithm (moments, diagonals, Thomas) in extended precision: the yardstick
    that separates algorithmic error from the conditioning of a stiff system."""
    L = np.longdouble
    f, v, nu, dt, dv = f.astype(L), v.astype(L), L(nu), L(dt), L(dv)
    tr = lambda y: (dv * (y[..., 1:] + y[..., :-1]) / 2).sum(-1)  # noqa: E731
    if op == "lb":
        vbar, T = np.zeros(f.shape[0], L), tr(f * v ** 2)
    else:
        vbar = tr(f * v)
        T = tr(f * (v[None, :] - vbar[:, None]) ** 2)
    nx, nv = f.shape
    a = nu * dt * (-T[:, None] / dv ** 2 + (v[None, :-1] - vbar[:, None]) / 2 / dv)
    b = 1 + nu * dt * np.ones((nx, nv), L) * (2 * T[:, None] / dv ** 2)
    c = nu * dt * (-T[:, None] / dv ** 2 - (v[None, 1:] - vbar[:, None]) / 2 / dv)
    return O.thomas_batched(a, b, c, f).astype(np.float64)


@pytest.mark.parametrize("op", ["lb", "dg"])
@pytest.mark.parametrize("nu", [1e-2, 1.0, 30.0])
@pytest.mark.parametrize("nv,m", [(8, 4), (24, 4), (100, 7), (1024, 0), (2048, 16), (4096, 0)])
def test_fp_program_sizes_and_stiffness(op, nu, nv, m):
    """Chunked elimination + cyclic reduction vs the reference's Thomas sweep.  For stiff systems
    (nu*dt*T/dv^2 >> 1, condition number ~1e6) the reference itself is only good to ~1e-10, so the
    bar is: within 1e-12 of the reference, or no further from the extended-precision solution
    than a small multiple of the reference's own distance."""
    dv, v, kv = O.velocity_grid(6.0, nv)
    f = O.shifted_maxwellian(3, v, 1.0, 0.7) * np.array([1.0, 0.7, 1.3])[:, None]
    ref = O.collision_step(f, v, nu, 0.1, dv, op)
    out = E.fp_step(f, v, nu, 0.1, dv, op, m=m)
    if rel_err(out, ref) < TOL:
        return
    truth = fp_longdouble(f, v, nu, 0.1, dv, op)
    assert rel_err(out, truth) < 4 * rel_err(ref, truth) + TOL


@pytest.mark.parametrize("op", ["lb", "dg"])
def test_two_stage_collision_programs_vs_reference(op):
    """csrc/tridiag.h: DiagProg against the diagonals the REFERENCE's get_batched_array_maker returned (golden
    ops_small lb_a/lb_b/lb_c, dg_*), TridiagProg on those diagonals against the reference's solve"""
    g = golden("ops_small")
    nu, dt, dv = float(g["nu"]), float(g["dt"]), float(g["dv"])
    a, b, c = E.fp_diagonals(g["fpos"], g["v"], nu, dt, dv, op)
    for got, name in ((a, "_a"), (b, "_b"), (c, "_c")):
        ref = g[op + name]
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref)), name
    assert rel_err(E.tridiag_solve(g[op + "_a"], g[op + "_b"], g[op + "_c"], g["fpos"]), g[op + "_solve"]) < TOL
    assert rel_err(E.tridiag_solve(a, b, c, g["fpos"]), g[op + "_solve"]) < TOL


@pytest.mark.parametrize("nv,m", [(8, 4), (24, 4), (100, 7), (1000, 0), (1024, 0), (4096, 0), (16384, 0)])
def test_tridiag_program_general_diagonals(nv, m):
    """general (non-constant, non-symmetric) diagonally dominant diagonals, ragged sizes, against the reference's
    Thomas sweep (oracle thomas_batched = vlapy/core/collisions.py:232-263)"""
    rng = np.random.default_rng(nv)
    rows = 3
    a = rng.uniform(-1, 1, (rows, nv - 1))
    c = rng.uniform(-1, 1, (rows, nv - 1))
    b = 2.5 + rng.uniform(0, 1, (rows, nv))
    d = rng.standard_normal((rows, nv))
    ref = O.thomas_batched(a, b, c, d)
    out = E.tridiag_solve(a, b, c, d, m=m)
    assert rel_err(out, ref) < TOL
    # the system is solved: residual of the tridiagonal product
    res = b * out
    res[:, 1:] += a * out[:, :-1]
    res[:, :-1] += c * out[:, 1:]
    assert rel_err(res, d) < 1e-13


def test_fp_collision_unit_cases_16_steps():
    g = golden("collisions_unit")
    v, dv, nu, dt = g["v"], float(g["dv"]), float(g["nu"]), float(g["dt"])
    for vshift in (0.0, 0.5, 1.5):
        for op in ("lb", "dg"):
            f = g["f_%g" % vshift].copy()
            for _ in range(16):
                f = E.fp_step(f, v, nu, dt, dv, op)
            assert rel_err(f, g["out_%s_%g" % (op, vshift)]) < TOL


@pytest.mark.parametrize("radix", [4, 8, 16])
def test_register_butterflies(radix):
    rng = np.random.default_rng(radix)
    x = rng.standard_normal(radix) + 1j * rng.standard_normal(radix)
    np.testing.assert_allclose(E.butterfly(x, radix, -1), np.fft.fft(x), rtol=0, atol=1e-14)
    np.testing.assert_allclose(E.butterfly(x, radix, +1), np.fft.ifft(x) * radix, rtol=0, atol=1e-14)


@pytest.mark.parametrize("nv", [4096, 8192, 16384])
def test_rowfft_program(nv, monkeypatch):
    """single-pass row kernel (rowfft.cuh: real row as one complex sequence of nv/2 points, pairs
    (k, M-k) un-mixed in registers) against the oracle: white noise (every bin, Nyquist included)
    and a smooth Maxwellian-like row, positive and negative sub-steps, odd row count.  Three thread
    orders: the kernel has no barrier after its pointwise phase and after the last phase of a row (the
    emulation runs such phases back to back per thread), a race there would make the result order dependent."""
    rng = np.random.default_rng(nv)
    dv, v, kv = O.velocity_grid(6.4, nv)
    f = rng.standard_normal((3, nv))
    f[1] = np.exp(-v ** 2 / 2) * (1 + 1e-3 * rng.standard_normal(nv))
    e = np.array([0.05, -0.7, 1.3])
    outs = []
    for order in ("0", "1", "2"):
        monkeypatch.setenv("VPFP_EMUL_ORDER", order)
        for dt in (0.37, -0.066):
            out = E.edfdv_rowfft(f, e, kv, dt)
            assert rel_err(out, O.edfdv_exponential(f, e, dt, kv)) < TOL
            outs.append(out)
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[0], outs[4])


# ---- Fokker-Planck __global__ kernels run thread by thread on the host (tests/emul/simt.h)
@pytest.mark.parametrize("which", [0, 1], ids=["fp_fast", "fp_reg"])
@pytest.mark.parametrize("op", ["lb", "dg"])
@pytest.mark.parametrize("nv", [4096, 16384])
def test_fp_global_kernels_on_host(which, op, nv):
    """fp_fast.cuh (row in shared memory) and fp_reg.cuh (row in registers, determinant recurrences,
    next row prefetched) against the reference's Thomas sweep and row moments: weak (the C5 value) and
    strong collisions, three rows over two CTAs so that the row loop and the prefetch are exercised."""
    dv, v, kv = O.velocity_grid(6.4, nv)
    rng = np.random.default_rng(nv)
    f = O.shifted_maxwellian(3, v, 1.0, 0.7) * np.array([1.0, 0.7, 1.3])[:, None]
    f = f * (1 + 1e-3 * rng.standard_normal(f.shape))
    for nu in (3e-6, 1e-2):
        ref = O.collision_step(f, v, nu, 0.25, dv, op)
        out, mom = E.fp_simt(which, f, v, nu, 0.25, dv, op)
        err = rel_err(out, ref)
        if err >= TOL:      # stiff: the reference itself is only good to ~1e-10 (see the test above)
            truth = fp_longdouble(f, v, nu, 0.25, dv, op)
            assert rel_err(out, truth) < 4 * rel_err(ref, truth) + TOL
        mref = np.asarray(O.field_moments(out, v, dv))
        for p in range(6):
            assert rel_err(mom[p], mref[p]) < 1e-13
        assert rel_err(mom[6], O.trapz_last(out ** 2, dv)) < 1e-13
        assert rel_err(mom[7], O.trapz_last(out * np.log(out), dv)) < 1e-13


def test_fp_reg_stiffness_and_nan_semantics():
    """fp_reg.cuh: very stiff systems (unit-diagonal scaling keeps the determinants in [2^-31, 1]) and
    numpy's NaN for f ln f of a non-positive cell (vlapy/core/step.py:222-224)."""
    nv = 4096
    dv, v, kv = O.velocity_grid(6.0, nv)
    f = O.shifted_maxwellian(2, v, 1.0, 0.3)
    for nu in (1.0, 30.0, 1e4):
        for op in ("lb", "dg"):
            ref = O.collision_step(f, v, nu, 0.1, dv, op)
            out, mom = E.fp_simt(1, f, v, nu, 0.1, dv, op)
            assert np.isfinite(out).all()
            if rel_err(out, ref) >= TOL:
                # the diagonal b = 1 + 2 nu dt T / dv^2 carries the "1" of (I - dt C) with an absolute
                # rounding error eps * b: every fp64 evaluation of this system is uncertain at that level
                bd = 1 + 2 * nu * 0.1 / dv ** 2
                truth = fp_longdouble(f, v, nu, 0.1, dv, op)
                assert rel_err(out, truth) < 4 * rel_err(ref, truth) + TOL + 0.25 * np.finfo(float).eps * bd
    g = f.copy()
    g[1, 100] = -1e-3
    out, mom = E.fp_simt(1, g, v, 0.0, 0.1, dv, "lb")
    assert rel_err(out, g) < 1e-15            # nu = 0: identity matrix
    assert np.isfinite(mom[7, 0]) and np.isnan(mom[7, 1])


def test_semi_lagrangian_programs_vs_reference():
    """csrc/spline.h (column sweep, row right-hand sides + tridiag.h solve with broadcast diagonals, evaluation at
    the clamped feet of the characteristics) against outputs of the REFERENCE's sl operators (golden sl_ops), and at
    ragged sizes against the oracle"""
    g = golden("sl_ops")
    for tag in ("small", "c1"):
        f, e, x, v = g[tag + "_f"], g[tag + "_e"], g[tag + "_x"], g[tag + "_v"]
        for name, dt in zip("abc", g[tag + "_dts"]):
            assert rel_err(E.vdfdx_sl(f, x, v, dt), g["%s_vdfdx_%s" % (tag, name)]) < TOL
            assert rel_err(E.edfdv_sl(f, e, v, dt), g["%s_edfdv_%s" % (tag, name)]) < TOL
    rng = np.random.default_rng(3)
    for nx, nv in ((5, 10), (7, 33), (100, 130)):
        x, v = np.linspace(0.3, 20.0, nx), np.linspace(-6.0, 6.0, nv)
        f = np.exp(-v ** 2 / 2)[None, :] * (1 + 0.1 * np.sin(0.3 * x))[:, None] + 1e-3 * rng.standard_normal((nx, nv))
        e = 0.7 * np.cos(0.3 * x)
        for dt in (0.2, -1.3):
            assert rel_err(E.vdfdx_sl(f, x, v, dt), O.vdfdx_sl(f, dt, x, v)) < TOL
            assert rel_err(E.edfdv_sl(f, e, v, dt), O.edfdv_sl(f, e, dt, x, v)) < TOL


@pytest.mark.parametrize("n", [256, 512, 1024, 2048])
def test_midfft_programs(n, monkeypatch):
    """midfft.cuh (single-pass mid-size transforms, 16 values per thread, L = R1 x R2 x 8) against the oracle: e df/dv
    with odd and even row counts, v df/dx with ragged column tiles and per-simulation wavenumbers, both signs of dt,
    in three thread orders (a race inside a phase would make the result order dependent)"""
    rng = np.random.default_rng(n)
    dv, v, kv = O.velocity_grid(6.4, n)
    f = rng.standard_normal((5, n))
    f[1] = np.exp(-v ** 2 / 2) * (1 + 1e-3 * rng.standard_normal(n))
    e = np.array([0.05, -0.7, 1.3, 0.0, 2.5])
    outs = []
    for order in ("0", "1", "2"):
        monkeypatch.setenv("VPFP_EMUL_ORDER", order)
        for dt in (0.37, -0.066):
            out = E.midfft_rows(f, e, kv, dt)
            assert rel_err(out, O.edfdv_exponential(f, e, dt, kv)) < TOL
            outs.append(out)
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[0], outs[4])
    monkeypatch.setenv("VPFP_EMUL_ORDER", "0")
    assert rel_err(E.midfft_rows(f[:4], e[:4], kv, 0.2), O.edfdv_exponential(f[:4], e[:4], 0.2, kv)) < TOL
    fr, er = rng.standard_normal((37, n)), rng.standard_normal(37)      # several tiles per CTA: the prefetched path
    assert rel_err(E.midfft_rows(fr, er, kv, -0.11), O.edfdv_exponential(fr, er, -0.11, kv)) < TOL
    ncols = 26
    vv = np.linspace(-6.4, 6.4, ncols)
    k0s = (0.3, 0.41)
    kx = np.stack([O.spatial_grid(0.0, 2 * np.pi / k, n)[2] for k in k0s])
    g = rng.standard_normal((2, n, ncols))
    for dt in (0.16, -0.05):
        out = E.midfft_cols(g, kx, vv, dt, batch=2)
        for s in range(2):
            assert rel_err(out[s], O.vdfdx_exponential(g[s], dt, kx[s], vv)) < TOL
    if n <= 1024:
        # charge density fused into the store phase: same f, n = trapz_v of it (whole grid and the slices of a v-shard)
        for edge in (3, 1, 2, 0):
            out2, dens = E.midfft_cols_density(g, kx, vv, 0.16, 0.05, edge_flags=edge, batch=2)
            w = np.full(ncols, 0.05)
            if edge & 1:
                w[0] *= 0.5
            if edge & 2:
                w[-1] *= 0.5
            for s in range(2):
                ref = O.vdfdx_exponential(g[s], 0.16, kx[s], vv)
                assert rel_err(out2[s], ref) < TOL
                assert rel_err(dens[s], (ref * w).sum(axis=1)) < TOL


@pytest.mark.parametrize("nx", [16, 32])
def test_tiny_transform_programs(nx):
    """tinyfft.cuh (nx = 16 / 32, the reference's Landau grid: whole transform in the registers of one thread) against the
    oracle: v df/dx of a batch with per-simulation wavenumbers, both signs of dt, a ragged last CTA; the field solve for
    odd and even numbers of rows with and without a driver"""
    rng = np.random.default_rng(nx)
    ncols = 2 * 131
    vv = np.linspace(-6.4, 6.4, ncols)
    k0s = (0.3, 0.41)
    kx = np.stack([O.spatial_grid(0.0, 2 * np.pi / k, nx)[2] for k in k0s])
    g = rng.standard_normal((2, nx, ncols))
    for dt in (0.16, -0.05, 0.5):
        out = E.tiny_cols(g, kx, vv, dt, batch=2)
        for s in range(2):
            assert rel_err(out[s], O.vdfdx_exponential(g[s], dt, kx[s], vv)) < TOL
    for batch in (1, 2, 5):
        k0b = 0.25 + 0.05 * np.arange(batch)
        grids = [O.spatial_grid(0.0, 2 * np.pi / k, nx) for k in k0b]
        ook = np.stack([gr[3] for gr in grids])
        n = 1.0 + 0.1 * rng.standard_normal((batch, nx))
        drv = np.stack([0.02 * np.sin(k * gr[1]) for k, gr in zip(k0b, grids)])
        ref = np.stack([O.solve_for_field(n[i], ook[i]) for i in range(batch)])
        assert np.max(np.abs(E.tiny_poisson(n, ook) - ref)) < TOL * np.abs(ref).max()
        assert np.max(np.abs(E.tiny_poisson(n, ook, drv) - (ref + drv))) < TOL * np.abs(ref).max()


@pytest.mark.parametrize("nx", [256, 512, 1024, 2048])
def test_poisson_mid_size_program(nx):
    """midfft.cuh in Poisson mode (two density rows packed per sequence, multiplier i one_over_kx, driver added by the
    store) against the oracle's vlapy/core/field.py:39-88: odd and even numbers of rows, per-row wavenumber grids (an
    ensemble), with and without driver rows"""
    rng = np.random.default_rng(nx)
    for batch in (1, 2, 5):
        k0s = 0.25 + 0.05 * np.arange(batch)
        grids = [O.spatial_grid(0.0, 2 * np.pi / k, nx) for k in k0s]
        ook = np.stack([g[3] for g in grids])
        n = 1.0 + 0.1 * rng.standard_normal((batch, nx))
        drv = np.stack([0.02 * np.sin(k * g[1]) for k, g in zip(k0s, grids)])
        ref = np.stack([O.solve_for_field(n[i], ook[i]) for i in range(batch)])
        assert np.max(np.abs(E.midfft_poisson(n, ook) - ref)) < TOL * np.abs(ref).max()
        assert np.max(np.abs(E.midfft_poisson(n, ook, drv) - (ref + drv))) < TOL * np.abs(ref).max()


@pytest.mark.parametrize("nx", [4096, 8192, 16384])
def test_poisson_single_launch_program(nx):
    """rowfft.cuh in Poisson mode (one CTA per density row: load 1 - n, multiplier i one_over_kx, add the driver)
    against the oracle's vlapy/core/field.py:39-88, with and without a driver row, three rows"""
    rng = np.random.default_rng(nx)
    dx, x, kx, ook = O.spatial_grid(0.0, 2 * np.pi / 0.35, nx)
    n = 1.0 + 0.1 * rng.standard_normal((3, nx))
    drv = 0.02 * np.sin(0.35 * x)
    ref = np.stack([O.solve_for_field(n[i], ook) for i in range(3)])
    assert np.max(np.abs(E.poisson_rowfft(n, ook, None) - ref)) < TOL * np.abs(ref).max()
    assert np.max(np.abs(E.poisson_rowfft(n, ook, drv) - (ref + drv))) < TOL * np.abs(ref).max()
