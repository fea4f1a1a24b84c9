"""GPU parity tests (-m gpu): the CUDA path, called through the operator mirrors / C ABI, against
the golden fixtures (outputs of the reference itself) and the CPU oracle on seeded inputs.

Tolerance: BASELINE.json asks for 1e-12 relative per operator in fp64, with the norm
max|a-b| / max|b|.  Sizes here are ones the oracle finishes in seconds; full-size
(16384^2) behaviour is covered by size-independent properties in test_gpu_fullsize.py."""
import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import vpfp_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vlapy_b200 import _lib
    _lib.lib()      # fail loudly if the extension is missing
    return torch.device("cuda:0")


def stuff_from(g_or_cfg, **extra):
    keys = ("kx", "kv", "v", "x", "one_over_kx")
    d = {k: np.asarray(g_or_cfg[k]) for k in keys}
    d["dv"] = float(g_or_cfg["dv"])
    d.update(extra)
    return d


@pytest.mark.parametrize("name", ["ops_small", "ops_c1", "ops_white"])
def test_operators_vs_reference_outputs(dev, name):
    from vlapy_b200.core import vlasov, field, step
    g = golden(name)
    f, e, dt, dv = g["f"], g["e"], float(g["dt"]), float(g["dv"])
    nx, nv = f.shape
    stuff = stuff_from(g, nx=nx, nv=nv)
    vdfdx = vlasov.get_vdfdx(stuff, "exponential")
    edfdv = vlasov.get_edfdv(stuff, "exponential")
    cd2 = vlasov.get_edfdv(stuff, "cd2")
    f_keep = f.copy()
    assert rel_err(vdfdx(f, dt), g["vdfdx"]) < TOL
    assert rel_err(vdfdx(f=f, dt=-0.066 * dt), g["vdfdx_neg"]) < TOL
    assert rel_err(edfdv(f, e, 0.5 * dt), g["edfdv"]) < TOL
    assert rel_err(edfdv(f=f, e=e, dt=-0.21 * dt), g["edfdv_neg"]) < TOL
    assert rel_err(cd2(f, e, 0.5 * dt), g["cd2"]) < TOL
    np.testing.assert_array_equal(f, f_keep)                      # operators are functional
    assert rel_err(field.compute_charges(f, dv), g["charges"]) < TOL
    fs = field.get_field_solver(stuff, "spectral")
    assert rel_err(fs(g["drv"], f), g["efield"]) < TOL
    assert rel_err(fs(driver_field=g["drv"], f=f), g["efield"]) < TOL
    # device tensors in -> device tensors out
    fd = torch.from_numpy(f).to(dev)
    out = vdfdx(fd, dt)
    assert isinstance(out, torch.Tensor) and out.is_cuda and out.data_ptr() != fd.data_ptr()
    assert rel_err(out.cpu().numpy(), g["vdfdx"]) < TOL
    # collisions through the reference's entry point (tests/test_collisions.py:205-211)
    for op in ("lb", "dg"):
        fp = step.get_collision_step(
            stuff_for_time_loop=dict(f=g["fpos"], v=g["v"], nv=nv, nx=nx, nu=float(g["nu"]), dt=dt, dv=dv),
            all_params={"fokker-planck": {"type": op, "solver": "batched_tridiagonal"}, "nu": float(g["nu"])})
        assert rel_err(fp(g["fpos"]), g[op + "_solve"]) < TOL
        assert rel_err(fp(f=g["fpos"]), g[op + "_solve"]) < TOL


def test_stored_quantities_vs_reference_outputs(dev):
    from vlapy_b200 import ops
    for name in ("ops_small", "ops_c1", "ops_white"):
        g = golden(name)
        dv = float(g["dv"])
        f = torch.from_numpy(g["fpos"]).to(dev)
        v = torch.from_numpy(g["v"]).to(dev)
        mom = ops.moments(f, v, dv)
        assert rel_err(mom[:6].cpu().numpy(), g["moments"]) < TOL
        ser = ops.series(mom, torch.from_numpy(g["e"]).to(dev), torch.from_numpy(g["drv"]).to(dev))
        np.testing.assert_allclose(ser.cpu().numpy(), g["series"], rtol=1e-12)
        modes = ops.xmodes(f, 2)[0].cpu().numpy()
        assert rel_err(modes, g["modes"]) < TOL
    # f ln f is NaN where f <= 0, as in numpy (vlapy/core/step.py:222-224)
    g = golden("ops_small")
    mom = ops.moments(torch.from_numpy(g["f"]).to(dev), torch.from_numpy(g["v"]).to(dev), float(g["dv"]))
    assert torch.isnan(mom[7]).any()


@pytest.mark.parametrize("shape", [(2, 4), (3, 64), (64, 512), (7, 2048), (6, 8192), (5, 16384), (34, 32768)])
def test_edfdv_sizes_vs_oracle(dev, shape):
    """every nv regime: one CTA per row pair (<= 8192) and the three-pass decomposition above,
    odd row counts included."""
    from vlapy_b200.core import vlasov
    nx, nv = shape
    rng = np.random.default_rng(nx + nv)
    dv, v, kv = O.velocity_grid(6.4, nv)
    f = O.maxwellian(nx, nv, 6.4) * (1 + 0.1 * rng.standard_normal((nx, 1))) + 1e-3 * rng.standard_normal((nx, nv))
    e = 0.05 * rng.standard_normal(nx)
    out = vlasov.get_edfdv_exponential(kv)(f, e, 0.125)
    assert rel_err(out, O.edfdv_exponential(f, e, 0.125, kv)) < TOL


@pytest.mark.parametrize("shape", [(2, 4), (32, 512), (256, 130), (2048, 34), (4096, 32), (16384, 18), (65536, 4)])
def test_vdfdx_sizes_vs_oracle(dev, shape):
    """every nx regime: whole columns in one CTA (<= 2048) and the three-pass decomposition,
    ragged column tiles included."""
    from vlapy_b200.core import vlasov
    nx, nv = shape
    rng = np.random.default_rng(nx + nv)
    v = np.linspace(-6.4, 6.4, nv)
    dx, x, kx, ook = O.spatial_grid(0.0, 2 * np.pi / 0.35, nx)
    f = np.exp(-v ** 2 / 2)[None, :] * (1 + 0.1 * np.sin(0.35 * x))[:, None] + 1e-3 * rng.standard_normal((nx, nv))
    out = vlasov.get_vdfdx_exponential(kx, v)(f, 0.25)
    assert rel_err(out, O.vdfdx_exponential(f, 0.25, kx, v)) < TOL


@pytest.mark.parametrize("flags", [0, 1, 2])
@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192, 16384])
def test_fast_and_generic_kernels_agree_with_oracle(dev, n, flags):
    """register-resident kernels with exact phases (0) and geometric tables (1), and the generic
    phase program (2), on white-noise-dominated input at the sizes the fast path serves."""
    from vlapy_b200 import ops
    rng = np.random.default_rng(n + flags)
    # e df/dv: an odd number of rows (phantom partner) of length n; enough cells (> 2^22) for the
    # three-pass kernels to be chosen at every n
    dv, v, kv = O.velocity_grid(6.4, n)
    nrows = max(5, (1 << 22) // n + 1)
    f = rng.standard_normal((nrows, n))
    e = 0.3 * rng.standard_normal(nrows)
    out = ops.edfdv_exp(torch.from_numpy(f).to(dev), torch.from_numpy(e).to(dev), torch.from_numpy(kv).to(dev),
                        0.125, flags=flags)
    assert rel_err(out.cpu().numpy(), O.edfdv_exponential(f, e, 0.125, kv)) < TOL
    # v df/dx: n rows, a ragged number of columns, two simulations with their own kx
    ncols = max(36, 2 * (((1 << 21) // n) // 2) + 4)
    vv = np.linspace(-6.4, 6.4, ncols)
    kxs = np.stack([O.spatial_grid(0.0, 2 * np.pi / k0, n)[2] for k0 in (0.3, 0.41)])
    g = rng.standard_normal((2, n, ncols))
    out = ops.vdfdx_exp(torch.from_numpy(g).to(dev), torch.from_numpy(kxs).to(dev), torch.from_numpy(vv).to(dev),
                        0.25, flags=flags)
    for b in range(2):
        assert rel_err(out[b].cpu().numpy(), O.vdfdx_exponential(g[b], 0.25, kxs[b], vv)) < TOL


@pytest.mark.parametrize("nx,ncols,edge", [(4096, 36, 3), (16384, 66, 1), (8192, 32, 2), (256, 64, 3), (16384, 32, 0)])
def test_vdfdx_fused_density(dev, nx, ncols, edge):
    """charge density reduced in the epilogue of the last pass == trapz_v of the result, for
    whole grids (edge = 3) and v-slices of a sharded grid (edge = 1, 2, 0)"""
    from vlapy_b200 import ops
    rng = np.random.default_rng(nx + ncols + edge)
    f = rng.standard_normal((nx, ncols)) + 1.0
    vv = np.linspace(-3.0, 3.0, ncols)
    kx = O.spatial_grid(0.0, 17.0, nx)[2]
    n = torch.empty(nx, dtype=torch.float64, device=dev)
    out = ops.vdfdx_exp(torch.from_numpy(f).to(dev), torch.from_numpy(kx).to(dev), torch.from_numpy(vv).to(dev), 0.3,
                        flags=1, density_out=n, dv=0.05, edge_flags=edge)
    ref = O.vdfdx_exponential(f, 0.3, kx, vv)
    assert rel_err(out.cpu().numpy(), ref) < TOL
    w = np.full(ncols, 0.05)
    if edge & 1:
        w[0] *= 0.5
    if edge & 2:
        w[-1] *= 0.5
    assert rel_err(n.cpu().numpy(), (ref * w).sum(axis=1)) < TOL


@pytest.mark.parametrize("nx,ncols", [(4096, 40), (512, 48), (256, 2048)])
def test_vdfdx_fused_density_of_an_ensemble(dev, nx, ncols):
    """the same for a batch of simulations with their own wavenumbers: three passes (nx = 4096: the partial sums are
    parked in a permuted order per simulation) and the mid-size single-pass kernel (density in the store phase)"""
    from vlapy_b200 import ops
    rng = np.random.default_rng(nx + ncols)
    B = 3
    f = rng.standard_normal((B, nx, ncols)) + 1.0
    vv = np.linspace(-3.0, 3.0, ncols)
    kx = np.stack([O.spatial_grid(0.0, 15.0 + 2.0 * b, nx)[2] for b in range(B)])
    n = torch.empty((B, nx), dtype=torch.float64, device=dev)
    out = ops.vdfdx_exp(torch.from_numpy(f).to(dev), torch.from_numpy(kx).to(dev), torch.from_numpy(vv).to(dev), 0.3,
                        flags=1, density_out=n, dv=0.05)
    w = np.full(ncols, 0.05)
    w[0] *= 0.5
    w[-1] *= 0.5
    for b in range(B):
        ref = O.vdfdx_exponential(f[b], 0.3, kx[b], vv)
        assert rel_err(out[b].cpu().numpy(), ref) < TOL
        assert rel_err(n[b].cpu().numpy(), (ref * w).sum(axis=1)) < TOL


def test_vdfdx_ensemble_with_per_simulation_kx(dev):
    from vlapy_b200.core import vlasov
    rng = np.random.default_rng(11)
    batch, nx, nv = 5, 64, 128
    dv, v, kv = O.velocity_grid(6.4, nv)
    k0s = np.linspace(0.25, 0.45, batch)
    kxs = np.stack([O.spatial_grid(0.0, 2 * np.pi / k0, nx)[2] for k0 in k0s])
    f = rng.standard_normal((batch, nx, nv))
    out = vlasov.get_vdfdx_exponential(kxs, v)(f, 0.2)
    for b in range(batch):
        assert rel_err(out[b], O.vdfdx_exponential(f[b], 0.2, kxs[b], v)) < TOL


def test_row_pitch_larger_than_row(dev):
    """v-sharded / padded layouts: operators take a row pitch (ld) different from the row length."""
    from vlapy_b200 import ops
    rng = np.random.default_rng(2)
    nx, nv = 64, 96
    big = torch.from_numpy(rng.standard_normal((nx, 2 * nv))).to(dev)
    f = big[:, 16:16 + nv]                       # pitch 192, 16-byte aligned offset
    v = torch.linspace(-3, 3, nv, dtype=torch.float64, device=dev)
    dx, x, kx, ook = O.spatial_grid(0.0, 20.0, nx)
    out = ops.vdfdx_exp(f, torch.from_numpy(kx).to(dev), v, 0.3)
    ref = O.vdfdx_exponential(f.cpu().numpy(), 0.3, kx, v.cpu().numpy())
    assert rel_err(out.cpu().numpy(), ref) < TOL
    mom = ops.moments(f, v, 0.1, nmom=6, edge_flags=1)   # local slice owning only the first global cell
    w = np.full(nv, 0.1); w[0] = 0.05
    np.testing.assert_allclose(mom[0].cpu().numpy(), (f.cpu().numpy() * w).sum(1), rtol=1e-13)


def test_field_solver_unit_cases(dev):
    """tests/test_fieldsolver.py of the reference (nx = 96, direct DFT path) + powers of two."""
    from vlapy_b200.core import field
    g = golden("fieldsolver_unit")
    x, kp = g["x"], 0.25
    analytic = [np.cos(kp * x) / kp, -np.sin(2 * kp * x) / 2.0 / kp,
                np.cos(2 * kp * x) / 2.0 / kp - np.sin(8 * kp * x) / 8.0 / kp]
    for i in range(3):
        e = field.solve_for_field(charge_density=g["rho_%d" % i], one_over_kx=g["one_over_kx"])
        assert rel_err(e, g["e_%d" % i]) < TOL
        np.testing.assert_almost_equal(e, analytic[i], decimal=4)
    rng = np.random.default_rng(4)
    for nx in (2, 8, 1024, 8192, 16384, 65536):
        dx, xx, kx, ook = O.spatial_grid(0.0, 17.0, nx)
        n = 1.0 + 0.1 * rng.standard_normal(nx)
        ref = O.solve_for_field(n, ook)
        err = np.max(np.abs(field.solve_for_field(n, ook) - ref))
        assert err < TOL * max(np.abs(ref).max(), 1e-3), (nx, err)     # nx = 2 has only DC + Nyquist: E = 0


@pytest.mark.parametrize("nx", [16, 32, 64, 256, 512, 1024, 2048, 4096, 16384])
def test_field_solve_of_an_ensemble_every_kernel(dev, nx):
    """the spectral field solve for batches of density rows with their own wavenumber grids and driver rows: in-register
    transforms (nx = 16, 32), generic program (64), mid-size kernel in Poisson mode (256 .. 2048, two rows per packed
    sequence: odd and even batch sizes), row-FFT kernel in Poisson mode (4096, 16384)"""
    from vlapy_b200 import ops
    rng = np.random.default_rng(nx)
    for batch in (1, 2, 5):
        k0s = 0.25 + 0.05 * np.arange(batch)
        grids = [O.spatial_grid(0.0, 2 * np.pi / k, nx) for k in k0s]
        ook = np.stack([g[3] for g in grids])
        n = 1.0 + 0.1 * rng.standard_normal((batch, nx))
        drv = np.stack([0.02 * np.sin(k * g[1]) for k, g in zip(k0s, grids)])
        ref = np.stack([O.solve_for_field(n[i], ook[i]) for i in range(batch)])
        nd, od, dd = (torch.from_numpy(a).to(dev) for a in (n, ook, drv))
        assert np.max(np.abs(ops.poisson(nd, od).cpu().numpy() - ref)) < TOL * np.abs(ref).max()
        assert np.max(np.abs(ops.poisson(nd, od, dd).cpu().numpy() - (ref + drv))) < TOL * np.abs(ref).max()


def test_collision_unit_cases(dev):
    """tests/test_collisions.py of the reference: 16 steps at nx=2, nv=1024, nu=1e-2, dt=0.1."""
    from vlapy_b200.core import step
    g = golden("collisions_unit")
    v, dv, nu, dt = g["v"], float(g["dv"]), float(g["nu"]), float(g["dt"])
    for vshift in (0.0, 0.5, 1.5):
        f0 = g["f_%g" % vshift]
        for op in ("lb", "dg"):
            fp = step.get_collision_step(
                stuff_for_time_loop=dict(f=f0, v=v, nv=1024, nx=2, nu=nu, dt=dt, dv=dv),
                all_params={"fokker-planck": {"type": op, "solver": "batched_tridiagonal"}, "nu": nu})
            f = f0.copy()
            for _ in range(16):
                f = fp(f)
            assert rel_err(f, g["out_%s_%g" % (op, vshift)]) < TOL
            if vshift == 0.0:
                np.testing.assert_almost_equal(f, f0, decimal=4)
                np.testing.assert_almost_equal(O.trapz_last(f * v, dv), O.trapz_last(f0 * v, dv), decimal=4)
            if vshift == 0.5:
                np.testing.assert_almost_equal(O.trapz_last(f, dv), O.trapz_last(f0, dv), decimal=4)
                np.testing.assert_almost_equal(O.trapz_last(f * v ** 2, dv), O.trapz_last(f0 * v ** 2, dv), decimal=4)
            if vshift == 1.5 and op == "lb":
                assert np.all(O.trapz_last(f * v, dv) < O.trapz_last(f0 * v, dv))
            if vshift == 1.5 and op == "dg":
                np.testing.assert_almost_equal(O.trapz_last(f * v, dv), O.trapz_last(f0 * v, dv), decimal=4)


@pytest.mark.parametrize("op", ["lb", "dg"])
@pytest.mark.parametrize("nv", [8, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("linspace_kernel", [False, True])
def test_fp_sizes_vs_oracle(dev, op, nv, linspace_kernel):
    """generic phase program (any v array) and the kernel specialised for np.linspace grids"""
    from vlapy_b200 import ops
    dv, v, kv = O.velocity_grid(6.4, nv)
    nu = 3.4e-6 if nv >= 4096 else 1e-3
    f = O.shifted_maxwellian(5, v, 1.0, 0.3) * np.array([1.0, 0.7, 1.3, 0.2, 2.0])[:, None]
    ref = O.collision_step(f, v, nu, 0.25, dv, op)
    mom = torch.zeros((8, 5), dtype=torch.float64, device=dev)
    vgrid = ops.linspace_params(v) if linspace_kernel else None
    assert (vgrid is not None) == linspace_kernel
    out = ops.fp_step(torch.from_numpy(f).to(dev), torch.from_numpy(v).to(dev), nu, 0.25, dv, op, moments_out=mom,
                      vgrid=vgrid)
    assert rel_err(out.cpu().numpy(), ref) < TOL
    assert rel_err(mom[:6].cpu().numpy(), O.field_moments(ref, v, dv)) < TOL
    ser = O.series_moments(ref, np.zeros(5), np.zeros(5), O.field_moments(ref, v, dv), dv)
    np.testing.assert_allclose(mom[6].cpu().numpy().mean(), ser[5], rtol=1e-12)
    np.testing.assert_allclose(mom[7].cpu().numpy().mean(), ser[6], rtol=1e-11)
    out2 = ops.fp_step(torch.from_numpy(f).to(dev), torch.from_numpy(v).to(dev), nu, 0.25, dv, op, vgrid=vgrid)
    assert torch.equal(out, out2)          # with and without the fused moments: same bits


@pytest.mark.parametrize("op", ["lb", "dg"])
def test_tridiag_two_stage_interface_vs_reference(dev, op):
    """the reference's explicit two-stage collision interface (vlapy/core/collisions.py:268-317) on the device:
    get_batched_array_maker against the diagonals the REFERENCE returned (golden ops_small), get_matrix_solver on
    them against the reference's solve; numpy in -> numpy out, CUDA in -> CUDA out, arguments untouched"""
    from vlapy_b200.core import collisions
    g = golden("ops_small")
    nu, dt, dv, v, f = float(g["nu"]), float(g["dt"]), float(g["dv"]), g["v"], g["fpos"]
    nx, nv = f.shape
    maker = collisions.get_batched_array_maker(v, nv, nx, nu, dt, dv, operator=op)
    a, b, c = maker(f)
    assert isinstance(a, np.ndarray)
    for got, name in ((a, "_a"), (b, "_b"), (c, "_c")):
        ref = g[op + name]
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= 1e-13 * np.max(np.abs(ref)), name
    for solver_name in ("batched_tridiagonal", "naive"):
        solve = collisions.get_matrix_solver(nx, nv, solver_name)
        keep = [t.copy() for t in (a, b, c, f)]
        x = solve(a, b, c, f)
        assert rel_err(x, g[op + "_solve"]) < TOL
        assert all(np.array_equal(k, t) for k, t in zip(keep, (a, b, c, f)))
    fd = torch.from_numpy(f).to(dev)
    ad, bd, cd = maker(fd)
    assert ad.is_cuda and rel_err(ad.cpu().numpy(), g[op + "_a"]) < 1e-13
    xd = collisions.get_matrix_solver(nx, nv)(ad, bd, cd, fd)
    assert xd.is_cuda and rel_err(xd.cpu().numpy(), g[op + "_solve"]) < TOL
    with pytest.raises(NotImplementedError):
        collisions.get_matrix_solver(nx, nv, "cholesky")
    with pytest.raises(NotImplementedError):
        collisions.get_batched_array_maker(v, nv, nx, nu, dt, dv, operator="krook")


@pytest.mark.parametrize("nv", [8, 100, 1000, 2048, 16384])
def test_tridiag_general_diagonals_vs_thomas(dev, nv):
    """general (non-constant, non-symmetric) diagonally dominant systems, ragged sizes, against the reference's
    Thomas sweep restated in the oracle (vlapy/core/collisions.py:232-263)"""
    from vlapy_b200 import ops
    rng = np.random.default_rng(nv)
    rows = 37
    a = rng.uniform(-1, 1, (rows, nv - 1))
    c = rng.uniform(-1, 1, (rows, nv - 1))
    b = 2.5 + rng.uniform(0, 1, (rows, nv))
    d = rng.standard_normal((rows, nv))
    ref = O.thomas_batched(a, b, c, d)
    t = [torch.from_numpy(z).to(dev) for z in (a, b, c, d)]
    out = ops.tridiag_solve(*t).cpu().numpy()
    assert rel_err(out, ref) < TOL
    with pytest.raises(NotImplementedError):            # nv = 4 < 8
        ops.tridiag_solve(t[0][:, :3].contiguous(), t[1][:, :4].contiguous(), t[2][:, :3].contiguous(),
                          t[3][:, :4].contiguous())


@pytest.mark.parametrize("ncols", [33, 64, 130, 512, 2048])
@pytest.mark.parametrize("nmom", [1, 6, 8])
def test_moments_many_short_rows(dev, ncols, nmom):
    """the warp-per-row moments kernel (>= 1024 rows of <= 2048 cells: ensembles, C4) against the oracle; odd and
    even column counts (scalar and 16-byte loads), v-slices with and without the global end cells"""
    from vlapy_b200 import ops
    rows = 1500
    rng = np.random.default_rng(ncols + nmom)
    dv, v, kv = O.velocity_grid(6.4, ncols) if ncols % 2 == 0 else (0.1, np.linspace(-1.6, 1.6, ncols), None)
    f = np.exp(-v ** 2 / 2)[None, :] * (1 + 0.1 * rng.random((rows, ncols)))
    fd, vd = torch.from_numpy(f).to(dev), torch.from_numpy(v).to(dev)
    for edge in (3, 1, 0):
        w = np.full(ncols, dv)
        if edge & 1:
            w[0] *= 0.5
        if edge & 2:
            w[-1] *= 0.5
        got = ops.moments(fd, vd, dv, nmom=nmom, edge_flags=edge).cpu().numpy()
        assert got.shape == (nmom, rows)
        for k in range(nmom):
            ref = (f * v ** k * w).sum(1) if k < 6 else ((f * f * w).sum(1) if k == 6 else (f * np.log(f) * w).sum(1))
            assert np.max(np.abs(got[k] - ref)) <= 1e-12 * max(np.max(np.abs(ref)), 1e-30), (k, edge)


def field_tolerances(cfg, fmax):
    """Per-field absolute tolerance that follows from 1e-12 relative parity on f (norm max|f|):
    a v-moment of order p is a linear functional of f with L1 weight int |v|^p dv, and E is the
    Poisson solve of the p = 0 moment (gain <= 1/k0 per mode, summed over a few modes)."""
    vmax = float(np.abs(cfg["v"]).max())
    tol = {}
    for p, name in enumerate(("n", "j", "T", "q", "fv4", "vN")):
        tol[name] = 1e-12 * fmax * 2.0 * vmax ** (p + 1) / (p + 1)
    tol["e"] = 10.0 * tol["n"] / cfg["k0"]
    tol["driver"] = 1e-15 * max(p_["a0"] * p_["k0"] for p_ in cfg["pulses"].values()) * 10
    return tol


def make_stuff(cfg, rules, with_pulses=True):
    stuff = {k: cfg[k] for k in ("kx", "x", "one_over_kx", "v", "kv", "nv", "nx", "dv", "dt", "nu",
                                 "driver_function")}
    stuff.update(e=cfg["e0"], f=cfg["f0"], rules_to_store_f=rules)
    if with_pulses:
        stuff["pulse_dictionary"] = cfg["pulses"]
    return stuff


def make_params(cfg, integ="leapfrog", op="lb", edfdv="exponential"):
    return {"backend": {"core": "b200"}, "nu": cfg["nu"],
            "vlasov-poisson": {"time": integ, "vdfdx": "exponential", "edfdv": edfdv, "poisson": "spectral"},
            "fokker-planck": {"type": op, "solver": "batched_tridiagonal"}}


RULES = {"time": "first-last", "space": ["k0", "k1"]}


@pytest.mark.parametrize("integ", ["leapfrog", "pefrl", "h-sixth"])
@pytest.mark.parametrize("device_driver", [True, False])
def test_vp50_schedules_vs_reference(dev, integ, device_driver):
    """50 Vlasov-Poisson steps at C1 for each splitting schedule."""
    from vlapy_b200.core import step
    g = golden("vp50_c1")
    cfg = O.landau_config()
    vp = step.get_vlasov_poisson_step(make_params(cfg, integ), make_stuff(cfg, RULES, device_driver))
    e = torch.from_numpy(cfg["e0"]).to(dev)
    f = torch.from_numpy(cfg["f0"]).to(dev)
    for i in range(50):
        e, f = vp(e=e, f=f, t=cfg["dt"] * i)
    assert rel_err(f.cpu().numpy(), g["f_" + integ]) < TOL
    assert np.max(np.abs(e.cpu().numpy() - g["e_" + integ])) < 5e-14


def run_inner_loops(cfg, params, steps_in_loop, n_loops):
    from vlapy_b200 import outer_loop
    stuff = make_stuff(cfg, RULES)
    sim, inner = outer_loop.get_sim_config_and_inner_loop_step(params, stuff, steps_in_loop, RULES)
    outs = []
    for li in range(n_loops):
        t = cfg["dt"] * np.arange(li * steps_in_loop, (li + 1) * steps_in_loop)
        drv = np.stack([cfg["driver_function"](ti) for ti in t])
        sim = inner(time_array=t, driver_array=drv, temp_storage=sim)
        outs.append({"fields": {k: v.copy() for k, v in sim["fields"].items()},
                     "series": {k: np.array(v).copy() for k, v in sim["series"].items()},
                     "stored_f": sim["stored_f"].copy(), "f": sim["f"].copy(), "e": sim["e"].copy(),
                     "time": sim["time_batch"].copy()})
    return outs


def test_landau_damping_integrated(dev):
    """tests/test_landau_damping.py of the reference re-expressed on the b200 backend: 800 steps
    (2 inner loops of 400, vlapy/manager.py:61-83), damping rate of E_k1 vs the dispersion root."""
    g = golden("landau_c1")
    cfg = O.landau_config()
    steps, loops = O.steps_in_loop_like_manager(32, 512, 500)
    assert (steps, loops) == (400, 2)
    for integ in ("leapfrog", "pefrl", "h-sixth"):
        outs = run_inner_loops(cfg, make_params(cfg, integ), steps, loops)
        e_hist = np.concatenate([o["fields"]["e"] for o in outs])
        tax = np.concatenate([o["time"] for o in outs])
        rate = O.damping_rate(e_hist, tax)
        assert abs(rate - float(g["nu_ld"])) < 1.5e-4                       # the reference's own bar
        assert abs(rate - float(g["rate_%s_exponential" % integ])) < 1e-7   # and its own number
        assert np.max(np.abs(outs[-1]["e"] - g["e_final_" + integ])) < 1e-13
        assert abs(outs[-1]["series"]["mean_n"][-1] - 1.0) < 1e-13
        if integ == "leapfrog":
            o = outs[0]
            tol = field_tolerances(cfg, 0.4)
            for k in ("e", "driver", "n", "j", "T", "q", "fv4", "vN"):
                err = np.max(np.abs(o["fields"][k][:12] - g["fields_" + k]))
                assert err < tol[k], (k, err, tol[k])
            for k in O.SERIES_KEYS + ("mean_cum_de2",):
                np.testing.assert_allclose(o["series"][k][:12], g["series_" + k], rtol=1e-9, atol=1e-14, err_msg=k)
            assert o["stored_f"].dtype == np.complex64
            assert rel_err(o["stored_f"][:12], g["stored_f"]) < 1e-6          # complex64 storage
            assert rel_err(outs[-1]["f"], g["f_final_leapfrog"]) < TOL
            assert np.max(np.abs(e_hist - g["e_hist_leapfrog"])) < 1e-13


def test_landau_damping_cd2(dev):
    g = golden("landau_c1")
    cfg = O.landau_config()
    outs = run_inner_loops(cfg, make_params(cfg, "leapfrog", edfdv="cd2"), 400, 2)
    e_hist = np.concatenate([o["fields"]["e"] for o in outs])
    tax = np.concatenate([o["time"] for o in outs])
    rate = O.damping_rate(e_hist, tax)
    assert abs(rate - float(g["nu_ld"])) < 1.5e-4
    assert abs(rate - float(g["rate_leapfrog_cd2"])) < 1e-7


@pytest.mark.parametrize("op", ["lb", "dg"])
def test_nlepw_c2_40_steps_vs_reference(dev, op):
    """C2 (256 x 2048, run_nlepw.py, k0 = 0.35): 40 collisional leapfrog steps."""
    g = golden("nlepw_c2")
    cfg = O.nlepw_config()
    outs = run_inner_loops(cfg, make_params(cfg, "leapfrog", op), 40, 1)
    f, e = outs[0]["f"], outs[0]["e"]
    assert rel_err(e, g["e_" + op]) < 1e-11
    assert rel_err(f[::8, ::16], g["f_sub_" + op]) < TOL
    assert abs(f.sum() / float(g["f_sum_" + op]) - 1) < 1e-13
    assert abs(f[100, 1300] / float(g["f_100_1300_" + op]) - 1) < 1e-12


def test_nlepw_c2_200_step_series_vs_reference(dev):
    """SURVEY 8d integrated acceptance: C2 (256 x 2048, run_nlepw.py, lb), 200 steps through the public inner loop;
    every per-step series (mean_n, mean_j, mean_T, mean_e2, mean_de2, mean_f2, mean_flogf and the cumulative driver
    energy) against the REFERENCE's own inner loop (golden nlepw_c2_series200) to 1e-9, fields and final state"""
    g = golden("nlepw_c2_series200")
    cfg = O.nlepw_config()
    outs = run_inner_loops(cfg, make_params(cfg, "leapfrog", "lb"), 200, 1)
    o = outs[0]
    for k in O.SERIES_KEYS + ("mean_cum_de2", "mean_t_plus_e2_minus_cum_de2", "mean_t_plus_e2_plus_cum_de2"):
        np.testing.assert_allclose(o["series"][k], g["series_" + k], rtol=1e-9, atol=1e-14, err_msg=k)
    tol = field_tolerances(cfg, 0.4)
    for k in ("e", "n", "T"):
        err = np.max(np.abs(o["fields"][k][::20] - g["fields_%s_sub" % k]))
        assert err < 10 * tol[k], (k, err, tol[k])        # 200 steps of accumulated rounding
    assert rel_err(o["e"], g["e_final"]) < 1e-10
    assert rel_err(o["f"][::8, ::16], g["f_final_sub"]) < 5e-12


def test_c3_three_steps_vs_oracle(dev):
    """BASELINE config 3 (4096 x 4096, collisional NLEPW): three integrated steps through the public inner loop
    against the oracle on the full grid (the oracle needs ~7 s per step and ~1 GB: the slowest test of the suite)"""
    cfg = O.nlepw_config(nx=4096, nv=4096)
    outs = run_inner_loops(cfg, make_params(cfg, "leapfrog", "lb"), 3, 1)
    e_ref, f_ref, hist = O.run_steps(cfg, 3, "leapfrog", "lb", collect=True)
    o = outs[0]
    assert rel_err(o["f"], f_ref) < TOL
    assert rel_err(o["e"], e_ref) < 1e-10
    for j, k in enumerate(O.SERIES_KEYS):
        np.testing.assert_allclose(o["series"][k], hist["series"][:, j], rtol=1e-9, atol=1e-14, err_msg=k)
    assert np.max(np.abs(o["fields"]["n"] - hist["mom"][:, 0])) < 1e-12


def test_semi_lagrangian_operators_vs_reference(dev):
    """N4: get_vdfdx(stuff, "sl") / get_edfdv(stuff, "sl") (vlapy/core/vlasov.py:42-80, 168-210, 228-260) against
    outputs of the REFERENCE's own sl operators (golden sl_ops: seeded noisy states, three time steps incl. shifts
    beyond one cell where FITPACK clamps the feet); numpy in -> numpy out, CUDA in -> CUDA out; larger and ragged
    grids against the oracle"""
    from vlapy_b200.core import vlasov
    g = golden("sl_ops")
    for tag in ("small", "c1"):
        f, e, x, v = g[tag + "_f"], g[tag + "_e"], g[tag + "_x"], g[tag + "_v"]
        stuff = {"x": x, "v": v}
        vdfdx, edfdv = vlasov.get_vdfdx(stuff, "sl"), vlasov.get_edfdv(stuff, "sl")
        for name, dt in zip("abc", g[tag + "_dts"]):
            keep = f.copy()
            assert rel_err(vdfdx(f, dt), g["%s_vdfdx_%s" % (tag, name)]) < TOL
            assert rel_err(edfdv(f, e, dt), g["%s_edfdv_%s" % (tag, name)]) < TOL
            assert np.array_equal(keep, f)
        out = vdfdx(torch.from_numpy(f).to(dev), 0.16)
        assert out.is_cuda and rel_err(out.cpu().numpy(), g[tag + "_vdfdx_a"]) < TOL
    rng = np.random.default_rng(11)
    for nx, nv in ((7, 33), (100, 130), (512, 2048), (2048, 1000)):
        x, v = np.linspace(0.3, 20.0, nx), np.linspace(-6.0, 6.0, nv)
        f = np.exp(-v ** 2 / 2)[None, :] * (1 + 0.1 * np.sin(0.3 * x))[:, None] + 1e-3 * rng.standard_normal((nx, nv))
        e = 0.7 * np.cos(0.3 * x)
        stuff = {"x": x, "v": v}
        for dt in (0.2, -1.3):
            # np.linspace axes are uniform only to rounding; FITPACK takes the knots as given, the kernels one spacing
            # (ax[2] - ax[1], as the reference pads): 1e-11 on white-noise states (looser parity)
            assert rel_err(vlasov.get_vdfdx(stuff, "sl")(f, dt), O.vdfdx_sl(f, dt, x, v)) < 1e-11
            assert rel_err(vlasov.get_edfdv(stuff, "sl")(f, e, dt), O.edfdv_sl(f, e, dt, x, v)) < 1e-11
    with pytest.raises(NotImplementedError):
        vlasov.get_vdfdx({"x": x, "v": v, "kx": x}, "weno")


@pytest.mark.parametrize("vdfdx,edfdv", [("sl", "exponential"), ("exponential", "sl"), ("sl", "sl")])
def test_landau_damping_semi_lagrangian(dev, vdfdx, edfdv):
    """tests/test_landau_damping.py of the reference with the sl flavours (leapfrog, 800 steps through the public
    inner loop): the reference's own bar on the damping rate, and its own numbers (golden sl_ops)"""
    g = golden("sl_ops")
    cfg = O.landau_config()
    params = make_params(cfg, "leapfrog", edfdv=edfdv)
    params["vlasov-poisson"]["vdfdx"] = vdfdx
    outs = run_inner_loops(cfg, params, 400, 2)
    e_hist = np.concatenate([o["fields"]["e"] for o in outs])
    tax = np.concatenate([o["time"] for o in outs])
    rate = O.damping_rate(e_hist, tax)
    key = "%s_%s" % (vdfdx, edfdv)
    assert abs(rate - float(g["nu_ld"])) < 1.5e-4
    assert abs(rate - float(g["rate_" + key])) < 1e-7
    assert np.max(np.abs(outs[-1]["e"] - g["e_final_" + key])) < 1e-12
    # 1600 spline operators at ~1e-13 each (the line-by-line form against FITPACK's tensor-product solve)
    assert rel_err(outs[-1]["f"][::2, ::8], g["f_final_sub_" + key]) < 1e-9


@pytest.mark.parametrize("n", [256, 512, 1024, 2048])
def test_mid_size_single_pass_kernels(dev, n):
    """midfft.cuh through the C ABI (uniform grids, VPFP_PHASE_TABLE): e df/dv at nv = n (odd row count, a row pitch
    larger than the row, many tiles) and v df/dx at nx = n (ragged column tile, per-simulation wavenumbers) against the
    oracle, and against the generic kernels the same call used before (VPFP_FORCE_GENERIC)"""
    from vlapy_b200 import ops
    rng = np.random.default_rng(n + 3)
    dv, v, kv = O.velocity_grid(6.4, n)
    rows = 301
    f = np.exp(-v ** 2 / 2)[None, :] * (1 + 0.1 * rng.standard_normal((rows, n)))
    f[::7] = rng.standard_normal((len(f[::7]), n))
    e = 0.3 * rng.standard_normal(rows)
    big = torch.zeros((rows, n + 8), dtype=torch.float64, device=dev)
    big[:, :n] = torch.from_numpy(f).to(dev)
    fd, ed, kd = big[:, :n], torch.from_numpy(e).to(dev), torch.from_numpy(kv).to(dev)
    for dt in (0.125, -0.033):
        ref = O.edfdv_exponential(f, e, dt, kv)
        out = ops.edfdv_exp(fd, ed, kd, dt, flags=ops.PHASE_TABLE).cpu().numpy()
        gen = ops.edfdv_exp(fd, ed, kd, dt, flags=ops.PHASE_TABLE | ops.FORCE_GENERIC).cpu().numpy()
        assert rel_err(out, ref) < TOL
        assert rel_err(out, gen) < TOL
        assert not np.array_equal(out, gen)               # two different kernels really ran
    ncols = 54
    vv = np.linspace(-6.4, 6.4, ncols)
    k0s = (0.3, 0.35, 0.41)
    kx = np.stack([O.spatial_grid(0.0, 2 * np.pi / k, n)[2] for k in k0s])
    g = rng.standard_normal((3, n, ncols))
    gd, kxd, vd = torch.from_numpy(g).to(dev), torch.from_numpy(kx).to(dev), torch.from_numpy(vv).to(dev)
    nd = torch.zeros((3, n), dtype=torch.float64, device=dev)
    for dt in (0.16, -0.05):
        out = ops.vdfdx_exp(gd, kxd, vd, dt, flags=ops.PHASE_TABLE, density_out=nd, dv=0.25).cpu().numpy()
        for s in range(3):
            ref = O.vdfdx_exponential(g[s], dt, kx[s], vv)
            assert rel_err(out[s], ref) < TOL
            w = np.full(ncols, 0.25); w[0] = w[-1] = 0.125
            assert np.max(np.abs(nd[s].cpu().numpy() - (ref * w).sum(1))) < 1e-12 * np.abs(ref).max() * ncols


def test_small_collisional_run_through_inner_loop(dev):
    """16 x 128 collisional run through two inner loops of the reference (everything the storage
    layer receives: fields, series, stored modes, final state)."""
    g = golden("nlepw_c2")
    cfg = O.nlepw_config(nx=16, nv=128, k0=0.35, log_nu=-2)
    assert abs(cfg["nu"] / float(g["small_nu"]) - 1) < 1e-14
    outs = run_inner_loops(cfg, make_params(cfg, "leapfrog", "lb"), 24, 2)
    tol = field_tolerances(cfg, 0.4)
    for li, o in enumerate(outs):
        for k in ("e", "driver", "n", "j", "T", "q", "fv4", "vN"):
            ref = g["small_fields_%s_%d" % (k, li)]
            err = np.max(np.abs(o["fields"][k] - ref))
            assert err < tol[k], (k, li, err, tol[k])
        for k in O.SERIES_KEYS + ("mean_cum_de2", "mean_t_plus_e2_minus_cum_de2", "mean_t_plus_e2_plus_cum_de2"):
            np.testing.assert_allclose(o["series"][k], g["small_series_%s_%d" % (k, li)], rtol=1e-9, atol=1e-13,
                                       err_msg=k)
        assert rel_err(o["stored_f"], g["small_stored_f_%d" % li]) < 1e-6
        assert rel_err(o["f"], g["small_f_%d" % li]) < TOL
        assert rel_err(o["e"], g["small_e_%d" % li]) < 1e-11


def test_cuda_graph_inner_loop_equals_eager(dev):
    """the captured-step inner loop (default for small grids) and the eager one give the same
    stored quantities and final state, for a schedule with several driver sub-step times"""
    from vlapy_b200 import outer_loop
    cfg = O.nlepw_config(nx=32, nv=256, k0=0.35, log_nu=-2)
    res = {}
    for mode in (True, False):
        params = make_params(cfg, "pefrl", "dg")
        params["backend"]["cuda_graph"] = mode
        stuff = make_stuff(cfg, RULES)
        sim, inner = outer_loop.get_sim_config_and_inner_loop_step(params, stuff, 12, RULES)
        for li in range(2):
            t = cfg["dt"] * np.arange(li * 12, (li + 1) * 12)
            drv = np.stack([cfg["driver_function"](ti) for ti in t])
            sim = inner(time_array=t, driver_array=drv, temp_storage=sim)
        res[mode] = {"f": sim["f"].copy(), "e": sim["e"].copy(), "T": sim["fields"]["T"].copy(),
                     "series": {k: np.array(v).copy() for k, v in sim["series"].items()},
                     "stored_f": sim["stored_f"].copy()}
    g, e = res[True], res[False]
    assert rel_err(g["f"], e["f"]) < 1e-14 and rel_err(g["e"], e["e"]) < 1e-13
    np.testing.assert_allclose(g["T"], e["T"], rtol=1e-13)
    for k in e["series"]:
        np.testing.assert_allclose(g["series"][k], e["series"][k], rtol=1e-12, atol=1e-16, err_msg=k)
    np.testing.assert_allclose(g["stored_f"], e["stored_f"], rtol=1e-6, atol=1e-9)
    # and both agree with the oracle
    e_ref, f_ref = O.run_steps(cfg, 24, "pefrl", "dg")
    assert rel_err(g["f"], f_ref) < TOL


@pytest.mark.parametrize("integ,nu_log", [("leapfrog", None), ("pefrl", -2)])
def test_ensemble_matches_independent_oracle_runs(dev, integ, nu_log):
    """BASELINE config 4 in miniature: a k-sweep of independent simulations in one batch; every
    member must follow the reference's timestep exactly as if it ran alone."""
    from vlapy_b200 import ensemble
    k0s = [0.27, 0.3, 0.35, 0.41, 0.44]
    cfgs = [O.make_config(32, 128, k0, tmax=80, nt=500, a0=1e-3, t_R=20, log_nu_over_nu_ld=nu_log) for k0 in k0s]
    nu = cfgs[1]["nu"]
    for c in cfgs:
        c["nu"] = nu                      # one collision frequency for the whole batch
    stuff = dict(kx=np.stack([c["kx"] for c in cfgs]), one_over_kx=np.stack([c["one_over_kx"] for c in cfgs]),
                 x=np.stack([c["x"] for c in cfgs]), v=cfgs[0]["v"], kv=cfgs[0]["kv"], dv=cfgs[0]["dv"],
                 dt=cfgs[0]["dt"], nu=nu, pulses=[c["pulses"] for c in cfgs])
    params = make_params(cfgs[0], integ, "lb")
    step_fn = ensemble.get_ensemble_timestep(params, stuff)
    state = {"e": torch.from_numpy(np.stack([c["e0"] for c in cfgs])).to(dev),
             "f": torch.from_numpy(np.stack([c["f0"] for c in cfgs])).to(dev)}
    nsteps = 6
    for i in range(nsteps):
        state = step_fn(state, cfgs[0]["dt"] * i)
    f, e = state["f"].cpu().numpy(), state["e"].cpu().numpy()
    mom, ser = state["moments"].cpu().numpy(), state["series"].cpu().numpy()
    for b, c in enumerate(cfgs):
        e_ref, f_ref, hist = O.run_steps(c, nsteps, integ, "lb", collect=True)
        assert rel_err(f[b], f_ref) < TOL
        assert np.max(np.abs(e[b] - e_ref)) < 1e-13
        assert rel_err(mom[:3, b], hist["mom"][-1][:3]) < TOL
        np.testing.assert_allclose(ser[b, :6], hist["series"][-1][:6], rtol=1e-9, atol=1e-13)
    assert ensemble.shard(1024, 3, 8) == (384, 512) and ensemble.shard(10, 3, 4) == (9, 10)


def test_ensemble_at_c4_member_size_vs_oracle(dev):
    """BASELINE config 4 at its real member size: a batch of 256 x 512 Landau-damping simulations with their own
    box lengths (k0 from the sweep's ends and middle), 12 collisionless leapfrog steps through the batched kernels
    (mid-size single-pass advection, warp-per-row moments) against the oracle run of every member"""
    from vlapy_b200 import ensemble
    k0s = [0.25, 0.3, 0.35, 0.45]
    cfgs = [O.make_config(256, 512, k0, tmax=80, nt=500, a0=1e-3, t_R=20) for k0 in k0s]
    stuff = dict(kx=np.stack([c["kx"] for c in cfgs]), one_over_kx=np.stack([c["one_over_kx"] for c in cfgs]),
                 x=np.stack([c["x"] for c in cfgs]), v=cfgs[0]["v"], kv=cfgs[0]["kv"], dv=cfgs[0]["dv"],
                 dt=cfgs[0]["dt"], nu=0.0, pulses=[c["pulses"] for c in cfgs])
    params = make_params(cfgs[0], "leapfrog", "lb")
    params["nu"] = 0.0
    step_fn = ensemble.get_ensemble_timestep(params, stuff)
    state = {"e": torch.from_numpy(np.stack([c["e0"] for c in cfgs])).to(dev),
             "f": torch.from_numpy(np.stack([c["f0"] for c in cfgs])).to(dev)}
    nsteps = 12
    for i in range(nsteps):
        state = step_fn(state, cfgs[0]["dt"] * i)
    f, e = state["f"].cpu().numpy(), state["e"].cpu().numpy()
    mom, ser = state["moments"].cpu().numpy(), state["series"].cpu().numpy()
    for b, c in enumerate(cfgs):
        c = dict(c, nu=0.0)
        e_ref, f_ref, hist = O.run_steps(c, nsteps, "leapfrog", "lb", collect=True)
        assert rel_err(f[b], f_ref) < TOL
        assert np.max(np.abs(e[b] - e_ref)) < 1e-13
        assert rel_err(mom[:3, b], hist["mom"][-1][:3]) < TOL
        np.testing.assert_allclose(ser[b, :7], hist["series"][-1][:7], rtol=1e-9, atol=1e-13)


def test_smoke_entry(dev):
    import __graft_entry__ as ge
    assert ge.smoke()


@pytest.mark.parametrize("nv,rows", [(4096, 1024), (8192, 513), (16384, 257), (16384, 1024)])
def test_single_pass_row_kernel(dev, nv, rows):
    """e df/dv as one kernel per row (csrc/rowfft.cuh) against the oracle and against the three-pass
    kernels; smooth + noise rows, both signs of dt, a row pitch larger than the row."""
    from vlapy_b200 import ops
    rng = np.random.default_rng(nv + rows)
    dv, v, kv = O.velocity_grid(6.4, nv)
    f = np.exp(-v ** 2 / 2)[None, :] * (1 + 0.1 * rng.standard_normal((rows, nv)))
    f[::7] = rng.standard_normal((len(f[::7]), nv))
    e = 0.3 * rng.standard_normal(rows)
    big = torch.zeros((rows, nv + 16), dtype=torch.float64, device=dev)
    big[:, :nv] = torch.from_numpy(f).to(dev)
    fd, ed, kd = big[:, :nv], torch.from_numpy(e).to(dev), torch.from_numpy(kv).to(dev)
    for dt in (0.125, -0.033):
        ref = O.edfdv_exponential(f, e, dt, kv)
        one = ops.edfdv_exp(fd, ed, kd, dt, flags=ops.PHASE_TABLE)
        three = ops.edfdv_exp(fd, ed, kd, dt, flags=ops.PHASE_TABLE | ops.FORCE_THREE_PASS)
        assert ops.rowfft_serves(rows, nv, ops.PHASE_TABLE)
        assert rel_err(one.cpu().numpy(), ref) < TOL
        assert rel_err(three.cpu().numpy(), ref) < TOL
        assert rel_err(one.cpu().numpy(), three.cpu().numpy()) < TOL


def test_scattering_vdfdx_with_emulated_ranks(dev):
    """the scattering v df/dx of the multi-GPU path (last pass stores into the peers' x-shards) with four ranks emulated
    in one process: every "peer" buffer is local.  Result and summed partial densities against the one-grid operator
    (same kernels per column) and the oracle."""
    from vlapy_b200 import ops
    nx, nv, P = 2048, 8192, 4
    nxl, nvl = nx // P, nv // P
    cfg = O.nlepw_config(nx=nx, nv=nv, k0=0.35, log_nu=-2)
    rng = np.random.default_rng(5)
    f = cfg["f0"] * (1.0 + 0.1 * np.sin(0.35 * cfg["x"]))[:, None] + 1e-3 * rng.standard_normal((nx, nv))
    fd = torch.from_numpy(f).to(dev)
    kx, v = torch.from_numpy(cfg["kx"]).to(dev), torch.from_numpy(cfg["v"]).to(dev)
    dt, dv = 0.25, float(cfg["dv"])
    n_one = torch.empty(nx, dtype=torch.float64, device=dev)
    one = ops.vdfdx_exp(fd, kx, v, dt, flags=1, density_out=n_one, dv=dv)
    shards = [torch.zeros((nxl, nv), dtype=torch.float64, device=dev) for _ in range(P)]
    ptrs = [t.data_ptr() for t in shards]
    scratch = torch.empty((nx, nvl), dtype=torch.float64, device=dev)
    n_sum = torch.zeros(nx, dtype=torch.float64, device=dev)
    for r in range(P):
        fv = fd[:, r * nvl:(r + 1) * nvl].contiguous()
        n_r = torch.empty(nx, dtype=torch.float64, device=dev)
        edge = (1 if r == 0 else 0) | (2 if r == P - 1 else 0)
        ops.vdfdx_exp_scatter(fv, kx, v[r * nvl:(r + 1) * nvl].contiguous(), dt, scratch, ptrs, r, flags=1,
                              density_out=n_r, dv=dv, edge_flags=edge)
        n_sum += n_r
    got = torch.cat(shards, dim=0)
    assert rel_err(got.cpu().numpy(), one.cpu().numpy()) < 1e-14      # same kernels, the phase tables start per shard
    assert rel_err(n_sum.cpu().numpy(), n_one.cpu().numpy()) < 1e-14
    assert rel_err(got.cpu().numpy(), O.vdfdx_exponential(f, dt, cfg["kx"], cfg["v"])) < TOL


def test_resume_from_a_stored_state(dev):
    """a FRESH inner loop restarted from the host copy of the state after loop 1 (what the
    storage layer wrote as full_distribution) reproduces loop 2 of an uninterrupted run"""
    from vlapy_b200 import outer_loop
    cfg = O.nlepw_config(nx=32, nv=256, k0=0.35, log_nu=-2)
    params = make_params(cfg, "leapfrog", "lb")
    nt = 10
    straight = run_inner_loops(cfg, params, nt, 2)
    sim, inner = outer_loop.get_sim_config_and_inner_loop_step(params, make_stuff(cfg, RULES), nt, RULES)
    sim = outer_loop.resume_from(sim, straight[0]["f"], straight[0]["e"],
                                 mean_cum_de2_previous=straight[0]["series"]["mean_cum_de2"][-1])
    t = cfg["dt"] * np.arange(nt, 2 * nt)
    drv = np.stack([cfg["driver_function"](ti) for ti in t])
    sim = inner(time_array=t, driver_array=drv, temp_storage=sim)
    assert np.array_equal(sim["f"], straight[1]["f"]) and np.array_equal(sim["e"], straight[1]["e"])
    for k in ("mean_n", "mean_T", "mean_e2", "mean_cum_de2"):
        np.testing.assert_allclose(sim["series"][k], straight[1]["series"][k], rtol=1e-14, atol=0)
    with pytest.raises(ValueError):
        outer_loop.resume_from(sim, straight[0]["f"][0], straight[0]["e"])


def test_cuda_graph_survives_growth_of_the_library_scratch(dev):
    """a captured step keeps the pointers of the library's reduction scratch (x-mode partials, density
    partials).  A larger problem in the same process makes that scratch grow; outgrown blocks are retired, not freed,
    so replaying the earlier graph stays correct -- and the inner loop re-captures when the generation changed."""
    from vlapy_b200 import ops, outer_loop
    cfg = O.nlepw_config(nx=32, nv=256, k0=0.35, log_nu=-2)
    params = make_params(cfg, "leapfrog", "lb")
    params["backend"]["cuda_graph"] = True
    stuff = make_stuff(cfg, RULES)
    gs = outer_loop._GraphStep(params, stuff, 32, 256, (2, 256), True, dev)
    e0, f0 = torch.from_numpy(cfg["e0"]).to(dev), torch.from_numpy(cfg["f0"]).to(dev)
    gs.e.copy_(e0); gs.f.copy_(f0)
    gs.capture()

    def replay():
        gs.e.copy_(e0); gs.f.copy_(f0)
        gs.inp.zero_()
        gs.graph.replay()
        torch.cuda.synchronize()
        return gs.f.clone(), gs.stage.clone()
    f_a, st_a = replay()
    gen = ops.scratch_generation()
    big = torch.rand((2048, 4096), dtype=torch.float64, device=dev)          # x-modes + fused density at a larger size
    ops.xmodes(big, 2)
    kx = torch.from_numpy(O.spatial_grid(0.0, 18.0, 2048)[2]).to(dev)
    vv = torch.linspace(-6.4, 6.4, 4096, dtype=torch.float64, device=dev)
    ops.vdfdx_exp(big, kx, vv, 0.1, flags=ops.PHASE_TABLE, density_out=torch.empty(2048, dtype=torch.float64, device=dev), dv=0.1)
    torch.cuda.synchronize()
    if ops.scratch_generation() > gen:       # (no growth when a larger test already ran in this process: run alone to see it)
        assert gs.scratch_generation != ops.scratch_generation()             # the inner loop would capture again
    f_b, st_b = replay()
    assert torch.equal(f_a, f_b) and torch.equal(st_a, st_b)
    e_ref, f_ref = O.run_steps(cfg, 1, "leapfrog", "lb")
    assert rel_err(f_b.cpu().numpy(), f_ref) < TOL


@pytest.mark.parametrize("graph", [True, False])
def test_run_loops_with_two_pinned_sets_equals_sequential_calls(dev, graph):
    """outer_loop.run_loops (storage hand-off overlapped with the next inner loop, backend.pinned_sets = 2): what
    the consumer sees for every batch, and the final state, equal a plain sequence of inner-loop calls"""
    from vlapy_b200 import outer_loop
    cfg = O.nlepw_config(nx=32, nv=256, k0=0.35, log_nu=-2)
    nloop, nt = 4, 6
    batches = []
    for li in range(nloop):
        t = cfg["dt"] * np.arange(li * nt, (li + 1) * nt)
        batches.append((t, np.stack([cfg["driver_function"](ti) for ti in t])))

    def grab(sim):
        return {"f": np.array(sim["f"]), "e": np.array(sim["e"]), "n": np.array(sim["fields"]["n"]),
                "T": np.array(sim["series"]["mean_T"]), "cum": np.array(sim["series"]["mean_cum_de2"]),
                "stored_f": np.array(sim["stored_f"]), "t": np.array(sim["time_batch"])}

    params = make_params(cfg, "leapfrog", "lb")
    params["backend"]["cuda_graph"] = graph
    sim, inner = outer_loop.get_sim_config_and_inner_loop_step(params, make_stuff(cfg, RULES), nt, RULES)
    seq = []
    for t, drv in batches:
        sim = inner(time_array=t, driver_array=drv, temp_storage=sim)
        seq.append(grab(sim))

    params = make_params(cfg, "leapfrog", "lb")
    params["backend"]["cuda_graph"] = graph
    params["backend"]["pinned_sets"] = 2
    sim2, inner2 = outer_loop.get_sim_config_and_inner_loop_step(params, make_stuff(cfg, RULES), nt, RULES)
    seen = []

    def consume(snap):
        import time
        first = grab(snap)
        time.sleep(0.05)                       # the next batch runs (and fills the OTHER pinned set) meanwhile
        again = grab(snap)
        assert all(np.array_equal(first[k], again[k]) for k in first)
        seen.append(first)

    final = outer_loop.run_loops(inner2, sim2, batches, consume)
    assert len(seen) == nloop
    for a, b in zip(seq, seen):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(np.array(final["f"]), seq[-1]["f"])
