"""world_size-2 (and 4) CPU tests of the multi-GPU orchestration (vlapy_b200/dist.py) over gloo:
the sharding, all-to-all transposes, partial density + all-reduce, sharded stored quantities.
The per-shard operators come from the oracle here (the CUDA operators have their own parity
tests); what is being tested is everything *around* them."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vpfp_oracle as O  # noqa: E402


class OracleBackend:
    """per-shard operators on CPU tensors, numpy/scipy arithmetic of the oracle"""

    def __init__(self, topo, cfg, fp_type):
        self.topo, self.cfg, self.fp_type = topo, cfg, fp_type
        self.v_loc = cfg["v"][topo.v0: topo.v0 + topo.nvl]

    def edfdv(self, fx, e_loc, dt):
        return torch.from_numpy(O.edfdv_exponential(fx.numpy(), e_loc.numpy(), dt, self.cfg["kv"]))

    def vdfdx(self, fv, dt):
        return torch.from_numpy(O.vdfdx_exponential(fv.numpy(), dt, self.cfg["kx"], self.v_loc))

    def density_partial(self, fv):
        w = np.full(self.topo.nvl, self.cfg["dv"])
        if self.topo.rank == 0:
            w[0] *= 0.5
        if self.topo.rank == self.topo.world - 1:
            w[-1] *= 0.5
        return torch.from_numpy((fv.numpy() * w).sum(axis=1))

    def poisson(self, n, driver):
        return torch.from_numpy(driver.numpy() + O.solve_for_field(n.numpy(), self.cfg["one_over_kx"]))

    def fp(self, fx, moments_out):
        c = self.cfg
        out = O.collision_step(fx.numpy(), c["v"], c["nu"], c["dt"], c["dv"], self.fp_type)
        if moments_out is not None:
            self.moments(torch.from_numpy(out), moments_out)
        return torch.from_numpy(np.ascontiguousarray(out))

    def moments(self, fx, out):
        c = self.cfg
        f = fx.numpy()
        out[:6] = torch.from_numpy(O.field_moments(f, c["v"], c["dv"]))
        out[6] = torch.from_numpy(O.trapz_last(f ** 2, c["dv"]))
        with np.errstate(invalid="ignore", divide="ignore"):
            out[7] = torch.from_numpy(O.trapz_last(f * np.log(f), c["dv"]))
        return out

    def driver(self, t):
        return torch.from_numpy(self.cfg["driver_function"](t))

    def xmodes_partial(self, fx, nmodes):
        x = np.arange(self.topo.x0, self.topo.x0 + self.topo.nxl)
        m = np.stack([(fx.numpy() * np.exp(-2j * np.pi * k * x / self.topo.nx)[:, None]).sum(0) for k in range(nmodes)])
        return torch.view_as_real(torch.from_numpy(m)).contiguous()

    def zeros(self, shape):
        return torch.zeros(shape, dtype=torch.float64)

    def upload(self, a):
        return a if isinstance(a, torch.Tensor) else torch.from_numpy(a)


def _worker(rank, world, port, integ, fp_type, nsteps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vlapy_b200 import dist as vd
        cfg = O.nlepw_config(nx=16, nv=32, k0=0.35, log_nu=-2)
        topo = vd.Topology(cfg["nx"], cfg["nv"])
        backend = OracleBackend(topo, cfg, fp_type)
        params = {"nu": cfg["nu"], "vlasov-poisson": {"time": integ}, "fokker-planck": {"type": fp_type}}
        stuff = {"dt": cfg["dt"]}
        step = vd.get_sharded_timestep(params, stuff, topo, backend=backend)
        f0 = torch.from_numpy(cfg["f0"][topo.x0: topo.x0 + topo.nxl].copy())
        state = {"e": torch.from_numpy(cfg["e0"].copy()), "f": vd.Sharded(f0, "x")}
        store = vd.make_store(topo, backend, nsteps)
        for i in range(nsteps):
            t = cfg["dt"] * i
            state = step(state, t, backend.driver(t), store)
        series, modes = vd.finish_store(topo, store)
        fx = vd.ops_to_x(state["f"], topo)
        # layout round trip is the identity
        rt = topo.v_to_x(topo.x_to_v(fx.clone()))
        q.put((rank, fx.numpy(), state["e"].numpy(), series.numpy(), modes.numpy(),
               store["fields_mom"].numpy(), store["fields_e"].numpy(), bool(torch.equal(rt, fx))))
    finally:
        dist.destroy_process_group()


def run_case(world, integ, fp_type, nsteps=4, port=29541):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, integ, fp_type, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return outs


@pytest.mark.parametrize("world,integ,fp_type", [(2, "leapfrog", "lb"), (2, "pefrl", "dg"), (2, "h-sixth", "lb"),
                                                 (4, "leapfrog", "dg")])
def test_sharded_step_matches_single_process(world, integ, fp_type):
    nsteps = 4
    outs = run_case(world, integ, fp_type, nsteps, port=29541 + world + len(integ))
    cfg = O.nlepw_config(nx=16, nv=32, k0=0.35, log_nu=-2)
    e_ref, f_ref, hist = O.run_steps(cfg, nsteps, integ, fp_type, collect=True)
    f = np.concatenate([o[1] for o in outs], axis=0)
    assert np.max(np.abs(f - f_ref)) / np.max(np.abs(f_ref)) < 1e-13
    for o in outs:
        assert np.max(np.abs(o[2] - e_ref)) < 1e-13            # e is replicated on every rank
        np.testing.assert_allclose(o[3], hist["series"], rtol=1e-11, atol=1e-15)   # series means after all-reduce
        assert o[7]                                            # x->v->x transposes are the identity
    mom = np.concatenate([o[5] for o in outs], axis=2)         # (steps, 6, nx) from the x-slabs
    for pw in range(6):     # a v^p moment amplifies rounding differences of f by int |v|^p dv
        assert np.max(np.abs(mom[:, pw] - hist["mom"][:, pw])) < 1e-13 * 2 * 6.4 ** (pw + 1) / (pw + 1), pw
    e_hist = np.concatenate([o[6] for o in outs], axis=1)
    assert np.max(np.abs(e_hist - hist["e"])) < 1e-13
    # stored x-modes: sum of the slab partials == fft_x(f)[:2] of the final state
    np.testing.assert_allclose(outs[0][4][-1], np.fft.fft(f_ref, axis=0)[:2], rtol=1e-11, atol=1e-13)


def test_topology_rejects_uneven_shards():
    from vlapy_b200 import dist as vd
    with pytest.raises(NotImplementedError):
        vd.Topology(10, 32, rank=0, world=4)
    with pytest.raises(NotImplementedError):
        vd.Topology(16, 6, rank=0, world=2)      # 3 columns per rank: odd


# ---------------------------------------------------------------------------------------------
# the sharded inner loop BEHIND THE REFERENCE API (vlapy_b200.outer_loop.get_sim_config_and_inner_loop_step
# under an initialised process group), against the reference's own outer_loop outputs (golden nlepw_c2 small_*)
# ---------------------------------------------------------------------------------------------
RULES = {"time": "first-last", "space": ["k0", "k1"]}


def _inner_loop_worker(rank, world, port, gather, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vlapy_b200 import outer_loop
        cfg = O.nlepw_config(nx=16, nv=128, k0=0.35, log_nu=-2)
        params = {"backend": {"core": "b200", "gather": gather,
                              "shard_backend": lambda topo, stuff, fp: OracleBackend(topo, cfg, fp)},
                  "nu": cfg["nu"],
                  "vlasov-poisson": {"time": "leapfrog", "vdfdx": "exponential", "edfdv": "exponential",
                                     "poisson": "spectral"},
                  "fokker-planck": {"type": "lb", "solver": "batched_tridiagonal"}}
        stuff = {k: cfg[k] for k in ("kx", "x", "one_over_kx", "v", "kv", "nv", "nx", "dv", "dt", "nu",
                                     "driver_function")}
        stuff.update(e=cfg["e0"], f=cfg["f0"], rules_to_store_f=RULES)
        nt = 24
        sim, inner = outer_loop.get_sim_config_and_inner_loop_step(params, stuff, nt, RULES)
        assert inner.topology.world == world
        outs = []
        for li in range(2):
            t = cfg["dt"] * np.arange(li * nt, (li + 1) * nt)
            drv = np.stack([cfg["driver_function"](ti) for ti in t])
            sim = inner(time_array=t, driver_array=drv, temp_storage=sim)
            outs.append({"fields": {k: np.array(v) for k, v in sim["fields"].items()},
                         "series": {k: np.array(v) for k, v in sim["series"].items()},
                         "stored_f": np.array(sim["stored_f"]), "f": np.array(sim["f"]), "e": np.array(sim["e"]),
                         "f_slab": sim["f_slab"]})
        q.put((rank, outs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("gather", ["rank0", "slab"])
def test_sharded_inner_loop_behind_the_reference_api(gather):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29571 + (gather == "slab")
    procs = [ctx.Process(target=_inner_loop_worker, args=(r, world, port, gather, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "nlepw_c2.npz"))
    nx = 16
    for li in range(2):
        for rank in range(world):
            o = res[rank][li]
            for k in ("e", "driver", "n", "j", "T", "q", "fv4", "vN"):          # every rank holds the global rows
                ref = g["small_fields_%s_%d" % (k, li)]
                assert o["fields"][k].shape == ref.shape
                # a v^p moment amplifies rounding differences of f by int |v|^p dv
                pw = {"n": 0, "j": 1, "T": 2, "q": 3, "fv4": 4, "vN": 5}.get(k, 0)
                assert np.max(np.abs(o["fields"][k] - ref)) < 1e-13 * 2 * 6.4 ** (pw + 1) / (pw + 1), (k, li)
            for k in O.SERIES_KEYS + ("mean_cum_de2", "mean_t_plus_e2_minus_cum_de2", "mean_t_plus_e2_plus_cum_de2"):
                np.testing.assert_allclose(o["series"][k], g["small_series_%s_%d" % (k, li)], rtol=1e-9, atol=1e-13,
                                           err_msg=k)
            assert o["stored_f"].dtype == np.complex64
            ref = g["small_stored_f_%d" % li]
            assert np.max(np.abs(o["stored_f"] - ref)) < 1e-6 * np.max(np.abs(ref))
            assert np.max(np.abs(o["e"] - g["small_e_%d" % li])) < 1e-12
        fref = g["small_f_%d" % li]
        if gather == "rank0":
            assert res[0][li]["f_slab"] == (0, nx)
            assert np.max(np.abs(res[0][li]["f"] - fref)) / np.max(np.abs(fref)) < 1e-12
            assert res[1][li]["f_slab"] == (nx // 2, nx)
        f = np.concatenate([res[0][li]["f"][:nx // 2] if gather == "rank0" else res[0][li]["f"], res[1][li]["f"]])
        assert np.max(np.abs(f - fref)) / np.max(np.abs(fref)) < 1e-12
