"""Multi-GPU parity (-m gpu, needs >= 2 devices): the x/v-sharded step over NCCL against the
reference output of the single-process run (golden nlepw_c2: C2, 40 collisional steps)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, op, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import vpfp_oracle as O
        from vlapy_b200 import dist as vd
        cfg = O.nlepw_config()
        topo = vd.Topology(cfg["nx"], cfg["nv"])
        params = {"nu": cfg["nu"], "vlasov-poisson": {"time": "leapfrog"}, "fokker-planck": {"type": op}}
        stuff = {k: cfg[k] for k in ("kx", "x", "one_over_kx", "v", "kv", "nv", "nx", "dv", "dt", "nu")}
        stuff.update(pulse_dictionary=cfg["pulses"], driver_function=cfg["driver_function"])
        step = vd.get_sharded_timestep(params, stuff, topo)
        dev = torch.device("cuda", rank)
        f0 = torch.from_numpy(cfg["f0"][topo.x0: topo.x0 + topo.nxl].copy()).to(dev)
        state = {"e": torch.from_numpy(cfg["e0"].copy()).to(dev), "f": vd.Sharded(f0, "x")}
        store = vd.make_store(topo, step.backend, 40)
        for i in range(40):
            t = cfg["dt"] * i
            state = step(state, t, step.backend.driver(t), store)
        series, modes = vd.finish_store(topo, store)
        fx = vd.ops_to_x(state["f"], topo)
        q.put((rank, fx.cpu().numpy(), state["e"].cpu().numpy(), series.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_c2_matches_reference(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "nlepw_c2.npz"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29620 + world, "lb", q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=600) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = np.concatenate([o[1] for o in outs], axis=0)
    e = outs[0][2]
    assert np.max(np.abs(e - g["e_lb"])) / np.max(np.abs(g["e_lb"])) < 1e-11
    assert np.max(np.abs(f[::8, ::16] - g["f_sub_lb"])) / np.max(np.abs(g["f_sub_lb"])) < 1e-12
    assert abs(f.sum() / float(g["f_sum_lb"]) - 1) < 1e-13
    from oracle import vpfp_oracle as O
    dv = 2 * 6.4 / 2048
    mean_n = O.trapz_last(f, dv).mean()                   # mean density of the assembled final state
    assert abs(outs[0][3][-1, 0] - mean_n) < 1e-12        # == the all-reduced series entry of the last step


def _worker_scatter(rank, world, port, nx, nv, nsteps, integrator, q):
    """the same steps twice: NCCL all-to-all transposes, then the layout change fused into the
    advection kernels' last pass (peer stores over NVLink)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from oracle import vpfp_oracle as O
        from vlapy_b200 import dist as vd
        cfg = O.nlepw_config(nx=nx, nv=nv, log_nu=-2)
        topo = vd.Topology(cfg["nx"], cfg["nv"])
        params = {"nu": cfg["nu"], "vlasov-poisson": {"time": integrator}, "fokker-planck": {"type": "lb"}}
        stuff = {k: cfg[k] for k in ("kx", "x", "one_over_kx", "v", "kv", "nv", "nx", "dv", "dt", "nu")}
        stuff.update(pulse_dictionary=cfg["pulses"], driver_function=cfg["driver_function"])
        dev = torch.device("cuda", rank)
        res = {}
        for mode in ("a2a", "scatter"):
            os.environ["VPFP_NO_SCATTER"] = "1" if mode == "a2a" else "0"
            step = vd.get_sharded_timestep(params, stuff, topo)
            assert step.backend.can_scatter == (mode == "scatter")
            f0 = torch.from_numpy(cfg["f0"][topo.x0: topo.x0 + topo.nxl].copy()).to(dev)
            state = {"e": torch.from_numpy(cfg["e0"].copy()).to(dev), "f": vd.Sharded(f0, "x")}
            store = vd.make_store(topo, step.backend, nsteps)
            for i in range(nsteps):
                t = cfg["dt"] * i
                state = step(state, t, step.backend.driver(t), store)
            series, modes = vd.finish_store(topo, store)
            fx = vd.ops_to_x(state["f"], topo)
            res[mode] = (fx.cpu().numpy(), state["e"].cpu().numpy(), series.cpu().numpy())
            step.backend.close()
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,nv,integrator", [(2, 2048, 4096, "leapfrog"), (2, 4096, 2048, "pefrl"),
                                                    (4, 4096, 4096, "leapfrog"), (8, 8192, 4096, "leapfrog")])
def test_peer_scatter_matches_all_to_all_and_oracle(world, nx, nv, integrator):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    nsteps = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_scatter, args=(r, world, 29640 + world, nx, nv, nsteps, integrator, q))
             for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=900) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    f = {m: np.concatenate([o[1][m][0] for o in outs], axis=0) for m in ("a2a", "scatter")}
    e = {m: outs[0][1][m][1] for m in ("a2a", "scatter")}
    # only the destination of the last pass' stores differs: bit-identical results
    assert np.array_equal(f["a2a"], f["scatter"])
    assert np.array_equal(e["a2a"], e["scatter"])
    assert np.array_equal(outs[0][1]["a2a"][2], outs[0][1]["scatter"][2])
    from oracle import vpfp_oracle as O
    cfg = O.nlepw_config(nx=nx, nv=nv, log_nu=-2)
    e_ref, f_ref = O.run_steps(cfg, nsteps, integrator, "lb")
    assert np.max(np.abs(f["scatter"] - f_ref)) / np.max(np.abs(f_ref)) < 1e-12
    assert np.max(np.abs(e["scatter"] - e_ref)) / np.max(np.abs(e_ref)) < 1e-10
