"""Mirror of the backend-dependent half of vlapy/outer_loop.py for ``backend.core == "b200"``.

``get_sim_config_and_inner_loop_step`` (vlapy/outer_loop.py:31-64) is the plug point: it returns the
storage dictionary and the inner-loop stepper.  Inside the inner loop everything is device
resident; on return the dictionary holds HOST numpy arrays under the keys the reference's
storage layer reads (vlapy/storage.py:78-91: time_batch, series, fields, stored_f, f, e), while
the device copies of e and f are kept under private keys for the next call.
"""
import numpy as np
import torch

from . import ops
from .core import step
from ._util import const, device

BACKEND = "b200"


def get_sim_config_and_inner_loop_step(all_params, stuff_for_time_loop, nt_in_loop, store_f_rules):
    """vlapy/outer_loop.py:31-64."""
    if all_params["backend"]["core"] != BACKEND:
        raise NotImplementedError(
            "The backend <" + all_params["backend"]["core"] + "> has not yet been implemented")
    if sharded_world(all_params) > 1:
        # launched under torchrun (one process per GPU): the same call returns the x/v-sharded inner loop
        do_inner_loop = get_sharded_inner_loop_stepper(all_params, stuff_for_time_loop, nt_in_loop)
    else:
        do_inner_loop = get_inner_loop_stepper(all_params, stuff_for_time_loop, nt_in_loop)
    sim_config = get_arrays_for_inner_loop(stuff_for_time_loop, nt_in_loop, store_f_rules, this_np=np)
    return sim_config, do_inner_loop


def get_arrays_for_inner_loop(stuff_for_time_loop, nt_in_loop, store_f_rules, this_np=np):
    """vlapy/outer_loop.py:149-215: same keys and shapes; device buffers are created lazily by the
    inner loop (``_dev``), the host-visible entries start as the reference initialises them."""
    f0 = np.asarray(stuff_for_time_loop["f"])
    e0 = np.asarray(stuff_for_time_loop["e"])
    if store_f_rules["space"] == "all":
        store_f = np.zeros((nt_in_loop,) + f0.shape, dtype=np.float64)
        store_f[0] = f0
    elif isinstance(store_f_rules["space"], list) and store_f_rules["space"][0] == "k0":
        store_f = np.zeros((nt_in_loop, len(store_f_rules["space"]), f0.shape[1]), dtype=np.complex64)
        store_f[0] = np.fft.fft(f0, axis=0)[: len(store_f_rules["space"])]
    else:
        raise NotImplementedError
    return {
        "time_batch": np.zeros(nt_in_loop),
        "e": np.array(e0),
        "f": np.array(f0),
        "stored_f": store_f,
        "mean_cum_de2_previous": 0.0,
        "series": {k: np.zeros(nt_in_loop) for k in step.SERIES_KEYS},
        "fields": {k: np.zeros((nt_in_loop,) + e0.shape) for k in step.FIELD_KEYS},
    }


def post_inner_loop_update(temp_storage, this_np=np):
    """vlapy/outer_loop.py:218-245 (host side, O(nt_in_loop))."""
    s = temp_storage["series"]
    s["mean_cum_de2"] = temp_storage["mean_cum_de2_previous"] + this_np.cumsum(s["mean_de2"])
    s["mean_t_plus_e2_minus_cum_de2"] = s["mean_T"] + (s["mean_e2"] - s["mean_cum_de2"])
    s["mean_t_plus_e2_plus_cum_de2"] = s["mean_T"] + s["mean_e2"] + s["mean_de2"]
    temp_storage["mean_cum_de2_previous"] = s["mean_cum_de2"][-1]
    return temp_storage


def _device_storage(temp_storage, nt, nx, nv, dev, pinned_sets=1):
    """Device twins of the per-loop buffers of vlapy/outer_loop.py:190-215 (key ``_dev``) and pinned
    host mirrors for the once-per-loop download (key ``_pin``).  The device copies of e and f are
    authoritative between calls; delete ``temp_storage["_dev"]`` to make the next call upload
    ``temp_storage["e"]`` / ``["f"]`` again.  ``pinned_sets = 2``: two sets of pinned mirrors used
    alternately, so that the arrays one call returns stay valid while the next call runs
    (``run_loops``)."""
    d = temp_storage.get("_dev")
    if d is None or d["nt"] != nt:
        stored = temp_storage["stored_f"]
        cplx = np.iscomplexobj(stored)
        d = {
            "nt": nt,
            "fields": torch.zeros((len(step.FIELD_KEYS), nt, nx), dtype=torch.float64, device=dev),
            "series_rows": torch.zeros((nt, 7), dtype=torch.float64, device=dev),
            "stored_f": torch.zeros(stored.shape, dtype=torch.complex128 if cplx else torch.float64, device=dev),
            "moments": torch.zeros((8, nx), dtype=torch.float64, device=dev),
            "e": torch.from_numpy(np.ascontiguousarray(temp_storage["e"], dtype=np.float64)).to(dev),
            "f": torch.from_numpy(np.ascontiguousarray(temp_storage["f"], dtype=np.float64)).to(dev),
        }
        temp_storage["_dev"] = d
    p = temp_storage.get("_pin")
    if p is None or p["nt"] != nt or len(p["sets"]) != pinned_sets:
        sf = d["stored_f"]
        pin_it = dev.type == "cuda"
        p = {"nt": nt, "next": 0, "sets": [{
            "fields": torch.empty(d["fields"].shape, dtype=torch.float64, pin_memory=pin_it),
            "series_rows": torch.empty((nt, 7), dtype=torch.float64, pin_memory=pin_it),
            "stored_f": torch.empty(sf.shape, dtype=torch.complex64 if sf.is_complex() else torch.float64,
                                    pin_memory=pin_it),
            "e": torch.empty(nx, dtype=torch.float64, pin_memory=pin_it),
            "f": torch.empty((nx, nv), dtype=torch.float64, pin_memory=pin_it),
        } for _ in range(pinned_sets)]}
        temp_storage["_pin"] = p
    pin = p["sets"][p["next"]]
    p["next"] = (p["next"] + 1) % len(p["sets"])
    return d, pin


class _GraphStep:
    """One timestep captured in a CUDA graph.  Small grids are launch bound: a step
    is ~25 kernels of a few microseconds, so the Python/launch overhead dominates.  The captured
    step reads its time and driver row from a small device input buffer and writes everything the
    storage step records into one staging row, so a replay costs three launches from Python
    (input copy, graph, staging copy).  The state ``e`` and the moment rows ARE slices of the staging row and the x-mode
    kernel writes its slice directly: no copies between them inside the graph (each was a launch of its own, 2-3 us of a
    60-90 us step); the driver rows are not staged at all, the caller has them.  Needs the device-side driver
    (``pulse_dictionary``)."""

    def __init__(self, all_params, stuff, nx, nv, nmodes_shape, modes_complex, dev):
        self.vp_step, self.fp_step, self.fused, self.store_f = step.get_step_parts(all_params, stuff)
        self.v_d, self.dv = const(stuff["v"]), float(stuff["dv"])
        self.nx, self.nv = nx, nv
        self.inp = torch.zeros(1 + nx, dtype=torch.float64, device=dev)            # [t, driver row]
        self.f = torch.zeros((nx, nv), dtype=torch.float64, device=dev)
        self.n_modes = int(np.prod(nmodes_shape))
        self.modes_complex = bool(modes_complex)
        # staging row: [e (nx) | eight moment rows (8 nx) | seven series entries + 1 pad | stored modes]
        nstage = 9 * nx + 8 + (2 * self.n_modes if self.modes_complex else self.n_modes)
        self.stage = torch.zeros(nstage, dtype=torch.float64, device=dev)
        self.e = self.stage[0:nx]
        self.mom = self.stage[nx:9 * nx].view(8, nx)
        rule = stuff["rules_to_store_f"]
        self.xmodes_direct = self.modes_complex and rule["space"][0] == "k0" and self.n_modes % nv == 0
        self.graph = None

    def _body(self):
        nx = self.nx
        t = ops.DevTime(self.inp[:1])
        de = self.inp[1:]
        e, f = self.vp_step(e=self.e, f=self.f, t=t)
        if self.fused:
            f = self.fp_step(f, moments_out=self.mom, out=self.f)      # straight into the state buffer (f is a new tensor)
        else:
            f = self.fp_step(f=f)
            ops.moments(f, self.v_d, self.dv, nmom=8, out=self.mom)
        st = self.stage
        ops.series(self.mom, e, de, out=st[9 * nx:9 * nx + 7])
        tail = st[9 * nx + 8:]
        if self.xmodes_direct:
            ops.xmodes(f, self.n_modes // self.nv, out=tail.view(1, self.n_modes // self.nv, self.nv, 2))
        else:
            m = self.store_f(f)
            tail.copy_(torch.view_as_real(m).reshape(-1) if self.modes_complex else m.reshape(-1))
        self.e.copy_(e)                                  # = the e slice of the staging row
        if f.data_ptr() != self.f.data_ptr():
            self.f.copy_(f)

    def capture(self):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        e0, f0 = self.e.clone(), self.f.clone()
        with torch.cuda.stream(side):
            self._body()                        # warm-up: library caches, allocator pools
        torch.cuda.current_stream().wait_stream(side)
        self.e.copy_(e0)
        self.f.copy_(f0)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        self.e.copy_(e0)
        self.f.copy_(f0)
        # the library's reduction scratch the graph was captured with (outgrown blocks stay allocated, so the graph
        # remains valid; a changed generation only means that a re-capture would use the current block)
        self.scratch_generation = ops.scratch_generation()


GRAPH_MAX_CELLS = 1 << 22      # above this a step is no longer launch bound (and the extra copy of f would cost)


def get_inner_loop_stepper(all_params, stuff_for_time_loop, steps_in_loop):
    """vlapy/outer_loop.py:248-281: inner_loop(time_array, driver_array, temp_storage).

    ``all_params["backend"]["cuda_graph"]``: "auto" (default: graphs for grids up to 2^22 cells when
    the device-side driver is available), True or False."""
    if all_params["backend"]["core"] != BACKEND:
        raise NotImplementedError(
            "The backend: <" + all_params["backend"]["core"] + "> has not yet been implemented")
    one_step = step.get_timestep(all_params=all_params, stuff_for_time_loop=stuff_for_time_loop)
    graph_opt = all_params["backend"].get("cuda_graph", "auto")
    pinned_sets = int(all_params["backend"].get("pinned_sets", 1))      # 2: see run_loops
    if pinned_sets not in (1, 2):
        raise ValueError("backend.pinned_sets must be 1 or 2")
    graph_state = {}

    def want_graph(nx, nv):
        if graph_opt is False or stuff_for_time_loop.get("pulse_dictionary") is None:
            return False
        return graph_opt is True or nx * nv <= GRAPH_MAX_CELLS

    def run_graph(time_array, drv, d, nx, nv):
        gs = graph_state.get("gs")
        if gs is not None and gs.scratch_generation != ops.scratch_generation():
            gs = None                                   # a larger problem ran in between: capture again
        if gs is None:
            gs = _GraphStep(all_params, stuff_for_time_loop, nx, nv, tuple(d["stored_f"].shape[1:]),
                            d["stored_f"].is_complex(), drv.device)
            gs.e.copy_(d["e"]); gs.f.copy_(d["f"])
            gs.capture()
            graph_state["gs"] = gs
        gs.e.copy_(d["e"]); gs.f.copy_(d["f"])
        nt = steps_in_loop
        inputs = torch.empty((nt, 1 + nx), dtype=torch.float64, device=drv.device)
        inputs[:, 0] = torch.as_tensor(np.asarray(time_array, dtype=np.float64)).to(drv.device)
        inputs[:, 1:] = drv
        rows = torch.empty((nt, gs.stage.numel()), dtype=torch.float64, device=drv.device)
        for it in range(nt):
            gs.inp.copy_(inputs[it])
            gs.graph.replay()
            rows[it].copy_(gs.stage)
        d["e"], d["f"] = gs.e.clone(), gs.f.clone()
        d["fields"][0].copy_(rows[:, :nx])                                                  # FIELD_KEYS: e, driver, six moments
        d["fields"][1].copy_(drv)
        d["fields"][2:8].copy_(rows[:, nx:7 * nx].reshape(nt, 6, nx).permute(1, 0, 2))
        d["series_rows"].copy_(rows[:, 9 * nx:9 * nx + 7])
        tail = rows[:, 9 * nx + 8:]
        if gs.modes_complex:
            d["stored_f"].copy_(torch.view_as_complex(tail.reshape(nt, -1, 2).contiguous()).reshape(d["stored_f"].shape))
        else:
            d["stored_f"].copy_(tail.reshape(d["stored_f"].shape))

    def inner_loop(time_array, driver_array, temp_storage):
        dev = device()
        nx, nv = np.asarray(temp_storage["f"]).shape[-2:]
        d, pin = _device_storage(temp_storage, steps_in_loop, nx, nv, dev, pinned_sets)
        # host -> device: the driver rows of this loop (asynchronous when the caller pinned them)
        drv = torch.as_tensor(np.ascontiguousarray(driver_array, dtype=np.float64)).to(dev, non_blocking=True)
        if want_graph(nx, nv):
            run_graph(time_array, drv, d, nx, nv)
        else:
            work = {
                "time_batch": np.asarray(time_array, dtype=np.float64),
                "driver_array_batch": drv,
                "e": d["e"], "f": d["f"],
                "stored_f": d["stored_f"],
                "fields": dict({k: d["fields"][j] for j, k in enumerate(step.FIELD_KEYS)}, _block=d["fields"]),
                "series": {"_rows": d["series_rows"]},
                "_moment_scratch": d["moments"],
            }
            for it in range(steps_in_loop):
                work, _ = one_step(work, it)
            d["e"], d["f"] = work["e"], work["f"]
        # device -> host once per inner loop (the storage cadence of vlapy/manager.py:138-150),
        # into pinned mirrors; the returned arrays are views that the next call overwrites, as the
        # reference's in-place temp_storage arrays are.
        pin["fields"].copy_(d["fields"], non_blocking=True)
        pin["series_rows"].copy_(d["series_rows"], non_blocking=True)
        sf = d["stored_f"]
        pin["stored_f"].copy_(sf.to(torch.complex64) if sf.is_complex() else sf, non_blocking=True)
        pin["e"].copy_(d["e"], non_blocking=True)
        pin["f"].copy_(d["f"], non_blocking=True)
        if dev.type == "cuda":
            torch.cuda.current_stream().synchronize()
        temp_storage["time_batch"] = np.asarray(time_array)
        temp_storage["driver_array_batch"] = np.asarray(driver_array)
        fields = pin["fields"].numpy()
        for j, k in enumerate(step.FIELD_KEYS):
            temp_storage["fields"][k] = fields[j]
        rows = pin["series_rows"].numpy()
        for j, k in enumerate(step.SERIES_KEYS):
            temp_storage["series"][k] = rows[:, j].copy()
        temp_storage["stored_f"] = pin["stored_f"].numpy()
        temp_storage["e"] = pin["e"].numpy()
        temp_storage["f"] = pin["f"].numpy()
        post_inner_loop_update(temp_storage, np)
        return temp_storage

    return inner_loop



def resume_from(temp_storage, f, e, mean_cum_de2_previous=0.0):
    """Restart hook (the reference leaves it as a TODO at vlapy/manager.py:118-119): continue a
    simulation from a stored state, e.g. ``f`` = the last ``full_distribution`` record that vlapy/storage.py:211-231
    wrote (shape (nx, nv), or this rank's x-slab under the sharded inner loop) and ``e`` = the matching field row.
    The device-resident copies are dropped, so the next ``inner_loop`` call uploads these arrays; the cumulative
    driver energy of vlapy/outer_loop.py:218-245 restarts from ``mean_cum_de2_previous``.  The caller starts the
    manager's loop at the step index of the record, so that ``time_array`` / ``driver_array`` continue from there."""
    f = np.array(f, dtype=np.float64, order="C")
    e = np.array(e, dtype=np.float64, order="C")
    if f.ndim != 2 or e.ndim != 1 or e.shape[0] < f.shape[0]:
        raise ValueError("resume_from: f must be (nx, nv) or an x-slab of it, e the full field (nx)")
    temp_storage.pop("_dev", None)
    temp_storage["f"], temp_storage["e"] = f, e
    temp_storage["mean_cum_de2_previous"] = float(mean_cum_de2_previous)
    return temp_storage


def sharded_world(all_params):
    """Number of ranks the inner loop is sharded over: the size of the default process group when
    torch.distributed is initialised (torchrun, one process per GPU) and ``backend.sharded`` is not False."""
    import torch.distributed as tdist
    if all_params["backend"].get("sharded", "auto") is False:
        return 1
    if not (tdist.is_available() and tdist.is_initialized()):
        return 1
    return tdist.get_world_size()


def get_sharded_inner_loop_stepper(all_params, stuff_for_time_loop, steps_in_loop):
    """vlapy/outer_loop.py:248-281 for one simulation sharded over the ranks of the default process group
    (rows x-sharded, v df/dx v-sharded, vlapy_b200/dist.py).  SPMD: every rank makes the same call
    with the same arguments (the reference's setup is deterministic, so every rank holds the same host arrays);
    a rank uploads only its x-slab of ``temp_storage["f"]``.  On return every rank holds the same host
    dictionary as the single-GPU inner loop returns -- time_batch, driver_array_batch, series, fields,
    stored_f, e (the per-loop buffers are all-gathered / all-reduced on the device once per loop, at the
    storage cadence of vlapy/manager.py:138-150) -- except ``f``:

      backend.gather = "rank0" (default): rank 0 receives the full f (nx, nv), as the storage layer of the
                       reference expects; the other ranks receive their x-slab;
      backend.gather = "slab":  every rank receives its x-slab only (parallel downloads, one PCIe link each).

    ``temp_storage["f_slab"] = (x_begin, x_end)`` says which rows ``f`` holds.  The device-resident slab is
    kept under the private key for the next call.  ``backend.shard_backend`` (tests): a factory
    (topology, stuff, fp_type) -> per-shard operators replacing the CUDA kernels (dist.DeviceBackend)."""
    import torch.distributed as tdist
    from . import dist as vd
    if all_params["backend"]["core"] != BACKEND:
        raise NotImplementedError(
            "The backend: <" + all_params["backend"]["core"] + "> has not yet been implemented")
    nx, nv = int(stuff_for_time_loop["nx"]), int(stuff_for_time_loop["nv"])
    topo = vd.Topology(nx, nv)
    factory = all_params["backend"].get("shard_backend")
    fp_type = all_params["fokker-planck"]["type"]
    backend = factory(topo, stuff_for_time_loop, fp_type) if factory else vd.DeviceBackend(
        topo, stuff_for_time_loop, fp_type, peer_scatter=all_params["backend"].get("peer_scatter", True))
    one_step = vd.get_sharded_timestep(all_params, stuff_for_time_loop, topo, backend=backend)
    gather = all_params["backend"].get("gather", "rank0")
    if gather not in ("rank0", "slab"):
        raise ValueError("backend.gather must be 'rank0' or 'slab'")
    nmodes = len(stuff_for_time_loop["rules_to_store_f"])       # the reference's quirk (vlapy/core/step.py:133)
    if stuff_for_time_loop["rules_to_store_f"]["space"] == "all":
        raise NotImplementedError("rules_to_store_f space = 'all' has not been implemented for the sharded inner loop")
    pins = {}

    def pinned(name, shape, dtype=torch.float64):
        t = pins.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dtype, pin_memory=torch.cuda.is_available())
            pins[name] = t
        return t

    def gather_x(t):
        """(nt, ..., nxl) slabs -> (nt, ..., nx) on every rank"""
        parts = [torch.empty_like(t) for _ in range(topo.world)]
        tdist.all_gather(parts, t.contiguous(), group=topo.group)
        return torch.cat(parts, dim=-1)

    def inner_loop(time_array, driver_array, temp_storage):
        nt = steps_in_loop
        d = temp_storage.get("_dev")
        if d is None:
            f_host = np.asarray(temp_storage["f"], dtype=np.float64)
            if f_host.shape[0] == nx:
                f_host = f_host[topo.x0: topo.x0 + topo.nxl]
            elif f_host.shape[0] != topo.nxl:
                raise ValueError("temp_storage['f'] must be the full grid or this rank's x-slab")
            d = {"f": vd.Sharded(backend.upload(np.ascontiguousarray(f_host)), "x"),
                 "e": backend.upload(np.ascontiguousarray(temp_storage["e"], dtype=np.float64))}
            temp_storage["_dev"] = d
        drv = backend.upload(np.ascontiguousarray(driver_array, dtype=np.float64))
        store = vd.make_store(topo, backend, nt, nmodes)
        state = {"e": d["e"], "f": d["f"]}
        times = np.asarray(time_array, dtype=np.float64)
        for it in range(nt):
            state = one_step(state, float(times[it]), drv[it], store)
        d["e"], d["f"] = state["e"], vd.Sharded(vd.ops_to_x(state["f"], topo), "x")
        # storage cadence: per-loop buffers become global on the device, then one download
        series, modes = vd.finish_store(topo, store)
        fe = gather_x(store["fields_e"])
        fd = gather_x(store["fields_driver"])
        fm = gather_x(store["fields_mom"])
        fx = d["f"].t
        if gather == "rank0" and topo.world > 1:
            parts = [torch.empty_like(fx) for _ in range(topo.world)] if topo.rank == 0 else None
            tdist.gather(fx.contiguous(), parts, dst=0, group=topo.group)
            f_dev = torch.cat(parts, dim=0) if topo.rank == 0 else fx
        else:
            f_dev = fx
        outs = {"fe": fe, "fd": fd, "fm": fm, "series": series, "e": d["e"], "f": f_dev,
                "modes": torch.view_as_real(modes.to(torch.complex64)) if modes.is_complex() else modes}
        host = {}
        for k, t in outs.items():
            h = pinned(k, t.shape, t.dtype)
            h.copy_(t, non_blocking=True)
            host[k] = h
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        temp_storage["time_batch"] = np.asarray(time_array)
        temp_storage["driver_array_batch"] = np.asarray(driver_array)
        temp_storage["fields"]["e"] = host["fe"].numpy()
        temp_storage["fields"]["driver"] = host["fd"].numpy()
        fm_h = host["fm"].numpy()
        for j, k in enumerate(("n", "j", "T", "q", "fv4", "vN")):
            temp_storage["fields"][k] = fm_h[:, j]
        rows = host["series"].numpy()
        for j, k in enumerate(step.SERIES_KEYS):
            temp_storage["series"][k] = rows[:, j].copy()
        temp_storage["stored_f"] = torch.view_as_complex(host["modes"]).numpy()
        temp_storage["e"] = host["e"].numpy()
        temp_storage["f"] = host["f"].numpy()
        full = f_dev.shape[0] == nx
        temp_storage["f_slab"] = (0, nx) if full else (topo.x0, topo.x0 + topo.nxl)
        post_inner_loop_update(temp_storage, np)
        return temp_storage

    inner_loop.topology, inner_loop.shard_backend = topo, backend
    return inner_loop


def run_loops(inner_loop, temp_storage, batches, consume):
    """The loop of vlapy/manager.py:118-150 with the storage hand-off overlapped.

    ``batches`` yields ``(time_array, driver_array)``; for every batch ``consume(snapshot)`` -- the reference's
    ``storage_manager.batch_update`` -- is called once, in order, on a worker thread WHILE the next batch computes.
    ``snapshot`` is a shallow copy of the dictionary the inner loop returned (time_batch, series, fields, stored_f,
    e, f as host arrays).  The inner loop must have been built with ``all_params["backend"]["pinned_sets"] = 2``:
    the arrays of batch i live in pinned set i % 2, which batch i+2 overwrites, so batch i+1 is not started before
    ``consume`` of batch i-1 has returned.  Exceptions of ``consume`` are re-raised here.  Returns the final
    dictionary (device-resident e, f under its private keys, as after a plain sequence of calls)."""
    import collections
    from concurrent.futures import ThreadPoolExecutor
    pending = collections.deque()
    with ThreadPoolExecutor(max_workers=1) as pool:
        for time_array, driver_array in batches:
            while len(pending) >= 2:                  # the pinned set about to be rewritten is still being read
                pending.popleft().result()
            temp_storage = inner_loop(time_array=time_array, driver_array=driver_array, temp_storage=temp_storage)
            snap = {k: (dict(v) if isinstance(v, dict) else v) for k, v in temp_storage.items()
                    if not k.startswith("_")}
            pending.append(pool.submit(consume, snap))
        while pending:
            pending.popleft().result()
    return temp_storage
