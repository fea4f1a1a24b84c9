"""Multi-GPU VPFP step: one process per GPU, f sharded along x for the row operators and along v
for the x-advection, torch.distributed all-to-all transposes between the two layouts.

The reference is single-process; this is the sharding of SURVEY 8(e):

    e df/dv, Fokker-Planck, v-moments   act on rows     -> x-sharded  f_x: (nx/P, nv)
    v df/dx                             acts on columns -> v-sharded  f_v: (nx, nv/P)
    density                             partial int over the local v-slice + all-reduce of nx doubles
    Poisson                             solved redundantly on every rank (nx values)

The orchestration is independent of where the operators run: ``backend`` supplies the per-shard
operators (``DeviceBackend`` = the CUDA kernels of this package; the CPU tests plug the oracle in
over gloo).  The operator closures built here accept and return ``Sharded`` handles, so the
schedule mirrors of ``vlapy_b200.core.vlasov_poisson`` compose them unchanged.
"""
import numpy as np
import torch
import torch.distributed as dist

from ._util import attach_density, cached_density


class Sharded:
    """A shard of f together with its layout: 'x' = rows [r*nx/P, (r+1)*nx/P) x all v,
    'v' = all x  x  columns [r*nv/P, (r+1)*nv/P).
    ``density``: the (globally reduced) charge density of this f when the producing operator already
    computed it.  ``pending`` = (f_x, e_local, dt): the shard is e df/dv of f_x, not evaluated yet --
    whoever consumes it decides where the result is stored (locally in x layout, or scattered over
    NVLink into the peers' v-shards)."""
    __slots__ = ("t", "layout", "density", "pending")

    def __init__(self, t, layout, density=None, pending=None):
        self.t, self.layout, self.density, self.pending = t, layout, density, pending


class Topology:
    def __init__(self, nx, nv, rank=None, world=None, group=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        P = self.world
        if nx % P or nv % P or (nv // P) % 2:
            raise NotImplementedError("sharding needs nx and nv divisible by the number of ranks "
                                      "(and an even number of v-columns per rank)")
        self.nx, self.nv, self.nxl, self.nvl = nx, nv, nx // P, nv // P
        self.x0, self.v0 = self.rank * self.nxl, self.rank * self.nvl

    # ---- layout changes -------------------------------------------------------------------
    def x_to_v(self, fx):
        """(nx/P, nv) -> (nx, nv/P): rank r sends its rows of column block q to rank q."""
        P, nxl, nvl = self.world, self.nxl, self.nvl
        if P == 1:
            return fx
        send = fx.view(nxl, P, nvl).permute(1, 0, 2).contiguous()          # block q first
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.view(P * nxl, nvl)                                      # blocks arrive ordered by source rank = x order

    def v_to_x(self, fv):
        """(nx, nv/P) -> (nx/P, nv): the row blocks of a v-shard are already contiguous."""
        P, nxl, nvl = self.world, self.nxl, self.nvl
        if P == 1:
            return fv
        send = fv.contiguous().view(P, nxl, nvl)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv.permute(1, 0, 2).reshape(nxl, P * nvl)

    def all_reduce_sum(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def stream_barrier(self, token):
        """order the ranks on the device (no host synchronisation): a one-element all-reduce"""
        return self.all_reduce_sum(token)


class PeerShards:
    """Receive buffers of this rank that the peers can address (CUDA IPC over NVLink), and the
    mapped pointers of every peer's: ``fx`` (nx/P, nv) receives the scattered result of v df/dx,
    ``fv`` (nx, nv/P) the scattered result of e df/dv."""

    def __init__(self, topo, device):
        from . import ops
        nbytes = 8 * topo.nxl * topo.nv
        self.bx, self.bv = ops.PeerBuffer(nbytes), ops.PeerBuffer(nbytes)
        self.fx = self.bx.tensor((topo.nxl, topo.nv), device)
        self.fv = self.bv.tensor((topo.nx, topo.nvl), device)
        mine = torch.tensor(list(self.bx.handle + self.bv.handle), dtype=torch.uint8, device=device)
        every = [torch.empty_like(mine) for _ in range(topo.world)]
        dist.all_gather(every, mine, group=topo.group)
        self.px, self.pv = [], []
        for q, h in enumerate(every):
            if q == topo.rank:
                self.px.append(self.bx.ptr); self.pv.append(self.bv.ptr)
            else:
                raw = bytes(h.cpu().numpy().tobytes())
                self.px.append(self.bx.open_peer(raw[:64])); self.pv.append(self.bv.open_peer(raw[64:]))
        self.token = torch.zeros(1, dtype=torch.float32, device=device)

    def close(self):
        torch.cuda.synchronize()
        self.fx = self.fv = None
        self.bx.close(); self.bv.close()


class DeviceBackend:
    """Per-shard operators on the CUDA kernels (vlapy_b200.ops)."""

    def __init__(self, topo, stuff, fp_type, peer_scatter=True):
        from . import ops
        from ._util import const
        from .core.vlasov import _phase_flags
        self.ops, self.topo = ops, topo
        s = stuff
        self.kx, self.kv = const(s["kx"]), const(s["kv"])
        self.v = const(s["v"])
        self.v_loc = self.v[topo.v0: topo.v0 + topo.nvl].contiguous()
        self.ook = const(s["one_over_kx"])
        self.x = const(s["x"])
        self.dv, self.dt, self.nu = float(s["dv"]), float(s["dt"]), float(s["nu"])
        self.fp_type = fp_type
        self.flags_x, self.flags_v = _phase_flags(s["kx"]), _phase_flags(s["kv"])
        self.vgrid = ops.linspace_params(s["v"])
        self.pulses = ops.pulses_to_array(s["pulse_dictionary"]) if s.get("pulse_dictionary") else None
        self.driver_host = s.get("driver_function")
        self.edge = (1 if topo.rank == 0 else 0) | (2 if topo.rank == topo.world - 1 else 0)
        # layout changes fused into the operators' last pass (stores over NVLink into the peers'
        # shards) when the register-resident kernels apply; otherwise NCCL all-to-all transposes
        def pow2(n):
            return n > 0 and (n & (n - 1)) == 0
        P = topo.world
        self.can_scatter = (P > 1 and P <= 8 and pow2(P) and pow2(topo.nx) and pow2(topo.nv)
                            and 256 <= topo.nx <= 16384 and 256 <= topo.nv <= 16384
                            and topo.nxl * topo.nv >= (1 << 22) and bool(peer_scatter))
        self.peers = PeerShards(topo, self.x.device) if self.can_scatter else None
        self.scratch_x = self.scratch_v = None

    def close(self):
        if self.peers is not None:
            self.peers.close()
            self.peers, self.can_scatter = None, False

    def edfdv(self, fx, e_loc, dt):
        return self.ops.edfdv_exp(fx, e_loc, self.kv, dt, flags=self.flags_v)

    def edfdv_scatter(self, fx, e_loc, dt):
        """e df/dv of the local x-shard, result delivered as everybody's v-shard"""
        topo, pr = self.topo, self.peers
        if self.scratch_x is None:
            self.scratch_x = torch.empty((topo.nxl, topo.nv), dtype=torch.float64, device=fx.device)
        self.ops.edfdv_exp_scatter(fx, e_loc, self.kv, dt, self.scratch_x, pr.pv, topo.rank, flags=self.flags_v)
        topo.stream_barrier(pr.token)            # every rank's block has landed in every v-shard
        return pr.fv

    def vdfdx_scatter(self, fv, dt):
        """v df/dx of the local v-shard, result delivered as everybody's x-shard; returns it with the
        all-reduced charge density (the all-reduce also orders the ranks)"""
        topo, pr = self.topo, self.peers
        if self.scratch_v is None:
            self.scratch_v = torch.empty((topo.nx, topo.nvl), dtype=torch.float64, device=fv.device)
        n = fv.new_empty(topo.nx)
        self.ops.vdfdx_exp_scatter(fv, self.kx, self.v_loc, dt, self.scratch_v, pr.px, topo.rank,
                                   flags=self.flags_x, density_out=n, dv=self.dv, edge_flags=self.edge)
        topo.all_reduce_sum(n)
        return pr.fx, n

    def vdfdx(self, fv, dt):
        n = fv.new_empty(fv.shape[0])            # partial density over the local columns, fused epilogue
        out = self.ops.vdfdx_exp(fv, self.kx, self.v_loc, dt, flags=self.flags_x, density_out=n, dv=self.dv,
                                 edge_flags=self.edge)
        attach_density(out, n)
        return out

    def density_partial(self, fv):
        n = cached_density(fv)
        if n is not None:
            return n
        return self.ops.moments(fv, self.v_loc, self.dv, nmom=1, edge_flags=self.edge)[0].contiguous()

    def poisson(self, n, driver):
        return self.ops.poisson(n, self.ook, driver)

    def fp(self, fx, moments_out):
        return self.ops.fp_step(fx, self.v, self.nu, self.dt, self.dv, self.fp_type, moments_out=moments_out,
                                vgrid=self.vgrid)

    def moments(self, fx, out):
        return self.ops.moments(fx, self.v, self.dv, nmom=8, out=out)

    def driver(self, t):
        if self.pulses is not None:
            return self.ops.driver(self.x, t, self.pulses)
        return torch.as_tensor(np.asarray(self.driver_host(t))).to(self.x.device)

    def xmodes_partial(self, fx, nmodes):
        m = self.ops.xmodes(fx, nmodes, x_offset=self.topo.x0, nx_total=self.topo.nx)[0]
        return torch.view_as_real(m).contiguous()

    def xmodes_partial_into(self, fx, out):
        """the same, written into ``out`` (1, nmodes, nv, 2) -- a row of the per-loop buffer -- without a copy"""
        self.ops.xmodes(fx, out.shape[1], out=out, x_offset=self.topo.x0, nx_total=self.topo.nx)

    def series_means(self, mom, e_loc, de_loc, out):
        """the seven series entries of vlapy/core/step.py:202-224 as means over this rank's x cells: one kernel"""
        self.ops.series(mom, e_loc, de_loc, out=out)

    def zeros(self, shape):
        return torch.zeros(shape, dtype=torch.float64, device=self.x.device)

    def upload(self, a):
        """host array (pinned or not) -> device tensor on this rank's GPU"""
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(a)
        return t.to(self.x.device, non_blocking=True)


def make_sharded_operators(topo, backend):
    """vdfdx / edfdv / field_solve / fp closures on Sharded handles (same call signatures as
    vlapy/core/vlasov.py, field.py, step.py closures)."""

    scatter = bool(getattr(backend, "can_scatter", False))

    def to_x(f):
        if f.pending is not None:
            return Sharded(backend.edfdv(*f.pending), "x")
        return f if f.layout == "x" else Sharded(topo.v_to_x(f.t), "x")

    def to_v(f):
        if f.pending is not None:
            if scatter:
                return Sharded(backend.edfdv_scatter(*f.pending), "v")
            f = to_x(f)
        return f if f.layout == "v" else Sharded(topo.x_to_v(f.t), "v")

    def vdfdx(f, dt):
        fv = to_v(f)
        if scatter:
            fx, n = backend.vdfdx_scatter(fv.t, dt)
            return Sharded(fx, "x", density=n)
        return Sharded(backend.vdfdx(fv.t, dt), "v")

    def edfdv(f, e, dt):
        e_loc = e[topo.x0: topo.x0 + topo.nxl].contiguous()
        fx = to_x(f)
        if scatter:
            return Sharded(None, "x", pending=(fx.t, e_loc, dt))     # evaluated by its consumer
        return Sharded(backend.edfdv(fx.t, e_loc, dt), "x")

    def field_solve(driver_field, f):
        if f.density is not None:
            return backend.poisson(f.density, driver_field)
        # density of a v-shard: partial integral over the local columns, then a sum over ranks
        n = topo.all_reduce_sum(backend.density_partial(to_v(f).t))
        return backend.poisson(n, driver_field)

    def fp_step(f, moments_out=None):
        return Sharded(backend.fp(to_x(f).t, moments_out), "x")

    return dict(vdfdx=vdfdx, edfdv=edfdv, field_solve=field_solve, fp_step=fp_step, to_x=to_x, to_v=to_v)


def get_sharded_timestep(all_params, stuff_for_time_loop, topo, backend=None):
    """The sharded counterpart of vlapy/core/step.py:286-328.  Returns
    timestep(state, t, de, store) -> state where state = {"e": full field, "f": Sharded};
    ``store`` (optional) receives the per-step stored quantities of this rank's x-slab."""
    from .core import vlasov_poisson
    if backend is None:
        # backend.peer_scatter = False keeps the NCCL all-to-all transposes (tests compare the two)
        backend = DeviceBackend(topo, stuff_for_time_loop, all_params["fokker-planck"]["type"],
                                peer_scatter=all_params.get("backend", {}).get("peer_scatter", True))
    ops_ = make_sharded_operators(topo, backend)
    stuff = dict(stuff_for_time_loop)
    stuff["driver_function"] = backend.driver
    vp_step = vlasov_poisson.get_time_integrator(
        all_params["vlasov-poisson"]["time"], ops_["vdfdx"], ops_["edfdv"], ops_["field_solve"], stuff)
    collide = all_params["nu"] > 0.0
    if all_params["nu"] < 0.0:
        raise NotImplementedError

    def timestep(state, t, de=None, store=None):
        e, f = vp_step(e=state["e"], f=state["f"], t=t)
        # the eight row moments of this step go straight into their row of the per-loop buffer (no copy)
        mom = store["moments_all"][store["i"]] if store is not None else None
        if collide:
            f = ops_["fp_step"](f, moments_out=mom)
        else:
            f = ops_["to_x"](f)
            if mom is not None:
                backend.moments(f.t, mom)
        peers = getattr(backend, "peers", None)
        if peers is not None and f.t is peers.fx:      # never hand out the receive buffer itself
            f = Sharded(f.t.clone(), "x")
        if store is not None:
            i = store["i"]
            sl = slice(topo.x0, topo.x0 + topo.nxl)
            el = e[sl]
            store["fields_e"][i] = el
            if de is not None:
                store["fields_driver"][i] = de[sl]
            # series: means over this rank's x cells (finished by an all-reduce at the storage cadence)
            if hasattr(backend, "series_means"):
                backend.series_means(mom, el, de[sl] if de is not None else None, store["series_mean"][i])
            else:
                row = store["series_mean"][i]
                row[0:3] = mom[0:3].mean(dim=1)
                row[3] = (el * el).mean()
                row[4] = (de[sl] * de[sl]).mean() if de is not None else 0.0
                row[5:7] = mom[6:8].mean(dim=1)
            if hasattr(backend, "xmodes_partial_into"):
                backend.xmodes_partial_into(f.t, store["modes_partial"][i:i + 1])
            else:
                store["modes_partial"][i] = backend.xmodes_partial(f.t, store["modes_partial"].shape[1])
            store["i"] = i + 1
        return {"e": e, "f": f}

    timestep.backend = backend
    return timestep


def make_store(topo, backend, nsteps, nmodes=2):
    """per-loop buffers of one rank: row i of every buffer belongs to step i of the loop.  ``moments_all`` receives the
    eight row moments of the kernels directly; ``fields_mom`` is its view of the six stored v-moments."""
    moments_all = backend.zeros((nsteps, 8, topo.nxl))
    return {
        "i": 0,
        "moments_all": moments_all,
        "fields_mom": moments_all[:, :6],
        "fields_e": backend.zeros((nsteps, topo.nxl)),
        "fields_driver": backend.zeros((nsteps, topo.nxl)),
        "series_mean": backend.zeros((nsteps, 7)),
        "modes_partial": backend.zeros((nsteps, nmodes, topo.nv, 2)),
    }


def finish_store(topo, store):
    """storage cadence: turn the slab means into global means (equal slabs: the mean of the means) and the slab
    partials of the x-modes into the modes (two small all-reduces)"""
    series = topo.all_reduce_sum(store["series_mean"].clone()) / topo.world
    modes = topo.all_reduce_sum(store["modes_partial"].clone())
    return series, torch.view_as_complex(modes.contiguous())


def ops_to_x(f, topo):
    return f.t if f.layout == "x" else topo.v_to_x(f.t)
