"""Generate the golden fixtures in tests/golden/ by running the REFERENCE ITSELF.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py [/path/to/reference]

The reference package is imported in place from its source tree (nothing is copied); ``mlflow``
is stubbed because ``vlapy/initializers.py`` imports it at module top and it is not installed.
Outputs are small ``.npz`` files committed next to this script; ``tests/test_oracle_vs_golden.py``
pins the oracle against them and the ``-m gpu`` tests pin the CUDA path against the same files.
"""
import os
import sys
import types
import warnings

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

warnings.filterwarnings("ignore", category=DeprecationWarning)
ml = types.ModuleType("mlflow")
ml.log_params = lambda *a, **k: None
ml.log_metrics = lambda *a, **k: None
sys.modules["mlflow"] = ml
sys.path.insert(0, REF)

from vlapy import initializers, outer_loop, field_driver  # noqa: E402
from vlapy.core import vlasov, field, collisions, step, vlasov_poisson  # noqa: E402
from vlapy.diagnostics import low_level_helpers as llh  # noqa: E402
import scipy  # noqa: E402


def save(name, **arrays):
    arrays["numpy_version"] = np.array(np.__version__)
    arrays["scipy_version"] = np.array(scipy.__version__)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def params_for(k0, nx, nv, tmax, nt, log_nu=None):
    p = initializers.make_default_params_dictionary()
    p = initializers.specify_epw_params_to_dict(k0=k0, all_params_dict=p)
    p = initializers.specify_collisions_to_dict(log_nu_over_nu_ld=log_nu, all_params_dict=p)
    p["nx"], p["nv"], p["tmax"], p["nt"] = nx, nv, tmax, nt
    return p


def pulse_for(p, k0, a0, t_R):
    return {"first pulse": {"start_time": 0, "t_L": 6, "t_wL": 2.5, "t_R": t_R, "t_wR": 2.5,
                            "w0": p["w_epw"], "a0": a0, "k0": k0}}


class Rules:
    rules_to_store_f = {"time": "first-last", "space": ["k0", "k1"]}


def manager_loop_sizes(p):
    # manager.py:61-83
    mem_f_store = 2 * len(Rules.rules_to_store_f["space"]) * p["nv"]
    mem_field_store = p["nx"] * 8
    steps = int(1e9 * p["backend"]["max_GB_for_device"] / (6 * (mem_f_store + mem_field_store) * 8))
    if steps > p["nt"]:
        steps = int(p["nt"] / 1.25)
    return steps, p["nt"] // steps + 1


# ---------------------------------------------------------------------------------------------
# 1. operator-level fixtures (each operator in isolation)
# ---------------------------------------------------------------------------------------------

def operator_fixture(name, nx, nv, k0, seed, noise, diagonals=False):
    rng = np.random.default_rng(seed)
    xmax = 2 * np.pi / k0
    dx, x, kx, ook = initializers.initialize_spatial_quantities(0.0, xmax, nx)
    dv, v, kv = initializers.initialize_velocity_quantities(6.4, nv)
    f = initializers.initialize_distribution(nx, nv, 6.4) * (1.0 + 0.1 * np.sin(k0 * x))[:, None]
    f = f + noise * rng.standard_normal((nx, nv))
    e = 0.05 * np.cos(k0 * x) + 0.01 * rng.standard_normal(nx)
    drv = 0.02 * np.sin(k0 * x)
    dt, nu = 0.16, 1e-3
    out = dict(f=f, e=e, drv=drv, x=x, kx=kx, one_over_kx=ook, v=v, kv=kv, dv=np.array(dv),
               dt=np.array(dt), nu=np.array(nu), k0=np.array(k0))
    out["vdfdx"] = vlasov.get_vdfdx_exponential(kx=kx, v=v)(f, dt)
    out["vdfdx_neg"] = vlasov.get_vdfdx_exponential(kx=kx, v=v)(f, -0.066 * dt)
    out["edfdv"] = vlasov.get_edfdv_exponential(kv=kv)(f, e, 0.5 * dt)
    out["edfdv_neg"] = vlasov.get_edfdv_exponential(kv=kv)(f, e, -0.21 * dt)
    out["cd2"] = vlasov.get_edfdv_center_differenced(dv=dv)(f, e, 0.5 * dt)
    out["charges"] = field.compute_charges(f, dv)
    out["efield"] = field.get_spectral_solver(dv=dv, one_over_kx=ook)(drv, f)
    fpos = np.abs(f) + 1e-12
    out["fpos"] = fpos
    for op in ("lb", "dg"):
        a, b, c = collisions.get_batched_array_maker(v, nv, nx, nu, dt, dv, operator=op)(fpos)
        if diagonals:
            out[op + "_a"], out[op + "_b"], out[op + "_c"] = a, b, c
        out[op + "_solve"] = collisions.get_batched_tridiag_solver(nv)(a, b, c, fpos)
    # per-step stored quantities of step.py:164-171, 202-224, 132-135
    fields = {k: np.zeros((1, nx)) for k in ("e", "driver", "n", "j", "T", "q", "fv4", "vN")}
    fields = step.get_fields_update(dv=dv, v=v)(fields, e, drv, fpos, 0)
    ts = {"fields": fields, "series": {k: np.zeros(1) for k in
          ("mean_n", "mean_j", "mean_T", "mean_e2", "mean_de2", "mean_f2", "mean_flogf")}}
    series = step.get_series_update(dv=dv)(ts, e, drv, fpos, 0)
    out["moments"] = np.stack([fields[k][0] for k in ("n", "j", "T", "q", "fv4", "vN")])
    out["series"] = np.array([series[k][0] for k in
                              ("mean_n", "mean_j", "mean_T", "mean_e2", "mean_de2", "mean_f2", "mean_flogf")])
    out["modes"] = step.get_f_update(Rules.rules_to_store_f)(fpos)
    save(name, **out)


# ---------------------------------------------------------------------------------------------
# 2. the reference's own unit-test cases
# ---------------------------------------------------------------------------------------------

def collision_test_fixture():
    # tests/test_collisions.py:168-213
    sys.path.insert(0, REF)
    from tests import helpers
    nx, nv, nu, dt, v0, vmax = 2, 1024, 1e-2, 0.1, 1.0, 6.0
    dv = 2 * vmax / nv
    v = np.linspace(-vmax + dv / 2.0, vmax - dv / 2.0, nv)
    out = dict(v=v, dv=np.array(dv), nu=np.array(nu), dt=np.array(dt))
    for vshift in (0.0, 0.5, 1.5):
        f = helpers.__initialize_f__(nx=nx, v=v, v0=v0, vshift=vshift)
        out["f_%g" % vshift] = f
        for op in ("lb", "dg"):
            fp = step.get_collision_step(
                stuff_for_time_loop=dict(f=f, v=v, nv=nv, nx=nx, nu=nu, dt=dt, dv=dv),
                all_params={"fokker-planck": {"type": op, "solver": "batched_tridiagonal"}, "nu": nu})
            g = f.copy()
            for _ in range(16):
                g = fp(g)
            out["out_%s_%g" % (op, vshift)] = g
    save("collisions_unit", **out)


def fieldsolver_test_fixture():
    # tests/test_fieldsolver.py:27-54 (nx = 96, not a power of two)
    nx, kp = 96, 0.25
    xmax = 2 * np.pi / kp
    dx = xmax / nx
    ax = np.linspace(dx / 2, xmax - dx / 2, nx)
    kx = np.fft.fftfreq(ax.size, d=dx) * 2.0 * np.pi
    ook = np.zeros_like(kx)
    ook[1:] = 1.0 / kx[1:]
    rho = [1.0 + np.sin(kp * ax), 1.0 + np.cos(2 * kp * ax),
           1.0 + np.sin(2 * kp * ax) + np.cos(8 * kp * ax)]
    out = dict(one_over_kx=ook, x=ax)
    for i, r in enumerate(rho):
        out["rho_%d" % i] = 1.0 - r          # what the test passes as charge_density
        out["e_%d" % i] = field.solve_for_field(charge_density=1.0 - r, one_over_kx=ook)
    save("fieldsolver_unit", **out)


# ---------------------------------------------------------------------------------------------
# 3. integrated runs through the reference's own outer_loop (mlflow stubbed)
# ---------------------------------------------------------------------------------------------

def run_reference(p, pulse, nsteps_total=None, steps_in_loop=None, n_loops=None):
    if steps_in_loop is None:
        steps_in_loop, n_loops = manager_loop_sizes(p)
    total = steps_in_loop * n_loops
    stuff = outer_loop.get_everything_ready_for_outer_loop(Rules, p, pulse, total)
    cfg, inner = outer_loop.get_sim_config_and_inner_loop_step(p, stuff, steps_in_loop, Rules.rules_to_store_f)
    import tqdm as _t
    outer_loop.tqdm = lambda it: it
    outs = []
    for it in range(0, total, steps_in_loop):
        idx = np.arange(it, it + steps_in_loop)
        cfg = inner(temp_storage=cfg, driver_array=np.array(stuff["driver"][idx]),
                    time_array=np.array(stuff["t"][idx]))
        outs.append({
            "e_hist": cfg["fields"]["e"].copy(), "time": cfg["time_batch"].copy(),
            "fields": {k: val.copy() for k, val in cfg["fields"].items()},
            "series": {k: np.array(val).copy() for k, val in cfg["series"].items()},
            "stored_f": cfg["stored_f"].copy(), "f": cfg["f"].copy(), "e": cfg["e"].copy()})
    return stuff, outs


class Shim:
    def __init__(self, data, t):
        self.data = data
        self.coords = {"time": types.SimpleNamespace(data=t)}


def landau_fixture():
    k0 = 0.3
    out = {}
    for integ in ("leapfrog", "pefrl", "h-sixth"):
        for edfdv in ("exponential", "cd2"):
            p = params_for(k0, 32, 512, 80, 500)
            p["vlasov-poisson"]["time"] = integ
            p["vlasov-poisson"]["edfdv"] = edfdv
            pulse = pulse_for(p, k0, 1e-7, 20)
            stuff, outs = run_reference(p, pulse)
            e_hist = np.concatenate([o["e_hist"] for o in outs])
            tax = np.concatenate([o["time"] for o in outs])
            rate = llh.get_damping_rate(Shim(e_hist, tax))
            key = integ + "_" + edfdv
            out["rate_" + key] = np.array(rate)
            print(key, "damping rate", rate, "nu_ld", p["nu_ld"])
            if edfdv == "exponential":
                out["e_final_" + integ] = outs[-1]["e"]
                out["mean_n_last_" + integ] = np.array(outs[-1]["series"]["mean_n"][-1])
            if key == "leapfrog_exponential":
                out["e_hist_leapfrog"] = e_hist
                out["time"] = tax
                out["f_final_leapfrog"] = outs[-1]["f"]
                # first inner loop: everything the storage layer receives, first 12 steps
                o = outs[0]
                for k in o["fields"]:
                    out["fields_" + k] = o["fields"][k][:12]
                for k in o["series"]:
                    out["series_" + k] = o["series"][k][:12]
                out["stored_f"] = o["stored_f"][:12]
        out["nu_ld"] = np.array(p["nu_ld"])
        out["w_epw"] = np.array(p["w_epw"])
        out["dt"] = np.array(stuff["dt"])
    save("landau_c1", **out)


def short_run_fixture():
    # 50 vp_steps from e = 0 at C1 (no FP, no storage) for the three schedules.
    k0 = 0.3
    out = {}
    for integ in ("leapfrog", "pefrl", "h-sixth"):
        p = params_for(k0, 32, 512, 80, 500)
        p["vlasov-poisson"]["time"] = integ
        pulse = pulse_for(p, k0, 1e-7, 20)
        stuff = outer_loop.get_everything_ready_for_outer_loop(Rules, p, pulse, 800)
        vp = step.get_vlasov_poisson_step(p, stuff)
        e, f = stuff["e"].copy(), stuff["f"].copy()
        for i in range(50):
            e, f = vp(e=e, f=f, t=stuff["t"][i])
        out["e_" + integ], out["f_" + integ] = e, f
        print(integ, "max|e| %.12e e[0] %.12e f[5,300] %.16g sum %.12f" %
              (np.abs(e).max(), e[0], f[5, 300], f.sum()))
    save("vp50_c1", **out)


def nlepw_fixture():
    # C2 at reduced nv for fixture size: 64 x 512 and the full 256 x 2048 summarised.
    k0 = 0.35
    out = {}
    for op in ("lb", "dg"):
        p = params_for(k0, 256, 2048, 1000, 4000, log_nu=-4)
        p["fokker-planck"]["type"] = op
        pulse = pulse_for(p, k0, 4e-2, 25)
        stuff = outer_loop.get_everything_ready_for_outer_loop(Rules, p, pulse, 64)
        vp = step.get_vlasov_poisson_step(p, stuff)
        fp = step.get_collision_step(stuff, p)
        e, f = stuff["e"].copy(), stuff["f"].copy()
        for i in range(40):
            e, f = vp(e=e, f=f, t=stuff["t"][i])
            f = fp(f=f)
        out["e_" + op] = e
        out["f_sub_" + op] = f[::8, ::16].copy()
        out["f_sum_" + op] = np.array(f.sum())
        out["f_min_" + op] = np.array(f.min())
        out["f_100_1300_" + op] = np.array(f[100, 1300])
        print(op, "max|e| %.15g e[0] %.15g f[100,1300] %.16g sum %.11f min %.4g" %
              (np.abs(e).max(), e[0], f[100, 1300], f.sum(), f.min()))
        out["nu"] = np.array(p["nu"])
        out["dt"] = np.array(stuff["dt"])
    # a small collisional run with the full storage step, through the reference inner loop
    p = params_for(k0, 16, 128, 1000, 4000, log_nu=-2)
    pulse = pulse_for(p, k0, 4e-2, 25)
    stuff, outs = run_reference(p, pulse, steps_in_loop=24, n_loops=2)
    for li, o in enumerate(outs):
        for k in o["fields"]:
            out["small_fields_%s_%d" % (k, li)] = o["fields"][k]
        for k in o["series"]:
            out["small_series_%s_%d" % (k, li)] = o["series"][k]
        out["small_stored_f_%d" % li] = o["stored_f"]
        out["small_f_%d" % li] = o["f"]
        out["small_e_%d" % li] = o["e"]
    out["small_nu"] = np.array(p["nu"])
    save("nlepw_c2", **out)


def sl_fixture():
    """N4: the reference's semi-Lagrangian operators (vlapy/core/vlasov.py:42-80, 168-210) in isolation on seeded
    inputs, and the Landau-damping run of tests/test_landau_damping.py with the sl flavours (leapfrog)."""
    out = {}
    for tag, nx, nv, k0, seed in (("small", 16, 64, 0.3, 0), ("c1", 32, 512, 0.3, 1)):
        rng = np.random.default_rng(seed)
        xmax = 2 * np.pi / k0
        dx, x, kx, ook = initializers.initialize_spatial_quantities(0.0, xmax, nx)
        dv, v, kv = initializers.initialize_velocity_quantities(6.4, nv)
        f = initializers.initialize_distribution(nx, nv, 6.4) * (1.0 + 0.1 * np.sin(k0 * x))[:, None]
        f = f + 1e-3 * rng.standard_normal((nx, nv))
        e = 0.05 * np.cos(k0 * x) + 0.01 * rng.standard_normal(nx)
        out.update({tag + "_f": f, tag + "_e": e, tag + "_x": x, tag + "_v": v})
        for name, dt in (("a", 0.16), ("b", -0.0106), ("c", 0.9)):          # c: shifts beyond one cell (clamped feet)
            out["%s_vdfdx_%s" % (tag, name)] = vlasov.get_vdfdx_sl(x=x, v=v)(f, dt)
            out["%s_edfdv_%s" % (tag, name)] = vlasov.get_edfdv_sl(x=x, v=v)(f, e, dt)
        out[tag + "_dts"] = np.array([0.16, -0.0106, 0.9])
    k0 = 0.3
    for vdfdx, edfdv in (("sl", "exponential"), ("exponential", "sl"), ("sl", "sl")):
        p = params_for(k0, 32, 512, 80, 500)
        p["vlasov-poisson"]["time"] = "leapfrog"
        p["vlasov-poisson"]["vdfdx"] = vdfdx
        p["vlasov-poisson"]["edfdv"] = edfdv
        pulse = pulse_for(p, k0, 1e-7, 20)
        stuff, outs = run_reference(p, pulse)
        e_hist = np.concatenate([o["e_hist"] for o in outs])
        tax = np.concatenate([o["time"] for o in outs])
        rate = llh.get_damping_rate(Shim(e_hist, tax))
        key = "%s_%s" % (vdfdx, edfdv)
        out["rate_" + key] = np.array(rate)
        out["e_final_" + key] = outs[-1]["e"]
        out["f_final_sub_" + key] = outs[-1]["f"][::2, ::8].copy()
        print(key, "damping rate", rate, "nu_ld", p["nu_ld"])
        out["nu_ld"] = np.array(p["nu_ld"])
    save("sl_ops", **out)


def nlepw_series200_fixture():
    """C2 exactly as run_nlepw.py sizes it (256 x 2048, k0 = 0.35, lb) through the reference's own inner loop for
    200 steps: the per-step series (SURVEY 8d integrated acceptance: mean_n / mean_T / mean_e2 to 1e-9) and fields
    at a few steps.  A separate small file so that it can be regenerated alone (about a minute of CPU)."""
    k0 = 0.35
    p = params_for(k0, 256, 2048, 1000, 4000, log_nu=-4)
    pulse = pulse_for(p, k0, 4e-2, 25)
    stuff, outs = run_reference(p, pulse, steps_in_loop=200, n_loops=1)
    o = outs[0]
    out = {"series_" + k: np.asarray(v) for k, v in o["series"].items()}
    for k in ("e", "n", "T"):
        out["fields_%s_sub" % k] = o["fields"][k][::20].copy()
    out["e_final"] = o["e"]
    out["f_final_sub"] = o["f"][::8, ::16].copy()
    out["nu"] = np.array(p["nu"])
    save("nlepw_c2_series200", **out)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "series200":
        nlepw_series200_fixture()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "sl":
        sl_fixture()
        sys.exit(0)
    operator_fixture("ops_small", nx=16, nv=64, k0=0.3, seed=0, noise=1e-3, diagonals=True)
    operator_fixture("ops_c1", nx=32, nv=512, k0=0.3, seed=1, noise=1e-3)
    operator_fixture("ops_white", nx=64, nv=128, k0=0.35, seed=2, noise=0.3)
    collision_test_fixture()
    fieldsolver_test_fixture()
    short_run_fixture()
    landau_fixture()
    nlepw_fixture()
