"""Mirror of vlapy/core/step.py: Vlasov-Poisson step, collision step, storage step, timestep.

The state (e, f) stays on the device across steps; the per-step stored quantities of the
reference (vlapy/core/step.py:116-283: six v-moments, the series means, the two lowest x-modes of
f) are produced by device reductions into device-resident ``(nt_in_loop, ...)`` buffers and are
copied to the host once per inner loop (vlapy_b200/outer_loop.py), which is where the
reference's storage layer picks them up (vlapy/manager.py:138-150).
"""
import numpy as np
import torch

from . import collisions, field, vlasov, vlasov_poisson
from .. import ops
from .._util import back, const, to_dev

FIELD_KEYS = ("e", "driver", "n", "j", "T", "q", "fv4", "vN")
SERIES_KEYS = ("mean_n", "mean_j", "mean_T", "mean_e2", "mean_de2", "mean_f2", "mean_flogf")


def get_device_driver_function(stuff_for_time_loop):
    """Device-side evaluation of vlapy/field_driver.py:24-50 when the pulse parameters are
    available (key ``pulse_dictionary``); otherwise the reference's host function is used and its
    result uploaded at every sub-step (correct, but one host->device copy per field solve)."""
    pulses = stuff_for_time_loop.get("pulse_dictionary")
    if pulses is None:
        return stuff_for_time_loop["driver_function"]
    arr = ops.pulses_to_array(pulses)
    x_d = const(stuff_for_time_loop["x"])

    def driver_function(current_time):
        if isinstance(current_time, ops.DevTime):
            return ops.driver_dev(x_d, current_time, arr)
        return ops.driver(x_d, current_time, arr)

    driver_function.on_device = True
    return driver_function


def get_vlasov_poisson_step(all_params, stuff_for_time_loop):
    """vlapy/core/step.py:29-67."""
    vdfdx = vlasov.get_vdfdx(
        stuff_for_time_loop=stuff_for_time_loop,
        vdfdx_implementation=all_params["vlasov-poisson"]["vdfdx"])
    edfdv = vlasov.get_edfdv(
        stuff_for_time_loop=stuff_for_time_loop,
        edfdv_implementation=all_params["vlasov-poisson"]["edfdv"])
    field_solver = field.get_field_solver(
        stuff_for_time_loop=stuff_for_time_loop,
        field_solver_implementation=all_params["vlasov-poisson"]["poisson"])
    stuff = dict(stuff_for_time_loop)
    stuff["driver_function"] = get_device_driver_function(stuff_for_time_loop)
    return vlasov_poisson.get_time_integrator(
        time_integrator_name=all_params["vlasov-poisson"]["time"],
        vdfdx=vdfdx, edfdv=edfdv, field_solver=field_solver, stuff_for_time_loop=stuff)


def get_collision_step(stuff_for_time_loop, all_params):
    """vlapy/core/step.py:70-113: identity for nu == 0, NotImplementedError for nu < 0."""
    if all_params["nu"] == 0.0:

        def take_collision_step(f):
            return f

    elif all_params["nu"] > 0.0:
        solver_name = all_params["fokker-planck"]["solver"]
        if solver_name not in ("naive", "batched_tridiagonal"):
            raise NotImplementedError(
                "Matrix Solver: <" + solver_name + "> has not yet been implemented on the b200 backend")
        collide = collisions.get_collision_operator(
            vax=stuff_for_time_loop["v"], nv=stuff_for_time_loop["nv"], nx=stuff_for_time_loop["nx"],
            nu=stuff_for_time_loop["nu"], dt=stuff_for_time_loop["dt"], dv=stuff_for_time_loop["dv"],
            operator=all_params["fokker-planck"]["type"])

        def take_collision_step(f, moments_out=None, out=None):
            # moments_out / out (device tensors): b200 extensions for callers that own their buffers (storage step,
            # captured step); the reference's signature is take_collision_step(f)
            f_d, host = to_dev(f)
            return back(collide(f_d.contiguous(), moments_out=moments_out, out=out), host)

        take_collision_step.fuses_moments = True
    else:
        raise NotImplementedError

    return take_collision_step


def get_f_update(store_f_rule):
    """vlapy/core/step.py:116-140."""
    if store_f_rule["space"] == "all":

        def get_f_to_store(f):
            return f

    elif store_f_rule["space"][0] == "k0":
        nmodes = len(store_f_rule)          # the reference's quirk: len of the rule DICT (== 2)

        def get_f_to_store(f):
            f_d, host = to_dev(f)
            m = ops.xmodes(f_d.contiguous(), nmodes)[0]
            return m.cpu().numpy() if host else m

    else:
        raise NotImplementedError
    return get_f_to_store


def get_fields_update(dv, v):
    """vlapy/core/step.py:143-175: e, driver and the six v-moments of f into row i."""
    v_d = const(v)

    def update_fields(temp_storage_fields, e, de, f, i, moments=None):
        if moments is None:
            moments = ops.moments(f, v_d, dv, nmom=8)
        block = temp_storage_fields.get("_block")
        if block is not None:
            # the eight fields are rows of ONE device buffer (len(FIELD_KEYS), nt, nx): three copies instead of eight
            block[0, i] = e
            block[1, i] = de
            block[2:8, i] = moments[:6]
        else:
            temp_storage_fields["e"][i] = e
            temp_storage_fields["driver"][i] = de
            for k, name in enumerate(("n", "j", "T", "q", "fv4", "vN")):
                temp_storage_fields[name][i] = moments[k]
        temp_storage_fields["_moments"] = moments
        return temp_storage_fields

    return update_fields


def get_series_update(dv):
    """vlapy/core/step.py:178-228: x-means of n, j, T, e^2, de^2, int f^2, int f ln f."""

    def update_series(temp_storage, e, de, f, i):
        series = temp_storage["series"]
        ops.series(temp_storage["fields"]["_moments"], e, de, out=series["_rows"][i])
        return series

    return update_series


def get_storage_step(stuff_for_time_loop):
    """vlapy/core/step.py:231-283."""
    dv = stuff_for_time_loop["dv"]
    v = stuff_for_time_loop["v"]
    store_f_function = get_f_update(store_f_rule=stuff_for_time_loop["rules_to_store_f"])
    update_fields = get_fields_update(dv=dv, v=v)
    update_series = get_series_update(dv=dv)

    def storage_step(temp_storage, e, de, f, i, moments=None):
        temp_storage["stored_f"][i] = store_f_function(f)
        temp_storage["e"] = e
        temp_storage["f"] = f
        temp_storage["fields"] = update_fields(
            temp_storage_fields=temp_storage["fields"], de=de, e=e, f=f, i=i, moments=moments)
        temp_storage["series"] = update_series(temp_storage=temp_storage, f=f, de=de, e=e, i=i)
        return temp_storage

    return storage_step


def get_step_parts(all_params, stuff_for_time_loop):
    """The pieces of a timestep for callers that compose them differently (the CUDA-graph inner loop
    of vlapy_b200/outer_loop.py): (vp_step, fp_step, fp_fuses_moments, store_f_function)."""
    vp_step = get_vlasov_poisson_step(all_params=all_params, stuff_for_time_loop=stuff_for_time_loop)
    fp_step = get_collision_step(all_params=all_params, stuff_for_time_loop=stuff_for_time_loop)
    store_f = get_f_update(store_f_rule=stuff_for_time_loop["rules_to_store_f"])
    return vp_step, fp_step, getattr(fp_step, "fuses_moments", False), store_f


def get_timestep(all_params, stuff_for_time_loop):
    """vlapy/core/step.py:286-328: timestep(temp_storage, i) = VP step, FP step, storage step.

    ``temp_storage`` is the device-resident dictionary built by
    vlapy_b200.outer_loop.get_arrays_for_inner_loop."""
    vp_step = get_vlasov_poisson_step(all_params=all_params, stuff_for_time_loop=stuff_for_time_loop)
    fp_step = get_collision_step(all_params=all_params, stuff_for_time_loop=stuff_for_time_loop)
    storage_step = get_storage_step(stuff_for_time_loop=stuff_for_time_loop)
    fused = getattr(fp_step, "fuses_moments", False)

    def timestep(temp_storage, i):
        e = temp_storage["e"]
        f = temp_storage["f"]
        t = temp_storage["time_batch"][i]
        de = temp_storage["driver_array_batch"][i]

        e, f = vp_step(e=e, f=f, t=t)
        if fused:
            mom = temp_storage["_moment_scratch"]
            f = fp_step(f, moments_out=mom)      # solve + moments of the new f in one kernel
            temp_storage = storage_step(temp_storage=temp_storage, e=e, de=de, f=f, i=i, moments=mom)
        else:
            f = fp_step(f=f)
            temp_storage = storage_step(temp_storage=temp_storage, e=e, de=de, f=f, i=i)
        return temp_storage, i

    return timestep
