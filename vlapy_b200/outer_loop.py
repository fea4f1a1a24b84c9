"""Mirror of the backend-dependent half of vlapy/outer_loop.py for ``backend.core == "b200"``.

``get_sim_config_and_inner_loop_step`` (vlapy/outer_loop.py:31-64) is the plug point: it returns the
storage dictionary and the inner-loop stepper.  Inside the inner loop everything is device
resident; on return the dictionary holds HOST numpy arrays under the keys the reference's
storage layer reads (vlapy/storage.py:78-91: time_batch, series, fields, stored_f, f, e), while
the device copies of e and f are kept under private keys for the next call.
"""
import numpy as np
import torch

from .core import step
from ._util import device

BACKEND = "b200"


def get_sim_config_and_inner_loop_step(all_params, stuff_for_time_loop, nt_in_loop, store_f_rules):
    """vlapy/outer_loop.py:31-64."""
    if all_params["backend"]["core"] != BACKEND:
        raise NotImplementedError(
            "The backend <" + all_params["backend"]["core"] + "> has not yet been implemented")
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


def _device_storage(temp_storage, nt, nx, nv, dev):
    """Device twins of the per-loop buffers of vlapy/outer_loop.py:190-215."""
    d = temp_storage.get("_dev")
    if d is not None and d["nt"] == nt:
        return d
    stored = temp_storage["stored_f"]
    d = {
        "nt": nt,
        "fields": {k: torch.zeros((nt, nx), dtype=torch.float64, device=dev) for k in step.FIELD_KEYS},
        "series_rows": torch.zeros((nt, 7), dtype=torch.float64, device=dev),
        "stored_f": torch.zeros(stored.shape, dtype=torch.complex128 if np.iscomplexobj(stored) else torch.float64,
                                device=dev),
        "moments": torch.zeros((8, nx), dtype=torch.float64, device=dev),
        "e": torch.from_numpy(np.ascontiguousarray(temp_storage["e"], dtype=np.float64)).to(dev),
        "f": torch.from_numpy(np.ascontiguousarray(temp_storage["f"], dtype=np.float64)).to(dev),
    }
    temp_storage["_dev"] = d
    return d


def get_inner_loop_stepper(all_params, stuff_for_time_loop, steps_in_loop):
    """vlapy/outer_loop.py:248-281: inner_loop(time_array, driver_array, temp_storage)."""
    if all_params["backend"]["core"] != BACKEND:
        raise NotImplementedError(
            "The backend: <" + all_params["backend"]["core"] + "> has not yet been implemented")
    one_step = step.get_timestep(all_params=all_params, stuff_for_time_loop=stuff_for_time_loop)

    def inner_loop(time_array, driver_array, temp_storage):
        dev = device()
        nx, nv = np.asarray(temp_storage["f"]).shape[-2:]
        d = _device_storage(temp_storage, steps_in_loop, nx, nv, dev)
        # host -> device: the driver rows and times of this loop (pinned when the caller pinned them)
        drv = torch.as_tensor(np.ascontiguousarray(driver_array, dtype=np.float64)).to(dev, non_blocking=True)
        work = {
            "time_batch": np.asarray(time_array, dtype=np.float64),
            "driver_array_batch": drv,
            "e": d["e"], "f": d["f"],
            "stored_f": d["stored_f"],
            "fields": dict(d["fields"]),
            "series": {"_rows": d["series_rows"]},
            "_moment_scratch": d["moments"],
        }
        for it in range(steps_in_loop):
            work, _ = one_step(work, it)
        d["e"], d["f"] = work["e"], work["f"]
        # device -> host, once per inner loop (the storage cadence of vlapy/manager.py:138-150)
        temp_storage["time_batch"] = np.asarray(time_array)
        temp_storage["driver_array_batch"] = np.asarray(driver_array)
        for k in step.FIELD_KEYS:
            temp_storage["fields"][k] = d["fields"][k].cpu().numpy()
        rows = d["series_rows"].cpu().numpy()
        for j, k in enumerate(step.SERIES_KEYS):
            temp_storage["series"][k] = rows[:, j].copy()
        sf = d["stored_f"].cpu().numpy()
        temp_storage["stored_f"] = sf.astype(np.complex64) if np.iscomplexobj(sf) else sf
        temp_storage["e"] = d["e"].cpu().numpy()
        temp_storage["f"] = d["f"].cpu().numpy()
        post_inner_loop_update(temp_storage, np)
        return temp_storage

    return inner_loop
