"""The reference's REAL setup through the drop-in boundary (CPU only; skipped where /root/reference is absent,
i.e. on the GPU box).

vlapy.outer_loop.get_everything_ready_for_outer_loop (vlapy/outer_loop.py:67-146) builds ``stuff_for_time_loop``
exactly as the reference's manager does; the two-line patch of INTEGRATION.md section 1 is applied IN MEMORY to
vlapy.outer_loop.get_sim_config_and_inner_loop_step (vlapy/outer_loop.py:31-64), so that
``all_params["backend"]["core"] = "b200"`` routes to vlapy_b200.outer_loop; both backends then run the same two
inner loops and every key / shape / dtype / value the storage layer reads (vlapy/storage.py:78-91) is compared.
Without a GPU the C-ABI calls are replaced by tests/stub_ops.py (oracle arithmetic): what is under test is
the host logic between the reference and the kernels, not the kernels."""
import copy
import os
import sys
import types
import warnings

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vlapy")), reason="reference tree not present")


@pytest.fixture()
def ref(monkeypatch):
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    ml = types.ModuleType("mlflow")
    ml.log_params = lambda *a, **k: None
    ml.log_metrics = lambda *a, **k: None
    monkeypatch.setitem(sys.modules, "mlflow", ml)
    monkeypatch.syspath_prepend(REF)
    from vlapy import initializers, outer_loop
    monkeypatch.setattr(outer_loop, "tqdm", lambda it: it)
    return types.SimpleNamespace(initializers=initializers, outer_loop=outer_loop)


@pytest.fixture()
def b200_on_cpu(monkeypatch):
    """vlapy_b200's host logic with the C-ABI wrappers replaced by the oracle-backed stand-in"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stub_ops
    import vlapy_b200.outer_loop as bo
    from vlapy_b200 import _util
    from vlapy_b200.core import collisions, field, step, vlasov
    cpu = torch.device("cpu")
    monkeypatch.setattr(_util, "device", lambda: cpu)
    monkeypatch.setattr(bo, "device", lambda: cpu)
    def to_dev(x):          # CPU tensors play the role of device tensors, numpy arrays that of host arrays
        if isinstance(x, torch.Tensor):
            return x.to(torch.float64), False
        return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))), True

    monkeypatch.setattr(_util, "to_dev", to_dev)
    for mod in (bo, collisions, field, step, vlasov):
        monkeypatch.setattr(mod, "ops", stub_ops)
        if hasattr(mod, "to_dev"):
            monkeypatch.setattr(mod, "to_dev", to_dev)
    stub_ops.calls.clear()
    return bo, stub_ops


class Rules:
    rules_to_store_f = {"time": "first-last", "space": ["k0", "k1"]}


def _params(ref, k0, nx, nv, tmax, nt, log_nu):
    p = ref.initializers.make_default_params_dictionary()
    p = ref.initializers.specify_epw_params_to_dict(k0=k0, all_params_dict=p)
    p = ref.initializers.specify_collisions_to_dict(log_nu_over_nu_ld=log_nu, all_params_dict=p)
    p["nx"], p["nv"], p["tmax"], p["nt"] = nx, nv, tmax, nt
    return p


def _run(outer_loop, p, pulse, steps_in_loop, n_loops):
    total = steps_in_loop * n_loops
    stuff = outer_loop.get_everything_ready_for_outer_loop(Rules, p, pulse, total)
    cfg, inner = outer_loop.get_sim_config_and_inner_loop_step(p, stuff, steps_in_loop, Rules.rules_to_store_f)
    outs = []
    for it in range(0, total, steps_in_loop):
        idx = np.arange(it, it + steps_in_loop)
        cfg = inner(temp_storage=cfg, driver_array=np.array(stuff["driver"][idx]), time_array=np.array(stuff["t"][idx]))
        outs.append({k: (copy.deepcopy({kk: np.array(vv) for kk, vv in v.items()}) if isinstance(v, dict) else np.array(v))
                     for k, v in cfg.items() if not k.startswith("_")})
    return stuff, outs


@pytest.mark.parametrize("case", ["landau_c1", "collisional"])
def test_reference_setup_through_the_b200_boundary(ref, b200_on_cpu, monkeypatch, case):
    bo, stub = b200_on_cpu
    ol = ref.outer_loop
    if case == "landau_c1":      # the grid and pulse of tests/test_landau_damping.py (C1), shortened
        k0, steps, loops = 0.3, 12, 2
        p = _params(ref, k0, 32, 512, 80, 500, None)
        pulse = {"first pulse": {"start_time": 0, "t_L": 6, "t_wL": 2.5, "t_R": 20, "t_wR": 2.5, "w0": p["w_epw"],
                                 "a0": 1e-7, "k0": k0}}
    else:
        k0, steps, loops = 0.35, 8, 2
        p = _params(ref, k0, 16, 128, 1000, 4000, -2)
        pulse = {"first pulse": {"start_time": 0, "t_L": 6, "t_wL": 2.5, "t_R": 25, "t_wR": 2.5, "w0": p["w_epw"],
                                 "a0": 4e-2, "k0": k0}}
    assert p["backend"]["core"] == "numpy"
    stuff_ref, outs_ref = _run(ol, copy.deepcopy(p), pulse, steps, loops)

    # ---- INTEGRATION.md section 1, applied in memory
    original = ol.get_sim_config_and_inner_loop_step

    def patched(all_params, stuff_for_time_loop, nt_in_loop, store_f_rules):
        if all_params["backend"]["core"] == "b200":
            stuff_for_time_loop["pulse_dictionary"] = pulse
            return bo.get_sim_config_and_inner_loop_step(all_params, stuff_for_time_loop, nt_in_loop, store_f_rules)
        return original(all_params, stuff_for_time_loop, nt_in_loop, store_f_rules)

    monkeypatch.setattr(ol, "get_sim_config_and_inner_loop_step", patched)
    pb = copy.deepcopy(p)
    pb["backend"]["core"] = "b200"
    pb["backend"]["cuda_graph"] = False
    stuff_b, outs_b = _run(ol, pb, pulse, steps, loops)
    assert "fp_step" in stub.calls if p["nu"] > 0 else "fp_step" not in stub.calls
    assert "driver" in stub.calls                       # device-side driver was selected via pulse_dictionary

    # every key the reference's stuff_for_time_loop carries was accepted as is
    assert set(stuff_ref) <= set(stuff_b)
    for a, b in zip(outs_ref, outs_b):
        assert set(a) <= set(b), set(a) - set(b)
        for k in a:
            if isinstance(a[k], dict):
                assert set(a[k]) == set(b[k]), (k, set(a[k]) ^ set(b[k]))
                for kk in a[k]:
                    ra, rb = np.asarray(a[k][kk]), np.asarray(b[k][kk])
                    assert ra.shape == rb.shape and ra.dtype == rb.dtype, (k, kk, ra.shape, rb.shape, ra.dtype, rb.dtype)
                    np.testing.assert_allclose(rb, ra, rtol=1e-9, atol=1e-12 * max(1.0, np.abs(ra).max()), err_msg=kk)
            else:
                ra, rb = np.asarray(a[k]), np.asarray(b[k])
                assert ra.shape == rb.shape, (k, ra.shape, rb.shape)
                assert ra.dtype == rb.dtype, (k, ra.dtype, rb.dtype)
                tol = 1e-6 if k == "stored_f" else 1e-11
                assert np.max(np.abs(rb - ra)) <= tol * max(1e-30, np.max(np.abs(ra))) + 1e-15, k


def test_unknown_backend_still_raises_like_the_reference(ref, b200_on_cpu):
    bo, _ = b200_on_cpu
    with pytest.raises(NotImplementedError):
        bo.get_sim_config_and_inner_loop_step({"backend": {"core": "jax"}}, {}, 1, Rules.rules_to_store_f)


def test_initial_storage_dictionary_matches_the_reference(ref, b200_on_cpu):
    """vlapy/outer_loop.py:149-215 against vlapy_b200.outer_loop.get_arrays_for_inner_loop on the reference's
    own stuff_for_time_loop: same keys, shapes, dtypes and initial values"""
    bo, _ = b200_on_cpu
    p = _params(ref, 0.3, 32, 512, 80, 500, None)
    pulse = {"first pulse": {"start_time": 0, "t_L": 6, "t_wL": 2.5, "t_R": 20, "t_wR": 2.5, "w0": p["w_epw"],
                             "a0": 1e-7, "k0": 0.3}}
    stuff = ref.outer_loop.get_everything_ready_for_outer_loop(Rules, p, pulse, 10)
    a = ref.outer_loop.get_arrays_for_inner_loop(stuff, 10, Rules.rules_to_store_f, this_np=np)
    b = bo.get_arrays_for_inner_loop(stuff, 10, Rules.rules_to_store_f, this_np=np)
    assert set(a) == set(b)
    for k in a:
        if isinstance(a[k], dict):
            assert set(a[k]) == set(b[k])
            for kk in a[k]:
                assert np.asarray(a[k][kk]).shape == np.asarray(b[k][kk]).shape
                assert np.asarray(a[k][kk]).dtype == np.asarray(b[k][kk]).dtype
        else:
            assert np.asarray(a[k]).shape == np.asarray(b[k]).shape, k
            assert np.asarray(a[k]).dtype == np.asarray(b[k]).dtype, k
            np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]))


def test_restart_from_a_stored_state_on_the_host_logic(ref, b200_on_cpu):
    """SURVEY 8f N3 (restart, the reference's TODO at vlapy/manager.py:118-119): a fresh inner loop that resumes from the
    host copy of the state after loop 1 reproduces loop 2 of an uninterrupted run -- and loop 2 of the reference's own
    numpy inner loop on the same setup"""
    bo, _ = b200_on_cpu
    ol = ref.outer_loop
    k0, steps, loops = 0.35, 6, 2
    p = _params(ref, k0, 16, 128, 1000, 4000, -2)
    pulse = {"first pulse": {"start_time": 0, "t_L": 6, "t_wL": 2.5, "t_R": 25, "t_wR": 2.5, "w0": p["w_epw"],
                             "a0": 4e-2, "k0": k0}}
    stuff, want = _run(ol, copy.deepcopy(p), pulse, steps, loops)
    pb = copy.deepcopy(p)
    pb["backend"]["core"] = "b200"
    pb["backend"]["cuda_graph"] = False
    stuff["pulse_dictionary"] = pulse

    def one_loop(cfg, inner, li):
        idx = np.arange(li * steps, (li + 1) * steps)
        return inner(temp_storage=cfg, driver_array=np.array(stuff["driver"][idx]), time_array=np.array(stuff["t"][idx]))

    cfg, inner = bo.get_sim_config_and_inner_loop_step(pb, stuff, steps, Rules.rules_to_store_f)
    cfg = one_loop(cfg, inner, 0)
    f1, e1, cum1 = np.array(cfg["f"]), np.array(cfg["e"]), float(cfg["series"]["mean_cum_de2"][-1])
    cfg = one_loop(cfg, inner, 1)
    straight = {"f": np.array(cfg["f"]), "e": np.array(cfg["e"]),
                "series": {k: np.array(v) for k, v in cfg["series"].items()}}
    # restart: a FRESH configuration and inner loop, fed with the stored state
    cfg2, inner2 = bo.get_sim_config_and_inner_loop_step(pb, stuff, steps, Rules.rules_to_store_f)
    cfg2 = bo.resume_from(cfg2, f1, e1, mean_cum_de2_previous=cum1)
    cfg2 = one_loop(cfg2, inner2, 1)
    np.testing.assert_array_equal(np.asarray(cfg2["f"]), straight["f"])
    np.testing.assert_array_equal(np.asarray(cfg2["e"]), straight["e"])
    for k in ("mean_n", "mean_T", "mean_e2", "mean_cum_de2"):
        np.testing.assert_allclose(np.asarray(cfg2["series"][k]), straight["series"][k], rtol=1e-14, atol=0)
        np.testing.assert_allclose(np.asarray(cfg2["series"][k]), want[1]["series"][k], rtol=1e-9, atol=1e-15)
    assert np.max(np.abs(np.asarray(cfg2["f"]) - want[1]["f"])) < 1e-11 * np.max(np.abs(want[1]["f"]))
    with pytest.raises(ValueError):
        bo.resume_from(cfg2, f1[0], e1)
