"""CPU tests: the C-ABI library loads and exports every symbol the header declares; the host-side
mirrors (schedules, dictionaries, error behaviour) follow the reference.  No GPU compute here."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden, rel_err
from oracle import vpfp_oracle as O
from vlapy_b200 import _lib, outer_loop
from vlapy_b200.core import vlasov_poisson, vlasov, step


def test_library_exports_every_header_symbol():
    so = _lib.build()
    assert os.path.exists(so)
    hdr = open(os.path.join(ROOT, "include", "vpfp_b200.h")).read()
    declared = set(re.findall(r"\b(vpfp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    import ctypes
    h = ctypes.CDLL(so)
    for name in declared:
        assert hasattr(h, name), name
    assert declared == set(_lib.exported_symbols())
    assert _lib.lib().vpfp_abi_version() == _lib.ABI_VERSION


@pytest.mark.parametrize("integ", ["leapfrog", "pefrl", "h-sixth"])
def test_schedule_mirrors_compose_like_the_reference(integ):
    """The b200 time-integrator mirrors, fed with the ORACLE's numpy operators as closures, must
    reproduce the reference run bit-for-tolerance (golden: 50 steps at C1)."""
    g = golden("vp50_c1")
    cfg = O.landau_config()
    vdfdx = lambda f, dt: O.vdfdx_exponential(f, dt, cfg["kx"], cfg["v"])            # noqa: E731
    edfdv = lambda f, e, dt: O.edfdv_exponential(f, e, dt, cfg["kv"])                # noqa: E731
    fs = lambda driver_field, f: O.field_solve(driver_field, f, cfg["dv"], cfg["one_over_kx"])  # noqa: E731
    vp = vlasov_poisson.get_time_integrator(integ, vdfdx, edfdv, fs,
                                            {"dt": cfg["dt"], "driver_function": cfg["driver_function"]})
    e, f = cfg["e0"].copy(), cfg["f0"].copy()
    for i in range(50):
        e, f = vp(e=e, f=f, t=cfg["dt"] * i)
    assert rel_err(f, g["f_" + integ]) < 1e-13
    assert np.max(np.abs(e - g["e_" + integ])) < 2e-14


def test_unknown_flavours_raise_like_the_reference():
    cfg = O.landau_config()
    with pytest.raises(NotImplementedError):
        vlasov_poisson.get_time_integrator("rk4", None, None, None, {"dt": 0.1, "driver_function": None})
    with pytest.raises(NotImplementedError):
        vlasov.get_vdfdx({"kx": cfg["kx"], "v": cfg["v"], "x": cfg["x"]}, "weno")
    with pytest.raises(NotImplementedError):
        vlasov.get_edfdv({"kv": cfg["kv"], "dv": cfg["dv"]}, "upwind")
    with pytest.raises(NotImplementedError):          # <sl> needs at least four points along the advected axis
        vlasov._axis_spacing(np.arange(3.0))
    with pytest.raises(NotImplementedError):
        step.get_collision_step({}, {"nu": -1.0})
    with pytest.raises(NotImplementedError):
        outer_loop.get_sim_config_and_inner_loop_step({"backend": {"core": "numpy"}}, {}, 1, {})
    # nu == 0 is the identity (vlapy/core/step.py:80-83)
    ident = step.get_collision_step({}, {"nu": 0.0})
    x = object()
    assert ident(x) is x


def test_storage_dictionary_matches_reference_layout():
    cfg = O.landau_config()
    rules = {"time": "first-last", "space": ["k0", "k1"]}
    d = outer_loop.get_arrays_for_inner_loop({"f": cfg["f0"], "e": cfg["e0"]}, 7, rules)
    assert d["stored_f"].shape == (7, 2, 512) and d["stored_f"].dtype == np.complex64
    assert set(d["fields"]) == {"e", "driver", "n", "j", "T", "q", "fv4", "vN"}
    assert all(v.shape == (7, 32) for v in d["fields"].values())
    assert set(d["series"]) == set(O.SERIES_KEYS)
    np.testing.assert_allclose(d["stored_f"][0], np.fft.fft(cfg["f0"], axis=0)[:2].astype(np.complex64))
    d["series"]["mean_de2"] = np.arange(7.0)
    d["series"]["mean_T"] = np.ones(7)
    d["series"]["mean_e2"] = np.ones(7)
    outer_loop.post_inner_loop_update(d)
    np.testing.assert_allclose(d["series"]["mean_cum_de2"], np.cumsum(np.arange(7.0)))
    assert d["mean_cum_de2_previous"] == 21.0


def test_non_fftfreq_wavenumbers_are_rejected():
    with pytest.raises(NotImplementedError):
        vlasov._check_wavenumbers(np.arange(8.0), "v df/dx")
    vlasov._check_wavenumbers(np.fft.fftfreq(8) * 3.0, "v df/dx")
    vlasov._check_wavenumbers(np.fft.fftfreq(2), "v df/dx")


def test_flag_constants_match_the_header():
    """the Python mirrors of the C-ABI flag macros (vlapy_b200/ops.py) carry the header's values"""
    from vlapy_b200 import ops
    hdr = open(os.path.join(ROOT, "include", "vpfp_b200.h")).read()
    macros = {m: int(v) for m, v in re.findall(r"#define\s+(VPFP_[A-Z0-9_]+)\s+(\d+)\b", hdr)}
    for name in ("PHASE_EXACT", "PHASE_TABLE", "FORCE_GENERIC", "FORCE_THREE_PASS"):
        assert macros["VPFP_" + name] == getattr(ops, name), name
    flags = [macros["VPFP_" + n] for n in ("PHASE_TABLE", "FORCE_GENERIC", "FORCE_THREE_PASS")]
    assert len(set(flags)) == len(flags) and all(f & (f - 1) == 0 for f in flags)      # distinct single bits


def test_run_loops_overlaps_storage_with_the_next_batch_and_respects_the_two_buffer_sets():
    """outer_loop.run_loops with a stand-in inner loop that, like the real one with
    backend.pinned_sets = 2, returns views into two alternating buffers: every batch is consumed once, in order,
    with the values of its own batch (a buffer is never rewritten while its consumer is still reading), and the
    consumer of batch i runs while batch i+1 computes."""
    import threading
    import time
    bufs = [np.zeros(4), np.zeros(4)]
    state = {"n": 0}
    log = []
    busy = threading.Event()
    overlapped = []

    def inner_loop(time_array, driver_array, temp_storage):
        overlapped.append(busy.is_set())                 # is a consumer running while this batch "computes"?
        b = bufs[state["n"] % 2]
        b[:] = time_array[0]
        time.sleep(0.02)
        state["n"] += 1
        out = dict(temp_storage)
        out.update(f=b, fields={"n": b}, time_batch=np.asarray(time_array), _dev="private")
        return out

    def consume(snap):
        busy.set()
        first = float(snap["f"][0])
        time.sleep(0.03)                                 # slower than a batch: the loop has to wait for the buffer
        assert "_dev" not in snap
        log.append((float(snap["time_batch"][0]), first, float(snap["f"][0]), float(snap["fields"]["n"][3])))
        busy.clear()

    batches = [(np.full(3, float(i)), None) for i in range(6)]
    final = outer_loop.run_loops(inner_loop, {"f": None}, batches, consume)
    assert [l[0] for l in log] == [float(i) for i in range(6)]
    assert all(l[0] == l[1] == l[2] == l[3] for l in log)     # never overwritten under the consumer
    assert any(overlapped[1:])
    assert final["_dev"] == "private" and state["n"] == 6

    def failing(snap):
        raise RuntimeError("disk full")
    with pytest.raises(RuntimeError, match="disk full"):
        outer_loop.run_loops(inner_loop, {"f": None}, batches[:3], failing)
