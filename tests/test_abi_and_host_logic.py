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
        vlasov.get_vdfdx({"kx": cfg["kx"], "v": cfg["v"], "x": cfg["x"]}, "sl")
    with pytest.raises(NotImplementedError):
        vlasov.get_edfdv({"kv": cfg["kv"], "dv": cfg["dv"]}, "sl")
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
