"""Mirror of vlapy/core/field.py for the b200 backend: charge density and spectral Poisson."""
import torch

from .. import ops
from .._util import back, cached_density, const, to_dev


def compute_charges(f, dv):
    """vlapy/core/field.py:27-36 -- n[x] = trapz_v f."""
    f_d, host = to_dev(f)
    v_dummy = torch.zeros(f_d.shape[-1], dtype=torch.float64, device=f_d.device)
    n = ops.moments(f_d.contiguous(), v_dummy, dv, nmom=1)[0].reshape(f_d.shape[:-1])
    return back(n, host)


def solve_for_field(charge_density, one_over_kx):
    """vlapy/core/field.py:52-63 -- E = Re ifft(1j * one_over_kx * fft(1 - charge_density))."""
    n_d, host = to_dev(charge_density)
    ook = const(one_over_kx)
    if ook.shape != n_d.shape:
        ook = ook.expand_as(n_d).contiguous()
    return back(ops.poisson(n_d.contiguous(), ook), host)


def get_spectral_solver(dv, one_over_kx):
    """vlapy/core/field.py:66-88 -- total field = driver + self-consistent field."""
    ook = const(one_over_kx)

    def solve_total_electric_field(driver_field, f):
        f_d, host = to_dev(f)
        drv, _ = to_dev(driver_field)
        n = cached_density(f) if isinstance(f, torch.Tensor) else None   # reduced by the v df/dx epilogue that produced f
        if n is None:
            n = compute_charges(f_d, dv)
        o = ook if ook.shape == n.shape else ook.expand_as(n).contiguous()
        return back(ops.poisson(n.contiguous(), o, drv.contiguous()), host)

    return solve_total_electric_field


def get_field_solver(stuff_for_time_loop, field_solver_implementation="spectral"):
    """vlapy/core/field.py:91-107."""
    if field_solver_implementation == "spectral":
        field_solver = get_spectral_solver(
            dv=stuff_for_time_loop["dv"], one_over_kx=stuff_for_time_loop["one_over_kx"])
    else:
        raise NotImplementedError
    return field_solver
