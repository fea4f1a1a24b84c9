"""Mirror of vlapy/core/collisions.py for the b200 backend.

The reference materialises the three diagonals (vlapy/core/collisions.py:44-81, 104-158) and
sweeps them with a Thomas loop (:232-263).  On the device the diagonals are two scalars per row
and the solve is fused with the moment computation (csrc/rowops.h FpProg), so the product path
is ``get_collision_operator`` below; ``get_batched_array_maker`` / ``get_matrix_solver`` keep the
reference's two-stage interface for callers that want (a, b, c) explicitly.
"""
import torch

from .. import ops
from .._util import back, const, to_dev


def get_collision_operator(vax, nv, nx, nu, dt, dv, operator="lb"):
    """Fused matrix build + solve: f -> f_new. Optionally also returns the 8 moments of f_new."""
    if operator not in ops.FP_OPS:
        raise NotImplementedError(
            "Collision Operator: <" + str(operator) + "> has not yet been implemented on the b200 backend")
    v_d = const(vax)
    vgrid = ops.linspace_params(vax)      # np.linspace grids get the specialised kernel

    def collide(f_xv, moments_out=None):
        return ops.fp_step(f_xv, v_d, nu, dt, dv, operator, moments_out=moments_out, vgrid=vgrid)

    return collide


def get_batched_array_maker(vax, nv, nx, nu, dt, dv, operator="lb"):
    """vlapy/core/collisions.py:292-317 -- f -> (a, b, c) diagonals (device tensors or numpy)."""
    if operator not in ops.FP_OPS:
        raise NotImplementedError(
            "Collision Operator: <" + str(operator) + "> has not yet been implemented on the b200 backend")
    v_d = const(vax)

    def make_arrays_for_matrix(f_xv):
        f_d, host = to_dev(f_xv)
        mom = ops.moments(f_d.contiguous(), v_d, dv, nmom=3)          # n, int f v, int f v^2
        if operator == "lb":
            vbar = torch.zeros_like(mom[1])
            v0t_sq = mom[2]
        else:
            vbar = mom[1]
            w = torch.full((nv,), dv, dtype=torch.float64, device=f_d.device)
            w[0] = w[-1] = 0.5 * dv
            v0t_sq = (f_d * (v_d[None, :] - vbar[:, None]) ** 2 * w).sum(dim=1)
        a = nu * dt * (-v0t_sq[:, None] / dv ** 2.0 + (v_d[None, :-1] - vbar[:, None]) / 2.0 / dv)
        b = 1.0 + nu * dt * torch.ones((nx, nv), dtype=torch.float64, device=f_d.device) * (
            2.0 * v0t_sq[:, None] / dv ** 2.0)
        c = nu * dt * (-v0t_sq[:, None] / dv ** 2.0 - (v_d[None, 1:] - vbar[:, None]) / 2.0 / dv)
        return back(a, host), back(b, host), back(c, host)

    return make_arrays_for_matrix


def get_matrix_solver(nx, nv, solver_name="batched_tridiagonal"):
    """vlapy/core/collisions.py:268-289.  ``naive`` (dense LAPACK loop in the reference) is accepted
    and routed to the same batched solver.  General (a, b, c) inputs are solved with a batched
    Thomas recurrence expressed in torch ops -- a convenience path, not the hot path (the hot path
    never materialises the diagonals, see get_collision_operator)."""
    if solver_name not in ("naive", "batched_tridiagonal"):
        raise NotImplementedError(
            "Matrix Solver: <" + solver_name + "> has not yet been implemented on the b200 backend")

    def _batched_tridiag_solver_(a, b, c, f):
        a_d, host = to_dev(a)
        b_d, _ = to_dev(b)
        c_d, _ = to_dev(c)
        d_d, _ = to_dev(f)
        bc, dc = b_d.clone(), d_d.clone()
        for it in range(1, nv):
            mc = a_d[:, it - 1] / bc[:, it - 1]
            bc[:, it] = bc[:, it] - mc * c_d[:, it - 1]
            dc[:, it] = dc[:, it] - mc * dc[:, it - 1]
        xc = bc
        xc[:, -1] = dc[:, -1] / bc[:, -1]
        for il in range(nv - 2, -1, -1):
            xc[:, il] = (dc[:, il] - c_d[:, il] * xc[:, il + 1]) / bc[:, il]
        return back(xc, host)

    return _batched_tridiag_solver_
