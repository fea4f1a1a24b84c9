"""Mirror of vlapy/core/collisions.py for the b200 backend.

The reference materialises the three diagonals (vlapy/core/collisions.py:44-81, 104-158) and
sweeps them with a Thomas loop (:232-263).  On the device the diagonals are two scalars per row
and the solve is fused with the moment computation (csrc/rowops.h FpProg), so the product path
is ``get_collision_operator`` below; ``get_batched_array_maker`` / ``get_matrix_solver`` keep the
reference's two-stage interface for callers that want (a, b, c) explicitly.
"""
from .. import ops
from .._util import back, const, to_dev


def get_collision_operator(vax, nv, nx, nu, dt, dv, operator="lb"):
    """Fused matrix build + solve: f -> f_new. Optionally also returns the 8 moments of f_new."""
    if operator not in ops.FP_OPS:
        raise NotImplementedError(
            "Collision Operator: <" + str(operator) + "> has not yet been implemented on the b200 backend")
    v_d = const(vax)
    vgrid = ops.linspace_params(vax)      # np.linspace grids get the specialised kernel

    def collide(f_xv, moments_out=None, out=None):
        return ops.fp_step(f_xv, v_d, nu, dt, dv, operator, out=out, moments_out=moments_out, vgrid=vgrid)

    return collide


def get_batched_array_maker(vax, nv, nx, nu, dt, dv, operator="lb"):
    """vlapy/core/collisions.py:292-317 -- f -> (a, b, c) diagonals (device tensors or numpy), one kernel
    (csrc/tridiag.h DiagProg: row moments, then the diagonals with the reference's association)."""
    if operator not in ops.FP_OPS:
        raise NotImplementedError(
            "Collision Operator: <" + str(operator) + "> has not yet been implemented on the b200 backend")
    v_d = const(vax)

    def make_arrays_for_matrix(f_xv):
        f_d, host = to_dev(f_xv)
        a, b, c = ops.fp_diagonals(f_d.contiguous(), v_d, nu, dt, dv, operator)
        return back(a, host), back(b, host), back(c, host)

    return make_arrays_for_matrix


def get_matrix_solver(nx, nv, solver_name="batched_tridiagonal"):
    """vlapy/core/collisions.py:268-289.  ``naive`` (dense LAPACK loop in the reference) is accepted
    and routed to the same batched solver: one kernel for general (a, b, c) (csrc/tridiag.h TridiagProg,
    partition method + cyclic reduction; the hot path never materialises the diagonals, see
    get_collision_operator).  Functional like the reference: the arguments are not modified."""
    if solver_name not in ("naive", "batched_tridiagonal"):
        raise NotImplementedError(
            "Matrix Solver: <" + solver_name + "> has not yet been implemented on the b200 backend")

    def _batched_tridiag_solver_(a, b, c, f):
        a_d, host = to_dev(a)
        b_d, _ = to_dev(b)
        c_d, _ = to_dev(c)
        d_d, _ = to_dev(f)
        return back(ops.tridiag_solve(a_d.contiguous(), b_d.contiguous(), c_d.contiguous(), d_d.contiguous()), host)

    return _batched_tridiag_solver_
