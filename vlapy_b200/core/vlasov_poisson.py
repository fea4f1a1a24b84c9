"""Mirror of vlapy/core/vlasov_poisson.py: the symplectic splitting schedules.

Pure composition of the operator closures -- the sub-step sizes, their order and the times at
which the driver is evaluated follow the reference line by line (leapfrog :53-56, PEFRL :100-148
with cumulative driver times, 6th order :199-228 with NON-cumulative driver times).
"""


def get_full_leapfrog_step(vdfdx, edfdv, field_solve, dt, driver_function):
    """vlapy/core/vlasov_poisson.py:24-60 (v half step, x full step, field at t+dt, v half step)."""

    def full_leapfrog_ps_step(e, f, t):
        f = edfdv(f=f, e=e, dt=0.5 * dt)
        f = vdfdx(f=f, dt=dt)
        e = field_solve(driver_field=driver_function(t + dt), f=f)
        f = edfdv(f=f, e=e, dt=0.5 * dt)
        return e, f

    return full_leapfrog_ps_step


def get_full_pefrl_step(vdfdx, edfdv, field_solve, dt, driver_function):
    """vlapy/core/vlasov_poisson.py:63-154 (Omelyan, Mryglod & Folk 2002, 4th order)."""
    xsi = 0.1786178958448091
    lambd = -0.2123418310626054
    chi = -0.6626458266981849e-1

    def full_pefrl_ps_step(e, f, t):
        dt1 = xsi * dt
        dt2 = chi * dt
        dt3 = (1.0 - 2.0 * (chi + xsi)) * dt
        dt4 = dt2
        dt5 = dt1
        vdt1 = 0.5 * (1.0 - 2.0 * lambd) * dt
        vdt2 = lambd * dt
        vdt3 = vdt2
        vdt4 = vdt1

        f = vdfdx(f, dt1)
        e = field_solve(driver_function(t + dt1), f=f)
        f = edfdv(f, e, vdt1)
        f = vdfdx(f, dt2)
        e = field_solve(driver_function(t + dt1 + dt2), f=f)
        f = edfdv(f, e, vdt2)
        f = vdfdx(f, dt3)
        e = field_solve(driver_function(t + dt1 + dt2 + dt3), f=f)
        f = edfdv(f, e, vdt3)
        f = vdfdx(f, dt4)
        e = field_solve(driver_function(t + dt1 + dt2 + dt3 + dt4), f=f)
        f = edfdv(f, e, vdt4)
        f = vdfdx(f, dt5)
        e = field_solve(driver_function(t + dt1 + dt2 + dt3 + dt4 + dt5), f=f)
        return e, f

    return full_pefrl_ps_step


def get_6th_order_integrator(vdfdx, edfdv, field_solve, dt, driver_function):
    """vlapy/core/vlasov_poisson.py:157-232 (Casas, Crouseilles, Faou & Mehrenberger 2017)."""
    a1 = 0.168735950563437422448196
    a2 = 0.377851589220928303880766
    a3 = -0.093175079568731452657924
    b1 = 0.049086460976116245491441
    b2 = 0.264177609888976700200146
    b3 = 0.186735929134907054308413
    c1 = -0.000069728715055305084099
    c2 = -0.000625704827430047189169
    c3 = -0.002213085124045325561636
    d2 = -2.916600457689847816445691e-6
    d3 = 3.048480261700038788680723e-5
    e3 = 4.985549387875068121593988e-7

    def sixth_order_step(e, f, t):
        D1 = b1 + 2.0 * c1 * dt ** 2.0
        D2 = b2 + 2.0 * c2 * dt ** 2.0 + 4.0 * d2 * dt ** 4.0
        D3 = b3 + 2.0 * c3 * dt ** 2.0 + 4.0 * d3 * dt ** 4.0 - 8.0 * e3 * dt ** 6.0
        for Dv, ax in ((D1, a1), (D2, a2), (D3, a3), (D3, a2), (D2, a1)):
            f = edfdv(f=f, e=e, dt=Dv * dt)
            f = vdfdx(f=f, dt=ax * dt)
            e = field_solve(driver_field=driver_function(t + ax * dt), f=f)
        f = edfdv(f=f, e=e, dt=D1 * dt)
        return e, f

    return sixth_order_step


def get_time_integrator(time_integrator_name, vdfdx, edfdv, field_solver, stuff_for_time_loop):
    """vlapy/core/vlasov_poisson.py:235-280."""
    makers = {"leapfrog": get_full_leapfrog_step, "pefrl": get_full_pefrl_step,
              "h-sixth": get_6th_order_integrator}
    if time_integrator_name not in makers:
        raise NotImplementedError("df/dt : <" + time_integrator_name + "> has not yet been implemented")
    return makers[time_integrator_name](
        vdfdx=vdfdx, edfdv=edfdv, field_solve=field_solver, dt=stuff_for_time_loop["dt"],
        driver_function=stuff_for_time_loop["driver_function"])
