"""Time-stepper selectors with the reference's names (jax_ib/base/time_stepping.py:367-387).  The
reference's steppers are forward-Euler instances of its RK driver (:109-365); here they select which
fused CUDA step `equations.semi_implicit_navier_stokes_*` builds."""
from __future__ import annotations


class _Stepper:
    def __init__(self, name, incremental_pressure):
        self.name, self.incremental_pressure = name, incremental_pressure

    def __repr__(self):
        return f"<time stepper {self.name}>"


# time_stepping.py:378-387: u* = u + dt F(u) - grad(p_old); IB force; project; p += q
forward_euler_updated = _Stepper("forward_euler_updated", True)
# time_stepping.py:367-376: u* = u + dt F(u); project (pressure is still accumulated, not used)
forward_euler_penalty = _Stepper("forward_euler_penalty", False)
