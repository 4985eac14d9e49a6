"""Body update with the reference's names (jax_ib/base/particle_motion.py:6-66)."""
from __future__ import annotations

import dataclasses

import numpy as np
import torch

from . import _probe, kinematics


def _moved(all_variables, dt):
    particles = all_variables.particles
    t = np.float32(all_variables.velocity[0].bc.time_stamp)
    tt = torch.tensor([t], dtype=torch.float32)
    _, rates = kinematics._values_and_rates(particles.Displacement_EQ, list(particles.displacement_param), tt)
    U0 = rates.reshape(2, -1).numpy().astype(np.float32)
    c = np.asarray(particles.particle_center, dtype=np.float32).reshape(-1, 2)
    new_c = np.stack([c[:, 0] + np.float32(dt) * U0[0], c[:, 1] + np.float32(dt) * U0[1]], axis=1).astype(np.float32)
    return dataclasses.replace(particles, particle_center=new_c)


def Update_particle_position_Multiple(all_variables, dt):
    """particle_motion.py:38-66: centre += dt * d/dt Displacement_EQ(t); Step_count += 1."""
    if _probe.is_probe(all_variables):
        return _probe.Tag("position", "kinematic", None)
    return dataclasses.replace(all_variables, particles=_moved(all_variables, dt), Step_count=all_variables.Step_count + 1)


def Update_particle_position_Multiple_and_MD_Step(step_fn, all_variables, dt):
    """particle_motion.py:6-35: as above, plus MD_var = step_fn(all_variables) (tracer sub-step)."""
    if _probe.is_probe(all_variables):
        return _probe.Tag("position", "kinematic", step_fn)
    md = step_fn(all_variables)
    return dataclasses.replace(all_variables, particles=_moved(all_variables, dt), Step_count=all_variables.Step_count + 1,
                               MD_var=md)
