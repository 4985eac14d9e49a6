"""jax_ib/penalty/parameteric_fns.py:3-26: polar shape functions R(theta)."""
import numpy as np


def param_ellipse(geometry_param, theta):
    A, B = geometry_param[0], geometry_param[1]
    return A * B / np.sqrt((B * np.cos(theta)) ** 2 + (A * np.sin(theta)) ** 2)


def param_rose(geometry_param, theta):
    return geometry_param[0] * np.sin(geometry_param[1] * theta)


def param_snail(geometry_param, theta):
    return geometry_param[0] * theta


def param_circle(geometry_param, theta):
    return geometry_param[0] * np.ones_like(theta)


def param_rot_ellipse(phi, geometry_param, theta):
    A, B = geometry_param[0], geometry_param[1]
    excc = np.sqrt(1 - np.round((B / A) ** 2, 6))
    return B / np.sqrt(1 - (excc * np.cos(theta - phi)) ** 2)
