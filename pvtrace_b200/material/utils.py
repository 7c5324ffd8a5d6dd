"""Optics and sampling helpers (host float64 versions).

Fresnel / Snell / mirror formulas: pvtrace/material/utils.py:8-45.  Phase functions and surface scattering
distributions: pvtrace/material/utils.py:104-186.  The device versions live in pvtrace_b200/csrc/pvt_math.cuh;
tests/test_gpu_helpers.py checks device, oracle and host against each other, against vectors generated from the
reference's functions (tests/golden/optics.npz) and against the reference's known answers
(/root/reference/tests/test_frensel_reflection.py, test_frensel_refraction.py).
"""
import math

import numpy as np

from pvtrace_b200.geometry.utils import close_to_zero, flip


def fresnel_reflectivity(angle, n1, n2):
    """Unpolarised power reflectivity at incidence `angle` (radians) going from n1 into n2."""
    if n2 < n1 and angle > math.asin(n2 / n1):
        return 1.0  # total internal reflection
    c, s = math.cos(angle), math.sin(angle)
    k = math.sqrt(1.0 - (n1 / n2 * s) ** 2)
    r_s = ((n1 * c - n2 * k) / (n1 * c + n2 * k)) ** 2
    r_p = ((n1 * k - n2 * c) / (n1 * k + n2 * c)) ** 2
    return 0.5 * (r_s + r_p)


def specular_reflection(direction, normal):
    """Mirror `direction` in the plane with `normal` (either orientation of the normal works)."""
    d = np.asarray(direction, dtype=float)
    n = np.asarray(normal, dtype=float)
    if float(n @ d) < 0.0:
        n = flip(n)
    return d - 2.0 * float(n @ d) * n


def fresnel_refraction(direction, normal, n1, n2):
    """Snell refraction of `direction` through a surface with `normal`, from index n1 into n2."""
    d = np.asarray(direction, dtype=float)
    n = np.asarray(normal, dtype=float)
    ratio = n1 / n2
    cos_i = float(d @ n)
    cos_t = math.sqrt(1.0 - ratio * ratio * (1.0 - cos_i * cos_i))
    sign = -1.0 if cos_i < 0.0 else 1.0
    return ratio * d + sign * (cos_t - sign * ratio * cos_i) * n


# line shapes --------------------------------------------------------------------------------------


def gaussian(x, c1, c2, c3):
    return c1 * np.exp(-(((c2 - x) / c3) ** 2))


def bandgap(x, cutoff, alpha):
    return (1 - np.heaviside(x - cutoff, 0.5)) * alpha


# directions ---------------------------------------------------------------------------------------


def spherical_to_cart(theta, phi, r=1):
    st = np.sin(theta)
    cart = np.column_stack((r * st * np.cos(phi), r * st * np.sin(phi), r * np.cos(theta)))
    return cart[0, :] if cart.size == 3 else cart


def isotropic():
    """Uniform direction on the unit sphere."""
    g1, g2 = np.random.uniform(0, 1, 2)
    return spherical_to_cart(math.acos(2.0 * g2 - 1.0), 2.0 * math.pi * g1)


def henyey_greenstein(g=0.0):
    """Henyey-Greenstein phase function with asymmetry g (isotropic in the g -> 0 limit)."""
    if close_to_zero(g):
        return isotropic()
    s = 2.0 * np.random.uniform(0, 1) - 1.0
    mu = (1.0 + g * g - ((1.0 - g * g) / (1.0 + g * s)) ** 2) / (2.0 * g)
    phi = 2.0 * math.pi * np.random.uniform()
    return spherical_to_cart(math.acos(mu), phi)


class HenyeyGreenstein(object):
    def __init__(self, g: float):
        self.g = float(g)

    def __call__(self):
        return henyey_greenstein(self.g)


def cone(theta_max: float):
    """Direction within a cone of half-angle `theta_max` about +z with cos-weighted polar density."""
    if np.isclose(theta_max, 0.0) or theta_max > math.pi / 2:
        raise ValueError("Expected 0 < theta_max <= pi/2")
    p1, p2 = np.random.uniform(0, 1, 2)
    return spherical_to_cart(math.asin(math.sqrt(p1) * math.sin(theta_max)), 2.0 * math.pi * p2)


class Cone(object):
    def __init__(self, theta_max: float):
        self.theta_max = float(theta_max)

    def __call__(self):
        return cone(self.theta_max)


def lambertian():
    """Lambertian direction about +z (never points into -z)."""
    p1, p2 = np.random.uniform(0, 1, 2)
    return spherical_to_cart(math.asin(math.sqrt(p1)), 2.0 * math.pi * p2)
