"""Light sources: three delegate callables (wavelength, position, direction) sampled per ray.

Public names follow pvtrace/light/light.py:48-262.  The engine recognises the built-in delegates (functions,
helper objects and functools.partial wrappers of them) and lowers them to an emission descriptor that is
sampled on the device (pvtrace_b200/engine/emit.py); anything else is called once per ray on the host.
"""
import math
from typing import Iterator

import numpy as np

from pvtrace_b200.light.ray import Ray


def default_wavelength():
    return 555.0


def default_position():
    return (0.0, 0.0, 0.0)


def default_direction():
    return (0.0, 0.0, 1.0)


def rectangular_mask(X, Y):
    """Uniform over the rectangle |x| <= X, |y| <= Y in the z = 0 plane."""
    return (np.random.uniform(-X, X), np.random.uniform(-Y, Y), 0.0)


def circular_mask(radius):
    """Uniform over the disc of `radius` in the z = 0 plane."""
    phi = np.random.uniform(0, 2.0 * math.pi)
    r = math.sqrt(np.random.uniform()) * radius
    return (r * math.cos(phi), r * math.sin(phi), 0.0)


def cube_mask(X, Y, Z):
    return (np.random.uniform(-X, X), np.random.uniform(-Y, Y), np.random.uniform(-Z, Z))


class DefaultWavelength(object):
    def __call__(self):
        return default_wavelength()


class DefaultPosition(object):
    def __call__(self):
        return default_position()


class DefaultDirection(object):
    def __call__(self):
        return default_direction()


class ConstantWavelengthMask(object):
    def __init__(self, nanometers):
        self.nanometers = float(nanometers)

    def __call__(self):
        return self.nanometers


class SpectrumWavelengthMask(object):
    """Wavelengths drawn from a `Distribution` by inverse-CDF sampling."""

    def __init__(self, distribution):
        self.distribution = distribution

    def __call__(self):
        return self.distribution.sample(np.random.uniform(0, 1))


class RectangularMask(object):
    def __init__(self, x, y):
        self.x = float(x)
        self.y = float(y)

    def __call__(self):
        return rectangular_mask(self.x, self.y)


class CircularMask(object):
    def __init__(self, radius):
        self.radius = radius

    def __call__(self):
        return circular_mask(self.radius)


class CubeMask(object):
    def __init__(self, x, y, z):
        self.x, self.y, self.z = x, y, z

    def __call__(self):
        return cube_mask(self.x, self.y, self.z)


class Light(object):
    """Emits along local +z from the local origin at 555 nm unless delegates say otherwise."""

    def __init__(self, wavelength=None, position=None, direction=None, name="Light"):
        self.wavelength = default_wavelength if wavelength is None else wavelength
        self.position = default_position if position is None else position
        self.direction = default_direction if direction is None else direction
        self.name = name

    def emit(self, num_rays=None) -> Iterator[Ray]:
        if not num_rays:
            return
        for _ in range(int(num_rays)):
            yield Ray(
                wavelength=self.wavelength(),
                position=tuple(self.position()),
                direction=tuple(self.direction()),
                source=self.name,
            )
