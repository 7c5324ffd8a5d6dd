"""Host material: refractive index, surface, dissolved components (pvtrace/material/material.py:10-63)."""
import math

import numpy as np

from pvtrace_b200.material.surface import Surface


class Material(object):
    def __init__(self, refractive_index, surface=None, components=None):
        self.refractive_index = refractive_index
        self.surface = Surface() if surface is None else surface
        self.components = [] if components is None else components

    def total_attenutation_coefficient(self, wavelength):  # (sic) reference spelling kept
        return float(sum(c.coefficient(wavelength) for c in self.components))

    def penetration_depth(self, wavelength):
        """Beer-Lambert free path: -ln(1 - U) / alpha, infinite when alpha is (close to) zero."""
        alpha = self.total_attenutation_coefficient(wavelength)
        if np.isclose(alpha, 0.0):
            return math.inf
        if not np.isfinite(alpha):
            return 0.0
        return -math.log(1 - np.random.uniform()) / alpha

    def is_absorbed(self, ray, full_distance):
        distance = self.penetration_depth(ray.wavelength)
        return distance < full_distance, distance

    def component(self, wavelength):
        """Pick the absorbing component with probability proportional to its coefficient."""
        coefs = np.array([c.coefficient(wavelength) for c in self.components], dtype=float)
        if np.any(coefs < 0.0):
            raise ValueError("Must be positive.")
        running = np.cumsum(coefs)
        target = np.random.uniform() * running[-1]
        return self.components[int(np.searchsorted(running, target, side="left").clip(0, len(coefs) - 1))]
