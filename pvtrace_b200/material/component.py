"""Volume components of a host material: Scatterer, Absorber, Reactor, Luminophore.

Constructor signatures and attribute names (`_abs_dist`, `_ems_dist`, `quantum_yield`, `tau_rad`, `tau_nr`,
`phase_function`) follow pvtrace/material/component.py:33-440; the flattener lowers them into the component and
spectrum tables (pvtrace/engine/compiler.py:251-326) and the device code does the sampling.  The host methods
(`coefficient`, `is_radiative`, `emit`, ...) are provided for API parity and for small host-side checks.
"""
import math
from dataclasses import replace

import numpy as np

from pvtrace_b200.material.distribution import Distribution
from pvtrace_b200.material.utils import gaussian, isotropic

q = 1.60217662e-19  # C
kB = 1.380649e-23 / q  # eV / K


def _make_distribution(values, x, hist, what):
    if values is None:
        raise ValueError(f"{what} must be specified.")
    if isinstance(values, float):
        return Distribution(x=None, y=values, hist=hist)
    if isinstance(values, np.ndarray):
        return Distribution(x=values[:, 0], y=values[:, 1], hist=hist)
    if isinstance(values, (list, tuple)):
        if x is None:
            raise ValueError("Requires `x`.")
        return Distribution.from_functions(x, values, hist=hist)
    raise ValueError(f"{what} has wrong type.")


class Component(object):
    """Anything that can be dissolved in a host material."""

    def __init__(self, name: str = "Component"):
        self.name = name

    def is_radiative(self, ray):
        return False

    def nonradiative_absorb(self, ray):
        return ray


class Scatterer(Component):
    """Scattering centre with an attenuation coefficient (cm^-1), constant or tabulated against nm."""

    def __init__(self, coefficient, x=None, quantum_yield=1.0, tau_rad=None, tau_nr=None,
                 phase_function=None, hist=False, name="Scatterer"):
        super(Scatterer, self).__init__(name=name)
        self._coefficient = coefficient
        self._abs_dist = _make_distribution(coefficient, x, hist, "Coefficient")
        if tau_rad is not None and tau_nr is not None:
            qy = tau_nr / (tau_nr + tau_rad)
        elif quantum_yield is not None:
            qy = quantum_yield
        else:
            qy = math.nan
        if not np.isfinite(qy):
            raise ValueError("Specify either `quantum yield` or both `tau_rad` and `tau_nr`")
        self.quantum_yield = qy
        self.tau_rad = tau_rad
        self.tau_nr = tau_nr
        self.phase_function = isotropic if phase_function is None else phase_function

    def coefficient(self, wavelength):
        return self._abs_dist(wavelength)

    def is_radiative(self, ray):
        return np.random.uniform() < self.quantum_yield

    def nonradiative_absorb(self, ray):
        if self.tau_nr:
            return replace(ray, duration=ray.duration - math.log(1 - np.random.uniform()) * self.tau_nr)
        return ray

    def emit(self, ray, **kwargs):
        return replace(ray, direction=tuple(np.asarray(self.phase_function()).tolist()), source=self.name)


class Absorber(Scatterer):
    """Non-radiative absorber (quantum yield zero)."""

    def __init__(self, coefficient, x=None, tau_nr=None, name="Absorber", hist=False):
        super(Absorber, self).__init__(coefficient, x=x, quantum_yield=0.0, tau_nr=tau_nr, tau_rad=0.0,
                                       phase_function=None, hist=hist, name=name)

    def is_radiative(self, ray):
        return False


class Reactor(Absorber):
    """Absorber whose absorption events are photochemical reactions (Event.REACT)."""

    def __init__(self, coefficient, x=None, name="Reactor", hist=False):
        super(Reactor, self).__init__(coefficient, x=x, hist=hist, name=name)


class Luminophore(Scatterer):
    """Absorbs and re-emits with a new direction and a wavelength drawn from an emission spectrum."""

    def __init__(self, coefficient, emission=None, x=None, hist=False, quantum_yield=1.0, tau_rad=None,
                 tau_nr=None, phase_function=None, name="Luminophore"):
        super(Luminophore, self).__init__(coefficient, x=x, quantum_yield=quantum_yield, tau_rad=tau_rad,
                                          tau_nr=tau_nr, phase_function=phase_function, hist=hist, name=name)
        self._emission = emission
        if emission is None:
            self._ems_dist = Distribution.from_functions(x, [lambda v: gaussian(v, 1.0, 600.0, 40.0)], hist=hist)
        elif isinstance(emission, (np.ndarray, tuple, list)):
            self._ems_dist = _make_distribution(emission, x, hist, "Luminophore `emission` arg")
        else:
            raise ValueError("Luminophore `emission` arg has wrong type.")

    def emit(self, ray, method="kT", T=300.0, **kwargs):
        """Host restatement of the emission rule (the device applies the same rule, pvt_trace.cuh)."""
        dist = self._ems_dist
        nm = ray.wavelength
        if method == "kT":
            nm = 1240.0 / (1240.0 / nm + 1.5 * kB * T)
            p1 = dist.lookup(nm)
        elif method == "redshift":
            p1 = dist.lookup(nm)
        elif method == "full":
            p1 = 0.0
        else:
            raise NotImplementedError(method)
        wavelength = dist.sample(np.random.uniform(p1, 1.0))
        delay = -math.log(1 - np.random.uniform()) * self.tau_rad if self.tau_rad else 0.0
        return replace(ray, direction=tuple(np.asarray(self.phase_function()).tolist()), wavelength=wavelength,
                       duration=ray.duration + delay, source=self.name)
