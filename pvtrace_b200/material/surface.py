"""Surface interaction delegates (names of pvtrace/material/surface.py:13-272).

On the device a surface is data, not a callback: the flattener maps `FresnelSurfaceDelegate` and
`NullSurfaceDelegate` to tags (pvtrace/engine/compiler.py:237-247) and `FacetSurfaceDelegate` (our data-driven
generalisation of the Python-only delegates in pvtrace/device/lsc.py:22-86) to a per-facet table.  The host
methods below restate the same rules for one ray and are used by the facet-table unit tests.
"""
import abc
from dataclasses import replace

import numpy as np

from pvtrace_b200.geometry.utils import angle_between, flip
from pvtrace_b200.material.utils import fresnel_reflectivity, fresnel_refraction, specular_reflection


class SurfaceDelegate(abc.ABC):
    @abc.abstractmethod
    def reflectivity(self, surface, ray, geometry, container, adjacent) -> float:
        ...

    @abc.abstractmethod
    def reflected_direction(self, surface, ray, geometry, container, adjacent):
        ...

    @abc.abstractmethod
    def transmitted_direction(self, surface, ray, geometry, container, adjacent):
        ...


class NullSurfaceDelegate(SurfaceDelegate):
    """Always transmits, without refraction (useful for counting surfaces)."""

    def reflectivity(self, surface, ray, geometry, container, adjacent):
        return 0.0

    def reflected_direction(self, surface, ray, geometry, container, adjacent):
        raise NotImplementedError("This surface delegate does not reflect.")

    def transmitted_direction(self, surface, ray, geometry, container, adjacent):
        return ray.direction


def _incident_normal(ray, geometry):
    normal = np.asarray(geometry.normal(ray.position), dtype=float)
    if float(normal @ np.asarray(ray.direction)) < 0.0:
        normal = flip(normal)
    return normal


class FresnelSurfaceDelegate(SurfaceDelegate):
    """Fresnel reflection probability, specular reflection, Snell refraction."""

    def reflectivity(self, surface, ray, geometry, container, adjacent):
        n1 = container.geometry.material.refractive_index
        n2 = adjacent.geometry.material.refractive_index
        angle = angle_between(_incident_normal(ray, geometry), np.asarray(ray.direction, dtype=float))
        return float(fresnel_reflectivity(angle, n1, n2))

    def reflected_direction(self, surface, ray, geometry, container, adjacent):
        return tuple(specular_reflection(ray.direction, geometry.normal(ray.position)).tolist())

    def transmitted_direction(self, surface, ray, geometry, container, adjacent):
        n1 = container.geometry.material.refractive_index
        n2 = adjacent.geometry.material.refractive_index
        return tuple(fresnel_refraction(ray.direction, _incident_normal(ray, geometry), n1, n2).tolist())


class Facet(object):
    """Optical override for the part of a surface whose LOCAL outward normal equals `normal`.

    reflectivity: None keeps Fresnel; a number in [0, 1] replaces it (1 = mirror, 0 = no reflection); an (n, 2) array of
                  (wavelength in nm, reflectivity) knots is a coating with a spectrum (linear interpolation, clamped to
                  the end values).
    transmit:     "refract" (Snell) or "straight" (index matched: direction unchanged).
    reflect:      "specular" or "lambertian" (about the normal, back into the side the ray came from).
    region:       None, or ((xmin, xmax), (ymin, ymax), (zmin, zmax)) in the node's local frame (None / +-inf entries
                  are unbounded): the facet covers only the points strictly inside -- a PARTIAL coating such as the
                  quarter mirror of examples/006 Coatings.ipynb cell 3.

    Where a facet states a reflectivity below 1 but no refracted ray exists (total internal reflection, transmit
    "refract"), the ray reflects: a coating does not repeal Snell's law.
    """

    def __init__(self, normal, reflectivity=None, transmit="refract", reflect="specular", atol=1e-6, region=None):
        if transmit not in ("refract", "straight") or reflect not in ("specular", "lambertian"):
            raise ValueError("transmit must be refract|straight and reflect specular|lambertian")
        self.normal = tuple(float(v) for v in normal)
        self.reflectivity = None
        self.spectrum = None  # (x, R) arrays of a tabulated reflectivity
        if reflectivity is not None and np.ndim(reflectivity) == 0:
            if not 0.0 <= float(reflectivity) <= 1.0:
                raise ValueError("reflectivity must be in [0, 1]")
            self.reflectivity = float(reflectivity)
        elif reflectivity is not None:
            table = np.asarray(reflectivity, dtype=float)
            if table.ndim != 2 or table.shape[1] != 2 or len(table) < 1 or (np.diff(table[:, 0]) < 0).any():
                raise ValueError("a reflectivity spectrum is an (n, 2) array of (wavelength, reflectivity), wavelengths ascending")
            if (table[:, 1] < 0.0).any() or (table[:, 1] > 1.0).any():
                raise ValueError("reflectivity must be in [0, 1]")
            self.spectrum = (table[:, 0].copy(), table[:, 1].copy())
        self.transmit = transmit
        self.reflect = reflect
        self.atol = float(atol)
        lo, hi = [-np.inf] * 3, [np.inf] * 3
        if region is not None:
            if len(region) != 3:
                raise ValueError("region is ((xmin, xmax), (ymin, ymax), (zmin, zmax))")
            for k, bounds in enumerate(region):
                if bounds is None:
                    continue
                lo[k] = -np.inf if bounds[0] is None else float(bounds[0])
                hi[k] = np.inf if bounds[1] is None else float(bounds[1])
        self.region = (tuple(lo), tuple(hi))

    def matches(self, normal, position=None) -> bool:
        if not all(abs(a - b) <= self.atol for a, b in zip(self.normal, normal)):
            return False
        if position is None:
            return True
        lo, hi = self.region
        return all(l < float(p) < h for l, p, h in zip(lo, position, hi))

    def reflectivity_at(self, wavelength):
        """The stated reflectivity at `wavelength`, None when the facet keeps Fresnel."""
        if self.spectrum is not None:
            return float(np.clip(np.interp(wavelength, self.spectrum[0], self.spectrum[1]), 0.0, 1.0))
        return self.reflectivity


class FacetSurfaceDelegate(FresnelSurfaceDelegate):
    """Fresnel surface with per-facet overrides; compiles to the device facet table."""

    def __init__(self, facets=None):
        super(FacetSurfaceDelegate, self).__init__()
        self._facets = [] if facets is None else list(facets)

    @property
    def facets(self):
        return list(self._facets)

    def _facet(self, ray, geometry):
        normal = geometry.normal(ray.position)
        for facet in self.facets:
            if facet.matches(normal, ray.position):
                return facet
        return None

    @staticmethod
    def _no_refracted_ray(ray, geometry, container, adjacent):
        n1 = container.geometry.material.refractive_index
        n2 = adjacent.geometry.material.refractive_index
        if not n2 < n1:
            return False
        c = float(np.clip(_incident_normal(ray, geometry) @ np.asarray(ray.direction, dtype=float), -1.0, 1.0))
        return np.sqrt(max(1.0 - c * c, 0.0)) * (n1 / n2) > 1.0

    def reflectivity(self, surface, ray, geometry, container, adjacent):
        facet = self._facet(ray, geometry)
        stated = None if facet is None else facet.reflectivity_at(ray.wavelength)
        if stated is not None:
            if facet.transmit == "refract" and self._no_refracted_ray(ray, geometry, container, adjacent):
                return 1.0
            return stated
        return super(FacetSurfaceDelegate, self).reflectivity(surface, ray, geometry, container, adjacent)

    def reflected_direction(self, surface, ray, geometry, container, adjacent):
        facet = self._facet(ray, geometry)
        if facet is not None and facet.reflect == "lambertian":
            from pvtrace_b200.material.utils import lambertian

            back = -_incident_normal(ray, geometry)  # the hemisphere the ray arrived from
            x, y, z = lambertian()  # about +z (material/utils.py:173-186)
            if back[2] < -0.9999999:
                t1, t2 = np.array((0.0, -1.0, 0.0)), np.array((-1.0, 0.0, 0.0))
            else:
                a = 1.0 / (1.0 + back[2])
                b = -back[0] * back[1] * a
                t1 = np.array((1.0 - back[0] * back[0] * a, b, -back[0]))
                t2 = np.array((b, 1.0 - back[1] * back[1] * a, -back[1]))
            return tuple((x * t1 + y * t2 + z * back).tolist())
        return super(FacetSurfaceDelegate, self).reflected_direction(surface, ray, geometry, container, adjacent)

    def transmitted_direction(self, surface, ray, geometry, container, adjacent):
        facet = self._facet(ray, geometry)
        if facet is not None and facet.transmit == "straight":
            return tuple(ray.direction)
        return super(FacetSurfaceDelegate, self).transmitted_direction(surface, ray, geometry, container, adjacent)


class Surface(object):
    """The set of things that can happen at a material's surface, decided by a delegate."""

    def __init__(self, delegate=None):
        self._delegate = FresnelSurfaceDelegate() if delegate is None else delegate

    @property
    def delegate(self):
        return self._delegate

    def is_reflected(self, ray, geometry, container, adjacent):
        r = self.delegate.reflectivity(self, ray, geometry, container, adjacent)
        if not isinstance(r, (int, float)):
            raise ValueError("Reflectivity must be a number.")
        if r == 0.0:
            return False  # no random number is consumed
        return np.random.uniform() < r

    @staticmethod
    def _checked(direction, method):
        if not isinstance(direction, tuple) or len(direction) != 3:
            raise ValueError(f"Delegate method `{method}` should return a tuple of length 3.")
        return direction

    def reflect(self, ray, geometry, container, adjacent):
        d = self.delegate.reflected_direction(self, ray, geometry, container, adjacent)
        return replace(ray, direction=self._checked(d, "reflected_direction"))

    def transmit(self, ray, geometry, container, adjacent):
        d = self.delegate.transmitted_direction(self, ray, geometry, container, adjacent)
        return replace(ray, direction=self._checked(d, "transmitted_direction"))
