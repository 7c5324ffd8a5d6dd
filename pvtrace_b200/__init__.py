"""pvtrace_b200 -- B200-native photon tracing behind pvtrace's Scene/Node/Light API and `engine.simulate()`.

Top-level names follow pvtrace/__init__.py:16-52 (minus the trimesh `Mesh` and the meshcat renderer, which are
outside the traced path).
"""
__version__ = "0.1.0"

from pvtrace_b200.algorithm import photon_tracer  # noqa: F401
from pvtrace_b200.data import fluro_red, lumogen_f_red_305  # noqa: F401
from pvtrace_b200.device.lsc import LSC  # noqa: F401
from pvtrace_b200.geometry.box import Box  # noqa: F401
from pvtrace_b200.geometry.cylinder import Cylinder  # noqa: F401
from pvtrace_b200.geometry.sphere import Sphere  # noqa: F401
from pvtrace_b200.light.event import Event  # noqa: F401
from pvtrace_b200.light.light import Light, circular_mask, cube_mask, rectangular_mask  # noqa: F401
from pvtrace_b200.light.ray import Ray  # noqa: F401
from pvtrace_b200.material.component import Absorber, Luminophore, Reactor, Scatterer  # noqa: F401
from pvtrace_b200.material.distribution import Distribution  # noqa: F401
from pvtrace_b200.material.material import Material  # noqa: F401
from pvtrace_b200.material.surface import (  # noqa: F401
    Facet,
    FacetSurfaceDelegate,
    FresnelSurfaceDelegate,
    NullSurfaceDelegate,
    Surface,
    SurfaceDelegate,
)
from pvtrace_b200.material.utils import cone, henyey_greenstein, isotropic, lambertian  # noqa: F401
from pvtrace_b200.scene.node import Node  # noqa: F401
from pvtrace_b200.scene.scene import Scene  # noqa: F401
from pvtrace_b200 import engine  # noqa: F401
