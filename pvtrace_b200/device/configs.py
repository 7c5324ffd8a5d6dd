"""The five benchmark/parity scenes of BASELINE.json `configs`, built with this package's host API.

1. hello_world        examples/hello_world.py:8-32           glass sphere in an air sphere, cone(pi/8)
2. lsc_default        LSC((5,5,1)), pvtrace/device/lsc.py:95-219      Lumogen F Red 305 + background, cone(20 deg)
3. nested_cylinders   examples/nested_cylinders.py:21-63     rotated cylinder with a protruding child cylinder
4. lsc_coated         config 2 + edge solar cells + back-surface mirror (lsc.py:280-291)
5. validation         examples/Validation.ipynb cell 6 == tests/test_3D_flux_comparison.py:22-64
                      (the notebook's lambda delegates are expressed with the equivalent built-in delegates so the
                      light is sampled on the device)
`record=True` attaches the `record: true` auto recorders of the reference's YAML front end (cli/parse.py:469-525)
to the LSC node plus an `exit` recorder on the world.
"""
import functools

import numpy as np

from pvtrace_b200.data import fluro_red
from pvtrace_b200.device.lsc import LSC
from pvtrace_b200.engine.recorder import Recorder, auto_recorders
from pvtrace_b200.geometry.cylinder import Cylinder
from pvtrace_b200.geometry.sphere import Sphere
from pvtrace_b200.light.light import Light, SpectrumWavelengthMask, rectangular_mask
from pvtrace_b200.material.distribution import Distribution
from pvtrace_b200.material.material import Material
from pvtrace_b200.material.utils import cone
from pvtrace_b200.scene.node import Node
from pvtrace_b200.scene.scene import Scene


def hello_world(record=True):
    world = Node(name="world", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    ball = Node(name="ball-lens", parent=world, geometry=Sphere(radius=1.0, material=Material(refractive_index=1.5)))
    ball.location = (0, 0, 2)
    Node(name="green-laser", parent=world, light=Light(direction=functools.partial(cone, np.pi / 8), name="green-laser"))
    if record:
        world.recorders.append(Recorder("exit", event="exit"))
        ball.recorders.extend([Recorder("ball-entering", event="entering"), Recorder("ball-escaping", event="escaping"),
                               Recorder("ball-reflected", event="reflected")])
    return Scene(world)


def _instrument_lsc(scene):
    world = scene.root
    lsc = next(n for n in world.children if n.name == "LSC")
    lsc.recorders.extend(auto_recorders("LSC", lsc.geometry))
    world.recorders.append(Recorder("exit", event="exit"))
    return scene


def lsc_default(size=(5.0, 5.0, 1.0), record=True):
    scene = LSC(size)._make_scene(record=False)
    return _instrument_lsc(scene) if record else scene


def lsc_coated(size=(5.0, 5.0, 1.0), record=True):
    lsc = LSC(size)
    lsc.add_solar_cell({"left", "right", "near", "far"})
    lsc.add_back_surface_mirror()
    scene = lsc._make_scene(record=False)
    return _instrument_lsc(scene) if record else scene


def nested_cylinders(record=True):
    world = Node(name="World", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    a = Node(name="A", parent=world, geometry=Cylinder(length=2, radius=0.5, material=Material(refractive_index=1.5)))
    a.translate((0, 0, 2))
    a.rotate(np.pi * 0.2, (0, 1, 0))
    b = Node(name="B", parent=a, geometry=Cylinder(length=2.0, radius=0.4, material=Material(refractive_index=1.5)))
    b.rotate(np.pi / 2, (1, 0, 0))
    light = Node(name="Light (555nm)", parent=world, light=Light(direction=functools.partial(cone, np.radians(30))))
    light.translate((0, 0, -1))
    if record:
        world.recorders.append(Recorder("exit", event="exit"))
        for node in (a, b):
            node.recorders.extend([Recorder(f"{node.name}-entering", event="entering"),
                                   Recorder(f"{node.name}-escaping", event="escaping"),
                                   Recorder(f"{node.name}-reflected", event="reflected")])
    return Scene(world)


def lamp_spectrum(x):
    """Two-Gaussian fit to the Oriel lamp + filter spectrum (tests/test_3D_flux_comparison.py:42-52)."""
    def g(v, a, p, w):
        return a * np.exp(-(((p - v) / w) ** 2))
    return g(x, 0.53025700136646192, 512.91400020614333, 93.491838802960473) + \
        g(x, 0.63578999789955015, 577.63100003089369, 66.031706473985736)


def validation(size=(4.8, 1.8, 0.26), record=True):
    x = np.arange(400, 801, dtype=float)
    (l, w, d) = size
    lsc = LSC(size, wavelength_range=x)
    lsc.add_luminophore("Fluro Red", np.column_stack((x, fluro_red.absorption(x) * 11.387815)),
                        np.column_stack((x, fluro_red.emission(x))), quantum_yield=0.95)
    lsc.add_absorber("PMMA", 0.02)
    lsc.add_light("Oriel Lamp + Filter", (0.0, 0.0, 0.5 * d + 0.01), rotation=(np.radians(180), (1, 0, 0)),
                  wavelength=SpectrumWavelengthMask(Distribution(x, lamp_spectrum(x))),
                  position=functools.partial(rectangular_mask, l / 2, w / 2))
    scene = lsc._make_scene(record=False)
    if record:
        _instrument_lsc(scene)
        node = next(n for n in scene.root.children if n.name == "LSC")
        node.recorders.append(Recorder("LSC-top-reflected", event="reflected", facet=(0, 0, 1)))
        node.recorders.append(Recorder("LSC-entering", event="entering"))
    return scene


CONFIGS = {
    "hello_world": (hello_world, {"emit_method": "kT"}),
    "lsc_default": (lsc_default, {"emit_method": "kT"}),
    "nested_cylinders": (nested_cylinders, {"emit_method": "kT"}),
    "lsc_coated": (lsc_coated, {"emit_method": "kT"}),
    "validation": (validation, {"emit_method": "redshift"}),
}
