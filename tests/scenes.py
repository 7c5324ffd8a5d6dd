"""Test scenes built with pvtrace_b200's host classes.  Each is the mirror of the scene of the same name that
tests/golden/make_golden.py builds with the REFERENCE's classes (and, for fresnel/lsc, of the helper scenes of the
reference's own tests/test_engine.py:36-96), so flattened tables and traces can be compared one to one."""
import functools

import numpy as np

from pvtrace_b200 import (Absorber, Box, Cylinder, Light, Luminophore, Material, Node, NullSurfaceDelegate, Reactor,
                          Scatterer, Scene, Sphere, Surface, cone)
from pvtrace_b200.data import lumogen_f_red_305
from pvtrace_b200.engine import Heatmap, Histogram, Recorder
from pvtrace_b200.material.utils import Cone, HenyeyGreenstein, gaussian


def fresnel():
    world = Node(name="world", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    box = Node(name="box", geometry=Box((1.0, 1.0, 1.0), material=Material(refractive_index=1.5)), parent=world)
    box.location = (0.0, 0.0, 2.0)
    Node(name="light", light=Light(direction=functools.partial(cone, np.pi / 16)), parent=world)
    world.recorders = [Recorder("exit", event="exit", histograms=[Histogram("angle", 0.0, 1.6, 16)])]
    box.recorders = [Recorder("in", event="entering", histograms=[Heatmap("x", "y", (-0.5, 0.5, 8), (-0.5, 0.5, 8))]),
                     Recorder("out-top", event="escaping", facet=(0, 0, 1)),
                     Recorder("refl", event="reflected")]
    return Scene(world)


def lsc():
    x = np.linspace(300.0, 1000.0, 200)
    absorption = np.column_stack((x, 5.0 * gaussian(x, 1.0, 480.0, 40.0)))
    emission = np.column_stack((x, gaussian(x, 1.0, 600.0, 40.0)))
    world = Node(name="world", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    slab = Node(name="slab", parent=world, geometry=Box((5.0, 5.0, 1.0), material=Material(
        refractive_index=1.5, components=[
            Luminophore(coefficient=absorption, emission=emission, quantum_yield=0.9, name="dye"),
            Absorber(coefficient=0.3, name="background")])))
    light = Node(name="light", light=Light(), parent=world)
    light.location = (0.0, 0.0, -3.0)
    world.recorders = [Recorder("exit", event="exit", histograms=[Histogram("wavelength", 300.0, 1000.0, 70)])]
    slab.recorders = [Recorder("lost", event="lost", histograms=[Histogram("pathlength", 0.0, 20.0, 40)]),
                      Recorder("edge+x", event="escaping", facet=(1, 0, 0),
                               histograms=[Heatmap("y", "z", (-2.5, 2.5, 10), (-0.5, 0.5, 4))]),
                      Recorder("top", event="escaping", facet=(0, 0, 1), histograms=[Histogram("angle", 0.0, 1.6, 16)]),
                      Recorder("entering", event="entering")]
    return Scene(world)


def mixed():
    x = np.linspace(350.0, 900.0, 111)
    world = Node(name="world", geometry=Box((30.0, 30.0, 30.0), material=Material(refractive_index=1.0)))
    cyl = Node(name="cyl", parent=world, geometry=Cylinder(length=3.0, radius=1.0, material=Material(
        refractive_index=1.4, components=[
            Scatterer(coefficient=0.4, phase_function=HenyeyGreenstein(0.6), name="hg"),
            Luminophore(coefficient=np.column_stack((x, 2.0 * gaussian(x, 1.0, 500.0, 50.0))),
                        emission=np.column_stack((x, gaussian(x, 1.0, 620.0, 35.0))), quantum_yield=0.8,
                        phase_function=Cone(0.5), name="lum")])))
    cyl.translate((0.5, -0.3, 4.0))
    cyl.rotate(0.7, (1.0, 0.3, 0.0))
    ball = Node(name="ball", parent=world, geometry=Sphere(radius=1.2, material=Material(
        refractive_index=1.6, components=[Reactor(coefficient=0.5, name="react"),
                                          Absorber(coefficient=0.2, name="abs")])))
    ball.translate((-2.0, 1.0, 6.5))
    ghost = Node(name="ghost", parent=world, geometry=Box((2.0, 2.0, 0.5), material=Material(
        refractive_index=1.0, surface=Surface(delegate=NullSurfaceDelegate()),
        components=[Scatterer(coefficient=1.0, quantum_yield=0.7, name="fog")])))
    ghost.translate((1.0, 1.5, 2.0))
    ghost.rotate(0.4, (0.0, 1.0, 0.2))
    Node(name="light", parent=world, light=Light(direction=functools.partial(cone, 0.35)))
    world.recorders = [Recorder("exit", event="exit"), Recorder("killed", event="killed")]
    cyl.recorders = [Recorder("cyl-in", event="entering"), Recorder("cyl-out", event="escaping"),
                     Recorder("cyl-lost", event="lost")]
    ball.recorders = [Recorder("ball-react", event="reacted"), Recorder("ball-refl", event="reflected")]
    ghost.recorders = [Recorder("ghost-in", event="entering"), Recorder("ghost-lost", event="lost")]
    return Scene(world)


def hello_world():
    from pvtrace_b200.device import configs

    return configs.hello_world(record=False)


def nested_cylinders():
    from pvtrace_b200.device import configs

    return configs.nested_cylinders(record=False)


def lsc_device():
    """LSC((5,5,1)) default with a plain Fresnel surface: what the reference engine can compile of it."""
    x = np.arange(400, 800)
    world = Node(name="World", geometry=Box((500.0, 500.0, 100.0), material=Material(refractive_index=1.0)))
    Node(name="LSC", parent=world, geometry=Box((5.0, 5.0, 1.0), material=Material(refractive_index=1.5, components=[
        Luminophore(np.column_stack((x, lumogen_f_red_305.absorption(x) * 10.0)),
                    emission=np.column_stack((x, lumogen_f_red_305.emission(x))), quantum_yield=1.0,
                    phase_function=None, name="Lumogen F Red 305"),
        Absorber(0.1, name="Background")])))
    light = Node(name="Light", parent=world, light=Light(name="Light", direction=functools.partial(cone, np.radians(20))))
    light.location = (0.0, 0.0, 5.0)
    light.rotate(np.radians(180), (1, 0, 0))
    return Scene(world)


SCENES = {"fresnel": fresnel, "lsc": lsc, "mixed": mixed, "hello_world": hello_world,
          "nested_cylinders": nested_cylinders, "lsc_device": lsc_device}
GOLDEN_METHODS = {"fresnel": ["kT"], "lsc": ["kT", "redshift", "full"], "mixed": ["kT", "redshift", "full"],
                  "hello_world": ["kT"], "nested_cylinders": ["kT"], "lsc_device": ["kT"]}
TABLES = ("geom_type geom_params local_to_world world_to_local refractive_index surface_type comp_start comp_count "
          "comp_type comp_qy comp_tau_rad comp_tau_nr comp_phase_type comp_phase_param comp_abs_start comp_abs_n "
          "comp_ems_start comp_ems_n abs_x abs_y ems_x ems_cdf rec_node rec_event rec_has_facet rec_facet rec_atol "
          "rec_hist_start rec_hist_n hist_prop_a hist_prop_b hist_na hist_nb hist_lo_a hist_hi_a hist_lo_b hist_hi_b "
          "hist_offset").split()
