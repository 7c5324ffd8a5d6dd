"""Golden fixture for the sqlite sink: histories of the UNMODIFIED reference tracer (photon_tracer.step_forward on the
hello_world scene, numpy seed 0) written by the reference's OWN writer functions -- prepare_database / write_ray /
write_event, taken verbatim (ast-extracted and exec'd) from /root/reference/pvtrace/cli/main.py, with the schema of
/root/reference/pvtrace/data/schema.sql -- and read back.  Stored next to the same histories flattened into the
engine's event-log columns, so tests/test_sinks.py can feed them to pvtrace_b200.engine.sinks and compare row for row.

Run in the build container only:   python tests/golden/make_sqlite_golden.py   ->  tests/golden/sqlite_hello_world.json
"""
import ast
import functools
import json
import os
import sqlite3
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

ref_loader.load_reference_package()

from pvtrace.algorithm import photon_tracer  # noqa: E402
from pvtrace.geometry.sphere import Sphere  # noqa: E402
from pvtrace.light.light import Light  # noqa: E402
from pvtrace.material.material import Material  # noqa: E402
from pvtrace.material.utils import cone  # noqa: E402
from pvtrace.scene.node import Node  # noqa: E402
from pvtrace.scene.scene import Scene, is_end_ray  # noqa: E402

REF = "/root/reference/pvtrace"


def reference_writer():
    """The three writer functions of cli/main.py, executed from the reference's own source text (importing the module
    would pull in typer and the meshcat renderer)."""
    tree = ast.parse(open(os.path.join(REF, "cli", "main.py")).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("prepare_database", "write_ray", "write_event")]
    assert len(wanted) == 3
    namespace = {"sqlite3": sqlite3, "SCHEMA": os.path.join(REF, "data", "schema.sql")}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), "cli/main.py", "exec"), namespace)
    return namespace


def main():
    world = Node(name="world (air)", geometry=Sphere(radius=10.0, material=Material(refractive_index=1.0)))
    ball = Node(name="sphere (glass)", parent=world, geometry=Sphere(radius=1.0, material=Material(refractive_index=1.5)))
    ball.location = (0, 0, 2)
    Node(name="Light (555nm)", parent=world, light=Light(direction=functools.partial(cone, np.pi / 8)))
    scene = Scene(world)
    np.random.seed(0)
    histories = [list(photon_tracer.step_forward(scene, ray)) for ray in scene.emit(150)]

    ref = reference_writer()
    out = {}
    for label, end_rays in (("all", False), ("end_rays", True)):
        path = os.path.join(tempfile.mkdtemp(), "golden.sqlite3")
        ref["prepare_database"](path)
        connection = sqlite3.connect(path)
        for throw, history in enumerate(histories):
            for ray, event, metadata in history:
                if end_rays and not is_end_ray(event, metadata):
                    continue
                cur = connection.cursor()
                meta = None if metadata is None else {k: (tuple(float(x) for x in v) if k == "normal" else v)
                                                      for k, v in metadata.items()}
                ray_db_id = ref["write_ray"](cur, ray, throw)
                ref["write_event"](cur, event, meta, ray_db_id)
                connection.commit()
        out[label] = {"ray": [list(r) for r in connection.execute("SELECT * FROM ray ORDER BY rowid")],
                      "event": [list(r) for r in connection.execute("SELECT * FROM event ORDER BY rowid")]}
        out["columns"] = {t: [c[1] for c in connection.execute(f"PRAGMA table_info({t})")] for t in ("ray", "event")}
        connection.close()

    # the same histories as engine log columns (node index: 0 world, 1 ball; -1 = None)
    names = [world.name, ball.name]
    max_events = max(len(h) for h in histories)
    n = len(histories)
    log = {"counts": [len(h) for h in histories], "kind": np.zeros(n * max_events, int), "hit": -np.ones(n * max_events, int),
           "container": -np.ones(n * max_events, int), "adjacent": -np.ones(n * max_events, int),
           "position": np.zeros((n * max_events, 3)), "direction": np.zeros((n * max_events, 3)),
           "normal": np.zeros((n * max_events, 3)), "wavelength": np.zeros(n * max_events),
           "travelled": np.zeros(n * max_events), "duration": np.zeros(n * max_events)}
    for j, history in enumerate(histories):
        for k, (ray, event, metadata) in enumerate(history):
            row = j * max_events + k
            log["kind"][row] = event.value
            for key in ("hit", "container", "adjacent"):
                value = (metadata or {}).get(key)
                log[key][row] = names.index(value) if value is not None else -1
            log["position"][row] = ray.position
            log["direction"][row] = ray.direction
            if metadata and "normal" in metadata:
                log["normal"][row] = metadata["normal"]
            log["wavelength"][row], log["travelled"][row], log["duration"][row] = ray.wavelength, ray.travelled, ray.duration
    out["log"] = {k: np.asarray(v).tolist() for k, v in log.items()}
    out["node_names"], out["max_events"], out["source"] = names, max_events, histories[0][0][0].source
    with open(os.path.join(HERE, "sqlite_hello_world.json"), "w") as fp:
        json.dump(out, fp)
    print("rows:", len(out["all"]["ray"]), "end rays:", len(out["end_rays"]["ray"]), "max_events", max_events)


if __name__ == "__main__":
    main()
