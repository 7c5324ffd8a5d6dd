"""The sqlite sink (pvtrace_b200/engine/sinks.py) against rows written by the reference's own writer functions
(tests/golden/sqlite_hello_world.json, made by tests/golden/make_sqlite_golden.py from /root/reference)."""
import json
import os
import sqlite3
import types

import numpy as np
import pytest

from pvtrace_b200.engine import sinks
from pvtrace_b200.engine.api import EngineResult

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sqlite_hello_world.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as fp:
        return json.load(fp)


def result_from(golden):
    log = golden["log"]
    data = {k: np.asarray(v) for k, v in log.items()}
    rows = len(data["kind"])
    data["component"] = -np.ones(rows, dtype=np.int64)
    data["source"] = -np.ones(rows, dtype=np.int64)
    compiled = types.SimpleNamespace(node_names=golden["node_names"], component_names=[], recorder_specs=[])
    n = len(data["counts"])
    return EngineResult(compiled, data, [golden["source"]] * n, golden["max_events"], 1, 0.0)


def read(path):
    connection = sqlite3.connect(path)
    try:
        return {t: [list(r) for r in connection.execute(f"SELECT * FROM {t} ORDER BY rowid")] for t in ("ray", "event")}, \
               {t: [c[1] for c in connection.execute(f"PRAGMA table_info({t})")] for t in ("ray", "event")}
    finally:
        connection.close()


@pytest.mark.parametrize("label,end_rays", [("all", False), ("end_rays", True)])
def test_rows_equal_the_reference_writers(golden, tmp_path, label, end_rays):
    path = str(tmp_path / "out.sqlite3")
    written = result_from(golden).to_sqlite(path, end_rays=end_rays)
    rows, columns = read(path)
    assert columns == golden["columns"]  # data/schema.sql: same tables, same columns, same order
    assert written == len(golden[label]["ray"]) == len(rows["ray"])
    assert rows["ray"] == golden[label]["ray"]      # floats are the same doubles: exact
    assert rows["event"] == golden[label]["event"]
    if end_rays:
        assert len(golden["end_rays"]["ray"]) < len(golden["all"]["ray"])  # the fixture does exercise the filter


def test_bundles_append_and_continue_the_numbering(golden, tmp_path):
    path = str(tmp_path / "out.sqlite3")
    result = result_from(golden)
    first = result.to_sqlite(path)
    second = result.to_sqlite(path, first_throw_id=result.num_recorded)
    rows, _ = read(path)
    assert len(rows["ray"]) == first + second == 2 * len(golden["all"]["ray"])
    assert rows["event"][first][0] == first + 1 and rows["ray"][first][0] == result.num_recorded
    assert [r[1:] for r in rows["event"][first:]] == [r[1:] for r in golden["all"]["event"]]


def test_empty_log_writes_only_the_schema(tmp_path):
    data = {"counts": np.zeros(0, dtype=np.int32)}
    for key in ("kind", "hit", "container", "adjacent", "component", "source", "wavelength", "travelled", "duration"):
        data[key] = np.zeros(0)
    for key in ("position", "direction", "normal"):
        data[key] = np.zeros((0, 3))
    compiled = types.SimpleNamespace(node_names=["world"], component_names=[], recorder_specs=[])
    result = EngineResult(compiled, data, ["Light"] * 10, 8, 0, 0.0)
    path = str(tmp_path / "out.sqlite3")
    assert result.to_sqlite(path) == 0
    rows, columns = read(path)
    assert rows == {"ray": [], "event": []} and columns["ray"][0] == "throw_id"
