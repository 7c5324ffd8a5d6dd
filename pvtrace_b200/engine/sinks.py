"""Result sinks: the event log of an EngineResult written into the reference CLI's sqlite database.

The layout is the reference's `pvtrace/data/schema.sql` (tables `ray` and `event`, same columns in the same order) and
the rows are what its writer loop produces (`pvtrace/cli/main.py:39-82` write_ray / write_event, `:84-160`
monitor_queue): one `ray` row and one `event` row per logged (ray, event) pair, `throw_id` = ordinal of the thrown
ray, `event.ray_id` = rowid of the pair's `ray` row, names instead of indices, the surface normal on REFLECT / TRANSMIT
rows only.  `end_rays=True` keeps what `pvtrace/scene/scene.py:35-60` (`is_end_ray`) keeps.  Rows are produced from
the log arrays directly (no Ray objects), a few hundred thousand rows per second.
"""
import sqlite3

import numpy as np

from pvtrace_b200.light.event import Event

SCHEMA = """
CREATE TABLE ray (
    throw_id NOT NULL,
    x DOUBLE, y DOUBLE, z DOUBLE,
    i DOUBLE, j DOUBLE, k DOUBLE,
    wavelength DOUBLE,
    source TEXT,
    travelled DOUBLE,
    duration DOUBLE
);
CREATE TABLE event (
    ray_id INTEGER NOT NULL,
    kind TEXT,
    component TEXT,
    hit TEXT,
    container TEXT,
    adjacent TEXT,
    facet TEXT,
    ni DOUBLE, nj DOUBLE, nk DOUBLE,
    FOREIGN KEY(ray_id) REFERENCES ray(rowid)
);
"""

_ALWAYS_END = (Event.GENERATE, Event.NONRADIATIVE, Event.REACT, Event.KILL, Event.EXIT)


def prepare_database(dbfilepath):
    """Create the two tables (cli/main.py:30-36)."""
    connection = sqlite3.connect(dbfilepath)
    connection.executescript(SCHEMA)
    connection.commit()
    return connection


def end_ray_mask(kind, hit, container, adjacent):
    """is_end_ray (scene/scene.py:35-60) over log columns: GENERATE / NONRADIATIVE / REACT / KILL / EXIT always;
    REFLECT or TRANSMIT with hit == adjacent (reflected from / transmitted into a node); TRANSMIT with
    hit == container (escaped a node)."""
    kind = np.asarray(kind)
    keep = np.isin(kind, [int(e.value) for e in _ALWAYS_END])
    surface = (kind == Event.REFLECT.value) | (kind == Event.TRANSMIT.value)
    keep |= surface & (hit == adjacent)
    keep |= (kind == Event.TRANSMIT.value) & (hit == container)
    return keep


def write_sqlite(result, dbfilepath, first_throw_id=None, end_rays=False, connection=None):
    """Append every logged (ray, event) pair of `result` to the database at `dbfilepath` (created with the schema if
    it has no `ray` table yet).  Returns the number of pairs written.  `throw_id` is the index of the THROWN ray
    (the reference CLI numbers every throw, logged or not): `first_throw_id` + the ray's index in the bundle, with
    `first_throw_id` defaulting to the bundle's first global ray index, so bundles of `simulate_stream` continue the
    numbering and `record_every > 1` leaves gaps instead of renumbering."""
    if first_throw_id is None:
        first_throw_id = int(getattr(result, "first_index", 0))
    own = connection is None
    if own:
        connection = sqlite3.connect(dbfilepath)
    cur = connection.cursor()
    if cur.execute("SELECT name FROM sqlite_master WHERE type='table' AND name='ray'").fetchone() is None:
        cur.executescript(SCHEMA)
    d, m = result.data, result.max_events
    counts = np.asarray(d["counts"], dtype=np.int64)
    recorded = result.recorded_indices
    rows = (np.arange(len(counts))[:, None] * m + np.arange(m)[None, :])[np.arange(m)[None, :] < counts[:, None]]
    throw = np.repeat(np.arange(len(counts)), counts)
    if end_rays and len(rows):
        keep = end_ray_mask(d["kind"][rows], d["hit"][rows], d["container"][rows], d["adjacent"][rows])
        rows, throw = rows[keep], throw[keep]
    start = cur.execute("SELECT COALESCE(MAX(rowid), 0) FROM ray").fetchone()[0]
    node_names, comp_names = list(result.compiled.node_names), list(result.compiled.component_names)
    kind_names = {e.value: e.name for e in Event}

    def name(table, index):
        return table[index] if index >= 0 else None

    pos, direc, nrm = d["position"], d["direction"], d["normal"]
    ray_rows, event_rows = [], []
    for k, (row, j) in enumerate(zip(rows.tolist(), throw.tolist())):
        src = int(d["source"][row])
        source = result.sources[int(recorded[j])] if src < 0 else comp_names[src]
        ray_rows.append((first_throw_id + int(recorded[j]), float(pos[row, 0]), float(pos[row, 1]), float(pos[row, 2]),
                         float(direc[row, 0]), float(direc[row, 1]), float(direc[row, 2]), float(d["wavelength"][row]),
                         source, float(d["travelled"][row]), float(d["duration"][row])))
        kind = int(d["kind"][row])
        normal = (None, None, None)
        if kind in (Event.REFLECT.value, Event.TRANSMIT.value):
            normal = (float(nrm[row, 0]), float(nrm[row, 1]), float(nrm[row, 2]))
        event_rows.append((start + k + 1, kind_names[kind], name(comp_names, int(d["component"][row])),
                           name(node_names, int(d["hit"][row])), name(node_names, int(d["container"][row])),
                           name(node_names, int(d["adjacent"][row])), None) + normal)
    cur.executemany("INSERT INTO ray VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?)", ray_rows)
    cur.executemany("INSERT INTO event VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)", event_rows)
    connection.commit()
    if own:
        connection.close()
    return len(ray_rows)
