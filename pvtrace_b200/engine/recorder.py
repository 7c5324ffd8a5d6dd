"""Recorder (tally) specifications attached to scene nodes.

Same public classes, property ids and event-selector ids as pvtrace/engine/recorder.py:33-117: the ids are part
of the table contract shared with the device code (include/pvtrace_b200.h, PVT_REC_*).  A recorder counts rays
interacting with its node in one particular way; memory is O(bins), not O(rays).  Counts, moments and histograms
are per DISTINCT ray (first matching interaction only); raw crossings are counted separately.
"""

# histogrammable ray properties; x/y/z are in the frame of the node owning the recorder
PROPERTIES = {"wavelength": 0, "angle": 1, "duration": 2, "pathlength": 3, "x": 4, "y": 5, "z": 6}

# interaction selectors: surface (entering/escaping/reflected), volume (lost/reacted/killed), root (exit)
EVENTS = {"entering": 0, "escaping": 1, "reflected": 2, "lost": 3, "reacted": 4, "killed": 5, "exit": 6}


class Histogram:
    """1-D histogram of a ray property over [start, stop) with `bins` equal bins."""

    def __init__(self, prop, start, stop, bins):
        if prop not in PROPERTIES:
            raise ValueError(f"Unknown property {prop!r}; use one of {sorted(PROPERTIES)}")
        if not stop > start:
            raise ValueError("Histogram range requires stop > start.")
        if bins < 1:
            raise ValueError("Histogram requires at least one bin.")
        self.prop, self.start, self.stop, self.bins = prop, float(start), float(stop), int(bins)

    def __repr__(self):
        return f"Histogram({self.prop!r}, {self.start}, {self.stop}, {self.bins})"


class Heatmap:
    """2-D histogram over a pair of ray properties; ranges are (start, stop, bins)."""

    def __init__(self, prop_a, prop_b, range_a, range_b):
        self.a = Histogram(prop_a, *range_a)
        self.b = Histogram(prop_b, *range_b)

    def __repr__(self):
        return f"Heatmap({self.a!r}, {self.b!r})"


class Recorder:
    """Tally of rays interacting with a node.

    name:   key of the result in EngineResult.recorders.
    event:  one of EVENTS.
    facet:  optional outward WORLD normal; a surface recorder then only matches interactions whose surface
            normal equals it within `atol` per component.
    histograms: list of Histogram / Heatmap.
    """

    def __init__(self, name, event="entering", facet=None, atol=1e-6, histograms=None):
        if event not in EVENTS:
            raise ValueError(f"Unknown event {event!r}; use one of {sorted(EVENTS)}")
        self.name = name
        self.event = event
        self.facet = None if facet is None else tuple(float(v) for v in facet)
        self.atol = float(atol)
        self.histograms = list(histograms or [])
        if not all(isinstance(h, (Histogram, Heatmap)) for h in self.histograms):
            raise ValueError("histograms must contain Histogram or Heatmap objects.")

    def __repr__(self):
        return f"Recorder({self.name!r}, event={self.event!r})"


def auto_recorders(node_name, geometry, wavelength=(300.0, 1000.0, 100), angle=(0.0, 1.5708, 18)):
    """Default instrumentation of a node -- what `record: true` desugars to in the reference's YAML front end
    (pvtrace/cli/parse.py:469-525).  Boxes get one `escaping` recorder per face (wavelength and angle
    histograms plus a position heatmap with clamp(int(10 * size), 10, 60) bins per in-plane axis) and a volume
    `lost` recorder; spheres and cylinders get whole-surface `escaping` and `lost` recorders.
    Returns the recorders in the reference's order (lost first)."""
    from pvtrace_b200.geometry.box import Box

    made = [Recorder(f"{node_name}-lost", event="lost", histograms=[Histogram("wavelength", *wavelength)])]
    if isinstance(geometry, Box):
        size = [float(v) for v in geometry.size]
        faces = (("top", (0, 0, 1)), ("bottom", (0, 0, -1)), ("east", (1, 0, 0)), ("west", (-1, 0, 0)),
                 ("north", (0, 1, 0)), ("south", (0, -1, 0)))
        for label, facet in faces:
            across = facet.index(1) if 1 in facet else facet.index(-1)
            u, v = (k for k in range(3) if k != across)
            spans = []
            for k in (u, v):
                spans.append((-size[k] / 2.0, size[k] / 2.0, max(10, min(60, int(size[k] * 10)))))
            made.append(Recorder(
                f"{node_name}-{label}", event="escaping", facet=facet,
                histograms=[Histogram("wavelength", *wavelength), Histogram("angle", *angle),
                            Heatmap("xyz"[u], "xyz"[v], spans[0], spans[1])]))
    else:
        made.append(Recorder(f"{node_name}-escaping", event="escaping",
                             histograms=[Histogram("wavelength", *wavelength), Histogram("angle", *angle)]))
    return made
