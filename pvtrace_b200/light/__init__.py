from pvtrace_b200.light.event import Event  # noqa: F401
from pvtrace_b200.light.ray import Ray  # noqa: F401
