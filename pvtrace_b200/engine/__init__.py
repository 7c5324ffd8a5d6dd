"""GPU tracing engine: the drop-in for `pvtrace.engine` (pvtrace/engine/__init__.py:19-44).

Same names as the reference namespace.  The tracing loop runs in hand-written sm_100a CUDA behind the C ABI of
include/pvtrace_b200.h (pvtrace_b200/csrc, loaded with ctypes by engine/_cuda.py); there is no CPU fallback.
Build the library with::

    python -m pvtrace_b200.csrc.build
"""
from pvtrace_b200.engine.compiler import (  # noqa: F401
    CompiledEmitter,
    CompiledScene,
    UnsupportedSceneError,
    compile_emitter,
    compile_scene,
)
from pvtrace_b200.engine.recorder import Heatmap, Histogram, Recorder, auto_recorders  # noqa: F401
from pvtrace_b200.engine.tally import tally_histories  # noqa: F401
from pvtrace_b200.engine.api import (  # noqa: F401
    EngineResult,
    RecorderResult,
    is_available,
    simulate,
    simulate_stream,
)

__all__ = [
    "CompiledScene", "UnsupportedSceneError", "compile_scene", "Recorder", "Histogram", "Heatmap", "EngineResult",
    "RecorderResult", "is_available", "simulate", "simulate_stream", "tally_histories",
    "CompiledEmitter", "compile_emitter", "auto_recorders",
]
