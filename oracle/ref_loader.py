"""Load the UNMODIFIED reference (pvtrace) from /root/reference for validation only.

TEST INFRASTRUCTURE.  Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py may import anything under oracle/.  The product
package (pvtrace_b200) never does.

Two loaders:

* ``load_ref_kernel()`` -- imports the compiled reference kernel built by ``make -C oracle ref``
  (oracle/_ref/_kernel*.so, compiled from /root/reference/pvtrace/engine/_kernel.pyx where it lies).
  The built module travels to the GPU box with the snapshot; it only needs numpy + libgomp.
* ``load_reference_package()`` -- imports the reference *Python* package from /root/reference.  Only works
  in the build container (the GPU box has no /root/reference).  pvtrace's third-party dependencies
  ``anytree``, ``trimesh`` and ``meshcat`` are not installed here, so minimal stand-ins are registered:
  a tree mixin with the four names the reference uses (NodeMixin, Walker, PreOrderIter, PostOrderIter,
  LevelOrderIter) and a ``trimesh.creation.box`` stub (the engine never queries the mesh).  Python-tracer
  runs through Box geometry therefore are NOT available (trimesh ray queries); Sphere/Cylinder scenes are.
"""
from __future__ import annotations

import glob
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("PVTRACE_REFERENCE", "/root/reference")


def ref_kernel_path():
    hits = sorted(glob.glob(os.path.join(HERE, "_ref", "_kernel*.so")))
    return hits[0] if hits else None


def load_ref_kernel():
    """Return the compiled reference kernel module (or None when it has not been built)."""
    if "pvt_ref_kernel" in sys.modules:
        return sys.modules["pvt_ref_kernel"]
    path = ref_kernel_path()
    if path is None:
        return None
    # The extension's init symbol is PyInit__kernel, so the spec name must end in "_kernel".
    spec = importlib.util.spec_from_file_location("_kernel", path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    sys.modules["pvt_ref_kernel"] = module
    return module


# ----------------------------------------------------------------------------------------------
# Stand-ins for the reference's missing third-party dependencies (validation only).


def _install_anytree_shim():
    if "anytree" in sys.modules:
        return
    mod = types.ModuleType("anytree")

    class NodeMixin:
        _pvt_parent = None

        @property
        def parent(self):
            return self._pvt_parent

        @parent.setter
        def parent(self, value):
            old = self._pvt_parent
            if old is not None:
                old._pvt_kids.remove(self)
            self._pvt_parent = value
            if value is not None:
                if "_pvt_kids" not in value.__dict__:
                    value.__dict__["_pvt_kids"] = []
                value._pvt_kids.append(self)

        @property
        def children(self):
            return tuple(self.__dict__.get("_pvt_kids", ()))

        @property
        def path(self):
            chain, node = [], self
            while node is not None:
                chain.append(node)
                node = node.parent
            return tuple(reversed(chain))

        @property
        def leaves(self):
            return tuple(n for n in PreOrderIter(self) if not n.children)

    def PreOrderIter(root):
        stack = [root]
        while stack:
            node = stack.pop()
            yield node
            stack.extend(reversed(node.children))

    def PostOrderIter(root):
        for child in root.children:
            yield from PostOrderIter(child)
        yield root

    def LevelOrderIter(root):
        level = [root]
        while level:
            nxt = []
            for node in level:
                yield node
                nxt.extend(node.children)
            level = nxt

    class Walker:
        def walk(self, start, end):
            a, b = start.path, end.path
            k = 0
            while k < min(len(a), len(b)) and a[k] is b[k]:
                k += 1
            common = a[k - 1]
            return tuple(reversed(a[k:])), common, tuple(b[k:])

    mod.NodeMixin = NodeMixin
    mod.PreOrderIter = PreOrderIter
    mod.PostOrderIter = PostOrderIter
    mod.LevelOrderIter = LevelOrderIter
    mod.Walker = Walker
    sys.modules["anytree"] = mod


def _install_trimesh_stub():
    if "trimesh" in sys.modules:
        return
    import numpy as np

    mod = types.ModuleType("trimesh")
    creation = types.ModuleType("trimesh.creation")

    class _BoxMesh:
        def __init__(self, size):
            half = 0.5 * np.asarray(size, dtype=float)
            signs = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
            self.vertices = signs * half
            self.center_mass = np.zeros(3)
            self.extents = 2.0 * half

    creation.box = lambda extents=None, **kw: _BoxMesh(extents)
    mod.creation = creation
    mod.Trimesh = _BoxMesh
    sys.modules["trimesh"] = mod
    sys.modules["trimesh.creation"] = creation


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pvtrace"))


def load_reference_package():
    """Make ``import pvtrace.<submodule>`` resolve to /root/reference WITHOUT running pvtrace/__init__.py
    (which imports meshcat).  Returns the stub top-level package."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected in the build container only)")
    if "pvtrace" in sys.modules and getattr(sys.modules["pvtrace"], "_pvt_ref_stub", False):
        return sys.modules["pvtrace"]
    _install_anytree_shim()
    _install_trimesh_stub()
    pkg = types.ModuleType("pvtrace")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pvtrace")]
    pkg._pvt_ref_stub = True
    sys.modules["pvtrace"] = pkg
    import logging

    logging.getLogger("pvtrace").setLevel(logging.WARNING)
    kernel = load_ref_kernel()
    # import order matters: photon_tracer before scene (circular import in the reference otherwise)
    import pvtrace.algorithm.photon_tracer  # noqa: F401
    import pvtrace.scene.scene  # noqa: F401
    import pvtrace.engine  # noqa: F401

    if kernel is not None:
        sys.modules["pvtrace.engine._kernel"] = kernel
        sys.modules["pvtrace.engine"]._kernel = kernel
    return pkg
