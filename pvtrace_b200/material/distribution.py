"""Tabulated spectrum / probability distribution (API of pvtrace/material/distribution.py:8-193).

`_x`, `_y`, `_cdf` keep the reference's attribute names because the flattener reads them to build the device
spectrum tables exactly as pvtrace/engine/compiler.py:251-326 does.
"""
import numpy as np

from pvtrace_b200.geometry.utils import allinrange


class Distribution(object):
    def __init__(self, x, y, hist=False):
        self.hist = hist
        if x is None and isinstance(y, float):
            self._x, self._y = None, y  # wavelength-independent constant
            return
        x = np.asarray(x, dtype=float)
        y = np.asarray(y, dtype=float)
        if not np.all(np.diff(x) > 0):
            raise ValueError("x must be sorted and ascending.")
        if not np.isfinite(y).any():
            raise ValueError("All values of y must be finite.")
        if np.any(y < 0.0):
            raise ValueError("Distributions are like histograms all counts must be positive.")
        self._x, self._y = x, y
        self._x_range = (float(x.min()), float(x.max()))
        if hist:
            total = np.cumsum(y, dtype=float)
            self._cdf = total / total[-1]
            self._edges = np.append(x, 2 * x[-1] - x[-2])
        else:
            # trapezoid rule, normalised, with a leading zero so cdf[i] pairs with x[i]
            area = np.cumsum(0.5 * (y[:-1] + y[1:]))
            self._cdf = np.concatenate(([0.0], area / area.max()))

    def _is_constant(self):
        return self._x is None

    def __call__(self, x):
        """Linearly interpolated value at x (ValueError outside the tabulated range)."""
        if self._is_constant():
            if isinstance(x, (list, tuple, np.ndarray)):
                return np.full(len(x), self._y)
            return self._y
        if not allinrange(x, self._x_range):
            raise ValueError("x is outside data range.", {"x": x, "x_range": self._x_range})
        if self.hist:
            return self._y[np.searchsorted(self._edges[:-1], x)]
        return np.interp(x, self._x, self._y, left=np.nan, right=np.nan)

    def lookup(self, x):
        """Cumulative probability at x."""
        if not allinrange(x, self._x_range):
            raise ValueError("x is outside data range.", {"x": x, "x_range": self._x_range})
        if self.hist:
            return self._cdf[np.searchsorted(self._edges[:-1], x)]
        p = np.interp(x, self._x, self._cdf, left=np.nan, right=np.nan)
        return p.tolist() if p.size == 1 else p

    def sample(self, p):
        """Inverse CDF: the x value at cumulative probability p in [0, 1]."""
        if not allinrange(p, (0.0, 1.0)):
            raise ValueError("p is outside valid range.")
        if self.hist:
            idx = np.minimum(np.searchsorted(self._cdf, p), len(self._x) - 1)
            return self._x[idx]
        x = np.interp(p, self._cdf, self._x, left=np.nan, right=np.nan)
        return x.tolist() if x.size == 1 else x

    @classmethod
    def from_functions(cls, x, callables, hist=False):
        x = np.array(x, dtype=float)
        if x.ndim != 1:
            raise ValueError("Requires a 1D array.")
        y = np.zeros(len(x))
        for f in callables:
            part = np.asarray(f(x), dtype=float)
            y += np.where(np.isfinite(part), part, 0.0)
        return cls(x=x, y=y, hist=hist)
