"""Coumarin "Fluro Red" model spectra used by the validation scene.

Fit parameters as shipped by the reference (pvtrace/data/fluro_red.py:4-93): four Gaussians for absorption
(peak normalised) and an exponentially modified Gaussian for emission.
"""
import numpy as np
from scipy.special import erf

# (centre nm, amplitude, width nm)
_ABSORPTION_TERMS = (
    (549.06438843562137, 439.06754804626956, 24.298601639828647),
    (379.48645797468572, 85.177292848284353, 13.513987279089216),
    (519.58858977131513, 660.1731296017241, 38.263352007649125),
    (490.05625608592726, 511.11501615291041, 52.213294432464529),
)
_EMG = (1.1477763237584664, 592.06478874548839, 19.981040318195117, 12.723704058786568)


def absorption(x):
    total = None
    for centre, amplitude, width in _ABSORPTION_TERMS:
        term = amplitude * np.exp(-(((centre - x) / width) ** 2))
        total = term if total is None else total + term
    return total / np.max(total)


def emission(x):
    a, b, c, d = _EMG
    r2 = np.sqrt(2)
    return (
        a * c * np.sqrt(2 * np.pi) / (2 * d)
        * np.exp((c ** 2 / (2 * d ** 2)) - ((x - b) / d))
        * (d / np.abs(d) + erf((x - b) / (r2 * c) - c / (r2 * d)))
    )
