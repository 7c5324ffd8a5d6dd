"""Lumogen F Red 305 model spectra.

The numbers are the published fit parameters the reference ships (pvtrace/data/lumogen_f_red_305.py:4-75):
five Gaussians for the absorption coefficient (normalised to unit peak over the sampled range) and a single
Gaussian for the emission line shape.  They are data, restated so that the LSC configs reproduce bit-identical
spectrum tables.
"""
import numpy as np

# (amplitude, centre nm, width nm)
_ABSORPTION_TERMS = (
    (0.9454846839252642, 578.6167306868869, 22.69760939870020),
    (0.6430326869158796, 535.1850303736512, 28.63029894331116),
    (0.1243340609168971, 494.5721783546976, 13.98438275367119),
    (0.3651471532322375, 440.4679754085741, 34.91923613222621),
    (0.7042787252835550, 336.0548556730901, 34.24136755250487),
)


def absorption(x):
    """Absorption coefficient spectrum at wavelengths `x` (nm), peak normalised to 1."""
    total = None
    for amplitude, centre, width in _ABSORPTION_TERMS:
        term = amplitude * np.exp(-(((centre - x) / width) ** 2))
        total = term if total is None else total + term
    return total / np.max(total)


def emission(x):
    """Emission line shape at wavelengths `x` (nm), peak 1 at 600 nm."""
    return 1.0 * np.exp(-(((600.0 - x) / 38.60) ** 2))
