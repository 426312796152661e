"""Latitude-band width profile (reference: PCONV_operator/base.py:5-35).

set_weight(npart, opt) returns one weight per band: with opt=True the widths are given in units of 1/64 of
the ERP width (resampled from a 32-entry profile with monotone PCHIP interpolation), otherwise the cosine
law weight 64*cos(latitude) that the native geometry code turns into widths (extension/math_cuda.cu:223-253).
"""
import os

import numpy as np

DEFAULT_PROFILE = [8, 18, 24, 36, 46, 58, 62, 62, 62, 62, 63, 63, 63, 63, 63, 63,
                   63, 63, 63, 63, 63, 63, 62, 62, 62, 62, 58, 46, 36, 24, 18, 8]


def load_param(file_name):
    """Optional override of the 32-entry profile: one comma-separated line (base.py:5-11)."""
    if os.path.exists(file_name):
        with open(file_name) as f:
            return [int(tok) for tok in f.readline().strip().split(",")]
    return list(DEFAULT_PROFILE)


def _band_centres(n):
    return np.cos((0.5 - (np.arange(float(n)) + 0.5) / n) * np.pi)


def set_weight(npart, opt=False, merge=False, config_file="./config/param.txt"):
    assert npart % 2 == 0, "npart should be the multiplier of 2 for the merge case"
    bands = npart * 2 if merge else npart
    if opt:
        import scipy.interpolate
        profile = np.array([v + 1 for v in load_param(config_file)])
        x32, xb, half = _band_centres(32), _band_centres(bands), bands // 2
        north = np.ceil(scipy.interpolate.pchip_interpolate(x32[:16], profile[:16], xb[:half]))
        south = np.ceil(scipy.interpolate.pchip_interpolate(x32[16:][::-1], profile[16:][::-1], xb[half:]))
        weights = north.tolist() + south.tolist()
    else:
        weights = np.ceil(_band_centres(bands) * 64.0).tolist()
    if merge:
        weights = [max(weights[2 * i], weights[2 * i + 1]) for i in range(bands // 2)]
    return weights
