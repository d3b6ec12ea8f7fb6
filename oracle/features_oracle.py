"""NumPy oracle of ``_kinematic_features`` (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates /root/reference/openglottal/features.py:38-68 step by step; the only change is that
the O(n^2) ``np.correlate(..., "full")`` (features.py:55) is replaced by the 50 lagged dot
products that features.py:56-58 actually reads, so the oracle stays usable at n = 10^6.
``exact_correlate=True`` uses the reference's np.correlate call verbatim (small n).
"""
from __future__ import annotations

import numpy as np


def kinematic_features(area_wave, exact_correlate: bool = False) -> dict | None:
    area = np.array(area_wave, dtype=np.float64)          # features.py:44
    if area.max() == 0:                                    # features.py:45-46
        return None
    mean_a = area.mean()                                   # features.py:47
    std_a = area.std()                                     # features.py:48
    oq = float(np.mean(area > mean_a * 0.1))               # features.py:49
    d = area - mean_a
    fft = np.abs(np.fft.rfft(d))                           # features.py:50
    freqs = np.fft.rfftfreq(len(area))                     # features.py:51
    peak_idx = int(np.argmax(fft[1:]) + 1)                 # features.py:52 (raises at n == 1)
    f0 = None if peak_idx == 1 else float(freqs[peak_idx])  # features.py:53-54
    n = len(area)
    if exact_correlate:
        ac = np.correlate(d, d, mode="full")               # features.py:55
        ac = ac[len(ac) // 2:]                             # features.py:56
    else:
        kmax = min(50, n)
        ac = np.array([np.dot(d[: n - k], d[k:]) for k in range(kmax)])
    ac = ac / (ac[0] + 1e-8)                               # features.py:57
    periodicity = float(ac[1: min(50, len(ac))].max())     # features.py:58 (lags 1..49)
    return {
        "area_mean": mean_a,
        "area_std": std_a,
        "area_range": area.max() - area.min(),
        "open_quotient": oq,
        "f0": f0,
        "periodicity": periodicity,
        "cv": std_a / (mean_a + 1e-8),
        "_area": area,
    }
