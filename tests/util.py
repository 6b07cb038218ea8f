import numpy as np


def peak_norm_err(a, b):
    """max-abs error and SNR (dB) after peak normalisation by the reference's
    peak, the scale the CLI writes (zen/offline.h:182-191)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    pk = np.abs(b).max()
    if pk == 0:
        return float(np.abs(a).max()), float("inf") if np.abs(a).max() == 0 else -float("inf")
    err = np.abs(a - b).max() / pk
    den = np.sum((a - b) ** 2)
    snr = float("inf") if den == 0 else 10 * np.log10(np.sum(b ** 2) / den)
    return float(err), float(snr)
