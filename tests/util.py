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


def flip_aware_compare(zen_hpr, oracle_hpr, audio, hop, flags, hard_mask, margin_tol=2e-5):
    """Hop-by-hop comparison of a zen_b200 HPR object with the oracle that
    accounts for hard-mask threshold flips.

    The hard mask is a comparison `a / (b + eps) >= beta` (libzen/hps.h:100-113).
    Our FFT and the oracle's differ in the last bits, so a bin whose ratio sits
    within rounding of beta can land on the other side: a whole bin then appears
    in one output and not the other, which is not a numerical error of either.
    For every hop we therefore compare the masks first; a mismatching bin must
    be borderline in the oracle (relative margin <= margin_tol), and the audio
    of that hop and the next (overlap-add tail) is excluded from the tolerance
    check.  Returns dict(per-hop errors, flips, hops_checked).
    """
    import numpy as np
    n_hops = audio.size // hop
    W, lag, nfft = oracle_hpr.stft_width, oracle_hpr.lag, oracle_hpr.nfft
    row = W - lag
    beta = np.float32(oracle_hpr_beta(oracle_hpr))
    eps = np.float32(np.finfo(np.float32).eps)
    got = [np.zeros(n_hops * hop, np.float32) for _ in range(3)]
    ref = [np.zeros(n_hops * hop, np.float32) for _ in range(3)]
    flips = np.zeros(n_hops, dtype=np.int64)
    worst_margin = 0.0
    import torch
    a_dev = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).cuda()
    tmp = [torch.zeros(hop, dtype=torch.float32, device="cuda") for _ in range(3)]
    for i in range(n_hops):
        oracle_hpr.process_next_hop(audio[i * hop:(i + 1) * hop])
        zen_hpr.process_hop_io(a_dev[i * hop:].data_ptr(), tmp[0].data_ptr(), tmp[1].data_ptr(), tmp[2].data_ptr())
        zen_hpr.synchronize()
        for o, nm in enumerate(("harmonic_out", "percussive_out", "residual_out")):
            ref[o][i * hop:(i + 1) * hop] = oracle_hpr.get(nm)[:hop]
            got[o][i * hop:(i + 1) * hop] = tmp[o].cpu().numpy()
        if hard_mask:
            m = zen_hpr.materialize()
            H, P = oracle_hpr.get("harmonic_matrix")[row], oracle_hpr.get("percussive_matrix")[row]
            with np.errstate(all="ignore"):
                rp = P / (H + eps)
                rh = H / (P + eps)
            bad = np.zeros(nfft, dtype=bool)
            if flags & 2:
                d = m["percussive_mask"][row] != oracle_hpr.get("percussive_mask")[row]
                if d.any():
                    worst_margin = max(worst_margin, float(np.max(np.abs(rp[d] - beta) / beta)))
                bad |= d
            if flags & 1:
                d = m["harmonic_mask"][row] != oracle_hpr.get("harmonic_mask")[row]
                if d.any():
                    worst_margin = max(worst_margin, float(np.max(np.abs(rh[d] - (beta - eps)) / beta)))
                bad |= d
            flips[i] = int(bad.sum())
    peaks = [max(float(np.abs(r).max()), 1e-30) for r in ref]
    clean = np.ones(n_hops, dtype=bool)
    for i in np.nonzero(flips)[0]:
        clean[i] = False
        if i + 1 < n_hops:
            clean[i + 1] = False
    errs, snrs = [], []
    for o in range(3):
        g = got[o].reshape(n_hops, hop)[clean]
        r = ref[o].reshape(n_hops, hop)[clean]
        e = float(np.abs(g - r).max() / peaks[o]) if g.size else 0.0
        den = float(np.sum((g.astype(np.float64) - r) ** 2))
        num = float(np.sum(r.astype(np.float64) ** 2))
        errs.append(e)
        snrs.append(float("inf") if den == 0 else 10 * np.log10(max(num, 1e-300) / den))
    return dict(err=errs, snr=snrs, flips=flips, worst_margin=worst_margin, hops_checked=int(clean.sum()),
                margin_ok=worst_margin <= margin_tol, got=got, ref=ref)


def oracle_hpr_beta(o):
    return getattr(o, "beta", None) if getattr(o, "beta", None) is not None else o._beta
