"""NumPy model of the ALGORITHM the CUDA kernels implement (zen_b200/csrc):
half-spectrum real FFT, one consumed row per hop instead of the full
stft_width x nfft matrix, mirrored-index frequency windows, effective masks for
the --nocopybord geometry.  tests/test_halfspec_model.py checks it against the
oracle's full-matrix restatement of the reference, which pins the index
arithmetic the kernels rely on without needing a GPU.

It is a derivation aid for tests only; the product never imports it.
"""
import numpy as np

EPS = np.float32(np.finfo(np.float32).eps)


def time_tap_ages(W, lag, l_harm, causal, copy_bord):
    """Ages (frames behind the newest) of the ring rows that enter the time
    median of the consumed row r = W - lag.  None => that row is never written
    by the reference (H stays 0).  SURVEY.md section 8(a')."""
    L = l_harm + (1 - l_harm % 2)
    mid = L // 2
    r = W - lag
    if copy_bord:
        rows = [(r - mid + t) % W for t in range(L)]
    elif causal:
        if not (L <= r < W):
            return None
        rows = [r - L + t for t in range(L)]
    else:
        if not (mid <= r < mid + W - L):
            return None
        rows = [r - mid + t for t in range(L)]
    return [W - 1 - row for row in rows]


def mirror(i, N):
    """index into the half spectrum 0..N/2 of full-spectrum bin i (any integer)"""
    i = np.mod(i, N)
    return np.where(i > N // 2, N - i, i)


def freq_medians(mag, N, l_perc, copy_bord):
    """Returns (P_fwd, P_bwd) over bins 0..N/2: the frequency median the
    reference holds at full-spectrum bins k and N-k."""
    M = N // 2
    L = l_perc + (1 - l_perc % 2)
    mid = L // 2
    k = np.arange(M + 1)
    if copy_bord:
        idx = mirror(k[:, None] - mid + np.arange(L)[None, :], N)
        P = np.sort(mag[idx], axis=1)[:, mid]
        return P, P
    idx = mirror(k[:, None] + np.arange(L)[None, :], N)
    F = np.sort(mag[idx], axis=1)[:, mid]          # F[s] = median(mag_full[s .. s+L-1])
    P_fwd = np.where(k < N - L, F, np.float32(0))   # always true for k <= M when L < M
    P_bwd = np.zeros(M + 1, dtype=np.float32)       # value at bin N-k: window N-k .. N-k+L-1
    ok = (k > L) & (k < M)                          # N-k < N-L  <=>  k > L
    P_bwd[ok] = F[k[ok] - L + 1]
    P_bwd[0] = P_fwd[0]
    P_bwd[M] = P_fwd[M]
    return P_fwd.astype(np.float32), P_bwd


def hard(x, y, beta):
    with np.errstate(all="ignore"):
        return ((x / (y + EPS)) >= np.float32(beta)).astype(np.float32)


def soft(x, y, power):
    with np.errstate(all="ignore"):
        px, py = np.power(x, np.float32(power)), np.power(y, np.float32(power))
        return (px / (px + py + EPS)).astype(np.float32)


class HalfSpecHPR:
    """Streaming model; same call surface as oracle.OracleHPR.run."""

    def __init__(self, fs, hop, beta, flags, causal, copy_bord, geom, window, cola, sse=False, soft_mask=False):
        self.hop, self.N, self.M = hop, 4 * hop, 2 * hop
        self.beta, self.flags, self.causal, self.copy_bord = np.float32(beta), flags, causal, copy_bord
        self.W, self.lag, self.l_harm, self.l_perc = geom["stft_width"], geom["lag"], geom["l_harm"], geom["l_perc"]
        self.win, self.cola = window.astype(np.float32), np.float32(cola)
        self.sse, self.soft_mask = sse, soft_mask
        self.mag_ring = np.zeros((self.W, self.M + 1), dtype=np.float32)   # slot = frame index mod W
        self.x_ring = np.zeros((self.lag, self.M + 1), dtype=np.complex64)
        self.prev = np.zeros(hop, dtype=np.float32)
        self.tails = {o: np.zeros(hop, dtype=np.float32) for o in "HPR"}
        self.i = 0
        k = np.arange(self.M + 1)
        self.tw = np.exp(-2j * np.pi * k / self.N)

    def _analysis(self, frame):
        M, N = self.M, self.N
        x = np.zeros(N, dtype=np.float32)
        x[:2 * self.hop] = frame * self.win
        z = (x[0::2] + 1j * x[1::2]).astype(np.complex64)
        Z = np.fft.fft(z.astype(np.complex128))
        Zk = np.concatenate([Z, Z[:1]])
        Zc = np.conj(Zk[::-1])
        X = 0.5 * (Zk + Zc) - 0.5j * self.tw * (Zk - Zc)
        return X.astype(np.complex64)

    def _synthesis(self, Y):
        """unnormalised inverse of the Hermitian extension of Y[0..M]; first nwin samples"""
        M = self.M
        Yk = Y.astype(np.complex128)
        Yc = np.conj(Yk[::-1])
        Zy = (Yk + Yc) + 1j * np.conj(self.tw) * (Yk - Yc)
        zy = np.fft.ifft(Zy[:M]) * M
        y = np.empty(2 * M, dtype=np.float64)
        y[0::2], y[1::2] = zy.real, zy.imag
        return y[: 2 * self.hop].astype(np.float32)

    def process_next_hop(self, cur):
        i, W, M, N = self.i, self.W, self.M, self.N
        X = self._analysis(np.concatenate([self.prev, cur]))
        self.prev = cur.copy()
        re, im = X.real.astype(np.float32), X.imag.astype(np.float32)
        mag = np.hypot(re, im).astype(np.float32)
        if self.sse:
            mag = np.power(mag, np.float32(2.0)).astype(np.float32)
        self.mag_ring[i % W] = mag
        self.x_ring[i % self.lag] = X
        jc = i - self.lag + 1                      # consumed frame
        Xc = self.x_ring[jc % self.lag] if jc >= 0 else np.zeros(M + 1, np.complex64)

        def row(age):
            j = i - age
            return self.mag_ring[j % W] if j >= 0 else np.zeros(M + 1, np.float32)

        cons = row(self.lag - 1)
        out = {}
        if not self.sse:
            ages = time_tap_ages(W, self.lag, self.l_harm, self.causal, self.copy_bord)
            if ages is None:
                H = np.zeros(M + 1, np.float32)
            else:
                H = np.sort(np.stack([row(a) for a in ages]), axis=0)[len(ages) // 2]
            Pf, Pb = freq_medians(cons, N, self.l_perc, self.copy_bord)
            mk = (lambda a, b, beta: soft(a, b, int(self.beta))) if self.soft_mask else hard
            z = np.zeros(M + 1, np.float32)
            Mp = 0.5 * (mk(Pf, H, self.beta) + mk(Pb, H, self.beta)) if self.flags & 2 else z
            Mh = 0.5 * (mk(H, Pf, self.beta - EPS) + mk(H, Pb, self.beta - EPS)) if self.flags & 1 else z
            masks = {"P": Mp if self.flags & 2 else None, "H": Mh if self.flags & 1 else None,
                     "R": (1 - (Mh + Mp)) if (self.flags & 4 and not self.soft_mask) else None}
        else:
            with np.errstate(all="ignore"):
                Lh = self.l_harm + (1 - self.l_harm % 2)
                Lp = self.l_perc + (1 - self.l_perc % 2)
                ages = time_tap_ages(W, self.lag, self.l_harm, self.causal, True)   # box filters always wrap
                rec_t = np.stack([np.float32(1.0) / row(a) for a in ages]).astype(np.float64)
                Hm = (rec_t.sum(axis=0) / Lh).astype(np.float32)
                rec = (np.float32(1.0) / cons)
                k = np.arange(M + 1)
                idx = mirror(k[:, None] - Lp // 2 + np.arange(Lp)[None, :], N)
                Pm = (rec[idx].astype(np.float64).sum(axis=1) / Lp).astype(np.float32)
                H = (np.float32(1.0) / Hm) * np.float32(self.l_harm + 1.0)
                P = (np.float32(1.0) / Pm) * np.float32(self.l_perc + 1.0)
                sm = lambda a, b: (a * a / (a * a + b * b + EPS)).astype(np.float32)  # noqa: E731
                masks = {"P": sm(P, H) if self.flags & 2 else None, "H": sm(H, P) if self.flags & 1 else None, "R": None}
        for o in "PHR":
            hop = self.hop
            if masks[o] is None:
                out[o] = np.zeros(hop, np.float32)
                continue
            with np.errstate(all="ignore"):
                y = self._synthesis(Xc * masks[o]) * self.cola
            out[o] = self.tails[o] + y[:hop]
            self.tails[o] = y[hop:].copy()
        self.i += 1
        return out

    def run(self, audio):
        n = audio.size // self.hop
        res = {o: np.zeros(n * self.hop, np.float32) for o in "HPR"}
        for t in range(n):
            o = self.process_next_hop(audio[t * self.hop:(t + 1) * self.hop])
            for k in "HPR":
                res[k][t * self.hop:(t + 1) * self.hop] = o[k]
        return res["H"], res["P"], res["R"]
