"""Shared pieces of the NumPy oracle (test infrastructure; see oracle/__init__.py)."""
import numpy as np
from scipy.special import comb


class DictAnneal(object):
    """Minimal stand-in for prosper.em.annealing.Annealing.__getitem__ semantics
    (annealing.py:90-94: a missing key reads as 0.0)."""

    def __init__(self, **kw):
        self.values = dict(kw)
        self.crit_params = []

    def __getitem__(self, k):
        return self.values.get(k, 0.0)

    def __setitem__(self, k, v):
        self.values[k] = v

    def as_dict(self):
        return dict(self.values)


def state_sqerr(W_HD, y, cand, SM, chunk_bytes=256 << 20):
    """||SM . W[cand_n] - y_n||^2 for every (n, state): (n,S).

    Direct restatement of `Wbar = np.dot(SM, W[cand]); ((Wbar-y)**2).sum(axis=1)`
    (bsc_et.py:180-184, tsc_et.py:347-351, dsc_et.py:576-579), batched over n.
    """
    n, D = y.shape
    S = SM.shape[0]
    out = np.empty((n, S))
    SMf = np.asarray(SM, dtype=np.float64)
    step = max(1, int(chunk_bytes // max(1, S * D * 8)))
    for a in range(0, n, step):
        b = min(n, a + step)
        Wc = W_HD[cand[a:b]]                      # (m, H', D)
        Wbar = np.matmul(SMf[None, :, :], Wc)     # (m, S, D)
        Wbar -= y[a:b, None, :]
        out[a:b] = np.einsum('msd,msd->ms', Wbar, Wbar)
    return out


def single_sqerr(W_HD, y):
    """||W_h - y_n||^2 for every (n, h): (n,H); `((W-y)**2).sum(axis=1)` (bsc_et.py:176)."""
    n = y.shape[0]
    H = W_HD.shape[0]
    out = np.empty((n, H))
    step = max(1, (64 << 20) // max(1, W_HD.size * 8))
    for a in range(0, n, step):
        b = min(n, a + step)
        d = W_HD[None, :, :] - y[a:b, None, :]
        out[a:b] = np.einsum('nhd,nhd->nh', d, d)
    return out


def binom_AB(H, gamma, pies):
    """A_pi_gamma, B_pi_gamma of bsc_et.py:239-244 (also mca_et.py:241-246, mmca_et.py:269-274)."""
    A = 0.0
    B = 0.0
    for g in range(gamma + 1):
        a = comb(H, g) * (pies ** g) * ((1. - pies) ** (H - g))
        A += a
        B += g * a
    return A, B


def numpy_rcond():
    """bsc_et.py:377-380 / dsc_et.py:732-735 pick rcond by parsing np.__version__[2:]
    as a float: '1.26.4' -> 26.4 -> None; '2.3.5' -> 3.5 -> -1."""
    try:
        return None if float(np.__version__[2:]) >= 14.0 else -1
    except ValueError:
        return -1


def truncate(all_denoms, N_use_target, strict, allsort=None):
    """Data truncation rule of bsc_et.py:250-254 (`>= cut`) / dsc_et.py:830-832 (`> cut`).

    `allsort` maps the local denominators to the globally sorted array
    (parallel.py:87-110); single rank: np.sort.
    """
    srt = np.sort(all_denoms) if allsort is None else allsort(all_denoms)
    cut = srt[-N_use_target]
    return (all_denoms > cut) if strict else (all_denoms >= cut)
