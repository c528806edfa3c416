"""Bars-test helpers of prosper/utils/barstest.py: ground-truth dictionaries, bars data and the permutation search the
acceptance test of SURVEY 8(c) uses (learned W ~ a permutation of 10 * bars).  Host-side, O(H^2 D); same results and the
same consumption of the global np.random stream as the reference."""
import numpy as np


def generate_bars_dict(H, neg_bars=False):
    """(D, H) dictionary of H/2 horizontal and H/2 vertical bars on an R x R grid, R = H // 2 (barstest.py:8-32).
    With `neg_bars` every bar gets a random sign (one np.random.randint(2, size=H) draw)."""
    R = H // 2
    W_gt = np.zeros((R, R, H))
    idx = np.arange(R)
    W_gt[idx, :, idx] = 1.           # bar i: row i
    W_gt[:, idx, R + idx] = 1.       # bar R + i: column i
    if neg_bars:
        sign = 1 - 2 * np.random.randint(2, size=(H,))
        W_gt = sign[None, None, :] * W_gt
    return W_gt.reshape((R * R, H))


def generate_bars_data(num, size, p_bar):
    """`num` images of size x size pixels, every bar on with probability p_bar (barstest.py:34-52).  The reference draws
    np.random.random() per image and position, horizontal bar first: one (num, size, 2) draw is the same stream."""
    u = np.random.random((num, size, 2)) <= p_bar
    data = np.zeros((num, size, size))
    data[u[:, :, 0]] = 1.                                   # rows
    data.transpose(0, 2, 1)[u[:, :, 1]] = 1.                # columns
    return data.reshape(num, size * size)


def _mae_matrix(W, Wgt):
    return np.abs(Wgt.T[:, None, :] - W.T[None, :, :]).sum(-1) / W.shape[0]


def find_permutation(W, Wgt):
    """perm[i] = column of W matched to ground-truth column i by greedy assignment on the mean absolute error
    (barstest.py:56-98): repeatedly take the smallest remaining entry of the (Hgt, H) error matrix."""
    D, H = W.shape
    Dgt, Hgt = Wgt.shape
    assert D == Dgt
    assert H >= Hgt
    mae = _mae_matrix(W, Wgt)
    perm = np.zeros(Hgt, dtype=int)
    for _ in range(Hgt):
        i, j = divmod(int(np.argmin(mae)), H)
        perm[i] = j
        mae[i, :] = np.inf
        mae[:, j] = np.inf
    return perm


def find_permutation2(W, Wgt):
    """The reference's dynamic-programming variant (barstest.py:101-174), restated with its index arithmetic kept as it
    is upstream (row `hr` of the tables extends the best partial assignment of row hr - 1; the flat argmin is split with
    Hgt, which is the row length only for H == Hgt, the case the bars tests use)."""
    D, H = W.shape
    Dgt, Hgt = Wgt.shape
    assert D == Dgt
    error = _mae_matrix(W, Wgt)
    mae_tab = np.zeros((Hgt, H))
    used_tab = np.empty((Hgt, H), dtype=object)
    for ht in range(H):
        row = error[0].copy()
        row[ht] = np.inf
        best = int(np.argmin(row))
        mae_tab[0, ht] = error[0, best]
        used_tab[0, ht] = [best]
    h0 = np.arange(Hgt)
    for hr in range(1, Hgt - 1):
        for ht in range(H):
            cost = mae_tab[hr - 1][None, :] + error[hr, h0][:, None] + np.zeros((Hgt, H))
            for h1 in range(H):
                used = used_tab[hr - 1, h1]
                if ht in used:
                    cost[:, h1] = np.inf
                else:
                    cost[[u for u in used if u < Hgt], h1] = np.inf
            if ht < Hgt:
                cost[ht, :] = np.inf
            flat = int(np.argmin(cost))
            e, prev = flat // Hgt, flat % Hgt
            mae_tab[hr, ht] = cost[e, prev]
            used_tab[hr, ht] = used_tab[hr - 1, prev] + [e]
    for ht in range(H):
        mae_tab[-1, ht] = mae_tab[-2, ht] + error[-1, ht]
        used_tab[-1, ht] = used_tab[-2, ht] + [ht]
    return np.array(used_tab[-1, int(np.argmin(mae_tab[-1]))])
