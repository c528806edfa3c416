"""Shared test helpers: synthetic inputs of the BASELINE.json shapes and parity metrics."""
import numpy as np


def bars_dict(H):
    """Ground-truth bars dictionary (D = (H/2)^2, H): same construction as the reference's
    utils/barstest.py:8-32 (horizontal then vertical bars)."""
    R = H // 2
    W = np.zeros((R, R, H))
    for i in range(R):
        W[i, :, i] = 1.
        W[:, i, R + i] = 1.
    return W.reshape(R * R, H)


def bsc_problem(D, H, N, seed, pi=None, sigma=1.0, bars=False, w_scale=10.0):
    """Synthetic BSC data + an initial parameter set (standard_init semantics)."""
    rng = np.random.RandomState(seed)
    if bars:
        Wgt = w_scale * bars_dict(H)
    else:
        Wgt = rng.standard_normal((D, H))
        Wgt *= w_scale / np.linalg.norm(Wgt, axis=0, keepdims=True)
    pi = pi if pi is not None else 2.0 / H
    s = rng.random_sample((N, H)) < pi
    y = s.astype(np.float64) @ Wgt.T + sigma * rng.standard_normal((N, D))
    W_mean = y.mean(axis=0)
    sig0 = np.sqrt(((y - W_mean) ** 2).mean(axis=0)).sum() / D
    W0 = W_mean[:, None] + rng.normal(scale=sig0 / 4., size=(D, H))
    params = {'W': W0, 'pi': 1. / H, 'sigma': sig0}
    return y, params, {'W': Wgt, 'pi': pi, 'sigma': sigma}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def cand_mismatch_gap(sim, cand_a, cand_b):
    """For rows whose candidate SETS differ, the largest score gap between the elements that
    differ (north_star: sets must be identical wherever the gap exceeds the tolerance)."""
    worst = 0.0
    bad = 0
    for n in range(sim.shape[0]):
        sa, sb = set(cand_a[n].tolist()), set(cand_b[n].tolist())
        if sa != sb:
            bad += 1
            diff = list(sa ^ sb)
            v = sim[n, diff]
            worst = max(worst, float(v.max() - v.min()))
    return bad, worst


def northstar_inputs(N=48, seed=5):
    """Inputs of the north-star shape (BASELINE configs[4]: D=676, H=1000, H'=12, gamma=5) regenerated from a seed with
    the law of SURVEY 8(d): W_gt ~ N(0,1) with columns of norm 10, s ~ Bernoulli(2/H), y = W_gt s + N(0,1), initial
    parameters with standard_init semantics.  tests/golden/northstar_bsc.npz holds what the UNMODIFIED reference
    returns for exactly these arrays (minted by tests/golden/make_golden.py northstar), so W0 is not stored twice."""
    D, H = 676, 1000
    rng = np.random.RandomState(seed)
    Wgt = rng.standard_normal((D, H))
    Wgt *= 10.0 / np.linalg.norm(Wgt, axis=0, keepdims=True)
    s = rng.random_sample((N, H)) < 2.0 / H
    y = s.astype(np.float64) @ Wgt.T + rng.standard_normal((N, D))
    W_mean = y.mean(axis=0)
    sig0 = np.sqrt(((y - W_mean) ** 2).mean(axis=0)).sum() / D
    W0 = W_mean[:, None] + rng.normal(scale=sig0 / 4., size=(D, H))
    return y, {'W': W0, 'pi': 1. / H, 'sigma': sig0}
