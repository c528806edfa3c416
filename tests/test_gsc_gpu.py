"""GPU parity tests of GSC (spike-and-slab) ET: moment tensors and parameter updates."""
import glob
import os

import numpy as np
import pytest

from helpers import bars_dict, rel_err
from oracle.common import DictAnneal
from oracle.gsc import GSC as OGSC

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-8


def cp(p):
    return dict((k, (np.copy(v) if isinstance(v, np.ndarray) else v)) for k, v in p.items())


def check(D, H, Hp, gam, stype, params, y, T, golden=None):
    from prosper_b200.em.camodels.gsc_et import GSC
    assert torch.cuda.is_available(), "GPU tests need a B200"
    an = DictAnneal(T=T)
    o = OGSC(D, H, Hp, gam, sigma_sq_type=stype)
    po = cp(params)
    od = o.select_hprimes(po, {'y': y.copy()})
    cand_o = od['candidates'].copy()
    osuff = o.e_step(an, po, od)
    onew = o.m_step(an, cp(po), osuff, od)
    m = GSC(D, H, Hp, gam, sigma_sq_type=stype)
    p1 = cp(params)
    d = m.select_Hprimes(p1, {'y': y.copy()})
    assert np.array_equal(m._cand, cand_o)
    assert sorted(len(c['ind']) for c in d['data_clusters'].values()) == sorted(np.unique(cand_o, axis=0, return_counts=True)[1].tolist())
    suff = m.E_step(an, p1, d)
    assert np.array_equal(d['y'], od['y']) and np.array_equal(d['candidates'], od['candidates'])     # cluster-major reorder
    for k in ('xpt_s', 'xpt_ss', 'xpt_sz', 'xpt_szsz'):
        assert suff[k].shape == osuff[k].shape
        assert np.abs(suff[k] - osuff[k]).max() < 1e-10 * max(1.0, np.abs(osuff[k]).max()), k
    got_c = m.M_step(an, cp(p1), osuff, d)                                  # compat M-step on the oracle's tensors
    got_f = m._fused_step(an, cp(params), {'y': y.copy()})                  # fused: nothing of size n*H*H exists
    for got in (got_c, got_f):
        for k in ('W', 'pi', 'mu', 'psi_sq', 'sigma_sq'):
            assert rel_err(got[k], onew[k]) < TOL, k
    if golden is not None:
        for k in ('W', 'pi', 'mu', 'psi_sq', 'sigma_sq'):
            assert rel_err(got_f[k], golden[k + '_new']) < TOL, k
        assert np.abs(suff['xpt_szsz'] - golden['xpt_szsz']).max() < 1e-9 * np.abs(golden['xpt_szsz']).max()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "gsc_*.npz"))),
                         ids=lambda p: os.path.basename(p))
def test_against_reference_golden(path):
    g = np.load(path)
    D, H, Hp, gam = (int(v) for v in g['meta'])
    s2 = g['sigma_sq0']
    params = {'W': g['W0'].copy(), 'pi': g['pi0'].copy(), 'mu': g['mu0'].copy(), 'psi_sq': g['psi_sq0'].copy(),
              'sigma_sq': (float(s2) if s2.ndim == 0 else s2.copy())}
    check(D, H, Hp, gam, str(g['sigma_sq_type']), params, g['y'], float(g['T']), golden=g)


def synth(D, H, N, seed, stype):
    rng = np.random.RandomState(seed)
    Wgt = rng.standard_normal((D, H))
    s = rng.random_sample((N, H)) < 2.0 / H
    z = s * (1.0 + rng.standard_normal((N, H)))
    y = z @ Wgt.T + rng.standard_normal((N, D))
    psi = np.diag(0.5 + rng.random_sample(H))
    u = rng.standard_normal((H, 2)) * 0.15
    psi = psi + u @ u.T                                   # a non-diagonal slab covariance
    params = {'W': Wgt + 0.3 * rng.standard_normal((D, H)), 'pi': 0.05 + 0.2 * rng.random_sample(H),
              'mu': rng.standard_normal(H) * 0.5, 'psi_sq': psi,
              'sigma_sq': (1.3 if stype == 'scalar' else 0.8 + rng.random_sample(D))}
    return y, params


@pytest.mark.parametrize("D,H,Hp,gam,stype,N,seed,T", [
    (144, 64, 8, 3, 'scalar', 150, 3, 1.0),          # BASELINE configs[3] shape
    (144, 64, 8, 3, 'scalar', 100, 4, 1.7),
    (40, 16, 6, 4, 'diagonal', 200, 5, 1.0),
    (25, 10, 5, 1, 'scalar', 120, 6, 1.2),           # gamma = 1: null + singletons only
    (33, 12, 12, 2, 'diagonal', 90, 7, 1.0),         # H' == H
])
def test_against_oracle(D, H, Hp, gam, stype, N, seed, T):
    y, params = synth(D, H, N, seed, stype)
    check(D, H, Hp, gam, stype, params, y, T)


def test_full_covariance_against_oracle():
    """sigma_sq_type='full' with a symmetric positive-definite Sigma (what the M-step produces, gsc_et.py:677-691)."""
    D, H, Hp, gam, N = 20, 9, 5, 3, 150
    y, params = synth(D, H, N, 8, 'scalar')
    rng = np.random.RandomState(2)
    u = rng.standard_normal((D, 3)) * 0.4
    params['sigma_sq'] = np.diag(0.8 + rng.random_sample(D)) + u @ u.T
    check(D, H, Hp, gam, 'full', params, y, 1.0)
    check(D, H, Hp, gam, 'full', params, y, 1.6)


def test_singular_full_covariance_is_rejected_loudly():
    from prosper_b200._lib import PetError
    from prosper_b200.em.camodels.gsc_et import GSC
    y, params = synth(16, 8, 30, 1, 'scalar')
    params['sigma_sq'] = np.zeros((16, 16))
    m = GSC(16, 8, 4, 2, sigma_sq_type='full')
    with pytest.raises((PetError, AssertionError)):
        m.select_Hprimes(m.check_params(params) if False else params, {'y': y})
