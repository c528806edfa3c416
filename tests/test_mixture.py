"""Mixture models: the oracle against outputs of the reference (CPU), the device path against both (GPU)."""
import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLD, "mix_*.npz")))


class Anneal(dict):
    crit_params = []

    def __missing__(self, k):
        return 0.0

    def as_dict(self):
        return dict(self)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def params0(g):
    return dict((k[3:], g[k].copy()) for k in g.files if k.startswith('p0_'))


def kind(g):
    case = str(g['case'])
    return case, ('diagonal' if 'diag' in case else 'full'), (40. if case.endswith('_A') else np.nan)


@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p))
def test_oracle_mixture_matches_reference_golden(path):
    from oracle import mixture as om
    g = np.load(path, allow_pickle=False)
    case, stype, A = kind(g)
    D, H = g['p0_W'].shape
    o = om.MoG(D, H, stype) if case.startswith('mog') else om.MoP(D, H, A)
    suff = o.e_step(float(g['T']), params0(g), g['y'])
    assert rel(suff['logpj'], g['logpj']) < 1e-12 and rel(suff['posteriors_h'], g['post']) < 1e-10
    new = o.m_step(params0(g), g['post'], g['y'])
    for k in new:
        assert rel(new[k], g['new_' + k]) < 1e-10, k


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p))
def test_device_mixture_matches_reference_golden(path):
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from prosper_b200.em.mixturemodels.MoG import MoG
    from prosper_b200.em.mixturemodels.MoP import MoP
    g = np.load(path, allow_pickle=False)
    case, stype, A = kind(g)
    D, H = g['p0_W'].shape
    m = MoG(D, H, sigmas_sq_type=stype) if case.startswith('mog') else MoP(D, H, A=A)
    an = Anneal(T=float(g['T']))
    data = {'y': g['y'].copy()}
    suff = m.E_step(an, params0(g), data)
    assert suff['posteriors_h'].shape == g['post'].shape
    assert rel(suff['logpj'], g['logpj']) < 1e-11 and rel(suff['posteriors_h'], g['post']) < 1e-9
    new = m.M_step(an, params0(g), {'posteriors_h': g['post'].copy()}, data)
    for k in new:
        assert rel(new[k], g['new_' + k]) < 1e-9, k
    # fused step = E then M without the host round trip
    fused = m.step(an, params0(g), data)
    for k in fused:
        assert rel(fused[k], g['new_' + k]) < 1e-8, k
    # a few iterations stay finite and keep pies normalised
    p = fused
    for _ in range(3):
        p = m.step(an, p, data)
    assert np.isfinite(p['W']).all() and abs(p['pies'].sum() - 1) < 1e-12
