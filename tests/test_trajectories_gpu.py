"""50-iteration annealed EM trajectories of MCA, TSC, DSC and GSC on the GPU against the NumPy oracle, iteration by
iteration (north_star: "trajectories tracked over 50 iterations").  Flow of examples/barstests/bars-learning.py:77-88:
LinearAnnealing(50), model.step per iteration.  BSC's trajectory lives in test_bsc_gpu.py.

Shapes are the BASELINE.json configurations (cfg 2: MCA bars with T 4 -> 1 as param-bars-mca.py:34-37; cfg 3: TSC / DSC
on 8x8 bars, H=16, H'=8, gamma=4; cfg 4: GSC D=144, H=64, H'=8, gamma=3) at an N the oracle steps through in a few
minutes.

MCA's preselection score sum_d max(W_hd - y_d, 0) (mca_et.py:105-106) is EXACTLY zero for every cause that lies below
the datapoint, so rows with tied scores are common on bars data and the reference's pick among them is whatever
np.argsort's introsort leaves; the MCA trajectory therefore checks the device selection with the north-star gap rule
and then runs each iteration on the oracle's candidate order.  GSC's W update inverts sum_n <s z z^T s> (gsc_et.py:624),
which at N = 160 < the ~2000 datapoints that would excite all 64 units is conditioned ~1e10: its trajectory is
teacher-forced (both sides step from the same parameters every iteration) with the bound on W and sigma_sq widened
accordingly, and a second, well-conditioned configuration runs free."""
import numpy as np
import pytest

from helpers import bars_dict, rel_err, cand_mismatch_gap
from oracle.common import DictAnneal

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL = 1e-6


def cp(p):
    return dict((k, (np.copy(v) if isinstance(v, np.ndarray) else v)) for k, v in p.items())


def schedule(T, ncut):
    from prosper_b200.em.annealing import LinearAnnealing
    an = LinearAnnealing(50)
    an['T'] = T
    an['Ncut_factor'] = ncut
    an['anneal_prior'] = False
    return an


def worst_err(new, po, keys):
    w = 0.0
    for k in keys:
        w = max(w, rel_err(new[k], po[k]))
    return w


def run_trajectory(m, o, anneal, y, params, keys, tol=None, forced=False, inject_candidates=False):
    """Steps both sides through the 50-iteration schedule and returns the worst relative difference per key.
    forced: the oracle continues from the CUDA parameters after every iteration (per-step comparison from identical
    inputs); inject_candidates: the device selection is checked with the gap rule, then the iteration runs on the
    oracle's candidates."""
    assert torch.cuda.is_available(), "GPU tests need a B200"
    po, pm = cp(params), cp(params)
    worst = dict((k, 0.0) for k in keys)
    it = 0
    while not anneal.finished:
        an = DictAnneal(**anneal.as_dict())
        if inject_candidates:
            poc = o.check_params(cp(po))
            od = o.select_hprimes(poc, {'y': y.copy()})
            po = o.m_step(an, poc, o.e_step(an, poc, od), od)
            pmc = m.check_params(cp(pm))
            md = m.select_Hprimes(pmc, {'y': y})
            bad, gap = cand_mismatch_gap(od['_sim'], od['candidates'], md['candidates'])
            assert bad == 0 or gap < 1e-9 * max(1.0, np.abs(od['_sim']).max()), (it, bad, gap)
            m.engine.set_candidates(od['candidates'])
            pm = m._m_step(an, pmc, None, fused=False)
        else:
            po = o.step(an, cp(po), {'y': y.copy()})
            pm = m.step(anneal, pm, {'y': y})
        anneal.next()
        for k in keys:
            worst[k] = max(worst[k], rel_err(pm[k], po[k]))
        if forced:
            po = dict((k, (np.copy(np.asarray(pm[k])) if isinstance(pm[k], np.ndarray) else pm[k])) for k in po if k in pm)
        it += 1
    assert it == 50
    return worst, pm


def test_mca_trajectory_cfg2():
    """BASELINE configs[1]: MCA-ET bars 5x5, H=10, H'=6, gamma=3, N=2000, T 4 -> 1 (param-bars-mca.py:34-37)."""
    from prosper_b200.em.camodels.mca_et import MCA_ET
    from oracle.mca import MCA
    D, H, Hp, gam, N = 25, 10, 6, 3, 2000
    rng = np.random.RandomState(1)
    W = 10.0 * bars_dict(H)
    s = rng.random_sample((N, H)) < 0.2
    y = np.where(s[:, None, :], W[None, :, :], 0.0).max(axis=2) + 2.0 * rng.standard_normal((N, D))
    mean = y.mean(0)
    sig0 = np.sqrt(((y - mean) ** 2).mean(0)).sum() / D
    params = {'W': np.abs(mean[:, None] + rng.normal(scale=sig0 / 4., size=(D, H))), 'pi': 1. / H, 'sigma': sig0}
    m, o = MCA_ET(D, H, Hp, gam), MCA(D, H, Hp, gam)
    worst, last = run_trajectory(m, o, schedule([(0, 4.), (.8, 1.)], [(0, 0.), (2. / 3, 1.)]), y, params,
                                 ('W', 'pi', 'sigma'), inject_candidates=True)
    assert max(worst.values()) < TOL, worst
    err = np.abs(last['W'][:, :, None] - W[:, None, :]).mean(axis=0)
    assert (err.min(axis=0) < 1.5).all()                      # every bar is found


@pytest.mark.parametrize("name", ["tsc", "dsc"])
def test_tsc_dsc_trajectory_cfg3(name):
    """BASELINE configs[2]: 8x8 bars, H=16, H'=8, gamma=4 (1697 / 1680 states), N reduced to 400 for the oracle."""
    D, H, Hp, gam, N = 64, 16, 8, 4, 400
    rng = np.random.RandomState(2)
    W = 10.0 * bars_dict(H)
    if name == 'tsc':
        from prosper_b200.em.camodels.tsc_et import TSC_ET
        from oracle.tsc import TSC
        u = rng.random_sample((N, H))
        s = np.where(u < 0.0625, -1.0, np.where(u < 0.125, 1.0, 0.0))
        m, o = TSC_ET(D, H, Hp, gam), TSC(D, H, Hp, gam)
        pi0 = 1. / H
    else:
        from prosper_b200.em.camodels.dsc_et import DSC_ET
        from oracle.dsc import DSC
        st = np.array([-1., 0., 1.])
        s = rng.choice(st, size=(N, H), p=[.06, .88, .06])
        m, o = DSC_ET(D, H, Hp, gam, st), DSC(D, H, Hp, gam, st)
        pi0 = np.array([0.5 / H, 1. - 1. / H, 0.5 / H])
    y = s @ W.T + 2.0 * rng.standard_normal((N, D))
    mean = y.mean(0)
    sig0 = np.sqrt(((y - mean) ** 2).mean(0)).sum() / D
    params = {'W': mean[:, None] + rng.normal(scale=sig0 / 4., size=(D, H)), 'pi': pi0, 'sigma': sig0}
    worst, _ = run_trajectory(m, o, schedule([(0, 2.), (.7, 1.)], [(0, 0.), (2. / 3, 1.)]), y, params, ('W', 'pi', 'sigma'))
    assert max(worst.values()) < TOL, worst


def test_gsc_trajectory_cfg4():
    """BASELINE configs[3]: GSC D=144, H=64, H'=8, gamma=3 with scalar noise, N reduced to 160 for the oracle."""
    from prosper_b200.em.camodels.gsc_et import GSC
    from oracle.gsc import GSC as OGSC
    D, H, Hp, gam, N = 144, 64, 8, 3, 160
    rng = np.random.RandomState(3)
    W = rng.standard_normal((D, H))
    s = rng.random_sample((N, H)) < 2.0 / H
    y = (s * (1.0 + rng.standard_normal((N, H)))) @ W.T + rng.standard_normal((N, D))
    mean = y.mean(0)
    var = ((y - mean) ** 2).mean(0)
    sig0 = np.sqrt(var).sum() / D
    params = {'W': mean[:, None] + rng.normal(scale=sig0 / 4., size=(D, H)), 'pi': np.maximum(rng.rand(H) * 0.95, 0.05),
              'sigma_sq': float(var.mean() + 0.001), 'mu': rng.normal(0, 1, H), 'psi_sq': np.diag(np.maximum(rng.rand(H) * 2, 0.05))}
    m, o = GSC(D, H, Hp, gam, sigma_sq_type='scalar'), OGSC(D, H, Hp, gam, sigma_sq_type='scalar')
    worst, _ = run_trajectory(m, o, schedule([(0, 1.2), (.6, 1.)], [(0, 0.)]), y, params,
                              ('W', 'pi', 'mu', 'psi_sq', 'sigma_sq'), forced=True)
    assert max(worst[k] for k in ('pi', 'mu', 'psi_sq')) < TOL, worst
    assert worst['W'] < 1e-3 and worst['sigma_sq'] < 1e-4, worst          # inverse of a ~1e10-conditioned sum_szsz


def test_gsc_trajectory_well_conditioned():
    """GSC running free for 50 annealed iterations where every unit is excited often (D=36, H=12, H'=6, gamma=3, N=240)."""
    from prosper_b200.em.camodels.gsc_et import GSC
    from oracle.gsc import GSC as OGSC
    D, H, Hp, gam, N = 36, 12, 6, 3, 240
    rng = np.random.RandomState(4)
    W = 3.0 * rng.standard_normal((D, H))
    s = rng.random_sample((N, H)) < 0.25
    y = (s * (1.0 + 0.5 * rng.standard_normal((N, H)))) @ W.T + rng.standard_normal((N, D))
    mean = y.mean(0)
    var = ((y - mean) ** 2).mean(0)
    sig0 = np.sqrt(var).sum() / D
    params = {'W': mean[:, None] + rng.normal(scale=sig0 / 4., size=(D, H)), 'pi': np.full(H, 0.25),
              'sigma_sq': float(var.mean() + 0.001), 'mu': np.ones(H), 'psi_sq': np.eye(H)}
    m, o = GSC(D, H, Hp, gam, sigma_sq_type='scalar'), OGSC(D, H, Hp, gam, sigma_sq_type='scalar')
    worst, _ = run_trajectory(m, o, schedule([(0, 1.5), (.6, 1.)], [(0, 0.)]), y, params,
                              ('W', 'pi', 'mu', 'psi_sq', 'sigma_sq'))
    assert max(worst.values()) < TOL, worst
