"""50-iteration annealed EM trajectories of MCA, TSC, DSC and GSC on the GPU against the NumPy oracle, iteration by
iteration (north_star: "trajectories tracked over 50 iterations").  Flow of examples/barstests/bars-learning.py:77-88:
LinearAnnealing(50), model.step per iteration.  BSC's trajectory lives in test_bsc_gpu.py.

Shapes are the BASELINE.json configurations (cfg 2: MCA bars with T 4 -> 1 as param-bars-mca.py:34-37; cfg 3: TSC / DSC
on 8x8 bars, H=16, H'=8, gamma=4; cfg 4: GSC D=144, H=64, H'=8, gamma=3) at an N the oracle steps through in a few
minutes.  A datapoint whose candidate scores tie within rounding may pick a different candidate on the two sides;
such a step is re-synchronised (the oracle continues from the CUDA parameters) and counted, it must stay rare."""
import numpy as np
import pytest

from helpers import bars_dict, rel_err
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


def run_trajectory(m, o, anneal, y, params, keys, check=None):
    """Steps both sides through the schedule; returns (worst relative difference over all iterations, number of
    iterations that had to be re-synchronised)."""
    assert torch.cuda.is_available(), "GPU tests need a B200"
    po, pm = cp(params), cp(params)
    worst, resync = 0.0, 0
    it = 0
    while not anneal.finished:
        an = DictAnneal(**anneal.as_dict())
        if check is not None:
            po = check(po)
        po = o.step(an, cp(po), {'y': y.copy()})
        pm = m.step(anneal, pm, {'y': y})
        anneal.next()
        err = worst_err(pm, po, keys)
        if err >= TOL and err < 1e-2:        # a flipped near-tie: continue from the same point, count it
            resync += 1
            po = dict((k, (np.copy(np.asarray(pm[k])) if isinstance(pm[k], np.ndarray) else pm[k])) for k in po if k in pm)
            err = 0.0
        worst = max(worst, err)
        it += 1
    assert it == 50
    return worst, resync, pm


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
    worst, resync, last = run_trajectory(m, o, schedule([(0, 4.), (.8, 1.)], [(0, 0.), (2. / 3, 1.)]), y, params,
                                         ('W', 'pi', 'sigma'), check=o.check_params)
    assert worst < TOL and resync <= 2, (worst, resync)
    err = np.abs(last['W'][:, :, None] - W[:, None, :]).mean(axis=0)
    assert (err.min(axis=0) < 1.5).all()                      # every bar is found


@pytest.mark.parametrize("name", ["tsc", "dsc"])
def test_tsc_dsc_trajectory_cfg3(name):
    """BASELINE configs[2]: 8x8 bars, H=16, H'=8, gamma=4 (1697 / 1680 states), N reduced to 600 for the oracle."""
    D, H, Hp, gam, N = 64, 16, 8, 4, 600
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
    worst, resync, _ = run_trajectory(m, o, schedule([(0, 2.), (.7, 1.)], [(0, 0.), (2. / 3, 1.)]), y, params,
                                      ('W', 'pi', 'sigma'))
    assert worst < TOL and resync <= 2, (worst, resync)


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
    worst, resync, _ = run_trajectory(m, o, schedule([(0, 1.2), (.6, 1.)], [(0, 0.)]), y, params,
                                      ('W', 'pi', 'mu', 'psi_sq', 'sigma_sq'))
    assert worst < TOL and resync <= 2, (worst, resync)
