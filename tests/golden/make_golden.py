"""Mint golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/prosper through oracle/ref_harness.py (fake mpi4py/tables, NumPy-2
aliases), runs select_Hprimes / E_step / M_step of each model on small seeded
inputs and stores inputs + every intermediate in tests/golden/<model>_<case>.npz.  The GPU box
has no /root/reference, so tests only ever read the .npz files.
Protocol per case: np.random.seed(seed); data = model.generate_data(gt, N);
params = model.standard_init(data); one step at the given (T, Ncut_factor, anneal_prior).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import ref_harness  # noqa: E402

ref_harness.load()
from prosper.em.annealing import LinearAnnealing  # noqa: E402
from prosper.utils.barstest import generate_bars_dict  # noqa: E402

log = ref_harness.KeepLog().install()


def run_case(name, case, model, gt, N, seed, T, ncut, ap, meta, mutate=None):
    np.random.seed(seed)
    data = model.generate_data(gt, N)
    params = model.standard_init(data)
    if mutate:
        params = mutate(params)
    an = LinearAnnealing(2)
    an['T'] = T
    an['Ncut_factor'] = ncut
    an['anneal_prior'] = ap
    params = model.check_params(params)
    p0 = dict((k, np.copy(v)) for k, v in params.items())
    data = model.select_Hprimes(params, data)
    suff = model.E_step(an, params, data)
    new = model.M_step(an, params, suff, data)
    out = dict(model=name, meta=np.array(meta), T=T, Ncut_factor=ncut, anneal_prior=ap,
               y=data['y'], W0=p0['W'], pi0=p0['pi'], sigma0=p0['sigma'],
               candidates=np.asarray(data['candidates'], dtype=np.int64), logpj=suff['logpj'],
               W_new=new['W'], pi_new=new['pi'], sigma_new=new['sigma'],
               L=log.values.get('L', [np.nan])[-1] if 'Q' not in new or name in ('tsc', 'dsc') else np.nan,
               Q=new.get('Q', np.nan), N_use=log.last('N_use'))
    if name == 'dsc':
        out['states'] = model.states
    path = os.path.join(HERE, "%s_%s.npz" % (name, case))
    np.savez_compressed(path, **out)
    print("wrote", path, "pi_new", new['pi'], "sigma_new", new['sigma'], "L/Q", out['L'], out['Q'], "N_use", out['N_use'])


def main():
    from prosper.em.camodels.bsc_et import BSC_ET
    from prosper.em.camodels.tsc_et import TSC_ET
    from prosper.em.camodels.dsc_et import DSC_ET
    from prosper.em.camodels.mca_et import MCA_ET
    from prosper.em.camodels.mmca_et import MMCA_ET
    gt10 = {'W': 10 * generate_bars_dict(10), 'pi': 0.2, 'sigma': 2.0}
    cases = [('t1', 1.0, 0.0, False), ('t2cut', 2.0, 0.5, False), ('prior', 1.5, 0.3, True)]
    for case, T, c, ap in cases:
        run_case('bsc', case, BSC_ET(25, 10, 6, 3), gt10, 300, 1, T, c, ap, (25, 10, 6, 3))
        run_case('mca', case, MCA_ET(25, 10, 6, 3), gt10, 300, 1, T, c, ap, (25, 10, 6, 3))
        run_case('mmca', case, MMCA_ET(25, 10, 6, 3), gt10, 300, 1, T, c, ap, (25, 10, 6, 3))
    gt12 = {'W': 10 * generate_bars_dict(12), 'pi': 0.125, 'sigma': 2.0}
    gt12d = {'W': 10 * generate_bars_dict(12), 'pi': np.array([.06, .88, .06]), 'sigma': 2.0}
    for case, T, c, ap in cases:
        run_case('tsc', case, TSC_ET(36, 12, 6, 3), gt12, 200, 1, T, c, ap, (36, 12, 6, 3))
        run_case('dsc', case, DSC_ET(36, 12, 6, 3, np.array([-1., 0., 1.])), gt12d, 200, 1, T, c, ap, (36, 12, 6, 3))
    # a second binary shape with gamma=4 and H'=8 (more states than columns of singles)
    gt16 = {'W': 10 * generate_bars_dict(16), 'pi': 0.125, 'sigma': 2.0}
    run_case('bsc', 'h16', BSC_ET(64, 16, 8, 4), gt16, 150, 2, 1.2, 0.7, False, (64, 16, 8, 4))


def run_gsc(case, D, H, Hp, gam, stype, N, seed, T):
    """GSC returns moment tensors instead of logpj and mutates its inputs (gsc_et.py:572-573,584-718)."""
    from prosper.em.camodels.gsc_et import GSC
    np.random.seed(seed)
    model = GSC(D, H, Hp, gam, sigma_sq_type=stype)
    sig = {'scalar': 1.0, 'diagonal': np.ones(D), 'full': np.eye(D)}[stype]
    gt = {'W': 10 * generate_bars_dict(H), 'pi': 0.2 * np.ones(H), 'mu': np.ones(H), 'psi_sq': np.eye(H), 'sigma_sq': sig}
    data = model.generate_data(gt, N)
    params = model.standard_init(data)
    an = LinearAnnealing(2)
    an['T'] = T
    if stype == 'full':      # standard_init's 'full' matrix is non-symmetric (broadcasting slip, gsc_et.py:85):
        v = np.diag(params['sigma_sq']).copy()       # use a proper symmetric positive-definite one instead
        u = np.random.randn(D) * 0.5
        params['sigma_sq'] = np.diag(v) + np.outer(u, u)
    params = model.check_params(params)
    p0 = dict((k, np.copy(v)) for k, v in params.items())
    y0 = data['y'].copy()
    data = model.select_Hprimes(params, data)
    cand = np.zeros((N, model.Hprime), dtype=np.int64)
    for cl in data['data_clusters'].values():
        cand[cl['ind']] = cl['hprimes'][None, :]
    suff = model.E_step(an, params, data)
    y_after = data['y'].copy()
    new = model.M_step(an, params, suff, data)
    out = dict(model='gsc', meta=np.array((D, H, Hp, gam)), sigma_sq_type=stype, T=T, y=y0, y_after=y_after,
               candidates=cand, candidates_after=np.asarray(data['candidates']),
               W0=p0['W'], pi0=p0['pi'], mu0=p0['mu'], psi_sq0=p0['psi_sq'], sigma_sq0=p0['sigma_sq'],
               xpt_s=suff['xpt_s'], xpt_ss=suff['xpt_ss'], xpt_sz=suff['xpt_sz'], xpt_szsz=suff['xpt_szsz'],
               W_new=new['W'], pi_new=new['pi'], mu_new=new['mu'], psi_sq_new=new['psi_sq'], sigma_sq_new=new['sigma_sq'])
    path = os.path.join(HERE, "gsc_%s.npz" % case)
    np.savez_compressed(path, **out)
    print("wrote", path, "pi_new[:3]", new['pi'][:3], "sigma_sq_new", np.ravel(new['sigma_sq'])[:2], "W sum", new['W'].sum())


def main_gsc():
    run_gsc('scalar_t1', 25, 10, 6, 3, 'scalar', 120, 1, 1.0)
    run_gsc('scalar_t2', 25, 10, 6, 3, 'scalar', 120, 1, 2.0)
    run_gsc('diag_t1', 25, 10, 5, 2, 'diagonal', 100, 2, 1.3)
    run_gsc('full_t1', 16, 8, 4, 2, 'full', 80, 3, 1.0)


def run_inference(name, case, model, gt, N, seed, T, meta, **kw):
    """CAModel.inference of the unmodified reference (camodels/__init__.py:255-375 and the TSC/DSC overrides)."""
    np.random.seed(seed)
    data = model.generate_data(gt, N)
    an = LinearAnnealing(2)
    an['T'] = T
    an['anneal_prior'] = False
    params = dict((k, np.copy(v)) for k, v in gt.items())
    if name == 'bsc':
        params['mu'] = np.zeros(meta[0])
    res = model.inference(an, dict(params), {'y': data['y'].copy()}, **kw)
    out = dict(model=name, meta=np.array(meta), T=T, y=data['y'], W=params['W'], pi=params['pi'], sigma=params['sigma'])
    for k, v in kw.items():
        out['kw_' + k] = -1 if v is None else v
    for k, v in res.items():
        out['res_' + k] = v
    if name == 'dsc':
        out['states'] = model.states
    path = os.path.join(HERE, "infer_%s_%s.npz" % (name, case))
    np.savez_compressed(path, **out)
    print("wrote", path, "gamma used", np.unique(res['gamma']), "p[0]", res['p'][0][:3])


def run_inference_gsc(case, D, H, Hp, gam, stype, N, seed, **kw):
    from prosper.em.camodels.gsc_et import GSC
    np.random.seed(seed)
    model = GSC(D, H, Hp, gam, sigma_sq_type=stype)
    sig = {'scalar': 1.5, 'diagonal': 1.0 + 0.5 * np.arange(D) / D}[stype]
    gt = {'W': 10 * generate_bars_dict(H), 'pi': 0.2 * np.ones(H), 'mu': np.ones(H) + 0.1 * np.arange(H),
          'psi_sq': np.eye(H) + 0.05, 'sigma_sq': sig}
    data = model.generate_data(gt, N)
    an = LinearAnnealing(2)
    an['T'] = 1.0
    params = dict((k, np.copy(v)) for k, v in gt.items())
    res = model.inference(an, dict((k, np.copy(v)) for k, v in params.items()), {'y': data['y'].copy()}, **kw)
    out = dict(model='gsc', meta=np.array((D, H, Hp, gam)), sigma_sq_type=stype, T=1.0, y=data['y'], W=params['W'],
               pi=params['pi'], mu=params['mu'], psi_sq=params['psi_sq'], sigma_sq=params['sigma_sq'])
    for k, v in kw.items():
        out['kw_' + k] = -1 if v is None else v
    for k, v in res.items():
        out['res_' + k] = v
    path = os.path.join(HERE, "infer_gsc_%s.npz" % case)
    np.savez_compressed(path, **out)
    print("wrote", path, "p[0]", res['p'][0][:3])


def main_inference():
    run_inference_gsc('scalar', 25, 10, 6, 3, 'scalar', 40, 4, topK=5, logprob=False, adaptive=False)
    run_inference_gsc('diag_logp', 25, 10, 5, 2, 'diagonal', 40, 5, topK=6, logprob=True, adaptive=False)
    from prosper.em.camodels.bsc_et import BSC_ET
    from prosper.em.camodels.tsc_et import TSC_ET
    from prosper.em.camodels.dsc_et import DSC_ET
    from prosper.em.camodels.mca_et import MCA_ET
    from prosper.em.camodels.mmca_et import MMCA_ET
    gt10 = {'W': 10 * generate_bars_dict(10), 'pi': 0.25, 'sigma': 2.0}
    for name, cls in (('bsc', BSC_ET), ('mca', MCA_ET), ('mmca', MMCA_ET)):
        run_inference(name, 'fixed', cls(25, 10, 6, 3), gt10, 60, 4, 1.0, (25, 10, 6, 3), topK=5, logprob=False, adaptive=False)
        run_inference(name, 'logp', cls(25, 10, 6, 3), gt10, 60, 5, 1.0, (25, 10, 6, 3), topK=7, logprob=True, adaptive=False)
    run_inference('bsc', 'adaptive', BSC_ET(25, 10, 5, 2), gt10, 80, 6, 1.0, (25, 10, 5, 2), topK=4, logprob=False,
                  adaptive=True, Hprime_max=8, gamma_max=5)
    gt12 = {'W': 10 * generate_bars_dict(12), 'pi': 0.2, 'sigma': 2.0}
    run_inference('tsc', 'fixed', TSC_ET(36, 12, 6, 3), gt12, 50, 4, 1.0, (36, 12, 6, 3), topK=5, logprob=False, adaptive=False)
    run_inference('tsc', 'logp', TSC_ET(36, 12, 6, 3), gt12, 50, 5, 1.0, (36, 12, 6, 3), topK=5, logprob=True, adaptive=False)
    run_inference('tsc', 'adaptive', TSC_ET(36, 12, 5, 2), gt12, 60, 6, 1.0, (36, 12, 5, 2), topK=3, logprob=False,
                  adaptive=True, Hprime_max=7, gamma_max=4)
    gt12d = {'W': 10 * generate_bars_dict(12), 'pi': np.array([.08, .84, .08]), 'sigma': 2.0}
    run_inference('dsc', 'fixed', DSC_ET(36, 12, 6, 3, np.array([-1., 0., 1.])), gt12d, 50, 4, 1.0, (36, 12, 6, 3),
                  topK=5, logprob=False, adaptive=False)
    run_inference('dsc', 'logp', DSC_ET(36, 12, 6, 3, np.array([-1., 0., 1.])), gt12d, 50, 5, 1.0, (36, 12, 6, 3),
                  topK=5, logprob=True, adaptive=False)


def run_mixture(case, model, gt, N, seed, T):
    """E_step + M_step of the unmodified mixture models (mixturemodels/MoG.py, MoP.py)."""
    np.random.seed(seed)
    data = model.generate_data(gt, N)
    params = model.standard_init({'y': data['y']})
    if 'pies' not in params:
        params['pies'] = np.ones(model.H) / model.H
    if 'sigmas_sq' in gt and 'sigmas_sq' not in params:
        params['sigmas_sq'] = gt['sigmas_sq'].copy()
    if case.startswith('mop'):
        params['W'] = np.abs(params['W']) + 0.5                 # Poisson rates must be positive
    an = LinearAnnealing(2)
    an['T'] = T
    p0 = dict((k, np.copy(v)) for k, v in params.items())
    suff = model.E_step(an, params, {'y': data['y']})
    new = model.M_step(an, dict((k, np.copy(v)) for k, v in params.items()), suff, {'y': data['y']})
    out = dict(case=case, T=T, y=data['y'], post=suff['posteriors_h'], logpj=suff['logpj'])
    for k, v in p0.items():
        out['p0_' + k] = v
    for k, v in new.items():
        out['new_' + k] = v
    path = os.path.join(HERE, "mix_%s.npz" % case)
    np.savez_compressed(path, **out)
    print("wrote", path, "pies_new", np.round(new['pies'], 4))


def main_mixture():
    from prosper.em.mixturemodels.MoG import MoG
    from prosper.em.mixturemodels.MoP import MoP
    rng = np.random.RandomState(0)
    D, H = 6, 4
    Wg = rng.randn(D, H) * 3
    pies = np.array([.1, .2, .3, .4])
    sd = 0.5 + rng.rand(H, D)
    run_mixture('mog_diag', MoG(D, H, sigmas_sq_type='diagonal'), {'W': Wg, 'pies': pies, 'sigmas_sq': sd}, 300, 1, 1.0)
    run_mixture('mog_diag_t2', MoG(D, H, sigmas_sq_type='diagonal'), {'W': Wg, 'pies': pies, 'sigmas_sq': sd}, 200, 2, 2.0)
    sf = np.stack([np.diag(sd[h]) + 0.1 for h in range(H)])
    run_mixture('mog_full', MoG(D, H, sigmas_sq_type='full'), {'W': Wg, 'pies': pies, 'sigmas_sq': sf}, 300, 3, 1.0)
    Wp = np.abs(rng.randn(D, H)) * 10 + 1
    run_mixture('mop', MoP(D, H), {'W': Wp, 'pies': pies}, 300, 4, 1.0)
    run_mixture('mop_A', MoP(D, H, A=40.), {'W': Wp, 'pies': pies}, 300, 5, 1.3)


def main_northstar():
    """The unmodified reference at the north-star shape (bsc_et.py:98-438, D=676 H=1000 H'=12 gamma=5) on 48 seeded
    datapoints (59 ms per datapoint and step here).  Inputs come from tests/helpers.northstar_inputs and are
    regenerated by the tests; stored: y (as a cross-check of the regeneration), candidates, logpj and the results
    (every fourth row of W_new, which keeps the file at 4 MB) of one step at T=1 without truncation and one at T=1.3 with Ncut_factor=1."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from helpers import northstar_inputs
    from prosper.em.camodels.bsc_et import BSC_ET
    y, p0 = northstar_inputs()
    out = dict(meta=np.array((676, 1000, 12, 5)), N=y.shape[0], seed=5, y=y, W0_checksum=np.array([p0['W'].sum(), np.abs(p0['W']).sum()]),
               pi0=p0['pi'], sigma0=p0['sigma'])
    model = BSC_ET(676, 1000, 12, 5)
    for tag, T, ncut in (('a', 1.0, 0.0), ('b', 1.3, 1.0)):
        an = LinearAnnealing(2)
        an['T'] = T
        an['Ncut_factor'] = ncut
        an['anneal_prior'] = False
        params = {'W': p0['W'].copy(), 'pi': p0['pi'], 'sigma': p0['sigma']}
        data = model.select_Hprimes(params, {'y': y.copy()})
        suff = model.E_step(an, params, data)
        new = model.M_step(an, params, suff, data)
        out.update({'T_' + tag: T, 'Ncut_' + tag: ncut, 'W_new_rows4_' + tag: new['W'][::4].copy(), 'pi_new_' + tag: new['pi'],
                    'sigma_new_' + tag: new['sigma'], 'L_' + tag: log.values['L'][-1], 'N_use_' + tag: log.last('N_use')})
        if tag == 'a':
            out['candidates'] = np.asarray(data['candidates'], dtype=np.int64)
            out['logpj'] = suff['logpj']
        print("northstar", tag, "pi_new", new['pi'], "sigma_new", new['sigma'], "L", out['L_' + tag], "N_use", out['N_use_' + tag])
    path = os.path.join(HERE, "northstar_bsc.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == 'northstar':
        main_northstar()
    elif len(sys.argv) > 1 and sys.argv[1] == 'gsc':
        main_gsc()
    elif len(sys.argv) > 1 and sys.argv[1] == 'inference':
        main_inference()
    elif len(sys.argv) > 1 and sys.argv[1] == 'mixture':
        main_mixture()
    else:
        main_mixture()
        main_inference()
        main()
        main_gsc()
        main_northstar()
