"""CPU tests (-m "not gpu"): the oracle against the golden vectors minted from the reference,
the host-side mirrors, and the C-ABI library's exports."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from helpers import rel_err
from oracle import ref_harness
from oracle.common import DictAnneal

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _oracle_for(name, meta):
    if name == 'bsc':
        from oracle.bsc import BSC
        return BSC(*meta)
    if name == 'tsc':
        from oracle.tsc import TSC
        return TSC(*meta)
    if name == 'dsc':
        from oracle.dsc import DSC
        return DSC(*meta)
    if name == 'mca':
        from oracle.mca import MCA
        return MCA(*meta)
    if name == 'mmca':
        from oracle.mca import MMCA
        return MMCA(*meta)
    raise KeyError(name)


def golden_files():
    return sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(p).startswith(("gsc_", "infer_", "mix_", "northstar_")))


@pytest.mark.parametrize("path", golden_files(), ids=[os.path.basename(p) for p in golden_files()])
def test_oracle_matches_reference_golden(path):
    """Every stage of the oracle reproduces what the unmodified reference produced."""
    g = np.load(path, allow_pickle=False)
    name = str(g['model'])
    o = _oracle_for(name, tuple(int(v) for v in g['meta']))
    an = DictAnneal(T=float(g['T']), Ncut_factor=float(g['Ncut_factor']), anneal_prior=bool(g['anneal_prior']))
    params = {'W': g['W0'].copy(), 'pi': (g['pi0'].copy() if g['pi0'].ndim else float(g['pi0'])), 'sigma': float(g['sigma0'])}
    data = {'y': g['y'].copy()}
    data = o.select_hprimes(params, data)
    assert np.array_equal(data['candidates'], g['candidates'])
    suff = o.e_step(an, params, data)
    assert np.abs(suff['logpj'] - g['logpj']).max() < 1e-9 * max(1.0, np.abs(g['logpj']).max())
    new = o.m_step(an, params, suff, data)
    assert rel_err(new['W'], g['W_new']) < 1e-9
    assert rel_err(new['pi'], g['pi_new']) < 1e-10
    assert rel_err(new['sigma'], g['sigma_new']) < 1e-10
    if name in ('mca', 'mmca'):
        assert abs(new['Q'] - float(g['Q'])) < 1e-9 * abs(float(g['Q']))
    else:
        assert abs(o.log['L'] - float(g['L'])) < 1e-9 * abs(float(g['L']))
    assert o.log['N_use'] == int(g['N_use'])


def test_oracle_matches_the_reference_at_the_north_star_shape():
    """BASELINE configs[4] shape (D=676, H=1000, H'=12, gamma=5): the oracle against what the unmodified reference
    returned for the 48 seeded datapoints of helpers.northstar_inputs (bsc_et.py:98-438), with and without truncation.
    N < H makes Wq rank deficient: the update is lstsq's minimum-norm solution, compared row by row."""
    from helpers import northstar_inputs
    from oracle.bsc import BSC
    g = np.load(os.path.join(GOLDEN, "northstar_bsc.npz"))
    y, p0 = northstar_inputs(int(g['N']), int(g['seed']))
    assert np.array_equal(y, g['y'])
    assert np.allclose([p0['W'].sum(), np.abs(p0['W']).sum()], g['W0_checksum'], rtol=1e-14)
    o = BSC(676, 1000, 12, 5)
    for tag in ('a', 'b'):
        an = DictAnneal(T=float(g['T_' + tag]), Ncut_factor=float(g['Ncut_' + tag]), anneal_prior=False)
        params = {'W': p0['W'].copy(), 'pi': p0['pi'], 'sigma': p0['sigma']}
        data = o.select_hprimes(params, {'y': y.copy()})
        suff = o.e_step(an, params, data)
        if tag == 'a':
            assert np.array_equal(data['candidates'], g['candidates'])
            assert np.abs(suff['logpj'] - g['logpj']).max() < 1e-10 * np.abs(g['logpj']).max()
        new = o.m_step(an, params, suff, data)
        assert rel_err(new['W'][::4], g['W_new_rows4_' + tag]) < 1e-8
        assert abs(new['pi'] - float(g['pi_new_' + tag])) < 1e-10 * float(g['pi_new_' + tag])
        assert abs(new['sigma'] - float(g['sigma_new_' + tag])) < 1e-10 * float(g['sigma_new_' + tag])
        assert abs(o.log['L'] - float(g['L_' + tag])) < 1e-10 * abs(float(g['L_' + tag]))
        assert o.log['N_use'] == int(g['N_use_' + tag])


def test_state_spaces_match_reference_orders():
    from oracle import states
    sm, sa = states.binary_states(6, 3)
    assert sm.shape == (35, 6) and sa.min() == 2 and sa.max() == 3
    assert sm[0].tolist() == [1, 1, 0, 0, 0, 0] and sm[-1].tolist() == [0, 0, 0, 1, 1, 1]
    sm12, _ = states.binary_states(12, 5)
    assert sm12.shape[0] == 1573
    ssm, tsm, n_all, sabs = states.ternary_states(8, 4, 16)
    assert tsm.shape == (1697, 8) and n_all == 3 ** 8 and ssm.shape == (32, 16)
    assert (ssm[:16] == -np.eye(16)).all() and (ssm[16:] == np.eye(16)).all()
    ssm, dsm, dabs, k0 = states.discrete_states(np.array([-1., 0., 1.]), 8, 4, 16)
    assert dsm.shape == (1680, 8) and k0 == 1 and dabs.shape == (3, 1680)
    assert np.all(dabs.sum(axis=0) == 16)


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box)")
def test_live_reference_state_matrices():
    ref_harness.load()
    from prosper.em.camodels import generate_state_matrix
    from prosper.em.camodels.tsc_et import generate_state_matrix as tsc_gen
    from oracle import states
    for hp, g in [(6, 3), (8, 5), (12, 5)]:
        _, _, sm, sa = generate_state_matrix(hp, g)
        osm, osa = states.binary_states(hp, g)
        assert np.array_equal(sm, osm) and np.array_equal(sa, osa)
    ssm, sm, n, sabs = tsc_gen(6, 3, 10, np.array([-1., 0., 1.]))
    o = states.ternary_states(6, 3, 10)
    assert np.array_equal(ssm, o[0]) and np.array_equal(sm, o[1]) and n == o[2] and np.array_equal(sabs, o[3])


def test_annealing_schedule():
    from prosper_b200.em.annealing import LinearAnnealing
    a = LinearAnnealing(50)
    a['T'] = [(0, 2.), (.7, 1.)]
    a['Ncut_factor'] = [(0, 0.), (2. / 3, 1.)]
    a['anneal_prior'] = False
    assert a['T'] == 2.0 and a['Ncut_factor'] == 0.0 and a['missing'] == 0.0
    seen = []
    while not a.finished:
        seen.append((a['T'], a['Ncut_factor'], a['step'], a['position']))
        a.next()
    seen = np.array(seen)
    assert len(seen) == 50 and np.allclose(seen.sum(0), [68., 33., 1225., 24.5])     # values minted from the reference
    assert a['T'] == 1.0 and a.as_dict()['max_step'] == 50
    with pytest.raises(RuntimeError):
        a.next()
    with pytest.raises(TypeError):
        a['bad'] = [1, 2]


def test_stride_data_rule():
    from prosper_b200.utils import parallel

    class C(object):
        def __init__(self, r, s):
            self.rank, self.size = r, s
    got = [parallel.stride_data(10, comm=C(r, 4)) for r in range(4)]
    assert got == [(0, 3), (3, 6), (6, 8), (8, 10)]                 # parallel.py:67-84
    got = [parallel.stride_data(10, balanced=True, comm=C(r, 4)) for r in range(4)]
    assert got == [(0, 2), (2, 4), (4, 6), (6, 8)]


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports exactly what include/prosper_b200.h declares."""
    from prosper_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "prosper_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pet_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.pet_abi_version() == 1


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly (no oracle / CPU route)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from prosper_b200 import _lib
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    lib = _lib.load()
    cfg = _lib.Config(0, 0, 25, 10, 6, 3, 0, None, 0)
    h = ctypes.c_void_p()
    assert lib.pet_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert b"no CPU path" in lib.pet_last_error() or b"CUDA" in lib.pet_last_error()
    m = BSC_ET(25, 10, 6, 3)
    with pytest.raises(RuntimeError):
        m.select_Hprimes({'W': np.zeros((25, 10)), 'pi': .1, 'sigma': 1.}, {'y': np.zeros((4, 25))})


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(GOLDEN), "..", "prosper_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "gsc_*.npz"))), ids=lambda p: os.path.basename(p))
def test_gsc_oracle_matches_reference_golden(path):
    """GSC returns posterior moment tensors and mutates its inputs; every piece is compared."""
    from oracle.gsc import GSC
    g = np.load(path)
    D, H, Hp, gam = (int(v) for v in g['meta'])
    o = GSC(D, H, Hp, gam, sigma_sq_type=str(g['sigma_sq_type']))
    an = DictAnneal(T=float(g['T']))
    s2 = g['sigma_sq0']
    params = {'W': g['W0'].copy(), 'pi': g['pi0'].copy(), 'mu': g['mu0'].copy(), 'psi_sq': g['psi_sq0'].copy(),
              'sigma_sq': (float(s2) if s2.ndim == 0 else s2.copy())}
    data = o.select_hprimes(params, {'y': g['y'].copy()})
    assert np.array_equal(data['candidates'], g['candidates'])
    suff = o.e_step(an, params, data)
    assert np.array_equal(data['y'], g['y_after']) and np.array_equal(data['candidates'], g['candidates_after'])
    for k in ('xpt_s', 'xpt_ss', 'xpt_sz', 'xpt_szsz'):
        assert np.abs(suff[k] - g[k]).max() < 1e-9 * max(1.0, np.abs(g[k]).max()), k
    new = o.m_step(an, params, suff, data)
    for k in ('W', 'pi', 'mu', 'psi_sq', 'sigma_sq'):
        assert rel_err(new[k], g[k + '_new']) < 1e-8, k


# ---- inference: the oracle's restatement against outputs of the reference ------------------
def _infer_close(a, b, tol=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    with np.errstate(invalid='ignore'):
        ok = (np.isnan(a) & np.isnan(b)) | (a == b) | (np.abs(a - b) <= tol * np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))
    return bool(ok.all())


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "infer_*.npz"))), ids=lambda p: os.path.basename(p))
def test_oracle_inference_matches_reference_golden(path):
    from oracle import inference as oinf
    from oracle.bsc import BSC
    from oracle.tsc import TSC
    from oracle.dsc import DSC
    from oracle.mca import MCA, MMCA
    from oracle.common import DictAnneal
    g = np.load(path, allow_pickle=False)
    name = str(g['model'])
    D, H, Hp, gam = [int(x) for x in g['meta']]
    cls = {'bsc': BSC, 'tsc': TSC, 'mca': MCA, 'mmca': MMCA}.get(name)
    make = (lambda hp, ga: DSC(D, H, hp, ga, g['states'])) if name == 'dsc' else (lambda hp, ga: cls(D, H, hp, ga))
    if name == 'gsc':
        from oracle.gsc import GSC
        make = lambda hp, ga: GSC(D, H, hp, ga, sigma_sq_type=str(g['sigma_sq_type']))
    kw = {}
    for k in g.files:
        if k.startswith('kw_'):
            v = g[k].item()
            kw[k[3:]] = None if (k[3:] in ('Hprime_max', 'gamma_max') and v == -1) else (bool(v) if k[3:] in ('logprob', 'adaptive') else int(v))
    an = DictAnneal(T=float(g['T']), anneal_prior=False)
    if name == 'gsc':
        params = dict((k, g[k].copy()) for k in ('W', 'pi', 'mu', 'psi_sq', 'sigma_sq'))
        if params['sigma_sq'].ndim == 0:
            params['sigma_sq'] = float(params['sigma_sq'])
    else:
        params = {'W': g['W'].copy(), 'pi': g['pi'] if g['pi'].ndim else float(g['pi']), 'sigma': float(g['sigma'])}
    if name == 'bsc':
        params['mu'] = np.zeros(D)
    res = oinf.inference(make, Hp, gam, an, params, g['y'].copy(), **kw)
    assert np.array_equal(res['gamma'], g['res_gamma']) and np.array_equal(res['Hprime'], g['res_Hprime'])
    assert _infer_close(res['p'], g['res_p'])
    assert _infer_close(res['m'], g['res_m'])
    if 'am' in res:
        assert _infer_close(res['am'], g['res_am'])
    assert (res['s'] == g['res_s']).all(axis=2).mean() > 0.97       # equal-probability ranks may swap


# ---- host-side mirror of the reference interface ---------------------------------------------------------
API_PAIRS = [('em.camodels', 'CAModel'), ('em.camodels.bsc_et', 'BSC_ET'), ('em.camodels.mca_et', 'MCA_ET'),
             ('em.camodels.mmca_et', 'MMCA_ET'), ('em.camodels.tsc_et', 'TSC_ET'), ('em.camodels.dsc_et', 'DSC_ET'),
             ('em', 'EM'), ('em', 'Model'), ('em.annealing', 'LinearAnnealing'), ('em.annealing', 'Annealing'),
             ('utils.datalog', 'DataLog'), ('utils.autotable', 'AutoTable'), ('em.camodels.gsc_et', 'GSC'),
             ('em.mixturemodels.MoG', 'MoG'), ('em.mixturemodels.MoP', 'MoP')]
# not mirrored, on purpose: `resume_init` uses undefined names upstream; the other three are the bodies
# of the reference's per-datapoint NumPy loops, which are CUDA kernels here (gsc_kernel.cu, mixture.cu) and have no host form
NOT_MIRRORED = {('GSC', 'resume_init'), ('MoG', 'resume_init'), ('MoP', 'resume_init'),
                ('GSC', 'component_scores'), ('GSC', 'compute_posterior_hprime'), ('MoG', 'log_p_y'), ('MoP', 'log_p_y')}


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box)")
def test_public_methods_and_positional_arguments_match_the_reference():
    """Every public method of the reference classes on the path exists here with the same positional parameters (a
    trailing **kwargs or extra defaulted parameter is allowed), so positional calls written for prosper keep working."""
    import importlib
    import inspect
    ref_harness.load()
    problems = []
    for mod, cls in API_PAIRS:
        r = getattr(importlib.import_module('prosper.' + mod), cls)
        o = getattr(importlib.import_module('prosper_b200.' + mod), cls)
        for name, fn in inspect.getmembers(r, callable):
            if (name.startswith('_') and name != '__init__') or (cls, name) in NOT_MIRRORED:
                continue
            if not hasattr(o, name):
                problems.append("%s.%s missing" % (cls, name))
                continue
            try:
                a = list(inspect.signature(fn).parameters)
                b = list(inspect.signature(getattr(o, name)).parameters)
            except (TypeError, ValueError):
                continue
            if b[:len(a)] != a:
                problems.append("%s.%s%s != %s" % (cls, name, b, a))
    assert not problems, problems


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box)")
def test_dsc_host_helpers_match_the_live_reference():
    """dsc_et.py:238-299 (generate_data incl. noise_on / gs / gp), :825-843, :845-870 on the same np.random stream."""
    ref_harness.load()
    from prosper.em.camodels.dsc_et import DSC_ET as R
    from prosper_b200.em.camodels.dsc_et import DSC_ET as O
    from oracle.common import DictAnneal
    states = np.array([-1., 0., 1., 2.])
    r, o = R(6, 4, 3, 2, states=states), O(6, 4, 3, 2, states=states)
    rng = np.random.RandomState(0)
    params = {'W': rng.randn(6, 4), 'pi': np.array([.1, .7, .1, .1]), 'sigma': 0.5}
    for kw in ({}, {'noise_on': False}, {'gs': rng.randint(-1, 3, size=(5, 4)).astype(float)},
               {'gs': rng.randint(-1, 3, size=(5, 3, 4)).astype(float), 'gp': rng.rand(5, 3, 4)}):
        np.random.seed(3)
        a = r.generate_data(params, 5, **kw)
        np.random.seed(3)
        b = o.generate_data(params, 5, **kw)
        assert a['s'].dtype == b['s'].dtype and np.array_equal(a['s'], b['s'])
        assert np.abs(a['y'] - b['y']).max() < 1e-14
    den, cand, lp, y = rng.rand(20), rng.randint(0, 4, size=(20, 3)), rng.randn(20, 9), rng.randn(20, 6)
    for nc in (0.0, 0.5, 1.0):
        an = DictAnneal(T=1.0, Ncut_factor=nc)
        a, b = r._get_sorted_data(20, an, 0.8, den, cand, lp, y), o._get_sorted_data(20, an, 0.8, den, cand, lp, y)
        assert a[0] == b[0] and all(np.array_equal(u, v) for u, v in zip(a[1:], b[1:]))
    assert abs(r.get_likelihood(6, 0.5, lp, 20) - o.get_likelihood(6, 0.5, lp, 20)) < 1e-12
    assert o.free_energy(params, {}) == 0.0 and o.gain(params, params) == 0.0


def test_dsc_generate_data_without_the_reference():
    from prosper_b200.em.camodels.dsc_et import DSC_ET
    m = DSC_ET(6, 4, 3, 2)
    rng = np.random.RandomState(1)
    params = {'W': rng.randn(6, 4), 'pi': np.array([.2, .6, .2]), 'sigma': 0.3}
    np.random.seed(0)
    d = m.generate_data(params, 50, noise_on=False)
    assert d['s'].dtype == np.int8 and set(np.unique(d['s'])) <= {-1, 0, 1}
    assert np.allclose(d['y'], d['s'].astype(float) @ params['W'].T)
    gs = rng.randint(-1, 2, size=(7, 4))
    d = m.generate_data(params, 7, noise_on=False, gs=gs)
    assert np.array_equal(d['s'], gs) and np.allclose(d['y'], gs @ params['W'].T)
    from scipy.special import logsumexp
    lp = rng.randn(9, 5)
    assert abs(m.get_likelihood(6, 0.3, lp, 9) - (-3 * np.log(2 * np.pi * 0.09) + logsumexp(lp, 1).sum() / 9)) < 1e-12


def test_bars_helpers(tmp_path, monkeypatch):
    """utils/barstest.py (barstest.py:8-98) and utils.create_output_path (utils/__init__.py:17-62), self-contained."""
    from prosper_b200.utils import barstest, create_output_path
    W = barstest.generate_bars_dict(10)
    assert W.shape == (25, 10) and (W.sum(0) == 5).all() and (W.reshape(5, 5, 10)[2, :, 2] == 1).all()
    assert (W.reshape(5, 5, 10)[:, 3, 8] == 1).all()
    np.random.seed(0)
    y = barstest.generate_bars_data(200, 5, 0.2).reshape(200, 5, 5)
    rows, cols = y.min(2) == 1, y.min(1) == 1
    assert ((y == 1) == (rows[:, :, None] | cols[:, None, :])).all() and 0.1 < rows.mean() < 0.3
    rng = np.random.RandomState(1)
    perm = rng.permutation(10)
    Wl = 10 * W[:, perm] + 0.5 * rng.randn(25, 10)
    found = barstest.find_permutation(Wl, 10 * W)
    assert (perm[found] == np.arange(10)).all() and (barstest.find_permutation2(Wl, 10 * W) == found).all()
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("SLURM_JOBID", "77")
    monkeypatch.delenv("PBS_JOBID", raising=False)
    assert create_output_path("run") == "output/run.d77/" and create_output_path("run") == "output/run.d77+1/"
    assert os.path.isdir(str(tmp_path / "output" / "run.d77+1"))


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box)")
def test_bars_helpers_match_the_live_reference():
    ref_harness.load()
    from prosper.utils import barstest as R
    from prosper_b200.utils import barstest as O
    for H in (4, 10, 16):
        for neg in (False, True):
            np.random.seed(1)
            a = R.generate_bars_dict(H, neg)
            np.random.seed(1)
            assert np.array_equal(a, O.generate_bars_dict(H, neg))
    np.random.seed(2)
    a = R.generate_bars_data(50, 5, 0.3)
    np.random.seed(2)
    assert np.array_equal(a, O.generate_bars_data(50, 5, 0.3))
    rng = np.random.RandomState(0)
    for _ in range(12):
        H = int(rng.choice([4, 6, 8, 10]))
        Wgt = 10 * O.generate_bars_dict(H)
        W = Wgt[:, rng.permutation(H)] + rng.randn(Wgt.shape[0], H) * rng.choice([0.1, 2.0, 6.0])
        assert np.array_equal(R.find_permutation(W, Wgt), O.find_permutation(W, Wgt))
        assert np.array_equal(R.find_permutation2(W, Wgt), O.find_permutation2(W, Wgt))
    Wgt = 10 * O.generate_bars_dict(6)
    W = np.concatenate([Wgt[:, ::-1] + rng.randn(9, 6), rng.randn(9, 3)], 1)
    assert np.array_equal(R.find_permutation(W, Wgt), O.find_permutation(W, Wgt))


def test_install_as_prosper_serves_every_import_of_the_reference_examples():
    """`prosper_b200.install_as_prosper()`: the import lines of the reference's examples/ resolve to this package."""
    import subprocess
    import sys
    code = r'''
import sys
sys.path.insert(0, %r)
import prosper_b200
prosper_b200.install_as_prosper()
from prosper.utils.barstest import generate_bars_dict
from prosper.em.annealing import LinearAnnealing
from prosper.utils.parallel import pprint, stride_data
from prosper.utils.datalog import dlog, StoreToH5, TextPrinter, StoreToTxt
from prosper.utils import create_output_path
from prosper.em import EM
from prosper.em.camodels.bsc_et import BSC_ET
from prosper.em.camodels.mca_et import MCA_ET
from prosper.em.camodels.mmca_et import MMCA_ET
from prosper.em.camodels.tsc_et import TSC_ET
from prosper.em.camodels.dsc_et import DSC_ET
from prosper.em.camodels.gsc_et import GSC
from prosper.em.mixturemodels.MoG import MoG
from prosper.em.mixturemodels.MoP import MoP
import prosper.em.camodels.bsc_et as m
assert m.__name__ == "prosper_b200.em.camodels.bsc_et" and BSC_ET.__module__ == m.__name__
model = BSC_ET(25, 10, 6, 3)                       # the parameter files of the examples build the model at import time
anneal = LinearAnnealing(50)
anneal['T'] = [(0, 2.), (.7, 1.)]
assert model.state_matrix.shape == (35, 6) and generate_bars_dict(10).shape == (25, 10)
print("ALIAS_OK")
''' % os.path.join(os.path.dirname(GOLDEN), "..")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ALIAS_OK" in out.stdout, out.stderr[-2000:]
