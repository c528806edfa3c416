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
    return sorted(p for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not os.path.basename(p).startswith(("gsc_", "infer_", "mix_")))


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
