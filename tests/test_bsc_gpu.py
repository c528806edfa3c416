"""GPU parity tests of the BSC-ET hot path: CUDA (through the C ABI) vs golden vectors minted
from the reference, vs the NumPy oracle on seeded inputs, plus size-independent properties."""
import glob
import os

import numpy as np
import pytest

from helpers import bsc_problem, rel_err, cand_mismatch_gap
from oracle.bsc import BSC
from oracle.common import DictAnneal

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-8          # float64 path; north_star asks for <= 1e-5 relative


def model(D, H, Hp, g, **kw):
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return BSC_ET(D, H, Hp, g, **kw)


def keep_log():
    from prosper_b200.utils.datalog import dlog, Keep
    return dlog, dlog.set_handler('*', Keep)


def copy_params(p):
    return dict((k, (np.copy(v) if isinstance(v, np.ndarray) else v)) for k, v in p.items())


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "bsc_*.npz"))),
                         ids=lambda p: os.path.basename(p))
def test_bsc_against_reference_golden(path):
    g = np.load(path)
    D, H, Hp, gam = (int(v) for v in g['meta'])
    m = model(D, H, Hp, gam)
    an = DictAnneal(T=float(g['T']), Ncut_factor=float(g['Ncut_factor']), anneal_prior=bool(g['anneal_prior']))
    params = {'W': g['W0'].copy(), 'pi': float(g['pi0']), 'sigma': float(g['sigma0'])}
    data = m.select_Hprimes(params, {'y': g['y'].copy()})
    assert data['candidates'].dtype == np.int64
    assert np.array_equal(data['candidates'], g['candidates'])
    suff = m.E_step(an, params, data)
    assert 'mu' in params                                   # bsc_et.py:145-149 side effect
    assert suff['logpj'].shape == g['logpj'].shape
    assert np.abs(suff['logpj'] - g['logpj']).max() < 1e-10 * np.abs(g['logpj']).max()
    dlog, keep = keep_log()
    try:
        new = m.M_step(an, params, suff, data)
        assert rel_err(new['W'], g['W_new']) < TOL
        assert abs(new['pi'] - g['pi_new']) < TOL * g['pi_new']
        assert abs(new['sigma'] - g['sigma_new']) < TOL * g['sigma_new']
        assert abs(keep.last('L') - float(g['L'])) < TOL * abs(float(g['L']))
        assert keep.last('N_use') == int(g['N_use'])
        # fused path (what CAModel.step runs): nothing of size n*C is materialised
        newf = m._fused_step(an, {'W': g['W0'].copy(), 'pi': float(g['pi0']), 'sigma': float(g['sigma0'])}, {'y': g['y'].copy()})
        assert rel_err(newf['W'], g['W_new']) < TOL
        assert abs(newf['pi'] - g['pi_new']) < TOL * g['pi_new']
        assert abs(newf['sigma'] - g['sigma_new']) < TOL * g['sigma_new']
        assert abs(keep.last('L') - float(g['L'])) < TOL * abs(float(g['L']))
        assert keep.last('N_use') == int(g['N_use'])
    finally:
        dlog.remove_handler(keep)


CASES = [
    # D, H, H', gamma, N, seed, T, Ncut, anneal_prior
    (25, 10, 6, 3, 1000, 1, 1.0, 0.0, False),
    (25, 10, 6, 3, 1000, 1, 2.0, 1.0, False),
    (100, 50, 8, 3, 2000, 2, 1.0, 0.0, False),
    (100, 50, 8, 4, 2000, 2, 1.3, 1.0, True),
    (31, 17, 5, 5, 333, 4, 1.1, 0.4, False),          # odd D and H, gamma == H'
    (40, 12, 4, 1, 257, 6, 1.0, 0.0, False),          # gamma = 1: no multi-cause states at all
    (64, 16, 16, 2, 100, 7, 1.0, 0.0, False),         # H' == H
    (676, 1000, 12, 5, 192, 5, 1.0, 0.0, False),      # north-star shape, small N
    (676, 1000, 12, 5, 192, 5, 1.2, 1.0, False),
    (30, 1100, 6, 3, 96, 8, 1.0, 0.5, False),         # H > 1024: generic row kernel (no register-resident scores)
    (24, 18, 8, 6, 150, 9, 1.1, 0.0, False),          # gamma = 6: 8-member state records, direct state evaluation
    (24, 18, 8, 8, 120, 10, 1.0, 0.7, True),          # gamma = H' = 8: all 247 states
]


@pytest.mark.parametrize("D,H,Hp,gam,N,seed,T,ncut,ap", CASES)
def test_bsc_against_oracle(D, H, Hp, gam, N, seed, T, ncut, ap):
    bars = (D == 25)
    y, params, _ = bsc_problem(D, H, N, seed, bars=bars, pi=(0.2 if bars else None), sigma=(2.0 if bars else 1.0))
    an = DictAnneal(T=T, Ncut_factor=ncut, anneal_prior=ap)
    o = BSC(D, H, Hp, gam)
    od = o.select_hprimes(copy_params(params), {'y': y.copy()})
    p0 = copy_params(params)
    oss = o.e_step(an, p0, od)
    onew = o.m_step(an, p0, oss, od)

    m = model(D, H, Hp, gam)
    p1 = copy_params(params)
    d = m.select_Hprimes(p1, {'y': y.copy()})
    bad, gap = cand_mismatch_gap(od['_sim'], od['candidates'], d['candidates'])
    assert bad == 0 or gap < 1e-12, (bad, gap)              # sets identical unless scores tie
    d['candidates'] = od['candidates'].copy()               # line the columns up for the elementwise check
    ss = m.E_step(an, p1, d)
    assert np.abs(ss['logpj'] - oss['logpj']).max() < 1e-11 * np.abs(oss['logpj']).max()
    dlog, keep = keep_log()
    try:
        new = m.M_step(an, p1, {'logpj': oss['logpj']}, d)
        # N < H: Wq is rank deficient and the update is lstsq's minimum-norm solution (bsc_et.py:377-380), which amplifies
        # rounding differences of the statistics by up to 1 / rcond; the bound for W is the north-star 1e-5 scaled down
        tol_w = TOL if N >= H else 1e-6
        for got in (new, m._fused_step(an, copy_params(params), {'y': y.copy()})):
            assert rel_err(got['W'], onew['W']) < tol_w
            assert abs(got['pi'] - onew['pi']) < TOL * onew['pi']
            assert abs(got['sigma'] - onew['sigma']) < TOL * onew['sigma']
            assert abs(keep.last('L') - o.log['L']) < TOL * abs(o.log['L'])
            assert keep.last('N_use') == o.log['N_use']
    finally:
        dlog.remove_handler(keep)


@pytest.mark.parametrize("state_kernel", [1, 2], ids=["scalar-state-kernel", "tensor-state-kernel"])
def test_bsc_against_reference_golden_at_north_star_shape(state_kernel):
    """BASELINE configs[4] shape: the CUDA path against what the UNMODIFIED reference returned (bsc_et.py:98-438) for
    the 48 seeded datapoints of helpers.northstar_inputs; compat calls and the fused step, both state kernels."""
    from helpers import northstar_inputs
    g = np.load(os.path.join(GOLDEN, "northstar_bsc.npz"))
    y, p0 = northstar_inputs(int(g['N']), int(g['seed']))
    assert np.array_equal(y, g['y'])
    m = model(676, 1000, 12, 5)
    m.engine.set_state_kernel(state_kernel)
    dlog, keep = keep_log()
    try:
        for tag in ('a', 'b'):
            an = DictAnneal(T=float(g['T_' + tag]), Ncut_factor=float(g['Ncut_' + tag]), anneal_prior=False)
            params = {'W': p0['W'].copy(), 'pi': p0['pi'], 'sigma': p0['sigma']}
            outs = []
            if state_kernel == 1:
                data = m.select_Hprimes(params, {'y': y.copy()})
                suff = m.E_step(an, params, data)
                if tag == 'a':
                    assert np.array_equal(data['candidates'], g['candidates'])
                    assert np.abs(suff['logpj'] - g['logpj']).max() < 1e-10 * np.abs(g['logpj']).max()
                outs.append(m.M_step(an, params, suff, data))
            outs.append(m._fused_step(an, {'W': p0['W'].copy(), 'pi': p0['pi'], 'sigma': p0['sigma']}, {'y': y.copy()}))
            for new in outs:
                # N = 48 < H: Wq is rank deficient, the update is the minimum-norm solution (lstsq, bsc_et.py:377-380)
                assert rel_err(new['W'][::4], g['W_new_rows4_' + tag]) < 1e-6
                assert abs(new['pi'] - float(g['pi_new_' + tag])) < TOL * float(g['pi_new_' + tag])
                assert abs(new['sigma'] - float(g['sigma_new_' + tag])) < TOL * float(g['sigma_new_' + tag])
                assert abs(keep.last('L') - float(g['L_' + tag])) < TOL * abs(float(g['L_' + tag]))
                assert keep.last('N_use') == int(g['N_use_' + tag])
    finally:
        dlog.remove_handler(keep)


@pytest.mark.parametrize("ncut", [0.0, 1.0])
def test_bsc_north_star_shape_many_chunks_against_oracle(ncut, monkeypatch):
    """North-star shape with N = 1400 and 512-row chunks: three chunks, the last one ending in a partial 128-row tile
    (score GEMM tiles, state-kernel tiles, split-K statistics GEMM all cross chunk borders), N > H so that the update
    is a well-posed solve.  The oracle takes about half a minute on the host."""
    monkeypatch.setenv("PET_CHUNK_ROWS", "512")
    D, H, Hp, gam, N = 676, 1000, 12, 5, 1400
    y, params, _ = bsc_problem(D, H, N, 9)
    an = DictAnneal(T=1.1, Ncut_factor=ncut, anneal_prior=False)
    want = BSC(D, H, Hp, gam).step(an, copy_params(params), {'y': y.copy()})
    for kern in (2, 1):
        m = model(D, H, Hp, gam)
        m.engine.set_state_kernel(kern)
        got = m._fused_step(an, copy_params(params), {'y': y.copy()})
        assert rel_err(got['W'], want['W']) < TOL, kern
        assert abs(got['pi'] - want['pi']) < TOL * want['pi'], kern
        assert abs(got['sigma'] - want['sigma']) < TOL * want['sigma'], kern


def test_bsc_trajectory_50_iterations_tracks_oracle():
    """BASELINE configs[0]: bars test 5x5, H=10, H'=6, gamma=3, N=1000, 50 annealed EM iterations."""
    from prosper_b200.em import EM
    from prosper_b200.em.annealing import LinearAnnealing
    D, H, Hp, gam, N = 25, 10, 6, 3, 1000
    y, params, gt = bsc_problem(D, H, N, 1, bars=True, pi=0.2, sigma=2.0)
    anneal = LinearAnnealing(50)
    anneal['T'] = [(0, 2.), (.7, 1.)]
    anneal['Ncut_factor'] = [(0, 0.), (2. / 3, 1.)]
    anneal['anneal_prior'] = False
    o = BSC(D, H, Hp, gam)
    po = copy_params(params)
    m = model(D, H, Hp, gam)
    em = EM(model=m, anneal=anneal, data={'y': y.copy()}, lparams=copy_params(params))
    worst = 0.0
    while not anneal.finished:
        an = DictAnneal(**anneal.as_dict())
        po = o.step(an, po, {'y': y})
        new = m.step(anneal, em.lparams, em.data)
        anneal.next()
        em.lparams = new
        worst = max(worst, rel_err(new['W'], po['W']), abs(new['pi'] - po['pi']) / po['pi'],
                    abs(new['sigma'] - po['sigma']) / po['sigma'])
    assert worst < 1e-6, worst
    # bars are recovered: every ground-truth bar is matched by some learned column
    Wl, Wg = new['W'], gt['W']
    err = np.abs(Wl[:, :, None] - Wg[:, None, :]).mean(axis=0)
    assert (err.min(axis=0) < 1.0).all()
    assert abs(new['sigma'] - 2.0) < 0.2 and abs(new['pi'] - 0.2) < 0.05


_scale_results = {}


@pytest.mark.parametrize("state_kernel", [1, 2], ids=["scalar-state-kernel", "tensor-state-kernel"])
def test_properties_at_scale(state_kernel):
    """Size-independent properties at a size the oracle cannot reach (N = 40k, north-star shape):
    permutation invariance over datapoints, chunking invariance, truncation count, select idempotence -- for either
    state kernel, and the two kernels against each other."""
    from prosper_b200.em.camodels import Engine
    D, H, Hp, gam, N = 676, 1000, 12, 5, 40000
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev); g.manual_seed(11)
    Wg = torch.randn(D, H, dtype=torch.float64, device=dev, generator=g)
    Wg *= 10 / Wg.norm(dim=0, keepdim=True)
    s = (torch.rand(N, H, device=dev, generator=g) < 2.0 / H).to(torch.float64)
    y = s @ Wg.T + torch.randn(N, D, dtype=torch.float64, device=dev, generator=g)
    W0 = (y.mean(0)[:, None] + 0.3 * torch.randn(D, H, dtype=torch.float64, device=dev, generator=g)).cpu().numpy()
    params = {'W': W0, 'pi': 1. / H, 'sigma': 1.2}
    an = DictAnneal(T=1.1, Ncut_factor=1.0, anneal_prior=False)
    m = model(D, H, Hp, gam)
    m.engine.set_state_kernel(state_kernel)
    dlog, keep = keep_log()
    try:
        a = m._fused_step(an, copy_params(params), {'y': y})
        L_a, nuse_a = keep.last('L'), keep.last('N_use')
        perm = torch.randperm(N, device=dev, generator=g)
        b = m._fused_step(an, copy_params(params), {'y': y[perm].contiguous()})
        L_b, nuse_b = keep.last('L'), keep.last('N_use')
        m2 = model(D, H, Hp, gam)
        m2._engine = Engine(m2.model_kind, D, H, Hp, gam, chunk_rows=4096)      # 10 chunks instead of 1
        m2.engine.set_state_kernel(state_kernel)
        c = m2._fused_step(an, copy_params(params), {'y': y})
        L_c, nuse_c = keep.last('L'), keep.last('N_use')
    finally:
        dlog.remove_handler(keep)
    for other, L_o, n_o in ((b, L_b, nuse_b), (c, L_c, nuse_c)):
        assert rel_err(other['W'], a['W']) < 1e-9
        assert abs(other['pi'] - a['pi']) < 1e-11 * a['pi'] and abs(other['sigma'] - a['sigma']) < 1e-11 * a['sigma']
        assert abs(L_o - L_a) < 1e-11 * abs(L_a) and n_o == nuse_a
    _scale_results[state_kernel] = (a, L_a, nuse_a)
    if len(_scale_results) == 2:        # scalar FP64 state kernel against the int8 tensor-core one
        (a1, L1, n1), (a2, L2, n2) = _scale_results[1], _scale_results[2]
        assert rel_err(a2['W'], a1['W']) < TOL
        assert abs(a2['pi'] - a1['pi']) < 1e-9 * a1['pi'] and abs(a2['sigma'] - a1['sigma']) < 1e-9 * a1['sigma']
        assert abs(L2 - L1) < 1e-11 * abs(L1) and n1 == n2
    # truncation keeps at least the model-predicted number of datapoints (bsc_et.py:250-257, '>=' rule)
    A, _ = m._AB(params['pi'])
    target = int(N * (1 - (1 - A) * 1.0))
    assert target <= nuse_a <= N
    # select is deterministic and its output is a valid index set, best candidate last
    c1 = m.select_Hprimes(copy_params(params), {'y': y})['candidates']
    c2 = m.select_Hprimes(copy_params(params), {'y': y})['candidates']
    assert np.array_equal(c1, c2) and c1.min() >= 0 and c1.max() < H
    assert all(len(set(r)) == Hp for r in c1[:500].tolist())


def test_edge_cases():
    m = model(25, 10, 6, 3)
    y, params, _ = bsc_problem(25, 10, 1, 3, bars=True, pi=0.2, sigma=2.0)
    params['sigma'] = 2.5                                            # (std of one datapoint is 0)
    params['W'] = params['W'] + np.random.RandomState(0).normal(size=(25, 10))
    an = DictAnneal(T=1.0, Ncut_factor=0.0, anneal_prior=False)
    o = BSC(25, 10, 6, 3)
    want = o.step(an, copy_params(params), {'y': y.copy()})           # a single datapoint
    got = m._fused_step(an, copy_params(params), {'y': y.copy()})
    assert abs(got['sigma'] - want['sigma']) < TOL * want['sigma'] and abs(got['pi'] - want['pi']) < TOL * want['pi']
    # a dead unit (never a candidate, singleton posterior underflows to exactly 0) leaves a zero row/column in Wq:
    # np.linalg.lstsq returns the minimum-norm answer (zero column of W), and so must the device path
    y2, p2, _ = bsc_problem(25, 10, 400, 5, bars=True, pi=0.2, sigma=2.0)
    p2['W'][:, 3] = 1e3
    want2 = BSC(25, 10, 6, 3).step(an, copy_params(p2), {'y': y2.copy()})
    m2 = model(25, 10, 6, 3)
    got2 = m2._fused_step(an, copy_params(p2), {'y': y2.copy()})
    assert m2.last_dropped_pivots == 1 and np.abs(got2['W'][:, 3]).max() < 1e-10 and np.abs(want2['W'][:, 3]).max() < 1e-10
    assert rel_err(got2['W'], want2['W']) < TOL and abs(got2['pi'] - want2['pi']) < TOL * want2['pi']
    # wrong feature dimension is rejected like the reference's shape asserts
    with pytest.raises(AssertionError):
        m.select_Hprimes(copy_params(params), {'y': np.zeros((5, 24))})
    # E-step before any candidates exist -> explicit error, not garbage
    from prosper_b200._lib import PetError
    m3 = model(25, 10, 6, 3)
    m3._bind({'y': np.zeros((4, 25))})
    with pytest.raises(PetError):
        m3.engine.e_step(m3.engine.anneal(an), m3._pack_params(copy_params(params)))
    # invalid constructor arguments (camodels/__init__.py:90-91)
    with pytest.raises(AssertionError):
        model(25, 10, 11, 3)
    with pytest.raises(AssertionError):
        model(25, 10, 6, 7)


def test_em_run_writes_result_h5(tmp_path):
    """EM.run() with the StoreToH5 handler leaves a result.h5 whose tables have one row per
    iteration (`/W (iters, D, H)`, `/pi`, `/sigma`, `/L`, `/N_use`, annealing keys), as bars-learning.py:61-69."""
    from prosper_b200.em import EM
    from prosper_b200.em.annealing import LinearAnnealing
    from prosper_b200.utils.datalog import dlog, StoreToH5
    from prosper_b200.utils import h5min
    D, H, Hp, gam, N, iters = 25, 10, 6, 3, 400, 6
    y, params, gt = bsc_problem(D, H, N, 3, bars=True, pi=0.2, sigma=2.0)
    anneal = LinearAnnealing(iters)
    anneal['T'] = [(0, 2.), (.7, 1.)]
    anneal['Ncut_factor'] = [(0, 0.), (2. / 3, 1.)]
    anneal['anneal_prior'] = False
    path = str(tmp_path / "result.h5")
    h = dlog.set_handler(('W', 'pi', 'sigma', 'mu', 'L', 'N_use', 'T', 'Ncut_factor'), StoreToH5, path)
    try:
        em = EM(model=model(D, H, Hp, gam), anneal=anneal, data={'y': y.copy()}, lparams=copy_params(params))
        em.run()
    finally:
        h.close()
        dlog.remove_handler(h)
    r = h5min.read_h5(path)
    assert r['W'].shape == (iters, D, H) and r['W'].dtype == np.float64
    for k in ('pi', 'sigma', 'L', 'N_use', 'T', 'Ncut_factor'):
        assert r[k].shape == (iters,), k
    assert r['mu'].shape == (iters, D)
    assert np.array_equal(r['W'][-1], em.lparams['W']) and r['sigma'][-1] == em.lparams['sigma']
    assert r['T'][0] == 2.0 and r['T'][-1] == 1.0
    assert np.isfinite(r['L']).all() and r['L'][-1] > r['L'][0]


@pytest.mark.parametrize("ncut", [0.0, 0.6])
def test_bsc_learns_mu_like_the_reference(ncut):
    """bsc_et.py:422-430 with 'mu' in to_learn, from a non-zero mu, with and without truncation; two iterations so
    that the second runs on the shard re-shifted by the learned mu."""
    D, H, Hp, gam, N = 25, 10, 6, 3, 500
    y, params, _ = bsc_problem(D, H, N, 9, bars=True, pi=0.2, sigma=2.0)
    rng = np.random.RandomState(3)
    y = y + 1.5                                              # a real offset to learn
    params['mu'] = rng.randn(D) * 0.3
    an = DictAnneal(T=1.3, Ncut_factor=ncut, anneal_prior=False)
    learn = ['W', 'pi', 'sigma', 'mu']
    o = BSC(D, H, Hp, gam, to_learn=learn)
    m = model(D, H, Hp, gam, to_learn=learn)
    po, pm = copy_params(params), copy_params(params)
    for it in range(2):
        po = o.step(an, po, {'y': y.copy()})
        pm = m._fused_step(an, pm, {'y': y})
        assert rel_err(pm['mu'], po['mu']) < TOL and rel_err(pm['W'], po['W']) < TOL
        assert abs(pm['sigma'] - po['sigma']) < TOL * po['sigma'] and abs(pm['pi'] - po['pi']) < TOL * po['pi']
    assert np.abs(pm['mu']).max() > 0.5                      # it moved towards the offset


@pytest.mark.parametrize("slices", ["0", "6", "7", "default"])
def test_gemm_fallback_paths_agree_with_oracle(monkeypatch, slices):
    """The FP64 DMMA kernels (PET_OZAKI=0: what runs when the int8 slices do not fit in memory), the all-6-slice and
    all-7-slice int8 variants and the default (score GEMM 6 slices, statistics GEMM 7) go through the same pipeline; all
    must reproduce the oracle (6 slices: ~1e-12 per product)."""
    if slices == "0":
        monkeypatch.setenv("PET_OZAKI", "0")
    elif slices != "default":
        monkeypatch.setenv("PET_OZAKI_SLICES", slices)
    D, H, Hp, gam, N = 40, 24, 8, 4, 3000
    y, params, _ = bsc_problem(D, H, N, 11)
    an = DictAnneal(T=1.2, Ncut_factor=0.4, anneal_prior=False)
    o = BSC(D, H, Hp, gam)
    onew = o.step(an, copy_params(params), {'y': y.copy()})
    m = model(D, H, Hp, gam)
    got = m._fused_step(an, copy_params(params), {'y': y.copy()})
    want = {"0": (0, 0), "6": (6, 6), "7": (7, 7), "default": (6, 7)}[slices]
    assert m.engine.gemm_slices() == want                           # known once a shard is bound
    assert m.engine.gemm_path() == want[1]
    tol = 1e-7 if slices == "6" else TOL
    assert rel_err(got['W'], onew['W']) < tol
    assert abs(got['pi'] - onew['pi']) < tol * onew['pi'] and abs(got['sigma'] - onew['sigma']) < tol * onew['sigma']


def test_device_resident_parameters_through_em_run(tmp_path):
    """with W given as a CUDA tensor the EM loop never copies it to the host unless a dlog handler asks;
    the trajectory equals the host-parameter one."""
    from prosper_b200.em import EM
    from prosper_b200.em.annealing import LinearAnnealing
    from prosper_b200.utils.datalog import dlog, Keep
    D, H, Hp, gam, N, iters = 25, 10, 6, 3, 600, 5
    y, params, _ = bsc_problem(D, H, N, 5, bars=True, pi=0.2, sigma=2.0)

    def run(p0, listen):
        anneal = LinearAnnealing(iters)
        anneal['T'] = [(0, 2.), (.7, 1.)]
        anneal['Ncut_factor'] = [(0, 0.), (2. / 3, 1.)]
        anneal['anneal_prior'] = False
        keep = dlog.set_handler(listen, Keep) if listen else None
        try:
            em = EM(model=model(D, H, Hp, gam), anneal=anneal, data={'y': y.copy()}, lparams=p0)
            em.run()
        finally:
            if keep is not None:
                dlog.remove_handler(keep)
        return em.lparams, keep

    host, _ = run(copy_params(params), None)
    pd = copy_params(params)
    pd['W'] = torch.as_tensor(pd['W']).cuda()
    devp, keep = run(pd, ('W', 'sigma'))
    assert isinstance(devp['W'], torch.Tensor) and devp['W'].is_cuda
    assert rel_err(devp['W'].cpu().numpy(), host['W']) < 1e-12 and abs(devp['sigma'] - host['sigma']) < 1e-12 * host['sigma']
    assert len(keep.values['W']) == iters and isinstance(keep.values['W'][-1], np.ndarray)      # the listener got host arrays
