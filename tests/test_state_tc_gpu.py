"""GPU parity tests of the tensor-core state kernel (csrc/gl_state_tc.cu) behind the fused BSC step: against the
NumPy oracle (bsc_et.py:98-438 restated in oracle/bsc.py) and against the scalar FP64 state kernel it replaces, on
the same seeded inputs.  Shapes cover partial 128-datapoint tiles, partial 64-state chunks of every state size,
gamma = 2..5, H' up to 12, annealing of the prior, and the truncated iteration (log-denominator sweep + cut)."""
import numpy as np
import pytest

from helpers import bsc_problem, rel_err
from oracle.bsc import BSC
from oracle.common import DictAnneal

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL = 1e-8


def model(D, H, Hp, g, mode):
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    assert torch.cuda.is_available(), "GPU tests need a B200"
    m = BSC_ET(D, H, Hp, g)
    m.engine.set_state_kernel(mode)
    assert m.engine.state_kernel_path() == mode
    return m


def copy_params(p):
    return dict((k, (np.copy(v) if isinstance(v, np.ndarray) else v)) for k, v in p.items())


CASES = [
    # D, H, H', gamma, N, seed, T, Ncut, anneal_prior
    (25, 10, 6, 3, 1000, 1, 1.0, 0.0, False),         # bars: 15 + 20 states, two partial chunks
    (25, 10, 6, 3, 1000, 1, 2.0, 1.0, False),
    (100, 50, 8, 4, 2000, 2, 1.3, 1.0, True),
    (31, 17, 5, 5, 333, 4, 1.1, 0.4, False),          # gamma == H', N not a multiple of 128
    (40, 30, 12, 2, 130, 11, 1.0, 0.0, False),        # pairs only: 66 states = one full + one 2-state chunk
    (40, 30, 10, 5, 77, 12, 1.0, 0.0, False),         # a single partial tile
    (60, 40, 12, 5, 300, 13, 1.4, 0.7, True),         # the north-star state space (1573 states) on a small model
    (676, 1000, 12, 5, 192, 5, 1.0, 0.0, False),      # north-star shape, small N
    (676, 1000, 12, 5, 192, 5, 1.2, 1.0, False),
    (64, 24, 9, 3, 4500, 14, 1.0, 0.0, False),        # more tiles than fit one wave of a small grid
]


@pytest.mark.parametrize("D,H,Hp,gam,N,seed,T,ncut,ap", CASES)
def test_tensor_state_kernel_against_oracle_and_scalar_kernel(D, H, Hp, gam, N, seed, T, ncut, ap):
    bars = (D == 25)
    y, params, _ = bsc_problem(D, H, N, seed, bars=bars, pi=(0.2 if bars else None), sigma=(2.0 if bars else 1.0))
    an = DictAnneal(T=T, Ncut_factor=ncut, anneal_prior=ap)
    o = BSC(D, H, Hp, gam)
    want = o.step(an, copy_params(params), {'y': y.copy()})
    got = {}
    lse = {}
    stats = {}
    for mode in (1, 2):
        m = model(D, H, Hp, gam, mode)
        got[mode] = m._fused_step(an, copy_params(params), {'y': y.copy()})
        eng = m.engine
        p = m._pack_params(copy_params(params))
        a = eng.anneal(an)
        from prosper_b200 import _lib
        lse[mode] = eng.log_denominators(a, p, None, _lib.PASS_SELECT).cpu().numpy().copy()
        stats[mode] = eng.m_step_stats(a, p, None, _lib.PASS_SELECT).cpu().numpy().copy()
    # N < H: rank-deficient Wq, minimum-norm update (see test_bsc_gpu.test_bsc_against_oracle)
    tol_w = TOL if N >= H else 1e-6
    for mode in (1, 2):
        assert rel_err(got[mode]['W'], want['W']) < tol_w, mode
        assert abs(got[mode]['pi'] - want['pi']) < TOL * want['pi'], mode
        assert abs(got[mode]['sigma'] - want['sigma']) < TOL * want['sigma'], mode
    # the two kernels against each other: log-denominators and the packed statistics (Wp, Wq, scalars)
    assert np.abs(lse[2] - lse[1]).max() < 1e-11 * max(1.0, np.abs(lse[1]).max())
    assert rel_err(stats[2], stats[1]) < 1e-9
    assert rel_err(got[2]['W'], got[1]['W']) < (1e-9 if N >= H else 1e-6)


def test_tensor_state_kernel_is_the_default_for_large_shards_of_the_north_star_state_space():
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    m = BSC_ET(60, 40, 12, 5)
    assert m.engine.state_kernel_path() == 2                            # no shard bound yet: the choice for a large one
    m._bind({'y': np.zeros((300, 60))})
    assert m.engine.state_kernel_path() == 1                            # 3 tiles would leave 145 SMs idle: scalar kernel
    m._bind({'y': np.zeros((20000, 60))})
    assert m.engine.state_kernel_path() == 2
    assert BSC_ET(25, 10, 6, 3).engine.state_kernel_path() == 1          # 35 states: the scalar kernel
    from prosper_b200 import _lib
    m = BSC_ET(24, 18, 8, 6)                                             # gamma = 6 has no tensor-core kernel
    with pytest.raises(_lib.PetError):
        m.engine.set_state_kernel(2)


def test_tensor_state_kernel_heavy_tailed_features():
    """Features spanning 1e6 inside a datapoint (one dominant cause): the per-datapoint power-of-two scale keeps 49
    bits below the LARGEST feature, so the log-joints stay accurate to ~1e-14 of that scale."""
    D, H, Hp, gam, N = 48, 20, 10, 4, 400
    y, params, _ = bsc_problem(D, H, N, 21)
    params['W'][:, 3] *= 300.0                       # ||W_3||^2 ~ 1e5 x the others
    y[::7] += 0.05 * params['W'][:, 3]
    params['sigma'] = float(np.sqrt(((y - y.mean(0)) ** 2).mean()))      # keeps the reference's unshifted exp finite
    an = DictAnneal(T=1.0, Ncut_factor=0.0, anneal_prior=False)
    want = BSC(D, H, Hp, gam).step(an, copy_params(params), {'y': y.copy()})
    got = model(D, H, Hp, gam, 2)._fused_step(an, copy_params(params), {'y': y.copy()})
    assert rel_err(got['W'], want['W']) < 1e-7
    assert abs(got['pi'] - want['pi']) < 1e-7 * want['pi']
    assert abs(got['sigma'] - want['sigma']) < 1e-7 * want['sigma']


def test_truncated_iteration_evaluates_the_posterior_once(monkeypatch):
    """bsc_et.py:250-257 needs every log-denominator before the cut is known.  With the tensor-core state kernel the
    log-denominator sweep parks the per-datapoint statistics and the statistics sweep only adds up the datapoints that
    stay (GLF_DEFER_STATS + gl_finalize_cut): same result as the two-evaluation form and as the oracle, over several
    chunks with a ragged last one, and the row / state kernels run once per chunk instead of twice."""
    D, H, Hp, gam, N = 60, 40, 12, 5, 1400
    y, params, _ = bsc_problem(D, H, N, 21)
    an = DictAnneal(T=1.1, Ncut_factor=0.8, anneal_prior=False)
    want = BSC(D, H, Hp, gam).step(an, copy_params(params), {'y': y.copy()})
    got, spans = {}, {}
    for single in (True, False):
        if single:
            monkeypatch.delenv("PET_GL_NO_SINGLE_EVAL", raising=False)
        else:
            monkeypatch.setenv("PET_GL_NO_SINGLE_EVAL", "1")
        monkeypatch.setenv("PET_CHUNK_ROWS", "512")
        m = model(D, H, Hp, gam, 2)
        m._bind({'y': y})
        m.engine.enable_timing(True)
        got[single] = m._fused_step(an, copy_params(params), {'y': y})
        st = m.engine.stage_times()
        m.engine.enable_timing(False)
        spans[single] = (st['row_kernel']['spans'], st['state_kernel']['spans'], st['score_gemm']['spans'])
    for single in (True, False):
        assert rel_err(got[single]['W'], want['W']) < TOL, single
        assert abs(got[single]['pi'] - want['pi']) < TOL * want['pi'], single
        assert abs(got[single]['sigma'] - want['sigma']) < TOL * want['sigma'], single
    assert rel_err(got[True]['W'], got[False]['W']) < 1e-10
    nchunks = spans[True][2]
    assert nchunks >= 3                                                   # several chunks, the last one ragged
    assert spans[False][0] == 2 * nchunks and spans[True][0] == nchunks   # one selection + singleton pass per chunk
    assert spans[True][1] == 2 * nchunks                                  # state kernel once, then gl_finalize_cut
