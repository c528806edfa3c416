"""Device-side data generation and standard_init: distributional checks (the reference uses
np.random's Mersenne Twister; stream parity is not required) and exact structural properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def bars(H):
    R = H // 2
    W = np.zeros((R * R, H))
    for i in range(R):
        W[:, i] = np.eye(R)[i].repeat(R)
        W[:, R + i] = np.tile(np.eye(R)[i], R)
    return W


def test_generate_bsc_statistics_and_shard_consistency():
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    D, H, N = 25, 10, 200000
    m = BSC_ET(D, H, 6, 3)
    gt = {'W': 10 * bars(H), 'pi': 0.2, 'sigma': 2.0}
    d = m.generate_data_device(gt, N, seed=11)
    y, s = d['y'], d['s']
    assert y.is_cuda and y.shape == (N, D) and s.shape == (N, H) and s.dtype == torch.int8
    assert set(np.unique(s.cpu().numpy())) == {0, 1}
    assert abs(float(s.double().mean()) - 0.2) < 4 * np.sqrt(0.2 * 0.8 / (N * H))
    resid = y - s.double() @ torch.as_tensor(gt['W']).to(y.device).T               # exactly the noise
    assert abs(float(resid.mean())) < 4 * 2.0 / np.sqrt(N * D)
    assert abs(float(resid.std()) - 2.0) < 0.01
    k = float(((resid / 2.0) ** 4).mean())                                         # Gaussian kurtosis 3
    assert abs(k - 3.0) < 0.05
    # independence across elements: lag-1 correlation of the noise along both axes
    r = (resid / 2.0).cpu().numpy()
    assert abs((r[:, 1:] * r[:, :-1]).mean()) < 0.003 and abs((r[1:] * r[:-1]).mean()) < 0.003
    # the same seed reproduces; shards with a row offset equal the rows of the big call; another seed differs
    d2 = m.generate_data_device(gt, N, seed=11)
    assert torch.equal(d2['y'], y)
    part = m.generate_data_device(gt, 1000, seed=11, row0=5000)
    assert torch.equal(part['y'], y[5000:6000]) and torch.equal(part['s'], s[5000:6000])
    assert not torch.equal(m.generate_data_device(gt, 1000, seed=12)['y'], y[:1000])


def test_generate_tsc_dsc_mca_laws():
    from prosper_b200.em.camodels.tsc_et import TSC_ET
    from prosper_b200.em.camodels.dsc_et import DSC_ET
    from prosper_b200.em.camodels.mca_et import MCA_ET
    N = 100000
    W12 = 10 * bars(12)
    t = TSC_ET(36, 12, 6, 3).generate_data_device({'W': W12, 'pi': 0.2, 'sigma': 1.0}, N, seed=3)
    s = t['s'].cpu().numpy()
    for v, p in ((-1, 0.1), (0, 0.8), (1, 0.1)):
        assert abs((s == v).mean() - p) < 0.003
    states = np.array([-2., 0., 1., 3.])
    pi = np.array([.05, .8, .1, .05])
    dd = DSC_ET(36, 12, 6, 3, states).generate_data_device({'W': W12, 'pi': pi, 'sigma': 0.0}, N, seed=4)
    s = dd['s'].cpu().numpy()
    for v, p in zip(states, pi):
        assert abs((s == v).mean() - p) < 0.003
    assert torch.allclose(dd['y'], dd['s'].double() @ torch.as_tensor(W12).to(dd['y'].device).T, atol=1e-12)   # sigma = 0
    # MCA: per pixel the entry of largest magnitude among the active causes (mca_et.py:83-85)
    Wm = np.abs(np.random.RandomState(0).randn(25, 10)) + 0.1
    mm = MCA_ET(25, 10, 6, 3).generate_data_device({'W': Wm, 'pi': 0.3, 'sigma': 0.0}, 5000, seed=5)
    s = mm['s'].cpu().numpy().astype(np.float64)
    ref = (s[:, None, :] * Wm[None, :, :]).max(axis=2)
    assert np.allclose(mm['y'].cpu().numpy(), ref, atol=1e-12)


def test_standard_init_on_device_matches_host_formula():
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    D, H, N = 25, 10, 50000
    m = BSC_ET(D, H, 6, 3)
    gt = {'W': 10 * bars(H), 'pi': 0.2, 'sigma': 2.0}
    y = m.generate_data_device(gt, N, seed=21, latents=False)['y']
    np.random.seed(5)
    p = m.standard_init({'y': y})
    yh = y.cpu().numpy()
    W_mean = yh.mean(axis=0)
    sigma_init = np.sqrt(((yh - W_mean) ** 2).mean(axis=0)).sum() / D              # camodels/__init__.py:217-221
    assert abs(p['sigma'] - sigma_init) < 1e-12 * sigma_init and p['pi'] == 1. / H
    noise = p['W'] - W_mean[:, None]
    assert p['W'].shape == (D, H) and isinstance(p['W'], np.ndarray)
    assert abs(noise.mean()) < 4 * (sigma_init / 4) / np.sqrt(D * H)
    assert abs(noise.std() - sigma_init / 4) < 0.15 * sigma_init / 4
    np.random.seed(5)
    again = m.standard_init({'y': y})['W']                                          # seeded through np.random on rank 0
    assert np.allclose(again, p['W'], rtol=0, atol=1e-12)                           # (column sums use atomics: last-bit noise)


def test_select_partial_data_on_the_device():
    """camodels/__init__.py:125-152 for a shard that lives on the device: ceil(partial N) distinct rows in ascending
    order, gathered by pet_gather_rows; reproducible from np.random's seed; the host path is untouched."""
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    from oracle.common import DictAnneal
    m = BSC_ET(20, 8, 4, 2)
    dev = torch.device('cuda', 0)
    y = torch.arange(1000 * 20, dtype=torch.float64, device=dev).reshape(1000, 20)
    an = DictAnneal(partial=0.3)
    np.random.seed(3)
    a = m.select_partial_data(an, {'y': y})['y']
    np.random.seed(3)
    b = m.select_partial_data(an, {'y': y})['y']
    assert a.is_cuda and a.shape == (300, 20) and torch.equal(a, b)
    rows = (a[:, 0] / 20).round().long()
    assert torch.equal(a, y[rows]) and bool((rows[1:] > rows[:-1]).all())
    c = m.select_partial_data(an, {'y': y})['y']
    assert not torch.equal(a, c)
    assert m.select_partial_data(DictAnneal(partial=1.0), {'y': y})['y'] is y
    yh = y.cpu().numpy()
    np.random.seed(5)
    h = m.select_partial_data(an, {'y': yh})['y']
    np.random.seed(5)
    sel = np.sort(np.random.permutation(1000)[:300])
    assert np.array_equal(h, yh[sel])


def test_noisify_params_on_the_device():
    """em/__init__.py:63-107 for a parameter matrix that lives on the device: noise of the annealed scale from the
    counter-based generator (every rank draws the same), bounds of the noise policy applied, host scalars as before."""
    from prosper_b200.em.camodels.bsc_et import BSC_ET
    from oracle.common import DictAnneal
    m = BSC_ET(64, 32, 4, 2)
    dev = torch.device('cuda', 0)
    W = torch.zeros((64, 32), dtype=torch.float64, device=dev)
    an = DictAnneal(W_noise=0.5)
    np.random.seed(1)
    p1 = m.noisify_params({'W': W, 'pi': 0.1, 'sigma': 1.0}, an)
    np.random.seed(1)
    p2 = m.noisify_params({'W': W, 'pi': 0.1, 'sigma': 1.0}, an)
    assert p1['W'].is_cuda and torch.equal(p1['W'], p2['W']) and p1['pi'] == 0.1 and p1['sigma'] == 1.0
    assert abs(float(p1['W'].std()) - 0.5) < 0.05 and abs(float(p1['W'].mean())) < 0.05
    m.noise_policy['W'] = (-0.2, 0.3, False)
    p3 = m.noisify_params({'W': W, 'pi': 0.1, 'sigma': 1.0}, an)
    assert float(p3['W'].min()) >= -0.2 and float(p3['W'].max()) <= 0.3
    assert m.noisify_params({'W': W, 'pi': 0.1, 'sigma': 1.0}, DictAnneal())['W'] is W
