"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): NumPy restatement of CAModel.inference.

Follows prosper/em/camodels/__init__.py:255-375 (base class: BSC, MCA, MMCA), tsc_et.py:546-680 and
dsc_et.py:927-1058 on top of the oracle's own select_hprimes / e_step.  `make(Hprime, gamma)` builds
the oracle model for a grown truncation (the reference mutates self.Hprime / self.gamma and regenerates
the state matrix in place, :357-368).  Pinned to outputs of the reference in tests/golden/infer_*.npz.
"""
import numpy as np
from scipy.special import logsumexp


def inference(make, Hprime, gamma, anneal, params, y, topK=10, logprob=False, adaptive=True, Hprime_max=None,
              gamma_max=None, abs_marginal=True):
    model = make(Hprime, gamma)
    kind = model.name
    if hasattr(model, 'check_params'):
        params = model.check_params(params)
    my_N, H = y.shape[0], model.H
    if topK == -1:
        topK = model.state_matrix.shape[0]
    res = {'s': np.zeros((my_N, topK, H), dtype=np.int8), 'm': np.zeros((my_N, H)), 'p': np.zeros((my_N, topK)),
           'gamma': np.zeros((my_N,)), 'Hprime': np.zeros((my_N,))}
    if kind == 'tsc':
        res['am'] = np.zeros((my_N, H))
    which = np.ones(my_N, dtype=bool)
    y_tmp = y
    while which.any():
        ind_n = np.where(which)[0]
        if kind == 'gsc':                                          # gsc_et.py:811-944 overrides compute_lpj only
            logpj, cand = model.compute_lpj(params, {'y': y_tmp})
        else:
            data = model.select_hprimes(params, {'y': y_tmp})
            logpj = model.e_step(anneal, params, data)['logpj']
            cand = np.asarray(data['candidates']).astype(np.int64)
        corr = logpj.max(axis=1)                                   # :307
        logpjc = logpj - corr[:, None]
        pjc = np.exp(logpjc)                                       # :309 (before the normalisation below)
        denomc = pjc.sum(axis=1)
        logpjc = logpjc + (-np.log(denomc))[:, None]               # :311
        idx = np.argsort(logpjc, axis=-1)[:, ::-1]                 # :312
        SM = model.state_matrix
        for n in range(len(ind_n)):
            n_ = ind_n[n]
            res['Hprime'][n_] = model.Hprime
            res['gamma'][n_] = model.gamma
            for m in range(topK):
                t = idx[n, m]
                if kind == 'tsc':
                    res['p'][n_, m] = logpjc[n, t] if logprob else pjc[n, t] / denomc[n]      # tsc_et.py:629-632
                    res['s'][n_, m, cand[n, :]] = SM[t]
                    continue
                res['p'][n_, m] = logpjc[n, t] if logprob else pjc[n, t]                      # :321-324
                if kind == 'dsc':
                    nb = (model.K - 1) * H
                    if t == 0:
                        pass
                    elif t < nb + 1:                                                            # dsc_et.py:1004-1007
                        res['s'][n_, m, (t - 1) % H] = model.single_state_matrix[t - 1, (t - 1) % H]
                    else:
                        res['s'][n_, m, cand[n, :]] = SM[t - nb - 1]
                else:
                    if t == 0:
                        pass
                    elif t < H + 1:
                        res['s'][n_, m, t - 1] = 1                                              # :327-328
                    else:
                        res['s'][n_, m, cand[n, :]] = SM[t - H - 1]                             # :330-331
            if kind == 'tsc':
                res['m'][n_, cand[n]] = (pjc[n][:, None] * SM / denomc[n]).sum(0)               # tsc_et.py:636
                if abs_marginal:
                    res['am'][n_, cand[n]] = (pjc[n][:, None] * np.abs(SM) / denomc[n]).sum(0)
            else:
                off = ((model.K - 1) * H + 1) if kind == 'dsc' else (H + 1)
                for h in range(H):                                                              # :336-342
                    if h in cand[n, :]:
                        j = np.where(cand[n] == h)[0][0]
                        logp = np.hstack([logpjc[n, h + 1], logpjc[n, off:][SM[:, j] == 1]])
                        res['m'][n_, h] = logsumexp(logp)
                    else:
                        res['m'][n_, h] = logpjc[n, h + 1]
        if not adaptive:
            break
        if kind == 'tsc':
            which = (res['s'][:, 0, :].astype(bool) != 0).sum(-1) == model.gamma
        else:
            which = (res['s'][:, 0, :] != 0).sum(-1) == model.gamma
        if not which.any():
            break
        if (Hprime_max is not None and model.Hprime == Hprime_max) and (gamma_max is not None and model.gamma == gamma_max):
            break
        y_tmp = y[which]
        hp, g = model.Hprime, model.gamma
        if not (hp == H or (Hprime_max is not None and hp == Hprime_max)):
            hp += 1
        if not (g == H or (gamma_max is not None and g == gamma_max)):
            g += 1
        model = make(hp, g)
    if kind == 'tsc':
        if logprob:
            with np.errstate(divide='ignore', invalid='ignore'):
                res['m'] = np.log(res['m'])
                res['am'] = np.log(res['am'])
    elif not logprob:
        res['m'] = np.exp(res['m'])
    return res
