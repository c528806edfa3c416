"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): NumPy restatement of the mixture models.

Follows prosper/em/mixturemodels/MoG.py (posterior :208-218, log_p_y :221-262, M_step :143-198) and MoP.py
(posterior :165-175, log_p_y :178-217, normalize :236-244, M_step :102-157), vectorised over datapoints.
Pinned to outputs of the reference in tests/golden/mix_*.npz.
"""
import numpy as np

TINY = np.finfo(np.float64).tiny
EPS = np.finfo(np.float64).eps


def _posterior(logp, pies, beta, H):
    lp = logp + np.log(pies)[None, :] * beta
    with np.errstate(over='ignore', under='ignore'):
        post = np.exp(lp)
    post[np.isnan(post)] = TINY
    post[post < TINY] = TINY
    post[np.isinf(post)] = np.finfo(np.float64).max / H
    return {'posteriors_h': post / post.sum(1)[:, None], 'logpj': lp}


class MoG(object):
    def __init__(self, D, H, sigmas_sq_type='full', to_learn=('pies', 'W', 'sigmas_sq')):
        self.D, self.H, self.sigmas_sq_type, self.to_learn = D, H, sigmas_sq_type, list(to_learn)

    def e_step(self, T, params, y):
        beta = 1. / T
        W = params['W'].T
        logp = np.zeros((y.shape[0], self.H))
        for h in range(self.H):
            yn = y - W[h]
            sig = params['sigmas_sq'][h]
            if self.sigmas_sq_type == 'full':
                quad = np.einsum('nd,de,ne->n', yn, np.linalg.inv(sig), yn)
                logdet = np.linalg.slogdet(sig)[1]
            else:
                quad = np.exp(2 * np.log(np.abs(yn) + TINY) - np.log(sig)[None, :]).sum(1)      # :257-258
                logdet = np.sum(np.log(sig))
            logp[:, h] = -(logdet + quad) * beta
        return _posterior(logp, params['pies'], beta, self.H)

    def m_step(self, params, post, y):
        params = dict(params)
        sum_post = post.sum(0) + TINY
        if 'W' in self.to_learn:
            params['W'] = (y.T @ post) * np.power(sum_post, -1)[None, :]
        if 'sigmas_sq' in self.to_learn:
            if self.sigmas_sq_type == 'full':
                sig = np.einsum('nh,nd,ne->hde', post, y, y) * np.power(sum_post, -1)[:, None, None]
                for h in range(self.H):
                    sig[h] -= np.outer(params['W'][:, h], params['W'][:, h])
            else:
                sig = ((y ** 2).T @ post).T * np.power(sum_post, -1)[:, None] - params['W'].T ** 2
            params['sigmas_sq'] = sig
        if 'pies' in self.to_learn:
            params['pies'] = sum_post / np.sum(sum_post)
        return params


class MoP(object):
    def __init__(self, D, H, A=np.nan, to_learn=('pies', 'W')):
        if not np.isnan(A) and A <= D:
            A = 10 * D
        self.D, self.H, self.A, self.to_learn = D, H, A, list(to_learn)

    def normalize(self, y):
        return ((self.A - self.D) / (y.sum(1) + EPS)[:, None]) * y + 1

    def e_step(self, T, params, y):
        beta = 1. / T
        if not np.isnan(self.A):
            y = self.normalize(y)
        W = params['W'].astype(np.longdouble)
        logp = y.astype(np.longdouble) @ np.log(W)
        if np.isnan(self.A):
            logp = logp - W.sum(0)[None, :]
        return _posterior(np.asarray(logp * beta, dtype=np.float64), params['pies'], beta, self.H)

    def m_step(self, params, post, y):
        params = dict(params)
        if not np.isnan(self.A):
            y = self.normalize(y)
        W_num = y.T @ post
        if 'W' in self.to_learn:
            denom = post.sum(0) if np.isnan(self.A) else W_num.sum(0) / self.A + EPS
            params['W'] = W_num / denom[None, :] + EPS
        if 'pies' in self.to_learn:
            sp = post.sum(0) + TINY
            params['pies'] = sp / sp.sum()
        return params
