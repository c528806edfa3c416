"""NumPy oracle for Binary Sparse Coding ET (test infrastructure; see oracle/__init__.py).

Follows prosper/em/camodels/bsc_et.py:
  select_hprimes  <- :98-115     e_step <- :119-192     m_step <- :195-438
"""
import math

import numpy as np

from . import common, states


class SerialComm(object):
    """What the oracle needs from a communicator; one rank = identity."""
    rank = 0
    size = 1

    def allreduce(self, x):
        return x

    def allsort(self, a):
        return np.sort(a)


class BSC(object):
    name = 'bsc'

    def __init__(self, D, H, Hprime, gamma, to_learn=('W', 'pi', 'sigma'), comm=None):
        assert Hprime <= H and gamma <= Hprime        # camodels/__init__.py:90-91
        self.D, self.H, self.Hprime, self.gamma = D, H, Hprime, gamma
        self.to_learn = list(to_learn)
        self.comm = comm or SerialComm()
        self.state_matrix, self.state_abs = states.binary_states(Hprime, gamma)
        self.no_states = self.state_matrix.shape[0]
        self.log = {}

    # bsc_et.py:98-115 -------------------------------------------------------
    def select_hprimes(self, params, data):
        y = data['y']
        W = params['W'].T                                      # (H,D)
        wn = np.sqrt(np.einsum('hd,hd->h', W, W))              # sqrt(diag(inner(W,W))), :111
        yn = np.sqrt(np.einsum('nd,nd->n', y, y))
        sim = (y @ W.T) / wn[None, :] / yn[:, None]            # :111
        data['candidates'] = np.argsort(sim, axis=1)[:, -self.Hprime:].astype(np.int64)   # :112
        data['_sim'] = sim                                     # oracle extra: lets tests measure score gaps
        return data

    # bsc_et.py:119-192 ------------------------------------------------------
    def e_step(self, anneal, params, data):
        H = self.H
        W = params['W'].T
        pies, sigma = params['pi'], params['sigma']
        if 'mu' not in params:                                 # :145-149 (mutates the caller's dict)
            params['mu'] = np.zeros(self.D)
        mu = params['mu']
        beta = 1. / anneal['T']                                # :152
        pre1 = -1. / 2. / sigma / sigma                        # :153
        pil_bar = np.log(pies / (1. - pies))                   # :154
        y = data['y'] - mu                                     # :169
        cand = data['candidates']
        n = y.shape[0]
        F = np.empty((n, 1 + H + self.no_states))
        F[:, 0] = pre1 * np.einsum('nd,nd->n', y, y)           # :172-173
        F[:, 1:1 + H] = pre1 * common.single_sqerr(W, y)       # :176-177
        F[:, 1 + H:] = pre1 * common.state_sqerr(W, y, cand, self.state_matrix)   # :180-185
        pre_F = np.empty(1 + H + self.no_states)               # :162-164
        pre_F[0] = 0.
        pre_F[1:1 + H] = pil_bar
        pre_F[1 + H:] = pil_bar * self.state_abs
        if anneal['anneal_prior']:                             # :187-190
            F = beta * (pre_F[None, :] + F)
        else:
            F = pre_F[None, :] + beta * F
        return {'logpj': F}

    # bsc_et.py:195-438 ------------------------------------------------------
    def m_step(self, anneal, params, suff, data):
        comm = self.comm
        H, gamma = self.H, self.gamma
        W = params['W'].T
        pies, sigma, mu = params['pi'], params['sigma'], params['mu']
        y = data['y'].copy()
        cand = data['candidates']
        logpj = suff['logpj']
        with np.errstate(over='ignore', under='ignore'):
            all_denoms = np.exp(logpj).sum(axis=1)             # :222 (not max-shifted)
        my_N, D = y.shape
        N = comm.allreduce(my_N)                               # :225
        SM = self.state_matrix.astype(np.float64)

        A, B = common.binom_AB(H, gamma, pies)                 # :239-243
        E = pies * H * A / B                                   # :244

        if anneal['Ncut_factor'] > 0.0:                        # :247-258
            N_use = int(N * (1 - (1 - A) * anneal['Ncut_factor']))
            which = common.truncate(all_denoms, N_use, strict=False, allsort=comm.allsort)
            cand, logpj, y = cand[which], logpj[which], y[which]
            my_N = y.shape[0]
            N_use = comm.allreduce(my_N)
        else:
            N_use = N
        self.log['N'] = N_use                                  # :261

        L = H * np.log(1 - pies) - 0.5 * D * np.log(2 * math.pi * sigma ** 2) - np.log(A)   # :264
        with np.errstate(over='ignore', under='ignore', divide='ignore'):
            Fs = np.log(np.exp(logpj).sum(axis=1)).sum()       # :265
        L += comm.allreduce(Fs) / N_use                        # :266
        self.log['L'] = L

        corr = logpj.max(axis=1)                               # :271
        pjb = np.exp(logpj - corr[:, None])                    # :272
        post = pjb / pjb.sum(axis=1)[:, None]                  # denom :362
        yc = y - mu                                            # :335
        single = post[:, 1:1 + H]
        multi = post[:, 1 + H:]

        # <s> per datapoint: singles (:349,:352) + multi-state marginals scattered to cand (:355,:360)
        exp_s = single.copy()
        marg = multi @ SM                                      # (n,H')
        rows = np.arange(my_N)[:, None]
        np.add.at(exp_s, (rows, cand), marg)
        my_Wp = exp_s.T @ yc                                   # :349,:355,:363
        my_Wq = np.diag(single.sum(axis=0))                    # :350,:364
        blocks = np.einsum('ns,sj,sk->njk', multi, SM, SM)     # :357
        np.add.at(my_Wq, (cand[:, :, None], cand[:, None, :]), blocks)   # :356-358
        my_pi = single.sum() + (multi * self.state_abs[None, :]).sum()   # :351,:359,:365
        my_mus = exp_s.sum(axis=0)                             # :352,:360,:366
        data_sum = y.sum(axis=0)                               # :283

        if 'W' in self.to_learn:                               # :369-382
            Wp = comm.allreduce(my_Wp)
            Wq = comm.allreduce(my_Wq)
            W_new = np.linalg.lstsq(Wq, Wp, rcond=common.numpy_rcond())[0]
        else:
            W_new = W

        if 'pi' in self.to_learn:                              # :385-389
            pi_new = E * comm.allreduce(my_pi) / H / N_use
        else:
            pi_new = pies

        if 'sigma' in self.to_learn:                           # :392-419, with the OLD W
            sq = np.empty_like(post)
            sq[:, 0] = np.einsum('nd,nd->n', yc, yc)
            sq[:, 1:1 + H] = common.single_sqerr(W, yc)
            sq[:, 1 + H:] = common.state_sqerr(W, yc, cand, self.state_matrix)
            my_sigma = (post * sq).sum()
            sigma_new = np.sqrt(comm.allreduce(my_sigma) / D / N_use)
        else:
            sigma_new = sigma

        if 'mu' in self.to_learn:                              # :422-430 (divides by LOCAL my_N, quirk B5)
            mus = comm.allreduce(my_mus)
            all_data_sum = comm.allreduce(data_sum)
            mu_new = all_data_sum / my_N - np.inner(W_new.T / my_N, mus)
        else:
            mu_new = mu

        self.log['N_use'] = N_use                              # :436
        return {'W': W_new.T, 'pi': pi_new, 'sigma': sigma_new, 'mu': mu_new}

    # camodels/__init__.py:163-193 (noise / partial data are host-side and left to the caller)
    def step(self, anneal, params, data):
        data = self.select_hprimes(params, data)
        suff = self.e_step(anneal, params, data)
        return self.m_step(anneal, params, suff, data)
