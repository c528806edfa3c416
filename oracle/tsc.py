"""NumPy oracle for Ternary Sparse Coding ET (test infrastructure; see oracle/__init__.py).

Follows prosper/em/camodels/tsc_et.py:
  select_hprimes <- :142-212    e_step <- :277-356    m_step <- :359-542
Reference quirks kept on purpose: the preselection ranks the 2H signed singletons and maps
them to cause indices, so a datapoint's candidate list may hold the same h twice (:208-210);
fancy-index `+=` with such duplicates keeps the last write only (:475-478); `no_states` is the
UNFILTERED 3**H' (:77); L carries no H*log(1-pi) term (:447).
"""
import math

import numpy as np
from scipy.special import comb

from . import common, states
from .bsc import SerialComm


class TSC(object):
    name = 'tsc'

    def __init__(self, D, H, Hprime, gamma, to_learn=('W', 'pi', 'sigma'), comm=None):
        self.D, self.H, self.Hprime, self.gamma = D, H, Hprime, gamma
        self.to_learn = list(to_learn)
        self.comm = comm or SerialComm()
        self.states = np.array([-1., 0., 1.])
        self.single_state_matrix, self.state_matrix, self.no_states, self.state_abs = \
            states.ternary_states(Hprime, gamma, H, self.states)
        self.log = {}

    def _log_prior(self, SM, pi):
        """sum_j log p(s_j), p(+-1)=pi/2, p(0)=1-pi over the entries of each row (:184-194,:316-326)."""
        nz = np.abs(SM).sum(axis=1)
        return nz * np.log(pi / 2) + (SM.shape[1] - nz) * np.log(1 - pi)

    # tsc_et.py:142-212 ----------------------------------------------------------
    def select_hprimes(self, params, data):
        y = data['y']
        W = params['W'].T
        pi, sigma = params['pi'], params['sigma']
        SSM = self.single_state_matrix.astype(np.float64)          # (2H,H): -1 block then +1 block
        pre1 = -1. / 2. / sigma / sigma
        pil_bar = self._log_prior(SSM, pi)                         # :184-194 (same for every row)
        Wbar = SSM @ W                                             # :203
        d = Wbar[None, :, :] - y[:, None, :]
        F = pil_bar[None, :] + pre1 * np.einsum('nkd,nkd->nk', d, d)   # :204-207
        tmp = np.argsort(F, axis=1)[:, -self.Hprime:]              # :208
        data['candidates'] = (tmp % self.H).astype(np.int64)       # np.nonzero(SM[tmp])[1], :209
        data['_sim'] = F
        return data

    # tsc_et.py:277-356 ----------------------------------------------------------
    def e_step(self, anneal, params, data):
        W = params['W'].T
        pi, sigma = params['pi'], params['sigma']
        SM = self.state_matrix
        beta = 1. / anneal['T']
        pre1 = -1. / 2. / sigma / sigma
        pil_bar = self._log_prior(SM, pi)                          # :316-326
        F = pre1 * common.state_sqerr(W, data['y'], data['candidates'], SM)   # :337-352
        if anneal['anneal_prior']:                                 # :354-359
            F += pil_bar[None, :]
            F *= beta
        else:
            F *= beta
            F += pil_bar[None, :]
        return {'logpj': F}

    # tsc_et.py:359-542 ----------------------------------------------------------
    def m_step(self, anneal, params, suff, data):
        comm = self.comm
        H, gamma = self.H, self.gamma
        W = params['W'].T
        pi, sigma = params['pi'], params['sigma']
        y = data['y'].copy()
        cand = data['candidates']
        logpj = suff['logpj']
        with np.errstate(over='ignore', under='ignore'):
            all_denoms = np.exp(logpj).sum(axis=1)                 # :412
        my_N, D = y.shape
        N = comm.allreduce(my_N)
        SM = self.state_matrix.astype(np.float64)
        state_abs = np.abs(SM).sum(axis=1)                         # :418

        A = 0.0                                                    # :422-431
        B = 0.0
        for g1 in range(gamma + 1):
            for g2 in range(gamma - g1 + 1):
                cmb = comb(g1, g1) * comb(g1 + g2, g2) * comb(H, H - g1 - g2)
                a = cmb * ((pi / 2) ** (g1 + g2)) * ((1 - pi) ** (H - g1 - g2))
                A += a
                B += (g1 + g2) * a
        E = pi * H * A / B

        if anneal['Ncut_factor'] > 0.0:                            # :434-444
            N_use = int(N * (1 - (1 - A) * anneal['Ncut_factor']))
            which = common.truncate(all_denoms, N_use, strict=False, allsort=comm.allsort)
            cand, logpj, y = cand[which], logpj[which], y[which]
            my_N = y.shape[0]
            N_use = comm.allreduce(my_N)
        else:
            N_use = N

        L = -0.5 * D * np.log(2 * math.pi * sigma ** 2) - np.log(A)    # :447
        with np.errstate(over='ignore', under='ignore', divide='ignore'):
            Fs = np.log(np.exp(logpj).sum(axis=1)).sum()
        L += comm.allreduce(Fs) / N_use
        self.log['L'] = L

        corr = logpj.max(axis=1)
        pjb = np.exp(logpj - corr[:, None])
        post = pjb / pjb.sum(axis=1)[:, None]

        # numpy fancy-index semantics for duplicate candidates: the LAST position holding an h wins
        Hp = self.Hprime
        live = np.ones((my_N, Hp), dtype=bool)
        for j in range(Hp):
            for j2 in range(j + 1, Hp):
                live[:, j] &= cand[:, j2] != cand[:, j]
        marg = post @ SM                                           # (n,H')  :475
        blocks = np.einsum('ns,sj,sk->njk', post, SM, SM)          # :477
        exp_s = np.zeros((my_N, H))
        rows = np.arange(my_N)[:, None]
        np.add.at(exp_s, (rows, cand), marg * live)
        my_Wp = exp_s.T @ y
        my_Wq = np.zeros((H, H))
        np.add.at(my_Wq, (cand[:, :, None], cand[:, None, :]), blocks * (live[:, :, None] & live[:, None, :]))
        my_pi = (post * state_abs[None, :]).sum()                  # :479

        if 'W' in self.to_learn:                                   # :487-495
            Wp = comm.allreduce(my_Wp)
            Wq = comm.allreduce(my_Wq)
            W_new = np.linalg.pinv(Wq) @ Wp
        else:
            W_new = W
        pi_new = E * comm.allreduce(my_pi) / H / N_use if 'pi' in self.to_learn else pi   # :499-503
        if 'sigma' in self.to_learn:                               # :505-534, OLD W
            sq = common.state_sqerr(W, y, cand, self.state_matrix)
            sigma_new = np.sqrt(comm.allreduce((post * sq).sum()) / D / N_use)
        else:
            sigma_new = sigma
        self.log['N_use'] = N_use
        return {'W': W_new.T, 'pi': pi_new, 'sigma': sigma_new, 'Q': 0.}

    def step(self, anneal, params, data):
        data = self.select_hprimes(params, data)
        return self.m_step(anneal, params, self.e_step(anneal, params, data), data)
