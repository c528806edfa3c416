"""NumPy oracle for Discrete Sparse Coding ET (test infrastructure; see oracle/__init__.py).

Follows prosper/em/camodels/dsc_et.py:
  select_hprimes <- :347-410   e_step <- :492-585   m_step <- :587-774
  scaling factor <- :798-823   truncation <- :825-843 (strict '>')   likelihood <- :845-870
"""
import itertools
import math

import numpy as np
from scipy.special import gammaln, logsumexp

from . import common, states
from .bsc import SerialComm


class DSC(object):
    name = 'dsc'

    def __init__(self, D, H, Hprime, gamma, states_=np.array([-1., 0., 1.]), to_learn=('W', 'pi', 'sigma'), comm=None):
        if not isinstance(states_, np.ndarray):
            raise TypeError("DSC: states must be of type numpy.ndarray")       # :139
        self.D, self.H, self.Hprime, self.gamma = D, H, Hprime, gamma
        self.to_learn = list(to_learn)
        self.comm = comm or SerialComm()
        self.states = states_
        self.K = len(states_)
        self.single_state_matrix, self.state_matrix, self.state_abs, self._K_0 = \
            states.discrete_states(states_, Hprime, gamma, H)
        self.no_states = self.state_matrix.shape[0]
        self.log = {}

    def _nonzero_values(self):
        return [i for i in range(self.K) if i != self._K_0]

    # dsc_et.py:347-410 ----------------------------------------------------------
    def select_hprimes(self, params, data):
        y = data['y']
        H = self.H
        W = params['W'].T
        pi, sigma = params['pi'], params['sigma']
        pre1 = -1. / 2. / sigma / sigma
        l_pis = np.concatenate([np.full(H, np.log(pi[i]) + (H - 1) * np.log(pi[self._K_0]))
                                for i in self._nonzero_values()])                 # :388-394
        Wbar = self.single_state_matrix @ W                                         # :401
        d = Wbar[None, :, :] - y[:, None, :]
        F = pre1 * np.einsum('nkd,nkd->nk', d, d) + l_pis[None, :]                  # :402-404
        n = y.shape[0]
        cand = np.zeros((n, self.Hprime), dtype=np.int64)
        for i in range(n):                                                          # :405-408
            order = np.mod(np.argsort(F[i]), H)[::-1]
            Fu, Si = np.unique(order, return_index=True)
            cand[i] = Fu[np.argsort(Si)][:self.Hprime]
        data['candidates'] = cand
        data['_sim'] = F.reshape(n, self.K - 1, H).max(axis=1)
        return data

    # dsc_et.py:492-585 ----------------------------------------------------------
    def e_step(self, anneal, params, data):
        H, K = self.H, self.K
        W = params['W'].T
        pi, sigma = params['pi'], params['sigma']
        y, cand = data['y'], data['candidates']
        beta = 1. / anneal['T']
        pre1 = -1. / 2. / sigma / sigma
        n = y.shape[0]
        ncol = 1 + (K - 1) * H + self.no_states
        l_pis = (self.state_abs * np.log(pi)[:, None]).sum(axis=0)                  # :525-527
        pre_F = np.empty(ncol)
        pre_F[0] = H * np.log(pi[self._K_0])                                        # :539
        for c, k in enumerate(self._nonzero_values()):                              # :540-545
            pre_F[c * H + 1:(c + 1) * H + 1] = np.log(pi[k]) + (H - 1) * np.log(pi[self._K_0])
        pre_F[(K - 1) * H + 1:] = l_pis
        F = np.empty((n, ncol))
        F[:, 0] = pre1 * np.einsum('nd,nd->n', y, y)                                # :552-553
        Wb = self.single_state_matrix @ W                                           # :558
        d = Wb[None, :, :] - y[:, None, :]
        F[:, 1:(K - 1) * H + 1] = pre1 * np.einsum('nkd,nkd->nk', d, d)
        if self.gamma > 1:                                                          # :561-567
            F[:, (K - 1) * H + 1:] = pre1 * common.state_sqerr(W, y, cand, self.state_matrix)
        if anneal['anneal_prior']:                                                  # :569-574
            F += pre_F[None, :]
            F *= beta
        else:
            F *= beta
            F += pre_F[None, :]
        return {'logpj': F}

    # dsc_et.py:798-823 ----------------------------------------------------------
    def scaling_factor(self, pi):
        A = 0.0
        for gp in itertools.product(range(self.gamma + 1), repeat=self.K - 1):
            ngp = np.array(gp)
            if ngp.sum() > self.gamma:
                continue
            abs_array = np.insert(ngp, self._K_0, self.H - ngp.sum())
            cmb = np.exp(gammaln(abs_array.sum() + 1) - gammaln(abs_array + 1).sum())   # multinom2 :23-39
            A += cmb * np.prod(pi ** abs_array)
        return A

    # dsc_et.py:587-774 ----------------------------------------------------------
    def m_step(self, anneal, params, suff, data):
        comm = self.comm
        H, K = self.H, self.K
        W = params['W'].T
        pi, sigma = params['pi'], params['sigma']
        y = data['y'].copy()
        cand = data['candidates']
        logpj = suff['logpj']
        with np.errstate(over='ignore', under='ignore'):
            all_denoms = np.exp(logpj).sum(axis=1)                                  # :638
        my_N, D = y.shape
        N = comm.allreduce(my_N)
        A = self.scaling_factor(pi)                                                 # :642
        self.log['prior_mass'] = A
        if anneal['Ncut_factor'] > 0.0:                                             # :825-843, strict '>'
            N_use = int(N * (1 - (1 - A) * anneal['Ncut_factor']))
            which = common.truncate(all_denoms, N_use, strict=True, allsort=comm.allsort)
            cand, logpj, y = cand[which], logpj[which], y[which]
            my_N = y.shape[0]
            N_use = comm.allreduce(my_N)
        else:
            N_use = N
        corr = logpj.max(axis=1)
        pjb = np.exp(logpj - corr[:, None])
        post = pjb / pjb.sum(axis=1)[:, None]
        L = -0.5 * D * np.log(2 * math.pi * sigma ** 2)                             # :866-869
        L += comm.allreduce(logsumexp(logpj, 1).sum()) / N_use
        self.log['L'] = L

        SM = self.state_matrix
        SSM = self.single_state_matrix
        ns = (K - 1) * H
        p0 = post[:, 0]
        ps = post[:, 1:ns + 1]
        pm = post[:, ns + 1:]
        my_pi = np.zeros(K)
        my_pi[self._K_0] += H * p0.sum() + (H - 1) * ps.sum()                       # :676,:697
        for c, k in enumerate(self._nonzero_values()):                              # :690
            my_pi[k] += ps[:, c * H:(c + 1) * H].sum()
        exp_s = ps @ SSM                                                            # (n,H)  :698
        my_Wq = SSM.T @ (ps.sum(axis=0)[:, None] * SSM)                             # :701
        sq0 = np.einsum('nd,nd->n', y, y)
        my_sigma = (p0 * sq0).sum()                                                 # :677
        Wb = SSM @ W
        dd = Wb[None, :, :] - y[:, None, :]
        my_sigma += (ps * np.einsum('nkd,nkd->nk', dd, dd)).sum()                   # :692-694
        if self.gamma > 1:                                                          # :705-718
            rows = np.arange(my_N)[:, None]
            np.add.at(exp_s, (rows, cand), pm @ SM)
            blocks = np.einsum('ns,sj,sk->njk', pm, SM, SM)
            np.add.at(my_Wq, (cand[:, :, None], cand[:, None, :]), blocks)
            my_pi += (pm[:, None, :] * self.state_abs[None, :, :]).sum(axis=(0, 2))
            my_sigma += (pm * common.state_sqerr(W, y, cand, SM)).sum()
        my_Wp = exp_s.T @ y
        my_sigma /= D                                                               # :724

        Wp = comm.allreduce(my_Wp)
        Wq = comm.allreduce(my_Wq)
        W_new = np.linalg.lstsq(Wq, Wp, rcond=common.numpy_rcond())[0]              # :732-735
        tot = comm.allreduce(my_pi)
        pi_new = tot / tot.sum()                                                    # :740-741
        eps = 1e-6                                                                  # :743-748
        if np.any(pi_new < eps):
            lo = pi_new < eps
            hi = ~lo
            pi_new[lo] += eps - pi_new[lo]
            pi_new[hi] -= (eps * lo.sum()) / hi.sum()
        sigma_new = np.sqrt(comm.allreduce(my_sigma) / N_use)                       # :759
        if 'W' not in self.to_learn:
            W_new = W
        if 'pi' not in self.to_learn:
            pi_new = pi
        if 'sigma' not in self.to_learn:
            sigma_new = sigma
        self.log['N_use'] = N_use
        return {'W': W_new.T, 'pi': pi_new, 'sigma': sigma_new, 'Q': 0.}

    def step(self, anneal, params, data):
        data = self.select_hprimes(params, data)
        return self.m_step(anneal, params, self.e_step(anneal, params, data), data)
