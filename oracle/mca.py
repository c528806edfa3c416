"""NumPy oracles for Maximal Causes Analysis ET and Max-Magnitude Causes ET
(test infrastructure; see oracle/__init__.py).

MCA  follows prosper/em/camodels/mca_et.py:  check_params :44-55, select_hprimes :88-111,
     e_step :114-179, m_step :182-377.
MMCA follows prosper/em/camodels/mmca_et.py: check_params :48-60, select_hprimes :95-124,
     e_step :127-202, m_step :205-426.
Both superimpose causes with a rho-norm (soft max), so there is no Gram shortcut: W-bar is a
D-vector per state.  NB the E-step does NOT apply the annealing beta (the M-step does,
mca_et.py:237-238), truncation ranks ANNEALED log-denominators (:250-254) and Q uses the
un-annealed ones (:327,:371-375).
"""
import math

import numpy as np
from scipy.special import comb

from . import common, states
from .bsc import SerialComm


class _MaxCauses(object):
    def __init__(self, D, H, Hprime, gamma, to_learn=('W', 'pi', 'sigma'), comm=None):
        assert Hprime <= H and gamma <= Hprime
        self.D, self.H, self.Hprime, self.gamma = D, H, Hprime, gamma
        self.to_learn = list(to_learn)
        self.comm = comm or SerialComm()
        self.state_matrix, self.state_abs = states.binary_states(Hprime, gamma)
        self.no_states = self.state_matrix.shape[0]
        self.log = {}

    # -- model specific -----------------------------------------------------------------------
    def rho(self, T):
        raise NotImplementedError

    def wbar(self, W, cand, rho, chunk):
        """(Wlbar, Wbar) of every (n, state): each (m, S, D)."""
        raise NotImplementedError

    # -- shared E-step: mca_et.py:140-179 / mmca_et.py:159-202 ---------------------------------
    def e_step(self, anneal, params, data):
        H = self.H
        W = params['W'].T
        pies, sigma = params['pi'], params['sigma']
        y, cand = data['y'], data['candidates']
        rho = self.rho(anneal['T'])
        pre1 = -1. / 2. / sigma / sigma
        pil_bar = np.log(pies / (1. - pies))
        n = y.shape[0]
        F = np.empty((n, 1 + H + self.no_states))
        F[:, 0] = pre1 * np.einsum('nd,nd->n', y, y)
        F[:, 1:1 + H] = pil_bar + pre1 * common.single_sqerr(W, y)
        step = max(1, (128 << 20) // max(1, self.no_states * self.D * 8))
        for a in range(0, n, step):
            b = min(n, a + step)
            _, Wbar = self.wbar(W, cand[a:b], rho)
            d = Wbar - y[a:b, None, :]
            F[a:b, 1 + H:] = pil_bar * self.state_abs[None, :] + pre1 * np.einsum('msd,msd->ms', d, d)
        assert np.isfinite(F).all()                      # mca_et.py:177
        return {'logpj': F}

    def _AB(self, pies):
        A = 0.
        B = 0.
        for gp in range(self.gamma + 1):
            a = comb(self.H, gp) * pies ** gp * (1. - pies) ** (self.H - gp)
            A += a
            B += gp * a
        return A, B

    def step(self, anneal, params, data):
        params = self.check_params(params)
        data = self.select_hprimes(params, data)
        return self.m_step(anneal, params, self.e_step(anneal, params, data), data)

    # -- shared M-step skeleton -----------------------------------------------------------------
    def m_step(self, anneal, params, suff, data):
        comm = self.comm
        H, D = self.H, self.D
        W = params['W'].T
        pies, sigma = params['pi'], params['sigma']
        y, cand, logpj = data['y'], data['candidates'], suff['logpj']
        my_N = y.shape[0]
        N = comm.allreduce(my_N)
        SM = self.state_matrix.astype(np.float64)
        T = anneal['T']
        rho = self.rho(T)
        beta = 1. / T
        corr = beta * logpj.max(axis=1)                                  # :237
        logpjb = beta * logpj - corr[:, None]
        pjb = np.exp(logpjb)                                             # :238
        A, B = self._AB(pies)
        if anneal['Ncut_factor'] > 0.0:                                  # :249-264
            denoms = np.log(pjb.sum(axis=1)) + corr
            N_use = int(N * (1 - (1 - A) * anneal['Ncut_factor']))
            sel = common.truncate(denoms, N_use, strict=False, allsort=comm.allsort)
            N_use = comm.allreduce(int(sel.sum()))
        else:
            sel = np.ones(my_N, dtype=bool)
            N_use = N
        self.log['N_use'] = N_use
        y, cand, logpj, logpjb, pjb = y[sel], cand[sel], logpj[sel], logpjb[sel], pjb[sel]
        n = y.shape[0]
        inv = 1.0 / pjb.sum(axis=1)
        p0, ps, pm = pjb[:, 0] * inv, pjb[:, 1:1 + H] * inv[:, None], pjb[:, 1 + H:] * inv[:, None]

        my_Wp, my_Wq = self.single_stats(W, y, ps)
        my_pi = ps.sum() + (pm * self.state_abs[None, :]).sum()
        my_sigma = (p0 * np.einsum('nd,nd->n', y, y)).sum() + (ps * common.single_sqerr(W, y)).sum()
        Wl = np.log(np.abs(W))
        step = max(1, (64 << 20) // max(1, self.no_states * self.Hprime * D * 8))
        for a in range(0, n, step):
            b = min(n, a + step)
            Wlbar, Wbar = self.wbar(W, cand[a:b], rho)                   # (m,S,D)
            expo = self.aid_exponent(logpjb[a:b, 1 + H:], Wlbar, Wl[cand[a:b]], rho)   # (m,S,H',D)
            with np.errstate(under='ignore'):
                Aid = (SM[None, :, :, None] * np.exp(expo)).sum(axis=1) * inv[a:b, None, None]   # (m,H',D)
            np.add.at(my_Wp, cand[a:b], Aid * y[a:b, None, :])
            np.add.at(my_Wq, cand[a:b], Aid)
            d = Wbar - y[a:b, None, :]
            my_sigma += (pm[a:b] * np.einsum('msd,msd->ms', d, d)).sum()
        with np.errstate(under='ignore', divide='ignore'):
            my_ldenom = np.log(np.exp(logpj).sum(axis=1)).sum()          # :327, un-annealed

        if 'W' in self.to_learn:
            W_new = self.update_W(W, comm.allreduce(my_Wp), comm.allreduce(my_Wq))
        else:
            W_new = W.T
        pi_new = A / B * pies * comm.allreduce(my_pi) / N_use if 'pi' in self.to_learn else pies
        sigma_new = np.sqrt(comm.allreduce(my_sigma) / D / N_use) if 'sigma' in self.to_learn else sigma
        lAi = (H * np.log(1. - pi_new)) - ((D / 2) * np.log(2 * math.pi)) - (D * np.log(sigma_new))   # :372
        Q = lAi * N_use + comm.allreduce(my_ldenom)
        return {'W': W_new, 'pi': pi_new, 'sigma': sigma_new, 'Q': Q}


class MCA(_MaxCauses):
    name = 'mca'
    rho_temp_bound = 1.05
    W_tol = 1e-4

    def check_params(self, params):                                      # mca_et.py:44-55
        params['W'] = np.maximum(params['W'], self.W_tol)
        return params

    def rho(self, T):                                                    # mca_et.py:143-145
        return 1. / (1. - 1. / np.maximum(T, self.rho_temp_bound))

    def select_hprimes(self, params, data):                              # mca_et.py:88-111
        y = data['y']
        W = params['W'].T
        sim = np.empty((y.shape[0], self.H))
        for a in range(0, y.shape[0], 4096):
            Wi = np.maximum(W[None, :, :], y[a:a + 4096, None, :])
            sim[a:a + 4096] = np.abs(Wi - y[a:a + 4096, None, :]).sum(axis=2)
        data['candidates'] = np.argsort(sim, axis=1)[:, :self.Hprime].astype(np.int64)
        data['_sim'] = -sim
        return data

    def wbar(self, W, cand, rho):                                        # mca_et.py:149-150,171-173
        Wrho = np.exp(rho * np.log(W))
        t = np.matmul(self.state_matrix.astype(np.float64)[None], Wrho[cand])
        Wlbar = np.log(t) / rho
        return Wlbar, np.exp(Wlbar)

    def single_stats(self, W, y, ps):                                    # mca_et.py:294-295
        W2 = W * W
        return W2 * (ps.T @ y), W2 * ps.sum(axis=0)[:, None]

    def aid_exponent(self, blpj, Wlbar, Wl_c, rho):                      # mca_et.py:304-309
        return blpj[:, :, None, None] + (1 - rho) * Wlbar[:, :, None, :] + (rho - 1) * Wl_c[:, None, :, :]

    def update_W(self, W, Wp, Wq):                                       # mca_et.py:343-348
        tiny = np.finfo(Wq.dtype).tiny
        Wp = Wp.copy()
        Wq = Wq.copy()
        Wp[Wq < tiny] = 0.
        Wq[Wq < tiny] = tiny
        return (Wp / Wq).T


class MMCA(_MaxCauses):
    name = 'mmca'
    rho_T_bound = 1.20
    rho_lbound = 1
    rho_ubound = 35
    tol = 1e-4

    def check_params(self, params):                                      # mmca_et.py:48-60 (in place)
        W = params['W']
        W[np.logical_and(W >= 0., W < +self.tol)] = +self.tol
        W[np.logical_and(W <= 0., W > -self.tol)] = -self.tol
        return params

    def rho(self, T):                                                    # mmca_et.py:163-165
        rho = 1. / (1. - 1. / np.maximum(T, self.rho_T_bound))
        return np.maximum(np.minimum(rho, self.rho_ubound), self.rho_lbound)

    def select_hprimes(self, params, data):                              # mmca_et.py:95-124
        sim = common.single_sqerr(params['W'].T, data['y'])
        data['candidates'] = np.argsort(sim, axis=1)[:, :self.Hprime].astype(np.int64)
        data['_sim'] = -sim
        return data

    def wbar(self, W, cand, rho):                                        # mmca_et.py:169-171,191-192
        Wrhos = np.sign(W) * np.exp(rho * np.log(np.abs(W)))
        t0 = np.matmul(self.state_matrix.astype(np.float64)[None], Wrhos[cand])
        with np.errstate(divide='ignore'):
            Wlbar = np.log(np.abs(t0)) / rho
        return Wlbar, np.sign(t0) * np.exp(Wlbar)

    def single_stats(self, W, y, ps):                                    # mmca_et.py:318-319
        return ps.T @ y, np.repeat(ps.sum(axis=0)[:, None], W.shape[1], axis=1)

    def aid_exponent(self, logpjb, Wlbar, Wl_c, rho):                    # mmca_et.py:338-340
        t = np.maximum(Wlbar[:, :, None, :] - Wl_c[:, None, :, :], 0.)
        return logpjb[:, :, None, None] - (rho - 1) * t

    def update_W(self, W, Wp, Wq):                                       # mmca_et.py:383-394
        Wq = np.maximum(Wq, self.tol)
        W_new = Wp / Wq
        inertia = np.maximum(1. - np.exp(-Wq / 2.5), 0.2)
        return (inertia * W_new + (1 - inertia) * W).T
