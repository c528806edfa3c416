"""NumPy oracle for Gaussian Sparse Coding (spike-and-slab) ET (test infrastructure; see oracle/__init__.py).

Follows prosper/em/camodels/gsc_et.py:
  component_scores :752-809, select_hprimes :721-749, compute_posterior_hprime :260-398,
  e_step :401-580, m_step :584-718.
Direct restatement: every (datapoint, state) builds W_s, Lambda_s, C^-1 explicitly like the reference
(no Gram shortcut).  Reference quirks kept: the exponent has no factor 1/2 (:346,:511); the null
state is NOT clamped to `tiny` while all others are (:461-463 vs :355-357,:520-522); psi_sq uses an
ELEMENT-WISE product with inv(sum<ss> + eps I) (:673); the E-step returns the datapoints grouped by
candidate set ("cluster-major") and overwrites my_data['y'] / ['candidates'] (:572-573).
"""
import numpy as np

from . import states
from .bsc import SerialComm

TINY = np.finfo(np.float64).tiny


class GSC(object):
    name = 'gsc'

    def __init__(self, D, H, Hprime=0, gamma=0, sigma_sq_type='scalar',
                 to_learn=('W', 'pi', 'mu', 'sigma_sq', 'psi_sq'), comm=None):
        # CAModel.__init__ builds the state matrix BEFORE the <=0 fix-ups (gsc_et.py:33,45-50)
        assert Hprime <= H and gamma <= Hprime
        self.D, self.H, self.Hprime, self.gamma = D, H, Hprime, gamma
        self.state_matrix, self.state_abs = states.binary_states(Hprime, gamma)
        self.no_states = self.state_matrix.shape[0]
        if gamma <= 0 or gamma > H:
            self.gamma = H
        if Hprime <= 0 or Hprime > H:
            self.Hprime = H
        elif Hprime < gamma:
            self.gamma = self.Hprime
        self.sigma_sq_type = sigma_sq_type
        self.to_learn = list(to_learn)
        self.comm = comm or SerialComm()

    def _B(self, params):
        s2 = params['sigma_sq']
        if self.sigma_sq_type == 'full':
            return np.linalg.inv(s2)
        if self.sigma_sq_type == 'diagonal':
            return np.diag(1. / s2)
        return (1. / s2) * np.eye(self.D)

    def _log_weight(self, params, B, y, comps):
        """-(logdet Psi_s + logdet Lambda_s) - (y - W_s mu_s)^T C^-1 (y - W_s mu_s) for rows y (m,D),
        plus kappa (m,k) and Lambda^-1 (k,k).  comps: indices into H (k,).
        Assumes a SYMMETRIC sigma_sq for the 'full' type (what the M-step produces, :685-691).  The
        reference's own standard_init builds a non-symmetric 'full' matrix by a broadcasting slip (:85)
        and then multiplies sigma_sq_inv from different sides in the singleton and multi-cause code
        (:481 vs :315); that corner is not reproduced."""
        W_s = params['W'][:, comps]                                   # (D,k)
        mu_s = params['mu'][comps]
        psi_s = params['psi_sq'][np.ix_(comps, comps)]
        BW = B @ W_s                                                  # "sigma_sq_inv_W_s"
        lam = W_s.T @ (B @ W_s) + np.linalg.inv(psi_s)                # :319 / :486
        lam_inv = np.linalg.inv(lam)
        lam_inv_W = lam_inv @ (W_s.T @ B)                             # (k,D)  :323 / :490
        yn = y - (W_s @ mu_s)[None, :]
        kappa = yn @ lam_inv_W.T + mu_s[None, :]                      # :329-331
        C_inv = B - BW @ lam_inv_W                                    # :337
        C_det = np.linalg.slogdet(psi_s)[1] + np.linalg.slogdet(lam)[1]
        post = -C_det - np.einsum('nd,de,ne->n', yn, C_inv, yn)       # :339-342
        return post, kappa, lam_inv

    # gsc_et.py:752-809 -----------------------------------------------------------------------
    def component_scores(self, params, y):
        B = self._B(params)
        log_tiny = np.finfo(np.float64).min
        out = np.zeros((y.shape[0], self.H))
        for h in range(self.H):
            post, _, _ = self._log_weight(params, B, y, np.array([h]))
            post[np.isnan(post)] = log_tiny
            post[post < log_tiny] = log_tiny
            post[np.isinf(post)] = 0
            out[:, h] = post
        return out

    # gsc_et.py:721-749 -----------------------------------------------------------------------
    def select_hprimes(self, params, data):
        scores = self.component_scores(params, data['y'])
        cand = np.argsort(scores, axis=1)[:, -self.Hprime:]
        data['candidates'] = np.sort(cand, axis=1).astype(np.int64)
        data['_sim'] = scores
        return data

    # gsc_et.py:811-944 -----------------------------------------------------------------------
    def compute_lpj(self, params, data):
        """Un-annealed, un-clamped log-joints [null | H singletons | S multi-cause states] per datapoint, rows in
        the order of data['y'] (the reference computes them cluster-major and sorts back, :941-944)."""
        data = self.select_hprimes(params, data)
        y, cand = data['y'], data['candidates']
        B = self._B(params)
        H, S = self.H, self.no_states
        log_pi_pr = np.log(params['pi']) - np.log(1 - np.asarray(params['pi']))          # :846
        out = np.zeros((y.shape[0], 1 + H + S))
        out[:, 0] = -np.einsum('nd,de,ne->n', y, B, y)                                   # :864
        for h in range(H):                                                               # :868-893
            post, _, _ = self._log_weight(params, B, y, np.array([h]))
            out[:, 1 + h] = post + log_pi_pr[h]
        order = self.cluster_order(cand)
        sc = cand[order]
        starts = np.flatnonzero(np.r_[True, np.any(sc[1:] != sc[:-1], axis=1)])
        ends = np.r_[starts[1:], len(order)]
        for a, b in zip(starts, ends):                                                   # :897-926
            rows = order[a:b]
            comps_all = sc[a]
            for s_i in range(S):
                act = np.nonzero(self.state_matrix[s_i] > 0)[0]
                comps = comps_all[act]
                post, _, _ = self._log_weight(params, B, y[rows], comps)
                out[rows, 1 + H + s_i] = post + log_pi_pr[comps].sum()
        return out, cand

    @staticmethod
    def cluster_order(cand):
        """Row permutation that groups equal candidate sets, clusters in first-appearance order,
        original order inside a cluster (dict insertion order of :733-745)."""
        _, first, inv = np.unique(cand, axis=0, return_index=True, return_inverse=True)
        rank_of_cluster = np.argsort(np.argsort(first))
        return np.argsort(rank_of_cluster[inv.ravel()], kind='stable')

    # gsc_et.py:401-580 + :260-398 ------------------------------------------------------------
    def e_step(self, anneal, params, data):
        H = self.H
        beta = 1. / anneal['T']
        B = self._B(params)
        perm = self.cluster_order(data['candidates'])
        y = data['y'][perm]
        cand = data['candidates'][perm]
        n = y.shape[0]
        log_pi_pr = np.log(params['pi']) - np.log(1 - np.array(params['pi']))
        nfac = np.zeros(n)
        s = np.zeros((n, H)); ss = np.zeros((n, H, H)); sz = np.zeros((n, H)); szsz = np.zeros((n, H, H))
        with np.errstate(under='ignore'):
            nfac += np.exp(-np.einsum('nd,de,ne->n', y, B, y) * beta)            # null state :459-463
            for h in range(H):                                                   # singletons :466-541
                post, kappa, lam_inv = self._log_weight(params, B, y, np.array([h]))
                w = np.exp((post + log_pi_pr[h]) * beta)
                w[np.isnan(w)] = TINY
                w[w < TINY] = TINY
                nfac += w
                s[:, h] += w
                ss[:, h, h] += w
                sz[:, h] += kappa[:, 0] * w
                szsz[:, h, h] += (kappa[:, 0] ** 2 + lam_inv[0, 0]) * w
            # multi-cause states, datapoint groups with the same candidate set        :546-560, :304-391
            start = 0
            while start < n:
                stop = start
                while stop < n and np.array_equal(cand[stop], cand[start]):
                    stop += 1
                comps_all = cand[start]
                for row in self.state_matrix:
                    comps = comps_all[row > 0]
                    post, kappa, lam_inv = self._log_weight(params, B, y[start:stop], comps)
                    w = np.exp((post + log_pi_pr[comps].sum()) * beta)
                    w[np.isnan(w)] = TINY
                    w[w < TINY] = TINY
                    nfac[start:stop] += w
                    s[start:stop, comps] += w[:, None]
                    ix = np.ix_(np.arange(start, stop), comps, comps)
                    ss[ix] += w[:, None, None]
                    sz[start:stop, comps] += kappa * w[:, None]
                    szsz[ix] += (kappa[:, :, None] * kappa[:, None, :] + lam_inv[None]) * w[:, None, None]
                start = stop
        nf = 1.0 / (nfac + TINY)                                                 # :564
        data['y'] = y                                                            # :572-573
        data['candidates'] = cand.astype(np.float64)
        data['_perm'] = perm
        return {'xpt_s': s * nf[:, None], 'xpt_ss': ss * nf[:, None, None],
                'xpt_sz': sz * nf[:, None], 'xpt_szsz': szsz * nf[:, None, None]}

    # gsc_et.py:584-718 -----------------------------------------------------------------------
    def m_step(self, anneal, params, suff, data):
        comm = self.comm
        xs, xss, xsz, xszsz = suff['xpt_s'], suff['xpt_ss'], suff['xpt_sz'], suff['xpt_szsz']
        y = data['y']
        my_N, D = y.shape
        H = self.H
        N = comm.allreduce(my_N)
        eps = 1e-5
        sum_s = comm.allreduce(xs.sum(axis=0))
        sum_sz = comm.allreduce(xsz.sum(axis=0))
        sum_szsz = comm.allreduce(xszsz.sum(axis=0))
        Wp = comm.allreduce(y.T @ xsz)                                           # :613-620
        W_n = Wp @ np.linalg.inv(sum_szsz)                                       # :624-626
        if 'pi' in self.to_learn:                                                # :640-645
            pi_new = sum_s / N
            pi_new[pi_new <= 5e-5] = 5e-5
            pi_new[pi_new >= 1 - 5e-5] = 1 - 5e-5
            params['pi'] = pi_new
        if 'W' in self.to_learn:
            params['W'] = W_n
        if 'mu' in self.to_learn:                                                # :653-654
            params['mu'] = sum_sz * 1. / (sum_s + np.finfo(np.float64).eps)
        if 'psi_sq' in self.to_learn:                                            # :657-673
            mu = params['mu']
            my_psi = np.outer(mu, mu) * xss.sum(axis=0) + xszsz.sum(axis=0) - 2 * (mu[None, :] * xs).T @ xsz
            psi = comm.allreduce(my_psi)
            sum_ss = comm.allreduce(xss.sum(axis=0))
            params['psi_sq'] = psi * np.linalg.inv(sum_ss + eps * np.eye(H)) + eps * np.eye(H)
        if 'sigma_sq' in self.to_learn:                                          # :675-715
            M_out = xsz.T @ xsz
            if self.sigma_sq_type == 'full':
                my = (y.T @ y - W_n @ M_out @ W_n.T) / N
                params['sigma_sq'] = comm.allreduce(my) + eps * np.eye(D)
            elif self.sigma_sq_type == 'diagonal':
                my = (y ** 2).sum(axis=0) - ((xsz @ W_n.T) ** 2).sum(axis=0)
                params['sigma_sq'] = comm.allreduce(my) / N + eps
            else:
                my = (y * y).sum() - np.trace(M_out @ (W_n.T @ W_n))
                params['sigma_sq'] = comm.allreduce(my) / N / D + eps
        return params

    def step(self, anneal, params, data):
        data = self.select_hprimes(params, data)
        return self.m_step(anneal, params, self.e_step(anneal, params, data), data)
