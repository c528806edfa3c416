"""State-space enumeration of the truncated posterior (oracle; test infrastructure).

Restates:
  * binary   -- prosper/em/camodels/__init__.py:21-47   (generate_state_matrix)
  * ternary  -- prosper/em/camodels/tsc_et.py:23-80     (generate_state_matrix)
  * discrete -- prosper/em/camodels/dsc_et.py:56-63,161-191 (get_states + ctor)
The column order of `logpj` depends on the row order produced here, so the
orders follow the reference exactly (itertools.combinations by size then
lexicographic; itertools.product order for ternary/discrete).
"""
import itertools

import numpy as np


def binary_states(Hprime, gamma):
    """All H'-vectors with 2..gamma ones -> (state_matrix uint8 (S,H'), state_abs (S,))."""
    rows = []
    for g in range(2, gamma + 1):
        for combo in itertools.combinations(range(Hprime), g):
            r = np.zeros(Hprime, dtype=np.uint8)
            r[list(combo)] = 1
            rows.append(r)
    sm = np.array(rows, dtype=np.uint8).reshape(len(rows), Hprime)
    return sm, sm.sum(axis=1)


def product_states(values, Hprime):
    """All len(values)**H' tuples in itertools.product order, as (n,H') array."""
    values = np.asarray(values)
    K = len(values)
    idx = np.indices((K,) * Hprime).reshape(Hprime, -1).T
    return values[idx]


def ternary_states(Hprime, gamma, H, values=(-1., 0., 1.)):
    """TSC: returns (single_state_matrix (2H,H) int8, state_matrix (S_t,H') int8,
    no_states = 3**H' (unfiltered count, reference quirk), states_abs (3, 3**H'))."""
    values = np.asarray(values, dtype=np.float64)
    blocks = [np.eye(H, dtype=np.int8) * int(v) for v in values if v != 0]
    ssm = np.concatenate(blocks)
    s = product_states(values, Hprime).astype(np.int8)
    states_abs = np.stack([(s == v).sum(axis=1) for v in values]).astype(np.float64)
    sm = s[np.abs(s).sum(axis=1) <= gamma]
    return ssm, sm, s.shape[0], states_abs


def discrete_states(values, Hprime, gamma, H):
    """DSC: returns (single_state_matrix ((K-1)H,H), state_matrix (S,H') float,
    state_abs (K,S) with the zero-row counting H - nnz, K_0)."""
    values = np.asarray(values, dtype=np.float64)
    K = len(values)
    k0 = int(np.argwhere(values == 0.)[0, 0])
    ssm = np.concatenate([np.eye(H) * values[i] for i in range(K) if i != k0])
    s = product_states(values, Hprime)
    nnz = (s != 0).sum(axis=1)
    sm = s[(nnz <= gamma) & (nnz > 1)]
    state_abs = np.stack([(sm == values[i]).sum(axis=1) for i in range(K)]).astype(np.float64)
    state_abs[k0] = H - state_abs.sum(axis=0) + state_abs[k0]
    return ssm, sm, state_abs, k0
