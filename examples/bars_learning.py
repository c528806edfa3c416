#!/usr/bin/env python
"""Bars test written against the REFERENCE's import names, run on this engine (BASELINE.json configs[0]).

    python examples/bars_learning.py [bsc|mca|mmca|tsc|dsc] [N]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/bars_learning.py bsc

`prosper_b200.install_as_prosper()` registers the package as `prosper` (and a stand-in for `mpi4py.MPI.COMM_WORLD`), so
everything below the first three lines is what a prosper user writes -- the same calls as the reference's
examples/barstests/bars-learning.py + param-bars-*.py.  Acceptance: the learned W is a permutation of the
generating bars; the script prints the mean absolute error after `find_permutation` and writes output/.../result.h5.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import prosper_b200  # noqa: E402

prosper_b200.install_as_prosper()

import numpy as np  # noqa: E402
from mpi4py import MPI  # noqa: E402

from prosper.em import EM  # noqa: E402
from prosper.em.annealing import LinearAnnealing  # noqa: E402
from prosper.utils import create_output_path  # noqa: E402
from prosper.utils.barstest import generate_bars_dict, find_permutation  # noqa: E402
from prosper.utils.datalog import dlog, StoreToH5, TextPrinter, StoreToTxt  # noqa: E402
from prosper.utils.parallel import pprint  # noqa: E402


def build(kind, size, H, Hprime, gamma):
    W_gt = 10 * generate_bars_dict(H)
    if kind == 'bsc':
        from prosper.em.camodels.bsc_et import BSC_ET
        return BSC_ET(size ** 2, H, Hprime, gamma), {'W': W_gt, 'pi': 2. / H, 'sigma': 2.0}
    if kind == 'mca':
        from prosper.em.camodels.mca_et import MCA_ET
        return MCA_ET(size ** 2, H, Hprime, gamma), {'W': W_gt, 'pi': 2. / H, 'sigma': 2.0}
    if kind == 'mmca':
        from prosper.em.camodels.mmca_et import MMCA_ET
        return MMCA_ET(size ** 2, H, Hprime, gamma), {'W': W_gt, 'pi': 2. / H, 'sigma': 2.0}
    if kind == 'tsc':
        from prosper.em.camodels.tsc_et import TSC_ET
        return TSC_ET(size ** 2, H, Hprime, gamma), {'W': W_gt, 'pi': 2. / H, 'sigma': 2.0}
    if kind == 'dsc':
        from prosper.em.camodels.dsc_et import DSC_ET
        states = np.array([-1., 0., 1.])
        return (DSC_ET(size ** 2, H, Hprime, gamma, states=states),
                {'W': W_gt, 'pi': np.array([1. / H, 1 - 2. / H, 1. / H]), 'sigma': 2.0})
    raise SystemExit("unknown model %r" % kind)


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else 'bsc'
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    size, H, Hprime, gamma = 5, 10, 6, 3
    comm = MPI.COMM_WORLD
    np.random.seed(1 + comm.rank)
    model, params_gt = build(kind, size, H, Hprime, gamma)
    output_path = create_output_path("bars-%s" % kind)
    pprint("Bars test %s: N=%d on %d process(es), results in %s" % (kind.upper(), N, comm.size, output_path))

    my_data = model.generate_data(params_gt, N // comm.size)
    print_list = ('T', 'L', 'pi', 'sigma', 'N_use')
    dlog.set_handler(print_list, TextPrinter)
    dlog.set_handler(print_list, StoreToTxt, output_path + 'terminal.txt')
    dlog.set_handler(('*'), StoreToH5, output_path + 'result.h5')

    model_params = model.standard_init(my_data)
    anneal = LinearAnnealing(50)
    anneal['T'] = [(0, 4. if kind in ('mca', 'mmca') else 2.), (.7, 1.)]
    anneal['Ncut_factor'] = [(0, 0.), (2. / 3, 1.)]
    anneal['anneal_prior'] = False

    em = EM(model=model, anneal=anneal)
    em.data = my_data
    em.lparams = model_params
    em.run()
    dlog.close()

    W = np.asarray(em.lparams['W'])
    perm = find_permutation(np.abs(W) if kind in ('tsc', 'dsc') else W, params_gt['W'])
    mae = np.abs(np.abs(W[:, perm]) - params_gt['W']).mean() if kind in ('tsc', 'dsc') else np.abs(W[:, perm] - params_gt['W']).mean()
    pprint("MAE between the learned W (permuted) and the generating bars: %.3f  (bar amplitude 10)" % mae)
    pprint("Done")
