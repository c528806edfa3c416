"""EM driver and Model base class with the reference's interface.

`EM(model, anneal, data, lparams).run()` and `Model.noisify_params/gain` follow
prosper/em/__init__.py:24-178.  They are O(1) host logic per iteration; the O(N) work is
inside `model.step`, i.e. the CUDA hot path.
"""
import numpy as np

from ..utils import parallel
from ..utils.datalog import dlog


class Model(object):
    """Base class (prosper/em/__init__.py:24-110)."""

    def __init__(self, comm=None):
        self.comm = comm if comm is not None else parallel.default_comm()
        self.noise_policy = {}

    def generate_data(self, model_params, N):
        raise NotImplementedError

    def step(self, anneal, model_params, my_data):
        raise NotImplementedError

    def standard_init(self, data):
        raise NotImplementedError

    def noisify_params(self, model_params, anneal):
        """Add annealed parameter noise: rank 0 draws, everyone receives (em/__init__.py:63-107).

        The scalar clamp follows the reference literally: it assigns to an unused name
        (`new_value`, em/__init__.py:85-88), i.e. scalar bounds are NOT applied.
        """
        comm = self.comm
        for param, policy in self.noise_policy.items():
            pvalue = model_params[param]
            scale = anneal[param + "_noise"]
            if scale != 0.0 and hasattr(pvalue, 'is_cuda') and pvalue.is_cuda and pvalue.dim() == 2 and pvalue.dtype.is_floating_point:
                model_params[param] = self._noisify_device(pvalue, scale, policy)
                continue
            on_device = None
            if scale != 0.0 and hasattr(pvalue, 'detach') and hasattr(pvalue, 'cpu'):   # other device-resident parameters
                on_device = pvalue.device
                pvalue = pvalue.detach().cpu().numpy()
            if scale != 0.0:
                if np.isscalar(pvalue):
                    new_pvalue = 0
                    if comm.rank == 0:
                        new_pvalue = pvalue + np.random.normal(scale=scale)
                        if policy[2]:
                            new_pvalue = np.abs(new_pvalue)
                    pvalue = comm.bcast(new_pvalue)
                else:
                    new_pvalue = pvalue
                    if comm.rank == 0:
                        low, up, absify = policy
                        new_pvalue = pvalue + np.random.normal(scale=scale, size=pvalue.shape)
                        new_pvalue = np.minimum(up, np.maximum(low, new_pvalue))
                        if absify:
                            new_pvalue = np.abs(new_pvalue)
                    pvalue = comm.bcast(new_pvalue)
            if on_device is not None:
                import torch
                pvalue = torch.as_tensor(pvalue).to(on_device)
            model_params[param] = pvalue
        return model_params

    def _noisify_device(self, pvalue, scale, policy):
        """Parameter noise for a matrix that lives on the device (em/__init__.py:82-103): rank 0 draws a SEED, every rank
        runs the same counter-based generator (`pet_normal_fill`, Philox keyed by seed and element index), so all
        ranks add identical noise and nothing but the seed is broadcast; clamp and abs follow on the device."""
        import ctypes as C
        import torch
        from .. import _lib
        seed = self.comm.bcast(int(np.random.randint(0, 2 ** 31 - 1)) if self.comm.rank == 0 else None)
        lib = _lib.load()
        src = pvalue if pvalue.dtype == torch.float64 else pvalue.to(torch.float64)
        noise = torch.empty(src.shape, dtype=torch.float64, device=src.device)
        st = C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream)
        _lib.check(lib.pet_normal_fill(C.c_void_p(noise.data_ptr()), noise.stride(0), noise.shape[0], noise.shape[1], None,
                                       float(scale), seed, st))
        low, up, absify = policy
        out = src + noise
        if np.isfinite(low) or np.isfinite(up):
            out = torch.clamp(out, min=(low if np.isfinite(low) else None), max=(up if np.isfinite(up) else None))
        if absify:
            out = torch.abs(out)
        return out.to(pvalue.dtype)

    def gain(self, old_params, new_params):
        return 0.


class EM(object):
    """Drives the annealed EM loop (prosper/em/__init__.py:115-178)."""

    def __init__(self, model=None, anneal=None, data=None, lparams=None, mpi_comm=None):
        self.model = model
        self.anneal = anneal
        self.data = data
        self.lparams = lparams
        self.mpi_comm = mpi_comm

    def step(self):
        self.model.step(self.anneal, self.lparams, self.data)

    def run(self, verbose=False):
        model, anneal, my_data = self.model, self.anneal, self.data
        model_params = self.lparams
        while not anneal.finished:
            if verbose:
                dlog.progress("EM step %d of %d" % (anneal['step'] + 1, anneal['max_step']), anneal['position'])
            new_model_params = model.step(anneal, model_params, my_data)
            gain = model.gain(model_params, new_model_params)
            anneal.next(gain)
            if anneal.accept:
                model_params = new_model_params
            self.lparams = model_params
