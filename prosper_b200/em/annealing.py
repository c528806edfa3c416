"""Annealing schedules read by the hot path through `anneal['T']`, `anneal['Ncut_factor']`, ...

Same interface and interpolation rule as prosper/em/annealing.py:49-139 (LinearAnnealing):
piecewise-linear (position, value) points, float positions are fractions of `steps`,
negative positions count from the end, a missing key reads as 0.0 (:93-94).
"""
import numpy as np


class Annealing(object):
    """Base class: subclasses provide reset(), next(gain), __getitem__ and as_dict()."""

    def reset(self):
        raise NotImplementedError

    def next(self, gain=0.0):
        raise NotImplementedError


class LinearAnnealing(Annealing):
    def __init__(self, steps=80):
        self.steps = steps
        self.anneal_params = {}
        self.crit_params = []
        self.reset()
        self['max_step'] = [(steps, steps)]
        self['position'] = [(0, 0.), (steps, 1.)]
        self['step'] = [(0, 0.), (steps, steps)]

    def add_param(self, param_name, points):
        if np.isscalar(points):
            points = [(0, points)]
        stored = []
        for point in points:
            if not isinstance(point, tuple):
                raise TypeError("points must be a list of (pos, val)-tuples")
            pos, val = point
            if isinstance(pos, float):
                pos = int(pos * self.steps)
            if pos < 0:
                pos = self.steps + pos
            stored.append((pos, val))
        if stored[0][0] != 0:                      # hold the first value from step 0
            stored.insert(0, (0, stored[0][1]))
        if stored[-1][0] != self.steps:            # hold the last value to the end
            stored.append((self.steps + 1, stored[-1][1]))
        self.anneal_params[param_name] = stored

    def __setitem__(self, param_name, points):
        self.add_param(param_name, points)

    def __getitem__(self, param_name):
        if param_name not in self.anneal_params:
            return 0.0
        points = self.anneal_params[param_name]
        i = 0
        for i in range(len(points)):
            if points[i][0] > self.cur_pos:
                break
        (lp, lv), (rp, rv) = points[i - 1], points[i]
        frac = float(self.cur_pos - lp) / (rp - lp)
        return frac * (rv - lv) + lv

    def reset(self):
        self.cur_pos = 0
        self.finished = False

    def next(self, gain=0.0):
        if self.finished:
            raise RuntimeError("Should not next() further when already finished!")
        self.accept = True
        self.cur_pos += 1
        if self.cur_pos >= self.steps:
            self.finished = True

    def as_dict(self):
        return dict((name, self[name]) for name in self.anneal_params)
