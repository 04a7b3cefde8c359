"""A small Bayesian optimiser for the autotune service (reference: bagua/service/bayesian_optimizer.py:1-79, a thin
wrapper over scikit-optimize).  scikit-optimize is not a dependency here: this is a self-contained
Gaussian-process / expected-improvement optimiser over integer, float and boolean dimensions with a Halton sequence for
the initial design (the reference also starts from 20 Halton points)."""
from __future__ import annotations

import math
from typing import Optional, Dict, List, Tuple, Union

import numpy as np

__all__ = ["IntParam", "FloatParam", "BoolParam", "BayesianOptimizer"]


class IntParam:
    """Integer search dimension ``[lo, hi]`` with its current value."""
    def __init__(self, val: int, space_dimension: Tuple[int, int]):
        self.val = int(val)
        self.space_dimension = (int(space_dimension[0]), int(space_dimension[1]))


class FloatParam:
    """Real-valued search dimension ``[lo, hi]`` with its current value."""
    def __init__(self, val: float, space_dimension: Tuple[float, float]):
        self.val = float(val)
        self.space_dimension = (float(space_dimension[0]), float(space_dimension[1]))


class BoolParam:
    """Boolean search dimension with its current value."""
    def __init__(self, val: bool):
        self.val = bool(val)
        self.space_dimension = (0, 1)


Param = Union[IntParam, FloatParam, BoolParam]


def _halton(index: int, base: int) -> float:
    f, r = 1.0, 0.0
    while index > 0:
        f /= base
        r += f * (index % base)
        index //= base
    return r


_PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29]


class BayesianOptimizer:
    """Maximises a black-box score.  ``tell(params, score)`` feeds an observation, ``ask()`` proposes the next point."""

    def __init__(self, param_declaration: Dict[str, Param], n_initial_points: int = 20, initial_point_generator: str = "halton",
                 random_state: Optional[int] = 0, seed: Optional[int] = None):
        """Signature of the reference (bayesian_optimizer.py:39-57, a thin wrapper of ``skopt.Optimizer``); ``seed`` is an alias of
        ``random_state``.  ``initial_point_generator``: ``"halton"`` (default) or ``"random"``."""
        if initial_point_generator not in ("halton", "random"):
            raise ValueError("initial_point_generator must be 'halton' or 'random'")
        self.initial_point_generator = initial_point_generator
        seed = seed if seed is not None else (random_state if random_state is not None else 0)
        self.param_declaration = dict(sorted(param_declaration.items()))
        self.names = list(self.param_declaration.keys())
        self.n_initial_points = n_initial_points
        self._asked = 0
        self._X: List[List[float]] = []
        self._y: List[float] = []
        self._rng = np.random.RandomState(seed)

    # -- encoding: every dimension → [0, 1] -----------------------------------------------------------------------
    def _encode(self, params: Dict[str, Union[int, float, bool]]) -> List[float]:
        x = []
        for n in self.names:
            lo, hi = self.param_declaration[n].space_dimension
            x.append((float(params[n]) - lo) / (hi - lo) if hi > lo else 0.0)
        return x

    def _decode(self, x) -> Dict[str, Union[int, float, bool]]:
        out = {}
        for n, u in zip(self.names, x):
            p = self.param_declaration[n]
            lo, hi = p.space_dimension
            v = lo + float(np.clip(u, 0.0, 1.0)) * (hi - lo)
            if isinstance(p, BoolParam):
                out[n] = bool(round(v))
            elif isinstance(p, IntParam):
                out[n] = int(round(v))
            else:
                out[n] = float(v)
        return out

    def tell(self, param_dict: Dict[str, Union[int, float, bool]], score: float) -> None:
        if score is None or not math.isfinite(score):
            return
        self._X.append(self._encode(param_dict))
        self._y.append(float(score))

    def _initial_point(self) -> List[float]:
        if self.initial_point_generator == "random":
            return self._rng.rand(len(self.names)).tolist()
        i = self._asked + 1
        return [_halton(i, _PRIMES[d % len(_PRIMES)]) for d in range(len(self.names))]

    def ask(self) -> Dict[str, Union[int, float, bool]]:
        self._asked += 1
        if self._asked <= self.n_initial_points or len(self._y) < 3:
            self._asked -= 1
            x = self._initial_point()
            self._asked += 1
            return self._decode(x)
        from scipy.stats import norm
        from sklearn.gaussian_process import GaussianProcessRegressor
        from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

        X = np.asarray(self._X)
        y = np.asarray(self._y)
        mu, sd = y.mean(), y.std() + 1e-12
        yn = (y - mu) / sd
        kernel = ConstantKernel(1.0, (1e-2, 1e2)) * Matern(length_scale=0.3, length_scale_bounds=(1e-2, 1e1), nu=2.5) + WhiteKernel(1e-3, (1e-6, 1e0))
        gp = GaussianProcessRegressor(kernel=kernel, normalize_y=False, n_restarts_optimizer=1, random_state=self._rng.randint(1 << 30))
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gp.fit(X, yn)
        cand = self._rng.rand(512, len(self.names))
        # snap candidates to the decodable lattice so integer/bool dims are evaluated where they will be sampled
        cand = np.asarray([self._encode(self._decode(c)) for c in cand])
        m, s = gp.predict(cand, return_std=True)
        best = yn.max()
        z = (m - best - 0.01) / np.maximum(s, 1e-9)
        ei = (m - best - 0.01) * norm.cdf(z) + s * norm.pdf(z)
        return self._decode(cand[int(np.argmax(ei))])
