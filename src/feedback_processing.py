"""Design-matrix builder (reference: src/feedback_processing.py:9-187).  Kept on the host on purpose: it consumes the
global legacy numpy RNG in a fixed order (grid jitter) and produces a few KB per query; the device sees only the final X.

Row layout of X (and X_full): one block of m+1 rows per query -- the chosen point alpha* xi + x first, then the m
pseudo-observations alpha_k xi + x -- which is the comparison-set layout every CUDA kernel of this package assumes."""
import numpy as np
import scipy.stats

from misc import alpha_bounds
from TGN_distribution import TGN_sample


class FeedbackProcessing:
    def __init__(self, D, m, original_bounds, alpha_grid_distribution, TGN_speed):
        self.D, self.m = D, m
        self.original_bounds = original_bounds
        self.bounds = ((0, 1),) * D
        self.alpha_grid_distribution = alpha_grid_distribution
        self.TGN_speed = TGN_speed
        self.iter_number = 1
        self.X_obs = self.X_full = self.X = self.N = None
        self.obs_indices = self.pseudobs_indices = self.latest_obs_indices = None

    # ---- data flow (:34-44)
    def initialize_data(self, X_obs):
        self.X_obs = X_obs
        self.create_X()
        self.create_indices_bookkeeping()

    def update_data(self, X_obs):
        self.iter_number += 1
        self.X_obs = X_obs
        self.update_X()
        self.create_indices_bookkeeping()

    # ---- alpha grids (:47-108)
    def _draw_alphas(self, dist, m, lo, hi, alpha_star):
        """m distinct alphas in [lo, hi]; redraw until no duplicates survive clipping (same retry loop as the reference,
        so the RNG stream stays aligned with it)."""
        width = np.abs(hi - lo)
        if dist == 'TGN':
            gamma = 3 / np.power(np.max([self.iter_number + 1 - self.D, 1]), self.TGN_speed) + 2
        while True:
            if dist == 'equispaced':
                edge = (hi - lo) * (0.01 / 2)
                a = np.linspace(lo + edge, hi - edge, num=m) + np.random.normal(0, width * 0.01, m)
            elif dist == 'Cauchy':
                a = scipy.stats.cauchy.rvs(loc=float(alpha_star), scale=width * 0.07, size=m)
            elif dist == 'TGN':
                a = TGN_sample(size=m, gamma=gamma, alpha=float(alpha_star), x_min=lo, x_max=hi)
            else:
                print('Uknown alpha-distribution: ' + str(dist))
                raise ValueError(dist)
            a = np.unique(np.clip(a, lo, hi))
            if len(a) == m:
                return a

    def xi_grid(self, xi, x=None, alpha_grid_distribution=None, alpha_star=None, m=None, is_scaled=False):
        dist = self.alpha_grid_distribution if alpha_grid_distribution is None else alpha_grid_distribution
        m = self.m if m is None else m
        if is_scaled:
            lo, hi = 0, 1
        else:
            lo, hi = alpha_bounds(xi, [b[0] for b in self.original_bounds], [b[1] for b in self.original_bounds])
        alpha = self._draw_alphas(dist, m, lo, hi, alpha_star).reshape(m, 1)
        line = alpha @ np.array(xi, dtype=float).reshape(1, self.D)
        if x is None:
            return line[:, ~(line == 0).all(axis=0)]          # drop the coordinates the projection does not move
        return line + np.tile(x, (m, 1))

    # ---- design matrix (:110-154)
    def _query_block(self, row):
        D, m = self.D, self.m
        point, xi, alpha_star = row[:D], row[D:2 * D], row[-1]
        fixed = xi == 0
        x = np.zeros(D)
        x[fixed] = point[fixed]
        pts = np.vstack([point, self.xi_grid(xi=xi, x=x, alpha_star=alpha_star)])
        flags = np.r_[0.0, np.ones(m)].reshape(m + 1, 1)
        return np.hstack([pts, np.tile(xi, (m + 1, 1)), flags])           # [point | xi | is_pseudo-observation]

    def _finish(self, X_full):
        self.X_full = X_full
        self.X = self.scale(X_full[:, :self.D])
        self.N = len(self.X)

    def create_X(self):
        blocks = [self._query_block(self.X_obs[i]) for i in range(self.X_obs.shape[0])]
        self._finish(np.concatenate([np.empty((0, 2 * self.D + 1))] + blocks, axis=0))

    def update_X(self):
        """X_obs gained one row: append its block, keep the pseudo-observation grids drawn earlier."""
        self._finish(np.concatenate([self.X_full, self._query_block(self.X_obs[-1])], axis=0))

    def is_pseudobs(self, i):
        return bool(self.X_full[i, 2 * self.D])

    def create_indices_bookkeeping(self):
        flags = self.X_full[:, 2 * self.D].astype(bool)
        idx = np.arange(self.N)
        self.obs_indices = [int(i) for i in idx[~flags]]
        self.pseudobs_indices = [int(i) for i in idx[flags]]
        last = np.maximum.accumulate(np.where(flags, -1, idx))             # latest true observation at or before row i
        self.latest_obs_indices = [int(i) for i in last]

    # ---- [0,1] scaling (:167-186)
    def _span(self):
        lo = np.array([b[0] for b in self.original_bounds])
        hi = np.array([b[1] for b in self.original_bounds])
        return lo, np.abs(hi - lo)

    def scale(self, X, retain_0_values=False):
        lo, width = self._span()
        out = (X - lo) / width
        if retain_0_values:
            out[np.asarray(X) == 0] = 0
        return out

    def unscale(self, X, retain_0_values=False):
        lo, width = self._span()
        out = X * width + lo
        if retain_0_values:
            out[np.asarray(X) == 0] = 0
        return out
