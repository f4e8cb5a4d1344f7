"""Sweep backend over the CPU oracle with the interface of jwas_b200.mcmc.GpuBackend, so that the
host-side chain logic (priors, schedules, hyper-parameter draws, outputs) can be tested without a
GPU and the GPU chain can be compared with an oracle chain end to end.  TEST INFRASTRUCTURE."""
import numpy as np

from helpers import canonical_sum_prod
from oracle import pyoracle as orc

SCHED_EXACT, SCHED_BLOCK, SCHED_INDEPENDENT = 0, 1, 2


class OracleBackend:
    name = "oracle"

    def __init__(self, packed, n, t, starts, means=None):
        self.packed = np.ascontiguousarray(packed); self.n, self.t = n, t
        self.p = packed.shape[0]
        self.starts = np.asarray(starts, dtype=np.int64)
        self.means, self.xpx = orc.marker_stats(self.packed, n)
        if means is not None:
            # centring on means computed on a larger sample (twin of jwas_set_marker_means): xpx from the integer
            # counts through the contract's closed form (gram_value), binary64, in that order
            self.means = np.asarray(means, dtype=np.float32).copy()
            codes = np.stack([(self.packed >> (2 * k)) & 3 for k in range(4)], axis=2).reshape(self.p, -1)[:, :n]
            n1 = (codes == 1).sum(axis=1).astype(np.float64); n2 = (codes == 2).sum(axis=1).astype(np.float64)
            nn = (codes != 3).sum(axis=1).astype(np.float64)
            ssum = n1 + 2 * n2; m64 = self.means.astype(np.float64)
            g = (n1 + 4 * n2) - m64 * ssum
            g = g - m64 * ssum
            g = g + (m64 * m64) * nn
            self.xpx = g.astype(np.float32)
        tp = t * self.p
        self.y = np.zeros(t * n, np.float32)
        self.alpha = np.zeros(tp, np.float32); self.beta = np.zeros(tp, np.float32); self.delta = np.zeros(tp, np.int32)
        self.ma = np.zeros(tp, np.float32); self.ma2 = np.zeros(tp, np.float32); self.md = np.zeros(tp, np.float32)
        self.ve = None; self.pi = None
        self.lag = 0

    def put_ycorr(self, y):
        self.y[:] = np.asarray(y, np.float32).reshape(-1)

    def get_ycorr(self):
        return self.y.copy()

    def shift_ycorr(self, trait, shift):
        sl = slice(trait * self.n, (trait + 1) * self.n)
        self.y[sl] = self.y[sl] + np.float32(shift)

    def ycorr_sum(self, trait):
        sl = slice(trait * self.n, (trait + 1) * self.n)
        return canonical_sum_prod(self.y[sl], np.ones(self.n))

    def put_state(self, alpha, beta, delta):
        if alpha is not None: self.alpha[:] = alpha
        if beta is not None: self.beta[:] = beta
        if delta is not None: self.delta[:] = delta

    def get_state(self):
        return self.alpha.copy(), self.beta.copy(), self.delta.copy()

    def sub_malpha(self):
        for k in range(self.t):
            a = self.alpha[k * self.p:(k + 1) * self.p]
            y = self.y[k * self.n:(k + 1) * self.n]
            for j in np.nonzero(a)[0]:
                x = orc.decode_marker(self.packed, self.n, int(j), float(self.means[j]))
                y[:] = (y.astype(np.float64) + (-np.float64(a[j])) * x.astype(np.float64)).astype(np.float32)

    def mul_alpha(self, trait):
        return orc.mul_alpha(self.packed, self.n, self.means, self.alpha[trait * self.p:(trait + 1) * self.p])

    def fill_hyper(self, which, value):
        if which == "var_effects":
            self.ve = np.full(self.p, float(value))
        else:
            self.pi = np.full(self.p, float(value))

    def sample_bayesb_variances(self, df, scale, seed, it):
        self.ve = orc.bayesb_variances(self.beta, df, scale, seed, it)

    def _stats(self, method):
        t, n, p = self.t, self.n, self.p
        Y = self.y.reshape(t, n); A = self.alpha.reshape(t, p); B = self.beta.reshape(t, p); D = self.delta.reshape(t, p)
        st = {"ycorr_ss": np.array([[canonical_sum_prod(Y[a], Y[b]) for b in range(t)] for a in range(t)]),
              "alpha_ss": np.array([[canonical_sum_prod(A[a], A[b]) for b in range(t)] for a in range(t)]),
              "beta_ss": np.array([[canonical_sum_prod(B[a], B[b]) for b in range(t)] for a in range(t)]),
              "ycorr_sum": np.array([canonical_sum_prod(Y[a], np.ones(n)) for a in range(t)]),
              "nnz_alpha": np.array([float(np.count_nonzero(A[a])) for a in range(t)]),
              "n_active": 0, "n_rounds": 0, "bayesr_ssq": 0.0}
        cc = np.zeros(16)
        if method == orc.METHOD_R:
            st["sum_delta"] = np.array([float((D[0] > 1).sum())])
            cc[:4] = np.bincount(D[0], minlength=5)[1:5]
            prod = np.where(D[0] > 1, A[0].astype(np.float64) ** 2 / np.array([1.0, 0.01, 0.1, 1.0])[np.clip(D[0] - 1, 0, 3)], 0.0)
            st["bayesr_ssq"] = canonical_sum_prod(prod, np.ones(p))
        else:
            st["sum_delta"] = np.array([float((D[a] != 0).sum()) for a in range(t)])
            state = sum((D[a] != 0).astype(int) << a for a in range(t))
            cc[:1 << t] = np.bincount(state, minlength=1 << t)
        st["class_counts"] = cc
        return st

    def _sweep(self, method, schedule, full_reps, seed, it, **kw):
        rc, _ = orc.sweep_contract(self.packed, self.n, self.means, self.xpx, self.starts, self.y, self.alpha,
                                   self.beta, self.delta, method=method,
                                   nreps_mode=0 if (schedule == SCHED_EXACT or not full_reps) else 1,
                                   independent=(schedule == SCHED_INDEPENDENT), seed=seed, it=it,
                                   lag=(self.lag if schedule == SCHED_EXACT else 0), **kw)
        assert rc == 0, "oracle fixed-point overflow"
        return self._stats(method)

    def sweep_bayesc(self, schedule, vare, var_effect, pi, seed, it):
        return self._sweep(orc.METHOD_ABC, schedule, 1, seed, it, vare=vare, varEffects=np.full(self.p, float(var_effect)),
                           pi=np.full(self.p, float(pi)))

    def sweep_bayesabc(self, schedule, vare, var_effects, pi, seed, it):
        ve = self.ve if var_effects is None else var_effects
        pv = self.pi if pi is None else pi
        return self._sweep(orc.METHOD_ABC, schedule, 1, seed, it, vare=vare, varEffects=ve, pi=pv)

    def sweep_bayesr(self, schedule, full_reps, vare, sigma_sq, pi, gamma, seed, it):
        return self._sweep(orc.METHOD_R, schedule, full_reps, seed, it, vare=vare, sigmaSq=sigma_sq, pi=pi, gamma=gamma)

    def sweep_mt1(self, schedule, R, G, big_pi, seed, it):
        return self._sweep(orc.METHOD_MT1, schedule, 1, seed, it, R=R, G=G, bigPi=big_pi)

    def sweep_mt2(self, schedule, R, G, big_pi, seed, it):
        return self._sweep(orc.METHOD_MT2, schedule, 1, seed, it, R=R, G=G, bigPi=big_pi)

    def sweep_mega(self, schedule, vare, var_effects, pi, seed, it):
        return self._sweep(orc.METHOD_MEGA, schedule, 1, seed, it, R=np.diag(vare), G=np.diag(var_effects), bigPi=np.asarray(pi, float))

    def accumulate(self, nsamples, bayesr=False):
        a = self.alpha.astype(np.float64)
        d = (self.delta > 1).astype(np.float64) if bayesr else self.delta.astype(np.float64)
        self.ma = (self.ma.astype(np.float64) + (a - self.ma) / nsamples).astype(np.float32)
        self.ma2 = (self.ma2.astype(np.float64) + (a * a - self.ma2) / nsamples).astype(np.float32)
        self.md = (self.md.astype(np.float64) + (d - self.md) / nsamples).astype(np.float32)

    def get_means(self):
        return self.ma.copy(), self.ma2.copy(), self.md.copy()


def factory(packed, n, t, starts, means=None):
    return OracleBackend(packed, n, t, starts, means=means)
