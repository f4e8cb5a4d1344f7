"""A stand-in for jwas_b200.GpuSweeper with the same method surface, backed by the CPU oracle (test infrastructure).
It lets the CPU suite execute the branch of api.runMCMC that builds GpuSweeper / mcmc.GpuBackend objects -- option and
call order, argument shapes, the second handle for EBV output rows -- which otherwise only runs on a B200."""
import numpy as np

from oracle_backend import OracleBackend


class _Stats:
    def __init__(self, d, t):
        z16, z4 = np.zeros(16), np.zeros(4)

        def pad(a, n):
            out = np.zeros(n); a = np.asarray(a, dtype=np.float64).reshape(-1); out[:len(a)] = a
            return out
        self.ycorr_ss = pad(d["ycorr_ss"], 16); self.alpha_ss = pad(d["alpha_ss"], 16); self.beta_ss = pad(d["beta_ss"], 16)
        self.ycorr_sum = pad(d["ycorr_sum"], 4); self.nnz_alpha = pad(d["nnz_alpha"], 4); self.sum_delta = pad(d["sum_delta"], 4)
        self.class_counts = pad(d["class_counts"], 16); self.bayesr_ssq = d["bayesr_ssq"]
        self.n_active = d["n_active"]; self.n_rounds = d["n_rounds"]


class FakeSweeper:
    """Same constructor and methods as _lib.GpuSweeper (the ones api.py / mcmc.GpuBackend use)."""
    log = []                                    # (method, args) of every call, for order checks

    def __init__(self, packed, n_obs, n_traits=1, device=0, rows=None):
        self.packed = np.ascontiguousarray(packed); self.n, self.t, self.p = n_obs, n_traits, packed.shape[0]
        self._means = None; self.b = None; self.options = {}
        FakeSweeper.log.append(("create", (self.p, n_obs, n_traits)))

    def set_marker_means(self, means):
        assert self.b is None, "jwas_set_marker_means must precede jwas_set_blocks"
        assert np.asarray(means).dtype == np.float32 and len(means) == self.p
        self._means = np.asarray(means, np.float32)
        FakeSweeper.log.append(("set_marker_means", None))

    def set_option(self, key, value):
        self.options[key] = int(value)

    def _backend(self):
        if self.b is None:                      # a handle that never gets set_blocks (the EBV rows) still serves M * alpha
            self.b = OracleBackend(self.packed, self.n, self.t, np.array([0, self.p], dtype=np.int64), means=self._means)
        return self.b

    def set_blocks(self, starts):
        self.b = OracleBackend(self.packed, self.n, self.t, starts, means=self._means)
        self.b.lag = self.options.get("lag", 0)
        FakeSweeper.log.append(("set_blocks", len(starts) - 1))

    def put_ycorr(self, y): self._backend().put_ycorr(y)
    def get_ycorr(self, out=None): return self._backend().get_ycorr()
    def put_state(self, alpha=None, beta=None, delta=None): self._backend().put_state(alpha, beta, delta)
    def get_state(self, out=None): return self._backend().get_state()
    def ycorr_sub_malpha(self): self._backend().sub_malpha()

    def shift_ycorr(self, trait, shift, want=True):
        b = self._backend()
        b.shift_ycorr(trait, shift)
        if want:
            return b.ycorr_sum(trait), 0.0

    def mul_alpha(self, trait=0): return self._backend().mul_alpha(trait)
    def fill_hyper(self, which, value): self._backend().fill_hyper(which, value)
    def sample_bayesb_variances(self, df, scale, seed, it, want=False): self._backend().sample_bayesb_variances(df, scale, seed, it)
    def accumulate(self, nsamples, bayesr=False): self._backend().accumulate(nsamples, bayesr)
    def get_means(self): return self._backend().get_means()
    def stream_kernel_ms(self): return (0.0, 0)
    last_sweep_ms = 0.0

    def sweep_bayesc(self, *a): return _Stats(self.b.sweep_bayesc(*a), 1)
    def sweep_bayesabc(self, schedule, vare, ve, pi, seed, it, u=None, z=None): return _Stats(self.b.sweep_bayesabc(schedule, vare, ve, pi, seed, it), 1)
    def sweep_bayesr(self, schedule, full, vare, s2, pi, gamma, seed, it, u=None, z=None): return _Stats(self.b.sweep_bayesr(schedule, full, vare, s2, pi, gamma, seed, it), 1)
    def sweep_mt1(self, schedule, R, G, big_pi, seed, it, u=None, z=None): return _Stats(self.b.sweep_mt1(schedule, R, G, big_pi, seed, it), self.t)
    def sweep_mt2(self, schedule, R, G, big_pi, seed, it, u=None, z=None): return _Stats(self.b.sweep_mt2(schedule, R, G, big_pi, seed, it), self.t)
    def sweep_mega(self, schedule, vare, ve, pi, seed, it, u=None, z=None): return _Stats(self.b.sweep_mega(schedule, vare, ve, pi, seed, it), self.t)
    def close(self): pass
