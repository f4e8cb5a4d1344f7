"""Chain driver for the marker-effects path: the part of MCMC_BayesianAlphabet
(MCMC/MCMC_BayesianAlphabet.jl:184-421) that surrounds the sweep -- intercept update,
pi / sigma^2_alpha / sigma^2_e draws and posterior accumulation -- written against an abstract
sweep backend so that tests can drive the identical host logic with the CPU oracle.

The product only ever instantiates GpuBackend (below); nothing here imports the oracle.
"""
import math

import numpy as np

from . import annotations as annot
from ._lib import GpuSweeper, SCHED_EXACT, SCHED_BLOCK, SCHED_INDEPENDENT

BAYESR_GAMMA = np.array([0.0, 0.01, 0.1, 1.0])      # JWAS.jl:12


class HostRng:
    """Host-side draws for the O(1) hyper-parameter updates (the reference uses Distributions.jl:
    Chisq, Beta, Dirichlet, InverseWishart -- variance_components.jl, Pi.jl)."""

    def __init__(self, seed):
        self.g = np.random.Generator(np.random.Philox(seed))

    def chisq(self, df):
        return float(self.g.chisquare(df))

    def beta(self, a, b):
        return float(self.g.beta(a, b))

    def dirichlet(self, a):
        return self.g.dirichlet(np.asarray(a, dtype=np.float64))

    def normal(self):
        return float(self.g.standard_normal())

    def gamma(self, shape, scale, size):
        return self.g.gamma(shape, scale, size)

    def uniform(self, size):
        return self.g.random(size)

    def inverse_wishart(self, df, scale):
        """InverseWishart(df, scale) via Bartlett; scale is the sum-of-squares matrix."""
        scale = np.asarray(scale, dtype=np.float64)
        t = scale.shape[0]
        L = np.linalg.cholesky(np.linalg.inv(scale))
        A = np.zeros((t, t))
        for i in range(t):
            A[i, i] = math.sqrt(self.g.chisquare(df - i))
            for j in range(i):
                A[i, j] = self.g.standard_normal()
        LA = L @ A
        return np.linalg.inv(LA @ LA.T)


class GpuBackend:
    """Sweep backend = the CUDA library.  Holds M, ycorr, alpha/beta/delta on the device."""

    name = "b200"

    def __init__(self, sweeper: GpuSweeper):
        self.s = sweeper
        # record = True (bench.py): per-sweep device times -- (streaming-kernel ms, launches) needs option "profile"
        self.record = False
        self.kernel_ms = []
        self.sweep_ms = []

    # ycorr lifecycle
    def put_ycorr(self, y):
        self.s.put_ycorr(y)

    def get_ycorr(self):
        return self.s.get_ycorr()

    def shift_ycorr(self, trait, shift):
        self.s.shift_ycorr(trait, shift, want=False)

    def ycorr_sum(self, trait):
        return self.s.shift_ycorr(trait, 0.0)[0]

    def put_state(self, alpha, beta, delta):
        self.s.put_state(alpha, beta, delta)

    def get_state(self):
        return self.s.get_state()

    def sub_malpha(self):
        self.s.ycorr_sub_malpha()

    # sweeps return a dict of the reductions the hyper-parameter draws need
    def _stats(self, st, t):
        if self.record:
            self.kernel_ms.append(self.s.stream_kernel_ms())
            self.sweep_ms.append(self.s.last_sweep_ms)
        return {
            "ycorr_ss": np.array(st.ycorr_ss[:t * t]).reshape(t, t),
            "alpha_ss": np.array(st.alpha_ss[:t * t]).reshape(t, t),
            "beta_ss": np.array(st.beta_ss[:t * t]).reshape(t, t),
            "ycorr_sum": np.array(st.ycorr_sum[:t]),
            "nnz_alpha": np.array(st.nnz_alpha[:t]), "sum_delta": np.array(st.sum_delta[:t]),
            "class_counts": np.array(st.class_counts[:16]), "bayesr_ssq": st.bayesr_ssq,
            "n_active": st.n_active, "n_rounds": st.n_rounds,
        }

    def sweep_bayesc(self, schedule, vare, var_effect, pi, seed, it):
        return self._stats(self.s.sweep_bayesc(schedule, vare, var_effect, pi, seed, it), 1)

    def sweep_bayesabc(self, schedule, vare, var_effects, pi, seed, it):
        return self._stats(self.s.sweep_bayesabc(schedule, vare, var_effects, pi, seed, it), 1)

    def sweep_bayesr(self, schedule, full_reps, vare, sigma_sq, pi, gamma, seed, it):
        return self._stats(self.s.sweep_bayesr(schedule, full_reps, vare, sigma_sq, pi, gamma, seed, it), 1)

    def sweep_mt1(self, schedule, R, G, big_pi, seed, it):
        return self._stats(self.s.sweep_mt1(schedule, R, G, big_pi, seed, it), self.s.t)

    def sweep_mt2(self, schedule, R, G, big_pi, seed, it):
        return self._stats(self.s.sweep_mt2(schedule, R, G, big_pi, seed, it), self.s.t)

    def sweep_mega(self, schedule, vare, var_effects, pi, seed, it):
        return self._stats(self.s.sweep_mega(schedule, vare, var_effects, pi, seed, it), self.s.t)

    def sample_bayesb_variances(self, df, scale, seed, it):
        self.s.sample_bayesb_variances(df, scale, seed, it)

    def fill_hyper(self, which, value):
        self.s.fill_hyper(which, value)

    def mul_alpha(self, trait):
        return self.s.mul_alpha(trait)

    def accumulate(self, nsamples, bayesr=False):
        self.s.accumulate(nsamples, bayesr)

    def get_means(self):
        return self.s.get_means()


def run_chain(backend, *, n, p, ntraits, method, schedule, chain_length, burnin, output_samples_frequency,
              seed, vare, var_effect, pi, df_effect, scale_effect, df_res, scale_res,
              estimate_pi=True, estimate_variance=True, estimate_vare=True, block_size=1,
              R=None, G=None, big_pi=None, scale_G=None, scale_R=None, sample_intercept=True,
              mu0=None, iter0=0, want_ebv=False, mt_sampler="I", constraint_G=False, constraint_R=False,
              sample_sink=None, annotations=None, ebv_backend=None, want_heritability=False, hyper_sink=None):
    """One MCMC run over an already-initialised backend (ycorr = y - mu0 - M*alpha on entry).

    Mirrors MCMC_BayesianAlphabet.jl:184-421 for `y = intercept + markers`:
      [1] intercept Gibbs (:207-220, solver.jl:143-151)  [2] marker sweep (:224-290)
      [3] pi (:294-317, Pi.jl)  [4] marker variance (:321-326, variance_components.jl:151-189)
      [5] residual variance (:355-371)  [6] posterior means every saved iteration (:399-413).
    With `annotations` (annotations.MarkerAnnotations) step [3] is update_marker_annotation_priors! (:296-305) and the
    sweep takes the marker-level priors it produces.
    Float32 re-casts of the variances follow :323-325, :368-370.
    """
    rng = HostRng([seed, iter0])
    t = ntraits
    mu = np.zeros(t) if mu0 is None else np.array(mu0, dtype=np.float64)
    out = {"mu_mean": np.zeros(t), "mu_mean2": np.zeros(t), "vare_mean": 0.0, "vare_mean2": 0.0,
           "vara_mean": 0.0, "vara_mean2": 0.0, "pi_mean": 0.0, "pi_mean2": 0.0, "nsamples": 0, "trace": [],
           "ebv_mean": None, "ebv_var": None}
    if method == "BayesR":
        pi = np.array(pi, dtype=np.float64)
        out["pi_mean"] = np.zeros_like(pi); out["pi_mean2"] = np.zeros_like(pi)
    if t > 1:
        R = np.array(R, dtype=np.float64); G = np.array(G, dtype=np.float64)
        big_pi = np.array(big_pi, dtype=np.float64)
        for key in ("vare_mean", "vare_mean2", "vara_mean", "vara_mean2"):
            out[key] = np.zeros((t, t))
        out["pi_mean"] = np.zeros_like(big_pi); out["pi_mean2"] = np.zeros_like(big_pi)
    ebv_m = ebv_s = None
    gvar_samples, h2_samples = [], []
    gamma_arr = None
    ann = annotations
    ann_prior = None                      # what the sweep takes: (p,) pi_j for BayesC, (p, 4) class / joint-state priors
    if ann is not None:
        if t == 1 and method == "BayesC":
            pi = np.array(pi, dtype=np.float64)
            if pi.shape != (p,):
                raise ValueError("BayesABC: pi vector length must match the number of markers")   # BayesABC.jl:17-23
            ann_prior = pi
            out["pi_mean"] = np.zeros(p); out["pi_mean2"] = np.zeros(p)
        elif (t == 1 and method == "BayesR") or (t == 2 and method == "BayesC" and not constraint_G and mt_sampler == "I"):
            ann_prior = ann.snp_pi
        else:
            raise ValueError("Unsupported annotation configuration.")
        estimate_pi = True                # normalize_annotation_estimatePi (readgenotypes.jl:152-158)
    if method == "BayesL":
        # Bayesian Lasso (BayesC0L.jl:25-47): marker j has variance var_effect * gamma_j; gamma ~ Gamma(1, 8) to
        # start with (MCMC_BayesianAlphabet.jl:72-77; api.runMCMC has already divided var_effect and its scale by 8)
        if t != 1:
            raise ValueError("BayesL: single-trait only in this backend")
        gamma_arr = rng.gamma(1.0, 8.0, p)
    first_bayesb = True
    nsamples = 0
    ysum = None
    for it in range(iter0 + 1, iter0 + chain_length + 1):
        # [1] intercept: ycorr += mu; mu ~ N(mean(ycorr), vare/n); ycorr -= mu
        if sample_intercept and t == 1:
            s = ysum[0] if ysum is not None else backend.ycorr_sum(0)
            new_mu = mu[0] + s / n + rng.normal() * math.sqrt(vare / n)      # Gibbs(A,x,b,vare), solver.jl:143-151
            backend.shift_ycorr(0, np.float32(mu[0] - new_mu))
            mu[0] = new_mu
        elif sample_intercept:
            # multi-trait: mmeLhs = X'RiX with Ri = kron(inv(R), I) (MCMC_BayesianAlphabet.jl:196-217) and one pass of
            # Gibbs(A,x,b) (solver.jl:154-162): each intercept from its FULL CONDITIONAL given the other traits'
            # current intercepts -- variance 1/(n Rinv_kk), mean through the off-diagonals of inv(R)
            Rinv = np.linalg.inv(R)
            ssum = np.array([ysum[k] if ysum is not None else backend.ycorr_sum(k) for k in range(t)], dtype=np.float64)
            for k in range(t):
                invlhs = 1.0 / (n * Rinv[k, k])
                new_mu = mu[k] + invlhs * float(Rinv[k] @ ssum) + rng.normal() * math.sqrt(invlhs)
                backend.shift_ycorr(k, np.float32(mu[k] - new_mu))
                ssum[k] += n * (mu[k] - new_mu)              # later traits see this trait's updated residual sum
                mu[k] = new_mu
        # [2] marker effects
        if method in ("BayesC", "BayesB", "BayesA") or (method == "RR-BLUP" and t > 1):
            # multi-trait RR-BLUP: MTBayesC0! (MTBayesC0L.jl:6-58) is sampler I with all the prior mass on the all-traits
            # state, megaBayesC0! (BayesC0L.jl:13-17) is megaBayesABC! with pi = 0 per trait -- api.runMCMC sets big_pi so
            if t == 1:
                if ann is not None:
                    # bayesabc_pi_vector (BayesABC.jl:17-23): one common variance, marker-level pi_j
                    st = backend.sweep_bayesabc(schedule, vare, np.full(p, float(var_effect)), ann_prior, seed, it)
                elif method == "BayesC":
                    st = backend.sweep_bayesc(schedule, vare, var_effect, pi, seed, it)
                else:
                    # BayesB/BayesA: G.val is a per-marker vector kept on the device
                    # (MCMC_BayesianAlphabet.jl:67-69); pi is refreshed when it is re-sampled
                    if first_bayesb:
                        backend.fill_hyper("var_effects", var_effect)
                        first_bayesb = False
                    backend.fill_hyper("pi", pi)
                    st = backend.sweep_bayesabc(schedule, vare, None, None, seed, it)
            elif constraint_G:
                # megaBayesABC! (MCMC_BayesianAlphabet.jl:233-234): per-trait pi vector, diagonal variances
                st = backend.sweep_mega(schedule, np.diag(R).copy(), np.diag(G).copy(), big_pi, seed, it)
            else:
                # annotated: MarkerSpecificPiPrior(snp_pi) (MTBayesABC.jl:28-30), columns 00, 10, 01, 11
                st = (backend.sweep_mt2 if mt_sampler == "II" else backend.sweep_mt1)(
                    schedule, R, G, big_pi if ann is None else ann_prior, seed, it)
        elif method == "BayesR":
            full = 1 if it > burnin else 0          # bayesr_block_nreps, BayesR.jl:22-25
            st = backend.sweep_bayesr(schedule, full, vare, var_effect, pi if ann is None else ann_prior,   # BayesR.jl:28
                                      BAYESR_GAMMA, seed, it)
        elif method == "RR-BLUP" and t == 1:
            # BayesC0! = BayesL! with gamma = [1.0] (BayesC0L.jl:19-23): every marker in the model with the common
            # variance, i.e. the BayesC step with pi = 0 (log pi = -inf: the inclusion test always passes)
            st = backend.sweep_bayesc(schedule, vare, var_effect, 0.0, seed, it)
        elif method == "BayesL":
            # BayesL! (BayesC0L.jl:25-47): lhs = xpx + (vare/var_effect)/gamma_j  <=>  marker variance var_effect*gamma_j
            st = backend.sweep_bayesabc(schedule, vare, var_effect * gamma_arr, np.zeros(p), seed, it)
        else:
            raise ValueError(method)
        # [3] pi
        if ann is not None:
            delta = np.asarray(backend.get_state()[2]).reshape(t, p)
            ann_prior, summary = annot.update_marker_annotation_priors(rng, ann, method, t, delta if t > 1 else delta[0])
            if t == 1:
                pi = summary
            else:
                big_pi = summary
        elif estimate_pi:
            if t == 1 and method == "BayesR":
                pi = rng.dirichlet(st["class_counts"][:len(pi)] + 1.0)          # Pi.jl:11-17
            elif t == 1:
                k_in = st["sum_delta"][0]
                pi = rng.beta(p - k_in + 1, k_in + 1)                            # Pi.jl:7-9
            elif constraint_G:                                                   # MCMC_BayesianAlphabet.jl:300-301
                big_pi = np.array([rng.beta(p - st["sum_delta"][k] + 1, st["sum_delta"][k] + 1) for k in range(t)])
            else:
                big_pi = rng.dirichlet(st["class_counts"][:1 << t] + 1.0)        # Pi.jl:20-42
        # [4] marker effect variance
        if estimate_variance:
            if t == 1 and method == "BayesC":
                k_in = st["sum_delta"][0]
                var_effect = float(np.float32((st["alpha_ss"][0, 0] + df_effect * scale_effect)
                                              / rng.chisq(k_in + df_effect)))   # variance_components.jl:160-162
            elif t == 1 and method == "BayesR":
                var_effect = float(np.float32((st["bayesr_ssq"] + df_effect * scale_effect)
                                              / rng.chisq(st["sum_delta"][0] + df_effect)))  # :166-168
            elif t == 1 and method == "RR-BLUP":
                var_effect = float(np.float32((st["alpha_ss"][0, 0] + df_effect * scale_effect)
                                              / rng.chisq(p + df_effect)))      # :160-162 with nloci = nMarkers
            elif method == "BayesL":
                # :152-165: sample_variance on alpha ./ sqrt(gamma), then the MH update of gammaArray (:191-204)
                a = backend.get_state()[0][:p].astype(np.float64)
                var_effect = float(np.float32((float(np.sum(a * a / gamma_arr)) + df_effect * scale_effect)
                                              / rng.chisq(p + df_effect)))
                Q = a * a / var_effect
                cand = 1.0 / rng.gamma(0.5, 4.0, p)
                accept = rng.uniform(p) < np.exp(Q / 4.0 * (2.0 / gamma_arr - cand))
                gamma_arr = np.where(accept, 2.0 / cand, gamma_arr)
            elif t == 1:
                backend.sample_bayesb_variances(df_effect, scale_effect, seed, it)  # :169-172
            elif constraint_G:                       # variance_components.jl:103-110: diagonal scaled-inv-chi2
                G = np.diag([float(np.float32((st["beta_ss"][k, k] + df_effect * scale_G[k, k]) / rng.chisq(p + df_effect)))
                             for k in range(t)])
            else:
                G = rng.inverse_wishart(df_effect + p, scale_G + st["beta_ss"]).astype(np.float32).astype(np.float64)
        # [5] residual variance
        if estimate_vare:
            if t == 1:
                vare = float(np.float32((st["ycorr_ss"][0, 0] + df_res * scale_res) / rng.chisq(n + df_res)))
            elif constraint_R:
                R = np.diag([float(np.float32((st["ycorr_ss"][k, k] + df_res * scale_R[k, k]) / rng.chisq(n + df_res)))
                             for k in range(t)])
            else:
                R = rng.inverse_wishart(df_res + n, scale_R + st["ycorr_ss"]).astype(np.float32).astype(np.float64)
        ysum = st["ycorr_sum"]
        out["trace"].append((it, float(st["sum_delta"][0]), st["n_active"], st["n_rounds"]))
        # [6] posterior means
        if it > burnin and (it - burnin) % output_samples_frequency == 0:
            nsamples += 1
            backend.accumulate(nsamples, bayesr=(method == "BayesR"))
            out["mu_mean"] += (mu - out["mu_mean"]) / nsamples
            out["mu_mean2"] += (mu ** 2 - out["mu_mean2"]) / nsamples
            ve_now = vare if t == 1 else R
            out["vare_mean"] = out["vare_mean"] + (ve_now - out["vare_mean"]) / nsamples
            out["vare_mean2"] = out["vare_mean2"] + (ve_now ** 2 - out["vare_mean2"]) / nsamples
            if method != "BayesB" and method != "BayesA":
                va_now = var_effect if t == 1 else G
                out["vara_mean"] = out["vara_mean"] + (va_now - out["vara_mean"]) / nsamples
                out["vara_mean2"] = out["vara_mean2"] + (va_now ** 2 - out["vara_mean2"]) / nsamples
            if estimate_pi:
                pi_now = pi if t == 1 else big_pi
                out["pi_mean"] = out["pi_mean"] + (pi_now - out["pi_mean"]) / nsamples
                out["pi_mean2"] = out["pi_mean2"] + (pi_now ** 2 - out["pi_mean2"]) / nsamples
            if ann is not None:             # output.jl:597-600
                ann.accumulate(nsamples)
            if sample_sink is not None:     # marker-effect sample rows (output.jl:467)
                sample_sink(backend.get_state()[0])
            if want_ebv:                    # getEBV per saved sample (output.jl:281-306, 489-495)
                eb = backend
                if ebv_backend is not None:     # output IDs other than the training rows: M_out * alpha (output.jl:302)
                    eb = ebv_backend
                    eb.put_state(*backend.get_state())
                e = np.array([eb.mul_alpha(k) for k in range(t)], dtype=np.float64)
                if ebv_m is None:
                    ebv_m = np.zeros_like(e); ebv_s = np.zeros_like(e)
                d = e - ebv_m
                ebv_m += d / nsamples
                ebv_s += d * (e - ebv_m)
                if want_heritability:
                    # output.jl:498-511: genetic (co)variance of this sample's breeding values over the output IDs
                    # (diagonal when the marker covariance is constrained), h2 = diag(g) / (diag(g) + diag(vare))
                    gv = np.atleast_2d(np.cov(e))
                    if constraint_G:
                        gv = np.diag(np.diag(gv))
                    vres = np.diag(np.atleast_2d(vare if t == 1 else R))
                    gvar_samples.append(gv.reshape(-1) if t > 1 else gv[0, 0])
                    h2_samples.append(np.diag(gv) / (np.diag(gv) + vres))
            if hyper_sink is not None:      # output_MCMC_samples (output.jl:444-515): this saved sample's hyper-parameters
                hyper_sink(dict(vare=(vare if t == 1 else R),
                                vara=(None if method in ("BayesB", "BayesA") else (var_effect if t == 1 else G)),
                                pi=((pi if t == 1 else big_pi) if estimate_pi else None),
                                ebv=(e if want_ebv else None),
                                gvar=(gvar_samples[-1] if want_ebv and want_heritability else None),
                                h2=(h2_samples[-1] if want_ebv and want_heritability else None)))
    if gvar_samples:                # means and standard deviations over the saved samples (output.jl:201-207)
        g_, h_ = np.array(gvar_samples, dtype=np.float64), np.array(h2_samples, dtype=np.float64)
        dd = 1 if len(g_) > 1 else 0
        out.update(gvar_mean=g_.mean(axis=0), gvar_sd=g_.std(axis=0, ddof=dd), h2_mean=h_.mean(axis=0), h2_sd=h_.std(axis=0, ddof=dd))
    if ebv_m is not None:
        out["ebv_mean"] = ebv_m
        out["ebv_var"] = ebv_s / max(nsamples - 1, 1)
    out.update(nsamples=nsamples, mu=mu, vare=vare if t == 1 else R, var_effect=var_effect if t == 1 else G,
               pi=pi if t == 1 else big_pi)
    return out
