"""Reference-arithmetic bridge: runs the faithful `*_ref` restatement of a reference sampler (dense
Float32 `sdot`/`saxpy`, Julia's Float32/Float64 promotion, libm; oracle/jwas_oracle.c) and ANOTHER
implementation of the same sweep side by side on shared replayed draws, sweep after sweep, each arm on
its own state.  The other arm is the contract-arithmetic sweep (CPU tests) or the CUDA library (GPU tests).

What is compared (north_star: "integer inclusion indicators bit-exact, fp effect samples within 1e-5
relative"): after every sweep delta must be EQUAL; alpha and ycorr are compared relative to their scale,
    rel(alpha) = max_j |alpha_j - alpha_ref_j| / max_j |alpha_ref_j|     (<= 1e-5)
    rel(ycorr) = max_i |y_i - y_ref_i| / max_i |y_ref_i|                 (<= 1e-5)
The first sweep at which delta forks (None when it never does) is reported back.

Cases: method in {BayesC, BayesB, BayesR, MT1, MT2} x schedule in {exact, block, independent}.
Reference routines: BayesABC.jl:60-80 / :118-188 / :190-255, BayesR.jl:45-97 / :111-193 / :195-273,
MTBayesABC.jl:57-127 / :129-210 / :243-333 / :335-437 / :439-537 / :539-646.
"""
import numpy as np

GAMMA = np.array([0.0, 0.01, 0.1, 1.0])          # JWAS.jl:12 BAYESR_GAMMA
PI_R = np.array([0.95, 0.03, 0.015, 0.005])      # tools4genotypes.jl:373-375

METHODS = ("BayesC", "BayesB", "BayesR", "MT1", "MT2")
SCHEDULES = ("exact", "block", "independent")


def fast_block_starts(n, p):
    """fast_blocks=true: block size floor(sqrt(nObs)) (JWAS.jl:293-297), 0-based boundaries."""
    b = max(1, int(np.floor(np.sqrt(n))))
    return np.array(list(range(0, p, b)) + [p], dtype=np.int64)


class Hyper:
    """Hyper-parameters of one case, set the way the reference sets its defaults
    (input_data_validation.jl:296-350; tools4genotypes.jl:353-421): half the phenotypic variance is genetic."""

    def __init__(self, prob, method, seed):
        rng = np.random.default_rng(seed + 77)
        p, t = prob.p, prob.t
        sum2pq = float((prob.means.astype(np.float64) * (1 - prob.means / 2)).sum())
        self.method = method
        self.t = t
        if t == 1:
            self.vare = float(np.float32(prob.vary / 2))
            pi0 = 0.95
            self.pi = np.full(p, pi0)
            ve = float(np.float32((prob.vary / 2) / ((1 - pi0) * sum2pq)))
            self.ve = np.full(p, ve)
            if method == "BayesB":                      # per-marker variances (variance_components.jl:169-172)
                self.ve = (ve * rng.uniform(0.5, 2.0, size=p)).astype(np.float32).astype(np.float64)
            # BayesR: sigmaSq = genetic variance / (sum2pq * sum(gamma .* pi)) (tools4genotypes.jl:388-396)
            self.sigma_sq = float(np.float32((prob.vary / 2) / (sum2pq * float((GAMMA * PI_R).sum()))))
        else:
            vy = np.array([prob.y[k].var() for k in range(t)])
            R = np.diag(vy / 2)
            G = np.diag(vy / 2 / (0.5 * sum2pq))
            for a in range(t):
                for b in range(a + 1, t):
                    R[a, b] = R[b, a] = 0.3 * np.sqrt(R[a, a] * R[b, b])
                    G[a, b] = G[b, a] = 0.2 * np.sqrt(G[a, a] * G[b, b])
            self.R = R.astype(np.float32).astype(np.float64)
            self.G = G.astype(np.float32).astype(np.float64)
            big = np.full(1 << t, 0.1 / ((1 << t) - 1))
            big[0] = 0.9
            self.big_pi = big


def ref_sweep(orc, prob, hyp, schedule, starts, state, u, z):
    """One sweep of the reference-arithmetic restatement.  u, z in the contract layout [(rep*t+k)*p + j]."""
    X, xpx, p, t = prob.X, prob.xpx, prob.p, prob.t
    y, a, b, d = state
    m = hyp.method
    indep = schedule == "independent"
    if m in ("BayesC", "BayesB"):
        if schedule == "exact":
            orc.bayesabc_ref(X, xpx, y, a, b, d, hyp.vare, hyp.ve, hyp.pi, u[:p], z[:p])
        else:
            orc.bayesabc_block_ref(X, xpx, starts, 0, indep, y, a, b, d, hyp.vare, hyp.ve, hyp.pi, u, z)
    elif m == "BayesR":
        if schedule == "exact":
            orc.bayesr_ref(X, xpx, y, a, d, hyp.vare, hyp.sigma_sq, PI_R, GAMMA, u[:p], z[:p])
        else:
            orc.bayesr_block_ref(X, xpx, starts, 0, indep, y, a, d, hyp.vare, hyp.sigma_sq, PI_R, GAMMA, u, z)
    elif m == "MT1":
        if schedule == "exact":
            orc.mtbayesabc_I_ref(X, xpx, y, a, b, d, hyp.R, hyp.G, hyp.big_pi, u[:t * p], z[:t * p])
        else:
            orc.mtbayesabc_block_ref(X, xpx, starts, 0, indep, 1, y, a, b, d, hyp.R, hyp.G, hyp.big_pi, u, z)
    elif m == "MT2":
        if schedule == "exact":
            z2 = np.ascontiguousarray(z[:2 * p].reshape(2, p).T)
            orc.mtbayesabc_II_ref(X, xpx, y, a, b, d, hyp.R, hyp.G, hyp.big_pi, u[:p], z2)
        else:
            orc.mtbayesabc_block_ref(X, xpx, starts, 0, indep, 2, y, a, b, d, hyp.R, hyp.G, hyp.big_pi, u, z)
    else:
        raise ValueError(m)


def contract_sweep(orc, prob, hyp, schedule, starts, state, u, z, it, lag=0):
    """One sweep of the oracle's contract-arithmetic twin of the CUDA kernels."""
    y, a, b, d = state
    m = hyp.method
    kw = dict(nreps_mode=0 if schedule == "exact" else 1, independent=(schedule == "independent"),
              seed=1, it=it, u=u, z=z, lag=lag if schedule == "exact" else 0)
    if m in ("BayesC", "BayesB"):
        rc, _ = orc.sweep_contract(prob.packed, prob.n, prob.means, prob.xpx, starts, y, a, b, d, method=orc.METHOD_ABC,
                                   vare=hyp.vare, varEffects=hyp.ve, pi=hyp.pi, **kw)
    elif m == "BayesR":
        rc, _ = orc.sweep_contract(prob.packed, prob.n, prob.means, prob.xpx, starts, y, a, b, d, method=orc.METHOD_R,
                                   vare=hyp.vare, sigmaSq=hyp.sigma_sq, pi=PI_R, gamma=GAMMA, **kw)
    else:
        rc, _ = orc.sweep_contract(prob.packed, prob.n, prob.means, prob.xpx, starts, y, a, b, d,
                                   method=orc.METHOD_MT1 if m == "MT1" else orc.METHOD_MT2,
                                   R=hyp.R, G=hyp.G, bigPi=hyp.big_pi, **kw)
    assert rc == 0


def gpu_sweep(jw, g, hyp, schedule, u, z, it):
    """One sweep of the CUDA library through the C ABI (replayed draws)."""
    sched = {"exact": jw.SCHED_EXACT, "block": jw.SCHED_BLOCK, "independent": jw.SCHED_INDEPENDENT}[schedule]
    m = hyp.method
    if m in ("BayesC", "BayesB"):
        g.sweep_bayesabc(sched, hyp.vare, hyp.ve, hyp.pi, 1, it, u, z)
    elif m == "BayesR":
        g.sweep_bayesr(sched, 1, hyp.vare, hyp.sigma_sq, PI_R, GAMMA, 1, it, u, z)
    elif m == "MT1":
        g.sweep_mt1(sched, hyp.R, hyp.G, hyp.big_pi, 1, it, u, z)
    else:
        g.sweep_mt2(sched, hyp.R, hyp.G, hyp.big_pi, 1, it, u, z)


def ref_state(prob, method):
    """State arrays in the dtypes the *_ref restatements take (delta Float32 as in the reference, Int for BayesR)."""
    t, p = prob.t, prob.p
    y = prob.ycorr0.copy()
    shape = (t, p) if t > 1 else (p,)
    a = np.zeros(shape, np.float32); b = np.zeros(shape, np.float32)
    d = np.zeros(shape, np.int32 if method == "BayesR" else np.float32)
    return [y, a, b, d]


def compare(state_ref, other, method):
    """-> (delta_equal, rel_alpha, rel_ycorr); `other` = (ycorr, alpha, beta, delta) flat arrays."""
    y_r, a_r, _, d_r = state_ref
    y_o, a_o, _, d_o = other
    a_r = a_r.reshape(-1).astype(np.float64); a_o = np.asarray(a_o).reshape(-1).astype(np.float64)
    d_equal = np.array_equal(d_r.reshape(-1).astype(np.int64), np.asarray(d_o).reshape(-1).astype(np.int64))
    sa = np.abs(a_r).max()
    rel_a = float(np.abs(a_o - a_r).max() / sa) if sa > 0 else float(np.abs(a_o).max())
    rel_y = float(np.abs(np.asarray(y_o, np.float64) - y_r).max() / np.abs(y_r).max())
    return d_equal, rel_a, rel_y


def draws(rng, schedule, starts, t, p):
    reps = int(np.diff(starts).max()) if schedule != "exact" else 1
    return rng.random(reps * t * p), rng.standard_normal(reps * t * p)
