"""Annotation-aware marker priors: the host-side step that turns the sampled inclusion indicators into
per-marker prior probabilities for the next sweep (MCMC/annotation_updates.jl, markers/annotation_setup.jl,
validation in markers/readgenotypes.jl:56-158).

The sweep itself is unchanged: `jwas_sweep_bayesabc` takes a per-marker pi, `jwas_sweep_bayesr` a p x 4 matrix
(`per_marker_pi`), `jwas_sweep_mt1` a p x 4 matrix of joint-state priors.  What lives here is O(p x annotations)
per iteration and stays on the host, like the other hyper-parameter draws in mcmc.py:

  binary probit per step:   z_j = 1(l_j > 0),  l_j = x_j' a + e_j,  e_j ~ N(0, 1)
    liabilities             truncated normal given z_j                      (annotation_updates.jl:43-59)
    coefficients            coordinate Gibbs, flat intercept, slopes ~ N(0, s2)         (:98-123)
    slope variance          s2 = (sum_{k>1} a_k^2 + 2) / chi2(ncoef + 1)                (:135-137)
  BayesC  (1 step):         pi_j = 1 - Phi(mu_j)                                        (:177-189)
  BayesR  (3 nested steps): delta > 1; delta > 2 | delta > 1; delta > 3 | delta > 2     (:202-212, :260-267)
  2-trait BayesC (3 steps): active; 11 | active; 10 | singleton                         (:269-304)
"""
import math

import numpy as np
from scipy.special import ndtr, ndtri

EPS = float(np.finfo(np.float64).eps)

# The O(markers x annotations) part of the update runs in libjwasio (jwann_probit_step, threaded C) when the library is
# built; the numpy statements below are the same algorithm and the fall-back.  Both consume the host generator in the
# same order (the uniforms of a step, then one normal per coefficient, then the chi-square), so a chain does not depend
# on which one runs beyond the last bits of the special functions.
USE_NATIVE = True


def _native():
    if not USE_NATIVE:
        return None
    from . import _io
    return _io if _io.available() else None

# joint states of 2-trait BayesC in the column order of snp_pi (annotation_setup.jl:15): 00, 10, 01, 11 --
# which is also the sweep's index sum(delta_k << k), so snp_pi is passed to jwas_sweep_mt1 as it is
MT_STATES = ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 1.0))
STEP_LABELS = {"BayesR": ["step1_zero_vs_nonzero", "step2_small_vs_larger", "step3_medium_vs_large"],
               "BayesC2": ["step1_zero_vs_active", "step2_11_vs_singleton", "step3_10_vs_01"]}


class AnnotationError(ValueError):
    pass


def validate_annotations_input(annotations, nmarkers, method):
    """readgenotypes.jl:56-70."""
    if annotations is False or annotations is None:
        return False
    if method not in ("BayesC", "BayesR"):
        raise AnnotationError('annotations are only supported with method="BayesC" or method="BayesR".')
    try:
        A = np.array(annotations, dtype=np.float64)
    except (TypeError, ValueError):
        A = None
    if A is None or A.ndim != 2:
        raise AnnotationError("annotations must be a numeric matrix with one row per marker.")
    if A.shape[0] != nmarkers:
        raise AnnotationError(f"annotations rows ({A.shape[0]}) must match the number of raw markers ({nmarkers}).")
    return A


def validate_annotation_design(A):
    """readgenotypes.jl:72-88: no constant columns (the intercept is added here), full column rank."""
    if A.shape[1] == 0:
        return
    const = [j + 1 for j in range(A.shape[1]) if len(np.unique(A[:, j])) == 1]
    if const:
        raise AnnotationError(f"annotations contain constant column(s) {const}. Remove constant columns because "
                              "JWAS automatically adds an intercept.")
    X = np.hstack([np.ones((A.shape[0], 1)), A])
    if np.linalg.matrix_rank(X) != X.shape[1]:
        raise AnnotationError("annotations are collinear after adding the intercept. Remove duplicate or perfectly "
                              "collinear annotation columns.")


def bayesr_annotation_probabilities(pi):
    """readgenotypes.jl:90-105: the three conditional step probabilities implied by a 4-class pi."""
    if len(pi) != 4:
        raise AnnotationError("BayesR Pi must have length 4.")
    nonzero = pi[1] + pi[2] + pi[3]
    larger = pi[2] + pi[3]
    if not nonzero > 0:
        raise AnnotationError("Annotated BayesR requires positive nonzero prior mass.")
    if not larger > 0:
        raise AnnotationError("Annotated BayesR requires positive prior mass in classes 3 or 4.")
    p1, p2, p3 = nonzero, larger / nonzero, pi[3] / larger
    if not 0.0 < p1 < 1.0:
        raise AnnotationError("Annotated BayesR requires 0 < Pr(delta > 1) < 1. Adjust Pi so the zero-vs-nonzero "
                              "split is nondegenerate.")
    if not 0.0 < p2 < 1.0:
        raise AnnotationError("Annotated BayesR requires 0 < Pr(delta > 2 | delta > 1) < 1. Adjust Pi so classes 2 "
                              "versus 3/4 are both represented.")
    if not 0.0 < p3 < 1.0:
        raise AnnotationError("Annotated BayesR requires 0 < Pr(delta > 3 | delta > 2) < 1. Adjust Pi so classes 3 "
                              "and 4 are both represented.")
    return p1, p2, p3


class MarkerAnnotations:
    """types.jl:167-216.  `design_matrix` already carries the intercept column."""

    def __init__(self, design_matrix, variance=1.0, nsteps=1, nclasses=2, coefficients=None, snp_pi=None):
        X = np.array(design_matrix, dtype=np.float64)
        if X.ndim != 2:
            raise AnnotationError("annotation design matrix must be a matrix.")
        m, k = X.shape
        self.design_matrix = X
        self.nsteps, self.nclasses = int(nsteps), int(nclasses)
        shape_c = (k,) if nsteps == 1 else (k, nsteps)
        shape_l = (m,) if nsteps == 1 else (m, nsteps)
        self.coefficients = np.zeros(shape_c) if coefficients is None else np.array(coefficients, dtype=np.float64)
        self.mean_coefficients = np.zeros(shape_c)
        self.mean_coefficients2 = np.zeros(shape_c)
        self.variance = float(variance) if nsteps == 1 else np.full(nsteps, float(variance))
        self.liability = np.zeros(shape_l)
        self.mu = np.zeros(shape_l)
        self.lower_bound = np.full(shape_l, -np.inf)
        self.upper_bound = np.full(shape_l, np.inf)
        self.snp_pi = False if snp_pi is None else np.array(snp_pi, dtype=np.float64)
        self.col_ss = (X * X).sum(axis=0)            # x_k'x_k of every column (the diagonal of lhs = X'X)
        self.design_cols = np.asfortranarray(X)      # the same matrix with contiguous columns (coordinate updates)

    def accumulate(self, nsamples):                  # output.jl:597-600
        self.mean_coefficients += (self.coefficients - self.mean_coefficients) / nsamples
        self.mean_coefficients2 += (self.coefficients ** 2 - self.mean_coefficients2) / nsamples

    def names(self):
        return ["Intercept"] + [f"Annotation_{i}" for i in range(1, self.design_matrix.shape[1])]


# ------------------------------------------------------------------------------------ set-up
def build_marker_annotations(A, method, Pi):
    """readgenotypes.jl:127-150: BayesR allocates its nested state at once; BayesC waits for build_model
    (the number of traits decides between one binary step and the three-step tree)."""
    validate_annotation_design(A)
    X = np.hstack([np.ones((A.shape[0], 1)), A])
    if method == "BayesR":
        pi = np.array([0.95, 0.03, 0.015, 0.005]) if (np.isscalar(Pi) and Pi == 0.0) else np.array(Pi, dtype=np.float64)
        bayesr_annotation_probabilities(pi)
        return MarkerAnnotations(X, nsteps=3, nclasses=4, coefficients=np.zeros((X.shape[1], 3)),
                                 snp_pi=np.tile(pi, (X.shape[0], 1)))
    return MarkerAnnotations(X)


def annotation_starting_pi(method, Pi, nmarkers):
    """readgenotypes.jl:111-125."""
    if method != "BayesC" or isinstance(Pi, dict):
        return Pi.copy() if hasattr(Pi, "copy") else Pi
    if isinstance(Pi, (list, tuple, np.ndarray)):
        if len(Pi) != nmarkers:
            raise AnnotationError(f"Annotated BayesC starting Pi vector length {len(Pi)} must match the number of "
                                  f"markers ({nmarkers}).")
        return np.array(Pi, dtype=np.float64)
    return np.full(nmarkers, float(Pi))


def mt_row_from_dict(pi):
    """annotation_setup.jl:41-48."""
    row = np.zeros(4)
    for state, prob in pi.items():
        if len(state) != 2:
            raise AnnotationError("Annotated multi-trait BayesC v1 expects 2-trait state labels.")
        d = (int(round(float(state[0]))), int(round(float(state[1]))))
        if d not in ((0, 0), (1, 0), (0, 1), (1, 1)):
            raise AnnotationError("Annotated multi-trait BayesC v1 expects binary 2-trait state labels.")
        row[d[0] + 2 * d[1]] = float(prob)
    if abs(row.sum() - 1.0) > 1e-8:
        raise AnnotationError("Summation of probabilities of Pi is not equal to one.")
    return row


def finalize_marker_annotation_setup(geno):
    """annotation_setup.jl:141-153, called from build_model once `ntraits` is known."""
    ann = geno.annotations
    if ann is False or ann is None or geno.method != "BayesC":
        return
    X = ann.design_matrix
    raw = geno.annotation_start_pi
    p = geno.nMarkers
    if geno.ntraits == 1:                                   # annotation_setup.jl:73-90, :64-71
        if isinstance(raw, dict):
            raise AnnotationError("Annotated BayesC genotypes initialized with a joint Pi dictionary cannot be "
                                  "rebuilt for a single-trait model. Use a fresh get_genotypes call with "
                                  "scalar/vector Pi for single-trait analysis.")
        start = np.array(raw, dtype=np.float64) if isinstance(raw, np.ndarray) else np.full(p, float(raw))
        if len(start) != p:
            raise AnnotationError(f"Annotated BayesC starting Pi vector length {len(start)} must match the number "
                                  f"of markers ({p}).")
        geno.annotations = ann = MarkerAnnotations(X)
        incl = min(max(float(np.mean(1.0 - start)), EPS), 1.0 - EPS)
        ann.coefficients[0] = float(ndtri(incl))
        ann.mu[:] = X @ ann.coefficients
        geno.π = start
        return
    if geno.ntraits != 2:
        raise AnnotationError("Annotated multi-trait BayesC currently supports exactly 2 traits.")
    if isinstance(raw, dict):                               # annotation_setup.jl:92-123
        row = mt_row_from_dict(raw)
    elif (np.isscalar(raw) and raw == 0.0) or (isinstance(raw, np.ndarray) and len(raw) == p and not np.any(raw)):
        row = np.array([0.0, 0.0, 0.0, 1.0])
    else:
        raise AnnotationError("Annotated multi-trait BayesC requires Pi=0.0 or a joint Pi dictionary.")
    if not row[1] + row[3] > 0.0:                           # annotation_setup.jl:52-61
        raise AnnotationError("Annotated multi-trait BayesC requires positive startup prior mass in states {10,11} "
                              "for trait 1.")
    if not row[2] + row[3] > 0.0:
        raise AnnotationError("Annotated multi-trait BayesC requires positive startup prior mass in states {01,11} "
                              "for trait 2.")
    if not row[3] > 0.0:
        raise AnnotationError("Annotated multi-trait BayesC requires positive startup prior mass in shared state 11.")
    geno.annotations = MarkerAnnotations(X, nsteps=3, nclasses=4, coefficients=np.zeros((X.shape[1], 3)),
                                         snp_pi=np.tile(row, (p, 1)))
    geno.π = {MT_STATES[i]: float(row[i]) for i in range(4)}


# ------------------------------------------------------------------------------------ the update
def sample_binary_annotation_liabilities(rng, mu, response):
    """annotation_updates.jl:20-59: l ~ N(mu, 1) truncated to (-inf, 0] when z = 0 and to [0, inf) when z = 1, by
    inversion in the lower tail of whichever side is kept (accurate for |mu| up to ~37).
    Returns (liability, lower, upper)."""
    mu = np.asarray(mu, dtype=np.float64)
    one = np.asarray(response) != 0
    lower = np.where(one, 0.0, -np.inf)
    upper = np.where(one, np.inf, 0.0)
    u = np.clip(rng.uniform(mu.shape[0]), 1e-300, 1.0)
    # z = 1: l = mu - w with w ~ N(0,1) | w < mu;   z = 0: l = mu + e with e ~ N(0,1) | e < -mu
    s = np.where(one, mu, -mu)
    w = ndtri(u * ndtr(s))
    w = np.where(np.isfinite(w), w, np.minimum(s, -37.0))      # ndtr underflow: the boundary itself
    liab = np.where(one, np.maximum(mu - w, 0.0), np.minimum(mu + w, 0.0))
    return liab, lower, upper


def gibbs_update_binary_probit_annotation_coefficients(rng, coeffs, X, latent_residual, coef_prior_var, col_ss=None):
    """annotation_updates.jl:98-123: one coordinate pass; `coeffs` and `latent_residual` are updated in place."""
    m = X.shape[0]
    old = coeffs[0]
    inv_lhs = 1.0 / m
    ahat = inv_lhs * (latent_residual.sum() + m * old)
    coeffs[0] = rng.normal() * math.sqrt(inv_lhs) + ahat
    latent_residual += old - coeffs[0]
    if not X.flags.f_contiguous and X.shape[1] > 1:
        X = np.asfortranarray(X)                     # contiguous columns: the dots below run at memory speed
    for k in range(1, X.shape[1]):
        old = coeffs[k]
        xk = X[:, k]
        diag = float(xk @ xk) if col_ss is None else float(col_ss[k])
        inv_lhs = 1.0 / (diag + 1.0 / coef_prior_var)
        ahat = inv_lhs * (float(xk @ latent_residual) + diag * old)
        coeffs[k] = rng.normal() * math.sqrt(inv_lhs) + ahat
        latent_residual += xk * (old - coeffs[k])


def sample_annotation_effect_variance(rng, coeffs):
    """annotation_updates.jl:135-137."""
    return (float(np.sum(coeffs[1:] ** 2)) + 2.0) / rng.chisq(len(coeffs) + 1.0)


def clamp_prob(x):
    return np.clip(x, EPS, 1.0 - EPS)


def _bounds(response):
    one = np.asarray(response) != 0
    return np.where(one, 0.0, -np.inf), np.where(one, np.inf, 0.0)


def update_bayesc_binary_priors(rng, ann, delta):
    """annotation_updates.jl:177-189.  Returns the per-marker pi (probability of a ZERO effect)."""
    nat = _native()
    if nat is not None:
        m, k = ann.design_cols.shape
        u = rng.uniform(m)
        zn = np.array([rng.normal() for _ in range(k)])
        ann.lower_bound, ann.upper_bound = _bounds(delta)
        ann.liability = np.ascontiguousarray(ann.liability, dtype=np.float64)
        ann.mu = np.ascontiguousarray(ann.mu, dtype=np.float64)
        nat.probit_step(ann.design_cols, None, np.ascontiguousarray(delta, dtype=np.int32), ann.coefficients, ann.variance,
                        u, zn, ann.liability, ann.mu)
        if k > 1:
            ann.variance = sample_annotation_effect_variance(rng, ann.coefficients)
        return nat.probit_probability(ann.mu, complement=True)
    ann.liability, ann.lower_bound, ann.upper_bound = sample_binary_annotation_liabilities(rng, ann.mu, delta)
    resid = ann.liability - ann.mu
    gibbs_update_binary_probit_annotation_coefficients(rng, ann.coefficients, ann.design_cols, resid,
                                                       ann.variance, ann.col_ss)
    ann.mu = ann.design_matrix @ ann.coefficients
    if len(ann.coefficients) > 1:
        ann.variance = sample_annotation_effect_variance(rng, ann.coefficients)
    return clamp_prob(1.0 - ndtr(ann.mu))


def sample_nested_annotation_probit_step(rng, ann, step, response, active):
    """annotation_updates.jl:224-258: one conditional binary step on the markers in `active`."""
    X = ann.design_matrix
    coeffs = ann.coefficients[:, step].copy()
    ann.mu[:, step] = X @ coeffs
    ann.lower_bound[:, step] = -np.inf
    ann.upper_bound[:, step] = np.inf
    if len(active) == 0:
        return
    nat = _native()
    if nat is not None:
        m, k = X.shape
        u = rng.uniform(len(active))
        zn = np.array([rng.normal() for _ in range(k)])
        resp = np.ascontiguousarray(response, dtype=np.int32)
        lo, up = _bounds(resp[active])
        ann.lower_bound[active, step] = lo
        ann.upper_bound[active, step] = up
        liab = np.ascontiguousarray(ann.liability[:, step]); mu = np.ascontiguousarray(ann.mu[:, step])
        nat.probit_step(ann.design_cols, (None if len(active) == m else active), resp, coeffs, ann.variance[step], u, zn,
                        liab, mu)
        ann.liability[:, step] = liab
        if k > 1:
            ann.variance[step] = sample_annotation_effect_variance(rng, coeffs)
        ann.coefficients[:, step] = coeffs
        ann.mu[:, step] = mu
        return
    Xa = X[active]
    mu_a = ann.mu[active, step]
    liab, lo, up = sample_binary_annotation_liabilities(rng, mu_a, response[active])
    ann.liability[active, step] = liab
    ann.lower_bound[active, step] = lo
    ann.upper_bound[active, step] = up
    resid = liab - mu_a
    gibbs_update_binary_probit_annotation_coefficients(rng, coeffs, Xa, resid, ann.variance[step])
    if X.shape[1] > 1:
        ann.variance[step] = sample_annotation_effect_variance(rng, coeffs)
    ann.coefficients[:, step] = coeffs
    ann.mu[:, step] = X @ coeffs


def bayesr_nested_step_indicators(delta):
    """annotation_updates.jl:202-212 (delta in 1..4)."""
    delta = np.asarray(delta)
    z = [(delta > c).astype(np.int64) for c in (1, 2, 3)]
    return z, [np.arange(len(delta)), np.flatnonzero(z[0]), np.flatnonzero(z[1])]


def _step_probabilities(mu):
    nat = _native()
    if nat is not None:
        return nat.probit_probability(mu).reshape(mu.shape)
    return clamp_prob(ndtr(mu))


def rebuild_bayesr_nested_priors(ann):
    """annotation_updates.jl:260-267."""
    pr = _step_probabilities(ann.mu)
    ann.snp_pi[:, 0] = 1.0 - pr[:, 0]
    ann.snp_pi[:, 1] = pr[:, 0] * (1.0 - pr[:, 1])
    ann.snp_pi[:, 2] = pr[:, 0] * pr[:, 1] * (1.0 - pr[:, 2])
    ann.snp_pi[:, 3] = pr[:, 0] * pr[:, 1] * pr[:, 2]


def bayesc_mt_tree_step_indicators(d1, d2):
    """annotation_updates.jl:269-292: states 1..4 = 00, 10, 01, 11."""
    d1 = np.asarray(d1).astype(np.int64); d2 = np.asarray(d2).astype(np.int64)
    states = 1 + (d1 != 0) + 2 * (d2 != 0)
    z = [(states != 1).astype(np.int64), (states == 4).astype(np.int64), (states == 2).astype(np.int64)]
    return z, [np.arange(len(states)), np.flatnonzero(z[0]), np.flatnonzero((states == 2) | (states == 3))]


def rebuild_bayesc_mt_tree_priors(ann):
    """annotation_updates.jl:294-304."""
    pr = _step_probabilities(ann.mu)
    p1, p2, p3 = pr[:, 0], pr[:, 1], pr[:, 2]
    ann.snp_pi[:, 0] = 1.0 - p1
    ann.snp_pi[:, 1] = p1 * (1.0 - p2) * p3
    ann.snp_pi[:, 2] = p1 * (1.0 - p2) * (1.0 - p3)
    ann.snp_pi[:, 3] = p1 * p2


def update_marker_annotation_priors(rng, ann, method, ntraits, delta):
    """annotation_updates.jl:322-364.  `delta`: the indicators of the sweep just finished -- (p,) for one trait,
    (2, p) for two.  Returns what the next sweep takes as its prior -- (p,) pi for BayesC, the (p, 4) class or
    joint-state matrix otherwise -- and the summary the reference keeps in Mi.pi (the same vector for BayesC,
    the column means otherwise)."""
    if ann.nsteps == 1:
        pi = update_bayesc_binary_priors(rng, ann, np.asarray(delta).reshape(-1))
        return pi, pi
    if method == "BayesR":
        z, active = bayesr_nested_step_indicators(np.asarray(delta).reshape(-1))
        for step in range(ann.nsteps):
            sample_nested_annotation_probit_step(rng, ann, step, z[step], active[step])
        rebuild_bayesr_nested_priors(ann)
        return ann.snp_pi, ann.snp_pi.mean(axis=0)
    if method == "BayesC" and ntraits == 2:
        d = np.asarray(delta).reshape(2, -1)
        z, active = bayesc_mt_tree_step_indicators(d[0], d[1])
        for step in range(ann.nsteps):
            sample_nested_annotation_probit_step(rng, ann, step, z[step], active[step])
        rebuild_bayesc_mt_tree_priors(ann)
        return ann.snp_pi, ann.snp_pi.mean(axis=0)
    raise AnnotationError("Unsupported annotation configuration.")
