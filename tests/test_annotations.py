"""Annotation-aware priors (MCMC/annotation_updates.jl, markers/annotation_setup.jl): validation, start-up state, the
probit Gibbs steps and whole annotated chains through runMCMC (driven over the CPU oracle backend here; the same host
code drives the B200 backend, see test_zz_gpu_late.py).  The cases follow the reference's own
test/unit/test_annotated_bayesc.jl and test_annotated_bayesr.jl; its expected values are built from its own random
stream, so "update == manual composition of the documented steps under one generator" is asserted the same way here."""
import math
import warnings

import numpy as np
import pandas as pd
import pytest
from scipy.special import ndtr, ndtri

import jwas_b200 as jw
from jwas_b200 import annotations as an
from jwas_b200.mcmc import HostRng
from oracle_backend import factory

CODES7 = np.array([[0, 1, 2, 1, 0], [1, 1, 0, 2, 1], [2, 0, 1, 1, 2], [0, 2, 1, 0, 1], [1, 0, 2, 2, 0],
                   [2, 1, 0, 1, 1], [1, 2, 1, 0, 2]])                      # 7 individuals x 5 markers


def _err(fn):
    with pytest.raises(jw.JwasError) as e:
        fn()
    return str(e.value)


# ---------------------------------------------------------------------------------- API and validation
def test_rejects_unsupported_methods_and_bad_shapes():
    A = np.random.default_rng(0).random((5, 2))
    assert "annotations" in _err(lambda: jw.get_genotypes(CODES7, 1.0, method="RR-BLUP", annotations=A, quality_control=False))
    assert "rows" in _err(lambda: jw.get_genotypes(CODES7, 1.0, method="BayesC", annotations=A[:4], quality_control=False))
    assert "multi_trait_sampler" in _err(lambda: jw.get_genotypes(CODES7, 1.0, method="BayesC", annotations=A,
                                                                  quality_control=False, multi_trait_sampler="bogus"))
    const = np.array([[1.0, 0.0], [1.0, 1.0], [1.0, 0.0], [1.0, 1.0], [1.0, 0.5]])
    assert "constant" in _err(lambda: jw.get_genotypes(CODES7, 1.0, method="BayesC", annotations=const, quality_control=False))
    coll = np.array([[0.0, 0.0], [1.0, 1.0], [0.0, 0.0], [1.0, 1.0], [0.5, 0.5]])
    assert "collinear" in _err(lambda: jw.get_genotypes(CODES7, 1.0, method="BayesC", annotations=coll, quality_control=False))
    assert "length" in _err(lambda: jw.get_genotypes(CODES7, 1.0, method="BayesC", annotations=A, Pi=[0.5, 0.5],
                                                     quality_control=False))


def test_qc_filters_annotations_and_prepends_intercept():
    codes = np.array([[0, 1, 2], [1, 1, 1], [2, 1, 0], [1, 1, 1]])          # m2 is fixed -> dropped by QC
    A = np.array([[10.0], [20.0], [30.0]])
    g = jw.get_genotypes(codes, 1.0, method="BayesC", annotations=A, quality_control=True, MAF=0.01)
    assert g.nMarkers == 2
    X = g.annotations.design_matrix
    assert X.shape == (2, 2) and np.all(X[:, 0] == 1.0)
    np.testing.assert_array_equal(X[:, 1:], A[[0, 2]])


def test_forces_estimate_pi():
    A = np.random.default_rng(1).random((5, 2))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        g = jw.get_genotypes(CODES7, 1.0, method="BayesC", annotations=A, estimatePi=False, quality_control=False)
    assert any("estimatePi=false is ignored when annotations are provided" in str(x.message) for x in w)
    assert g.estimatePi is True
    assert hasattr(g.annotations, "variance") and not hasattr(g.annotations, "sd")


@pytest.mark.parametrize("Pi", [0.99, 0.3, np.array([0.9, 0.8, 0.7, 0.6, 0.5])])
def test_bayesc_startup_initialises_probit_intercept_from_starting_pi(Pi):
    A = np.array([[0.0, 1.0], [1.0, 0.0], [1.0, 1.0], [0.0, 0.0], [0.5, 0.5]])
    g = jw.get_genotypes(CODES7, 1.0, method="BayesC", annotations=A, Pi=Pi, quality_control=False)
    jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": g})
    start = np.full(5, Pi) if np.isscalar(Pi) else Pi
    icpt = ndtri(np.mean(1.0 - start))
    ann = g.annotations
    np.testing.assert_array_equal(g.π, start)
    assert ann.coefficients[0] == pytest.approx(icpt) and np.all(ann.coefficients[1:] == 0.0)
    np.testing.assert_allclose(ann.mu, ann.design_matrix @ ann.coefficients)
    np.testing.assert_allclose(ann.mu, icpt)
    # rebuilding the model from the same genotypes gives the same state
    jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": g})
    np.testing.assert_array_equal(g.π, start)
    assert g.annotations.coefficients[0] == pytest.approx(icpt)


def test_bayesr_startup_and_degenerate_pi():
    A = np.array([[0.0, 1.0], [1.0, 0.0], [1.0, 1.0], [0.0, 0.0], [0.5, 0.5]])
    g = jw.get_genotypes(CODES7, 1.0, method="BayesR", annotations=A, quality_control=False)
    ann = g.annotations
    assert ann.nsteps == 3 and ann.nclasses == 4 and ann.coefficients.shape == (3, 3) and ann.snp_pi.shape == (5, 4)
    np.testing.assert_array_equal(ann.snp_pi, np.tile([0.95, 0.03, 0.015, 0.005], (5, 1)))
    assert an.bayesr_annotation_probabilities([0.9, 0.05, 0.03, 0.02]) == pytest.approx((0.1, 0.5, 0.4))
    for bad, word in (([1.0, 0.0, 0.0, 0.0], "nonzero"), ([0.9, 0.1, 0.0, 0.0], "classes 3 or 4"),
                      ([0.0, 0.5, 0.3, 0.2], "delta > 1"), ([0.5, 0.0, 0.3, 0.2], "delta > 2"),
                      ([0.5, 0.2, 0.3, 0.0], "delta > 3")):
        assert word in _err(lambda: jw.get_genotypes(CODES7, 1.0, method="BayesR", annotations=A, Pi=bad,
                                                     quality_control=False))


def test_two_trait_startup_from_joint_pi_and_rejections():
    A = np.array([[0.0, 1.0], [1.0, 0.0], [1.0, 1.0], [0.0, 0.0], [0.5, 0.5]])
    Pi = {(0.0, 0.0): 0.45, (1.0, 0.0): 0.20, (0.0, 1.0): 0.15, (1.0, 1.0): 0.20}
    G = np.array([[1.0, 0.3], [0.3, 1.0]])
    g = jw.get_genotypes(CODES7, G, method="BayesC", annotations=A, Pi=Pi, quality_control=False)
    jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", G, genotypes={"geno": g})
    ann = g.annotations
    assert isinstance(g.π, dict) and ann.nsteps == 3 and ann.nclasses == 4
    assert ann.coefficients.shape == (3, 3) and ann.snp_pi.shape == (5, 4)
    assert np.all(ann.coefficients == 0.0) and np.all(ann.mu == 0.0)
    np.testing.assert_array_equal(ann.snp_pi, np.tile([0.45, 0.20, 0.15, 0.20], (5, 1)))     # columns 00, 10, 01, 11
    # Pi omitted: all markers start in state 11 (annotation_setup.jl:99-103)
    g0 = jw.get_genotypes(CODES7, G, method="BayesC", annotations=A, quality_control=False)
    jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", G, genotypes={"geno": g0})
    np.testing.assert_array_equal(g0.annotations.snp_pi[0], [0.0, 0.0, 0.0, 1.0])
    # three traits, a joint dictionary reused for one trait, no shared-state mass
    g3 = jw.get_genotypes(CODES7, np.eye(3), method="BayesC", annotations=A, quality_control=False)
    assert "supports exactly 2 traits" in _err(lambda: jw.build_model(
        "y1 = intercept + geno\ny2 = intercept + geno\ny3 = intercept + geno", np.eye(3), genotypes={"geno": g3}))
    g1 = jw.get_genotypes(CODES7, 1.0, method="BayesC", annotations=A, Pi=Pi, quality_control=False)
    assert "rebuil" in _err(lambda: jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": g1}))
    bad = {(0.0, 0.0): 0.5, (1.0, 0.0): 0.3, (0.0, 1.0): 0.2, (1.0, 1.0): 0.0}
    gb = jw.get_genotypes(CODES7, G, method="BayesC", annotations=A, Pi=bad, quality_control=False)
    assert "shared state 11" in _err(lambda: jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", G,
                                                            genotypes={"geno": gb}))


# ---------------------------------------------------------------------------------- the sampler
@pytest.fixture(params=[False, True], ids=["numpy", "native"])
def native(request):
    """The update runs in libjwasio (jwann_probit_step) when built, else in numpy: both paths, same expectations (the
    native special functions differ from scipy's in the last bits: equal to 1e-12)."""
    old = an.USE_NATIVE
    an.USE_NATIVE = request.param
    yield request.param
    an.USE_NATIVE = old


def same(actual, desired, native):
    if native:
        np.testing.assert_allclose(actual, desired, rtol=1e-12, atol=1e-13)
    else:
        np.testing.assert_array_equal(actual, desired)


def test_update_is_the_documented_composition_standard_probit(native):
    """annotation sampler uses the standard probit latent variance (test_annotated_bayesc.jl:336-375)."""
    ann = an.MarkerAnnotations(np.ones((1, 1)), variance=4.0)
    ann.mu[0] = 0.3
    prior, summary = an.update_marker_annotation_priors(HostRng([5, 6]), ann, "BayesC", 1, np.array([1]))
    rng = HostRng([5, 6])
    liab, lo, up = an.sample_binary_annotation_liabilities(rng, np.array([0.3]), np.array([1]))
    assert lo[0] == 0.0 and up[0] == np.inf and liab[0] >= 0.0
    coeffs = np.zeros(1); resid = liab - 0.3
    an.gibbs_update_binary_probit_annotation_coefficients(rng, coeffs, np.ones((1, 1)), resid, 4.0)
    same(ann.liability, liab, native)
    same(ann.coefficients, coeffs, native)
    same(ann.mu, coeffs, native)
    same(prior, np.clip(1.0 - ndtr(coeffs), an.EPS, 1 - an.EPS), native)
    assert prior is summary and ann.variance == 4.0            # intercept only: slope variance untouched


def test_update_coordinate_probit_with_shrunken_slopes(native):
    """test_annotated_bayesc.jl:377-433."""
    X = np.array([[1.0, 0.0], [1.0, 1.0], [1.0, 2.0]])
    delta = np.array([0, 1, 0])
    ann = an.MarkerAnnotations(X, variance=0.25)
    ann.coefficients[:] = [-0.4, 0.8]
    ann.mu[:] = X @ ann.coefficients
    prior, _ = an.update_marker_annotation_priors(HostRng([7]), ann, "BayesC", 1, delta)

    rng = HostRng([7])
    mu = X @ np.array([-0.4, 0.8])
    liab, lo, up = an.sample_binary_annotation_liabilities(rng, mu, delta)
    assert np.all(liab[delta == 0] <= 0) and np.all(liab[delta == 1] >= 0)
    coeffs = np.array([-0.4, 0.8]); resid = liab - mu
    # the coordinate update written out: flat intercept, slope shrunk by 1/variance
    z0 = rng.normal()
    c0 = z0 * math.sqrt(1 / 3) + (resid.sum() + 3 * coeffs[0]) / 3
    resid = resid + (coeffs[0] - c0)
    z1 = rng.normal()
    xx = 5.0
    inv = 1.0 / (xx + 1 / 0.25)
    c1 = z1 * math.sqrt(inv) + inv * (X[:, 1] @ resid + xx * coeffs[1])
    var = (c1 ** 2 + 2.0) / rng.chisq(3.0)
    same(ann.liability, liab, native)
    np.testing.assert_allclose(ann.coefficients, [c0, c1], rtol=0, atol=(1e-12 if native else 1e-15))
    assert ann.variance == pytest.approx(var, rel=(1e-11 if native else 1e-14))
    np.testing.assert_allclose(ann.mu, X @ ann.coefficients)
    np.testing.assert_allclose(prior, np.clip(1 - ndtr(ann.mu), an.EPS, 1 - an.EPS))
    np.testing.assert_array_equal(ann.lower_bound, lo); np.testing.assert_array_equal(ann.upper_bound, up)


def test_truncated_liabilities_have_the_right_law():
    rng = HostRng([11])
    m = 200000
    for mu0 in (-2.5, 0.4, 3.0):
        mu = np.full(m, mu0)
        l1, _, _ = an.sample_binary_annotation_liabilities(rng, mu, np.ones(m, int))
        l0, _, _ = an.sample_binary_annotation_liabilities(rng, mu, np.zeros(m, int))
        phi = math.exp(-0.5 * mu0 * mu0) / math.sqrt(2 * math.pi)
        assert l1.min() >= 0 and l0.max() <= 0
        assert l1.mean() == pytest.approx(mu0 + phi / ndtr(mu0), abs=0.01)       # E[l | l > 0]
        assert l0.mean() == pytest.approx(mu0 - phi / ndtr(-mu0), abs=0.01)      # E[l | l < 0]
    # far tails stay finite and on the right side
    l, _, _ = an.sample_binary_annotation_liabilities(rng, np.array([-60.0, 60.0]), np.array([1, 0]))
    assert np.isfinite(l).all() and l[0] >= 0 and l[1] <= 0


def test_probit_gibbs_recovers_the_generating_coefficients(native):
    rng0 = np.random.default_rng(3)
    m = 20000
    A = np.column_stack([rng0.integers(0, 2, m).astype(float), rng0.normal(size=m)])
    truth = np.array([-1.2, 0.9, -0.5])
    X = np.column_stack([np.ones(m), A])
    delta = (X @ truth + rng0.normal(size=m) > 0).astype(int)
    ann = an.MarkerAnnotations(X)
    rng = HostRng([13])
    keep = []
    for it in range(300):
        an.update_marker_annotation_priors(rng, ann, "BayesC", 1, delta)
        if it >= 100:
            keep.append(ann.coefficients.copy())
    np.testing.assert_allclose(np.mean(keep, axis=0), truth, atol=0.06)
    assert 0.05 < ann.variance < 50


def test_nested_indicators_and_prior_rebuild():
    z, active = an.bayesr_nested_step_indicators(np.array([1, 2, 3, 4, 1, 4]))
    np.testing.assert_array_equal(z[0], [0, 1, 1, 1, 0, 1])
    np.testing.assert_array_equal(z[1], [0, 0, 1, 1, 0, 1])
    np.testing.assert_array_equal(z[2], [0, 0, 0, 1, 0, 1])
    assert list(active[0]) == [0, 1, 2, 3, 4, 5] and list(active[1]) == [1, 2, 3, 5] and list(active[2]) == [2, 3, 5]
    z, active = an.bayesc_mt_tree_step_indicators(np.array([0, 1, 0, 1]), np.array([0, 0, 1, 1]))
    np.testing.assert_array_equal(z[0], [0, 1, 1, 1]); np.testing.assert_array_equal(z[1], [0, 0, 0, 1])
    np.testing.assert_array_equal(z[2], [0, 1, 0, 0])
    assert list(active[1]) == [1, 2, 3] and list(active[2]) == [1, 2]
    ann = an.MarkerAnnotations(np.ones((2, 1)), nsteps=3, nclasses=4, coefficients=np.zeros((1, 3)), snp_pi=np.zeros((2, 4)))
    ann.mu[:] = [[0.0, 0.5, -0.5], [40.0, -40.0, 0.0]]
    pr = np.clip(ndtr(ann.mu), an.EPS, 1 - an.EPS)
    an.rebuild_bayesr_nested_priors(ann)
    np.testing.assert_allclose(ann.snp_pi.sum(axis=1), 1.0)
    np.testing.assert_allclose(ann.snp_pi[0], [0.5, 0.5 * (1 - pr[0, 1]), 0.5 * pr[0, 1] * (1 - pr[0, 2]), 0.5 * pr[0, 1] * pr[0, 2]])
    assert np.all(ann.snp_pi > 0)                       # clamped: no class is ever ruled out
    an.rebuild_bayesc_mt_tree_priors(ann)
    np.testing.assert_allclose(ann.snp_pi.sum(axis=1), 1.0)
    np.testing.assert_allclose(ann.snp_pi[0], [0.5, 0.5 * (1 - pr[0, 1]) * pr[0, 2], 0.5 * (1 - pr[0, 1]) * (1 - pr[0, 2]), 0.5 * pr[0, 1]])


def test_nested_step_only_touches_active_markers(native):
    X = np.column_stack([np.ones(6), [0.0, 1.0, 0.0, 1.0, 0.0, 1.0]])
    ann = an.MarkerAnnotations(X, nsteps=3, nclasses=4, coefficients=np.zeros((2, 3)), snp_pi=np.full((6, 4), 0.25))
    delta = np.array([1, 2, 3, 4, 1, 4])
    an.update_marker_annotation_priors(HostRng([17]), ann, "BayesR", 1, delta)
    assert np.all(ann.liability[[0, 4], 1] == 0.0) and np.all(ann.liability[[0, 1, 4], 2] == 0.0)
    assert np.all(np.isinf(ann.lower_bound[[0, 4], 1])) and np.all(np.isinf(ann.upper_bound[[0, 4], 1]))
    assert np.all(ann.liability[[1, 2, 3, 5], 0] >= 0) and np.all(ann.liability[[0, 4], 0] <= 0)
    np.testing.assert_allclose(ann.mu, X @ ann.coefficients)
    np.testing.assert_allclose(ann.snp_pi.sum(axis=1), 1.0)
    # an empty active set leaves the step's coefficients alone
    ann2 = an.MarkerAnnotations(X, nsteps=3, nclasses=4, coefficients=np.zeros((2, 3)), snp_pi=np.full((6, 4), 0.25))
    an.update_marker_annotation_priors(HostRng([17]), ann2, "BayesR", 1, np.ones(6, int))
    assert np.all(ann2.coefficients[:, 1:] == 0.0) and ann2.coefficients[0, 0] != 0.0


# ---------------------------------------------------------------------------------- whole chains
def _annotated_data(n=300, p=400, seed=31, ntraits=1, nqtl=30):
    """QTL sit on markers with annotation 1 = 1 (a fifth of the markers); annotation 2 is noise."""
    from helpers import make_codes
    rng = np.random.default_rng(seed)
    codes = make_codes(n, p, seed)
    a1 = (rng.random(p) < 0.2).astype(float)
    A = np.column_stack([a1, rng.normal(size=p)])
    X = codes - codes.mean(axis=0)
    ys = {}
    for k in range(ntraits):
        qtl = rng.choice(np.flatnonzero(a1), size=nqtl, replace=False)
        gv = X[:, qtl] @ rng.normal(size=nqtl)
        ys[f"y{k + 1}"] = gv + rng.normal(size=n) * gv.std() * 0.6 + 2.0
    ids = [f"id{i}" for i in range(n)]
    return codes, ids, pd.DataFrame({"ID": ids, **ys}), A


def test_annotated_bayesc_chain_finds_the_enriched_annotation(tmp_path):
    codes, ids, ph, A = _annotated_data()
    geno = jw.get_genotypes(codes, False, method="BayesC", Pi=0.9, annotations=A, obsID=ids)
    model = jw.build_model("y1 = intercept + geno", False, genotypes={"geno": geno})
    out = jw.runMCMC(model, ph, chain_length=250, burnin=50, seed=5, _backend_factory=factory, outputEBV=False)
    co = out["annotation coefficients geno"]
    assert list(co.columns) == ["Annotation", "Estimate", "SD"]
    assert list(co["Annotation"]) == ["Intercept", "Annotation_1", "Annotation_2"]
    est = co["Estimate"].to_numpy()
    assert est[1] > 0.3 and est[1] > 2 * co["SD"][1] and abs(est[2]) < 0.5      # enrichment found, noise is not
    pi = out["pi_geno"]
    assert len(pi) == geno.nMarkers and pi["Estimate"].between(0, 1).all()
    a1 = A[:, 0] == 1
    assert pi["Estimate"].to_numpy()[a1].mean() < pi["Estimate"].to_numpy()[~a1].mean()   # annotated: more likely in
    assert out["marker effects geno"]["Model_Frequency"].between(0, 1).all()


def test_annotated_bayesr_and_two_trait_chains_run_and_label_their_steps():
    codes, ids, ph, A = _annotated_data(n=200, p=300, seed=33)
    geno = jw.get_genotypes(codes, False, method="BayesR", annotations=A, obsID=ids)
    model = jw.build_model("y1 = intercept + geno", False, genotypes={"geno": geno})
    out = jw.runMCMC(model, ph, chain_length=40, burnin=10, seed=6, _backend_factory=factory, outputEBV=False)
    co = out["annotation coefficients geno"]
    assert list(co.columns) == ["Annotation", "Step", "Estimate", "SD"] and len(co) == 9
    assert list(co["Step"][:3]) == ["step1_zero_vs_nonzero", "step2_small_vs_larger", "step3_medium_vs_large"]
    assert list(co["Annotation"][:4]) == ["Intercept"] * 3 + ["Annotation_1"]
    assert out["pi_geno"]["Estimate"].sum() == pytest.approx(1.0)
    np.testing.assert_allclose(geno.annotations.snp_pi.sum(axis=1), 1.0)

    codes, ids, ph, A = _annotated_data(n=200, p=300, seed=35, ntraits=2)
    Pi = {(0.0, 0.0): 0.45, (1.0, 0.0): 0.20, (0.0, 1.0): 0.15, (1.0, 1.0): 0.20}
    geno = jw.get_genotypes(codes, False, method="BayesC", annotations=A, Pi=Pi, obsID=ids)
    model = jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", False, genotypes={"geno": geno})
    out = jw.runMCMC(model, ph, chain_length=30, burnin=10, seed=7, _backend_factory=factory, outputEBV=False)
    co = out["annotation coefficients geno"]
    assert list(co["Step"][:3]) == ["step1_zero_vs_active", "step2_11_vs_singleton", "step3_10_vs_01"]
    assert out["pi_geno"]["Estimate"].sum() == pytest.approx(1.0)
    assert list(out["pi_geno"]["π"]) == ["[0.0, 0.0]", "[1.0, 0.0]", "[0.0, 1.0]", "[1.0, 1.0]"]
    # sampler II and constraint=true are refused for annotated 2-trait runs
    g2 = jw.get_genotypes(codes, False, method="BayesC", annotations=A, Pi=Pi, obsID=ids, multi_trait_sampler="II")
    m2 = jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", False, genotypes={"geno": g2})
    assert "sampler I" in _err(lambda: jw.runMCMC(m2, ph, chain_length=5, _backend_factory=factory))
    g3 = jw.get_genotypes(codes, False, method="BayesC", annotations=A, Pi=Pi, obsID=ids, constraint=True)
    m3 = jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", False, genotypes={"geno": g3})
    assert "constraint=false" in _err(lambda: jw.runMCMC(m3, ph, chain_length=5, _backend_factory=factory))


def test_oracle_marker_level_priors_reduce_to_the_global_ones(oracle):
    """The oracle's per_marker_pi path (jwas_oracle.c: BayesR and sampler I) with identical rows is the global-prior
    sweep, bit for bit -- the checker the GPU per-marker parity tests lean on."""
    from helpers import Problem, uniform_starts
    gamma = np.array([0.0, 0.01, 0.1, 1.0]); pi_r = np.array([0.9, 0.05, 0.03, 0.02])
    prob = Problem(oracle, 150, 200, seed=61)
    starts = uniform_starts(200, 64)
    res = []
    for pi in (pi_r, np.tile(pi_r, (200, 1))):
        yc, al, be, de = prob.fresh_state(); de[:] = 1
        for it in (1, 2):
            rc, _ = oracle.sweep_contract(prob.packed, 150, prob.means, prob.xpx, starts, yc, al, None, de,
                                          method=oracle.METHOD_R, nreps_mode=0, independent=False, vare=prob.vary * 0.5,
                                          sigmaSq=prob.vary * 0.01, pi=pi, gamma=gamma, seed=3, it=it, lag=2)
            assert rc == 0
        res.append((yc, al, de))
    for a, b in zip(*res):
        np.testing.assert_array_equal(a, b)
    assert (res[0][2] > 1).sum() > 0
    prob = Problem(oracle, 150, 120, seed=62, ntraits=2)
    starts = uniform_starts(120, 50)
    R = np.array([[1.0, 0.3], [0.3, 1.2]]) * prob.vary * 0.5
    G = np.array([[1.0, 0.4], [0.4, 0.8]]) * prob.vary * 0.02
    big = np.array([0.6, 0.15, 0.1, 0.15])
    res = []
    for bp in (big, np.tile(big, (120, 1))):
        yc, al, be, de = prob.fresh_state()
        for it in (1, 2):
            rc, _ = oracle.sweep_contract(prob.packed, 150, prob.means, prob.xpx, starts, yc, al, be, de,
                                          method=oracle.METHOD_MT1, nreps_mode=0, independent=False, R=R, G=G,
                                          bigPi=bp, seed=9, it=it, lag=1)
            assert rc == 0
        res.append((yc, al, be, de))
    for a, b in zip(*res):
        np.testing.assert_array_equal(a, b)
    assert res[0][3].sum() > 0


def test_packed_backend_keeps_the_raw_marker_mapping_for_annotations(tmp_path):
    """test_annotated_bayesc.jl:190-265: annotations are given per RAW marker; a backend prepared with QC records which
    raw markers it kept (selected.i32, nMarkersAll) and get_genotypes(prefix, annotations=...) filters with it; a
    legacy backend without the mapping is refused."""
    path = str(tmp_path / "annotated_stream_qc.csv")
    open(path, "w").write("ID,m1,m2,m3\na1,0,1,2\na2,1,1,1\na3,2,1,0\na4,1,1,2\n")
    A = np.array([[10.0], [20.0], [30.0]])
    prefix = jw.prepare_streaming_genotypes(path, quality_control=True, MAF=0.01)
    be = jw.load_streaming_backend(prefix)
    assert be["nMarkers"] == 2 and be["nMarkersAll"] == 3 and be["selected_marker_indices"].tolist() == [1, 3]
    assert be["has_raw_marker_mapping"]
    geno = jw.get_genotypes(prefix, 1.0, method="BayesC", annotations=A)
    X = geno.annotations.design_matrix
    assert geno.nMarkers == 2 and X.shape == (2, 2) and np.all(X[:, 0] == 1.0)
    np.testing.assert_array_equal(X[:, 1:], A[[0, 2]])
    assert "rows" in _err(lambda: jw.get_genotypes(prefix, 1.0, method="BayesC", annotations=A[[0, 2]]))
    # legacy manifest: no selected_path / nMarkersAll lines
    meta = prefix + ".meta"
    lines = [l for l in open(meta) if not l.startswith(("selected_path\t", "nMarkersAll\t"))]
    open(meta, "w").writelines(lines)
    assert not jw.load_streaming_backend(prefix)["has_raw_marker_mapping"]
    assert "rebuild the backend" in _err(lambda: jw.get_genotypes(prefix, 1.0, method="BayesC", annotations=A))
    assert jw.get_genotypes(prefix, 1.0, method="BayesC").nMarkers == 2          # without annotations it still loads
    # one of the two entries alone is an inconsistent manifest
    open(meta, "a").write("nMarkersAll\t3\n")
    assert "inconsistent" in _err(lambda: jw.load_streaming_backend(prefix))


@pytest.mark.parametrize("kind", ["BayesC", "BayesR", "BayesC2"])
def test_native_update_equals_numpy_update(kind):
    """jwann_probit_step (threaded C, include/jwas_io.h) against the numpy statements of annotations.py on the same
    generator: several updates in a row, coefficients / liabilities / priors equal to the last bits of the special
    functions, for one thread and for all of them."""
    rng0 = np.random.default_rng(9)
    m = 30011
    X = np.column_stack([np.ones(m), (rng0.random(m) < 0.15).astype(float), rng0.normal(size=m), rng0.random(m)])
    outs = []
    for use in (False, True):
        old = an.USE_NATIVE
        an.USE_NATIVE = use
        try:
            if kind == "BayesC":
                ann = an.MarkerAnnotations(X); ann.coefficients[0] = -1.0; ann.mu[:] = X @ ann.coefficients
            else:
                ann = an.MarkerAnnotations(X, nsteps=3, nclasses=4, coefficients=np.zeros((4, 3)), snp_pi=np.full((m, 4), 0.25))
            rng = HostRng([3, 4])
            r2 = np.random.default_rng(5)
            for it in range(4):
                if kind == "BayesC":
                    delta = (r2.random(m) < 0.2).astype(np.int32)
                    prior, _ = an.update_marker_annotation_priors(rng, ann, "BayesC", 1, delta)
                elif kind == "BayesR":
                    delta = r2.choice([1, 2, 3, 4], size=m, p=[0.8, 0.1, 0.06, 0.04]).astype(np.int32)
                    prior, _ = an.update_marker_annotation_priors(rng, ann, "BayesR", 1, delta)
                else:
                    delta = (r2.random((2, m)) < 0.15).astype(np.int32)
                    prior, _ = an.update_marker_annotation_priors(rng, ann, "BayesC", 2, delta)
            outs.append((ann.coefficients.copy(), np.array(ann.variance, dtype=float), ann.liability.copy(), ann.mu.copy(),
                         np.array(prior).copy(), rng.normal()))
        finally:
            an.USE_NATIVE = old
    for a, b in zip(*outs):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-10)
    assert outs[0][5] == outs[1][5]                     # both paths consumed the generator identically


@pytest.mark.parametrize("kw", [dict(fast_blocks=True), dict(fast_blocks=[1, 40, 41, 100], independent_blocks=True)])
@pytest.mark.parametrize("method,t", [("BayesC", 1), ("BayesR", 1), ("BayesC", 2)])
def test_annotated_chains_on_the_block_schedules(method, t, kw):
    """The reference runs its annotated samplers on the block and independent-block schedules too (CHANGELOG:
    "Independent-block sampler coverage across ... annotated BayesC/BayesR ... dense 2-trait annotated BayesC")."""
    codes, ids, ph, A = _annotated_data(n=120, p=150, seed=71, ntraits=t, nqtl=10)
    eqs = "y1 = intercept + geno" + ("\ny2 = intercept + geno" if t == 2 else "")
    Pi = {(0.0, 0.0): 0.45, (1.0, 0.0): 0.2, (0.0, 1.0): 0.15, (1.0, 1.0): 0.2} if t == 2 else (0.9 if method == "BayesC" else 0.0)
    geno = jw.get_genotypes(codes, False, method=method, Pi=Pi, annotations=A, obsID=ids, center=False)
    model = jw.build_model(eqs, False, genotypes={"geno": geno})
    out = jw.runMCMC(model, ph.iloc[:100], chain_length=200, burnin=2, seed=3, _backend_factory=factory,
                     output_heritability=True, **kw)
    assert "annotation coefficients geno" in out and len(out["EBV_y1"]) == 120
    assert np.isfinite(out["marker effects geno"]["Estimate"].to_numpy(float)).all()
    assert 0 < out["heritability"]["Estimate"][0] < 1
    with pytest.raises(jw.JwasError, match="outer iterations"):
        jw.runMCMC(model, ph.iloc[:100], chain_length=60, burnin=10, seed=3, _backend_factory=factory, fast_blocks=True)


def test_native_special_functions_against_scipy():
    """jwann_phi_cdf / jwann_phi_inv (csrc/io/jw_annot.c) in the regime the sampler uses them: p = u * Phi(s)."""
    import ctypes as C
    from jwas_b200 import _io
    L = _io.lib()
    L.jwann_phi_inv.restype = C.c_double; L.jwann_phi_inv.argtypes = [C.c_double]
    L.jwann_phi_cdf.restype = C.c_double; L.jwann_phi_cdf.argtypes = [C.c_double]
    rng = np.random.default_rng(0)
    s = rng.normal(size=20000) * 4; u = rng.random(20000)
    p = u * ndtr(s)
    mine = np.array([L.jwann_phi_inv(float(v)) for v in p])
    np.testing.assert_allclose(mine, ndtri(p), rtol=0, atol=1e-11)
    xs = np.linspace(-37.5, 9, 4000)
    np.testing.assert_allclose([L.jwann_phi_cdf(float(x)) for x in xs], ndtr(xs), rtol=1e-12)
    assert L.jwann_phi_inv(0.0) == -np.inf and L.jwann_phi_inv(1.0) == np.inf
    assert abs(L.jwann_phi_inv(1e-300) - ndtri(1e-300)) < 1e-6        # far tail: the rational start alone
