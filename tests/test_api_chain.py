"""Host-side logic of the reference-facing API (get_genotypes / build_model / runMCMC) driven over
the CPU oracle backend: schedules, priors, outputs, error behaviour.  Mirrors the reference's own
smoke / reproducibility / schedule / constraint tests."""
import os

import numpy as np
import pandas as pd
import pytest

import jwas_b200 as jw
from helpers import make_codes, make_phenotype
from oracle_backend import factory
from oracle import pyoracle as orc


def make_data(n=120, p=90, seed=3, ntraits=1, missing=0.0):
    codes = make_codes(n, p, seed, missing=missing).astype(float)
    packed = orc.pack_codes(np.where(codes == 9, 9, codes).astype(int))
    means, _ = orc.marker_stats(packed, n)
    X = orc.dense_centered(packed, n, means)
    Y = make_phenotype(X, seed, ntraits=ntraits)
    ids = [f"a{i + 1}" for i in range(n)]
    ph = pd.DataFrame({"ID": ids, **{f"y{k + 1}": Y[k] for k in range(ntraits)}})
    return codes, ids, ph


def test_bayesc_run_outputs_and_reproducibility():
    codes, ids, ph = make_data()
    outs = []
    for _ in range(2):
        geno = jw.get_genotypes(codes, 1.0, method="BayesC", Pi=0.9, obsID=ids, quality_control=False)
        model = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
        out = jw.runMCMC(model, ph, chain_length=30, burnin=10, seed=2026, _backend_factory=factory)
        outs.append(out)
    out = outs[0]
    # runtests.jl:265-283: keys, ranges
    for key in ("location parameters", "residual variance", "marker effects geno", "pi_geno", "EBV_y1"):
        assert key in out
    me = out["marker effects geno"]
    assert list(me.columns) == ["Trait", "Marker_ID", "Estimate", "SD", "Model_Frequency"]
    assert len(me) == 90 and me["Model_Frequency"].between(0, 1).all()
    assert 0 < out["pi_geno"]["Estimate"][0] < 1
    assert len(out["EBV_y1"]) == 120
    # runtests.jl:302-320: same seed -> identical results
    np.testing.assert_array_equal(outs[0]["marker effects geno"]["Estimate"], outs[1]["marker effects geno"]["Estimate"])
    assert outs[0]["residual variance"]["Estimate"][0] == outs[1]["residual variance"]["Estimate"][0]
    assert model.MCMCinfo.chain_length == 30


@pytest.mark.parametrize("method", ["BayesA", "BayesB", "BayesR"])
def test_other_methods_run(method):
    codes, ids, ph = make_data(seed=5)
    geno = jw.get_genotypes(codes, 1.0, method=method, obsID=ids, Pi=(0.0 if method != "BayesB" else 0.8))
    model = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
    out = jw.runMCMC(model, ph, chain_length=12, burnin=2, seed=1, outputEBV=False, _backend_factory=factory)
    me = out["marker effects geno"]
    assert me["Model_Frequency"].between(0, 1).all()
    if method == "BayesR":           # test_bayesr.jl:339-342
        assert out["pi_geno"]["Estimate"].sum() == pytest.approx(1.0)
        assert len(out["pi_geno"]) == 4
    if method == "BayesA":           # every marker stays in the model
        assert (me["Model_Frequency"] == 1).all()


@pytest.mark.parametrize("method", ["RR-BLUP", "BayesL"])
def test_dense_methods_run_and_predict(method):
    """RR-BLUP (BayesC0! = BayesL! with gamma = 1, BayesC0L.jl:19-47) and the Bayesian Lasso: every marker is in
    the model every iteration, pi is not estimated (input_data_validation.jl:24-31), EBVs follow the simulated
    genetic values, and the same seed reproduces the run."""
    n, p = 150, 60
    codes, ids, ph = make_data(n=n, p=p, seed=8)
    outs = []
    for _ in range(2):
        geno = jw.get_genotypes(codes, 1.0, method=method, obsID=ids, Pi=0.5, estimatePi=True, quality_control=False)
        assert geno.π == 0.0 and geno.estimatePi is False
        model = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
        outs.append(jw.runMCMC(model, ph, chain_length=60, burnin=20, seed=11, _backend_factory=factory))
    out = outs[0]
    me = out["marker effects geno"]
    assert (me["Model_Frequency"] == 1).all()
    assert "pi_geno" not in out
    np.testing.assert_array_equal(outs[0]["marker effects geno"]["Estimate"], outs[1]["marker effects geno"]["Estimate"])
    # the fitted genetic values explain the phenotype they were simulated from
    ebv = out["EBV_y1"]["EBV"].to_numpy(dtype=float)
    y = ph["y1"].to_numpy(dtype=float)
    assert np.corrcoef(ebv, y)[0, 1] > 0.5
    assert out["residual variance"]["Estimate"][0] > 0


def test_bayesl_multitrait_is_rejected():
    """Multi-trait BayesL (MTBayesL!, per-marker gamma) is outside this backend; multi-trait RR-BLUP is not
    (test_multitrait_rrblup_chain)."""
    codes, ids, ph = make_data(ntraits=2, seed=4)
    geno = jw.get_genotypes(codes, np.eye(2), method="BayesL", obsID=ids, quality_control=False)
    model = jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", np.eye(2), genotypes={"geno": geno})
    with pytest.raises(jw.JwasError, match="multi-trait"):
        jw.runMCMC(model, ph, chain_length=4, seed=1, _backend_factory=factory)


def test_multitrait_bayesc_run():
    codes, ids, ph = make_data(ntraits=2, seed=7)
    geno = jw.get_genotypes(codes, np.array([[1.0, 0.5], [0.5, 1.0]]), method="BayesC", obsID=ids)
    model = jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", np.array([[1.0, 0.5], [0.5, 1.0]]),
                           genotypes={"geno": geno})
    out = jw.runMCMC(model, ph, chain_length=12, burnin=2, seed=123, _backend_factory=factory)
    assert len(out["residual variance"]) == 4          # test_multitrait_mcmc.jl:120
    assert len(out["marker effects geno"]) == 2 * geno.nMarkers
    assert out["pi_geno"]["Estimate"].sum() == pytest.approx(1.0)
    assert "EBV_y1" in out and "EBV_y2" in out


def test_fast_blocks_schedule_semantics():
    # JWAS.jl:293-316; test_misc_coverage.jl:116-209
    codes, ids, ph = make_data(n=50, p=40, seed=9)

    def run(**kw):
        geno = jw.get_genotypes(codes, 1.0, method="BayesC", obsID=ids, quality_control=False)
        model = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
        out = jw.runMCMC(model, ph, seed=123, outputEBV=False, _backend_factory=factory, **kw)
        return model, out

    m, _ = run(chain_length=21, fast_blocks=True)                 # block = floor(sqrt(50)) = 7
    assert m.MCMCinfo.fast_blocks == list(range(1, 41, 7)) and m.MCMCinfo.chain_length == 3
    m, _ = run(chain_length=6, fast_blocks=[1, 3, 5], independent_blocks=True)
    assert m.MCMCinfo.fast_blocks == [1, 3, 5] and m.MCMCinfo.chain_length == 6
    assert m.MCMCinfo.independent_blocks is True
    with pytest.raises(jw.JwasError, match="independent_blocks=true requires fast_blocks"):
        run(chain_length=6, independent_blocks=True)
    for bad in ([2, 4], [1, 3, 3], [3, 1], [1, 100]):
        with pytest.raises(jw.JwasError, match="fast_blocks"):
            run(chain_length=6, fast_blocks=bad)
    with pytest.raises(jw.JwasError, match="at least two block starts"):
        run(chain_length=6, fast_blocks=1000)


def test_constraint_errors():
    # input_data_validation.jl:45-66, 81-111 style guards for what this backend does not cover
    codes, ids, ph = make_data(n=30, p=20)
    with pytest.raises(jw.JwasError, match="outside the GPU marker-sweep path"):
        jw.get_genotypes(codes, 1.0, method="GBLUP")
    with pytest.raises(jw.JwasError, match="Only 0/1/2 genotypes"):
        jw.get_genotypes(codes + 0.5, 1.0)
    with pytest.raises(jw.JwasError, match="outside the GPU marker-sweep path"):
        jw.set_random(None, "x")
    geno = jw.get_genotypes(codes, 1.0, obsID=ids, quality_control=False)
    with pytest.raises(jw.JwasError, match="not a genotype term"):
        jw.build_model("y1 = intercept + x1 + geno", 1.0, genotypes={"geno": geno})
    model = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
    with pytest.raises(jw.JwasError, match="heterogeneous_residuals"):
        jw.runMCMC(model, ph, heterogeneous_residuals=True, _backend_factory=factory)
    ph2 = ph.copy(); ph2.loc[0, "ID"] = "zzz"
    with pytest.raises(jw.JwasError, match="without genotypes"):
        jw.runMCMC(model, ph2, _backend_factory=factory)


def test_quality_control_and_phenotype_subset():
    codes, ids, ph = make_data(n=60, p=30, seed=11)
    codes[:, 3] = 1.0                       # fixed locus -> removed (readgenotypes.jl:388-399)
    codes[:, 7] = 0.0; codes[0, 7] = 1.0    # MAF = 1/120 < 0.01 -> removed
    geno = jw.get_genotypes(codes, 1.0, obsID=ids)
    assert geno.nMarkers == 28 and "m4" not in geno.markerID and "m8" not in geno.markerID
    model = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
    sub = ph.iloc[::-1].iloc[:40].reset_index(drop=True)     # subset + reorder: genotypes follow phenotypes
    seen = {}

    def fac(packed, n, t, starts, means=None):
        b = factory(packed, n, t, starts, means=means)
        seen["out" if "b" in seen else "b"] = b          # first call: the training backend; second: the EBV rows
        return b

    out = jw.runMCMC(model, sub, chain_length=5, seed=1, _backend_factory=fac)
    # check_outputID (input_data_validation.jl:143-196): EBVs for ALL genotyped individuals by default, in genotype order
    assert list(out["EBV_y1"]["ID"]) == ids
    # ... computed as M_out * alpha with the same full-sample centring (align_genotypes; getEBV, output.jl:300-304)
    keep = [j for j in range(30) if j not in (3, 7)]
    xall = codes[:, keep].astype(np.float64) - np.asarray(geno.marker_means, np.float64)[None, :]
    last = xall @ seen["b"].alpha.astype(np.float64)
    assert np.corrcoef(out["EBV_y1"]["EBV"].to_numpy(float), last)[0, 1] > 0.5
    np.testing.assert_allclose(seen["out"].mul_alpha(0), last, rtol=1e-4, atol=1e-5)
    # outputEBV(model, IDs) (output.jl:60-69) names the individuals (honoured with output_heritability=false, as in the
    # reference); IDs without genotypes are dropped with a warning
    jw.outputEBV(model, ["a3", "a50", "nobody"])
    with pytest.warns(UserWarning, match="not a subset of genotyped individuals"):
        out2 = jw.runMCMC(model, sub, chain_length=5, seed=1, output_heritability=False, _backend_factory=factory)
    assert list(out2["EBV_y1"]["ID"]) == ["a3", "a50"]
    full = out["EBV_y1"].set_index("ID")["EBV"]
    np.testing.assert_allclose(out2["EBV_y1"]["EBV"].to_numpy(float), [full["a3"], full["a50"]], rtol=1e-6)
    # the training rows themselves, in training order: no second backend
    jw.outputEBV(model, list(sub["ID"]))
    out3 = jw.runMCMC(model, sub, chain_length=5, seed=1, output_heritability=False, _backend_factory=factory)
    assert list(out3["EBV_y1"]["ID"]) == list(sub["ID"])
    # with output_heritability=true (the default, as in the reference) every genotyped individual is an output ID again
    # (check_outputID, input_data_validation.jl:167-174)
    out4 = jw.runMCMC(model, sub, chain_length=5, seed=1, _backend_factory=factory)
    assert list(out4["EBV_y1"]["ID"]) == ids and "heritability" in out4
    np.testing.assert_allclose(out3["EBV_y1"]["EBV"].to_numpy(float), [full[i] for i in sub["ID"]], rtol=1e-6)
    # centring stays on ALL genotyped individuals (readgenotypes.jl:372-385 runs before the alignment of
    # JWAS.jl:381-402): the backend keeps the full-sample means and xpx is the subset's sum of squares about them
    b = seen["b"]
    np.testing.assert_array_equal(b.means, np.asarray(geno.marker_means, np.float32))
    rows = [ids.index(i) for i in sub["ID"]]
    keep = [j for j in range(30) if j not in (3, 7)]
    x = codes[np.ix_(rows, keep)].astype(np.float64) - np.asarray(geno.marker_means, np.float64)[None, :]
    np.testing.assert_allclose(b.xpx, (x * x).sum(axis=0), rtol=1e-6)


@pytest.mark.parametrize("t", [2, 3, 4])
def test_multitrait_prior_df_and_scale_follow_the_reference(t):
    """build_MME.jl:108-110, 128-134: df += nModels for t > 1; scale = val * (df - t - 1) = val * (df_user - 1)
    (tools4genotypes.jl:417, input_data_validation.jl:345); constraint=true takes the traits back out and uses
    diag(scale / (df - 1)) * (df - 2) / df (input_data_validation.jl:530-558)."""
    codes, ids, ph = make_data(ntraits=t, seed=5)
    for constraint in (False, True):
        geno = jw.get_genotypes(codes, np.eye(t), method="BayesC", obsID=ids, constraint=constraint)
        eq = "\n".join(f"y{k + 1} = intercept + geno" for k in range(t))
        model = jw.build_model(eq, np.eye(t) * 2.0, genotypes={"geno": geno}, constraint=constraint)
        assert model.R.df == 4.0 + t and geno.G.df == 4.0 + t
        seen = {}
        orig = jw.mcmc.run_chain

        def spy(backend, **kw):
            seen.update(kw)
            return orig(backend, **kw)

        jw.api.mcmc.run_chain = spy
        try:
            jw.runMCMC(model, ph, chain_length=2, seed=1, outputEBV=False, _backend_factory=factory)
        finally:
            jw.api.mcmc.run_chain = orig
        if not constraint:
            assert seen["df_res"] == 4.0 + t and seen["df_effect"] == 4.0 + t
            np.testing.assert_allclose(seen["scale_R"], np.eye(t) * 2.0 * 3.0)              # R * (df_user - 1)
            np.testing.assert_allclose(seen["scale_G"], np.asarray(geno.G.val) * 3.0)
        else:
            assert seen["df_res"] == 4.0 and seen["df_effect"] == 4.0
            np.testing.assert_allclose(seen["scale_R"], np.eye(t) * 2.0 * 3.0 / 3.0 * 2.0 / 4.0)   # diag(R) * (nu-2)/nu
            assert np.all(np.linalg.eigvalsh(seen["scale_G"]) > 0)


def test_jgb2_backend_files_roundtrip(tmp_path):
    """prepare_streaming_genotypes / load_streaming_backend write and read the reference's packed
    backend files (streaming_genotypes.jl:77-95, 636-654, 884-971): same manifest keys, .jgb2 bit
    layout and Float32 side-cars."""
    codes = np.array([[0, 1, 2, 0], [1, 0, 1, 2], [2, 9, 0, 1], [0, 2, 1, 0], [1, 1, 2, 2], [2, 0, 0, 1]], float)
    csv = tmp_path / "geno_missing.csv"
    pd.DataFrame(np.column_stack([[f"a{i}" for i in range(1, 7)], codes.astype(int)]),
                 columns=["ID", "m1", "m2", "m3", "m4"]).to_csv(csv, index=False)
    prefix = jw.prepare_streaming_genotypes(str(csv), quality_control=True)
    assert os.path.isfile(prefix + ".jgb2") and os.path.isfile(prefix + ".meta")
    meta = dict(l.rstrip("\n").split("\t", 1) for l in open(prefix + ".meta"))
    for key in ("version", "data_path", "obs_path", "marker_path", "selected_path", "mean_path", "xp_path",
                "afreq_path", "nObs", "nMarkers", "nMarkersAll", "stride_bytes", "centered", "sum2pq"):
        assert key in meta
    raw = np.fromfile(prefix + ".jgb2", dtype=np.uint8).reshape(4, 2)
    assert raw[1].tolist() == [0xB1, 0x01]                    # test_streaming_codec.jl fixture, m2 with missing
    g = jw.get_genotypes(prefix, 1.0, method="BayesC")
    assert g.nObs == 6 and g.nMarkers == 4 and g.obsID == [f"a{i}" for i in range(1, 7)]
    means = np.fromfile(meta["mean_path"], dtype=np.float32)
    xp = np.fromfile(meta["xp_path"], dtype=np.float32)
    for j in range(4):                                        # decode == dense, xpRinvx == dot(decoded, decoded)
        d = orc.decode_marker(g.packed, 6, j, float(means[j]))
        assert float(d @ d) == pytest.approx(float(xp[j]), abs=1e-5)
    assert means[1] == pytest.approx(0.8)


def test_contract_arithmetic_tracks_reference_arithmetic():
    """The contract sweep (fixed-point dots, binary64 scalars) and the faithful Float32 restatement of
    BayesABC! agree to Float32 rounding on a sweep with shared draws (same pattern as the reference's
    dense-vs-stream check, test_streaming_codec.jl:53-105, atol 1e-4)."""
    n, p = 300, 200
    codes = make_codes(n, p, 13)
    packed = orc.pack_codes(codes)
    means, xpx = orc.marker_stats(packed, n)
    X = orc.dense_centered(packed, n, means)
    y = make_phenotype(X, 13)[0]
    yc = (y - y.mean()).astype(np.float32)
    rng = np.random.default_rng(0)
    ve = np.full(p, 0.02); pi = np.full(p, 0.8)
    y1, a1, b1, d1 = yc.copy(), np.zeros(p, np.float32), np.zeros(p, np.float32), np.zeros(p, np.float32)
    y2, a2, b2, d2 = yc.copy(), np.zeros(p, np.float32), np.zeros(p, np.float32), np.zeros(p, np.int32)
    y3, a3, b3, d3 = yc.copy(), np.zeros(p, np.float32), np.zeros(p, np.float32), np.zeros(p, np.float32)
    for it in range(3):
        u = rng.random(p); z = rng.standard_normal(p)
        orc.bayesabc_ref(X, xpx, y1, a1, b1, d1, 1.0, ve, pi, u, z)
        orc.sweep_contract(packed, n, means, xpx, [0, 64, 128, p], y2, a2, b2, d2, vare=1.0, varEffects=ve, pi=pi, u=u, z=z)
        orc.bayesabc_streaming_ref(packed, n, means, xpx, y3, a3, b3, d3, 1.0, ve, pi, u, z)
        np.testing.assert_array_equal(d1.astype(np.int32), d2)
        np.testing.assert_allclose(a2, a1, atol=1e-4)
        np.testing.assert_allclose(y2, y1, atol=1e-3)
        np.testing.assert_allclose(a3, a1, atol=1e-4)       # dense == stream (reference's own check)
    assert d2.sum() > 5


def test_multitrait_constraint_runs_independent_traits():
    # constraint=true: megaBayesABC! -- no genetic covariance among traits, one pi per trait
    codes, ids, ph = make_data(ntraits=2, seed=8)
    geno = jw.get_genotypes(codes, np.array([[1.0, 0.0], [0.0, 1.0]]), method="BayesC", obsID=ids, constraint=True)
    model = jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", np.array([[1.0, 0.2], [0.2, 1.0]]),
                           genotypes={"geno": geno})
    out = jw.runMCMC(model, ph, chain_length=12, burnin=2, seed=5, outputEBV=False, _backend_factory=factory)
    assert len(out["pi_geno"]) == 2 and out["pi_geno"]["Estimate"].between(0, 1).all()
    assert len(out["marker effects geno"]) == 2 * geno.nMarkers


def test_heritability_output_and_result_files(tmp_path):
    """output_heritability (output.jl:196-209, 498-511): per saved sample the (co)variance of the breeding values over
    all genotyped individuals and h2 = g / (g + vare); result tables written like JWAS.jl:479-482 into a folder that is
    never an existing one (JWAS.jl:255-262)."""
    codes, ids, ph = make_data(n=150, p=120, seed=41, ntraits=2)
    geno = jw.get_genotypes(codes, np.eye(2), method="BayesC", Pi=0.0, obsID=ids)
    model = jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", np.eye(2), genotypes={"geno": geno})
    folder = str(tmp_path / "res")
    os.makedirs(folder)                                  # exists already -> results go to res1
    out = jw.runMCMC(model, ph.iloc[:100], chain_length=40, burnin=10, output_samples_frequency=2, seed=3,
                     output_heritability=True, output_folder=folder, _backend_factory=factory)
    gv, h2 = out["genetic_variance"], out["heritability"]
    assert list(gv["Covariance"]) == ["y1_y1", "y1_y2", "y2_y1", "y2_y2"] and list(h2["Covariance"]) == ["y1", "y2"]
    assert h2["Estimate"].between(0, 1).all() and (gv["Estimate"].to_numpy()[[0, 3]] > 0).all()
    assert gv["Estimate"][1] == pytest.approx(gv["Estimate"][2])
    assert len(out["EBV_y1"]) == 150                     # heritability uses every genotyped individual
    # h2 of the posterior-mean breeding values is in the same ballpark as the mean of the per-sample h2
    rv = out["residual variance"]["Estimate"].to_numpy()[[0, 3]]
    assert np.all(np.abs(h2["Estimate"].to_numpy() - gv["Estimate"].to_numpy()[[0, 3]] / (gv["Estimate"].to_numpy()[[0, 3]] + rv)) < 0.15)
    assert os.listdir(folder) == []
    written = sorted(os.listdir(folder + "1"))
    samples = ["MCMC_samples_" + x + ".txt" for x in ("residual_variance", "marker_effects_variances_geno", "pi_geno", "EBV_y1",
                                                         "EBV_y2", "genetic_variance", "heritability")]
    assert written == sorted([k.replace(" ", "_") + ".txt" for k in out] + samples)
    # sample files (output.jl:318-515): header + one row per saved sample; their means are the tables of the output
    ns = (40 - 10) // 2
    ebv_s = pd.read_csv(os.path.join(folder + "1", "MCMC_samples_EBV_y1.txt"))
    assert list(ebv_s.columns) == ids and len(ebv_s) == ns
    np.testing.assert_allclose(ebv_s.mean(axis=0).to_numpy(), out["EBV_y1"]["EBV"].to_numpy(float), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(ebv_s.var(axis=0, ddof=1).to_numpy(), out["EBV_y1"]["PEV"].to_numpy(float), rtol=1e-5, atol=1e-9)
    h2_s = pd.read_csv(os.path.join(folder + "1", "MCMC_samples_heritability.txt"))
    assert list(h2_s.columns) == ["y1", "y2"] and len(h2_s) == ns
    np.testing.assert_allclose(h2_s.mean(axis=0).to_numpy(), h2["Estimate"].to_numpy(float), rtol=1e-9)
    rv_s = pd.read_csv(os.path.join(folder + "1", "MCMC_samples_residual_variance.txt"))
    assert list(rv_s.columns) == ["y1_y1", "y1_y2", "y2_y1", "y2_y2"] and len(rv_s) == ns
    np.testing.assert_allclose(rv_s.mean(axis=0).to_numpy(), out["residual variance"]["Estimate"].to_numpy(float), rtol=1e-6)
    lines = open(os.path.join(folder + "1", "MCMC_samples_pi_geno.txt")).read().split("\n")
    assert len(lines) == ns * 5 + 1 and lines[4] == ""              # four joint-state probabilities, then a blank line
    assert len(open(os.path.join(folder + "1", "MCMC_samples_marker_effects_variances_geno.txt")).readlines()) == 2 * ns
    back = pd.read_csv(os.path.join(folder + "1", "marker_effects_geno.txt"))
    assert list(back.columns) == ["Trait", "Marker_ID", "Estimate", "SD", "Model_Frequency"] and len(back) == 2 * geno.nMarkers
    # single trait: scalar tables
    codes, ids, ph = make_data(n=90, p=60, seed=42)
    geno = jw.get_genotypes(codes, 1.0, method="BayesC", Pi=0.5, obsID=ids)
    model = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
    out = jw.runMCMC(model, ph, chain_length=20, seed=3, output_heritability=True, _backend_factory=factory)
    assert list(out["genetic_variance"]["Covariance"]) == ["y1"] and 0 < out["heritability"]["Estimate"][0] < 1
    assert not os.path.exists("results")                 # no folder given: nothing is written


def test_uncentred_genotypes(tmp_path):
    """center=false (readgenotypes.jl:384; decode_marker!, streaming_genotypes.jl:993-994): x_ij is the code itself,
    xpRinvx the sum of squared codes, EBV = M * alpha without centring.  Supported when no call is missing (a missing
    call would have to carry the column mean); the reference's own benchmark scripts run this way."""
    codes, ids, ph = make_data(n=140, p=90, seed=51)
    geno = jw.get_genotypes(codes, 1.0, method="BayesC", Pi=0.9, obsID=ids, center=False)
    assert geno.centered is False
    np.testing.assert_allclose(geno.marker_means, codes.mean(axis=0), rtol=1e-6)      # still the allele-frequency source
    model = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
    seen = {}

    def fac(packed, n, t, starts, means=None):
        seen["b"] = factory(packed, n, t, starts, means=means)
        return seen["b"]

    out = jw.runMCMC(model, ph, chain_length=200, burnin=50, seed=4, _backend_factory=fac)
    b = seen["b"]
    assert np.all(b.means == 0.0)
    np.testing.assert_allclose(b.xpx, (codes.astype(np.float64) ** 2).sum(axis=0), rtol=1e-6)
    np.testing.assert_allclose(b.mul_alpha(0), codes @ b.alpha.astype(np.float64), rtol=1e-4, atol=1e-4)
    # same model, centred: the intercept absorbs the column means, the marker effects agree statistically
    g2 = jw.get_genotypes(codes, 1.0, method="BayesC", Pi=0.9, obsID=ids)
    m2 = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": g2})
    out2 = jw.runMCMC(m2, ph, chain_length=200, burnin=50, seed=4, _backend_factory=factory)
    e1 = out["EBV_y1"]["EBV"].to_numpy(float); e2 = out2["EBV_y1"]["EBV"].to_numpy(float)
    assert np.corrcoef(e1, e2)[0, 1] > 0.97
    assert abs((e1 - e1.mean()).std() / (e2 - e2.mean()).std() - 1) < 0.2
    # missing calls cannot be represented uncentred
    codes[3, 5] = 9
    with pytest.raises(jw.JwasError, match="center=false with missing genotypes"):
        jw.get_genotypes(codes, 1.0, obsID=ids, center=False)
    # a prepared backend records its flag and xpRinvx follows it (test_streaming_prepare_lowmem.jl:22-66)
    path = str(tmp_path / "g.csv")
    df = pd.DataFrame(codes.astype(int), columns=[f"m{j + 1}" for j in range(90)]); df.insert(0, "ID", ids)
    df.to_csv(path, index=False)
    prefix = jw.prepare_streaming_genotypes(path, quality_control=False, center=False)
    be = jw.load_streaming_backend(prefix)
    assert be["centered"] is False
    x = np.where(codes == 9, np.nan, codes)
    m = np.nanmean(x, axis=0)
    raw = np.where(np.isnan(x), m, x)
    np.testing.assert_allclose(np.fromfile(prefix + ".xpRinvx.f32", np.float32), (raw * raw).sum(axis=0), rtol=2e-6)


@pytest.mark.parametrize("constraint", [False, True])
def test_multitrait_rrblup_chain(constraint):
    """Multi-trait RR-BLUP: MTBayesC0! / megaBayesC0! (MCMC_BayesianAlphabet.jl:259-268) through the sampler-I and
    megaBayesABC sweeps with every marker in the model for every trait."""
    codes, ids, ph = make_data(n=160, p=120, seed=61, ntraits=2)
    G = np.array([[1.0, 0.0 if constraint else 0.4], [0.0 if constraint else 0.4, 1.0]])
    geno = jw.get_genotypes(codes, G, method="RR-BLUP", obsID=ids, constraint=constraint)
    model = jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", np.eye(2), genotypes={"geno": geno}, constraint=constraint)
    out = jw.runMCMC(model, ph, chain_length=120, burnin=30, seed=8, _backend_factory=factory)
    me = out["marker effects geno"]
    assert (me["Model_Frequency"] == 1.0).all() and "pi_geno" not in out
    for tr in ("y1", "y2"):
        assert np.corrcoef(out["EBV_" + tr]["EBV"].to_numpy(float), ph[tr].to_numpy(float))[0, 1] > 0.6
    g2 = jw.get_genotypes(codes, G, method="RR-BLUP", obsID=ids, multi_trait_sampler="II")
    m2 = jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", np.eye(2), genotypes={"geno": g2})
    with pytest.raises(jw.JwasError, match="supported for BayesC only"):
        jw.runMCMC(m2, ph, chain_length=5, _backend_factory=factory)
