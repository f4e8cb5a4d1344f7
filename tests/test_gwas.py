"""GWAS post-processing (src/3.GWAS/src/GWAS.jl) on marker-effect samples written by runMCMC: the structural
checks of the reference's own tests (test/unit/test_gwas_windows.jl, runtests.jl:324-350) plus an exact
comparison of the sparse window statistics with the dense computation the reference performs."""
import numpy as np
import pandas as pd
import pytest

import jwas_b200 as jw
from oracle_backend import factory
from oracle import pyoracle as orc
from test_api_chain import make_data


@pytest.fixture(scope="module")
def run(tmp_path_factory):
    out_dir = tmp_path_factory.mktemp("gwas")
    n, p = 90, 60
    codes, ids, ph = make_data(n=n, p=p, seed=12, missing=0.02)
    geno = jw.get_genotypes(codes, 1.0, method="BayesC", Pi=0.8, obsID=ids, quality_control=False)
    model = jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
    jw.runMCMC(model, ph, chain_length=60, burnin=10, output_samples_frequency=5, seed=123,
               output_folder=str(out_dir), output_marker_effect_samples=True, _backend_factory=factory)
    # map: 3 chromosomes, 20 markers each, 0.25 Mb apart -> 1 Mb windows of 4 markers (the last one shorter)
    mp = pd.DataFrame({"markerID": geno.markerID, "chromosome": np.repeat(["1", "2", "3"], 20),
                       "position": np.tile(np.arange(20) * 250_000 + 1_000, 3)})
    map_file = out_dir / "map.txt"
    mp.to_csv(map_file, index=False)
    return model, geno, str(map_file), model.sample_files["y1"], codes


def test_sample_file_format(run):
    model, geno, map_file, sfile, codes = run
    lines = open(sfile).read().strip().split("\n")
    assert lines[0].split(",") == [str(m) for m in geno.markerID]          # output.jl:411
    assert len(lines) - 1 == 10                                             # (60 - 10) / 5 saved iterations
    assert all(len(l.split(",")) == 60 for l in lines[1:])


def test_model_frequency(run):
    model, geno, map_file, sfile, codes = run
    mf = jw.GWAS(sfile)                                                     # runtests.jl:340-346
    assert list(mf.columns) == ["marker_ID", "modelfrequency"]
    assert len(mf) == 60 and mf["modelfrequency"].between(0, 1).all()
    samples = np.loadtxt(sfile, delimiter=",", skiprows=1)
    np.testing.assert_allclose(mf["modelfrequency"], (samples != 0).mean(axis=0))


@pytest.mark.parametrize("sliding", [False, True])
def test_window_statistics_match_dense_computation(run, sliding):
    model, geno, map_file, sfile, codes = run
    (tab,), (props,) = jw.GWAS(model, map_file, sfile, window_size="1 Mb", sliding_window=sliding,
                              threshold=0.01, output_winVarProps=True)
    for col in ("trait", "window", "chr", "wStart", "wEnd", "start_SNP", "end_SNP", "numSNP", "estimateGenVar",
                "stdGenVar", "prGenVar", "WPPA", "PPA_t"):                  # GWAS.jl:182-194
        assert col in tab.columns
    assert tab["WPPA"].between(0, 1).all() and (np.diff(tab["WPPA"]) <= 1e-15).all()    # sorted by WPPA, descending
    # dense restatement of GWAS.jl:141-176 with the centred genotype matrix
    means, _ = orc.marker_stats(geno.packed, geno.nObs)
    X = orc.dense_centered(geno.packed, geno.nObs, means).astype(np.float64)
    samples = np.loadtxt(sfile, delimiter=",", skiprows=1)
    pos = np.tile(np.arange(20) * 250_000 + 1_000, 3); chrom = np.repeat([0, 1, 2], 20)
    # column ranges exactly as GWAS.jl:94-137 computes them.  NB: with sliding windows the reference advances the
    # column cursor by ONE per window and carries it across chromosomes, so from the second chromosome on its
    # column ranges no longer start at that chromosome's first marker; parity means reproducing that.
    wins = []
    index_start = 0
    for c in range(3):
        pc = pos[chrom == c]
        nwin = int(np.ceil(pc[-1] / 1e6)) if not sliding else int(np.argmax(pc >= pc[-1] - 1_000_000)) + 1
        for j in range(nwin):
            start = j * 1_000_000 if not sliding else pc[j]
            k = int(((pc >= start) & (pc < start + 1_000_000)).sum())
            if k:
                wins.append(np.arange(index_start, index_start + k))
            index_start += k if not sliding else 1
    assert len(wins) == len(tab) == props.shape[1]
    dense = np.zeros((len(samples), len(wins)))
    for i, a in enumerate(samples):
        gv = np.var(X @ a, ddof=1)
        for w, idx in enumerate(wins):
            dense[i, w] = np.var(X[:, idx] @ a[idx], ddof=1) / gv if gv > 0 else 0.0
    dense[np.isnan(dense)] = 0.0
    np.testing.assert_allclose(props, dense, rtol=1e-9, atol=1e-12)
    by_window = tab.sort_values("window")
    np.testing.assert_allclose(by_window["WPPA"], (dense > 0.01).mean(axis=0))
    assert list(by_window["numSNP"]) == [len(w) for w in wins]


def test_fake_map_and_errors(run):
    model, geno, map_file, sfile, codes = run
    (tab,) = jw.GWAS(model, False, sfile, window_size=10)                   # GWAS.jl:67-76: 10 markers per window
    assert len(tab) == 6 and (tab["numSNP"] == 10).all()
    with pytest.raises(jw.JwasError, match="window_size"):
        jw.GWAS(model, map_file, sfile, window_size="1 kb")
