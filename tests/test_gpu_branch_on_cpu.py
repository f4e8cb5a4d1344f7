"""The branch of api.runMCMC that builds GpuSweeper / mcmc.GpuBackend objects (no `_backend_factory`) only runs for real
on a B200.  Here it runs against tests/fake_sweeper.py -- the same method surface over the CPU oracle -- so that its
option / call order and argument shapes are exercised by the CPU suite, and its output must equal the oracle-factory
branch's (same host code, same sweeps)."""
import numpy as np
import pandas as pd
import pytest

import jwas_b200 as jw
from fake_sweeper import FakeSweeper
from oracle_backend import factory
from test_annotations import _annotated_data
from test_api_chain import make_data


@pytest.fixture
def fake(monkeypatch):
    monkeypatch.setattr(jw.api, "GpuSweeper", FakeSweeper)
    FakeSweeper.log.clear()
    return FakeSweeper


def same_output(a, b):
    assert a.keys() == b.keys()
    for key in a:
        for col in a[key].columns:
            x, y = a[key][col].to_numpy(), b[key][col].to_numpy()
            if x.dtype.kind == "f":
                np.testing.assert_array_equal(x, y, err_msg=f"{key}/{col}")
            else:
                assert list(x) == list(y), f"{key}/{col}"


def both(make_model, ph, **kw):
    outs = []
    for bf in (None, factory):
        outs.append(jw.runMCMC(make_model(), ph, seed=77, _backend_factory=bf, **kw))
    same_output(*outs)
    return outs[0]


def test_phenotyped_subset_second_handle(fake):
    codes, ids, ph = make_data(n=130, p=100, seed=29, missing=0.01)
    sub = ph.iloc[::-1].iloc[:90].reset_index(drop=True)

    def mk():
        geno = jw.get_genotypes(codes, 1.0, method="BayesC", Pi=0.9, obsID=ids)
        return jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
    out = both(mk, sub, chain_length=20, burnin=4, output_heritability=True)
    assert list(out["EBV_y1"]["ID"]) == ids
    calls = [c[0] for c in fake.log]
    # training handle: means of the full sample before the blocks; EBV handle: created, centred, never given blocks
    assert calls == ["create", "set_marker_means", "set_blocks", "create", "set_marker_means"]
    assert fake.log[0][1] == (100, 90, 1) and fake.log[3][1] == (100, 130, 1)


@pytest.mark.parametrize("constraint", [False, True])
def test_multitrait_rrblup(fake, constraint):
    codes, ids, ph = make_data(n=120, p=90, seed=63, ntraits=2)
    G = np.array([[1.0, 0.0 if constraint else 0.4], [0.0 if constraint else 0.4, 1.0]])

    def mk():
        geno = jw.get_genotypes(codes, G, method="RR-BLUP", obsID=ids, constraint=constraint)
        return jw.build_model("y1 = intercept + geno\ny2 = intercept + geno", np.eye(2), genotypes={"geno": geno},
                              constraint=constraint)
    out = both(mk, ph, chain_length=16, burnin=4)
    assert (out["marker effects geno"]["Model_Frequency"] == 1.0).all()


@pytest.mark.parametrize("case", ["BayesC", "BayesR", "BayesC2"])
@pytest.mark.parametrize("center", [True, False])
def test_annotated_and_uncentred(fake, case, center):
    two = case == "BayesC2"
    codes, ids, ph, A = _annotated_data(n=110, p=130, seed=51, ntraits=2 if two else 1, nqtl=10)
    eqs = "y1 = intercept + geno\ny2 = intercept + geno" if two else "y1 = intercept + geno"
    Pi = {(0.0, 0.0): 0.45, (1.0, 0.0): 0.20, (0.0, 1.0): 0.15, (1.0, 1.0): 0.20} if two else (0.9 if case == "BayesC" else 0.0)

    def mk():
        geno = jw.get_genotypes(codes, False, method="BayesR" if case == "BayesR" else "BayesC", Pi=Pi, annotations=A,
                                obsID=ids, center=center)
        return jw.build_model(eqs, False, genotypes={"geno": geno})
    out = both(mk, ph, chain_length=20, burnin=4)
    assert "annotation coefficients geno" in out
    if not center:
        assert ("set_marker_means", None) in fake.log          # means of zero reach the handle before its blocks


@pytest.mark.parametrize("method,Pi", [("BayesB", 0.8), ("BayesA", 0.0), ("BayesL", 0.0), ("RR-BLUP", 0.0)])
def test_other_single_trait_methods(fake, method, Pi):
    codes, ids, ph = make_data(n=100, p=80, seed=21)

    def mk():
        geno = jw.get_genotypes(codes, 1.0, method=method, Pi=Pi, obsID=ids)
        return jw.build_model("y1 = intercept + geno", 1.0, genotypes={"geno": geno})
    both(mk, ph, chain_length=15, burnin=3, fast_blocks=False)
