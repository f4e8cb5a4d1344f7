"""Pins the CPU oracle against every RNG-free known answer the reference's own tests hold for
the marker-sweep path.  Reference paths are relative to /root/reference.
"""
import itertools

import numpy as np
import pytest

# test/unit/test_streaming_codec.jl:6-16 -- the inline CSV fixture (rows a1..a6, markers m1..m4)
CODEC_ROWS = np.array([[0, 1, 2, 0],
                       [1, 0, 1, 2],
                       [2, 1, 0, 1],
                       [0, 2, 1, 0],
                       [1, 1, 2, 2],
                       [2, 0, 0, 1]])


def dense_centered_like_reference(codes, missing=9):
    """markers/readgenotypes.jl:372-385: missing -> column mean, then center! (Float32)."""
    X = codes.astype(np.float32)
    out = np.empty_like(X)
    for j in range(X.shape[1]):
        col = X[:, j].copy()
        miss = col == missing
        mu = np.float32(col[~miss].sum(dtype=np.float32) / np.float32((~miss).sum()))
        col[miss] = mu
        out[:, j] = col - col.mean(dtype=np.float32)
    return out


def test_codec_bit_layout(oracle):
    # streaming_genotypes.jl:364-367 / 622-627: LSB-first, 4 individuals per byte, code 3 = missing
    codes = CODEC_ROWS.copy()
    codes[2, 1] = 9  # with_missing=true replaces a3/m2
    packed = oracle.pack_codes(codes)
    assert packed.shape == (4, 2)
    assert packed[0].tolist() == [0x24, 0x09]          # m1 = 0,1,2,0 | 1,2
    assert packed[1].tolist() == [0xB1, 0x01]          # m2 = 1,0,3,2 | 1,0
    # every code readable back
    for j in range(4):
        for i in range(6):
            c = (packed[j, i >> 2] >> ((i & 3) << 1)) & 3
            assert c == (3 if codes[i, j] == 9 else codes[i, j])


@pytest.mark.parametrize("with_missing", [False, True])
def test_decode_matches_dense_column(oracle, with_missing):
    # test_streaming_codec.jl:46-50: decode_marker! == dense centred column, atol 1e-5
    codes = CODEC_ROWS.copy()
    if with_missing:
        codes[2, 1] = 9
    packed = oracle.pack_codes(codes)
    means, xpx, af = oracle.marker_stats_ref(packed, 6, center=True)
    dense = dense_centered_like_reference(codes)
    for j in range(4):
        got = oracle.decode_marker(packed, 6, j, float(means[j]), centered=True)
        np.testing.assert_allclose(got, dense[:, j], atol=1e-5)
    if with_missing:
        assert means[1] == pytest.approx(0.8)
        np.testing.assert_allclose(oracle.decode_marker(packed, 6, 1, float(means[1])),
                                   [0.2, -0.8, 0.0, 1.2, 0.2, -0.8], atol=1e-6)
    np.testing.assert_allclose(af, means / 2)


@pytest.mark.parametrize("center", [True, False])
def test_xprinvx_parity(oracle, center):
    # test_streaming_prepare_lowmem.jl:49-50, 64-66: dot(decoded, decoded) == xpRinvx, atol 1e-5
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 3, size=(37, 11))
    codes[rng.random(codes.shape) < 0.08] = 9
    packed = oracle.pack_codes(codes)
    means, xpx, _ = oracle.marker_stats_ref(packed, 37, center=center)
    for j in range(11):
        d = oracle.decode_marker(packed, 37, j, float(means[j]), centered=center)
        assert float(np.dot(d, d)) == pytest.approx(float(xpx[j]), abs=1e-4 if not center else 1e-5)
    if center:
        # contract statistics agree with the reference-arithmetic ones to Float32 rounding
        m2, x2 = oracle.marker_stats(packed, 37)
        np.testing.assert_array_equal(m2, means)
        np.testing.assert_allclose(x2, xpx, rtol=2e-6)
        # Gram diagonal is xpx; off-diagonals match the decoded columns
        G = oracle.gram_block(packed, 37, means, 2, 5)
        X = oracle.dense_centered(packed, 37, means).astype(np.float64)
        np.testing.assert_allclose(G, (X[:, 2:7].T @ X[:, 2:7]), rtol=1e-6, atol=1e-5)
        np.testing.assert_array_equal(np.diag(G), x2[2:7])


def test_mul_alpha(oracle):
    # streaming_genotypes.jl:1009-1027
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 3, size=(23, 9))
    packed = oracle.pack_codes(codes)
    means, _ = oracle.marker_stats(packed, 23)
    alpha = rng.normal(size=9).astype(np.float32); alpha[[1, 4]] = 0
    X = oracle.dense_centered(packed, 23, means)
    np.testing.assert_allclose(oracle.mul_alpha(packed, 23, means, alpha), X @ alpha, rtol=1e-5, atol=1e-5)


def test_bayesr_block_nreps_table(oracle):
    # test/unit/test_bayesr.jl:244-250
    assert oracle.bayesr_block_nreps(1, 10, 7) == 1
    assert oracle.bayesr_block_nreps(10, 10, 7) == 1
    assert oracle.bayesr_block_nreps(11, 10, 7) == 7
    assert oracle.bayesr_block_nreps(25, 0, 7) == 7
    assert oracle.bayesr_block_nreps(3, 8, 1) == 1
    assert oracle.bayesr_block_nreps(3, 8, 0) < 0      # "block_size must be at least 1"


def test_bayesr_sigma_sufficient_statistics(oracle):
    # test/unit/test_bayesr.jl:252-262
    alpha = np.array([0.0, 0.4, -0.3, 0.1, 0.0])
    delta = np.array([1, 2, 4, 3, 1])
    gamma = np.array([0.0, 0.01, 0.1, 1.0])
    ssq, nnz = oracle.bayesr_sigma_sufficient_statistics(alpha, delta, gamma)
    a32 = alpha.astype(np.float32).astype(np.float64)
    expected = a32[1] ** 2 / gamma[1] + a32[2] ** 2 / gamma[3] + a32[3] ** 2 / gamma[2]
    assert ssq == pytest.approx(expected, rel=1e-12)
    assert nnz == 3


def test_block_start_validation(oracle):
    # test/unit/test_misc_coverage.jl:195-208 with demo_7animals (5 markers); JWAS.jl:73-79
    assert oracle.validate_block_starts([1, 3, 5], 5) == 0
    for bad in ([2, 4], [1, 3, 3], [3, 1], [1, 10], []):
        assert oracle.validate_block_starts(bad, 5) != 0


def test_independent_block_crossproduct_identity():
    # test/unit/test_misc_coverage.jl:211-227
    y = np.array([1.2, -0.3, 0.7, 1.4]); W = np.diag([1.0, 2.0, 3.0, 4.0])
    x1 = np.array([1.0, 0, 0, 0]); x2i = np.array([0, 0, 1.0, 0]); x2c = np.array([1.0, 0, 1.0, 0])
    a1, a2 = 0.4, -0.6
    assert x1 @ W @ x2i == 0.0
    assert x1 @ W @ (y - x1 * a1 - x2i * a2) == pytest.approx(x1 @ W @ (y - x1 * a1))
    assert x1 @ W @ x2c != 0.0
    assert x1 @ W @ (y - x1 * a1 - x2c * a2) != pytest.approx(x1 @ W @ (y - x1 * a1))


# ----------------------------------------------------------------------------------------------
# closed-form multi-trait state probabilities: test/unit/test_multitrait_mcmc.jl:6-30, 557-642
# ----------------------------------------------------------------------------------------------
MT_X = np.array([1.0, -0.5, 0.75])
MT_Y = np.array([[0.8, -0.1, 0.3], [0.2, 0.6, -0.4]])
MT_R = np.array([[1.0, 0.25], [0.25, 0.9]])
MT_G = np.array([[0.7, 0.15], [0.15, 0.8]])
MT_PI = np.array([0.35, 0.20, 0.15, 0.30])      # states (0,0),(1,0),(0,1),(1,1)


def exact_mt_state_probs(x, Y, R, G, Pi):
    xp = x @ x
    Rinv, Ginv = np.linalg.inv(R), np.linalg.inv(G)
    w = Y @ x
    ld = np.zeros(4)
    for s in range(4):
        D = np.diag([s & 1, (s >> 1) & 1]).astype(float)
        lhs = D @ Rinv @ D * xp + Ginv
        rhs = (Rinv @ D).T @ w
        ghat = np.linalg.solve(lhs, rhs)
        ld[s] = -0.5 * (np.log(np.linalg.det(lhs)) - rhs @ ghat) + np.log(Pi[s])
    pr = np.exp(ld - ld.max())
    return pr / pr.sum()


@pytest.mark.parametrize("mode", ["I", "II"])
def test_mt_sampler_targets_closed_form(oracle, mode):
    exact = exact_mt_state_probs(MT_X, MT_Y, MT_R, MT_G, MT_PI)
    golden = np.loadtxt(__file__.replace("test_oracle_pins.py", "golden/mt_state_probs.txt"))
    np.testing.assert_allclose(exact, golden, rtol=1e-10)
    rng = np.random.default_rng(20260410)
    X = np.asfortranarray(MT_X.astype(np.float32).reshape(3, 1))
    xpx = np.array([MT_X @ MT_X], np.float32)
    ycorr = MT_Y.astype(np.float32).reshape(-1).copy()
    alpha = np.zeros((2, 1), np.float32); beta = np.zeros((2, 1), np.float32); delta = np.zeros((2, 1), np.float32)
    niter, burn = 20000, 3000
    counts = np.zeros(4)
    for it in range(niter):
        if mode == "I":
            oracle.mtbayesabc_I_ref(X, xpx, ycorr, alpha, beta, delta, MT_R, MT_G, MT_PI,
                                    rng.random(2), rng.standard_normal(2))
        else:
            oracle.mtbayesabc_II_ref(X, xpx, ycorr, alpha, beta, delta, MT_R, MT_G, MT_PI,
                                     rng.random(1), rng.standard_normal(2))
        if it >= burn:
            counts[int(delta[0, 0]) + 2 * int(delta[1, 0])] += 1
    emp = counts / counts.sum()
    assert np.abs(emp - exact).max() < 0.02


@pytest.mark.parametrize("method_name", ["MT1", "MT2"])
def test_mt_contract_sampler_targets_closed_form(oracle, method_name):
    """Same closed form, through the contract-arithmetic sweeps (sampler I and sampler II) on a packed
    one-marker problem."""
    codes = np.array([[0], [1], [2], [1], [0], [2], [1]])
    n = 7
    packed = oracle.pack_codes(codes)
    means, xpx = oracle.marker_stats(packed, n)
    x = oracle.decode_marker(packed, n, 0, float(means[0])).astype(np.float64)
    Y = np.array([[0.8, -0.1, 0.3, 0.5, -0.7, 0.2, 0.1], [0.2, 0.6, -0.4, -0.3, 0.1, 0.9, -0.2]])
    exact = exact_mt_state_probs(x, Y, MT_R, MT_G, MT_PI)
    ycorr = Y.astype(np.float32).reshape(-1).copy()
    alpha = np.zeros(2, np.float32); beta = np.zeros(2, np.float32); delta = np.zeros(2, np.int32)
    counts = np.zeros(4)
    for it in range(1, 20001):
        rc, _ = oracle.sweep_contract(packed, n, means, xpx, [0, 1], ycorr, alpha, beta, delta,
                                      method=getattr(oracle, "METHOD_" + method_name), R=MT_R, G=MT_G, bigPi=MT_PI,
                                      seed=77, it=it)
        assert rc == 0
        if it > 3000:
            counts[delta[0] + 2 * delta[1]] += 1
    emp = counts / counts.sum()
    assert np.abs(emp - exact).max() < 0.02
