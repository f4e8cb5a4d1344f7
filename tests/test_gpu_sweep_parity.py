"""GPU parity: the CUDA sweep (through the C ABI) against the CPU oracle's contract-arithmetic
sweep on the same seeded inputs.  Bar: inclusion indicators, effects and ycorr BIT-EXACT
(north_star asks for indicators bit-exact and effects within 1e-5 relative; the arithmetic
contract gives equality, so that is what is asserted)."""
import numpy as np
import pytest

from helpers import Problem, uniform_starts, canonical_sum_prod

pytestmark = pytest.mark.gpu

GAMMA = np.array([0.0, 0.01, 0.1, 1.0])          # JWAS.jl:12 BAYESR_GAMMA
PI_R = np.array([0.95, 0.03, 0.015, 0.005])      # tools4genotypes.jl:373-375


@pytest.fixture(scope="module")
def jw():
    import jwas_b200
    assert jwas_b200.device_count() > 0, "no CUDA device: the gpu-marked tests need a B200"
    return jwas_b200


def run_pair_abc(jw, oracle, prob, starts, schedule, nsweeps, *, pi=0.9, bayesb=False, replay=False, engine=0,
                 seed=11, lag=0, chain_ctas=0, gather=1, ws=0):
    n, p = prob.n, prob.p
    g = jw.GpuSweeper(prob.packed, n, 1)
    g.set_blocks(starts)
    g.set_option("engine", engine)
    g.set_option("lag", lag)
    g.set_option("chain_ctas", chain_ctas)
    g.set_option("gather", gather)
    g.set_option("ws", ws)
    gm, gx = g.marker_stats()
    np.testing.assert_array_equal(gm, prob.means)
    np.testing.assert_array_equal(gx, prob.xpx)
    yc, al, be, de = prob.fresh_state()
    g.put_ycorr(yc); g.put_state(al, be, de)
    rng = np.random.default_rng(seed)
    vare = prob.vary * 0.5
    ve = np.full(p, prob.vary * 0.5 / (0.1 * prob.xpx.mean() / n * p))
    if bayesb:
        ve = ve * rng.uniform(0.5, 2.0, size=p)
    piv = np.full(p, pi)
    nreps_mode = 0 if schedule == jw.SCHED_EXACT else 1
    maxb = int(np.diff(starts).max())
    for it in range(1, nsweeps + 1):
        u = z = None
        if replay:
            reps = maxb if nreps_mode else 1
            u = rng.random(reps * p); z = rng.standard_normal(reps * p)
        rc, S = oracle.sweep_contract(prob.packed, n, prob.means, prob.xpx, starts, yc, al, be, de,
                                      method=oracle.METHOD_ABC, nreps_mode=nreps_mode,
                                      independent=(schedule == jw.SCHED_INDEPENDENT), vare=vare,
                                      varEffects=ve, pi=piv, seed=seed, it=it, u=u, z=z, lag=lag)
        assert rc == 0
        st = g.sweep_bayesabc(schedule, vare, ve, piv, seed, it, u, z)
        ga, gb, gd = g.get_state()
        gy = g.get_ycorr()
        assert st.scale_exp == S
        np.testing.assert_array_equal(gd, de, err_msg=f"delta differs at sweep {it}")
        np.testing.assert_array_equal(ga.view(np.uint32), al.view(np.uint32), err_msg=f"alpha differs at sweep {it}")
        np.testing.assert_array_equal(gb.view(np.uint32), be.view(np.uint32), err_msg=f"beta differs at sweep {it}")
        np.testing.assert_array_equal(gy.view(np.uint32), yc.view(np.uint32), err_msg=f"ycorr differs at sweep {it}")
        assert st.sum_delta[0] == de.sum()
        assert st.nnz_alpha[0] == np.count_nonzero(al)
        assert st.alpha_ss[0] == canonical_sum_prod(al, al)
        assert st.ycorr_ss[0] == canonical_sum_prod(yc, yc)
        assert st.ycorr_maxabs == np.abs(yc).max()
    assert de.sum() > 0, "degenerate test: nothing ever entered the model"
    g.close()
    return st


@pytest.mark.parametrize("missing", [0.0, 0.03])
@pytest.mark.parametrize("n,p,b", [(500, 2000, 256), (501, 333, 64), (67, 50, 1), (1030, 700, 700)])
def test_bayesc_exact_schedule(jw, oracle, n, p, b, missing):
    prob = Problem(oracle, n, p, seed=n + p, missing=missing)
    run_pair_abc(jw, oracle, prob, uniform_starts(p, b), jw.SCHED_EXACT, nsweeps=4)


def test_bayesb_per_marker_variances_and_replayed_draws(jw, oracle):
    prob = Problem(oracle, 300, 400, seed=5)
    run_pair_abc(jw, oracle, prob, uniform_starts(400, 128), jw.SCHED_EXACT, nsweeps=3, bayesb=True, replay=True)


def test_bayesa_pi_zero(jw, oracle):
    # BayesA is BayesB with pi = 0 (input_data_validation.jl:33-36): every marker is included
    prob = Problem(oracle, 200, 60, seed=8)
    st = run_pair_abc(jw, oracle, prob, uniform_starts(60, 16), jw.SCHED_EXACT, nsweeps=2, pi=0.0)
    assert st.sum_delta[0] == 60


@pytest.mark.parametrize("schedule_name", ["block", "independent"])
@pytest.mark.parametrize("missing", [0.0, 0.02])
def test_bayesc_block_schedules(jw, oracle, schedule_name, missing):
    # BayesABC_block! nreps = block size (BayesABC.jl:153); explicit ragged starts
    prob = Problem(oracle, 400, 157, seed=21, missing=missing)
    starts = np.array([0, 20, 21, 60, 100, 157], dtype=np.int64)
    sched = jw.SCHED_BLOCK if schedule_name == "block" else jw.SCHED_INDEPENDENT
    run_pair_abc(jw, oracle, prob, starts, sched, nsweeps=3, replay=(schedule_name == "block"))


def run_pair_r(jw, oracle, prob, starts, schedule, full_reps, nsweeps, seed=3, engine=0, lag=0, chain_ctas=0, gather=1, ws=0):
    n, p = prob.n, prob.p
    g = jw.GpuSweeper(prob.packed, n, 1)
    g.set_blocks(starts)
    g.set_option("engine", engine)
    g.set_option("lag", lag)
    g.set_option("chain_ctas", chain_ctas)
    g.set_option("gather", gather)
    g.set_option("ws", ws)
    yc, al, be, de = prob.fresh_state()
    de[:] = 1
    g.put_ycorr(yc); g.put_state(al, be, de)
    vare = prob.vary * 0.5
    sigma = prob.vary * 0.5 / (prob.xpx.mean() / n * p * float(GAMMA @ PI_R))
    nreps_mode = 0 if (schedule == jw.SCHED_EXACT or not full_reps) else 1
    for it in range(1, nsweeps + 1):
        rc, S = oracle.sweep_contract(prob.packed, n, prob.means, prob.xpx, starts, yc, al, None, de,
                                      method=oracle.METHOD_R, nreps_mode=nreps_mode,
                                      independent=(schedule == jw.SCHED_INDEPENDENT), vare=vare,
                                      sigmaSq=sigma, pi=PI_R, gamma=GAMMA, seed=seed, it=it, lag=lag)
        assert rc == 0
        st = g.sweep_bayesr(schedule, full_reps, vare, sigma, PI_R, GAMMA, seed, it)
        ga, _, gd = g.get_state()
        gy = g.get_ycorr()
        np.testing.assert_array_equal(gd, de)
        np.testing.assert_array_equal(ga.view(np.uint32), al.view(np.uint32))
        np.testing.assert_array_equal(gy.view(np.uint32), yc.view(np.uint32))
        counts = np.bincount(de, minlength=5)[1:5]
        assert [st.class_counts[k] for k in range(4)] == counts.tolist()
        ssq, nnz = oracle.bayesr_sigma_sufficient_statistics(al, de, GAMMA)
        assert st.sum_delta[0] == nnz
        assert st.bayesr_ssq == pytest.approx(ssq, rel=1e-12)
    assert (de > 1).sum() > 0
    g.close()


@pytest.mark.parametrize("missing", [0.0, 0.03])
def test_bayesr_exact(jw, oracle, missing):
    prob = Problem(oracle, 500, 1000, seed=31, missing=missing)
    run_pair_r(jw, oracle, prob, uniform_starts(1000, 256), jw.SCHED_EXACT, 1, nsweeps=4)


@pytest.mark.parametrize("schedule_name,full_reps", [("block", 1), ("block", 0), ("independent", 1)])
def test_bayesr_block_schedules(jw, oracle, schedule_name, full_reps):
    # bayesr_block_nreps burn-in gate (BayesR.jl:22-25): full_reps=0 during burn-in
    prob = Problem(oracle, 300, 90, seed=33)
    sched = jw.SCHED_BLOCK if schedule_name == "block" else jw.SCHED_INDEPENDENT
    run_pair_r(jw, oracle, prob, uniform_starts(90, 17), sched, full_reps, nsweeps=3)


def run_pair_mt(jw, oracle, prob, starts, schedule, nsweeps, seed=9, engine=0, lag=0, sampler="I", chain_ctas=0, gather=1):
    n, p, t = prob.n, prob.p, prob.t
    g = jw.GpuSweeper(prob.packed, n, t)
    g.set_blocks(starts)
    g.set_option("engine", engine)
    g.set_option("lag", lag)
    g.set_option("chain_ctas", chain_ctas)
    g.set_option("gather", gather)
    yc, al, be, de = prob.fresh_state()
    g.put_ycorr(yc); g.put_state(al, be, de)
    R = np.array([[1.0, 0.3], [0.3, 1.2]]) * prob.vary * 0.5
    G = np.array([[1.0, 0.4], [0.4, 0.8]]) * prob.vary * 0.5 / (0.2 * prob.xpx.mean() / n * p)
    bigPi = np.array([0.7, 0.1, 0.1, 0.1])
    nreps_mode = 0 if schedule == jw.SCHED_EXACT else 1
    for it in range(1, nsweeps + 1):
        rc, S = oracle.sweep_contract(prob.packed, n, prob.means, prob.xpx, starts, yc, al, be, de,
                                      method=(oracle.METHOD_MT2 if sampler == "II" else oracle.METHOD_MT1),
                                      nreps_mode=nreps_mode,
                                      independent=(schedule == jw.SCHED_INDEPENDENT), R=R, G=G, bigPi=bigPi,
                                      seed=seed, it=it, lag=lag)
        assert rc == 0
        st = (g.sweep_mt2 if sampler == "II" else g.sweep_mt1)(schedule, R, G, bigPi, seed, it)
        ga, gb, gd = g.get_state()
        gy = g.get_ycorr()
        np.testing.assert_array_equal(gd, de)
        np.testing.assert_array_equal(ga.view(np.uint32), al.view(np.uint32))
        np.testing.assert_array_equal(gb.view(np.uint32), be.view(np.uint32))
        np.testing.assert_array_equal(gy.view(np.uint32), yc.view(np.uint32))
        states = de[:p] + 2 * de[p:]
        assert [st.class_counts[k] for k in range(4)] == np.bincount(states, minlength=4).tolist()
        assert st.beta_ss[1] == canonical_sum_prod(be[:p], be[p:])
        assert st.ycorr_ss[1] == canonical_sum_prod(yc[:n], yc[n:])
    assert de.sum() > 0
    g.close()


@pytest.mark.parametrize("missing", [0.0, 0.03])
def test_mt_sampler1_exact(jw, oracle, missing):
    prob = Problem(oracle, 403, 600, seed=41, missing=missing, ntraits=2)
    run_pair_mt(jw, oracle, prob, uniform_starts(600, 200), jw.SCHED_EXACT, nsweeps=3)


@pytest.mark.parametrize("schedule_name", ["block", "independent"])
def test_mt_sampler1_block_schedules(jw, oracle, schedule_name):
    prob = Problem(oracle, 200, 70, seed=43, ntraits=2)
    sched = jw.SCHED_BLOCK if schedule_name == "block" else jw.SCHED_INDEPENDENT
    run_pair_mt(jw, oracle, prob, uniform_starts(70, 16), sched, nsweeps=2)


def test_ycorr_lifecycle_and_accumulators(jw, oracle):
    """ycorr init (MCMC_BayesianAlphabet.jl:131-147), intercept shift (:207-220), getEBV product
    (output.jl:302) and the running posterior means (output.jl:568-577)."""
    prob = Problem(oracle, 333, 210, seed=51, missing=0.02)
    n, p = prob.n, prob.p
    g = jw.GpuSweeper(prob.packed, n, 1)
    g.set_blocks(uniform_starts(p, 64))
    rng = np.random.default_rng(1)
    alpha = np.where(rng.random(p) < 0.2, rng.normal(size=p), 0.0).astype(np.float32)
    y = prob.ycorr0.copy()
    g.put_ycorr(y); g.put_state(alpha, np.zeros(p, np.float32), (alpha != 0).astype(np.int32))
    ebv = g.mul_alpha(0)
    np.testing.assert_allclose(ebv, oracle.mul_alpha(prob.packed, n, prob.means, alpha), rtol=1e-5, atol=1e-5)
    g.ycorr_sub_malpha()
    np.testing.assert_allclose(g.get_ycorr(), y - ebv, rtol=1e-5, atol=1e-5)
    s, ss = g.shift_ycorr(0, 0.25)
    yy = g.get_ycorr()
    assert ss == canonical_sum_prod(yy, yy)
    assert s == canonical_sum_prod(yy, np.ones_like(yy))
    # accumulators
    ma = np.zeros(p, np.float32); ma2 = np.zeros(p, np.float32); md = np.zeros(p, np.float32)
    for k in range(1, 4):
        a = (alpha * k).astype(np.float32); d = (rng.random(p) < 0.5).astype(np.int32)
        g.put_state(a, None, d)
        g.accumulate(k)
        ma = (ma.astype(np.float64) + (a.astype(np.float64) - ma) / k).astype(np.float32)
        ma2 = (ma2.astype(np.float64) + (a.astype(np.float64) ** 2 - ma2) / k).astype(np.float32)
        md = (md.astype(np.float64) + (d - md.astype(np.float64)) / k).astype(np.float32)
    gma, gma2, gmd = g.get_means()
    np.testing.assert_array_equal(gma, ma); np.testing.assert_array_equal(gma2, ma2); np.testing.assert_array_equal(gmd, md)
    g.close()


def test_error_behaviour(jw, oracle):
    """error("...") -> ErrorException in the reference; non-zero rc + message here."""
    prob = Problem(oracle, 40, 10, seed=61)
    g = jw.GpuSweeper(prob.packed, 40, 1)
    with pytest.raises(jw.JwasError, match="jwas_set_blocks must be called"):
        g.sweep_bayesc(jw.SCHED_EXACT, 1.0, 1.0, 0.9, 1, 1)
    for bad in ([1, 5, 10], [0, 5, 5, 10], [0, 7, 3, 10], [0, 5, 12]):   # test_misc_coverage.jl:195-208
        with pytest.raises(jw.JwasError, match="fast_blocks"):
            g.set_blocks(np.array(bad, dtype=np.int64))
    g.set_blocks(np.array([0, 5, 10], dtype=np.int64))
    with pytest.raises(jw.JwasError, match="BayesR pi must sum to 1"):
        g.sweep_bayesr(jw.SCHED_EXACT, 1, 1.0, 1.0, np.array([0.5, 0.2, 0.2, 0.2]), GAMMA, 1, 1)
    with pytest.raises(jw.JwasError, match="sigmaSq must be positive"):
        g.sweep_bayesr(jw.SCHED_EXACT, 1, 1.0, 0.0, PI_R, GAMMA, 1, 1)
    g.close()
    with pytest.raises(jw.JwasError, match="Genotype data is empty"):
        jw.GpuSweeper(np.zeros((0, 1), np.uint8), 0, 1)


def test_full_size_invariants(jw, oracle):
    """At a size the oracle cannot run in seconds, check size-independent properties: after K
    sweeps ycorr + M*alpha reproduces the input (linearity of the fused updates), counts and
    sums agree with the state that comes back, and the same seed reproduces the same bits."""
    rng = np.random.default_rng(7)
    n, p = 20000, 30000
    f = rng.uniform(0.05, 0.5, size=p)
    packed = np.zeros((p, (n + 3) // 4), np.uint8)
    for k in range(4):
        c = (rng.random((p, (n + 3) // 4)) < f[:, None]).astype(np.uint8) + (rng.random((p, (n + 3) // 4)) < f[:, None]).astype(np.uint8)
        packed |= c << (2 * k)
    y = rng.normal(size=n).astype(np.float32)
    outs = []
    for rep in range(2):
        g = jw.GpuSweeper(packed, n, 1)
        g.set_blocks(uniform_starts(p, 512))
        g.put_ycorr(y)
        for it in range(1, 4):
            st = g.sweep_bayesc(jw.SCHED_EXACT, 0.5, 1e-4, 0.98, 2026, it)
        a, b, d = g.get_state(); yc = g.get_ycorr()
        assert st.sum_delta[0] == d.sum() and st.nnz_alpha[0] == np.count_nonzero(a)
        assert np.array_equal(d != 0, a != 0)
        recon = yc + g.mul_alpha(0)
        np.testing.assert_allclose(recon, y, atol=5e-4)
        outs.append((a.copy(), d.copy(), yc.copy()))
        g.close()
    for x, y2 in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(x, y2)


@pytest.mark.parametrize("missing", [0.0, 0.05])
def test_gram_blocks_match_oracle(jw, oracle, missing):
    # GibbsMats XpRinvX (tools4genotypes.jl:263): every block, bit for bit
    prob = Problem(oracle, 517, 300, seed=71, missing=missing)
    g = jw.GpuSweeper(prob.packed, 517, 1)
    starts = np.array([0, 1, 70, 199, 300], dtype=np.int64)
    g.set_blocks(starts)
    for ib in range(4):
        G = g.get_gram(ib)
        ref = oracle.gram_block(prob.packed, 517, prob.means, int(starts[ib]), int(starts[ib + 1] - starts[ib]))
        np.testing.assert_array_equal(G.view(np.uint32), ref.view(np.uint32))
    g.close()


def test_synthetic_generator_roundtrip(jw, oracle):
    g = jw.GpuSweeper.synthetic(1001, 257, 1, seed=5, missing_rate=0.01)
    packed = g.get_packed()
    means, xpx = oracle.marker_stats(packed, 1001)
    gm, gx = g.marker_stats()
    np.testing.assert_array_equal(gm, means); np.testing.assert_array_equal(gx, xpx)
    af = means / 2
    assert 0.02 < af.min() and af.max() < 0.56
    codes = np.array([[(packed[j, i >> 2] >> ((i & 3) << 1)) & 3 for j in range(257)] for i in range(1001)])
    assert 0.003 < (codes == 3).mean() < 0.03
    assert (packed[:, -1] >> 2).max() == 0      # padding individuals are code 0 (1001 = 4*250 + 1)
    g.close()


# ---------------------------------------------------------------------------------------------
# engine 1: the persistent fused kernel must give the same bits as the oracle (and engine 0)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("missing", [0.0, 0.03])
@pytest.mark.parametrize("n,p,b", [(500, 2000, 256), (501, 333, 64), (67, 50, 1), (1030, 700, 700),
                                   (60013, 150, 64)])
def test_fused_bayesc_exact(jw, oracle, n, p, b, missing):
    # (60013, ...) needs more row slices than SMs: exercises the multi-slice path
    prob = Problem(oracle, n, p, seed=n + p + 1, missing=missing)
    run_pair_abc(jw, oracle, prob, uniform_starts(p, b), jw.SCHED_EXACT, nsweeps=3, engine=1)


def test_fused_bayesc_block_schedule(jw, oracle):
    prob = Problem(oracle, 400, 157, seed=22, missing=0.02)
    starts = np.array([0, 20, 21, 60, 100, 157], dtype=np.int64)
    run_pair_abc(jw, oracle, prob, starts, jw.SCHED_BLOCK, nsweeps=2, engine=1)


@pytest.mark.parametrize("missing", [0.0, 0.03])
def test_fused_bayesr(jw, oracle, missing):
    prob = Problem(oracle, 700, 900, seed=35, missing=missing)
    run_pair_r(jw, oracle, prob, uniform_starts(900, 256), jw.SCHED_EXACT, 1, nsweeps=3, engine=1)


def test_fused_mt_sampler1(jw, oracle):
    prob = Problem(oracle, 803, 500, seed=45, ntraits=2)
    run_pair_mt(jw, oracle, prob, uniform_starts(500, 128), jw.SCHED_EXACT, nsweeps=3, engine=1)


@pytest.mark.parametrize("engine", [0, 1])
def test_large_panels_walked_in_subblocks(jw, oracle, engine):
    """Panels above 1024 markers (exact schedule only): the chain walks them in sub-blocks and replays the
    commits of earlier sub-blocks from a list -- same bits as the one-marker-per-thread chain/oracle."""
    prob = Problem(oracle, 160, 3100, seed=91, missing=0.0)
    run_pair_abc(jw, oracle, prob, uniform_starts(3100, 1500), jw.SCHED_EXACT, nsweeps=3, engine=engine, pi=0.97)


def test_large_panels_rejected_for_block_schedules(jw, oracle):
    prob = Problem(oracle, 64, 1300, seed=92)
    g = jw.GpuSweeper(prob.packed, 64, 1)
    g.set_blocks(np.array([0, 1300], dtype=np.int64))
    g.put_ycorr(prob.ycorr0)
    with pytest.raises(jw.JwasError, match="exact schedule only"):
        g.sweep_bayesc(jw.SCHED_BLOCK, 1.0, 0.01, 0.9, 1, 1)
    g.close()


@pytest.mark.parametrize("missing", [0.0, 0.03])
@pytest.mark.parametrize("n,p,b", [(500, 2000, 256), (501, 333, 64), (67, 50, 1), (1030, 700, 700),
                                   (60013, 150, 64), (160, 3100, 1500)])
def test_fused_lagged_schedule(jw, oracle, n, p, b, missing):
    """option lag=1: chain of block k overlaps the streaming of block k+1; the rhs of a block is computed
    from ycorr without the previous block's updates and corrected with the cross-Gram -- bit-exact against
    the oracle's lagged schedule."""
    if b > 1024 and missing > 0:
        pytest.skip("panels above 1024 with missing calls run on engine 0")
    prob = Problem(oracle, n, p, seed=n + p + 7, missing=missing)
    run_pair_abc(jw, oracle, prob, uniform_starts(p, b), jw.SCHED_EXACT, nsweeps=3, engine=1, lag=1,
                 pi=(0.97 if b > 1024 else 0.9))


@pytest.mark.parametrize("chain_ctas", [2, 4])
@pytest.mark.parametrize("n,p,b,missing", [(500, 2000, 256, 0.0), (501, 333, 64, 0.03), (67, 50, 1, 0.0),
                                           (1030, 700, 700, 0.03), (60013, 150, 64, 0.0), (160, 3100, 1500, 0.0),
                                           (300, 9000, 4096, 0.0), (200, 2500, 2048, 0.0), (160, 3100, 1500, 0.03)])
def test_fused_pipelined_chain(jw, oracle, n, p, b, missing, chain_ctas):
    """option chain_ctas: the chain of the lagged schedule walks units of <= 1024 markers on several chain CTAs
    that hand each other commit records (jw_chain_pipe.cuh).  Same sums in the same order: bit-exact against
    the oracle's lagged schedule, whatever the number of chain CTAs."""
    if b >= 4096 and chain_ctas != 2:
        pytest.skip("largest panel: one chain layout (keeps the oracle side short)")
    prob = Problem(oracle, n, p, seed=n + p + 7, missing=missing)
    run_pair_abc(jw, oracle, prob, uniform_starts(p, b), jw.SCHED_EXACT, nsweeps=(2 if b >= 4096 else 3), engine=1, lag=1,
                 pi=(0.97 if b > 1024 else 0.9), chain_ctas=chain_ctas)


@pytest.mark.parametrize("chain_ctas,gather", [(1, 0), (4, 0), (2, 1)])
@pytest.mark.parametrize("n,p,b,missing", [(500, 2000, 256, 0.0), (501, 333, 64, 0.03), (67, 50, 1, 0.0),
                                           (1030, 700, 700, 0.03), (60013, 150, 64, 0.0), (160, 3100, 1500, 0.0),
                                           (300, 9000, 4096, 0.0), (200, 2500, 2048, 0.0), (160, 3100, 1500, 0.03),
                                           (300, 4100, 1024, 0.0)])
def test_fused_lag2_pipelined_chain(jw, oracle, n, p, b, missing, chain_ctas, gather):
    """option lag=2: the stream of block k carries the updates of blocks <= k-3; the chain corrects the rhs with the
    cross-Grams of blocks k-2 and k-1 (oldest first, commit order).  Three panels in flight hide the chain and the
    multi-GPU hand-off behind the stream.  Bit-exact against the oracle's lag-2 schedule."""
    if b >= 4096 and chain_ctas != 4:
        pytest.skip("largest panel: one chain layout (keeps the oracle side short)")
    prob = Problem(oracle, n, p, seed=n + p + 9, missing=missing)
    run_pair_abc(jw, oracle, prob, uniform_starts(p, b), jw.SCHED_EXACT, nsweeps=(2 if b >= 4096 else 3), engine=1, lag=2,
                 pi=(0.97 if b > 1024 else 0.9), chain_ctas=chain_ctas, gather=gather)


def test_fused_lag2_other_methods(jw, oracle):
    prob = Problem(oracle, 700, 900, seed=36, missing=0.02)
    run_pair_r(jw, oracle, prob, uniform_starts(900, 128), jw.SCHED_EXACT, 1, nsweeps=3, engine=1, lag=2, chain_ctas=4)
    prob = Problem(oracle, 803, 500, seed=46, ntraits=2)
    run_pair_mt(jw, oracle, prob, uniform_starts(500, 64), jw.SCHED_EXACT, nsweeps=3, engine=1, lag=2, chain_ctas=2)
    prob = Problem(oracle, 403, 2500, seed=47, ntraits=2)
    run_pair_mt(jw, oracle, prob, uniform_starts(2500, 1300), jw.SCHED_EXACT, nsweeps=2, engine=1, lag=2, sampler="II",
                chain_ctas=3)
    prob = Problem(oracle, 300, 1200, seed=48)
    run_pair_abc(jw, oracle, prob, uniform_starts(1200, 100), jw.SCHED_EXACT, nsweeps=2, engine=1, lag=2, pi=0.0,
                 chain_ctas=2)          # BayesA regime: every marker commits


@pytest.mark.parametrize("chain_ctas", [2, 3])
def test_fused_pipelined_chain_other_methods(jw, oracle, chain_ctas):
    prob = Problem(oracle, 700, 900, seed=36, missing=0.02)
    run_pair_r(jw, oracle, prob, uniform_starts(900, 256), jw.SCHED_EXACT, 1, nsweeps=3, engine=1, lag=1,
               chain_ctas=chain_ctas)
    prob = Problem(oracle, 350, 2600, seed=37)
    run_pair_r(jw, oracle, prob, uniform_starts(2600, 2048), jw.SCHED_EXACT, 1, nsweeps=2, engine=1, lag=1,
               chain_ctas=chain_ctas)
    prob = Problem(oracle, 803, 500, seed=46, ntraits=2)
    run_pair_mt(jw, oracle, prob, uniform_starts(500, 128), jw.SCHED_EXACT, nsweeps=3, engine=1, lag=1,
                chain_ctas=chain_ctas)
    prob = Problem(oracle, 403, 2500, seed=47, ntraits=2)
    run_pair_mt(jw, oracle, prob, uniform_starts(2500, 1300), jw.SCHED_EXACT, nsweeps=2, engine=1, lag=1, sampler="II",
                chain_ctas=chain_ctas)


@pytest.mark.parametrize("n,p,b,missing", [(500, 2000, 256, 0.0), (1030, 700, 700, 0.03), (300, 9000, 4096, 0.0)])
def test_pipelined_chain_inline_replay(jw, oracle, n, p, b, missing):
    """option gather=0: the streaming CTAs replay the commit records in line (kernel mode 1) instead of on a
    gather warp (mode 2, the default when every streaming CTA owns one slice)."""
    prob = Problem(oracle, n, p, seed=n + p + 9, missing=missing)
    run_pair_abc(jw, oracle, prob, uniform_starts(p, b), jw.SCHED_EXACT, nsweeps=(2 if b >= 4096 else 3), engine=1, lag=1,
                 pi=(0.97 if b > 1024 else 0.9), chain_ctas=2, gather=0)


def test_pipelined_chain_dense_start(jw, oracle):
    """pi = 0.2: most markers commit, units carry hundreds of records (the dense start of a chain)."""
    prob = Problem(oracle, 400, 3000, seed=5)
    run_pair_abc(jw, oracle, prob, uniform_starts(3000, 2048), jw.SCHED_EXACT, nsweeps=2, engine=1, lag=1, pi=0.2,
                 chain_ctas=4)


def test_fused_lagged_bayesr_and_multitrait(jw, oracle):
    prob = Problem(oracle, 700, 900, seed=36, missing=0.02)
    run_pair_r(jw, oracle, prob, uniform_starts(900, 256), jw.SCHED_EXACT, 1, nsweeps=3, engine=1, lag=1)
    prob = Problem(oracle, 803, 500, seed=46, ntraits=2)
    run_pair_mt(jw, oracle, prob, uniform_starts(500, 128), jw.SCHED_EXACT, nsweeps=3, engine=1, lag=1)


def test_gram_gemm_equals_popcount_at_full_n(jw):
    """The bf16 tensor-core GEMM and the popcount kernel must give the same Gram blocks bit for bit
    (integer pair counts below 2^24 are exact in FP32), at the benchmark's n, with and without missing."""
    for miss in (0.0, 0.01):
        outs = []
        for popc in (0, 1):
            g = jw.GpuSweeper.synthetic(50000, 700, 1, seed=3, missing_rate=miss)
            g.set_option("gram_popcount", popc)
            g.set_option("lag", 1)
            g.set_blocks(np.array([0, 300, 700], dtype=np.int64))
            outs.append([g.get_gram(0), g.get_gram(1)])
            g.close()
        for a, b in zip(*outs):
            np.testing.assert_array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("engine,lag", [(0, 0), (1, 0), (1, 1)])
def test_mt_sampler2_joint_states(jw, oracle, engine, lag):
    # _MTBayesABC_samplerII! (MTBayesABC.jl:129-210): joint draw of the 4 inclusion states, 2 traits
    prob = Problem(oracle, 403, 600, seed=47, ntraits=2)
    run_pair_mt(jw, oracle, prob, uniform_starts(600, 200), jw.SCHED_EXACT, nsweeps=3, engine=engine, lag=lag, sampler="II")


def test_mt_sampler2_block_schedule(jw, oracle):
    prob = Problem(oracle, 200, 70, seed=48, ntraits=2)
    run_pair_mt(jw, oracle, prob, uniform_starts(70, 16), jw.SCHED_BLOCK, nsweeps=2, sampler="II")


@pytest.mark.parametrize("t", [3, 4])
def test_mt_sampler1_three_and_four_traits(jw, oracle, t):
    """Sampler I is written for any number of traits (MTBayesABC.jl:78-125); the device path covers t <= 4."""
    n, p = 257, 180
    prob = Problem(oracle, n, p, seed=60 + t, ntraits=t, missing=0.01)
    g = jw.GpuSweeper(prob.packed, n, t)
    starts = uniform_starts(p, 64)
    g.set_blocks(starts)
    yc, al, be, de = prob.fresh_state()
    g.put_ycorr(yc); g.put_state(al, be, de)
    rng = np.random.default_rng(t)
    A = rng.normal(size=(t, t)); R = (A @ A.T + t * np.eye(t)) * prob.vary * 0.2
    B = rng.normal(size=(t, t)); G = (B @ B.T + t * np.eye(t)) * prob.vary * 0.02 / (prob.xpx.mean() / n * p)
    bigPi = rng.dirichlet(np.ones(1 << t))
    for it in (1, 2, 3):
        rc, _ = oracle.sweep_contract(prob.packed, n, prob.means, prob.xpx, starts, yc, al, be, de,
                                      method=oracle.METHOD_MT1, R=R, G=G, bigPi=bigPi, seed=4, it=it)
        assert rc == 0
        st = g.sweep_mt1(jw.SCHED_EXACT, R, G, bigPi, 4, it)
        ga, gb, gd = g.get_state()
        np.testing.assert_array_equal(gd, de)
        np.testing.assert_array_equal(ga.view(np.uint32), al.view(np.uint32))
        np.testing.assert_array_equal(g.get_ycorr().view(np.uint32), yc.view(np.uint32))
        states = sum(de[k * p:(k + 1) * p] << k for k in range(t))
        assert [st.class_counts[q] for q in range(1 << t)][:16] == np.bincount(states, minlength=1 << t).tolist()[:16]
    assert de.sum() > 0
    g.close()


def test_tiny_problems(jw, oracle):
    """n < 4 individuals, a single marker, blocks of one marker."""
    for n, p in ((3, 5), (1, 1), (9, 1)):
        codes = np.array([[(i + j) % 3 for j in range(p)] for i in range(n)])
        if n == 1:
            codes[:] = 1
        packed = oracle.pack_codes(codes)
        means, xpx = oracle.marker_stats(packed, n)
        starts = uniform_starts(p, 1)
        for engine in (0, 1):
            g = jw.GpuSweeper(packed, n, 1)
            g.set_blocks(starts); g.set_option("engine", engine)
            yc = np.linspace(-1, 1, n).astype(np.float32) if n > 1 else np.array([0.5], np.float32)
            al = np.zeros(p, np.float32); be = np.zeros(p, np.float32); de = np.zeros(p, np.int32)
            g.put_ycorr(yc); g.put_state(al, be, de)
            ve = np.full(p, 0.5); pi = np.full(p, 0.5)
            rc, _ = oracle.sweep_contract(packed, n, means, xpx, starts, yc, al, be, de, vare=1.0, varEffects=ve, pi=pi, seed=2, it=1)
            g.sweep_bayesabc(jw.SCHED_EXACT, 1.0, ve, pi, 2, 1)
            ga, gb, gd = g.get_state()
            np.testing.assert_array_equal(gd, de); np.testing.assert_array_equal(ga, al)
            np.testing.assert_array_equal(g.get_ycorr(), yc)
            g.close()


def test_fixed_point_overflow_is_reported(jw, oracle):
    """ycorr growing more than 4x inside one sweep hits the fixed-point clamp: the call fails loudly
    (sticky flag), it never returns silently wrong numbers."""
    prob = Problem(oracle, 64, 40, seed=77)
    g = jw.GpuSweeper(prob.packed, 64, 1)
    g.set_blocks(uniform_starts(40, 8))
    y = (prob.ycorr0 * 1e-6).astype(np.float32)          # tiny residuals ...
    g.put_ycorr(y)
    alpha = np.zeros(40, np.float32); alpha[:4] = 50.0   # ... but huge effects that the first block removes
    g.put_state(alpha, alpha.copy(), (alpha != 0).astype(np.int32))
    with pytest.raises(jw.JwasError, match="fixed-point overflow"):
        g.sweep_bayesc(jw.SCHED_EXACT, 1.0, 1e-12, 1.0 - 1e-12, 1, 1)
    g.close()


@pytest.mark.parametrize("engine,lag,t,chain_ctas", [(0, 0, 2, 0), (1, 1, 2, 0), (1, 1, 2, 2), (0, 0, 3, 0)])
def test_mega_bayesabc_independent_traits(jw, oracle, engine, lag, t, chain_ctas):
    """megaBayesABC! (BayesABC.jl:1-7, constraint=true): one single-trait step per trait, one column read."""
    n, p = 300, 400
    prob = Problem(oracle, n, p, seed=80 + t, ntraits=t)
    g = jw.GpuSweeper(prob.packed, n, t)
    g.set_option("engine", engine); g.set_option("lag", lag); g.set_option("chain_ctas", chain_ctas)
    starts = uniform_starts(p, 128)
    g.set_blocks(starts)
    yc, al, be, de = prob.fresh_state()
    g.put_ycorr(yc); g.put_state(al, be, de)
    vare = np.array([0.5, 0.7, 0.6, 0.9][:t]) * prob.vary
    ve = np.array([1.0, 2.0, 0.5, 1.5][:t]) * prob.vary * 0.5 / (0.1 * prob.xpx.mean() / n * p)
    pi = np.array([0.9, 0.8, 0.95, 0.85][:t])
    for it in (1, 2, 3):
        rc, _ = oracle.sweep_contract(prob.packed, n, prob.means, prob.xpx, starts, yc, al, be, de,
                                      method=oracle.METHOD_MEGA, R=np.diag(vare), G=np.diag(ve), bigPi=pi, seed=6, it=it, lag=lag)
        assert rc == 0
        st = g.sweep_mega(jw.SCHED_EXACT, vare, ve, pi, 6, it)
        ga, gb, gd = g.get_state()
        np.testing.assert_array_equal(gd, de)
        np.testing.assert_array_equal(ga.view(np.uint32), al.view(np.uint32))
        np.testing.assert_array_equal(gb.view(np.uint32), be.view(np.uint32))
        np.testing.assert_array_equal(g.get_ycorr().view(np.uint32), yc.view(np.uint32))
        assert [st.sum_delta[k] for k in range(t)] == [de[k * p:(k + 1) * p].sum() for k in range(t)]
    assert all(de[k * p:(k + 1) * p].sum() > 0 for k in range(t))
    g.close()


def test_external_marker_means(jw, oracle):
    """jwas_set_marker_means: centring on means computed on a larger sample (get_genotypes centres on all genotyped
    individuals, readgenotypes.jl:372-385, before JWAS.jl:381-402 aligns rows): xpx, Gram blocks, dots and the axpy all
    follow the supplied means -- bit-exact against the contract sweep fed the same means / xpx."""
    from oracle_backend import OracleBackend
    prob = Problem(oracle, 300, 200, seed=77, missing=0.02)
    means = (prob.means + np.random.default_rng(1).normal(0, 0.05, 200)).astype(np.float32)
    starts = uniform_starts(200, 64)
    ob = OracleBackend(prob.packed, 300, 1, starts, means=means)
    for engine, lag, cc in ((0, 0, 0), (1, 2, 2)):
        g = jw.GpuSweeper(prob.packed, 300, 1)
        g.set_marker_means(means)
        g.set_option("engine", engine); g.set_option("lag", lag); g.set_option("chain_ctas", cc)
        g.set_blocks(starts)
        gm, gx = g.marker_stats()
        np.testing.assert_array_equal(gm, means)
        np.testing.assert_array_equal(gx, ob.xpx)
        yc, al, be, de = prob.fresh_state()
        g.put_ycorr(yc); g.put_state(al, be, de)
        ve = np.full(200, 0.02); pi = np.full(200, 0.8)
        for it in (1, 2):
            rc, _ = oracle.sweep_contract(prob.packed, 300, means, ob.xpx, starts, yc, al, be, de, vare=1.0, varEffects=ve,
                                          pi=pi, seed=4, it=it, lag=lag)
            assert rc == 0
            g.sweep_bayesabc(jw.SCHED_EXACT, 1.0, ve, pi, 4, it)
        ga, gb, gd = g.get_state()
        np.testing.assert_array_equal(gd, de)
        np.testing.assert_array_equal(ga.view(np.uint32), al.view(np.uint32))
        np.testing.assert_array_equal(g.get_ycorr().view(np.uint32), yc.view(np.uint32))
        assert de.sum() > 0
        g.close()


def test_host_array_sweep_call(jw, oracle):
    """jwas_sweep_bayesc_host = BayesABC!(..., yCorr, alpha, beta, delta, ...) mutating the caller's arrays
    (BayesABC.jl:60-63): same result as put / sweep / get."""
    prob = Problem(oracle, 400, 300, seed=12)
    starts = uniform_starts(300, 128)
    outs = []
    for host in (False, True):
        g = jw.GpuSweeper(prob.packed, 400, 1)
        g.set_option("engine", 1); g.set_option("lag", 2); g.set_option("chain_ctas", 2)
        g.set_blocks(starts)
        yc, al, be, de = prob.fresh_state()
        for it in (1, 2, 3):
            if host:
                g.sweep_bayesc_host(jw.SCHED_EXACT, 1.0, 0.02, 0.9, 9, it, yc, al, be, de)
            else:
                g.put_ycorr(yc); g.put_state(al, be, de)
                g.sweep_bayesc(jw.SCHED_EXACT, 1.0, 0.02, 0.9, 9, it)
                al, be, de = g.get_state(); yc = g.get_ycorr()
        outs.append((yc.copy(), al.copy(), be.copy(), de.copy()))
        g.close()
    for a, b in zip(*outs):
        np.testing.assert_array_equal(a, b)
    assert outs[0][3].sum() > 0


@pytest.mark.parametrize("lag,chain_ctas", [(1, 2), (2, 4)])
@pytest.mark.parametrize("n,p,b", [(500, 2000, 256), (501, 333, 64), (67, 50, 1), (1030, 700, 700), (60013, 150, 64),
                                   (160, 3100, 1500), (300, 9000, 4096), (200, 2500, 2048), (52000, 4000, 448)])
def test_fused_warp_specialised_stream(jw, oracle, n, p, b, lag, chain_ctas):
    """option ws=1 (kernel MODE 3, jw_fused_ws.cuh): builder warps rebuild one table set while the streaming warps
    run through the other; per-warp release of the panel, no CTA barrier.  Same sums, same order of the per-row
    updates: bit-exact against the oracle's lagged schedules."""
    big = b >= 4096 or n * p > 5e7
    if big and lag != 2:
        pytest.skip("largest cases: the default lag only (keeps the oracle side short)")
    prob = Problem(oracle, n, p, seed=n + p + 13)
    run_pair_abc(jw, oracle, prob, uniform_starts(p, b), jw.SCHED_EXACT, nsweeps=(2 if big else 3), engine=1, lag=lag,
                 pi=(0.97 if b > 1024 else 0.9), chain_ctas=chain_ctas, gather=0, ws=1)


def test_fused_warp_specialised_stream_bayesr_and_dense(jw, oracle):
    prob = Problem(oracle, 700, 900, seed=36)
    run_pair_r(jw, oracle, prob, uniform_starts(900, 128), jw.SCHED_EXACT, 1, nsweeps=3, engine=1, lag=2, chain_ctas=4, ws=1)
    prob = Problem(oracle, 300, 1200, seed=48)
    run_pair_abc(jw, oracle, prob, uniform_starts(1200, 100), jw.SCHED_EXACT, nsweeps=2, engine=1, lag=2, pi=0.0,
                 chain_ctas=2, gather=0, ws=1)      # every marker commits: the builders replay long record lists
