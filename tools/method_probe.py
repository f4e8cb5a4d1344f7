"""Sweep time of BayesR (cfg3 shape) and 2-trait BayesC (cfg4 shape) at full size on one GPU."""
import argparse, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jwas_b200
ap = argparse.ArgumentParser()
ap.add_argument("--method", default="R"); ap.add_argument("--n", type=int, default=50000); ap.add_argument("--p", type=int, default=1000000)
ap.add_argument("--panel", type=int, default=2048); ap.add_argument("--chain-ctas", type=int, default=2); ap.add_argument("--sweeps", type=int, default=12); ap.add_argument("--lag", type=int, default=1)
a = ap.parse_args()
t = 2 if a.method == "MT" else 1
t0 = time.time()
g = jwas_b200.GpuSweeper.synthetic(a.n, a.p, t, seed=2026)
g.set_option("lag", a.lag); g.set_option("chain_ctas", a.chain_ctas if a.lag else 0)
g.set_blocks(np.array(list(range(0, a.p, a.panel)) + [a.p], dtype=np.int64)); g.set_option("engine", 1)
print("setup %.1fs" % (time.time() - t0))
rng = np.random.default_rng(1)
g.put_ycorr(rng.standard_normal(t * a.n).astype(np.float32))
GAMMA = np.array([0.0, 0.01, 0.1, 1.0]); PI = np.array([0.999, 0.0006, 0.0003, 0.0001])
if a.method == "R":
    g.put_state(None, None, np.ones(a.p, np.int32))
for it in range(1, a.sweeps + 1):
    if a.method == "R":
        st = g.sweep_bayesr(jwas_b200.SCHED_EXACT, 1, 1.0, 2e-3, PI, GAMMA, 5, it)
    else:
        st = g.sweep_mt1(jwas_b200.SCHED_EXACT, np.array([[1.0, 0.3], [0.3, 1.0]]), np.array([[2e-3, 5e-4], [5e-4, 2e-3]]),
                         np.array([0.999, 0.0004, 0.0004, 0.0002]), 5, it)
    print(f"{a.method} sweep {it}: {g.last_sweep_ms:.2f} ms  in-model={int(st.sum_delta[0])} active={st.n_active} rounds={st.n_rounds}")
