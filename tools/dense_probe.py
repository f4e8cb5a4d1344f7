"""Small dense-regime run for ncu: every marker commits (pi = 0).  PROBE_P markers, 4 sweeps."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jwas_b200
n = int(os.environ.get("PROBE_N", "50000")); p = int(os.environ.get("PROBE_P", "61440"))
g = jwas_b200.GpuSweeper.synthetic(n, p, 1, seed=2026)
g.set_option("engine", 1); g.set_option("lag", int(os.environ.get("PROBE_LAG", "2"))); g.set_option("chain_ctas", int(os.environ.get("PROBE_CC", "4")))
g.set_blocks(np.array(list(range(0, p, 2048)) + [p], dtype=np.int64))
g.put_ycorr(np.random.default_rng(1).standard_normal(n).astype(np.float32))
for it in range(1, 5):
    st = g.sweep_bayesc(jwas_b200.SCHED_EXACT, 1.0, 1e-4, float(os.environ.get("PROBE_PI", "0.0")), 5, it)
    print(f"sweep {it}: {g.last_sweep_ms:.2f} ms model={int(st.sum_delta[0])} active={st.n_active} rounds={st.n_rounds}", flush=True)
g.close()
