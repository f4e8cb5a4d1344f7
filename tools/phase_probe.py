"""Prints engine-1 phase timers (ns per block) at a given size: where a block's time goes.
The timers are a BUILD option (their presence alone costs the sweep ~10 %): build the library with
    JWAS_B200_BUILD_FLAGS=-DJW_TIMERS python jwas.jl_b200/build.py
before running this tool; with the default build every phase reads 0."""
import argparse, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jwas_b200

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=50000); ap.add_argument("--p", type=int, default=600000)
ap.add_argument("--panel", type=int, default=1024); ap.add_argument("--sweeps", type=int, default=6)
ap.add_argument("--lag", type=int, default=0); ap.add_argument("--chain-ctas", type=int, default=0); ap.add_argument("--gather", action="store_true")
ap.add_argument("--pi", type=float, default=0.999); ap.add_argument("--ve", type=float, default=2e-3)
a = ap.parse_args()
g = jwas_b200.GpuSweeper.synthetic(a.n, a.p, 1, seed=2026)
starts = np.array(list(range(0, a.p, a.panel)) + [a.p], dtype=np.int64)
g.set_option("chain_ctas", a.chain_ctas); g.set_option("gather", 1 if a.gather else 0)
g.set_blocks(starts); g.set_option("engine", 1); g.set_option("lag", a.lag); g.set_option("timers", 1)
rng = np.random.default_rng(1)
g.put_ycorr(rng.standard_normal(a.n).astype(np.float32))
nb = len(starts) - 1
for it in range(1, a.sweeps + 1):
    st = g.sweep_bayesc(jwas_b200.SCHED_EXACT, 1.0, a.ve, a.pi, 5, it)
    ph = g.phase_ns().astype(np.float64) / nb
    if a.chain_ctas and a.lag:
        raw = g.phase_ns().astype(np.float64)
        nu = max(raw[29], 1.0)       # units walked by chain CTA 0
        print(f"sweep {it}: {g.last_sweep_ms:.2f} ms  model={int(st.sum_delta[0])} active={st.n_active} rounds={st.n_rounds} | "
              f"stream CTA0 ns/block: records+axpy={ph[0]:.0f} tables={ph[1]:.0f} stream={ph[2]:.0f} (since block top: warp0 done {ph[6]:.0f}, last stream warp {ph[15]:.0f}, gather warp {ph[5]:.0f}, barrier {ph[7]:.0f}, arrive {ph[13]:.0f}) | "
              f"chain CTA0 ns/unit ({int(nu)} units): preload={raw[24]/nu:.0f} wait_rhs={raw[25]/nu:.0f} rhs+records={raw[26]/nu:.0f} rounds={raw[27]/nu:.0f} epilogue={raw[28]/nu:.0f}")
        continue
    print(f"sweep {it}: {g.last_sweep_ms:.2f} ms  model={int(st.sum_delta[0])} active={st.n_active} rounds={st.n_rounds} | "
          f"CTA0 ns/block: wait_prev={ph[0]:.0f} rebuild={ph[1]:.0f} stream={ph[2]:.0f} wait_all={ph[3]:.0f} chain={ph[4]:.0f} | "
          f"CTA1: wait_prev={ph[8]:.0f} rebuild={ph[9]:.0f} stream={ph[10]:.0f} | chainCTA: wait_all={ph[19]:.0f} chain={ph[20]:.0f} | in-chain: preload={ph[24]:.0f} wait={ph[25]:.0f} rhs={ph[26]:.0f} rounds={ph[27]:.0f} epilogue={ph[28]:.0f}")
