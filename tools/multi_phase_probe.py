"""torchrun --nproc-per-node N tools/multi_phase_probe.py : engine-1 phase timers (library built with -DJW_TIMERS,
selected with JWAS_B200_LIB) of the row-sharded sweep, rank 0; N = 1 works without torchrun."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jwas_b200
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
n = int(os.environ.get("PROBE_N", "50000")); p = int(os.environ.get("PROBE_P", "600000")); panel = int(os.environ.get("PROBE_PANEL", "2048"))
cc = int(os.environ.get("PROBE_CC", "2")); lag = int(os.environ.get("PROBE_LAG", "1")); ppi = float(os.environ.get("PROBE_PI", "0.999"))
if world > 1:
    import torch
    torch.cuda.set_device(local)
    from jwas_b200 import multigpu
    multigpu.init_process_group("nccl")
rows = jwas_b200.shard_range(n, rank, world) if world > 1 else None
g = jwas_b200.GpuSweeper.synthetic(n, p, 1, seed=2026, device=local, rows=rows)
if world > 1:
    multigpu.shard(g, rank, world)
g.set_option("engine", 1); g.set_option("lag", lag); g.set_option("chain_ctas", cc); g.set_option("timers", 1)
g.set_blocks(np.array(list(range(0, p, panel)) + [p], dtype=np.int64))
if world > 1:
    multigpu.connect(g, world)
g.put_ycorr(np.random.default_rng(1).standard_normal(n).astype(np.float32))
nb = (p + panel - 1) // panel
for it in range(1, 7):
    st = g.sweep_bayesc(jwas_b200.SCHED_EXACT, 1.0, 2e-3, ppi, 5, it)
    raw = g.phase_ns().astype(np.float64)
    ph = raw / nb
    nu = max(raw[29], 1.0)
    if rank == 0:
        print(f"N={world} sweep {it}: {g.last_sweep_ms:.2f} ms model={int(st.sum_delta[0])} active={st.n_active} | stream CTA0 ns/block: records+axpy={ph[0]:.0f} tables={ph[1]:.0f} "
              f"stream={ph[2]:.0f} (warp0 done {ph[6]:.0f}, last warp {ph[15]:.0f}, barrier {ph[7]:.0f}, arrive {ph[13]:.0f}) | chain CTA0 ns/unit ({int(nu)} units): "
              f"preload={raw[24]/nu:.0f} wait_rhs={raw[25]/nu:.0f} rhs+records={raw[26]/nu:.0f} rounds={raw[27]/nu:.0f} epilogue={raw[28]/nu:.0f}", flush=True)
g.close()
if world > 1:
    import torch.distributed as dist
    dist.barrier(); dist.destroy_process_group()
