"""torchrun --nproc-per-node N tools/multigpu_check.py : the row-sharded sweep on N GPUs (every rank stores only
its own rows; marker statistics and Gram blocks summed over the ranks) must give the same bits as the single-GPU
sweep (engine 0 and engine 1) for BayesC (with missing data), BayesR and 2-trait BayesC -- state, ycorr, marker
statistics and M*alpha."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import jwas_b200
from jwas_b200 import multigpu

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
rank, world = multigpu.init_process_group("nccl")
n, p = 30011, 3000
GAMMA = np.array([0.0, 0.01, 0.1, 1.0]); PI_R = np.array([0.95, 0.03, 0.015, 0.005])


def run(t, method, sharded, engine, miss, lag=0, chain_ctas=0, from_full=False, ws=0):
    # sharded: this rank creates (and stores) only its own rows, or -- from_full -- creates the whole matrix and
    # lets jwas_init_sharding drop the other ranks' rows
    rows = jwas_b200.shard_range(n, rank, world) if (sharded and not from_full) else None
    g = jwas_b200.GpuSweeper.synthetic(n, p, t, seed=11, missing_rate=miss, device=local, rows=rows)
    if sharded:
        multigpu.shard(g, rank, world)
        assert g.row_range() == jwas_b200.shard_range(n, rank, world)
    g.set_option("engine", engine)
    g.set_option("lag", lag)
    g.set_option("chain_ctas", chain_ctas)
    g.set_option("ws", ws)
    g.set_blocks(np.array(list(range(0, p, 512)) + [p], dtype=np.int64))
    if sharded and engine == 1:
        multigpu.connect(g, world)
    y = np.random.default_rng(3).standard_normal(t * n).astype(np.float32)
    g.put_ycorr(y)
    if method == "R":
        g.put_state(None, None, np.ones(p, np.int32))
    for it in (1, 2, 3):
        if method == "C":
            g.sweep_bayesc(jwas_b200.SCHED_EXACT, 1.0, 2e-3, 0.99, 5, it)
        elif method == "R":
            g.sweep_bayesr(jwas_b200.SCHED_EXACT, 1, 1.0, 5e-3, PI_R, GAMMA, 5, it)
        elif method == "I":
            g.sweep_bayesc(jwas_b200.SCHED_INDEPENDENT, 1.0, 2e-3, 0.99, 5, it)
        else:
            g.sweep_mt1(jwas_b200.SCHED_EXACT, np.array([[1.0, 0.3], [0.3, 1.0]]), np.array([[2e-3, 5e-4], [5e-4, 2e-3]]),
                        np.array([0.97, 0.01, 0.01, 0.01]), 5, it)
    a, b, d = g.get_state(); yc = g.get_ycorr()
    m, x = g.marker_stats()
    ebv = g.mul_alpha(0)
    g.close()
    return a, b, d, yc, m, x, ebv


ok = True
for t, method, miss in ((1, "C", 0.01), (1, "R", 0.0), (2, "M", 0.0), (1, "I", 0.0)):
    ref = run(t, method, False, 0, miss)
    fused = run(t, method, False, 1, miss) if method != "I" else ref
    sh = run(t, method, True, 0, miss)
    same = all(np.array_equal(x, y) for x, y in zip(ref, sh)) and all(np.array_equal(x, y) for x, y in zip(ref, fused))
    if method != "I":
        # fused persistent kernel with the in-kernel NVLink reduction (lagged schedule) vs one GPU
        ref1 = run(t, method, False, 1, miss, lag=1)
        sh1 = run(t, method, True, 1, miss, lag=1)
        sh1f = run(t, method, True, 1, miss, lag=1, from_full=True)
        same1 = all(np.array_equal(x, y) for x, y in zip(ref1, sh1)) and all(np.array_equal(x, y) for x, y in zip(ref1, sh1f))
        print(f"rank {rank}/{world} method {method} t={t}: fused multi-GPU (NVLink push) == fused single GPU: {same1}", flush=True)
        same = same and same1
        # the same with the chain pipelined over two chain CTAs (commit records), one GPU and sharded
        ref2 = run(t, method, False, 1, miss, lag=1, chain_ctas=2)
        sh2 = run(t, method, True, 1, miss, lag=1, chain_ctas=2)
        same2 = all(np.array_equal(x, y) for x, y in zip(ref1, ref2)) and all(np.array_equal(x, y) for x, y in zip(ref1, sh2))
        print(f"rank {rank}/{world} method {method} t={t}: pipelined chain, sharded == single == one-CTA chain: {same2}", flush=True)
        same = same and same2
        # lag 2 (three panels in flight), four chain CTAs
        ref3 = run(t, method, False, 1, miss, lag=2, chain_ctas=4)
        sh3 = run(t, method, True, 1, miss, lag=2, chain_ctas=4)
        same3 = all(np.array_equal(x, y) for x, y in zip(ref3, sh3))
        print(f"rank {rank}/{world} method {method} t={t}: lag 2, sharded == single: {same3}", flush=True)
        same = same and same3
        if miss == 0.0 and t == 1:
            # warp-specialised streaming role (kernel MODE 3), sharded, against the plain role on one GPU
            sh4 = run(t, method, True, 1, miss, lag=2, chain_ctas=4, ws=1)
            same4 = all(np.array_equal(x, y) for x, y in zip(ref3, sh4))
            print(f"rank {rank}/{world} method {method} t={t}: warp-specialised stream, sharded == plain single: {same4}", flush=True)
            same = same and same4
    nz = int(np.count_nonzero(ref[0]))
    print(f"rank {rank}/{world} method {method} t={t}: sharded==single==fused: {same} (nonzero effects {nz})", flush=True)
    ok = ok and same and nz > 0
res = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(res, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("MULTIGPU_CHECK", "PASS" if int(res.item()) == 1 else "FAIL")
sys.exit(0 if int(res.item()) == 1 else 1)
