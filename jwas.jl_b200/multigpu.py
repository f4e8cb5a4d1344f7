"""One-process-per-GPU driver for the row-sharded sweep.

torch.distributed is plumbing only: rendezvous, broadcasting the 128-byte NCCL unique id, barriers and
the max-over-ranks of the timings.  The data-path collectives (one int64 all-reduce of the block rhs per
marker block, one gather of the ycorr shards per sweep) are issued by libjwasb200 itself on its own
stream (csrc/jw_nccl.cuh), so no Python runs between a block's GEMV and its chain.
"""
import json
import math
import os
import time

import numpy as np


def shard_bounds(n, world):
    """Row boundaries of every rank: words of 64 individuals (16 packed bytes) split evenly, the remainder to
    the first ranks (mirrors jwas_shard_range)."""
    nw = (n + 63) // 64
    b = [min(n, (nw // world * r + min(r, nw % world)) * 64) for r in range(world)]
    return b + [n]


def init_process_group(backend=None):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group(backend=backend)
    return dist.get_rank(), dist.get_world_size()


def broadcast_bytes(payload, nbytes, src=0):
    """Broadcast a byte string from rank `src` (the NCCL unique id) over the default group."""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def shard(sweeper, rank, world):
    """Make the sweeper rank `rank` of a `world`-way row-sharded sweep (before set_blocks): a sweeper holding the
    full matrix keeps only its rows, one created with rows=shard_range(...) is checked; marker statistics are
    summed over the ranks inside the library (NCCL)."""
    from ._lib import nccl_unique_id
    uid = broadcast_bytes(nccl_unique_id() if rank == 0 else b"", 128, src=0) if world > 1 else None
    sweeper.init_sharding(rank, world, uid)
    return sweeper


def connect(sweeper, world):
    """After set_blocks: exchange buffers of the in-kernel NVLink reduction (all-gather of the 64-byte IPC handles)."""
    if world > 1:
        sweeper.ipc_import(all_gather_bytes(sweeper.ipc_export(), 64))
    return sweeper


def all_gather_bytes(payload, nbytes):
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    outs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, mine)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def bench_main(args, cfg, config):
    """bench.py --gpus N (N > 1): the same chain on the same synthetic data, rows sharded over N GPUs."""
    import torch
    import torch.distributed as dist
    import jwas_b200
    from jwas_b200 import mcmc
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    rank, world = init_process_group("nccl")
    n = args.n or cfg["n"]; p = args.p or cfg["p"]
    g = jwas_b200.GpuSweeper.synthetic(n, p, 1, seed=2026, device=local)      # identical on every rank
    panel = args.panel
    g.set_option("engine", args.engine)
    g.set_option("lag", 1 if args.engine == 1 else 0)
    g.set_option("chain_ctas", args.chain_ctas if args.engine == 1 else 0)
    starts = np.array(list(range(0, p, panel)) + [p], dtype=np.int64)
    g.set_blocks(starts)
    attach(g, rank, world, fused=(args.engine == 1))
    means, _ = g.marker_stats()
    rng = np.random.default_rng(7)
    nq = max(1, p // 1000)
    a_true = np.zeros(p, np.float32)
    a_true[rng.choice(p, nq, replace=False)] = rng.standard_normal(nq).astype(np.float32)
    g.put_state(a_true, None, None)
    gv = g.mul_alpha(0).astype(np.float64)
    y = gv + rng.standard_normal(n) * gv.std() + 10.0
    vary = float(y.var()); mu0 = float(y.mean())
    g.put_state(np.zeros(p, np.float32), np.zeros(p, np.float32), np.zeros(p, np.int32))
    g.put_ycorr((y - mu0).astype(np.float32))
    pi0 = 0.95
    sum2pq = float((means.astype(np.float64) * (1 - means / 2)).sum())
    vare = float(np.float32(vary / 2)); var_effect = float(np.float32((vary / 2) / ((1 - pi0) * sum2pq)))
    df = 4.0
    be = mcmc.GpuBackend(g)
    common = dict(n=n, p=p, ntraits=1, method="BayesC", schedule=jwas_b200.SCHED_EXACT,
                  output_samples_frequency=10 ** 9, seed=2026, df_effect=df, scale_effect=var_effect * (df - 2) / df,
                  df_res=df, scale_res=vare * (df - 2) / df, estimate_pi=not args.fixed_pi)
    state = dict(vare=vare, var_effect=var_effect, pi=pi0, mu0=[mu0])
    it0 = 0

    def advance(k):
        nonlocal state, it0
        out = mcmc.run_chain(be, chain_length=k, burnin=10 ** 9, iter0=it0, **common, **state)
        state = dict(vare=out["vare"], var_effect=out["var_effect"], pi=out["pi"], mu0=out["mu"])
        it0 += k
        return out

    advance(args.burnin); advance(args.warmup)
    g.set_option("profile", 1)
    dist.barrier(); torch.cuda.synchronize()
    launches0 = g.kernel_launches
    t0 = time.perf_counter()
    out = advance(args.steps)
    torch.cuda.synchronize(); dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    k_ms, k_launches = g.stream_kernel_ms()
    launches = g.kernel_launches - launches0
    if rank == 0:
        from bench import measured_peaks
        peak, src = measured_peaks()
        bytes_local = p * math.ceil(n / 4) / world
        ach = bytes_local / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
        line = {"metric": "gibbs_marker_sweeps_per_sec", "value": args.steps / dt, "unit": "sweeps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "int64 dots / f64 scalars / f32 state", "data": "synthetic",
                "config": dict(config, parallelism=(f"rows sharded over {world} GPUs; persistent kernel pushes the int64 block rhs "
                               "into peer memory over NVLink (IPC), chain replicated" if args.engine == 1 else
                               f"rows sharded over {world} GPUs; one int64 NCCL all-reduce of the block rhs per marker block, "
                               "chain replicated"), engine=args.engine, panel=panel,
                               markers_in_model=float(np.mean([t[1] for t in out["trace"]]))),
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": (ach / peak) if ach else None, "traffic": None, "peak_source": src,
                             "kernel": ("jw_k_fused (per-GPU bytes)" if args.engine == 1 else
                                        "jw_k_block_dot (per-GPU bytes, summed over the sweep's launches)"),
                             "kernel_ms_per_sweep": k_ms, "kernel_launches_per_sweep": k_launches},
                "e2e": None, "gpu_launches": int(launches)}
        print(json.dumps(line))
    dist.barrier()
    g.close()
    dist.destroy_process_group()
