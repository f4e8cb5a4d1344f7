#!/usr/bin/env python
"""bench.py -- Gibbs marker-sweeps/sec of the B200 marker-effects sampler (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]      (N > 1: under torchrun)

MAIN LINE (`value`).  A step is one full marker sweep (p single-site updates + the hyper-parameter draws of one
MCMC iteration) of the running chain.  The workload is BASELINE.json configs[1] per GPU:
    N = 1 : single-trait BayesC, 50,000 x 600,000, 2-bit packed (the configuration the metric is quoted on);
    N > 1 : the SAME 50,000-row x 600,000-marker shard on every GPU, i.e. an (N*50,000) x 600,000 problem whose
            rows are sharded over the GPUs (north_star: "individuals shard naturally across the 8 GPUs"; N = 8 is
            400,000 rows, the row count of configs[4]).  Per-GPU work is fixed: "scaling": "weak".
  value  = shard-sweeps per second summed over the GPUs = N * (sweeps of the whole problem per second); for N = 1
           this is plain sweeps/s of cfg2.  `problem_sweeps_per_s` is the un-multiplied figure.
  e2e    = the same through the plugin-style call BayesABC!(..., yCorr, alpha, beta, delta, ...) on HOST arrays
           (jwas_sweep_bayesc_host): pinned host buffers copied in and out around every sweep.
  roofline = algorithmic bytes of the persistent sweep kernel (this GPU's packed rows, read once) / its CUDA-event
           duration averaged over the timed sweeps, against MEASURED_PEAKS.json's HBM copy bandwidth.
NESTED RESULTS (same JSON line, so nothing hides behind the headline):
  regimes   (N = 1) cfg2 with pi fixed at 0.95 (the reference perf scripts' setting) and with pi = 0 (BayesA)
  schedules (N = 1) cfg2 exact / exact-block (nreps = b) / independent-block with b = 223 (fast_blocks=true);
            one outer iteration counts as b sweeps (JWAS.jl:312)
  configs   cfg3 (BayesR) and cfg4 (2-trait BayesC-pi) on N GPUs; cfg5 (BayesB 400,000 x 1,000,000) when N = 8
  ingest    (N = 1) host only: genotype text file -> .jgb2 through libjwasio at 10,000 x 5,000 (tools/ingest_bench.py)
  strong    (N > 1) cfg2 itself, rows sharded over the N GPUs (fixed total work); same burn-in / warm-up / step counts
            as the main line, so its state_crc equals the N = 1 line's state_crc (bit-identical chains at any N)
--impl reference : the reference algorithm on the host cores (oracle restatement, dense Float32 dot + axpy per
  marker, all cores) on a bounded marker sample of the same workload; ms_per_step is the time of the sampled step,
  value the rate extrapolated linearly in p (the reference's own method, docs/src/manual/benchmark.md:74-77).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHARD_ROWS = 50000
CONFIGS = {
    "cfg1": dict(n=500, p=2000, t=1, method="BayesC", desc="single-trait BayesC pi=0.95, 500 x 2,000"),
    "cfg2": dict(n=50000, p=600000, t=1, method="BayesC", desc="single-trait BayesC, 50,000 x 600,000, 2-bit packed"),
    "cfg3": dict(n=50000, p=1000000, t=1, method="BayesR", desc="single-trait BayesR 4-component mixture, 50,000 x 1,000,000"),
    "cfg4": dict(n=20000, p=500000, t=2, method="BayesC", desc="2-trait BayesC-pi (sampler I), 20,000 x 500,000"),
    "cfg5": dict(n=400000, p=1000000, t=1, method="BayesB", desc="single-trait BayesB, 400,000 x 1,000,000"),
}
GAMMA = np.array([0.0, 0.01, 0.1, 1.0])
PI_R = np.array([0.95, 0.03, 0.015, 0.005])


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(name, n, p, world):
    """The workload description both arms print (identical in the b200 and the reference line)."""
    cfg = CONFIGS[name]
    return {"workload": f"{cfg['desc']}" + (f" per GPU: rows sharded over {world} GPUs, {n} x {p} in total" if world > 1 and name == "cfg2" else
                                            f" (n={n}, p={p})") + "; synthetic Binomial(2,f_j) genotypes, 0.1% QTL, h2=0.5",
            "config": name, "n_obs": n, "n_markers": p, "n_traits": cfg["t"], "method": cfg["method"],
            "l2": "inputs (packed M, %.2f GB per GPU) exceed the 126 MB L2" % (p * math.ceil(n / 4) / max(world, 1) / 1e9)}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, dev=0):
        super().__init__(daemon=True)
        self.dev = dev; self.stop = threading.Event(); self.samples = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- CPU legs
def _dense_sample(n, p_cpu, rng):
    f = rng.uniform(0.05, 0.5, size=p_cpu).astype(np.float32)
    X = np.empty((n, p_cpu), dtype=np.float32, order="F")
    for j in range(p_cpu):
        c = (rng.random(n, dtype=np.float32) < f[j]).astype(np.float32) + (rng.random(n, dtype=np.float32) < f[j])
        X[:, j] = c - c.mean(dtype=np.float32)
    return X


def cpu_reference_leg(name, n, p, steps, warmup, nthreads, p_cpu=None, variants=False, seed=11):
    """The reference algorithm on the host: the oracle's faithful restatements (dense Float32 dot + axpy per marker;
    BayesABC.jl:24-80, BayesR.jl:45-97, MTBayesABC.jl:57-127) on p_cpu markers at full n, threads over n like BLAS
    threads (benchmarks/jwas_nonblock_benchmark.jl:21-22).  Returns (cpu_baseline dict, seconds per sampled sweep)."""
    from oracle import pyoracle as orc
    cfg = CONFIGS[name]
    t, method = cfg["t"], cfg["method"]
    if p_cpu is None:
        p_cpu = max(200, int(4000 * SHARD_ROWS / n))
    p_cpu = min(p_cpu, p)
    rng = np.random.default_rng(seed)
    X = _dense_sample(n, p_cpu, rng)
    xpx = np.einsum("ij,ij->j", X, X).astype(np.float32)
    y = rng.standard_normal(t * n).astype(np.float32)

    def one_sweep(nt):
        u = rng.random(t * p_cpu); z = rng.standard_normal(t * p_cpu)
        t0 = time.perf_counter()
        if method == "BayesR":
            orc.bayesr_ref(X, xpx, st["y"], st["a"], st["d"], 0.5, 1e-3, PI_R, GAMMA, u, z, nthreads=nt)
        elif t > 1:
            orc.mtbayesabc_I_ref(X, xpx, st["y"], st["a"], st["b"], st["d"], np.array([[0.5, 0.1], [0.1, 0.5]]),
                                 np.array([[1e-4, 2e-5], [2e-5, 1e-4]]), np.array([0.95, 0.02, 0.02, 0.01]), u, z)
        else:
            orc.bayesabc_ref(X, xpx, st["y"], st["a"], st["b"], st["d"], 0.5, st["ve"], st["pi"], u, z, nthreads=nt)
        return time.perf_counter() - t0

    def fresh():
        shape = (t, p_cpu) if t > 1 else (p_cpu,)
        return {"y": y.copy(), "a": np.zeros(shape, np.float32), "b": np.zeros(shape, np.float32),
                "d": np.zeros(shape, np.int32 if method == "BayesR" else np.float32),
                "ve": np.full(p_cpu, 1e-4, np.float32), "pi": np.full(p_cpu, 0.95)}

    st = fresh()
    nt_used = nthreads if t == 1 else 1                       # the multi-trait restatement is serial (as the reference's loop)
    times = [one_sweep(nt_used) for _ in range(warmup + steps)][warmup:]
    t_sample = float(np.mean(times))
    scale = p / p_cpu
    base = {"value": 1.0 / (t_sample * scale), "unit": "sweeps/s", "cores": nt_used, "kind": "port",
            "sample": f"oracle restatement of the reference sampler ({method}, dense Float32 dot+axpy per marker), n={n}, "
                      f"{p_cpu} of {p} markers x {steps} sweeps, scaled linearly in p; {t_sample * 1e3:.1f} ms per sampled sweep"}
    if variants and t == 1 and method != "BayesR":
        st = fresh()
        t1 = float(np.mean([one_sweep(1) for _ in range(2)][1:]))
        base["one_thread"] = {"value": 1.0 / (t1 * scale), "unit": "sweeps/s", "cores": 1}
        # packed-decode variant (BayesABC_streaming! + decode_marker!, streaming_genotypes.jl:978-1002), 1 thread
        pc = min(p_cpu, 1000)
        codes = (rng.random((n, pc)) < 0.3).astype(np.int8) + (rng.random((n, pc)) < 0.3).astype(np.int8)
        packed = orc.pack_codes(codes)
        means, xp = orc.marker_stats(packed, n)
        ys = y[:n].copy(); a = np.zeros(pc, np.float32); b = np.zeros(pc, np.float32); d = np.zeros(pc, np.float32)
        tt = []
        for _ in range(2):
            u = rng.random(pc); z = rng.standard_normal(pc)
            t0 = time.perf_counter()
            orc.bayesabc_streaming_ref(packed, n, means, xp, ys, a, b, d, 0.5, np.full(pc, 1e-4, np.float32), np.full(pc, 0.95), u, z)
            tt.append(time.perf_counter() - t0)
        base["packed_decode_one_thread"] = {"value": 1.0 / (tt[-1] * p / pc), "unit": "sweeps/s", "cores": 1,
                                            "sample": f"{pc} markers, decode_marker! semantics"}
        base["published_julia"] = "JWAS.jl BayesC non-block, N=50k: 0.11-0.16 sweeps/s at P=100k, 0.037 at P=200k (docs/src/manual/benchmark.md:81-90)"
    return base, t_sample


def cpu_leg_subprocess(name, n, p, steps, warmup, p_cpu=None, variants=False, world=1):
    """cpu_reference_leg in a fresh interpreter: inside the GPU arm the OpenMP runtime is shared with torch, which
    torchrun has told to use ONE thread (OMP_NUM_THREADS=1) -- 24 oracle threads then crawl (measured: 22x slower than one
    thread).  A clean process with OMP_NUM_THREADS = all cores gives the same number as `--impl reference`."""
    # under torchrun the other ranks stay alive (spinning in the final barrier): leave them their cores, or every OpenMP
    # barrier of the 24-thread team waits for a pre-empted thread (measured: 1.0 s instead of 12 ms per sampled sweep)
    nthreads = max(1, (os.cpu_count() or 1) - (2 * world if world > 1 else 0))
    env = dict(os.environ, OMP_NUM_THREADS=str(nthreads))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "OMP_PROC_BIND", "GOMP_CPU_AFFINITY"):
        env.pop(k, None)
    spec = json.dumps(dict(name=name, n=n, p=p, steps=steps, warmup=warmup, p_cpu=p_cpu, variants=variants, nthreads=nthreads))
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-leg", spec], capture_output=True, text=True,
                         timeout=600, env=env)
    if out.returncode != 0:
        raise RuntimeError("cpu leg failed: " + out.stderr[-300:])
    return json.loads(out.stdout.strip().splitlines()[-1])


# ---------------------------------------------------------------------------------------------- GPU workload
class Workload:
    """One chain on one configuration: device-resident genotypes (this rank's rows), simulated phenotypes, the
    reference's default priors, driven through mcmc.run_chain (the host side of MCMC_BayesianAlphabet)."""

    def __init__(self, name, n, p, rank, world, device, *, schedule="exact", block=0, panel=4096, engine=1, lag=2,
                 chain_ctas=6, estimate_pi=True, pi0=None, method=None, seed=2026, opts=None):
        import jwas_b200
        from jwas_b200 import mcmc, multigpu
        self.jw, self.mcmc = jwas_b200, mcmc
        cfg = CONFIGS[name]
        self.name, self.n, self.p, self.t = name, n, p, cfg["t"]
        self.method = method or cfg["method"]
        self.rank, self.world = rank, world
        t = self.t
        t0 = time.perf_counter()
        rows = jwas_b200.shard_range(n, rank, world) if world > 1 else None
        g = jwas_b200.GpuSweeper.synthetic(n, p, t, seed=seed, device=device, rows=rows)
        if world > 1:
            multigpu.shard(g, rank, world)
        self.g = g
        self.schedule_name = schedule
        self.sched = {"exact": jwas_b200.SCHED_EXACT, "block": jwas_b200.SCHED_BLOCK,
                      "independent": jwas_b200.SCHED_INDEPENDENT}[schedule]
        self.block = block if schedule != "exact" else 0
        step = self.block if self.block else panel
        use_lag = int(lag) if (engine == 1 and lag and schedule == "exact") else 0
        g.set_option("engine", engine); g.set_option("lag", use_lag)
        g.set_option("chain_ctas", chain_ctas if use_lag else 0)
        for kv in (opts or []):                                   # A/B knobs: --opt key=value (jwas_set_option)
            key, val = kv.split("=")
            g.set_option(key, int(val))
        g.set_blocks(np.array(list(range(0, p, step)) + [p], dtype=np.int64))
        if world > 1 and engine == 1:
            multigpu.connect(g, world)
        self.engine = dict(engine=engine, lag=use_lag, chain_ctas=chain_ctas if use_lag else 0, panel=step, schedule=schedule)
        means, _ = g.marker_stats()
        # phenotypes: y_k = sum_qtl x_j a_jk + e_k, h2 = 0.5 (cfg4: residual correlation 0.3)
        rng = np.random.default_rng(7)
        nq = max(1, p // 1000)
        a_true = np.zeros((t, p), np.float32)
        qtl = rng.choice(p, nq, replace=False)
        a_true[:, qtl] = rng.standard_normal((t, nq)).astype(np.float32)
        g.put_state(a_true.reshape(-1), None, None)
        gv = np.array([g.mul_alpha(k).astype(np.float64) for k in range(t)])
        e = rng.standard_normal((t, n))
        if t == 2:
            e[1] = 0.3 * e[0] + math.sqrt(1 - 0.09) * e[1]
        y = gv + e * gv.std(axis=1, keepdims=True) + 10.0
        self.vary = y.var(axis=1)
        mu0 = y.mean(axis=1)
        g.put_state(np.zeros(t * p, np.float32), np.zeros(t * p, np.float32), np.zeros(t * p, np.int32))
        self.y0 = (y - mu0[:, None]).astype(np.float32).reshape(-1)
        g.put_ycorr(self.y0)
        self.setup_s = time.perf_counter() - t0
        # priors as the reference sets them (input_data_validation.jl:296-350; tools4genotypes.jl:353-421;
        # build_MME.jl:108-134 for the multi-trait df)
        sum2pq = float((means.astype(np.float64) * (1 - means / 2)).sum())
        df = 4.0
        self.be = mcmc.GpuBackend(g)
        self.be.record = True
        common = dict(n=n, p=p, ntraits=t, method=self.method, schedule=self.sched, output_samples_frequency=10 ** 9,
                      seed=seed, estimate_pi=estimate_pi, block_size=max(1, self.block))
        if t == 1:
            vare = float(np.float32(self.vary[0] / 2))
            if self.method == "BayesR":
                pi = PI_R.copy()
                ve = float(np.float32((self.vary[0] / 2) / (sum2pq * float((GAMMA * PI_R).sum()))))
            else:
                pi = 0.95 if pi0 is None else float(pi0)
                ve = float(np.float32((self.vary[0] / 2) / ((1 - pi) * sum2pq)))
            common.update(df_effect=df, scale_effect=ve * (df - 2) / df, df_res=df, scale_res=vare * (df - 2) / df)
            self.state = dict(vare=vare, var_effect=ve, pi=pi, mu0=list(mu0))
        else:
            big = np.array([0.95, 0.02, 0.02, 0.01])
            R = np.diag(self.vary / 2).astype(np.float32).astype(np.float64)
            denom = np.array([[sum2pq * sum(big[s] for s in range(1 << t) if (s >> i) & 1 and (s >> j) & 1) for j in range(t)]
                              for i in range(t)])
            G = np.diag((self.vary / 2) / np.diag(denom)).astype(np.float32).astype(np.float64)
            dft = df + t
            common.update(df_effect=dft, scale_effect=None, df_res=dft, scale_res=None, scale_G=G * (dft - t - 1),
                          scale_R=R * (dft - t - 1))
            self.state = dict(vare=None, var_effect=None, pi=None, R=R, G=G, big_pi=big, mu0=list(mu0))
        self.common = common
        self.it0 = 0

    def reset(self):
        t, p = self.t, self.p
        self.g.put_state(np.zeros(t * p, np.float32), np.zeros(t * p, np.float32), np.zeros(t * p, np.int32))
        self.g.put_ycorr(self.y0)

    def advance(self, k):
        st = dict(self.state)
        out = self.mcmc.run_chain(self.be, chain_length=k, burnin=10 ** 9, iter0=self.it0, **self.common, **st)
        self.it0 += k
        if self.t == 1:
            self.state.update(vare=out["vare"], var_effect=out["var_effect"], pi=out["pi"], mu0=out["mu"])
        else:
            self.state.update(R=out["vare"], G=out["var_effect"], big_pi=out["pi"], mu0=out["mu"])
        return out

    def timed(self, steps):
        """K steps bracketed by barrier + synchronize on both sides; MAX over ranks.  Returns a result dict."""
        import torch
        g = self.g
        g.set_option("profile", 1)
        self.be.kernel_ms = []; self.be.sweep_ms = []
        _barrier(self.world); torch.cuda.synchronize()
        l0 = g.kernel_launches
        t0 = time.perf_counter()
        out = self.advance(steps)
        torch.cuda.synchronize(); _barrier(self.world)
        dt = _max_over_ranks(time.perf_counter() - t0, self.world)
        g.set_option("profile", 0)
        k_ms = _max_over_ranks(float(np.mean([x[0] for x in self.be.kernel_ms])), self.world)
        k_launch = int(np.mean([x[1] for x in self.be.kernel_ms]))
        dev_ms = _max_over_ranks(float(np.mean(self.be.sweep_ms)), self.world)
        tr = out["trace"]
        mult = self.block if self.schedule_name != "exact" and self.block else 1     # one outer iteration = b sweeps (JWAS.jl:312)
        return {"sweeps_per_s": mult * steps / dt, "ms_per_step": 1e3 * dt / steps, "device_ms_per_step": dev_ms,
                "kernel_ms_per_sweep": k_ms, "kernel_launches_per_sweep": k_launch, "gpu_launches": int(g.kernel_launches - l0),
                "markers_in_model": float(np.mean([x[1] for x in tr])), "active_updates_per_sweep": float(np.mean([x[2] for x in tr])),
                "chain_rounds_per_sweep": float(np.mean([x[3] for x in tr])), "sweeps_per_step": mult}

    def crc(self):
        a, b, d = self.g.get_state()
        return "%08x" % (zlib.crc32(self.g.get_ycorr().tobytes(), zlib.crc32(a.tobytes(), zlib.crc32(d.tobytes()))) & 0xffffffff)

    def roofline(self, res, peak, peak_src, note):
        bytes_launch = self.p * math.ceil(self.n / 4) / self.world / max(res["kernel_launches_per_sweep"], 1)
        kms = res["kernel_ms_per_sweep"] / max(res["kernel_launches_per_sweep"], 1)
        ach = bytes_launch / (kms * 1e-3) / 1e9 if kms > 0 else None
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                "traffic": None, "kernel": "jw_k_fused (persistent sweep kernel)" if self.engine["engine"] == 1 and self.schedule_name != "independent"
                else "packed GEMV kernel(s) of the sweep", "kernel_ms_per_launch": kms,
                "kernel_launches_per_sweep": res["kernel_launches_per_sweep"], "algorithmic_bytes_per_launch": bytes_launch,
                "peak_source": peak_src, "note": note}

    def close(self):
        self.g.close()


def _barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def _max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    v = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v.item())


def e2e_leg(w, steps, units):
    """Plugin-style call with HOST buffers (pinned): BayesABC!(..., yCorr, alpha, beta, delta, ...) mutating host arrays."""
    import torch
    g = w.g
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    a, b, d = g.get_state()
    h_y, h_a, h_b, h_d = pin(g.get_ycorr()), pin(a), pin(b), pin(d)
    st = w.state
    for i in range(2):                                            # warm the pinned path
        g.sweep_bayesc_host(w.sched, st["vare"], st["var_effect"], st["pi"], 2026, w.it0 + 1 + i, h_y, h_a, h_b, h_d)
    _barrier(w.world); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        g.sweep_bayesc_host(w.sched, st["vare"], st["var_effect"], st["pi"], 2026, w.it0 + 3 + i, h_y, h_a, h_b, h_d)
    torch.cuda.synchronize(); _barrier(w.world)
    dt = _max_over_ranks(time.perf_counter() - t0, w.world)
    w.it0 += steps + 2
    io = (h_y.nbytes + h_a.nbytes + h_b.nbytes + h_d.nbytes) * w.world
    return {"value": units * steps / dt, "unit": UNIT, "problem_sweeps_per_s": steps / dt,
            "h2d_bytes_per_step": io, "d2h_bytes_per_step": io,
            "call": "jwas_sweep_bayesc_host (host arrays mutated in place, pinned; every rank moves its replica of ycorr and the state)"}


UNIT = "sweeps/s (per 50,000-row shard, summed over GPUs; N=1: plain sweeps/s)"


def nested(name, n, p, rank, world, device, args, burnin, steps, peak, peak_src, warm=3, **kw):
    """One more configuration / regime / schedule, reported inside the main line."""
    w = Workload(name, n, p, rank, world, device, engine=args.engine, lag=args.lag, chain_ctas=args.chain_ctas,
                 panel=args.panel, **kw)
    try:
        if burnin:
            w.advance(burnin)
        w.advance(warm)
        res = w.timed(steps)
        out = {"workload": workload_config(name, n, p, world)["workload"], "sweeps_per_s": res["sweeps_per_s"],
               "ms_per_step": res["ms_per_step"], "device_ms_per_step": res["device_ms_per_step"], "steps": steps,
               "burnin": burnin, "markers_in_model": res["markers_in_model"], "active_updates_per_sweep": res["active_updates_per_sweep"],
               "engine": w.engine, "setup_s": w.setup_s, "state_crc": w.crc(),
               "roofline": w.roofline(res, peak, peak_src, "per-GPU bytes; see DESIGN.md section 5 for the ceiling that binds")}
        if res["sweeps_per_step"] > 1:
            out["sweeps_per_step"] = res["sweeps_per_step"]
        return out
    finally:
        w.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=list(CONFIGS))
    ap.add_argument("--n", "--nobs", dest="n", type=int, default=0)          # --nobs/--nmarkers: safe under torchrun
    ap.add_argument("--p", "--nmarkers", dest="p", type=int, default=0)
    ap.add_argument("--strong", action="store_true", help="N > 1: shard the configuration's own rows instead of growing them with N")
    ap.add_argument("--schedule", default="exact", choices=["exact", "block", "independent"])
    ap.add_argument("--block", type=int, default=223, help="fast_blocks block size for --schedule block|independent")
    ap.add_argument("--panel", type=int, default=4096, help="look-ahead panel of the exact schedule (markers per block)")
    ap.add_argument("--burnin", type=int, default=40, help="untimed chain iterations before warm-up")
    ap.add_argument("--engine", type=int, default=1)
    ap.add_argument("--lag", type=int, default=2, help="L = lagged exact schedule (the chains of blocks k-L..k-1 overlap the stream of block k)")
    ap.add_argument("--chain-ctas", type=int, default=6, help="chain CTAs of the pipelined chain (0 = one-CTA chain)")
    ap.add_argument("--fixed-pi", action="store_true", help="keep pi fixed at its start value (reference perf scripts: estimatePi=false)")
    ap.add_argument("--pi0", type=float, default=None)
    ap.add_argument("--cpu-markers", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (jwas_set_option), repeatable")
    ap.add_argument("--cpu-leg", default=None, help=argparse.SUPPRESS)       # internal: one CPU leg in a clean process
    ap.add_argument("--no-extras", action="store_true", help="main line only (no nested regimes / schedules / configs)")
    args = ap.parse_args()
    if args.cpu_leg:
        a = json.loads(args.cpu_leg)
        try:        # the parent (a GPU rank) may carry a narrowed CPU affinity mask (NCCL / launcher): use every core
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except Exception:
            pass
        base, _ = cpu_reference_leg(a["name"], a["n"], a["p"], a["steps"], a["warmup"], a.get("nthreads") or os.cpu_count() or 1,
                                    p_cpu=a["p_cpu"], variants=a["variants"])
        base["affinity_cores"] = len(os.sched_getaffinity(0))
        print(json.dumps(base))
        return
    if args.warmup < 3:
        args.warmup = 3
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    weak = world > 1 and args.config == "cfg2" and not args.strong and not args.n
    n = args.n or (cfg["n"] * world if weak else cfg["n"])
    p = args.p or cfg["p"]
    config = workload_config(args.config, n, p, world)
    units = world if (weak or world == 1) else 1
    scaling = "weak" if (weak or world == 1) else "strong"
    ncpu = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        base, t_sample = cpu_reference_leg(args.config, n, p, args.steps, args.warmup, ncpu, p_cpu=args.cpu_markers or None)
        v = base["value"] * units
        line = {"metric": "gibbs_marker_sweeps_per_sec", "value": v, "unit": UNIT, "impl": "reference",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_sample,
                "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "problem_sweeps_per_s": base["value"], "units_per_step": units,
                "note": "ms_per_step is the measured time of one SAMPLED step (a bounded marker sample at full n); value is the rate "
                        "of the full workload extrapolated linearly in p",
                "cpu_baseline": dict(base, value=v, unit=UNIT),
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import jwas_b200
    import torch
    if jwas_b200.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (libjwasb200 has no CPU fallback)")
    device = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(device)
    if world > 1:
        from jwas_b200 import multigpu
        multigpu.init_process_group("nccl")
    peak, peak_src = measured_peaks()

    w = Workload(args.config, n, p, rank, world, device, schedule=args.schedule, block=args.block, panel=args.panel,
                 engine=args.engine, lag=args.lag, chain_ctas=args.chain_ctas, estimate_pi=not args.fixed_pi, pi0=args.pi0,
                 opts=args.opt)
    w.advance(args.burnin)
    w.advance(args.warmup)
    sampler = ClockSampler(device); sampler.start()
    res = w.timed(args.steps)
    sampler.stop.set(); sampler.join()
    crc = w.crc()
    e2e = None
    if w.t == 1 and w.method == "BayesC":
        e2e = e2e_leg(w, max(3, min(args.steps, 10)), units)
    roof = w.roofline(res, peak, peak_src, "exact single-site chain; the lookup-table stream is bound by CUDA-core issue slots and per-panel "
                                           "fixed costs before HBM (DESIGN.md section 5)")
    if (args.config, n, p, world, args.schedule) == ("cfg2", 50000, 600000, 1, "exact"):
        try:        # DRAM bytes of one `ncu --set full` capture of this kernel (tracked under profiles/), per launch
            rd = {l.split(",")[0]: l.strip().split(",") for l in open(os.path.join(ROOT, "profiles", "r2_fused_kernel_ncu_summary.csv"))}
            sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            roof["traffic"] = sum(float(rd[k][2]) * sc[rd[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            roof["traffic_source"] = "profiles/r2_fused_kernel_ncu_summary.csv (ncu --set full capture of this kernel, not measured in this run)"
        except Exception:
            pass
    value = units * res["sweeps_per_s"]
    line = {"metric": "gibbs_marker_sweeps_per_sec", "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "int64 dots / f64 scalars / f32 state", "data": "synthetic",
            "config": config, "problem_sweeps_per_s": res["sweeps_per_s"], "units_per_step": units,
            "unit_definition": "one unit = one sweep (p single-site updates) over one 50,000-row shard; a step sweeps all shards, "
                               "i.e. the whole (N*50,000)-row problem, once; value = units_per_step * 1000 / ms_per_step",
            "engine": dict(w.engine, burnin=args.burnin, pi=("fixed" if args.fixed_pi else "estimated"),
                           parallelism=(f"rows sharded over {world} GPUs (each stores only its rows); int64 block rhs pushed into peer "
                                        "memory over NVLink inside the persistent kernel; chain replicated; one NCCL all-gather of "
                                        "ycorr per sweep") if world > 1 else "one GPU"),
            "device_ms_per_step": res["device_ms_per_step"], "markers_in_model": res["markers_in_model"],
            "active_updates_per_sweep": res["active_updates_per_sweep"], "chain_rounds_per_sweep": res["chain_rounds_per_sweep"],
            "setup_s": w.setup_s, "state_crc": crc, "roofline": roof, "e2e": e2e, "gpu_launches": res["gpu_launches"],
            "clocks": sampler.summary()}
    w.close()

    # ------------------------------------------------------------------ nested results
    if not args.no_extras and args.config == "cfg2" and args.schedule == "exact" and not args.n and not args.p:
        def guard(dst, key, fn):
            try:
                dst[key] = fn()
            except Exception as e:                                   # a nested leg never takes the main line down
                dst[key] = {"error": str(e)[:300]}

        c2 = CONFIGS["cfg2"]
        if world == 1:
            regimes = {}
            guard(regimes, "fixed_pi_0.95", lambda: nested("cfg2", c2["n"], c2["p"], 0, 1, device, args, 10, 5, peak, peak_src,
                                                            estimate_pi=False))
            guard(regimes, "pi_0_bayesA", lambda: nested("cfg2", c2["n"], c2["p"], 0, 1, device, args, 0, 2, peak, peak_src,
                                                          estimate_pi=False, pi0=0.0, method="BayesB"))
            line["regimes"] = regimes
            sch = {"exact": {"sweeps_per_s": res["sweeps_per_s"], "ms_per_step": res["ms_per_step"], "panel": args.panel}}
            for s_ in ("block", "independent"):
                guard(sch, s_, lambda s_=s_: nested("cfg2", c2["n"], c2["p"], 0, 1, device, args, 6, 4, peak, peak_src,
                                                    schedule=s_, block=223))
            line["schedules"] = sch
        else:
            # same burn-in / warm-up / step counts as the main line, so that its state_crc equals the N = 1 line's
            guard(line, "strong", lambda: nested("cfg2", c2["n"], c2["p"], rank, world, device, args, args.burnin, args.steps,
                                                 peak, peak_src, warm=args.warmup))
        cfgs = {}
        for name in ("cfg3", "cfg4") + (("cfg5",) if world == 8 else ()):
            c = CONFIGS[name]
            guard(cfgs, name, lambda c=c, name=name: nested(name, c["n"], c["p"], rank, world, device, args, 30, 8, peak, peak_src))
        line["configs"] = cfgs
    if rank == 0 and not args.no_cpu:
        try:
            base = cpu_leg_subprocess(args.config, n, p, 2, 1, p_cpu=args.cpu_markers or None, variants=True, world=world)
            base["value"] *= units; base["unit"] = UNIT
            line["cpu_baseline"] = base
            if "configs" in line and world == 1:
                for name in ("cfg3", "cfg4"):
                    if isinstance(line["configs"].get(name), dict) and "error" not in line["configs"][name]:
                        c = CONFIGS[name]
                        line["configs"][name]["cpu_baseline"] = cpu_leg_subprocess(name, c["n"], c["p"], 1, 1,
                                                                                   p_cpu=1500 if c["t"] == 1 else 600)
        except Exception as e:                                       # the CPU leg never takes the GPU line down
            line.setdefault("cpu_baseline", {"error": str(e)[:300]})
        if world == 1 and not args.no_extras:
            # SURVEY 8(f) N2, host only: genotype text file -> .jgb2 through libjwasio at the size the reference documents
            # (10,000 x 5,000: prepare_streaming_genotypes 11.99 s, docs/src/manual/streaming_genotype_backend.md:178-182)
            try:
                out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "ingest_bench.py"),
                                      "--reps", "2"], capture_output=True, text=True, timeout=180)
                line["ingest"] = json.loads(out.stdout.strip().splitlines()[-1])
            except Exception as e:
                line["ingest"] = {"error": str(e)[:300]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
