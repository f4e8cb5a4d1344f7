#!/usr/bin/env python
"""bench.py -- Gibbs marker-sweeps/sec of the B200 marker-effects sampler.

Metric (BASELINE.json): Gibbs marker-sweeps/sec; achieved HBM GB/s vs the measured copy peak;
next to the reference algorithm on the box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one full marker sweep (p single-site updates) of the running BayesC chain, with the
hyper-parameter draws of one MCMC iteration in between (a few host scalars).
  value : sweeps/s with every input resident in HBM (bracketed by device synchronisation)
  e2e   : sweeps/s through the plugin-style call with HOST buffers: ycorr, alpha, beta, delta are
          copied host->device before and device->host after every sweep (BayesABC!(...) mutates
          host arrays in the reference, BayesABC.jl:60-63); the genotype matrix stays resident,
          as Genotypes does in the reference.
  --impl reference : the reference algorithm (oracle restatement of BayesABC!, dense Float32
          dot+axpy per marker, all host threads) on a bounded marker sample, scaled linearly in p
          (the reference's own extrapolation method, docs/src/manual/benchmark.md:74-77).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (fits one GPU)
    "cfg2": dict(n=50000, p=600000, method="BayesC", desc="single-trait BayesC, 50,000 x 600,000, 2-bit packed"),
    "cfg1": dict(n=500, p=2000, method="BayesC", desc="single-trait BayesC pi=0.95, 500 x 2,000"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


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
            self.stop.wait(0.2)

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


def cpu_reference_leg(n, p, p_cpu, nthreads, steps, warmup, seed=11):
    """The reference algorithm on the host: oracle's faithful BayesABC! (dense Float32 dot+axpy per
    marker; BayesABC.jl:24-80) on p_cpu markers at full n, threads over n like BLAS threads."""
    from oracle import pyoracle as orc
    rng = np.random.default_rng(seed)
    f = rng.uniform(0.05, 0.5, size=p_cpu).astype(np.float32)
    X = np.empty((n, p_cpu), dtype=np.float32, order="F")
    for j in range(p_cpu):
        c = (rng.random(n, dtype=np.float32) < f[j]).astype(np.float32) + (rng.random(n, dtype=np.float32) < f[j])
        X[:, j] = c - c.mean(dtype=np.float32)
    xpx = np.einsum("ij,ij->j", X, X).astype(np.float32)
    y = rng.standard_normal(n).astype(np.float32)
    alpha = np.zeros(p_cpu, np.float32); beta = np.zeros(p_cpu, np.float32); delta = np.zeros(p_cpu, np.float32)
    ve = np.full(p_cpu, 1e-4, np.float32); pi = np.full(p_cpu, 0.95)
    nt = nthreads if nthreads > 0 else orc.max_threads()
    times = []
    for it in range(warmup + steps):
        u = rng.random(p_cpu); z = rng.standard_normal(p_cpu)
        t0 = time.perf_counter()
        orc.bayesabc_ref(X, xpx, y, alpha, beta, delta, 0.5, ve, pi, u, z, nthreads=nt)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t_sample = float(np.mean(times))
    sweeps_per_s = 1.0 / (t_sample * (p / p_cpu))
    return {"value": sweeps_per_s, "unit": "sweeps/s", "cores": nt, "kind": "port",
            "sample": f"oracle restatement of BayesABC! (dense Float32 dot+axpy per marker), n={n}, "
                      f"{p_cpu} of {p} markers x {steps} sweeps, scaled linearly in p; "
                      f"{t_sample * 1e3:.1f} ms per sampled sweep"}, t_sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=list(CONFIGS))
    ap.add_argument("--n", "--nobs", dest="n", type=int, default=0)          # --nobs/--nmarkers: safe under torchrun
    ap.add_argument("--p", "--nmarkers", dest="p", type=int, default=0)
    ap.add_argument("--panel", type=int, default=2048, help="look-ahead panel (markers per block)")
    ap.add_argument("--burnin", type=int, default=40, help="untimed chain iterations before warm-up")
    ap.add_argument("--engine", type=int, default=1)
    ap.add_argument("--lag", type=int, default=1, help="1 = lagged exact schedule (chain k overlaps stream k+1)")
    ap.add_argument("--chain-ctas", type=int, default=2, help="chain CTAs of the pipelined chain (0 = one-CTA chain)")
    ap.add_argument("--gather", action="store_true", help="gather warp per streaming CTA (kernel mode 2; use a panel that is a multiple of 496)")
    ap.add_argument("--fixed-pi", action="store_true", help="keep pi=0.95 fixed (reference perf scripts: estimatePi=false)")
    ap.add_argument("--cpu-markers", type=int, default=4000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    cfg = dict(CONFIGS[args.config])
    n = args.n or cfg["n"]; p = args.p or cfg["p"]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": f"{cfg['desc']} (n={n}, p={p}); synthetic Binomial(2,f_j) genotypes generated on device, "
                          "0.1% QTL, h2=0.5",
              "n_obs": n, "n_markers": p, "panel": args.panel,
              "l2": "inputs (packed M, %.2f GB) exceed the 126 MB L2" % (p * math.ceil(n / 4) / 1e9)}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        base, t_sample = cpu_reference_leg(n, p, args.cpu_markers, 0, args.steps, args.warmup)
        line = {"metric": "gibbs_marker_sweeps_per_sec", "value": base["value"], "unit": "sweeps/s", "impl": "reference",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import jwas_b200
    from jwas_b200 import mcmc
    if world > 1:
        from jwas_b200 import multigpu
        return multigpu.bench_main(args, cfg, config)

    if jwas_b200.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (libjwasb200 has no CPU fallback)")
    t_setup = time.perf_counter()
    g = jwas_b200.GpuSweeper.synthetic(n, p, 1, seed=2026)
    starts = np.array(list(range(0, p, args.panel)) + [p], dtype=np.int64)
    g.set_option("engine", args.engine)
    g.set_option("lag", args.lag if args.engine == 1 else 0)
    g.set_option("chain_ctas", args.chain_ctas if (args.engine == 1 and args.lag) else 0)
    for key, val in (("gather", 1 if args.gather else 0),):
        try:
            g.set_option(key, val)
        except jwas_b200.JwasError:
            pass                                  # an older build under JWAS_B200_LIB does not know this A/B knob
    g.set_blocks(starts)
    means, xpx = g.marker_stats()
    # phenotype: y = sum_qtl x_j a_j + e, h2 = 0.5
    rng = np.random.default_rng(7)
    nq = max(1, p // 1000)
    a_true = np.zeros(p, np.float32)
    a_true[rng.choice(p, nq, replace=False)] = rng.standard_normal(nq).astype(np.float32)
    g.put_state(a_true, None, None)
    gv = g.mul_alpha(0).astype(np.float64)
    y = gv + rng.standard_normal(n) * gv.std() + 10.0
    vary = float(y.var())
    g.put_state(np.zeros(p, np.float32), np.zeros(p, np.float32), np.zeros(p, np.int32))
    mu0 = float(y.mean())
    g.put_ycorr((y - mu0).astype(np.float32))
    setup_s = time.perf_counter() - t_setup

    # priors as the reference sets them (input_data_validation.jl:296-350; tools4genotypes.jl:353-421)
    pi0 = 0.95
    sum2pq = float((means.astype(np.float64) * (1 - means / 2)).sum())
    genetic_var = vary / 2; vare = float(np.float32(vary / 2))
    var_effect = float(np.float32(genetic_var / ((1 - pi0) * sum2pq)))
    df = 4.0
    scale_effect = var_effect * (df - 2) / df; scale_res = vare * (df - 2) / df
    be = mcmc.GpuBackend(g)
    common = dict(n=n, p=p, ntraits=1, method="BayesC", schedule=jwas_b200.SCHED_EXACT, output_samples_frequency=10 ** 9,
                  seed=2026, df_effect=df, scale_effect=scale_effect, df_res=df, scale_res=scale_res,
                  estimate_pi=not args.fixed_pi)
    state = dict(vare=vare, var_effect=var_effect, pi=pi0, mu0=[mu0])

    def advance(k_iters, first_iter):
        # run_chain restarts its iteration counter; offset the seed stream via the iteration base
        nonlocal state
        out = mcmc.run_chain(be, chain_length=k_iters, burnin=10 ** 9, iter0=first_iter, **common, **state)
        state = dict(vare=out["vare"], var_effect=out["var_effect"], pi=out["pi"], mu0=out["mu"])
        return out

    it0 = 0
    out = advance(args.burnin, it0); it0 += args.burnin
    out = advance(args.warmup, it0); it0 += args.warmup

    import torch
    torch.cuda.synchronize()
    sampler = ClockSampler(0); sampler.start()
    g.set_option("profile", 1)
    launches0 = g.kernel_launches
    t0 = time.perf_counter()
    out = advance(args.steps, it0); it0 += args.steps
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = g.kernel_launches - launches0
    sampler.stop.set(); sampler.join()
    k_ms, k_launches = g.stream_kernel_ms()      # the last sweep's streaming kernel(s)
    g.set_option("profile", 0)
    value = args.steps / dt
    trace = out["trace"]
    model_size = float(np.mean([tr[1] for tr in trace]))
    active = float(np.mean([tr[2] for tr in trace])); rounds = float(np.mean([tr[3] for tr in trace]))

    # roofline of the dominant kernel (algorithmic bytes = packed M read once per sweep)
    peak, peak_src = measured_peaks()
    bytes_per_sweep = p * math.ceil(n / 4)
    ach = bytes_per_sweep / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None
    traffic = None      # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture (profiles/)
    try:
        if args.engine == 1 and (n, p) == (50000, 600000):
            rd = {l.split(",")[0]: l.strip().split(",") for l in open(os.path.join(ROOT, "profiles", "r1_fused_kernel_ncu_summary.csv"))}
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            traffic = sum(float(rd[k][2]) * scale[rd[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None, "traffic": traffic,
                "kernel": "jw_k_fused (persistent sweep)" if args.engine == 1 else "jw_k_block_dot (summed over the sweep's launches)",
                "kernel_ms_per_sweep": k_ms, "kernel_launches_per_sweep": k_launches,
                "algorithmic_bytes_per_launch": bytes_per_sweep / max(k_launches, 1), "peak_source": peak_src,
                "note": "exact single-site chain: sequential depth (active markers + blocks), not HBM, bounds the sweep"}

    # e2e: plugin-style call with host buffers (pinned), copies inside the timed region
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    h_y = pin(g.get_ycorr()); al, bt, dl = g.get_state()
    h_a, h_b, h_d = pin(al), pin(bt), pin(dl)
    e_steps = max(3, min(args.steps, 10))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e_steps):
        g.put_ycorr(h_y); g.put_state(h_a, h_b, h_d)
        st = g.sweep_bayesc(jwas_b200.SCHED_EXACT, state["vare"], state["var_effect"], state["pi"], 2026, it0 + i + 1)
        a2, b2, d2 = g.get_state(); y2 = g.get_ycorr()
        h_y[:] = y2; h_a[:] = a2; h_b[:] = b2; h_d[:] = d2
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    io_bytes = h_y.nbytes + h_a.nbytes + h_b.nbytes + h_d.nbytes
    e2e = {"value": e_steps / e_dt, "unit": "sweeps/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": io_bytes}

    line = {"metric": "gibbs_marker_sweeps_per_sec", "value": value, "unit": "sweeps/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64 dots / f64 scalars / f32 state",
            "data": "synthetic", "config": dict(config, engine=args.engine, lag=args.lag, chain_ctas=args.chain_ctas, burnin=args.burnin, pi=("fixed 0.95" if args.fixed_pi else "estimated"),
                                                markers_in_model=model_size, active_updates_per_sweep=active,
                                                chain_rounds_per_sweep=rounds, setup_s=setup_s),
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary()}
    if not args.no_cpu:
        base, _ = cpu_reference_leg(n, p, args.cpu_markers, 0, 3, 1)
        line["cpu_baseline"] = base
    print(json.dumps(line))


if __name__ == "__main__":
    main()
