"""Chain-level comparison with numbers the reference PUBLISHED for this path: its `simulated_annotations` benchmark
(benchmarks/simulated_annotations_multitrait_comparison.jl, standard mode) on its own packaged fixture
(src/4.Datasets/data/simulated_annotations: 400 individuals x 964 markers, two traits, 14 active markers per trait),
report benchmarks/reports/2026-04-10-multitrait-annotated-bayesc-sampler-benchmark-report.md (seeds 101 and 202,
chain_length 2000, burnin 500, output_samples_frequency 20, starting h2 0.5).

Same settings here through jwas_b200.get_genotypes / build_model / runMCMC: Pi, starting variances, annotations
(annotations_mt.csv, four columns), outputEBV on the phenotyped IDs, in-sample metrics exactly as summarize_case
computes them (:816-890): cor(y, EBV), cor(estimate, true effect), top-k recall of the model frequencies with
k = number of active markers, any-active recall for the two-trait runs.  The random streams differ (Julia's
Xoshiro vs Philox), so agreement is statistical: the reference's two-seed means should lie inside the seed-to-seed
spread measured here.  Differences in set-up: this backend centres the genotypes (the benchmark passes center=false;
with a flat-prior intercept the marker-effect posterior is the same), and annotated 2-trait BayesC runs sampler I only.

Script, not a test (reads /root/reference; minutes of CPU).  Default backend: the CPU oracle backend (the same host
code over the bit-exact twin of the CUDA sweep); `--backend gpu` runs the B200 library instead.

    python tests/ref_benchmark_annotations.py --seeds 101,202,303,404,505,606 --out profiles/r2_annotations_benchmark.md
"""
import argparse
import os
import sys
import time

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import jwas_b200 as jw  # noqa: E402

MT_START_PI = {(0.0, 0.0): 0.96, (1.0, 0.0): 0.015, (0.0, 1.0): 0.015, (1.0, 1.0): 0.01}
ST_BAYESC_PI = 0.98
ST_BAYESR_PI = [0.99, 0.006, 0.003, 0.001]

# variant, method, annotated, multitrait, trait, sampler
CASES = [
    ("MT_BayesC", "BayesC", False, True, "", "I"),
    ("MT_BayesC_II", "BayesC", False, True, "", "II"),
    ("MT_Annotated_BayesC_I", "BayesC", True, True, "", "I"),
    ("BayesC_y1", "BayesC", False, False, "y1", "I"),
    ("Annotated_BayesC_y1", "BayesC", True, False, "y1", "I"),
    ("BayesC_y2", "BayesC", False, False, "y2", "I"),
    ("Annotated_BayesC_y2", "BayesC", True, False, "y2", "I"),
    ("BayesR_y1", "BayesR", False, False, "y1", "I"),
    ("BayesR_y2", "BayesR", False, False, "y2", "I"),
    ("Annotated_BayesR_y1", "BayesR", True, False, "y1", "I"),
    ("Annotated_BayesR_y2", "BayesR", True, False, "y2", "I"),
]

# (variant, trait) -> cor(y, EBV), effect corr, top-k recall, any-active recall: the reference's two-seed means,
# benchmarks/reports/2026-04-10-multitrait-annotated-bayesc-sampler-benchmark-report.md:27-42 (MT_BayesC_II from
# 2026-04-11-multitrait-bayesc-plain-empty-sampler-report.md when present there; otherwise not published)
PUBLISHED = {
    ("MT_BayesC", "y1"): (0.7731, 0.5057, 0.2857, 0.3000), ("MT_BayesC", "y2"): (0.7353, 0.7573, 0.3571, 0.3000),
    ("MT_Annotated_BayesC_I", "y1"): (0.7607, 0.4970, 0.3571, 0.4250),
    ("MT_Annotated_BayesC_I", "y2"): (0.7146, 0.7725, 0.4643, 0.4250),
    ("BayesC_y1", "y1"): (0.7835, 0.4723, 0.2500, None), ("Annotated_BayesC_y1", "y1"): (0.7580, 0.5403, 0.3214, None),
    ("BayesC_y2", "y2"): (0.7279, 0.7799, 0.3571, None), ("Annotated_BayesC_y2", "y2"): (0.7069, 0.8396, 0.6071, None),
    ("BayesR_y1", "y1"): (0.8029, 0.4752, 0.2500, None), ("BayesR_y2", "y2"): (0.7468, 0.8085, 0.3929, None),
    ("Annotated_BayesR_y1", "y1"): (0.7805, 0.5942, 0.2500, None),
    ("Annotated_BayesR_y2", "y2"): (0.7320, 0.8620, 0.5357, None),
}


def top_k_recall(scores, truth, k):
    order = np.argsort(-np.asarray(scores, float), kind="stable")[:k]
    return float(np.sum(truth[order])) / k


# family -> trait-mean held-out cor(y, EBV), 5 folds x seeds 101 and 202, chain_length 1500, burnin 500, frequency 50:
# benchmarks/reports/2026-04-11-simulated-annotations-cv-report.md:62-91
PUBLISHED_CV = {"MT_BayesC": 0.6397, "MT_BayesC_II": 0.6423, "MT_Annotated_BayesC_I": 0.6338, "BayesC_single": 0.6424,
                "Annotated_BayesC_single": 0.6469, "BayesR_single": 0.6497, "Annotated_BayesR_single": 0.6484}


def family(variant):
    return variant[:-3] + "_single" if variant.endswith(("_y1", "_y2")) else variant


def run_case(case, seed, data, args, factory, heldout=None):
    """heldout: IDs whose phenotypes are hidden (cross-validation mode): they are left out of the training rows --
    what masking y (single-trait) or y1 and y2 (two-trait) does in the reference (masked_phenotype_frame, :156-169) --
    and predicted through outputEBV on all IDs.  Starting variances come from all phenotypes, as there (:725, :749)."""
    variant, method, annotated, multitrait, trait, sampler = case
    geno_df, ph_all, A, truth = data
    ph = ph_all if heldout is None else ph_all[~ph_all["ID"].isin(heldout)].reset_index(drop=True)
    kw = dict(method=method, estimatePi=True, quality_control=False, center=not args.uncentred)
    if annotated:
        kw["annotations"] = A
    if multitrait:
        cov = np.cov(ph_all[["y1", "y2"]].to_numpy(float).T)
        geno = jw.get_genotypes(geno_df, cov * args.start_h2, Pi=dict(MT_START_PI), multi_trait_sampler=sampler, **kw)
        model = jw.build_model("y1 = intercept + bench_geno\ny2 = intercept + bench_geno", cov * (1 - args.start_h2),
                               genotypes={"bench_geno": geno})
        traits = ["y1", "y2"]
    else:
        v = float(np.var(ph_all[trait].to_numpy(float), ddof=1))
        Pi = ST_BAYESC_PI if method == "BayesC" else list(ST_BAYESR_PI)
        geno = jw.get_genotypes(geno_df, v * args.start_h2, Pi=Pi, **kw)
        model = jw.build_model(f"{trait} = intercept + bench_geno", v * (1 - args.start_h2), genotypes={"bench_geno": geno})
        traits = [trait]
    jw.outputEBV(model, list(ph_all["ID"]))
    t0 = time.time()
    extra = dict(_backend_factory=factory, lag=0, panel=args.panel) if factory is not None else {}
    out = jw.runMCMC(model, ph, chain_length=args.chain_length, burnin=args.burnin,
                     output_samples_frequency=args.freq, seed=seed, outputEBV=True, output_heritability=False, **extra)
    dt = time.time() - t0
    rows = []
    if heldout is not None:                      # summarize_case_cv (:892-903): held-out cor(y, EBV) per trait
        test = ph_all[ph_all["ID"].isin(heldout)]
        for tr in traits:
            ebv = out["EBV_" + tr].set_index("ID").loc[test["ID"]]["EBV"].to_numpy(float)
            rows.append(dict(variant=variant, family=family(variant), trait=tr, seed=seed, runtime_s=dt,
                             heldout_cor=float(np.corrcoef(ebv, test[tr].to_numpy(float))[0, 1])))
        return rows
    me = out["marker effects bench_geno"]
    pips = {}
    for tr in traits:
        sub = me[me["Trait"] == tr].set_index("Marker_ID").loc[truth["marker_id"]]
        active = truth[f"is_active_{tr}"].to_numpy(bool)
        ebv = out["EBV_" + tr].set_index("ID").loc[ph["ID"]]["EBV"].to_numpy(float)
        pips[tr] = sub["Model_Frequency"].to_numpy(float)
        rows.append(dict(variant=variant, trait=tr, seed=seed, runtime_s=dt,
                         ebv_cor=float(np.corrcoef(ebv, ph[tr].to_numpy(float))[0, 1]),
                         effect_cor=float(np.corrcoef(sub["Estimate"].to_numpy(float),
                                                      truth[f"true_effect_{tr}"].to_numpy(float))[0, 1]),
                         topk=top_k_recall(pips[tr], active, int(active.sum())), any_active=np.nan))
    if multitrait:
        any_active = truth["is_active_y1"].to_numpy(bool) | truth["is_active_y2"].to_numpy(bool)
        rec = top_k_recall(np.maximum(pips["y1"], pips["y2"]), any_active, int(any_active.sum()))
        for r in rows:
            r["any_active"] = rec
    if annotated:
        co = out["annotation coefficients bench_geno"]
        for r in rows:
            r["annotation_coefficients"] = co.to_dict("records")
    return rows


# cross-seed agreement of two runs (seeds 100 and 110; chain_length 5000, burnin 1000, frequency 10) on the single-trait
# fixture: benchmarks/simulated_annotations_method_matrix.jl, reports/2026-04-02-simulated-annotations-method-matrix-report.md
# variant -> marker corr, PIP corr, EBV corr, annotation coefficient corr, pi vector corr
PUBLISHED_MATRIX = {"BayesC_dense": (0.8502, 0.6409, 0.9813, None, None),
                    "Annotated_BayesC_dense": (0.9911, 0.9932, 0.9997, 0.9956, 0.99997),
                    "BayesR_dense": (0.9760, 0.8255, 0.9995, None, 0.99992),
                    "Annotated_BayesR_dense": (0.9985, 0.9982, 0.9997, 0.9844, 0.999996)}


def matrix_mode(args, factory):
    """simulated_annotations_method_matrix.jl:78-198 (dense variants): BayesC starts from Pi = 0, BayesR from
    (0.99, 0.006, 0.003, 0.001); every pair of seeds gives one set of cross-seed correlations (summarize_pair)."""
    import itertools
    geno_df = pd.read_csv(os.path.join(args.data, "genotypes.csv"))
    ph = pd.read_csv(os.path.join(args.data, "phenotypes.csv"))
    A = pd.read_csv(os.path.join(args.data, "annotations.csv"))[["functional", "random_anno"]].to_numpy(float)
    v = float(np.var(ph["y1"].to_numpy(float), ddof=1))
    seeds = [int(s) for s in args.seeds.split(",")]
    extra = dict(_backend_factory=factory, lag=0, panel=args.panel) if factory is not None else {}
    rows = []
    # Annotated BayesC from Pi = 0: the source under /root/reference starts the probit intercept at
    # quantile(Normal, 1 - eps) = 8.13 (annotation_setup.jl:64-71), which leaves every marker in the model until the
    # intercept has random-walked down -- a metastable start whose exit time depends on the seed.  The report was
    # produced on 2026-04-02 with the start-up of docs/plans/2026-04-02-annotated-bayesc-jian-startup-design.md
    # (coefficients and mu at zero), so that start-up is run as well.
    for variant, method, annotated, zero_start in (("BayesC_dense", "BayesC", False, False),
                                                   ("Annotated_BayesC_dense", "BayesC", True, False),
                                                   ("Annotated_BayesC_dense (report-time start-up: coefficients 0)", "BayesC", True, True),
                                                   ("BayesR_dense", "BayesR", False, False),
                                                   ("Annotated_BayesR_dense", "BayesR", True, False)):
        runs = {}
        for seed in seeds:
            kw = dict(method=method, estimatePi=True, quality_control=False, center=not args.uncentred,
                      Pi=(0.0 if method == "BayesC" else list(ST_BAYESR_PI)))
            if annotated:
                kw["annotations"] = A
            geno = jw.get_genotypes(geno_df, v * args.start_h2, **kw)
            model = jw.build_model("y1 = intercept + geno", v * (1 - args.start_h2), genotypes={"geno": geno})
            jw.outputEBV(model, list(geno_df.iloc[:, 0]))
            if zero_start:
                geno.annotations.coefficients[:] = 0.0
                geno.annotations.mu[:] = 0.0
            t0 = time.time()
            out = jw.runMCMC(model, ph, chain_length=args.chain_length, burnin=args.burnin, output_samples_frequency=args.freq,
                             seed=seed, outputEBV=True, output_heritability=False, **extra)
            me = out["marker effects geno"]
            pi = out["pi_geno"]["Estimate"].to_numpy(float)
            runs[seed] = dict(est=me["Estimate"].to_numpy(float), pip=me["Model_Frequency"].to_numpy(float),
                              ebv=out["EBV_y1"]["EBV"].to_numpy(float),
                              ann=(out["annotation coefficients geno"]["Estimate"].to_numpy(float) if annotated else None),
                              pi=(pi if len(pi) > 1 else None))
            print(f"{variant:24s} seed {seed}: {time.time() - t0:.0f} s", flush=True)

        def cor(a, b):
            return float(np.corrcoef(a, b)[0, 1]) if a is not None and np.std(a) > 0 and np.std(b) > 0 else np.nan
        for sa, sb in itertools.combinations(seeds, 2):
            ra, rb = runs[sa], runs[sb]
            rows.append(dict(variant=variant, seed_a=sa, seed_b=sb, marker=cor(ra["est"], rb["est"]), pip=cor(ra["pip"], rb["pip"]),
                             ebv=cor(ra["ebv"], rb["ebv"]), ann=cor(ra["ann"], rb["ann"]), pi=cor(ra["pi"], rb["pi"])))
    df = pd.DataFrame(rows)
    lines = ["| Variant | seed pairs | Marker corr: here min / mean / max (reference) | PIP corr | EBV corr | Annotation coeff corr | pi vector corr |",
             "|---|---|---|---|---|---|---|"]
    for variant, g in df.groupby("variant", sort=False):
        ref = PUBLISHED_MATRIX[variant.split(" (")[0]]
        cells = []
        for col, r in zip(("marker", "pip", "ebv", "ann", "pi"), ref):
            x = g[col].to_numpy(float)
            if np.all(np.isnan(x)):
                cells.append("—")
            else:
                cells.append(f"{np.nanmin(x):.4f} / {np.nanmean(x):.4f} / {np.nanmax(x):.4f} ({'%.4f' % r if r is not None else 'NA'})")
        lines.append(f"| `{variant}`{' (center=false)' if args.uncentred else ''} | {len(g)} | " + " | ".join(cells) + " |")
    table = "\n".join(lines)
    print(table)
    if args.out:
        with open(args.out, "w") as f:
            f.write(f"<!-- python tests/ref_benchmark_annotations.py --matrix --seeds {args.seeds} --chain-length {args.chain_length} "
                    f"--burnin {args.burnin} --freq {args.freq} --backend {args.backend} -->\n" + table + "\n")
        df.to_csv(os.path.splitext(args.out)[0] + "_runs.csv", index=False)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default="/root/reference/src/4.Datasets/data/simulated_annotations")
    ap.add_argument("--seeds", default="101,202")
    ap.add_argument("--chain-length", type=int, default=2000)
    ap.add_argument("--burnin", type=int, default=500)
    ap.add_argument("--freq", type=int, default=20)
    ap.add_argument("--start-h2", type=float, default=0.5)
    ap.add_argument("--panel", type=int, default=16, help="oracle backend only: small panels keep its Gram work small")
    ap.add_argument("--backend", choices=["oracle", "gpu"], default="oracle")
    ap.add_argument("--variants", default="")
    ap.add_argument("--matrix", action="store_true", help="cross-seed method matrix on the single-trait fixture "
                    "(the reference's report: --seeds 100,110 --chain-length 5000 --burnin 1000 --freq 10)")
    ap.add_argument("--uncentred", action="store_true", help="center=false, as the reference's benchmark scripts pass it")
    ap.add_argument("--cv", type=int, default=0, help="K-fold cross-validation mode (the reference's report: 5 folds, "
                    "--chain-length 1500 --burnin 500 --freq 50, seeds 101,202)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    factory = None
    if args.backend == "oracle":
        from oracle_backend import factory
    if args.matrix:
        return matrix_mode(args, factory)
    geno_df = pd.read_csv(os.path.join(args.data, "genotypes.csv"))
    ph = pd.read_csv(os.path.join(args.data, "phenotypes_mt.csv"))
    ann = pd.read_csv(os.path.join(args.data, "annotations_mt.csv"))
    truth = pd.read_csv(os.path.join(args.data, "truth_mt.csv"))
    for c in ("is_active_y1", "is_active_y2"):
        truth[c] = truth[c].astype(str).str.upper().isin(["TRUE", "1"])
    assert list(ann["marker_id"]) == list(geno_df.columns[1:]) == list(truth["marker_id"])
    data = (geno_df, ph, ann.iloc[:, 1:].to_numpy(float), truth)
    seeds = [int(s) for s in args.seeds.split(",")]
    want = set(v for v in args.variants.split(",") if v)
    rows = []
    if args.cv:
        # cv_fold_assignments (:142-154): shuffle the sorted IDs, deal them round-robin into the folds; the same folds for
        # every method within a seed
        for seed in seeds:
            rng = np.random.default_rng(seed)
            shuffled = rng.permutation(sorted(ph["ID"]))
            fold_of = {i: k % args.cv for k, i in enumerate(shuffled)}
            for fold in range(args.cv):
                held = [i for i in ph["ID"] if fold_of[i] == fold]
                for case in CASES:
                    if want and case[0] not in want:
                        continue
                    r = run_case(case, seed, data, args, factory, heldout=held)
                    for x in r:
                        x["fold"] = fold
                        print(f"{x['variant']:24s} {x['trait']} seed {seed} fold {fold}: held-out cor {x['heldout_cor']:.4f} "
                              f"({x['runtime_s']:.0f} s)", flush=True)
                    rows += r
        df = pd.DataFrame(rows)
        lines = ["| Family | Trait-mean held-out cor(y, EBV): here, mean over seeds x folds x traits (sd of the seed means; "
                 "se over folds) | reference |", "|---|---|---|"]
        for fam, g in df.groupby("family", sort=False):
            seed_means = g.groupby("seed")["heldout_cor"].mean().to_numpy()
            se = g["heldout_cor"].std(ddof=1) / np.sqrt(len(g))
            ref = PUBLISHED_CV.get(fam)
            lines.append(f"| `{fam}` | {g['heldout_cor'].mean():.4f} (sd {seed_means.std(ddof=1) if len(seed_means) > 1 else 0:.4f}; "
                         f"se {se:.4f}) | {'%.4f' % ref if ref is not None else 'not published'} |")
        table = "\n".join(lines)
        print(table)
        if args.out:
            with open(args.out, "w") as f:
                f.write(f"<!-- python tests/ref_benchmark_annotations.py --cv {args.cv} --seeds {args.seeds} --chain-length "
                        f"{args.chain_length} --burnin {args.burnin} --freq {args.freq} --backend {args.backend} -->\n")
                f.write(table + "\n")
            df.to_csv(os.path.splitext(args.out)[0] + "_runs.csv", index=False)
        return 0
    for case in CASES:
        if want and case[0] not in want:
            continue
        for seed in seeds:
            r = run_case(case, seed, data, args, factory)
            rows += r
            for x in r:
                print(f"{x['variant']:24s} {x['trait']} seed {seed}: cor(y,EBV) {x['ebv_cor']:.4f}  effect {x['effect_cor']:.4f}  "
                      f"top-k {x['topk']:.4f}  any {x['any_active']:.4f}  ({x['runtime_s']:.0f} s)", flush=True)
    df = pd.DataFrame(rows)
    lines = ["| Variant | Trait | cor(y, EBV): here mean ± sd (reference) | Effect corr: here (reference) | "
             "Top-k recall: here (reference) | Any-active recall: here (reference) |", "|---|---|---|---|---|---|"]

    def cell(g, col, ref):
        v = g[col].to_numpy(float)
        if np.all(np.isnan(v)):
            return "—"
        s = f"{np.nanmean(v):.4f} ± {np.nanstd(v, ddof=1) if len(v) > 1 else 0.0:.4f}"
        return s + (f" ({ref:.4f})" if ref is not None else " (not published)")

    for (variant, trait), g in df.groupby(["variant", "trait"], sort=False):
        ref = PUBLISHED.get((variant, trait), (None,) * 4)
        lines.append(f"| `{variant}` | {trait} | {cell(g, 'ebv_cor', ref[0])} | {cell(g, 'effect_cor', ref[1])} | "
                     f"{cell(g, 'topk', ref[2])} | {cell(g, 'any_active', ref[3])} |")
    table = "\n".join(lines)
    print(table)
    if args.out:
        with open(args.out, "w") as f:
            f.write(f"<!-- python tests/ref_benchmark_annotations.py --seeds {args.seeds} --chain-length {args.chain_length} "
                    f"--burnin {args.burnin} --freq {args.freq} --backend {args.backend} -->\n")
            f.write(table + "\n")
        df.drop(columns=[c for c in ("annotation_coefficients",) if c in df]).to_csv(
            os.path.splitext(args.out)[0] + "_runs.csv", index=False)
    return 0


if __name__ == "__main__":
    sys.exit(main())
