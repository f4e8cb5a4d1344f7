"""Ingest timing at the size the reference documents for its packed backend: N = 10,000 individuals x P = 5,000 markers,
`prepare_streaming_genotypes` = 11.99 s (docs/src/manual/streaming_genotype_backend.md:178-182).  Host only (no GPU):
text file -> .jgb2 + side-cars through libjwasio (include/jwas_io.h).

    python tools/ingest_bench.py [--n 10000 --p 5000]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jwas_b200 as jw  # noqa: E402
from jwas_b200 import _io  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--p", type=int, default=5000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    d = tempfile.mkdtemp()
    path = os.path.join(d, "geno.csv")
    f = rng.uniform(0.05, 0.5, a.p)
    codes = (rng.random((a.n, a.p)) < f).astype(np.int8) + (rng.random((a.n, a.p)) < f).astype(np.int8)
    with open(path, "w") as fh:
        fh.write("ID," + ",".join(f"m{j + 1}" for j in range(a.p)) + "\n")
        for i in range(a.n):
            fh.write(f"id_{i}," + ",".join(map(str, codes[i].tolist())) + "\n")
    size_mb = os.path.getsize(path) / 1e6
    res = {"n": a.n, "p": a.p, "file_MB": round(size_mb, 1), "cores": os.cpu_count()}
    for th in (1, 0):
        best = 1e9
        for _ in range(a.reps):
            t0 = time.perf_counter(); _io.read_genotype_text(path, nthreads=th); best = min(best, time.perf_counter() - t0)
        res["parse_pack_s_%s" % ("1_thread" if th == 1 else "all_cores")] = round(best, 3)
    best = 1e9
    for _ in range(a.reps):
        t0 = time.perf_counter(); prefix = jw.prepare_streaming_genotypes(path); best = min(best, time.perf_counter() - t0)
    res["prepare_streaming_genotypes_s"] = round(best, 3)
    t0 = time.perf_counter(); g = jw.get_genotypes(prefix, 1.0); res["get_genotypes_from_jgb2_s"] = round(time.perf_counter() - t0, 3)
    res["markers_after_qc"] = g.nMarkers
    res["reference_prepare_s_published"] = 11.99
    res["MB_per_s_all_cores"] = round(size_mb / res["parse_pack_s_all_cores"], 0)
    shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
