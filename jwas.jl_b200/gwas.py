"""GWAS post-processing of marker-effect samples (src/3.GWAS/src/GWAS.jl of the reference).

Two entry points, as in the reference:
  GWAS(marker_effects_file)                      -> model frequency per marker        (GWAS.jl:6-19)
  GWAS(model, map_file, *marker_effects_files)   -> window posterior probability of association (:49-196)
                                                    and window genomic correlations             (:197-243)

The window statistics need  var(X_w * alpha_w)  for every saved sample and window.  The reference
multiplies the dense n x p genotype matrix by every sample; here the product uses what the sampler
produces: a BayesC / BayesR sample has a few hundred non-zero effects out of p, so only those columns
are decoded from the 2-bit packed image (jwas_b200.Genotypes.packed) and accumulated per window --
O(n * nnz) per sample instead of O(n * p).  Host-side code: the samples are already on the host when
they are written (output.jl:467)."""
import math
import os

import numpy as np

try:
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None

from ._lib import JwasError


def error(msg):
    raise JwasError(msg)


def _read_samples(path, header=True):
    """MCMC_samples_marker_effects_<geno>_<trait>.txt: optional header of marker IDs, one row per sample."""
    with open(path) as f:
        first = f.readline().rstrip("\n")
        ids = [s.strip().strip('"') for s in first.split(",")] if header else None
        if not header:
            f.seek(0)
        rows = [np.array(line.split(","), dtype=np.float64) for line in f if line.strip()]
    samples = np.vstack(rows) if rows else np.zeros((0, len(ids) if ids else 0))
    if ids is None:
        ids = [str(i + 1) for i in range(samples.shape[1])]
    return samples, ids


def write_samples(path, marker_ids, rows):
    """output.jl:411 + :467: header line of marker IDs, then one comma-separated row per saved sample."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        f.write(",".join(str(m) for m in marker_ids) + "\n")
        for r in rows:
            f.write(",".join(repr(float(x)) for x in r) + "\n")


def _frame(cols):
    if pd is None:  # pragma: no cover
        return cols
    return pd.DataFrame(cols)


def model_frequency(marker_effects_file, header=True):
    """GWAS(marker_effects_file; header=true): share of samples in which each marker's effect is non-zero."""
    samples, ids = _read_samples(marker_effects_file, header)
    mf = (samples != 0.0).mean(axis=0) if len(samples) else np.zeros(len(ids))
    return _frame({"marker_ID": ids, "modelfrequency": mf})


def _windows(chr_, pos, window_size_bp, sliding_window):
    """GWAS.jl:94-137: windows per chromosome; empty windows are dropped (non-sliding).  The column cursor is
    advanced exactly as the reference does: by the window's marker count for non-overlapping windows, by one per
    window for sliding windows -- carried across chromosomes, so with sliding windows the column ranges of later
    chromosomes are the reference's (offset) ones, not that chromosome's own markers."""
    W = {k: [] for k in ("chr", "pos_start", "pos_end", "snp_start", "snp_end", "col_start", "col_end", "nsnp")}
    index_start = 0
    seen = []
    for c in chr_:
        if c not in seen:
            seen.append(c)
    for c in seen:
        pc = pos[chr_ == c]                                   # assumed sorted, as in the reference
        if not sliding_window:
            nwin = int(math.ceil(pc[-1] / window_size_bp))
        else:
            nwin = int(np.argmax(pc >= pc[-1] - window_size_bp)) + 1
        for j in range(nwin):
            start = window_size_bp * j if not sliding_window else int(pc[j])
            end = start + window_size_bp
            inw = (pc >= start) & (pc < end)
            k = int(inw.sum())
            if k != 0:
                first = int(np.argmax(inw)); last = len(inw) - 1 - int(np.argmax(inw[::-1]))
                W["snp_start"].append(int(pc[first])); W["snp_end"].append(int(pc[last]))
                W["col_start"].append(index_start); W["col_end"].append(index_start + k - 1)
                W["chr"].append(c); W["pos_start"].append(start); W["pos_end"].append(end); W["nsnp"].append(k)
            index_start += k if not sliding_window else 1
    return W


class _Columns:
    """Centred genotype columns decoded on demand from the packed image (missing -> 0 after centring)."""

    def __init__(self, geno):
        self.packed = geno.packed; self.n = geno.nObs; self.means = np.asarray(geno.marker_means, dtype=np.float32)
        self.cache = {}                      # bounded (LRU-ish): dense samples (RR-BLUP / BayesL) must not
        self.cache_cap = max(64, int(2 ** 28 // max(8 * self.n, 1)))   # materialise an n x p Float64 matrix (256 MB cap)

    def col(self, j):
        x = self.cache.get(j)
        if x is None:
            if len(self.cache) >= self.cache_cap:
                self.cache.pop(next(iter(self.cache)))
            b = self.packed[j]
            codes = np.stack([(b >> s) & 3 for s in (0, 2, 4, 6)], axis=1).reshape(-1)[:self.n]
            # Float32(code) - mean in Float32, 0 where missing: the centred Float32 genotypes of the reference
            # (decode_marker!, streaming_genotypes.jl:978-1002)
            x = np.where(codes == 3, np.float32(0), codes.astype(np.float32) - self.means[j]).astype(np.float64)
            self.cache[j] = x
        return x


def _var(x):
    return float(np.var(x, ddof=1)) if len(x) > 1 else float("nan")


def GWAS(model_or_file, map_file=None, *marker_effects_files, window_size="1 Mb", sliding_window=False,
         GWAS=True, threshold=0.001, genetic_correlation=False, local_EBV=False, header=True,
         output_winVarProps=False, write_files=False):
    """See the module docstring.  `model_or_file`: a samples file (model-frequency form) or the MME used in the
    analysis.  Returns a tuple of DataFrames (one per samples file, plus the correlation table), and the
    per-sample window variance proportions when output_winVarProps=True -- like the reference.  The reference
    always writes GWAS_<file> and MCMC_samples_local_genomic_variance<i>.txt into the working directory; here
    only with write_files=True."""
    if isinstance(model_or_file, (str, os.PathLike)) and map_file is None:
        return model_frequency(model_or_file, header=header)
    model = model_or_file
    if not marker_effects_files:
        error("GWAS: at least one marker effects file is required.")
    if isinstance(window_size, str):
        parts = window_size.split()
        if len(parts) != 2 or parts[1] != "Mb":
            error('The format for window_size is "1 Mb".')
    geno = model.M[0]
    snp_id = [str(m) for m in geno.markerID]
    if map_file is False and isinstance(window_size, (int, np.integer)):
        # GWAS.jl:67-76: fake map with window_size markers per 1 Mb window
        step = 1_000_000 / window_size
        ids = _read_samples(marker_effects_files[0], header)[1]
        mp_ids = [str(i) for i in ids]
        mp_chr = np.array(["1"] * len(ids)); mp_pos = np.floor(1 + step * np.arange(len(ids))).astype(np.int64)
        window_size = "1 Mb"
    else:
        if pd is None:  # pragma: no cover
            error("pandas is required to read the map file.")
        mp = pd.read_csv(map_file, header=0 if header else None, dtype={0: str, 1: str})
        mp_ids = [str(x) for x in mp.iloc[:, 0]]
        mp_chr = mp.iloc[:, 1].astype(str).to_numpy(); mp_pos = mp.iloc[:, 2].to_numpy(dtype=np.int64)
    window_size_bp = int(float(window_size.split()[0]) * 1_000_000)
    known = set(snp_id)
    keep = np.array([m in known for m in mp_ids])
    if not keep.any():
        error("Please check the 1st column of the mapfile (i.e., marker ID)")
    chr_, pos = mp_chr[keep], mp_pos[keep]
    W = _windows(chr_, pos, window_size_bp, sliding_window)
    nwin = len(W["nsnp"])
    cs = np.array(W["col_start"]); ce = np.array(W["col_end"])
    cols = _Columns(geno)
    n = geno.nObs

    def window_bvs(alpha):
        """{window: X_w alpha_w} for the windows that hold a non-zero effect, and the total X alpha."""
        nz = np.nonzero(alpha)[0]
        total = np.zeros(n)
        bv = {}
        for j in nz:
            xa = cols.col(int(j)) * alpha[j]
            total += xa
            for w in np.nonzero((cs <= j) & (j <= ce))[0]:        # one window, or several when sliding
                if w in bv:
                    bv[w] = bv[w] + xa
                else:
                    bv[w] = xa.copy()
        return bv, total

    out, props_out = [], []
    if GWAS:
        for fi, path in enumerate(marker_effects_files, start=1):
            samples, _ = _read_samples(path, header)
            ns = samples.shape[0]
            win_var = np.zeros((ns, nwin)); win_prop = np.zeros((ns, nwin))
            local = np.zeros((n, nwin)) if local_EBV else None
            for i in range(ns):
                bv, total = window_bvs(samples[i])
                gen_var = _var(total)
                for w, x in bv.items():
                    v = _var(x)
                    win_var[i, w] = v
                    win_prop[i, w] = v / gen_var if gen_var != 0 else float("nan")
                if gen_var == 0 or gen_var != gen_var:
                    win_prop[i, :] = float("nan")                 # 0/0 in the reference
                if local_EBV:
                    for w in range(nwin):
                        local[:, w] += ((bv[w] if w in bv else 0.0) - local[:, w]) / (i + 1)
            win_prop[np.isnan(win_prop)] = 0.0                    # GWAS.jl:174
            wppa = (win_prop > threshold).mean(axis=0) if ns else np.zeros(nwin)
            prop = np.round(win_prop.mean(axis=0) * 100, 6) if ns else np.zeros(nwin)
            vmean = win_var.mean(axis=0) if ns else np.zeros(nwin)
            vstd = win_var.std(axis=0, ddof=1) if ns > 1 else np.full(nwin, float("nan"))
            order = np.argsort(-wppa, kind="stable")              # sortperm(WPPA, rev=true)
            tab = _frame({
                "trait": [fi] * nwin, "window": (np.arange(nwin) + 1)[order],
                "chr": np.array(W["chr"])[order], "wStart": np.array(W["pos_start"])[order],
                "wEnd": np.array(W["pos_end"])[order], "start_SNP": np.array(W["snp_start"])[order],
                "end_SNP": np.array(W["snp_end"])[order], "numSNP": np.array(W["nsnp"])[order],
                "estimateGenVar": vmean[order], "stdGenVar": vstd[order], "prGenVar": prop[order],
                "WPPA": wppa[order], "PPA_t": np.cumsum(wppa[order]) / (np.arange(nwin) + 1)})
            out.append(tab)
            if write_files and local_EBV:                         # GWAS.jl: localEBV<i>.txt, one column per window
                np.savetxt(f"localEBV{fi}.txt", local, delimiter=",")
            if write_files and pd is not None:
                np.savetxt(f"MCMC_samples_local_genomic_variance{fi}.txt", win_var, delimiter=",")
                tab.to_csv("GWAS_" + str(path).replace("/", "_"), index=False)
            if output_winVarProps:
                props_out.append(win_prop)
    if genetic_correlation and len(marker_effects_files) == 2:
        s1, _ = _read_samples(marker_effects_files[0], header); s2, _ = _read_samples(marker_effects_files[1], header)
        ns = s1.shape[0]
        gcov = np.zeros((ns, nwin)); gcor = np.zeros((ns, nwin))
        for i in range(ns):
            b1, _ = window_bvs(s1[i]); b2, _ = window_bvs(s2[i])
            for w in set(b1) & set(b2):
                c = np.cov(b1[w], b2[w], ddof=1)
                gcov[i, w] = c[0, 1]
                d = math.sqrt(c[0, 0] * c[1, 1])
                gcor[i, w] = c[0, 1] / d if d > 0 else 0.0
        out.append(_frame({
            "trait": ["cor(t1,t2)"] * nwin, "window": np.arange(nwin) + 1, "chr": W["chr"],
            "wStart": W["pos_start"], "wEnd": W["pos_end"], "start_SNP": W["snp_start"], "end_SNP": W["snp_end"],
            "numSNP": W["nsnp"], "estimate_cov": gcov.mean(axis=0), "std_cov": gcov.std(axis=0, ddof=1) if ns > 1 else np.nan,
            "estimate_cor": gcor.mean(axis=0), "std_cor": gcor.std(axis=0, ddof=1) if ns > 1 else np.nan}))
    return (tuple(out), tuple(props_out)) if output_winVarProps else tuple(out)
