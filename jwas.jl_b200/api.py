"""Host-side mirror of the reference interface for the marker-effects path.

Same names, argument meaning and error behaviour as JWAS.jl for the calls a user makes around the
marker sweep:  get_genotypes -> build_model -> (set_covariate/set_random) -> runMCMC, with the
Genotypes / MME / MCMCinfo / Variance types (types.jl:56-64, 98-165, 225-248, 264-346).  The
backend seam is the one the reference itself uses for `storage=:stream`
(readgenotypes.jl:228-295, MCMC_BayesianAlphabet.jl:53-65, 243-251): `storage="gpu"` keeps the
genotypes 2-bit packed in HBM behind a libjwasb200 handle and every sweep runs there.

Only what the sweep needs is implemented: `y = intercept + <genotypes>` models (single- or
multi-trait), BayesA/B/C, BayesR and BayesL (single-trait), RR-BLUP (single- and multi-trait), multi-trait BayesC
samplers I / II, and the
annotation-aware priors of BayesC / BayesR / 2-trait BayesC (annotations.py).
Everything else the reference offers (pedigree, covariates, random terms, SEM, RRM, categorical traits,
GBLUP) is outside this backend's scope and raises JwasError with a message saying so.
"""
import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import _io
from . import annotations as annot
from . import mcmc
from .memory import check_marker_memory_guard, device_memory_bytes, estimate_marker_memory, format_bytes_human
from ._lib import GpuSweeper, JwasError, SCHED_EXACT, SCHED_BLOCK, SCHED_INDEPENDENT

try:  # pandas is what a DataFrame is here
    import pandas as pd
except Exception:  # pragma: no cover
    pd = None

DEFAULT_PANEL = 4096          # look-ahead panel of the exact schedule (GPU-internal)
DEFAULT_CHAIN_CTAS = 6        # chain CTAs of the pipelined chain (engine 1, lag >= 1); 0 = one chain CTA
DEFAULT_LAG = 2               # lagged exact schedule: three look-ahead panels in flight (0 = plain, 1, 2)


def error(msg):
    raise JwasError(msg)


# ------------------------------------------------------------------------------------ types
@dataclass
class Variance:                      # types.jl:56-64
    val: object = False
    df: float = 4.0
    scale: object = False
    estimate_variance: bool = True
    estimate_scale: bool = False
    constraint: bool = False


@dataclass
class Genotypes:                     # types.jl:98-165 (fields this path uses)
    name: str = ""
    obsID: list = field(default_factory=list)
    markerID: list = field(default_factory=list)
    nObs: int = 0
    nMarkers: int = 0
    alleleFreq: np.ndarray = None
    sum2pq: float = 0.0
    centered: bool = True
    packed: np.ndarray = None        # marker-major 2-bit image (.jgb2 layout), host copy until runMCMC
    marker_means: np.ndarray = None
    method: str = "BayesC"
    estimatePi: bool = True
    π: object = 0.0
    genetic_variance: Variance = field(default_factory=Variance)
    G: Variance = field(default_factory=Variance)
    ntraits: int = 1
    α: list = None
    β: list = None
    δ: list = None
    meanAlpha: list = None
    meanAlpha2: list = None
    meanDelta: list = None
    mean_pi: object = None
    multi_trait_sampler: str = "I"
    storage_mode: str = "gpu"
    stream_backend: object = None    # GpuSweeper once the chain starts
    starting_value: object = False
    annotations: object = False      # annotations.MarkerAnnotations (types.jl:167-216)
    annotation_start_pi: object = 0.0
    nMarkersAll: int = 0             # markers in the raw file and the (1-based) raw indices of the ones kept by QC:
    selected_marker_indices: np.ndarray = None     # the raw-marker mapping of the packed backend (streaming_genotypes.jl:913-930)


@dataclass
class MCMCinfo:                      # types.jl:225-248
    chain_length: int = 100
    burnin: int = 0
    output_samples_frequency: int = 1
    seed: object = False
    fast_blocks: object = False
    independent_blocks: bool = False
    outputEBV: bool = True
    double_precision: bool = False
    output_folder: str = "results"
    printout_model_info: bool = False


@dataclass
class MME:                           # types.jl:264-346 (fields this path uses)
    model_equations: str = ""
    lhsVec: list = field(default_factory=list)
    nModels: int = 1
    M: list = field(default_factory=list)
    R: Variance = field(default_factory=Variance)
    MCMCinfo: MCMCinfo = None
    obsID: list = None
    output_ID: object = False        # outputEBV(model, IDs) (output.jl:60-69); default: all genotyped (check_outputID)
    sol: np.ndarray = None
    output: dict = None


# ------------------------------------------------------------------------------------ codec / QC
def _pack_codes(codes):
    """(n, p) codes 0/1/2 with 3 = missing -> (p, cld(n,4)) uint8, LSB first
    (streaming_genotypes.jl:622-627)."""
    n, p = codes.shape
    stride = (n + 3) // 4
    padded = np.zeros((stride * 4, p), dtype=np.uint8)
    padded[:n] = codes
    q = padded.reshape(stride, 4, p)
    packed = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)
    return np.ascontiguousarray(packed.T)


def _unpack_codes(packed, n):
    p, stride = packed.shape
    out = np.empty((stride * 4, p), dtype=np.uint8)
    for k in range(4):
        out[k::4] = ((packed >> (2 * k)) & 3).T
    return out[:n]


def _codes_from_matrix(X, missing_value):
    X = np.asarray(X)
    miss = (X == missing_value) | ~np.isfinite(X.astype(np.float64))
    Xr = np.where(miss, 0, X)
    if not np.all((Xr == 0) | (Xr == 1) | (Xr == 2)):
        error(f"Only 0/1/2 genotypes (and missing_value={missing_value}) are supported in storage=:gpu.")
    codes = Xr.astype(np.uint8)
    codes[miss] = 3
    return codes


def _call_counts(packed, n):
    """(p, 3) int64: 1s, 2s and missing calls per marker, from the packed image (libjwasio when built)."""
    if _io.available():
        return _io.packed_counts(packed, n)
    codes = _unpack_codes(packed, n)
    return np.stack([(codes == 1).sum(axis=0), (codes == 2).sum(axis=0), (codes == 3).sum(axis=0)], axis=1).astype(np.int64)


def _select_rows(packed, n, rows):
    """Rows `rows` of every packed column as a new image (aligning genotypes to other individuals, JWAS.jl:381-402)."""
    if _io.available():
        return _io.packed_rows(packed, rows)
    return _pack_codes(_unpack_codes(packed, n)[np.asarray(rows)])


def _stats_from_counts(counts, n):
    """Means over observed calls (Float32 like the reference, readgenotypes.jl:372-385, streaming_genotypes.jl:560-585),
    number of observed calls, their sum and their sum of squares."""
    n1, n2, nm = counts[:, 0], counts[:, 1], counts[:, 2]
    nn = n - nm
    if np.any(nn == 0):
        error("Marker %d has only missing values." % int(np.argmin(nn) + 1))
    s = n1 + 2 * n2
    means = (s.astype(np.float32) / nn.astype(np.float32)).astype(np.float32)
    return means, nn, s, (n1 + 4 * n2)


def get_genotypes(file, G=False, *, method="BayesC", Pi=0.0, estimatePi=True, G_is_marker_variance=False,
                  df=4.0, estimate_variance=True, estimate_scale=False, constraint=False, separator=",",
                  header=True, double_precision=False, quality_control=True, MAF=0.01, missing_value=9.0,
                  center=True, starting_value=False, annotations=False, multi_trait_sampler="I",
                  storage="gpu", name="geno", obsID=None, markerID=None):
    """readgenotypes.jl:213-448 for the GPU backend.

    `file`: (n, p) array of 0/1/2 (missing_value for missing), a DataFrame whose first column holds
    IDs, a CSV path, or the prefix of a packed `.jgb2` backend written by
    prepare_streaming_genotypes (same files the reference's storage=:stream reads)."""
    if multi_trait_sampler not in ("auto", "I", "II"):
        error("multi_trait_sampler must be one of :auto, :I, or :II.")
    if storage not in ("gpu",):
        error("storage must be :gpu in this backend (:dense and :stream live in JWAS.jl).")
    if method not in ("BayesA", "BayesB", "BayesC", "BayesR", "RR-BLUP", "BayesL"):
        error(f"method {method} is outside the GPU marker-sweep path (BayesA/B/C, BayesR, RR-BLUP and BayesL only).")
    if double_precision:
        error("double_precision=true is not supported with storage=:gpu.")
    if estimate_scale:
        error("estimate_scale=true is not supported with storage=:gpu.")
    if method == "BayesR" and not isinstance(Pi, (list, tuple, np.ndarray)) and Pi != 0.0:
        error("BayesR Pi must have length 4.")

    be = None
    if isinstance(file, str) and (os.path.exists(file + ".meta") or file.endswith((".jgb2", ".meta"))):
        be = load_streaming_backend(file)
        packed, n = be["packed"], be["nObs"]
        obs, mk = be["obsID"], be["markerID"]
        quality_control = False            # QC happened when the backend was prepared
        if annotations is not False and not be["has_raw_marker_mapping"]:      # readgenotypes.jl:251-253
            error("Annotated storage=:stream requires a backend prepared with the current prepare_streaming_genotypes; "
                  "rebuild the backend to include raw-marker mapping metadata.")
    elif isinstance(file, str) and _io.available() and len(separator) == 1:
        # text file -> 2-bit image in one parallel pass, no dense matrix (libjwasio, include/jwas_io.h)
        if not os.path.isfile(file):
            error(f"genotype file {file} is not found.")
        obs, mk, packed = _io.read_genotype_text(file, separator, header, missing_value)
        n = len(obs)
    else:
        if isinstance(file, str):
            if pd is None:
                error("pandas is required to read genotype files.")
            dfm = pd.read_csv(file, sep=separator, header=0 if header else None)
            obs = [str(x) for x in dfm.iloc[:, 0]]
            mk = [str(c) for c in dfm.columns[1:]] if header else [f"m{j + 1}" for j in range(dfm.shape[1] - 1)]
            X = dfm.iloc[:, 1:].to_numpy()
        elif pd is not None and isinstance(file, pd.DataFrame):
            obs = [str(x) for x in file.iloc[:, 0]]
            mk = [str(c) for c in file.columns[1:]]
            X = file.iloc[:, 1:].to_numpy()
        else:
            X = np.asarray(file)
            if X.ndim != 2:
                error("genotypes must be a matrix.")
            obs = [str(x) for x in obsID] if obsID is not None else [str(i + 1) for i in range(X.shape[0])]
            mk = [str(x) for x in markerID] if markerID is not None else [f"m{j + 1}" for j in range(X.shape[1])]
        if X.size == 0:
            error("Genotype data is empty.")
        codes = _codes_from_matrix(X, missing_value)
        packed, n = _pack_codes(codes), codes.shape[0]
        del codes

    from_backend = be is not None
    n_all = be["nMarkersAll"] if from_backend else packed.shape[0]
    selected = be["selected_marker_indices"] if from_backend else np.arange(1, packed.shape[0] + 1, dtype=np.int32)
    try:                                    # readgenotypes.jl:254-258: one annotation row per RAW marker
        ann_matrix = annot.validate_annotations_input(annotations, n_all, method)
    except annot.AnnotationError as e:
        error(str(e))
    if ann_matrix is not False and from_backend:
        ann_matrix = ann_matrix[np.asarray(selected, dtype=np.int64) - 1]      # rows of the markers the backend kept
    if ann_matrix is not False and not estimatePi:
        import warnings
        warnings.warn(f"estimatePi=false is ignored when annotations are provided; Annotated {method} requires "
                      "estimatePi=true.")
        estimatePi = True
    # per-marker statistics from the call counts of the packed image (never from a dense matrix)
    counts = _call_counts(packed, n)
    means, nn, s, sq = _stats_from_counts(counts, n)
    centered = bool(center) if be is None else be["centered"]     # a prepared backend keeps its own flag (readgenotypes.jl:259-262)
    if not centered and np.any(counts[:, 2] > 0):
        # uncentred columns keep the column mean at missing calls (decode_marker!, streaming_genotypes.jl:993-994); the
        # device image has no value for a missing call other than "the mean, i.e. 0 after centring"
        error("center=false with missing genotypes is not supported with storage=:gpu.")
    af = (means / np.float32(2.0)).astype(np.float32)
    if quality_control:                     # readgenotypes.jl:388-399: MAF filter + fixed loci
        ss = sq.astype(np.float64) - (s.astype(np.float64) ** 2) / nn
        keep = (af > MAF) & (af < 1 - MAF) & (ss > 0)
        if not keep.any():
            error("No markers remain after streaming genotype quality control.")
        packed = np.ascontiguousarray(packed[keep]); means = means[keep]; af = af[keep]
        mk = [m for m, k in zip(mk, keep) if k]
        selected = selected[keep]
        if ann_matrix is not False:
            ann_matrix = ann_matrix[keep]   # annotations follow the markers that survive QC (readgenotypes.jl:256)
    p = packed.shape[0]
    g = Genotypes(name=name, obsID=obs, markerID=mk, nObs=n, nMarkers=p, alleleFreq=af,
                  sum2pq=float((2.0 * af.astype(np.float64) * (1 - af.astype(np.float64))).sum()),
                  centered=centered, packed=np.ascontiguousarray(packed), marker_means=means, method=method,
                  estimatePi=bool(estimatePi), multi_trait_sampler=multi_trait_sampler,
                  starting_value=starting_value, nMarkersAll=int(n_all),
                  selected_marker_indices=np.asarray(selected, dtype=np.int32))
    g.π = Pi if not isinstance(Pi, (list, tuple)) else np.array(Pi, dtype=np.float64)
    g.G = Variance(val=G if G_is_marker_variance else False, df=df, estimate_variance=estimate_variance,
                   estimate_scale=estimate_scale, constraint=constraint)
    g.genetic_variance = Variance(val=False if G_is_marker_variance else G, df=df)
    if method == "BayesA":                 # input_data_validation.jl:33-36
        g.method = "BayesB"; g.π = 0.0; g.estimatePi = False
    if method in ("RR-BLUP", "BayesL"):    # input_data_validation.jl:24-31: "runs with π = false / estimatePi = false"
        g.π = 0.0; g.estimatePi = False
    if ann_matrix is not False:            # readgenotypes.jl:127-150, :111-125
        try:
            g.annotation_start_pi = annot.annotation_starting_pi(method, g.π, p)
            g.annotations = annot.build_marker_annotations(ann_matrix, method, g.π)
        except annot.AnnotationError as e:
            error(str(e))
        if method == "BayesC":
            g.π = g.annotation_start_pi
    return g


# ------------------------------------------------------------------------------------ .jgb2 backend files
def prepare_streaming_genotypes(file, *, output_prefix=None, separator=",", header=True, quality_control=True,
                                MAF=0.01, missing_value=9.0, center=True):
    """Writes the reference's packed backend files (streaming_genotypes.jl:819-877, dense path
    :520-660): <prefix>.jgb2 + .meta and the Float32/Int32/text side-cars, readable by JWAS.jl's
    storage=:stream and by get_genotypes(prefix) here."""
    g = get_genotypes(file, 1.0, separator=separator, header=header, quality_control=quality_control, MAF=MAF,
                      missing_value=missing_value, center=True)
    g.centered = bool(center)        # recorded in the manifest; the image itself holds codes either way
    prefix = os.path.abspath(output_prefix or (os.path.splitext(file)[0] + "_stream"))
    paths = {k: prefix + ext for k, ext in (("data_path", ".jgb2"), ("obs_path", ".obsid.txt"),
                                            ("marker_path", ".markerid.txt"), ("selected_path", ".selected.i32"),
                                            ("mean_path", ".mean.f32"), ("xp_path", ".xpRinvx.f32"),
                                            ("afreq_path", ".afreq.f32"))}
    g.packed.tofile(paths["data_path"])
    open(paths["obs_path"], "w").write("".join(x + "\n" for x in g.obsID))
    open(paths["marker_path"], "w").write("".join(x + "\n" for x in g.markerID))
    g.selected_marker_indices.astype(np.int32).tofile(paths["selected_path"])        # raw (1-based) indices kept by QC
    g.marker_means.astype(np.float32).tofile(paths["mean_path"])
    # xpRinvx, missing calls at the mean (streaming_genotypes.jl:283-285, 560-585), in closed form from the call counts:
    # centred   sum (x - m)^2 = (n1 + 4 n2) - 2 m (n1 + 2 n2) + m^2 * (observed calls)
    # uncentred sum x^2       = (n1 + 4 n2) + m^2 * (missing calls)
    cnt = _call_counts(g.packed, g.nObs)
    m64 = g.marker_means.astype(np.float64)
    s64 = (cnt[:, 0] + 2 * cnt[:, 1]).astype(np.float64)
    if g.centered:
        xp = (cnt[:, 0] + 4 * cnt[:, 1]) - 2 * m64 * s64 + m64 * m64 * (g.nObs - cnt[:, 2])
    else:
        xp = (cnt[:, 0] + 4 * cnt[:, 1]) + m64 * m64 * cnt[:, 2]
    xp.astype(np.float32).tofile(paths["xp_path"])
    g.alleleFreq.astype(np.float32).tofile(paths["afreq_path"])
    with open(prefix + ".meta", "w") as io:
        for k, v in [("version", "1")] + list(paths.items()) + [("nObs", g.nObs), ("nMarkers", g.nMarkers),
                                                               ("nMarkersAll", g.nMarkersAll),
                                                               ("stride_bytes", (g.nObs + 3) // 4),
                                                               ("centered", int(g.centered)), ("sum2pq", repr(g.sum2pq))]:
            io.write(f"{k}\t{v}\n")
    return prefix


def load_streaming_backend(path):
    """streaming_genotypes.jl:884-971 (manifest: tab-separated key/value lines, :77-95)."""
    prefix = os.path.abspath(path)
    for ext in (".meta", ".jgb2"):
        if prefix.endswith(ext):
            prefix = prefix[:-len(ext)]
    meta_path = prefix + ".meta"
    if not os.path.isfile(meta_path):
        error(f"Streaming manifest is not found: {meta_path}")
    meta = {}
    for line in open(meta_path):
        parts = line.rstrip("\n").split("\t", 1)
        if len(parts) == 2:
            meta[parts[0]] = parts[1]
    n, p, stride = int(meta["nObs"]), int(meta["nMarkers"]), int(meta["stride_bytes"])
    if os.path.getsize(meta["data_path"]) != p * stride:
        error(f"Packed genotype file size does not match metadata for {meta['data_path']}")
    packed = np.fromfile(meta["data_path"], dtype=np.uint8).reshape(p, stride)
    obs = [l.rstrip("\n") for l in open(meta["obs_path"])]
    mk = [l.rstrip("\n") for l in open(meta["marker_path"])]
    if len(obs) != n:
        error(f"Number of IDs in {meta['obs_path']} does not match nObs in manifest.")
    if len(mk) != p:
        error(f"Number of markers in {meta['marker_path']} does not match nMarkers in manifest.")
    # raw-marker mapping (streaming_genotypes.jl:913-944): older backends have neither entry
    if ("nMarkersAll" in meta) != bool(meta.get("selected_path")):
        error("Streaming backend metadata is inconsistent. Rebuild the backend with prepare_streaming_genotypes.")
    has_map = "nMarkersAll" in meta and bool(meta.get("selected_path"))
    n_all = int(meta["nMarkersAll"]) if "nMarkersAll" in meta else p
    if has_map:
        if not os.path.isfile(meta["selected_path"]):
            error("Streaming backend selected-marker metadata is missing. Rebuild the backend with prepare_streaming_genotypes.")
        sel = np.fromfile(meta["selected_path"], dtype=np.int32)
        if len(sel) != p:
            error("Number of selected raw-marker indices does not match nMarkers in manifest.")
        if len(sel) and (sel.min() < 1 or sel.max() > n_all):
            error("Selected raw-marker indices are out of bounds for the recorded raw marker count.")
    else:
        sel = np.arange(1, p + 1, dtype=np.int32)
    return {"packed": packed, "nObs": n, "nMarkers": p, "obsID": obs, "markerID": mk,
            "centered": int(meta["centered"]) == 1, "nMarkersAll": n_all, "selected_marker_indices": sel,
            "has_raw_marker_mapping": has_map}


# ------------------------------------------------------------------------------------ model
def build_model(model_equations, R=False, *, df=4.0, genotypes=None, estimate_variance=True, constraint=False):
    """build_MME.jl:42-156 for `trait = intercept + <genotype term>` equations (one per line or ';')."""
    if not isinstance(model_equations, str) or not model_equations.strip():
        error("Model equations are wrong.")
    eqs = [e.strip() for e in model_equations.replace(";", "\n").split("\n") if e.strip()]
    lhs, geno_names = [], None
    for e in eqs:
        if "=" not in e:
            error("Model equations are wrong.")
        l, r = [x.strip() for x in e.split("=", 1)]
        terms = [t.strip() for t in r.split("+")]
        if "intercept" not in terms:
            error("storage=:gpu models must contain an intercept.")
        others = [t for t in terms if t != "intercept"]
        if geno_names is None:
            geno_names = others
        elif others != geno_names:
            error("every equation must contain the same genotype term with storage=:gpu.")
        lhs.append(l)
    if genotypes is None:
        import inspect
        fr = inspect.currentframe().f_back
        genotypes = {k: v for k, v in {**fr.f_globals, **fr.f_locals}.items() if isinstance(v, Genotypes)}
    M = []
    for nm in geno_names or []:
        if nm not in genotypes:
            error(f"{nm} is not a genotype term known to this backend: covariates, factors, pedigree and "
                  "random terms are outside the GPU marker-sweep path.")
        gi = genotypes[nm]; gi.name = nm; gi.ntraits = len(lhs)
        try:                                                 # build_MME.jl -> annotation_setup.jl:141-153
            annot.finalize_marker_annotation_setup(gi)
        except annot.AnnotationError as e:
            error(str(e))
        if len(lhs) != 1 and not getattr(gi, "_df_bumped", False):
            gi.G.df = gi.G.df + len(lhs)                     # build_MME.jl:108-110
            gi._df_bumped = True
        M.append(gi)
    if len(M) != 1:
        error("exactly one genotype term is supported with storage=:gpu.")
    df_R = df if len(lhs) == 1 else df + len(lhs)            # build_MME.jl:128-134
    return MME(model_equations=model_equations, lhsVec=lhs, nModels=len(lhs), M=M,
               R=Variance(val=R, df=df_R, estimate_variance=estimate_variance, constraint=constraint))


def outputEBV(model, IDs):
    """output.jl:60-69: estimated breeding values and prediction error variances for these IDs."""
    model.output_ID = [str(x) for x in np.asarray(IDs).reshape(-1)]


def set_covariate(*a, **k):
    error("set_covariate: covariates are outside the GPU marker-sweep path.")


def set_random(*a, **k):
    error("set_random: random / pedigree terms are outside the GPU marker-sweep path.")


def validate_fast_block_starts(block_starts, nmarkers):     # JWAS.jl:73-79 (1-based starts)
    bs = list(block_starts)
    if len(bs) == 0:
        error("fast_blocks block start vector cannot be empty.")
    if bs[0] != 1:
        error("fast_blocks block starts must begin with 1.")
    if not all(1 <= b <= nmarkers for b in bs):
        error("fast_blocks block starts must be within 1:nMarkers.")
    if not all(b2 > b1 for b1, b2 in zip(bs, bs[1:])):
        error("fast_blocks block starts must be sorted and unique.")


def resolve_fast_blocks(fast_blocks, chain_length, n_obs, n_markers):
    """JWAS.jl:293-316.  Returns (0-based boundaries or None, chain_length)."""
    if fast_blocks is False or fast_blocks is None:
        return None, chain_length
    if fast_blocks is True:
        bsize = int(math.floor(math.sqrt(n_obs)))
    elif isinstance(fast_blocks, (int, float, np.integer, np.floating)):
        bsize = int(math.floor(fast_blocks))
    elif isinstance(fast_blocks, (list, tuple, np.ndarray)) and all(float(b).is_integer() for b in fast_blocks):
        validate_fast_block_starts([int(b) for b in fast_blocks], n_markers)
        st = [int(b) - 1 for b in fast_blocks] + [n_markers]
        return np.array(st, dtype=np.int64), chain_length          # explicit starts keep chain_length
    else:
        error("fast_blocks must be false, true, a positive number, or a vector of block start positions.")
    if bsize < 1:
        error("fast_blocks block size must be at least 1.")
    starts1 = list(range(1, n_markers + 1, bsize))
    if len(starts1) <= 1:
        error("fast_blocks block size must create at least two block starts.")
    chain_length = int(math.floor(chain_length / (starts1[1] - starts1[0])))   # number of outer loops
    return np.array([s - 1 for s in starts1] + [n_markers], dtype=np.int64), chain_length


def _frame(rows, cols):
    return pd.DataFrame(rows, columns=cols) if pd is not None else {"columns": cols, "rows": rows}


def runMCMC(model, df, *, chain_length=100, burnin=0, output_samples_frequency=None, seed=False,
            fast_blocks=False, independent_blocks=False, outputEBV=True, output_heritability=True,
            double_precision=False, heterogeneous_residuals=False, output_folder=None, device=0,
            panel=DEFAULT_PANEL, engine=1, lag=DEFAULT_LAG, chain_ctas=DEFAULT_CHAIN_CTAS, output_marker_effect_samples=False,
            memory_guard="error", memory_guard_ratio=0.80,
            _backend_factory=None, **other):
    """JWAS.jl:161-511 -> MCMC_BayesianAlphabet (MCMC/MCMC_BayesianAlphabet.jl:4) for the GPU backend.
    Returns the reference's output dictionary keys for this path: "location parameters",
    "residual variance", "marker effects <name>", "pi_<name>", "annotation coefficients <name>", "EBV_<trait>",
    "genetic_variance" and "heritability" (output_heritability=true, output.jl:196-209, 498-511).

    output_folder: the reference always creates a folder ("results", "results1", ... JWAS.jl:255-262) and writes every
    output table there as <key with _ for spaces>.txt (JWAS.jl:479-482).  Here that happens when output_folder is
    given -- together with the MCMC sample files of the hyper-parameters, EBVs, genetic variance and heritability
    (output.jl:318-515); the default (None) writes nothing but requested marker-effect sample files (into "results")."""
    write_results = output_folder is not None
    if output_folder is None:
        output_folder = "results"
    elif os.path.exists(output_folder):          # JWAS.jl:255-262: never write into an existing folder
        base, k = output_folder, 1
        while os.path.exists(output_folder):
            output_folder = base + str(k); k += 1
    if write_results:
        os.makedirs(output_folder)
    # the reference's remaining keyword arguments (JWAS.jl:161-200): the ones that only steer printing or are deprecated
    # are accepted, the ones that select a model this backend does not run are refused instead of being ignored
    passive = {"printout_model_info", "printout_frequency", "big_memory", "fitting_J_vector", "methods",
               "output_samples_for_all_parameters", "Pi", "estimatePi", "estimate_scale", "estimate_variance"}
    refused = {"starting_value": False, "update_priors_frequency": 0, "single_step_analysis": False, "pedigree": False,
               "causal_structure": False, "RRM": False, "prediction_equation": False, "missing_phenotypes": True,
               "categorical_trait": False, "censored_trait": False}
    for key, value in other.items():
        if key in passive:
            continue
        if key in refused:
            if not (value is refused[key] or value == refused[key]):
                error(f"{key}={value!r} is outside the GPU marker-sweep path (storage=:gpu).")
            continue
        error(f"runMCMC got an unknown keyword argument: {key}")
    # output_heritability (default true as in the reference) needs the EBVs; it makes every genotyped individual an
    # output ID (check_outputID, input_data_validation.jl:167-174) and is skipped when outputEBV is off (output.jl:498)
    output_heritability = bool(output_heritability) and bool(outputEBV)
    if output_heritability:
        model.output_ID = False
    if heterogeneous_residuals:
        error("heterogeneous_residuals=true is not supported with storage=:gpu (unit weights only).")
    if double_precision:
        error("double_precision=true is not supported with storage=:gpu.")
    if independent_blocks and fast_blocks is False:
        error("independent_blocks=true requires fast_blocks != false.")
    if burnin >= chain_length:
        error("burnin must be smaller than chain_length.")
    Mi = model.M[0]
    t = model.nModels
    annotated = Mi.annotations is not False and Mi.annotations is not None
    if annotated and t > 1 and not (Mi.method == "BayesC" and t == 2 and not Mi.G.constraint):
        # MCMC_BayesianAlphabet.jl:19-25
        error("Annotated multi-trait BayesC currently supports exactly 2 traits with storage=:dense and "
              "constraint=false.")
    if annotated and t > 1 and Mi.multi_trait_sampler == "II":
        error("annotated 2-trait BayesC runs sampler I with storage=:gpu (jwas_sweep_mt2 takes no per-marker prior).")
    if t > 1 and Mi.method not in ("BayesC", "RR-BLUP"):
        error("multi-trait analysis with storage=:gpu supports BayesC (samplers I / II) and RR-BLUP only.")
    mt_sampler = "I"
    if Mi.multi_trait_sampler == "II" and Mi.method != "BayesC":       # build_MME.jl:104-107
        error("multi_trait_sampler overrides are supported for BayesC only.")
    if t > 1 and Mi.multi_trait_sampler == "II":
        if t != 2:
            error("multi_trait_sampler=:II is supported for exactly 2 traits with storage=:gpu.")
        mt_sampler = "II"
    if t > 4:
        error("at most 4 traits are supported with storage=:gpu.")
    # ---- phenotypes aligned to the genotype IDs (JWAS.jl:381-402)
    ids = [str(x) for x in (df.iloc[:, 0] if pd is not None and isinstance(df, pd.DataFrame) else df["ID"])]
    pos = {k: i for i, k in enumerate(Mi.obsID)}
    missing_ids = [i for i in ids if i not in pos]
    if missing_ids:
        error("phenotyped individuals without genotypes are not supported with storage=:gpu.")
    rows = np.array([pos[i] for i in ids])
    Y = np.array([np.asarray(df[tr], dtype=np.float64) for tr in model.lhsVec])
    if not np.all(np.isfinite(Y)):
        error("missing phenotypes are not supported with storage=:gpu.")
    n = len(ids); p = Mi.nMarkers
    packed = Mi.packed
    subset_means = None
    if n != Mi.nObs or not np.array_equal(rows, np.arange(n)):
        # the reference centres on ALL genotyped individuals in get_genotypes (readgenotypes.jl:372-385) and only then
        # aligns rows to the phenotyped ones (JWAS.jl:381-402): keep the full-sample means, recompute xpx for them
        packed = _select_rows(Mi.packed, Mi.nObs, rows)
        if np.any(_call_counts(packed, n)[:, 2] == n):
            error("a marker has no observed genotype among the phenotyped individuals.")
        subset_means = np.asarray(Mi.marker_means, dtype=np.float32)
    if not Mi.centered:
        # center=false (no missing calls, checked in get_genotypes): the device centres on the means it is given, so
        # means of zero make x_ij the code itself -- xpRinvx = sum of squared codes, EBV = M * alpha uncentred
        subset_means = np.zeros(p, dtype=np.float32)

    if output_samples_frequency is None:                 # evaluated on the user's chain_length (JWAS.jl:168), before :312
        output_samples_frequency = chain_length // 1000 if chain_length > 1000 else 1
    starts, chain_length = resolve_fast_blocks(fast_blocks, chain_length, n, p)
    if burnin >= chain_length:
        # fast_blocks divides chain_length by the block size (JWAS.jl:312) but leaves burnin alone: nothing would be saved
        error(f"burnin ({burnin}) must be smaller than the number of outer iterations ({chain_length}) that fast_blocks "
              "leaves of chain_length.")
    # seed=false: unseeded run (JWAS.jl:239-241 only seeds when a number is given)
    seed_v = int.from_bytes(os.urandom(4), "little") if seed is False else int(seed)
    model.MCMCinfo = MCMCinfo(chain_length=chain_length, burnin=burnin, output_samples_frequency=output_samples_frequency,
                              seed=seed, fast_blocks=(False if starts is None else [int(s) + 1 for s in starts[:-1]]),
                              independent_blocks=independent_blocks, outputEBV=outputEBV, output_folder=output_folder)
    schedule = SCHED_EXACT if starts is None else (SCHED_INDEPENDENT if independent_blocks else SCHED_BLOCK)
    if starts is None:
        has_missing = bool(np.any((packed & (packed >> 1) & 0x55) != 0))
        pipelined = bool(lag) and engine == 1 and chain_ctas > 0
        pmax = 4096 if pipelined else (2048 if (t == 1 and not has_missing) else 1024)
        starts = np.array(list(range(0, p, max(1, min(panel, pmax)))) + [p], dtype=np.int64)

    # ---- default priors (input_data_validation.jl:296-350) and marker hyper-parameters
    #      (tools4genotypes.jl:353-424)
    vary = Y.var(axis=1, ddof=1)
    if t == 1:
        varg = vary[0] * 0.5
        if model.R.val is False:
            model.R.val = float(np.float32(vary[0] * 0.5))
        model.R.scale = model.R.val * (model.R.df - 2) / model.R.df
        if Mi.method == "BayesR":
            pi = np.array([0.95, 0.03, 0.015, 0.005]) if np.isscalar(Mi.π) and Mi.π == 0.0 else np.array(Mi.π, float)
            if len(pi) != 4:
                error("BayesR Pi must have length 4.")
            denom = Mi.sum2pq * float(mcmc.BAYESR_GAMMA @ pi)
        elif Mi.annotations is not False:
            # marker-level starting pi (genetic2marker, tools4genotypes.jl:468-475)
            pi = np.array(Mi.π, dtype=np.float64)
            if len(pi) != p:
                error(f"BayesC marker-level Pi must have length {p}.")
            af64 = Mi.alleleFreq.astype(np.float64)
            denom = float(np.sum(2.0 * af64 * (1.0 - af64) * (1.0 - np.clip(pi, 0.0, 1.0))))
            if not denom > 0:
                error("BayesC implied variance denominator must be positive.")
        else:
            pi = float(Mi.π)
            denom = (1 - pi) * Mi.sum2pq
        if Mi.G.val is False:
            gv = Mi.genetic_variance.val if Mi.genetic_variance.val is not False else varg
            Mi.G.val = float(np.float32(gv / denom))
        if not Mi.G.val > 0:
            error("Marker effects variance is negative!")
        Mi.G.scale = Mi.G.val * (Mi.G.df - 2) / Mi.G.df
        if Mi.method == "BayesL":          # MCMC_BayesianAlphabet.jl:70-74: G.val is the scale "Sigma" of the lasso prior
            Mi.G.val = Mi.G.val / 8; Mi.G.scale = Mi.G.scale / 8
        Mi.π = pi
    else:
        if model.R.val is False:
            model.R.val = np.diag(vary * 0.5)
        model.R.val = np.array(model.R.val, dtype=np.float64)
        # df already carries the + nModels of build_model (build_MME.jl:128-134): scale = R * (df_user - 1);
        # R_constraint! (input_data_validation.jl:530-540) then takes the traits back out and makes it diagonal
        df_R = model.R.df
        model.R.scale = model.R.val * (df_R - t - 1)         # input_data_validation.jl:345
        if model.R.constraint:
            df_R = df_R - t
            model.R.scale = np.diag(np.diag(model.R.scale) / (df_R - 1)) * (df_R - 2) / df_R
        if Mi.G.constraint:
            # megaBayesABC!: one pi per trait (MCMC_BayesianAlphabet.jl:96-99 starts them at zero)
            big = np.zeros(t) if (np.isscalar(Mi.π) or isinstance(Mi.π, dict)) else np.array(Mi.π, dtype=np.float64)
            if big.shape != (t,):
                error("constraint=true needs one Pi per trait.")
        elif isinstance(Mi.π, dict):
            big = np.zeros(1 << t)
            for key, v in Mi.π.items():
                big[sum(int(round(k)) << i for i, k in enumerate(key))] = v
        elif np.isscalar(Mi.π) and Mi.π == 0.0:
            big = np.zeros(1 << t); big[-1] = 1.0      # "all markers have effects on all traits"
        else:
            big = np.array(Mi.π, dtype=np.float64)
        if not Mi.G.constraint and abs(big.sum() - 1.0) > 1e-8:
            error("Summation of probabilities of Pi is not equal to one.")
        if Mi.G.val is False and annotated:
            # genetic2marker with the marker-level joint priors (tools4genotypes.jl:440-455)
            gv = np.array(Mi.genetic_variance.val, float) if Mi.genetic_variance.val is not False else np.diag(vary * 0.5)
            af64 = Mi.alleleFreq.astype(np.float64)
            twopq = 2.0 * af64 * (1.0 - af64)
            sp = Mi.annotations.snp_pi
            d12 = float(np.sum(twopq * sp[:, 3]))
            denom = np.array([[float(np.sum(twopq * (sp[:, 1] + sp[:, 3]))), d12],
                              [d12, float(np.sum(twopq * (sp[:, 2] + sp[:, 3])))]])
            if np.any(denom <= 0):
                error("Annotated multi-trait BayesC implied covariance denominator must be positive.")
            Mi.G.val = gv / denom
        if Mi.G.val is False:
            gv = np.array(Mi.genetic_variance.val, float) if Mi.genetic_variance.val is not False else np.diag(vary * 0.5)
            denom = np.zeros((t, t))
            for i in range(t):
                for j in range(t):
                    if Mi.G.constraint:
                        denom[i, j] = Mi.sum2pq * (1 - big[i]) * (1 - big[j]) if i != j else Mi.sum2pq * (1 - big[i])
                    else:
                        denom[i, j] = Mi.sum2pq * sum(big[s] for s in range(1 << t) if (s >> i) & 1 and (s >> j) & 1)
            if np.any(denom <= 0):
                error("Marker effects covariance matrix is not postive definite! Please modify the argument: Pi.")
            Mi.G.val = gv / denom
        Mi.G.val = np.array(Mi.G.val, dtype=np.float64)
        if np.any(np.linalg.eigvalsh(Mi.G.val) <= 0):
            error("Marker effects covariance matrix is not postive definite! Please modify the argument: Pi.")
        df_G = Mi.G.df                                       # user df + nModels (build_MME.jl:108-110)
        Mi.G.scale = Mi.G.val * (df_G - t - 1)               # tools4genotypes.jl:417
        if Mi.G.constraint:                                  # G_constraint! (input_data_validation.jl:543-558)
            df_G = df_G - t
            Mi.G.scale = np.diag(np.diag(Mi.G.scale) / (df_G - 1)) * (df_G - 2) / df_G
        Mi.π = big

    # ---- device-resident backend: GibbsMats (MCMC_BayesianAlphabet.jl:58) + ycorr (:131-147)
    use_lag = int(lag) if (schedule == SCHED_EXACT and engine == 1) else 0
    if use_lag >= 2 and not chain_ctas:
        use_lag = 1                                     # lag 2 needs the pipelined chain
    # marker-memory precheck (JWAS.jl:415-458) against the HBM of the device instead of the host's RAM
    est = estimate_marker_memory(n, p, element_bytes=4, block_starts=[int(s_) + 1 for s_ in starts[:-1]], storage_mode="gpu",
                                 n_traits=t, lag=use_lag)
    check_marker_memory_guard(mode=memory_guard, ratio=memory_guard_ratio, estimated_bytes=est["bytes_total"],
                              total_memory_bytes=device_memory_bytes(),
                              context_string=f"geno={Mi.name}, storage=:gpu, nObs={n}, nMarkers={p}, traits={t}, "
                                             f"packed+tiled={format_bytes_human(est['bytes_packed'] + est['bytes_tiled'])}, "
                                             f"Gram={format_bytes_human(est['bytes_XpRinvX'] + est['bytes_cross_gram'])}")
    if _backend_factory is not None:
        backend = _backend_factory(packed, n, t, starts) if subset_means is None else \
            _backend_factory(packed, n, t, starts, means=subset_means)
        backend.lag = use_lag
    else:
        sw = GpuSweeper(packed, n, t, device=device)
        if subset_means is not None:
            sw.set_marker_means(subset_means)
        sw.set_option("engine", engine)
        sw.set_option("lag", use_lag)
        sw.set_option("chain_ctas", int(chain_ctas) if use_lag else 0)
        sw.set_blocks(starts)
        backend = mcmc.GpuBackend(sw)
        Mi.stream_backend = sw
    # ---- EBV output IDs (check_outputID, input_data_validation.jl:143-196): every genotyped individual unless
    #      outputEBV(model, IDs) named others; IDs without genotypes are dropped.  When they are not exactly the training
    #      rows, a second handle holds their rows (centred on the same full-sample means, align_genotypes,
    #      tools4genotypes.jl:288-296) and serves the getEBV product M_out * alpha (output.jl:300-304)
    ebv_backend = None
    ebv_ids = ids
    if outputEBV:
        want_ids = list(Mi.obsID) if model.output_ID is False else list(model.output_ID)
        if any(i not in pos for i in want_ids):
            import warnings
            warnings.warn("Testing individuals are not a subset of genotyped individuals (complete genomic data,"
                          "non-single-step). Only output EBV for tesing individuals with genotypes.")
            want_ids = [i for i in want_ids if i in pos]
        if want_ids != ids:
            ebv_ids = want_ids
            out_rows = np.array([pos[i] for i in want_ids], dtype=np.int64)
            out_packed = _select_rows(Mi.packed, Mi.nObs, out_rows)
            out_means = np.asarray(Mi.marker_means, dtype=np.float32) if Mi.centered else np.zeros(p, dtype=np.float32)
            if _backend_factory is not None:
                ebv_backend = _backend_factory(out_packed, len(want_ids), t, np.array([0, p], dtype=np.int64), means=out_means)
            else:
                sw_out = GpuSweeper(out_packed, len(want_ids), t, device=device)
                sw_out.set_marker_means(out_means)
                ebv_backend = mcmc.GpuBackend(sw_out)
    mu0 = Y.mean(axis=1)
    alpha0 = np.zeros(t * p, np.float32)
    if Mi.starting_value is not False:
        alpha0 = np.asarray(Mi.starting_value, dtype=np.float32).reshape(-1).copy()
    delta0 = np.ones(t * p, np.int32)
    backend.put_state(alpha0, alpha0.copy(), delta0)       # beta = copy(alpha), delta = ones (:89-90,119)
    backend.put_ycorr((Y - mu0[:, None]).astype(np.float32).reshape(-1))
    if np.any(alpha0 != 0):
        backend.sub_malpha()

    # marker-effect sample files (output.jl:411, 467): <output_folder>/MCMC_samples_marker_effects_<geno>_<trait>.txt,
    # header of marker IDs + one row per saved iteration -- what GWAS() (and the reference's GWAS) reads.  Off by
    # default here: a row is p numbers per saved iteration and trait.
    sample_files = None
    if output_marker_effect_samples:              # rows are streamed to the files as they are produced (output.jl:467)
        os.makedirs(output_folder, exist_ok=True)
        model.sample_files = {}
        sample_files = []
        for tr in model.lhsVec:
            path = os.path.join(output_folder, f"MCMC_samples_marker_effects_{Mi.name}_{tr}.txt")
            f = open(path, "w")
            f.write(",".join(str(m) for m in Mi.markerID) + "\n")
            sample_files.append(f); model.sample_files[tr] = path

    # hyper-parameter and EBV sample files (output_MCMC_samples_setup / output_MCMC_samples, output.jl:318-515), written
    # when an output folder was asked for: residual variance, marker effect variances, pi, EBVs, genetic variance and
    # heritability.  (BayesB's per-marker variances stay on the device and are not written.)
    hyper_files = {}
    if write_results:
        def _open(name, header=None):
            f = open(os.path.join(output_folder, f"MCMC_samples_{name}.txt"), "w")
            if header is not None:
                f.write(",".join(header) + "\n")
            hyper_files[name] = f
        pairs = [f"{a}_{b}" for a in model.lhsVec for b in model.lhsVec]
        _open("residual_variance", list(model.lhsVec) if (t == 1 or model.R.constraint) else pairs)
        if Mi.method not in ("BayesB", "BayesA"):
            _open("marker_effects_variances_" + Mi.name)
        if Mi.estimatePi:
            _open("pi_" + Mi.name)
        if outputEBV:
            for tr in model.lhsVec:
                _open("EBV_" + tr, list(ebv_ids))
            if output_heritability:
                _open("genetic_variance", pairs if t > 1 else list(model.lhsVec))
                _open("heritability", list(model.lhsVec))

    def _row(x):
        return ",".join(repr(float(v)) for v in np.atleast_1d(np.asarray(x, dtype=np.float64)).reshape(-1)) + "\n"

    def hyper_sink(smp):
        ve = np.atleast_2d(smp["vare"])
        hyper_files["residual_variance"].write(_row(np.diag(ve) if (t > 1 and model.R.constraint) else ve))
        f = hyper_files.get("marker_effects_variances_" + Mi.name)
        if f is not None and smp["vara"] is not None:
            for line in np.atleast_2d(smp["vara"]):          # a scalar, or the t x t matrix row by row (writedlm)
                f.write(_row(line))
        f = hyper_files.get("pi_" + Mi.name)
        if f is not None and smp["pi"] is not None:
            pv = np.atleast_1d(smp["pi"])
            for v in pv:                                      # one value per line, a blank line after a vector
                f.write(repr(float(v)) + "\n")
            if len(pv) > 1:
                f.write("\n")
        if smp["ebv"] is not None:
            for k, tr in enumerate(model.lhsVec):
                hyper_files["EBV_" + tr].write(_row(smp["ebv"][k]))
        if smp["gvar"] is not None:
            hyper_files["genetic_variance"].write(_row(smp["gvar"]))
            hyper_files["heritability"].write(_row(smp["h2"]))

    def sink(alpha):
        for k in range(t):
            sample_files[k].write(",".join(repr(float(x)) for x in np.asarray(alpha[k * p:(k + 1) * p], dtype=np.float32)) + "\n")

    out = mcmc.run_chain(backend, n=n, p=p, ntraits=t, method=Mi.method, schedule=schedule,
                         sample_sink=(sink if output_marker_effect_samples else None),
                         chain_length=chain_length, burnin=burnin, output_samples_frequency=output_samples_frequency,
                         seed=seed_v, vare=model.R.val if t == 1 else None, var_effect=Mi.G.val if t == 1 else None,
                         pi=Mi.π if t == 1 else None, df_effect=(Mi.G.df if t == 1 else df_G),
                         scale_effect=Mi.G.scale if t == 1 else None,
                         df_res=(model.R.df if t == 1 else df_R), scale_res=model.R.scale if t == 1 else None,
                         estimate_pi=Mi.estimatePi, estimate_variance=Mi.G.estimate_variance,
                         estimate_vare=model.R.estimate_variance,
                         R=model.R.val if t > 1 else None, G=Mi.G.val if t > 1 else None,
                         big_pi=Mi.π if t > 1 else None, scale_G=Mi.G.scale if t > 1 else None,
                         scale_R=model.R.scale if t > 1 else None, mu0=mu0, want_ebv=outputEBV, mt_sampler=mt_sampler,
                         constraint_G=bool(t > 1 and Mi.G.constraint), constraint_R=bool(t > 1 and model.R.constraint),
                         annotations=(Mi.annotations if annotated else None), ebv_backend=ebv_backend,
                         want_heritability=bool(output_heritability), hyper_sink=(hyper_sink if write_results else None))
    for f in hyper_files.values():
        f.close()

    # ---- output dictionary (output.jl:108-212)
    ma, ma2, md = backend.get_means()
    Mi.meanAlpha = [ma[k * p:(k + 1) * p] for k in range(t)]
    Mi.meanAlpha2 = [ma2[k * p:(k + 1) * p] for k in range(t)]
    Mi.meanDelta = [md[k * p:(k + 1) * p] for k in range(t)]
    al, be, de = backend.get_state()
    Mi.α = [al[k * p:(k + 1) * p] for k in range(t)]
    Mi.β = [be[k * p:(k + 1) * p] for k in range(t)]
    Mi.δ = [de[k * p:(k + 1) * p] for k in range(t)]
    output = {}
    output["location parameters"] = _frame(
        [[tr, "intercept", "intercept", out["mu_mean"][k], math.sqrt(abs(out["mu_mean2"][k] - out["mu_mean"][k] ** 2))]
         for k, tr in enumerate(model.lhsVec)], ["Trait", "Effect", "Level", "Estimate", "SD"])
    vm = np.atleast_2d(out["vare_mean"]); vm2 = np.atleast_2d(out["vare_mean2"])
    output["residual variance"] = _frame(
        [[f"{model.lhsVec[i]}_{model.lhsVec[j]}", vm[i, j], math.sqrt(abs(vm2[i, j] - vm[i, j] ** 2))]
         for i in range(t) for j in range(t)], ["Covariance", "Estimate", "SD"])
    rows_me = []
    for k, tr in enumerate(model.lhsVec):
        sd = np.sqrt(np.abs(Mi.meanAlpha2[k].astype(np.float64) - Mi.meanAlpha[k].astype(np.float64) ** 2))
        rows_me += [[tr, mid, float(e), float(s_), float(d)] for mid, e, s_, d in
                    zip(Mi.markerID, Mi.meanAlpha[k], sd, Mi.meanDelta[k])]
    output["marker effects " + Mi.name] = _frame(rows_me, ["Trait", "Marker_ID", "Estimate", "SD", "Model_Frequency"])
    if Mi.estimatePi:
        pm = np.atleast_1d(out["pi_mean"]); pm2 = np.atleast_1d(out["pi_mean2"])
        if t == 1 and Mi.method != "BayesR" and annotated:
            labels = [str(j + 1) for j in range(p)]          # marker-level pi_j (output.jl:256-266)
        elif t == 1 and Mi.method != "BayesR":
            labels = ["π"]
        elif t == 1:
            labels = [f"class{c + 1}" for c in range(len(pm))]
        elif Mi.G.constraint:
            labels = list(model.lhsVec)
        else:
            labels = [str([float((s >> i) & 1) for i in range(t)]) for s in range(1 << t)]
        output["pi_" + Mi.name] = _frame([[l, pm[i], math.sqrt(abs(pm2[i] - pm[i] ** 2))] for i, l in enumerate(labels)],
                                         ["π", "Estimate", "SD"])
        Mi.mean_pi = out["pi_mean"]
    if annotated:                                            # output.jl:151-177
        ann = Mi.annotations
        asd = np.sqrt(np.abs(ann.mean_coefficients2 - ann.mean_coefficients ** 2))
        if ann.nsteps == 1:
            output["annotation coefficients " + Mi.name] = _frame(
                [[nm, float(e), float(s_)] for nm, e, s_ in zip(ann.names(), ann.mean_coefficients, asd)],
                ["Annotation", "Estimate", "SD"])
        else:
            steps = annot.STEP_LABELS["BayesR" if Mi.method == "BayesR" else "BayesC2"]
            output["annotation coefficients " + Mi.name] = _frame(
                [[nm, steps[c], float(ann.mean_coefficients[k, c]), float(asd[k, c])]
                 for k, nm in enumerate(ann.names()) for c in range(ann.nsteps)],
                ["Annotation", "Step", "Estimate", "SD"])
    if outputEBV and out.get("ebv_mean") is not None:
        for k, tr in enumerate(model.lhsVec):
            em, ev = out["ebv_mean"][k], out["ebv_var"][k]
            output["EBV_" + tr] = _frame([[i, float(a), float(v)] for i, a, v in zip(ebv_ids, em, ev)], ["ID", "EBV", "PEV"])
    if output_heritability and out.get("gvar_mean") is not None:      # output.jl:196-209
        names = [f"{a}_{b}" for a in model.lhsVec for b in model.lhsVec] if t > 1 else list(model.lhsVec)
        gm, gs = np.atleast_1d(out["gvar_mean"]).reshape(-1), np.atleast_1d(out["gvar_sd"]).reshape(-1)
        output["genetic_variance"] = _frame([[nm, float(a), float(b)] for nm, a, b in zip(names, gm, gs)],
                                            ["Covariance", "Estimate", "SD"])
        output["heritability"] = _frame([[tr, float(a), float(b)] for tr, a, b in
                                         zip(model.lhsVec, np.atleast_1d(out["h2_mean"]), np.atleast_1d(out["h2_sd"]))],
                                        ["Covariance", "Estimate", "SD"])
    if sample_files is not None:
        for f in sample_files:
            f.close()
    if write_results and pd is not None:         # JWAS.jl:479-482
        for key, value in output.items():
            value.to_csv(os.path.join(output_folder, key.replace(" ", "_") + ".txt"), index=False)
    model.output = output
    model.sol = out["mu"]
    return output
