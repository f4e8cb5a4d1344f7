"""ctypes binding of the CPU oracle (oracle/libjwas_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (jwas.jl_b200/) never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libjwas_oracle.so")

METHOD_ABC, METHOD_R, METHOD_MT1, METHOD_MT2, METHOD_MEGA = 0, 1, 2, 3, 4


def build(force=False):
    # make is mtime-aware: a library older than its sources is rebuilt, an up-to-date one costs nothing
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    except (OSError, subprocess.CalledProcessError):
        if force or not os.path.exists(_SO):
            raise
    return _SO


class _SweepArgs(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("p", C.c_int64), ("stride", C.c_int64),
        ("packed", C.c_void_p), ("means", C.c_void_p), ("xpx", C.c_void_p),
        ("starts", C.c_void_p), ("nblocks", C.c_int64),
        ("nreps_mode", C.c_int), ("independent", C.c_int),
        ("method", C.c_int), ("ntraits", C.c_int),
        ("ycorr", C.c_void_p), ("alpha", C.c_void_p), ("beta", C.c_void_p), ("delta", C.c_void_p),
        ("vare", C.c_double), ("varEffects", C.c_void_p), ("pi", C.c_void_p), ("per_marker_pi", C.c_int),
        ("sigmaSq", C.c_double), ("gamma", C.c_void_p), ("nclasses", C.c_int),
        ("Rmat", C.c_void_p), ("Gmat", C.c_void_p), ("per_marker_G", C.c_int),
        ("bigPi", C.c_void_p),
        ("seed", C.c_uint64), ("iter", C.c_uint32),
        ("u", C.c_void_p), ("z", C.c_void_p),
        ("overflow", C.c_int), ("scale_exp", C.c_int),
        ("row_begin", C.c_int64), ("row_end", C.c_int64),
        ("allreduce", C.c_void_p), ("ctx", C.c_void_p),
        ("lag", C.c_int),
    ]


ALLREDUCE_CB = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int64), C.c_int64)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.jwo_sdot.restype = C.c_float
        _lib.jwo_bayesr_block_nreps.argtypes = [C.c_int64] * 3
        _lib.jwo_validate_block_starts.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        _lib.jwo_sweep_contract.argtypes = [C.POINTER(_SweepArgs)]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- codec / stats
def pack_codes(codes):
    """codes: (n, p) integer array, 0/1/2, anything else = missing. Returns (p, stride) uint8."""
    codes = np.asarray(codes)
    n, p = codes.shape
    cm = np.asfortranarray(codes.astype(np.int8))
    stride = (n + 3) // 4
    packed = np.zeros((p, stride), dtype=np.uint8)
    lib().jwo_pack_codes(_p(cm), C.c_int64(n), C.c_int64(p), _p(packed), C.c_int64(stride))
    return packed


def decode_marker(packed, n, j, mean, centered=True):
    dest = np.empty(n, dtype=np.float32)
    col = packed[j]
    lib().jwo_decode_marker(_p(col), C.c_int64(n), C.c_float(mean), C.c_int(int(centered)), _p(dest))
    return dest


def mul_alpha(packed, n, means, alpha):
    p, stride = packed.shape
    out = np.empty(n, dtype=np.float32)
    lib().jwo_mul_alpha(_p(packed), C.c_int64(n), C.c_int64(p), C.c_int64(stride),
                        _p(_f32(means)), _p(_f32(alpha)), _p(out))
    return out


def marker_stats_ref(packed, n, center=True):
    p, stride = packed.shape
    means = np.empty(p, np.float32); xpx = np.empty(p, np.float32); af = np.empty(p, np.float32)
    lib().jwo_marker_stats_ref(_p(packed), C.c_int64(n), C.c_int64(p), C.c_int64(stride),
                               C.c_int(int(center)), _p(means), _p(xpx), _p(af))
    return means, xpx, af


def marker_stats(packed, n):
    p, stride = packed.shape
    means = np.empty(p, np.float32); xpx = np.empty(p, np.float32)
    lib().jwo_marker_stats(_p(packed), C.c_int64(n), C.c_int64(p), C.c_int64(stride), _p(means), _p(xpx))
    return means, xpx


def gram_block(packed, n, means, j0, b):
    p, stride = packed.shape
    G = np.empty((b, b), np.float32)
    lib().jwo_gram_block(_p(packed), C.c_int64(n), C.c_int64(stride), _p(_f32(means)),
                         C.c_int64(j0), C.c_int64(b), _p(G))
    return G


def dense_centered(packed, n, means):
    """(n, p) Fortran-ordered Float32 matrix exactly as decode_marker! would produce it."""
    p = packed.shape[0]
    X = np.empty((n, p), dtype=np.float32, order="F")
    for j in range(p):
        X[:, j] = decode_marker(packed, n, j, float(means[j]))
    return X


# ---------------------------------------------------------------- reference-arithmetic samplers
def bayesabc_ref(X, xpx, ycorr, alpha, beta, delta, vare, varEffects, pi, u, z, nthreads=1):
    n, p = X.shape
    assert X.flags.f_contiguous and X.dtype == np.float32
    lib().jwo_bayesabc_ref(_p(X), C.c_int64(n), C.c_int64(p), _p(_f32(xpx)), _p(ycorr), _p(alpha), _p(beta),
                           _p(delta), C.c_float(vare), _p(_f32(varEffects)), _p(_f64(pi)), _p(_f64(u)),
                           _p(_f64(z)), C.c_int(nthreads))


def bayesl_ref(X, xpx, ycorr, alpha, gamma, v_res, v_eff, z, nthreads=1):
    """BayesL! (BayesC0L.jl:25-47); gamma = [1.0] is BayesC0! (RR-BLUP)."""
    n, p = X.shape
    assert X.flags.f_contiguous and X.dtype == np.float32
    g = _f64(gamma)
    assert len(g) in (1, p)
    lib().jwo_bayesl_ref(_p(X), C.c_int64(n), C.c_int64(p), _p(_f32(xpx)), _p(ycorr), _p(alpha), _p(g),
                         C.c_int64(len(g)), C.c_float(v_res), C.c_float(v_eff), _p(_f64(z)), C.c_int(nthreads))


def mtbayesl_ref(X, xpx, ycorr, alpha, gamma, R, G, z):
    """MTBayesL! (MTBayesC0L.jl:11-58); gamma = [1.0] is MTBayesC0! (multi-trait RR-BLUP).  alpha: (t, p) Float32."""
    n, p = X.shape
    t = alpha.shape[0]
    g = _f64(gamma)
    assert len(g) in (1, p) and alpha.dtype == np.float32 and alpha.flags.c_contiguous
    lib().jwo_mtbayesl_ref(_p(X), C.c_int64(n), C.c_int64(p), C.c_int(t), _p(_f32(xpx)), _p(ycorr), _p(alpha), _p(g),
                           C.c_int64(len(g)), _p(_f64(R)), _p(_f64(G)), _p(_f64(z)))


def bayesabc_streaming_ref(packed, n, means, xpx, ycorr, alpha, beta, delta, vare, varEffects, pi, u, z):
    p, stride = packed.shape
    lib().jwo_bayesabc_streaming_ref(_p(packed), C.c_int64(n), C.c_int64(p), C.c_int64(stride),
                                     _p(_f32(means)), _p(_f32(xpx)), _p(ycorr), _p(alpha), _p(beta),
                                     _p(delta), C.c_float(vare), _p(_f32(varEffects)), _p(_f64(pi)),
                                     _p(_f64(u)), _p(_f64(z)))


def bayesabc_block_ref(X, xpx, starts, nreps, independent, ycorr, alpha, beta, delta, vare, varEffects,
                       pi, u, z):
    n, p = X.shape
    st = np.ascontiguousarray(starts, dtype=np.int64)
    lib().jwo_bayesabc_block_ref(_p(X), C.c_int64(n), C.c_int64(p), _p(_f32(xpx)), _p(st),
                                 C.c_int64(len(st) - 1), C.c_int(nreps), C.c_int(int(independent)),
                                 _p(ycorr), _p(alpha), _p(beta), _p(delta), C.c_float(vare),
                                 _p(_f32(varEffects)), _p(_f64(pi)), _p(_f64(u)), _p(_f64(z)))


def bayesr_ref(X, xpx, ycorr, alpha, delta, vare, sigmaSq, pi, gamma, u, z, nthreads=1):
    n, p = X.shape
    pi = _f64(pi); gamma = _f64(gamma)
    lib().jwo_bayesr_ref(_p(X), C.c_int64(n), C.c_int64(p), _p(_f32(xpx)), _p(ycorr), _p(alpha), _p(delta),
                         C.c_float(vare), C.c_float(sigmaSq), _p(pi), C.c_int(int(pi.ndim == 2)),
                         _p(gamma), C.c_int(len(gamma)), _p(_f64(u)), _p(_f64(z)), C.c_int(nthreads))


def mtbayesabc_I_ref(X, xpx, ycorr, alpha, beta, delta, R, G, bigPi, u, z):
    n, p = X.shape
    t = alpha.shape[0]
    G = _f64(G); bigPi = _f64(bigPi)
    lib().jwo_mtbayesabc_I_ref(_p(X), C.c_int64(n), C.c_int64(p), C.c_int(t), _p(_f32(xpx)), _p(ycorr),
                               _p(alpha), _p(beta), _p(delta), _p(_f64(R)), _p(G), C.c_int(int(G.ndim == 3)),
                               _p(bigPi), C.c_int(int(bigPi.ndim == 2)), _p(_f64(u)), _p(_f64(z)))


def mtbayesabc_II_ref(X, xpx, ycorr, alpha, beta, delta, R, G, bigPi, u, z2):
    n, p = X.shape
    lib().jwo_mtbayesabc_II_ref(_p(X), C.c_int64(n), C.c_int64(p), _p(_f32(xpx)), _p(ycorr), _p(alpha),
                                _p(beta), _p(delta), _p(_f64(R)), _p(_f64(G)), _p(_f64(bigPi)),
                                _p(_f64(u)), _p(_f64(z2)))


def bayesr_block_ref(X, xpx, starts, nreps, independent, ycorr, alpha, delta, vare, sigmaSq, pi, gamma, u, z):
    n, p = X.shape
    st = np.ascontiguousarray(starts, dtype=np.int64)
    pi = _f64(pi); gamma = _f64(gamma)
    lib().jwo_bayesr_block_ref(_p(X), C.c_int64(n), C.c_int64(p), _p(_f32(xpx)), _p(st), C.c_int64(len(st) - 1),
                               C.c_int(nreps), C.c_int(int(independent)), _p(ycorr), _p(alpha), _p(delta),
                               C.c_float(vare), C.c_float(sigmaSq), _p(pi), C.c_int(int(pi.ndim == 2)),
                               _p(gamma), C.c_int(len(gamma)), _p(_f64(u)), _p(_f64(z)))


def mtbayesabc_block_ref(X, xpx, starts, nreps, independent, sampler, ycorr, alpha, beta, delta, R, G, bigPi, u, z):
    """sampler: 1 or 2.  alpha/beta/delta: (t, p) Float32; ycorr (t*n,)."""
    n, p = X.shape
    t = alpha.shape[0]
    st = np.ascontiguousarray(starts, dtype=np.int64)
    lib().jwo_mtbayesabc_block_ref(_p(X), C.c_int64(n), C.c_int64(p), C.c_int(t), C.c_int(sampler), _p(_f32(xpx)),
                                   _p(st), C.c_int64(len(st) - 1), C.c_int(nreps), C.c_int(int(independent)),
                                   _p(ycorr), _p(alpha), _p(beta), _p(delta), _p(_f64(R)), _p(_f64(G)),
                                   _p(_f64(bigPi)), _p(_f64(u)), _p(_f64(z)))


def bayesr_block_nreps(it, burnin, bs):
    return lib().jwo_bayesr_block_nreps(it, burnin, bs)


def validate_block_starts(starts, nmarkers):
    st = np.ascontiguousarray(starts, dtype=np.int64)
    return lib().jwo_validate_block_starts(_p(st), C.c_int64(len(st)), C.c_int64(nmarkers))


def bayesr_sigma_sufficient_statistics(alpha, delta, gamma):
    ssq = C.c_double(); nnz = C.c_int64()
    a = _f32(alpha); d = np.ascontiguousarray(delta, np.int32); g = _f64(gamma)
    lib().jwo_bayesr_sigma_sufficient_statistics(_p(a), _p(d), _p(g), C.c_int64(len(a)),
                                                 C.byref(ssq), C.byref(nnz))
    return ssq.value, nnz.value


def bayesb_variances(beta, df, scale, seed, it):
    b = _f32(beta)
    ve = np.empty(len(b), np.float64)
    lib().jwo_bayesb_variances(_p(b), C.c_int64(len(b)), C.c_double(df), C.c_double(scale), C.c_uint64(seed),
                               C.c_uint32(it), _p(ve))
    return ve


def max_threads():
    return lib().jwo_max_threads()


# ---------------------------------------------------------------- contract sweep
def sweep_contract(packed, n, means, xpx, starts, ycorr, alpha, beta, delta, *, method=METHOD_ABC,
                   nreps_mode=0, independent=False, vare=1.0, varEffects=None, pi=None,
                   sigmaSq=0.0, gamma=None, R=None, G=None, bigPi=None, seed=0, it=1, u=None, z=None,
                   row_range=None, allreduce=None, lag=0):
    """State arrays are modified in place: ycorr (t*n,) f32, alpha/beta (t*p,) f32, delta (t*p,) i32.
    starts: 0-based block boundaries of length nblocks+1."""
    p, stride = packed.shape
    a = _SweepArgs()
    keep = []

    def hold(x, dt):
        if x is None:
            return None
        x = np.ascontiguousarray(x, dtype=dt); keep.append(x); return _p(x)

    t = ycorr.size // n
    for arr, dt in ((ycorr, np.float32), (alpha, np.float32), (beta, np.float32), (delta, np.int32)):
        assert arr is None or (arr.dtype == dt and arr.flags.c_contiguous)
    a.n, a.p, a.stride = n, p, stride
    a.packed = _p(packed); a.means = hold(means, np.float32); a.xpx = hold(xpx, np.float32)
    st = np.ascontiguousarray(starts, dtype=np.int64); keep.append(st)
    a.starts = _p(st); a.nblocks = len(st) - 1
    a.nreps_mode = int(nreps_mode); a.independent = int(independent)
    a.method = method; a.ntraits = t
    a.ycorr = _p(ycorr); a.alpha = _p(alpha); a.beta = _p(beta); a.delta = _p(delta)
    a.vare = float(vare)
    a.varEffects = hold(varEffects, np.float64)
    if pi is not None:
        pi_arr = np.ascontiguousarray(pi, dtype=np.float64); keep.append(pi_arr)
        a.pi = _p(pi_arr); a.per_marker_pi = int(method == METHOD_R and pi_arr.ndim == 2)
    a.sigmaSq = float(sigmaSq)
    if gamma is not None:
        g = np.ascontiguousarray(gamma, dtype=np.float64); keep.append(g)
        a.gamma = _p(g); a.nclasses = len(g)
    a.Rmat = hold(R, np.float64)
    if G is not None:
        Ga = np.ascontiguousarray(G, dtype=np.float64); keep.append(Ga)
        a.Gmat = _p(Ga); a.per_marker_G = int(Ga.ndim == 3)
    if bigPi is not None:
        bp = np.ascontiguousarray(bigPi, dtype=np.float64); keep.append(bp)
        a.bigPi = _p(bp)
        if method in (METHOD_MT1, METHOD_MT2):
            a.per_marker_pi = int(bp.ndim == 2)
    a.seed = int(seed); a.iter = int(it)
    a.u = hold(u, np.float64); a.z = hold(z, np.float64)
    if row_range is not None:
        a.row_begin, a.row_end = int(row_range[0]), int(row_range[1])
    if allreduce is not None:
        def _cb(ctx, buf, count):
            arr = np.ctypeslib.as_array(buf, shape=(count,))
            allreduce(arr)
        cb = ALLREDUCE_CB(_cb); keep.append(cb)
        a.allreduce = C.cast(cb, C.c_void_p)
    a.lag = int(lag)
    rc = lib().jwo_sweep_contract(C.byref(a))
    return rc, a.scale_exp
