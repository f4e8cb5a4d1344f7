"""ctypes binding of libjwasb200.so (include/jwas_b200.h).  No CPU fallback: a missing
library or a machine without a CUDA device makes every compute call raise."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# JWAS_B200_LIB selects another build of the same library (A/B measurements of two commits on one box)
SO_PATH = os.environ.get("JWAS_B200_LIB") or os.path.join(_HERE, "libjwasb200.so")

SCHED_EXACT, SCHED_BLOCK, SCHED_INDEPENDENT = 0, 1, 2


class SweepStats(C.Structure):
    _fields_ = [
        ("ycorr_ss", C.c_double * 16), ("ycorr_sum", C.c_double * 4),
        ("alpha_ss", C.c_double * 16), ("beta_ss", C.c_double * 16),
        ("nnz_alpha", C.c_double * 4), ("sum_delta", C.c_double * 4),
        ("class_counts", C.c_double * 16), ("bayesr_ssq", C.c_double),
        ("ycorr_maxabs", C.c_double), ("scale_exp", C.c_int32), ("overflow", C.c_int32),
        ("n_active", C.c_int64), ("n_rounds", C.c_int64),
    ]


class JwasError(RuntimeError):
    """Mirrors the reference's ErrorException (error("...") in JWAS.jl)."""


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise JwasError(
            f"{SO_PATH} is missing: build it with `python jwas.jl_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, i64, i32, dbl, u64, u32 = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64, C.c_uint32
    sig = {
        "jwas_create": [i64, i64, i32, vp, i64, i32, C.POINTER(vp)],
        "jwas_create_synthetic": [i64, i64, i32, u64, dbl, i32, C.POINTER(vp)],
        "jwas_shard_range": [i64, i32, i32, C.POINTER(i64), C.POINTER(i64)],
        "jwas_create_shard": [i64, i64, i32, i64, i64, vp, i64, i32, C.POINTER(vp)],
        "jwas_create_synthetic_shard": [i64, i64, i32, u64, dbl, i64, i64, i32, C.POINTER(vp)],
        "jwas_set_marker_means": [vp, vp],
        "jwas_get_packed": [vp, vp, i64],
        "jwas_get_gram": [vp, i64, vp],
        "jwas_last_stream_kernel_ms": [vp, C.POINTER(i64)],
        "jwas_get_phase_ns": [vp, vp],
        "jwas_nccl_unique_id": [vp],
        "jwas_init_sharding": [vp, i32, i32, vp],
        "jwas_get_row_range": [vp, C.POINTER(i64), C.POINTER(i64)],
        "jwas_ipc_export": [vp, vp],
        "jwas_ipc_import": [vp, vp],
        "jwas_destroy": [vp],
        "jwas_device_count": [],
        "jwas_get_marker_stats": [vp, vp, vp],
        "jwas_set_blocks": [vp, vp, i64],
        "jwas_put_ycorr": [vp, vp], "jwas_get_ycorr": [vp, vp],
        "jwas_ycorr_sub_malpha": [vp],
        "jwas_shift_ycorr": [vp, i32, C.c_float, C.POINTER(dbl), C.POINTER(dbl)],
        "jwas_mul_alpha": [vp, i32, vp],
        "jwas_put_state": [vp, vp, vp, vp], "jwas_get_state": [vp, vp, vp, vp],
        "jwas_sweep_bayesabc": [vp, i32, dbl, vp, vp, u64, u32, vp, vp, C.POINTER(SweepStats)],
        "jwas_sweep_bayesc": [vp, i32, dbl, dbl, dbl, u64, u32, C.POINTER(SweepStats)],
        "jwas_sweep_bayesc_host": [vp, i32, dbl, dbl, dbl, u64, u32, vp, vp, vp, vp, C.POINTER(SweepStats)],
        "jwas_sweep_bayesr": [vp, i32, i32, dbl, dbl, vp, i32, vp, i32, u64, u32, vp, vp, C.POINTER(SweepStats)],
        "jwas_sweep_mt1": [vp, i32, vp, vp, i32, vp, i32, u64, u32, vp, vp, C.POINTER(SweepStats)],
        "jwas_sweep_mt2": [vp, i32, vp, vp, vp, u64, u32, vp, vp, C.POINTER(SweepStats)],
        "jwas_sweep_mega": [vp, i32, vp, vp, vp, u64, u32, vp, vp, C.POINTER(SweepStats)],
        "jwas_sample_bayesb_variances": [vp, dbl, dbl, u64, u32, vp],
        "jwas_fill_hyper": [vp, i32, dbl],
        "jwas_accumulate": [vp, dbl, i32],
        "jwas_get_means": [vp, vp, vp, vp],
        "jwas_kernel_launches": [vp],
        "jwas_set_option": [vp, C.c_char_p, i64],
        "jwas_last_sweep_ms": [vp],
        "jwas_stream": [vp],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.jwas_last_error.restype = C.c_char_p
    L.jwas_last_error.argtypes = []
    L.jwas_kernel_launches.restype = C.c_int64
    L.jwas_last_sweep_ms.restype = C.c_double
    L.jwas_stream.restype = C.c_void_p
    L.jwas_last_stream_kernel_ms.restype = C.c_double
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise JwasError(lib().jwas_last_error().decode() or f"libjwasb200 error {rc}")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def nccl_unique_id():
    out = np.zeros(128, np.uint8)
    _check(lib().jwas_nccl_unique_id(_p(out)))
    return out.tobytes()


def device_count():
    return lib().jwas_device_count()


def shard_range(n_obs, rank, world):
    """Rows [begin, end) stored by `rank` of a `world`-way row-sharded problem (jwas_shard_range)."""
    b, e = C.c_int64(), C.c_int64()
    _check(lib().jwas_shard_range(int(n_obs), int(rank), int(world), C.byref(b), C.byref(e)))
    return b.value, e.value


class GpuSweeper:
    """Owns one device-resident genotype matrix and the sampler state that lives beside it."""

    def __init__(self, packed, n_obs, n_traits=1, device=0, rows=None):
        """packed: (p, stride) uint8 .jgb2 image of all n_obs individuals, or -- with rows=(begin, end), the range
        shard_range() gives this rank -- of those rows only (a row shard; follow with init_sharding)."""
        self._h = C.c_void_p()
        self.starts = None
        if packed is None:
            return
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        assert packed.ndim == 2
        self.n, self.p, self.t = int(n_obs), int(packed.shape[0]), int(n_traits)
        if rows is None:
            _check(lib().jwas_create(self.n, self.p, self.t, _p(packed), packed.shape[1], device, C.byref(self._h)))
        else:
            _check(lib().jwas_create_shard(self.n, self.p, self.t, int(rows[0]), int(rows[1]), _p(packed), packed.shape[1],
                                           device, C.byref(self._h)))

    @classmethod
    def synthetic(cls, n_obs, n_markers, n_traits=1, seed=0, missing_rate=0.0, device=0, rows=None):
        """Synthetic genotypes generated on the device; rows=(begin, end) generates this rank's shard only (the
        same bytes the full matrix would hold in those rows)."""
        self = cls(None, 0)
        self.n, self.p, self.t = int(n_obs), int(n_markers), int(n_traits)
        if rows is None:
            _check(lib().jwas_create_synthetic(self.n, self.p, self.t, int(seed), float(missing_rate), device,
                                               C.byref(self._h)))
        else:
            _check(lib().jwas_create_synthetic_shard(self.n, self.p, self.t, int(seed), float(missing_rate),
                                                     int(rows[0]), int(rows[1]), device, C.byref(self._h)))
        return self

    def get_packed(self):
        """The packed rows this handle stores (all rows unless it is a shard)."""
        b, e = self.row_range()
        out = np.empty((self.p, (e - b + 3) // 4), np.uint8)
        _check(lib().jwas_get_packed(self._h, _p(out), out.shape[1]))
        return out

    def set_marker_means(self, means):
        m = _arr(means, np.float32); assert m.size == self.p
        _check(lib().jwas_set_marker_means(self._h, _p(m)))

    def get_gram(self, ib):
        b = int(self.starts[ib + 1] - self.starts[ib])
        out = np.empty((b, b), np.float32)
        _check(lib().jwas_get_gram(self._h, int(ib), _p(out)))
        return out

    def stream_kernel_ms(self):
        nl = C.c_int64()
        ms = lib().jwas_last_stream_kernel_ms(self._h, C.byref(nl))
        return ms, nl.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().jwas_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- setup
    def marker_stats(self):
        means = np.empty(self.p, np.float32); xpx = np.empty(self.p, np.float32)
        _check(lib().jwas_get_marker_stats(self._h, _p(means), _p(xpx)))
        return means, xpx

    def set_blocks(self, starts):
        st = np.ascontiguousarray(starts, dtype=np.int64)
        _check(lib().jwas_set_blocks(self._h, _p(st), len(st) - 1))
        self.starts = st

    def set_option(self, key, value):
        _check(lib().jwas_set_option(self._h, key.encode(), int(value)))

    # -- state
    def put_ycorr(self, y):
        y = _arr(y, np.float32); assert y.size == self.t * self.n
        _check(lib().jwas_put_ycorr(self._h, _p(y)))

    def get_ycorr(self, out=None):
        """out: optional caller-owned Float32 buffer (e.g. pinned memory) the device copies straight into."""
        y = np.empty(self.t * self.n, np.float32) if out is None else out
        assert y.dtype == np.float32 and y.size == self.t * self.n and y.flags.c_contiguous
        _check(lib().jwas_get_ycorr(self._h, _p(y)))
        return y

    def put_state(self, alpha=None, beta=None, delta=None):
        a, b, d = _arr(alpha, np.float32), _arr(beta, np.float32), _arr(delta, np.int32)
        for x in (a, b, d):
            assert x is None or x.size == self.t * self.p
        _check(lib().jwas_put_state(self._h, _p(a), _p(b), _p(d)))

    def get_state(self, out=None):
        """out: optional (alpha, beta, delta) caller-owned buffers (e.g. pinned memory) filled in place."""
        if out is None:
            a = np.empty(self.t * self.p, np.float32); b = np.empty_like(a); d = np.empty(self.t * self.p, np.int32)
        else:
            a, b, d = out
            assert a.dtype == np.float32 and b.dtype == np.float32 and d.dtype == np.int32
        _check(lib().jwas_get_state(self._h, _p(a), _p(b), _p(d)))
        return a, b, d

    def ycorr_sub_malpha(self):
        _check(lib().jwas_ycorr_sub_malpha(self._h))

    def shift_ycorr(self, trait, shift, want=True):
        if not want:
            _check(lib().jwas_shift_ycorr(self._h, trait, float(shift), None, None))
            return None
        s, ss = C.c_double(), C.c_double()
        _check(lib().jwas_shift_ycorr(self._h, trait, float(shift), C.byref(s), C.byref(ss)))
        return s.value, ss.value

    def mul_alpha(self, trait=0):
        out = np.empty(self.n, np.float32)
        _check(lib().jwas_mul_alpha(self._h, trait, _p(out)))
        return out

    # -- sweeps
    def sweep_bayesabc(self, schedule, vare, var_effects, pi, seed, it, u=None, z=None):
        st = SweepStats()
        ve, pv, uu, zz = _arr(var_effects, np.float64), _arr(pi, np.float64), _arr(u, np.float64), _arr(z, np.float64)
        _check(lib().jwas_sweep_bayesabc(self._h, schedule, float(vare), _p(ve), _p(pv), int(seed), int(it),
                                         _p(uu), _p(zz), C.byref(st)))
        return st

    def sweep_bayesc(self, schedule, vare, var_effect, pi, seed, it):
        st = SweepStats()
        _check(lib().jwas_sweep_bayesc(self._h, schedule, float(vare), float(var_effect), float(pi),
                                       int(seed), int(it), C.byref(st)))
        return st

    def sweep_bayesc_host(self, schedule, vare, var_effect, pi, seed, it, ycorr, alpha, beta, delta):
        """BayesABC!(…, yCorr, α, β, δ, …) on the caller's HOST arrays, mutated in place (BayesABC.jl:60-63)."""
        st = SweepStats()
        for a, dt, sz in ((ycorr, np.float32, self.n), (alpha, np.float32, self.p), (beta, np.float32, self.p),
                          (delta, np.int32, self.p)):
            assert a.dtype == dt and a.size == sz and a.flags.c_contiguous
        _check(lib().jwas_sweep_bayesc_host(self._h, schedule, float(vare), float(var_effect), float(pi), int(seed),
                                            int(it), _p(ycorr), _p(alpha), _p(beta), _p(delta), C.byref(st)))
        return st

    def sweep_bayesr(self, schedule, full_reps, vare, sigma_sq, pi, gamma, seed, it, u=None, z=None):
        st = SweepStats()
        pv, g = _arr(pi, np.float64), _arr(gamma, np.float64)
        uu, zz = _arr(u, np.float64), _arr(z, np.float64)
        _check(lib().jwas_sweep_bayesr(self._h, schedule, int(full_reps), float(vare), float(sigma_sq), _p(pv),
                                       int(pv.ndim == 2), _p(g), len(g), int(seed), int(it), _p(uu), _p(zz),
                                       C.byref(st)))
        return st

    def sweep_mt1(self, schedule, R, G, big_pi, seed, it, u=None, z=None):
        st = SweepStats()
        Rm, Gm, bp = _arr(R, np.float64), _arr(G, np.float64), _arr(big_pi, np.float64)
        uu, zz = _arr(u, np.float64), _arr(z, np.float64)
        _check(lib().jwas_sweep_mt1(self._h, schedule, _p(Rm), _p(Gm), int(Gm.ndim == 3), _p(bp),
                                    int(bp.ndim == 2), int(seed), int(it), _p(uu), _p(zz), C.byref(st)))
        return st

    def sweep_mt2(self, schedule, R, G, big_pi, seed, it, u=None, z=None):
        st = SweepStats()
        Rm, Gm, bp = _arr(R, np.float64), _arr(G, np.float64), _arr(big_pi, np.float64)
        uu, zz = _arr(u, np.float64), _arr(z, np.float64)
        _check(lib().jwas_sweep_mt2(self._h, schedule, _p(Rm), _p(Gm), _p(bp), int(seed), int(it), _p(uu), _p(zz),
                                    C.byref(st)))
        return st

    def sweep_mega(self, schedule, vare, var_effects, pi, seed, it, u=None, z=None):
        st = SweepStats()
        v, ve, pv = _arr(vare, np.float64), _arr(var_effects, np.float64), _arr(pi, np.float64)
        uu, zz = _arr(u, np.float64), _arr(z, np.float64)
        _check(lib().jwas_sweep_mega(self._h, schedule, _p(v), _p(ve), _p(pv), int(seed), int(it), _p(uu), _p(zz),
                                     C.byref(st)))
        return st

    def sample_bayesb_variances(self, df, scale, seed, it, want=False):
        out = np.empty(self.p, np.float64) if want else None
        _check(lib().jwas_sample_bayesb_variances(self._h, float(df), float(scale), int(seed), int(it), _p(out)))
        return out

    def fill_hyper(self, which, value):
        _check(lib().jwas_fill_hyper(self._h, {"var_effects": 0, "pi": 1}[which], float(value)))

    # -- posterior accumulators
    def accumulate(self, nsamples, bayesr=False):
        _check(lib().jwas_accumulate(self._h, float(nsamples), int(bayesr)))

    def get_means(self):
        tp = self.t * self.p
        ma = np.empty(tp, np.float32); ma2 = np.empty(tp, np.float32); md = np.empty(tp, np.float32)
        _check(lib().jwas_get_means(self._h, _p(ma), _p(ma2), _p(md)))
        return ma, ma2, md

    def init_sharding(self, rank, world, unique_id=None):
        uid = None if unique_id is None else np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        _check(lib().jwas_init_sharding(self._h, int(rank), int(world), _p(uid)))

    def ipc_export(self):
        out = np.zeros(64, np.uint8)
        _check(lib().jwas_ipc_export(self._h, _p(out)))
        return out.tobytes()

    def ipc_import(self, handles):
        buf = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
        _check(lib().jwas_ipc_import(self._h, _p(buf)))

    def row_range(self):
        b, e = C.c_int64(), C.c_int64()
        _check(lib().jwas_get_row_range(self._h, C.byref(b), C.byref(e)))
        return b.value, e.value

    def phase_ns(self):
        out = np.zeros(32, np.uint64)
        _check(lib().jwas_get_phase_ns(self._h, _p(out)))
        return out

    # -- introspection
    @property
    def kernel_launches(self):
        return lib().jwas_kernel_launches(self._h)

    @property
    def last_sweep_ms(self):
        return lib().jwas_last_sweep_ms(self._h)
