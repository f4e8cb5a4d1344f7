"""ctypes binding of libjwasio.so (include/jwas_io.h): genotype text files straight to the 2-bit marker-major image,
call counts and row / column selection on the packed image -- the dense n x p matrix is never built."""
import ctypes as C
import mmap
import os

import numpy as np

from ._lib import JwasError

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libjwasio.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise JwasError(f"{SO_PATH} is missing: build it with `python jwas.jl_b200/build.py` (gcc).")
        L = C.CDLL(SO_PATH)
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        L.jwio_last_error.restype = C.c_char_p
        L.jwio_csv_dims.argtypes = [C.c_char_p, i32, i32, C.POINTER(i64), C.POINTER(i64)]
        L.jwio_csv_pack.argtypes = [C.c_char_p, i32, i32, C.c_double, i64, i64, vp, i64, vp, vp, i32]
        L.jwio_packed_counts.argtypes = [vp, i64, i64, i64, vp, i32]
        L.jwio_packed_select.argtypes = [vp, i64, i64, vp, i64]
        L.jwio_packed_rows.argtypes = [vp, i64, i64, vp, i64, vp, i64, i32]
        L.jwann_probit_step.argtypes = [i64, i32, vp, vp, i64, vp, vp, C.c_double, vp, vp, vp, vp, i32]
        L.jwann_probit_probability.argtypes = [vp, i64, i32, vp, i32]
        _lib = L
    return _lib


def available():
    return os.path.exists(SO_PATH)


def _check(rc):
    if rc:
        raise JwasError(lib().jwio_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def read_genotype_text(path, separator=",", header=True, missing_value=9.0, nthreads=0):
    """-> (obsID, markerID, packed (p, cld(n,4)) uint8).  readgenotypes.jl:296-330 for 0/1/2 files."""
    if len(separator) != 1:
        raise JwasError("separator must be a single character.")
    n = C.c_int64(0); f = C.c_int64(0)
    bpath = os.fsencode(path)
    _check(lib().jwio_csv_dims(bpath, ord(separator), int(bool(header)), C.byref(n), C.byref(f)))
    n, p = n.value, f.value - 1
    stride = (n + 3) // 4
    packed = np.empty((p, stride), dtype=np.uint8)
    ib = np.empty(n, np.int64); ie = np.empty(n, np.int64)
    _check(lib().jwio_csv_pack(bpath, ord(separator), int(bool(header)), float(missing_value), n, p, _p(packed), stride,
                               _p(ib), _p(ie), int(nthreads)))
    with open(path, "rb") as fh:
        mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            obs = [mm[a:b].decode() for a, b in zip(ib.tolist(), ie.tolist())]
            if header:
                first = mm[:mm.find(b"\n") if mm.find(b"\n") >= 0 else len(mm)].decode().rstrip("\r")
                names = [x.strip().strip('"') for x in first.split(separator)][1:]
                if len(names) != p:
                    raise JwasError(f"the header names {len(names)} markers, the first row holds {p}")
            else:
                names = [f"m{j + 1}" for j in range(p)]
        finally:
            mm.close()
    return obs, names, packed


def packed_counts(packed, n, nthreads=0):
    """(p, 3) int64: number of 1s, 2s and missing calls per marker among the first n individuals."""
    packed = np.ascontiguousarray(packed)
    p, stride = packed.shape
    out = np.empty((p, 3), np.int64)
    _check(lib().jwio_packed_counts(_p(packed), n, p, stride, _p(out), int(nthreads)))
    return out


def packed_rows(packed, rows, nthreads=0):
    """Rows `rows` of every column, as a new image (p, cld(len(rows),4))."""
    packed = np.ascontiguousarray(packed)
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    p, stride = packed.shape
    out = np.empty((p, (len(rows) + 3) // 4), np.uint8)
    _check(lib().jwio_packed_rows(_p(packed), p, stride, _p(rows), len(rows), _p(out), out.shape[1], int(nthreads)))
    return out


def probit_step(Xc, active, response, coeffs, prior_var, uniforms, normals, liability, mu, nthreads=0):
    """jwann_probit_step (include/jwas_io.h): one binary probit step of the annotation update.  Xc (m, k) with contiguous
    columns; active: int64 indices or None; response int32 (m); coeffs (k), liability (m), mu (m) float64, updated in
    place."""
    m, k = Xc.shape
    assert Xc.flags.f_contiguous and Xc.dtype == np.float64
    for a in (coeffs, liability, mu):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    assert response.dtype == np.int32 and response.flags.c_contiguous and len(response) == m
    act = None if active is None else np.ascontiguousarray(active, dtype=np.int64)
    n_act = m if act is None else len(act)
    u = np.ascontiguousarray(uniforms, dtype=np.float64); z = np.ascontiguousarray(normals, dtype=np.float64)
    assert len(u) >= n_act and len(z) >= k
    rc = lib().jwann_probit_step(m, k, _p(Xc), None if act is None else _p(act), n_act, _p(response), _p(coeffs),
                                 float(prior_var), _p(u), _p(z), _p(liability), _p(mu), int(nthreads))
    if rc:
        raise JwasError("jwann_probit_step failed (code %d)" % rc)


def probit_probability(mu, complement=False, nthreads=0):
    """clamp(Phi(mu)) (or of 1 - Phi(mu)) to [eps, 1 - eps]."""
    mu = np.ascontiguousarray(mu, dtype=np.float64)
    out = np.empty_like(mu)
    rc = lib().jwann_probit_probability(_p(mu), mu.size, int(bool(complement)), _p(out), int(nthreads))
    if rc:
        raise JwasError("jwann_probit_probability failed (code %d)" % rc)
    return out
