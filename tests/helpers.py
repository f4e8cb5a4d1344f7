"""Shared builders for the parity tests (synthetic genotypes in the style of
benchmarks/bayesr_parity_common.jl:28-59: Binomial(2, f_j) codes via two Bernoullis)."""
import numpy as np


def make_codes(n, p, seed, missing=0.0, fmin=0.05, fmax=0.5):
    rng = np.random.default_rng(seed)
    f = rng.uniform(fmin, fmax, size=p)
    codes = (rng.random((n, p)) < f).astype(np.int8) + (rng.random((n, p)) < f).astype(np.int8)
    # keep every marker polymorphic
    for j in range(p):
        if codes[:, j].min() == codes[:, j].max():
            codes[0, j] = (codes[0, j] + 1) % 3
    if missing > 0:
        codes[rng.random((n, p)) < missing] = 9
    return codes


def make_phenotype(X, seed, nqtl=10, h2=0.5, ntraits=1):
    rng = np.random.default_rng(seed + 1000)
    n, p = X.shape
    ys = []
    for _ in range(ntraits):
        qtl = rng.choice(p, size=min(nqtl, p), replace=False)
        g = X[:, qtl].astype(np.float64) @ rng.normal(size=len(qtl))
        vg = g.var() if g.var() > 0 else 1.0
        e = rng.normal(size=n) * np.sqrt(vg * (1 - h2) / h2)
        ys.append(g + e + 3.0)
    return np.array(ys)


def uniform_starts(p, b):
    st = list(range(0, p, b)) + [p]
    return np.array(st, dtype=np.int64)


def canonical_sum_prod(a, b):
    """Same order as jw_k_chunk_prod/jw_k_chunk_final: 256-element chunks summed in index order, then
    groups of 256 chunk sums summed in index order, then the group sums added in order; binary64."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    v = a * b

    def level(x):
        pad = (-x.size) % 256
        if pad:
            x = np.concatenate([x, np.zeros(pad)])
        rows = x.reshape(-1, 256)
        acc = np.zeros(rows.shape[0])
        for c in range(256):          # sequential within a chunk, vectorised over chunks
            acc = acc + rows[:, c]
        return acc

    part = level(v)
    groups = level(part)
    assert groups.size <= 256
    total = 0.0
    for g in groups:
        total = total + g
    return float(total)


class Problem:
    def __init__(self, oracle, n, p, seed, missing=0.0, ntraits=1):
        self.n, self.p, self.t = n, p, ntraits
        self.codes = make_codes(n, p, seed, missing)
        self.packed = oracle.pack_codes(self.codes)
        self.means, self.xpx = oracle.marker_stats(self.packed, n)
        self.X = oracle.dense_centered(self.packed, n, self.means)
        y = make_phenotype(self.X, seed, ntraits=ntraits)
        self.y = y
        self.ycorr0 = (y - y.mean(axis=1, keepdims=True)).astype(np.float32).reshape(-1)
        self.vary = float(y.var())

    def fresh_state(self):
        tp = self.t * self.p
        return (self.ycorr0.copy(), np.zeros(tp, np.float32), np.zeros(tp, np.float32),
                np.zeros(tp, np.int32))
