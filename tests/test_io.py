"""libjwasio.so (include/jwas_io.h): genotype text files straight to the 2-bit marker-major image -- against the
numpy packer / pandas path on the same files, the reference's codec fixture (test_streaming_codec.jl:6-16, bit layout
streaming_genotypes.jl:622-627), error behaviour, and get_genotypes / prepare_streaming_genotypes through it."""
import os

import numpy as np
import pandas as pd
import pytest

import jwas_b200 as jw
from jwas_b200 import _io, api


def write_csv(path, codes, sep=",", floats=False, quoted=False, crlf=False, header=True):
    n, p = codes.shape
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        if header:
            hdr = ["ID"] + [f"m{j + 1}" for j in range(p)]
            f.write(sep.join(('"%s"' % h if quoted else h) for h in hdr) + nl)
        for i in range(n):
            vals = [("%.1f" % v if floats else str(v)) for v in codes[i]]
            f.write(('"id_%d"' % i if quoted else "id_%d" % i) + sep + sep.join(vals) + nl)


def random_codes(rng, n, p, missing=0.02):
    codes = rng.integers(0, 3, size=(n, p))
    codes[rng.random((n, p)) < missing] = 9
    return codes


@pytest.mark.parametrize("kw", [dict(), dict(floats=True), dict(quoted=True, crlf=True), dict(sep="\t"), dict(header=False)])
@pytest.mark.parametrize("n,p", [(7, 5), (401, 333), (64, 1), (1030, 70)])
def test_text_to_packed_equals_numpy_packer(tmp_path, kw, n, p):
    rng = np.random.default_rng(n + p)
    codes = random_codes(rng, n, p)
    path = str(tmp_path / "g.txt")
    write_csv(path, codes, **kw)
    obs, names, packed = _io.read_genotype_text(path, separator=kw.get("sep", ","), header=kw.get("header", True))
    c = codes.copy(); c[c == 9] = 3
    assert obs == [f"id_{i}" for i in range(n)] and names == [f"m{j + 1}" for j in range(p)]
    np.testing.assert_array_equal(packed, api._pack_codes(c.astype(np.uint8)))
    cnt = _io.packed_counts(packed, n)
    np.testing.assert_array_equal(cnt, np.stack([(c == 1).sum(0), (c == 2).sum(0), (c == 3).sum(0)], axis=1))
    rows = rng.permutation(n)[:max(1, n // 2)]
    np.testing.assert_array_equal(_io.packed_rows(packed, rows), api._pack_codes(c[rows].astype(np.uint8)))
    for th in (1, 3):
        np.testing.assert_array_equal(_io.read_genotype_text(path, kw.get("sep", ","), kw.get("header", True), nthreads=th)[2], packed)


def test_reference_codec_fixture(tmp_path):
    """test_streaming_codec.jl:6-16: four individuals x three markers with a missing call (9); individual i sits in
    byte (i-1)>>2 at bits ((i-1)&3)<<1, LSB first, code 3 = missing (streaming_genotypes.jl:622-627)."""
    path = str(tmp_path / "geno.csv")
    open(path, "w").write("ID,m1,m2,m3\na1,0,1,2\na2,1,9,0\na3,2,1,1\na4,0,2,9\n")
    obs, names, packed = _io.read_genotype_text(path)
    assert obs == ["a1", "a2", "a3", "a4"] and names == ["m1", "m2", "m3"]
    assert packed.shape == (3, 1)
    assert packed[:, 0].tolist() == [0 | 1 << 2 | 2 << 4 | 0 << 6, 1 | 3 << 2 | 1 << 4 | 2 << 6, 2 | 0 << 2 | 1 << 4 | 3 << 6]


def test_errors_and_missing_tokens(tmp_path):
    path = str(tmp_path / "bad.csv")
    open(path, "w").write("ID,m1,m2\na,0,1\nb,0.5,2\n")
    with pytest.raises(jw.JwasError, match=r"Only 0/1/2 genotypes.*row 2, marker 1"):
        _io.read_genotype_text(path)
    for body in ("a,0,1\nb,1\n", "a,0,1\nb,1,2,0\n"):
        open(path, "w").write("ID,m1,m2\n" + body)
        with pytest.raises(jw.JwasError, match="row 2 does not hold 2 genotype fields"):
            _io.read_genotype_text(path)
    open(path, "w").write("")
    with pytest.raises(jw.JwasError, match="Genotype data is empty"):
        _io.read_genotype_text(path)
    with pytest.raises(jw.JwasError, match="cannot open"):
        _io.read_genotype_text(str(tmp_path / "nowhere.csv"))
    open(path, "w").write("ID,m1,m2\na,0,1\nb,NA,\nc,NaN,9.0\n\n")          # blank trailing line, NA / empty / NaN / 9.0
    obs, names, packed = _io.read_genotype_text(path)
    assert obs == ["a", "b", "c"]
    assert packed[:, 0].tolist() == [0 | 3 << 2 | 3 << 4, 1 | 3 << 2 | 3 << 4]


def test_get_genotypes_from_text_equals_dataframe_path(tmp_path):
    rng = np.random.default_rng(5)
    codes = random_codes(rng, 203, 150, missing=0.03)
    codes[:, 7] = 1                      # fixed locus: removed by QC
    codes[:, 11] = 0; codes[0, 11] = 1   # MAF below 0.01: removed
    path = str(tmp_path / "g.csv")
    write_csv(path, codes)
    A = rng.random((150, 2))
    g1 = jw.get_genotypes(path, 1.0, method="BayesC", annotations=A)
    g2 = jw.get_genotypes(pd.read_csv(path), 1.0, method="BayesC", annotations=A)
    assert g1.nMarkers == g2.nMarkers == 148 and g1.obsID == g2.obsID and g1.markerID == g2.markerID
    np.testing.assert_array_equal(g1.packed, g2.packed)
    np.testing.assert_array_equal(g1.marker_means, g2.marker_means)
    np.testing.assert_array_equal(g1.alleleFreq, g2.alleleFreq)
    assert g1.sum2pq == g2.sum2pq
    np.testing.assert_array_equal(g1.annotations.design_matrix, g2.annotations.design_matrix)
    # means as the reference computes them: mean of the observed calls, Float32
    c = np.where(codes == 9, np.nan, codes).astype(np.float64)
    keep = [j for j in range(150) if j not in (7, 11)]
    np.testing.assert_allclose(g1.marker_means, np.nanmean(c[:, keep], axis=0), rtol=1e-6)


def test_prepare_streaming_genotypes_files(tmp_path):
    """streaming_genotypes.jl:819-877: <prefix>.jgb2 + side-cars; xpRinvx of the centred columns with missing calls at
    the mean (test_streaming_prepare_lowmem.jl:22-66 checks the same quantity)."""
    rng = np.random.default_rng(9)
    codes = random_codes(rng, 57, 40, missing=0.05)
    path = str(tmp_path / "g.csv")
    write_csv(path, codes)
    prefix = jw.prepare_streaming_genotypes(path, quality_control=False)
    be = jw.load_streaming_backend(prefix)
    c = codes.copy(); c[c == 9] = 3
    np.testing.assert_array_equal(be["packed"], api._pack_codes(c.astype(np.uint8)))
    x = np.where(codes == 9, np.nan, codes).astype(np.float64)
    m = np.nanmean(x, axis=0)
    xc = np.where(np.isnan(x), 0.0, x - m)
    np.testing.assert_allclose(np.fromfile(prefix + ".xpRinvx.f32", np.float32), (xc * xc).sum(axis=0), rtol=2e-6)
    np.testing.assert_allclose(np.fromfile(prefix + ".mean.f32", np.float32), m, rtol=1e-6)
    g = jw.get_genotypes(prefix, 1.0)
    assert g.nObs == 57 and g.nMarkers == 40
    np.testing.assert_array_equal(g.packed, be["packed"])


@pytest.mark.parametrize("kw", [dict(trail_nl=False), dict(pad_to=4096), dict(pad_to=4096, trail_nl=False), dict(spaces=True)])
@pytest.mark.parametrize("n,p", [(1, 1), (3, 1), (5, 2), (4, 20000), (4097, 5)])
def test_file_shapes_at_the_edges(tmp_path, kw, n, p):
    """No trailing newline, a file that ends exactly on a page boundary (the mapping has no slack behind the last
    field), padded fields, one row, one marker, a row count that is not a multiple of four."""
    rng = np.random.default_rng(n * 31 + p)
    codes = random_codes(rng, n, p, missing=0.05)
    lines = ["ID," + ",".join(f"m{j + 1}" for j in range(p))]
    for i in range(n):
        lines.append(f"i{i}," + ",".join((" %d " % v if kw.get("spaces") else str(v)) for v in codes[i]))
    end = "\n" if kw.get("trail_nl", True) else ""
    txt = "\n".join(lines) + end
    if kw.get("pad_to"):
        need = (-len(txt)) % kw["pad_to"]
        lines[-1] = "i" + "x" * need + lines[-1][1:]
        txt = "\n".join(lines) + end
        assert len(txt) % kw["pad_to"] == 0
    path = str(tmp_path / "f.csv")
    open(path, "w", newline="").write(txt)
    obs, names, packed = _io.read_genotype_text(path)
    c = codes.copy(); c[c == 9] = 3
    assert len(obs) == n and len(names) == p
    np.testing.assert_array_equal(packed, api._pack_codes(c.astype(np.uint8)))
