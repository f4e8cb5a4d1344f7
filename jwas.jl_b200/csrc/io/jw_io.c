/* libjwasio.so -- genotype text file -> 2-bit marker-major image (include/jwas_io.h).  Host C, OpenMP. */
#define _GNU_SOURCE
#include "../../../include/jwas_io.h"
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char jwio_err[512];
static char jwio_err_shared[512];
const char* jwio_last_error(void) { return jwio_err[0] ? jwio_err : jwio_err_shared; }
#define FAIL(...) do { snprintf(jwio_err, sizeof jwio_err, __VA_ARGS__); \
                       memcpy(jwio_err_shared, jwio_err, sizeof jwio_err); return 1; } while (0)

typedef struct { const char* p; int64_t size; int fd; } mapped;

static int map_file(const char* path, mapped* m) {
    m->fd = open(path, O_RDONLY);
    if (m->fd < 0) FAIL("cannot open %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(m->fd, &st)) { close(m->fd); FAIL("cannot stat %s", path); }
    m->size = (int64_t)st.st_size;
    if (m->size == 0) { close(m->fd); FAIL("Genotype data is empty."); }
    m->p = (const char*)mmap(NULL, (size_t)m->size, PROT_READ, MAP_PRIVATE, m->fd, 0);
    if (m->p == MAP_FAILED) { close(m->fd); FAIL("cannot map %s: %s", path, strerror(errno)); }
    madvise((void*)m->p, (size_t)m->size, MADV_SEQUENTIAL);
    return 0;
}
static void unmap_file(mapped* m) { munmap((void*)m->p, (size_t)m->size); close(m->fd); }

/* end of the line starting at `a` (position of '\n' or size) */
static inline int64_t line_end(const mapped* m, int64_t a) {
    const char* q = (const char*)memchr(m->p + a, '\n', (size_t)(m->size - a));
    return q ? (int64_t)(q - m->p) : m->size;
}
static inline int blank_line(const mapped* m, int64_t a, int64_t e) {
    for (int64_t i = a; i < e; ++i) if (m->p[i] != '\r' && m->p[i] != ' ' && m->p[i] != '\t') return 0;
    return 1;
}

/* starts[k] = offset of data row k; returns the number of data rows found (at most cap when starts != NULL) */
static int64_t data_rows(const mapped* m, int header, int64_t* starts, int64_t cap) {
    int64_t a = 0, n = 0;
    int skip = header ? 1 : 0;
    while (a < m->size) {
        int64_t e = line_end(m, a);
        if (!blank_line(m, a, e)) {
            if (skip) skip = 0;
            else { if (starts) { if (n < cap) starts[n] = a; } ++n; }
        }
        a = e + 1;
    }
    return n;
}

static int64_t count_fields(const mapped* m, int64_t a, int64_t e, char sep) {
    int64_t k = 1;
    for (int64_t i = a; i < e; ++i) k += (m->p[i] == sep);
    return k;
}

int jwio_csv_dims(const char* path, int separator, int header, int64_t* n_rows, int64_t* n_fields) {
    jwio_err[0] = 0;
    if (!path || !n_rows || !n_fields) FAIL("jwio_csv_dims: null argument");
    mapped m;
    if (map_file(path, &m)) return 1;
    int64_t first = -1;
    *n_rows = data_rows(&m, header, &first, 1);
    *n_fields = 0;
    if (*n_rows > 0) *n_fields = count_fields(&m, first, line_end(&m, first), (char)separator);
    unmap_file(&m);
    if (*n_rows == 0 || *n_fields < 2) FAIL("Genotype data is empty.");
    return 0;
}

/* one field [a, e) -> code 0..3, or -1 */
static inline int field_code(const char* p, int64_t a, int64_t e, double missing_value) {
    while (a < e && (p[a] == ' ' || p[a] == '"')) ++a;
    while (e > a && (p[e - 1] == ' ' || p[e - 1] == '"' || p[e - 1] == '\r')) --e;
    const int64_t len = e - a;
    if (len == 1 && p[a] >= '0' && p[a] <= '2' && !(missing_value == (double)(p[a] - '0'))) return p[a] - '0';
    if (len == 0) return 3;
    if (len > 63) return -1;
    char buf[64];
    memcpy(buf, p + a, (size_t)len); buf[len] = 0;
    if (!strcmp(buf, "NA") || !strcmp(buf, "NaN") || !strcmp(buf, "nan") || !strcmp(buf, "missing")) return 3;
    char* endp = NULL;
    const double v = strtod(buf, &endp);
    if (endp == buf || *endp != 0) return -1;
    if (v == missing_value || isnan(v)) return 3;
    if (v == 0.0) return 0;
    if (v == 1.0) return 1;
    if (v == 2.0) return 2;
    return -1;
}

int jwio_csv_pack(const char* path, int separator, int header, double missing_value,
                  int64_t n_rows, int64_t n_markers, uint8_t* packed, int64_t stride,
                  int64_t* id_begin, int64_t* id_end, int n_threads) {
    jwio_err[0] = 0;
    if (!path || !packed) FAIL("jwio_csv_pack: null argument");
    if (n_rows <= 0 || n_markers <= 0) FAIL("Genotype data is empty.");
    if (stride < (n_rows + 3) / 4) FAIL("jwio_csv_pack: stride_bytes below cld(n_rows,4)");
    mapped m;
    if (map_file(path, &m)) return 1;
    int64_t* starts = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_rows);
    if (!starts) { unmap_file(&m); FAIL("out of memory"); }
    const int64_t found = data_rows(&m, header, starts, n_rows);
    if (found != n_rows) {
        free(starts); unmap_file(&m);
        FAIL("jwio_csv_pack: the file holds %lld data rows, %lld expected", (long long)found, (long long)n_rows);
    }
    memset(packed, 0, (size_t)stride * (size_t)n_markers);
    const char sep = (char)separator;
    const int64_t groups = (n_rows + 3) / 4;
    int64_t bad_row = -1, bad_col = -1;
    volatile int bad_kind = 0;             /* 1 = field count, 2 = value */
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#endif
    (void)n_threads;
    /* One group = four consecutive individuals = one byte of every column.  A thread takes a tile of G consecutive
     * groups, parses their rows into a private G x n_markers byte tile (sequential stores while the text streams by),
     * then transposes the tile into the image in 64-column blocks: every column receives G contiguous bytes. */
    int64_t G = (8ll << 20) / n_markers;
    if (G > 64) G = 64;
    if (G < 1) G = 1;
    const int64_t tiles = (groups + G - 1) / G;
    int oom = 0;
#pragma omp parallel
    {
        uint8_t* tmp = (uint8_t*)malloc((size_t)G * (size_t)n_markers);
        uint8_t blk[64][64];
        if (!tmp) {
#pragma omp atomic write
            oom = 1;
        }
#pragma omp for schedule(dynamic, 1)
        for (int64_t t = 0; t < tiles; ++t) {
            if (bad_kind || !tmp) continue;
            const int64_t g0 = t * G;
            const int64_t gn = (g0 + G <= groups) ? G : groups - g0;
            memset(tmp, 0, (size_t)gn * (size_t)n_markers);
            const char* p = m.p;
            for (int64_t i = 4 * g0; i < 4 * (g0 + gn) && i < n_rows && !bad_kind; ++i) {
                const int64_t a0 = starts[i];
                const int64_t e = line_end(&m, a0);
                const char* q = (const char*)memchr(p + a0, sep, (size_t)(e - a0));
                if (!q) {
#pragma omp critical
                    { if (!bad_kind) { bad_kind = 1; bad_row = i; } }
                    break;
                }
                int64_t ia = a0, ie = (int64_t)(q - p);
                while (ia < ie && (p[ia] == ' ' || p[ia] == '"')) ++ia;
                while (ie > ia && (p[ie - 1] == ' ' || p[ie - 1] == '"')) --ie;
                if (id_begin) id_begin[i] = ia;
                if (id_end) id_end[i] = ie;
                int64_t a = (int64_t)(q - p) + 1;
                const int shift = 2 * (int)(i & 3);
                uint8_t* row = tmp + ((i >> 2) - g0) * n_markers;
                int short_row = 0;
                for (int64_t j = 0; j < n_markers; ++j) {
                    if (a > e) { short_row = 1; break; }
                    const int last = (j + 1 == n_markers);
                    int64_t fe;
                    /* fast path: a lone digit followed by the separator */
                    if (!last && a + 1 < e && p[a + 1] == sep) fe = a + 1;
                    else {
                        const char* s2 = (e > a) ? (const char*)memchr(p + a, sep, (size_t)(e - a)) : NULL;
                        if ((!last && !s2) || (last && s2)) { short_row = 1; break; }      /* too few / too many fields */
                        fe = last ? e : (int64_t)(s2 - p);
                    }
                    int code;
                    if (fe - a == 1 && p[a] >= '0' && p[a] <= '2' && missing_value != (double)(p[a] - '0')) code = p[a] - '0';
                    else code = field_code(p, a, fe, missing_value);
                    if (code < 0) {
#pragma omp critical
                        { if (!bad_kind) { bad_kind = 2; bad_row = i; bad_col = j; } }
                        break;
                    }
                    row[j] |= (uint8_t)(code << shift);
                    a = fe + 1;
                }
                if (short_row) {
#pragma omp critical
                    { if (!bad_kind) { bad_kind = 1; bad_row = i; } }
                }
            }
            if (bad_kind) continue;
            for (int64_t jb = 0; jb < n_markers; jb += 64) {
                const int64_t jn = (jb + 64 <= n_markers) ? 64 : n_markers - jb;
                for (int64_t gi = 0; gi < gn; ++gi) {
                    const uint8_t* src = tmp + gi * n_markers + jb;
                    for (int64_t jj = 0; jj < jn; ++jj) blk[jj][gi] = src[jj];
                }
                for (int64_t jj = 0; jj < jn; ++jj) memcpy(packed + (jb + jj) * stride + g0, blk[jj], (size_t)gn);
            }
        }
        free(tmp);
    }
    if (oom) { free(starts); unmap_file(&m); FAIL("out of memory"); }
    free(starts);
    unmap_file(&m);
    if (bad_kind == 1) FAIL("row %lld does not hold %lld genotype fields", (long long)(bad_row + 1), (long long)n_markers);
    if (bad_kind == 2) FAIL("Only 0/1/2 genotypes (and missing_value=%g) are supported in storage=:gpu (row %lld, marker %lld).",
                            missing_value, (long long)(bad_row + 1), (long long)(bad_col + 1));
    return 0;
}

int jwio_packed_counts(const uint8_t* packed, int64_t n_rows, int64_t n_markers, int64_t stride,
                       int64_t* counts, int n_threads) {
    jwio_err[0] = 0;
    if (!packed || !counts) FAIL("jwio_packed_counts: null argument");
    if (stride < (n_rows + 3) / 4) FAIL("jwio_packed_counts: stride_bytes below cld(n_rows,4)");
    /* per byte value: number of fields equal to 1, 2, 3 */
    uint8_t lut[256][3];
    for (int b = 0; b < 256; ++b) {
        lut[b][0] = lut[b][1] = lut[b][2] = 0;
        for (int k = 0; k < 4; ++k) { const int c = (b >> (2 * k)) & 3; if (c) ++lut[b][c - 1]; }
    }
    const int64_t full = n_rows / 4;
    const int tail = (int)(n_rows % 4);
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#endif
    (void)n_threads;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n_markers; ++j) {
        const uint8_t* col = packed + j * stride;
        int64_t c1 = 0, c2 = 0, c3 = 0;
        for (int64_t b = 0; b < full; ++b) { c1 += lut[col[b]][0]; c2 += lut[col[b]][1]; c3 += lut[col[b]][2]; }
        if (tail) {
            const uint8_t v = (uint8_t)(col[full] & ((1u << (2 * tail)) - 1u));
            c1 += lut[v][0]; c2 += lut[v][1]; c3 += lut[v][2];
        }
        counts[3 * j] = c1; counts[3 * j + 1] = c2; counts[3 * j + 2] = c3;
    }
    return 0;
}

int jwio_packed_select(uint8_t* packed, int64_t n_markers, int64_t stride, const int64_t* keep, int64_t n_keep) {
    jwio_err[0] = 0;
    if (!packed || (!keep && n_keep)) FAIL("jwio_packed_select: null argument");
    for (int64_t k = 0; k < n_keep; ++k) {
        if (keep[k] < 0 || keep[k] >= n_markers || (k && keep[k] <= keep[k - 1])) FAIL("jwio_packed_select: keep must be ascending marker indices");
        if (keep[k] != k) memmove(packed + k * stride, packed + keep[k] * stride, (size_t)stride);
    }
    return 0;
}

int jwio_packed_rows(const uint8_t* packed, int64_t n_markers, int64_t stride, const int64_t* rows, int64_t n_out,
                     uint8_t* out, int64_t out_stride, int n_threads) {
    jwio_err[0] = 0;
    if (!packed || !rows || !out) FAIL("jwio_packed_rows: null argument");
    if (out_stride < (n_out + 3) / 4) FAIL("jwio_packed_rows: out_stride_bytes below cld(n_out,4)");
    for (int64_t i = 0; i < n_out; ++i) if (rows[i] < 0 || (rows[i] >> 2) >= stride) FAIL("jwio_packed_rows: row index out of range");
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#endif
    (void)n_threads;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < n_markers; ++j) {
        const uint8_t* col = packed + j * stride;
        uint8_t* o = out + j * out_stride;
        memset(o, 0, (size_t)out_stride);
        for (int64_t i = 0; i < n_out; ++i) {
            const int64_t r = rows[i];
            const unsigned c = (col[r >> 2] >> ((r & 3) << 1)) & 3u;
            o[i >> 2] |= (uint8_t)(c << ((i & 3) << 1));
        }
    }
    return 0;
}
