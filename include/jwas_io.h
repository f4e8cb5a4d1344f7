/* jwas_io.h -- C ABI of libjwasio.so: the genotype-file side of the marker-sweep path.
 *
 * A delimited genotype text file (first field = individual ID, then one 0/1/2 call per marker; the reference's
 * get_genotypes input, markers/readgenotypes.jl:213-330) goes straight to the marker-major 2-bit image the
 * reference's storage=:stream backend keeps in its .jgb2 file (markers/streaming_genotypes.jl:520-660,
 * bit layout :622-627: individual i in byte i>>2, bits (i&3)<<1, code 3 = missing) -- which is also the layout
 * jwas_create (jwas_b200.h) uploads.  The dense n x p matrix is never materialised: the file is memory-mapped,
 * rows are parsed in parallel in groups of four (one byte of every column per group), and the per-marker call
 * counts that the QC / centring statistics need (readgenotypes.jl:372-401, streaming_genotypes.jl:560-585) are
 * taken from the packed image.  Plain C, host only, no CUDA.  0 = ok, otherwise jwio_last_error(). */
#ifndef JWAS_IO_H
#define JWAS_IO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* jwio_last_error(void);

/* Number of data rows (individuals; a header line is not counted, blank trailing lines are ignored) and of
 * fields per row including the ID, taken from the first data row. */
int jwio_csv_dims(const char* path, int separator, int header, int64_t* n_rows, int64_t* n_fields);

/* Parse and pack.  packed: n_markers columns of stride_bytes >= cld(n_rows,4) bytes, zero-filled by the callee.
 * A field equal to missing_value (or empty, "NA", "NaN", "missing") becomes code 3; anything else than 0, 1, 2
 * is an error naming the row and marker.  id_begin/id_end (n_rows each, may be NULL): byte offsets of every
 * row's ID field in the file (quotes stripped), so that the host language slices the IDs without re-parsing.
 * n_threads <= 0: all cores. */
int jwio_csv_pack(const char* path, int separator, int header, double missing_value,
                  int64_t n_rows, int64_t n_markers, uint8_t* packed, int64_t stride_bytes,
                  int64_t* id_begin, int64_t* id_end, int n_threads);

/* Per-marker call counts from the packed image: counts[3*j + 0..2] = number of 1s, of 2s, of missing calls
 * among the first n_rows individuals of column j. */
int jwio_packed_counts(const uint8_t* packed, int64_t n_rows, int64_t n_markers, int64_t stride_bytes,
                       int64_t* counts, int n_threads);

/* Keep the columns listed in `keep` (ascending, n_keep entries), in place: the QC filter applied to the image. */
int jwio_packed_select(uint8_t* packed, int64_t n_markers, int64_t stride_bytes, const int64_t* keep, int64_t n_keep);

/* Rows `rows` (n_out entries, any order, repeats allowed) of every column -> out (n_markers columns of
 * out_stride_bytes >= cld(n_out,4)): aligning genotypes to phenotyped / output individuals (JWAS.jl:381-402,
 * tools4genotypes.jl:288-296) without unpacking. */
int jwio_packed_rows(const uint8_t* packed, int64_t n_markers, int64_t stride_bytes, const int64_t* rows, int64_t n_out,
                     uint8_t* out, int64_t out_stride_bytes, int n_threads);

/* ---- annotation prior update (MCMC/annotation_updates.jl): the O(markers x annotations) part, threaded ----------
 * One binary probit step (:43-123) on the markers listed in active[0..n_active) (NULL: all m markers).
 * Xc: m x k design matrix, COLUMN-major (first column = intercept).  response (m): the step's indicator (non-zero = 1).
 * mu (m): linear predictor on entry (read at the active markers), Xc * coeffs for ALL markers on exit.
 * liability (m): written at the active markers.  coeffs (k): updated in place.  uniforms (n_active) and normals (k)
 * come from the host's generator (one uniform per truncated-normal draw, one normal per coefficient, in this order).
 * Deterministic for any thread count (fixed chunks, partial sums added in chunk order). */
int jwann_probit_step(int64_t m, int k, const double* Xc, const int64_t* active, int64_t n_active,
                      const int32_t* response, double* coeffs, double prior_var,
                      const double* uniforms, const double* normals,
                      double* liability, double* mu, int n_threads);
/* prob[j] = clamp(Phi(mu[j]), eps, 1 - eps), or clamp(1 - Phi(mu[j]), ...) with complement (:177-189, :260-304) */
int jwann_probit_probability(const double* mu, int64_t m, int complement, double* prob, int n_threads);
double jwann_phi_inv(double p);
double jwann_phi_cdf(double x);

#ifdef __cplusplus
}
#endif
#endif
