/* The C-ABI entry points integration/updes_jax_ffi.cc calls, forwarded to callbacks registered from Python (the CPU
 * emulation of tests/cpu_abi_emulation.py) -- lets the compiled adapter run without libupdes_b200.so / a GPU.  Test
 * infrastructure; signatures are those of include/updes_b200.h. */
#include "updes_b200.h"

typedef int (*assemble_rows_fn)(int, double, int, int, const double *, const UpdesRows *, int64_t, int64_t, int, double *, int64_t, void *);
typedef int (*lu_create_fn)(UpdesLU **, int64_t, int64_t);
typedef int (*lu_destroy_fn)(UpdesLU *);
typedef int (*lu_factor_fn)(UpdesLU *, double *, int32_t *, int32_t *, void *);
typedef int (*lu_solve_fn)(UpdesLU *, const double *, const int32_t *, double *, int64_t, int, int, void *);
typedef size_t (*ws_fn)(int, int, int);
typedef int (*eval_jets_fn)(int, double, int, int, const double *, const double *, int64_t, int, const double *, int, const int32_t *, double *, double *, void *, void *);

static assemble_rows_fn cb_assemble_rows; static lu_create_fn cb_lu_create; static lu_destroy_fn cb_lu_destroy;
static lu_factor_fn cb_lu_factor; static lu_solve_fn cb_lu_solve; static ws_fn cb_ws; static eval_jets_fn cb_eval_jets;

void updes_mock_register(void *a, void *b, void *c, void *d, void *e, void *f, void *g) {
  cb_assemble_rows = (assemble_rows_fn)a; cb_lu_create = (lu_create_fn)b; cb_lu_destroy = (lu_destroy_fn)c;
  cb_lu_factor = (lu_factor_fn)d; cb_lu_solve = (lu_solve_fn)e; cb_ws = (ws_fn)f; cb_eval_jets = (eval_jets_fn)g;
}

int updes_assemble_rows(int k, double p, int N, int M, const double *ctr, const UpdesRows *rows, int64_t r0, int64_t nr, int mask,
                        double *out, int64_t ld, void *st) { return cb_assemble_rows(k, p, N, M, ctr, rows, r0, nr, mask, out, ld, st); }
int updes_lu_create(UpdesLU **h, int64_t n, int64_t ld) { return cb_lu_create(h, n, ld); }
int updes_lu_destroy(UpdesLU *h) { return cb_lu_destroy(h); }
int updes_lu_factor(UpdesLU *h, double *K, int32_t *ipiv, int32_t *info, void *st) { return cb_lu_factor(h, K, ipiv, info, st); }
int updes_lu_solve(UpdesLU *h, const double *LU, const int32_t *ipiv, double *B, int64_t ldb, int nrhs, int tr, void *st) {
  return cb_lu_solve(h, LU, ipiv, B, ldb, nrhs, tr, st);
}
size_t updes_eval_jets_workspace_bytes(int N, int npts, int nf) { return cb_ws(N, npts, nf); }
int updes_eval_jets(int k, double p, int N, int M, const double *ctr, const double *cf, int64_t ldc, int nf, const double *pts, int npts,
                    const int32_t *skip, double *jphi, double *jpol, void *ws, void *st) {
  return cb_eval_jets(k, p, N, M, ctr, cf, ldc, nf, pts, npts, skip, jphi, jpol, ws, st);
}
