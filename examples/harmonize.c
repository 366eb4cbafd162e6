/* harmonize.c — the reference's driver loop (R/utils.R:15-46, harmonize()) written against the C ABI
 * alone: what any host language's binding does, without Python or R in between.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/harmonize.c -Lharmony_b200 -lharmony_b200 \
 *       -Wl,-rpath,$PWD/harmony_b200 -lm -o harmonize_demo && ./harmonize_demo [cells]
 *
 * Synthetic input: `cells` cells x 20 PCs in two batches whose means differ by a constant shift, K = 10.
 * Exit codes: 0 = converged or ran max_iter_harmony iterations, 3 = no handle (no CUDA device: there is no CPU
 * path), 4 = any other library error.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "harmony_b200.h"

static double uniform01(uint64_t* s) { /* xorshift64*, good enough for a demo */
  *s ^= *s >> 12;
  *s ^= *s << 25;
  *s ^= *s >> 27;
  return (double)((*s * 0x2545F4914F6CDD1DULL) >> 11) / 9007199254740992.0;
}

static int fail(hb_handle* h, const char* what) {
  fprintf(stderr, "%s: %s\n", what, h ? hb_last_error(h) : "no handle");
  if (h) hb_destroy(h);
  return 4;
}

int main(int argc, char** argv) {
  const int64_t N = argc > 1 ? atoll(argv[1]) : 20000;
  const int d = 20, K = 10, C = 1, max_iter_harmony = 10;
  const int32_t B_vec[1] = {2};
  /* harmony_options() defaults (R/harmony_option.R:33-52) and RunHarmony()'s (R/ui.R:91-107) */
  const double theta[2] = {2.0, 2.0}, lambda[3] = {0.0, 1.0, 1.0}, alpha = 0.2, block_size = 0.05;
  const double epsilon_cluster = 1e-3, epsilon_harmony = 1e-2, batch_proportion_cutoff = 1e-5;
  const int max_iter_cluster = 20;
  double sigma[10];
  for (int k = 0; k < K; ++k) sigma[k] = 0.1;

  hb_handle* h = NULL;
  if (hb_create(&h, -1) != 0) {
    fprintf(stderr, "hb_create: %s\n", h ? hb_last_error(h) : "no usable CUDA device (the library has no CPU path)");
    if (h) hb_destroy(h);
    return 3;
  }

  /* Z: d x N column-major (cells are columns), Phi as its row-index slot: one level per cell */
  double* Z = (double*)malloc(sizeof(double) * (size_t)d * (size_t)N);
  int32_t* phi_i = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  if (!Z || !phi_i) return fail(h, "malloc");
  uint64_t s = 88172645463325252ULL;
  for (int64_t n = 0; n < N; ++n) {
    const int batch = uniform01(&s) < 0.4;
    const int type = (int)(uniform01(&s) * 5.0);
    phi_i[n] = batch;
    for (int j = 0; j < d; ++j)
      Z[n * d + j] = cos(1.3 * (type + 1) * (j + 1)) + 0.3 * (uniform01(&s) - 0.5) + (batch ? 0.25 : 0.0);
  }

  if (hb_setup(h, Z, d, N, phi_i, B_vec, C, sigma, theta, lambda, alpha, max_iter_cluster, epsilon_cluster,
               epsilon_harmony, K, block_size, batch_proportion_cutoff, 0))
    return fail(h, "hb_setup");
  char warning[256];
  while (hb_pop_warning(h, warning, sizeof warning)) fprintf(stderr, "warning: %s\n", warning);
  if (hb_set_seed(h, 1)) return fail(h, "hb_set_seed");
  if (hb_init_cluster(h, NULL)) return fail(h, "hb_init_cluster"); /* native kmeans_centers (utils.cpp:10-64) */

  int converged = 0, iter = 0;
  for (iter = 1; iter <= max_iter_harmony && !converged; ++iter) { /* R/utils.R:22-41 */
    const int rc = hb_cluster(h, NULL);                            /* native update orders */
    if (rc < 0) {
      fprintf(stderr, "aborted\n");
      break;
    }
    if (rc > 0) return fail(h, "hb_cluster");
    if (hb_moe_correct_ridge(h)) return fail(h, "hb_moe_correct_ridge");
    converged = hb_check_convergence(h, 1);
    if (converged < 0) return fail(h, "hb_check_convergence");
  }

  double obj[64];
  const int64_t n_obj = hb_trace(h, HB_OBJECTIVE_HARMONY, obj, 64);
  for (int64_t i = 0; i < n_obj && i < 64; ++i) printf("objective_harmony[%lld] = %.6f\n", (long long)i, obj[i]);
  printf("%s after %d iteration(s), %lld kernel launches\n", converged ? "converged" : "stopped", iter - 1,
         (long long)hb_kernel_launches(h));

  double* Zc = (double*)malloc(sizeof(double) * (size_t)hb_field_size(h, HB_Z_CORR));
  if (!Zc || hb_get_field(h, HB_Z_CORR, Zc)) return fail(h, "hb_get_field(HB_Z_CORR)");
  double shift0 = 0, shift1 = 0, n0 = 0, n1 = 0; /* the batch shift of PC 1 before / after */
  double before0 = 0, before1 = 0;
  for (int64_t n = 0; n < N; ++n) {
    if (phi_i[n]) { shift1 += Zc[n * d]; before1 += Z[n * d]; n1 += 1; }
    else { shift0 += Zc[n * d]; before0 += Z[n * d]; n0 += 1; }
  }
  printf("mean(PC1 | batch 1) - mean(PC1 | batch 0): %.4f before, %.4f after\n", before1 / n1 - before0 / n0,
         shift1 / n1 - shift0 / n0);
  free(Zc);
  free(Z);
  free(phi_i);
  hb_destroy(h);
  return 0;
}
