/* Plain-C consumer of include/arrowgpu.h: proves the boundary is a C ABI (no C++/CUDA/torch types)
 * that a cgo / JNI / ctypes binding can link.  Exit code 0 = all checks passed on a GPU,
 * 77 = library loaded but no usable device (every compute call refused with AG_ERR_CUDA). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "arrowgpu.h"

#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      char msg[512];                                                         \
      ag_last_error(msg, sizeof msg);                                        \
      fprintf(stderr, "FAIL %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, msg); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(void) {
  enum { N = 100000 };
  int ndev = 0;
  if (ag_device_count(&ndev) != AG_OK || ndev == 0) {
    double x[4] = {1, 2, 3, 4}, r = -1;
    ag_status st = ag_sum_f64(x, 4, &r);
    if (st != AG_ERR_CUDA || r != -1) { fprintf(stderr, "expected AG_ERR_CUDA without a device, got %d\n", st); return 1; }
    printf("no device: compute refused with AG_ERR_CUDA (no CPU fallback)\n");
    return 77;
  }
  CHECK(ag_init(-1) == AG_OK);
  double *a = malloc(N * sizeof *a), *b = malloc(N * sizeof *b), *c = malloc(N * sizeof *c);
  int64_t* v = malloc(N * sizeof *v);
  int64_t* out = malloc(N * sizeof *out);
  int32_t* idx = malloc(N * sizeof *idx);
  uint8_t* mask = calloc(N / 8 + 8, 1);
  for (int i = 0; i < N; ++i) { a[i] = i; b[i] = 2.0 * i; v[i] = i % 100; idx[i] = (int32_t)((i * 7919LL) % N); }

  /* arrow/math Sum: sum(0..N-1) */
  double s = 0;
  CHECK(ag_sum_f64(a, N, &s) == AG_OK);
  CHECK(s == (double)N * (N - 1) / 2);
  /* compute.Add(float64) through the reference-shaped entry point */
  CHECK(ag_arith_binary(AG_TYPE_FLOAT64, AG_OP_ADD_CHECKED, a, b, c, N) == AG_OK);
  for (int i = 0; i < N; ++i) CHECK(c[i] == 3.0 * i);
  /* greater(int64, 89) -> bitmap at bit offset 0, then filter */
  int64_t scalar = 89;
  CHECK(ag_cmp_gt_as(AG_TYPE_INT64, v, &scalar, mask, N, 0) == AG_OK);
  int64_t out_len = 0;
  CHECK(ag_filter_primitive(64, v, NULL, 0, mask, NULL, 0, N, AG_DROP_NULLS, out, NULL, &out_len, NULL) == AG_OK);
  CHECK(out_len == N / 10);
  for (int64_t i = 0; i < out_len; ++i) CHECK(out[i] > 89);
  /* take + bounds error */
  int64_t bad_pos = 0, bad_idx = 0;
  CHECK(ag_take_primitive(64, v, NULL, 0, N, 32, 1, idx, NULL, 0, N, 1, out, NULL, NULL, &bad_pos, &bad_idx) == AG_OK);
  for (int i = 0; i < N; ++i) CHECK(out[i] == v[idx[i]]);
  idx[1234] = N;
  CHECK(ag_take_primitive(64, v, NULL, 0, N, 32, 1, idx, NULL, 0, N, 1, out, NULL, NULL, &bad_pos, &bad_idx) == AG_ERR_INDEX);
  CHECK(bad_pos == 1234 && bad_idx == N);
  /* checked int64 add: INT64_MAX + 1 -> "overflow" */
  int64_t big[2] = {INT64_MAX, 1}, one[2] = {1, 1}, res[2], first_bad = 0;
  CHECK(ag_arith_checked(AG_TYPE_INT64, AG_OP_ADD_CHECKED, AG_SHAPE_AA, big, NULL, 0, one, NULL, 0, res, 2, &first_bad) == AG_ERR_INVALID);
  char msg[64];
  ag_last_error(msg, sizeof msg);
  CHECK(strcmp(msg, "overflow") == 0 && first_bad == 0);
  printf("abi_smoke ok: %llu kernel launches, %s\n", (unsigned long long)ag_kernel_launch_count(), ag_version());
  free(a); free(b); free(c); free(v); free(out); free(idx); free(mask);
  return 0;
}
