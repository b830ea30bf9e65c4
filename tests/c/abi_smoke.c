/* Plain-C consumer of libtdvp_b200: proves that include/tdvp_b200.h is valid C, that every declared entry point
 * links, and (when run on a GPU box) that the library works without Python: one small ZGEMM checked on the host.
 *   gcc -std=c11 -I include -I /usr/local/cuda/include tests/c/abi_smoke.c -L pytdscf_b200 -ltdvp_b200 \
 *       -L /usr/local/cuda/lib64 -lcudart -o abi_smoke */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "tdvp_b200.h"

/* taking the address of every entry point makes the link step fail if a declaration has no definition */
static const void* const entry_points[] = {
    (const void*)tdvp_create, (const void*)tdvp_destroy, (const void*)tdvp_last_error, (const void*)tdvp_abi_version,
    (const void*)tdvp_launch_count, (const void*)tdvp_get_stats, (const void*)tdvp_reset_stats, (const void*)tdvp_gemm_profile,
    (const void*)tdvp_profile_json, (const void*)tdvp_heff_apply, (const void*)tdvp_keff_apply, (const void*)tdvp_env_update,
    (const void*)tdvp_set_krylov_size, (const void*)tdvp_krylov_expm, (const void*)tdvp_lanczos_eigvec, (const void*)tdvp_qr_shift, (const void*)tdvp_absorb,
    (const void*)tdvp_svd_truncate, (const void*)tdvp_svd, (const void*)tdvp_pinv, (const void*)tdvp_inner,
    (const void*)tdvp_overlap_site, (const void*)tdvp_zgemm, (const void*)tdvp_set_gemm_config};

int main(int argc, char** argv) {
  printf("abi_version=%d entry_points=%d\n", tdvp_abi_version(), (int)(sizeof(entry_points) / sizeof(entry_points[0])));
  if (argc > 1 && argv[1][0] == 'l') return 0; /* link check only (no GPU) */
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fprintf(stderr, "no CUDA device\n"); return 2; }
  tdvp_handle_t h = NULL;
  if (tdvp_create(0, NULL, &h) != 0) { fprintf(stderr, "tdvp_create failed\n"); return 3; }
  enum { M = 37, N = 29, K = 41 };
  tdvp_c128 *A = malloc(sizeof(tdvp_c128) * M * K), *B = malloc(sizeof(tdvp_c128) * K * N), *C = malloc(sizeof(tdvp_c128) * M * N);
  for (int i = 0; i < M * K; ++i) { A[i].re = sin(0.37 * i); A[i].im = cos(0.11 * i); }
  for (int i = 0; i < K * N; ++i) { B[i].re = cos(0.23 * i); B[i].im = sin(0.05 * i); }
  tdvp_c128 *dA, *dB, *dC;
  cudaMalloc((void**)&dA, sizeof(tdvp_c128) * M * K);
  cudaMalloc((void**)&dB, sizeof(tdvp_c128) * K * N);
  cudaMalloc((void**)&dC, sizeof(tdvp_c128) * M * N);
  cudaMemcpy(dA, A, sizeof(tdvp_c128) * M * K, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B, sizeof(tdvp_c128) * K * N, cudaMemcpyHostToDevice);
  int rc = tdvp_zgemm(h, 0, 0, M, N, K, 1.0, 0.0, dA, K, dB, N, 0.0, 0.0, dC, N);
  if (rc != 0) { fprintf(stderr, "tdvp_zgemm rc=%d: %s\n", rc, tdvp_last_error(h)); return 4; }
  cudaDeviceSynchronize();
  cudaMemcpy(C, dC, sizeof(tdvp_c128) * M * N, cudaMemcpyDeviceToHost);
  double err = 0.0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double re = 0.0, im = 0.0;
      for (int k = 0; k < K; ++k) {
        const tdvp_c128 a = A[i * K + k], b = B[k * N + j];
        re += a.re * b.re - a.im * b.im;
        im += a.re * b.im + a.im * b.re;
      }
      const double d = hypot(C[i * N + j].re - re, C[i * N + j].im - im);
      if (d > err) err = d;
    }
  /* error behaviour: a bad argument returns a negative code and a message, nothing aborts */
  rc = tdvp_zgemm(h, 7, 0, M, N, K, 1.0, 0.0, dA, K, dB, N, 0.0, 0.0, dC, N);
  printf("max_abs_err=%.3e bad_arg_rc=%d msg=\"%s\" launches=%llu\n", err, rc, tdvp_last_error(h), (unsigned long long)tdvp_launch_count());
  tdvp_destroy(h);
  return (err < 1e-12 && rc < 0) ? 0 : 5;
}
