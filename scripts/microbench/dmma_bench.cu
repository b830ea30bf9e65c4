// FP64 pipe micro-benchmark for sm_100a: which double-precision instruction
// shape feeds the B200 FP64 units fastest?  Measures sustained FMA/clk/SM for
//   - DFMA (plain fma.rn.f64)
//   - mma.sync.m8n8k4.f64   (Ampere+ DMMA)
//   - mma.sync.m16n8k4 / m16n8k8 / m16n8k16.f64 (sm_90+ DMMA shapes)
// Output is one JSON line per variant; used by DESIGN.md to pick the ZGEMM
// inner instruction and to state the FP64 roof the roofline is quoted against.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_bench dmma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

template <int NACC>
__global__ void k_dfma(double* out, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_m8n8k4(double* out, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_m16n8k4(double* out, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_m16n8k8(double* out, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_m16n8k16(double* out, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                   "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b),
                     "d"(a), "d"(b), "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F launch, double fma_per_thread_iter_times32, int warps, int nsm, double clk_ghz) {
  // fma_per_thread_iter_times32: FMAs issued per WARP per outer iteration
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch();  // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  double fmas = fma_per_thread_iter_times32 * (double)ITERS * warps * nsm;
  double tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
  double fma_clk_sm = fmas / nsm / (best * 1e-3 * clk_ghz * 1e9);
  printf("{\"variant\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"fma_per_clk_per_sm_at_%.3fGHz\": %.1f}\n",
         name, warps, best, tflops, clk_ghz, fma_clk_sm);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount;
  double clk = p.clockRate * 1e-6;  // GHz (max)
  printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_ghz\": %.3f}\n", p.name, nsm, clk);
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 1024));
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32;
    constexpr int NACC = 8;
    run("dfma", [&] { k_dfma<NACC><<<nsm, threads>>>(out, 1.0000001, 1e-9); }, 32.0 * NACC, warps, nsm, clk);
    run("dmma_m8n8k4", [&] { k_m8n8k4<NACC><<<nsm, threads>>>(out, 1.0000001, 1e-9); }, 256.0 * NACC, warps, nsm, clk);
    run("dmma_m16n8k4", [&] { k_m16n8k4<NACC><<<nsm, threads>>>(out, 1.0000001, 1e-9); }, 512.0 * NACC, warps, nsm, clk);
    run("dmma_m16n8k8", [&] { k_m16n8k8<NACC><<<nsm, threads>>>(out, 1.0000001, 1e-9); }, 1024.0 * NACC, warps, nsm, clk);
    run("dmma_m16n8k16", [&] { k_m16n8k16<NACC><<<nsm, threads>>>(out, 1.0000001, 1e-9); }, 2048.0 * NACC, warps, nsm, clk);
  }
  CK(cudaFree(out));
  return 0;
}
