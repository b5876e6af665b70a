// fp64_issue_bench.cu — micro-benchmark behind DESIGN.md's issue model for the resident kernels (sm_100a):
// FP64 pipe rate per SM, cost of integer instructions beside it, dependent-chain latency at 4 warps per scheduler.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue_bench fp64_issue_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NCHAIN, int NINT>
__global__ void __launch_bounds__(256, 2) k(double *out, int *iout, int iters, double b, double c, int ib) {
  double a[NCHAIN];
  int q[NINT > 0 ? NINT : 1];
#pragma unroll
  for (int j = 0; j < NCHAIN; ++j) a[j] = 1.0 + threadIdx.x * 1e-9 + j * 1e-3;
#pragma unroll
  for (int j = 0; j < (NINT > 0 ? NINT : 1); ++j) q[j] = threadIdx.x + j;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < NCHAIN; ++j) a[j] = fma(a[j], b, c);
#pragma unroll
      for (int j = 0; j < NINT; ++j) q[j] = q[j] * ib + (q[j] >> 3);    // IMAD + SHF-class work
    }
  }
  double s = 0.0; int t = 0;
#pragma unroll
  for (int j = 0; j < NCHAIN; ++j) s += a[j];
#pragma unroll
  for (int j = 0; j < (NINT > 0 ? NINT : 1); ++j) t += q[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int NCHAIN, int NINT>
void run(const char *name, int sms, double *out, int *iout) {
  const int iters = 4000, blocks = sms * 2, threads = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<NCHAIN, NINT><<<blocks, threads>>>(out, iout, iters, 0.999999, 1e-7, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double cycles = best * 1e-3 * khz * 1e3;                       // at the nominal max clock
  const double dfma_warp = (double)iters * 8 * NCHAIN * 16;            // warp-level DFMAs per SM (16 warps)
  const double int_warp = (double)iters * 8 * NINT * 2 * 16;           // ~2 integer instr per q update
  printf("%-26s %8.3f ms  DFMA/clk/SM %.3f (warp-instr)  int/clk/SM %.3f  cycles per DFMA per warp %.2f\n", name, best,
         dfma_warp / cycles, int_warp / cycles, cycles / ((double)iters * 8 * NCHAIN));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out; int *iout;
  cudaMalloc(&out, sms * 2 * 256 * sizeof(double)); cudaMalloc(&iout, sms * 2 * 256 * sizeof(int));
  printf("16 warps per SM (2 CTAs x 256 threads), like the resident kernels\n");
  run<1, 0>("1 dependent chain", sms, out, iout);
  run<2, 0>("2 chains", sms, out, iout);
  run<4, 0>("4 chains", sms, out, iout);
  run<8, 0>("8 chains", sms, out, iout);
  run<1, 1>("1 chain + 1 int update", sms, out, iout);
  run<1, 2>("1 chain + 2 int updates", sms, out, iout);
  run<4, 2>("4 chains + 2 int updates", sms, out, iout);
  run<4, 4>("4 chains + 4 int updates", sms, out, iout);
  run<4, 8>("4 chains + 8 int updates", sms, out, iout);
  run<8, 8>("8 chains + 8 int updates", sms, out, iout);
  return 0;
}
