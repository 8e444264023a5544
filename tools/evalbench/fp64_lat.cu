// FP64 pipe micro-benchmarks for sm_100a: dependent-issue latency of DFMA and the per-SMSP throughput as a
// function of independent chains per warp and warps per scheduler (what the evaluation kernel's point loop
// can hope for).   ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void
chains(double* out, long long* cycles, int iters, double a, double b)
{
  double v[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
    v[c] = threadIdx.x + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c)
        v[c] = __fma_rn(v[c], a, b);
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
    s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    *cycles = t1 - t0;
}

// 8 independent DFMA chains + INT_PER_DFMA independent integer ops per DFMA: does an integer instruction
// issue in the shadow of a DFMA's second pipe cycle, or does the DFMA hold the issue port for both?
template <int INT_PER_DFMA>
__global__ void
mixed(double* out, long long* cycles, int iters, double a, double b, unsigned m)
{
  double v[8];
  unsigned w[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    v[c] = threadIdx.x + c;
    w[c] = threadIdx.x * 7 + c;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        v[c] = __fma_rn(v[c], a, b);
#pragma unroll
        for (int k = 0; k < INT_PER_DFMA; ++k)
          w[(c + k) & 7] = (w[(c + k) & 7] ^ m) + (w[(c + k + 1) & 7] & 0x55u);  // LOP3 + IADD3 (2 ops)
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
  unsigned t = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    s += v[c];
    t += w[c];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    *cycles = t1 - t0;
}

template <int INT_PER_DFMA>
void
run_mixed(double* out, long long* cyc)
{
  const int iters = 1000, warps = 32;
  mixed<INT_PER_DFMA><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9, 0x1234567u);
  cudaDeviceSynchronize();
  const double dfma_per_smsp = 32.0 * iters * (warps / 4.0);
  printf("{\"int_pairs_per_dfma\": %d, \"warps\": %d, \"smsp_cycles_per_dfma\": %.3f}\n", INT_PER_DFMA, warps,
         *cyc / dfma_per_smsp);
}

template <int CHAINS>
void
run(int warps_per_block, double* out, long long* cyc)
{
  const int iters = 2000;
  chains<CHAINS><<<1, 32 * warps_per_block>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  const double per_warp_instr = 8.0 * CHAINS * iters;
  // one block on one SM: warps spread over the 4 schedulers
  const double warps_per_smsp = warps_per_block / 4.0 < 1 ? 1 : warps_per_block / 4.0;
  printf("{\"chains\": %d, \"warps\": %d, \"cycles_per_dfma_per_warp\": %.2f, \"smsp_dfma_per_cycle\": %.3f}\n", CHAINS,
         warps_per_block, *cyc / per_warp_instr, per_warp_instr * warps_per_smsp / *cyc);
}

int
main()
{
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMallocManaged(&cyc, sizeof(long long));
  for (int w : {1, 4, 8, 16, 32}) {
    run<1>(w, out, cyc);
    run<2>(w, out, cyc);
    run<4>(w, out, cyc);
    run<8>(w, out, cyc);
  }
  run_mixed<0>(out, cyc);
  run_mixed<1>(out, cyc);
  run_mixed<2>(out, cyc);
  return 0;
}
