// Roofline denominators measured on the device bench.py runs on: FP64 FMA peak (the driver's
// MEASURED_PEAKS.json has no FP64 figure) and a plain HBM copy.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/pagani_b200.h"
#include "engine.h"

namespace {

// 8 independent FMA chains per thread: enough ILP to saturate the FP64 pipe at any occupancy
__global__ void __launch_bounds__(256)
k_dfma(double* out, int iters, double a, double b)
{
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
         x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = __fma_rn(x0, a, b);
      x1 = __fma_rn(x1, a, b);
      x2 = __fma_rn(x2, a, b);
      x3 = __fma_rn(x3, a, b);
      x4 = __fma_rn(x4, a, b);
      x5 = __fma_rn(x5, a, b);
      x6 = __fma_rn(x6, a, b);
      x7 = __fma_rn(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678)
    out[0] = s;
}

__global__ void __launch_bounds__(256)
k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n)
{
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = in[i];
}

}  // namespace

extern "C" int
pgn_bench_fp64_peak(int device, int iters, double* tflops)
{
  if (!tflops || iters < 1)
    return PGN_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    pgn::set_last_error("no CUDA device");
    return PGN_ERR_NO_DEVICE;
  }
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess)
    return PGN_ERR_CUDA;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess)
    return PGN_ERR_NOMEM;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = sms * 8, block = 256;
  k_dfma<<<grid, block>>>(d, 16, 1.0000001, 1e-9);  // warm-up
  cudaDeviceSynchronize();
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_dfma<<<grid, block>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * 16 * static_cast<double>(iters) * grid * block;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best)
      best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? PGN_OK : PGN_ERR_CUDA;
}

extern "C" int
pgn_bench_hbm_copy(int device, size_t bytes, int iters, double* gbs)
{
  if (!gbs || iters < 1 || bytes < 1024)
    return PGN_ERR_INVALID;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    pgn::set_last_error("no CUDA device");
    return PGN_ERR_NO_DEVICE;
  }
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess)
    return PGN_ERR_CUDA;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t n = bytes / sizeof(double2);
  double2 *a = nullptr, *b = nullptr;
  if (cudaMalloc(&a, n * sizeof(double2)) != cudaSuccess ||
      cudaMalloc(&b, n * sizeof(double2)) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(a);
    return PGN_ERR_NOMEM;
  }
  cudaMemset(a, 0, n * sizeof(double2));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_copy<<<sms * 16, 256>>>(a, b, n);
  cudaDeviceSynchronize();
  double best = 0;
  for (int rep = 0; rep < iters; ++rep) {
    cudaEventRecord(e0);
    k_copy<<<sms * 16, 256>>>(a, b, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9;
    if (g > best)
      best = g;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  *gbs = best;
  return cudaGetLastError() == cudaSuccess ? PGN_OK : PGN_ERR_CUDA;
}
