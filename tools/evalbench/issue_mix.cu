// Issue-slot model of an sm_100a scheduler for the evaluation kernel's instruction mix: how many cycles does
// one SMSP need for a DFMA plus K instructions of another class, with enough warps and independent chains that
// no dependent-issue latency is exposed?  (tools/evalbench/fp64_lat.cu showed DFMA = 2 cycles / warp instruction
// and LOP3/IMAD adding ~1 cycle each; this separates the classes the point loop actually uses.)
//   ./issue_mix            -> one JSON line per (class, K); check the SASS mix with
//   cuobjdump -sass issue_mix | grep -c ...   (every "other" instruction is an asm volatile)
#include <cstdio>
#include <cuda_runtime.h>

enum Kind { kNone, kLop3, kPrmt, kIadd, kFsel, kLds64, kFfma, kCvt, kShfl, kLdc, kImad, kIsetpSel, kDaddMagic, kLds128, kMov, kKinds };
static const char* kNames[] = {"none", "lop3", "prmt", "iadd3", "fsel", "lds64", "ffma", "f2i+i2f(f64)", "shfl", "ldc", "imad",
                               "isetp+sel", "dadd_magic_pair", "lds128", "mov"};

__constant__ double c_tab[64];

template <int KIND>
__device__ __forceinline__ void
other(unsigned& a, unsigned a2, unsigned& b, float& f, double& d, unsigned saddr)
{
  if constexpr (KIND == kLop3)
    asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(a) : "r"(b), "r"(a2));
  else if constexpr (KIND == kPrmt)
    asm volatile("prmt.b32 %0, %0, %1, 0x7650;" : "+r"(a) : "r"(b));
  else if constexpr (KIND == kIadd)
    asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(a2));
  else if constexpr (KIND == kFsel)
    asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.f32 %0, %0, 1.5, p;}" : "+f"(f) : "r"(b));
  else if constexpr (KIND == kLds64) {
    double t;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(t) : "r"(saddr));
    d = t;
  } else if constexpr (KIND == kLds128) {
    double t, u;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(t), "=d"(u) : "r"(saddr));
    d = t;
  } else if constexpr (KIND == kFfma)
    asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f));
  else if constexpr (KIND == kCvt) {
    int q;
    asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(q) : "d"(d));
    asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d) : "r"(q));
  } else if constexpr (KIND == kShfl)
    asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(a));
  else if constexpr (KIND == kLdc) {
    double t;
    asm volatile("ld.const.f64 %0, [%1];" : "=d"(t) : "l"(reinterpret_cast<unsigned long long>(c_tab) + (b & 0x38)));
    d = t;
  } else if constexpr (KIND == kImad)
    asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(a) : "r"(b));
  else if constexpr (KIND == kIsetpSel)
    asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %0, %1, p;}" : "+r"(a) : "r"(b));
  else if constexpr (KIND == kDaddMagic) {
    asm volatile("add.rn.f64 %0, %0, 0d4338000000000000;" : "+d"(d));
    asm volatile("sub.rn.f64 %0, %0, 0d4338000000000000;" : "+d"(d));
  } else if constexpr (KIND == kMov)
    asm volatile("mov.b32 %0, %1;" : "=r"(a) : "r"(b));
}

// CHAINS independent DFMA chains; after every DFMA, K "other" instructions on rotating state
template <int KIND, int K, int NDFMA>
__global__ void __launch_bounds__(1024)
mix(double* out, long long* cycles, int iters, double x, double y)
{
  __shared__ double s_tab[64];
  if (threadIdx.x < 64)
    s_tab[threadIdx.x] = threadIdx.x;
  double v[8];
  unsigned a[4], b[4];
  float f[4];
  double d[4];
#pragma unroll
  for (int c = 0; c < 8; ++c)
    v[c] = threadIdx.x + c;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    a[c] = threadIdx.x * 7 + c;
    b[c] = threadIdx.x * 3 + c + 1;
    f[c] = threadIdx.x + c;
    d[c] = threadIdx.x * 0.25 + c;
  }
  const unsigned saddr = static_cast<unsigned>(__cvta_generic_to_shared(s_tab)) + (threadIdx.x & 3) * 16;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if constexpr (NDFMA > 0)
        v[u] = __fma_rn(v[u], x, y);
#pragma unroll
      for (int k = 0; k < K; ++k)
        other<KIND>(a[(u + k) & 3], a[(u + k + 2) & 3], b[(u + k + 1) & 3], f[(u + k) & 3], d[(u + k) & 3], saddr + ((u + k) & 3) * 64);
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    s += v[c];
#pragma unroll
  for (int c = 0; c < 4; ++c)
    s += a[c] + b[c] + f[c] + d[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    *cycles = t1 - t0;
}

template <int KIND, int K, int NDFMA>
void
run(double* out, long long* cyc)
{
  const int iters = 500, warps = 32;
  mix<KIND, K, NDFMA><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaError_t e = cudaDeviceSynchronize();
  const double groups_per_smsp = 8.0 * iters * (warps / 4.0);  // (DFMA + K others) groups issued per scheduler
  printf("{\"other\": \"%s\", \"k_per_group\": %d, \"dfma_per_group\": %d, \"smsp_cycles_per_group\": %.3f%s}\n", kNames[KIND], K,
         NDFMA, *cyc / groups_per_smsp, e == cudaSuccess ? "" : ", \"error\": true");
}

template <int KIND>
void
sweep(double* out, long long* cyc)
{
  run<KIND, 1, 1>(out, cyc);
  run<KIND, 2, 1>(out, cyc);
  run<KIND, 4, 1>(out, cyc);
  run<KIND, 2, 0>(out, cyc);  // the other class alone
  run<KIND, 4, 0>(out, cyc);
}

int
main()
{
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1024 * sizeof(double));
  cudaMallocManaged(&cyc, sizeof(long long));
  double h[64];
  for (int i = 0; i < 64; ++i)
    h[i] = i;
  cudaMemcpyToSymbol(c_tab, h, sizeof h);
  run<kNone, 0, 1>(out, cyc);
  sweep<kLop3>(out, cyc);
  sweep<kPrmt>(out, cyc);
  sweep<kIadd>(out, cyc);
  sweep<kFsel>(out, cyc);
  sweep<kLds64>(out, cyc);
  sweep<kLds128>(out, cyc);
  sweep<kFfma>(out, cyc);
  sweep<kCvt>(out, cyc);
  sweep<kShfl>(out, cyc);
  sweep<kImad>(out, cyc);
  sweep<kDaddMagic>(out, cyc);
  return 0;
}
