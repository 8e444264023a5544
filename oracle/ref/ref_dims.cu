// TEST INFRASTRUCTURE (oracle side).  One translation unit per dimension (-DREF_NDIM=n) so the
// reference templates compile in parallel.  The Genz functors at 5..8 dimensions are the
// reference's own (common/cuda/integrands.cuh:1303-1932); the 10-D C0/discontinuous ones and the
// small 2-D test integrands are defined by this project (include/pagani_b200/genz.cuh, below) and
// are compiled here into the REFERENCE's kernels, so both sides evaluate the same source.
#include "ref_pagani.cuh"
#include "common/cuda/integrands.cuh"
#include "pagani_b200/genz.cuh"

#ifndef REF_NDIM
#error "compile with -DREF_NDIM=<n>"
#endif

REF_EXPOSE_WORKSPACE(REF_NDIM)

#if REF_NDIM == 2
// test/cuda/pagani/Easy_Integrals.cu:14-21 and do_integration_from_c.cu:11-17 use these two shapes
struct RefConst2 {
  __device__ __host__ double operator()(double, double) { return 15.37; }
};
struct RefXY2 {
  __device__ __host__ double operator()(double x, double y) { return x * y; }
};
REF_REGISTER(const_2, RefConst2, 2)
REF_REGISTER(xy_2, RefXY2, 2)
#elif REF_NDIM == 3
// products of the coordinates at 3 and 4 axes: multilinear, so the fourth differences are pure rounding noise and the
// split axes pin the reference's machine code for these dimensions (its contraction choices depend on ndim)
struct RefXY3 {
  __device__ __host__ double operator()(double x, double y, double z) { return x * y * z; }
};
REF_REGISTER(xy_3, RefXY3, 3)
#elif REF_NDIM == 4
struct RefXY4 {
  __device__ __host__ double operator()(double x, double y, double z, double w) { return x * y * z * w; }
};
REF_REGISTER(xy_4, RefXY4, 4)
#elif REF_NDIM == 5
REF_REGISTER(f1_5, F_1_5D, 5)
REF_REGISTER(f2_5, F_2_5D, 5)
REF_REGISTER(f3_5, F_3_5D, 5)
REF_REGISTER(f4_5, F_4_5D, 5)
REF_REGISTER(f5_5, F_5_5D, 5)
REF_REGISTER(f6_5, F_6_5D, 5)
#elif REF_NDIM == 6
REF_REGISTER(f1_6, F_1_6D, 6)
REF_REGISTER(f2_6, F_2_6D, 6)
REF_REGISTER(f3_6, F_3_6D, 6)
REF_REGISTER(f4_6, F_4_6D, 6)
REF_REGISTER(f5_6, F_5_6D, 6)
REF_REGISTER(f6_6, F_6_6D, 6)
#elif REF_NDIM == 7
REF_REGISTER(f1_7, F_1_7D, 7)
REF_REGISTER(f2_7, F_2_7D, 7)
REF_REGISTER(f3_7, F_3_7D, 7)
REF_REGISTER(f4_7, F_4_7D, 7)
REF_REGISTER(f5_7, F_5_7D, 7)
REF_REGISTER(f6_7, F_6_7D, 7)
#elif REF_NDIM == 8
REF_REGISTER(f1_8, F_1_8D, 8)
REF_REGISTER(f2_8, F_2_8D, 8)
REF_REGISTER(f3_8, F_3_8D, 8)
REF_REGISTER(f4_8, F_4_8D, 8)
REF_REGISTER(f5_8, F_5_8D, 8)
REF_REGISTER(f6_8, F_6_8D, 8)
#elif REF_NDIM == 10
REF_REGISTER(f5_10, pagani_b200::genz::CZero<10>, 10)
REF_REGISTER(f6_10, pagani_b200::genz::Discontinuous<10>, 10)
#else
#error "no integrands registered for this REF_NDIM"
#endif
