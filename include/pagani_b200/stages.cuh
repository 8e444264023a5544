// The reference's stage-level classes on top of the B200-native engine: same names, members, ownership and call
// semantics as the reference's headers, implemented through the device-pointer stage entry points of the C ABI
// (pgn_dev_*, include/pagani_b200.h).  Code written against the reference's stage interface -- notably its own unit
// tests test/cuda/pagani/{RegionSplitting,RegionFiltering,RegionSampling,Uniform_sub_division}.cu -- compiles
// unchanged when include/pagani_b200/compat precedes the reference tree on the include path.
//
//   Region_estimates<T,ndim>        cuda/pagani/quad/GPUquad/Region_estimates.cuh:8-38
//   Region_characteristics<ndim>    cuda/pagani/quad/GPUquad/Region_characteristics.cuh:7-27
//   Sub_regions<T,ndim>             cuda/pagani/quad/GPUquad/Sub_regions.cuh:38-258
//   Sub_regions_filter<T,ndim>      cuda/pagani/quad/GPUquad/Sub_region_filter.cuh:48-139
//   Sub_region_splitter<T,ndim>     cuda/pagani/quad/GPUquad/Sub_region_splitter.cuh:48-87
//   Cubature_rules<T,ndim>          cuda/pagani/quad/GPUquad/PaganiUtils.cuh:40-343
//   compute_finished_estimates, accuracy_reached      PaganiUtils.cuh:368-405
//
// Ownership is the reference's: every object owns its device arrays (cudaMalloc / cudaFree), filter() and split()
// free the arrays they replace and re-point the caller's objects, active flags are doubles, split axes ints.
#ifndef PAGANI_B200_STAGES_CUH
#define PAGANI_B200_STAGES_CUH

#include <cuda_runtime.h>

#include <array>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <ostream>
#include <vector>

#include "../pagani_b200.h"
#include "evaluate.cuh"

// numint::integration_result and quad::Volume share the reference's include guards: whichever definition is seen
// first wins and the other header becomes a no-op, so a translation unit may include both trees.
#ifndef GPUINTEGRATION_COMMON_INTEGRATION_RESULT_H
#define GPUINTEGRATION_COMMON_INTEGRATION_RESULT_H
namespace numint {
  // common/integration_result.hh:11-21
  struct integration_result {
    double estimate = 0.;
    double errorest = 0.;
    size_t neval = 0;
    size_t nregions = 0;
    size_t nFinishedRegions = 0;
    int status = -1;
    int lastPhase = -1;
    double chi_sq = 0.;
    size_t iters = 0;
  };
  inline std::ostream&
  operator<<(std::ostream& os, integration_result const& res)
  {
    os << res.estimate << "," << res.errorest << "," << res.nregions << "," << res.chi_sq << ","
       << res.status;
    return os;
  }
}
#endif

#ifndef CUDACUHRE_QUAD_GPUQUAD_VOLUME_CUH
#define CUDACUHRE_QUAD_GPUQUAD_VOLUME_CUH
namespace quad {
  // common/cuda/Volume.cuh:13-52
  template <typename T, int NDIM>
  struct Volume {
    T lows[NDIM] = {0.0};
    T highs[NDIM];
    Volume()
    {
      for (T& x : highs)
        x = 1.0;
    }
    Volume(std::array<T, NDIM> l, std::array<T, NDIM> h)
    {
      std::memcpy(lows, l.data(), NDIM * sizeof(T));
      std::memcpy(highs, h.data(), NDIM * sizeof(T));
    }
    Volume(T const* l, T const* h)
    {
      std::memcpy(lows, l, NDIM * sizeof(T));
      std::memcpy(highs, h, NDIM * sizeof(T));
    }
    Volume(T l, T h)
    {
      for (int d = 0; d < NDIM; ++d) {
        lows[d] = l;
        highs[d] = h;
      }
    }
  };
}
#endif

namespace pagani_b200 {
  // error convention of the reference: CUDA failures abort with a message (common/cuda/cudaDebugUtil.h:118-151),
  // allocation failure throws std::bad_alloc (cudaMemoryUtil.h:198-208)
  inline void
  check(int rc, const char* what)
  {
    if (rc == PGN_OK)
      return;
    if (rc == PGN_ERR_NOMEM)
      throw std::bad_alloc();
    std::fprintf(stderr, "pagani_b200: %s failed (%d): %s\n", what, rc, pgn_last_error());
    std::abort();
  }

  namespace compat {
    template <typename T>
    T*
    device_alloc(size_t n)
    {
      T* p = nullptr;
      if (cudaMalloc(reinterpret_cast<void**>(&p), sizeof(T) * (n > 0 ? n : 1)) != cudaSuccess) {
        cudaGetLastError();
        throw std::bad_alloc();
      }
      return p;
    }
    template <typename T>
    void
    copy(T* dst, const T* src, size_t n, cudaMemcpyKind kind)
    {
      if (n > 0 && cudaMemcpy(dst, src, sizeof(T) * n, kind) != cudaSuccess) {
        std::fprintf(stderr, "pagani_b200: cudaMemcpy failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        std::abort();
      }
    }
    // One engine context per dimension for the stage objects that carry no state of their own (region lists,
    // filter, splitter); created on first use, lives until the process ends.
    template <size_t ndim>
    pgn_ctx*
    stage_context()
    {
      static pgn_ctx* ctx = [] {
        pgn_ctx* c = nullptr;
        check(pgn_create(&c, static_cast<int>(ndim), nullptr), "pgn_create");
        return c;
      }();
      return ctx;
    }
  }
}

template <typename T, size_t ndim>
class Region_estimates {
public:
  Region_estimates() {}
  Region_estimates(size_t num_regions) { reallocate(num_regions); }
  Region_estimates(const Region_estimates&) = delete;
  Region_estimates& operator=(const Region_estimates&) = delete;

  void
  reallocate(size_t num_regions)
  {
    cudaFree(integral_estimates);
    cudaFree(error_estimates);
    integral_estimates = pagani_b200::compat::device_alloc<T>(num_regions);
    error_estimates = pagani_b200::compat::device_alloc<T>(num_regions);
    size = num_regions;
  }

  ~Region_estimates()
  {
    cudaFree(integral_estimates);
    cudaFree(error_estimates);
  }

  T* integral_estimates = nullptr;
  T* error_estimates = nullptr;
  size_t size = 0;
};

template <size_t ndim>
class Region_characteristics {
public:
  Region_characteristics(size_t num_regions)
  {
    active_regions = pagani_b200::compat::device_alloc<double>(num_regions);
    sub_dividing_dim = pagani_b200::compat::device_alloc<int>(num_regions);
    size = num_regions;
  }
  Region_characteristics(const Region_characteristics&) = delete;
  Region_characteristics& operator=(const Region_characteristics&) = delete;

  ~Region_characteristics()
  {
    cudaFree(active_regions);
    cudaFree(sub_dividing_dim);
  }

  size_t size = 0;
  double* active_regions = nullptr;  // 0. / 1., as in the reference
  int* sub_dividing_dim = nullptr;
};

template <typename T, size_t ndim>
struct Sub_regions {
  static_assert(sizeof(T) == sizeof(double), "the engine is FP64 only, like every use in the reference");

  Sub_regions() {}
  Sub_regions(const size_t partitions_per_axis) { uniform_split(partitions_per_axis); }

  // deep copy (the reference's copy constructor copies `size` instead of `size * ndim` elements,
  // Sub_regions.cuh:52-59; a copy of a region list that drops most of it is of no use to anybody)
  Sub_regions(const Sub_regions<T, ndim>& other)
  {
    device_init(other.size);
    pagani_b200::compat::copy(dLeftCoord, other.dLeftCoord, size * ndim, cudaMemcpyDeviceToDevice);
    pagani_b200::compat::copy(dLength, other.dLength, size * ndim, cudaMemcpyDeviceToDevice);
  }
  Sub_regions& operator=(const Sub_regions&) = delete;

  ~Sub_regions()
  {
    delete[] LeftCoord;
    delete[] Length;
    cudaFree(dLeftCoord);
    cudaFree(dLength);
    if (snapshot_dLeftCoord != dLeftCoord)
      cudaFree(snapshot_dLeftCoord);
    if (snapshot_dLength != dLength)
      cudaFree(snapshot_dLength);
  }

  void
  host_device_init(const size_t numRegions)
  {
    delete[] LeftCoord;
    delete[] Length;
    LeftCoord = new T[numRegions * ndim];
    Length = new T[numRegions * ndim];
    device_init(numRegions);
    host_data_size = numRegions;
  }

  void
  refresh_host_device()
  {
    pagani_b200::compat::copy(LeftCoord, dLeftCoord, size * ndim, cudaMemcpyDeviceToHost);
    pagani_b200::compat::copy(Length, dLength, size * ndim, cudaMemcpyDeviceToHost);
  }

  void
  refresh_device_from_host()
  {
    pagani_b200::compat::copy(dLeftCoord, LeftCoord, size * ndim, cudaMemcpyHostToDevice);
    pagani_b200::compat::copy(dLength, Length, size * ndim, cudaMemcpyHostToDevice);
  }

  void
  host_init()
  {
    delete[] LeftCoord;
    delete[] Length;
    host_data_size = size;
    LeftCoord = new T[size * ndim];
    Length = new T[size * ndim];
  }

  void
  device_init(size_t const numRegions)
  {
    cudaFree(dLeftCoord);
    cudaFree(dLength);
    size = numRegions;
    dLeftCoord = pagani_b200::compat::device_alloc<T>(numRegions * ndim);
    dLength = pagani_b200::compat::device_alloc<T>(numRegions * ndim);
  }

  // explicit list from host arrays, SoA [dim * n + i] like dLeftCoord / dLength
  void
  set_host(const T* left, const T* length, size_t n)
  {
    device_init(n);
    pagani_b200::compat::copy(dLeftCoord, left, n * ndim, cudaMemcpyHostToDevice);
    pagani_b200::compat::copy(dLength, length, n * ndim, cudaMemcpyHostToDevice);
  }

  void
  print_bounds()
  {
    host_init();
    refresh_host_device();
    for (size_t i = 0; i < size; i++) {
      for (size_t dim = 0; dim < ndim; dim++)
        std::printf("region %lu, %lu, %f, %f, %f", i, dim, LeftCoord[size * dim + i],
                    LeftCoord[size * dim + i] + Length[size * dim + i], Length[size * dim + i]);
      std::printf("\n");
    }
  }

  T
  compute_region_volume(size_t const regionID)
  {
    T reg_vol = 1.;
    for (size_t dim = 0; dim < ndim; dim++)
      reg_vol *= Length[size * dim + regionID];
    return reg_vol;
  }

  T
  compute_total_volume()
  {
    host_init();
    refresh_host_device();
    T total_vol = 0.;
    for (size_t regID = 0; regID < size; regID++)
      total_vol += compute_region_volume(regID);
    return total_vol;
  }

  // create_uniform_split, Sub_regions.cuh:11-36, 164-185
  void
  uniform_split(size_t numOfDivisionPerRegionPerDimension)
  {
    size_t n = 1;
    for (size_t d = 0; d < ndim; ++d)
      n *= numOfDivisionPerRegionPerDimension;
    device_init(n);
    pagani_b200::check(pgn_dev_uniform_split(pagani_b200::compat::stage_context<ndim>(),
                                             static_cast<int>(numOfDivisionPerRegionPerDimension), dLeftCoord, dLength, n),
                       "pgn_dev_uniform_split");
  }

  quad::Volume<T, static_cast<int>(ndim)>
  extract_region(size_t const regionID)
  {
    if (LeftCoord == nullptr || Length == nullptr || host_data_size != size)
      host_init();
    refresh_host_device();
    quad::Volume<T, static_cast<int>(ndim)> regionID_bounds;
    for (size_t dim = 0; dim < ndim; dim++) {
      size_t region_index = size * dim + regionID;
      regionID_bounds.lows[dim] = LeftCoord[region_index];
      regionID_bounds.highs[dim] = LeftCoord[region_index] + Length[region_index];
    }
    return regionID_bounds;
  }

  void
  set_ptr_to_estimates(Region_estimates<T, ndim>* estimates)
  {
    assert(estimates != nullptr && estimates->size == this->size);
    region_estimates = estimates;
  }

  void
  set_ptr_to_characteristics(Region_characteristics<ndim>* charactrs)
  {
    assert(charactrs != nullptr && charactrs->size == this->size);
    characteristics = charactrs;
  }

  // Sub_regions.cuh:219-239: a device copy of the list that a later run can be restarted from
  void
  take_snapshot()
  {
    if (snapshot_dLeftCoord != dLeftCoord)
      cudaFree(snapshot_dLeftCoord);
    if (snapshot_dLength != dLength)
      cudaFree(snapshot_dLength);
    snapshot_size = size;
    snapshot_dLeftCoord = pagani_b200::compat::device_alloc<T>(size * ndim);
    snapshot_dLength = pagani_b200::compat::device_alloc<T>(size * ndim);
    pagani_b200::compat::copy(snapshot_dLeftCoord, dLeftCoord, size * ndim, cudaMemcpyDeviceToDevice);
    pagani_b200::compat::copy(snapshot_dLength, dLength, size * ndim, cudaMemcpyDeviceToDevice);
  }

  void
  load_snapshot()
  {
    if (snapshot_dLeftCoord == nullptr)
      return;
    cudaFree(dLeftCoord);
    cudaFree(dLength);
    dLeftCoord = snapshot_dLeftCoord;
    dLength = snapshot_dLength;
    size = snapshot_size;
    // the snapshot now IS the list (the reference aliases the pointers the same way); a new snapshot re-copies
    snapshot_dLeftCoord = nullptr;
    snapshot_dLength = nullptr;
  }

  // for accessing on the host side, may need to invoke refresh_host_device() to do the copy
  T* LeftCoord = nullptr;
  T* Length = nullptr;

  // device side
  T* dLeftCoord = nullptr;
  T* dLength = nullptr;

  T* snapshot_dLeftCoord = nullptr;
  T* snapshot_dLength = nullptr;
  Region_characteristics<ndim>* characteristics = nullptr;
  Region_estimates<T, ndim>* region_estimates = nullptr;

  size_t size = 0;
  size_t host_data_size = 0;
  size_t snapshot_size = 0;
};

template <typename T, size_t ndim, bool use_custom = false>
class Sub_regions_filter {
public:
  using Regions = Sub_regions<T, ndim>;
  using Region_char = Region_characteristics<ndim>;
  using Region_ests = Region_estimates<T, ndim>;

  Sub_regions_filter(const size_t num_regions) { (void)num_regions; }

  size_t
  get_num_active_regions(T* active_regions, const size_t num_regions)
  {
    size_t m = 0;
    pagani_b200::check(
      pgn_dev_count_active(pagani_b200::compat::stage_context<ndim>(), active_regions, num_regions, &m),
      "pgn_dev_count_active");
    return m;
  }

  // filter out finished regions: stable compaction of bounds and split axes; the estimates of the survivors become
  // the parents' (Sub_region_filter.cuh:85-139)
  size_t
  filter(Regions& sub_regions, Region_char& region_characteristics, const Region_ests& region_ests,
         Region_ests& parent_ests)
  {
    const size_t n = sub_regions.size;
    const size_t m = get_num_active_regions(region_characteristics.active_regions, n);
    if (m == 0)
      return 0;
    T* filtered_leftCoord = pagani_b200::compat::device_alloc<T>(m * ndim);
    T* filtered_length = pagani_b200::compat::device_alloc<T>(m * ndim);
    int* filtered_sub_dividing_dim = pagani_b200::compat::device_alloc<int>(m);
    parent_ests.reallocate(m);
    pagani_b200::check(
      pgn_dev_filter(pagani_b200::compat::stage_context<ndim>(), sub_regions.dLeftCoord, sub_regions.dLength,
                     region_ests.integral_estimates, region_ests.error_estimates,
                     region_characteristics.sub_dividing_dim, region_characteristics.active_regions, n, m,
                     filtered_leftCoord, filtered_length, parent_ests.integral_estimates,
                     parent_ests.error_estimates, filtered_sub_dividing_dim),
      "pgn_dev_filter");
    cudaFree(sub_regions.dLeftCoord);
    cudaFree(sub_regions.dLength);
    cudaFree(region_characteristics.sub_dividing_dim);
    sub_regions.dLeftCoord = filtered_leftCoord;
    sub_regions.dLength = filtered_length;
    region_characteristics.sub_dividing_dim = filtered_sub_dividing_dim;
    sub_regions.size = m;
    region_characteristics.size = m;
    return m;
  }
};

template <typename T, size_t ndim>
class Sub_region_splitter {
public:
  size_t num_regions;
  Sub_region_splitter(size_t size) : num_regions(size) {}

  // every region is bisected along classifiers.sub_dividing_dim: children at i and num_regions + i
  // (Sub_region_splitter.cuh:55-87)
  void
  split(Sub_regions<T, ndim>& sub_regions, const Region_characteristics<ndim>& classifiers)
  {
    if (num_regions == 0)
      return;
    T* children_left_coord = pagani_b200::compat::device_alloc<T>(num_regions * ndim * 2);
    T* children_length = pagani_b200::compat::device_alloc<T>(num_regions * ndim * 2);
    pagani_b200::check(pgn_dev_split(pagani_b200::compat::stage_context<ndim>(), sub_regions.dLeftCoord,
                                     sub_regions.dLength, classifiers.sub_dividing_dim, num_regions,
                                     children_left_coord, children_length),
                       "pgn_dev_split");
    cudaFree(sub_regions.dLeftCoord);
    cudaFree(sub_regions.dLength);
    sub_regions.size = num_regions * 2;
    sub_regions.dLeftCoord = children_left_coord;
    sub_regions.dLength = children_length;
  }
};

template <typename T, size_t ndim, bool use_custom = false, int debug = 0>
class Cubature_rules {
public:
  static_assert(sizeof(T) == sizeof(double), "the engine is FP64 only, like every use in the reference");
  using Reg_estimates = Region_estimates<T, ndim>;
  using Sub_regs = Sub_regions<T, ndim>;
  using Regs_characteristics = Region_characteristics<ndim>;

  // rule tables are built once per object (Rule.cuh:311-510 + ComputeGenerators, PaganiUtils.cuh:52-96)
  Cubature_rules()
  {
    pagani_b200::check(pgn_create(&ctx_, static_cast<int>(ndim), nullptr), "pgn_create");
  }
  ~Cubature_rules() { pgn_destroy(ctx_); }
  Cubature_rules(const Cubature_rules&) = delete;
  Cubature_rules& operator=(const Cubature_rules&) = delete;

  // host arrays of ndim bounds; no arguments = unit cube (PaganiUtils.cuh:98-112)
  void
  set_device_volume(T const* lows = nullptr, T const* highs = nullptr)
  {
    pagani_b200::check(pgn_set_volume(ctx_, lows, highs), "pgn_set_volume");
  }

  // d_integrand: device-visible functor (quad::make_gpu_integrand / cuda_copy_to_managed).  Fills the estimates and
  // split axes of every region, sets every active flag, returns the two sums (PaganiUtils.cuh:291-343).
  template <typename IntegT>
  numint::integration_result
  apply_cubature_integration_rules(IntegT* d_integrand, int it, const Sub_regs& subregions,
                                   const Reg_estimates& subregion_estimates,
                                   const Regs_characteristics& region_characteristics, bool compute_error = false)
  {
    (void)it;
    cudaDeviceSynchronize();  // the functor copy may still be in flight (managed memory)
    double est = 0., err = 0.;
    pagani_b200::check(
      pgn_dev_evaluate(ctx_, &pagani_b200::detail::trampoline<IntegT, static_cast<int>(ndim)>, d_integrand,
                       subregions.dLeftCoord, subregions.dLength, subregions.size,
                       subregion_estimates.integral_estimates, subregion_estimates.error_estimates,
                       region_characteristics.sub_dividing_dim, region_characteristics.active_regions, &est, &err),
      "pgn_dev_evaluate");
    numint::integration_result res;
    res.estimate = est;
    res.errorest = compute_error ? err : std::numeric_limits<T>::infinity();
    return res;
  }

  pgn_ctx* context() { return ctx_; }

private:
  pgn_ctx* ctx_ = nullptr;
};

// PaganiUtils.cuh:368-385: what the regions that just became inactive contribute
template <typename T, size_t ndim, bool use_custom = false>
numint::integration_result
compute_finished_estimates(const Region_estimates<T, ndim>& estimates, const Region_characteristics<ndim>& classifiers,
                           const numint::integration_result& iter)
{
  const size_t n = estimates.size;
  std::vector<double> est(n), err(n), act(n);
  pagani_b200::compat::copy(est.data(), estimates.integral_estimates, n, cudaMemcpyDeviceToHost);
  pagani_b200::compat::copy(err.data(), estimates.error_estimates, n, cudaMemcpyDeviceToHost);
  pagani_b200::compat::copy(act.data(), classifiers.active_regions, n, cudaMemcpyDeviceToHost);
  double dot_est = 0., dot_err = 0.;
  for (size_t i = 0; i < n; ++i) {
    dot_est += act[i] * est[i];
    dot_err += act[i] * err[i];
  }
  numint::integration_result finished;
  finished.estimate = iter.estimate - dot_est;
  finished.errorest = iter.errorest - dot_err;
  return finished;
}

// PaganiUtils.cuh:387-405
template <typename T>
bool
accuracy_reached(T epsrel, T epsabs, T estimate, T errorest)
{
  return errorest / estimate <= epsrel || errorest <= epsabs;
}

template <typename T>
bool
accuracy_reached(T epsrel, T epsabs, numint::integration_result res)
{
  return res.errorest / res.estimate <= epsrel || res.errorest <= epsabs;
}

#endif
