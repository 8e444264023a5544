// Include-swap shim: put include/pagani_b200/compat (and include/) BEFORE the reference tree on the include path and
// "cuda/pagani/quad/GPUquad/Sub_region_filter.cuh" resolves here -- the reference's classes of that header, implemented on the
// B200-native engine (include/pagani_b200/stages.cuh, Workspace.cuh).  Everything else of the reference tree (its
// utility headers, integrands, Catch2) keeps resolving to the reference's own files.
#ifndef PAGANI_B200_COMPAT_SUB_REGION_FILTER_CUH
#define PAGANI_B200_COMPAT_SUB_REGION_FILTER_CUH
#include "pagani_b200/Workspace.cuh"
// the reference's header of this name drags in its CUDA utility helpers (quad::cuda_malloc, host_alloc,
// make_gpu_integrand, copy_to_host ...), which callers use without including them: keep that transitive include
#if defined(__has_include)
#if __has_include("common/cuda/cudaMemoryUtil.h")
#include "common/cuda/cudaMemoryUtil.h"
#endif
#endif
#endif
