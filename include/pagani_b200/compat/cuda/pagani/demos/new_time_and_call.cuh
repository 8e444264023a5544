// Include-swap shim for cuda/pagani/demos/new_time_and_call.cuh (clean_time_and_call, the CSV driver of the
// reference's demos) on the B200-native engine: include/pagani_b200/demos/time_and_call.cuh.
#ifndef PAGANI_B200_COMPAT_NEW_TIME_AND_CALL_CUH
#define PAGANI_B200_COMPAT_NEW_TIME_AND_CALL_CUH
#include "pagani_b200/Workspace.cuh"
// the reference's header of this name drags in its CUDA utility helpers (quad::cuda_malloc, host_alloc,
// make_gpu_integrand, copy_to_host ...), which callers use without including them: keep that transitive include
#if defined(__has_include)
#if __has_include("common/cuda/cudaMemoryUtil.h")
#include "common/cuda/cudaMemoryUtil.h"
#endif
#endif
#include "pagani_b200/demos/time_and_call.cuh"
#endif
