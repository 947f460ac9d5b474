// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -DUSE_FLOAT -I include -I approximatenn_b200/csrc -c scratch/next_round/compile_check.cu -o /tmp/s5draft.o
#include "annb_common.cuh"
#include <cuda_fp16.h>
#include "supercharge_screen_draft.cuh"
template __global__ void supercharge_screen_kernel<8>(const float *, const float *, const unsigned short *, const float2 *,
    const unsigned *, const u32 *, const float *, const u32 *, size_t, int, size_t, size_t, int, u32 *, float *, TieList);
