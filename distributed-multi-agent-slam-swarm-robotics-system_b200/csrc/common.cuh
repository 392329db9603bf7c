// Shared host-side helpers of the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/occgrid_b200.h"
#include "beam_expand.cuh"

namespace occ {

void set_last_error(const char* fmt, ...);
bool cuda_ok(cudaError_t e, const char* what);

#define OCC_CUDA_TRY(call)                                  \
    do {                                                    \
        if (!occ::cuda_ok((call), #call)) return OCCGRID_E_CUDA; \
    } while (0)

inline Geom to_geom(const occgrid_geom* g) {
    Geom o;
    o.ox = g->ox; o.oy = g->oy; o.res = g->res;
    o.size_x = g->size_x; o.size_y = g->size_y;
    o.win_x0 = g->win_x0; o.win_y0 = g->win_y0; o.win_w = g->win_w; o.win_h = g->win_h;
    return o;
}

int validate_geom(const occgrid_geom* g);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Block-wide accumulation of the first N uint64 counter slots: per-thread values -> warp
// shuffle -> shared memory -> one global atomic per slot per CTA.  All threads must call.
template <int N>
__device__ __forceinline__ void block_add_counters(const unsigned long long (&v)[N],
                                                   unsigned long long* smem_acc /* >= N */, uint64_t* d_counters) {
    if (d_counters == nullptr) return;
    if (threadIdx.x < N) smem_acc[threadIdx.x] = 0ull;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        unsigned long long x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&smem_acc[i], x);
    }
    __syncthreads();
    if (threadIdx.x < N && smem_acc[threadIdx.x])
        atomicAdd(reinterpret_cast<unsigned long long*>(d_counters) + threadIdx.x, smem_acc[threadIdx.x]);
}

}  // namespace occ
