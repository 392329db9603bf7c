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
    o.ox = g->ox; o.oy = g->oy; o.res = g->res; o.inv_res = 1.0 / g->res;
    o.size_x = g->size_x; o.size_y = g->size_y;
    o.win_x0 = g->win_x0; o.win_y0 = g->win_y0; o.win_w = g->win_w; o.win_h = g->win_h;
    return o;
}

int validate_geom(const occgrid_geom* g);

// A decoded, pose-corrected packet (struct occgrid_pose_rec in the public header): what the
// TILED strategy bins, and what travels between GPUs after routing.
struct __align__(16) PoseRec {       // 48 bytes
    double rx, ry;                   // corrected pose (dual_bot_mapper.py:851-857)
    float yaw;
    float d[4];                      // front, left, back, right (:882-885)
    unsigned int k;                  // index of the source record (informational)
    unsigned int pad[2];
};
static_assert(sizeof(PoseRec) == 48, "PoseRec must be 48 bytes");

// Optional per-kernel timing (occgrid_profile_begin/_end): when enabled, every launch of one
// of our kernels is bracketed by cudaEventRecord on its stream, so bench.py can report the
// dominant kernel's own duration and the exact number of launches.
enum KernelId { K_INTEGRATE_GLOBAL = 0, K_RESOLVE, K_UPDATE_RAYS, K_TILE_COUNT, K_TILE_SCAN, K_TILE_SCATTER,
                K_TILE_RAYCAST, K_TILE_RESOLVE, K_MERGE_EXTRACT, K_MERGE_BOUNDS, K_MERGE_VOXEL, K_MERGE_RASTER,
                K_MERGE_FUSE, K_PROBE, K_ROUTE, K_FRONTIER, K_FRONTIER_CLUSTER, K_CHAIN_PROBE, K_CHAIN_INCR,
                K_CHAIN_REBUILD, K_ICP, K_N_KERNELS };
bool profile_enabled();
void profile_mark(int kernel_id, cudaStream_t st, bool begin, int n_kernels);

struct ProfileScope {
    int id; cudaStream_t st; bool on; int nk;
    ProfileScope(int id_, cudaStream_t st_, int n_kernels = 1) : id(id_), st(st_), on(profile_enabled()), nk(n_kernels) {
        if (on) profile_mark(id, st, true, nk);
    }
    ~ProfileScope() { if (on) profile_mark(id, st, false, nk); }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Block-wide accumulation of the first N uint64 counter slots: per-thread values -> warp
// shuffle -> per-warp partials in shared memory -> one global atomic per slot per CTA.
// All threads must call.  smem_acc needs N * 32 entries.
template <int N>
__device__ __forceinline__ void block_add_counters(const unsigned long long (&v)[N],
                                                   unsigned long long* smem_acc /* >= N * 32 */, uint64_t* d_counters) {
    if (d_counters == nullptr) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        unsigned long long x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) smem_acc[i * 32 + warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        unsigned long long t = 0;
        for (int w = 0; w < nwarps; ++w) t += smem_acc[threadIdx.x * 32 + w];
        if (t) atomicAdd(reinterpret_cast<unsigned long long*>(d_counters) + threadIdx.x, t);
    }
}

}  // namespace occ
