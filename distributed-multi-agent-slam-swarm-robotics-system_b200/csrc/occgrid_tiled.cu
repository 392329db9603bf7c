// Strategy TILED of occgrid_integrate_packets: packets binned by the HOME TILE of the robot
// cell, last-writer-wins resolved with SHARED-MEMORY atomics.
//
// Same reference semantics as occgrid_integrate.cu (server_nodes/dual_bot_mapper.py:826-903,
// :136-179; every cell keeps the largest  (beam ordinal + 1) << 1 | occupied  stamp), but the
// per-cell atomicMax runs on a stamp window in shared memory, where B200 sustains
// ~1.0-1.5e12 atomics/s (measured, occgrid_scatter_probe kind 2) against ~2.2e11/s for
// red.global (kind 0).  Only the surviving stamp of each touched cell goes to L2.
//
// All four beams of a packet start at the robot cell (:142) and are at most
// R = ceil(MAX_DIST_M / res) + 2 cells long (:57, :900), so they lie inside the robot's 64x64
// home tile grown by R on every side.  Binning therefore needs only the robot cell — no
// trigonometry, one record per packet, no duplicates:
//
//   k_home_count    thread/packet: decode (:828-843), pose correction (:851-857), robot cell
//                   (:142) -> 48-byte pose record (packet order) + packets per home tile
//   k_tile_plan     one CTA over the active tiles (listed by the count pass): exclusive scan of
//                   their counts -> bin offsets, (tile, chunk) work items of <= kChunkPk packets
//   k_home_scatter  thread/packet: write the packet's index into its tile's bin
//   k_home_raycast  persistent CTAs pull work items: zero the (64+2R)^2 smem window, expand
//                   each record into its 4 beams (fp64 endpoints, :887-902) and walk the exact
//                   reference Bresenham (:158-179) with atomicMax in shared memory; then flush
//                   the non-zero stamps with coalesced red.global.max to the stamp plane
//   k_home_resolve  persistent CTAs over the active tiles' windows: stamps -> int8 grid, := 0
//
// Windows of neighbouring tiles overlap; the global atomicMax merges them, and chunks of one
// tile likewise, so the result is independent of how work was cut.
#include "common.cuh"

namespace occ {

constexpr int kTile = 64;            // home-tile side in cells
constexpr int kTileShift = 6;
constexpr int kTT = 256;             // threads per CTA
constexpr int kChunkPk = 2048;       // packets per work item
constexpr int kMaxStrideT = 64;
constexpr int kMaxWindowBytes = 100 * 1024;   // smem window budget (two CTAs per SM at least)

struct TilePlanHeader {
    unsigned int n_items, n_active, total_records, work_counter, resolve_counter, overflow;
    unsigned int active_count;   // tiles appended to the active list by the count pass of the current call
    unsigned int pad;
};

struct TileGeom {
    int reach;            // R
    int pad;              // cells added around the window so that every home tile is >= 0
    int tiles_x, tiles_y, n_tiles;
    int win_side;         // smem window side = kTile + 2R
    int pitch;            // smem row pitch (odd)
};

static inline int reach_cells(double res) { return (int)ceil(OCC_MAX_DIST_M / res) + 2; }

static inline TileGeom tile_geom(const occgrid_geom* g) {
    TileGeom t;
    t.reach = reach_cells(g->res);
    t.pad = ((t.reach + kTile - 1) / kTile) * kTile;
    t.tiles_x = (g->win_w + 2 * t.pad + kTile - 1) >> kTileShift;
    t.tiles_y = (g->win_h + 2 * t.pad + kTile - 1) >> kTileShift;
    t.n_tiles = t.tiles_x * t.tiles_y;
    t.win_side = kTile + 2 * t.reach;
    t.pitch = t.win_side | 1;
    return t;
}

__device__ __forceinline__ void stage_records_t(const uint8_t* __restrict__ src, size_t bytes, uint8_t* smem) {
    const size_t nvec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? bytes / 16 : 0;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(smem);
    for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) d4[i] = __ldg(s4 + i);
    for (size_t i = nvec * 16 + threadIdx.x; i < bytes; i += blockDim.x) smem[i] = __ldg(src + i);
}

// Home tile of a corrected pose, or -1 when no beam of the packet can reach the window.
__device__ __forceinline__ int home_tile(const Geom& g, const TileGeom& tg, double rx, double ry) {
    const double qx = cell_quotient(rx, g.ox, g.res), qy = cell_quotient(ry, g.oy, g.res);
    if (!quotient_in_range(qx) || !quotient_in_range(qy)) return -1;
    const int px = trunc_cell(qx) - g.win_x0 + tg.pad, py = trunc_cell(qy) - g.win_y0 + tg.pad;
    if (px < tg.pad - tg.reach || py < tg.pad - tg.reach || px >= tg.pad + g.win_w + tg.reach ||
        py >= tg.pad + g.win_h + tg.reach)
        return -1;
    return (py >> kTileShift) * tg.tiles_x + (px >> kTileShift);
}

// ---- binning passes ------------------------------------------------------------------------
// The hot tiles of a swarm are few (two robots per room), so per-packet global atomics on the
// tile counters serialise on a handful of L2 addresses.  Each CTA therefore bins kPkPerCta
// packets in a shared-memory hash table first and touches every distinct tile once.
constexpr int kPkPerCta = 2048;
constexpr int kSub = kPkPerCta / kTT;           // packets per thread
constexpr int kHash = 4096;                     // >= 2 * kPkPerCta would be ideal; distinct tiles <= kPkPerCta
constexpr unsigned int kEmpty = 0xffffffffu;
static_assert(kHash >= 2 * kPkPerCta, "hash table must stay at most half full");

__device__ __forceinline__ int hash_insert(unsigned int* keys, unsigned int tile) {
    unsigned int h = (tile * 2654435761u) >> 20;             // 12 bits
    for (;;) {
        const unsigned int old = atomicCAS(&keys[h], kEmpty, tile);
        if (old == kEmpty || old == tile) return (int)h;
        h = (h + 1) & (kHash - 1);
    }
}

__global__ void __launch_bounds__(kTT)
k_home_count(Geom g, TileGeom tg, const uint8_t* __restrict__ pkts, long long n, int stride,
             const int32_t* __restrict__ agent_idx, const double* __restrict__ drift,
             const double* __restrict__ agent_off, int n_agents,
             unsigned int* __restrict__ tile_count, int* __restrict__ tile_ids, PoseRec* __restrict__ recs,
             unsigned int* __restrict__ active, TilePlanHeader* __restrict__ hdr, uint64_t* counters) {
    __shared__ __align__(16) uint8_t s_rec[kTT * kMaxStrideT];
    __shared__ unsigned int s_keys[kHash];
    __shared__ unsigned int s_vals[kHash];
    for (int i = threadIdx.x; i < kHash; i += kTT) { s_keys[i] = kEmpty; s_vals[i] = 0u; }
    unsigned long long c[OCCGRID_C_HITS + 1] = {};
    const long long cta_first = (long long)blockIdx.x * kPkPerCta;
    // Software-pipelined staging: the 16-byte loads of sub-batch s+1 are in flight while
    // sub-batch s is decoded out of shared memory.
    constexpr int kVec = (kTT * kMaxStrideT / 16 + kTT - 1) / kTT;      // uint4 per thread per sub-batch (<= 4)
    uint4 pre[kVec];
    auto prefetch = [&](int sub) {
        const long long first = cta_first + (long long)sub * kTT;
        const long long left = n - first;
        const size_t bytes = left <= 0 ? 0 : (size_t)min((long long)kTT, left) * stride;
        const uint4* s4 = reinterpret_cast<const uint4*>(pkts + (size_t)first * stride);
#pragma unroll
        for (int j = 0; j < kVec; ++j) {
            const size_t i = (size_t)j * kTT + threadIdx.x;
            pre[j] = (i * 16 + 16 <= bytes) ? __ldg(s4 + i) : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    const bool vec_ok = (reinterpret_cast<uintptr_t>(pkts) & 15) == 0 && ((kTT * stride) & 15) == 0;
    if (vec_ok) prefetch(0);
    for (int sub = 0; sub < kSub; ++sub) {
        const long long first = cta_first + (long long)sub * kTT;
        if (first >= n) break;
        const int count = (int)min((long long)kTT, n - first);
        __syncthreads();
        if (vec_ok) {
            const size_t bytes = (size_t)count * stride;
            uint4* d4 = reinterpret_cast<uint4*>(s_rec);
#pragma unroll
            for (int j = 0; j < kVec; ++j) {
                const size_t i = (size_t)j * kTT + threadIdx.x;
                if (i * 16 + 16 <= bytes) d4[i] = pre[j];
            }
            for (size_t i = (bytes / 16) * 16 + threadIdx.x; i < bytes; i += kTT) s_rec[i] = __ldg(pkts + (size_t)first * stride + i);
            if (sub + 1 < kSub) prefetch(sub + 1);
        } else {
            stage_records_t(pkts + (size_t)first * stride, (size_t)count * stride, s_rec);
        }
        __syncthreads();
        if ((int)threadIdx.x < count) {
            const long long k = first + threadIdx.x;
            double rx, ry, ryaw;
            float dist[4];
            c[OCCGRID_C_PACKETS] += 1;
            int tile = -1;
            const int st = decode_packet(s_rec + threadIdx.x * stride, k, agent_idx, drift, agent_off, n_agents, &rx, &ry, &ryaw, dist);
            if (st == PKT_DROPPED) c[OCCGRID_C_DROPPED] += 1;
            else if (st == PKT_BAD_POSE) c[OCCGRID_C_BAD_POSE] += 1;
            else {
                c[OCCGRID_C_ACCEPTED] += 1;
                c[OCCGRID_C_BEAMS] += 4;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    const double d = (double)dist[s];
                    c[OCCGRID_C_HITS] += (OCC_MIN_DIST_M < d && d <= OCC_MAX_DIST_M) ? 1 : 0;     // :888
                }
                tile = home_tile(g, tg, rx, ry);
                if (tile >= 0) {
                    atomicAdd(&s_vals[hash_insert(s_keys, (unsigned int)tile)], 1u);
                    PoseRec r;
                    r.rx = rx; r.ry = ry; r.yaw = (float)ryaw;      // ryaw came from an fp32 field: exact
                    r.d[0] = dist[0]; r.d[1] = dist[1]; r.d[2] = dist[2]; r.d[3] = dist[3];
                    r.k = (unsigned int)k;
                    r.pad[0] = r.pad[1] = 0;
                    recs[k] = r;                                    // packet order: coalesced 48-byte stores
                }
            }
            tile_ids[k] = tile;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHash; i += kTT)
        if (s_keys[i] != kEmpty && atomicAdd(&tile_count[s_keys[i]], s_vals[i]) == 0u)
            active[atomicAdd(&hdr->active_count, 1u)] = s_keys[i];          // first toucher lists the tile
    // the staging buffer is free now: reuse it for the counter reduction (48 KB static limit)
    block_add_counters(c, reinterpret_cast<unsigned long long*>(s_rec), counters);
}

// Same as k_home_count for input that is already decoded (routed pose records): the records
// are binned in place.
__global__ void __launch_bounds__(kTT)
k_home_count_poses(Geom g, TileGeom tg, const PoseRec* __restrict__ recs, long long n,
                   unsigned int* __restrict__ tile_count, int* __restrict__ tile_ids,
                   unsigned int* __restrict__ active, TilePlanHeader* __restrict__ hdr, uint64_t* counters) {
    __shared__ unsigned int s_keys[kHash];
    __shared__ unsigned int s_vals[kHash];
    __shared__ unsigned long long s_acc[(OCCGRID_C_HITS + 1) * 32];
    for (int i = threadIdx.x; i < kHash; i += kTT) { s_keys[i] = kEmpty; s_vals[i] = 0u; }
    __syncthreads();
    unsigned long long c[OCCGRID_C_HITS + 1] = {};
    const long long cta_first = (long long)blockIdx.x * kPkPerCta;
    for (int sub = 0; sub < kSub; ++sub) {
        const long long k = cta_first + (long long)sub * kTT + threadIdx.x;
        if (k >= n) break;
        const PoseRec r = recs[k];
        c[OCCGRID_C_PACKETS] += 1;
        int tile = -1;
        if (!(isfinite(r.rx) && isfinite(r.ry) && isfinite(r.yaw))) c[OCCGRID_C_BAD_POSE] += 1;
        else {
            c[OCCGRID_C_ACCEPTED] += 1;
            c[OCCGRID_C_BEAMS] += 4;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const double d = (double)r.d[s];
                c[OCCGRID_C_HITS] += (OCC_MIN_DIST_M < d && d <= OCC_MAX_DIST_M) ? 1 : 0;
            }
            tile = home_tile(g, tg, r.rx, r.ry);
            if (tile >= 0) atomicAdd(&s_vals[hash_insert(s_keys, (unsigned int)tile)], 1u);
        }
        tile_ids[k] = tile;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHash; i += kTT)
        if (s_keys[i] != kEmpty && atomicAdd(&tile_count[s_keys[i]], s_vals[i]) == 0u)
            active[atomicAdd(&hdr->active_count, 1u)] = s_keys[i];          // first toucher lists the tile
    block_add_counters(c, s_acc, counters);
}

__global__ void __launch_bounds__(kTT)
k_home_scatter(long long n, const int* __restrict__ tile_ids, const unsigned int* __restrict__ tile_offset,
               unsigned int* __restrict__ tile_cursor, const TilePlanHeader* __restrict__ hdr,
               unsigned int* __restrict__ bins) {
    __shared__ unsigned int s_keys[kHash];
    __shared__ unsigned int s_vals[kHash];
    if (hdr->overflow) return;
    for (int i = threadIdx.x; i < kHash; i += kTT) { s_keys[i] = kEmpty; s_vals[i] = 0u; }
    __syncthreads();
    const long long cta_first = (long long)blockIdx.x * kPkPerCta;
    unsigned int where[kSub];          // (hash slot << 16) | rank within this CTA's share of the tile
#pragma unroll
    for (int sub = 0; sub < kSub; ++sub) {
        const long long k = cta_first + (long long)sub * kTT + threadIdx.x;
        where[sub] = kEmpty;
        if (k < n) {
            const int tile = tile_ids[k];
            if (tile >= 0) {
                const int h = hash_insert(s_keys, (unsigned int)tile);
                where[sub] = ((unsigned int)h << 16) | atomicAdd(&s_vals[h], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHash; i += kTT)          // one reservation per distinct tile
        if (s_keys[i] != kEmpty) s_vals[i] = tile_offset[s_keys[i]] + atomicAdd(&tile_cursor[s_keys[i]], s_vals[i]);
    __syncthreads();
#pragma unroll
    for (int sub = 0; sub < kSub; ++sub) {
        const long long k = cta_first + (long long)sub * kTT + threadIdx.x;
        if (where[sub] != kEmpty) bins[s_vals[where[sub] >> 16] + (where[sub] & 0xffffu)] = (unsigned int)k;
    }
}

// One CTA over the ACTIVE tiles only (listed by the count pass, in arbitrary order — nothing
// downstream depends on the order): exclusive scan of their counts -> bin offsets, cursors
// zeroed, (tile, chunk) work items built.  tile_count is re-zeroed for the next call.
__global__ void __launch_bounds__(1024)
k_tile_plan(unsigned int* __restrict__ tile_count, unsigned int* __restrict__ tile_offset,
            unsigned int* __restrict__ tile_cursor, uint4* __restrict__ items, unsigned int max_items,
            const unsigned int* __restrict__ active, unsigned long long max_records, TilePlanHeader* __restrict__ hdr,
            uint64_t* counters) {
    __shared__ unsigned int s_warp[2][33];
    __shared__ unsigned int s_carry[2];
    if (threadIdx.x < 2) s_carry[threadIdx.x] = 0;
    __syncthreads();
    const int n_active = (int)hdr->active_count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < n_active; start += blockDim.x) {
        const int i = start + threadIdx.x;
        const unsigned int t = i < n_active ? active[i] : 0u;
        const unsigned int cnt = i < n_active ? tile_count[t] : 0u;
        unsigned int v[2] = {cnt, (cnt + kChunkPk - 1) / kChunkPk};
        unsigned int ex[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            unsigned int inc = v[j];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            if (lane == 31) s_warp[j][warp] = inc;
            ex[j] = inc - v[j];
        }
        __syncthreads();
        if (warp < 2) {
            unsigned int w = s_warp[warp][lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned int u = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += u; }
            s_warp[warp][lane] = winc - w;
            if (lane == 31) s_warp[warp][32] = winc;
        }
        __syncthreads();
        const unsigned int off = s_carry[0] + s_warp[0][warp] + ex[0];
        const unsigned int item0 = s_carry[1] + s_warp[1][warp] + ex[1];
        if (i < n_active) {
            tile_offset[t] = off;
            tile_cursor[t] = 0u;
            tile_count[t] = 0u;
            for (unsigned int j = 0; j < v[1]; ++j) {
                const unsigned int b = off + j * kChunkPk;
                const unsigned int e = min(off + cnt, b + kChunkPk);
                if (item0 + j < max_items) items[item0 + j] = make_uint4(t, b, e, 0u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 2) s_carry[threadIdx.x] += s_warp[threadIdx.x][32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        TilePlanHeader h;
        h.total_records = s_carry[0];
        h.n_items = s_carry[1];
        h.n_active = (unsigned int)n_active;
        h.work_counter = 0;
        h.resolve_counter = 0;
        h.active_count = 0;                  // ready for the next call's count pass
        h.overflow = (s_carry[0] > max_records || s_carry[1] > max_items) ? 1u : 0u;
        if (h.overflow) { h.n_items = 0; h.n_active = 0; }
        h.pad = 0;
        *hdr = h;
        if (counters) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + OCCGRID_C_RECORDS, (unsigned long long)s_carry[0]);
    }
}

// Walk one beam into the shared-memory window (exact reference Bresenham, :158-179: strict
// `e2 > -dy` / `e2 < dx`, both may fire, sx = -1 when x0 == x1).  (x, y) are window-local.
// Both end points inside the (convex) window => every cell is inside, so the hot loop carries
// no bounds test and steps a running shared-memory offset instead of recomputing y*pitch+x.
__device__ __forceinline__ void draw_beam_smem(unsigned int* __restrict__ s_win, int side, int pitch, int x, int y,
                                               int ddx, int ddy, unsigned int free_stamp, bool hit, bool skip_first) {
    const int dx = abs(ddx), dy = abs(ddy);
    int err = dx - dy;
    const int n = max(dx, dy);
    const unsigned int us = (unsigned int)side;
    if ((unsigned int)x < us && (unsigned int)y < us && (unsigned int)(x + ddx) < us && (unsigned int)(y + ddy) < us) {
        // Major/minor form of the same recurrence.  With err = dx - dy the reference steps the
        // major axis on EVERY iteration (for dx > dy: 2*err > -dy always holds; for dy > dx:
        // 2*err < dx always holds; both step when dx == dy), and the minor axis steps iff
        //   x-major: 2*err < dx        y-major: 2*err > -dy  <=>  2*(-err) < dy
        // i.e. with E = |err-form| = dmaj - dmin initially:  minor iff 2*E < dmaj (strict, like
        // the reference), then E += (minor ? dmaj : 0) - dmin.  Branch-free across octants, so
        // lanes walking different octants do not diverge.  (Exhaustive KAT on device:
        // tests/test_gpu_integrate.py::test_bresenham_exhaustive_on_device.)
        const int sxo = ddx > 0 ? 1 : -1;
        const int syo = ddy > 0 ? pitch : -pitch;
        const bool xmajor = dx >= dy;
        const int dmaj = xmajor ? dx : dy, dmin = xmajor ? dy : dx;
        const int step_maj = xmajor ? sxo : syo;
        const int step_both = sxo + syo;
        const int e_minor = dmaj - dmin;         // E change when the minor axis steps too
        int E = dmaj - dmin;
        int off = y * pitch + x;
        int i = 0;
        if (skip_first) {                     // the start cell is overwritten by a later beam of the packet
            if (n == 0) return;
            const bool minor = 2 * E < dmaj;
            off += minor ? step_both : step_maj;
            E += minor ? e_minor : -dmin;
            i = 1;
        }
        for (; i < n; ++i) {
            atomicMax(&s_win[off], free_stamp);
            const bool minor = 2 * E < dmaj;
            off += minor ? step_both : step_maj;
            E += minor ? e_minor : -dmin;
        }
        if (hit) atomicMax(&s_win[off], free_stamp | 1u);
        return;
    }
    // guarded path (cannot happen while the reach bound holds): per-cell window test
    const int sx = ddx > 0 ? 1 : -1, sy = ddy > 0 ? 1 : -1;
    for (int i = 0; i < n; ++i) {
        if (!(i == 0 && skip_first) && (unsigned int)x < us && (unsigned int)y < us) atomicMax(&s_win[y * pitch + x], free_stamp);
        const int e2 = 2 * err;
        if (e2 > -dy) { err -= dy; x += sx; }
        if (e2 < dx)  { err += dx; y += sy; }
    }
    if (hit && !(n == 0 && skip_first) && (unsigned int)x < us && (unsigned int)y < us) atomicMax(&s_win[y * pitch + x], free_stamp | 1u);
}

// Hit/miss COUNT mode (extension): the same walk, but every cell update_ray (:148-156)
// would have stored FREE to counts one miss (low half of the window word) and the OCCUPIED end
// cell of a valid hit counts one hit (high half).  A work item holds <= 4 * kChunkPk beams and a
// beam visits a cell once, so 16 bits per half cannot overflow.
__device__ __forceinline__ void count_beam_smem(unsigned int* __restrict__ s_win, int side, int pitch, int x, int y,
                                                int ddx, int ddy, bool hit) {
    const int dx = abs(ddx), dy = abs(ddy);
    const int n = max(dx, dy);
    const unsigned int us = (unsigned int)side;
    if ((unsigned int)x < us && (unsigned int)y < us && (unsigned int)(x + ddx) < us && (unsigned int)(y + ddy) < us) {
        const int sxo = ddx > 0 ? 1 : -1;
        const int syo = ddy > 0 ? pitch : -pitch;
        const bool xmajor = dx >= dy;
        const int dmaj = xmajor ? dx : dy, dmin = xmajor ? dy : dx;
        const int step_maj = xmajor ? sxo : syo;
        const int step_both = sxo + syo;
        const int e_minor = dmaj - dmin;
        int E = dmaj - dmin;
        int off = y * pitch + x;
        for (int i = 0; i < n; ++i) {
            atomicAdd(&s_win[off], 1u);
            const bool minor = 2 * E < dmaj;
            off += minor ? step_both : step_maj;
            E += minor ? e_minor : -dmin;
        }
        if (hit) atomicAdd(&s_win[off], 0x10000u);
        return;
    }
    int err = dx - dy;
    const int sx = ddx > 0 ? 1 : -1, sy = ddy > 0 ? 1 : -1;
    for (int i = 0; i < n; ++i) {
        if ((unsigned int)x < us && (unsigned int)y < us) atomicAdd(&s_win[y * pitch + x], 1u);
        const int e2 = 2 * err;
        if (e2 > -dy) { err -= dy; x += sx; }
        if (e2 < dx)  { err += dx; y += sy; }
    }
    if (hit && (unsigned int)x < us && (unsigned int)y < us) atomicAdd(&s_win[y * pitch + x], 0x10000u);
}

template <bool kCounts>
__global__ void __launch_bounds__(kTT)
k_home_raycast(Geom g, TileGeom tg, const uint4* __restrict__ items, TilePlanHeader* __restrict__ hdr,
               const unsigned int* __restrict__ bins, const PoseRec* __restrict__ recs, int ordinals_in_records,
               unsigned int* __restrict__ stamps, uint64_t* counters) {
    extern __shared__ unsigned int s_win[];
    __shared__ unsigned int s_item;
    __shared__ unsigned long long s_acc[3 * 32];
    const unsigned int n_items = hdr->n_items;
    const int side = tg.win_side, pitch = tg.pitch, words = side * pitch;
    unsigned long long c[3] = {0, 0, 0};     // updates, slowpath, owned updates
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&hdr->work_counter, 1u);
        for (int i = threadIdx.x; i < words; i += kTT) s_win[i] = 0u;
        __syncthreads();
        const unsigned int it = s_item;
        if (it >= n_items) break;
        const uint4 item = items[it];
        const int ttx = item.x % tg.tiles_x, tty = item.x / tg.tiles_x;
        // global cell of window-local (0, 0)
        const int wx0 = g.win_x0 - tg.pad + (ttx << kTileShift) - tg.reach;
        const int wy0 = g.win_y0 - tg.pad + (tty << kTileShift) - tg.reach;
        for (unsigned int r = item.y + threadIdx.x; r < item.z; r += kTT) {
            const unsigned int idx = bins[r];
            const PoseRec rec = recs[idx];
            const unsigned int k = ordinals_in_records ? rec.k : idx;   // packet ordinal in this batch
            const float dist[4] = {rec.d[0], rec.d[1], rec.d[2], rec.d[3]};
            Beam b[4];
            expand_packet(g, rec.rx, rec.ry, (double)rec.yaw, dist, LibSinCos(), b);
            bool later_writes_first = false;
#pragma unroll
            for (int s = 3; s >= 0; --s) {
                if (!b[s].valid) continue;
                const int cells = beam_cells(b[s]);
                c[0] += cells;
                c[1] += b[s].slow;
                if (b[s].x0 >= g.win_x0 && b[s].x0 < g.win_x0 + g.win_w && b[s].y0 >= g.win_y0 && b[s].y0 < g.win_y0 + g.win_h)
                    c[2] += cells;
                if (kCounts)
                    count_beam_smem(s_win, side, pitch, b[s].x0 - wx0, b[s].y0 - wy0, b[s].x1 - b[s].x0, b[s].y1 - b[s].y0,
                                    b[s].hit != 0);
                else
                    draw_beam_smem(s_win, side, pitch, b[s].x0 - wx0, b[s].y0 - wy0, b[s].x1 - b[s].x0, b[s].y1 - b[s].y0,
                                   (k * 4u + (unsigned int)s + 1u) << 1, b[s].hit != 0, later_writes_first);
                later_writes_first = later_writes_first || (cells > 1 || b[s].hit);
            }
        }
        __syncthreads();
        // flush: window rows are contiguous in the stamp plane -> coalesced reductions
        const int lx_lo = max(0, g.win_x0 - wx0), lx_hi = min(side, g.win_x0 + g.win_w - wx0);
        const int ly_lo = max(0, g.win_y0 - wy0), ly_hi = min(side, g.win_y0 + g.win_h - wy0);
        if (lx_hi > lx_lo) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            for (int ly = ly_lo + warp; ly < ly_hi; ly += kTT / 32) {          // one warp per window row
                const unsigned int* row = s_win + ly * pitch;
                unsigned int* out = stamps + (size_t)(wy0 + ly - g.win_y0) * g.win_w + (wx0 - g.win_x0);
                for (int lx = lx_lo + lane; lx < lx_hi; lx += 32) {
                    const unsigned int v = row[lx];
                    if (!v) continue;
                    if (kCounts)     // {miss, hit} int32 pair of the cell: one 64-bit add (the halves cannot carry into each other)
                        atomicAdd(reinterpret_cast<unsigned long long*>(stamps) + ((size_t)(wy0 + ly - g.win_y0) * g.win_w + (wx0 - g.win_x0) + lx),
                                  ((unsigned long long)(v >> 16) << 32) | (unsigned long long)(v & 0xffffu));
                    else
                        atomicMax(out + lx, v);
                }
            }
        }
    }
    if (counters) {
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            unsigned long long x = c[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) s_acc[i * 32 + warp] = x;
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            unsigned long long t = 0;
            for (int w = 0; w < kTT / 32; ++w) t += s_acc[threadIdx.x * 32 + w];
            const int slot = threadIdx.x == 0 ? OCCGRID_C_UPDATES : (threadIdx.x == 1 ? OCCGRID_C_SLOWPATH : OCCGRID_C_OWNED_UPDATES);
            if (t) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + slot, t);
        }
    }
}

constexpr int kResolveSplit = 8;      // CTAs per active window (row slices) — few tiles are active, keep all SMs busy

__global__ void __launch_bounds__(kTT)
k_home_resolve(Geom g, TileGeom tg, const unsigned int* __restrict__ active, TilePlanHeader* __restrict__ hdr,
               unsigned int* __restrict__ stamps, int8_t* __restrict__ grid) {
    __shared__ unsigned int s_item;
    const unsigned int n_work = hdr->n_active * kResolveSplit;
    const int side = tg.win_side;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&hdr->resolve_counter, 1u);
        __syncthreads();
        const unsigned int it = s_item;
        if (it >= n_work) break;
        const unsigned int t = active[it / kResolveSplit];
        const int part = (int)(it % kResolveSplit);
        const int ttx = t % tg.tiles_x, tty = t / tg.tiles_x;
        const int wx0 = -tg.pad + (ttx << kTileShift) - tg.reach;      // window-relative
        const int wy0 = -tg.pad + (tty << kTileShift) - tg.reach;
        const int lx_lo = max(0, -wx0), lx_hi = min(side, g.win_w - wx0);
        const int ly_lo = max(0, -wy0), ly_hi = min(side, g.win_h - wy0);
        if (lx_hi <= lx_lo) continue;
        for (int ly = ly_lo + part * (kTT / 32) + warp; ly < ly_hi; ly += kResolveSplit * (kTT / 32)) {   // one warp per row
            const size_t row = (size_t)(wy0 + ly) * g.win_w + wx0;
            for (int lx = lx_lo + lane; lx < lx_hi; lx += 32) {
                const unsigned int s = stamps[row + lx];
                if (s) {
                    grid[row + lx] = (s & 1u) ? OCCGRID_CELL_OCCUPIED : OCCGRID_CELL_FREE;
                    stamps[row + lx] = 0u;
                }
            }
        }
    }
}

// ---- host side ----------------------------------------------------------------------------

// Persistent raycast CTAs per SM (0 = as many as fit).  A multi-GPU pipeline lowers it so that
// the routing kernel of the next batch can be co-resident (occgrid_set_raycast_ctas_per_sm).
static int g_raycast_cta_cap = 0;
void set_raycast_cta_cap(int cap) { g_raycast_cta_cap = cap < 0 ? 0 : cap; }

struct TiledLayout {
    size_t off_stamps, off_count, off_offset, off_cursor, off_active, off_hdr, off_items, off_ids, off_bins, off_recs, total;
    unsigned long long max_records;
    unsigned int max_items;
};

static TiledLayout tiled_layout(const occgrid_geom* geom, int64_t max_packets) {
    TiledLayout L;
    const TileGeom tg = tile_geom(geom);
    L.max_records = (unsigned long long)max_packets;
    unsigned long long items = L.max_records / kChunkPk + (unsigned long long)tg.n_tiles + 1;
    if (items > L.max_records + 1) items = L.max_records + 1;
    L.max_items = (unsigned int)items;
    size_t o = 0;
    L.off_stamps = o; o += align_up((size_t)geom->win_w * geom->win_h * 4, 256);
    L.off_count = o;  o += align_up((size_t)(tg.n_tiles + 1) * 4, 256);
    L.off_offset = o; o += align_up((size_t)(tg.n_tiles + 1) * 4, 256);
    L.off_cursor = o; o += align_up((size_t)(tg.n_tiles + 1) * 4, 256);
    L.off_active = o; o += align_up((size_t)(tg.n_tiles + 1) * 4, 256);
    L.off_hdr = o;    o += 256;
    L.off_items = o;  o += align_up((size_t)L.max_items * sizeof(uint4), 256);
    L.off_ids = o;    o += align_up((size_t)L.max_records * sizeof(int), 256);
    L.off_bins = o;   o += align_up((size_t)L.max_records * sizeof(unsigned int), 256);
    L.off_recs = o;   o += align_up((size_t)L.max_records * sizeof(PoseRec), 256);
    L.total = o;
    return L;
}

bool tiled_supported(const occgrid_geom* geom) {
    const TileGeom tg = tile_geom(geom);
    return (size_t)tg.win_side * tg.pitch * 4 <= (size_t)kMaxWindowBytes && tg.n_tiles <= (1 << 24);
}

size_t tiled_workspace_bytes(const occgrid_geom* geom, int64_t max_packets) {
    return tiled_layout(geom, max_packets < 1 ? 1 : max_packets).total;
}

// `d_poses` != NULL: input is n PoseRec (already decoded and corrected), `d_packets` etc. unused.
int integrate_tiled(const occgrid_geom* geom, const uint8_t* d_packets, const PoseRec* d_poses, int ordinals_in_records,
                    int64_t n, int stride,
                    const int32_t* d_agent_idx, const double* d_drift, const double* d_agent_off,
                    int n_agents, int8_t* d_grid, int32_t* d_counts, void* d_ws, size_t ws_bytes, uint64_t* d_counters,
                    cudaStream_t st) {
    const TiledLayout L = tiled_layout(geom, n);
    if (ws_bytes < L.total) {
        set_last_error("workspace %zu B < %zu B needed by TILED for %lld records", ws_bytes, L.total, (long long)n);
        return OCCGRID_E_WORKSPACE;
    }
    char* ws = reinterpret_cast<char*>(d_ws);
    unsigned int* stamps = reinterpret_cast<unsigned int*>(ws + L.off_stamps);
    unsigned int* tile_count = reinterpret_cast<unsigned int*>(ws + L.off_count);
    unsigned int* tile_offset = reinterpret_cast<unsigned int*>(ws + L.off_offset);
    unsigned int* tile_cursor = reinterpret_cast<unsigned int*>(ws + L.off_cursor);
    unsigned int* active = reinterpret_cast<unsigned int*>(ws + L.off_active);
    TilePlanHeader* hdr = reinterpret_cast<TilePlanHeader*>(ws + L.off_hdr);
    uint4* items = reinterpret_cast<uint4*>(ws + L.off_items);
    unsigned int* bins = reinterpret_cast<unsigned int*>(ws + L.off_bins);
    const PoseRec* recs = d_poses ? d_poses : reinterpret_cast<const PoseRec*>(ws + L.off_recs);
    int* tile_ids = reinterpret_cast<int*>(ws + L.off_ids);
    const Geom g = to_geom(geom);
    const TileGeom tg = tile_geom(geom);
    const unsigned int blocks = (unsigned int)((n + kPkPerCta - 1) / kPkPerCta);
    const size_t win_bytes = (size_t)tg.win_side * tg.pitch * 4;
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    static thread_local size_t configured_smem = 0;
    if (win_bytes > configured_smem) {
        OCC_CUDA_TRY(cudaFuncSetAttribute(k_home_raycast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_bytes));
        OCC_CUDA_TRY(cudaFuncSetAttribute(k_home_raycast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_bytes));
        configured_smem = win_bytes;
    }
    int ctas_per_sm = 1;
    const cudaError_t occ_rc = d_counts ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_home_raycast<true>, kTT, win_bytes)
                                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_home_raycast<false>, kTT, win_bytes);
    if (occ_rc != cudaSuccess || ctas_per_sm < 1)
        ctas_per_sm = 1;
    if (g_raycast_cta_cap > 0 && ctas_per_sm > g_raycast_cta_cap) ctas_per_sm = g_raycast_cta_cap;
    {
        ProfileScope ps(K_TILE_COUNT, st);
        if (d_poses)
            k_home_count_poses<<<blocks, kTT, 0, st>>>(g, tg, d_poses, n, tile_count, tile_ids, active, hdr, d_counters);
        else
            k_home_count<<<blocks, kTT, 0, st>>>(g, tg, d_packets, n, stride, d_agent_idx, d_drift, d_agent_off, n_agents,
                                                tile_count, tile_ids, reinterpret_cast<PoseRec*>(ws + L.off_recs), active, hdr,
                                                d_counters);
    }
    {
        ProfileScope ps(K_TILE_SCAN, st);
        k_tile_plan<<<1, 1024, 0, st>>>(tile_count, tile_offset, tile_cursor, items, L.max_items, active, L.max_records, hdr,
                                        d_counters);
    }
    {
        ProfileScope ps(K_TILE_SCATTER, st);
        k_home_scatter<<<blocks, kTT, 0, st>>>(n, tile_ids, tile_offset, tile_cursor, hdr, bins);
    }
    {
        ProfileScope ps(K_TILE_RAYCAST, st);
        if (d_counts)
            k_home_raycast<true><<<sms * ctas_per_sm, kTT, win_bytes, st>>>(g, tg, items, hdr, bins, recs, ordinals_in_records,
                                                                            reinterpret_cast<unsigned int*>(d_counts), d_counters);
        else
            k_home_raycast<false><<<sms * ctas_per_sm, kTT, win_bytes, st>>>(g, tg, items, hdr, bins, recs, ordinals_in_records, stamps,
                                                                             d_counters);
    }
    if (!d_counts) {
        ProfileScope ps(K_TILE_RESOLVE, st);
        k_home_resolve<<<sms * 8, kTT, 0, st>>>(g, tg, active, hdr, stamps, d_grid);
    }
    OCC_CUDA_TRY(cudaGetLastError());
    return OCCGRID_OK;
}

int integrate_packets_tiled(const occgrid_geom* geom, const uint8_t* d_packets, int64_t n, int stride,
                            const int32_t* d_agent_idx, const double* d_drift, const double* d_agent_off,
                            int n_agents, int8_t* d_grid, void* d_ws, size_t ws_bytes, uint64_t* d_counters,
                            cudaStream_t st) {
    return integrate_tiled(geom, d_packets, nullptr, 0, n, stride, d_agent_idx, d_drift, d_agent_off, n_agents, d_grid, nullptr, d_ws,
                           ws_bytes, d_counters, st);
}

int accumulate_packets_tiled(const occgrid_geom* geom, const uint8_t* d_packets, int64_t n, int stride,
                             const int32_t* d_agent_idx, const double* d_drift, const double* d_agent_off,
                             int n_agents, int32_t* d_counts, void* d_ws, size_t ws_bytes, uint64_t* d_counters,
                             cudaStream_t st) {
    return integrate_tiled(geom, d_packets, nullptr, 0, n, stride, d_agent_idx, d_drift, d_agent_off, n_agents, nullptr, d_counts, d_ws,
                           ws_bytes, d_counters, st);
}

int integrate_poses_tiled(const occgrid_geom* geom, const void* d_poses, int64_t n, int ordinals_in_records, int8_t* d_grid,
                          void* d_ws, size_t ws_bytes, uint64_t* d_counters, cudaStream_t st) {
    return integrate_tiled(geom, nullptr, reinterpret_cast<const PoseRec*>(d_poses), ordinals_in_records, n, 0, nullptr, nullptr,
                           nullptr, 0,
                           d_grid, nullptr, d_ws, ws_bytes, d_counters, st);
}

}  // namespace occ
