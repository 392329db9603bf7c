// Strategy TILED of occgrid_integrate_packets: beams binned by grid tile, last-writer-wins
// resolved with SHARED-MEMORY atomics.
//
// Same reference semantics as occgrid_integrate.cu (server_nodes/dual_bot_mapper.py:826-903,
// :136-179; every cell keeps the largest  (beam ordinal + 1) << 1 | occupied  stamp), but the
// per-cell atomicMax runs on a 64x64-cell stamp tile in shared memory, where B200 sustains
// ~1.0-1.5e12 atomics/s (measured, occgrid_scatter_probe kind 2) against ~2.2e11/s for
// red.global (kind 0).  Only the surviving stamp of each touched cell goes to L2.
//
//   k_tile_count     thread/packet: decode, expand, count (tile, beam) records per tile
//   k_tile_plan      one CTA: exclusive scan of the tile counts -> bin offsets, list of
//                    (tile, chunk) work items of <= kChunk records, list of active tiles
//   k_tile_scatter   thread/packet: expand again, write 16-byte beam records into the bins
//   k_tile_raycast   persistent CTAs pull work items: zero the smem tile, walk every record's
//                    Bresenham line (exact reference tie-breaks) with atomicMax in smem on the
//                    cells that fall in the tile, then flush non-zero stamps with coalesced
//                    red.global.max to the global stamp plane
//   k_tile_resolve   persistent CTAs over the active tiles: stamps -> int8 grid, stamps := 0
//
// A beam is recorded in every tile its bounding box overlaps (1-4 for <=25-cell rays and
// 64-cell tiles); each copy walks the same global line and masks to its tile, so tiles — like
// multi-GPU windows — reproduce the reference's per-cell clipping (:149, :155) exactly.
#include "common.cuh"

namespace occ {

constexpr int kTile = 64;            // cells per tile side
constexpr int kTileShift = 6;
constexpr int kTilePitch = kTile + 1;   // smem row pitch in words: spreads vertical rays over banks
constexpr int kTT = 256;             // threads per CTA
constexpr int kChunk = 8192;         // records per work item
constexpr int kMaxStrideT = 64;

struct __align__(16) BeamRec {
    short lx0, ly0;        // start cell relative to the tile origin (may lie outside the tile)
    short ddx, ddy;        // x1 - x0, y1 - y0
    unsigned int stamp;    // ((ordinal + 1) << 1) | hit_valid
    unsigned int flags;    // bit 0: skip the write to the start cell
};

struct TilePlanHeader {
    unsigned int n_items;
    unsigned int n_active;
    unsigned int total_records;
    unsigned int work_counter;
    unsigned int resolve_counter;
    unsigned int overflow;
    unsigned int pad[2];
};

struct TileGeom {
    int tiles_x, tiles_y;
    int n_tiles;
};

__host__ __device__ inline TileGeom tile_geom(int win_w, int win_h) {
    TileGeom t;
    t.tiles_x = (win_w + kTile - 1) >> kTileShift;
    t.tiles_y = (win_h + kTile - 1) >> kTileShift;
    t.n_tiles = t.tiles_x * t.tiles_y;
    return t;
}

// Tile range of a beam's bounding box in window-relative tile coordinates (clamped).
// Returns false when the box misses the window's tiles entirely.
__device__ __forceinline__ bool beam_tile_range(const Geom& g, const TileGeom& tg, const Beam& b,
                                                int* tx0, int* tx1, int* ty0, int* ty1) {
    if (!b.valid) return false;
    const int ax = min(b.x0, b.x1) - g.win_x0, bx = max(b.x0, b.x1) - g.win_x0;
    const int ay = min(b.y0, b.y1) - g.win_y0, by = max(b.y0, b.y1) - g.win_y0;
    if (bx < 0 || by < 0 || ax >= g.win_w || ay >= g.win_h) return false;
    *tx0 = max(ax, 0) >> kTileShift;
    *tx1 = min(bx, g.win_w - 1) >> kTileShift;
    *ty0 = max(ay, 0) >> kTileShift;
    *ty1 = min(by, g.win_h - 1) >> kTileShift;
    return true;
}

__device__ __forceinline__ void stage_records_t(const uint8_t* __restrict__ src, size_t bytes, uint8_t* smem) {
    const size_t nvec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? bytes / 16 : 0;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(smem);
    for (size_t i = threadIdx.x; i < nvec; i += blockDim.x) d4[i] = __ldg(s4 + i);
    for (size_t i = nvec * 16 + threadIdx.x; i < bytes; i += blockDim.x) smem[i] = __ldg(src + i);
}

// Shared by the count and scatter passes: decode + expand one packet and hand every
// (tile, beam) pair to `emit(tile, beam index, beam, skip_first)`.
template <bool kCountStats, class Emit>
__device__ __forceinline__ void for_each_record(const Geom& g, const TileGeom& tg, const uint8_t* rec, long long k,
                                                const int32_t* agent_idx, const double* drift,
                                                const double* agent_off, int n_agents,
                                                unsigned long long (&c)[OCCGRID_C_OWNED_UPDATES + 1], Emit&& emit) {
    double rx, ry, ryaw;
    float dist[4];
    if (kCountStats) c[OCCGRID_C_PACKETS] = 1;
    const int st = decode_packet(rec, k, agent_idx, drift, agent_off, n_agents, &rx, &ry, &ryaw, dist);
    if (st == PKT_DROPPED) { if (kCountStats) c[OCCGRID_C_DROPPED] = 1; return; }
    if (st == PKT_BAD_POSE) { if (kCountStats) c[OCCGRID_C_BAD_POSE] = 1; return; }
    if (kCountStats) c[OCCGRID_C_ACCEPTED] = 1;
    Beam b[4];
    expand_packet(g, rx, ry, ryaw, dist, LibSinCos(), b);
    bool later_writes_first = false;
#pragma unroll
    for (int s = 3; s >= 0; --s) {
        const int cells = b[s].valid ? beam_cells(b[s]) : 0;
        if (kCountStats) {
            c[OCCGRID_C_BEAMS] += 1;
            c[OCCGRID_C_HITS] += b[s].hit;
            c[OCCGRID_C_UPDATES] += cells;
            c[OCCGRID_C_SLOWPATH] += b[s].slow;
            if (b[s].valid && b[s].x0 >= g.win_x0 && b[s].x0 < g.win_x0 + g.win_w && b[s].y0 >= g.win_y0 &&
                b[s].y0 < g.win_y0 + g.win_h)
                c[OCCGRID_C_OWNED_UPDATES] += cells;
        }
        int tx0, tx1, ty0, ty1;
        if (beam_tile_range(g, tg, b[s], &tx0, &tx1, &ty0, &ty1)) {
            for (int ty = ty0; ty <= ty1; ++ty)
                for (int tx = tx0; tx <= tx1; ++tx) emit(ty * tg.tiles_x + tx, tx, ty, s, b[s], later_writes_first);
        }
        later_writes_first = later_writes_first || (b[s].valid && (cells > 1 || b[s].hit));
    }
}

__global__ void __launch_bounds__(kTT)
k_tile_count(Geom g, TileGeom tg, const uint8_t* __restrict__ pkts, long long n, int stride,
             const int32_t* __restrict__ agent_idx, const double* __restrict__ drift,
             const double* __restrict__ agent_off, int n_agents,
             unsigned int* __restrict__ tile_count, uint64_t* counters) {
    __shared__ __align__(16) uint8_t s_rec[kTT * kMaxStrideT];
    __shared__ unsigned long long s_acc[OCCGRID_C_OWNED_UPDATES + 1];
    const long long first = (long long)blockIdx.x * kTT;
    const int count = (int)min((long long)kTT, n - first);
    stage_records_t(pkts + (size_t)first * stride, (size_t)count * stride, s_rec);
    __syncthreads();
    unsigned long long c[OCCGRID_C_OWNED_UPDATES + 1] = {};
    if ((int)threadIdx.x < count) {
        int last_tile = -1;
        unsigned int run = 0;
        for_each_record<true>(g, tg, s_rec + threadIdx.x * stride, first + threadIdx.x, agent_idx, drift, agent_off,
                              n_agents, c,
                              [&](int tile, int, int, int, const Beam&, bool) {
                                  if (tile == last_tile) { ++run; return; }
                                  if (run) atomicAdd(&tile_count[last_tile], run);
                                  last_tile = tile;
                                  run = 1;
                              });
        if (run) atomicAdd(&tile_count[last_tile], run);
    }
    block_add_counters(c, s_acc, counters);
}

// One CTA.  tile_count[t] -> tile_offset[t] (exclusive), cursors zeroed, work items and the
// active-tile list built.  tile_count is re-zeroed for the next call.
__global__ void __launch_bounds__(1024)
k_tile_plan(unsigned int* __restrict__ tile_count, unsigned int* __restrict__ tile_offset,
            unsigned int* __restrict__ tile_cursor, int n_tiles, uint4* __restrict__ items, unsigned int max_items,
            unsigned int* __restrict__ active, unsigned long long max_records, TilePlanHeader* __restrict__ hdr,
            uint64_t* counters) {
    __shared__ unsigned int s_warp[3][33];
    __shared__ unsigned int s_carry[3];
    if (threadIdx.x < 3) s_carry[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < n_tiles; start += blockDim.x) {
        const int t = start + threadIdx.x;
        const unsigned int cnt = t < n_tiles ? tile_count[t] : 0u;
        unsigned int v[3] = {cnt, (cnt + kChunk - 1) / kChunk, cnt ? 1u : 0u};
        unsigned int ex[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            unsigned int inc = v[j];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            if (lane == 31) s_warp[j][warp] = inc;
            ex[j] = inc - v[j];
        }
        __syncthreads();
        if (warp < 3) {
            unsigned int w = s_warp[warp][lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned int u = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += u; }
            s_warp[warp][lane] = winc - w;
            if (lane == 31) s_warp[warp][32] = winc;
        }
        __syncthreads();
        const unsigned int off = s_carry[0] + s_warp[0][warp] + ex[0];
        const unsigned int item0 = s_carry[1] + s_warp[1][warp] + ex[1];
        const unsigned int act = s_carry[2] + s_warp[2][warp] + ex[2];
        if (t < n_tiles) {
            tile_offset[t] = off;
            tile_cursor[t] = 0u;
            tile_count[t] = 0u;
            if (cnt) {
                active[act] = (unsigned int)t;
                for (unsigned int j = 0; j < v[1]; ++j) {
                    const unsigned int b = off + j * kChunk;
                    const unsigned int e = min(off + cnt, b + kChunk);
                    if (item0 + j < max_items) items[item0 + j] = make_uint4((unsigned int)t, b, e, 0u);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < 3) s_carry[threadIdx.x] += s_warp[threadIdx.x][32];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        TilePlanHeader h;
        h.total_records = s_carry[0];
        h.n_items = s_carry[1];
        h.n_active = s_carry[2];
        h.work_counter = 0;
        h.resolve_counter = 0;
        h.overflow = (s_carry[0] > max_records || s_carry[1] > max_items) ? 1u : 0u;
        if (h.overflow) { h.n_items = 0; }
        h.pad[0] = h.pad[1] = 0;
        *hdr = h;
        if (counters) atomicAdd(reinterpret_cast<unsigned long long*>(counters) + OCCGRID_C_RECORDS, (unsigned long long)s_carry[0]);
    }
}

__global__ void __launch_bounds__(kTT)
k_tile_scatter(Geom g, TileGeom tg, const uint8_t* __restrict__ pkts, long long n, int stride,
               const int32_t* __restrict__ agent_idx, const double* __restrict__ drift,
               const double* __restrict__ agent_off, int n_agents,
               const unsigned int* __restrict__ tile_offset, unsigned int* __restrict__ tile_cursor,
               const TilePlanHeader* __restrict__ hdr, BeamRec* __restrict__ bins) {
    __shared__ __align__(16) uint8_t s_rec[kTT * kMaxStrideT];
    if (hdr->overflow) return;
    const long long first = (long long)blockIdx.x * kTT;
    const int count = (int)min((long long)kTT, n - first);
    stage_records_t(pkts + (size_t)first * stride, (size_t)count * stride, s_rec);
    __syncthreads();
    if ((int)threadIdx.x >= count) return;
    const long long k = first + threadIdx.x;
    unsigned long long c[OCCGRID_C_OWNED_UPDATES + 1];
    // Records of one packet that fall in the same tile arrive consecutively (the start tile is
    // shared by all four beams): reserve their slots with one atomic per run.
    int run_tile = -1;
    unsigned int run_len = 0;
    BeamRec pend[4];
    auto flush = [&]() {
        if (!run_len) return;
        const unsigned int slot = tile_offset[run_tile] + atomicAdd(&tile_cursor[run_tile], run_len);
        for (unsigned int j = 0; j < run_len; ++j) bins[slot + j] = pend[j];
        run_len = 0;
    };
    for_each_record<false>(g, tg, s_rec + threadIdx.x * stride, k, agent_idx, drift, agent_off, n_agents, c,
                           [&](int tile, int tx, int ty, int s, const Beam& b, bool skip_first) {
                               if (tile != run_tile || run_len == 4) { flush(); run_tile = tile; }
                               BeamRec r;
                               r.lx0 = (short)(b.x0 - g.win_x0 - (tx << kTileShift));
                               r.ly0 = (short)(b.y0 - g.win_y0 - (ty << kTileShift));
                               r.ddx = (short)(b.x1 - b.x0);
                               r.ddy = (short)(b.y1 - b.y0);
                               r.stamp = ((unsigned int)(k * 4 + s + 1) << 1) | (unsigned int)b.hit;
                               r.flags = skip_first ? 1u : 0u;
                               pend[run_len++] = r;
                           });
    flush();
}

__global__ void __launch_bounds__(kTT)
k_tile_raycast(Geom g, TileGeom tg, const uint4* __restrict__ items, TilePlanHeader* __restrict__ hdr,
               const BeamRec* __restrict__ bins, unsigned int* __restrict__ stamps) {
    __shared__ unsigned int s_tile[kTile * kTilePitch];
    __shared__ unsigned int s_item;
    const unsigned int n_items = hdr->n_items;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&hdr->work_counter, 1u);
        for (int i = threadIdx.x; i < kTile * kTilePitch; i += kTT) s_tile[i] = 0u;
        __syncthreads();
        const unsigned int it = s_item;
        if (it >= n_items) break;
        const uint4 item = items[it];
        for (unsigned int r = item.y + threadIdx.x; r < item.z; r += kTT) {
            const BeamRec rec = bins[r];
            int x = rec.lx0, y = rec.ly0;
            const int ddx = rec.ddx, ddy = rec.ddy;
            const int dx = abs(ddx), dy = abs(ddy);
            const int sx = ddx > 0 ? 1 : -1, sy = ddy > 0 ? 1 : -1;     // sx = -1 when x0 == x1 (:163)
            int err = dx - dy;
            const int n = max(dx, dy);
            const unsigned int free_stamp = rec.stamp & ~1u;
            bool was_inside = false;
            int i = 0;
            if (rec.flags & 1u) {               // start cell is overwritten by a later beam of the packet
                if (n == 0) continue;
                const int e2 = 2 * err;
                if (e2 > -dy) { err -= dy; x += sx; }
                if (e2 < dx)  { err += dx; y += sy; }
                i = 1;
            }
            for (; i < n; ++i) {
                const bool in = ((unsigned int)x < (unsigned int)kTile) && ((unsigned int)y < (unsigned int)kTile);
                if (in) { atomicMax(&s_tile[y * kTilePitch + x], free_stamp); was_inside = true; }
                else if (was_inside) break;     // a straight line never re-enters a convex tile
                const int e2 = 2 * err;
                if (e2 > -dy) { err -= dy; x += sx; }
                if (e2 < dx)  { err += dx; y += sy; }
            }
            if (i == n && (rec.stamp & 1u) && ((unsigned int)x < (unsigned int)kTile) && ((unsigned int)y < (unsigned int)kTile))
                atomicMax(&s_tile[y * kTilePitch + x], rec.stamp);
        }
        __syncthreads();
        const int tx = item.x % tg.tiles_x, ty = item.x / tg.tiles_x;
        const int gx0 = tx << kTileShift, gy0 = ty << kTileShift;
        for (int idx = threadIdx.x; idx < kTile * kTile; idx += kTT) {
            const int lx = idx & (kTile - 1), ly = idx >> kTileShift;
            const unsigned int v = s_tile[ly * kTilePitch + lx];
            if (v && gx0 + lx < g.win_w && gy0 + ly < g.win_h)
                atomicMax(&stamps[(size_t)(gy0 + ly) * g.win_w + (gx0 + lx)], v);
        }
    }
}

__global__ void __launch_bounds__(kTT)
k_tile_resolve(Geom g, TileGeom tg, const unsigned int* __restrict__ active, TilePlanHeader* __restrict__ hdr,
               unsigned int* __restrict__ stamps, int8_t* __restrict__ grid) {
    __shared__ unsigned int s_item;
    const unsigned int n_active = hdr->overflow ? 0u : hdr->n_active;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&hdr->resolve_counter, 1u);
        __syncthreads();
        const unsigned int it = s_item;
        if (it >= n_active) break;
        const unsigned int t = active[it];
        const int gx0 = (int)(t % tg.tiles_x) << kTileShift, gy0 = (int)(t / tg.tiles_x) << kTileShift;
        for (int idx = threadIdx.x; idx < kTile * kTile; idx += kTT) {
            const int gx = gx0 + (idx & (kTile - 1)), gy = gy0 + (idx >> kTileShift);
            if (gx < g.win_w && gy < g.win_h) {
                const size_t cidx = (size_t)gy * g.win_w + gx;
                const unsigned int s = stamps[cidx];
                if (s) {
                    grid[cidx] = (s & 1u) ? OCCGRID_CELL_OCCUPIED : OCCGRID_CELL_FREE;
                    stamps[cidx] = 0u;
                }
            }
        }
    }
}

__global__ void k_tile_check(const TilePlanHeader* __restrict__ hdr, int* __restrict__ status_word) {
    if (hdr->overflow) atomicOr(status_word, 1);
}

// ---- host side ----------------------------------------------------------------------------

static int tiles_per_beam_bound(const occgrid_geom* geom) {
    const int reach = (int)ceil(OCC_MAX_DIST_M / geom->res) + 2;
    const int per_axis = reach / kTile + 2;
    return per_axis * per_axis;
}

struct TiledLayout {
    size_t off_stamps, off_count, off_offset, off_cursor, off_active, off_hdr, off_items, off_bins, total;
    unsigned long long max_records;
    unsigned int max_items;
};

static TiledLayout tiled_layout(const occgrid_geom* geom, int64_t max_packets) {
    TiledLayout L;
    const TileGeom tg = tile_geom(geom->win_w, geom->win_h);
    L.max_records = (unsigned long long)max_packets * 4ull * (unsigned long long)tiles_per_beam_bound(geom);
    unsigned long long items = L.max_records / kChunk + (unsigned long long)tg.n_tiles + 1;
    if (items > L.max_records + 1) items = L.max_records + 1;
    L.max_items = (unsigned int)(items > 0xfffffff0ull ? 0xfffffff0ull : items);
    size_t o = 0;
    L.off_stamps = o; o += align_up((size_t)geom->win_w * geom->win_h * 4, 256);
    L.off_count = o;  o += align_up((size_t)(tg.n_tiles + 1) * 4, 256);
    L.off_offset = o; o += align_up((size_t)(tg.n_tiles + 1) * 4, 256);
    L.off_cursor = o; o += align_up((size_t)(tg.n_tiles + 1) * 4, 256);
    L.off_active = o; o += align_up((size_t)(tg.n_tiles + 1) * 4, 256);
    L.off_hdr = o;    o += 256;
    L.off_items = o;  o += align_up((size_t)L.max_items * sizeof(uint4), 256);
    L.off_bins = o;   o += align_up((size_t)L.max_records * sizeof(BeamRec), 256);
    L.total = o;
    return L;
}

bool tiled_supported(const occgrid_geom* geom) {
    const TileGeom tg = tile_geom(geom->win_w, geom->win_h);
    const int reach = (int)ceil(OCC_MAX_DIST_M / geom->res) + 2;
    return reach <= 30000 && tg.n_tiles <= (1 << 24);
}

size_t tiled_workspace_bytes(const occgrid_geom* geom, int64_t max_packets) {
    return tiled_layout(geom, max_packets < 1 ? 1 : max_packets).total;
}

int integrate_packets_tiled(const occgrid_geom* geom, const uint8_t* d_packets, int64_t n, int stride,
                            const int32_t* d_agent_idx, const double* d_drift, const double* d_agent_off,
                            int n_agents, int8_t* d_grid, void* d_ws, size_t ws_bytes, uint64_t* d_counters,
                            cudaStream_t st) {
    const TiledLayout L = tiled_layout(geom, n);
    if (ws_bytes < L.total) {
        set_last_error("workspace %zu B < %zu B needed by TILED for %lld records", ws_bytes, L.total, (long long)n);
        return OCCGRID_E_WORKSPACE;
    }
    if (L.max_records >= 0xffffffffull) { set_last_error("TILED: batch too large (record offsets are 32-bit)"); return OCCGRID_E_ARG; }
    char* ws = reinterpret_cast<char*>(d_ws);
    unsigned int* stamps = reinterpret_cast<unsigned int*>(ws + L.off_stamps);
    unsigned int* tile_count = reinterpret_cast<unsigned int*>(ws + L.off_count);
    unsigned int* tile_offset = reinterpret_cast<unsigned int*>(ws + L.off_offset);
    unsigned int* tile_cursor = reinterpret_cast<unsigned int*>(ws + L.off_cursor);
    unsigned int* active = reinterpret_cast<unsigned int*>(ws + L.off_active);
    TilePlanHeader* hdr = reinterpret_cast<TilePlanHeader*>(ws + L.off_hdr);
    uint4* items = reinterpret_cast<uint4*>(ws + L.off_items);
    BeamRec* bins = reinterpret_cast<BeamRec*>(ws + L.off_bins);
    const Geom g = to_geom(geom);
    const TileGeom tg = tile_geom(geom->win_w, geom->win_h);
    const unsigned int blocks = (unsigned int)((n + kTT - 1) / kTT);
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    {
        ProfileScope ps(K_TILE_COUNT, st);
        k_tile_count<<<blocks, kTT, 0, st>>>(g, tg, d_packets, n, stride, d_agent_idx, d_drift, d_agent_off, n_agents,
                                            tile_count, d_counters);
    }
    {
        ProfileScope ps(K_TILE_SCAN, st);
        k_tile_plan<<<1, 1024, 0, st>>>(tile_count, tile_offset, tile_cursor, tg.n_tiles, items, L.max_items, active,
                                        L.max_records, hdr, d_counters);
    }
    {
        ProfileScope ps(K_TILE_SCATTER, st);
        k_tile_scatter<<<blocks, kTT, 0, st>>>(g, tg, d_packets, n, stride, d_agent_idx, d_drift, d_agent_off, n_agents,
                                              tile_offset, tile_cursor, hdr, bins);
    }
    {
        ProfileScope ps(K_TILE_RAYCAST, st);
        k_tile_raycast<<<sms * 8, kTT, 0, st>>>(g, tg, items, hdr, bins, stamps);
    }
    {
        ProfileScope ps(K_TILE_RESOLVE, st);
        k_tile_resolve<<<sms * 4, kTT, 0, st>>>(g, tg, active, hdr, stamps, d_grid);
    }
    OCC_CUDA_TRY(cudaGetLastError());
    return OCCGRID_OK;
}

}  // namespace occ
