// Host-side producer of the per-packet drift table.
//
// Reference: the sequential half of main()'s ingest loop, server_nodes/dual_bot_mapper.py:826-857
// and :908-914, with PoseGraphSLAM.add_pose / _check_closure (:261-322):
//   for every datagram, in arrival order:
//     decode (v2 42 B / v1 41 B, :828-838), drop bad magic or agent_id not in {1, 2} (:840-843)
//     rx += separation for agent 2 (:851-852); (cdx, cdy) = drift_correction[agent]  <- the table entry
//     rx += cdx; ry += cdy (:856-857)
//     closure, dx, dy = slam.add_pose(rx, ry, ryaw, agent, landmark, t) (:908-909)
//     if closure: drift_correction[agent] += (dx, dy) (:910-914)
//   add_pose: a pose with a landmark is matched against every EARLIER landmark of the same type
//   that is >= 30 poses older, unless this agent closed a loop < 30 poses ago; the first one
//   (insertion order) closer than 0.60 m closes the loop with half the error as correction.
// The chain is inherently sequential (every packet's pose depends on the closures before it)
// and it is pure host work in the reference; this is the same loop in C++ so that feeding the
// device path does not cost seconds of Python per batch.  The reference scans the whole landmark
// list per landmark pose (quadratic); here landmarks sit in a spatial hash per type (cell =
// closure radius) and the match with the smallest insertion index among the 3x3 neighbouring
// cells is taken — the same landmark the linear scan returns first.  All arithmetic is the
// reference's fp64 sequence (compiled without FMA contraction): results are bit-identical.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <unordered_map>
#include <vector>

#include "../../include/occgrid_b200.h"

namespace occ {
void set_last_error(const char* fmt, ...);
}

namespace {

constexpr double kClosureRadius = 0.60;      // :97
constexpr long long kMinPosesBetween = 30;   // :98
constexpr double kClosureCorrection = 0.5;   // :99

struct Landmark { double x, y; int type; long long node; };
struct Closure { long long lm_node, node; double dx, dy; int agent; };

struct CellKey {
    long long cx, cy; int type;
    bool operator==(const CellKey& o) const { return cx == o.cx && cy == o.cy && type == o.type; }
};
struct CellHash {
    size_t operator()(const CellKey& k) const {
        uint64_t h = (uint64_t)k.cx * 0x9E3779B97F4A7C15ull;
        h ^= (uint64_t)k.cy * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
        h ^= (uint64_t)k.type * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
        return (size_t)h;
    }
};

}  // namespace

struct occgrid_slam {
    long long n_nodes = 0;
    std::vector<Landmark> landmarks;                                     // insertion order, as the reference's list
    std::unordered_map<CellKey, std::vector<int>, CellHash> cells;       // landmark ids per (cell, type), ascending
    std::vector<Closure> closures;
    long long last_closure[3] = {0, -kMinPosesBetween, -kMinPosesBetween};   // per agent id (:271)
    double drift[3][2] = {{0, 0}, {0, 0}, {0, 0}};                       // drift_correction (:782)
};

static inline long long cell_of(double v) { return (long long)std::floor(v / kClosureRadius); }

// PoseGraphSLAM.add_pose (:261-280) + _check_closure (:282-322)
static bool slam_add_pose(occgrid_slam* s, double x, double y, int agent, int lm_type, double* dx, double* dy) {
    const long long idx = s->n_nodes++;
    if (lm_type == 0) return false;                                      // LM_NONE
    bool closed = false;
    const bool finite_pose = std::isfinite(x) && std::isfinite(y);
    const bool recent = idx - s->last_closure[agent] < kMinPosesBetween;
    if (!recent && finite_pose) {
        int best = -1;
        const long long cx = cell_of(x), cy = cell_of(y);
        for (long long oy = -1; oy <= 1; ++oy)
            for (long long ox = -1; ox <= 1; ++ox) {
                auto it = s->cells.find(CellKey{cx + ox, cy + oy, lm_type});
                if (it == s->cells.end()) continue;
                for (int id : it->second) {                              // ascending insertion index
                    if (best >= 0 && id > best) break;
                    const Landmark& lm = s->landmarks[id];
                    if (idx - lm.node < kMinPosesBetween) break;         // later ones are younger still
                    const double ex = x - lm.x, ey = y - lm.y;
                    const double dist = std::sqrt(ex * ex + ey * ey);    // math.sqrt((dx)**2 + (dy)**2)
                    if (dist < kClosureRadius) { best = id; break; }
                }
            }
        if (best >= 0) {
            const Landmark& lm = s->landmarks[best];
            *dx = (lm.x - x) * kClosureCorrection;
            *dy = (lm.y - y) * kClosureCorrection;
            s->closures.push_back(Closure{lm.node, idx, *dx, *dy, agent});
            s->last_closure[agent] = idx;
            closed = true;
        }
    }
    const int id = (int)s->landmarks.size();
    s->landmarks.push_back(Landmark{x, y, lm_type, idx});
    if (finite_pose) s->cells[CellKey{cell_of(x), cell_of(y), lm_type}].push_back(id);
    return closed;
}

static inline float load_f32(const uint8_t* p) { float f; std::memcpy(&f, p, 4); return f; }

extern "C" {

occgrid_slam* occgrid_slam_create(void) { return new (std::nothrow) occgrid_slam(); }

void occgrid_slam_destroy(occgrid_slam* s) { delete s; }

int occgrid_slam_drift_table(occgrid_slam* s, const uint8_t* packets, int64_t n, int32_t rec_stride, int32_t rec_size,
                             const int32_t* rec_sizes, double separation, double* drift_out) {
    if (!s || (!packets && n > 0) || n < 0 || rec_stride < 41 || !drift_out || (!rec_sizes && rec_size != 41 && rec_size != 42)) {
        occ::set_last_error("occgrid_slam_drift_table: bad arguments");
        return OCCGRID_E_ARG;
    }
    for (int64_t k = 0; k < n; ++k) {
        const uint8_t* p = packets + (size_t)k * (size_t)rec_stride;
        drift_out[2 * k] = 0.0;
        drift_out[2 * k + 1] = 0.0;
        const int size = rec_sizes ? rec_sizes[k] : rec_size;
        if (size != 42 && size != 41) continue;                          // :836-838
        if (size > rec_stride) continue;
        if (std::memcmp(p, "QSRL", 4) != 0) continue;                    // :840-841
        const int agent = p[4];
        if (agent != 1 && agent != 2) continue;                          // :842-843
        double rx = (double)load_f32(p + 5), ry = (double)load_f32(p + 9);
        const double ryaw = (double)load_f32(p + 13);
        const int lm = size == 42 ? p[41] : 0;                           // v1 has no landmark byte (:832-835)
        if (agent == 2) rx += separation;                                // :851-852
        const double cdx = s->drift[agent][0], cdy = s->drift[agent][1]; // :855
        drift_out[2 * k] = cdx;
        drift_out[2 * k + 1] = cdy;
        rx += cdx;                                                       // :856
        ry += cdy;                                                       // :857
        if (!(std::isfinite(rx) && std::isfinite(ry) && std::isfinite(ryaw))) continue;   // int(nan) would crash the reference (:123)
        double dx = 0.0, dy = 0.0;
        if (slam_add_pose(s, rx, ry, agent, lm, &dx, &dy)) {             // :908-914
            s->drift[agent][0] += dx;
            s->drift[agent][1] += dy;
        }
    }
    return OCCGRID_OK;
}

int occgrid_slam_counts(const occgrid_slam* s, int64_t* n_nodes, int64_t* n_landmarks, int64_t* n_closures) {
    if (!s) { occ::set_last_error("occgrid_slam_counts: NULL"); return OCCGRID_E_ARG; }
    if (n_nodes) *n_nodes = s->n_nodes;
    if (n_landmarks) *n_landmarks = (int64_t)s->landmarks.size();
    if (n_closures) *n_closures = (int64_t)s->closures.size();
    return OCCGRID_OK;
}

int occgrid_slam_closures(const occgrid_slam* s, int64_t capacity, int64_t* lm_node, int64_t* node, double* corr_xy, int32_t* agent) {
    if (!s || capacity < 0) { occ::set_last_error("occgrid_slam_closures: bad arguments"); return OCCGRID_E_ARG; }
    const int64_t n = (int64_t)s->closures.size() < capacity ? (int64_t)s->closures.size() : capacity;
    for (int64_t i = 0; i < n; ++i) {
        const Closure& c = s->closures[(size_t)i];
        if (lm_node) lm_node[i] = c.lm_node;
        if (node) node[i] = c.node;
        if (corr_xy) { corr_xy[2 * i] = c.dx; corr_xy[2 * i + 1] = c.dy; }
        if (agent) agent[i] = c.agent;
    }
    return OCCGRID_OK;
}

// PoseGraphSLAM.get_correction_for_agent (:324-332): the closures of that agent summed in order.
int occgrid_slam_correction_for_agent(const occgrid_slam* s, int32_t agent_id, double* out_xy) {
    if (!s || !out_xy) { occ::set_last_error("occgrid_slam_correction_for_agent: bad arguments"); return OCCGRID_E_ARG; }
    double tx = 0.0, ty = 0.0;
    for (const Closure& c : s->closures)
        if (c.agent == agent_id) { tx += c.dx; ty += c.dy; }
    out_xy[0] = tx; out_xy[1] = ty;
    return OCCGRID_OK;
}

}  // extern "C"
