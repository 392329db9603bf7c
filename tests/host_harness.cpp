// CPU harness around the host+device headers of the product (beam_expand.cuh, sincos_dd.cuh):
// lets the CPU test-suite check the packet expansion logic — including the screened fast path
// and its exact re-evaluation — without a GPU.  Built by tests/test_host_expand.py with
//   g++ -O2 -ffp-contract=off -shared -fPIC
// Test infrastructure only; nothing in the product links against it.
#include <cstdint>
#include <cmath>
#include <cstring>

#include "../distributed-multi-agent-slam-swarm-robotics-system_b200/csrc/beam_expand.cuh"

namespace {

// A deliberately sloppy sincos: glibc's result moved by up to 2 ulp, direction and amount
// chosen from the argument's bits (mode 1), or always +2 / -2 ulp (modes 2 / 3).
struct PerturbedSinCos {
    int mode;
    static double bump(double v, int k) {
        for (int i = 0; i < (k < 0 ? -k : k); ++i) v = std::nextafter(v, k > 0 ? INFINITY : -INFINITY);
        return v;
    }
    void operator()(double a, double* s, double* c) const {
        double sv = std::sin(a), cv = std::cos(a);
        if (mode == 1) {
            uint64_t bits;
            std::memcpy(&bits, &a, 8);
            bits *= 0x9E3779B97F4A7C15ull;
            sv = bump(sv, (int)((bits >> 60) % 5) - 2);
            cv = bump(cv, (int)((bits >> 56) % 5) - 2);
        } else if (mode == 2) { sv = bump(sv, 2); cv = bump(cv, 2); }
        else if (mode == 3) { sv = bump(sv, -2); cv = bump(cv, -2); }
        *s = sv;
        *c = cv;
    }
};

}  // namespace

extern "C" {

int hh_sincos_dd(double x, double* s, double* c) { return occ::sincos_dd(x, s, c) ? 1 : 0; }

// out: n x 4 x 7 int32 (x0, y0, x1, y1, hit, valid, slow); status: n (PKT_*)
void hh_expand_packets(const uint8_t* pkts, int64_t n, int stride, const int32_t* agent_idx, const double* drift,
                       const double* agent_off, int n_agents, double ox, double oy, double res, int mode,
                       int32_t* out, int32_t* status) {
    occ::Geom g;
    g.ox = ox; g.oy = oy; g.res = res; g.inv_res = 1.0 / res;
    g.size_x = g.size_y = 1 << 30;
    g.win_x0 = g.win_y0 = 0; g.win_w = g.win_h = 1 << 30;
    PerturbedSinCos fsc{mode};
    for (int64_t k = 0; k < n; ++k) {
        double rx, ry, ryaw;
        float dist[4];
        int st = occ::decode_packet(pkts + k * stride, k, agent_idx, drift, agent_off, n_agents, &rx, &ry, &ryaw, dist);
        status[k] = st;
        int32_t* o = out + k * 28;
        std::memset(o, 0, 28 * sizeof(int32_t));
        if (st != occ::PKT_OK) continue;
        occ::Beam b[4];
        occ::expand_packet(g, rx, ry, ryaw, dist, fsc, b);
        for (int s = 0; s < 4; ++s) {
            o[s * 7 + 0] = b[s].x0; o[s * 7 + 1] = b[s].y0; o[s * 7 + 2] = b[s].x1; o[s * 7 + 3] = b[s].y1;
            o[s * 7 + 4] = b[s].hit; o[s * 7 + 5] = b[s].valid; o[s * 7 + 6] = b[s].slow;
        }
    }
}

int64_t hh_bresenham(int x0, int y0, int x1, int y1, int32_t* xs, int32_t* ys, int64_t cap) {
    int64_t n = 0;
    occ::bresenham_walk(x0, y0, x1, y1, [&](int x, int y, bool) {
        if (n < cap) { xs[n] = x; ys[n] = y; }
        ++n;
    });
    return n;
}

}  // extern "C"
