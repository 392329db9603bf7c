// Occupancy overlay of the reference's dashboard as a headless RGB image:
// MapRenderer._draw_occupancy, server_nodes/dual_bot_mapper.py:492-527, with world_to_screen
// (:404-408) and grid_to_world (:127-131).  The reference paints, every frame and in Python, one
// rectangle per visible cell that is neither UNKNOWN nor OCCUPIED (:513-520) in CELL_COLOR_FREE;
// all rectangles have one colour, so the order of the cells does not matter and every visible
// cell gets its own thread.  PyGame's primitives are restated from their documented behaviour:
// set_at outside the surface has no effect, draw.rect fills the rectangle clipped to the surface.
#include "common.cuh"

namespace occ {

constexpr int kRT2 = 256;

struct RenderView {
    double ox, oy, res, scale, offset_x, offset_y;
    int size_x, size_y, width, height;
    int gx_min, gx_max, gy_min, gy_max, cell_px;
    unsigned char bg[3], fg[3];
};

__global__ void __launch_bounds__(kRT2)
k_render_fill(unsigned char* __restrict__ rgb, long long n_pixels, unsigned char r, unsigned char g, unsigned char b) {
    // 12 bytes = 4 pixels per step when the buffer is 4-byte aligned
    const unsigned int w0 = r | (g << 8) | (b << 16) | (r << 24), w1 = g | (b << 8) | (r << 16) | (g << 24), w2 = b | (r << 8) | (g << 16) | (b << 24);
    const long long quads = ((reinterpret_cast<uintptr_t>(rgb) & 3) == 0) ? n_pixels / 4 : 0;
    unsigned int* w = reinterpret_cast<unsigned int*>(rgb);
    for (long long i = (long long)blockIdx.x * kRT2 + threadIdx.x; i < quads; i += (long long)gridDim.x * kRT2) {
        w[3 * i] = w0; w[3 * i + 1] = w1; w[3 * i + 2] = w2;
    }
    for (long long i = quads * 4 + (long long)blockIdx.x * kRT2 + threadIdx.x; i < n_pixels; i += (long long)gridDim.x * kRT2) {
        rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
    }
}

__global__ void __launch_bounds__(kRT2)
k_render_cells(const int8_t* __restrict__ grid, RenderView v, unsigned char* __restrict__ rgb) {
    const int nx = v.gx_max - v.gx_min;
    const long long n = (long long)nx * (v.gy_max - v.gy_min);
    for (long long i = (long long)blockIdx.x * kRT2 + threadIdx.x; i < n; i += (long long)gridDim.x * kRT2) {
        const int gy = v.gy_min + (int)(i / nx), gx = v.gx_min + (int)(i - (long long)(i / nx) * nx);
        const int8_t val = grid[(size_t)gy * v.size_x + gx];
        if (val == OCCGRID_CELL_UNKNOWN || val == OCCGRID_CELL_OCCUPIED) continue;          // :513-520
        const double wx = OCC_DADD(v.ox, OCC_DMUL(OCC_DADD((double)gx, 0.5), v.res));         // :129-130
        const double wy = OCC_DADD(v.oy, OCC_DMUL(OCC_DADD((double)gy, 0.5), v.res));
        const int sx = __double2int_rz(OCC_DADD(v.offset_x, OCC_DMUL(wx, v.scale)));          // :406
        const int sy = __double2int_rz(OCC_DADD(v.offset_y, -OCC_DMUL(wy, v.scale)));         // :407 (y is flipped)
        int x0 = sx, y0 = sy, side = 1;
        if (v.cell_px > 2) { x0 = sx - v.cell_px / 2; y0 = sy - v.cell_px / 2; side = v.cell_px; }   // :523-527
        const int xa = max(0, x0), xb = min(v.width, x0 + side), ya = max(0, y0), yb = min(v.height, y0 + side);
        for (int y = ya; y < yb; ++y) {
            unsigned char* row = rgb + ((size_t)y * v.width + xa) * 3;
            for (int x = xa; x < xb; ++x, row += 3) { row[0] = v.fg[0]; row[1] = v.fg[1]; row[2] = v.fg[2]; }
        }
    }
}

}  // namespace occ

using namespace occ;

extern "C" {

int occgrid_render_overlay(const int8_t* d_grid, int32_t size_x, int32_t size_y, double ox, double oy, double res,
                           double scale, double offset_x, double offset_y, int32_t width, int32_t height,
                           const uint8_t* bg_rgb_host, const uint8_t* fg_rgb_host, uint8_t* d_rgb, void* stream) {
    if (!d_grid || !d_rgb || !bg_rgb_host || !fg_rgb_host || size_x <= 0 || size_y <= 0 || width <= 0 || height <= 0 ||
        !(res > 0.0) || !(scale > 0.0)) {
        set_last_error("occgrid_render_overlay: bad arguments");
        return OCCGRID_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    RenderView v;
    v.ox = ox; v.oy = oy; v.res = res; v.scale = scale; v.offset_x = offset_x; v.offset_y = offset_y;
    v.size_x = size_x; v.size_y = size_y; v.width = width; v.height = height;
    for (int i = 0; i < 3; ++i) { v.bg[i] = bg_rgb_host[i]; v.fg[i] = fg_rgb_host[i]; }
    const double cp = res * scale;
    v.cell_px = cp >= 2147483647.0 ? 2147483647 : ((int)cp > 1 ? (int)cp : 1);                 // :494
    // visible world bounds and cell ranges (:500-508); int() truncates toward zero like the C cast
    const double world_left = -offset_x / scale, world_right = ((double)width - offset_x) / scale;
    const double world_top = offset_y / scale, world_bottom = -((double)height - offset_y) / scale;
    auto cell = [](double q) { return q >= 2147483000.0 ? 2147483000 : (q <= -2147483000.0 ? -2147483000 : (int)q); };
    v.gx_min = cell((world_left - ox) / res) - 1;  if (v.gx_min < 0) v.gx_min = 0;
    v.gx_max = cell((world_right - ox) / res) + 1; if (v.gx_max > size_x) v.gx_max = size_x;
    v.gy_min = cell((world_bottom - oy) / res) - 1; if (v.gy_min < 0) v.gy_min = 0;
    v.gy_max = cell((world_top - oy) / res) + 1;    if (v.gy_max > size_y) v.gy_max = size_y;
    ProfileScope ps(K_RENDER, st, 2);
    const long long n_pixels = (long long)width * height;
    long long fb = (n_pixels / 4 + kRT2 - 1) / kRT2;
    if (fb < 1) fb = 1;
    if (fb > device_sm_count() * 16) fb = device_sm_count() * 16;
    k_render_fill<<<(int)fb, kRT2, 0, st>>>(d_rgb, n_pixels, v.bg[0], v.bg[1], v.bg[2]);
    if (v.cell_px >= 2 && v.gx_max > v.gx_min && v.gy_max > v.gy_min) {                        // :495-496
        long long cb = ((long long)(v.gx_max - v.gx_min) * (v.gy_max - v.gy_min) + kRT2 - 1) / kRT2;
        if (cb > device_sm_count() * 16) cb = device_sm_count() * 16;
        k_render_cells<<<(int)cb, kRT2, 0, st>>>(d_grid, v, d_rgb);
    }
    OCC_CUDA_TRY(cudaGetLastError());
    return OCCGRID_OK;
}

}  // extern "C"
