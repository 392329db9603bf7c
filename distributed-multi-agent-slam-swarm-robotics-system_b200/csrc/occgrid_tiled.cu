// Strategy TILED — placeholder until the shared-memory tile kernels land.
#include "common.cuh"

namespace occ {
bool tiled_supported(const occgrid_geom*) { return false; }
size_t tiled_workspace_bytes(const occgrid_geom*, int64_t) { return 0; }
int integrate_packets_tiled(const occgrid_geom*, const uint8_t*, int64_t, int, const int32_t*, const double*,
                            const double*, int, int8_t*, void*, size_t, uint64_t*, cudaStream_t) {
    set_last_error("TILED strategy not built");
    return OCCGRID_E_ARG;
}
}  // namespace occ
