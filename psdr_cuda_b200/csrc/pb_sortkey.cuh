// psdr-b200: key of the counting sort that regroups a wavefront's rays before traversal (pb_sort.cu). Shared with k_shade, which knows
// every ray it emits and leaves the 2-byte key next to it, so that the histogram pass does not read the 32-byte rays again.
#pragma once
#include "pb_trace.cuh"

namespace pb {

constexpr int kSortBins = 4096;   // 12-bit keys; key kSortBins = inactive lane (compacted away)
enum { SORT_CELL_OCTANT = 5, SORT_DIR_CELL = 0, SORT_DIRECTION = 7 };

PB_D int morton3(int x, int y, int z, int bits) {
    int m = 0;
    for (int k = 0; k < bits; ++k) m |= (((x >> k) & 1) << (3 * k)) | (((y >> k) & 1) << (3 * k + 1)) | (((z >> k) & 1) << (3 * k + 2));
    return m;
}

// SORT_CELL_OCTANT (wavefront rays, which start on surfaces): 9 bits origin cell (8x8x8, Morton) x 3 bits direction octant, cell-major.
// SORT_DIRECTION (rays that share an origin: camera rays of the edge terms): 12 bits direction (octahedral 64x64, Morton).
// SORT_DIR_CELL (round 1, kept for the A/B): 6 bits direction (octahedral 8x8) x 6 bits origin cell (4x4x4).
PB_D int sort_key(float3 o, float tmax, float3 d, float3 lo, float3 inv_ext, int mode) {
    if (!(tmax > 0.f)) return kSortBins;
    if (mode == SORT_DIRECTION) {
        const float inv = 1.f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z));
        float px = d.x * inv, py = d.y * inv;
        if (d.z < 0.f) {
            const float qx = (1.f - fabsf(py)) * (px >= 0.f ? 1.f : -1.f), qy = (1.f - fabsf(px)) * (py >= 0.f ? 1.f : -1.f);
            px = qx; py = qy;
        }
        const int ux = min(63, max(0, (int)((px * .5f + .5f) * 64.f))), uy = min(63, max(0, (int)((py * .5f + .5f) * 64.f)));
        int m = 0;
#pragma unroll
        for (int b = 0; b < 6; ++b) m |= (((ux >> b) & 1) << (2 * b)) | (((uy >> b) & 1) << (2 * b + 1));
        return m;
    }
    if (mode == SORT_DIR_CELL) {
        const int cx = min(3, max(0, (int)((o.x - lo.x) * inv_ext.x * 4.f)));
        const int cy = min(3, max(0, (int)((o.y - lo.y) * inv_ext.y * 4.f)));
        const int cz = min(3, max(0, (int)((o.z - lo.z) * inv_ext.z * 4.f)));
        return (direction_bin(d) << 6) | morton3(cx, cy, cz, 2);
    }
    const int cx = min(7, max(0, (int)((o.x - lo.x) * inv_ext.x * 8.f)));
    const int cy = min(7, max(0, (int)((o.y - lo.y) * inv_ext.y * 8.f)));
    const int cz = min(7, max(0, (int)((o.z - lo.z) * inv_ext.z * 8.f)));
    return morton3(cx, cy, cz, 3) << 3 | (((d.x < 0.f) ? 4 : 0) | ((d.y < 0.f) ? 2 : 0) | ((d.z < 0.f) ? 1 : 0));
}

}  // namespace pb
