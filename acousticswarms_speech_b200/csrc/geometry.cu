// geometry.cu -- host-side hypercube (Grid_cluster) table build.
//
// Reference: SRP_PHAT.Map_3D_TDoA / search_cluster (sep/Traditional_SP/SRP_Prunning.py:277-344): voxels
// whose quantised TDoA vectors are identical and that are 26-connected form one cluster; clusters are
// numbered in the order the C-order scan (ix, iy, iz) meets their first voxel, members are listed in
// the reference's breadth-first order (FIFO queue, neighbours visited dx, dy, dz ascending).  The
// reference spends ~29 s here in Python (a fresh visited volume per cluster); this is the same walk in
// C++ (milliseconds).  Member order matters because the steering position is the mean of the member
// positions accumulated in that order (:90-91, :340).
#include <string.h>

#include <vector>

#include "common.cuh"

extern "C" int asw_geometry_cluster(const int64_t* offsets, const uint8_t* valid, int Lx, int Ly, int Lz, int D,
                                    int32_t* label, int32_t* order, int32_t* cluster_start, int32_t* n_clusters) {
    if (!offsets || !valid || !label || !order || !cluster_start || !n_clusters || Lx < 1 || Ly < 1 || Lz < 1 ||
        D < 1) {
        asw::set_error("asw_geometry_cluster: null argument or empty volume");
        return ASW_ERR_ARG;
    }
    const size_t nvox = (size_t)Lx * Ly * Lz;
    for (size_t i = 0; i < nvox; ++i) label[i] = -1;
    std::vector<int32_t> queue;
    queue.reserve(1024);
    int32_t ncl = 0;
    size_t n_out = 0;
    auto flat = [=](int x, int y, int z) { return ((size_t)x * Ly + y) * Lz + z; };
    for (int ix = 0; ix < Lx; ++ix)
        for (int iy = 0; iy < Ly; ++iy)
            for (int iz = 0; iz < Lz; ++iz) {
                const size_t s = flat(ix, iy, iz);
                if (!valid[s] || label[s] >= 0) continue;
                const int64_t* base = offsets + s * D;
                cluster_start[ncl] = (int32_t)n_out;
                label[s] = ncl;
                order[n_out++] = (int32_t)s;
                queue.clear();
                queue.push_back((int32_t)s);
                for (size_t head = 0; head < queue.size(); ++head) {
                    const int32_t c = queue[head];
                    const int cz = c % Lz, cy = (c / Lz) % Ly, cx = c / (Lz * Ly);
                    for (int nx = cx - 1; nx <= cx + 1; ++nx) {
                        if (nx < 0 || nx >= Lx) continue;
                        for (int ny = cy - 1; ny <= cy + 1; ++ny) {
                            if (ny < 0 || ny >= Ly) continue;
                            for (int nz = cz - 1; nz <= cz + 1; ++nz) {
                                if (nz < 0 || nz >= Lz) continue;
                                const size_t t = flat(nx, ny, nz);
                                if (!valid[t] || label[t] >= 0) continue;
                                if (memcmp(base, offsets + t * D, sizeof(int64_t) * D) != 0) continue;
                                label[t] = ncl;
                                order[n_out++] = (int32_t)t;
                                queue.push_back((int32_t)t);
                            }
                        }
                    }
                }
                ++ncl;
            }
    cluster_start[ncl] = (int32_t)n_out;
    *n_clusters = ncl;
    return ASW_OK;
}
