// plan.cu — plan-time graph building on the device: exact k-nearest-neighbour search in 2-D on a uniform
// cell grid.  It replaces the host k-d tree behind the kNN connectivity (transforms/connect.py:58 -> torch_cluster.knn_graph)
// and the interpolation lists (transforms/interpolate.py:125 -> torch_cluster.knn) when a mesh is built on the GPU; the rest
// of the builders (grid clustering, the closed-form angle lists) are torch device ops in graphs4cfd_b200/mesh.py.
//
// One thread per query point.  Points are pre-sorted by cell (host side of the wrapper: torch.sort of the cell ids), a cell's
// points are sorted_idx[cell_start[c] .. cell_start[c + 1]).  The thread scans rings of cells around its own cell, keeps the
// k best (distance^2, index) pairs in an insertion-sorted list, and stops when the k-th distance is not larger than the
// distance to the border of the scanned window (no unseen point can be closer).  Distances are formed in double without
// contraction, like the host k-d tree (scipy cKDTree) does, so both orderings agree bit for bit; ties go to the lower index.
#include "common.cuh"

namespace g4c {

constexpr int KNN_MAX_K = 16;

__global__ void knn_grid_kernel(const G4cKnnDesc d) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= d.n_queries) return;
    const double qx = (double)d.query[2 * q], qy = (double)d.query[2 * q + 1];
    const double c = (double)d.cell, x0 = (double)d.x0, y0 = (double)d.y0;
    int cx = (int)floor((qx - x0) / c), cy = (int)floor((qy - y0) / c);
    cx = min(max(cx, 0), d.gx - 1);
    cy = min(max(cy, 0), d.gy - 1);
    double bd[KNN_MAX_K];
    int bi[KNN_MAX_K];
    int found = 0;
    const int k = d.k;
    const int self = d.exclude_self ? (int)q : -1;
    const int rmax = max(d.gx, d.gy);
    for (int r = 0; r <= rmax; ++r) {
        for (int yy = cy - r; yy <= cy + r; ++yy) {
            if (yy < 0 || yy >= d.gy) continue;
            const bool edge_row = (yy == cy - r) || (yy == cy + r);
            for (int xx = cx - r; xx <= cx + r; xx += (edge_row ? 1 : 2 * r)) {      // the ring only: full rows at the top / bottom, two cells else
                if (xx >= 0 && xx < d.gx) {
                    const int cell = yy * d.gx + xx;
                    for (int s = d.cell_start[cell]; s < d.cell_start[cell + 1]; ++s) {
                        const int p = d.sorted_idx[s];
                        if (p == self) continue;
                        const double dx = __dsub_rn((double)d.pos[2 * p], qx), dy = __dsub_rn((double)d.pos[2 * p + 1], qy);
                        const double dist = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                        if (found < k || dist < bd[found - 1] || (dist == bd[found - 1] && p < bi[found - 1])) {
                            int j = found < k ? found : k - 1;
                            while (j > 0 && (bd[j - 1] > dist || (bd[j - 1] == dist && bi[j - 1] > p))) {
                                bd[j] = bd[j - 1];
                                bi[j] = bi[j - 1];
                                --j;
                            }
                            bd[j] = dist;
                            bi[j] = p;
                            if (found < k) ++found;
                        }
                    }
                }
                if (r == 0) break;
            }
        }
        if (found == k) {
            // distance from the query to the border of the scanned window (sides outside the grid do not bound anything)
            double bound = 1e300;
            if (cx - r > 0) bound = fmin(bound, qx - (x0 + (cx - r) * c));
            if (cx + r < d.gx - 1) bound = fmin(bound, (x0 + (cx + r + 1) * c) - qx);
            if (cy - r > 0) bound = fmin(bound, qy - (y0 + (cy - r) * c));
            if (cy + r < d.gy - 1) bound = fmin(bound, (y0 + (cy + r + 1) * c) - qy);
            if (bound >= 1e300 || (bound > 0 && bd[k - 1] <= bound * bound)) break;
        }
    }
    for (int j = 0; j < k; ++j) d.nbr[q * k + j] = j < found ? bi[j] : -1;
}

int knn_launch(const G4cKnnDesc& d, cudaStream_t st) {
    if (d.n_queries == 0) return G4C_OK;
    const int64_t blocks = (d.n_queries + 127) / 128;
    knn_grid_kernel<<<(unsigned)blocks, 128, 0, st>>>(d);
    count_launch();
    return check_launch("knn_grid_kernel");
}

}  // namespace g4c
