/*
 * arp_rings.cu -- ring -> residue assignment.
 *
 * Replaces the search of InteractionComplex._assign_aromatic_rings_to_residues
 * (interactions.py:1453-1492): for every ring centroid, the closest atom of the structure within
 * `radius` (3.0 A there) --
 *     atoms_near_ring = NeighborSearch(s_atoms).search(ring_centroid, 3.0)       double, d2 <= r*r
 *     distance = np.linalg.norm(nearby_atom.coord - ring_centroid)               float64 (f32 - f64)
 *     strict `<` keeps the first minimum; ties go to the lowest atom index here (the reference's
 *     order is the KD-tree traversal)
 * One warp per ring, lanes stride over the atoms (coordinates staged through shared memory in
 * tiles shared by the block's rings), warp argmin by (distance, index).
 */
#include "arp_ctx.cuh"

#define RING_WARPS 8
#define RING_TILE  1024

__global__ void __launch_bounds__(RING_WARPS * 32) k_ring_nearest(int n_rings, int n_atoms, const double* __restrict__ center,
                                                                  const float* __restrict__ xyz, double r2, int blas_fma,
                                                                  int32_t* __restrict__ atom_out, double* __restrict__ dist_out)
{
    __shared__ float s_xyz[RING_TILE * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ring = blockIdx.x * RING_WARPS + warp;
    const bool live = ring < n_rings;
    double cx = 0.0, cy = 0.0, cz = 0.0;
    if (live) { cx = center[3 * (size_t)ring]; cy = center[3 * (size_t)ring + 1]; cz = center[3 * (size_t)ring + 2]; }
    double best = 0.0;
    int best_i = -1;
    for (int t0 = 0; t0 < n_atoms; t0 += RING_TILE) {
        const int m = min(RING_TILE, n_atoms - t0);
        __syncthreads();
        for (int k = threadIdx.x; k < 3 * m; k += blockDim.x) s_xyz[k] = xyz[3 * (size_t)t0 + k];
        __syncthreads();
        if (!live) continue;
        for (int k = lane; k < m; k += 32) {
            const double dx = d_sub((double)s_xyz[3 * k], cx), dy = d_sub((double)s_xyz[3 * k + 1], cy),
                         dz = d_sub((double)s_xyz[3 * k + 2], cz);
            double s = d_mul(dx, dx);
            s = d_add(s, d_mul(dy, dy));
            s = d_add(s, d_mul(dz, dz));
            if (!(s <= r2)) continue;                                   /* Bio.PDB.kdtrees radius test */
            const double d = np_norm3_f64(dx, dy, dz, blas_fma);        /* :1469 */
            if (best_i < 0 || d < best) { best = d; best_i = t0 + k; }  /* ascending index per lane: first minimum */
        }
    }
    if (!live) return;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
        /* distances of atoms inside the radius are finite (s <= r2), so this is a plain argmin, lowest index on ties */
        const bool take = oi >= 0 && (best_i < 0 || od < best || (od == best && oi < best_i));
        if (take) { best = od; best_i = oi; }
    }
    if (lane == 0) { atom_out[ring] = best_i; dist_out[ring] = best_i >= 0 ? best : 0.0; }
}

int arp_ring_nearest_run(arp_ctx* c, const float* xyz, int n_atoms, const double* centers, int n_rings, double radius,
                         int32_t* atom_out, double* dist_out)
{
    const size_t bx = sizeof(float) * 3 * (size_t)n_atoms, bc = sizeof(double) * 3 * (size_t)n_rings;
    const size_t o_c = (bx + 255) / 256 * 256, o_a = o_c + (bc + 255) / 256 * 256;
    const size_t o_d = o_a + (sizeof(int32_t) * (size_t)n_rings + 255) / 256 * 256;
    ARP_TRY(dbuf_reserve(c, c->ring_scratch, o_d + sizeof(double) * (size_t)n_rings));
    char* base = c->ring_scratch.as<char>();
    if (n_atoms) ARP_CUDA(c, cudaMemcpyAsync(base, xyz, bx, cudaMemcpyHostToDevice, c->stream));
    ARP_CUDA(c, cudaMemcpyAsync(base + o_c, centers, bc, cudaMemcpyHostToDevice, c->stream));
    k_ring_nearest<<<(unsigned)((n_rings + RING_WARPS - 1) / RING_WARPS), RING_WARPS * 32, 0, c->stream>>>(
        n_rings, n_atoms, (const double*)(base + o_c), (const float*)base, radius * radius, c->rp.blas_fma,
        (int32_t*)(base + o_a), (double*)(base + o_d));
    ARP_LAUNCHED(c);
    ARP_CUDA(c, cudaMemcpyAsync(atom_out, base + o_a, sizeof(int32_t) * (size_t)n_rings, cudaMemcpyDeviceToHost, c->stream));
    ARP_CUDA(c, cudaMemcpyAsync(dist_out, base + o_d, sizeof(double) * (size_t)n_rings, cudaMemcpyDeviceToHost, c->stream));
    ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    return ARP_OK;
}
