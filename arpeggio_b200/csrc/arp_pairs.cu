/*
 * arp_pairs.cu -- atom-atom contacts on the GPU (sm_100a).
 *
 * Replaces, for one `selection_plus` atom list (or a batch of independent lists):
 *   Bio.PDB.NeighborSearch(selection_plus).search_all(cutoff)   interactions.py:1442, :707
 *   InteractionComplex._calculate_atom_contacts                 interactions.py:693-936
 *
 * Pipeline of one run (all on the context's stream, no host round trip):
 *   memset     one zero region: counters | bounding boxes | cell counts | scan state
 *   k_bbox     per-structure bounding box (ordered-int atomics)
 *   k_geom     per-structure cell grid (edge >= cutoff), global cell numbering, prefilter band
 *   k_cellid   cell of every atom + rank inside the cell
 *   k_scan     exclusive scan of the cell counts (single pass, decoupled look-back)
 *   k_scatter  atoms into cell order as two 16-byte records: pos4 (x, y, z, original index)
 *              and att4 (packed feature word, residue, residue's prev/next link)
 *   k_search   one warp per home cell (dynamic tickets): the 13 forward neighbour cells + the
 *              home cell are 5 contiguous runs of the cell-sorted array; lane = candidate,
 *              home atoms broadcast; float32 prefilter, exact double test inside the band;
 *              ballot compaction into a per-warp shared-memory ring; dense filters; survivors
 *              appended to the hit list (8 B per pair, L2 resident) behind one cursor atomic
 *   k_classify one thread per hit: fused float32 distance + 15-bit mask rules (arp_rules.cuh),
 *              hydrogen scans and rare exact predicates compacted and run densely, records
 *              staged in shared memory and stored tile by tile with cp.async.bulk
 */
#include <cooperative_groups.h>

#include "arp_ctx.cuh"

namespace cg = cooperative_groups;

#define FULL 0xffffffffu

/* ---- programmatic dependent launch (PDL) ------------------------------------------------------
 * The pair kernels are launched with the programmatic-stream-serialization attribute: their blocks
 * become resident while the previous kernel of the stream drains and park in pdl_wait() until that
 * kernel has completed and its writes are visible.  pdl_trigger() in the primary lets the dependent
 * grid be scheduled as soon as every primary block has passed it (or exited).                     */
__device__ __forceinline__ void pdl_wait()    { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef PAIR_PROFILE       /* diagnostic build: first and last instruction of every block on the global timer */
__device__ unsigned long long g_prof[4][2048][2];
#define PROF_STAMP(kern, which) do { if (threadIdx.x == 0 && blockIdx.x < 2048) { unsigned long long t_; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_prof[kern][blockIdx.x][which] = t_; } } while (0)
#else
#define PROF_STAMP(kern, which) do { } while (0)
#endif

template <class... KArgs, class... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                            bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

/* ---- ordered-int encoding of float32 for atomic min/max ------------------------------ */
__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

/* structure of atom i (struct_off ascending, S >= 1); null struct_off = one structure */
__device__ __forceinline__ int struct_of(const int* __restrict__ struct_off, int S, int i)
{
    if (!struct_off || S == 1) return 0;
    int lo = 0, hi = S;            /* invariant: struct_off[lo] <= i < struct_off[hi] */
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (struct_off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

/* ======================================================================================
 * Grid build.  Five phases, written as block-level device functions over a VIRTUAL block id so
 * that they run either as five small kernels (batches of structures, very large inputs) or as
 * one cooperative persistent kernel with grid-wide barriers in between (k_grid_fused: a single
 * launch instead of five -- the phases are latency-, not throughput-bound at 10^5 atoms).
 * ====================================================================================== */
#define GRID_THREADS       256
#define BBOX_ATOMS_PER_VB  (GRID_THREADS * 4)

/* the threads' partial boxes (s = structure, -1: none; v = running maxima) go to bbox[s]: warp reduction when
   every live lane sees the same structure, then one block reduction, at most 6 atomics per block */
template <int WARPS>
__device__ __forceinline__ void bbox_commit(int s, unsigned (&v)[6], unsigned* __restrict__ bbox)
{
    __shared__ unsigned s_v[WARPS][6];
    __shared__ int s_s[WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int smax = __reduce_max_sync(FULL, s);
    bool uniform = __all_sync(FULL, s == smax || s == -1);
    if (uniform) {
        for (int k = 0; k < 6; ++k) v[k] = __reduce_max_sync(FULL, v[k]);
    } else if (s >= 0) {
        for (int k = 0; k < 6; ++k) atomicMax(&bbox[6 * (size_t)s + k], v[k]);
    }
    __syncthreads();                  /* the previous use's readers are done with s_v / s_s */
    if (lane == 0) {
        s_s[warp] = uniform ? smax : -1;
        for (int k = 0; k < 6; ++k) s_v[warp][k] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
        int ws = lane < WARPS ? s_s[lane] : -1;
        int bs = __reduce_max_sync(FULL, ws);
        bool buni = __all_sync(FULL, ws == bs || ws == -1);
        if (buni) {
            if (bs >= 0 && lane < 6) {
                unsigned m = 0;
                for (int w = 0; w < WARPS; ++w) if (s_s[w] >= 0) m = max(m, s_v[w][lane]);
                atomicMax(&bbox[6 * (size_t)bs + lane], m);
            }
        } else if (lane < WARPS && ws >= 0) {
            for (int k = 0; k < 6; ++k) atomicMax(&bbox[6 * (size_t)ws + k], s_v[lane][k]);
        }
    }
}

/* ---- phase 1: bounding boxes --------------------------------------------------------------
 * bbox[s][0..2] = max over atoms of ~ord(coord)  (i.e. the minimum), [3..5] = max of ord(coord).
 * Zero-initialised, so a structure without atoms keeps all zeros.                        */
__device__ __forceinline__ void dev_bbox(const float* __restrict__ xyz, const int* __restrict__ struct_off,
                                         int S, int N, unsigned* __restrict__ bbox, int vb)
{
    const int base = vb * BBOX_ATOMS_PER_VB;
    int s = -1;                       /* structure of this thread's atoms */
    unsigned v[6] = {0, 0, 0, 0, 0, 0};
    for (int t = 0; t < BBOX_ATOMS_PER_VB / GRID_THREADS; ++t) {
        int i = base + t * GRID_THREADS + threadIdx.x;
        if (i >= N) break;
        int si = struct_of(struct_off, S, i);
        if (s == -1) s = si;
        if (si != s) {                /* rare: a structure boundary inside the thread's atoms */
            for (int k = 0; k < 3; ++k) {
                unsigned o = f2ord(xyz[3 * (size_t)i + k]);
                atomicMax(&bbox[6 * (size_t)si + k], ~o);
                atomicMax(&bbox[6 * (size_t)si + 3 + k], o);
            }
            continue;
        }
        for (int k = 0; k < 3; ++k) {
            unsigned o = f2ord(xyz[3 * (size_t)i + k]);
            v[k] = max(v[k], ~o); v[3 + k] = max(v[3 + k], o);
        }
    }
    bbox_commit<GRID_THREADS / 32>(s, v, bbox);
}

/* ---- phase 2: per-structure grids -------------------------------------------------------------- */

/* grid of one structure with n > 0 atoms from its bounding box (cell_base is filled in by the caller);
   returns false when a coordinate is NaN / Inf (one cell, every test exact) */
__device__ __forceinline__ bool geom_make(const unsigned* __restrict__ bb, int n, double cutoff, int tile_x, StructGeom& g)
{
    double mn[3], mx[3], amax = 0.0;
    bool finite = true;
    for (int k = 0; k < 3; ++k) {
        mn[k] = (double)ord2f(~__ldcg(&bb[k]));
        mx[k] = (double)ord2f(__ldcg(&bb[3 + k]));
        if (!(fabs(mn[k]) <= 3.0e38) || !(fabs(mx[k]) <= 3.0e38)) finite = false;
        amax = fmax(amax, fmax(fabs(mn[k]), fabs(mx[k])));
    }
    if (!finite) for (int k = 0; k < 3; ++k) { mn[k] = 0.0; mx[k] = 0.0; }
    double r = cutoff;
    double w = r > 1e-3 ? r * 1.0001 + 1e-4 : 1e-3;
    long long d[3];
    for (;;) {
        for (int k = 0; k < 3; ++k) d[k] = (long long)floor((mx[k] - mn[k]) / w) + 1;
        /* at most 4 n + 64 cells per structure (host sizes the tables on that bound) */
        if ((double)d[0] * (double)d[1] * (double)d[2] <= 4.0 * (double)n + 64.0) break;
        w *= 1.5;
    }
    g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
    g.inv_w = 1.0 / w;
    g.dx = (int)d[0]; g.dy = (int)d[1]; g.dz = (int)d[2];
    g.ncell = g.dx * g.dy * g.dz;
    {   /* work units of k_tiles: segments of seg_x home cells of one x-row; a unit stages the segment's five
           neighbour-row windows, (5 seg_x + 9) cells, which should fit ARP_TILE_CAP atoms with room for the
           fluctuations of the cell population (a unit that does not fit is still correct: it reads global memory) */
        const double per_cell = (double)n / (double)g.ncell;
        int sx = (int)floor(((double)ARP_TILE_CAP / (1.5 * per_cell) - 9.0) / 5.0);
        sx = sx < 1 ? 1 : sx;
        sx = sx > tile_x ? tile_x : sx;
        sx = sx > g.dx ? g.dx : sx;
        g.seg_x = sx;
        g.nseg = (g.dx + sx - 1) / sx;
        g.n_units = g.nseg * g.dy * g.dz;
        g.unit_base = 0;
    }
    /* float32 prefilter band around r^2: u bounds one ulp of any coordinate */
    double u = amax * 1.1920928955078125e-07;
    double r2 = r * r;
    double band = 4.0 * r * u + 4.0 * u * u + r2 * 1.9073486328125e-06;
    double lo2 = r2 - band, hi2 = r2 + band;
    g.r2_lo = finite ? __double2float_rd(lo2) : -1.0f;
    g.r2_hi = finite ? __double2float_ru(hi2) : 3.4e38f;
    if (!(lo2 > 0.0)) g.r2_lo = -1.0f;
    return finite;
}

/* k_classify uses one conservative lower edge of the band for all structures: the minimum */
__device__ __forceinline__ unsigned r2_lo_key(float r2_lo)
{
    return r2_lo > 0.f ? 0x7f800000u - __float_as_uint(r2_lo) : 0x7f800000u;
}

/* all structures, one block of THREADS threads: grids + global cell numbering (block scan) */
template <int THREADS>
__device__ __forceinline__ void dev_geom(const unsigned* __restrict__ bbox, const int* __restrict__ struct_off,
                                         int S, int N, double cutoff, int tile_x, StructGeom* __restrict__ geom,
                                         RunMeta* __restrict__ meta)
{
    __shared__ long long s_sum[THREADS];      /* units << 32 | cells: both totals stay below 2^31 */
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int s0 = 0; s0 < S; s0 += blockDim.x) {
        int s = s0 + threadIdx.x;
        StructGeom g;
        memset(&g, 0, sizeof g);
        long long mine = 0;
        if (s < S) {
            int lo = struct_off ? struct_off[s] : 0, hi = struct_off ? struct_off[s + 1] : N;
            int n = hi - lo;
            if (n > 0) {
                if (!geom_make(bbox + 6 * (size_t)s, n, cutoff, tile_x, g)) atomicOr(&meta->fault, 1u);
                mine = ((long long)g.n_units << 32) | (long long)g.ncell;
                atomicMax(&meta->r2_lo_inv, r2_lo_key(g.r2_lo));
            }
        }
        /* block exclusive scan of (units, cells) */
        s_sum[threadIdx.x] = mine;
        __syncthreads();
        for (int off = 1; off < blockDim.x; off <<= 1) {
            long long t = threadIdx.x >= off ? s_sum[threadIdx.x - off] : 0;
            __syncthreads();
            s_sum[threadIdx.x] += t;
            __syncthreads();
        }
        long long incl = s_sum[threadIdx.x];
        long long carry = s_carry;
        if (s < S) {
            const long long excl = carry + incl - mine;
            g.cell_base = (int)(excl & 0xffffffffll);
            g.unit_base = (int)(excl >> 32);
            geom[s] = g;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        meta->n_cells = (unsigned)(s_carry & 0xffffffffll);
        meta->n_units = (unsigned)(s_carry >> 32);
    }
}

__device__ __forceinline__ int cell_coord(double v, double o, double inv_w, int dim)
{
    double t = floor((v - o) * inv_w);
    t = fmin(fmax(t, 0.0), (double)(dim - 1));     /* NaN -> 0 */
    return (int)t;
}

/* ---- phase 3: cell of every atom, rank inside the cell ---------------------------------------- */
/* Hydrogens take no part in atom-atom contacts (interactions.py:712-713 skips every pair with one before anything else):
   they get no cell and never enter the cell-sorted arrays, so the pair kernels neither test nor classify them (a
   hydrogenated structure has half of its atoms and three quarters of its within-cutoff pairs out of the way). */
__device__ __forceinline__ void dev_cellid(const float* __restrict__ xyz, const uint32_t* __restrict__ feat,
                                           const int* __restrict__ struct_off,
                                           int S, int N, const StructGeom* __restrict__ geom,
                                           int* __restrict__ cell_cnt, int* __restrict__ cell_of,
                                           int* __restrict__ rank, int i)
{
    if (i >= N) return;
    if (feat[i] & ARP_F_ELEM_H) { cell_of[i] = -1; return; }
    int s = struct_of(struct_off, S, i);
    const StructGeom* g = geom + s;
    int cx = cell_coord((double)xyz[3 * (size_t)i + 0], g->ox, g->inv_w, g->dx);
    int cy = cell_coord((double)xyz[3 * (size_t)i + 1], g->oy, g->inv_w, g->dy);
    int cz = cell_coord((double)xyz[3 * (size_t)i + 2], g->oz, g->inv_w, g->dz);
    int c = g->cell_base + (cz * g->dy + cy) * g->dx + cx;
    cell_of[i] = c;
    rank[i] = atomicAdd(&cell_cnt[c], 1);
}

/* ---- phase 4: single-pass exclusive scan (decoupled look-back) ------------------------------ */
#define SCAN_THREADS GRID_THREADS
#define ST_AGG  (1ull << 62)
#define ST_INCL (2ull << 62)

/* one tile of ARP_SCAN_TILE items by a block of THREADS threads; every tile below `tile` has been started by a
   resident block.  state[t] = flag (bits 62..63) | value: first the tile's own total (AGG), then the inclusive
   total of tiles 0..t (INCL); warp 0 looks back over 32 predecessors per step. */
template <int THREADS>
__device__ __forceinline__ void dev_scan_tile(const int* __restrict__ in, int* __restrict__ out,
                                              unsigned long long* state, long long n, int tile)
{
    constexpr int ITEMS = ARP_SCAN_TILE / THREADS;
    constexpr int WARPS = THREADS / 32;
    __shared__ int s_warp[WARPS];
    __shared__ int s_prefix;
    const long long base = (long long)tile * ARP_SCAN_TILE;
    const long long first = base + (long long)threadIdx.x * ITEMS;
    int v[ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        v[k] = (first + k < n) ? __ldcg(&in[first + k]) : 0;
        sum += v[k];
    }
    /* block scan of the thread sums */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int t = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += t;
    }
    __syncthreads();                  /* readers of the previous tile are done with the shared words */
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < WARPS ? s_warp[lane] : 0;
#pragma unroll
        for (int off = 1; off < WARPS; off <<= 1) {
            int t = __shfl_up_sync(FULL, w, off);
            if (lane >= off) w += t;
        }
        if (lane < WARPS) s_warp[lane] = w;                   /* inclusive warp totals */
        const int tile_total = __shfl_sync(FULL, w, WARPS - 1);
        volatile unsigned long long* st = state;
        int prefix = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = ST_INCL | (unsigned long long)(unsigned)tile_total;
        } else {
            if (lane == 0) st[tile] = ST_AGG | (unsigned long long)(unsigned)tile_total;
            int p = tile - 1;                                  /* nearest predecessor of this window */
            for (;;) {
                const int idx = p - lane;
                const unsigned long long w64 = idx >= 0 ? st[idx] : ST_INCL;   /* before tile 0: inclusive total 0 */
                const unsigned f = (unsigned)(w64 >> 62);
                const unsigned m_incl = __ballot_sync(FULL, f == 2u);
                const unsigned m_none = __ballot_sync(FULL, f == 0u);
                const int stop = m_incl ? __ffs(m_incl) - 1 : 31;                /* lanes 0..stop contribute */
                const unsigned need = stop == 31 ? FULL : ((2u << stop) - 1u);
                if (m_none & need) continue;                                     /* a predecessor has not published yet */
                int c = lane <= stop ? (int)(unsigned)(w64 & 0xffffffffull) : 0;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(FULL, c, off);
                prefix += c;
                if (m_incl) break;
                p -= 32;
            }
            if (lane == 0) st[tile] = ST_INCL | (unsigned long long)(unsigned)(prefix + tile_total);
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    const int warp_excl = warp ? s_warp[warp - 1] : 0;
    int run = s_prefix + warp_excl + incl - sum;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (first + k < n) out[first + k] = run;
        run += v[k];
    }
}

/* ---- phase 5: atoms into cell order ------------------------------------------------------------ */
struct ScatterArgs {
    const float* xyz; const uint32_t* feat; const int32_t* res_id; const uint16_t* rad_class;
    const int32_t* res_prev; const int32_t* res_next; const uint8_t* res_flags; const int32_t* bond_off;
    const int32_t* h_off;            /* null: no hydrogens uploaded */
    const double* h_xyz;             /* hydrogens of the donors (k_grid_reg pulls them into L2 for k_hscan) */
    const int* cell_of; const int* rank; const int* cell_start;
    float4* pos4; uint4* att4;
    uint4* arec;                     /* non-null: 32-byte records (pos4 | att4) for k_tiles instead of the three arrays above */
};

__device__ __forceinline__ void dev_scatter(const ScatterArgs& A, int N, int i)
{
    if (i >= N) return;
    if (A.cell_of[i] < 0) return;                         /* a hydrogen: not in the cell-sorted arrays (dev_cellid) */
    int dst = __ldcg(&A.cell_start[A.cell_of[i]]) + A.rank[i];
    int r = A.res_id[i];
    const uint32_t w = arp_pack_word(A.feat[i], A.res_flags[r], A.rad_class[i],
                                     A.bond_off && A.bond_off[i + 1] > A.bond_off[i]);
    const float4 p4 = make_float4(A.xyz[3 * (size_t)i], A.xyz[3 * (size_t)i + 1], A.xyz[3 * (size_t)i + 2], __int_as_float(i));
    const uint4 a4 = make_uint4(w, (uint32_t)r, (uint32_t)A.res_prev[r], (uint32_t)A.res_next[r]);
    if (A.arec) {
        A.arec[2 * (size_t)dst] = make_uint4(__float_as_uint(p4.x), __float_as_uint(p4.y), __float_as_uint(p4.z), (uint32_t)i);
        A.arec[2 * (size_t)dst + 1] = a4;
        return;
    }
    A.pos4[dst] = p4;
    A.att4[dst] = a4;
}

/* ---- phase 6 (k_tiles only): run table of every cell ---------------------------------------------
 * The home cell and its 13 forward neighbours are five contiguous runs of the cell-sorted records.  One THREAD
 * per cell writes them once -- runtab[6 j + r] = (first position, length) of run r, runtab[6 j + 5] = (home
 * atoms, float bits of the structure's upper band edge) -- instead of one warp per cell deriving them in the pair
 * kernel.  cs(i) returns cell_start[i] (from shared or global memory). */
template <class CS>
__device__ __forceinline__ void dev_runtab(const StructGeom* __restrict__ geom, int S, int j, CS cs, int2* __restrict__ runtab)
{
    int s = 0;
    if (S > 1) {                      /* last structure whose first cell is <= j (empty structures share their successor's) */
        int lo = 0, hi = S;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (geom[mid].cell_base <= j) lo = mid; else hi = mid;
        }
        s = lo;
    }
    const StructGeom* g = geom + s;
    const int gdx = g->dx, gdy = g->dy, gdz = g->dz, base = g->cell_base;
    const int lc = j - base;
    const int t2 = lc / gdx, cx = lc - t2 * gdx;
    const int cz = t2 / gdy, cy = t2 - cz * gdy;
    int4* dst = reinterpret_cast<int4*>(runtab + 6 * (size_t)j);
    const int xh = min(cx + 1, gdx - 1), xm = max(cx - 1, 0);
    auto run = [&](int r, int& beg, int& len) {
        const int y = cy + (r == 1 || r == 4 ? 1 : (r == 2 ? -1 : 0));
        const int z = cz + (r >= 2 ? 1 : 0);
        beg = 0; len = 0;
        if (y >= 0 && y < gdy && z < gdz) {
            const int row = base + (z * gdy + y) * gdx;
            beg = cs(row + (r == 0 ? cx : xm));
            len = cs(row + xh + 1) - beg;
        }
    };
    int b0, l0, b1, l1;
    run(0, b0, l0); run(1, b1, l1);
    const int nh = cs(j + 1) - b0;    /* run 0 starts with the home cell */
    dst[0] = make_int4(b0, l0, b1, l1);
    run(2, b0, l0); run(3, b1, l1);
    dst[1] = make_int4(b0, l0, b1, l1);
    run(4, b0, l0);
    dst[2] = make_int4(b0, l0, nh, __float_as_int(g->r2_hi));
}

__global__ void __launch_bounds__(256) k_runtab(const StructGeom* __restrict__ geom, int S, const RunMeta* __restrict__ meta,
                                                const int* __restrict__ cell_start, int2* __restrict__ runtab)
{
    const int n_cells = (int)meta->n_cells;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_cells; j += gridDim.x * blockDim.x)
        dev_runtab(geom, S, j, [&](int i) { return __ldcg(&cell_start[i]); }, runtab);
}

/* ---- the phases as separate kernels ------------------------------------------------------------- */
__global__ void __launch_bounds__(GRID_THREADS) k_bbox(const float* __restrict__ xyz, const int* __restrict__ struct_off,
                                                       int S, int N, unsigned* __restrict__ bbox)
{
    dev_bbox(xyz, struct_off, S, N, bbox, blockIdx.x);
}

__global__ void __launch_bounds__(GRID_THREADS) k_geom(const unsigned* __restrict__ bbox, const int* __restrict__ struct_off,
                                                       int S, int N, double cutoff, int tile_x, StructGeom* __restrict__ geom,
                                                       RunMeta* __restrict__ meta)
{
    dev_geom<GRID_THREADS>(bbox, struct_off, S, N, cutoff, tile_x, geom, meta);
}

__global__ void __launch_bounds__(GRID_THREADS) k_cellid(const float* __restrict__ xyz, const uint32_t* __restrict__ feat, const int* __restrict__ struct_off,
                                                         int S, int N, const StructGeom* __restrict__ geom,
                                                         int* __restrict__ cell_cnt, int* __restrict__ cell_of,
                                                         int* __restrict__ rank)
{
    dev_cellid(xyz, feat, struct_off, S, N, geom, cell_cnt, cell_of, rank, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(const int* __restrict__ in, int* __restrict__ out,
                                                       unsigned long long* state, unsigned int* ticket,
                                                       const unsigned int* n_dev, int n_add)
{
    __shared__ int s_tile;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);   /* tiles in order of execution start */
    __syncthreads();
    const int tile = s_tile;
    const long long n = (long long)(n_dev ? *n_dev : 0u) + n_add;
    if ((long long)tile * ARP_SCAN_TILE >= n) return;
    dev_scan_tile<SCAN_THREADS>(in, out, state, n, tile);
}

int arp_scan_exclusive(arp_ctx* c, const int* in, int* out, unsigned long long* state, unsigned int* ticket,
                       const unsigned int* n_dev, int n_add, size_t n_bound)
{
    size_t tiles = (n_bound + ARP_SCAN_TILE - 1) / ARP_SCAN_TILE;
    if (tiles == 0) tiles = 1;
    k_scan<<<(unsigned)tiles, SCAN_THREADS, 0, c->stream>>>(in, out, state, ticket, n_dev, n_add);
    ARP_LAUNCHED(c);
    return ARP_OK;
}

__global__ void __launch_bounds__(GRID_THREADS) k_scatter(int N, ScatterArgs A)
{
    dev_scatter(A, N, blockIdx.x * blockDim.x + threadIdx.x);
}

/* ---- the phases as one cooperative kernel ----------------------------------------------------- */
struct GridArgs {
    int N, S;
    int tile_x;                     /* most home cells per unit of k_tiles */
    double cutoff;
    const int* struct_off;
    unsigned* bbox;
    StructGeom* geom;
    RunMeta* meta;
    int* cell_cnt;
    unsigned long long* scan_state;
    int2* runtab;                   /* non-null: per-cell run tables for k_tiles (dev_runtab) */
    ScatterArgs sc;                 /* sc.cell_of / rank / cell_start are written by earlier phases */
    int* cell_of; int* rank; int* cell_start;
};

__global__ void __launch_bounds__(GRID_THREADS) k_grid_fused(GridArgs G)
{
    cg::grid_group grid = cg::this_grid();
    const int N = G.N;
    const int vb_atoms = (N + GRID_THREADS - 1) / GRID_THREADS;
    for (int vb = blockIdx.x; vb * BBOX_ATOMS_PER_VB < N; vb += gridDim.x)
        dev_bbox(G.sc.xyz, G.struct_off, G.S, N, G.bbox, vb);
    grid.sync();
    if (blockIdx.x == 0) dev_geom<GRID_THREADS>(G.bbox, G.struct_off, G.S, N, G.cutoff, G.tile_x, G.geom, G.meta);
    grid.sync();
    for (int vb = blockIdx.x; vb < vb_atoms; vb += gridDim.x)
        dev_cellid(G.sc.xyz, G.sc.feat, G.struct_off, G.S, N, G.geom, G.cell_cnt, G.cell_of, G.rank, vb * GRID_THREADS + threadIdx.x);
    grid.sync();
    {
        const long long n = (long long)__ldcg(&G.meta->n_cells) + 1;
        const int tiles = (int)((n + ARP_SCAN_TILE - 1) / ARP_SCAN_TILE);
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)     /* ascending per block, blocks co-resident */
            dev_scan_tile<GRID_THREADS>(G.cell_cnt, G.cell_start, G.scan_state, n, tile);
    }
    grid.sync();
    for (int vb = blockIdx.x; vb < vb_atoms; vb += gridDim.x)
        dev_scatter(G.sc, N, vb * GRID_THREADS + threadIdx.x);
    if (G.runtab) {
        const int n_cells = (int)__ldcg(&G.meta->n_cells);
        for (int j = blockIdx.x * GRID_THREADS + threadIdx.x; j < n_cells; j += gridDim.x * GRID_THREADS)
            dev_runtab(G.geom, G.S, j, [&](int i) { return __ldcg(&G.cell_start[i]); }, G.runtab);
    }
}

/* ---- the phases as one cooperative kernel, atoms held in registers ------------------------------
 * For inputs of up to REG_APT atoms per thread of a one-block-per-SM grid (about 3 * 10^5 atoms on 148 SMs):
 * every thread loads its atoms once and keeps coordinates, cell, rank and the attribute values in registers
 * across the phases; the attribute loads are issued before the first barrier and consumed after it.  One
 * structure: every block derives the grid from the bounding box itself (no barrier for the grid).  A small
 * table of cell counts (up to REG_SMEM_CELLS entries, about 45 000 atoms at protein density) is scanned by
 * every block for itself in shared memory and the atoms go to their places straight from there: two grid
 * barriers in all.  Larger tables take the single-pass scan with decoupled look-back and a third barrier
 * (the redundant scan grows with the table, the look-back does not).
 *
 * No memset in front of this kernel: block 0 zeroes the run's counters (RunMeta) here, the bounding boxes and
 * the scan state are zeroed again as soon as their last reader is past a barrier, and the cell counts are
 * zeroed by k_hscan (the last kernel of the run), so the next run finds everything clean.             */
#define REG_THREADS 1024
#define REG_APT     2
#ifndef REG_SMEM_CELLS
#define REG_SMEM_CELLS 8192     /* measured: 3 600 cells (20k atoms) 1.8 us faster in shared memory, 17 600 cells (100k) 2 us slower */
#endif
#define REG_SIDX(j)    ((j) + ((j) >> 5))
#define REG_SMEM_BYTES ((REG_SMEM_CELLS + REG_SMEM_CELLS / 32 + 32) * sizeof(int))

#ifdef GRID_PROFILE      /* diagnostic build: phase boundaries of block 0 on the global timer */
#define GRID_STAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(stamp[k])); } while (0)
#else
#define GRID_STAMP(k) do { } while (0)
#endif

template <int APT>
__global__ void __launch_bounds__(REG_THREADS, 1) k_grid_reg(GridArgs G)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) int s_off[];        /* exclusive cell offsets (REG_SMEM_CELLS entries + padding) */
    __shared__ StructGeom s_geom;
    __shared__ int s_wsum[REG_THREADS / 32];
#ifdef GRID_PROFILE
    unsigned long long stamp[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    GRID_STAMP(0);
    PROF_STAMP(0, 0);
    const int N = G.N, S = G.S;
    const int T = gridDim.x * REG_THREADS;
    const int gtid = blockIdx.x * REG_THREADS + threadIdx.x;
    if (blockIdx.x == 0) {              /* the run's counters; their first writers come after the first grid barrier */
        unsigned* m = reinterpret_cast<unsigned*>(G.meta);
        for (int k = threadIdx.x; k < (int)(sizeof(RunMeta) / sizeof(unsigned)); k += REG_THREADS) m[k] = 0u;
    }
    float x[APT], y[APT], z[APT];
    int st[APT];
    /* ---- phase 1: load; the attribute loads are in flight during the bounding-box reductions ---- */
    uint32_t af[APT], arc[APT];
    int ar[APT], ab0[APT], ab1[APT];
    int2 ah[APT];
#pragma unroll
    for (int k = 0; k < APT; ++k) {
        const int i = k * T + gtid;
        st[k] = -1;
        x[k] = y[k] = z[k] = 0.f;
        af[k] = 0; arc[k] = 0; ar[k] = 0; ab0[k] = ab1[k] = 0; ah[k] = make_int2(0, 0);
        if (i < N) {
            x[k] = G.sc.xyz[3 * (size_t)i]; y[k] = G.sc.xyz[3 * (size_t)i + 1]; z[k] = G.sc.xyz[3 * (size_t)i + 2];
            ar[k] = G.sc.res_id[i];
            af[k] = G.sc.feat[i];
            arc[k] = G.sc.rad_class[i];
            if (G.sc.bond_off) { ab0[k] = G.sc.bond_off[i]; ab1[k] = G.sc.bond_off[i + 1]; }
            if (G.sc.h_off) ah[k] = make_int2(G.sc.h_off[i], G.sc.h_off[i + 1]);
            st[k] = struct_of(G.struct_off, S, i);
        }
    }
#pragma unroll
    for (int k = 0; k < APT; ++k) {     /* block-contiguous atoms per k: one reduction each */
        unsigned v[6] = {0, 0, 0, 0, 0, 0};
        if (st[k] >= 0) {
            const unsigned ox = f2ord(x[k]), oy = f2ord(y[k]), oz = f2ord(z[k]);
            v[0] = ~ox; v[1] = ~oy; v[2] = ~oz; v[3] = ox; v[4] = oy; v[5] = oz;
        }
        if (k * T + blockIdx.x * REG_THREADS < N) bbox_commit<REG_THREADS / 32>(st[k], v, G.bbox);   /* block-uniform */
    }
    GRID_STAMP(1);
    uint32_t arf[APT];
    int ap[APT], an[APT];
#pragma unroll
    for (int k = 0; k < APT; ++k) {     /* residue of the atom: its flags and chain links */
        arf[k] = 0; ap[k] = an[k] = 0;
        if (st[k] >= 0) { arf[k] = G.sc.res_flags[ar[k]]; ap[k] = G.sc.res_prev[ar[k]]; an[k] = G.sc.res_next[ar[k]]; }
    }
    GRID_STAMP(2);
    grid.sync();
    GRID_STAMP(3);
    /* ---- phase 2: grids ---- */
    if (S == 1) {
        if (threadIdx.x == 0) {
            StructGeom g;
            memset(&g, 0, sizeof g);
            const bool finite = geom_make(G.bbox, N, G.cutoff, G.tile_x, g);
            s_geom = g;
            if (blockIdx.x == 0) {
                G.geom[0] = g;
                G.meta->n_cells = (unsigned)g.ncell;
                G.meta->n_units = (unsigned)g.n_units;
                G.meta->r2_lo_inv = r2_lo_key(g.r2_lo);
                if (!finite) atomicOr(&G.meta->fault, 1u);
            }
        }
        __syncthreads();
    } else {
        if (blockIdx.x == 0) dev_geom<REG_THREADS>(G.bbox, G.struct_off, S, N, G.cutoff, G.tile_x, G.geom, G.meta);
        grid.sync();
    }
    GRID_STAMP(4);
    /* ---- phase 3: cell of every atom, rank inside the cell ---- */
    int cell[APT], rank[APT];
#pragma unroll
    for (int k = 0; k < APT; ++k) {
        cell[k] = 0; rank[k] = 0;
        if (st[k] >= 0 && (af[k] & ARP_F_ELEM_H)) st[k] = -1;     /* hydrogens stay out of the cell grid (dev_cellid) */
        if (st[k] >= 0) {
            const StructGeom* g = S == 1 ? &s_geom : G.geom + st[k];
            const int cx = cell_coord((double)x[k], g->ox, g->inv_w, g->dx);
            const int cy = cell_coord((double)y[k], g->oy, g->inv_w, g->dy);
            const int cz = cell_coord((double)z[k], g->oz, g->inv_w, g->dz);
            cell[k] = g->cell_base + (cz * g->dy + cy) * g->dx + cx;
            rank[k] = atomicAdd(&G.cell_cnt[cell[k]], 1);
        }
    }
    uint32_t aw[APT];
#pragma unroll
    for (int k = 0; k < APT; ++k) aw[k] = arp_pack_word(af[k], arf[k], arc[k], ab1[k] > ab0[k]);
    GRID_STAMP(5);
    grid.sync();
    GRID_STAMP(6);
    if (blockIdx.x == 0)                /* every reader of the bounding boxes is past a barrier: clean for the next run */
        for (int k = threadIdx.x; k < 6 * S; k += REG_THREADS) G.bbox[k] = 0u;
    /* ---- phase 4: scan of the cell counts ---- */
    const int n = (int)(S == 1 ? (unsigned)s_geom.ncell : __ldcg(&G.meta->n_cells)) + 1;
    const bool in_smem = n <= REG_SMEM_CELLS;
    if (in_smem) {
        /* every block scans the whole table in its own shared memory: no barrier, no look-back.  The counts
           arrive with independent loads (one L2 latency for all of them); thread t then owns C consecutive
           entries; one padding word per 32 entries keeps the threads' walks off each other's banks */
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        {
            const int n4 = n >> 2;
            const int4* src = reinterpret_cast<const int4*>(G.cell_cnt);
#pragma unroll 8
            for (int j4 = threadIdx.x; j4 < n4; j4 += REG_THREADS) {
                const int4 v = __ldcg(src + j4);
                const int j = 4 * j4;
                s_off[REG_SIDX(j)] = v.x; s_off[REG_SIDX(j + 1)] = v.y; s_off[REG_SIDX(j + 2)] = v.z; s_off[REG_SIDX(j + 3)] = v.w;
            }
            if (threadIdx.x < (n & 3)) { const int j = 4 * n4 + threadIdx.x; s_off[REG_SIDX(j)] = __ldcg(G.cell_cnt + j); }
        }
        __syncthreads();
        const int C = (n + REG_THREADS - 1) / REG_THREADS;
        const int j0 = threadIdx.x * C, j1 = min(n, j0 + C);
        int sum = 0;
        for (int j = j0; j < j1; ++j) sum += s_off[REG_SIDX(j)];
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_wsum[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(FULL, w, off);
                if (lane >= off) w += t;
            }
            s_wsum[lane] = w;                                     /* inclusive warp totals */
        }
        __syncthreads();
        int run = (warp ? s_wsum[warp - 1] : 0) + incl - sum;
        for (int j = j0; j < j1; ++j) {
            const int v = s_off[REG_SIDX(j)];
            s_off[REG_SIDX(j)] = run;
            run += v;
        }
        __syncthreads();
        /* the global table (k_search reads it): every block writes its share */
        const int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
        const int lo = (int)blockIdx.x * per, hi = min(n, lo + per);
        for (int j = lo + threadIdx.x; j < hi; j += REG_THREADS) G.cell_start[j] = s_off[REG_SIDX(j)];
        GRID_STAMP(7);
        GRID_STAMP(8);
    } else {
        const int tiles = (int)(((long long)n + ARP_SCAN_TILE - 1) / ARP_SCAN_TILE);
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x)     /* ascending per block, blocks co-resident */
            dev_scan_tile<REG_THREADS>(G.cell_cnt, G.cell_start, G.scan_state, n, tile);
        GRID_STAMP(7);
        grid.sync();
        GRID_STAMP(8);
        if (blockIdx.x == 0)            /* the look-backs are over: clean for the next run */
            for (int k = threadIdx.x; k < tiles; k += REG_THREADS) G.scan_state[k] = 0ull;
    }
    /* ---- phase 5: atoms into cell order ---- */
#pragma unroll
    for (int k = 0; k < APT; ++k) {
        if (st[k] < 0) continue;
        const int i = k * T + gtid;
        const int dst = (in_smem ? s_off[REG_SIDX(cell[k])] : __ldcg(&G.cell_start[cell[k]])) + rank[k];
        if (G.sc.arec) {
            G.sc.arec[2 * (size_t)dst] = make_uint4(__float_as_uint(x[k]), __float_as_uint(y[k]), __float_as_uint(z[k]), (uint32_t)i);
            G.sc.arec[2 * (size_t)dst + 1] = make_uint4(aw[k], (uint32_t)ar[k], (uint32_t)ap[k], (uint32_t)an[k]);
        } else {
            G.sc.pos4[dst] = make_float4(x[k], y[k], z[k], __int_as_float(i));
            G.sc.att4[dst] = make_uint4(aw[k], (uint32_t)ar[k], (uint32_t)ap[k], (uint32_t)an[k]);
        }
        /* the donor's hydrogens are read once, by k_hscan at the end of the run, on its critical path: bring
           their line into L2 now (up to five hydrogens per 128-byte line) */
        if (ah[k].y > ah[k].x && G.sc.h_xyz)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(G.sc.h_xyz + 3 * (size_t)ah[k].x));
    }
    if (G.runtab) {                     /* run tables of the cells for k_tiles: one thread per cell */
        const int n_cells = n - 1;
        if (in_smem) {
            for (int j = gtid; j < n_cells; j += T)
                dev_runtab(S == 1 ? &s_geom : G.geom, S, j, [&](int i) { return s_off[REG_SIDX(i)]; }, G.runtab);
        } else {
            for (int j = gtid; j < n_cells; j += T)
                dev_runtab(S == 1 ? &s_geom : G.geom, S, j, [&](int i) { return __ldcg(&G.cell_start[i]); }, G.runtab);
        }
    }
    GRID_STAMP(9);
    PROF_STAMP(0, 1);
#ifdef GRID_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 0)
        printf("grid ns: load+bbox %llu attr %llu sync1 %llu geom %llu cellid %llu sync2 %llu scan %llu sync3 %llu scatter %llu | total %llu\n",
               stamp[1] - stamp[0], stamp[2] - stamp[1], stamp[3] - stamp[2], stamp[4] - stamp[3], stamp[5] - stamp[4],
               stamp[6] - stamp[5], stamp[7] - stamp[6], stamp[8] - stamp[7], stamp[9] - stamp[8], stamp[9] - stamp[0]);
#endif
}

/* ---- k_search ---------------------------------------------------------------------------------
 * Neighbour search.  One warp per home cell (dynamic tickets of SEARCH_CELLS cells): the home cell and
 * its 13 forward neighbours are 5 contiguous runs of the cell-sorted array; lane = candidate (held in
 * registers), home atoms broadcast by shuffle; float32 FMA d^2 against the upper edge of the band around
 * r^2; ballot compaction into a per-warp shared-memory queue that is appended to the global candidate
 * list behind one cursor atomic per >= 128 candidates.  The exact test, the orientation and the filters
 * happen densely in k_classify.  The kernel is small on purpose: many warps in flight hide the loads. */
#define SEARCH_WARPS  8
#define SEARCH_SLOTS  4                     /* candidates held per lane */
#define SEARCH_CHUNK  (32 * SEARCH_SLOTS)
#define SEARCH_QCAP   256                   /* queue capacity per warp: < DRAIN before a home atom, + <= CHUNK hits */
#define SEARCH_DRAIN  128
#ifndef SEARCH_CELLS
#define SEARCH_CELLS  4
#endif
#ifndef SEARCH_NC
#define SEARCH_NC 1                         /* > 1: dynamic tickets from that many class counters (as in k_classify) */
#endif
/* cells per ticket (<= 4); 8 lanes describe one cell's runs */

struct SearchArgs {
    const float4* pos4;
    const int*    cell_start;
    const StructGeom* geom;
    RunMeta*      meta;
    uint2*        raw;                      /* candidate pairs (cell-sorted indices), float32 d^2 <= r2_hi */
    unsigned long long cap;
};

/* the warp's queued candidates go to the global candidate list behind one cursor atomic */
__device__ __forceinline__ void search_flush(const SearchArgs& A, const uint2* q, unsigned n, int lane)
{
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&A.meta->n_raw, (unsigned long long)n);
    base = __shfl_sync(FULL, base, 0);
    /* one 64-bit store per entry: k_classify, which may already be reading the list, sees an entry whole or not at all */
    for (unsigned r = lane; r < n; r += 32)
        if (base + r < A.cap)
            reinterpret_cast<unsigned long long*>(A.raw)[base + r] = reinterpret_cast<const unsigned long long*>(q)[r];
    __syncwarp();
}

/* n / d for 0 <= n < 2^24, d >= 1: float reciprocal estimate, then exact correction */
__device__ __forceinline__ int fast_div(int n, int d)
{
    if (n >= (1 << 24)) return n / d;
    int q = __float2int_rz(__fdividef((float)n, (float)d));
    int r = n - q * d;
    while (r < 0) { --q; r += d; }
    while (r >= d) { ++q; r -= d; }
    return q;
}

/* one chunk of <= 32 * NS candidates (registers) against the home atoms [0, h_end) of the cell */
template <int NS>
__device__ __forceinline__ void search_chunk(const SearchArgs& A, uint2* q, uint32_t q_addr, unsigned& qcount,
                                             const float (&cxs)[SEARCH_SLOTS], const float (&cys)[SEARCH_SLOTS],
                                             const float (&czs)[SEARCH_SLOTS], const int (&cg)[SEARCH_SLOTS],
                                             int k_lane, bool first_chunk, int hb, int h_end, float r2_hi,
                                             int lane, unsigned lt_mask)
{
    for (int h = 0; h < h_end; ++h) {
        float hx, hy, hz;
        if (first_chunk && h < 32) {       /* candidate k = h of run 0 is home atom h */
            hx = __shfl_sync(FULL, cxs[0], h); hy = __shfl_sync(FULL, cys[0], h); hz = __shfl_sync(FULL, czs[0], h);
        } else {
            const float4 hp = A.pos4[hb + h];
            hx = hp.x; hy = hp.y; hz = hp.z;
        }
        const unsigned home = (unsigned)(hb + h);
        const int hk = h - k_lane;         /* candidate k = k_lane + 32 sl is live iff 32 sl > hk */
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            const float ddx = hx - cxs[sl], ddy = hy - cys[sl], ddz = hz - czs[sl];
            const float d2 = __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, __fmul_rn(ddx, ddx)));
            const bool hit = (d2 <= r2_hi) && (32 * sl > hk);          /* padding lanes sit at +inf */
            const unsigned m = __ballot_sync(FULL, hit);
            if (hit) {
                const uint32_t addr = q_addr + 8u * (qcount + __popc(m & lt_mask));
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(addr), "r"(home), "r"((unsigned)cg[sl]) : "memory");
            }
            qcount += __popc(m);
        }
        if (qcount >= SEARCH_DRAIN) {
            __syncwarp();
            search_flush(A, q, qcount, lane);
            qcount = 0;
        }
    }
}

#ifndef SEARCH_MINB
#define SEARCH_MINB 4
#endif
#ifndef SEARCH_GRID_MULT
#define SEARCH_GRID_MULT SEARCH_MINB
#endif
__global__ void __launch_bounds__(SEARCH_WARPS * 32, SEARCH_MINB) k_search(SearchArgs A)
{
    __shared__ uint2 s_queue[SEARCH_WARPS][SEARCH_QCAP];
    __shared__ int2  s_runs[SEARCH_WARPS][SEARCH_CELLS][8];      /* per cell: 5 x (first index, prefix); [5] = (total, nh); [6] = band */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2* q = s_queue[warp];
    const uint32_t q_addr = (uint32_t)__cvta_generic_to_shared(q);
    unsigned qcount = 0;                        /* linear queue: drained completely once it holds >= SEARCH_DRAIN hits */
    unsigned long long ncand = 0;
    unsigned nonempty = 0;
    const unsigned lt_mask = (1u << lane) - 1u;

    pdl_wait();                                 /* the grid build has completed */
    pdl_trigger();
    PROF_STAMP(1, 0);
    const int n_cells = (int)A.meta->n_cells;
    int s = 0;                                  /* warp-uniform: structure of the ticket's first cell */
    int s_end = A.geom[0].cell_base + A.geom[0].ncell;

    /* tickets: the first one is static (ticket = global warp id), the rest come from a shared counter that
       starts at the number of warps; a plain read screens the counter so that late warps leave without
       an atomic.  Measured alternatives, all slower on the 100k-atom job (profiles/README.md): a counter per
       block in shared memory over a contiguous cell range per block, static tickets strided over the warps,
       and a grid of up to four times the resident blocks (the block scheduler as the balancer). */
    const unsigned n_warps_total = gridDim.x * SEARCH_WARPS;
    const unsigned n_tickets = ((unsigned)n_cells + SEARCH_CELLS - 1) / SEARCH_CELLS;
#if SEARCH_NC > 1
    const unsigned nc = min((unsigned)SEARCH_NC, gridDim.x);
    const unsigned cls = blockIdx.x % nc;
    unsigned* const ticket_ctr = &A.meta->ticket_srch[cls].v;
    const unsigned class_warps = ((gridDim.x - cls + nc - 1) / nc) * SEARCH_WARPS;
#else
    unsigned* const ticket_ctr = &A.meta->ticket_search;
#endif
    bool first = true;
    for (;;) {
        unsigned t = 0;
        if (first) {
#if SEARCH_NC > 1
            t = ((blockIdx.x / nc) * SEARCH_WARPS + warp) * nc + cls;
#else
            t = blockIdx.x * SEARCH_WARPS + warp;
#endif
            first = false;
        } else {
#if SEARCH_NC > 1
            if (lane == 0) t = (class_warps + atomicAdd(ticket_ctr, 1u)) * nc + cls;
#else
            if (lane == 0) {
                t = n_tickets;
                if (*(volatile unsigned*)ticket_ctr + n_warps_total < n_tickets)
                    t = atomicAdd(ticket_ctr, 1u) + n_warps_total;
            }
#endif
            t = __shfl_sync(FULL, t, 0);
        }
        if (t >= n_tickets) break;
        const long long c0l = (long long)t * SEARCH_CELLS;
        const int c0 = (int)c0l;
        while (c0 >= s_end) { ++s; s_end = A.geom[s].cell_base + A.geom[s].ncell; }   /* tickets ascend */
        /* ---- run tables of the ticket's cells: lane = (cell q, run r) ---- */
        {
            const int qc = lane >> 3, r = lane & 7;
            const int c = c0 + qc;
            int rbeg = 0, rlen = 0, nh = 0;
            float band_lo = 0.f, band_hi = 0.f;
            if (qc < SEARCH_CELLS && c < n_cells) {
                const StructGeom* gp = A.geom + s;
                while (c >= gp->cell_base + gp->ncell) ++gp;                  /* the cell may lie in a later structure */
                band_lo = gp->r2_lo; band_hi = gp->r2_hi;
                if (r < 5) {
                    const int gdx = gp->dx, gdy = gp->dy, gdz = gp->dz, gbase = gp->cell_base;
                    const int lc = c - gbase;
                    const int t2 = fast_div(lc, gdx), cx = lc - t2 * gdx;
                    const int cz = fast_div(t2, gdy), cy = t2 - cz * gdy;
                    const int y = cy + (r == 1 || r == 4 ? 1 : (r == 2 ? -1 : 0));
                    const int z = cz + (r >= 2 ? 1 : 0);
                    const int x0 = r == 0 ? cx : max(cx - 1, 0);
                    const int x1 = min(cx + 1, gdx - 1);
                    if (y >= 0 && y < gdy && z < gdz) {
                        const int row = gbase + (z * gdy + y) * gdx;
                        rbeg = A.cell_start[row + x0];
                        rlen = A.cell_start[row + x1 + 1] - rbeg;
                    }
                    if (r == 0) nh = A.cell_start[c + 1] - rbeg;
                }
            }
            /* exclusive prefix of the run lengths inside each group of 8 lanes */
            int incl = rlen;
#pragma unroll
            for (int off = 1; off < 8; off <<= 1) {
                int v = __shfl_up_sync(FULL, incl, off, 8);
                if (r >= off) incl += v;
            }
            const int total = __shfl_sync(FULL, incl, 7, 8);
            nh = __shfl_sync(FULL, nh, 0, 8);
            __syncwarp();
            if (qc < SEARCH_CELLS) {
                if (r < 5) s_runs[warp][qc][r] = make_int2(rbeg, incl - rlen);
                else if (r == 5) s_runs[warp][qc][5] = make_int2(total, nh);
                else if (r == 6) s_runs[warp][qc][6] = make_int2(__float_as_int(band_lo), __float_as_int(band_hi));
            }
            __syncwarp();
        }
        for (int qc = 0; qc < SEARCH_CELLS; ++qc) {
            const int2 tn = s_runs[warp][qc][5];
            const int total = tn.x, nh = tn.y;
            if (nh == 0) continue;
            ++nonempty;
            const int2 r0 = s_runs[warp][qc][0], r1 = s_runs[warp][qc][1], r2 = s_runs[warp][qc][2],
                       r3 = s_runs[warp][qc][3], r4 = s_runs[warp][qc][4];
            const int hb = r0.x;
            const float r2_hi = __int_as_float(s_runs[warp][qc][6].y);
            /* tests of this cell: home atom h meets candidates k > h */
            ncand += (unsigned long long)((long long)nh * total - (long long)nh * (nh + 1) / 2);

            for (int k0 = 0; k0 < total; k0 += SEARCH_CHUNK) {
                float cxs[SEARCH_SLOTS], cys[SEARCH_SLOTS], czs[SEARCH_SLOTS];
                int   cg[SEARCH_SLOTS];
#pragma unroll
                for (int sl = 0; sl < SEARCH_SLOTS; ++sl) {
                    const int k = k0 + sl * 32 + lane;
                    cg[sl] = 0;
                    cxs[sl] = cys[sl] = czs[sl] = __int_as_float(0x7f800000);     /* +inf: never within any cutoff */
                    if (k < total) {
                        const int g = k < r1.y ? r0.x + k : k < r2.y ? r1.x + (k - r1.y) : k < r3.y ? r2.x + (k - r2.y)
                                    : k < r4.y ? r3.x + (k - r3.y) : r4.x + (k - r4.y);
                        const float4 p = A.pos4[g];
                        cxs[sl] = p.x; cys[sl] = p.y; czs[sl] = p.z;
                        cg[sl] = g;
                    }
                }
                /* home atoms that still have candidates in this chunk: k > h */
                const int h_end = min(nh, k0 + SEARCH_CHUNK - 1);
                const int left = total - k0;
                if (left > 96)      search_chunk<4>(A, q, q_addr, qcount, cxs, cys, czs, cg, k0 + lane, k0 == 0, hb, h_end, r2_hi, lane, lt_mask);
                else if (left > 64) search_chunk<3>(A, q, q_addr, qcount, cxs, cys, czs, cg, k0 + lane, k0 == 0, hb, h_end, r2_hi, lane, lt_mask);
                else if (left > 32) search_chunk<2>(A, q, q_addr, qcount, cxs, cys, czs, cg, k0 + lane, k0 == 0, hb, h_end, r2_hi, lane, lt_mask);
                else                search_chunk<1>(A, q, q_addr, qcount, cxs, cys, czs, cg, k0 + lane, k0 == 0, hb, h_end, r2_hi, lane, lt_mask);
            }
        }
    }
    if (qcount) {
        __syncwarp();
        search_flush(A, q, qcount, lane);
    }
    if (lane == 0) {
        if (ncand) atomicAdd(&A.meta->n_candidates, ncand);
        if (nonempty) atomicAdd(&A.meta->n_cells_nonempty, nonempty);
    }
#ifdef PAIR_PROFILE
    __syncthreads();
    PROF_STAMP(1, 1);
#endif
}

/* ---- k_classify ---------------------------------------------------------------------------------
 * Everything per pair, one lane per pair.  Every WARP owns tiles of CLS_TILE consecutive candidates and
 * runs them without any block barrier:
 *   stage 0  32 candidates per round: the exact double test of Bio.PDB.kdtrees inside the float32 band,
 *            orientation (atom_bgn = lower list index), the reference's `continue` filters
 *            (interactions.py:712-741)
 *   stage 1  the survivors, still in their lanes: exact float32 distance, proximity bit, metal, the
 *            feature bits that need no angle (rule_classify_core, interactions.py:743-936); records are
 *            compacted (ballot) into the warp's staging tile; pairs that need a hydrogen scan (is_hbond /
 *            is_weak_hbond) or a rarer predicate (halogen weak hbond, xbond) append a 32-bit work item
 *   stage 2  the work list is processed densely -- 32 items per round -- and each result bit is OR-ed
 *            into the staged record
 *   stage 3  the finished tile leaves with one bulk asynchronous copy shared -> global (cp.async.bulk,
 *            TMA engine) behind one cursor atomic, double buffered so that the store of one tile overlaps
 *            the arithmetic of the next.                                                              */
#define CLS_WARPS   8
#ifndef CLS_TILE
#define CLS_TILE    64
#endif
#define CLS_TAB_K   8                       /* radius tables up to K x K = 256 entries are staged in shared memory */
#define CLS_ITEMS   (3 * CLS_TILE)          /* per pair at most: is_hbond scan + (is_weak_hbond scan | halogen) + xbond */
#define CLS_SMEM_PER_WARP (2 * CLS_TILE * 16 + CLS_ITEMS * 4 + CLS_TILE * 8)
#define CLS_SMEM    (CLS_WARPS * CLS_SMEM_PER_WARP)

struct ClassifyArgs {
    const unsigned long long* h_reach;      /* upload generation << 32 | float bits of the longest donor-hydrogen distance (k_hreach) */
    unsigned      h_gen;                    /* generation of the current upload */
    const float4* pos4;
    const uint4*  att4;
    uint2*        raw;                      /* candidate list; with the early start every entry is zeroed again once it is read */
    RunMeta*      meta;
    arp_pair*     out;
    unsigned long long cap;                 /* capacity of the candidate list */
    uint4*        work;                     /* deferred predicates: (donor, acceptor | halogen, record index, kind) */
    unsigned long long work_cap;
    double        r2;
    int           include_seq_adjacent;
    ArpSide       side;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store_tile(void* gdst, const void* ssrc, uint32_t bytes)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read_1()
{
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

/* item = radius class of the acceptor / halogen << 16 | record index in the tile << 4 | direction << 3 | kind (bits 0..2):
   kind 1..3 = ARP_HB_NEED_* of a hydrogen scan, 4 = halogen weak hbond, 5 = xbond;
   direction 0: donor = bgn, 1: donor = end */
#define CLS_KIND_HAL   4u
#define CLS_KIND_XBOND 5u


/* candidate `pos` of the list: read from L2, waited for while k_search is still running (the cursor is advanced
   before the entries are stored), and zeroed for the next run */
#define CLS_FAULT_HANDOFF 2u
template <bool EARLY>
__device__ __forceinline__ uint2 cls_fetch(const ClassifyArgs& A, unsigned long long pos, bool final)
{
    if (!EARLY) return A.raw[pos];
    /* the entry travels as ONE 64-bit word in both directions (search_flush), so "non-zero" means "in place" */
    const unsigned long long* p64 = reinterpret_cast<const unsigned long long*>(A.raw) + pos;
    unsigned long long w = __ldcg(p64);
    if (!final) {
        unsigned spins = 0;
        while (w == 0ull) {
            w = __ldcg(p64);
            /* a producer that slow (time slicing, a debugger): flagged, and the host repeats the run with k_classify
               waiting for k_search (arp_api.cu, pairs_finish) */
            if (++spins > (1u << 22)) { atomicOr(&A.meta->fault, CLS_FAULT_HANDOFF); break; }
        }
    }
    reinterpret_cast<unsigned long long*>(A.raw)[pos] = 0ull;
    return make_uint2((unsigned)(w & 0xffffffffull), (unsigned)(w >> 32));
}
#ifndef CLS_MINB
#define CLS_MINB 4
#endif
template <bool EARLY>
__global__ void __launch_bounds__(CLS_WARPS * 32, CLS_MINB) k_classify(ClassifyArgs A, ArpRuleParams P)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = s_dyn + (size_t)warp * CLS_SMEM_PER_WARP;
    int4* const rec0 = reinterpret_cast<int4*>(mine);                       /* 2 staging tiles */
    uint32_t* const items = reinterpret_cast<uint32_t*>(mine + 2 * CLS_TILE * 16);
    uint2* const surv = reinterpret_cast<uint2*>(mine + 2 * CLS_TILE * 16 + CLS_ITEMS * 4);
    const unsigned lt_mask = (1u << lane) - 1u;
    /* small radius tables live in shared memory: one dependent global load less per pair */
    __shared__ float4 s_radtab[CLS_TAB_K * CLS_TAB_K];
    __shared__ double s_vdw[CLS_TAB_K];
    __shared__ float s_hlim[CLS_TAB_K];
    A.side.hlim = nullptr;
    if (A.side.K <= CLS_TAB_K) {
        for (int k = threadIdx.x; k < A.side.K * A.side.K; k += blockDim.x) s_radtab[k] = A.side.radtab[k];
        for (int k = threadIdx.x; k < A.side.K; k += blockDim.x) {
            s_vdw[k] = A.side.vdw[k];
            /* donor farther than this from an acceptor of class k: none of its hydrogens can be within
               h_vdw + vdw_k + comp (utils.py:89, :149) of it; generous rounding margin on top */
            const unsigned long long hr = *A.h_reach;     /* an older generation: this upload has no hydrogens at all */
            const float reach = (unsigned)(hr >> 32) == A.h_gen ? __uint_as_float((unsigned)hr) : __int_as_float(0xff800000);
            const double lim = (double)reach + P.h_vdw + A.side.vdw[k] + P.vdw_comp;
            s_hlim[k] = lim == lim ? __double2float_ru(lim) * 1.000001f + 2e-3f : __int_as_float(0x7f800000);
        }
        __syncthreads();
        A.side.radtab = s_radtab;
        A.side.vdw = s_vdw;
        A.side.hlim = s_hlim;
    }
    pdl_trigger();
    PROF_STAMP(2, 0);
    /* The blocks of this kernel become resident while k_search drains (programmatic dependent launch) and start
       on the candidates that are already there instead of waiting for its last warp: the candidate cursor tells
       how far the list has been handed out, and an entry is in place once it is non-zero (a pair never has two
       zero indices; the list is all zero before a run because every reader zeroes what it has read).  A warp
       whose next tile is not handed out yet sleeps in griddepcontrol.wait until k_search has completed; from
       then on the count is final.  What k_search leaves in r2_lo_inv was written by the grid build, which had
       completed before k_search could trigger this launch. */
    const volatile unsigned long long* const n_raw_p = &A.meta->n_raw;
    const float r2_lo = A.meta->r2_lo_inv == 0x7f800000u ? -1.0f : __uint_as_float(0x7f800000u - A.meta->r2_lo_inv);
    bool final = !EARLY;                                         /* k_search has completed: n and n_tiles are valid */
    unsigned long long n = 0, n_tiles = 0;
    int buf = 0;
    if (!EARLY) {                                                /* large inputs: the drain of k_search is short against the job */
        pdl_wait();
        n = *n_raw_p;
        if (n > A.cap) n = A.cap;                                /* overflowing run: host repeats it with a larger buffer */
        n_tiles = (n + CLS_TILE - 1) / CLS_TILE;
    }
    /* Tiles.  Without the early start all blocks begin together: plain stride over the warps (tile = global warp
       id + k * warps; contiguous shares per warp were measured 2 us slower -- the candidate list is not uniform
       along its length and the stride spreads the expensive stretches).  With the early start the blocks begin
       whenever k_search frees a slot, so the tiles are handed out dynamically: ARP_CLS_COUNTERS counters share
       the work, counter c owns the tiles congruent to c and serves the blocks congruent to c; a warp's first tile
       is static (its rank among the warps of its class), the rest come from the counter, requested one tile
       ahead so that the atomic's latency is hidden. */
    const unsigned nc = min((unsigned)ARP_CLS_COUNTERS, gridDim.x);      /* classes in use: every one needs a block */
    const unsigned cls = blockIdx.x % nc;
    unsigned* const ticket_ctr = &A.meta->ticket_cls[cls].v;
    const unsigned long long class_warps = (unsigned long long)((gridDim.x - cls + nc - 1) / nc) * CLS_WARPS;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * CLS_WARPS;
    unsigned long long tile = EARLY ? ((unsigned long long)(blockIdx.x / nc) * CLS_WARPS + warp) * nc + cls
                                    : (unsigned long long)blockIdx.x * CLS_WARPS + warp;
    unsigned long long snap = 0;                                 /* the cursor when it was last read: a lower bound */
    for (;;) {
        if (EARLY && !final && (tile + 1) * CLS_TILE > min(snap, A.cap)) {
            if (lane == 0) snap = *n_raw_p;                      /* look again */
            snap = __shfl_sync(FULL, snap, 0);
            if ((tile + 1) * CLS_TILE > min(snap, A.cap)) {      /* not handed out yet (or the last, partial tile) */
                pdl_wait();                                      /* k_search has completed */
                final = true;
                n = *n_raw_p;
                if (n > A.cap) n = A.cap;
                n_tiles = (n + CLS_TILE - 1) / CLS_TILE;
            }
        }
        if (final && tile >= n_tiles) break;
        unsigned next_ticket = 0;                                /* requested now, used after this tile */
        unsigned long long seen = 0;
        if (EARLY) {
            if (lane == 0) next_ticket = atomicAdd(ticket_ctr, 1u);
            if (lane == 1 && !final) seen = *n_raw_p;
        }
        const unsigned long long base = tile * CLS_TILE;
        const unsigned cnt = final ? (unsigned)min((unsigned long long)CLS_TILE, n - base) : (unsigned)CLS_TILE;
        int4* rec = rec0 + buf * CLS_TILE;
        if (lane == 0) bulk_store_wait_read_1();                 /* the store that last used this buffer has read it */
        __syncwarp();
        /* ---- stages 0 + 1, 32 candidates per round ---- */
        unsigned nsurv = 0, n_items = 0, n_rare = 0;             /* rare items are staged from the back of `items` */
        unsigned long long o_early = 0;
        uint2 e_next = lane < cnt ? cls_fetch<EARLY>(A, base + lane, final) : make_uint2(0, 0);   /* candidates are fetched one round ahead */
#pragma unroll 1
        for (unsigned i0 = 0; i0 < cnt; i0 += 32) {
            const unsigned idx = i0 + lane;
            bool keep = false;
            uint2 e = e_next;
            if (idx + 32 < cnt) e_next = cls_fetch<EARLY>(A, base + idx + 32, final);
            float4 pa, pb;
            uint4 ab, ae;
            if (idx < cnt) {
                pa = A.pos4[e.x]; pb = A.pos4[e.y];
                ab = A.att4[e.x]; ae = A.att4[e.y];
                const float ddx = pa.x - pb.x, ddy = pa.y - pb.y, ddz = pa.z - pb.z;
                const float d2 = __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, __fmul_rn(ddx, ddx)));   /* as in k_search */
                keep = true;
                if (!(d2 <= r2_lo)) keep = kd_within(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, A.r2);
                if (__float_as_int(pb.w) < __float_as_int(pa.w)) {          /* atom_bgn = lower list index */
                    const unsigned t = e.x; e.x = e.y; e.y = t;
                    const uint4 tt = ab; ab = ae; ae = tt;
                    const float4 tp = pa; pa = pb; pb = tp;
                }
                keep = keep && rule_pair_survives(ab.x, (int)ab.y, (int)ab.z, (int)ab.w, ae.x, (int)ae.y, (int)ae.z, (int)ae.w,
                                                  A.include_seq_adjacent);
            }
            /* survivors keep their lane for the rules; their records are compacted into the staging tile */
            const unsigned mk = __ballot_sync(FULL, keep);
            const unsigned slot = nsurv + __popc(mk & lt_mask);
            nsurv += __popc(mk);
            uint32_t work = 0;
            /* the record cursor is reserved as soon as the tile's last round knows how many pairs survive: the atomic's
               round trip then runs under the rules of that round instead of in front of the store */
            if (i0 + 32 >= cnt && lane == 0 && nsurv) o_early = atomicAdd(&A.meta->n_pairs, (unsigned long long)nsurv);
            if (keep) {
                const int ib = __float_as_int(pa.w), ie = __float_as_int(pb.w);
                /* atom_bgn's range in the bond table: fetched now, needed at the very end of the rules */
                int bb0 = 0, bb1 = 0;
                if (ab.x & ARPK_HAS_BOND) { bb0 = A.side.bond_off[ib]; bb1 = A.side.bond_off[ib + 1]; }
                uint32_t mask; float dist;
                rule_classify_core(A.side, P, ib, ie, pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, ab.x, ae.x, &mask, &dist, &work, bb0, bb1);
                rec[slot] = make_int4(ib, ie, (int)mask, __float_as_int(dist));
                if (work) surv[slot] = make_uint2((unsigned)ib, (unsigned)ie);   /* k_hscan finds donor / acceptor through this: ORIGINAL indices */
            }
            /* append the work items of this round, one kind of slot at a time (ballot compaction) */
            if (__any_sync(FULL, work != 0)) {
                const uint32_t it0 = ((ae.x & ARPK_RAD_MASK) << 16) | (slot << 4);         /* donor = bgn, acceptor = end */
                const uint32_t it1 = ((ab.x & ARPK_RAD_MASK) << 16) | (slot << 4) | 8u;    /* donor = end, acceptor = bgn */
                unsigned m = __ballot_sync(FULL, (work & ARP_WORK_SCAN0) != 0);
                if (work & ARP_WORK_SCAN0) items[n_items + __popc(m & lt_mask)] = it0 | (work & 3u);
                n_items += __popc(m);
                m = __ballot_sync(FULL, (work & ARP_WORK_SCAN1) != 0);
                if (work & ARP_WORK_SCAN1) items[n_items + __popc(m & lt_mask)] = it1 | ((work >> 2) & 3u);
                n_items += __popc(m);
                const uint32_t rare = work & (ARP_WORK_HAL0 | ARP_WORK_HAL1 | ARP_WORK_XB0 | ARP_WORK_XB1);
                if (__any_sync(FULL, rare != 0)) {
                    m = __ballot_sync(FULL, (rare & (ARP_WORK_HAL0 | ARP_WORK_HAL1)) != 0);
                    if (rare & (ARP_WORK_HAL0 | ARP_WORK_HAL1))
                        items[CLS_ITEMS - 1 - (n_rare + __popc(m & lt_mask))] = ((rare & ARP_WORK_HAL1) ? it1 : it0) | CLS_KIND_HAL;
                    n_rare += __popc(m);
                    m = __ballot_sync(FULL, (rare & (ARP_WORK_XB0 | ARP_WORK_XB1)) != 0);
                    if (rare & (ARP_WORK_XB0 | ARP_WORK_XB1))
                        items[CLS_ITEMS - 1 - (n_rare + __popc(m & lt_mask))] = ((rare & ARP_WORK_XB1) ? it1 : it0) | CLS_KIND_XBOND;
                    n_rare += __popc(m);
                }
            }
        }
        tile = EARLY ? (class_warps + __shfl_sync(FULL, next_ticket, 0)) * nc + cls : tile + n_warps;
        if (EARLY) snap = __shfl_sync(FULL, seen, 1);
        if (nsurv == 0) continue;                                /* nothing staged: the buffer stays free */
        __syncwarp();
        /* ---- stage 3: the tile leaves through the TMA engine; its work items go to the global work list ---- */
        /* The hydrogen scans fill the work list from its front, the rare predicates (halogen weak hbond, xbond:
           a few per cent of the items) from its back: one of them in a chunk of 32 would make the whole warp of
           k_hscan walk a second, different chain of loads. */
        unsigned long long o = 0, ow = 0, owr = 0;
        if (lane == 0) {
            o = o_early;
            bulk_store_tile(A.out + o, rec, nsurv * (uint32_t)sizeof(arp_pair));
            if (n_items) ow = atomicAdd(&A.meta->n_work, (unsigned long long)n_items);
            if (n_rare) owr = atomicAdd(&A.meta->n_work_rare, (unsigned long long)n_rare);
        }
        o = __shfl_sync(FULL, o, 0);
        ow = __shfl_sync(FULL, ow, 0);
        owr = __shfl_sync(FULL, owr, 0);
        for (unsigned w = lane; w < n_items + n_rare; w += 32) {
            const bool is_rare = w >= n_items;
            const uint32_t it = is_rare ? items[CLS_ITEMS - 1 - (w - n_items)] : items[w];
            const unsigned slot = (it >> 4) & 0xfffu;
            uint2 e = surv[slot];
            if (it & 8u) { const unsigned t = e.x; e.x = e.y; e.y = t; }   /* e.x = donor, e.y = acceptor / halogen */
            const unsigned long long pos = is_rare ? owr + (w - n_items) : ow + w;
            if (pos < A.work_cap)
                A.work[is_rare ? A.work_cap - 1 - pos : pos] = make_uint4(e.x, e.y, (uint32_t)(o + slot), (it & 7u) | ((it >> 16) << 8));
        }
        __syncwarp();
        buf ^= 1;
    }
    if (lane == 0) bulk_store_wait_read_all();                   /* shared memory must outlive the copies */
#ifdef PAIR_PROFILE
    __syncthreads();
    PROF_STAMP(2, 1);
#endif
}

/* ---- k_hscan -----------------------------------------------------------------------------------
 * The deferred predicates of k_classify, one thread per work item, dense: utils.is_hbond /
 * is_weak_hbond (one pass over the donor's hydrogens serves both), is_halogen_weak_hbond, is_xbond
 * (utils.py:73-179).  A true predicate ORs its SIFt bit into the finished record.                  */
struct HscanArgs {
    const float*  xyz;                      /* coordinates in upload order (the work items carry original atom indices) */
    const uint4*  work;
    RunMeta*      meta;
    unsigned long long work_cap;
    arp_pair*     out;
    ArpSide       side;
    int*          clean_cnt;                /* the cell counts of the grid build, zeroed here for the next run */
};

#ifndef HSCAN_MINB
#define HSCAN_MINB 4
#endif
#define HSCAN_WARPS 8
__global__ void __launch_bounds__(HSCAN_WARPS * 32, HSCAN_MINB) k_hscan(HscanArgs A, ArpRuleParams P)
{
    __shared__ double s_vdw[CLS_TAB_K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (A.side.K <= CLS_TAB_K) {                                 /* uploaded table, not written by the run */
        if (threadIdx.x < A.side.K) s_vdw[threadIdx.x] = A.side.vdw[threadIdx.x];
        __syncthreads();
        A.side.vdw = s_vdw;
    }
    pdl_wait();                                                  /* k_classify has completed */
    PROF_STAMP(3, 0);
    unsigned long long n = A.meta->n_work, n_rare = A.meta->n_work_rare;
    if (n + n_rare > A.work_cap) n = n_rare = 0;                 /* the two ends of the list met: host repeats the run with a larger list */
    /* chunks of 32 items per warp, strided over the warps (chunks from a counter, one or sixteen, were measured
       1-2 us slower: a chunk is 4-7 us of latency and a warp sees two of them) */
    const unsigned long long c_front = (n + 31) / 32, n_chunks = c_front + (n_rare + 31) / 32;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * HSCAN_WARPS;
    for (unsigned long long chunk = (unsigned long long)blockIdx.x * HSCAN_WARPS + warp; chunk < n_chunks; chunk += n_warps) {
        const bool back = chunk >= c_front;                      /* a chunk of rare predicates, from the back of the list */
        const unsigned long long w = (back ? chunk - c_front : chunk) * 32 + lane;
        if (w < (back ? n_rare : n)) {
        const uint4 it = A.work[back ? A.work_cap - 1 - w : w];
        const float* dp = A.xyz + 3 * (size_t)it.x;
        const float* ap = A.xyz + 3 * (size_t)it.y;
        const float4 pd = make_float4(dp[0], dp[1], dp[2], __int_as_float((int)it.x));
        const float4 pa = make_float4(ap[0], ap[1], ap[2], __int_as_float((int)it.y));
        const uint32_t kind = it.w & 7u;
        const double vdw_a = A.side.vdw[it.w >> 8];
        uint32_t bits = 0;
        if (kind <= 3u) {
            const int2 hr = A.side.h_off ? make_int2(A.side.h_off[it.x], A.side.h_off[it.x + 1]) : make_int2(0, 0);
            const int got = rule_hbond_scan_range(A.side, P, hr.x, hr.y, pd.x, pd.y, pd.z, pa.x, pa.y, pa.z, vdw_a, (int)kind);
            if (got & ARP_HB_NEED_H) bits |= 1u << ARP_SIFT_HBOND;
            if (got & ARP_HB_NEED_W) bits |= 1u << ARP_SIFT_WEAK_HBOND;
        } else if (kind == CLS_KIND_HAL) {
            if (rule_is_halogen_weak_hbond(A.side, P, __float_as_int(pd.w), __float_as_int(pa.w), pa.x, pa.y, pa.z,
                                           A.side.feat[__float_as_int(pa.w)], vdw_a)) bits = 1u << ARP_SIFT_WEAK_HBOND;
        } else {
            uint32_t fault = 0;
            if (rule_is_xbond(A.side, P, __float_as_int(pd.w), pd.x, pd.y, pd.z, pa.x, pa.y, pa.z,
                              A.side.feat[__float_as_int(pd.w)], &fault)) bits = 1u << ARP_SIFT_XBOND;
            bits |= fault;
        }
        if (bits) atomicOr(&A.out[it.z].mask, bits);
        }
    }
    if (A.clean_cnt) {                                           /* the grid build is long over: leave its counts clean */
        const unsigned nc = A.meta->n_cells + 1u;
        for (unsigned k = blockIdx.x * (HSCAN_WARPS * 32) + threadIdx.x; k < nc; k += gridDim.x * (HSCAN_WARPS * 32))
            A.clean_cnt[k] = 0;
    }
#ifdef PAIR_PROFILE
    __syncthreads();
    PROF_STAMP(3, 1);
#endif
}

#include "arp_tiles.cuh"

/* longest donor-hydrogen distance of the upload (float, rounded up; +inf when a distance is not finite).
   *reach = upload generation << 32 | float bits: non-negative floats order like their bit patterns and a newer
   generation beats every older value, so the word never needs resetting */
__global__ void __launch_bounds__(256) k_hreach(int N, const float* __restrict__ xyz, const int32_t* __restrict__ h_off,
                                                const double* __restrict__ h_xyz, unsigned long long* __restrict__ reach,
                                                unsigned gen)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = -1.f;
    if (i < N) {
        const double x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
        for (int k = h_off[i]; k < h_off[i + 1]; ++k) {
            const double dx = h_xyz[3 * (size_t)k] - x, dy = h_xyz[3 * (size_t)k + 1] - y, dz = h_xyz[3 * (size_t)k + 2] - z;
            const double d = sqrt(dx * dx + dy * dy + dz * dz);
            float f = __double2float_ru(d);
            if (!(f < 3.0e38f)) f = __int_as_float(0x7f800000);
            m = fmaxf(m, f);
        }
    }
    int key = m < 0.f ? -1 : __float_as_int(m);
    key = __reduce_max_sync(FULL, key);
    if ((threadIdx.x & 31) == 0 && key >= 0) atomicMax(reach, ((unsigned long long)gen << 32) | (unsigned)key);
}

/* K x K table of the float32 proximity thresholds (interactions.py:717-718, :760-768): NumPy narrows
   the python-float sums to float32 before comparing with the float32 distance (NEP 50) */
__global__ void k_radtab(int K, const double* __restrict__ vdw, const double* __restrict__ cov, double comp,
                         float4* __restrict__ tab)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= K * K) return;
    int a = t / K, b = t % K;
    double sum_cov = d_add(cov[a], cov[b]);
    double sum_vdw = d_add(vdw[a], vdw[b]);
    tab[t] = make_float4((float)sum_cov, (float)sum_vdw, (float)d_add(sum_vdw, comp), 0.f);
}

/* ---- host side ------------------------------------------------------------------------------- */

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int arp_pairs_prepare(arp_ctx* c)
{
    const size_t N = (size_t)c->N, S = (size_t)c->S;
    c->cell_bound = 4 * N + 64 * S;
    size_t tiles = (c->cell_bound + 1 + ARP_SCAN_TILE - 1) / ARP_SCAN_TILE + 1;
    c->off_bbox = align_up(sizeof(RunMeta), 256);
    c->off_cnt = align_up(c->off_bbox + 6 * sizeof(unsigned) * S, 256);
    c->off_state = align_up(c->off_cnt + sizeof(int) * (c->cell_bound + 2), 256);
    c->zero_bytes = c->off_state + tiles * sizeof(unsigned long long);
    ARP_TRY(dbuf_reserve(c, c->zero, c->zero_bytes));
    ARP_TRY(dbuf_reserve(c, c->geom, sizeof(StructGeom) * S));
    ARP_TRY(dbuf_reserve(c, c->cell_start, sizeof(int) * (c->cell_bound + 2)));
    if (c->use_tiles) ARP_TRY(dbuf_reserve(c, c->runtab, sizeof(int2) * 6 * (c->cell_bound + 2)));
    ARP_TRY(dbuf_reserve(c, c->cell_of, sizeof(int) * N));
    ARP_TRY(dbuf_reserve(c, c->rank, sizeof(int) * N));
    ARP_TRY(dbuf_reserve(c, c->pos4, sizeof(float4) * N * (c->use_tiles ? 2 : 1)));
    ARP_TRY(dbuf_reserve(c, c->att4, sizeof(uint4) * N));
    /* longest donor-hydrogen distance: a property of the uploaded atoms, computed once per upload */
    if (!c->hreach.p) {
        ARP_TRY(dbuf_reserve(c, c->hreach, 16));
        ARP_CUDA(c, cudaMemsetAsync(c->hreach.p, 0, 16, c->stream));
    }
    c->upload_gen++;
    if (c->has_h && N > 0 && c->H > 0) {
        k_hreach<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>((int)N, c->xyz.as<float>(), c->h_off.as<int32_t>(),
                                                                      c->h_xyz.as<double>(), c->hreach.as<unsigned long long>(),
                                                                      c->upload_gen);
        ARP_LAUNCHED(c);
    }
    return ARP_OK;
}

int arp_pairs_enqueue(arp_ctx* c, int with_events)
{
    const int N = c->N, S = c->S;
    char* z = c->zero.as<char>();
    RunMeta* meta = (RunMeta*)z;
    unsigned* bbox = (unsigned*)(z + c->off_bbox);
    int* cell_cnt = (int*)(z + c->off_cnt);
    unsigned long long* state = (unsigned long long*)(z + c->off_state);
    const int* so = S > 1 ? c->struct_off.as<int>() : nullptr;
    cudaStream_t st = c->stream;

    /* K x K float32 proximity thresholds: once per (atoms, params); ahead of the run so that the pair
       kernels follow each other directly (programmatic dependent launch) */
    if (N > 0 && !c->radtab_valid) {
        ARP_TRY(dbuf_reserve(c, c->radtab, sizeof(float4) * (size_t)c->K * c->K));
        k_radtab<<<(unsigned)((c->K * c->K + 127) / 128), 128, 0, st>>>(c->K, c->vdw.as<double>(), c->cov.as<double>(),
                                                                       c->params.vdw_comp, c->radtab.as<float4>());
        ARP_LAUNCHED(c);
        c->radtab_valid = 1;
    }
    c->events_level = with_events;
    const bool grid_event = with_events >= 2;         /* events between the kernels keep them from overlapping */
    const bool split_events = with_events >= 3;
    /* k_classify starts while k_search drains when that drain is a visible part of the job; the price is one
       8-byte store per candidate (the list is left all zero), which costs more than it gains beyond ~5 * 10^5
       atoms.  The diagnostic timing with events between the kernels measures the kernels one after the other.
       A list that an earlier run without the early start left dirty is zeroed first (state repair, not part of
       this run: ahead of the first event). */
    const int early = !c->use_tiles && c->use_early_cls && N <= 500000 && !split_events;
    if (early && c->hits_dirty && c->hits.p) {
        ARP_CUDA(c, cudaMemsetAsync(c->hits.p, 0, c->hits.cap, st));
        c->hits_dirty = 0;
    }
    if (with_events) ARP_CUDA(c, cudaEventRecord(c->ev[0], st));
    /* Which grid build?  The register kernel needs no memset when the previous run on this layout of the zero
       region was one of its own (it zeroes its counters itself and every run leaves the rest clean). */
    bool reg_path = false;
    int reg_apt = 1;
    if (N > 0) {
        if (c->coop_blocks < 0) {           /* once per context: can the grid build run as one cooperative kernel? */
            int coop = 0, per_sm = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
            if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_grid_fused, GRID_THREADS, 0) == cudaSuccess)
                c->coop_blocks = per_sm * c->sm_count;
            else
                c->coop_blocks = 0;
            (void)cudaGetLastError();
        }
        if (c->coop_blocks > 0 && c->use_fused_grid >= 2) {
            if (c->reg_blocks < 0) {        /* one 1024-thread block per SM with the shared-memory cell table, co-resident? */
                int per_sm = 0;
                c->reg_blocks = 0;
                const size_t smem = REG_SMEM_BYTES;
                if (cudaFuncSetAttribute(k_grid_reg<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                    cudaFuncSetAttribute(k_grid_reg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_grid_reg<REG_APT>, REG_THREADS, smem) == cudaSuccess && per_sm >= 1)
                    c->reg_blocks = c->sm_count;
                (void)cudaGetLastError();
            }
            const long long reg_threads = (long long)c->reg_blocks * REG_THREADS;
            if (reg_threads > 0 && (long long)N <= reg_threads * REG_APT) {
                reg_path = true;
                reg_apt = (long long)N <= reg_threads ? 1 : 2;
            }
        }
    }
    const bool clean = c->zero_clean && c->clean_ptr == (void*)z && c->clean_bytes == c->zero_bytes &&
                       c->clean_off_cnt == c->off_cnt && c->clean_off_state == c->off_state;
    c->zero_clean = 0;
    if (!(reg_path && clean)) ARP_CUDA(c, cudaMemsetAsync(z, 0, c->zero_bytes, st));
    if (N > 0) {
        ScatterArgs SC;
        SC.xyz = c->xyz.as<float>(); SC.feat = c->feat.as<uint32_t>(); SC.res_id = c->res_id.as<int32_t>();
        SC.rad_class = c->rad_class.as<uint16_t>(); SC.res_prev = c->res_prev.as<int32_t>();
        SC.res_next = c->res_next.as<int32_t>(); SC.res_flags = c->res_flags.as<uint8_t>();
        SC.bond_off = c->has_bonds ? c->bond_off.as<int32_t>() : nullptr;
        SC.h_off = c->has_h ? c->h_off.as<int32_t>() : nullptr;
        SC.h_xyz = c->has_h && c->H > 0 ? c->h_xyz.as<double>() : nullptr;
        SC.cell_of = c->cell_of.as<int>(); SC.rank = c->rank.as<int>(); SC.cell_start = c->cell_start.as<int>();
        SC.pos4 = c->pos4.as<float4>(); SC.att4 = c->att4.as<uint4>();
        SC.arec = c->use_tiles ? c->pos4.as<uint4>() : nullptr;      /* the records take the place of pos4 (sized for both layouts) */
        unsigned blocks = (unsigned)((N + GRID_THREADS - 1) / GRID_THREADS);
        if (c->coop_blocks > 0 && c->use_fused_grid) {
            GridArgs GA;
            GA.runtab = c->use_tiles ? c->runtab.as<int2>() : nullptr; GA.N = N; GA.S = S; GA.tile_x = c->tile_x; GA.cutoff = c->params.interacting_cutoff; GA.struct_off = so; GA.bbox = bbox;
            GA.geom = c->geom.as<StructGeom>(); GA.meta = meta; GA.cell_cnt = cell_cnt; GA.scan_state = state;
            GA.sc = SC; GA.cell_of = c->cell_of.as<int>(); GA.rank = c->rank.as<int>(); GA.cell_start = c->cell_start.as<int>();
            void* args[] = { &GA };
            if (reg_path) {
                unsigned g = (unsigned)(((long long)N + (long long)reg_apt * REG_THREADS - 1) / ((long long)reg_apt * REG_THREADS));
                ARP_CUDA(c, cudaLaunchCooperativeKernel(reg_apt == 1 ? (void*)k_grid_reg<1> : (void*)k_grid_reg<2>, dim3(g),
                                                        dim3(REG_THREADS), args, REG_SMEM_BYTES, st));
                c->launches++;
            } else {
                unsigned cap_blocks = (unsigned)c->coop_blocks;
#ifdef GRID_FUSED_BLOCKS_PER_SM
                if (cap_blocks > (unsigned)(c->sm_count * GRID_FUSED_BLOCKS_PER_SM)) cap_blocks = (unsigned)(c->sm_count * GRID_FUSED_BLOCKS_PER_SM);
#endif
                unsigned g = blocks < cap_blocks ? blocks : cap_blocks;
                ARP_CUDA(c, cudaLaunchCooperativeKernel((void*)k_grid_fused, dim3(g), dim3(GRID_THREADS), args, 0, st));
                c->launches++;
            }
        } else {
            k_bbox<<<(unsigned)((N + BBOX_ATOMS_PER_VB - 1) / BBOX_ATOMS_PER_VB), GRID_THREADS, 0, st>>>(c->xyz.as<float>(), so, S, N, bbox);
            ARP_LAUNCHED(c);
            k_geom<<<1, GRID_THREADS, 0, st>>>(bbox, so, S, N, c->params.interacting_cutoff, c->tile_x, c->geom.as<StructGeom>(), meta);
            ARP_LAUNCHED(c);
            k_cellid<<<blocks, GRID_THREADS, 0, st>>>(c->xyz.as<float>(), c->feat.as<uint32_t>(), so, S, N, c->geom.as<StructGeom>(), cell_cnt,
                                                      c->cell_of.as<int>(), c->rank.as<int>());
            ARP_LAUNCHED(c);
            ARP_TRY(arp_scan_exclusive(c, cell_cnt, c->cell_start.as<int>(), state, &meta->ticket_scan, &meta->n_cells, 1,
                                       c->cell_bound + 1));
            k_scatter<<<blocks, GRID_THREADS, 0, st>>>(N, SC);
            ARP_LAUNCHED(c);
            if (c->use_tiles) {
                k_runtab<<<(unsigned)(c->sm_count * 4), 256, 0, st>>>(c->geom.as<StructGeom>(), S, meta, c->cell_start.as<int>(), c->runtab.as<int2>());
                ARP_LAUNCHED(c);
            }
        }
    }
    if (grid_event) ARP_CUDA(c, cudaEventRecord(c->ev[1], st));
    if (N > 0) {
        ArpSide side;
        memset(&side, 0, sizeof side);
        side.vdw = c->vdw.as<double>(); side.cov = c->cov.as<double>();
        side.K = c->K;
        side.feat = c->feat.as<uint32_t>();
        side.radtab = c->radtab.as<float4>();
        side.bond_off = c->has_bonds ? c->bond_off.as<int32_t>() : nullptr;
        side.bond_nbr = c->has_bonds ? c->bond_nbr.as<int32_t>() : nullptr;
        side.h_off = c->has_h ? c->h_off.as<int32_t>() : nullptr;
        side.h_xyz = c->has_h ? c->h_xyz.as<double>() : nullptr;
        side.xnbr = c->has_xnbr ? c->xnbr.as<float>() : nullptr;
        side.hlim = nullptr;                 /* k_classify builds it in shared memory */

        const bool pdl = c->use_pdl != 0;
        if (!c->cls_smem_set) {             /* per device: > 48 KB of dynamic shared memory is opt-in */
            ARP_CUDA(c, cudaFuncSetAttribute(k_classify<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CLS_SMEM));
            ARP_CUDA(c, cudaFuncSetAttribute(k_classify<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CLS_SMEM));
            c->cls_smem_set = 1;
        }
        if (!c->hscan_blocks) {             /* a persistent grid: exactly the blocks that are resident together */
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hscan, HSCAN_WARPS * 32, 0) != cudaSuccess || per_sm < 1)
                per_sm = 1;
            (void)cudaGetLastError();
            c->hscan_blocks = per_sm * c->sm_count;
        }
        if (!early && !c->use_tiles) c->hits_dirty = 1;

        const bool tiles = c->use_tiles != 0;
        if (tiles) {
            /* search + classify in one kernel from shared-memory tiles (arp_tiles.cuh) */
            if (!c->tiles_blocks) {             /* a persistent grid: exactly the blocks that are resident together */
                ARP_CUDA(c, cudaFuncSetAttribute(k_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM));
                int per_sm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tiles, TW_THREADS, TW_SMEM) != cudaSuccess || per_sm < 1)
                    per_sm = 1;
                (void)cudaGetLastError();
                c->tiles_blocks = per_sm * c->sm_count;
            }
            if (split_events) ARP_CUDA(c, cudaEventRecord(c->ev[2], st));      /* no separate search kernel */
            TileArgs TA;
            TA.arec = c->pos4.as<uint4>(); TA.runtab = c->runtab.as<int2>();
            TA.meta = meta; TA.out = c->out.as<arp_pair>(); TA.cap = c->out_cap;
            TA.work = c->work.as<uint4>(); TA.work_cap = c->work_cap;
            TA.h_reach = c->hreach.as<unsigned long long>(); TA.h_gen = c->upload_gen;
            TA.r2 = c->rp.r2; TA.include_seq_adjacent = c->rp.include_seq_adjacent; TA.side = side;
            /* about one warp per two cells of ~6 atoms on small inputs */
            unsigned tgrid = (unsigned)c->tiles_blocks;
            size_t want = (size_t)N / (12 * TW_WARPS) + 1;
            if (want < tgrid) tgrid = (unsigned)want;
            ARP_CUDA(c, launch_k(k_tiles, tgrid, TW_THREADS, TW_SMEM, st, pdl, TA, c->rp));
            c->launches++;
            if (split_events) ARP_CUDA(c, cudaEventRecord(c->ev[4], st));
        } else {
            SearchArgs SA;
            SA.pos4 = c->pos4.as<float4>(); SA.cell_start = c->cell_start.as<int>();
            SA.geom = c->geom.as<StructGeom>(); SA.meta = meta; SA.raw = c->hits.as<uint2>(); SA.cap = c->out_cap;
            unsigned grid = (unsigned)(c->sm_count * SEARCH_GRID_MULT);
    #ifndef SEARCH_ATOMS_PER_CELL
    #define SEARCH_ATOMS_PER_CELL 5      /* 5.7 at protein density; a block too many exits at once, one too few leaves tickets to the counter */
    #endif
            size_t want = ((size_t)N / (SEARCH_ATOMS_PER_CELL * SEARCH_CELLS)) / SEARCH_WARPS + 1;     /* about one warp per ticket on small inputs */
            if (want < grid) grid = (unsigned)want;
            ARP_CUDA(c, launch_k(k_search, grid, SEARCH_WARPS * 32, 0, st, pdl, SA));
            c->launches++;
            if (split_events) ARP_CUDA(c, cudaEventRecord(c->ev[2], st));

            ClassifyArgs CA;
            CA.h_reach = c->hreach.as<unsigned long long>(); CA.h_gen = c->upload_gen;
            CA.pos4 = SA.pos4; CA.att4 = c->att4.as<uint4>(); CA.raw = SA.raw; CA.cap = c->out_cap; CA.meta = meta;
            CA.out = c->out.as<arp_pair>(); CA.side = side;
            CA.work = c->work.as<uint4>(); CA.work_cap = c->work_cap;
            CA.r2 = c->rp.r2; CA.include_seq_adjacent = c->rp.include_seq_adjacent;
            size_t tiles = (size_t)((c->out_cap + CLS_TILE - 1) / CLS_TILE);
            size_t blocks_needed = (tiles + CLS_WARPS - 1) / CLS_WARPS;
            unsigned cgrid = (unsigned)(c->sm_count * CLS_MINB);
            if (blocks_needed < cgrid) cgrid = (unsigned)(blocks_needed ? blocks_needed : 1);
            ARP_CUDA(c, launch_k(early ? k_classify<true> : k_classify<false>, cgrid, CLS_WARPS * 32, CLS_SMEM, st, pdl, CA, c->rp));
            c->launches++;
            if (split_events) ARP_CUDA(c, cudaEventRecord(c->ev[4], st));
        }

        HscanArgs HA;
        HA.work = c->work.as<uint4>(); HA.meta = meta; HA.work_cap = c->work_cap;
        HA.out = c->out.as<arp_pair>(); HA.side = side; HA.clean_cnt = cell_cnt; HA.xyz = c->xyz.as<float>();
        size_t hb = (size_t)((c->work_cap + 255) / 256);
        unsigned hgrid = (unsigned)c->hscan_blocks;
        if (hb < hgrid) hgrid = (unsigned)(hb ? hb : 1);
        ARP_CUDA(c, launch_k(k_hscan, hgrid, HSCAN_WARPS * 32, 0, st, pdl, HA, c->rp));
        c->launches++;
    } else if (split_events) {
        ARP_CUDA(c, cudaEventRecord(c->ev[2], st));
        ARP_CUDA(c, cudaEventRecord(c->ev[4], st));
    }
    if (with_events) ARP_CUDA(c, cudaEventRecord(c->ev[3], st));
    ARP_CUDA(c, cudaMemcpyAsync(c->h_meta, meta, sizeof(RunMeta), cudaMemcpyDeviceToHost, st));
#ifdef PAIR_PROFILE
    {
        static unsigned long long h[4][2048][2];
        static int runs = 0;
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_prof, sizeof h);
        if (++runs % 8 == 0) {
            unsigned long long t0 = ~0ull;
            for (int b = 0; b < 2048; ++b) if (h[0][b][0] && h[0][b][0] < t0) t0 = h[0][b][0];
            const char* nm[4] = { "grid", "search", "classify", "hscan" };
            for (int k = 0; k < 4; ++k) {
                unsigned long long s0 = ~0ull, s1 = 0, e0 = ~0ull, e1 = 0; double es = 0; int nb = 0;
                unsigned long long ends[2048];
                for (int b = 0; b < 2048; ++b) if (h[k][b][0] >= t0 && h[k][b][1] >= h[k][b][0]) {
                    s0 = h[k][b][0] < s0 ? h[k][b][0] : s0; s1 = h[k][b][0] > s1 ? h[k][b][0] : s1;
                    e0 = h[k][b][1] < e0 ? h[k][b][1] : e0; e1 = h[k][b][1] > e1 ? h[k][b][1] : e1;
                    es += (double)(h[k][b][1] - t0); ends[nb++] = h[k][b][1] - t0;
                }
                if (!nb) continue;
                for (int a = 1; a < nb; ++a) { unsigned long long v = ends[a]; int b = a - 1; while (b >= 0 && ends[b] > v) { ends[b + 1] = ends[b]; --b; } ends[b + 1] = v; }
                fprintf(stderr, "%-8s blocks %4d | first start %6llu last start %6llu | first end %6llu median end %6llu p90 end %6llu last end %6llu (ns after the grid kernel's first block)\n",
                        nm[k], nb, s0 - t0, s1 - t0, e0 - t0, ends[nb / 2], ends[nb * 9 / 10], e1 - t0);
            }
        }
    }
#endif
    if (reg_path) {                         /* everything enqueued: this layout of the zero region is clean after the run */
        c->zero_clean = 1; c->clean_ptr = (void*)z; c->clean_bytes = c->zero_bytes;
        c->clean_off_cnt = c->off_cnt; c->clean_off_state = c->off_state;
    }
    return ARP_OK;
}

/* ---- canonical (i, j) order of the record stream -----------------------------------------
 * The kernels take the record count either from the host (n_dev == NULL: n_max records) or from the run's own counter on
 * the device (n_dev = &meta->n_pairs, clamped to the capacity n_max), so that the view can be enqueued behind a run that
 * has not been waited for (arp_pairs_fetch_packed_async). */
__device__ __forceinline__ unsigned long long sort_n(const unsigned long long* n_dev, unsigned long long n_max)
{
    if (!n_dev) return n_max;
    const unsigned long long n = *n_dev;
    return n < n_max ? n : n_max;
}

__global__ void __launch_bounds__(256) k_sort_count(const arp_pair* __restrict__ rec, const unsigned long long* __restrict__ n_dev,
                                                    unsigned long long n_max, int* __restrict__ cnt, unsigned* __restrict__ fault_cnt)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *fault_cnt = 0u;          /* counted by k_sort_place<2>, two kernels later */
    const unsigned long long n = sort_n(n_dev, n_max), step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += step) atomicAdd(&cnt[rec[r].i], 1);
}

__global__ void __launch_bounds__(256) k_sort_scatter(const arp_pair* __restrict__ rec, const unsigned long long* __restrict__ n_dev,
                                                      unsigned long long n_max, const int* __restrict__ off, int* __restrict__ cur,
                                                      arp_pair* __restrict__ tmp)
{
    const unsigned long long n = sort_n(n_dev, n_max), step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += step) {
        int4 v = reinterpret_cast<const int4*>(rec)[r];
        int pos = off[v.x] + atomicAdd(&cur[v.x], 1);
        reinterpret_cast<int4*>(tmp)[pos] = v;
    }
}

/* records of one i are contiguous in tmp; (i, j) is unique, so the rank of j inside the segment
   is the final position.  In the compact and packed views i is implied by the row offsets `off` and the distances
   leave as a stream of their own (arp_pairs_fetch_compact / arp_pairs_fetch_packed). */
/* MODE 0: 16-byte records; 1: the compact view, 8-byte records (j, mask) + distance stream; 2: the PACKED view, the 15 SIFt
   bits above the bits_j bits of j in one word of 32 (+ 8) bits per record + distance stream.  The packed word carries
   neither the entity class (a function of the two atoms' selection / water flags, which the host has) nor the
   xbond-without-neighbour fault bit: records that have it are counted in *n_fault instead. */
template <int MODE>
__global__ void __launch_bounds__(256) k_sort_place(const arp_pair* __restrict__ tmp, const unsigned long long* __restrict__ n_dev,
                                                    unsigned long long n_max,
                                                    const int* __restrict__ off, arp_pair* __restrict__ out,
                                                    arp_pair_c* __restrict__ outc, float* __restrict__ outd,
                                                    uint32_t* __restrict__ lo32, uint8_t* __restrict__ hi8, int bits_j,
                                                    unsigned* __restrict__ n_fault, const int* __restrict__ struct_off, int S)
{
    const unsigned long long n = sort_n(n_dev, n_max), step = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += step) {
        int4 v = reinterpret_cast<const int4*>(tmp)[r];
        int b = off[v.x], e = off[v.x + 1];
        int rank = 0;
        for (int k = b; k < e; ++k) rank += tmp[k].j < v.y ? 1 : 0;
        if (MODE == 1) {
            reinterpret_cast<int2*>(outc)[b + rank] = make_int2(v.y, v.z);
            outd[b + rank] = __int_as_float(v.w);
        } else if (MODE == 2) {
            int jl = v.y;
            if (S > 1) {                        /* a batch: j local to the structure of row i (no pair spans two structures) */
                int lo = 0, hi = S;             /* last structure whose first atom is <= i */
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (struct_off[mid] <= v.x) lo = mid; else hi = mid; }
                jl -= struct_off[lo];
            }
            const unsigned long long w = (unsigned long long)(unsigned)jl | ((unsigned long long)((unsigned)v.z & 0x7fffu) << bits_j);
            lo32[b + rank] = (uint32_t)w;
            if (hi8) hi8[b + rank] = (uint8_t)(w >> 32);
            outd[b + rank] = __int_as_float(v.w);
            if ((unsigned)v.z & ARPK_FAULT_XBOND_NO_NBR) atomicAdd(n_fault, 1u);
        } else {
            reinterpret_cast<int4*>(out)[b + rank] = v;
        }
    }
}

/* bits of an end-atom index in the packed view: indices are local to the structure, so the largest structure decides */
int arp_pairs_bits_j(const arp_ctx* c)
{
    int bits = 1;
    while (bits < 31 && (1ll << bits) < (long long)c->max_struct_atoms) ++bits;
    return bits;
}

/* view 0: sort_out (arp_pair[n]); 1: sort_c (arp_pair_c[n]) + sort_d (float[n]); 2: sort_lo (u32[n]) [+ sort_hi (u8[n])] + sort_d;
   all with sort_off (row offsets).  blind: the run has not been waited for -- the kernels read the record count on the
   device and the buffers are sized for the capacity of the record stream; the caller marks the view valid once the run
   turned out complete. */
int arp_pairs_sorted_build(arp_ctx* c, int view, int blind)
{
    int& valid = view == 0 ? c->sorted_valid : view == 1 ? c->compact_valid : c->packed_valid;
    if (valid && !blind) return ARP_OK;
    const unsigned long long n = blind ? c->out_cap : c->n_pairs;          /* bound of the record count */
    const size_t N = (size_t)c->N;
    if (n >= (1ull << 31)) return arp_fail(c, ARP_E_CAPACITY, "too many records for the sorted view", __FILE__, __LINE__);
    const int bits_j = arp_pairs_bits_j(c);
    const bool need_hi = bits_j + 15 > 32;
    if (view == 1) {
        ARP_TRY(dbuf_reserve(c, c->sort_c, sizeof(arp_pair_c) * (size_t)n));
        ARP_TRY(dbuf_reserve(c, c->sort_d, sizeof(float) * (size_t)n));
    } else if (view == 2) {
        ARP_TRY(dbuf_reserve(c, c->sort_lo, sizeof(uint32_t) * (size_t)n));
        if (need_hi) ARP_TRY(dbuf_reserve(c, c->sort_hi, (size_t)n));
        ARP_TRY(dbuf_reserve(c, c->sort_d, sizeof(float) * (size_t)n));
    } else {
        ARP_TRY(dbuf_reserve(c, c->sort_out, sizeof(arp_pair) * (size_t)n));
    }
    ARP_TRY(dbuf_reserve(c, c->sort_off, sizeof(int) * (N + 2)));
    /* zero region: cnt[N+1] | cur[N+1] | ticket | scan state.  The fault counter of the packed view is the spare entry
       behind the row offsets, sort_off[N + 1], so that one copy fetches both. */
    const size_t tiles = (N + 1 + ARP_SCAN_TILE - 1) / ARP_SCAN_TILE + 1;
    const size_t o_cur = align_up(sizeof(int) * (N + 2), 256);
    const size_t o_tick = align_up(o_cur + sizeof(int) * (N + 2), 256);
    const size_t o_state = o_tick + 256;
    const size_t zb = o_state + tiles * sizeof(unsigned long long);
    ARP_TRY(dbuf_reserve(c, c->sort_zero, zb));
    char* z = c->sort_zero.as<char>();
    c->sort_fault = c->sort_off.as<unsigned>() + (N + 1);
    if (n == 0) {
        ARP_CUDA(c, cudaMemsetAsync(c->sort_off.p, 0, sizeof(int) * (N + 2), c->stream));
        valid = 1;
        return ARP_OK;
    }
    const unsigned long long* n_dev = blind ? &((const RunMeta*)c->zero.p)->n_pairs : nullptr;
    size_t want = (size_t)((n + 255) / 256), most = (size_t)c->sm_count * 16;
    const unsigned blocks = (unsigned)(want < most ? want : most);
    if (!c->sort_tmp_valid || blind) {           /* records grouped by i (tmp) + row offsets: shared by the views */
        ARP_TRY(dbuf_reserve(c, c->sort_tmp, sizeof(arp_pair) * (size_t)n));
        int* cnt = (int*)z; int* cur = (int*)(z + o_cur);
        unsigned* ticket = (unsigned*)(z + o_tick);
        unsigned long long* state = (unsigned long long*)(z + o_state);
        ARP_CUDA(c, cudaMemsetAsync(z, 0, zb, c->stream));
        k_sort_count<<<blocks, 256, 0, c->stream>>>(c->out.as<arp_pair>(), n_dev, n, cnt, c->sort_fault);
        ARP_LAUNCHED(c);
        ARP_TRY(arp_scan_exclusive(c, cnt, c->sort_off.as<int>(), state, ticket, nullptr, (int)(N + 1), N + 1));
        k_sort_scatter<<<blocks, 256, 0, c->stream>>>(c->out.as<arp_pair>(), n_dev, n, c->sort_off.as<int>(), cur,
                                                      c->sort_tmp.as<arp_pair>());
        ARP_LAUNCHED(c);
        if (!blind) c->sort_tmp_valid = 1;
    } else if (view == 2) {
        ARP_CUDA(c, cudaMemsetAsync(c->sort_fault, 0, sizeof(unsigned), c->stream));
    }
    const arp_pair* tmp = c->sort_tmp.as<arp_pair>();
    const int* off = c->sort_off.as<int>();
    if (view == 1)
        k_sort_place<1><<<blocks, 256, 0, c->stream>>>(tmp, n_dev, n, off, nullptr, c->sort_c.as<arp_pair_c>(), c->sort_d.as<float>(), nullptr, nullptr, 0, nullptr, nullptr, 1);
    else if (view == 2)
        k_sort_place<2><<<blocks, 256, 0, c->stream>>>(tmp, n_dev, n, off, nullptr, nullptr, c->sort_d.as<float>(), c->sort_lo.as<uint32_t>(),
                                                       need_hi ? c->sort_hi.as<uint8_t>() : nullptr, bits_j, c->sort_fault,
                                                       c->S > 1 ? c->struct_off.as<int>() : nullptr, c->S);
    else
        k_sort_place<0><<<blocks, 256, 0, c->stream>>>(tmp, n_dev, n, off, c->sort_out.as<arp_pair>(), nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 1);
    ARP_LAUNCHED(c);
    if (!blind) valid = 1;
    return ARP_OK;
}

/* ---- binding-site expansion (interactions.py:1420-1424) ------------------------------------
 * flag[i] = in selection, or within `radius` of a selected atom of the same structure
 * (Bio.PDB.kdtrees test: double, d2 <= r*r).  Selected atoms are few (a ligand), so each
 * block stages a tile of selected atoms in shared memory and every thread tests its atom. */
__global__ void __launch_bounds__(256) k_within_collect(int N, const uint32_t* __restrict__ feat, int* __restrict__ n_sel,
                                                        int* __restrict__ sel)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && (feat[i] & ARP_F_IN_SELECTION)) sel[atomicAdd(n_sel, 1)] = i;
}

__global__ void __launch_bounds__(256) k_within(int N, int S, const float* __restrict__ xyz, const uint32_t* __restrict__ feat,
                                                const int* __restrict__ struct_off, const int* __restrict__ n_sel_p,
                                                const int* __restrict__ sel, double r2,
                                                uint8_t* __restrict__ flags)
{
    __shared__ float4 s_sel[256];
    __shared__ int s_struct[256];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_sel = *n_sel_p;
    bool live = i < N;
    float x = 0.f, y = 0.f, z = 0.f;
    int s = 0;
    bool flag = false;
    if (live) {
        x = xyz[3 * (size_t)i]; y = xyz[3 * (size_t)i + 1]; z = xyz[3 * (size_t)i + 2];
        s = struct_of(struct_off, S, i);
        flag = (feat[i] & ARP_F_IN_SELECTION) != 0;
    }
    for (int t0 = 0; t0 < n_sel; t0 += 256) {
        int k = t0 + threadIdx.x;
        __syncthreads();
        if (k < n_sel) {
            int a = sel[k];
            s_sel[threadIdx.x] = make_float4(xyz[3 * (size_t)a], xyz[3 * (size_t)a + 1], xyz[3 * (size_t)a + 2], 0.f);
            s_struct[threadIdx.x] = struct_of(struct_off, S, a);
        }
        __syncthreads();
        int m = min(256, n_sel - t0);
        if (live && !flag) {
            for (int j = 0; j < m; ++j) {
                float4 p = s_sel[j];
                if (s_struct[j] == s && kd_within(p.x, p.y, p.z, x, y, z, r2)) { flag = true; break; }
            }
        }
    }
    if (live) flags[i] = flag ? 1 : 0;
}

/* ---- the same through a cell grid (large inputs) -------------------------------------------------
 * The double loop above costs N x n_sel tests; for a large selection (a chain, several ligands) the selected
 * atoms are binned into cells of edge >= radius instead and every other atom visits the 27 cells around it.
 * Cell coordinates are clamped to the grid (monotone, so two atoms within one edge of each other always land in
 * neighbouring cells, whatever the box); a batch shares one grid, the structure index travels with the point. */
struct WithinGrid {
    double ox, oy, oz, inv_w;
    int dx, dy, dz, ncell;
};

__device__ __forceinline__ void wg_make(const unsigned* __restrict__ bb, int n, double radius, WithinGrid& g)
{
    double mn[3], mx[3], amax = 0.0;
    for (int k = 0; k < 3; ++k) {
        const unsigned lo = bb[k], hi = bb[3 + k];
        mn[k] = hi ? (double)ord2f(~lo) : 0.0;
        mx[k] = hi ? (double)ord2f(hi) : 0.0;
        amax = fmax(amax, fmax(fabs(mn[k]), fabs(mx[k])));
    }
    double w = (radius > 1e-3 ? radius * (1.0 + 1e-6) + 1e-6 : 1e-3) + 2e-16 * amax;
    long long d[3];
    for (;;) {
        for (int k = 0; k < 3; ++k) d[k] = (long long)floor((mx[k] - mn[k]) / w) + 1;
        if ((double)d[0] * (double)d[1] * (double)d[2] <= 4.0 * (double)n + 64.0) break;
        w *= 1.5;
    }
    g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2]; g.inv_w = 1.0 / w;
    g.dx = (int)d[0]; g.dy = (int)d[1]; g.dz = (int)d[2]; g.ncell = g.dx * g.dy * g.dz;
}

__global__ void __launch_bounds__(256) k_wg_bbox(int N, const float* __restrict__ xyz, unsigned* __restrict__ bbox)
{
    __shared__ unsigned s_v[8][6];
    unsigned v[6] = {0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
        if (fabsf(x) < 3e38f && fabsf(y) < 3e38f && fabsf(z) < 3e38f) {                 /* finite atoms only */
            const unsigned ox = f2ord(x), oy = f2ord(y), oz = f2ord(z);
            v[0] = max(v[0], ~ox); v[1] = max(v[1], ~oy); v[2] = max(v[2], ~oz);
            v[3] = max(v[3], ox); v[4] = max(v[4], oy); v[5] = max(v[5], oz);
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        v[k] = __reduce_max_sync(FULL, v[k]);
        if ((threadIdx.x & 31) == 0) s_v[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned m = 0;
        for (int w = 0; w < 8; ++w) m = max(m, s_v[w][threadIdx.x]);
        if (m) atomicMax(&bbox[threadIdx.x], m);
    }
}

/* binning of the selected atoms: cell + rank (count pass), then (x, y, z, structure) into cell order */
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_wg_bin(int N, int S, const float* __restrict__ xyz, const uint32_t* __restrict__ feat,
                                                const int* __restrict__ struct_off, const unsigned* __restrict__ bbox, double radius,
                                                int* __restrict__ cell_cnt, const int* __restrict__ cell_start,
                                                int* __restrict__ cell_of, int* __restrict__ rank, float4* __restrict__ spos,
                                                unsigned* __restrict__ n_cells)
{
    __shared__ WithinGrid s_g;
    if (threadIdx.x == 0) {
        wg_make(bbox, N, radius, s_g);
        if (!SCATTER && blockIdx.x == 0) *n_cells = (unsigned)s_g.ncell;
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        if (!(feat[i] & ARP_F_IN_SELECTION)) continue;
        const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
        if (!SCATTER) {
            const int c = (cell_coord((double)z, s_g.oz, s_g.inv_w, s_g.dz) * s_g.dy + cell_coord((double)y, s_g.oy, s_g.inv_w, s_g.dy)) * s_g.dx +
                          cell_coord((double)x, s_g.ox, s_g.inv_w, s_g.dx);
            cell_of[i] = c;
            rank[i] = atomicAdd(&cell_cnt[c], 1);
        } else {
            spos[cell_start[cell_of[i]] + rank[i]] = make_float4(x, y, z, __int_as_float(struct_of(struct_off, S, i)));
        }
    }
}

__global__ void __launch_bounds__(256) k_wg_query(int N, int S, const float* __restrict__ xyz, const uint32_t* __restrict__ feat,
                                                  const int* __restrict__ struct_off, const unsigned* __restrict__ bbox, double radius,
                                                  const int* __restrict__ cell_start, const float4* __restrict__ spos,
                                                  uint8_t* __restrict__ flags)
{
    __shared__ WithinGrid s_g;
    if (threadIdx.x == 0) wg_make(bbox, N, radius, s_g);
    __syncthreads();
    const double r2 = radius * radius;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        bool flag = (feat[i] & ARP_F_IN_SELECTION) != 0;
        if (!flag) {
            const float x = xyz[3 * (size_t)i], y = xyz[3 * (size_t)i + 1], z = xyz[3 * (size_t)i + 2];
            const int s = struct_of(struct_off, S, i);
            const int cx = cell_coord((double)x, s_g.ox, s_g.inv_w, s_g.dx), cy = cell_coord((double)y, s_g.oy, s_g.inv_w, s_g.dy),
                      cz = cell_coord((double)z, s_g.oz, s_g.inv_w, s_g.dz);
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, s_g.dx - 1);
            for (int k = 0; k < 9 && !flag; ++k) {
                const int yy = cy + k % 3 - 1, zz = cz + k / 3 - 1;
                if (yy < 0 || yy >= s_g.dy || zz < 0 || zz >= s_g.dz) continue;
                const int row = (zz * s_g.dy + yy) * s_g.dx;
                for (int q = cell_start[row + x0], e = cell_start[row + x1 + 1]; q < e; ++q) {
                    const float4 p = spos[q];
                    if (__float_as_int(p.w) == s && kd_within(p.x, p.y, p.z, x, y, z, r2)) { flag = true; break; }
                }
            }
        }
        flags[i] = flag ? 1 : 0;
    }
}

#ifndef WITHIN_GRID_MIN_ATOMS
#define WITHIN_GRID_MIN_ATOMS 20000
#endif

int arp_flag_within_run(arp_ctx* c, double radius)
{
    const int N = c->N;
    if (N >= WITHIN_GRID_MIN_ATOMS && c->use_within_grid) {
        /* zero region: bbox | n_cells | ticket | cell_cnt | scan state ; then cell_start, cell_of, rank, spos, flags */
        const size_t cells = 4 * (size_t)N + 66;
        const size_t tiles = (cells + ARP_SCAN_TILE - 1) / ARP_SCAN_TILE + 1;
        const size_t o_cnt = 256, o_state = align_up(o_cnt + cells * 4, 256), zero_bytes = align_up(o_state + tiles * 8, 256);
        const size_t o_start = zero_bytes, o_cellof = align_up(o_start + cells * 4, 256), o_rank = align_up(o_cellof + (size_t)N * 4, 256);
        const size_t o_spos = align_up(o_rank + (size_t)N * 4, 256), o_flags = align_up(o_spos + (size_t)N * 16, 256);
        ARP_TRY(dbuf_reserve(c, c->within, o_flags + (size_t)N));
        char* z = c->within.as<char>();
        unsigned* bbox = (unsigned*)z; unsigned* n_cells = (unsigned*)(z + 64); unsigned* ticket = (unsigned*)(z + 128);
        int* cell_cnt = (int*)(z + o_cnt); unsigned long long* state = (unsigned long long*)(z + o_state);
        int* cell_start = (int*)(z + o_start); int* cell_of = (int*)(z + o_cellof); int* rank = (int*)(z + o_rank);
        float4* spos = (float4*)(z + o_spos);
        c->within_flags_off = o_flags;
        const int* so = c->S > 1 ? c->struct_off.as<int>() : nullptr;
        ARP_CUDA(c, cudaMemsetAsync(z, 0, zero_bytes, c->stream));
        unsigned blocks = (unsigned)((N + 255) / 256);
        const unsigned cap = (unsigned)c->sm_count * 8;
        blocks = blocks < cap ? blocks : cap;
        k_wg_bbox<<<blocks, 256, 0, c->stream>>>(N, c->xyz.as<float>(), bbox);
        ARP_LAUNCHED(c);
        k_wg_bin<false><<<blocks, 256, 0, c->stream>>>(N, c->S, c->xyz.as<float>(), c->feat.as<uint32_t>(), so, bbox, radius, cell_cnt,
                                                       nullptr, cell_of, rank, nullptr, n_cells);
        ARP_LAUNCHED(c);
        ARP_TRY(arp_scan_exclusive(c, cell_cnt, cell_start, state, ticket, n_cells, 1, cells));
        k_wg_bin<true><<<blocks, 256, 0, c->stream>>>(N, c->S, c->xyz.as<float>(), c->feat.as<uint32_t>(), so, bbox, radius, nullptr,
                                                      cell_start, cell_of, rank, spos, nullptr);
        ARP_LAUNCHED(c);
        k_wg_query<<<blocks, 256, 0, c->stream>>>(N, c->S, c->xyz.as<float>(), c->feat.as<uint32_t>(), so, bbox, radius, cell_start, spos,
                                                  (uint8_t*)(z + o_flags));
        ARP_LAUNCHED(c);
        return ARP_OK;
    }
    ARP_TRY(dbuf_reserve(c, c->within, (size_t)N * 5 + 64));
    char* base = c->within.as<char>();
    int* n_sel = (int*)base;
    int* sel = (int*)(base + 16);
    uint8_t* flags = (uint8_t*)(base + 16 + (size_t)N * 4);
    c->within_flags_off = 16 + (size_t)N * 4;
    ARP_CUDA(c, cudaMemsetAsync(n_sel, 0, 16, c->stream));
    if (N == 0) return ARP_OK;
    unsigned blocks = (unsigned)((N + 255) / 256);
    k_within_collect<<<blocks, 256, 0, c->stream>>>(N, c->feat.as<uint32_t>(), n_sel, sel);
    ARP_LAUNCHED(c);
    double r2 = radius * radius;
    k_within<<<blocks, 256, 0, c->stream>>>(N, c->S, c->xyz.as<float>(), c->feat.as<uint32_t>(),
                                            c->S > 1 ? c->struct_off.as<int>() : nullptr, n_sel, sel, r2, flags);
    ARP_LAUNCHED(c);
    return ARP_OK;
}
