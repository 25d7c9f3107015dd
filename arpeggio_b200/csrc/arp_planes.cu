/*
 * arp_planes.cu -- ring / amide plane terms on the GPU (sm_100a).
 *
 *   arp_ring_ring_run    __calculate_plane_plane_contacts   interactions.py:1064-1194
 *   arp_atom_ring_run    __calculate_atom_plane_contacts    interactions.py:947-1062
 *   arp_amide_amide_run  __calculate_group_group_contacts   interactions.py:1217-1300
 *   arp_amide_ring_run   __calculate_group_plane_contacts   interactions.py:1302-1382
 *   with utils.group_angle / group_group_angle              utils.py:638-693
 *
 * Every term is a (row, column) double loop in the reference.  Here one warp owns a row and its
 * lanes stride over the columns; hits are compacted with a ballot, which keeps the columns of a
 * row in ascending order.  A counting pass, the single-pass scan of arp_pairs.cu and an emitting
 * pass give records sorted by (row, column) -- the order the reference's loops create them in.
 *
 * Arithmetic follows the dtype flows of SURVEY 8a: ring-ring, atom-ring and amide-ring in float64
 * (BLAS dot/norm model), amide-amide in float32.  arccos is never evaluated: the folded angle
 * |deg| <= 30/60/90 tests compare the cosine with the host-computed images in arp_params.
 */
#include <math.h>

#include "arp_ctx.cuh"

#define FULL 0xffffffffu

struct PlaneRec {            /* layout of arp_plane_pair and arp_atom_plane */
    int32_t  a, b;
    uint32_t code, pad;
    double   dist;
};
static_assert(sizeof(PlaneRec) == sizeof(arp_plane_pair) && sizeof(PlaneRec) == sizeof(arp_atom_plane), "record layout");

struct PlaneArgs {
    /* rings: float64 */
    int nr; const double* rc; const double* rn; const int32_t* rres; const uint32_t* rflags;
    /* amides: float32 */
    int na; const float* ac; const float* an; const int32_t* ares; const uint32_t* aflags;
    /* atoms */
    int n_atoms; const float* xyz; const uint32_t* feat; const int32_t* res_id;
    /* thresholds */
    int blas_fma;
    double ring_centroid, amide_centroid, atom_ring, met_sulphur;
    float  amide_centroid_f32;
    double split64, pos64[3], neg64[3];
    float  split32, pos32[3], neg32[3];
};

enum { KIND_RING_RING = 0, KIND_ATOM_RING = 1, KIND_AMIDE_AMIDE = 2, KIND_AMIDE_RING = 3 };

/* |fold(arccos(c))| <= bins[k] on the cosine; NaN angle -> false (every comparison with NaN is) */
__device__ __forceinline__ bool le64(const PlaneArgs& A, double c, int k)
{
    if (acos_is_nan(c)) return false;
    return c <= A.split64 ? c <= A.neg64[k] : c >= A.pos64[k];
}
__device__ __forceinline__ bool le32(const PlaneArgs& A, float c, int k)
{
    if (acos_is_nan_f(c)) return false;
    return c <= A.split32 ? c <= A.neg32[k] : c >= A.pos32[k];
}

/* interactions.py:1095-1108 (and :1253-1265, :1334-1346): four ifs, last true wins */
__device__ __forceinline__ uint32_t plane_class(uint32_t fa, uint32_t fb)
{
    bool sa = fa & ARP_P_IN_SELECTION, sb = fb & ARP_P_IN_SELECTION;
    bool pa = fa & ARP_P_IN_SELECTION_PLUS, pb = fb & ARP_P_IN_SELECTION_PLUS;
    uint32_t c = 7;
    if (!sa && !sb) c = ARP_CLASS_INTRA_NON_SELECTION;
    if (pa && pb) c = ARP_CLASS_INTRA_BINDING_SITE;
    if (sa && sb) c = ARP_CLASS_INTRA_SELECTION;
    if (sa != sb) c = ARP_CLASS_INTER;
    return c;
}

/* one visit (a -> b) of the ring double loop, interactions.py:1110-1155 */
__device__ bool ring_visit(const PlaneArgs& A, int a, int b, uint32_t* geom, double* dist)
{
    const double* ca = A.rc + 3 * (size_t)a; const double* cb = A.rc + 3 * (size_t)b;
    const double* na = A.rn + 3 * (size_t)a; const double* nb = A.rn + 3 * (size_t)b;
    const bool intra = A.rres[a] == A.rres[b];                                        /* :1091 */
    const double tx = d_sub(ca[0], cb[0]), ty = d_sub(ca[1], cb[1]), tz = d_sub(ca[2], cb[2]);
    const double distance = np_norm3_f64(tx, ty, tz, A.blas_fma);                     /* :1111 */
    if (distance > A.ring_centroid) return false;                                     /* :1113 */
    const double nna = np_norm3_f64(na[0], na[1], na[2], A.blas_fma);
    const double cd = d_div(np_dot3_f64(na[0], na[1], na[2], nb[0], nb[1], nb[2], A.blas_fma),
                            d_mul(nna, np_norm3_f64(nb[0], nb[1], nb[2], A.blas_fma)));
    const double ct = d_div(np_dot3_f64(na[0], na[1], na[2], tx, ty, tz, A.blas_fma), d_mul(nna, distance));
    const bool d0 = le64(A, cd, 0), d1 = le64(A, cd, 1), d2 = le64(A, cd, 2), dn = acos_is_nan(cd);
    const bool t0 = le64(A, ct, 0), t1 = le64(A, ct, 1), t2 = le64(A, ct, 2);
    const bool g0 = !dn && !d0, g1 = !dn && !d1;          /* 30 < dihedral, 60 < dihedral */
    uint32_t g = ARP_G_NONE;                              /* :1127-1148 */
    if (d0 && t0) g = ARP_G_FF;
    else if (d0 && t1) g = ARP_G_OF;
    else if (d0 && t2) g = ARP_G_EE;
    else if (g0 && d1 && t0) g = ARP_G_FT;
    else if (g0 && d1 && t1) g = ARP_G_OT;
    else if (g0 && d1 && t2) g = ARP_G_ET;
    else if (g1 && d2 && t0) g = ARP_G_FE;
    else if (g1 && d2 && t1) g = ARP_G_OE;
    else if (g1 && d2 && t2) g = ARP_G_EF;
    if (intra && g == ARP_G_EE) return false;                                         /* :1154 */
    *geom = g; *dist = distance;
    return true;
}

template <int KIND> __device__ __forceinline__ bool row_live(const PlaneArgs& A, int row)
{
    if (KIND == KIND_RING_RING || KIND == KIND_ATOM_RING) return (A.rflags[row] & ARP_P_IN_SELECTION_PLUS) != 0;
    return (A.aflags[row] & ARP_P_IN_SELECTION_PLUS) != 0;
}

template <int KIND> __device__ bool plane_eval(const PlaneArgs& A, int row, int col, PlaneRec* out)
{
    out->pad = 0;
    if (KIND == KIND_RING_RING) {
        /* the record of an unordered pair is created by its first surviving visit; the outer loop
           runs a ascending, so visit (min, max) comes first (interactions.py:1181-1194) */
        const int a = row, b = col;
        if (a == b || !(A.rflags[b] & ARP_P_IN_SELECTION_PLUS)) return false;         /* :1081, :1085 */
        {   /* cheap reject: squared centroid distance far beyond the threshold */
            const double* ca = A.rc + 3 * (size_t)a; const double* cb = A.rc + 3 * (size_t)b;
            double x = ca[0] - cb[0], y = ca[1] - cb[1], z = ca[2] - cb[2];
            if (x * x + y * y + z * z > A.ring_centroid * A.ring_centroid * 1.000001 + 1e-9) return false;
        }
        uint32_t g1 = 0, g2 = 0; double d1 = 0, d2 = 0;
        const bool v1 = ring_visit(A, a, b, &g1, &d1);
        const bool v2 = ring_visit(A, b, a, &g2, &d2);
        if (!v1) return false;
        uint32_t second = 0xF;
        if (a < b) { if (v2 && g2 != g1) second = g2; }
        else if (v2) return false;                      /* created by row b */
        out->a = a; out->b = b; out->dist = d1;
        out->code = g1 | (second << 4) | (plane_class(A.rflags[a], A.rflags[b]) << 8) |
                    ((uint32_t)(A.rres[a] == A.rres[b]) << 11);
        return true;
    }
    if (KIND == KIND_AMIDE_AMIDE) {                    /* float32 throughout, :1217-1300 */
        const int a = row, b = col;
        if (a == b || !(A.aflags[b] & ARP_P_IN_SELECTION_PLUS)) return false;         /* :1233, :1237 */
        const float* ca = A.ac + 3 * (size_t)a; const float* cb = A.ac + 3 * (size_t)b;
        const float* na = A.an + 3 * (size_t)a; const float* nb = A.an + 3 * (size_t)b;
        const float tx = f_sub(ca[0], cb[0]), ty = f_sub(ca[1], cb[1]), tz = f_sub(ca[2], cb[2]);
        const float distance = np_norm3_f32(tx, ty, tz);                              /* :1268 */
        if (distance > A.amide_centroid_f32) return false;                            /* :1270 */
        const float nna = np_norm3_f32(na[0], na[1], na[2]);
        const float cd = f_div(np_dot3_f32(na[0], na[1], na[2], nb[0], nb[1], nb[2]),
                               f_mul(nna, np_norm3_f32(nb[0], nb[1], nb[2])));
        const float ct = f_div(np_dot3_f32(na[0], na[1], na[2], tx, ty, tz), f_mul(nna, distance));
        /* skip when dihedral > 30 or theta > 30 (:1282); a NaN angle compares false and stays */
        if (!(acos_is_nan_f(cd) || le32(A, cd, 0)) || !(acos_is_nan_f(ct) || le32(A, ct, 0))) return false;
        out->a = a; out->b = b; out->dist = (double)distance;
        out->code = 0xFFu | (plane_class(A.aflags[a], A.aflags[b]) << 8) | ((uint32_t)(A.ares[a] == A.ares[b]) << 11);
        return true;
    }
    if (KIND == KIND_AMIDE_RING) {                     /* float32 amide widened by the float64 ring, :1302-1382 */
        const int a = row, r = col;
        if (!(A.rflags[r] & ARP_P_IN_SELECTION_PLUS)) return false;                   /* :1318 */
        const float* ca = A.ac + 3 * (size_t)a; const float* na = A.an + 3 * (size_t)a;
        const double* cr = A.rc + 3 * (size_t)r; const double* nr = A.rn + 3 * (size_t)r;
        const double tx = d_sub((double)ca[0], cr[0]), ty = d_sub((double)ca[1], cr[1]), tz = d_sub((double)ca[2], cr[2]);
        const double distance = np_norm3_f64(tx, ty, tz, A.blas_fma);                 /* :1349 */
        if (distance > A.amide_centroid) return false;                                /* :1351 */
        const double nna = (double)np_norm3_f32(na[0], na[1], na[2]);                 /* float32 scalar, widened */
        const double ax = (double)na[0], ay = (double)na[1], az = (double)na[2];
        const double cd = d_div(np_dot3_f64(ax, ay, az, nr[0], nr[1], nr[2], A.blas_fma),
                                d_mul(nna, np_norm3_f64(nr[0], nr[1], nr[2], A.blas_fma)));
        const double ct = d_div(np_dot3_f64(ax, ay, az, tx, ty, tz, A.blas_fma), d_mul(nna, distance));
        if (!(acos_is_nan(cd) || le64(A, cd, 0)) || !(acos_is_nan(ct) || le64(A, ct, 0))) return false;   /* :1363 */
        out->a = a; out->b = r; out->dist = distance;
        out->code = 0xFFu | (plane_class(A.aflags[a], A.rflags[r]) << 8) | ((uint32_t)(A.ares[a] == A.rres[r]) << 11);
        return true;
    }
    if (KIND == KIND_ATOM_RING) {                      /* :947-1062; row = ring, column = atom */
        const int r = row, i = col;
        const double* c = A.rc + 3 * (size_t)r; const double* nr = A.rn + 3 * (size_t)r;
        const float* x = A.xyz + 3 * (size_t)i;
        /* NeighborSearch.search(center, met_sulphur_aromatic_distance) (:960): double, d2 <= r*r */
        const double dx = d_sub((double)x[0], c[0]), dy = d_sub((double)x[1], c[1]), dz = d_sub((double)x[2], c[2]);
        double s = d_mul(dx, dx);
        s = d_add(s, d_mul(dy, dy));
        s = d_add(s, d_mul(dz, dz));
        if (!(s <= d_mul(A.met_sulphur, A.met_sulphur))) return false;
        const uint32_t f = A.feat[i];
        if (f & ARP_F_ELEM_H) return false;                                           /* :964 */
        const double distance = np_norm3_f64(dx, dy, dz, A.blas_fma);                 /* :972 */
        if (f & ARP_F_AROMATIC) return false;                                         /* :975 */
        const bool intra = A.rres[r] == A.res_id[i];                                  /* :981 */
        const bool sr = A.rflags[r] & ARP_P_IN_SELECTION, sa = f & ARP_F_IN_SELECTION;
        uint32_t cls = ARP_CLASS_INTRA_BINDING_SITE;                                  /* :985-997 */
        if (sr && sa) cls = ARP_CLASS_INTRA_SELECTION;
        if (sr != sa) cls = ARP_CLASS_INTER;
        const double px = d_sub(c[0], (double)x[0]), py = d_sub(c[1], (double)x[1]), pz = d_sub(c[2], (double)x[2]);
        const double ct = d_div(np_dot3_f64(nr[0], nr[1], nr[2], px, py, pz, A.blas_fma),
                                d_mul(np_norm3_f64(nr[0], nr[1], nr[2], A.blas_fma), np_norm3_f64(px, py, pz, A.blas_fma)));
        uint32_t lab = 0;
        if (distance <= A.atom_ring && le64(A, ct, 0)) {                              /* :1007 */
            if ((f & ARP_F_ELEM_C) && (f & ARP_F_WEAK_HBOND_DONOR)) lab |= ARP_AP_CARBONPI;
            if (f & ARP_F_POS_IONISABLE) lab |= ARP_AP_CATIONPI;
            if (f & ARP_F_HBOND_DONOR) lab |= ARP_AP_DONORPI;
            if (f & ARP_F_XBOND_DONOR) lab |= ARP_AP_HALOGENPI;
        }
        if (distance <= A.met_sulphur && (f & ARP_F_MET_SULPHUR)) lab |= ARP_AP_METSULPHURPI;   /* :1021 */
        if (!lab) return false;                                                       /* :1026 */
        out->a = i; out->b = r; out->dist = distance;            /* arp_atom_plane: atom, ring */
        out->code = lab | (cls << 8) | ((uint32_t)intra << 11);
        return true;
    }
    return false;
}

#define PLANE_WARPS 8

template <int KIND, bool EMIT>
__global__ void __launch_bounds__(PLANE_WARPS * 32) k_plane_rows(PlaneArgs A, int n_rows, int n_cols,
                                                                 int* __restrict__ row_cnt, const int* __restrict__ row_off,
                                                                 PlaneRec* __restrict__ rec)
{
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * PLANE_WARPS + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    int cnt = 0;
    if (row_live<KIND>(A, row)) {
        const int base = EMIT ? row_off[row] : 0;
        const unsigned lt = (1u << lane) - 1u;
        for (int c0 = 0; c0 < n_cols; c0 += 32) {
            const int col = c0 + lane;
            PlaneRec r;
            bool hit = col < n_cols && plane_eval<KIND>(A, row, col, &r);
            const unsigned m = __ballot_sync(FULL, hit);
            if (EMIT && hit) rec[base + cnt + __popc(m & lt)] = r;
            cnt += __popc(m);
        }
    }
    if (!EMIT && lane == 0) row_cnt[row] = cnt;
}

template <int KIND> __device__ __forceinline__ float4 row_point(const PlaneArgs& A, int row)
{
    float x, y, z;
    if (KIND == KIND_RING_RING || KIND == KIND_ATOM_RING) {
        x = (float)A.rc[3 * (size_t)row]; y = (float)A.rc[3 * (size_t)row + 1]; z = (float)A.rc[3 * (size_t)row + 2];
    } else {
        x = A.ac[3 * (size_t)row]; y = A.ac[3 * (size_t)row + 1]; z = A.ac[3 * (size_t)row + 2];
    }
    return make_float4(x, y, z, fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))));
}

template <int KIND> __device__ __forceinline__ float screen_radius(const PlaneArgs& A)
{
    return (float)(KIND == KIND_RING_RING ? A.ring_centroid : KIND == KIND_ATOM_RING ? A.met_sulphur : A.amide_centroid);
}

/* ---- grid variant ----------------------------------------------------------------------------------
 * The double loops visit n_rows x n_cols pairs of which a few thousand lie within the 6 A of the centroid /
 * search tests (interactions.py:960 searches the KD-tree around the centroid; :1113, :1270, :1351 skip on the
 * centroid distance).  Here the column points of all terms -- atoms (atom-ring), ring centroids (ring-ring,
 * amide-ring), amide centres (amide-amide) -- are binned into three cell grids of edge >= the largest radius in
 * ONE launch sequence (bounding box, count, scan, scatter), one warp per row then visits the 27 cells around the
 * row point (9 contiguous runs of the cell-sorted float32 copies), a conservative float32 screen leaves the exact
 * predicate of the reference (plane_eval) to the few pairs that can be within the radius, and count -> scan ->
 * emit keeps the reference's creation order (row, column ascending) for all four terms at once.  O(rows) work.
 *
 * Cell coordinates are clamped to the grid, which is monotone: two points within one cell edge of each other land
 * in cells that differ by at most one per axis whatever the grid's origin and size, so one bounding box (of all
 * points) serves all three grids and a point outside it is still found.  The screen keeps NaN distances
 * (`!(d2 > T2)`); non-finite plane centres, whose pairs the reference does NOT skip, are detected at upload and
 * take the plain double loops (k_plane_rows).                                                                  */
#define PLG_SETS 3                       /* 0 atoms, 1 ring centroids, 2 amide centres */
#define PLG_TERMS 4
#define PLG_SLOTS 8                      /* records a row can leave in the counting pass; rows with more are searched again */

struct PlaneGrid {
    double ox, oy, oz, inv_w;
    int dx, dy, dz, ncell;
    int cell_base;                       /* first global cell id of the set */
    int n;                               /* points of the set */
    int pt_base;                         /* first global point id of the set */
    int pad;
};

struct PlaneGridArgs {
    PlaneArgs A;
    int term_mask;                       /* bit t: term t runs (KIND_* order) */
    double edge;                         /* cell edge before widening: >= every radius */
    unsigned* bbox;                      /* [6] ordered-int min (as max of ~ord) and max */
    PlaneGrid* grids;                    /* [PLG_SETS] */
    unsigned* n_cells;                   /* cells of the three grids together */
    int* cell_cnt;                       /* [total cells + 1] */
    int* cell_start;
    int* cell_of; int* rank;             /* per point */
    float4* spos;                        /* cell-sorted points: x, y, z (float32), index in its set */
    int* row_cnt; int* row_off;          /* rows of the four terms, concatenated */
    int row_base[PLG_TERMS + 1];
    PlaneRec* slots;                     /* counting pass: the first PLG_SLOTS records of every row, discovery order */
    int* d_tot;                          /* the record offsets at the term boundaries, for the host */
    PlaneRec* tmp;                       /* rows with more records: all of them in discovery order (positions = row_off) */
    unsigned long long tmp_cap;
    PlaneRec* rec[PLG_TERMS];            /* per term: records in (row, column) order */
    unsigned long long cap[PLG_TERMS];
};

__device__ __forceinline__ unsigned pl_f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pl_ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

/* point p of the concatenated sets -> (set, index, float32 coordinates, binned?) */
__device__ __forceinline__ bool pl_point(const PlaneArgs& A, int n_atoms, int p, int* set, int* idx, float* x, float* y, float* z)
{
    if (p < n_atoms) {
        *set = 0; *idx = p;
        *x = A.xyz[3 * (size_t)p]; *y = A.xyz[3 * (size_t)p + 1]; *z = A.xyz[3 * (size_t)p + 2];
        return !(A.feat[p] & ARP_F_ELEM_H);                                           /* interactions.py:964 */
    }
    p -= n_atoms;
    if (p < A.nr) {
        *set = 1; *idx = p;
        *x = (float)A.rc[3 * (size_t)p]; *y = (float)A.rc[3 * (size_t)p + 1]; *z = (float)A.rc[3 * (size_t)p + 2];
        return (A.rflags[p] & ARP_P_IN_SELECTION_PLUS) != 0;                          /* :1085, :1318 */
    }
    p -= A.nr;
    *set = 2; *idx = p;
    *x = A.ac[3 * (size_t)p]; *y = A.ac[3 * (size_t)p + 1]; *z = A.ac[3 * (size_t)p + 2];
    return (A.aflags[p] & ARP_P_IN_SELECTION_PLUS) != 0;                              /* :1237 */
}

__global__ void __launch_bounds__(256) k_pl_bbox(PlaneGridArgs G, int n_atoms, int n_points)
{
    __shared__ unsigned s_v[8][6];
    unsigned v[6] = {0, 0, 0, 0, 0, 0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_points; p += gridDim.x * blockDim.x) {
        int set, idx; float x, y, z;
        pl_point(G.A, n_atoms, p, &set, &idx, &x, &y, &z);
        if (x == x && y == y && z == z && fabsf(x) < 3e38f && fabsf(y) < 3e38f && fabsf(z) < 3e38f) {   /* finite points only */
            const unsigned ox = pl_f2ord(x), oy = pl_f2ord(y), oz = pl_f2ord(z);
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
    if (threadIdx.x < 6) {                      /* six atomics per block */
        unsigned m = 0;
        for (int w = 0; w < 8; ++w) m = max(m, s_v[w][threadIdx.x]);
        if (m) atomicMax(&G.bbox[threadIdx.x], m);
    }
}

__device__ __forceinline__ int pl_cell_coord(double v, double o, double inv_w, int dim)
{
    double t = floor((v - o) * inv_w);
    t = fmin(fmax(t, 0.0), (double)(dim - 1));         /* clamped (monotone); NaN -> 0 */
    return (int)t;
}
__device__ __forceinline__ int pl_cell(const PlaneGrid& g, float x, float y, float z, int* cx, int* cy, int* cz)
{
    *cx = pl_cell_coord((double)x, g.ox, g.inv_w, g.dx);
    *cy = pl_cell_coord((double)y, g.oy, g.inv_w, g.dy);
    *cz = pl_cell_coord((double)z, g.oz, g.inv_w, g.dz);
    return (*cz * g.dy + *cy) * g.dx + *cx;
}

/* every block derives the three grids from the bounding box itself; block 0 publishes them */
__device__ __forceinline__ void pl_make_grids(const PlaneGridArgs& G, int n_atoms, PlaneGrid* out)
{
    double mn[3], mx[3];
    bool any = false;
    for (int k = 0; k < 3; ++k) {
        const unsigned lo = G.bbox[k], hi = G.bbox[3 + k];
        any = any || hi != 0u;
        mn[k] = hi ? (double)pl_ord2f(~lo) : 0.0;
        mx[k] = hi ? (double)pl_ord2f(hi) : 0.0;
    }
    (void)any;
    double amax = 0.0;
    for (int k = 0; k < 3; ++k) amax = fmax(amax, fmax(fabs(mn[k]), fabs(mx[k])));
    const double edge = G.edge + 5e-7 * amax;          /* the points are binned by their float32 images */
    const int n_set[PLG_SETS] = { n_atoms, G.A.nr, G.A.na };
    int cell_base = 0, pt_base = 0;
    for (int s = 0; s < PLG_SETS; ++s) {
        PlaneGrid g;
        double w = edge;
        long long d[3];
        for (;;) {
            for (int k = 0; k < 3; ++k) d[k] = (long long)floor((mx[k] - mn[k]) / w) + 1;
            if ((double)d[0] * (double)d[1] * (double)d[2] <= 4.0 * (double)n_set[s] + 64.0) break;   /* the host sizes the tables on this bound */
            w *= 1.5;
        }
        g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2]; g.inv_w = 1.0 / w;
        g.dx = (int)d[0]; g.dy = (int)d[1]; g.dz = (int)d[2]; g.ncell = g.dx * g.dy * g.dz;
        g.cell_base = cell_base; g.n = n_set[s]; g.pt_base = pt_base; g.pad = 0;
        cell_base += g.ncell; pt_base += n_set[s];
        out[s] = g;
    }
}
/* total number of cells of the three grids (block 0 of k_pl_count leaves it for the scan) */

__global__ void __launch_bounds__(256) k_pl_count(PlaneGridArgs G, int n_atoms, int n_points)
{
    __shared__ PlaneGrid s_g[PLG_SETS];
    if (threadIdx.x == 0) {
        pl_make_grids(G, n_atoms, s_g);
        if (blockIdx.x == 0) {
            for (int s = 0; s < PLG_SETS; ++s) G.grids[s] = s_g[s];
            *G.n_cells = (unsigned)(s_g[PLG_SETS - 1].cell_base + s_g[PLG_SETS - 1].ncell);
        }
    }
    __syncthreads();
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_points; p += gridDim.x * blockDim.x) {
        int set, idx, cx, cy, cz; float x, y, z;
        int c = -1, r = 0;
        if (pl_point(G.A, n_atoms, p, &set, &idx, &x, &y, &z)) {
            c = s_g[set].cell_base + pl_cell(s_g[set], x, y, z, &cx, &cy, &cz);
            r = atomicAdd(&G.cell_cnt[c], 1);
        }
        G.cell_of[p] = c; G.rank[p] = r;
    }
}

__global__ void __launch_bounds__(256) k_pl_scatter(PlaneGridArgs G, int n_atoms, int n_points)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n_points; p += gridDim.x * blockDim.x) {
        const int c = G.cell_of[p];
        if (c < 0) continue;
        int set, idx; float x, y, z;
        pl_point(G.A, n_atoms, p, &set, &idx, &x, &y, &z);
        G.spos[G.cell_start[c] + G.rank[p]] = make_float4(x, y, z, __int_as_float(idx));
    }
}

template <int KIND> __device__ __forceinline__ int pl_col_set() { return KIND == KIND_ATOM_RING ? 0 : KIND == KIND_AMIDE_AMIDE ? 2 : 1; }

/* one warp, one row of term KIND: the 27 cells around the row point are 9 contiguous runs of the cell-sorted points;
   lanes 0..8 fetch the runs' bounds together (one round trip), then the lanes stride over the concatenated
   candidates.  EMIT: records in discovery order to tmp, then by ascending column into rec */
template <int KIND> __device__ __forceinline__ int pl_key(const PlaneRec& r) { return KIND == KIND_ATOM_RING ? r.a : r.b; }

/* EMIT == false (counting pass): the row's records are counted and the first PLG_SLOTS of them kept in the row's slots;
   EMIT == true (emitting pass): a row with at most PLG_SLOTS records only moves them from its slots to their places,
   by ascending column; a row with more is searched again, its records go through tmp */
template <int KIND, bool EMIT>
__device__ __forceinline__ void pl_row(const PlaneGridArgs& G, const PlaneGrid& g, int row, int grow, int lane)
{
    const PlaneArgs& A = G.A;
    int cnt = 0;
    int base = 0;
    if (EMIT) {
        base = G.row_off[grow];
        const int n = G.row_off[grow + 1] - base;
        if (n == 0 || (unsigned long long)base + n > G.cap[KIND]) return;
        if (n <= PLG_SLOTS) {
            if (lane < n) {
                const PlaneRec mine = G.slots[(size_t)grow * PLG_SLOTS + lane];
                const int key = pl_key<KIND>(mine);
                int rk = 0;
                for (int k = 0; k < n; ++k) rk += pl_key<KIND>(G.slots[(size_t)grow * PLG_SLOTS + k]) < key ? 1 : 0;
                G.rec[KIND][base + rk] = mine;                  /* one buffer for all terms: the row offsets are global */
            }
            return;
        }
        if ((unsigned long long)base + n > G.tmp_cap) return;
    }
    if (row_live<KIND>(A, row)) {
        const float4 rp = row_point<KIND>(A, row);
        const float r = screen_radius<KIND>(A);
        const float t0 = r * 1.000004f + 1e-6f;
        int cx, cy, cz;
        pl_cell(g, rp.x, rp.y, rp.z, &cx, &cy, &cz);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
        const unsigned lt = (1u << lane) - 1u;
        int rb = 0, rl = 0;                                        /* lane k < 9: run k = (first position, length) */
        if (lane < 9) {
            const int y = cy + lane % 3 - 1, z = cz + lane / 3 - 1;
            if (y >= 0 && y < g.dy && z >= 0 && z < g.dz) {
                const int rowc = g.cell_base + (z * g.dy + y) * g.dx;
                rb = G.cell_start[rowc + x0];
                rl = G.cell_start[rowc + x1 + 1] - rb;
            }
        }
        int incl = rl;
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 8);
        const int cofs = rb - (incl - rl);                         /* position = cofs + k for the k of this run */
        int co[9], pe[9];                                          /* per run: offset, end of its k range */
#pragma unroll
        for (int k = 0; k < 9; ++k) { co[k] = __shfl_sync(FULL, cofs, k); pe[k] = __shfl_sync(FULL, incl, k); }
        for (int k0 = 0; k0 < total; k0 += 32) {
            const int k = k0 + lane;
            bool hit = false;
            PlaneRec rec;
            if (k < total) {
                int q = k + co[8];
#pragma unroll
                for (int t = 7; t >= 0; --t) if (k < pe[t]) q = k + co[t];
                const float4 cp = G.spos[q];
                const float dx = rp.x - cp.x, dy = rp.y - cp.y, dz = rp.z - cp.z;
                const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                const float T = t0 + 4e-7f * (rp.w + fmaxf(fabsf(cp.x), fmaxf(fabsf(cp.y), fabsf(cp.z))));
                if (!(d2 > T * T)) hit = plane_eval<KIND>(A, row, __float_as_int(cp.w), &rec);    /* rare: the exact predicate decides */
            }
            const unsigned m = __ballot_sync(FULL, hit);
            if (hit) {
                const int pos = cnt + __popc(m & lt);
                if (EMIT) G.tmp[base + pos] = rec;
                else if (pos < PLG_SLOTS) G.slots[(size_t)grow * PLG_SLOTS + pos] = rec;
            }
            cnt += __popc(m);
        }
    }
    if (!EMIT) {
        if (lane == 0) G.row_cnt[grow] = cnt;
        return;
    }
    __syncwarp();
    /* column order inside the row: the rank of a record is the number of the row's records with a smaller column.
       The column is field b, except for atom-ring records (atom, ring), where it is field a. */
    for (int h = lane; h < cnt; h += 32) {
        const PlaneRec mine = G.tmp[base + h];
        const int key = pl_key<KIND>(mine);
        int rk = 0;
        for (int k = 0; k < cnt; ++k) rk += pl_key<KIND>(G.tmp[base + k]) < key ? 1 : 0;
        G.rec[KIND][base + rk] = mine;
    }
}

#define PLG_WARPS 8
template <bool EMIT>
__global__ void __launch_bounds__(PLG_WARPS * 32, 4) k_pl_pass(PlaneGridArgs G)
{
    const int lane = threadIdx.x & 31;
    const int grow = blockIdx.x * PLG_WARPS + (threadIdx.x >> 5);
    if (EMIT && blockIdx.x == 0 && threadIdx.x <= PLG_TERMS)      /* the record offsets at the term boundaries, for the host */
        G.d_tot[threadIdx.x] = G.row_off[G.row_base[threadIdx.x]];
    if (grow >= G.row_base[PLG_TERMS]) return;
    if (grow < G.row_base[1])      pl_row<KIND_RING_RING, EMIT>(G, G.grids[pl_col_set<KIND_RING_RING>()], grow - G.row_base[0], grow, lane);
    else if (grow < G.row_base[2]) pl_row<KIND_ATOM_RING, EMIT>(G, G.grids[pl_col_set<KIND_ATOM_RING>()], grow - G.row_base[1], grow, lane);
    else if (grow < G.row_base[3]) pl_row<KIND_AMIDE_AMIDE, EMIT>(G, G.grids[pl_col_set<KIND_AMIDE_AMIDE>()], grow - G.row_base[2], grow, lane);
    else                           pl_row<KIND_AMIDE_RING, EMIT>(G, G.grids[pl_col_set<KIND_AMIDE_RING>()], grow - G.row_base[3], grow, lane);
}

/* ---- host side ------------------------------------------------------------------------------- */

static int plane_upload(arp_ctx* c, PlaneSet& ps, const arp_planes* p)
{
    ps.n = 0; ps.is_f32 = 0;
    if (!p || p->n == 0) return ARP_OK;
    ARP_REQUIRE(c, p->n > 0, ARP_E_INVALID_ARG, "negative plane count");
    ARP_REQUIRE(c, p->center && p->normal && p->res_id && p->flags, ARP_E_INVALID_ARG, "a plane array is NULL");
    size_t el = p->is_f32 ? 4 : 8, n = (size_t)p->n;
    ARP_TRY(dbuf_reserve(c, ps.center, n * 3 * el));
    ARP_TRY(dbuf_reserve(c, ps.normal, n * 3 * el));
    ARP_TRY(dbuf_reserve(c, ps.res_id, n * 4));
    ARP_TRY(dbuf_reserve(c, ps.flags, n * 4));
    ARP_CUDA(c, cudaMemcpyAsync(ps.center.p, p->center, n * 3 * el, cudaMemcpyHostToDevice, c->stream));
    ARP_CUDA(c, cudaMemcpyAsync(ps.normal.p, p->normal, n * 3 * el, cudaMemcpyHostToDevice, c->stream));
    ARP_CUDA(c, cudaMemcpyAsync(ps.res_id.p, p->res_id, n * 4, cudaMemcpyHostToDevice, c->stream));
    ARP_CUDA(c, cudaMemcpyAsync(ps.flags.p, p->flags, n * 4, cudaMemcpyHostToDevice, c->stream));
    ps.n = p->n; ps.is_f32 = p->is_f32;
    return ARP_OK;
}

static void plane_args(arp_ctx* c, PlaneArgs* A)
{
    memset(A, 0, sizeof *A);
    A->nr = c->rings.n; A->rc = c->rings.center.as<double>(); A->rn = c->rings.normal.as<double>();
    A->rres = c->rings.res_id.as<int32_t>(); A->rflags = c->rings.flags.as<uint32_t>();
    A->na = c->amides.n; A->ac = c->amides.center.as<float>(); A->an = c->amides.normal.as<float>();
    A->ares = c->amides.res_id.as<int32_t>(); A->aflags = c->amides.flags.as<uint32_t>();
    A->n_atoms = c->have_atoms ? c->N : 0; A->xyz = c->xyz.as<float>(); A->feat = c->feat.as<uint32_t>();
    A->res_id = c->res_id.as<int32_t>();
    const arp_params& p = c->params;
    A->blas_fma = p.blas_fma;
    A->ring_centroid = p.ring_centroid_dist; A->amide_centroid = p.amide_centroid_dist;
    A->atom_ring = p.atom_ring_dist; A->met_sulphur = p.met_sulphur_dist;
    A->amide_centroid_f32 = (float)p.amide_centroid_dist;
    A->split64 = p.cos_split_f64; A->split32 = p.cos_split_f32;
    for (int k = 0; k < 3; ++k) {
        A->pos64[k] = p.cos_pos_f64[k]; A->neg64[k] = p.cos_neg_f64[k];
        A->pos32[k] = p.cos_pos_f32[k]; A->neg32[k] = p.cos_neg_f32[k];
    }
}

/* the plain double loop of one term: fallback for non-finite plane centres, and the A/B knob ARPEGGIO_NO_PLANE_GRID */
template <int KIND> static int plane_run_rows(arp_ctx* c, PlaneResult& R, int n_rows, int n_cols)
{
    R.valid = 0; R.n = 0; R.ptr = nullptr; R.host = nullptr;
    if (n_rows > 0 && n_cols > 0) {
        PlaneArgs A; plane_args(c, &A);
        /* zero region: row_cnt[n_rows + 1] | ticket | scan state ; then row_off */
        size_t tiles = ((size_t)n_rows + 1 + ARP_SCAN_TILE - 1) / ARP_SCAN_TILE + 1;
        size_t o_tick = ((size_t)(n_rows + 2) * 4 + 255) / 256 * 256;
        size_t o_state = o_tick + 256;
        size_t o_off = o_state + tiles * 8;
        size_t bytes = o_off + (size_t)(n_rows + 2) * 4;
        ARP_TRY(dbuf_reserve(c, R.cnt, bytes));
        char* z = R.cnt.as<char>();
        int* row_cnt = (int*)z; unsigned* ticket = (unsigned*)(z + o_tick);
        unsigned long long* state = (unsigned long long*)(z + o_state);
        int* row_off = (int*)(z + o_off);
        ARP_CUDA(c, cudaMemsetAsync(z, 0, o_off, c->stream));
        unsigned blocks = (unsigned)((n_rows + PLANE_WARPS - 1) / PLANE_WARPS);
        k_plane_rows<KIND, false><<<blocks, PLANE_WARPS * 32, 0, c->stream>>>(A, n_rows, n_cols, row_cnt, nullptr, nullptr);
        ARP_LAUNCHED(c);
        ARP_TRY(arp_scan_exclusive(c, row_cnt, row_off, state, ticket, nullptr, n_rows + 1, (size_t)n_rows + 1));
        int total = 0;
        ARP_CUDA(c, cudaMemcpyAsync(&total, row_off + n_rows, 4, cudaMemcpyDeviceToHost, c->stream));
        ARP_CUDA(c, cudaStreamSynchronize(c->stream));
        if (total > 0) {
            ARP_TRY(dbuf_reserve(c, R.rec, (size_t)total * sizeof(PlaneRec)));
            k_plane_rows<KIND, true><<<blocks, PLANE_WARPS * 32, 0, c->stream>>>(A, n_rows, n_cols, nullptr, row_off, R.rec.as<PlaneRec>());
            ARP_LAUNCHED(c);
        }
        R.n = (uint64_t)total;
        R.ptr = R.rec.p;
    }
    R.valid = 1;
    return ARP_OK;
}

static PlaneResult& plane_result(arp_ctx* c, int term)
{
    return term == KIND_RING_RING ? c->ring_ring : term == KIND_ATOM_RING ? c->atom_ring : term == KIND_AMIDE_AMIDE ? c->amide_amide : c->amide_ring;
}

/* The terms of `mask` through the cell grids: ONE launch sequence for all of them (bounding box, count, scan,
   scatter, counting pass, scan, emitting pass) and one wait, for the record totals. */
static int planes_run_grid(arp_ctx* c, int mask)
{
    PlaneGridArgs G;
    memset(&G, 0, sizeof G);
    plane_args(c, &G.A);
    const int nr = c->rings.n, na = c->amides.n;
    const int n_atoms = (mask & (1 << KIND_ATOM_RING)) ? c->N : 0;      /* the atoms are binned only for atom-ring */
    const int rows_t[PLG_TERMS] = { (mask & 1) ? nr : 0, (mask & 2) ? nr : 0, (mask & 4) ? na : 0, (mask & 8) ? na : 0 };
    G.row_base[0] = 0;
    for (int t = 0; t < PLG_TERMS; ++t) G.row_base[t + 1] = G.row_base[t] + rows_t[t];
    const int rows = G.row_base[PLG_TERMS];
    for (int t = 0; t < PLG_TERMS; ++t) if (mask & (1 << t)) { PlaneResult& R = plane_result(c, t); R.valid = 0; R.n = 0; R.ptr = nullptr; R.host = nullptr; }
    if (rows == 0) {
        for (int t = 0; t < PLG_TERMS; ++t) if (mask & (1 << t)) plane_result(c, t).valid = 1;
        return ARP_OK;
    }
    G.term_mask = mask;
    const arp_params& p = c->params;
    double r = p.ring_centroid_dist > p.amide_centroid_dist ? p.ring_centroid_dist : p.amide_centroid_dist;
    r = p.met_sulphur_dist > r ? p.met_sulphur_dist : r;
    G.edge = r > 1e-3 ? r * (1.0 + 1e-5) + 1e-5 : 1e-3;
    const size_t P = (size_t)n_atoms + nr + na;
    const size_t cells = 4 * P + 64 * PLG_SETS + 2;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    /* zero region: bbox | n_cells | cell_cnt | scan tickets and states | row_cnt */
    const size_t tiles_c = (cells + ARP_SCAN_TILE - 1) / ARP_SCAN_TILE + 1, tiles_r = ((size_t)rows + 1 + ARP_SCAN_TILE - 1) / ARP_SCAN_TILE + 1;
    const size_t o_cnt = 256, o_tick = up(o_cnt + cells * 4), o_state_c = o_tick + 256, o_state_r = up(o_state_c + tiles_c * 8);
    const size_t o_rcnt = up(o_state_r + tiles_r * 8), zero_bytes = up(o_rcnt + ((size_t)rows + 2) * 4);
    /* scratch: grids | cell_start | cell_of | rank | spos | row_off */
    const size_t o_start = up(zero_bytes + sizeof(PlaneGrid) * PLG_SETS), o_cellof = up(o_start + cells * 4), o_rank = up(o_cellof + P * 4);
    const size_t o_spos = up(o_rank + P * 4), o_roff = up(o_spos + P * 16), o_slots = up(o_roff + ((size_t)rows + 2) * 4);
    const size_t bytes = up(o_slots + (size_t)rows * PLG_SLOTS * sizeof(PlaneRec));
    ARP_TRY(dbuf_reserve(c, c->plane_scratch, bytes));
    char* z = c->plane_scratch.as<char>();
    G.bbox = (unsigned*)z; G.n_cells = (unsigned*)(z + 64); G.cell_cnt = (int*)(z + o_cnt);
    unsigned* tick_c = (unsigned*)(z + o_tick); unsigned* tick_r = (unsigned*)(z + o_tick + 128);
    unsigned long long* state_c = (unsigned long long*)(z + o_state_c); unsigned long long* state_r = (unsigned long long*)(z + o_state_r);
    G.row_cnt = (int*)(z + o_rcnt);
    G.grids = (PlaneGrid*)(z + zero_bytes); G.cell_start = (int*)(z + o_start); G.cell_of = (int*)(z + o_cellof);
    G.rank = (int*)(z + o_rank); G.spos = (float4*)(z + o_spos); G.row_off = (int*)(z + o_roff);
    G.slots = (PlaneRec*)(z + o_slots); G.d_tot = (int*)(z + 128);       /* d_tot: inside the zero region, behind the bounding box and the cell total */
    cudaStream_t st = c->stream;
    ARP_CUDA(c, cudaMemsetAsync(z, 0, zero_bytes, st));
    const unsigned pb = (unsigned)((P + 255) / 256 < (size_t)c->sm_count * 8 ? (P + 255) / 256 : (size_t)c->sm_count * 8);
    k_pl_bbox<<<pb ? pb : 1, 256, 0, st>>>(G, n_atoms, (int)P);
    ARP_LAUNCHED(c);
    k_pl_count<<<pb ? pb : 1, 256, 0, st>>>(G, n_atoms, (int)P);
    ARP_LAUNCHED(c);
    ARP_TRY(arp_scan_exclusive(c, G.cell_cnt, G.cell_start, state_c, tick_c, G.n_cells, 1, cells));
    k_pl_scatter<<<pb ? pb : 1, 256, 0, st>>>(G, n_atoms, (int)P);
    ARP_LAUNCHED(c);
    const unsigned rb = (unsigned)((rows + PLG_WARPS - 1) / PLG_WARPS);
    k_pl_pass<false><<<rb, PLG_WARPS * 32, 0, st>>>(G);
    ARP_LAUNCHED(c);
    ARP_TRY(arp_scan_exclusive(c, G.row_cnt, G.row_off, state_r, tick_r, nullptr, rows + 1, (size_t)rows + 1));
    int* h_tot = c->h_plane_tot;                    /* pinned: the record offsets at the term boundaries */
    /* every run of this path leaves its records in ONE buffer (term t at offset h_tot[t]) and, with one copy, in a pinned
       host mirror: the fetch calls then cost no CUDA call.  Results of earlier plane runs are no longer valid.
       The emitting pass is enqueued BLIND behind the counting pass, with the capacity the last runs needed, so that
       the whole sequence costs one wait; a run that needs more is emitted again with larger buffers. */
    for (PlaneResult* R : { &c->ring_ring, &c->atom_ring, &c->amide_amide, &c->amide_ring }) { R->valid = 0; R->host = nullptr; }
    for (int attempt = 0; attempt < 2; ++attempt) {
        size_t cap = c->plane_cap_guess > 1024 ? c->plane_cap_guess : 1024;
        if (attempt == 1) cap = (size_t)h_tot[PLG_TERMS] + (size_t)h_tot[PLG_TERMS] / 4 + 1024;
        ARP_TRY(dbuf_reserve(c, c->plane_tmp, cap * sizeof(PlaneRec)));
        ARP_TRY(dbuf_reserve(c, c->plane_rec, cap * sizeof(PlaneRec)));
        if (c->h_plane_cap < cap) {
            if (c->h_plane_rec) cudaFreeHost(c->h_plane_rec);
            c->h_plane_rec = nullptr; c->h_plane_cap = 0;
            ARP_CUDA(c, cudaMallocHost(&c->h_plane_rec, cap * sizeof(PlaneRec)));
            c->h_plane_cap = cap;
        }
        G.tmp = c->plane_tmp.as<PlaneRec>(); G.tmp_cap = (unsigned long long)cap;
        /* the term's first record sits at row_off[row_base[t]], which only the device knows yet: one buffer, term offsets
           applied in the kernel (rec[t] = buffer start + that offset) */
        for (int t = 0; t < PLG_TERMS; ++t) { G.rec[t] = c->plane_rec.as<PlaneRec>(); G.cap[t] = (unsigned long long)cap; }
        k_pl_pass<true><<<rb, PLG_WARPS * 32, 0, st>>>(G);
        ARP_LAUNCHED(c);
        ARP_CUDA(c, cudaMemcpyAsync(h_tot, G.d_tot, (PLG_TERMS + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
        ARP_CUDA(c, cudaMemcpyAsync(c->h_plane_rec, c->plane_rec.p, cap * sizeof(PlaneRec), cudaMemcpyDeviceToHost, st));
        ARP_CUDA(c, cudaStreamSynchronize(st));
        if ((size_t)h_tot[PLG_TERMS] <= cap) break;
        ARP_REQUIRE(c, attempt == 0, ARP_E_CAPACITY, "plane records overflowed repeatedly");
    }
    c->plane_cap_guess = (size_t)h_tot[PLG_TERMS] + (size_t)h_tot[PLG_TERMS] / 2 + 1024;
    for (int t = 0; t < PLG_TERMS; ++t) if (mask & (1 << t)) {
        PlaneResult& R = plane_result(c, t);
        R.n = (uint64_t)(h_tot[t + 1] - h_tot[t]);
        R.ptr = c->plane_rec.as<PlaneRec>() + h_tot[t];
        R.host = R.n ? (const char*)c->h_plane_rec + (size_t)h_tot[t] * sizeof(PlaneRec) : nullptr;
        R.valid = 1;
    }
    return ARP_OK;
}

/* runs the terms of `mask`; n_out[t] (may be null) receives the record counts */
static int planes_run(arp_ctx* c, int mask, uint64_t* n_out)
{
    const bool need_atoms = (mask & (1 << KIND_ATOM_RING)) != 0;
    if (need_atoms) {
        ARP_REQUIRE(c, c->have_atoms, ARP_E_NOT_READY, "arp_atom_ring_run before arp_upload_atoms");
        ARP_REQUIRE(c, c->S == 1, ARP_E_INVALID_ARG, "plane terms need a single structure");
    }
    if (c->use_plane_grid && c->planes_finite) {
        ARP_TRY(planes_run_grid(c, mask));
    } else {
        if (mask & 1) ARP_TRY(plane_run_rows<KIND_RING_RING>(c, c->ring_ring, c->rings.n, c->rings.n));
        if (mask & 2) ARP_TRY(plane_run_rows<KIND_ATOM_RING>(c, c->atom_ring, c->rings.n, c->N));
        if (mask & 4) ARP_TRY(plane_run_rows<KIND_AMIDE_AMIDE>(c, c->amide_amide, c->amides.n, c->amides.n));
        if (mask & 8) ARP_TRY(plane_run_rows<KIND_AMIDE_RING>(c, c->amide_ring, c->amides.n, c->rings.n));
    }
    if (n_out) for (int t = 0; t < PLG_TERMS; ++t) if (mask & (1 << t)) n_out[t] = plane_result(c, t).n;
    return ARP_OK;
}

static int plane_fetch(arp_ctx* c, PlaneResult& R, void* dst, uint64_t cap)
{
    ARP_REQUIRE(c, R.valid, ARP_E_NOT_READY, "fetch before run");
    ARP_REQUIRE(c, cap >= R.n, ARP_E_CAPACITY, "destination holds fewer records than the run produced");
    if (R.n == 0) return ARP_OK;
    ARP_REQUIRE(c, dst != nullptr, ARP_E_INVALID_ARG, "dst is NULL");
    if (R.host) {                       /* the run has already brought the records to the host */
        memcpy(dst, R.host, (size_t)R.n * sizeof(PlaneRec));
        return ARP_OK;
    }
    ARP_TRY(arp_bind(c));
    ARP_CUDA(c, cudaMemcpyAsync(dst, R.ptr, (size_t)R.n * sizeof(PlaneRec), cudaMemcpyDeviceToHost, c->stream));
    ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    return ARP_OK;
}

void arp_planes_release(arp_ctx* c)
{
    for (PlaneSet* ps : { &c->rings, &c->amides }) {
        dbuf_free(ps->center); dbuf_free(ps->normal); dbuf_free(ps->res_id); dbuf_free(ps->flags);
    }
    for (PlaneResult* r : { &c->ring_ring, &c->atom_ring, &c->amide_amide, &c->amide_ring }) {
        dbuf_free(r->rec); dbuf_free(r->tmp); dbuf_free(r->cnt);
    }
    dbuf_free(c->plane_scratch); dbuf_free(c->plane_tmp); dbuf_free(c->plane_rec);
    if (c->h_plane_tot) { cudaFreeHost(c->h_plane_tot); c->h_plane_tot = nullptr; }
    if (c->h_plane_rec) { cudaFreeHost(c->h_plane_rec); c->h_plane_rec = nullptr; c->h_plane_cap = 0; }
}

extern "C" {

int arp_upload_planes(arp_ctx* c, const arp_planes* rings, const arp_planes* amides)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, !rings || rings->n == 0 || !rings->is_f32, ARP_E_INVALID_ARG, "rings must be float64");
    ARP_REQUIRE(c, !amides || amides->n == 0 || amides->is_f32, ARP_E_INVALID_ARG, "amides must be float32");
    ARP_TRY(arp_bind(c));
    c->have_planes = 0;
    c->ring_ring.valid = c->atom_ring.valid = c->amide_amide.valid = c->amide_ring.valid = 0;
    ARP_TRY(plane_upload(c, c->rings, rings));
    ARP_TRY(plane_upload(c, c->amides, amides));
    /* a non-finite centre makes the reference's `distance > threshold: continue` tests false, i.e. it pairs that plane
       with every other one: such inputs take the plain double loops */
    c->planes_finite = 1;
    if (rings && rings->n > 0) {
        const double* p = (const double*)rings->center;
        for (size_t k = 0; k < (size_t)rings->n * 3; ++k) if (!(fabs(p[k]) <= 1.7e308)) { c->planes_finite = 0; break; }
    }
    if (amides && amides->n > 0) {
        const float* p = (const float*)amides->center;
        for (size_t k = 0; k < (size_t)amides->n * 3; ++k) if (!(fabsf(p[k]) <= 3.4e38f)) { c->planes_finite = 0; break; }
    }
    c->have_planes = 1;
    return ARP_OK;
}

static int planes_entry(arp_ctx* c, int mask, uint64_t* n_out, const char* what)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->have_planes, ARP_E_NOT_READY, what);
    ARP_TRY(arp_bind(c));
    if (!c->h_plane_tot) ARP_CUDA(c, cudaMallocHost((void**)&c->h_plane_tot, 64));
    return planes_run(c, mask, n_out);
}

int arp_ring_ring_run(arp_ctx* c, uint64_t* n)
{
    uint64_t v[PLG_TERMS] = {0, 0, 0, 0};
    int rc = planes_entry(c, 1 << KIND_RING_RING, v, "arp_ring_ring_run before arp_upload_planes");
    if (rc == ARP_OK && n) *n = v[KIND_RING_RING];
    return rc;
}
int arp_ring_ring_fetch(arp_ctx* c, arp_plane_pair* dst, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    return plane_fetch(c, c->ring_ring, dst, cap);
}

int arp_atom_ring_run(arp_ctx* c, uint64_t* n)
{
    uint64_t v[PLG_TERMS] = {0, 0, 0, 0};
    int rc = planes_entry(c, 1 << KIND_ATOM_RING, v, "arp_atom_ring_run before arp_upload_planes");
    if (rc == ARP_OK && n) *n = v[KIND_ATOM_RING];
    return rc;
}
int arp_atom_ring_fetch(arp_ctx* c, arp_atom_plane* dst, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    return plane_fetch(c, c->atom_ring, dst, cap);
}

int arp_amide_amide_run(arp_ctx* c, uint64_t* n)
{
    uint64_t v[PLG_TERMS] = {0, 0, 0, 0};
    int rc = planes_entry(c, 1 << KIND_AMIDE_AMIDE, v, "arp_amide_amide_run before arp_upload_planes");
    if (rc == ARP_OK && n) *n = v[KIND_AMIDE_AMIDE];
    return rc;
}
int arp_amide_amide_fetch(arp_ctx* c, arp_plane_pair* dst, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    return plane_fetch(c, c->amide_amide, dst, cap);
}

int arp_amide_ring_run(arp_ctx* c, uint64_t* n)
{
    uint64_t v[PLG_TERMS] = {0, 0, 0, 0};
    int rc = planes_entry(c, 1 << KIND_AMIDE_RING, v, "arp_amide_ring_run before arp_upload_planes");
    if (rc == ARP_OK && n) *n = v[KIND_AMIDE_RING];
    return rc;
}
int arp_amide_ring_fetch(arp_ctx* c, arp_plane_pair* dst, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    return plane_fetch(c, c->amide_ring, dst, cap);
}

/* all four terms in one launch sequence: n[0..3] = records of ring-ring, atom-ring, amide-amide, amide-ring
   (atom-ring is left out -- n[1] = 0, its fetch fails with ARP_E_NOT_READY -- when no single structure is uploaded) */
int arp_planes_run_all(arp_ctx* c, uint64_t* n)
{
    if (!c) return ARP_E_INVALID_ARG;
    int mask = 0xF;
    if (!c->have_atoms || c->S != 1) { mask &= ~(1 << KIND_ATOM_RING); c->atom_ring.valid = 0; }
    uint64_t v[PLG_TERMS] = {0, 0, 0, 0};
    int rc = planes_entry(c, mask, v, "arp_planes_run_all before arp_upload_planes");
    if (rc == ARP_OK && n) for (int t = 0; t < PLG_TERMS; ++t) n[t] = v[t];
    return rc;
}

}  /* extern "C" */
