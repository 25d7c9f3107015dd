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

/* ---- screened variant ------------------------------------------------------------------------------
 * The double loops visit n_rows x n_cols pairs of which a few thousand lie within the 6 A of the centroid /
 * search tests.  k_plane_scan tiles the column points through shared memory as float32 (one tile serves the
 * eight rows of a block) and runs the exact predicate only where a float32 distance screen cannot exclude the
 * pair; the exact hits of a row are kept as a bit per column, so the emitting pass (k_plane_emit) re-evaluates
 * those pairs only.  The screen is conservative: threshold widened by the float32 rounding of the largest
 * coordinate of the row point and of the tile, `!(d2 > T2)` keeps NaN for the exact code to judge.       */
#define PL_TILE 1024

template <int KIND> __device__ __forceinline__ float4 col_point(const PlaneArgs& A, int col)
{
    float x, y, z;
    bool live = true;
    if (KIND == KIND_ATOM_RING) {
        x = A.xyz[3 * (size_t)col]; y = A.xyz[3 * (size_t)col + 1]; z = A.xyz[3 * (size_t)col + 2];
    } else if (KIND == KIND_AMIDE_AMIDE) {
        x = A.ac[3 * (size_t)col]; y = A.ac[3 * (size_t)col + 1]; z = A.ac[3 * (size_t)col + 2];
        live = (A.aflags[col] & ARP_P_IN_SELECTION_PLUS) != 0;
    } else {
        x = (float)A.rc[3 * (size_t)col]; y = (float)A.rc[3 * (size_t)col + 1]; z = (float)A.rc[3 * (size_t)col + 2];
        live = (A.rflags[col] & ARP_P_IN_SELECTION_PLUS) != 0;
    }
    const float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    if (!live) return make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);      /* +inf: beyond every threshold */
    return make_float4(x, y, z, mag);
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

#define PL_ROWS (PLANE_WARPS * 32)          /* rows per block: one per thread */

template <int KIND>
__global__ void __launch_bounds__(PL_ROWS) k_plane_scan(PlaneArgs A, int n_rows, int n_cols, int words,
                                                       int* __restrict__ row_cnt, uint32_t* __restrict__ mask)
{
    /* thread = row (its point in registers), the column points of a tile are broadcast from shared memory:
       one conflict-free LDS serves 32 rows, and the 32 screen bits a thread collects for 32 consecutive columns
       are exactly its row's word of the hit bitmask -- no ballots.  blockIdx.y splits the columns so that every
       SM has work when the rows are few (atom-ring). */
    __shared__ float4 s_col[PL_TILE];
    __shared__ float s_mag[PLANE_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * PL_ROWS + threadIdx.x;
    const bool live = row < n_rows && row_live<KIND>(A, row);
    float4 rp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) rp = row_point<KIND>(A, row);
    const float r = screen_radius<KIND>(A);
    int cnt = 0;
    const int tiles = (n_cols + PL_TILE - 1) / PL_TILE;
    const int t_lo = (int)((long long)tiles * blockIdx.y / gridDim.y), t_hi = (int)((long long)tiles * (blockIdx.y + 1) / gridDim.y);
    for (int c0 = t_lo * PL_TILE; c0 < t_hi * PL_TILE && c0 < n_cols; c0 += PL_TILE) {
        const int m = min(PL_TILE, n_cols - c0);
        __syncthreads();                                   /* the previous tile has been consumed */
        float mag = 0.f;
        for (int k = threadIdx.x; k < PL_TILE; k += PL_ROWS) {
            float4 q = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);
            if (k < m) q = col_point<KIND>(A, c0 + k);
            s_col[k] = q;
            mag = q.w == q.w ? fmaxf(mag, q.w) : q.w;      /* a NaN coordinate poisons the tile: nothing is screened out */
        }
        {   /* warp maximum; non-negative floats order like their bit patterns */
            const bool nan = __any_sync(FULL, mag != mag);
            mag = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(mag == mag ? mag : 0.f)));
            if (nan) mag = __int_as_float(0x7fc00000);
        }
        if (lane == 0) s_mag[warp] = mag;
        __syncthreads();
        float tile_mag = 0.f;
#pragma unroll
        for (int w = 0; w < PLANE_WARPS; ++w) tile_mag = s_mag[w] == s_mag[w] ? fmaxf(tile_mag, s_mag[w]) : s_mag[w];
        const float T = r * 1.000004f + 1e-6f + 4e-7f * (rp.w + tile_mag);
        const float T2 = T * T;
#pragma unroll 1
        for (int g = 0; g < PL_TILE / 32 && c0 + 32 * g < n_cols; ++g) {
            uint32_t w = 0;                                /* bit b: column c0 + 32 g + b survives the screen */
#pragma unroll
            for (int b = 0; b < 32; ++b) {
                const float4 q = s_col[g * 32 + b];
                const float dx = rp.x - q.x, dy = rp.y - q.y, dz = rp.z - q.z;
                const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                if (!(d2 > T2)) w |= 1u << b;
            }
            if (live && w) {                               /* rare: the exact predicate decides */
                uint32_t hits = 0;
                while (w) {
                    const int b = __ffs(w) - 1;
                    w &= w - 1;
                    const int col = c0 + g * 32 + b;
                    PlaneRec rec;
                    if (col < n_cols && plane_eval<KIND>(A, row, col, &rec)) hits |= 1u << b;
                }
                if (hits) {
                    mask[(size_t)row * words + (c0 >> 5) + g] = hits;
                    cnt += __popc(hits);
                }
            }
        }
    }
    if (cnt) atomicAdd(&row_cnt[row], cnt);                /* row_cnt is zeroed by the caller */
}

template <int KIND>
__global__ void __launch_bounds__(PLANE_WARPS * 32) k_plane_emit(PlaneArgs A, int n_rows, int words,
                                                                 const int* __restrict__ row_off, const uint32_t* __restrict__ mask,
                                                                 PlaneRec* __restrict__ rec)
{
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * PLANE_WARPS + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int base = row_off[row], n = row_off[row + 1] - base;
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    for (int w0 = 0; w0 < words && cnt < n; w0 += 32) {
        const uint32_t wv = w0 + lane < words ? mask[(size_t)row * words + w0 + lane] : 0u;
        unsigned nz = __ballot_sync(FULL, wv != 0u);
        while (nz) {
            const int j = __ffs(nz) - 1;
            nz &= nz - 1;
            const uint32_t word = __shfl_sync(FULL, wv, j);
            if (word >> lane & 1u) {
                PlaneRec r;
                plane_eval<KIND>(A, row, (w0 + j) * 32 + lane, &r);
                rec[base + cnt + __popc(word & lt)] = r;
            }
            cnt += __popc(word);
        }
    }
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

template <int KIND> static int plane_run(arp_ctx* c, PlaneResult& R, int n_rows, int n_cols, uint64_t* n_out)
{
    R.valid = 0; R.n = 0;
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
        /* one bit per (row, column) for the exact hits, if that fits: the emitting pass then touches only those */
        const int words = (n_cols + 31) / 32;
        const size_t mask_bytes = (size_t)n_rows * (size_t)words * 4;
        const bool screened = c->use_plane_screen && mask_bytes <= ((size_t)1 << 30);
        if (screened) {
            ARP_TRY(dbuf_reserve(c, R.tmp, mask_bytes));
            ARP_CUDA(c, cudaMemsetAsync(R.tmp.p, 0, mask_bytes, c->stream));
            const unsigned tiles = (unsigned)((n_cols + PL_TILE - 1) / PL_TILE);
            const unsigned row_blocks = (unsigned)((n_rows + PL_ROWS - 1) / PL_ROWS);
            unsigned splits = ((unsigned)c->sm_count * 6 + row_blocks - 1) / row_blocks;
            splits = splits < 1 ? 1 : splits > tiles ? tiles : splits;
            k_plane_scan<KIND><<<dim3(row_blocks, splits), PL_ROWS, 0, c->stream>>>(A, n_rows, n_cols, words, row_cnt,
                                                                                   R.tmp.as<uint32_t>());
        } else {
            k_plane_rows<KIND, false><<<blocks, PLANE_WARPS * 32, 0, c->stream>>>(A, n_rows, n_cols, row_cnt, nullptr, nullptr);
        }
        ARP_LAUNCHED(c);
        ARP_TRY(arp_scan_exclusive(c, row_cnt, row_off, state, ticket, nullptr, n_rows + 1, (size_t)n_rows + 1));
        int total = 0;
        ARP_CUDA(c, cudaMemcpyAsync(&total, row_off + n_rows, 4, cudaMemcpyDeviceToHost, c->stream));
        ARP_CUDA(c, cudaStreamSynchronize(c->stream));
        if (total > 0) {
            ARP_TRY(dbuf_reserve(c, R.rec, (size_t)total * sizeof(PlaneRec)));
            if (screened)
                k_plane_emit<KIND><<<blocks, PLANE_WARPS * 32, 0, c->stream>>>(A, n_rows, words, row_off, R.tmp.as<uint32_t>(),
                                                                               R.rec.as<PlaneRec>());
            else
                k_plane_rows<KIND, true><<<blocks, PLANE_WARPS * 32, 0, c->stream>>>(A, n_rows, n_cols, nullptr, row_off,
                                                                                     R.rec.as<PlaneRec>());
            ARP_LAUNCHED(c);
            ARP_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        R.n = (uint64_t)total;
    }
    R.valid = 1;
    if (n_out) *n_out = R.n;
    return ARP_OK;
}

static int plane_fetch(arp_ctx* c, PlaneResult& R, void* dst, uint64_t cap)
{
    ARP_REQUIRE(c, R.valid, ARP_E_NOT_READY, "fetch before run");
    ARP_REQUIRE(c, cap >= R.n, ARP_E_CAPACITY, "destination holds fewer records than the run produced");
    if (R.n == 0) return ARP_OK;
    ARP_REQUIRE(c, dst != nullptr, ARP_E_INVALID_ARG, "dst is NULL");
    ARP_TRY(arp_bind(c));
    ARP_CUDA(c, cudaMemcpyAsync(dst, R.rec.p, (size_t)R.n * sizeof(PlaneRec), cudaMemcpyDeviceToHost, c->stream));
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
    c->have_planes = 1;
    return ARP_OK;
}

int arp_ring_ring_run(arp_ctx* c, uint64_t* n)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->have_planes, ARP_E_NOT_READY, "arp_ring_ring_run before arp_upload_planes");
    ARP_TRY(arp_bind(c));
    return plane_run<KIND_RING_RING>(c, c->ring_ring, c->rings.n, c->rings.n, n);
}
int arp_ring_ring_fetch(arp_ctx* c, arp_plane_pair* dst, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    return plane_fetch(c, c->ring_ring, dst, cap);
}

int arp_atom_ring_run(arp_ctx* c, uint64_t* n)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->have_planes, ARP_E_NOT_READY, "arp_atom_ring_run before arp_upload_planes");
    ARP_REQUIRE(c, c->have_atoms, ARP_E_NOT_READY, "arp_atom_ring_run before arp_upload_atoms");
    ARP_REQUIRE(c, c->S == 1, ARP_E_INVALID_ARG, "plane terms need a single structure");
    ARP_TRY(arp_bind(c));
    return plane_run<KIND_ATOM_RING>(c, c->atom_ring, c->rings.n, c->N, n);
}
int arp_atom_ring_fetch(arp_ctx* c, arp_atom_plane* dst, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    return plane_fetch(c, c->atom_ring, dst, cap);
}

int arp_amide_amide_run(arp_ctx* c, uint64_t* n)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->have_planes, ARP_E_NOT_READY, "arp_amide_amide_run before arp_upload_planes");
    ARP_TRY(arp_bind(c));
    return plane_run<KIND_AMIDE_AMIDE>(c, c->amide_amide, c->amides.n, c->amides.n, n);
}
int arp_amide_amide_fetch(arp_ctx* c, arp_plane_pair* dst, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    return plane_fetch(c, c->amide_amide, dst, cap);
}

int arp_amide_ring_run(arp_ctx* c, uint64_t* n)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->have_planes, ARP_E_NOT_READY, "arp_amide_ring_run before arp_upload_planes");
    ARP_TRY(arp_bind(c));
    return plane_run<KIND_AMIDE_RING>(c, c->amide_ring, c->amides.n, c->rings.n, n);
}
int arp_amide_ring_fetch(arp_ctx* c, arp_plane_pair* dst, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    return plane_fetch(c, c->amide_ring, dst, cap);
}

}  /* extern "C" */
