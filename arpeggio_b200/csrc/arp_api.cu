/*
 * arp_api.cu -- the C ABI of libarpeggio_cuda.so (include/arpeggio_cuda.h): life cycle,
 * parameters, uploads, the atom-atom run/fetch calls and the benchmark hooks.
 * The plane entry points live in arp_planes.cu.
 */
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <math.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <sched.h>
#include <stdlib.h>
#include <new>
#include <thread>
#include <vector>

#include "arp_ctx.cuh"

static char g_create_err[ARP_ERRLEN] = {0};

/* ---- cosine images of the angle thresholds, from the C library's acos (see arp_params) ---- */
namespace {

inline long long ord64(double x) { long long b; memcpy(&b, &x, 8); return b < 0 ? -(b & 0x7fffffffffffffffLL) : b; }
inline double unord64(long long k) { long long b = k >= 0 ? k : (long long)((unsigned long long)(-k) | 0x8000000000000000ULL); double x; memcpy(&x, &b, 8); return x; }
inline long long ord32(float x) { int b; memcpy(&b, &x, 4); return b < 0 ? -(long long)(b & 0x7fffffff) : (long long)b; }
inline float unord32(long long k) { unsigned b = k >= 0 ? (unsigned)k : ((unsigned)(-k) | 0x80000000u); float x; memcpy(&x, &b, 4); return x; }

/* last argument on the True side of a monotone predicate over [lo, hi]; *constant set if it never switches */
template <class Pred> double edge64(Pred pred, double lo, double hi, bool true_below, bool* constant)
{
    long long klo = ord64(lo), khi = ord64(hi);
    bool plo = pred(unord64(klo)), phi = pred(unord64(khi));
    *constant = plo == phi;
    if (*constant) return 0.0;
    while (khi - klo > 1) {
        long long mid = klo + (khi - klo) / 2;
        if (pred(unord64(mid)) == plo) klo = mid; else khi = mid;
    }
    return unord64(true_below ? klo : khi);
}
template <class Pred> float edge32(Pred pred, float lo, float hi, bool true_below, bool* constant)
{
    long long klo = ord32(lo), khi = ord32(hi);
    bool plo = pred(unord32(klo)), phi = pred(unord32(khi));
    *constant = plo == phi;
    if (*constant) return 0.f;
    while (khi - klo > 1) {
        long long mid = klo + (khi - klo) / 2;
        if (pred(unord32(mid)) == plo) klo = mid; else khi = mid;
    }
    return unord32(true_below ? klo : khi);
}

inline double fold_deg64(double c)
{
    double rad = acos(c);
    if (rad > M_PI / 2) rad = rad - M_PI;
    return fabs(rad * 180 / M_PI);
}
inline float fold_deg32(float c)
{
    float rad = acosf(c);
    if (rad > (float)(M_PI / 2)) { volatile float t = rad - (float)M_PI; rad = t; }
    volatile float deg = rad * 180.0f;
    deg = deg / (float)M_PI;
    return fabsf(deg);
}

void cosine_images(arp_params* p)
{
    bool k;
    auto ge = [&](double thr) {
        double e = edge64([&](double c) { return acos(c) >= thr; }, -1.0, 1.0, true, &k);
        return k ? (acos(1.0) >= thr ? 2.0 : -2.0) : e;
    };
    p->cos_hbond = ge(p->hbond_angle);
    p->cos_weak_hbond = ge(p->weak_hbond_angle);
    p->cos_cx_min = ge(p->cx_angle_min);
    {
        double mx = p->cx_angle_max;
        double e = edge64([&](double c) { return acos(c) <= mx; }, -1.0, 1.0, false, &k);
        p->cos_cx_max = k ? (acos(-1.0) <= mx ? -2.0 : 2.0) : e;
    }
    {
        float thr = (float)p->xbond_angle;
        float e = edge32([&](float c) { return acosf(c) >= thr; }, -1.f, 1.f, true, &k);
        p->cos_xbond_f32 = k ? (acosf(1.f) >= thr ? 2.f : -2.f) : e;
    }
    {
        double split = edge64([&](double c) { return acos(c) > M_PI / 2; }, -1.0, 1.0, true, &k);
        double nxt = unord64(ord64(split) + 1);
        p->cos_split_f64 = split;
        for (int i = 0; i < 3; ++i) {
            double b = p->plane_bins_deg[i];
            double e = edge64([&](double c) { return fold_deg64(c) <= b; }, nxt, 1.0, false, &k);
            p->cos_pos_f64[i] = k ? (fold_deg64(1.0) <= b ? -2.0 : 2.0) : e;
            e = edge64([&](double c) { return fold_deg64(c) <= b; }, -1.0, split, true, &k);
            p->cos_neg_f64[i] = k ? (fold_deg64(-1.0) <= b ? 2.0 : -2.0) : e;
        }
    }
    {
        float split = edge32([&](float c) { return acosf(c) > (float)(M_PI / 2); }, -1.f, 1.f, true, &k);
        float nxt = unord32(ord32(split) + 1);
        p->cos_split_f32 = split;
        for (int i = 0; i < 3; ++i) {
            float b = (float)p->plane_bins_deg[i];
            float e = edge32([&](float c) { return fold_deg32(c) <= b; }, nxt, 1.f, false, &k);
            p->cos_pos_f32[i] = k ? (fold_deg32(1.f) <= b ? -2.f : 2.f) : e;
            e = edge32([&](float c) { return fold_deg32(c) <= b; }, -1.f, split, true, &k);
            p->cos_neg_f32[i] = k ? (fold_deg32(-1.f) <= b ? 2.f : -2.f) : e;
        }
    }
}

void derive_rule_params(const arp_params& p, ArpRuleParams* r) { arp_derive_rule_params(&p, r); }

int upload(arp_ctx* c, DBuf& b, const void* src, size_t bytes)
{
    ARP_TRY(dbuf_reserve(c, b, bytes));
    if (bytes) ARP_CUDA(c, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    c->input_bytes += bytes;
    return ARP_OK;
}

/* the record stream of the last run is no longer current; a run that is still in flight is waited for first */
void pairs_invalidate(arp_ctx* c)
{
    if (c->run_pending) { cudaStreamSynchronize(c->stream); (void)cudaGetLastError(); c->run_pending = 0; }
    c->pk.pending = 0;
    c->pairs_valid = 0; c->sorted_valid = 0; c->compact_valid = 0; c->packed_valid = 0; c->sort_tmp_valid = 0; c->sifts_valid = 0;
}

}  // namespace

/* ---- a batch of structures in ONE upload: packed on the device ---------------------------------------------
 * arp_upload_atoms takes a batch as ONE arp_atoms whose arrays the caller has already concatenated (struct_off,
 * residue / bond / hydrogen indices rebased).  arp_upload_atoms_batch takes the structures as they are -- one
 * arp_atoms each, indices local to the structure, radius tables of their own -- moves every structure with its own
 * DMA(s) into a staging arena and lets the device do the concatenation: indices rebased, radius classes mapped
 * onto one merged table.  The batch then runs as one launch sequence (struct_off inside): 16 structures of 20k atoms
 * cost the kernels of one 320k-atom structure instead of 16 separate launch sequences. */
namespace {

struct HostTimer {                     /* accumulates the host time of a scope into ctx->ht_ns[slot] (ARPEGGIO_HOST_TIMING prints them) */
    arp_ctx* c; int slot; timespec t0; bool live;
    HostTimer(arp_ctx* c_, int s) : c(c_), slot(s), live(true) { clock_gettime(CLOCK_MONOTONIC, &t0); }
    void stop() {
        if (!live) return;
        timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
        c->ht_ns[slot] += (unsigned long long)((t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec));
        ++c->ht_calls[slot];
        live = false;
    }
    ~HostTimer() { stop(); }
};

/* ---- wire forms of arp_atoms, decoded on the device after the copy ------------------------------------------
 * per-atom counts (1 byte) in place of CSR offsets (4 bytes), the halogen neighbours as (atom index, coordinate)
 * rows in place of a dense [N][3] array, hydrogen coordinates as int32 fixed point in place of float64.  The
 * kernels below turn them into the arrays every other kernel reads, so nothing downstream knows the difference. */
constexpr int CNT_THREADS = 256, CNT_ITEMS = 16, CNT_TILE = CNT_THREADS * CNT_ITEMS;

template <class T> __device__ __forceinline__ int cnt_load(const T* __restrict__ cnt, int n, int base, int (&v)[CNT_ITEMS])
{
    int s = 0;
#pragma unroll
    for (int k = 0; k < CNT_ITEMS; ++k) { const int i = base + k; v[k] = i < n ? (int)cnt[i] : 0; s += v[k]; }
    return s;
}

/* exclusive prefix of `s` over the block's threads, and the block total */
__device__ __forceinline__ int cnt_block_scan(int s, int* total)
{
    __shared__ int w_sum[CNT_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) w_sum[w] = inc;
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < CNT_THREADS / 32; ++k) { const int t = w_sum[k]; if (k < w) before += t; tot += t; }
    *total = tot;
    return before + inc - s;
}

/* counts of up to two arrays over the same n atoms (blockIdx.y): tile sums, their prefix, the offsets */
template <class T> __global__ void __launch_bounds__(CNT_THREADS) k_cnt_sums(const T* __restrict__ cnt_a, const T* __restrict__ cnt_b, int n,
                                                                             int* __restrict__ sums, int tiles)
{
    int v[CNT_ITEMS], tot;
    const int s = cnt_load(blockIdx.y ? cnt_b : cnt_a, n, blockIdx.x * CNT_TILE + threadIdx.x * CNT_ITEMS, v);
    cnt_block_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.y * tiles + blockIdx.x] = tot;
}

__global__ void __launch_bounds__(64) k_cnt_top(int* __restrict__ sums, int tiles, int arrays)
{
    const int lane = threadIdx.x & 31, y = threadIdx.x >> 5;
    if (y >= arrays) return;
    int carry = 0;
    for (int b = 0; b < tiles; b += 32) {
        const int v = b + lane < tiles ? sums[y * tiles + b + lane] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
        if (b + lane < tiles) sums[y * tiles + b + lane] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

/* sum of `v` over the block's threads, in every thread */
__device__ __forceinline__ int cnt_block_total(int v)
{
    __shared__ int w_tot[CNT_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) w_tot[w] = v;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int k = 0; k < CNT_THREADS / 32; ++k) tot += w_tot[k];
    __syncthreads();
    return tot;
}

template <class T, bool SUM_TILES> __global__ void __launch_bounds__(CNT_THREADS) k_cnt_fill(const T* __restrict__ cnt_a, const T* __restrict__ cnt_b, int n,
                                                                             const int* __restrict__ sums, int tiles,
                                                                             int* __restrict__ off_a, int* __restrict__ off_b)
{
    int v[CNT_ITEMS], tot;
    const int base = blockIdx.x * CNT_TILE + threadIdx.x * CNT_ITEMS;
    const int s = cnt_load(blockIdx.y ? cnt_b : cnt_a, n, base, v);
    int before = 0;
    if (SUM_TILES) {                    /* few tiles: every block adds up the sums of the tiles before it (no k_cnt_top pass) */
        for (int t = threadIdx.x; t < (int)blockIdx.x; t += CNT_THREADS) before += sums[blockIdx.y * tiles + t];
        before = cnt_block_total(before);
    } else {
        before = sums[blockIdx.y * tiles + blockIdx.x];
    }
    int run = before + cnt_block_scan(s, &tot);
    int* __restrict__ off = blockIdx.y ? off_b : off_a;
#pragma unroll
    for (int k = 0; k < CNT_ITEMS; ++k) {
        const int i = base + k;
        if (i < n) { off[i] = run; run += v[k]; if (i == n - 1) off[n] = run; }
    }
}

/* offsets [n + 1] of one or two count arrays on c->stream; scratch: 2 * tiles ints */
template <class T> int counts_to_offsets(arp_ctx* c, const T* cnt_a, int* off_a, const T* cnt_b, int* off_b, int n)
{
    if (n <= 0) return ARP_OK;
    if (!cnt_a) { cnt_a = cnt_b; off_a = off_b; cnt_b = nullptr; off_b = nullptr; }
    if (!cnt_a) return ARP_OK;
    const int tiles = (n + CNT_TILE - 1) / CNT_TILE, arrays = cnt_b ? 2 : 1;
    ARP_TRY(dbuf_reserve(c, c->wire_sums, (size_t)tiles * 2 * sizeof(int)));
    int* sums = c->wire_sums.as<int>();
    k_cnt_sums<T><<<dim3((unsigned)tiles, (unsigned)arrays), CNT_THREADS, 0, c->stream>>>(cnt_a, cnt_b, n, sums, tiles);
    ARP_LAUNCHED(c);
    if (tiles <= 1024) {
        k_cnt_fill<T, true><<<dim3((unsigned)tiles, (unsigned)arrays), CNT_THREADS, 0, c->stream>>>(cnt_a, cnt_b, n, sums, tiles, off_a, off_b);
        ARP_LAUNCHED(c);
    } else {
        k_cnt_top<<<1, 64, 0, c->stream>>>(sums, tiles, arrays);
        ARP_LAUNCHED(c);
        k_cnt_fill<T, false><<<dim3((unsigned)tiles, (unsigned)arrays), CNT_THREADS, 0, c->stream>>>(cnt_a, cnt_b, n, sums, tiles, off_a, off_b);
        ARP_LAUNCHED(c);
    }
    return ARP_OK;
}

/* fixed-point hydrogens -> float64 (IEEE division, the caller verified the round trip); sparse halogen neighbours -> dense
   rows: every atom that carries ARP_F_HAS_XNBR looks its row up in the ascending index list (zeros when it is missing);
   the rows of the other atoms are never read (arp_rules.cuh tests the flag first) and stay as they are */
__global__ void __launch_bounds__(256) k_wire_rest(const int32_t* __restrict__ h_fix, double scale, double* __restrict__ h_xyz, long long n_h3,
                                                   const uint32_t* __restrict__ feat, const int32_t* __restrict__ x_idx,
                                                   const float* __restrict__ x_src, float* __restrict__ xnbr, int n_x, int n_atoms)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_h3) h_xyz[t] = __ddiv_rn((double)h_fix[t], scale);
    if (xnbr && t < n_atoms && (feat[t] & ARP_F_HAS_XNBR)) {
        int lo = 0, hi = n_x;                   /* first entry >= t */
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (x_idx[mid] < (int)t) lo = mid + 1; else hi = mid; }
        float a = 0.f, b = 0.f, c = 0.f;
        if (lo < n_x && x_idx[lo] == (int)t) { a = x_src[3 * (size_t)lo]; b = x_src[3 * (size_t)lo + 1]; c = x_src[3 * (size_t)lo + 2]; }
        xnbr[3 * (size_t)t] = a; xnbr[3 * (size_t)t + 1] = b; xnbr[3 * (size_t)t + 2] = c;
    }
}

struct PartDesc {                      /* device: where the arrays of one structure sit in the staging arena (byte offsets, -1: absent) */
    int atom_base, n_atoms, res_base, n_res, bond_base, h_base, class_base, x_base;
    long long o_xyz, o_feat, o_res_id, o_rad, o_prev, o_next, o_flags, o_boff, o_bnbr, o_hoff, o_hxyz, o_xnbr;
    long long o_bcnt, o_hcnt, o_hfix, o_xidx;      /* wire forms (arp_atoms.bond_cnt, h_cnt, h_fix, xnbr_idx) */
    double h_scale;
};

struct MergeArgs {
    const PartDesc* parts; int n_parts;
    const char* stage; const unsigned short* class_map;
    int N, Rs, E, H, X;                 /* X: rows of all sparse neighbour lists */
    float* xyz; uint32_t* feat; int32_t* res_id; uint16_t* rad_class; int32_t* res_prev; int32_t* res_next; uint8_t* res_flags;
    int32_t* bond_off; int32_t* bond_nbr; int32_t* h_off; double* h_xyz; float* xnbr;
    int32_t* bond_cnt; int32_t* h_cnt;  /* merged per-atom counts (scanned into bond_off / h_off afterwards) when any structure came with counts */
};

/* part of global index i for bases ascending with duplicates (empty structures): last part whose base is <= i */
template <class F> __device__ __forceinline__ int part_of(const PartDesc* p, int n, int i, F base)
{
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (base(p[mid]) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_merge_atoms(MergeArgs M)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        if (M.bond_off) M.bond_off[M.N] = M.E;
        if (M.h_off) M.h_off[M.N] = M.H;
    }
    if (i >= M.N) return;
    const PartDesc& d = M.parts[part_of(M.parts, M.n_parts, i, [](const PartDesc& q) { return q.atom_base; })];
    const int li = i - d.atom_base;
    const float* x = (const float*)(M.stage + d.o_xyz) + 3 * (size_t)li;
    M.xyz[3 * (size_t)i] = x[0]; M.xyz[3 * (size_t)i + 1] = x[1]; M.xyz[3 * (size_t)i + 2] = x[2];
    M.feat[i] = ((const uint32_t*)(M.stage + d.o_feat))[li];
    M.res_id[i] = ((const int32_t*)(M.stage + d.o_res_id))[li] + d.res_base;
    M.rad_class[i] = M.class_map[d.class_base + ((const uint16_t*)(M.stage + d.o_rad))[li]];
    if (M.bond_cnt) {                   /* counts now, offsets by the scan that follows */
        int n = 0;
        if (d.o_bcnt >= 0) n = ((const uint8_t*)(M.stage + d.o_bcnt))[li];
        else if (d.o_boff >= 0) { const int32_t* o = (const int32_t*)(M.stage + d.o_boff); n = o[li + 1] - o[li]; }
        M.bond_cnt[i] = n;
    } else if (M.bond_off) M.bond_off[i] = d.bond_base + (d.o_boff >= 0 ? ((const int32_t*)(M.stage + d.o_boff))[li] : 0);
    if (M.h_cnt) {
        int n = 0;
        if (d.o_hcnt >= 0) n = ((const uint8_t*)(M.stage + d.o_hcnt))[li];
        else if (d.o_hoff >= 0) { const int32_t* o = (const int32_t*)(M.stage + d.o_hoff); n = o[li + 1] - o[li]; }
        M.h_cnt[i] = n;
    } else if (M.h_off) M.h_off[i] = d.h_base + (d.o_hoff >= 0 ? ((const int32_t*)(M.stage + d.o_hoff))[li] : 0);
    if (M.xnbr) {                       /* dense rows here; sparse rows are scattered over the zeros by k_merge_rest */
        float a = 0.f, b = 0.f, c = 0.f;
        if (d.o_xnbr >= 0 && d.o_xidx < 0) { const float* q = (const float*)(M.stage + d.o_xnbr) + 3 * (size_t)li; a = q[0]; b = q[1]; c = q[2]; }
        M.xnbr[3 * (size_t)i] = a; M.xnbr[3 * (size_t)i + 1] = b; M.xnbr[3 * (size_t)i + 2] = c;
    }
}

__global__ void __launch_bounds__(256) k_merge_rest(MergeArgs M)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < M.Rs) {
        const PartDesc& d = M.parts[part_of(M.parts, M.n_parts, t, [](const PartDesc& q) { return q.res_base; })];
        const int lr = t - d.res_base;
        const int pv = ((const int32_t*)(M.stage + d.o_prev))[lr], nx = ((const int32_t*)(M.stage + d.o_next))[lr];
        M.res_prev[t] = pv >= 0 ? pv + d.res_base : pv;
        M.res_next[t] = nx >= 0 ? nx + d.res_base : nx;
        M.res_flags[t] = ((const uint8_t*)(M.stage + d.o_flags))[lr];
    }
    if (t < M.E) {
        const PartDesc& d = M.parts[part_of(M.parts, M.n_parts, t, [](const PartDesc& q) { return q.bond_base; })];
        M.bond_nbr[t] = ((const int32_t*)(M.stage + d.o_bnbr))[t - d.bond_base] + d.atom_base;
    }
    if (t < M.H) {
        const PartDesc& d = M.parts[part_of(M.parts, M.n_parts, t, [](const PartDesc& q) { return q.h_base; })];
        double x, y, z;
        if (d.o_hfix >= 0) {
            const int32_t* h = (const int32_t*)(M.stage + d.o_hfix) + 3 * (size_t)(t - d.h_base);
            x = __ddiv_rn((double)h[0], d.h_scale); y = __ddiv_rn((double)h[1], d.h_scale); z = __ddiv_rn((double)h[2], d.h_scale);
        } else {
            const double* h = (const double*)(M.stage + d.o_hxyz) + 3 * (size_t)(t - d.h_base);
            x = h[0]; y = h[1]; z = h[2];
        }
        M.h_xyz[3 * (size_t)t] = x; M.h_xyz[3 * (size_t)t + 1] = y; M.h_xyz[3 * (size_t)t + 2] = z;
    }
}

/* sparse halogen-neighbour rows of all structures -> the dense merged array (after k_merge_atoms zeroed / filled it) */
__global__ void __launch_bounds__(256) k_merge_xnbr(MergeArgs M)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= M.X) return;
    const PartDesc& d = M.parts[part_of(M.parts, M.n_parts, t, [](const PartDesc& q) { return q.x_base; })];
    const int r = t - d.x_base;
    const int i = d.atom_base + ((const int32_t*)(M.stage + d.o_xidx))[r];
    const float* q = (const float*)(M.stage + d.o_xnbr) + 3 * (size_t)r;
    M.xnbr[3 * (size_t)i] = q[0]; M.xnbr[3 * (size_t)i + 1] = q[1]; M.xnbr[3 * (size_t)i + 2] = q[2];
}

/* totals and consistency of the wire forms of one arp_atoms (host): E, H, rows of the neighbour list */
struct WireSizes { long long E, H, X; bool cnt_b, cnt_h, fix_h, sparse_x; };

int wire_sizes(arp_ctx* c, const arp_atoms* a, WireSizes* w)
{
    const int N = a->n_atoms;
    memset(w, 0, sizeof *w);
    w->cnt_b = a->bond_cnt != nullptr; w->cnt_h = a->h_cnt != nullptr; w->fix_h = a->h_fix != nullptr; w->sparse_x = a->xnbr_idx != nullptr;
    if (N <= 0) return ARP_OK;
    ARP_REQUIRE(c, !(a->bond_cnt && a->bond_off), ARP_E_INVALID_ARG, "bond_cnt and bond_off are alternatives");
    ARP_REQUIRE(c, !(a->h_cnt && a->h_off), ARP_E_INVALID_ARG, "h_cnt and h_off are alternatives");
    ARP_REQUIRE(c, !(a->h_fix && a->h_xyz), ARP_E_INVALID_ARG, "h_fix and h_xyz are alternatives");
    ARP_REQUIRE(c, !a->h_fix || (a->h_fix_scale > 0.0 && a->h_fix_scale < 1e12), ARP_E_INVALID_ARG, "h_fix_scale out of range");
    auto total = [&](const uint8_t* cnt) {         /* byte sum: 16 at a time (psadbw) where SSE2 is there, else 8 in four 16-bit lanes */
        long long s = 0;
        int i = 0;
#if defined(__SSE2__)
        __m128i acc = _mm_setzero_si128();
        const __m128i zero = _mm_setzero_si128();
        for (; i + 16 <= N; i += 16) acc = _mm_add_epi64(acc, _mm_sad_epu8(_mm_loadu_si128((const __m128i*)(cnt + i)), zero));
        s = (long long)_mm_cvtsi128_si64(acc) + (long long)_mm_cvtsi128_si64(_mm_unpackhi_epi64(acc, acc));
#else
        const uint64_t M = 0x00ff00ff00ff00ffull;
        while (i + 8 <= N) {
            uint64_t acc = 0;
            for (int k = 0; k < 128 && i + 8 <= N; ++k, i += 8) {
                uint64_t x;
                memcpy(&x, cnt + i, 8);
                acc += (x & M) + ((x >> 8) & M);
            }
            s += (long long)((acc & 0xffff) + ((acc >> 16) & 0xffff) + ((acc >> 32) & 0xffff) + (acc >> 48));
        }
#endif
        for (; i < N; ++i) s += cnt[i];
        return s;
    };
    if (a->bond_cnt) {
        ARP_REQUIRE(c, a->n_bond_nbr >= 0 && total(a->bond_cnt) == a->n_bond_nbr, ARP_E_INVALID_ARG, "bond_cnt does not sum to n_bond_nbr");
        w->E = a->n_bond_nbr;
    } else if (a->bond_off) w->E = a->bond_off[N];
    if (a->h_cnt) {
        ARP_REQUIRE(c, a->n_h >= 0 && total(a->h_cnt) == a->n_h, ARP_E_INVALID_ARG, "h_cnt does not sum to n_h");
        w->H = a->n_h;
    } else if (a->h_off) w->H = a->h_off[N];
    ARP_REQUIRE(c, w->E >= 0 && w->H >= 0, ARP_E_INVALID_ARG, "negative CSR size");
    ARP_REQUIRE(c, w->E == 0 || a->bond_nbr, ARP_E_INVALID_ARG, "bond_nbr is NULL");
    ARP_REQUIRE(c, w->H == 0 || a->h_xyz || a->h_fix, ARP_E_INVALID_ARG, "h_xyz is NULL");
    if (a->xnbr_idx) {
        ARP_REQUIRE(c, a->n_xnbr >= 0 && a->n_xnbr <= N && (a->n_xnbr == 0 || a->xnbr_xyz), ARP_E_INVALID_ARG, "xnbr_idx without rows");
        for (int k = 0; k < a->n_xnbr; ++k)
            ARP_REQUIRE(c, a->xnbr_idx[k] >= 0 && a->xnbr_idx[k] < N && (k == 0 || a->xnbr_idx[k] > a->xnbr_idx[k - 1]), ARP_E_INVALID_ARG,
                        "xnbr_idx must ascend strictly inside the atoms");
        w->X = a->n_xnbr;
    }
    return ARP_OK;
}

int check_part(arp_ctx* c, const arp_atoms* a)
{
    ARP_REQUIRE(c, a != nullptr, ARP_E_INVALID_ARG, "a structure of the batch is NULL");
    ARP_REQUIRE(c, a->n_atoms >= 0 && a->n_residues >= 0 && a->n_rad_classes >= 0, ARP_E_INVALID_ARG, "negative size");
    ARP_REQUIRE(c, a->n_structures <= 1, ARP_E_INVALID_ARG, "arp_upload_atoms_batch takes single structures");
    if (a->n_atoms > 0) {
        ARP_REQUIRE(c, a->xyz && a->feat && a->res_id && a->rad_class && a->vdw && a->cov && a->res_prev &&
                       a->res_next && a->res_flags, ARP_E_INVALID_ARG, "a required atom array is NULL");
        ARP_REQUIRE(c, a->n_residues > 0 && a->n_rad_classes > 0, ARP_E_INVALID_ARG, "atoms without residues or radius classes");
        ARP_REQUIRE(c, !a->bond_off || a->bond_off[0] == 0, ARP_E_INVALID_ARG, "bond_off[0] != 0");
        ARP_REQUIRE(c, !a->h_off || a->h_off[0] == 0, ARP_E_INVALID_ARG, "h_off[0] != 0");
    }
    return ARP_OK;
}

}  // namespace

extern "C" {

int arp_abi_version(void) { return ARP_ABI_VERSION; }

int arp_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) { (void)cudaGetLastError(); return 0; }
    if (e != cudaSuccess) { (void)cudaGetLastError(); return ARP_E_CUDA; }
    return n;
}

int arp_params_default(arp_params* p)
{
    if (!p) return ARP_E_INVALID_ARG;
    memset(p, 0, sizeof *p);
    p->interacting_cutoff = 5.0;         /* process_protein_cli.py -i default */
    p->vdw_comp = 0.1;                   /* -co default */
    p->include_sequence_adjacent = 0;
    p->blas_fma = 1;
    p->h_vdw = 1.2;                      /* config.py:23-25 */
    p->dist_max = 4.5;                   /* config.py:592-660 */
    p->hbond_polar_dist = 3.5;
    p->weak_polar_dist = 3.5;
    p->ionic_dist = 4.0;
    p->carbonyl_dist = 3.6;
    p->aromatic_dist = 4.0;
    p->hydrophobic_dist = 4.5;
    p->metal_dist = 2.8;
    p->hbond_angle = 1.57;
    p->weak_hbond_angle = 2.27;
    p->cx_angle_min = 0.52;
    p->cx_angle_max = 2.62;
    p->xbond_angle = 2.09;
    p->ring_centroid_dist = 6.0;
    p->atom_ring_dist = 4.5;
    p->met_sulphur_dist = 6.0;
    p->amide_centroid_dist = 6.0;
    p->plane_bins_deg[0] = 30.0; p->plane_bins_deg[1] = 60.0; p->plane_bins_deg[2] = 90.0;
    cosine_images(p);
    return ARP_OK;
}

int arp_create(int device, arp_ctx** out)
{
    if (!out) return ARP_E_INVALID_ARG;
    *out = nullptr;
    int n = arp_device_count();
    if (n <= 0) {
        snprintf(g_create_err, ARP_ERRLEN, "no CUDA device available (there is no CPU fallback)");
        return n < 0 ? ARP_E_CUDA : ARP_E_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        snprintf(g_create_err, ARP_ERRLEN, "device %d out of range (0..%d)", device, n - 1);
        return ARP_E_INVALID_ARG;
    }
    arp_ctx* c = new (std::nothrow) arp_ctx();
    if (!c) return ARP_E_OOM;
    c->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int k = 0; k < 5 && e == cudaSuccess; ++k) e = cudaEventCreate(&c->ev[k]);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&c->h_meta, sizeof(RunMeta));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        snprintf(g_create_err, ARP_ERRLEN, "%s", cudaGetErrorString(e));
        arp_destroy(c);
        return ARP_E_CUDA;
    }
    memset(c->h_meta, 0, sizeof(RunMeta));
    if (getenv("ARPEGGIO_NO_PLANE_GRID")) c->use_plane_grid = 0;     /* A-B knob: plain double loops for the plane terms */
    if (getenv("ARPEGGIO_NO_WITHIN_GRID")) c->use_within_grid = 0;    /* A-B knob: binding-site flags by the double loop */
    if (getenv("ARPEGGIO_NO_PDL")) c->use_pdl = 0;                   /* A-B knob: plain stream-ordered launches */
    if (getenv("ARPEGGIO_NO_FUSED_GRID")) c->use_fused_grid = 0;
    if (getenv("ARPEGGIO_TILES")) c->use_tiles = 1;                                     /* A-B knob: the fused tile kernel instead of k_search + k_classify */
    if (const char* tx = getenv("ARPEGGIO_TILE_X")) { int v = atoi(tx); if (v >= 1 && v <= ARP_TILE_XMAX) c->tile_x = v; }
    if (getenv("ARPEGGIO_NO_EARLY_CLASSIFY")) c->use_early_cls = 0;                     /* A-B knob: k_classify waits for k_search */
    if (getenv("ARPEGGIO_NO_REG_GRID") && c->use_fused_grid > 1) c->use_fused_grid = 1;     /* debugging / A-B knob: five-kernel grid build */
    memset(&c->stats, 0, sizeof c->stats);
    arp_params_default(&c->params);
    derive_rule_params(c->params, &c->rp);
    c->have_params = 1;
    *out = c;
    return ARP_OK;
}

void arp_destroy(arp_ctx* c)
{
    if (!c) return;
    if (getenv("ARPEGGIO_HOST_TIMING")) {         /* diagnostic: host time inside the entry points of the end-to-end step */
        static const char* nm[5] = { "upload_atoms (enqueue)", "pairs_run_async (enqueue)", "fetch_packed: wait for the run",
                                     "fetch_packed: enqueue sort + copies", "fetch_packed: wait for the copies" };
        for (int k = 0; k < 5; ++k)
            if (c->ht_calls[k]) fprintf(stderr, "[host timing] %-38s %8llu calls  %8.1f us/call\n", nm[k], c->ht_calls[k], c->ht_ns[k] / 1e3 / c->ht_calls[k]);
    }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DBuf* bufs[] = { &c->xyz, &c->feat, &c->res_id, &c->rad_class, &c->vdw, &c->cov, &c->res_prev, &c->res_next,
                     &c->res_flags, &c->bond_off, &c->bond_nbr, &c->h_off, &c->h_xyz, &c->xnbr, &c->struct_off,
                     &c->zero, &c->geom, &c->cell_start, &c->cell_of, &c->rank, &c->pos4, &c->att4, &c->runtab, &c->sift_acc, &c->sift_out, &c->ring_scratch, &c->hreach, &c->arena, &c->out, &c->hits, &c->work,
                     &c->batch_small, &c->batch_stage, &c->w_bcnt, &c->w_hcnt, &c->w_hfix, &c->w_xidx, &c->w_xnbr, &c->wire_sums, &c->w_cnt_merge, &c->radtab, &c->sort_tmp, &c->sort_out, &c->sort_c, &c->sort_d, &c->sort_lo, &c->sort_hi, &c->sort_zero, &c->sort_off, &c->within, &c->flush };
    for (DBuf* b : bufs) dbuf_free(*b);
    arp_planes_release(c);
    for (int k = 0; k < 5; ++k) if (c->ev[k]) cudaEventDestroy(c->ev[k]);
    if (c->h_meta) cudaFreeHost(c->h_meta);
    if (c->h_batch) cudaFreeHost(c->h_batch);
    if (c->ev_batch) cudaEventDestroy(c->ev_batch);
    if (c->stream) cudaStreamDestroy(c->stream);
    (void)cudaGetLastError();
    delete c;
}

const char* arp_last_error(arp_ctx* c) { return c ? c->err : g_create_err; }

int arp_set_params(arp_ctx* c, const arp_params* p)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, p != nullptr, ARP_E_INVALID_ARG, "params is NULL");
    ARP_REQUIRE(c, p->interacting_cutoff >= 0.0 && p->interacting_cutoff < 1e6 && p->vdw_comp == p->vdw_comp,
                ARP_E_INVALID_ARG, "interacting_cutoff / vdw_comp out of range");
    c->params = *p;
    derive_rule_params(c->params, &c->rp);
    c->have_params = 1;
    c->radtab_valid = 0;
    pairs_invalidate(c);
    c->ring_ring.valid = c->atom_ring.valid = c->amide_amide.valid = c->amide_ring.valid = 0;
    return ARP_OK;
}

/* CPUs local to the calling thread's current CUDA device (its PCI function's local_cpulist in sysfs), or an empty set */
static bool device_local_cpus(cpu_set_t* set)
{
    CPU_ZERO(set);
    int dev = 0;
    char bdf[32] = {0}, path[128], line[4096];
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bdf, sizeof bdf, dev) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    for (char* p = bdf; *p; ++p) if (*p >= 'A' && *p <= 'F') *p += 'a' - 'A';
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bdf);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    const bool ok = fgets(line, sizeof line, f) != nullptr;
    fclose(f);
    if (!ok) return false;
    int n = 0;
    for (char* p = line; *p && *p != '\n'; ) {                  /* "0-15,32-47" */
        char* end = nullptr;
        long lo = strtol(p, &end, 10), hi = lo;
        if (end == p) break;
        p = end;
        if (*p == '-') { hi = strtol(p + 1, &end, 10); if (end == p + 1) break; p = end; }
        for (long c = lo; c <= hi && c < CPU_SETSIZE; ++c) if (c >= 0) { CPU_SET((int)c, set); ++n; }
        if (*p == ',') ++p;
    }
    return n > 0;
}

int arp_host_alloc(void** ptr, uint64_t bytes)
{
    if (!ptr) return ARP_E_INVALID_ARG;
    /* ARPEGGIO_NUMA_PIN=1: the buffer is the target of the device's DMA, so place it on the NUMA node of the device:
       pages land where the allocating thread runs (first touch), so the thread is confined to the device-local CPUs
       for the duration of the allocation and released again.  Off by default: on the boxes measured (8 GPUs behind
       ONE NUMA node, profiles/README.md round 2) it changes nothing -- the end-to-end figures stop scaling on the
       shared host PCIe / memory fabric, which the `pcie` probe of bench.py records. */
    cpu_set_t before, local, both;
    bool confined = false;
    if (getenv("ARPEGGIO_NUMA_PIN") && sched_getaffinity(0, sizeof before, &before) == 0 && device_local_cpus(&local)) {
        CPU_AND(&both, &before, &local);
        if (CPU_COUNT(&both) > 0 && !CPU_EQUAL(&both, &before)) confined = sched_setaffinity(0, sizeof both, &both) == 0;
    }
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 16);
    if (confined) sched_setaffinity(0, sizeof before, &before);
    if (e != cudaSuccess) { (void)cudaGetLastError(); *ptr = nullptr; return e == cudaErrorMemoryAllocation ? ARP_E_OOM : ARP_E_CUDA; }
    return ARP_OK;
}

int arp_host_free(void* ptr)
{
    if (!ptr) return ARP_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return ARP_E_CUDA; }
    return ARP_OK;
}

int arp_upload_atoms(arp_ctx* c, const arp_atoms* a)
{
    if (!c) return ARP_E_INVALID_ARG;
    HostTimer ht(c, 0);
    ARP_REQUIRE(c, a != nullptr, ARP_E_INVALID_ARG, "atoms is NULL");
    ARP_REQUIRE(c, a->n_atoms >= 0 && a->n_residues >= 0 && a->n_rad_classes >= 0, ARP_E_INVALID_ARG, "negative size");
    ARP_REQUIRE(c, a->n_atoms <= 500000000, ARP_E_INVALID_ARG, "more than 5e8 atoms");
    ARP_REQUIRE(c, a->n_rad_classes <= ARPK_MAX_RAD, ARP_E_INVALID_ARG, "more than 512 radius classes");
    const int N = a->n_atoms;
    const int S = a->n_structures > 0 ? a->n_structures : 1;
    ARP_REQUIRE(c, S == 1 || a->struct_off != nullptr, ARP_E_INVALID_ARG, "struct_off required for a batch");
    if (N > 0) {
        ARP_REQUIRE(c, a->xyz && a->feat && a->res_id && a->rad_class && a->vdw && a->cov && a->res_prev &&
                       a->res_next && a->res_flags, ARP_E_INVALID_ARG, "a required atom array is NULL");
        ARP_REQUIRE(c, a->n_residues > 0 && a->n_rad_classes > 0, ARP_E_INVALID_ARG, "atoms without residues or radius classes");
    }
    WireSizes w;
    ARP_TRY(wire_sizes(c, a, &w));
    ARP_REQUIRE(c, !a->bond_off || N == 0 || a->bond_off[0] == 0, ARP_E_INVALID_ARG, "bond_off[0] != 0");
    ARP_REQUIRE(c, !a->h_off || N == 0 || a->h_off[0] == 0, ARP_E_INVALID_ARG, "h_off[0] != 0");
    if (a->struct_off) {
        ARP_REQUIRE(c, a->struct_off[0] == 0 && a->struct_off[S] == N, ARP_E_INVALID_ARG, "struct_off must span the atoms");
        for (int s = 0; s < S; ++s)
            ARP_REQUIRE(c, a->struct_off[s] <= a->struct_off[s + 1], ARP_E_INVALID_ARG, "struct_off must ascend");
    }
    ARP_TRY(arp_bind(c));
    c->have_atoms = 0; pairs_invalidate(c);
    /* the radius-sum table depends on (vdw, cov, vdw_comp) only: a structure with the same radius tables keeps it */
    {
        const size_t kb = (size_t)a->n_rad_classes * sizeof(double);
        const bool same = c->radtab_valid && N > 0 && c->rad_host.size() == 2 * (size_t)a->n_rad_classes &&
                          memcmp(c->rad_host.data(), a->vdw, kb) == 0 && memcmp(c->rad_host.data() + a->n_rad_classes, a->cov, kb) == 0;
        if (!same) {
            c->radtab_valid = 0;
            c->rad_host.clear();
            if (N > 0) { c->rad_host.insert(c->rad_host.end(), a->vdw, a->vdw + a->n_rad_classes); c->rad_host.insert(c->rad_host.end(), a->cov, a->cov + a->n_rad_classes); }
        }
    }
    c->atom_ring.valid = 0;
    c->input_bytes = 0;
    c->N = N; c->Rs = a->n_residues; c->K = a->n_rad_classes; c->S = S;
    c->max_struct_atoms = N;
    if (S > 1) {
        c->max_struct_atoms = 0;
        for (int s = 0; s < S; ++s) { const int m = a->struct_off[s + 1] - a->struct_off[s]; if (m > c->max_struct_atoms) c->max_struct_atoms = m; }
    }
    c->has_bonds = a->bond_off != nullptr || w.cnt_b;
    c->has_h = a->h_off != nullptr || w.cnt_h;
    c->has_xnbr = a->xnbr_xyz != nullptr || w.sparse_x;
    ARP_REQUIRE(c, w.E < (1ll << 31) && w.H < (1ll << 31), ARP_E_INVALID_ARG, "CSR size beyond int32");
    c->E = (int)w.E;
    c->H = (int)w.H;
    const size_t n = (size_t)N;
    struct Piece { DBuf* dst; const void* src; size_t bytes; };
    Piece pieces[18];
    int np = 0;
    auto piece = [&](DBuf& d, const void* src, size_t bytes) { pieces[np++] = Piece{ &d, src, bytes }; };
    piece(c->xyz, a->xyz, n * 12);
    piece(c->feat, a->feat, n * 4);
    piece(c->res_id, a->res_id, n * 4);
    piece(c->rad_class, a->rad_class, n * 2);
    piece(c->vdw, a->vdw, (size_t)c->K * 8);
    piece(c->cov, a->cov, (size_t)c->K * 8);
    piece(c->res_prev, a->res_prev, (size_t)c->Rs * 4);
    piece(c->res_next, a->res_next, (size_t)c->Rs * 4);
    piece(c->res_flags, a->res_flags, (size_t)c->Rs);
    /* the wire forms land in staging buffers and are decoded below into the arrays the kernels read */
    if (c->has_bonds) {
        if (w.cnt_b) piece(c->w_bcnt, a->bond_cnt, n); else piece(c->bond_off, a->bond_off, (n + 1) * 4);
        piece(c->bond_nbr, a->bond_nbr, (size_t)c->E * 4);
    }
    if (c->has_h) {
        if (w.cnt_h) piece(c->w_hcnt, a->h_cnt, n); else piece(c->h_off, a->h_off, (n + 1) * 4);
        if (w.fix_h) piece(c->w_hfix, a->h_fix, (size_t)c->H * 12); else piece(c->h_xyz, a->h_xyz, (size_t)c->H * 24);
    }
    if (c->has_xnbr) {
        if (w.sparse_x) { piece(c->w_xidx, a->xnbr_idx, (size_t)w.X * 4); piece(c->w_xnbr, a->xnbr_xyz, (size_t)w.X * 12); }
        else piece(c->xnbr, a->xnbr_xyz, n * 12);
    }
    if (S > 1) piece(c->struct_off, a->struct_off, (size_t)(S + 1) * 4);
    /* One host block (e.g. engine.pinned_soa: every array at a 256-byte offset of one pinned allocation) goes up in
       ONE copy and the device arrays become views of the arena: 14 small DMA transfers cost about 110 us for a
       20k-atom structure, one transfer of the same 1.4 MB about 35 us.  Otherwise one copy per array. */
    uintptr_t lo = UINTPTR_MAX, hi = 0;
    size_t sum = 0;
    bool packed = N > 0;
    for (int k = 0; k < np; ++k) {
        if (!pieces[k].bytes) continue;
        const uintptr_t p0 = (uintptr_t)pieces[k].src;
        lo = p0 < lo ? p0 : lo;
        hi = p0 + pieces[k].bytes > hi ? p0 + pieces[k].bytes : hi;
        sum += pieces[k].bytes;
    }
    for (int k = 0; k < np && packed; ++k)
        if (pieces[k].bytes && ((uintptr_t)pieces[k].src - lo) % 16 != 0) packed = false;
    if (packed && (hi - lo) > sum + 512 * (size_t)np) packed = false;          /* gaps beyond alignment padding */
    for (int k = 0; k < np && packed; ++k)                                      /* no overlaps: aliasing arrays stay separate */
        for (int j = 0; j < k; ++j) {
            const uintptr_t a0 = (uintptr_t)pieces[k].src, a1 = a0 + pieces[k].bytes;
            const uintptr_t b0 = (uintptr_t)pieces[j].src, b1 = b0 + pieces[j].bytes;
            if (pieces[k].bytes && pieces[j].bytes && a0 < b1 && b0 < a1) packed = false;
        }
    if (packed) {
        ARP_TRY(dbuf_reserve(c, c->arena, hi - lo));
        ARP_CUDA(c, cudaMemcpyAsync(c->arena.p, (const void*)lo, hi - lo, cudaMemcpyHostToDevice, c->stream));
        for (int k = 0; k < np; ++k) {
            DBuf& d = *pieces[k].dst;
            dbuf_free(d);
            if (pieces[k].bytes) {
                d.p = c->arena.as<char>() + ((uintptr_t)pieces[k].src - lo);
                d.view = true;
            } else {
                ARP_TRY(dbuf_reserve(c, d, 0));          /* kernels may form (never dereference) the address */
            }
        }
        c->input_bytes = sum;
    } else {
        for (int k = 0; k < np; ++k) ARP_TRY(upload(c, *pieces[k].dst, pieces[k].src, pieces[k].bytes));
    }
    if (w.cnt_b) ARP_TRY(dbuf_reserve(c, c->bond_off, (n + 1) * 4));
    if (w.cnt_h) ARP_TRY(dbuf_reserve(c, c->h_off, (n + 1) * 4));
    if (w.cnt_b || w.cnt_h)
        ARP_TRY(counts_to_offsets<uint8_t>(c, w.cnt_b ? c->w_bcnt.as<uint8_t>() : nullptr, c->bond_off.as<int>(),
                                           w.cnt_h ? c->w_hcnt.as<uint8_t>() : nullptr, c->h_off.as<int>(), N));
    if (w.fix_h) ARP_TRY(dbuf_reserve(c, c->h_xyz, (size_t)c->H * 24));
    if (w.sparse_x) ARP_TRY(dbuf_reserve(c, c->xnbr, n * 12));
    const long long h3 = w.fix_h ? 3ll * c->H : 0, xn = w.sparse_x ? (long long)N : 0;
    if (h3 > 0 || xn > 0) {
        const long long m = h3 > xn ? h3 : xn;
        k_wire_rest<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(c->w_hfix.as<int32_t>(), a->h_fix_scale, c->h_xyz.as<double>(), h3,
                                                                        c->feat.as<uint32_t>(), c->w_xidx.as<int32_t>(), c->w_xnbr.as<float>(),
                                                                        w.sparse_x ? c->xnbr.as<float>() : nullptr, (int)w.X, N);
        ARP_LAUNCHED(c);
    }
    ARP_TRY(arp_pairs_prepare(c));
    c->have_atoms = 1;
    return ARP_OK;
}

static int pairs_out_reserve(arp_ctx* c, uint64_t records)
{
    if (records <= c->out_cap && c->out.p) return ARP_OK;
    ARP_TRY(dbuf_reserve(c, c->out, (size_t)records * sizeof(arp_pair)));
    c->out_cap = c->out.cap / sizeof(arp_pair);
    if (!c->use_tiles) {   /* the candidate list is all zero between runs (k_classify zeroes what it reads): a new allocation starts so */
        const void* before = c->hits.p;
        const size_t cap_before = c->hits.cap;
        ARP_TRY(dbuf_reserve(c, c->hits, (size_t)c->out_cap * sizeof(uint2)));
        if (c->hits.p != before || c->hits.cap != cap_before) {
            ARP_CUDA(c, cudaMemsetAsync(c->hits.p, 0, c->hits.cap, c->stream));
            c->hits_dirty = 0;
        }
    }
    if (c->work_cap < c->out_cap / 2 + 1024) {       /* first guess: one deferred predicate per two records */
        ARP_TRY(dbuf_reserve(c, c->work, (size_t)(c->out_cap / 2 + 1024) * 16));
        c->work_cap = c->work.cap / 16;
    }
    return ARP_OK;
}

static void fill_stats(arp_ctx* c, int with_events)
{
    arp_stats& s = c->stats;
    s.n_pairs = c->h_meta->n_pairs;
    s.n_candidates = c->h_meta->n_candidates;
    s.n_cells = c->h_meta->n_cells;
    s.n_cells_nonempty = c->h_meta->n_cells_nonempty;
    s.input_bytes = c->input_bytes;
    s.output_bytes = s.n_pairs * sizeof(arp_pair);
    s.faults = c->h_meta->fault;
    if (with_events > c->events_level) with_events = c->events_level;      /* only what the run recorded */
    if (with_events >= 3) {
        float a = 0.f, b = 0.f, d = 0.f;
        cudaEventElapsedTime(&a, c->ev[0], c->ev[1]);
        cudaEventElapsedTime(&b, c->ev[1], c->ev[2]);
        cudaEventElapsedTime(&d, c->ev[2], c->ev[3]);
        float h = 0.f;
        cudaEventElapsedTime(&h, c->ev[4], c->ev[3]);
        s.ms_grid = a; s.ms_search = b; s.ms_classify = d; s.ms_hscan = h; s.ms_pairs = b + d; s.ms_total = a + b + d;
    } else if (with_events >= 1) {
        float w = 0.f;
        cudaEventElapsedTime(&w, c->ev[0], c->ev[3]);
        s.ms_total = w; s.ms_grid = s.ms_search = s.ms_classify = s.ms_hscan = s.ms_pairs = 0.f;
    } else {
        s.ms_total = s.ms_grid = s.ms_search = s.ms_classify = s.ms_hscan = s.ms_pairs = 0.f;      /* a run enqueued without events */
    }
}

/* waits for the enqueued run and repeats it with larger buffers when the record stream or the work list overflowed
   (an overflowing run still counts, so one repetition is enough) */
static int pairs_finish(arp_ctx* c)
{
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (attempt > 0) { ++c->finish_reruns; ARP_TRY(arp_pairs_enqueue(c, 1)); }
        ARP_CUDA(c, cudaStreamSynchronize(c->stream));
        c->run_pending = 0;
        /* candidates >= records: both lists share the capacity (k_tiles has no candidate list: n_raw stays 0) */
        const uint64_t n = c->h_meta->n_raw > c->h_meta->n_pairs ? c->h_meta->n_raw : c->h_meta->n_pairs;
        const uint64_t nw = c->h_meta->n_work + c->h_meta->n_work_rare;     /* the list is filled from both ends */
        if (c->h_meta->fault & 2u) {
            /* k_classify gave up waiting for a candidate that k_search had announced (slow producer: time slicing,
               a debugger).  Not an error: clean the list and repeat the run with k_classify waiting for k_search. */
            if (c->hits.p) ARP_CUDA(c, cudaMemsetAsync(c->hits.p, 0, c->hits.cap, c->stream));
            ARP_REQUIRE(c, attempt < 2, ARP_E_CUDA, "candidate hand-off between k_search and k_classify timed out repeatedly");
            c->use_early_cls = 0;
            continue;
        }
        if (n <= c->out_cap && nw <= c->work_cap) break;
        ARP_REQUIRE(c, attempt < 2, ARP_E_CAPACITY, "record stream overflowed repeatedly");
        if (n > c->out_cap) ARP_TRY(pairs_out_reserve(c, n + n / 16 + 1024));
        else {                                  /* the work-item count is exact once the records fit */
            ARP_TRY(dbuf_reserve(c, c->work, (size_t)(nw + nw / 16 + 1024) * 16));
            c->work_cap = c->work.cap / 16;
        }
    }
    ARP_REQUIRE(c, c->h_meta->n_pairs < (1ull << 32), ARP_E_CAPACITY, "more than 2^32 records in one run");
    c->n_pairs = c->h_meta->n_pairs;
    c->pairs_valid = 1;
    fill_stats(c, 1);
    return ARP_OK;
}

int arp_upload_atoms_batch(arp_ctx* c, const arp_atoms* const* parts, int32_t n_parts)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, parts != nullptr && n_parts >= 1, ARP_E_INVALID_ARG, "no structures");
    long long N = 0, Rs = 0, E = 0, H = 0, X = 0;
    bool any_bonds = false, any_h = false, any_x = false, cnt_bonds = false, cnt_h = false;
    std::vector<WireSizes> ws((size_t)n_parts);
    for (int s = 0; s < n_parts; ++s) {
        ARP_TRY(check_part(c, parts[s]));
        const arp_atoms* a = parts[s];
        WireSizes& w = ws[(size_t)s];
        ARP_TRY(wire_sizes(c, a, &w));
        N += a->n_atoms; Rs += a->n_atoms > 0 ? a->n_residues : 0;
        E += w.E; H += w.H; X += w.X;
        if (a->n_atoms > 0 && (a->bond_off || w.cnt_b)) { any_bonds = true; cnt_bonds |= w.cnt_b; }
        if (a->n_atoms > 0 && (a->h_off || w.cnt_h)) { any_h = true; cnt_h |= w.cnt_h; }
        if (a->n_atoms > 0 && (a->xnbr_xyz || w.sparse_x)) any_x = true;
    }
    ARP_REQUIRE(c, N <= 500000000 && E < (1ll << 31) && H < (1ll << 31), ARP_E_INVALID_ARG, "batch too large");
    ARP_TRY(arp_bind(c));
    c->have_atoms = 0; pairs_invalidate(c); c->radtab_valid = 0; c->rad_host.clear();
    c->atom_ring.valid = 0;
    /* ---- host tables: part descriptors, struct_off, merged radius table + class maps (one pinned block) ---- */
    std::vector<double> vdw, cov;
    std::vector<unsigned short> cmap;
    std::vector<PartDesc> desc((size_t)n_parts);
    std::vector<int32_t> soff((size_t)n_parts + 1);
    size_t stage_bytes = 0;
    auto up256 = [](size_t v) { return (v + 255) / 256 * 256; };
    struct Copy { size_t dst; const void* src; size_t bytes; };
    std::vector<Copy> copies;
    long long ab = 0, rb = 0, bb = 0, hb = 0, xb = 0;
    c->input_bytes = 0;
    for (int s = 0; s < n_parts; ++s) {
        const arp_atoms* a = parts[s];
        PartDesc& d = desc[(size_t)s];
        memset(&d, 0, sizeof d);
        soff[(size_t)s] = (int32_t)ab;
        const size_t n = (size_t)a->n_atoms, nr = n ? (size_t)a->n_residues : 0;
        const WireSizes& w = ws[(size_t)s];
        const size_t ne = (size_t)w.E, nh = (size_t)w.H, nx = (size_t)w.X;
        d.atom_base = (int)ab; d.n_atoms = (int)n; d.res_base = (int)rb; d.n_res = (int)nr; d.bond_base = (int)bb; d.h_base = (int)hb;
        d.x_base = (int)xb; d.h_scale = a->h_fix_scale;
        d.class_base = (int)cmap.size();
        for (int k = 0; n && k < a->n_rad_classes; ++k) {          /* merged radius table: identical (vdw, cov) pairs share a class */
            size_t m = 0;
            while (m < vdw.size() && !(memcmp(&vdw[m], &a->vdw[k], 8) == 0 && memcmp(&cov[m], &a->cov[k], 8) == 0)) ++m;
            if (m == vdw.size()) { vdw.push_back(a->vdw[k]); cov.push_back(a->cov[k]); }
            cmap.push_back((unsigned short)m);
        }
        struct Piece { long long* off; const void* src; size_t bytes; };
        Piece pc[16] = {
            { &d.o_xyz, a->xyz, n * 12 }, { &d.o_feat, a->feat, n * 4 }, { &d.o_res_id, a->res_id, n * 4 }, { &d.o_rad, a->rad_class, n * 2 },
            { &d.o_prev, a->res_prev, nr * 4 }, { &d.o_next, a->res_next, nr * 4 }, { &d.o_flags, a->res_flags, nr },
            { &d.o_boff, n ? a->bond_off : nullptr, (n && a->bond_off) ? (n + 1) * 4 : 0 }, { &d.o_bnbr, a->bond_nbr, ne * 4 },
            { &d.o_hoff, n ? a->h_off : nullptr, (n && a->h_off) ? (n + 1) * 4 : 0 }, { &d.o_hxyz, w.fix_h ? nullptr : a->h_xyz, w.fix_h ? 0 : nh * 24 },
            { &d.o_xnbr, n ? a->xnbr_xyz : nullptr, w.sparse_x ? nx * 12 : ((n && a->xnbr_xyz) ? n * 12 : 0) },
            { &d.o_bcnt, a->bond_cnt, w.cnt_b ? n : 0 }, { &d.o_hcnt, a->h_cnt, w.cnt_h ? n : 0 },
            { &d.o_hfix, a->h_fix, w.fix_h ? nh * 12 : 0 }, { &d.o_xidx, a->xnbr_idx, w.sparse_x ? nx * 4 : 0 } };
        /* one host block (engine.pinned_soa) -> one DMA for the structure; otherwise one per array */
        uintptr_t lo = UINTPTR_MAX, hi = 0;
        size_t sum = 0;
        for (const Piece& q : pc) if (q.bytes) {
            const uintptr_t p0 = (uintptr_t)q.src;
            lo = p0 < lo ? p0 : lo; hi = p0 + q.bytes > hi ? p0 + q.bytes : hi; sum += q.bytes;
        }
        bool packed = sum > 0 && (hi - lo) <= sum + 512 * 16;
        for (const Piece& q : pc) if (q.bytes && ((uintptr_t)q.src - lo) % 8 != 0) packed = false;

        if (packed) {
            copies.push_back(Copy{ stage_bytes, (const void*)lo, hi - lo });
            for (const Piece& q : pc) *q.off = q.bytes ? (long long)(stage_bytes + ((uintptr_t)q.src - lo)) : -1;
            stage_bytes = up256(stage_bytes + (hi - lo));
        } else {
            for (const Piece& q : pc) {
                *q.off = q.bytes ? (long long)stage_bytes : -1;
                if (q.bytes) { copies.push_back(Copy{ stage_bytes, q.src, q.bytes }); stage_bytes = up256(stage_bytes + q.bytes); }
            }
        }
        c->input_bytes += sum + (n ? (size_t)a->n_rad_classes * 16 : 0);
        ab += (long long)n; rb += (long long)nr; bb += (long long)ne; hb += (long long)nh; xb += (long long)nx;
    }
    soff[(size_t)n_parts] = (int32_t)ab;
    ARP_REQUIRE(c, vdw.size() <= ARPK_MAX_RAD, ARP_E_INVALID_ARG, "more than 512 distinct radius classes in the batch");
    const int K = (int)vdw.size();
    /* small tables in one pinned block, kept until the next upload */
    const size_t o_desc = 0, o_soff = up256(o_desc + desc.size() * sizeof(PartDesc)), o_vdw = up256(o_soff + soff.size() * 4);
    const size_t o_cov = up256(o_vdw + (size_t)(K ? K : 1) * 8), o_cmap = up256(o_cov + (size_t)(K ? K : 1) * 8);
    const size_t small = up256(o_cmap + (cmap.size() ? cmap.size() : 1) * 2);
    if (c->h_batch_cap < small) {
        if (c->ev_batch) ARP_CUDA(c, cudaEventSynchronize(c->ev_batch));      /* the old image may still be a copy's source */
        if (c->h_batch) cudaFreeHost(c->h_batch);
        c->h_batch = nullptr; c->h_batch_cap = 0;
        ARP_CUDA(c, cudaMallocHost(&c->h_batch, small + small / 2));
        c->h_batch_cap = small + small / 2;
    }
    /* the pinned image may still be the source of the previous batch's copy: wait for that copy (not for the stream) */
    if (!c->ev_batch) ARP_CUDA(c, cudaEventCreateWithFlags(&c->ev_batch, cudaEventDisableTiming));
    else ARP_CUDA(c, cudaEventSynchronize(c->ev_batch));
    char* hb8 = (char*)c->h_batch;
    memcpy(hb8 + o_desc, desc.data(), desc.size() * sizeof(PartDesc));
    memcpy(hb8 + o_soff, soff.data(), soff.size() * 4);
    if (K) { memcpy(hb8 + o_vdw, vdw.data(), (size_t)K * 8); memcpy(hb8 + o_cov, cov.data(), (size_t)K * 8); }
    if (!cmap.empty()) memcpy(hb8 + o_cmap, cmap.data(), cmap.size() * 2);
    ARP_TRY(dbuf_reserve(c, c->batch_small, small));
    ARP_CUDA(c, cudaMemcpyAsync(c->batch_small.p, hb8, small, cudaMemcpyHostToDevice, c->stream));
    ARP_CUDA(c, cudaEventRecord(c->ev_batch, c->stream));
    ARP_TRY(dbuf_reserve(c, c->batch_stage, stage_bytes));
    for (const Copy& q : copies)
        ARP_CUDA(c, cudaMemcpyAsync(c->batch_stage.as<char>() + q.dst, q.src, q.bytes, cudaMemcpyHostToDevice, c->stream));
    /* ---- the merged arrays ---- */
    c->N = (int)N; c->Rs = (int)Rs; c->K = K; c->S = n_parts; c->E = (int)E; c->H = (int)H;
    c->max_struct_atoms = 0;
    for (int s = 0; s < n_parts; ++s) if (parts[s]->n_atoms > c->max_struct_atoms) c->max_struct_atoms = parts[s]->n_atoms;
    c->has_bonds = any_bonds; c->has_h = any_h; c->has_xnbr = any_x;
    const size_t n = (size_t)N;
    ARP_TRY(dbuf_reserve(c, c->xyz, n * 12)); ARP_TRY(dbuf_reserve(c, c->feat, n * 4)); ARP_TRY(dbuf_reserve(c, c->res_id, n * 4));
    ARP_TRY(dbuf_reserve(c, c->rad_class, n * 2)); ARP_TRY(dbuf_reserve(c, c->res_prev, (size_t)Rs * 4));
    ARP_TRY(dbuf_reserve(c, c->res_next, (size_t)Rs * 4)); ARP_TRY(dbuf_reserve(c, c->res_flags, (size_t)Rs));
    if (any_bonds) { ARP_TRY(dbuf_reserve(c, c->bond_off, (n + 1) * 4)); ARP_TRY(dbuf_reserve(c, c->bond_nbr, (size_t)E * 4)); }
    if (any_h) { ARP_TRY(dbuf_reserve(c, c->h_off, (n + 1) * 4)); ARP_TRY(dbuf_reserve(c, c->h_xyz, (size_t)H * 24)); }
    if (any_x) ARP_TRY(dbuf_reserve(c, c->xnbr, n * 12));
    /* vdw / cov / struct_off: views of the small device block */
    for (DBuf* b : { &c->vdw, &c->cov, &c->struct_off }) dbuf_free(*b);
    c->vdw.p = c->batch_small.as<char>() + o_vdw; c->vdw.view = true;
    c->cov.p = c->batch_small.as<char>() + o_cov; c->cov.view = true;
    c->struct_off.p = c->batch_small.as<char>() + o_soff; c->struct_off.view = true;
    if (N > 0) {
        MergeArgs M;
        memset(&M, 0, sizeof M);
        M.parts = (const PartDesc*)(c->batch_small.as<char>() + o_desc); M.n_parts = n_parts;
        M.stage = c->batch_stage.as<char>(); M.class_map = (const unsigned short*)(c->batch_small.as<char>() + o_cmap);
        M.N = (int)N; M.Rs = (int)Rs; M.E = (int)E; M.H = (int)H; M.X = (int)X;
        if (cnt_bonds || cnt_h) {
            ARP_TRY(dbuf_reserve(c, c->w_cnt_merge, n * 8));
            M.bond_cnt = cnt_bonds ? c->w_cnt_merge.as<int32_t>() : nullptr;
            M.h_cnt = cnt_h ? c->w_cnt_merge.as<int32_t>() + n : nullptr;
        }
        M.xyz = c->xyz.as<float>(); M.feat = c->feat.as<uint32_t>(); M.res_id = c->res_id.as<int32_t>();
        M.rad_class = c->rad_class.as<uint16_t>(); M.res_prev = c->res_prev.as<int32_t>(); M.res_next = c->res_next.as<int32_t>();
        M.res_flags = c->res_flags.as<uint8_t>();
        M.bond_off = any_bonds ? c->bond_off.as<int32_t>() : nullptr; M.bond_nbr = any_bonds ? c->bond_nbr.as<int32_t>() : nullptr;
        M.h_off = any_h ? c->h_off.as<int32_t>() : nullptr; M.h_xyz = any_h ? c->h_xyz.as<double>() : nullptr;
        M.xnbr = any_x ? c->xnbr.as<float>() : nullptr;
        k_merge_atoms<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(M);
        ARP_LAUNCHED(c);
        long long rest = Rs > E ? Rs : E;
        rest = H > rest ? H : rest;
        if (rest > 0) {
            k_merge_rest<<<(unsigned)((rest + 255) / 256), 256, 0, c->stream>>>(M);
            ARP_LAUNCHED(c);
        }
        if (X > 0) {
            k_merge_xnbr<<<(unsigned)((X + 255) / 256), 256, 0, c->stream>>>(M);
            ARP_LAUNCHED(c);
        }
        if (M.bond_cnt || M.h_cnt)
            ARP_TRY(counts_to_offsets<int32_t>(c, M.bond_cnt, M.bond_off, M.h_cnt, M.h_off, (int)N));
    }
    ARP_TRY(arp_pairs_prepare(c));
    c->have_atoms = 1;
    return ARP_OK;
}

static int pairs_run_enqueue(arp_ctx* c, int with_events)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->have_atoms, ARP_E_NOT_READY, "arp_pairs_run before arp_upload_atoms");
    ARP_TRY(arp_bind(c));
    HostTimer ht(c, 1);
    pairs_invalidate(c);
    /* first guess of the stream length; an overflowing run still counts, then is repeated once */
    uint64_t want = c->out_cap ? c->out_cap : (uint64_t)c->N * 16 + 4096;
    ARP_TRY(pairs_out_reserve(c, want));
    ARP_TRY(arp_pairs_enqueue(c, with_events));
    c->run_pending = 1;
    return ARP_OK;
}

/* no timing events around an asynchronous run: two API calls less per step (arp_stats.ms_* are 0 for such a run) */
int arp_pairs_run_async(arp_ctx* c) { return pairs_run_enqueue(c, 0); }

int arp_pairs_run(arp_ctx* c, uint64_t* n_pairs)
{
    ARP_TRY(pairs_run_enqueue(c, 1));
    ARP_TRY(pairs_finish(c));
    if (n_pairs) *n_pairs = c->n_pairs;
    return ARP_OK;
}

/* a run enqueued by arp_pairs_run_async is waited for by the first call that needs its result */
static int pairs_ready(arp_ctx* c, const char* what)
{
    if (c->run_pending) { ARP_TRY(arp_bind(c)); ARP_TRY(pairs_finish(c)); }
    ARP_REQUIRE(c, c->pairs_valid, ARP_E_NOT_READY, what);
    return ARP_OK;
}

int arp_pairs_fetch(arp_ctx* c, arp_pair* dst, uint64_t cap, int sorted)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_TRY(pairs_ready(c, "arp_pairs_fetch before arp_pairs_run"));
    ARP_REQUIRE(c, cap >= c->n_pairs, ARP_E_CAPACITY, "destination holds fewer records than the run produced");
    if (c->n_pairs == 0) return ARP_OK;
    ARP_REQUIRE(c, dst != nullptr, ARP_E_INVALID_ARG, "dst is NULL");
    ARP_TRY(arp_bind(c));
    const arp_pair* src = c->out.as<arp_pair>();
    if (sorted) {
        ARP_TRY(arp_pairs_sorted_build(c, 0));
        src = c->sort_out.as<arp_pair>();
    }
    ARP_CUDA(c, cudaMemcpyAsync(dst, src, (size_t)c->n_pairs * sizeof(arp_pair), cudaMemcpyDeviceToHost, c->stream));
    ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    return ARP_OK;
}

int arp_pairs_count(arp_ctx* c, uint64_t* n_pairs)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, n_pairs != nullptr, ARP_E_INVALID_ARG, "n_pairs is NULL");
    ARP_TRY(pairs_ready(c, "arp_pairs_count before arp_pairs_run"));
    *n_pairs = c->n_pairs;
    return ARP_OK;
}

int arp_pairs_fetch_compact(arp_ctx* c, uint32_t* row_off, arp_pair_c* rec, uint64_t cap, float* dist, uint64_t* n_pairs)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_TRY(pairs_ready(c, "arp_pairs_fetch_compact before arp_pairs_run"));
    if (n_pairs) *n_pairs = c->n_pairs;
    ARP_REQUIRE(c, cap >= c->n_pairs, ARP_E_CAPACITY, "destination holds fewer records than the run produced");
    ARP_REQUIRE(c, row_off != nullptr, ARP_E_INVALID_ARG, "row_off is NULL");
    ARP_REQUIRE(c, c->n_pairs == 0 || rec != nullptr, ARP_E_INVALID_ARG, "rec is NULL");
    ARP_TRY(arp_bind(c));
    ARP_TRY(arp_pairs_sorted_build(c, 1));
    ARP_CUDA(c, cudaMemcpyAsync(row_off, c->sort_off.p, ((size_t)c->N + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (c->n_pairs) {
        ARP_CUDA(c, cudaMemcpyAsync(rec, c->sort_c.p, (size_t)c->n_pairs * sizeof(arp_pair_c), cudaMemcpyDeviceToHost, c->stream));
        if (dist) ARP_CUDA(c, cudaMemcpyAsync(dist, c->sort_d.p, (size_t)c->n_pairs * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    return ARP_OK;
}

int arp_pairs_fetch_packed(arp_ctx* c, uint32_t* row_off, uint32_t* lo32, uint8_t* hi8, uint64_t cap, float* dist,
                           uint64_t* n_pairs, int32_t* bits_j, uint32_t* n_faults)
{
    if (!c) return ARP_E_INVALID_ARG;
    { HostTimer ht(c, 2); ARP_TRY(pairs_ready(c, "arp_pairs_fetch_packed before arp_pairs_run")); }
    HostTimer ht3(c, 3);
    const int bj = arp_pairs_bits_j(c);
    const bool need_hi = bj + 15 > 32;
    if (n_pairs) *n_pairs = c->n_pairs;
    if (bits_j) *bits_j = bj;
    ARP_REQUIRE(c, cap >= c->n_pairs, ARP_E_CAPACITY, "destination holds fewer records than the run produced");
    ARP_REQUIRE(c, row_off != nullptr && n_faults != nullptr, ARP_E_INVALID_ARG, "row_off or n_faults is NULL");
    ARP_REQUIRE(c, c->n_pairs == 0 || (lo32 != nullptr && (!need_hi || hi8 != nullptr)), ARP_E_INVALID_ARG,
                "lo32 is NULL (or hi8, which more than 131072 atoms need)");
    ARP_TRY(arp_bind(c));
    ARP_TRY(arp_pairs_sorted_build(c, 2));
    ARP_CUDA(c, cudaMemcpyAsync(row_off, c->sort_off.p, ((size_t)c->N + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    *n_faults = 0;
    if (c->n_pairs) {
        ARP_CUDA(c, cudaMemcpyAsync(lo32, c->sort_lo.p, (size_t)c->n_pairs * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        if (need_hi) ARP_CUDA(c, cudaMemcpyAsync(hi8, c->sort_hi.p, (size_t)c->n_pairs, cudaMemcpyDeviceToHost, c->stream));
        if (dist) ARP_CUDA(c, cudaMemcpyAsync(dist, c->sort_d.p, (size_t)c->n_pairs * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        ARP_CUDA(c, cudaMemcpyAsync(&c->h_meta->pad0[0], c->sort_fault, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
    }
    ht3.stop();
    { HostTimer ht(c, 4); ARP_CUDA(c, cudaStreamSynchronize(c->stream)); }
    if (c->n_pairs) *n_faults = c->h_meta->pad0[0];
    return ARP_OK;
}

/* The whole fetch enqueued behind a run that has not been waited for: the sorted packed view is built with the record
   count read on the device and the first min(cap, expect) words are copied; arp_pairs_fetch_packed_wait is the one wait of
   the step.  One host thread can so keep several contexts (streams) busy without ever blocking in the middle of a step. */
int arp_pairs_fetch_packed_async(arp_ctx* c, uint32_t* row_off, uint32_t* lo32, uint8_t* hi8, uint64_t cap, float* dist, uint64_t expect)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->have_atoms && (c->run_pending || c->pairs_valid), ARP_E_NOT_READY, "arp_pairs_fetch_packed_async before arp_pairs_run_async");
    ARP_REQUIRE(c, row_off != nullptr, ARP_E_INVALID_ARG, "row_off is NULL");          /* n_atoms + 2 entries: the last one is scratch */
    const bool need_hi = arp_pairs_bits_j(c) + 15 > 32;
    ARP_REQUIRE(c, cap == 0 || (lo32 != nullptr && (!need_hi || hi8 != nullptr)), ARP_E_INVALID_ARG,
                "lo32 is NULL (or hi8, which more than 131072 atoms need)");
    ARP_TRY(arp_bind(c));
    HostTimer ht(c, 3);
    const int blind = c->run_pending ? 1 : 0;
    uint64_t ncopy;
    if (blind) {
        ncopy = expect && expect < cap ? expect : cap;
        if (ncopy > c->out_cap) ncopy = c->out_cap;
    } else {
        ncopy = c->n_pairs <= cap ? c->n_pairs : 0;          /* a stream that does not fit is reported by the wait */
    }
    ARP_TRY(arp_pairs_sorted_build(c, 2, blind));
    /* row offsets [N + 1] and the fault counter behind them in one copy */
    ARP_CUDA(c, cudaMemcpyAsync(row_off, c->sort_off.p, ((size_t)c->N + 2) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    static const bool no_d2h = getenv("ARPEGGIO_DEBUG_NO_D2H") != nullptr;      /* diagnostic (tools/e2e_breakdown.py): the words stay on the device */
    if (no_d2h) ncopy = ncopy < 64 ? ncopy : 64;
    if (ncopy) {
        ARP_CUDA(c, cudaMemcpyAsync(lo32, c->sort_lo.p, (size_t)ncopy * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        if (need_hi) ARP_CUDA(c, cudaMemcpyAsync(hi8, c->sort_hi.p, (size_t)ncopy, cudaMemcpyDeviceToHost, c->stream));
        if (dist) ARP_CUDA(c, cudaMemcpyAsync(dist, c->sort_d.p, (size_t)ncopy * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (no_d2h) ncopy = c->out_cap;               /* pretend: the wait then has nothing left to fetch */
    c->pk.row_off = row_off; c->pk.lo32 = lo32; c->pk.hi8 = hi8; c->pk.cap = cap; c->pk.dist = dist;
    c->pk.copied = ncopy; c->pk.blind = blind; c->pk.pending = 1;
    return ARP_OK;
}

int arp_pairs_fetch_packed_wait(arp_ctx* c, uint64_t* n_pairs, int32_t* bits_j, uint32_t* n_faults)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->pk.pending, ARP_E_NOT_READY, "arp_pairs_fetch_packed_wait without arp_pairs_fetch_packed_async");
    ARP_REQUIRE(c, n_faults != nullptr, ARP_E_INVALID_ARG, "n_faults is NULL");
    ARP_TRY(arp_bind(c));
    HostTimer ht(c, 4);
    c->pk.pending = 0;
    if (c->pk.blind && c->run_pending) {
        c->finish_reruns = 0;
        ARP_TRY(pairs_finish(c));                 /* the wait; repeats a run that overflowed */
        if (c->finish_reruns == 0) { c->packed_valid = 1; c->sort_tmp_valid = 1; }      /* the view was built from the complete stream */
    } else {
        ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    ARP_REQUIRE(c, c->pairs_valid, ARP_E_NOT_READY, "the run was invalidated between arp_pairs_fetch_packed_async and the wait");
    if (!c->packed_valid)                         /* the run was repeated after the view had been built: the plain way */
        return arp_pairs_fetch_packed(c, c->pk.row_off, c->pk.lo32, c->pk.hi8, c->pk.cap, c->pk.dist, n_pairs, bits_j, n_faults);
    const int bj = arp_pairs_bits_j(c);
    if (n_pairs) *n_pairs = c->n_pairs;
    if (bits_j) *bits_j = bj;
    ARP_REQUIRE(c, c->pk.cap >= c->n_pairs, ARP_E_CAPACITY, "destination holds fewer records than the run produced");
    if (c->n_pairs > c->pk.copied) {              /* more records than expected: the rest of the streams */
        const size_t o = (size_t)c->pk.copied, m = (size_t)(c->n_pairs - c->pk.copied);
        ARP_CUDA(c, cudaMemcpyAsync(c->pk.lo32 + o, c->sort_lo.as<uint32_t>() + o, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        if (bj + 15 > 32) ARP_CUDA(c, cudaMemcpyAsync(c->pk.hi8 + o, c->sort_hi.as<uint8_t>() + o, m, cudaMemcpyDeviceToHost, c->stream));
        if (c->pk.dist) ARP_CUDA(c, cudaMemcpyAsync(c->pk.dist + o, c->sort_d.as<float>() + o, m * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    *n_faults = c->pk.row_off[(size_t)c->N + 1];
    return ARP_OK;
}

/* host only: the packed view back into 16-byte records.  feat: the ARP_F_* words of the uploaded atoms (the entity class
   is a function of their selection / water bits, rule_entity_class_bools = interactions.py:643-691) */
/* rows [r0, r1) of the packed view -> records */
static int unpack_packed_rows(const uint32_t* row_off, const uint32_t* lo32, const uint8_t* hi8, const float* dist, int32_t n_atoms,
                              int32_t bits_j, const uint32_t* feat, const int32_t* struct_off, int32_t n_structures, arp_pair* dst,
                              uint64_t n, int32_t r0, int32_t r1)
{
    const uint64_t jmask = (1ull << bits_j) - 1ull;
    uint32_t cls_of[16];                               /* branch-free: the six ifs as a table over (sel_i, sel_j, water_i, water_j) */
    for (int k = 0; k < 16; ++k) cls_of[k] = rule_entity_class_bools(k & 1, (k >> 1) & 1, (k >> 2) & 1, (k >> 3) & 1) << ARP_CLASS_SHIFT;
    int32_t s = 0, base = 0, end = n_atoms;            /* structure of row i: its first atom and the one behind its last */
    if (struct_off) {
        int lo = 0, hi = n_structures;                 /* last structure whose first atom is <= r0 */
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (struct_off[mid] <= r0) lo = mid; else hi = mid; }
        s = lo; base = struct_off[s]; end = struct_off[s + 1];
    }
    for (int32_t i = r0; i < r1; ++i) {
        while (struct_off && i >= end) { ++s; base = struct_off[s]; end = struct_off[s + 1]; if (end < base) return ARP_E_INVALID_ARG; }
        if (row_off[i + 1] < row_off[i] || row_off[i + 1] > n) return ARP_E_INVALID_ARG;
        const uint32_t fi = feat[i];
        const uint32_t row_bits = ((fi & ARP_F_IN_SELECTION) ? 1u : 0u) | ((fi & ARP_F_IS_WATER) ? 4u : 0u);
        const uint64_t k1 = row_off[i + 1];
        for (uint64_t k = row_off[i]; k < k1; ++k) {
            const uint64_t w = (uint64_t)lo32[k] | (hi8 ? (uint64_t)hi8[k] << 32 : 0ull);
            const int64_t j = (int64_t)(w & jmask) + base;              /* the word holds j local to the structure */
            if (j < base || j >= end) return ARP_E_INVALID_ARG;
            const uint32_t fj = feat[j];
            const uint32_t idx = row_bits | ((fj & ARP_F_IN_SELECTION) ? 2u : 0u) | ((fj & ARP_F_IS_WATER) ? 8u : 0u);
            dst[k].i = i; dst[k].j = (int32_t)j;
            dst[k].mask = (uint32_t)((w >> bits_j) & 0x7fffu) | cls_of[idx];
            dst[k].dist = dist ? dist[k] : 0.f;
        }
    }
    return ARP_OK;
}

int arp_pairs_unpack_packed(const uint32_t* row_off, const uint32_t* lo32, const uint8_t* hi8, const float* dist, int32_t n_atoms,
                            int32_t bits_j, const uint32_t* feat, const int32_t* struct_off, int32_t n_structures, arp_pair* dst, uint64_t cap,
                            int32_t threads)
{
    if (n_atoms < 0 || (n_atoms > 0 && (!row_off || !feat)) || bits_j < 1 || bits_j > 31) return ARP_E_INVALID_ARG;
    if (struct_off && (n_structures < 1 || struct_off[0] != 0 || struct_off[n_structures] != n_atoms)) return ARP_E_INVALID_ARG;
    if (n_atoms == 0) return ARP_OK;
    const uint64_t n = row_off[n_atoms];
    if (n > cap) return ARP_E_CAPACITY;
    if (n && (!lo32 || !dst || (bits_j + 15 > 32 && !hi8))) return ARP_E_INVALID_ARG;
    if (struct_off) for (int32_t s = 0; s < n_structures; ++s) if (struct_off[s + 1] < struct_off[s]) return ARP_E_INVALID_ARG;
    int T = threads > 1 ? threads : 1;
    if (T > 64) T = 64;
    if (n < 200000 || n_atoms < 4 * T) T = 1;          /* not worth a thread start */
    if (T == 1) return unpack_packed_rows(row_off, lo32, hi8, dist, n_atoms, bits_j, feat, struct_off, n_structures, dst, n, 0, n_atoms);
    /* rows cut where the record count crosses k * n / T (row_off must ascend for the search to mean anything: each share checks its rows) */
    std::vector<int32_t> cut((size_t)T + 1);
    cut[0] = 0; cut[(size_t)T] = n_atoms;
    for (int t = 1; t < T; ++t) {
        const uint64_t want = n / (uint64_t)T * (uint64_t)t;
        int32_t lo = 0, hi = n_atoms;
        while (lo < hi) { const int32_t mid = lo + (hi - lo) / 2; if (row_off[mid] < want) lo = mid + 1; else hi = mid; }
        cut[(size_t)t] = lo < cut[(size_t)t - 1] ? cut[(size_t)t - 1] : lo;
    }
    std::vector<int> rc((size_t)T, ARP_OK);
    std::vector<std::thread> pool;
    for (int t = 1; t < T; ++t)
        pool.emplace_back([&, t] { rc[(size_t)t] = unpack_packed_rows(row_off, lo32, hi8, dist, n_atoms, bits_j, feat, struct_off, n_structures, dst, n,
                                                                       cut[(size_t)t], cut[(size_t)t + 1]); });
    rc[0] = unpack_packed_rows(row_off, lo32, hi8, dist, n_atoms, bits_j, feat, struct_off, n_structures, dst, n, cut[0], cut[1]);
    for (std::thread& th : pool) th.join();
    for (int t = 0; t < T; ++t) if (rc[(size_t)t] != ARP_OK) return rc[(size_t)t];
    return ARP_OK;
}

int arp_pairs_fetch_dist(arp_ctx* c, float* dist, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_TRY(pairs_ready(c, "arp_pairs_fetch_dist before arp_pairs_run"));
    ARP_REQUIRE(c, cap >= c->n_pairs, ARP_E_CAPACITY, "destination holds fewer distances than the run produced");
    if (c->n_pairs == 0) return ARP_OK;
    ARP_REQUIRE(c, dist != nullptr, ARP_E_INVALID_ARG, "dist is NULL");
    ARP_TRY(arp_bind(c));
    if (!c->packed_valid) ARP_TRY(arp_pairs_sorted_build(c, 1));          /* the compact and the packed view share the distance stream */
    ARP_CUDA(c, cudaMemcpyAsync(dist, c->sort_d.p, (size_t)c->n_pairs * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    return ARP_OK;
}

/* host only: the compact view back into 16-byte records (dist == NULL: distance 0) */
int arp_pairs_unpack(const uint32_t* row_off, const arp_pair_c* rec, const float* dist, int32_t n_atoms, arp_pair* dst, uint64_t cap)
{
    if (n_atoms < 0 || (n_atoms > 0 && !row_off)) return ARP_E_INVALID_ARG;
    if (n_atoms == 0) return ARP_OK;
    const uint64_t n = row_off[n_atoms];
    if (n > cap) return ARP_E_CAPACITY;
    if (n && (!rec || !dst)) return ARP_E_INVALID_ARG;
    for (int32_t i = 0; i < n_atoms; ++i) {
        if (row_off[i + 1] < row_off[i] || row_off[i + 1] > n) return ARP_E_INVALID_ARG;
        for (uint64_t k = row_off[i]; k < row_off[i + 1]; ++k) {
            dst[k].i = i; dst[k].j = rec[k].j; dst[k].mask = rec[k].mask; dst[k].dist = dist ? dist[k] : 0.f;
        }
    }
    return ARP_OK;
}

int arp_pairs_device_ptr(arp_ctx* c, const arp_pair** dptr)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, dptr != nullptr, ARP_E_INVALID_ARG, "dptr is NULL");
    ARP_TRY(pairs_ready(c, "no record stream yet"));
    *dptr = c->out.as<arp_pair>();
    return ARP_OK;
}

int arp_ring_nearest_atom(arp_ctx* c, const float* xyz, int32_t n_atoms, const double* centers, int32_t n_rings,
                          double radius, int32_t* atom_out, double* dist_out)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, n_atoms >= 0 && n_rings >= 0 && radius >= 0.0, ARP_E_INVALID_ARG, "negative size or radius");
    if (n_rings == 0) return ARP_OK;
    ARP_REQUIRE(c, centers && atom_out && dist_out && (xyz || n_atoms == 0), ARP_E_INVALID_ARG, "NULL array");
    ARP_REQUIRE(c, c->have_params, ARP_E_NOT_READY, "arp_ring_nearest_atom before arp_set_params");
    ARP_TRY(arp_bind(c));
    return arp_ring_nearest_run(c, xyz, n_atoms, centers, n_rings, radius, atom_out, dist_out);
}

int arp_atom_sifts_run(arp_ctx* c)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_TRY(pairs_ready(c, "arp_atom_sifts_run before arp_pairs_run"));
    ARP_TRY(arp_bind(c));
    ARP_TRY(arp_atom_sifts_enqueue(c));
    c->sifts_valid = 1;
    return ARP_OK;
}

int arp_atom_sifts_fetch(arp_ctx* c, arp_atom_sift* dst, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->sifts_valid && c->pairs_valid, ARP_E_NOT_READY, "arp_atom_sifts_fetch before arp_atom_sifts_run");
    ARP_REQUIRE(c, cap >= (uint64_t)c->N, ARP_E_CAPACITY, "destination holds fewer entries than there are atoms");
    if (c->N == 0) return ARP_OK;
    ARP_REQUIRE(c, dst != nullptr, ARP_E_INVALID_ARG, "dst is NULL");
    ARP_TRY(arp_bind(c));
    ARP_CUDA(c, cudaMemcpyAsync(dst, c->sift_out.p, (size_t)c->N * sizeof(arp_atom_sift), cudaMemcpyDeviceToHost, c->stream));
    ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    return ARP_OK;
}

int arp_flag_within(arp_ctx* c, double radius, uint8_t* flags_out, uint64_t cap)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->have_atoms, ARP_E_NOT_READY, "arp_flag_within before arp_upload_atoms");
    ARP_REQUIRE(c, radius >= 0.0, ARP_E_INVALID_ARG, "negative radius");
    ARP_REQUIRE(c, cap >= (uint64_t)c->N, ARP_E_CAPACITY, "flags_out too small");
    if (c->N == 0) return ARP_OK;
    ARP_REQUIRE(c, flags_out != nullptr, ARP_E_INVALID_ARG, "flags_out is NULL");
    ARP_TRY(arp_bind(c));
    ARP_TRY(arp_flag_within_run(c, radius));
    const uint8_t* d = c->within.as<uint8_t>() + c->within_flags_off;
    ARP_CUDA(c, cudaMemcpyAsync(flags_out, d, (size_t)c->N, cudaMemcpyDeviceToHost, c->stream));
    ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    return ARP_OK;
}

int arp_sync(arp_ctx* c)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_TRY(arp_bind(c));
    ARP_CUDA(c, cudaStreamSynchronize(c->stream));
    return ARP_OK;
}

int arp_get_stats(arp_ctx* c, arp_stats* out)
{
    if (!c || !out) return ARP_E_INVALID_ARG;
    *out = c->stats;
    return ARP_OK;
}

uint64_t arp_launch_count(arp_ctx* c) { return c ? c->launches : 0; }

/*
 * Benchmark hook: repeats the whole atom-atom job (memset + grid build + pair kernel) `iters`
 * times on the resident inputs.  Every iteration is bracketed by CUDA events on the context's
 * stream; with flush_l2 a buffer larger than L2 is overwritten between iterations, outside the
 * brackets.  *ms_per_iter = mean whole-job time; arp_get_stats then reports the mean grid-build
 * and pair-kernel times of the same iterations.
 */
int arp_timing_iters(arp_ctx* c, int iters, int flush_l2, float* ms_per_iter)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, c->have_atoms, ARP_E_NOT_READY, "arp_timing_iters before arp_upload_atoms");
    ARP_REQUIRE(c, iters > 0, ARP_E_INVALID_ARG, "iters must be positive");
    ARP_TRY(arp_bind(c));
    if (c->run_pending) ARP_TRY(pairs_finish(c));
    if (!c->pairs_valid) ARP_TRY(arp_pairs_run(c, nullptr));      /* sizes the record buffer */
    const size_t flush_bytes = (size_t)384 << 20;
    if (flush_l2) ARP_TRY(dbuf_reserve(c, c->flush, flush_bytes));
    /* the job as the API runs it: events around the whole job only, kernels free to overlap */
    double tot = 0.0, grid = 0.0, search = 0.0, classify = 0.0, hscan = 0.0;
    for (int it = 0; it < iters; ++it) {
        if (flush_l2) ARP_CUDA(c, cudaMemsetAsync(c->flush.p, it & 0xff, flush_bytes, c->stream));
        ARP_TRY(arp_pairs_enqueue(c, 1));
        ARP_CUDA(c, cudaStreamSynchronize(c->stream));
        ARP_REQUIRE(c, c->h_meta->n_pairs == c->n_pairs, ARP_E_CUDA, "record count changed between iterations");
        float w = 0.f;
        ARP_CUDA(c, cudaEventElapsedTime(&w, c->ev[0], c->ev[3]));
        tot += w;
    }
    /* grid build | pair kernels: a bounded number of extra iterations with ONE event in between */
    const int split_iters = iters < 32 ? iters : 32;
    double pairs = 0.0, grid2 = 0.0;
    for (int it = 0; it < split_iters; ++it) {
        if (flush_l2) ARP_CUDA(c, cudaMemsetAsync(c->flush.p, it & 0xff, flush_bytes, c->stream));
        ARP_TRY(arp_pairs_enqueue(c, 2));
        ARP_CUDA(c, cudaStreamSynchronize(c->stream));
        float a = 0.f, b = 0.f;
        ARP_CUDA(c, cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
        ARP_CUDA(c, cudaEventElapsedTime(&b, c->ev[1], c->ev[3]));
        grid2 += a; pairs += b;
    }
    /* the full split (diagnostic): events between all kernels */
    for (int it = 0; it < split_iters; ++it) {
        if (flush_l2) ARP_CUDA(c, cudaMemsetAsync(c->flush.p, it & 0xff, flush_bytes, c->stream));
        ARP_TRY(arp_pairs_enqueue(c, 3));
        ARP_CUDA(c, cudaStreamSynchronize(c->stream));
        float a = 0.f, b = 0.f, d = 0.f, h = 0.f;
        ARP_CUDA(c, cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
        ARP_CUDA(c, cudaEventElapsedTime(&b, c->ev[1], c->ev[2]));
        ARP_CUDA(c, cudaEventElapsedTime(&d, c->ev[2], c->ev[3]));
        ARP_CUDA(c, cudaEventElapsedTime(&h, c->ev[4], c->ev[3]));
        grid += a; search += b; classify += d; hscan += h;
    }
    fill_stats(c, 0);
    c->stats.ms_grid = (float)(grid2 / split_iters);
    c->stats.ms_pairs = (float)(pairs / split_iters);
    (void)grid;
    c->stats.ms_search = (float)(search / split_iters);
    c->stats.ms_classify = (float)(classify / split_iters);
    c->stats.ms_hscan = (float)(hscan / split_iters);
    c->stats.ms_total = (float)(tot / iters);
    c->sorted_valid = 0; c->compact_valid = 0; c->packed_valid = 0; c->sort_tmp_valid = 0; c->sifts_valid = 0;
    if (ms_per_iter) *ms_per_iter = (float)(tot / iters);
    return ARP_OK;
}

int arp_memcpy_probe(arp_ctx* c, uint64_t h2d_bytes, uint64_t d2h_bytes, int iters, float* ms)
{
    if (!c) return ARP_E_INVALID_ARG;
    ARP_REQUIRE(c, ms != nullptr && iters > 0 && h2d_bytes > 0 && d2h_bytes > 0, ARP_E_INVALID_ARG, "bad probe arguments");
    ARP_TRY(arp_bind(c));
    void *hu = nullptr, *hd = nullptr, *du = nullptr, *dd = nullptr;
    cudaStream_t s2 = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaError_t e = cudaMallocHost(&hu, h2d_bytes);
    if (e == cudaSuccess) e = cudaMallocHost(&hd, d2h_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&du, h2d_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dd, d2h_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaEventCreate(&ev[k]);
    if (e == cudaSuccess) {
        memset(hu, 1, h2d_bytes);
        cudaMemsetAsync(dd, 2, d2h_bytes, c->stream);
        for (int w = 0; w < 2; ++w) {       /* one warm-up pass, one timed */
            cudaStreamSynchronize(c->stream); cudaStreamSynchronize(s2);
            cudaEventRecord(ev[0], c->stream); cudaEventRecord(ev[2], s2);
            for (int it = 0; it < iters; ++it) {
                cudaMemcpyAsync(du, hu, h2d_bytes, cudaMemcpyHostToDevice, c->stream);
                cudaMemcpyAsync(hd, dd, d2h_bytes, cudaMemcpyDeviceToHost, s2);
            }
            cudaEventRecord(ev[1], c->stream); cudaEventRecord(ev[3], s2);
        }
        cudaStreamSynchronize(c->stream);
        e = cudaStreamSynchronize(s2);
        if (e == cudaSuccess) { cudaEventElapsedTime(&ms[0], ev[0], ev[1]); cudaEventElapsedTime(&ms[1], ev[2], ev[3]); }
    }
    for (int k = 0; k < 4; ++k) if (ev[k]) cudaEventDestroy(ev[k]);
    if (s2) cudaStreamDestroy(s2);
    if (hu) cudaFreeHost(hu);
    if (hd) cudaFreeHost(hd);
    if (du) cudaFree(du);
    if (dd) cudaFree(dd);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return arp_fail(c, ARP_E_CUDA, cudaGetErrorString(e), __FILE__, __LINE__); }
    return ARP_OK;
}

}  /* extern "C" */
