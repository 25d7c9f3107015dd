/*
 * arp_rules.cuh -- the per-pair CREDO rule evaluation, written once for the device.
 *
 * Everything here is a pure function of the two atoms' records and of the side arrays
 * (hydrogens, bonds, halogen neighbours) addressed by ORIGINAL atom index.  The code is
 * __host__ __device__ so that tests/emu can also compile it for the host and check the rule
 * logic in the GPU-less build container; the product only ever runs it inside the CUDA
 * kernels of arp_pairs.cu.
 *
 * Arithmetic contract (measured against NumPy 2.3 / OpenBLAS):
 *   - float32 distance (interactions.py:745): component differences rounded to float32,
 *     squares rounded to float32, summed in double, rounded to float32, sqrtf.
 *   - float64 norm/dot of 3-vectors through BLAS (utils.py:87, :147): FMA chain when
 *     params.blas_fma, else sequential.
 *   - utils.get_angle (utils.py:696-745): plain scalar arithmetic, one rounding per op,
 *     in the dtype NumPy promotes to.  arccos is never evaluated: the angle tests are
 *     comparisons of the cosine with the host-computed images in arp_params (cos_*).
 * The translation unit must be compiled with -fmad=false (no implicit contraction);
 * every fused multiply-add below is explicit.
 *
 * Speed without changing a bit: the expensive exactly-rounded chains (float64 sqrt and six
 * divisions per hydrogen) sit behind cheap certainly-true / certainly-false screens
 *   - squared hydrogen distance against lim^2 (1 +- 4 ulp) before the float64 sqrt,
 *   - a float32 estimate of the cosine against the threshold +- 2e-5 before get_angle's chain,
 * and only the undecided sliver (and every NaN-prone degenerate geometry) runs the exact code,
 * which is kept out of line so that the hot loop stays small.
 */
#ifndef ARP_RULES_CUH
#define ARP_RULES_CUH

#include <stdint.h>
#include <string.h>
#include <math.h>

#include "../../include/arpeggio_cuda.h"

#if defined(__CUDACC__)
#define ARP_HD __host__ __device__ __forceinline__
#define ARP_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define ARP_HD static inline
#define ARP_HD_NOINLINE static __attribute__((noinline))
struct float4 { float x, y, z, w; };      /* host build of the rules (tests/emu) */
#endif

/* ---- packed per-atom word (cell-sorted attribute record, word 0) ---------------------
 * Internal layout, chosen so that every "p on one atom, q on the other" rule is one AND of the
 * bgn word with the pair-swapped end word (arp_pack_word builds it from the uploaded arrays):
 *   bits 0..8   radius class (K <= 512)
 *   bit  9      element H                      bits 10..11 ARP_R_* of the atom's residue
 *   bit  12     atom has bond CSR entries      bit  13 water      bit 14 in selection
 *   bit  15     aromatic                       bit  16 hydrophobe (bit 17 unused)
 *   pairs (even bit p, odd bit q):
 *   18 hbond acceptor   | 19 hbond donor                     A: is_hbond directions
 *   20 hbond acceptor   | 21 weak hbond donor                B: is_weak_hbond directions
 *   22 xbond acceptor   | 23 xbond donor                     C: is_xbond directions
 *   24 pos ionisable    | 25 neg ionisable                   D: ionic
 *   26 carbonyl oxygen  | 27 carbonyl carbon                 E: carbonyl
 *   28 hbond acceptor   | 29 is metal                        F: metal complex
 *   30 weak hbond acceptor AND halogen | 31 hbond donor OR weak hbond donor   G: halogen weak hbond  */
#define ARPK_RAD_MASK    0x1FFu
#define ARPK_MAX_RAD     512
#define ARPK_ELEM_H      (1u << 9)
#define ARPK_RES_SHIFT   10
#define ARPK_HAS_BOND    (1u << 12)
#define ARPK_WATER       (1u << 13)
#define ARPK_SELECTION   (1u << 14)
#define ARPK_AROMATIC    (1u << 15)
#define ARPK_HYDROPHOBE  (1u << 16)
#define ARPK_PAIR_EVEN   0x55540000u     /* bits 18, 20, ..., 30 */
#define ARPK_PAIR_ODD    0xAAA80000u     /* bits 19, 21, ..., 31 */
#define ARPK_ACC         (1u << 18)
#define ARPK_DON         (1u << 19)

#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
static inline
#endif
uint32_t arp_pack_word(uint32_t feat, uint32_t res_flags, uint32_t rad_class, int has_bond)
{
    uint32_t w = (rad_class & ARPK_RAD_MASK) | ((res_flags & 3u) << ARPK_RES_SHIFT);
    const uint32_t acc = (feat & ARP_F_HBOND_ACCEPTOR) != 0, don = (feat & ARP_F_HBOND_DONOR) != 0;
    const uint32_t wdon = (feat & ARP_F_WEAK_HBOND_DONOR) != 0;
    if (feat & ARP_F_ELEM_H) w |= ARPK_ELEM_H;
    if (has_bond) w |= ARPK_HAS_BOND;
    if (feat & ARP_F_IS_WATER) w |= ARPK_WATER;
    if (feat & ARP_F_IN_SELECTION) w |= ARPK_SELECTION;
    if (feat & ARP_F_AROMATIC) w |= ARPK_AROMATIC;
    if (feat & ARP_F_HYDROPHOBE) w |= ARPK_HYDROPHOBE;
    w |= acc << 18 | don << 19 | acc << 20 | wdon << 21 | acc << 28;
    if (feat & ARP_F_XBOND_ACCEPTOR) w |= 1u << 22;
    if (feat & ARP_F_XBOND_DONOR) w |= 1u << 23;
    if (feat & ARP_F_POS_IONISABLE) w |= 1u << 24;
    if (feat & ARP_F_NEG_IONISABLE) w |= 1u << 25;
    if (feat & ARP_F_CARBONYL_OXYGEN) w |= 1u << 26;
    if (feat & ARP_F_CARBONYL_CARBON) w |= 1u << 27;
    if (feat & ARP_F_IS_METAL) w |= 1u << 29;
    if ((feat & ARP_F_WEAK_HBOND_ACCEPTOR) && (feat & ARP_F_IS_HALOGEN)) w |= 1u << 30;
    w |= (don | wdon) << 31;
    return w;
}

struct ArpSide {               /* side arrays, original atom order (device pointers) */
    const double*   vdw;       /* [K] */
    const double*   cov;       /* [K] */
    const float4*   radtab;    /* [K][K] (f32(cov_a+cov_b), f32(vdw_a+vdw_b), f32((vdw_a+vdw_b)+comp), -) */
    int             K;
    const uint32_t* feat;      /* [N] ARP_F_* as uploaded (rare predicates only) */
    const int32_t*  bond_off;  /* [N+1] or null */
    const int32_t*  bond_nbr;
    const int32_t*  h_off;     /* [N+1] or null */
    const double*   h_xyz;     /* [H][3] */
    const float*    xnbr;      /* [N][3] or null */
    const float*    hlim;      /* [K] or null: no hydrogen of ANY atom lies within utils.py:89's limit of an acceptor of
                                  radius class k once the donor is farther than hlim[k] (triangle inequality over the
                                  longest donor-hydrogen distance of the upload, rounded up) */
};

struct ArpRuleParams {         /* arp_params narrowed the way NumPy (NEP 50) narrows it */
    double r2;                 /* interacting_cutoff^2, double (Bio.PDB.kdtrees) */
    double vdw_comp;
    double h_vdw;
    float  dist_max, hbond_polar, weak_polar, ionic, carbonyl, aromatic, hydrophobic, metal;
    double cos_hbond, cos_weak_hbond, cos_cx_min, cos_cx_max;
    float  cos_xbond_f32;
    int    blas_fma;
    int    include_seq_adjacent;
    /* what the comparisons give when get_angle() falls back to np.pi (utils.py:741-743) */
    int    pi_ge_hbond, pi_ge_weak_hbond, pi_in_cx, pi_ge_xbond;
};

/* arp_params -> ArpRuleParams (host) */
static inline void arp_derive_rule_params(const arp_params* p, ArpRuleParams* r)
{
    const double pi = 3.141592653589793;      /* np.pi */
    memset(r, 0, sizeof *r);
    r->r2 = p->interacting_cutoff * p->interacting_cutoff;
    r->vdw_comp = p->vdw_comp;
    r->h_vdw = p->h_vdw;
    r->dist_max = (float)p->dist_max;         /* NEP 50: np.float32 <op> python float compares in float32 */
    r->hbond_polar = (float)p->hbond_polar_dist;
    r->weak_polar = (float)p->weak_polar_dist;
    r->ionic = (float)p->ionic_dist;
    r->carbonyl = (float)p->carbonyl_dist;
    r->aromatic = (float)p->aromatic_dist;
    r->hydrophobic = (float)p->hydrophobic_dist;
    r->metal = (float)p->metal_dist;
    r->cos_hbond = p->cos_hbond;
    r->cos_weak_hbond = p->cos_weak_hbond;
    r->cos_cx_min = p->cos_cx_min;
    r->cos_cx_max = p->cos_cx_max;
    r->cos_xbond_f32 = p->cos_xbond_f32;
    r->blas_fma = p->blas_fma;
    r->include_seq_adjacent = p->include_sequence_adjacent;
    r->pi_ge_hbond = pi >= p->hbond_angle;
    r->pi_ge_weak_hbond = pi >= p->weak_hbond_angle;
    r->pi_in_cx = p->cx_angle_min <= pi && pi <= p->cx_angle_max;
    r->pi_ge_xbond = pi >= p->xbond_angle;
}

/* ---- exactly rounded primitives ------------------------------------------------------ */
ARP_HD float f_sub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
ARP_HD float f_add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
ARP_HD float f_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
ARP_HD float f_div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
ARP_HD float f_sqrt(float a) {
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(a);
#else
    return sqrtf(a);
#endif
}
ARP_HD double d_sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    volatile double r = a - b; return r;
#endif
}
ARP_HD double d_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
ARP_HD double d_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
ARP_HD double d_div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b; return r;
#endif
}
ARP_HD double d_sqrt(double a) {
#ifdef __CUDA_ARCH__
    return __dsqrt_rn(a);
#else
    return sqrt(a);
#endif
}
ARP_HD double d_fma(double a, double b, double c) { return fma(a, b, c); }
ARP_HD float fast_rsqrt(float a) {
#ifdef __CUDA_ARCH__
    return rsqrtf(a);
#else
    return 1.0f / sqrtf(a);
#endif
}

/* ---- NumPy / OpenBLAS models ------------------------------------------------------------ */

/* np.dot of float64 3-vectors (OpenBLAS ddot tail loop) */
ARP_HD double np_dot3_f64(double x0, double x1, double x2, double y0, double y1, double y2, int blas_fma)
{
    double d = d_mul(x0, y0);
    if (blas_fma) { d = d_fma(x1, y1, d); d = d_fma(x2, y2, d); }
    else          { d = d_add(d, d_mul(x1, y1)); d = d_add(d, d_mul(x2, y2)); }
    return d;
}
ARP_HD double np_norm3_f64(double x, double y, double z, int blas_fma)
{
    return d_sqrt(np_dot3_f64(x, y, z, x, y, z, blas_fma));
}
/* np.dot of float32 3-vectors (OpenBLAS sdot: float products, double accumulator) */
ARP_HD float np_dot3_f32(float x0, float x1, float x2, float y0, float y1, float y2)
{
    double d = (double)f_mul(x0, y0);
    d = d_add(d, (double)f_mul(x1, y1));
    d = d_add(d, (double)f_mul(x2, y2));
    return (float)d;
}
ARP_HD float np_norm3_f32(float x, float y, float z) { return f_sqrt(np_dot3_f32(x, y, z, x, y, z)); }

/* Bio.PDB.kdtrees radius test: double coordinates, sequential sum of squares <= r*r */
ARP_HD bool kd_within(float ax, float ay, float az, float bx, float by, float bz, double r2)
{
    double dx = d_sub((double)ax, (double)bx);
    double dy = d_sub((double)ay, (double)by);
    double dz = d_sub((double)az, (double)bz);
    double s = d_mul(dx, dx);
    s = d_add(s, d_mul(dy, dy));
    s = d_add(s, d_mul(dz, dz));
    return s <= r2;
}

/* np.linalg.norm(bgn.coord - end.coord), float32 (interactions.py:745) */
ARP_HD float np_dist_f32(float ax, float ay, float az, float bx, float by, float bz)
{
    return np_norm3_f32(f_sub(ax, bx), f_sub(ay, by), f_sub(az, bz));
}

/* arccos(c) is NaN (-> np.pi, utils.py:741-743) exactly when c is NaN or outside [-1, 1] */
ARP_HD bool acos_is_nan(double c) { return !(c >= -1.0 && c <= 1.0); }
ARP_HD bool acos_is_nan_f(float c) { return !(c >= -1.0f && c <= 1.0f); }

/* ---- cosines of utils.get_angle in its three dtype flows (exact chains, out of line) ---- */

/* v1 = a - b, v2 = c - b already formed in float64 (a, c float32, b = hydrogen float64):
   float64 throughout (utils.py:90, :113) */
ARP_HD_NOINLINE double cos_angle_dd_exact(double v1x, double v1y, double v1z, double v2x, double v2y, double v2z)
{
    double m1 = d_sqrt(d_add(d_add(d_mul(v1x, v1x), d_mul(v1y, v1y)), d_mul(v1z, v1z)));
    double m2 = d_sqrt(d_add(d_add(d_mul(v2x, v2x), d_mul(v2y, v2y)), d_mul(v2z, v2z)));
    double n1x = d_div(v1x, m1), n1y = d_div(v1y, m1), n1z = d_div(v1z, m1);
    double n2x = d_div(v2x, m2), n2y = d_div(v2y, m2), n2z = d_div(v2z, m2);
    return d_add(d_add(d_mul(n1x, n2x), d_mul(n1y, n2y)), d_mul(n1z, n2z));
}

/* a, b float32 (v1 stays float32), c float64 (v2 float64): utils.py:151 */
ARP_HD_NOINLINE double cos_angle_ffd(const float* a, const float* b, const double* c)
{
    float v1x = f_sub(a[0], b[0]), v1y = f_sub(a[1], b[1]), v1z = f_sub(a[2], b[2]);
    double v2x = d_sub(c[0], (double)b[0]), v2y = d_sub(c[1], (double)b[1]), v2z = d_sub(c[2], (double)b[2]);
    float m1 = f_sqrt(f_add(f_add(f_mul(v1x, v1x), f_mul(v1y, v1y)), f_mul(v1z, v1z)));
    float n1x = f_div(v1x, m1), n1y = f_div(v1y, m1), n1z = f_div(v1z, m1);
    double m2 = d_sqrt(d_add(d_add(d_mul(v2x, v2x), d_mul(v2y, v2y)), d_mul(v2z, v2z)));
    double n2x = d_div(v2x, m2), n2y = d_div(v2y, m2), n2z = d_div(v2z, m2);
    return d_add(d_add(d_mul((double)n1x, n2x), d_mul((double)n1y, n2y)), d_mul((double)n1z, n2z));
}

/* all float32 (utils.py:174) */
ARP_HD_NOINLINE float cos_angle_fff(const float* a, float bx, float by, float bz, float cx, float cy, float cz)
{
    float v1x = f_sub(a[0], bx), v1y = f_sub(a[1], by), v1z = f_sub(a[2], bz);
    float v2x = f_sub(cx, bx), v2y = f_sub(cy, by), v2z = f_sub(cz, bz);
    float m1 = f_sqrt(f_add(f_add(f_mul(v1x, v1x), f_mul(v1y, v1y)), f_mul(v1z, v1z)));
    float m2 = f_sqrt(f_add(f_add(f_mul(v2x, v2x), f_mul(v2y, v2y)), f_mul(v2z, v2z)));
    float n1x = f_div(v1x, m1), n1y = f_div(v1y, m1), n1z = f_div(v1z, m1);
    float n2x = f_div(v2x, m2), n2y = f_div(v2y, m2), n2z = f_div(v2z, m2);
    return f_add(f_add(f_mul(n1x, n2x), f_mul(n1y, n2y)), f_mul(n1z, n2z));
}

/* ---- predicates ---------------------------------------------------------------------------- */

#define ARP_HB_NEED_H 1        /* utils.is_hbond      threshold (utils.py:73-93)  */
#define ARP_HB_NEED_W 2        /* utils.is_weak_hbond threshold (utils.py:96-116) */

/*
 * One pass over the donor's hydrogens for utils.is_hbond and/or utils.is_weak_hbond: the two
 * differ only in the angle threshold, so hydrogen distance and cosine are shared.
 * Returns the subset of `need` whose predicate is true.
 */
ARP_HD int rule_hbond_scan_range(const ArpSide& S, const ArpRuleParams& P, int h0, int h1, float dcx, float dcy, float dcz,
                                 float acx, float acy, float acz, double vdw_acc, int need)
{
    if (h0 == h1) return 0;
    const double lim = d_add(d_add(P.h_vdw, vdw_acc), P.vdw_comp);                  /* utils.py:89 */
    const double lim2 = lim * lim;
    const double lim2_lo = lim2 * (1.0 - 1e-15), lim2_hi = lim2 * (1.0 + 1e-15);
    int got = 0;
#pragma unroll 1
    for (int k = h0; k < h1; ++k) {
        const double hx = S.h_xyz[3 * (size_t)k], hy = S.h_xyz[3 * (size_t)k + 1], hz = S.h_xyz[3 * (size_t)k + 2];
        /* h_dist = np.linalg.norm(h_coord - acceptor.coord) (utils.py:87) */
        const double ux = d_sub(hx, (double)acx), uy = d_sub(hy, (double)acy), uz = d_sub(hz, (double)acz);
        const double s = np_dot3_f64(ux, uy, uz, ux, uy, uz, P.blas_fma);
        bool within;
        if (s < lim2_lo) within = true;
        else if (s > lim2_hi) within = false;
        else within = d_sqrt(s) <= lim;                     /* also takes NaN */
        if (!within) continue;
        /* get_angle(donor.coord, h_coord, acceptor.coord) (utils.py:90): v1 = donor - h, v2 = acceptor - h */
        const double v1x = d_sub((double)dcx, hx), v1y = d_sub((double)dcy, hy), v1z = d_sub((double)dcz, hz);
        const double v2x = d_sub((double)acx, hx), v2y = d_sub((double)acy, hy), v2z = d_sub((double)acz, hz);
        /* float32 estimate of the cosine; |estimate - exact chain| < 2e-6 for non-degenerate vectors */
        const float f1x = (float)v1x, f1y = (float)v1y, f1z = (float)v1z;
        const float f2x = (float)v2x, f2y = (float)v2y, f2z = (float)v2z;
        const float q1 = f1x * f1x + f1y * f1y + f1z * f1z;
        const float q2 = (float)s;                         /* v2 = -(h - acceptor) */
        const float dt = f1x * f2x + f1y * f2y + f1z * f2z;
        const float ce = dt * fast_rsqrt(q1 * q2);
        const float tol = 2e-5f;
        int sure_true = 0, sure_false = 0;
        if (q1 > 1e-12f && q2 > 1e-12f && q1 < 1e12f && q2 < 1e12f && ce > -0.9999f && ce < 0.9999f) {
            if (need & ARP_HB_NEED_H) {
                if (ce < (float)P.cos_hbond - tol) sure_true |= ARP_HB_NEED_H;
                else if (ce > (float)P.cos_hbond + tol) sure_false |= ARP_HB_NEED_H;
            }
            if (need & ARP_HB_NEED_W) {
                if (ce < (float)P.cos_weak_hbond - tol) sure_true |= ARP_HB_NEED_W;
                else if (ce > (float)P.cos_weak_hbond + tol) sure_false |= ARP_HB_NEED_W;
            }
        }
        got |= sure_true;
        if ((need & ~got & ~sure_false) != 0) {              /* something still undecided for this hydrogen */
            const double c = cos_angle_dd_exact(v1x, v1y, v1z, v2x, v2y, v2z);
            const bool nan = acos_is_nan(c);
            if ((need & ARP_HB_NEED_H) && (nan ? P.pi_ge_hbond : (c <= P.cos_hbond))) got |= ARP_HB_NEED_H;
            if ((need & ARP_HB_NEED_W) && (nan ? P.pi_ge_weak_hbond : (c <= P.cos_weak_hbond))) got |= ARP_HB_NEED_W;
        }
        if ((need & ~got) == 0) break;                       /* the reference returns at the first success */
    }
    return got & need;
}

/* the same with the donor's hydrogens looked up in the CSR (h_off of the ORIGINAL atom index) */
ARP_HD int rule_hbond_scan(const ArpSide& S, const ArpRuleParams& P, int donor, float dcx, float dcy, float dcz,
                           float acx, float acy, float acz, double vdw_acc, int need)
{
    if (!S.h_off) return 0;
    return rule_hbond_scan_range(S, P, S.h_off[donor], S.h_off[donor + 1], dcx, dcy, dcz, acx, acy, acz, vdw_acc, need);
}

/* utils.is_halogen_weak_hbond (utils.py:119-155), screened like rule_hbond_scan */
ARP_HD int rule_is_halogen_weak_hbond(const ArpSide& S, const ArpRuleParams& P, int donor, int halogen,
                                      float hcx, float hcy, float hcz, uint32_t feat_hal, double vdw_hal)
{
    if (!(feat_hal & ARP_F_HAS_XNBR) || !S.xnbr || !S.h_off) return 0;            /* utils.py:139-141 */
    const int h0 = S.h_off[donor], h1 = S.h_off[donor + 1];
    if (h0 == h1) return 0;
    const float* nb = S.xnbr + 3 * (size_t)halogen;
    const float hc[3] = { hcx, hcy, hcz };
    const double lim = d_add(d_add(P.h_vdw, vdw_hal), P.vdw_comp);                 /* utils.py:149 */
    const double lim2 = lim * lim;
    const double lim2_lo = lim2 * (1.0 - 1e-15), lim2_hi = lim2 * (1.0 + 1e-15);
    const float w1x = nb[0] - hcx, w1y = nb[1] - hcy, w1z = nb[2] - hcz;           /* estimate of neighbour - halogen */
    const float q1 = w1x * w1x + w1y * w1y + w1z * w1z;
#pragma unroll 1
    for (int k = h0; k < h1; ++k) {
        const double* h = S.h_xyz + 3 * (size_t)k;
        const double vx = d_sub((double)hcx, h[0]), vy = d_sub((double)hcy, h[1]), vz = d_sub((double)hcz, h[2]);
        const double s = np_dot3_f64(vx, vy, vz, vx, vy, vz, P.blas_fma);          /* h_dist^2, utils.py:147 */
        bool within;
        if (s < lim2_lo) within = true;
        else if (s > lim2_hi) within = false;
        else within = d_sqrt(s) <= lim;
        if (!within) continue;
        /* get_angle(neighbour, halogen, h): v2 = h - halogen = -(vx, vy, vz) */
        const float q2 = (float)s;
        const float dt = -(w1x * (float)vx + w1y * (float)vy + w1z * (float)vz);
        const float ce = dt * fast_rsqrt(q1 * q2);
        const float tol = 2e-5f;
        if (q1 > 1e-12f && q2 > 1e-12f && q1 < 1e12f && q2 < 1e12f && ce > -0.9999f && ce < 0.9999f) {
            if (ce < (float)P.cos_cx_min - tol && ce > (float)P.cos_cx_max + tol) return 1;
            if (ce > (float)P.cos_cx_min + tol || ce < (float)P.cos_cx_max - tol) continue;
        }
        const double c = cos_angle_ffd(nb, hc, h);
        if (acos_is_nan(c) ? P.pi_in_cx : (c <= P.cos_cx_min && c >= P.cos_cx_max)) return 1;  /* utils.py:151 */
    }
    return 0;
}

#define ARPK_FAULT_XBOND_NO_NBR (1u << 31)

/* utils.is_xbond (utils.py:158-179); a donor without single-bond neighbour makes the reference
   raise (utils.py:173): reported through the fault bit, like the oracle */
ARP_HD_NOINLINE int rule_is_xbond(const ArpSide& S, const ArpRuleParams& P, int donor, float dcx, float dcy, float dcz,
                                  float acx, float acy, float acz, uint32_t feat_donor, uint32_t* fault)
{
    if (!(feat_donor & ARP_F_HAS_XNBR) || !S.xnbr) { *fault |= ARPK_FAULT_XBOND_NO_NBR; return 0; }
    float c = cos_angle_fff(S.xnbr + 3 * (size_t)donor, dcx, dcy, dcz, acx, acy, acz);
    if (acos_is_nan_f(c)) return P.pi_ge_xbond;
    return c <= P.cos_xbond_f32;
}

/* InteractionComplex.__get_contact_type (interactions.py:643-691): six ifs, last true wins */
ARP_HD uint32_t rule_entity_class_bools(bool sb, bool se, bool wb, bool we)
{
    uint32_t c = 7;
    if (!sb && !se) c = ARP_CLASS_INTRA_NON_SELECTION;
    if (sb && se) c = ARP_CLASS_INTRA_SELECTION;
    if (sb != se) c = ARP_CLASS_INTER;
    if ((sb && we) || (se && wb)) c = ARP_CLASS_SELECTION_WATER;
    if ((!sb && we) || (!se && wb)) c = ARP_CLASS_NON_SELECTION_WATER;
    if (wb && we) c = ARP_CLASS_WATER_WATER;
    return c;
}
/* the same as a 16 x 3-bit table indexed by sb | se << 1 | wb << 2 | we << 3 (built from the ifs above) */
ARP_HD uint32_t rule_entity_class(uint32_t wb_word, uint32_t we_word)
{
    unsigned long long lut = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k)
        lut |= (unsigned long long)rule_entity_class_bools(k & 1, (k >> 1) & 1, (k >> 2) & 1, (k >> 3) & 1) << (3 * k);
    const uint32_t idx = ((wb_word >> 14) & 1u) | (((we_word >> 14) & 1u) << 1) | (((wb_word >> 13) & 1u) << 2) |
                         (((we_word >> 13) & 1u) << 3);
    return (uint32_t)(lut >> (3 * idx)) & 7u;
}

/*
 * The filters that make _calculate_atom_contacts `continue` (interactions.py:712-741).
 * (fb, rb, pb, nb) belong to the atom with the LOWER list index (atom_bgn), e to atom_end.
 * f* are packed words (ARPK_*), r* residue index, p and n the residue's prev/next links.
 */
ARP_HD bool rule_pair_survives(uint32_t fb, int rb, int pb, int nb, uint32_t fe, int re, int pe, int ne,
                               int include_seq_adjacent)
{
    if ((fb | fe) & ARPK_ELEM_H) return false;                                     /* :712-713 */
    if (rb == re) return false;                                                    /* :729-730 */
    if (!include_seq_adjacent) {                                                   /* :733 */
        uint32_t flb = fb >> ARPK_RES_SHIFT, fle = fe >> ARPK_RES_SHIFT;
        if ((fle & ARP_R_IS_POLYPEPTIDE) &&                                        /* :734, res_end twice */
            (flb & ARP_R_HAS_LINKS) && (fle & ARP_R_HAS_LINKS) &&                  /* :736-737 */
            (nb == re || pb == re || ne == rb || pe == rb))                        /* :739-740 */
            return false;
    }
    return true;
}

/* ---- deferred work of one pair (rule_classify_core -> the dense second stage) --------------------- */
#define ARP_WORK_SCAN0  0x03u   /* bits 0..1: ARP_HB_NEED_* for direction 0 (donor = bgn, acceptor = end) */
#define ARP_WORK_SCAN1  0x0Cu   /* bits 2..3: ARP_HB_NEED_* for direction 1 (donor = end, acceptor = bgn) */
#define ARP_WORK_HAL0   0x10u   /* is_halogen_weak_hbond(donor = bgn, halogen = end) */
#define ARP_WORK_HAL1   0x20u   /* is_halogen_weak_hbond(donor = end, halogen = bgn) */
#define ARP_WORK_XB0    0x40u   /* is_xbond(donor = bgn, acceptor = end) */
#define ARP_WORK_XB1    0x80u   /* is_xbond(donor = end, acceptor = bgn) */
#define ARP_RARE_HAL    1u
#define ARP_RARE_XBOND  2u

/*
 * Loop body of _calculate_atom_contacts after the filters (interactions.py:743-936), without the
 * predicates that walk hydrogens or need an exact angle: those come back in *work and are OR-ed into
 * the mask by the caller (bits hbond, weak_hbond, xbond).
 * b = atom_bgn (lower list index), e = atom_end; coordinates and packed words are passed in registers.
 * X = wb & pairswap(we): bit 2k = "p on bgn, q on end", bit 2k + 1 = "q on bgn, p on end".
 */
/* bb0 / bb1: atom_bgn's range in the bond CSR when the caller has already fetched it (bb0 < 0: fetch it here).  The bond
   walk (two dependent loads, only for atoms that have bonds) comes LAST: everything that does not depend on `bonded` is
   computed first, so that the caller's early loads of the range fly under that arithmetic.  The reference gates the
   feature rules on `not clash` (interactions.py:786); here they are evaluated on the distance alone and dropped again
   for a clash, which is the same thing (they have no other effect). */
ARP_HD void rule_classify_core(const ArpSide& S, const ArpRuleParams& P, int b, int e,
                               float bx, float by, float bz, float ex, float ey, float ez,
                               uint32_t wb, uint32_t we, uint32_t* mask_out, float* dist_out, uint32_t* work_out,
                               int bb0 = -1, int bb1 = -1)
{
    const float4 t = S.radtab[(wb & ARPK_RAD_MASK) * (uint32_t)S.K + (we & ARPK_RAD_MASK)];   /* :717-718 narrowed */
    const float t_cov = t.x, t_vdw = t.y, vdwc = t.z;
    const float d = np_dist_f32(bx, by, bz, ex, ey, ez);                           /* :745 */

    const uint32_t sw = ((we & ARPK_PAIR_EVEN) << 1) | ((we & ARPK_PAIR_ODD) >> 1);
    const uint32_t X = wb & sw, Y = wb & we;
#define ARP_XB(n) ((X >> (n)) & 1u)
    uint32_t m = 0;
    if (d <= P.metal) m |= (ARP_XB(28) | ARP_XB(29)) << ARP_SIFT_METAL;            /* :777-783 */

    uint32_t mf = 0, work = 0;          /* the feature bits and deferred predicates of a pair that is not a clash */
    if (d <= P.dist_max) {                                                         /* :786 */
        const bool in_vdwc = d <= vdwc;
        /* hbond / polar :791-819 */
        const bool ws_b = (wb & ARPK_WATER) && in_vdwc;
        const bool ws_e = !ws_b && (we & ARPK_WATER) && in_vdwc;
        if (ws_b) {
            if (we & (ARPK_ACC | ARPK_DON)) mf |= (1u << ARP_SIFT_HBOND) | (1u << ARP_SIFT_POLAR);
        } else if (ws_e) {
            if (wb & (ARPK_ACC | ARPK_DON)) mf |= (1u << ARP_SIFT_HBOND) | (1u << ARP_SIFT_POLAR);
        } else if (X & (3u << 18)) {
            work |= ARP_XB(19) ? ARP_HB_NEED_H : (ARP_HB_NEED_H << 2);             /* is_hbond(bgn, end) before (end, bgn) */
            if (d <= P.hbond_polar) mf |= 1u << ARP_SIFT_POLAR;
        }
        /* weak hbond / weak polar: four independent ifs, each ASSIGNS SIFt[6] (:857-886), so only the
           last applicable one decides the bit; any applicable one enables weak polar */
        if (X & ((3u << 20) | (3u << 30))) {
            if (ARP_XB(31))      work |= ARP_WORK_HAL0;                            /* halogen end, donor bgn */
            else if (ARP_XB(30)) work |= ARP_WORK_HAL1;                            /* halogen bgn, donor end */
            else if (ARP_XB(21)) work |= ARP_HB_NEED_W;                            /* is_weak_hbond(bgn, end) */
            else                 work |= ARP_HB_NEED_W << 2;                       /* is_weak_hbond(end, bgn) */
            if (d <= P.weak_polar) mf |= 1u << ARP_SIFT_WEAK_POLAR;
        }
        /* xbond :889-895 */
        if (in_vdwc && (X & (3u << 22))) work |= ARP_XB(23) ? ARP_WORK_XB0 : ARP_WORK_XB1;
        /* ionic :898-904, carbonyl :907-913, aromatic :916-917, hydrophobic :920-921 */
        /* hydrogen scans that cannot succeed are not scheduled: |H - acceptor| >= d - |donor - H| > limit */
        if (S.hlim && (work & 0x3Fu)) {
            if (d > S.hlim[we & ARPK_RAD_MASK]) work &= ~(ARP_WORK_SCAN0 | ARP_WORK_HAL0);   /* acceptor / halogen = end */
            if (d > S.hlim[wb & ARPK_RAD_MASK]) work &= ~(ARP_WORK_SCAN1 | ARP_WORK_HAL1);   /* acceptor / halogen = bgn */
        }
        if (d <= P.ionic) mf |= (ARP_XB(24) | ARP_XB(25)) << ARP_SIFT_IONIC;
        if (d <= P.carbonyl) mf |= (ARP_XB(26) | ARP_XB(27)) << ARP_SIFT_CARBONYL;
        if (d <= P.aromatic) mf |= ((Y >> 15) & 1u) << ARP_SIFT_AROMATIC;
        if (d <= P.hydrophobic) mf |= ((Y >> 16) & 1u) << ARP_SIFT_HYDROPHOBIC;
    }
#undef ARP_XB
    m |= rule_entity_class(wb, we) << ARP_CLASS_SHIFT;

    bool bonded = false;                                                           /* :750-754 */
    if (wb & ARPK_HAS_BOND) {           /* only atom_bgn's neighbour list is consulted */
        if (bb0 < 0) { bb0 = S.bond_off[b]; bb1 = S.bond_off[b + 1]; }
        for (int k = bb0; k < bb1; ++k)
            if (S.bond_nbr[k] == e) { bonded = true; break; }
    }
    const bool clash = !bonded && d < t_cov;
    m |= bonded ? 1u << ARP_SIFT_COVALENT                                          /* :756-757 */
       : clash ? 1u << ARP_SIFT_CLASH                                              /* :760 */
       : d < t_vdw ? 1u << ARP_SIFT_VDW_CLASH                                      /* :764 */
       : d <= vdwc ? 1u << ARP_SIFT_VDW                                            /* :768 */
       : 1u << ARP_SIFT_PROXIMAL;                                                  /* :772 */
    if (clash) { mf = 0; work = 0; }                                               /* :786: the feature rules are skipped for a clash */
    *mask_out = m | mf;
    *dist_out = d;
    *work_out = work;
}

/* the whole loop body for one pair: core + its deferred work, evaluated in place
   (host emulation and single-pair uses; the kernels run the deferred work densely instead) */
ARP_HD void rule_classify(const ArpSide& S, const ArpRuleParams& P, int b, int e,
                          float bx, float by, float bz, float ex, float ey, float ez,
                          uint32_t wb, uint32_t we, uint32_t* mask_out, float* dist_out)
{
    uint32_t m, work;
    rule_classify_core(S, P, b, e, bx, by, bz, ex, ey, ez, wb, we, &m, dist_out, &work);
    const double vdw_b = S.vdw[wb & ARPK_RAD_MASK], vdw_e = S.vdw[we & ARPK_RAD_MASK];
    int got = 0, weak = 0;
    uint32_t fault = 0;
    if (work & ARP_WORK_SCAN0) got |= rule_hbond_scan(S, P, b, bx, by, bz, ex, ey, ez, vdw_e, (int)(work & 3u));
    if (work & ARP_WORK_SCAN1) got |= rule_hbond_scan(S, P, e, ex, ey, ez, bx, by, bz, vdw_b, (int)((work >> 2) & 3u));
    if (work & ARP_WORK_HAL0) weak = rule_is_halogen_weak_hbond(S, P, b, e, ex, ey, ez, S.feat[e], vdw_e);
    if (work & ARP_WORK_HAL1) weak = rule_is_halogen_weak_hbond(S, P, e, b, bx, by, bz, S.feat[b], vdw_b);
    if (got & ARP_HB_NEED_H) m |= 1u << ARP_SIFT_HBOND;
    if ((got & ARP_HB_NEED_W) || weak) m |= 1u << ARP_SIFT_WEAK_HBOND;
    if ((work & ARP_WORK_XB0) && rule_is_xbond(S, P, b, bx, by, bz, ex, ey, ez, S.feat[b], &fault)) m |= 1u << ARP_SIFT_XBOND;
    if ((work & ARP_WORK_XB1) && rule_is_xbond(S, P, e, ex, ey, ez, bx, by, bz, S.feat[e], &fault)) m |= 1u << ARP_SIFT_XBOND;
    *mask_out = m | fault;
}

#endif /* ARP_RULES_CUH */
