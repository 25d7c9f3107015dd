/*
 * arp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code)
 *
 * A plain-C restatement of the interatomic-contact hot path of pdbe-arpeggio
 * at the structure-of-arrays boundary declared in include/arpeggio_cuda.h.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product (arpeggio_b200/) never does.
 *
 * Every function cites the reference code it follows (paths relative to the
 * pdbe-arpeggio tree, arpeggio/core/...).  Third-party arithmetic that is not
 * in the reference tree is restated from its published behaviour:
 *   - Bio.PDB.NeighborSearch / Bio/PDB/kdtrees.c (BioPython >= 1.80, unpinned
 *     in setup.py:38): coordinates held as double, pair reported when
 *     dx*dx + dy*dy + dz*dz <= r*r (sequential double), index1 < index2.
 *   - NumPy 2.3.5 + OpenBLAS 0.3.30 (x86-64 Haswell/SkylakeX kernels):
 *     np.linalg.norm / np.dot of 3-vectors.  float32: products rounded to
 *     float32, accumulated in double, result rounded to float32.  float64:
 *     fma chain fma(x2,y2,fma(x1,y1,x0*y0)) (params.blas_fma=1) -- both
 *     measured against live NumPy, see tests/test_numpy_model.py.
 *   - NEP 50: np.float32 <op> python-float is evaluated in float32.
 *
 * Parity pin: the fixtures under tests/golden/ are produced by the reference's own
 * _calculate_atom_contacts/_calculate_ring_contacts/_calculate_group_contacts
 * (imported from /root/reference with stubbed Bio/openbabel/gemmi) and this
 * oracle is checked against them in tests/test_oracle_golden.py.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: no implicit FMA).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/arpeggio_cuda.h"

#define ORC_FAULT_XBOND_NO_NBR (1u << 31)

/* ------------------------------------------------------------------------ */
/* NumPy / OpenBLAS arithmetic models                                        */
/* ------------------------------------------------------------------------ */

/* np.dot(x, y) for float64 3-vectors (OpenBLAS ddot tail loop). */
static double dot3_f64(const double* x, const double* y, int blas_fma)
{
    if (blas_fma) {
        double d = x[0] * y[0];
        d = fma(x[1], y[1], d);
        d = fma(x[2], y[2], d);
        return d;
    }
    double d = x[0] * y[0];
    d = d + x[1] * y[1];
    d = d + x[2] * y[2];
    return d;
}

/* np.linalg.norm(v) for a float64 3-vector = sqrt(dot(v, v)). */
static double norm3_f64(const double* v, int blas_fma)
{
    return sqrt(dot3_f64(v, v, blas_fma));
}

/* np.dot(x, y) for float32 3-vectors (OpenBLAS sdot: float products, double accumulator). */
static float dot3_f32(const float* x, const float* y)
{
    volatile float p0 = x[0] * y[0];
    volatile float p1 = x[1] * y[1];
    volatile float p2 = x[2] * y[2];
    double d = 0.0;
    d += (double)p0;
    d += (double)p1;
    d += (double)p2;
    return (float)d;
}

static float norm3_f32(const float* v)
{
    return sqrtf(dot3_f32(v, v));
}

/* ------------------------------------------------------------------------ */
/* utils.get_angle (utils.py:696-745) in its three dtype flows.              */
/* Returns the angle in radians as a double (float32 results widened);      */
/* NaN -> pi (utils.py:741-743).                                             */
/* ------------------------------------------------------------------------ */

/* all three points float32 (is_xbond, utils.py:174): float32 throughout */
static double get_angle_fff(const float* a, const float* b, const float* c, int* is_f32_result, float* f32_out)
{
    float v1[3], v2[3];
    for (int k = 0; k < 3; ++k) { v1[k] = a[k] - b[k]; v2[k] = c[k] - b[k]; }
    volatile float s1 = v1[0] * v1[0]; volatile float t1 = v1[1] * v1[1]; volatile float u1 = v1[2] * v1[2];
    volatile float q1 = s1 + t1; q1 = q1 + u1;
    float m1 = sqrtf(q1);
    volatile float s2 = v2[0] * v2[0]; volatile float t2 = v2[1] * v2[1]; volatile float u2 = v2[2] * v2[2];
    volatile float q2 = s2 + t2; q2 = q2 + u2;
    float m2 = sqrtf(q2);
    volatile float n1x = v1[0] / m1, n1y = v1[1] / m1, n1z = v1[2] / m1;
    volatile float n2x = v2[0] / m2, n2y = v2[1] / m2, n2z = v2[2] / m2;
    volatile float r0 = n1x * n2x; volatile float r1 = n1y * n2y; volatile float r2 = n1z * n2z;
    volatile float res = r0 + r1; res = res + r2;
    float ang = acosf(res);
    if (isnan(ang)) { *is_f32_result = 0; return M_PI; }   /* angle = np.pi (python float) */
    *is_f32_result = 1; *f32_out = ang;
    return (double)ang;
}

/* v1 float32 (a, b float32), v2 float64 (c float64): is_halogen_weak_hbond, utils.py:151 */
static double get_angle_ffd(const float* a, const float* b, const double* c)
{
    float v1[3]; double v2[3];
    for (int k = 0; k < 3; ++k) { v1[k] = a[k] - b[k]; v2[k] = c[k] - (double)b[k]; }
    volatile float s1 = v1[0] * v1[0]; volatile float t1 = v1[1] * v1[1]; volatile float u1 = v1[2] * v1[2];
    volatile float q1 = s1 + t1; q1 = q1 + u1;
    float m1 = sqrtf(q1);
    volatile float n1x = v1[0] / m1, n1y = v1[1] / m1, n1z = v1[2] / m1;
    double q2 = v2[0] * v2[0] + v2[1] * v2[1]; q2 = q2 + v2[2] * v2[2];
    double m2 = sqrt(q2);
    double n2x = v2[0] / m2, n2y = v2[1] / m2, n2z = v2[2] / m2;
    double res = (double)n1x * n2x + (double)n1y * n2y; res = res + (double)n1z * n2z;
    double ang = acos(res);
    if (isnan(ang)) return M_PI;
    return ang;
}

/* a float32, b float64, c float32 (is_hbond / is_weak_hbond, utils.py:90,113): float64 throughout */
static double get_angle_fdf(const float* a, const double* b, const float* c)
{
    double v1[3], v2[3];
    for (int k = 0; k < 3; ++k) { v1[k] = (double)a[k] - b[k]; v2[k] = (double)c[k] - b[k]; }
    double q1 = v1[0] * v1[0] + v1[1] * v1[1]; q1 = q1 + v1[2] * v1[2];
    double m1 = sqrt(q1);
    double q2 = v2[0] * v2[0] + v2[1] * v2[1]; q2 = q2 + v2[2] * v2[2];
    double m2 = sqrt(q2);
    double n1x = v1[0] / m1, n1y = v1[1] / m1, n1z = v1[2] / m1;
    double n2x = v2[0] / m2, n2y = v2[1] / m2, n2z = v2[2] / m2;
    double res = n1x * n2x + n1y * n2y; res = res + n1z * n2z;
    double ang = acos(res);
    if (isnan(ang)) return M_PI;
    return ang;
}

/* ------------------------------------------------------------------------ */
/* pair predicates                                                           */
/* ------------------------------------------------------------------------ */

typedef struct {
    const arp_atoms* A;
    const arp_params* P;
} orc_env;

static int n_hyd(const arp_atoms* A, int i) { return A->h_off ? A->h_off[i + 1] - A->h_off[i] : 0; }
static const double* hyd(const arp_atoms* A, int i, int k) { return A->h_xyz + 3 * (size_t)(A->h_off[i] + k); }

/* utils.is_hbond (utils.py:73-93) and utils.is_weak_hbond (utils.py:96-116):
   identical but for the angle threshold */
static int orc_is_hbond(const orc_env* E, int donor, int acceptor, double angle_thr)
{
    const arp_atoms* A = E->A; const arp_params* P = E->P;
    const float* dc = A->xyz + 3 * (size_t)donor;
    const float* ac = A->xyz + 3 * (size_t)acceptor;
    double lim = P->h_vdw + A->vdw[A->rad_class[acceptor]] + P->vdw_comp;   /* utils.py:89 */
    for (int k = 0; k < n_hyd(A, donor); ++k) {
        const double* h = hyd(A, donor, k);
        double v[3] = { h[0] - (double)ac[0], h[1] - (double)ac[1], h[2] - (double)ac[2] };
        double h_dist = norm3_f64(v, P->blas_fma);                             /* utils.py:87 */
        if (h_dist <= lim) {
            if (get_angle_fdf(dc, h, ac) >= angle_thr) return 1;              /* utils.py:90 */
        }
    }
    return 0;
}

/* utils.is_halogen_weak_hbond (utils.py:119-155) */
static int orc_is_halogen_weak_hbond(const orc_env* E, int donor, int halogen)
{
    const arp_atoms* A = E->A; const arp_params* P = E->P;
    if (!(A->feat[halogen] & ARP_F_HAS_XNBR) || !A->xnbr_xyz) return 0;      /* utils.py:139-141 */
    const float* nb = A->xnbr_xyz + 3 * (size_t)halogen;
    const float* hc = A->xyz + 3 * (size_t)halogen;
    double lim = P->h_vdw + A->vdw[A->rad_class[halogen]] + P->vdw_comp;     /* utils.py:149 */
    for (int k = 0; k < n_hyd(A, donor); ++k) {
        const double* h = hyd(A, donor, k);
        double v[3] = { (double)hc[0] - h[0], (double)hc[1] - h[1], (double)hc[2] - h[2] };
        double h_dist = norm3_f64(v, P->blas_fma);                             /* utils.py:147 */
        if (h_dist <= lim) {
            double ang = get_angle_ffd(nb, hc, h);
            if (P->cx_angle_min <= ang && ang <= P->cx_angle_max) return 1;   /* utils.py:151 */
        }
    }
    return 0;
}

/* utils.is_xbond (utils.py:158-179).  The reference dereferences None when the
   donor has no single-bond neighbour (utils.py:173); that is reported as a fault. */
static int orc_is_xbond(const orc_env* E, int donor, int acceptor, uint32_t* fault)
{
    const arp_atoms* A = E->A; const arp_params* P = E->P;
    if (!(A->feat[donor] & ARP_F_HAS_XNBR) || !A->xnbr_xyz) { *fault |= ORC_FAULT_XBOND_NO_NBR; return 0; }
    int isf = 0; float tf = 0.f;
    double theta = get_angle_fff(A->xnbr_xyz + 3 * (size_t)donor, A->xyz + 3 * (size_t)donor,
                                 A->xyz + 3 * (size_t)acceptor, &isf, &tf);
    if (isf) return tf >= (float)P->xbond_angle;     /* np.float32 >= python float -> float32 compare */
    return theta >= P->xbond_angle;                  /* np.pi >= 2.09 */
}

/* InteractionComplex.__get_contact_type (interactions.py:643-691): six ifs, last true wins */
static uint32_t orc_entity_class(uint32_t fb, uint32_t fe)
{
    int sb = (fb & ARP_F_IN_SELECTION) != 0, se = (fe & ARP_F_IN_SELECTION) != 0;
    int wb = (fb & ARP_F_IS_WATER) != 0,     we = (fe & ARP_F_IS_WATER) != 0;
    uint32_t c = 7;
    if (!sb && !se) c = ARP_CLASS_INTRA_NON_SELECTION;
    if (sb && se) c = ARP_CLASS_INTRA_SELECTION;
    if ((sb && !se) || (se && !sb)) c = ARP_CLASS_INTER;
    if ((sb && we) || (se && wb)) c = ARP_CLASS_SELECTION_WATER;
    if ((!sb && we) || (!se && wb)) c = ARP_CLASS_NON_SELECTION_WATER;
    if (wb && we) c = ARP_CLASS_WATER_WATER;
    return c;
}

static int orc_bonded(const arp_atoms* A, int b, int e)
{
    if (!A->bond_off) return 0;
    for (int k = A->bond_off[b]; k < A->bond_off[b + 1]; ++k)       /* interactions.py:750-754 */
        if (A->bond_nbr[k] == e) return 1;
    return 0;
}

/* np.linalg.norm(atom_bgn.coord - atom_end.coord) (interactions.py:745), float32 */
static float orc_dist_f32(const float* a, const float* b)
{
    float v[3] = { a[0] - b[0], a[1] - b[1], a[2] - b[2] };
    return norm3_f32(v);
}

/*
 * One (bgn, end) pair of the loop body of _calculate_atom_contacts
 * (interactions.py:707-936).  Returns 0 when the reference `continue`s,
 * else 1 with the record filled.
 */
static int orc_classify_pair(const orc_env* E, int b, int e, arp_pair* out)
{
    const arp_atoms* A = E->A; const arp_params* P = E->P;
    uint32_t fb = A->feat[b], fe = A->feat[e];

    if ((fb & ARP_F_ELEM_H) || (fe & ARP_F_ELEM_H)) return 0;                 /* :712-713 */
    uint32_t cls = orc_entity_class(fb, fe);                                   /* :715 */

    double sum_cov = A->cov[A->rad_class[b]] + A->cov[A->rad_class[e]];        /* :717 */
    double sum_vdw = A->vdw[A->rad_class[b]] + A->vdw[A->rad_class[e]];        /* :718 */

    int rb = A->res_id[b], re = A->res_id[e];
    if (rb == re) return 0;                                                    /* :729-730 */

    if (!P->include_sequence_adjacent) {                                       /* :733 */
        if (A->res_flags[re] & ARP_R_IS_POLYPEPTIDE) {                         /* :734 (res_end twice) */
            if ((A->res_flags[rb] & ARP_R_HAS_LINKS) && (A->res_flags[re] & ARP_R_HAS_LINKS)) { /* :736-737 */
                if (A->res_next[rb] == re || A->res_prev[rb] == re ||
                    A->res_next[re] == rb || A->res_prev[re] == rb)            /* :739-740 */
                    return 0;
            }
        }
    }

    float d = orc_dist_f32(A->xyz + 3 * (size_t)b, A->xyz + 3 * (size_t)e);   /* :745 */
    float vdwc = (float)(sum_vdw + P->vdw_comp);
    uint32_t m = 0, fault = 0;

    if (orc_bonded(A, b, e))            m |= 1u << ARP_SIFT_COVALENT;          /* :756-757 */
    else if (d < (float)sum_cov)        m |= 1u << ARP_SIFT_CLASH;             /* :760 */
    else if (d < (float)sum_vdw)        m |= 1u << ARP_SIFT_VDW_CLASH;         /* :764 */
    else if (d <= vdwc)                 m |= 1u << ARP_SIFT_VDW;               /* :768 */
    else                                m |= 1u << ARP_SIFT_PROXIMAL;          /* :772 */

    if (d <= (float)P->metal_dist) {                                           /* :777 */
        if ((fb & ARP_F_HBOND_ACCEPTOR) && (fe & ARP_F_IS_METAL)) m |= 1u << ARP_SIFT_METAL;
        else if ((fe & ARP_F_HBOND_ACCEPTOR) && (fb & ARP_F_IS_METAL)) m |= 1u << ARP_SIFT_METAL;
    }

    if (!(m & (1u << ARP_SIFT_CLASH)) && d <= (float)P->dist_max) {            /* :786 */
        /* HBOND / POLAR  :791-819 */
        if ((fb & ARP_F_IS_WATER) && d <= vdwc) {
            if (fe & (ARP_F_HBOND_ACCEPTOR | ARP_F_HBOND_DONOR)) m |= (1u << ARP_SIFT_HBOND) | (1u << ARP_SIFT_POLAR);
        } else if ((fe & ARP_F_IS_WATER) && d <= vdwc) {
            if (fb & (ARP_F_HBOND_ACCEPTOR | ARP_F_HBOND_DONOR)) m |= (1u << ARP_SIFT_HBOND) | (1u << ARP_SIFT_POLAR);
        } else {
            if ((fb & ARP_F_HBOND_DONOR) && (fe & ARP_F_HBOND_ACCEPTOR)) {
                if (orc_is_hbond(E, b, e, P->hbond_angle)) m |= 1u << ARP_SIFT_HBOND;
                if (d <= (float)P->hbond_polar_dist) m |= 1u << ARP_SIFT_POLAR;
            } else if ((fe & ARP_F_HBOND_DONOR) && (fb & ARP_F_HBOND_ACCEPTOR)) {
                if (orc_is_hbond(E, e, b, P->hbond_angle)) m |= 1u << ARP_SIFT_HBOND;
                if (d <= (float)P->hbond_polar_dist) m |= 1u << ARP_SIFT_POLAR;
            }
        }
        /* WEAK HBOND / WEAK POLAR: four independent ifs, each assigns SIFt[6]  :857-886 */
        int weak = 0;
        if ((fb & ARP_F_HBOND_ACCEPTOR) && (fe & ARP_F_WEAK_HBOND_DONOR)) {
            weak = orc_is_hbond(E, e, b, P->weak_hbond_angle);
            if (d <= (float)P->weak_polar_dist) m |= 1u << ARP_SIFT_WEAK_POLAR;
        }
        if ((fb & ARP_F_WEAK_HBOND_DONOR) && (fe & ARP_F_HBOND_ACCEPTOR)) {
            weak = orc_is_hbond(E, b, e, P->weak_hbond_angle);
            if (d <= (float)P->weak_polar_dist) m |= 1u << ARP_SIFT_WEAK_POLAR;
        }
        if ((fb & ARP_F_WEAK_HBOND_ACCEPTOR) && (fb & ARP_F_IS_HALOGEN) &&
            (fe & (ARP_F_HBOND_DONOR | ARP_F_WEAK_HBOND_DONOR))) {
            weak = orc_is_halogen_weak_hbond(E, e, b);
            if (d <= (float)P->weak_polar_dist) m |= 1u << ARP_SIFT_WEAK_POLAR;
        }
        if ((fe & ARP_F_WEAK_HBOND_ACCEPTOR) && (fe & ARP_F_IS_HALOGEN) &&
            (fb & (ARP_F_HBOND_DONOR | ARP_F_WEAK_HBOND_DONOR))) {
            weak = orc_is_halogen_weak_hbond(E, b, e);
            if (d <= (float)P->weak_polar_dist) m |= 1u << ARP_SIFT_WEAK_POLAR;
        }
        if (weak) m |= 1u << ARP_SIFT_WEAK_HBOND;
        /* XBOND :889-895 */
        if (d <= vdwc) {
            if ((fb & ARP_F_XBOND_DONOR) && (fe & ARP_F_XBOND_ACCEPTOR)) {
                if (orc_is_xbond(E, b, e, &fault)) m |= 1u << ARP_SIFT_XBOND;
            } else if ((fe & ARP_F_XBOND_DONOR) && (fb & ARP_F_XBOND_ACCEPTOR)) {
                if (orc_is_xbond(E, e, b, &fault)) m |= 1u << ARP_SIFT_XBOND;
            }
        }
        /* IONIC :898-904 */
        if (d <= (float)P->ionic_dist) {
            if ((fb & ARP_F_POS_IONISABLE) && (fe & ARP_F_NEG_IONISABLE)) m |= 1u << ARP_SIFT_IONIC;
            else if ((fb & ARP_F_NEG_IONISABLE) && (fe & ARP_F_POS_IONISABLE)) m |= 1u << ARP_SIFT_IONIC;
        }
        /* CARBONYL :907-913 */
        if (d <= (float)P->carbonyl_dist) {
            if ((fb & ARP_F_CARBONYL_OXYGEN) && (fe & ARP_F_CARBONYL_CARBON)) m |= 1u << ARP_SIFT_CARBONYL;
            else if ((fe & ARP_F_CARBONYL_OXYGEN) && (fb & ARP_F_CARBONYL_CARBON)) m |= 1u << ARP_SIFT_CARBONYL;
        }
        /* AROMATIC :916-917, HYDROPHOBIC :920-921 */
        if ((fb & ARP_F_AROMATIC) && (fe & ARP_F_AROMATIC) && d <= (float)P->aromatic_dist) m |= 1u << ARP_SIFT_AROMATIC;
        if ((fb & ARP_F_HYDROPHOBE) && (fe & ARP_F_HYDROPHOBE) && d <= (float)P->hydrophobic_dist) m |= 1u << ARP_SIFT_HYDROPHOBIC;
    }

    out->i = b; out->j = e; out->mask = m | (cls << ARP_CLASS_SHIFT) | fault; out->dist = d;   /* :936 */
    return 1;
}

/* ------------------------------------------------------------------------ */
/* neighbour search: Bio.PDB.NeighborSearch(atom_list).search_all(radius)    */
/* (call sites interactions.py:1442, :707) -- restated, see header.          */
/* ------------------------------------------------------------------------ */

static int kd_within(const float* a, const float* b, double r2)
{
    double dx = (double)a[0] - (double)b[0];
    double dy = (double)a[1] - (double)b[1];
    double dz = (double)a[2] - (double)b[2];
    double s = 0.0;
    s += dx * dx; s += dy * dy; s += dz * dz;
    return s <= r2;
}

typedef struct { arp_pair* p; uint64_t n, cap; } pair_vec;

static int pv_push(pair_vec* v, const arp_pair* r)
{
    if (v->n == v->cap) {
        uint64_t nc = v->cap ? v->cap * 2 : 4096;
        arp_pair* q = (arp_pair*)realloc(v->p, nc * sizeof(arp_pair));
        if (!q) return -1;
        v->p = q; v->cap = nc;
    }
    v->p[v->n++] = *r;
    return 0;
}

static int cmp_pair(const void* x, const void* y)
{
    const arp_pair* a = (const arp_pair*)x; const arp_pair* b = (const arp_pair*)y;
    if (a->i != b->i) return a->i < b->i ? -1 : 1;
    if (a->j != b->j) return a->j < b->j ? -1 : 1;
    return 0;
}

/* all pairs of one structure [lo, hi) through a uniform grid (cell >= radius) */
static int orc_pairs_range(const orc_env* E, int lo, int hi, pair_vec* out, uint64_t* n_within)
{
    const arp_atoms* A = E->A;
    int n = hi - lo;
    if (n <= 1) return 0;
    double r = E->P->interacting_cutoff, r2 = r * r;
    double w = r > 1e-3 ? r * 1.0001 + 1e-4 : 1e-3;
    float mn[3] = { INFINITY, INFINITY, INFINITY }, mx[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int i = lo; i < hi; ++i)
        for (int k = 0; k < 3; ++k) {
            float v = A->xyz[3 * (size_t)i + k];
            if (v < mn[k]) mn[k] = v;
            if (v > mx[k]) mx[k] = v;
        }
    int64_t dim[3];
    for (;;) {
        for (int k = 0; k < 3; ++k) dim[k] = (int64_t)floor(((double)mx[k] - (double)mn[k]) / w) + 1;
        if (dim[0] * dim[1] * dim[2] <= 4 * (int64_t)n + 64) break;
        w *= 1.5;
    }
    int64_t nc = dim[0] * dim[1] * dim[2];
    int* head = (int*)malloc(sizeof(int) * (size_t)(nc + 1));
    int* order = (int*)malloc(sizeof(int) * (size_t)n);
    int* cell = (int*)malloc(sizeof(int) * (size_t)n);
    if (!head || !order || !cell) { free(head); free(order); free(cell); return -1; }
    memset(head, 0, sizeof(int) * (size_t)(nc + 1));
    for (int i = 0; i < n; ++i) {
        int64_t c[3];
        for (int k = 0; k < 3; ++k) {
            c[k] = (int64_t)floor(((double)A->xyz[3 * (size_t)(lo + i) + k] - (double)mn[k]) / w);
            if (c[k] < 0) c[k] = 0;
            if (c[k] >= dim[k]) c[k] = dim[k] - 1;
        }
        cell[i] = (int)((c[2] * dim[1] + c[1]) * dim[0] + c[0]);
        head[cell[i] + 1]++;
    }
    for (int64_t c = 0; c < nc; ++c) head[c + 1] += head[c];
    int* fill = (int*)malloc(sizeof(int) * (size_t)nc);
    if (!fill) { free(head); free(order); free(cell); return -1; }
    memcpy(fill, head, sizeof(int) * (size_t)nc);
    for (int i = 0; i < n; ++i) order[fill[cell[i]]++] = i;
    free(fill);

    int rc = 0;
    for (int i = 0; i < n && rc == 0; ++i) {
        int c = cell[i];
        int cx = (int)(c % dim[0]), cy = (int)((c / dim[0]) % dim[1]), cz = (int)(c / (dim[0] * dim[1]));
        for (int dz = -1; dz <= 1; ++dz) for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
            int x = cx + dx, y = cy + dy, z = cz + dz;
            if (x < 0 || y < 0 || z < 0 || x >= dim[0] || y >= dim[1] || z >= dim[2]) continue;
            int64_t cc = ((int64_t)z * dim[1] + y) * dim[0] + x;
            for (int k = head[cc]; k < head[cc + 1]; ++k) {
                int j = order[k];
                if (j <= i) continue;                                   /* index1 < index2 */
                if (!kd_within(A->xyz + 3 * (size_t)(lo + i), A->xyz + 3 * (size_t)(lo + j), r2)) continue;
                if (n_within) ++*n_within;
                arp_pair rec;
                if (orc_classify_pair(E, lo + i, lo + j, &rec))
                    if (pv_push(out, &rec)) { rc = -1; break; }
            }
        }
    }
    free(head); free(order); free(cell);
    return rc;
}

/*
 * Oracle for arp_upload_atoms + arp_pairs_run + arp_pairs_fetch(sorted=1).
 * *out is malloc'ed (free with orc_free); records sorted by (i, j).
 */
int orc_pairs(const arp_atoms* A, const arp_params* P, arp_pair** out, uint64_t* n_out, uint64_t* n_within)
{
    orc_env E = { A, P };
    pair_vec v = { 0, 0, 0 };
    if (n_within) *n_within = 0;
    int S = A->n_structures > 0 ? A->n_structures : 1;
    for (int s = 0; s < S; ++s) {
        int lo = A->struct_off ? A->struct_off[s] : 0;
        int hi = A->struct_off ? A->struct_off[s + 1] : A->n_atoms;
        if (orc_pairs_range(&E, lo, hi, &v, n_within)) { free(v.p); return ARP_E_OOM; }
    }
    qsort(v.p, v.n, sizeof(arp_pair), cmp_pair);
    *out = v.p; *n_out = v.n;
    return ARP_OK;
}

/* the per-pair body alone, for explicit (b, e) lists (truth tables, golden checks) */
int orc_classify(const arp_atoms* A, const arp_params* P, const int32_t* b, const int32_t* e, int64_t n,
                 arp_pair* out, uint8_t* emitted)
{
    orc_env E = { A, P };
    for (int64_t k = 0; k < n; ++k) {
        arp_pair r = { b[k], e[k], 0, 0.f };
        emitted[k] = (uint8_t)orc_classify_pair(&E, b[k], e[k], &r);
        out[k] = r;
    }
    return ARP_OK;
}

/* brute-force variant of orc_pairs (O(N^2)); cross-checks the grid search on small inputs */
int orc_pairs_bruteforce(const arp_atoms* A, const arp_params* P, arp_pair** out, uint64_t* n_out)
{
    orc_env E = { A, P };
    pair_vec v = { 0, 0, 0 };
    double r2 = P->interacting_cutoff * P->interacting_cutoff;
    int S = A->n_structures > 0 ? A->n_structures : 1;
    for (int s = 0; s < S; ++s) {
        int lo = A->struct_off ? A->struct_off[s] : 0;
        int hi = A->struct_off ? A->struct_off[s + 1] : A->n_atoms;
        for (int i = lo; i < hi; ++i) for (int j = i + 1; j < hi; ++j) {
            if (!kd_within(A->xyz + 3 * (size_t)i, A->xyz + 3 * (size_t)j, r2)) continue;
            arp_pair rec;
            if (orc_classify_pair(&E, i, j, &rec))
                if (pv_push(&v, &rec)) { free(v.p); return ARP_E_OOM; }
        }
    }
    *out = v.p; *n_out = v.n;
    return ARP_OK;
}

/* ---- per-atom SIFt side effects of the pair loop -------------------------------------
 * The loop of _calculate_atom_contacts, replayed over finished records IN THE GIVEN ORDER
 * (interactions.py:822-852 counters, :924-934 SIFt updates):
 *   utils.update_atom_integer_sift (utils.py:225-242): integer_sift* = sift* + SIFt   (an assignment, from
 *       the binary sift BEFORE this contact -- it is called first, interactions.py:925-926)
 *   utils.update_atom_sift (utils.py:182-199):         sift* |= SIFt
 *   utils.update_atom_fsift (utils.py:202-222):        actual_fsift* |= SIFt[5:]  (= sift* >> 5, not stored)
 * categories: every contact; contact_type == 'INTER'; 'INTRA' in contact_type; 'WATER' in contact_type. */
int orc_atom_sifts(const arp_pair* rec, uint64_t n, int n_atoms, arp_atom_sift* out)
{
    typedef struct { int sift[4][ARP_SIFT_NBITS]; int integer[4][ARP_SIFT_NBITS]; uint32_t hb[4], pl[4]; } atom_state;
    atom_state* st = (atom_state*)calloc((size_t)(n_atoms > 0 ? n_atoms : 1), sizeof(atom_state));
    if (!st) return -1;
    for (uint64_t r = 0; r < n; ++r) {
        int sift[ARP_SIFT_NBITS];
        for (int b = 0; b < ARP_SIFT_NBITS; ++b) sift[b] = (int)(rec[r].mask >> b & 1u);
        const uint32_t cls = (rec[r].mask >> ARP_CLASS_SHIFT) & 7u;
        /* the class names (interactions.py:643-691) as the substring tests see them */
        const int is_inter = cls == ARP_CLASS_INTER;
        const int has_intra = cls == ARP_CLASS_INTRA_NON_SELECTION || cls == ARP_CLASS_INTRA_SELECTION;
        const int has_water = cls == ARP_CLASS_SELECTION_WATER || cls == ARP_CLASS_NON_SELECTION_WATER || cls == ARP_CLASS_WATER_WATER;
        const int atoms[2] = { rec[r].i, rec[r].j };
        for (int side = 0; side < 2; ++side) {
            if (atoms[side] < 0 || atoms[side] >= n_atoms) { free(st); return -2; }
            atom_state* a = st + atoms[side];
            /* counters, interactions.py:822-852: if 'INTRA' ... elif 'INTER' ... elif 'WATER' */
            const int cat = has_intra ? 2 : is_inter ? 1 : has_water ? 3 : 0;
            if (sift[ARP_SIFT_HBOND]) { a->hb[0]++; if (cat) a->hb[cat]++; }
            if (sift[ARP_SIFT_POLAR]) { a->pl[0]++; if (cat) a->pl[cat]++; }
            /* update_atom_integer_sift, then update_atom_sift */
            const int on[4] = { 1, is_inter, has_intra, has_water };
            for (int c = 0; c < 4; ++c) {
                if (!on[c]) continue;
                for (int b = 0; b < ARP_SIFT_NBITS; ++b) a->integer[c][b] = a->sift[c][b] + sift[b];
            }
            for (int c = 0; c < 4; ++c) {
                if (!on[c]) continue;
                for (int b = 0; b < ARP_SIFT_NBITS; ++b) a->sift[c][b] = a->sift[c][b] || sift[b];
            }
        }
    }
    for (int i = 0; i < n_atoms; ++i) {
        memset(&out[i], 0, sizeof out[i]);
        for (int c = 0; c < 4; ++c) {
            for (int b = 0; b < ARP_SIFT_NBITS; ++b) {
                out[i].sift[c] |= (uint16_t)(st[i].sift[c][b] << b);
                out[i].integer_sift[c] |= (uint32_t)st[i].integer[c][b] << (2 * b);
            }
            out[i].hbonds[c] = st[i].hb[c];
            out[i].polars[c] = st[i].pl[c];
        }
    }
    free(st);
    return 0;
}

/* _assign_aromatic_rings_to_residues (interactions.py:1453-1492): closest atom within `radius` of every ring
   centroid; NeighborSearch.search restated as a double-precision scan in ascending atom order, the strict `<`
   of :1471 keeps the first minimum */
int orc_ring_nearest(const float* xyz, int n_atoms, const double* centers, int n_rings, double radius, int blas_fma,
                     int32_t* atom_out, double* dist_out)
{
    const double r2 = radius * radius;
    for (int r = 0; r < n_rings; ++r) {
        const double* c = centers + 3 * (size_t)r;
        int best_i = -1; double best = 0.0;
        for (int i = 0; i < n_atoms; ++i) {
            const float* x = xyz + 3 * (size_t)i;
            double dv[3] = { (double)x[0] - c[0], (double)x[1] - c[1], (double)x[2] - c[2] };
            double s = 0.0; s += dv[0] * dv[0]; s += dv[1] * dv[1]; s += dv[2] * dv[2];
            if (!(s <= r2)) continue;
            double distance = norm3_f64(dv, blas_fma);                          /* :1469 */
            if (best_i < 0 || distance < best) { best_i = i; best = distance; } /* :1471 */
        }
        atom_out[r] = best_i;
        dist_out[r] = best_i >= 0 ? best : 0.0;
    }
    return 0;
}

void orc_free(void* p) { free(p); }

/* binding-site expansion of _make_selection (interactions.py:1420-1424) */
int orc_flag_within(const arp_atoms* A, double radius, uint8_t* flags)
{
    double r2 = radius * radius;
    int S = A->n_structures > 0 ? A->n_structures : 1;
    for (int i = 0; i < A->n_atoms; ++i) flags[i] = (A->feat[i] & ARP_F_IN_SELECTION) ? 1 : 0;
    for (int s = 0; s < S; ++s) {
        int lo = A->struct_off ? A->struct_off[s] : 0;
        int hi = A->struct_off ? A->struct_off[s + 1] : A->n_atoms;
        for (int i = lo; i < hi; ++i) {
            if (!(A->feat[i] & ARP_F_IN_SELECTION)) continue;
            for (int j = lo; j < hi; ++j)
                if (!flags[j] && kd_within(A->xyz + 3 * (size_t)i, A->xyz + 3 * (size_t)j, r2)) flags[j] = 1;
        }
    }
    return ARP_OK;
}

/* ------------------------------------------------------------------------ */
/* plane terms                                                               */
/* ------------------------------------------------------------------------ */

/* fold of utils.group_angle / group_group_angle with degrees=True, signed=True
   followed by the caller's abs() (utils.py:649-660, :680-693), float64 */
static double fold_deg_f64(double cosangle)
{
    double rad = acos(cosangle);
    if (rad > M_PI / 2) rad = rad - M_PI;
    double deg = rad * 180 / M_PI;
    return fabs(deg);
}

/* same in float32 (amide-amide: every operand is float32, python floats are weak) */
static float fold_deg_f32(float cosangle)
{
    float rad = acosf(cosangle);
    if (rad > (float)(M_PI / 2)) { volatile float t = rad - (float)M_PI; rad = t; }
    volatile float deg = rad * 180.0f;
    deg = deg / (float)M_PI;
    return fabsf(deg);
}

static uint32_t plane_class(uint32_t fa, uint32_t fb)
{
    /* interactions.py:1095-1108 (and :985-997, :1253-1265, :1334-1346): four ifs, last true wins;
       the loops only see planes in selection_plus, so INTRA_NON_SELECTION is always overwritten */
    int sa = (fa & ARP_P_IN_SELECTION) != 0, sb = (fb & ARP_P_IN_SELECTION) != 0;
    int pa = (fa & ARP_P_IN_SELECTION_PLUS) != 0, pb = (fb & ARP_P_IN_SELECTION_PLUS) != 0;
    uint32_t c = 7;
    if (!sa && !sb) c = ARP_CLASS_INTRA_NON_SELECTION;
    if (pa && pb) c = ARP_CLASS_INTRA_BINDING_SITE;
    if (sa && sb) c = ARP_CLASS_INTRA_SELECTION;
    if ((sa && !sb) || (sb && !sa)) c = ARP_CLASS_INTER;
    return c;
}

static uint32_t ring_geometry(double dihedral, double theta, const double* bins)
{
    /* interactions.py:1127-1148 */
    double b0 = bins[0], b1 = bins[1], b2 = bins[2];
    if (dihedral <= b0 && theta <= b0) return ARP_G_FF;
    else if (dihedral <= b0 && theta <= b1) return ARP_G_OF;
    else if (dihedral <= b0 && theta <= b2) return ARP_G_EE;
    else if (b0 < dihedral && dihedral <= b1 && theta <= b0) return ARP_G_FT;
    else if (b0 < dihedral && dihedral <= b1 && theta <= b1) return ARP_G_OT;
    else if (b0 < dihedral && dihedral <= b1 && theta <= b2) return ARP_G_ET;
    else if (b1 < dihedral && dihedral <= b2 && theta <= b0) return ARP_G_FE;
    else if (b1 < dihedral && dihedral <= b2 && theta <= b1) return ARP_G_OE;
    else if (b1 < dihedral && dihedral <= b2 && theta <= b2) return ARP_G_EF;
    return ARP_G_NONE;
}

/* one visit (a -> b) of the ring double loop, interactions.py:1110-1155.
   returns 0 if the visit `continue`s */
static int ring_visit(const arp_planes* R, const arp_params* P, int a, int b, uint32_t* geom, double* dist, int* intra)
{
    const double* C = (const double*)R->center; const double* Nn = (const double*)R->normal;
    const double* ca = C + 3 * (size_t)a; const double* cb = C + 3 * (size_t)b;
    const double* na = Nn + 3 * (size_t)a; const double* nb = Nn + 3 * (size_t)b;
    *intra = (R->res_id[a] == R->res_id[b]);                                   /* :1091 */
    double tp[3] = { ca[0] - cb[0], ca[1] - cb[1], ca[2] - cb[2] };
    double distance = norm3_f64(tp, P->blas_fma);                              /* :1111 */
    if (distance > P->ring_centroid_dist) return 0;                            /* :1113 */
    double cd = dot3_f64(na, nb, P->blas_fma) / (norm3_f64(na, P->blas_fma) * norm3_f64(nb, P->blas_fma));
    double ct = dot3_f64(na, tp, P->blas_fma) / (norm3_f64(na, P->blas_fma) * norm3_f64(tp, P->blas_fma));
    double dihedral = fold_deg_f64(cd);                                        /* :1122 */
    double theta = fold_deg_f64(ct);                                           /* :1123 */
    uint32_t g = ring_geometry(dihedral, theta, P->plane_bins_deg);
    if (*intra && g == ARP_G_EE) return 0;                                     /* :1154 */
    *geom = g; *dist = distance;
    return 1;
}

typedef struct { arp_plane_pair* p; uint64_t n, cap; } pp_vec;
static int ppv_push(pp_vec* v, const arp_plane_pair* r)
{
    if (v->n == v->cap) {
        uint64_t nc = v->cap ? v->cap * 2 : 1024;
        arp_plane_pair* q = (arp_plane_pair*)realloc(v->p, nc * sizeof(arp_plane_pair));
        if (!q) return -1;
        v->p = q; v->cap = nc;
    }
    v->p[v->n++] = *r;
    return 0;
}

/* __calculate_plane_plane_contacts (interactions.py:1064-1194), literally: every ordered
   visit (a, b) that survives either creates a record (bgn=a, end=b) or appends its geometry
   label to the record the earlier visit of the same unordered pair created (:1181-1194).
   Records come out in creation order. */
int orc_ring_ring(const arp_planes* R, const arp_params* P, arp_plane_pair** out, uint64_t* n_out)
{
    pp_vec v = { 0, 0, 0 };
    /* per-ring list of record indices whose bgn or end is that ring (the `identity` filter :1181) */
    int** adj = (int**)calloc((size_t)(R->n > 0 ? R->n : 1), sizeof(int*));
    int* adj_n = (int*)calloc((size_t)(R->n > 0 ? R->n : 1), sizeof(int));
    int* adj_c = (int*)calloc((size_t)(R->n > 0 ? R->n : 1), sizeof(int));
    int rc = ARP_OK;
    for (int a = 0; a < R->n && rc == ARP_OK; ++a) {
        for (int b = 0; b < R->n && rc == ARP_OK; ++b) {
            if (!(R->flags[a] & ARP_P_IN_SELECTION_PLUS) || !(R->flags[b] & ARP_P_IN_SELECTION_PLUS)) continue; /* :1081 */
            if (a == b) continue;                                              /* :1085 */
            uint32_t g; double d; int intra;
            if (!ring_visit(R, P, a, b, &g, &d, &intra)) continue;
            int found = -1;
            for (int k = 0; k < adj_n[a]; ++k) {
                arp_plane_pair* q = &v.p[adj[a][k]];
                if ((q->a == a && q->b == b) || (q->a == b && q->b == a)) { found = adj[a][k]; break; }
            }
            if (found >= 0) {
                arp_plane_pair* q = &v.p[found];
                uint32_t g1 = q->code & 0xF, g2 = (q->code >> 4) & 0xF;
                if (g != g1 && (g2 == 0xF || g != g2)) {                       /* :1184-1185 */
                    if (g2 == 0xF) q->code = (q->code & ~0xF0u) | (g << 4);
                }
                continue;
            }
            arp_plane_pair rec; memset(&rec, 0, sizeof rec);
            rec.a = a; rec.b = b; rec.dist = d;
            rec.code = g | (0xFu << 4) | (plane_class(R->flags[a], R->flags[b]) << 8) | ((uint32_t)intra << 11);
            if (ppv_push(&v, &rec)) { rc = ARP_E_OOM; break; }
            int idx = (int)v.n - 1;
            int ends[2] = { a, b };
            for (int t = 0; t < 2; ++t) {
                int r = ends[t];
                if (adj_n[r] == adj_c[r]) {
                    adj_c[r] = adj_c[r] ? adj_c[r] * 2 : 4;
                    int* q = (int*)realloc(adj[r], sizeof(int) * (size_t)adj_c[r]);
                    if (!q) { rc = ARP_E_OOM; break; }
                    adj[r] = q;
                }
                adj[r][adj_n[r]++] = idx;
            }
        }
    }
    for (int r = 0; r < R->n; ++r) free(adj[r]);
    free(adj); free(adj_n); free(adj_c);
    if (rc != ARP_OK) { free(v.p); return rc; }
    *out = v.p; *n_out = v.n;
    return ARP_OK;
}

/* __calculate_group_group_contacts (interactions.py:1217-1300): float32 throughout */
int orc_amide_amide(const arp_planes* Am, const arp_params* P, arp_plane_pair** out, uint64_t* n_out)
{
    pp_vec v = { 0, 0, 0 };
    const float* C = (const float*)Am->center; const float* Nn = (const float*)Am->normal;
    for (int a = 0; a < Am->n; ++a) {
        if (!(Am->flags[a] & ARP_P_IN_SELECTION_PLUS)) continue;              /* :1224 */
        for (int b = 0; b < Am->n; ++b) {
            if (a == b) continue;                                             /* :1233 */
            if (!(Am->flags[b] & ARP_P_IN_SELECTION_PLUS)) continue;          /* :1237 */
            int intra = Am->res_id[a] == Am->res_id[b];
            const float* ca = C + 3 * (size_t)a; const float* cb = C + 3 * (size_t)b;
            const float* na = Nn + 3 * (size_t)a; const float* nb = Nn + 3 * (size_t)b;
            float tp[3] = { ca[0] - cb[0], ca[1] - cb[1], ca[2] - cb[2] };
            float distance = norm3_f32(tp);                                   /* :1268 */
            if (distance > (float)P->amide_centroid_dist) continue;           /* :1270 */
            volatile float den1 = norm3_f32(na) * norm3_f32(nb);
            volatile float cd = dot3_f32(na, nb) / den1;
            volatile float den2 = norm3_f32(na) * norm3_f32(tp);
            volatile float ct = dot3_f32(na, tp) / den2;
            float dihedral = fold_deg_f32(cd), theta = fold_deg_f32(ct);       /* :1278-1279 */
            if (dihedral > (float)P->plane_bins_deg[0] || theta > (float)P->plane_bins_deg[0]) continue;  /* :1282 */
            arp_plane_pair rec; memset(&rec, 0, sizeof rec);
            rec.a = a; rec.b = b; rec.dist = (double)distance;
            rec.code = 0xFFu | (plane_class(Am->flags[a], Am->flags[b]) << 8) | ((uint32_t)intra << 11);
            if (ppv_push(&v, &rec)) { free(v.p); return ARP_E_OOM; }
        }
    }
    *out = v.p; *n_out = v.n;
    return ARP_OK;
}

/* __calculate_group_plane_contacts (interactions.py:1302-1382): float32 amide operands are
   widened by the float64 ring operands */
int orc_amide_ring(const arp_planes* Am, const arp_planes* R, const arp_params* P, arp_plane_pair** out, uint64_t* n_out)
{
    pp_vec v = { 0, 0, 0 };
    const float* AC = (const float*)Am->center; const float* AN = (const float*)Am->normal;
    const double* RC = (const double*)R->center; const double* RN = (const double*)R->normal;
    for (int a = 0; a < Am->n; ++a) {
        if (!(Am->flags[a] & ARP_P_IN_SELECTION_PLUS)) continue;              /* :1309 */
        for (int r = 0; r < R->n; ++r) {
            if (!(R->flags[r] & ARP_P_IN_SELECTION_PLUS)) continue;           /* :1318 */
            int intra = Am->res_id[a] == R->res_id[r];
            const float* ca = AC + 3 * (size_t)a; const float* na = AN + 3 * (size_t)a;
            const double* cr = RC + 3 * (size_t)r; const double* nr = RN + 3 * (size_t)r;
            double tp[3] = { (double)ca[0] - cr[0], (double)ca[1] - cr[1], (double)ca[2] - cr[2] };
            double distance = norm3_f64(tp, P->blas_fma);                     /* :1349 */
            if (distance > P->amide_centroid_dist) continue;                  /* :1351 */
            double nad[3] = { (double)na[0], (double)na[1], (double)na[2] };
            float na_norm = norm3_f32(na);                                    /* float32 scalar */
            double cd = dot3_f64(nad, nr, P->blas_fma) / ((double)na_norm * norm3_f64(nr, P->blas_fma));
            double ct = dot3_f64(nad, tp, P->blas_fma) / ((double)na_norm * norm3_f64(tp, P->blas_fma));
            double dihedral = fold_deg_f64(cd), theta = fold_deg_f64(ct);      /* :1359-1360 */
            if (dihedral > P->plane_bins_deg[0] || theta > P->plane_bins_deg[0]) continue;   /* :1363 */
            arp_plane_pair rec; memset(&rec, 0, sizeof rec);
            rec.a = a; rec.b = r; rec.dist = distance;
            rec.code = 0xFFu | (plane_class(Am->flags[a], R->flags[r]) << 8) | ((uint32_t)intra << 11);
            if (ppv_push(&v, &rec)) { free(v.p); return ARP_E_OOM; }
        }
    }
    *out = v.p; *n_out = v.n;
    return ARP_OK;
}

/* __calculate_atom_plane_contacts (interactions.py:947-1062).  NeighborSearch.search(center, r)
   is restated as a double-precision scan; results sorted by (ring, atom). */
int orc_atom_ring(const arp_atoms* A, const arp_planes* R, const arp_params* P, arp_atom_plane** out, uint64_t* n_out)
{
    arp_atom_plane* v = 0; uint64_t n = 0, cap = 0;
    const double* RC = (const double*)R->center; const double* RN = (const double*)R->normal;
    double r2 = P->met_sulphur_dist * P->met_sulphur_dist;                     /* :960 search radius */
    for (int r = 0; r < R->n; ++r) {
        if (!(R->flags[r] & ARP_P_IN_SELECTION_PLUS)) continue;               /* :957 */
        const double* c = RC + 3 * (size_t)r; const double* nr = RN + 3 * (size_t)r;
        for (int i = 0; i < A->n_atoms; ++i) {
            const float* x = A->xyz + 3 * (size_t)i;
            double dx = (double)x[0] - c[0], dy = (double)x[1] - c[1], dz = (double)x[2] - c[2];
            double s = 0.0; s += dx * dx; s += dy * dy; s += dz * dz;
            if (!(s <= r2)) continue;                                          /* kdtrees: within radius */
            uint32_t f = A->feat[i];
            if (f & ARP_F_ELEM_H) continue;                                    /* :964 */
            double dv[3] = { dx, dy, dz };
            double distance = norm3_f64(dv, P->blas_fma);                      /* :972 */
            if (f & ARP_F_AROMATIC) continue;                                  /* :975 */
            int intra = R->res_id[r] == A->res_id[i];                          /* :981 */
            /* contact type :985-997 (atom in selection_plus always holds for the SoA atoms) */
            int sr = (R->flags[r] & ARP_P_IN_SELECTION) != 0, sa = (f & ARP_F_IN_SELECTION) != 0;
            uint32_t cls = ARP_CLASS_INTRA_BINDING_SITE;
            if (sr && sa) cls = ARP_CLASS_INTRA_SELECTION;
            if ((sr && !sa) || (sa && !sr)) cls = ARP_CLASS_INTER;
            double p[3] = { c[0] - (double)x[0], c[1] - (double)x[1], c[2] - (double)x[2] };
            double ct = dot3_f64(nr, p, P->blas_fma) / (norm3_f64(nr, P->blas_fma) * norm3_f64(p, P->blas_fma));
            double theta = fold_deg_f64(ct);                                   /* :1005 */
            uint32_t lab = 0;
            if (distance <= P->atom_ring_dist && theta <= P->plane_bins_deg[0]) {   /* :1007 */
                if ((f & ARP_F_ELEM_C) && (f & ARP_F_WEAK_HBOND_DONOR)) lab |= ARP_AP_CARBONPI;
                if (f & ARP_F_POS_IONISABLE) lab |= ARP_AP_CATIONPI;
                if (f & ARP_F_HBOND_DONOR) lab |= ARP_AP_DONORPI;
                if (f & ARP_F_XBOND_DONOR) lab |= ARP_AP_HALOGENPI;
            }
            if (distance <= P->met_sulphur_dist) {                             /* :1021 */
                if (f & ARP_F_MET_SULPHUR) lab |= ARP_AP_METSULPHURPI;
            }
            if (!lab) continue;                                                /* :1026 */
            if (n == cap) {
                cap = cap ? cap * 2 : 1024;
                arp_atom_plane* q = (arp_atom_plane*)realloc(v, cap * sizeof *v);
                if (!q) { free(v); return ARP_E_OOM; }
                v = q;
            }
            memset(&v[n], 0, sizeof v[n]);
            v[n].atom = i; v[n].ring = r; v[n].dist = distance;
            v[n].code = lab | (cls << 8) | ((uint32_t)intra << 11);
            ++n;
        }
    }
    *out = v; *n_out = n;
    return ARP_OK;
}

/* exposed for tests of the arithmetic models against live NumPy */
double orc_dot3_f64(const double* x, const double* y, int blas_fma) { return dot3_f64(x, y, blas_fma); }
double orc_norm3_f64(const double* v, int blas_fma) { return norm3_f64(v, blas_fma); }
float  orc_dot3_f32(const float* x, const float* y) { return dot3_f32(x, y); }
float  orc_norm3_f32(const float* v) { return norm3_f32(v); }
float  orc_dist_f32_pub(const float* a, const float* b) { return orc_dist_f32(a, b); }
double orc_fold_deg_f64(double c) { return fold_deg_f64(c); }
float  orc_fold_deg_f32(float c) { return fold_deg_f32(c); }
double orc_get_angle_fdf(const float* a, const double* b, const float* c) { return get_angle_fdf(a, b, c); }
double orc_get_angle_ffd(const float* a, const float* b, const double* c) { return get_angle_ffd(a, b, c); }
double orc_get_angle_fff(const float* a, const float* b, const float* c) { int f; float t; return get_angle_fff(a, b, c, &f, &t); }
