/*
 * arpeggio_cuda.h -- C ABI of libarpeggio_cuda.so
 *
 * B200-native (sm_100a) implementation of the interatomic-contact hot path of
 * pdbe-arpeggio.  The reference has no FFI of its own: the seam is the Python
 * method surface of arpeggio.core.InteractionComplex.  Every entry point below
 * names the reference code (file:line, relative to the pdbe-arpeggio tree) whose
 * work it replaces.  The Python host side (arpeggio_b200/) binds these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C linkage, plain pointers and sizes; no C++/torch types.
 *   - every function returns ARP_OK (0) or a negative ARP_E_* code; the text of
 *     the last error is available from arp_last_error().
 *   - host input buffers are borrowed for the duration of the call only.
 *   - a context is bound to one device and one stream; it is not re-entrant.
 *     Different contexts may be driven from different host threads.
 *   - there is no CPU fallback: without a CUDA device arp_create() fails.
 */
#ifndef ARPEGGIO_CUDA_H
#define ARPEGGIO_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARP_ABI_VERSION 6

/* ---- error codes -------------------------------------------------------- */
#define ARP_OK              0
#define ARP_E_INVALID_ARG  -1
#define ARP_E_CUDA         -2
#define ARP_E_OOM          -3
#define ARP_E_CAPACITY     -4   /* caller buffer too small; nothing written   */
#define ARP_E_NOT_READY    -5   /* fetch before run, run before upload, ...   */
#define ARP_E_NO_DEVICE    -6

/* ---- SIFt bit positions of arp_pair.mask (interactions.py:178-180) ------ */
#define ARP_SIFT_CLASH        0
#define ARP_SIFT_COVALENT     1
#define ARP_SIFT_VDW_CLASH    2
#define ARP_SIFT_VDW          3
#define ARP_SIFT_PROXIMAL     4
#define ARP_SIFT_HBOND        5
#define ARP_SIFT_WEAK_HBOND   6
#define ARP_SIFT_XBOND        7
#define ARP_SIFT_IONIC        8
#define ARP_SIFT_METAL        9
#define ARP_SIFT_AROMATIC    10
#define ARP_SIFT_HYDROPHOBIC 11
#define ARP_SIFT_CARBONYL    12
#define ARP_SIFT_POLAR       13
#define ARP_SIFT_WEAK_POLAR  14
#define ARP_SIFT_NBITS       15
/* entity class (interacting_entities) lives in mask bits 16..18 */
#define ARP_CLASS_SHIFT      16
#define ARP_CLASS_MASK       0x7u
#define ARP_CLASS_INTRA_NON_SELECTION 0   /* interactions.py:669-670 */
#define ARP_CLASS_INTRA_SELECTION     1   /* :672-673 */
#define ARP_CLASS_INTER               2   /* :675-676 */
#define ARP_CLASS_SELECTION_WATER     3   /* :678-679 */
#define ARP_CLASS_NON_SELECTION_WATER 4   /* :681-682 */
#define ARP_CLASS_WATER_WATER         5   /* :684-685 */
#define ARP_CLASS_INTRA_BINDING_SITE  6   /* planes only: :991, :1101, :1259, :1340 */

/* ---- per-atom feature bits (arp_atoms.feat) ----------------------------- */
/* bits 0..11: the 12 keys of config.ATOM_TYPES (config.py:53-145)           */
#define ARP_F_HBOND_ACCEPTOR      (1u << 0)
#define ARP_F_HBOND_DONOR         (1u << 1)
#define ARP_F_WEAK_HBOND_ACCEPTOR (1u << 2)
#define ARP_F_WEAK_HBOND_DONOR    (1u << 3)
#define ARP_F_XBOND_ACCEPTOR      (1u << 4)
#define ARP_F_XBOND_DONOR         (1u << 5)
#define ARP_F_POS_IONISABLE       (1u << 6)
#define ARP_F_NEG_IONISABLE       (1u << 7)
#define ARP_F_HYDROPHOBE          (1u << 8)
#define ARP_F_CARBONYL_OXYGEN     (1u << 9)
#define ARP_F_CARBONYL_CARBON     (1u << 10)
#define ARP_F_AROMATIC            (1u << 11)
#define ARP_F_IS_METAL            (1u << 12)  /* interactions.py:1990 */
#define ARP_F_IS_HALOGEN          (1u << 13)  /* interactions.py:1991 */
#define ARP_F_IS_WATER            (1u << 14)  /* get_full_id()[3][0]=='W' */
#define ARP_F_IN_SELECTION        (1u << 15)  /* atom in self.selection    */
#define ARP_F_ELEM_H              (1u << 16)  /* element.strip()=='H' (:712, :964) */
#define ARP_F_ELEM_C              (1u << 17)  /* element=='C'  (:1009)     */
#define ARP_F_MET_SULPHUR         (1u << 18)  /* resname=='MET' and element=='S' (:1023) */
#define ARP_F_HAS_XNBR            (1u << 19)  /* get_single_bond_neighbour() is not None (utils.py:612-635) */

/* ---- per-residue flag bits (arp_atoms.res_flags) ------------------------ */
#define ARP_R_IS_POLYPEPTIDE  (1u << 0)   /* residue.is_polypeptide (interactions.py:1671, :1860) */
#define ARP_R_HAS_LINKS       (1u << 1)   /* hasattr prev_residue and next_residue (:736-737, :1687-1688) */

/* ---- plane flags (arp_planes.flags) -------------------------------------- */
#define ARP_P_IN_SELECTION       (1u << 0)  /* id in selection_ring_ids / selection_amide_ids (:1416-1417) */
#define ARP_P_IN_SELECTION_PLUS  (1u << 1)  /* id in selection_plus_*_ids (:1434-1437) */

/* ---- plane-plane geometry codes (interactions.py:1129-1148) ------------- */
enum { ARP_G_FF = 0, ARP_G_OF, ARP_G_EE, ARP_G_FT, ARP_G_OT, ARP_G_ET,
       ARP_G_FE, ARP_G_OE, ARP_G_EF, ARP_G_NONE /* '' : NaN angle */ };

/* ---- atom-plane label bits (interactions.py:1007-1024) ------------------ */
#define ARP_AP_CARBONPI      (1u << 0)
#define ARP_AP_CATIONPI      (1u << 1)
#define ARP_AP_DONORPI       (1u << 2)
#define ARP_AP_HALOGENPI     (1u << 3)
#define ARP_AP_METSULPHURPI  (1u << 4)

typedef struct arp_ctx arp_ctx;

/*
 * Run-time parameters.  The distance/angle thresholds are the values of
 * config.CONTACT_TYPES (config.py:592-660), kept as doubles exactly as the
 * reference holds them (Python floats); the kernels narrow them to float32
 * where NumPy (NEP 50) would.  The cos_* members are the images of the angle
 * thresholds under the HOST's arccos, found by bisection (see
 * arpeggio_b200/params.py): the GPU never evaluates acos, it compares the
 * cosine, which is what makes the angle bits identical to the host's libm.
 * arp_params_default() fills all of it from the C library's acos/acosf.
 */
typedef struct arp_params {
    double interacting_cutoff;      /* run_arpeggio(..., interacting_cutoff) interactions.py:329 */
    double vdw_comp;                /* vdw_comp_factor */
    int32_t include_sequence_adjacent;
    int32_t blas_fma;               /* 1: np.dot/np.linalg.norm on float64 = FMA chain (OpenBLAS Haswell/SkylakeX ddot); 0: plain sequential */
    double h_vdw;                   /* config.VDW_RADII['H'] = 1.2 (config.py:23-25) */
    double dist_max;                /* CONTACT_TYPES_DIST_MAX 4.5 */
    double hbond_polar_dist;        /* 3.5 */
    double weak_polar_dist;         /* 3.5 */
    double ionic_dist;              /* 4.0 */
    double carbonyl_dist;           /* 3.6 */
    double aromatic_dist;           /* 4.0 */
    double hydrophobic_dist;        /* 4.5 */
    double metal_dist;              /* 2.8 */
    double hbond_angle;             /* 1.57 rad */
    double weak_hbond_angle;        /* 2.27 rad */
    double cx_angle_min;            /* 0.52 rad */
    double cx_angle_max;            /* 2.62 rad */
    double xbond_angle;             /* 2.09 rad */
    double ring_centroid_dist;      /* 6.0 aromatic.centroid_distance */
    double atom_ring_dist;          /* 4.5 aromatic.atom_aromatic_distance */
    double met_sulphur_dist;        /* 6.0 aromatic.met_sulphur_aromatic_distance */
    double amide_centroid_dist;     /* 6.0 amide.centroid_distance */
    double plane_bins_deg[3];       /* 30, 60, 90 (interactions.py:1129-1148, :1007, :1282) */
    /* ---- cosine-domain images (used by the CUDA path only) ---- */
    double cos_hbond;               /* largest c in [-1,1] with arccos(c) >= hbond_angle */
    double cos_weak_hbond;          /* ... >= weak_hbond_angle */
    double cos_cx_min;              /* largest c with arccos(c) >= cx_angle_min */
    double cos_cx_max;              /* smallest c with arccos(c) <= cx_angle_max */
    float  cos_xbond_f32;           /* largest float32 c with arccosf(c) >= f32(xbond_angle) */
    float  _pad0;
    /* folded plane angle |deg(c)| <= bins[k]:  pos branch: c >= cos_pos[k];
       neg branch (arccos(c) > pi/2, i.e. c <= cos_split): c <= cos_neg[k]   */
    double cos_split_f64;           /* largest c with arccos(c) > pi/2 */
    double cos_pos_f64[3];
    double cos_neg_f64[3];
    float  cos_split_f32;
    float  cos_pos_f32[3];
    float  cos_neg_f32[3];
    float  _pad1;
} arp_params;

/*
 * Structure-of-arrays atom input.  Index = position in the reference's
 * `selection_plus` list (interactions.py:1426, :1442), so that
 * atom_bgn = lower index (Bio.PDB.NeighborSearch.search_all reports
 * index1 < index2).  A batch of S independent structures is the concatenation
 * of their atoms (and residues); struct_off gives the atom ranges and no pair
 * ever spans two structures.
 */
typedef struct arp_atoms {
    int32_t n_atoms;          /* N  */
    int32_t n_residues;       /* Rs */
    int32_t n_rad_classes;    /* K  */
    int32_t n_structures;     /* S >= 1 */
    const float*    xyz;        /* [N][3] exact Atom.coord (float32, protein_reader.py:327) */
    const uint32_t* feat;       /* [N] ARP_F_* */
    const int32_t*  res_id;     /* [N] index into the residue arrays */
    const uint16_t* rad_class;  /* [N] index into vdw/cov */
    const double*   vdw;        /* [K] ob.GetVdwRad      (interactions.py:1501) */
    const double*   cov;        /* [K] ob.GetCovalentRad (interactions.py:1509) */
    const int32_t*  res_prev;   /* [Rs] residue index or -1 (interactions.py:1690) */
    const int32_t*  res_next;   /* [Rs] residue index or -1 (interactions.py:1693) */
    const uint8_t*  res_flags;  /* [Rs] ARP_R_* */
    const int32_t*  bond_off;   /* [N+1] CSR of OBAtomAtomIter neighbours (interactions.py:750); NULL = no bonds */
    const int32_t*  bond_nbr;   /* [E] atom indices */
    const int32_t*  h_off;      /* [N+1] CSR of atom.h_coords (interactions.py:1529); NULL = none */
    const double*   h_xyz;      /* [H][3] float64 */
    const float*    xnbr_xyz;   /* [N][3] coord of get_single_bond_neighbour (valid where ARP_F_HAS_XNBR); NULL allowed */
    const int32_t*  struct_off; /* [S+1] atom offsets; NULL when S == 1 */
    /* ---- optional wire forms (all NULL / 0 in a zero-initialised struct): the same arrays in fewer bytes over PCIe,
       decoded on the device after the copy; results are those of the plain forms ---- */
    const uint8_t*  bond_cnt;   /* [N] neighbours per atom in place of bond_off (which must then be NULL); bond_nbr holds n_bond_nbr indices */
    const uint8_t*  h_cnt;      /* [N] hydrogens per atom in place of h_off; n_h hydrogens follow in h_xyz or h_fix */
    const int32_t*  h_fix;      /* [n_h][3] fixed-point hydrogen coordinates in place of h_xyz: coordinate = (double)h_fix / h_fix_scale
                                   (IEEE division).  Lossless for coordinates that are decimal fractions of the file's precision
                                   (3 decimals: scale 1000) -- the CALLER checks h_fix / scale == h_xyz before choosing this form */
    const int32_t*  xnbr_idx;   /* [n_xnbr] strictly ascending atom indices: xnbr_xyz is then [n_xnbr][3], the rows of just these atoms */
    double          h_fix_scale;
    int32_t         n_bond_nbr; /* with bond_cnt: sum of the counts (checked) */
    int32_t         n_h;        /* with h_cnt: sum of the counts (checked) */
    int32_t         n_xnbr;
    int32_t         _pad;
} arp_atoms;

/* ring: centre/normal float64 (OBRing.findCenterAndNormal, interactions.py:1708-1717);
   amide: centre/normal float32 (interactions.py:1564-1578) */
typedef struct arp_planes {
    int32_t n;
    int32_t is_f32;             /* 0: center/normal point at double[n][3]; 1: at float[n][3] */
    const void*     center;
    const void*     normal;
    const int32_t*  res_id;     /* [n] residue index of ring['residue'] / amide['residue'] (-1: None) */
    const uint32_t* flags;      /* [n] ARP_P_* */
} arp_planes;

/* ---- output records ------------------------------------------------------ */
typedef struct arp_pair {       /* AtomAtomContact, interactions.py:28-29, :936 */
    int32_t  i;                 /* bgn atom index (i < j) */
    int32_t  j;                 /* end atom index */
    uint32_t mask;              /* bits 0..14 SIFt, bits 16..18 entity class */
    float    dist;              /* np.linalg.norm(bgn.coord - end.coord), float32 */
} arp_pair;

/* compact view of the (i, j)-sorted stream (arp_pairs_fetch_compact): the records of bgn atom i are rows
   row_off[i] .. row_off[i + 1] - 1; `i` is not stored and the float32 distances travel as a separate stream that is
   fetched only on demand: 8 bytes per record + 4 per atom over PCIe instead of 16 per record. */
typedef struct arp_pair_c {
    int32_t  j;                 /* end atom index (ascending inside a row) */
    uint32_t mask;              /* as arp_pair.mask */
} arp_pair_c;

typedef struct arp_plane_pair { /* PlanePlaneContact, interactions.py:23-26 */
    int32_t  a;                 /* bgn plane index */
    int32_t  b;                 /* end plane index */
    uint32_t code;              /* bits 0..3 first geometry code, 4..7 second (0xF = none),
                                   bits 8..10 entity class, bit 11 intra-residue */
    uint32_t _pad;
    double   dist;              /* centroid distance (float32 value widened for amide-amide) */
} arp_plane_pair;

typedef struct arp_atom_plane { /* AtomPlaneContact, interactions.py:19-21 */
    int32_t  atom;
    int32_t  ring;
    uint32_t code;              /* bits 0..4 ARP_AP_*, bits 8..10 entity class, bit 11 intra-residue */
    uint32_t _pad;
    double   dist;
} arp_atom_plane;

/* per-atom SIFt reductions: what the pair loop leaves on every atom as a side effect
   (interactions.py:822-852, :924-934 through utils.update_atom_integer_sift / update_atom_sift /
   update_atom_fsift, utils.py:182-242).  Category k: 0 every contact, 1 contact_type == 'INTER',
   2 'INTRA' in contact_type, 3 'WATER' in contact_type. */
typedef struct arp_atom_sift {
    uint16_t sift[4];           /* atom.sift, .sift_inter_only, .sift_intra_only, .sift_water_only: bit b = SIFt[b];
                                   atom.actual_fsift* is the same word >> 5 (SIFt[5:]) */
    uint32_t integer_sift[4];   /* atom.integer_sift*: 2 bits per SIFt position, values 0..2.  The reference ASSIGNS
                                   sift-before-this-contact + SIFt at every contact (utils.py:233), so the value is
                                   that of the atom's LAST contact of the category in loop order */
    uint32_t hbonds[4];         /* atom.actual_hbonds, _inter_only, _intra_only, _water_only (interactions.py:822-836) */
    uint32_t polars[4];         /* atom.actual_polars, ...                                  (interactions.py:838-852) */
} arp_atom_sift;

typedef struct arp_stats {
    uint64_t n_pairs;           /* records emitted by the last arp_pairs_run */
    uint64_t n_candidates;      /* distance tests performed */
    uint64_t n_cells;           /* grid cells (all structures) */
    uint64_t n_cells_nonempty;
    uint64_t input_bytes;       /* algorithmic input bytes (sum of uploaded array sizes) */
    uint64_t output_bytes;      /* 16 * n_pairs */
    float    ms_total;          /* CUDA-event time of the last run, first to last kernel */
    float    ms_grid;           /* cell build part   */
    float    ms_search;         /* search kernel: neighbour search + filters -> hit list */
    float    ms_classify;       /* classify kernels: distance + angle + bitmask rules -> records (includes ms_hscan) */
    float    ms_hscan;          /* of which the deferred hydrogen / halogen / xbond predicates */
    float    ms_pairs;          /* the three pair kernels back to back (no events between them) */
    uint32_t faults;            /* device-side diagnostics of the last run: ARP_FAULT_* */
    uint32_t _pad;
} arp_stats;

#define ARP_FAULT_NONFINITE  1u  /* a NaN / Inf coordinate: that structure was searched as ONE cell (every pair tested);
                                    pairs with such an atom never compare true, as in the reference */
#define ARP_FAULT_HANDOFF    2u  /* the early start of k_classify timed out once; the run was repeated without it */

/* ---- life cycle ---------------------------------------------------------- */
int  arp_abi_version(void);
int  arp_device_count(void);                       /* <0 on error, 0 if no CUDA device */
int  arp_create(int device, arp_ctx** out);
void arp_destroy(arp_ctx* ctx);
const char* arp_last_error(arp_ctx* ctx);          /* ctx may be NULL: last create error */

/* fills *p with config.CONTACT_TYPES defaults (config.py:592-660) and the
   cosine images computed from the C library's acos()/acosf() */
int  arp_params_default(arp_params* p);
int  arp_set_params(arp_ctx* ctx, const arp_params* p);

/* pinned host memory for zero-staging uploads/fetches (optional) */
int  arp_host_alloc(void** ptr, uint64_t bytes);
int  arp_host_free(void* ptr);

/* ---- atom-atom contacts --------------------------------------------------
 * replaces NeighborSearch(selection_plus) (interactions.py:1442) +
 * InteractionComplex._calculate_atom_contacts (interactions.py:693-936) +
 * the predicates utils.is_hbond/is_weak_hbond/is_halogen_weak_hbond/is_xbond
 * (utils.py:73-179) and utils.get_angle (utils.py:696-745).                  */
/* Preconditions of arp_upload_atoms / arp_upload_atoms_batch the library does NOT check (a host pass over the arrays would
   cost as much as the kernels of a small structure; arpeggio_b200.soa.AtomSoA.validate checks them on the Python side):
   0 <= res_id[i] < n_residues, rad_class[i] < n_rad_classes, 0 <= bond_nbr[k] < n_atoms, res_prev / res_next in
   [-1, n_residues), bond_off / h_off non-decreasing.  Out-of-range values are out-of-bounds reads on the device.
   Checked: sizes, NULL arrays, CSR first entries, struct_off.  Non-finite coordinates are legal (ARP_FAULT_NONFINITE). */
int  arp_upload_atoms(arp_ctx* ctx, const arp_atoms* atoms);     /* async H2D */
/* A batch of independent structures as they are -- one arp_atoms each (n_structures <= 1), indices local to the
   structure, radius tables of their own: every structure travels with its own DMA(s) and the DEVICE concatenates
   them (residue / bond / hydrogen indices rebased, radius classes merged), so that the batch runs as one launch
   sequence.  Atom indices of the results are global: structure s owns [sum of n_atoms before s, + n_atoms). */
int  arp_upload_atoms_batch(arp_ctx* ctx, const arp_atoms* const* parts, int32_t n_parts);
int  arp_pairs_run(arp_ctx* ctx, uint64_t* n_pairs);             /* kernels; returns the record count */
int  arp_pairs_fetch(arp_ctx* ctx, arp_pair* dst, uint64_t cap, int sorted); /* D2H; sorted!=0: (i,j) ascending */
int  arp_pairs_device_ptr(arp_ctx* ctx, const arp_pair** dptr);  /* device pointer of the record stream */
/* The same job without the host waiting for it ("arp_pairs_run is asynchronous; fetch / sync block"):
   arp_pairs_run_async only enqueues; the first of arp_pairs_count / arp_pairs_fetch* / arp_atom_sifts_run /
   arp_pairs_device_ptr waits for the run (and repeats it if the record buffer was too small). */
int  arp_pairs_run_async(arp_ctx* ctx);
int  arp_pairs_count(arp_ctx* ctx, uint64_t* n_pairs);
/* Compact D2H of the (i, j)-sorted stream: row_off[n_atoms + 1], rec[cap >= n_pairs], dist[cap] or NULL (the distances
   stay on the device and can be fetched later with arp_pairs_fetch_dist, same order).  *n_pairs is set even when the
   call fails with ARP_E_CAPACITY, so the caller can size its buffers and call again (the run is not repeated). */
int  arp_pairs_fetch_compact(arp_ctx* ctx, uint32_t* row_off, arp_pair_c* rec, uint64_t cap, float* dist, uint64_t* n_pairs);
int  arp_pairs_fetch_dist(arp_ctx* ctx, float* dist, uint64_t cap);
/* The sorted stream PACKED: one word per record = j in its low *bits_j bits (bits_j = bits of n_atoms - 1; of the largest structure's in a batch, where j is local to the structure), the 15 SIFt
   bits above them; lo32[k] holds bits 0..31 of record k, hi8[k] bits 32..39 (only when bits_j + 15 > 32, i.e. more than
   131072 atoms; may be NULL otherwise): 4 (or 5) bytes per record + 4 per atom over PCIe.  Not in the word: the entity
   class -- a function of the two atoms' ARP_F_IN_SELECTION / ARP_F_IS_WATER flags (interactions.py:643-691), which
   arp_pairs_unpack_packed recomputes from the caller's feat array -- and the xbond-without-neighbour fault bit:
   *n_faults counts the records that have it (fetch the compact view to see which; the reference raises there).
   dist as for arp_pairs_fetch_compact.  ARP_E_CAPACITY sets *n_pairs and *bits_j. */
int  arp_pairs_fetch_packed(arp_ctx* ctx, uint32_t* row_off, uint32_t* lo32, uint8_t* hi8, uint64_t cap, float* dist,
                            uint64_t* n_pairs, int32_t* bits_j, uint32_t* n_faults);
/* host only.  In a batch (several structures in one upload) the j of the words are LOCAL to the structure of row i -- no
   pair spans two structures -- so bits_j is that of the largest structure.  struct_off [n_structures + 1] (NULL: one
   structure) gives the structures' first atoms: the records come out with the run's global i and j.  To unpack ONE
   structure of a batch pass its rows (row_off rebased to start at 0), words and feat with struct_off = NULL: i and j
   are then local to it.  threads > 1: that many host threads share the rows (1.25 M records: about 11 ms with one). */
int  arp_pairs_unpack_packed(const uint32_t* row_off, const uint32_t* lo32, const uint8_t* hi8, const float* dist,
                             int32_t n_atoms, int32_t bits_j, const uint32_t* feat, const int32_t* struct_off, int32_t n_structures,
                             arp_pair* dst, uint64_t cap, int32_t threads);

/* arp_pairs_fetch_packed split for pipelining: _async enqueues the sorted packed view and its copies BEHIND a run that
   has not been waited for (arp_pairs_run_async) -- the record count is read on the device and the first
   min(cap, expect) words are copied (expect: the caller's guess, 0 = cap) -- and returns at once; _wait is the single
   host wait of the step: it repeats an overflowed run, fetches what `expect` missed, and reports ARP_E_CAPACITY (with
   *n_pairs set) when cap was too small; the views stay valid for a second, plain arp_pairs_fetch_packed.  The
   destination buffers must stay untouched until _wait returns, and row_off must hold n_atoms + 2 entries here (the last
   one is scratch: the fault counter travels behind the offsets).  Replaces nothing in the reference (no overlap there). */
int  arp_pairs_fetch_packed_async(arp_ctx* ctx, uint32_t* row_off, uint32_t* lo32, uint8_t* hi8, uint64_t cap, float* dist, uint64_t expect);
int  arp_pairs_fetch_packed_wait(arp_ctx* ctx, uint64_t* n_pairs, int32_t* bits_j, uint32_t* n_faults);
/* host only, no context: compact view -> 16-byte records (dist NULL: distance 0) */
int  arp_pairs_unpack(const uint32_t* row_off, const arp_pair_c* rec, const float* dist, int32_t n_atoms,
                      arp_pair* dst, uint64_t cap);

/* ---- plane terms -----------------------------------------------------------
 * arp_ring_ring_run   replaces __calculate_plane_plane_contacts (interactions.py:1064-1194)
 * arp_atom_ring_run   replaces __calculate_atom_plane_contacts  (interactions.py:947-1062)
 * arp_amide_amide_run replaces __calculate_group_group_contacts (interactions.py:1217-1300)
 * arp_amide_ring_run  replaces __calculate_group_plane_contacts (interactions.py:1302-1382)
 * with utils.group_angle / group_group_angle (utils.py:638-693).
 * Planes belong to the single structure last uploaded (n_structures == 1).
 * Results are returned sorted (ring-ring: by first visit in the reference's
 * double loop; the others by (a, b)).  A plane run makes the results of earlier
 * plane runs invalid (their fetch fails with ARP_E_NOT_READY): fetch before the
 * next run, or run all four at once with arp_planes_run_all.                  */
int  arp_upload_planes(arp_ctx* ctx, const arp_planes* rings, const arp_planes* amides);
int  arp_ring_ring_run(arp_ctx* ctx, uint64_t* n);
int  arp_ring_ring_fetch(arp_ctx* ctx, arp_plane_pair* dst, uint64_t cap);
int  arp_atom_ring_run(arp_ctx* ctx, uint64_t* n);
int  arp_atom_ring_fetch(arp_ctx* ctx, arp_atom_plane* dst, uint64_t cap);
int  arp_amide_amide_run(arp_ctx* ctx, uint64_t* n);
int  arp_amide_amide_fetch(arp_ctx* ctx, arp_plane_pair* dst, uint64_t cap);
int  arp_amide_ring_run(arp_ctx* ctx, uint64_t* n);
int  arp_amide_ring_fetch(arp_ctx* ctx, arp_plane_pair* dst, uint64_t cap);
/* all four terms in ONE launch sequence and one wait: n[0..3] = record counts of ring-ring, atom-ring, amide-amide,
   amide-ring; the four fetch calls then return them (atom-ring is left out when no single structure is uploaded) */
int  arp_planes_run_all(arp_ctx* ctx, uint64_t* n);

/* ---- per-atom SIFt reductions ----------------------------------
 * Segmented OR / count / last-contact reduction of the record stream of the last arp_pairs_run onto
 * the atoms, for the loop order = ascending (bgn, end) (the order of the sorted stream; the reference's
 * own order is the KD-tree traversal of Bio.PDB.NeighborSearch, which only matters for integer_sift).
 * dst[i] belongs to atom i of the uploaded list; cap >= n_atoms.                                  */
int  arp_atom_sifts_run(arp_ctx* ctx);
int  arp_atom_sifts_fetch(arp_ctx* ctx, arp_atom_sift* dst, uint64_t cap);

/* ---- ring -> residue assignment -----------------------------------
 * replaces the search of _assign_aromatic_rings_to_residues (interactions.py:1453-1492): for every ring
 * centroid (double[n_rings][3], OBRing.findCenterAndNormal) the closest atom of xyz (float[n_atoms][3], the
 * structure's s_atoms) within `radius` as NeighborSearch.search tests it (double, d2 <= r*r), distance =
 * np.linalg.norm(atom.coord - centroid) in float64; atom_out[r] = -1 when no atom is that close (the reference
 * then sets ring['residue'] = None).  Ties go to the lowest atom index.  Host pointers; synchronous.          */
int  arp_ring_nearest_atom(arp_ctx* ctx, const float* xyz, int32_t n_atoms, const double* centers, int32_t n_rings,
                           double radius, int32_t* atom_out, double* dist_out);

/* ---- contact JSON ------------------------------------------------
 * Host-side emitter of the atom-atom entries of InteractionComplex.get_contacts (interactions.py:172-196)
 * exactly as json.dump(contacts, fp, indent=4, sort_keys=True) writes them (process_protein_cli.py:187-188):
 * entries joined by ",\n", WITHOUT the enclosing "[\n" ... "\n]" (the plane entries follow in the same list).
 * frag[a] / frag_len[a]: the rendered 'bgn' / 'end' object of atom a (utils.make_pymol_json, utils.py:530-564,
 * + label_comp_type) as json.dumps leaves it at nesting depth 2, not NUL-terminated.  distance is
 * float.__repr__(round(np.float64(dist), 2)).  rec: host memory (what arp_pairs_fetch wrote).
 * threads: host threads to use (>= 1).  No device, no context.                                      */
int  arp_pairs_json_size(const arp_pair* rec, uint64_t n, int32_t n_atoms, const uint32_t* frag_len, int threads,
                         uint64_t* bytes);
int  arp_pairs_json_write(const arp_pair* rec, uint64_t n, int32_t n_atoms, const char* const* frag,
                          const uint32_t* frag_len, int threads, char* dst, uint64_t cap, uint64_t* written);

/* ---- binding-site expansion -----------------------------------
 * replaces the search_all(6.0) loop of _make_selection (interactions.py:1420-1424):
 * flag[i] = 1 iff atom i is selected or lies within `radius` of a selected atom
 * (ARP_F_IN_SELECTION in feat), distances tested as Bio.PDB.kdtrees does.      */
int  arp_flag_within(arp_ctx* ctx, double radius, uint8_t* flags_out, uint64_t cap);

/* ---- misc ------------------------------------------------------------------ */
int  arp_sync(arp_ctx* ctx);
int  arp_get_stats(arp_ctx* ctx, arp_stats* out);
/* bench hook: repeats the whole atom-atom job (grid build + pair kernel) `iters` times on the
   resident inputs, each iteration bracketed by CUDA events on the context's stream; flush_l2 != 0
   overwrites a buffer larger than L2 between iterations, outside the brackets.  *ms_per_iter is the
   mean whole-job time; arp_get_stats then holds the mean grid-build / pair-kernel split. */
int  arp_timing_iters(arp_ctx* ctx, int iters, int flush_l2, float* ms_per_iter);
uint64_t arp_launch_count(arp_ctx* ctx);           /* kernels this context has launched so far */
/* bench hook: `iters` pinned cudaMemcpyAsync of h2d_bytes host -> device and of d2h_bytes device -> host, the two
   directions concurrently on two streams; ms[0] = CUDA-event time of the H2D copies, ms[1] of the D2H copies.
   What PCIe gives this GPU for the transfer sizes of one end-to-end step (no kernels involved). */
int  arp_memcpy_probe(arp_ctx* ctx, uint64_t h2d_bytes, uint64_t d2h_bytes, int iters, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* ARPEGGIO_CUDA_H */
