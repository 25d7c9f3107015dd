/*
 * arp_ctx.cuh -- context object, device buffers and error plumbing shared by the
 * translation units of libarpeggio_cuda.so (not part of the public ABI).
 */
#ifndef ARP_CTX_CUH
#define ARP_CTX_CUH

#include <cuda_runtime.h>
#include <vector>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/arpeggio_cuda.h"
#include "arp_rules.cuh"

#define ARP_ERRLEN 512

/* growable device allocation */
struct DBuf {
    void*  p = nullptr;
    size_t cap = 0;
    bool   view = false;       /* p points into another allocation (the upload arena): never freed, never grown in place */
    template <class T> T* as() const { return (T*)p; }
};

/* per-structure cell grid (device) */
struct StructGeom {
    double ox, oy, oz;      /* origin = bounding-box minimum */
    double inv_w;           /* 1 / cell edge; edge >= cutoff * 1.0001 */
    int    dx, dy, dz;      /* cells per axis */
    int    ncell;           /* dx * dy * dz (0 for an empty structure) */
    int    cell_base;       /* global id of cell (0,0,0) */
    float  r2_lo, r2_hi;    /* float32 d^2 at or below r2_lo is certainly within the cutoff,
                               above r2_hi certainly outside; in between the exact double
                               test of Bio.PDB.kdtrees decides */
    /* work units of the tile kernel (k_tiles): a unit is a segment of seg_x consecutive home cells of one x-row */
    int    seg_x;           /* home cells per unit (1 .. ARP_TILE_XMAX), chosen from the structure's mean cell population */
    int    nseg;            /* segments per row = ceil(dx / seg_x) */
    int    n_units;         /* nseg * dy * dz */
    int    unit_base;       /* global id of the structure's first unit */
    int    pad;
};

/* tile kernel: most home cells per unit, atoms staged per unit (shared-memory capacity) */
#define ARP_TILE_XMAX 8
#define ARP_TILE_CAP  512

/* k_classify hands out its tiles from ARP_CLS_COUNTERS counters, each serving the tickets and the blocks of one
   residue class modulo ARP_CLS_COUNTERS (same-address atomics serialise in L2: one counter for 20 000 tiles
   costs more than the balancing gains) */
#ifndef ARP_CLS_COUNTERS
#define ARP_CLS_COUNTERS 16
#endif

/* counters the kernels leave behind (device, copied to pinned host memory after a run).
   Every counter that many warps hammer with atomics sits on its own 128-byte line. */
struct alignas(128) RunMeta {
    /* line 0: written once by the grid build, read by everybody */
    unsigned int n_cells;
    unsigned int r2_lo_inv;           /* 0x7f800000 - bits(min over structures of r2_lo); 0x7f800000 = no quick accept */
    unsigned int fault;               /* sticky device-side diagnostics */
    unsigned int n_units;             /* work units of k_tiles (all structures) */
    unsigned int pad0[28];
    /* line 1 */
    unsigned long long n_raw;         /* candidate cursor of k_search */
    unsigned int pad1[30];
    /* line 2 */
    unsigned long long n_pairs;       /* record cursor: total records the run produced */
    unsigned int pad2[30];
    /* line 3 */
    unsigned int ticket_search;       /* dynamic cell tickets of k_search */
    unsigned int pad3[31];
    /* line 4 */
    unsigned int ticket_scan;         /* dynamic tile ids of the cell scan */
    unsigned int pad4[31];
    /* line 5 */
    unsigned long long n_work;        /* work-item cursor of k_classify: hydrogen scans, from the front of the list */
    unsigned int pad6[30];
    unsigned long long n_work_rare;   /* the rare predicates (halogen weak hbond, xbond), from the back of the list */
    unsigned int pad9[30];
    /* line 6: end-of-kernel statistics */
    unsigned long long n_candidates;  /* distance tests performed */
    unsigned int n_cells_nonempty;
    unsigned int pad5[29];
    /* dynamic tile tickets of k_classify */
    struct alignas(128) { unsigned int v; unsigned int pad[31]; } ticket_cls[ARP_CLS_COUNTERS];
    /* dynamic cell tickets of k_search when it uses class counters (SEARCH_NC > 1) */
    struct alignas(128) { unsigned int v; unsigned int pad[31]; } ticket_srch[ARP_CLS_COUNTERS];
};

struct PlaneSet {
    int n = 0, is_f32 = 0;
    DBuf center, normal, res_id, flags;
};

struct PlaneResult {
    const void* ptr = nullptr;   /* the records of the last run (device) */
    const char* host = nullptr;  /* ... and their pinned host mirror when the run has already copied them */
    DBuf rec;              /* arp_plane_pair / arp_atom_plane records, sorted */
    DBuf tmp;              /* unsorted stream */
    DBuf cnt;              /* record cursor (device) */
    uint64_t n = 0;
    uint64_t cap = 0;
    int valid = 0;
};

struct arp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    char err[ARP_ERRLEN] = {0};
    int sm_count = 148;

    arp_params params;
    ArpRuleParams rp;
    int have_params = 0;

    /* uploaded atoms (original order) */
    int N = 0, Rs = 0, K = 0, S = 1, E = 0, H = 0;
    int have_atoms = 0;
    DBuf xyz, feat, res_id, rad_class, vdw, cov, res_prev, res_next, res_flags;
    DBuf bond_off, bond_nbr, h_off, h_xyz, xnbr, struct_off;
    DBuf arena;                   /* one block for all of the above when the caller's arrays are one host block */
    DBuf w_bcnt, w_hcnt, w_hfix, w_xidx, w_xnbr;   /* wire forms of arp_atoms as uploaded (bond_cnt, h_cnt, h_fix, xnbr_idx + rows) */
    unsigned* sort_fault = nullptr;   /* device: records of the packed view that carried the fault bit (= sort_off[N + 1]) */
    int max_struct_atoms = 0;         /* atoms of the largest structure of the upload: the j of the packed view are structure-local */
    int events_level = 0;             /* with_events of the last arp_pairs_enqueue */
    /* arp_pairs_fetch_packed_async -> arp_pairs_fetch_packed_wait */
    struct { uint32_t* row_off; uint32_t* lo32; uint8_t* hi8; uint64_t cap; float* dist; uint64_t copied; int pending, blind; } pk = {};
    int finish_reruns = 0;            /* runs pairs_finish had to repeat (overflow, hand-off fault) */
    unsigned long long ht_ns[8] = {0}, ht_calls[8] = {0};   /* host time inside entry points (diagnostic, ARPEGGIO_HOST_TIMING) */
    DBuf wire_sums, w_cnt_merge;  /* tile sums of the count scan; merged per-atom counts of a batch */
    DBuf batch_stage, batch_small;/* arp_upload_atoms_batch: the structures as uploaded; descriptors, struct_off, merged radius table */
    void* h_batch = nullptr;      /* pinned image of batch_small */
    size_t h_batch_cap = 0;
    cudaEvent_t ev_batch = nullptr;   /* after the H2D copy of h_batch: the image is free again */
    int has_bonds = 0, has_h = 0, has_xnbr = 0;
    uint64_t input_bytes = 0;

    /* cell grid + cell-sorted atom records */
    DBuf zero;                    /* one memset per run: RunMeta | bbox | cell_cnt | scan_state */
    size_t zero_bytes = 0, off_bbox = 0, off_cnt = 0, off_state = 0;
    size_t cell_bound = 0;        /* upper bound of the number of cells, all structures */
    DBuf geom, cell_start, cell_of, rank, pos4, att4;
    DBuf runtab;                  /* per-cell run tables of k_tiles: int2[6] per cell (arp_pairs.cu dev_runtab) */
    DBuf radtab;                  /* K x K float32 proximity thresholds */
    DBuf hreach;                  /* upload generation << 32 | float bits of the longest donor-hydrogen distance */
    unsigned upload_gen = 0;
    int radtab_valid = 0;
    std::vector<double> rad_host;     /* vdw | cov of the upload the radius-sum table was built for (bitwise comparison at the next upload) */
    int cls_smem_set = 0;
    int hscan_blocks = 0;         /* co-resident blocks of k_hscan (probed once) */
    int coop_blocks = -1;         /* co-resident blocks of k_grid_fused (0: not available, -1: not probed) */
    int use_fused_grid = 2;       /* 0: five kernels, 1: cooperative kernel through global memory, 2: + register kernel when the atoms fit */
    int reg_blocks = -1;          /* co-resident blocks of k_grid_reg (0: not available, -1: not probed) */
    int use_pdl = 1;              /* pair kernels launched with programmatic stream serialization */
    int use_tiles = 0;            /* 1: search + classify as ONE kernel on TMA-staged shared-memory tiles (k_tiles, arp_tiles.cuh;
                                     ARPEGGIO_TILES=1).  Bit-exact, but measured slower than k_search + k_classify (72 vs 58 us at
                                     100k atoms, 463 vs 374 us at 1M): 16 warps per SM against 32 (profiles/README.md, round 2) */
    int tile_x = ARP_TILE_XMAX;   /* most home cells per unit of k_tiles (ARPEGGIO_TILE_X) */
    int tiles_blocks = 0;         /* co-resident blocks of k_tiles (probed once) */
    int use_early_cls = 1;        /* k_classify consumes candidates while k_search drains (ARPEGGIO_NO_EARLY_CLASSIFY turns it off) */
    int hits_dirty = 0;           /* the candidate list may hold non-zero entries (a run without the early start) */
    RunMeta* h_meta = nullptr;    /* pinned */
    /* the zero region as the last register-grid run left it: counters, bounding boxes, cell counts and scan state
       all zero again, for exactly this allocation and layout (a run of another path, a failed enqueue or a new
       layout brings the memset back) */
    int zero_clean = 0;
    void* clean_ptr = nullptr;
    size_t clean_bytes = 0, clean_off_cnt = 0, clean_off_state = 0;

    /* output stream */
    DBuf out;                     /* arp_pair records */
    DBuf hits;                    /* uint2 candidate list of the search kernel, same capacity as out */
    DBuf work;                    /* uint4 deferred-predicate items of the classify kernel */
    uint64_t work_cap = 0;
    uint64_t out_cap = 0;         /* records */
    uint64_t n_pairs = 0;
    int pairs_valid = 0;
    DBuf sort_tmp, sort_out, sort_zero, sort_off;
    DBuf sort_c, sort_d;          /* compact view of the sorted stream: arp_pair_c[n] and float[n] (row offsets: sort_off) */
    DBuf sort_lo, sort_hi;        /* packed view: low 32 bits (+ the fault counter behind them) and, for N > 2^17, bits 32..39 of the record words */
    int sorted_valid = 0, compact_valid = 0, packed_valid = 0, sort_tmp_valid = 0;
    int run_pending = 0;          /* arp_pairs_run_async has enqueued a run that nobody has waited for yet */

    /* planes */
    PlaneSet rings, amides;
    int have_planes = 0;
    int planes_finite = 1;        /* every plane centre is finite (checked at upload): the grid path applies */
    int use_plane_grid = 1;       /* plane terms through cell grids (0: plain double loops, ARPEGGIO_NO_PLANE_GRID) */
    PlaneResult ring_ring, atom_ring, amide_amide, amide_ring;
    DBuf plane_scratch, plane_tmp, plane_rec;
    void* h_plane_rec = nullptr;  /* pinned mirror of plane_rec */
    size_t h_plane_cap = 0;       /* ... in records */
    size_t plane_cap_guess = 0;   /* records the last grid run produced (+ 50 %): capacity of the next blind emitting pass */
    int* h_plane_tot = nullptr;   /* pinned: record offsets at the term boundaries */

    /* binding-site flags */
    DBuf within;
    size_t within_flags_off = 0;  /* where arp_flag_within_run left the flags inside `within` */
    int use_within_grid = 1;      /* binding-site flags through a cell grid for large inputs (ARPEGGIO_NO_WITHIN_GRID: double loop) */

    /* ring -> nearest atom scratch (coordinates, centroids, results) */
    DBuf ring_scratch;

    /* per-atom SIFt reductions */
    DBuf sift_acc, sift_out;
    int sifts_valid = 0;

    /* timing */
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   /* [4]: between k_classify and k_hscan */
    arp_stats stats;
    DBuf flush;                   /* L2 flush buffer for arp_timing_iters */
    unsigned long long launches = 0;   /* kernels launched by this context so far */
};

static inline int arp_fail(arp_ctx* c, int code, const char* what, const char* file, int line)
{
    if (c) snprintf(c->err, ARP_ERRLEN, "%s (%s:%d)", what, file, line);
    return code;
}

#define ARP_CUDA(c, call)                                                                   \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            (void)cudaGetLastError();                                                       \
            return arp_fail((c), e_ == cudaErrorMemoryAllocation ? ARP_E_OOM : ARP_E_CUDA,  \
                            cudaGetErrorString(e_), __FILE__, __LINE__);                    \
        }                                                                                   \
    } while (0)

#define ARP_TRY(expr)                     \
    do {                                  \
        int rc_ = (expr);                 \
        if (rc_ != ARP_OK) return rc_;    \
    } while (0)

#define ARP_REQUIRE(c, cond, code, msg)                                   \
    do {                                                                  \
        if (!(cond)) return arp_fail((c), (code), (msg), __FILE__, __LINE__); \
    } while (0)

/* kernel launch check: configuration errors surface here, execution errors at the next sync */
#define ARP_LAUNCHED(c)                                                   \
    do {                                                                  \
        (c)->launches++;                                                  \
        ARP_CUDA((c), cudaGetLastError());                                \
    } while (0)

static inline int dbuf_reserve(arp_ctx* c, DBuf& b, size_t bytes)
{
    if (b.view) { b.p = nullptr; b.cap = 0; b.view = false; }
    if (bytes <= b.cap && b.p) return ARP_OK;
    if (bytes == 0) bytes = 16;
    size_t want = bytes + bytes / 8 + 256;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            b.p = nullptr;
            return arp_fail(c, ARP_E_OOM, "device allocation failed", __FILE__, __LINE__);
        }
    }
    b.cap = want;
    return ARP_OK;
}

static inline void dbuf_free(DBuf& b)
{
    if (b.p && !b.view) cudaFree(b.p);
    b.p = nullptr; b.cap = 0; b.view = false;
}

static inline int arp_bind(arp_ctx* c)
{
    ARP_CUDA(c, cudaSetDevice(c->device));
    return ARP_OK;
}

/* entry points implemented across translation units */
int  arp_pairs_prepare(arp_ctx* c);                       /* arp_pairs.cu: size the grid buffers after an upload */
/* arp_pairs.cu: memset + grid build + pair kernels; with_events 0: none, 1: ev[0] / ev[3] around the whole
   job, 2: also ev[1] between the grid build and the pair kernels, 3: also ev[2] / ev[4] between the pair kernels
   (every event between two kernels keeps them from overlapping and costs a few microseconds) */
int  arp_pairs_enqueue(arp_ctx* c, int with_events);
int  arp_pairs_sorted_build(arp_ctx* c, int view, int blind = 0);        /* arp_pairs.cu: (i, j)-ascending copy of the stream: 0 16-byte records, 1 compact, 2 packed */
int  arp_pairs_bits_j(const arp_ctx* c);                  /* bits of an atom index in the packed view */
int  arp_flag_within_run(arp_ctx* c, double radius);      /* arp_pairs.cu */
int  arp_ring_nearest_run(arp_ctx* c, const float* xyz, int n_atoms, const double* centers, int n_rings, double radius,
                          int32_t* atom_out, double* dist_out);   /* arp_rings.cu */
int  arp_atom_sifts_enqueue(arp_ctx* c);                  /* arp_sifts.cu: sorted stream -> arp_atom_sift[N] */
void arp_planes_release(arp_ctx* c);                      /* arp_planes.cu */

/* exclusive scan of n ints (n read from *n_dev + n_add when n_dev != null), single pass;
   state must hold one zeroed 64-bit word per tile of ARP_SCAN_TILE items */
#define ARP_SCAN_TILE 2048
int  arp_scan_exclusive(arp_ctx* c, const int* in, int* out, unsigned long long* state, unsigned int* ticket,
                        const unsigned int* n_dev, int n_add, size_t n_bound);

#endif /* ARP_CTX_CUH */
