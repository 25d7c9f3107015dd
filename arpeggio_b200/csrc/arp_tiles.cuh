/*
 * arp_tiles.cuh -- neighbour search and per-pair classification in ONE kernel, from shared-memory tiles.
 * Included by arp_pairs.cu (uses its launch helpers and the bulk-store helpers).
 *
 * Replaces, together with the grid build in front of it and k_hscan behind it:
 *   Bio.PDB.NeighborSearch(selection_plus).search_all(cutoff)   interactions.py:1442, :707
 *   InteractionComplex._calculate_atom_contacts                 interactions.py:693-936
 *
 * Every WARP works for itself (no block barrier after the prologue).  Work unit = one home cell, handed out by
 * ticket counters two cells ahead.  The home cell and its 13 forward neighbours are five contiguous runs of the
 * cell-sorted 32-byte atom records (x, y, z, original index | packed feature word, residue, chain links):
 *
 *   stage    the runs are copied global -> the warp's tile in shared memory by the TMA engine (cp.async.bulk,
 *            completion on the warp's mbarrier); two tiles per warp, the next cell's copy is issued as soon as
 *            the search of the current cell is over and lands while its hits are classified
 *   search   lane = candidate (registers, hydrogens parked at +inf), home atom broadcast from the tile, float32
 *            FMA d^2 against the upper edge of the band around r^2, ballot compaction of the hits into the
 *            warp's queue as 16-bit tile positions
 *   classify 32 hits per round, one lane per pair, operands from the tile: the exact double test of
 *            Bio.PDB.kdtrees inside the band, orientation (atom_bgn = lower list index), the reference's
 *            `continue` filters, float32 distance, proximity bit, bit-parallel feature rules
 *            (rule_classify_core); records are compacted into the warp's staging tile, which leaves with
 *            cp.async.bulk shared -> global; pairs that need a hydrogen scan / halogen / xbond predicate append
 *            a work item (original atom indices) for k_hscan.  Hits that do not fill a round stay queued and
 *            are classified together with the next cell's (their tile is still in place).
 *
 * There is no candidate list in global memory and no hand-off between kernels.  A cell with more than TW_WCAP
 * candidates or TW_HC home atoms (a locally very dense structure) is processed as several jobs: home chunks of
 * TW_HC atoms against windows of TW_WCAP candidates, same code.
 */
#ifndef ARP_TILES_CUH
#define ARP_TILES_CUH

#define TW_WARPS    4
#define TW_THREADS  (TW_WARPS * 32)
#define TW_WCAP     112                     /* candidates staged per job (4 slots of 32 lanes hold up to 128) */
#define TW_HC       16                      /* home atoms per job */
#define TW_TILE     (TW_WCAP + TW_HC)       /* records per tile; [TW_WCAP, TW_TILE): the home chunk of a job whose window does not contain it */
#define TW_SLOTS    4
#define TW_QCAP     256                     /* hit queue per warp: < DRAIN before a home atom, + <= WCAP hits */
#define TW_DRAIN    128
#define TW_RB       96                      /* record staging tile (records); it leaves once it holds >= TW_FLUSH */
#define TW_FLUSH    64
#define TW_ITEMS    192                     /* work items staged with the tile; a round adds at most 3 per lane (is_hbond scan +
                                               (is_weak_hbond scan | halogen) + xbond), so the tile also leaves when fewer than 96 are free */
#define TW_NC       ARP_CLS_COUNTERS        /* ticket counters */
/* per-warp shared memory (bytes) */
#define TW_O_TILE   0
#define TW_O_REC    (2 * TW_TILE * 32)
#define TW_O_QUEUE  (TW_O_REC + 2 * TW_RB * 16)
#define TW_O_ITEMS  (TW_O_QUEUE + TW_QCAP * 2)
#define TW_O_SURV   (TW_O_ITEMS + TW_ITEMS * 4)
#define TW_O_MBAR   (TW_O_SURV + TW_RB * 8)
#define TW_WARP_BYTES (TW_O_MBAR + 16)
#define TW_SMEM     (TW_WARPS * TW_WARP_BYTES)
#ifndef TW_MINB
#define TW_MINB     4
#endif

struct TileArgs {
    const uint4*  arec;                     /* cell-sorted atom records, 2 x 16 bytes per atom */
    const int2*   runtab;                   /* per cell: 5 x (first position, length) of its runs, (home atoms, r2_hi bits) */
    RunMeta*      meta;
    arp_pair*     out;
    unsigned long long cap;                 /* capacity of the record stream */
    uint4*        work;                     /* deferred predicates: (donor, acceptor | halogen, record index, kind); ORIGINAL atom indices */
    unsigned long long work_cap;
    const unsigned long long* h_reach;      /* upload generation << 32 | float bits of the longest donor-hydrogen distance */
    unsigned      h_gen;
    double        r2;
    int           include_seq_adjacent;
    ArpSide       side;
};

/* ---- mbarrier / bulk-copy primitives (PTX) ---- */
__device__ __forceinline__ void mbar_init(uint32_t bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, unsigned bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int2 lds_i2(uint32_t a)
{
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

/* the candidate chunk (registers) against the home atoms whose bits are left in hmask; stops early once the queue
   holds TW_DRAIN hits */
template <int NS>
__device__ __forceinline__ unsigned tile_search(uint32_t home_addr, uint32_t q_addr, unsigned& qcount, unsigned hmask,
                                                const float (&cxs)[TW_SLOTS], const float (&cys)[TW_SLOTS],
                                                const float (&czs)[TW_SLOTS], int tk, unsigned ebase, float r2_hi,
                                                unsigned lt_mask)
{
    while (hmask && qcount < TW_DRAIN) {
        const int hl = __ffs(hmask) - 1;
        hmask &= hmask - 1u;
        const float4 hp = lds_f4(home_addr + 32u * (unsigned)hl);      /* every lane reads the same address: one broadcast */
        const int hk = hl - tk;            /* candidate at tile position lane + 32 sl is later in the cell's list than the home atom iff 32 sl > hk */
        const unsigned eh = ebase + ((unsigned)hl << 7);
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            const float ddx = hp.x - cxs[sl], ddy = hp.y - cys[sl], ddz = hp.z - czs[sl];
            const float d2 = __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, __fmul_rn(ddx, ddx)));
            const bool hit = (d2 <= r2_hi) && (32 * sl > hk);          /* padding lanes and hydrogens sit at +inf */
            const unsigned m = __ballot_sync(FULL, hit);
            if (hit) {
                const uint32_t addr = q_addr + 2u * (qcount + __popc(m & lt_mask));
                asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)(eh + 32u * sl)) : "memory");
            }
            qcount += __popc(m);
        }
    }
    return hmask;
}

__global__ void __launch_bounds__(TW_THREADS, TW_MINB) k_tiles(TileArgs A, ArpRuleParams P)
{
    extern __shared__ __align__(128) unsigned char s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbase = smem_u32(s_dyn) + (uint32_t)warp * TW_WARP_BYTES;
    unsigned char* const wptr = s_dyn + (size_t)warp * TW_WARP_BYTES;
    int4* const rec0 = reinterpret_cast<int4*>(wptr + TW_O_REC);                        /* 2 staging tiles */
    const uint32_t q_addr = wbase + TW_O_QUEUE;
    uint32_t* const items = reinterpret_cast<uint32_t*>(wptr + TW_O_ITEMS);
    uint2* const surv = reinterpret_cast<uint2*>(wptr + TW_O_SURV);
    const uint32_t mbar0 = wbase + TW_O_MBAR;
    const unsigned lt_mask = (1u << lane) - 1u;

    __shared__ float4 s_radtab[CLS_TAB_K * CLS_TAB_K];
    __shared__ double s_vdw[CLS_TAB_K];
    __shared__ float s_hlim[CLS_TAB_K];

    /* small radius tables live in shared memory (uploaded data, not written by the run) */
    A.side.hlim = nullptr;
    if (A.side.K <= CLS_TAB_K) {
        for (int k = threadIdx.x; k < A.side.K * A.side.K; k += TW_THREADS) s_radtab[k] = A.side.radtab[k];
        for (int k = threadIdx.x; k < A.side.K; k += TW_THREADS) {
            s_vdw[k] = A.side.vdw[k];
            /* donor farther than this from an acceptor of class k: none of its hydrogens can be within
               h_vdw + vdw_k + comp (utils.py:89, :149) of it; generous rounding margin on top */
            const unsigned long long hr = *A.h_reach;     /* an older generation: this upload has no hydrogens at all */
            const float reach = (unsigned)(hr >> 32) == A.h_gen ? __uint_as_float((unsigned)hr) : __int_as_float(0xff800000);
            const double lim = (double)reach + P.h_vdw + A.side.vdw[k] + P.vdw_comp;
            s_hlim[k] = lim == lim ? __double2float_ru(lim) * 1.000001f + 2e-3f : __int_as_float(0x7f800000);
        }
    }
    if (lane == 0) {
        mbar_init(mbar0, 1);
        mbar_init(mbar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (A.side.K <= CLS_TAB_K) { A.side.radtab = s_radtab; A.side.vdw = s_vdw; A.side.hlim = s_hlim; }

    pdl_wait();                                 /* the grid build has completed */
    pdl_trigger();
    PROF_STAMP(1, 0);
    const int n_cells = (int)A.meta->n_cells;
    /* queued hits outlive their cell (and, in a batch, their structure): one conservative lower band edge for all */
    const float r2_lo = A.meta->r2_lo_inv == 0x7f800000u ? -1.0f : __uint_as_float(0x7f800000u - A.meta->r2_lo_inv);

    /* ---- tickets: cells are dealt in TW_NC residue classes, counter c serves the cells and the blocks congruent
            to c; a warp's first cell is static (its rank among the warps of its class) ---- */
    const unsigned nc = min((unsigned)TW_NC, gridDim.x);
    const unsigned cls = blockIdx.x % nc;
    unsigned* const ticket_ctr = &A.meta->ticket_srch[cls].v;
    const unsigned class_warps = ((gridDim.x - cls + nc - 1) / nc) * TW_WARPS;

    /* ---- per-warp state ---- */
    unsigned qcount = 0, n_old = 0;             /* queued hits; how many of them (at the front) point into the other tile */
    unsigned nrec = 0, n_items = 0, n_rare = 0;
    int rbuf = 0;                               /* record staging tile in use */
    unsigned par = 0;                           /* phase parity of the two mbarriers (bit b) */
    unsigned long long ncand = 0;
    unsigned nonempty = 0;

    /* job = (cell, home chunk [hs, he), window [ws, we) of the cell's candidate list); the current job lives in
       tile `cb`, the next one is prefetched into tile cb ^ 1.  Run tables in registers, lane r < 5 = run r:
       (first cell-sorted position, exclusive prefix of the lengths, length). */
    int cur_valid = 0, cb = 1;
    int nh = 0, total = 0, hs = 0, ws = 0;      /* current cell / job */
    float r2_hi = 0.f;
    int cu_beg = 0, cu_pre = 0, cu_len = 0;     /* run table of the current cell */
    int nx_beg = 0, nx_pre = 0, nx_len = 0;     /* ... of the next job's cell */
    /* iterator over the warp's cells: c_nx = next cell (its table load is in flight in tab_nx), t_pend = ticket of
       the cell after that (atomic in flight) */
    int c_nx = (int)((((blockIdx.x / nc) * TW_WARPS + warp)) * nc + cls);
    int2 tab_nx = make_int2(0, 0);
    if (lane < 6 && c_nx < n_cells) tab_nx = A.runtab[6 * (size_t)c_nx + lane];
    unsigned t_pend = 0;
    if (lane == 0) t_pend = atomicAdd(ticket_ctr, 1u);
    int nx_valid = 0, nx_nh = 0, nx_total = 0, nx_hs = 0, nx_ws = 0;
    float nx_hi = 0.f;
    bool have_nx = false, hook = false;
    /* search state of the current job */
    bool sdone = true;
    unsigned hmask = 0;
    int hlen = 0, wlen = 0;

    for (;;) {
        const uint32_t tile_cur = wbase + TW_O_TILE + (uint32_t)cb * (TW_TILE * 32);
        const bool idle = !cur_valid || sdone;  /* nothing (left) to search in the current job */
        /* ---- S: search the current job until it is finished or the queue holds TW_DRAIN hits ---- */
        if (!idle) {
            const uint32_t home_addr = tile_cur + (ws == hs ? 0u : (uint32_t)TW_WCAP * 32u);
            float cxs[TW_SLOTS], cys[TW_SLOTS], czs[TW_SLOTS];
#pragma unroll
            for (int sl = 0; sl < TW_SLOTS; ++sl) {
                const int t = sl * 32 + lane;
                cxs[sl] = cys[sl] = czs[sl] = __int_as_float(0x7f800000);     /* +inf: never within any cutoff */
                if (t < wlen) {
                    const float4 p = lds_f4(tile_cur + 32u * (unsigned)t);
                    const uint32_t w = lds_u32(tile_cur + 32u * (unsigned)t + 16u);
                    if (!(w & ARPK_ELEM_H)) { cxs[sl] = p.x; cys[sl] = p.y; czs[sl] = p.z; }      /* interactions.py:712-713 */
                }
            }
            const int tk = ws - hs + lane;      /* list position of the lane's first candidate relative to the home chunk */
            const unsigned ebase = ((unsigned)cb << 15) | ((ws == hs ? 0u : (unsigned)TW_WCAP) << 7) | (unsigned)lane;
            if (wlen > 96)      hmask = tile_search<4>(home_addr, q_addr, qcount, hmask, cxs, cys, czs, tk, ebase, r2_hi, lt_mask);
            else if (wlen > 64) hmask = tile_search<3>(home_addr, q_addr, qcount, hmask, cxs, cys, czs, tk, ebase, r2_hi, lt_mask);
            else if (wlen > 32) hmask = tile_search<2>(home_addr, q_addr, qcount, hmask, cxs, cys, czs, tk, ebase, r2_hi, lt_mask);
            else                hmask = tile_search<1>(home_addr, q_addr, qcount, hmask, cxs, cys, czs, tk, ebase, r2_hi, lt_mask);
            sdone = hmask == 0;
            __syncwarp();
        }
        const bool over = !cur_valid || sdone;
        /* ---- the next job: another window / home chunk of this cell (dense cells only), or the next cell ---- */
        if (over && !have_nx) {
            have_nx = true;
            nx_valid = 0;
            if (cur_valid && (total > TW_WCAP || nh > TW_HC)) {
                nx_nh = nh; nx_total = total; nx_hi = r2_hi;
                nx_beg = cu_beg; nx_pre = cu_pre; nx_len = cu_len;
                nx_hs = hs; nx_ws = ws + TW_WCAP;
                if (nx_ws >= total) { nx_hs = hs + TW_HC; nx_ws = nx_hs; }
                nx_valid = nx_hs < nh;
            }
            while (!nx_valid && c_nx < n_cells) {
                /* the table of cell c_nx has arrived: lanes 0..4 hold the runs, lane 5 (home atoms, band) */
                nx_nh = __shfl_sync(FULL, tab_nx.x, 5);
                if (nx_nh > 0) {
                    nx_hi = __int_as_float(__shfl_sync(FULL, tab_nx.y, 5));
                    nx_beg = tab_nx.x;
                    nx_len = lane < 5 ? tab_nx.y : 0;
                    int incl = nx_len;
#pragma unroll
                    for (int off = 1; off < 8; off <<= 1) {
                        const int v = __shfl_up_sync(FULL, incl, off);
                        if (lane >= off) incl += v;
                    }
                    nx_pre = incl - nx_len;
                    nx_total = __shfl_sync(FULL, incl, 4);
                    nx_valid = 1; nx_hs = 0; nx_ws = 0;
                    ++nonempty;
                    ncand += (unsigned long long)((long long)nx_nh * nx_total - (long long)nx_nh * (nx_nh + 1) / 2);
                }
                /* advance the cell iterator: the pending ticket has arrived; request the one after */
                const unsigned tk_now = __shfl_sync(FULL, t_pend, 0);
                c_nx = (int)((class_warps + tk_now) * nc + cls);
                tab_nx = make_int2(0, 0);
                if (lane < 6 && c_nx < n_cells) tab_nx = A.runtab[6 * (size_t)c_nx + lane];
                if (lane == 0) t_pend = atomicAdd(ticket_ctr, 1u);
            }
            hook = nx_valid != 0;
        }
        /* ---- C: classify.  Full rounds only while the search goes on; once the job is finished, the hits that
                still point into the other tile are retired first (a partial round if need be), then the next job's
                copy is issued into that tile, then the remaining full rounds run under the copy.  When there is no
                next job everything is drained. ---- */
        const bool finishing = over && !nx_valid;
        unsigned nround = qcount / 32u;
        if (finishing) nround = (qcount + 31u) / 32u;
        else if (over && n_old > 0 && nround == 0) nround = 1;
        unsigned consumed = 0;
        for (unsigned r = 0;; ++r) {
            const bool last = r == nround;
            if (hook && n_old == 0) {
                /* ---- issue the next job's copies into tile cb ^ 1 ---- */
                hook = false;
                const int nb = cb ^ 1;
                const uint32_t tile_nx = wbase + TW_O_TILE + (uint32_t)nb * (TW_TILE * 32);
                const uint32_t bar = mbar0 + 8u * (unsigned)nb;
                const int nx_he = min(nx_nh, nx_hs + TW_HC), nx_we = min(nx_total, nx_ws + TW_WCAP);
                const int a = max(nx_ws, nx_pre), b = min(nx_we, nx_pre + nx_len);
                const bool own_home = nx_ws != nx_hs;                                 /* the window does not contain the home chunk */
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive_expect_tx(bar, (unsigned)((nx_we - nx_ws) + (own_home ? nx_he - nx_hs : 0)) * 32u);
                }
                __syncwarp();
                if (lane < 5 && b > a)
                    bulk_load(tile_nx + 32u * (unsigned)(a - nx_ws), A.arec + 2 * (size_t)(nx_beg + (a - nx_pre)), (unsigned)(b - a) * 32u, bar);
                if (own_home) {
                    const int beg0 = __shfl_sync(FULL, nx_beg, 0);
                    if (lane == 5) bulk_load(tile_nx + 32u * TW_WCAP, A.arec + 2 * (size_t)(beg0 + nx_hs), (unsigned)(nx_he - nx_hs) * 32u, bar);
                }
            }
            if (nrec >= TW_FLUSH || n_items + n_rare > TW_ITEMS - 96 || (last && finishing && nrec > 0)) {
                /* ---- the staging tile leaves through the TMA engine; its work items go to the global work list ---- */
                int4* const rec = rec0 + rbuf * TW_RB;
                __syncwarp();
                unsigned long long o = 0, ow = 0, owr = 0;
                if (lane == 0) {
                    o = atomicAdd(&A.meta->n_pairs, (unsigned long long)nrec);
                    if (n_items) ow = atomicAdd(&A.meta->n_work, (unsigned long long)n_items);
                    if (n_rare) owr = atomicAdd(&A.meta->n_work_rare, (unsigned long long)n_rare);
                    if (o + nrec <= A.cap) bulk_store_tile(A.out + o, rec, nrec * (uint32_t)sizeof(arp_pair));
                    bulk_store_wait_read_1();        /* the other tile's store has read it: free for the next records */
                }
                o = __shfl_sync(FULL, o, 0);
                ow = __shfl_sync(FULL, ow, 0);
                owr = __shfl_sync(FULL, owr, 0);
                /* The hydrogen scans fill the work list from its front, the rare predicates (halogen weak hbond,
                   xbond) from its back (they get chunks of their own in k_hscan). */
                for (unsigned w = lane; w < n_items + n_rare; w += 32) {
                    const bool is_rare = w >= n_items;
                    const uint32_t it = is_rare ? items[TW_ITEMS - 1 - (w - n_items)] : items[w];
                    const unsigned slot = (it >> 4) & 0xfffu;
                    uint2 e = surv[slot];
                    if (it & 8u) { const unsigned t = e.x; e.x = e.y; e.y = t; }   /* e.x = donor, e.y = acceptor / halogen */
                    const unsigned long long wpos = is_rare ? owr + (w - n_items) : ow + w;
                    if (wpos < A.work_cap)      /* a record beyond the stream's capacity (the host repeats the run): an item that asks for nothing */
                        A.work[is_rare ? A.work_cap - 1 - wpos : wpos] =
                            o + slot < A.cap ? make_uint4(e.x, e.y, (uint32_t)(o + slot), (it & 7u) | ((it >> 16) << 8)) : make_uint4(0u, 0u, 0u, 0u);
                }
                __syncwarp();
                rbuf ^= 1; nrec = 0; n_items = 0; n_rare = 0;
            }
            if (last) break;
            /* ---- one round: 32 queued hits ---- */
            int4* const rec = rec0 + rbuf * TW_RB;
            const unsigned qi = 32u * r + lane;
            bool keep = false;
            float4 pa, pb;
            uint4 ab, ae;
            if (qi < qcount) {
                unsigned short e16;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(e16) : "r"(q_addr + 2u * qi));
                const unsigned e = e16;
                const uint32_t tile = wbase + TW_O_TILE + (e >> 15) * (TW_TILE * 32);
                uint32_t aa = tile + 32u * ((e >> 7) & 0xffu), bb = tile + 32u * (e & 0x7fu);
                /* atom_bgn = lower list index: the operands are fetched in that order */
                if ((int)lds_u32(bb + 12u) < (int)lds_u32(aa + 12u)) { const uint32_t t = aa; aa = bb; bb = t; }
                pa = lds_f4(aa); pb = lds_f4(bb);
                ab = lds_u4(aa + 16u); ae = lds_u4(bb + 16u);
                const float ddx = pa.x - pb.x, ddy = pa.y - pb.y, ddz = pa.z - pb.z;
                const float d2 = __fmaf_rn(ddz, ddz, __fmaf_rn(ddy, ddy, __fmul_rn(ddx, ddx)));   /* the search's value up to the sign of the differences */
                keep = true;
                if (!(d2 <= r2_lo)) keep = kd_within(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, A.r2);
                keep = keep && rule_pair_survives(ab.x, (int)ab.y, (int)ab.z, (int)ab.w, ae.x, (int)ae.y, (int)ae.z, (int)ae.w,
                                                  A.include_seq_adjacent);
            }
            /* survivors keep their lane for the rules; their records are compacted into the staging tile */
            const unsigned mk = __ballot_sync(FULL, keep);
            const unsigned slot = nrec + __popc(mk & lt_mask);
            nrec += __popc(mk);
            uint32_t work = 0;
            if (keep) {
                const int ib = __float_as_int(pa.w), ie = __float_as_int(pb.w);
                uint32_t mask; float dist;
                rule_classify_core(A.side, P, ib, ie, pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, ab.x, ae.x, &mask, &dist, &work);
                rec[slot] = make_int4(ib, ie, (int)mask, __float_as_int(dist));
                if (work) surv[slot] = make_uint2((unsigned)ib, (unsigned)ie);     /* k_hscan finds donor / acceptor through this */
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
                        items[TW_ITEMS - 1 - (n_rare + __popc(m & lt_mask))] = ((rare & ARP_WORK_HAL1) ? it1 : it0) | CLS_KIND_HAL;
                    n_rare += __popc(m);
                    m = __ballot_sync(FULL, (rare & (ARP_WORK_XB0 | ARP_WORK_XB1)) != 0);
                    if (rare & (ARP_WORK_XB0 | ARP_WORK_XB1))
                        items[TW_ITEMS - 1 - (n_rare + __popc(m & lt_mask))] = ((rare & ARP_WORK_XB1) ? it1 : it0) | CLS_KIND_XBOND;
                    n_rare += __popc(m);
                }
            }
            const unsigned took = min(32u, qcount - 32u * r);
            consumed += took;
            n_old = n_old > took ? n_old - took : 0u;
        }
        /* the hits of an incomplete round move to the front of the queue */
        if (consumed) {
            const unsigned rem = qcount - consumed;
            unsigned short mv = 0;
            if ((unsigned)lane < rem) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(mv) : "r"(q_addr + 2u * (consumed + lane)));
            __syncwarp();
            if ((unsigned)lane < rem) asm volatile("st.shared.u16 [%0], %1;" :: "r"(q_addr + 2u * lane), "h"(mv) : "memory");
            __syncwarp();
            qcount = rem;
        }
        if (finishing) break;
        if (over) {
            /* ---- switch to the next job: everything still queued points into the tile that is now the other one ---- */
            cb ^= 1;
            nh = nx_nh; total = nx_total; hs = nx_hs; ws = nx_ws; r2_hi = nx_hi;
            cu_beg = nx_beg; cu_pre = nx_pre; cu_len = nx_len;
            cur_valid = 1; have_nx = false; sdone = false;
            n_old = qcount;
            hlen = min(nh, hs + TW_HC) - hs;
            wlen = min(total, ws + TW_WCAP) - ws;
            mbar_wait(mbar0 + 8u * (unsigned)cb, (par >> cb) & 1u);
            par ^= 1u << cb;
            /* home atoms of the chunk that are not hydrogens (interactions.py:712-713) */
            const uint32_t home_addr = (wbase + TW_O_TILE + (uint32_t)cb * (TW_TILE * 32)) + (ws == hs ? 0u : (uint32_t)TW_WCAP * 32u);
            uint32_t hw = ARPK_ELEM_H;
            if (lane < hlen) hw = lds_u32(home_addr + 32u * (unsigned)lane + 16u);
            hmask = __ballot_sync(FULL, !(hw & ARPK_ELEM_H));
            if (hmask == 0) sdone = true;
        }
    }
    if (lane == 0) {
        bulk_store_wait_read_all();             /* shared memory must outlive the copies */
        if (ncand) atomicAdd(&A.meta->n_candidates, ncand);
        if (nonempty) atomicAdd(&A.meta->n_cells_nonempty, nonempty);
    }
#ifdef PAIR_PROFILE
    __syncthreads();
    PROF_STAMP(1, 1);
#endif
}

#endif /* ARP_TILES_CUH */
