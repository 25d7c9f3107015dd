/*
 * arp_sifts.cu -- per-atom SIFt reductions.
 *
 * Replaces the side effects of the pair loop on the atoms:
 *   utils.update_atom_integer_sift / update_atom_sift / update_atom_fsift     utils.py:182-242
 *   (called at interactions.py:924-934) and the hbond / polar counters        interactions.py:822-852
 *
 * The reference runs them once per contact, in loop order.  Everything but integer_sift is an
 * order-free OR / count.  integer_sift is ASSIGNED at every contact as `sift before the contact + SIFt`
 * (utils.py:233), so it ends as  OR(all earlier contacts of the category) + SIFt(last contact):
 * with  seen = OR of all contacts,  dup = bits seen in at least two contacts,  last = SIFt of the
 * last contact:   integer[b] = last[b] + (last[b] ? dup[b] : seen[b]).
 * Loop order here = position in the (bgn, end)-sorted record stream.
 *
 *   k_sift_accumulate  one thread per record: atomicOr (its return value feeds dup), atomicMax of
 *                      (position << 16 | SIFt) for the last contact, atomicAdd for the counters,
 *                      on both atoms, for category 0 and the record's own category
 *   k_sift_finalize    one thread per atom: arp_atom_sift
 */
#include "arp_ctx.cuh"

struct SiftAcc {                       /* 96 bytes per atom, zero-initialised */
    unsigned int       seen[4];
    unsigned int       dup[4];
    unsigned long long last[4];        /* (position + 1) << 16 | SIFt */
    unsigned int       hb[4];
    unsigned int       pl[4];
};

/* category of an entity class: 'INTER' -> 1, 'INTRA_*' -> 2, '*WATER*' -> 3 (utils.py:189-199) */
__host__ __device__ static inline int sift_category(uint32_t cls)
{
    return cls == ARP_CLASS_INTER ? 1 : (cls == ARP_CLASS_INTRA_NON_SELECTION || cls == ARP_CLASS_INTRA_SELECTION) ? 2 : 3;
}

__global__ void __launch_bounds__(256) k_sift_accumulate(const arp_pair* __restrict__ rec, unsigned long long n,
                                                         SiftAcc* __restrict__ acc)
{
    const unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int4 v = reinterpret_cast<const int4*>(rec)[r];
    const uint32_t m = (uint32_t)v.z & 0x7fffu;
    const int cat = sift_category(((uint32_t)v.z >> ARP_CLASS_SHIFT) & 7u);
    const unsigned long long key = ((r + 1) << 16) | m;
    const bool hb = m >> ARP_SIFT_HBOND & 1u, pl = m >> ARP_SIFT_POLAR & 1u;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        SiftAcc* a = acc + (side ? v.y : v.x);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int c = t ? cat : 0;
            const unsigned old = atomicOr(&a->seen[c], m);
            if (old & m) atomicOr(&a->dup[c], old & m);
            atomicMax(&a->last[c], key);
            if (hb) atomicAdd(&a->hb[c], 1u);
            if (pl) atomicAdd(&a->pl[c], 1u);
        }
    }
}

__global__ void __launch_bounds__(256) k_sift_finalize(int N, const SiftAcc* __restrict__ acc, arp_atom_sift* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const SiftAcc a = acc[i];
    arp_atom_sift o;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t seen = a.seen[c], dup = a.dup[c], last = (uint32_t)(a.last[c] & 0x7fffu);
        uint32_t integer = 0;
#pragma unroll
        for (int b = 0; b < ARP_SIFT_NBITS; ++b) {
            const uint32_t l = last >> b & 1u;
            const uint32_t before = l ? (dup >> b & 1u) : (seen >> b & 1u);
            integer |= (l + before) << (2 * b);
        }
        o.sift[c] = (uint16_t)seen;
        o.integer_sift[c] = integer;
        o.hbonds[c] = a.hb[c];
        o.polars[c] = a.pl[c];
    }
    out[i] = o;
}

int arp_atom_sifts_enqueue(arp_ctx* c)
{
    const size_t N = (size_t)c->N;
    ARP_TRY(arp_pairs_sorted_build(c, 0));
    ARP_TRY(dbuf_reserve(c, c->sift_acc, sizeof(SiftAcc) * N));
    ARP_TRY(dbuf_reserve(c, c->sift_out, sizeof(arp_atom_sift) * N));
    if (N == 0) return ARP_OK;
    ARP_CUDA(c, cudaMemsetAsync(c->sift_acc.p, 0, sizeof(SiftAcc) * N, c->stream));
    const unsigned long long n = c->n_pairs;
    if (n) {
        k_sift_accumulate<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->sort_out.as<arp_pair>(), n, c->sift_acc.as<SiftAcc>());
        ARP_LAUNCHED(c);
    }
    k_sift_finalize<<<(unsigned)((N + 255) / 256), 256, 0, c->stream>>>((int)N, c->sift_acc.as<SiftAcc>(), c->sift_out.as<arp_atom_sift>());
    ARP_LAUNCHED(c);
    return ARP_OK;
}
