/*
 * rules_emu.cpp -- TEST INFRASTRUCTURE: the device rule header (arpeggio_b200/csrc/arp_rules.cuh)
 * compiled for the host, so that the rule logic the CUDA kernels run (filters, screens in front of
 * the exact chains, direction sharing of the hydrogen loops) can be checked against the oracle in
 * the GPU-less build container.  Not linked into, loaded by or shipped with the product.
 */
#include "../../arpeggio_b200/csrc/arp_rules.cuh"

extern "C" int emu_classify(const arp_atoms* A, const arp_params* P, const int32_t* b, const int32_t* e, int64_t n,
                            arp_pair* out, uint8_t* emitted, int use_table)
{
    ArpRuleParams R;
    arp_derive_rule_params(P, &R);
    ArpSide S;
    memset(&S, 0, sizeof S);
    S.vdw = A->vdw; S.cov = A->cov; S.K = A->n_rad_classes;
    S.feat = A->feat;
    S.bond_off = A->bond_off; S.bond_nbr = A->bond_nbr; S.h_off = A->h_off; S.h_xyz = A->h_xyz; S.xnbr = A->xnbr_xyz;
    float4* tab = 0;
    /* use_table != 0: also the hydrogen-reach screen of k_classify (S.hlim), built as k_hreach + the kernel prologue do */
    float* hlim = 0;
    if (use_table) {
        float reach = -INFINITY;
        for (int a = 0; A->h_off && a < A->n_atoms; ++a)
            for (int k = A->h_off[a]; k < A->h_off[a + 1]; ++k) {
                double dx = A->h_xyz[3 * (size_t)k] - A->xyz[3 * (size_t)a], dy = A->h_xyz[3 * (size_t)k + 1] - A->xyz[3 * (size_t)a + 1],
                       dz = A->h_xyz[3 * (size_t)k + 2] - A->xyz[3 * (size_t)a + 2];
                float f = nextafterf((float)sqrt(dx * dx + dy * dy + dz * dz), INFINITY);
                if (!(f < 3.0e38f)) f = INFINITY;
                if (f > reach) reach = f;
            }
        hlim = new float[S.K];
        for (int k = 0; k < S.K; ++k) {
            double lim = (double)reach + R.h_vdw + A->vdw[k] + R.vdw_comp;
            hlim[k] = lim == lim ? nextafterf((float)lim, INFINITY) * 1.000001f + 2e-3f : INFINITY;
        }
        S.hlim = hlim;
    }
    {
        int K = S.K;
        tab = new float4[(size_t)K * K];
        for (int x = 0; x < K; ++x) for (int y = 0; y < K; ++y) {
            double sc = d_add(A->cov[x], A->cov[y]), sv = d_add(A->vdw[x], A->vdw[y]);
            tab[x * K + y] = float4{ (float)sc, (float)sv, (float)d_add(sv, P->vdw_comp), 0.f };
        }
        S.radtab = tab;
    }
    for (int64_t k = 0; k < n; ++k) {
        int i = b[k], j = e[k];
        auto word = [&](int a) {
            return arp_pack_word(A->feat[a], A->res_flags[A->res_id[a]], A->rad_class[a],
                                 A->bond_off && A->bond_off[a + 1] > A->bond_off[a]);
        };
        uint32_t fb = word(i), fe = word(j);
        int rb = A->res_id[i], re = A->res_id[j];
        bool keep = rule_pair_survives(fb, rb, A->res_prev[rb], A->res_next[rb], fe, re, A->res_prev[re], A->res_next[re],
                                       R.include_seq_adjacent);
        emitted[k] = keep;
        out[k].i = i; out[k].j = j; out[k].mask = 0; out[k].dist = 0.f;
        if (!keep) continue;
        const float* pb = A->xyz + 3 * (size_t)i; const float* pe = A->xyz + 3 * (size_t)j;
        rule_classify(S, R, i, j, pb[0], pb[1], pb[2], pe[0], pe[1], pe[2], fb, fe, &out[k].mask, &out[k].dist);
    }
    delete[] tab;
    delete[] hlim;
    return 0;
}
