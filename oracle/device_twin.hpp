// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product (see rng.hpp header).
//
// "Twin" mode of the oracle (rng_mode 1): the free-running Philox mode of the CUDA library samples
// PedigreeSim's distributions directly (pedigreesim_b200/csrc/psim_meiosis.cuh: a host + device
// header), so its draw sequence is not the reference's and cannot be restated from the Java. What
// the tests need from the CPU side in that mode is
//   * the very same sampler run on the HOST (this file includes the product header and calls it with
//     the streams the kernel uses), so that the kernel's plumbing - stream keys, task compaction,
//     chromatid trace, segment assembly, descriptors - is held bit for bit, and
//   * everything downstream of the random choices done by the oracle's own restatement of the Java
//     (Multivalent.slotsToPatterns / fillGamete, Gamete.fertilization in meiosis.hpp).
// That the sampler's LAWS are the reference's is a separate matter: tests/test_sampler_laws.py and
// tests/test_gpu_stats.py hold twin and GPU against the Java-order restatement (rng_mode 0,
// java.util.Random) by chi-square, and against the manual's tables.
#pragma once
#include <vector>

#include "../pedigreesim_b200/csrc/psim_meiosis.cuh"
#include "meiosis.hpp"

namespace orc {

inline psim::ModelParams twin_model(const Model& m) {
    psim::ModelParams mp{};
    mp.ploidy = m.ploidy;
    mp.chiasma_interference = m.chiasma_interference;
    mp.natural_pairing = m.natural_pairing;
    mp.allow_no_recomb = m.allow_no_recomb;
    mp.parallel_multivalents = m.parallel_multivalents;
    mp.paired_centromeres = m.paired_centromeres;
    mp.seed = m.twin_seed;
    mp.unreduced = m.unreduced_gametes;
    return mp;
}
inline psim::ChromParams twin_chrom(const Model& m, const Chrom& ch) {
    psim::ChromParams cp{};
    cp.head = ch.head; cp.tail = ch.tail; cp.centro = ch.centro;
    cp.pref_pairing = ch.pref_pairing;
    cp.recomb_dist_noempty = calc_recomb_dist(ch.length());
    for (int q = 0; q < psim::MAXP / 4; q++) cp.cumul_binomial[q] = q < (int)ch.cumul_binomial.size() ? ch.cumul_binomial[q] : 1.0;
    (void)m;
    return cp;
}

template <bool QUAD>
inline uint32_t twin_multivalent(psim::ChiasmaList& x, psim::PhiloxStream& rng, const psim::ModelParams& mp, const psim::ChromParams& cp) {
    if (mp.chiasma_interference) return mp.allow_no_recomb ? psim::draw_multivalent<QUAD, true, true>(x, rng, mp, cp)
                                                           : psim::draw_multivalent<QUAD, true, false>(x, rng, mp, cp);
    return mp.allow_no_recomb ? psim::draw_multivalent<QUAD, false, true>(x, rng, mp, cp)
                              : psim::draw_multivalent<QUAD, false, false>(x, rng, mp, cp);
}

// The kept gametes of one meiosis of `parent` (the one that makes `child`): out_hs[c][g * P/2 + h] is
// filled for the kept gametes g only (gamete 0; with 2NGAMETES also gamete `second`).
inline void twin_do_meiosis(const Model& m, const Individual& parent, uint32_t child, uint32_t par,
                            std::vector<std::vector<HS>>& out_hs, std::vector<ChromConfig>& out_conf, int* second_gamete,
                            MeiosisStats* stats) {
    const int P = m.ploidy, C = (int)m.chrom.size(), p2 = P / 2;
    const psim::ModelParams mp = twin_model(m);
    out_hs.assign(C, std::vector<HS>(2 * P));
    out_conf.assign(C, ChromConfig());
    std::vector<double> xpos(4096);
    std::vector<uint8_t> xab(4096);
    for (int c = 0; c < C; c++) {
        const Chrom& ch = m.chrom[c];
        const psim::ChromParams cp = twin_chrom(m, ch);
        psim::Nib16 g_head, g_tail;
        g_head.v = 0; g_tail.v = 0;
        if (P > 2)
            for (int h = 0; h < P; h++) {
                const HS& hs = parent.hs[c][h];
                g_head.set(h, hs.founder[hs.find_segment(ch.head)] % p2);
                g_tail.set(h, hs.founder[hs.find_segment(ch.tail)] % p2);
            }
        psim::PhiloxStream rng;
        rng.init(m.twin_seed, m.twin_replicate);
        rng.seek(child, (uint32_t)c, par, psim::STAGE_PAIRING);
        const psim::PairingResult pr = psim::draw_pairing(rng, mp, cp, g_head, g_tail);
        out_conf[c].quad = pr.quad;
        out_conf[c].chromseq.resize(P);
        for (int h = 0; h < P; h++) out_conf[c].chromseq[h] = pr.seq.get(h);
        if (stats) {
            if ((int)stats->quad_config_count.size() <= pr.quad) stats->quad_config_count.resize(pr.quad + 1, 0);
            stats->quad_config_count[pr.quad]++;
        }
        const int n_mv = pr.quad + (P - 4 * pr.quad) / 2;
        for (int v = 0; v < n_mv; v++) {
            const bool is_quad = v < pr.quad;
            const int k = is_quad ? 4 : 2;
            const int base = is_quad ? 4 * v : 4 * pr.quad + 2 * (v - pr.quad);
            std::vector<const HS*> orig(k);
            for (int h = 0; h < k; h++) orig[h] = &parent.hs[c][pr.seq.get(base + h)];
            psim::ChiasmaList x;
            x.pos = xpos.data(); x.ab = xab.data(); x.stride = 1; x.cap = (int)xpos.size(); x.n = 0; x.overflow = false;
            rng.seek(child, (uint32_t)c, par, psim::STAGE_MULTIVALENT + (uint32_t)v);
            const uint32_t at = is_quad ? twin_multivalent<true>(x, rng, mp, cp) : twin_multivalent<false>(x, rng, mp, cp);
            if (x.overflow) throw std::runtime_error("oracle twin: chiasma list overflow");
            if (stats) { if (is_quad) stats->cross_quads++; else stats->bivalents++; }
            const psim::Transmission tr = psim::draw_transmission(rng, mp, child, (uint32_t)c, par);
            if (second_gamete) *second_gamete = tr.gamete[1];
            // the oracle's own chromatid machinery from here on (meiosis.hpp)
            Multivalent mv(m, ch, orig);
            mv.slots.assign(2 * k, HS(ch.head, -1));
            for (int e = 0; e < x.n; e++) {
                mv.slots[x.a(e)].add_segment(x.p(e), x.b(e));
                mv.slots[x.b(e)].add_segment(x.p(e), x.a(e));
            }
            const std::vector<HS> patterns = mv.slots_to_patterns();
            for (int kk = 0; kk < tr.n_gametes; kk++) {
                const int g = tr.gamete[kk];
                for (int o = 0; o < (is_quad ? 2 : 1); o++) {
                    const int slot = is_quad ? 2 * v + o : 2 * pr.quad + (v - pr.quad);
                    const int chromatid = (int)((at >> (4 * (is_quad ? 2 * g + o : g))) & 15u);
                    out_hs[c][g * p2 + tr.dest[kk].get(slot)] = mv.fill_gamete(patterns[chromatid]);
                }
            }
        }
    }
}

}  // namespace orc
