// Device-side meiosis of one parent on one chromosome ("half unit"): pairing configuration,
// chiasma placement, chromatid tracing, centromere segregation, selection of the transmitted
// gamete. Replaces, for the transmitted gamete only (the reference builds all four and discards
// three, Individual.java:258-259), the call tree
//   Individual.doMeiosis / calcChromConfig            Individual.java:281-349, 601-670
//   Multivalent.doMeiosis and helpers                 Multivalent.java:79-182, 332-527
//   Quadrivalent.doCrossingOver / getCentromereSortOrder   Quadrivalent.java:108-243, 371-510
// (paths relative to /root/reference/src/PedigreeSim/). The draw order inside each Philox stream
// is the reference's, which is what tests/ check bit-for-bit against the CPU oracle.
#pragma once
#include <cstdint>
#include <math_constants.h>

#include "psim_rng.cuh"

namespace psim {

constexpr int MAXP = 16;        // maximum ploidy
constexpr int MAXE = 192;       // maximum chiasmata per multivalent held on the device
constexpr double RECOMBDIST = 0.5;   // Multivalent.java:18
constexpr double GAMMAFACTOR = 2.63; // Tools.java:317

// Model flags + per-chromosome constants, passed by value to the kernels.
struct ModelParams {
    int ploidy;
    int chiasma_interference;
    int natural_pairing;
    int allow_no_recomb;
    double parallel_multivalents;
    double paired_centromeres;
    uint32_t seed;
    int unreduced;   // 2NGAMETES: 0 = haploid gametes, 1 = FDR, 2 = SDR (Individual.java:341-383)
};
struct ChromParams {
    double head, tail, centro;
    double pref_pairing;
    double recomb_dist_noempty;  // Tools.calcRecombDist(length), used when allow_no_recomb == 0
    double cumul_binomial[MAXP / 4];  // TetraploidChromosome.java:102-108
};

// Read-only view of the parent's P homologs on this chromosome (segments in the slab).
struct HomologView {
    const double* pos;
    const int32_t* founder;
    int n;
    // HaploStruct.java:107-112
    __device__ __forceinline__ int find_segment(double position) const {
        int seg = n - 1;
        while (seg > 0 && pos[seg] > position) seg--;
        return seg;
    }
};

// ---- Tools.java samplers
__device__ __forceinline__ double ran_exp(PhiloxStream& r, double mean) {  // Tools.java:238-240
    return -mean * log(1.0 - r.next_double());
}
__device__ inline double ran_gamma_cheng(PhiloxStream& rng, double k, double theta) {  // Tools.java:294-314 (k >= 1)
    const double b = k - log(4.0);
    const double lam = sqrt(2 * k - 1);
    const double c = k + lam;
    const double cheng = 1 + log(4.5);
    double x;
    for (;;) {
        rng.prefetch();
        const double u = rng.next_double();
        const double v = rng.next_double();
        const double y = (1 / lam) * log(v / (1 - v));
        x = k * exp(y);
        const double z = u * v * v;
        const double r = b + (c * y) - x;
        if ((r >= ((4.5 * z) - cheng)) || (r >= log(z))) break;
    }
    return x * theta;
}
__device__ __forceinline__ double ran_dist_interference(PhiloxStream& r, double meandist) {  // Tools.java:329-331
    return ran_gamma_cheng(r, GAMMAFACTOR, meandist / GAMMAFACTOR);
}
__device__ inline double dist_to_first_recomb(PhiloxStream& r, const ModelParams& m, double meandist) {  // Multivalent.java:84-97
    double p = ran_exp(r, meandist);
    if (m.chiasma_interference) {
        const double burnin = 10.0 * meandist;
        while (p < burnin) p += ran_dist_interference(r, meandist);
        return p - burnin;
    }
    return p;
}
__device__ __forceinline__ double dist_to_next_recomb(PhiloxStream& r, const ModelParams& m, double meandist) {  // Multivalent.java:99-105
    return m.chiasma_interference ? ran_dist_interference(r, meandist) : ran_exp(r, meandist);
}
// Tools.java:33-51 on a small int8 array
__device__ __forceinline__ void shuffle_small(PhiloxStream& r, int8_t* a, int n) {
    int last = n - 1;
    while (last > 0) {
        const int ix = r.next_int(last + 1);
        const int8_t tmp = a[last]; a[last] = a[ix]; a[ix] = tmp;
        last--;
    }
}

// ---------------------------------------------------------------- pairing configuration
struct PairingState {
    int8_t initorder[MAXP], backorder[MAXP];
    int8_t paired[2][MAXP];
    int8_t genomenr[2][MAXP];
    int8_t ghs[2][MAXP / 2][MAXP];  // Individual.genomehs for this chromosome
    int8_t ghs_n[2][MAXP / 2];
    int8_t gen[MAXP / 2][MAXP];     // Individual.genohs (working copy)
    int8_t gen_n[MAXP / 2];
    bool hsdone[MAXP];
    int possible[2];
    int primary, secondary;
    int bivalentcount, quadcount;
    int8_t chromseq[MAXP];
    int quad;
};

__device__ __forceinline__ int list_index_of(const int8_t* v, int n, int x) {
    for (int i = 0; i < n; i++) if (v[i] == x) return i;
    return -1;
}
__device__ __forceinline__ void list_remove(int8_t* v, int8_t& n, int idx) {
    for (int i = idx; i + 1 < n; i++) v[i] = v[i + 1];
    n--;
}
// Individual.java:437-470 for one chromosome
__device__ inline void calc_genome(PairingState& s, const HomologView* hv, int P, const ChromParams& ch) {
    for (int side = 0; side < 2; side++) {
        for (int g = 0; g < P / 2; g++) s.ghs_n[side][g] = 0;
        for (int h = 0; h < P; h++) {
            const double at = side == 0 ? ch.head : ch.tail;
            const int g = hv[h].founder[hv[h].find_segment(at)] % (P / 2);
            s.genomenr[side][h] = (int8_t)g;
            s.ghs[side][g][s.ghs_n[side][g]++] = (int8_t)h;
        }
        s.possible[side] = 0;
        for (int g = 0; g < P / 2; g++) s.possible[side] += s.ghs_n[side][g] / 2;
    }
}
// Individual.java:574-593 (removal uses genomenr at `secondary`, whatever `side` is)
__device__ inline void calc_genohs(PairingState& s, int P, int side) {
    for (int g = 0; g < P / 2; g++) {
        s.gen_n[g] = s.ghs_n[side][g];
        for (int i = 0; i < s.gen_n[g]; i++) s.gen[g][i] = s.ghs[side][g][i];
    }
    for (int h = 0; h < P; h++) {
        if (s.hsdone[h]) {
            const int gnm = s.genomenr[s.secondary][s.initorder[h]];
            const int idx = list_index_of(s.gen[gnm], s.gen_n[gnm], s.initorder[h]);
            list_remove(s.gen[gnm], s.gen_n[gnm], idx);
        }
    }
}
// Individual.java:483-523
__device__ inline void get_pairing(PairingState& s, PhiloxStream& rng, int P, int side, double pref) {
    bool done[MAXP];
    for (int h = 0; h < P; h++) done[h] = false;
    calc_genohs(s, P, side);
    int firsth = 0;
    while (firsth < P - 1) {
        while (firsth < P - 1 && done[firsth]) firsth++;
        if (firsth < P - 1) {
            done[firsth] = true;
            int gnm = s.genomenr[side][s.initorder[firsth]];
            int hidx = list_index_of(s.gen[gnm], s.gen_n[gnm], s.initorder[firsth]);
            list_remove(s.gen[gnm], s.gen_n[gnm], hidx);
            int secondh;
            if (s.gen_n[gnm] != 0 && rng.next_double() < pref) {
                hidx = rng.next_int(s.gen_n[gnm]);
                secondh = s.backorder[s.gen[gnm][hidx]];
            } else {
                secondh = firsth;
                while (done[secondh]) secondh = rng.next_int(P - firsth - 1) + firsth + 1;
                gnm = s.genomenr[side][s.initorder[secondh]];
                hidx = list_index_of(s.gen[gnm], s.gen_n[gnm], s.initorder[secondh]);
            }
            done[secondh] = true;
            list_remove(s.gen[gnm], s.gen_n[gnm], hidx);
            s.paired[side][firsth] = (int8_t)secondh;
            s.paired[side][secondh] = (int8_t)firsth;
            firsth++;
        }
    }
}
__device__ __forceinline__ void add_bivalent(PairingState& s, int P, int firsth, int secondh) {  // Individual.java:672-677
    s.bivalentcount++;
    s.chromseq[P - 2 * s.bivalentcount] = s.initorder[firsth];
    s.chromseq[P - 2 * s.bivalentcount + 1] = s.initorder[secondh];
}
__device__ __forceinline__ void add_quadrivalent(PairingState& s, int firsth, int secondh, int thirdh, int fourth) {  // Individual.java:679-703
    const int b = s.quadcount * 4;
    if (s.primary == 0) {
        s.chromseq[b] = s.initorder[firsth]; s.chromseq[b + 1] = s.initorder[secondh];
        s.chromseq[b + 2] = s.initorder[thirdh]; s.chromseq[b + 3] = s.initorder[fourth];
    } else {
        s.chromseq[b] = s.initorder[secondh]; s.chromseq[b + 1] = s.initorder[thirdh];
        s.chromseq[b + 2] = s.initorder[fourth]; s.chromseq[b + 3] = s.initorder[firsth];
    }
    s.quadcount++;
}
__device__ inline void get_spontaneous_bivalents(PairingState& s, int P, int max_bivalents) {  // Individual.java:705-725
    int firsth = 0;
    while (firsth < P - 1 && s.bivalentcount < max_bivalents) {
        while (firsth < P - 1 && (s.hsdone[firsth] || s.paired[s.primary][firsth] != s.paired[s.secondary][firsth])) firsth++;
        if (firsth < P - 1) {
            const int secondh = s.paired[s.primary][firsth];
            add_bivalent(s, P, firsth, secondh);
            s.hsdone[firsth] = true;
            s.hsdone[secondh] = true;
        }
        firsth++;
    }
}
__device__ inline void get_spontaneous_quadrivalents(PairingState& s, int P, int max_quadrivalents) {  // Individual.java:727-787
    bool tmp_bi[MAXP];
    for (int h = 0; h < P; h++) tmp_bi[h] = false;
    int firsth = 0;
    while (firsth < P - 1) {
        while (firsth < P - 1 && (s.hsdone[firsth] || s.paired[s.primary][firsth] != s.paired[s.secondary][firsth])) firsth++;
        if (firsth < P - 1) {
            const int secondh = s.paired[s.primary][firsth];
            s.hsdone[firsth] = true; s.hsdone[secondh] = true;
            tmp_bi[firsth] = true; tmp_bi[secondh] = true;
        }
        firsth++;
    }
    firsth = 0;
    while (firsth < P - 3 && s.quadcount < max_quadrivalents) {
        while (firsth < P - 3 && s.hsdone[firsth]) firsth++;
        if (firsth < P - 3) {
            const int secondh = s.paired[s.secondary][firsth];
            const int thirdh = s.paired[s.primary][secondh];
            const int fourth = s.paired[s.primary][firsth];
            if (s.paired[s.secondary][thirdh] == fourth) {
                add_quadrivalent(s, firsth, secondh, thirdh, fourth);
                s.hsdone[firsth] = true; s.hsdone[secondh] = true;
                s.hsdone[thirdh] = true; s.hsdone[fourth] = true;
            }
        }
        firsth++;
    }
    for (int h = 0; h < P; h++) if (tmp_bi[h]) s.hsdone[h] = false;
}
__device__ inline void get_forced_bivalent(PairingState& s, int P) {  // Individual.java:789-799
    int firsth = 0;
    while (firsth < P && s.hsdone[firsth]) firsth++;
    const int secondh = s.paired[s.primary][firsth];
    add_bivalent(s, P, firsth, secondh);
    s.hsdone[firsth] = true;
    s.hsdone[secondh] = true;
}
__device__ inline void get_forced_bi_or_quadri(PairingState& s, PhiloxStream& rng, int P, bool bi_allowed, double pref) {  // Individual.java:822-904
    int firsth = 0, secondh, thirdh, fourth;
    while (firsth < P - 1 && s.hsdone[firsth]) firsth++;
    fourth = s.paired[s.primary][firsth];
    int gnm = s.genomenr[s.secondary][s.initorder[firsth]];
    int gnm4 = s.genomenr[s.secondary][s.initorder[fourth]];
    if (s.gen_n[gnm4] > s.gen_n[gnm]) {
        int t = firsth; firsth = fourth; fourth = t;
        t = gnm; gnm = gnm4; gnm4 = t;
    }
    int hidx = list_index_of(s.gen[gnm], s.gen_n[gnm], s.initorder[firsth]);
    list_remove(s.gen[gnm], s.gen_n[gnm], hidx);
    s.hsdone[firsth] = true;
    // the reference takes fourth's index now and removes it later (after other removals)
    int hidx4 = list_index_of(s.gen[gnm4], s.gen_n[gnm4], s.initorder[fourth]);
    if (((bi_allowed && s.gen_n[gnm] > 0) || (!bi_allowed && s.gen_n[gnm] > 1)) && rng.next_double() < pref) {
        do {
            hidx = rng.next_int(s.gen_n[gnm]);
            secondh = s.backorder[s.gen[gnm][hidx]];
        } while (secondh == fourth && !bi_allowed);
        s.hsdone[fourth] = true;
        list_remove(s.gen[gnm4], s.gen_n[gnm4], hidx4);
        if (secondh == fourth) {
            add_bivalent(s, P, firsth, secondh);
        } else {
            gnm = s.genomenr[s.secondary][s.initorder[secondh]];
            hidx = list_index_of(s.gen[gnm], s.gen_n[gnm], s.initorder[secondh]);
            list_remove(s.gen[gnm], s.gen_n[gnm], hidx);
            s.hsdone[secondh] = true;
            thirdh = s.paired[s.primary][secondh];
            gnm = s.genomenr[s.secondary][s.initorder[thirdh]];
            hidx = list_index_of(s.gen[gnm], s.gen_n[gnm], s.initorder[thirdh]);
            list_remove(s.gen[gnm], s.gen_n[gnm], hidx);
            s.hsdone[thirdh] = true;
            add_quadrivalent(s, firsth, secondh, thirdh, fourth);
        }
    } else {
        secondh = firsth;
        while (s.hsdone[secondh] || (secondh == fourth && !bi_allowed)) secondh = rng.next_int(P);
        s.hsdone[fourth] = true;
        list_remove(s.gen[gnm4], s.gen_n[gnm4], hidx4);
        if (secondh == fourth) {
            add_bivalent(s, P, firsth, secondh);
            s.hsdone[secondh] = true;
        } else {
            thirdh = s.paired[s.primary][secondh];
            s.hsdone[secondh] = true;
            gnm = s.genomenr[s.secondary][s.initorder[secondh]];
            hidx = list_index_of(s.gen[gnm], s.gen_n[gnm], s.initorder[secondh]);
            list_remove(s.gen[gnm], s.gen_n[gnm], hidx);
            s.hsdone[thirdh] = true;
            gnm = s.genomenr[s.secondary][s.initorder[thirdh]];
            hidx = list_index_of(s.gen[gnm], s.gen_n[gnm], s.initorder[thirdh]);
            list_remove(s.gen[gnm], s.gen_n[gnm], hidx);
            add_quadrivalent(s, firsth, secondh, thirdh, fourth);
        }
    }
}
// Individual.java:601-670. Leaves s.chromseq / s.quad.
__device__ inline void calc_chrom_config(PairingState& s, PhiloxStream& rng, const ModelParams& m,
                                         const ChromParams& ch, const HomologView* hv) {
    const int P = m.ploidy;
    if (P == 2) {
        s.chromseq[0] = 0; s.chromseq[1] = 1;
        shuffle_small(rng, s.chromseq, 2);
        s.quad = 0;
        return;
    }
    for (int h = 0; h < P; h++) s.initorder[h] = (int8_t)h;
    shuffle_small(rng, s.initorder, P);
    for (int h = 0; h < P; h++) s.backorder[s.initorder[h]] = (int8_t)h;
    for (int h = 0; h < P; h++) { s.hsdone[h] = false; s.chromseq[h] = 0; }
    s.secondary = 0;
    calc_genome(s, hv, P, ch);
    for (int side = 0; side < 2; side++) get_pairing(s, rng, P, side, ch.pref_pairing);
    {   // Individual.java:536-542
        const int c0 = s.possible[0], c1 = s.possible[1];
        const double probside0 = c0 == c1 ? 0.5 : (double)c0 / (double)(c0 + c1);
        s.primary = rng.next_double() < probside0 ? 0 : 1;
    }
    s.secondary = 1 - s.primary;
    s.bivalentcount = 0;
    s.quadcount = 0;
    if (m.natural_pairing) {
        get_spontaneous_bivalents(s, P, P / 2);
        get_spontaneous_quadrivalents(s, P, (P - 2 * s.bivalentcount) / 4);
        calc_genohs(s, P, s.secondary);
        while (2 * s.bivalentcount + 4 * s.quadcount < P) {
            if (2 * s.bivalentcount + 4 * s.quadcount == P - 2) get_forced_bivalent(s, P);
            else get_forced_bi_or_quadri(s, rng, P, true, ch.pref_pairing);
        }
    } else {
        // TetraploidChromosome.java:122-128
        const double p = rng.next_double();
        int num_quad = 0;
        while (num_quad < P / 4 && ch.cumul_binomial[num_quad] < p) num_quad++;
        const int num_bi = (P - 4 * num_quad) / 2;
        get_spontaneous_bivalents(s, P, num_bi);
        get_spontaneous_quadrivalents(s, P, num_quad);
        while (s.bivalentcount < num_bi) get_forced_bivalent(s, P);
        if (s.quadcount < num_quad) {
            calc_genohs(s, P, s.secondary);
            while (s.quadcount < num_quad) get_forced_bi_or_quadri(s, rng, P, false, ch.pref_pairing);
        }
    }
    s.quad = s.quadcount;
}

// ---------------------------------------------------------------- chiasmata of one multivalent
// All chiasmata of a multivalent as one position-sorted list; each involves chromatid slots a, b
// (slot s belongs to chromosome s/2 of the multivalent). Per slot this is the reference's record
// list (Multivalent.java:158-177, Quadrivalent.makeRecombination :276-284).
struct Chiasmata {
    double pos[MAXE];
    uint8_t a[MAXE], b[MAXE];
    int n;
    bool overflow;
    __device__ __forceinline__ void clear() { n = 0; }
    __device__ __forceinline__ void append(double p, int sa, int sb) {
        if (n >= MAXE) { overflow = true; return; }
        pos[n] = p; a[n] = (uint8_t)sa; b[n] = (uint8_t)sb; n++;
    }
    // HaploStruct.insertSegment (:129-133): after the last record whose position is <= p
    __device__ __forceinline__ void insert_sorted(double p, int sa, int sb) {
        if (n >= MAXE) { overflow = true; return; }
        int i = n;
        while (i > 0 && pos[i - 1] > p) { pos[i] = pos[i - 1]; a[i] = a[i - 1]; b[i] = b[i - 1]; i--; }
        pos[i] = p; a[i] = (uint8_t)sa; b[i] = (uint8_t)sb; n++;
    }
    // Multivalent.java:587-603
    __device__ __forceinline__ bool all_chrom_recomb(int k) const {
        if (k == 2) return n > 0;
        unsigned seen = 0;
        for (int e = 0; e < n; e++) seen |= (1u << (a[e] >> 1)) | (1u << (b[e] >> 1));
        return seen == 0xFu;
    }
};

// Multivalent.java:139-182 — bivalent (k = 2) and parallel quadrivalent (k = 4)
__device__ inline void parallel_crossing_over(Chiasmata& x, PhiloxStream& rng, const ModelParams& m,
                                              const ChromParams& ch, int k) {
    const int slotcount = 2 * k;
    double recomb_dist = k == 2 ? RECOMBDIST : RECOMBDIST * 2 / (double)k;
    if (!m.allow_no_recomb) recomb_dist = recomb_dist / 0.5 * ch.recomb_dist_noempty;
    do {
        x.clear();
        double p = ch.head + dist_to_first_recomb(rng, m, recomb_dist);
        while (p < ch.tail) {
            rng.prefetch();
            const int i0 = rng.next_int(slotcount);
            int i1;
            do { i1 = rng.next_int(slotcount); } while (i1 / 2 == i0 / 2);
            x.append(p, i0, i1);
            p += dist_to_next_recomb(rng, m, recomb_dist);
            if (x.overflow) return;
        }
    } while (!(m.allow_no_recomb || x.all_chrom_recomb(k)));
}

// Quadrivalent.java:371-424 (quadriMethod 2, quadriEachArm false). Returns exchangeLim[1].
__device__ inline double quadrivalent_crossing_over(Chiasmata& x, PhiloxStream& rng, const ModelParams& m,
                                                    const ChromParams& ch) {
    const bool paral = m.parallel_multivalents == 0.0 ? false
                     : m.parallel_multivalents == 1.0 ? true
                     : rng.next_double() < m.parallel_multivalents;
    if (paral) { parallel_crossing_over(x, rng, m, ch, 4); return 0.0; }  // fresh object: exchangeLim = {0,0}
    // Quadrivalent.java:14-19: chromatids with their head in arm 0 / tail in arm 1 / head in arm 2 / tail in arm 3
    const int ARM[4][4] = {{0, 1, 6, 7}, {0, 1, 2, 3}, {2, 3, 4, 5}, {4, 5, 6, 7}};
    double exchange_lim[2];
    do {
        const double recomb_dist = m.allow_no_recomb ? RECOMBDIST : ch.recomb_dist_noempty;
        x.clear();
        double lastpos[4];
        bool arm_finished[4];
        for (int a = 0; a < 4; a++) { lastpos[a] = CUDART_NAN; arm_finished[a] = false; }
        bool finished = false;
        exchange_lim[0] = ch.head;
        exchange_lim[1] = ch.tail;
        do {
            rng.prefetch();
            const int a = rng.next_int(4);
            if (arm_finished[a]) continue;
            // generateRecombination, Quadrivalent.java:108-139
            const int c0 = ARM[a][rng.next_int(2)];
            const int c1 = ARM[a][2 + rng.next_int(2)];
            const int side = a & 1;
            const double sf = side == 0 ? 1.0 : -1.0;
            const double side_pos = side == 1 ? ch.tail : ch.head;
            const double oppo_pos = side == 1 ? ch.head : ch.tail;
            double pos;
            if (isnan(lastpos[a])) pos = side_pos + sf * dist_to_first_recomb(rng, m, recomb_dist);
            else pos = lastpos[a] + sf * dist_to_next_recomb(rng, m, recomb_dist);
            const int oa0 = side == 0 ? 1 : 0, oa1 = side == 0 ? 3 : 2;  // oppoArms
            double mindist = CUDART_NAN;
            int nearest = -1;
            if (!isnan(lastpos[oa0])) { mindist = sf * (lastpos[oa0] - pos); nearest = oa0; }
            if (!isnan(lastpos[oa1]) && (isnan(mindist) || mindist > sf * (lastpos[oa1] - pos))) {
                mindist = sf * (lastpos[oa1] - pos);
                nearest = oa1;
            }
            // testRecomb2, Quadrivalent.java:141-243
            double temp_lim = exchange_lim[side];
            const double dist_exchangelim = sf * (pos - exchange_lim[1 - side]);
            bool reject = dist_exchangelim >= 0.0;
            if (!reject && !isnan(mindist)) {
                // Multivalent.rejectDist :216-220
                reject = mindist <= 0 || (m.chiasma_interference && mindist < 0.5 &&
                                          rng.next_double() < pow(1.0 - 2 * mindist, 2.25));
            }
            if (reject) {
                arm_finished[a] = true;
                if (dist_exchangelim >= 0.0) {
                    temp_lim = exchange_lim[1 - side];
                    for (int oa = 0; oa < 2; oa++) {
                        const int arm_o = oa == 0 ? oa0 : oa1;
                        const double lpoa = lastpos[arm_o];
                        if (temp_lim == oppo_pos || (!isnan(lpoa) && (sf * (lpoa - temp_lim)) <= 0)) arm_finished[arm_o] = true;
                    }
                } else {
                    if (sf * (pos - temp_lim) > 0) temp_lim = pos;
                    arm_finished[nearest] = true;
                }
                finished = arm_finished[0] && arm_finished[1] && arm_finished[2] && arm_finished[3];
            } else {
                lastpos[a] = pos;
                if ((sf * (pos - temp_lim)) > 0) temp_lim = pos;
            }
            exchange_lim[side] = temp_lim;
            if (!arm_finished[a]) {  // makeRecombination :276-284
                x.insert_sorted(pos, c0, c1);
                lastpos[a] = pos;
                if (x.overflow) return exchange_lim[1];
            }
        } while (!finished);
    } while (!(m.allow_no_recomb || x.all_chrom_recomb(4)));
    return exchange_lim[1];
}

// Chromosome (0..k-1) carried at `position` by the chromatid that starts in slot p
// (Multivalent.slotsToPatterns :332-357 followed by getFounderAtCentromere).
__device__ __forceinline__ int pattern_chrom_at(const Chiasmata& x, int p, double position) {
    int cur = p;
    for (int e = 0; e < x.n && x.pos[e] <= position; e++) {
        if (x.a[e] == cur) cur = x.b[e];
        else if (x.b[e] == cur) cur = x.a[e];
    }
    return cur >> 1;
}

// Multivalent.getCentromereSortOrder :397-437 (k = 2 or parallel / unpaired quadrivalent)
__device__ inline void multivalent_centromere_order(PhiloxStream& rng, int k, int8_t* centro) {
    if (k == 2) {
        if (rng.next_bool()) { centro[0] = 0; centro[1] = 0; centro[2] = 1; centro[3] = 1; }
        else { centro[0] = 1; centro[1] = 1; centro[2] = 0; centro[3] = 0; }
        return;
    }
    int8_t chr[4] = {0, 1, 2, 3};
    shuffle_small(rng, chr, 4);
    int8_t gam[4][2] = {{chr[0], chr[1]}, {chr[0], chr[1]}, {chr[2], chr[3]}, {chr[2], chr[3]}};
    for (int g = 0; g < 4; g++) { shuffle_small(rng, gam[g], 2); centro[2 * g] = gam[g][0]; centro[2 * g + 1] = gam[g][1]; }
}
// Quadrivalent.getCentromereSortOrder :465-510
__device__ inline void quadrivalent_centromere_order(PhiloxStream& rng, const ModelParams& m, const ChromParams& ch,
                                                     double exchange_lim1, int8_t* centro) {
    const bool paired = m.paired_centromeres == 0.0 ? false
                      : m.paired_centromeres == 1.0 ? true
                      : rng.next_double() < m.paired_centromeres;
    if (!paired) { multivalent_centromere_order(rng, 4, centro); return; }
    int8_t top[2], bot[2];
    if (rng.next_bool()) {
        const bool head_arm = ch.centro <= exchange_lim1;
        if (head_arm) {
            if (rng.next_bool()) { top[0] = 0; top[1] = 1; bot[0] = 2; bot[1] = 3; }
            else { top[0] = 2; top[1] = 3; bot[0] = 0; bot[1] = 1; }
        } else {
            if (rng.next_bool()) { top[0] = 0; top[1] = 3; bot[0] = 1; bot[1] = 2; }
            else { top[0] = 1; top[1] = 2; bot[0] = 0; bot[1] = 3; }
        }
    } else {
        if (rng.next_bool()) { top[0] = 0; top[1] = 2; bot[0] = 1; bot[1] = 3; }
        else { top[0] = 1; top[1] = 3; bot[0] = 0; bot[1] = 2; }
    }
    int8_t gam[4][2] = {{top[0], top[1]}, {top[0], top[1]}, {bot[0], bot[1]}, {bot[0], bot[1]}};
    for (int g = 0; g < 4; g++) { shuffle_small(rng, gam[g], 2); centro[2 * g] = gam[g][0]; centro[2 * g + 1] = gam[g][1]; }
}

// Multivalent.patternsToGametes :486-518 on indices: idx[s] = the starting slot of the chromatid
// that ends up as output pattern s. Gamete g receives pattern g of a bivalent, patterns 2g and
// 2g+1 of a quadrivalent.
__device__ inline void select_gametes(const Chiasmata& x, PhiloxStream& rng, int k, const ChromParams& ch,
                                      const int8_t* centro, int8_t* idx) {
    const int n = 2 * k;
    int8_t fac[8];
    for (int p = 0; p < n; p++) { idx[p] = (int8_t)p; fac[p] = (int8_t)pattern_chrom_at(x, p, ch.centro); }
    for (int s = 0; s < n - 1; s++) {
        if (fac[idx[s]] != centro[s]) {
            int t = s + 1;
            while (fac[idx[t]] != centro[s]) t++;
            const int8_t tmp = idx[s]; idx[s] = idx[t]; idx[t] = tmp;
        }
    }
    if (n > 4) {
        for (int all = 0; all < 4; all++) {
            if (rng.next_bool()) {
                int first = 0; while (centro[first] != all) first++;
                int second = first + 1; while (centro[second] != all) second++;
                const int8_t tmp = idx[first]; idx[first] = idx[second]; idx[second] = tmp;
            }
        }
    }
}

}  // namespace psim
