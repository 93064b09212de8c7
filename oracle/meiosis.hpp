// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product (see rng.hpp header).
//
// CPU restatement of PedigreeSim's per-individual meiosis path. Every function cites the
// reference lines it follows (paths relative to /root/reference/src/PedigreeSim/).
// Draw order is the reference's, so with JavaRandom(seed) this reproduces the jar's draw
// sequence; with PhiloxStream the same code is the specification the CUDA kernels must match.
//
// PARITY STATUS: the reference ships no tests or golden outputs and cannot run here (no JVM),
// so this restatement is pinned by (i) the closed forms coded in the reference
// (Tools.java:116,140), (ii) manual Tables 1 and 2, (iii) structural identities — see
// tests/test_oracle_*.py. Draw-for-draw equality with the jar is UNPINNED until a box with a
// JVM has run tests/golden/make_jar_vectors.sh.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "rng.hpp"

namespace orc {

// HaploStruct.java:29-48 — founder allele per segment + START position (Morgan) per segment.
struct HS {
    std::vector<int> founder;
    std::vector<double> pos;
    HS() {}
    HS(double head, int f) { founder.push_back(f); pos.push_back(head); }
    int count() const { return (int)founder.size(); }
    // HaploStruct.java:107-112 — last segment whose start is <= position (scan from the tail).
    int find_segment(double position) const {
        int seg = (int)pos.size() - 1;
        while (seg > 0 && pos[seg] > position) seg--;
        return seg;
    }
    // HaploStruct.java:119-122
    void add_segment(double position, int f) { pos.push_back(position); founder.push_back(f); }
    // HaploStruct.java:129-133
    void insert_segment(double position, int f) {
        int seg = find_segment(position);
        pos.insert(pos.begin() + seg + 1, position);
        founder.insert(founder.begin() + seg + 1, f);
    }
};

// Chromosome.java:65-88 + TetraploidChromosome.java:80-108.
struct Chrom {
    std::string name;
    double head = 0.0, tail = 0.0, centro = 0.0;  // Morgan
    double pref_pairing = 0.0, frac_quad = 0.0;
    std::vector<double> cumul_binomial;  // TetraploidChromosome.java:102-108
    double length() const { return tail - head; }
    double side_pos(int side) const { return side == 1 ? tail : head; }  // Chromosome.java:120
    // Chromosome.java:65-80: head/tail are widened to include the centromere.
    void set(double head_, double tail_, double centro_) {
        centro = centro_;
        head = centro_ < head_ ? centro_ : head_;
        tail = centro_ > tail_ ? centro_ : tail_;
    }
    // JSci BinomialDistribution.cumulative restated: sum_{k<=i} C(n,k) p^k (1-p)^(n-k) with
    // C(n,k) = exp(lgamma(n+1)-lgamma(k+1)-lgamma(n-k+1)). For ploidy 4 and 6 (n = 1) this is
    // exactly {1-p} (lgamma(1) = lgamma(2) = 0).
    void set_cumul_binomial(int ploidy) {
        const int n = ploidy / 4;
        cumul_binomial.assign(n, 0.0);
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int k = 0; k <= i; k++) {
                double binom = std::exp(std::lgamma(n + 1.0) - std::lgamma(k + 1.0) - std::lgamma(n - k + 1.0));
                s += binom * std::pow(frac_quad, k) * std::pow(1.0 - frac_quad, n - k);
            }
            cumul_binomial[i] = s;
        }
    }
};

// PopulationData.java:187-222 (model flags only).
struct Model {
    int ploidy = 2;
    bool chiasma_interference = false;  // MAPFUNCTION=KOSAMBI
    double parallel_multivalents = 0.0;
    double paired_centromeres = 0.0;
    bool natural_pairing = true;
    bool allow_no_recomb = true;
    int unreduced_gametes = 0;          // 2NGAMETES: 0 = haploid gametes, 1 = FDR, 2 = SDR (Main.java:179-185)
    std::vector<Chrom> chrom;
    Rng* rng = nullptr;
    // rng_mode 1 ("twin", device_twin.hpp): the CUDA library's own sampler run on the host
    bool twin = false;
    uint64_t twin_seed = 0;
    uint32_t twin_replicate = 0;
};

// Genotype.java:38-51
struct ChromConfig {
    int quad = 0;
    std::vector<int> chromseq;
};

// ---------------------------------------------------------------- Tools.java
// Tools.java:33-51 — Fisher–Yates from the top.
inline std::vector<int> shuffle_array(Rng& r, std::vector<int> a) {
    int last = (int)a.size() - 1;
    while (last > 0) {
        int ix = r.next_int(last + 1);
        int tmp = a[last]; a[last] = a[ix]; a[ix] = tmp;
        last--;
    }
    return a;
}
inline std::vector<int> seq(int from, int to) {  // Tools.java:74-83 (inclusive)
    std::vector<int> v;
    int step = to < from ? -1 : 1;
    for (int x = from;; x += step) { v.push_back(x); if (x == to) break; }
    return v;
}
inline std::vector<int> seq(int n) {  // Tools.java:58-64
    std::vector<int> v(n);
    for (int i = 0; i < n; i++) v[i] = i;
    return v;
}
// Tools.java:238-240
inline double ran_exp(Rng& r, double mean) { return -mean * std::log(1.0 - r.next_double()); }
// Tools.java:275-315 — Cheng's algorithm for k >= 1, Weibull-based for k < 1.
inline double ran_gamma(Rng& rng, double k, double theta) {
    bool accept = false;
    if (k < 1) {
        double c = 1 / k;
        double d = (1 - k) * std::pow(k, k / (1 - k));
        double u, v, z, e, x;
        do {
            u = rng.next_double();
            v = rng.next_double();
            z = -std::log(u);
            e = -std::log(v);
            x = std::pow(z, c);
            if ((z + e) >= (d + x)) accept = true;
        } while (!accept);
        return x * theta;
    } else {
        double b = k - std::log(4.0);
        double c = k + std::sqrt(2 * k - 1);
        double lam = std::sqrt(2 * k - 1);
        double cheng = 1 + std::log(4.5);
        double u, v, x, y, z, r;
        do {
            u = rng.next_double();
            v = rng.next_double();
            y = (1 / lam) * std::log(v / (1 - v));
            x = k * std::exp(y);
            z = u * v * v;
            r = b + (c * y) - x;
            if ((r >= ((4.5 * z) - cheng)) || (r >= std::log(z))) accept = true;
        } while (!accept);
        return x * theta;
    }
}
constexpr double GAMMAFACTOR = 2.63;  // Tools.java:317
inline double ran_dist_interference(Rng& r, double meandist) {  // Tools.java:329-331
    return ran_gamma(r, GAMMAFACTOR, meandist / GAMMAFACTOR);
}
// Tools.java:333-386 — mean chiasma spacing when bivalents without recombination are rejected.
inline double mdiff(double L, double m) { return std::fabs(m - 0.5 / (1 - std::exp(-L / m))); }
inline double calc_recomb_dist(double chrom_length) {
    if (chrom_length <= 0.500) return std::nan("");
    if (chrom_length > 5.0) return 0.5;
    double toler = 0.00001;
    double lastdiff = mdiff(chrom_length, 0.5);
    double m1 = 10.5;
    while (mdiff(chrom_length, m1) < lastdiff) {
        lastdiff = mdiff(chrom_length, m1);
        m1 += 10.0;
    }
    double m2, m3;
    if (m1 == 10.5) { m2 = 5.5; m3 = 0.5; }
    else { m2 = m1 - 10.0; m3 = m1 - 20.0; }
    while (m1 - m3 > toler) {
        if (mdiff(chrom_length, m1 - 0.5 * (m1 - m2)) < mdiff(chrom_length, m2)) {
            m3 = m2; m2 = m1 - 0.5 * (m1 - m2);
        } else if (mdiff(chrom_length, m3 + 0.5 * (m2 - m3)) < mdiff(chrom_length, m2)) {
            m1 = m2; m2 = m3 + 0.5 * (m2 - m3);
        } else {
            m1 = m1 - 0.5 * (m1 - m2);
            m3 = m3 + 0.5 * (m2 - m3);
        }
    }
    return m2;
}
// Tools.java:116-118, 140-142 — the closed forms the reference's TEST mode checks against.
inline double haldane_to_recomb(double morgan) { return 0.5 * (1.0 - std::exp(-2 * std::fabs(morgan))); }
inline double kosambi_to_recomb(double morgan) { return 0.5 * std::tanh(2 * std::fabs(morgan)); }

// ---------------------------------------------------------------- Multivalent.java / Quadrivalent.java
constexpr double RECOMBDIST = 0.5;  // Multivalent.java:18

struct Multivalent {
    const Model& m;
    const Chrom& chrom;
    Rng& rng;
    std::vector<const HS*> hs;  // the 2 or 4 parental homologs, caller's order
    // Quadrivalent.java:68-80 scratch (exchangeLim starts {0,0} in a fresh object)
    std::vector<HS> slots;
    int side = 0, nearest_arm = -1;
    double side_fact = 1.0, mindist = 0.0, pos = 0.0;
    int chromatid[2] = {0, 0};
    double exchange_lim[2] = {0.0, 0.0};
    double lastpos[4] = {0, 0, 0, 0};
    bool arm_finished[4] = {false, false, false, false};
    bool crossover_finished = false;
    // statistics mirrored from the reference's TEST mode (Quadrivalent.java:426-455)
    int recomb_count = 0;
    bool was_cross_type = false;

    Multivalent(const Model& model, const Chrom& c, const std::vector<const HS*>& h)
        : m(model), chrom(c), rng(*model.rng), hs(h) {
        assert(h.size() == 2 || h.size() == 4);  // Bivalent.java:26-39, Quadrivalent.java:29-45
    }
    bool is_quad() const { return hs.size() == 4; }

    // Multivalent.java:84-97
    double dist_to_first_recomb(double meandist) {
        double p = ran_exp(rng, meandist);
        if (m.chiasma_interference) {
            double burnin = 10.0 * meandist;
            while (p < burnin) p += ran_dist_interference(rng, meandist);
            return p - burnin;
        }
        return p;
    }
    // Multivalent.java:99-105
    double dist_to_next_recomb(double meandist) {
        return m.chiasma_interference ? ran_dist_interference(rng, meandist) : ran_exp(rng, meandist);
    }
    // Multivalent.java:587-603
    static bool all_chrom_recomb(const std::vector<HS>& s) {
        if (s.size() == 4) return s[0].count() > 1 || s[1].count() > 1;
        bool result;
        int p = (int)s.size() / 2 - 1;
        do {
            result = s[2 * p].count() > 1 || s[2 * p + 1].count() > 1;
            p--;
        } while (result && p >= 0);
        return result;
    }
    // Multivalent.java:139-182 — bivalent and parallel multivalent crossing-over.
    void parallel_crossing_over() {
        const int slotcount = (int)hs.size() * 2;
        double recomb_dist = hs.size() == 2 ? RECOMBDIST : RECOMBDIST * 2 / (double)hs.size();
        if (!m.allow_no_recomb) recomb_dist = recomb_dist / 0.5 * calc_recomb_dist(chrom.length());
        do {
            slots.assign(slotcount, HS(chrom.head, -1));
            double p = chrom.head + dist_to_first_recomb(recomb_dist);
            int ix[2];
            while (p < chrom.tail) {
                ix[0] = rng.next_int(slotcount);
                do { ix[1] = rng.next_int(slotcount); } while (ix[1] / 2 == ix[0] / 2);
                slots[ix[0]].add_segment(p, ix[1]);
                slots[ix[1]].add_segment(p, ix[0]);
                p += dist_to_next_recomb(recomb_dist);
            }
        } while (!(m.allow_no_recomb || all_chrom_recomb(slots)));
    }
    // Multivalent.java:216-220
    bool reject_dist(double dist) {
        return dist <= 0 || (m.chiasma_interference && dist < 0.5 &&
                             rng.next_double() < std::pow(1.0 - 2 * dist, 2.25));
    }

    // Quadrivalent.java:14-19, 71
    static constexpr int ARM[4][4] = {{0, 1, 6, 7}, {0, 1, 2, 3}, {2, 3, 4, 5}, {4, 5, 6, 7}};
    static constexpr int OPPO[2][2] = {{1, 3}, {0, 2}};

    // Quadrivalent.java:108-139
    void generate_recombination(int a, double recomb_dist) {
        chromatid[0] = ARM[a][rng.next_int(2)];
        chromatid[1] = ARM[a][2 + rng.next_int(2)];
        side = a % 2;
        side_fact = side == 0 ? 1.0 : -1.0;
        if (std::isnan(lastpos[a])) pos = chrom.side_pos(side) + side_fact * dist_to_first_recomb(recomb_dist);
        else pos = lastpos[a] + side_fact * dist_to_next_recomb(recomb_dist);
        mindist = std::nan("");
        nearest_arm = -1;
        if (!std::isnan(lastpos[OPPO[side][0]])) {
            mindist = side_fact * (lastpos[OPPO[side][0]] - pos);
            nearest_arm = OPPO[side][0];
        }
        if (!std::isnan(lastpos[OPPO[side][1]]) &&
            (std::isnan(mindist) || mindist > side_fact * (lastpos[OPPO[side][1]] - pos))) {
            mindist = side_fact * (lastpos[OPPO[side][1]] - pos);
            nearest_arm = OPPO[side][1];
        }
    }
    // Quadrivalent.java:141-243
    void test_recomb2(int a) {
        double temp_lim = exchange_lim[side];
        double dist_exchangelim = side_fact * (pos - exchange_lim[1 - side]);
        if (dist_exchangelim >= 0.0 || (!std::isnan(mindist) && reject_dist(mindist))) {
            arm_finished[a] = true;
            if (dist_exchangelim >= 0.0) {
                temp_lim = exchange_lim[1 - side];
                for (int oa = 0; oa < 2; oa++) {
                    double lpoa = lastpos[OPPO[side][oa]];
                    if (temp_lim == chrom.side_pos(1 - side) ||
                        (!std::isnan(lpoa) && (side_fact * (lpoa - temp_lim)) <= 0)) {
                        arm_finished[OPPO[side][oa]] = true;
                    }
                }
            } else {
                if (side_fact * (pos - temp_lim) > 0) temp_lim = pos;
                assert(nearest_arm > -1);
                arm_finished[nearest_arm] = true;
            }
            crossover_finished = true;
            for (int i = 0; i < 4; i++) crossover_finished = crossover_finished && arm_finished[i];
        } else {
            lastpos[a] = pos;
            if ((side_fact * (pos - temp_lim)) > 0) temp_lim = pos;
        }
        exchange_lim[side] = temp_lim;
    }
    // Quadrivalent.java:276-284
    void make_recombination(int a) {
        recomb_count++;
        slots[chromatid[0]].insert_segment(pos, chromatid[1]);
        slots[chromatid[1]].insert_segment(pos, chromatid[0]);
        lastpos[a] = pos;
    }
    // Quadrivalent.java:371-457 (quadriMethod == 2, quadriEachArm == false: PopulationData.java:206-207)
    void quadrivalent_crossing_over() {
        bool paral = m.parallel_multivalents == 0.0 ? false
                   : m.parallel_multivalents == 1.0 ? true
                   : rng.next_double() < m.parallel_multivalents;
        if (paral) { parallel_crossing_over(); return; }
        was_cross_type = true;
        do {
            double recomb_dist = RECOMBDIST;
            if (!m.allow_no_recomb) recomb_dist = calc_recomb_dist(chrom.length());
            slots.assign(8, HS(chrom.head, -1));
            for (int a = 0; a < 4; a++) { lastpos[a] = std::nan(""); arm_finished[a] = false; }
            crossover_finished = false;
            recomb_count = 0;
            exchange_lim[0] = chrom.head;
            exchange_lim[1] = chrom.tail;
            do {
                int a = rng.next_int(4);
                if (!arm_finished[a]) {
                    generate_recombination(a, recomb_dist);
                    test_recomb2(a);
                    if (!arm_finished[a]) make_recombination(a);
                }
            } while (!crossover_finished);
        } while (!(m.allow_no_recomb || all_chrom_recomb(slots)));
    }

    // Multivalent.java:332-357
    std::vector<HS> slots_to_patterns() const {
        const int slotcount = (int)slots.size();
        std::vector<HS> patterns(slotcount);
        for (int p = 0; p < slotcount; p++) {
            patterns[p] = HS(chrom.head, p / 2);
            int currslot = p;
            int64_t nextseg = 1;
            while (nextseg < slots[currslot].count()) {
                double ps = slots[currslot].pos[nextseg];
                int slt = slots[currslot].founder[nextseg];
                patterns[p].add_segment(ps, slt / 2);
                currslot = slt;
                nextseg = slots[currslot].find_segment(ps);
                if (slots[currslot].pos[nextseg] != ps) nextseg = INT32_MAX;
                nextseg++;
            }
        }
        return patterns;
    }
    // Multivalent.java:397-437
    std::vector<int> multivalent_centromere_order() {
        const int k = (int)hs.size();
        if (k == 2) {
            if (rng.next_bool()) return {0, 0, 1, 1};
            return {1, 1, 0, 0};
        }
        std::vector<int> chr = shuffle_array(rng, seq(k));
        std::vector<std::vector<int>> gam(4, std::vector<int>(k / 2));
        for (int p = 0; p < k / 2; p++) {
            gam[0][p] = chr[p]; gam[1][p] = chr[p];
            gam[2][p] = chr[p + k / 2]; gam[3][p] = chr[p + k / 2];
        }
        for (int g = 0; g < 4; g++) gam[g] = shuffle_array(rng, gam[g]);
        std::vector<int> centro;
        for (int g = 0; g < 4; g++) centro.insert(centro.end(), gam[g].begin(), gam[g].end());
        return centro;
    }
    // Quadrivalent.java:465-510
    std::vector<int> quadrivalent_centromere_order() {
        bool paired = m.paired_centromeres == 0.0 ? false
                    : m.paired_centromeres == 1.0 ? true
                    : rng.next_double() < m.paired_centromeres;
        if (!paired) return multivalent_centromere_order();
        std::vector<std::vector<int>> gam;
        if (rng.next_bool()) {
            bool centromeres_in_head_arm = chrom.centro <= exchange_lim[1];
            if (centromeres_in_head_arm) {
                if (rng.next_bool()) gam = {{0, 1}, {0, 1}, {2, 3}, {2, 3}};
                else gam = {{2, 3}, {2, 3}, {0, 1}, {0, 1}};
            } else {
                if (rng.next_bool()) gam = {{0, 3}, {0, 3}, {1, 2}, {1, 2}};
                else gam = {{1, 2}, {1, 2}, {0, 3}, {0, 3}};
            }
        } else {
            if (rng.next_bool()) gam = {{0, 2}, {0, 2}, {1, 3}, {1, 3}};
            else gam = {{1, 3}, {1, 3}, {0, 2}, {0, 2}};
        }
        for (int g = 0; g < 4; g++) gam[g] = shuffle_array(rng, gam[g]);
        std::vector<int> centro;
        for (int g = 0; g < 4; g++) centro.insert(centro.end(), gam[g].begin(), gam[g].end());
        return centro;
    }
    // HaploStruct.java:135-151 applied to a pattern (its "founder" is the chromosome-in-multivalent).
    int founder_at_centromere(const HS& h) const { return h.founder[h.find_segment(chrom.centro)]; }
    // Multivalent.java:539-577
    HS fill_gamete(const HS& pattern) const {
        if (pattern.count() == 1) return *hs[pattern.founder[0]];
        HS gamete;
        int patseg = 0;
        while (patseg < pattern.count()) {
            double patstart = pattern.pos[patseg];
            double patend = patseg < pattern.count() - 1 ? pattern.pos[patseg + 1] : chrom.tail;
            const HS& h = *hs[pattern.founder[patseg]];
            int hsseg = h.find_segment(patstart);
            gamete.pos.push_back(patstart);
            gamete.founder.push_back(h.founder[hsseg]);
            hsseg++;
            while (hsseg < h.count() && h.pos[hsseg] < patend) {
                gamete.pos.push_back(h.pos[hsseg]);
                gamete.founder.push_back(h.founder[hsseg]);
                hsseg++;
            }
            patseg++;
        }
        return gamete;
    }
    // Multivalent.java:486-527
    std::vector<HS> patterns_to_gametes(std::vector<HS> patterns, const std::vector<int>& centro) {
        const int n = (int)patterns.size();
        for (int s = 0; s < n - 1; s++) {
            if (founder_at_centromere(patterns[s]) != centro[s]) {
                int t = s + 1;
                while (founder_at_centromere(patterns[t]) != centro[s]) t++;
                std::swap(patterns[s], patterns[t]);
            }
        }
        if (n > 4) {
            for (int all = 0; all < 4; all++) {
                if (rng.next_bool()) {
                    int first = 0; while (centro[first] != all) first++;
                    int second = first + 1; while (centro[second] != all) second++;
                    std::swap(patterns[first], patterns[second]);
                }
            }
        }
        std::vector<HS> gametes(n);
        for (int s = 0; s < n; s++) gametes[s] = fill_gamete(patterns[s]);
        return gametes;
    }
    // Multivalent.java:79-82 — Java evaluates the arguments left to right: crossing-over draws,
    // then centromere draws, then (quadrivalents) the four chromatid-swap booleans.
    std::vector<HS> do_meiosis() {
        if (is_quad()) quadrivalent_crossing_over(); else parallel_crossing_over();
        std::vector<HS> patterns = slots_to_patterns();
        std::vector<int> centro = is_quad() ? quadrivalent_centromere_order() : multivalent_centromere_order();
        return patterns_to_gametes(std::move(patterns), centro);
    }
};

// ---------------------------------------------------------------- Individual.java
struct Individual {
    int parent[2] = {-1, -1};
    bool has_hs = false;
    std::vector<std::vector<HS>> hs;  // [chrom][homolog]
    bool has_parconfig = false;
    std::vector<ChromConfig> parconfig[2];  // [parent][chrom]
    // Individual.java:49-71, built lazily by calc_genomehs (cached per individual, :488)
    bool genome_ready = false;
    std::vector<std::vector<std::vector<int>>> genomenr;               // [chrom][side][h]
    std::vector<std::vector<std::vector<std::vector<int>>>> genomehs;  // [chrom][side][genome] -> homologs
    std::vector<std::vector<int>> possible_pair_count;                 // [chrom][side]
    bool is_founder() const { return parent[0] < 0; }
};

// Per-meiosis bookkeeping the statistical tests read (mirrors TEST-mode counters,
// TetraploidChromosome.java:54-58 and Individual.java:668).
struct MeiosisStats {
    std::vector<int64_t> quad_config_count;  // [#quadrivalents]
    int64_t bivalents = 0, parallel_quads = 0, cross_quads = 0;
};

// device_twin.hpp
struct Individual;
struct MeiosisStats;
inline void twin_do_meiosis(const Model& m, const Individual& parent, uint32_t child, uint32_t par,
                            std::vector<std::vector<HS>>& out_hs, std::vector<ChromConfig>& out_conf, int* second_gamete,
                            MeiosisStats* stats);

struct MeiosisEngine {
    const Model& m;
    Rng& rng;
    MeiosisStats* stats = nullptr;
    // Individual.java:553-567 scratch
    std::vector<int> initorder, backorder;
    int primaryside = 0, secondaryside = 0;
    std::vector<int> paired_with[2];
    int bivalentcount = 0, quadrivalentcount = 0;
    ChromConfig chrconf;
    int twin_second = 1;
    std::vector<bool> hsdone;
    std::vector<std::vector<int>> genohs;

    explicit MeiosisEngine(const Model& model) : m(model), rng(*model.rng) {}

    static int index_of(const std::vector<int>& v, int x) {
        for (size_t i = 0; i < v.size(); i++) if (v[i] == x) return (int)i;
        throw std::runtime_error("oracle: indexOf failed (ArrayList.remove(-1) in the reference)");
    }
    // Individual.java:437-470
    void calc_genomehs(Individual& ind) const {
        const int P = m.ploidy, C = (int)m.chrom.size();
        ind.genomenr.assign(C, {});
        ind.genomehs.assign(C, {});
        ind.possible_pair_count.assign(C, {0, 0});
        for (int chr = 0; chr < C; chr++) {
            ind.genomenr[chr].assign(2, std::vector<int>(P));
            ind.genomehs[chr].assign(2, std::vector<std::vector<int>>(P / 2));
            for (int side = 0; side < 2; side++) {
                for (int h = 0; h < P; h++) {
                    const HS& hs = ind.hs[chr][h];
                    double at = side == 0 ? m.chrom[chr].head : m.chrom[chr].tail;
                    int g = hs.founder[hs.find_segment(at)] % (P / 2);
                    ind.genomenr[chr][side][h] = g;
                    ind.genomehs[chr][side][g].push_back(h);
                }
                for (int g = 0; g < P / 2; g++)
                    ind.possible_pair_count[chr][side] += (int)ind.genomehs[chr][side][g].size() / 2;
            }
        }
        ind.genome_ready = true;
    }
    // Individual.java:574-593 — NB: removal uses genomenr at `secondaryside` whatever `side` is.
    void calc_genohs(const Individual& ind, int chr, int side) {
        genohs = ind.genomehs[chr][side];
        for (int h = 0; h < m.ploidy; h++) {
            if (hsdone[h]) {
                int gnm = ind.genomenr[chr][secondaryside][initorder[h]];
                int hidx = index_of(genohs[gnm], initorder[h]);
                genohs[gnm].erase(genohs[gnm].begin() + hidx);
            }
        }
    }
    // Individual.java:483-523
    std::vector<int> get_pairing(Individual& ind, int chr, int side) {
        const int P = m.ploidy;
        std::vector<int> pw(P, 0);
        std::vector<bool> done(P, false);
        if (!ind.genome_ready) calc_genomehs(ind);
        calc_genohs(ind, chr, side);
        int firsth = 0, secondh, gnm, hidx;
        while (firsth < P - 1) {
            while (firsth < P - 1 && done[firsth]) firsth++;
            if (firsth < P - 1) {
                done[firsth] = true;
                gnm = ind.genomenr[chr][side][initorder[firsth]];
                hidx = index_of(genohs[gnm], initorder[firsth]);
                genohs[gnm].erase(genohs[gnm].begin() + hidx);
                if (!genohs[gnm].empty() && rng.next_double() < m.chrom[chr].pref_pairing) {
                    hidx = rng.next_int((int)genohs[gnm].size());
                    secondh = backorder[genohs[gnm][hidx]];
                } else {
                    secondh = firsth;
                    while (done[secondh]) secondh = rng.next_int(P - firsth - 1) + firsth + 1;
                    gnm = ind.genomenr[chr][side][initorder[secondh]];
                    hidx = index_of(genohs[gnm], initorder[secondh]);
                }
                done[secondh] = true;
                genohs[gnm].erase(genohs[gnm].begin() + hidx);
                pw[firsth] = secondh;
                pw[secondh] = firsth;
                firsth++;
            }
        }
        return pw;
    }
    // Individual.java:536-542
    int get_primary_side(const Individual& ind, int chr) {
        int c0 = ind.possible_pair_count[chr][0], c1 = ind.possible_pair_count[chr][1];
        double probside0 = c0 == c1 ? 0.5 : (double)c0 / (double)(c0 + c1);
        return rng.next_double() < probside0 ? 0 : 1;
    }
    // Individual.java:672-677
    void add_bivalent(int firsth, int secondh) {
        bivalentcount++;
        chrconf.chromseq[m.ploidy - 2 * bivalentcount] = initorder[firsth];
        chrconf.chromseq[m.ploidy - 2 * bivalentcount + 1] = initorder[secondh];
    }
    // Individual.java:679-703
    void add_quadrivalent(int firsth, int secondh, int thirdh, int fourth) {
        int b = quadrivalentcount * 4;
        if (primaryside == 0) {
            chrconf.chromseq[b] = initorder[firsth];
            chrconf.chromseq[b + 1] = initorder[secondh];
            chrconf.chromseq[b + 2] = initorder[thirdh];
            chrconf.chromseq[b + 3] = initorder[fourth];
        } else {
            chrconf.chromseq[b] = initorder[secondh];
            chrconf.chromseq[b + 1] = initorder[thirdh];
            chrconf.chromseq[b + 2] = initorder[fourth];
            chrconf.chromseq[b + 3] = initorder[firsth];
        }
        quadrivalentcount++;
    }
    // Individual.java:705-725
    void get_spontaneous_bivalents(int max_bivalents) {
        const int P = m.ploidy;
        int firsth = 0;
        while (firsth < P - 1 && bivalentcount < max_bivalents) {
            while (firsth < P - 1 &&
                   (hsdone[firsth] || paired_with[primaryside][firsth] != paired_with[secondaryside][firsth]))
                firsth++;
            if (firsth < P - 1) {
                int secondh = paired_with[primaryside][firsth];
                add_bivalent(firsth, secondh);
                hsdone[firsth] = true;
                hsdone[secondh] = true;
            }
            firsth++;
        }
    }
    // Individual.java:727-787
    void get_spontaneous_quadrivalents(int max_quadrivalents) {
        const int P = m.ploidy;
        std::vector<int> tmp_bi;
        int firsth = 0, secondh;
        while (firsth < P - 1) {
            while (firsth < P - 1 &&
                   (hsdone[firsth] || paired_with[primaryside][firsth] != paired_with[secondaryside][firsth]))
                firsth++;
            if (firsth < P - 1) {
                secondh = paired_with[primaryside][firsth];
                hsdone[firsth] = true;
                hsdone[secondh] = true;
                tmp_bi.push_back(firsth);
                tmp_bi.push_back(secondh);
            }
            firsth++;
        }
        firsth = 0;
        int thirdh, fourth;
        while (firsth < P - 3 && quadrivalentcount < max_quadrivalents) {
            while (firsth < P - 3 && hsdone[firsth]) firsth++;
            if (firsth < P - 3) {
                secondh = paired_with[secondaryside][firsth];
                thirdh = paired_with[primaryside][secondh];
                fourth = paired_with[primaryside][firsth];
                if (paired_with[secondaryside][thirdh] == fourth) {
                    add_quadrivalent(firsth, secondh, thirdh, fourth);
                    hsdone[firsth] = true; hsdone[secondh] = true;
                    hsdone[thirdh] = true; hsdone[fourth] = true;
                }
            }
            firsth++;
        }
        for (int i : tmp_bi) hsdone[i] = false;
    }
    // Individual.java:789-799
    void get_forced_bivalent() {
        int firsth = 0;
        while (firsth < (int)hsdone.size() && hsdone[firsth]) firsth++;
        int secondh = paired_with[primaryside][firsth];
        add_bivalent(firsth, secondh);
        hsdone[firsth] = true;
        hsdone[secondh] = true;
    }
    // Individual.java:822-904
    void get_forced_bi_or_quadri(const Individual& ind, int chr, bool bi_allowed) {
        const int P = m.ploidy;
        int firsth = 0, secondh, thirdh, fourth;
        while (firsth < P - 1 && hsdone[firsth]) firsth++;
        fourth = paired_with[primaryside][firsth];
        int gnm = ind.genomenr[chr][secondaryside][initorder[firsth]];
        int gnm4 = ind.genomenr[chr][secondaryside][initorder[fourth]];
        if (genohs[gnm4].size() > genohs[gnm].size()) {
            std::swap(firsth, fourth);
            std::swap(gnm, gnm4);
        }
        int hidx = index_of(genohs[gnm], initorder[firsth]);
        genohs[gnm].erase(genohs[gnm].begin() + hidx);
        hsdone[firsth] = true;
        int hidx4 = index_of(genohs[gnm4], initorder[fourth]);
        if (((bi_allowed && genohs[gnm].size() > 0) || (!bi_allowed && genohs[gnm].size() > 1)) &&
            rng.next_double() < m.chrom[chr].pref_pairing) {
            do {
                hidx = rng.next_int((int)genohs[gnm].size());
                secondh = backorder[genohs[gnm][hidx]];
            } while (secondh == fourth && !bi_allowed);
            hsdone[fourth] = true;
            genohs[gnm4].erase(genohs[gnm4].begin() + hidx4);
            if (secondh == fourth) {
                add_bivalent(firsth, secondh);
            } else {
                gnm = ind.genomenr[chr][secondaryside][initorder[secondh]];
                hidx = index_of(genohs[gnm], initorder[secondh]);
                genohs[gnm].erase(genohs[gnm].begin() + hidx);
                hsdone[secondh] = true;
                thirdh = paired_with[primaryside][secondh];
                gnm = ind.genomenr[chr][secondaryside][initorder[thirdh]];
                hidx = index_of(genohs[gnm], initorder[thirdh]);
                genohs[gnm].erase(genohs[gnm].begin() + hidx);
                hsdone[thirdh] = true;
                add_quadrivalent(firsth, secondh, thirdh, fourth);
            }
        } else {
            secondh = firsth;
            while (hsdone[secondh] || (secondh == fourth && !bi_allowed)) secondh = rng.next_int(P);
            hsdone[fourth] = true;
            genohs[gnm4].erase(genohs[gnm4].begin() + hidx4);
            if (secondh == fourth) {
                add_bivalent(firsth, secondh);
                hsdone[secondh] = true;
            } else {
                thirdh = paired_with[primaryside][secondh];
                hsdone[secondh] = true;
                gnm = ind.genomenr[chr][secondaryside][initorder[secondh]];
                hidx = index_of(genohs[gnm], initorder[secondh]);
                genohs[gnm].erase(genohs[gnm].begin() + hidx);
                hsdone[thirdh] = true;
                gnm = ind.genomenr[chr][secondaryside][initorder[thirdh]];
                hidx = index_of(genohs[gnm], initorder[thirdh]);
                genohs[gnm].erase(genohs[gnm].begin() + hidx);
                add_quadrivalent(firsth, secondh, thirdh, fourth);
            }
        }
    }
    // TetraploidChromosome.java:122-128
    int ran_quadri(const Chrom& ch) {
        double p = rng.next_double();
        int i = 0;
        while (i < (int)ch.cumul_binomial.size() && ch.cumul_binomial[i] < p) i++;
        return i;
    }
    // Individual.java:601-670
    void calc_chrom_config(Individual& ind, int c) {
        const int P = m.ploidy;
        chrconf = ChromConfig();
        chrconf.chromseq.assign(P, 0);
        if (P == 2) {
            chrconf.chromseq = shuffle_array(rng, seq(2));
            chrconf.quad = 0;
            return;
        }
        const Chrom& chr = m.chrom[c];
        initorder = shuffle_array(rng, seq(P));
        backorder.assign(P, 0);
        for (int h = 0; h < P; h++) backorder[initorder[h]] = h;
        hsdone.assign(P, false);
        for (int side = 0; side < 2; side++) paired_with[side] = get_pairing(ind, c, side);
        primaryside = get_primary_side(ind, c);
        secondaryside = 1 - primaryside;
        bivalentcount = 0;
        quadrivalentcount = 0;
        if (m.natural_pairing) {
            get_spontaneous_bivalents(P / 2);
            get_spontaneous_quadrivalents((P - 2 * bivalentcount) / 4);
            calc_genohs(ind, c, secondaryside);
            while (2 * bivalentcount + 4 * quadrivalentcount < P) {
                if (2 * bivalentcount + 4 * quadrivalentcount == P - 2) get_forced_bivalent();
                else get_forced_bi_or_quadri(ind, c, true);
            }
        } else {
            int num_quad = ran_quadri(chr);
            int num_bi = (P - 4 * num_quad) / 2;
            get_spontaneous_bivalents(num_bi);
            get_spontaneous_quadrivalents(num_quad);
            while (bivalentcount < num_bi) get_forced_bivalent();
            if (quadrivalentcount < num_quad) {
                calc_genohs(ind, c, secondaryside);
                while (quadrivalentcount < num_quad) get_forced_bi_or_quadri(ind, c, false);
            }
        }
        chrconf.quad = quadrivalentcount;
        if (stats) {
            if ((int)stats->quad_config_count.size() <= chrconf.quad) stats->quad_config_count.resize(chrconf.quad + 1, 0);
            stats->quad_config_count[chrconf.quad]++;
        }
    }

    // Individual.java:281-349 — returns hs[c][0 .. 2P-1] (4 gametes of P/2 homologs each) and
    // the per-chromosome configurations. `ind_index`/`parent_slot` only select Philox streams.
    void do_meiosis(Individual& ind, uint32_t child_index, uint32_t parent_slot,
                    std::vector<std::vector<HS>>& out_hs, std::vector<ChromConfig>& out_conf) {
        const int P = m.ploidy, C = (int)m.chrom.size(), p2 = P / 2;
        if (m.twin) { twin_do_meiosis(m, ind, child_index, parent_slot, out_hs, out_conf, &twin_second, stats); return; }
        out_hs.assign(C, std::vector<HS>(2 * P));
        out_conf.assign(C, ChromConfig());
        for (int c = 0; c < C; c++) {
            rng.seek(child_index, (uint32_t)c, parent_slot, 0);
            calc_chrom_config(ind, c);
            out_conf[c] = chrconf;
            uint32_t stage = 1;
            for (int v = 0; v < chrconf.quad; v++) {
                std::vector<const HS*> orig(4);
                for (int h = 0; h < 4; h++) orig[h] = &ind.hs[c][chrconf.chromseq[4 * v + h]];
                rng.seek(child_index, (uint32_t)c, parent_slot, stage++);
                Multivalent mv(m, m.chrom[c], orig);
                std::vector<HS> qh = mv.do_meiosis();
                if (stats) { if (mv.was_cross_type) stats->cross_quads++; else stats->parallel_quads++; }
                for (int g = 0; g < 4; g++) {
                    out_hs[c][2 * v + p2 * g] = qh[2 * g];
                    out_hs[c][2 * v + p2 * g + 1] = qh[2 * g + 1];
                }
            }
            int firstbi = chrconf.quad * 4;
            if (firstbi < P) {
                for (int v = 0; v < (P - firstbi) / 2; v++) {
                    std::vector<const HS*> orig(2);
                    for (int h = 0; h < 2; h++) orig[h] = &ind.hs[c][chrconf.chromseq[firstbi + 2 * v + h]];
                    rng.seek(child_index, (uint32_t)c, parent_slot, stage++);
                    Multivalent mv(m, m.chrom[c], orig);
                    std::vector<HS> dh = mv.do_meiosis();
                    if (stats) stats->bivalents++;
                    for (int g = 0; g < 4; g++) out_hs[c][firstbi / 2 + v + p2 * g] = dh[g];
                }
            }
            // Individual.java:325-335 — shuffle the P/2 homolog slots inside each of the 4 gametes.
            rng.seek(child_index, (uint32_t)c, parent_slot, 1023);
            for (int g = 0; g < 4; g++) {
                std::vector<int> shuf = shuffle_array(rng, seq(g * p2, (g + 1) * p2 - 1));
                std::vector<HS> shufhs(p2);
                for (int h = 0; h < p2; h++) shufhs[h] = out_hs[c][shuf[h]];
                for (int h = 0; h < p2; h++) out_hs[c][g * p2 + h] = shufhs[h];
            }
        }
    }

    // Individual.java:255-261 + Gamete.java:52-72,95-116: gamete 0 of each parent, parent 0 first.
    void conception(std::vector<Individual>& pop, int child) {
        const int P = m.ploidy, C = (int)m.chrom.size(), p2 = P / 2;
        Individual& z = pop[child];
        std::vector<std::vector<HS>> g0, g1;
        std::vector<ChromConfig> c0, c1;
        do_meiosis(pop[z.parent[0]], (uint32_t)child, 0, g0, c0);
        do_meiosis(pop[z.parent[1]], (uint32_t)child, 1, g1, c1);
        z.hs.assign(C, std::vector<HS>(P));
        for (int p = 0; p < p2; p++)
            for (int c = 0; c < C; c++) {
                z.hs[c][p] = g0[c][p];
                z.hs[c][p2 + p] = g1[c][p];
            }
        z.has_hs = true;
        z.parconfig[0] = c0;
        z.parconfig[1] = c1;
        z.has_parconfig = true;
    }

    // Individual.java:341-383 — the FIRST unreduced (2n) gamete of a meiosis: P homologs per
    // chromosome taken from the 4 x P/2 homologs do_meiosis leaves.
    //   FDR (:350-372): gamete 0 ("top") + one of the two "bottom" gametes, 2 unless nextFloat() < 0.5
    //                   (one draw per meiosis, after all chromosomes);
    //   SDR (:373-382): gametes 0 and 1 (the second division undone).
    // The reference only analyses such gametes (TEST mode, Main.testMeiosis); here an "individual"
    // is formed from the one 2n gamete of its parent 0 alone, so that the emission and statistics
    // paths see it. parconfig[1] stays missing. In Philox mode the FDR draw has its own stream
    // (chromosome 0, stage 1022).
    void unreduced_conception(std::vector<Individual>& pop, int child) {
        const int P = m.ploidy, C = (int)m.chrom.size(), p2 = P / 2;
        Individual& z = pop[child];
        std::vector<std::vector<HS>> g;
        std::vector<ChromConfig> c0;
        do_meiosis(pop[z.parent[0]], (uint32_t)child, 0, g, c0);
        int second = 1;
        if (m.twin) second = twin_second;
        else if (m.unreduced_gametes == 1) {
            rng.seek((uint32_t)child, 0, 0, 1022);
            second = rng.next_float() < 0.5f ? 3 : 2;
        }
        z.hs.assign(C, std::vector<HS>(P));
        for (int c = 0; c < C; c++)
            for (int h = 0; h < p2; h++) {
                z.hs[c][h] = g[c][h];
                z.hs[c][p2 + h] = g[c][second * p2 + h];
            }
        z.has_hs = true;
        ChromConfig none;
        none.quad = -1;
        none.chromseq.assign(P, -1);
        z.parconfig[0] = c0;
        z.parconfig[1].assign(C, none);
        z.has_parconfig = true;
    }

    // Main.java:1504-1520 + Individual.java:155-171
    void simulate_individuals(std::vector<Individual>& pop, int maxfounderallele) {
        const int P = m.ploidy, C = (int)m.chrom.size();
        for (size_t i = 0; i < pop.size(); i++) {
            Individual& ind = pop[i];
            if (ind.has_hs) continue;
            if (ind.is_founder()) {
                ind.hs.assign(C, std::vector<HS>(P));
                int last = maxfounderallele;
                for (int p = 0; p < P; p++) {
                    last++;
                    for (int c = 0; c < C; c++) ind.hs[c][p] = HS(m.chrom[c].head, last);
                }
                ind.has_hs = true;
                maxfounderallele += P;
            } else if (m.unreduced_gametes != 0) {
                unreduced_conception(pop, (int)i);
            } else {
                conception(pop, (int)i);
            }
        }
    }
};

}  // namespace orc
