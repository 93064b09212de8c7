// Device-side meiosis of one parent on one chromosome: the pairing configuration, the chiasmata of
// one multivalent, the choice of the transmitted chromatids and their trace.
//
// What is modelled is PedigreeSim's meiosis (paths relative to /root/reference/src/PedigreeSim/):
//   pairing of the homolog ends, cycles -> bivalents / quadrivalents   Individual.java:483-904
//   chiasmata of a bivalent / parallel quadrivalent                    Multivalent.java:84-182
//   chiasmata of a cross-type quadrivalent (four arms)                 Quadrivalent.java:108-243, 371-424
//   centromere segregation, the transmitted gamete                     Multivalent.java:397-527, Quadrivalent.java:465-510
// How it is computed is this library's own: everything lives in position space (a random order of
// the homologs) as bit masks and 4-bit vectors in registers, every random choice is sampled directly
// from its distribution (a free partner, a non-sister chromatid, an unfinished arm: one bounded draw,
// no rejection loop, no list surgery), the stationary start of the interference (gamma) renewal
// process is drawn from its closed form instead of a burn-in, and only the chromatids that reach the
// transmitted gamete are traced (the reference builds all 4 gametes and keeps one,
// Individual.java:258-259). Outcome laws are the reference's: tests/test_gpu_stats.py and
// tests/test_sampler_laws.py hold them (chi-square) against the CPU oracle's restatement of the Java
// driven by java.util.Random. Everything here is PSIM_HD (host + device) so that the tests can run the
// very same sampler on the host and hold the kernels against it bit for bit.
#pragma once
#include <cmath>
#include <cstdint>

#include "psim_rng.cuh"

namespace psim {

constexpr int MAXP = 16;             // ploidy limit of the 4-bit position vectors
constexpr double RECOMBDIST = 0.5;   // mean chiasma spacing on a 4-strand bundle, Morgan (Multivalent.java:18)
constexpr double GAMMAFACTOR = 2.63; // shape of the inter-chiasma gamma under interference (Tools.java:317)

// Model flags + per-chromosome constants, passed by value to the kernels.
struct ModelParams {
    int ploidy;
    int chiasma_interference;
    int natural_pairing;
    int allow_no_recomb;
    double parallel_multivalents;
    double paired_centromeres;
    uint64_t seed;
    int unreduced;   // 2NGAMETES: 0 = haploid gametes, 1 = FDR, 2 = SDR (Individual.java:341-383)
};
struct ChromParams {
    double head, tail, centro;
    double pref_pairing;
    double recomb_dist_noempty;  // Tools.calcRecombDist(length), used when allow_no_recomb == 0
    double cumul_binomial[MAXP / 4];  // TetraploidChromosome.java:102-108
};

// Read-only view of one parental homolog (segments in the slab).
struct HomologView {
    const double* pos;
    const int32_t* founder;
    int n;
    // HaploStruct.findSegment (HaploStruct.java:107-112): last segment that starts at or before `position`
    PSIM_HD int find_segment(double position) const {
        int seg = n - 1;
        while (seg > 0 && pos[seg] > position) seg--;
        return seg;
    }
};

// Sixteen 4-bit entries in one register pair: order of the homologs, partners, genome numbers.
struct Nib16 {
    uint64_t v;
    PSIM_HD int get(int i) const { return (int)((v >> (4 * i)) & 15u); }
    PSIM_HD void set(int i, int x) { v = (v & ~(15ull << (4 * i))) | ((uint64_t)x << (4 * i)); }
    PSIM_HD void swap(int i, int j) { const int a = get(i), b = get(j); set(i, b); set(j, a); }
};
PSIM_HD Nib16 nib_identity() { Nib16 r; r.v = 0xFEDCBA9876543210ull; return r; }
PSIM_HD int pick_bit(PhiloxStream& rng, uint32_t mask) {   // a uniformly chosen member of a non-empty set
    const int n = popc32(mask);
    return n == 1 ? lowest_bit(mask) : nth_set_bit(mask, rng.next_below(n));
}
PSIM_HD void shuffle_nib(PhiloxStream& rng, Nib16& a, int n) {   // uniform random permutation of entries 0..n-1
    for (int last = n - 1; last > 0; last--) a.swap(last, rng.next_below(last + 1));
}

// ---------------------------------------------------------------- distances between chiasmata
// (the logarithm is a CALL on the device, by value like philox4x32_10: the exponential draw is inlined at every gap of
// every kind of multivalent, and fp64 log is ~100 instructions of a kernel that waits on its instruction cache)
#if defined(__CUDACC__) && !defined(PSIM_LOG_INLINE)
__host__ __device__ __noinline__
#else
PSIM_HD
#endif
double log_called(double x) { return log(x); }
PSIM_HD double draw_exponential(PhiloxStream& r, double mean) { return -mean * log_called(r.next_open()); }
// Gamma(k, theta) for k >= 1 by Cheng's GA rejection sampler (the algorithm Tools.ranGamma uses, Tools.java:294-314)
PSIM_HD_NOINLINE double draw_gamma(PhiloxStream& rng, double k, double theta) {
    const double lam = sqrt(2.0 * k - 1.0);
    const double b = k - 1.3862943611198906;     // k - ln 4
    const double c = k + lam;
    for (;;) {
        const double u = rng.next_open(), v = rng.next_unit();
        if (v <= 0.0) continue;
        const double y = log(v / (1.0 - v)) / lam;
        const double x = k * exp(y);
        const double z = u * v * v;
        const double r = b + c * y - x;
        if (r + 2.504077396776274 >= 4.5 * z || r >= log(z)) return x * theta;   // 1 + ln 4.5
    }
}
// Distance from a chromosome end (or the previous chiasma) to the next chiasma of a renewal process
// with mean spacing `mean`: exponential gaps without interference (Haldane), gamma(2.63) gaps with
// (Kosambi, Multivalent.java:84-105). FIRST: the process is stationary at the chromosome end. The
// reference gets there by a burn-in of ten mean spacings; the forward recurrence time of a
// stationary renewal process has the closed form U * G with G the length-biased gap,
// gamma(2.63 + 1), which is drawn here directly.
template <bool KOS, bool FIRST>
PSIM_HD double draw_gap(PhiloxStream& r, double mean) {
    if (!KOS) return draw_exponential(r, mean);
    if (FIRST) return r.next_unit() * draw_gamma(r, GAMMAFACTOR + 1.0, mean / GAMMAFACTOR);
    return draw_gamma(r, GAMMAFACTOR, mean / GAMMAFACTOR);
}

// ---------------------------------------------------------------- pairing configuration
// Input: the genome number (founder allele % (P/2), Individual.java:456-459) of every homolog at the
// head and at the tail of the chromosome. Output: Genotype.ChromConfig (Genotype.java:38-51) - the
// number of quadrivalents and the order of the homologs (quadrivalents first, bivalents filled from
// the end backwards).
struct PairingResult {
    Nib16 seq;
    int quad;
};

struct PairingWork {
    int P;
    Nib16 order;          // order.get(position) = homolog
    Nib16 genpos[2];      // genome number by position, at the head / tail
    Nib16 partner[2];     // partner position at the head / tail
    uint32_t done;        // positions already placed in a bivalent / quadrivalent
    int primary, n_bi, n_quad;
    Nib16 seq;

    PSIM_HD uint32_t genome_mask(int side, int g) const {   // positions whose end `side` belongs to genome g
        uint32_t m = 0;
        for (int p = 0; p < P; p++) m |= (uint32_t)(genpos[side].get(p) == g) << p;
        return m;
    }
    // Pairing of one chromosome end (Individual.getPairing, Individual.java:483-523): the lowest free
    // position takes, with probability `pref` and if one is free, a partner of its own genome,
    // otherwise any free position; uniformly in both cases.
    PSIM_HD_NOINLINE void pair_end(PhiloxStream& rng, int side, double pref) {
        uint32_t free_mask = (1u << P) - 1u;
        Nib16 pw; pw.v = 0;
        while (free_mask) {
            const int first = lowest_bit(free_mask);
            free_mask &= free_mask - 1;
            const uint32_t same = genome_mask(side, genpos[side].get(first)) & free_mask;
            const int second = (same && rng.next_unit() < pref) ? pick_bit(rng, same) : pick_bit(rng, free_mask);
            free_mask &= ~(1u << second);
            pw.set(first, second);
            pw.set(second, first);
        }
        partner[side] = pw;
    }
    PSIM_HD void put_bivalent(int a, int b) {   // Individual.addBivalent, :672-677
        n_bi++;
        seq.set(P - 2 * n_bi, order.get(a));
        seq.set(P - 2 * n_bi + 1, order.get(b));
        done |= (1u << a) | (1u << b);
    }
    // Individual.addQuadrivalent, :679-703: a-b and c-d are paired at the secondary end, a-d and b-c
    // at the primary end; stored so that 1st-4th & 2nd-3rd pair at the head, 1st-2nd & 3rd-4th at the tail
    PSIM_HD void put_quadrivalent(int a, int b, int c, int d) {
        const int base = 4 * n_quad;
        if (primary == 0) { seq.set(base, order.get(a)); seq.set(base + 1, order.get(b)); seq.set(base + 2, order.get(c)); seq.set(base + 3, order.get(d)); }
        else { seq.set(base, order.get(b)); seq.set(base + 1, order.get(c)); seq.set(base + 2, order.get(d)); seq.set(base + 3, order.get(a)); }
        n_quad++;
        done |= (1u << a) | (1u << b) | (1u << c) | (1u << d);
    }
    PSIM_HD uint32_t two_cycles() const {   // free positions with the same partner at both ends
        uint32_t m = 0;
        for (int p = 0; p < P; p++) m |= (uint32_t)(partner[0].get(p) == partner[1].get(p)) << p;
        return m & ~done;
    }
    // pairs that chose each other at both ends (getSpontaneousBivalents, :705-725), lowest position first
    PSIM_HD_NOINLINE void take_two_cycles(int limit) {
        uint32_t cand = two_cycles();
        while (cand && n_bi < limit) {
            const int a = lowest_bit(cand), b = partner[primary].get(a);
            put_bivalent(a, b);
            cand &= ~done;
        }
    }
    // closed rings of four (getSpontaneousQuadrivalents, :727-787); pairs of a two-cycle are not rings
    PSIM_HD_NOINLINE void take_four_cycles(int limit) {
        uint32_t cand = ~(done | two_cycles()) & ((1u << P) - 1u);
        const int sec = 1 - primary;
        while (cand && n_quad < limit) {
            const int a = lowest_bit(cand);
            cand &= cand - 1;
            if (a >= P - 3) break;
            const int b = partner[sec].get(a), c = partner[primary].get(b), d = partner[primary].get(a);
            if (partner[sec].get(c) == d) { put_quadrivalent(a, b, c, d); cand &= ~done; }
        }
    }
    PSIM_HD void force_bivalent() {   // getForcedBivalent, :789-799: the lowest free position and its primary partner
        const int a = lowest_bit(~done);
        put_bivalent(a, partner[primary].get(a));
    }
    // A ring longer than four is cut (getForcedBiOrQuadri, :822-904): the lowest free position a and its
    // primary partner d (the one with more free positions of its own genome at the secondary end leads)
    // get a new secondary partner b for a - of a's genome with probability `pref` when one is
    // available, else any free position - and close either as the bivalent a-d (b == d, if allowed)
    // or as the quadrivalent a, b, c = primary partner of b, d.
    PSIM_HD_NOINLINE void force_cut(PhiloxStream& rng, bool bivalent_allowed, double pref) {
        const int sec = 1 - primary;
        const uint32_t all = (1u << P) - 1u;
        int a = lowest_bit(~done), d = partner[primary].get(a);
        {
            const int na = popc32(genome_mask(sec, genpos[sec].get(a)) & ~done), nd = popc32(genome_mask(sec, genpos[sec].get(d)) & ~done);
            if (nd > na) { const int t = a; a = d; d = t; }
        }
        done |= 1u << a;
        const uint32_t free_mask = ~done & all;                                    // still holds d
        const uint32_t own = genome_mask(sec, genpos[sec].get(a)) & free_mask;     // a's genome, may hold d
        const uint32_t banned = bivalent_allowed ? 0u : (1u << d);
        int b;
        if ((bivalent_allowed ? own != 0 : popc32(own) > 1) && rng.next_unit() < pref) b = pick_bit(rng, own & ~banned);
        else b = pick_bit(rng, free_mask & ~banned);
        if (b == d) { done &= ~(1u << a); put_bivalent(a, d); }
        else { done &= ~(1u << a); put_quadrivalent(a, b, partner[primary].get(b), d); }
    }
};

// Individual.calcChromConfig (:601-670) for ploidy > 2.
PSIM_HD_NOINLINE PairingResult draw_pairing(PhiloxStream& rng, const ModelParams& m, const ChromParams& ch, Nib16 genome_head, Nib16 genome_tail) {
    PairingWork w;
    const int P = m.ploidy;
    if (P == 2) {   // one bivalent; which homolog comes first is a coin (Individual.java:603-608)
        PairingResult r2;
        r2.seq = nib_identity();
        shuffle_nib(rng, r2.seq, 2);
        r2.quad = 0;
        return r2;
    }
    w.P = P;
    w.order = nib_identity();
    shuffle_nib(rng, w.order, P);
    w.genpos[0].v = 0; w.genpos[1].v = 0;
    for (int p = 0; p < P; p++) { w.genpos[0].set(p, genome_head.get(w.order.get(p))); w.genpos[1].set(p, genome_tail.get(w.order.get(p))); }
    w.pair_end(rng, 0, ch.pref_pairing);
    w.pair_end(rng, 1, ch.pref_pairing);
    {   // the end with more same-genome pairs to offer is more often the primary one (getPrimarySide, :536-542)
        int c0 = 0, c1 = 0;
        for (int g = 0; g < P / 2; g++) { c0 += popc32(w.genome_mask(0, g)) >> 1; c1 += popc32(w.genome_mask(1, g)) >> 1; }
        const double p0 = c0 == c1 ? 0.5 : (double)c0 / (double)(c0 + c1);
        w.primary = rng.next_unit() < p0 ? 0 : 1;
    }
    w.done = 0; w.n_bi = 0; w.n_quad = 0; w.seq.v = 0;
    if (m.natural_pairing) {
        w.take_two_cycles(P / 2);
        w.take_four_cycles((P - 2 * w.n_bi) / 4);
        while (2 * w.n_bi + 4 * w.n_quad < P) {
            if (2 * w.n_bi + 4 * w.n_quad == P - 2) w.force_bivalent();
            else w.force_cut(rng, true, ch.pref_pairing);
        }
    } else {
        // number of quadrivalents ~ Binomial(P/4, fraction) (TetraploidChromosome.ranQuadri, :122-128)
        const double u = rng.next_unit();
        int want_quad = 0;
        while (want_quad < P / 4 && ch.cumul_binomial[want_quad] < u) want_quad++;
        const int want_bi = (P - 4 * want_quad) / 2;
        w.take_two_cycles(want_bi);
        w.take_four_cycles(want_quad);
        while (w.n_bi < want_bi) w.force_bivalent();
        while (w.n_quad < want_quad) w.force_cut(rng, false, ch.pref_pairing);
    }
    PairingResult r;
    r.seq = w.seq;
    r.quad = w.n_quad;
    return r;
}

// ---------------------------------------------------------------- chiasmata of one multivalent
// All chiasmata of a multivalent as one position-sorted list in per-thread global scratch
// (entry e of thread slot t at [e * stride + t], like local memory but sized at run time from the
// longest chromosome): position and the two chromatid strands that exchange (strand s belongs to
// chromosome s / 2 of the multivalent).
struct ChiasmaList {
    double* pos;
    uint8_t* ab;
    int64_t stride;
    int cap, n;
    bool overflow;
    PSIM_HD double p(int e) const { return pos[e * stride]; }
    PSIM_HD int a(int e) const { return ab[e * stride] >> 4; }
    PSIM_HD int b(int e) const { return ab[e * stride] & 15; }
    PSIM_HD void append(double x, int sa, int sb) {
        if (n >= cap) { overflow = true; return; }
        pos[n * stride] = x; ab[n * stride] = (uint8_t)((sa << 4) | sb); n++;
    }
    // keeps the list sorted; among equal positions the later chiasma comes last (HaploStruct.insertSegment, :129-133)
    PSIM_HD void insert(double x, int sa, int sb) {
        if (n >= cap) { overflow = true; return; }
        int i = n;
        while (i > 0 && pos[(i - 1) * stride] > x) { pos[i * stride] = pos[(i - 1) * stride]; ab[i * stride] = ab[(i - 1) * stride]; i--; }
        pos[i * stride] = x; ab[i * stride] = (uint8_t)((sa << 4) | sb); n++;
    }
    PSIM_HD uint32_t chromosomes_hit() const {
        uint32_t seen = 0;
        for (int e = 0; e < n; e++) seen |= (1u << (a(e) >> 1)) | (1u << (b(e) >> 1));
        return seen;
    }
};

// Bivalent (K = 2) or parallel quadrivalent (K = 4): one renewal process along the bundle of 2K
// chromatids, every chiasma between two chromatids of different chromosomes chosen uniformly
// (Multivalent.doCrossingOver, Multivalent.java:139-182).
template <int K, bool KOS, bool ANR>
PSIM_HD_NOINLINE void chiasmata_parallel(ChiasmaList& x, PhiloxStream& rng, const ChromParams& ch) {
    double mean = RECOMBDIST * 2.0 / K;
    if (!ANR) mean = mean / 0.5 * ch.recomb_dist_noempty;
    for (;;) {
        x.n = 0;
        double at = ch.head + draw_gap<KOS, true>(rng, mean);
        while (at < ch.tail) {
            int s0, s1;
            if (K == 2) {
                const uint32_t bits = rng.next_bits(3);
                s0 = (int)(bits >> 1);                        // any of the four chromatids
                s1 = (int)((((bits >> 2) ^ 1u) << 1) | (bits & 1u));   // either chromatid of the other chromosome
            } else {
                s0 = (int)rng.next_bits(3);
                const int r = rng.next_below(6);              // the six chromatids of the three other chromosomes
                s1 = r + (r >= (s0 & ~1) ? 2 : 0);
            }
            x.append(at, s0, s1);
            if (x.overflow) return;
            at += draw_gap<KOS, false>(rng, mean);
        }
        if (ANR || x.chromosomes_hit() == (1u << K) - 1u) return;   // ALLOWNORECOMBINATION=0: every chromosome exchanges at least once
    }
}

// Cross-type quadrivalent (Quadrivalent.doCrossingOver with quadriMethod 2, Quadrivalent.java:371-424 and
// generateRecombination / testRecomb2 / makeRecombination, :108-284). Chromosomes 0-3 and 1-2 are
// paired at the head, 0-1 and 2-3 at the tail: four arms, arm a growing from the head (a even) or the
// tail (a odd) towards the other end, each a renewal process of its own. A candidate chiasma of an
// arm is refused when it passes the point the arms of the other end have reached (the exchange limit)
// or meets their last chiasma (with interference: comes too close to it); a refusal finishes the arm
// and, depending on the case, the arms it ran into. Returns the exchange limit of the tail end.
template <bool KOS, bool ANR>
PSIM_HD_NOINLINE double chiasmata_cross(ChiasmaList& x, PhiloxStream& rng, const ChromParams& ch) {
    const double mean = ANR ? RECOMBDIST : ch.recomb_dist_noempty;
    double lim_head, lim_tail;
    for (;;) {
        x.n = 0;
        double last0 = 0.0, last1 = 0.0, last2 = 0.0, last3 = 0.0;   // last chiasma of each arm
        uint32_t has = 0, open = 0xFu;                              // arms that have a chiasma / can still get one
        lim_head = ch.head;
        lim_tail = ch.tail;
        while (open) {
            const int a = pick_bit(rng, open);
            const uint32_t bits = rng.next_bits(2);
            // arm 0 joins chromosomes 0 and 3, arm a > 0 chromosomes a-1 and a: one chromatid of each
            const int s0 = 2 * (a == 0 ? 0 : a - 1) + (int)(bits >> 1), s1 = 2 * (a == 0 ? 3 : a) + (int)(bits & 1u);
            const bool from_tail = a & 1;
            const double dir = from_tail ? -1.0 : 1.0;
            const double mine = a == 0 ? last0 : a == 1 ? last1 : a == 2 ? last2 : last3;
            const double at = (has >> a) & 1u ? mine + dir * draw_gap<KOS, false>(rng, mean)
                                              : (from_tail ? ch.tail : ch.head) + dir * draw_gap<KOS, true>(rng, mean);
            // the two arms growing from the other end, and how far ahead of `at` their nearest last chiasma lies
            const int o0 = from_tail ? 0 : 1, o1 = o0 + 2;
            const double lo0 = from_tail ? last0 : last1, lo1 = from_tail ? last2 : last3;
            bool met = false;
            double ahead = 0.0;
            int nearest = -1;
            if ((has >> o0) & 1u) { ahead = dir * (lo0 - at); nearest = o0; met = true; }
            if (((has >> o1) & 1u) && (!met || ahead > dir * (lo1 - at))) { ahead = dir * (lo1 - at); nearest = o1; met = true; }
            const double my_lim = from_tail ? lim_tail : lim_head, other_lim = from_tail ? lim_head : lim_tail;
            const bool passed = dir * (at - other_lim) >= 0.0;
            bool refused = passed;
            if (!refused && met)   // Multivalent.rejectDist, Multivalent.java:216-220
                refused = ahead <= 0.0 || (KOS && ahead < 0.5 && rng.next_unit() < pow(1.0 - 2.0 * ahead, 2.25));
            double new_lim = my_lim;
            if (refused) {
                open &= ~(1u << a);
                if (passed) {
                    // this end's limit jumps to the other end's; arms of the other end that stand at or behind it
                    // (or all of them while that limit is still the chromosome end) are finished too
                    new_lim = other_lim;
                    const bool at_end = other_lim == (from_tail ? ch.head : ch.tail);
                    if (at_end || (((has >> o0) & 1u) && dir * (lo0 - other_lim) <= 0.0)) open &= ~(1u << o0);
                    if (at_end || (((has >> o1) & 1u) && dir * (lo1 - other_lim) <= 0.0)) open &= ~(1u << o1);
                } else {
                    if (dir * (at - my_lim) > 0.0) new_lim = at;
                    open &= ~(1u << nearest);
                }
            } else {
                if (dir * (at - my_lim) > 0.0) new_lim = at;
                if (a == 0) last0 = at; else if (a == 1) last1 = at; else if (a == 2) last2 = at; else last3 = at;
                has |= 1u << a;
                x.insert(at, s0, s1);
                if (x.overflow) return lim_tail;
            }
            if (from_tail) lim_tail = new_lim; else lim_head = new_lim;
        }
        if (ANR || x.chromosomes_hit() == 0xFu) return lim_tail;
    }
}

// ---------------------------------------------------------------- segregation
// Which chromosome each of the 2K chromatids (numbered by the strand they start on at the head)
// carries at the centromere: the strands are walked once, swapping the two chromatids of every
// chiasma at or before the centromere (Multivalent.slotsToPatterns :332-357 + getFounderAtCentromere).
template <int K>
PSIM_HD uint32_t centromere_chromosomes(const ChiasmaList& x, double centro) {
    uint32_t on_strand = 0x76543210u;   // nibble s = chromatid currently on strand s
    for (int e = 0; e < x.n && x.p(e) <= centro; e++) {
        const int a = x.a(e), b = x.b(e);
        const uint32_t ca = (on_strand >> (4 * a)) & 15u, cb = (on_strand >> (4 * b)) & 15u;
        on_strand = (on_strand & ~((15u << (4 * a)) | (15u << (4 * b)))) | (cb << (4 * a)) | (ca << (4 * b));
    }
    uint32_t chrom_of = 0;              // nibble c = chromosome chromatid c carries at the centromere
    for (int s = 0; s < 2 * K; s++) chrom_of |= (uint32_t)(s >> 1) << (4 * ((on_strand >> (4 * s)) & 15u));
    return chrom_of;
}
PSIM_HD uint32_t nib8_swap(uint32_t v, int i, int j) {
    const uint32_t a = (v >> (4 * i)) & 15u, b = (v >> (4 * j)) & 15u;
    return (v & ~((15u << (4 * i)) | (15u << (4 * j)))) | (b << (4 * i)) | (a << (4 * j));
}
// The centromeres the gametes receive: nibble 2g, 2g+1 (K = 4) or g (K = 2) = chromosome whose centromere goes to
// gamete g. Bivalent: first division separates the two chromosomes (Multivalent.getCentromereSortOrder :397-437).
PSIM_HD uint32_t centromeres_bivalent(PhiloxStream& rng) { return rng.next_bits(1) ? 0x1100u : 0x0011u; }
// Quadrivalent: two chromosomes go to each pole - a random two without centromere pairing (Multivalent.java:407-436),
// with pairing adjacent or alternate segregation of the ring (Quadrivalent.getCentromereSortOrder, :465-510); the two
// centromeres of a gamete in random order.
PSIM_HD_NOINLINE uint32_t centromeres_quadrivalent(PhiloxStream& rng, const ModelParams& m, const ChromParams& ch, double lim_tail) {
    const bool paired = m.paired_centromeres == 0.0 ? false : m.paired_centromeres == 1.0 ? true : rng.next_unit() < m.paired_centromeres;
    int top0, top1, bot0, bot1;
    if (!paired) {
        Nib16 c = nib_identity();
        shuffle_nib(rng, c, 4);
        top0 = c.get(0); top1 = c.get(1); bot0 = c.get(2); bot1 = c.get(3);
    } else {
        const uint32_t bits = rng.next_bits(2);
        const bool flip = bits & 1u;
        if (bits >> 1) {   // adjacent: the partners of the arm that holds the centromeres go together
            if (ch.centro <= lim_tail) { top0 = 0; top1 = 1; bot0 = 2; bot1 = 3; }
            else { top0 = 0; top1 = 3; bot0 = 1; bot1 = 2; }
        } else { top0 = 0; top1 = 2; bot0 = 1; bot1 = 3; }   // alternate
        if (!flip) { int t = top0; top0 = bot0; bot0 = t; t = top1; top1 = bot1; bot1 = t; }
    }
    const uint32_t coins = rng.next_bits(4);
    uint32_t out = 0;
    for (int g = 0; g < 4; g++) {
        const int u = g < 2 ? top0 : bot0, v = g < 2 ? top1 : bot1;
        const bool sw = (coins >> g) & 1u;
        out |= (uint32_t)(sw ? v : u) << (8 * g) | (uint32_t)(sw ? u : v) << (8 * g + 4);
    }
    return out;
}
// Chromatids to gametes (Multivalent.patternsToGametes :486-518): position s of the result takes the first chromatid
// not yet placed whose centromere is the one position s receives; in a quadrivalent the two chromatids of every
// centromere then change places with probability 1/2. Returns nibble s = chromatid at position s
// (gamete g = position g of a bivalent, positions 2g, 2g+1 of a quadrivalent). No data-dependent branch: the
// searches are nibble matches (a zero nibble of x ^ pattern), so a warp stays converged.
PSIM_HD uint32_t nib8_match(uint32_t v, uint32_t want) {   // bit 4t set where nibble t of v equals `want`
    const uint32_t x = v ^ (want * 0x11111111u);
    return ~(x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;
}
template <int K>
PSIM_HD uint32_t assign_chromatids(PhiloxStream& rng, uint32_t chrom_of, uint32_t centro) {
    constexpr int n = 2 * K;
    constexpr uint32_t used = n == 8 ? 0xFFFFFFFFu : 0x0000FFFFu;
    uint32_t at = 0x76543210u & used;
    uint32_t cof = chrom_of & used;                 // nibble t = centromere of the chromatid now at position t
    for (int s = 0; s < n - 1; s++) {
        const uint32_t want = (centro >> (4 * s)) & 15u;
        const uint32_t hit = nib8_match(cof, want) & used & (0xFFFFFFFFu << (4 * s));
        const int t = lowest_bit(hit) >> 2;         // a chromatid with that centromere is always left
        at = nib8_swap(at, s, t);
        cof = nib8_swap(cof, s, t);
    }
    if (K == 4) {
        const uint32_t coins = rng.next_bits(4);
        for (int c = 0; c < 4; c++) {
            const uint32_t hit = nib8_match(centro, (uint32_t)c);           // the two positions that receive centromere c
            const int first = lowest_bit(hit) >> 2, second = lowest_bit(hit & (hit - 1)) >> 2;
            if ((coins >> c) & 1u) at = nib8_swap(at, first, second);
        }
    }
    return at;
}

// ---------------------------------------------------------------- one multivalent, start to end
// Two steps, so that a kernel can let its warp reconverge between them: the chiasmata (loops whose
// trip counts differ from thread to thread), then centromere segregation and the assignment of the
// 2K chromatids to the gametes. draw_segregation returns nibble s = chromatid (numbered by the strand
// it starts on at the head) at gamete position s: gamete g is position g of a bivalent, positions
// 2g and 2g+1 of a quadrivalent.
template <bool QUAD, bool KOS, bool ANR>
PSIM_HD_NOINLINE double draw_chiasmata(ChiasmaList& x, PhiloxStream& rng, const ModelParams& m, const ChromParams& ch) {
    if (!QUAD) { chiasmata_parallel<2, KOS, ANR>(x, rng, ch); return 0.0; }
    // a quadrivalent pairs as a cross (four arms) or, with probability PARALLELMULTIVALENTS, as one
    // parallel bundle (Quadrivalent.doCrossingOver, Quadrivalent.java:371-378)
    const bool parallel = m.parallel_multivalents == 0.0 ? false : m.parallel_multivalents == 1.0 ? true : rng.next_unit() < m.parallel_multivalents;
    if (parallel) { chiasmata_parallel<4, KOS, ANR>(x, rng, ch); return 0.0; }   // no exchange limits: they stay 0 (Quadrivalent.java:42)
    return chiasmata_cross<KOS, ANR>(x, rng, ch);
}
template <bool QUAD>
PSIM_HD_NOINLINE uint32_t draw_segregation(const ChiasmaList& x, PhiloxStream& rng, const ModelParams& m, const ChromParams& ch, double lim_tail) {
    if (!QUAD) return assign_chromatids<2>(rng, centromere_chromosomes<2>(x, ch.centro), centromeres_bivalent(rng));
    return assign_chromatids<4>(rng, centromere_chromosomes<4>(x, ch.centro), centromeres_quadrivalent(rng, m, ch, lim_tail));
}
template <bool QUAD, bool KOS, bool ANR>
PSIM_HD_NOINLINE uint32_t draw_multivalent(ChiasmaList& x, PhiloxStream& rng, const ModelParams& m, const ChromParams& ch) {
    const double lim_tail = draw_chiasmata<QUAD, KOS, ANR>(x, rng, m, ch);
    if (x.overflow) return 0;
    return draw_segregation<QUAD>(x, rng, m, ch, lim_tail);
}

// ---------------------------------------------------------------- which gametes are kept, and their homolog order
// Gamete 0 of a meiosis makes the individual (Individual.java:258-259). With 2NGAMETES the unreduced
// gamete is gamete 0 plus gamete 1 (SDR) or plus one of the two gametes of the other pole (FDR, one
// coin per meiosis shared by its chromosomes; Individual.java:341-383). The P/2 homologs of every
// gamete are put in random order (Individual.java:325-335): dest[k].get(s) = position, inside kept
// gamete k, of the chromatid that segregation left at slot s.
struct Transmission {
    int n_gametes;
    int gamete[2];
    Nib16 dest[2];
};
PSIM_HD_NOINLINE Transmission draw_transmission(PhiloxStream& rng, const ModelParams& m, uint32_t child, uint32_t chrom, uint32_t parent) {
    Transmission t;
    t.n_gametes = m.unreduced ? 2 : 1;
    t.gamete[0] = 0;
    t.gamete[1] = 1;
    if (m.unreduced == 1) {
        rng.seek(child, 0u, parent, STAGE_UNREDUCED);
        t.gamete[1] = rng.next_bits(1) ? 3 : 2;
    }
    rng.seek(child, chrom, parent, STAGE_ORDER);
    const int p2 = m.ploidy / 2;
    for (int k = 0; k < t.n_gametes; k++) {
        Nib16 from = nib_identity();        // from.get(position) = slot
        shuffle_nib(rng, from, p2);
        t.dest[k].v = 0;
        for (int h = 0; h < p2; h++) t.dest[k].set(from.get(h), h);
    }
    return t;
}

}  // namespace psim
