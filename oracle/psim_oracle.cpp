// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product (see rng.hpp header).
//
// Flat C entry points (ctypes) over the CPU restatement in meiosis.hpp / popio.hpp. The calls
// deliberately mirror include/psim.h so that a parity test drives both sides identically.
// Homolog order in every array is the .hsa file order: [individual][chromosome][homolog].
#include <cstring>

#include "popio.hpp"
#include "device_twin.hpp"

using namespace orc;

namespace {

struct Handle {
    Population pop;
    int rng_mode = 0;
    int64_t seed = 0;
    uint32_t replicate = 0;
    // matrix-level emission inputs
    std::vector<int64_t> locus_off;      // C+1
    std::vector<double> locus_pos;       // L (Morgan, sorted per chromosome)
    int A = 0;
    std::vector<uint8_t> code;           // [L][A] allele code per founder allele (optional)
    std::string err;
};

int fail(Handle* h, const std::string& m) { h->err = m; return 1; }

}  // namespace

extern "C" {

void* orc_create(int ploidy, int chiasma_interference, int natural_pairing, int allow_no_recomb,
                 double parallel_multivalents, double paired_centromeres,
                 int rng_mode, int64_t seed, uint32_t replicate) {
    Handle* h = new Handle();
    Model& m = h->pop.model;
    m.ploidy = ploidy;
    m.chiasma_interference = chiasma_interference != 0;
    m.natural_pairing = natural_pairing != 0;
    m.allow_no_recomb = allow_no_recomb != 0;
    m.parallel_multivalents = parallel_multivalents;
    m.paired_centromeres = paired_centromeres;
    h->rng_mode = rng_mode;
    h->seed = seed;
    h->replicate = replicate;
    if (rng_mode == 0) h->pop.rng_holder.reset(new JavaRandom(seed));
    else {
        h->pop.rng_holder.reset(new PhiloxStream((uint32_t)seed, replicate));
        m.twin = true; m.twin_seed = (uint64_t)seed; m.twin_replicate = replicate;
    }
    m.rng = h->pop.rng_holder.get();
    return h;
}
// 2NGAMETES (Main.java:179-185): 0 = haploid gametes, 1 = FDR, 2 = SDR
int orc_set_unreduced_gametes(void* p, int kind) {
    Handle* h = (Handle*)p;
    if (kind < 0 || kind > 2) return fail(h, "unreduced gametes: 0, 1 (FDR) or 2 (SDR)");
    h->pop.model.unreduced_gametes = kind;
    return 0;
}
void orc_destroy(void* p) { delete (Handle*)p; }
const char* orc_last_error(void* p) { return ((Handle*)p)->err.c_str(); }

// length / centromere in Morgan (already cM/100.0, Main.java:619-620)
int orc_set_chromosomes(void* p, int C, const double* length, const double* centromere,
                        const double* pref_pairing, const double* frac_quad) {
    Handle* h = (Handle*)p;
    Model& m = h->pop.model;
    m.chrom.clear();
    for (int c = 0; c < C; c++) {
        Chrom ch;
        ch.name = "chr" + std::to_string(c + 1);
        ch.set(0.0, length[c], centromere[c]);
        if (m.ploidy > 2) {
            ch.pref_pairing = pref_pairing ? pref_pairing[c] : 0.0;
            ch.frac_quad = frac_quad ? frac_quad[c] : 0.0;
            ch.set_cumul_binomial(m.ploidy);
        }
        m.chrom.push_back(ch);
    }
    h->pop.locus.assign(C, {});
    return 0;
}

int orc_set_pedigree(void* p, int64_t N, const int32_t* parent0, const int32_t* parent1) {
    Handle* h = (Handle*)p;
    Population& pop = h->pop;
    pop.ind.assign(N, Individual());
    pop.ind_name.resize(N);
    pop.founder_allele_count = 0;
    for (int64_t i = 0; i < N; i++) {
        pop.ind_name[i] = "I" + std::to_string(i);
        if (parent0[i] < 0) { pop.founder_allele_count += pop.P(); continue; }
        if (parent0[i] >= i || parent1[i] >= i || parent1[i] < 0) return fail(h, "pedigree is not parents-first");
        pop.ind[i].parent[0] = parent0[i];
        pop.ind[i].parent[1] = parent1[i];
    }
    return 0;
}

int orc_set_haplostructs(void* p, int64_t first, int64_t count, const int64_t* seg_off,
                         const int32_t* founder, const double* pos) {
    Handle* h = (Handle*)p;
    Population& pop = h->pop;
    const int C = pop.C(), P = pop.P();
    if (first < 0 || first + count > (int64_t)pop.ind.size()) return fail(h, "individual range out of bounds");
    for (int64_t i = 0; i < count; i++) {
        Individual& x = pop.ind[first + i];
        x.hs.assign(C, std::vector<HS>(P));
        for (int c = 0; c < C; c++)
            for (int k = 0; k < P; k++) {
                int64_t hid = (i * C + c) * P + k;
                HS hs;
                for (int64_t s = seg_off[hid]; s < seg_off[hid + 1]; s++) { hs.founder.push_back(founder[s]); hs.pos.push_back(pos[s]); }
                if (hs.count() == 0) return fail(h, "empty haplostruct");
                x.hs[c][k] = hs;
            }
        x.has_hs = true;
    }
    return 0;
}

// Main.java:1504-1520; maxfounderallele as returned by readHaploStructFiles (-1 for a fresh run).
int orc_simulate(void* p, int maxfounderallele) {
    Handle* h = (Handle*)p;
    try {
        MeiosisEngine eng(h->pop.model);
        eng.stats = &h->pop.stats;
        eng.simulate_individuals(h->pop.ind, maxfounderallele);
    } catch (std::exception& ex) { return fail(h, ex.what()); }
    return 0;
}

int64_t orc_segment_count(void* p, int64_t first, int64_t count) {
    Handle* h = (Handle*)p;
    int64_t n = 0;
    for (int64_t i = first; i < first + count; i++)
        for (auto& cv : h->pop.ind[i].hs) for (auto& hs : cv) n += hs.count();
    return n;
}

int orc_get_haplostructs(void* p, int64_t first, int64_t count, int64_t* seg_off, int32_t* founder, double* pos) {
    Handle* h = (Handle*)p;
    Population& pop = h->pop;
    const int C = pop.C(), P = pop.P();
    int64_t n = 0, hid = 0;
    for (int64_t i = first; i < first + count; i++) {
        if (!pop.ind[i].has_hs) return fail(h, "individual has no haplostruct");
        for (int c = 0; c < C; c++)
            for (int k = 0; k < P; k++) {
                const HS& hs = pop.ind[i].hs[c][k];
                seg_off[hid++] = n;
                for (int s = 0; s < hs.count(); s++) { founder[n] = hs.founder[s]; pos[n] = hs.pos[s]; n++; }
            }
    }
    seg_off[hid] = n;
    return 0;
}

// quad: [count][C][2], chromseq: [count][C][2][P]; founders get -1 everywhere.
int orc_get_meiotic_configs(void* p, int64_t first, int64_t count, int32_t* quad, int32_t* chromseq) {
    Handle* h = (Handle*)p;
    Population& pop = h->pop;
    const int C = pop.C(), P = pop.P();
    for (int64_t i = 0; i < count; i++) {
        const Individual& x = pop.ind[first + i];
        for (int c = 0; c < C; c++)
            for (int par = 0; par < 2; par++) {
                int64_t q = (i * C + c) * 2 + par;
                if (!x.has_parconfig) { quad[q] = -1; for (int k = 0; k < P; k++) chromseq[q * P + k] = -1; continue; }
                quad[q] = x.parconfig[par][c].quad;
                for (int k = 0; k < P; k++) chromseq[q * P + k] = x.parconfig[par][c].chromseq[k];
            }
    }
    return 0;
}

// Positions in Morgan, ascending per chromosome (the order Chromosome.addLocus produces).
int orc_set_map(void* p, const int64_t* locus_off, const double* pos) {
    Handle* h = (Handle*)p;
    const int C = h->pop.C();
    h->locus_off.assign(locus_off, locus_off + C + 1);
    h->locus_pos.assign(pos, pos + locus_off[C]);
    return 0;
}

// code[l][a]: index of founder allele a's name among the locus' sorted unique names
// (Locus.java:88-103); the MAX allele is the largest code present at the locus.
int orc_set_allele_codes(void* p, int A, const uint8_t* code) {
    Handle* h = (Handle*)p;
    h->A = A;
    h->code.assign(code, code + (size_t)A * h->locus_pos.size());
    return 0;
}

static int chrom_of_locus(const Handle* h, int64_t l) {
    int c = 0;
    while (l >= h->locus_off[c + 1]) c++;
    return c;
}

// kind 0: founder alleles  u8 out[l][i*P+k]            (Main.java:903-939, founders == true)
// kind 1: allele codes     u8 out[l][i*P+k]            (Main.java:903-939, founders == false)
// kind 2: dose             u8 out[l][i], `which` as Main.java:969-1022:
//           with codes:   -2 MAX_ALLELE, -1 MIN_ALLELE, >=0 name of founder allele `which`
//           without codes: count of founder allele `which`
// kind 3: founder alleles  u16 out[l][i*P+k]
int orc_emit(void* p, int kind, int which, int64_t ind_first, int64_t ind_count, int64_t loc_first,
             int64_t loc_count, void* out, int64_t pitch_bytes) {
    Handle* h = (Handle*)p;
    Population& pop = h->pop;
    const int P = pop.P();
    const bool has_codes = !h->code.empty();
    try {
        for (int64_t l = loc_first; l < loc_first + loc_count; l++) {
            const int c = chrom_of_locus(h, l);
            const double position = h->locus_pos[l];
            uint8_t* row = (uint8_t*)out + (l - loc_first) * pitch_bytes;
            const uint8_t* codes = has_codes ? &h->code[(size_t)l * h->A] : nullptr;
            int match = which;
            if (kind == 2 && has_codes) {
                if (which == MAX_ALLELE) { match = 0; for (int a = 0; a < h->A; a++) match = std::max(match, (int)codes[a]); }
                else if (which == MIN_ALLELE) { match = 255; for (int a = 0; a < h->A; a++) match = std::min(match, (int)codes[a]); }
                else { if (which < 0 || which >= pop.founder_allele_count) which = 0; match = codes[which]; }
            }
            for (int64_t i = 0; i < ind_count; i++) {
                int dose = 0;
                for (int k = 0; k < P; k++) {
                    int f = pop.founder_at((int)(ind_first + i), c, k, position);
                    if (kind == 0) row[i * P + k] = (uint8_t)f;
                    else if (kind == 3) ((uint16_t*)row)[i * P + k] = (uint16_t)f;
                    else if (kind == 1) { if (!has_codes) return fail(h, "no allele codes set"); row[i * P + k] = codes[f]; }
                    else if (has_codes ? (codes[f] == match) : (f == match)) dose++;
                }
                if (kind == 2) row[i] = (uint8_t)dose;
            }
        }
    } catch (std::exception& ex) { return fail(h, ex.what()); }
    return 0;
}

// Reference statistics (TetraploidChromosome.java:54-58): out[0..7] = #meioses*chromosomes with
// k quadrivalents; out[8] bivalents, out[9] parallel quadrivalents, out[10] cross-type quadrivalents.
int orc_get_stats(void* p, int64_t* out) {
    Handle* h = (Handle*)p;
    for (int k = 0; k < 8; k++) out[k] = k < (int)h->pop.stats.quad_config_count.size() ? h->pop.stats.quad_config_count[k] : 0;
    out[8] = h->pop.stats.bivalents;
    out[9] = h->pop.stats.parallel_quads;
    out[10] = h->pop.stats.cross_quads;
    return 0;
}

// Whole-program restatement (Main.java:75-450, normal mode): reads a .par, writes the same files.
// The reference's stdout text is returned in `log` (truncated to log_cap).
int orc_run_parameter_file(const char* par_path, int rng_mode, char* log, int64_t log_cap) {
    std::string out;
    int rc = run_parameter_file(par_path, rng_mode, out);
    if (log && log_cap > 0) {
        size_t n = std::min((size_t)log_cap - 1, out.size());
        std::memcpy(log, out.data(), n);
        log[n] = 0;
    }
    return rc;
}

// ---- primitives exposed for known-answer tests
int32_t orc_java_next_int32(int64_t seed, int skip) { JavaRandom r(seed); int32_t v = 0; for (int i = 0; i <= skip; i++) v = r.next_int32(); return v; }
int32_t orc_java_next_int(int64_t seed, int bound, int skip) { JavaRandom r(seed); int v = 0; for (int i = 0; i <= skip; i++) v = r.next_int(bound); return v; }
double orc_java_next_double(int64_t seed, int skip) { JavaRandom r(seed); double v = 0; for (int i = 0; i <= skip; i++) v = r.next_double(); return v; }
void orc_philox_block(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { Philox4x32::block(ctr, key, out); }
// fills out[n] with the raw 32-bit words of stream (individual, chrom, parent, stage)
void orc_philox_stream(uint32_t seed, uint32_t replicate, uint32_t individual, uint32_t chrom, uint32_t parent,
                       uint32_t stage, int n, uint32_t* out) {
    PhiloxStream s(seed, replicate);
    s.seek(individual, chrom, parent, stage);
    for (int i = 0; i < n; i++) out[i] = s.next_u32();
}
// the product's own stream (pedigreesim_b200/csrc/psim_rng.cuh compiled for the host): held word for
// word against the oracle's restatement of Philox above (tests/test_oracle_pins.py)
void orc_twin_stream(uint64_t seed, uint32_t replicate, uint32_t individual, uint32_t chrom, uint32_t parent,
                     uint32_t stage, int n, uint32_t* out) {
    psim::PhiloxStream s;
    s.init(seed, replicate);
    s.seek(individual, chrom, parent, stage);
    for (int i = 0; i < n; i++) out[i] = s.next_u32();
}
double orc_calc_recomb_dist(double len) { return calc_recomb_dist(len); }
void orc_java_double_to_string(double d, char* buf, int cap) {
    std::string s = java_double_to_string(d);
    std::strncpy(buf, s.c_str(), cap - 1);
    buf[cap - 1] = 0;
}
double orc_ran_gamma(int64_t seed, double k, double theta, int skip) {
    JavaRandom r(seed); double v = 0; for (int i = 0; i <= skip; i++) v = ran_gamma(r, k, theta); return v;
}

}  // extern "C"
