// Input side of the libpsim host: PedigreeSim's .par / .chrom / .map / .ped / .gen / .hsa / .hsb /
// _meioticconfigs.dat readers and the pedigree ordering, producing the flat arrays the C ABI
// (include/psim.h) takes. Column-oriented, hash-indexed and whole-file tokenised so that
// pedigrees of 10^6..10^7 individuals load (the reference's getIndiv(name) and sortPedigree are
// O(N^2), PopulationData.java:244-252,309-430). Messages follow the reference because users and
// scripts read them (Main.java prints every failure to stdout and exits 0).
#pragma once

#include <algorithm>
#include <array>
#include <deque>
#include <map>
#include <unordered_map>

#include "text.hpp"

namespace psimhost {

// ---------------------------------------------------------------------------------------------
// .par  (Main.readParameterFile Main.java:499-579, getKeyandValue :1373-1403)
// ---------------------------------------------------------------------------------------------
using ParMap = std::map<std::string, std::string>;

inline bool is_boolean_word(const std::string& v) {
    std::string s = upper(v);
    return s == "0" || s == "1" || s == "FALSE" || s == "TRUE" || s == "NO" || s == "YES";
}
inline bool to_boolean(const std::string& v) {
    std::string s = upper(v);
    return s == "1" || s == "TRUE" || s == "YES";
}

// Splits "KEY = value", "KEY=value", "KEY =value", "KEY= value"; the key is upper-cased.
// Returns false when the line yields only one of the two.
inline bool key_and_value(sv line, std::string& key, std::string& value) {
    std::vector<sv> w;
    split_words(line, w);
    key.clear();
    value.clear();
    if (w.empty() || w[0][0] == ';') return true;
    size_t eq = w[0].find('=');
    if (eq != sv::npos) {
        // Java's "KEY=".split("=", 2) is {"KEY", ""}: a value in a separate word after "KEY=" is
        // never picked up (the words[1] branch of Main.java:1381-1388 is unreachable)
        key = std::string(w[0].substr(0, eq));
        value = std::string(w[0].substr(eq + 1));
    } else {
        key = std::string(w[0]);
        if (w.size() > 1) {
            if (w[1] == "=" && w.size() > 2)
                value = std::string(w[2]);
            else if (w[1][0] == '=')
                value = std::string(w[1].substr(1));
        }
    }
    key = upper(key);
    return key.empty() == value.empty();
}

inline bool valid_par_entry(const std::string& k, const std::string& v) {
    auto in = [&](std::initializer_list<const char*> l) {
        for (auto s : l)
            if (k == s) return true;
        return false;
    };
    int32_t iv;
    double dv;
    if (k == "PLOIDY") return parse_int32(v, iv) && iv > 0 && iv % 2 == 0;
    if (k == "MAPFUNCTION") return upper(v) == "HALDANE" || upper(v) == "KOSAMBI";
    if (in({"CHROMFILE", "PEDFILE", "MAPFILE", "FOUNDERFILE", "HAPLOSTRUCT", "OUTPUT", "MISSING"})) return true;
    if (k == "POPTYPE") return v == "F1" || v == "F2" || v == "BC" || v == "S1";  // case-sensitive, Main.java:536-540
    if (in({"POPSIZE", "TESTITER"})) return parse_int32(v, iv) && iv > 0;
    if (in({"ALLOWNORECOMBINATION", "ALLOWNOCHIASMATA", "NATURALPAIRING", "BIVALENTSBIDIRECTIONAL", "TEST",
            "PRINTEACHITER", "PRINTMAPDISTANCES", "PRINTPUBTABLES"}))
        return is_boolean_word(v);
    if (in({"PARALLELMULTIVALENTS", "PAIREDCENTROMERES"})) return parse_double(v, dv) && dv >= 0.0 && dv <= 1.0;
    if (in({"SEED", "PRINTGAMETES"})) return parse_int32(v, iv);
    if (k == "2NGAMETES") return upper(v) == "FDR" || upper(v) == "SDR";
    return false;
}

// Returns false after printing the reference's message for an invalid line.
inline bool read_parameter_file(const std::string& path, const std::string& shown_name, ParMap& par, std::string& log) {
    LineFile f;
    if (!f.open(path)) {
        log += "Error opening " + shown_name + " for reading\n";
        return false;
    }
    sv line;
    while (f.next(line)) {
        std::string k, v;
        bool ok = key_and_value(line, k, v);
        if (!ok) {
            log += "Invalid line in parameter file:\n" + std::string(line) + "\n";
            return false;
        }
        if (k.empty()) continue;
        if (!valid_par_entry(k, v)) {
            log += "Invalid line in parameter file:\n" + std::string(line) + "\n";
            return false;
        }
        par[k] = v;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Model settings (PopulationData.java:187-222) and the tables the C ABI consumes
// ---------------------------------------------------------------------------------------------
struct Settings {
    int ploidy = 2;
    bool kosambi = false, natural_pairing = true, allow_no_recomb = true;
    double parallel_multivalents = 0.0, paired_centromeres = 0.0;
    std::string missing = "NA";
    int32_t seed = 0;
};

struct Chromosomes {  // Chromosome.java:65-88, TetraploidChromosome.java:80-89
    std::vector<std::string> name;
    std::vector<double> length, centromere, pref_pairing, frac_quad, head, tail;
    int count() const { return (int)name.size(); }
};

struct Pedigree {  // parents-first; -1/-1 = founder
    std::vector<std::string> name;
    std::vector<int32_t> parent0, parent1;
    std::unordered_map<std::string, int32_t> index;  // first occurrence wins (PopulationData.getIndiv)
    int founder_allele_count = 0;
    int64_t count() const { return (int64_t)name.size(); }
    int32_t find(sv nm) const {
        auto it = index.find(std::string(nm));
        return it == index.end() ? -1 : it->second;
    }
};

struct LinkageMap {  // per chromosome, position-sorted as Chromosome.addLocus leaves it
    std::vector<std::vector<std::string>> name;
    std::vector<std::vector<double>> pos;
    std::vector<int64_t> locus_off;  // C+1
    std::vector<double> flat_pos;    // L
    std::vector<const std::string*> flat_name;
    // founder genotypes (.gen): allele names per locus, their codes, and the sorted unique names
    bool has_alleles = false;
    std::vector<std::vector<std::string>> unique_names;  // [L][k] sorted (Locus.java:88-103)
    std::vector<uint8_t> code;                           // [L][A]
    int max_allele_count = 0;
    int64_t total() const { return locus_off.empty() ? 0 : locus_off.back(); }
};

// .chrom  (Main.readChromosomeFile Main.java:581-634)
inline void read_chromosome_file(const std::string& path, const Settings& st, Chromosomes& ch) {
    LineFile f;
    if (!f.open(path)) throw HostError("cannot open file");
    sv line;
    std::vector<sv> w;
    bool have = f.next_data(line);
    if (have) split_words(line, w);
    if (!have || w.size() < 3 || upper(w[0]) != "CHROMOSOME" || upper(w[1]) != "LENGTH" || upper(w[2]) != "CENTROMERE")
        throw HostError("header line not found");
    size_t need = st.ploidy == 2 ? 3 : 5;
    while (f.next_data(line)) {
        split_words(line, w);
        if (w.size() < need) throw HostError("insufficient data on line '" + std::string(line) + "'");
        for (size_t i = 0; i < need; i++)
            if (w[i] == st.missing) throw HostError("missing data are not allowed");
        double len, cen, pp = 0, fq = 0;
        auto num = [&](sv s, double& d) {
            if (!parse_double(s, d)) throw HostError("For input string: \"" + std::string(s) + "\"");
        };
        num(w[1], len);
        num(w[2], cen);
        len /= 100.0;
        cen /= 100.0;
        if (st.ploidy != 2) {
            num(w[3], pp);
            num(w[4], fq);
        }
        ch.name.emplace_back(w[0]);
        ch.length.push_back(len);
        ch.centromere.push_back(cen);
        ch.pref_pairing.push_back(pp);
        ch.frac_quad.push_back(fq);
        ch.head.push_back(cen < 0.0 ? cen : 0.0);  // Chromosome.java:71-80: widened to hold the centromere
        ch.tail.push_back(cen > len ? cen : len);
    }
}

// .map  (Main.readMapFile Main.java:636-695; Chromosome.addLocus Chromosome.java:133-162)
inline void read_map_file(const std::string& path, const Settings& st, const Chromosomes& ch, LinkageMap& mp) {
    LineFile f;
    if (!f.open(path)) throw HostError("cannot open file");
    int C = ch.count();
    mp.name.assign(C, {});
    mp.pos.assign(C, {});
    sv line;
    std::vector<sv> w;
    bool have = f.next_data(line);
    if (have) split_words(line, w);
    if (!have || w.size() < 3 || upper(w[0]) != "MARKER" || upper(w[1]) != "CHROMOSOME" || upper(w[2]) != "POSITION")
        throw HostError("header line not found");
    std::string last;
    int cur = -1;
    while (f.next_data(line)) {
        split_words(line, w);
        if (w.size() < 3) throw HostError("too few items on line '" + std::string(line) + "'");
        for (int i = 0; i < 3; i++)
            if (w[i] == st.missing) throw HostError("missing data are not allowed");
        double pos;
        if (!parse_double(w[2], pos)) throw HostError("For input string: \"" + std::string(w[2]) + "\"");
        pos /= 100.0;
        if (cur < 0 || w[1] != last) {
            cur = -1;
            for (int c = 0; c < C; c++)
                if (ch.name[c] == w[1]) {
                    cur = c;
                    break;
                }
            if (cur < 0) throw HostError("chromName " + std::string(w[1]) + " not present");
            last = std::string(w[1]);
        }
        const double tol = 1.0e-7;
        if (pos < ch.head[cur]) {
            if (pos < ch.head[cur] - tol) throw HostError("addLocus: position invalid");
            pos = ch.head[cur];
        }
        if (pos > ch.tail[cur]) {
            if (pos > ch.tail[cur] + tol) throw HostError("addLocus: position invalid");
            pos = ch.tail[cur];
        }
        auto& P = mp.pos[cur];
        auto& N = mp.name[cur];
        if (P.empty() || P.back() <= pos) {  // the common, already sorted case
            P.push_back(pos);
            N.emplace_back(w[0]);
        } else {  // stable insert: after every locus at a position <= the new one
            size_t at = (size_t)(std::upper_bound(P.begin(), P.end(), pos) - P.begin());
            P.insert(P.begin() + at, pos);
            N.insert(N.begin() + at, std::string(w[0]));
        }
    }
    mp.locus_off.assign(1, 0);
    mp.flat_pos.clear();
    mp.flat_name.clear();
    for (int c = 0; c < C; c++) {
        mp.locus_off.push_back(mp.locus_off.back() + (int64_t)mp.pos[c].size());
        mp.flat_pos.insert(mp.flat_pos.end(), mp.pos[c].begin(), mp.pos[c].end());
        for (auto& s : mp.name[c]) mp.flat_name.push_back(&s);
    }
}

// .ped  (Main.readPedigreeFile Main.java:724-767)
using PedRows = std::vector<std::array<std::string, 3>>;
inline PedRows read_pedigree_file(const std::string& path, std::string& err) {
    PedRows rows;
    LineFile f;
    if (!f.open(path)) {
        err = "cannot open file";
        return rows;
    }
    sv line;
    std::vector<sv> w;
    bool have = f.next_data(line);
    if (have) split_words(line, w);
    if (!have || w.size() < 3 || upper(w[0]) != "NAME" || upper(w[1]) != "PARENT1" || upper(w[2]) != "PARENT2") {
        err = "header line not found";
        return rows;
    }
    while (f.next_data(line)) {
        split_words(line, w);
        if (w.size() < 3 || w[1][0] == ';' || w[2][0] == ';') {
            err = "incomplete line : " + std::string(line);
            return rows;
        }
        rows.push_back({std::string(w[0]), std::string(w[1]), std::string(w[2])});
    }
    return rows;
}

// Standard populations (Main.generateStandardPopulation Main.java:1466-1502); names are
// zero-padded to the width of POPSIZE.
inline PedRows generate_standard_population(const std::string& type, int popsize, const std::string& missing) {
    PedRows r;
    int width = (int)std::to_string(popsize).size();
    auto numbered = [&](const char* prefix, int i) {
        std::string n = std::to_string(i);
        return std::string(prefix) + std::string(width - (int)n.size(), '0') + n;
    };
    auto sibs = [&](const char* prefix, const char* a, const char* b) {
        r.reserve(r.size() + popsize);
        for (int i = 1; i <= popsize; i++) r.push_back({numbered(prefix, i), a, b});
    };
    if (type == "F1") {
        r.push_back({"P1", missing, missing});
        r.push_back({"P2", missing, missing});
        sibs("F1_", "P1", "P2");
    } else if (type == "F2") {
        r.push_back({"P1", missing, missing});
        r.push_back({"P2", missing, missing});
        r.push_back({"F1", "P1", "P2"});
        sibs("F2_", "F1", "F1");
    } else if (type == "BC") {
        r.push_back({"P1", missing, missing});
        r.push_back({"P2", missing, missing});
        r.push_back({"F1", "P1", "P2"});
        sibs("BC_", "P1", "F1");
    } else if (type == "S1") {
        r.push_back({"P1", missing, missing});
        sibs("S1_", "P1", "P1");
    }
    return r;
}

// PopulationData.sortPedigree (PopulationData.java:309-430): validation, then N passes that each
// place the first remaining individual whose parents both stand somewhere before it and rotate
// the individuals skipped on the way to the end of the list (so blocks of sibs keep their order).
// Same result, but name lookups are hashed and each scan step is O(1): a pass costs the number
// of individuals it skips instead of N times that.
inline PedRows sort_pedigree(const PedRows& ped, const std::string& missing) {
    const size_t n = ped.size();
    for (size_t j = 0; j < n; j++)
        if (ped[j][1] == ped[j][0] || ped[j][2] == ped[j][0])
            throw HostError("Individual '" + ped[j][0] + " is it's own parent!");
    for (size_t j = 0; j < n; j++)
        if (ped[j][0] == missing)
            throw HostError("Individual '" + std::to_string(j) + " has a missing name, which is not allowed");
    for (size_t j = 0; j < n; j++)
        if ((ped[j][1] == missing) != (ped[j][2] == missing))
            throw HostError("Individual " + ped[j][0] + " has one known and one unknown parent, which is not allowed");
    std::unordered_map<sv, int32_t> id;
    id.reserve(n * 2);
    for (size_t j = 0; j < n; j++)
        if (!id.emplace(sv(ped[j][0]), (int32_t)j).second)
            throw HostError("Individual " + ped[j][0] + " occurs more than once in pedigree");
    // parents as row ids: -1 = unknown (always "present"), -2 = a name that is nowhere in the file
    std::vector<int32_t> pa(n), pb(n);
    auto lookup = [&](const std::string& s) -> int32_t {
        if (s == missing) return -1;
        auto it = id.find(sv(s));
        return it == id.end() ? -2 : it->second;
    };
    for (size_t j = 0; j < n; j++) {
        pa[j] = lookup(ped[j][1]);
        pb[j] = lookup(ped[j][2]);
    }
    std::vector<uint8_t> placed(n, 0);
    std::vector<int64_t> skipped_in(n, -1);  // pass in which a row was last skipped
    std::deque<int32_t> rest;
    for (size_t j = 0; j < n; j++) rest.push_back((int32_t)j);
    PedRows out;
    out.reserve(n);
    auto before = [&](int32_t p, int64_t pass) { return p == -1 || (p >= 0 && (placed[p] || skipped_in[p] == pass)); };
    for (size_t pass = 0; pass < n; pass++) {
        size_t scanned = 0, len = rest.size();
        // rows skipped in this pass are popped to the back at once: the rotation of the reference
        while (scanned < len && !(before(pa[rest.front()], (int64_t)pass) && before(pb[rest.front()], (int64_t)pass))) {
            int32_t r = rest.front();
            rest.pop_front();
            skipped_in[r] = (int64_t)pass;
            rest.push_back(r);
            scanned++;
        }
        if (scanned == len) {
            // nobody can be placed: name the first remaining row with a parent that is not in the file
            for (int32_t r : rest) {
                if (pa[r] == -2)
                    throw HostError(ped[r][1] + " is listed as parent of " + ped[r][0] + " but it does not occur in the population");
                if (pb[r] == -2)
                    throw HostError(ped[r][2] + " is listed as parent of " + ped[r][0] + " but it does not occur in the population");
            }
            throw HostError("The pedigree could not be sorted, probably due to circular references");
        }
        int32_t r = rest.front();
        rest.pop_front();
        placed[r] = 1;
        out.push_back(ped[r]);
    }
    return out;
}

// PopulationData.createPopulation (PopulationData.java:432-464)
// Returns the reference's message ("" = ok) instead of throwing, like the original.
inline std::string create_population(const PedRows& rows, const Settings& st, Pedigree& pd) {
    size_t n = rows.size();
    pd.name.clear();
    pd.parent0.clear();
    pd.parent1.clear();
    pd.index.clear();
    pd.index.reserve(n * 2);
    pd.founder_allele_count = 0;
    for (size_t i = 0; i < n; i++) {
        int32_t a = -1, b = -1;
        if (rows[i][1] == st.missing) {
            pd.founder_allele_count += st.ploidy;
        } else {
            a = pd.find(rows[i][1]);
            b = pd.find(rows[i][2]);
            if (a < 0 || b < 0)  // Individual.setParents, Individual.java:191-199; the population stays partial
                return "Individual.setParents: parents may not be null";
        }
        pd.name.push_back(rows[i][0]);
        pd.parent0.push_back(a);
        pd.parent1.push_back(b);
        pd.index.emplace(rows[i][0], (int32_t)i);
    }
    return "";
}

// ---------------------------------------------------------------------------------------------
// HaploStructs as the C ABI takes them: CSR over homologs in .hsa order
// ---------------------------------------------------------------------------------------------
struct HaploBlock {
    std::vector<int64_t> seg_off{0};
    std::vector<int32_t> founder;
    std::vector<double> pos;
};

struct PresetHaplo {          // what a .hsa/.hsb pair holds
    std::vector<int32_t> ind; // individuals in file order
    HaploBlock hs;            // their homologs, file order
    int max_founder_allele = -1;
};

// .hsa/.hsb  (Main.readHaploStructFiles Main.java:1166-1291)
inline void read_haplostruct_files(const std::string& base, const Settings& st, const Chromosomes& ch,
                                   const Pedigree& pd, PresetHaplo& out) {
    LineFile hsa, hsb;
    if (!hsa.open(base + ".hsa")) throw HostError(base + ".hsa (No such file or directory)");
    if (!hsb.open(base + ".hsb")) throw HostError(base + ".hsb (No such file or directory)");
    const int C = ch.count(), P = st.ploidy;
    const int64_t per_ind = (int64_t)C * P;
    std::vector<uint8_t> seen((size_t)pd.count(), 0);
    std::vector<sv> w, wb;
    sv la, lb;
    int64_t line = 0;
    std::string iname;
    auto lineno = [&] { return std::to_string(line + 1); };
    while (hsa.next(la)) {
        la = jtrim(la);
        if (la.empty()) break;
        split_words(la, w);
        if (w.size() < 4) throw HostError(base + ".hsa: line " + lineno() + " has too few elements");
        int64_t indline = line % per_ind;
        if (indline == 0) {
            int32_t ix = pd.find(w[0]);
            if (ix < 0) throw HostError(base + ".hsa: individual " + std::string(w[0]) + " doesn't exist");
            if (seen[ix]) throw HostError(base + ".hsa: individual " + std::string(w[0]) + " occurs more than once");
            iname = std::string(w[0]);
            out.ind.push_back(ix);
        } else if (w[0] != iname) {
            throw HostError(base + ".hsa: insufficient lines for individual " + iname);
        }
        int c = (int)(indline / P), h = (int)(indline % P);
        if (w[1] != ch.name[c])
            throw HostError(base + ".hsa, line " + lineno() + ": chromosome " + ch.name[c] + " expected but " +
                            std::string(w[1]) + " found");
        int32_t homnr;
        if (!parse_int32(w[2], homnr) || homnr != h + 1)
            throw HostError(base + ".hsa, line " + lineno() + ": homolog " + std::to_string(h + 1) + " expected, but " +
                            std::string(w[2]) + " found");
        if (!hsb.next(lb)) throw HostError(base + ".hsb: less lines than " + base + ".hsa");
        auto founder_of = [&](sv s) {
            int32_t f = -1;
            if (!parse_int32(s, f) || f < 0 || f >= pd.founder_allele_count)
                throw HostError(base + ".hsa, line " + lineno() + ": invalid founder allele '" + std::string(s) + "'");
            if (f > out.max_founder_allele) out.max_founder_allele = f;
            return f;
        };
        int32_t f0 = founder_of(w[3]);
        split_words(lb, wb);
        if (wb.size() != w.size() - 4)
            throw HostError(base + ".hsa/hsb, line " + lineno() + ": number of items in both files don't match");
        out.hs.founder.push_back(f0);
        out.hs.pos.push_back(ch.head[c]);
        for (size_t r = 0; r < wb.size(); r++) {
            if (w[4 + r] == st.missing) {
                if (wb[r] != st.missing)
                    throw HostError(base + ".hsa/hsb, line " + lineno() +
                                    ": number of non-missing items in both files don't match");
                break;
            }
            int32_t fa = founder_of(w[4 + r]);
            double rp;
            if (!parse_double(wb[r], rp) || std::isnan(rp))
                throw HostError(base + ".hsa/hsb, line " + lineno() +
                                ": number of non-missing items in both files don't match");
            out.hs.founder.push_back(fa);
            out.hs.pos.push_back(rp / 100.0);  // cM to Morgan
        }
        out.hs.seg_off.push_back((int64_t)out.hs.founder.size());
        line++;
        if (line % per_ind == 0) seen[out.ind.back()] = 1;
    }
    if (line % per_ind != 0) {
        // the reference silently leaves an individual with an incomplete set without a HaploStruct
        out.ind.pop_back();
        int64_t keep = (int64_t)out.ind.size() * per_ind;
        out.hs.seg_off.resize((size_t)keep + 1);
        out.hs.founder.resize((size_t)out.hs.seg_off.back());
        out.hs.pos.resize((size_t)out.hs.seg_off.back());
    }
}

// _meioticconfigs.dat  (Main.readMeioticconfigsFile Main.java:1293-1353); values are kept as
// written for the individuals that are not re-simulated.
struct MeioticConfigs {
    std::vector<uint8_t> known;     // [N]
    std::vector<int32_t> quad;      // [N][C][2]
    std::vector<int32_t> chromseq;  // [N][C][2][P], 0-based
};

inline void read_meioticconfigs_file(const std::string& base, const Settings& st, const Chromosomes& ch,
                                     const Pedigree& pd, MeioticConfigs& mc, std::string& log) {
    std::string path = base + "_meioticconfigs.dat";
    LineFile f;
    if (!f.open(path)) {
        log += "Warning: " + path + " not found\n";
        return;
    }
    const int C = ch.count(), P = st.ploidy;
    std::vector<sv> w;
    sv l;
    int64_t line = 0;
    std::string iname;
    int32_t ix = 0;
    bool has = false;
    while (f.next(l)) {
        l = jtrim(l);
        if (l.empty()) break;
        split_words(l, w);
        if ((int)w.size() != 2 + 2 * (P + 1))
            throw HostError(path + ": line " + std::to_string(line + 1) + " doesn't have the correct number of fields");
        int c = (int)(line % C);
        if (c == 0) {
            ix = pd.find(w[0]);
            if (ix < 0) throw HostError(path + ": individual " + std::string(w[0]) + " doesn't exist");
            has = w[2] != st.missing;
            mc.known[ix] = has;
            iname = std::string(w[0]);
        } else if (w[0] != iname) {
            throw HostError(path + ": Insufficient lines for individual " + iname);
        }
        if (w[1] != ch.name[c])
            throw HostError(path + ", line " + std::to_string(line + 1) + ": chromosome " + ch.name[c] + " expected but " +
                            std::string(w[1]) + " found");
        if (has) {
            for (int par = 0; par < 2; par++) {
                int32_t v;
                sv q = w[2 + par * (P + 1)];
                if (!parse_int32(q, v)) throw HostError("For input string: \"" + std::string(q) + "\"");
                mc.quad[((size_t)ix * C + c) * 2 + par] = v;
                for (int h = 0; h < P; h++) {
                    sv s = w[3 + par * (P + 1) + h];
                    if (!parse_int32(s, v)) throw HostError("For input string: \"" + std::string(s) + "\"");
                    mc.chromseq[(((size_t)ix * C + c) * 2 + par) * P + h] = v - 1;
                }
            }
        }
        line++;
    }
}

// .gen  (Main.readFounderGenotypesFile Main.java:783-891). head_allele[f*C + c] = founder allele at
// the head of homolog 0 of the f-th pedigree individual that is a founder (in pedigree order
// this is what getHaploStruct(cix,0).getFounderAt(headPos) returns); founder_ind lists those
// individuals. Fills LinkageMap::unique_names / code / max_allele_count.
inline void read_founder_genotypes_file(const std::string& path, const Settings& st, const Chromosomes& ch,
                                        const Pedigree& pd, const std::vector<int32_t>& head_allele_of_ind,
                                        LinkageMap& mp) {
    LineFile f;
    if (!f.open(path)) throw HostError("cannot open file");
    const int P = st.ploidy, A = pd.founder_allele_count, C = ch.count();
    sv line;
    std::vector<sv> w;
    bool have = f.next_data(line);
    if (have) split_words(line, w);
    if (!have || w.empty() || upper(w[0]) != "MARKER") throw HostError("header line not found");
    if ((int)w.size() != 1 + A)
        throw HostError("must contain one column for the marker name and ploidy columns for each founder");
    int nf = A / P;
    std::vector<int32_t> indnr(nf);
    for (int i = 0; i < nf; i++) {
        sv cap = w[1 + i * P];
        size_t hs = cap.rfind('_');
        if (hs == sv::npos) throw HostError("String index out of range: -1");  // substring(0,-1) in the reference
        std::string nm(cap.substr(0, hs));
        int32_t ix = pd.find(nm);
        if (ix < 0 || pd.parent0[ix] >= 0) throw HostError(nm + " is not a founder");
        indnr[i] = ix;
        for (int hom = 0; hom < P; hom++) {
            sv c2 = w[1 + i * P + hom];
            bool ok = c2.size() >= 3 && hs < c2.size() && c2[hs] == '_' && c2.substr(0, hs) == sv(pd.name[ix]);
            int32_t hn = 0;
            ok = ok && parse_int32(c2.substr(hs + 1), hn) && hn == hom + 1;
            if (!ok)
                throw HostError(std::string(c2) + " found but " + pd.name[ix] + "_" + std::to_string(hom + 1) + " expected");
        }
    }
    for (int i = 1; i < nf; i++)
        for (int j = 0; j < i; j++)
            if (indnr[j] == indnr[i]) throw HostError("Founder " + pd.name[indnr[i]] + " occurs more than once in header");

    const int64_t L = mp.total();
    mp.unique_names.assign((size_t)L, {});
    mp.code.assign((size_t)L * A, 0);
    mp.max_allele_count = 0;
    int cix = 0;
    while (cix < C && mp.pos[cix].empty()) {
        // the reference compares against chromosome 0's (possibly empty) locus list first: an empty
        // chromosome makes every marker "invalid" (Main.java:863-867). Keep that.
        break;
    }
    int64_t loc = 0;      // within chromosome
    bool done = (C == 0);
    std::vector<std::string> names((size_t)A);
    std::vector<int> order((size_t)A);
    while (f.next_data(line)) {
        split_words(line, w);
        if ((int)w.size() != 1 + A) throw HostError("not all lines have " + std::to_string(1 + A) + " items");
        for (auto& x : w)
            if (x == st.missing) throw HostError("missing data are not allowed");
        if (done || loc >= (int64_t)mp.pos[cix].size() || mp.name[cix][(size_t)loc] != w[0])
            throw HostError("invalid marker: " + std::string(w[0]));
        for (auto& s : names) s.clear();
        for (int i = 0; i < nf; i++) {
            int32_t head = head_allele_of_ind[(size_t)indnr[i] * C + cix];
            for (int hom = 0; hom < P; hom++) {
                if (head + hom < 0 || head + hom >= A) throw HostError("Index " + std::to_string(head + hom) + " out of bounds for length " + std::to_string(A));
                names[(size_t)(head + hom)] = std::string(w[1 + i * P + hom]);
            }
        }
        int64_t g = mp.locus_off[cix] + loc;
        auto& uniq = mp.unique_names[(size_t)g];
        uniq = names;
        std::sort(uniq.begin(), uniq.end());  // Java String order == byte order for ASCII names
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        if (uniq.size() > 256) throw HostError("more than 256 different alleles at marker " + std::string(w[0]));
        for (int a = 0; a < A; a++)
            mp.code[(size_t)g * A + a] =
                (uint8_t)(std::lower_bound(uniq.begin(), uniq.end(), names[(size_t)a]) - uniq.begin());
        mp.max_allele_count = std::max(mp.max_allele_count, (int)uniq.size());
        loc++;
        if (loc == (int64_t)mp.pos[cix].size()) {
            loc = 0;
            cix++;
            if (cix == C) done = true;
        }
    }
    if (!done) throw HostError("end of file found while reading chromosome " + ch.name[cix]);
    mp.has_alleles = true;
}

}  // namespace psimhost
