// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product (see rng.hpp header).
//
// CPU restatement of PedigreeSim's population container, file formats, genotype emission and the
// `simulate` driver (reference: src/PedigreeSim/Main.java, PopulationData.java, Chromosome.java,
// Locus.java). Error texts follow the reference because they are part of the drop-in contract.
#pragma once
#include <algorithm>
#include <array>
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "meiosis.hpp"

namespace orc {

// Locus.java:15-47
struct Locus {
    std::string name;
    double position = 0.0;  // Morgan
    std::vector<std::string> allele_name;  // one per founder allele
    std::vector<std::string> unique_names;  // sorted; filled lazily (Locus.java:88-103)
    void count_allele_names() {
        unique_names.clear();
        for (size_t i = 0; i < allele_name.size(); i++) {
            bool seen = false;
            for (size_t j = 0; j < i && !seen; j++) seen = allele_name[i] == allele_name[j];
            if (!seen) unique_names.push_back(allele_name[i]);
        }
        std::sort(unique_names.begin(), unique_names.end());  // Java String order == byte order for ASCII
    }
    const std::vector<std::string>& uniques() { if (unique_names.empty()) count_allele_names(); return unique_names; }
};

struct Population {
    Model model;
    std::string missing = "NA";
    std::vector<std::string> ind_name;
    std::vector<Individual> ind;
    std::vector<std::vector<Locus>> locus;  // [chrom] sorted by position (Chromosome.java:133-162)
    int founder_allele_count = 0;
    std::unique_ptr<Rng> rng_holder;
    MeiosisStats stats;

    int C() const { return (int)model.chrom.size(); }
    int P() const { return model.ploidy; }
    int find_ind(const std::string& n) const {  // PopulationData.java:244-252 (linear scan)
        for (size_t i = 0; i < ind_name.size(); i++) if (ind_name[i] == n) return (int)i;
        return -1;
    }
    // Chromosome.java:133-162 — clamp within 1e-7, stable insert after equal positions.
    void add_locus(int c, Locus l) {
        const Chrom& ch = model.chrom[c];
        const double tol = 1.0e-7;
        if (l.position < ch.head) {
            if (l.position < ch.head - tol) throw std::runtime_error("addLocus: position invalid");
            l.position = ch.head;
        }
        if (l.position > ch.tail) {
            if (l.position > ch.tail + tol) throw std::runtime_error("addLocus: position invalid");
            l.position = ch.tail;
        }
        if ((int)locus.size() < C()) locus.resize(C());
        auto& v = locus[c];
        size_t index = 0;
        while (index < v.size() && v[index].position <= l.position) index++;
        v.insert(v.begin() + index, std::move(l));
    }
    int max_allele_count() {  // PopulationData.getMaxAlleleCount
        int mx = 0;
        for (auto& cl : locus) for (auto& l : cl) mx = std::max(mx, (int)l.uniques().size());
        return mx;
    }
    // HaploStruct.java:135-142
    int founder_at(int i, int c, int h, double position) const {
        const Chrom& ch = model.chrom[c];
        if (position < ch.head || position > ch.tail) throw std::runtime_error("getFounderAt: invalid position");
        const HS& hs = ind[i].hs[c][h];
        return hs.founder[hs.find_segment(position)];
    }
};

// ---------------------------------------------------------------- text helpers
inline std::string trim(const std::string& s) {  // Java String.trim(): strips chars <= ' '
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') a++;
    while (b > a && (unsigned char)s[b - 1] <= ' ') b--;
    return s.substr(a, b - a);
}
inline std::vector<std::string> read_words(const std::string& line) {  // Tools.java:392-401
    std::vector<std::string> w;
    std::string t = trim(line);
    size_t i = 0;
    while (i < t.size()) {
        while (i < t.size() && std::isspace((unsigned char)t[i])) i++;
        size_t j = i;
        while (j < t.size() && !std::isspace((unsigned char)t[j])) j++;
        if (j > i) w.push_back(t.substr(i, j - i));
        i = j;
    }
    return w;
}
inline std::string upper(std::string s) { for (auto& c : s) c = (char)std::toupper((unsigned char)c); return s; }
inline bool getline_any(std::istream& in, std::string& s) {
    if (!std::getline(in, s)) return false;
    if (!s.empty() && s.back() == '\r') s.pop_back();
    return true;
}
inline bool parse_int(const std::string& s, long& out) {  // Integer.parseInt: optional sign + digits only
    if (s.empty()) return false;
    size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (i == s.size()) return false;
    for (size_t k = i; k < s.size(); k++) if (!std::isdigit((unsigned char)s[k])) return false;
    errno = 0;
    long long v = std::strtoll(s.c_str(), nullptr, 10);
    if (errno || v < INT32_MIN || v > INT32_MAX) return false;
    out = (long)v;
    return true;
}
inline bool parse_double(const std::string& s, double& out) {  // Double.parseDouble (subset: no hex, no suffix)
    std::string t = trim(s);
    if (t.empty()) return false;
    char* end = nullptr;
    out = std::strtod(t.c_str(), &end);
    return end && *end == 0;
}
// Java Double.toString: shortest digits that round-trip, decimal notation for 1e-3 <= |d| < 1e7,
// otherwise "d.dddE[-]n" (Main.java:1122 prints breakpoints with string concatenation).
inline std::string java_double_to_string(double d) {
    if (std::isnan(d)) return "NaN";
    if (std::isinf(d)) return d > 0 ? "Infinity" : "-Infinity";
    if (d == 0.0) return std::signbit(d) ? "-0.0" : "0.0";
    char buf[64];
    int prec = 1;
    for (; prec <= 17; prec++) {
        std::snprintf(buf, sizeof buf, "%.*e", prec - 1, d);
        if (std::strtod(buf, nullptr) == d) break;
    }
    // buf = [-]D.DDDDe[+-]XX
    std::string s(buf);
    bool neg = s[0] == '-';
    if (neg) s = s.substr(1);
    size_t epos = s.find('e');
    std::string mant = s.substr(0, epos);
    int exp10 = std::atoi(s.c_str() + epos + 1);
    std::string digits;
    for (char ch : mant) if (ch != '.') digits.push_back(ch);
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    std::string out;
    double ad = std::fabs(d);
    if (ad >= 1e-3 && ad < 1e7) {
        if (exp10 >= 0) {
            std::string ip, fp;
            if ((int)digits.size() <= exp10 + 1) { ip = digits + std::string(exp10 + 1 - digits.size(), '0'); fp = "0"; }
            else { ip = digits.substr(0, exp10 + 1); fp = digits.substr(exp10 + 1); }
            out = ip + "." + fp;
        } else {
            out = "0." + std::string(-exp10 - 1, '0') + digits;
        }
    } else {
        std::string fp = digits.size() > 1 ? digits.substr(1) : "0";
        out = digits.substr(0, 1) + "." + fp + "E" + std::to_string(exp10);
    }
    return neg ? "-" + out : out;
}

// ---------------------------------------------------------------- readers
const char COMMENT = ';';

// Main.java:1373-1403
inline std::pair<std::string, std::string> key_and_value(const std::string& line, bool& ok) {
    ok = true;
    auto words = read_words(line);
    if (words.empty() || words[0][0] == COMMENT) return {"", ""};
    std::string key, value;
    size_t eq = words[0].find('=');
    if (eq != std::string::npos) {
        key = words[0].substr(0, eq);
        std::string rest = words[0].substr(eq + 1);
        // Java split("=",2): "K=" yields {"K",""} (length 2, empty value) — value stays "" then.
        value = rest;
        if (eq + 1 == words[0].size()) {
            // subwords.length is 2 with an empty second element for "KEY=" -> value ""
            value = "";
        }
    } else {
        key = words[0];
        if (words.size() > 1) {
            if (words[1] == "=" && words.size() > 2) value = words[2];
            else if (words[1][0] == '=') value = words[1].substr(1);
        }
    }
    return {upper(key), value};
}
inline bool is_boolean(const std::string& v) {
    std::string s = upper(v);
    return s == "0" || s == "1" || s == "FALSE" || s == "TRUE" || s == "NO" || s == "YES";
}
inline bool to_boolean(const std::string& v) { std::string s = upper(v); return s == "1" || s == "TRUE" || s == "YES"; }

// Main.java:499-579. Returns false (after printing the reference's message) on an invalid line.
inline bool read_parameter_file(const std::string& path, std::map<std::string, std::string>& par, std::string& out_msg) {
    std::ifstream in(path);
    if (!in) { out_msg = "Error opening " + path + " for reading"; return false; }
    std::string s;
    while (getline_any(in, s)) {
        bool ok;
        auto kv = key_and_value(s, ok);
        const std::string &k = kv.first, &v = kv.second;
        if (k.empty() != v.empty()) { out_msg = "Invalid line in parameter file:\n" + s; return false; }
        if (k.empty()) continue;
        long iv; double dv;
        bool valid =
            (k == "PLOIDY" && parse_int(v, iv) && iv > 0 && iv % 2 == 0) ||
            (k == "MAPFUNCTION" && (upper(v) == "HALDANE" || upper(v) == "KOSAMBI")) ||
            (k == "CHROMFILE" || k == "PEDFILE" || k == "MAPFILE" || k == "FOUNDERFILE" || k == "HAPLOSTRUCT" || k == "OUTPUT") ||
            (k == "POPTYPE" && (v == "F1" || v == "F2" || v == "BC" || v == "S1")) ||
            ((k == "POPSIZE" || k == "TESTITER") && parse_int(v, iv) && iv > 0) ||
            ((k == "ALLOWNORECOMBINATION" || k == "ALLOWNOCHIASMATA" || k == "NATURALPAIRING" ||
              k == "BIVALENTSBIDIRECTIONAL" || k == "TEST" || k == "PRINTEACHITER" ||
              k == "PRINTMAPDISTANCES" || k == "PRINTPUBTABLES") && is_boolean(v)) ||
            ((k == "PARALLELMULTIVALENTS" || k == "PAIREDCENTROMERES") && parse_double(v, dv) && dv >= 0.0 && dv <= 1.0) ||
            (k == "MISSING") ||
            ((k == "SEED" || k == "PRINTGAMETES") && parse_int(v, iv)) ||
            (k == "2NGAMETES" && (upper(v) == "FDR" || upper(v) == "SDR"));
        if (!valid) { out_msg = "Invalid line in parameter file:\n" + s; return false; }
        par[k] = v;
    }
    return true;
}

inline bool next_data_line(std::istream& in, std::string& s) {
    while (getline_any(in, s)) {
        s = trim(s);
        if (!s.empty() && s[0] != COMMENT) return true;
    }
    return false;
}

// Main.java:581-634
inline void read_chromosome_file(const std::string& path, Population& pop) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open file");
    std::string s;
    std::vector<std::string> w;
    if (next_data_line(in, s)) w = read_words(s);
    if (w.size() < 3 || upper(w[0]) != "CHROMOSOME" || upper(w[1]) != "LENGTH" || upper(w[2]) != "CENTROMERE")
        throw std::runtime_error("header line not found");
    const size_t needed = pop.P() == 2 ? 3 : 5;
    while (next_data_line(in, s)) {
        w = read_words(s);
        if (w.size() < needed) throw std::runtime_error("insufficient data on line '" + s + "'");
        for (size_t i = 0; i < needed; i++)
            if (w[i] == pop.missing) throw std::runtime_error("missing data are not allowed");
        Chrom ch;
        ch.name = w[0];
        double len, cen;
        if (!parse_double(w[1], len) || !parse_double(w[2], cen)) throw std::runtime_error("For input string: \"" + w[1] + "\"");
        ch.set(0.0, len / 100.0, cen / 100.0);
        if (pop.P() > 2) {
            if (!parse_double(w[3], ch.pref_pairing) || !parse_double(w[4], ch.frac_quad))
                throw std::runtime_error("For input string: \"" + w[3] + "\"");
            ch.set_cumul_binomial(pop.P());
        }
        pop.model.chrom.push_back(ch);
    }
    pop.locus.resize(pop.C());
}

// Main.java:636-695
inline void read_map_file(const std::string& path, Population& pop) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open file");
    std::string s;
    std::vector<std::string> w;
    if (next_data_line(in, s)) w = read_words(s);
    if (w.size() < 3 || upper(w[0]) != "MARKER" || upper(w[1]) != "CHROMOSOME" || upper(w[2]) != "POSITION")
        throw std::runtime_error("header line not found");
    std::string last_name;
    int chrom = -1;
    while (next_data_line(in, s)) {
        w = read_words(s);
        if (w.size() < 3) throw std::runtime_error("too few items on line '" + s + "'");
        for (int i = 0; i < 3; i++) if (w[i] == pop.missing) throw std::runtime_error("missing data are not allowed");
        double pos;
        if (!parse_double(w[2], pos)) throw std::runtime_error("For input string: \"" + w[2] + "\"");
        pos /= 100.0;
        if (w[1] != last_name) {
            int chr = 0;
            do { chrom = chr; chr++; } while (chr < pop.C() && w[1] != pop.model.chrom[chrom].name);
            if (w[1] != pop.model.chrom[chrom].name) throw std::runtime_error("chromName " + w[1] + " not present");
            last_name = w[1];
        }
        Locus l;
        l.name = w[0];
        l.position = pos;
        // Locus.java:34-46 with alleleNames {"0","1"}: names "0","1", then ""+i
        l.allele_name.resize(pop.founder_allele_count);
        for (int i = 0; i < pop.founder_allele_count; i++) l.allele_name[i] = std::to_string(i);
        pop.add_locus(chrom, std::move(l));
    }
}

// Main.java:724-767
inline std::vector<std::array<std::string, 3>> read_pedigree_file(const std::string& path, std::string& err) {
    std::vector<std::array<std::string, 3>> ped;
    err.clear();
    std::ifstream in(path);
    if (!in) { err = "cannot open file"; return ped; }
    std::string s;
    std::vector<std::string> w;
    if (next_data_line(in, s)) w = read_words(s);
    if (w.size() < 3 || upper(w[0]) != "NAME" || upper(w[1]) != "PARENT1" || upper(w[2]) != "PARENT2") {
        err = "header line not found";
        return ped;
    }
    while (next_data_line(in, s)) {
        w = read_words(s);
        if (w.size() < 3 || w[1][0] == COMMENT || w[2][0] == COMMENT) { err = "incomplete line : " + s; return ped; }
        ped.push_back({w[0], w[1], w[2]});
    }
    return ped;
}

// PopulationData.java:309-430 — validation + the reference's stable "move blockers to the end" sort.
inline std::vector<std::array<std::string, 3>> sort_pedigree(const std::vector<std::array<std::string, 3>>& ped,
                                                             const std::string& missing) {
    const int n = (int)ped.size();
    for (int j = 0; j < n; j++)
        if (ped[j][1] == ped[j][0] || ped[j][2] == ped[j][0])
            throw std::runtime_error("Individual '" + ped[j][0] + " is it's own parent!");
    for (int j = 0; j < n; j++)
        if (ped[j][0] == missing)
            throw std::runtime_error("Individual '" + std::to_string(j) + " has a missing name, which is not allowed");
    for (int j = 0; j < n; j++)
        if ((ped[j][1] == missing) != (ped[j][2] == missing))
            throw std::runtime_error("Individual " + ped[j][0] + " has one known and one unknown parent, which is not allowed");
    for (int j = 1; j < n; j++)
        for (int k = j - 1; k >= 0; k--)
            if (ped[j][0] == ped[k][0])
                throw std::runtime_error("Individual " + ped[j][0] + " occurs more than once in pedigree");
    std::vector<std::array<std::string, 3>> ind(2 * n);
    for (int i = 0; i < n; i++) ind[i] = ped[i];
    auto present = [&](const std::string& name, int i) {
        if (name == missing) return true;
        for (int j = i - 1; j >= 0; j--) if (name == ind[j][0]) return true;
        return false;
    };
    for (int pass = 0; pass < n; pass++) {
        int i = pass;
        while (i < n && (!present(ind[i][1], i) || !present(ind[i][2], i))) i++;
        if (i >= n) {
            i = pass;
            while (i < n && present(ind[i][1], n) && present(ind[i][2], n)) i++;
            if (i < n) {
                const std::string& bad = !present(ind[i][1], n) ? ind[i][1] : ind[i][2];
                throw std::runtime_error(bad + " is listed as parent of " + ind[i][0] + " but it does not occur in the population");
            }
            throw std::runtime_error("The pedigree could not be sorted, probably due to circular references");
        }
        if (i > pass) {
            for (int k = pass; k < i; k++) ind[n + k - pass] = ind[k];
            for (int k = 0; k < n - pass; k++) ind[pass + k] = ind[i + k];
        }
    }
    ind.resize(n);
    return ind;
}

// PopulationData.java:432-464
inline void create_population(Population& pop, const std::vector<std::array<std::string, 3>>& ped) {
    pop.ind.clear();
    pop.ind_name.clear();
    pop.founder_allele_count = 0;
    for (auto& row : ped) {
        Individual x;
        if (row[1] == pop.missing) {
            pop.founder_allele_count += pop.P();
        } else {
            x.parent[0] = pop.find_ind(row[1]);
            x.parent[1] = pop.find_ind(row[2]);
            if (x.parent[0] < 0 || x.parent[1] < 0)
                throw std::runtime_error("Individual.setParents: parents may not be null");
        }
        pop.ind_name.push_back(row[0]);
        pop.ind.push_back(std::move(x));
    }
}

// Main.java:1466-1502
inline std::vector<std::array<std::string, 3>> generate_standard_population(const std::string& type, int popsize,
                                                                            const std::string& missing) {
    std::vector<std::array<std::string, 3>> ped;
    const int width = (int)std::to_string(popsize).size();
    auto num = [&](int i) { std::string s = std::to_string(i); return std::string(width - std::min(width, (int)s.size()), '0') + s; };
    if (type == "F1") {
        ped.push_back({"P1", missing, missing});
        ped.push_back({"P2", missing, missing});
        for (int i = 1; i <= popsize; i++) ped.push_back({"F1_" + num(i), "P1", "P2"});
    } else if (type == "F2") {
        ped.push_back({"P1", missing, missing});
        ped.push_back({"P2", missing, missing});
        ped.push_back({"F1", "P1", "P2"});
        for (int i = 1; i <= popsize; i++) ped.push_back({"F2_" + num(i), "F1", "F1"});
    } else if (type == "BC") {
        ped.push_back({"P1", missing, missing});
        ped.push_back({"P2", missing, missing});
        ped.push_back({"F1", "P1", "P2"});
        for (int i = 1; i <= popsize; i++) ped.push_back({"BC_" + num(i), "P1", "F1"});
    } else if (type == "S1") {
        ped.push_back({"P1", missing, missing});
        for (int i = 1; i <= popsize; i++) ped.push_back({"S1_" + num(i), "P1", "P1"});
    }
    return ped;
}

// Main.java:783-891
inline void read_founder_genotypes_file(const std::string& path, Population& pop) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open file");
    std::string s;
    std::vector<std::string> w;
    if (next_data_line(in, s)) w = read_words(s);
    if (w.empty() || upper(w[0]) != "MARKER") throw std::runtime_error("header line not found");
    const int A = pop.founder_allele_count, P = pop.P();
    if ((int)w.size() != 1 + A)
        throw std::runtime_error("must contain one column for the marker name and ploidy columns for each founder");
    const int nf = A / P;
    std::vector<int> indnr(nf);
    for (int i = 0; i < nf; i++) {
        std::string iname = w[1 + i * P];
        size_t hspos = iname.rfind('_');
        std::string base = hspos == std::string::npos ? std::string() : iname.substr(0, hspos);
        int ind = hspos == std::string::npos ? -1 : pop.find_ind(base);
        if (ind < 0 || !pop.ind[ind].is_founder()) throw std::runtime_error(base + " is not a founder");
        indnr[i] = ind;
        for (int hom = 0; hom < P; hom++) {
            const std::string& cap = w[1 + i * P + hom];
            bool ok = cap.size() >= 3 && cap.size() > hspos && cap[hspos] == '_' && cap.substr(0, hspos) == pop.ind_name[ind];
            long p = 0;
            ok = ok && parse_int(cap.substr(hspos + 1), p) && p == hom + 1;
            if (!ok) throw std::runtime_error(cap + " found but " + pop.ind_name[ind] + "_" + std::to_string(hom + 1) + " expected");
        }
    }
    for (int i = 1; i < nf; i++)
        for (int j = 0; j < i; j++)
            if (indnr[j] == indnr[i]) throw std::runtime_error("Founder " + pop.ind_name[indnr[i]] + " occurs more than once in header");
    int cix = 0, loc = 0;
    bool chrom_ok = pop.C() > 0;
    while (next_data_line(in, s)) {
        w = read_words(s);
        if ((int)w.size() != 1 + A) throw std::runtime_error("not all lines have " + std::to_string(1 + A) + " items");
        for (auto& x : w) if (x == pop.missing) throw std::runtime_error("missing data are not allowed");
        if (!chrom_ok || loc >= (int)pop.locus[cix].size() || pop.locus[cix][loc].name != w[0])
            throw std::runtime_error("invalid marker: " + w[0]);
        std::vector<std::string> names(A);
        for (int i = 0; i < nf; i++) {
            int head_allele = pop.founder_at(indnr[i], cix, 0, pop.model.chrom[cix].head);
            for (int hom = 0; hom < P; hom++) names[head_allele + hom] = w[1 + i * P + hom];
        }
        pop.locus[cix][loc].allele_name = names;
        pop.locus[cix][loc].unique_names.clear();
        loc++;
        if (loc == (int)pop.locus[cix].size()) {
            loc = 0;
            cix++;
            // the reference does not skip chromosomes without loci; neither do we
            if (cix == pop.C()) chrom_ok = false;
        }
    }
    if (chrom_ok) throw std::runtime_error("end of file found while reading chromosome " + pop.model.chrom[cix].name);
}

// Main.java:1166-1291. Returns maxfounderallele.
inline int read_haplostruct_files(const std::string& base, Population& pop) {
    std::ifstream hsa(base + ".hsa"), hsb(base + ".hsb");
    if (!hsa) throw std::runtime_error(base + ".hsa (No such file or directory)");
    if (!hsb) throw std::runtime_error(base + ".hsb (No such file or directory)");
    for (auto& x : pop.ind) { x.has_hs = false; x.hs.clear(); }
    const int C = pop.C(), P = pop.P();
    std::string s, sb;
    long line = 0;
    int maxfounder = -1;
    int cur = -1;
    std::vector<std::vector<HS>> hs;
    while (getline_any(hsa, s)) {
        s = trim(s);
        if (s.empty()) break;
        auto w = read_words(s);
        const std::string ln = std::to_string(line + 1);
        if (w.size() < 4) throw std::runtime_error(base + ".hsa: line " + ln + " has too few elements");
        if (line % (C * P) == 0) {
            cur = pop.find_ind(w[0]);
            if (cur < 0) throw std::runtime_error(base + ".hsa: individual " + w[0] + " doesn't exist");
            if (pop.ind[cur].has_hs) throw std::runtime_error(base + ".hsa: individual " + w[0] + " occurs more than once");
            hs.assign(C, std::vector<HS>(P));
        } else if (w[0] != pop.ind_name[cur]) {
            throw std::runtime_error(base + ".hsa: insufficient lines for individual " + pop.ind_name[cur]);
        }
        int indline = (int)(line % (C * P));
        int chrnum = indline / P, homol = indline % P;
        const Chrom& ch = pop.model.chrom[chrnum];
        if (w[1] != ch.name)
            throw std::runtime_error(base + ".hsa, line " + ln + ": chromosome " + ch.name + " expected but " + w[1] + " found");
        long homnr;
        if (!parse_int(w[2], homnr) || homnr != homol + 1)
            throw std::runtime_error(base + ".hsa, line " + ln + ": homolog " + std::to_string(homol + 1) + " expected, but " + w[2] + " found");
        if (!getline_any(hsb, sb)) throw std::runtime_error(base + ".hsb: less lines than " + base + ".hsa");
        long fa = -1;
        if (!parse_int(w[3], fa)) fa = -1;
        if (fa < 0 || fa >= pop.founder_allele_count)
            throw std::runtime_error(base + ".hsa, line " + ln + ": invalid founder allele '" + w[3] + "'");
        if (fa > maxfounder) maxfounder = (int)fa;
        auto wb = read_words(sb);
        if (wb.size() != w.size() - 4)
            throw std::runtime_error(base + ".hsa/hsb, line " + ln + ": number of items in both files don't match");
        HS h(ch.head, (int)fa);
        for (size_t r = 0; r < wb.size(); r++) {
            if (w[4 + r] == pop.missing) {
                if (wb[r] != pop.missing)
                    throw std::runtime_error(base + ".hsa/hsb, line " + ln + ": number of non-missing items in both files don't match");
                break;
            }
            fa = -1;
            if (!parse_int(w[4 + r], fa)) fa = -1;
            if (fa < 0 || fa >= pop.founder_allele_count)
                throw std::runtime_error(base + ".hsa, line " + ln + ": invalid founder allele '" + w[4 + r] + "'");
            if (fa > maxfounder) maxfounder = (int)fa;
            double rp;
            if (!parse_double(wb[r], rp) || std::isnan(rp))
                throw std::runtime_error(base + ".hsa/hsb, line " + ln + ": number of non-missing items in both files don't match");
            h.add_segment(rp / 100.0, (int)fa);
        }
        hs[chrnum][homol] = h;
        line++;
        if (line % (C * P) == 0) { pop.ind[cur].hs = hs; pop.ind[cur].has_hs = true; }
    }
    return maxfounder;
}

// Main.java:1293-1353
inline void read_meioticconfigs_file(const std::string& base, Population& pop, std::string& warning) {
    const std::string fname = base + "_meioticconfigs.dat";
    std::ifstream in(fname);
    if (!in) { warning = "Warning: " + fname + " not found"; return; }
    const int C = pop.C(), P = pop.P();
    std::string s;
    long line = 0;
    int cur = 0;
    while (getline_any(in, s)) {
        s = trim(s);
        if (s.empty()) break;
        auto w = read_words(s);
        const std::string ln = std::to_string(line + 1);
        if ((int)w.size() != 2 + 2 * (P + 1)) throw std::runtime_error(fname + ": line " + ln + " doesn't have the correct number of fields");
        if (line % C == 0) {
            cur = pop.find_ind(w[0]);
            if (cur < 0) throw std::runtime_error(fname + ": individual " + w[0] + " doesn't exist");
            Individual& x = pop.ind[cur];
            if (w[2] == pop.missing) { x.has_parconfig = false; }
            else { x.has_parconfig = true; x.parconfig[0].assign(C, ChromConfig()); x.parconfig[1].assign(C, ChromConfig()); }
        } else if (w[0] != pop.ind_name[cur]) {
            throw std::runtime_error(fname + ": Insufficient lines for individual " + pop.ind_name[cur]);
        }
        int indline = (int)(line % C);
        if (w[1] != pop.model.chrom[indline].name)
            throw std::runtime_error(fname + ", line " + ln + ": chromosome " + pop.model.chrom[indline].name + " expected but " + w[1] + " found");
        Individual& x = pop.ind[cur];
        if (x.has_parconfig) {
            for (int par = 0; par < 2; par++) {
                ChromConfig cc;
                long v;
                if (!parse_int(w[2 + par * (P + 1)], v)) throw std::runtime_error("For input string: \"" + w[2 + par * (P + 1)] + "\"");
                cc.quad = (int)v;
                cc.chromseq.resize(P);
                for (int h = 0; h < P; h++) {
                    if (!parse_int(w[3 + par * (P + 1) + h], v)) throw std::runtime_error("For input string: \"" + w[3 + par * (P + 1) + h] + "\"");
                    cc.chromseq[h] = (int)v - 1;
                }
                x.parconfig[par][indline] = cc;
            }
        }
        line++;
    }
}

// ---------------------------------------------------------------- writers
struct OutFile {
    FILE* f;
    explicit OutFile(const std::string& path) : f(std::fopen(path.c_str(), "wb")) {
        if (!f) throw std::runtime_error(path + " (cannot create)");
        std::setvbuf(f, nullptr, _IOFBF, 1 << 20);
    }
    ~OutFile() { if (f) std::fclose(f); }
    void str(const std::string& s) { std::fwrite(s.data(), 1, s.size(), f); }
    void tab_str(const std::string& s) { std::fputc('\t', f); str(s); }
    void tab_int(int v) { std::fprintf(f, "\t%d", v); }
    void nl() { std::fputc('\n', f); }
};

// Main.java:697-713
inline void write_pedigree_file(const std::string& path, const Population& pop) {
    OutFile out(path);
    out.str("Name\tParent1\tParent2"); out.nl();
    for (size_t i = 0; i < pop.ind.size(); i++) {
        out.str(pop.ind_name[i]);
        if (pop.ind[i].is_founder()) { out.tab_str(pop.missing); out.tab_str(pop.missing); }
        else { out.tab_str(pop.ind_name[pop.ind[i].parent[0]]); out.tab_str(pop.ind_name[pop.ind[i].parent[1]]); }
        out.nl();
    }
}
// Main.java:1100-1136
inline void write_haplostruct_files(const std::string& base, const Population& pop) {
    OutFile hsa(base + ".hsa"), hsb(base + ".hsb");
    int segcount = 0;
    for (auto& x : pop.ind) for (auto& cv : x.hs) for (auto& h : cv) segcount = std::max(segcount, h.count());
    for (size_t i = 0; i < pop.ind.size(); i++)
        for (int c = 0; c < pop.C(); c++)
            for (int p = 0; p < pop.P(); p++) {
                hsa.str(pop.ind_name[i]); hsa.tab_str(pop.model.chrom[c].name); hsa.tab_int(p + 1);
                const HS& h = pop.ind[i].hs[c][p];
                for (int s = 0; s < h.count(); s++) {
                    hsa.tab_int(h.founder[s]);
                    if (s > 0) hsb.tab_str(java_double_to_string(h.pos[s] * 100.0));
                }
                for (int s = h.count(); s < segcount; s++) { hsa.tab_str(pop.missing); hsb.tab_str(pop.missing); }
                hsa.nl(); hsb.nl();
            }
}
// Main.java:1138-1164
inline void write_meiotic_configs_file(const std::string& path, const Population& pop) {
    OutFile out(path);
    for (size_t i = 0; i < pop.ind.size(); i++)
        for (int c = 0; c < pop.C(); c++) {
            out.str(pop.ind_name[i]); out.tab_str(pop.model.chrom[c].name);
            const Individual& x = pop.ind[i];
            if (!x.has_parconfig) {
                for (int k = 0; k < 2 * (pop.P() + 1); k++) out.tab_str(pop.missing);
            } else {
                for (int par = 0; par < 2; par++) {
                    out.tab_int(x.parconfig[par][c].quad);
                    for (int p = 0; p < pop.P(); p++) out.tab_int(x.parconfig[par][c].chromseq[p] + 1);
                }
            }
            out.nl();
        }
}
// Main.java:903-939
inline void write_genotypes_file(const std::string& path, bool founders, Population& pop, std::string* log = nullptr) {
    OutFile out(path);
    out.str("marker");
    for (size_t i = 0; i < pop.ind.size(); i++)
        for (int h = 0; h < pop.P(); h++) out.tab_str(pop.ind_name[i] + "_" + std::to_string(h + 1));
    out.nl();
    for (int c = 0; c < pop.C(); c++) {
        if (log) *log += "chrom " + pop.model.chrom[c].name + "\n";  // System.out.println("chrom "+...)
        for (auto& l : pop.locus[c]) {
            out.str(l.name);
            for (size_t i = 0; i < pop.ind.size(); i++)
                for (int h = 0; h < pop.P(); h++) {
                    int f = pop.founder_at((int)i, c, h, l.position);
                    if (founders) out.tab_int(f); else out.tab_str(l.allele_name[f]);
                }
            out.nl();
        }
    }
}
constexpr int MIN_ALLELE = -1, MAX_ALLELE = -2;  // Main.java:941-942
// Main.java:969-1022
inline void write_alleledose_file(const std::string& path, int which, bool actual, Population& pop, std::string* log = nullptr) {
    OutFile out(path);
    out.str("marker");
    for (auto& n : pop.ind_name) out.tab_str(n);
    out.nl();
    for (int c = 0; c < pop.C(); c++) {
        if (log) *log += "chrom " + pop.model.chrom[c].name + "\n";  // System.out.println("chrom "+...)
        for (auto& l : pop.locus[c]) {
            out.str(l.name);
            std::string match;
            if (actual) {
                if (which == MIN_ALLELE) match = l.uniques().front();
                else if (which == MAX_ALLELE) match = l.uniques().back();
                else { if (which < 0 || which >= pop.founder_allele_count) which = 0; match = l.allele_name[which]; }
            }
            for (size_t i = 0; i < pop.ind.size(); i++) {
                int dose = 0;
                for (int h = 0; h < pop.P(); h++) {
                    int f = pop.founder_at((int)i, c, h, l.position);
                    if (actual ? (match == l.allele_name[f]) : (which == f)) dose++;
                }
                out.tab_int(dose);
            }
            out.nl();
        }
    }
}
// Main.java:1037-1088
inline void write_all_alleledose_file(const std::string& path, bool actual, Population& pop, std::string* log = nullptr) {
    OutFile out(path);
    out.str("marker\tallele");
    for (auto& n : pop.ind_name) out.tab_str(n);
    out.nl();
    for (int c = 0; c < pop.C(); c++) {
        if (log) *log += "chrom " + pop.model.chrom[c].name + "\n";  // System.out.println("chrom "+...)
        for (auto& l : pop.locus[c]) {
            const int rows = actual ? (int)l.uniques().size() : pop.founder_allele_count;
            for (int a = 0; a < rows; a++) {
                out.str(l.name);
                out.tab_str(actual ? l.uniques()[a] : std::to_string(a));
                for (size_t i = 0; i < pop.ind.size(); i++) {
                    int dose = 0;
                    for (int h = 0; h < pop.P(); h++) {
                        int f = pop.founder_at((int)i, c, h, l.position);
                        if (actual ? (l.uniques()[a] == l.allele_name[f]) : (a == f)) dose++;
                    }
                    out.tab_int(dose);
                }
                out.nl();
            }
        }
    }
}

// ---------------------------------------------------------------- Main.simulate (normal mode)
// Main.java:75-450 without TEST mode. `log` receives what the reference prints to stdout.
// rng_mode 0 = java.util.Random(SEED) (the reference), 1 = Philox streams (replicate = 0).
inline int run_parameter_file(const std::string& par_path, int rng_mode, std::string& log) {
    auto say = [&](const std::string& s) { log += s; log += "\n"; };
    std::string base_name = par_path.substr(par_path.find_last_of('/') == std::string::npos ? 0 : par_path.find_last_of('/') + 1);
    // java.io.File semantics: relative names resolve against the working directory, not the .par file
    auto resolve = [&](const std::string& p) { return p; };
    say("Parameter file=" + base_name);
    std::map<std::string, std::string> par;
    std::string msg;
    if (!read_parameter_file(par_path, par, msg)) { say(msg); return 0; }
    if (par.count("TEST") && to_boolean(par["TEST"])) { say("TEST mode is out of scope for this restatement"); return 0; }
    bool interf = false;
    if (!par.count("PLOIDY")) { say(base_name + ": no valid PLOIDY found"); return 0; }
    if (par.count("MAPFUNCTION")) interf = upper(par["MAPFUNCTION"]) == "KOSAMBI";
    else if (!par.count("HAPLOSTRUCT")) { say(base_name + ": no valid MAPFUNCTION found"); return 0; }
    Population pop;
    Model& m = pop.model;
    m.ploidy = std::atoi(par["PLOIDY"].c_str());
    m.chiasma_interference = interf;
    if (par.count("PARALLELMULTIVALENTS")) m.parallel_multivalents = std::strtod(par["PARALLELMULTIVALENTS"].c_str(), nullptr);
    if (par.count("PAIREDCENTROMERES")) m.paired_centromeres = std::strtod(par["PAIREDCENTROMERES"].c_str(), nullptr);
    if (par.count("ALLOWNOCHIASMATA")) m.allow_no_recomb = to_boolean(par["ALLOWNOCHIASMATA"]);
    if (par.count("ALLOWNORECOMBINATION")) m.allow_no_recomb = to_boolean(par["ALLOWNORECOMBINATION"]);
    if (par.count("NATURALPAIRING")) m.natural_pairing = to_boolean(par["NATURALPAIRING"]);
    if (par.count("MISSING")) pop.missing = par["MISSING"];
    int32_t seed = par.count("SEED") ? (int32_t)std::atoi(par["SEED"].c_str()) : 12345;
    auto check_infile = [&](const std::string& descr, const std::string& p) {  // Main.java:58-71
        std::ifstream f(resolve(p));
        if (f.good()) return true;
        say(descr + " " + p + " not found or not readable");
        return false;
    };
    auto can_create = [&](const std::string& p) {  // Main.java:1449-1462
        std::remove(p.c_str());
        std::ofstream f(p);
        if (!f.good()) return false;
        f.close();
        return std::remove(p.c_str()) == 0;
    };
    {   // Main.java:142-160: a later message overwrites an earlier one
        std::string e;
        if (!par.count("CHROMFILE")) e = base_name + ": no CHROMFILE specified";
        if (!par.count("OUTPUT")) e = base_name + ": no OUTPUT specified";
        else if (!can_create(resolve(par["OUTPUT"]) + ".hsa") || !can_create(resolve(par["OUTPUT"]) + ".hsb") ||
                 !can_create(resolve(par["OUTPUT"]) + "_meioticconfigs.dat"))
            e = base_name + ": cannot create " + par["OUTPUT"] + ".*";
        if (!e.empty()) { say(e); return 0; }
    }
    const std::string output = resolve(par["OUTPUT"]);
    if (rng_mode == 0) pop.rng_holder.reset(new JavaRandom((int64_t)seed));
    else {
        pop.rng_holder.reset(new PhiloxStream((uint32_t)seed, 0));
        m.twin = true; m.twin_seed = (uint64_t)seed; m.twin_replicate = 0;
    }
    m.rng = pop.rng_holder.get();
    if (!check_infile("CHROMFILE", par["CHROMFILE"])) return 0;
    try {
        read_chromosome_file(resolve(par["CHROMFILE"]), pop);
        if (pop.C() == 0) throw std::runtime_error("CHROMFILE: no chromosomes");
        if (!m.allow_no_recomb)
            for (auto& ch : m.chrom)
                if (ch.length() < 0.70)
                    throw std::runtime_error("if ALLOWNORECOMBINATION is false the minimum chromosome length is 70.0 cM");
    } catch (std::exception& ex) {
        say("Error reading CHROMFILE " + par["CHROMFILE"] + ": " + ex.what());
        return 0;
    }
    std::string err;
    if (par.count("PEDFILE")) {
        if (!check_infile("PEDFILE", par["PEDFILE"])) return 0;
        auto orig = read_pedigree_file(resolve(par["PEDFILE"]), err);
        if (!err.empty()) { say("Error reading PEDFILE " + par["PEDFILE"] + ": " + err); return 0; }
        try {
            auto sorted = sort_pedigree(orig, pop.missing);
            // PopulationData.createPopulation returns its failure as a string (PopulationData.java:432-464):
            // the (partial) pedigree is still written and the bare message is printed last (Main.java:244-258,444)
            std::string create_err;
            try { create_population(pop, sorted); } catch (std::exception& ex) { create_err = ex.what(); }
            bool changed = false;
            for (size_t i = 0; i < sorted.size(); i++) if (sorted[i][0] != orig[i][0]) changed = true;
            if (changed) { say("write file " + par["OUTPUT"] + ".ped"); write_pedigree_file(output + ".ped", pop); }
            if (!create_err.empty()) { say(create_err); return 0; }
        } catch (std::exception& ex) { say(std::string("Error in pedigree: ") + ex.what()); return 0; }
    } else {
        if (!par.count("POPTYPE") || !par.count("POPSIZE")) {
            say("Either a valid PEDFILE or both POPTYPE and POPSIZE must be specified");
            return 0;
        }
        auto ped = generate_standard_population(upper(par["POPTYPE"]), std::atoi(par["POPSIZE"].c_str()), pop.missing);
        if (ped.empty()) { say("Invalid population type or population size==0"); return 0; }
        create_population(pop, ped);
        say("write file " + par["OUTPUT"] + ".ped");
        write_pedigree_file(output + ".ped", pop);
    }
    int maxfounderallele = -1;
    size_t hapstruct_read = 0;
    if (par.count("HAPLOSTRUCT")) {
        if (!check_infile("HAPLOSTRUCT file", par["HAPLOSTRUCT"] + ".hsa") ||
            !check_infile("HAPLOSTRUCT file", par["HAPLOSTRUCT"] + ".hsb")) return 0;
        try { maxfounderallele = read_haplostruct_files(resolve(par["HAPLOSTRUCT"]), pop); }
        catch (std::exception& ex) { say(std::string("Error reading HAPLOSTRUCT file ") + ex.what()); return 0; }
    }
    for (auto& x : pop.ind) if (x.has_hs) hapstruct_read++;
    try {
        if (hapstruct_read > 0 && hapstruct_read < pop.ind.size()) {
            std::string warn;
            read_meioticconfigs_file(resolve(par["HAPLOSTRUCT"]), pop, warn);
            if (!warn.empty()) say(warn);
        }
        if (hapstruct_read < pop.ind.size()) {
            MeiosisEngine eng(m);
            eng.stats = &pop.stats;
            eng.simulate_individuals(pop.ind, maxfounderallele);
            say("write file " + par["OUTPUT"] + ".hsa");
            say("write file " + par["OUTPUT"] + ".hsb");
            write_haplostruct_files(output, pop);
            say("write file " + par["OUTPUT"] + "_meioticconfigs.dat");
            write_meiotic_configs_file(output + "_meioticconfigs.dat", pop);
        }
    } catch (std::exception& ex) { say(ex.what()); return 0; }
    if (par.count("MAPFILE")) {
        if (!check_infile("MAPFILE", par["MAPFILE"])) return 0;
        try { read_map_file(resolve(par["MAPFILE"]), pop); }
        catch (std::exception& ex) { say("Error reading MAPFILE " + par["MAPFILE"] + ": " + ex.what()); return 0; }
        bool founder_genotypes = false;
        if (par.count("FOUNDERFILE")) {
            if (!check_infile("FOUNDERFILE", par["FOUNDERFILE"])) return 0;
            try { read_founder_genotypes_file(resolve(par["FOUNDERFILE"]), pop); founder_genotypes = true; }
            catch (std::exception& ex) { say("Error reading FOUNDERFILE " + par["FOUNDERFILE"] + ": " + ex.what()); return 0; }
        }
        try {
            int which = 0;
            if (founder_genotypes) {
                say("write file " + par["OUTPUT"] + "_genotypes.dat");
                write_genotypes_file(output + "_genotypes.dat", false, pop, &log);
                which = MAX_ALLELE;
            }
            say("write file " + par["OUTPUT"] + "_founderalleles.dat");
            write_genotypes_file(output + "_founderalleles.dat", true, pop, &log);
            say("write file " + par["OUTPUT"] + "_alleledose.dat");
            write_alleledose_file(output + "_alleledose.dat", which, founder_genotypes, pop, &log);
            if (!founder_genotypes || pop.max_allele_count() > 2) {
                say("write file " + par["OUTPUT"] + "_allAlleledose.dat");
                write_all_alleledose_file(output + "_allAlleledose.dat", founder_genotypes, pop, &log);
            }
        } catch (std::exception& ex) { say(std::string("Error writing files: ") + ex.what()); return 0; }
    } else if (par.count("HAPLOSTRUCT") && hapstruct_read == pop.ind.size()) {
        say(base_name + ": when HAPLOSTRUCT is given also MAPFILE must be specified");
    }
    return 0;
}

}  // namespace orc
