// Output side of the libpsim host: PedigreeSim's .ped / .hsa / .hsb / _meioticconfigs.dat /
// _genotypes.dat / _founderalleles.dat / _alleledose.dat / _allAlleledose.dat writers
// (Main.java:697-713,903-1164). The reference prints every cell through a PrintWriter; here the
// binary tables that libpsim emitted for a block of loci are formatted row-parallel on all host
// threads (token tables, no per-cell allocation) and appended to the files in order, so the text
// side keeps pace with the PCIe copy of the tables.
#pragma once

#include <thread>

#include "popfiles.hpp"

namespace psimhost {

inline int host_threads() {
    static int n = [] {
        const char* e = std::getenv("PSIM_HOST_THREADS");
        int v = e ? std::atoi(e) : (int)std::thread::hardware_concurrency();
        return std::max(1, std::min(v, 64));
    }();
    return n;
}

// Runs fn(part, begin, end) over [0, n) split into contiguous parts, one per thread.
template <class F>
inline int parallel_parts(int64_t n, int64_t min_per_part, F&& fn) {
    int parts = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), n / std::max<int64_t>(1, min_per_part)));
    if (parts == 1) {
        fn(0, (int64_t)0, n);
        return 1;
    }
    std::vector<std::thread> th;
    for (int p = 0; p < parts; p++) {
        int64_t b = n * p / parts, e = n * (p + 1) / parts;
        th.emplace_back([&fn, p, b, e] { fn(p, b, e); });
    }
    for (auto& t : th) t.join();
    return parts;
}

// "\t" + decimal of every value a u8 table can hold, packed for one unaligned 4-byte store.
struct TabTokens {
    uint32_t word[256];
    uint8_t len[256];
    TabTokens() {
        for (int v = 0; v < 256; v++) {
            char b[4] = {'\t', 0, 0, 0};
            int n = 1;
            if (v >= 100) b[n++] = (char)('0' + v / 100);
            if (v >= 10) b[n++] = (char)('0' + v / 10 % 10);
            b[n++] = (char)('0' + v % 10);
            std::memcpy(&word[v], b, 4);
            len[v] = (uint8_t)n;
        }
    }
};
inline const TabTokens& tab_tokens() {
    static TabTokens t;
    return t;
}

// Appends "\t<v>" for each u8 cell of a row.
inline void format_u8_row(const uint8_t* cells, int64_t n, std::string& o) {
    const TabTokens& T = tab_tokens();
    size_t at = o.size();
    o.resize(at + (size_t)n * 4 + 4);
    char* p = &o[at];
    for (int64_t i = 0; i < n; i++) {
        std::memcpy(p, &T.word[cells[i]], 4);
        p += T.len[cells[i]];
    }
    o.resize((size_t)(p - o.data()));
}
inline void format_u16_row(const uint16_t* cells, int64_t n, std::string& o) {
    char b[8];
    for (int64_t i = 0; i < n; i++) {
        auto r = std::to_chars(b, b + 8, (unsigned)cells[i]);
        o.push_back('\t');
        o.append(b, (size_t)(r.ptr - b));
    }
}

// Main.writePedigreeFile Main.java:697-713
inline void write_pedigree_file(const std::string& path, const Pedigree& pd, const Settings& st) {
    OutFile f(path);
    std::string o = "Name\tParent1\tParent2\n";
    for (int64_t i = 0; i < pd.count(); i++) {
        o += pd.name[i];
        if (pd.parent0[i] < 0) {
            o += "\t" + st.missing + "\t" + st.missing + "\n";
        } else {
            o += "\t" + pd.name[pd.parent0[i]] + "\t" + pd.name[pd.parent1[i]] + "\n";
        }
        if (o.size() > (1u << 20)) {
            f.write(o);
            o.clear();
        }
    }
    f.write(o);
}

// Main.writeHaploStructFiles Main.java:1100-1136 for the individuals [first, first+count) whose
// haplostructs are in `hs` (CSR, .hsa order). segcount = the population-wide maximum.
inline void append_haplostruct_text(OutFile& hsa, OutFile& hsb, const Pedigree& pd, const Chromosomes& ch,
                                    const Settings& st, int64_t first, int64_t count, const HaploBlock& hs,
                                    int64_t segcount) {
    const int C = ch.count(), P = st.ploidy;
    std::vector<std::string> ta((size_t)host_threads()), tb((size_t)host_threads());
    int parts = parallel_parts(count, 64, [&](int part, int64_t b, int64_t e) {
        std::string& a = ta[part];
        std::string& bb = tb[part];
        char num[16];
        for (int64_t i = b; i < e; i++) {
            const std::string& nm = pd.name[first + i];
            for (int c = 0; c < C; c++)
                for (int p = 0; p < P; p++) {
                    int64_t h = (i * C + c) * P + p;
                    a += nm;
                    a.push_back('\t');
                    a += ch.name[c];
                    a.push_back('\t');
                    auto r = std::to_chars(num, num + 16, p + 1);
                    a.append(num, (size_t)(r.ptr - num));
                    int64_t s0 = hs.seg_off[h], n = hs.seg_off[h + 1] - s0;
                    for (int64_t s = 0; s < n; s++) {
                        a.push_back('\t');
                        r = std::to_chars(num, num + 16, hs.founder[s0 + s]);
                        a.append(num, (size_t)(r.ptr - num));
                        if (s > 0) {
                            bb.push_back('\t');
                            java_double(hs.pos[s0 + s] * 100.0, bb);  // Morgan to cM
                        }
                    }
                    for (int64_t s = n; s < segcount; s++) {
                        a.push_back('\t');
                        a += st.missing;
                        bb.push_back('\t');
                        bb += st.missing;
                    }
                    a.push_back('\n');
                    bb.push_back('\n');
                }
        }
    });
    for (int p = 0; p < parts; p++) {
        hsa.write(ta[p]);
        hsb.write(tb[p]);
    }
}

// Main.writeMeioticConfigsFile Main.java:1138-1164
inline void write_meiotic_configs_file(const std::string& path, const Pedigree& pd, const Chromosomes& ch,
                                       const Settings& st, const MeioticConfigs& mc) {
    OutFile f(path);
    const int C = ch.count(), P = st.ploidy;
    std::vector<std::string> t((size_t)host_threads());
    int parts = parallel_parts(pd.count(), 256, [&](int part, int64_t b, int64_t e) {
        std::string& o = t[part];
        for (int64_t i = b; i < e; i++)
            for (int c = 0; c < C; c++) {
                o += pd.name[i];
                o.push_back('\t');
                o += ch.name[c];
                if (!mc.known[i]) {
                    for (int k = 0; k < 2 * (P + 1); k++) {
                        o.push_back('\t');
                        o += st.missing;
                    }
                } else {
                    for (int par = 0; par < 2; par++) {
                        size_t q = ((size_t)i * C + c) * 2 + par;
                        o.push_back('\t');
                        o += std::to_string(mc.quad[q]);
                        for (int h = 0; h < P; h++) {
                            o.push_back('\t');
                            o += std::to_string(mc.chromseq[q * P + h] + 1);
                        }
                    }
                }
                o.push_back('\n');
            }
    });
    for (int p = 0; p < parts; p++) f.write(t[p]);
}

// Header lines of the three genotype tables (Main.java:907-913,976-980,1044-1048)
inline void write_table_header(OutFile& f, const Pedigree& pd, int ploidy_columns, bool allele_column) {
    std::string o = allele_column ? "marker\tallele" : "marker";
    for (int64_t i = 0; i < pd.count(); i++) {
        if (ploidy_columns == 0) {
            o.push_back('\t');
            o += pd.name[i];
        } else {
            for (int h = 0; h < ploidy_columns; h++) {
                o.push_back('\t');
                o += pd.name[i];
                o.push_back('_');
                o += std::to_string(h + 1);
            }
        }
        if (o.size() > (1u << 20)) {
            f.write(o);
            o.clear();
        }
    }
    o.push_back('\n');
    f.write(o);
}

// Rows [0, rows) of a block: label(r, out) writes the row label, cells(r, out) the "\t"-led cells.
template <class Label, class Cells>
inline void append_rows(OutFile& f, int64_t rows, Label&& label, Cells&& cells) {
    std::vector<std::string> t((size_t)host_threads());
    int parts = parallel_parts(rows, 1, [&](int part, int64_t b, int64_t e) {
        std::string& o = t[part];
        for (int64_t r = b; r < e; r++) {
            label(r, o);
            cells(r, o);
            o.push_back('\n');
        }
    });
    for (int p = 0; p < parts; p++) f.write(t[p]);
}

}  // namespace psimhost
