// pedigreesim — the host program around libpsim: `pedigreesim x.par` is the drop-in for
// `java -jar PedigreeSim.jar x.par` on the meiosis + genotype-emission path (Main.simulate,
// Main.java:75-450, normal mode: simulate / replay / extend). It keeps the reference's
// .par/.chrom/.map/.ped/.gen/.hsa/.hsb inputs, its output files and its stdout messages, walks
// the pedigree on the host and hands every generation's meioses and every block of the genotype
// tables to the CUDA library through the C ABI of include/psim.h — the same calls a Java host
// makes through FFM/JNI. There is no CPU path: without libpsim and a GPU the
// simulation and emission steps fail with the library's message.
//
//   pedigreesim x.par [--parse-only] [--device N] [--block-mib M]
//
// --parse-only stops before the first libpsim call (reads and validates every input, writes the
// sorted/generated .ped) — used by the CPU-only tests of the host logic.
#include <chrono>
#include <random>

#include "../../include/psim.h"
#include "writers.hpp"

using namespace psimhost;

namespace {

struct Options {
    std::string par_path = "PedigreeSim.par";
    bool parse_only = false;
    int device = 0;
    int64_t block_bytes = 512LL << 20;
};

void say(const std::string& s) {
    std::fputs(s.c_str(), stdout);
    std::fputc('\n', stdout);
}

std::string base_name(const std::string& p) {
    size_t k = p.find_last_of('/');
    return k == std::string::npos ? p : p.substr(k + 1);
}

bool check_infile(const std::string& descr, const std::string& path) {  // Main.checkInfile Main.java:58-71
    if (file_readable(path)) return true;
    say(descr + " " + path + " not found or not readable");
    return false;
}

bool can_create_file(const std::string& path) {  // Main.canCreateFile Main.java:1449-1462
    std::remove(path.c_str());
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fclose(f);
    return std::remove(path.c_str()) == 0;
}

struct Lib {  // the C ABI with errors turned into the exceptions the reference would print
    psim_ctx* ctx = nullptr;
    ~Lib() {
        if (ctx) psim_destroy(ctx);
    }
    void chk(int rc) const {
        if (rc != PSIM_OK) throw HostError(std::string("libpsim: ") + psim_last_error(ctx));
    }
};

// Founder allele at the head of homolog 0, per chromosome (HaploStruct.getFounderAt(headPos),
// HaploStruct.java:107-112,135-142): what .gen columns are keyed on (Main.java:869-876).
int32_t head_allele(const HaploBlock& hs, int64_t homolog, double head) {
    int64_t s0 = hs.seg_off[homolog], seg = hs.seg_off[homolog + 1] - s0 - 1;
    while (seg > 0 && hs.pos[s0 + seg] > head) seg--;
    return hs.founder[s0 + seg];
}

int run(const Options& opt) {
    using clk = std::chrono::steady_clock;
    if (!file_readable(opt.par_path)) {
        say("PedigreeSim: parameter file " + opt.par_path + " not found or not readable.");
        return 0;
    }
    const std::string par_name = base_name(opt.par_path);
    say("Parameter file=" + par_name);
    ParMap par;
    {
        std::string log;
        bool ok = read_parameter_file(opt.par_path, par_name, par, log);
        std::fputs(log.c_str(), stdout);
        if (!ok) return 0;
    }
    auto has = [&](const char* k) { return par.count(k) != 0; };
    const bool test = has("TEST") && to_boolean(par["TEST"]);
    Settings st;
    if (!has("PLOIDY")) {
        say(par_name + ": no valid PLOIDY found");
        return 0;
    }
    if (has("MAPFUNCTION")) {
        st.kosambi = upper(par["MAPFUNCTION"]) == "KOSAMBI";
    } else if (!has("HAPLOSTRUCT")) {
        say(par_name + ": no valid MAPFUNCTION found");
        return 0;
    }
    {
        int32_t v = 2;
        parse_int32(par["PLOIDY"], v);
        st.ploidy = v;
    }
    if (has("PARALLELMULTIVALENTS")) parse_double(par["PARALLELMULTIVALENTS"], st.parallel_multivalents);
    if (has("PAIREDCENTROMERES")) parse_double(par["PAIREDCENTROMERES"], st.paired_centromeres);
    if (has("ALLOWNOCHIASMATA")) st.allow_no_recomb = to_boolean(par["ALLOWNOCHIASMATA"]);
    if (has("ALLOWNORECOMBINATION")) st.allow_no_recomb = to_boolean(par["ALLOWNORECOMBINATION"]);
    if (has("NATURALPAIRING")) st.natural_pairing = to_boolean(par["NATURALPAIRING"]);
    if (has("MISSING")) st.missing = par["MISSING"];
    if (has("SEED"))
        parse_int32(par["SEED"], st.seed);
    else
        st.seed = (int32_t)std::random_device{}();
    std::string err;
    if (!has("CHROMFILE")) err = par_name + ": no CHROMFILE specified";
    if (!has("OUTPUT")) {
        err = par_name + ": no OUTPUT specified";
    } else {
        const std::string& o = par["OUTPUT"];
        if (!can_create_file(o + ".hsa") || !can_create_file(o + ".hsb") || !can_create_file(o + "_meioticconfigs.dat"))
            err = par_name + ": cannot create " + o + ".*";
    }
    if (!err.empty()) {
        say(err);
        return 0;
    }
    const std::string out_base = par["OUTPUT"];

    Chromosomes ch;
    try {
        if (!check_infile("CHROMFILE", par["CHROMFILE"])) return 0;
        read_chromosome_file(par["CHROMFILE"], st, ch);
        if (ch.count() == 0) throw HostError("CHROMFILE: no chromosomes");
        if (!st.allow_no_recomb)
            for (int c = 0; c < ch.count(); c++)
                if (ch.tail[c] - ch.head[c] < 0.70)
                    throw HostError("if ALLOWNORECOMBINATION is false the minimum chromosome length is 70.0 cM");
    } catch (const HostError& ex) {
        say("Error reading CHROMFILE " + par["CHROMFILE"] + ": " + ex.what());
        return 0;
    }
    if (test) {
        say("TEST mode (Main.testMeiosis) is outside libpsim's path: run the reference jar for the TEST report");
        return 0;
    }
    if (has("2NGAMETES")) {
        say("2NGAMETES (unreduced FDR/SDR gametes) is not supported by libpsim");
        return 0;
    }

    // ---- pedigree
    Pedigree pd;
    if (has("PEDFILE")) {
        if (!check_infile("PEDFILE", par["PEDFILE"])) return 0;
        std::string perr;
        PedRows orig = read_pedigree_file(par["PEDFILE"], perr);
        if (!perr.empty()) {
            say("Error reading PEDFILE " + par["PEDFILE"] + ": " + perr);
            return 0;
        }
        try {
            PedRows sorted = sort_pedigree(orig, st.missing);
            std::string cerr = create_population(sorted, st, pd);
            bool changed = false;
            for (size_t i = 0; i < sorted.size() && !changed; i++) changed = sorted[i][0] != orig[i][0];
            if (changed) {
                say("write file " + out_base + ".ped");
                write_pedigree_file(out_base + ".ped", pd, st);
            }
            if (!cerr.empty()) {  // Main.java:244-258,444: reported last, without a prefix
                say(cerr);
                return 0;
            }
        } catch (const HostError& ex) {
            say(std::string("Error in pedigree: ") + ex.what());
            return 0;
        }
    } else {
        if (!has("POPTYPE") || !has("POPSIZE")) {
            say("Either a valid PEDFILE or both POPTYPE and POPSIZE must be specified");
            return 0;
        }
        int32_t popsize = 0;
        parse_int32(par["POPSIZE"], popsize);
        PedRows rows = generate_standard_population(upper(par["POPTYPE"]), popsize, st.missing);
        if (rows.empty()) {
            say("Invalid population type or population size==0");
            return 0;
        }
        std::string cerr = create_population(rows, st, pd);
        if (!cerr.empty()) {
            say(cerr);
            return 0;
        }
        say("write file " + out_base + ".ped");
        write_pedigree_file(out_base + ".ped", pd, st);
    }
    const int64_t N = pd.count();
    const int C = ch.count(), P = st.ploidy;

    // ---- recorded haplostructs (replay / extend)
    PresetHaplo preset;
    if (has("HAPLOSTRUCT")) {
        if (!check_infile("HAPLOSTRUCT file", par["HAPLOSTRUCT"] + ".hsa") ||
            !check_infile("HAPLOSTRUCT file", par["HAPLOSTRUCT"] + ".hsb"))
            return 0;
        try {
            read_haplostruct_files(par["HAPLOSTRUCT"], st, ch, pd, preset);
        } catch (const HostError& ex) {
            say(std::string("Error reading HAPLOSTRUCT file ") + ex.what());
            return 0;
        }
    }
    const int64_t n_read = (int64_t)preset.ind.size();
    std::vector<int64_t> preset_slot((size_t)N, -1);  // position of an individual's block in `preset`
    for (int64_t k = 0; k < n_read; k++) preset_slot[preset.ind[k]] = k;
    MeioticConfigs mc;
    mc.known.assign((size_t)N, 0);
    mc.quad.assign((size_t)N * C * 2, 0);
    mc.chromseq.assign((size_t)N * C * 2 * P, 0);
    if (n_read > 0 && n_read < N) {
        try {
            std::string log;
            read_meioticconfigs_file(par["HAPLOSTRUCT"], st, ch, pd, mc, log);
            std::fputs(log.c_str(), stdout);
        } catch (const HostError& ex) {
            say(ex.what());  // the reference lets this exception escape simulate()
            return 0;
        }
    }
    const bool simulate_new = n_read < N;
    if (!simulate_new && !has("MAPFILE")) {
        say(par_name + ": when HAPLOSTRUCT is given also MAPFILE must be specified");
        return 0;
    }

    // founder alleles at the head of homolog 0: known on the host for every founder
    // (Individual.calcFounderHaplostruct Individual.java:155-171 / the recorded haplostructs)
    std::vector<int32_t> head_of((size_t)N * C, -1);
    {
        int next_allele = preset.max_founder_allele;
        for (int64_t i = 0; i < N; i++) {
            if (preset_slot[i] >= 0) {
                for (int c = 0; c < C; c++)
                    head_of[(size_t)i * C + c] = head_allele(preset.hs, (preset_slot[i] * C + c) * P, ch.head[c]);
            } else if (pd.parent0[i] < 0) {
                for (int c = 0; c < C; c++) head_of[(size_t)i * C + c] = next_allele + 1;
                next_allele += P;
            }
        }
    }

    // ---- map and founder genotypes are validated before any GPU work in --parse-only mode;
    // in a normal run the reference reads them after the simulation, and so do we (same messages
    // in the same order), but the parsing code is the same
    LinkageMap mp;
    bool founder_genotypes = false;
    auto read_map_and_founders = [&]() -> bool {
        if (!has("MAPFILE")) return true;
        if (!check_infile("MAPFILE", par["MAPFILE"])) return false;
        try {
            read_map_file(par["MAPFILE"], st, ch, mp);
        } catch (const HostError& ex) {
            say("Error reading MAPFILE " + par["MAPFILE"] + ": " + ex.what());
            return false;
        }
        if (has("FOUNDERFILE")) {
            if (!check_infile("FOUNDERFILE", par["FOUNDERFILE"])) return false;
            try {
                read_founder_genotypes_file(par["FOUNDERFILE"], st, ch, pd, head_of, mp);
                founder_genotypes = true;
            } catch (const HostError& ex) {
                say("Error reading FOUNDERFILE " + par["FOUNDERFILE"] + ": " + ex.what());
                return false;
            }
        }
        return true;
    };
    if (opt.parse_only) {
        if (simulate_new) {
            say("write file " + out_base + ".hsa");
            say("write file " + out_base + ".hsb");
            say("write file " + out_base + "_meioticconfigs.dat");
        }
        if (!read_map_and_founders()) return 0;
        std::printf("parse-only: %lld individuals, %d founder alleles, %d chromosomes, %lld loci, %lld recorded haplostructs\n",
                    (long long)N, pd.founder_allele_count, C, (long long)mp.total(), (long long)n_read);
        return 0;
    }

    // ---- libpsim context (seam A inputs)
    Lib lib;
    auto t0 = clk::now();
    auto t_ctx = t0, t_sim = t0;   // (PSIM_HOST_TIMING: context creation, simulation, the segment files)
    try {
        psim_config cfg{};
        cfg.ploidy = P;
        cfg.chiasma_interference = st.kosambi;
        cfg.natural_pairing = st.natural_pairing;
        cfg.allow_no_recomb = st.allow_no_recomb;
        cfg.parallel_multivalents = st.parallel_multivalents;
        cfg.paired_centromeres = st.paired_centromeres;
        cfg.seed = st.seed;
        cfg.replicate_first = 0;
        cfg.replicate_count = 1;
        cfg.device = opt.device;
        if (psim_create(&cfg, &lib.ctx) != PSIM_OK) throw HostError(std::string("libpsim: ") + psim_last_error(nullptr));
        t_ctx = clk::now();
        lib.chk(psim_set_chromosomes(lib.ctx, C, ch.length.data(), ch.centromere.data(),
                                     P == 2 ? nullptr : ch.pref_pairing.data(), P == 2 ? nullptr : ch.frac_quad.data()));
        lib.chk(psim_set_pedigree(lib.ctx, N, pd.parent0.data(), pd.parent1.data()));
        // recorded haplostructs go in as runs of consecutive individuals
        if (n_read > 0) {
            std::vector<int64_t> order((size_t)n_read);
            for (int64_t k = 0; k < n_read; k++) order[k] = k;
            std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return preset.ind[a] < preset.ind[b]; });
            const int64_t per = (int64_t)C * P;
            for (int64_t a = 0; a < n_read;) {
                int64_t b = a + 1;
                while (b < n_read && preset.ind[order[b]] == preset.ind[order[b - 1]] + 1) b++;
                HaploBlock run;
                for (int64_t k = a; k < b; k++) {
                    int64_t h0 = order[k] * per;
                    for (int64_t h = h0; h < h0 + per; h++) {
                        int64_t s0 = preset.hs.seg_off[h], s1 = preset.hs.seg_off[h + 1];
                        run.founder.insert(run.founder.end(), preset.hs.founder.begin() + s0, preset.hs.founder.begin() + s1);
                        run.pos.insert(run.pos.end(), preset.hs.pos.begin() + s0, preset.hs.pos.begin() + s1);
                        run.seg_off.push_back((int64_t)run.founder.size());
                    }
                }
                lib.chk(psim_set_haplostructs(lib.ctx, 0, preset.ind[order[a]], b - a, run.seg_off.data(), run.founder.data(),
                                              run.pos.data()));
                a = b;
            }
        }
    } catch (const HostError& ex) {
        say(ex.what());
        return 0;
    }

    // ---- seam A: simulate what is missing, write .hsa/.hsb and _meioticconfigs.dat
    const int64_t ind_block = 1 << 16;
    if (simulate_new) {
        try {
            lib.chk(psim_simulate(lib.ctx, preset.max_founder_allele));
            t_sim = clk::now();
            say("write file " + out_base + ".hsa");
            say("write file " + out_base + ".hsb");
            // two passes over blocks of individuals: the widest haplostruct sets the column count
            int64_t segcount = 0;
            std::vector<HaploBlock> kept;
            const bool keep = N <= 4 * ind_block;
            for (int pass = 0; pass < 2; pass++) {
                std::unique_ptr<OutFile> hsa, hsb;
                if (pass == 1) {
                    hsa.reset(new OutFile(out_base + ".hsa"));
                    hsb.reset(new OutFile(out_base + ".hsb"));
                }
                size_t kb = 0;
                for (int64_t first = 0; first < N; first += ind_block, kb++) {
                    int64_t cnt = std::min(ind_block, N - first);
                    HaploBlock local;
                    HaploBlock* hb = &local;
                    if (pass == 0 || !keep) {
                        int64_t nseg = 0;
                        lib.chk(psim_segment_count(lib.ctx, 0, first, cnt, &nseg));
                        local.seg_off.assign((size_t)(cnt * C * P + 1), 0);
                        local.founder.resize((size_t)nseg);
                        local.pos.resize((size_t)nseg);
                        lib.chk(psim_get_haplostructs(lib.ctx, 0, first, cnt, local.seg_off.data(), local.founder.data(),
                                                      local.pos.data()));
                        if (pass == 0) {
                            for (size_t h = 0; h + 1 < local.seg_off.size(); h++)
                                segcount = std::max(segcount, local.seg_off[h + 1] - local.seg_off[h]);
                            if (keep) kept.push_back(std::move(local));
                            continue;
                        }
                    } else {
                        hb = &kept[kb];
                    }
                    append_haplostruct_text(*hsa, *hsb, pd, ch, st, first, cnt, *hb, segcount);
                }
            }
            // meiotic configurations of the individuals simulated now; recorded ones are kept
            for (int64_t first = 0; first < N; first += ind_block) {
                int64_t cnt = std::min(ind_block, N - first);
                std::vector<int32_t> q((size_t)cnt * C * 2), s((size_t)cnt * C * 2 * P);
                lib.chk(psim_get_meiotic_configs(lib.ctx, 0, first, cnt, q.data(), s.data()));
                for (int64_t i = 0; i < cnt; i++) {
                    if (preset_slot[first + i] >= 0 || pd.parent0[first + i] < 0) continue;
                    mc.known[first + i] = 1;
                    std::copy(q.begin() + i * C * 2, q.begin() + (i + 1) * C * 2, mc.quad.begin() + (first + i) * C * 2);
                    std::copy(s.begin() + i * C * 2 * P, s.begin() + (i + 1) * C * 2 * P,
                              mc.chromseq.begin() + (first + i) * C * 2 * P);
                }
            }
            say("write file " + out_base + "_meioticconfigs.dat");
            write_meiotic_configs_file(out_base + "_meioticconfigs.dat", pd, ch, st, mc);
        } catch (const HostError& ex) {
            say(ex.what());
            return 0;
        }
    }
    auto t1 = clk::now();

    // ---- seam B: genotype tables
    if (!has("MAPFILE")) return 0;
    if (!read_map_and_founders()) return 0;
    const int64_t L = mp.total();
    try {
        if (L == 0) throw HostError("no loci on the map");
        const int A = pd.founder_allele_count;
        lib.chk(psim_set_map(lib.ctx, mp.locus_off.data(), mp.flat_pos.data()));
        if (founder_genotypes) lib.chk(psim_set_allele_codes(lib.ctx, A, mp.code.data()));
        const bool all_dose = !founder_genotypes || mp.max_allele_count > 2;
        // the reference writes the files one after the other and reports each chromosome
        auto announce = [&](const std::string& name) {
            say("write file " + name);
            for (int c = 0; c < C; c++) say("chrom " + ch.name[c]);
        };
        std::unique_ptr<OutFile> f_gen, f_fnd, f_dose, f_all;
        if (founder_genotypes) {
            announce(out_base + "_genotypes.dat");
            f_gen.reset(new OutFile(out_base + "_genotypes.dat"));
            write_table_header(*f_gen, pd, P, false);
        }
        announce(out_base + "_founderalleles.dat");
        f_fnd.reset(new OutFile(out_base + "_founderalleles.dat"));
        write_table_header(*f_fnd, pd, P, false);
        announce(out_base + "_alleledose.dat");
        f_dose.reset(new OutFile(out_base + "_alleledose.dat"));
        write_table_header(*f_dose, pd, 0, false);
        if (all_dose) {
            announce(out_base + "_allAlleledose.dat");
            f_all.reset(new OutFile(out_base + "_allAlleledose.dat"));
            write_table_header(*f_all, pd, 0, true);
        }

        // The three big tables leave the GPU as finished text (psim_emit_text): one fused call per
        // block of loci formats every requested table on the device into page-locked buffers; a
        // writer thread appends block k to the files while block k+1 is produced.
        {
            std::vector<int64_t> off((size_t)L + 1, 0);
            std::string chars;
            for (int64_t l = 0; l < L; l++) {
                chars += *mp.flat_name[(size_t)l];
                off[(size_t)l + 1] = (int64_t)chars.size();
            }
            chars.push_back('\0');
            lib.chk(psim_set_locus_names(lib.ctx, off.data(), chars.data()));
        }
        int64_t label_max = 0, name_max = 0;
        for (int64_t l = 0; l < L; l++) label_max = std::max<int64_t>(label_max, (int64_t)mp.flat_name[(size_t)l]->size());
        if (founder_genotypes) {
            std::vector<int64_t> first((size_t)L + 1, 0), noff(1, 0);
            std::string chars;
            for (int64_t l = 0; l < L; l++) {
                for (const std::string& nm : mp.unique_names[(size_t)l]) {
                    chars += nm;
                    noff.push_back((int64_t)chars.size());
                    name_max = std::max<int64_t>(name_max, (int64_t)nm.size());
                }
                first[(size_t)l + 1] = (int64_t)noff.size() - 1;
            }
            chars.push_back('\0');
            lib.chk(psim_set_allele_names(lib.ctx, first.data(), noff.data(), chars.data()));
        }
        const int64_t cols = N * P;
        auto digits = [](int64_t v) { int d = 1; while (v >= 10) { v /= 10; d++; } return d; };
        const int64_t line_fnd = label_max + cols * (1 + digits(std::max(0, A - 1))) + 1;
        const int64_t line_gen = founder_genotypes ? label_max + cols * (1 + name_max) + 1 : 0;
        const int64_t line_dose = label_max + N * (1 + digits(P)) + 1;
        const int64_t rows_blk = std::max<int64_t>(1, std::min<int64_t>(L, opt.block_bytes / (line_fnd + line_gen + line_dose)));
        struct TextBufs { void *fnd = nullptr, *gen = nullptr, *dose = nullptr; psim_text_targets t{}; };
        TextBufs tb[2];
        for (TextBufs& x : tb) {
            lib.chk(psim_alloc_host(lib.ctx, (size_t)(rows_blk * line_fnd), &x.fnd));
            if (founder_genotypes) lib.chk(psim_alloc_host(lib.ctx, (size_t)(rows_blk * line_gen), &x.gen));
            lib.chk(psim_alloc_host(lib.ctx, (size_t)(rows_blk * line_dose), &x.dose));
        }
        std::thread writer;
        int cur = 0;
        try {
            for (int64_t l0 = 0; l0 < L; l0 += rows_blk, cur ^= 1) {
                const int64_t nr = std::min(rows_blk, L - l0);
                TextBufs& x = tb[cur];
                x.t = psim_text_targets{};
                x.t.founder = (char*)x.fnd;
                x.t.founder_capacity = rows_blk * line_fnd;
                x.t.allele = (char*)x.gen;
                x.t.allele_capacity = rows_blk * line_gen;
                x.t.dose = (char*)x.dose;
                x.t.dose_capacity = rows_blk * line_dose;
                x.t.dose_which = founder_genotypes ? PSIM_DOSE_MAX_ALLELE : 0;  // Main.java:402-418
                lib.chk(psim_emit_text(lib.ctx, 0, N, l0, nr, &x.t));
                if (writer.joinable()) writer.join();
                writer = std::thread([&, cur] {
                    const TextBufs& y = tb[cur];
                    if (f_gen) f_gen->write(y.t.allele, (size_t)y.t.allele_bytes);
                    f_fnd->write(y.t.founder, (size_t)y.t.founder_bytes);
                    f_dose->write(y.t.dose, (size_t)y.t.dose_bytes);
                });
            }
        } catch (...) {
            if (writer.joinable()) writer.join();
            throw;
        }
        if (writer.joinable()) writer.join();
        for (TextBufs& x : tb) {
            psim_free_host(lib.ctx, x.fnd);
            if (x.gen) psim_free_host(lib.ctx, x.gen);
            psim_free_host(lib.ctx, x.dose);
        }

        if (all_dose) {
            // rows per locus: founder alleles 0..A-1, or the locus' unique allele names in sorted order;
            // also formatted on the device (psim_emit_all_dose_text), same writer-thread pipeline
            const int64_t max_rows_per_locus = founder_genotypes ? mp.max_allele_count : A;
            const int64_t sub_max = founder_genotypes ? name_max : digits(std::max(0, A - 1));
            const int64_t line_all = label_max + 1 + sub_max + N * (1 + digits(P)) + 1;
            const int64_t loci_blk = std::max<int64_t>(1, std::min<int64_t>(L, opt.block_bytes / (max_rows_per_locus * line_all)));
            void* buf[2] = {nullptr, nullptr};
            int64_t bytes[2] = {0, 0};
            for (void*& b2 : buf) lib.chk(psim_alloc_host(lib.ctx, (size_t)(loci_blk * max_rows_per_locus * line_all), &b2));
            std::thread wr;
            int cur2 = 0;
            try {
                for (int64_t l0 = 0; l0 < L; l0 += loci_blk, cur2 ^= 1) {
                    const int64_t nl = std::min(loci_blk, L - l0);
                    lib.chk(psim_emit_all_dose_text(lib.ctx, 0, N, l0, nl, (char*)buf[cur2], loci_blk * max_rows_per_locus * line_all, &bytes[cur2]));
                    if (wr.joinable()) wr.join();
                    wr = std::thread([&, cur2] { f_all->write((const char*)buf[cur2], (size_t)bytes[cur2]); });
                }
            } catch (...) {
                if (wr.joinable()) wr.join();
                throw;
            }
            if (wr.joinable()) wr.join();
            for (void* b2 : buf) psim_free_host(lib.ctx, b2);
        }
    } catch (const HostError& ex) {
        say(std::string("Error writing files: ") + ex.what());
        return 0;
    }
    auto t2 = clk::now();
    if (std::getenv("PSIM_HOST_TIMING")) {
        psim_stats s{};
        psim_get_stats(lib.ctx, &s);
        std::fprintf(stderr, "[pedigreesim] simulate+write %.3f s (CUDA context %.3f s, inputs + psim_simulate %.3f s, .hsa/.hsb/_meioticconfigs.dat %.3f s), "
                             "emit+write %.3f s, %lld meioses, %lld libpsim kernel launches\n",
                     std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(t_ctx - t0).count(),
                     std::chrono::duration<double>(t_sim - t_ctx).count(), std::chrono::duration<double>(t1 - std::max(t_sim, t_ctx)).count(),
                     std::chrono::duration<double>(t2 - t1).count(), (long long)s.n_meioses, (long long)s.kernel_launches);
    }
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    Options opt;
    bool have_par = false, format_doubles = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--parse-only")
            opt.parse_only = true;
        else if (a == "--format-doubles")
            format_doubles = true;
        else if (a == "--device" && i + 1 < argc)
            opt.device = std::atoi(argv[++i]);
        else if (a == "--block-mib" && i + 1 < argc)
            opt.block_bytes = (int64_t)std::atoll(argv[++i]) << 20;
        else if (!have_par) {
            opt.par_path = a;
            have_par = true;
        }
    }
    if (format_doubles) {
        // self-check hook for the Double.toString twin: bit patterns (hex) in, Java text out
        char w[64];
        while (std::scanf("%63s", w) == 1) {
            uint64_t bits = std::strtoull(w, nullptr, 16);
            double d;
            std::memcpy(&d, &bits, 8);
            std::string o;
            java_double(d, o);
            say(o);
        }
        return 0;
    }
    try {
        return run(opt);
    } catch (const std::exception& ex) {
        say(ex.what());
        return 0;  // the reference reports every failure on stdout and exits 0 (Main.java:106-109,446-448)
    }
}
