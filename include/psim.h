/*
 * psim.h — C ABI of libpsim, the B200 (sm_100a) replacement for PedigreeSim's meiosis and
 * genotype-emission hot path.
 *
 * The reference (PedigreeSim 2.2, pure Java) has no FFI. The path sits behind two Java call seams
 * (file:line relative to /root/reference/src/PedigreeSim/):
 *
 *   Seam A, meiosis:  Main.simulateIndividuals(popdata, maxfounderallele)      Main.java:1504-1520
 *                     called from Main.simulate                                Main.java:358-361
 *   Seam B, emission: Main.writeGenotypesFile / writeAlleledoseFile /
 *                     writeAllAlleledoseFile                                   Main.java:903-1088
 *                     called from Main.simulate                                Main.java:399-426
 *
 * A Java host binds these entry points with java.lang.foreign downcall handles or JNI
 * (INTEGRATION.md shows the stub). Conventions:
 *   - plain pointers and sizes; the caller owns every host buffer it passes, the library copies
 *     in and owns all device memory; output buffers are caller-allocated (query sizes first);
 *   - every call returns 0 on success, otherwise a PSIM_E* code with text in psim_last_error().
 *     (The reference throws java.lang.Exception(message), Main.java:106-109,446-448; the Java
 *     stub rethrows the text.)
 *   - one context per host thread; results do not depend on thread or GPU count (counter-based
 *     Philox streams keyed by seed, replicate, individual, chromosome, parent);
 *   - positions are MORGAN, i.e. the host has already done the reference's `cM / 100.0`
 *     (Main.java:619-620,678,1275) in fp64;
 *   - homolog order in every haplostruct array is the .hsa file order
 *     [individual][chromosome][homolog] (Main.java:1112-1131);
 *   - there is no CPU fallback: without a CUDA device psim_create fails.
 */
#ifndef PSIM_H
#define PSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSIM_ABI_VERSION 2

typedef struct psim_ctx psim_ctx;

enum {
    PSIM_OK = 0,
    PSIM_EINVAL = 1,    /* bad argument or call order */
    PSIM_ECUDA = 2,     /* CUDA runtime error (no device, out of memory, launch failure) */
    PSIM_ECAPACITY = 3, /* a device-side buffer limit was exceeded (e.g. chiasmata per bivalent) */
    PSIM_ESTATE = 4     /* required input not set (pedigree, map, haplostructs ...) */
};

/* The model flags of PopulationData (PopulationData.java:187-222) plus the RNG identity. */
typedef struct psim_config {
    int32_t ploidy;                /* PLOIDY: positive, even, <= 16 (4-bit homolog numbers on the device) */
    int32_t chiasma_interference;  /* MAPFUNCTION: 0 = HALDANE, 1 = KOSAMBI */
    int32_t natural_pairing;       /* NATURALPAIRING */
    int32_t allow_no_recomb;       /* ALLOWNORECOMBINATION */
    double parallel_multivalents;  /* PARALLELMULTIVALENTS in [0,1] */
    double paired_centromeres;     /* PAIREDCENTROMERES in [0,1] */
    int64_t seed;                  /* SEED (the reference parses an int32, Main.java:139) */
    uint32_t replicate_first;      /* first replicate id; replicate r uses Philox key (seed, r) */
    uint32_t replicate_count;      /* >= 1: independent replicate pedigrees simulated side by side */
    int32_t device;                /* CUDA device ordinal */
    int32_t unreduced_gametes;     /* 2NGAMETES (Main.java:179-185; Individual.java:341-383): 0 = haploid gametes,
                                      1 = FDR, 2 = SDR. Non-zero: every non-founder is formed from ONE unreduced (2n)
                                      gamete of its parent 0 alone (all P homologs; parent 1 is not used, its meiotic
                                      configuration stays missing) - the reference analyses such gametes in TEST mode only */
} psim_config;

int psim_abi_version(void);

/* Lifetime. */
int psim_create(const psim_config* cfg, psim_ctx** out);
void psim_destroy(psim_ctx* ctx);
const char* psim_last_error(const psim_ctx* ctx); /* ctx may be NULL: error of the last failed psim_create */

/* Chromosome table (Chromosome.java:65-88, TetraploidChromosome.java:80-89; read by
 * Main.readChromosomeFile, Main.java:581-634). pref_pairing / frac_quadrivalents may be NULL for
 * ploidy 2. */
int psim_set_chromosomes(psim_ctx* ctx, int32_t n_chrom, const double* length_morgan,
                         const double* centromere_morgan, const double* pref_pairing,
                         const double* frac_quadrivalents);

/* Pedigree in parents-first order (PopulationData.sortPedigree guarantees it,
 * PopulationData.java:309-430): parent0[i] = parent1[i] = -1 marks a founder, otherwise both are
 * indices < i. Every replicate shares this pedigree. */
int psim_set_pedigree(psim_ctx* ctx, int64_t n_ind, const int32_t* parent0, const int32_t* parent1);

/* Preset haplostructs for individuals [first, first+count) of replicate `replicate` — the replay
 * / extend modes (Main.readHaploStructFiles, Main.java:1166-1291). seg_off has
 * count*n_chrom*ploidy+1 entries; pos_morgan[seg_off[h]] is the start of homolog h's first segment
 * (the chromosome head, HaploStruct.java:38-48). */
int psim_set_haplostructs(psim_ctx* ctx, uint32_t replicate, int64_t first, int64_t count,
                          const int64_t* seg_off, const int32_t* founder, const double* pos_morgan);

/* Seam A. Gives founders without a haplostruct their founder alleles
 * (Individual.calcFounderHaplostruct, Individual.java:155-171: the k-th such founder gets
 * max_founder_allele+1+k*ploidy ...) and runs conception() (Individual.java:255-261) for every
 * other individual without one, one pedigree generation per kernel batch, in every replicate.
 * max_founder_allele is what readHaploStructFiles returned (-1 for a fresh simulation). */
int psim_simulate(psim_ctx* ctx, int32_t max_founder_allele);

/* psim_simulate without the final wait: the whole simulation (founders, then one kernel launch per generation) is
 * enqueued on the context's stream and the call returns. psim_wait - or the next call on the context that needs the
 * population - completes it and reports its status. Lets one host thread keep several contexts (replicate populations,
 * GPUs) busy: simulate population k+1 while population k is being emitted. */
int psim_simulate_async(psim_ctx* ctx, int32_t max_founder_allele);
int psim_wait(psim_ctx* ctx);

/* Same, but only the listed non-founders are simulated (their parents must already have a
 * haplostruct or be listed too). For drivers that shard the individuals of a generation over
 * several GPUs; founders without a haplostruct are initialised on every call. */
int psim_simulate_subset(psim_ctx* ctx, int32_t max_founder_allele, const int32_t* individuals, int64_t count);

/* Read back haplostructs of individuals [first, first+count) of one replicate
 * (what Main.writeHaploStructFiles serialises, Main.java:1100-1136). */
int psim_segment_count(psim_ctx* ctx, uint32_t replicate, int64_t first, int64_t count, int64_t* n_segments);
int psim_get_haplostructs(psim_ctx* ctx, uint32_t replicate, int64_t first, int64_t count,
                          int64_t* seg_off, int32_t* founder, double* pos_morgan);

/* Parental meiotic configurations (Individual.parconfig, Genotype.java:38-51; serialised by
 * Main.writeMeioticConfigsFile, Main.java:1138-1164). quad[count][n_chrom][2],
 * chromseq[count][n_chrom][2][ploidy] (0-based homolog numbers); -1 for individuals that were
 * not simulated in this context (founders, preset). */
int psim_get_meiotic_configs(psim_ctx* ctx, uint32_t replicate, int64_t first, int64_t count,
                             int32_t* quad, int32_t* chromseq);

/* Linkage map: locus_off[n_chrom+1], positions ascending within each chromosome exactly as
 * Chromosome.addLocus leaves them (Chromosome.java:133-162; read by Main.readMapFile,
 * Main.java:636-695). */
int psim_set_map(psim_ctx* ctx, const int64_t* locus_off, const double* pos_morgan);

/* Optional founder genotypes (Main.readFounderGenotypesFile, Main.java:783-891), as codes:
 * code[l*n_alleles + a] = rank of founder allele a's name among the locus' sorted unique allele
 * names (Locus.java:88-103), so the "MAX_ALLELE" of Main.java:407,991-992 is the largest code at
 * the locus. n_alleles = founder allele count. Pass code = NULL to clear. */
int psim_set_allele_codes(psim_ctx* ctx, int32_t n_alleles, const uint8_t* code);

/* Seam B. One fused pass writes every requested table for individuals
 * [ind_first, ind_first+ind_count) x all replicates of the context and loci
 * [locus_first, locus_first+locus_count) (global locus index = map order).
 * Tables are locus-major like the reference's files; a NULL pointer skips a table.
 * Column of (replicate r, individual i, homolog h) in founder / allele tables:
 *     ((r - replicate_first) * ind_count + (i - ind_first)) * ploidy + h
 * and in the dose table ((r - replicate_first) * ind_count + (i - ind_first)). */
#define PSIM_DOSE_MIN_ALLELE (-1) /* Main.MIN_ALLELE, Main.java:941 */
#define PSIM_DOSE_MAX_ALLELE (-2) /* Main.MAX_ALLELE, Main.java:942 */
#define PSIM_MEM_HOST 0
#define PSIM_MEM_DEVICE 1

typedef struct psim_emit_targets {
    void* founder;         /* u8 (founder_bytes = 1) or u16 (= 2) founder allele per cell:   */
    int64_t founder_pitch; /*   _founderalleles.dat, Main.writeGenotypesFile(founders=true)  */
    int32_t founder_bytes;
    int32_t memory;        /* PSIM_MEM_HOST or PSIM_MEM_DEVICE (all pointers of this call) */
    void* allele;          /* u8 allele code per cell: _genotypes.dat, needs allele codes */
    int64_t allele_pitch;
    void* dose;            /* u8 dose per individual: _alleledose.dat, Main.writeAlleledoseFile */
    int64_t dose_pitch;
    int32_t dose_which;    /* with codes: MAX/MIN_ALLELE or the name of founder allele `which`; */
    int32_t flags;         /*   without codes: founder allele `which` (Main.java:402-418,987-1015). flags: PSIM_EMIT_* */
} psim_emit_targets;

/* psim_emit_targets.flags (0 = u8 / u16 tables, the call returns when the tables are complete). */
#define PSIM_EMIT_ASYNC 0x1   /* device targets: return once the work is enqueued on the context's stream; the tables are
                                 complete after psim_wait (or any later call on the context that waits for the stream) */
#define PSIM_EMIT_PACKED4 0x2 /* founder and dose tables hold TWO 4-bit cells per byte (cell 2k in the low, 2k+1 in the high
                                 nibble; pitches in bytes; a row's last byte is zero-padded when the cell count is odd):
                                 needs <= 16 founder alleles, ploidy <= 14 and no allele table. Halves the bytes written
                                 to HBM and sent over PCIe for the same genotype calls */
/* Which emission kernels serve the call: PSIM_PATH_AUTO picks; the others are for tests and measurements (a request
 * the chosen kernels cannot serve falls through to the generic ones). */
#define PSIM_PATH_AUTO 0
#define PSIM_PATH_ROWS 1
#define PSIM_PATH_TILES 2
#define PSIM_PATH_GENERIC 3
#define PSIM_EMIT_PATH(p) ((p) << 8)

int psim_emit(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int64_t locus_first,
              int64_t locus_count, const psim_emit_targets* targets);

/* _allAlleledose.dat (Main.writeAllAlleledoseFile, Main.java:1037-1088): one row per
 * (locus, allele). Without codes the alleles are the founder alleles 0..A-1; with codes they are
 * the codes 0..max at each locus. row_off[locus_count+1] receives the first row of each locus. */
int psim_all_dose_rows(psim_ctx* ctx, int64_t locus_first, int64_t locus_count, int64_t* row_off);
int psim_emit_all_dose(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int64_t locus_first,
                       int64_t locus_count, void* out, int64_t pitch, int32_t memory);

/* Seam B as TEXT: the lines of _founderalleles.dat / _genotypes.dat / _alleledose.dat for loci
 * [locus_first, locus_first+locus_count), formatted on the device — what the reference produces
 * with one PrintWriter.print per cell (Main.writeGenotypesFile Main.java:903-939,
 * writeAlleledoseFile :969-1022). Every line is
 *     <locus name> { '\t' <token> } '\n'
 * with one token per cell in psim_emit's column order: the founder allele number, the allele NAME
 * of the cell's founder allele at the locus, or the dose. The host appends the bytes to its files
 * (header lines stay with the host). Locus names come from psim_set_locus_names (map order);
 * allele names from psim_set_allele_names: for every locus its sorted unique allele names
 * (Locus.getUniqueAlleleNames, Locus.java:88-103), name k belonging to code k of
 * psim_set_allele_codes. Offsets are into the concatenated characters (no terminators). */
int psim_set_locus_names(psim_ctx* ctx, const int64_t* name_off /* [L+1] */, const char* chars);
int psim_set_allele_names(psim_ctx* ctx, const int64_t* locus_first_name /* [L+1] */, const int64_t* name_off /* [names+1] */,
                          const char* chars);
typedef struct psim_text_targets {
    char* founder; int64_t founder_capacity; int64_t founder_bytes;   /* bytes = out: length of the text written */
    char* allele;  int64_t allele_capacity;  int64_t allele_bytes;
    char* dose;    int64_t dose_capacity;    int64_t dose_bytes;
    int32_t dose_which;    /* as in psim_emit_targets */
    int32_t reserved;
} psim_text_targets;
/* Host buffers (page-locked ones from psim_alloc_host are fastest); a NULL pointer skips a table.
 * Fails with PSIM_ECAPACITY (and the bytes needed in *_bytes) when a buffer is too small: an upper
 * bound per line is name + cells * (1 + longest token) + 1. */
int psim_emit_text(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int64_t locus_first, int64_t locus_count,
                   psim_text_targets* targets);

/* TEST-mode statistics (the counters of Main.testMeiosis, Main.java:1660-1721) as device
 * reductions over the GAMETES that made individuals [ind_first, ind_first+ind_count) of every
 * replicate: homologs [0, P/2) of an individual are the gamete of parent 0's meiosis, [P/2, P) that
 * of parent 1's (Gamete.fertilization, Gamete.java:95-116). One chromosome per call, Lc = its loci,
 * A = founder alleles of the pedigree. Host pointers; NULL skips an output. With an S1 population
 * of one founder this is the reference's TEST setting (one parent, POPSIZE meioses). With
 * unreduced_gametes != 0 an individual IS one 2n gamete: G = P homologs per gamete and one gamete
 * per individual (Main.java:1546-1548); otherwise G = P/2 and two gametes per individual. */
typedef struct psim_gamete_stats {
    int64_t n_gametes;            /* out: (P/G) * ind_count * replicate_count */
    uint32_t* allele_count;       /* [A][Lc]        gamete homologs with founder allele f at the locus (founderAlleleCountSel) */
    uint32_t* dose_count;         /* [A][Lc][G+1]   gametes by their dose of founder allele f (founderAlleleDoseSel) */
    uint32_t* homozygous_count;   /* [Lc]           gametes holding a founder allele twice or more (founderHomSel) */
    uint32_t* recomb_count;       /* [Lc][Lc]       entry [loc][loc2], loc2 < loc: gamete homologs whose founder alleles
                                                    at the two loci differ (recombCountSel); other entries are 0 */
    uint32_t* segments_hist;      /* [64]           gamete homologs by segment count - 1 (recombPoints; last bin: >= 63) */
    uint32_t* quadrivalent_hist;  /* [P/4+1]        meioses by number of quadrivalents on the chromosome */
} psim_gamete_stats;
int psim_gamete_stats_run(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int32_t chrom, psim_gamete_stats* out);

/* _allAlleledose.dat as text (Main.writeAllAlleledoseFile, Main.java:1037-1088): per locus one line
 * per allele,  <locus name> '\t' <allele> { '\t' <dose> } '\n',  the allele being the founder allele
 * number (no allele codes set) or the locus' k-th sorted unique allele name. Host buffer; *bytes
 * receives the length (the bytes needed when the call fails with PSIM_ECAPACITY). */
int psim_emit_all_dose_text(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int64_t locus_first, int64_t locus_count,
                            char* text, int64_t capacity, int64_t* bytes);

/* Drop every haplostruct and meiotic configuration (pedigree, chromosomes, map and allele codes
 * stay resident) and give the context a new RNG identity: the next psim_simulate is an
 * independent draw. Used by Monte Carlo drivers that re-simulate the same pedigree. */
int psim_reset(psim_ctx* ctx, int64_t seed, uint32_t replicate_first);

/* Run all device work of this context on the caller's CUDA stream (a cudaStream_t passed as a
 * pointer-sized integer; 0 restores the context's own stream). Lets a host that owns streams
 * order and time libpsim's work with its own events. */
int psim_set_stream(psim_ctx* ctx, uint64_t cuda_stream);

/* Page-locked host memory for emission targets and haplostruct buffers: a table emitted into
 * memory from psim_alloc_host leaves the GPU as one DMA at PCIe rate instead of being staged
 * through the driver's bounce buffer (what the reference's `PrintWriter` sink becomes on the
 * host side, Main.java:903-1088). Any other host pointer still works, only slower. */
int psim_alloc_host(psim_ctx* ctx, size_t bytes, void** out);
int psim_free_host(psim_ctx* ctx, void* p);

/* Introspection for benchmarks and sharded drivers. */
typedef struct psim_stats {
    int64_t n_meioses;          /* Individual.doMeiosis() equivalents run by psim_simulate (running total since create) */
    int64_t n_units;            /* (meiosis, chromosome) units (running total) */
    int64_t n_segments;         /* segments resident on the device */
    int64_t kernel_launches;    /* libpsim kernels launched since create */
    double  ms_simulate;        /* device time of the last psim_simulate (CUDA events) */
    double  ms_emit_kernel;     /* device time of the streaming emission kernel (k_emit_rows / k_emit_tiles / k_emit_generic)
                                   launches of the last psim_emit*, CUDA events on the launching stream */
    double  ms_emit_total;      /* device time of the last psim_emit*: the breakpoint -> locus-run index of a new population
                                   (k_index_runs), plan kernels, streaming kernels and copies */
    int64_t emit_kernel_launches; /* streaming-kernel launches of the last psim_emit* */
    double  ms_emit_plan;       /* device time of the tile-plan / dense helper kernels of the last psim_emit* */
} psim_stats;
int psim_get_stats(psim_ctx* ctx, psim_stats* out);

/* Device-resident haplostruct exchange for a pedigree sharded over several GPUs (one context per
 * rank; DESIGN.md "Multi-GPU"). export gives device pointers to a packed copy of the haplostructs
 * of [first, first+count) (valid until the next call on the context); import appends them on the
 * receiving context. The all-gather between the two is the caller's NCCL call. */
typedef struct psim_device_hs {
    int64_t n_homologs;   /* count * n_chrom * ploidy * replicate_count */
    int64_t n_segments;
    const int64_t* seg_off;  /* device, n_homologs+1 */
    const int32_t* founder;  /* device, n_segments */
    const double* pos;       /* device, n_segments */
} psim_device_hs;
int psim_export_haplostructs_device(psim_ctx* ctx, int64_t first, int64_t count, psim_device_hs* out);
int psim_import_haplostructs_device(psim_ctx* ctx, int64_t first, int64_t count, const psim_device_hs* in);

/* ---- One pedigree over the GPUs of a box, inside the library. The dependency is the reference's own
 * (Main.simulateIndividuals, Main.java:1504-1520: an individual needs its parents' haplostructs): every rank holds the whole
 * pedigree in its context (same psim_set_chromosomes / psim_set_pedigree / seed and the same individuals simulated so far
 * on all ranks, replicate_count 1), simulates its share of every generation, and after a generation the ranks hand each
 * other the new haplostructs that another rank needs: the sizes first (one 8-byte all-gather, the one host wait of the
 * exchange), then the haplostructs straight into the other ranks' segment slabs - packed and sent as one group of
 * ncclBroadcast (an all-gather of ragged parts) when few are shared, or, when every rank shares all it has just
 * simulated, straight out of its slab as three ncclAllGather of equal padded parts; no exchange at all in generations
 * where nobody shares. Results are bit-identical to psim_simulate on one GPU (counter-based streams keyed by individual).
 *   PSIM_PARTITION_BLOCK    every generation is cut into equal blocks and every new haplostruct is shared: any pedigree,
 *                           every rank ends with the whole population;
 *   PSIM_PARTITION_LINEAGE  an individual lives on the rank that owns its parents where possible; parents with children on
 *                           other ranks (or with a large progeny in one generation) are shared: a selfing chain or an F1
 *                           exchanges next to nothing, and every rank ends with its own lines (psim_shard_owner), which it
 *                           emits itself; psim_shard_gather brings listed individuals together when somebody needs them.
 * Transports: psim_shard_init - one process per GPU over NCCL (libnccl.so.2 is loaded at run time, by psim_shard_unique_id
 * / psim_shard_init only; rank 0 calls psim_shard_unique_id and hands the bytes to the other ranks by any means it has);
 * psim_shard_init_local - several contexts of ONE process (one per GPU, or several on one GPU), every context driven by its
 * own thread making the same calls, exchanging by peer copies. The communicator belongs to the context. */
#define PSIM_UNIQUE_ID_BYTES 128
#define PSIM_SHARD_MAX_WORLD 16
#define PSIM_PARTITION_BLOCK 0
#define PSIM_PARTITION_LINEAGE 1
int psim_shard_unique_id(uint8_t* id /* [PSIM_UNIQUE_ID_BYTES] */);
int psim_shard_init(psim_ctx* ctx, int32_t rank, int32_t world, const uint8_t* id);
typedef struct psim_shard_group psim_shard_group;
int psim_shard_group_create(int32_t world, psim_shard_group** out);
void psim_shard_group_destroy(psim_shard_group* group);   /* after the contexts that joined it */
int psim_shard_init_local(psim_ctx* ctx, int32_t rank, psim_shard_group* group);
int psim_simulate_sharded(psim_ctx* ctx, int32_t max_founder_allele, int32_t partition);
/* owner[i] = rank that holds individual i after psim_simulate_sharded, -1 = every rank (founders, shared individuals,
 * everybody with PSIM_PARTITION_BLOCK) */
int psim_shard_owner(psim_ctx* ctx, int32_t* owner /* [n_ind] */);
/* collective: the listed individuals go from their owners to every rank (the same list on all ranks) */
int psim_shard_gather(psim_ctx* ctx, const int32_t* individuals, int64_t count);
typedef struct psim_shard_stats {
    int64_t collectives;         /* exchanges steps of the last psim_simulate_sharded (sizes + payloads) */
    int64_t bytes_sent;          /* payload bytes this rank contributed */
    int64_t segments_received;   /* segments imported from other ranks */
    double  ms_total;            /* wall time of the call on this rank */
} psim_shard_stats;
int psim_shard_get_stats(psim_ctx* ctx, psim_shard_stats* out);
/* The partition on its own (pure host, no context, no device): generation, owner rank and shared mark of every individual
 * of a parents-first pedigree; has_haplostruct (may be NULL) marks individuals that need no simulation. */
int psim_shard_plan(int64_t n_ind, const int32_t* parent0, const int32_t* parent1, const uint8_t* has_haplostruct, int32_t world,
                    int32_t partition, int32_t* level /* [n_ind] or NULL */, int32_t* owner /* [n_ind] or NULL */,
                    uint8_t* share /* [n_ind] or NULL */);

#ifdef __cplusplus
}
#endif
#endif /* PSIM_H */
