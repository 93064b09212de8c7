// libpsim — B200 (sm_100a) meiosis + genotype emission behind the C ABI of include/psim.h.
//
// Device layout (DESIGN.md "Data layout in HBM"):
//   homolog id  hid = ((c * R + r) * N + i) * P + h      chromosome-major, so that the columns
//                                                         of one emission tile are contiguous
//   hom_desc[hid]  = begin (40 bit) | count << 40         into the segment slab
//   seg_pos[] f64 Morgan / seg_founder[] i32              HaploStruct.recombPos / founder, unmerged
//   run_desc / run_lb[] u32 / run_founder[] u16           K0 output: per homolog the runs of loci
//                                                         that carry one founder (locus-index space)
// Kernels (one header per group, all included below inside namespace psim):
//   psim_k1.cuh            k_init_founders, k_meiosis (K1: one launch per generation batch)
//   psim_k0.cuh            k_index_runs (K0)
//   psim_emit_rows.cuh     k_band_plan + k_emit_rows (K2, row image + TMA bulk stores: the roofline kernel)
//   psim_emit_tiles.cuh    k_emit_plan + k_emit_tiles + k_emit_dense (K2, tiled path)
//   psim_emit_generic.cuh  k_emit_generic, k_emit_all_dose
//   psim_text.cuh          k_text_rows (K3)      psim_stats.cuh  k_stats_* (K4)
//   psim_misc.cuh          haplostruct exchange helpers, k_match_codes
// This file: the host side of the C ABI.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <numeric>
#include <string>
#include <type_traits>
#include <vector>

#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/psim.h"
#include "psim_meiosis.cuh"

namespace psim {

constexpr int DESC_SHIFT = 40;
constexpr uint64_t DESC_MASK = (1ULL << DESC_SHIFT) - 1;
__host__ __device__ __forceinline__ uint64_t make_desc(uint64_t begin, uint64_t count) { return begin | (count << DESC_SHIFT); }
__host__ __device__ __forceinline__ uint64_t desc_begin(uint64_t d) { return d & DESC_MASK; }
__host__ __device__ __forceinline__ uint32_t desc_count(uint64_t d) { return (uint32_t)(d >> DESC_SHIFT); }

#include "psim_k1.cuh"
#include "psim_k0.cuh"
#include "psim_emit_tiles.cuh"
#include "psim_emit_rows.cuh"
#include "psim_emit_generic.cuh"
#include "psim_text.cuh"
#include "psim_stats.cuh"
#include "psim_misc.cuh"

}  // namespace psim

// ==========================================================================================
// host side
using namespace psim;

namespace {

std::string g_create_error;

struct ShardState;   // psim_shard.inl

struct psim_ctx_impl {
    psim_config cfg{};
    std::string err;
    int C = 0, P = 0, R = 1;
    int64_t N = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr, copy_stream = nullptr, aux_stream = nullptr;   // aux: lowest-priority stream of the meiosis kernels
    cudaEvent_t ev_aux0 = nullptr, ev_aux1 = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr;
    std::vector<cudaEvent_t> kev;   // event pool: triples (before plan, before stream, after stream) per launch
    int kev_used = 0;
    ModelParams model{};
    std::vector<ChromParams> chrom_h;
    ChromParams* chrom_d = nullptr;
    double max_len = 0.0;
    int chiasma_cap = 0;           // entries of a thread's column of the chiasma scratch
    std::vector<int32_t> par0_h, par1_h;
    int32_t *par0_d = nullptr, *par1_d = nullptr;
    std::vector<int32_t> hs_reps;  // per individual: number of replicates that have a haplostruct
    // store
    uint64_t* desc_d = nullptr;
    int64_t n_hom = 0;
    double* seg_pos_d = nullptr;
    int32_t* seg_founder_d = nullptr;
    int64_t slab_cap = 0, slab_used = 0;
    // runs
    uint64_t* run_desc_d = nullptr;
    uint32_t* run_lb_d = nullptr;
    uint16_t* run_founder_d = nullptr;
    int64_t run_cap = 0;
    bool runs_valid = false;
    // configs
    int8_t *cfg_quad_d = nullptr, *cfg_seq_d = nullptr;
    // map
    std::vector<int64_t> locus_off_h;
    int64_t* locus_off_d = nullptr;
    double* locus_pos_d = nullptr;
    int32_t* grid_d = nullptr; int64_t* grid_off_d = nullptr; double* grid_lo_d = nullptr; double* grid_inv_d = nullptr;   // MapGrid (k_index_runs)
    int64_t L = 0;
    // codes
    int A = 0;
    uint8_t* code_d = nullptr;
    std::vector<uint8_t> code_h;
    uint8_t* match_d = nullptr;
    uint32_t* code16_d = nullptr;   // [L][4]
    uint32_t* nibmask_d = nullptr;  // [L][2]
    uint4* mtab_d = nullptr;        // [L]
    // text emission (psim_set_locus_names / psim_set_allele_names)
    int64_t* label_off_d = nullptr; char* labels_d = nullptr; int64_t label_max = 0;
    int64_t* code_off_d = nullptr; int64_t* aname_off_d = nullptr; char* anames_d = nullptr; int64_t aname_max = 0;
    std::vector<int64_t> code_off_h;
    int match_which = 12345678;
    int founder_allele_count = 0;
    // scratch
    int* status_d = nullptr;                  // [2]: see k_slab_mark
    unsigned long long* cursor_d = nullptr;   // [0] segments of the slab in use (device copy while a simulation runs), [1] founders' base
    unsigned long long* mark_d = nullptr;     // slab cursor at the start of every batch (b_mark)
    int* tile_counter_d = nullptr;
    // Page-locked and mapped: [0] segment total of a batch, [1] error flag, [2] overflowed emission
    // units. Written by k_publish (a store over PCIe), not by a D2H memcpy: a small copy queues
    // behind whatever large transfer occupies the device-to-host copy engine.
    int64_t* scalars_h = nullptr;
    int64_t* scalars_dev = nullptr;   // the same memory as the device sees it
    void* cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;
    void* staging = nullptr;
    size_t staging_bytes = 0;
    // grow-only scratch of psim_simulate (no allocation in the steady state of a Monte Carlo loop)
    struct Buf { void* p = nullptr; size_t cap = 0; };
    Buf b_lev, b_xpos, b_xab, b_mark, b_find, b_fal;
    Buf b_plan, b_dense_list, b_text, b_text2, b_rowlen, b_rowoff;
    // map and allele-code arrays (psim_set_map / psim_set_allele_codes in a loop: no cudaFree / cudaMalloc per call - they
    // cost 2 - 17 ms of an end-to-end step; the *_d pointers above point into these, or are null when nothing is set)
    Buf b_locoff, b_locpos, b_grid, b_gridoff, b_gridlo, b_gridinv, b_code, b_match, b_code16, b_nibmask, b_mtab;
    // export buffers
    int64_t* exp_off_d = nullptr; int32_t* exp_founder_d = nullptr; double* exp_pos_d = nullptr;
    int64_t exp_off_cap = 0, exp_seg_cap = 0;
    int64_t pop_cap_ind = 0, pop_cap_hom = 0, pop_cap_cfg = 0;   // what the population arrays can hold
    // founder list and generation lists of a simulation that starts from the empty population
    struct SimPlan {
        bool valid = false, on_device = false;
        int32_t maxf_in = 0, max_level = 0;
        std::vector<int32_t> f_ind, f_allele, order;
        std::vector<int64_t> level_off;
    } plan, plan_once;   // plan: kept across psim_reset; plan_once: a simulation that starts from a partly filled population
    SimPlan *lev_plan = nullptr, *find_plan = nullptr;   // the plans whose generation / founder lists the device buffers hold
    // the simulation that has been enqueued (sim_begin) and not yet settled (sim_finish)
    struct SimRun {
        struct Batch { int level; int64_t first, count; };   // level 0: the founders; else `count` individuals of plan->order from `first`
        std::vector<Batch> batches;
        const SimPlan* plan = nullptr;
        int units_per_block = 0;
        bool pending = false;
    } run;
    bool emit_pending = false, emit_pending_device = false;   // an asynchronous psim_emit has not been waited for
    bool fresh = false;   // no individual has a haplostruct (after psim_set_pedigree / psim_reset)
    bool check_reps = false;            // psim_set_haplostructs was called: replicate coverage to be verified
    std::vector<int32_t> founder_idx;   // founders of the pedigree, ascending
    std::vector<int32_t> level_scratch; // N zeros between calls (simulate_impl)
    psim_stats stats{};
    int sm_count = 148;
    uint64_t ped_version = 0;           // counts psim_set_pedigree calls
    ShardState* shard = nullptr;        // psim_shard_init / psim_shard_init_local
};
void shard_free(psim_ctx_impl* c);

#define CTX ((psim_ctx_impl*)ctx)

int fail(psim_ctx_impl* c, int code, const std::string& msg) { c->err = msg; return code; }
int settle(psim_ctx_impl* c);

#define CUDA_TRY(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(c, PSIM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));  \
    } while (0)

template <typename T>
cudaError_t dev_alloc(T** p, size_t n) { return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }

// Tools.calcRecombDist (Tools.java:333-386): mean chiasma spacing such that 2L chiasmata remain
// on average when multivalents without any are rejected. Per-chromosome constant, host side.
double mdiff(double L, double m) { return std::fabs(m - 0.5 / (1 - std::exp(-L / m))); }
double calc_recomb_dist(double len) {
    if (len <= 0.500) return std::nan("");
    if (len > 5.0) return 0.5;
    const double toler = 0.00001;
    double lastdiff = mdiff(len, 0.5);
    double m1 = 10.5;
    while (mdiff(len, m1) < lastdiff) { lastdiff = mdiff(len, m1); m1 += 10.0; }
    double m2, m3;
    if (m1 == 10.5) { m2 = 5.5; m3 = 0.5; } else { m2 = m1 - 10.0; m3 = m1 - 20.0; }
    while (m1 - m3 > toler) {
        if (mdiff(len, m1 - 0.5 * (m1 - m2)) < mdiff(len, m2)) { m3 = m2; m2 = m1 - 0.5 * (m1 - m2); }
        else if (mdiff(len, m3 + 0.5 * (m2 - m3)) < mdiff(len, m2)) { m1 = m2; m2 = m3 + 0.5 * (m2 - m3); }
        else { m1 = m1 - 0.5 * (m1 - m2); m3 = m3 + 0.5 * (m2 - m3); }
    }
    return m2;
}

int ensure_slab(psim_ctx_impl* c, int64_t need) {
    if (need <= c->slab_cap) return 0;
    // grow generously: a Monte Carlo loop re-simulates populations whose segment totals vary by a
    // few percent, and a reallocation (cudaMalloc + copy + cudaFree) costs far more than the memory
    int64_t cap = std::max<int64_t>(need + need / 4, 2 * c->slab_cap);
    cap = std::max<int64_t>(cap, 1 << 16);
    double* np = nullptr; int32_t* nf = nullptr;
    CUDA_TRY(dev_alloc(&np, cap));
    CUDA_TRY(dev_alloc(&nf, cap));
    if (c->slab_used > 0) {
        CUDA_TRY(cudaMemcpyAsync(np, c->seg_pos_d, c->slab_used * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(nf, c->seg_founder_d, c->slab_used * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    cudaFree(c->seg_pos_d); cudaFree(c->seg_founder_d);
    c->seg_pos_d = np; c->seg_founder_d = nf; c->slab_cap = cap;
    return 0;
}
int ensure_buf(psim_ctx_impl* c, psim_ctx_impl::Buf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    CUDA_TRY(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return 0;
}
int ensure_cub(psim_ctx_impl* c, size_t bytes) {
    if (bytes <= c->cub_tmp_bytes) return 0;
    cudaFree(c->cub_tmp);
    c->cub_tmp = nullptr; c->cub_tmp_bytes = 0;
    CUDA_TRY(cudaMalloc(&c->cub_tmp, bytes));
    c->cub_tmp_bytes = bytes;
    return 0;
}
int exclusive_scan_i64(psim_ctx_impl* c, const int64_t* in, int64_t* out, int64_t n) {
    size_t bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, c->stream));
    if (int rc = ensure_cub(c, bytes)) return rc;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->cub_tmp, bytes, in, out, n, c->stream));
    c->stats.kernel_launches += 2;
    return 0;
}
inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

void free_population(psim_ctx_impl* c) {
    cudaFree(c->par0_d); cudaFree(c->par1_d); cudaFree(c->desc_d);
    cudaFree(c->run_desc_d); cudaFree(c->cfg_quad_d); cudaFree(c->cfg_seq_d);
    c->par0_d = c->par1_d = nullptr; c->desc_d = nullptr; c->run_desc_d = nullptr;
    c->cfg_quad_d = c->cfg_seq_d = nullptr;
    c->pop_cap_ind = c->pop_cap_hom = c->pop_cap_cfg = 0;
    c->slab_used = 0;
    c->runs_valid = false;
}

template <int P, int G, int MODE>
void launch_tiles_mode(psim_ctx_impl* c, const EmitArgs& ea, int blocks) {
    k_emit_tiles<P, G, MODE><<<blocks, EMIT_THREADS, 0, c->stream>>>(ea);
}
cudaEvent_t next_kev(psim_ctx_impl* c) {
    if (c->kev_used == (int)c->kev.size()) { cudaEvent_t e; cudaEventCreate(&e); c->kev.push_back(e); }
    return c->kev[c->kev_used++];
}
template <int P, int G>
void launch_tiles(psim_ctx_impl* c, const EmitArgs& ea, int blocks, int mode) {
    const int64_t n_cols = ea.n_cols_ind * P;
    k_emit_plan<<<dim3((unsigned)((n_cols + 255) / 256), (unsigned)ea.n_chunks), 256, 0, c->stream>>>(ea, G * P);
    cudaEventRecord(next_kev(c), c->stream);
    if (mode == 0) launch_tiles_mode<P, G, 0>(c, ea, blocks);
    else if (mode == 1) launch_tiles_mode<P, G, 1>(c, ea, blocks);
    else if (mode == 2) launch_tiles_mode<P, G, 2>(c, ea, blocks);
    else launch_tiles_mode<P, G, 3>(c, ea, blocks);
    cudaEventRecord(next_kev(c), c->stream);
    k_emit_dense<P, G><<<c->sm_count * 4, EMIT_THREADS, 0, c->stream>>>(ea);
}
// Row-image emission (k_band_plan + k_emit_rows) of loci [l_begin, l_end). Sets `done` when the request
// fits the path; otherwise leaves it false for the tiled / generic paths. Nothing is read back: units
// whose run-start list overflows are listed on the device and redone by k_emit_rows_overflow.
template <int DK>
static void launch_rows(psim_ctx_impl* c, const RowsArgs& ra, int blocks, size_t smem) {
    cudaFuncSetAttribute(k_emit_rows<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS_SMEM);
    k_emit_rows<DK><<<blocks, ROWS_THREADS, smem, c->stream>>>(ra);
}
template <int DK>
static void launch_rows_packed(psim_ctx_impl* c, const RowsArgs& ra, int blocks, size_t smem) {
    cudaFuncSetAttribute(k_emit_rows_packed<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS_SMEM);
    k_emit_rows_packed<DK><<<blocks, ROWS_THREADS, smem, c->stream>>>(ra);
}
#ifdef PSIM_TUNING
static int tuning_int(const char* name) { const char* v = getenv(name); return v ? atoi(v) : 0; }
#else
static int tuning_int(const char*) { return 0; }   // the knobs exist in tuning builds only (-DPSIM_TUNING)
#endif

static int emit_rows_path(psim_ctx_impl* c, int64_t ind_first, int64_t ind_count, int64_t l_begin, int64_t l_end, int64_t row0_locus,
                          uint8_t* founder, int64_t founder_pitch, uint8_t* dose, int64_t dose_pitch, int dose_mode, int dose_which, int mode,
                          int path, bool packed, bool& done) {
    done = false;
    const int P = c->P;
    if (packed && P > 8) return PSIM_OK;
    if (path == PSIM_PATH_TILES || path == PSIM_PATH_GENERIC || !(mode == 0 || (mode == 1 && (P == 2 || P == 4)))) return PSIM_OK;
    const int DK = !dose ? 0 : mode == 0 ? 1 : P == 4 ? 2 : 3;
    const int64_t n_cols_ind = (int64_t)c->R * ind_count, n_cols = n_cols_ind * P;
    static const int env_cap = tuning_int("PSIM_ROWS_CAP");
    static const int64_t env_seg = tuning_int("PSIM_SEG_MAX");
    static const int env_band = tuning_int("PSIM_BAND_ROWS");
    static const int env_units = tuning_int("PSIM_ROWS_UNITS");
    int cap = env_cap > 0 ? env_cap : 8192;
    const bool dose_rows = DK == 1 || (DK >= 2 && !ROWS_DOSE_DIRECT);
    const double per_cell = packed ? 0.5 * ROWS_NBUF + (DK >= 1 ? 0.5 * ROWS_NBUF : 0.0) / P
                                   : (double)ROWS_NBUF + (DK >= 2 ? 1.0 : 0.0) + (dose_rows ? (double)ROWS_NBUF : 0.0) / P;
    const int64_t smem_lists = ROWS_SMEM - 2 * (ROWS_MAX_BAND + 2) * 4;   // what the row buffers and the run-start list share
    const int64_t budget = smem_lists - cap * 5;
    const int64_t quantum = (packed ? 32 : 16) * P;
    int64_t seg_max = (int64_t)(budget / per_cell) / quantum * quantum;
    seg_max = std::min<int64_t>(seg_max, 65535 / quantum * quantum);
    if (env_seg > 0) seg_max = std::max<int64_t>(quantum, std::min<int64_t>(seg_max, env_seg / quantum * quantum));
    // run starts per cell and row, from the resident haplostructs (as in the tiled path)
    const double seg_per_hom = c->n_hom > 0 ? (double)c->slab_used / (double)c->n_hom : 1.0;
    const double loci_per_chrom = std::max(1.0, (double)c->L / c->C);
    const double starts_per_cell = std::max(0.0, seg_per_hom - 1.0) / loci_per_chrom;
    // A coarse map over many columns has many run starts per row: the row segment is then made
    // narrower, so that a band of at least 16 rows keeps its expected run starts within half the list
    // (a narrow segment still leaves the SM as one bulk store per row; below 4 KB the path is not taken).
    // PSIM_PATH_ROWS takes the path as it is: lists that overflow are redone per cell.
    // The list then takes the shared memory the narrower row buffers leave (5 bytes per run start): the segment is
    // the widest one whose band of `min_band` rows keeps its expected run starts within half of THAT list, because
    // what a short band costs is its start (state load, counting sort, buffer copies: ~5 us) against ~0.6 us per row.
    const int min_band = 48;
    bool coarse = false;
    if (path != PSIM_PATH_ROWS && env_cap <= 0 && starts_per_cell * (double)seg_max * min_band > cap / 2) {
        const int64_t fit = (int64_t)((double)smem_lists / (per_cell + 10.0 * starts_per_cell * min_band)) / quantum * quantum;
        if (fit < 4096) return PSIM_OK;   // sparse map: tiled / dense path
        seg_max = std::min(seg_max, fit);
        coarse = true;
    }
    const int n_seg = (int)((n_cols + seg_max - 1) / seg_max);
    const int64_t seg_cells = ((n_cols + n_seg - 1) / n_seg + quantum - 1) / quantum * quantum;
    if (coarse) cap = (int)std::min<int64_t>(1 << 20, std::max<int64_t>(cap, (smem_lists - (int64_t)(per_cell * (double)seg_cells) - 64) / 5));
    const int n_blocks = c->sm_count;
    // tallest band whose expected run starts stay within half the list capacity
    const double per_row = starts_per_cell * (double)seg_cells;
    const int64_t band_cap = env_band > 0 ? std::min(env_band, ROWS_MAX_BAND)
                           : std::max<int64_t>(8, std::min<int64_t>(ROWS_MAX_BAND, per_row > 0 ? (int64_t)((cap / 2) / per_row) : ROWS_MAX_BAND));

    if (n_cols >= (int64_t)1 << 31) return PSIM_OK;   // (k_band_plan does its column arithmetic in 32 bits)
    for (int k = 0; k < c->C; k++) {                  // (and keeps the band bounds of a chromosome piece in shared memory)
        const int64_t a0 = std::max(l_begin, c->locus_off_h[k]), a1 = std::min(l_end, c->locus_off_h[k + 1]);
        if (a1 > a0 && (a1 - a0 + band_cap - 1) / band_cap > ROWS_MAX_PLAN_BANDS / 2) return PSIM_OK;
    }
    RowsArgs ra{};
    ra.C = c->C; ra.R = c->R; ra.P = P; ra.N = c->N;
    ra.run_desc = c->run_desc_d; ra.run_lb = c->run_lb_d; ra.run_founder = c->run_founder_d;
    ra.ind_first = ind_first; ra.ind_count = ind_count; ra.n_cols_ind = n_cols_ind;
    ra.locus_first = row0_locus;
    ra.n_seg = n_seg; ra.seg_cells = (int)seg_cells; ra.cap = cap;
    ra.plan_stride = (16 + seg_cells + (int64_t)cap * 5 + 127) / 128 * 128;
    ra.founder = founder; ra.founder_pitch = founder_pitch;
    ra.dose = dose; ra.dose_pitch = dose_pitch;
    ra.dose_which = dose_which; ra.mtab = c->mtab_d;
    ra.unit_counter = c->tile_counter_d;
    {   // many run starts per unit: a global atomic each is what k_band_plan waits for, so they are counted per block first
        // (costs a second walk over the runs; tuning builds: PSIM_PLAN_SHARED = 1 always, -1 never)
        static const int env_shared = tuning_int("PSIM_PLAN_SHARED");
        ra.plan_shared = env_shared ? env_shared > 0 : starts_per_cell * 256.0 * (double)std::min<int64_t>(band_cap, 64) >= 24.0;
    }
    const size_t smem = (packed ? ROWS_NBUF * (size_t)seg_cells / 2 + (size_t)(DK >= 1 ? ROWS_NBUF : 0) * (seg_cells / P / 2)
                                : ROWS_NBUF * (size_t)seg_cells + (DK >= 2 ? seg_cells : 0) + (size_t)(dose_rows ? ROWS_NBUF : 0) * (seg_cells / P)) +
                        (size_t)cap * 5 + 2 * (ROWS_MAX_BAND + 2) * 4;
    RowsOverflowArgs oa{};
    oa.dose_mode = dose_mode; oa.code = c->code_d; oa.match_code = c->match_d; oa.A = c->A;
    int k0 = 0;
    while (k0 < c->C && c->locus_off_h[k0 + 1] <= l_begin) k0++;
    while (k0 < c->C && c->locus_off_h[k0] < l_end) {
        ra.n_chunks = 0;
        ra.chunk_unit_base[0] = 0;
        while (k0 < c->C && c->locus_off_h[k0] < l_end && ra.n_chunks < 64) {
            const int64_t a0 = std::max(l_begin, c->locus_off_h[k0]), a1 = std::min(l_end, c->locus_off_h[k0 + 1]);
            if (a1 > a0) {
                const int q = ra.n_chunks++;
                ra.chunk_chrom[q] = k0; ra.chunk_l0[q] = a0; ra.chunk_l1[q] = a1; ra.chunk_coff[q] = c->locus_off_h[k0];
            }
            k0++;
        }
        if (ra.n_chunks == 0) break;
        {   // Bands per chromosome piece: proportional to its rows, largest remainders get the extra
            // band. The total is the smallest one (tallest bands) whose units fill whole waves of
            // resident blocks; no band is taller than band_cap. (Handing the last rows out in many short
            // bands so that fast blocks take more was tried - tools/exp_tail.sh - and does not pay: the
            // blocks of a launch end 163-216 us after its start, but their sum is what the memory system
            // takes; a block that ends early only leaves its share to the others.)
            int64_t total = 0;
            for (int q = 0; q < ra.n_chunks; q++) total += ra.chunk_l1[q] - ra.chunk_l0[q];
            auto allocate = [&](int64_t target) -> int64_t {
                int64_t given = 0;
                double rem[64];
                for (int q = 0; q < ra.n_chunks; q++) {
                    const int64_t len = ra.chunk_l1[q] - ra.chunk_l0[q];
                    const double ideal = (double)len * (double)target / (double)total;
                    int64_t nb = std::max<int64_t>(1, (int64_t)ideal);
                    nb = std::max(nb, (len + band_cap - 1) / band_cap);
                    nb = std::min(nb, std::min<int64_t>(len, ROWS_MAX_PLAN_BANDS));
                    ra.chunk_bands[q] = (int)nb;
                    rem[q] = ideal - (double)nb;
                    given += nb;
                }
                while (given < target) {
                    int best = -1;
                    for (int q = 0; q < ra.n_chunks; q++)
                        if (ra.chunk_bands[q] < std::min<int64_t>(ra.chunk_l1[q] - ra.chunk_l0[q], ROWS_MAX_PLAN_BANDS) && (best < 0 || rem[q] > rem[best])) best = q;
                    if (best < 0) break;
                    ra.chunk_bands[best]++;
                    rem[best] -= 1.0;
                    given++;
                }
                return given;
            };
            const int64_t t_min = std::max<int64_t>(1, (total + band_cap - 1) / band_cap);
            int64_t best_t = t_min;
            double best_eff = 0.0;
            // (a band's start - state load, counting sort, buffer copies, while the previous band's last rows drain - costs
            // about as much as eight of its rows: a fuller last wave does not pay for bands that are much shorter. The
            // candidates are the fewest bands and the band counts next to the first few whole numbers of waves.)
            const double start_rows = 8.0;
            auto consider = [&](int64_t tg) {
                if (tg < t_min || tg > std::max<int64_t>(t_min, total / (env_band > 0 ? 1 : 16))) return;
                const int64_t bands = allocate(tg), units = bands * n_seg;
                const double rows_per_band = (double)total / (double)bands;
                const double eff = (double)units / (double)((units + n_blocks - 1) / n_blocks * n_blocks) * rows_per_band / (rows_per_band + start_rows);
                if (eff > best_eff + 1e-9) { best_eff = eff; best_t = tg; }
            };
            consider(t_min);
            if (env_band <= 0) {
                const int64_t w0 = (t_min * n_seg + n_blocks - 1) / n_blocks;
                for (int64_t wv = w0; wv < w0 + 4; wv++)
                    for (int64_t tg = wv * n_blocks / n_seg - 1; tg <= wv * n_blocks / n_seg + 1; tg++) consider(tg);
            }
            if (env_units > 0) best_t = std::max<int64_t>(t_min, env_units);
            allocate(best_t);
            for (int q = 0; q < ra.n_chunks; q++) ra.chunk_unit_base[q + 1] = ra.chunk_unit_base[q] + ra.chunk_bands[q] * n_seg;
        }
        ra.n_units = ra.chunk_unit_base[ra.n_chunks];
        {
            static const int env_trace = tuning_int("PSIM_ROWS_TRACE");
            if (env_trace) fprintf(stderr, "[rows] cols %lld x %d segments of %lld cells, list %d, starts/cell/row %.5f, tallest band %lld rows, %d pieces, %d units (first piece: %lld rows in %d bands)\n",
                                   (long long)n_cols, n_seg, (long long)seg_cells, cap, starts_per_cell, (long long)band_cap, ra.n_chunks, ra.n_units,
                                   (long long)(ra.chunk_l1[0] - ra.chunk_l0[0]), ra.chunk_bands[0]);
        }
        if (int rc = ensure_buf(c, c->b_plan, (size_t)ra.n_units * ra.plan_stride)) return rc;
        if (int rc = ensure_buf(c, c->b_dense_list, ((size_t)ra.n_units + 1) * sizeof(int))) return rc;
        ra.plan = (uint8_t*)c->b_plan.p;
        ra.overflow = (int*)c->b_dense_list.p;
        // list counters, overflow count and ticket: one small kernel (a 2-D memset is a copy-engine job
        // and queues behind the PCIe copy of the previous chunk)
        k_rows_reset<<<grid_for(ra.n_units + 2, 256), 256, 0, c->stream>>>(ra.plan, ra.plan_stride, ra.n_units, ra.overflow, c->tile_counter_d);
        k_band_plan<<<dim3((unsigned)((n_cols + 255) / 256), (unsigned)ra.n_chunks), 256, 0, c->stream>>>(ra);
        cudaEventRecord(next_kev(c), c->stream);
        const int blocks = std::min(ra.n_units, n_blocks);
        if (packed) {
            if (DK == 0) launch_rows_packed<0>(c, ra, blocks, smem);
            else if (DK == 1) launch_rows_packed<1>(c, ra, blocks, smem);
            else if (DK == 2) launch_rows_packed<2>(c, ra, blocks, smem);
            else launch_rows_packed<3>(c, ra, blocks, smem);
        } else {
            if (DK == 0) launch_rows<0>(c, ra, blocks, smem);
            else if (DK == 1) launch_rows<1>(c, ra, blocks, smem);
            else if (DK == 2) launch_rows<2>(c, ra, blocks, smem);
            else launch_rows<3>(c, ra, blocks, smem);
        }
        cudaEventRecord(next_kev(c), c->stream);
        if (packed) k_emit_rows_overflow_packed<<<dim3((unsigned)c->sm_count, 8), 256, 0, c->stream>>>(ra, oa, c->locus_off_d);
        else k_emit_rows_overflow<<<dim3((unsigned)c->sm_count, 8), 256, 0, c->stream>>>(ra, oa);
        c->stats.kernel_launches += 4;
        c->stats.emit_kernel_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    done = true;
    return PSIM_OK;
}

}  // namespace

extern "C" {

int psim_abi_version(void) { return PSIM_ABI_VERSION; }

const char* psim_last_error(const psim_ctx* ctx) {
    if (!ctx) return g_create_error.c_str();
    return ((const psim_ctx_impl*)ctx)->err.c_str();
}

int psim_create(const psim_config* cfg, psim_ctx** out) {
    if (!cfg || !out) { g_create_error = "psim_create: null argument"; return PSIM_EINVAL; }
    *out = nullptr;
    if (cfg->ploidy <= 0 || cfg->ploidy % 2 != 0) { g_create_error = "Error in PopulationData: ploidy must be even"; return PSIM_EINVAL; }
    if (cfg->ploidy > MAXP) { g_create_error = "ploidy above 16 is not supported by the device path"; return PSIM_EINVAL; }
    if (cfg->replicate_count < 1) { g_create_error = "replicate_count must be >= 1"; return PSIM_EINVAL; }
    if (cfg->unreduced_gametes < 0 || cfg->unreduced_gametes > 2) { g_create_error = "unreduced_gametes must be 0, 1 (FDR) or 2 (SDR)"; return PSIM_EINVAL; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available (libpsim has no CPU fallback): ") + cudaGetErrorString(e);
        return PSIM_ECUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "invalid CUDA device ordinal"; return PSIM_EINVAL; }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return PSIM_ECUDA; }
    psim_ctx_impl* c = new psim_ctx_impl();
    c->cfg = *cfg;
    c->P = cfg->ploidy;
    c->R = (int)cfg->replicate_count;
    c->model.ploidy = cfg->ploidy;
    c->model.chiasma_interference = cfg->chiasma_interference != 0;
    c->model.natural_pairing = cfg->natural_pairing != 0;
    c->model.allow_no_recomb = cfg->allow_no_recomb != 0;
    c->model.parallel_multivalents = cfg->parallel_multivalents;
    c->model.paired_centromeres = cfg->paired_centromeres;
    c->model.seed = (uint64_t)cfg->seed;
    c->model.unreduced = cfg->unreduced_gametes;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              [&] { int least = 0, greatest = 0; cudaDeviceGetStreamPriorityRange(&least, &greatest);
                    return cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, least); }() == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_aux0, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_aux1, cudaEventDisableTiming) == cudaSuccess &&
              cudaHostAlloc((void**)&c->scalars_h, 8 * sizeof(int64_t), cudaHostAllocMapped) == cudaSuccess &&
              cudaHostGetDevicePointer((void**)&c->scalars_dev, c->scalars_h, 0) == cudaSuccess &&
              cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess &&
              cudaEventCreate(&c->ev2) == cudaSuccess && cudaEventCreate(&c->ev3) == cudaSuccess && cudaEventCreate(&c->ev4) == cudaSuccess &&
              cudaMalloc((void**)&c->status_d, 2 * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void**)&c->cursor_d, 2 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc((void**)&c->tile_counter_d, sizeof(int)) == cudaSuccess;
    if (!ok) { g_create_error = std::string("CUDA setup failed: ") + cudaGetErrorString(cudaGetLastError()); delete c; return PSIM_ECUDA; }
    c->stream = c->own_stream;
    cudaMemset(c->status_d, 0, 2 * sizeof(int));
    cudaMemset(c->cursor_d, 0, 2 * sizeof(unsigned long long));
    *out = (psim_ctx*)c;
    return PSIM_OK;
}

void psim_destroy(psim_ctx* ctx) {
    if (!ctx) return;
    psim_ctx_impl* c = CTX;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    shard_free(c);
    free_population(c);
    cudaFree(c->chrom_d); cudaFree(c->seg_pos_d); cudaFree(c->seg_founder_d);
    cudaFree(c->run_lb_d); cudaFree(c->run_founder_d);
    cudaFree(c->status_d); cudaFree(c->cursor_d); cudaFree(c->tile_counter_d); cudaFree(c->cub_tmp); cudaFree(c->staging);
    cudaFree(c->exp_off_d); cudaFree(c->exp_founder_d); cudaFree(c->exp_pos_d);
    for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
    for (psim_ctx_impl::Buf* b : {&c->b_lev, &c->b_xpos, &c->b_xab, &c->b_mark, &c->b_find, &c->b_fal,
                                 &c->b_plan, &c->b_dense_list, &c->b_text, &c->b_text2, &c->b_rowlen, &c->b_rowoff,
                                 &c->b_locoff, &c->b_locpos, &c->b_grid, &c->b_gridoff, &c->b_gridlo, &c->b_gridinv,
                                 &c->b_code, &c->b_match, &c->b_code16, &c->b_nibmask, &c->b_mtab}) cudaFree(b->p);
    cudaFree(c->label_off_d); cudaFree(c->labels_d); cudaFree(c->code_off_d); cudaFree(c->aname_off_d); cudaFree(c->anames_d);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    if (c->ev4) cudaEventDestroy(c->ev4);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->ev_aux0) cudaEventDestroy(c->ev_aux0);
    if (c->ev_aux1) cudaEventDestroy(c->ev_aux1);
    if (c->scalars_h) cudaFreeHost(c->scalars_h);
    delete c;
}

int psim_set_chromosomes(psim_ctx* ctx, int32_t n_chrom, const double* length, const double* centromere,
                         const double* pref_pairing, const double* frac_quad) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (n_chrom <= 0 || !length || !centromere) return fail(c, PSIM_EINVAL, "CHROMFILE: no chromosomes");
    if (n_chrom >= (1 << 20)) return fail(c, PSIM_EINVAL, "too many chromosomes");
    if (c->N > 0) return fail(c, PSIM_EINVAL, "psim_set_chromosomes must precede psim_set_pedigree");
    cudaSetDevice(c->cfg.device);
    c->C = n_chrom;
    c->chrom_h.assign(n_chrom, ChromParams());
    c->max_len = 0.0;
    for (int k = 0; k < n_chrom; k++) {
        ChromParams& ch = c->chrom_h[k];
        // Chromosome.java:65-80: head 0, tail = length, both widened to include the centromere
        ch.centro = centromere[k];
        ch.head = centromere[k] < 0.0 ? centromere[k] : 0.0;
        ch.tail = centromere[k] > length[k] ? centromere[k] : length[k];
        ch.pref_pairing = pref_pairing ? pref_pairing[k] : 0.0;
        const double len = ch.tail - ch.head;
        if (!(len > 0.0)) return fail(c, PSIM_EINVAL, "chromosome length must be positive");
        if (!c->model.allow_no_recomb && len < 0.70)
            return fail(c, PSIM_EINVAL, "if ALLOWNORECOMBINATION is false the minimum chromosome length is 70.0 cM");
        ch.recomb_dist_noempty = calc_recomb_dist(len);
        for (int q = 0; q < MAXP / 4; q++) ch.cumul_binomial[q] = 1.0;
        if (c->P > 2) {
            // JSci BinomialDistribution.cumulative (TetraploidChromosome.java:102-108) restated
            const int n = c->P / 4;
            const double fq = frac_quad ? frac_quad[k] : 0.0;
            for (int q = 0; q < n; q++) {
                double s = 0.0;
                for (int kk = 0; kk <= q; kk++)
                    s += std::exp(std::lgamma(n + 1.0) - std::lgamma(kk + 1.0) - std::lgamma(n - kk + 1.0)) *
                         std::pow(fq, kk) * std::pow(1.0 - fq, n - kk);
                ch.cumul_binomial[q] = s;
            }
        }
        c->max_len = std::max(c->max_len, len);
    }
    // capacity of a multivalent's chiasma list: chiasmata on a 4-strand bundle are Poisson(2 len)
    // (4 len for a parallel quadrivalent); a list that overflows all the same is reported as PSIM_ECAPACITY
    const double mean_e = 4.0 * c->max_len;
    c->chiasma_cap = (int)std::ceil(mean_e + 12.0 * std::sqrt(mean_e) + 16.0);
    cudaFree(c->chrom_d); c->chrom_d = nullptr;
    CUDA_TRY(dev_alloc(&c->chrom_d, (size_t)n_chrom));
    CUDA_TRY(cudaMemcpy(c->chrom_d, c->chrom_h.data(), n_chrom * sizeof(ChromParams), cudaMemcpyHostToDevice));
    return PSIM_OK;
}

int psim_set_pedigree(psim_ctx* ctx, int64_t n_ind, const int32_t* parent0, const int32_t* parent1) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (n_ind <= 0 || !parent0 || !parent1) return fail(c, PSIM_EINVAL, "empty pedigree");
    if (c->C == 0) return fail(c, PSIM_ESTATE, "psim_set_chromosomes must be called first");
    if (n_ind > INT32_MAX) return fail(c, PSIM_EINVAL, "more than 2^31-1 individuals");
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    c->ped_version++;
    int nf = 0;
    for (int64_t i = 0; i < n_ind; i++) {
        const bool f0 = parent0[i] < 0, f1 = parent1[i] < 0;
        if (f0 != f1) return fail(c, PSIM_EINVAL, "Individual " + std::to_string(i) + " has one known and one unknown parent, which is not allowed");
        if (f0) { nf++; continue; }
        if (parent0[i] >= i || parent1[i] >= i) return fail(c, PSIM_EINVAL, "pedigree is not in parents-first order");
    }
    // A Monte Carlo loop sets a pedigree of the same size again and again: the population arrays are
    // grow-only (cudaFree / cudaMalloc cost 1-10 ms, far more than the simulation of such a pedigree).
    const int64_t n_hom = (int64_t)c->C * c->R * n_ind * c->P;
    const int64_t ncfg = (int64_t)c->R * n_ind * c->C * 2;
    if (n_ind > c->pop_cap_ind || n_hom > c->pop_cap_hom || ncfg > c->pop_cap_cfg) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        free_population(c);
        CUDA_TRY(dev_alloc(&c->par0_d, (size_t)n_ind));
        CUDA_TRY(dev_alloc(&c->par1_d, (size_t)n_ind));
        CUDA_TRY(dev_alloc(&c->desc_d, (size_t)n_hom));
        CUDA_TRY(dev_alloc(&c->run_desc_d, (size_t)n_hom));
        CUDA_TRY(dev_alloc(&c->cfg_quad_d, (size_t)ncfg));
        CUDA_TRY(dev_alloc(&c->cfg_seq_d, (size_t)ncfg * c->P));
        c->pop_cap_ind = n_ind; c->pop_cap_hom = n_hom; c->pop_cap_cfg = ncfg;
    }
    c->slab_used = 0;
    c->runs_valid = false;
    c->fresh = true;
    c->plan.valid = false;
    c->check_reps = false;
    c->level_scratch.clear();
    c->founder_idx.clear();
    for (int64_t i = 0; i < n_ind; i++) if (parent0[i] < 0) c->founder_idx.push_back((int32_t)i);
    c->N = n_ind;
    c->founder_allele_count = nf * c->P;
    c->par0_h.assign(parent0, parent0 + n_ind);
    c->par1_h.assign(parent1, parent1 + n_ind);
    c->hs_reps.assign(n_ind, 0);
    c->n_hom = n_hom;
    // stream-ordered with the kernels that follow; the sources are the library's own copies (the
    // caller's arrays may be page-locked, and then an asynchronous copy would outlive this call)
    CUDA_TRY(cudaMemcpyAsync(c->par0_d, c->par0_h.data(), n_ind * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->par1_d, c->par1_h.data(), n_ind * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->desc_d, 0, n_hom * sizeof(uint64_t), c->stream));
    CUDA_TRY(cudaMemsetAsync(c->cfg_quad_d, 0xFF, ncfg, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->cfg_seq_d, 0xFF, ncfg * c->P, c->stream));
    c->stats.n_segments = 0;
    return PSIM_OK;
}

int psim_set_haplostructs(psim_ctx* ctx, uint32_t replicate, int64_t first, int64_t count,
                          const int64_t* seg_off, const int32_t* founder, const double* pos) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (!seg_off || !founder || !pos) return fail(c, PSIM_EINVAL, "null argument");
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    if (first < 0 || count <= 0 || first + count > c->N) return fail(c, PSIM_EINVAL, "individual range out of bounds");
    if (replicate < c->cfg.replicate_first || replicate >= c->cfg.replicate_first + c->cfg.replicate_count)
        return fail(c, PSIM_EINVAL, "replicate not held by this context");
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    const int r = (int)(replicate - c->cfg.replicate_first);
    const int64_t nh = count * c->C * c->P;
    const int64_t nseg = seg_off[nh];
    for (int64_t t = 0; t < nh; t++) {
        if (seg_off[t + 1] <= seg_off[t]) return fail(c, PSIM_EINVAL, "every homolog needs at least one segment");
        if (seg_off[t + 1] - seg_off[t] >= (1 << 24)) return fail(c, PSIM_ECAPACITY, "more than 2^24 segments in one homolog");
    }
    for (int64_t s = 0; s < nseg; s++)   // founder alleles are numbered 0 .. founderAlleleCount-1 (Main.java:1235)
        if (founder[s] < 0 || founder[s] >= c->founder_allele_count || founder[s] > 65535)
            return fail(c, PSIM_EINVAL, "invalid founder allele '" + std::to_string(founder[s]) + "'");
    if (int rc = ensure_slab(c, c->slab_used + nseg)) return rc;
    int64_t* off_d = nullptr;
    CUDA_TRY(dev_alloc(&off_d, (size_t)nh + 1));
    CUDA_TRY(cudaMemcpyAsync(off_d, seg_off, (nh + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->seg_pos_d + c->slab_used, pos, nseg * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->seg_founder_d + c->slab_used, founder, nseg * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    k_scatter_desc<<<grid_for(nh, 256), 256, 0, c->stream>>>(nh, first, r, c->C, c->R, c->N, c->P, off_d, c->slab_used, c->desc_d);
    c->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(off_d);
    c->slab_used += nseg;
    c->stats.n_segments = c->slab_used;
    for (int64_t i = first; i < first + count; i++) c->hs_reps[i] = std::min(c->hs_reps[i] + 1, c->R);
    c->runs_valid = false;
    c->fresh = false;
    if (c->R > 1) c->check_reps = true;
    return PSIM_OK;
}

// ---- psim_simulate: the whole simulation is enqueued (founders, then one k_meiosis launch per
// generation batch) without the host reading anything back; sim_finish waits once, and only if the
// slab turned out too small grows it and runs the batches from the first incomplete one again.
static int sim_enqueue(psim_ctx_impl* c, int from_batch, int64_t cursor_at) {
    psim_ctx_impl::SimRun& run = c->run;
    const psim_ctx_impl::SimPlan& pl = *run.plan;
    const int C = c->C, P = c->P, R = c->R;
    const int64_t N = c->N;
    int32_t* const lev_all_d = (int32_t*)c->b_lev.p;
    // The meiosis kernels run on a stream of the lowest priority (between two events of the context's stream): when
    // another context of the process emits at the same time, the blocks of ITS streaming kernel - one per SM, bound by
    // HBM - are placed first, and the meiosis blocks (bound by latency, next to no memory traffic) fill what is left.
    static const int env_prio = tuning_int("PSIM_SIM_PRIO");
    cudaStream_t const ctx_stream = c->stream;
    const bool low = env_prio >= 0 && c->aux_stream;
    if (low) {
        CUDA_TRY(cudaEventRecord(c->ev_aux0, ctx_stream));
        CUDA_TRY(cudaStreamWaitEvent(c->aux_stream, c->ev_aux0, 0));
    }
    struct Swap {   // (the launches below name c->stream)
        psim_ctx_impl* c; cudaStream_t keep;
        ~Swap() { c->stream = keep; }
    } swap{c, ctx_stream};
    if (low) c->stream = c->aux_stream;
    for (int b = from_batch; b < (int)run.batches.size(); b++) {
        const psim_ctx_impl::SimRun::Batch& bt = run.batches[b];
        const long long restart = b == from_batch ? (long long)cursor_at : -1;
        if (bt.level == 0) {   // founders (Main.java:1510-1513)
            const int n_new = (int)pl.f_ind.size();
            const int64_t nseg = (int64_t)n_new * R * C * P;
            k_slab_mark<<<1, 1, 0, c->stream>>>(c->cursor_d, c->mark_d, b, nseg, c->slab_cap, c->cursor_d + 1, c->status_d, restart);
            k_init_founders<<<grid_for(nseg, 256), 256, 0, c->stream>>>(n_new * R, n_new, (const int32_t*)c->b_find.p, (const int32_t*)c->b_fal.p,
                                                                        C, R, N, P, c->chrom_d, c->desc_d, c->seg_pos_d, c->seg_founder_d,
                                                                        c->cursor_d + 1, c->slab_cap);
            c->stats.kernel_launches += 2;
            continue;
        }
        k_slab_mark<<<1, 1, 0, c->stream>>>(c->cursor_d, c->mark_d, b, 0, c->slab_cap, c->cursor_d + 1, c->status_d, restart);
        MeiosisArgs ma;
        ma.model = c->model; ma.chrom = c->chrom_d; ma.C = C; ma.R = R; ma.P = P; ma.N = N;
        ma.replicate_first = c->cfg.replicate_first;
        ma.n_lev = (int)bt.count; ma.lev_ind = lev_all_d + bt.first; ma.par0 = c->par0_d; ma.par1 = c->par1_d;
        ma.desc = c->desc_d; ma.seg_pos = c->seg_pos_d; ma.seg_founder = c->seg_founder_d;
        ma.slab_cap = c->slab_cap; ma.cursor = c->cursor_d; ma.status = c->status_d; ma.batch = b;
        ma.cfg_quad = c->cfg_quad_d; ma.cfg_seq = c->cfg_seq_d;
        ma.x_cap = c->chiasma_cap;
        ma.units_per_block = run.units_per_block;
        const int64_t n_units = (int64_t)bt.count * C * R * 2;
        const int64_t blocks = (n_units + run.units_per_block - 1) / run.units_per_block;
        ma.x_pos = (double*)c->b_xpos.p; ma.x_ab = (uint8_t*)c->b_xab.p;
        if (c->model.chiasma_interference) {
            if (c->model.allow_no_recomb) k_meiosis<true, true><<<(unsigned)blocks, MEI_THREADS, 0, c->stream>>>(ma);
            else k_meiosis<true, false><<<(unsigned)blocks, MEI_THREADS, 0, c->stream>>>(ma);
        } else {
            if (c->model.allow_no_recomb) k_meiosis<false, true><<<(unsigned)blocks, MEI_THREADS, 0, c->stream>>>(ma);
            else k_meiosis<false, false><<<(unsigned)blocks, MEI_THREADS, 0, c->stream>>>(ma);
        }
        c->stats.kernel_launches += 2;
    }
    CUDA_TRY(cudaGetLastError());
    if (low) {
        CUDA_TRY(cudaEventRecord(c->ev_aux1, c->aux_stream));
        c->stream = ctx_stream;
        CUDA_TRY(cudaStreamWaitEvent(ctx_stream, c->ev_aux1, 0));
    }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    k_publish<<<1, 32, 0, c->stream>>>(c->scalars_dev, (const int64_t*)c->cursor_d, c->status_d, c->status_d + 1);
    c->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    return PSIM_OK;
}

// subset == nullptr: every individual without a haplostruct; otherwise only the listed non-founders
// (founders without a haplostruct are always initialised: it is cheap and every rank of a sharded
// pedigree needs them).
static int sim_start(psim_ctx_impl* c, psim_ctx_impl::SimPlan& plan, bool upload, int lv_lo, int lv_hi, bool founders);
static int sim_begin(psim_ctx_impl* c, int32_t max_founder_allele, const int32_t* subset, int64_t n_subset) {
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    cudaSetDevice(c->cfg.device);
    const int P = c->P, R = c->R;
    const int64_t N = c->N;
    // A Monte Carlo loop (psim_reset, psim_simulate, ...) simulates the same pedigree from the same
    // empty state again and again: the founder list and the generation lists are then kept, on the
    // host and on the device, instead of being rebuilt and uploaded every time.
    psim_ctx_impl::SimPlan& sp = c->plan;
    const bool fresh = c->fresh && !subset;
    const bool reuse = fresh && sp.valid && sp.maxf_in == max_founder_allele;
    if (c->check_reps) {   // only after psim_set_haplostructs with several replicates
        for (int64_t i = 0; i < N; i++)
            if (c->hs_reps[i] != 0 && c->hs_reps[i] != R)
                return fail(c, PSIM_EINVAL, "preset haplostructs must cover the same individuals in every replicate");
        c->check_reps = false;
    }
    psim_ctx_impl::SimPlan& pl = reuse ? sp : c->plan_once;
    if (!reuse) {
        pl = psim_ctx_impl::SimPlan();
        // --- founders (Main.java:1510-1513)
        int32_t maxf = max_founder_allele;
        for (int32_t i : c->founder_idx) {
            if (c->hs_reps[i]) continue;
            pl.f_ind.push_back(i); pl.f_allele.push_back(maxf + 1); maxf += P;
        }
        // founder alleles are numbered 0 .. founder_allele_count-1 (Main.java:1235 rejects others): the count
        // decides the width of the founder tables and sizes the statistics histograms
        if (!pl.f_ind.empty() && maxf >= c->founder_allele_count)
            return fail(c, PSIM_EINVAL, "invalid founder allele '" + std::to_string(maxf) + "': the pedigree has " +
                                            std::to_string(c->founder_allele_count) + " founder alleles");
        if (maxf > 65535) return fail(c, PSIM_ECAPACITY, "more than 65536 founder alleles");
        // --- generations: level = 1 + max(level of parents) over individuals still to simulate.
        // `level` is a context-owned array of N zeros; only the entries touched here are reset
        // afterwards, so that a subset call costs O(subset), not O(N) (a sharded pedigree of 10^6+
        // individuals makes one such call per generation and rank).
        if ((int64_t)c->level_scratch.size() != N) c->level_scratch.assign((size_t)N, 0);
        std::vector<int32_t>& level = c->level_scratch;
        std::vector<int32_t> touched;
        int32_t max_level = 0;
        auto reset_levels = [&]() { for (int32_t i : touched) level[i] = 0; };
        auto visit = [&](int32_t i) -> bool {   // false: a parent is neither simulated nor scheduled
            if (c->hs_reps[i] || c->par0_h[i] < 0) return true;   // has a haplostruct, or is a founder about to get one
            int32_t lv = 0;
            for (int32_t par : {c->par0_h[i], c->par1_h[i]}) {
                if (c->hs_reps[par] || c->par0_h[par] < 0) continue;
                if (level[par] <= 0) return false;
                lv = std::max(lv, level[par]);
            }
            level[i] = lv + 1;
            touched.push_back(i);
            max_level = std::max(max_level, lv + 1);
            return true;
        };
        int64_t bad = -1;
        if (subset) {
            // parents-first order: a parent has the smaller index, so the list is walked ascending.
            // Drivers pass strictly ascending lists: only others are copied, sorted and de-duplicated.
            bool ascending = true;
            for (int64_t k = 1; k < n_subset && ascending; k++) ascending = subset[k - 1] < subset[k];
            std::vector<int32_t> sorted_copy;
            const int32_t* sub = subset;
            int64_t n_sub = n_subset;
            if (!ascending) {
                sorted_copy.assign(subset, subset + n_subset);
                std::sort(sorted_copy.begin(), sorted_copy.end());
                sorted_copy.erase(std::unique(sorted_copy.begin(), sorted_copy.end()), sorted_copy.end());
                sub = sorted_copy.data();
                n_sub = (int64_t)sorted_copy.size();
            }
            if (n_sub > 0 && (sub[0] < 0 || sub[n_sub - 1] >= N)) return fail(c, PSIM_EINVAL, "individual index out of bounds");
            touched.reserve((size_t)n_sub);
            for (int64_t k = 0; k < n_sub; k++) if (!visit(sub[k])) { bad = sub[k]; break; }
        } else {
            for (int64_t i = 0; i < N; i++) if (!visit((int32_t)i)) { bad = i; break; }
        }
        if (bad >= 0) {
            reset_levels();
            return fail(c, PSIM_ESTATE, "a parent of individual " + std::to_string(bad) + " has no haplostruct yet");
        }
        std::vector<int64_t> cnt(max_level + 2, 0);
        for (int32_t i : touched) cnt[level[i] + 1]++;
        for (int lv = 1; lv <= max_level + 1; lv++) cnt[lv] += cnt[lv - 1];
        pl.level_off = cnt;                                   // individuals of level lv: order[level_off[lv] .. level_off[lv+1])
        pl.order.assign((size_t)cnt[max_level + 1], 0);
        std::vector<int64_t> fill(cnt.begin(), cnt.end());
        for (int32_t i : touched) pl.order[(size_t)fill[level[i]]++] = i;   // (touched is ascending: so is every level)
        reset_levels();
        pl.max_level = max_level;
        pl.maxf_in = max_founder_allele;
        pl.on_device = false;
        if (fresh) { sp = std::move(pl); sp.valid = true; }
    }
    psim_ctx_impl::SimPlan& plan = fresh ? sp : c->plan_once;
    return sim_start(c, plan, !(reuse && sp.on_device), 1, plan.max_level, true);
}

// Enqueues the founders (if asked) and the generations lv_lo .. lv_hi of a plan whose lists are, or are now put, on the device.
static int sim_start(psim_ctx_impl* c, psim_ctx_impl::SimPlan& plan, bool upload, int lv_lo, int lv_hi, bool founders) {
    const int C = c->C, P = c->P, R = c->R;
    if (upload && !plan.order.empty()) {   // the device copies of another plan's generation lists are about to be overwritten
        if (c->lev_plan && c->lev_plan != &plan) c->lev_plan->on_device = false;
        c->lev_plan = &plan;
    }
    if (upload && founders && !plan.f_ind.empty()) {   // ... and of its founder lists
        if (c->find_plan && c->find_plan != &plan) c->find_plan->on_device = false;
        c->find_plan = &plan;
    }
    // --- batches: the founders, then every generation cut into pieces whose chiasma scratch stays
    // below about 1 GiB (a thread's column holds chiasma_cap entries of 9 bytes)
    psim_ctx_impl::SimRun& run = c->run;
    run.plan = &plan;
    run.batches.clear();
    {   // a unit per thread; its multivalents go round by round (tuning builds: PSIM_MEI_UPB units per block)
        static const int env_upb = tuning_int("PSIM_MEI_UPB");
        run.units_per_block = env_upb > 0 ? std::min(env_upb, MEI_THREADS) : env_upb < 0 ? (P == 2 ? MEI_THREADS : (MEI_THREADS - 32) / (P / 2)) : MEI_THREADS;
    }
    const int64_t max_blocks = std::max<int64_t>(1, (int64_t)(1LL << 30) / ((int64_t)MEI_THREADS * c->chiasma_cap * 9));
    const int64_t max_batch = std::max<int64_t>(1, max_blocks * run.units_per_block / ((int64_t)C * R * 2));
    if (founders && !plan.f_ind.empty()) run.batches.push_back({0, 0, (int64_t)plan.f_ind.size()});
    int64_t most_blocks = 0;
    for (int lv = std::max(1, lv_lo); lv <= std::min(lv_hi, (int)plan.max_level); lv++) {
        const int64_t lev_begin = plan.level_off[lv], lev_size = plan.level_off[lv + 1] - lev_begin;
        for (int64_t b0 = 0; b0 < lev_size; b0 += max_batch) {
            const int64_t n = std::min<int64_t>(max_batch, lev_size - b0);
            run.batches.push_back({lv, lev_begin + b0, n});
            most_blocks = std::max(most_blocks, (n * C * R * 2 + run.units_per_block - 1) / run.units_per_block);
        }
    }
    if (run.batches.empty()) {   // nothing to simulate (e.g. the founders of a sharded run that all exist already): the
        run.pending = false;     // device cursor is only valid behind a batch, so there is nothing to enqueue or settle
        return PSIM_OK;
    }
    if (founders && !plan.f_ind.empty()) {
        const size_t n_new = plan.f_ind.size();
        if (int rc = ensure_buf(c, c->b_find, n_new * sizeof(int32_t))) return rc;
        if (int rc = ensure_buf(c, c->b_fal, n_new * sizeof(int32_t))) return rc;
        if (upload) {   // (pageable sources: cudaMemcpyAsync has staged them before it returns)
            CUDA_TRY(cudaMemcpyAsync(c->b_find.p, plan.f_ind.data(), n_new * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(cudaMemcpyAsync(c->b_fal.p, plan.f_allele.data(), n_new * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        }
    }
    if (most_blocks > 0 || upload) {
        const size_t n_slots = (size_t)most_blocks * MEI_THREADS;
        if (int rc = ensure_buf(c, c->b_lev, plan.order.size() * sizeof(int32_t))) return rc;
        if (int rc = ensure_buf(c, c->b_xpos, n_slots * c->chiasma_cap * sizeof(double))) return rc;
        if (int rc = ensure_buf(c, c->b_xab, n_slots * c->chiasma_cap)) return rc;
        if (upload && !plan.order.empty())   // every generation list in one copy
            CUDA_TRY(cudaMemcpyAsync(c->b_lev.p, plan.order.data(), plan.order.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    }
    if (int rc = ensure_buf(c, c->b_mark, (run.batches.size() + 1) * sizeof(unsigned long long))) return rc;
    c->mark_d = (unsigned long long*)c->b_mark.p;
    if (upload) plan.on_device = true;
    c->fresh = false;
    // room for the founders and about two segments per homolog, so that the first run of a typical
    // pedigree does not have to be repeated (a slab that is too small is grown in sim_finish)
    {
        int64_t n_sim = 0;
        for (const auto& bt : run.batches) n_sim += bt.count;
        if (int rc = ensure_slab(c, c->slab_used + n_sim * R * C * P * 2)) return rc;
    }
    for (const auto& bt : run.batches)
        for (int64_t k = 0; k < bt.count; k++) c->hs_reps[bt.level == 0 ? plan.f_ind[(size_t)k] : plan.order[(size_t)(bt.first + k)]] = R;
    c->scalars_h[0] = 0; c->scalars_h[1] = 0; c->scalars_h[2] = 0;
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    run.pending = true;
    c->runs_valid = false;
    return sim_enqueue(c, 0, c->slab_used);
}

// Waits for an enqueued simulation and settles the slab: PSIM_OK, or the error of the first batch
// that failed (individuals of that batch and of later ones are left without a haplostruct).
static int sim_finish(psim_ctx_impl* c) {
    psim_ctx_impl::SimRun& run = c->run;
    if (!run.pending) return PSIM_OK;
    const psim_ctx_impl::SimPlan& pl = *run.plan;
    const int R = c->R;
    int rc_out = PSIM_OK;
    for (;;) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        const int64_t cursor = c->scalars_h[0];
        const int status = (int)c->scalars_h[1], failed = (int)c->scalars_h[2];
        if (status == 0) { c->slab_used = cursor; break; }
        unsigned long long mark = 0;
        CUDA_TRY(cudaMemcpy(&mark, c->mark_d + failed, sizeof(mark), cudaMemcpyDeviceToHost));
        c->slab_used = (int64_t)mark;
        if (status == STATUS_SLAB_FULL) {
            // `cursor` counts what the batches so far asked for; the rest is extrapolated
            int64_t done_ind = 0, all_ind = 0;
            for (size_t b = 0; b < run.batches.size(); b++) { all_ind += run.batches[b].count; if ((int)b <= failed) done_ind += run.batches[b].count; }
            const double per_ind = (double)std::max<int64_t>(cursor, c->slab_cap) / (double)std::max<int64_t>(1, done_ind);
            const int64_t want = (int64_t)(per_ind * (double)all_ind * 1.25) + 1024;
            if (int rc = ensure_slab(c, std::max<int64_t>(want, c->slab_cap + c->slab_cap / 2))) { rc_out = rc; }
            else {
                c->scalars_h[0] = 0; c->scalars_h[1] = 0; c->scalars_h[2] = 0;
                if (int rc2 = sim_enqueue(c, failed, c->slab_used)) rc_out = rc2; else continue;
            }
        } else {
            rc_out = fail(c, status, "more chiasmata on one multivalent than the device list holds");
        }
        // batches from `failed` on did not complete
        for (size_t b = (size_t)failed; b < run.batches.size(); b++) {
            const auto& bt = run.batches[b];
            for (int64_t k = 0; k < bt.count; k++) c->hs_reps[bt.level == 0 ? pl.f_ind[(size_t)k] : pl.order[(size_t)(bt.first + k)]] = 0;
        }
        k_slab_mark<<<1, 1, 0, c->stream>>>(c->cursor_d, c->mark_d, 0, 0, c->slab_cap, c->cursor_d + 1, c->status_d, c->slab_used);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        run.batches.resize((size_t)failed);
        break;
    }
    run.pending = false;
    for (const auto& bt : run.batches) {
        if (bt.level == 0) continue;
        c->stats.n_meioses += bt.count * R * (c->model.unreduced ? 1 : 2);
        c->stats.n_units += bt.count * R * (c->model.unreduced ? 1 : 2) * c->C;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->stats.ms_simulate = ms;
    c->stats.n_segments = c->slab_used;
    return rc_out;
}

static int simulate_impl(psim_ctx_impl* c, int32_t max_founder_allele, const int32_t* subset, int64_t n_subset) {
    if (int rc = settle(c)) return rc;
    if (int rc = sim_begin(c, max_founder_allele, subset, n_subset)) return rc;
    return sim_finish(c);
}

int psim_simulate(psim_ctx* ctx, int32_t max_founder_allele) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    return simulate_impl(c, max_founder_allele, nullptr, 0);
}

int psim_simulate_async(psim_ctx* ctx, int32_t max_founder_allele) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (int rc = settle(c)) return rc;
    return sim_begin(c, max_founder_allele, nullptr, 0);
}

int psim_wait(psim_ctx* ctx) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    return settle(c);
}

int psim_simulate_subset(psim_ctx* ctx, int32_t max_founder_allele, const int32_t* individuals, int64_t count) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (count > 0 && !individuals) return fail(c, PSIM_EINVAL, "null individual list");
    static const int32_t none = 0;
    return simulate_impl(c, max_founder_allele, individuals ? individuals : &none, count);
}

static int check_range(psim_ctx_impl* c, uint32_t replicate, int64_t first, int64_t count, int& r) {
    if (int rc = settle(c)) return rc;
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    if (first < 0 || count <= 0 || first + count > c->N) return fail(c, PSIM_EINVAL, "individual range out of bounds");
    if (replicate < c->cfg.replicate_first || replicate >= c->cfg.replicate_first + c->cfg.replicate_count)
        return fail(c, PSIM_EINVAL, "replicate not held by this context");
    for (int64_t i = first; i < first + count; i++)
        if (c->hs_reps[i] != c->R) return fail(c, PSIM_ESTATE, "individual " + std::to_string(i) + " has no haplostruct yet");
    r = (int)(replicate - c->cfg.replicate_first);
    return PSIM_OK;
}

static int gather_to_export(psim_ctx_impl* c, int r, int64_t first, int64_t count, int64_t& nseg_out) {
    const int64_t nh = count * c->C * c->P;
    if (nh + 1 > c->exp_off_cap) {
        cudaFree(c->exp_off_d); c->exp_off_d = nullptr;
        CUDA_TRY(dev_alloc(&c->exp_off_d, (size_t)(nh + 1) * 2));  // counts + offsets
        c->exp_off_cap = nh + 1;
    }
    int64_t* cnt = c->exp_off_d + c->exp_off_cap;
    int64_t* off = c->exp_off_d;
    k_gather_counts<<<grid_for(nh, 256), 256, 0, c->stream>>>(nh, first, r, c->C, c->R, c->N, c->P, c->desc_d, cnt);
    c->stats.kernel_launches++;
    CUDA_TRY(cudaMemsetAsync(cnt + nh, 0, sizeof(int64_t), c->stream));
    if (int rc = exclusive_scan_i64(c, cnt, off, nh + 1)) return rc;
    int64_t nseg = 0;
    CUDA_TRY(cudaMemcpyAsync(&nseg, off + nh, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (nseg > c->exp_seg_cap) {
        cudaFree(c->exp_founder_d); cudaFree(c->exp_pos_d);
        c->exp_founder_d = nullptr; c->exp_pos_d = nullptr;
        CUDA_TRY(dev_alloc(&c->exp_founder_d, (size_t)nseg));
        CUDA_TRY(dev_alloc(&c->exp_pos_d, (size_t)nseg));
        c->exp_seg_cap = nseg;
    }
    k_gather_segments<<<grid_for(nh, 256), 256, 0, c->stream>>>(nh, first, r, c->C, c->R, c->N, c->P, c->desc_d, off,
                                                                  c->seg_pos_d, c->seg_founder_d, c->exp_pos_d, c->exp_founder_d);
    c->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    nseg_out = nseg;
    return PSIM_OK;
}

int psim_segment_count(psim_ctx* ctx, uint32_t replicate, int64_t first, int64_t count, int64_t* n_segments) {
    psim_ctx_impl* c = CTX;
    if (!c || !n_segments) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    int r;
    if (int rc = check_range(c, replicate, first, count, r)) return rc;
    return gather_to_export(c, r, first, count, *n_segments);
}

int psim_get_haplostructs(psim_ctx* ctx, uint32_t replicate, int64_t first, int64_t count,
                          int64_t* seg_off, int32_t* founder, double* pos) {
    psim_ctx_impl* c = CTX;
    if (!c || !seg_off || !founder || !pos) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    int r;
    if (int rc = check_range(c, replicate, first, count, r)) return rc;
    int64_t nseg = 0;
    if (int rc = gather_to_export(c, r, first, count, nseg)) return rc;
    const int64_t nh = count * c->C * c->P;
    CUDA_TRY(cudaMemcpyAsync(seg_off, c->exp_off_d, (nh + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(founder, c->exp_founder_d, nseg * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(pos, c->exp_pos_d, nseg * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PSIM_OK;
}

int psim_get_meiotic_configs(psim_ctx* ctx, uint32_t replicate, int64_t first, int64_t count, int32_t* quad, int32_t* chromseq) {
    psim_ctx_impl* c = CTX;
    if (!c || !quad || !chromseq) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    if (first < 0 || count <= 0 || first + count > c->N) return fail(c, PSIM_EINVAL, "individual range out of bounds");
    if (replicate < c->cfg.replicate_first || replicate >= c->cfg.replicate_first + c->cfg.replicate_count)
        return fail(c, PSIM_EINVAL, "replicate not held by this context");
    const int r = (int)(replicate - c->cfg.replicate_first);
    const int64_t nq = count * c->C * 2;
    std::vector<int8_t> q(nq), s(nq * c->P);
    const int64_t base = ((int64_t)r * c->N + first) * c->C * 2;
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // psim_set_pedigree / psim_reset clear the tables on the stream
    CUDA_TRY(cudaMemcpy(q.data(), c->cfg_quad_d + base, nq, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(s.data(), c->cfg_seq_d + base * c->P, nq * c->P, cudaMemcpyDeviceToHost));
    for (int64_t k = 0; k < nq; k++) quad[k] = q[k];
    for (int64_t k = 0; k < nq * c->P; k++) chromseq[k] = s[k];
    return PSIM_OK;
}

int psim_set_map(psim_ctx* ctx, const int64_t* locus_off, const double* pos) {
    psim_ctx_impl* c = CTX;
    if (!c || !locus_off || !pos) return PSIM_EINVAL;
    if (c->C == 0) return fail(c, PSIM_ESTATE, "psim_set_chromosomes must be called first");
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    if (locus_off[0] != 0) return fail(c, PSIM_EINVAL, "locus_off[0] must be 0");
    for (int k = 0; k < c->C; k++) {
        if (locus_off[k + 1] < locus_off[k]) return fail(c, PSIM_EINVAL, "locus_off must be non-decreasing");
        for (int64_t l = locus_off[k]; l < locus_off[k + 1]; l++) {
            // HaploStruct.getFounderAt (:135-142) rejects positions outside [head, tail]
            if (pos[l] < c->chrom_h[k].head || pos[l] > c->chrom_h[k].tail) return fail(c, PSIM_EINVAL, "getFounderAt: invalid position");
            if (l > locus_off[k] && pos[l] < pos[l - 1]) return fail(c, PSIM_EINVAL, "map positions must be ascending within a chromosome");
        }
    }
    c->L = locus_off[c->C];
    c->locus_off_h.assign(locus_off, locus_off + c->C + 1);
    c->locus_off_d = nullptr; c->locus_pos_d = nullptr; c->code_d = nullptr; c->match_d = nullptr;
    c->code16_d = nullptr; c->nibmask_d = nullptr; c->mtab_d = nullptr;
    c->A = 0; c->code_h.clear();
    if (int rc = ensure_buf(c, c->b_locoff, ((size_t)c->C + 1) * sizeof(int64_t))) return rc;
    if (int rc = ensure_buf(c, c->b_locpos, std::max<size_t>(1, (size_t)c->L) * sizeof(double))) return rc;
    c->locus_off_d = (int64_t*)c->b_locoff.p; c->locus_pos_d = (double*)c->b_locpos.p;
    CUDA_TRY(cudaMemcpyAsync(c->locus_off_d, locus_off, (c->C + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->locus_pos_d, pos, c->L * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    {   // the grid k_index_runs searches the map through: about four buckets per locus
        std::vector<int64_t> goff((size_t)c->C + 1, 0);
        std::vector<double> glo((size_t)c->C, 0.0), ginv((size_t)c->C, 0.0);
        std::vector<int32_t> grid;
        for (int k = 0; k < c->C; k++) {
            const double* p0 = pos + locus_off[k];
            const int64_t Lc = locus_off[k + 1] - locus_off[k];
            const double span = Lc > 0 ? p0[Lc - 1] - p0[0] : 0.0;
            const int64_t nb = span > 0.0 ? std::max<int64_t>(16, std::min<int64_t>(4 * Lc, 1 << 22)) : 1;
            const double width = span > 0.0 ? span / (double)nb : 0.0;
            glo[(size_t)k] = Lc > 0 ? p0[0] : 0.0;
            ginv[(size_t)k] = width > 0.0 ? 1.0 / width : 0.0;
            int64_t at = 0;   // (bucket starts and positions both ascend: one merge pass)
            for (int64_t q = 0; q < nb; q++) {
                const double start = glo[(size_t)k] + (double)q * width;
                while (at < Lc && p0[at] < start) at++;
                grid.push_back((int32_t)at);
            }
            grid.push_back((int32_t)Lc);
            goff[(size_t)k + 1] = (int64_t)grid.size();
        }
        if (int rc = ensure_buf(c, c->b_grid, grid.size() * sizeof(int32_t))) return rc;
        if (int rc = ensure_buf(c, c->b_gridoff, goff.size() * sizeof(int64_t))) return rc;
        if (int rc = ensure_buf(c, c->b_gridlo, glo.size() * sizeof(double))) return rc;
        if (int rc = ensure_buf(c, c->b_gridinv, ginv.size() * sizeof(double))) return rc;
        c->grid_d = (int32_t*)c->b_grid.p; c->grid_off_d = (int64_t*)c->b_gridoff.p;
        c->grid_lo_d = (double*)c->b_gridlo.p; c->grid_inv_d = (double*)c->b_gridinv.p;
        CUDA_TRY(cudaMemcpyAsync(c->grid_d, grid.data(), grid.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->grid_off_d, goff.data(), goff.size() * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->grid_lo_d, glo.data(), glo.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->grid_inv_d, ginv.data(), ginv.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));   // (the sources are this call's and the caller's arrays)
    }
    c->runs_valid = false;
    return PSIM_OK;
}

int psim_set_allele_codes(psim_ctx* ctx, int32_t n_alleles, const uint8_t* code) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (c->L == 0) return fail(c, PSIM_ESTATE, "psim_set_map must be called first");
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    c->code_d = nullptr; c->match_d = nullptr; c->code16_d = nullptr; c->nibmask_d = nullptr; c->mtab_d = nullptr;
    c->A = 0; c->code_h.clear(); c->match_which = 12345678;
    if (!code) return PSIM_OK;
    // (up to 256 unique allele NAMES per locus - the codes are bytes - for any number of founder alleles: tables
    // with more than 256 founder alleles are emitted by the per-cell kernels, which index code[l][founder])
    if (n_alleles <= 0 || n_alleles > 65536) return fail(c, PSIM_EINVAL, "allele codes need 1..65536 founder alleles");
    c->A = n_alleles;
    c->code_h.assign(code, code + (size_t)c->L * n_alleles);
    if (int rc = ensure_buf(c, c->b_code, (size_t)c->L * n_alleles)) return rc;
    if (int rc = ensure_buf(c, c->b_match, (size_t)c->L)) return rc;
    if (int rc = ensure_buf(c, c->b_code16, (size_t)c->L * 4 * sizeof(uint32_t))) return rc;
    if (int rc = ensure_buf(c, c->b_nibmask, (size_t)c->L * 2 * sizeof(uint32_t))) return rc;
    if (int rc = ensure_buf(c, c->b_mtab, (size_t)c->L * sizeof(uint4))) return rc;
    c->code_d = (uint8_t*)c->b_code.p; c->match_d = (uint8_t*)c->b_match.p; c->code16_d = (uint32_t*)c->b_code16.p;
    c->nibmask_d = (uint32_t*)c->b_nibmask.p; c->mtab_d = (uint4*)c->b_mtab.p;
    CUDA_TRY(cudaMemcpyAsync(c->code_d, code, (size_t)c->L * n_alleles, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // (c->code_h holds a copy, but the contract is that `code` is free on return)
    return PSIM_OK;
}

static int ensure_runs(psim_ctx_impl* c) {
    if (c->runs_valid) return PSIM_OK;
    if (c->run_cap < c->slab_used) {
        cudaFree(c->run_lb_d); cudaFree(c->run_founder_d);
        c->run_lb_d = nullptr; c->run_founder_d = nullptr;
        CUDA_TRY(dev_alloc(&c->run_lb_d, (size_t)c->slab_cap));
        CUDA_TRY(dev_alloc(&c->run_founder_d, (size_t)c->slab_cap));
        c->run_cap = c->slab_cap;
    }
    k_index_runs<<<grid_for(c->n_hom, 128), 128, 0, c->stream>>>(c->n_hom, (int64_t)c->R * c->N * c->P, c->desc_d, c->seg_pos_d,
                                                                 c->seg_founder_d, c->locus_off_d, c->locus_pos_d,
                                                                 MapGrid{c->grid_d, c->grid_off_d, c->grid_lo_d, c->grid_inv_d}, c->run_desc_d,
                                                                 c->run_lb_d, c->run_founder_d);
    c->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    c->runs_valid = true;
    return PSIM_OK;
}

static int ensure_staging(psim_ctx_impl* c, size_t bytes) {
    if (bytes <= c->staging_bytes) return PSIM_OK;
    cudaFree(c->staging); c->staging = nullptr; c->staging_bytes = 0;
    CUDA_TRY(cudaMalloc(&c->staging, bytes));
    c->staging_bytes = bytes;
    return PSIM_OK;
}

// Emit loci [l_begin, l_end) (global) into device tables whose row 0 is locus `row0_locus`.
static int emit_device(psim_ctx_impl* c, int64_t ind_first, int64_t ind_count, int64_t l_begin, int64_t l_end, int64_t row0_locus,
                       void* founder, int64_t founder_pitch, int founder_bytes, uint8_t* allele, int64_t allele_pitch,
                       uint8_t* dose, int64_t dose_pitch, int dose_mode, int dose_which, int path = PSIM_PATH_AUTO, bool packed = false) {
    const int P = c->P;
    const int64_t n_cols_ind = (int64_t)c->R * ind_count;
    if (packed) {   // two 4-bit cells per byte: the row-image kernels, else one thread per pair of individuals
        const bool aligned16 = (!founder || ((uintptr_t)founder % 16 == 0 && founder_pitch % 16 == 0)) &&
                               (!dose || ((uintptr_t)dose % 16 == 0 && dose_pitch % 16 == 0));
        if (aligned16 && c->run_cap < (int64_t)0xFFFFFFFF && path != PSIM_PATH_GENERIC) {
            bool done = false;
            const int mode0 = !c->code_d ? 0 : c->A <= 8 ? 1 : 2;
            if (int rc = emit_rows_path(c, ind_first, ind_count, l_begin, l_end, row0_locus, (uint8_t*)founder, founder_pitch, dose, dose_pitch,
                                        dose_mode, dose_which, mode0, path, true, done)) return rc;
            if (done) return PSIM_OK;
        }
        EmitPackedArgs pa{};
        pa.C = c->C; pa.R = c->R; pa.P = P; pa.N = c->N;
        pa.run_desc = c->run_desc_d; pa.run_lb = c->run_lb_d; pa.run_founder = c->run_founder_d; pa.locus_off = c->locus_off_d;
        pa.ind_first = ind_first; pa.ind_count = ind_count; pa.n_cols_ind = n_cols_ind;
        pa.locus_first = row0_locus; pa.l_begin = l_begin; pa.l_count = l_end - l_begin;
        pa.q_first = 0; pa.q_count = n_cols_ind;
        pa.founder = (uint8_t*)founder; pa.founder_pitch = founder_pitch; pa.dose = dose; pa.dose_pitch = dose_pitch;
        pa.dose_mode = dose_mode; pa.dose_which = dose_which; pa.code = c->code_d; pa.match_code = c->match_d; pa.A = c->A;
        cudaEventRecord(next_kev(c), c->stream);
        k_emit_generic_packed<<<grid_for(pa.l_count * ((n_cols_ind + 1) / 2), 256), 256, 0, c->stream>>>(pa);
        cudaEventRecord(next_kev(c), c->stream);
        c->stats.kernel_launches++;
        c->stats.emit_kernel_launches++;
        CUDA_TRY(cudaGetLastError());
        return PSIM_OK;
    }
    int G = 0;
    if (P == 2) G = 8; else if (P == 4) G = 4; else if (P == 6) G = 8; else if (P == 8) G = 2;
    const bool aligned = (!founder || ((uintptr_t)founder % 16 == 0 && founder_pitch % 16 == 0)) &&
                         (!allele || ((uintptr_t)allele % 16 == 0 && allele_pitch % 16 == 0)) &&
                         (!dose || ((uintptr_t)dose % 16 == 0 && dose_pitch % 16 == 0));
    const int64_t max_pitch = (int64_t)1 << 25;   // row offsets inside a tile are 32-bit (64 rows)
    const bool small_pitch = founder_pitch < max_pitch && allele_pitch < max_pitch && dose_pitch < max_pitch;
    bool tiled = G != 0 && aligned && small_pitch && founder_bytes == 1 && c->founder_allele_count <= 256 &&
                 c->run_cap < (int64_t)0xFFFFFFFF;
    if (tiled && !allele && (founder || dose)) {
        bool done = false;
        const int mode0 = !c->code_d ? 0 : c->A <= 8 ? 1 : c->A <= 16 ? 2 : 3;
        if (int rc = emit_rows_path(c, ind_first, ind_count, l_begin, l_end, row0_locus, (uint8_t*)founder, founder_pitch, dose, dose_pitch,
                                    dose_mode, dose_which, mode0, path, false, done)) return rc;
        if (done) return PSIM_OK;
    }
    if (path == PSIM_PATH_GENERIC) tiled = false;
    if (tiled) {
        EmitArgs ea{};
        ea.C = c->C; ea.R = c->R; ea.P = P; ea.N = c->N;
        ea.run_desc = c->run_desc_d; ea.run_lb = c->run_lb_d; ea.run_founder = c->run_founder_d;
        ea.locus_off = c->locus_off_d;
        ea.ind_first = ind_first; ea.ind_count = ind_count; ea.n_cols_ind = n_cols_ind;
        ea.locus_first = row0_locus;
        static const int env_tile = tuning_int("PSIM_TILE_LOCI");
        static const int env_bps = tuning_int("PSIM_BLOCKS_PER_SM");
        {   // rows per tile: as tall as possible (the plan is read once per tile) while a warp's expected
            // number of run starts per tile stays far below EMIT_EV
            const double seg_per_hom = c->n_hom > 0 ? (double)c->slab_used / (double)c->n_hom : 1.0;
            const double loci_per_chrom = std::max(1.0, (double)c->L / c->C);
            const double starts_per_cell = std::max(0.0, seg_per_hom - 1.0) / loci_per_chrom;
            int rows = 64;
            while (rows > 8 && 32.0 * G * P * rows * starts_per_cell > 18.0) rows >>= 1;
            ea.tile_loci = env_tile > 0 ? std::min(env_tile, EMIT_MAX_TILE_LOCI) : rows;
        }
        ea.n_col_tiles = (int)((n_cols_ind + (int64_t)EMIT_THREADS * G - 1) / ((int64_t)EMIT_THREADS * G));
        ea.founder = (uint8_t*)founder; ea.founder_pitch = founder_pitch;
        ea.allele = allele; ea.allele_pitch = allele_pitch;
        ea.dose = dose; ea.dose_pitch = dose_pitch;
        ea.dose_which = dose_which;
        ea.code = c->code_d; ea.match_code = c->match_d; ea.A = c->A;
        ea.code16 = c->code16_d; ea.nibmask = c->nibmask_d; ea.mtab = c->mtab_d;
        const int mode = !c->code_d ? 0 : c->A <= 8 ? 1 : c->A <= 16 ? 2 : 3;
        ea.tile_counter = c->tile_counter_d;
        // chromosomes touched by [l_begin, l_end), at most 64 per launch
        int k0 = 0;
        while (k0 < c->C && c->locus_off_h[k0 + 1] <= l_begin) k0++;
        while (k0 < c->C && c->locus_off_h[k0] < l_end) {
            ea.n_chunks = 0;
            ea.chunk_tile_base[0] = 0;
            while (k0 < c->C && c->locus_off_h[k0] < l_end && ea.n_chunks < 64) {
                const int64_t a0 = std::max(l_begin, c->locus_off_h[k0]), a1 = std::min(l_end, c->locus_off_h[k0 + 1]);
                if (a1 > a0) {
                    const int q = ea.n_chunks++;
                    ea.chunk_chrom[q] = k0; ea.chunk_l0[q] = a0; ea.chunk_l1[q] = a1;
                    const int64_t n_lt = (a1 - a0 + ea.tile_loci - 1) / ea.tile_loci;
                    ea.chunk_tile_base[q + 1] = ea.chunk_tile_base[q] + (int)(n_lt * ea.n_col_tiles);
                }
                k0++;
            }
            if (ea.n_chunks == 0) break;
            ea.n_tiles = ea.chunk_tile_base[ea.n_chunks];
            {
                const size_t nt = (size_t)ea.n_tiles;
                ea.plan_blk_size = 2 + (G * P / 16) * EMIT_THREADS + EMIT_EV;
                if (int rc = ensure_buf(c, c->b_plan, nt * ea.plan_blk_size * sizeof(uint4))) return rc;
                if (int rc = ensure_buf(c, c->b_dense_list, (nt * (EMIT_THREADS / 32) + 1) * sizeof(int))) return rc;
                ea.plan_blk = (uint4*)c->b_plan.p;
                ea.dense_list = (int*)c->b_dense_list.p;
                CUDA_TRY(cudaMemsetAsync(ea.plan_blk, 0, nt * ea.plan_blk_size * sizeof(uint4), c->stream));
                CUDA_TRY(cudaMemsetAsync(ea.dense_list, 0, sizeof(int), c->stream));
            }
            CUDA_TRY(cudaMemsetAsync(c->tile_counter_d, 0, sizeof(int), c->stream));
            const int blocks = std::min(ea.n_tiles, c->sm_count * (env_bps > 0 ? env_bps : 4));
            if (P == 2) launch_tiles<2, 8>(c, ea, blocks, mode);
            else if (P == 4) launch_tiles<4, 4>(c, ea, blocks, mode);
            else if (P == 6) launch_tiles<6, 8>(c, ea, blocks, mode);
            else launch_tiles<8, 2>(c, ea, blocks, mode);
            c->stats.kernel_launches += 3;
            c->stats.emit_kernel_launches++;
            CUDA_TRY(cudaGetLastError());
        }
    } else {
        EmitGenericArgs ga{};
        ga.C = c->C; ga.R = c->R; ga.P = P; ga.N = c->N;
        ga.run_desc = c->run_desc_d; ga.run_lb = c->run_lb_d; ga.run_founder = c->run_founder_d;
        ga.locus_off = c->locus_off_d;
        ga.ind_first = ind_first; ga.ind_count = ind_count; ga.n_cols_ind = n_cols_ind;
        ga.locus_first = row0_locus; ga.l_begin = l_begin; ga.l_count = l_end - l_begin;
        ga.q_first = 0; ga.q_count = n_cols_ind;
        ga.founder = founder; ga.founder_pitch = founder_pitch; ga.founder_bytes = founder_bytes;
        ga.allele = allele; ga.allele_pitch = allele_pitch; ga.dose = dose; ga.dose_pitch = dose_pitch;
        ga.dose_mode = dose_mode; ga.dose_which = dose_which;
        ga.code = c->code_d; ga.match_code = c->match_d; ga.A = c->A;
        const int64_t nt = ga.l_count * n_cols_ind;
        cudaEventRecord(next_kev(c), c->stream);
        k_emit_generic<<<grid_for(nt, 256), 256, 0, c->stream>>>(ga);
        cudaEventRecord(next_kev(c), c->stream);
        c->stats.kernel_launches++;
        c->stats.emit_kernel_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return PSIM_OK;
}

// device times of the last emission from the events recorded around its kernels (after the stream has been waited for)
static void emit_times(psim_ctx_impl* c, bool device_targets) {
    float ms_total = 0.f, ms_kernel = 0.f;
    cudaEventElapsedTime(&ms_total, c->ev0, c->ev1);
    for (int i = device_targets ? 0 : 2; i + 1 < c->kev_used; i += 2) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->kev[i], c->kev[i + 1]);
        ms_kernel += ms;
    }
    c->stats.ms_emit_kernel = ms_kernel;
    c->stats.ms_emit_total = ms_total;
    c->stats.ms_emit_plan = 0.0;
}
}  // extern "C"
namespace {
// Completes what psim_simulate_async / an asynchronous psim_emit left running. Every entry point that
// reads or changes the population calls it first.
int settle(psim_ctx_impl* c) {
    int rc = sim_finish(c);
    if (c->emit_pending) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        c->emit_pending = false;
        if (e != cudaSuccess) return fail(c, PSIM_ECUDA, std::string("asynchronous emission: ") + cudaGetErrorString(e));
        emit_times(c, c->emit_pending_device);
    }
    return rc;
}
}  // namespace
extern "C" {

static int emit_prepare(psim_ctx_impl* c, int64_t ind_first, int64_t ind_count, int64_t locus_first, int64_t locus_count) {
    if (int rc = settle(c)) return rc;
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    if (c->L == 0) return fail(c, PSIM_ESTATE, "psim_set_map must be called first");
    if (ind_first < 0 || ind_count <= 0 || ind_first + ind_count > c->N) return fail(c, PSIM_EINVAL, "individual range out of bounds");
    if (locus_first < 0 || locus_count <= 0 || locus_first + locus_count > c->L) return fail(c, PSIM_EINVAL, "locus range out of bounds");
    for (int64_t i = ind_first; i < ind_first + ind_count; i++)
        if (c->hs_reps[i] != c->R) return fail(c, PSIM_ESTATE, "individual " + std::to_string(i) + " has no haplostruct yet");
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));   // psim_emit's ms_emit_total counts K0 too (the other callers record their own start)
    return ensure_runs(c);
}

int psim_emit(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int64_t locus_first, int64_t locus_count,
              const psim_emit_targets* t) {
    psim_ctx_impl* c = CTX;
    if (!c || !t) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (int rc = emit_prepare(c, ind_first, ind_count, locus_first, locus_count)) return rc;
    if (!t->founder && !t->allele && !t->dose) return fail(c, PSIM_EINVAL, "no output table requested");
    const int fb = t->founder ? (t->founder_bytes == 0 ? 1 : t->founder_bytes) : 1;
    if (fb != 1 && fb != 2) return fail(c, PSIM_EINVAL, "founder_bytes must be 1 or 2");
    if (fb == 1 && t->founder && c->founder_allele_count > 256) return fail(c, PSIM_EINVAL, "more than 256 founder alleles need founder_bytes = 2");
    if (t->allele && !c->code_d) return fail(c, PSIM_ESTATE, "allele table requested but no allele codes set");
    const int64_t n_cols_ind = (int64_t)c->R * ind_count;
    const int64_t n_cols = n_cols_ind * c->P;
    const bool packed = (t->flags & PSIM_EMIT_PACKED4) != 0;
    if (packed) {
        if (t->allele) return fail(c, PSIM_EINVAL, "PSIM_EMIT_PACKED4 has no allele table");
        if (fb != 1) return fail(c, PSIM_EINVAL, "PSIM_EMIT_PACKED4 needs founder_bytes = 1");
        if (c->founder_allele_count > 16) return fail(c, PSIM_EINVAL, "PSIM_EMIT_PACKED4 needs at most 16 founder alleles");
        if (c->P > 14) return fail(c, PSIM_EINVAL, "PSIM_EMIT_PACKED4 needs ploidy <= 14");
    }
    const int64_t w_founder = packed ? (n_cols + 1) / 2 : n_cols * fb, w_dose = packed ? (n_cols_ind + 1) / 2 : n_cols_ind;
    if (t->founder && t->founder_pitch < w_founder) return fail(c, PSIM_EINVAL, "founder_pitch too small");
    if (t->allele && t->allele_pitch < n_cols) return fail(c, PSIM_EINVAL, "allele_pitch too small");
    if (t->dose && t->dose_pitch < w_dose) return fail(c, PSIM_EINVAL, "dose_pitch too small");
    int dose_mode = 0, dose_which = t->dose_which;
    if (c->code_d) {
        dose_mode = 1;
        if (c->match_which != t->dose_which) {
            k_match_codes<<<grid_for(c->L, 256), 256, 0, c->stream>>>(c->L, c->A, t->dose_which, c->code_d, c->match_d,
                                                                       c->code16_d, c->nibmask_d, c->mtab_d);
            c->stats.kernel_launches++;
            CUDA_TRY(cudaGetLastError());
            c->match_which = t->dose_which;
        }
    } else if (dose_which < 0) {
        dose_which = 0x7FFFFFFF;  // Main.java:1010-1014: no founder allele equals a negative index -> all zero
    }
    c->stats.emit_kernel_launches = 0;
    c->kev_used = 0;
    // (ms_emit_total starts in emit_prepare, before the breakpoint -> locus-run index K0 of a new population)
    const int path = (t->flags >> 8) & 3;
    if (t->memory == PSIM_MEM_DEVICE) {
        if (int rc = emit_device(c, ind_first, ind_count, locus_first, locus_first + locus_count, locus_first, t->founder,
                                 t->founder_pitch, fb, (uint8_t*)t->allele, t->allele_pitch, (uint8_t*)t->dose, t->dose_pitch,
                                 dose_mode, dose_which, path, packed)) return rc;
        CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
        if (t->flags & PSIM_EMIT_ASYNC) { c->emit_pending = true; c->emit_pending_device = true; return PSIM_OK; }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    } else {
        if (t->flags & PSIM_EMIT_ASYNC) return fail(c, PSIM_EINVAL, "PSIM_EMIT_ASYNC needs device targets (PSIM_MEM_DEVICE)");
        // Host targets: locus chunks are emitted into one of two device slots and copied out on a
        // second stream while the next chunk is computed. A table whose host pitch is not 16-byte
        // aligned is written with an aligned pitch and repacked on the device (2-D device copy),
        // so that the PCIe transfer is always one large contiguous copy.
        // direct: dense and aligned host rows, the kernel writes the host layout; repack: dense but
        // unaligned; strided: padded host rows (a view into a wider table) are copied row by row
        struct Tab { void* host; int64_t hpitch, width; bool direct; int64_t spitch; bool strided; };
        auto up = [](int64_t x, int64_t m) { return (x + m - 1) / m * m; };
        Tab tab[3] = {{t->founder, t->founder_pitch, w_founder, false, 0, false},
                      {t->allele, t->allele_pitch, n_cols, false, 0, false},
                      {t->dose, t->dose_pitch, w_dose, false, 0, false}};
        int64_t row_bytes = 0;
        for (Tab& x : tab) {
            if (!x.host) continue;
            x.strided = x.hpitch != x.width;
            x.direct = !x.strided && x.width % 16 == 0;
            x.spitch = x.direct ? x.hpitch : up(x.width, 128);
            row_bytes += x.spitch + (x.direct || x.strided ? 0 : x.hpitch);
        }
        const int64_t slot_budget = 384LL << 20;   // few large chunks: every chunk boundary is a small bubble on the PCIe link
        int64_t rows = std::max<int64_t>(1, std::min<int64_t>(locus_count, slot_budget / row_bytes));
        // chunk boundaries on 256-byte boundaries of every host table where the pitches allow it: a
        // PCIe copy into an unaligned host address runs ~10 % slower
        int64_t row_quantum = 1;
        for (const Tab& x : tab)
            if (x.host) row_quantum = std::lcm(row_quantum, (int64_t)256 / std::gcd((int64_t)256, x.hpitch));
        if (rows >= 2 * row_quantum) rows = rows / row_quantum * row_quantum; else row_quantum = 1;
        const int64_t slot_bytes = up(rows * row_bytes, 256) + 6 * 256;   // every sub-buffer starts 256-byte aligned
        if (int rc = ensure_staging(c, (size_t)(2 * slot_bytes))) return rc;
        cudaEvent_t ready[2] = {c->ev2, c->ev3};
        cudaEvent_t freed[2] = {next_kev(c), next_kev(c)};
        static const bool trace = tuning_int("PSIM_TRACE") != 0;
        std::vector<cudaEvent_t> tev;
        int k = 0;
        // the first chunk is short so that the PCIe link starts early
        for (int64_t l0 = locus_first, l1; l0 < locus_first + locus_count; l0 = l1, k++) {
            l1 = std::min(l0 + (k == 0 ? std::max<int64_t>(1, rows / 16 / row_quantum) * row_quantum : rows), locus_first + locus_count);
            const int64_t nr = l1 - l0, r0 = l0 - locus_first;
            const int sl = k & 1;
            uint8_t* base = (uint8_t*)c->staging + sl * slot_bytes;
            uint8_t* sp[3]; uint8_t* dn[3];
            {
                uint8_t* p = base;
                for (int q = 0; q < 3; q++) {
                    sp[q] = dn[q] = nullptr;
                    if (!tab[q].host) continue;
                    sp[q] = p; p += up(rows * tab[q].spitch, 256);
                    if (!tab[q].direct && !tab[q].strided) { dn[q] = p; p += up(rows * tab[q].hpitch, 256); }
                }
            }
            if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(c->stream, freed[sl], 0));   // slot copied out
            if (trace) { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); fprintf(stderr, "psim_emit chunk %d begins at host %.3f ms\n", k, ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6); }
            if (int rc = emit_device(c, ind_first, ind_count, l0, l1, l0, sp[0], tab[0].spitch, fb, sp[1], tab[1].spitch,
                                     sp[2], tab[2].spitch, dose_mode, dose_which, path, packed)) return rc;
            for (int q = 0; q < 3; q++)
                if (dn[q]) {   // (a 2-D device copy with an unaligned pitch runs at a fraction of this kernel's pace)
                    k_repack_dense<<<c->sm_count * 8, 256, 0, c->stream>>>(sp[q], tab[q].spitch, dn[q], tab[q].width, nr);
                    c->stats.kernel_launches++;
                }
            CUDA_TRY(cudaEventRecord(ready[sl], c->stream));
            CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, ready[sl], 0));
            if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->copy_stream); tev.push_back(e); }
            for (int q = 0; q < 3; q++) {
                if (!tab[q].host) continue;
                uint8_t* dst = (uint8_t*)tab[q].host + r0 * tab[q].hpitch;
                if (tab[q].strided)
                    CUDA_TRY(cudaMemcpy2DAsync(dst, tab[q].hpitch, sp[q], tab[q].spitch, tab[q].width, nr, cudaMemcpyDeviceToHost, c->copy_stream));
                else
                    CUDA_TRY(cudaMemcpyAsync(dst, tab[q].direct ? sp[q] : dn[q], (size_t)(nr * tab[q].width), cudaMemcpyDeviceToHost, c->copy_stream));
            }
            CUDA_TRY(cudaEventRecord(freed[sl], c->copy_stream));
            if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->copy_stream); tev.push_back(e); }
            if (trace) { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); fprintf(stderr, "psim_emit chunk %d enqueued at host %.3f ms\n", k, ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6); }
        }
        CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
        CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (trace) {   // PSIM_TRACE: when every chunk's copies started and ended, ms after the call began
            for (size_t i = 0; i + 1 < tev.size(); i += 2) {
                float a = 0.f, b = 0.f;
                cudaEventElapsedTime(&a, c->ev0, tev[i]);
                cudaEventElapsedTime(&b, c->ev0, tev[i + 1]);
                fprintf(stderr, "psim_emit chunk %zu: copies start %.3f ms end %.3f ms\n", i / 2, a, b);
            }
            for (cudaEvent_t e : tev) cudaEventDestroy(e);
        }
    }
    emit_times(c, t->memory == PSIM_MEM_DEVICE);
    return PSIM_OK;
}

int psim_set_locus_names(psim_ctx* ctx, const int64_t* name_off, const char* chars) {
    psim_ctx_impl* c = CTX;
    if (!c || !name_off || !chars) return PSIM_EINVAL;
    if (c->L == 0) return fail(c, PSIM_ESTATE, "psim_set_map must be called first");
    cudaSetDevice(c->cfg.device);
    if (name_off[0] != 0) return fail(c, PSIM_EINVAL, "name_off[0] must be 0");
    c->label_max = 0;
    for (int64_t l = 0; l < c->L; l++) {
        if (name_off[l + 1] < name_off[l]) return fail(c, PSIM_EINVAL, "name_off must not decrease");
        c->label_max = std::max(c->label_max, name_off[l + 1] - name_off[l]);
    }
    cudaFree(c->label_off_d); cudaFree(c->labels_d);
    c->label_off_d = nullptr; c->labels_d = nullptr;
    CUDA_TRY(dev_alloc(&c->label_off_d, (size_t)c->L + 1));
    CUDA_TRY(dev_alloc(&c->labels_d, (size_t)name_off[c->L]));
    CUDA_TRY(cudaMemcpy(c->label_off_d, name_off, ((size_t)c->L + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->labels_d, chars, (size_t)name_off[c->L], cudaMemcpyHostToDevice));
    return PSIM_OK;
}

int psim_set_allele_names(psim_ctx* ctx, const int64_t* locus_first_name, const int64_t* name_off, const char* chars) {
    psim_ctx_impl* c = CTX;
    if (!c || !locus_first_name || !name_off || !chars) return PSIM_EINVAL;
    if (c->code_h.empty()) return fail(c, PSIM_ESTATE, "psim_set_allele_codes must be called first");
    cudaSetDevice(c->cfg.device);
    if (locus_first_name[0] != 0 || name_off[0] != 0) return fail(c, PSIM_EINVAL, "offsets must start at 0");
    const int64_t n_names = locus_first_name[c->L];
    c->aname_max = 0;
    for (int64_t l = 0; l < c->L; l++) {
        int mx = 0;
        const uint8_t* cd = &c->code_h[(size_t)l * c->A];
        for (int a = 0; a < c->A; a++) mx = std::max(mx, (int)cd[a]);
        if (locus_first_name[l + 1] - locus_first_name[l] != mx + 1)
            return fail(c, PSIM_EINVAL, "locus " + std::to_string(l) + " needs one name per allele code");
    }
    for (int64_t k = 0; k < n_names; k++) {
        if (name_off[k + 1] < name_off[k]) return fail(c, PSIM_EINVAL, "name_off must not decrease");
        c->aname_max = std::max(c->aname_max, name_off[k + 1] - name_off[k]);
    }
    cudaFree(c->code_off_d); cudaFree(c->aname_off_d); cudaFree(c->anames_d);
    c->code_off_d = nullptr; c->aname_off_d = nullptr; c->anames_d = nullptr;
    CUDA_TRY(dev_alloc(&c->code_off_d, (size_t)c->L + 1));
    CUDA_TRY(dev_alloc(&c->aname_off_d, (size_t)n_names + 1));
    CUDA_TRY(dev_alloc(&c->anames_d, (size_t)name_off[n_names]));
    CUDA_TRY(cudaMemcpy(c->code_off_d, locus_first_name, ((size_t)c->L + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->aname_off_d, name_off, ((size_t)n_names + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->anames_d, chars, (size_t)name_off[n_names], cudaMemcpyHostToDevice));
    return PSIM_OK;
}

int psim_emit_text(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int64_t locus_first, int64_t locus_count,
                   psim_text_targets* t) {
    psim_ctx_impl* c = CTX;
    if (!c || !t) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (int rc = emit_prepare(c, ind_first, ind_count, locus_first, locus_count)) return rc;
    if (!t->founder && !t->allele && !t->dose) return fail(c, PSIM_EINVAL, "no output table requested");
    if (!c->label_off_d) return fail(c, PSIM_ESTATE, "psim_set_locus_names must be called first");
    if (t->allele && !c->code_d) return fail(c, PSIM_ESTATE, "allele table requested but no allele codes set");
    if (t->allele && !c->code_off_d) return fail(c, PSIM_ESTATE, "allele table requested but no allele names set");
    const int fb = c->founder_allele_count > 256 ? 2 : 1;
    const int64_t n_cols_ind = (int64_t)c->R * ind_count, n_cols = n_cols_ind * c->P;
    int dose_mode = 0, dose_which = t->dose_which;
    if (c->code_d) {
        dose_mode = 1;
        if (c->match_which != t->dose_which) {
            k_match_codes<<<grid_for(c->L, 256), 256, 0, c->stream>>>(c->L, c->A, t->dose_which, c->code_d, c->match_d,
                                                                       c->code16_d, c->nibmask_d, c->mtab_d);
            c->stats.kernel_launches++;
            CUDA_TRY(cudaGetLastError());
            c->match_which = t->dose_which;
        }
    } else if (dose_which < 0) {
        dose_which = 0x7FFFFFFF;
    }
    c->stats.emit_kernel_launches = 0;
    c->kev_used = 0;
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    auto up = [](int64_t x, int64_t m) { return (x + m - 1) / m * m; };
    struct Tab { char* host; int64_t cap; int64_t* bytes; int64_t pitch; int cell_bytes; int kind; int max_tok; int64_t n_cells; };
    Tab tab[3] = {{t->founder, t->founder_capacity, &t->founder_bytes, up(n_cols * fb, 128), fb, 0, fb == 1 ? 4 : 6, n_cols},
                  {t->allele, t->allele_capacity, &t->allele_bytes, up(n_cols, 128), 1, 1, (int)(1 + c->aname_max), n_cols},
                  {t->dose, t->dose_capacity, &t->dose_bytes, up(n_cols_ind, 128), 1, 0, 3, n_cols_ind}};
    int64_t row_bytes = 0;
    for (Tab& x : tab) { if (x.host) row_bytes += x.pitch; *x.bytes = 0; }
    const int64_t rows = std::max<int64_t>(1, std::min<int64_t>(locus_count, (64LL << 20) / row_bytes));
    if (int rc = ensure_staging(c, (size_t)(rows * row_bytes))) return rc;
    if (int rc = ensure_buf(c, c->b_rowlen, ((size_t)rows + 1) * sizeof(int64_t))) return rc;
    if (int rc = ensure_buf(c, c->b_rowoff, ((size_t)rows + 1) * sizeof(int64_t))) return rc;
    bool too_small = false;
    // Two text buffers: while the lines of one table block cross the PCIe link, the next block is
    // formatted into the other buffer.
    psim_ctx_impl::Buf* text_buf[2] = {&c->b_text, &c->b_text2};
    cudaEvent_t text_out[2] = {c->ev3, c->ev4};
    bool in_flight[2] = {false, false};
    int tb = 0;
    for (int64_t l0 = locus_first; l0 < locus_first + locus_count; l0 += rows) {
        const int64_t l1 = std::min(l0 + rows, locus_first + locus_count), nr = l1 - l0;
        uint8_t* sp[3];
        {
            uint8_t* p = (uint8_t*)c->staging;
            for (int q = 0; q < 3; q++) { sp[q] = nullptr; if (tab[q].host) { sp[q] = p; p += rows * tab[q].pitch; } }
        }
        // (the cells in `staging` are read by the text kernels on the same stream as the emission that rewrites them)
        if (int rc = emit_device(c, ind_first, ind_count, l0, l1, l0, sp[0], tab[0].pitch, fb, sp[1], tab[1].pitch, sp[2], tab[2].pitch,
                                 dose_mode, dose_which)) return rc;
        for (int q = 0; q < 3; q++) {
            if (!tab[q].host) continue;
            TextArgs ta{};
            ta.cells = sp[q]; ta.pitch = tab[q].pitch; ta.cell_bytes = tab[q].cell_bytes;
            ta.n_cells = tab[q].n_cells; ta.row0_locus = l0; ta.n_rows = (int)nr;
            ta.label_off = c->label_off_d; ta.labels = c->labels_d;
            ta.kind = tab[q].kind; ta.code_off = c->code_off_d; ta.name_off = c->aname_off_d; ta.names = c->anames_d;
            ta.max_tok = tab[q].max_tok;
            int chunk = (int)(40960 / (tab[q].cell_bytes + tab[q].max_tok)) / TEXT_THREADS * TEXT_THREADS;
            chunk = std::max(TEXT_THREADS, std::min(chunk, 8192));
            ta.chunk = chunk;
            const size_t smem = (size_t)chunk * (tab[q].cell_bytes + tab[q].max_tok) + 16;
            if (smem > 200 * 1024) return fail(c, PSIM_ECAPACITY, "allele names longer than 700 characters are not supported in text emission");
            ta.row_len = (int64_t*)c->b_rowlen.p;
            const int blocks = (int)std::min<int64_t>(nr, (int64_t)c->sm_count * 8);
            cudaFuncSetAttribute(k_text_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_text_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            CUDA_TRY(cudaMemsetAsync(ta.row_len + nr, 0, sizeof(int64_t), c->stream));
            k_text_rows<false><<<blocks, TEXT_THREADS, smem, c->stream>>>(ta);
            c->stats.kernel_launches++;
            CUDA_TRY(cudaGetLastError());
            if (int rc = exclusive_scan_i64(c, (int64_t*)c->b_rowlen.p, (int64_t*)c->b_rowoff.p, nr + 1)) return rc;
            k_publish<<<1, 32, 0, c->stream>>>(c->scalars_dev, (int64_t*)c->b_rowoff.p + nr, nullptr, nullptr);
            c->stats.kernel_launches++;
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            const int64_t total = c->scalars_h[0];
            if (*tab[q].bytes + total > tab[q].cap) too_small = true;
            if (!too_small) {
                if (in_flight[tb]) CUDA_TRY(cudaEventSynchronize(text_out[tb]));   // this buffer's previous text has left (it may also grow)
                // the text starts at the same offset modulo 256 on both sides of the PCIe copy
                const int64_t head = *tab[q].bytes & 255;
                if (int rc = ensure_buf(c, *text_buf[tb], (size_t)(total + 256))) return rc;
                ta.row_off = (int64_t*)c->b_rowoff.p;
                ta.text = (char*)text_buf[tb]->p + head;
                k_text_rows<true><<<blocks, TEXT_THREADS, smem, c->stream>>>(ta);
                c->stats.kernel_launches++;
                CUDA_TRY(cudaGetLastError());
                CUDA_TRY(cudaEventRecord(c->ev2, c->stream));
                CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev2, 0));
                CUDA_TRY(cudaMemcpyAsync(tab[q].host + *tab[q].bytes, ta.text, (size_t)total, cudaMemcpyDeviceToHost, c->copy_stream));
                CUDA_TRY(cudaEventRecord(text_out[tb], c->copy_stream));
                in_flight[tb] = true;
                tb ^= 1;
            }
            *tab[q].bytes += total;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    float ms_total = 0.f;
    cudaEventElapsedTime(&ms_total, c->ev0, c->ev1);
    c->stats.ms_emit_total = ms_total;
    if (too_small) return fail(c, PSIM_ECAPACITY, "a text buffer is too small (the bytes needed are in the *_bytes fields)");
    return PSIM_OK;
}

int psim_gamete_stats_run(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int32_t chrom, psim_gamete_stats* out) {
    psim_ctx_impl* c = CTX;
    if (!c || !out) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (chrom < 0 || chrom >= c->C) return fail(c, PSIM_EINVAL, "chromosome out of bounds");
    if (int rc = emit_prepare(c, ind_first, ind_count, c->locus_off_h[chrom], std::max<int64_t>(1, c->locus_off_h[chrom + 1] - c->locus_off_h[chrom]))) return rc;
    const int64_t Lc = c->locus_off_h[chrom + 1] - c->locus_off_h[chrom];
    if (Lc <= 0) return fail(c, PSIM_EINVAL, "the chromosome has no loci");
    const int A = c->founder_allele_count, P = c->P;
    const int gp = c->model.unreduced ? P : P / 2, nd = gp + 1;   // homologs per gamete (Main.java:1546-1548)
    if (A <= 0 || A > 4096) return fail(c, PSIM_EINVAL, "gamete statistics need 1..4096 founder alleles");
    const size_t n_ac = (size_t)A * Lc, n_dc = n_ac * nd, n_diff = (size_t)(Lc + 1) * (Lc + 1);
    const size_t n_words = n_ac + n_dc + (size_t)Lc + (out->recomb_count ? n_diff : 0) + 64 + (size_t)(P / 4 + 1);
    if (int rc = ensure_buf(c, c->b_text, n_words * 4)) return rc;
    uint32_t* w = (uint32_t*)c->b_text.p;
    CUDA_TRY(cudaMemsetAsync(w, 0, n_words * 4, c->stream));
    StatsArgs sa{};
    sa.C = c->C; sa.R = c->R; sa.P = P; sa.A = A; sa.N = c->N; sa.gp = gp;
    sa.run_desc = c->run_desc_d; sa.run_lb = c->run_lb_d; sa.run_founder = c->run_founder_d;
    sa.desc = c->desc_d; sa.cfg_quad = c->cfg_quad_d;
    sa.ind_first = ind_first; sa.ind_count = ind_count; sa.chrom = chrom; sa.Lc = (uint32_t)Lc;
    sa.allele_count = w; w += n_ac;
    sa.dose_count = w; w += n_dc;
    sa.homozygous_count = w; w += Lc;
    sa.diff = out->recomb_count ? (int32_t*)w : nullptr; if (out->recomb_count) w += n_diff;
    sa.segments_hist = w; w += 64;
    sa.quadrivalent_hist = w;
    const size_t smem = ((size_t)A + (size_t)A * nd + 1) * 4;
    if (smem > 160 * 1024) return fail(c, PSIM_ECAPACITY, "too many founder alleles for the per-locus histogram");
    cudaFuncSetAttribute(k_stats_locus, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_stats_locus<<<(unsigned)Lc, 256, smem, c->stream>>>(sa);
    k_stats_homolog<<<grid_for(ind_count * c->R * P, 256), 256, 0, c->stream>>>(sa);
    c->stats.kernel_launches += 2;
    if (sa.diff) {
        k_prefix_rows<<<grid_for(Lc + 1, 128), 128, 0, c->stream>>>(sa.diff, (uint32_t)(Lc + 1));
        k_prefix_cols<<<grid_for(Lc + 1, 128), 128, 0, c->stream>>>(sa.diff, (uint32_t)(Lc + 1));
        c->stats.kernel_launches += 2;
    }
    if (c->cfg_quad_d) {
        k_stats_quads<<<grid_for(ind_count * c->R * 2, 256), 256, 0, c->stream>>>(sa);
        c->stats.kernel_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    out->n_gametes = (P / gp) * ind_count * c->R;
    if (out->allele_count) CUDA_TRY(cudaMemcpyAsync(out->allele_count, sa.allele_count, n_ac * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out->dose_count) CUDA_TRY(cudaMemcpyAsync(out->dose_count, sa.dose_count, n_dc * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out->homozygous_count) CUDA_TRY(cudaMemcpyAsync(out->homozygous_count, sa.homozygous_count, (size_t)Lc * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out->recomb_count)   // [Lc][Lc] = the interior of the prefix-summed difference array
        CUDA_TRY(cudaMemcpy2DAsync(out->recomb_count, (size_t)Lc * 4, sa.diff, (size_t)(Lc + 1) * 4, (size_t)Lc * 4, (size_t)Lc,
                                   cudaMemcpyDeviceToHost, c->stream));
    if (out->segments_hist) CUDA_TRY(cudaMemcpyAsync(out->segments_hist, sa.segments_hist, 64 * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out->quadrivalent_hist) CUDA_TRY(cudaMemcpyAsync(out->quadrivalent_hist, sa.quadrivalent_hist, (size_t)(P / 4 + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return PSIM_OK;
}

int psim_all_dose_rows(psim_ctx* ctx, int64_t locus_first, int64_t locus_count, int64_t* row_off) {
    psim_ctx_impl* c = CTX;
    if (!c || !row_off) return PSIM_EINVAL;
    if (c->L == 0) return fail(c, PSIM_ESTATE, "psim_set_map must be called first");
    if (locus_first < 0 || locus_count <= 0 || locus_first + locus_count > c->L) return fail(c, PSIM_EINVAL, "locus range out of bounds");
    int64_t rows = 0;
    for (int64_t k = 0; k < locus_count; k++) {
        row_off[k] = rows;
        if (c->code_h.empty()) rows += c->founder_allele_count;
        else {
            int mx = 0;
            const uint8_t* cd = &c->code_h[(size_t)(locus_first + k) * c->A];
            for (int a = 0; a < c->A; a++) mx = std::max(mx, (int)cd[a]);
            rows += mx + 1;
        }
    }
    row_off[locus_count] = rows;
    return PSIM_OK;
}

int psim_emit_all_dose(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int64_t locus_first, int64_t locus_count,
                       void* out, int64_t pitch, int32_t memory) {
    psim_ctx_impl* c = CTX;
    if (!c || !out) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (int rc = emit_prepare(c, ind_first, ind_count, locus_first, locus_count)) return rc;
    const int64_t n_cols_ind = (int64_t)c->R * ind_count;
    if (pitch < n_cols_ind) return fail(c, PSIM_EINVAL, "pitch too small");
    std::vector<int64_t> row_off(locus_count + 1);
    if (int rc = psim_all_dose_rows(ctx, locus_first, locus_count, row_off.data())) return rc;
    int64_t* row_off_d = nullptr;
    CUDA_TRY(dev_alloc(&row_off_d, (size_t)locus_count + 1));
    CUDA_TRY(cudaMemcpyAsync(row_off_d, row_off.data(), (locus_count + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    c->stats.emit_kernel_launches = 0;
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    AllDoseArgs aa{};
    aa.C = c->C; aa.R = c->R; aa.P = c->P; aa.N = c->N;
    aa.run_desc = c->run_desc_d; aa.run_lb = c->run_lb_d; aa.run_founder = c->run_founder_d; aa.locus_off = c->locus_off_d;
    aa.ind_first = ind_first; aa.ind_count = ind_count; aa.n_cols_ind = n_cols_ind;
    aa.code = c->code_d; aa.A = c->A;
    if (memory == PSIM_MEM_DEVICE) {
        aa.l_begin = locus_first; aa.l_count = locus_count; aa.row_off = row_off_d; aa.row_first = 0;
        aa.out = (uint8_t*)out; aa.pitch = pitch;
        k_emit_all_dose<<<grid_for(locus_count * n_cols_ind, 256), 256, 0, c->stream>>>(aa);
        c->stats.kernel_launches++; c->stats.emit_kernel_launches++;
        CUDA_TRY(cudaGetLastError());
    } else {
        const int64_t sp = (n_cols_ind + 255) / 256 * 256;
        int64_t k0 = 0;
        while (k0 < locus_count) {
            int64_t k1 = k0 + 1;
            while (k1 < locus_count && (row_off[k1 + 1] - row_off[k0]) * sp <= (1LL << 30)) k1++;
            const int64_t nrows = row_off[k1] - row_off[k0];
            if (int rc = ensure_staging(c, (size_t)(nrows * sp))) { cudaFree(row_off_d); return rc; }
            aa.l_begin = locus_first + k0; aa.l_count = k1 - k0; aa.row_off = row_off_d + k0; aa.row_first = row_off[k0];
            aa.out = (uint8_t*)c->staging; aa.pitch = sp;
            k_emit_all_dose<<<grid_for(aa.l_count * n_cols_ind, 256), 256, 0, c->stream>>>(aa);
            c->stats.kernel_launches++; c->stats.emit_kernel_launches++;
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpy2DAsync((uint8_t*)out + row_off[k0] * pitch, pitch, c->staging, sp, n_cols_ind, nrows, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            k0 = k1;
        }
    }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaFree(row_off_d);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->stats.ms_emit_total = ms;
    c->stats.ms_emit_kernel = ms;
    return PSIM_OK;
}

int psim_emit_all_dose_text(psim_ctx* ctx, int64_t ind_first, int64_t ind_count, int64_t locus_first, int64_t locus_count,
                            char* text, int64_t capacity, int64_t* bytes) {
    psim_ctx_impl* c = CTX;
    if (!c || !text || !bytes) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    *bytes = 0;
    if (int rc = emit_prepare(c, ind_first, ind_count, locus_first, locus_count)) return rc;
    if (!c->label_off_d) return fail(c, PSIM_ESTATE, "psim_set_locus_names must be called first");
    if (c->code_d && !c->code_off_d) return fail(c, PSIM_ESTATE, "allele codes are set but no allele names");
    const int64_t n_cols_ind = (int64_t)c->R * ind_count;
    std::vector<int64_t> row_off(locus_count + 1);
    if (int rc = psim_all_dose_rows(ctx, locus_first, locus_count, row_off.data())) return rc;
    int64_t* row_off_d = nullptr;
    CUDA_TRY(dev_alloc(&row_off_d, (size_t)locus_count + 1));
    CUDA_TRY(cudaMemcpyAsync(row_off_d, row_off.data(), (locus_count + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    AllDoseArgs aa{};
    aa.C = c->C; aa.R = c->R; aa.P = c->P; aa.N = c->N;
    aa.run_desc = c->run_desc_d; aa.run_lb = c->run_lb_d; aa.run_founder = c->run_founder_d; aa.locus_off = c->locus_off_d;
    aa.ind_first = ind_first; aa.ind_count = ind_count; aa.n_cols_ind = n_cols_ind;
    aa.code = c->code_d; aa.A = c->A;
    const int64_t sp = (n_cols_ind + 255) / 256 * 256;
    bool too_small = false;
    int rc_out = PSIM_OK;
    int64_t k0 = 0;
    while (k0 < locus_count && rc_out == PSIM_OK) {
        int64_t k1 = k0 + 1;
        while (k1 < locus_count && (row_off[k1 + 1] - row_off[k0]) * sp <= (256LL << 20)) k1++;
        const int64_t nrows = row_off[k1] - row_off[k0];
        if ((rc_out = ensure_staging(c, (size_t)(nrows * sp)))) break;
        aa.l_begin = locus_first + k0; aa.l_count = k1 - k0; aa.row_off = row_off_d + k0; aa.row_first = row_off[k0];
        aa.out = (uint8_t*)c->staging; aa.pitch = sp;
        k_emit_all_dose<<<grid_for(aa.l_count * n_cols_ind, 256), 256, 0, c->stream>>>(aa);
        c->stats.kernel_launches++;
        // row -> (locus, allele) for the labels
        std::vector<int32_t> rl((size_t)nrows * 2);
        for (int64_t k = k0; k < k1; k++)
            for (int64_t r = row_off[k]; r < row_off[k + 1]; r++) {
                rl[(size_t)(r - row_off[k0])] = (int32_t)(k - k0);
                rl[(size_t)(nrows + r - row_off[k0])] = (int32_t)(r - row_off[k]);
            }
        if ((rc_out = ensure_buf(c, c->b_rowlen, ((size_t)nrows + 1) * sizeof(int64_t)))) break;
        if ((rc_out = ensure_buf(c, c->b_rowoff, ((size_t)nrows + 1) * sizeof(int64_t) + (size_t)nrows * 8))) break;
        int32_t* rl_d = (int32_t*)((int64_t*)c->b_rowoff.p + nrows + 1);
        CUDA_TRY(cudaMemcpyAsync(rl_d, rl.data(), (size_t)nrows * 8, cudaMemcpyHostToDevice, c->stream));
        TextArgs ta{};
        ta.cells = (const uint8_t*)c->staging; ta.pitch = sp; ta.cell_bytes = 1;
        ta.n_cells = n_cols_ind; ta.row0_locus = locus_first + k0; ta.n_rows = (int)nrows;
        ta.label_off = c->label_off_d; ta.labels = c->labels_d;
        ta.kind = 0; ta.code_off = c->code_off_d; ta.name_off = c->aname_off_d; ta.names = c->anames_d;
        ta.row_locus = rl_d; ta.row_sub = rl_d + nrows; ta.sub_kind = c->code_d ? 2 : 1;
        ta.max_tok = 3;
        ta.chunk = 8192;
        const size_t smem = (size_t)ta.chunk * (1 + ta.max_tok) + 16;
        ta.row_len = (int64_t*)c->b_rowlen.p;
        const int blocks = (int)std::min<int64_t>(nrows, (int64_t)c->sm_count * 8);
        cudaFuncSetAttribute(k_text_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_text_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        CUDA_TRY(cudaMemsetAsync(ta.row_len + nrows, 0, sizeof(int64_t), c->stream));
        k_text_rows<false><<<blocks, TEXT_THREADS, smem, c->stream>>>(ta);
        c->stats.kernel_launches++;
        CUDA_TRY(cudaGetLastError());
        if ((rc_out = exclusive_scan_i64(c, (int64_t*)c->b_rowlen.p, (int64_t*)c->b_rowoff.p, nrows + 1))) break;
        int64_t total = 0;
        CUDA_TRY(cudaMemcpyAsync(&total, (int64_t*)c->b_rowoff.p + nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));   // also: rl (host vector) has been read
        if (*bytes + total > capacity) too_small = true;
        if (!too_small) {
            if ((rc_out = ensure_buf(c, c->b_text, (size_t)total))) break;
            ta.row_off = (int64_t*)c->b_rowoff.p;
            ta.text = (char*)c->b_text.p;
            k_text_rows<true><<<blocks, TEXT_THREADS, smem, c->stream>>>(ta);
            c->stats.kernel_launches++;
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(text + *bytes, c->b_text.p, (size_t)total, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
        *bytes += total;
        k0 = k1;
    }
    cudaFree(row_off_d);
    if (rc_out != PSIM_OK) return rc_out;
    if (too_small) return fail(c, PSIM_ECAPACITY, "the text buffer is too small (the bytes needed are in *bytes)");
    return PSIM_OK;
}

int psim_reset(psim_ctx* ctx, int64_t seed, uint32_t replicate_first) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    c->cfg.seed = seed;
    c->cfg.replicate_first = replicate_first;
    c->model.seed = (uint64_t)seed;
    CUDA_TRY(cudaMemsetAsync(c->desc_d, 0, c->n_hom * sizeof(uint64_t), c->stream));
    const int64_t ncfg = (int64_t)c->R * c->N * c->C * 2;
    CUDA_TRY(cudaMemsetAsync(c->cfg_quad_d, 0xFF, ncfg, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->cfg_seq_d, 0xFF, ncfg * c->P, c->stream));
    std::fill(c->hs_reps.begin(), c->hs_reps.end(), 0);
    c->slab_used = 0;
    c->stats.n_segments = 0;
    c->runs_valid = false;
    c->fresh = true;
    c->check_reps = false;
    return PSIM_OK;
}

int psim_set_stream(psim_ctx* ctx, uint64_t cuda_stream) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    cudaStreamSynchronize(c->stream);
    c->stream = cuda_stream ? (cudaStream_t)(uintptr_t)cuda_stream : c->own_stream;
    return PSIM_OK;
}

int psim_alloc_host(psim_ctx* ctx, size_t bytes, void** out) {
    psim_ctx_impl* c = CTX;
    if (!c || !out) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    *out = nullptr;
    CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return PSIM_OK;
}

int psim_free_host(psim_ctx* ctx, void* p) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (!p) return PSIM_OK;
    cudaSetDevice(c->cfg.device);
    CUDA_TRY(cudaFreeHost(p));
    return PSIM_OK;
}

int psim_get_stats(psim_ctx* ctx, psim_stats* out) {
    psim_ctx_impl* c = CTX;
    if (!c || !out) return PSIM_EINVAL;
    *out = c->stats;
    return PSIM_OK;
}

int psim_export_haplostructs_device(psim_ctx* ctx, int64_t first, int64_t count, psim_device_hs* out) {
    psim_ctx_impl* c = CTX;
    if (!c || !out) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (c->R != 1) return fail(c, PSIM_EINVAL, "device export supports replicate_count == 1");
    int r;
    if (int rc = check_range(c, c->cfg.replicate_first, first, count, r)) return rc;
    int64_t nseg = 0;
    if (int rc = gather_to_export(c, r, first, count, nseg)) return rc;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    out->n_homologs = count * c->C * c->P;
    out->n_segments = nseg;
    out->seg_off = c->exp_off_d;
    out->founder = c->exp_founder_d;
    out->pos = c->exp_pos_d;
    return PSIM_OK;
}

int psim_import_haplostructs_device(psim_ctx* ctx, int64_t first, int64_t count, const psim_device_hs* in) {
    psim_ctx_impl* c = CTX;
    if (!c || !in) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
    if (int rc = settle(c)) return rc;
    if (c->R != 1) return fail(c, PSIM_EINVAL, "device import supports replicate_count == 1");
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    if (first < 0 || count <= 0 || first + count > c->N) return fail(c, PSIM_EINVAL, "individual range out of bounds");
    const int64_t nh = count * c->C * c->P;
    if (in->n_homologs != nh) return fail(c, PSIM_EINVAL, "homolog count does not match the individual range");
    if (int rc = ensure_slab(c, c->slab_used + in->n_segments)) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->seg_pos_d + c->slab_used, in->pos, in->n_segments * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->seg_founder_d + c->slab_used, in->founder, in->n_segments * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    k_scatter_desc<<<grid_for(nh, 256), 256, 0, c->stream>>>(nh, first, 0, c->C, c->R, c->N, c->P, in->seg_off, c->slab_used, c->desc_d);
    c->scalars_h[0] = 0;
    k_check_founders<<<grid_for(in->n_segments, 256), 256, 0, c->stream>>>(in->founder, in->n_segments, c->founder_allele_count, c->scalars_dev);
    c->stats.kernel_launches += 2;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->scalars_h[0] != 0) return fail(c, PSIM_EINVAL, "invalid founder allele '" + std::to_string(c->scalars_h[0] - 1) + "'");
    c->slab_used += in->n_segments;
    c->stats.n_segments = c->slab_used;
    for (int64_t i = first; i < first + count; i++) c->hs_reps[i] = c->R;
    c->runs_valid = false;
    c->fresh = false;
    return PSIM_OK;
}

}  // extern "C"

#include "psim_shard.inl"
