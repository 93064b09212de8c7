// libpsim — B200 (sm_100a) meiosis + genotype emission behind the C ABI of include/psim.h.
//
// Device layout (DESIGN.md "Data layout in HBM"):
//   homolog id  hid = ((c * R + r) * N + i) * P + h      chromosome-major, so that the columns
//                                                         of one emission tile are contiguous
//   hom_desc[hid]  = begin (40 bit) | count << 40         into the segment slab
//   seg_pos[] f64 Morgan / seg_founder[] i32              HaploStruct.recombPos / founder, unmerged
//   run_desc / run_lb[] u32 / run_founder[] u16           K0 output: per homolog the runs of loci
//                                                         that carry one founder (locus-index space)
// Kernels: k_init_founders, k_meiosis_config + k_meiosis_mv (K1a), k_meiosis_assemble (K1b), k_index_runs (K0),
//          k_emit_tiles<P,G> (K2, HBM-write-bound), k_emit_generic, gather/scatter helpers.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "../../include/psim.h"
#include "psim_meiosis.cuh"

namespace psim {

constexpr int DESC_SHIFT = 40;
constexpr uint64_t DESC_MASK = (1ULL << DESC_SHIFT) - 1;
__host__ __device__ __forceinline__ uint64_t make_desc(uint64_t begin, uint64_t count) { return begin | (count << DESC_SHIFT); }
__host__ __device__ __forceinline__ uint64_t desc_begin(uint64_t d) { return d & DESC_MASK; }
__host__ __device__ __forceinline__ uint32_t desc_count(uint64_t d) { return (uint32_t)(d >> DESC_SHIFT); }

// ------------------------------------------------------------------------------------------
// Founders: Individual.calcFounderHaplostruct (Individual.java:155-171)
// job j = (founder slot f, replicate r): individual ind[f], first allele allele0[f]
__global__ void k_init_founders(int n_jobs, int n_new, const int32_t* __restrict__ ind, const int32_t* __restrict__ allele0,
                                int C, int R, int64_t N, int P, const ChromParams* __restrict__ chrom,
                                uint64_t* __restrict__ desc, double* __restrict__ seg_pos, int32_t* __restrict__ seg_founder,
                                int64_t slab_base) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)n_jobs * C * P;
    if (t >= total) return;
    const int h = (int)(t % P);
    const int c = (int)((t / P) % C);
    const int job = (int)(t / ((int64_t)P * C));
    const int f = job % n_new, r = job / n_new;
    const int64_t hid = (((int64_t)c * R + r) * N + ind[f]) * P + h;
    const int64_t s = slab_base + t;
    seg_pos[s] = chrom[c].head;
    seg_founder[s] = allele0[f] + h;
    desc[hid] = make_desc((uint64_t)s, 1);
}

// ------------------------------------------------------------------------------------------
// K1a — one thread per (child, chromosome, parent): meiosis plan for the transmitted gamete.
// Output per child homolog u (scan order ((c*R + r)*n_lev + j)*P + h): the pieces
// (start position, parental homolog) that make up the chromatid, and its segment count.
struct PlanArgs {
    ModelParams model;
    const ChromParams* chrom;
    int C, R, P;
    int64_t N;
    uint32_t replicate_first;
    int n_lev;                   // individuals of this batch
    const int32_t* lev_ind;      // their indices
    const int32_t* par0;
    const int32_t* par1;
    const uint64_t* desc;
    const double* seg_pos;
    const int32_t* seg_founder;
    int piece_cap;
    int32_t* plan_np;            // [n_units_homologs]
    int64_t* plan_cnt;           // [n_units_homologs] -> segment counts (scanned in place later)
    double* plan_pos;            // [n_units_homologs][piece_cap]
    uint8_t* plan_hom;           // [n_units_homologs][piece_cap]
    int8_t* cfg_quad;            // [R*N][C][2]
    int8_t* cfg_seq;             // [R*N][C][2][P]
    int* error_flag;
    int64_t* task_bi;            // bivalents of the batch: (thread id of the unit) * 8 + multivalent number
    int64_t* task_quad;          // quadrivalents
    int* task_count;             // [0] bivalents, [1] quadrivalents
};

// K1a is split in two so that the warps of the heavy part run one code path:
//   k_meiosis_config  one thread per (child, chromosome, parent): pairing configuration only
//                     (Individual.calcChromConfig); it also files every multivalent of the unit as a
//                     task in the bivalent or the quadrivalent list (one atomic per warp and list)
//   k_meiosis_mv<Q>   one thread per task of one kind: crossing-over, centromere order, choice of
//                     the transmitted chromatids and their trace into the plan
// Every multivalent has its own Philox stream (stage 1 + v), so the split changes no draw.
__device__ __forceinline__ void plan_decode(const PlanArgs& a, int64_t t, int& par, int& j, int& r, int& c) {
    par = (int)(t & 1);
    const int64_t unit = t >> 1;
    j = (int)(unit % a.n_lev);
    r = (int)((unit / a.n_lev) % a.R);
    c = (int)(unit / ((int64_t)a.n_lev * a.R));
}

__global__ void __launch_bounds__(128) k_meiosis_config(PlanArgs a) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)a.C * a.R * a.n_lev * 2;
    int n_quad = 0, n_bi = 0;
    if (t < total) {
        int par, j, r, c;
        plan_decode(a, t, par, j, r, c);
        const int P = a.P;
        const int32_t child = a.lev_ind[j];
        const int32_t parent = par == 0 ? a.par0[child] : a.par1[child];
        const ChromParams ch = a.chrom[c];
        HomologView hv[MAXP];
        const int64_t phid0 = (((int64_t)c * a.R + r) * a.N + parent) * P;
        for (int h = 0; h < P; h++) {
            const uint64_t d = a.desc[phid0 + h];
            hv[h].pos = a.seg_pos + desc_begin(d);
            hv[h].founder = a.seg_founder + desc_begin(d);
            hv[h].n = (int)desc_count(d);
        }
        PhiloxStream rng;
        rng.init(a.model.seed, a.replicate_first + (uint32_t)r);
        PairingState ps;
        rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, STAGE_CONFIG);
        calc_chrom_config(ps, rng, a.model, ch, hv);
        const int64_t q = (((int64_t)r * a.N + child) * a.C + c) * 2 + par;
        a.cfg_quad[q] = (int8_t)ps.quad;
        for (int h = 0; h < P; h++) a.cfg_seq[q * P + h] = ps.chromseq[h];
        n_quad = ps.quad;
        n_bi = (P - 4 * ps.quad) / 2;
    }
    // file the tasks: exclusive prefix inside the warp, one atomic per warp and list
    const int lane = threadIdx.x & 31;
    int sq = n_quad, sb = n_bi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int yq = __shfl_up_sync(0xFFFFFFFFu, sq, o), yb = __shfl_up_sync(0xFFFFFFFFu, sb, o);
        if (lane >= o) { sq += yq; sb += yb; }
    }
    int base_q = 0, base_b = 0;
    if (lane == 31) {
        if (sq) base_q = atomicAdd(a.task_count + 1, sq);
        if (sb) base_b = atomicAdd(a.task_count, sb);
    }
    base_q = __shfl_sync(0xFFFFFFFFu, base_q, 31) + sq - n_quad;
    base_b = __shfl_sync(0xFFFFFFFFu, base_b, 31) + sb - n_bi;
    for (int v = 0; v < n_quad; v++) a.task_quad[base_q + v] = t * 8 + v;
    for (int v = 0; v < n_bi; v++) a.task_bi[base_b + v] = t * 8 + n_quad + v;
}

template <bool QUAD>
__global__ void __launch_bounds__(128) k_meiosis_mv(PlanArgs a) {
    const int64_t ti = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (ti >= (int64_t)a.task_count[QUAD ? 1 : 0]) return;
    const int64_t code = (QUAD ? a.task_quad : a.task_bi)[ti];
    const int v = (int)(code & 7);
    int par, j, r, c;
    plan_decode(a, code >> 3, par, j, r, c);
    const int P = a.P, p2 = P / 2;
    const int32_t child = a.lev_ind[j];
    const int32_t parent = par == 0 ? a.par0[child] : a.par1[child];
    const ChromParams ch = a.chrom[c];
    const ModelParams& m = a.model;
    const int64_t phid0 = (((int64_t)c * a.R + r) * a.N + parent) * P;
    const int64_t q = (((int64_t)r * a.N + child) * a.C + c) * 2 + par;
    const int quad = a.cfg_quad[q];
    const int8_t* chromseq = a.cfg_seq + q * P;
    PhiloxStream rng;
    rng.init(m.seed, a.replicate_first + (uint32_t)r);
    // Each transmitted chromatid is traced straight into the plan entry of the child homolog it
    // will occupy after the gamete's homolog shuffle. The shuffle draws come from their own stream:
    // slot h of the transmitted gamete receives pre-shuffle slot shuf[h] (Individual.java:325-335).
    const int64_t u0 = (((int64_t)c * a.R + r) * a.n_lev + j) * P + (int64_t)par * p2;
    int8_t shuf[MAXP / 2];
    for (int s = 0; s < p2; s++) shuf[s] = (int8_t)s;
    rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, STAGE_SHUFFLE);
    shuffle_small(rng, shuf, p2);
    int8_t dest_of_slot[MAXP / 2];  // pre-shuffle slot -> final homolog position
    for (int h = 0; h < p2; h++) dest_of_slot[shuf[h]] = (int8_t)h;

    const int firstbi = quad * 4;
    constexpr int k = QUAD ? 4 : 2;
    const int base = QUAD ? 4 * v : firstbi + 2 * (v - quad);
    Chiasmata x;
    x.overflow = false;
    rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, 1u + (uint32_t)v);
    double exchange_lim1 = 0.0;
    if (QUAD) exchange_lim1 = quadrivalent_crossing_over(x, rng, m, ch);
    else parallel_crossing_over(x, rng, m, ch, 2);
    if (x.overflow) { atomicExch(a.error_flag, PSIM_ECAPACITY); return; }
    int8_t centro[8];
    if (QUAD) quadrivalent_centromere_order(rng, m, ch, exchange_lim1, centro);
    else multivalent_centromere_order(rng, 2, centro);
    int start_slots[2];
    select_gamete0(x, rng, k, ch, centro, start_slots);
    constexpr int n_out = QUAD ? 2 : 1;
    for (int o = 0; o < n_out; o++) {
        // Individual.java:297-321: quadrivalent v fills gamete slots 2v, 2v+1; bivalent v' fills firstbi/2 + v'
        const int pre_slot = QUAD ? 2 * v + o : firstbi / 2 + (v - quad);
        const int64_t u = u0 + dest_of_slot[pre_slot];
        double* ppos = a.plan_pos + u * a.piece_cap;
        uint8_t* phom = a.plan_hom + u * a.piece_cap;
        // Multivalent.slotsToPatterns (:332-357) for the selected chromatid
        int cur = start_slots[o];
        int np = 0;
        ppos[0] = ch.head;
        phom[0] = (uint8_t)chromseq[base + (cur >> 1)];
        np = 1;
        bool over = false;
        for (int e = 0; e < x.n; e++) {
            int nxt = -1;
            if (x.a[e] == cur) nxt = x.b[e];
            else if (x.b[e] == cur) nxt = x.a[e];
            if (nxt >= 0) {
                if (np >= a.piece_cap) { over = true; break; }
                ppos[np] = x.pos[e];
                phom[np] = (uint8_t)chromseq[base + (nxt >> 1)];
                np++;
                cur = nxt;
            }
        }
        if (over) { atomicExch(a.error_flag, PSIM_ECAPACITY); return; }
        a.plan_np[u] = np;
        // Multivalent.fillGamete (:539-577): number of segments the chromatid will have
        auto view = [&](int hom) {
            const uint64_t d = a.desc[phid0 + hom];
            HomologView hvw;
            hvw.pos = a.seg_pos + desc_begin(d);
            hvw.founder = a.seg_founder + desc_begin(d);
            hvw.n = (int)desc_count(d);
            return hvw;
        };
        int64_t cnt = 0;
        if (np == 1) {
            cnt = view(phom[0]).n;
        } else {
            for (int qq = 0; qq < np; qq++) {
                const HomologView h = view(phom[qq]);
                const double patstart = ppos[qq];
                const double patend = qq < np - 1 ? ppos[qq + 1] : ch.tail;
                int seg = h.find_segment(patstart) + 1;
                cnt++;
                while (seg < h.n && h.pos[seg] < patend) { cnt++; seg++; }
            }
        }
        a.plan_cnt[u] = cnt;
    }
}

// K1b — one thread per child homolog: Multivalent.fillGamete (:539-577) + Gamete.fertilization
// (Gamete.java:95-116: homologs 0..P/2-1 from parent 0's gamete, P/2..P-1 from parent 1's).
struct AssembleArgs {
    const ChromParams* chrom;
    int C, R, P;
    int64_t N;
    int n_lev;
    const int32_t* lev_ind;
    const int32_t* par0;
    const int32_t* par1;
    uint64_t* desc;
    double* seg_pos;
    int32_t* seg_founder;
    int piece_cap;
    const int32_t* plan_np;
    const int64_t* plan_cnt;   // per-homolog counts
    const int64_t* plan_off;   // exclusive scan of plan_cnt
    const double* plan_pos;
    const uint8_t* plan_hom;
    int64_t slab_base;
};

__global__ void __launch_bounds__(128) k_meiosis_assemble(AssembleArgs a) {
    const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)a.C * a.R * a.n_lev * a.P;
    if (u >= total) return;
    const int P = a.P, p2 = P / 2;
    const int h = (int)(u % P);
    const int j = (int)((u / P) % a.n_lev);
    const int r = (int)((u / ((int64_t)P * a.n_lev)) % a.R);
    const int c = (int)(u / ((int64_t)P * a.n_lev * a.R));
    const int32_t child = a.lev_ind[j];
    const int32_t parent = h < p2 ? a.par0[child] : a.par1[child];
    const int64_t phid0 = (((int64_t)c * a.R + r) * a.N + parent) * P;
    const double tail = a.chrom[c].tail;
    const int np = a.plan_np[u];
    const double* ppos = a.plan_pos + u * a.piece_cap;
    const uint8_t* phom = a.plan_hom + u * a.piece_cap;
    int64_t w = a.slab_base + a.plan_off[u];
    const int64_t w0 = w;
    if (np == 1) {
        const uint64_t d = a.desc[phid0 + phom[0]];
        const int64_t b = (int64_t)desc_begin(d);
        const int n = (int)desc_count(d);
        for (int s = 0; s < n; s++) { a.seg_pos[w] = a.seg_pos[b + s]; a.seg_founder[w] = a.seg_founder[b + s]; w++; }
    } else {
        for (int q = 0; q < np; q++) {
            const uint64_t d = a.desc[phid0 + phom[q]];
            HomologView hv;
            hv.pos = a.seg_pos + desc_begin(d);
            hv.founder = a.seg_founder + desc_begin(d);
            hv.n = (int)desc_count(d);
            const double patstart = ppos[q];
            const double patend = q < np - 1 ? ppos[q + 1] : tail;
            int seg = hv.find_segment(patstart);
            a.seg_pos[w] = patstart; a.seg_founder[w] = hv.founder[seg]; w++;
            seg++;
            while (seg < hv.n && hv.pos[seg] < patend) { a.seg_pos[w] = hv.pos[seg]; a.seg_founder[w] = hv.founder[seg]; w++; seg++; }
        }
    }
    const int64_t hid = (((int64_t)c * a.R + r) * a.N + child) * P + h;
    a.desc[hid] = make_desc((uint64_t)w0, (uint64_t)(w - w0));
}

// ------------------------------------------------------------------------------------------
// K0 — breakpoints (fp64 Morgan) -> runs in locus-index space, so that emission is integer-only.
// Locus l belongs to the last segment s with pos[s] <= locus position (HaploStruct.findSegment
// :107-112, segment 0 catches everything below); with lb[s] = #loci strictly left of pos[s] that is
// "last s with lb[s] <= l" (ties: a locus exactly on a breakpoint belongs to the segment starting
// there). A suffix minimum makes lb monotone without changing that answer, empty runs are
// dropped and neighbours with the same founder merged.
__global__ void __launch_bounds__(128) k_index_runs(int64_t n_hom, int64_t hom_per_chrom, const uint64_t* __restrict__ desc,
                                                    const double* __restrict__ seg_pos, const int32_t* __restrict__ seg_founder,
                                                    const int64_t* __restrict__ locus_off, const double* __restrict__ locus_pos,
                                                    uint64_t* __restrict__ run_desc, uint32_t* __restrict__ run_lb,
                                                    uint16_t* __restrict__ run_founder) {
    const int64_t hid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (hid >= n_hom) return;
    const uint64_t d = desc[hid];
    const int64_t b = (int64_t)desc_begin(d);
    const int n = (int)desc_count(d);
    if (n == 0) { run_desc[hid] = 0; return; }
    const int c = (int)(hid / hom_per_chrom);
    const double* lp = locus_pos + locus_off[c];
    const uint32_t Lc = (uint32_t)(locus_off[c + 1] - locus_off[c]);
    // pass 1: lb[s]
    for (int s = 0; s < n; s++) {
        uint32_t lo = 0;
        if (s > 0) {
            const double p = seg_pos[b + s];
            uint32_t hi = Lc;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (lp[mid] < p) lo = mid + 1; else hi = mid; }
        }
        run_lb[b + s] = lo;
    }
    // pass 2: suffix minimum
    uint32_t mn = Lc;
    for (int s = n - 1; s >= 0; s--) { const uint32_t v = run_lb[b + s]; mn = v < mn ? v : mn; run_lb[b + s] = mn; }
    // pass 3: compact in place
    int w = 0;
    for (int s = 0; s < n; s++) {
        const uint32_t start = run_lb[b + s];
        const uint32_t end = s + 1 < n ? run_lb[b + s + 1] : Lc;
        if (start >= end) continue;
        const uint16_t f = (uint16_t)seg_founder[b + s];
        if (w > 0 && run_founder[b + w - 1] == f) continue;
        run_lb[b + w] = start;
        run_founder[b + w] = f;
        w++;
    }
    run_desc[hid] = make_desc((uint64_t)b, (uint64_t)w);
}

// ------------------------------------------------------------------------------------------
// K2 — emission. One thread owns G individuals (NC = G*P cells of a locus row, a multiple of 16
// bytes) and walks the tile's loci in map order; the cells only change where one of its homologs
// starts a new run, so the steady state is "re-store the same registers on the next row":
// 128-bit coalesced stores, 32*NC contiguous bytes per warp and row.
// With founder genotypes (allele codes) the per-locus work is kept off the cells:
//   dose   = horizontal nibble sum of (multiplicity-of-each-founder-allele & per-locus nibble mask)
//   allele = PRMT byte lookup of 4 cells at a time in the locus' code table
// both tables are staged per tile in shared memory. MODE 0: no codes, 1: A <= 8, 2: A <= 16,
// 3: any A (per-cell lookup).
struct EmitArgs {
    int C, R, P;
    int64_t N;
    const uint64_t* run_desc;
    const uint32_t* run_lb;
    const uint16_t* run_founder;
    const int64_t* locus_off;     // device, C+1
    int64_t ind_first, ind_count; // emitted individuals per replicate
    int64_t n_cols_ind;           // R * ind_count
    int64_t locus_first;          // first global locus of the whole request (row 0 of the outputs)
    // tiles: for every chromosome touched, [l_begin, l_end) global loci; tile_base prefix over chromosomes
    int n_chunks;                 // chromosomes segments in this launch (<= 64)
    int chunk_chrom[64];
    int64_t chunk_l0[64], chunk_l1[64];
    int chunk_tile_base[65];
    int tile_loci;
    int n_col_tiles;
    int n_tiles;
    int* tile_counter;
    // tile plan written by k_emit_plan, read by k_emit_tiles: one contiguous block per tile
    //   uint4 [0]                     run starts inside the tile per warp (4 counters)
    //   uint4 [1]                     tile coordinates: l0, l1 (chromosome-local rows), column tile, row of l0 in the output
    //   uint4 [2 .. NG*128+1]         founder bytes of the tile's first row: [cg][slot], 16 cells each
    //   uint4 [2 + NG*128 .. ]        the run starts: [warp][EMIT_EV] words row_rel << 22 | lane << 15 | column << 8 | founder
    uint4* plan_blk;
    int plan_blk_size;            // uint4 per tile: 2 + NG * 128 + EMIT_EV
    int* dense_list;              // [1 + n_units] count, then (tile, warp) units with more than EMIT_EV run starts
    uint8_t* founder; int64_t founder_pitch;
    uint8_t* allele;  int64_t allele_pitch;
    uint8_t* dose;    int64_t dose_pitch;
    int dose_which;               // MODE 0: founder allele to count
    const uint32_t* code16;       // [L][4] allele codes of founder alleles 0..15, one byte each (MODE 1, 2)
    const uint32_t* nibmask;      // [L][2] nibble a = 0xF iff founder allele a carries the counted allele (MODE 1, 2)
    const uint4* mtab;            // [L] bytes 0..7: 1 iff founder allele a carries the counted allele; bytes 8..15: the same << 4
    const uint8_t* code;          // [L][A] (MODE 3)
    const uint8_t* match_code;    // [L] code to count (MODE 3)
    int A;
};

#ifndef PSIM_EMIT_MINB
#define PSIM_EMIT_MINB 4
#endif
constexpr int EMIT_THREADS = 128;
constexpr int EMIT_MAX_TILE_LOCI = 64;    // rows of a tile (its per-row dose tables are staged in shared memory)
constexpr int EMIT_EV = 32;           // run starts a warp may have inside a tile on the fast path (one per lane)
constexpr int EMIT_K = 6;             // ... and a single lane (kept sorted in registers)
#ifndef PSIM_EMIT_SYNC_ROWS
#define PSIM_EMIT_SYNC_ROWS 8
#endif
// The four warps of a block re-align every EMIT_SYNC rows so that the 2 KB row pieces of a tile
// reach L2 (and later DRAM) together; without it they drift apart by up to a tile (0 = off).
constexpr int EMIT_SYNC = PSIM_EMIT_SYNC_ROWS;
constexpr uint32_t NO_EVENT = 0xFFFFFFFFu;

struct TileCoord {
    int c, ct;
    int64_t gl0, gl1, c_off;
    uint32_t l0, l1;
};
__device__ __forceinline__ TileCoord decode_tile(const EmitArgs& a, int tile) {
    TileCoord t;
    int k = 0;
    while (tile >= a.chunk_tile_base[k + 1]) k++;
    t.c = a.chunk_chrom[k];
    const int t_in = tile - a.chunk_tile_base[k];
    const int lt = t_in / a.n_col_tiles;
    t.ct = t_in % a.n_col_tiles;
    t.gl0 = a.chunk_l0[k] + (int64_t)lt * a.tile_loci;
    t.gl1 = min(t.gl0 + a.tile_loci, a.chunk_l1[k]);
    t.c_off = a.locus_off[t.c];
    t.l0 = (uint32_t)(t.gl0 - t.c_off);
    t.l1 = (uint32_t)(t.gl1 - t.c_off);
    return t;
}
template <int P, int G>
__device__ __forceinline__ void thread_homologs(const EmitArgs& a, int c, int64_t q0, int64_t (&hid0)[G]) {
#pragma unroll
    for (int g = 0; g < G; g++) {
        const int64_t q = q0 + g;
        if (q < a.n_cols_ind) {
            const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
            hid0[g] = (((int64_t)c * a.R + r) * a.N + i) * P;
        } else hid0[g] = -1;
    }
}

// K2a — tile plan, homolog-major. One thread per (chromosome chunk, emitted column) walks that
// homolog's run list once over the chunk's loci and leaves, for every tile it crosses, the founder
// byte at the tile's first row (coalesced byte stores into the owning thread slot's 16-byte
// vectors) and the run starts inside the tile, appended to the list of the warp that owns the
// column (atomic ticket; the streaming warp sorts its list). Reads cost the streaming kernel
// about 3.4x their byte share of DRAM time (profiles/write_ceiling_r1.txt), so the plan is kept
// as small as it can be: 1 byte per cell and tile plus 4 bytes per run start.
__global__ void __launch_bounds__(256) k_emit_plan(const EmitArgs a, int NC) {
    const int k = blockIdx.y;
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int P = a.P;
    if (col >= a.n_cols_ind * P) return;
    const int c = a.chunk_chrom[k];
    const int64_t q = col / P;
    const int h = (int)(col % P);
    const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
    const int64_t hid = (((int64_t)c * a.R + r) * a.N + i) * P + h;
    const int slot = (int)((col / NC) % EMIT_THREADS);
    const int ct = (int)(col / ((int64_t)NC * EMIT_THREADS));
    const int jcol = (int)(col % NC);
    const int ncg = NC / 16;
    const uint64_t d = __ldg(a.run_desc + hid);
    const uint32_t b = (uint32_t)desc_begin(d), n = desc_count(d);
    const int64_t c_off = a.locus_off[c];
    const uint32_t l_begin = (uint32_t)(a.chunk_l0[k] - c_off), l_end = (uint32_t)(a.chunk_l1[k] - c_off);
    uint32_t lo = 0, hi = n;   // last run with lb <= l_begin
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.run_lb + b + mid) <= l_begin) lo = mid; else hi = mid; }
    uint32_t f = n ? __ldg(a.run_founder + b + lo) : 0;
    uint32_t kx = lo + 1;
    uint32_t nxt = kx < n ? __ldg(a.run_lb + b + kx) : NO_EVENT;
    const int n_lt = (int)((l_end - l_begin + a.tile_loci - 1) / a.tile_loci);
    const uint32_t ev_lo = ((uint32_t)(slot & 31) << 15) | ((uint32_t)jcol << 8);
    for (int lt = 0; lt < n_lt; lt++) {
        const int tile = a.chunk_tile_base[k] + lt * a.n_col_tiles + ct;
        const uint32_t tl0 = l_begin + (uint32_t)lt * a.tile_loci;
        const uint32_t tl1 = min(tl0 + (uint32_t)a.tile_loci, l_end);
        uint4* blk = a.plan_blk + (int64_t)tile * a.plan_blk_size;
        while (nxt <= tl0) {   // a run starting exactly on the tile's first row is its start state
            f = __ldg(a.run_founder + b + kx);
            kx++;
            nxt = kx < n ? __ldg(a.run_lb + b + kx) : NO_EVENT;
        }
        if (slot == 0 && jcol == 0)   // one thread per tile leaves the coordinates the streaming kernel needs
            blk[1] = make_uint4(tl0, tl1, (uint32_t)ct, (uint32_t)(c_off + tl0 - a.locus_first));
        reinterpret_cast<uint8_t*>(blk + 2 + (jcol >> 4) * EMIT_THREADS + slot)[jcol & 15] = (uint8_t)f;
        while (nxt < tl1) {
            f = __ldg(a.run_founder + b + kx);
            uint32_t* cnt = reinterpret_cast<uint32_t*>(blk) + (slot >> 5);
            const uint32_t idx = atomicAdd(cnt, 1u);
            if (idx < EMIT_EV)
                reinterpret_cast<uint32_t*>(blk + 2 + ncg * EMIT_THREADS)[(slot >> 5) * EMIT_EV + idx] = ((nxt - tl0) << 22) | ev_lo | f;
            kx++;
            nxt = kx < n ? __ldg(a.run_lb + b + kx) : NO_EVENT;
        }
    }
}

template <int NW>
__device__ __forceinline__ void store_cells(uint8_t* p, const uint32_t (&v)[NW], bool full, int64_t col0, int64_t n_cols) {
    if (full) {
#pragma unroll
        for (int q = 0; q < NW / 4; q++) reinterpret_cast<uint4*>(p)[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < NW * 4; j++) if (col0 + j < n_cols) p[j] = (uint8_t)(v[j / 4] >> (8 * (j % 4)));
    }
}

// PRMT with the selector taken as is (nibbles 0..7 pick a byte; __byte_perm would first mask it)
__device__ __forceinline__ uint32_t prmt(uint32_t lo, uint32_t hi, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
    return r;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// K2b — streaming. A block's four warps take the same tile (2 KB contiguous per row written at the
// same time), tiles are handed out in row-band order by an atomic ticket requested a tile early;
// the plan block of the next tile (and, with founder genotypes, its per-locus tables) is copied
// into shared memory with cp.async while the current one streams, so no global round trip is ever
// waited for (under a saturated write stream each costs thousands of cycles).
// Run starts: the warp's unsorted list (one per lane) is routed to the owning lanes with six
// ballots and a few shuffles; every lane keeps its own (at most EMIT_K) in registers, sorted by
// row. The row loop then has a fixed trip count and no warp-wide breaks: a lane compares the row
// with its next run start and patches its registers in a short divergent branch.
template <int P, int G, int MODE>
__global__ void __launch_bounds__(EMIT_THREADS, PSIM_EMIT_MINB) k_emit_tiles(const EmitArgs a) {
    constexpr int NC = G * P;
    static_assert(NC % 16 == 0, "a thread's cells must be whole 16-byte vectors");
    constexpr int NW = NC / 4;          // 32-bit words of packed cells
    constexpr int NG = NC / 16;         // 16-byte vectors of packed cells
    constexpr int DW = (G + 3) / 4;     // 32-bit words of packed doses
    constexpr int W = MODE == 2 ? 2 : 1;  // nibble words per individual (8 founder alleles each)
    // A <= 8 and 2 or 4 homologs: the dose is a byte-table lookup of the PRMT selectors (no extra state)
    constexpr bool FASTD = MODE == 1 && (P == 4 || P == 2);
    constexpr int SW = MODE == 0 ? DW : (MODE == 3 || FASTD) ? 1 : G * W;   // words of derived per-thread state
    constexpr int PB = 2 + NG * EMIT_THREADS + EMIT_EV;          // uint4 per plan block
    constexpr int K = EMIT_K;
    __shared__ uint4 s_blk[2][PB];
    __shared__ uint4 s_tab[2][FASTD ? EMIT_MAX_TILE_LOCI + 1 : 1];   // per-row dose tables of the tile (+1: read-ahead)
    __shared__ int s_ticket[2];
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    const bool want_f = a.founder != nullptr, want_a = a.allele != nullptr, want_d = a.dose != nullptr;
    const int64_t n_cols = a.n_cols_ind * P;

    auto prefetch_plan = [&](int tile, int buf) {
        const uint4* src = a.plan_blk + (int64_t)tile * PB;
        for (int i = tid; i < PB; i += EMIT_THREADS) cp_async16(&s_blk[buf][i], src + i);
        if (FASTD && want_d) {   // tile coordinates from the launch arguments: no dependent global load
            int k = 0;
            while (tile >= a.chunk_tile_base[k + 1]) k++;
            const int lt = (tile - a.chunk_tile_base[k]) / a.n_col_tiles;
            const int64_t g0 = a.chunk_l0[k] + (int64_t)lt * a.tile_loci;
            const int64_t g1 = min(g0 + a.tile_loci, a.chunk_l1[k]);
            if (tid < (int)(g1 - g0)) cp_async16(&s_tab[buf][tid], a.mtab + g0 + tid);
        }
        cp_async_commit();
    };
    int ticket = 0;     // thread 0: the outstanding atomic result
    int tile = blockIdx.x, tile_next = 0;
    if (tid == 0) { s_ticket[0] = gridDim.x + atomicAdd(a.tile_counter, 1); ticket = atomicAdd(a.tile_counter, 1); }
    __syncthreads();
    tile_next = s_ticket[0];
    int par = 1, buf = 0;
    if (tile < a.n_tiles) prefetch_plan(tile, 0);
    for (; tile < a.n_tiles;) {
        cp_async_wait_all();
        if (tid == 0) { s_ticket[par] = gridDim.x + ticket; ticket = atomicAdd(a.tile_counter, 1); }
        __syncthreads();   // plan block of `tile` complete and visible; next ticket published
        const int cb = buf;
        buf ^= 1;
        const int tile_cur = tile;
        tile = tile_next;
        tile_next = s_ticket[par];
        par ^= 1;
        if (tile < a.n_tiles) prefetch_plan(tile, buf);

        const uint32_t n_ev = reinterpret_cast<const uint32_t*>(&s_blk[cb][0])[wq];
        // ---- route the warp's run starts to the lanes that own them
        uint32_t q[K];
        bool dense = n_ev > (uint32_t)EMIT_EV;
        if (!dense) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&s_blk[cb][2 + NG * EMIT_THREADS]) + wq * EMIT_EV;
            const uint32_t evr = (uint32_t)lane < n_ev ? src[lane] : NO_EVENT;
            const uint32_t tgt = (evr >> 15) & 31;
            uint32_t hm = __ballot_sync(0xFFFFFFFFu, evr != NO_EVENT);   // lanes holding a run start of mine
#pragma unroll
            for (int b = 0; b < 5; b++) {
                const uint32_t B = __ballot_sync(0xFFFFFFFFu, (tgt >> b) & 1);
                hm &= ((lane >> b) & 1) ? B : ~B;
            }
#pragma unroll
            for (int i = 0; i < K; i++) {
                const uint32_t e = __shfl_sync(0xFFFFFFFFu, evr, (__ffs(hm) - 1) & 31);
                q[i] = hm ? e : NO_EVENT;
                hm &= hm - 1;
            }
            dense = __any_sync(0xFFFFFFFFu, hm != 0);   // a lane with more than K: rare, searched per cell
            // sort by row (the top bits): 12-comparator network for 6 inputs
            auto ce = [&](int x, int y) { const uint32_t lo = min(q[x], q[y]), hi = max(q[x], q[y]); q[x] = lo; q[y] = hi; };
            static_assert(K == 6, "the sorting network below is for 6 entries");
            ce(0, 5); ce(1, 3); ce(2, 4); ce(1, 2); ce(3, 4); ce(0, 3); ce(2, 5); ce(0, 1); ce(2, 3); ce(4, 5); ce(1, 2); ce(3, 4);
        }
        const uint4 tc = s_blk[cb][1];     // tile coordinates left by k_emit_plan
        const uint32_t l0 = tc.x, n_rows = tc.y - tc.x;
        if (dense) {   // sparse map: this (tile, warp) unit goes to k_emit_dense
            if (lane == 0) a.dense_list[1 + atomicAdd(a.dense_list, 1)] = tile_cur * (EMIT_THREADS / 32) + wq;
            if (EMIT_SYNC > 0)   // keep the block's barrier schedule
                for (uint32_t r = EMIT_SYNC > 0 ? EMIT_SYNC - 1 : 0; r < n_rows; r += EMIT_SYNC > 0 ? EMIT_SYNC : 1) __syncthreads();
            continue;
        }
        const int64_t row0 = (int64_t)tc.w;
        const int64_t c_off = row0 + a.locus_first - l0;
        const int64_t q0 = ((int64_t)tc.z * EMIT_THREADS + tid) * G;   // first emitted individual of this lane
        const int64_t colf = q0 * P;                                    // first founder/allele column
        const bool full = colf + NC <= n_cols;   // lanes beyond the last column run along with every store masked off

        // ---- state of this lane at the tile's first row
        uint32_t pk[NW];
#pragma unroll
        for (int cg = 0; cg < NG; cg++) {
            const uint4 v = s_blk[cb][2 + cg * EMIT_THREADS + tid];
            pk[4 * cg] = v.x; pk[4 * cg + 1] = v.y; pk[4 * cg + 2] = v.z; pk[4 * cg + 3] = v.w;
        }
        uint32_t st[SW];           // MODE 0: packed doses; MODE 1/2 (not FASTD): nibble multiplicities [g][w]
        uint32_t sel[NW];          // MODE 1/2: PRMT selectors (low 3 bits of every cell)
        uint32_t him[NW];          // MODE 2: 0xFF per cell whose founder allele is >= 8
#pragma unroll
        for (int w = 0; w < SW; w++) st[w] = 0;
        if (!FASTD) {
#pragma unroll
            for (int j = 0; j < NC; j++) {
                const uint32_t f = (pk[j / 4] >> (8 * (j % 4))) & 0xFF;
                if (MODE == 0) { if (f == (uint32_t)a.dose_which) st[(j / P) / 4] += 1u << (8 * ((j / P) % 4)); }
                else if (MODE == 1) st[j / P] += 1u << (4 * (f & 7));
                else if (MODE == 2) st[(j / P) * W + (f >> 3 & 1)] += 1u << (4 * (f & 7));
            }
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const uint32_t v = pk[w];
                sel[w] = (v & 0x7) | ((v >> 4) & 0x70) | ((v >> 8) & 0x700) | ((v >> 12) & 0x7000);
                him[w] = ((v >> 3) & 0x01010101u) * 0xFFu;
            }
        }

        // one of this lane's homologs starts a new run on row `nxt`: patch the registers
        uint32_t nxt = q[0] >> 22;
        auto take_run_starts = [&](uint32_t r) {
            do {
                const uint32_t e = q[0];
                const uint32_t col = (e >> 8) & 0x7F, f = e & 0xFF;
                const uint32_t cw = col >> 2, sh = (col & 3) * 8, g = col / P;
                uint32_t old = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) {
                    const bool hit = (uint32_t)w == cw;
                    const uint32_t msk = hit ? 0xFFu << sh : 0u;
                    old |= pk[w] & msk;
                    pk[w] = (pk[w] & ~msk) | ((f << sh) & msk);
                    if (MODE == 1 || MODE == 2) {
                        const uint32_t nsh = sh >> 1, nmsk = hit ? 0xFu << nsh : 0u;
                        sel[w] = (sel[w] & ~nmsk) | (((f & 7) << nsh) & nmsk);
                        if (MODE == 2) him[w] = (him[w] & ~msk) | (f & 8 ? msk : 0u);
                    }
                }
                old >>= sh;
                if (MODE == 0) {
                    const uint32_t delta = ((f == (uint32_t)a.dose_which) - (old == (uint32_t)a.dose_which)) << (8 * (g & 3));
#pragma unroll
                    for (int w = 0; w < DW; w++) st[w] += (uint32_t)w == (g >> 2) ? delta : 0u;
                } else if (MODE == 1 && !FASTD) {
                    const uint32_t delta = (1u << (4 * (f & 7))) - (1u << (4 * (old & 7)));
#pragma unroll
                    for (int gg = 0; gg < G; gg++) st[gg] += (uint32_t)gg == g ? delta : 0u;
                } else if (MODE == 2) {
                    const uint32_t dec = 1u << (4 * (old & 7)), inc = 1u << (4 * (f & 7));
                    const uint32_t wi_old = g * W + (old >> 3 & 1), wi_new = g * W + (f >> 3 & 1);
#pragma unroll
                    for (int w = 0; w < SW; w++) st[w] += ((uint32_t)w == wi_new ? inc : 0u) - ((uint32_t)w == wi_old ? dec : 0u);
                }
#pragma unroll
                for (int i = 0; i + 1 < K; i++) q[i] = q[i + 1];
                q[K - 1] = NO_EVENT;
                nxt = q[0] >> 22;
            } while (nxt == r);
        };

        // ---- stream the rows
        uint8_t* const pf0 = a.founder + row0 * a.founder_pitch + colf;
        uint8_t* const pa0 = a.allele + row0 * a.allele_pitch + colf;
        uint8_t* const pd0 = a.dose + row0 * a.dose_pitch + q0;
        const uint32_t fpitch32 = (uint32_t)a.founder_pitch, apitch32 = (uint32_t)a.allele_pitch, dpitch32 = (uint32_t)a.dose_pitch;
        const uint4* code16 = reinterpret_cast<const uint4*>(a.code16) + c_off + l0;    // per-locus tables: warp-uniform
        const uint2* nibmask = reinterpret_cast<const uint2*>(a.nibmask) + c_off + l0;  // loads through L1
        // compiled once per combination of requested tables so that the hot loop carries no
        // predicated-off work (WF / WA / WD shadow the runtime flags)
        // (`full` stays a per-lane runtime flag: the loop holds block barriers, so the lanes of a warp
        // must not split into differently compiled copies of it)
        auto emit_rows_t = [&](auto wf_tag, auto wa_tag, auto wd_tag) {
            const bool FULL = full;
            constexpr bool want_f = decltype(wf_tag)::value, want_a = decltype(wa_tag)::value, want_d = decltype(wd_tag)::value;
            uint4 Tn = s_tab[cb][0];
#pragma unroll 4
            for (uint32_t r = 0; r < n_rows; r++) {
                if (r == nxt) take_run_starts(r);
                uint32_t dose_pk[DW];
                uint8_t* const pf = pf0 + (uint64_t)(r * fpitch32);   // a tile spans far less than 4 GB
                uint8_t* const pa = pa0 + (uint64_t)(r * apitch32);
                uint8_t* const pd = pd0 + (uint64_t)(r * dpitch32);
                if (want_f) store_cells<NW>(pf, pk, FULL, colf, n_cols);
                if (MODE == 0) {
#pragma unroll
                    for (int w = 0; w < DW; w++) dose_pk[w] = st[w];
                } else if (MODE == 1 || MODE == 2) {
                    if (want_a) {
                        const uint4 tb = __ldg(code16 + r);
                        uint32_t ak[NW];
#pragma unroll
                        for (int w = 0; w < NW; w++) {
                            const uint32_t lov = __byte_perm(tb.x, tb.y, sel[w]);
                            if (MODE == 2) { const uint32_t hiv = __byte_perm(tb.z, tb.w, sel[w]); ak[w] = (lov & ~him[w]) | (hiv & him[w]); }
                            else ak[w] = lov;
                        }
                        store_cells<NW>(pa, ak, FULL, colf, n_cols);
                    }
                    if (want_d && FASTD) {
                        // T = {1 per founder allele that carries the counted allele (bytes 0..7), the same << 4}
                        const uint4 T = Tn;
                        Tn = s_tab[cb][r + 1];   // next row's table: its latency hides behind this row's stores
                        if (P == 4) {          // one word = one individual
                            const uint32_t y01 = prmt(T.x, T.y, sel[0]) + prmt(T.z, T.w, sel[1]);
                            const uint32_t y23 = prmt(T.x, T.y, sel[2]) + prmt(T.z, T.w, sel[3]);
                            const uint32_t v = __dp4a(y01, 0x01010101u, 0u) | (__dp4a(y23, 0x01010101u, 0u) << 8);
                            dose_pk[0] = prmt(0x03020100u, 0x07060504u, v);   // dose nibbles (<= 4) to bytes
                        } else {               // P == 2: one word = two individuals
                            uint32_t s2[NW];
#pragma unroll
                            for (int w = 0; w < NW; w++) { const uint32_t x = prmt(T.x, T.y, sel[w]); s2[w] = x + (x >> 8); }
#pragma unroll
                            for (int w = 0; w < DW; w++) dose_pk[w] = prmt(s2[2 * w], s2[2 * w + 1], 0x6420u);
                        }
                    } else if (want_d) {
                        uint2 m;
                        if (W == 2) m = __ldg(nibmask + r);
                        else { m.x = __ldg(reinterpret_cast<const uint32_t*>(nibmask + r)); m.y = 0; }
#pragma unroll
                        for (int w = 0; w < DW; w++) dose_pk[w] = 0;
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            uint32_t d = ((st[(g * W) % SW] & m.x) * 0x11111111u) >> 28;
                            if (W == 2) d += ((st[(g * W + W - 1) % SW] & m.y) * 0x11111111u) >> 28;
                            dose_pk[g / 4] |= d << (8 * (g % 4));
                        }
                    }
                } else {
                    const uint8_t* codes = a.code + (c_off + l0 + r) * a.A;
                    const uint32_t match = a.match_code[c_off + l0 + r];
                    uint32_t ak[NW];
#pragma unroll
                    for (int w = 0; w < DW; w++) dose_pk[w] = 0;
#pragma unroll
                    for (int w = 0; w < NW; w++) ak[w] = 0;
#pragma unroll
                    for (int j = 0; j < NC; j++) {
                        const uint32_t f = (pk[j / 4] >> (8 * (j % 4))) & 0xFF;
                        const uint32_t al = __ldg(codes + f);
                        ak[j / 4] |= al << (8 * (j % 4));
                        if (al == match) dose_pk[(j / P) / 4] += 1u << (8 * ((j / P) % 4));
                    }
                    if (want_a) store_cells<NW>(pa, ak, FULL, colf, n_cols);
                }
                if (want_d) {
                    if (FULL) {
                        if (G % 16 == 0) {
#pragma unroll
                            for (int v = 0; v < G / 16; v++)
                                reinterpret_cast<uint4*>(pd)[v] = make_uint4(dose_pk[(4 * v) % DW], dose_pk[(4 * v + 1) % DW], dose_pk[(4 * v + 2) % DW], dose_pk[(4 * v + 3) % DW]);
                        } else if (G == 8) {
                            *reinterpret_cast<uint2*>(pd) = make_uint2(dose_pk[0], dose_pk[DW - 1]);
                        } else if (G == 4) {
                            *reinterpret_cast<uint32_t*>(pd) = dose_pk[0];
                        } else if (G == 2) {
                            *reinterpret_cast<uint16_t*>(pd) = (uint16_t)dose_pk[0];
                        } else {
#pragma unroll
                            for (int g = 0; g < G; g++) pd[g] = (uint8_t)(dose_pk[g / 4] >> (8 * (g % 4)));
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < G; g++) if (q0 + g < a.n_cols_ind) pd[g] = (uint8_t)(dose_pk[g / 4] >> (8 * (g % 4)));
                    }
                }
                if (EMIT_SYNC > 0 && (r % (EMIT_SYNC > 0 ? EMIT_SYNC : 1)) == EMIT_SYNC - 1) __syncthreads();
            }
        };
        {
            using T = std::true_type;
            using F = std::false_type;
            if (MODE == 0) {   // no allele table without codes
                if (want_f) { if (want_d) emit_rows_t(T{}, F{}, T{}); else emit_rows_t(T{}, F{}, F{}); }
                else emit_rows_t(F{}, F{}, T{});
            } else if (want_f) {
                if (want_a) { if (want_d) emit_rows_t(T{}, T{}, T{}); else emit_rows_t(T{}, T{}, F{}); }
                else { if (want_d) emit_rows_t(T{}, F{}, T{}); else emit_rows_t(T{}, F{}, F{}); }
            } else {
                if (want_a) { if (want_d) emit_rows_t(F{}, T{}, T{}); else emit_rows_t(F{}, T{}, F{}); }
                else emit_rows_t(F{}, F{}, T{});
            }
        }
    }
}

// Dense units (a warp has more than EMIT_EV run starts inside the tile — sparse maps, where the
// tables are tiny): every cell is searched. One warp per unit listed by k_emit_tiles.
template <int P, int G>
__global__ void __launch_bounds__(EMIT_THREADS) k_emit_dense(const EmitArgs a) {
    const int n_dense = a.dense_list[0];
    const bool want_f = a.founder != nullptr, want_a = a.allele != nullptr, want_d = a.dose != nullptr;
    const bool coded = a.code != nullptr;
    for (int di = blockIdx.x; di < n_dense; di += gridDim.x) {
        const int u = a.dense_list[1 + di];
        const TileCoord t = decode_tile(a, u >> 2);
        const int n_rows = (int)(t.l1 - t.l0);
        // work items: (row, lane of the unit, individual of the lane)
        for (int it = threadIdx.x; it < n_rows * 32 * G; it += EMIT_THREADS) {
            const int g = it % G, lane = (it / G) % 32;
            const uint32_t l = t.l0 + it / (G * 32);
            const int64_t q = ((int64_t)t.ct * EMIT_THREADS + (u & 3) * 32 + lane) * G + g;
            if (q >= a.n_cols_ind) continue;
            const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
            const int64_t hid0 = (((int64_t)t.c * a.R + r) * a.N + i) * P;
            const int64_t row = t.gl0 - a.locus_first + (l - t.l0);
            const uint8_t* codes = coded ? a.code + (t.c_off + l) * a.A : nullptr;
            const uint32_t match = coded ? (uint32_t)a.match_code[t.c_off + l] : (uint32_t)a.dose_which;
            uint32_t dose = 0;
#pragma unroll 1
            for (int h = 0; h < P; h++) {
                const uint64_t d = a.run_desc[hid0 + h];
                const uint32_t b = (uint32_t)desc_begin(d), n = desc_count(d);
                uint32_t lo = 0, hi = n;
                while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.run_lb[b + mid] <= l) lo = mid; else hi = mid; }
                const uint32_t f = a.run_founder[b + lo];
                const uint32_t al = coded ? (uint32_t)codes[f] : f;
                if (want_f) a.founder[row * a.founder_pitch + q * P + h] = (uint8_t)f;
                if (want_a) a.allele[row * a.allele_pitch + q * P + h] = (uint8_t)al;
                dose += al == match;
            }
            if (want_d) a.dose[row * a.dose_pitch + q] = (uint8_t)dose;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2 "row image" path — the roofline kernel for founder + dose tables.
// A block owns a band of consecutive locus rows of one chromosome and a segment of at most ~46 KB of
// the founder row. It keeps the CURRENT ROW of that segment in shared memory (1 byte per cell; with
// founder genotypes also a 4-bit copy used as PRMT selectors), patches the few cells whose homolog
// starts a new run on the row, and sends the row to global memory as ONE contiguous TMA bulk store
// (cp.async.bulk shared -> global): every table row leaves the SM as 40 KB + 10 KB of consecutive
// bytes, which is the write pattern HBM likes (tools/write_rowimage.cu, profiles/README.md), and it
// stays that way however the blocks drift against each other. The dose row is recomputed per locus
// from the 4-bit image (A <= 8, ploidy 2 or 4) or patched like the founder row (no genotypes).
// The plan (k_band_plan) holds, per (band, segment) unit, the founder byte of every cell on the
// band's first row and the unordered list of run starts inside the band: 1 byte per cell and BAND,
// bands being ~80 rows, so plan reads are ~1 % of the bytes written.
struct RowsArgs {
    int C, R, P;
    int64_t N;
    const uint64_t* run_desc;
    const uint32_t* run_lb;
    const uint16_t* run_founder;
    int64_t ind_first, ind_count, n_cols_ind;
    int64_t locus_first;          // global locus of output row 0
    int n_chunks;                 // chromosome pieces of this launch (<= 64)
    int chunk_chrom[64];
    int64_t chunk_l0[64], chunk_l1[64], chunk_coff[64];
    int chunk_bands[64];          // bands the piece is cut into (equal row counts, +-1)
    int chunk_unit_base[65];
    int n_seg, seg_cells;         // segments per row; cells per segment (a multiple of 16 * P)
    int n_units, cap;             // units = sum(bands) * n_seg; run starts a unit's list can hold
    uint8_t* plan;                // per unit: [count u32, 12 bytes pad][seg_cells state bytes][cap run starts]
    int64_t plan_stride;
    int* unit_counter;            // ticket
    int* overflow;                // [0] = number of units whose list overflowed, then their ids
    uint8_t* founder; int64_t founder_pitch;
    uint8_t* dose;    int64_t dose_pitch;
    int dose_which;               // without genotypes: founder allele to count
    const uint4* mtab;            // with genotypes: per-locus byte tables of the counted allele (k_match_codes)
};

__device__ __forceinline__ void band_rows(const RowsArgs& a, int k, int j, int64_t& g0, int64_t& g1) {
    const int64_t len = a.chunk_l1[k] - a.chunk_l0[k];
    g0 = a.chunk_l0[k] + len * j / a.chunk_bands[k];
    g1 = a.chunk_l0[k] + len * (j + 1) / a.chunk_bands[k];
}

// One thread per (chromosome piece, emitted cell column): walks the homolog's runs once down the
// piece and leaves the state byte of every band plus the run starts inside it.
__global__ void __launch_bounds__(256) k_band_plan(const RowsArgs a) {
    const int k = blockIdx.y;
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int P = a.P;
    if (col >= a.n_cols_ind * P) return;
    const int c = a.chunk_chrom[k];
    const int64_t q = col / P;
    const int h = (int)(col % P);
    const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
    const int64_t hid = (((int64_t)c * a.R + r) * a.N + i) * P + h;
    const int s = (int)(col / a.seg_cells);
    const uint32_t cs = (uint32_t)(col % a.seg_cells);
    const uint64_t d = __ldg(a.run_desc + hid);
    const uint32_t b = (uint32_t)desc_begin(d), n = desc_count(d);
    const int64_t c_off = a.chunk_coff[k];
    const uint32_t l_begin = (uint32_t)(a.chunk_l0[k] - c_off);
    uint32_t lo = 0, hi = n;   // last run with lb <= l_begin
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.run_lb + b + mid) <= l_begin) lo = mid; else hi = mid; }
    uint32_t f = n ? __ldg(a.run_founder + b + lo) : 0;
    uint32_t kx = lo + 1;
    uint32_t nxt = kx < n ? __ldg(a.run_lb + b + kx) : NO_EVENT;
    for (int j = 0; j < a.chunk_bands[k]; j++) {
        int64_t g0, g1;
        band_rows(a, k, j, g0, g1);
        const uint32_t tl0 = (uint32_t)(g0 - c_off), tl1 = (uint32_t)(g1 - c_off);
        uint8_t* pu = a.plan + (int64_t)(a.chunk_unit_base[k] + j * a.n_seg + s) * a.plan_stride;
        while (nxt <= tl0) {   // a run starting exactly on the band's first row is its start state
            f = __ldg(a.run_founder + b + kx);
            kx++;
            nxt = kx < n ? __ldg(a.run_lb + b + kx) : NO_EVENT;
        }
        pu[16 + cs] = (uint8_t)f;
        while (nxt < tl1) {
            f = __ldg(a.run_founder + b + kx);
            const uint32_t idx = atomicAdd(reinterpret_cast<uint32_t*>(pu), 1u);
            if (idx < (uint32_t)a.cap)
                reinterpret_cast<uint32_t*>(pu + 16 + a.seg_cells)[idx] = ((nxt - tl0) << 24) | (cs << 8) | f;
            kx++;
            nxt = kx < n ? __ldg(a.run_lb + b + kx) : NO_EVENT;
        }
    }
}

constexpr int ROWS_THREADS = 512;
constexpr int ROWS_MAX_BAND = 255;     // the row of a run start inside its band is 8 bits
constexpr int ROWS_SMEM = 220 * 1024;  // one block per SM
#ifndef PSIM_ROWS_NBUF
#define PSIM_ROWS_NBUF 2
#endif
constexpr int ROWS_NBUF = PSIM_ROWS_NBUF;   // row buffers: NBUF-1 rows can be queued at the bulk engine while the next is made

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

// DK: how the dose row is made. 0 = no dose table, 1 = no genotypes (count of one founder allele:
// patched like the founder row), 2 = genotypes, A <= 8, ploidy 4, 3 = genotypes, A <= 8, ploidy 2.
// Founder and dose rows are double-buffered: while the bulk engine streams row r-1 out of one
// buffer, row r is made in the other (copy + patch, or recomputed), so the engine always has a
// row queued and the block never waits for the row it has just issued.
template <int DK>
__global__ void __launch_bounds__(ROWS_THREADS, 1) k_emit_rows(const RowsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int P = a.P;
    const int seg_cells = a.seg_cells, seg_ind = seg_cells / P;
    constexpr int NB = ROWS_NBUF;
    uint8_t* const img0 = smem;                                              // founder byte of every cell on the current row (x NB)
    uint32_t* const nib = reinterpret_cast<uint32_t*>(img0 + (size_t)NB * seg_cells);   // DK >= 2: the same, 4 bits per cell
    uint8_t* const dose0 = reinterpret_cast<uint8_t*>(nib) + (DK >= 2 ? seg_cells / 2 : 0);
    uint32_t* const evs = reinterpret_cast<uint32_t*>(dose0 + (DK >= 1 ? (size_t)NB * seg_ind : 0));   // the unit's run starts sorted by row
    uint32_t* const start = evs + a.cap;                                     // [band rows + 1] first run start of each row
    uint32_t* const cursor = start + ROWS_MAX_BAND + 2;
    __shared__ int s_unit;
    const int64_t n_cols = a.n_cols_ind * P;
    const bool count_ok = a.dose_which >= 0 && a.dose_which < 256;           // DK 1: a founder allele that exists

    for (;;) {
        __syncthreads();
        if (tid == 0) s_unit = atomicAdd(a.unit_counter, 1);
        __syncthreads();
        const int u = s_unit;
        if (u >= a.n_units) break;
        int k = 0;
        while (u >= a.chunk_unit_base[k + 1]) k++;
        const int j = (u - a.chunk_unit_base[k]) / a.n_seg, s = (u - a.chunk_unit_base[k]) % a.n_seg;
        int64_t g0, g1;
        band_rows(a, k, j, g0, g1);
        const int n_rows = (int)(g1 - g0);
        const int64_t out_row0 = g0 - a.locus_first;
        const int64_t c0 = (int64_t)s * seg_cells;                            // first cell column of the segment
        const int n_cell = (int)min((int64_t)seg_cells, n_cols - c0);         // a multiple of P
        const int n_ind = n_cell / P;
        const uint8_t* pu = a.plan + (int64_t)u * a.plan_stride;
        const uint32_t n_ev = *reinterpret_cast<const uint32_t*>(pu);
        if (n_ev > (uint32_t)a.cap) {   // sparse map: searched per cell afterwards (k_emit_generic)
            if (tid == 0) a.overflow[1 + atomicAdd(a.overflow, 1)] = u;
            continue;
        }
        // the bulk engine must be done reading the previous unit's rows out of shared memory
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();

        // ---- start state of the band, run starts bucketed by row
        for (int i = tid * 16; i < seg_cells; i += ROWS_THREADS * 16)
            *reinterpret_cast<uint4*>(img0 + i) = __ldg(reinterpret_cast<const uint4*>(pu + 16 + i));
        for (int i = tid; i <= n_rows; i += ROWS_THREADS) start[i] = 0;
        __syncthreads();
        const uint32_t* ev_g = reinterpret_cast<const uint32_t*>(pu + 16 + seg_cells);
        for (uint32_t e = tid; e < n_ev; e += ROWS_THREADS) atomicAdd(&start[__ldg(ev_g + e) >> 24], 1u);
        __syncthreads();
        if (tid < 32) {   // exclusive scan of the row histogram (<= 256 bins) by one warp
            uint32_t carry = 0;
            for (int base = 0; base <= n_rows; base += 32) {
                const int i = base + tid;
                const uint32_t v = i <= n_rows ? start[i] : 0;
                uint32_t x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (tid >= o) x += y; }
                if (i <= n_rows) { start[i] = carry + x - v; cursor[i] = carry + x - v; }
                carry += __shfl_sync(0xFFFFFFFFu, x, 31);
            }
        }
        __syncthreads();
        for (uint32_t e = tid; e < n_ev; e += ROWS_THREADS) {
            const uint32_t w = __ldg(ev_g + e);
            evs[atomicAdd(&cursor[w >> 24], 1u)] = w;
        }
        if (DK >= 2) {   // 4-bit image: low 3 bits of every cell, the PRMT selectors of the dose lookup
            for (int i = tid; i < seg_cells / 16; i += ROWS_THREADS) {
                const uint4 v = reinterpret_cast<const uint4*>(img0)[i];
                auto pack = [](uint32_t x) { return (x & 0x7) | ((x >> 4) & 0x70) | ((x >> 8) & 0x700) | ((x >> 12) & 0x7000); };
                reinterpret_cast<uint2*>(nib)[i] = make_uint2(pack(v.x) | (pack(v.y) << 16), pack(v.z) | (pack(v.w) << 16));
            }
        }
        if (DK == 1) {   // dose of one founder allele on the band's first row
            for (int i = tid; i < seg_ind; i += ROWS_THREADS) {
                int dsum = 0;
                for (int h = 0; h < P; h++) dsum += img0[i * P + h] == (uint8_t)a.dose_which;
                dose0[i] = (uint8_t)(count_ok ? dsum : 0);
            }
        }
        uint4 Tn = make_uint4(0, 0, 0, 0);
        if (DK >= 2) Tn = __ldg(a.mtab + g0);
        __syncthreads();
        // the other buffers start from the same state (they become rows 1 .. NB-1)
        for (int b = 1; b < NB; b++) {
            for (int i = tid; i < seg_cells / 16; i += ROWS_THREADS)
                reinterpret_cast<uint4*>(img0 + (size_t)b * seg_cells)[i] = reinterpret_cast<const uint4*>(img0)[i];
            if (DK == 1)
                for (int i = tid; i < seg_ind / 16; i += ROWS_THREADS)
                    reinterpret_cast<uint4*>(dose0 + (size_t)b * seg_ind)[i] = reinterpret_cast<const uint4*>(dose0)[i];
        }

        // ---- the rows. Buffer r % NB holds row r-NB when row r begins: it takes the run starts of rows r-NB+1 .. r.
        auto patch = [&](uint8_t* img, uint8_t* drow, uint32_t e0, uint32_t e1) {
            for (uint32_t e = e0 + tid; e < e1; e += ROWS_THREADS) {
                const uint32_t w = evs[e], cs = (w >> 8) & 0xFFFF, f = w & 0xFF;
                if (DK == 1) {
                    const int delta = (int)(f == (uint32_t)a.dose_which) - (int)(count_ok && img[cs] == (uint8_t)a.dose_which);
                    const uint32_t ind = cs / P;
                    if (delta) atomicAdd(reinterpret_cast<uint32_t*>(drow) + (ind >> 2), (uint32_t)delta << (8 * (ind & 3)));
                }
                img[cs] = (uint8_t)f;
            }
        };
        for (int r = 0; r < n_rows; r++) {
            const uint32_t e0 = start[r], e1 = start[r + 1];   // this row's run starts (row 0 has none: they are the state)
            uint8_t* const img = img0 + (size_t)(r % NB) * seg_cells;
            uint8_t* const drow = dose0 + (size_t)(r % NB) * seg_ind;
            // buffers of row r were last read by the bulk stores of row r-NB
            if (r >= NB && tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NB - 1) : "memory");
            if (DK >= 2) {
                for (uint32_t e = e0 + tid; e < e1; e += ROWS_THREADS) {
                    const uint32_t w = evs[e], cs = (w >> 8) & 0xFFFF, sh = (cs & 7) * 4;
                    atomicAnd(&nib[cs >> 3], ~(0xFu << sh));
                    atomicOr(&nib[cs >> 3], (w & 7) << sh);
                }
            }
            __syncthreads();
            // catch up with the rows this buffer missed, oldest first (a cell may change on two of them)
            for (int m = NB - 1; m >= 1; m--) {
                if (r - m >= 1) patch(img, drow, start[r - m], start[r - m + 1]);
                if (m > 1) __syncthreads();
            }
            if (DK >= 2) {
                const uint4 T = Tn;
                if (r + 1 < n_rows) Tn = __ldg(a.mtab + g0 + r + 1);
                if (DK == 2) {   // 8 individuals per step: 4 words of the 4-bit image in, 8 dose bytes out
                    for (int i = tid; i < seg_ind / 8; i += ROWS_THREADS) {
                        const uint4 v = reinterpret_cast<const uint4*>(nib)[i];
                        auto two = [&](uint32_t w) { return __dp4a(prmt(T.x, T.y, w) + prmt(T.z, T.w, w >> 16), 0x01010101u, 0u); };
                        const uint32_t lo = two(v.x) | (two(v.y) << 8), hi = two(v.z) | (two(v.w) << 8);   // dose nibbles
                        reinterpret_cast<uint2*>(drow)[i] = make_uint2(prmt(0x03020100u, 0x07060504u, lo), prmt(0x03020100u, 0x07060504u, hi));
                    }
                } else {         // ploidy 2: 16 individuals per step
                    for (int i = tid; i < seg_ind / 16; i += ROWS_THREADS) {
                        const uint4 v = reinterpret_cast<const uint4*>(nib)[i];
                        auto four = [&](uint32_t w) {
                            const uint32_t xl = prmt(T.x, T.y, w), xh = prmt(T.x, T.y, w >> 16);
                            return prmt(xl + (xl >> 8), xh + (xh >> 8), 0x6420u);
                        };
                        reinterpret_cast<uint4*>(drow)[i] = make_uint4(four(v.x), four(v.y), four(v.z), four(v.w));
                    }
                }
            }
            __syncthreads();
            patch(img, drow, e0, e1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            uint8_t* const gf = a.founder ? a.founder + (out_row0 + r) * a.founder_pitch + c0 : nullptr;
            uint8_t* const gd = a.dose ? a.dose + (out_row0 + r) * a.dose_pitch + c0 / P : nullptr;
            if (tid == 0) {
                if (gf && (n_cell & ~15)) bulk_store(gf, img, (uint32_t)(n_cell & ~15));
                if (DK >= 1 && gd && (n_ind & ~15)) bulk_store(gd, drow, (uint32_t)(n_ind & ~15));
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            // ragged ends (a bulk store moves multiples of 16 bytes)
            if (gf && tid < (n_cell & 15)) gf[(n_cell & ~15) + tid] = img[(n_cell & ~15) + tid];
            if (DK >= 1 && gd && tid >= 32 && tid - 32 < (n_ind & 15)) gd[(n_ind & ~15) + tid - 32] = drow[(n_ind & ~15) + tid - 32];
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Generic emission: one thread per (locus row, emitted individual); any ploidy, u8 or u16 founders,
// any pitch. Used for ploidies without a tiled instantiation and for unaligned device targets.
struct EmitGenericArgs {
    int C, R, P;
    int64_t N;
    const uint64_t* run_desc;
    const uint32_t* run_lb;
    const uint16_t* run_founder;
    const int64_t* locus_off;
    int64_t ind_first, ind_count, n_cols_ind;
    int64_t locus_first;          // row 0 of the outputs
    int64_t l_begin, l_count;     // global loci of this launch
    int64_t q_first, q_count;     // emitted individuals (columns) of this launch
    void* founder; int64_t founder_pitch; int founder_bytes;
    uint8_t* allele; int64_t allele_pitch;
    uint8_t* dose; int64_t dose_pitch;
    int dose_mode, dose_which;
    const uint8_t* code; const uint8_t* match_code; int A;
};

__global__ void __launch_bounds__(256) k_emit_generic(const EmitGenericArgs a) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= a.l_count * a.q_count) return;
    const int64_t q = a.q_first + t % a.q_count;
    const int64_t gl = a.l_begin + t / a.q_count;
    int c = 0;
    while (gl >= a.locus_off[c + 1]) c++;
    const uint32_t l = (uint32_t)(gl - a.locus_off[c]);
    const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
    const int64_t hid0 = (((int64_t)c * a.R + r) * a.N + i) * a.P;
    const int64_t row = gl - a.locus_first;
    int dose = 0;
    const uint8_t* codes = a.code ? a.code + gl * a.A : nullptr;
    const int match = a.dose_mode == 1 ? (int)a.match_code[gl] : a.dose_which;
    for (int h = 0; h < a.P; h++) {
        const uint64_t d = a.run_desc[hid0 + h];
        const uint32_t b = (uint32_t)desc_begin(d), n = desc_count(d);
        uint32_t lo = 0, hi = n;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.run_lb[b + mid] <= l) lo = mid; else hi = mid; }
        const int f = a.run_founder[b + lo];
        if (a.founder) {
            if (a.founder_bytes == 2) *reinterpret_cast<uint16_t*>((uint8_t*)a.founder + row * a.founder_pitch + (q * a.P + h) * 2) = (uint16_t)f;
            else ((uint8_t*)a.founder)[row * a.founder_pitch + q * a.P + h] = (uint8_t)f;
        }
        int al = f;
        if (codes) al = codes[f];
        if (a.allele) a.allele[row * a.allele_pitch + q * a.P + h] = (uint8_t)al;
        if ((a.dose_mode == 1 ? al : f) == match) dose++;
    }
    if (a.dose) a.dose[row * a.dose_pitch + q] = (uint8_t)dose;
}

// ------------------------------------------------------------------------------------------
// K3 — the reference's table files as text, formatted on the device (Main.writeGenotypesFile /
// writeAlleledoseFile, Main.java:903-1022: `label { '\t' token } '\n'` per locus, one
// PrintWriter.print per cell in the reference). One block per table row: the row's cells are read
// into shared memory a chunk at a time, every thread sizes (pass 1) or writes (pass 2) the tokens of
// its run of consecutive cells at the offset a block scan gives it, and the chunk's text leaves
// shared memory with coalesced stores. Pass 1 only measures; a device scan of the row lengths
// places the rows back to back.
constexpr int TEXT_THREADS = 256;

struct TextArgs {
    const uint8_t* cells; int64_t pitch; int cell_bytes;   // device table [rows][pitch], u8 or u16 cells
    int64_t n_cells;               // tokens per row
    int64_t row0_locus;            // global locus of table row 0
    int n_rows;
    const int64_t* label_off;      // [L+1] locus names (psim_set_locus_names)
    const char* labels;
    int kind;                      // 0 = decimal numbers, 1 = the locus' allele names by code
    const int64_t* code_off;       // kind 1: [L+1] first name of each locus
    const int64_t* name_off;       //         [names+1]
    const char* names;
    int max_tok;                   // longest token including its tab
    int chunk;                     // cells per chunk, a multiple of TEXT_THREADS
    int64_t* row_len;              // pass 1 out [n_rows]
    const int64_t* row_off;        // pass 2 in  [n_rows+1]
    char* text;                    // pass 2 out
};

template <bool WRITE>
__global__ void __launch_bounds__(TEXT_THREADS) k_text_rows(const TextArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_warp[TEXT_THREADS / 32];
    __shared__ uint32_t s_nlen[256], s_nbeg[256];
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    uint8_t* const s_cells = smem;                                   // chunk * cell_bytes
    char* const s_text = reinterpret_cast<char*>(smem + (size_t)a.chunk * a.cell_bytes);   // chunk * max_tok (pass 2)
    const int cpt = a.chunk / TEXT_THREADS;
    for (int r = blockIdx.x; r < a.n_rows; r += gridDim.x) {
        const int64_t gl = a.row0_locus + r;
        const uint8_t* row = a.cells + (int64_t)r * a.pitch;
        const int64_t lab0 = a.label_off[gl];
        const int lab_n = (int)(a.label_off[gl + 1] - lab0);
        char* out = WRITE ? a.text + a.row_off[r] : nullptr;
        if (WRITE)
            for (int i = tid; i < lab_n; i += TEXT_THREADS) out[i] = a.labels[lab0 + i];
        int64_t base = 0;   // first name of the locus
        if (a.kind == 1) {
            base = a.code_off[gl];
            const int k = (int)(a.code_off[gl + 1] - base);
            __syncthreads();
            for (int i = tid; i < k && i < 256; i += TEXT_THREADS) {
                s_nbeg[i] = (uint32_t)(a.name_off[base + i] - a.name_off[base]);
                s_nlen[i] = (uint32_t)(a.name_off[base + i + 1] - a.name_off[base + i]);
            }
        }
        int64_t running = lab_n;
        for (int64_t c0 = 0; c0 < a.n_cells; c0 += a.chunk) {
            const int n = (int)min((int64_t)a.chunk, a.n_cells - c0);
            __syncthreads();   // previous chunk's text has been copied out; name table loaded
            {   // the chunk's cells, coalesced (rows start 4-byte aligned: staging pitches are multiples of 128)
                const int nbytes = n * a.cell_bytes;
                const uint8_t* src = row + c0 * a.cell_bytes;
                for (int i = tid * 4; i < (nbytes & ~3); i += TEXT_THREADS * 4)
                    *reinterpret_cast<uint32_t*>(s_cells + i) = *reinterpret_cast<const uint32_t*>(src + i);
                if (tid < (nbytes & 3)) s_cells[(nbytes & ~3) + tid] = src[(nbytes & ~3) + tid];
            }
            __syncthreads();
            const int i0 = tid * cpt, i1 = min(i0 + cpt, n);
            auto cell = [&](int i) -> uint32_t {
                return a.cell_bytes == 1 ? (uint32_t)s_cells[i] : (uint32_t)reinterpret_cast<const uint16_t*>(s_cells)[i];
            };
            auto tok_len = [&](uint32_t v) -> uint32_t {
                if (a.kind == 1) return 1 + s_nlen[v & 255];
                return v < 10 ? 2 : v < 100 ? 3 : v < 1000 ? 4 : v < 10000 ? 5 : 6;
            };
            uint32_t len = 0;
            for (int i = i0; i < i1; i++) len += tok_len(cell(i));
            // block exclusive scan of the per-thread lengths
            uint32_t x = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
            if (lane == 31) s_warp[wq] = x;
            __syncthreads();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int w = 0; w < TEXT_THREADS / 32; w++) { const uint32_t v = s_warp[w]; if (w < wq) wbase += v; total += v; }
            if (WRITE) {
                char* p = s_text + wbase + x - len;
                for (int i = i0; i < i1; i++) {
                    const uint32_t v = cell(i);
                    *p++ = '\t';
                    if (a.kind == 1) {
                        const char* nm = a.names + a.name_off[base] + s_nbeg[v & 255];
                        const uint32_t nl = s_nlen[v & 255];
                        for (uint32_t q = 0; q < nl; q++) *p++ = nm[q];
                    } else {
                        if (v >= 10000) *p++ = (char)('0' + v / 10000);
                        if (v >= 1000) *p++ = (char)('0' + v / 1000 % 10);
                        if (v >= 100) *p++ = (char)('0' + v / 100 % 10);
                        if (v >= 10) *p++ = (char)('0' + v / 10 % 10);
                        *p++ = (char)('0' + v % 10);
                    }
                }
                __syncthreads();
                char* dst = out + running;
                // coalesced copy: head bytes up to 4-byte alignment of the destination, then words
                const uint32_t mis = (uint32_t)((4 - ((uintptr_t)dst & 3)) & 3);
                const uint32_t head = min(mis, total);
                if (tid < head) dst[tid] = s_text[tid];
                const uint32_t words = (total - head) >> 2;
                for (uint32_t i = tid; i < words; i += TEXT_THREADS) {
                    const char* s = s_text + head + 4 * i;   // shared memory side may be misaligned: assemble bytes
                    const uint32_t w = (uint32_t)(uint8_t)s[0] | ((uint32_t)(uint8_t)s[1] << 8) | ((uint32_t)(uint8_t)s[2] << 16) | ((uint32_t)(uint8_t)s[3] << 24);
                    *reinterpret_cast<uint32_t*>(dst + head + 4 * i) = w;
                }
                const uint32_t done = head + 4 * words;
                if (tid < total - done) dst[done + tid] = s_text[done + tid];
            }
            running += total;
        }
        if (tid == 0) {
            if (WRITE) out[running] = '\n';
            else a.row_len[r] = running + 1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4 — TEST-mode statistics (Main.testMeiosis, Main.java:1660-1721) as device reductions over the
// gametes that made the simulated individuals: homologs [0, P/2) of an individual are the gamete
// of parent 0's meiosis, [P/2, P) that of parent 1's. All of it works on the locus-index runs of K0.
struct StatsArgs {
    int C, R, P, A;
    int64_t N;
    const uint64_t* run_desc; const uint32_t* run_lb; const uint16_t* run_founder;
    const uint64_t* desc;             // unmerged segment lists (segment counts as the reference holds them)
    const int8_t* cfg_quad;           // [R*N][C][2]
    int64_t ind_first, ind_count;
    int chrom; uint32_t Lc;
    uint32_t* allele_count;           // [A][Lc]
    uint32_t* dose_count;             // [A][Lc][P/2+1]
    uint32_t* homozygous_count;       // [Lc]
    int32_t* diff;                    // [(Lc+1)][(Lc+1)] 2-D difference array of the recombination counts
    uint32_t* segments_hist;          // [64]
    uint32_t* quadrivalent_hist;      // [P/4+1]
};

__device__ __forceinline__ uint32_t founder_at_row(const StatsArgs& a, int64_t hid, uint32_t l) {
    const uint64_t d = a.run_desc[hid];
    const uint32_t b = (uint32_t)desc_begin(d), n = desc_count(d);
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.run_lb[b + mid] <= l) lo = mid; else hi = mid; }
    return a.run_founder[b + lo];
}

// one block per locus: founder-allele counts, dose histogram and homozygosity over all gametes
__global__ void __launch_bounds__(256) k_stats_locus(const StatsArgs a) {
    extern __shared__ uint32_t s_hist[];     // [A] counts, then [A][P/2+1] doses, then 1 homozygous
    const uint32_t l = blockIdx.x;
    const int p2 = a.P / 2, nd = p2 + 1;
    const int n_bins = a.A + a.A * nd + 1;
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t n_gam = 2 * a.ind_count * a.R;
    uint32_t zero_gametes = 0;   // gametes seen by this thread: dose 0 bins are filled in bulk at the end
    for (int64_t g = threadIdx.x; g < n_gam; g += blockDim.x) {
        const int half = (int)(g & 1);
        const int64_t q = g >> 1;
        const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
        const int64_t hid0 = (((int64_t)a.chrom * a.R + r) * a.N + i) * a.P + (int64_t)half * p2;
        uint32_t f[MAXP / 2];
        for (int h = 0; h < p2; h++) f[h] = founder_at_row(a, hid0 + h, l);
        bool homoz = false;
        for (int h = 0; h < p2; h++) {
            atomicAdd(&s_hist[f[h]], 1u);
            int copies = 0, first = 1;
            for (int k = 0; k < p2; k++) { if (f[k] == f[h]) { copies++; if (k < h) first = 0; } }
            if (first) atomicAdd(&s_hist[a.A + f[h] * nd + copies], 1u);   // once per distinct allele of the gamete
            if (copies > 1) homoz = true;
        }
        if (homoz) atomicAdd(&s_hist[a.A + a.A * nd], 1u);
        zero_gametes++;
    }
    __syncthreads();
    // dose 0 of allele f = gametes - gametes with dose >= 1
    for (int i = threadIdx.x; i < a.A; i += blockDim.x) {
        uint32_t with = 0;
        for (int dd = 1; dd < nd; dd++) with += s_hist[a.A + i * nd + dd];
        s_hist[a.A + i * nd] = (uint32_t)n_gam - with;
    }
    __syncthreads();
    (void)zero_gametes;
    if (a.allele_count) for (int i = threadIdx.x; i < a.A; i += blockDim.x) a.allele_count[(size_t)i * a.Lc + l] = s_hist[i];
    if (a.dose_count)
        for (int i = threadIdx.x; i < a.A * nd; i += blockDim.x)
            a.dose_count[((size_t)(i / nd) * a.Lc + l) * nd + i % nd] = s_hist[a.A + i];
    if (a.homozygous_count && threadIdx.x == 0) a.homozygous_count[l] = s_hist[a.A + a.A * nd];
}

// one thread per gamete homolog: every pair of runs with different founder alleles is a rectangle
// of (loc, loc2 < loc) pairs that count one recombinant (recombCountSel): four corner updates of a
// 2-D difference array; segment-count histogram on the way
__global__ void __launch_bounds__(256) k_stats_homolog(const StatsArgs a) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n_hom = a.ind_count * a.R * a.P;
    if (t >= n_hom) return;
    const int h = (int)(t % a.P);
    const int64_t q = t / a.P;
    const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
    const int64_t hid = (((int64_t)a.chrom * a.R + r) * a.N + i) * a.P + h;
    if (a.segments_hist) {
        const uint32_t nseg = desc_count(a.desc[hid]);
        atomicAdd(&a.segments_hist[min(nseg - 1, 63u)], 1u);
    }
    if (!a.diff) return;
    const uint64_t d = a.run_desc[hid];
    const uint32_t b = (uint32_t)desc_begin(d), n = desc_count(d);
    const size_t W = (size_t)a.Lc + 1;
    for (uint32_t j = 1; j < n; j++) {
        const uint32_t j0 = a.run_lb[b + j], j1 = j + 1 < n ? a.run_lb[b + j + 1] : a.Lc;
        const uint32_t fj = a.run_founder[b + j];
        for (uint32_t k = 0; k < j; k++) {
            if (a.run_founder[b + k] == fj) continue;
            const uint32_t k0 = a.run_lb[b + k], k1 = a.run_lb[b + k + 1];
            // rows loc in [j0, j1), columns loc2 in [k0, k1)
            atomicAdd(&a.diff[(size_t)j0 * W + k0], 1);
            atomicAdd(&a.diff[(size_t)j0 * W + k1], -1);
            atomicAdd(&a.diff[(size_t)j1 * W + k0], -1);
            atomicAdd(&a.diff[(size_t)j1 * W + k1], 1);
        }
    }
}

// 2-D inclusive prefix sum of the difference array, in place: along rows, then along columns
__global__ void k_prefix_rows(int32_t* d, uint32_t W) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= W) return;
    int32_t acc = 0;
    for (uint32_t c = 0; c < W; c++) { acc += d[(size_t)row * W + c]; d[(size_t)row * W + c] = acc; }
}
__global__ void k_prefix_cols(int32_t* d, uint32_t W) {
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= W) return;
    int32_t acc = 0;
    for (uint32_t r = 0; r < W; r++) { acc += d[(size_t)r * W + col]; d[(size_t)r * W + col] = acc; }
}
__global__ void k_stats_quads(const StatsArgs a) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n = a.ind_count * a.R * 2;
    if (t >= n) return;
    const int par = (int)(t & 1);
    const int64_t q = t >> 1;
    const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
    const int8_t v = a.cfg_quad[((r * a.N + i) * a.C + a.chrom) * 2 + par];
    if (v >= 0) atomicAdd(&a.quadrivalent_hist[v], 1u);
}

// _allAlleledose rows: one thread per (row, emitted individual)
struct AllDoseArgs {
    int C, R, P;
    int64_t N;
    const uint64_t* run_desc;
    const uint32_t* run_lb;
    const uint16_t* run_founder;
    const int64_t* locus_off;
    int64_t ind_first, ind_count, n_cols_ind;
    int64_t l_begin, l_count;
    const int64_t* row_off;   // device, per locus of the request (l_count + 1), relative rows
    int64_t row_first;        // rows before this launch (outputs are offset by the caller)
    uint8_t* out; int64_t pitch;
    const uint8_t* code; int A;   // code == null: founder alleles
};
__global__ void __launch_bounds__(256) k_emit_all_dose(const AllDoseArgs a) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= a.l_count * a.n_cols_ind) return;
    const int64_t q = t % a.n_cols_ind;
    const int64_t li = t / a.n_cols_ind;
    const int64_t gl = a.l_begin + li;
    int c = 0;
    while (gl >= a.locus_off[c + 1]) c++;
    const uint32_t l = (uint32_t)(gl - a.locus_off[c]);
    const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
    const int64_t hid0 = (((int64_t)c * a.R + r) * a.N + i) * a.P;
    const int64_t r0 = a.row_off[li] - a.row_first, r1 = a.row_off[li + 1] - a.row_first;
    for (int64_t rr = r0; rr < r1; rr++) a.out[rr * a.pitch + q] = 0;
    const uint8_t* codes = a.code ? a.code + gl * a.A : nullptr;
    for (int h = 0; h < a.P; h++) {
        const uint64_t d = a.run_desc[hid0 + h];
        const uint32_t b = (uint32_t)desc_begin(d), n = desc_count(d);
        uint32_t lo = 0, hi = n;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.run_lb[b + mid] <= l) lo = mid; else hi = mid; }
        int f = a.run_founder[b + lo];
        if (codes) f = codes[f];
        if (r0 + f < r1) a.out[(r0 + f) * a.pitch + q] += 1;
    }
}

// ------------------------------------------------------------------------------------------
// host-order <-> device-order helpers ([i][c][h] on the host side)
__global__ void k_scatter_desc(int64_t n, int64_t first, int r, int C, int R, int64_t N, int P,
                               const int64_t* __restrict__ seg_off, int64_t slab_base, uint64_t* __restrict__ desc) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int h = (int)(t % P);
    const int c = (int)((t / P) % C);
    const int64_t i = first + t / ((int64_t)P * C);
    const int64_t hid = (((int64_t)c * R + r) * N + i) * P + h;
    desc[hid] = make_desc((uint64_t)(slab_base + seg_off[t]), (uint64_t)(seg_off[t + 1] - seg_off[t]));
}
__global__ void k_gather_counts(int64_t n, int64_t first, int r, int C, int R, int64_t N, int P,
                                const uint64_t* __restrict__ desc, int64_t* __restrict__ cnt) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int h = (int)(t % P);
    const int c = (int)((t / P) % C);
    const int64_t i = first + t / ((int64_t)P * C);
    const int64_t hid = (((int64_t)c * R + r) * N + i) * P + h;
    cnt[t] = desc_count(desc[hid]);
}
__global__ void k_gather_segments(int64_t n, int64_t first, int r, int C, int R, int64_t N, int P,
                                  const uint64_t* __restrict__ desc, const int64_t* __restrict__ off,
                                  const double* __restrict__ seg_pos, const int32_t* __restrict__ seg_founder,
                                  double* __restrict__ out_pos, int32_t* __restrict__ out_founder) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int h = (int)(t % P);
    const int c = (int)((t / P) % C);
    const int64_t i = first + t / ((int64_t)P * C);
    const int64_t hid = (((int64_t)c * R + r) * N + i) * P + h;
    const uint64_t d = desc[hid];
    const int64_t b = (int64_t)desc_begin(d);
    const int cn = (int)desc_count(d);
    int64_t w = off[t];
    for (int s = 0; s < cn; s++, w++) { out_pos[w] = seg_pos[b + s]; out_founder[w] = seg_founder[b + s]; }
}
// Per-locus tables for the coded emission: the code to count for the dose table (Main.java:987-1000:
// MAX / MIN allele name or the name of founder allele `which`), the padded 16-entry code table and
// the nibble mask of founder alleles that carry the counted code.
__global__ void k_match_codes(int64_t L, int A, int which, const uint8_t* __restrict__ code, uint8_t* __restrict__ match,
                              uint32_t* __restrict__ code16, uint32_t* __restrict__ nibmask, uint4* __restrict__ mtab) {
    const int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (l >= L) return;
    const uint8_t* cd = code + l * A;
    int m;
    if (which == PSIM_DOSE_MAX_ALLELE) { m = 0; for (int k = 0; k < A; k++) m = max(m, (int)cd[k]); }
    else if (which == PSIM_DOSE_MIN_ALLELE) { m = 255; for (int k = 0; k < A; k++) m = min(m, (int)cd[k]); }
    else m = cd[(which < 0 || which >= A) ? 0 : which];
    match[l] = (uint8_t)m;
    if (A <= 16) {
        uint32_t t[4] = {0, 0, 0, 0}, nm[2] = {0, 0};
        for (int k = 0; k < A; k++) {
            t[k >> 2] |= (uint32_t)cd[k] << (8 * (k & 3));
            if (cd[k] == m) nm[k >> 3] |= 0xFu << (4 * (k & 7));
        }
        for (int q = 0; q < 4; q++) code16[l * 4 + q] = t[q];
        nibmask[l * 2] = nm[0];
        nibmask[l * 2 + 1] = nm[1];
        uint32_t mt[2] = {0, 0};   // byte table for the PRMT dose lookup (A <= 8 paths)
        for (int k = 0; k < A && k < 8; k++)
            if (cd[k] == m) mt[k >> 2] |= 1u << (8 * (k & 3));
        mtab[l] = make_uint4(mt[0], mt[1], mt[0] << 4, mt[1] << 4);
    }
}

}  // namespace psim

// ==========================================================================================
// host side
using namespace psim;

namespace {

std::string g_create_error;

struct psim_ctx_impl {
    psim_config cfg{};
    std::string err;
    int C = 0, P = 0, R = 1;
    int64_t N = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    std::vector<cudaEvent_t> kev;   // event pool: triples (before plan, before stream, after stream) per launch
    int kev_used = 0;
    ModelParams model{};
    std::vector<ChromParams> chrom_h;
    ChromParams* chrom_d = nullptr;
    double max_len = 0.0;
    int piece_cap = 0;
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
    int* error_flag_d = nullptr;
    int* tile_counter_d = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;
    void* staging = nullptr;
    size_t staging_bytes = 0;
    // grow-only scratch of psim_simulate (no allocation in the steady state of a Monte Carlo loop)
    struct Buf { void* p = nullptr; size_t cap = 0; };
    Buf b_lev, b_np, b_cnt, b_off, b_ppos, b_phom, b_find, b_fal;
    Buf b_plan, b_dense_list, b_task, b_text, b_rowlen, b_rowoff;
    // export buffers
    int64_t* exp_off_d = nullptr; int32_t* exp_founder_d = nullptr; double* exp_pos_d = nullptr;
    int64_t exp_off_cap = 0, exp_seg_cap = 0;
    psim_stats stats{};
    int sm_count = 148;
};

#define CTX ((psim_ctx_impl*)ctx)

int fail(psim_ctx_impl* c, int code, const std::string& msg) { c->err = msg; return code; }

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
// Row-image emission (k_band_plan + k_emit_rows) of loci [l_begin, l_end). Returns PSIM_OK and sets
// `done` when the request fits the path; otherwise leaves `done` false for the tiled / generic paths.
template <int DK>
static void launch_rows(psim_ctx_impl* c, const RowsArgs& ra, int blocks, size_t smem) {
    cudaFuncSetAttribute(k_emit_rows<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS_SMEM);
    k_emit_rows<DK><<<blocks, ROWS_THREADS, smem, c->stream>>>(ra);
}

static int emit_rows_path(psim_ctx_impl* c, int64_t ind_first, int64_t ind_count, int64_t l_begin, int64_t l_end, int64_t row0_locus,
                          uint8_t* founder, int64_t founder_pitch, uint8_t* dose, int64_t dose_pitch, int dose_which, int mode,
                          bool& done) {
    done = false;
    static const bool disabled = getenv("PSIM_NO_ROWS") != nullptr;
    const int P = c->P;
    if (disabled || !(mode == 0 || (mode == 1 && (P == 2 || P == 4)))) return PSIM_OK;
    const int DK = !dose ? 0 : mode == 0 ? 1 : P == 4 ? 2 : 3;
    const int64_t n_cols_ind = (int64_t)c->R * ind_count, n_cols = n_cols_ind * P;
    const int cap = 8192;
    const double per_cell = (double)ROWS_NBUF + (DK >= 2 ? 0.5 : 0.0) + (DK >= 1 ? (double)ROWS_NBUF : 0.0) / P;
    const int64_t budget = ROWS_SMEM - 2 * (ROWS_MAX_BAND + 2) * 4 - cap * 4;
    const int64_t quantum = 16 * P;
    int64_t seg_max = (int64_t)(budget / per_cell) / quantum * quantum;
    seg_max = std::min<int64_t>(seg_max, 65535 / quantum * quantum);
    const int n_seg = (int)((n_cols + seg_max - 1) / seg_max);
    const int64_t seg_cells = ((n_cols + n_seg - 1) / n_seg + quantum - 1) / quantum * quantum;
    // run starts per cell and row, from the resident haplostructs (as in the tiled path)
    const double seg_per_hom = c->n_hom > 0 ? (double)c->slab_used / (double)c->n_hom : 1.0;
    const double loci_per_chrom = std::max(1.0, (double)c->L / c->C);
    const double starts_per_cell = std::max(0.0, seg_per_hom - 1.0) / loci_per_chrom;
    static const int env_band = getenv("PSIM_BAND_ROWS") ? atoi(getenv("PSIM_BAND_ROWS")) : 0;
    static const int env_bps = getenv("PSIM_ROWS_BLOCKS_PER_SM") ? atoi(getenv("PSIM_ROWS_BLOCKS_PER_SM")) : 0;
    const int n_blocks = c->sm_count * (env_bps > 0 ? env_bps : 1);
    // band height: one unit per resident block when the request is large, never so tall that a unit's
    // expected run starts come near the list capacity
    // tallest band whose expected run starts stay within half the list capacity
    const double per_row = starts_per_cell * (double)seg_cells;
    const bool forced = getenv("PSIM_ROWS_FORCE") != nullptr;   // test knob: take this path even if lists will overflow
    if (!forced && per_row * 8.0 > cap / 2) return PSIM_OK;      // sparse map: tiled / dense path
    const int64_t band_cap = env_band > 0 ? std::min(env_band, ROWS_MAX_BAND)
                           : std::max<int64_t>(8, std::min<int64_t>(ROWS_MAX_BAND, per_row > 0 ? (int64_t)((cap / 2) / per_row) : ROWS_MAX_BAND));

    RowsArgs ra{};
    ra.C = c->C; ra.R = c->R; ra.P = P; ra.N = c->N;
    ra.run_desc = c->run_desc_d; ra.run_lb = c->run_lb_d; ra.run_founder = c->run_founder_d;
    ra.ind_first = ind_first; ra.ind_count = ind_count; ra.n_cols_ind = n_cols_ind;
    ra.locus_first = row0_locus;
    ra.n_seg = n_seg; ra.seg_cells = (int)seg_cells; ra.cap = cap;
    ra.plan_stride = (16 + seg_cells + (int64_t)cap * 4 + 127) / 128 * 128;
    ra.founder = founder; ra.founder_pitch = founder_pitch;
    ra.dose = dose; ra.dose_pitch = dose_pitch;
    ra.dose_which = dose_which; ra.mtab = c->mtab_d;
    ra.unit_counter = c->tile_counter_d;
    const size_t smem = ROWS_NBUF * (size_t)seg_cells + (DK >= 2 ? seg_cells / 2 : 0) + (size_t)(DK >= 1 ? ROWS_NBUF : 0) * (seg_cells / P) +
                        (size_t)cap * 4 + 2 * (ROWS_MAX_BAND + 2) * 4;
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
            // resident blocks; no band is taller than band_cap.
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
                    nb = std::min(nb, len);
                    ra.chunk_bands[q] = (int)nb;
                    rem[q] = ideal - (double)nb;
                    given += nb;
                }
                while (given < target) {
                    int best = -1;
                    for (int q = 0; q < ra.n_chunks; q++)
                        if (ra.chunk_bands[q] < ra.chunk_l1[q] - ra.chunk_l0[q] && (best < 0 || rem[q] > rem[best])) best = q;
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
            for (int64_t tg = t_min; tg < t_min + 4 * n_blocks && tg <= std::max<int64_t>(t_min, total / (env_band > 0 ? 1 : 16)); tg++) {
                const int64_t units = allocate(tg) * n_seg;
                const double eff = (double)units / (double)((units + n_blocks - 1) / n_blocks * n_blocks);
                if (eff > best_eff + 1e-9) { best_eff = eff; best_t = tg; }
                if (env_band > 0 || (best_eff >= 0.97 && tg >= best_t + 8)) break;   // good enough; look a little further for a full wave
            }
            allocate(best_t);
            for (int q = 0; q < ra.n_chunks; q++) ra.chunk_unit_base[q + 1] = ra.chunk_unit_base[q] + ra.chunk_bands[q] * n_seg;
        }
        ra.n_units = ra.chunk_unit_base[ra.n_chunks];
        if (int rc = ensure_buf(c, c->b_plan, (size_t)ra.n_units * ra.plan_stride)) return rc;
        if (int rc = ensure_buf(c, c->b_dense_list, ((size_t)ra.n_units + 1) * sizeof(int))) return rc;
        ra.plan = (uint8_t*)c->b_plan.p;
        ra.overflow = (int*)c->b_dense_list.p;
        CUDA_TRY(cudaMemset2DAsync(ra.plan, (size_t)ra.plan_stride, 0, 16, (size_t)ra.n_units, c->stream));   // list counters
        CUDA_TRY(cudaMemsetAsync(ra.overflow, 0, sizeof(int), c->stream));
        CUDA_TRY(cudaMemsetAsync(c->tile_counter_d, 0, sizeof(int), c->stream));
        k_band_plan<<<dim3((unsigned)((n_cols + 255) / 256), (unsigned)ra.n_chunks), 256, 0, c->stream>>>(ra);
        cudaEventRecord(next_kev(c), c->stream);
        const int blocks = std::min(ra.n_units, n_blocks);
        if (DK == 0) launch_rows<0>(c, ra, blocks, smem);
        else if (DK == 1) launch_rows<1>(c, ra, blocks, smem);
        else if (DK == 2) launch_rows<2>(c, ra, blocks, smem);
        else launch_rows<3>(c, ra, blocks, smem);
        cudaEventRecord(next_kev(c), c->stream);
        c->stats.kernel_launches += 2;
        c->stats.emit_kernel_launches++;
        CUDA_TRY(cudaGetLastError());
        // units whose run-start list overflowed (not expected with the band heuristic): searched per cell
        int n_over = 0;
        CUDA_TRY(cudaMemcpyAsync(&n_over, ra.overflow, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (n_over > 0) {
            std::vector<int> units((size_t)n_over);
            CUDA_TRY(cudaMemcpy(units.data(), ra.overflow + 1, (size_t)n_over * sizeof(int), cudaMemcpyDeviceToHost));
            for (int u : units) {
                int k = 0;
                while (u >= ra.chunk_unit_base[k + 1]) k++;
                const int j = (u - ra.chunk_unit_base[k]) / n_seg, sg = (u - ra.chunk_unit_base[k]) % n_seg;
                const int64_t len = ra.chunk_l1[k] - ra.chunk_l0[k];
                const int64_t g0 = ra.chunk_l0[k] + len * j / ra.chunk_bands[k], g1 = ra.chunk_l0[k] + len * (j + 1) / ra.chunk_bands[k];
                EmitGenericArgs ga{};
                ga.C = c->C; ga.R = c->R; ga.P = P; ga.N = c->N;
                ga.run_desc = c->run_desc_d; ga.run_lb = c->run_lb_d; ga.run_founder = c->run_founder_d;
                ga.locus_off = c->locus_off_d;
                ga.ind_first = ind_first; ga.ind_count = ind_count; ga.n_cols_ind = n_cols_ind;
                ga.locus_first = row0_locus; ga.l_begin = g0; ga.l_count = g1 - g0;
                ga.q_first = (int64_t)sg * (seg_cells / P);
                ga.q_count = std::min<int64_t>(seg_cells / P, n_cols_ind - ga.q_first);
                ga.founder = founder; ga.founder_pitch = founder_pitch; ga.founder_bytes = 1;
                ga.dose = dose; ga.dose_pitch = dose_pitch;
                ga.dose_mode = mode == 0 ? 0 : 1; ga.dose_which = dose_which;
                ga.code = c->code_d; ga.match_code = c->match_d; ga.A = c->A;
                k_emit_generic<<<grid_for(ga.l_count * ga.q_count, 256), 256, 0, c->stream>>>(ga);
                c->stats.kernel_launches++;
                CUDA_TRY(cudaGetLastError());
            }
        }
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
    c->model.seed = (uint32_t)cfg->seed;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess &&
              cudaEventCreate(&c->ev2) == cudaSuccess && cudaEventCreate(&c->ev3) == cudaSuccess &&
              cudaMalloc((void**)&c->error_flag_d, sizeof(int)) == cudaSuccess &&
              cudaMalloc((void**)&c->tile_counter_d, sizeof(int)) == cudaSuccess;
    if (!ok) { g_create_error = std::string("CUDA setup failed: ") + cudaGetErrorString(cudaGetLastError()); delete c; return PSIM_ECUDA; }
    c->stream = c->own_stream;
    cudaMemset(c->error_flag_d, 0, sizeof(int));
    *out = (psim_ctx*)c;
    return PSIM_OK;
}

void psim_destroy(psim_ctx* ctx) {
    if (!ctx) return;
    psim_ctx_impl* c = CTX;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    free_population(c);
    cudaFree(c->chrom_d); cudaFree(c->seg_pos_d); cudaFree(c->seg_founder_d);
    cudaFree(c->run_lb_d); cudaFree(c->run_founder_d);
    cudaFree(c->locus_off_d); cudaFree(c->locus_pos_d); cudaFree(c->code_d); cudaFree(c->match_d);
    cudaFree(c->code16_d); cudaFree(c->nibmask_d); cudaFree(c->mtab_d);
    cudaFree(c->error_flag_d); cudaFree(c->tile_counter_d); cudaFree(c->cub_tmp); cudaFree(c->staging);
    cudaFree(c->exp_off_d); cudaFree(c->exp_founder_d); cudaFree(c->exp_pos_d);
    for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
    for (psim_ctx_impl::Buf* b : {&c->b_lev, &c->b_np, &c->b_cnt, &c->b_off, &c->b_ppos, &c->b_phom, &c->b_find, &c->b_fal,
                                 &c->b_plan, &c->b_dense_list, &c->b_task, &c->b_text, &c->b_rowlen, &c->b_rowoff}) cudaFree(b->p);
    cudaFree(c->label_off_d); cudaFree(c->labels_d); cudaFree(c->code_off_d); cudaFree(c->aname_off_d); cudaFree(c->anames_d);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
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
    // capacity of the per-chromatid piece list and the per-multivalent chiasma list (MAXE):
    // chiasmata on a 4-strand bundle are Poisson(2 len) (4 len for a parallel quadrivalent).
    const double mean_e = 4.0 * c->max_len;
    if (mean_e + 12.0 * std::sqrt(mean_e) + 16.0 > MAXE)
        return fail(c, PSIM_ECAPACITY, "chromosome too long for the device chiasma buffer (limit about 14 Morgan)");
    c->piece_cap = (int)std::ceil(c->max_len * 2.0 + 12.0 * std::sqrt(c->max_len * 2.0) + 8.0);
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
    int nf = 0;
    for (int64_t i = 0; i < n_ind; i++) {
        const bool f0 = parent0[i] < 0, f1 = parent1[i] < 0;
        if (f0 != f1) return fail(c, PSIM_EINVAL, "Individual " + std::to_string(i) + " has one known and one unknown parent, which is not allowed");
        if (f0) { nf++; continue; }
        if (parent0[i] >= i || parent1[i] >= i) return fail(c, PSIM_EINVAL, "pedigree is not in parents-first order");
    }
    free_population(c);
    c->N = n_ind;
    c->founder_allele_count = nf * c->P;
    c->par0_h.assign(parent0, parent0 + n_ind);
    c->par1_h.assign(parent1, parent1 + n_ind);
    c->hs_reps.assign(n_ind, 0);
    c->n_hom = (int64_t)c->C * c->R * c->N * c->P;
    CUDA_TRY(dev_alloc(&c->par0_d, (size_t)n_ind));
    CUDA_TRY(dev_alloc(&c->par1_d, (size_t)n_ind));
    CUDA_TRY(cudaMemcpy(c->par0_d, parent0, n_ind * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->par1_d, parent1, n_ind * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_TRY(dev_alloc(&c->desc_d, (size_t)c->n_hom));
    CUDA_TRY(cudaMemset(c->desc_d, 0, c->n_hom * sizeof(uint64_t)));
    CUDA_TRY(dev_alloc(&c->run_desc_d, (size_t)c->n_hom));
    const int64_t ncfg = (int64_t)c->R * c->N * c->C * 2;
    CUDA_TRY(dev_alloc(&c->cfg_quad_d, (size_t)ncfg));
    CUDA_TRY(dev_alloc(&c->cfg_seq_d, (size_t)ncfg * c->P));
    CUDA_TRY(cudaMemset(c->cfg_quad_d, 0xFF, ncfg));
    CUDA_TRY(cudaMemset(c->cfg_seq_d, 0xFF, ncfg * c->P));
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
    const int r = (int)(replicate - c->cfg.replicate_first);
    const int64_t nh = count * c->C * c->P;
    const int64_t nseg = seg_off[nh];
    for (int64_t t = 0; t < nh; t++) {
        if (seg_off[t + 1] <= seg_off[t]) return fail(c, PSIM_EINVAL, "every homolog needs at least one segment");
        if (seg_off[t + 1] - seg_off[t] >= (1 << 24)) return fail(c, PSIM_ECAPACITY, "more than 2^24 segments in one homolog");
    }
    for (int64_t s = 0; s < nseg; s++)
        if (founder[s] < 0 || founder[s] > 65535) return fail(c, PSIM_EINVAL, "invalid founder allele '" + std::to_string(founder[s]) + "'");
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
    return PSIM_OK;
}

// subset == nullptr: every individual without a haplostruct; otherwise only the listed non-founders
// (founders without a haplostruct are always initialised: it is cheap and every rank of a sharded
// pedigree needs them).
static int simulate_impl(psim_ctx_impl* c, int32_t max_founder_allele, const int32_t* subset, int64_t n_subset) {
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    cudaSetDevice(c->cfg.device);
    const int C = c->C, P = c->P, R = c->R;
    const int64_t N = c->N;
    for (int64_t i = 0; i < N; i++)
        if (c->hs_reps[i] != 0 && c->hs_reps[i] != R)
            return fail(c, PSIM_EINVAL, "preset haplostructs must cover the same individuals in every replicate");
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    // --- founders (Main.java:1510-1513)
    std::vector<int32_t> f_ind, f_allele;
    int32_t maxf = max_founder_allele;
    for (int64_t i = 0; i < N; i++) {
        if (c->hs_reps[i]) continue;
        if (c->par0_h[i] < 0) { f_ind.push_back((int32_t)i); f_allele.push_back(maxf + 1); maxf += P; }
    }
    if (maxf > 65535) return fail(c, PSIM_ECAPACITY, "more than 65536 founder alleles");
    if (!f_ind.empty()) {
        const int n_new = (int)f_ind.size();
        const int64_t nseg = (int64_t)n_new * R * C * P;
        if (int rc = ensure_slab(c, c->slab_used + nseg)) return rc;
        if (int rc = ensure_buf(c, c->b_find, n_new * sizeof(int32_t))) return rc;
        if (int rc = ensure_buf(c, c->b_fal, n_new * sizeof(int32_t))) return rc;
        int32_t *ind_d = (int32_t*)c->b_find.p, *al_d = (int32_t*)c->b_fal.p;
        CUDA_TRY(cudaMemcpyAsync(ind_d, f_ind.data(), n_new * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(al_d, f_allele.data(), n_new * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        k_init_founders<<<grid_for(nseg, 256), 256, 0, c->stream>>>(n_new * R, n_new, ind_d, al_d, C, R, N, P, c->chrom_d, c->desc_d,
                                                                    c->seg_pos_d, c->seg_founder_d, c->slab_used);
        c->stats.kernel_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // f_ind / f_allele are stack-owned host buffers
        c->slab_used += nseg;
        for (int32_t i : f_ind) c->hs_reps[i] = R;
    }
    // --- generations: level = 1 + max(level of parents) over individuals still to simulate
    std::vector<int32_t> level(N, 0);
    int32_t max_level = 0;
    std::vector<uint8_t> wanted;
    if (subset) {
        wanted.assign(N, 0);
        for (int64_t k = 0; k < n_subset; k++) {
            if (subset[k] < 0 || subset[k] >= N) return fail(c, PSIM_EINVAL, "individual index out of bounds");
            wanted[subset[k]] = 1;
        }
    }
    for (int64_t i = 0; i < N; i++) {
        if (c->hs_reps[i]) continue;
        if (subset && !wanted[i]) { level[i] = -1; continue; }   // left for another rank / a later call
        const int32_t l0 = level[c->par0_h[i]], l1 = level[c->par1_h[i]];
        if (l0 < 0 || l1 < 0) return fail(c, PSIM_ESTATE, "a parent of individual " + std::to_string(i) + " has no haplostruct yet");
        level[i] = 1 + std::max(l0, l1);
        max_level = std::max(max_level, level[i]);
    }
    std::vector<std::vector<int32_t>> by_level(max_level + 1);
    for (int64_t i = 0; i < N; i++) if (level[i] > 0) by_level[level[i]].push_back((int32_t)i);
    // batch size: bound the plan scratch to about 1.5 GiB
    const int64_t per_ind = (int64_t)C * R * P * (c->piece_cap * 9 + 4 + 16);
    const int64_t max_batch = std::max<int64_t>(1, (int64_t)(1536LL << 20) / per_ind);
    int64_t scratch_ind = 0;
    for (auto& v : by_level) scratch_ind = std::max<int64_t>(scratch_ind, std::min<int64_t>((int64_t)v.size(), max_batch));
    int32_t* lev_d = nullptr; int32_t* plan_np = nullptr; int64_t* plan_cnt = nullptr; int64_t* plan_off = nullptr;
    double* plan_pos = nullptr; uint8_t* plan_hom = nullptr;
    if (scratch_ind > 0) {
        const size_t nu = (size_t)scratch_ind * C * R * P;
        if (int rc = ensure_buf(c, c->b_lev, scratch_ind * sizeof(int32_t))) return rc;
        if (int rc = ensure_buf(c, c->b_np, nu * sizeof(int32_t))) return rc;
        if (int rc = ensure_buf(c, c->b_cnt, (nu + 1) * sizeof(int64_t))) return rc;
        if (int rc = ensure_buf(c, c->b_off, (nu + 1) * sizeof(int64_t))) return rc;
        if (int rc = ensure_buf(c, c->b_ppos, nu * c->piece_cap * sizeof(double))) return rc;
        if (int rc = ensure_buf(c, c->b_phom, nu * c->piece_cap)) return rc;
        lev_d = (int32_t*)c->b_lev.p; plan_np = (int32_t*)c->b_np.p; plan_cnt = (int64_t*)c->b_cnt.p;
        plan_off = (int64_t*)c->b_off.p; plan_pos = (double*)c->b_ppos.p; plan_hom = (uint8_t*)c->b_phom.p;
    }
    int rc_out = PSIM_OK;
    for (int lv = 1; lv <= max_level && rc_out == PSIM_OK; lv++) {
        const std::vector<int32_t>& inds = by_level[lv];
        for (int64_t b0 = 0; b0 < (int64_t)inds.size() && rc_out == PSIM_OK; b0 += max_batch) {
            const int n_lev = (int)std::min<int64_t>(max_batch, (int64_t)inds.size() - b0);
            const int64_t nu = (int64_t)n_lev * C * R * P;
            CUDA_TRY(cudaMemcpyAsync(lev_d, inds.data() + b0, n_lev * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            PlanArgs pa;
            pa.model = c->model; pa.chrom = c->chrom_d; pa.C = C; pa.R = R; pa.P = P; pa.N = N;
            pa.replicate_first = c->cfg.replicate_first;
            pa.n_lev = n_lev; pa.lev_ind = lev_d; pa.par0 = c->par0_d; pa.par1 = c->par1_d;
            pa.desc = c->desc_d; pa.seg_pos = c->seg_pos_d; pa.seg_founder = c->seg_founder_d;
            pa.piece_cap = c->piece_cap; pa.plan_np = plan_np; pa.plan_cnt = plan_cnt;
            pa.plan_pos = plan_pos; pa.plan_hom = plan_hom;
            pa.cfg_quad = c->cfg_quad_d; pa.cfg_seq = c->cfg_seq_d; pa.error_flag = c->error_flag_d;
            const int64_t n_half = (int64_t)n_lev * C * R * 2;
            const int64_t cap_bi = n_half * (P / 2), cap_quad = n_half * (P / 4);
            if (int rc = ensure_buf(c, c->b_task, (size_t)(cap_bi + cap_quad) * sizeof(int64_t) + 16)) { rc_out = rc; break; }
            pa.task_count = (int*)c->b_task.p;
            pa.task_bi = (int64_t*)((uint8_t*)c->b_task.p + 16);
            pa.task_quad = pa.task_bi + cap_bi;
            CUDA_TRY(cudaMemsetAsync(pa.task_count, 0, 16, c->stream));
            k_meiosis_config<<<grid_for(n_half, 128), 128, 0, c->stream>>>(pa);
            k_meiosis_mv<false><<<grid_for(cap_bi, 128), 128, 0, c->stream>>>(pa);
            c->stats.kernel_launches += 2;
            if (cap_quad > 0) {
                k_meiosis_mv<true><<<grid_for(cap_quad, 128), 128, 0, c->stream>>>(pa);
                c->stats.kernel_launches++;
            }
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemsetAsync(plan_cnt + nu, 0, sizeof(int64_t), c->stream));
            if (int rc = exclusive_scan_i64(c, plan_cnt, plan_off, nu + 1)) { rc_out = rc; break; }
            int64_t total = 0; int flag = 0;
            CUDA_TRY(cudaMemcpyAsync(&total, plan_off + nu, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaMemcpyAsync(&flag, c->error_flag_d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            if (flag) {
                cudaMemset(c->error_flag_d, 0, sizeof(int));
                rc_out = fail(c, flag, "meiosis plan exceeded a device buffer (too many chiasmata on one chromatid)");
                break;
            }
            if (int rc = ensure_slab(c, c->slab_used + total)) { rc_out = rc; break; }
            AssembleArgs aa;
            aa.chrom = c->chrom_d; aa.C = C; aa.R = R; aa.P = P; aa.N = N; aa.n_lev = n_lev; aa.lev_ind = lev_d;
            aa.par0 = c->par0_d; aa.par1 = c->par1_d; aa.desc = c->desc_d; aa.seg_pos = c->seg_pos_d;
            aa.seg_founder = c->seg_founder_d; aa.piece_cap = c->piece_cap; aa.plan_np = plan_np; aa.plan_cnt = plan_cnt;
            aa.plan_off = plan_off; aa.plan_pos = plan_pos; aa.plan_hom = plan_hom; aa.slab_base = c->slab_used;
            k_meiosis_assemble<<<grid_for(nu, 128), 128, 0, c->stream>>>(aa);
            c->stats.kernel_launches++;
            CUDA_TRY(cudaGetLastError());
            c->slab_used += total;
            c->stats.n_meioses += (int64_t)n_lev * R * 2;
            c->stats.n_units += (int64_t)n_lev * R * 2 * C;
        }
        for (int32_t i : inds) c->hs_reps[i] = R;
    }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->stats.ms_simulate = ms;
    c->stats.n_segments = c->slab_used;
    c->runs_valid = false;
    return rc_out;
}

int psim_simulate(psim_ctx* ctx, int32_t max_founder_allele) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    return simulate_impl(c, max_founder_allele, nullptr, 0);
}

int psim_simulate_subset(psim_ctx* ctx, int32_t max_founder_allele, const int32_t* individuals, int64_t count) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (count > 0 && !individuals) return fail(c, PSIM_EINVAL, "null individual list");
    static const int32_t none = 0;
    return simulate_impl(c, max_founder_allele, individuals ? individuals : &none, count);
}

static int check_range(psim_ctx_impl* c, uint32_t replicate, int64_t first, int64_t count, int& r) {
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
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    if (first < 0 || count <= 0 || first + count > c->N) return fail(c, PSIM_EINVAL, "individual range out of bounds");
    if (replicate < c->cfg.replicate_first || replicate >= c->cfg.replicate_first + c->cfg.replicate_count)
        return fail(c, PSIM_EINVAL, "replicate not held by this context");
    const int r = (int)(replicate - c->cfg.replicate_first);
    const int64_t nq = count * c->C * 2;
    std::vector<int8_t> q(nq), s(nq * c->P);
    const int64_t base = ((int64_t)r * c->N + first) * c->C * 2;
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
    cudaFree(c->locus_off_d); cudaFree(c->locus_pos_d); cudaFree(c->code_d); cudaFree(c->match_d);
    cudaFree(c->code16_d); cudaFree(c->nibmask_d); cudaFree(c->mtab_d);
    c->locus_off_d = nullptr; c->locus_pos_d = nullptr; c->code_d = nullptr; c->match_d = nullptr;
    c->code16_d = nullptr; c->nibmask_d = nullptr; c->mtab_d = nullptr;
    c->A = 0; c->code_h.clear();
    CUDA_TRY(dev_alloc(&c->locus_off_d, (size_t)c->C + 1));
    CUDA_TRY(dev_alloc(&c->locus_pos_d, (size_t)c->L));
    CUDA_TRY(cudaMemcpy(c->locus_off_d, locus_off, (c->C + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->locus_pos_d, pos, c->L * sizeof(double), cudaMemcpyHostToDevice));
    c->runs_valid = false;
    return PSIM_OK;
}

int psim_set_allele_codes(psim_ctx* ctx, int32_t n_alleles, const uint8_t* code) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (c->L == 0) return fail(c, PSIM_ESTATE, "psim_set_map must be called first");
    cudaSetDevice(c->cfg.device);
    cudaFree(c->code_d); cudaFree(c->match_d); cudaFree(c->code16_d); cudaFree(c->nibmask_d); cudaFree(c->mtab_d);
    c->code_d = nullptr; c->match_d = nullptr; c->code16_d = nullptr; c->nibmask_d = nullptr; c->mtab_d = nullptr;
    c->A = 0; c->code_h.clear(); c->match_which = 12345678;
    if (!code) return PSIM_OK;
    if (n_alleles <= 0 || n_alleles > 256) return fail(c, PSIM_EINVAL, "allele codes need 1..256 founder alleles");
    c->A = n_alleles;
    c->code_h.assign(code, code + (size_t)c->L * n_alleles);
    CUDA_TRY(dev_alloc(&c->code_d, (size_t)c->L * n_alleles));
    CUDA_TRY(dev_alloc(&c->match_d, (size_t)c->L));
    CUDA_TRY(dev_alloc(&c->code16_d, (size_t)c->L * 4));
    CUDA_TRY(dev_alloc(&c->nibmask_d, (size_t)c->L * 2));
    CUDA_TRY(dev_alloc(&c->mtab_d, (size_t)c->L));
    CUDA_TRY(cudaMemcpy(c->code_d, code, (size_t)c->L * n_alleles, cudaMemcpyHostToDevice));
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
                                                                 c->seg_founder_d, c->locus_off_d, c->locus_pos_d, c->run_desc_d,
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
                       uint8_t* dose, int64_t dose_pitch, int dose_mode, int dose_which) {
    const int P = c->P;
    const int64_t n_cols_ind = (int64_t)c->R * ind_count;
    int G = 0;
    if (P == 2) G = 8; else if (P == 4) G = 4; else if (P == 6) G = 8; else if (P == 8) G = 2;
    const bool aligned = (!founder || ((uintptr_t)founder % 16 == 0 && founder_pitch % 16 == 0)) &&
                         (!allele || ((uintptr_t)allele % 16 == 0 && allele_pitch % 16 == 0)) &&
                         (!dose || ((uintptr_t)dose % 16 == 0 && dose_pitch % 16 == 0));
    const int64_t max_pitch = (int64_t)1 << 25;   // row offsets inside a tile are 32-bit (64 rows)
    const bool small_pitch = founder_pitch < max_pitch && allele_pitch < max_pitch && dose_pitch < max_pitch;
    const bool tiled = G != 0 && aligned && small_pitch && founder_bytes == 1 && c->founder_allele_count <= 256 &&
                       c->run_cap < (int64_t)0xFFFFFFFF;
    if (tiled && !allele && (founder || dose)) {
        bool done = false;
        const int mode0 = !c->code_d ? 0 : c->A <= 8 ? 1 : c->A <= 16 ? 2 : 3;
        if (int rc = emit_rows_path(c, ind_first, ind_count, l_begin, l_end, row0_locus, (uint8_t*)founder, founder_pitch, dose, dose_pitch,
                                    dose_which, mode0, done)) return rc;
        if (done) return PSIM_OK;
    }
    if (tiled) {
        EmitArgs ea{};
        ea.C = c->C; ea.R = c->R; ea.P = P; ea.N = c->N;
        ea.run_desc = c->run_desc_d; ea.run_lb = c->run_lb_d; ea.run_founder = c->run_founder_d;
        ea.locus_off = c->locus_off_d;
        ea.ind_first = ind_first; ea.ind_count = ind_count; ea.n_cols_ind = n_cols_ind;
        ea.locus_first = row0_locus;
        static const int env_tile = getenv("PSIM_TILE_LOCI") ? atoi(getenv("PSIM_TILE_LOCI")) : 0;       // tuning knobs
        static const int env_bps = getenv("PSIM_BLOCKS_PER_SM") ? atoi(getenv("PSIM_BLOCKS_PER_SM")) : 0;
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

static int emit_prepare(psim_ctx_impl* c, int64_t ind_first, int64_t ind_count, int64_t locus_first, int64_t locus_count) {
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    if (c->L == 0) return fail(c, PSIM_ESTATE, "psim_set_map must be called first");
    if (ind_first < 0 || ind_count <= 0 || ind_first + ind_count > c->N) return fail(c, PSIM_EINVAL, "individual range out of bounds");
    if (locus_first < 0 || locus_count <= 0 || locus_first + locus_count > c->L) return fail(c, PSIM_EINVAL, "locus range out of bounds");
    for (int64_t i = ind_first; i < ind_first + ind_count; i++)
        if (c->hs_reps[i] != c->R) return fail(c, PSIM_ESTATE, "individual " + std::to_string(i) + " has no haplostruct yet");
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
    if (t->founder && t->founder_pitch < n_cols * fb) return fail(c, PSIM_EINVAL, "founder_pitch too small");
    if (t->allele && t->allele_pitch < n_cols) return fail(c, PSIM_EINVAL, "allele_pitch too small");
    if (t->dose && t->dose_pitch < n_cols_ind) return fail(c, PSIM_EINVAL, "dose_pitch too small");
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
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    if (t->memory == PSIM_MEM_DEVICE) {
        if (int rc = emit_device(c, ind_first, ind_count, locus_first, locus_first + locus_count, locus_first, t->founder,
                                 t->founder_pitch, fb, (uint8_t*)t->allele, t->allele_pitch, (uint8_t*)t->dose, t->dose_pitch,
                                 dose_mode, dose_which)) return rc;
        CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    } else {
        // Host targets: locus chunks are emitted into one of two device slots and copied out on a
        // second stream while the next chunk is computed. A table whose host pitch is not 16-byte
        // aligned is written with an aligned pitch and repacked on the device (2-D device copy),
        // so that the PCIe transfer is always one large contiguous copy.
        // direct: dense and aligned host rows, the kernel writes the host layout; repack: dense but
        // unaligned; strided: padded host rows (a view into a wider table) are copied row by row
        struct Tab { void* host; int64_t hpitch, width; bool direct; int64_t spitch; bool strided; };
        auto up = [](int64_t x, int64_t m) { return (x + m - 1) / m * m; };
        Tab tab[3] = {{t->founder, t->founder_pitch, n_cols * fb, false, 0, false},
                      {t->allele, t->allele_pitch, n_cols, false, 0, false},
                      {t->dose, t->dose_pitch, n_cols_ind, false, 0, false}};
        int64_t row_bytes = 0;
        for (Tab& x : tab) {
            if (!x.host) continue;
            x.strided = x.hpitch != x.width;
            x.direct = !x.strided && x.width % 16 == 0;
            x.spitch = x.direct ? x.hpitch : up(x.width, 128);
            row_bytes += x.spitch + (x.direct || x.strided ? 0 : x.hpitch);
        }
        const int64_t slot_budget = 96LL << 20;
        const int64_t rows = std::max<int64_t>(1, std::min<int64_t>(locus_count, slot_budget / row_bytes));
        const int64_t slot_bytes = up(rows * row_bytes, 256);
        if (int rc = ensure_staging(c, (size_t)(2 * slot_bytes))) return rc;
        cudaEvent_t ready[2] = {c->ev2, c->ev3};
        cudaEvent_t freed[2] = {next_kev(c), next_kev(c)};
        int k = 0;
        for (int64_t l0 = locus_first; l0 < locus_first + locus_count; l0 += rows, k++) {
            const int64_t l1 = std::min(l0 + rows, locus_first + locus_count);
            const int64_t nr = l1 - l0, r0 = l0 - locus_first;
            const int sl = k & 1;
            uint8_t* base = (uint8_t*)c->staging + sl * slot_bytes;
            uint8_t* sp[3]; uint8_t* dn[3];
            {
                uint8_t* p = base;
                for (int q = 0; q < 3; q++) {
                    sp[q] = dn[q] = nullptr;
                    if (!tab[q].host) continue;
                    sp[q] = p; p += rows * tab[q].spitch;
                    if (!tab[q].direct && !tab[q].strided) { dn[q] = p; p += rows * tab[q].hpitch; }
                }
            }
            if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(c->stream, freed[sl], 0));   // slot copied out
            if (int rc = emit_device(c, ind_first, ind_count, l0, l1, l0, sp[0], tab[0].spitch, fb, sp[1], tab[1].spitch,
                                     sp[2], tab[2].spitch, dose_mode, dose_which)) return rc;
            for (int q = 0; q < 3; q++)
                if (dn[q])
                    CUDA_TRY(cudaMemcpy2DAsync(dn[q], tab[q].hpitch, sp[q], tab[q].spitch, tab[q].width, nr, cudaMemcpyDeviceToDevice, c->stream));
            CUDA_TRY(cudaEventRecord(ready[sl], c->stream));
            CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, ready[sl], 0));
            for (int q = 0; q < 3; q++) {
                if (!tab[q].host) continue;
                uint8_t* dst = (uint8_t*)tab[q].host + r0 * tab[q].hpitch;
                if (tab[q].strided)
                    CUDA_TRY(cudaMemcpy2DAsync(dst, tab[q].hpitch, sp[q], tab[q].spitch, tab[q].width, nr, cudaMemcpyDeviceToHost, c->copy_stream));
                else
                    CUDA_TRY(cudaMemcpyAsync(dst, tab[q].direct ? sp[q] : dn[q], (size_t)(nr * tab[q].width), cudaMemcpyDeviceToHost, c->copy_stream));
            }
            CUDA_TRY(cudaEventRecord(freed[sl], c->copy_stream));
        }
        CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
        CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    float ms_total = 0.f, ms_kernel = 0.f;
    cudaEventElapsedTime(&ms_total, c->ev0, c->ev1);
    const int first_pair = t->memory == PSIM_MEM_DEVICE ? 0 : 2;
    for (int i = first_pair; i + 1 < c->kev_used; i += 2) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->kev[i], c->kev[i + 1]);
        ms_kernel += ms;
    }
    c->stats.ms_emit_kernel = ms_kernel;
    c->stats.ms_emit_total = ms_total;
    c->stats.ms_emit_plan = 0.0;
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
    for (int64_t l0 = locus_first; l0 < locus_first + locus_count; l0 += rows) {
        const int64_t l1 = std::min(l0 + rows, locus_first + locus_count), nr = l1 - l0;
        uint8_t* sp[3];
        {
            uint8_t* p = (uint8_t*)c->staging;
            for (int q = 0; q < 3; q++) { sp[q] = nullptr; if (tab[q].host) { sp[q] = p; p += rows * tab[q].pitch; } }
        }
        CUDA_TRY(cudaStreamSynchronize(c->copy_stream));   // the text of the previous block has left b_text
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
            int64_t total = 0;
            CUDA_TRY(cudaMemcpyAsync(&total, (int64_t*)c->b_rowoff.p + nr, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            if (*tab[q].bytes + total > tab[q].cap) too_small = true;
            if (!too_small) {
                CUDA_TRY(cudaStreamSynchronize(c->copy_stream));   // b_text may grow: nothing may be in flight from it
                if (int rc = ensure_buf(c, c->b_text, (size_t)total)) return rc;
                ta.row_off = (int64_t*)c->b_rowoff.p;
                ta.text = (char*)c->b_text.p;
                k_text_rows<true><<<blocks, TEXT_THREADS, smem, c->stream>>>(ta);
                c->stats.kernel_launches++;
                CUDA_TRY(cudaGetLastError());
                CUDA_TRY(cudaEventRecord(c->ev2, c->stream));
                CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev2, 0));
                CUDA_TRY(cudaMemcpyAsync(tab[q].host + *tab[q].bytes, c->b_text.p, (size_t)total, cudaMemcpyDeviceToHost, c->copy_stream));
                // the next table reuses b_text: its write kernel must wait for this copy
                CUDA_TRY(cudaEventRecord(c->ev3, c->copy_stream));
                CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev3, 0));
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
    const int A = c->founder_allele_count, P = c->P, nd = P / 2 + 1;
    if (A <= 0 || A > 4096) return fail(c, PSIM_EINVAL, "gamete statistics need 1..4096 founder alleles");
    const size_t n_ac = (size_t)A * Lc, n_dc = n_ac * nd, n_diff = (size_t)(Lc + 1) * (Lc + 1);
    const size_t n_words = n_ac + n_dc + (size_t)Lc + (out->recomb_count ? n_diff : 0) + 64 + (size_t)(P / 4 + 1);
    if (int rc = ensure_buf(c, c->b_text, n_words * 4)) return rc;
    uint32_t* w = (uint32_t*)c->b_text.p;
    CUDA_TRY(cudaMemsetAsync(w, 0, n_words * 4, c->stream));
    StatsArgs sa{};
    sa.C = c->C; sa.R = c->R; sa.P = P; sa.A = A; sa.N = c->N;
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
    out->n_gametes = 2 * ind_count * c->R;
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

int psim_reset(psim_ctx* ctx, int64_t seed, uint32_t replicate_first) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    cudaSetDevice(c->cfg.device);
    c->cfg.seed = seed;
    c->cfg.replicate_first = replicate_first;
    c->model.seed = (uint32_t)seed;
    CUDA_TRY(cudaMemsetAsync(c->desc_d, 0, c->n_hom * sizeof(uint64_t), c->stream));
    const int64_t ncfg = (int64_t)c->R * c->N * c->C * 2;
    CUDA_TRY(cudaMemsetAsync(c->cfg_quad_d, 0xFF, ncfg, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->cfg_seq_d, 0xFF, ncfg * c->P, c->stream));
    std::fill(c->hs_reps.begin(), c->hs_reps.end(), 0);
    c->slab_used = 0;
    c->stats.n_segments = 0;
    c->runs_valid = false;
    return PSIM_OK;
}

int psim_set_stream(psim_ctx* ctx, uint64_t cuda_stream) {
    psim_ctx_impl* c = CTX;
    if (!c) return PSIM_EINVAL;
    cudaSetDevice(c->cfg.device);
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
    if (c->R != 1) return fail(c, PSIM_EINVAL, "device import supports replicate_count == 1");
    if (c->N == 0) return fail(c, PSIM_ESTATE, "psim_set_pedigree must be called first");
    if (first < 0 || count <= 0 || first + count > c->N) return fail(c, PSIM_EINVAL, "individual range out of bounds");
    const int64_t nh = count * c->C * c->P;
    if (in->n_homologs != nh) return fail(c, PSIM_EINVAL, "homolog count does not match the individual range");
    if (int rc = ensure_slab(c, c->slab_used + in->n_segments)) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->seg_pos_d + c->slab_used, in->pos, in->n_segments * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->seg_founder_d + c->slab_used, in->founder, in->n_segments * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    k_scatter_desc<<<grid_for(nh, 256), 256, 0, c->stream>>>(nh, first, 0, c->C, c->R, c->N, c->P, in->seg_off, c->slab_used, c->desc_d);
    c->stats.kernel_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->slab_used += in->n_segments;
    c->stats.n_segments = c->slab_used;
    for (int64_t i = first; i < first + count; i++) c->hs_reps[i] = c->R;
    c->runs_valid = false;
    return PSIM_OK;
}

}  // extern "C"
