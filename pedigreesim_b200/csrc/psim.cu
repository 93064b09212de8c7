// libpsim — B200 (sm_100a) meiosis + genotype emission behind the C ABI of include/psim.h.
//
// Device layout (DESIGN.md "Data layout in HBM"):
//   homolog id  hid = ((c * R + r) * N + i) * P + h      chromosome-major, so that the columns
//                                                         of one emission tile are contiguous
//   hom_desc[hid]  = begin (40 bit) | count << 40         into the segment slab
//   seg_pos[] f64 Morgan / seg_founder[] i32              HaploStruct.recombPos / founder, unmerged
//   run_desc / run_lb[] u32 / run_founder[] u16           K0 output: per homolog the runs of loci
//                                                         that carry one founder (locus-index space)
// Kernels: k_init_founders, k_meiosis_plan (K1a), k_meiosis_assemble (K1b), k_index_runs (K0),
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
};

__global__ void __launch_bounds__(128) k_meiosis_plan(PlanArgs a) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)a.C * a.R * a.n_lev * 2;
    if (t >= total) return;
    const int par = (int)(t & 1);
    const int64_t unit = t >> 1;
    const int j = (int)(unit % a.n_lev);
    const int r = (int)((unit / a.n_lev) % a.R);
    const int c = (int)(unit / ((int64_t)a.n_lev * a.R));
    const int P = a.P, p2 = P / 2;
    const int32_t child = a.lev_ind[j];
    const int32_t parent = par == 0 ? a.par0[child] : a.par1[child];
    const ChromParams ch = a.chrom[c];
    const ModelParams& m = a.model;

    HomologView hv[MAXP];
    const int64_t phid0 = (((int64_t)c * a.R + r) * a.N + parent) * P;
    for (int h = 0; h < P; h++) {
        const uint64_t d = a.desc[phid0 + h];
        hv[h].pos = a.seg_pos + desc_begin(d);
        hv[h].founder = a.seg_founder + desc_begin(d);
        hv[h].n = (int)desc_count(d);
    }
    PhiloxStream rng;
    rng.init(m.seed, a.replicate_first + (uint32_t)r);
    PairingState ps;
    rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, STAGE_CONFIG);
    calc_chrom_config(ps, rng, m, ch, hv);
    {
        const int64_t q = (((int64_t)r * a.N + child) * a.C + c) * 2 + par;
        a.cfg_quad[q] = (int8_t)ps.quad;
        for (int h = 0; h < P; h++) a.cfg_seq[q * P + h] = ps.chromseq[h];
    }
    // gamete 0 before the homolog shuffle: slot -> (multivalent base in chromseq, starting chromatid slot, stage)
    Chiasmata x;
    x.overflow = false;
    // Multivalents are processed one at a time; each transmitted chromatid is traced straight into
    // the plan entry of the child homolog it will occupy after the gamete's homolog shuffle.
    const int64_t u0 = (((int64_t)c * a.R + r) * a.n_lev + j) * P + (int64_t)par * p2;
    // The shuffle draws come from their own stream, so take them first: slot h of the transmitted
    // gamete receives pre-shuffle slot shuf[h] (Individual.java:325-335).
    int8_t shuf[MAXP / 2];
    for (int s = 0; s < p2; s++) shuf[s] = (int8_t)s;
    rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, STAGE_SHUFFLE);
    shuffle_small(rng, shuf, p2);
    int8_t dest_of_slot[MAXP / 2];  // pre-shuffle slot -> final homolog position
    for (int h = 0; h < p2; h++) dest_of_slot[shuf[h]] = (int8_t)h;

    uint32_t stage = 1;
    const int firstbi = ps.quad * 4;
    const int n_mv = ps.quad + (P - firstbi) / 2;
    for (int v = 0; v < n_mv; v++) {
        const bool is_quad = v < ps.quad;
        const int k = is_quad ? 4 : 2;
        const int base = is_quad ? 4 * v : firstbi + 2 * (v - ps.quad);
        rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, stage++);
        double exchange_lim1 = 0.0;
        if (is_quad) exchange_lim1 = quadrivalent_crossing_over(x, rng, m, ch);
        else parallel_crossing_over(x, rng, m, ch, 2);
        if (x.overflow) { atomicExch(a.error_flag, PSIM_ECAPACITY); return; }
        int8_t centro[8];
        if (is_quad) quadrivalent_centromere_order(rng, m, ch, exchange_lim1, centro);
        else multivalent_centromere_order(rng, 2, centro);
        int start_slots[2];
        select_gamete0(x, rng, k, ch, centro, start_slots);
        const int n_out = is_quad ? 2 : 1;
        for (int o = 0; o < n_out; o++) {
            // Individual.java:297-321: quadrivalent v fills gamete slots 2v, 2v+1; bivalent v' fills firstbi/2 + v'
            const int pre_slot = is_quad ? 2 * v + o : firstbi / 2 + (v - ps.quad);
            const int64_t u = u0 + dest_of_slot[pre_slot];
            double* ppos = a.plan_pos + u * a.piece_cap;
            uint8_t* phom = a.plan_hom + u * a.piece_cap;
            // Multivalent.slotsToPatterns (:332-357) for the selected chromatid
            int cur = start_slots[o];
            int np = 0;
            ppos[0] = ch.head;
            phom[0] = (uint8_t)ps.chromseq[base + (cur >> 1)];
            np = 1;
            bool over = false;
            for (int e = 0; e < x.n; e++) {
                int nxt = -1;
                if (x.a[e] == cur) nxt = x.b[e];
                else if (x.b[e] == cur) nxt = x.a[e];
                if (nxt >= 0) {
                    if (np >= a.piece_cap) { over = true; break; }
                    ppos[np] = x.pos[e];
                    phom[np] = (uint8_t)ps.chromseq[base + (nxt >> 1)];
                    np++;
                    cur = nxt;
                }
            }
            if (over) { atomicExch(a.error_flag, PSIM_ECAPACITY); return; }
            a.plan_np[u] = np;
            // Multivalent.fillGamete (:539-577): number of segments the chromatid will have
            int64_t cnt = 0;
            if (np == 1) {
                cnt = hv[phom[0]].n;
            } else {
                for (int q = 0; q < np; q++) {
                    const HomologView& h = hv[phom[q]];
                    const double patstart = ppos[q];
                    const double patend = q < np - 1 ? ppos[q + 1] : ch.tail;
                    int seg = h.find_segment(patstart) + 1;
                    cnt++;
                    while (seg < h.n && h.pos[seg] < patend) { cnt++; seg++; }
                }
            }
            a.plan_cnt[u] = cnt;
        }
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
    const uint8_t* code;          // [L][A] (MODE 3)
    const uint8_t* match_code;    // [L] code to count (MODE 3)
    int A;
};

#ifndef PSIM_EMIT_MINB
#define PSIM_EMIT_MINB 6
#endif
constexpr int EMIT_THREADS = 128;
constexpr int EMIT_MAX_TILE_LOCI = 256;   // row index inside a tile must fit 16 bits
constexpr int EMIT_EV = 64;           // run starts a warp may have inside a tile on the fast path (two per lane)
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

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// K2b — streaming. A block's four warps take the same tile (2 KB contiguous per row written at the
// same time), tiles are handed out in row-band order by an atomic ticket requested a tile early;
// the plan block of the next tile is copied into shared memory with cp.async while the current
// one streams, so no global round trip is ever waited for (under a saturated write stream each
// costs thousands of cycles). Each warp sorts its run starts once (bitonic, one entry per lane)
// and then walks them with warp-uniform control flow: write rows without any per-row checks up to
// the next run start, let the owning lane patch its registers, go on.
template <int P, int G, int MODE>
__global__ void __launch_bounds__(EMIT_THREADS, PSIM_EMIT_MINB) k_emit_tiles(const EmitArgs a) {
    constexpr int NC = G * P;
    static_assert(NC % 16 == 0, "a thread's cells must be whole 16-byte vectors");
    constexpr int NW = NC / 4;          // 32-bit words of packed cells
    constexpr int NG = NC / 16;         // 16-byte vectors of packed cells
    constexpr int DW = (G + 3) / 4;     // 32-bit words of packed doses
    constexpr int W = MODE == 2 ? 2 : 1;  // nibble words per individual (8 founder alleles each)
    constexpr int SW = MODE == 0 ? DW : MODE == 3 ? 1 : G * W;   // words of derived per-thread state
    constexpr int PB = 2 + NG * EMIT_THREADS + EMIT_EV;          // uint4 per plan block
    __shared__ uint4 s_blk[2][PB];
    __shared__ int s_ticket[2];
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    const bool want_f = a.founder != nullptr, want_a = a.allele != nullptr, want_d = a.dose != nullptr;
    const int64_t n_cols = a.n_cols_ind * P;

    auto prefetch_plan = [&](int tile, int buf) {
        const uint4* src = a.plan_blk + (int64_t)tile * PB;
        for (int i = tid; i < PB; i += EMIT_THREADS) cp_async16(&s_blk[buf][i], src + i);
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
        if (n_ev > (uint32_t)EMIT_EV) {   // sparse map: this (tile, warp) unit goes to k_emit_dense
            if (lane == 0) a.dense_list[1 + atomicAdd(a.dense_list, 1)] = tile_cur * (EMIT_THREADS / 32) + wq;
            continue;
        }
        const uint4 tc = s_blk[cb][1];     // tile coordinates left by k_emit_plan
        const uint32_t l0 = tc.x, l1 = tc.y;
        const int64_t row0 = (int64_t)tc.w;
        const int64_t c_off = row0 + a.locus_first - l0;
        const int64_t q0 = ((int64_t)tc.z * EMIT_THREADS + tid) * G;   // first emitted individual of this lane
        const int64_t colf = q0 * P;                                    // first founder/allele column
        const bool full = colf + NC <= n_cols;

        // ---- the warp's run starts, sorted by row: element i = lane + 32 r lives in ev[r] of the
        // lane, bitonic network over the 64 slots (unused ones hold NO_EVENT and sort to the end)
        uint32_t ev[2];
        {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&s_blk[cb][2 + NG * EMIT_THREADS]) + wq * EMIT_EV;
            ev[0] = (uint32_t)lane < n_ev ? src[lane] : NO_EVENT;
            ev[1] = (uint32_t)lane + 32 < n_ev ? src[lane + 32] : NO_EVENT;
        }
        if (n_ev > 1) {
#pragma unroll
            for (int k2 = 2; k2 <= 64; k2 <<= 1) {
#pragma unroll
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    if (j == 32) {   // partner is the lane's other register
                        const uint32_t lo_v = min(ev[0], ev[1]), hi_v = max(ev[0], ev[1]);
                        ev[0] = lo_v; ev[1] = hi_v;     // k2 == 64: the whole array ascends
                    } else {
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            const int i = lane + 32 * r;
                            const uint32_t other = __shfl_xor_sync(0xFFFFFFFFu, ev[r], j);
                            const bool keep_min = ((i & k2) == 0) == ((i & j) == 0);
                            ev[r] = keep_min ? min(ev[r], other) : max(ev[r], other);
                        }
                    }
                }
            }
        }
        // ---- state of this lane at the tile's first row
        uint32_t pk[NW];
#pragma unroll
        for (int cg = 0; cg < NG; cg++) {
            const uint4 v = s_blk[cb][2 + cg * EMIT_THREADS + tid];
            pk[4 * cg] = v.x; pk[4 * cg + 1] = v.y; pk[4 * cg + 2] = v.z; pk[4 * cg + 3] = v.w;
        }
        uint32_t st[SW];           // MODE 0: packed doses; MODE 1/2: nibble multiplicities [g][w]
        uint32_t sel[NW];          // MODE 1/2 with allele table: PRMT selectors
        uint32_t him[NW];          // MODE 2: 0xFF per cell whose founder allele is >= 8
#pragma unroll
        for (int w = 0; w < SW; w++) st[w] = 0;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const uint32_t f = (pk[j / 4] >> (8 * (j % 4))) & 0xFF;
            if (MODE == 0) { if (f == (uint32_t)a.dose_which) st[(j / P) / 4] += 1u << (8 * ((j / P) % 4)); }
            else if (MODE == 1) st[j / P] += 1u << (4 * (f & 7));
            else if (MODE == 2) st[(j / P) * W + (f >> 3 & 1)] += 1u << (4 * (f & 7));
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int w = 0; w < NW; w++) {
                const uint32_t v = pk[w];
                sel[w] = (v & 0x7) | ((v >> 4) & 0x70) | ((v >> 8) & 0x700) | ((v >> 12) & 0x7000);
                him[w] = ((v >> 3) & 0x01010101u) * 0xFFu;
            }
        }

        // ---- stream the rows
        uint32_t l = l0;
        uint8_t* pf = a.founder + row0 * a.founder_pitch + colf;
        uint8_t* pa = a.allele + row0 * a.allele_pitch + colf;
        uint8_t* pd = a.dose + row0 * a.dose_pitch + q0;
        const uint4* code16 = reinterpret_cast<const uint4*>(a.code16) + c_off;    // per-locus tables: warp-uniform
        const uint2* nibmask = reinterpret_cast<const uint2*>(a.nibmask) + c_off;  // loads through L1
        // compiled once per combination of requested tables so that the hot loop carries no
        // predicated-off work (WF / WA / WD shadow the runtime flags)
        auto emit_rows_t = [&](auto full_tag, auto wf_tag, auto wa_tag, auto wd_tag, uint32_t stop) {
            constexpr bool FULL = decltype(full_tag)::value;
            constexpr bool want_f = decltype(wf_tag)::value, want_a = decltype(wa_tag)::value, want_d = decltype(wd_tag)::value;
#pragma unroll 4
            for (; l < stop; l++) {
                uint32_t dose_pk[DW];
                if (want_f) { store_cells<NW>(pf, pk, FULL, colf, n_cols); pf += a.founder_pitch; }
                if (MODE == 0) {
#pragma unroll
                    for (int w = 0; w < DW; w++) dose_pk[w] = st[w];
                } else if (MODE == 1 || MODE == 2) {
                    if (want_a) {
                        const uint4 tb = __ldg(code16 + l);
                        uint32_t ak[NW];
#pragma unroll
                        for (int w = 0; w < NW; w++) {
                            const uint32_t lov = __byte_perm(tb.x, tb.y, sel[w]);
                            if (MODE == 2) { const uint32_t hiv = __byte_perm(tb.z, tb.w, sel[w]); ak[w] = (lov & ~him[w]) | (hiv & him[w]); }
                            else ak[w] = lov;
                        }
                        store_cells<NW>(pa, ak, FULL, colf, n_cols);
                        pa += a.allele_pitch;
                    }
                    if (want_d) {
                        uint2 m;
                        if (W == 2) m = __ldg(nibmask + l);
                        else { m.x = __ldg(reinterpret_cast<const uint32_t*>(nibmask + l)); m.y = 0; }
#pragma unroll
                        for (int w = 0; w < DW; w++) dose_pk[w] = 0;
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            uint32_t d = ((st[g * W] & m.x) * 0x11111111u) >> 28;
                            if (W == 2) d += ((st[g * W + W - 1] & m.y) * 0x11111111u) >> 28;
                            dose_pk[g / 4] |= d << (8 * (g % 4));
                        }
                    }
                } else {
                    const uint8_t* codes = a.code + (c_off + l) * a.A;
                    const uint32_t match = a.match_code[c_off + l];
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
                    if (want_a) { store_cells<NW>(pa, ak, FULL, colf, n_cols); pa += a.allele_pitch; }
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
                    pd += a.dose_pitch;
                }
            }
        };
        auto emit_rows = [&](auto full_tag, uint32_t stop) {
            using T = std::true_type;
            using F = std::false_type;
            if (MODE == 0) {   // no allele table without codes
                if (want_f) { if (want_d) emit_rows_t(full_tag, T{}, F{}, T{}, stop); else emit_rows_t(full_tag, T{}, F{}, F{}, stop); }
                else emit_rows_t(full_tag, F{}, F{}, T{}, stop);
            } else if (want_f) {
                if (want_a) { if (want_d) emit_rows_t(full_tag, T{}, T{}, T{}, stop); else emit_rows_t(full_tag, T{}, T{}, F{}, stop); }
                else { if (want_d) emit_rows_t(full_tag, T{}, F{}, T{}, stop); else emit_rows_t(full_tag, T{}, F{}, F{}, stop); }
            } else {
                if (want_a) { if (want_d) emit_rows_t(full_tag, F{}, T{}, T{}, stop); else emit_rows_t(full_tag, F{}, T{}, F{}, stop); }
                else emit_rows_t(full_tag, F{}, F{}, T{}, stop);
            }
        };
        for (uint32_t k = 0;; k++) {   // warp-uniform walk over the sorted run starts
            const uint32_t e = k < n_ev ? __shfl_sync(0xFFFFFFFFu, k < 32 ? ev[0] : ev[1], k & 31) : NO_EVENT;
            const uint32_t wstop = min(l0 + (e >> 22), l1);
            if (full) emit_rows(std::true_type{}, wstop);
            else if (colf < n_cols) emit_rows(std::false_type{}, wstop);
            else l = wstop;
            if (e == NO_EVENT) break;
            if (((e >> 15) & 31) == (uint32_t)lane) {      // this lane's run starts here: patch the registers
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
                } else if (MODE == 1) {
                    const uint32_t delta = (1u << (4 * (f & 7))) - (1u << (4 * (old & 7)));
#pragma unroll
                    for (int gg = 0; gg < G; gg++) st[gg] += (uint32_t)gg == g ? delta : 0u;
                } else if (MODE == 2) {
                    const uint32_t dec = 1u << (4 * (old & 7)), inc = 1u << (4 * (f & 7));
                    const uint32_t wi_old = g * W + (old >> 3 & 1), wi_new = g * W + (f >> 3 & 1);
#pragma unroll
                    for (int w = 0; w < SW; w++) st[w] += ((uint32_t)w == wi_new ? inc : 0u) - ((uint32_t)w == wi_old ? dec : 0u);
                }
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
    void* founder; int64_t founder_pitch; int founder_bytes;
    uint8_t* allele; int64_t allele_pitch;
    uint8_t* dose; int64_t dose_pitch;
    int dose_mode, dose_which;
    const uint8_t* code; const uint8_t* match_code; int A;
};

__global__ void __launch_bounds__(256) k_emit_generic(const EmitGenericArgs a) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= a.l_count * a.n_cols_ind) return;
    const int64_t q = t % a.n_cols_ind;
    const int64_t gl = a.l_begin + t / a.n_cols_ind;
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
                              uint32_t* __restrict__ code16, uint32_t* __restrict__ nibmask) {
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
    Buf b_plan, b_dense_list;
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
    cudaFree(c->code16_d); cudaFree(c->nibmask_d);
    cudaFree(c->error_flag_d); cudaFree(c->tile_counter_d); cudaFree(c->cub_tmp); cudaFree(c->staging);
    cudaFree(c->exp_off_d); cudaFree(c->exp_founder_d); cudaFree(c->exp_pos_d);
    for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
    for (psim_ctx_impl::Buf* b : {&c->b_lev, &c->b_np, &c->b_cnt, &c->b_off, &c->b_ppos, &c->b_phom, &c->b_find, &c->b_fal,
                                 &c->b_plan, &c->b_dense_list}) cudaFree(b->p);
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
            k_meiosis_plan<<<grid_for(n_half, 128), 128, 0, c->stream>>>(pa);
            c->stats.kernel_launches++;
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
    cudaFree(c->code16_d); cudaFree(c->nibmask_d);
    c->locus_off_d = nullptr; c->locus_pos_d = nullptr; c->code_d = nullptr; c->match_d = nullptr;
    c->code16_d = nullptr; c->nibmask_d = nullptr;
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
    cudaFree(c->code_d); cudaFree(c->match_d); cudaFree(c->code16_d); cudaFree(c->nibmask_d);
    c->code_d = nullptr; c->match_d = nullptr; c->code16_d = nullptr; c->nibmask_d = nullptr;
    c->A = 0; c->code_h.clear(); c->match_which = 12345678;
    if (!code) return PSIM_OK;
    if (n_alleles <= 0 || n_alleles > 256) return fail(c, PSIM_EINVAL, "allele codes need 1..256 founder alleles");
    c->A = n_alleles;
    c->code_h.assign(code, code + (size_t)c->L * n_alleles);
    CUDA_TRY(dev_alloc(&c->code_d, (size_t)c->L * n_alleles));
    CUDA_TRY(dev_alloc(&c->match_d, (size_t)c->L));
    CUDA_TRY(dev_alloc(&c->code16_d, (size_t)c->L * 4));
    CUDA_TRY(dev_alloc(&c->nibmask_d, (size_t)c->L * 2));
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
    const bool tiled = G != 0 && aligned && founder_bytes == 1 && c->founder_allele_count <= 256 &&
                       c->run_cap < (int64_t)0xFFFFFFFF;
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
            while (rows > 8 && 32.0 * G * P * rows * starts_per_cell > 24.0) rows >>= 1;
            ea.tile_loci = env_tile > 0 ? std::min(env_tile, EMIT_MAX_TILE_LOCI) : rows;
        }
        ea.n_col_tiles = (int)((n_cols_ind + (int64_t)EMIT_THREADS * G - 1) / ((int64_t)EMIT_THREADS * G));
        ea.founder = (uint8_t*)founder; ea.founder_pitch = founder_pitch;
        ea.allele = allele; ea.allele_pitch = allele_pitch;
        ea.dose = dose; ea.dose_pitch = dose_pitch;
        ea.dose_which = dose_which;
        ea.code = c->code_d; ea.match_code = c->match_d; ea.A = c->A;
        ea.code16 = c->code16_d; ea.nibmask = c->nibmask_d;
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
                                                                       c->code16_d, c->nibmask_d);
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
