// K2 (tiled path) — tile plan, streaming tiles, dense units
// (part of libpsim: included by psim.cu inside namespace psim)
#pragma once

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

