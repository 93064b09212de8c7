// K2 (row-image path) — band plan and shared-memory row + TMA bulk store emission
// (part of libpsim: included by psim.cu inside namespace psim)
#pragma once

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
    uint8_t* plan;                // per unit: [count u32, 12 bytes pad][seg_cells state bytes][cap run starts u32][cap distances u8]
    int64_t plan_stride;
    int* unit_counter;            // ticket
    int* overflow;                // [0] = number of units whose list overflowed, then their ids
    uint8_t* founder; int64_t founder_pitch;
    uint8_t* dose;    int64_t dose_pitch;
    int dose_which;               // without genotypes: founder allele to count
    const uint4* mtab;            // with genotypes: per-locus byte tables of the counted allele (k_match_codes)
    int plan_shared;              // k_band_plan: count a block's run starts per band in shared memory, one global atomic per band
};
constexpr int ROWS_MAX_PLAN_BANDS = 2048;   // bands of one chromosome piece in one launch (more: the launch is cut)
__device__ __forceinline__ void band_rows(const RowsArgs& a, int k, int j, int64_t& g0, int64_t& g1) {
    const int64_t len = a.chunk_l1[k] - a.chunk_l0[k];
    g0 = a.chunk_l0[k] + len * j / a.chunk_bands[k];
    g1 = a.chunk_l0[k] + len * (j + 1) / a.chunk_bands[k];
}

// One thread per (chromosome piece, emitted cell column): walks the homolog's runs down the piece and leaves the state
// byte of every band plus the run starts inside it. The band bounds are computed once per block and the column
// arithmetic is 32-bit (64-bit divisions per thread and band were a tenth of this kernel's time). Filing a run start
// needs a place in its unit's list: a global atomic per run start was over half of the kernel's time (ncu: the threads
// wait for the counter's old value), so a block counts its run starts per band in shared memory on a first walk,
// reserves them with ONE global atomic per band and segment, and files them on a second walk (the loads hit L1 / L2).
// Pieces of more than ROWS_PLAN_SHARED_BANDS bands keep the atomic per run start.
constexpr int ROWS_PLAN_SHARED_BANDS = 512;
struct PlanWalk {   // a homolog's runs from the first row of the piece on
    uint64_t b; uint32_t n, kx, nxt, f;
};
__global__ void __launch_bounds__(256) k_band_plan(const RowsArgs a) {
    __shared__ uint32_t s_tl[ROWS_MAX_PLAN_BANDS + 1];          // first row of every band, relative to the chromosome
    __shared__ uint32_t s_cnt[2][ROWS_PLAN_SHARED_BANDS];        // [segment of the block][band]: run starts, then the cursor
    __shared__ uint32_t s_base[2][ROWS_PLAN_SHARED_BANDS];       // first entry reserved in the unit's list
    const int k = blockIdx.y;
    const int n_bands = a.chunk_bands[k];
    const bool shared_counts = a.plan_shared && n_bands <= ROWS_PLAN_SHARED_BANDS;
    const int64_t c_off = a.chunk_coff[k];
    for (int j = threadIdx.x; j <= n_bands; j += blockDim.x) {
        int64_t g0, g1;
        band_rows(a, k, min(j, n_bands - 1), g0, g1);
        s_tl[j] = (uint32_t)((j < n_bands ? g0 : g1) - c_off);
    }
    if (shared_counts) for (int j = threadIdx.x; j < 2 * n_bands; j += blockDim.x) s_cnt[j / n_bands][j % n_bands] = 0;
    __syncthreads();
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const uint32_t P = (uint32_t)a.P;
    const bool live = col < a.n_cols_ind * P;   // (threads past the last column still reach the barriers)
    const int c = a.chunk_chrom[k];
    const uint32_t col32 = live ? (uint32_t)col : 0u;   // (the row-image path is taken for fewer than 2^31 cell columns)
    const uint32_t q = col32 / P, h = col32 - q * P;
    const uint32_t r = q / (uint32_t)a.ind_count;
    const int64_t i = a.ind_first + (q - r * (uint32_t)a.ind_count);
    const int64_t hid = (((int64_t)c * a.R + r) * a.N + i) * P + h;
    const uint32_t s = col32 / (uint32_t)a.seg_cells;
    const uint32_t cs = col32 - s * (uint32_t)a.seg_cells;
    const uint32_t s_first = (uint32_t)((blockIdx.x * (int64_t)blockDim.x) / a.seg_cells);   // (a block spans at most two segments)
    const uint32_t sl = s - s_first;
    PlanWalk w0{0, 0, 1, NO_EVENT, 0};
    if (live) {
        const uint64_t d = __ldg(a.run_desc + hid);
        w0.b = desc_begin(d);   // (40 bits: the slab may hold more than 2^32 segments)
        w0.n = desc_count(d);
        const uint32_t l_begin = s_tl[0];
        uint32_t lo = 0, hi = w0.n;   // last run with lb <= l_begin
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(a.run_lb + w0.b + mid) <= l_begin) lo = mid; else hi = mid; }
        w0.f = w0.n ? __ldg(a.run_founder + w0.b + lo) : 0;
        w0.kx = lo + 1;
        w0.nxt = w0.kx < w0.n ? __ldg(a.run_lb + w0.b + w0.kx) : NO_EVENT;
    }
    auto advance = [&](PlanWalk& w) {
        w.f = __ldg(a.run_founder + w.b + w.kx);
        w.kx++;
        w.nxt = w.kx < w.n ? __ldg(a.run_lb + w.b + w.kx) : NO_EVENT;
    };
    uint8_t* const pu0 = a.plan + (int64_t)(a.chunk_unit_base[k] + s) * a.plan_stride;
    const int64_t band_stride = (int64_t)a.n_seg * a.plan_stride;
    auto file = [&](uint8_t* pu, uint32_t idx, uint32_t row, uint32_t f, uint32_t again) {
        if (idx < (uint32_t)a.cap) {
            reinterpret_cast<uint32_t*>(pu + 16 + a.seg_cells)[idx] = (row << 24) | (cs << 8) | f;
            // rows until the same cell changes again (255: not soon): a row buffer that takes the run starts
            // of several rows in one pass skips those that are overwritten inside its window
            (pu + 16 + a.seg_cells + (size_t)a.cap * 4)[idx] = (uint8_t)min(again, 255u);
        }
    };
    // ---- first walk: the state byte of every band; the run starts are counted (or filed, in a piece of many bands)
    if (live) {
        PlanWalk w = w0;
        uint8_t* pu = pu0;
        for (int j = 0; j < n_bands; j++, pu += band_stride) {
            const uint32_t tl0 = s_tl[j], tl1 = s_tl[j + 1];
            while (w.nxt <= tl0) advance(w);   // a run starting exactly on the band's first row is its start state
            pu[16 + cs] = (uint8_t)w.f;
            uint32_t cnt = 0;
            while (w.nxt < tl1) {
                const uint32_t row = w.nxt - tl0;
                advance(w);
                if (shared_counts) cnt++;
                else file(pu, atomicAdd(reinterpret_cast<uint32_t*>(pu), 1u), row, w.f, w.nxt - (row + tl0));
            }
            if (cnt) atomicAdd(&s_cnt[sl][j], cnt);
        }
    }
    if (!shared_counts) return;   // (the same for every thread of the block)
    __syncthreads();
    // ---- one reservation per band and segment of the block
    for (int t = threadIdx.x; t < 2 * n_bands; t += blockDim.x) {
        const int sg = t / n_bands, j = t % n_bands;
        const uint32_t cnt = s_cnt[sg][j];
        if (cnt) {
            uint8_t* pu = a.plan + (int64_t)(a.chunk_unit_base[k] + j * a.n_seg + (int)s_first + sg) * a.plan_stride;
            s_base[sg][j] = atomicAdd(reinterpret_cast<uint32_t*>(pu), cnt);
            s_cnt[sg][j] = 0;
        }
    }
    __syncthreads();
    // ---- second walk: the run starts are filed
    if (live) {
        PlanWalk w = w0;
        uint8_t* pu = pu0;
        for (int j = 0; j < n_bands; j++, pu += band_stride) {
            const uint32_t tl0 = s_tl[j], tl1 = s_tl[j + 1];
            while (w.nxt <= tl0) advance(w);
            while (w.nxt < tl1) {
                const uint32_t row = w.nxt - tl0;
                advance(w);
                file(pu, s_base[sl][j] + atomicAdd(&s_cnt[sl][j], 1u), row, w.f, w.nxt - (row + tl0));
            }
        }
    }
}

// Before every launch: run-start counters of the units, overflow count, unit ticket.
__global__ void k_rows_reset(uint8_t* plan, int64_t plan_stride, int n_units, int* overflow, int* ticket) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n_units) *reinterpret_cast<uint4*>(plan + (int64_t)u * plan_stride) = make_uint4(0, 0, 0, 0);
    else if (u == n_units) *overflow = 0;
    else if (u == n_units + 1) *ticket = 0;
}

#ifndef PSIM_ROWS_COMPUTE
#define PSIM_ROWS_COMPUTE 384
#endif
constexpr int ROWS_COMPUTE = PSIM_ROWS_COMPUTE;   // threads that make the rows
constexpr int ROWS_MAX_BAND = 255;     // the row of a run start inside its band is 8 bits
constexpr int ROWS_SMEM = 220 * 1024;  // one block per SM
#ifndef PSIM_ROWS_NBUF
#define PSIM_ROWS_NBUF 3
#endif
constexpr int ROWS_NBUF = PSIM_ROWS_NBUF;   // row buffers in flight
#ifndef PSIM_DOSE_DIRECT
#define PSIM_DOSE_DIRECT 0
#endif
constexpr bool ROWS_DOSE_DIRECT = PSIM_DOSE_DIRECT != 0;   // recomputed dose rows go from registers to global memory, not through a row buffer
#ifndef PSIM_ROWS_INFLIGHT
#define PSIM_ROWS_INFLIGHT 0
#endif
// 0: one issuing warp per row buffer (every buffer's row is handed to the bulk-copy engine as soon as it is made);
// k > 0: ONE issuing warp that keeps at most k rows in the engine (it hands over row r when row r - k has been read)
constexpr int ROWS_INFLIGHT = PSIM_ROWS_INFLIGHT;
static_assert(ROWS_INFLIGHT <= ROWS_NBUF, "a row in the engine occupies a buffer");
constexpr int ROWS_ISSUE_WARPS = ROWS_INFLIGHT > 0 ? 1 : ROWS_NBUF;
constexpr int ROWS_THREADS = ROWS_COMPUTE + 32 * ROWS_ISSUE_WARPS;   // + the warps that hand rows to the bulk-copy engine
#ifndef PSIM_ROWS_GROUPS
#define PSIM_ROWS_GROUPS 2
#endif
// The compute threads work as ROWS_GROUPS independent groups, group g making rows g, g + GROUPS, ...: making a row is
// a chain of latencies (wait for the buffer, dependent shared-memory patches, a barrier, the dose pass, a fence), so
// two groups half a row apart keep the bulk-copy engine fed where one group of twice the threads leaves it idle
// between rows. A group owns the 4-bit image of its row parity; the row buffers go round all groups.
constexpr int ROWS_GROUPS = PSIM_ROWS_GROUPS;
static_assert(ROWS_GROUPS == 1 || ROWS_GROUPS == 2, "the 4-bit image is kept per row parity");
constexpr int ROWS_GROUP_THREADS = ROWS_COMPUTE / ROWS_GROUPS;

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
// named barriers (0 is __syncthreads): FULL + b / EMPTY + b for row buffer b, WORK among the compute threads
// (EMPTY + b + NBUF * g: group g waits for buffer b - two groups may wait for the same buffer at the same time, for rows
// NBUF apart; WORK + g among the threads of group g)
constexpr int BAR_FULL = 1, BAR_EMPTY = 1 + ROWS_NBUF, BAR_WORK = 1 + ROWS_NBUF + ROWS_NBUF * ROWS_GROUPS;
static_assert(BAR_WORK + ROWS_GROUPS <= 16, "named barriers");
__device__ __forceinline__ int bar_empty(int row) { return BAR_EMPTY + row % ROWS_NBUF + ROWS_NBUF * (row % ROWS_GROUPS); }   // waited for by the group that makes `row`
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// The run starts of a unit, bucketed by row (counting sort by all threads of the block): evs / evd sorted by row,
// start[i] = first entry of row i. (A histogram per warp instead of one for the block - fewer threads on a bin - was
// tried for bands of few rows and many run starts and is slower: profiles/README.md, round 2.)
__device__ __forceinline__ void bucket_run_starts(const uint8_t* pu, int seg_cells, int cap, uint32_t n_ev, int n_rows, uint32_t* evs, uint8_t* evd,
                                                  uint32_t* start, uint32_t* cursor) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const uint32_t* ev_g = reinterpret_cast<const uint32_t*>(pu + 16 + seg_cells);
    const uint8_t* evd_g = pu + 16 + seg_cells + (size_t)cap * 4;
    for (int i = tid; i <= n_rows; i += NT) start[i] = 0;
    __syncthreads();
    for (uint32_t e = tid; e < n_ev; e += NT) atomicAdd(&start[__ldg(ev_g + e) >> 24], 1u);
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
    for (uint32_t e = tid; e < n_ev; e += NT) {
        const uint32_t w = __ldg(ev_g + e);
        const uint32_t at = atomicAdd(&cursor[w >> 24], 1u);
        evs[at] = w;
        evd[at] = __ldg(evd_g + e);
    }
}

// DK: how the dose row is made. 0 = no dose table, 1 = no genotypes (count of one founder allele:
// patched like the founder row), 2 = genotypes, A <= 8, ploidy 4, 3 = genotypes, A <= 8, ploidy 2.
// Warp-specialised: 512 threads make row r in buffer r % NBUF; every buffer has a warp of its own that
// hands its rows to the bulk-copy engine and waits until they have been read (issuing a bulk store
// keeps the issuing warp busy for 800-1600 cycles whatever the size of the store).
// Making a row is latency, not arithmetic: every phase that ends in a block-wide barrier costs
// 300-500 cycles (dependent shared-memory loads, then the slowest of 16 warps), and a row may take
// ~2100 cycles before HBM is the limit. So a row has as few phases as possible:
//   1. when the engine has read the buffer's previous row: ONE pass patches the run starts of all the
//      rows the buffer missed (k_band_plan marks the starts that are overwritten again inside such a
//      window, so the pass touches every cell at most once and needs no order);
//   2. with genotypes: barrier, then the dose row is recomputed from a 4-bit image of the row (kept
//      twice, for even and odd rows, patched the same way) and goes from registers straight to global
//      memory; without genotypes the dose row is patched like the founder row and leaves as a bulk store.
// Sized to share an SM with the meiosis kernel: 384 compute threads + 3 issuing warps at 40 registers (the compiler
// takes 87 when left alone, with no gain) leave room for three blocks of k_meiosis at 56 registers, and the pipelined
// step - population k + 1 simulated while population k is emitted - is 0.40 ms instead of 0.45 (512 threads at 87
// registers next to k_meiosis at 64: whichever kernel was resident first left the other next to nothing). The
// kernel's own time does not change (configs[1]); profiles/README.md has the variants.
#ifndef PSIM_ROWS_MAXREG
#define PSIM_ROWS_MAXREG 40
#endif
#define ROWS_KERNEL_BOUNDS __maxnreg__(PSIM_ROWS_MAXREG)
template <int DK>
__global__ void ROWS_KERNEL_BOUNDS k_emit_rows(const RowsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int NT = ROWS_THREADS;
    const int P = a.P;
    const int seg_cells = a.seg_cells, seg_ind = seg_cells / P;
    constexpr int NB = ROWS_NBUF;
    uint8_t* const img0 = smem;                                              // founder byte of every cell on the current row (x NB)
    uint32_t* const nib0 = reinterpret_cast<uint32_t*>(img0 + (size_t)NB * seg_cells);   // DK >= 2: the same, 4 bits per cell (x 2)
    constexpr bool DSM = DK == 1 || (DK >= 2 && !ROWS_DOSE_DIRECT);          // dose rows in shared memory, sent as bulk stores
    uint8_t* const dose0 = reinterpret_cast<uint8_t*>(nib0) + (DK >= 2 ? seg_cells : 0);   // dose rows (x NB)
    uint32_t* const evs = reinterpret_cast<uint32_t*>(dose0 + (DSM ? (size_t)NB * seg_ind : 0));   // the unit's run starts sorted by row
    uint32_t* const start = evs + a.cap;                                     // [band rows + 1] first run start of each row
    uint32_t* const cursor = start + ROWS_MAX_BAND + 2;
    uint8_t* const evd = reinterpret_cast<uint8_t*>(cursor + ROWS_MAX_BAND + 2);   // rows until the run start's cell changes again
    __shared__ int s_unit;
    const int64_t n_cols = a.n_cols_ind * P;
    const bool count_ok = a.dose_which >= 0 && a.dose_which < 256;           // DK 1: a founder allele that exists

    for (;;) {
        __syncthreads();   // (an issuing warp gets here only after the engine has read its last row of the previous unit)
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
        if (n_ev > (uint32_t)a.cap) {   // more run starts than the list holds: k_emit_rows_overflow searches per cell
            if (tid == 0) a.overflow[1 + atomicAdd(a.overflow, 1)] = u;
            continue;
        }

        // ---- start state of the band, run starts bucketed by row (all threads)
        for (int i = tid * 16; i < seg_cells; i += NT * 16)
            *reinterpret_cast<uint4*>(img0 + i) = __ldg(reinterpret_cast<const uint4*>(pu + 16 + i));
        bucket_run_starts(pu, seg_cells, a.cap, n_ev, n_rows, evs, evd, start, cursor);
        if (DK >= 2) {   // 4-bit image: low 3 bits of every cell, the PRMT selectors of the dose lookup
            for (int i = tid; i < seg_cells / 16; i += NT) {
                const uint4 v = reinterpret_cast<const uint4*>(img0)[i];
                auto pack = [](uint32_t x) { return (x & 0x7) | ((x >> 4) & 0x70) | ((x >> 8) & 0x700) | ((x >> 12) & 0x7000); };
                const uint2 pk = make_uint2(pack(v.x) | (pack(v.y) << 16), pack(v.z) | (pack(v.w) << 16));
                reinterpret_cast<uint2*>(nib0)[i] = pk;
                reinterpret_cast<uint2*>(nib0 + seg_cells / 8)[i] = pk;
            }
        }
        if (DK == 1) {   // dose of one founder allele on the band's first row
            for (int i = tid; i < seg_ind; i += NT) {
                int dsum = 0;
                for (int h = 0; h < P; h++) dsum += img0[i * P + h] == (uint8_t)a.dose_which;
                dose0[i] = (uint8_t)(count_ok ? dsum : 0);
            }
        }
        __syncthreads();
        // the other buffers start from the same state (they become rows 1 .. NB-1)
        for (int b = 1; b < NB; b++) {
            for (int i = tid; i < seg_cells / 16; i += NT)
                reinterpret_cast<uint4*>(img0 + (size_t)b * seg_cells)[i] = reinterpret_cast<const uint4*>(img0)[i];
            if (DK == 1)
                for (int i = tid; i < seg_ind / 16; i += NT)
                    reinterpret_cast<uint4*>(dose0 + (size_t)b * seg_ind)[i] = reinterpret_cast<const uint4*>(dose0)[i];
        }
        __syncthreads();

        if (tid >= ROWS_COMPUTE) {
            // ---- the issuing warp of buffer b: lane 0 sends the founder row, lane 1 the dose row, lanes 2.. the ragged ends
            const int lane = tid & 31, b0 = (tid - ROWS_COMPUTE) >> 5;
            for (int r = b0; r < n_rows; r += ROWS_ISSUE_WARPS) {
                const int b = r % NB;
                const uint8_t* const img = img0 + (size_t)b * seg_cells;
                const uint8_t* const drow = dose0 + (size_t)b * seg_ind;
                uint8_t* const gf = a.founder ? a.founder + (out_row0 + r) * a.founder_pitch + c0 : nullptr;
                uint8_t* const gd = a.dose ? a.dose + (out_row0 + r) * a.dose_pitch + c0 / P : nullptr;
                bar_sync(BAR_FULL + b, ROWS_GROUP_THREADS + 32);
                if (ROWS_INFLIGHT > 0 ? lane == 0 : lane < 2) {
                    if ((ROWS_INFLIGHT > 0 || lane == 0) && gf && (n_cell & ~15)) bulk_store(gf, img, (uint32_t)(n_cell & ~15));
                    if ((ROWS_INFLIGHT > 0 || lane == 1) && DSM && gd && (n_ind & ~15)) bulk_store(gd, drow, (uint32_t)(n_ind & ~15));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                } else if (lane >= 2) {   // a bulk store moves multiples of 16 bytes
                    const int t = lane - 2;
                    if (gf && t < 15 && t < (n_cell & 15)) gf[(n_cell & ~15) + t] = img[(n_cell & ~15) + t];
                    if (DSM && gd && t >= 15 && t - 15 < (n_ind & 15)) gd[(n_ind & ~15) + t - 15] = drow[(n_ind & ~15) + t - 15];
                }
                // the engine has read the row (one warp per buffer) / row r - INFLIGHT + 1 (one issuing warp)
                if (lane < 2) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(ROWS_INFLIGHT > 0 ? ROWS_INFLIGHT - 1 : 0) : "memory");
                __syncwarp();
                const int rr = r - (ROWS_INFLIGHT > 0 ? ROWS_INFLIGHT - 1 : 0);
                if (rr >= 0 && rr + NB < n_rows) bar_arrive(bar_empty(rr + NB), ROWS_GROUP_THREADS + 32);
            }
            if (ROWS_INFLIGHT > 1 && lane < 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            continue;
        }

        // ---- the rows. Buffer r % NB holds row r-NB when row r begins and takes the run starts of rows r-NB+1 .. r;
        // the 4-bit image of row parity r & 1 holds row r-2 and takes those of rows r-1 and r.
        constexpr int CT = ROWS_GROUP_THREADS;
        const int grp = tid / CT, lt = tid % CT;
        uint4 Tn = make_uint4(0, 0, 0, 0);
        if (DK >= 2 && grp < n_rows) Tn = __ldg(a.mtab + g0 + grp);
        for (int r = grp; r < n_rows; r += ROWS_GROUPS) {
            const int b = r % NB;
            uint8_t* const img = img0 + (size_t)b * seg_cells;
            uint8_t* const drow = dose0 + (size_t)b * seg_ind;
            uint32_t* const nib = nib0 + (size_t)(r & 1) * (seg_cells / 8);
            const int w0 = max(1, r - NB + 1);                    // first row of the buffer's window (row 0 has no run starts: they are the state)
            const uint32_t e0 = start[w0], en = start[max(1, r - 1)], e1 = start[r + 1];
            if (r >= NB) bar_sync(bar_empty(r), ROWS_GROUP_THREADS + 32);   // the engine has read row r-NB out of this buffer
            for (uint32_t e = e0 + lt; e < e1; e += CT) {
                const uint32_t w = evs[e], cs = (w >> 8) & 0xFFFF, f = w & 0xFF;
                if ((int)(w >> 24) + (int)evd[e] <= r) continue;   // the cell changes again inside the window
                if (DK == 1) {
                    const int delta = (int)(f == (uint32_t)a.dose_which) - (int)(count_ok && img[cs] == (uint8_t)a.dose_which);
                    const uint32_t ind = cs / P;
                    if (delta) atomicAdd(reinterpret_cast<uint32_t*>(drow) + (ind >> 2), (uint32_t)delta << (8 * (ind & 3)));
                }
                img[cs] = (uint8_t)f;
                if (DK >= 2 && e >= en) {
                    const uint32_t sh = (cs & 7) * 4;
                    atomicAnd(&nib[cs >> 3], ~(0xFu << sh));
                    atomicOr(&nib[cs >> 3], (w & 7) << sh);
                }
            }
            if (DK >= 2) {
                bar_sync(BAR_WORK + grp, CT);
                const uint4 T = Tn;
                if (r + ROWS_GROUPS < n_rows) Tn = __ldg(a.mtab + g0 + r + ROWS_GROUPS);
                // the dose row is recomputed from the 4-bit image and goes from registers straight to global memory
                // (16 / 8 bytes per thread, consecutive threads consecutive bytes): no shared-memory row, no second
                // bulk store per table row
                uint8_t* const gd = a.dose ? a.dose + (out_row0 + r) * a.dose_pitch + c0 / P : nullptr;
                if (DK == 2) {   // 8 individuals per step: 4 words of the 4-bit image in, 8 dose bytes out
                    for (int i = lt; i < (n_ind + 7) / 8; i += CT) {
                        const uint4 v = reinterpret_cast<const uint4*>(nib)[i];
                        auto two = [&](uint32_t w) { return __dp4a(prmt(T.x, T.y, w) + prmt(T.z, T.w, w >> 16), 0x01010101u, 0u); };
                        const uint32_t lo = two(v.x) | (two(v.y) << 8), hi = two(v.z) | (two(v.w) << 8);   // dose nibbles
                        const uint2 out = make_uint2(prmt(0x03020100u, 0x07060504u, lo), prmt(0x03020100u, 0x07060504u, hi));
                        if (DSM) reinterpret_cast<uint2*>(drow)[i] = out;
                        else if (i * 8 + 8 <= n_ind) *reinterpret_cast<uint2*>(gd + i * 8) = out;
                        else for (int q = 0; i * 8 + q < n_ind; q++) gd[i * 8 + q] = (uint8_t)((q < 4 ? out.x : out.y) >> (8 * (q & 3)));
                    }
                } else {         // ploidy 2: 16 individuals per step
                    for (int i = lt; i < (n_ind + 15) / 16; i += CT) {
                        const uint4 v = reinterpret_cast<const uint4*>(nib)[i];
                        auto four = [&](uint32_t w) {
                            const uint32_t xl = prmt(T.x, T.y, w), xh = prmt(T.x, T.y, w >> 16);
                            return prmt(xl + (xl >> 8), xh + (xh >> 8), 0x6420u);
                        };
                        const uint4 out = make_uint4(four(v.x), four(v.y), four(v.z), four(v.w));
                        if (DSM) reinterpret_cast<uint4*>(drow)[i] = out;
                        else if (i * 16 + 16 <= n_ind) *reinterpret_cast<uint4*>(gd + i * 16) = out;
                        else {
                            const uint32_t w4[4] = {out.x, out.y, out.z, out.w};
                            for (int q = 0; i * 16 + q < n_ind; q++) gd[i * 16 + q] = (uint8_t)(w4[q >> 2] >> (8 * (q & 3)));
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bar_arrive(BAR_FULL + b, ROWS_GROUP_THREADS + 32);
        }
    }
    if (tid >= ROWS_COMPUTE && (tid & 31) < 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Units whose run-start list overflowed (not expected with the band heuristic; the list is filled by
// k_emit_rows on the device, so no host round trip): every cell searched in its homolog's runs.
struct RowsOverflowArgs {
    int dose_mode;                // 0: count founder allele dose_which; 1: count the locus' matched allele code
    const uint8_t* code; const uint8_t* match_code; int A;
};
__global__ void __launch_bounds__(256) k_emit_rows_overflow(const RowsArgs a, const RowsOverflowArgs o) {
    const int n_over = a.overflow[0];
    const int P = a.P;
    const int64_t n_cols = a.n_cols_ind * P;
    for (int x = blockIdx.y; x < n_over; x += gridDim.y) {
        const int u = a.overflow[1 + x];
        int k = 0;
        while (u >= a.chunk_unit_base[k + 1]) k++;
        const int j = (u - a.chunk_unit_base[k]) / a.n_seg, s = (u - a.chunk_unit_base[k]) % a.n_seg;
        int64_t g0, g1;
        band_rows(a, k, j, g0, g1);
        const int64_t c0 = (int64_t)s * a.seg_cells;
        const int64_t n_ind = min((int64_t)a.seg_cells, n_cols - c0) / P, q0 = c0 / P;
        const int c = a.chunk_chrom[k];
        for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (g1 - g0) * n_ind; t += (int64_t)gridDim.x * blockDim.x) {
            const int64_t gl = g0 + t / n_ind, q = q0 + t % n_ind;
            const uint32_t l = (uint32_t)(gl - a.chunk_coff[k]);
            const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
            const int64_t hid0 = (((int64_t)c * a.R + r) * a.N + i) * P;
            const int64_t row = gl - a.locus_first;
            const uint8_t* codes = o.code ? o.code + gl * o.A : nullptr;
            const int match = o.dose_mode == 1 ? (int)o.match_code[gl] : a.dose_which;
            int dose = 0;
            for (int h = 0; h < P; h++) {
                const uint64_t d = a.run_desc[hid0 + h];
                const uint64_t b = desc_begin(d);
                const uint32_t n = desc_count(d);
                uint32_t lo = 0, hi = n;
                while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.run_lb[b + mid] <= l) lo = mid; else hi = mid; }
                const int f = a.run_founder[b + lo];
                if (a.founder) a.founder[row * a.founder_pitch + q * P + h] = (uint8_t)f;
                if ((codes ? (int)codes[f] : f) == match) dose++;
            }
            if (a.dose) a.dose[row * a.dose_pitch + q] = (uint8_t)dose;
        }
    }
}

// ------------------------------------------------------------------------------------------
// PSIM_EMIT_PACKED4: the same tables with two 4-bit cells per byte (cell 2k in the low nibble). The row
// image IS the packed founder row (4 bits per cell), so there is no byte image at all: a buffer is
// patched with one atomicXor per run start (cells that share a word are patched by different threads),
// the dose nibbles come straight out of the PRMT / IDP.4A lookup, and a table row leaves the SM as
// half the bytes. Same plan, same issuing warps, same barriers as k_emit_rows.
template <int DK>
__global__ void ROWS_KERNEL_BOUNDS k_emit_rows_packed(const RowsArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int NT = ROWS_THREADS;
    const int P = a.P;
    const int seg_cells = a.seg_cells, seg_ind = seg_cells / P;   // seg_cells is a multiple of 32 * P
    constexpr int NB = ROWS_NBUF;
    const int img_bytes = seg_cells / 2, dose_bytes = seg_ind / 2;
    uint8_t* const img0 = smem;                                              // 4-bit founder of every cell on the buffer's row (x NB)
    uint8_t* const dose0 = img0 + (size_t)NB * img_bytes;                    // DK >= 1: 4-bit dose rows (x NB)
    uint32_t* const evs = reinterpret_cast<uint32_t*>(dose0 + (DK >= 1 ? (size_t)NB * dose_bytes : 0));
    uint32_t* const start = evs + a.cap;
    uint32_t* const cursor = start + ROWS_MAX_BAND + 2;
    uint8_t* const evd = reinterpret_cast<uint8_t*>(cursor + ROWS_MAX_BAND + 2);
    __shared__ int s_unit;
    const int64_t n_cols = a.n_cols_ind * P;
    const bool count_ok = a.dose_which >= 0 && a.dose_which < 16;            // DK 1: a founder allele that exists

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
        const int64_t c0 = (int64_t)s * seg_cells;
        const int n_cell = (int)min((int64_t)seg_cells, n_cols - c0);
        const int n_ind = n_cell / P;
        const int row_bytes = (n_cell + 1) / 2, drow_bytes = (n_ind + 1) / 2;   // (cells past the end are zero)
        const uint8_t* pu = a.plan + (int64_t)u * a.plan_stride;
        const uint32_t n_ev = *reinterpret_cast<const uint32_t*>(pu);
        if (n_ev > (uint32_t)a.cap) {
            if (tid == 0) a.overflow[1 + atomicAdd(a.overflow, 1)] = u;
            continue;
        }
        // ---- start state of the band packed to 4 bits, run starts bucketed by row (all threads)
        for (int i = tid; i < seg_cells / 16; i += NT) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(pu + 16) + i);
            if (i * 16 + 16 > n_cell) {   // the plan's state bytes beyond the last column were never written
                uint32_t w4[4] = {v.x, v.y, v.z, v.w};
                for (int q = 0; q < 16; q++) if (i * 16 + q >= n_cell) w4[q >> 2] &= ~(0xFFu << (8 * (q & 3)));
                v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
            auto pack = [](uint32_t x) { return (x & 0xF) | ((x >> 4) & 0xF0) | ((x >> 8) & 0xF00) | ((x >> 12) & 0xF000); };
            reinterpret_cast<uint2*>(img0)[i] = make_uint2(pack(v.x) | (pack(v.y) << 16), pack(v.z) | (pack(v.w) << 16));
        }
        bucket_run_starts(pu, seg_cells, a.cap, n_ev, n_rows, evs, evd, start, cursor);
        if (DK == 1) {   // dose of one founder allele on the band's first row, 4 bits per individual
            for (int i = tid; i < seg_ind / 8; i += NT) {   // 8 individuals = one word of dose nibbles
                uint32_t out = 0;
                for (int q = 0; q < 8; q++) {
                    const int ind = i * 8 + q;
                    int dsum = 0;
                    for (int h = 0; h < P; h++) {
                        const int cell = ind * P + h;
                        dsum += ((img0[cell >> 1] >> (4 * (cell & 1))) & 15) == a.dose_which;
                    }
                    if (count_ok && ind < n_ind) out |= (uint32_t)dsum << (4 * q);
                }
                reinterpret_cast<uint32_t*>(dose0)[i] = out;
            }
        }
        __syncthreads();
        for (int b = 1; b < NB; b++) {
            for (int i = tid; i < img_bytes / 16; i += NT)
                reinterpret_cast<uint4*>(img0 + (size_t)b * img_bytes)[i] = reinterpret_cast<const uint4*>(img0)[i];
            if (DK == 1)
                for (int i = tid; i < dose_bytes / 16; i += NT)
                    reinterpret_cast<uint4*>(dose0 + (size_t)b * dose_bytes)[i] = reinterpret_cast<const uint4*>(dose0)[i];
        }
        __syncthreads();

        if (tid >= ROWS_COMPUTE) {
            const int lane = tid & 31, b0 = (tid - ROWS_COMPUTE) >> 5;
            for (int r = b0; r < n_rows; r += ROWS_ISSUE_WARPS) {
                const int b = r % NB;
                const uint8_t* const img = img0 + (size_t)b * img_bytes;
                const uint8_t* const drow = dose0 + (size_t)b * dose_bytes;
                uint8_t* const gf = a.founder ? a.founder + (out_row0 + r) * a.founder_pitch + c0 / 2 : nullptr;
                uint8_t* const gd = a.dose ? a.dose + (out_row0 + r) * a.dose_pitch + c0 / P / 2 : nullptr;
                bar_sync(BAR_FULL + b, ROWS_GROUP_THREADS + 32);
                if (ROWS_INFLIGHT > 0 ? lane == 0 : lane < 2) {
                    if ((ROWS_INFLIGHT > 0 || lane == 0) && gf && (row_bytes & ~15)) bulk_store(gf, img, (uint32_t)(row_bytes & ~15));
                    if ((ROWS_INFLIGHT > 0 || lane == 1) && DK >= 1 && gd && (drow_bytes & ~15)) bulk_store(gd, drow, (uint32_t)(drow_bytes & ~15));
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                } else if (lane >= 2) {
                    const int t = lane - 2;
                    if (gf && t < 15 && t < (row_bytes & 15)) gf[(row_bytes & ~15) + t] = img[(row_bytes & ~15) + t];
                    if (DK >= 1 && gd && t >= 15 && t - 15 < (drow_bytes & 15)) gd[(drow_bytes & ~15) + t - 15] = drow[(drow_bytes & ~15) + t - 15];
                }
                if (lane < 2) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(ROWS_INFLIGHT > 0 ? ROWS_INFLIGHT - 1 : 0) : "memory");
                __syncwarp();
                const int rr = r - (ROWS_INFLIGHT > 0 ? ROWS_INFLIGHT - 1 : 0);
                if (rr >= 0 && rr + NB < n_rows) bar_arrive(bar_empty(rr + NB), ROWS_GROUP_THREADS + 32);
            }
            if (ROWS_INFLIGHT > 1 && lane < 2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            continue;
        }

        constexpr int CT = ROWS_GROUP_THREADS;
        const int grp = tid / CT, lt = tid % CT;
        uint4 Tn = make_uint4(0, 0, 0, 0);
        if (DK >= 2 && grp < n_rows) Tn = __ldg(a.mtab + g0 + grp);
        for (int r = grp; r < n_rows; r += ROWS_GROUPS) {
            const int b = r % NB;
            uint32_t* const img = reinterpret_cast<uint32_t*>(img0 + (size_t)b * img_bytes);
            uint32_t* const drow = reinterpret_cast<uint32_t*>(dose0 + (size_t)b * dose_bytes);
            const int w0 = max(1, r - NB + 1);
            const uint32_t e0 = start[w0], e1 = start[r + 1];
            if (r >= NB) bar_sync(bar_empty(r), ROWS_GROUP_THREADS + 32);
            for (uint32_t e = e0 + lt; e < e1; e += CT) {
                const uint32_t w = evs[e], cs = (w >> 8) & 0xFFFF, f = w & 0xF;
                if ((int)(w >> 24) + (int)evd[e] <= r) continue;   // the cell changes again inside the window
                const uint32_t sh = (cs & 7) * 4;
                const uint32_t old = (img[cs >> 3] >> sh) & 15u;   // (only this thread changes this nibble)
                if (DK == 1) {
                    const int delta = (int)(f == (uint32_t)a.dose_which) - (int)(count_ok && old == (uint32_t)a.dose_which);
                    const uint32_t ind = cs / P;
                    if (delta) atomicAdd(drow + (ind >> 3), (uint32_t)delta << (4 * (ind & 7)));
                }
                atomicXor(&img[cs >> 3], (old ^ f) << sh);
            }
            if (DK >= 2) {
                bar_sync(BAR_WORK + grp, CT);
                const uint4 T = Tn;
                if (r + ROWS_GROUPS < n_rows) Tn = __ldg(a.mtab + g0 + r + ROWS_GROUPS);
                if (DK == 2) {   // 8 individuals per step: 4 words of the image in, 8 dose nibbles out
                    for (int i = lt; i < seg_ind / 8; i += CT) {
                        const uint4 v = reinterpret_cast<const uint4*>(img)[i];
                        auto two = [&](uint32_t w) { return __dp4a(prmt(T.x, T.y, w) + prmt(T.z, T.w, w >> 16), 0x01010101u, 0u); };
                        uint32_t out = two(v.x) | (two(v.y) << 8) | (two(v.z) << 16) | (two(v.w) << 24);
                        if (i * 8 + 8 > n_ind) out &= i * 8 >= n_ind ? 0u : 0xFFFFFFFFu >> (4 * (i * 8 + 8 - n_ind));   // nibbles past the last individual are zero
                        drow[i] = out;
                    }
                } else {         // ploidy 2: 16 individuals per step
                    for (int i = lt; i < seg_ind / 16; i += CT) {
                        const uint4 v = reinterpret_cast<const uint4*>(img)[i];
                        auto four = [&](uint32_t w) {   // 4 individuals of one word -> 16 bits of dose nibbles
                            const uint32_t xl = prmt(T.x, T.y, w), xh = prmt(T.x, T.y, w >> 16);
                            const uint32_t d = prmt(xl + (xl >> 8), xh + (xh >> 8), 0x6420u);
                            return (d & 0xF) | ((d >> 4) & 0xF0) | ((d >> 8) & 0xF00) | ((d >> 12) & 0xF000);
                        };
                        uint32_t o0 = four(v.x) | (four(v.y) << 16), o1 = four(v.z) | (four(v.w) << 16);
                        if (i * 16 + 8 > n_ind) o0 &= i * 16 >= n_ind ? 0u : 0xFFFFFFFFu >> (4 * (i * 16 + 8 - n_ind));
                        if (i * 16 + 16 > n_ind) o1 &= i * 16 + 8 >= n_ind ? 0u : 0xFFFFFFFFu >> (4 * (i * 16 + 16 - n_ind));
                        reinterpret_cast<uint2*>(drow)[i] = make_uint2(o0, o1);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bar_arrive(BAR_FULL + b, ROWS_GROUP_THREADS + 32);
        }
    }
    if (tid >= ROWS_COMPUTE && (tid & 31) < 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// per-cell kernels for packed tables: one thread per (row, pair of emitted individuals)
struct EmitPackedArgs {
    int C, R, P;
    int64_t N;
    const uint64_t* run_desc;
    const uint32_t* run_lb;
    const uint16_t* run_founder;
    const int64_t* locus_off;
    int64_t ind_first, ind_count, n_cols_ind;
    int64_t locus_first, l_begin, l_count;
    int64_t q_first, q_count;     // emitted individuals of this launch; q_first even
    uint8_t* founder; int64_t founder_pitch;
    uint8_t* dose; int64_t dose_pitch;
    int dose_mode, dose_which;
    const uint8_t* code; const uint8_t* match_code; int A;
};
__device__ __forceinline__ void emit_packed_pair(const EmitPackedArgs& a, int c, uint32_t l, int64_t gl, int64_t q) {
    const int P = a.P;
    const int64_t row = gl - a.locus_first;
    const uint8_t* codes = a.code ? a.code + gl * a.A : nullptr;
    const int match = a.dose_mode == 1 ? (int)a.match_code[gl] : a.dose_which;
    uint8_t fl[32];       // founder alleles of the pair's 2 * P cells
    for (int k = 0; k < 2 * P; k++) fl[k] = 0;
    int dose2 = 0;
    for (int x = 0; x < 2; x++) {
        if (q + x >= a.q_first + a.q_count) break;
        const int64_t r = (q + x) / a.ind_count, i = a.ind_first + (q + x) % a.ind_count;
        const int64_t hid0 = (((int64_t)c * a.R + r) * a.N + i) * P;
        int dose = 0;
        for (int h = 0; h < P; h++) {
            const uint64_t d = a.run_desc[hid0 + h];
            const uint64_t b = desc_begin(d);
            const uint32_t n = desc_count(d);
            uint32_t lo = 0, hi = n;
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.run_lb[b + mid] <= l) lo = mid; else hi = mid; }
            const int f = a.run_founder[b + lo];
            fl[x * P + h] = (uint8_t)(f & 15);
            if ((codes ? (int)codes[f] : f) == match) dose++;
        }
        dose2 |= dose << (4 * x);
    }
    if (a.founder) for (int k = 0; k < P; k++) a.founder[row * a.founder_pitch + (q / 2) * P + k] = (uint8_t)(fl[2 * k] | (fl[2 * k + 1] << 4));
    if (a.dose) a.dose[row * a.dose_pitch + q / 2] = (uint8_t)dose2;
}
__global__ void __launch_bounds__(256) k_emit_generic_packed(const EmitPackedArgs a) {
    const int64_t pairs = (a.q_count + 1) / 2;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= a.l_count * pairs) return;
    const int64_t gl = a.l_begin + t / pairs, q = a.q_first + 2 * (t % pairs);
    int c = 0;
    while (gl >= a.locus_off[c + 1]) c++;
    emit_packed_pair(a, c, (uint32_t)(gl - a.locus_off[c]), gl, q);
}
__global__ void __launch_bounds__(256) k_emit_rows_overflow_packed(const RowsArgs a, const RowsOverflowArgs o, const int64_t* locus_off) {
    const int n_over = a.overflow[0];
    const int P = a.P;
    const int64_t n_cols = a.n_cols_ind * P;
    for (int x = blockIdx.y; x < n_over; x += gridDim.y) {
        const int u = a.overflow[1 + x];
        int k = 0;
        while (u >= a.chunk_unit_base[k + 1]) k++;
        const int j = (u - a.chunk_unit_base[k]) / a.n_seg, s = (u - a.chunk_unit_base[k]) % a.n_seg;
        int64_t g0, g1;
        band_rows(a, k, j, g0, g1);
        const int64_t c0 = (int64_t)s * a.seg_cells;
        EmitPackedArgs e{};
        e.C = a.C; e.R = a.R; e.P = P; e.N = a.N;
        e.run_desc = a.run_desc; e.run_lb = a.run_lb; e.run_founder = a.run_founder; e.locus_off = locus_off;
        e.ind_first = a.ind_first; e.ind_count = a.ind_count; e.n_cols_ind = a.n_cols_ind;
        e.locus_first = a.locus_first;
        e.q_first = c0 / P; e.q_count = min((int64_t)a.seg_cells, n_cols - c0) / P;
        e.founder = a.founder; e.founder_pitch = a.founder_pitch; e.dose = a.dose; e.dose_pitch = a.dose_pitch;
        e.dose_mode = o.dose_mode; e.dose_which = a.dose_which; e.code = o.code; e.match_code = o.match_code; e.A = o.A;
        const int64_t pairs = (e.q_count + 1) / 2;
        for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (g1 - g0) * pairs; t += (int64_t)gridDim.x * blockDim.x) {
            const int64_t gl = g0 + t / pairs;
            emit_packed_pair(e, a.chunk_chrom[k], (uint32_t)(gl - a.chunk_coff[k]), gl, e.q_first + 2 * (t % pairs));
        }
    }
}
