// K0 — breakpoints to locus-index runs
// (part of libpsim: included by psim.cu inside namespace psim)
#pragma once

// ------------------------------------------------------------------------------------------
// K0 — breakpoints (fp64 Morgan) -> runs in locus-index space, so that emission is integer-only.
// Locus l belongs to the last segment s with pos[s] <= locus position (HaploStruct.findSegment
// :107-112, segment 0 catches everything below); with lb[s] = #loci strictly left of pos[s] that is
// "last s with lb[s] <= l" (ties: a locus exactly on a breakpoint belongs to the segment starting
// there). A suffix minimum makes lb monotone without changing that answer, empty runs are
// dropped and neighbours with the same founder merged.
// A map is searched through a grid over its positions (psim_set_map): bucket k of chromosome c starts at
// lo + k * width and grid[k] = #loci left of that, so that a breakpoint's answer lies between the entries one bucket
// below and two above its own - a bisection of one to three steps instead of log2(loci) (the search was half of this
// kernel's instructions). The grid only narrows the interval; the answer is the plain lower bound.
struct MapGrid {
    const int32_t* grid;      // per chromosome: n_buckets + 1 entries from grid_off[c]
    const int64_t* grid_off;  // [C + 1]
    const double* lo;         // [C] position of the chromosome's first locus
    const double* inv_width;  // [C] buckets per Morgan (0: all loci at one position)
};
__global__ void __launch_bounds__(128) k_index_runs(int64_t n_hom, int64_t hom_per_chrom, const uint64_t* __restrict__ desc,
                                                    const double* __restrict__ seg_pos, const int32_t* __restrict__ seg_founder,
                                                    const int64_t* __restrict__ locus_off, const double* __restrict__ locus_pos,
                                                    const MapGrid mg,
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
    const int32_t* const grid = mg.grid + mg.grid_off[c];
    const int n_buckets = (int)(mg.grid_off[c + 1] - mg.grid_off[c]) - 1;
    const double g_lo = mg.lo[c], g_inv = mg.inv_width[c];
    auto lower_bound = [&](double p) {   // #loci strictly left of p
        const double t = (p - g_lo) * g_inv;
        const int k = t > 0.0 ? (t < (double)n_buckets ? (int)t : n_buckets - 1) : 0;
        uint32_t lo = (uint32_t)__ldg(grid + max(k - 1, 0)), hi = (uint32_t)__ldg(grid + min(k + 2, n_buckets));
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(lp + mid) < p) lo = mid + 1; else hi = mid; }
        return lo;
    };
    int w = 0;
    constexpr int SMALL = 8;
    if (n <= SMALL) {
        // the common case (a homolog a few generations from its founders): everything in registers,
        // one pass over global memory instead of three
        uint32_t lb[SMALL];
        uint16_t fo[SMALL];
#pragma unroll
        for (int s = 0; s < SMALL; s++) {
            lb[s] = Lc; fo[s] = 0;
            if (s < n) { lb[s] = s > 0 ? lower_bound(seg_pos[b + s]) : 0u; fo[s] = (uint16_t)seg_founder[b + s]; }
        }
        uint32_t mn = Lc;
#pragma unroll
        for (int s = SMALL - 1; s >= 0; s--) { mn = lb[s] < mn ? lb[s] : mn; lb[s] = mn; }   // suffix minimum (entries >= n hold Lc)
        uint32_t last_f = 0x10000u;
#pragma unroll
        for (int s = 0; s < SMALL; s++) {
            const uint32_t end = s + 1 < SMALL ? lb[s + 1] : Lc;
            if (s < n && lb[s] < end && fo[s] != last_f) {
                run_lb[b + w] = lb[s];
                run_founder[b + w] = fo[s];
                last_f = fo[s];
                w++;
            }
        }
    } else {
        // pass 1: lb[s]
        for (int s = 0; s < n; s++) run_lb[b + s] = s > 0 ? lower_bound(seg_pos[b + s]) : 0u;
        // pass 2: suffix minimum
        uint32_t mn = Lc;
        for (int s = n - 1; s >= 0; s--) { const uint32_t v = run_lb[b + s]; mn = v < mn ? v : mn; run_lb[b + s] = mn; }
        // pass 3: compact in place
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
    }
    run_desc[hid] = make_desc((uint64_t)b, (uint64_t)w);
}

