// K4 — TEST-mode statistics
// (part of libpsim: included by psim.cu inside namespace psim)
#pragma once

// ------------------------------------------------------------------------------------------
// K4 — TEST-mode statistics (Main.testMeiosis, Main.java:1660-1721) as device reductions over the
// gametes that made the simulated individuals: homologs [0, P/2) of an individual are the gamete
// of parent 0's meiosis, [P/2, P) that of parent 1's. All of it works on the locus-index runs of K0.
struct StatsArgs {
    int C, R, P, A;
    int gp;                           // homologs per gamete: P/2, or P with unreduced gametes
    int64_t N;
    const uint64_t* run_desc; const uint32_t* run_lb; const uint16_t* run_founder;
    const uint64_t* desc;             // unmerged segment lists (segment counts as the reference holds them)
    const int8_t* cfg_quad;           // [R*N][C][2]
    int64_t ind_first, ind_count;
    int chrom; uint32_t Lc;
    uint32_t* allele_count;           // [A][Lc]
    uint32_t* dose_count;             // [A][Lc][gp+1]
    uint32_t* homozygous_count;       // [Lc]
    int32_t* diff;                    // [(Lc+1)][(Lc+1)] 2-D difference array of the recombination counts
    uint32_t* segments_hist;          // [64]
    uint32_t* quadrivalent_hist;      // [P/4+1]
};

__device__ __forceinline__ uint32_t founder_at_row(const StatsArgs& a, int64_t hid, uint32_t l) {
    const uint64_t d = a.run_desc[hid];
    const uint64_t b = desc_begin(d);   // (40 bits: the slab may hold more than 2^32 segments)
        const uint32_t n = desc_count(d);
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.run_lb[b + mid] <= l) lo = mid; else hi = mid; }
    return a.run_founder[b + lo];
}

// one block per locus: founder-allele counts, dose histogram and homozygosity over all gametes
__global__ void __launch_bounds__(256) k_stats_locus(const StatsArgs a) {
    extern __shared__ uint32_t s_hist[];     // [A] counts, then [A][gp+1] doses, then 1 homozygous
    const uint32_t l = blockIdx.x;
    const int p2 = a.gp, nd = p2 + 1, per_ind = a.P / a.gp;
    const int n_bins = a.A + a.A * nd + 1;
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t n_gam = (int64_t)per_ind * a.ind_count * a.R;
    uint32_t zero_gametes = 0;   // gametes seen by this thread: dose 0 bins are filled in bulk at the end
    for (int64_t g = threadIdx.x; g < n_gam; g += blockDim.x) {
        const int half = (int)(g % per_ind);
        const int64_t q = g / per_ind;
        const int64_t r = q / a.ind_count, i = a.ind_first + q % a.ind_count;
        const int64_t hid0 = (((int64_t)a.chrom * a.R + r) * a.N + i) * a.P + (int64_t)half * p2;
        uint32_t f[MAXP];
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
    const uint64_t b = desc_begin(d);   // (40 bits: the slab may hold more than 2^32 segments)
        const uint32_t n = desc_count(d);
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

