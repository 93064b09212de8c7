// K1 — meiosis: founders, and ONE kernel per generation batch that takes every (child, chromosome,
// parent) unit from the parent's haplostructs to the child's segments in the slab
// (part of libpsim: included by psim.cu inside namespace psim)
#pragma once

// ------------------------------------------------------------------------------------------
// Founders: Individual.calcFounderHaplostruct (Individual.java:155-171)
// job j = (founder slot f, replicate r): individual ind[f], first allele allele0[f]
__global__ void k_init_founders(int n_jobs, int n_new, const int32_t* __restrict__ ind, const int32_t* __restrict__ allele0,
                                int C, int R, int64_t N, int P, const ChromParams* __restrict__ chrom,
                                uint64_t* __restrict__ desc, double* __restrict__ seg_pos, int32_t* __restrict__ seg_founder,
                                const unsigned long long* __restrict__ taken_at, int64_t slab_cap) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)n_jobs * C * P;
    if (t >= total) return;
    const int64_t slab_base = (int64_t)*taken_at;   // reserved by k_slab_mark just before this kernel
    if (slab_base + total > slab_cap) return;       // (which has also set the status)
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

// Slab bookkeeping on the device, so that a whole simulation is enqueued without the host reading
// anything back. status[0]: 0, PSIM_ECAPACITY (a chiasma list overflowed) or STATUS_SLAB_FULL;
// status[1]: the first batch that did not complete (everything before it is intact).
constexpr int STATUS_SLAB_FULL = -1;
// mark[batch] = slab cursor when the batch begins (a batch that is run again starts there); `take`
// segments are then reserved for a kernel that knows its size in advance (founders): old cursor -> *taken_at.
// restart >= 0: the first batch of a (re)started simulation - the cursor is set and the status cleared.
__global__ void k_slab_mark(unsigned long long* cursor, unsigned long long* mark, int batch, long long take, int64_t slab_cap,
                            unsigned long long* taken_at, int* status, long long restart) {
    if (threadIdx.x || blockIdx.x) return;
    if (restart >= 0) { *cursor = (unsigned long long)restart; status[0] = 0; status[1] = 0x7FFFFFFF; }
    if (status[0] != 0) return;
    const unsigned long long at = *cursor;
    mark[batch] = at;
    if (take > 0) {
        *taken_at = at;
        *cursor = at + (unsigned long long)take;
        if ((int64_t)at + take > slab_cap) { status[0] = STATUS_SLAB_FULL; status[1] = batch; }
    }
}

// ------------------------------------------------------------------------------------------
// The meiosis kernel. A block owns `units_per_block` consecutive units (unit = child, chromosome,
// parent: one meiosis seen on one chromosome) and works in three steps:
//   1. one thread per unit draws the pairing configuration (bit masks in registers) and leaves it in
//      shared memory and in the configuration tables;
//   2. the multivalents of all the block's units are compacted by kind - quadrivalents first,
//      bivalents from the next warp boundary - and every thread takes one per round: chiasmata (into its
//      column of the chiasma scratch), segregation, the chromatids of the kept gamete(s) traced
//      once to count the segments they will have;
//   3. every warp reserves its segments in the slab with one atomic, and the same threads trace
//      again, writing the segments (Multivalent.fillGamete :539-577) and the child's descriptors
//      (Gamete.fertilization, Gamete.java:95-116: homologs 0..P/2-1 from parent 0's gamete,
//      P/2..P-1 from parent 1's).
// Warps therefore run one kind of multivalent, no task list goes through global memory, and a
// generation is one launch. Every multivalent owns a Philox stream, so results do not depend on the
// block size or on how the units are batched.
#ifndef PSIM_MEI_THREADS
#define PSIM_MEI_THREADS 256
#endif
constexpr int MEI_THREADS = PSIM_MEI_THREADS;
#ifndef PSIM_MEI_MIN_BLOCKS
#define PSIM_MEI_MIN_BLOCKS 4
#endif

struct MeiosisArgs {
    ModelParams model;
    const ChromParams* chrom;
    int C, R, P;
    int64_t N;
    uint32_t replicate_first;
    int n_lev;                   // individuals of this batch
    const int32_t* lev_ind;      // their indices
    const int32_t* par0;
    const int32_t* par1;
    uint64_t* desc;
    double* seg_pos;
    int32_t* seg_founder;
    int64_t slab_cap;
    unsigned long long* cursor;  // segments of the slab in use
    int* status;
    int batch;
    int8_t* cfg_quad;            // [R*N][C][2]
    int8_t* cfg_seq;             // [R*N][C][2][P]
    double* x_pos;               // chiasma scratch: entry e of thread slot t at [e * n_slots + t]
    uint8_t* x_ab;
    int x_cap;
    int units_per_block;
};

__device__ __forceinline__ void unit_decode(const MeiosisArgs& a, int64_t t, int& par, int& j, int& r, int& c) {
    par = (int)(t & 1);
    const int64_t unit = t >> 1;
    j = (int)(unit % a.n_lev);
    r = (int)((unit / a.n_lev) % a.R);
    c = (int)(unit / ((int64_t)a.n_lev * a.R));
}
__device__ __forceinline__ HomologView homolog_view(const MeiosisArgs& a, int64_t hid) {
    const uint64_t d = a.desc[hid];
    HomologView hv;
    hv.pos = a.seg_pos + desc_begin(d);
    hv.founder = a.seg_founder + desc_begin(d);
    hv.n = (int)desc_count(d);
    return hv;
}

// One chromatid of the kept gamete, followed from the head along the chiasmata it takes part in:
// the stretch between two of them is copied from the parental homolog the chromatid is on
// (Multivalent.slotsToPatterns :332-357 + fillGamete :539-577). WRITE = false only counts.
template <bool WRITE>
__device__ __forceinline__ int trace_chromatid(const MeiosisArgs& a, const ChiasmaList& x, int strand, Nib16 seq, int base,
                                               int64_t phid0, double head, double tail, int64_t w) {
    int first = 0;
    while (first < x.n && x.a(first) != strand && x.b(first) != strand) first++;
    if (first == x.n) {   // no exchange: the parental homolog as it is
        const HomologView hv = homolog_view(a, phid0 + seq.get(base + (strand >> 1)));
        if (WRITE) for (int s = 0; s < hv.n; s++) { a.seg_pos[w + s] = hv.pos[s]; a.seg_founder[w + s] = hv.founder[s]; }
        return hv.n;
    }
    int n = 0;
    double from = head;
    for (int e = first; e <= x.n; e++) {
        int next = -1;
        double to = tail;
        if (e < x.n) {
            const int sa = x.a(e), sb = x.b(e);
            if (sa == strand) next = sb; else if (sb == strand) next = sa; else continue;
            to = x.p(e);
        }
        const HomologView hv = homolog_view(a, phid0 + seq.get(base + (strand >> 1)));
        int seg = hv.find_segment(from);
        if (WRITE) { a.seg_pos[w + n] = from; a.seg_founder[w + n] = hv.founder[seg]; }
        n++;
        for (seg++; seg < hv.n && hv.pos[seg] < to; seg++) {
            if (WRITE) { a.seg_pos[w + n] = hv.pos[seg]; a.seg_founder[w + n] = hv.founder[seg]; }
            n++;
        }
        from = to;
        strand = next;
    }
    return n;
}

// KOS / ANR: MAPFUNCTION=KOSAMBI and ALLOWNORECOMBINATION as compile-time constants, so that the code
// of the other model (gamma sampler, pow, retry loop) is not in the kernel at all.
// 56 registers (a few spills; 64 would be spill-free and 2 % faster alone): four blocks then leave a slice of the
// register file free, and three fit next to the emission kernel (psim_emit_rows.cuh) when a caller pipelines the two.
#ifndef PSIM_MEI_MAXREG
#define PSIM_MEI_MAXREG 56
#endif
#if PSIM_MEI_MAXREG > 0
#define MEI_KERNEL_BOUNDS __maxnreg__(PSIM_MEI_MAXREG)
#else
#define MEI_KERNEL_BOUNDS __launch_bounds__(MEI_THREADS, PSIM_MEI_MIN_BLOCKS)
#endif
template <bool KOS, bool ANR>
__global__ void MEI_KERNEL_BOUNDS k_meiosis(const MeiosisArgs a) {
    typedef cub::BlockScan<int, MEI_THREADS> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ uint64_t s_seq[MEI_THREADS];
    __shared__ uint8_t s_nquad[MEI_THREADS];
    __shared__ uint16_t s_task[MEI_THREADS * (MAXP / 2) + 32];
    __shared__ int s_stop;
    const int tid = threadIdx.x;
    if (tid == 0) s_stop = *(volatile int*)a.status != 0;   // an earlier batch failed: the host runs this one again
    __syncthreads();
    if (s_stop) return;
    const int P = a.P, p2 = P / 2;
    const int64_t n_units = (int64_t)a.C * a.R * a.n_lev * 2;
    const int64_t unit0 = (int64_t)blockIdx.x * a.units_per_block;

    // ---- 1. pairing configuration of the block's units
    int nq = 0, nb = 0;
    if (tid < a.units_per_block && unit0 + tid < n_units) {
        int par, j, r, c;
        unit_decode(a, unit0 + tid, par, j, r, c);
        if (!(a.model.unreduced && par == 1)) {   // 2NGAMETES: the individual is parent 0's 2n gamete alone
            const int32_t child = a.lev_ind[j];
            const int32_t parent = par == 0 ? a.par0[child] : a.par1[child];
            const ChromParams& ch = a.chrom[c];
            const int64_t phid0 = (((int64_t)c * a.R + r) * a.N + parent) * P;
            Nib16 g_head, g_tail;   // genome = founder allele % (P/2) at both chromosome ends (Individual.java:437-470)
            g_head.v = 0; g_tail.v = 0;
            if (P > 2)
                for (int h = 0; h < P; h++) {
                    const HomologView hv = homolog_view(a, phid0 + h);
                    g_head.set(h, hv.founder[hv.find_segment(ch.head)] % p2);
                    g_tail.set(h, hv.founder[hv.find_segment(ch.tail)] % p2);
                }
            PhiloxStream rng;
            rng.init(a.model.seed, a.replicate_first + (uint32_t)r);
            rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, STAGE_PAIRING);
            const PairingResult pr = draw_pairing(rng, a.model, ch, g_head, g_tail);
            s_seq[tid] = pr.seq.v;
            s_nquad[tid] = (uint8_t)pr.quad;
            nq = pr.quad;
            nb = (P - 4 * pr.quad) / 2;
            const int64_t q = (((int64_t)r * a.N + child) * a.C + c) * 2 + par;
            a.cfg_quad[q] = (int8_t)pr.quad;
            for (int h = 0; h < P; h++) a.cfg_seq[q * P + h] = (int8_t)pr.seq.get(h);
        }
    }
    // ---- 2. one multivalent per thread, quadrivalents and bivalents in different warps
    int before, all;
    Scan(scan_tmp).ExclusiveSum(nq | (nb << 16), before, all);
    const int n_quad = all & 0xFFFF, n_bi = all >> 16, bi0 = (n_quad + 31) & ~31;
    for (int v = 0; v < nq; v++) s_task[(before & 0xFFFF) + v] = (uint16_t)((tid << 3) | v);
    for (int v = 0; v < nb; v++) s_task[bi0 + (before >> 16) + v] = (uint16_t)((tid << 3) | (nq + v));
    __syncthreads();
    // (a block takes as many units as it has threads, so the multivalents may need several rounds: the warps go
    // round by round on their own - there is no block-wide barrier below)
    const int n_slots = bi0 + n_bi;
    for (int slot0 = tid & ~31; slot0 < n_slots; slot0 += MEI_THREADS) {
    const int slot = slot0 + (tid & 31);
    const bool is_quad = slot < n_quad, is_bi = slot >= bi0 && slot < bi0 + n_bi;
    bool live = is_quad || is_bi;
    ChiasmaList x;
    x.stride = (int64_t)gridDim.x * MEI_THREADS;
    x.pos = a.x_pos + (int64_t)blockIdx.x * MEI_THREADS + tid;
    x.ab = a.x_ab + (int64_t)blockIdx.x * MEI_THREADS + tid;
    x.cap = a.x_cap; x.n = 0; x.overflow = false;
    int n_out = 0;            // chromatids this thread delivers (<= 4: two per quadrivalent, twice with 2NGAMETES)
    uint32_t strands = 0, dests = 0;   // nibble o: the strand chromatid o starts on, the child's homolog it becomes
    int count[4];
    Nib16 seq; seq.v = 0;
    int base = 0, v = 0, quad = 0;
    int64_t phid0 = 0, chid0 = 0;
    double head = 0.0, tail = 0.0, lim_tail = 0.0;
    int32_t child = 0;
    int par = 0, c = 0;
    PhiloxStream rng;
    rng.init(a.model.seed, a.replicate_first);
    const ChromParams* ch = a.chrom;
    // the chiasmata: loops whose trip counts differ from thread to thread ...
    if (live) {
        const int code = s_task[slot], i = code >> 3;
        v = code & 7;
        int j, r;
        unit_decode(a, unit0 + i, par, j, r, c);
        child = a.lev_ind[j];
        const int32_t parent = par == 0 ? a.par0[child] : a.par1[child];
        ch = a.chrom + c;
        head = ch->head; tail = ch->tail;
        phid0 = (((int64_t)c * a.R + r) * a.N + parent) * P;
        chid0 = (((int64_t)c * a.R + r) * a.N + child) * P + (int64_t)par * p2;
        seq.v = s_seq[i];
        quad = s_nquad[i];
        rng.init(a.model.seed, a.replicate_first + (uint32_t)r);
        rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, STAGE_MULTIVALENT + (uint32_t)v);
        if (is_quad) lim_tail = draw_chiasmata<true, KOS, ANR>(x, rng, a.model, *ch);
        else draw_chiasmata<false, KOS, ANR>(x, rng, a.model, *ch);
        if (x.overflow) {
            atomicExch(a.status, PSIM_ECAPACITY);
            atomicMin(a.status + 1, a.batch);
            live = false;
        }
    }
    __syncwarp();
    // ... then the warp runs together again: segregation and the kept gametes are branch-free bit work
    if (live) {
        const uint32_t at = is_quad ? draw_segregation<true>(x, rng, a.model, *ch, lim_tail) : draw_segregation<false>(x, rng, a.model, *ch, lim_tail);
        const Transmission tr = draw_transmission(rng, a.model, (uint32_t)child, (uint32_t)c, (uint32_t)par);
        // Individual.java:297-321: quadrivalent v fills gamete slots 2v, 2v+1; bivalent v' slot 2 * quad + v'
        base = is_quad ? 4 * v : 4 * quad + 2 * (v - quad);
        const int per_gamete = is_quad ? 2 : 1;
        for (int k = 0; k < tr.n_gametes; k++)
            for (int o = 0; o < per_gamete; o++) {
                const int g = tr.gamete[k];
                const int slot = is_quad ? 2 * v + o : 2 * quad + (v - quad);
                strands |= ((at >> (4 * (is_quad ? 2 * g + o : g))) & 15u) << (4 * n_out);
                dests |= (uint32_t)(k * p2 + tr.dest[k].get(slot)) << (4 * n_out);
                n_out++;
            }
    }
    __syncwarp();
    // (the traces are not unrolled: eight inlined copies of trace_chromatid were a fifth of the kernel's code)
    int mine = 0;
#pragma unroll 1
    for (int o = 0; o < n_out; o++) {
        count[o] = trace_chromatid<false>(a, x, (int)((strands >> (4 * o)) & 15u), seq, base, phid0, head, tail, 0);
        mine += count[o];
    }
    // ---- 3. the warp's segments: one reservation, then the same traces write (no block-wide barrier
    // here: the threads of a block finish at very different times)
    int offset = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, offset, o); if ((tid & 31) >= o) offset += y; }
    const int warp_total = __shfl_sync(0xFFFFFFFFu, offset, 31);
    offset -= mine;
    long long warp_base = 0;
    if ((tid & 31) == 0 && warp_total > 0) {
        warp_base = (long long)atomicAdd(a.cursor, (unsigned long long)warp_total);
        if (warp_base + warp_total > a.slab_cap) {
            atomicCAS(a.status, 0, STATUS_SLAB_FULL);
            atomicMin(a.status + 1, a.batch);
            warp_base = -1;
        }
    }
    warp_base = __shfl_sync(0xFFFFFFFFu, warp_base, 0);
    if (warp_base < 0) return;   // (the whole warp)
    int64_t w = warp_base + offset;
#pragma unroll 1
    for (int o = 0; o < n_out; o++) {
        trace_chromatid<true>(a, x, (int)((strands >> (4 * o)) & 15u), seq, base, phid0, head, tail, w);
        a.desc[chid0 + ((dests >> (4 * o)) & 15u)] = make_desc((uint64_t)w, (uint64_t)count[o]);
        w += count[o];
    }
    }
}
