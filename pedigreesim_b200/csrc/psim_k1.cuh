// K1 — meiosis kernels: founders, pairing configuration, multivalents, assembly
// (part of libpsim: included by psim.cu inside namespace psim)
#pragma once

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
    int direct_diploid;          // ploidy 2: no configuration kernel and no task lists - thread t of the bivalent
                                 // kernel IS unit t and draws the unit's configuration (one shuffle of two) itself
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
    if (t < total && !(a.model.unreduced && (t & 1))) {   // 2NGAMETES: the individual is parent 0's 2n gamete alone
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

// KOS / ANR: MAPFUNCTION=KOSAMBI and ALLOWNORECOMBINATION as compile-time constants, so that the code
// of the other model (gamma sampler, pow, retry loop) is not in the kernel at all: the kernels are
// instruction-cache bound (ncu: 3.2 'no instruction' stall cycles per issue with everything inlined).
template <bool QUAD, bool KOS, bool ANR>
__global__ void __launch_bounds__(128) k_meiosis_mv(PlanArgs a) {
    const int64_t ti = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool direct = !QUAD && a.direct_diploid;
    if (direct ? ti >= (int64_t)a.C * a.R * a.n_lev * 2 : ti >= (int64_t)a.task_count[QUAD ? 1 : 0]) return;
    const int64_t code = direct ? ti * 8 : (QUAD ? a.task_quad : a.task_bi)[ti];
    const int v = (int)(code & 7);
    int par, j, r, c;
    plan_decode(a, code >> 3, par, j, r, c);
    if (direct && a.model.unreduced && par == 1) return;   // 2NGAMETES: parent 0's gamete alone
    const int P = a.P, p2 = P / 2;
    const int32_t child = a.lev_ind[j];
    const int32_t parent = par == 0 ? a.par0[child] : a.par1[child];
    const ChromParams ch = a.chrom[c];
    ModelParams m = a.model;
    m.chiasma_interference = KOS;
    m.allow_no_recomb = ANR;
    const int64_t phid0 = (((int64_t)c * a.R + r) * a.N + parent) * P;
    const int64_t q = (((int64_t)r * a.N + child) * a.C + c) * 2 + par;
    PhiloxStream rng;
    rng.init(m.seed, a.replicate_first + (uint32_t)r);
    int8_t own_seq[2] = {0, 1};
    if (direct) {   // Individual.calcChromConfig for ploidy 2 (Individual.java:603-608): shuffle of (0, 1), no quadrivalent
        rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, STAGE_CONFIG);
        shuffle_small(rng, own_seq, 2);
        a.cfg_quad[q] = 0;
        a.cfg_seq[q * 2] = own_seq[0];
        a.cfg_seq[q * 2 + 1] = own_seq[1];
    }
    const int quad = direct ? 0 : a.cfg_quad[q];
    const int8_t* chromseq = direct ? own_seq : a.cfg_seq + q * P;
    // Each transmitted chromatid is traced straight into the plan entry of the child homolog it
    // will occupy after the gamete's homolog shuffle. The shuffle draws come from their own stream:
    // slot h of the transmitted gamete receives pre-shuffle slot shuf[h] (Individual.java:325-335).
    const int64_t u0 = (((int64_t)c * a.R + r) * a.n_lev + j) * P + (int64_t)par * p2;
    // Transmitted gametes: gamete 0 (Individual.java:258-259), plus with 2NGAMETES the second half
    // of the unreduced gamete (:350-382): gamete 1 (SDR) or bottom gamete 3 / 2 (FDR, one draw per
    // meiosis); their chromatids fill the child's homologs P/2 .. P-1.
    const int unred = m.unreduced;
    int g_second = 0;
    if (unred == 2) g_second = 1;
    else if (unred == 1) {
        rng.seek((uint32_t)child, 0u, (uint32_t)par, STAGE_UNREDUCED);
        g_second = rng.next_float() < 0.5f ? 3 : 2;
    }
    const int n_gam = unred ? 2 : 1;
    int8_t dest_of_slot[2][MAXP / 2];  // [transmitted gamete][pre-shuffle slot] -> final homolog position in that gamete
    rng.seek((uint32_t)child, (uint32_t)c, (uint32_t)par, STAGE_SHUFFLE);
    for (int g = 0; g <= g_second; g++) {   // the four gametes are shuffled in order on one stream
        int8_t shuf[MAXP / 2];
        for (int s = 0; s < p2; s++) shuf[s] = (int8_t)s;
        shuffle_small(rng, shuf, p2);
        if (g == 0) for (int h = 0; h < p2; h++) dest_of_slot[0][shuf[h]] = (int8_t)h;
        if (g == g_second) for (int h = 0; h < p2; h++) dest_of_slot[1][shuf[h]] = (int8_t)h;
    }

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
    int8_t pattern_slot[8];
    select_gametes(x, rng, k, ch, centro, pattern_slot);
    constexpr int n_out = QUAD ? 2 : 1;
    for (int go = 0; go < n_gam * n_out; go++) {
        const int gi = go / n_out, o = go % n_out;
        const int g = gi == 0 ? 0 : g_second;
        // Individual.java:297-321: quadrivalent v fills gamete slots 2v, 2v+1; bivalent v' fills firstbi/2 + v'
        const int pre_slot = QUAD ? 2 * v + o : firstbi / 2 + (v - quad);
        const int64_t u = u0 + gi * p2 + dest_of_slot[gi][pre_slot];
        double* ppos = a.plan_pos + u * a.piece_cap;
        uint8_t* phom = a.plan_hom + u * a.piece_cap;
        // Multivalent.slotsToPatterns (:332-357) for the selected chromatid
        int cur = pattern_slot[QUAD ? 2 * g + o : g];
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
    int64_t slab_cap;          // segments the slab holds: a batch that does not fit is not written at all
    const int* error_flag;     // set by the plan kernels: the plan is incomplete, nothing is assembled
    int unreduced;             // ModelParams::unreduced
};

__global__ void __launch_bounds__(128) k_meiosis_assemble(AssembleArgs a) {
    const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)a.C * a.R * a.n_lev * a.P;
    if (u >= total) return;
    if (*a.error_flag || a.slab_base + a.plan_off[total] > a.slab_cap) return;   // the host reports the error / grows the slab and launches again
    const int P = a.P, p2 = P / 2;
    const int h = (int)(u % P);
    const int j = (int)((u / P) % a.n_lev);
    const int r = (int)((u / ((int64_t)P * a.n_lev)) % a.R);
    const int c = (int)(u / ((int64_t)P * a.n_lev * a.R));
    const int32_t child = a.lev_ind[j];
    const int32_t parent = (h < p2 || a.unreduced) ? a.par0[child] : a.par1[child];   // 2NGAMETES: all P homologs from parent 0
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

