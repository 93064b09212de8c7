// K2 (generic path) and the _allAlleledose rows
// (part of libpsim: included by psim.cu inside namespace psim)
#pragma once

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
        const uint64_t b = desc_begin(d);   // (40 bits: the slab may hold more than 2^32 segments)
        const uint32_t n = desc_count(d);
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
        const uint64_t b = desc_begin(d);   // (40 bits: the slab may hold more than 2^32 segments)
        const uint32_t n = desc_count(d);
        uint32_t lo = 0, hi = n;
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.run_lb[b + mid] <= l) lo = mid; else hi = mid; }
        int f = a.run_founder[b + lo];
        if (codes) f = codes[f];
        if (r0 + f < r1) a.out[(r0 + f) * a.pitch + q] += 1;
    }
}

