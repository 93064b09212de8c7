// haplostruct exchange helpers and per-locus code tables
// (part of libpsim: included by psim.cu inside namespace psim)
#pragma once

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

