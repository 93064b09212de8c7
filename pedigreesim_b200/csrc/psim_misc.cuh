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


// Rows with a 16-byte-aligned pitch -> the same rows back to back (pitch == width, any width): the
// layout of a dense host table whose row length is not a multiple of 16. One 16-byte word of the
// destination per thread and step: two aligned loads and a funnel shift when the source bytes are
// not aligned, byte by byte for the one word per row that straddles two source rows.
__global__ void __launch_bounds__(256) k_repack_dense(const uint8_t* __restrict__ src, int64_t spitch, uint8_t* __restrict__ dst,
                                                      int64_t width, int64_t rows) {
    const int64_t total = rows * width;
    const int64_t step = (int64_t)gridDim.x * blockDim.x * 16;
    for (int64_t o = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 16; o < total; o += step) {
        const int64_t row = o / width, col = o - row * width;
        if (col + 16 <= width) {
            const uint8_t* p = src + row * spitch + col;
            const int mis = (int)(col & 15);
            const uint4* q = reinterpret_cast<const uint4*>(p - mis);
            const uint4 lo = __ldg(q);
            uint4 out = lo;
            if (mis) {
                const uint4 hi = __ldg(q + 1);   // still inside the padded source row
                const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                const int wo = mis >> 2, sh = (mis & 3) * 8;
                uint32_t r[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    uint32_t a = 0, b = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) if (wo == j) { a = w[j + k]; b = w[j + k + 1 < 8 ? j + k + 1 : 7]; }
                    r[k] = __funnelshift_r(a, b, sh);
                }
                out = make_uint4(r[0], r[1], r[2], r[3]);
            }
            *reinterpret_cast<uint4*>(dst + o) = out;
        } else {
            for (int b = 0; b < 16 && o + b < total; b++) {
                const int64_t oo = o + b, rr = oo / width;
                dst[oo] = src[rr * spitch + (oo - rr * width)];
            }
        }
    }
}

// Founder alleles of imported segments must be 0 .. founderAlleleCount-1 (Main.java:1235); host[0] = 1 + an offender.
__global__ void k_check_founders(const int32_t* __restrict__ founder, int64_t n, int count, volatile int64_t* host) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n && (founder[t] < 0 || founder[t] >= count)) host[0] = (int64_t)founder[t] + 1;
}

// Small device values the host needs after the next stream synchronisation, stored straight into
// mapped page-locked memory (see psim_ctx_impl::scalars_h).
__global__ void k_publish(volatile int64_t* host, const int64_t* v0, const int* v1, const int* v2) {
    if (threadIdx.x == 0) {
        if (v0) host[0] = *v0;
        if (v1) host[1] = *v1;
        if (v2) host[2] = *v2;
        __threadfence_system();
    }
}

// ------------------------------------------------------------------------------------------
// One pedigree over several GPUs (psim_simulate_sharded): the haplostructs a rank shares after a
// generation travel as one byte payload
//   [first segment of every homolog, and the total: u32 x (n_h + 1)] [founder: i32 x n_seg] [pos: f64 x n_seg]
// (each part padded to 8 bytes) with the homologs in the order [individual of the list][chromosome][homolog]
// (replicate_count is 1). The receiver needs no scan: a homolog's segments start at slab base + offset.
__device__ __forceinline__ int64_t listed_hid(int64_t t, const int32_t* __restrict__ list, int C, int64_t N, int P) {
    const int h = (int)(t % P);
    const int c = (int)((t / P) % C);
    const int64_t i = list[t / ((int64_t)P * C)];
    return ((int64_t)c * N + i) * P + h;
}
__global__ void k_pack_counts(int64_t n_h, const int32_t* __restrict__ list, int C, int64_t N, int P, const uint64_t* __restrict__ desc,
                              int64_t* __restrict__ cnt64) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t > n_h) return;
    cnt64[t] = t < n_h ? desc_count(desc[listed_hid(t, list, C, N, P)]) : 0;   // (entry n_h: the scan leaves the total there)
}
__global__ void k_pack_segments(int64_t n_h, const int32_t* __restrict__ list, int C, int64_t N, int P, const uint64_t* __restrict__ desc,
                                const int64_t* __restrict__ off, const double* __restrict__ seg_pos, const int32_t* __restrict__ seg_founder,
                                uint32_t* __restrict__ out_off, int32_t* __restrict__ out_founder, double* __restrict__ out_pos) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t > n_h) return;
    int64_t w = off[t];
    out_off[t] = (uint32_t)w;
    if (t == n_h) return;
    const uint64_t d = desc[listed_hid(t, list, C, N, P)];
    const int64_t b = (int64_t)desc_begin(d);
    const int n = (int)desc_count(d);
    for (int s = 0; s < n; s++, w++) { out_pos[w] = seg_pos[b + s]; out_founder[w] = seg_founder[b + s]; }
}
__global__ void k_unpack_desc(int64_t n_h, const int32_t* __restrict__ list, int C, int64_t N, int P, const uint32_t* __restrict__ in_off,
                              int64_t slab_base, uint64_t* __restrict__ desc) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_h) return;
    desc[listed_hid(t, list, C, N, P)] = make_desc((uint64_t)(slab_base + in_off[t]), (uint64_t)(in_off[t + 1] - in_off[t]));
}
// The same when a rank shares everything it has just simulated: the segments travel straight out of the slab, headed by
// the homologs' descriptors relative to the first of them.
__global__ void k_pack_desc(int64_t n_h, const int32_t* __restrict__ list, int C, int64_t N, int P, const uint64_t* __restrict__ desc,
                            int64_t slab_from, uint64_t* __restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_h) return;
    const uint64_t d = desc[listed_hid(t, list, C, N, P)];
    out[t] = make_desc(desc_begin(d) - (uint64_t)slab_from, desc_count(d));
}
__global__ void k_unpack_desc_rel(int64_t n_h, const int32_t* __restrict__ list, int C, int64_t N, int P, const uint64_t* __restrict__ in,
                                  int64_t slab_base, uint64_t* __restrict__ desc) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_h) return;
    desc[listed_hid(t, list, C, N, P)] = make_desc(desc_begin(in[t]) + (uint64_t)slab_base, desc_count(in[t]));
}
// (device value -> mapped page-locked host memory, like k_publish)
__global__ void k_publish_i64(volatile int64_t* host, const int64_t* v, int n) {
    if ((int)threadIdx.x < n) host[threadIdx.x] = v[threadIdx.x];
    __threadfence_system();
}
