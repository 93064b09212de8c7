// Philox4x32-10 counter streams for the meiosis kernels (DESIGN.md "RNG").
//   key = { (u32) seed, replicate }
//   ctr = { block#, individual, chrom | parent << 20 | stage << 21, 0x5053494d }
// One stream per (replicate, individual, chromosome, parent, stage): stage 0 = pairing
// configuration, 1 + v = multivalent v, 1023 = gamete homolog shuffle, 1022 = FDR draw. The draw order inside a
// stream is the reference's (SURVEY.md §8a "RNG call sites"); the primitives replace
// java.util.Random's (Tools.java:22-27): next_int is Lemire's unbiased multiply-shift,
// next_double takes 53 bits from two words, next_bool the top bit of one word.
#pragma once
#include <cstdint>

namespace psim {

constexpr uint32_t STAGE_CONFIG = 0;
constexpr uint32_t STAGE_SHUFFLE = 1023;
constexpr uint32_t STAGE_UNREDUCED = 1022;   // the FDR draw of a meiosis (chromosome 0)

struct PhiloxStream {
    uint32_t k0, k1;
    uint32_t c0, c1, c2;
    uint32_t b0, b1, b2, b3;
    uint32_t n0, n1, n2, n3;   // the block after the current one, computed ahead by prefetch()
    int have;
    bool has_next;

    __device__ __forceinline__ void init(uint32_t seed, uint32_t replicate) { k0 = seed; k1 = replicate; have = 0; has_next = false; c0 = c1 = c2 = 0; }
    __device__ __forceinline__ void seek(uint32_t individual, uint32_t chrom, uint32_t parent, uint32_t stage) {
        c0 = 0; c1 = individual; c2 = chrom | (parent << 20) | (stage << 21); have = 0; has_next = false;
    }
    __device__ __forceinline__ void block(uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
        uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = 0x5053494du;
        uint32_t ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            if (r) { ka += 0x9E3779B9u; kb += 0xBB67AE85u; }
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t n0_ = hi1 ^ x1 ^ ka;
            const uint32_t n2_ = hi0 ^ x3 ^ kb;
            x0 = n0_; x1 = lo1; x2 = n2_; x3 = lo0;
        }
        o0 = x0; o1 = x1; o2 = x2; o3 = x3;
        c0++;
    }
    // Called at the top of the chiasma loops: every lane of the warp computes its next block at ONE
    // place in the code (the draws inside the loop body then only move registers) and the ten
    // dependent rounds overlap the body's log / compare chains instead of stalling a draw. The
    // sequence of words a stream yields is unchanged.
    __device__ __forceinline__ void prefetch() {
        if (!has_next) { block(n0, n1, n2, n3); has_next = true; }
    }
    __device__ __forceinline__ void refill() {
        if (has_next) { b0 = n0; b1 = n1; b2 = n2; b3 = n3; has_next = false; }
        else block(b0, b1, b2, b3);
        have = 4;
    }
    __device__ __forceinline__ uint32_t next_u32() {
        if (have == 0) refill();
        const uint32_t r = have == 4 ? b0 : have == 3 ? b1 : have == 2 ? b2 : b3;
        have--;
        return r;
    }
    __device__ __forceinline__ int next_int(int bound) {
        const uint32_t n = (uint32_t)bound;
        uint64_t m = (uint64_t)next_u32() * n;
        uint32_t l = (uint32_t)m;
        if (l < n) {
            const uint32_t t = (0u - n) % n;
            while (l < t) { m = (uint64_t)next_u32() * n; l = (uint32_t)m; }
        }
        return (int)(m >> 32);
    }
    __device__ __forceinline__ double next_double() {
        const uint32_t a = next_u32(), b = next_u32();
        return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * 0x1.0p-53;
    }
    __device__ __forceinline__ bool next_bool() { return (next_u32() >> 31) != 0; }
    __device__ __forceinline__ float next_float() { return (float)(next_u32() >> 8) * 0x1.0p-24f; }   // java.util.Random.nextFloat: 24 bits
};

}  // namespace psim
