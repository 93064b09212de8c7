// Philox4x32-10 counter streams for the meiosis kernel (DESIGN.md "RNG").
//   key = { low 32 bits of the seed, replicate }
//   ctr = { block#, individual, chrom | parent << 20 | stage << 21, "PSIM" ^ high 32 bits of the seed }
// Streams of one (replicate, individual, chromosome, parent): stage 0 = the pairing configuration,
// stage 1 + v = multivalent v (chiasmata, segregation), stage 1021 = the homolog order of the kept
// gametes, stage 1022 = the FDR coin of a meiosis (shared by its chromosomes). The reference has ONE java.util.Random for
// everything (Tools.java:22-27); here every unit of work owns a stream, so results do not depend on
// launch geometry, batch size or GPU count. The distributions the reference draws by rejection (a
// partner that is still free, a non-sister chromatid, an arm that is not finished) are sampled
// directly from these words: the law of every outcome is the reference's, the draw sequence is not.
//
// The header is plain C++ apart from PSIM_HD, so that the tests can compile the same sampler for
// the host and hold the kernel against it bit for bit.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define PSIM_HD __host__ __device__ __forceinline__
#define PSIM_HD_NOINLINE __host__ __device__ __forceinline__
#else
#define PSIM_HD inline
#define PSIM_HD_NOINLINE inline
#endif

namespace psim {

constexpr uint32_t STAGE_PAIRING = 0;
constexpr uint32_t STAGE_MULTIVALENT = 1;    // + number of the multivalent
constexpr uint32_t STAGE_ORDER = 1021;
constexpr uint32_t STAGE_UNREDUCED = 1022;   // the FDR coin of a meiosis (chromosome 0)

PSIM_HD int popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
PSIM_HD int lowest_bit(uint32_t x) {   // index of the lowest set bit of a non-zero word
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
PSIM_HD int nth_set_bit(uint32_t mask, int n) {   // index of the n-th (0-based) set bit
#ifdef __CUDA_ARCH__
    return (int)__fns(mask, 0, n + 1);
#else
    for (int i = 0; i < n; i++) mask &= mask - 1;
    return __builtin_ctz(mask);
#endif
}

// One Philox4x32-10 block. By value in, by value out: the device code CALLS it (registers in, registers out) instead
// of carrying ten unrolled rounds at every draw site of the meiosis kernel, whose divergent warps otherwise wait on
// the instruction cache (inlined: 26 % more code, simulate 8 - 16 % slower; -DPSIM_PHILOX_INLINE brings that back).
struct PhiloxBlock { uint32_t x0, x1, x2, x3; };
#ifndef PSIM_PHILOX_INLINE
#define PSIM_PHILOX_CALL 1
#endif
#if defined(__CUDACC__) && defined(PSIM_PHILOX_CALL)
__host__ __device__ __noinline__
#else
PSIM_HD
#endif
PhiloxBlock philox4x32_10(uint32_t ka, uint32_t kb, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
        if (r) { ka += 0x9E3779B9u; kb += 0xBB67AE85u; }
        const uint64_t p0 = (uint64_t)0xD2511F53u * x0, p1 = (uint64_t)0xCD9E8D57u * x2;
        const uint32_t n0_ = (uint32_t)(p1 >> 32) ^ x1 ^ ka;
        const uint32_t n2_ = (uint32_t)(p0 >> 32) ^ x3 ^ kb;
        x0 = n0_; x1 = (uint32_t)p1; x2 = n2_; x3 = (uint32_t)p0;
    }
    return PhiloxBlock{x0, x1, x2, x3};
}

struct PhiloxStream {
    uint32_t k0, k1;
    uint32_t c0, c1, c2, c3;
    uint32_t b0, b1, b2, b3;
    int have;

    PSIM_HD void init(uint64_t seed, uint32_t replicate) {
        k0 = (uint32_t)seed; k1 = replicate; c3 = 0x5053494du ^ (uint32_t)(seed >> 32);
        have = 0; c0 = c1 = c2 = 0;
        b0 = b1 = b2 = b3 = 0;
    }
    PSIM_HD void seek(uint32_t individual, uint32_t chrom, uint32_t parent, uint32_t stage) {
        c0 = 0; c1 = individual; c2 = chrom | (parent << 20) | (stage << 21); have = 0;
    }
    PSIM_HD void refill() {
        const PhiloxBlock b = philox4x32_10(k0, k1, c0, c1, c2, c3);
        b0 = b.x0; b1 = b.x1; b2 = b.x2; b3 = b.x3;
        c0++;
        have = 4;
    }
    PSIM_HD uint32_t next_u32() {
        if (have == 0) refill();
        const uint32_t r = have == 4 ? b0 : have == 3 ? b1 : have == 2 ? b2 : b3;
        have--;
        return r;
    }
    // uniform integer in [0, n): multiply-shift with the (practically never taken) rejection that
    // makes it exactly unbiased
    PSIM_HD int next_below(int bound) {
        const uint32_t n = (uint32_t)bound;
        uint64_t m = (uint64_t)next_u32() * n;
        uint32_t l = (uint32_t)m;
        if (l < n) {
            const uint32_t t = (0u - n) % n;
            while (l < t) { m = (uint64_t)next_u32() * n; l = (uint32_t)m; }
        }
        return (int)(m >> 32);
    }
    // the top `n` bits of one word (1 <= n <= 32): n fair coins
    PSIM_HD uint32_t next_bits(int n) { return next_u32() >> (32 - n); }
    // uniform in [0, 1), 53 bits
    PSIM_HD double next_unit() {
        const uint32_t a = next_u32(), b = next_u32();
        return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * 0x1.0p-53;
    }
    // uniform in (0, 1], 53 bits: the argument of a logarithm
    PSIM_HD double next_open() {
        const uint32_t a = next_u32(), b = next_u32();
        return (double)((((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) + 1ull) * 0x1.0p-53;
    }
    PSIM_HD float next_float() { return (float)(next_u32() >> 8) * 0x1.0p-24f; }   // java.util.Random.nextFloat: 24 bits
};

}  // namespace psim
