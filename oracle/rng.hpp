// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link or call anything under oracle/.
//
// Random sources for the CPU restatement of PedigreeSim's meiosis path.
//
//  * JavaRandom  — java.util.Random, the reference's only RNG
//                  (reference: src/PedigreeSim/Tools.java:22-27 `rand = new Random(randomSeed)`).
//                  The algorithm is the one the JDK Javadoc specifies (48-bit LCG); it is a
//                  third-party dependency of the reference that is not under /root/reference,
//                  so it is restated here from the published contract and pinned by the
//                  well-known first outputs for seeds 0 and 42 (tests/test_oracle_rng.py).
//  * PhiloxStream — Philox4x32-10 counter-based streams, the free-running mode the
//                  north_star prescribes for the GPU: one stream per
//                  (replicate, individual, chromosome, parent, stage). Own restatement of the
//                  published Philox algorithm (Salmon et al., SC'11), pinned by the Random123
//                  known-answer vectors.
#pragma once
#include <cstdint>

namespace orc {

struct Rng {
    virtual ~Rng() {}
    virtual int next_int(int bound) = 0;   // uniform in [0, bound)
    virtual double next_double() = 0;      // uniform in [0, 1), 53 bits
    virtual bool next_bool() = 0;
    virtual float next_float() = 0;        // uniform in [0, 1), 24 bits (java.util.Random.nextFloat)
    // Select the stream for one unit of work. A no-op for the sequential Java LCG.
    virtual void seek(uint32_t individual, uint32_t chrom, uint32_t parent, uint32_t stage) {
        (void)individual; (void)chrom; (void)parent; (void)stage;
    }
};

// java.util.Random (JDK contract).
struct JavaRandom : Rng {
    uint64_t s;
    static constexpr uint64_t MULT = 0x5DEECE66DULL;
    static constexpr uint64_t MASK = (1ULL << 48) - 1;
    explicit JavaRandom(int64_t seed) { s = ((uint64_t)seed ^ MULT) & MASK; }
    int32_t next(int bits) {
        s = (s * MULT + 0xBULL) & MASK;
        return (int32_t)(uint32_t)(s >> (48 - bits));
    }
    int32_t next_int32() { return next(32); }
    int next_int(int bound) override {
        int32_t r = next(31);
        int32_t m = bound - 1;
        if ((bound & m) == 0) {
            r = (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        } else {
            // Java: for (int u = r; u - (r = u % bound) + m < 0; u = next(31));  (int32 overflow)
            int32_t u = r;
            for (;;) {
                r = u % bound;
                int32_t t = (int32_t)((uint32_t)u - (uint32_t)r + (uint32_t)m);
                if (t >= 0) break;
                u = next(31);
            }
        }
        return r;
    }
    double next_double() override {
        int64_t hi = (int64_t)next(26);
        int64_t lo = (int64_t)next(27);
        return (double)((hi << 27) + lo) * 0x1.0p-53;
    }
    bool next_bool() override { return next(1) != 0; }
    float next_float() override { return next(24) / (float)(1 << 24); }
};

// Philox4x32-10.
struct Philox4x32 {
    static inline void round(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c[1] ^ k[0];
        const uint32_t n2 = hi0 ^ c[3] ^ k[1];
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    }
    static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k[2] = {key[0], key[1]};
        for (int r = 0; r < 10; r++) {
            if (r) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
            round(c, k);
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};

// Stream layout shared (by specification, not by code) with the CUDA kernels — DESIGN.md §RNG:
//   key = { (u32) seed, (u32) replicate }
//   ctr = { block#, individual, chrom | parent << 20 | stage << 21, 0x5053494d ("PSIM") }
//   the stream yields the four words of block 0, then block 1, ...
//   next_double = ((a >> 5) * 2^26 + (b >> 6)) * 2^-53  with a, b the next two words
//   next_int(n) = Lemire multiply-shift with rejection (unbiased), one word per attempt
//   next_bool   = top bit of the next word;  next_float = top 24 bits of the next word * 2^-24
struct PhiloxStream : Rng {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    int have;
    PhiloxStream(uint32_t seed, uint32_t replicate) {
        key[0] = seed; key[1] = replicate;
        ctr[0] = ctr[1] = ctr[2] = 0; ctr[3] = 0x5053494du; have = 0;
    }
    void seek(uint32_t individual, uint32_t chrom, uint32_t parent, uint32_t stage) override {
        ctr[0] = 0; ctr[1] = individual;
        ctr[2] = chrom | (parent << 20) | (stage << 21);
        have = 0;
    }
    uint32_t next_u32() {
        if (have == 0) { Philox4x32::block(ctr, key, buf); ctr[0]++; have = 4; }
        return buf[4 - have--];
    }
    int next_int(int bound) override {
        const uint32_t n = (uint32_t)bound;
        uint64_t m = (uint64_t)next_u32() * n;
        uint32_t l = (uint32_t)m;
        if (l < n) {
            const uint32_t t = (0u - n) % n;
            while (l < t) { m = (uint64_t)next_u32() * n; l = (uint32_t)m; }
        }
        return (int)(m >> 32);
    }
    double next_double() override {
        const uint32_t a = next_u32(), b = next_u32();
        return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * 0x1.0p-53;
    }
    bool next_bool() override { return (next_u32() >> 31) != 0; }
    float next_float() override { return (float)(next_u32() >> 8) / (float)(1 << 24); }   // top 24 bits of one word
};

}  // namespace orc
