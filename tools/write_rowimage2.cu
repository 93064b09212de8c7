// Store-mechanism sweep for the row-image emission kernel (round 2): the same band / row-buffer
// structure as tools/write_rowimage.cu, varying HOW a finished row leaves shared memory:
//   mech 0  one cp.async.bulk per table and row, issued by one thread            (what k_emit_rows does)
//   mech 1  the row cut into `pieces` bulk stores issued by `pieces` threads
//   mech 2  as 0/1 with an L2 cache-policy hint on the bulk store (evict_first / evict_last / no_allocate-like)
//   mech 3  ordinary coalesced 16-byte st.global from shared memory by all threads (default / .cs / .wt)
// `delay` dependent IMADs per thread and row stand in for the per-row work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_rowimage2 write_rowimage2.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t spin(uint32_t x, int n) {
    for (int i = 0; i < n; i++) x = x * 1664525u + 1013904223u;
    return x;
}

struct Args {
    uint8_t* f; uint32_t fpitch; uint8_t* d; uint32_t dpitch;
    int rows_total, rows_per_band; int* counter;
    int delay, pieces, nbuf, mech, hint, interleave;
};

__device__ __forceinline__ void bulk_plain(void* g, const void* s, uint32_t n) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"((uint32_t)__cvta_generic_to_shared(s)), "r"(n) : "memory");
}
__device__ __forceinline__ void bulk_hint(void* g, const void* s, uint32_t n, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(g),
                 "r"((uint32_t)__cvta_generic_to_shared(s)), "r"(n), "l"(pol) : "memory");
}

template <int NBUF>
__global__ void __launch_bounds__(512) k_rowimage(Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_band;
    const uint32_t img = a.fpitch + a.dpitch;
    const int n_bands = (a.rows_total + a.rows_per_band - 1) / a.rows_per_band;
    uint64_t pol = 0;
    if (a.hint == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (a.hint == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (a.hint == 3) asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
    if (a.hint == 4) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    for (;;) {
        if (threadIdx.x == 0) s_band = atomicAdd(a.counter, 1);
        __syncthreads();
        const int band = s_band;
        __syncthreads();
        if (band >= n_bands) break;
        const int G = a.interleave;
        const int grp = band / G, mem = band % G;
        const int r0 = G == 1 ? band * a.rows_per_band : grp * G * a.rows_per_band + mem;
        const int r1 = G == 1 ? min(r0 + a.rows_per_band, a.rows_total) : min((grp + 1) * G * a.rows_per_band, a.rows_total);
        uint32_t x = band * 977u + threadIdx.x;
        int buf = 0;
        for (int r = r0; r < r1; r += G) {
            if (a.mech != 3 && threadIdx.x < a.pieces) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
            __syncthreads();
            uint8_t* row = smem + (size_t)buf * img;
            x = spin(x, a.delay);
            for (uint32_t i = threadIdx.x * 4; i < a.dpitch; i += blockDim.x * 4) *reinterpret_cast<uint32_t*>(row + a.fpitch + i) = x + i;
            row[(x >> 8) % a.fpitch] = (uint8_t)x;
            if (a.mech != 3) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (a.mech == 3) {
                uint8_t* gf = a.f + (size_t)r * a.fpitch;
                uint8_t* gd = a.d + (size_t)r * a.dpitch;
                for (uint32_t i = threadIdx.x * 16; i < img; i += blockDim.x * 16) {
                    const uint4 v = *reinterpret_cast<const uint4*>(row + i);
                    uint4* dst = reinterpret_cast<uint4*>(i < a.fpitch ? gf + i : gd + (i - a.fpitch));
                    if (a.hint == 0) *dst = v;
                    else if (a.hint == 1) __stcs(dst, v);
                    else if (a.hint == 2) __stwt(dst, v);
                    else __stcg(dst, v);
                }
            } else if (threadIdx.x < a.pieces) {
                const uint32_t fp = (a.fpitch / a.pieces + 127) / 128 * 128, dp = (a.dpitch / a.pieces + 127) / 128 * 128;
                const uint32_t fo = threadIdx.x * fp, dof = threadIdx.x * dp;
                if (fo < a.fpitch) {
                    const uint32_t n = min(fp, a.fpitch - fo);
                    if (a.hint) bulk_hint(a.f + (size_t)r * a.fpitch + fo, row + fo, n, pol); else bulk_plain(a.f + (size_t)r * a.fpitch + fo, row + fo, n);
                }
                if (dof < a.dpitch) {
                    const uint32_t n = min(dp, a.dpitch - dof);
                    if (a.hint) bulk_hint(a.d + (size_t)r * a.dpitch + dof, row + a.fpitch + dof, n, pol);
                    else bulk_plain(a.d + (size_t)r * a.dpitch + dof, row + a.fpitch + dof, n);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf = buf + 1 == NBUF ? 0 : buf + 1;
        }
    }
    if (a.mech != 3 && threadIdx.x < a.pieces) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const int rows = 24000;
    const uint32_t fpitch = 40064, dpitch = 10112;
    uint8_t *f, *d;
    int* counter;
    CK(cudaMalloc(&f, (size_t)rows * fpitch));
    CK(cudaMalloc(&d, (size_t)rows * dpitch));
    CK(cudaMalloc(&counter, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double bytes = (double)rows * (fpitch + dpitch);
    struct Res { std::string name; float best, mean; };
    std::vector<Res> all;
    auto time = [&](const std::string& name, auto fn) {
        for (int i = 0; i < 2; i++) fn();
        float best = 1e9, sum = 0;
        for (int i = 0; i < 6; i++) {
            CK(cudaEventRecord(e0));
            fn();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
            sum += ms;
        }
        all.push_back({name, best, sum / 6});
    };
    CK(cudaFuncSetAttribute(k_rowimage<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (fpitch + dpitch)));
    CK(cudaFuncSetAttribute(k_rowimage<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (fpitch + dpitch)));
    CK(cudaFuncSetAttribute(k_rowimage<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (fpitch + dpitch)));
    auto run = [&](int mech, int hint, int pieces, int nbuf, int delay, int band, int interleave, int threads) {
        char nm[160];
        snprintf(nm, 160, "mech %d hint %d pieces %d nbuf %d delay %3d band %3d interleave %3d thr %3d", mech, hint, pieces, nbuf, delay, band, interleave, threads);
        Args a{f, fpitch, d, dpitch, rows, band, counter, delay, pieces, nbuf, mech, hint, interleave};
        time(nm, [&] {
            CK(cudaMemsetAsync(counter, 0, 4));
            const size_t sm = (size_t)nbuf * (fpitch + dpitch);
            if (nbuf == 2) k_rowimage<2><<<148, threads, sm>>>(a);
            else if (nbuf == 3) k_rowimage<3><<<148, threads, sm>>>(a);
            else k_rowimage<4><<<148, threads, sm>>>(a);
            CK(cudaGetLastError());
        });
    };
    // memset reference
    time("cudaMemsetAsync of both tables", [&] { CK(cudaMemsetAsync(f, 1, (size_t)rows * fpitch)); CK(cudaMemsetAsync(d, 1, (size_t)rows * dpitch)); });
    for (int delay : {0, 100, 200, 400})
        for (int nbuf : {2, 3, 4}) {
            run(0, 0, 1, nbuf, delay, 163, 1, 512);
            for (int hint : {1, 2, 3, 4}) run(2, hint, 1, nbuf, delay, 163, 1, 512);
            for (int pieces : {2, 4, 8}) run(1, 0, pieces, nbuf, delay, 163, 1, 512);
            run(2, 1, 4, nbuf, delay, 163, 1, 512);
        }
    for (int delay : {0, 200})
        for (int hint : {0, 1, 2, 3}) run(3, hint, 1, 2, delay, 163, 1, 512);
    for (int band : {20, 41, 82, 163})
        for (int il : {1, 148}) run(0, 0, 1, 3, 200, band, il, 512);
    CK(cudaDeviceSynchronize());
    for (const Res& r : all) printf("%-96s best %7.1f us %7.1f GB/s   mean %7.1f us\n", r.name.c_str(), r.best * 1e3, bytes / r.best / 1e6, r.mean * 1e3);
    return 0;
}
