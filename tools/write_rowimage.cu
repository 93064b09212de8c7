// Store-pattern experiment for a row-image emission kernel: every block owns a band of consecutive
// locus rows, keeps the current row of the founder table (40 KB) and of the dose table (10 KB) in
// shared memory and sends each finished row to global memory as ONE contiguous TMA bulk store
// (cp.async.bulk shared -> global). `delay` dependent ALU operations per thread and row stand in
// for the per-row work (dose recomputation); `jitter` desynchronises the blocks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_rowimage write_rowimage.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t spin(uint32_t x, int n) {
    for (int i = 0; i < n; i++) x = x * 1664525u + 1013904223u;
    return x;
}

// mode 0: one bulk store per table and row; mode 1: the row is cut into `pieces` bulk stores issued by different threads
__global__ void __launch_bounds__(512) k_rowimage(uint8_t* f, uint32_t fpitch, uint8_t* d, uint32_t dpitch, int rows_total,
                                                  int rows_per_band, int* counter, int delay, int jitter, int pieces, int nbuf, int G) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int s_band;
    const uint32_t img = fpitch + dpitch;   // bytes of one row image (both multiples of 128)
    const int n_bands = (rows_total + rows_per_band - 1) / rows_per_band;
    for (;;) {
        if (threadIdx.x == 0) s_band = atomicAdd(counter, 1);
        __syncthreads();
        const int band = s_band;
        __syncthreads();
        if (band >= n_bands) break;
        // G > 1: the G bands of a group deal their rows round-robin (concurrently written rows are neighbours)
        const int grp = band / G, mem = band % G;
        const int r0 = G == 1 ? band * rows_per_band : grp * G * rows_per_band + mem;
        const int r1 = G == 1 ? min(r0 + rows_per_band, rows_total) : min((grp + 1) * G * rows_per_band, rows_total);
        const int rstep = G;
        uint32_t x = spin(band * 977u + threadIdx.x, jitter ? (band * 37 % 11) * jitter : 0);
        int buf = 0;
        for (int r = r0; r < r1; r += rstep) {
            // the buffer we are about to touch must have been read by the bulk engine
            if (threadIdx.x < pieces) {
                if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            }
            __syncthreads();
            uint8_t* row = smem + (size_t)buf * img;
            x = spin(x, delay);
            // touch the image: a few patched cells and the whole dose row (what the real kernel rewrites every row)
            for (uint32_t i = threadIdx.x * 4; i < dpitch; i += blockDim.x * 4) *reinterpret_cast<uint32_t*>(row + fpitch + i) = x + i;
            row[(x >> 8) % fpitch] = (uint8_t)x;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x < pieces) {
                const uint32_t fp = (fpitch / pieces + 127) / 128 * 128, dp = (dpitch / pieces + 127) / 128 * 128;
                const uint32_t fo = threadIdx.x * fp, dof = threadIdx.x * dp;
                if (fo < fpitch) {
                    const uint32_t n = min(fp, fpitch - fo);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(f + (size_t)r * fpitch + fo),
                                 "r"((uint32_t)__cvta_generic_to_shared(row + fo)), "r"(n) : "memory");
                }
                if (dof < dpitch) {
                    const uint32_t n = min(dp, dpitch - dof);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + (size_t)r * dpitch + dof),
                                 "r"((uint32_t)__cvta_generic_to_shared(row + fpitch + dof)), "r"(n) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf = buf + 1 == nbuf ? 0 : buf + 1;
        }
    }
    if (threadIdx.x < pieces) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
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
    auto time = [&](const char* name, auto fn) {
        for (int i = 0; i < 2; i++) fn();
        float best = 1e9, sum = 0;
        for (int i = 0; i < 8; i++) {
            CK(cudaEventRecord(e0));
            fn();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
            sum += ms;
        }
        printf("%-64s best %7.1f us %7.1f GB/s   mean %7.1f us\n", name, best * 1e3, bytes / best / 1e6, sum / 8 * 1e3);
    };
    CK(cudaFuncSetAttribute(k_rowimage, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (fpitch + dpitch)));
    for (int rep = 0; rep < 2; rep++)
    for (int G : {1, 4, 8, 16, 37, 148})
        for (int delay : {0, 200}) {
            const int nbuf = 3, threads = 512, band = 163, pieces = 1, jitter = delay ? 400 : 0;
            char nm[128];
            snprintf(nm, 128, "rowimage nbuf 3 thr 512 band 163 interleave %3d delay %3d jitter %3d", G, delay, jitter);
            time(nm, [&] {
                CK(cudaMemsetAsync(counter, 0, 4));
                k_rowimage<<<148, threads, (size_t)nbuf * (fpitch + dpitch)>>>(f, fpitch, d, dpitch, rows, band, counter, delay, jitter, pieces, nbuf, G);
                CK(cudaGetLastError());
            });
        }
    CK(cudaDeviceSynchronize());
    return 0;
}
