// Why does the emission kernel sit at ~5.0 TB/s when the bare store pattern reaches 6.8 TB/s?
// Experiments on the K2 store pattern (tile = 64 rows x 2 KB founder + 512 B dose per block):
//   pace   the same stores, but every row is preceded by `delay` dependent ALU operations, so a warp
//          emits rows at a slow steady pace (like the real kernel) instead of in a burst
//   bulk   rows are first written to shared memory and leave as cp.async.bulk (TMA, 1-D) stores of
//          2 KB / 512 B, `rows_per_flush` rows at a time
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_pace write_pace.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t spin(uint32_t x, int n) {
    for (int i = 0; i < n; i++) x = x * 1664525u + 1013904223u;   // dependent IMAD chain, ~4 cycles each
    return x;
}

__global__ void __launch_bounds__(128) k_pace(uint8_t* f, size_t fpitch, uint8_t* d, size_t dpitch, int rows_total, int tile_rows,
                                              int n_col_tiles, int n_tiles, int* counter, int delay, int sync_rows) {
    __shared__ int s_tile;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1);
        __syncthreads();
        const int tile = s_tile;
        __syncthreads();
        if (tile >= n_tiles) break;
        const int lt = tile / n_col_tiles, ct = tile % n_col_tiles;
        const int r0 = lt * tile_rows, r1 = min(r0 + tile_rows, rows_total);
        const size_t col = ((size_t)ct * blockDim.x + threadIdx.x) * 16;
        uint4 v = make_uint4(tile, ct, lt, threadIdx.x);
        uint8_t* pf = f + (size_t)r0 * fpitch + col;
        uint8_t* pd = d + (size_t)r0 * dpitch + col / 4;
        const bool ok = col + 16 <= fpitch;
        // per-warp skew like the real kernel's run starts
        const int my_delay = delay + ((threadIdx.x >> 5) * 7 + tile) % 5 * (delay / 8);
        for (int r = r0; r < r1; r++) {
            v.x = spin(v.x, my_delay);
            if (ok) {
                *reinterpret_cast<uint4*>(pf) = v;
                *reinterpret_cast<uint32_t*>(pd) = v.x;
            }
            pf += fpitch;
            pd += dpitch;
            if (sync_rows && ((r - r0) % sync_rows) == sync_rows - 1) __syncthreads();
        }
    }
}

// rows staged in shared memory, flushed with 1-D bulk stores (one elected thread issues them)
template <int FLUSH>
__global__ void __launch_bounds__(128) k_bulk(uint8_t* f, size_t fpitch, uint8_t* d, size_t dpitch, int rows_total, int tile_rows,
                                              int n_col_tiles, int n_tiles, int* counter, int delay) {
    __shared__ int s_tile;
    __shared__ __align__(128) uint4 s_f[2][FLUSH][128];
    __shared__ __align__(128) uint32_t s_d[2][FLUSH][128];
    int buf = 0;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1);
        __syncthreads();
        const int tile = s_tile;
        __syncthreads();
        if (tile >= n_tiles) break;
        const int lt = tile / n_col_tiles, ct = tile % n_col_tiles;
        const int r0 = lt * tile_rows, r1 = min(r0 + tile_rows, rows_total);
        const size_t col0 = (size_t)ct * 128 * 16;
        const int width = (int)min((size_t)2048, fpitch - col0);   // bytes of this tile's rows (multiple of 16)
        uint4 v = make_uint4(tile, ct, lt, threadIdx.x);
        const int my_delay = delay + ((threadIdx.x >> 5) * 7 + tile) % 5 * (delay / 8);
        for (int rb = r0; rb < r1; rb += FLUSH) {
            const int nr = min(FLUSH, r1 - rb);
            // the buffer we are about to fill was handed to the bulk engine two flushes ago
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncthreads();
            for (int k = 0; k < nr; k++) {
                v.x = spin(v.x, my_delay);
                s_f[buf][k][threadIdx.x] = v;
                s_d[buf][k][threadIdx.x] = v.x;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int k = 0; k < nr; k++) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(f + (size_t)(rb + k) * fpitch + col0),
                                 "r"((uint32_t)__cvta_generic_to_shared(&s_f[buf][k][0])), "r"(width) : "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + (size_t)(rb + k) * dpitch + col0 / 4),
                                 "r"((uint32_t)__cvta_generic_to_shared(&s_d[buf][k][0])), "r"(width / 4) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            buf ^= 1;
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const int rows = 24000;
    const size_t fpitch = 40064, dpitch = 10112;
    uint8_t *f, *d;
    int* counter;
    CK(cudaMalloc(&f, rows * fpitch));
    CK(cudaMalloc(&d, rows * dpitch));
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
        printf("%-56s best %7.1f us %7.1f GB/s   mean %7.1f us\n", name, best * 1e3, bytes / best / 1e6, sum / 8 * 1e3);
    };
    const int tile_rows = 64, nct = (int)((fpitch / 16 + 127) / 128), nt = (rows + tile_rows - 1) / tile_rows * nct;
    for (int bps : {2, 4})
        for (int delay : {0, 8, 32, 64, 128})
            for (int sync_rows : {0, 8}) {
                char nm[96];
                snprintf(nm, 96, "pace bps %d delay %3d sync %d", bps, delay, sync_rows);
                time(nm, [&] {
                    CK(cudaMemsetAsync(counter, 0, 4));
                    k_pace<<<std::min(nt, 148 * bps), 128>>>(f, fpitch, d, dpitch, rows, tile_rows, nct, nt, counter, delay, sync_rows);
                });
            }
    for (int bps : {2, 4})
        for (int delay : {0, 32, 64, 128}) {
            char nm[96];
            snprintf(nm, 96, "bulk flush 8  bps %d delay %3d", bps, delay);
            time(nm, [&] {
                CK(cudaMemsetAsync(counter, 0, 4));
                k_bulk<8><<<std::min(nt, 148 * bps), 128>>>(f, fpitch, d, dpitch, rows, tile_rows, nct, nt, counter, delay);
            });
            snprintf(nm, 96, "bulk flush 4  bps %d delay %3d", bps, delay);
            time(nm, [&] {
                CK(cudaMemsetAsync(counter, 0, 4));
                k_bulk<4><<<std::min(nt, 148 * bps), 128>>>(f, fpitch, d, dpitch, rows, tile_rows, nct, nt, counter, delay);
            });
        }
    CK(cudaDeviceSynchronize());
    return 0;
}
