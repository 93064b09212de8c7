// Write-only HBM ceiling for the emission kernel's store pattern on this B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_ceiling write_ceiling.cu
// Variants:
//   memset      cudaMemsetAsync of the whole buffer
//   linear      grid-stride uint4 stores, each warp 512 contiguous bytes
//   tiles       the K2 pattern: a block of T threads owns T*16 contiguous bytes of a row and walks
//               `rows` consecutive rows (row pitch = table width), tiles handed out by an atomic
//   tiles+dose  same plus a second table with a quarter of the width (4-byte stores)
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_linear(uint4* p, size_t n) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void k_tiles(uint8_t* f, size_t fpitch, uint8_t* d, size_t dpitch, int rows_total, int tile_rows, int n_col_tiles,
                        int n_tiles, int* counter) {
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
        uint8_t* pd = d ? d + (size_t)r0 * dpitch + col / 4 : nullptr;
        if (col + 16 <= fpitch) {
#pragma unroll 4
            for (int r = r0; r < r1; r++) {
                *reinterpret_cast<uint4*>(pf) = v;
                pf += fpitch;
                if (pd) { *reinterpret_cast<uint32_t*>(pd) = v.x; pd += dpitch; }
                v.x++;
            }
        }
    }
}

// per-warp units (tile, warp quarter): ticket per warp; optional plan reads (48 B per lane per unit)
// mode bit0: dynamic ticket (else round-robin); bit1: read a 48-byte plan record per lane and unit
// through cp.async one unit ahead; bit2: plain (non-async) plan loads at unit start
__global__ void __launch_bounds__(128) k_warp_units(uint8_t* f, size_t fpitch, int rows_total, int tile_rows, int n_col_tiles,
                                                    int n_tiles, int* counter, const uint4* plan, int mode) {
    __shared__ uint4 s_plan[2][3][128];
    const int lane = threadIdx.x & 31, tid = threadIdx.x;
    const int n_units = n_tiles * 4;
    const int stride = gridDim.x * 4;
    auto fetch = [&](int prev) -> int {
        if (mode & 1) {
            int u = 0;
            if (lane == 0) u = atomicAdd(counter, 1);
            return __shfl_sync(0xFFFFFFFFu, u, 0);
        }
        return prev + stride;
    };
    auto prefetch = [&](int u, int buf) {
        const uint4* src = plan + ((size_t)(u >> 2) * 128 + (u & 3) * 32 + lane) * 3;
        for (int k = 0; k < 3; k++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_plan[buf][k][tid])), "l"(src + k));
        asm volatile("cp.async.commit_group;");
    };
    if (mode & 16) {   // static block-level tiles: the block's warps walk the same tile sequence without talking
        int buf = 0;
        int tile = blockIdx.x;
        const int wq = tid >> 5;
        if ((mode & 2) && tile < n_tiles) prefetch(tile * 4 + wq, 0);
        for (; tile < n_tiles; tile += gridDim.x) {
            if (mode & 32) __syncthreads();
            uint4 v = make_uint4(tile, lane, 0, 0);
            if (mode & 2) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                v = s_plan[buf][0][tid];
                v.y += s_plan[buf][1][tid].x + s_plan[buf][2][tid].y;
                buf ^= 1;
                if (tile + (int)gridDim.x < n_tiles) prefetch((tile + gridDim.x) * 4 + wq, buf);
            }
            const int lt = tile / n_col_tiles, ct = tile % n_col_tiles;
            const int r0 = lt * tile_rows, r1 = min(r0 + tile_rows, rows_total);
            const size_t col = ((size_t)ct * 128 + tid) * 16;
            uint8_t* pf = f + (size_t)r0 * fpitch + col;
            if (col + 16 <= fpitch) {
#pragma unroll 4
                for (int r = r0; r < r1; r++) { *reinterpret_cast<uint4*>(pf) = v; pf += fpitch; }
            }
        }
        return;
    }
    int u = (mode & 1) ? fetch(0) : blockIdx.x * 4 + (tid >> 5);
    int buf = 0;
    if ((mode & 2) && u < n_units) prefetch(u, 0);
    while (u < n_units) {
        uint4 v = make_uint4(u, lane, 0, 0);
        if (mode & 2) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            v = s_plan[buf][0][tid];
            v.y += s_plan[buf][1][tid].x + s_plan[buf][2][tid].y;
        }
        if (mode & 4) {
            const uint4* src = plan + ((size_t)(u >> 2) * 128 + (u & 3) * 32 + lane) * 3;
            v = src[0];
            v.y += src[1].x + src[2].y;
        }
        const int un = fetch(u);
        buf ^= 1;
        if ((mode & 2) && un < n_units) prefetch(un, buf);
        const int tile = u >> 2;
        const int lt = tile / n_col_tiles, ct = tile % n_col_tiles;
        const int r0 = lt * tile_rows, r1 = min(r0 + tile_rows, rows_total);
        const size_t col = ((size_t)ct * 128 + (u & 3) * 32 + lane) * 16;
        uint8_t* pf = f + (size_t)r0 * fpitch + col;
        if (col + 16 <= fpitch) {
#pragma unroll 4
            if (mode & 8) {
#pragma unroll 4
                for (int r = r0; r < r1; r++) {
                    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(pf), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    pf += fpitch;
                }
            } else {
#pragma unroll 4
                for (int r = r0; r < r1; r++) {
                    *reinterpret_cast<uint4*>(pf) = v;
                    pf += fpitch;
                }
            }
        }
        u = un;
    }
}

// Block-level dynamic tiles (like k_tiles) that also read a 6 KB plan block per tile, one tile ahead:
// how = 0 none, 1 = per-lane cp.async (3 x 16 B per lane, three separate arrays), 2 = one TMA bulk copy
__global__ void __launch_bounds__(128) k_tiles_plan(uint8_t* f, size_t fpitch, int rows_total, int tile_rows, int n_col_tiles,
                                                    int n_tiles, int* counter, const uint4* plan, int how) {
    __shared__ __align__(128) uint4 s_plan[2][384];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ int s_tile[2];
    const int tid = threadIdx.x;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
        asm volatile("fence.mbarrier_init.release.cluster;");
        s_tile[0] = atomicAdd(counter, 1);
    }
    __syncthreads();
    auto prefetch = [&](int tile, int buf) {
        if (how == 1) {
            for (int k = 0; k < 3; k++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_plan[buf][k * 128 + tid])),
                             "l"(plan + ((size_t)k * n_tiles + tile) * 128 + tid));
            asm volatile("cp.async.commit_group;");
        } else if (how == 2 && tid == 0) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_plan[buf][0]);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * buf), "r"(6144));
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(plan + (size_t)tile * 384), "r"(6144), "r"(bar0 + 8 * buf) : "memory");
        }
    };
    int tile = s_tile[0];
    int buf = 0, phase[2] = {0, 0};
    if (tile < n_tiles) prefetch(tile, 0);
    int par = 1;
    while (tile < n_tiles) {
        if (tid == 0) s_tile[par] = atomicAdd(counter, 1);
        uint4 v = make_uint4(tile, tid, 0, 0);
        if (how == 1) { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
        if (how == 2) {
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar0 + 8 * buf), "r"(phase[buf]));
            }
            phase[buf] ^= 1;
        }
        __syncthreads();
        const int next = s_tile[par];
        par ^= 1;
        if (how) { v = s_plan[buf][tid]; v.y += s_plan[buf][128 + tid].x + s_plan[buf][256 + tid].y; }
        buf ^= 1;
        if (next < n_tiles) prefetch(next, buf);
        const int lt = tile / n_col_tiles, ct = tile % n_col_tiles;
        const int r0 = lt * tile_rows, r1 = min(r0 + tile_rows, rows_total);
        const size_t col = ((size_t)ct * 128 + tid) * 16;
        uint8_t* pf = f + (size_t)r0 * fpitch + col;
        if (col + 16 <= fpitch) {
#pragma unroll 4
            for (int r = r0; r < r1; r++) { *reinterpret_cast<uint4*>(pf) = v; pf += fpitch; }
        }
        tile = next;
    }
}

int main(int argc, char** argv) {
    const int rows = 24000;
    const size_t fpitch = 40064, dpitch = 10112;  // configs[1]: 10002 x 4 columns, padded to 128
    uint8_t *f, *d;
    int* counter;
    CK(cudaMalloc(&f, rows * fpitch));
    CK(cudaMalloc(&d, rows * dpitch));
    CK(cudaMalloc(&counter, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto time = [&](const char* name, double bytes, auto fn) {
        for (int i = 0; i < 3; i++) fn();
        float best = 1e9;
        for (int i = 0; i < 10; i++) {
            CK(cudaEventRecord(e0));
            fn();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf("%-44s %8.1f us  %7.1f GB/s\n", name, best * 1e3, bytes / best / 1e6);
    };
    const double fb = (double)rows * fpitch, db = (double)rows * dpitch;
    time("memset founder table", fb, [&] { CK(cudaMemsetAsync(f, 1, rows * fpitch)); });
    for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32})
        for (int t : {128, 256, 512}) {
            char nm[64];
            snprintf(nm, 64, "linear grid %d x %d", g, t);
            time(nm, fb, [&] { k_linear<<<g, t>>>((uint4*)f, rows * fpitch / 16); });
        }
    for (int threads : {64, 128, 256})
        for (int tile_rows : {32, 64, 128, 256})
            for (int bps : {4, 8, 16}) {
                const int nct = (int)((fpitch / 16 + threads - 1) / threads);
                const int nt = (rows + tile_rows - 1) / tile_rows * nct;
                for (int with_dose = 0; with_dose < 2; with_dose++) {
                    char nm[96];
                    snprintf(nm, 96, "tiles thr %d rows %d bps %d %s", threads, tile_rows, bps, with_dose ? "+dose" : "");
                    time(nm, fb + (with_dose ? db : 0), [&] {
                        CK(cudaMemsetAsync(counter, 0, 4));
                        k_tiles<<<min(nt, 148 * bps), threads>>>(f, fpitch, with_dose ? d : nullptr, dpitch, rows, tile_rows, nct, nt, counter);
                    });
                }
            }
    uint4* plan;
    const int tile_rows = 64;
    const int nct = (int)((fpitch / 16 + 127) / 128);
    const int nt = (rows + tile_rows - 1) / tile_rows * nct;
    CK(cudaMalloc(&plan, (size_t)nt * 128 * 48));
    CK(cudaMemset(plan, 1, (size_t)nt * 128 * 48));
    for (int bps : {3, 4, 6})
        for (int how : {0, 1, 2}) {
            char nm[96];
            snprintf(nm, 96, "block tiles + plan reads bps %d how %d (1=cp.async 2=TMA bulk 6KB)", bps, how);
            time(nm, fb, [&] {
                CK(cudaMemsetAsync(counter, 0, 4));
                k_tiles_plan<<<min(nt, 148 * bps), 128>>>(f, fpitch, rows, tile_rows, nct, nt, counter, plan, how);
            });
        }
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    for (int persist : {0}) {
    if (persist) {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, 0));
        printf("persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB, L2 %d MB\n", prop.persistingL2CacheMaxSize >> 20,
               prop.accessPolicyMaxWindowSize >> 20, prop.l2CacheSize >> 20);
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(prop.persistingL2CacheMaxSize, (size_t)nt * 128 * 48)));
        cudaStreamAttrValue av{};
        av.accessPolicyWindow.base_ptr = plan;
        av.accessPolicyWindow.num_bytes = std::min<size_t>(prop.accessPolicyMaxWindowSize, (size_t)nt * 128 * 48);
        av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
    }
    for (int bps : {4})
        for (int mode : {16}) {
            char nm[96];
            snprintf(nm, 96, "warp units persist %d bps %d mode %d (1=ticket 2=cp.async plan 8=st.cs)", persist, bps, mode);
            time(nm, fb, [&] {
                CK(cudaMemsetAsync(counter, 0, 4, st));
                k_warp_units<<<min(nt, 148 * bps), 128, 0, st>>>(f, fpitch, rows, tile_rows, nct, nt, counter, plan, mode);
                CK(cudaStreamSynchronize(st));
            });
        }
    }
    return 0;
}
