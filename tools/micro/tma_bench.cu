// Microbenchmark: per-SM streaming bandwidth, cp.async.bulk ring vs LDG.128, for N CTAs (one per SM).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void __launch_bounds__(512, 1) k_tma(const uint8_t *src, uint64_t per_cta, uint32_t tile, uint32_t nst, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[8], empty[8];
    const uint8_t *base = src + (uint64_t)blockIdx.x * per_cta;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nst; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(16) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ntiles = (uint32_t)(per_cta / tile);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t issued = 0;
    unsigned long long acc = 0;
    for (uint32_t t = 0; t < ntiles; t++) {
        if (warp == 0 && lane == 0) {
            while (issued < ntiles && issued < t + nst) {
                const uint32_t s = issued % nst, use = issued / nst;
                while (!try_wait(&empty[s], (use & 1u) ^ 1u)) { }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(tile) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(sm + (size_t)s * tile)), "l"(base + (uint64_t)issued * tile), "r"(tile), "r"(s32(&full[s])) : "memory");
                issued++;
            }
        }
        const uint32_t s = t % nst;
        while (!try_wait(&full[s], (t / nst) & 1u)) { }
        acc += *reinterpret_cast<const unsigned long long *>(sm + (size_t)s * tile + threadIdx.x * 16);   // touch the tile
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
    }
    if (acc == 0x1234567) sink[0] = acc;
}
__global__ void __launch_bounds__(512, 1) k_ldg(const uint8_t *src, uint64_t per_cta, unsigned long long *sink) {
    const int4 *p = reinterpret_cast<const int4 *>(src + (uint64_t)blockIdx.x * per_cta);
    const uint64_t n = per_cta / 16;
    int acc = 0;
    for (uint64_t i = threadIdx.x; i + 3 * 512 < n; i += 4 * 512) {
        int4 a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p + i));
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + i + 512));
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(p + i + 1024));
        asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(p + i + 1536));
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 0x1234567) sink[0] = acc;
}
// dedicated producer (NL lanes of warp 0, lane l issues tiles l, l+NL, ...), 15 consumer warps, one consumer warp per tile
template <int NL>
__global__ void __launch_bounds__(512, 1) k_tma2(const uint8_t *src, uint64_t per_cta, uint32_t tile, uint32_t nst, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[16], empty[16];
    __shared__ volatile uint32_t tid_s[16];
    const uint8_t *base = src + (uint64_t)blockIdx.x * per_cta;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nst; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(1) : "memory");
            tid_s[s] = 0xffffffffu;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ntiles = (uint32_t)(per_cta / tile);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long acc = 0;
    if (warp == 0) {
        if (lane < NL) {
            for (uint32_t i = lane; i < ntiles; i += NL) {
                const uint32_t s = i % nst, use = i / nst;
                while (!try_wait(&empty[s], (use & 1u) ^ 1u)) { }
                tid_s[s] = i;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(tile) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(sm + (size_t)s * tile)), "l"(base + (uint64_t)i * tile), "r"(tile), "r"(s32(&full[s])) : "memory");
            }
        }
    } else {
        for (uint32_t t = warp - 1; t < ntiles; t += 15) {
            const uint32_t s = t % nst, use = t / nst;
            while (tid_s[s] != t) { }
            while (!try_wait(&full[s], use & 1u)) { }
            acc += *reinterpret_cast<const unsigned long long *>(sm + (size_t)s * tile + lane * 16);
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
    }
    if (acc == 0x1234567) sink[0] = acc;
}

int main() {
    const uint64_t total = 2ull << 30;
    uint8_t *buf; unsigned long long *sink;
    CK(cudaMalloc(&buf, total)); CK(cudaMalloc(&sink, 64)); CK(cudaMemset(buf, 1, total));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CK(cudaFuncSetAttribute((const void *)k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int ncta : {16}) {
        const uint64_t per = ((total / ncta) / (1 << 20)) * (1 << 20) > (256ull << 20) ? (256ull << 20) : ((total / ncta) / (1 << 20)) * (1 << 20);
        for (uint32_t tile : {8192u, 16384u, 32768u}) for (uint32_t nst : {2u, 4u, 6u}) {
            if ((uint64_t)tile * nst > 190 * 1024) continue;
            cudaEventRecord(a);
            k_tma<<<ncta, 512, tile * nst>>>(buf, per, tile, nst, sink);
            cudaEventRecord(b); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("TMA  ncta=%3d tile=%5u nst=%u : %7.1f GB/s total, %6.1f GB/s per SM\n", ncta, tile, nst, per * ncta / ms / 1e6, per / ms / 1e6);
        }
        CK(cudaFuncSetAttribute((const void *)k_tma2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute((const void *)k_tma2<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (uint32_t tile : {8192u, 16384u, 32768u, 65536u}) for (uint32_t nst : {2u, 4u, 8u, 12u}) {
            if ((uint64_t)tile * nst > 190 * 1024) continue;
            for (int nl : {1, 4}) {
                cudaEventRecord(a);
                if (nl == 1) k_tma2<1><<<ncta, 512, tile * nst>>>(buf, per, tile, nst, sink); else k_tma2<4><<<ncta, 512, tile * nst>>>(buf, per, tile, nst, sink);
                cudaEventRecord(b); CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, a, b);
                printf("TMA2 ncta=%3d tile=%5u nst=%2u lanes=%d : %7.1f GB/s total, %6.1f GB/s per SM\n", ncta, tile, nst, nl, per * ncta / ms / 1e6, per / ms / 1e6);
            }
        }
        cudaEventRecord(a);
        k_ldg<<<ncta, 512>>>(buf, per, sink);
        cudaEventRecord(b); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("LDG  ncta=%3d (512 thr x 4 x 16B in flight)   : %7.1f GB/s total, %6.1f GB/s per SM\n", ncta, per * ncta / ms / 1e6, per / ms / 1e6);
    }
    return 0;
}
