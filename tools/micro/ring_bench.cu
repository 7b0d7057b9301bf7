// Microbenchmark: what can a cp.async.bulk ring deliver per SM?  One CTA per SM; lane 0 of warp 15 issues tiles of B bytes
// round-robin into NST shared-memory stages (mbarrier full/empty), warp 0 waits for each tile and releases it at once.
// Source: a 2 GB buffer streamed once (HBM), every CTA its own contiguous slice.  Reports aggregate GB/s.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void __launch_bounds__(512, 1) k_ring(const unsigned char *src, size_t per_cta, uint32_t B, uint32_t nst, uint32_t touch, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char buf[];
    __shared__ uint64_t full[32], empty[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nst; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ntiles = (uint32_t)(per_cta / B);
    const unsigned char *base = src + (size_t)blockIdx.x * per_cta;
    if (warp == 15) {
        if (lane == 0) {
            uint32_t s = 0, par = 1;
            for (uint32_t t = 0; t < ntiles; t++) {
                while (!try_wait(&empty[s], par)) { }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(B) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(buf + (size_t)s * B)), "l"(base + (size_t)t * B), "r"(B), "r"(smem_u32(&full[s])) : "memory");
                if (++s == nst) { s = 0; par ^= 1u; }
            }
        }
    } else if (warp == 0) {
        uint32_t s = 0, par = 0; unsigned acc = 0;
        for (uint32_t t = 0; t < ntiles; t++) {
            while (!try_wait(&full[s], par)) { }
            if (touch) for (uint32_t i = lane * 16u; i < B; i += 512u) { const int4 v = *reinterpret_cast<const int4 *>(buf + (size_t)s * B + i); acc += v.x ^ v.w; }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
            if (++s == nst) { s = 0; par ^= 1u; }
        }
        if (acc == 0x12345u) sink[0] = acc;
    }
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t total = 2ull << 30;
    unsigned char *src; unsigned *sink;
    CK(cudaMalloc(&src, total)); CK(cudaMalloc(&sink, 64)); CK(cudaMemset(src, 1, total));
    CK(cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (uint32_t touch : {0u, 1u})
        for (uint32_t B : {2048u, 4096u, 8192u, 12288u, 16384u, 32768u})
            for (uint32_t nst : {2u, 4u, 8u, 16u}) {
                if ((size_t)B * nst > 196 * 1024) continue;
                const size_t per_cta = (total / sms) / B * B;
                k_ring<<<sms, 512, B * nst>>>(src, per_cta, B, nst, touch, sink);     // warm
                CK(cudaDeviceSynchronize());
                cudaEventRecord(a);
                k_ring<<<sms, 512, B * nst>>>(src, per_cta, B, nst, touch, sink);
                cudaEventRecord(b);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, a, b);
                printf("touch=%u tile %5u B x %2u stages: %7.1f GB/s aggregate (%5.1f per SM)\n", touch, B, nst, per_cta * sms / ms / 1e6, per_cta / ms / 1e6);
            }
    return 0;
}
