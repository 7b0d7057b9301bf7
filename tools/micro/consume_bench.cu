// Microbenchmark: throughput of the warp-owned-tile consumer of k_decode_stream (15 warps, full-warp K-split Q80 dots) fed
// by the cp.async.bulk ring, streaming a large buffer from HBM.  Tile = T rows x (n + 16) codes + T x (G|1) scales.
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
template <int LPG, int RB>
__device__ __forceinline__ void rows_q80_warp(const unsigned char *wrow, uint32_t row_stride, const unsigned char *srow, uint32_t aux_stride,
                                              uint32_t n, const unsigned char *act, float (&val)[RB]) {
    constexpr uint32_t gs = LPG * 16;
    constexpr int GPS = 32 / LPG;
    const int lane = threadIdx.x & 31;
    const float *xs = reinterpret_cast<const float *>(act + ((n + 15u) & ~15u));
#pragma unroll
    for (int r2 = 0; r2 < RB; r2++) val[r2] = 0.0f;
    for (uint32_t k0 = 0; k0 < n; k0 += 512u) {
        const uint32_t k = k0 + lane * 16u;
        const bool on = k < n;
        const uint32_t kc = on ? k : 0u;
        const int4 xq = on ? *reinterpret_cast<const int4 *>(act + kc) : make_int4(0, 0, 0, 0);
        const float xsc = xs[kc / gs];
        float term[RB];
#pragma unroll
        for (int r2 = 0; r2 < RB; r2++) {
            const int4 w = *reinterpret_cast<const int4 *>(wrow + (size_t)r2 * row_stride + kc);
            const float ws = reinterpret_cast<const float *>(srow + (size_t)r2 * aux_stride)[kc / gs];
            int isum = __dp4a(w.x, xq.x, 0);
            isum = __dp4a(w.y, xq.y, isum); isum = __dp4a(w.z, xq.z, isum); isum = __dp4a(w.w, xq.w, isum);
#pragma unroll
            for (int o = 1; o < LPG; o <<= 1) isum += __shfl_xor_sync(0xffffffffu, isum, o);
            term[r2] = __fmul_rn(__fmul_rn((float)isum, ws), xsc);
        }
#pragma unroll
        for (int gq = 0; gq < GPS; gq++) {
#pragma unroll
            for (int r2 = 0; r2 < RB; r2++) {
                const float t = __shfl_sync(0xffffffffu, term[r2], gq * LPG);
                if (k0 + gq * gs < n) val[r2] = __fadd_rn(val[r2], t);
            }
        }
    }
}
__global__ void __launch_bounds__(512, 1) k_cons(const unsigned char *src, uint32_t ntiles, uint32_t T, uint32_t n, uint32_t nst, uint32_t stage_bytes, float *out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[32], empty[32];
    __shared__ volatile uint32_t tile_id[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t G = n / 128, row_stride = n + 16, aux_stride = (G | 1) * 4, tile_bytes = (T * (row_stride + aux_stride) + 15) & ~15u;
    unsigned char *act = sm, *ring = sm + 16384;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nst; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])) : "memory");
            tile_id[s] = 0xffffffffu;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = threadIdx.x; i < 4096; i += 512) reinterpret_cast<uint32_t *>(act)[i] = 0x01fe02ffu * (i + 1);
    __syncthreads();
    const unsigned char *base = src + (size_t)blockIdx.x * ntiles * tile_bytes;
    if (warp == 15) {
        if (lane == 0) {
            uint32_t s = 0, par = 1;
            for (uint32_t t = 0; t < ntiles; t++) {
                while (!try_wait(&empty[s], par)) { }
                tile_id[s] = t;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(tile_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(ring + (size_t)s * stage_bytes)), "l"(base + (size_t)t * tile_bytes), "r"(tile_bytes), "r"(smem_u32(&full[s])) : "memory");
                if (++s == nst) { s = 0; par ^= 1u; }
            }
        }
        return;
    }
    float acc = 0.0f;
    uint32_t s = 0, par = 0;
    for (uint32_t t = 0; t < ntiles; t++) {
        if (t % 15 == (uint32_t)warp) {
            while (tile_id[s] != t) { }
            while (!try_wait(&full[s], par)) { }
            const unsigned char *tile = ring + (size_t)s * stage_bytes, *aux = tile + (size_t)T * row_stride;
            for (uint32_t rr = 0; rr + 1 < T; rr += 2) {
                float v[2];
                rows_q80_warp<8, 2>(tile + (size_t)rr * row_stride, row_stride, aux + (size_t)rr * aux_stride, aux_stride, n, act, v);
                acc += v[0] + v[1];
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        }
        if (++s == nst) { s = 0; par ^= 1u; }
    }
    if (acc == 12345.678f) out[0] = acc;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t total = 2ull << 30;
    unsigned char *src; float *out;
    CK(cudaMalloc(&src, total)); CK(cudaMalloc(&out, 64)); CK(cudaMemset(src, 1, total));
    CK(cudaFuncSetAttribute(k_cons, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    struct Cfg { uint32_t n, T, nst; } cfgs[] = {{1024, 4, 16}, {1024, 10, 16}, {1024, 10, 8}, {2560, 4, 16}, {2560, 4, 8}, {768, 8, 16}, {768, 14, 16}, {9728, 1, 16}, {9728, 2, 8}};
    for (auto c : cfgs) {
        const uint32_t G = c.n / 128, tile_bytes = (c.T * (c.n + 16 + (G | 1) * 4) + 15) & ~15u, stage = (tile_bytes + 127) & ~127u;
        if ((size_t)stage * c.nst + 16384 > 216 * 1024) { printf("n=%u T=%u nst=%u: does not fit\n", c.n, c.T, c.nst); continue; }
        const uint32_t ntiles = (uint32_t)((total / sms) / tile_bytes);
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(a);
            k_cons<<<sms, 512, stage * c.nst + 16384>>>(src, ntiles, c.T, c.n, c.nst, stage, out);
            cudaEventRecord(b);
            CK(cudaDeviceSynchronize());
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("n=%5u rows/tile %2u (%6u B) x %2u stages: %7.1f GB/s aggregate (%5.1f per SM)\n", c.n, c.T, tile_bytes, c.nst, (double)ntiles * tile_bytes * sms / ms / 1e6, (double)ntiles * tile_bytes / ms / 1e6);
    }
    return 0;
}
