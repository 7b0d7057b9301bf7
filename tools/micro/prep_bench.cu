// Microbenchmark: intrinsic cost (SM cycles) of the activation prologue and of one Q80 row block, one CTA, L2-resident data.
#include <cstdio>
#include <vector>
#include "../../nano_b200/csrc/kernels.cuh"
using namespace nb;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) kb(const float *x, const float *gain, const int8_t *W, const float *S, float *out, int n, int rows,
                                                  long long *res, int iters) {
    extern __shared__ __align__(16) unsigned char act[];
    __shared__ float red[32];
    float *stage = reinterpret_cast<float *>(act + act_region_bytes(0x80, n, 128));
    long long t0 = clock64();
    float acc = 0.0f;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { stage_vector<kThreads>(x, gain, n, stage); acc += stage[threadIdx.x % n]; }
        if (MODE == 1) { stage_vector<kThreads>(x, gain, n, stage); acc += rms_inverse<kThreads>(stage, n, false, red); }
        if (MODE == 2) prep_q80<kThreads>(x, gain, n, 128, false, act, stage, red, nullptr, nullptr);
        if (MODE == 3) prep_q80<kThreads>(x, nullptr, n, 128, false, act, stage, red, nullptr, nullptr);
        if (MODE == 4 || MODE == 5) {
            if (it == 0) prep_q80<kThreads>(x, gain, n, 128, false, act, stage, red, nullptr, nullptr);
            const int warp = threadIdx.x >> 5;
            float val[2] = {0.0f, 0.0f};
            Q80Tile<2> t;
            q80_load<2, 8>(t, W, S, (warp * 2 + it * 32) % rows, rows, n, 0);
            q80_consume<2, 8>(t, n, 0, act, val);
            acc += val[0] + val[1];
            if (MODE == 5) __syncthreads();
        }
        if (MODE == 6) { __syncthreads(); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) res[0] = (t1 - t0) / iters;
    if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
int run(const char *name, int n, const float *x, const float *g, const int8_t *W, const float *S, float *out, long long *res) {
    int iters = 2000, rows = 4096;
    uint32_t smem = act_smem_bytes(0x80, n, 128);
    CK(cudaFuncSetAttribute((const void *)kb<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    kb<MODE><<<1, kThreads, smem>>>(x, g, W, S, out, n, rows, res, iters);
    CK(cudaDeviceSynchronize());
    long long h; CK(cudaMemcpy(&h, res, 8, cudaMemcpyDeviceToHost));
    printf("n=%5d  %-58s %6lld cycles/iter\n", n, name, h);
    return 0;
}

int main() {
    float *x, *g, *S, *out; int8_t *W; long long *res;
    CK(cudaMalloc(&x, 16384 * 4)); CK(cudaMalloc(&g, 16384 * 4)); CK(cudaMalloc(&W, 4096 * 4096)); CK(cudaMalloc(&S, 4096 * 32 * 4));
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&res, 64));
    std::vector<float> h(16384);
    for (int i = 0; i < 16384; i++) h[i] = (float)((i * 37) % 101) / 50.0f - 1.0f;
    CK(cudaMemcpy(x, h.data(), 16384 * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(g, h.data(), 16384 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(W, 1, 4096 * 4096)); CK(cudaMemset(S, 0, 4096 * 32 * 4));
    for (int n : {768, 1024, 3072}) {
        run<6>("__syncthreads only", n, x, g, W, S, out, res);
        run<0>("stage_vector (x + gain via L2, 1 sync)", n, x, g, W, S, out, res);
        run<1>("stage + rms_inverse (tree, fdiv, sqrt)", n, x, g, W, S, out, res);
        run<2>("prep_q80 with rmsnorm (stage + rms + quantise)", n, x, g, W, S, out, res);
        run<3>("prep_q80 plain (stage + quantise)", n, x, g, W, S, out, res);
        run<4>("q80 tile load (L2) + consume, RB=2, no sync", n, x, g, W, S, out, res);
        run<5>("q80 tile load (L2) + consume, RB=2, + sync", n, x, g, W, S, out, res);
    }
    return 0;
}
