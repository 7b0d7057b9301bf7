// cluster.cuh -- cluster-resident decode kernel (Q80 path): one 16-CTA thread-block cluster runs the whole token.
//
// Why: at batch 1 a 168M..1.7B-parameter model is bound by dependency latency, not bandwidth.  In the grid-wide
// persistent kernel (kernels.cuh:k_decode_mega) every one of the 5L+1 phases costs ~5 L2 round trips (~900 cycles
// each on B200: grid barrier x2, activation fetch, weight tile, store visibility).  Here the activations never
// leave the SMs:
//   * every CTA keeps a full replica of the activation vectors (x, q/k/v of the position, attention output, SwiGLU
//     output) in its own shared memory; a CTA computes 1/16 of the rows of each matvec and writes each result into
//     all 16 replicas with distributed-shared-memory stores (st.shared::cluster), then a hardware cluster barrier
//     (~400 cycles) replaces the L2 counter barrier (~2400 cycles);
//   * weights arrive through a per-CTA shared-memory ring filled by cp.async.bulk (TMA bulk copies, mbarrier
//     complete_tx).  The per-token weight schedule is static, so loads are issued up to NST tiles ahead -- across
//     phase and layer boundaries -- and HBM latency is never exposed;
//   * the weight stream is re-laid-out at load time: per rank, per phase, tiles of T rows = [T x n int8 codes]
//     [T x n/gs fp32 scales]; rmsnorm gains ride in the same stream as header tiles.
// The arithmetic (activation quantisation, integer dots, ordered fp32 combine, attention partials, argmax) is the
// same as in kernels.cuh, so results are bit-identical to the other two paths (tests/test_gpu_engine.py).
#pragma once
#include "kernels.cuh"

namespace nb {

constexpr int kCluster = 16;
constexpr int kMaxStages = 8;

struct ClPhase {              // one matvec phase of the per-token schedule (identical for all ranks)
    uint64_t stream_off;      // byte offset of this phase's first weight tile inside a rank's stream
    uint64_t gain_off;        // byte offset of the gain vector inside the shared region (has_gain)
    uint32_t tile_base;       // index of the phase's first ring tile within the token (header tile included)
    uint32_t ntiles;          // weight tiles
    uint32_t has_gain;
    uint32_t rows_per_rank, rows_per_tile, tile_stride, n, epi, layer, pad;   // pad: source vector (0 x, 1 attention output, 2 SwiGLU output)
};

struct ClusterArgs {
    const uint8_t *stream; uint64_t rank_stride; const uint8_t *shared_base;
    const ClPhase *phases; uint32_t nphases, tiles_per_token, nstages, stage_bytes;
    // byte offsets of the regions inside dynamic shared memory
    uint32_t off_ring, off_phases, off_x, off_q, off_kraw, off_vrow, off_xba, off_hb, off_part, off_act, off_slots, off_attn;
    const void *emb_w, *emb_aux;
    float *logits, *kc, *vc;            // caches: [L][KV][max_seq][hd]
    const float *qnorm, *knorm, *rope_cos, *rope_sin;
    DevState *st; uint32_t *ids; uint8_t *seen;
    uint32_t n_steps;
    Dims d;
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) { } }
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// write one float into the same shared-memory slot of every CTA of the cluster (lanes 0..15 -> rank = lane)
__device__ __forceinline__ void scatter_f32(float *local_slot, float v) {
    const int lane = threadIdx.x & 31;
    if (lane < kCluster) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_slot)), "r"(lane));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
    }
}
__device__ __forceinline__ void scatter_u32(uint32_t *local_slot, uint32_t v) {
    const int lane = threadIdx.x & 31;
    if (lane < kCluster) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_slot)), "r"(lane));
        asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(remote), "r"(v) : "memory");
    }
}

// ---------------------------------------------------------------- weight ring
struct Ring {
    uint64_t *full, *empty;       // [nstages]
    unsigned char *buf;           // nstages * stage_bytes
    uint32_t nstages, stage_bytes;
};

// producer cursor (lives in the registers of warp 0 / lane 0)
struct Producer {
    uint64_t issued, total;       // global tile counters
    uint32_t phase, j;            // next tile to issue: phase index within the token, tile index within the phase (-1 => header)
    int32_t jj;
};

// issue tiles until the ring is NST ahead of tile `t` (called by warp 0 / lane 0 only)
__device__ __forceinline__ void ring_refill(const ClusterArgs &g, const ClPhase *ph, const Ring &r, Producer &p, uint64_t t, uint32_t rank) {
    while (p.issued < p.total && p.issued < t + r.nstages) {
        const uint32_t s = (uint32_t)(p.issued % r.nstages);
        const uint32_t use = (uint32_t)(p.issued / r.nstages);
        mbar_wait(&r.empty[s], (use & 1u) ^ 1u);                 // previous occupant fully consumed (passes at once on first use)
        const ClPhase &c = ph[p.phase];
        const uint8_t *src; uint32_t bytes;
        if (p.jj < 0) {                                          // header tile: rmsnorm gain, shared by all ranks
            src = g.shared_base + c.gain_off; bytes = (c.n * 4u + 15u) & ~15u;
        } else {
            const uint32_t rows = min(c.rows_per_tile, c.rows_per_rank - (uint32_t)p.jj * c.rows_per_tile);
            const uint32_t G = c.n / g.d.gs;
            src = g.stream + (uint64_t)rank * g.rank_stride + c.stream_off + (uint64_t)p.jj * c.tile_stride;
            bytes = (rows * (c.n + 4u * G) + 15u) & ~15u;
        }
        mbar_expect_tx(&r.full[s], bytes);
        bulk_g2s(r.buf + (size_t)s * r.stage_bytes, src, bytes, &r.full[s]);
        p.issued++;
        p.jj++;
        if (p.jj >= (int32_t)c.ntiles) {                         // next phase (wrap to the next token)
            p.phase++;
            if (p.phase >= g.nphases) p.phase = 0;
            p.jj = ph[p.phase].has_gain ? -1 : 0;
        }
    }
}

// ---------------------------------------------------------------- activation prologue from a local smem vector
// act layout as in kernels.cuh: int8 codes[n] | pad16 | float scales[n/gs]
template <int NT>
__device__ __forceinline__ void cl_prep_q80(const float *src, const float *gain, int n, int gs, unsigned char *act, float *red) {
    int8_t *codes = reinterpret_cast<int8_t *>(act);
    float *scales = reinterpret_cast<float *>(act + ((n + 15) & ~15));
    float inv = 1.0f;
    if (gain) {
        float acc = 0.0f;
        for (int i = threadIdx.x; i < n; i += NT) { const float v = src[i]; acc = fmaf(v, v, acc); }
        float ss = block_sum<NT>(acc, red);
        ss = __fdiv_rn(ss, (float)n);
        ss = __fadd_rn(ss, 1e-5f);
        inv = __fdiv_rn(1.0f, __fsqrt_rn(ss));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = n / gs, epl = gs / 32;
    for (int gi = warp; gi < G; gi += NT / 32) {
        float v[8];
        float amax = 0.0f;
        const int base = gi * gs + lane * epl;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j < epl) {
                const float t = src[base + j];
                v[j] = gain ? __fmul_rn(gain[base + j], __fmul_rn(inv, t)) : t;
                amax = fmaxf(amax, fabsf(v[j]));
            }
        }
        amax = warp_max(amax);
        const float sc = __fdiv_rn(amax, 127.0f);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j < epl) codes[base + j] = (int8_t)((sc == 0.0f) ? 0 : (int)roundf(__fdiv_rn(v[j], sc)));
        }
        if (lane == 0) scales[gi] = sc;
    }
    __syncthreads();
}

// tile rows from shared memory into the register tile used by q80_consume
template <int LPG>
__device__ __forceinline__ void q80_load_smem(Q80Tile<2> &t, const unsigned char *codes, const float *scales, uint32_t n, uint32_t k0) {
    constexpr uint32_t gs = LPG * 16;
    const int lane = threadIdx.x & 31;
    const uint32_t G = n / gs;
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const uint32_t k = k0 + s * 512 + lane * 16;
        const bool on = k < n;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            t.w[s][r] = on ? *reinterpret_cast<const int4 *>(codes + (size_t)r * n + k) : make_int4(0, 0, 0, 0);
            t.ws[s][r] = on ? scales[r * G + k / gs] : 0.0f;
        }
    }
}

// ---------------------------------------------------------------- attention partial of one (kv head, split) from smem q/k/v
// part layout per (kv head g, split): [KVM][hd + 2] floats: acc[hd], m, l
template <int KVM, int NT>
__device__ __forceinline__ void cl_attn_partial(const ClusterArgs &g, uint32_t layer, uint32_t kvh, uint32_t split, uint32_t rpk, uint32_t pos,
                                                uint32_t range, const float *q_s, const float *kraw_s, const float *vrow_s, float *part_slot,
                                                float *ws) {
    constexpr int NW = NT / 32;
    const Dims &d = g.d;
    const uint32_t hd = d.hd;
    uint32_t chunk = (range + rpk - 1) / rpk;
    chunk = (chunk + 7u) & ~7u;
    const uint32_t t0 = min(range, split * chunk), t1 = min(range, t0 + chunk), len = t1 - t0;
    const bool owner = (pos >= t0 && pos < t1);
    uint32_t lpr = 1; while (lpr * 4 < hd) lpr <<= 1;
    const uint32_t rpw = 32 / lpr;
    // workspace: qs[KVM*hd] | krow[hd] | stat[2*KVM] (+pad) | sc[KVM*cap] | red part[NW*rpw*KVM*hd]
    float *qs = ws;
    float *krow = qs + KVM * hd;
    float *stat = krow + hd;
    float *sc = stat + ((2 * KVM + 3) & ~3);
    const uint32_t cap = (chunk + 7u) & ~7u;
    float *part = sc + KVM * cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t sub = lane / lpr, li = lane % lpr;
    const float *cr = g.rope_cos + (size_t)pos * (hd / 2), *ci = g.rope_sin + (size_t)pos * (hd / 2);
    const float *qn = g.qnorm ? g.qnorm + (size_t)layer * hd : nullptr, *kn = g.knorm ? g.knorm + (size_t)layer * hd : nullptr;

    for (uint32_t i = threadIdx.x; i < KVM * hd; i += NT) qs[i] = q_s[(size_t)kvh * KVM * hd + i];
    if (owner) for (uint32_t i = threadIdx.x; i < hd; i += NT) krow[i] = kraw_s[(size_t)kvh * hd + i];
    __syncthreads();
    for (uint32_t m = warp; m < KVM + (owner ? 1u : 0u); m += NW) {
        if (m < KVM) head_norm_rope(qs + m * hd, qn, cr, ci, d, false);
        else head_norm_rope(krow, kn, cr, ci, d, false);
    }
    __syncthreads();
    const size_t kvl = (size_t)d.KV * d.max_seq * hd;
    float *kbase = g.kc + layer * kvl + (size_t)kvh * d.max_seq * hd, *vbase = g.vc + layer * kvl + (size_t)kvh * d.max_seq * hd;
    if (owner) for (uint32_t i = threadIdx.x; i < hd; i += NT) kbase[(size_t)pos * hd + i] = krow[i];
    const float dv = sqrtf((float)hd);
    const uint32_t col = li * 4;
    const bool colon = col < hd;
    float4 qv[KVM];
#pragma unroll
    for (int m = 0; m < KVM; m++) qv[m] = colon ? *reinterpret_cast<const float4 *>(qs + m * hd + col) : make_float4(0, 0, 0, 0);
    for (uint32_t tb = warp * rpw; tb < len; tb += NW * rpw) {
        const uint32_t tl = tb + sub;
        float4 kv = make_float4(0, 0, 0, 0);
        if (tl < len && colon) {
            const uint32_t t = t0 + tl;
            kv = (t == pos) ? *reinterpret_cast<const float4 *>(krow + col) : __ldcg(reinterpret_cast<const float4 *>(kbase + (size_t)t * hd + col));
        }
#pragma unroll
        for (int m = 0; m < KVM; m++) {
            float acc = kv.x * qv[m].x;
            acc = fmaf(kv.y, qv[m].y, acc); acc = fmaf(kv.z, qv[m].z, acc); acc = fmaf(kv.w, qv[m].w, acc);
            for (uint32_t o = lpr >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (tl < len && li == 0) sc[m * cap + tl] = __fdiv_rn(acc, dv);
        }
    }
    __syncthreads();
    if (warp < KVM) {
        float *s = sc + warp * cap;
        float mx = -FLT_MAX;
        for (uint32_t t = lane; t < len; t += 32) mx = fmaxf(mx, s[t]);
        mx = warp_max(mx);
        float ls = 0.0f;
        for (uint32_t t = lane; t < len; t += 32) { const float e = expf(s[t] - mx); s[t] = e; ls += e; }
        ls = warp_sum(ls);
        if (lane == 0) { stat[2 * warp] = mx; stat[2 * warp + 1] = ls; }
    }
    __syncthreads();
    float4 av[KVM];
#pragma unroll
    for (int m = 0; m < KVM; m++) av[m] = make_float4(0, 0, 0, 0);
    for (uint32_t tb = warp * rpw; tb < len; tb += NW * rpw) {
        const uint32_t tl = tb + sub;
        if (tl < len && colon) {
            const uint32_t t = t0 + tl;
            const float4 vv = (t == pos) ? *reinterpret_cast<const float4 *>(vrow_s + (size_t)kvh * hd + col)
                                         : __ldcg(reinterpret_cast<const float4 *>(vbase + (size_t)t * hd + col));
#pragma unroll
            for (int m = 0; m < KVM; m++) {
                const float e = sc[m * cap + tl];
                av[m].x = fmaf(e, vv.x, av[m].x); av[m].y = fmaf(e, vv.y, av[m].y);
                av[m].z = fmaf(e, vv.z, av[m].z); av[m].w = fmaf(e, vv.w, av[m].w);
            }
        }
    }
    if (colon) {
#pragma unroll
        for (int m = 0; m < KVM; m++)
            *reinterpret_cast<float4 *>(part + ((size_t)(warp * rpw + sub) * KVM + m) * hd + col) = av[m];
    }
    __syncthreads();
    const uint32_t np = NW * rpw;
    // reduce across warps and publish the partial to every CTA of the cluster; one warp handles one (m, i) element batch
    for (uint32_t idx = warp; idx < KVM * (hd + 2); idx += NW) {
        const uint32_t m = idx / (hd + 2), i = idx % (hd + 2);
        float v;
        if (i < hd) { v = 0.0f; for (uint32_t p = 0; p < np; p++) v += part[((size_t)p * KVM + m) * hd + i]; }
        else v = (len == 0) ? (i == hd ? -FLT_MAX : 0.0f) : stat[2 * m + (i - hd)];
        scatter_f32(part_slot + (size_t)m * (hd + 2) + i, v);
    }
}

// ---------------------------------------------------------------- the kernel
template <int LPG, int KVM>
__global__ void __launch_bounds__(kThreads, 1) k_decode_cluster(const ClusterArgs g) {
    extern __shared__ __align__(128) unsigned char csm[];
    unsigned char *sm = csm;
    __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
    __shared__ MatvecSmem ms;
    constexpr uint32_t gs = LPG * 16;
    const Dims &d = g.d;
    const uint32_t rank = cluster_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    Ring ring{full_bar, empty_bar, sm + g.off_ring, g.nstages, g.stage_bytes};
    ClPhase *ph = reinterpret_cast<ClPhase *>(sm + g.off_phases);
    float *x_s = reinterpret_cast<float *>(sm + g.off_x), *q_s = reinterpret_cast<float *>(sm + g.off_q);
    float *kraw_s = reinterpret_cast<float *>(sm + g.off_kraw), *vrow_s = reinterpret_cast<float *>(sm + g.off_vrow);
    float *xba_s = reinterpret_cast<float *>(sm + g.off_xba), *hb_s = reinterpret_cast<float *>(sm + g.off_hb);
    float *part_s = reinterpret_cast<float *>(sm + g.off_part);
    unsigned char *act = sm + g.off_act;
    float *slot_v = reinterpret_cast<float *>(sm + g.off_slots); uint32_t *slot_i = reinterpret_cast<uint32_t *>(slot_v + kCluster);
    float *attn_ws = reinterpret_cast<float *>(sm + g.off_attn);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < g.nstages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint64_t *src = reinterpret_cast<const uint64_t *>(g.phases);
        uint64_t *dst = reinterpret_cast<uint64_t *>(ph);
        for (uint32_t i = threadIdx.x; i < g.nphases * (uint32_t)(sizeof(ClPhase) / 8); i += kThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    cluster_sync_all();

    // step state (identical in every CTA; advanced locally)
    uint32_t pos = __ldcg(&g.st->pos);
    const uint32_t causal = __ldcg(&g.st->is_causal), n_prompt = __ldcg(&g.st->n_prompt), advance = __ldcg(&g.st->advance);
    const float pen = __ldcg(&g.st->penalty);
    uint32_t tok = __ldcg(&g.st->use_token) ? __ldcg(&g.st->token) : __ldcg(g.ids + pos);

    Producer prod{0, (uint64_t)g.n_steps * g.tiles_per_token, 0, 0, ph[0].has_gain ? -1 : 0};
    const uint32_t rpk = kCluster / d.KV;                 // ranks per kv head
    const uint32_t part_stride = KVM * (d.hd + 2);

    for (uint32_t step = 0; step < g.n_steps; step++) {
        const uint64_t tok_tile0 = (uint64_t)step * g.tiles_per_token;
        if (warp == 0 && lane == 0) ring_refill(g, ph, ring, prod, tok_tile0, rank);      // start streaming before the embedding fetch
        embed_row<kThreads>(g.emb_w, g.emb_aux, x_s, tok, d);
        __syncthreads();
        const uint32_t range = causal ? pos + 1 : d.max_seq;
        float bestv = -FLT_MAX; uint32_t besti = 0xffffffffu;

        for (uint32_t p = 0; p < g.nphases; p++) {
            const ClPhase &c = ph[p];
            const uint32_t G = c.n / gs;
            const float *src = (c.pad == 1u) ? xba_s : (c.pad == 2u) ? hb_s : x_s;       // pad = activation source selector
            uint64_t t = tok_tile0 + c.tile_base;
            // ---- header tile (gain) + activation prologue ----
            const float *gain = nullptr;
            uint32_t hs = 0;
            if (c.has_gain) {
                hs = (uint32_t)(t % ring.nstages);
                if (warp == 0 && lane == 0) ring_refill(g, ph, ring, prod, t, rank);
                mbar_wait(&ring.full[hs], (uint32_t)(t / ring.nstages) & 1u);
                gain = reinterpret_cast<const float *>(ring.buf + (size_t)hs * ring.stage_bytes);
            }
            cl_prep_q80<kThreads>(src, gain, (int)c.n, (int)gs, act, ms.red);
            if (c.has_gain) { if (lane == 0) mbar_arrive(&ring.empty[hs]); t++; }
            // ---- weight tiles ----
            const uint32_t ppt = c.rows_per_tile / 2;                 // row pairs per full tile
            for (uint32_t j = 0; j < c.ntiles; j++, t++) {
                const uint32_t s = (uint32_t)(t % ring.nstages);
                if (warp == 0 && lane == 0) ring_refill(g, ph, ring, prod, t, rank);
                mbar_wait(&ring.full[s], (uint32_t)(t / ring.nstages) & 1u);
                const uint32_t rows = min(c.rows_per_tile, c.rows_per_rank - j * c.rows_per_tile);
                const unsigned char *tile = ring.buf + (size_t)s * ring.stage_bytes;
                const float *tscales = reinterpret_cast<const float *>(tile + (size_t)rows * c.n);
                const uint32_t first = (uint32_t)((warp + kWarps - (j * ppt) % kWarps) % kWarps);
                for (uint32_t i = first; i < rows / 2; i += kWarps) {
                    float val[2] = {0.0f, 0.0f};
                    for (uint32_t k0 = 0; k0 < c.n; k0 += 1024) {
                        Q80Tile<2> tl;
                        q80_load_smem<LPG>(tl, tile + (size_t)(2 * i) * c.n, tscales + (size_t)(2 * i) * G, c.n, k0);
                        q80_consume<2, LPG>(tl, c.n, k0, act, val);
                    }
                    const uint32_t lrow = j * c.rows_per_tile + 2 * i;           // row within this rank's slice
                    const uint32_t row = rank * c.rows_per_rank + lrow;          // row of the fused matrix
                    if (c.epi == EPI_SWIGLU) {
                        const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-val[0])));
                        scatter_f32(hb_s + (row >> 1), __fmul_rn(__fmul_rn(val[0], sg), val[1]));
                    } else {
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            const uint32_t rr = row + r;
                            float v = val[r];
                            if (c.epi == EPI_RESID) scatter_f32(x_s + rr, __fadd_rn(x_s[rr], v));
                            else if (c.epi == EPI_QKV) {
                                if (rr < d.q_dim) scatter_f32(q_s + rr, v);
                                else if (rr < d.q_dim + d.kv_dim) scatter_f32(kraw_s + (rr - d.q_dim), v);
                                else {
                                    const uint32_t cc = rr - d.q_dim - d.kv_dim, h = cc / d.hd, e = cc % d.hd;
                                    scatter_f32(vrow_s + cc, v);
                                    if (lane == 0) g.vc[(size_t)c.layer * d.KV * d.max_seq * d.hd + ((size_t)h * d.max_seq + pos) * d.hd + e] = v;
                                }
                            } else {        // EPI_CLS: infer.c:1156-1167 penalty, then first-max argmax :1026-1037
                                if (pen != 1.0f && __ldcg(g.seen + rr)) v = __fdiv_rn(v, pen);
                                if (lane == 0) g.logits[rr] = v;
                                if (v > bestv) { bestv = v; besti = rr; }
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ring.empty[s]);
            }
            // ---- phase boundary ----
            if (c.epi == EPI_CLS) {
                if (lane == 0) { ms.best_v[warp] = bestv; ms.best_i[warp] = besti; }
                __syncthreads();
                if (warp == 0) {
                    float bv = -FLT_MAX; uint32_t bi = 0xffffffffu;
                    if (lane < kWarps) { bv = ms.best_v[lane]; bi = ms.best_i[lane]; }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (oi != 0xffffffffu && (ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
                    }
                    scatter_f32(slot_v + rank, bv);
                    scatter_u32(slot_i + rank, bi);
                }
            }
            cluster_sync_all();
            if (c.epi == EPI_QKV) {
                // ---- attention: this rank's (kv head, split) partial -> all replicas; then every CTA merges all heads ----
                const uint32_t kvh = rank / rpk, split = rank % rpk;
                cl_attn_partial<KVM, kThreads>(g, c.layer, kvh, split, rpk, pos, range, q_s, kraw_s, vrow_s,
                                               part_s + (size_t)(kvh * rpk + split) * part_stride, attn_ws);
                cluster_sync_all();
                for (uint32_t idx = threadIdx.x; idx < d.H * d.hd; idx += kThreads) {
                    const uint32_t h = idx / d.hd, i = idx % d.hd, kh = h / KVM, m = h % KVM;
                    float M = -FLT_MAX;
                    for (uint32_t sp = 0; sp < rpk; sp++) M = fmaxf(M, part_s[(size_t)(kh * rpk + sp) * part_stride + m * (d.hd + 2) + d.hd]);
                    float L = 0.0f, o = 0.0f;
                    for (uint32_t sp = 0; sp < rpk; sp++) {
                        const float *pp = part_s + (size_t)(kh * rpk + sp) * part_stride + m * (d.hd + 2);
                        const float w = expf(pp[d.hd] - M);
                        L += pp[d.hd + 1] * w;
                        o = fmaf(pp[i], w, o);
                    }
                    xba_s[idx] = __fdiv_rn(o, L);
                }
                __syncthreads();
            }
        }
        // ---- every CTA knows all 16 partial argmaxes: pick the token, advance the state ----
        float bv = -FLT_MAX; uint32_t bi = 0xffffffffu;
        for (int c2 = 0; c2 < kCluster; c2++) {
            const float v = slot_v[c2]; const uint32_t i = slot_i[c2];
            if (i != 0xffffffffu && (v > bv || (v == bv && i < bi))) { bv = v; bi = i; }
        }
        if (bi == 0xffffffffu) bi = 0;
        uint32_t nxt = bi;
        if (advance) {
            const bool forced = (pos + 1 < n_prompt);
            if (forced) nxt = __ldcg(g.ids + pos + 1);
            if (rank == 0 && threadIdx.x == 0) {
                g.seen[tok] = 1;
                if (!forced) g.ids[pos + 1] = bi;
                g.st->pos = pos + 1;
                g.st->next_token = nxt;
            }
            tok = nxt; pos = pos + 1;
        } else if (rank == 0 && threadIdx.x == 0) {
            g.st->next_token = bi;
        }
        cluster_sync_all();          // slots / seen[] are reused by the next token
    }
}

}  // namespace nb
