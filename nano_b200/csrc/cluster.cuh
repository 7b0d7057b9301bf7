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
constexpr int kMaxStages = 16;

struct ClPhase {              // one matvec phase of the per-token schedule (identical for all ranks)
    uint64_t stream_off;      // byte offset of this phase's first weight tile inside a rank's stream
    uint64_t gain_off;        // byte offset of the gain vector inside the shared region (has_gain)
    uint32_t tile_base;       // index of the phase's first ring tile within the token (header tile included)
    uint32_t ntiles;          // weight tiles
    uint32_t has_gain;
    uint32_t rows_per_rank, rows_per_tile, tile_stride, n, epi, layer, pad;   // pad: source vector (0 x, 1 attention output, 2 SwiGLU output)
    uint32_t row_stride;      // bytes between rows inside a tile: n + 16 (keeps lane-per-row 128-bit smem reads conflict-free)
    uint32_t gs_stride;       // floats between the scale rows of a tile: (n/gs) | 1 (odd => conflict-free)
    uint32_t pad2[2];
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
    unsigned long long *trace;          // optional: rank 0 / thread 0 clock64() stamps during the LAST step
    Dims d;
};
#define CL_STAMP() do { if (g.trace && rank == 0 && threadIdx.x == 0 && step + 1 == g.n_steps && ti < 1000) g.trace[ti++] = clock64(); } while (0)

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
    volatile uint32_t *tile_id;   // [nstages] global index of the tile currently assigned to the stage (written by the producer).
                                  // mbarrier parity waits are only unambiguous one phase apart; tiles of one stage are consumed
                                  // by DIFFERENT warps, and a fast warp can be several uses ahead -- it first waits for its tile
                                  // to own the stage, then for the bytes.
    uint64_t *full, *empty;       // [nstages]
    unsigned char *buf;           // nstages * stage_bytes
    uint32_t nstages, stage_bytes;
};

// consumer side: wait until tile `t` owns its stage and its bytes have landed (32-bit tile counters: a run is < 2^32 tiles)
__device__ __forceinline__ uint32_t ring_wait_tile(const Ring &r, uint32_t t) {
    const uint32_t use = t / r.nstages, s = t - use * r.nstages;
    while (r.tile_id[s] != t) { }
    mbar_wait(&r.full[s], use & 1u);
    return s;
}

// producer cursor (lives in the registers of warp 0 / lane 0); everything incremental: no divisions on the issue path
struct Producer {
    uint32_t issued, total;       // global tile counters
    uint32_t s, par;              // stage and empty-barrier parity of the next tile to issue
    uint32_t phase;               // phase (within the token) of the next tile to issue
    int32_t jj;                   // its index within the phase (-1 => the header tile carrying the rmsnorm gain)
    const uint8_t *src;           // its source address
    uint32_t rows_left;           // rows of the phase not yet issued
};

__device__ __forceinline__ void producer_enter_phase(const ClusterArgs &g, const ClPhase *ph, Producer &p, uint32_t rank) {
    const ClPhase &c = ph[p.phase];
    p.jj = c.has_gain ? -1 : 0;
    p.src = g.stream + (uint64_t)rank * g.rank_stride + c.stream_off;
    p.rows_left = c.rows_per_rank;
}

// Issue ring tiles with global index < limit (warp 0 / lane 0 only).  blocking: wait until the stage's previous
// occupant has been released; otherwise stop at the first busy stage (opportunistic prefetch of later phases).
__device__ __forceinline__ void produce(const ClusterArgs &g, const ClPhase *ph, const Ring &r, Producer &p, uint32_t limit, bool blocking, uint32_t rank) {
    if (limit > p.total) limit = p.total;
    while (p.issued < limit) {
        if (blocking) mbar_wait(&r.empty[p.s], p.par);
        else if (!mbar_try_wait(&r.empty[p.s], p.par)) return;
        const ClPhase &c = ph[p.phase];
        const uint8_t *src; uint32_t bytes;
        if (p.jj < 0) {                                          // header tile: rmsnorm gain, shared by all ranks
            src = g.shared_base + c.gain_off; bytes = (c.n * 4u + 15u) & ~15u;
        } else {
            const uint32_t rows = min(c.rows_per_tile, p.rows_left);
            src = p.src; bytes = (rows * (c.row_stride + 4u * c.gs_stride) + 15u) & ~15u;
            p.src += c.tile_stride; p.rows_left -= rows;
        }
        r.tile_id[p.s] = p.issued;
        mbar_expect_tx(&r.full[p.s], bytes);
        bulk_g2s(r.buf + p.s * r.stage_bytes, src, bytes, &r.full[p.s]);
        p.issued++;
        if (++p.s == r.nstages) { p.s = 0; p.par ^= 1u; }
        p.jj++;
        if (p.jj >= (int32_t)c.ntiles) {                         // next phase (wraps to the next token)
            if (++p.phase >= g.nphases) p.phase = 0;
            producer_enter_phase(g, ph, p, rank);
        }
    }
}

// One lane = one weight row resident in shared memory: no cross-lane reductions at all.  Integer group dots are exact;
// the fp32 combine is the reference's left-to-right sum over groups (infer.c:668-674).
template <int LPG>
__device__ __forceinline__ float cl_row_dot(const unsigned char *wrow, const float *srow, uint32_t n, const unsigned char *act) {
    constexpr uint32_t gs = LPG * 16;
    const unsigned char *codes = act;
    const float *xs = reinterpret_cast<const float *>(act + ((n + 15) & ~15));
    const uint32_t G = n / gs;
    float val = 0.0f;
    for (uint32_t gi = 0; gi < G; gi++) {
        int i0 = 0, i1 = 0;
#pragma unroll
        for (int ch = 0; ch < LPG; ch += 2) {
            const int4 w0 = *reinterpret_cast<const int4 *>(wrow + gi * gs + ch * 16), x0 = *reinterpret_cast<const int4 *>(codes + gi * gs + ch * 16);
            const int4 w1 = *reinterpret_cast<const int4 *>(wrow + gi * gs + ch * 16 + 16), x1 = *reinterpret_cast<const int4 *>(codes + gi * gs + ch * 16 + 16);
            i0 = __dp4a(w0.x, x0.x, i0); i0 = __dp4a(w0.y, x0.y, i0); i0 = __dp4a(w0.z, x0.z, i0); i0 = __dp4a(w0.w, x0.w, i0);
            i1 = __dp4a(w1.x, x1.x, i1); i1 = __dp4a(w1.y, x1.y, i1); i1 = __dp4a(w1.z, x1.z, i1); i1 = __dp4a(w1.w, x1.w, i1);
        }
        val = __fadd_rn(val, __fmul_rn(__fmul_rn((float)(i0 + i1), srow[gi]), xs[gi]));
    }
    return val;
}

// every lane writes its own value into the same slot of all 16 replicas
__device__ __forceinline__ void scatter_lane_f32(float *local_slot, float v, bool active) {
    const uint32_t a = smem_u32(local_slot);
#pragma unroll
    for (int r = 0; r < kCluster; r++) {
        if (active) {
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(a), "r"(r));
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
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
        const float rinv = __frcp_rn(sc);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j < epl) codes[base + j] = (int8_t)((sc == 0.0f) ? 0 : q80_code(v[j], sc, rinv));
        }
        if (lane == 0) scales[gi] = sc;
    }
    __syncthreads();
}

// ---------------------------------------------------------------- the kernel
template <int LPG, int KVM>
__global__ void __launch_bounds__(kThreads, 1) k_decode_cluster(const ClusterArgs g) {
    extern __shared__ __align__(128) unsigned char csm[];
    unsigned char *sm = csm;
    __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
    __shared__ volatile uint32_t stage_tile[kMaxStages];
    __shared__ MatvecSmem ms;
    constexpr uint32_t gs = LPG * 16;
    const Dims &d = g.d;
    const uint32_t rank = cluster_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    Ring ring{stage_tile, full_bar, empty_bar, sm + g.off_ring, g.nstages, g.stage_bytes};
    ClPhase *ph = reinterpret_cast<ClPhase *>(sm + g.off_phases);
    float *x_s = reinterpret_cast<float *>(sm + g.off_x), *q_s = reinterpret_cast<float *>(sm + g.off_q);
    float *kraw_s = reinterpret_cast<float *>(sm + g.off_kraw), *vrow_s = reinterpret_cast<float *>(sm + g.off_vrow);
    float *xba_s = reinterpret_cast<float *>(sm + g.off_xba), *hb_s = reinterpret_cast<float *>(sm + g.off_hb);
    float *part_s = reinterpret_cast<float *>(sm + g.off_part);
    unsigned char *act = sm + g.off_act;
    float *slot_v = reinterpret_cast<float *>(sm + g.off_slots); uint32_t *slot_i = reinterpret_cast<uint32_t *>(slot_v + kCluster);
    float *attn_ws = reinterpret_cast<float *>(sm + g.off_attn);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < g.nstages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); stage_tile[s] = 0xffffffffu; }
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

    Producer prod{0, g.n_steps * g.tiles_per_token, 0, 1u, 0, 0, nullptr, 0};      // parity 1 passes at once on a stage's first use
    producer_enter_phase(g, ph, prod, rank);
    const uint32_t rpk = kCluster / d.KV;                 // ranks per kv head
    const uint32_t part_stride = KVM * (d.hd + 2);

    uint32_t ti = 0;
    for (uint32_t step = 0; step < g.n_steps; step++) {
        const uint32_t tok_tile0 = step * g.tiles_per_token;
        CL_STAMP();
        if (warp == 0 && lane == 0) produce(g, ph, ring, prod, prod.total, false, rank);   // keep the ring full across the token boundary
        embed_row<kThreads>(g.emb_w, g.emb_aux, x_s, tok, d);
        __syncthreads();
        const uint32_t range = causal ? pos + 1 : d.max_seq;
        float bestv = -FLT_MAX; uint32_t besti = 0xffffffffu;

        for (uint32_t p = 0; p < g.nphases; p++) {
            const ClPhase &c = ph[p];
            const uint32_t G = c.n / gs;
            const float *src = (c.pad == 1u) ? xba_s : (c.pad == 2u) ? hb_s : x_s;       // pad = activation source selector
            const uint32_t t_hdr = tok_tile0 + c.tile_base;
            const uint32_t t0w = t_hdr + (c.has_gain ? 1u : 0u);          // first weight tile of the phase
            // ---- header tile (gain) + activation prologue ----
            const float *gain = nullptr;
            uint32_t hs = 0;
            if (c.has_gain) {
                if (warp == 0 && lane == 0) produce(g, ph, ring, prod, t_hdr + 1, true, rank);
                hs = ring_wait_tile(ring, t_hdr);
                gain = reinterpret_cast<const float *>(ring.buf + (size_t)hs * ring.stage_bytes);
            }
            cl_prep_q80<kThreads>(src, gain, (int)c.n, (int)gs, act, ms.red);        // ends with __syncthreads()
            if (c.has_gain && threadIdx.x == 32) mbar_arrive(&ring.empty[hs]);
            CL_STAMP();
            // ---- weight tiles: warp 0 streams, warps 1..15 each own whole tiles (one lane per row) ----
            if (warp == 0) {
                if (lane == 0) {
                    produce(g, ph, ring, prod, t0w + c.ntiles, true, rank);          // everything this phase needs
                    produce(g, ph, ring, prod, prod.total, false, rank);             // and whatever later phases fit
                }
                __syncwarp();
            } else {
                for (uint32_t j = (uint32_t)warp - 1; j < c.ntiles; j += kWarps - 1) {
                    const uint32_t s = ring_wait_tile(ring, t0w + j);
                    const uint32_t rows = min(c.rows_per_tile, c.rows_per_rank - j * c.rows_per_tile);
                    const unsigned char *tile = ring.buf + (size_t)s * ring.stage_bytes;
                    const float *tscales = reinterpret_cast<const float *>(tile + (size_t)rows * c.row_stride);
                    const bool on = (uint32_t)lane < rows;
                    float v = 0.0f;
                    if (on) v = cl_row_dot<LPG>(tile + (size_t)lane * c.row_stride, tscales + (size_t)lane * c.gs_stride, c.n, act);
                    // release the stage before the epilogue (the arrive has release semantics; DSMEM stores issued before it
                    // would have to be acknowledged by 16 SMs first)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ring.empty[s]);
                    const uint32_t row = rank * c.rows_per_rank + j * c.rows_per_tile + (uint32_t)lane;   // row of the fused matrix
                    if (c.epi == EPI_SWIGLU) {            // lanes (2i, 2i+1) = (w1 row i, w3 row i); infer.c:937-944
                        const float v3 = __shfl_down_sync(0xffffffffu, v, 1);
                        const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
                        scatter_lane_f32(hb_s + (row >> 1), __fmul_rn(__fmul_rn(v, sg), v3), on && !(lane & 1));
                    } else if (c.epi == EPI_RESID) {
                        scatter_lane_f32(x_s + row, on ? __fadd_rn(x_s[row], v) : 0.0f, on);
                    } else if (c.epi == EPI_QKV) {
                        float *dst = q_s + row;
                        if (row >= d.q_dim + d.kv_dim) {
                            const uint32_t cc = row - d.q_dim - d.kv_dim, h = cc / d.hd, e = cc % d.hd;
                            dst = vrow_s + cc;
                            if (on) g.vc[(size_t)c.layer * d.KV * d.max_seq * d.hd + ((size_t)h * d.max_seq + pos) * d.hd + e] = v;
                        } else if (row >= d.q_dim) dst = kraw_s + (row - d.q_dim);
                        scatter_lane_f32(dst, v, on);
                    } else if (on) {       // EPI_CLS: infer.c:1156-1167 penalty, then first-max argmax :1026-1037 (rows ascend per lane)
                        if (pen != 1.0f && __ldcg(g.seen + row)) v = __fdiv_rn(v, pen);
                        g.logits[row] = v;
                        if (v > bestv) { bestv = v; besti = row; }
                    }
                }
            }
            CL_STAMP();
            // ---- phase boundary ----
            if (c.epi == EPI_CLS) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bestv, o); const uint32_t oi = __shfl_xor_sync(0xffffffffu, besti, o);
                    if (oi != 0xffffffffu && (ov > bestv || (ov == bestv && oi < besti))) { bestv = ov; besti = oi; }
                }
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
            CL_STAMP();
            if (c.epi == EPI_QKV) {
                // ---- attention: this rank's (kv head, split) partial -> all replicas; then every CTA merges all heads ----
                const uint32_t kvh = rank / rpk, split = rank % rpk;
                {
                    uint32_t chunk = (range + rpk - 1) / rpk;
                    chunk = (chunk + 7u) & ~7u;
                    const uint32_t at0 = min(range, split * chunk), at1 = min(range, at0 + chunk);
                    const size_t kvl = (size_t)d.KV * d.max_seq * d.hd;
                    float *kb = g.kc + c.layer * kvl + (size_t)kvh * d.max_seq * d.hd;
                    const float *vb = g.vc + c.layer * kvl + (size_t)kvh * d.max_seq * d.hd;
                    float *outp = attn_ws + (size_t)kWarps * KVM * (d.hd + 4);
                    attn_stream_partial<KVM, kThreads, false>(d, q_s + (size_t)kvh * KVM * d.hd, kraw_s + (size_t)kvh * d.hd, vrow_s + (size_t)kvh * d.hd, kb, vb,
                                                              g.qnorm ? g.qnorm + (size_t)c.layer * d.hd : nullptr, g.knorm ? g.knorm + (size_t)c.layer * d.hd : nullptr,
                                                              g.rope_cos + (size_t)pos * (d.hd / 2), g.rope_sin + (size_t)pos * (d.hd / 2), pos, at0, at1 - at0, attn_ws, outp,
                                                              (g.trace && rank == 0 && step + 1 == g.n_steps && c.layer == d.L / 2) ? g.trace + 1040 : nullptr);
                    float *slot = part_s + (size_t)(kvh * rpk + split) * part_stride;
                    for (uint32_t idx = warp; idx < part_stride; idx += kWarps) scatter_f32(slot + idx, outp[idx]);
                }
                CL_STAMP();
                cluster_sync_all();
                CL_STAMP();
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
                CL_STAMP();
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
