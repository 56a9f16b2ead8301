// Backward of the dense per-image self-attention (self_attention_tc.cu) on the same tcgen05 / TMEM path:
// autograd of softmax(Q K^T * scale) V for the ViT blocks of ImageSelfAttention
// (/root/reference/src/model/encoder/epipolar/image_self_attention.py:57-79 ->
//  /root/reference/src/model/transformer/attention.py:54-70 with z = None).
//
// Per (image, head), with Pn the forward's probabilities -- rebuilt from the saved per-row (max, 1 / sum) with
// the same TF32 roundings, so the backward differentiates the forward that actually ran:
//     dPn = dO V^T      D_i = dO_i . O_i      dS = Pn o (dPn - D) * scale
//     dQ = dS K         dK = dS^T Q           dV = Pn^T dO
// Two CTA roles (grid.x = 4), 8 warps, every contraction a tcgen05.mma with FP32 accumulation in TMEM:
//   role 0/1 "query half I" -> dQ_I.  S = Q_I K^T (TMEM cols 0..255) and dPn = dO_I V^T (cols 256..511) with
//            A / B staged K-major in shared memory; thread = query row turns S into dS in place; then
//            dQ_I = dS K with A = dS read straight from TMEM and B = K^T staged transposed (the forward's P V
//            step with other operands), accumulated over dPn's dead columns.
//   role 2/3 "key half J"   -> dK_J, dV_J.  The transposed problem, so that the rows a CTA owns are the rows it
//            sums over: S^T = K_J Q_I'^T and dPn^T = V_J dO_I'^T for the two query halves I' in turn (128 TMEM
//            columns each), thread = key row builds Pn^T and dS^T in place (the per-query constants max, 1 / sum
//            and D are per COLUMN here: 256-entry shared arrays), then dV_J += Pn^T dO_I' and dK_J += dS^T Q_I'
//            with A from TMEM and B = dO_I'^T / Q_I'^T staged transposed; the two accumulators own the other 256
//            TMEM columns across both I'.
// Shared memory: 3 x 64 KB operand buffers (K_J, V_J persistent + one rotating buffer in role 2/3; Q/dO + K/V in
// role 0/1) -- the forward's 192 KB; operands are rounded to the nearest TF32 on the way in, like the forward's.
#include "umma_tf32.cuh"

namespace ps {

namespace {

constexpr uint32_t kLbo128 = 128 * 16;   // bytes between 16-byte K chunks of a 128-row tile
constexpr uint32_t kLbo256 = 256 * 16;

__device__ __forceinline__ void sync_before_mma() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic -> async proxy (smem operands)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void wait_mma(uint32_t bar, uint32_t &phase) {
    mbar_wait(bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

}  // namespace

__global__ void __launch_bounds__(kSaThreads, 1)
k_self_attention_tc_bwd(const float *__restrict__ qkv, const float *__restrict__ out, const float *__restrict__ d_out,
                        const float *__restrict__ stats, float *__restrict__ d_qkv, int n_heads, float scale,
                        float scale_log2e) {
    extern __shared__ __align__(128) unsigned char s_sa[];
    unsigned char *buf0 = s_sa, *buf1 = s_sa + 64 * 1024, *buf2 = s_sa + 128 * 1024;
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_sa + 192 * 1024);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_sa + 192 * 1024 + 16);
    float *s_mb = reinterpret_cast<float *>(s_sa + 192 * 1024 + 64);     // [256] row max * scale * log2 e
    float *s_inv = s_mb + 256;                                            // [256] 1 / row sum
    float *s_D = s_inv + 256;                                             // [256] dO_i . O_i
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = blockIdx.x >> 1, half = blockIdx.x & 1, head = blockIdx.y, img = blockIdx.z;
    const int inner = n_heads * kSaD;
    const size_t rs3 = 3 * (size_t)inner, rs1 = (size_t)inner;            // floats per token in qkv / out
    const float *q_img = qkv + (size_t)img * kSaL * rs3 + (size_t)head * kSaD;
    const float *k_img = q_img + inner, *v_img = q_img + 2 * inner;
    const float *o_img = out + (size_t)img * kSaL * rs1 + (size_t)head * kSaD;
    const float *do_img = d_out + (size_t)img * kSaL * rs1 + (size_t)head * kSaD;
    float *dq_img = d_qkv + (size_t)img * kSaL * rs3 + (size_t)head * kSaD;
    float *dk_img = dq_img + inner, *dv_img = dq_img + 2 * inner;
    const uint32_t bar_a = smem_u32(bar);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "n"(kSaTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // per-query constants of all 256 queries: (max, 1 / sum) saved by the forward, D = dO . O
    {
        const int i = tid;                                                  // kSaThreads == kSaL
        const float2 st = reinterpret_cast<const float2 *>(stats)[((size_t)img * n_heads + head) * kSaL + i];
        s_mb[i] = st.x;
        s_inv[i] = st.y;
        const float4 *a = reinterpret_cast<const float4 *>(do_img + (size_t)i * rs1);
        const float4 *b = reinterpret_cast<const float4 *>(o_img + (size_t)i * rs1);
        float acc = 0.0f;
#pragma unroll 8
        for (int c = 0; c < kSaD / 4; ++c) {
            const float4 x = __ldg(a + c), y = __ldg(b + c);
            acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        }
        s_D[i] = acc;
    }
    uint32_t phase = 0;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;             // this warp's TMEM lane quarter
    const int row = (warp & 3) * 32 + lane;                                   // row of the CTA's 128-row half

    if (role == 0) {
        // ================================================================= dQ for query half `half`
        unsigned char *sQ = buf0, *sK = buf1;                                 // 64 KB + 128 KB
        stage_natural<128>(sQ, q_img + (size_t)half * 128 * rs3, rs3, tid, kSaThreads);
        stage_natural<256>(sK, k_img, rs3, tid, kSaThreads);
        sync_before_mma();
        const uint32_t tmem = *tmem_slot;
        const uint32_t tmem_S = tmem, tmem_dP = tmem + 256, tmem_dQ = tmem + 256;
        if (tid == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 256);
#pragma unroll 1
            for (int k = 0; k < kSaD / 8; ++k)
                mma_tf32_ss(tmem_S, umma_desc(smem_u32(sQ) + k * 2 * kLbo128, kLbo128, 128),
                            umma_desc(smem_u32(sK) + k * 2 * kLbo256, kLbo256, 128), idesc, k > 0);
            umma_commit(bar_a);
        }
        wait_mma(bar_a, phase);
        // dPn = dO_I V^T
        stage_natural<128>(sQ, do_img + (size_t)half * 128 * rs1, rs1, tid, kSaThreads);
        stage_natural<256>(sK, v_img, rs3, tid, kSaThreads);
        sync_before_mma();
        if (tid == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 256);
#pragma unroll 1
            for (int k = 0; k < kSaD / 8; ++k)
                mma_tf32_ss(tmem_dP, umma_desc(smem_u32(sQ) + k * 2 * kLbo128, kLbo128, 128),
                            umma_desc(smem_u32(sK) + k * 2 * kLbo256, kLbo256, 128), idesc, k > 0);
            umma_commit(bar_a);
        }
        wait_mma(bar_a, phase);
        if (warp >= 4) {
            // K^T into the K/V buffer (V is dead: its MMAs have completed)
            stage_transposed<256>(sK, k_img, rs3, tid - 128, 128);
        } else {
            // thread = query row: S -> dS in place (TF32)
            const int i = half * 128 + row;
            const float mb = s_mb[i], inv = s_inv[i], Di = s_D[i];
            for (int c = 0; c < 256; c += 32) {
                float sv[32], dp[32];
                tmem_ld32(tmem_S + lane_addr + c, sv);
                tmem_ld32(tmem_dP + lane_addr + c, dp);
#pragma unroll
                for (int x = 0; x < 32; ++x) {
                    const float p = to_tf32(exp2f(sv[x] * scale_log2e - mb)) * inv;
                    sv[x] = to_tf32(p * (dp[x] - Di) * scale);
                }
                tmem_st32(tmem_S + lane_addr + c, sv);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        sync_before_mma();
        if (tid == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, 128);
#pragma unroll 1
            for (int k = 0; k < kSaL / 8; ++k)
                mma_tf32_ts(tmem_dQ, tmem_S + k * 8, umma_desc(smem_u32(sK) + k * 2 * kLbo128, kLbo128, 128), idesc, k > 0);
            umma_commit(bar_a);
        }
        wait_mma(bar_a, phase);
        {
            const int c0 = (warp >> 2) * 64;
            float *dst = dq_img + (size_t)(half * 128 + row) * rs3 + c0;
            for (int c = 0; c < 64; c += 32) {
                float v[32];
                tmem_ld32(tmem_dQ + lane_addr + c0 + c, v);
#pragma unroll
                for (int x = 0; x < 32; x += 4)
                    *reinterpret_cast<float4 *>(dst + c + x) = make_float4(v[x], v[x + 1], v[x + 2], v[x + 3]);
            }
        }
    } else {
        // ================================================================= dK, dV for key half `half`
        unsigned char *sKj = buf0, *sVj = buf1, *sC = buf2;
        stage_natural<128>(sKj, k_img + (size_t)half * 128 * rs3, rs3, tid, kSaThreads);
        stage_natural<128>(sVj, v_img + (size_t)half * 128 * rs3, rs3, tid, kSaThreads);
        sync_before_mma();
        const uint32_t tmem = *tmem_slot;
        const uint32_t tmem_ST = tmem, tmem_DPT = tmem + 128, tmem_dV = tmem + 256, tmem_dK = tmem + 384;
        const uint32_t idesc = umma_idesc_tf32(128, 128);
#pragma unroll 1
        for (int ih = 0; ih < 2; ++ih) {
            const float *q_i = q_img + (size_t)ih * 128 * rs3, *do_i = do_img + (size_t)ih * 128 * rs1;
            // S^T = K_J Q_I'^T
            stage_natural<128>(sC, q_i, rs3, tid, kSaThreads);
            sync_before_mma();
            if (tid == 0) {
#pragma unroll 1
                for (int k = 0; k < kSaD / 8; ++k)
                    mma_tf32_ss(tmem_ST, umma_desc(smem_u32(sKj) + k * 2 * kLbo128, kLbo128, 128),
                                umma_desc(smem_u32(sC) + k * 2 * kLbo128, kLbo128, 128), idesc, k > 0);
                umma_commit(bar_a);
            }
            wait_mma(bar_a, phase);
            // dPn^T = V_J dO_I'^T
            stage_natural<128>(sC, do_i, rs1, tid, kSaThreads);
            sync_before_mma();
            if (tid == 0) {
#pragma unroll 1
                for (int k = 0; k < kSaD / 8; ++k)
                    mma_tf32_ss(tmem_DPT, umma_desc(smem_u32(sVj) + k * 2 * kLbo128, kLbo128, 128),
                                umma_desc(smem_u32(sC) + k * 2 * kLbo128, kLbo128, 128), idesc, k > 0);
                umma_commit(bar_a);
            }
            wait_mma(bar_a, phase);
            if (warp >= 4) {
                stage_transposed<128>(sC, do_i, rs1, tid - 128, 128);         // dO_I'^T (B of dV)
            } else {
                // thread = key row: S^T -> Pn^T, dPn^T -> dS^T, in place; the query constants are per column
                for (int c = 0; c < 128; c += 32) {
                    float sv[32], dp[32];
                    tmem_ld32(tmem_ST + lane_addr + c, sv);
                    tmem_ld32(tmem_DPT + lane_addr + c, dp);
#pragma unroll
                    for (int x = 0; x < 32; ++x) {
                        const int i = ih * 128 + c + x;
                        const float p = to_tf32(exp2f(sv[x] * scale_log2e - s_mb[i])) * s_inv[i];
                        sv[x] = to_tf32(p);
                        dp[x] = to_tf32(p * (dp[x] - s_D[i]) * scale);
                    }
                    tmem_st32(tmem_ST + lane_addr + c, sv);
                    tmem_st32(tmem_DPT + lane_addr + c, dp);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            sync_before_mma();
            if (tid == 0) {                                                   // dV_J += Pn^T dO_I'
#pragma unroll 1
                for (int k = 0; k < 128 / 8; ++k)
                    mma_tf32_ts(tmem_dV, tmem_ST + k * 8, umma_desc(smem_u32(sC) + k * 2 * kLbo128, kLbo128, 128), idesc,
                                (ih > 0) || (k > 0));
                umma_commit(bar_a);
            }
            wait_mma(bar_a, phase);
            stage_transposed<128>(sC, q_i, rs3, tid, kSaThreads);            // Q_I'^T (B of dK)
            sync_before_mma();
            if (tid == 0) {                                                   // dK_J += dS^T Q_I'
#pragma unroll 1
                for (int k = 0; k < 128 / 8; ++k)
                    mma_tf32_ts(tmem_dK, tmem_DPT + k * 8, umma_desc(smem_u32(sC) + k * 2 * kLbo128, kLbo128, 128), idesc,
                                (ih > 0) || (k > 0));
                umma_commit(bar_a);
            }
            wait_mma(bar_a, phase);
        }
        {
            const int c0 = (warp >> 2) * 64;
            float *dk = dk_img + (size_t)(half * 128 + row) * rs3 + c0;
            float *dv = dv_img + (size_t)(half * 128 + row) * rs3 + c0;
            for (int c = 0; c < 64; c += 32) {
                float v[32];
                tmem_ld32(tmem_dK + lane_addr + c0 + c, v);
#pragma unroll
                for (int x = 0; x < 32; x += 4)
                    *reinterpret_cast<float4 *>(dk + c + x) = make_float4(v[x], v[x + 1], v[x + 2], v[x + 3]);
                tmem_ld32(tmem_dV + lane_addr + c0 + c, v);
#pragma unroll
                for (int x = 0; x < 32; x += 4)
                    *reinterpret_cast<float4 *>(dv + c + x) = make_float4(v[x], v[x + 1], v[x + 2], v[x + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(*tmem_slot), "n"(kSaTmemCols) : "memory");
}

}  // namespace ps

extern "C" PS_API int ps_self_attention_backward(int32_t n_images, int32_t tokens, int32_t heads, int32_t dim_head,
                                                 const float *qkv, const float *out, const float *d_out,
                                                 const float *stats, float scale, float *d_qkv, void *stream) {
    using namespace ps;
    if (n_images < 1 || heads < 1 || heads > 16 || !qkv || !out || !d_out || !stats || !d_qkv) {
        set_error("ps_self_attention_backward: bad argument");
        return PS_ERR_INVALID_ARGUMENT;
    }
    if (tokens != kSaL || dim_head != kSaD) {
        set_error("ps_self_attention_backward: only 256 tokens x 128-dim heads are supported (got %d x %d)", tokens, dim_head);
        return PS_ERR_UNSUPPORTED;
    }
    if (((uintptr_t)qkv | (uintptr_t)out | (uintptr_t)d_out | (uintptr_t)d_qkv | (uintptr_t)stats) & 15) {
        set_error("ps_self_attention_backward: pointers must be 16-byte aligned");
        return PS_ERR_INVALID_ARGUMENT;
    }
    const size_t smem = 192 * 1024 + 64 + 3 * 256 * sizeof(float);
    static unsigned long long attr_devices = 0;
    if (first_use_on_device(attr_devices)) {
        PS_CUDA_CHECK(cudaFuncSetAttribute(k_self_attention_tc_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(4, heads, n_images);
    k_self_attention_tc_bwd<<<grid, kSaThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        qkv, out, d_out, stats, d_qkv, heads, scale, scale * 1.4426950408889634f);
    PS_LAUNCH_CHECK("k_self_attention_tc_bwd");
    return PS_OK;
}
