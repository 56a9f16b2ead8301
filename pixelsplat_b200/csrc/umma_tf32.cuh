// tcgen05 / TMEM / UMMA-descriptor helpers shared by the self-attention forward and backward kernels
// (TF32 operands in the canonical K-major no-swizzle layout; see self_attention_tc.cu for the walk-through).
#pragma once
#include "ps_common.cuh"

namespace ps {

constexpr int kSaL = 256;        // tokens per image
constexpr int kSaD = 128;        // head dimension
constexpr int kSaThreads = 256;      // warps 0-3: soft-max rows; warps 4-7: V^T staging; all: Q/K staging, epilogue
constexpr uint32_t kSaTmemCols = 512;

// fp32 -> nearest TF32 (ties away), kept in an fp32 container: the tensor core ignores the low 13
// mantissa bits, so rounding here instead of letting it truncate halves the operand error and removes its bias.
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ float4 to_tf32(float4 v) { return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w)); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes (contiguous 128 B); 8-row groups are
// adjacent (SBO = 128 B), 16-byte K chunks are `lbo` bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
    return d;                                // base offset 0, LBO mode 0, layout type 0 (SWIZZLE_NONE)
}

// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_saddr) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar_saddr, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(bar_saddr), "r"(parity) : "memory");
    }
}

// 32 consecutive TMEM columns of this thread's lane -> registers (and back).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}


// ---- staging of fp32 global tiles into the K-major no-swizzle layout (rounded to the nearest TF32) ---------
// natural: tile rows = MMA rows (M or N), the 128 channels are the K dimension.  Consecutive threads take
// consecutive rows of the same 16-byte chunk, so the shared stores are contiguous (chunk c of row r lives at
// c * LBO + r * 16, LBO = ROWS * 16); eight independent 16-byte loads are in flight per thread.
template <int ROWS>
__device__ __forceinline__ void stage_natural(unsigned char *dst, const float *__restrict__ src, size_t row_stride,
                                              int tid, int nthreads) {
    constexpr int kBatch = 8;
    constexpr int kShift = ROWS == 256 ? 8 : 7;
    static_assert(ROWS == 128 || ROWS == 256, "tile rows");
#pragma unroll 1
    for (int i0 = tid; i0 < ROWS * 32; i0 += nthreads * kBatch) {
        float4 v[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int i = i0 + j * nthreads;
            v[j] = (i < ROWS * 32) ? __ldg(reinterpret_cast<const float4 *>(src + (size_t)(i & (ROWS - 1)) * row_stride) + (i >> kShift))
                                   : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int i = i0 + j * nthreads;
            if (i < ROWS * 32)
                *reinterpret_cast<float4 *>(dst + (size_t)(i >> kShift) * (ROWS * 16) + (i & (ROWS - 1)) * 16) = to_tf32(v[j]);
        }
    }
}

// transposed: MMA rows (N) = the 128 channels, K dimension = TOKENS tokens; one 16-byte chunk = 4 consecutive
// tokens of one channel (LBO = 128 * 16), 16 scalar loads in flight per thread.
template <int TOKENS>
__device__ __forceinline__ void stage_transposed(unsigned char *dst, const float *__restrict__ src, size_t row_stride,
                                                 int tid, int nthreads) {
#pragma unroll 1
    for (int i0 = tid; i0 < 128 * (TOKENS / 4); i0 += nthreads * 4) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * nthreads;
            v[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (i < 128 * (TOKENS / 4)) {
                const float *p = src + (size_t)(4 * (i >> 7)) * row_stride + (i & 127);
                v[j].x = __ldg(p);
                v[j].y = __ldg(p + row_stride);
                v[j].z = __ldg(p + 2 * row_stride);
                v[j].w = __ldg(p + 3 * row_stride);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = i0 + j * nthreads;
            if (i < 128 * (TOKENS / 4))
                *reinterpret_cast<float4 *>(dst + (size_t)(i >> 7) * (128 * 16) + (i & 127) * 16) = to_tf32(v[j]);
        }
    }
}

}  // namespace ps
