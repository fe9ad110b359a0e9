// Tensor-core building blocks for the Tucker-fit iteration: small FP32-grade GEMMs
//   D[128 x N] = A[128 x K] * B[N x K]^T      (K = 8 or 16, N <= 256)
// on tcgen05 with kind::tf32 and the 3xTF32 operand split (hi = tf32(v), lo = v - hi; lo*hi + hi*lo + hi*hi), the
// A operand written into shared memory by the threads themselves (one thread = one sample = one row = one TMEM lane).
//
// Operand layout: the canonical K-major NO-SWIZZLE form of the UMMA shared-memory descriptor
// (cute::UMMA::make_umma_desc, LayoutType::INTERLEAVE: ((8,n),2):((1,SBO),LBO) in 16-byte units):
//   element (row, k)  ->  (row / 8) * SBO + (k / 4) * LBO + (row % 8) * 16 + (k % 4) * 4   bytes
// i.e. 8-row x 16-byte "core matrices" of 128 contiguous bytes; LBO = distance between the K chunks of one MMA,
// SBO = distance between 8-row groups.  We store [row group][k chunk][8 rows][16 B]: LBO = 128, SBO = (K/4) * 128.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace nlml {
namespace ttc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) of an operand tile with K columns (K % 4 == 0)
__host__ __device__ constexpr int op_offset(int row, int k, int K) {
    return (row / 8) * ((K / 4) * 128) + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4;
}
__host__ __device__ constexpr int op_bytes(int rows, int K) { return rows * K * 4; }

// descriptor for the K-slice [k0, k0+8) of such a tile
__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t tile_addr, int K, int k0) {
    const uint32_t addr = tile_addr + (uint32_t)((k0 / 4) * 128);
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);              // start address
    d |= (uint64_t)(128 >> 4) << 16;                      // LBO: next K chunk (4 tf32) of the same rows
    d |= (uint64_t)(((K / 4) * 128) >> 4) << 32;          // SBO: next 8-row group
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    // layout type 0 = no swizzle
    return d;
}
// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// the same with the B operand MN-major ("transposed"): bit 16
__host__ __device__ constexpr uint32_t make_idesc_tf32_bt(int M, int N) { return make_idesc_tf32(M, N) | (1u << 16); }
// B operand read from the K-major image of its TRANSPOSE: the tile was stored as rows = k (Kt of them), contiguous
// dimension = n (N of them), i.e. element (n, k) at op_offset(k, n, N).  As an MN-major operand (canonical no-swizzle
// form ((4,1,m),(8,k)):((1,4,SBO),(16 B,LBO))): 4 consecutive n are contiguous, 8 consecutive k are 16 B apart,
// SBO = 128 B (next 4 n), LBO = (N/4)*128 B (next 8 k).  Descriptor for the K-slice [k0, k0+8).
__device__ __forceinline__ uint64_t make_desc_noswz_bt(uint32_t tile_addr, int N, int k0) {
    const uint32_t lbo = (uint32_t)((N / 4) * 128);
    const uint32_t addr = tile_addr + (uint32_t)(k0 / 8) * lbo;
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same MMA with the A operand in TENSOR MEMORY (lane = row, 8 consecutive 32-bit columns = one K step)
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_to(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// generic-proxy shared-memory writes -> visible to the tensor core's async proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// hi = value rounded to TF32 (10 mantissa bits, round to nearest even on the dropped 13 bits), lo = v - hi (exact in FP32)
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    uint32_t u = __float_as_uint(v);
    u += 0x00000fffu + ((u >> 13) & 1u);
    u &= 0xffffe000u;
    hi = __uint_as_float(u);
    lo = v - hi;
}

// The same split for operands rebuilt every iteration: one cvt.rna (round to nearest, ties away) instead of the
// integer sequence; lo = v - hi is exact either way.
__device__ __forceinline__ void split_tf32_fast(float v, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    lo = v - hi;
}

// D += A * B^T over K (multiple of 8) with the 3-pass split; a/b tiles in the no-swizzle layout above.
// Issued by ONE thread.  `first` = overwrite the accumulator with the first MMA.
__device__ __forceinline__ void issue_gemm_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                                  int K, int N, bool first) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    for (int k0 = 0; k0 < K; k0 += 8) {
        const uint64_t dah = make_desc_noswz(a_hi, K, k0), dal = make_desc_noswz(a_lo, K, k0);
        const uint64_t dbh = make_desc_noswz(b_hi, K, k0), dbl = make_desc_noswz(b_lo, K, k0);
        umma_tf32(tmem_d, dal, dbh, idesc, !(first && k0 == 0));   // small terms first
        umma_tf32(tmem_d, dah, dbl, idesc, 1);
        umma_tf32(tmem_d, dah, dbh, idesc, 1);
    }
}

// the same with the B tile stored as the K-major image of its transpose (one copy of a tile then serves two GEMMs:
// as the K-major B operand of one and the MN-major B operand of the other)
__device__ __forceinline__ void issue_gemm_3xtf32_bt(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                                     int K, int N, bool first) {
    const uint32_t idesc = make_idesc_tf32_bt(128, N);
    for (int k0 = 0; k0 < K; k0 += 8) {
        const uint64_t dah = make_desc_noswz(a_hi, K, k0), dal = make_desc_noswz(a_lo, K, k0);
        const uint64_t dbh = make_desc_noswz_bt(b_hi, N, k0), dbl = make_desc_noswz_bt(b_lo, N, k0);
        umma_tf32(tmem_d, dal, dbh, idesc, !(first && k0 == 0));
        umma_tf32(tmem_d, dah, dbl, idesc, 1);
        umma_tf32(tmem_d, dah, dbh, idesc, 1);
    }
}

// 3xTF32 GEMM with the A operand in tensor memory: hi part in columns [a_tmem, a_tmem + K), lo part in the next K.
__device__ __forceinline__ void issue_gemm_3xtf32_ta(uint32_t tmem_d, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, int K, int N,
                                                     bool first) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    for (int k0 = 0; k0 < K; k0 += 8) {
        const uint64_t dbh = make_desc_noswz(b_hi, K, k0), dbl = make_desc_noswz(b_lo, K, k0);
        umma_tf32_ta(tmem_d, a_tmem + K + k0, dbh, idesc, !(first && k0 == 0));   // lo * hi: small terms first
        umma_tf32_ta(tmem_d, a_tmem + k0, dbl, idesc, 1);
        umma_tf32_ta(tmem_d, a_tmem + k0, dbh, idesc, 1);
    }
}

// Round-to-nearest (ties away) TF32 split without cvt.rna (which sm_100a emulates with four instructions incl. an
// infinity test): add half an ulp of TF32 to the bit pattern, clear the 13 dropped bits; lo = v - hi is exact.
__device__ __forceinline__ void split_tf32_bits(float v, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
    lo = v - hi;
}

// ---- issue loops for the run-time-rank kernel (tucker_gen.cuh): the descriptors are built once per GEMM and advanced by
// one add per k-step (the K-slice [k0, k0+8) starts 256 B further: +16 in the descriptor's address field), the three MMAs
// of a k-step share one asm block.  Meant to be executed by a whole warp with warp-uniform arguments (descriptor
// arithmetic stays in the uniform datapath); only the elected lane issues.
__device__ __forceinline__ void mma3_ss(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 q, %5, %5;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %2, %3, %5, p;\n\t"   // lo * hi (small terms first)
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %4, %5, q;\n\t"   // hi * lo
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %3, %5, q;\n\t}"  // hi * hi
        ::"r"(d), "l"(ah), "l"(al), "l"(bh), "l"(bl), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma3_ts(uint32_t d, uint32_t ah, uint32_t al, uint64_t bh, uint64_t bl, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 q, %5, %5;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %3, %5, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %4, %5, q;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %3, %5, q;\n\t}"
        ::"r"(d), "r"(ah), "r"(al), "l"(bh), "l"(bl), "r"(idesc), "r"(acc)
        : "memory");
}
// commit by the elected lane of a converged warp
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}
// D = A * B^T (fresh accumulator), A and B tiles in shared memory (no-swizzle K-major), K multiple of 8.
// Executed by a CONVERGED warp with warp-uniform arguments.
__device__ __forceinline__ void gemm3_ss(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int K, int N) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    uint64_t ah = make_desc_noswz(a_hi, K, 0), al = make_desc_noswz(a_lo, K, 0);
    uint64_t bh = make_desc_noswz(b_hi, K, 0), bl = make_desc_noswz(b_lo, K, 0);
    for (int k = 0; k < K; k += 8) {
        mma3_ss(d, ah, al, bh, bl, idesc, k > 0 ? 1u : 0u);
        ah += 16; al += 16; bh += 16; bl += 16;
    }
}
// the same with the A operand in tensor memory (hi in columns [a, a+K), lo in [a+K, a+2K))
__device__ __forceinline__ void gemm3_ts(uint32_t d, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, int K, int N) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    uint64_t bh = make_desc_noswz(b_hi, K, 0), bl = make_desc_noswz(b_lo, K, 0);
    for (int k = 0; k < K; k += 8) {
        mma3_ts(d, a_tmem + k, a_tmem + K + k, bh, bl, idesc, k > 0 ? 1u : 0u);
        bh += 16; bl += 16;
    }
}

// ---- FP16 hi/lo variant of the same building blocks: kind::f16, K = 16 per MMA (half the MMA count of 3xTF32 at the same
// cycle cost per instruction and half the operand bytes; the same 22 operand bits: hi = fp16(v), lo = fp16(v - hi)).  The
// operands must sit in FP16's range: the callers scale them by powers of two.
__host__ __device__ constexpr int op16_offset(int row, int k, int K) {   // K-major no-swizzle: core matrix = 8 rows x 8 halves
    return (row / 8) * ((K / 8) * 128) + (k / 8) * 128 + (row % 8) * 16 + (k % 8) * 2;
}
__host__ __device__ constexpr int op16_bytes(int rows, int K) { return rows * K * 2; }
__device__ __forceinline__ uint64_t make_desc16_noswz(uint32_t tile_addr, int K, int k0) {
    const uint32_t addr = tile_addr + (uint32_t)((k0 / 8) * 128);
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;                      // LBO: next K chunk (8 halves) of the same rows
    d |= (uint64_t)(((K / 8) * 128) >> 4) << 32;          // SBO: next 8-row group
    d |= (uint64_t)1 << 46;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_f16_f32(int M, int N) {   // D = F32, A = B = F16, both K-major
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma3h_ss(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 q, %5, %5;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %3, %5, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %4, %5, q;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %3, %5, q;\n\t}"
        ::"r"(d), "l"(ah), "l"(al), "l"(bh), "l"(bl), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma3h_ts(uint32_t d, uint32_t ah, uint32_t al, uint64_t bh, uint64_t bl, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p, q, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.b32 q, %5, %5;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], %3, %5, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %4, %5, q;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %3, %5, q;\n\t}"
        ::"r"(d), "r"(ah), "r"(al), "l"(bh), "l"(bl), "r"(idesc), "r"(acc)
        : "memory");
}
// D = A * B^T (fresh accumulator), K multiple of 16; executed by a converged warp with warp-uniform arguments
__device__ __forceinline__ void gemm3h_ss(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int K, int N) {
    const uint32_t idesc = make_idesc_f16_f32(128, N);
    uint64_t ah = make_desc16_noswz(a_hi, K, 0), al = make_desc16_noswz(a_lo, K, 0);
    uint64_t bh = make_desc16_noswz(b_hi, K, 0), bl = make_desc16_noswz(b_lo, K, 0);
    for (int k = 0; k < K; k += 16) {
        mma3h_ss(d, ah, al, bh, bl, idesc, k > 0 ? 1u : 0u);
        ah += 16; al += 16; bh += 16; bl += 16;   // 16 halves = two 128-byte chunks further
    }
}
// A operand in tensor memory, two halves per 32-bit column: hi in columns [a, a + K/2), lo in [a + K/2, a + K)
__device__ __forceinline__ void gemm3h_ts(uint32_t d, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, int K, int N) {
    const uint32_t idesc = make_idesc_f16_f32(128, N);
    uint64_t bh = make_desc16_noswz(b_hi, K, 0), bl = make_desc16_noswz(b_lo, K, 0);
    for (int k = 0; k < K; k += 16) {
        mma3h_ts(d, a_tmem + k / 2, a_tmem + K / 2 + k / 2, bh, bl, idesc, k > 0 ? 1u : 0u);
        bh += 16; bl += 16;
    }
}
__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// check kernel of the FP16 variant: mode 2 = A in shared memory, mode 3 = A in tensor memory (packed pairs)
__global__ void __launch_bounds__(128) tc16_check_kernel(const float* A, const float* B, float* D, int K, int N, int mode) {
    extern __shared__ __align__(1024) uint8_t csm16[];
    uint8_t* a_hi = csm16;
    uint8_t* a_lo = a_hi + op16_bytes(128, K);
    uint8_t* b_hi = a_lo + op16_bytes(128, K);
    uint8_t* b_lo = b_hi + op16_bytes(N, K);
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + op16_bytes(N, K));
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int k = 0; k < K; ++k) {
        __half hi, lo;
        split_half(A[tid * K + k], hi, lo);
        *reinterpret_cast<__half*>(a_hi + op16_offset(tid, k, K)) = hi;
        *reinterpret_cast<__half*>(a_lo + op16_offset(tid, k, K)) = lo;
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        __half hi, lo;
        split_half(B[idx], hi, lo);
        *reinterpret_cast<__half*>(b_hi + op16_offset(n, k, K)) = hi;
        *reinterpret_cast<__half*>(b_lo + op16_offset(n, k, K)) = lo;
    }
    fence_async_smem();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t col_a = 256;
    if (mode == 3) {   // this thread's row of A, packed (element 2c in the low half of column c)
        for (int c = 0; c < K / 2; ++c) {
            __half h0, l0, h1, l1;
            split_half(A[tid * K + 2 * c], h0, l0);
            split_half(A[tid * K + 2 * c + 1], h1, l1);
            const uint32_t wh = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
            const uint32_t wl = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + col_a + c), "r"(wh) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + col_a + K / 2 + c), "r"(wl) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) {
        if (mode == 3) gemm3h_ts(tmem, tmem + col_a, smem_u32(b_hi), smem_u32(b_lo), K, N);
        else gemm3h_ss(tmem, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo), K, N);
        umma_commit_elect(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_load32(taddr + c0, v);
        for (int j = 0; j < 32 && c0 + j < N; ++j) D[tid * N + c0 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ---- stand-alone check kernel: one CTA, D[128][N] = A[128][K] * B[N][K]^T --------------------------------
struct TcCheckArgs {
    const float* A;   // [128][K]
    const float* B;   // [N][K]
    float* D;         // [128][N]
    int K, N;
    int mode;         // 0: B stored K-major.  1: B stored as the K-major image of its transpose, read as an MN-major operand
};

__global__ void __launch_bounds__(128) tc_check_kernel(const __grid_constant__ TcCheckArgs a) {
    extern __shared__ __align__(1024) uint8_t csm[];
    const int K = a.K, N = a.N;
    uint8_t* a_hi = csm;
    uint8_t* a_lo = a_hi + op_bytes(128, K);
    uint8_t* b_hi = a_lo + op_bytes(128, K);
    uint8_t* b_lo = b_hi + op_bytes(N, K);
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + op_bytes(N, K));
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // operands: thread = row of A; B rows spread over the threads
    for (int k = 0; k < K; ++k) {
        float hi, lo;
        split_tf32(a.A[tid * K + k], hi, lo);
        *reinterpret_cast<float*>(a_hi + op_offset(tid, k, K)) = hi;
        *reinterpret_cast<float*>(a_lo + op_offset(tid, k, K)) = lo;
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int n = idx / K, k = idx % K;
        float hi, lo;
        split_tf32(a.B[idx], hi, lo);
        const int off = a.mode == 1 ? op_offset(k, n, N) : op_offset(n, k, K);
        *reinterpret_cast<float*>(b_hi + off) = hi;
        *reinterpret_cast<float*>(b_lo + off) = lo;
    }
    fence_async_smem();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (tid == 0) {
        if (a.mode == 1) issue_gemm_3xtf32_bt(tmem, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo), K, N, true);
        else issue_gemm_3xtf32(tmem, smem_u32(a_hi), smem_u32(a_lo), smem_u32(b_hi), smem_u32(b_lo), K, N, true);
        umma_commit_to(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_load32(taddr + c0, v);
        for (int j = 0; j < 32 && c0 + j < N; ++j) a.D[tid * N + c0 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

}  // namespace ttc
}  // namespace nlml
