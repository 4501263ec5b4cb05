// tf32_tile.cuh -- helpers for hand-written tcgen05 kind::tf32 tiles with a 3-term hi/lo split (fp32-class accuracy):
//   x = hi + lo (hi = tf32-rounded x):  A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo
// Storage convention for every operand tile: column blocks of [rows][32 fp32] (128-byte rows, 8-row / 1024-byte swizzle
// atoms, block stride = rows * 128 B), written by ordinary threads through sw_off().  The SAME storage serves two uses:
//   K-major  operand: rows = M or N index, columns = K   (descriptor: SBO = 1024 B, K advances by +32 B inside the row)
//   MN-major operand: rows = K index,      columns = M/N (descriptor: LBO = block stride, SBO = 1024 B, K advances by +1024 B)
// (cute::UMMA canonical layouts  K-major ((8,n),2):((8,SBO),1)  and  MN-major ((8,n),(8,k)):((1,LBO),(8,SBO)), in 16-byte units.)
#pragma once
#include "tc_gemm.cuh"

namespace ftc {

// byte offset of element (row, k) inside one [rows x 32 fp32] block (block base 1024-aligned)
__device__ __forceinline__ uint32_t sw_off(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2));
}
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r);
}
// cute::UMMA::InstrDescriptor, kind::tf32: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, a_major @15, b_major @16 (1 = MN-major)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// SW128 descriptor with an explicit leading-dimension byte offset (MN-major: distance between 32-element column blocks)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// four 4-column chunks, 8 columns apart: v[4j .. 4j+3] = columns taddr + 8j .. + 3 of this thread's TMEM lane; one wait for all four
__device__ __forceinline__ void tmem_ld4x4(uint32_t taddr, uint32_t (&v)[16]) {
    #pragma unroll
    for (int j = 0; j < 4; j++)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v[4 * j]), "=r"(v[4 * j + 1]), "=r"(v[4 * j + 2]), "=r"(v[4 * j + 3]) : "r"(taddr + 8u * (uint32_t)j));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ int g_swap_lbo_sbo = 0;   // experiment switch used by the unit-test kernel only
// One operand = (hi base, lo base) of its first block + how to walk it.
struct Operand { uint32_t hi, lo; uint32_t blk_stride; int mn_major; };

// D[128 x N] (+)= A . B over K elements with the 3-term split.  K-major operands walk K as (block, +32 B); MN-major operands as (+1024 B).
// `acc` in: 0 -> the first MMA overwrites D.  Returns 1.
__device__ __forceinline__ uint32_t issue_3tf32(uint32_t tmem_d, const Operand& A, const Operand& B, int K, uint32_t idesc, uint32_t acc) {
    for (int k8 = 0; k8 < K / 8; k8++) {
        const uint32_t oa = A.mn_major ? (uint32_t)k8 * 1024u : (uint32_t)(k8 >> 2) * A.blk_stride + (uint32_t)(k8 & 3) * 32u;
        const uint32_t ob = B.mn_major ? (uint32_t)k8 * 1024u : (uint32_t)(k8 >> 2) * B.blk_stride + (uint32_t)(k8 & 3) * 32u;
        const uint32_t la = A.mn_major ? A.blk_stride : 16u, lb = B.mn_major ? B.blk_stride : 16u;
        uint64_t dah = desc_sw128(A.hi + oa, la), dal = desc_sw128(A.lo + oa, la);
        uint64_t dbh = desc_sw128(B.hi + ob, lb), dbl = desc_sw128(B.lo + ob, lb);
        if (g_swap_lbo_sbo) {
            auto sw = [](uint64_t d) { const uint64_t l = (d >> 16) & 0x3FFF, s = (d >> 32) & 0x3FFF; return (d & ~((0x3FFFull << 16) | (0x3FFFull << 32))) | (s << 16) | (l << 32); };
            if (A.mn_major) { dah = sw(dah); dal = sw(dal); }
            if (B.mn_major) { dbh = sw(dbh); dbl = sw(dbl); }
        }
        umma_tf32(tmem_d, dah, dbh, idesc, acc); acc = 1;
        umma_tf32(tmem_d, dal, dbh, idesc, 1);
        umma_tf32(tmem_d, dah, dbl, idesc, 1);
    }
    return acc;
}

}  // namespace ftc

// ---------------------------------------------------------------------------------------------------------------------
// bf16 3-way split tiles (kind::f16, bf16 operands): x = h + m + l (each bf16, together ~24 bits) -> fp32-class products,
// used by the fused field BACKWARD because 16-bit operands may be consumed MN-major
// (kind::tf32 has no working MN-major form; the 16-bit one is verified by tools/explore_mn.py / mi3d_gemm_f16_bt).
// Storage: [rows][64 bf16] = 128-byte rows, 8-row swizzle atoms; the same tile is read K-major (rows = M/N, +32 B per
// K = 16 step) or MN-major (rows = K, +2048 B per K = 16 step, LBO = distance to the next 64-column block).
// ---------------------------------------------------------------------------------------------------------------------
#include <cuda_bf16.h>
namespace fbf {
using ftc::mbar_arrive;
using ftc::tmem_ld16;
using ftc::tmem_ld4x4;

// byte offset of element (row, c) inside a [rows x 64 bf16] tile
__device__ __forceinline__ uint32_t sw_off16(int row, int c) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((c >> 3) ^ (row & 7)) & 7) << 4) + ((c & 7) << 1));
}
// 3-way split: x = h + m + l exactly to ~24 bits (fp32-class), each part bf16
__device__ __forceinline__ void split3(float x, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
    h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    m = __float2bfloat16_rn(r1);
    l = __float2bfloat16_rn(r1 - __bfloat162float(m));
}
// Two values at a time: one packed conversion (F2FP.BF16.F32.PACK_AB, full-rate ALU) per part instead of two scalar F2F.BF16.F32, which
// issue at a quarter of that rate -- the chain kernel's owners convert 576 values per evaluation and were bound by that pipe.
// Same round-to-nearest-even as split3: bit-identical parts.  x0 lands in the low half (the lower column / K index).
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void split3x2(float x0, float x1, uint32_t& h, uint32_t& m, uint32_t& l) {
    h = pack_bf16x2(x0, x1);
    const float r0 = x0 - __uint_as_float(h << 16), r1 = x1 - __uint_as_float(h & 0xffff0000u);
    m = pack_bf16x2(r0, r1);
    l = pack_bf16x2(r0 - __uint_as_float(m << 16), r1 - __uint_as_float(m & 0xffff0000u));
}
// pack 8 floats into one 16-byte chunk per part
__device__ __forceinline__ void split8(const float (&x)[8], uint4& ph, uint4& pm, uint4& pl) {
    uint32_t h[4], m[4], l[4];
    #pragma unroll
    for (int i = 0; i < 4; i++) split3x2(x[2 * i], x[2 * i + 1], h[i], m[i], l[i]);
    ph = make_uint4(h[0], h[1], h[2], h[3]); pm = make_uint4(m[0], m[1], m[2], m[3]); pl = make_uint4(l[0], l[1], l[2], l[3]);
}
// kind::f16 instruction descriptor with BF16 operands (format 1), fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// one operand: three part tiles laid out back to back (h, m, l at +0, +part_stride, +2 part_stride)
struct Operand { uint32_t base; uint32_t part_stride; uint32_t lbo; int mn_major; };

// D[128 x N] (+)= A . B over K elements (K % 16 == 0; K-major operands: K <= 64).  Six MMAs per K = 16 step keep every
// product term down to 2^-16 of the leading one: hh, hm, mh, hl, lh, mm  -> fp32-class accuracy.
// The descriptors of one operand differ only in the 14-bit start-address field: one base descriptor per operand, then a 32-bit add of
// (byte offset >> 4) per use (all tiles live below 256 KB, the field cannot carry into the LBO field).
__device__ __forceinline__ uint32_t issue_bf16x3(uint32_t tmem_d, const Operand& A, const Operand& B, int K, uint32_t idesc, uint32_t acc) {
    const uint64_t a0 = ftc::desc_sw128(A.base, A.mn_major ? A.lbo : 16u), b0 = ftc::desc_sw128(B.base, B.mn_major ? B.lbo : 16u);
    const uint32_t a_lo = (uint32_t)a0, b_lo = (uint32_t)b0;
    const uint64_t a_hi = a0 & 0xFFFFFFFF00000000ull, b_hi = b0 & 0xFFFFFFFF00000000ull;
    #pragma unroll
    for (int k16 = 0; k16 < K / 16; k16++) {
        const uint32_t oa = (A.mn_major ? (uint32_t)k16 * 2048u : (uint32_t)k16 * 32u) >> 4;
        const uint32_t ob = (B.mn_major ? (uint32_t)k16 * 2048u : (uint32_t)k16 * 32u) >> 4;
        uint64_t da[3], db[3];
        #pragma unroll
        for (int q = 0; q < 3; q++) {
            da[q] = a_hi | (uint64_t)(a_lo + oa + q * (A.part_stride >> 4));
            db[q] = b_hi | (uint64_t)(b_lo + ob + q * (B.part_stride >> 4));
        }
        tc::umma_f16(tmem_d, da[0], db[0], idesc, acc); acc = 1;
        tc::umma_f16(tmem_d, da[0], db[1], idesc, 1);
        tc::umma_f16(tmem_d, da[1], db[0], idesc, 1);
        tc::umma_f16(tmem_d, da[0], db[2], idesc, 1);
        tc::umma_f16(tmem_d, da[2], db[0], idesc, 1);
        tc::umma_f16(tmem_d, da[1], db[1], idesc, 1);
    }
    return acc;
}
}  // namespace fbf
