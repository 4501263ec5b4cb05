// tf32_test.cu -- C-ABI unit-test entry for the hand-written tcgen05 kind::tf32 tiles used by the fused field kernels:
// validates the thread-written 128B-swizzle layout, the 3-term split accuracy, and K-major / MN-major operand descriptors.
#include "tf32_tile.cuh"
#include "mi3d_common.cuh"
#include "../../include/mi3d.h"

namespace {

// stage X [R x C] (fp32 row-major, global) into hi/lo column blocks of [R x 32]
__device__ void stage(const float* __restrict__ X, int R, int C, uint8_t* hi, uint8_t* lo, int tid, int nthreads) {
    for (int i = tid; i < R * C; i += nthreads) {
        const int r = i / C, c = i % C;
        const float v = X[i], h = ftc::tf32_hi(v);
        const uint32_t o = (uint32_t)(c >> 5) * (uint32_t)R * 128u + ftc::sw_off(r, c & 31);
        *reinterpret_cast<float*>(hi + o) = h; *reinterpret_cast<float*>(lo + o) = v - h;
    }
}

// mode 0: D = A[128xK] . B[NxK]^T (both K-major) ; mode 1: D = A'[Kx128]^T . B'[KxN] (both MN-major) ; mode 2: A K-major, B' MN-major
__global__ void __launch_bounds__(160, 1) k_tf32_tile_test(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K, int mode) {
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* sm = smem_dyn + ((1024u - (tc::smem_u32(smem_dyn) & 1023u)) & 1023u);   // offset arithmetic keeps the shared address space (STS / LDS, not generic ST / LD)
    const int swapv = mode >= 3; if (swapv) mode -= 2;
    const int a_mn = mode == 1, b_mn = mode >= 1;
    const int Ra = a_mn ? K : 128, Ca = a_mn ? 128 : K, Rb = b_mn ? K : N, Cb = b_mn ? N : K;
    const uint32_t a_bytes = (uint32_t)Ra * Ca * 4, b_bytes = (uint32_t)Rb * Cb * 4;
    uint8_t *a_hi = sm, *a_lo = sm + a_bytes, *b_hi = sm + 2 * a_bytes, *b_lo = sm + 2 * a_bytes + b_bytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * a_bytes + 2 * b_bytes);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    stage(A, Ra, Ca, a_hi, a_lo, tid, blockDim.x);
    stage(B, Rb, Cb, b_hi, b_lo, tid, blockDim.x);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
    if (warp == 4) tc::tmem_alloc(slot, 64);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *slot;
    if (warp == 4 && (tid & 31) == 0) {
        ftc::Operand oa{tc::smem_u32(a_hi), tc::smem_u32(a_lo), (uint32_t)Ra * 128u, a_mn};
        ftc::Operand ob{tc::smem_u32(b_hi), tc::smem_u32(b_lo), (uint32_t)Rb * 128u, b_mn};
        ftc::g_swap_lbo_sbo = swapv;
        ftc::issue_3tf32(tmem, oa, ob, K, ftc::idesc_tf32(128, N, a_mn, b_mn), 0);
        tc::umma_commit(bar);
    }
    if (warp < 4) {
        tc::mbar_wait(bar, 0);
        tc::tc_fence_after();
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            ftc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            for (int j = 0; j < 16; j++) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 64); }
}

}  // namespace

extern "C" int mi3d_tf32_tile_test(const float* a, const float* b, float* d, int N, int K, int mode, mi3d_stream_t stream) {
    if ((N != 16 && N != 32 && N != 64) || K % 32 || K > 128 || mode < 0 || mode > 4) return MI3D_ERR_ARG;
    const size_t smem = 1024 + 2 * (size_t)128 * K * 4 + 2 * (size_t)N * K * 4 + 64;
    MI3D_CHECK(cudaFuncSetAttribute(k_tf32_tile_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tf32_tile_test<<<1, 160, smem, (cudaStream_t)stream>>>(a, b, d, N, K, mode);
    MI3D_RETURN_LAUNCH();
}
