// Hardware-semantics probe (test infrastructure of the tensor-core kernels, not on the product path):
// does a K-major SWIZZLE_128B shared-memory descriptor whose start address is offset by r0 rows (r0*128 B, i.e.
// not 1024 B aligned) read rows r0..r0+127 of an image that was written with the swizzle phase of the ABSOLUTE
// row index?  Two encodings are tried: base_offset = 0 and base_offset = (start >> 7) & 7.
#include "../../include/srlz.h"
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace srlz {

__global__ void __launch_bounds__(128, 1) desc_shift_probe_kernel(float* __restrict__ out, int r0, int mode, int mn_major) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    // A image: 256 rows x 128 B (64 bf16), B image: 64 rows x 128 B = identity
    unsigned char* A = smem;
    unsigned char* B = smem + 256 * 128;
    const uint32_t bar = base + 256 * 128 + 64 * 128;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 256 * 128 + 64 * 128 + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < 256 * 64; e += 128) {
        const int r = e >> 6, k = e & 63;
        const float v = (k == 0) ? (float)(r % 250) : (float)k;
        const int byte = r * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(A + byte) = __float2bfloat16_rn(v);
    }
    for (int e = tid; e < 64 * 64; e += 128) {
        const int n = e >> 6, k = e & 63;
        const int byte = n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(B + byte) = __float2bfloat16_rn(n == k ? 1.f : 0.f);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(smem_u32(tmem_ptr_smem), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp == 0 && lane == 0) {
        const uint32_t a_addr = base + r0 * 128;
        uint64_t adesc = make_desc_sw128(a_addr);
        if (mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;
        const uint64_t bdesc = make_desc_sw128(base + 256 * 128);
        uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        if (mn_major) idesc |= (1u << 15);  // A read as MN-major: D[m][n] = sum_k A[k][m] * B[n][k]  (rows = K)
        const int ksteps = mn_major ? 4 : 4;
        for (int k = 0; k < ksteps; ++k) {
            const uint64_t adv_a = mn_major ? (uint64_t)((k * 2048) >> 4) : (uint64_t)((k * 32) >> 4);
            umma_bf16(tmem_base, adesc + adv_a, bdesc + (uint64_t)((k * 32) >> 4), idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + h * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) out[tid * 64 + h * 32 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

}  // namespace srlz

extern "C" int srlz_probe_desc_shift(float* out, int r0, int mode, int mn_major, void* stream) {
    using namespace srlz;
    const int smem = 256 * 128 + 64 * 128 + 1024 + 256;
    cudaFuncSetAttribute(desc_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    desc_shift_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(out, r0, mode, mn_major);
    return check_launch("desc_shift_probe");
}
