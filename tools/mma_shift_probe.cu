// tcgen05.mma operand-alignment probe (test infrastructure): cycles per M=128 x N x K=16 kind::f16 MMA (K-major SWIZZLE_128B
// operands in shared memory, issued back to back, four K steps of one 128-byte row image as in the halo kernels) as a function
// of N and of a ROW SHIFT of the A or of the B start address (the halo kernels serve a conv tap by shifting the start row of a
// pixel-row image: is a start that is not a multiple of the 8-row swizzle atom slower, and for which operand?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_shift_probe tools/mma_shift_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../srl_zoo_b200/csrc/tc_common.cuh"

using namespace srlz;

__global__ void __launch_bounds__(128, 1) probe_kernel(long long* out, int N, int a_shift, int b_shift, int iters) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t a_base = base, b_base = base + 65536, bar = b_base + 98304;   // A: 512 rows, B: 768 rows of 128 B
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 65536 + 98304 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (65536 + 98304) / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x3f803f80u;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp == 0) {
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        uint64_t ad[4], bd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ad[k] = make_desc_sw128(a_base + a_shift * 128) + (uint64_t)((k * 32) >> 4);
            bd[k] = make_desc_sw128(b_base + b_shift * 128) + (uint64_t)((k * 32) >> 4);
        }
        const uint32_t d0 = tmem_base, d1 = tmem_base + (N <= 256 ? N : 0);
        const long long t0 = clock64();
        if (leader) {
            for (int i = 0; i < iters; ++i) {
                umma_bf16(d0, ad[0], bd[0], idesc, 1u); umma_bf16(d0, ad[1], bd[1], idesc, 1u);
                umma_bf16(d0, ad[2], bd[2], idesc, 1u); umma_bf16(d0, ad[3], bd[3], idesc, 1u);
                umma_bf16(d1, ad[0], bd[0], idesc, 1u); umma_bf16(d1, ad[1], bd[1], idesc, 1u);
                umma_bf16(d1, ad[2], bd[2], idesc, 1u); umma_bf16(d1, ad[3], bd[3], idesc, 1u);
            }
        }
        __syncwarp();
        if (leader) umma_commit(bar);
        __syncwarp();
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        if (leader) out[blockIdx.x] = t2 - t0;
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
    const int smem = 65536 + 98304 + 1024 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    long long h[148];
    const int iters = 256;  // x8 MMAs
    const int shifts[] = {0, 1, 2, 3, 4, 7, 8, 58, 59, 64, -1};
    printf("  N  a_shift b_shift | cycles per MMA (CTA 0, max over 148 CTAs)\n");
    for (int N = 64; N <= 256; N *= 2)
        for (int which = 0; which < 2; ++which)
            for (int si = 0; shifts[si] >= 0; ++si) {
                if (which == 1 && shifts[si] == 0) continue;
                const int as = which == 0 ? shifts[si] : 0, bs = which == 1 ? shifts[si] : 0;
                for (int rep = 0; rep < 2; ++rep) {
                    probe_kernel<<<148, 128, smem>>>(d, N, as, bs, iters);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < 148; ++i) if (h[i] > mx) mx = h[i];
                printf("%4d %6d %6d | %7.1f %7.1f\n", N, as, bs, (double)h[0] / (iters * 8), (double)mx / (iters * 8));
            }
    return 0;
}
