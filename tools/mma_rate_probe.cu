// tcgen05.mma issue-rate probe (test infrastructure, not on the product path): how many cycles does one
// M=128 x N x K=16 kind::f16 MMA take when issued back to back, as a function of N, of the operand layout
// (K-major / MN-major SWIZZLE_128B, A from shared memory or from TMEM) and of concurrent shared-memory
// store traffic from "producer" warps?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/mma_rate_probe tests/mma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../srl_zoo_b200/csrc/tc_common.cuh"

using namespace srlz;

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// mode 0: A, B K-major from smem | 1: A, B MN-major from smem | 2: A from TMEM, B K-major from smem
// writers: number of extra warps hammering shared memory with 16 B stores (0..8) ; wkind 0 = stores, 1 = loads
__global__ void __launch_bounds__(512, 1) probe_kernel(long long* out, int N, int mode, int writers, int wkind, int iters, int distinct, int nacc) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    // A: 4 stages x 32 KB (hi+lo planes) ; B: 64 KB ; scratch for writers: 32 KB ; bar
    const uint32_t a_base = base, b_base = base + 4 * 32768, scr = b_base + 65536, bar = scr + 32768;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 4 * 32768 + 65536 + 32768 + 64);
    volatile int* done = reinterpret_cast<volatile int*>(smem + 4 * 32768 + 65536 + 32768 + 128);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < (4 * 32768 + 65536) / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x3f803f80u + e * 2654435761u % 7u;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); *done = 0; }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp == 0) {
        const bool leader = elect_one();
        {
            uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
            if (mode == 1) idesc |= (1u << 15) | (1u << 16);
            // straight-line issue: descriptors precomputed, 8 MMAs per iteration, accumulators rotated at compile time
            uint64_t ad[4], bd[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mode == 1) {
                    ad[k] = desc_mn(a_base, 32768 / 2) + (uint64_t)((k * 2048) >> 4);
                    bd[k] = desc_mn(b_base, 16384) + (uint64_t)((k * 2048) >> 4);
                } else {
                    ad[k] = make_desc_sw128(a_base) + (uint64_t)((k * 32) >> 4);
                    bd[k] = make_desc_sw128(b_base) + (uint64_t)((k * 32) >> 4);
                }
            }
            const uint32_t d0 = tmem_base, d1 = tmem_base + (nacc > 1 ? N : 0), d2 = tmem_base + (nacc > 2 ? 2 * N : 0), d3 = tmem_base + (nacc > 2 ? 3 * N : (nacc > 1 ? N : 0));
            const long long t0 = clock64();
            if (leader) {
                for (int i = 0; i < iters / 2; ++i) {
                    if (mode == 2) {
                        umma_bf16_ts(d0, tmem_base + 448, bd[0], idesc, 1u); umma_bf16_ts(d1, tmem_base + 456, bd[1], idesc, 1u);
                        umma_bf16_ts(d2, tmem_base + 464, bd[2], idesc, 1u); umma_bf16_ts(d3, tmem_base + 472, bd[3], idesc, 1u);
                        umma_bf16_ts(d0, tmem_base + 448, bd[0], idesc, 1u); umma_bf16_ts(d1, tmem_base + 456, bd[1], idesc, 1u);
                        umma_bf16_ts(d2, tmem_base + 464, bd[2], idesc, 1u); umma_bf16_ts(d3, tmem_base + 472, bd[3], idesc, 1u);
                    } else {
                        umma_bf16(d0, ad[0], bd[0], idesc, 1u); umma_bf16(d1, ad[1], bd[1], idesc, 1u);
                        umma_bf16(d2, ad[2], bd[2], idesc, 1u); umma_bf16(d3, ad[3], bd[3], idesc, 1u);
                        umma_bf16(d0, ad[0], bd[0], idesc, 1u); umma_bf16(d1, ad[1], bd[1], idesc, 1u);
                        umma_bf16(d2, ad[2], bd[2], idesc, 1u); umma_bf16(d3, ad[3], bd[3], idesc, 1u);
                    }
                }
            }
            __syncwarp();
            if (leader) umma_commit(bar);
            __syncwarp();
            const long long t1 = clock64();
            mbar_wait(bar, 0);
            const long long t2 = clock64();
            if (leader) {
                out[blockIdx.x * 2] = t1 - t0;
                out[blockIdx.x * 2 + 1] = t2 - t0;
                *done = 1;
            }
        }
        __syncwarp();
    } else if (warp >= 8 && warp < 8 + writers) {
        // producer-like traffic: 16 B stores (or loads) into the scratch region, conflict-free pattern
        const uint32_t my = scr + (uint32_t)((tid - 256) * 16) % 32768u;
        uint32_t acc = 0;
        while (!*done) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t addr = scr + ((my - scr + j * 4096u) & 32767u);
                if (wkind == 0) {
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(acc) : "memory");
                } else {
                    uint32_t x, y, z, w;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr) : "memory");
                    acc += x + y + z + w;
                }
            }
            ++acc;
        }
        if (acc == 0xdeadbeefu) out[0] = acc;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
    const int smem = 4 * 32768 + 65536 + 32768 + 1024 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* d;
    cudaMalloc(&d, 148 * 2 * sizeof(long long));
    long long h[296];
    const int iters = 512;  // x4 MMAs
    printf("mode N nacc writers grid | issue cyc/MMA | complete cyc/MMA (CTA0, max over CTAs)\n");
    const int Ns[6] = {16, 32, 64, 128, 256, 0};
    for (int mode = 0; mode < 3; ++mode)
        for (int ni = 0; Ns[ni]; ++ni)
            for (int nacc = 1; nacc <= 4; nacc *= 2)
                for (int writers = 0; writers <= 8; writers += 8) {
                    const int N = Ns[ni], grid = 148;
                    if (N * nacc > 448 || (mode == 1 && N > 128)) continue;
                    if (mode == 1 && N < 64) continue;
                    for (int rep = 0; rep < 2; ++rep) {
                        probe_kernel<<<grid, 512, smem>>>(d, N, mode, writers, 0, iters, 1, nacc);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    }
                    cudaMemcpy(h, d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
                    long long mx = 0;
                    for (int i = 0; i < grid; ++i) if (h[2 * i + 1] > mx) mx = h[2 * i + 1];
                    printf("%d %3d %d %d %3d | %7.1f | %7.1f %7.1f\n", mode, N, nacc, writers, grid,
                           (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), (double)mx / (iters * 4));
                }
    return 0;
}
