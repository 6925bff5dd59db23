// Probe: which (row, column) of a TMEM tile each thread receives from tcgen05.ld.16x256b (sm_100a).
// The tile is written with tcgen05.st.32x32b (thread t of warp w owns row 32w + t; value = row * 1000 + column).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_ld_probe tools/tmem_ld_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(int* out) {
    __shared__ uint32_t tptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tptr;
    const uint32_t mine = tb + ((uint32_t)(warp * 32) << 16);
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = (uint32_t)((warp * 32 + lane) * 1000 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(mine),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                 "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // 16x256b.x2: 16 rows x 16 columns -> 8 registers per thread; first from row offset 0, then row offset 16 of the warp's quarter
    for (int half = 0; half < 2; ++half) {
        uint32_t r[8];
        const uint32_t ta = tb + ((uint32_t)(warp * 32 + half * 16) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int i = 0; i < 8; ++i) out[((warp * 2 + half) * 32 + lane) * 8 + i] = (int)r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb));
}

int main() {
    int* d;
    cudaMalloc(&d, 4 * 2 * 32 * 8 * sizeof(int));
    cudaMemset(d, 0xff, 4 * 2 * 32 * 8 * sizeof(int));
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    static int h[4 * 2 * 32 * 8];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int warp = 0; warp < 2; ++warp)
        for (int half = 0; half < 2; ++half) {
            printf("warp %d, lane offset %d: thread -> (row.col) of r0..r7\n", warp, half * 16);
            for (int lane = 0; lane < 32; ++lane) {
                printf("  t%2d:", lane);
                for (int i = 0; i < 8; ++i) { const int x = h[((warp * 2 + half) * 32 + lane) * 8 + i]; printf(" %3d.%02d", x / 1000, x % 1000); }
                printf("\n");
            }
        }
    return 0;
}
