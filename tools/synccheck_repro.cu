// Reproducer for the compute-sanitizer synccheck report recorded in profiles/r2_sanitizer.md ("Barrier error detected. Missing
// init." on a barrier that thread 0 initialised before a __syncthreads).  Same barrier layout, thread roles and wait / arrive order
// as gconv64_halo_kernel at the 6x6 -> 13x13 geometry (21 mbarriers behind 2 x 31 KB image planes and 144 KB of weights; eight
// producer warps wait on row_free(j) with the pre-first-phase parity and arrive on row_full(j), row by row; one consumer warp waits
// on row_full and releases row_free), without any tensor-core work.  Features are switched on one at a time:
//   bit 0: zero the image planes before the __syncthreads          bit 1: 9 x 16 KB bulk copies (cp.async.bulk) onto a tx barrier
//   bit 2: tcgen05.alloc / dealloc by warp 4                        bit 3: release row_free through tcgen05.commit instead of arrive
// build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/synccheck_repro tools/synccheck_repro.cu
// run:   for m in 0 1 2 4 8 15; do compute-sanitizer --tool synccheck tools/synccheck_repro $m; done
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int PLANE = 248 * 128, W_BYTES = 9 * 2 * 64 * 128, MAXNR = 8, THREADS = 512;
constexpr int SMEM_BYTES = 2 * PLANE + W_BYTES + 1024 + 256 + 6 * 64 * 4 + 4 * 128 * 4 + 4 * 4096;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni WAIT_DONE;\n\t"
                 "bra.uni WAIT_LOOP;\n\tWAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1) repro(const unsigned char* __restrict__ w, int* out, int mode, int tiles, int HW, int NR) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    unsigned char* smem = smem_raw + (base - raw);
    const uint32_t wsm = base + 2 * PLANE, bars = wsm + W_BYTES;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + 2 * PLANE + W_BYTES + 192);
    auto row_full = [&](int j) { return bars + 8u * j; };
    auto row_free = [&](int j) { return bars + 8u * (MAXNR + j); };
    auto tfull_bar = [&](int i) { return bars + 8u * (2 * MAXNR + i); };
    auto tempty_bar = [&](int i) { return bars + 8u * (2 * MAXNR + 2 + i); };
    const uint32_t wfull = bars + 8u * (2 * MAXNR + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int j = 0; j < MAXNR; ++j) { mbar_init(row_full(j), 8); mbar_init(row_free(j), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull_bar(i), 1); mbar_init(tempty_bar(i), 4); }
        mbar_init(wfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (mode & 1)
        for (int e = tid; e < 2 * PLANE / 16; e += THREADS) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if ((mode & 4) && warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if ((mode & 2) && tid == 0) {
        mbar_arrive_expect_tx(wfull, 9 * 16384);
        for (int t = 0; t < 9; ++t) bulk_g2s(wsm + t * 16384, w + (size_t)t * 16384, 16384, wfull);
    }
    if (warp >= 8) {   // producers
        const int pw = warp - 8, items_per_row = 2 * HW, nitems = NR * items_per_row;
        for (int it = 0; it < tiles; ++it) {
            const int fph = it & 1;
            int arrived = 0;
            auto pass_rows = [&](int upto) {
                for (; arrived < upto; ++arrived) {
                    mbar_wait(row_free(arrived), fph ^ 1);
                    if (lane == 0) mbar_arrive(row_full(arrived));
                }
            };
            for (int k = 0; k < 2; ++k) {
                const int lo_i = 256 * k + 32 * pw;
                if (lo_i >= nitems) break;
                const int lo_row = lo_i / items_per_row;
                int hi_row = (lo_i + 31) / items_per_row;
                if (hi_row >= NR) hi_row = NR - 1;
                pass_rows(lo_row);
                for (int row = lo_row; row <= hi_row; ++row) {
                    mbar_wait(row_free(row), fph ^ 1);
                    smem[(row * HW + (lane >> 1)) * 128 + (lane & 1) * 64] = (unsigned char)it;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(row_full(row));
                    arrived = row + 1;
                }
            }
            pass_rows(NR);
        }
    } else if (warp == 4) {   // consumer (the MMA issuer's waits and releases)
        if (mode & 2) mbar_wait(wfull, 0);
        int acc = 0;
        for (int it = 0; it < tiles; ++it) {
            const int fph = it & 1;
            for (int j = 0; j < NR; ++j) {
                mbar_wait(row_full(j), fph);
                acc += smem[j * HW * 128];
                if (mode & 8) {
                    if (lane == 0)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(row_free(j)) : "memory");
                } else if (lane == 0) {
                    mbar_arrive(row_free(j));
                }
                __syncwarp();
            }
        }
        if (lane == 0) out[blockIdx.x] = acc;
    }
    __syncthreads();
    if ((mode & 4) && warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_ptr_smem), "r"(512) : "memory");
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0, tiles = argc > 2 ? atoi(argv[2]) : 1;
    const int HW = argc > 3 ? atoi(argv[3]) : 8, NR = argc > 4 ? atoi(argv[4]) : 8;
    unsigned char* w; int* out;
    cudaMalloc(&w, 9 * 16384); cudaMemset(w, 1, 9 * 16384); cudaMalloc(&out, 2 * sizeof(int));
    cudaFuncSetAttribute(repro, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    repro<<<2, THREADS, SMEM_BYTES>>>(w, out, mode, tiles, HW, NR);
    cudaError_t e = cudaDeviceSynchronize();
    int h[2] = {-1, -1};
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("mode %d tiles %d HW %d NR %d: %s (out %d %d)\n", mode, tiles, HW, NR, cudaGetErrorString(e), h[0], h[1]);
    return e != cudaSuccess;
}
