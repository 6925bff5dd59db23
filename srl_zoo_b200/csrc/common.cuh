// Shared device/host helpers for libsrlz (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SRLZ_C 64          // channel width of every hidden conv layer (models/models.py:49-79)
#define SRLZ_MAX_PART 1184 // max per-CTA partial rows any kernel writes (8 CTAs x 148 SMs)

namespace srlz {

// thread-local last error (returned through srlz_last_error)
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> error code (0 ok)
int sm_count();

// ---- programmatic dependent launch ----
// A train step is ~130 dependent launches; between two of them the GPU idles for the grid-completion -> launch -> CTA-dispatch
// latency (a one-CTA kernel in the middle of the step costs ~12 us, profiles/r2_fused_finalize.md).  Every kernel of the library is
// launched with the programmatic-stream-serialization attribute and starts with pdl_enter() = `griddepcontrol.wait`, which
// blocks until the PREVIOUS kernel of the stream has completed and its writes are visible -- before the first global-memory
// access, so the ordering of a plain stream is kept.  The next grid is then launched while the last CTAs of this one retire
// (the implicit trigger at CTA exit) instead of after the grid has drained: measured -0.26 ms of 11.5 ms per step.
// (An explicit `griddepcontrol.launch_dependents` at the top of every kernel -- the next grid resident and waiting from the
// start -- measured +0.5 ms instead, before the wait (SRLZ_PDL=1) as well as after it.)  Without the attribute the instruction
// is a no-op.
#ifndef SRLZ_PDL
#define SRLZ_PDL 2   // 0: plain launches | 2: attribute + wait (product) | 1: + early trigger (experiment)
#endif
__device__ __forceinline__ void pdl_enter() {
#if SRLZ_PDL == 1
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
#if SRLZ_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = SRLZ_PDL ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through check_launch()
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// 256-bit read-only load (LDG.E.256, sm_100): p must be 32-byte aligned.  The producers' "one thread owns 128 contiguous
// bytes" pattern touches 32 different lines per warp instruction; half as many instructions = half as many L1 requests.
__device__ __forceinline__ void ldg8(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

__device__ __forceinline__ float4 bn_relu4(float4 v, float4 sc, float4 sh) {
    v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
    v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
    v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
    v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
    return v;
}

// geometry of a "gather convolution" over 64-channel NHWC tensors.
//   big side  : spatial (BH,BW)   small side : spatial (SH,SW)
//   relation  : by = sy*stride - pad + ky ,  bx = sx*stride - pad + kx
// direct conv fwd      : out = small, gathered = big            (transposed = 0)
// transposed conv fwd  : out = big,   gathered = small          (transposed = 1)
// dgrad of direct conv : out = big,   gathered = small (dy)     (transposed = 1)
// dgrad of transposed  : out = small, gathered = big   (dy)     (transposed = 0)
struct ConvGeom {
    int B;
    int BH, BW;   // big side
    int SH, SW;   // small side
    int KH, KW, stride, pad;
};

}  // namespace srlz
