// Shared device/host helpers for libmvfusion (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <atomic>
#include <utility>
#include "mvfusion.h"

namespace mvf {

extern std::atomic<unsigned long long> g_launches;

inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Launch with programmatic stream serialization: the kernel may start while its predecessor in the stream is still running (as soon
// as every CTA of the predecessor has executed griddepcontrol.launch_dependents or exited).  Kernels launched this way must not rely on
// the predecessor's memory being flushed: they synchronise through device counters (unproject_tc.cu).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
#ifdef MVF_DEBUG_ENV
    if (e != cudaSuccess) fprintf(stderr, "libmvfusion: launch failed: %s\n", cudaGetErrorString(e));
#endif
    return e == cudaSuccess ? MVF_OK : MVF_ECUDA;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// A/B measurement switches (MVF_K1_*, MVF_TC_*) exist only in builds made with -DMVF_DEBUG_ENV (make DEBUG_ENV=1): the product
// library reads no environment variables, so an operand format or kernel variant is a pure function of the call's arguments.
#ifdef MVF_DEBUG_ENV
inline int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
#else
inline int env_int(const char*, int dflt) { return dflt; }
#endif

// ---- pinned fp32 arithmetic: every op individually rounded, no FMA contraction --------------
// (dot products left-to-right in ascending k).
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ float dot3_rn(float a0, float a1, float a2, float b0, float b1, float b2) {
    return add_rn(add_rn(mul_rn(a0, b0), mul_rn(a1, b1)), mul_rn(a2, b2));
}
__device__ __forceinline__ float dot4_rn(float a0, float a1, float a2, float a3,
                                         float b0, float b1, float b2, float b3) {
    return add_rn(add_rn(add_rn(mul_rn(a0, b0), mul_rn(a1, b1)), mul_rn(a2, b2)), mul_rn(a3, b3));
}

// [R^T | -(R^T t)] of a camera->world pose P (3x4 row-major) -> out (3x4 row-major)
// mrcnn/model_multi.py:137-143 / :279-281
__device__ __forceinline__ void inverse_pose(const float* P, float* out) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float r0 = P[0 * 4 + i], r1 = P[1 * 4 + i], r2 = P[2 * 4 + i];   // row i of R^T
        out[i * 4 + 0] = r0; out[i * 4 + 1] = r1; out[i * 4 + 2] = r2;
        out[i * 4 + 3] = -dot3_rn(r0, r1, r2, P[3], P[7], P[11]);
    }
}

// 3x4 . (x,y,z,1) row i, ascending k; the homogeneous 1 multiplies exactly.
__device__ __forceinline__ float affine_row(const float* M, int i, float x, float y, float z) {
    return add_rn(add_rn(add_rn(mul_rn(M[i * 4 + 0], x), mul_rn(M[i * 4 + 1], y)),
                         mul_rn(M[i * 4 + 2], z)), M[i * 4 + 3]);
}

__device__ __forceinline__ bool usable_coord(float v) {
    return (fabsf(v) < 1073741824.0f);      // finite and |v| < 2^30 (NaN compares false)
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcs4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float4 fma4(float w, float4 f, float4 a) {
    a.x = fmaf(w, f.x, a.x); a.y = fmaf(w, f.y, a.y); a.z = fmaf(w, f.z, a.z); a.w = fmaf(w, f.w, a.w);
    return a;
}
__device__ __forceinline__ float4 mul4(float w, float4 f) { return make_float4(w * f.x, w * f.y, w * f.z, w * f.w); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 max4(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }
__device__ __forceinline__ float4 relu4(float4 a) { return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f)); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 issues two fp32 FMAs per instruction slot) ---------
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 dup2(float w) { u64 d; asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(w)); return d; }   // folds into FFMA2's scalar operand
__device__ __forceinline__ u64 pack2(float a, float b) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float2 unpack2(u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float lo2(u64 v) { return unpack2(v).x; }
__device__ __forceinline__ float hi2(u64 v) { return unpack2(v).y; }
__device__ __forceinline__ ulonglong2 ldg2x2(const void* p) { return __ldg(reinterpret_cast<const ulonglong2*>(p)); }
__device__ __forceinline__ void stcs2x2(float* p, ulonglong2 v) {
    asm volatile("st.global.cs.v2.b64 [%0], {%1, %2};" :: "l"(p), "l"(v.x), "l"(v.y) : "memory");
}
// acc + w * t for four channels held as two packed pairs
__device__ __forceinline__ ulonglong2 fma2x2(float w, ulonglong2 t, ulonglong2 a) {
    const u64 ww = dup2(w);
    a.x = fma2(t.x, ww, a.x); a.y = fma2(t.y, ww, a.y);
    return a;
}
__device__ __forceinline__ float4 unpack4(ulonglong2 v) { const float2 a = unpack2(v.x), b = unpack2(v.y); return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ ulonglong2 pack4(float4 v) { ulonglong2 r; r.x = pack2(v.x, v.y); r.y = pack2(v.z, v.w); return r; }

// ---- fp16 operand split of the tensor-core convolutions (convlstm_tc.cu; also written directly by unproject.cu) ----
// scale 2^s with s = 14 - exponent(amax): a*2^s = a1 + a2 in two fp16 halves; *inv = 2^-s
__device__ __forceinline__ float pow2_scale(unsigned amax_bits, float* inv) {
    const float amax = __uint_as_float(amax_bits);
    int sa = 0;
    if (amax > 0.f && amax < 3.0e38f) { int e; frexpf(amax, &e); sa = min(max(14 - e, -100), 100); }
    *inv = ldexpf(1.0f, -sa);
    return ldexpf(1.0f, sa);
}
__device__ __forceinline__ void split_half4(float4 v, float scale, uint2* hi, uint2* lo) {
    const float x[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
    unsigned short h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half a1 = __float2half_rn(x[i]);
        h[i] = __half_as_ushort(a1);
        l[i] = __half_as_ushort(__float2half_rn(x[i] - __half2float(a1)));
    }
    *hi = make_uint2((unsigned)h[0] | ((unsigned)h[1] << 16), (unsigned)h[2] | ((unsigned)h[3] << 16));
    *lo = make_uint2((unsigned)l[0] | ((unsigned)l[1] << 16), (unsigned)l[2] | ((unsigned)l[3] << 16));
}

// ---- host: TF1 RangeOp<float> / LinSpaceOp<float> fill order ---------------------------------
// (third-party kernels restated; call sites mrcnn/model_multi.py:157-160, :267)
inline int tf1_range(double start_d, double limit_d, double delta_d, float* out, int cap) {
    const float start = (float)start_d, limit = (float)limit_d, delta = (float)delta_d;
    volatile float q = (limit - start) / delta;           // volatile: keep it a rounded float
    const int size = (int)ceilf(fabsf(q));
    if (size > cap || size < 0) return -1;
    volatile float val = start;
    for (int i = 0; i < size; ++i) { out[i] = val; val = val + delta; }
    return size;
}
inline void tf1_linspace(double start_d, double stop_d, int num, float* out) {
    const float start = (float)start_d, stop = (float)stop_d;
    if (num == 1) { out[0] = start; return; }
    volatile float step = (stop - start) / (float)(num - 1);
    for (int i = 0; i < num; ++i) { volatile float m = step * (float)i; out[i] = start + m; }
}

}  // namespace mvf
