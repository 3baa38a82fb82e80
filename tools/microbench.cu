// Pipe-rate microbenchmarks that size the K1 design (run on one B200; prints SM-cycles per warp instruction).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int ITERS = 2048;
__device__ __forceinline__ float4 lds128(const float* p) { float4 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ float4 ldgca(const float4* p) { float4 v;
    asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ float2 lds64(const float* p) { float2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ float lds32(const float* p) { float v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
// mode 0: FFMA x16 independent chains; 1: FFMA2 x16; 2: FFMA2 x16 + 4 FMUL + 2 FADD; 3: LDS.128 broadcast x8;
// 4: LDS.128 distinct x8; 5: LDS.32 broadcast x8; 6: LDS.64 broadcast x8; 7: LDG.128 L1-hit x8 (512B/warp coalesced)
// 8: FFMA2 x16 + 1 LDS.128 broadcast ; 9: FFMA2 x16 + 2 LDG.128 ; 10: SHFL x8 ; 11: FFMA2 x16 + FMUL2/FADD2 x3
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, const float4* __restrict__ g, long long* cyc, float s) {
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)i * s;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a[16]; u64 p[16];
    for (int i = 0; i < 16; ++i) { a[i] = s * (i + lane); p[i] = (u64)__float_as_uint(a[i]) | ((u64)__float_as_uint(a[i]) << 32); }
    float w0 = s, w1 = s + 1.f, w2 = 0.5f, w3 = 0.25f;
    float4 acc4 = make_float4(0, 0, 0, 0);
    const u64 t = ((u64)__float_as_uint(1.0001f) << 32) | __float_as_uint(0.9999f);
    const float4* gp = g + lane + warp * 32;
    long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
            #pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w0, w1);
        } else if (MODE == 1 || MODE == 2 || MODE == 8 || MODE == 9 || MODE == 11) {
            if (MODE == 2) { float bx = 1.f - w2, by = 1.f - w3; w0 = w2 * w3; w1 = w2 * by; float wc = bx * w3, wd = bx * by; w2 = wc + 0.1f; w3 = wd + 0.1f; }
            if (MODE == 11) {
                u64 one = ((u64)__float_as_uint(1.0f) << 32) | __float_as_uint(1.0f);
                u64 axy = ((u64)__float_as_uint(w3) << 32) | __float_as_uint(w2), bxy, wab, wcd;
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(bxy) : "l"(one), "l"(axy));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(wab) : "l"(axy), "l"(bxy));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(wcd) : "l"(bxy), "l"(wab));
                w0 = __uint_as_float((unsigned)wab); w1 = __uint_as_float((unsigned)(wab >> 32)); w2 = __uint_as_float((unsigned)wcd) ; w3 = __uint_as_float((unsigned)(wcd >> 32));
            }
            if (MODE == 8) { float4 v = lds128(&sm[(it & 255) * 4]); w0 = v.x; w1 = v.y; acc4.z += v.z + v.w; }
            if (MODE == 9) { float4 v = ldgca(gp + ((it & 3) * 2048)); float4 v2 = ldgca(gp + 1024 + ((it & 3) * 2048)); w0 = v.x + v2.x; w1 = v.y + v2.y; acc4.z += v.z + v2.w; }
            const u64 ww0 = (u64)__float_as_uint(w0) | ((u64)__float_as_uint(w0) << 32);
            const u64 ww1 = (u64)__float_as_uint(w1) | ((u64)__float_as_uint(w1) << 32);
            #pragma unroll
            for (int i = 0; i < 16; ++i) p[i] = fma2(p[i], (i & 1) ? ww1 : ww0, t);
        } else if (MODE == 3 || MODE == 4) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = (MODE == 3) ? ((it + i) & 255) * 4 : (((it + i) & 7) * 128 + lane * 4);
                float4 v = lds128(&sm[idx]);
                acc4.x += v.x; acc4.y += v.y; acc4.z += v.z; acc4.w += v.w;
            }
        } else if (MODE == 5) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) acc4.x += lds32(&sm[(it + i) & 1023]);
        } else if (MODE == 6) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) { float2 v = lds64(&sm[((it + i) & 511) * 2]); acc4.x += v.x; acc4.y += v.y; }
        } else if (MODE == 7) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) { float4 v = ldgca(gp + i * 1024); acc4.x += v.x; acc4.y += v.y; acc4.z += v.z; acc4.w += v.w; }
        } else if (MODE == 10) {
            #pragma unroll
            for (int i = 0; i < 8; ++i) a[i] += __shfl_sync(0xffffffffu, a[i], (it + i) & 31);
        }
    }
    long long t1 = clock64();
    float r = acc4.x + acc4.y + acc4.z + acc4.w;
    for (int i = 0; i < 16; ++i) r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, int per_iter, float* out, float4* g, long long* cyc) {
    for (int warps : {4, 8, 16, 32}) {
        k<MODE><<<148, warps * 32>>>(out, g, cyc, 1e-6f);
        cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, g, cyc, 1e-6f);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-44s warps/SM=%2d  SM-cycles per warp-instr = %.3f  (%s)\n", name, warps, avg / ((double)ITERS * per_iter * warps), cudaGetErrorString(e));
    }
}
int main() {
    float* out; float4* g; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&g, 64 << 20); cudaMemset(g, 0, 64 << 20); cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA x16", 16, out, g, cyc);
    run<1>("FFMA2 x16", 16, out, g, cyc);
    run<2>("FFMA2 x16 + 4FMUL+2FADD (per 16 FFMA2)", 16, out, g, cyc);
    run<11>("FFMA2 x16 + 3 packed weight ops (per 16)", 16, out, g, cyc);
    run<3>("LDS.128 broadcast", 8, out, g, cyc);
    run<4>("LDS.128 distinct conflict-free", 8, out, g, cyc);
    run<5>("LDS.32 broadcast", 8, out, g, cyc);
    run<6>("LDS.64 broadcast", 8, out, g, cyc);
    run<7>("LDG.128 coalesced L1-hit", 8, out, g, cyc);
    run<8>("FFMA2 x16 + 1 LDS.128 bcast (per 16 FFMA2)", 16, out, g, cyc);
    run<9>("FFMA2 x16 + 2 LDG.128 (per 16 FFMA2)", 16, out, g, cyc);
    run<10>("SHFL.IDX", 8, out, g, cyc);
    return 0;
}
