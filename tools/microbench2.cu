// L1 load-path microbenchmarks: how many SM-cycles does a coalesced 512 B warp load cost when it hits L1?
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
__device__ __forceinline__ float4 ldnc(const float4* p) { float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }
// MODE 0: 8 independent ld.global.nc.v4 per iteration, addresses walk a per-SM footprint of FOOT bytes (L1 resident if small)
// MODE 1: same but every lane reads the SAME 16 B (broadcast global load)
// MODE 2: 32-bit coalesced loads (128 B per warp)
template <int MODE>
__global__ void __launch_bounds__(1024) k(float* out, const float4* __restrict__ g, long long* cyc, int foot_vec4) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4* base = g + (size_t)blockIdx.x * foot_vec4;
    float4 acc = make_float4(0, 0, 0, 0);
    unsigned pos = warp * 32 * 8;
    // warm the footprint
    for (int i = threadIdx.x; i < foot_vec4; i += blockDim.x) { float4 v = ldnc(base + i); acc.x += v.x; }
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i) {
            unsigned o = (pos + i * 32) & (unsigned)(foot_vec4 - 1);
            float4 v;
            if (MODE == 0) v = ldnc(base + o + lane);
            else if (MODE == 1) v = ldnc(base + o);
            else { float f; const float* pf = reinterpret_cast<const float*>(base + o) + lane; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(f) : "l"(pf)); v = make_float4(f, 0, 0, 0); }
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        pos += 8 * 32 * 7;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int foot_kb, float* out, float4* g, long long* cyc) {
    for (int warps : {8, 16, 32}) {
        const int fv = foot_kb * 1024 / 16;
        k<MODE><<<148, warps * 32>>>(out, g, cyc, fv); cudaDeviceSynchronize();
        k<MODE><<<148, warps * 32>>>(out, g, cyc, fv);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("%-40s footprint/SM=%4d KB warps/SM=%2d  SM-cycles per warp load = %.3f (%s)\n", name, foot_kb, warps, avg / ((double)ITERS * 8 * warps), cudaGetErrorString(e));
    }
}
int main() {
    float* out; float4* g; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&g, 256 << 20); cudaMemset(g, 0, 256 << 20); cudaMalloc(&cyc, 148 * 8);
    for (int kb : {32, 128, 1024}) {
        run<0>("ld.global.nc.v4 coalesced (512 B/warp)", kb, out, g, cyc);
    }
    run<1>("ld.global.nc.v4 broadcast (16 B/warp)", 32, out, g, cyc);
    run<2>("ld.global.nc.f32 coalesced (128 B/warp)", 32, out, g, cyc);
    return 0;
}
