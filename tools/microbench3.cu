// K1 decision microbenchmark (i) of VERDICT r1: register-return rate of tcgen05.ld (TMEM -> registers) against the
// 128 B/clk LSU write-back path that bounds unproject_slot_kernel (LDG.128 / LDS.128: profiles/microbench_r1.txt).
// Decision rule: stage tap footprints in TMEM and read them with LDTM only if TMEM -> RF sustains clearly MORE than 128 B/clk/SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench3 tools/microbench3.cu && tools/microbench3
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 1024;

template <int NREG> __device__ __forceinline__ void ldtm(unsigned taddr, unsigned* r);
template <> __device__ __forceinline__ void ldtm<8>(unsigned taddr, unsigned* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
template <> __device__ __forceinline__ void ldtm<16>(unsigned taddr, unsigned* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
template <> __device__ __forceinline__ void ldtm<32>(unsigned taddr, unsigned* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
}

// MODE 0: NREG-register tcgen05.ld, 4 KB per warp in flight per wait;  MODE 1: coalesced ld.global.v4 (L1 hit), 8 in flight;  MODE 2: ld.shared.v4
template <int MODE, int NREG>
__global__ void __launch_bounds__(512) k(unsigned* out, const uint4* __restrict__ g, long long* cyc) {
    __shared__ unsigned tmem_slot;
    __shared__ __align__(16) uint4 sm[2048];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i, i, i, i);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((unsigned)__cvta_generic_to_shared(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = tmem_slot + ((unsigned)((warp & 3) * 32) << 16);
    unsigned acc = 0;
    unsigned r[32];
    for (int i = 0; i < 32; ++i) r[i] = 0;
    const uint4* gp = g + warp * 32 + lane;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
            // 32 registers (4 KB per warp) in flight per wait, as 4 x8, 2 x16 or 1 x32 loads at different columns
#pragma unroll
            for (int j = 0; j < 32 / NREG; ++j)
                ldtm<NREG>(tbase + (unsigned)(((it * (32 / NREG) + j) * NREG) & 511), r + j * NREG);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 32; ++q) acc ^= r[q];
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 v; const uint4* p = gp + ((it * 8 + j) & 31) * 512;
                asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 v; const unsigned a = (unsigned)__cvta_generic_to_shared(&sm[(((it * 8 + j) & 63) * 32 + lane)]);
                asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_slot) : "memory");
}

template <int MODE, int NREG> void run(const char* name, int per_iter, int bytes_per_instr, unsigned* out, uint4* g, long long* cyc) {
    for (int warps : {4, 8, 16}) {
        k<MODE, NREG><<<148, warps * 32>>>(out, g, cyc);
        cudaDeviceSynchronize();
        k<MODE, NREG><<<148, warps * 32>>>(out, g, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        const double per = avg / ((double)ITERS * per_iter * warps);
        printf("%-34s warps/SM=%2d  SM-cycles per warp-instr = %7.3f  -> %6.1f B/clk/SM into registers  (%s)\n", name, warps, per,
               bytes_per_instr / per, cudaGetErrorString(e));
    }
}
int main() {
    unsigned* out; uint4* g; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&g, 64 << 20); cudaMemset(g, 0, 64 << 20); cudaMalloc(&cyc, 148 * 8);
    run<0, 8>("tcgen05.ld 32x32b.x8  (1 KB/warp)", 4, 1024, out, g, cyc);
    run<0, 16>("tcgen05.ld 32x32b.x16 (2 KB/warp)", 2, 2048, out, g, cyc);
    run<0, 32>("tcgen05.ld 32x32b.x32 (4 KB/warp)", 1, 4096, out, g, cyc);
    run<1, 8>("ld.global.v4 coalesced, L1 hit", 8, 512, out, g, cyc);
    run<2, 8>("ld.shared.v4 conflict-free", 8, 512, out, g, cyc);
    return 0;
}
