// TMA instruction-throughput microbenchmark for the K1T design: how many SM cycles does one
// cp.async.bulk.tensor load / store occupy the SM's TMA unit, as a function of the box shape (bytes, contiguous segments)?
// Sources are L2-resident (13 MB of features / a 32 MB window of the grid), the queue is deep (up to 32 loads in flight).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I mulit_view_object_detection_b200/csrc -o tools/microbench4 tools/microbench4.cu && tools/microbench4
#include <cstdio>
#include <cuda_fp16.h>
#include "tc_ptx.cuh"
using namespace mvf;

constexpr int NOPS = 4096, MAXDEPTH = 32, RING = 196608;

__device__ __forceinline__ void tma_load_nd(int nd, uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    if (nd == 5) tma_load_5d(dst, tm, bar, c0, c1, c2, c3, c4);
    else if (nd == 4) asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                              :: "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else tma_load_2d(dst, tm, bar, c0, c1);
}

// lane 0 of each of `nw` warps issues NOPS / nw loads of `bytes` each into its own `depth`-deep ring (depth * nw * bytes <= RING)
__global__ void __launch_bounds__(256) load_kernel(const __grid_constant__ CUtensorMap tm, int nd, uint32_t bytes, int depth, int m1, int m2, int m3, long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t base = ((smem_u32(smem_raw) + 1023u) & ~1023u) + (uint32_t)w * depth * bytes;
    __shared__ unsigned long long bars[8][MAXDEPTH];
    unsigned long long* bar = bars[w];
    if ((threadIdx.x & 31) == 0) {
        for (int i = 0; i < MAXDEPTH; ++i) mbar_init(smem_u32(&bar[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        const int dsh = 31 - __clz(depth), nops = NOPS / nw;
        const long long t0 = clock64();
        unsigned h = (blockIdx.x * 8 + w) * 7919u;
        for (int i = 0; i < nops; ++i) {
            const int s = i & (depth - 1);
            if (i >= depth) mbar_wait(smem_u32(&bar[s]), ((i >> dsh) - 1) & 1);
            mbar_expect_tx(smem_u32(&bar[s]), bytes);
            h = h * 1664525u + 1013904223u;
            const int a = (h >> 8) & m1, b = (h >> 16) & m2, c = (h >> 24) & m3;     // masks: the issuing thread must not be the bottleneck
            if (nd == 5) tma_load_nd(5, base + s * bytes, &tm, smem_u32(&bar[s]), 0, a, b, 0, c);
            else if (nd == 4) tma_load_nd(4, base + s * bytes, &tm, smem_u32(&bar[s]), 0, a, b, c * 8, 0);
            else tma_load_nd(2, base + s * bytes, &tm, smem_u32(&bar[s]), 0, a * 40 + b + c * 12800, 0, 0, 0);
        }
        for (int i = nops; i < nops + depth; ++i) mbar_wait(smem_u32(&bar[i & (depth - 1)]), ((i >> dsh) - 1) & 1);
        if (w == 0) cyc[blockIdx.x] = clock64() - t0;
    }
}

// one thread stores NOPS boxes from smem into an x < 8 window (32 MB) of a [64][64][64][256] fp32 grid
__global__ void __launch_bounds__(64) store_kernel(const __grid_constant__ CUtensorMap tm, int bc, int bz, int by, int bx, long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        unsigned h = blockIdx.x * 7919u;
        for (int i = 0; i < NOPS; ++i) {
            h = h * 1664525u + 1013904223u;
            const int c = ((h >> 4) & (256 / bc - 1)) * bc, z = ((h >> 8) & (64 / bz - 1)) * bz, y = ((h >> 14) & (64 / by - 1)) * by, x = ((h >> 20) & (8 / bx - 1)) * bx;
            bulk_wait_read<7>();
            tma_store_5d(&tm, base + (i & 7) * 16384u, c, z, y, x, 0);
            bulk_commit();
        }
        bulk_wait<0>();
        cyc[blockIdx.x] = clock64() - t0;
    }
}

static void report(const char* name, long long* cyc, double bytes) {
    long long h[148]; cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    printf("%-72s %7.1f SM-cycles per instruction  %6.1f B/clk/SM  (%s)\n", name, avg / NOPS, bytes / (avg / NOPS), cudaGetErrorString(e));
}

int main() {
    const int BV = 8, fh = 40, fw = 40, nblk = 4;                           // 13 MB of fp16 halves: L2-resident
    __half* f; cudaMalloc(&f, (size_t)2 * BV * nblk * fh * fw * 64 * 2); cudaMemset(f, 0, (size_t)2 * BV * nblk * fh * fw * 64 * 2);
    float* g; cudaMalloc(&g, (size_t)64 * 64 * 64 * 256 * 4);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const int SMEM = RING + 1024;
    cudaFuncSetAttribute(load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    char nm[128];
    for (int rep = 0; rep < 2; ++rep) {
        for (int nw : {1, 2, 4, 8}) for (int blk : {1, 4}) {   // 5-D {64 ch, px, rows, blk, 1}
            const int rows = blk == 1 ? 1 : 2, px = blk == 1 ? 1 : 4;
            CUtensorMap tm;
            const cuuint64_t dims[5] = {64, (cuuint64_t)fw, (cuuint64_t)fh, (cuuint64_t)nblk, (cuuint64_t)2 * BV};
            const cuuint64_t str[4] = {128, (cuuint64_t)fw * 128, (cuuint64_t)fh * fw * 128, (cuuint64_t)nblk * fh * fw * 128};
            const cuuint32_t box[5] = {64, (cuuint32_t)px, (cuuint32_t)rows, (cuuint32_t)blk, 1};
            encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, f, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            const uint32_t bytes = 128u * px * rows * blk;
            const int depth = 32 / nw;
            load_kernel<<<148, 32 * nw, SMEM>>>(tm, 5, bytes, depth, 31, 31, 2 * BV - 1, cyc);
            snprintf(nm, sizeof nm, "load 5-D {64ch,%dpx,%drows,%dblk,1} %5u B, %d issuing threads x depth %d", px, rows, blk, bytes, nw, depth);
            if (rep) report(nm, cyc, bytes);
        }
        for (int rows : {2, 4}) for (int px : {4, 8}) {   // hi and lo merged: 4-D {64 ch, px, rows, 8 (hl, blk)}
            CUtensorMap tm;
            const cuuint64_t dims[4] = {64, (cuuint64_t)fw, (cuuint64_t)fh, (cuuint64_t)2 * nblk * BV};
            const cuuint64_t str[3] = {128, (cuuint64_t)fw * 128, (cuuint64_t)fh * fw * 128};
            const cuuint32_t box[4] = {64, (cuuint32_t)px, (cuuint32_t)rows, 8};
            encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, f, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            const uint32_t bytes = 128u * px * rows * 8;
            if (bytes > 16384) continue;
            const int depth = bytes <= 4096 ? 32 : bytes <= 8192 ? 16 : 8;
            for (int nw : {1, 4}) {
            load_kernel<<<148, 32 * nw, SMEM>>>(tm, 4, bytes, depth / nw, 31, 31, BV - 1, cyc);
            snprintf(nm, sizeof nm, "load 4-D {64ch,%dpx,%drows,8 (hi/lo x blk)} %5u B, %d issuing threads x depth %d", px, rows, bytes, nw, depth / nw);
            if (rep) report(nm, cyc, bytes);
            }
        }
        for (int n : {4, 32, 128}) {                      // contiguous: 2-D {64 ch, n pixels}
            CUtensorMap tm;
            const cuuint64_t dims[2] = {64, (cuuint64_t)2 * BV * nblk * fh * fw};
            const cuuint64_t str[1] = {128};
            const cuuint32_t box[2] = {64, (cuuint32_t)n};
            encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, f, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            const uint32_t bytes = 128u * n;
            const int depth = bytes <= 4096 ? 32 : bytes <= 8192 ? 16 : 8;
            load_kernel<<<148, 32, SMEM>>>(tm, 2, bytes, depth, 31, 31, 2 * BV * nblk - 2, cyc);
            snprintf(nm, sizeof nm, "load 2-D {64ch,%d px} %5u B contiguous, depth %d", n, bytes, depth);
            if (rep) report(nm, cyc, bytes);
        }
        const int shapes[][5] = {{32, 8, 4, 1, 1}, {32, 8, 4, 4, 1}, {64, 8, 4, 2, 0}, {256, 8, 2, 1, 0}, {256, 8, 4, 1, 0}, {32, 1, 1, 1, 1}};
        for (auto& sh : shapes) {
            CUtensorMap tm;
            const cuuint64_t dims[5] = {256, 64, 64, 64, 1};
            const cuuint64_t str[4] = {1024, 64 * 1024, 64 * 64 * 1024, (cuuint64_t)64 * 64 * 64 * 1024};
            const cuuint32_t box[5] = {(cuuint32_t)sh[0], (cuuint32_t)sh[1], (cuuint32_t)sh[2], (cuuint32_t)sh[3], 1};
            CUresult r = encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, g, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           sh[4] ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { if (rep) printf("store box {%d,%d,%d,%d}: encode failed %d\n", sh[0], sh[1], sh[2], sh[3], (int)r); continue; }
            const double bytes = 4.0 * sh[0] * sh[1] * sh[2] * sh[3];
            store_kernel<<<148, 64, SMEM>>>(tm, sh[0], sh[1], sh[2], sh[3], cyc);
            snprintf(nm, sizeof nm, "store 5-D {%d f32, %dz, %dy, %dx, 1} %5.0f B, %3d rows x %4d B, L2 window", sh[0], sh[1], sh[2], sh[3], bytes, sh[1] * sh[2] * sh[3], 4 * sh[0]);
            if (rep) report(nm, cyc, bytes);
        }
    }
    return 0;
}
