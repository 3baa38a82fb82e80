// Epilogue store-path microbenchmark for K1T: each lane owns one voxel row and writes 32 consecutive floats
// (128 B) of it per step -- the register layout tcgen05.ld 32x32b.x32 produces.  Compared: (a) 8 x st.global.v4 per lane, (b) 4 x
// st.global.v8 (256-bit) per lane, (c) the staged path K1T uses (st.shared swizzled + fence + TMA tensor store of a {32,8,4,1} box).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I mulit_view_object_detection_b200/csrc -o tools/microbench5 tools/microbench5.cu && tools/microbench5
#include <cstdio>
#include "tc_ptx.cuh"
using namespace mvf;

constexpr int TILES = 256;     // tiles per CTA; tile = 128 voxel rows x 256 channels, as in K1T

template <int MODE>
__global__ void __launch_bounds__(128) k(float* out, const __grid_constant__ CUtensorMap tm, long long* cyc, int ntiles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float v[32];
    for (int i = 0; i < 32; ++i) v[i] = (float)(threadIdx.x + i);
    const long long t0 = clock64();
    uint32_t nstore = 0;
    for (int t = 0; t < TILES; ++t) {
        const int tile = (blockIdx.x + t * gridDim.x) % ntiles;
        const int tz = tile % 8, ty = (tile / 8) % 16, tx = (tile / 128) % 16, b = tile / 2048;
        const int dz = lane & 7, dy = lane >> 3, dx = warp;
        float* row = out + ((((size_t)b * 64 + tx * 4 + dx) * 64 + ty * 4 + dy) * 64 + tz * 8 + dz) * 256;
        for (int c = 0; c < 8; ++c) {
            for (int i = 0; i < 32; ++i) v[i] += 1.0f;
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(row + c * 32 + 4 * i), "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(row + c * 32 + 8 * i), "f"(v[8 * i]), "f"(v[8 * i + 1]), "f"(v[8 * i + 2]),
                                 "f"(v[8 * i + 3]), "f"(v[8 * i + 4]), "f"(v[8 * i + 5]), "f"(v[8 * i + 6]), "f"(v[8 * i + 7]) : "memory");
            } else {
                const uint32_t sb = base + warp * 8192u + (nstore & 1u) * 4096u;
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(sb + lane * 128u + (((uint32_t)i ^ (lane & 7u)) << 4)), "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3]) : "memory");
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { tma_store_5d(&tm, sb, c * 32, tz * 8, ty * 4, tx * 4 + dx, b); bulk_commit(); }
                ++nstore;
            }
        }
    }
    if (MODE == 2 && lane == 0) bulk_wait<0>();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
    const int B = 8, ntiles = B * 2048;
    float* g; cudaMalloc(&g, (size_t)B * 64 * 64 * 64 * 256 * 4);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    CUtensorMap tm;
    const cuuint64_t dims[5] = {256, 64, 64, 64, (cuuint64_t)B};
    const cuuint64_t str[4] = {1024, 64 * 1024, 64 * 64 * 1024, (cuuint64_t)64 * 64 * 64 * 1024};
    const cuuint32_t box[5] = {32, 8, 4, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    encode_tiled()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, g, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int SMEM = 4 * 8192 + 1024;
    const char* names[3] = {"8 x st.global.v4 per lane (rows 1 KB apart)", "4 x st.global.v8 per lane", "st.shared + fence + TMA store {32,8,4,1}"};
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 3; ++mode) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, 128, SMEM>>>(g, tm, cyc, ntiles);
            if (mode == 1) k<1><<<148, 128, SMEM>>>(g, tm, cyc, ntiles);
            if (mode == 2) k<2><<<148, 128, SMEM>>>(g, tm, cyc, ntiles);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            if (rep) printf("%-48s %7.0f SM-cycles per tile (128 KB), %6.1f B/clk/SM, %6.0f GB/s  (%s)\n", names[mode], avg / TILES, 131072.0 / (avg / TILES),
                            148.0 * TILES * 131072 / (ms * 1e6), cudaGetErrorString(e));
        }
    return 0;
}
