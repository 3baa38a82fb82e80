// K1  unproject_fuse: unproj_feat (+ fused view reduction / BN / ReLU) for sm_100a.
//
// Replaces mrcnn/model_multi.py:130-228 (unproj_feat) and :401-404 (grid_reas 'add'), plus the
// world-frame notebook variant Notebook/projection.py:47-151.  The per-view grids
// [B,V,X,Y,Z,C] are never materialised unless mode == MVF_FUSE_NONE.
//
// Two kernels live here.  The production kernel is unproject_run_kernel (further down): a warp
// owns a z-run of one voxel column and caches the bilinear 2x2 patch in registers.  The first
// generation brick kernel below is kept for A/B measurement only (MVF_K1_VARIANT=0).
//
// Brick kernel mapping: one CTA = one 4x4x16 brick of voxels of one scene; a warp owns 32 voxels of the
// brick.  Lanes first work as 32 independent (voxel, view) coordinate units -- voxel->pixel
// projection, floor, the four weights and the tap-validity bits are computed ONCE per pair in
// registers with individually rounded fp32 ops (bit-exact against the oracle) -- and then the
// warp walks the 32 pairs: every pair's parameters are broadcast with shuffles and the 32 lanes
// become 32 float4 channel slots, so every tap is one coalesced 128-bit read-only load per lane
// (512 B per warp instruction) and every voxel one coalesced streaming 128-bit store.
// The view reduction lives in the accumulator registers.
#include "mvf_common.cuh"
#include <stdlib.h>

namespace mvf {

constexpr int BRICK_X = 4, BRICK_Y = 4, BRICK_Z = 16;
constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;
static_assert(BRICK_X * BRICK_Y * BRICK_Z == K1_WARPS * 32, "one voxel per (warp, lane)");

struct UnprojParams {
    const float* feats; const float* Rcam; const float* Rmain; const float* Kmat;
    const float* bn_scale; const float* bn_shift;
    float* out; int32_t* out_idx; uint8_t* out_valid; float* out_grid_pos;
    int B, V, fh, fw, C, X, Y, Z, x_begin, Xs;
    int mode, flags;
    float sx, sy, inv_v, grid_dist;
    float gx[MVF_MAX_DIM], gy[MVF_MAX_DIM], gz[MVF_MAX_DIM];
};

template <int CPL, int MODE>
__global__ void __launch_bounds__(K1_THREADS)
unproject_fuse_kernel(const __grid_constant__ UnprojParams p) {
    __shared__ float sKR[MVF_MAX_VIEWS][12];
    __shared__ float sOff[3];

    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const bool world = (p.flags & MVF_FLAG_WORLD_GRID) != 0;

    // ---- prologue: KR_v = (K . [R_v^T | -R_v^T t_v]) . [[R_0|t_0],[0 0 0 1]]   (:137-147, :175-180)
    if (tid < p.V) {
        const float* P = p.Rcam + ((size_t)b * p.V + tid) * 12;
        const float* K = p.Kmat + (size_t)b * 9;
        const float* P0 = p.Rmain ? p.Rmain + (size_t)b * 12 : p.Rcam + (size_t)b * p.V * 12;
        float Rinv[12], M[12];
        inverse_pose(P, Rinv);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                M[i * 4 + j] = dot3_rn(K[i * 3 + 0], K[i * 3 + 1], K[i * 3 + 2],
                                       Rinv[0 * 4 + j], Rinv[1 * 4 + j], Rinv[2 * 4 + j]);
        if (world) {
#pragma unroll
            for (int e = 0; e < 12; ++e) sKR[tid][e] = M[e];
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float t3 = (j == 3) ? 1.0f : 0.0f;            // last row of T0
                    sKR[tid][i * 4 + j] = dot4_rn(M[i * 4 + 0], M[i * 4 + 1], M[i * 4 + 2], M[i * 4 + 3],
                                                  P0[0 * 4 + j], P0[1 * 4 + j], P0[2 * 4 + j], t3);
                }
        }
    }
    if (tid == 32 && world) {
        // grid_position = [R0|t0] . (0,0,grid_dist,1)   (Notebook/projection.py:86-91)
        const float* P0 = p.Rmain ? p.Rmain + (size_t)b * 12 : p.Rcam + (size_t)b * p.V * 12;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float v = dot4_rn(P0[i * 4 + 0], P0[i * 4 + 1], P0[i * 4 + 2], P0[i * 4 + 3],
                                    0.0f, 0.0f, p.grid_dist, 1.0f);
            sOff[i] = v;
            if (p.out_grid_pos && blockIdx.x == 0) p.out_grid_pos[b * 3 + i] = v;
        }
    }
    __syncthreads();

    // ---- brick / voxel assignment
    const int tiles_z = (p.Z + BRICK_Z - 1) / BRICK_Z;
    const int tiles_y = (p.Y + BRICK_Y - 1) / BRICK_Y;
    int t = blockIdx.x;
    const int tz = t % tiles_z; t /= tiles_z;
    const int ty = t % tiles_y; t /= tiles_y;
    const int tx = t;
    const int warp = tid >> 5, lane = tid & 31;
    const unsigned FULL = 0xffffffffu;

    const int C = p.C, V = p.V;
    const int C4 = C >> 2;
    const size_t view_stride = (size_t)p.fh * p.fw * C;
    const float* feats_b = p.feats + (size_t)b * V * view_stride;
    const int row_stride = p.fw * C;

    float4 acc[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[k] = zero4();

    const int npairs = 32 * V;
    for (int q0 = 0; q0 < npairs; q0 += 32) {
        // ---- phase 1: lane = one (voxel, view) pair
        const int q = q0 + lane;
        const int lv = q / V;                    // voxel slot 0..31 of this warp
        const int v = q - lv * V;
        const int l = warp * 32 + lv;            // voxel index inside the brick (z fastest)
        const int iz = tz * BRICK_Z + (l % BRICK_Z);
        const int iy = ty * BRICK_Y + ((l / BRICK_Z) % BRICK_Y);
        const int ixs = tx * BRICK_X + (l / (BRICK_Z * BRICK_Y));      // index inside the slab
        const bool in_grid = (iz < p.Z) && (iy < p.Y) && (ixs < p.Xs);
        int my_bits = 0, my_off = 0;
        float wa = 0.f, wb = 0.f, wc = 0.f, wd = 0.f;
        if (in_grid) {
            float x = p.gx[p.x_begin + ixs], y = p.gy[iy], z = p.gz[iz];
            if (world) { x = add_rn(x, sOff[0]); y = add_rn(y, sOff[1]); z = add_rn(z, sOff[2]); }
            const float* KR = sKR[v];
            const float px = affine_row(KR, 0, x, y, z);
            const float py = affine_row(KR, 1, x, y, z);
            const float pz = affine_row(KR, 2, x, y, z);
            const float u = mul_rn(div_rn(px, pz), p.sx);               // :187
            const float w = mul_rn(div_rn(py, pz), p.sy);               // :188
            int x0 = INT32_MIN, y0 = INT32_MIN;
            if (usable_coord(u) && usable_coord(w)) {
                const float x0f = floorf(u), y0f = floorf(w);            // :192-195
                x0 = (int)x0f; y0 = (int)y0f;
                const float x1f = (float)(x0 + 1), y1f = (float)(y0 + 1);
                const float ax = sub_rn(x1f, u), bx = sub_rn(u, x0f);
                const float ay = sub_rn(y1f, w), by = sub_rn(w, y0f);
                wa = mul_rn(ax, ay); wb = mul_rn(ax, by); wc = mul_rn(bx, ay); wd = mul_rn(bx, by);   // :214-217
                const bool inx0 = (x0 >= 0) && (x0 < p.fw), inx1 = (x0 + 1 >= 0) && (x0 + 1 < p.fw);
                const bool iny0 = (y0 >= 0) && (y0 < p.fh), iny1 = (y0 + 1 >= 0) && (y0 + 1 < p.fh);
                my_bits = (int)(iny0 && inx0) | ((int)(iny1 && inx0) << 1) | ((int)(iny0 && inx1) << 2) |
                          ((int)(iny1 && inx1) << 3);
                if (my_bits) my_off = (y0 * p.fw + x0) * C;              // may be negative: only valid taps are read
            }
            const size_t vox = (((size_t)b * V + v) * p.Xs + ixs) * p.Y * p.Z + (size_t)iy * p.Z + iz;
            if (p.out_idx) { p.out_idx[vox * 2 + 0] = y0; p.out_idx[vox * 2 + 1] = x0; }
            if (p.out_valid) p.out_valid[vox] = (uint8_t)my_bits;
            my_bits |= 16;                                               // bit4: voxel is inside the grid
        }

        // ---- phase 2: lanes = float4 channel slots; walk the 32 pairs of this chunk
        const int jmax = min(32, npairs - q0);
        for (int j = 0; j < jmax; ++j) {
            const int bits = __shfl_sync(FULL, my_bits, j);
            const int qq = q0 + j;
            const int jlv = qq / V;
            const int jv = qq - jlv * V;
            if (!(bits & 16)) continue;                                  // warp-uniform
            float4 val[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) val[k] = zero4();
            if (bits & 15) {
                const int off = __shfl_sync(FULL, my_off, j);
                const float fa = __shfl_sync(FULL, wa, j), fb = __shfl_sync(FULL, wb, j);
                const float fc = __shfl_sync(FULL, wc, j), fd = __shfl_sync(FULL, wd, j);
                const float* base = feats_b + (size_t)jv * view_stride + off;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const int c4 = lane + 32 * k;
                    if (c4 < C4) {
                        const float* pc = base + 4 * c4;
                        if (bits & 1) val[k] = fma4(fa, ldg4(pc), val[k]);
                        if (bits & 2) val[k] = fma4(fb, ldg4(pc + row_stride), val[k]);
                        if (bits & 4) val[k] = fma4(fc, ldg4(pc + C), val[k]);
                        if (bits & 8) val[k] = fma4(fd, ldg4(pc + row_stride + C), val[k]);
                    }
                }
            }
            if (p.flags & MVF_FLAG_RELU_IN) {
#pragma unroll
                for (int k = 0; k < CPL; ++k) val[k] = relu4(val[k]);
            }
            // voxel coordinates of pair j (warp-uniform)
            const int jl = warp * 32 + jlv;
            const int jz = tz * BRICK_Z + (jl % BRICK_Z);
            const int jy = ty * BRICK_Y + ((jl / BRICK_Z) % BRICK_Y);
            const int jx = tx * BRICK_X + (jl / (BRICK_Z * BRICK_Y));
            if (MODE == MVF_FUSE_NONE) {
                float* o = p.out + ((((size_t)b * V + jv) * p.Xs + jx) * p.Y * p.Z + (size_t)jy * p.Z + jz) * C;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const int c4 = lane + 32 * k;
                    if (c4 < C4) stcs4(o + 4 * c4, val[k]);
                }
                continue;
            }
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                if (MODE == MVF_FUSE_MAX) acc[k] = (jv == 0) ? val[k] : max4(acc[k], val[k]);
                else acc[k] = (jv == 0) ? val[k] : add4(acc[k], val[k]);
            }
            if (jv == V - 1) {
                float* o = p.out + ((((size_t)b * p.Xs + jx) * p.Y + jy) * p.Z + jz) * C;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    const int c4 = lane + 32 * k;
                    if (c4 < C4) {
                        float4 r = acc[k];
                        if (MODE == MVF_FUSE_MEAN) r = mul4(p.inv_v, r);
                        if (p.bn_scale) {
                            const float4 s = ldg4(p.bn_scale + 4 * c4), h = ldg4(p.bn_shift + 4 * c4);
                            r = make_float4(fmaf(r.x, s.x, h.x), fmaf(r.y, s.y, h.y), fmaf(r.z, s.z, h.z), fmaf(r.w, s.w, h.w));
                        }
                        if (p.flags & MVF_FLAG_RELU_OUT) r = relu4(r);
                        stcs4(o + 4 * c4, r);
                    }
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Run kernel (production).  A warp owns L consecutive z-voxels of ONE (ix,iy) column for ONE
// chunk of 128*CPL channels.  All L accumulators stay in registers across the whole view loop
// (the per-view grids never exist), and the four taps of the current bilinear cell are cached in
// registers: along a z-run the projected pixel moves ~0.3 px per voxel, so only about one step in
// three changes cell and has to touch the load pipe at all.  On-chip load bandwidth (128 B/clk/SM
// through L1), not HBM, is what bounds this op, so cutting tap loads is the lever.
//   phase A (lanes = (view, z-step) pairs): voxel->pixel projection, floor, weights, validity and
//           the "cell changed" flag, individually rounded fp32 (bit-exact vs the oracle), written
//           to a 32-slot per-warp shared-memory table;
//   phase B (lanes = float4 channel slots): walk the run; per step one broadcast LDS.128 of the
//           four weights, a predicated patch reload (4 coalesced 128-bit read-only loads, 512 B
//           per warp instruction) when the cell changed, and 16 FMAs straight into the accumulators.
constexpr int RUN_WARPS = 8, RUN_TX = 4, RUN_TY = 2;

// FULLC: C/4 is a multiple of 32*CPL, so no lane ever falls off the channel vector.
//
// What bounds this kernel is the register-return bandwidth of the load/store unit (128 B/clk/SM):
// every byte a lane receives -- tap data, but also BROADCAST shared-memory reads -- crosses it.
// So the per-step broadcast is kept to 8 bytes: ax = x1 - u and ay = y1 - w (two steps per
// LDS.128); each lane forms bx = 1 - ax, by = 1 - ay and the four products itself (<= 1 ulp from
// the reference's (u - x0), inside the 1e-5 feature tolerance; indices and validity masks stay
// bit-exact, they come from phase A).  A reload reads one 32-bit word: the byte offset of the
// cell's (y0,x0) tap with the complement of the 4 validity bits in its low nibble -- interior
// cells (nibble 0) take the fast path with constant strides, border cells zero-fill per tap.
template <int CPL, int L, int MODE, bool RELU_IN, bool FULLC>
__global__ void __launch_bounds__(RUN_WARPS * 32, (CPL * L <= 16) ? 2 : 1)
unproject_run_kernel(const __grid_constant__ UnprojParams p, int nchunk) {
    constexpr int VPP = 32 / L;                       // views handled per phase A
    static_assert(VPP * L == 32 && (L % 2) == 0, "L must be even and divide 32");
    __shared__ float sKR[MVF_MAX_VIEWS][12];
    __shared__ float sOff[3];
    __shared__ __align__(16) float2 sA[RUN_WARPS][32];   // (ax, ay) per (view, z-step) slot
    __shared__ __align__(16) uint4 sO[RUN_WARPS][32];    // four tap byte offsets (clamped into the map); .x low nibble = ~valid bits

    const int b = blockIdx.y / nchunk, chunk = blockIdx.y - b * nchunk;
    const int tid = threadIdx.x;
    const bool world = (p.flags & MVF_FLAG_WORLD_GRID) != 0;
    if (tid < p.V) {
        const float* P = p.Rcam + ((size_t)b * p.V + tid) * 12;
        const float* K = p.Kmat + (size_t)b * 9;
        const float* P0 = p.Rmain ? p.Rmain + (size_t)b * 12 : p.Rcam + (size_t)b * p.V * 12;
        float Rinv[12], M[12];
        inverse_pose(P, Rinv);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                M[i * 4 + j] = dot3_rn(K[i * 3 + 0], K[i * 3 + 1], K[i * 3 + 2],
                                       Rinv[0 * 4 + j], Rinv[1 * 4 + j], Rinv[2 * 4 + j]);
        if (world) {
#pragma unroll
            for (int e = 0; e < 12; ++e) sKR[tid][e] = M[e];
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float t3 = (j == 3) ? 1.0f : 0.0f;
                    sKR[tid][i * 4 + j] = dot4_rn(M[i * 4 + 0], M[i * 4 + 1], M[i * 4 + 2], M[i * 4 + 3],
                                                  P0[0 * 4 + j], P0[1 * 4 + j], P0[2 * 4 + j], t3);
                }
        }
    }
    if (tid == 32 && world) {
        const float* P0 = p.Rmain ? p.Rmain + (size_t)b * 12 : p.Rcam + (size_t)b * p.V * 12;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float v = dot4_rn(P0[i * 4 + 0], P0[i * 4 + 1], P0[i * 4 + 2], P0[i * 4 + 3], 0.0f, 0.0f, p.grid_dist, 1.0f);
            sOff[i] = v;
            if (p.out_grid_pos && blockIdx.x == 0 && chunk == 0) p.out_grid_pos[b * 3 + i] = v;
        }
    }
    __syncthreads();

    const int tiles_z = (p.Z + L - 1) / L;
    const int tiles_y = (p.Y + RUN_TY - 1) / RUN_TY;
    int t = blockIdx.x;
    const int tz = t % tiles_z; t /= tiles_z;
    const int ty = t % tiles_y; t /= tiles_y;
    const int warp = tid >> 5, lane = tid & 31;
    const int ixs = t * RUN_TX + (warp % RUN_TX);
    const int iy = ty * RUN_TY + (warp / RUN_TX);
    const int z0 = tz * L;
    if (ixs >= p.Xs || iy >= p.Y) return;             // warp-uniform; no block-wide sync below
    const unsigned FULL = 0xffffffffu;

    const int C = p.C, V = p.V, C4 = C >> 2;
    const size_t view_stride = (size_t)p.fh * p.fw * C;
    const int c4base = chunk * 32 * CPL + lane;
    // lane's channel slot(s); a slot past the end of the channel vector re-reads the last one and
    // is dropped at the store
    unsigned lane_off[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) lane_off[c] = 16u * (unsigned)(FULLC ? (c4base + 32 * c) : min(c4base + 32 * c, C4 - 1));
    const float* feats_b = p.feats + (size_t)b * V * view_stride;
    const long long strideX = 4ll * C, strideY = 4ll * C * p.fw;      // bytes between x / y neighbours

    float gxv = p.gx[p.x_begin + ixs], gyv = p.gy[iy];
    if (world) { gxv = add_rn(gxv, sOff[0]); gyv = add_rn(gyv, sOff[1]); }

    // accumulators and the cached bilinear patch live as packed fp32x2 pairs (FFMA2 operands)
    const ulonglong2 zz = make_ulonglong2(0ull, 0ull);
    ulonglong2 acc[L][CPL];
#pragma unroll
    for (int k = 0; k < L; ++k)
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[k][c] = zz;
    ulonglong2 tA[CPL], tB[CPL], tC[CPL], tD[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) { tA[c] = zz; tB[c] = zz; tC[c] = zz; tD[c] = zz; }

    for (int v0 = 0; v0 < V; v0 += VPP) {
        // ---- phase A: lane = (view v0 + lane / L, z-step lane % L)
        unsigned rmask;
        {
            const int sub = lane / L, k = lane - sub * L;
            const int v = v0 + sub, iz = z0 + k;
            float ax = 0.f, ay = 0.f;
            int bits = 0, x0 = INT32_MIN, y0 = INT32_MIN;
            if (v < V && iz < p.Z) {
                float z = p.gz[iz];
                if (world) z = add_rn(z, sOff[2]);
                const float* KR = sKR[v];
                const float px = affine_row(KR, 0, gxv, gyv, z);
                const float py = affine_row(KR, 1, gxv, gyv, z);
                const float pz = affine_row(KR, 2, gxv, gyv, z);
                const float u = mul_rn(div_rn(px, pz), p.sx);               // :187
                const float w = mul_rn(div_rn(py, pz), p.sy);               // :188
                if (usable_coord(u) && usable_coord(w)) {
                    const float x0f = floorf(u), y0f = floorf(w);            // :192-195
                    x0 = (int)x0f; y0 = (int)y0f;
                    const bool inx0 = (x0 >= 0) && (x0 < p.fw), inx1 = (x0 + 1 >= 0) && (x0 + 1 < p.fw);
                    const bool iny0 = (y0 >= 0) && (y0 < p.fh), iny1 = (y0 + 1 >= 0) && (y0 + 1 < p.fh);
                    bits = (int)(iny0 && inx0) | ((int)(iny1 && inx0) << 1) | ((int)(iny0 && inx1) << 2) |
                           ((int)(iny1 && inx1) << 3);
                    if (bits) { ax = sub_rn((float)(x0 + 1), u); ay = sub_rn((float)(y0 + 1), w); }   // :214-217
                }
                if (chunk == 0 && (p.out_idx || p.out_valid)) {
                    const size_t vox = (((size_t)b * V + v) * p.Xs + ixs) * p.Y * p.Z + (size_t)iy * p.Z + iz;
                    if (p.out_idx) { p.out_idx[vox * 2 + 0] = y0; p.out_idx[vox * 2 + 1] = x0; }
                    if (p.out_valid) p.out_valid[vox] = (uint8_t)bits;
                }
            }
            // sampling state of this step: the cell, or "nothing" (-1).  The patch registers are
            // refreshed whenever the state differs from the previous step's (always at a run start).
            const int cell = bits ? (y0 * p.fw + x0) : INT32_MIN;
            const int pcell = __shfl_up_sync(FULL, cell, 1), pbits = __shfl_up_sync(FULL, bits, 1);
            const bool reload = (k == 0) || (cell != pcell) || (bits != pbits);   // bits: (y, fw-1) and (y+1, -1) share a cell id
            rmask = __ballot_sync(FULL, reload);
            sA[warp][lane] = make_float2(ax, ay);
            // tap byte offsets, each clamped into the map so that every load is legal; the complement of the
            // validity bits rides in the low nibble of .x (offsets are multiples of 16): a tap that is outside
            // the map is zeroed after the load (TF-GPU gather_nd zero fill)
            uint4 o4 = make_uint4(15u, 0u, 0u, 0u);
            if (bits) {
                const int xa = min(max(x0, 0), p.fw - 1), xb = min(max(x0 + 1, 0), p.fw - 1);
                const int ya = min(max(y0, 0), p.fh - 1), yb = min(max(y0 + 1, 0), p.fh - 1);
                const unsigned CB = 4u * (unsigned)C;
                const unsigned oa = (unsigned)(ya * p.fw + xa) * CB;
                const unsigned dX = (xb != xa) ? CB : 0u, dY = (yb != ya) ? (unsigned)p.fw * CB : 0u;
                o4 = make_uint4(oa | (unsigned)(15 ^ bits), oa + dY, oa + dX, oa + dX + dY);
            }
            sO[warp][lane] = o4;
        }
        __syncwarp();
        // ---- phase B: lanes = float4 channel slots.  The view loop stays ROLLED (only the L z-steps are
        // unrolled, they index the accumulator registers): the hot loop must fit the instruction cache.
#pragma unroll 1
        for (int sub = 0; sub < VPP; ++sub) {
            const int vv = v0 + sub;
            if (vv >= V) break;
            // per-lane 64-bit base of this view; a tap address is base + warp-uniform byte offset
            const char* vb[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) vb[c] = (const char*)(feats_b + (size_t)vv * view_stride) + lane_off[FULLC ? 0 : c];
            const unsigned vmask = rmask >> (sub * L);
            const float2* aslot = &sA[warp][sub * L];
            const uint4* oslot = &sO[warp][sub * L];
            float4 aa = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < L; ++k) {
                if ((vmask >> k) & 1u) {                                   // warp-uniform
                    const uint4 o = oslot[k];                              // broadcast LDS.128
                    const unsigned ox = o.x & ~15u;
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        const char* q = FULLC ? vb[0] + 512 * c : vb[c];
                        tA[c] = ldg2x2(q + ox); tB[c] = ldg2x2(q + o.y); tC[c] = ldg2x2(q + o.z); tD[c] = ldg2x2(q + o.w);
                    }
                    if (o.x & 15u) {                                       // rare: border cell / nothing sampled
#pragma unroll
                        for (int c = 0; c < CPL; ++c) {
                            if (o.x & 1u) tA[c] = zz;
                            if (o.x & 2u) tB[c] = zz;
                            if (o.x & 4u) tC[c] = zz;
                            if (o.x & 8u) tD[c] = zz;
                        }
                    }
                }
                if ((k & 1) == 0) aa = *reinterpret_cast<const float4*>(aslot + k);   // broadcast LDS.128: two steps
                const float ax = (k & 1) ? aa.z : aa.x, ay = (k & 1) ? aa.w : aa.y;
                const float bx = 1.0f - ax, by = 1.0f - ay;
                const float wa = ax * ay, wb = ax * by, wc = bx * ay, wd = bx * by;
                if (MODE != MVF_FUSE_NONE && MODE != MVF_FUSE_MAX && !RELU_IN) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c)                          // 8 FFMA2 = 16 fp32 FMAs
                        acc[k][c] = fma2x2(wd, tD[c], fma2x2(wc, tC[c], fma2x2(wb, tB[c], fma2x2(wa, tA[c], acc[k][c]))));
                } else {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        float4 val = unpack4(fma2x2(wd, tD[c], fma2x2(wc, tC[c], fma2x2(wb, tB[c], fma2x2(wa, tA[c], zz)))));
                        if (RELU_IN) val = relu4(val);
                        if (MODE == MVF_FUSE_NONE) {
                            if (z0 + k < p.Z && (FULLC || c4base + 32 * c < C4)) {
                                float* o = p.out + ((((size_t)b * V + vv) * p.Xs + ixs) * p.Y * p.Z + (size_t)iy * p.Z + z0 + k) * C;
                                stcs4(o + 4 * (c4base + 32 * c), val);
                            }
                        } else if (MODE == MVF_FUSE_MAX) {
                            acc[k][c] = (vv == 0) ? pack4(val) : pack4(max4(unpack4(acc[k][c]), val));
                        } else {
                            acc[k][c] = pack4(add4(unpack4(acc[k][c]), val));
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    if (MODE == MVF_FUSE_NONE) return;
#pragma unroll
    for (int k = 0; k < L; ++k) {
        if (z0 + k >= p.Z) break;
        float* o = p.out + ((((size_t)b * p.Xs + ixs) * p.Y + iy) * p.Z + z0 + k) * C;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
            const int c4 = c4base + 32 * c;
            if (!FULLC && c4 >= C4) continue;
            float4 r = unpack4(acc[k][c]);
            if (MODE == MVF_FUSE_MEAN) r = mul4(p.inv_v, r);
            if (p.bn_scale) {
                const float4 s = ldg4(p.bn_scale + 4 * c4), h = ldg4(p.bn_shift + 4 * c4);
                r = make_float4(fmaf(r.x, s.x, h.x), fmaf(r.y, s.y, h.y), fmaf(r.z, s.z, h.z), fmaf(r.w, s.w, h.w));
            }
            if (p.flags & MVF_FLAG_RELU_OUT) r = relu4(r);
            stcs4(o + 4 * c4, r);
        }
    }
}

template <int CPL, int L, bool RELU_IN, bool FULLC>
static int launch_run_mode(const UnprojParams& p, dim3 grid, int nchunk, cudaStream_t s) {
    switch (p.mode) {
        case MVF_FUSE_NONE: unproject_run_kernel<CPL, L, MVF_FUSE_NONE, RELU_IN, FULLC><<<grid, RUN_WARPS * 32, 0, s>>>(p, nchunk); break;
        case MVF_FUSE_SUM:  unproject_run_kernel<CPL, L, MVF_FUSE_SUM, RELU_IN, FULLC><<<grid, RUN_WARPS * 32, 0, s>>>(p, nchunk); break;
        case MVF_FUSE_MEAN: unproject_run_kernel<CPL, L, MVF_FUSE_MEAN, RELU_IN, FULLC><<<grid, RUN_WARPS * 32, 0, s>>>(p, nchunk); break;
        case MVF_FUSE_MAX:  unproject_run_kernel<CPL, L, MVF_FUSE_MAX, RELU_IN, FULLC><<<grid, RUN_WARPS * 32, 0, s>>>(p, nchunk); break;
        default: return MVF_EINVAL;
    }
    count_launch();
    return check_launch();
}

template <int CPL, int L>
static int launch_run(const UnprojParams& p, int B, cudaStream_t s) {
    const int C4 = p.C / 4;
    const int nchunk = (C4 + 32 * CPL - 1) / (32 * CPL);
    if ((long long)B * nchunk > 65535) return MVF_EUNSUPPORTED;
    const int tiles = ((p.Xs + RUN_TX - 1) / RUN_TX) * ((p.Y + RUN_TY - 1) / RUN_TY) * ((p.Z + L - 1) / L);
    dim3 grid(tiles, B * nchunk);
    const bool fullc = (C4 % (32 * CPL)) == 0;
    const bool relu_in = (p.flags & MVF_FLAG_RELU_IN) != 0;
    if (fullc) return relu_in ? launch_run_mode<CPL, L, true, true>(p, grid, nchunk, s) : launch_run_mode<CPL, L, false, true>(p, grid, nchunk, s);
    return relu_in ? launch_run_mode<CPL, L, true, false>(p, grid, nchunk, s) : launch_run_mode<CPL, L, false, false>(p, grid, nchunk, s);
}

template <int CPL>
static int launch_k1(const UnprojParams& p, dim3 grid, cudaStream_t s) {
    switch (p.mode) {
        case MVF_FUSE_NONE: unproject_fuse_kernel<CPL, MVF_FUSE_NONE><<<grid, K1_THREADS, 0, s>>>(p); break;
        case MVF_FUSE_SUM:  unproject_fuse_kernel<CPL, MVF_FUSE_SUM><<<grid, K1_THREADS, 0, s>>>(p); break;
        case MVF_FUSE_MEAN: unproject_fuse_kernel<CPL, MVF_FUSE_MEAN><<<grid, K1_THREADS, 0, s>>>(p); break;
        case MVF_FUSE_MAX:  unproject_fuse_kernel<CPL, MVF_FUSE_MAX><<<grid, K1_THREADS, 0, s>>>(p); break;
        default: return MVF_EINVAL;
    }
    count_launch();
    return check_launch();
}

// Fill the voxel-centre arrays the way the reference's tf.range calls do.
int fill_centres(const MvfGrid* g, int flags, float* gx, float* gy, float* gz) {
    if (g->nvox <= 0 || g->nvox_z <= 0 || g->nvox > MVF_MAX_DIM || g->nvox_z > MVF_MAX_DIM) return MVF_EUNSUPPORTED;
    const int nx = tf1_range(g->vmin + g->vsize / 2.0, g->vmax, g->vsize, gx, MVF_MAX_DIM);   // :157
    if (nx != g->nvox) return MVF_EINVAL;
    for (int i = 0; i < nx; ++i) gy[i] = gx[i];
    int nz;
    if (flags & MVF_FLAG_WORLD_GRID)      // Notebook/projection.py:80
        nz = tf1_range(-(g->nvox_z - 1) * 0.5 * g->vsize, (g->nvox_z - 1) * 0.5 * g->vsize + g->vsize / 2, g->vsize, gz, MVF_MAX_DIM);
    else                                  // model_multi.py:159
        nz = tf1_range(g->vmin_z + g->vsize_z / 2.0, g->vmax_z, g->vsize_z, gz, MVF_MAX_DIM);
    if (nz != g->nvox_z) return MVF_EINVAL;
    return MVF_OK;
}

}  // namespace mvf

using namespace mvf;

extern "C" int mvf_unproject_fuse(const float* feats, const float* Rcam, const float* Rmain, const float* Kmat,
                                  const MvfGrid* g, int B, int V, int fh, int fw, int C, int img_h, int img_w,
                                  int mode, int flags, double grid_dist, int x_begin, int x_count,
                                  const float* bn_scale, const float* bn_shift,
                                  float* out, int32_t* out_idx, uint8_t* out_valid, float* out_grid_pos,
                                  void* stream) {
    if (!feats || !Rcam || !Kmat || !g || !out) return MVF_ENULL;
    if (B <= 0 || V <= 0 || fh <= 0 || fw <= 0 || C <= 0 || img_h <= 0 || img_w <= 0) return MVF_EINVAL;
    if (mode < MVF_FUSE_NONE || mode > MVF_FUSE_MAX) return MVF_EINVAL;
    if ((bn_scale == nullptr) != (bn_shift == nullptr)) return MVF_ENULL;
    if (C % 4 != 0 || !aligned16(feats) || !aligned16(out) || (bn_scale && (!aligned16(bn_scale) || !aligned16(bn_shift))))
        return MVF_EALIGN;
    if (V > MVF_MAX_VIEWS || C > 1024 || B > 65535) return MVF_EUNSUPPORTED;
    if ((size_t)fh * fw * C >= (size_t)1 << 30) return MVF_EUNSUPPORTED;   // byte offsets inside a view fit 32 bits
    UnprojParams p;
    int rc = fill_centres(g, flags, p.gx, p.gy, p.gz);
    if (rc != MVF_OK) return rc;
    if (x_count == 0) { x_begin = 0; x_count = g->nvox; }
    if (x_begin < 0 || x_count < 0 || x_begin + x_count > g->nvox) return MVF_EINVAL;
    p.feats = feats; p.Rcam = Rcam; p.Rmain = Rmain; p.Kmat = Kmat;
    p.bn_scale = (mode == MVF_FUSE_NONE) ? nullptr : bn_scale;
    p.bn_shift = (mode == MVF_FUSE_NONE) ? nullptr : bn_shift;
    p.out = out; p.out_idx = out_idx; p.out_valid = out_valid; p.out_grid_pos = out_grid_pos;
    p.B = B; p.V = V; p.fh = fh; p.fw = fw; p.C = C;
    p.X = g->nvox; p.Y = g->nvox; p.Z = g->nvox_z; p.x_begin = x_begin; p.Xs = x_count;
    p.mode = mode; p.flags = (mode == MVF_FUSE_NONE) ? (flags & ~MVF_FLAG_RELU_OUT) : flags;
    p.sy = (float)((double)fh / (double)img_h);          // :153  float(fh) / IMAGE_SHAPE[0]
    p.sx = (float)((double)fw / (double)img_w);          // :154
    p.inv_v = 1.0f / (float)V;
    p.grid_dist = (float)grid_dist;
    const int tiles = ((p.Xs + BRICK_X - 1) / BRICK_X) * ((p.Y + BRICK_Y - 1) / BRICK_Y) * ((p.Z + BRICK_Z - 1) / BRICK_Z);
    dim3 grid(tiles, B);
    cudaStream_t s = (cudaStream_t)stream;
    const int C4 = C / 4;
    // Default: the run kernel; one warp covers 256 channels x 8 z-steps when C is a multiple of 256 (the FPN
    // width), else 128 channels x 16 z-steps.  MVF_K1_VARIANT (debug / A-B measurement) overrides:
    // 0 = first-generation brick kernel, 1 = run 128ch x 16, 2 = run 256ch x 8, 3 = run 128ch x 8.
    static const int variant = [] { const char* e = getenv("MVF_K1_VARIANT"); return e ? atoi(e) : -1; }();
    if (variant < 0) return (C4 % 64 == 0) ? launch_run<2, 8>(p, B, s) : launch_run<1, 16>(p, B, s);
    if (variant == 1) return launch_run<1, 16>(p, B, s);
    if (variant == 2) return launch_run<2, 8>(p, B, s);
    if (variant == 3) return launch_run<1, 8>(p, B, s);
    if (C4 <= 32) return launch_k1<1>(p, grid, s);
    if (C4 <= 64) return launch_k1<2>(p, grid, s);
    if (C4 <= 128) return launch_k1<4>(p, grid, s);
    return launch_k1<8>(p, grid, s);
}
