// K1  unproject_fuse: unproj_feat (+ fused view reduction / BN / ReLU) for sm_100a.
//
// Replaces mrcnn/model_multi.py:130-228 (unproj_feat) and :401-404 (grid_reas 'add'), plus the
// world-frame notebook variant Notebook/projection.py:47-151.  The per-view grids
// [B,V,X,Y,Z,C] are never materialised unless mode == MVF_FUSE_NONE.
//
// One kernel lives here: unproject_slot_kernel (design notes above its definition).
#include "mvf_common.cuh"
#include <stdlib.h>

namespace mvf {

struct UnprojParams {
    const float* feats; const float* Rcam; const float* Rmain; const float* Kmat;
    const float* bn_scale; const float* bn_shift;
    float* out; int32_t* out_idx; uint8_t* out_valid; float* out_grid_pos;
    // mode NONE only: write the per-view grids as the fp16 (hi, lo) halves of a stride-2 conv operand (parity sub-lattice layout
    // [B*V, 8, X/2, Y/2, Z/2, C]) instead of fp32; split_tail[0] = bits of the bound on max|value|, split_tail[1] <- 2^-s
    uint2* out16_hi; uint2* out16_lo; unsigned* split_tail; int out16_s2d;
    int B, V, fh, fw, C, X, Y, Z, x_begin, Xs;
    int mode, flags;
    float sx, sy, inv_v, grid_dist;
    float gx[MVF_MAX_DIM], gy[MVF_MAX_DIM], gz[MVF_MAX_DIM];
};


// ---------------------------------------------------------------------------------------------
// Slot kernel (production).  A warp owns L consecutive z-voxels of ONE (ix,iy) column for ONE
// chunk of 128*CPL channels; a CTA (8 warps = 4x2 columns) walks a range of z-tiles so that the
// pose prologue is paid once per many voxels.  All L accumulators stay in registers across the
// whole view loop (the per-view grids never exist).
//
// What bounds this op on B200 is neither HBM nor FMA but the register-return path of the
// load/store unit: measured (tools/microbench*.cu, profiles/microbench_r1.txt) a coalesced
// LDG.128 costs 4.0 SM-cycles wherever it hits, a broadcast LDS.128 2.4, while the 16 FFMA2 of
// one (voxel, view) step cost 8.2.  A naive gather needs 4 taps x 1 KB = 32 LSU cycles per step.
// So the kernel is organised around loading as few tap vectors as possible:
//   * the bilinear 2x2 patch lives in four register SLOTS addressed by coordinate PARITY:
//     slot (r,c) holds the tap whose row has parity r and whose column has parity c.  When the
//     projected pixel crosses into the next cell along the z-run only the slots whose
//     coordinate changed are reloaded (2 of 4 for an axis-aligned move) and nothing is shuffled
//     between registers; the weights are assigned per slot instead (0.95 tap loads per
//     (voxel, view) on workload T instead of 4);
//   * steps whose four taps are all outside the map (21 % on T) skip loads and FMAs altogether;
//   * the per-step broadcast is 8 bytes: the weights of slot column 0 / slot row 0; the lane
//     forms the complements (exact for coordinates >= 1, Sterbenz) and the four products itself.
//   phase A (lanes = (view, z-step) pairs): voxel->pixel projection, floor, weights, validity and
//           the per-slot "reload" flags, individually rounded fp32 (bit-exact vs the oracle),
//           written to a 32-slot per-warp shared-memory table; masks travel as ballots;
//   phase B (lanes = float4 channel slots): walk the run.
#ifndef MVF_K1_TX
#define MVF_K1_TX 4
#endif
constexpr int RUN_TX = MVF_K1_TX, RUN_TY = 2, RUN_WARPS = RUN_TX * RUN_TY;

template <int CPL, int L, int MODE, bool RELU_IN, bool FULLC, bool AUX>
__global__ void __launch_bounds__(RUN_WARPS * 32, (CPL * L <= 16) ? 2 : 1)
unproject_slot_kernel(const __grid_constant__ UnprojParams p, int nchunk, int zsplit) {
    constexpr int VPP = 32 / L;                       // views handled per phase A
    constexpr bool PLAIN = (MODE == MVF_FUSE_SUM || MODE == MVF_FUSE_MEAN) && !RELU_IN;   // accumulate straight into acc
    static_assert(VPP * L == 32 && (L % 2) == 0, "L must be even and divide 32");
    __shared__ float sKR[MVF_MAX_VIEWS][12];
    __shared__ float sOff[3];
    // the four slot weights (w00, w01, w10, w11) of every (view, z-step); a slot outside the map has weight 0
    __shared__ __align__(16) float4 sW[RUN_WARPS][32];
    __shared__ __align__(16) uint4 sO[RUN_WARPS][32];    // byte offsets of slots 00,01,10,11 (clamped into the map); .x low nibble = out-of-map bits

    const int b = blockIdx.y / nchunk, chunk = blockIdx.y - b * nchunk;
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
            const float v = dot4_rn(P0[i * 4 + 0], P0[i * 4 + 1], P0[i * 4 + 2], P0[i * 4 + 3], 0.0f, 0.0f, p.grid_dist, 1.0f);
            sOff[i] = v;
            if (p.out_grid_pos && blockIdx.x == 0 && chunk == 0) p.out_grid_pos[b * 3 + i] = v;
        }
    }
    __syncthreads();

    const int tiles_z = (p.Z + L - 1) / L;
    const int tiles_y = (p.Y + RUN_TY - 1) / RUN_TY;
    int t = blockIdx.x;
    const int zpart = t % zsplit; t /= zsplit;
    const int ty = t % tiles_y; t /= tiles_y;
    const int warp = tid >> 5, lane = tid & 31;
    const int ixs = t * RUN_TX + (warp % RUN_TX);
    const int iy = ty * RUN_TY + (warp / RUN_TX);
    if (ixs >= p.Xs || iy >= p.Y) return;             // warp-uniform; no block-wide sync below
    const int tz_begin = (int)(((long long)zpart * tiles_z) / zsplit), tz_end = (int)(((long long)(zpart + 1) * tiles_z) / zsplit);
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
    const unsigned CB = 4u * (unsigned)C;                              // bytes per tap vector

    float gxv = p.gx[p.x_begin + ixs], gyv = p.gy[iy];
    if (world) { gxv = add_rn(gxv, sOff[0]); gyv = add_rn(gyv, sOff[1]); }
    const int sub_a = lane / L, k_a = lane - sub_a * L;               // phase A role of this lane
    constexpr unsigned LMASK = (1u << L) - 1u;
    // byte distance of the lane's c-th channel slot from its first (a compile-time constant when FULLC)
    auto coff = [&](int c) -> int { return FULLC ? 512 * c : (int)(lane_off[c] - lane_off[0]); };
    const size_t view_bytes = view_stride * sizeof(float);
    const char* vbase = (const char*)feats_b + lane_off[0];
    const bool has_bn = p.bn_scale != nullptr, relu_out = (p.flags & MVF_FLAG_RELU_OUT) != 0;
    float split_scale = 1.0f;
    if (MODE == MVF_FUSE_NONE && p.out16_hi) {
        float inv;
        split_scale = pow2_scale(p.split_tail[0], &inv);
        if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) reinterpret_cast<float*>(p.split_tail)[1] = inv;
    }
    float* const obase0 = (MODE == MVF_FUSE_NONE) ? nullptr : p.out + (((size_t)b * p.Xs + ixs) * p.Y + iy) * p.Z * C + 4 * c4base;

    const ulonglong2 zz = make_ulonglong2(0ull, 0ull);
    for (int tz = tz_begin; tz < tz_end; ++tz) {
    const int z0 = tz * L;
    // accumulators and the cached bilinear patch live as packed fp32x2 pairs (FFMA2 operands)
    ulonglong2 acc[L][CPL];
#pragma unroll
    for (int k = 0; k < L; ++k)
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[k][c] = zz;
    ulonglong2 T00[CPL], T01[CPL], T10[CPL], T11[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) { T00[c] = zz; T01[c] = zz; T10[c] = zz; T11[c] = zz; }

    for (int v0 = 0; v0 < V; v0 += VPP) {
        // ---- phase A: lane = (view v0 + lane / L, z-step lane % L)
        unsigned vmask, lm0, lm1, lm2, lm3;
        {
            const int v = v0 + sub_a, iz = z0 + k_a;
            float4 w4 = zero4();
            int bits = 0, x0 = INT32_MIN, y0 = INT32_MIN;
            int row0 = 0, row1 = 0, col0 = 0, col1 = 0;
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
                    if (bits) {
                        const int ox = x0 & 1, oy = y0 & 1;
                        col0 = x0 + ox; col1 = x0 + 1 - ox;                  // the even / the odd column of {x0, x0+1}
                        row0 = y0 + oy; row1 = y0 + 1 - oy;
                        // weight of the tap column x0 is (x1 - u), of x0+1 it is (u - x0); a column / row outside
                        // the map gets weight 0 = the zero fill of TF-GPU gather_nd   (:209-219)
                        const float wxa = inx0 ? sub_rn((float)(x0 + 1), u) : 0.f, wxb = inx1 ? sub_rn(u, x0f) : 0.f;
                        const float wya = iny0 ? sub_rn((float)(y0 + 1), w) : 0.f, wyb = iny1 ? sub_rn(w, y0f) : 0.f;
                        const float wx0 = ox ? wxb : wxa, wx1 = ox ? wxa : wxb;
                        const float wy0 = oy ? wyb : wya, wy1 = oy ? wya : wyb;
                        w4 = make_float4(mul_rn(wy0, wx0), mul_rn(wy0, wx1), mul_rn(wy1, wx0), mul_rn(wy1, wx1));
                    }
                }
                if (AUX && chunk == 0) {                  // index / validity side outputs (parity tests): compiled out of the production kernel
                    const size_t vox = (((size_t)b * V + v) * p.Xs + ixs) * p.Y * p.Z + (size_t)iy * p.Z + iz;
                    if (p.out_idx) { p.out_idx[vox * 2 + 0] = y0; p.out_idx[vox * 2 + 1] = x0; }
                    if (p.out_valid) p.out_valid[vox] = (uint8_t)bits;
                }
            }
            const bool valid = bits != 0;
            // a slot is (re)loaded when the coordinate it has to hold differs from the previous step's
            // (always at a run start and after a step that sampled nothing)
            const int prow0 = __shfl_up_sync(FULL, row0, 1), prow1 = __shfl_up_sync(FULL, row1, 1);
            const int pcol0 = __shfl_up_sync(FULL, col0, 1), pcol1 = __shfl_up_sync(FULL, col1, 1);
            const bool pvalid = __shfl_up_sync(FULL, (int)valid, 1) != 0;
            const bool first = (k_a == 0) || !pvalid;
            const bool cr0 = first || row0 != prow0, cr1 = first || row1 != prow1;
            const bool cc0 = first || col0 != pcol0, cc1 = first || col1 != pcol1;
            vmask = __ballot_sync(FULL, valid);
            lm0 = __ballot_sync(FULL, valid && (cr0 || cc0));
            lm1 = __ballot_sync(FULL, valid && (cr0 || cc1));
            lm2 = __ballot_sync(FULL, valid && (cr1 || cc0));
            lm3 = __ballot_sync(FULL, valid && (cr1 || cc1));
            sW[warp][lane] = w4;
            // slot byte offsets, clamped into the map so that every load is legal; a slot whose coordinate is
            // outside the map is zeroed after the load (TF-GPU gather_nd zero fill), flagged in the low nibble
            // of .x (offsets are multiples of 16)
            uint4 o4 = make_uint4(0u, 0u, 0u, 0u);
            if (valid) {
                const int r0c = min(max(row0, 0), p.fh - 1), r1c = min(max(row1, 0), p.fh - 1);
                const int c0c = min(max(col0, 0), p.fw - 1), c1c = min(max(col1, 0), p.fw - 1);
                o4 = make_uint4((unsigned)(r0c * p.fw + c0c) * CB, (unsigned)(r0c * p.fw + c1c) * CB,
                                (unsigned)(r1c * p.fw + c0c) * CB, (unsigned)(r1c * p.fw + c1c) * CB);
            }
            sO[warp][lane] = o4;
        }
        __syncwarp();
        // ---- phase B: lanes = float4 channel slots.  The view loop stays ROLLED (only the L z-steps are
        // unrolled, they index the accumulator registers): the hot loop must fit the instruction cache.
        const char* vptr = vbase + (size_t)v0 * view_bytes;                 // per-lane base of view v0 (+ channel slot)
#pragma unroll 1
        for (int sub = 0; sub < VPP && v0 + sub < V; ++sub, vptr += view_bytes) {
            const int vv = v0 + sub;
            const unsigned vm = (vmask >> (sub * L)) & LMASK;
            if (PLAIN && vm == 0u) continue;                           // the whole run misses this view
            // reload flags of this view: slots 00 | 01 in mlo, 10 | 11 in mhi (bit k / bit L+k)
            unsigned mlo = ((lm0 >> (sub * L)) & LMASK) | (((lm1 >> (sub * L)) & LMASK) << L);
            unsigned mhi = ((lm2 >> (sub * L)) & LMASK) | (((lm3 >> (sub * L)) & LMASK) << L);
            unsigned many = mlo | mhi;
            // pin the three words in registers: every per-step test below is then ONE LOP3 against a constant
            // (left alone, ptxas re-derives them from the ballots with shifts at every use)
            asm volatile("" : "+r"(mlo), "+r"(mhi), "+r"(many));
            const float4* wslot = &sW[warp][sub * L];
            const uint4* oslot = &sO[warp][sub * L];
            float4 wn = wslot[0];                                      // weights travel one step ahead of their use
            if (PLAIN) {
                // Software pipeline over the run: the tap loads of step k+1 are issued BETWEEN the FMA groups of step k
                // (a slot may be overwritten as soon as the FMAs that read it have been issued), so every reload has
                // 8-16 FFMA2 of this warp's own work to hide behind instead of stalling the first FMA of its step.
                // The per-accumulator FMA order (slots 00, 01, 10, 11) is unchanged.
                auto reload_lo = [&](int kk, const uint4& o) {
                    if (mlo & (1u << kk)) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) T00[c] = ldg2x2(vptr + coff(c) + o.x);
                    }
                    if (mlo & (1u << (L + kk))) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) T01[c] = ldg2x2(vptr + coff(c) + o.y);
                    }
                };
                auto reload_hi = [&](int kk, const uint4& o) {
                    if (mhi & (1u << kk)) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) T10[c] = ldg2x2(vptr + coff(c) + o.z);
                    }
                    if (mhi & (1u << (L + kk))) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) T11[c] = ldg2x2(vptr + coff(c) + o.w);
                    }
                };
                if (many & (1u | (1u << L))) {                          // step 0: nothing of this run to hide behind
                    const uint4 o = oslot[0];
                    reload_lo(0, o); reload_hi(0, o);
                }
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    const bool valid = (vm >> k) & 1u;                      // warp-uniform
                    const float4 ww = wn;
                    if (k + 1 < L) wn = wslot[k + 1];                       // broadcast LDS.128
                    const bool nxt = (k + 1 < L) && (many & ((1u << (k + 1)) | (1u << (L + k + 1))));   // step k+1 reloads something
                    if (!valid) {
                        if (nxt) { const uint4 o = oslot[k + 1]; reload_lo(k + 1, o); reload_hi(k + 1, o); }
                        continue;
                    }
                    if (nxt) {
                        const uint4 o = oslot[k + 1];                       // broadcast LDS.128
#pragma unroll
                        for (int c = 0; c < CPL; ++c) acc[k][c] = fma2x2(ww.y, T01[c], fma2x2(ww.x, T00[c], acc[k][c]));
                        reload_lo(k + 1, o);
#pragma unroll
                        for (int c = 0; c < CPL; ++c) acc[k][c] = fma2x2(ww.w, T11[c], fma2x2(ww.z, T10[c], acc[k][c]));
                        reload_hi(k + 1, o);
                    } else {
#pragma unroll
                        for (int c = 0; c < CPL; ++c)                          // 8 FFMA2 = 16 fp32 FMAs
                            acc[k][c] = fma2x2(ww.w, T11[c], fma2x2(ww.z, T10[c], fma2x2(ww.y, T01[c], fma2x2(ww.x, T00[c], acc[k][c]))));
                    }
                }
                continue;
            }
#pragma unroll
            for (int k = 0; k < L; ++k) {
                const bool valid = (vm >> k) & 1u;                          // warp-uniform
                const float4 ww = wn;
                if (k + 1 < L) wn = wslot[k + 1];                           // broadcast LDS.128
                if (PLAIN && !valid) continue;
                if (many & ((1u << k) | (1u << (L + k)))) {
                    const uint4 o = oslot[k];                               // broadcast LDS.128
                    if (mlo & (1u << k)) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) T00[c] = ldg2x2(vptr + coff(c) + o.x);
                    }
                    if (mlo & (1u << (L + k))) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) T01[c] = ldg2x2(vptr + coff(c) + o.y);
                    }
                    if (mhi & (1u << k)) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) T10[c] = ldg2x2(vptr + coff(c) + o.z);
                    }
                    if (mhi & (1u << (L + k))) {
#pragma unroll
                        for (int c = 0; c < CPL; ++c) T11[c] = ldg2x2(vptr + coff(c) + o.w);
                    }
                }
                const float w00 = ww.x, w01 = ww.y, w10 = ww.z, w11 = ww.w;
                if (PLAIN) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c)                          // 8 FFMA2 = 16 fp32 FMAs
                        acc[k][c] = fma2x2(w11, T11[c], fma2x2(w10, T10[c], fma2x2(w01, T01[c], fma2x2(w00, T00[c], acc[k][c]))));
                } else {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) {
                        float4 val = zero4();
                        if (valid) val = unpack4(fma2x2(w11, T11[c], fma2x2(w10, T10[c], fma2x2(w01, T01[c], fma2x2(w00, T00[c], zz)))));
                        if (RELU_IN) val = relu4(val);
                        if (MODE == MVF_FUSE_NONE) {
                            if (z0 + k < p.Z && (FULLC || c4base + 32 * c < C4)) {
                                if (p.out16_hi) {                  // operand halves for the U-Net's stride-2 conv (whole grid, even dims)
                                    const int z = z0 + k, sub = ((ixs & 1) * 2 + (iy & 1)) * 2 + (z & 1);
                                    const size_t o = p.out16_s2d
                                        ? ((((((size_t)b * V + vv) * 8 + sub) * (p.X / 2) + (ixs >> 1)) * (p.Y / 2) + (iy >> 1)) * (p.Z / 2)
                                           + (z >> 1)) * C4 + c4base + 32 * c
                                        : (((((size_t)b * V + vv) * p.X + ixs) * p.Y + iy) * p.Z + z) * C4 + c4base + 32 * c;
                                    uint2 h2, l2;
                                    split_half4(val, split_scale, &h2, &l2);
                                    p.out16_hi[o] = h2; p.out16_lo[o] = l2;
                                } else {
                                    float* o = p.out + ((((size_t)b * V + vv) * p.Xs + ixs) * p.Y * p.Z + (size_t)iy * p.Z + z0 + k) * C;
                                    stcs4(o + 4 * (c4base + 32 * c), val);
                                }
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
    if (MODE != MVF_FUSE_NONE) {
        float* obase = obase0 + (size_t)z0 * C;
        if (!has_bn && !relu_out && MODE != MVF_FUSE_MEAN) {
            // plain sum / max: 2 streaming 128-bit stores per voxel, nothing else
#pragma unroll
            for (int k = 0; k < L; ++k) {
                if (z0 + k >= p.Z) break;
#pragma unroll
                for (int c = 0; c < CPL; ++c)
                    if (FULLC || c4base + 32 * c < C4) stcs2x2(obase + (size_t)k * C + 128 * c, acc[k][c]);
            }
        } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                if (!FULLC && c4base + 32 * c >= C4) continue;
                float4 bs = make_float4(1.f, 1.f, 1.f, 1.f), bh = zero4();
                if (has_bn) { bs = ldg4(p.bn_scale + 4 * (c4base + 32 * c)); bh = ldg4(p.bn_shift + 4 * (c4base + 32 * c)); }
#pragma unroll
                for (int k = 0; k < L; ++k) {
                    if (z0 + k >= p.Z) break;
                    float4 r = unpack4(acc[k][c]);
                    if (MODE == MVF_FUSE_MEAN) r = mul4(p.inv_v, r);
                    if (has_bn) r = make_float4(fmaf(r.x, bs.x, bh.x), fmaf(r.y, bs.y, bh.y), fmaf(r.z, bs.z, bh.z), fmaf(r.w, bs.w, bh.w));
                    if (relu_out) r = relu4(r);
                    stcs4(obase + (size_t)k * C + 128 * c, r);
                }
            }
        }
    }
    }   // z-tile loop
}

template <int CPL, int L, bool RELU_IN, bool FULLC, bool AUX>
static int launch_run_mode(const UnprojParams& p, dim3 grid, int nchunk, int zsplit, cudaStream_t s) {
    switch (p.mode) {
        case MVF_FUSE_NONE: unproject_slot_kernel<CPL, L, MVF_FUSE_NONE, RELU_IN, FULLC, AUX><<<grid, RUN_WARPS * 32, 0, s>>>(p, nchunk, zsplit); break;
        case MVF_FUSE_SUM:  unproject_slot_kernel<CPL, L, MVF_FUSE_SUM, RELU_IN, FULLC, AUX><<<grid, RUN_WARPS * 32, 0, s>>>(p, nchunk, zsplit); break;
        case MVF_FUSE_MEAN: unproject_slot_kernel<CPL, L, MVF_FUSE_MEAN, RELU_IN, FULLC, AUX><<<grid, RUN_WARPS * 32, 0, s>>>(p, nchunk, zsplit); break;
        case MVF_FUSE_MAX:  unproject_slot_kernel<CPL, L, MVF_FUSE_MAX, RELU_IN, FULLC, AUX><<<grid, RUN_WARPS * 32, 0, s>>>(p, nchunk, zsplit); break;
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
    const int tiles_z = (p.Z + L - 1) / L;
    const long long cols = (long long)((p.Xs + RUN_TX - 1) / RUN_TX) * ((p.Y + RUN_TY - 1) / RUN_TY);
    // a CTA walks tiles_z / zsplit z-tiles of its 4x2 columns; split z only as far as needed to give every
    // SM several CTAs (2 resident per SM, >= 4 waves)
    static const int zs_env = env_int("MVF_K1_ZSPLIT", 0);
    int zsplit = 1;
    while (zsplit < tiles_z && cols * zsplit * B * nchunk < 148ll * 2 * 4) zsplit *= 2;
    if (zs_env > 0) zsplit = zs_env;
    if (zsplit > tiles_z) zsplit = tiles_z;
    dim3 grid((unsigned)(cols * zsplit), B * nchunk);
    const bool fullc = (C4 % (32 * CPL)) == 0;
    const bool relu_in = (p.flags & MVF_FLAG_RELU_IN) != 0;
    if (p.out_idx || p.out_valid) {       // side outputs requested: the (rare) instrumented build
        if (fullc) return relu_in ? launch_run_mode<CPL, L, true, true, true>(p, grid, nchunk, zsplit, s) : launch_run_mode<CPL, L, false, true, true>(p, grid, nchunk, zsplit, s);
        return relu_in ? launch_run_mode<CPL, L, true, false, true>(p, grid, nchunk, zsplit, s) : launch_run_mode<CPL, L, false, false, true>(p, grid, nchunk, zsplit, s);
    }
    if (fullc) return relu_in ? launch_run_mode<CPL, L, true, true, false>(p, grid, nchunk, zsplit, s) : launch_run_mode<CPL, L, false, true, false>(p, grid, nchunk, zsplit, s);
    return relu_in ? launch_run_mode<CPL, L, true, false, false>(p, grid, nchunk, zsplit, s) : launch_run_mode<CPL, L, false, false, false>(p, grid, nchunk, zsplit, s);
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

static int unproject_impl(const float* feats, const float* Rcam, const float* Rmain, const float* Kmat,
                          const MvfGrid* g, int B, int V, int fh, int fw, int C, int img_h, int img_w,
                          int mode, int flags, double grid_dist, int x_begin, int x_count,
                          const float* bn_scale, const float* bn_shift,
                          float* out, int32_t* out_idx, uint8_t* out_valid, float* out_grid_pos,
                          uint2* out16_hi, uint2* out16_lo, unsigned* split_tail, int out16_s2d, void* stream) {
    if (!feats || !Rcam || !Kmat || !g) return MVF_ENULL;
    if (B <= 0 || V <= 0 || fh <= 0 || fw <= 0 || C <= 0 || img_h <= 0 || img_w <= 0) return MVF_EINVAL;
    if (mode < MVF_FUSE_NONE || mode > MVF_FUSE_MAX) return MVF_EINVAL;
    if ((bn_scale == nullptr) != (bn_shift == nullptr)) return MVF_ENULL;
    if (C % 4 != 0 || !aligned16(feats) || (out && !aligned16(out)) || (bn_scale && (!aligned16(bn_scale) || !aligned16(bn_shift))))
        return MVF_EALIGN;
    if (V > MVF_MAX_VIEWS || C > 1024 || B > 65535) return MVF_EUNSUPPORTED;
    if ((size_t)fh * fw * C >= (size_t)1 << 30) return MVF_EUNSUPPORTED;   // byte offsets inside a view fit 32 bits
    UnprojParams p;
    int rc = fill_centres(g, flags, p.gx, p.gy, p.gz);
    if (rc != MVF_OK) return rc;
    if (x_count < 0) { x_begin = 0; x_count = g->nvox; }                   // MVF_WHOLE_GRID
    if (x_begin < 0 || x_begin + x_count > g->nvox) return MVF_EINVAL;
    if (x_count == 0) return MVF_OK;                                        // empty slab of a sharded caller: nothing to write
    if (!out && !out16_hi) return MVF_ENULL;
    p.feats = feats; p.Rcam = Rcam; p.Rmain = Rmain; p.Kmat = Kmat;
    p.bn_scale = (mode == MVF_FUSE_NONE) ? nullptr : bn_scale;
    p.bn_shift = (mode == MVF_FUSE_NONE) ? nullptr : bn_shift;
    p.out = out; p.out_idx = out_idx; p.out_valid = out_valid; p.out_grid_pos = out_grid_pos;
    p.out16_hi = out16_hi; p.out16_lo = out16_lo; p.split_tail = split_tail; p.out16_s2d = out16_s2d;
    p.B = B; p.V = V; p.fh = fh; p.fw = fw; p.C = C;
    p.X = g->nvox; p.Y = g->nvox; p.Z = g->nvox_z; p.x_begin = x_begin; p.Xs = x_count;
    p.mode = mode; p.flags = (mode == MVF_FUSE_NONE) ? (flags & ~MVF_FLAG_RELU_OUT) : flags;
    p.sy = (float)((double)fh / (double)img_h);          // :153  float(fh) / IMAGE_SHAPE[0]
    p.sx = (float)((double)fw / (double)img_w);          // :154
    p.inv_v = 1.0f / (float)V;
    p.grid_dist = (float)grid_dist;
    cudaStream_t s = (cudaStream_t)stream;
    const int C4 = C / 4;
    // One warp covers 256 channels x 8 z-steps when C is a multiple of 256 (the FPN width), else
    // 128 channels x 16 z-steps.  MVF_K1_VARIANT (debug / A-B measurement) overrides:
    // 1 = 128ch x 16, 2 = 256ch x 8, 3 = 128ch x 8.
    static const int variant = env_int("MVF_K1_VARIANT", -1);
    if (variant == 1) return launch_run<1, 16>(p, B, s);
    if (variant == 2) return launch_run<2, 8>(p, B, s);
    if (variant == 3) return launch_run<1, 8>(p, B, s);
    return (C4 % 64 == 0) ? launch_run<2, 8>(p, B, s) : launch_run<1, 16>(p, B, s);
}

extern "C" int mvf_unproject_fuse(const float* feats, const float* Rcam, const float* Rmain, const float* Kmat,
                                  const MvfGrid* g, int B, int V, int fh, int fw, int C, int img_h, int img_w,
                                  int mode, int flags, double grid_dist, int x_begin, int x_count,
                                  const float* bn_scale, const float* bn_shift,
                                  float* out, int32_t* out_idx, uint8_t* out_valid, float* out_grid_pos,
                                  void* stream) {
    if (!out && x_count != 0) return MVF_ENULL;
    return unproject_impl(feats, Rcam, Rmain, Kmat, g, B, V, fh, fw, C, img_h, img_w, mode, flags, grid_dist, x_begin, x_count,
                          bn_scale, bn_shift, out, out_idx, out_valid, out_grid_pos, nullptr, nullptr, nullptr, 0, stream);
}

// unproj_feat straight into the operand format of the convolution that follows it (the U-Net's first conv, model_multi.py:411-421:
// sublattices = 1, the parity-sub-lattice layout of MVF_CONV_S2; 'ident', :446-453: sublattices = 0, plain layout): the per-view
// grids are written once, as the fp16 (hi, lo) halves mvf_conv3d_tc(MVF_FLAG_PRESPLIT) reads, instead of fp32 grids that a split pass would re-read and re-write (4.3 GB at 64^3 with 8 views).
// act_amax: DEVICE pointer to a bound on max|value| -- max|feats| is one (bilinear weights are in [0,1] and sum to <= 1).
extern "C" int mvf_unproject_split_f16(const float* feats, const float* Rcam, const float* Rmain, const float* Kmat,
                                       const MvfGrid* g, int B, int V, int fh, int fw, int C, int img_h, int img_w,
                                       int flags, int sublattices, const float* act_amax, void* conv_ws, size_t ws_bytes, void* stream) {
    if (!g || !act_amax || !conv_ws) return MVF_ENULL;
    if (B <= 0 || V <= 0 || C <= 0) return MVF_EINVAL;
    if (C % 64 != 0 || (flags & MVF_FLAG_WORLD_GRID)) return MVF_EUNSUPPORTED;
    if (sublattices && ((g->nvox & 1) || (g->nvox_z & 1))) return MVF_EUNSUPPORTED;
    if (!aligned16(conv_ws)) return MVF_EALIGN;
    const size_t n1 = (size_t)B * V * g->nvox * g->nvox * g->nvox_z * C;             // elements of the per-view grids
    if (ws_bytes < 4 * n1 + 256) return MVF_EWORKSPACE;                                // [hi n1 halves][lo n1 halves][tail]
    __half* w0 = (__half*)conv_ws;
    unsigned* tail = (unsigned*)(((uintptr_t)(w0 + 2 * n1) + 15) & ~(uintptr_t)15);
    if (cudaMemcpyAsync(tail, act_amax, 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) return MVF_ECUDA;
    return unproject_impl(feats, Rcam, Rmain, Kmat, g, B, V, fh, fw, C, img_h, img_w, MVF_FUSE_NONE, flags & MVF_FLAG_RELU_IN, 0.0, 0, MVF_WHOLE_GRID,
                          nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, (uint2*)w0, (uint2*)(w0 + n1), tail, sublattices != 0, stream);
}
