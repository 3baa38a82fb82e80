// K4  pyramid_roi_align: FPN level assignment + crop_and_resize (1 bilinear sample per bin) for
// all four levels in ONE launch, writing straight into box order.
//
// Replaces PyramidROIAlign.call  mrcnn/model_multi.py:799-885 (level rule :816-828,
// tf.image.crop_and_resize :856-858, re-ordering :861-882).  The crop arithmetic follows
// TensorFlow's crop_and_resize_op (third-party, restated: SURVEY.md spec E) with individually
// rounded fp32 ops so crops are bit-identical to the oracle.
//
// Mapping: one CTA per box; thread 0 derives the level and the sampling steps, then warps
// stride over the ph*pw bins with lanes as float4 channel slots: four coalesced 128-bit
// read-only loads per lane, two lerps, one streaming 128-bit store.
#include "mvf_common.cuh"

namespace mvf {

struct RoiParams {
    const float* boxes; const float* maps[4];
    int H[4], W[4];
    float* out; int32_t* out_level;
    int B, R, C, ph, pw;
    float image_area;
};

__device__ __forceinline__ float4 lerp4_rn(float4 a, float4 b, float t) {
    // a + (b - a) * t, each op rounded (crop_and_resize_op)
    return make_float4(add_rn(a.x, mul_rn(sub_rn(b.x, a.x), t)), add_rn(a.y, mul_rn(sub_rn(b.y, a.y), t)),
                       add_rn(a.z, mul_rn(sub_rn(b.z, a.z), t)), add_rn(a.w, mul_rn(sub_rn(b.w, a.w), t)));
}

// FPN level of a box (model_multi.py:816-828). log is evaluated in double and rounded once
// (= correctly rounded fp32 log, the oracle's definition); round-half-even; non-finite -> 2.
__device__ __forceinline__ int roi_level(float y1, float x1, float y2, float x2, float image_area) {
    const float h = sub_rn(y2, y1), w = sub_rn(x2, x1);
    const float denom = div_rn(224.0f, sqrtf(image_area));
    const float ratio = div_rn(sqrtf(mul_rn(h, w)), denom);
    const float lg = (float)log((double)ratio);
    const float lvl_f = div_rn(lg, 0.693147182464599609375f);        // float32(log 2)
    if (!(fabsf(lvl_f) <= 3.0e38f)) return 2;                        // inf / NaN
    float r = rintf(lvl_f);
    r = fminf(fmaxf(r, -64.f), 64.f);
    return min(5, max(2, 4 + (int)r));
}

__global__ void __launch_bounds__(256)
pyramid_roi_align_kernel(const __grid_constant__ RoiParams p) {
    __shared__ float s_f[6];      // y1*(H-1), hs, x1*(W-1), ws, H-1, W-1
    __shared__ int s_lvl;
    const int box = blockIdx.x;                  // b*R + r
    const int b = box / p.R;
    const float* bx = p.boxes + (size_t)box * 4;
    if (threadIdx.x == 0) {
        const float y1 = bx[0], x1 = bx[1], y2 = bx[2], x2 = bx[3];
        const int lvl = roi_level(y1, x1, y2, x2, p.image_area);
        s_lvl = lvl;
        if (p.out_level && blockIdx.y == 0) p.out_level[box] = lvl;
        const float Hm1 = (float)(p.H[lvl - 2] - 1), Wm1 = (float)(p.W[lvl - 2] - 1);
        s_f[4] = Hm1; s_f[5] = Wm1;
        if (p.ph > 1) { s_f[0] = mul_rn(y1, Hm1); s_f[1] = div_rn(mul_rn(sub_rn(y2, y1), Hm1), (float)(p.ph - 1)); }
        else          { s_f[0] = mul_rn(mul_rn(0.5f, add_rn(y1, y2)), Hm1); s_f[1] = 0.f; }
        if (p.pw > 1) { s_f[2] = mul_rn(x1, Wm1); s_f[3] = div_rn(mul_rn(sub_rn(x2, x1), Wm1), (float)(p.pw - 1)); }
        else          { s_f[2] = mul_rn(mul_rn(0.5f, add_rn(x1, x2)), Wm1); s_f[3] = 0.f; }
    }
    __syncthreads();
    const int lvl = s_lvl;
    const int Hl = p.H[lvl - 2], Wl = p.W[lvl - 2];
    const float* map = p.maps[lvl - 2] + (size_t)b * Hl * Wl * p.C;
    const float Hm1 = s_f[4], Wm1 = s_f[5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int C4 = p.C >> 2;
    const int nbins = p.ph * p.pw;
    // the bins of a box are dealt round-robin to the warps of its gridDim.y CTAs
    for (int bin = blockIdx.y * nwarps + warp; bin < nbins; bin += nwarps * gridDim.y) {
        const int iy = bin / p.pw, ix = bin % p.pw;
        const float in_y = (p.ph > 1) ? add_rn(s_f[0], mul_rn((float)iy, s_f[1])) : s_f[0];
        const float in_x = (p.pw > 1) ? add_rn(s_f[2], mul_rn((float)ix, s_f[3])) : s_f[2];
        float* o = p.out + ((size_t)box * nbins + bin) * p.C;
        const bool ok = (in_y >= 0.f) && (in_y <= Hm1) && (in_x >= 0.f) && (in_x <= Wm1);   // NaN -> extrapolate
        if (!ok) {
            for (int c4 = lane; c4 < C4; c4 += 32) stcs4(o + 4 * c4, zero4());
            continue;
        }
        const float ty = floorf(in_y), tx = floorf(in_x);
        const int top = (int)ty, bot = (int)ceilf(in_y), left = (int)tx, right = (int)ceilf(in_x);
        const float ly = sub_rn(in_y, ty), lx = sub_rn(in_x, tx);
        const float* ptl = map + ((size_t)top * Wl + left) * p.C;
        const float* ptr = map + ((size_t)top * Wl + right) * p.C;
        const float* pbl = map + ((size_t)bot * Wl + left) * p.C;
        const float* pbr = map + ((size_t)bot * Wl + right) * p.C;
        for (int c4 = lane; c4 < C4; c4 += 32) {
            const float4 tl = ldg4(ptl + 4 * c4), tr = ldg4(ptr + 4 * c4);
            const float4 bl = ldg4(pbl + 4 * c4), br = ldg4(pbr + 4 * c4);
            const float4 t = lerp4_rn(tl, tr, lx);
            const float4 bm = lerp4_rn(bl, br, lx);
            stcs4(o + 4 * c4, lerp4_rn(t, bm, ly));
        }
    }
}

}  // namespace mvf

using namespace mvf;

extern "C" int mvf_pyramid_roi_align(const float* boxes, const float* const maps[4], const int H[4],
                                     const int W[4], int B, int R, int C, int image_h, int image_w,
                                     int pool_h, int pool_w, float* out, int32_t* out_level, void* stream) {
    if (!boxes || !maps || !H || !W || !out) return MVF_ENULL;
    if (B <= 0 || R <= 0 || C <= 0 || pool_h <= 0 || pool_w <= 0 || image_h <= 0 || image_w <= 0) return MVF_EINVAL;
    if (C % 4 != 0 || !aligned16(out)) return MVF_EALIGN;
    RoiParams p;
    for (int l = 0; l < 4; ++l) {
        if (!maps[l]) return MVF_ENULL;
        if (H[l] <= 0 || W[l] <= 0) return MVF_EINVAL;
        if (!aligned16(maps[l])) return MVF_EALIGN;
        p.maps[l] = maps[l]; p.H[l] = H[l]; p.W[l] = W[l];
    }
    p.boxes = boxes; p.out = out; p.out_level = out_level;
    p.B = B; p.R = R; p.C = C; p.ph = pool_h; p.pw = pool_w;
    p.image_area = (float)((double)image_h * (double)image_w);       // tf.cast(h*w, float32), :821
    const int nbins = pool_h * pool_w;
    const int threads = (nbins >= 8) ? 256 : 32 * nbins;
    // few boxes (the 100-box mask head): split each box's bins over several CTAs so that every SM gets work
    int split = 1;
    while ((long long)B * R * split < 148ll * 4 && (split + 1) * 8 <= nbins) ++split;
    pyramid_roi_align_kernel<<<dim3(B * R, split), threads, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return check_launch();
}
