// K3  project_rays: proj_grid + nearest3 (+ fused depth_sampling epilogue) for sm_100a.
//
// Replaces mrcnn/model_multi.py:231-322 (proj_grid), :357-369 (nearest3) and the non-conv3d
// branch of depth_sampling (:481-487); notebook variant Notebook/projection.py:253-339.
//
// Mapping: lanes first act as 32 independent ray samples (s,i,j): pixel -> camera ray ->
// world -> grid frame -> voxel index, all in registers with individually rounded fp32 ops
// (bit-exact against the oracle, round-half-even via rintf).  Then the warp walks its 32
// samples: the source voxel offset is broadcast by shuffle and the lanes become float4
// channel slots copying one C-vector per sample with coalesced 128-bit loads (read-only
// path) and streaming 128-bit stores.  The validity mask ("depth visibility": the sample
// lies inside the voxel box) is emitted by the same kernel.
#include "mvf_common.cuh"

namespace mvf {

constexpr int K3_THREADS = 256;

struct ProjParams {
    const float* grid; const float* Rview; const float* Rmain; const float* Kmat; const float* grid_pos;
    const float* w;                     // [S] depth-collapse weights (device) or null
    float* out; int32_t* out_vox; uint8_t* out_valid;
    int B, C, ph, pw, S, X, Y, Z, x_begin, Xs, flags;
    int rv_stride;                      // floats between two scenes' view poses (12: a dense [B,3,4]; V*12: view 0 of Rcam [B,V,3,4])
    float r, lo[3], hi[3], n[3];
    float bias, bn_scale, bn_shift;
    float zs[MVF_MAX_SAMPLES];
};

struct SceneXf {          // per-scene transforms, built once per CTA in shared memory
    float Kp[9];          // r * K            (:239)
    float A[12];          // [R_view | t_view] camera -> world   (:283)
    float RT[12];         // [R_0^T | -R_0^T t_0] world -> grid  (:279-281,:290)
    float gp[3];          // notebook grid_position
};

__device__ __forceinline__ void build_scene(const ProjParams& p, int b, SceneXf* xf) {
    const float* K = p.Kmat + (size_t)b * 9;
#pragma unroll
    for (int e = 0; e < 9; ++e) xf->Kp[e] = mul_rn(K[e], p.r);
    const float* A = p.Rview + (size_t)b * p.rv_stride;
#pragma unroll
    for (int e = 0; e < 12; ++e) xf->A[e] = A[e];
    const float* P0 = p.Rmain ? p.Rmain + (size_t)b * 12 : A;
    inverse_pose(P0, xf->RT);
    if (p.grid_pos) { xf->gp[0] = p.grid_pos[b * 3 + 0]; xf->gp[1] = p.grid_pos[b * 3 + 1]; xf->gp[2] = p.grid_pos[b * 3 + 2]; }
    else { xf->gp[0] = xf->gp[1] = xf->gp[2] = 0.f; }
}

// voxel index of ray sample (s, i, j); returns validity, writes id[3]
__device__ __forceinline__ bool sample_voxel(const ProjParams& p, const SceneXf& xf, int s, int i, int j, int id[3]) {
    const float u = (float)j + 0.5f, v = (float)i + 0.5f;                 // :252 pixel centres
    // back substitution of the upper-triangular solve K' Xc = (u, v, r)   (:263-264)
    const float zc = div_rn(p.r, xf.Kp[8]);
    const float yc = div_rn(sub_rn(v, mul_rn(xf.Kp[5], zc)), xf.Kp[4]);
    const float xc = div_rn(sub_rn(sub_rn(u, mul_rn(xf.Kp[1], yc)), mul_rn(xf.Kp[2], zc)), xf.Kp[0]);
    const float zs = p.zs[s];
    const float X0 = mul_rn(xc, zs), X1 = mul_rn(yc, zs), X2 = mul_rn(zc, zs);   // :270-272
    const float w0 = affine_row(xf.A, 0, X0, X1, X2), w1 = affine_row(xf.A, 1, X0, X1, X2), w2 = affine_row(xf.A, 2, X0, X1, X2);
    float g[3];
    if (p.flags & MVF_FLAG_WORLD_GRID) {                                   // Notebook/projection.py:311
        g[0] = sub_rn(w0, xf.gp[0]); g[1] = sub_rn(w1, xf.gp[1]); g[2] = sub_rn(w2, xf.gp[2]);
    } else {
        g[0] = affine_row(xf.RT, 0, w0, w1, w2); g[1] = affine_row(xf.RT, 1, w0, w1, w2); g[2] = affine_row(xf.RT, 2, w0, w1, w2);
    }
    bool ok = true;
    float q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        q[a] = mul_rn(div_rn(sub_rn(g[a], p.lo[a]), sub_rn(p.hi[a], p.lo[a])), p.n[a]);   // :297-298
        ok = ok && usable_coord(q[a]);
    }
    if (!ok) { id[0] = id[1] = id[2] = INT32_MIN; return false; }
#pragma unroll
    for (int a = 0; a < 3; ++a) id[a] = (int)rintf(q[a]);                  // tf.round: half-to-even (:361)
    return id[0] >= 0 && id[0] < p.X && id[1] >= 0 && id[1] < p.Y && id[2] >= 0 && id[2] < p.Z;
}

// source offset (in floats) of a voxel inside this rank's slab, or -1
__device__ __forceinline__ long long slab_offset(const ProjParams& p, int b, const int id[3], bool valid) {
    if (!valid) return -1;
    const int xs = id[0] - p.x_begin;
    if (xs < 0 || xs >= p.Xs) return -1;
    return ((((long long)b * p.Xs + xs) * p.Y + id[1]) * p.Z + id[2]) * p.C;
}

__global__ void __launch_bounds__(K3_THREADS)
project_rays_kernel(const __grid_constant__ ProjParams p) {
    __shared__ SceneXf xf;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) build_scene(p, b, &xf);
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int per_scene = p.S * p.ph * p.pw;
    const int t = blockIdx.x * K3_THREADS + threadIdx.x;          // sample index within the scene
    __syncthreads();
    long long src = -1;
    if (t < per_scene) {
        const int j = t % p.pw;
        const int i = (t / p.pw) % p.ph;
        const int s = t / (p.pw * p.ph);
        int id[3];
        const bool valid = sample_voxel(p, xf, s, i, j, id);
        src = slab_offset(p, b, id, valid);
        const size_t o = (size_t)b * per_scene + t;
        if (p.out_vox) { p.out_vox[o * 3 + 0] = id[0]; p.out_vox[o * 3 + 1] = id[1]; p.out_vox[o * 3 + 2] = id[2]; }
        if (p.out_valid) p.out_valid[o] = (uint8_t)valid;
    }
    const int t0 = t - lane;
    const int C4 = p.C >> 2;
    for (int jj = 0; jj < 32; ++jj) {
        const long long sj = __shfl_sync(FULL, src, jj);
        const int tj = t0 + jj;
        if (tj >= per_scene) break;                               // warp-uniform
        float* o = p.out + ((size_t)b * per_scene + tj) * p.C;
        if (sj >= 0) {
            const float* g = p.grid + sj;
            for (int c4 = lane; c4 < C4; c4 += 32)
                stcs4(o + 4 * c4, ldg4(g + 4 * c4));
        } else {
            for (int c4 = lane; c4 < C4; c4 += 32) stcs4(o + 4 * c4, zero4());
        }
    }
}

// proj_grid + depth_sampling: one warp per pixel, lanes = float4 channel slots,
// out = act(bn_scale * (sum_s w_s * sample_s + bias) + bn_shift)
template <int CPL>
__global__ void __launch_bounds__(K3_THREADS)
project_collapse_kernel(const __grid_constant__ ProjParams p) {
    __shared__ SceneXf xf;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) build_scene(p, b, &xf);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    const int npix = p.ph * p.pw;
    const int pix = blockIdx.x * (K3_THREADS / 32) + warp;
    if (pix >= npix) return;
    const int i = pix / p.pw, j = pix % p.pw;
    const int C4 = p.C >> 2;
    float4 acc[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[k] = zero4();
    for (int s0 = 0; s0 < p.S; s0 += 32) {
        long long src = -1;
        float ws = 0.f;
        if (s0 + lane < p.S) {
            int id[3];
            const bool valid = sample_voxel(p, xf, s0 + lane, i, j, id);
            src = slab_offset(p, b, id, valid);
            ws = __ldg(p.w + s0 + lane);
        }
        const int smax = min(32, p.S - s0);
        for (int s = 0; s < smax; ++s) {
            const long long sj = __shfl_sync(FULL, src, s);
            const float wj = __shfl_sync(FULL, ws, s);
            if (sj < 0) continue;
            const float* g = p.grid + sj;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int c4 = lane + 32 * k;
                if (c4 < C4) acc[k] = fma4(wj, ldg4(g + 4 * c4), acc[k]);
            }
        }
    }
    float* o = p.out + ((size_t)b * npix + pix) * p.C;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int c4 = lane + 32 * k;
        if (c4 < C4) {
            float4 r = acc[k];
            r = make_float4(r.x + p.bias, r.y + p.bias, r.z + p.bias, r.w + p.bias);
            r = make_float4(fmaf(r.x, p.bn_scale, p.bn_shift), fmaf(r.y, p.bn_scale, p.bn_shift),
                            fmaf(r.z, p.bn_scale, p.bn_shift), fmaf(r.w, p.bn_scale, p.bn_shift));
            if (p.flags & MVF_FLAG_RELU_OUT) r = relu4(r);
            stcs4(o + 4 * c4, r);
        }
    }
}

// depth_sampling on a materialised tensor: in [B,S,npix,C] -> out [B,npix,C]
__global__ void __launch_bounds__(256)
depth_collapse_kernel(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                      int S, long long npix, int C4, float bias, float bn_scale, float bn_shift, int relu) {
    const long long per = npix * C4;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (e >= per) return;
    const float4* src = reinterpret_cast<const float4*>(in) + (long long)b * S * per + e;
    float4 acc = zero4();
    for (int s = 0; s < S; ++s) acc = fma4(__ldg(w + s), __ldg(src + (long long)s * per), acc);
    acc = make_float4(acc.x + bias, acc.y + bias, acc.z + bias, acc.w + bias);
    acc = make_float4(fmaf(acc.x, bn_scale, bn_shift), fmaf(acc.y, bn_scale, bn_shift),
                      fmaf(acc.z, bn_scale, bn_shift), fmaf(acc.w, bn_scale, bn_shift));
    if (relu) acc = relu4(acc);
    __stcs(reinterpret_cast<float4*>(out) + (long long)b * per + e, acc);
}

static int fill_proj_params(ProjParams& p, const float* grid, const float* Rview, const float* Rmain, const float* Kmat,
                            const float* grid_pos, const MvfGrid* g, int B, int C, int img_h, int proj_h, int proj_w,
                            int samples, int flags, double grid_dist, int x_begin, int x_count) {
    if (!Rview || !Kmat || !g) return MVF_ENULL;
    if (!grid && x_count != 0) return MVF_ENULL;                          // an empty slab has no grid memory
    if (B <= 0 || C <= 0 || img_h <= 0 || proj_h <= 0 || proj_w <= 0 || samples <= 0) return MVF_EINVAL;
    if (C % 4 != 0 || (grid && !aligned16(grid))) return MVF_EALIGN;
    if (samples > MVF_MAX_SAMPLES || g->nvox > MVF_MAX_DIM || g->nvox_z > MVF_MAX_DIM || B > 65535 || C > 1024) return MVF_EUNSUPPORTED;
    if ((flags & MVF_FLAG_WORLD_GRID) && !grid_pos) return MVF_ENULL;
    if (x_count < 0) { x_begin = 0; x_count = g->nvox; }                  // MVF_WHOLE_GRID; x_count == 0: every sample misses the slab
    if (x_begin < 0 || x_begin + x_count > g->nvox) return MVF_EINVAL;
    p.grid = grid; p.Rview = Rview; p.Rmain = Rmain; p.Kmat = Kmat;
    p.grid_pos = (flags & MVF_FLAG_WORLD_GRID) ? grid_pos : nullptr;
    p.w = nullptr; p.out = nullptr; p.out_vox = nullptr; p.out_valid = nullptr;
    p.B = B; p.C = C; p.ph = proj_h; p.pw = proj_w; p.S = samples;
    p.X = g->nvox; p.Y = g->nvox; p.Z = g->nvox_z; p.x_begin = x_begin; p.Xs = x_count; p.flags = flags;
    p.r = (float)((double)proj_h / (double)img_h);                        // :238
    if (flags & MVF_FLAG_WORLD_GRID) {                                    // Notebook/projection.py:293,307-308
        tf1_linspace(grid_dist - g->vmax * 0.8, grid_dist + g->vmax * 0.8, samples, p.zs);
        p.lo[0] = p.lo[1] = (float)g->vmin; p.lo[2] = (float)(-g->nvox_z * 0.5 * g->vsize);
        p.hi[0] = p.hi[1] = (float)g->vmax; p.hi[2] = (float)(g->nvox_z * 0.5 * g->vsize);
    } else {                                                              // model_multi.py:267,294-295
        tf1_linspace(g->vmin_z + g->vsize_z / 2.0, g->vmax_z - g->vsize_z / 2.0, samples, p.zs);
        p.lo[0] = p.lo[1] = (float)g->vmin; p.lo[2] = (float)(g->vmin_z + g->vsize_z / 2.0);
        p.hi[0] = p.hi[1] = (float)g->vmax; p.hi[2] = (float)g->vmax_z;
    }
    p.n[0] = p.n[1] = (float)(g->nvox * 1.0); p.n[2] = (float)(g->nvox_z * 1.0);   // :296
    p.bias = 0.f; p.bn_scale = 1.f; p.bn_shift = 0.f; p.rv_stride = 12;
    return MVF_OK;
}

}  // namespace mvf

using namespace mvf;

// rview_stride: floats between two scenes' poses in Rview -- the pipelines of api.cu project into view 0 of Rcam [B,V,3,4] in place
// (stride V*12) instead of gathering the main-view poses into a dense [B,3,4] first.
namespace mvf {
int project_rays_strided(const float* grid, const float* Rview, int rview_stride, const float* Kmat, const MvfGrid* g, int B, int C,
                         int img_h, int proj_h, int proj_w, int samples, float* out, void* stream) {
    ProjParams p;
    int rc = fill_proj_params(p, grid, Rview, nullptr, Kmat, nullptr, g, B, C, img_h, proj_h, proj_w, samples, 0, 0.0, 0, MVF_WHOLE_GRID);
    if (rc != MVF_OK) return rc;
    if (!out) return MVF_ENULL;
    if (!aligned16(out)) return MVF_EALIGN;
    p.out = out; p.rv_stride = rview_stride;
    dim3 grd((samples * proj_h * proj_w + K3_THREADS - 1) / K3_THREADS, B);
    project_rays_kernel<<<grd, K3_THREADS, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return check_launch();
}
}  // namespace mvf

extern "C" int mvf_project_rays(const float* grid, const float* Rview, const float* Rmain, const float* Kmat,
                                const float* grid_pos, const MvfGrid* g, int B, int C, int img_h,
                                int proj_h, int proj_w, int samples, int flags, double grid_dist,
                                int x_begin, int x_count,
                                float* out, int32_t* out_vox, uint8_t* out_valid, void* stream) {
    ProjParams p;
    int rc = fill_proj_params(p, grid, Rview, Rmain, Kmat, grid_pos, g, B, C, img_h, proj_h, proj_w, samples, flags,
                              grid_dist, x_begin, x_count);
    if (rc != MVF_OK) return rc;
    if (!out) return MVF_ENULL;
    if (!aligned16(out)) return MVF_EALIGN;
    p.out = out; p.out_vox = out_vox; p.out_valid = out_valid;
    const int per_scene = samples * proj_h * proj_w;
    dim3 grd((per_scene + K3_THREADS - 1) / K3_THREADS, B);
    project_rays_kernel<<<grd, K3_THREADS, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return check_launch();
}

static int collapse_launch(ProjParams& p, int B, int C, int proj_h, int proj_w, const float* w, float bias, float bn_scale,
                           float bn_shift, float* out, void* stream);

namespace mvf {
int project_collapse_strided(const float* grid, const float* Rview, int rview_stride, const float* Kmat, const MvfGrid* g, int B, int C,
                             int img_h, int proj_h, int proj_w, int samples, int flags, const float* w, float bias, float bn_scale,
                             float bn_shift, float* out, void* stream) {
    ProjParams p;
    int rc = fill_proj_params(p, grid, Rview, nullptr, Kmat, nullptr, g, B, C, img_h, proj_h, proj_w, samples, flags, 0.0, 0, MVF_WHOLE_GRID);
    if (rc != MVF_OK) return rc;
    p.rv_stride = rview_stride;
    return collapse_launch(p, B, C, proj_h, proj_w, w, bias, bn_scale, bn_shift, out, stream);
}
}  // namespace mvf

extern "C" int mvf_project_depth_collapse(const float* grid, const float* Rview, const float* Rmain,
                                          const float* Kmat, const float* grid_pos, const MvfGrid* g,
                                          int B, int C, int img_h, int proj_h, int proj_w, int samples,
                                          int flags, double grid_dist, int x_begin, int x_count,
                                          const float* w, float bias, float bn_scale, float bn_shift,
                                          float* out, void* stream) {
    ProjParams p;
    int rc = fill_proj_params(p, grid, Rview, Rmain, Kmat, grid_pos, g, B, C, img_h, proj_h, proj_w, samples, flags,
                              grid_dist, x_begin, x_count);
    if (rc != MVF_OK) return rc;
    return collapse_launch(p, B, C, proj_h, proj_w, w, bias, bn_scale, bn_shift, out, stream);
}

static int collapse_launch(ProjParams& p, int B, int C, int proj_h, int proj_w, const float* w, float bias, float bn_scale,
                           float bn_shift, float* out, void* stream) {
    if (!out || !w) return MVF_ENULL;
    if (!aligned16(out)) return MVF_EALIGN;
    p.out = out; p.w = w; p.bias = bias; p.bn_scale = bn_scale; p.bn_shift = bn_shift;
    const int npix = proj_h * proj_w;
    dim3 grd((npix + K3_THREADS / 32 - 1) / (K3_THREADS / 32), B);
    cudaStream_t s = (cudaStream_t)stream;
    const int C4 = C / 4;
    if (C4 <= 32) project_collapse_kernel<1><<<grd, K3_THREADS, 0, s>>>(p);
    else if (C4 <= 64) project_collapse_kernel<2><<<grd, K3_THREADS, 0, s>>>(p);
    else if (C4 <= 128) project_collapse_kernel<4><<<grd, K3_THREADS, 0, s>>>(p);
    else project_collapse_kernel<8><<<grd, K3_THREADS, 0, s>>>(p);
    count_launch();
    return check_launch();
}

extern "C" int mvf_depth_collapse(const float* in, int B, int S, long long npix, int C, const float* w,
                                  float bias, float bn_scale, float bn_shift, int flags,
                                  float* out, void* stream) {
    if (!in || !w || !out) return MVF_ENULL;
    if (B <= 0 || S <= 0 || npix <= 0 || C <= 0) return MVF_EINVAL;
    if (C % 4 != 0 || !aligned16(in) || !aligned16(out)) return MVF_EALIGN;
    if (B > 65535) return MVF_EUNSUPPORTED;
    const long long per = npix * (C / 4);
    dim3 grd((unsigned)((per + 255) / 256), B);
    depth_collapse_kernel<<<grd, 256, 0, (cudaStream_t)stream>>>(in, w, out, S, npix, C / 4, bias, bn_scale, bn_shift,
                                                                  (flags & MVF_FLAG_RELU_OUT) != 0);
    count_launch();
    return check_launch();
}
