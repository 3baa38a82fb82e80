// Library-level entry points: error strings, launch counter and the HOST-buffer pipeline.
#include "mvf_common.cuh"
#include <new>

namespace mvf {
std::atomic<unsigned long long> g_launches{0};
static size_t align_up256(size_t v) { return (v + 255) & ~(size_t)255; }
int project_rays_strided(const float* grid, const float* Rview, int rview_stride, const float* Kmat, const MvfGrid* g, int B, int C,
                         int img_h, int proj_h, int proj_w, int samples, float* out, void* stream);             // project.cu
int k1t_launch(const float* feats, const float* Rcam, const float* Rmain, const float* Kmat, const MvfGrid* g, int B, int V, int fh, int fw,
               int C, int img_h, int img_w, int mode, int flags, double grid_dist, int x_begin, int x_count, const float* bn_scale,
               const float* bn_shift, float* out, void* ws, size_t ws_bytes, void* stream);                                     // unproject_tc.cu
int project_collapse_strided(const float* grid, const float* Rview, int rview_stride, const float* Kmat, const MvfGrid* g, int B, int C,
                             int img_h, int proj_h, int proj_w, int samples, int flags, const float* w, float bias, float bn_scale,
                             float bn_shift, float* out, void* stream);                                           // project.cu
}  // namespace mvf

using namespace mvf;

extern "C" const char* mvf_error_string(int code) {
    switch (code) {
        case MVF_OK: return "ok";
        case MVF_EINVAL: return "invalid shape, size or enum value";
        case MVF_ENULL: return "required pointer is NULL";
        case MVF_EALIGN: return "channel count not a multiple of 4 or pointer not 16-byte aligned";
        case MVF_ECUDA: return "CUDA runtime call or kernel launch failed";
        case MVF_EUNSUPPORTED: return "request outside the compiled limits";
        case MVF_EWORKSPACE: return "workspace too small";
        default: return "unknown error";
    }
}

extern "C" const char* mvf_version(void) { return "mvfusion 0.1.0 (sm_100a)"; }

extern "C" unsigned long long mvf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// device scratch layout of the host pipeline: feats | Rcam | Kmat | fused grid | ray slices
struct HostWs { float *feats, *Rcam, *Kmat, *grid, *out; char* k1t; size_t k1t_scene, bytes; };

static HostWs carve_host(void* ws, const MvfGrid* g, int B, int V, int fh, int fw, int C, int ph, int pw, int S) {
    HostWs w;
    char* p = (char*)ws;
    size_t off = 0;
    w.feats = (float*)(p + off); off += align_up256((size_t)B * V * fh * fw * C * sizeof(float));
    w.Rcam = (float*)(p + off); off += align_up256((size_t)B * V * 12 * sizeof(float));
    w.Kmat = (float*)(p + off); off += align_up256((size_t)B * 9 * sizeof(float));
    w.grid = (float*)(p + off); off += align_up256((size_t)B * g->nvox * g->nvox * g->nvox_z * C * sizeof(float));
    w.out = (float*)(p + off); off += align_up256((size_t)B * S * ph * pw * C * sizeof(float));
    // scratch of the tensor-core unprojection (fp16 operand halves of one scene's features), one region per scene
    w.k1t_scene = align_up256(mvf_unproject_fuse_tc_workspace_bytes(1, V, fh, fw, C));
    w.k1t = p + off; off += (size_t)B * w.k1t_scene;
    w.bytes = off;
    return w;
}

extern "C" size_t mvf_pipeline_host_workspace_bytes(const MvfGrid* g, int B, int V, int fh, int fw, int C,
                                                    int proj_h, int proj_w, int samples) {
    if (!g || B <= 0 || V <= 0 || fh <= 0 || fw <= 0 || C <= 0 || proj_h <= 0 || proj_w <= 0 || samples <= 0) return 0;
    return carve_host(nullptr, g, B, V, fh, fw, C, proj_h, proj_w, samples).bytes;
}

// Auxiliary streams/events of the host entry points live in a caller-owned handle (mvf_host_aux_create): copy-in runs on
// `sh`, kernels on `sc`, copy-out on `sd`, so the H2D of scene-chunk i+1, the kernels of chunk i and the D2H of chunk i-1
// overlap.  The library itself keeps no streams, events or per-device state.
namespace {
constexpr int kMaxChunks = 32;
}
struct MvfHostAux {
    int device = -1;
    cudaStream_t sh = nullptr, sc = nullptr, sd = nullptr;
    cudaEvent_t start = nullptr, done_c = nullptr, done_d = nullptr;
    cudaEvent_t ev[kMaxChunks] = {}, ev_in[kMaxChunks + 1] = {};
};

extern "C" int mvf_host_aux_destroy(MvfHostAux* a) {
    if (!a) return MVF_OK;
    if (a->sh) { cudaStreamSynchronize(a->sh); cudaStreamDestroy(a->sh); }
    if (a->sc) { cudaStreamSynchronize(a->sc); cudaStreamDestroy(a->sc); }
    if (a->sd) { cudaStreamSynchronize(a->sd); cudaStreamDestroy(a->sd); }
    if (a->start) cudaEventDestroy(a->start);
    if (a->done_c) cudaEventDestroy(a->done_c);
    if (a->done_d) cudaEventDestroy(a->done_d);
    for (int i = 0; i < kMaxChunks; ++i) if (a->ev[i]) cudaEventDestroy(a->ev[i]);
    for (int i = 0; i <= kMaxChunks; ++i) if (a->ev_in[i]) cudaEventDestroy(a->ev_in[i]);
    delete a;
    return MVF_OK;
}

extern "C" int mvf_host_aux_create(MvfHostAux** out) {
    if (!out) return MVF_ENULL;
    *out = nullptr;
    MvfHostAux* a = new (std::nothrow) MvfHostAux();
    if (!a) return MVF_EINVAL;
    bool ok = cudaGetDevice(&a->device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&a->sh, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&a->sc, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&a->sd, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&a->start, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&a->done_c, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&a->done_d, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < kMaxChunks; ++i) ok = cudaEventCreateWithFlags(&a->ev[i], cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i <= kMaxChunks; ++i) ok = cudaEventCreateWithFlags(&a->ev_in[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { mvf_host_aux_destroy(a); return MVF_ECUDA; }       // nothing leaks on a partial creation
    *out = a;
    return MVF_OK;
}

// The reference crosses host->device once per predict() call (mrcnn/model_multi.py:3067-3068);
// this is the same crossing for the fusion path alone: H2D inputs, K1, K3, D2H ray slices,
// software-pipelined over chunks of scenes.
namespace {
struct Collapse { const float* d_w; float bias, bn_scale, bn_shift; };       // depth_sampling fused into the projection (K3b)
}

static int host_pipeline(const float* h_feats, const float* h_Rcam, const float* h_Kmat,
                         const MvfGrid* g, int B, int V, int fh, int fw, int C,
                         int img_h, int img_w, int mode, int flags,
                         const float* d_bn_scale, const float* d_bn_shift,
                         int proj_h, int proj_w, int samples, const Collapse* col,
                         float* h_out, void* dev_ws, size_t dev_ws_bytes, MvfHostAux* ax, void* stream) {
    if (!h_feats || !h_Rcam || !h_Kmat || !g || !h_out || !dev_ws || !ax) return MVF_ENULL;
    if (B <= 0 || V <= 0 || fh <= 0 || fw <= 0 || C <= 0 || proj_h <= 0 || proj_w <= 0 || samples <= 0) return MVF_EINVAL;
    if (mode < MVF_FUSE_SUM || mode > MVF_FUSE_MAX) return MVF_EINVAL;
    if (!aligned16(dev_ws)) return MVF_EALIGN;
    HostWs w = carve_host(dev_ws, g, B, V, fh, fw, C, proj_h, proj_w, samples);
    if (dev_ws_bytes < w.bytes) return MVF_EWORKSPACE;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return MVF_ECUDA;
    if (dev != ax->device) return MVF_EINVAL;                     // the handle's streams belong to another device
    cudaStream_t s = (cudaStream_t)stream;
    const size_t feat_scene = (size_t)V * fh * fw * C;
    const size_t grid_scene = (size_t)g->nvox * g->nvox * g->nvox_z * C;
    const size_t out_scene = (size_t)(col ? 1 : samples) * proj_h * proj_w * C;      // collapsed: [P,P,C] per scene
    // On any failure after work has been queued, drain the three auxiliary streams before returning: the caller owns the
    // host buffers they read and write, and the next call on this handle must not find stale work in front of it.
    auto fail = [&](int code) { cudaStreamSynchronize(ax->sh); cudaStreamSynchronize(ax->sc); cudaStreamSynchronize(ax->sd); return code; };
#define MVF_TRY(x) do { if ((x) != cudaSuccess) return fail(MVF_ECUDA); } while (0)
    // order the auxiliary streams after whatever the caller queued on `stream`
    MVF_TRY(cudaEventRecord(ax->start, s));
    MVF_TRY(cudaStreamWaitEvent(ax->sh, ax->start, 0));
    MVF_TRY(cudaStreamWaitEvent(ax->sc, ax->start, 0));
    MVF_TRY(cudaStreamWaitEvent(ax->sd, ax->start, 0));
    MVF_TRY(cudaMemcpyAsync(w.Rcam, h_Rcam, (size_t)B * V * 12 * sizeof(float), cudaMemcpyHostToDevice, ax->sh));
    MVF_TRY(cudaMemcpyAsync(w.Kmat, h_Kmat, (size_t)B * 9 * sizeof(float), cudaMemcpyHostToDevice, ax->sh));
    // (proj_grid projects into view 0, Rcam[:,0], model_multi.py:245: K3 reads those poses in place with the scene stride V*12 --
    //  a strided 2-D device copy into a dense [B,3,4] measured 0.1-0.2 ms here, as much as K3 itself)
    MVF_TRY(cudaEventRecord(ax->ev_in[kMaxChunks], ax->sh));
    MVF_TRY(cudaStreamWaitEvent(ax->sc, ax->ev_in[kMaxChunks], 0));
    const int per = (B + kMaxChunks - 1) / kMaxChunks;          // scenes per chunk (1 unless B > 32)
    int chunk = 0;
    for (int b0 = 0; b0 < B; b0 += per, ++chunk) {
        const int nb = (b0 + per <= B) ? per : (B - b0);
        MVF_TRY(cudaMemcpyAsync(w.feats + b0 * feat_scene, h_feats + b0 * feat_scene, nb * feat_scene * sizeof(float),
                                cudaMemcpyHostToDevice, ax->sh));
        MVF_TRY(cudaEventRecord(ax->ev_in[chunk], ax->sh));
        MVF_TRY(cudaStreamWaitEvent(ax->sc, ax->ev_in[chunk], 0));
        // K1T (tensor cores) for the linear reductions it supports, else the CUDA-core slot kernel -- the same choice the
        // Python mirror makes, so both entries return the same bits
        int rc;
        if (mvf_unproject_fuse_tc_supported(V, C, mode, flags & ~MVF_FLAG_WORLD_GRID))
            rc = mvf_unproject_fuse_tc(w.feats + b0 * feat_scene, w.Rcam + (size_t)b0 * V * 12, nullptr, w.Kmat + (size_t)b0 * 9, g,
                                       nb, V, fh, fw, C, img_h, img_w, mode, flags & ~MVF_FLAG_WORLD_GRID, 0.0, 0, MVF_WHOLE_GRID,
                                       d_bn_scale, d_bn_shift, w.grid + b0 * grid_scene, w.k1t + (size_t)b0 * w.k1t_scene,
                                       (size_t)nb * w.k1t_scene, ax->sc);
        else
            rc = mvf_unproject_fuse(w.feats + b0 * feat_scene, w.Rcam + (size_t)b0 * V * 12, nullptr, w.Kmat + (size_t)b0 * 9, g,
                                    nb, V, fh, fw, C, img_h, img_w, mode, flags & ~MVF_FLAG_WORLD_GRID, 0.0, 0, MVF_WHOLE_GRID,
                                    d_bn_scale, d_bn_shift, w.grid + b0 * grid_scene, nullptr, nullptr, nullptr, ax->sc);
        if (rc != MVF_OK) return fail(rc);
        if (col)
            rc = project_collapse_strided(w.grid + b0 * grid_scene, w.Rcam + (size_t)b0 * V * 12, V * 12, w.Kmat + (size_t)b0 * 9, g,
                                          nb, C, img_h, proj_h, proj_w, samples, MVF_FLAG_RELU_OUT,
                                          col->d_w, col->bias, col->bn_scale, col->bn_shift, w.out + b0 * out_scene, ax->sc);
        else
            rc = project_rays_strided(w.grid + b0 * grid_scene, w.Rcam + (size_t)b0 * V * 12, V * 12, w.Kmat + (size_t)b0 * 9, g,
                                      nb, C, img_h, proj_h, proj_w, samples, w.out + b0 * out_scene, ax->sc);
        if (rc != MVF_OK) return fail(rc);
        MVF_TRY(cudaEventRecord(ax->ev[chunk], ax->sc));
        MVF_TRY(cudaStreamWaitEvent(ax->sd, ax->ev[chunk], 0));
        MVF_TRY(cudaMemcpyAsync(h_out + b0 * out_scene, w.out + b0 * out_scene, nb * out_scene * sizeof(float),
                                cudaMemcpyDeviceToHost, ax->sd));
    }
    MVF_TRY(cudaEventRecord(ax->done_c, ax->sc));
    MVF_TRY(cudaEventRecord(ax->done_d, ax->sd));
    MVF_TRY(cudaStreamWaitEvent(s, ax->done_c, 0));
    MVF_TRY(cudaStreamWaitEvent(s, ax->done_d, 0));
    MVF_TRY(cudaStreamSynchronize(s));
#undef MVF_TRY
    return MVF_OK;
}

extern "C" int mvf_unproject_fuse_project_host(const float* h_feats, const float* h_Rcam, const float* h_Kmat,
                                               const MvfGrid* g, int B, int V, int fh, int fw, int C,
                                               int img_h, int img_w, int mode, int flags,
                                               const float* d_bn_scale, const float* d_bn_shift,
                                               int proj_h, int proj_w, int samples,
                                               float* h_out, void* dev_ws, size_t dev_ws_bytes, MvfHostAux* aux, void* stream) {
    return host_pipeline(h_feats, h_Rcam, h_Kmat, g, B, V, fh, fw, C, img_h, img_w, mode, flags, d_bn_scale, d_bn_shift,
                         proj_h, proj_w, samples, nullptr, h_out, dev_ws, dev_ws_bytes, aux, stream);
}

// ---- the fused pipeline on DEVICE buffers: one call, one stream ----------------------------------------------------------------
// unproj_feat -> grid_reas -> proj_grid for a batch of scenes: K1T with its feature split running under it (unproject_tc.cu), or the
// slot kernel for the modes K1T does not take, then K3 reading the main-view poses in place.  (Measured and not shipped in round 2:
// K3 as a programmatic dependent that follows the tensor-core kernel scene by scene through per-scene store counters -- correct, but
// the projection's loads and stores on the same SMs slowed unproject_tc_kernel by about as much as they hid: DESIGN.md section 3.1b.)
extern "C" size_t mvf_unproject_fuse_project_workspace_bytes(int B, int V, int fh, int fw, int C) {
    return mvf_unproject_fuse_tc_workspace_bytes(B, V, fh, fw, C);
}

extern "C" int mvf_unproject_fuse_project(const float* feats, const float* Rcam, const float* Kmat,
                                          const MvfGrid* g, int B, int V, int fh, int fw, int C,
                                          int img_h, int img_w, int mode, int flags,
                                          const float* bn_scale, const float* bn_shift,
                                          int proj_h, int proj_w, int samples,
                                          float* grid_out, float* rays_out, void* ws, size_t ws_bytes, void* stream) {
    if (!feats || !Rcam || !Kmat || !g || !grid_out || !rays_out) return MVF_ENULL;
    if (B <= 0 || V <= 0 || fh <= 0 || fw <= 0 || C <= 0 || proj_h <= 0 || proj_w <= 0 || samples <= 0) return MVF_EINVAL;
    if (mode < MVF_FUSE_SUM || mode > MVF_FUSE_MAX) return MVF_EINVAL;
    if (flags & MVF_FLAG_WORLD_GRID) return MVF_EUNSUPPORTED;      // (the notebook variant's projection needs grid_pos: use the two calls)
    int rc;
    if (mvf_unproject_fuse_tc_supported(V, C, mode, flags)) {
        if (!ws) return MVF_ENULL;
        rc = k1t_launch(feats, Rcam, nullptr, Kmat, g, B, V, fh, fw, C, img_h, img_w, mode, flags, 0.0, 0, MVF_WHOLE_GRID,
                        bn_scale, bn_shift, grid_out, ws, ws_bytes, stream);
    } else {                                                       // CUDA-core slot kernel, then the projection as an ordinary launch
        rc = mvf_unproject_fuse(feats, Rcam, nullptr, Kmat, g, B, V, fh, fw, C, img_h, img_w, mode, flags, 0.0, 0, MVF_WHOLE_GRID,
                                bn_scale, bn_shift, grid_out, nullptr, nullptr, nullptr, stream);
    }
    if (rc != MVF_OK) return rc;
    // proj_grid projects into view 0 (Rcam[:,0], model_multi.py:245): read in place with the scene stride V*12
    return project_rays_strided(grid_out, Rcam, V * 12, Kmat, g, B, C, img_h, proj_h, proj_w, samples, rays_out, stream);
}

// One pyramid level of the fusion neck (model_multi.py:2382-2404) from HOST buffers: unproj_feat -> grid_reas(sum|mean|max
// [+BN+ReLU]) -> proj_grid -> depth_sampling (non-conv3d branch), i.e. features in, PG [B,P,P,C] out -- neither the fused
// grid nor the ray slices cross the bus.  d_depth_w [S] device; the scalars are the folded depth conv bias / BN.
extern "C" int mvf_fusion_neck_level_host(const float* h_feats, const float* h_Rcam, const float* h_Kmat,
                                          const MvfGrid* g, int B, int V, int fh, int fw, int C,
                                          int img_h, int img_w, int mode, int flags,
                                          const float* d_bn_scale, const float* d_bn_shift,
                                          int proj_h, int proj_w, int samples,
                                          const float* d_depth_w, float depth_bias, float depth_bn_scale, float depth_bn_shift,
                                          float* h_out, void* dev_ws, size_t dev_ws_bytes, MvfHostAux* aux, void* stream) {
    if (!d_depth_w) return MVF_ENULL;
    const Collapse col = {d_depth_w, depth_bias, depth_bn_scale, depth_bn_shift};
    return host_pipeline(h_feats, h_Rcam, h_Kmat, g, B, V, fh, fw, C, img_h, img_w, mode, flags, d_bn_scale, d_bn_shift,
                         proj_h, proj_w, samples, &col, h_out, dev_ws, dev_ws_bytes, aux, stream);
}
