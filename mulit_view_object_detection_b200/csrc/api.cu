// Library-level entry points: error strings, launch counter and the HOST-buffer pipeline.
#include "mvf_common.cuh"

namespace mvf {
std::atomic<unsigned long long> g_launches{0};
static size_t align_up256(size_t v) { return (v + 255) & ~(size_t)255; }
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
struct HostWs { float *feats, *Rcam, *R0, *Kmat, *grid, *out; size_t bytes; };

static HostWs carve_host(void* ws, const MvfGrid* g, int B, int V, int fh, int fw, int C, int ph, int pw, int S) {
    HostWs w;
    char* p = (char*)ws;
    size_t off = 0;
    w.feats = (float*)(p + off); off += align_up256((size_t)B * V * fh * fw * C * sizeof(float));
    w.Rcam = (float*)(p + off); off += align_up256((size_t)B * V * 12 * sizeof(float));
    w.R0 = (float*)(p + off); off += align_up256((size_t)B * 12 * sizeof(float));
    w.Kmat = (float*)(p + off); off += align_up256((size_t)B * 9 * sizeof(float));
    w.grid = (float*)(p + off); off += align_up256((size_t)B * g->nvox * g->nvox * g->nvox_z * C * sizeof(float));
    w.out = (float*)(p + off); off += align_up256((size_t)B * S * ph * pw * C * sizeof(float));
    w.bytes = off;
    return w;
}

extern "C" size_t mvf_pipeline_host_workspace_bytes(const MvfGrid* g, int B, int V, int fh, int fw, int C,
                                                    int proj_h, int proj_w, int samples) {
    if (!g || B <= 0 || V <= 0 || fh <= 0 || fw <= 0 || C <= 0 || proj_h <= 0 || proj_w <= 0 || samples <= 0) return 0;
    return carve_host(nullptr, g, B, V, fh, fw, C, proj_h, proj_w, samples).bytes;
}

// The reference crosses host->device once per predict() call (mrcnn/model_multi.py:3067-3068);
// this is the same crossing for the fusion path alone: H2D inputs, K1, K3, D2H ray slices.
extern "C" int mvf_unproject_fuse_project_host(const float* h_feats, const float* h_Rcam, const float* h_Kmat,
                                               const MvfGrid* g, int B, int V, int fh, int fw, int C,
                                               int img_h, int img_w, int mode, int flags,
                                               const float* d_bn_scale, const float* d_bn_shift,
                                               int proj_h, int proj_w, int samples,
                                               float* h_out, void* dev_ws, size_t dev_ws_bytes, void* stream) {
    if (!h_feats || !h_Rcam || !h_Kmat || !g || !h_out || !dev_ws) return MVF_ENULL;
    if (B <= 0 || V <= 0 || fh <= 0 || fw <= 0 || C <= 0 || proj_h <= 0 || proj_w <= 0 || samples <= 0) return MVF_EINVAL;
    if (mode < MVF_FUSE_SUM || mode > MVF_FUSE_MAX) return MVF_EINVAL;
    if (!aligned16(dev_ws)) return MVF_EALIGN;
    HostWs w = carve_host(dev_ws, g, B, V, fh, fw, C, proj_h, proj_w, samples);
    if (dev_ws_bytes < w.bytes) return MVF_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t feat_bytes = (size_t)B * V * fh * fw * C * sizeof(float);
    const size_t out_bytes = (size_t)B * samples * proj_h * proj_w * C * sizeof(float);
    if (cudaMemcpyAsync(w.feats, h_feats, feat_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) return MVF_ECUDA;
    if (cudaMemcpyAsync(w.Rcam, h_Rcam, (size_t)B * V * 12 * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess) return MVF_ECUDA;
    if (cudaMemcpyAsync(w.Kmat, h_Kmat, (size_t)B * 9 * sizeof(float), cudaMemcpyHostToDevice, s) != cudaSuccess) return MVF_ECUDA;
    int rc = mvf_unproject_fuse(w.feats, w.Rcam, nullptr, w.Kmat, g, B, V, fh, fw, C, img_h, img_w, mode, flags, 0.0, 0, 0,
                                d_bn_scale, d_bn_shift, w.grid, nullptr, nullptr, nullptr, stream);
    if (rc != MVF_OK) return rc;
    // proj_grid reads the main-view poses as a contiguous [B,3,4] tensor (Rcam[:,0], model_multi.py:245):
    // compact them out of [B,V,3,4] with one strided device copy.
    if (cudaMemcpy2DAsync(w.R0, 12 * sizeof(float), w.Rcam, (size_t)V * 12 * sizeof(float), 12 * sizeof(float), B,
                          cudaMemcpyDeviceToDevice, s) != cudaSuccess) return MVF_ECUDA;
    rc = mvf_project_rays(w.grid, w.R0, nullptr, w.Kmat, nullptr, g, B, C, img_h, proj_h, proj_w, samples,
                          flags & ~MVF_FLAG_WORLD_GRID, 0.0, 0, 0, w.out, nullptr, nullptr, stream);
    if (rc != MVF_OK) return rc;
    if (cudaMemcpyAsync(h_out, w.out, out_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess) return MVF_ECUDA;
    if (cudaStreamSynchronize(s) != cudaSuccess) return MVF_ECUDA;
    return MVF_OK;
}
