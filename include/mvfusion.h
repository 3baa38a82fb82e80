/*
 * mvfusion.h -- C ABI of libmvfusion.so: the B200 (sm_100a) implementation of the multi-view
 * fusion hot path of juliuserbach/mulit_view_object_detection
 *     unproject -> fuse across views -> project -> PyramidROIAlign -> per-class NMS.
 *
 * The reference is pure Python/TF1 (no FFI of its own); each entry point below replaces the
 * reference layer named in its comment (file:line relative to the reference root) and is
 * what a ctypes / cffi / pybind binding of that layer would call.  INTEGRATION.md shows the
 * reference-side stub for each.
 *
 * Conventions
 *  - All tensors are contiguous fp32, channel-last, in DEVICE memory unless a parameter is
 *    documented as host memory.  Channel counts must be multiples of 4 and tensor base
 *    addresses 16-byte aligned (128-bit vector access); violations return MVF_EALIGN.
 *  - The caller owns every buffer (inputs, outputs, workspaces).  The library never
 *    allocates or frees device memory, never synchronises the stream (except the *_host
 *    entry points, documented there) and keeps no state besides a launch counter: the copy
 *    streams and events of the *_host entry points live in a caller-owned MvfHostAux handle.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - Return value: MVF_OK (0) or a negative MVF_E* code; mvf_error_string() describes it.
 *  - Thread-safe and re-entrant per stream; the *_host entry points are re-entrant per MvfHostAux
 *    handle (one in-flight call per handle; any number of handles).
 */
#ifndef MVFUSION_H_
#define MVFUSION_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVF_OK            0
#define MVF_EINVAL       -1   /* bad shape / size / enum value                     */
#define MVF_ENULL        -2   /* required pointer is NULL                          */
#define MVF_EALIGN       -3   /* C % 4 != 0 or pointer not 16-byte aligned         */
#define MVF_ECUDA        -4   /* a CUDA runtime call or kernel launch failed       */
#define MVF_EUNSUPPORTED -5   /* valid request outside the compiled limits         */
#define MVF_EWORKSPACE   -6   /* workspace too small                               */

/* compiled limits */
#define MVF_MAX_VIEWS     32    /* views per scene handled by one mvf_unproject_fuse call */
#define MVF_MAX_DIM       192   /* voxels per grid axis                                    */
#define MVF_MAX_SAMPLES   64    /* depth samples per ray                                   */
#define MVF_MAX_NMS_BOXES 8192  /* candidates per NMS problem                              */
#define MVF_MAX_CLASSES   256
#define MVF_WHOLE_GRID    (-1)  /* x_count value meaning "no slab: the whole grid" */

/* Voxel box, the attributes the reference reads from `config`
 * (samples/interior/interior_multi.py:379-386; read at mrcnn/model_multi.py:157-160,267,294-296).
 * Doubles on purpose: the reference forms e.g. `vmin + vsize/2.0` in Python doubles before
 * TensorFlow casts to float32, and the library reproduces that order. */
typedef struct MvfGrid {
    int32_t nvox;      /* X = Y */
    int32_t nvox_z;    /* Z     */
    double  vmin, vmax, vsize;
    double  vmin_z, vmax_z, vsize_z;
} MvfGrid;

/* view reduction of mvf_unproject_fuse / mvf_view_reduce */
#define MVF_FUSE_NONE 0   /* no reduction: write the per-view grids [B,V,X,Y,Z,C] (= unproj_feat) */
#define MVF_FUSE_SUM  1   /* K.sum(axis=1), model_multi.py:402                                     */
#define MVF_FUSE_MEAN 2   /* sum * (1/V)                                                           */
#define MVF_FUSE_MAX  3   /* max over views of the zero-filled per-view samples                    */

/* flags */
#define MVF_FLAG_RELU_IN    1  /* ReLU on each per-view sample before the reduction (:448,:459) */
#define MVF_FLAG_RELU_OUT   2  /* ReLU after the (optional) BatchNorm affine (:404)             */
#define MVF_FLAG_PRESPLIT   8  /* mvf_conv3d_tc: `ws` already holds the operand halves (mvf_unproject_split_f16)   */
#define MVF_FLAG_WORLD_GRID 4  /* notebook variant: world-axis-aligned grid centred at
                                  [R0|t0].(0,0,grid_dist,1) (Notebook/projection.py:47-151,253-339) */

/* ---- K1: unproj_feat (+ fused grid_reas add/mean/max) --------------------------------------
 * replaces  unproj_feat([feats,Rcam,Kmat], config)   mrcnn/model_multi.py:130-228
 *      and  grid_reas(..)  'add' branch               mrcnn/model_multi.py:401-404
 * feats  [B,V,fh,fw,C]        Rcam [B,V,3,4] camera->world        Kmat [B,3,3]
 * Rmain  [B,3,4] or NULL: pose of the MAIN view (reference: Rcam[:,0]); pass it when `Rcam`
 *        holds only a shard of the views (multi-GPU view sharding).
 * x_begin/x_count: compute only the x-slab [x_begin, x_begin+x_count) of the grid; outputs are slab-shaped.
 *        x_count == MVF_WHOLE_GRID (any negative value) -> the whole grid.  x_count == 0 is a legitimate EMPTY slab
 *        (a sharded caller with more ranks than x-planes): nothing is written and MVF_OK is returned.
 * bn_scale/bn_shift [C] or NULL: frozen BatchNorm as x*scale+shift (applied when mode != NONE).
 * out       mode NONE: [B,V,Xs,Y,Z,C]   else [B,Xs,Y,Z,C]
 * out_idx   NULL or int32 [B,V,Xs,Y,Z,2] = (y0,x0) of the bilinear cell (INT32_MIN if unusable)
 * out_valid NULL or uint8 [B,V,Xs,Y,Z]: bit0=(y0,x0) bit1=(y1,x0) bit2=(y0,x1) bit3=(y1,x1)
 * out_grid_pos NULL or [B,3]: grid_position of the MVF_FLAG_WORLD_GRID variant. */
int mvf_unproject_fuse(const float* feats, const float* Rcam, const float* Rmain, const float* Kmat,
                       const MvfGrid* g, int B, int V, int fh, int fw, int C, int img_h, int img_w,
                       int mode, int flags, double grid_dist, int x_begin, int x_count,
                       const float* bn_scale, const float* bn_shift,
                       float* out, int32_t* out_idx, uint8_t* out_valid, float* out_grid_pos,
                       void* stream);

/* ---- K1T: the same fused unproject + view reduction on the tensor cores (tcgen05) ----------------------------------------
 * Same arguments and results as mvf_unproject_fuse (features within the fp32 tolerance) for the LINEAR view
 * reductions: mode MVF_FUSE_SUM | MVF_FUSE_MEAN, no MVF_FLAG_RELU_IN, C % 64 == 0, 64 <= C <= 256; no index / mask side
 * outputs.  mvf_unproject_fuse_tc_supported() tells whether a configuration qualifies (else MVF_EUNSUPPORTED: use
 * mvf_unproject_fuse).  ws: mvf_unproject_fuse_tc_workspace_bytes(...) bytes of device scratch (the fp16 operand halves of the
 * features).  Bilinear sampling of a 4x4x8 voxel tile is evaluated as out[128, C] += W_v[128, K] . F_v[K, C] per view with
 * the accumulators in TMEM; voxel->pixel coordinates and weights are computed exactly as in mvf_unproject_fuse. */
int mvf_unproject_fuse_tc_supported(int V, int C, int mode, int flags);
size_t mvf_unproject_fuse_tc_workspace_bytes(int B, int V, int fh, int fw, int C);
int mvf_unproject_fuse_tc(const float* feats, const float* Rcam, const float* Rmain, const float* Kmat,
                          const MvfGrid* g, int B, int V, int fh, int fw, int C, int img_h, int img_w,
                          int mode, int flags, double grid_dist, int x_begin, int x_count,
                          const float* bn_scale, const float* bn_shift, float* out,
                          void* ws, size_t ws_bytes, void* stream);

/* unproj_feat written straight into the operand format of the convolution that consumes it -- the 'conv3d' U-Net's first
 * convolution (model_multi.py:411-421; sublattices = 1: parity-sub-lattice layout of MVF_CONV_S2) or the 'ident' 1x1x1 conv
 * (:446-453; sublattices = 0) -- as fp16 (hi, lo) halves, into the workspace a following
 * mvf_conv3d_tc(flags | MVF_FLAG_PRESPLIT, in = NULL, same ws) reads: the fp32 per-view grids and the split pass over them never
 * exist.  The weights of that convolution are prepared with chan_interleave = -1 (fp16 format whatever the kernel size).
 * act_amax: DEVICE pointer to a bound on max|value| (max|feats| is one).  conv_ws: 4 * B*V*X*Y*Z*C + 256 bytes at least.
 * Needs C % 64 == 0 (and even grid dims with sublattices). */
int mvf_unproject_split_f16(const float* feats, const float* Rcam, const float* Rmain, const float* Kmat,
                            const MvfGrid* g, int B, int V, int fh, int fw, int C, int img_h, int img_w,
                            int flags, int sublattices, const float* act_amax, void* conv_ws, size_t ws_bytes, void* stream);

/* ---- grid_reas on a materialised [B,V,N,C] tensor -------------------------------------------
 * replaces grid_reas(x, scope, config) 'add' (model_multi.py:401-404) and the oracle-defined
 * mean / max.  N = voxels per scene.  in [B,V,N,C] -> out [B,N,C]. */
int mvf_view_reduce(const float* in, int B, int V, long long N, int C, int mode, int flags,
                    const float* bn_scale, const float* bn_shift, float* out, void* stream);

/* ---- notebook grid_reas 'mean' ---------------------------------------------------------------
 * replaces Notebook/projection.py:526-529,549: the mean is taken over the CHANNEL axis and the V per-view scalars
 * become the channels, then ReLU.  in [B,V,N,C] -> out [B,N,V]. */
int mvf_channel_mean(const float* in, int B, int V, long long N, int C, float* out, void* stream);

/* ---- grid_reas 'ident' -----------------------------------------------------------------------
 * replaces model_multi.py:443-455: ReLU -> concat views on channels (v*C+c) -> Conv3D 1x1x1
 * (+bias) -> BN -> ReLU.   in [B,V,N,C], weight [V*C,Cout], bias [Cout] -> out [B,N,Cout]. */
int mvf_ident_fuse(const float* in, const float* weight, const float* bias,
                   const float* bn_scale, const float* bn_shift,
                   int B, int V, long long N, int C, int Cout, float* out, void* stream);

/* ---- the conv3d family on the tensor cores (tcgen05, 3xTF32 split) ---------------------------------
 * One implicit-GEMM kernel serves every plain 3-D convolution of the fusion path:
 *   MVF_CONV_S1   Conv3D k=1|3 stride 1 SAME     grid_reas 'ident' (model_multi.py:443-455), depth_sampling 1x1 convs (:472-480)
 *   MVF_CONV_S2   Conv3D k=3 stride 2 SAME       grid_reas 'conv3d' encoder (:415-428); X,Y,Z must be even
 *   MVF_DECONV_S2 Conv3DTranspose k=3 stride 2   grid_reas 'conv3d' decoder (:430-441)
 * in  [B,V,X,Y,Z,C]: V tensors concatenated on channels (view-major, v*C + c) -- the reference's transpose+reshape
 *     of the per-view grids (:411-413) without moving data;  in2 [B,X,Y,Z,C2] or NULL: appended after them (the skip
 *     concat of :438).  X,Y,Z are the INPUT dims; out is [B,X,Y,Z,Cout] (S1), [B,X/2,Y/2,Z/2,Cout] (S2) or
 *     [B,2X,2Y,2Z,Cout] (DECONV).
 * out = act(bn(conv(pre(in)) + bias)); pre = optional per-input-channel affine pre_scale/pre_shift [V*C] (a depthwise
 *     1x1 conv, :472,:477) then ReLU when MVF_FLAG_RELU_IN; act = ReLU when MVF_FLAG_RELU_OUT; bn_* [Cout] or NULL.
 * Precision: k=3 members whose sources all have a multiple of 64 channels run the split as fp16 halves (a*2^s = a1 + a2,
 *     three f16 MMAs, twice the tf32 rate, ~5x closer to fp32 than the tf32 split); the others use the tf32 split.
 * Weights: mvf_conv3d_prepare(W) once per tensor (same V, C, C2 as the call; the prepared format depends on them).  W is the Keras kernel: Conv3D [k,k,k,Cin,Cout], Conv3DTranspose
 *     [3,3,3,Cout,Cin]; Cin = V*C + C2.  chan_interleave = S > 1: the reference orders the input channels (c*S + s)
 *     while `in` supplies S sources of C channels (depth_sampling, :468-470).
 * ws: mvf_conv3d_tc_workspace_bytes(...) bytes of device scratch for the hi/lo halves of the activations -- 0 (ws may be
 *     NULL) when the split is fused into the GEMM, which the library chooses for large 1x1x1 convolutions.
 * act_amax: NULL, or a DEVICE pointer to an upper bound of max|in|, |in2| (after the optional ReLU) for the fp16 operand scale;
 *     saves the max-reduction pass when the caller knows one (unprojected grids are bounded by max|features|).
 * Needs C % 32 == 0, C2 % 32 == 0, Cout % 16 == 0 (MVF_EUNSUPPORTED otherwise). */
#define MVF_CONV_S1   0
#define MVF_CONV_S2   1
#define MVF_DECONV_S2 2
size_t mvf_conv3d_wsplit_bytes(int kind, int ksize, int Cin, int Cout);
int mvf_conv3d_prepare(const float* W, int kind, int ksize, int V, int C, int C2, int Cout, int chan_interleave,
                       float* wsplit, void* stream);
size_t mvf_conv3d_tc_workspace_bytes(int kind, int ksize, int B, int V, int X, int Y, int Z, int C, int C2, int Cout);
int mvf_conv3d_tc(const float* in, const float* in2, const float* wsplit, const float* bias,
                  const float* bn_scale, const float* bn_shift, const float* pre_scale, const float* pre_shift,
                  int kind, int ksize, int B, int V, int X, int Y, int Z, int C, int C2, int Cout, int flags,
                  float* out, void* ws, size_t ws_bytes, const float* act_amax, void* stream);

/* ---- grid_reas 'ident' on the tensor cores (tcgen05, 3xTF32 split) ------------------------------
 * Same contract as mvf_ident_fuse (model_multi.py:443-455) for C % 32 == 0 and Cout % 16 == 0
 * (MVF_EUNSUPPORTED otherwise: use mvf_ident_fuse); the grid shape is passed as X,Y,Z (N = X*Y*Z).
 *   mvf_ident_prepare(weight [V*C,Cout]) -> wsplit (mvf_ident_wsplit_bytes): K-major hi/lo halves, once per weight.
 * ws: mvf_ident_tc_workspace_bytes(...) bytes of device scratch (hi/lo halves of relu(in); 0 when the split is fused). */
size_t mvf_ident_wsplit_bytes(int V, int C, int Cout);
int mvf_ident_prepare(const float* weight, int V, int C, int Cout, float* wsplit, void* stream);
size_t mvf_ident_tc_workspace_bytes(int B, int V, int X, int Y, int Z, int C, int Cout);
int mvf_ident_fuse_tc(const float* in, const float* wsplit, const float* bias,
                      const float* bn_scale, const float* bn_shift,
                      int B, int V, int X, int Y, int Z, int C, int Cout,
                      float* out, void* ws, size_t ws_bytes, void* stream);

/* ---- ConvLSTM cell step ----------------------------------------------------------------------
 * replaces ConvLSTMCell.call  mrcnn/recurrent.py:442-479 (driven over the view axis by
 * ConvRNN3D, recurrent.py:230-280; wrapper convlstm() model_multi.py:109-123).
 * x [B,X,Y,Z,C] (ReLU applied on load when MVF_FLAG_RELU_IN, model_multi.py:459),
 * h_prev,c_prev [B,X,Y,Z,F] (NULL = zeros, the initial state), W [3,3,3,C+F,4F], bias [4F],
 * gate order j,i,f,o; -> h_out,c_out [B,X,Y,Z,F].  In-place (h_out==h_prev) is NOT allowed. */
int mvf_convlstm_step(const float* x, const float* h_prev, const float* c_prev,
                      const float* W, const float* bias, float forget_bias,
                      int B, int X, int Y, int Z, int C, int F, int flags,
                      float* h_out, float* c_out, void* stream);

/* ---- ConvLSTM cell step on the tensor cores (tcgen05, 3xTF32 split) ------------------------------
 * Same contract as mvf_convlstm_step (recurrent.py:442-479) for C % 32 == 0 and F % 64 == 0
 * (MVF_EUNSUPPORTED otherwise: use mvf_convlstm_step).  Weights are prepared once per weight tensor:
 *   mvf_convlstm_prepare(W [3,3,3,C+F,4F]) -> wsplit (mvf_convlstm_wsplit_bytes): K-major hi/lo halves,
 *   the four gates of each 64-filter group adjacent.
 * ws: mvf_convlstm_tc_workspace_bytes(...) bytes of device scratch (hi/lo halves of x and h_prev). */
size_t mvf_convlstm_wsplit_bytes(int C, int F);
int mvf_convlstm_prepare(const float* W, int C, int F, float* wsplit, void* stream);
size_t mvf_convlstm_tc_workspace_bytes(int B, int X, int Y, int Z, int C, int F);
int mvf_convlstm_step_tc(const float* x, const float* h_prev, const float* c_prev, const float* wsplit,
                         const float* bias, float forget_bias, int B, int X, int Y, int Z, int C, int F,
                         int flags, float* h_out, float* c_out, void* ws, size_t ws_bytes, void* stream);

/* Slab form for the multi-GPU recurrence (x-slabs + 1-voxel halo of x and h): X is the slab's OWN extent;
 * x, h_prev and h_out carry halo_lo + X + halo_hi planes in x (halo_* in {0,1}: 1 where a neighbouring slab exists,
 * its plane filled by the caller's halo exchange; 0 at the grid border, where SAME padding applies); c_prev and
 * c_out carry X planes.  h_out is written at planes [halo_lo, halo_lo + X); its halo planes are left untouched.
 * ws: mvf_convlstm_tc_workspace_bytes(B, halo_lo + X + halo_hi, Y, Z, C, F).
 * act_amax: NULL, or a DEVICE pointer to max(|x| after the optional ReLU, |h_prev|) over the WHOLE grid: the fp16 operand
 *     split scales by a power of two derived from it, and slabs of one grid must use the same scale for the sharded
 *     recurrence to be bit-identical to the unsharded one (dist.lstm_slab all-reduces it).  NULL: computed over this call's
 *     tensors. */
int mvf_convlstm_step_tc_slab(const float* x, const float* h_prev, const float* c_prev, const float* wsplit,
                              const float* bias, float forget_bias, int B, int X, int Y, int Z, int C, int F,
                              int halo_lo, int halo_hi, int flags, float* h_out, float* c_out,
                              void* ws, size_t ws_bytes, const float* act_amax, void* stream);

/* ---- K3: proj_grid ---------------------------------------------------------------------------
 * replaces proj_grid([grid,Rcam,Kmat], config, proj_size)  model_multi.py:231-322 + nearest3 :357-369
 * grid [B,Xs,Y,Z,C] (slab [x_begin, x_begin+x_count) of the full grid; x_count < 0 = MVF_WHOLE_GRID -> whole;
 *      x_count == 0 -> empty slab: `grid` may be NULL and every output sample is 0)
 * Rview [B,3,4]: pose of the camera the rays belong to (reference: Rcam[:,0], :245)
 * Rmain [B,3,4] or NULL (= Rview): pose defining the grid frame (:279-290)
 * grid_pos [B,3] or NULL: required with MVF_FLAG_WORLD_GRID
 * out [B,S,ph,pw,C]; out_vox NULL or int32 [B,S,ph,pw,3]; out_valid NULL or uint8 [B,S,ph,pw]. */
int mvf_project_rays(const float* grid, const float* Rview, const float* Rmain, const float* Kmat,
                     const float* grid_pos, const MvfGrid* g, int B, int C, int img_h,
                     int proj_h, int proj_w, int samples, int flags, double grid_dist,
                     int x_begin, int x_count,
                     float* out, int32_t* out_vox, uint8_t* out_valid, void* stream);

/* ---- K3 + depth_sampling fused ---------------------------------------------------------------
 * replaces proj_grid followed by depth_sampling (non-conv3d branch) model_multi.py:481-487:
 * out[b,i,j,c] = act(bn_scale*(sum_s w[s]*sample[b,s,i,j,c] + bias) + bn_shift).
 * w [S] device; bias/bn_scale/bn_shift host scalars; relu when MVF_FLAG_RELU_OUT.
 * out [B,ph,pw,C]. */
int mvf_project_depth_collapse(const float* grid, const float* Rview, const float* Rmain,
                               const float* Kmat, const float* grid_pos, const MvfGrid* g,
                               int B, int C, int img_h, int proj_h, int proj_w, int samples,
                               int flags, double grid_dist, int x_begin, int x_count,
                               const float* w, float bias, float bn_scale, float bn_shift,
                               float* out, void* stream);

/* ---- depth_sampling on a materialised [B,S,P,P,C] tensor (model_multi.py:481-487) ---------- */
int mvf_depth_collapse(const float* in, int B, int S, long long npix, int C, const float* w,
                       float bias, float bn_scale, float bn_shift, int flags,
                       float* out, void* stream);

/* ---- K4: PyramidROIAlign ---------------------------------------------------------------------
 * replaces PyramidROIAlign(pool_shape)([boxes,image_meta]+maps)  model_multi.py:779-885
 * boxes [B,R,4] normalised (y1,x1,y2,x2); maps[l] = P(2+l) [B,H[l],W[l],C], l=0..3;
 * image_h/image_w = image_meta[0,4:6] (:812).  out [B,R,ph,pw,C] in box order;
 * out_level NULL or int32 [B,R]. */
int mvf_pyramid_roi_align(const float* boxes, const float* const maps[4], const int H[4],
                          const int W[4], int B, int R, int C, int image_h, int image_w,
                          int pool_h, int pool_w, float* out, int32_t* out_level, void* stream);

/* ---- K5: greedy NMS (bitmask, warp ballots) --------------------------------------------------
 * replaces tf.image.non_max_suppression as called at model_multi.py:754,1171.
 * nprob independent problems, problem p uses boxes[p*n .. p*n+n).  class_ids NULL or int32
 * [nprob,n]: suppression only between equal classes and at most max_out kept PER CLASS
 * (= the per-class map_fn of :1166-1187); NULL: class-agnostic.  Class ids must lie in [0, MVF_MAX_CLASSES); a box whose
 * id is outside that range is treated as excluded (never kept, suppresses nothing).
 * keep int32 [nprob,max_total] (indices into the problem's boxes, selection order, -1 padded),
 * keep_count int32 [nprob].  ws: mvf_nms_workspace_bytes(nprob,n) bytes of device scratch. */
size_t mvf_nms_workspace_bytes(int nprob, int n);
int mvf_nms(const float* boxes, const float* scores, const int32_t* class_ids, int nprob, int n,
            float iou_threshold, int max_out, int max_total, int32_t* keep, int32_t* keep_count,
            void* ws, size_t ws_bytes, void* stream);

/* ---- refine_detections_graph / DetectionLayer ------------------------------------------------
 * replaces refine_detections_graph(rois,probs,deltas,window,config) model_multi.py:1119-1214
 * batched as DetectionLayer does (:1245-1248).
 * rois [B,N,4], probs [B,N,K], deltas [B,N,K,4], windows [B,4] (normalised), bbox_std host [4]
 * -> detections [B,max_inst,6] (y1,x1,y2,x2,class,score) zero padded;
 *    out_keep NULL or int32 [B,max_inst] roi indices (-1 padded); out_count NULL or int32 [B]. */
size_t mvf_refine_detections_workspace_bytes(int B, int N);
int mvf_refine_detections(const float* rois, const float* probs, const float* deltas,
                          const float* windows, const float bbox_std[4], int B, int N, int K,
                          float min_confidence, float nms_threshold, int max_inst,
                          float* detections, int32_t* out_keep, int32_t* out_count,
                          void* ws, size_t ws_bytes, void* stream);

/* ---- ProposalLayer ---------------------------------------------------------------------------
 * replaces ProposalLayer(proposal_count,nms_threshold,config)([probs,bbox,anchors])
 * model_multi.py:690-767.  rpn_probs [B,A,2], rpn_bbox [B,A,4], anchors [B,A,4]
 * -> proposals [B,proposal_count,4] zero padded. */
size_t mvf_proposals_workspace_bytes(int B, int A, int pre_nms_limit);
int mvf_proposals(const float* rpn_probs, const float* rpn_bbox, const float* anchors,
                  const float bbox_std[4], int B, int A, int pre_nms_limit, int proposal_count,
                  float nms_threshold, float* proposals, int32_t* out_count,
                  void* ws, size_t ws_bytes, void* stream);

/* ---- fused pipeline through HOST buffers (the end-to-end entry) -----------------------------
 * unproj_feat -> grid_reas(sum|mean|max [+BN+ReLU]) -> proj_grid for B scenes whose inputs and
 * outputs live in (preferably pinned) HOST memory: copies feats/Rcam/Kmat host->device, runs
 * K1 + K3 and copies the ray slices [B,S,ph,pw,C] back, software-pipelined per scene over the three
 * streams of `aux` (ordered after `stream`) so H2D, kernels and D2H overlap, and SYNCHRONISES
 * `stream` before returning.  dev_ws: mvf_pipeline_host_workspace_bytes(...) bytes of device scratch.
 * aux: created once by the caller on the device it will be used on (mvf_host_aux_create: three non-blocking streams and the
 * events that order them); at most one call may be in flight per handle.  When a call fails after work has been queued,
 * the handle's streams are drained before the error is returned, so the host buffers are no longer referenced. */
typedef struct MvfHostAux MvfHostAux;
int mvf_host_aux_create(MvfHostAux** out);
int mvf_host_aux_destroy(MvfHostAux* aux);
size_t mvf_pipeline_host_workspace_bytes(const MvfGrid* g, int B, int V, int fh, int fw, int C,
                                         int proj_h, int proj_w, int samples);
int mvf_unproject_fuse_project_host(const float* h_feats, const float* h_Rcam, const float* h_Kmat,
                                    const MvfGrid* g, int B, int V, int fh, int fw, int C,
                                    int img_h, int img_w, int mode, int flags,
                                    const float* d_bn_scale, const float* d_bn_shift,
                                    int proj_h, int proj_w, int samples,
                                    float* h_out, void* dev_ws, size_t dev_ws_bytes, MvfHostAux* aux, void* stream);

/* ---- the fused pipeline on DEVICE buffers ----------------------------------------------------
 * replaces unproj_feat -> grid_reas(sum|mean|max [+BN+ReLU]) -> proj_grid (model_multi.py:130-228, :401-404, :231-322) for a
 * batch of scenes resident in device memory: feats [B,V,fh,fw,C], Rcam [B,V,3,4], Kmat [B,3,3] in, fused grid [B,X,Y,Z,C] and
 * ray slices [B,S,ph,pw,C] out -- the same kernels and bits as mvf_unproject_fuse(_tc) followed by mvf_project_rays(Rcam[:,0]).
 * One call instead of two plus the gather of the main-view poses; for the configurations mvf_unproject_fuse_tc_supported()
 * accepts, the feature split runs under the tensor-core unprojection (as in mvf_unproject_fuse_tc).  Asynchronous.
 * ws: mvf_unproject_fuse_project_workspace_bytes(...) bytes of device scratch (may be NULL when the tensor-core path does not apply). */
size_t mvf_unproject_fuse_project_workspace_bytes(int B, int V, int fh, int fw, int C);
int mvf_unproject_fuse_project(const float* feats, const float* Rcam, const float* Kmat,
                               const MvfGrid* g, int B, int V, int fh, int fw, int C,
                               int img_h, int img_w, int mode, int flags,
                               const float* bn_scale, const float* bn_shift,
                               int proj_h, int proj_w, int samples,
                               float* grid_out, float* rays_out, void* ws, size_t ws_bytes, void* stream);

/* One pyramid level of the fusion neck (model_multi.py:2382-2404) from HOST buffers: as above, with depth_sampling (non-conv3d
 * branch, :481-487) fused into the projection, so h_out is PG [B,ph,pw,C] and only features go in / PG comes out.
 * d_depth_w [S] device; depth_bias / depth_bn_scale / depth_bn_shift: the folded scalars of the depth conv and its BatchNorm.
 * Same workspace as mvf_unproject_fuse_project_host. */
int mvf_fusion_neck_level_host(const float* h_feats, const float* h_Rcam, const float* h_Kmat,
                               const MvfGrid* g, int B, int V, int fh, int fw, int C,
                               int img_h, int img_w, int mode, int flags,
                               const float* d_bn_scale, const float* d_bn_shift,
                               int proj_h, int proj_w, int samples,
                               const float* d_depth_w, float depth_bias, float depth_bn_scale, float depth_bn_shift,
                               float* h_out, void* dev_ws, size_t dev_ws_bytes, MvfHostAux* aux, void* stream);

/* ---- misc ------------------------------------------------------------------------------------ */
const char* mvf_error_string(int code);
const char* mvf_version(void);
/* kernels launched by this library in this process so far (monotonic; for bench accounting) */
unsigned long long mvf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MVFUSION_H_ */
