// K5  bitmask greedy NMS (warp ballots), refine_detections_graph / DetectionLayer, ProposalLayer.
//
// Replaces tf.image.non_max_suppression as called from mrcnn/model_multi.py:754 and :1171,
// refine_detections_graph (:1119-1214, batched as DetectionLayer :1245-1248) and
// ProposalLayer.call (:705-764).  TensorFlow's NMS / top_k are third-party kernels restated
// from their published algorithms (SURVEY.md spec F): candidates in descending score order
// (ties -> lower index), IoU on min/max-normalised corners with individually rounded fp32
// ops, area <= 0 -> IoU 0, suppress iff IoU > threshold.
//
// Pipeline per problem:  keys (score, index) -> counting sort (<= 8192 keys; the ProposalLayer's top-k over all anchors keeps the
// bitonic network) -> boxes scattered into score order -> IoU bit matrix,
// one __ballot_sync word per (row, 32 columns) -> one-CTA greedy scan over 32-row blocks.
// The per-class NMS of the detection head is ONE pass: a bit is set only between boxes of
// the same class and the scan keeps per-class counters (= map_fn over classes, :1166-1187).
#include "mvf_common.cuh"

namespace mvf {

typedef unsigned long long u64;
constexpr int SORT_CHUNK = 4096;
constexpr u64 KEY_EXCLUDED = ~0ull;

__device__ __forceinline__ unsigned score_key_desc(float s) {
    if (s != s) return 0xFFFFFFFEu;                               // NaN sorts last among candidates
    const unsigned bits = __float_as_uint(s);
    const unsigned asc = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
    return ~asc;                                                  // ascending key == descending score
}
__device__ __forceinline__ u64 make_key(float s, unsigned idx) { return ((u64)score_key_desc(s) << 32) | idx; }
// excluded boxes sort behind every candidate; the index keeps the key unique (the counting sort below ranks by strict comparison)
__device__ __forceinline__ u64 excluded_key(unsigned idx) { return (0xFFFFFFFFull << 32) | idx; }
__device__ __forceinline__ bool is_excluded(u64 key) { return (key >> 32) == 0xFFFFFFFFull; }

__device__ __forceinline__ void cmpx(u64& a, u64& b, bool asc) {
    if ((a > b) == asc) { const u64 t = a; a = b; b = t; }
}
__device__ __forceinline__ int pair_lo(int p, int j) { return ((p & ~(j - 1)) << 1) | (p & (j - 1)); }

// full bitonic sort of each chunk (k = 2..chunk) when k_merge == 0, else the j < chunk tail of stage k_merge
__global__ void __launch_bounds__(512)
bitonic_local_kernel(u64* keys, int n_pad, int chunk, int k_merge) {
    __shared__ u64 s[SORT_CHUNK];
    u64* kp = keys + (size_t)blockIdx.y * n_pad;
    const int base = blockIdx.x * chunk;
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) s[i] = kp[base + i];
    __syncthreads();
    const int k_lo = k_merge ? k_merge : 2, k_hi = k_merge ? k_merge : chunk;
    for (int k = k_lo; k <= k_hi; k <<= 1) {
        for (int j = min(k >> 1, chunk >> 1); j > 0; j >>= 1) {
            for (int p = threadIdx.x; p < (chunk >> 1); p += blockDim.x) {
                const int i = pair_lo(p, j);
                cmpx(s[i], s[i | j], ((base + i) & k) == 0);
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < chunk; i += blockDim.x) kp[base + i] = s[i];
}

__global__ void __launch_bounds__(256)
bitonic_global_kernel(u64* keys, int n_pad, int k, int j) {
    u64* kp = keys + (size_t)blockIdx.y * n_pad;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (n_pad >> 1)) return;
    const int i = pair_lo(p, j);
    u64 a = kp[i], b = kp[i | j];
    const u64 a0 = a;
    cmpx(a, b, (i & k) == 0);
    if (a != a0) { kp[i] = a; kp[i | j] = b; }
}

static int next_pow2(int n) { int p = 64; while (p < n) p <<= 1; return p; }

static int sort_keys(u64* keys, int nprob, int n_pad, cudaStream_t s) {
    const int chunk = n_pad < SORT_CHUNK ? n_pad : SORT_CHUNK;
    dim3 gl(n_pad / chunk, nprob);
    bitonic_local_kernel<<<gl, 512, 0, s>>>(keys, n_pad, chunk, 0);
    count_launch();
    for (int k = chunk << 1; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j >= chunk; j >>= 1) {
            dim3 gg((n_pad / 2 + 255) / 256, nprob);
            bitonic_global_kernel<<<gg, 256, 0, s>>>(keys, n_pad, k, j);
            count_launch();
        }
        bitonic_local_kernel<<<gl, 512, 0, s>>>(keys, n_pad, chunk, k);
        count_launch();
    }
    return check_launch();
}

// ---- box arithmetic (fp32, individually rounded) -------------------------------------------
struct Box { float y1, x1, y2, x2; };

// apply_box_deltas_graph (model_multi.py:648-669); exp = correctly rounded fp32 (double exp rounded once)
__device__ __forceinline__ Box apply_deltas(Box b, float d0, float d1, float d2, float d3) {
    float height = sub_rn(b.y2, b.y1), width = sub_rn(b.x2, b.x1);
    float cy = add_rn(b.y1, mul_rn(0.5f, height)), cx = add_rn(b.x1, mul_rn(0.5f, width));
    cy = add_rn(cy, mul_rn(d0, height));
    cx = add_rn(cx, mul_rn(d1, width));
    height = mul_rn(height, (float)exp((double)d2));
    width = mul_rn(width, (float)exp((double)d3));
    Box r;
    r.y1 = sub_rn(cy, mul_rn(0.5f, height));
    r.x1 = sub_rn(cx, mul_rn(0.5f, width));
    r.y2 = add_rn(r.y1, height);
    r.x2 = add_rn(r.x1, width);
    return r;
}
// clip_boxes_graph (:672-687)
__device__ __forceinline__ Box clip_box(Box b, float wy1, float wx1, float wy2, float wx2) {
    Box r;
    r.y1 = fmaxf(fminf(b.y1, wy2), wy1); r.x1 = fmaxf(fminf(b.x1, wx2), wx1);
    r.y2 = fmaxf(fminf(b.y2, wy2), wy1); r.x2 = fmaxf(fminf(b.x2, wx2), wx1);
    return r;
}
// ---- NMS stages --------------------------------------------------------------------------------
// A class id outside [0, MVF_MAX_CLASSES) would index the scan's shared per-class counters out of bounds: such a box gets the
// EXCLUDED key here, so it sorts behind every candidate (never kept, suppresses nothing) -- the documented contract of mvf_nms.
__global__ void nms_keys_kernel(const float* scores, const int32_t* class_ids, u64* keys, int32_t* rank, int n, int n_pad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int pb = blockIdx.y;
    if (i >= n_pad) return;
    bool ok = i < n;
    if (ok && class_ids) { const int c = class_ids[(size_t)pb * n + i]; ok = c >= 0 && c < MVF_MAX_CLASSES; }
    keys[(size_t)pb * n_pad + i] = ok ? make_key(scores[(size_t)pb * n + i], (unsigned)i) : excluded_key((unsigned)i);
    if (rank && i < n) rank[(size_t)pb * n + i] = 0;
}

// ---- counting sort for <= 8192 keys: rank(i) = #{j : key_j < key_i} (keys are unique), computed as RANK_SPLIT partial counts per
// element over slices of j and combined with atomicAdd; one pass over n^2 / 2^17 comparisons per SM instead of the bitonic network's
// launches (6000 keys: ~5 us instead of 56 us).  The scatter kernel then writes the score-ordered arrays directly.
constexpr int RANK_SPLIT = 8;
__global__ void __launch_bounds__(256)
nms_rank_kernel(const u64* keys, int n, int n_pad, int32_t* rank) {
    __shared__ u64 tile[256];
    const int pb = blockIdx.z;
    const u64* kp = keys + (size_t)pb * n_pad;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const u64 mine = i < n ? kp[i] : 0ull;
    const int per = (n + RANK_SPLIT - 1) / RANK_SPLIT;
    const int j0 = blockIdx.y * per, j1 = min(n, j0 + per);
    int cnt = 0;
    for (int jb = j0; jb < j1; jb += 256) {
        const int j = jb + threadIdx.x;
        __syncthreads();
        tile[threadIdx.x] = j < j1 ? kp[j] : ~0ull;              // pad: never smaller than a real key
        __syncthreads();
#pragma unroll 8
        for (int t = 0; t < 256; ++t) cnt += tile[t] < mine;     // broadcast reads
    }
    if (i < n && cnt) atomicAdd(rank + (size_t)pb * n + i, cnt);
}
// element i -> sorted position rank(i): (original index | -1, box, class)
__global__ void nms_scatter_kernel(const u64* keys, const int32_t* rank, const float4* boxes, const int32_t* class_ids, int n, int n_pad,
                                   int32_t* s_idx, float4* s_box, int32_t* s_cls) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int pb = blockIdx.y;
    if (i >= n) return;
    const u64 key = keys[(size_t)pb * n_pad + i];
    const size_t o = (size_t)pb * n + rank[(size_t)pb * n + i];
    if (is_excluded(key)) { s_idx[o] = -1; s_box[o] = make_float4(0.f, 0.f, 0.f, 0.f); s_cls[o] = -1; return; }
    s_idx[o] = i;
    s_box[o] = boxes[(size_t)pb * n + i];
    s_cls[o] = class_ids ? class_ids[(size_t)pb * n + i] : 0;
}
static int rank_scatter(const u64* keys, int32_t* rank, const float4* boxes, const int32_t* class_ids, int nprob, int n, int n_pad,
                        int32_t* s_idx, float4* s_box, int32_t* s_cls, cudaStream_t s) {
    nms_rank_kernel<<<dim3((n + 255) / 256, RANK_SPLIT, nprob), 256, 0, s>>>(keys, n, n_pad, rank);
    count_launch();
    nms_scatter_kernel<<<dim3((n + 255) / 256, nprob), 256, 0, s>>>(keys, rank, boxes, class_ids, n, n_pad, s_idx, s_box, s_cls);
    count_launch();
    return check_launch();
}

constexpr int MASK_ROWS = 256;   // rows per CTA (8 warps x 32 rows), staged once in shared memory

// min/max-normalised corners (tf.image.non_max_suppression accepts either corner order) and the area, as iou_above computes them
struct NBox { float ymin, xmin, ymax, xmax, area; int cls; };
__device__ __forceinline__ NBox norm_box(const float4 a, int cls) {
    NBox b;
    b.ymin = fminf(a.x, a.z); b.ymax = fmaxf(a.x, a.z); b.xmin = fminf(a.y, a.w); b.xmax = fmaxf(a.y, a.w);
    b.area = mul_rn(sub_rn(b.ymax, b.ymin), sub_rn(b.xmax, b.xmin));
    b.cls = cls;
    return b;
}
// iou_above on pre-normalised boxes (same operations in the same order: bit-identical decisions)
__device__ __forceinline__ bool iou_above_n(const NBox& i, const NBox& j, float thr) {
    if (i.area <= 0.f || j.area <= 0.f) return false;
    const float iy0 = fmaxf(i.ymin, j.ymin), ix0 = fmaxf(i.xmin, j.xmin);
    const float iy1 = fminf(i.ymax, j.ymax), ix1 = fminf(i.xmax, j.xmax);
    const float inter = mul_rn(fmaxf(sub_rn(iy1, iy0), 0.f), fmaxf(sub_rn(ix1, ix0), 0.f));
    if (!(inter > 0.f)) return 0.f > thr;
    const float uni = sub_rn(add_rn(i.area, j.area), inter);
    const float q = __fdividef(inter, uni);
    if (fabsf(q - thr) > 1e-5f) return q > thr;
    return div_rn(inter, uni) > thr;
}

// bit (r, c) = c > r  &&  same class  &&  IoU(r, c) > thr   -- one ballot word per (row, 32 columns).  A CTA owns 256 rows x 32
// columns: the row boxes are normalised once into shared memory (uniform reads in the loop), the column box lives in registers.
__global__ void __launch_bounds__(256)
nms_mask_kernel(const float4* s_box, const int32_t* s_cls, int n, int nwords, int pitch, float thr, unsigned* mask) {
    __shared__ NBox rows[MASK_ROWS];
    const int w = blockIdx.x, pb = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row0 = blockIdx.y * MASK_ROWS;
    if (w * 32 + 31 <= row0) return;                              // word entirely at or below the diagonal
    const float4* bx = s_box + (size_t)pb * n;
    const int32_t* cl = s_cls + (size_t)pb * n;
    {
        const int r = row0 + threadIdx.x;
        rows[threadIdx.x] = r < n ? norm_box(bx[r], cl[r]) : norm_box(make_float4(0.f, 0.f, 0.f, 0.f), -1);
    }
    const int col = w * 32 + lane;
    const bool col_ok = col < n;
    const NBox cb = col_ok ? norm_box(bx[col], cl[col]) : norm_box(make_float4(0.f, 0.f, 0.f, 0.f), -2);
    unsigned* mp = mask + (size_t)pb * n * pitch;                 // rows padded to 16 B: the scan reads them with 128-bit loads
    __syncthreads();
    for (int rr = warp * 32; rr < warp * 32 + 32; ++rr) {
        const int r = row0 + rr;
        if (r >= n) break;
        if (w * 32 + 31 <= r) break;                              // the remaining rows of this warp lie below the diagonal too
        const NBox rb = rows[rr];                                 // uniform address: broadcast
        const bool bit = col_ok && col > r && cb.cls == rb.cls && rb.cls >= 0 && iou_above_n(rb, cb, thr);
        const unsigned word = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) mp[(size_t)r * pitch + w] = word;
    }
}

// greedy selection over the bit matrix; one CTA (256 threads, thread t owns removed-word t) per problem.
// The sweep is a latency chain (one dependent round per 32 candidates), so nothing on it may wait for memory:
//   * warp 0 resolves a block with lane b holding candidate b (index, class) and two mask words of its row, both loaded one block
//     ahead: the DIAGONAL word (suppression inside the block) and the NEXT word (this block's effect on the next block, passed on in
//     a register as `carry`), so the chain of the next block never waits for the removal words below;
//   * with one class and no per-class cap every lane fetches the 32 diagonal words by (independent) shuffles and runs the same
//     32-step chain on registers; with classes the chain walks the surviving candidates through shuffles and shared counters
//     (tf NMS stops a class at max_output_size);
//   * thread t > wi + 1 loads word t of the kept rows straight from global memory (no staging) into a small register buffer and ORs
//     it into its removal word one round LATER: the L2 round trip overlaps the next block's chain (more than 8 kept rows in a block --
//     the case that fills the output within a few blocks -- are ORed at once).
__global__ void __launch_bounds__(256)
nms_scan_kernel(const unsigned* mask, const int32_t* s_idx, const int32_t* s_cls, int n, int nwords, int pitch,
                int max_per_class, int max_total, int by_position, int single_class, int32_t* keep, int32_t* keep_count) {
    __shared__ unsigned sh_cur, sh_kept;
    __shared__ int sh_total, sh_done;
    __shared__ int sh_cnt[MVF_MAX_CLASSES];
    const int pb = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const unsigned* mp = mask + (size_t)pb * n * pitch;
    const int32_t* ip = s_idx + (size_t)pb * n;
    const int32_t* cp = s_cls + (size_t)pb * n;
    int32_t* kp = keep + (size_t)pb * max_total;
    for (int i = tid; i < MVF_MAX_CLASSES; i += blockDim.x) sh_cnt[i] = 0;
    if (tid == 0) { sh_total = 0; sh_done = 0; }
    // warp 0: candidate `lane` of the next block -- index, class, diagonal word and next word of its mask row
    int nidx = -1, ncls = -1;
    unsigned ndiag = 0, nnext = 0;
    auto prefetch = [&](int blk) {
        const int row = blk * 32 + lane;
        const bool ok = blk < nwords && row < n;
        nidx = ok ? ip[row] : -1;
        ncls = ok ? cp[row] : -1;
        ndiag = ok ? __ldg(mp + (size_t)row * pitch + blk) : 0u;
        nnext = (ok && blk + 1 < nwords) ? __ldg(mp + (size_t)row * pitch + blk + 1) : 0u;
    };
    if (tid < 32) prefetch(0);
    unsigned remv = 0, carry = 0;
    unsigned pend[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};          // words loaded in the previous round
    int total = 0;                                                // warp 0: running number of kept boxes (uniform)
    for (int wi = 0; wi < nwords; ++wi) {
        if (tid == wi) sh_cur = remv;                             // contributions of the blocks before wi - 1
        __syncthreads();
        if (tid < 32) {
            const int idx = nidx, cls = ncls;
            const unsigned diag = ndiag, next = nnext;
            prefetch(wi + 1);                                     // issued now, used next iteration
            unsigned cur = sh_cur | carry, kept = 0;
            const unsigned present = __ballot_sync(0xffffffffu, idx >= 0);     // candidates of this block (a prefix of the lanes)
            int done = present != 0xffffffffu;                                 // a short block is the last one
            if (single_class) {
                unsigned dg[32];
#pragma unroll
                for (int b = 0; b < 32; ++b) dg[b] = __shfl_sync(0xffffffffu, diag, b);
                int room = max_total - total;
                cur |= ~present;                                               // absent candidates count as suppressed
                if (room >= 32) {                                              // two dependent logic ops per candidate
#pragma unroll
                    for (int b = 0; b < 32; ++b)
                        if (!((cur >> b) & 1u)) { kept |= 1u << b; cur |= dg[b]; }
                } else {
#pragma unroll
                    for (int b = 0; b < 32; ++b)
                        if (!((cur >> b) & 1u) && room > 0) { kept |= 1u << b; cur |= dg[b]; --room; }
                }
                if ((kept >> lane) & 1u) kp[total + __popc(kept & ((1u << lane) - 1u))] = by_position ? (wi * 32 + lane) : idx;
                total += __popc(kept);
            } else {
                unsigned alive = present & ~cur;                               // visit only candidates that are not suppressed yet
#pragma unroll 1
                while (alive) {
                    if (total >= max_total) { done = 1; break; }               // output full
                    const int b = __ffs(alive) - 1;
                    alive &= ~(1u << b);
                    const int c = __shfl_sync(0xffffffffu, cls, b);
                    const int cnt = sh_cnt[c];                                // broadcast read
                    if (cnt >= max_per_class) continue;                       // tf NMS stops a class at max_output_size
                    __syncwarp();
                    if (lane == 0) sh_cnt[c] = cnt + 1;
                    __syncwarp();
                    const int bi = __shfl_sync(0xffffffffu, idx, b);
                    if (lane == 0) kp[total] = by_position ? (wi * 32 + b) : bi;
                    ++total;
                    kept |= 1u << b;
                    cur |= __shfl_sync(0xffffffffu, diag, b);
                    alive &= ~cur;
                }
            }
            carry = __reduce_or_sync(0xffffffffu, ((kept >> lane) & 1u) ? next : 0u);   // this block's suppression of block wi + 1
            if (total >= max_total) done = 1;
            if (lane == 0) { sh_total = total; sh_kept = kept; sh_done = done; }
        }
        __syncthreads();
        if (sh_done) break;
        unsigned kept = sh_kept;
#pragma unroll
        for (int k = 0; k < 8; ++k) { remv |= pend[k]; pend[k] = 0u; }           // issued one round ago
        if (tid > wi + 1 && tid < nwords) {                       // word wi + 1 travels in `carry`
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (kept) {
                    const int b = __ffs(kept) - 1;
                    kept &= kept - 1;
                    pend[k] = __ldg(mp + (size_t)(wi * 32 + b) * pitch + tid);
                }
            }
            if (kept) {                                           // many kept rows: independent predicated loads, one wait
#pragma unroll 8
                for (int b = 0; b < 32; ++b)
                    if ((kept >> b) & 1u) remv |= __ldg(mp + (size_t)(wi * 32 + b) * pitch + tid);
            }
        }
    }
    __syncthreads();
    const int ntotal = sh_total;
    for (int i = ntotal + tid; i < max_total; i += blockDim.x) kp[i] = -1;
    if (tid == 0 && keep_count) keep_count[pb] = ntotal;
}

struct NmsWs { u64* keys; int32_t* rank; int32_t* s_idx; float4* s_box; int32_t* s_cls; unsigned* mask; size_t bytes; };

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

static NmsWs carve_nms(void* ws, int nprob, int n, int n_pad) {
    NmsWs w;
    char* p = (char*)ws;
    size_t off = 0;
    const int nwords = (n + 31) / 32;
    w.keys = (u64*)(p + off); off += align_up((size_t)nprob * n_pad * sizeof(u64));
    w.rank = (int32_t*)(p + off); off += align_up((size_t)nprob * n * sizeof(int32_t));
    w.s_idx = (int32_t*)(p + off); off += align_up((size_t)nprob * n * sizeof(int32_t));
    w.s_box = (float4*)(p + off); off += align_up((size_t)nprob * n * sizeof(float4));
    w.s_cls = (int32_t*)(p + off); off += align_up((size_t)nprob * n * sizeof(int32_t));
    w.mask = (unsigned*)(p + off); off += align_up((size_t)nprob * n * ((nwords + 3) & ~3) * sizeof(unsigned));
    w.bytes = off;
    return w;
}

// mask + scan on already gathered (score-ordered) boxes
static int run_mask_scan(const NmsWs& w, int nprob, int n, float thr, int max_per_class, int max_total, int by_position,
                         int single_class, int32_t* keep, int32_t* keep_count, cudaStream_t s) {
    const int nwords = (n + 31) / 32;
    dim3 gm(nwords, (n + MASK_ROWS - 1) / MASK_ROWS, nprob);
    const int pitch = (nwords + 3) & ~3;
    nms_mask_kernel<<<gm, 256, 0, s>>>(w.s_box, w.s_cls, n, nwords, pitch, thr, w.mask);
    count_launch();
    nms_scan_kernel<<<nprob, 256, 0, s>>>(w.mask, w.s_idx, w.s_cls, n, nwords, pitch, max_per_class, max_total, by_position,
                                          single_class, keep, keep_count);
    count_launch();
    return check_launch();
}

// ---- refine_detections ---------------------------------------------------------------------------
__global__ void refine_prepare_kernel(const float* rois, const float* probs, const float* deltas, const float* windows,
                                      float s0, float s1, float s2, float s3, int N, int K, float min_conf, int n_pad,
                                      float4* refined, int32_t* cls_out, float* score_out, u64* keys, int32_t* rank) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= n_pad) return;
    if (i >= N) { keys[(size_t)b * n_pad + i] = excluded_key((unsigned)i); return; }
    const size_t o = (size_t)b * N + i;
    const float* pr = probs + o * K;
    int cls = 0; float best = pr[0];
    for (int k = 1; k < K; ++k) { const float v = pr[k]; if (v > best) { best = v; cls = k; } }      // argmax: first maximum (:1135)
    const float* d = deltas + (o * K + cls) * 4;
    Box r; r.y1 = rois[o * 4 + 0]; r.x1 = rois[o * 4 + 1]; r.y2 = rois[o * 4 + 2]; r.x2 = rois[o * 4 + 3];
    r = apply_deltas(r, mul_rn(d[0], s0), mul_rn(d[1], s1), mul_rn(d[2], s2), mul_rn(d[3], s3));    // :1143-1144
    const float* w = windows + (size_t)b * 4;
    r = clip_box(r, w[0], w[1], w[2], w[3]);                                                      // :1146
    refined[o] = make_float4(r.y1, r.x1, r.y2, r.x2);
    cls_out[o] = cls;
    score_out[o] = best;
    const bool cand = (cls > 0) && (min_conf == 0.f || best >= min_conf);                         // :1151-1157
    keys[(size_t)b * n_pad + i] = cand ? make_key(best, (unsigned)i) : excluded_key((unsigned)i);
    rank[o] = 0;
}

__global__ void write_detections_kernel(const int32_t* keep, const float4* refined, const int32_t* cls, const float* score,
                                        int N, int max_inst, float* det, int32_t* out_keep) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    const int b = blockIdx.y;
    if (i >= max_inst) return;
    const int k = keep[(size_t)b * max_inst + i];
    float* o = det + ((size_t)b * max_inst + i) * 6;
    if (out_keep) out_keep[(size_t)b * max_inst + i] = k;
    if (k < 0) { for (int e = 0; e < 6; ++e) o[e] = 0.f; return; }
    const float4 r = refined[(size_t)b * N + k];
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
    o[4] = (float)cls[(size_t)b * N + k];                                                         // :1207
    o[5] = score[(size_t)b * N + k];
}

// ---- ProposalLayer -------------------------------------------------------------------------------
__global__ void proposal_keys_kernel(const float* rpn_probs, u64* keys, int A, int n_pad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= n_pad) return;
    keys[(size_t)b * n_pad + i] = (i < A) ? make_key(rpn_probs[((size_t)b * A + i) * 2 + 1], (unsigned)i) : KEY_EXCLUDED;   // fg score :707
}

// top-`limit` anchors in score order -> refined, clipped boxes (:723-745)
__global__ void proposal_boxes_kernel(const u64* keys, const float* rpn_bbox, const float* anchors, float s0, float s1,
                                      float s2, float s3, int A, int n_pad, int limit, int32_t* s_idx, float4* s_box,
                                      int32_t* s_cls) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (r >= limit) return;
    const unsigned idx = (unsigned)(keys[(size_t)b * n_pad + r] & 0xFFFFFFFFu);
    const float* d = rpn_bbox + ((size_t)b * A + idx) * 4;
    const float* a = anchors + ((size_t)b * A + idx) * 4;
    Box bx; bx.y1 = a[0]; bx.x1 = a[1]; bx.y2 = a[2]; bx.x2 = a[3];
    bx = apply_deltas(bx, mul_rn(d[0], s0), mul_rn(d[1], s1), mul_rn(d[2], s2), mul_rn(d[3], s3));
    bx = clip_box(bx, 0.f, 0.f, 1.f, 1.f);
    const size_t o = (size_t)b * limit + r;
    s_idx[o] = (int)idx; s_box[o] = make_float4(bx.y1, bx.x1, bx.y2, bx.x2); s_cls[o] = 0;
}

__global__ void write_proposals_kernel(const int32_t* keep, const float4* s_box, int limit, int count, float4* out) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    const int b = blockIdx.y;
    if (i >= count) return;
    const int k = keep[(size_t)b * count + i];
    out[(size_t)b * count + i] = (k < 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : s_box[(size_t)b * limit + k];    // tf.pad :759-760
}

}  // namespace mvf

using namespace mvf;

extern "C" size_t mvf_nms_workspace_bytes(int nprob, int n) {
    if (nprob <= 0 || n <= 0) return 0;
    return carve_nms(nullptr, nprob, n, next_pow2(n)).bytes;
}

extern "C" int mvf_nms(const float* boxes, const float* scores, const int32_t* class_ids, int nprob, int n,
                       float iou_threshold, int max_out, int max_total, int32_t* keep, int32_t* keep_count,
                       void* ws, size_t ws_bytes, void* stream) {
    if (!boxes || !scores || !keep || !ws) return MVF_ENULL;
    if (nprob <= 0 || n <= 0 || max_out <= 0 || max_total <= 0) return MVF_EINVAL;
    if (n > MVF_MAX_NMS_BOXES || nprob > 65535) return MVF_EUNSUPPORTED;
    if (!aligned16(boxes) || !aligned16(ws)) return MVF_EALIGN;
    const int n_pad = next_pow2(n);
    NmsWs w = carve_nms(ws, nprob, n, n_pad);
    if (ws_bytes < w.bytes) return MVF_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 gk((n_pad + 255) / 256, nprob);
    nms_keys_kernel<<<gk, 256, 0, s>>>(scores, class_ids, w.keys, w.rank, n, n_pad);
    count_launch();
    int rc = rank_scatter(w.keys, w.rank, (const float4*)boxes, class_ids, nprob, n, n_pad, w.s_idx, w.s_box, w.s_cls, s);
    if (rc != MVF_OK) return rc;
    return run_mask_scan(w, nprob, n, iou_threshold, max_out, max_total, 0, class_ids == nullptr && max_out >= max_total, keep, keep_count, s);
}

struct RefineWs { float4* refined; int32_t* cls; float* score; int32_t* keep; NmsWs nms; size_t bytes; };

static RefineWs carve_refine(void* ws, int B, int N, int max_inst) {
    RefineWs r;
    char* p = (char*)ws;
    size_t off = 0;
    r.refined = (float4*)(p + off); off += align_up((size_t)B * N * sizeof(float4));
    r.cls = (int32_t*)(p + off); off += align_up((size_t)B * N * sizeof(int32_t));
    r.score = (float*)(p + off); off += align_up((size_t)B * N * sizeof(float));
    r.keep = (int32_t*)(p + off); off += align_up((size_t)B * max_inst * sizeof(int32_t));
    r.nms = carve_nms(ws ? p + off : nullptr, B, N, next_pow2(N));
    r.bytes = off + r.nms.bytes;
    return r;
}

extern "C" size_t mvf_refine_detections_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    return carve_refine(nullptr, B, N, 4096).bytes;
}

extern "C" int mvf_refine_detections(const float* rois, const float* probs, const float* deltas,
                                     const float* windows, const float bbox_std[4], int B, int N, int K,
                                     float min_confidence, float nms_threshold, int max_inst,
                                     float* detections, int32_t* out_keep, int32_t* out_count,
                                     void* ws, size_t ws_bytes, void* stream) {
    if (!rois || !probs || !deltas || !windows || !bbox_std || !detections || !ws) return MVF_ENULL;
    if (B <= 0 || N <= 0 || K <= 0 || max_inst <= 0) return MVF_EINVAL;
    if (N > MVF_MAX_NMS_BOXES || K > MVF_MAX_CLASSES || max_inst > 4096 || B > 65535) return MVF_EUNSUPPORTED;
    if (!aligned16(ws)) return MVF_EALIGN;
    RefineWs w = carve_refine(ws, B, N, max_inst);
    if (ws_bytes < w.bytes) return MVF_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int n_pad = next_pow2(N);
    dim3 gp((n_pad + 255) / 256, B);
    refine_prepare_kernel<<<gp, 256, 0, s>>>(rois, probs, deltas, windows, bbox_std[0], bbox_std[1], bbox_std[2], bbox_std[3],
                                             N, K, min_confidence, n_pad, w.refined, w.cls, w.score, w.nms.keys, w.nms.rank);
    count_launch();
    int rc = rank_scatter(w.nms.keys, w.nms.rank, w.refined, w.cls, B, N, n_pad, w.nms.s_idx, w.nms.s_box, w.nms.s_cls, s);
    if (rc != MVF_OK) return rc;
    // per-class NMS (<= max_inst per class, :1171-1174) and top-max_inst by score (:1197-1201) in one scan
    rc = run_mask_scan(w.nms, B, N, nms_threshold, max_inst, max_inst, 0, 0, w.keep, out_count, s);
    if (rc != MVF_OK) return rc;
    dim3 gw((max_inst + 127) / 128, B);
    write_detections_kernel<<<gw, 128, 0, s>>>(w.keep, w.refined, w.cls, w.score, N, max_inst, detections, out_keep);
    count_launch();
    return check_launch();
}

struct PropWs { u64* keys; int32_t* keep; NmsWs nms; size_t bytes; };

static PropWs carve_prop(void* ws, int B, int A, int limit, int count) {
    PropWs r;
    char* p = (char*)ws;
    size_t off = 0;
    const int a_pad = next_pow2(A);
    r.keys = (u64*)(p + off); off += align_up((size_t)B * a_pad * sizeof(u64));
    r.keep = (int32_t*)(p + off); off += align_up((size_t)B * count * sizeof(int32_t));
    r.nms = carve_nms(ws ? p + off : nullptr, B, limit, 64);      // keys of the inner problem unused
    r.bytes = off + r.nms.bytes;
    return r;
}

extern "C" size_t mvf_proposals_workspace_bytes(int B, int A, int pre_nms_limit) {
    if (B <= 0 || A <= 0 || pre_nms_limit <= 0) return 0;
    const int limit = pre_nms_limit < A ? pre_nms_limit : A;
    return carve_prop(nullptr, B, A, limit, MVF_MAX_NMS_BOXES).bytes;
}

extern "C" int mvf_proposals(const float* rpn_probs, const float* rpn_bbox, const float* anchors,
                             const float bbox_std[4], int B, int A, int pre_nms_limit, int proposal_count,
                             float nms_threshold, float* proposals, int32_t* out_count,
                             void* ws, size_t ws_bytes, void* stream) {
    if (!rpn_probs || !rpn_bbox || !anchors || !bbox_std || !proposals || !ws) return MVF_ENULL;
    if (B <= 0 || A <= 0 || pre_nms_limit <= 0 || proposal_count <= 0) return MVF_EINVAL;
    const int limit = pre_nms_limit < A ? pre_nms_limit : A;             // :721
    if (limit > MVF_MAX_NMS_BOXES || proposal_count > MVF_MAX_NMS_BOXES || A > (1 << 24) || B > 65535) return MVF_EUNSUPPORTED;
    if (!aligned16(ws) || !aligned16(proposals)) return MVF_EALIGN;
    PropWs w = carve_prop(ws, B, A, limit, proposal_count);
    if (ws_bytes < w.bytes) return MVF_EWORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const int a_pad = next_pow2(A);
    dim3 gk((a_pad + 255) / 256, B);
    proposal_keys_kernel<<<gk, 256, 0, s>>>(rpn_probs, w.keys, A, a_pad);
    count_launch();
    int rc = sort_keys(w.keys, B, a_pad, s);                              // tf.nn.top_k(sorted=True) :723
    if (rc != MVF_OK) return rc;
    dim3 gb((limit + 255) / 256, B);
    proposal_boxes_kernel<<<gb, 256, 0, s>>>(w.keys, rpn_bbox, anchors, bbox_std[0], bbox_std[1], bbox_std[2], bbox_std[3],
                                             A, a_pad, limit, w.nms.s_idx, w.nms.s_box, w.nms.s_cls);
    count_launch();
    rc = run_mask_scan(w.nms, B, limit, nms_threshold, proposal_count, proposal_count, 1, 1, w.keep, out_count, s);
    if (rc != MVF_OK) return rc;
    dim3 gw((proposal_count + 127) / 128, B);
    write_proposals_kernel<<<gw, 128, 0, s>>>(w.keep, w.nms.s_box, limit, proposal_count, (float4*)proposals);
    count_launch();
    return check_launch();
}
