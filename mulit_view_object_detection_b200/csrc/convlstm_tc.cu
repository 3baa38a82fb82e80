// K2  ConvLSTM gates on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces ConvLSTMCell.call  mrcnn/recurrent.py:442-479:
//     y = conv3d_SAME([x ; h_prev], W) + b      W [3,3,3,C+F,4F]   (recurrent.py:423-431, :453-459)
//     j,i,f,o = split(y)  ;  c = c_prev*sig(f + forget_bias) + sig(i)*tanh(j)  ;  h = tanh(c)*sig(o)   (:460-477)
//
// Implicit GEMM:  M = voxels, N = 4F, K = 27*(C+F).  One CTA owns a 128-voxel box (BX x BY x BZ,
// z fastest) and 64 filters x 4 gates = 256 accumulator columns in TMEM, so the gate
// non-linearities run in the epilogue on values that never leave the SM.
//   * A operand: for tap (dx,dy,dz) and a 32-channel chunk the 128 x 32 fp32 tile is ONE 5-D TMA box
//     load of the channel-last tensor [B,X,Y,Z,C] at coordinates shifted by the tap; coordinates
//     outside the grid are zero-filled by the TMA unit, which IS the conv's SAME padding -- no
//     im2col, no halo logic, no per-element predicates.  The box lands in the 128-byte-swizzled
//     K-major layout tcgen05.mma reads directly.
//   * B operand: the weights, transposed once to K-major [4F, 27*(C+F)] with the four gates of a
//     64-filter group adjacent (mvf_convlstm_prepare), 256 x 32 tile per chunk, 2-D TMA.
//   * fp32 parity: the reference convolves in fp32; a plain TF32 MMA (10-bit mantissa) is 1e-3
//     off.  Operands are split a = a_hi + a_lo (a_hi = the top 19 bits, exact in TF32) and three
//     MMAs  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi  accumulate in fp32 TMEM: ~2^-21 per product.
//     a_hi / a_lo of the activations are produced by a tiny elementwise pass (which also applies the
//     ReLU of model_multi.py:459), b_hi / b_lo once per weight tensor.
//   * accumulation: TMEM adds truncate (measured: the error of one long chain grows linearly with K and is
//     biased toward zero, 1.2e-4 at K = 13 824), so the chain is cut every `promote` K-chunks: the epilogue
//     warps add the partial accumulator into a MASTER accumulator (second half of TMEM) with
//     round-to-nearest fp32 adds (tcgen05.ld -> FADD -> tcgen05.st) and the MMAs restart from zero.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM
//     allocator, warps 4-7 = epilogue (tcgen05.ld -> gates -> c/h stores).  2-stage smem ring
//     (96 KB per stage: A_hi, A_lo, B_hi, B_lo), full/empty mbarriers, tcgen05.commit.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "mvf_common.cuh"
#include "tc_ptx.cuh"

namespace mvf {

constexpr int TC_M = 128, TC_N = 256, TC_K = 32, TC_STAGES = 2, TC_THREADS = 256;   // TC_K: tf32 elements per 128-byte K chunk
constexpr int TC_K16 = 64;                                   // fp16 elements per 128-byte K chunk
constexpr int TC_FPT = 64;                                   // filters per CTA tile (x 4 gates = TC_N)
constexpr uint32_t TC_A_BYTES = TC_M * TC_K * 4;             // 16 KB
constexpr uint32_t TC_B_BYTES = TC_N * TC_K * 4;             // 32 KB
constexpr uint32_t TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;
constexpr uint32_t TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int TC_CONV_WARPS = 2;                            // warps 2-3 convert raw tiles into hi/lo halves

struct TcArgs {
    const float* bias; const float* c_prev; float* h_out; float* c_out;
    int B, X, Y, Z, C, F;
    int BX, BY, BZ, tiles_x, tiles_y, tiles_z;
    int has_h;
    int halo_lo, halo_hi;         // slab mode: x and h_prev / h_out carry halo_lo + X + halo_hi planes in x, c_prev / c_out carry X
    // plain mode (conv3d family): x = V tensors [B,V,(sub-lattices),X,Y,Z,C] concatenated on channels, then the F channels of `h`
    int V, Cout;                  // view-sources of x, output channels
    int kind, ksize;              // MVF_CONV_S1 / _S2 / MVF_DECONV_S2, kernel size per axis (1 or 3)
    int relu_out;
    // fused operand split: the TMA loads the RAW fp32 activation tile and warps 2-3 turn it into (hi, lo) in shared memory
    int fused_split, relu_x, relu_h;
    // fp16 operand split: activations and weights were scaled by powers of two; the epilogue multiplies by inv_a * inv_w
    const float* inv_scale_a; const float* inv_scale_w;
    const float* pre_scale; const float* pre_shift;
    const float* bn_scale; const float* bn_shift; float* out;
    int promote;                  // K-chunks per promotion of the partial accumulator into the master (0 = never)
    float forget_bias;
};

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=256 (cute::UMMA::InstrDescriptor)
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
// the same with A = B = f16 (format 0), for kind::f16
constexpr uint32_t TC_IDESC_F16 = (1u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC_F16), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// IDENT = false: ConvLSTM step (27 taps, gates in the epilogue).
// IDENT = true : the plain conv3d family with a bias -> BN -> ReLU epilogue, on the same pipeline:
//   MVF_CONV_S1   Conv3D k=1|3, stride 1, SAME.  k=1, V views = grid_reas 'ident' (model_multi.py:443-455).
//   MVF_CONV_S2   Conv3D k=3, stride 2, SAME on even dims (TF pads only at the high end: in = 2*o + k).  The split pass
//                 stores the input as its 8 parity sub-lattices, so tap k reads a PLAIN box of sub-lattice (k & 1) shifted by
//                 (k == 2); the TMA zero fill at the high end is TF's pad_after = 1   (model_multi.py:415-428).
//   MVF_DECONV_S2 Conv3DTranspose k=3, stride 2, SAME: out[2o + k] += in[o] * W[k].  blockIdx.z = output parity class;
//                 an even output coordinate gets taps (o, k=0) and (o-1, k=2), an odd one (o, k=1)   (:430-441).
struct TapRef { int sub, dx, dy, dz, kidx; };
__device__ __forceinline__ TapRef decode_tap(int kind, int ksize, int cls, int t) {
    TapRef r = {0, 0, 0, 0, 0};
    if (kind == MVF_CONV_S1) {
        if (ksize == 3) { r.dx = t / 9 - 1; r.dy = (t / 3) % 3 - 1; r.dz = t % 3 - 1; r.kidx = t; }
    } else if (kind == MVF_CONV_S2) {
        const int kx = t / 9, ky = (t / 3) % 3, kz = t % 3;
        r.sub = ((kx & 1) * 2 + (ky & 1)) * 2 + (kz & 1);
        r.dx = kx >> 1; r.dy = ky >> 1; r.dz = kz >> 1; r.kidx = t;
    } else {                                                 // MVF_DECONV_S2: cls = (px,py,pz); bit j of t picks the second tap of the j-th even axis
        const int px = (cls >> 2) & 1, py = (cls >> 1) & 1, pz = cls & 1;
        int bit = 0, kx, ky, kz;
        if (px) kx = 1; else { const int s2 = (t >> bit) & 1; ++bit; kx = s2 ? 2 : 0; r.dx = -s2; }
        if (py) ky = 1; else { const int s2 = (t >> bit) & 1; ++bit; ky = s2 ? 2 : 0; r.dy = -s2; }
        if (pz) kz = 1; else { const int s2 = (t >> bit) & 1; ++bit; kz = s2 ? 2 : 0; r.dz = -s2; }
        r.kidx = (kx * 3 + ky) * 3 + kz;
    }
    return r;
}
__device__ __forceinline__ int tap_count(int kind, int ksize, int cls) {
    if (kind == MVF_DECONV_S2) return 1 << (3 - (((cls >> 2) & 1) + ((cls >> 1) & 1) + (cls & 1)));
    return ksize == 3 ? 27 : 1;
}

template <bool IDENT, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
convlstm_tc_kernel(const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl,
                   const __grid_constant__ CUtensorMap tm_hh, const __grid_constant__ CUtensorMap tm_hl,
                   const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_wl,
                   const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;              // swizzle-128B tiles need 1024 B alignment
    const uint32_t bar_base = smem_base + TC_STAGES * TC_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
    const uint32_t part_full = bar_base + 8u * (2 * TC_STAGES);                    // MMA -> epilogue: a partial accumulator is complete
    const uint32_t part_empty = bar_base + 8u * (2 * TC_STAGES + 1);               // epilogue -> MMA: it has been added into the master
    const uint32_t tmem_slot = bar_base + 8u * (2 * TC_STAGES + 2);                // the allocator writes the TMEM base address here
    auto conv_bar = [&](int s) { return bar_base + 8u * (2 * TC_STAGES + 3 + s); };  // converter warps -> MMA: hi/lo halves are in place
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates: blockIdx.x -> (b, tx, ty, tz) box of BX x BY x BZ voxels, blockIdx.y -> 64-filter group
    int t = blockIdx.x;
    const int tz = t % a.tiles_z; t /= a.tiles_z;
    const int ty = t % a.tiles_y; t /= a.tiles_y;
    const int tx = t % a.tiles_x; const int b = t / a.tiles_x;
    const int x0 = tx * a.BX, y0 = ty * a.BY, z0 = tz * a.BZ;
    const int ntile = blockIdx.y;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); mbar_init(conv_bar(s), TC_CONV_WARPS); }
        mbar_init(part_full, 1); mbar_init(part_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tmem_slot), "n"(2 * TC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_d;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_d) : "r"(tmem_slot) : "memory");

    constexpr int KC = F16 ? TC_K16 : TC_K;                       // channels per 128-byte K chunk
    const int cx = a.C / KC, ch = a.has_h ? a.F / KC : 0;         // K chunks of x and of h_prev per tap
    const int cls = IDENT ? (int)blockIdx.z : 0;                  // output parity class (deconv)
    const int per_tap = IDENT ? a.V * cx + ch : cx + ch;
    const int nchunks = IDENT ? tap_count(a.kind, a.ksize, cls) * per_tap : 27 * per_tap;
    const int CF = IDENT ? a.V * a.C + a.F : a.C + a.F;
    const int nsub = (IDENT && a.kind == MVF_CONV_S2) ? 8 : 1;
    const int gsz = a.promote > 0 ? a.promote : nchunks;         // K-chunks per partial accumulation chain
    const int ngroups = (nchunks + gsz - 1) / gsz;

    // K-chunk `it` -> source (x view v / h), channel offset, tap shift, weight row, batch slice of the TMA tensor.
    // Channel-chunk-major, tap-minor: consecutive chunks read the SAME channels of boxes shifted by one voxel, so all but the
    // first of a chunk's 27 A tiles are L2 hits (tap-major order re-read each box only after a full sweep of the channels,
    // 40-100 MB later across the 148 CTAs in flight).
    struct Chunk { int dx, dy, dz, c0, krow, bidx, v; bool from_h; };
    const int ntaps = IDENT ? tap_count(a.kind, a.ksize, cls) : 27;
    auto decode_chunk = [&](int it) {
        Chunk q = {0, 0, 0, 0, 0, b, 0, false};
        const int kc = it / ntaps, tap = it - kc * ntaps;
        if (IDENT) {                                     // kc: view 0 chunks, ..., view V-1 chunks, h chunks
            const TapRef r = decode_tap(a.kind, a.ksize, cls, tap);
            q.dx = r.dx; q.dy = r.dy; q.dz = r.dz;
            q.from_h = kc >= a.V * cx;
            if (q.from_h) { q.c0 = (kc - a.V * cx) * KC; q.krow = r.kidx * CF + a.V * a.C + q.c0; q.bidx = b * nsub + r.sub; }
            else { q.v = kc / cx; q.c0 = (kc - q.v * cx) * KC; q.krow = r.kidx * CF + q.v * a.C + q.c0; q.bidx = (b * a.V + q.v) * nsub + r.sub; }
        } else {
            q.dx = tap / 9 - 1; q.dy = (tap / 3) % 3 - 1; q.dz = tap % 3 - 1;           // W[kx][ky][kz], SAME padding
            q.from_h = kc >= cx;
            q.c0 = (q.from_h ? kc - cx : kc) * KC;
            q.krow = tap * CF + (q.from_h ? a.C : 0) + q.c0;                            // row of the K-major weight matrix
        }
        return q;
    };

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        for (int it = 0; it < nchunks; ++it) {
            const int s = it % TC_STAGES;
            const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
            mbar_wait(empty_bar(s), phase ^ 1u);
            const Chunk q = decode_chunk(it);
            const uint32_t st = smem_base + s * TC_STAGE_BYTES;
            const int xin = x0 + q.dx + a.halo_lo;                    // halo planes hold the neighbour slab's data; past them the TMA
            if (a.fused_split) {                                      // zero fill is the grid border
                mbar_expect_tx(full_bar(s), TC_A_BYTES + 2 * TC_B_BYTES);
                tma_load_5d(st, q.from_h ? &tm_hh : &tm_xh, full_bar(s), q.c0, z0 + q.dz, y0 + q.dy, xin, q.bidx);      // raw fp32 tile
            } else {
                mbar_expect_tx(full_bar(s), TC_STAGE_BYTES);
                tma_load_5d(st, q.from_h ? &tm_hh : &tm_xh, full_bar(s), q.c0, z0 + q.dz, y0 + q.dy, xin, q.bidx);
                tma_load_5d(st + TC_A_BYTES, q.from_h ? &tm_hl : &tm_xl, full_bar(s), q.c0, z0 + q.dz, y0 + q.dy, xin, q.bidx);
            }
            tma_load_2d(st + 2 * TC_A_BYTES, &tm_wh, full_bar(s), q.krow, ntile * TC_N);
            tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &tm_wl, full_bar(s), q.krow, ntile * TC_N);
        }
    } else if (!F16 && (warp == 2 || warp == 3) && a.fused_split) {
        // ===== operand converter: raw fp32 tile -> a_hi (in place) and a_lo, with the ReLU / depthwise affine in front of the conv.
        // The 128-byte swizzle only permutes 16-byte chunks inside a row, so the split is position-wise; the channel of a chunk
        // (needed for the per-channel affine) is (chunk ^ (row & 7)) * 4.
        const int ct = (warp - 2) * 32 + lane;
        for (int it = 0; it < nchunks; ++it) {
            const int s = it % TC_STAGES;
            const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
            const Chunk q = decode_chunk(it);
            const bool relu = q.from_h ? a.relu_h != 0 : a.relu_x != 0;
            const bool affine = a.pre_scale != nullptr && !q.from_h;
            const int chbase = q.v * a.C + q.c0;
            mbar_wait(full_bar(s), phase);
            const uint32_t st = smem_base + s * TC_STAGE_BYTES;
#pragma unroll 4
            for (int i = ct; i < (int)(TC_A_BYTES / 16); i += TC_CONV_WARPS * 32) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(st + 16u * i));
                if (affine) {
                    const int ch = chbase + (((i & 7) ^ ((i >> 3) & 7)) << 2);
                    const float4 sc = ldg4(a.pre_scale + ch), sh = ldg4(a.pre_shift + ch);
                    v = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
                }
                if (relu) v = relu4(v);
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(st + 16u * i), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(st + TC_A_BYTES + 16u * i), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core's async reads
            __syncwarp();
            if (lane == 0) mbar_arrive(conv_bar(s));
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer: D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  (3xTF32) =====
        for (int it = 0; it < nchunks; ++it) {
            const int s = it % TC_STAGES;
            const uint32_t phase = (uint32_t)(it / TC_STAGES) & 1u;
            const int grp = it / gsz, in_grp = it - grp * gsz;
            if (in_grp == 0 && grp > 0) {                          // the previous partial must have been promoted before it is overwritten
                mbar_wait(part_empty, (uint32_t)(grp - 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            mbar_wait((!F16 && a.fused_split) ? conv_bar(s) : full_bar(s), phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t st = smem_base + s * TC_STAGE_BYTES;
            const uint64_t dah = umma_desc_sw128(st), dal = umma_desc_sw128(st + TC_A_BYTES);
            const uint64_t dbh = umma_desc_sw128(st + 2 * TC_A_BYTES), dbl = umma_desc_sw128(st + 2 * TC_A_BYTES + TC_B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {                          // UMMA_K = 32 B (8 tf32 / 16 f16): advance the start address inside the swizzle atom
                const uint64_t adv = (uint64_t)((k * 32) >> 4);
                if (F16) {
                    umma_f16(tmem_d, dal + adv, dbh + adv, (in_grp | k) != 0);
                    umma_f16(tmem_d, dah + adv, dbl + adv, 1u);
                    umma_f16(tmem_d, dah + adv, dbh + adv, 1u);
                } else {
                    umma_tf32(tmem_d, dal + adv, dbh + adv, (in_grp | k) != 0);
                    umma_tf32(tmem_d, dah + adv, dbl + adv, 1u);
                    umma_tf32(tmem_d, dah + adv, dbh + adv, 1u);
                }
            }
            umma_commit(empty_bar(s));                              // frees the smem stage when these MMAs have read it
            if (in_grp == gsz - 1 || it == nchunks - 1) umma_commit(part_full);   // partial accumulator complete
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> LSTM gates -> c, h =====
        const int q = warp - 4;                                     // TMEM lane quadrant of this warp (warp % 4)
        const uint32_t tpart = tmem_d + ((uint32_t)(q * 32) << 16), tmast = tpart + TC_N;
        // promotion: master (+)= partial with round-to-nearest fp32 adds; the last group is folded into the gate epilogue
        for (int grp = 0; grp + 1 < ngroups; ++grp) {
            mbar_wait(part_full, (uint32_t)grp & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c32 = 0; c32 < TC_N; c32 += 32) {              // 32 columns per round trip: 4 loads in flight, one wait
                float pa[16], pb[16], ma[16], mb[16];
                tmem_ld16(tpart + c32, pa); tmem_ld16(tpart + c32 + 16, pb);
                if (grp > 0) { tmem_ld16(tmast + c32, ma); tmem_ld16(tmast + c32 + 16, mb); }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (grp > 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { pa[i] += ma[i]; pb[i] += mb[i]; }
                }
                tmem_st16(tmast + c32, pa); tmem_st16(tmast + c32 + 16, pb);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(part_empty);
        }
        mbar_wait(part_full, (uint32_t)(ngroups - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int m = q * 32 + lane;                                // accumulator row = voxel of the box, z fastest
        const int bz = m % a.BZ, by = (m / a.BZ) % a.BY, bx = m / (a.BZ * a.BY);
        const int x = x0 + bx, y = y0 + by, z = z0 + bz;
        const bool ok = x < a.X && y < a.Y && z < a.Z;
        const long long vox = (((long long)b * a.X + x) * a.Y + y) * a.Z + z;
        const long long voxh = (((long long)b * (a.X + a.halo_lo + a.halo_hi) + x + a.halo_lo) * a.Y + y) * a.Z + z;   // h_out keeps the halo planes
        const float inv = F16 ? __ldg(a.inv_scale_a) * __ldg(a.inv_scale_w) : 1.0f;         // exact powers of two
        if (IDENT) {
            // out[b, vox_out, co] = relu(bn(acc + bias))     (model_multi.py:449-455, :418-441)
            long long ovox = vox;
            if (a.kind == MVF_DECONV_S2)
                ovox = (((long long)b * (2 * a.X) + 2 * x + ((cls >> 2) & 1)) * (2 * a.Y) + 2 * y + ((cls >> 1) & 1)) * (2 * a.Z) + 2 * z + (cls & 1);
#pragma unroll 1
            for (int c16 = 0; c16 < TC_N; c16 += 16) {
                float v[16];
                tmem_ld16(tpart + c16, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ngroups > 1) {
                    float mv[16];
                    tmem_ld16(tmast + c16, mv);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += mv[i];
                }
                const int co0 = ntile * TC_N + c16;
                if (ok && co0 < a.Cout) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float y = (F16 ? v[i] * inv : v[i]) + a.bias[co0 + i];
                        if (a.bn_scale) y = fmaf(y, a.bn_scale[co0 + i], a.bn_shift[co0 + i]);
                        v[i] = a.relu_out ? fmaxf(y, 0.f) : y;
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 4) st4(a.out + ovox * a.Cout + co0 + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
                }
            }
        } else {
        const int fbase = ntile * TC_FPT;
    #pragma unroll 1
            for (int g16 = 0; g16 < TC_FPT / 16; ++g16) {
                float gj[16], gi[16], gf[16], go[16];
                const uint32_t ta = tpart + (uint32_t)(g16 * 16);
                tmem_ld16(ta + 0 * TC_FPT, gj); tmem_ld16(ta + 1 * TC_FPT, gi);      // gate order j,i,f,o  (recurrent.py:460-461)
                tmem_ld16(ta + 2 * TC_FPT, gf); tmem_ld16(ta + 3 * TC_FPT, go);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ngroups > 1) {                                       // last partial + master
                    float mj[16], mi[16], mf[16], mo[16];
                    tmem_ld16(ta + TC_N + 0 * TC_FPT, mj); tmem_ld16(ta + TC_N + 1 * TC_FPT, mi);
                    tmem_ld16(ta + TC_N + 2 * TC_FPT, mf); tmem_ld16(ta + TC_N + 3 * TC_FPT, mo);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
                    for (int i = 0; i < 16; ++i) { gj[i] += mj[i]; gi[i] += mi[i]; gf[i] += mf[i]; go[i] += mo[i]; }
                }
                if (ok) {
                    const int f0 = fbase + g16 * 16;
                    float cp[16];
    #pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 v = a.c_prev ? ldg4(a.c_prev + vox * a.F + f0 + i) : zero4();
                        cp[i] = v.x; cp[i + 1] = v.y; cp[i + 2] = v.z; cp[i + 3] = v.w;
                    }
                    float cn[16], hn[16];
    #pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int f = f0 + i;
                        if (F16) { gj[i] *= inv; gi[i] *= inv; gf[i] *= inv; go[i] *= inv; }
                        const float vj = gj[i] + a.bias[0 * a.F + f], vi = gi[i] + a.bias[1 * a.F + f];
                        const float vf = gf[i] + a.bias[2 * a.F + f], vo = go[i] + a.bias[3 * a.F + f];
                        const float c = cp[i] * sigmoid_acc(vf + a.forget_bias) + sigmoid_acc(vi) * tanhf(vj);   // :470-472
                        cn[i] = c; hn[i] = tanhf(c) * sigmoid_acc(vo);                                           // :477
                    }
    #pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        st4(a.c_out + vox * a.F + f0 + i, make_float4(cn[i], cn[i + 1], cn[i + 2], cn[i + 3]));
                        st4(a.h_out + voxh * a.F + f0 + i, make_float4(hn[i], hn[i + 1], hn[i + 2], hn[i + 3]));
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(2 * TC_N) : "memory");
    }
}

// a -> (a_hi, a_lo): a_hi keeps the top 19 bits (exactly representable in TF32), a_lo = a - a_hi (exact in fp32).
__global__ void __launch_bounds__(256)
tf32_split_kernel(const float4* __restrict__ in, float4* __restrict__ hi, float4* __restrict__ lo, long long n4, int relu) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = __ldg(in + i);
    if (relu) v = relu4(v);
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    hi[i] = h; lo[i] = l;
}

// W [27*(C+F), 4F] (reference layout) -> K-major [4F, 27*(C+F)] hi and lo, rows permuted so that the four gates of a
// 64-filter group are adjacent: row n' = (f / 64) * 256 + gate * 64 + f % 64  <-  column gate * F + f.
__global__ void __launch_bounds__(256)
convlstm_prepare_kernel(const float* __restrict__ W, float* __restrict__ whi, float* __restrict__ wlo, int K, int F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // output element, k fastest
    const long long total = (long long)K * 4 * F;
    if (i >= total) return;
    const int k = (int)(i % K);
    const int np = (int)(i / K);
    const int grp = np / TC_N, gate = (np % TC_N) / TC_FPT, fl = np % TC_FPT;
    const float v = W[(long long)k * 4 * F + gate * F + grp * TC_FPT + fl];
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    whi[i] = h; wlo[i] = v - h;
}

// General activation split for the conv3d family: v = relu?(in * pre_scale[ch] + pre_shift[ch]) -> (hi, lo), optionally
// re-laid out as the 8 parity sub-lattices of a stride-2 conv:  in [NB, X, Y, Z, C] -> out [NB, 8, X/2, Y/2, Z/2, C].
// pre_scale / pre_shift (NULL = identity) are indexed by (v * C + c) with v = (nb % V): the per-input-channel affine of a
// depthwise 1x1 conv in front of the GEMM (model_multi.py:472,477).
__global__ void __launch_bounds__(256)
act_split_kernel(const float4* __restrict__ in, float4* __restrict__ hi, float4* __restrict__ lo, long long n4,
                 int X, int Y, int Z, int C4, int V, int relu, int s2d, const float4* __restrict__ pre_scale, const float4* __restrict__ pre_shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = __ldg(in + i);
    long long t = i;
    const int c4 = (int)(t % C4); t /= C4;
    const int z = (int)(t % Z); t /= Z;
    const int y = (int)(t % Y); t /= Y;
    const int x = (int)(t % X); t /= X;                         // t = nb
    if (pre_scale) {
        const int ch = (int)(t % V) * C4 + c4;
        const float4 sc = __ldg(pre_scale + ch), sh = __ldg(pre_shift + ch);
        v = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
    }
    if (relu) v = relu4(v);
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    long long o = i;
    if (s2d) {
        const int sub = ((x & 1) * 2 + (y & 1)) * 2 + (z & 1);
        o = ((((t * 8 + sub) * (X / 2) + (x >> 1)) * (Y / 2) + (y >> 1)) * (Z / 2) + (z >> 1)) * C4 + c4;
    }
    hi[o] = h; lo[o] = l;
}

// W [K, N] (reference layout, N fastest) -> K-major [N, K] hi and lo halves.  `transposed`: the Conv3DTranspose kernel
// layout [taps, N, Cin] (Keras: kernel_size + (filters, input_dim)), K index = tap * Cin + ci.  `S` > 1: the reference's
// input channel order is (c * S + s) (depth_sampling's reshape, model_multi.py:468-470) while the activations arrive as S
// sources of C channels: row k = s * C + c reads reference row c * S + s.
__global__ void __launch_bounds__(256)
weight_split_kernel(const float* __restrict__ W, float* __restrict__ whi, float* __restrict__ wlo, int K, int N, int transposed, int Cin, int S) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // output element, k fastest
    if (i >= (long long)K * N) return;
    const int k = (int)(i % K), n = (int)(i / K);
    long long src;
    if (transposed) { const int tap = k / Cin, ci = k - tap * Cin; src = ((long long)tap * N + n) * Cin + ci; }
    else if (S > 1) { const int tap = k / Cin, r = k - tap * Cin, Cc = Cin / S, sidx = r / Cc, c = r - sidx * Cc; src = ((long long)tap * Cin + c * S + sidx) * N + n; }
    else src = (long long)k * N + n;
    const float v = W[src];
    const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    whi[i] = h; wlo[i] = v - h;
}

// ---- fp16 operand split --------------------------------------------------------------------------------------------
// a -> a1 + a2 with a1 = fp16(a * 2^s), a2 = fp16(a * 2^s - a1) (round to nearest): 22 mantissa bits in two halves that the
// tensor core multiplies exactly and accumulates in fp32 -- a1*b1 + a1*b2 + a2*b1 at the f16 MMA rate (twice the tf32 rate) and,
// with rounding instead of truncation, ~5x closer to fp32 than the tf32 split.  s = 14 - exponent(max|a|) per tensor (one
// max-reduction pass, no host sync) keeps a1 below 2^14 and pushes a2 well into the normal fp16 range; the GEMM epilogue
// multiplies by 2^-(s_a + s_w).
__global__ void __launch_bounds__(256)
amax_kernel(const float4* __restrict__ in, long long n4, int relu, unsigned* __restrict__ amax_bits,
            const float4* __restrict__ pre_scale = nullptr, const float4* __restrict__ pre_shift = nullptr, long long inner4 = 1, int V = 1, int C4 = 1) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = __ldg(in + i);
        if (pre_scale) {                                  // per-input-channel affine (depthwise 1x1) in front of the conv: channel (v, c)
            const int ch = (int)((i / inner4) % V) * C4 + (int)(i % C4);
            const float4 sc = __ldg(pre_scale + ch), sh = __ldg(pre_shift + ch);
            v = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
        }
        if (relu) v = relu4(v);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));     // non-negative floats order like their bit patterns
}
// activations: same indexing as act_split_kernel (optional ReLU, optional parity sub-lattice re-layout); tail[0] = amax bits, tail[1] <- 2^-s
__global__ void __launch_bounds__(256)
act_split_f16_kernel(const float4* __restrict__ in, uint2* __restrict__ hi, uint2* __restrict__ lo, long long n4,
                     int X, int Y, int Z, int C4, int relu, int s2d, unsigned* __restrict__ tail,
                     const float4* __restrict__ pre_scale = nullptr, const float4* __restrict__ pre_shift = nullptr, int V = 1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float inv;
    const float scale = pow2_scale(tail[0], &inv);
    if (i == 0) reinterpret_cast<float*>(tail)[1] = inv;
    if (i >= n4) return;
    float4 v = __ldg(in + i);
    if (pre_scale) {
        const int ch = (int)((i / ((long long)X * Y * Z * C4)) % V) * C4 + (int)(i % C4);
        const float4 sc = __ldg(pre_scale + ch), sh = __ldg(pre_shift + ch);
        v = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
    }
    if (relu) v = relu4(v);
    long long o = i;
    if (s2d) {
        long long t = i;
        const int c4 = (int)(t % C4); t /= C4;
        const int z = (int)(t % Z); t /= Z;
        const int y = (int)(t % Y); t /= Y;
        const int x = (int)(t % X); t /= X;
        const int sub = ((x & 1) * 2 + (y & 1)) * 2 + (z & 1);
        o = ((((t * 8 + sub) * (X / 2) + (x >> 1)) * (Y / 2) + (y >> 1)) * (Z / 2) + (z >> 1)) * C4 + c4;
    }
    uint2 h, l;
    split_half4(v, scale, &h, &l);
    hi[o] = h; lo[o] = l;
}
// weights -> K-major [N, K] fp16 halves.  mode 0: W [K, N];  1: Conv3DTranspose [taps, N, Cin];  2: ConvLSTM gate permutation
// (row n' = (f / 64) * 256 + gate * 64 + f % 64  <-  column gate * F + f, F = N / 4)
__global__ void __launch_bounds__(256)
weight_split_f16_kernel(const float* __restrict__ W, __half* __restrict__ whi, __half* __restrict__ wlo, int K, int N, int mode, int Cin,
                        unsigned* __restrict__ tail, int S = 1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // output element, k fastest
    float inv;
    const float scale = pow2_scale(tail[0], &inv);
    if (i == 0) reinterpret_cast<float*>(tail)[1] = inv;
    if (i >= (long long)K * N) return;
    const int k = (int)(i % K), n = (int)(i / K);
    long long src;
    if (mode == 1) { const int tap = k / Cin, ci = k - tap * Cin; src = ((long long)tap * N + n) * Cin + ci; }
    else if (mode == 2) { const int F = N / 4, grp = n / TC_N, gate = (n % TC_N) / TC_FPT, fl = n % TC_FPT; src = (long long)k * N + gate * F + grp * TC_FPT + fl; }
    else if (mode == 3) { const int tap = k / Cin, r = k - tap * Cin, Cc = Cin / S, sidx = r / Cc, c = r - sidx * Cc; src = ((long long)tap * Cin + c * S + sidx) * N + n; }   // row s*C+c <- reference row c*S+s
    else src = (long long)k * N + n;
    const float v = W[src] * scale;
    const __half a1 = __float2half_rn(v);
    whi[i] = a1; wlo[i] = __float2half_rn(v - __half2float(a1));
}

// ---- host: TMA descriptors through the driver entry point (no link-time dependency on libcuda) ----
static bool make_act_map(CUtensorMap* tm, const void* base, int B, int X, int Y, int Z, int C, int BX, int BY, int BZ, bool f16 = false) {
    const cuuint64_t es = f16 ? 2 : 4;
    const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)Z, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)B};
    const cuuint64_t strides[4] = {(cuuint64_t)C * es, (cuuint64_t)Z * C * es, (cuuint64_t)Y * Z * C * es, (cuuint64_t)X * Y * Z * C * es};
    const cuuint32_t box[5] = {(cuuint32_t)(f16 ? TC_K16 : TC_K), (cuuint32_t)BZ, (cuuint32_t)BY, (cuuint32_t)BX, 1u};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return encode_tiled()(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool make_w_map(CUtensorMap* tm, const void* base, int K, int N, bool f16 = false) {
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    const cuuint64_t strides[1] = {(cuuint64_t)K * (f16 ? 2 : 4)};
    const cuuint32_t box[2] = {(cuuint32_t)(f16 ? TC_K16 : TC_K), (cuuint32_t)TC_N};
    const cuuint32_t estr[2] = {1, 1};
    return encode_tiled()(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// Operand format of the k=3 convolutions: fp16 halves (twice the MMA rate) whenever every source has a multiple of 64 channels.
// MVF_TC_TF32=1 forces the tf32 split everywhere (A/B measurement); the fused in-kernel converter is tf32 only.
static bool env_flag(const char* name) { return env_int(name, 0) != 0; }
// ONE predicate decides the fp16 weight / activation format, at prepare time and at call time alike (the debug switches are
// read once, together; in the product build they do not exist).
static bool want_f16(int C, int C2) {
    static const bool off = env_flag("MVF_TC_TF32") || env_flag("MVF_TC_FUSED_SPLIT");
    return !off && C % TC_K16 == 0 && C2 % TC_K16 == 0;
}
static unsigned amax_grid(long long n4) { const long long b = (n4 + 255) / 256; return (unsigned)(b < 148 * 8 ? (b > 0 ? b : 1) : 148 * 8); }

}  // namespace mvf

using namespace mvf;

extern "C" size_t mvf_convlstm_wsplit_bytes(int C, int F) {
    if (C <= 0 || F <= 0) return 0;
    return (size_t)2 * 27 * (size_t)(C + F) * 4 * F * sizeof(float);
}

extern "C" int mvf_convlstm_prepare(const float* W, int C, int F, float* wsplit, void* stream) {
    if (!W || !wsplit) return MVF_ENULL;
    if (C <= 0 || F <= 0) return MVF_EINVAL;
    if (C % TC_K != 0 || F % TC_FPT != 0) return MVF_EUNSUPPORTED;
    const int K = 27 * (C + F);
    const long long total = (long long)K * 4 * F;
    cudaStream_t s = (cudaStream_t)stream;
    if (want_f16(C, F)) {                               // [hi K*4F halves][lo K*4F halves][amax bits, 2^-s]
        __half* whi = (__half*)wsplit;
        unsigned* tail = (unsigned*)(whi + 2 * total);
        if (cudaMemsetAsync(tail, 0, 8, s) != cudaSuccess) return MVF_ECUDA;
        amax_kernel<<<amax_grid(total / 4), 256, 0, s>>>((const float4*)W, total / 4, 0, tail);
        weight_split_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(W, whi, whi + total, K, 4 * F, 2, C + F, tail);
        count_launch(2);
        return check_launch();
    }
    convlstm_prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(W, wsplit, wsplit + total, K, F);
    count_launch();
    return check_launch();
}

extern "C" size_t mvf_convlstm_tc_workspace_bytes(int B, int X, int Y, int Z, int C, int F) {
    if (B <= 0 || X <= 0 || Y <= 0 || Z <= 0 || C <= 0 || F <= 0) return 0;
    const size_t vox = (size_t)B * X * Y * Z;
    return 2 * vox * (size_t)(C + F) * sizeof(float) + 256;           // x_hi, x_lo, h_hi, h_lo (+ the scale cell of the fp16 split)
}

extern "C" int mvf_convlstm_step_tc_slab(const float* x, const float* h_prev, const float* c_prev, const float* wsplit,
                                         const float* bias, float forget_bias, int B, int X, int Y, int Z, int C, int F,
                                         int halo_lo, int halo_hi, int flags, float* h_out, float* c_out,
                                         void* ws, size_t ws_bytes, const float* act_amax, void* stream) {
    if (!x || !wsplit || !bias || !h_out || !c_out) return MVF_ENULL;
    if (B <= 0 || X <= 0 || Y <= 0 || Z <= 0 || C <= 0 || F <= 0) return MVF_EINVAL;
    if (halo_lo < 0 || halo_lo > 1 || halo_hi < 0 || halo_hi > 1) return MVF_EINVAL;
    const int Xin = X + halo_lo + halo_hi;
    if ((h_prev == nullptr) != (c_prev == nullptr)) return MVF_ENULL;
    if (h_out == h_prev || c_out == h_prev) return MVF_EINVAL;
    if (C % TC_K != 0 || F % TC_FPT != 0) return MVF_EUNSUPPORTED;
    if (!aligned16(x) || !aligned16(wsplit) || (ws && !aligned16(ws)) || !aligned16(h_out) || !aligned16(c_out) ||
        (h_prev && (!aligned16(h_prev) || !aligned16(c_prev)))) return MVF_EALIGN;
    // The hi/lo halves of x and h are produced by a separate elementwise pass: every element is read 27 x 4F/256 times by the
    // GEMM, so the pass costs 1 % while converting inside the 2-stage ring costs 10 % (measured: 35.7 vs 32.3 ms at c3 size).
    // MVF_TC_FUSED_SPLIT=1 selects the in-kernel converter for A/B measurement.
    static const bool split_pass = !env_flag("MVF_TC_FUSED_SPLIT");
    if (split_pass && (!ws || ws_bytes < mvf_convlstm_tc_workspace_bytes(B, Xin, Y, Z, C, F))) return MVF_EWORKSPACE;
    if (!encode_tiled()) return MVF_ECUDA;
    cudaStream_t s = (cudaStream_t)stream;
    const long long vox = (long long)B * Xin * Y * Z;
    const bool f16 = split_pass && want_f16(C, F);
    const void *xh = x, *xl = x, *hh = h_prev ? h_prev : x, *hl = hh;
    const float* inv_a = nullptr;
    const int relu_in = (flags & MVF_FLAG_RELU_IN) != 0;
    if (f16) {
        __half* w0 = (__half*)ws;
        __half* w1 = w0 + vox * C; __half* w2 = w1 + vox * C; __half* w3 = w2 + vox * F;
        unsigned* tail = (unsigned*)(((uintptr_t)(w3 + vox * F) + 15) & ~(uintptr_t)15);
        const long long n4 = vox * C / 4, m4 = vox * F / 4;
        if (act_amax) {                                  // the caller's max(|relu?(x)|, |h|): slabs of one grid must agree on the scale
            if (cudaMemcpyAsync(tail, act_amax, 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return MVF_ECUDA;
        } else {
            if (cudaMemsetAsync(tail, 0, 8, s) != cudaSuccess) return MVF_ECUDA;
            amax_kernel<<<amax_grid(n4), 256, 0, s>>>((const float4*)x, n4, relu_in, tail);
            if (h_prev) amax_kernel<<<amax_grid(m4), 256, 0, s>>>((const float4*)h_prev, m4, 0, tail);   // x and h share the accumulator: one scale
            count_launch(h_prev ? 2 : 1);
        }
        act_split_f16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>((const float4*)x, (uint2*)w0, (uint2*)w1, n4, Xin, Y, Z, C / 4, relu_in, 0, tail);
        if (h_prev) act_split_f16_kernel<<<(unsigned)((m4 + 255) / 256), 256, 0, s>>>((const float4*)h_prev, (uint2*)w2, (uint2*)w3, m4, Xin, Y, Z, F / 4, 0, 0, tail);
        count_launch(h_prev ? 2 : 1);
        xh = w0; xl = w1; hh = h_prev ? (const void*)w2 : (const void*)w0; hl = h_prev ? (const void*)w3 : (const void*)w1;
        inv_a = (const float*)tail + 1;
    } else if (split_pass) {
        float* w0 = (float*)ws;
        float* w1 = w0 + vox * C; float* w2 = w1 + vox * C; float* w3 = w2 + vox * F;
        const long long n4 = vox * C / 4;
        tf32_split_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>((const float4*)x, (float4*)w0, (float4*)w1, n4, relu_in);
        count_launch();
        if (h_prev) {
            const long long m4 = vox * F / 4;
            tf32_split_kernel<<<(unsigned)((m4 + 255) / 256), 256, 0, s>>>((const float4*)h_prev, (float4*)w2, (float4*)w3, m4, 0);
            count_launch();
        }
        xh = w0; xl = w1; hh = w2; hl = w3;
    }
    TcArgs a;
    a.bias = bias; a.c_prev = c_prev; a.h_out = h_out; a.c_out = c_out;
    a.B = B; a.X = X; a.Y = Y; a.Z = Z; a.C = C; a.F = F;
    a.BZ = pow2ceil(Z) < TC_M ? pow2ceil(Z) : TC_M;
    a.BY = pow2ceil(Y) < TC_M / a.BZ ? pow2ceil(Y) : TC_M / a.BZ;
    a.BX = TC_M / (a.BZ * a.BY);
    a.tiles_z = (Z + a.BZ - 1) / a.BZ; a.tiles_y = (Y + a.BY - 1) / a.BY; a.tiles_x = (X + a.BX - 1) / a.BX;
    a.has_h = h_prev != nullptr; a.forget_bias = forget_bias;
    a.halo_lo = halo_lo; a.halo_hi = halo_hi;
    a.fused_split = !split_pass; a.relu_x = relu_in; a.relu_h = 0; a.pre_scale = nullptr; a.pre_shift = nullptr;
    a.V = 1; a.Cout = 0; a.kind = 0; a.ksize = 3; a.relu_out = 0; a.bn_scale = nullptr; a.bn_shift = nullptr; a.out = nullptr;
    // K-chunks (of 32) per partial accumulation chain; MVF_TC_PROMOTE overrides (0 = one long chain, for A/B measurement)
    static const int promote_env = env_int("MVF_TC_PROMOTE", -1);
    a.promote = promote_env >= 0 ? promote_env : 8;
    const int K = 27 * (C + F);
    const long long wtotal = (long long)K * 4 * F;
    const void *whi = wsplit, *wlo = wsplit + wtotal;
    a.inv_scale_a = inv_a; a.inv_scale_w = nullptr;
    if (f16) { whi = wsplit; wlo = (const __half*)wsplit + wtotal; a.inv_scale_w = (const float*)((const __half*)wsplit + 2 * wtotal) + 1; }
    CUtensorMap tm_xh, tm_xl, tm_hh, tm_hl, tm_wh, tm_wl;
    bool ok = make_act_map(&tm_xh, xh, B, Xin, Y, Z, C, a.BX, a.BY, a.BZ, f16) && make_act_map(&tm_xl, xl, B, Xin, Y, Z, C, a.BX, a.BY, a.BZ, f16) &&
              make_act_map(&tm_hh, hh, B, Xin, Y, Z, h_prev ? F : C, a.BX, a.BY, a.BZ, f16) &&
              make_act_map(&tm_hl, hl, B, Xin, Y, Z, h_prev ? F : C, a.BX, a.BY, a.BZ, f16) &&
              make_w_map(&tm_wh, whi, K, 4 * F, f16) && make_w_map(&tm_wl, wlo, K, 4 * F, f16);
    if (!ok) return MVF_ECUDA;
    const long long mtiles = (long long)B * a.tiles_x * a.tiles_y * a.tiles_z;
    if (mtiles > 2147483647ll || F / TC_FPT > 65535) return MVF_EUNSUPPORTED;
    dim3 grid((unsigned)mtiles, F / TC_FPT);
    // per-device attribute: set on every call (a process may drive several GPUs)
    if (f16) {
        if (cudaFuncSetAttribute(convlstm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess) return MVF_ECUDA;
        convlstm_tc_kernel<false, true><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(tm_xh, tm_xl, tm_hh, tm_hl, tm_wh, tm_wl, a);
    } else {
        if (cudaFuncSetAttribute(convlstm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess) return MVF_ECUDA;
        convlstm_tc_kernel<false, false><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(tm_xh, tm_xl, tm_hh, tm_hl, tm_wh, tm_wl, a);
    }
    count_launch();
    return check_launch();
}

extern "C" int mvf_convlstm_step_tc(const float* x, const float* h_prev, const float* c_prev, const float* wsplit,
                                    const float* bias, float forget_bias, int B, int X, int Y, int Z, int C, int F,
                                    int flags, float* h_out, float* c_out, void* ws, size_t ws_bytes, void* stream) {
    return mvf_convlstm_step_tc_slab(x, h_prev, c_prev, wsplit, bias, forget_bias, B, X, Y, Z, C, F, 0, 0, flags, h_out, c_out,
                                     ws, ws_bytes, nullptr, stream);
}

// ---- the plain conv3d family on the tensor cores: grid_reas 'ident' / 'conv3d' (model_multi.py:406-455), depth_sampling
// 'conv3d' branch (:467-480) ---------------------------------------------------------------------------------------
static int conv_taps(int kind, int ksize) { return (kind == MVF_CONV_S1 && ksize == 1) ? 1 : 27; }

extern "C" size_t mvf_conv3d_wsplit_bytes(int kind, int ksize, int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0 || (ksize != 1 && ksize != 3)) return 0;
    // k = 1: both operand formats (tf32 for the fused converter, fp16 for the split-pass path, chosen per call); k = 3: one of them
    return (size_t)(ksize == 1 ? 3 : 2) * conv_taps(kind, ksize) * Cin * Cout * sizeof(float) + 512;
}

static bool conv3d_f16(int ksize, int C, int C2, bool pre_affine) { return ksize == 3 && !pre_affine && want_f16(C, C2); }
static bool conv3d_f16_presplit(int C, int C2) { return want_f16(C, C2); }

extern "C" int mvf_conv3d_prepare(const float* W, int kind, int ksize, int V, int C, int C2, int Cout, int chan_interleave,
                                  float* wsplit, void* stream) {
    if (!W || !wsplit) return MVF_ENULL;
    const bool for_presplit = chan_interleave == -1;      // weights for mvf_conv3d_tc(MVF_FLAG_PRESPLIT): always the fp16 format
    if (for_presplit) chan_interleave = 0;
    if (V <= 0 || C <= 0 || C2 < 0 || Cout <= 0 || chan_interleave < 0) return MVF_EINVAL;
    if (for_presplit && !conv3d_f16_presplit(C, C2)) return MVF_EUNSUPPORTED;
    if (kind < MVF_CONV_S1 || kind > MVF_DECONV_S2 || (ksize != 1 && ksize != 3) || (kind != MVF_CONV_S1 && ksize != 3)) return MVF_EINVAL;
    if (C % TC_K != 0 || C2 % TC_K != 0 || Cout % 16 != 0) return MVF_EUNSUPPORTED;
    const int Cin = V * C + C2;
    if (chan_interleave > 1 && (ksize != 1 || Cin % chan_interleave != 0)) return MVF_EINVAL;
    const int K = conv_taps(kind, ksize) * Cin;
    const long long total = (long long)K * Cout;
    cudaStream_t s = (cudaStream_t)stream;
    if (for_presplit || conv3d_f16(ksize, C, C2, false)) {       // [hi K*Cout halves][lo][amax bits, 2^-s]
        __half* whi = (__half*)wsplit;
        unsigned* tail = (unsigned*)(whi + 2 * total);
        if (cudaMemsetAsync(tail, 0, 8, s) != cudaSuccess) return MVF_ECUDA;
        amax_kernel<<<amax_grid(total / 4), 256, 0, s>>>((const float4*)W, total / 4, 0, tail);
        weight_split_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(W, whi, whi + total, K, Cout, kind == MVF_DECONV_S2 ? 1 : 0, Cin, tail);
        count_launch(2);
        return check_launch();
    }
    weight_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(W, wsplit, wsplit + total, K, Cout,
                                                                      kind == MVF_DECONV_S2, Cin, chan_interleave);
    count_launch();
    if (ksize == 1 && want_f16(C, C2)) {                  // second copy for small 1x1x1 problems, which take the split-pass path at the f16 rate
        __half* whi = (__half*)(wsplit + 2 * total);
        unsigned* tail = (unsigned*)(whi + 2 * total);
        if (cudaMemsetAsync(tail, 0, 8, s) != cudaSuccess) return MVF_ECUDA;
        amax_kernel<<<amax_grid(total / 4), 256, 0, s>>>((const float4*)W, total / 4, 0, tail);
        weight_split_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(W, whi, whi + total, K, Cout, chan_interleave > 1 ? 3 : 0, Cin, tail,
                                                                               chan_interleave > 1 ? chan_interleave : 1);
        count_launch(2);
    }
    return check_launch();
}

// Where the activation split happens.  A separate elementwise pass (workspace = hi and lo halves of both sources) is the
// default: with k=3 every element is read 27 times by the GEMM, and small problems are latency-bound chains that the extra
// converter stage lengthens.  The 1x1x1 members at scale read every activation exactly once and are HBM-bound, so they convert
// inside the GEMM (warps 2-3) and touch no workspace: 'ident' 2048->256 at 64^3 1.67 ms instead of 2.60.
// MVF_TC_SPLIT_PASS=1 / MVF_TC_FUSED_SPLIT=1 force either choice (the stride-2 conv always needs the pass: it re-lays the operand).
static bool conv3d_fused_split(int kind, int ksize, int B, int X, int Y, int Z, int Cout) {
    static const int force = env_flag("MVF_TC_SPLIT_PASS") ? 1 : env_flag("MVF_TC_FUSED_SPLIT") ? 2 : 0;
    if (kind == MVF_CONV_S2 || force == 1) return false;
    if (force == 2) return true;
    const long long tiles = (((long long)B * X * Y * Z + TC_M - 1) / TC_M) * ((Cout + TC_N - 1) / TC_N);
    return kind == MVF_CONV_S1 && ksize == 1 && tiles >= 2 * 148;
}

extern "C" size_t mvf_conv3d_tc_workspace_bytes(int kind, int ksize, int B, int V, int X, int Y, int Z, int C, int C2, int Cout) {
    if (B <= 0 || V <= 0 || X <= 0 || Y <= 0 || Z <= 0 || C <= 0 || C2 < 0 || Cout <= 0) return 0;
    if (conv3d_fused_split(kind, ksize, B, X, Y, Z, Cout)) return 0;
    return (size_t)2 * B * X * Y * Z * ((size_t)V * C + C2) * sizeof(float) + 256;
}

extern "C" int mvf_conv3d_tc(const float* in, const float* in2, const float* wsplit, const float* bias,
                             const float* bn_scale, const float* bn_shift, const float* pre_scale, const float* pre_shift,
                             int kind, int ksize, int B, int V, int X, int Y, int Z, int C, int C2, int Cout, int flags,
                             float* out, void* ws, size_t ws_bytes, const float* act_amax, void* stream) {
    const bool presplit = (flags & MVF_FLAG_PRESPLIT) != 0;      // the workspace already holds the fp16 halves of `in` (+ scale)
    if ((!in && !presplit) || !wsplit || !bias || !out) return MVF_ENULL;
    if ((bn_scale == nullptr) != (bn_shift == nullptr) || (pre_scale == nullptr) != (pre_shift == nullptr)) return MVF_ENULL;
    if ((in2 == nullptr) != (C2 == 0)) return MVF_EINVAL;
    if (presplit && (in2 || pre_scale || kind == MVF_DECONV_S2)) return MVF_EUNSUPPORTED;
    if (B <= 0 || V <= 0 || X <= 0 || Y <= 0 || Z <= 0 || C <= 0 || C2 < 0 || Cout <= 0) return MVF_EINVAL;
    if (kind < MVF_CONV_S1 || kind > MVF_DECONV_S2 || (ksize != 1 && ksize != 3) || (kind != MVF_CONV_S1 && ksize != 3)) return MVF_EINVAL;
    if (C % TC_K != 0 || C2 % TC_K != 0 || Cout % 16 != 0) return MVF_EUNSUPPORTED;
    if (kind == MVF_CONV_S2 && ((X | Y | Z) & 1)) return MVF_EUNSUPPORTED;      // odd sizes pad on both sides in TF; not built
    if (pre_scale && in2) return MVF_EUNSUPPORTED;
    if ((in && !aligned16(in)) || !aligned16(wsplit) || (ws && !aligned16(ws)) || !aligned16(out) || (in2 && !aligned16(in2)) ||
        (pre_scale && (!aligned16(pre_scale) || !aligned16(pre_shift)))) return MVF_EALIGN;
    if (pre_scale && ksize != 1) return MVF_EUNSUPPORTED;                       // the affine must not touch the SAME padding
    const int relu_in = (flags & MVF_FLAG_RELU_IN) != 0, s2d = kind == MVF_CONV_S2;
    const bool split_pass = presplit || !conv3d_fused_split(kind, ksize, B, X, Y, Z, Cout);
    if (split_pass && !presplit && (!ws || ws_bytes < mvf_conv3d_tc_workspace_bytes(kind, ksize, B, V, X, Y, Z, C, C2, Cout))) return MVF_EWORKSPACE;
    if (!encode_tiled()) return MVF_ECUDA;
    cudaStream_t s = (cudaStream_t)stream;
    const long long n1 = (long long)B * V * X * Y * Z * C, n2 = (long long)B * X * Y * Z * C2;
    // k = 3: fp16 halves whenever the channel counts allow; k = 1 on the split-pass path (small problems): the fp16 copy of the
    // weights that mvf_conv3d_prepare wrote behind the tf32 one (with the depthwise affine applied by the split pass)
    const bool f16_k1 = !presplit && split_pass && ksize == 1 && want_f16(C, C2);
    const bool f16 = presplit ? conv3d_f16_presplit(C, C2) : (f16_k1 || (split_pass && conv3d_f16(ksize, C, C2, pre_scale != nullptr)));
    if (presplit && (!f16 || !ws || ws_bytes < 4 * (size_t)n1 + 256)) return presplit && !f16 ? MVF_EUNSUPPORTED : MVF_EWORKSPACE;
    const void *xh = in, *xl = in, *hh = in2, *hl = in2;
    const float* inv_a = nullptr;
    if (f16 && presplit) {                               // written by mvf_unproject_split_f16
        __half* w0 = (__half*)ws;
        xh = w0; xl = w0 + n1; hh = xh; hl = xl;
        inv_a = (const float*)(((uintptr_t)(w0 + 2 * n1) + 15) & ~(uintptr_t)15) + 1;
    } else if (f16) {
        __half* w0 = (__half*)ws; __half* w1 = w0 + n1; __half* w2 = w1 + n1; __half* w3 = w2 + n2;
        unsigned* tail = (unsigned*)(((uintptr_t)(w3 + n2) + 15) & ~(uintptr_t)15);
        if (act_amax) {                                  // the caller's bound on max|operand| (e.g. max|features| for unprojected grids)
            if (cudaMemcpyAsync(tail, act_amax, 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return MVF_ECUDA;
        } else {
            if (cudaMemsetAsync(tail, 0, 8, s) != cudaSuccess) return MVF_ECUDA;
            amax_kernel<<<amax_grid(n1 / 4), 256, 0, s>>>((const float4*)in, n1 / 4, relu_in, tail, (const float4*)pre_scale, (const float4*)pre_shift,
                                                          (long long)X * Y * Z * (C / 4), V, C / 4);
            if (in2) amax_kernel<<<amax_grid(n2 / 4), 256, 0, s>>>((const float4*)in2, n2 / 4, relu_in, tail);  // sources share the accumulator: one scale
            count_launch(in2 ? 2 : 1);
        }
        act_split_f16_kernel<<<(unsigned)((n1 / 4 + 255) / 256), 256, 0, s>>>((const float4*)in, (uint2*)w0, (uint2*)w1, n1 / 4, X, Y, Z, C / 4, relu_in, s2d, tail,
                                                                            (const float4*)pre_scale, (const float4*)pre_shift, V);
        if (in2) act_split_f16_kernel<<<(unsigned)((n2 / 4 + 255) / 256), 256, 0, s>>>((const float4*)in2, (uint2*)w2, (uint2*)w3, n2 / 4, X, Y, Z, C2 / 4, relu_in, s2d, tail);
        count_launch(in2 ? 2 : 1);
        xh = w0; xl = w1; hh = w2; hl = w3;
        inv_a = (const float*)tail + 1;
    } else if (split_pass) {
        float* w0 = (float*)ws; float* w1 = w0 + n1; float* w2 = w1 + n1; float* w3 = w2 + n2;
        act_split_kernel<<<(unsigned)((n1 / 4 + 255) / 256), 256, 0, s>>>((const float4*)in, (float4*)w0, (float4*)w1, n1 / 4, X, Y, Z, C / 4, V,
                                                                        relu_in, s2d, (const float4*)pre_scale, (const float4*)pre_shift);
        count_launch();
        if (in2) {
            act_split_kernel<<<(unsigned)((n2 / 4 + 255) / 256), 256, 0, s>>>((const float4*)in2, (float4*)w2, (float4*)w3, n2 / 4, X, Y, Z, C2 / 4, 1,
                                                                            relu_in, s2d, nullptr, nullptr);
            count_launch();
        }
        xh = w0; xl = w1; hh = w2; hl = w3;
    }
    // M space: the output lattice for a conv, the INPUT lattice for the transposed conv; a 1x1x1 conv has no neighbourhood,
    // so its voxels are flattened into one axis (full 128-row tiles whatever the grid shape)
    int MX = X, MY = Y, MZ = Z, nb_mul = 1;
    if (s2d) { MX = X / 2; MY = Y / 2; MZ = Z / 2; nb_mul = 8; }
    if (kind == MVF_CONV_S1 && ksize == 1) {
        const long long flat = (long long)X * Y * Z;
        if (flat > 2147483647ll) return MVF_EUNSUPPORTED;
        MX = 1; MY = 1; MZ = (int)flat;
    }
    TcArgs a;
    a.bias = bias; a.c_prev = nullptr; a.h_out = nullptr; a.c_out = nullptr;
    a.B = B; a.X = MX; a.Y = MY; a.Z = MZ; a.C = C; a.F = C2;
    a.BZ = pow2ceil(MZ) < TC_M ? pow2ceil(MZ) : TC_M;
    a.BY = pow2ceil(MY) < TC_M / a.BZ ? pow2ceil(MY) : TC_M / a.BZ;
    a.BX = TC_M / (a.BZ * a.BY);
    a.tiles_z = (MZ + a.BZ - 1) / a.BZ; a.tiles_y = (MY + a.BY - 1) / a.BY; a.tiles_x = (MX + a.BX - 1) / a.BX;
    a.has_h = in2 != nullptr; a.forget_bias = 0.f; a.halo_lo = 0; a.halo_hi = 0;
    a.fused_split = !split_pass; a.relu_x = relu_in; a.relu_h = relu_in; a.pre_scale = pre_scale; a.pre_shift = pre_shift;
    a.V = V; a.Cout = Cout; a.kind = kind; a.ksize = ksize; a.relu_out = (flags & MVF_FLAG_RELU_OUT) != 0;
    a.bn_scale = bn_scale; a.bn_shift = bn_shift; a.out = out;
    static const int promote_env = env_int("MVF_TC_PROMOTE", -1);
    a.promote = promote_env >= 0 ? promote_env : 8;
    const int K = conv_taps(kind, ksize) * (V * C + C2);
    const long long wtotal = (long long)K * Cout;
    const void *whi = wsplit, *wlo = wsplit + wtotal;
    a.inv_scale_a = inv_a; a.inv_scale_w = nullptr;
    if (f16) {
        const __half* w16 = f16_k1 ? (const __half*)(wsplit + 2 * wtotal) : (const __half*)wsplit;      // k = 1: behind the tf32 copy
        whi = w16; wlo = w16 + wtotal; a.inv_scale_w = (const float*)(w16 + 2 * wtotal) + 1;
    }
    CUtensorMap tm_xh, tm_xl, tm_hh, tm_hl, tm_wh, tm_wl;
    bool ok = make_act_map(&tm_xh, xh, B * V * nb_mul, MX, MY, MZ, C, a.BX, a.BY, a.BZ, f16) &&
              make_act_map(&tm_xl, xl, B * V * nb_mul, MX, MY, MZ, C, a.BX, a.BY, a.BZ, f16) &&
              make_w_map(&tm_wh, whi, K, Cout, f16) && make_w_map(&tm_wl, wlo, K, Cout, f16);
    if (ok && in2) ok = make_act_map(&tm_hh, hh, B * nb_mul, MX, MY, MZ, C2, a.BX, a.BY, a.BZ, f16) &&
                        make_act_map(&tm_hl, hl, B * nb_mul, MX, MY, MZ, C2, a.BX, a.BY, a.BZ, f16);
    if (!ok) return MVF_ECUDA;
    if (!in2) { tm_hh = tm_xh; tm_hl = tm_xl; }
    const long long mtiles = (long long)B * a.tiles_x * a.tiles_y * a.tiles_z;
    const int ntiles = (Cout + TC_N - 1) / TC_N;
    if (mtiles > 2147483647ll || ntiles > 65535) return MVF_EUNSUPPORTED;
    dim3 grid((unsigned)mtiles, ntiles, kind == MVF_DECONV_S2 ? 8 : 1);
    // per-device attribute: set on every call (a process may drive several GPUs)
    if (f16) {
        if (cudaFuncSetAttribute(convlstm_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess) return MVF_ECUDA;
        convlstm_tc_kernel<true, true><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(tm_xh, tm_xl, tm_hh, tm_hl, tm_wh, tm_wl, a);
    } else {
        if (cudaFuncSetAttribute(convlstm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess) return MVF_ECUDA;
        convlstm_tc_kernel<true, false><<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(tm_xh, tm_xl, tm_hh, tm_hl, tm_wh, tm_wl, a);
    }
    count_launch();
    return check_launch();
}

// grid_reas 'ident' (model_multi.py:443-455) = the 1x1x1 member of the family: ReLU -> conv over V*C channels -> bias -> BN -> ReLU
extern "C" size_t mvf_ident_wsplit_bytes(int V, int C, int Cout) { return mvf_conv3d_wsplit_bytes(MVF_CONV_S1, 1, V * C, Cout); }
extern "C" int mvf_ident_prepare(const float* weight, int V, int C, int Cout, float* wsplit, void* stream) {
    if (V <= 0 || C <= 0) return MVF_EINVAL;
    if (C % TC_K != 0) return MVF_EUNSUPPORTED;
    return mvf_conv3d_prepare(weight, MVF_CONV_S1, 1, V, C, 0, Cout, 0, wsplit, stream);
}
extern "C" size_t mvf_ident_tc_workspace_bytes(int B, int V, int X, int Y, int Z, int C, int Cout) {
    return mvf_conv3d_tc_workspace_bytes(MVF_CONV_S1, 1, B, V, X, Y, Z, C, 0, Cout);
}
extern "C" int mvf_ident_fuse_tc(const float* in, const float* wsplit, const float* bias,
                                 const float* bn_scale, const float* bn_shift,
                                 int B, int V, int X, int Y, int Z, int C, int Cout,
                                 float* out, void* ws, size_t ws_bytes, void* stream) {
    return mvf_conv3d_tc(in, nullptr, wsplit, bias, bn_scale, bn_shift, nullptr, nullptr, MVF_CONV_S1, 1, B, V, X, Y, Z, C, 0, Cout,
                         MVF_FLAG_RELU_IN | MVF_FLAG_RELU_OUT, out, ws, ws_bytes, nullptr, stream);
}
