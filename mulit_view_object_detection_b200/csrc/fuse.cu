// View fusion on materialised tensors: grid_reas 'add'/mean/max reduction, 'ident' 1x1x1 conv,
// and the ConvLSTM cell step (fp32 CUDA-core implicit GEMM; exact-fp32 baseline path).
//
// Replaces mrcnn/model_multi.py:394-463 and mrcnn/recurrent.py:442-479.
#include "mvf_common.cuh"

namespace mvf {

// ---------------------------------------------------------------------------------------------
// [B,V,N,C] -> [B,N,C]   (model_multi.py:401-404; mean/max per SURVEY.md spec B)
template <int MODE>
__global__ void __launch_bounds__(256)
view_reduce_kernel(const float4* __restrict__ in, float4* __restrict__ out, const float4* __restrict__ bn_scale,
                   const float4* __restrict__ bn_shift, int V, long long per /* N*C4 */, int C4, int flags, float inv_v) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (e >= per) return;
    const float4* src = in + (long long)b * V * per + e;
    float4 acc = __ldcs(src);
    if (flags & MVF_FLAG_RELU_IN) acc = relu4(acc);
    for (int v = 1; v < V; ++v) {
        float4 x = __ldcs(src + (long long)v * per);
        if (flags & MVF_FLAG_RELU_IN) x = relu4(x);
        acc = (MODE == MVF_FUSE_MAX) ? max4(acc, x) : add4(acc, x);
    }
    if (MODE == MVF_FUSE_MEAN) acc = mul4(inv_v, acc);
    if (bn_scale) {
        const int c4 = (int)(e % C4);
        const float4 s = __ldg(bn_scale + c4), h = __ldg(bn_shift + c4);
        acc = make_float4(fmaf(acc.x, s.x, h.x), fmaf(acc.y, s.y, h.y), fmaf(acc.z, s.z, h.z), fmaf(acc.w, s.w, h.w));
    }
    if (flags & MVF_FLAG_RELU_OUT) acc = relu4(acc);
    __stcs(out + (long long)b * per + e, acc);
}

// ---------------------------------------------------------------------------------------------
// 'ident': out[b,n,co] = relu(bn(sum_{v,c} relu(in[b,v,n,c]) * W[v*C+c, co] + bias[co]))
// 64 voxels x 64 couts per CTA, K chunks of 16, 4x4 register tile per thread.
constexpr int GT_M = 64, GT_N = 64, GT_K = 16;

__global__ void __launch_bounds__(256)
ident_fuse_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                  const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, float* __restrict__ out,
                  int V, long long N, int C, int Cout) {
    __shared__ float sA[GT_K][GT_M + 4];
    __shared__ float sB[GT_K][GT_N + 4];
    const int b = blockIdx.z;
    const long long n0 = (long long)blockIdx.x * GT_M;
    const int co0 = blockIdx.y * GT_N;
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;
    float acc[4][4] = {};
    const int KT = V * C;
    for (int k0 = 0; k0 < KT; k0 += GT_K) {
        {   // A: 64 voxels x 16 k  -> one float4 (4 consecutive k) per thread
            const int m = tid / 4, kq = (tid % 4) * 4;
            const int k = k0 + kq;
            float4 a = zero4();
            if (n0 + m < N && k < KT) {
                const int v = k / C, c = k % C;        // C % 4 == 0: the float4 stays inside one view
                a = relu4(ldg4(in + (((long long)b * V + v) * N + n0 + m) * C + c));
            }
            sA[kq + 0][m] = a.x; sA[kq + 1][m] = a.y; sA[kq + 2][m] = a.z; sA[kq + 3][m] = a.w;
        }
        {   // B: 16 k x 64 couts
            const int kk = tid / 16, cq = (tid % 16) * 4;
            float4 w = zero4();
            if (k0 + kk < KT && co0 + cq < Cout) w = ldg4(W + (long long)(k0 + kk) * Cout + co0 + cq);
            sB[kk][cq + 0] = w.x; sB[kk][cq + 1] = w.y; sB[kk][cq + 2] = w.z; sB[kk][cq + 3] = w.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GT_K; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = sB[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tx * 4 + j;
            if (co >= Cout) continue;
            float y = acc[i][j] + bias[co];
            if (bn_scale) y = fmaf(y, bn_scale[co], bn_shift[co]);
            out[((long long)b * N + n) * Cout + co] = fmaxf(y, 0.f);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// ConvLSTM cell step, fp32 implicit GEMM on CUDA cores.
//   y = conv3d_SAME([x ; h_prev], W) + b ; j,i,f,o = split(y) ; c = c_prev*sig(f+fb) + sig(i)*tanh(j) ;
//   h = tanh(c)*sig(o)                                   (mrcnn/recurrent.py:453-477)
// CTA tile: 64 voxels x (4 gates x 16 filters); a thread owns 4 voxels x 4 gates of one filter
// so the gate non-linearities run on registers in the epilogue.
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(256)
convlstm_step_kernel(const float* __restrict__ x, const float* __restrict__ h_prev, const float* __restrict__ c_prev,
                     const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ h_out,
                     float* __restrict__ c_out, int B, int X, int Y, int Z, int C, int F, float forget_bias, int relu_in) {
    __shared__ float sA[GT_K][GT_M + 4];
    __shared__ float sB[GT_K][GT_N + 4];
    const long long M = (long long)B * X * Y * Z;
    const long long m0 = (long long)blockIdx.x * GT_M;
    const int f0 = blockIdx.y * 16;
    const int tid = threadIdx.x;
    const int ty = tid / 16, tx = tid % 16;
    const int CF = C + F, G = 4 * F;
    float acc[4][4] = {};          // [voxel][gate]

    // this thread's A-load voxel
    const int am = tid / 4, akq = (tid % 4) * 4;
    const long long avox = m0 + am;
    int ab = 0, ax = 0, ay = 0, az = 0;
    if (avox < M) {
        long long r = avox;
        az = (int)(r % Z); r /= Z; ay = (int)(r % Y); r /= Y; ax = (int)(r % X); ab = (int)(r / X);
    }
    const int bk = tid / 16, bq = tid % 16;        // B-load: k row, float4 slot (gate = bq/4, 4 filters)

    for (int tap = 0; tap < 27; ++tap) {
        const int dx = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dz = tap % 3 - 1;
        const int nx = ax + dx, ny = ay + dy, nz = az + dz;
        const bool nb_ok = (avox < M) && nx >= 0 && nx < X && ny >= 0 && ny < Y && nz >= 0 && nz < Z;
        const long long nvox = (((long long)ab * X + nx) * Y + ny) * Z + nz;
        for (int k0 = 0; k0 < CF; k0 += GT_K) {
            {
                const int ch = k0 + akq;
                float4 a = zero4();
                if (nb_ok && ch < CF) {
                    if (ch < C) { a = ldg4(x + nvox * C + ch); if (relu_in) a = relu4(a); }
                    else if (h_prev) a = ldg4(h_prev + nvox * F + (ch - C));
                }
                sA[akq + 0][am] = a.x; sA[akq + 1][am] = a.y; sA[akq + 2][am] = a.z; sA[akq + 3][am] = a.w;
            }
            {
                const int ch = k0 + bk;
                const int gate = bq / 4, fq = (bq % 4) * 4;
                float4 w = zero4();
                if (ch < CF && f0 + fq < F) w = ldg4(W + ((long long)tap * CF + ch) * G + gate * F + f0 + fq);
                const int col = gate * 16 + fq;
                sB[bk][col + 0] = w.x; sB[bk][col + 1] = w.y; sB[bk][col + 2] = w.z; sB[bk][col + 3] = w.w;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < GT_K; ++kk) {
                float a[4], w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
                for (int g = 0; g < 4; ++g) w[g] = sB[kk][g * 16 + tx];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int g = 0; g < 4; ++g) acc[i][g] = fmaf(a[i], w[g], acc[i][g]);
            }
            __syncthreads();
        }
    }
    const int f = f0 + tx;
    if (f >= F) return;
    const float bj = bias[0 * F + f], bi = bias[1 * F + f], bf = bias[2 * F + f], bo = bias[3 * F + f];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long vox = m0 + ty * 4 + i;
        if (vox >= M) continue;
        const float gj = acc[i][0] + bj, gi = acc[i][1] + bi, gf = acc[i][2] + bf, go = acc[i][3] + bo;   // :460 order j,i,f,o
        const float cp = c_prev ? c_prev[vox * F + f] : 0.f;
        const float c = cp * sigmoidf_acc(gf + forget_bias) + sigmoidf_acc(gi) * tanhf(gj);           // :470-472
        const float h = tanhf(c) * sigmoidf_acc(go);                                                  // :477
        c_out[vox * F + f] = c;
        h_out[vox * F + f] = h;
    }
}

// Notebook GRID_REAS='mean' (Notebook/projection.py:526-529,549): mean over the CHANNEL axis, the V per-view scalars become the
// channels, ReLU.  One warp per (b, v, n) channel vector: lanes stride over float4s, xor-shuffle tree.
__global__ void __launch_bounds__(256)
channel_mean_kernel(const float4* __restrict__ in, float* __restrict__ out, int V, long long N, int C4, long long rows) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);          // row = (b * V + v) * N + n
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int c = lane; c < C4; c += 32) {
        const float4 v = __ldg(in + row * C4 + c);
        s += (v.x + v.y) + (v.z + v.w);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) {
        const long long n = row % N, bv = row / N;
        const long long b = bv / V, v = bv % V;
        out[(b * N + n) * V + v] = fmaxf(s / (float)(4 * C4), 0.f);
    }
}

}  // namespace mvf

using namespace mvf;

extern "C" int mvf_channel_mean(const float* in, int B, int V, long long N, int C, float* out, void* stream) {
    if (!in || !out) return MVF_ENULL;
    if (B <= 0 || V <= 0 || N <= 0 || C <= 0) return MVF_EINVAL;
    if (C % 4 != 0 || !aligned16(in)) return MVF_EALIGN;
    const long long rows = (long long)B * V * N;
    if ((rows + 7) / 8 > 2147483647ll) return MVF_EUNSUPPORTED;
    channel_mean_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const float4*)in, out, V, N, C / 4, rows);
    count_launch();
    return check_launch();
}

extern "C" int mvf_view_reduce(const float* in, int B, int V, long long N, int C, int mode, int flags,
                               const float* bn_scale, const float* bn_shift, float* out, void* stream) {
    if (!in || !out) return MVF_ENULL;
    if (B <= 0 || V <= 0 || N <= 0 || C <= 0) return MVF_EINVAL;
    if (mode < MVF_FUSE_SUM || mode > MVF_FUSE_MAX) return MVF_EINVAL;
    if ((bn_scale == nullptr) != (bn_shift == nullptr)) return MVF_ENULL;
    if (C % 4 != 0 || !aligned16(in) || !aligned16(out) || (bn_scale && (!aligned16(bn_scale) || !aligned16(bn_shift)))) return MVF_EALIGN;
    if (B > 65535) return MVF_EUNSUPPORTED;
    const long long per = N * (C / 4);
    dim3 grid((unsigned)((per + 255) / 256), B);
    cudaStream_t s = (cudaStream_t)stream;
    const float inv_v = 1.0f / (float)V;
    const float4 *i4 = (const float4*)in, *sc = (const float4*)bn_scale, *sh = (const float4*)bn_shift;
    float4* o4 = (float4*)out;
    if (mode == MVF_FUSE_SUM) view_reduce_kernel<MVF_FUSE_SUM><<<grid, 256, 0, s>>>(i4, o4, sc, sh, V, per, C / 4, flags, inv_v);
    else if (mode == MVF_FUSE_MEAN) view_reduce_kernel<MVF_FUSE_MEAN><<<grid, 256, 0, s>>>(i4, o4, sc, sh, V, per, C / 4, flags, inv_v);
    else view_reduce_kernel<MVF_FUSE_MAX><<<grid, 256, 0, s>>>(i4, o4, sc, sh, V, per, C / 4, flags, inv_v);
    count_launch();
    return check_launch();
}

extern "C" int mvf_ident_fuse(const float* in, const float* weight, const float* bias,
                              const float* bn_scale, const float* bn_shift,
                              int B, int V, long long N, int C, int Cout, float* out, void* stream) {
    if (!in || !weight || !bias || !out) return MVF_ENULL;
    if (B <= 0 || V <= 0 || N <= 0 || C <= 0 || Cout <= 0) return MVF_EINVAL;
    if ((bn_scale == nullptr) != (bn_shift == nullptr)) return MVF_ENULL;
    if (C % 4 != 0 || Cout % 4 != 0 || !aligned16(in) || !aligned16(weight) || !aligned16(out)) return MVF_EALIGN;
    if (B > 65535) return MVF_EUNSUPPORTED;
    dim3 grid((unsigned)((N + GT_M - 1) / GT_M), (Cout + GT_N - 1) / GT_N, B);
    ident_fuse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, weight, bias, bn_scale, bn_shift, out, V, N, C, Cout);
    count_launch();
    return check_launch();
}

extern "C" int mvf_convlstm_step(const float* x, const float* h_prev, const float* c_prev,
                                 const float* W, const float* bias, float forget_bias,
                                 int B, int X, int Y, int Z, int C, int F, int flags,
                                 float* h_out, float* c_out, void* stream) {
    if (!x || !W || !bias || !h_out || !c_out) return MVF_ENULL;
    if (B <= 0 || X <= 0 || Y <= 0 || Z <= 0 || C <= 0 || F <= 0) return MVF_EINVAL;
    if (h_out == h_prev || c_out == h_prev) return MVF_EINVAL;
    if (C % 4 != 0 || F % 4 != 0 || !aligned16(x) || !aligned16(W) || (h_prev && !aligned16(h_prev))) return MVF_EALIGN;
    const long long M = (long long)B * X * Y * Z;
    dim3 grid((unsigned)((M + GT_M - 1) / GT_M), (F + 15) / 16);
    convlstm_step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, h_prev, c_prev, W, bias, h_out, c_out, B, X, Y, Z, C, F,
                                                                  forget_bias, (flags & MVF_FLAG_RELU_IN) != 0);
    count_launch();
    return check_launch();
}
