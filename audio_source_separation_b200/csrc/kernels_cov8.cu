// FastMNMF update_diagonalizer, covariance step, for eight channels (src/bss/mnmf.py:853-875):
//     R[f,t,m] = max(sum_n (W_n H_n)[f,t] g[n,f,m], eps)          (:867-868)
//     V[m,f]   = (1/T) sum_t x_ft x_ft^H / R[f,t,m]               (:875), all M = 8 weight sets from ONE pass over the bin
// The 8 x 8 Hermitian outer product of a frame has 36 distinct entries = 4 pairs of diagonals + 28 complex off-diagonals
// = exactly 32 two-float slots: lane L of a warp owns slot L for ALL frames of the bin and keeps its 8 weighted sums
// (one per weight set m) in registers -- 8 paired FMAs per frame and lane, every product formed exactly once, no
// cross-lane reduction at the end.  The frame's samples and its 8 inverse variances are broadcast reads from shared
// memory; the inverse variances are computed in the kernel, one frame per lane, from (W, H, g) -- the weight tensor of
// the earlier two-kernel form (4 M F T bytes written and read back) does not exist.
// Arithmetic: fp32 products and partial sums over one 128-frame block, fp64 totals across the blocks.
#include <algorithm>
#include <cstdlib>

#include "handle.h"

namespace {

constexpr int C8 = 8;
constexpr int C8_WARPS = 16;      // at most; the launch picks the count that spreads the bins over all SMs
constexpr int C8_XROW = 9;        // float2 per frame record in shared memory: 8 channels + 1 pad (bank spread)

struct Cov8Params {
    const cf* X;          // [B][F] bin tiles, block-interleaved [blk][8][128]
    const float* basis;   // [B][N][F][K]
    const float* act;     // [B][N][K][Tp]
    const float* G;       // [B][N][F][8]
    double* U;            // [B][8][F][64] packed Hermitian
    int B, N, F, T, Tp, K;
    float eps;
    double inv_T;
    uint32_t warp_off, warp_stride;   // per-warp shared memory: frame records | inverse variances | g | basis row
};

// One CTA serves one mixture (blockIdx.y); its warps take the bins blockIdx.x * warps + warp, + gridDim.x * warps, ...
__global__ void __launch_bounds__(C8_WARPS * 32, 1) cov8_kernel(const Cov8Params p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int Tp = p.Tp, N = p.N, K = p.K;
    // activation rows of this mixture, shared by all bins: [N][K][Tp]
    float* hs = reinterpret_cast<float*>(smem);
    {
        const float4* src = reinterpret_cast<const float4*>(p.act + (size_t)b * N * K * Tp);
        float4* dst = reinterpret_cast<float4*>(hs);
        const int n4 = ((N * K * Tp) % 4 == 0) ? N * K * Tp / 4 : 0;   // 16-byte copies when every mixture's rows start aligned
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
        for (int i = n4 * 4 + threadIdx.x; i < N * K * Tp; i += blockDim.x) hs[i] = __ldg(p.act + (size_t)b * N * K * Tp + i);
    }
    __syncthreads();
    unsigned char* mine = smem + p.warp_off + (size_t)warp * p.warp_stride;
    float2* xrec = reinterpret_cast<float2*>(mine);                            // [32][C8_XROW]
    float* wrec = reinterpret_cast<float*>(mine + 32 * C8_XROW * 8);           // [32][8] inverse variances
    float* gs = wrec + 32 * 8;                                                 // [N][8]
    float* ts = gs + 8 * 8;                                                    // [N][K]

    // slot of this lane: lanes 0..3 hold the diagonal pairs (2L, 2L+1), lanes 4..31 the strictly lower entries (i > j), row major
    int ia, ib;
    const bool diag = lane < 4;
    if (diag) {
        ia = 2 * lane;
        ib = 2 * lane + 1;
    } else {
        int e = lane - 4, i = 1;
        while (e >= i) {
            e -= i;
            ++i;
        }
        ia = i;
        ib = e;
    }
    const int n_blocks = (Tp + BSS_XSLAB - 1) / BSS_XSLAB;

    const int n_warps = blockDim.x >> 5;
#pragma unroll 1
    for (int f = (int)blockIdx.x * n_warps + warp; f < p.F; f += (int)gridDim.x * n_warps) {
        const size_t bf = (size_t)b * p.F + f;
        for (int i = lane; i < N * 8; i += 32) gs[i] = __ldg(p.G + (((size_t)b * N + i / 8) * p.F + f) * 8 + (i & 7));
        for (int i = lane; i < N * K; i += 32) ts[i] = __ldg(p.basis + (((size_t)b * N + i / K) * p.F + f) * K + (i % K));
        __syncwarp();
        const cf* tile = p.X + bf * C8 * Tp;
        double tot[8][2];
#pragma unroll
        for (int m = 0; m < 8; ++m) tot[m][0] = tot[m][1] = 0.0;

#pragma unroll 1
        for (int blk = 0; blk < n_blocks; ++blk) {
            const int L = min(BSS_XSLAB, Tp - blk * BSS_XSLAB);     // frames of this block
            const cf* xblk = tile + (size_t)blk * BSS_XSLAB * C8;    // [8][L]
            float2 acc[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[m] = make_float2(0.f, 0.f);
            // the block in groups of 32 frames; the samples of the next group are fetched while this one is accumulated
            float2 xn[C8];
            {
                const bool live = lane < L;
#pragma unroll
                for (int c = 0; c < C8; ++c) xn[c] = live ? __ldg(xblk + (size_t)c * L + lane) : make_float2(0.f, 0.f);
            }
#pragma unroll 1
            for (int t0 = 0; t0 < L; t0 += 32) {
                // ---- stage the group: frame records [frame][channel], and the frame's 8 inverse variances ----------------
#pragma unroll
                for (int c = 0; c < C8; ++c) xrec[lane * C8_XROW + c] = xn[c];
                {
                    const int tn = t0 + 32 + lane;
                    const bool live = tn < L;
#pragma unroll
                    for (int c = 0; c < C8; ++c) xn[c] = live ? __ldg(xblk + (size_t)c * L + tn) : make_float2(0.f, 0.f);
                }
                {
                    const int t = blk * BSS_XSLAB + t0 + lane;     // this lane's frame of the group
                    float R[8];
#pragma unroll
                    for (int m = 0; m < 8; ++m) R[m] = 0.f;
                    if (t0 + lane < L) {
                        for (int n = 0; n < N; ++n) {
                            float lam = 0.f;
                            for (int k = 0; k < K; ++k) lam = fmaf(ts[n * K + k], hs[(n * K + k) * Tp + t], lam);
                            const float4 g0 = *reinterpret_cast<const float4*>(gs + n * 8), g1 = *reinterpret_cast<const float4*>(gs + n * 8 + 4);
                            R[0] = fmaf(lam, g0.x, R[0]);
                            R[1] = fmaf(lam, g0.y, R[1]);
                            R[2] = fmaf(lam, g0.z, R[2]);
                            R[3] = fmaf(lam, g0.w, R[3]);
                            R[4] = fmaf(lam, g1.x, R[4]);
                            R[5] = fmaf(lam, g1.y, R[5]);
                            R[6] = fmaf(lam, g1.z, R[6]);
                            R[7] = fmaf(lam, g1.w, R[7]);
                        }
#pragma unroll
                        for (int m = 0; m < 8; ++m) R[m] = rcp_fast(fmaxf(R[m], p.eps));
                    }
                    // frames past the end of the block carry zero samples and zero weights
                    *reinterpret_cast<float4*>(wrec + lane * 8) = make_float4(R[0], R[1], R[2], R[3]);
                    *reinterpret_cast<float4*>(wrec + lane * 8 + 4) = make_float4(R[4], R[5], R[6], R[7]);
                }
                __syncwarp();
                // ---- accumulate: one product per frame and lane, eight paired FMAs ----------------------------------------
#pragma unroll 4
                for (int tt = 0; tt < 32; ++tt) {
                    const float2 A = xrec[tt * C8_XROW + ia], Bv = xrec[tt * C8_XROW + ib];
                    const float4 w0 = *reinterpret_cast<const float4*>(wrec + tt * 8), w1 = *reinterpret_cast<const float4*>(wrec + tt * 8 + 4);
                    // off-diagonal: A conj(B) = (A.x B.x + A.y B.y, A.y B.x - A.x B.y);  diagonal pair: (|A|^2, |B|^2)
                    const float2 Pq = diag ? A : Bv;
                    const float2 Qq = diag ? Bv : make_float2(A.y, -A.x);
                    float2 v;
                    v.x = fmaf(A.x, Pq.x, A.y * Pq.y);
                    v.y = fmaf(Qq.x, Bv.x, Qq.y * Bv.y);
                    acc[0] = __ffma2_rn(v, make_float2(w0.x, w0.x), acc[0]);
                    acc[1] = __ffma2_rn(v, make_float2(w0.y, w0.y), acc[1]);
                    acc[2] = __ffma2_rn(v, make_float2(w0.z, w0.z), acc[2]);
                    acc[3] = __ffma2_rn(v, make_float2(w0.w, w0.w), acc[3]);
                    acc[4] = __ffma2_rn(v, make_float2(w1.x, w1.x), acc[4]);
                    acc[5] = __ffma2_rn(v, make_float2(w1.y, w1.y), acc[5]);
                    acc[6] = __ffma2_rn(v, make_float2(w1.z, w1.z), acc[6]);
                    acc[7] = __ffma2_rn(v, make_float2(w1.w, w1.w), acc[7]);
                }
                __syncwarp();
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                tot[m][0] += (double)acc[m].x;
                tot[m][1] += (double)acc[m].y;
            }
        }
        // packed Hermitian output: 8 diagonals, then the strictly lower entries row major as (re, im)
        const int o = diag ? 2 * lane : 8 + 2 * (lane - 4);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            double* dst = p.U + (((size_t)b * 8 + m) * p.F + f) * 64 + o;
            *reinterpret_cast<double2*>(dst) = make_double2(tot[m][0] * p.inv_T, tot[m][1] * p.inv_T);
        }
        __syncwarp();
    }
}

}  // namespace

// all 8 weighted covariances of every bin of a FastMNMF handle with 8 channels; *done = false when the shape is not covered
int launch_covariance8(bss_handle* h, bool* done) {
    *done = false;
    static const bool disabled = getenv("BSSGPU_NO_COV8") != nullptr;
    if (disabled || h->C != C8 || h->N > 8 || h->N < 1 || h->Tp < 2) return BSS_OK;
    const size_t act_bytes = (size_t)round_up((int)((size_t)h->N * h->K * h->Tp * sizeof(float)), 16);
    const size_t per_warp = (size_t)round_up(32 * C8_XROW * 8 + 32 * 8 * 4 + 8 * 8 * 4 + 8 * h->K * 4, 16);
    const size_t smem_bytes = act_bytes + C8_WARPS * per_warp;
    if (smem_bytes > (size_t)h->max_smem) return BSS_OK;
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(cov8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    Cov8Params p{};
    p.X = h->X;
    p.basis = h->basis;
    p.act = h->act;
    p.G = h->G;
    p.U = h->U;
    p.B = h->B;
    p.N = h->N;
    p.F = h->F;
    p.T = h->T;
    p.Tp = h->Tp;
    p.K = h->K;
    p.eps = (float)h->cfg.eps;
    p.inv_T = 1.0 / (double)h->T;
    p.warp_off = (uint32_t)act_bytes;
    p.warp_stride = (uint32_t)per_warp;
    // one resident CTA per SM over all mixtures; as many warps per CTA as it takes to give every warp one bin (at most 16)
    const int gx = std::max(1, std::min((int)cdiv(h->F, 4), h->n_sm / h->B > 0 ? h->n_sm / h->B : 1));
    int wpc = (int)cdiv(h->F, gx);
    if (wpc > C8_WARPS) wpc = C8_WARPS;
    if (wpc < 4) wpc = 4;
    dim3 grid((unsigned)gx, (unsigned)h->B);
    cov8_kernel<<<grid, wpc * 32, smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    *done = true;
    return BSS_OK;
}
