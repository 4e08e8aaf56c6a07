// Weighted covariances of ALL weight sets of a bin as one small tensor-core contraction (FastMNMF diagonaliser,
// src/bss/mnmf.py:867-875; up to 8 channels x 8 weight sets):
//     U_m[i][j] = (1/T) sum_t w_m(t) x_i(t) conj(x_j(t))     <=>     D[(i, j, re|im)][m] = sum_t P[(i, j, re|im)][t] Wt[t][m]
// i.e. a [128 x T] x [T x 8] product per bin whose left operand P (the 64 complex outer products of a frame) is
// shared by every weight set.  On CUDA cores each lane recomputes and re-accumulates P for every weight set
// (8 x 108 paired FMAs per frame pair); here P is computed once and the 8-fold accumulate runs on the tensor cores.
//
// Mapping onto mma.sync.m16n8k8 (TF32 inputs, FP32 accumulate), one warp per bin, K = 8 frames per step:
//   lane = (g, tig), g = lane / 4, tig = lane % 4.  The C (C + 1) / 2 entries (i >= j) of the Hermitian matrix are dealt
//   out eight per M-tile: tile s, lane group g owns entry e = 8 s + g, whose (Re, Im) are rows g and g + 8 of the tile:
//   A fragment {a0, a1, a2, a3} = {Re, Im of x_i conj(x_j) at frame tig, Re, Im at frame tig + 4};
//   B fragment {b0, b1} = weight set g at frames tig, tig + 4; the D fragment of tile s then carries
//   U_m[i][j] for m = 2 tig, 2 tig + 1.  Five tiles cover the 36 entries of an 8 x 8 matrix.  Every lane computes exactly
//   the products its fragment needs from the staged frame block: nothing passes through shared memory, which is why
//   this is the register-fragment mma.sync and not tcgen05.mma (whose operands must sit in shared memory / TMEM: the
//   products of every frame would have to be written out and read back, the very traffic this formulation avoids).
// Precision: 3xTF32 (hi/lo split of both operands, three MMAs per tile) keeps fp32-level accuracy; plain TF32 would
// put ~5e-4 relative error on U, which the per-bin solve amplifies by the condition number.
#include <cstdlib>

#include "handle.h"

namespace {

constexpr int CM_STAGES = 2;   // 12 KB stages at C = NW = 8: two stages x 8 warps fit, and 96 KB in flight per SM cover the HBM latency
constexpr int CM_WARPS = 8;

struct CovMmaParams {
    const cf* X;       // [B][F] tiles, rows = C
    const float* iw;   // [B][F] tiles, rows = NW (block-interleaved like X), inverse weights
    double* U;         // [B][NW][F][C*C] packed Hermitian
    int B, F, C, NW, T, Tp;
    int n_items, n_blocks;
    double inv_T;
    uint32_t stage_bytes, w_off;
};

// hi = v rounded to the nearest TF32 (10 mantissa bits; ties away from zero, exactly cvt.rna.tf32.f32 on finite values),
// lo = the exact remainder v - hi, left for the tensor core to truncate: v = hi + lo up to 2^-21 |v| with errors of either sign.  (Truncating
// instead leaves remainders that all carry the sign of v; the tensor core's own truncating adder then accumulates a
// systematic error that the per-bin solve of an 8 x 8 system amplifies past the 2e-4 parity bound -- measured.)  Integer
// add-and-mask instead of the cvt instruction, which ptxas expands into a ~5-instruction sequence: with 68 values to split
// per 8-frame step the conversions, not the MMAs, set the speed of this kernel.
__device__ __forceinline__ uint32_t round_tf32(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = round_tf32(v);
    lo = __float_as_uint(v - __uint_as_float(hi));   // the tensor core ignores the 13 low mantissa bits itself: 2^-21 |v|, either sign
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CT>
__global__ void __launch_bounds__(CM_WARPS * 32) cov_mma_kernel(const CovMmaParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tig = lane & 3;
    constexpr int C = CT;
    const int NW = p.NW, Tp = p.Tp;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * CM_STAGES;
    unsigned char* ring = smem + 256 + (size_t)warp * CM_STAGES * p.stage_bytes;
    const uint32_t bars_sa = smem_u32(bars), ring_sa = smem_u32(ring);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < CM_STAGES; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const int stride = (int)gridDim.x * CM_WARPS;
    const int first = (int)blockIdx.x * CM_WARPS + warp;
    const int nb = p.n_blocks;
    // producer cursor (bin, block) and consumer cursor
    int p_item = first, p_blk = 0, pstage = 0;
    auto issue = [&]() {
        if (p_item < p.n_items) {
            if (lane == 0) {
                const int t0 = p_blk * BSS_XSLAB;
                const int L = min(BSS_XSLAB, Tp - t0);
                const uint32_t xb = (uint32_t)(C * L * 8), wb = (uint32_t)(NW * L * 4);
                const uint32_t bar = bars_sa + 8u * (uint32_t)pstage;
                const uint32_t dst = ring_sa + (uint32_t)pstage * p.stage_bytes;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(xb + wb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                             "l"(p.X + (size_t)p_item * C * Tp + (size_t)t0 * C), "r"(xb), "r"(bar)
                             : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + p.w_off),
                             "l"(p.iw + (size_t)p_item * NW * Tp + (size_t)t0 * NW), "r"(wb), "r"(bar)
                             : "memory");
            }
            if (++p_blk == nb) {
                p_blk = 0;
                p_item += stride;
            }
            pstage = pstage + 1 == CM_STAGES ? 0 : pstage + 1;
        }
    };
#pragma unroll 1
    for (int i = 0; i < CM_STAGES - 1; ++i) issue();

    int cstage = 0;
    uint32_t cphase = 0;
    const bool w_live = g < NW;       // lane group's weight set exists
    // entries (i >= j) in row-major order: e = i (i + 1) / 2 + j; this lane group's entry of tile s is e = 8 s + g
    constexpr int n_entries = C * (C + 1) / 2;
    constexpr int n_tiles = (n_entries + 7) / 8;
    int ei[n_tiles], ej[n_tiles];
#pragma unroll
    for (int s2 = 0; s2 < n_tiles; ++s2) {
        const int e = 8 * s2 + g;
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        ei[s2] = e < n_entries ? i : -1;
        ej[s2] = e - i * (i + 1) / 2;
    }
#pragma unroll 1
    for (int item = first; item < p.n_items; item += stride) {
        // main term (hi x hi) and the two correction terms accumulate separately: shorter dependent MMA chains
        float d[n_tiles][4], dc[n_tiles][4];
#pragma unroll
        for (int j = 0; j < n_tiles; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) d[j][q] = dc[j][q] = 0.f;
#pragma unroll 1
        for (int blk = 0; blk < nb; ++blk) {
            issue();
            {
                const uint32_t bar = bars_sa + 8u * (uint32_t)cstage;
                uint32_t done = 0;
                while (!done) {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t}"
                        : "=r"(done)
                        : "r"(bar), "r"(cphase)
                        : "memory");
                }
            }
            const int t0 = blk * BSS_XSLAB;
            const int L = min(BSS_XSLAB, Tp - t0);
            const unsigned char* stage = ring + (size_t)cstage * p.stage_bytes;
            const cf* xs = reinterpret_cast<const cf*>(stage);
            const float* ws = reinterpret_cast<const float*>(stage + p.w_off);
#pragma unroll 1
            for (int k0 = 0; k0 < L; k0 += 8) {
                const int tA = k0 + tig, tB = tA + 4;
                const bool lA = tA < L, lB = tB < L;
                const cf zero = cf_make(0.f, 0.f);
                uint32_t bh[2], bl[2];
                split_tf32((w_live && lA) ? ws[g * L + tA] : 0.f, bh[0], bl[0]);
                split_tf32((w_live && lB) ? ws[g * L + tB] : 0.f, bh[1], bl[1]);
#pragma unroll
                for (int s2 = 0; s2 < n_tiles; ++s2) {
                    {
                        const bool live = ei[s2] >= 0;
                        const int i = live ? ei[s2] : 0, j = ej[s2];
                        const cf xiA = (live && lA) ? xs[i * L + tA] : zero;
                        const cf xiB = (live && lB) ? xs[i * L + tB] : zero;
                        const cf xjA = (live && lA) ? xs[j * L + tA] : zero;
                        const cf xjB = (live && lB) ? xs[j * L + tB] : zero;
                        // x_i conj(x_j)
                        const float a[4] = {fmaf(xiA.x, xjA.x, xiA.y * xjA.y), fmaf(xiA.y, xjA.x, -xiA.x * xjA.y),
                                            fmaf(xiB.x, xjB.x, xiB.y * xjB.y), fmaf(xiB.y, xjB.x, -xiB.x * xjB.y)};
                        uint32_t ah[4], al[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) split_tf32(a[q], ah[q], al[q]);
                        mma_tf32(dc[s2], al, bh);
                        mma_tf32(dc[s2], ah, bl);
                        mma_tf32(d[s2], ah, bh);
                    }
                }
            }
            __syncwarp();
            if (++cstage == CM_STAGES) {
                cstage = 0;
                cphase ^= 1u;
            }
        }
        // D fragment of tile s: {Re, Re, Im, Im} of U_m[i][j] for m = 2 tig, 2 tig + 1
        const int b = item / p.F, f = item - b * p.F;
        const int CC = C * C;
#pragma unroll
        for (int s2 = 0; s2 < n_tiles; ++s2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) d[s2][q] += dc[s2][q];
            if (ei[s2] >= 0) {
                const int i = ei[s2], j = ej[s2];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int m = 2 * tig + h2;
                    if (m < NW) {
                        double* u = p.U + (((size_t)b * NW + m) * p.F + f) * CC;
                        if (i == j) {
                            u[i] = (double)d[s2][h2] * p.inv_T;
                        } else {
                            const int e = C + 2 * (i * (i - 1) / 2 + j);
                            u[e] = (double)d[s2][h2] * p.inv_T;
                            u[e + 1] = (double)d[s2][2 + h2] * p.inv_T;
                        }
                    }
                }
            }
        }
    }
}

template <int CT>
int launch_cov_mma_t(bss_handle* h, const CovMmaParams& p, size_t smem, long long n_items) {
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(cov_mma_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    int ctas = 1;
    BSS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, cov_mma_kernel<CT>, CM_WARPS * 32, smem));
    if (ctas < 1) return bss_fail(h, BSS_ECUDA, "tensor-core covariance kernel does not fit");
    long long grid = cdiv(n_items, CM_WARPS);
    if (grid > (long long)h->n_sm * ctas) grid = (long long)h->n_sm * ctas;
    cov_mma_kernel<CT><<<(unsigned)grid, CM_WARPS * 32, smem, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

}  // namespace

// iw must be in the block-interleaved tile layout (rows = NW); returns *done = false when the shape is not covered
int launch_covariance_mma(bss_handle* h, const cf* X, const float* iw_tiled, double* U, int B, int F, int C, int NW, int T, int Tp,
                          bool* done) {
    *done = false;
    if (C > 8 || NW > 8 || getenv("BSSGPU_NO_MMA")) return BSS_OK;
    const long long n_items = (long long)B * F;
    if (n_items > 0x7fffffffLL || n_items == 0) return BSS_OK;
    CovMmaParams p{};
    p.X = X;
    p.iw = iw_tiled;
    p.U = U;
    p.B = B;
    p.F = F;
    p.C = C;
    p.NW = NW;
    p.T = T;
    p.Tp = Tp;
    p.n_items = (int)n_items;
    p.n_blocks = (Tp + BSS_XSLAB - 1) / BSS_XSLAB;
    p.inv_T = 1.0 / (double)T;
    const int L = Tp < BSS_XSLAB ? Tp : BSS_XSLAB;
    p.w_off = (uint32_t)round_up(C * L * 8, 16);
    p.stage_bytes = (uint32_t)round_up((int)p.w_off + NW * L * 4, 128);
    const size_t smem = 256 + (size_t)CM_WARPS * CM_STAGES * p.stage_bytes;
    if (smem > (size_t)h->max_smem) return BSS_OK;
    // bulk copies need 16-byte sizes: C L 8 always is (L even); NW L 4 needs NW L % 4 == 0
    for (int blk = 0; blk < p.n_blocks; ++blk) {
        const int Lb = (Tp - blk * BSS_XSLAB) < BSS_XSLAB ? (Tp - blk * BSS_XSLAB) : BSS_XSLAB;
        if ((NW * Lb * 4) % 16 != 0 || ((size_t)blk * BSS_XSLAB * NW * 4) % 16 != 0) return BSS_OK;
    }
    if (((size_t)NW * Tp * 4) % 16 != 0) return BSS_OK;
    int rc = BSS_OK;
    switch (C) {
        case 2: rc = launch_cov_mma_t<2>(h, p, smem, n_items); break;
        case 3: rc = launch_cov_mma_t<3>(h, p, smem, n_items); break;
        case 4: rc = launch_cov_mma_t<4>(h, p, smem, n_items); break;
        case 5: rc = launch_cov_mma_t<5>(h, p, smem, n_items); break;
        case 6: rc = launch_cov_mma_t<6>(h, p, smem, n_items); break;
        case 7: rc = launch_cov_mma_t<7>(h, p, smem, n_items); break;
        default: rc = launch_cov_mma_t<8>(h, p, smem, n_items); break;
    }
    if (rc != BSS_OK) return rc;
    *done = true;
    return BSS_OK;
}
