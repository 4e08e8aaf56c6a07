// Source-model update of Gauss-ILRMA (domain 2, n_basis == 2) in ONE pass over the mixture:
//   basis T      src/bss/ilrma.py:413-419   T <- T sqrt( sum_t (P / TV^2) V / max(sum_t (1 / TV) V, eps) )
//   activation V src/bss/ilrma.py:422-428   V <- V sqrt( sum_f (P / T'V^2) T' / max(sum_f (1 / T'V) T', eps) ),  T' = the new basis
// with P = |W x|^2.  The three-pass form (kernels_mu.cu) streams X for the basis update, writes P (half the bytes of X),
// and streams P again for the activation update, because there a warp owns a bin in the first kernel and a frame block in
// the second.  Here a CTA owns a run of bins of one mixture and each of its warps owns one 128-frame block of EVERY bin
// of the run: the powers of a lane's frames stay in registers between the two updates, the warps of the CTA exchange
// only their 16 partial sums of the basis statistics per bin (one CTA barrier), every warp forms the new basis row
// and adds the bin's contribution to its activation accumulators.  P never exists in memory: an iteration moves
// 2 x 8 C F T bytes (this kernel and the covariance kernel) instead of 3 x.
// The sums over frames are taken lane-butterfly first, then over the warps in block order; the sums over bins chunk by
// chunk in mu_act_finish_kernel (kernels_mu.cu): deterministic, and independent of the batch size for a fixed chunk count.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "handle.h"

namespace {

constexpr int FU_STAGES = 4;      // ring stages per warp (one 128-frame block of one bin + its packed parameters each); 3 when four CTAs share an SM
constexpr int FU_MAX_WARPS = 4;   // frame blocks per bin tile covered by one CTA (Tp <= 512)
// shared-memory head: [FU_MAX_WARPS][FU_STAGES] ring mbarriers, then exchange[2][FU_MAX_WARPS][16] floats; the stages follow
constexpr int FU_HEAD_BYTES = (FU_MAX_WARPS * FU_STAGES * 8 + 2 * FU_MAX_WARPS * 16 * 4 + 127) / 128 * 128;

struct FusedParams {
    MuArgs a;
    float* part;                  // [B][n_chunks][N][K][2][Tp] partial sums of the activation statistics
    const unsigned char* pbin;    // packed per-bin parameters: Wf [C][C] complex64 | T [N][K] float, pb_stride bytes per bin
    int pb_stride;
    int n_chunks, bins_per_chunk, n_blocks;
    uint32_t stage_bytes, par_off;
};

// Packed per-bin parameters of the fused kernel: the bin's filter rows Wf [C][C] complex64 followed by its basis values
// T [N][K] (160 bytes at C = 4, K = 2), so that both arrive in the ring stage with the bin's block by one bulk copy.
// One thread per 16-byte piece: 8 pieces of filter (straight copies), then N K / 4 pieces gathered from [B][N][F][K].
__global__ void __launch_bounds__(256) fused_pack_kernel(const cf* Wf, const float* basis, unsigned char* out, int B, int N, int C, int F,
                                                         int K, int stride) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int pieces = stride >> 4;
    if (idx >= (long long)B * F * pieces) return;
    const int q = (int)(idx % pieces);
    const long long bf = idx / pieces;
    const int f = (int)(bf % F), b = (int)(bf / F);
    const int wq = C * C / 2;   // 16-byte pieces of the filter
    float4 v;
    if (q < wq) {
        v = __ldg(reinterpret_cast<const float4*>(Wf + (size_t)bf * C * C) + q);
    } else {
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = (q - wq) * 4 + j, n = i / K, k = i - n * K;
            t[j] = i < N * K ? __ldg(basis + (((size_t)b * N + n) * F + f) * K + k) : 0.f;
        }
        v = make_float4(t[0], t[1], t[2], t[3]);
    }
    reinterpret_cast<float4*>(out)[idx] = v;
}

__device__ __forceinline__ float2 rcp2f(float2 v) { return make_float2(rcp_fast(v.x), rcp_fast(v.y)); }

// Butterfly reduce-scatter of 16 per-lane values over the 32 lanes: afterwards lane L holds in the return value the
// warp-wide sum of element (L >> 1) (lanes 2e and 2e+1 both hold element e).  16 shuffles.
__device__ __forceinline__ float reduce16(float (&v)[16], int lane) {
#pragma unroll
    for (int lvl = 0; lvl < 4; ++lvl) {
        const int off = 16 >> lvl;          // 16, 8, 4, 2
        const int cnt = 8 >> lvl;           // 8, 4, 2, 1
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
            const float lo = v[i], hi = v[i + cnt];
            const float send = up ? lo : hi;
            const float keep = up ? hi : lo;
            v[i] = keep + __shfl_xor_sync(BSS_FULL, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(BSS_FULL, v[0], 1);
}

// One CTA = blockDim.x / 32 = n_blocks warps; warp g owns frame block g of the bins [f_begin, f_end) of mixture b.
// MINB: resident CTAs per SM the register allocation aims at (2: 204 registers, 3: 168, 4: 128 with a few spills)
template <int C, int KC, int MINB>
__global__ void __launch_bounds__(FU_MAX_WARPS * 32, MINB) mu_fused_kernel(const FusedParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    static_assert(C * KC * 2 == 16, "the partial-sum exchange is laid out for 16 statistics per bin (C = 4, K = 2)");
    constexpr int N = C;
    constexpr int STG = MINB >= 4 ? FU_STAGES - 1 : FU_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;
    const MuArgs& a = p.a;
    const int chunk = (int)blockIdx.x % p.n_chunks;
    const int b = (int)blockIdx.x / p.n_chunks;
    const int f_begin = chunk * p.bins_per_chunk;
    const int f_end = min(a.F, f_begin + p.bins_per_chunk);
    const int Tp = a.Tp;
    const int blk0 = warp * BSS_XSLAB;
    const int L = min(BSS_XSLAB, Tp - blk0);          // frames of this warp's block (even, > 0)
    const uint32_t blk_bytes = (uint32_t)(C * L * 8);
    const size_t bin_bytes = (size_t)C * Tp * 8;
    const unsigned char* src0 = reinterpret_cast<const unsigned char*>(a.X) + ((size_t)b * a.F * C * Tp + (size_t)blk0 * C) * 8;
    const unsigned char* par0 = p.pbin + (size_t)b * a.F * p.pb_stride;

    // shared memory: [warps][STG] mbarriers | exchange[2][FU_MAX_WARPS][16] floats | [warps][STG] stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * FU_STAGES;
    float* exch = reinterpret_cast<float*>(smem + FU_MAX_WARPS * FU_STAGES * 8);
    unsigned char* ring = smem + FU_HEAD_BYTES + (size_t)warp * STG * p.stage_bytes;
    const uint32_t bars_sa = smem_u32(bars), ring_sa = smem_u32(ring);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STG; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncwarp();
    auto issue = [&](int f, int stage) {
        if (lane == 0) {
            const uint32_t bar = bars_sa + 8u * (uint32_t)stage;
            const uint32_t dst = ring_sa + (uint32_t)stage * p.stage_bytes;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(blk_bytes + (uint32_t)p.pb_stride) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src0 + (size_t)f * bin_bytes), "r"(blk_bytes), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + p.par_off),
                         "l"(par0 + (size_t)f * p.pb_stride), "r"((uint32_t)p.pb_stride), "r"(bar)
                         : "memory");
        }
    };
    int fp = f_begin, pstage = 0;
#pragma unroll 1
    for (int i = 0; i < STG - 1 && fp < f_end; ++i, ++fp) {
        issue(fp, pstage);
        pstage = pstage + 1 == STG ? 0 : pstage + 1;
    }

    // loop invariants of this lane: activation values of its two frame pairs, and the activation statistics it accumulates
    float2 vreg[2][N][KC];
    float2 vnum[2][N][KC], vden[2][N][KC];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int tt = 2 * lane + 64 * j;
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                vreg[j][n][kk] = tt < L ? __ldg(reinterpret_cast<const float2*>(a.act + (((size_t)b * N + n) * KC + kk) * Tp + blk0 + tt))
                                        : make_float2(0.f, 0.f);
                vnum[j][n][kk] = vden[j][n][kk] = make_float2(0.f, 0.f);
            }
    }

    int tts[2];   // first frame of this lane's pairs inside the block, clamped for lanes past the end of a ragged block
#pragma unroll
    for (int j = 0; j < 2; ++j) tts[j] = (2 * lane + 64 * j) < L ? 2 * lane + 64 * j : 0;

    int cstage = 0;
    uint32_t cphase = 0;
#pragma unroll 1
    for (int f = f_begin; f < f_end; ++f) {
        if (fp < f_end) {
            issue(fp, pstage);
            ++fp;
            pstage = pstage + 1 == STG ? 0 : pstage + 1;
        }
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
        const unsigned char* stage = ring + (size_t)cstage * p.stage_bytes;
        const cf* xs = reinterpret_cast<const cf*>(stage);
        const float2* wf = reinterpret_cast<const float2*>(stage + p.par_off);
        const float* tb = reinterpret_cast<const float*>(stage + p.par_off) + C * C * 2;
        float tk[N][KC];
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) tk[n][kk] = tb[n * KC + kk];

        // ---- source powers of this lane's frames, basis statistics with the old basis ---------------------------------
        // (no per-lane branches: lanes past the end of a ragged block read frame 0 instead; their activation values are
        // zero, so they add nothing to the basis statistics, and their activation statistics are never stored)
        float2 P[2][N];
        float2 tnum[N][KC], tden[N][KC];
        float4 xv[2][C];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c) xv[j][c] = *reinterpret_cast<const float4*>(xs + c * L + tts[j]);
#pragma unroll
        for (int n = 0; n < N; ++n) {
            float2 y[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j) y[j][0] = y[j][1] = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                // w x = x * w.x + (-x.y, x.x) * w.y   (the operation order of frame_power2, kernels_mu.cu)
                const float2 w = wf[n * C + c];
                const float2 wx = make_float2(w.x, w.x), wy = make_float2(w.y, w.y);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float2 x0 = make_float2(xv[j][c].x, xv[j][c].y), x1 = make_float2(xv[j][c].z, xv[j][c].w);
                    y[j][0] = __ffma2_rn(x0, wx, y[j][0]);
                    y[j][0] = __ffma2_rn(make_float2(-x0.y, x0.x), wy, y[j][0]);
                    y[j][1] = __ffma2_rn(x1, wx, y[j][1]);
                    y[j][1] = __ffma2_rn(make_float2(-x1.y, x1.x), wy, y[j][1]);
                }
            }
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) tnum[n][kk] = tden[n][kk] = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 s0 = __fmul2_rn(y[j][0], y[j][0]), s1 = __fmul2_rn(y[j][1], y[j][1]);
                P[j][n] = make_float2(s0.x + s0.y, s1.x + s1.y);
                float2 tv = make_float2(0.f, 0.f);
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) tv = __ffma2_rn(vreg[j][n][kk], make_float2(tk[n][kk], tk[n][kk]), tv);
                tv.x = fmaxf(tv.x, a.eps);
                tv.y = fmaxf(tv.y, a.eps);
                const float2 sb = rcp2f(tv);
                const float2 sa = __fmul2_rn(P[j][n], __fmul2_rn(sb, sb));   // P / TV^2
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    tnum[n][kk] = __ffma2_rn(sa, vreg[j][n][kk], tnum[n][kk]);
                    tden[n][kk] = __ffma2_rn(sb, vreg[j][n][kk], tden[n][kk]);
                }
            }
        }
        __syncwarp();

        // ---- 16 statistics: over the lanes (butterfly), then over the warps in block order ---------------------------
        float flat[16];
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                flat[(n * KC + kk) * 2] = tnum[n][kk].x + tnum[n][kk].y;
                flat[(n * KC + kk) * 2 + 1] = tden[n][kk].x + tden[n][kk].y;
            }
        const float mine = reduce16(flat, lane);      // lanes 2e, 2e+1: statistic e = (n, k, num | den)
        float* ex = exch + ((f & 1) * FU_MAX_WARPS + warp) * 16;
        if ((lane & 1) == 0) ex[lane >> 1] = mine;
        __syncthreads();
        // lane l < 8 forms the new basis value of (n, k) = l; every warp does (same inputs, same order: identical values)
        float tnew = 0.f;
        if (lane < N * KC) {
            const float* e0 = exch + (f & 1) * FU_MAX_WARPS * 16;
            float nm = 0.f, dn = 0.f;
            for (int g = 0; g < n_warps; ++g) {
                nm += e0[g * 16 + 2 * lane];
                dn += e0[g * 16 + 2 * lane + 1];
            }
            dn = fmaxf(dn, a.eps);
            const float told = tb[lane];
            tnew = told * sqrtf(nm / dn);
            if (warp == 0) a.basis_out[(((size_t)b * N + lane / KC) * a.F + f) * KC + (lane % KC)] = tnew;
        }
        __syncwarp();   // the last read of the stage (tb) is behind every lane: lane 0 may refill it at the top of the next bin
        float tn[N][KC];
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) tn[n][kk] = __shfl_sync(BSS_FULL, tnew, n * KC + kk);

        // ---- activation statistics with the new basis --------------------------------------------------------------------
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int n = 0; n < N; ++n) {
                float2 tv = make_float2(0.f, 0.f);
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) tv = __ffma2_rn(vreg[j][n][kk], make_float2(tn[n][kk], tn[n][kk]), tv);
                tv.x = fmaxf(tv.x, a.eps);
                tv.y = fmaxf(tv.y, a.eps);
                const float2 sb = rcp2f(tv);
                const float2 sa = __fmul2_rn(P[j][n], __fmul2_rn(sb, sb));
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    const float2 t2 = make_float2(tn[n][kk], tn[n][kk]);
                    vnum[j][n][kk] = __ffma2_rn(sa, t2, vnum[j][n][kk]);
                    vden[j][n][kk] = __ffma2_rn(sb, t2, vden[j][n][kk]);
                }
            }
        }
        if (++cstage == STG) {
            cstage = 0;
            cphase ^= 1u;
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int tt = 2 * lane + 64 * j;
        if (tt < L) {
#pragma unroll
            for (int n = 0; n < N; ++n)
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    float* dst = p.part + (((((size_t)b * p.n_chunks + chunk) * N + n) * KC + kk) * 2) * Tp + blk0 + tt;
                    *reinterpret_cast<float2*>(dst) = vnum[j][n][kk];
                    *reinterpret_cast<float2*>(dst + Tp) = vden[j][n][kk];
                }
        }
    }
}

template <int MINB>
static int launch_mu_fused_t(bss_handle* h, const MuArgs& a, int* n_chunks_out, bool* done) {
    constexpr int C = 4, KC = 2;
    constexpr int STG = MINB >= 4 ? FU_STAGES - 1 : FU_STAGES;
    const int n_blocks = (a.Tp + BSS_XSLAB - 1) / BSS_XSLAB;
    FusedParams p{};
    p.a = a;
    p.n_blocks = n_blocks;
    const int blk_frames = a.Tp < BSS_XSLAB ? a.Tp : BSS_XSLAB;
    p.pb_stride = round_up(C * C * 8 + C * KC * 4, 16);
    p.par_off = (uint32_t)round_up(C * blk_frames * 8, 16);
    p.stage_bytes = (uint32_t)round_up((int)p.par_off + p.pb_stride, 128);
    const size_t smem_bytes = (size_t)FU_HEAD_BYTES + (size_t)n_blocks * STG * p.stage_bytes;
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(mu_fused_kernel<C, KC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    int ctas_per_sm = 1;
    BSS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, mu_fused_kernel<C, KC, MINB>, n_blocks * 32, smem_bytes));
    if (ctas_per_sm < 1) return BSS_OK;
    // bin chunks: whole waves of CTAs, chunks of at least 16 bins, at most 4 waves (ties: fewer chunks = fewer partial sums)
    const long long slots = (long long)h->n_sm * ctas_per_sm;
    int best = 1;
    double best_eff = -1.0;
    const int c_max = (int)std::max<long long>(1, std::min<long long>(a.F / 16, cdiv(4 * slots, a.B)));
    for (int c = 1; c <= c_max; ++c) {
        const double waves = (double)((long long)a.B * c) / (double)slots;
        const double eff = waves / std::ceil(waves);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = c;
        }
    }
    if (h->opt_act_chunks > 0) best = std::min(h->opt_act_chunks, a.F);   // BSS_OPT_ACT_CHUNKS: reproduce another batch's order
    p.bins_per_chunk = (int)cdiv(a.F, best);
    p.n_chunks = (int)cdiv(a.F, p.bins_per_chunk);
    const long long n_ctas = (long long)a.B * p.n_chunks;
    if (n_ctas > 0x7fffffffLL) return BSS_OK;
    const size_t need = (size_t)a.B * p.n_chunks * C * KC * 2 * a.Tp;
    if (need > h->part_elems) {
        if (h->part) cudaFree(h->part);
        h->part = nullptr;
        h->part_elems = 0;
        BSS_CUDA(h, cudaMalloc(&h->part, need * sizeof(float)));
        h->part_elems = need;
    }
    const size_t pb_bytes = (size_t)a.B * a.F * p.pb_stride;
    BSS_TRY(ensure_staging(h, pb_bytes));
    p.pbin = (const unsigned char*)h->staging;
    p.part = h->part;
    const long long pieces = (long long)a.B * a.F * (p.pb_stride >> 4);
    fused_pack_kernel<<<(unsigned)cdiv(pieces, 256), 256, 0, h->stream>>>(a.Wf, a.basis, (unsigned char*)h->staging, a.B, C, C, a.F, KC,
                                                                        p.pb_stride);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    mu_fused_kernel<C, KC, MINB><<<(unsigned)n_ctas, n_blocks * 32, smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    if (n_chunks_out) *n_chunks_out = p.n_chunks;
    h->last_act_chunks = p.n_chunks;
    *done = true;
    return BSS_OK;
}

}  // namespace

// Basis update (into a.basis_out) and the activation statistics (into h->part, *n_chunks_out chunks) in one pass over X.
// *done = false when the configuration is not covered (the caller then runs the three-pass form).
int launch_mu_fused(bss_handle* h, const MuArgs& a, int* n_chunks_out, bool* done) {
    *done = false;
    if (a.C != 4 || a.K != 2 || a.Y || a.raw || a.sel_m >= 0 || a.mode != 0 || a.p_exp != 2.f || a.q_exp != 0.5f) return BSS_OK;
    const int n_blocks = (a.Tp + BSS_XSLAB - 1) / BSS_XSLAB;
    if (n_blocks > FU_MAX_WARPS || a.Tp < 2) return BSS_OK;
    // resident CTAs per SM the kernel is compiled for (measurement aid: BSSGPU_FUSED_CTAS = 2, 3 or 4)
    static const int ctas = [] {
        const char* e = getenv("BSSGPU_FUSED_CTAS");
        const int v = e ? atoi(e) : 3;
        return v >= 2 && v <= 4 ? v : 3;
    }();
    if (ctas == 2) return launch_mu_fused_t<2>(h, a, n_chunks_out, done);
    if (ctas == 4) return launch_mu_fused_t<4>(h, a, n_chunks_out, done);
    return launch_mu_fused_t<3>(h, a, n_chunks_out, done);
}
