// FastMNMF (src/bss/mnmf.py:637-946): jointly-diagonalisable multichannel NMF.
//   x~[f,t,m] = |sum_c Q[f,m,c] x[c,f,t]|^2,  Lambda[n,f,t] = sum_k W[n,f,k] H[n,k,t],
//   R[f,t,m]  = max(sum_n Lambda[n,f,t] g[n,f,m], eps)
// None of x~, Lambda, R (each (F,T,M)-sized, 134 MB at cfg4) nor the reference's (N,F,T,M) broadcast
// temporaries (1.07 GB) is stored: every kernel recomputes them from the staged bin tile of X and the
// small factors.  One warp owns a bin; lanes own frame pairs; per-bin parameters (Q, g, W rows) sit in
// the warp's shared-memory scratch and are read as broadcasts.
//   basis W      per-bin reduction over frames            mnmf.py:790-800
//   activation H cross-bin, two deterministic stages      mnmf.py:802-813
//   spatial g    per-bin reduction over frames            mnmf.py:832-844
//   weights 1/R  for the diagonaliser's covariances       mnmf.py:867-868
//   normalise    mnmf.py:753-767,  loss :890-917,  separate (multichannel Wiener filter) :919-946
#include <cstdlib>

#include "handle.h"
#include "smallmat.cuh"

namespace {

constexpr int MN_STAGES = 2;   // two stages leave L1 room for the activation rows (cfg4: 1.39 -> 1.16 ms per iteration)
constexpr int MN_SLAB = 128;
constexpr int MN_NMAX = 8;     // sources
constexpr int MN_KC = 2;       // basis vectors accumulated per pass
constexpr int MN_NG = 2;       // sources per item of the spatial update

struct MnArgs {
    const cf* X;          // [B][F][M][Tp]
    const float* xt;      // [B][F][M][Tp] x~ = |Q x|^2 in the block-interleaved tile layout (rows = M), see mnmf_xt_kernel
    const cf* Qf;         // [B][F][M][M]
    const float* G;       // [B][N][F][M]
    const float* basis;   // [B][N][F][K]
    const float* act;     // [B][N][K][Tp]
    int B, F, M, N, T, Tp, K;
    float eps;
};

struct MnParams {
    MnArgs a;
    TileGeom g;
    long long n_items;
    int per_bin;          // items per bin
    uint32_t scratch_off, scratch_stride, ring_off;
    float* out_f;         // basis_out / G_out
    double* out_d;        // loss terms
    cf* out_c;            // separated output
    const cf* qinv;       // [B][F][M] row `reference_id` of Q^-1
};

// scratch (floats): Qs [2 M M] | gs [N M] | tb [N K] | red [64]
template <int M>
__device__ __forceinline__ void mn_scratch(float* base, int N, int K, float*& Qs, float*& gs, float*& tb, float*& red) {
    Qs = base;
    gs = Qs + 2 * M * M;
    tb = gs + MN_NMAX * M;
    red = tb + MN_NMAX * K;
}
template <int M>
static inline size_t mn_scratch_bytes(int K) {
    return (size_t)(2 * M * M + MN_NMAX * M + MN_NMAX * K + 64) * sizeof(float);
}

template <int M>
__device__ __forceinline__ void mn_load_bin(const MnArgs& a, long long bf, int b, int f, float* Qs, float* gs, float* tb, int lane) {
    const cf* q = a.Qf + (size_t)bf * M * M;
    for (int i = lane; i < M * M; i += 32) reinterpret_cast<float2*>(Qs)[i] = __ldg(q + i);
    for (int i = lane; i < a.N * M; i += 32) {
        const int n = i / M, m = i - n * M;
        gs[i] = __ldg(a.G + (((size_t)b * a.N + n) * a.F + f) * M + m);
    }
    for (int i = lane; i < a.N * a.K; i += 32) {
        const int n = i / a.K, k = i - n * a.K;
        tb[i] = __ldg(a.basis + (((size_t)b * a.N + n) * a.F + f) * a.K + k);
    }
    __syncwarp();
}

// y_m = sum_c Q[m][c] x_c for the two frames held in xv
template <int M>
__device__ __forceinline__ void mn_project(const float4 (&xv)[M], const float* Qs, float2 (&y0)[M], float2 (&y1)[M]) {
#pragma unroll
    for (int m = 0; m < M; ++m) {
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < M; ++c) {
            const float2 w = reinterpret_cast<const float2*>(Qs)[m * M + c];
            const float2 x0 = make_float2(xv[c].x, xv[c].y), x1 = make_float2(xv[c].z, xv[c].w);
            const float2 wx = make_float2(w.x, w.x), wy = make_float2(w.y, w.y);
            a0 = __ffma2_rn(x0, wx, a0);
            a0 = __ffma2_rn(make_float2(-x0.y, x0.x), wy, a0);
            a1 = __ffma2_rn(x1, wx, a1);
            a1 = __ffma2_rn(make_float2(-x1.y, x1.x), wy, a1);
        }
        y0[m] = a0;
        y1[m] = a1;
    }
}

template <int M>
__device__ __forceinline__ void mn_power(const float4 (&xv)[M], const float* Qs, float2 (&xt)[M]) {
    float2 y0[M], y1[M];
    mn_project<M>(xv, Qs, y0, y1);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const float2 s0 = __fmul2_rn(y0[m], y0[m]), s1 = __fmul2_rn(y1[m], y1[m]);
        xt[m] = make_float2(s0.x + s0.y, s1.x + s1.y);
    }
}

// Lambda[n] for the frame pair at hrow = act + b N K Tp + t
__device__ __forceinline__ void mn_lambda(const MnArgs& a, const float* tb, const float* hrow, float2 (&lam)[MN_NMAX]) {
#pragma unroll
    for (int n = 0; n < MN_NMAX; ++n) {
        float2 l = make_float2(0.f, 0.f);
        if (n < a.N) {
            for (int k = 0; k < a.K; ++k) {
                const float2 hv = __ldg(reinterpret_cast<const float2*>(hrow + ((size_t)n * a.K + k) * a.Tp));
                const float tk = tb[n * a.K + k];
                l = __ffma2_rn(hv, make_float2(tk, tk), l);
            }
        }
        lam[n] = l;
    }
}

// raw[m] = sum_n Lambda[n] g[n][m]
template <int M>
__device__ __forceinline__ void mn_variance(const MnArgs& a, const float* gs, const float2 (&lam)[MN_NMAX], float2 (&raw)[M]) {
#pragma unroll
    for (int m = 0; m < M; ++m) {
        float2 r = make_float2(0.f, 0.f);
#pragma unroll
        for (int n = 0; n < MN_NMAX; ++n)
            if (n < a.N) {
                const float g = gs[n * M + m];
                r = __ffma2_rn(lam[n], make_float2(g, g), r);
            }
        raw[m] = r;
    }
}

__device__ __forceinline__ float2 floor2(float2 v, float eps) { return make_float2(fmaxf(v.x, eps), fmaxf(v.y, eps)); }
__device__ __forceinline__ float2 rcp2n(float2 v) { return make_float2(__frcp_rn(v.x), __frcp_rn(v.y)); }

// ------------------------------------------------------------------------------------------- x~ = |Q x|^2
// Q only changes at the end of an iteration, so the diagonalised power is computed once per iteration (one pass over X)
// and the three source / spatial-model passes stream x~ (half the bytes of X, no 8 x 8 complex product per frame).
template <int M>
__global__ void __launch_bounds__(256, 1) mnmf_xt_kernel(const MnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const MnArgs& a = p.a;
    float *Qs, *gs, *tb, *red;
    mn_scratch<M>(reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride), a.N, a.K, Qs, gs, tb, red);
    WarpStream<MN_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * MN_STAGES,
             smem + p.ring_off + (size_t)warp * MN_STAGES * p.g.stage_bytes, a.X, (int)(blockIdx.x * wpc + warp),
             (int)(gridDim.x * wpc), (int)p.n_items, 1, lane);
#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, a.X, 1);
        const int bf = st.cons.item;
        if (st.first_slab()) {
            __syncwarp();
            const cf* q = a.Qf + (size_t)bf * M * M;
            for (int i = lane; i < M * M; i += 32) reinterpret_cast<float2*>(Qs)[i] = __ldg(q + i);
            __syncwarp();
        }
        const cf* xs = st.acquire(p.g);
        const int nf = st.frames(p.g);
        const int tbase = st.frame0(p.g);
        constexpr int MP = (M + 1) & ~1;   // x~ tiles have an even number of rows: every block is a multiple of 16 bytes
        float* out = p.out_f + (size_t)bf * MP * a.Tp + (size_t)tbase * MP;   // this block of the x~ tile: [MP][nf]
#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float4 xv[M];
#pragma unroll
            for (int c = 0; c < M; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * nf + tt);
            float2 xt[M];
            mn_power<M>(xv, Qs, xt);
#pragma unroll
            for (int m = 0; m < M; ++m) *reinterpret_cast<float2*>(out + (size_t)m * nf + tt) = xt[m];
        }
        st.release(p.g);
    }
}

// ------------------------------------------------------------------------------------------- basis W
// item = (bin, chunk of MN_KC basis vectors)
template <int M>
__global__ void __launch_bounds__(256, 1) mnmf_basis_kernel(const MnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const MnArgs& a = p.a;
    float *Qs, *gs, *tb, *red;
    mn_scratch<M>(reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride), a.N, a.K, Qs, gs, tb, red);
    WarpStream<MN_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * MN_STAGES,
             smem + p.ring_off + (size_t)warp * MN_STAGES * p.g.stage_bytes, reinterpret_cast<const cf*>(a.xt), (int)(blockIdx.x * wpc + warp),
             (int)(gridDim.x * wpc), (int)p.n_items, p.per_bin, lane);
    float2 num[MN_NMAX][MN_KC], den[MN_NMAX][MN_KC];
#pragma unroll
    for (int n = 0; n < MN_NMAX; ++n)
#pragma unroll
        for (int kk = 0; kk < MN_KC; ++kk) num[n][kk] = den[n][kk] = make_float2(0.f, 0.f);
    int b = 0, f = 0, k0 = 0;
    int bf = 0;
#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, reinterpret_cast<const cf*>(a.xt), p.per_bin);
        if (st.first_slab()) {
            bf = st.cons.item / p.per_bin;
            k0 = (st.cons.item - bf * p.per_bin) * MN_KC;
            b = bf / a.F;
            f = bf - b * a.F;
            __syncwarp();
            mn_load_bin<M>(a, bf, b, f, Qs, gs, tb, lane);
        }
        const cf* xs = st.acquire(p.g);
        // the stream moves x~ as pairs of floats: st.frames() counts pairs, nf counts frames
        const int nfp = st.frames(p.g);
        const int nf = 2 * nfp;
        const int tbase = 2 * st.frame0(p.g);
#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float2 xt[M];
#pragma unroll
            for (int m = 0; m < M; ++m) xt[m] = reinterpret_cast<const float2*>(xs)[m * nfp + (tt >> 1)];
            const float* hrow = a.act + (size_t)b * a.N * a.K * a.Tp + tbase + tt;
            float2 lam[MN_NMAX];
            mn_lambda(a, tb, hrow, lam);
            float2 R[M];
            mn_variance<M>(a, gs, lam, R);
            float2 u[M], ri[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                ri[m] = rcp2n(floor2(R[m], a.eps));
                u[m] = __fmul2_rn(xt[m], __fmul2_rn(ri[m], ri[m]));
            }
#pragma unroll
            for (int n = 0; n < MN_NMAX; ++n)
                if (n < a.N) {
                    float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const float g = gs[n * M + m];
                        sa = __ffma2_rn(u[m], make_float2(g, g), sa);
                        sb = __ffma2_rn(ri[m], make_float2(g, g), sb);
                    }
#pragma unroll
                    for (int kk = 0; kk < MN_KC; ++kk)
                        if (k0 + kk < a.K) {
                            const float2 hv = __ldg(reinterpret_cast<const float2*>(hrow + ((size_t)n * a.K + k0 + kk) * a.Tp));
                            num[n][kk] = __ffma2_rn(sa, hv, num[n][kk]);
                            den[n][kk] = __ffma2_rn(sb, hv, den[n][kk]);
                        }
                }
        }
        if (st.last_slab(p.g)) {
            constexpr int MP = MN_NMAX * MN_KC * 2;   // 32
            float flat[MP];
#pragma unroll
            for (int n = 0; n < MN_NMAX; ++n)
#pragma unroll
                for (int kk = 0; kk < MN_KC; ++kk) {
                    flat[(n * MN_KC + kk) * 2] = num[n][kk].x + num[n][kk].y;
                    flat[(n * MN_KC + kk) * 2 + 1] = den[n][kk].x + den[n][kk].y;
                    num[n][kk] = den[n][kk] = make_float2(0.f, 0.f);
                }
            warp_reduce_scatter<MP>(flat, lane);
            // lane L holds element L: (n, kk, which) = (L / 4, (L / 2) % 2, L % 2)
            const float mine = flat[0];
            const float other = __shfl_down_sync(BSS_FULL, mine, 1);
            const int n = lane / (2 * MN_KC), kk = (lane >> 1) % MN_KC, k = k0 + kk;
            if ((lane & 1) == 0 && n < a.N && k < a.K) {
                const float dn = fmaxf(other, a.eps);
                const size_t idx = (((size_t)b * a.N + n) * a.F + f) * a.K + k;
                p.out_f[idx] = a.basis[idx] * sqrtf(mine / dn);
            }
        }
        st.release(p.g);
    }
}

// ------------------------------------------------------------------------------------------- activation H
// Stage 1: a warp owns 64 frames of one mixture and walks a chunk of bins; part [B][n_chunks][N][K][2][Tp]
template <int M>
__global__ void __launch_bounds__(128) mnmf_act_partial_kernel(const MnArgs a, float* part, int n_chunks, int bins_per_chunk,
                                                              int n_slabs, int n_kc, long long n_items, uint32_t scratch_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const long long item = (int)(blockIdx.x * wpc + warp);
    if (item >= n_items) return;
    long long r = item;
    const int kc = (int)(r % n_kc);
    r /= n_kc;
    const int slab = (int)(r % n_slabs);
    r /= n_slabs;
    const int chunk = (int)(r % n_chunks);
    const int b = (int)(r / n_chunks);
    const int k0 = kc * MN_KC;
    const int t0 = slab * 64 + 2 * lane;
    const bool live = t0 < a.Tp;
    const int tl = live ? t0 : 0;
    float *Qs, *gs, *tb, *red;
    mn_scratch<M>(reinterpret_cast<float*>(smem + (size_t)warp * scratch_stride), a.N, a.K, Qs, gs, tb, red);
    float2 num[MN_NMAX][MN_KC], den[MN_NMAX][MN_KC];
#pragma unroll
    for (int n = 0; n < MN_NMAX; ++n)
#pragma unroll
        for (int kk = 0; kk < MN_KC; ++kk) num[n][kk] = den[n][kk] = make_float2(0.f, 0.f);
    const float* hrow = a.act + (size_t)b * a.N * a.K * a.Tp + tl;
    constexpr int MP = (M + 1) & ~1;   // rows of an x~ tile (see mnmf_xt_kernel)
    const size_t xoff = tile_off(MP, a.Tp, 0, tl);
    const int xlen = (int)(tile_off(MP, a.Tp, 1, tl) - xoff);
    const int f_begin = chunk * bins_per_chunk;
    const int f_end = min(a.F, f_begin + bins_per_chunk);
#pragma unroll 1
    for (int f = f_begin; f < f_end; ++f) {
        const long long bf = (long long)b * a.F + f;
        __syncwarp();
        mn_load_bin<M>(a, bf, b, f, Qs, gs, tb, lane);
        float2 xt[M];
#pragma unroll
        for (int m = 0; m < M; ++m)
            xt[m] = live ? __ldg(reinterpret_cast<const float2*>(a.xt + (size_t)bf * MP * a.Tp + xoff + (size_t)m * xlen)) : make_float2(0.f, 0.f);
        float2 lam[MN_NMAX];
        mn_lambda(a, tb, hrow, lam);
        float2 R[M];
        mn_variance<M>(a, gs, lam, R);
        float2 u[M], ri[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            ri[m] = rcp2n(floor2(R[m], a.eps));
            u[m] = __fmul2_rn(xt[m], __fmul2_rn(ri[m], ri[m]));
        }
#pragma unroll
        for (int n = 0; n < MN_NMAX; ++n)
            if (n < a.N) {
                float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float g = gs[n * M + m];
                    sa = __ffma2_rn(u[m], make_float2(g, g), sa);
                    sb = __ffma2_rn(ri[m], make_float2(g, g), sb);
                }
#pragma unroll
                for (int kk = 0; kk < MN_KC; ++kk)
                    if (k0 + kk < a.K) {
                        const float w = tb[n * a.K + k0 + kk];
                        num[n][kk] = __ffma2_rn(sa, make_float2(w, w), num[n][kk]);
                        den[n][kk] = __ffma2_rn(sb, make_float2(w, w), den[n][kk]);
                    }
            }
    }
    if (!live) return;
#pragma unroll
    for (int n = 0; n < MN_NMAX; ++n)
#pragma unroll
        for (int kk = 0; kk < MN_KC; ++kk) {
            const int k = k0 + kk;
            if (n < a.N && k < a.K) {
                float* dst = part + (((((size_t)b * n_chunks + chunk) * a.N + n) * a.K + k) * 2) * a.Tp + t0;
                *reinterpret_cast<float2*>(dst) = num[n][kk];
                *reinterpret_cast<float2*>(dst + a.Tp) = den[n][kk];
            }
        }
}

__global__ void __launch_bounds__(256) mnmf_act_finish_kernel(const MnArgs a, const float* part, float* act, int n_chunks) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)a.B * a.N * a.K * a.Tp;
    if (idx >= total) return;
    const int t = (int)(idx % a.Tp);
    long long r = idx / a.Tp;
    const int k = (int)(r % a.K);
    r /= a.K;
    const int n = (int)(r % a.N);
    const int b = (int)(r / a.N);
    if (t >= a.T) {
        act[idx] = 0.f;
        return;
    }
    float num = 0.f, den = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
        const float* src = part + (((((size_t)b * n_chunks + c) * a.N + n) * a.K + k) * 2) * a.Tp + t;
        num += src[0];
        den += src[a.Tp];
    }
    den = fmaxf(den, a.eps);
    act[idx] = act[idx] * sqrtf(num / den);
}

// ------------------------------------------------------------------------------------------- spatial g
// item = (bin, group of MN_MG channels):  g[n,f,m] *= sqrt(sum_t Lambda x~/R^2 / max(sum_t Lambda/R, eps)).
// Grouping by channel (not by source) keeps the per-item front end small: only MN_MG rows of Q x and of R are needed.
constexpr int MN_MG = 2;
template <int M>
__global__ void __launch_bounds__(256, 1) mnmf_scm_kernel(const MnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const MnArgs& a = p.a;
    float *Qs, *gs, *tb, *red;
    mn_scratch<M>(reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride), a.N, a.K, Qs, gs, tb, red);
    WarpStream<MN_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * MN_STAGES,
             smem + p.ring_off + (size_t)warp * MN_STAGES * p.g.stage_bytes, reinterpret_cast<const cf*>(a.xt), (int)(blockIdx.x * wpc + warp),
             (int)(gridDim.x * wpc), (int)p.n_items, p.per_bin, lane);
    float2 A[MN_NMAX][MN_MG], Bq[MN_NMAX][MN_MG];
#pragma unroll
    for (int n = 0; n < MN_NMAX; ++n)
#pragma unroll
        for (int j = 0; j < MN_MG; ++j) A[n][j] = Bq[n][j] = make_float2(0.f, 0.f);
    int b = 0, f = 0, m0 = 0;
    int bf = 0;
#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, reinterpret_cast<const cf*>(a.xt), p.per_bin);
        if (st.first_slab()) {
            bf = st.cons.item / p.per_bin;
            m0 = (st.cons.item - bf * p.per_bin) * MN_MG;
            b = bf / a.F;
            f = bf - b * a.F;
            __syncwarp();
            mn_load_bin<M>(a, bf, b, f, Qs, gs, tb, lane);
        }
        const cf* xs = st.acquire(p.g);
        // the stream moves x~ as pairs of floats: st.frames() counts pairs, nf counts frames
        const int nfp = st.frames(p.g);
        const int nf = 2 * nfp;
        const int tbase = 2 * st.frame0(p.g);
#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            const float* hrow = a.act + (size_t)b * a.N * a.K * a.Tp + tbase + tt;
            float2 lam[MN_NMAX];
            mn_lambda(a, tb, hrow, lam);
            float2 u[MN_MG], ri[MN_MG];
#pragma unroll
            for (int j = 0; j < MN_MG; ++j) {
                const int m = min(m0 + j, M - 1);
                const float2 xt = reinterpret_cast<const float2*>(xs)[m * nfp + (tt >> 1)];
                float2 r = make_float2(0.f, 0.f);
#pragma unroll
                for (int n = 0; n < MN_NMAX; ++n)
                    if (n < a.N) {
                        const float g = gs[n * M + m];
                        r = __ffma2_rn(lam[n], make_float2(g, g), r);
                    }
                ri[j] = rcp2n(floor2(r, a.eps));
                u[j] = __fmul2_rn(xt, __fmul2_rn(ri[j], ri[j]));
            }
#pragma unroll
            for (int n = 0; n < MN_NMAX; ++n)
#pragma unroll
                for (int j = 0; j < MN_MG; ++j) {
                    A[n][j] = __ffma2_rn(lam[n], u[j], A[n][j]);
                    Bq[n][j] = __ffma2_rn(lam[n], ri[j], Bq[n][j]);
                }
        }
        if (st.last_slab(p.g)) {
            constexpr int MP = MN_NMAX * MN_MG * 2;   // 32
            static_assert(MP == 32, "one value per lane");
            float flat[MP];
#pragma unroll
            for (int n = 0; n < MN_NMAX; ++n)
#pragma unroll
                for (int j = 0; j < MN_MG; ++j) {
                    flat[(n * MN_MG + j) * 2] = A[n][j].x + A[n][j].y;
                    flat[(n * MN_MG + j) * 2 + 1] = Bq[n][j].x + Bq[n][j].y;
                    A[n][j] = Bq[n][j] = make_float2(0.f, 0.f);
                }
            warp_reduce_scatter<MP>(flat, lane);
            // lane L holds element L: (n, j, which) = (L / 4, (L / 2) % 2, L % 2)
            const float mine = flat[0];
            const float other = __shfl_down_sync(BSS_FULL, mine, 1);
            const int n = lane / (2 * MN_MG), m = m0 + ((lane >> 1) % MN_MG);
            if ((lane & 1) == 0 && n < a.N && m < M) {
                const float bv = fmaxf(other, a.eps);
                const size_t idx = (((size_t)b * a.N + n) * a.F + f) * M + m;
                p.out_f[idx] = a.G[idx] * sqrtf(mine / bv);
            }
        }
        st.release(p.g);
    }
}

// ------------------------------------------------------------------------------------------- spatial g on tensor cores
// The same update as one small contraction per bin:  A[n][m] = sum_t Lambda[n,t] u[m,t],  B[n][m] = sum_t Lambda[n,t] / R[m,t]
// with u = x~ / R^2: two [8 x T] x [T x 8] products sharing their left operand.  One warp per bin, mma.sync.m16n8k8 (3xTF32),
// K = 8 frames per step: lane (g, tig) computes Lambda of source g and u, 1/R of channel g at frames tig and tig + 4 (the
// Lambda of the other sources, needed for R, come by shuffle), i.e. exactly its own fragment entries; the D fragments hold
// A[g][2 tig .. 2 tig + 1] and B[g][..], from which the lane updates its two entries of g.  11 warp instructions per frame
// instead of 41 for the CUDA-core kernel above (kept for reference and for BSSGPU_NO_MMA).
__device__ __forceinline__ uint32_t mn_round_tf32(float v) { return (__float_as_uint(v) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void mn_split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = mn_round_tf32(v);
    lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mn_mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int M>
__global__ void __launch_bounds__(512, 1) mnmf_scm_mma_kernel(const MnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const MnArgs& a = p.a;
    float *Qs, *gs, *tb, *red;
    mn_scratch<M>(reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride), a.N, a.K, Qs, gs, tb, red);
    WarpStream<MN_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * MN_STAGES,
             smem + p.ring_off + (size_t)warp * MN_STAGES * p.g.stage_bytes, reinterpret_cast<const cf*>(a.xt),
             (int)(blockIdx.x * wpc + warp), (int)(gridDim.x * wpc), (int)p.n_items, 1, lane);
    const bool n_live = g < a.N, m_live = g < M;
    float dA[4], dAc[4], dB[4], dBc[4];
    int b = 0, f = 0;
#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, reinterpret_cast<const cf*>(a.xt), 1);
        const int bf = st.cons.item;
        if (st.first_slab()) {
            b = bf / a.F;
            f = bf - b * a.F;
            __syncwarp();
            mn_load_bin<M>(a, bf, b, f, Qs, gs, tb, lane);
#pragma unroll
            for (int q = 0; q < 4; ++q) dA[q] = dAc[q] = dB[q] = dBc[q] = 0.f;
        }
        const float* xs = reinterpret_cast<const float*>(st.acquire(p.g));
        const int nf = 2 * st.frames(p.g);          // frames of this block
        const int tbase = 2 * st.frame0(p.g);
        const float* hrow = a.act + ((size_t)b * a.N + (n_live ? g : 0)) * a.K * a.Tp + tbase;
#pragma unroll 1
        for (int k0 = 0; k0 < nf; k0 += 8) {
            const int tA = k0 + tig, tB = tA + 4;
            const bool lA = tA < nf, lB = tB < nf;
            // Lambda of this lane group's source at its two frames
            float lamA = 0.f, lamB = 0.f;
            if (n_live) {
                for (int k = 0; k < a.K; ++k) {
                    const float tk = tb[g * a.K + k];
                    if (lA) lamA = fmaf(tk, __ldg(hrow + (size_t)k * a.Tp + tA), lamA);
                    if (lB) lamB = fmaf(tk, __ldg(hrow + (size_t)k * a.Tp + tB), lamB);
                }
            }
            // R of this lane group's channel: the other sources' Lambda come from lanes (n, tig)
            float RA = 0.f, RB = 0.f;
#pragma unroll
            for (int n = 0; n < MN_NMAX; ++n) {
                const float la = __shfl_sync(BSS_FULL, lamA, n * 4 + tig);
                const float lb = __shfl_sync(BSS_FULL, lamB, n * 4 + tig);
                if (n < a.N && m_live) {
                    const float gg = gs[n * M + g];
                    RA = fmaf(la, gg, RA);
                    RB = fmaf(lb, gg, RB);
                }
            }
            const float riA = __frcp_rn(fmaxf(RA, a.eps)), riB = __frcp_rn(fmaxf(RB, a.eps));
            const float xA = (m_live && lA) ? xs[g * nf + tA] : 0.f;
            const float xB = (m_live && lB) ? xs[g * nf + tB] : 0.f;
            const float uA = xA * riA * riA, uB = xB * riB * riB;
            uint32_t ah[4], al[4], uh[2], ul[2], rh[2], rl[2];
            mn_split_tf32(lamA, ah[0], al[0]);
            mn_split_tf32(lamB, ah[2], al[2]);
            ah[1] = ah[3] = al[1] = al[3] = 0u;
            mn_split_tf32(uA, uh[0], ul[0]);
            mn_split_tf32(uB, uh[1], ul[1]);
            mn_split_tf32((m_live && lA) ? riA : 0.f, rh[0], rl[0]);
            mn_split_tf32((m_live && lB) ? riB : 0.f, rh[1], rl[1]);
            mn_mma_tf32(dAc, al, uh);
            mn_mma_tf32(dBc, al, rh);
            mn_mma_tf32(dAc, ah, ul);
            mn_mma_tf32(dBc, ah, rl);
            mn_mma_tf32(dA, ah, uh);
            mn_mma_tf32(dB, ah, rh);
        }
        if (st.last_slab(p.g)) {
            // c0, c1 = rows n = g, columns m = 2 tig, 2 tig + 1
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int m = 2 * tig + h2;
                if (n_live && m < M) {
                    const float av = dA[h2] + dAc[h2];
                    const float bv = fmaxf(dB[h2] + dBc[h2], a.eps);
                    const size_t idx = (((size_t)b * a.N + g) * a.F + f) * M + m;
                    p.out_f[idx] = a.G[idx] * sqrtf(av / bv);
                }
            }
        }
        st.release(p.g);
    }
}

// ------------------------------------------------------------------------------------------- weights 1/R
// iw [B][F][M][Tp]: inverse of the floored variance (no pass over X needed)
template <int M>
__global__ void __launch_bounds__(128) mnmf_weights_kernel(const MnArgs a, float* iw, long long n_items, int tiled,
                                                          uint32_t scratch_stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + warp;   // (bin, 128-frame block)
    if (item >= n_items) return;
    const int n_blocks = (a.Tp + BSS_XSLAB - 1) / BSS_XSLAB;
    const long long bf = item / n_blocks;
    const int blk = (int)(item - bf * n_blocks);
    const int b = (int)(bf / a.F), f = (int)(bf - (long long)b * a.F);
    float *Qs, *gs, *tb, *red;
    mn_scratch<M>(reinterpret_cast<float*>(smem + (size_t)warp * scratch_stride), a.N, a.K, Qs, gs, tb, red);
    // only g and the basis row are needed here
    for (int i = lane; i < a.N * M; i += 32) {
        const int n = i / M, m = i - n * M;
        gs[i] = __ldg(a.G + (((size_t)b * a.N + n) * a.F + f) * M + m);
    }
    for (int i = lane; i < a.N * a.K; i += 32) {
        const int n = i / a.K, k = i - n * a.K;
        tb[i] = __ldg(a.basis + (((size_t)b * a.N + n) * a.F + f) * a.K + k);
    }
    __syncwarp();
    const int t_end = min(a.Tp, (blk + 1) * BSS_XSLAB);
    for (int t = blk * BSS_XSLAB + 2 * lane; t < t_end; t += 64) {
        float2 lam[MN_NMAX];
        mn_lambda(a, tb, a.act + (size_t)b * a.N * a.K * a.Tp + t, lam);
        float2 R[M];
        mn_variance<M>(a, gs, lam, R);
#pragma unroll
        for (int m = 0; m < M; ++m)
            // tiled: block-interleaved like X (rows = M), so that the tensor-core covariance kernel stages a block with one copy
            *reinterpret_cast<float2*>(iw + (tiled ? (size_t)bf * M * a.Tp + tile_off(M, a.Tp, m, t) : ((size_t)bf * M + m) * a.Tp + t)) =
                rcp2n(floor2(R[m], a.eps));
    }
}

// ------------------------------------------------------------------------------------------- loss
// per bin: sum_{t<T, m} (x~ + eps)/(y~ + eps) + log(y~ + eps), y~ unfloored     mnmf.py:907-915
template <int M>
__global__ void __launch_bounds__(256, 1) mnmf_loss_kernel(const MnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const MnArgs& a = p.a;
    float *Qs, *gs, *tb, *red;
    mn_scratch<M>(reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride), a.N, a.K, Qs, gs, tb, red);
    WarpStream<MN_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * MN_STAGES,
             smem + p.ring_off + (size_t)warp * MN_STAGES * p.g.stage_bytes, reinterpret_cast<const cf*>(a.xt), (int)(blockIdx.x * wpc + warp),
             (int)(gridDim.x * wpc), (int)p.n_items, 1, lane);
    double total = 0.0;
    int b = 0, f = 0;
#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, reinterpret_cast<const cf*>(a.xt), 1);
        const int bf = st.cons.item;
        if (st.first_slab()) {
            b = bf / a.F;
            f = bf - b * a.F;
            __syncwarp();
            mn_load_bin<M>(a, bf, b, f, Qs, gs, tb, lane);
            total = 0.0;
        }
        const cf* xs = st.acquire(p.g);
        // the stream moves x~ as pairs of floats: st.frames() counts pairs, nf counts frames
        const int nfp = st.frames(p.g);
        const int nf = 2 * nfp;
        const int tbase = 2 * st.frame0(p.g);
        float part = 0.f;
#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float2 xt[M];
#pragma unroll
            for (int m = 0; m < M; ++m) xt[m] = reinterpret_cast<const float2*>(xs)[m * nfp + (tt >> 1)];
            const int t = tbase + tt;
            float2 lam[MN_NMAX];
            mn_lambda(a, tb, a.act + (size_t)b * a.N * a.K * a.Tp + t, lam);
            float2 R[M];
            mn_variance<M>(a, gs, lam, R);
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float y0 = R[m].x + a.eps, y1 = R[m].y + a.eps;
                if (t < a.T) part += (xt[m].x + a.eps) / y0 + logf(y0);
                if (t + 1 < a.T) part += (xt[m].y + a.eps) / y1 + logf(y1);
            }
        }
        total += (double)part;
        if (st.last_slab(p.g)) {
            const double s = warp_sum(total);
            if (lane == 0) p.out_d[bf] = s;
        }
        st.release(p.g);
    }
}

// ------------------------------------------------------------------------------------------- separate
// x^[n,f,t] = sum_m Qinv[ref][m] (Qx)[m] Lambda[n] g[n][m] / max(sum_n' Lambda g, eps)      mnmf.py:923-946
template <int M>
__global__ void __launch_bounds__(256, 1) mnmf_separate_kernel(const MnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const MnArgs& a = p.a;
    float *Qs, *gs, *tb, *red;
    mn_scratch<M>(reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride), a.N, a.K, Qs, gs, tb, red);
    WarpStream<MN_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * MN_STAGES,
             smem + p.ring_off + (size_t)warp * MN_STAGES * p.g.stage_bytes, a.X, (int)(blockIdx.x * wpc + warp),
             (int)(gridDim.x * wpc), (int)p.n_items, 1, lane);
    int b = 0, f = 0;
    float2 qi[M];
#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, a.X, 1);
        const int bf = st.cons.item;
        if (st.first_slab()) {
            b = bf / a.F;
            f = bf - b * a.F;
            __syncwarp();
            mn_load_bin<M>(a, bf, b, f, Qs, gs, tb, lane);
#pragma unroll
            for (int m = 0; m < M; ++m) qi[m] = __ldg(p.qinv + (size_t)bf * M + m);
        }
        const cf* xs = st.acquire(p.g);
        const int nf = st.frames(p.g);
        const int tbase = st.frame0(p.g);
#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float4 xv[M];
#pragma unroll
            for (int c = 0; c < M; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * nf + tt);
            float2 y0[M], y1[M];
            mn_project<M>(xv, Qs, y0, y1);
            const int t = tbase + tt;
            float2 lam[MN_NMAX];
            mn_lambda(a, tb, a.act + (size_t)b * a.N * a.K * a.Tp + t, lam);
            float2 R[M];
            mn_variance<M>(a, gs, lam, R);
            // z[m] = Qinv[ref][m] * y[m] / y~[m]
            float2 z0[M], z1[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float2 ri = rcp2n(floor2(R[m], a.eps));
                z0[m] = make_float2((qi[m].x * y0[m].x - qi[m].y * y0[m].y) * ri.x, (qi[m].x * y0[m].y + qi[m].y * y0[m].x) * ri.x);
                z1[m] = make_float2((qi[m].x * y1[m].x - qi[m].y * y1[m].y) * ri.y, (qi[m].x * y1[m].y + qi[m].y * y1[m].x) * ri.y);
            }
#pragma unroll
            for (int n = 0; n < MN_NMAX; ++n)
                if (n < a.N) {
                    float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const float g = gs[n * M + m];
                        o0 = __ffma2_rn(z0[m], make_float2(g, g), o0);
                        o1 = __ffma2_rn(z1[m], make_float2(g, g), o1);
                    }
                    cf* o = p.out_c + (((size_t)b * a.N + n) * a.F + f) * a.T + t;
                    if (t < a.T) o[0] = cf_make(o0.x * lam[n].x, o0.y * lam[n].x);
                    if (t + 1 < a.T) o[1] = cf_make(o1.x * lam[n].y, o1.y * lam[n].y);
                }
        }
        st.release(p.g);
    }
}

// row `ref` of Q^-1 per bin, fp64 -> complex64
template <int M>
__global__ void __launch_bounds__(64) mnmf_qinv_kernel(const double2* Qg, cf* out, long long n_bins, int ref, int32_t* flags) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_bins) return;
    Mat<M> Q, Qi;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j) {
            const double2 v = Qg[(size_t)idx * M * M + i * M + j];
            Q.a[i][j] = cd_make(v.x, v.y);
        }
    if (!mat_inverse(Q, Qi)) atomicAdd(flags, 1);
#pragma unroll
    for (int m = 0; m < M; ++m) {
        cd v = Qi.a[0][m];
#pragma unroll
        for (int r = 1; r < M; ++r)
            if (r == ref) v = Qi.a[r][m];
        out[(size_t)idx * M + m] = cf_make((float)v.x, (float)v.y);
    }
}

// ------------------------------------------------------------------------------------------- normalisation
// per bin: s = max(mean_m sum_c |Q[m][c]|^2, eps); Q /= sqrt(s); g /= s; gamma[n] = max(sum_m g, eps); g /= gamma;
// W[n,f,:] *= gamma          mnmf.py:753-762.  One warp per bin (fp64, fixed-order lane reductions).
__global__ void __launch_bounds__(128) mnmf_norm_bin_kernel(double2* Q, cf* Qf, float* G, float* basis, int B, int N, int F, int M,
                                                           int K, double eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (idx >= (long long)B * F) return;
    const int b = (int)(idx / F), f = (int)(idx - (long long)b * F);
    double2* q = Q + (size_t)idx * M * M;
    double s = 0.0;
    for (int i = lane; i < M * M; i += 32) s += q[i].x * q[i].x + q[i].y * q[i].y;
    s = warp_sum(s) / (double)M;
    if (s < eps) s = eps;
    const double is = 1.0 / sqrt(s);
    for (int i = lane; i < M * M; i += 32) {
        double2 v = q[i];
        v.x *= is;
        v.y *= is;
        q[i] = v;
        Qf[(size_t)idx * M * M + i] = cf_make((float)v.x, (float)v.y);
    }
    for (int n = 0; n < N; ++n) {
        float* g = G + (((size_t)b * N + n) * F + f) * M;
        const double gv = lane < M ? (double)g[lane] / s : 0.0;
        double sum = warp_sum(gv);
        if (sum < eps) sum = eps;
        if (lane < M) g[lane] = (float)(gv / sum);
        float* w = basis + (((size_t)b * N + n) * F + f) * K;
        for (int k = lane; k < K; k += 32) w[k] = (float)((double)w[k] * sum);
    }
}

// block per (b, n, k): omega = max(sum_f W[n,f,k], eps); W /= omega; H[n,k,:] *= omega       mnmf.py:764-767
__global__ void __launch_bounds__(256) mnmf_norm_basis_kernel(float* basis, float* act, int N, int F, int K, int Tp, double eps) {
    __shared__ double red[8];
    __shared__ double omega_s;
    const long long bnk = blockIdx.x;
    const int k = (int)(bnk % K);
    const long long bn = bnk / K;
    float* w = basis + (size_t)bn * F * K + k;
    double s = 0.0;
    for (int f = threadIdx.x; f < F; f += blockDim.x) s += (double)w[(size_t)f * K];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        omega_s = t < eps ? eps : t;
    }
    __syncthreads();
    const double om = omega_s;
    for (int f = threadIdx.x; f < F; f += blockDim.x) w[(size_t)f * K] = (float)((double)w[(size_t)f * K] / om);
    float* hh = act + (size_t)bnk * Tp;
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) hh[t] = (float)((double)hh[t] * om);
}

// f32_tiles: the kernel streams x~ (float tiles) as pairs of floats: rows of Tp / 2 pairs, 64-pair blocks
template <int M, typename Kern>
int launch_stream(bss_handle* h, Kern kern, MnParams& p, int per_bin, int slab, int max_wpc, bool f32_tiles = true) {
    p.g = make_tile_geom(M, p.a.Tp, slab);
    if (f32_tiles) {
        p.g.n_rows = (M + 1) & ~1;
        p.g.row_len = p.a.Tp / 2;
        p.g.slab = p.g.n_slabs == 1 ? p.g.row_len : BSS_XSLAB / 2;
        p.g.stage_bytes = (uint32_t)(((size_t)p.g.n_rows * p.g.slab * 8 + 127) / 128 * 128);
    }
    p.per_bin = per_bin;
    p.n_items = (long long)p.a.B * p.a.F * per_bin;
    StreamPlan sp;
    if (!plan_stream(h, p.g, MN_STAGES, mn_scratch_bytes<M>(p.a.K), (int)p.n_items, max_wpc, &sp))
        return bss_fail(h, BSS_EINVAL, "FastMNMF: frame tile does not fit in shared memory");
    p.scratch_off = sp.scratch_off;
    p.scratch_stride = sp.scratch_stride;
    p.ring_off = sp.ring_off;
    BSS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
    kern<<<sp.grid, sp.wpc * 32, sp.smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

MnArgs mn_args(bss_handle* h) {
    MnArgs a{};
    a.X = h->X;
    a.xt = h->xt;
    a.Qf = h->Wf;
    a.G = h->G;
    a.basis = h->basis;
    a.act = h->act;
    a.B = h->B;
    a.F = h->F;
    a.M = h->C;
    a.N = h->N;
    a.T = h->T;
    a.Tp = h->Tp;
    a.K = h->K;
    a.eps = (float)h->cfg.eps;
    return a;
}

template <int M>
int mn_xt(bss_handle* h) {
    MnParams p{};
    p.a = mn_args(h);
    p.out_f = h->xt;
    return launch_stream<M>(h, mnmf_xt_kernel<M>, p, 1, MN_SLAB, 8, false);
}

template <int M>
int mn_update_basis(bss_handle* h) {
    MnParams p{};
    p.a = mn_args(h);
    p.out_f = h->basis2;
    const int n_kc = (int)cdiv(h->K, MN_KC);
    BSS_TRY((launch_stream<M>(h, mnmf_basis_kernel<M>, p, n_kc, MN_SLAB, 8)));
    float* t = h->basis;
    h->basis = h->basis2;
    h->basis2 = t;
    return BSS_OK;
}

template <int M>
int mn_update_act(bss_handle* h) {
    const MnArgs a = mn_args(h);
    const int n_kc = (int)cdiv(a.K, MN_KC);
    const int n_slabs = (a.Tp + 63) / 64;
    long long want = (long long)h->n_sm * 16;
    long long per_chunk_items = (long long)a.B * n_slabs * n_kc;
    int n_chunks = (int)cdiv(want, per_chunk_items);
    if (n_chunks < 1) n_chunks = 1;
    int bins_per_chunk = (int)cdiv(a.F, n_chunks);
    if (bins_per_chunk < 4) bins_per_chunk = a.F < 4 ? a.F : 4;
    n_chunks = (int)cdiv(a.F, bins_per_chunk);
    const size_t need = (size_t)a.B * n_chunks * a.N * a.K * 2 * a.Tp;
    if (need > h->part_elems) {
        if (h->part) cudaFree(h->part);
        h->part = nullptr;
        h->part_elems = 0;
        BSS_CUDA(h, cudaMalloc(&h->part, need * sizeof(float)));
        h->part_elems = need;
    }
    const long long n_items = per_chunk_items * n_chunks;
    const int wpc = 4;
    const uint32_t stride = (uint32_t)round_up((int)mn_scratch_bytes<M>(a.K), 16);
    BSS_CUDA(h, cudaFuncSetAttribute(mnmf_act_partial_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
    mnmf_act_partial_kernel<M><<<(unsigned)cdiv(n_items, wpc), wpc * 32, (size_t)wpc * stride, h->stream>>>(
        a, h->part, n_chunks, bins_per_chunk, n_slabs, n_kc, n_items, stride);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    const long long total = (long long)a.B * a.N * a.K * a.Tp;
    mnmf_act_finish_kernel<<<(unsigned)cdiv(total, 256), 256, 0, h->stream>>>(a, h->part, h->act, n_chunks);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

template <int M>
int mn_update_scm(bss_handle* h) {
    MnParams p{};
    p.a = mn_args(h);
    p.out_f = h->G2;
    if (getenv("BSSGPU_NO_MMA"))
        BSS_TRY((launch_stream<M>(h, mnmf_scm_kernel<M>, p, (int)cdiv(M, MN_MG), MN_SLAB, 8)));
    else
        BSS_TRY((launch_stream<M>(h, mnmf_scm_mma_kernel<M>, p, 1, MN_SLAB, 16)));
    float* t = h->G;
    h->G = h->G2;
    h->G2 = t;
    return BSS_OK;
}

template <int M>
int mn_weights(bss_handle* h, int tiled) {
    const MnArgs a = mn_args(h);
    const long long n_items = (long long)a.B * a.F * ((a.Tp + BSS_XSLAB - 1) / BSS_XSLAB);
    const int wpc = 4;
    const uint32_t stride = (uint32_t)round_up((int)mn_scratch_bytes<M>(a.K), 16);
    mnmf_weights_kernel<M><<<(unsigned)cdiv(n_items, wpc), wpc * 32, (size_t)wpc * stride, h->stream>>>(a, h->iw, n_items, tiled, stride);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

template <int M>
int mn_loss(bss_handle* h) {
    MnParams p{};
    p.a = mn_args(h);
    p.out_d = h->lossbuf;
    return launch_stream<M>(h, mnmf_loss_kernel<M>, p, 1, MN_SLAB, 8);
}

template <int M>
int mn_separate(bss_handle* h, cf* out) {
    const long long n_bins = (long long)h->B * h->F;
    cf* qinv = reinterpret_cast<cf*>(h->scale);   // [B][N][F] double2 >= [B][F][M] float2 when N >= 1 ... sized in mnmf_allocate
    mnmf_qinv_kernel<M><<<(unsigned)cdiv(n_bins, 64), 64, 0, h->stream>>>(h->W, qinv, n_bins, h->cfg.reference_id, h->flags);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    MnParams p{};
    p.a = mn_args(h);
    p.out_c = out;
    p.qinv = qinv;
    return launch_stream<M>(h, mnmf_separate_kernel<M>, p, 1, MN_SLAB, 8, false);
}

}  // namespace

#define MN_DISPATCH(Mval, CALL)                                                        \
    switch (Mval) {                                                                    \
        case 2: { constexpr int MM_ = 2; CALL; } break;                                \
        case 3: { constexpr int MM_ = 3; CALL; } break;                                \
        case 4: { constexpr int MM_ = 4; CALL; } break;                                \
        case 5: { constexpr int MM_ = 5; CALL; } break;                                \
        case 6: { constexpr int MM_ = 6; CALL; } break;                                \
        case 7: { constexpr int MM_ = 7; CALL; } break;                                \
        case 8: { constexpr int MM_ = 8; CALL; } break;                                \
        default: return bss_fail(h, BSS_EINVAL, "n_channels must be between 2 and 8"); \
    }

int launch_mnmf_xt(bss_handle* h) {
    int rc = BSS_OK;
    MN_DISPATCH(h->C, (rc = mn_xt<MM_>(h)))
    return rc;
}
int launch_mnmf_basis(bss_handle* h) {
    int rc = BSS_OK;
    MN_DISPATCH(h->C, (rc = mn_update_basis<MM_>(h)))
    return rc;
}
int launch_mnmf_act(bss_handle* h) {
    int rc = BSS_OK;
    MN_DISPATCH(h->C, (rc = mn_update_act<MM_>(h)))
    return rc;
}
int launch_mnmf_scm(bss_handle* h) {
    int rc = BSS_OK;
    MN_DISPATCH(h->C, (rc = mn_update_scm<MM_>(h)))
    return rc;
}
int launch_mnmf_weights(bss_handle* h, int tiled) {
    int rc = BSS_OK;
    MN_DISPATCH(h->C, (rc = mn_weights<MM_>(h, tiled)))
    return rc;
}
int launch_mnmf_loss_terms(bss_handle* h) {
    int rc = BSS_OK;
    MN_DISPATCH(h->C, (rc = mn_loss<MM_>(h)))
    return rc;
}
int launch_mnmf_separate(bss_handle* h, cf* out) {
    int rc = BSS_OK;
    MN_DISPATCH(h->C, (rc = mn_separate<MM_>(h, out)))
    return rc;
}
int launch_mnmf_normalize(bss_handle* h) {
    const long long n_bins = (long long)h->B * h->F;
    mnmf_norm_bin_kernel<<<(unsigned)cdiv(n_bins, 4), 128, 0, h->stream>>>(h->W, h->Wf, h->G, h->basis, h->B, h->N, h->F, h->C, h->K,
                                                                            h->cfg.eps);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    mnmf_norm_basis_kernel<<<(unsigned)((long long)h->B * h->N * h->K), 256, 0, h->stream>>>(h->basis, h->act, h->N, h->F, h->K, h->Tp,
                                                                                            h->cfg.eps);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}
