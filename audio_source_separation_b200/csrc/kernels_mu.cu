// Multiplicative updates of the NMF source model of (t-)ILRMA:
//   basis T      : per-bin reductions over frames                    src/bss/ilrma.py:413-419, :915-927
//   activation V : reductions across ALL bins, two deterministic stages   src/bss/ilrma.py:422-428, :929-938
// The estimates Y = W X are never stored: their power is recomputed from the staged bin tile.
// Arithmetic is paired fp32 (FFMA2): a complex multiply-add is two instructions and the two frames a
// lane handles per step share every weight computation.
#include <algorithm>
#include <cmath>

#include "handle.h"

namespace {

constexpr int MU_STAGES = 3;
constexpr int MU_SLAB = 128;
constexpr int MU_MAX_WARPS = 16;

__device__ __forceinline__ float2 rcp2(float2 v) { return make_float2(rcp_fast(v.x), rcp_fast(v.y)); }

// statistics of the majorisation for a pair of frames: a multiplies the numerator, b = 1/TV the denominator
__device__ __forceinline__ void mu_stats2(int mode, float2 P, float2 tv, float p_exp, float nu, float2& a, float2& b) {
    b = rcp2(tv);
    if (mode == 0) {
        if (p_exp == 2.f) {
            a = __fmul2_rn(P, __fmul2_rn(b, b));   // P / TV^2
        } else {
            a = make_float2(P.x / powf(tv.x, p_exp), P.y / powf(tv.y, p_exp));
        }
    } else {
        // Student-t: h / TV^2 with h = 1 / (2/((2+nu) TV) + nu/((2+nu) P))    src/bss/ilrma.py:922
        const float c = 2.f + nu;
        const float hx = 1.f / (2.f / (c * tv.x) + nu / (c * P.x));
        const float hy = 1.f / (2.f / (c * tv.y) + nu / (c * P.y));
        a = __fmul2_rn(make_float2(hx, hy), __fmul2_rn(b, b));
    }
}

__device__ __forceinline__ float pow_q(float r, float q) { return q == 0.5f ? sqrtf(r) : powf(r, q); }

template <int C, bool FROM_Y>
__device__ __forceinline__ void load_filter(float2 (&w)[C][C], const cf* Wf) {
    if (!FROM_Y) {
#pragma unroll
        for (int n = 0; n < C; ++n)
#pragma unroll
            for (int c = 0; c < C; ++c) w[n][c] = __ldg(Wf + n * C + c);
    }
}

// P[n] = (|y_n(t0)|^2, |y_n(t1)|^2) for the two frames held in xv (float4 = 2 complex frames per row)
template <int C, bool FROM_Y>
__device__ __forceinline__ void frame_power2(const float4 (&xv)[C], const float2 (&w)[C][C], float2 (&P)[C]) {
#pragma unroll
    for (int n = 0; n < C; ++n) {
        float2 y0, y1;
        if (FROM_Y) {
            y0 = make_float2(xv[n].x, xv[n].y);
            y1 = make_float2(xv[n].z, xv[n].w);
        } else {
            y0 = make_float2(0.f, 0.f);
            y1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                // w x = x * w.x + (-x.y, x.x) * w.y
                const float2 x0 = make_float2(xv[c].x, xv[c].y), x1 = make_float2(xv[c].z, xv[c].w);
                const float2 wx = make_float2(w[n][c].x, w[n][c].x), wy = make_float2(w[n][c].y, w[n][c].y);
                y0 = __ffma2_rn(x0, wx, y0);
                y0 = __ffma2_rn(make_float2(-x0.y, x0.x), wy, y0);
                y1 = __ffma2_rn(x1, wx, y1);
                y1 = __ffma2_rn(make_float2(-x1.y, x1.x), wy, y1);
            }
        }
        const float2 s0 = __fmul2_rn(y0, y0), s1 = __fmul2_rn(y1, y1);
        P[n] = make_float2(s0.x + s0.y, s1.x + s1.y);
    }
}

// ------------------------------------------------------------------------------------------- basis
struct MuParams {
    MuArgs a;
    TileGeom g;
    long long n_items;
    int n_kc;
    uint32_t scratch_off, scratch_stride, ring_off;
    uint32_t cache_off;   // CACHE: shared-memory copy of the activation rows of the mixtures this CTA touches
};

// KFIX: n_basis == KC at compile time (basis row in registers, one item per bin);
// otherwise n_basis is a run-time value, an item is a (bin, chunk of KC basis vectors) pair.
// CACHE: CTA-contiguous bin ranges with the activation rows in shared memory (see cov_kernel).
// WP: also store the source power |y|^2 of every frame as float bin tiles (a.Pout) for the activation update.
// GAUSS2 (only together with WP): mode 0 with p_exp == 2 decided at compile time, see mu_act_stream_kernel.
template <int C, int KC, bool KFIX, bool FROM_Y, bool CACHE, bool WP, bool GAUSS2 = false>
__global__ void __launch_bounds__(MU_MAX_WARPS * 32, 1) mu_basis_kernel(const MuParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const MuArgs& a = p.a;
    constexpr int N = C;
    constexpr int M = N * KC * 2;
    constexpr int MP = (M + 31) / 32 * 32;
    constexpr int Q = MP / 32;
    const int K = KFIX ? KC : a.K;
    const int Tp = a.Tp;

    float* tb = reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride);   // [N][K] (run-time K only)
    float* red = tb + (KFIX ? 0 : N * K);                                                           // [MP]
    int lo, hi;
    cta_item_range((int)p.n_items, lo, hi);
    // without the shared-memory cache the activation rows are read through L1: a two-stage ring leaves L1 room for them
    constexpr int STG = CACHE ? MU_STAGES : 2;
    WarpStream<STG> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * STG,
             smem + p.ring_off + (size_t)warp * STG * p.g.stage_bytes, FROM_Y ? a.Y : a.X, lo + warp, wpc, hi, p.n_kc, lane);
    const float* vcache = reinterpret_cast<const float*>(smem + p.cache_off);
    int b_lo = 0;
    if (CACHE) b_lo = load_act_cache(reinterpret_cast<float*>(smem + p.cache_off), a.act, N * K * Tp, lo, hi, p.n_kc, a.F);
    int voff = 0;

    float2 num[N][KC], den[N][KC];
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) num[n][kk] = den[n][kk] = make_float2(0.f, 0.f);
    float2 w[C][C];
    float tk[N][KC];
    int b = 0, f = 0, k0 = 0;
    const float* vrow = nullptr;

#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, (FROM_Y ? a.Y : a.X), p.n_kc);
        if (st.first_slab()) {
            const int bf = st.cons.item / p.n_kc;
            k0 = (st.cons.item - bf * p.n_kc) * KC;
            b = bf / a.F;
            f = bf - b * a.F;
            vrow = a.act + (size_t)b * N * K * Tp;
            voff = (b - b_lo) * N * K * Tp;
            if (KFIX) {
#pragma unroll
                for (int n = 0; n < N; ++n)
#pragma unroll
                    for (int kk = 0; kk < KC; ++kk) tk[n][kk] = __ldg(a.basis + (((size_t)b * N + n) * a.F + f) * KC + kk);
            } else {
                for (int i = lane; i < N * K; i += 32) {
                    const int n = i / K, k = i - n * K;
                    tb[i] = a.basis[(((size_t)b * N + n) * a.F + f) * K + k];
                }
                __syncwarp();
            }
            load_filter<C, FROM_Y>(w, a.Wf + (size_t)bf * C * C);
        }

        const cf* xs = st.acquire(p.g);
        const int nf = st.frames(p.g);
        const int tbase = st.frame0(p.g);
#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float4 xv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * nf + tt);
            float2 P[C];
            frame_power2<C, FROM_Y>(xv, w, P);
            const int t = tbase + tt;
            if (WP) {
                // block-major power tiles [b][128-frame block][f][N rows of nf floats] (stride N * 128 floats per bin): for a
                // fixed block the bins are consecutive, so the activation kernel fetches several bins with one bulk copy
                float* pd = a.Pout + ((((size_t)b * ((Tp + MU_SLAB - 1) / MU_SLAB) + tbase / MU_SLAB) * a.F + f) * N * MU_SLAB + tt);
#pragma unroll
                for (int n = 0; n < N; ++n) *reinterpret_cast<float2*>(pd + n * nf) = P[n];
            }
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float* v = CACHE ? vcache + voff + n * K * Tp + t : vrow + (size_t)n * K * Tp + t;
                float2 tv = make_float2(0.f, 0.f);
                float2 vk[KC];
                if (KFIX) {
#pragma unroll
                    for (int kk = 0; kk < KC; ++kk) {
                        vk[kk] = CACHE ? *reinterpret_cast<const float2*>(v + kk * Tp) : __ldg(reinterpret_cast<const float2*>(v + (size_t)kk * Tp));
                        tv = __ffma2_rn(vk[kk], make_float2(tk[n][kk], tk[n][kk]), tv);
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < KC; ++kk) vk[kk] = make_float2(0.f, 0.f);
                    for (int k = 0; k < K; ++k) {
                        const float2 vv = CACHE ? *reinterpret_cast<const float2*>(v + k * Tp) : __ldg(reinterpret_cast<const float2*>(v + (size_t)k * Tp));
                        const float tkv = tb[n * K + k];
                        tv = __ffma2_rn(vv, make_float2(tkv, tkv), tv);
#pragma unroll
                        for (int kk = 0; kk < KC; ++kk)
                            if (k == k0 + kk) vk[kk] = vv;
                    }
                }
                tv.x = fmaxf(tv.x, a.eps);
                tv.y = fmaxf(tv.y, a.eps);
                float2 sa, sb;
                if (GAUSS2) {
                    sb = rcp2(tv);
                    sa = __fmul2_rn(P[n], __fmul2_rn(sb, sb));   // the p_exp == 2 branch of mu_stats2
                } else {
                    mu_stats2(a.mode, P[n], tv, a.p_exp, a.nu, sa, sb);
                }
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    num[n][kk] = __ffma2_rn(sa, vk[kk], num[n][kk]);
                    den[n][kk] = __ffma2_rn(sb, vk[kk], den[n][kk]);
                }
            }
        }

        if (st.last_slab(p.g)) {
            float flat[MP];
#pragma unroll
            for (int i = 0; i < MP; ++i) flat[i] = 0.f;
#pragma unroll
            for (int n = 0; n < N; ++n)
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    flat[(n * KC + kk) * 2] = num[n][kk].x + num[n][kk].y;
                    flat[(n * KC + kk) * 2 + 1] = den[n][kk].x + den[n][kk].y;
                    num[n][kk] = den[n][kk] = make_float2(0.f, 0.f);
                }
            warp_reduce_scatter<MP>(flat, lane);
#pragma unroll
            for (int q = 0; q < Q; ++q) red[Q * lane + q] = flat[q];
            __syncwarp();
            for (int i = lane; i < N * KC; i += 32) {
                const int n = i / KC, kk = i - n * KC, k = k0 + kk;
                if (k < K) {
                    const float nm = red[2 * i];
                    const float dn = fmaxf(red[2 * i + 1], a.eps);
                    const size_t idx = (((size_t)b * N + n) * a.F + f) * K + k;
                    if (a.raw) {   // partitioned model: the caller combines the raw statistics across sources / bins
                        a.raw[2 * idx] = nm;
                        a.raw[2 * idx + 1] = red[2 * i + 1];
                        continue;
                    }
                    const float told = a.basis[idx];
                    const bool sel = a.sel_m < 0 || n == a.sel_m || n == a.sel_n;
                    a.basis_out[idx] = sel ? told * pow_q(nm / dn, a.q_exp) : told;
                }
            }
        }
        st.release(p.g);
    }
}

template <int C, int KC, bool KFIX, bool FROM_Y, bool CACHE, bool WP = false, bool GAUSS2 = false>
int launch_mu_basis_c(bss_handle* h, const MuParams& p, const StreamPlan& sp, size_t smem_bytes) {
    if constexpr (!WP && KFIX && !FROM_Y) {
        if (p.a.Pout) {
            if (p.a.mode == 0 && p.a.p_exp == 2.f) return launch_mu_basis_c<C, KC, KFIX, FROM_Y, CACHE, true, true>(h, p, sp, smem_bytes);
            return launch_mu_basis_c<C, KC, KFIX, FROM_Y, CACHE, true, false>(h, p, sp, smem_bytes);
        }
    }
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(mu_basis_kernel<C, KC, KFIX, FROM_Y, CACHE, WP, GAUSS2>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    mu_basis_kernel<C, KC, KFIX, FROM_Y, CACHE, WP, GAUSS2><<<sp.grid, sp.wpc * 32, smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

template <int C, int KC, bool KFIX, bool FROM_Y>
int launch_mu_basis_t(bss_handle* h, const MuArgs& a) {
    MuParams p;
    p.a = a;
    p.g = make_tile_geom(C, a.Tp, MU_SLAB);
    p.n_kc = KFIX ? 1 : (a.K + KC - 1) / KC;
    p.n_items = (long long)a.B * a.F * p.n_kc;
    p.cache_off = 0;
    constexpr int MP = (C * KC * 2 + 31) / 32 * 32;
    const size_t scratch = ((KFIX ? 0 : (size_t)C * a.K) + MP) * 4;   // [N][K] basis row (run-time K only) + reduction buffer
    StreamPlan sp;
    size_t smem_bytes = 0;
    const bool cached = plan_stream_cached(h, p.g, MU_STAGES, scratch, p.n_items, MU_MAX_WARPS, (size_t)C * a.K * a.Tp * sizeof(float),
                                           (long long)a.F * p.n_kc, &sp, &p.cache_off, &smem_bytes);
    if (!cached && !plan_stream(h, p.g, 2, scratch, p.n_items, MU_MAX_WARPS, &sp))
        return bss_fail(h, BSS_EINVAL, "source model: frame tile does not fit in shared memory");
    p.scratch_off = sp.scratch_off;
    p.scratch_stride = sp.scratch_stride;
    p.ring_off = sp.ring_off;
    if (cached) return launch_mu_basis_c<C, KC, KFIX, FROM_Y, true>(h, p, sp, smem_bytes);
    return launch_mu_basis_c<C, KC, KFIX, FROM_Y, false>(h, p, sp, sp.smem_bytes);
}

// ------------------------------------------------------------------------------------------- activation
// Stage 1: a warp owns 64 frames (two per lane) of one mixture and walks a chunk of bins, reading
// the 512-byte row segments straight from global memory; partial sums go to `part`.
// part layout: [B][n_chunks][N][K][2][Tp]
template <int C, int KC, bool KFIX, bool FROM_Y>
__global__ void __launch_bounds__(128) mu_act_partial_kernel(const MuArgs a, float* part, int n_chunks, int bins_per_chunk,
                                                            int n_slabs, int n_kc, long long n_items) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const long long item = (int)(blockIdx.x * wpc + warp);
    if (item >= n_items) return;
    constexpr int N = C;
    const int K = KFIX ? KC : a.K;
    // item -> (b, chunk, slab, kc), kc fastest so that the k-chunks of a tile run side by side
    long long r = item;
    const int kc = (int)(r % n_kc);
    r /= n_kc;
    const int slab = (int)(r % n_slabs);
    r /= n_slabs;
    const int chunk = (int)(r % n_chunks);
    const int b = (int)(r / n_chunks);
    const int k0 = kc * KC;
    const int t0 = slab * 64 + 2 * lane;
    const bool live = t0 < a.Tp;

    // activation values of this lane's two frames: registers when K is fixed, shared memory otherwise
    float2 vreg[N][KC];
    float* vs = reinterpret_cast<float*>(smem) + (size_t)warp * N * K * 64;
    if (KFIX) {
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk)
                vreg[n][kk] = live ? __ldg(reinterpret_cast<const float2*>(a.act + (((size_t)b * N + n) * KC + kk) * a.Tp + t0))
                                   : make_float2(0.f, 0.f);
    } else {
        for (int i = 0; i < N * K; ++i) {
            float2 vv = make_float2(0.f, 0.f);
            if (live) vv = __ldg(reinterpret_cast<const float2*>(a.act + ((size_t)b * N * K + i) * a.Tp + t0));
            *reinterpret_cast<float2*>(vs + (size_t)i * 64 + 2 * lane) = vv;
        }
        __syncwarp();
    }

    float2 num[N][KC], den[N][KC];
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) num[n][kk] = den[n][kk] = make_float2(0.f, 0.f);

    const int f_begin = chunk * bins_per_chunk;
    const int f_end = min(a.F, f_begin + bins_per_chunk);
    const cf* src = FROM_Y ? a.Y : a.X;
    // this lane's two frames inside the block-interleaved bin tile (common.cuh: tile_off)
    const size_t xoff = tile_off(C, a.Tp, 0, live ? t0 : 0);
    const int xlen = (int)(tile_off(C, a.Tp, 1, live ? t0 : 0) - xoff);
#pragma unroll 2
    for (int f = f_begin; f < f_end; ++f) {
        const size_t bf = (size_t)b * a.F + f;
        float4 xv[C];
#pragma unroll
        for (int c = 0; c < C; ++c)
            xv[c] = live ? __ldg(reinterpret_cast<const float4*>(src + bf * C * a.Tp + xoff + (size_t)c * xlen)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float2 w[C][C];
        load_filter<C, FROM_Y>(w, a.Wf + bf * C * C);
        float2 P[C];
        frame_power2<C, FROM_Y>(xv, w, P);
#pragma unroll
        for (int n = 0; n < N; ++n) {
            const float* tbn = a.basis + (((size_t)b * N + n) * a.F + f) * K;
            float2 tv = make_float2(0.f, 0.f);
            float tk[KC];
            if (KFIX) {
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    tk[kk] = __ldg(tbn + kk);
                    tv = __ffma2_rn(vreg[n][kk], make_float2(tk[kk], tk[kk]), tv);
                }
            } else {
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) tk[kk] = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float tkv = __ldg(tbn + k);
                    const float2 vv = *reinterpret_cast<const float2*>(vs + ((size_t)n * K + k) * 64 + 2 * lane);
                    tv = __ffma2_rn(vv, make_float2(tkv, tkv), tv);
#pragma unroll
                    for (int kk = 0; kk < KC; ++kk)
                        if (k == k0 + kk) tk[kk] = tkv;
                }
            }
            tv.x = fmaxf(tv.x, a.eps);
            tv.y = fmaxf(tv.y, a.eps);
            float2 sa, sb;
            mu_stats2(a.mode, P[n], tv, a.p_exp, a.nu, sa, sb);
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const float2 t2 = make_float2(tk[kk], tk[kk]);
                num[n][kk] = __ffma2_rn(sa, t2, num[n][kk]);
                den[n][kk] = __ffma2_rn(sb, t2, den[n][kk]);
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const int k = k0 + kk;
            if (k < K) {
                float* dst = part + (((((size_t)b * n_chunks + chunk) * N + n) * K + k) * 2) * a.Tp + t0;
                *reinterpret_cast<float2*>(dst) = num[n][kk];
                *reinterpret_cast<float2*>(dst + a.Tp) = den[n][kk];
            }
        }
}

// Stage 2: fixed-order sum over the chunks (deterministic), then V <- V (num/den)^q, in place.
// Block = 32 frames x 8 chunk lanes: lane y sums chunks y, y+8, ..., then the 8 partial sums are added in order.
__global__ void __launch_bounds__(256) mu_act_finish_kernel(const MuArgs a, const float* part, float* act, int N, int n_chunks) {
    __shared__ float red[2][8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int t_tiles = (a.Tp + 31) / 32;
    const long long row = blockIdx.x / t_tiles;          // (b, n, k)
    const int t = (int)(blockIdx.x % t_tiles) * 32 + tx;
    const int k = (int)(row % a.K);
    const long long bn = row / a.K;
    const int n = (int)(bn % N);
    const int b = (int)(bn / N);
    float num = 0.f, den = 0.f;
    if (t < a.Tp) {
        for (int c = ty; c < n_chunks; c += 8) {
            const float* src = part + (((((size_t)b * n_chunks + c) * N + n) * a.K + k) * 2) * a.Tp + t;
            num += src[0];
            den += src[a.Tp];
        }
    }
    red[0][ty][tx] = num;
    red[1][ty][tx] = den;
    __syncthreads();
    if (ty != 0 || t >= a.Tp) return;
    const size_t idx = (size_t)row * a.Tp + t;
    if (t >= a.T) {
        act[idx] = 0.f;
        return;
    }
    const bool sel = a.sel_m < 0 || n == a.sel_m || n == a.sel_n;
    if (!sel) return;
    num = den = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
        num += red[0][y][tx];
        den += red[1][y][tx];
    }
    den = fmaxf(den, a.eps);
    act[idx] = act[idx] * pow_q(num / den, a.q_exp);
}

// ---- streamed stage 1 (n_basis fixed at compile time) ------------------------------------------------------
// A warp owns ONE 128-frame block index s of one chunk of bins and streams that block of every bin of the chunk
// through a private ring: each stage holds the block (one bulk copy, 4 KB at C = 4) followed by the bin's packed
// parameters (demixing filter rows and basis values, one more bulk copy), so the loop body touches only shared
// memory and registers.  The lane's activation values are loop invariants (registers), as are its accumulators.
constexpr int ACT_STAGES = 4;
constexpr int ACT_BINS_P = 2;     // power tiles (FROM_P): bins per ring stage (one bulk copy).  More resident warps instead (four CTAs
                                  // per SM, six single-bin stages, 128 registers) measured slower: 299 vs 285 us (profiles/r5f_*)
constexpr int ACT_WARPS = 4;

struct ActParams {
    MuArgs a;
    float* part;                  // [B][n_chunks][N][K][2][Tp]
    const unsigned char* pbin;    // packed per-bin parameters, pb_stride bytes per bin: Wf [C][C] complex64 | T [N][K] float
    int pb_stride;
    int n_chunks, bins_per_chunk, n_blocks;
    int n_items;                  // B * n_chunks * n_blocks
    uint32_t stage_bytes, par_off;
};

__global__ void __launch_bounds__(256) pack_bin_params_kernel(const cf* Wf, const float* basis, unsigned char* out, int B, int N,
                                                              int C, int F, int K, int stride, int with_filter) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int words = stride >> 2;
    if (idx >= (long long)B * F * words) return;
    const int w = (int)(idx % words);
    const long long bf = idx / words;
    const int f = (int)(bf % F), b = (int)(bf / F);
    const int nw = with_filter ? C * C * 2 : 0;
    float v = 0.f;
    if (w < nw) {
        v = reinterpret_cast<const float*>(Wf + (size_t)bf * C * C)[w];
    } else if (w < nw + N * K) {
        const int i = w - nw, n = i / K, k = i - n * K;
        v = basis[(((size_t)b * N + n) * F + f) * K + k];
    }
    reinterpret_cast<float*>(out)[idx] = v;
}

// FROM_P: the ring carries the float power tiles the basis kernel stored (a.Pin) instead of the mixture: no filter in the
// packed parameters, no y = W x, half the bytes.
// GAUSS2: Gauss source model with domain 2 (mode 0, p_exp == 2) decided at compile time: the run-time dispatch of mu_stats2,
// inlined once per source and frame pair with its powf / Student-t branches, tripled the code of the loop body and showed up
// as instruction-fetch and branch stalls in the (latency bound) power-tile form.
template <int C, int KC, bool FROM_Y, bool FROM_P, bool GAUSS2>
__global__ void __launch_bounds__(ACT_WARPS * 32, (C <= 4 ? 3 : 1)) mu_act_stream_kernel(const ActParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const MuArgs& a = p.a;
    constexpr int N = C;
    constexpr int STG = ACT_STAGES;
    // bins per ring stage.  A 2 KB power block per bulk copy is too small: the per-copy cost of the copy engine, not DRAM,
    // bounded the kernel (4.0 TB/s); the power tiles are laid out block-major so that one copy brings two bins.
    constexpr int G = FROM_P ? ACT_BINS_P : 1;
    const int item = (int)blockIdx.x * ACT_WARPS + warp;
    if (item >= p.n_items) return;
    // item -> (b, chunk, block), block fastest: the warps of a CTA read consecutive blocks of the same bins
    const int s = item % p.n_blocks;
    const int bc = item / p.n_blocks;
    const int chunk = bc % p.n_chunks;
    const int b = bc / p.n_chunks;
    const int f_begin = chunk * p.bins_per_chunk;
    const int f_end = min(a.F, f_begin + p.bins_per_chunk);
    const int Tp = a.Tp;
    const int blk0 = s * BSS_XSLAB;
    const int L = min(BSS_XSLAB, Tp - blk0);          // frames of this block (even)
    constexpr int ESZ = FROM_P ? 4 : 8;               // bytes per tile element
    const uint32_t blk_bytes = (uint32_t)(C * L * ESZ);
    // distance between the blocks of consecutive bins, and where bin f_begin's block starts
    const size_t bin_bytes = FROM_P ? (size_t)C * BSS_XSLAB * ESZ : (size_t)C * Tp * ESZ;
    const unsigned char* src0 =
        FROM_P ? reinterpret_cast<const unsigned char*>(a.Pin) + ((size_t)b * p.n_blocks + s) * a.F * bin_bytes
               : reinterpret_cast<const unsigned char*>(FROM_Y ? a.Y : a.X) + ((size_t)b * a.F * C * Tp + (size_t)blk0 * C) * ESZ;
    const unsigned char* par0 = p.pbin + (size_t)b * a.F * p.pb_stride;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * STG;
    unsigned char* ring = smem + ACT_WARPS * STG * 8 + (size_t)warp * STG * p.stage_bytes;   // barriers first
    const uint32_t bars_sa = smem_u32(bars), ring_sa = smem_u32(ring);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STG; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncwarp();
    // one stage = the blocks of up to G consecutive bins (bin g at g * bin_bytes) followed by their packed parameters
    auto issue = [&](int f, int stage) {
        if (lane == 0) {
            const int g = min(G, f_end - f);
            const uint32_t bar = bars_sa + 8u * (uint32_t)stage;
            const uint32_t dst = ring_sa + (uint32_t)stage * p.stage_bytes;
            const uint32_t tile_bytes = (uint32_t)(g - 1) * (uint32_t)bin_bytes + blk_bytes;
            const uint32_t par_bytes = (uint32_t)g * (uint32_t)p.pb_stride;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tile_bytes + par_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(src0 + (size_t)f * bin_bytes), "r"(tile_bytes), "r"(bar)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + p.par_off),
                         "l"(par0 + (size_t)f * p.pb_stride), "r"(par_bytes), "r"(bar)
                         : "memory");
        }
    };
    int fp = f_begin, pstage = 0;
#pragma unroll 1
    for (int i = 0; i < STG - 1 && fp < f_end; ++i, fp += G) {
        issue(fp, pstage);
        pstage = pstage + 1 == STG ? 0 : pstage + 1;
    }

    // loop invariants of this lane: activation values of its frame pairs (two pairs per 128-frame block)
    float2 vreg[2][N][KC];
    float2 num[2][N][KC], den[2][N][KC];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int tt = 2 * lane + 64 * j;
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                vreg[j][n][kk] = tt < L ? __ldg(reinterpret_cast<const float2*>(a.act + (((size_t)b * N + n) * KC + kk) * Tp + blk0 + tt))
                                        : make_float2(0.f, 0.f);
                num[j][n][kk] = den[j][n][kk] = make_float2(0.f, 0.f);
            }
    }

    int cstage = 0;
    uint32_t cphase = 0;
#pragma unroll 1
    for (int f = f_begin; f < f_end; f += G) {
        if (fp < f_end) {
            issue(fp, pstage);
            fp += G;
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
        const unsigned char* stage0 = ring + (size_t)cstage * p.stage_bytes;
        const int g_n = G == 1 ? 1 : min(G, f_end - f);
#pragma unroll 1
        for (int g = 0; g < g_n; ++g) {
            const unsigned char* stage = stage0 + (size_t)g * bin_bytes;
            const unsigned char* par = stage0 + p.par_off + g * p.pb_stride;
            const cf* xs = reinterpret_cast<const cf*>(stage);
            const float2* wf = reinterpret_cast<const float2*>(par);
            const float* tb = reinterpret_cast<const float*>(par) + (FROM_Y || FROM_P ? 0 : C * C * 2);
            float2 w[C][C];
            if (!FROM_Y && !FROM_P) {
#pragma unroll
                for (int n = 0; n < C; ++n)
#pragma unroll
                    for (int c = 0; c < C; ++c) w[n][c] = wf[n * C + c];
            }
            float tk[N][KC];
#pragma unroll
            for (int n = 0; n < N; ++n)
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) tk[n][kk] = tb[n * KC + kk];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int tt = 2 * lane + 64 * j;
                if (tt < L) {
                    float2 P[C];
                    if (FROM_P) {
#pragma unroll
                        for (int n = 0; n < N; ++n) P[n] = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(stage) + n * L + tt);
                    } else {
                        float4 xv[C];
#pragma unroll
                        for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * L + tt);
                        frame_power2<C, FROM_Y>(xv, w, P);
                    }
#pragma unroll
                    for (int n = 0; n < N; ++n) {
                        float2 tv = make_float2(0.f, 0.f);
#pragma unroll
                        for (int kk = 0; kk < KC; ++kk) tv = __ffma2_rn(vreg[j][n][kk], make_float2(tk[n][kk], tk[n][kk]), tv);
                        tv.x = fmaxf(tv.x, a.eps);
                        tv.y = fmaxf(tv.y, a.eps);
                        float2 sa, sb;
                        if (GAUSS2) {
                            sb = rcp2(tv);
                            sa = __fmul2_rn(P[n], __fmul2_rn(sb, sb));   // the p_exp == 2 branch of mu_stats2
                        } else {
                            mu_stats2(a.mode, P[n], tv, a.p_exp, a.nu, sa, sb);
                        }
#pragma unroll
                        for (int kk = 0; kk < KC; ++kk) {
                            const float2 t2 = make_float2(tk[n][kk], tk[n][kk]);
                            num[j][n][kk] = __ffma2_rn(sa, t2, num[j][n][kk]);
                            den[j][n][kk] = __ffma2_rn(sb, t2, den[j][n][kk]);
                        }
                    }
                }
            }
        }
        __syncwarp();
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
                    *reinterpret_cast<float2*>(dst) = num[j][n][kk];
                    *reinterpret_cast<float2*>(dst + Tp) = den[j][n][kk];
                }
        }
    }
}

// host: choose the number of bin chunks so that the warps fill the machine in whole waves
template <int C, int KC, bool FROM_Y, bool FROM_P = false, bool GAUSS2 = false>
int launch_mu_act_stream(bss_handle* h, const MuArgs& a, int* n_chunks_out, bool* done) {
    if constexpr (!FROM_P && !FROM_Y) {
        if (a.Pin) {
            if (a.mode == 0 && a.p_exp == 2.f) return launch_mu_act_stream<C, KC, FROM_Y, true, true>(h, a, n_chunks_out, done);
            return launch_mu_act_stream<C, KC, FROM_Y, true, false>(h, a, n_chunks_out, done);
        }
    }
    *done = false;
    ActParams p{};
    p.a = a;
    p.n_blocks = (a.Tp + BSS_XSLAB - 1) / BSS_XSLAB;
    const int blk_frames = a.Tp < BSS_XSLAB ? a.Tp : BSS_XSLAB;
    p.pb_stride = round_up((FROM_Y || FROM_P ? 0 : C * C * 8) + C * KC * 4, 16);
    constexpr int G = FROM_P ? ACT_BINS_P : 1;
    // FROM_P: G bins per stage at the fixed block-major stride (C * 128 floats), then their G parameter records
    p.par_off = FROM_P ? (uint32_t)(G * C * BSS_XSLAB * 4) : (uint32_t)round_up(C * blk_frames * 8, 16);
    p.stage_bytes = (uint32_t)round_up((int)p.par_off + G * p.pb_stride, 128);
    constexpr int STG = ACT_STAGES;
    const size_t smem_bytes = (size_t)ACT_WARPS * STG * 8 + (size_t)ACT_WARPS * STG * p.stage_bytes;
    if (smem_bytes > (size_t)h->max_smem) return BSS_OK;   // fall back to the direct-load kernel
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(mu_act_stream_kernel<C, KC, FROM_Y, FROM_P, GAUSS2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    int ctas_per_sm = 1;
    BSS_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, mu_act_stream_kernel<C, KC, FROM_Y, FROM_P, GAUSS2>, ACT_WARPS * 32, smem_bytes));
    if (ctas_per_sm < 1) return BSS_OK;
    const long long slots = (long long)h->n_sm * ctas_per_sm * ACT_WARPS;
    const long long per_chunk = (long long)a.B * p.n_blocks;
    // candidates: chunks of at least 16 bins, at most 4 waves; keep the most efficient (ties: fewer chunks)
    int best = 1;
    double best_eff = -1.0;
    const int c_max = (int)std::max<long long>(1, std::min<long long>(a.F / 16, cdiv(4 * slots, per_chunk)));
    for (int c = 1; c <= c_max; ++c) {
        const double waves = (double)(per_chunk * c) / (double)slots;
        const double eff = waves / std::ceil(waves);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = c;
        }
    }
    if (h->opt_act_chunks > 0) best = std::min(h->opt_act_chunks, a.F);   // BSS_OPT_ACT_CHUNKS: reproduce another batch's order
    p.bins_per_chunk = (int)cdiv(a.F, best);
    p.n_chunks = (int)cdiv(a.F, p.bins_per_chunk);
    const long long n_items = per_chunk * p.n_chunks;
    if (n_items > 0x7fffffffLL) return BSS_OK;
    p.n_items = (int)n_items;
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
    const long long words = (long long)a.B * a.F * (p.pb_stride >> 2);
    pack_bin_params_kernel<<<(unsigned)cdiv(words, 256), 256, 0, h->stream>>>(a.Wf, a.basis, (unsigned char*)h->staging, a.B, C, C, a.F,
                                                                             KC, p.pb_stride, FROM_Y || FROM_P ? 0 : 1);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    mu_act_stream_kernel<C, KC, FROM_Y, FROM_P, GAUSS2><<<(unsigned)cdiv(n_items, ACT_WARPS), ACT_WARPS * 32, smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    if (n_chunks_out) *n_chunks_out = p.n_chunks;
    h->last_act_chunks = p.n_chunks;
    *done = true;
    return BSS_OK;
}

// act == nullptr: stage 1 only (the partial sums stay in h->part, *n_chunks_out tells how many)
template <int C, int KC, bool KFIX, bool FROM_Y>
int launch_mu_act_t(bss_handle* h, const MuArgs& a, float* act, int* n_chunks_out) {
    if constexpr (KFIX) {
        bool done = false;
        int n_chunks_s = 0;
        BSS_TRY((launch_mu_act_stream<C, KC, FROM_Y>(h, a, &n_chunks_s, &done)));
        if (done) {
            if (n_chunks_out) *n_chunks_out = n_chunks_s;
            if (!act) return BSS_OK;
            const long long blocks = (long long)a.B * C * a.K * ((a.Tp + 31) / 32);
            mu_act_finish_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(a, h->part, act, C, n_chunks_s);
            h->launches++;
            BSS_CUDA(h, cudaGetLastError());
            return BSS_OK;
        }
    }
    const int n_kc = KFIX ? 1 : (a.K + KC - 1) / KC;
    const int n_slabs = (a.Tp + 63) / 64;
    // enough warps to fill the machine a few times over, but chunks of at least 4 bins
    long long want = (long long)h->n_sm * 32;
    long long per_chunk_items = (long long)a.B * n_slabs * n_kc;
    int n_chunks = (int)cdiv(want, per_chunk_items);
    if (n_chunks < 1) n_chunks = 1;
    if (h->opt_act_chunks > 0) n_chunks = std::min(h->opt_act_chunks, a.F);
    int bins_per_chunk = (int)cdiv(a.F, n_chunks);
    if (bins_per_chunk < 4 && h->opt_act_chunks <= 0) bins_per_chunk = a.F < 4 ? a.F : 4;
    n_chunks = (int)cdiv(a.F, bins_per_chunk);
    h->last_act_chunks = n_chunks;
    const size_t need = (size_t)a.B * n_chunks * C * a.K * 2 * a.Tp;
    if (need > h->part_elems) {
        if (h->part) cudaFree(h->part);
        h->part = nullptr;
        h->part_elems = 0;
        BSS_CUDA(h, cudaMalloc(&h->part, need * sizeof(float)));
        h->part_elems = need;
    }
    const long long n_items = per_chunk_items * n_chunks;
    const int wpc = 4;
    const size_t smem_bytes = KFIX ? 0 : (size_t)wpc * C * a.K * 64 * sizeof(float);
    if (smem_bytes > (size_t)h->max_smem) return bss_fail(h, BSS_EINVAL, "source model: n_basis too large");
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(mu_act_partial_kernel<C, KC, KFIX, FROM_Y>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    mu_act_partial_kernel<C, KC, KFIX, FROM_Y><<<(unsigned)cdiv(n_items, wpc), wpc * 32, smem_bytes, h->stream>>>(
        a, h->part, n_chunks, bins_per_chunk, n_slabs, n_kc, n_items);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    if (n_chunks_out) *n_chunks_out = n_chunks;
    if (!act) return BSS_OK;
    const long long blocks = (long long)a.B * C * a.K * ((a.Tp + 31) / 32);
    mu_act_finish_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(a, h->part, act, C, n_chunks);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

}  // namespace

#define BSS_DISPATCH_C(Cval, CALL)                                                     \
    switch (Cval) {                                                                    \
        case 2: { constexpr int CC_ = 2; CALL; } break;                                \
        case 3: { constexpr int CC_ = 3; CALL; } break;                                \
        case 4: { constexpr int CC_ = 4; CALL; } break;                                \
        case 5: { constexpr int CC_ = 5; CALL; } break;                                \
        case 6: { constexpr int CC_ = 6; CALL; } break;                                \
        case 7: { constexpr int CC_ = 7; CALL; } break;                                \
        case 8: { constexpr int CC_ = 8; CALL; } break;                                \
        default: return bss_fail(h, BSS_EINVAL, "n_channels must be between 2 and 8"); \
    }

int launch_mu_basis(bss_handle* h, const MuArgs& a) {
    int rc = BSS_OK;
    const bool from_y = a.Y != nullptr;
    if (a.K == 2) {
        if (from_y) { BSS_DISPATCH_C(a.C, (rc = launch_mu_basis_t<CC_, 2, true, true>(h, a))) }
        else { BSS_DISPATCH_C(a.C, (rc = launch_mu_basis_t<CC_, 2, true, false>(h, a))) }
    } else {
        if (from_y) { BSS_DISPATCH_C(a.C, (rc = launch_mu_basis_t<CC_, 4, false, true>(h, a))) }
        else { BSS_DISPATCH_C(a.C, (rc = launch_mu_basis_t<CC_, 4, false, false>(h, a))) }
    }
    return rc;
}

// second stage alone: the partial sums of `n_chunks` bin chunks already sit in h->part (launch_mu_fused)
int launch_mu_act_finish(bss_handle* h, const MuArgs& a, float* act, int n_chunks) {
    const long long blocks = (long long)a.B * a.C * a.K * ((a.Tp + 31) / 32);
    mu_act_finish_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(a, h->part, act, a.C, n_chunks);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_mu_act(bss_handle* h, const MuArgs& a, float* act, int* n_chunks_out) {
    int rc = BSS_OK;
    const bool from_y = a.Y != nullptr;
    if (a.K == 2) {
        if (from_y) { BSS_DISPATCH_C(a.C, (rc = launch_mu_act_t<CC_, 2, true, true>(h, a, act, n_chunks_out))) }
        else { BSS_DISPATCH_C(a.C, (rc = launch_mu_act_t<CC_, 2, true, false>(h, a, act, n_chunks_out))) }
    } else {
        if (from_y) { BSS_DISPATCH_C(a.C, (rc = launch_mu_act_t<CC_, 4, false, true>(h, a, act, n_chunks_out))) }
        else { BSS_DISPATCH_C(a.C, (rc = launch_mu_act_t<CC_, 4, false, false>(h, a, act, n_chunks_out))) }
    }
    return rc;
}
