// Multiplicative updates of the NMF source model of (t-)ILRMA:
//   basis T      : per-bin reductions over frames                    src/bss/ilrma.py:413-419, :915-927
//   activation V : reductions across ALL bins, two deterministic stages   src/bss/ilrma.py:422-428, :929-938
// The estimates Y = W X are never stored: their power is recomputed from the staged bin tile.
// Arithmetic is paired fp32 (FFMA2): a complex multiply-add is two instructions and the two frames a
// lane handles per step share every weight computation.
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
};

// KFIX: n_basis == KC at compile time (basis row in registers, one item per bin);
// otherwise n_basis is a run-time value, an item is a (bin, chunk of KC basis vectors) pair.
template <int C, int KC, bool KFIX, bool FROM_Y>
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
    WarpStream<MU_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * MU_STAGES,
             smem + p.ring_off + (size_t)warp * MU_STAGES * p.g.stage_bytes, FROM_Y ? a.Y : a.X,
             (int)(blockIdx.x * wpc + warp), (int)(gridDim.x * wpc), (int)p.n_items, p.n_kc, lane);

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
        st.issue_next();
        if (st.first_slab()) {
            const int bf = st.cons.item / p.n_kc;
            k0 = (st.cons.item - bf * p.n_kc) * KC;
            b = bf / a.F;
            f = bf - b * a.F;
            vrow = a.act + (size_t)b * N * K * Tp;
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

        const cf* xs = st.acquire();
        const int nf = st.frames();
        const int tbase = st.frame0();
#pragma unroll 2
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float4 xv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * nf + tt);
            float2 P[C];
            frame_power2<C, FROM_Y>(xv, w, P);
            const int t = tbase + tt;
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float* v = vrow + (size_t)n * K * Tp + t;
                float2 tv = make_float2(0.f, 0.f);
                float2 vk[KC];
                if (KFIX) {
#pragma unroll
                    for (int kk = 0; kk < KC; ++kk) {
                        vk[kk] = __ldg(reinterpret_cast<const float2*>(v + (size_t)kk * Tp));
                        tv = __ffma2_rn(vk[kk], make_float2(tk[n][kk], tk[n][kk]), tv);
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < KC; ++kk) vk[kk] = make_float2(0.f, 0.f);
                    for (int k = 0; k < K; ++k) {
                        const float2 vv = __ldg(reinterpret_cast<const float2*>(v + (size_t)k * Tp));
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
                mu_stats2(a.mode, P[n], tv, a.p_exp, a.nu, sa, sb);
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    num[n][kk] = __ffma2_rn(sa, vk[kk], num[n][kk]);
                    den[n][kk] = __ffma2_rn(sb, vk[kk], den[n][kk]);
                }
            }
        }

        if (st.last_slab()) {
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
        st.release();
    }
}

template <int C, int KC, bool KFIX, bool FROM_Y>
int launch_mu_basis_t(bss_handle* h, const MuArgs& a) {
    MuParams p;
    p.a = a;
    p.g = make_tile_geom(C, a.Tp, MU_SLAB);
    p.n_kc = KFIX ? 1 : (a.K + KC - 1) / KC;
    p.n_items = (long long)a.B * a.F * p.n_kc;
    constexpr int MP = (C * KC * 2 + 31) / 32 * 32;
    StreamPlan sp;
    if (!plan_stream(h, p.g, MU_STAGES, ((size_t)C * a.K + MP) * 4, (int)p.n_items, MU_MAX_WARPS, &sp))
        return bss_fail(h, BSS_EINVAL, "source model: frame tile does not fit in shared memory");
    p.scratch_off = sp.scratch_off;
    p.scratch_stride = sp.scratch_stride;
    p.ring_off = sp.ring_off;
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(mu_basis_kernel<C, KC, KFIX, FROM_Y>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         h->max_smem));
        attr_done = true;
    }
    mu_basis_kernel<C, KC, KFIX, FROM_Y><<<sp.grid, sp.wpc * 32, sp.smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
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
__global__ void __launch_bounds__(256) mu_act_finish_kernel(const MuArgs a, const float* part, float* act, int N, int n_chunks) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)a.B * N * a.K * a.Tp;
    if (idx >= total) return;
    const int t = (int)(idx % a.Tp);
    long long r = idx / a.Tp;
    const int k = (int)(r % a.K);
    r /= a.K;
    const int n = (int)(r % N);
    const int b = (int)(r / N);
    if (t >= a.T) {
        act[idx] = 0.f;
        return;
    }
    const bool sel = a.sel_m < 0 || n == a.sel_m || n == a.sel_n;
    if (!sel) return;
    float num = 0.f, den = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
        const float* src = part + (((((size_t)b * n_chunks + c) * N + n) * a.K + k) * 2) * a.Tp + t;
        num += src[0];
        den += src[a.Tp];
    }
    den = fmaxf(den, a.eps);
    act[idx] = act[idx] * pow_q(num / den, a.q_exp);
}

// act == nullptr: stage 1 only (the partial sums stay in h->part, *n_chunks_out tells how many)
template <int C, int KC, bool KFIX, bool FROM_Y>
int launch_mu_act_t(bss_handle* h, const MuArgs& a, float* act, int* n_chunks_out) {
    const int n_kc = KFIX ? 1 : (a.K + KC - 1) / KC;
    const int n_slabs = (a.Tp + 63) / 64;
    // enough warps to fill the machine a few times over, but chunks of at least 4 bins
    long long want = (long long)h->n_sm * 32;
    long long per_chunk_items = (long long)a.B * n_slabs * n_kc;
    int n_chunks = (int)cdiv(want, per_chunk_items);
    if (n_chunks < 1) n_chunks = 1;
    int bins_per_chunk = (int)cdiv(a.F, n_chunks);
    if (bins_per_chunk < 4) bins_per_chunk = a.F < 4 ? a.F : 4;
    n_chunks = (int)cdiv(a.F, bins_per_chunk);
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
    const long long total = (long long)a.B * C * a.K * a.Tp;
    mu_act_finish_kernel<<<(unsigned)cdiv(total, 256), 256, 0, h->stream>>>(a, h->part, act, C, n_chunks);
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
