// Weighted spatial-covariance accumulate:  U[w,f] = (1/T) sum_t x_ft x_ft^H * iw[w,f,t]
// (src/bss/ilrma.py:503-511, src/bss/iva.py:491-499, src/bss/mnmf.py:875).
//
// One warp owns one (mixture, bin, weight-group) item at a time.  Lane 0 streams the bin's
// C x Tp complex64 tile from HBM into the warp's private shared-memory ring with bulk async
// copies (TMA engine, UBLKCP) signalled through mbarriers; the 32 lanes then walk the frames two
// at a time (one conflict-free LDS.128 per channel), recompute the weights from the low-rank
// source model in registers, and accumulate the packed Hermitian outer products with paired
// fp32 FMAs (FFMA2: one instruction updates the (re, im) pair of an entry; the weight and the
// swapped/negated operand come for free as operand modifiers).  A butterfly reduce-scatter over
// the lanes finishes the bin.  Nothing of size (N,F,T,C,C) -- the tensor the reference
// materialises -- ever exists.
#include "handle.h"

namespace {

constexpr int COV_STAGES = 3;
constexpr int COV_SLAB = 128;   // frames per ring stage
constexpr int COV_MAX_WARPS = 16;

struct CovParams {
    CovArgs a;
    TileGeom g;
    long long n_items;   // B * F * n_groups
    int n_groups;
    double inv_T;
    uint32_t scratch_off, scratch_stride, ring_off;
    uint32_t cache_off;  // CACHE: shared-memory copy of the activation rows of the (at most two) mixtures a CTA touches
};

// Pair layout of one Hermitian accumulator: DP = ceil(C/2) diagonal pairs (d0,d1), (d2,d3), ...
// followed by the strictly-lower entries (i > j, row major) as (re, im).
template <int C>
struct Pairs {
    static constexpr int DP = (C + 1) / 2;
    static constexpr int NP = DP + C * (C - 1) / 2;
};

template <int C, int NS>
__device__ __forceinline__ void accumulate_frame(float2 (&acc)[NS][Pairs<C>::NP], const float2 (&x)[C], const float (&w)[NS]) {
    constexpr int DP = Pairs<C>::DP;
    float2 v[Pairs<C>::NP];
#pragma unroll
    for (int i = 0; i < DP; ++i) {
        const float d0 = fmaf(x[2 * i].x, x[2 * i].x, x[2 * i].y * x[2 * i].y);
        float d1 = 0.f;
        if (2 * i + 1 < C) d1 = fmaf(x[2 * i + 1].x, x[2 * i + 1].x, x[2 * i + 1].y * x[2 * i + 1].y);
        v[i] = make_float2(d0, d1);
    }
    int e = DP;
#pragma unroll
    for (int i = 1; i < C; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) {
            // x_i conj(x_j) = x_i * x_j.x + (x_i.y, -x_i.x) * x_j.y
            const float2 t = __fmul2_rn(x[i], make_float2(x[j].x, x[j].x));
            v[e] = __ffma2_rn(make_float2(x[i].y, -x[i].x), make_float2(x[j].y, x[j].y), t);
            ++e;
        }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const float2 ww = make_float2(w[s], w[s]);
#pragma unroll
        for (int q = 0; q < Pairs<C>::NP; ++q) acc[s][q] = __ffma2_rn(v[q], ww, acc[s][q]);
    }
}

// KT > 0: n_basis known at compile time (basis row in registers); KT == 0: run-time K, basis row in smem.
// CACHE: every CTA owns a contiguous range of at most F bins, i.e. at most two mixtures; their activation rows
// (N K Tp floats each, shared by all bins of a mixture) are copied to shared memory once, so the weight
// computation reads LDS instead of L2-latency LDG (the loads it replaces cost 30% of the kernel time).
template <int C, int NS, int WM, int KT, bool POW, bool CACHE>
__global__ void __launch_bounds__(COV_MAX_WARPS * 32, 1) cov_kernel(const CovParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const CovArgs& a = p.a;
    constexpr int NP = Pairs<C>::NP;
    constexpr int DP = Pairs<C>::DP;
    constexpr int M = NS * NP * 2;
    constexpr int MP = (M + 31) / 32 * 32;
    constexpr int Q = MP / 32;
    constexpr int CC = C * C;

    float* tb = reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride);
    const int Tp = a.Tp;
    const int K = KT > 0 ? KT : a.K;
    // a CTA walks the contiguous item range [lo, hi); its warps interleave inside it
    int lo, hi;
    cta_item_range((int)p.n_items, lo, hi);
    // without the shared-memory cache the activation rows are read through L1: a two-stage ring leaves L1 room for them
    constexpr int STG = (WM == WM_ILRMA && !CACHE) ? 2 : COV_STAGES;
    WarpStream<STG> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * STG,
             smem + p.ring_off + (size_t)warp * STG * p.g.stage_bytes, a.X, lo + warp, wpc, hi, p.n_groups, lane);
    const float* vcache = reinterpret_cast<const float*>(smem + p.cache_off);
    int b_lo = 0;
    if (CACHE) b_lo = load_act_cache(reinterpret_cast<float*>(smem + p.cache_off), a.act, a.NW * K * Tp, lo, hi, p.n_groups, a.F);

    float2 acc[NS][NP];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[s][q] = make_float2(0.f, 0.f);

    // per-item state
    int b = 0, f = 0, w0 = 0;
    const float* wrow[NS];          // per weight set: activation rows / frame weights / explicit weights of this item
    int woff[NS];                   // CACHE: float offset of the weight set's activation rows inside vcache
    float tbr[NS][KT > 0 ? KT : 1];

#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, a.X, p.n_groups);
        if (st.first_slab()) {
            const int bf = st.cons.item / p.n_groups;
            const int grp = st.cons.item - bf * p.n_groups;
            b = bf / a.F;
            f = bf - b * a.F;
            w0 = grp * NS;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                // weight sets past n_sel reuse set 0: they accumulate values that are never stored
                const int sw = w0 + s < a.n_sel ? a.wsel[w0 + s] : 0;
                if (WM == WM_ILRMA) {
                    wrow[s] = a.act + ((size_t)b * a.NW + sw) * K * Tp;
                    woff[s] = ((b - b_lo) * a.NW + sw) * K * Tp;
                    const float* tsrc = a.basis + (((size_t)b * a.NW + sw) * a.F + f) * K;
                    if (KT > 0) {
#pragma unroll
                        for (int k = 0; k < (KT > 0 ? KT : 1); ++k) tbr[s][k] = __ldg(tsrc + k);
                    } else {
                        for (int k = lane; k < K; k += 32) tb[s * K + k] = __ldg(tsrc + k);
                    }
                } else if (WM == WM_FRAME) {
                    wrow[s] = a.wfr + ((size_t)b * a.NW + sw) * Tp;
                } else if (WM == WM_EXPLICIT) {
                    wrow[s] = a.iw + (((size_t)b * a.F + f) * a.NW + sw) * Tp;
                } else {
                    wrow[s] = nullptr;
                }
            }
            if (WM == WM_ILRMA && KT == 0) __syncwarp();
        }

        const cf* xs = st.acquire(p.g);
        const int nf = st.frames(p.g);
        const int tbase = st.frame0(p.g);

#pragma unroll 1   // unrolling spills: the 64 accumulator registers leave no room for a second frame pair in flight
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float4 xv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * nf + tt);
            const int t = tbase + tt;
            float wa[NS], wb[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (WM == WM_UNIT) {
                    wa[s] = wb[s] = 1.f;
                } else if (WM == WM_ILRMA) {
                    float2 r = make_float2(0.f, 0.f);
                    if (KT > 0) {
#pragma unroll
                        for (int k = 0; k < (KT > 0 ? KT : 1); ++k) {
                            const float2 vv = CACHE ? *reinterpret_cast<const float2*>(vcache + woff[s] + k * Tp + t)
                                                    : __ldg(reinterpret_cast<const float2*>(wrow[s] + (size_t)k * Tp + t));
                            r = __ffma2_rn(vv, make_float2(tbr[s][k], tbr[s][k]), r);
                        }
                    } else {
                        for (int k = 0; k < K; ++k) {
                            const float2 vv = CACHE ? *reinterpret_cast<const float2*>(vcache + woff[s] + k * Tp + t)
                                                    : __ldg(reinterpret_cast<const float2*>(wrow[s] + (size_t)k * Tp + t));
                            const float tk = tb[s * K + k];
                            r = __ffma2_rn(vv, make_float2(tk, tk), r);
                        }
                    }
                    if (POW) {
                        r.x = powf(r.x, a.expo);
                        r.y = powf(r.y, a.expo);
                    }
                    r.x = fmaxf(r.x, a.eps);
                    r.y = fmaxf(r.y, a.eps);
                    wa[s] = rcp_fast(r.x);
                    wb[s] = rcp_fast(r.y);
                } else {
                    const float2 vv = __ldg(reinterpret_cast<const float2*>(wrow[s] + t));
                    wa[s] = vv.x;
                    wb[s] = vv.y;
                }
            }
            float2 x0[C], x1[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                x0[c] = make_float2(xv[c].x, xv[c].y);
                x1[c] = make_float2(xv[c].z, xv[c].w);
            }
            accumulate_frame<C, NS>(acc, x0, wa);
            accumulate_frame<C, NS>(acc, x1, wb);
        }

        if (st.last_slab(p.g)) {
            float flat[MP];
#pragma unroll
            for (int i = 0; i < MP; ++i) flat[i] = 0.f;
#pragma unroll
            for (int s = 0; s < NS; ++s)
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    flat[(s * NP + q) * 2] = acc[s][q].x;
                    flat[(s * NP + q) * 2 + 1] = acc[s][q].y;
                    acc[s][q] = make_float2(0.f, 0.f);
                }
            warp_reduce_scatter<MP>(flat, lane);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int e = Q * lane + q;
                const int s = e / (2 * NP);
                int r = e - s * 2 * NP;   // float index inside the pair layout
                // pair layout -> packed output layout (C diagonal reals, then lower (re, im) pairs)
                bool valid = s < NS && w0 + s < a.n_sel;
                if (r >= 2 * DP)
                    r = r - 2 * DP + C;
                else if (r >= C)
                    valid = false;   // padding slot of an odd C
                if (valid) {
                    const int sw = a.wsel[w0 + s];
                    a.U[(((size_t)b * a.NW + sw) * a.F + f) * CC + r] = (double)flat[q] * p.inv_T;
                }
            }
        }
        st.release(p.g);
    }
}

template <int C, int NS, int WM, int KT, bool POW, bool CACHE>
int launch_cov_c(bss_handle* h, CovParams& p, const StreamPlan& sp, size_t smem_bytes) {
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(cov_kernel<C, NS, WM, KT, POW, CACHE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         h->max_smem));
        attr_done = true;
    }
    cov_kernel<C, NS, WM, KT, POW, CACHE><<<sp.grid, sp.wpc * 32, smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

template <int C, int NS, int WM, int KT, bool POW>
int launch_cov_t(bss_handle* h, const CovArgs& a) {
    CovParams p;
    p.a = a;
    p.g = make_tile_geom(C, a.Tp, COV_SLAB);
    p.n_groups = (a.n_sel + NS - 1) / NS;
    p.n_items = (long long)a.B * a.F * p.n_groups;
    p.inv_T = 1.0 / (double)a.T;
    p.cache_off = 0;
    if (p.n_items == 0) return BSS_OK;
    const size_t scratch = (size_t)NS * (a.K > 0 ? a.K : 1) * 4;
    StreamPlan sp;
    size_t smem_bytes = 0;
    if (WM == WM_ILRMA && plan_stream_cached(h, p.g, COV_STAGES, scratch, p.n_items, COV_MAX_WARPS, (size_t)a.NW * a.K * a.Tp * sizeof(float),
                                             (long long)a.F * p.n_groups, &sp, &p.cache_off, &smem_bytes)) {
        p.scratch_off = sp.scratch_off;
        p.scratch_stride = sp.scratch_stride;
        p.ring_off = sp.ring_off;
        return launch_cov_c<C, NS, WM, KT, POW, true>(h, p, sp, smem_bytes);
    }
    if (!plan_stream(h, p.g, WM == WM_ILRMA ? 2 : COV_STAGES, scratch, (int)p.n_items, COV_MAX_WARPS, &sp))
        return bss_fail(h, BSS_EINVAL, "covariance: frame tile does not fit in shared memory");
    p.scratch_off = sp.scratch_off;
    p.scratch_stride = sp.scratch_stride;
    p.ring_off = sp.ring_off;
    return launch_cov_c<C, NS, WM, KT, POW, false>(h, p, sp, sp.smem_bytes);
}

template <int C, int NS>
int launch_cov_wm(bss_handle* h, const CovArgs& a) {
    switch (a.wmode) {
        case WM_UNIT: return launch_cov_t<C, NS, WM_UNIT, 1, false>(h, a);
        case WM_ILRMA:
            if (a.expo != 1.f) return launch_cov_t<C, NS, WM_ILRMA, 0, true>(h, a);
            if (a.K == 2) return launch_cov_t<C, NS, WM_ILRMA, 2, false>(h, a);
            return launch_cov_t<C, NS, WM_ILRMA, 0, false>(h, a);
        case WM_FRAME: return launch_cov_t<C, NS, WM_FRAME, 1, false>(h, a);
        case WM_EXPLICIT: return launch_cov_t<C, NS, WM_EXPLICIT, 1, false>(h, a);
    }
    return bss_fail(h, BSS_EINVAL, "covariance: unknown weight mode");
}

// Plain covariance Cx[f] = (1/T) sum_t x x^H with fp64 accumulation of the (exact in fp64) products of the complex64 samples.
// Computed once per input.  Power normalisation and projection back evaluate w^H Cx w and Cx W^H (W Cx W^H)^-1 from it
// (src/bss/ilrma.py:305-311, src/algorithm/projection_back.py:15-21): on real recordings a separated source can be 1e3 - 1e5
// times weaker than |w| |x| in a bin, so these forms cancel almost completely and the 1e-7 relative error of an fp32-accumulated
// Cx showed up as a 0.3 % error of the normalisation constants (tests/test_gpu_parity_large.py, sample-2 recording).
// One warp per (mixture, bin); lanes stride over the frames of the block-interleaved tile; packed Hermitian output.
template <int C>
__global__ void __launch_bounds__(256) cx_kernel(const cf* X, double* Cx, long long n_bins, int Tp, double inv_T) {
    const long long bin = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (bin >= n_bins) return;
    const int lane = threadIdx.x & 31;
    const cf* tile = X + (size_t)bin * C * Tp;
    double d[C];
    double lr[C * (C - 1) / 2 + 1], li[C * (C - 1) / 2 + 1];
#pragma unroll
    for (int i = 0; i < C; ++i) d[i] = 0.0;
#pragma unroll
    for (int e = 0; e < C * (C - 1) / 2; ++e) lr[e] = li[e] = 0.0;
    for (int t = lane; t < Tp; t += 32) {
        double xr[C], xi[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const cf v = __ldg(tile + tile_off(C, Tp, c, t));
            xr[c] = (double)v.x;
            xi[c] = (double)v.y;
        }
        int e = 0;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            d[i] = fma(xr[i], xr[i], fma(xi[i], xi[i], d[i]));
#pragma unroll
            for (int j = 0; j < i; ++j) {
                // x_i conj(x_j)
                lr[e] = fma(xr[i], xr[j], fma(xi[i], xi[j], lr[e]));
                li[e] = fma(xi[i], xr[j], fma(-xr[i], xi[j], li[e]));
                ++e;
            }
        }
    }
    double* out = Cx + (size_t)bin * C * C;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        const double v = warp_sum(d[i]);
        if (lane == 0) out[i] = v * inv_T;
    }
#pragma unroll
    for (int e = 0; e < C * (C - 1) / 2; ++e) {
        const double vr = warp_sum(lr[e]), vi = warp_sum(li[e]);
        if (lane == 0) {
            out[C + 2 * e] = vr * inv_T;
            out[C + 2 * e + 1] = vi * inv_T;
        }
    }
}

}  // namespace

int launch_plain_covariance(bss_handle* h, const cf* X, double* Cx, int B, int F, int C, int T, int Tp) {
    const long long n_bins = (long long)B * F;
    if (n_bins == 0) return BSS_OK;
    const unsigned grid = (unsigned)cdiv(n_bins, 8);
    const double inv_T = 1.0 / (double)T;
    switch (C) {
        case 2: cx_kernel<2><<<grid, 256, 0, h->stream>>>(X, Cx, n_bins, Tp, inv_T); break;
        case 3: cx_kernel<3><<<grid, 256, 0, h->stream>>>(X, Cx, n_bins, Tp, inv_T); break;
        case 4: cx_kernel<4><<<grid, 256, 0, h->stream>>>(X, Cx, n_bins, Tp, inv_T); break;
        case 5: cx_kernel<5><<<grid, 256, 0, h->stream>>>(X, Cx, n_bins, Tp, inv_T); break;
        case 6: cx_kernel<6><<<grid, 256, 0, h->stream>>>(X, Cx, n_bins, Tp, inv_T); break;
        case 7: cx_kernel<7><<<grid, 256, 0, h->stream>>>(X, Cx, n_bins, Tp, inv_T); break;
        case 8: cx_kernel<8><<<grid, 256, 0, h->stream>>>(X, Cx, n_bins, Tp, inv_T); break;
        default: return bss_fail(h, BSS_EINVAL, "n_channels must be between 2 and 8");
    }
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_covariance(bss_handle* h, const CovArgs& a) {
    switch (a.C) {
        case 2: return launch_cov_wm<2, 2>(h, a);
        case 3: return launch_cov_wm<3, 3>(h, a);
        case 4: return launch_cov_wm<4, 4>(h, a);
        case 5: return launch_cov_wm<5, 2>(h, a);
        case 6: return launch_cov_wm<6, 2>(h, a);
        case 7: return launch_cov_wm<7, 1>(h, a);
        case 8: return launch_cov_wm<8, 1>(h, a);
    }
    return bss_fail(h, BSS_EINVAL, "n_channels must be between 2 and 8");
}
