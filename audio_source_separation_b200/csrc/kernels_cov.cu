// Weighted spatial-covariance accumulate:  U[w,f] = (1/T) sum_t x_ft x_ft^H * iw[w,f,t]
// (src/bss/ilrma.py:503-511, src/bss/iva.py:491-499, src/bss/mnmf.py:875).
//
// One warp owns one (mixture, bin, weight-group) item at a time.  Lane 0 streams the bin's
// C x Tp complex64 tile from HBM into the warp's private shared-memory ring with bulk async
// copies (TMA engine) signalled through mbarriers; the 32 lanes then walk the frames two at a
// time (one conflict-free LDS.128 per channel), recompute the weights from the low-rank source
// model in registers, accumulate the packed Hermitian outer products in fp32 registers and
// finish with a butterfly reduce-scatter over the lanes.  Nothing of size (N,F,T,C,C) -- the
// tensor the reference materialises -- ever exists.
#include "handle.h"

namespace {

constexpr int COV_STAGES = 3;

struct CovParams {
    CovArgs a;
    TileGeom g;
    long long n_items;   // B * F * n_groups
    int n_groups;
    double inv_T;
    uint32_t scratch_off, scratch_stride, ring_off;
};

template <int C, int NS>
__device__ __forceinline__ void accumulate_frame(float* acc, const cf* x, const float* w) {
    constexpr int CC = C * C;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        const float d = fmaf(x[i].x, x[i].x, x[i].y * x[i].y);
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[s * CC + i] = fmaf(w[s], d, acc[s * CC + i]);
    }
    int e = C;
#pragma unroll
    for (int i = 1; i < C; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) {
            // x_i conj(x_j)
            const float re = fmaf(x[i].x, x[j].x, x[i].y * x[j].y);
            const float im = fmaf(x[i].y, x[j].x, -x[i].x * x[j].y);
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                acc[s * CC + e] = fmaf(w[s], re, acc[s * CC + e]);
                acc[s * CC + e + 1] = fmaf(w[s], im, acc[s * CC + e + 1]);
            }
            e += 2;
        }
}

template <int C, int NS, int WM>
__global__ void __launch_bounds__(256) cov_kernel(const CovParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const CovArgs& a = p.a;
    constexpr int CC = C * C;
    constexpr int M = NS * CC;
    constexpr int MP = (M + 31) / 32 * 32;
    constexpr int Q = MP / 32;

    float* tb = reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride);
    WarpStream<COV_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * COV_STAGES,
             smem + p.ring_off + (size_t)warp * COV_STAGES * p.g.stage_bytes, a.X, (long long)blockIdx.x * wpc + warp,
             (long long)gridDim.x * wpc, p.n_items, p.n_groups, lane);
    const int row_stride = p.g.row_stride;

    float acc[MP];
#pragma unroll
    for (int i = 0; i < MP; ++i) acc[i] = 0.f;

#pragma unroll 1
    while (st.active()) {
        st.issue_next();
        const long long bf = st.cons.item / p.n_groups;
        const int grp = (int)(st.cons.item - bf * p.n_groups);
        const int b = (int)(bf / a.F), f = (int)(bf - (long long)b * a.F);
        const int w0 = grp * NS;

        int src[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) src[s] = (w0 + s < a.n_sel) ? a.wsel[w0 + s] : -1;

        if (WM == WM_ILRMA && st.first_slab()) {
            for (int i = lane; i < NS * a.K; i += 32) {
                const int s = i / a.K, k = i - s * a.K;
                const int sw = (w0 + s < a.n_sel) ? a.wsel[w0 + s] : 0;
                tb[i] = a.basis[(((size_t)b * a.NW + sw) * a.F + f) * a.K + k];
            }
            __syncwarp();
        }

        const cf* xs = st.acquire();
        const int nf = st.frames();
        const int tbase = st.frame0();

#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float4 xv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * row_stride + tt);
            const int t = tbase + tt;
            float wa[NS], wb[NS];
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                if (WM == WM_UNIT) {
                    wa[s] = wb[s] = (src[s] >= 0) ? 1.f : 0.f;
                } else if (src[s] < 0) {
                    wa[s] = wb[s] = 0.f;
                } else if (WM == WM_ILRMA) {
                    const float* v = a.act + ((size_t)b * a.NW + src[s]) * a.K * a.Tp + t;
                    float ra = 0.f, rb = 0.f;
                    for (int k = 0; k < a.K; ++k) {
                        const float2 vv = __ldg(reinterpret_cast<const float2*>(v + (size_t)k * a.Tp));
                        const float tk = tb[s * a.K + k];
                        ra = fmaf(tk, vv.x, ra);
                        rb = fmaf(tk, vv.y, rb);
                    }
                    if (a.expo != 1.f) {
                        ra = powf(ra, a.expo);
                        rb = powf(rb, a.expo);
                    }
                    ra = ra < a.eps ? a.eps : ra;
                    rb = rb < a.eps ? a.eps : rb;
                    wa[s] = __frcp_rn(ra);
                    wb[s] = __frcp_rn(rb);
                } else if (WM == WM_FRAME) {
                    const float2 vv =
                        __ldg(reinterpret_cast<const float2*>(a.wfr + ((size_t)b * a.NW + src[s]) * a.Tp + t));
                    wa[s] = vv.x;
                    wb[s] = vv.y;
                } else {
                    const float2 vv = __ldg(reinterpret_cast<const float2*>(
                        a.iw + (((size_t)b * a.F + f) * a.NW + src[s]) * a.Tp + t));
                    wa[s] = vv.x;
                    wb[s] = vv.y;
                }
            }
            cf x0[C], x1[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                x0[c] = cf_make(xv[c].x, xv[c].y);
                x1[c] = cf_make(xv[c].z, xv[c].w);
            }
            accumulate_frame<C, NS>(acc, x0, wa);
            accumulate_frame<C, NS>(acc, x1, wb);
        }

        if (st.last_slab()) {
            warp_reduce_scatter<MP>(acc, lane);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const int e = Q * lane + q;
                const int s = e / CC, r = e - s * CC;
                if (s < NS && w0 + s < a.n_sel) {
                    const int sw = a.wsel[w0 + s];
                    a.U[(((size_t)b * a.NW + sw) * a.F + f) * CC + r] = (double)acc[q] * p.inv_T;
                }
            }
#pragma unroll
            for (int i = 0; i < MP; ++i) acc[i] = 0.f;
        }
        st.release();
    }
}

template <int C, int NS, int WM>
int launch_cov_t(bss_handle* h, const CovArgs& a) {
    CovParams p;
    p.a = a;
    p.g = make_tile_geom(C, a.Tp);
    p.n_groups = (a.n_sel + NS - 1) / NS;
    p.n_items = (long long)a.B * a.F * p.n_groups;
    p.inv_T = 1.0 / (double)a.T;
    if (p.n_items == 0) return BSS_OK;
    StreamPlan sp;
    if (!plan_stream(h, p.g, COV_STAGES, (size_t)NS * (a.K > 0 ? a.K : 1) * 4, p.n_items, 8, &sp))
        return bss_fail(h, BSS_EINVAL, "covariance: frame tile does not fit in shared memory");
    p.scratch_off = sp.scratch_off;
    p.scratch_stride = sp.scratch_stride;
    p.ring_off = sp.ring_off;
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(cov_kernel<C, NS, WM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         h->max_smem));
        attr_done = true;
    }
    cov_kernel<C, NS, WM><<<sp.grid, sp.wpc * 32, sp.smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

template <int C, int NS>
int launch_cov_wm(bss_handle* h, const CovArgs& a) {
    switch (a.wmode) {
        case WM_UNIT: return launch_cov_t<C, NS, WM_UNIT>(h, a);
        case WM_ILRMA: return launch_cov_t<C, NS, WM_ILRMA>(h, a);
        case WM_FRAME: return launch_cov_t<C, NS, WM_FRAME>(h, a);
        case WM_EXPLICIT: return launch_cov_t<C, NS, WM_EXPLICIT>(h, a);
    }
    return bss_fail(h, BSS_EINVAL, "covariance: unknown weight mode");
}

}  // namespace

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
