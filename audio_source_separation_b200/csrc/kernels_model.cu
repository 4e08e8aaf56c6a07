// Source-model kernels of the ILRMA family and the kernels around them:
//   - multiplicative update of the NMF basis T (per-bin reductions over frames)        src/bss/ilrma.py:413-419, :915-927
//   - multiplicative update of the activation V (cross-bin reduction, two stages)      src/bss/ilrma.py:422-428, :929-938
//   - power normalisation                                                             src/bss/ilrma.py:304-322
//   - demixing Y = W X                                                                 src/bss/ilrma.py:153-165
//   - negative log-likelihood                                                          src/bss/ilrma.py:648-677, :993-1020
// The estimates Y = W X are never stored by the update loop: every consumer recomputes them from
// the bin tile it has just staged.
#include "handle.h"

namespace {

constexpr int MU_STAGES = 3;

template <int C, bool FROM_Y>
__device__ __forceinline__ void load_filter(cf (&w)[C][C], const cf* Wf) {
    if (!FROM_Y) {
#pragma unroll
        for (int n = 0; n < C; ++n)
#pragma unroll
            for (int c = 0; c < C; ++c) w[n][c] = __ldg(Wf + n * C + c);
    }
}

// power of the estimates for two consecutive frames held in xv (float4 = 2 complex frames per row)
template <int C, bool FROM_Y>
__device__ __forceinline__ void frame_power(const float4 (&xv)[C], const cf (&w)[C][C], float (&P0)[C], float (&P1)[C]) {
#pragma unroll
    for (int n = 0; n < C; ++n) {
        if (FROM_Y) {
            P0[n] = fmaf(xv[n].x, xv[n].x, xv[n].y * xv[n].y);
            P1[n] = fmaf(xv[n].z, xv[n].z, xv[n].w * xv[n].w);
        } else {
            cf y0 = cf_make(0.f, 0.f), y1 = cf_make(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                cf_fma(y0, w[n][c], cf_make(xv[c].x, xv[c].y));
                cf_fma(y1, w[n][c], cf_make(xv[c].z, xv[c].w));
            }
            P0[n] = cf_abs2(y0);
            P1[n] = cf_abs2(y1);
        }
    }
}

// ------------------------------------------------------------------------------------------- normalisation
// aux_n = max(sqrt(mean_f pw[n,f]), eps);  W[:,n,:] /= aux_n;  T[n] /= aux_n^domain     src/bss/ilrma.py:305-322
// Two launches: one block per (mixture, source) reduces the per-bin powers in a fixed order (fp64, deterministic),
// then a fully parallel pass rescales the filters, their fp32 mirror and the basis.
__global__ void __launch_bounds__(256) power_aux_kernel(const double* pw, double* aux, int F, double eps) {
    __shared__ double red[8];
    const long long bn = blockIdx.x;
    double s = 0.0;
    for (int f = threadIdx.x; f < F; f += blockDim.x) s += pw[(size_t)bn * F + f];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        const double a = sqrt(t / (double)F);
        aux[bn] = a < eps ? eps : a;
    }
}

__global__ void __launch_bounds__(256) normalize_power_kernel(double2* W, cf* Wf, float* basis, const double* aux, int B, int N, int C,
                                                              int F, int K, double domain) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nW = (long long)B * F * N * C;
    const long long nT = basis ? (long long)B * N * F * K : 0;
    if (idx < nW) {
        const int n = (int)((idx / C) % N);
        const int b = (int)(idx / ((long long)F * N * C));
        const double inv = 1.0 / aux[(size_t)b * N + n];
        double2 v = W[idx];
        v.x *= inv;
        v.y *= inv;
        W[idx] = v;
        Wf[idx] = cf_make((float)v.x, (float)v.y);
    } else if (idx < nW + nT) {
        const long long i = idx - nW;
        const long long bn = i / ((long long)F * K);
        const double a = aux[bn];
        const double sc = domain == 2.0 ? a * a : pow(a, domain);
        basis[i] = (float)((double)basis[i] / sc);
    }
}

// W[f,n,:] *= scale[n,f];  T[n,f,:] *= |scale[n,f]|^domain     src/bss/ilrma.py:323-330
__global__ void __launch_bounds__(256) normalize_pb_kernel(double2* W, cf* Wf, float* basis, const double2* scale, int B, int N,
                                                           int C, int F, int K, double domain) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * F * N) return;
    const int n = (int)(idx % N);
    const long long bf = idx / N;
    const int f = (int)(bf % F);
    const int b = (int)(bf / F);
    const double2 s = scale[((size_t)b * N + n) * F + f];
    for (int c = 0; c < C; ++c) {
        const double2 v = W[(size_t)idx * C + c];
        const double2 r = make_double2(v.x * s.x - v.y * s.y, v.x * s.y + v.y * s.x);
        W[(size_t)idx * C + c] = r;
        Wf[(size_t)idx * C + c] = cf_make((float)r.x, (float)r.y);
    }
    if (basis) {
        const double m = hypot(s.x, s.y);
        const double sc = domain == 2.0 ? m * m : pow(m, domain);
        for (int k = 0; k < K; ++k) {
            float* p = basis + (((size_t)b * N + n) * F + f) * K + k;
            *p = (float)((double)*p * sc);
        }
    }
}

__global__ void __launch_bounds__(256) widen_kernel(const cf* in, double2* out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2((double)in[i].x, (double)in[i].y);
}

__global__ void __launch_bounds__(256) sync_wf_kernel(const double2* W, cf* Wf, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Wf[i] = cf_make((float)W[i].x, (float)W[i].y);
}

// W = I for every bin (src/bss/ilrma.py:67-69), both precisions, without a host round trip
__global__ void __launch_bounds__(256) identity_filter_kernel(double2* W, cf* Wf, long long n, int N, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int rc = (int)(i % ((long long)N * C));
    const double v = (rc / C) == (rc % C) ? 1.0 : 0.0;
    W[i] = make_double2(v, 0.0);
    Wf[i] = cf_make((float)v, 0.f);
}

// ------------------------------------------------------------------------------------------- demixing
struct SepParams {
    const cf* X;
    const cf* Wf;
    const double2* scale;   // [B][N][F] or null
    cf* Y;                  // [B][F][N][Tp] or null
    cf* out;                // [B][N][F][T]  or null (reference layout)
    int B, F, T, Tp;
    TileGeom g;
    long long n_items;
    uint32_t scratch_off, scratch_stride, ring_off;
};

template <int C>
__global__ void __launch_bounds__(256) separate_kernel(const SepParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    WarpStream<MU_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * MU_STAGES,
             smem + p.ring_off + (size_t)warp * MU_STAGES * p.g.stage_bytes, p.X, (int)(blockIdx.x * wpc + warp),
             (int)(gridDim.x * wpc), (int)p.n_items, 1, lane);
    cf w[C][C];
#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, p.X, 1);
        const int bf = st.cons.item;
        const int b = bf / p.F, f = bf - b * p.F;
        if (st.first_slab()) {
            load_filter<C, false>(w, p.Wf + (size_t)bf * C * C);
            if (p.scale) {
#pragma unroll
                for (int n = 0; n < C; ++n) {
                    const double2 sd = p.scale[((size_t)b * C + n) * p.F + f];
                    const cf s = cf_make((float)sd.x, (float)sd.y);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const cf v = w[n][c];
                        w[n][c] = cf_make(v.x * s.x - v.y * s.y, v.x * s.y + v.y * s.x);
                    }
                }
            }
        }
        const cf* xs = st.acquire(p.g);
        const int nf = st.frames(p.g);
        const int tbase = st.frame0(p.g);
#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float4 xv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * nf + tt);
            const int t = tbase + tt;
#pragma unroll
            for (int n = 0; n < C; ++n) {
                cf y0 = cf_make(0.f, 0.f), y1 = cf_make(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    cf_fma(y0, w[n][c], cf_make(xv[c].x, xv[c].y));
                    cf_fma(y1, w[n][c], cf_make(xv[c].z, xv[c].w));
                }
                if (p.Y)
                    *reinterpret_cast<float4*>(p.Y + (size_t)bf * C * p.Tp + tile_off(C, p.Tp, n, t)) = make_float4(y0.x, y0.y, y1.x, y1.y);
                if (p.out) {
                    cf* o = p.out + (((size_t)b * C + n) * p.F + f) * p.T + t;
                    if (t < p.T) o[0] = y0;
                    if (t + 1 < p.T) o[1] = y1;
                }
            }
        }
        st.release(p.g);
    }
}

// Y tile -> reference layout with optional per-(n,f) complex scale (ISS output path)
__global__ void __launch_bounds__(256) export_y_kernel(const cf* Y, const double2* scale, cf* out, int B, int N, int F, int T, int Tp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * N * F * T) return;
    const int t = (int)(idx % T);
    long long r = idx / T;
    const int f = (int)(r % F);
    r /= F;
    const int n = (int)(r % N);
    const int b = (int)(r / N);
    cf v = Y[((size_t)b * F + f) * N * Tp + tile_off(N, Tp, n, t)];
    if (scale) {
        const double2 sd = scale[((size_t)b * N + n) * F + f];
        const cf s = cf_make((float)sd.x, (float)sd.y);
        v = cf_make(v.x * s.x - v.y * s.y, v.x * s.y + v.y * s.x);
    }
    out[idx] = v;
}

// ------------------------------------------------------------------------------------------- loss
// per bin: sum_{n,t} (P/R + log R)            (mode 0, src/bss/ilrma.py:669-676)
//          sum_{n,t} ((1+nu/2) log(1 + (2/nu) P/R) + log R)   (mode 1, src/bss/ilrma.py:1012-1019)
// Same streaming structure as the covariance kernel: CTA-contiguous bin ranges, activation rows cached in shared
// memory (CACHE), n_basis = 2 resolved at compile time (KT).
struct LossParams {
    MuArgs a;
    double* out;   // [B][F]
    float expo;    // 2/domain
    TileGeom g;
    long long n_items;
    uint32_t scratch_off, scratch_stride, ring_off, cache_off;
};

template <int C, int KT, bool FROM_Y, bool CACHE>
__global__ void __launch_bounds__(512, 1) ilrma_loss_kernel(const LossParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const MuArgs& a = p.a;
    constexpr int N = C;
    const int K = KT > 0 ? KT : a.K;
    const int Tp = a.Tp;
    float* tb = reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride);
    int lo, hi;
    cta_item_range((int)p.n_items, lo, hi);
    WarpStream<MU_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * MU_STAGES,
             smem + p.ring_off + (size_t)warp * MU_STAGES * p.g.stage_bytes, FROM_Y ? a.Y : a.X, lo + warp, wpc, hi, 1, lane);
    const float* vcache = reinterpret_cast<const float*>(smem + p.cache_off);
    int b_lo = 0;
    if (CACHE) b_lo = load_act_cache(reinterpret_cast<float*>(smem + p.cache_off), a.act, N * K * Tp, lo, hi, 1, a.F);
    cf w[C][C];
    float tkr[N][KT > 0 ? KT : 1];
    double total = 0.0;
    int b = 0, voff = 0;
#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, (FROM_Y ? a.Y : a.X), 1);
        const int bf = st.cons.item;
        if (st.first_slab()) {
            b = bf / a.F;
            const int f = bf - b * a.F;
            voff = (b - b_lo) * N * K * Tp;
            if (KT > 0) {
#pragma unroll
                for (int n = 0; n < N; ++n)
#pragma unroll
                    for (int k = 0; k < (KT > 0 ? KT : 1); ++k) tkr[n][k] = __ldg(a.basis + (((size_t)b * N + n) * a.F + f) * K + k);
            } else {
                for (int i = lane; i < N * K; i += 32) {
                    const int n = i / K, k = i - n * K;
                    tb[i] = a.basis[(((size_t)b * N + n) * a.F + f) * K + k];
                }
            }
            load_filter<C, FROM_Y>(w, a.Wf + (size_t)bf * C * C);
            total = 0.0;
            __syncwarp();
        }
        const cf* xs = st.acquire(p.g);
        const int nf = st.frames(p.g);
        const int tbase = st.frame0(p.g);
        float part = 0.f;
#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float4 xv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * nf + tt);
            float P0[C], P1[C];
            frame_power<C, FROM_Y>(xv, w, P0, P1);
            const int t = tbase + tt;
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float* v = CACHE ? vcache + voff + n * K * Tp + t : a.act + ((size_t)b * N + n) * K * Tp + t;
                float r0 = 0.f, r1 = 0.f;
                if (KT > 0) {
#pragma unroll
                    for (int k = 0; k < (KT > 0 ? KT : 1); ++k) {
                        const float2 vv = CACHE ? *reinterpret_cast<const float2*>(v + k * Tp) : __ldg(reinterpret_cast<const float2*>(v + (size_t)k * Tp));
                        r0 = fmaf(tkr[n][k], vv.x, r0);
                        r1 = fmaf(tkr[n][k], vv.y, r1);
                    }
                } else {
                    for (int k = 0; k < K; ++k) {
                        const float2 vv = CACHE ? *reinterpret_cast<const float2*>(v + k * Tp) : __ldg(reinterpret_cast<const float2*>(v + (size_t)k * Tp));
                        const float tk = tb[n * K + k];
                        r0 = fmaf(tk, vv.x, r0);
                        r1 = fmaf(tk, vv.y, r1);
                    }
                }
                if (p.expo != 1.f) {
                    r0 = powf(r0, p.expo);
                    r1 = powf(r1, p.expo);
                }
                r0 = r0 < a.eps ? a.eps : r0;
                r1 = r1 < a.eps ? a.eps : r1;
                float l0, l1;
                if (a.mode == 0) {
                    l0 = P0[n] * rcp_fast(r0) + __logf(r0);
                    l1 = P1[n] * rcp_fast(r1) + __logf(r1);
                } else {
                    l0 = (1.f + 0.5f * a.nu) * log1pf((2.f / a.nu) * (P0[n] / r0)) + logf(r0);
                    l1 = (1.f + 0.5f * a.nu) * log1pf((2.f / a.nu) * (P1[n] / r1)) + logf(r1);
                }
                if (t < a.T) part += l0;
                if (t + 1 < a.T) part += l1;
            }
        }
        total += (double)part;
        if (st.last_slab(p.g)) {
            const double s = warp_sum(total);
            if (lane == 0) p.out[bf] = s;
        }
        st.release(p.g);
    }
}

// loss[b] = sum_f terms[b,f] - coef * sum_f logdet[b,f]   (fixed-order block reduction, fp64)
__global__ void __launch_bounds__(256) loss_finish_kernel(const double* terms, const double* logdet, double coef, int F, double* out) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    double s = 0.0;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        double v = 0.0;
        if (terms) v += terms[(size_t)b * F + f];
        if (logdet) v -= coef * logdet[(size_t)b * F + f];
        s += v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        out[b] += t;
    }
}

// ------------------------------------------------------------------------------------------- host <-> device layouts
// host (B,C,F,T) complex{64,128} staged on the device -> X [B][F][C][Tp] complex64 (pad frames zeroed)
template <typename TIn>
__global__ void __launch_bounds__(256) import_x_kernel(const TIn* in, cf* X, int B, int C, int F, int T, int Tp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * F * C * Tp) return;
    const int t = (int)(idx % Tp);
    long long r = idx / Tp;
    const int c = (int)(r % C);
    r /= C;
    const int f = (int)(r % F);
    const int b = (int)(r / F);
    cf v = cf_make(0.f, 0.f);
    if (t < T) {
        const TIn s = in[(((size_t)b * C + c) * F + f) * T + t];
        v = cf_make((float)s.x, (float)s.y);
    }
    X[((size_t)b * F + f) * C * Tp + tile_off(C, Tp, c, t)] = v;
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

int launch_normalize_power(bss_handle* h, double2* W, cf* Wf, float* basis, const double* pw, int B, int N, int C, int F, int K,
                           double domain, double eps, double* aux_out) {
    power_aux_kernel<<<B * N, 256, 0, h->stream>>>(pw, aux_out, F, eps);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    const long long n = (long long)B * F * N * C + (basis ? (long long)B * N * F * K : 0);
    normalize_power_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(W, Wf, basis, aux_out, B, N, C, F, K, domain);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_normalize_pb(bss_handle* h, double2* W, cf* Wf, float* basis, const double2* scale, int B, int N, int C, int F, int K,
                        double domain) {
    const long long n = (long long)B * F * N;
    normalize_pb_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(W, Wf, basis, scale, B, N, C, F, K, domain);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_widen(bss_handle* h, const cf* in, double2* out, long long n) {
    widen_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(in, out, n);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_sync_wf(bss_handle* h, const double2* W, cf* Wf, long long n) {
    sync_wf_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(W, Wf, n);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_identity_filter(bss_handle* h, double2* W, cf* Wf, long long n_bins, int N, int C) {
    const long long n = n_bins * N * C;
    identity_filter_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(W, Wf, n, N, C);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

template <int C>
static int launch_separate_t(bss_handle* h, const cf* X, const cf* Wf, const double2* scale, cf* Y, cf* out, int B, int F, int T,
                             int Tp) {
    SepParams p;
    p.X = X;
    p.Wf = Wf;
    p.scale = scale;
    p.Y = Y;
    p.out = out;
    p.B = B;
    p.F = F;
    p.T = T;
    p.Tp = Tp;
    p.g = make_tile_geom(C, Tp);
    p.n_items = (long long)B * F;
    StreamPlan sp;
    if (!plan_stream(h, p.g, MU_STAGES, 16, (int)p.n_items, 8, &sp))
        return bss_fail(h, BSS_EINVAL, "separate: frame tile does not fit in shared memory");
    p.scratch_off = sp.scratch_off;
    p.scratch_stride = sp.scratch_stride;
    p.ring_off = sp.ring_off;
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(separate_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    separate_kernel<C><<<sp.grid, sp.wpc * 32, sp.smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_separate(bss_handle* h, const cf* X, const cf* Wf, const double2* scale, cf* Y, cf* out, int B, int C, int F, int T,
                    int Tp) {
    int rc = BSS_OK;
    BSS_DISPATCH_C(C, (rc = launch_separate_t<CC_>(h, X, Wf, scale, Y, out, B, F, T, Tp)))
    return rc;
}

int launch_export_y(bss_handle* h, const cf* Y, const double2* scale, cf* out, int B, int N, int F, int T, int Tp) {
    const long long n = (long long)B * N * F * T;
    export_y_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(Y, scale, out, B, N, F, T, Tp);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

template <int C, int KT, bool FROM_Y, bool CACHE>
static int launch_ilrma_loss_c(bss_handle* h, const LossParams& p, const StreamPlan& sp, size_t smem_bytes) {
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(ilrma_loss_kernel<C, KT, FROM_Y, CACHE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         h->max_smem));
        attr_done = true;
    }
    ilrma_loss_kernel<C, KT, FROM_Y, CACHE><<<sp.grid, sp.wpc * 32, smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

template <int C, int KT, bool FROM_Y>
static int launch_ilrma_loss_t(bss_handle* h, const MuArgs& a, float expo, double* terms) {
    LossParams p;
    p.a = a;
    p.out = terms;
    p.expo = expo;
    p.g = make_tile_geom(C, a.Tp);
    p.n_items = (long long)a.B * a.F;
    p.cache_off = 0;
    const size_t scratch = KT > 0 ? 16 : (size_t)C * a.K * 4;
    StreamPlan sp;
    size_t smem_bytes = 0;
    const bool cached = plan_stream_cached(h, p.g, MU_STAGES, scratch, p.n_items, 16, (size_t)C * a.K * a.Tp * sizeof(float), a.F, &sp,
                                           &p.cache_off, &smem_bytes);
    if (!cached && !plan_stream(h, p.g, MU_STAGES, scratch, p.n_items, 16, &sp))
        return bss_fail(h, BSS_EINVAL, "loss: frame tile does not fit in shared memory");
    p.scratch_off = sp.scratch_off;
    p.scratch_stride = sp.scratch_stride;
    p.ring_off = sp.ring_off;
    if (cached) return launch_ilrma_loss_c<C, KT, FROM_Y, true>(h, p, sp, smem_bytes);
    return launch_ilrma_loss_c<C, KT, FROM_Y, false>(h, p, sp, sp.smem_bytes);
}

int launch_ilrma_loss(bss_handle* h, const MuArgs& a, float expo, double* terms) {
    int rc = BSS_OK;
    if (a.K == 2) {
        if (a.Y) { BSS_DISPATCH_C(a.C, (rc = launch_ilrma_loss_t<CC_, 2, true>(h, a, expo, terms))) }
        else { BSS_DISPATCH_C(a.C, (rc = launch_ilrma_loss_t<CC_, 2, false>(h, a, expo, terms))) }
    } else {
        if (a.Y) { BSS_DISPATCH_C(a.C, (rc = launch_ilrma_loss_t<CC_, 0, true>(h, a, expo, terms))) }
        else { BSS_DISPATCH_C(a.C, (rc = launch_ilrma_loss_t<CC_, 0, false>(h, a, expo, terms))) }
    }
    return rc;
}

// hist[counter][b] = result[b]; ++counter -- the destination of a recorded loss lives in device memory so that the same launch
// can be replayed from a CUDA graph for every iteration (bss_run_record)
__global__ void loss_append_kernel(const double* result, double* hist, int* counter, int B, int capacity) {
    const int i = *counter;
    if (i < capacity && (int)threadIdx.x < B) hist[(size_t)i * B + threadIdx.x] = result[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) *counter = i + 1;
}

int launch_loss_append(bss_handle* h, const double* result, double* hist, int* counter, int B, int capacity) {
    if (B > 1024) return bss_fail(h, BSS_EINVAL, "loss history: at most 1024 mixtures per handle");
    loss_append_kernel<<<1, ((B + 31) / 32) * 32, 0, h->stream>>>(result, hist, counter, B, capacity);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_loss_finish(bss_handle* h, const double* terms, const double* logdet, double coef, int B, int F, double* out) {
    loss_finish_kernel<<<B, 256, 0, h->stream>>>(terms, logdet, coef, F, out);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_import_x(bss_handle* h, const void* staged, int dtype, cf* X, int B, int C, int F, int T, int Tp) {
    const long long n = (long long)B * F * C * Tp;
    const unsigned grid = (unsigned)cdiv(n, 256);
    if (dtype == BSS_C128)
        import_x_kernel<double2><<<grid, 256, 0, h->stream>>>((const double2*)staged, X, B, C, F, T, Tp);
    else
        import_x_kernel<float2><<<grid, 256, 0, h->stream>>>((const float2*)staged, X, B, C, F, T, Tp);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}
