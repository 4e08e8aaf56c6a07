// Auxiliary-function weights that need more than the low-rank model:
//   - AuxIVA frame weights r[n,t] (a reduction across ALL bins)          src/bss/iva.py:489-497, :722-730
//   - tILRMA weights xi = (nu R + 2 P) / (nu + 2)                        src/bss/ilrma.py:962-964
#include "handle.h"

namespace {

template <int C, bool FROM_Y>
__device__ __forceinline__ void power2(const float4 (&xv)[C], const cf* Wf, float (&P0)[C], float (&P1)[C]) {
#pragma unroll
    for (int n = 0; n < C; ++n) {
        if (FROM_Y) {
            P0[n] = fmaf(xv[n].x, xv[n].x, xv[n].y * xv[n].y);
            P1[n] = fmaf(xv[n].z, xv[n].z, xv[n].w * xv[n].w);
        } else {
            cf y0 = cf_make(0.f, 0.f), y1 = cf_make(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const cf w = __ldg(Wf + n * C + c);
                cf_fma(y0, w, cf_make(xv[c].x, xv[c].y));
                cf_fma(y1, w, cf_make(xv[c].z, xv[c].w));
            }
            P0[n] = cf_abs2(y0);
            P1[n] = cf_abs2(y1);
        }
    }
}

// Stage 1: a warp owns 64 frames of one mixture and sums |y|^2 over a chunk of bins.
// part: [B][n_chunks][N][Tp]
template <int C, bool FROM_Y>
__global__ void __launch_bounds__(128) frame_power_partial_kernel(const cf* src, const cf* Wf, float* part, int B, int F, int Tp,
                                                                 int n_chunks, int bins_per_chunk, int n_slabs,
                                                                 long long n_items) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (item >= n_items) return;
    long long r = item;
    const int slab = (int)(r % n_slabs);
    r /= n_slabs;
    const int chunk = (int)(r % n_chunks);
    const int b = (int)(r / n_chunks);
    const int t0 = slab * 64 + 2 * lane;
    if (t0 >= Tp) return;
    float s0[C], s1[C];
#pragma unroll
    for (int n = 0; n < C; ++n) s0[n] = s1[n] = 0.f;
    const int f_begin = chunk * bins_per_chunk;
    const int f_end = min(F, f_begin + bins_per_chunk);
    const size_t xoff = tile_off(C, Tp, 0, t0);
    const int xlen = (int)(tile_off(C, Tp, 1, t0) - xoff);
#pragma unroll 2
    for (int f = f_begin; f < f_end; ++f) {
        const size_t bf = (size_t)b * F + f;
        float4 xv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) xv[c] = __ldg(reinterpret_cast<const float4*>(src + bf * C * Tp + xoff + (size_t)c * xlen));
        float P0[C], P1[C];
        power2<C, FROM_Y>(xv, Wf + bf * C * C, P0, P1);
#pragma unroll
        for (int n = 0; n < C; ++n) {
            s0[n] += P0[n];
            s1[n] += P1[n];
        }
    }
#pragma unroll
    for (int n = 0; n < C; ++n)
        *reinterpret_cast<float2*>(part + (((size_t)b * n_chunks + chunk) * C + n) * Tp + t0) = make_float2(s0[n], s1[n]);
}

// Stage 2: r = sqrt(sum) (Laplace) or sum / F (Gauss); inverse of the floored value for the covariance
// kernel, raw value for the loss.  kind: 0 Laplace, 1 Gauss.  Block = 32 frames x 8 chunk lanes, fixed summation order.
__global__ void __launch_bounds__(256) frame_weight_finish_kernel(const float* part, float* winv, float* raw, int B, int N, int F,
                                                                 int T, int Tp, int n_chunks, int kind, float eps) {
    __shared__ float red[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int t_tiles = (Tp + 31) / 32;
    const long long bn = blockIdx.x / t_tiles;
    const int t = (int)(blockIdx.x % t_tiles) * 32 + tx;
    const int n = (int)(bn % N);
    const int b = (int)(bn / N);
    float s = 0.f;
    if (t < Tp)
        for (int c = ty; c < n_chunks; c += 8) s += part[(((size_t)b * n_chunks + c) * N + n) * Tp + t];
    red[ty][tx] = s;
    __syncthreads();
    if (ty != 0 || t >= Tp) return;
    const size_t idx = (size_t)bn * Tp + t;
    if (t >= T) {
        if (winv) winv[idx] = 1.f;
        if (raw) raw[idx] = 0.f;
        return;
    }
    s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += red[y][tx];
    const float r = kind == 0 ? sqrtf(s) : s / (float)F;
    if (raw) raw[idx] = r;
    if (winv) winv[idx] = __frcp_rn(r < eps ? eps : r);
}

constexpr int TW_STAGES = 3;
struct TwParams {
    const cf* X;
    const cf* Wf;
    const float* basis;
    const float* act;
    float* iw;   // [B][F][N][Tp]
    int B, F, K, Tp;
    float nu, eps;
    TileGeom g;
    long long n_items;
    uint32_t scratch_off, scratch_stride, ring_off;
};

template <int C>
__global__ void __launch_bounds__(256) t_weights_kernel(const TwParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    constexpr int N = C;
    const int K = p.K;
    float* tb = reinterpret_cast<float*>(smem + p.scratch_off + (size_t)warp * p.scratch_stride);
    WarpStream<TW_STAGES> st;
    st.start(p.g, reinterpret_cast<uint64_t*>(smem) + warp * TW_STAGES,
             smem + p.ring_off + (size_t)warp * TW_STAGES * p.g.stage_bytes, p.X, (int)(blockIdx.x * wpc + warp),
             (int)(gridDim.x * wpc), (int)p.n_items, 1, lane);
#pragma unroll 1
    while (st.active()) {
        st.issue_next(p.g, p.X, 1);
        const int bf = st.cons.item;
        const int b = bf / p.F, f = bf - b * p.F;
        if (st.first_slab()) {
            for (int i = lane; i < N * K; i += 32) {
                const int n = i / K, k = i - n * K;
                tb[i] = p.basis[(((size_t)b * N + n) * p.F + f) * K + k];
            }
            __syncwarp();
        }
        const cf* xs = st.acquire(p.g);
        const int nf = st.frames(p.g);
        const int tbase = st.frame0(p.g);
#pragma unroll 1
        for (int tt = 2 * lane; tt < nf; tt += 64) {
            float4 xv[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (size_t)c * nf + tt);
            float P0[C], P1[C];
            power2<C, false>(xv, p.Wf + (size_t)bf * C * C, P0, P1);
            const int t = tbase + tt;
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float* v = p.act + ((size_t)b * N + n) * K * p.Tp + t;
                float r0 = 0.f, r1 = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float2 vv = __ldg(reinterpret_cast<const float2*>(v + (size_t)k * p.Tp));
                    const float tk = tb[n * K + k];
                    r0 = fmaf(tk, vv.x, r0);
                    r1 = fmaf(tk, vv.y, r1);
                }
                r0 = r0 < p.eps ? p.eps : r0;
                r1 = r1 < p.eps ? p.eps : r1;
                const float x0 = (p.nu * r0 + 2.f * P0[n]) / (p.nu + 2.f);
                const float x1 = (p.nu * r1 + 2.f * P1[n]) / (p.nu + 2.f);
                *reinterpret_cast<float2*>(p.iw + ((size_t)bf * N + n) * p.Tp + t) = make_float2(1.f / x0, 1.f / x1);
            }
        }
        st.release(p.g);
    }
}

// host (N,F,T) float64 weights (already floored) -> iw [F][N][Tp] inverse fp32 (B = 1 primitive path)
__global__ void __launch_bounds__(256) import_weights_kernel(const double* r, float* iw, int N, int F, int T, int Tp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)F * N * Tp) return;
    const int t = (int)(idx % Tp);
    const long long fn = idx / Tp;
    const int n = (int)(fn % N);
    const int f = (int)(fn / N);
    iw[idx] = t < T ? (float)(1.0 / r[((size_t)n * F + f) * T + t]) : 0.f;
}

// GaussIDLMA: staged host (B,N,F,T) float64 variances R -> iw [B][F][N][Tp] = 1 / max(R, eps)
// (the floor of src/sss/idlma.py:190; the pad frame gets weight 0)
__global__ void __launch_bounds__(256) import_variance_kernel(const double* r, float* iw, int B, int N, int F, int T, int Tp,
                                                              double eps) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * F * N * Tp) return;
    const int t = (int)(idx % Tp);
    long long q = idx / Tp;
    const int n = (int)(q % N);
    q /= N;
    const int f = (int)(q % F);
    const int b = (int)(q / F);
    float v = 0.f;
    if (t < T) {
        double x = r[(((size_t)b * N + n) * F + f) * T + t];
        x = x < eps ? eps : x;
        v = (float)(1.0 / x);
    }
    iw[idx] = v;
}

// GaussIDLMA loss terms (src/sss/idlma.py:246-258): terms[bf] = sum_n sum_{t<T} (|y_n|^2 / R + log R), y = W_f x,
// one warp per bin; R is held as its inverse
__global__ void __launch_bounds__(128) idlma_loss_kernel(const cf* X, const cf* Wf, const float* iw, double* terms,
                                                         long long n_items, int C, int T, int Tp) {
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= n_items) return;
    const int lane = threadIdx.x & 31;
    const cf* x = X + (size_t)item * C * Tp;
    const cf* w = Wf + (size_t)item * C * C;
    const float* r = iw + (size_t)item * C * Tp;
    double s = 0.0;
    for (int t = lane; t < T; t += 32) {
        cf xv[8];
        for (int c = 0; c < C; ++c) xv[c] = x[tile_off(C, Tp, c, t)];
        for (int n = 0; n < C; ++n) {
            float yr = 0.f, yi = 0.f;
            for (int c = 0; c < C; ++c) {
                const cf wv = __ldg(w + n * C + c);
                yr = fmaf(wv.x, xv[c].x, fmaf(-wv.y, xv[c].y, yr));
                yi = fmaf(wv.x, xv[c].y, fmaf(wv.y, xv[c].x, yi));
            }
            const float inv = r[(size_t)n * Tp + t];
            s += (double)((yr * yr + yi * yi) * inv) - log((double)inv);
        }
    }
    s = warp_sum(s);
    if (lane == 0) terms[item] = s;
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

// src: X (from_y = 0, demixed with Wf on the fly) or the Y state (from_y = 1)
int launch_frame_weights(bss_handle* h, const cf* src, const cf* Wf, int from_y, float* winv, float* raw, int B, int C, int F,
                         int T, int Tp, int kind, float eps) {
    const int n_slabs = (Tp + 63) / 64;
    long long want = (long long)h->n_sm * 24;
    int n_chunks = (int)cdiv(want, (long long)B * n_slabs);
    if (n_chunks < 1) n_chunks = 1;
    int bins_per_chunk = (int)cdiv(F, n_chunks);
    if (bins_per_chunk < 4) bins_per_chunk = F < 4 ? F : 4;
    n_chunks = (int)cdiv(F, bins_per_chunk);
    const size_t need = (size_t)B * n_chunks * C * Tp;
    if (need > h->part_elems) {
        if (h->part) cudaFree(h->part);
        h->part = nullptr;
        BSS_CUDA(h, cudaMalloc(&h->part, need * sizeof(float)));
        h->part_elems = need;
    }
    const long long n_items = (long long)B * n_chunks * n_slabs;
    const unsigned grid = (unsigned)cdiv(n_items, 4);
    if (from_y) {
        BSS_DISPATCH_C(C, (frame_power_partial_kernel<CC_, true><<<grid, 128, 0, h->stream>>>(
                              src, Wf, h->part, B, F, Tp, n_chunks, bins_per_chunk, n_slabs, n_items)))
    } else {
        BSS_DISPATCH_C(C, (frame_power_partial_kernel<CC_, false><<<grid, 128, 0, h->stream>>>(
                              src, Wf, h->part, B, F, Tp, n_chunks, bins_per_chunk, n_slabs, n_items)))
    }
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    const long long blocks = (long long)B * C * ((Tp + 31) / 32);
    frame_weight_finish_kernel<<<(unsigned)blocks, 256, 0, h->stream>>>(h->part, winv, raw, B, C, F, T, Tp, n_chunks, kind, eps);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

template <int C>
static int launch_t_weights_t(bss_handle* h, const cf* X, const cf* Wf, const float* basis, const float* act, float* iw, int B,
                              int F, int K, int Tp, float nu, float eps) {
    TwParams p;
    p.X = X;
    p.Wf = Wf;
    p.basis = basis;
    p.act = act;
    p.iw = iw;
    p.B = B;
    p.F = F;
    p.K = K;
    p.Tp = Tp;
    p.nu = nu;
    p.eps = eps;
    p.g = make_tile_geom(C, Tp);
    p.n_items = (long long)B * F;
    StreamPlan sp;
    if (!plan_stream(h, p.g, TW_STAGES, (size_t)C * K * 4, (int)p.n_items, 8, &sp))
        return bss_fail(h, BSS_EINVAL, "t weights: frame tile does not fit in shared memory");
    p.scratch_off = sp.scratch_off;
    p.scratch_stride = sp.scratch_stride;
    p.ring_off = sp.ring_off;
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(t_weights_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    t_weights_kernel<C><<<sp.grid, sp.wpc * 32, sp.smem_bytes, h->stream>>>(p);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_t_weights(bss_handle* h, const cf* X, const cf* Wf, const float* basis, const float* act, float* iw, int B, int C,
                     int F, int K, int Tp, float nu, float eps) {
    int rc = BSS_OK;
    BSS_DISPATCH_C(C, (rc = launch_t_weights_t<CC_>(h, X, Wf, basis, act, iw, B, F, K, Tp, nu, eps)))
    return rc;
}

int launch_import_weights(bss_handle* h, const double* r_dev, float* iw, int N, int F, int T, int Tp) {
    const long long n = (long long)F * N * Tp;
    import_weights_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(r_dev, iw, N, F, T, Tp);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_import_variance(bss_handle* h, const double* r_dev, float* iw, int B, int N, int F, int T, int Tp, double eps) {
    const long long n = (long long)B * F * N * Tp;
    import_variance_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(r_dev, iw, B, N, F, T, Tp, eps);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_idlma_loss(bss_handle* h, const cf* X, const cf* Wf, const float* iw, double* terms, int B, int C, int F, int T, int Tp) {
    const long long n_items = (long long)B * F;
    idlma_loss_kernel<<<(unsigned)cdiv(n_items, 4), 128, 0, h->stream>>>(X, Wf, iw, terms, n_items, C, T, Tp);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

// AuxIVA loss terms: out[b] += coef * sum_{n, t < T} g(raw[b,n,t]); g = identity (kind 0, Laplace,
// src/bss/iva.py:617) or log(max(., eps)) (kind 1, Gauss, src/bss/iva.py:798-800)
__global__ void __launch_bounds__(256) sum_frames_kernel(const float* raw, int N, int T, int Tp, int kind, double coef, double eps,
                                                         double* out) {
    __shared__ double red[8];
    const int b = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < N * Tp; i += blockDim.x) {
        const int t = i % Tp;
        if (t >= T) continue;
        const double r = (double)raw[(size_t)b * N * Tp + i];
        s += kind == 0 ? r : log(r < eps ? eps : r);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        out[b] += coef * t;
    }
}

int launch_sum_frames(bss_handle* h, const float* raw, int B, int N, int T, int Tp, int kind, double coef, double eps, double* out) {
    sum_frames_kernel<<<B, 256, 0, h->stream>>>(raw, N, T, Tp, kind, coef, eps, out);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}
