// Iterative source steering: rank-1 updates applied directly to the estimates Y.
//   src/bss/ilrma.py:537-564, src/bss/iva.py:525-542 and :758-775
// A warp stages one bin tile Y[f] (N rows x Tp frames) in shared memory with one bulk copy, runs
// ALL N sequential steering steps on it there (the reference re-reads and re-writes the whole
// (N,F,T) tensor N times per iteration) and writes it back with one bulk store.  Each lane owns
// a fixed set of frames, so the tile needs no intra-warp synchronisation between steps.
#include "handle.h"

namespace {

struct IssParams {
    cf* Y;                 // [B][F][N][Tp]
    const float* basis;    // ILRMA weights (mode 0)
    const float* act;
    const float* wfr;      // AuxIVA inverse frame weights [B][N][Tp] (mode 1)
    double* pw;            // [B][N][F] mean_t |y|^2 after the sweep (may be null)
    int B, F, T, Tp, K;
    int mode;
    float expo, eps;
    long long n_items;
    uint32_t tile_bytes, warp_bytes;
};

template <int N>
__global__ void __launch_bounds__(256) iss_kernel(const IssParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    // per warp: [2] barriers (16 B, padded to 128) | rinv [N][Tp] floats | 2 tiles
    unsigned char* base = smem + (size_t)warp * p.warp_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base);
    float* rinv = reinterpret_cast<float*>(base + 128);
    const uint32_t rinv_bytes = (uint32_t)(((size_t)N * p.Tp * 4 + 127) / 128 * 128);
    unsigned char* tiles = base + 128 + rinv_bytes;
    const int Tp = p.Tp;

    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const long long stride = (long long)gridDim.x * wpc;
    long long item = (long long)blockIdx.x * wpc + warp;
    int buf = 0;
    uint32_t phase0 = 0u, phase1 = 0u;
    if (item < p.n_items && lane == 0) {
        mbar_expect_tx(&bars[0], p.tile_bytes);
        bulk_g2s(tiles, p.Y + (size_t)item * N * Tp, p.tile_bytes, &bars[0]);
    }
#pragma unroll 1
    for (; item < p.n_items; item += stride, buf ^= 1) {
        const long long next = item + stride;
        if (next < p.n_items && lane == 0) {
            // the other buffer was handed to a bulk store two items ago: its reads must be done
            bulk_wait_read0();
            mbar_expect_tx(&bars[buf ^ 1], p.tile_bytes);
            bulk_g2s(tiles + (size_t)(buf ^ 1) * p.tile_bytes, p.Y + (size_t)next * N * Tp, p.tile_bytes, &bars[buf ^ 1]);
        }
        const int b = (int)(item / p.F), f = (int)(item - (long long)b * p.F);
        // inverse weights of this bin
        if (p.mode == 0) {
            for (int n = 0; n < N; ++n) {
                const float* tb = p.basis + (((size_t)b * N + n) * p.F + f) * p.K;
                const float* v = p.act + ((size_t)b * N + n) * p.K * Tp;
                for (int t = 2 * lane; t < Tp; t += 64) {
                    float r0 = 0.f, r1 = 0.f;
                    for (int k = 0; k < p.K; ++k) {
                        const float2 vv = __ldg(reinterpret_cast<const float2*>(v + (size_t)k * Tp + t));
                        const float tk = __ldg(tb + k);
                        r0 = fmaf(tk, vv.x, r0);
                        r1 = fmaf(tk, vv.y, r1);
                    }
                    if (p.expo != 1.f) {
                        r0 = powf(r0, p.expo);
                        r1 = powf(r1, p.expo);
                    }
                    r0 = r0 < p.eps ? p.eps : r0;
                    r1 = r1 < p.eps ? p.eps : r1;
                    *reinterpret_cast<float2*>(rinv + (size_t)n * Tp + t) = make_float2(__frcp_rn(r0), __frcp_rn(r1));
                }
            }
        } else {
            for (int n = 0; n < N; ++n)
                for (int t = 2 * lane; t < Tp; t += 64)
                    *reinterpret_cast<float2*>(rinv + (size_t)n * Tp + t) =
                        __ldg(reinterpret_cast<const float2*>(p.wfr + ((size_t)b * N + n) * Tp + t));
        }
        mbar_wait(&bars[buf], buf ? phase1 : phase0);
        if (buf)
            phase1 ^= 1u;
        else
            phase0 ^= 1u;
        cf* y = reinterpret_cast<cf*>(tiles + (size_t)buf * p.tile_bytes);

#pragma unroll 1
        for (int n = 0; n < N; ++n) {
            float ur[N], ui[N], d[N];
#pragma unroll
            for (int k = 0; k < N; ++k) ur[k] = ui[k] = d[k] = 0.f;
            for (int t = 2 * lane; t < Tp; t += 64) {
                const float4 yn = *reinterpret_cast<const float4*>(y + tile_off(N, Tp, n, t));
                const float p0 = fmaf(yn.x, yn.x, yn.y * yn.y), p1 = fmaf(yn.z, yn.z, yn.w * yn.w);
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const float4 yk = *reinterpret_cast<const float4*>(y + tile_off(N, Tp, k, t));
                    const float2 ri = *reinterpret_cast<const float2*>(rinv + (size_t)k * Tp + t);
                    // y_k conj(y_n) / r_k
                    ur[k] = fmaf(ri.x, fmaf(yk.x, yn.x, yk.y * yn.y), ur[k]);
                    ui[k] = fmaf(ri.x, fmaf(yk.y, yn.x, -yk.x * yn.y), ui[k]);
                    ur[k] = fmaf(ri.y, fmaf(yk.z, yn.z, yk.w * yn.w), ur[k]);
                    ui[k] = fmaf(ri.y, fmaf(yk.w, yn.z, -yk.z * yn.w), ui[k]);
                    d[k] = fmaf(ri.x, p0, d[k]);
                    d[k] = fmaf(ri.y, p1, d[k]);
                }
            }
            float vr[N], vi[N];
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const double sur = warp_sum((double)ur[k]);
                const double sui = warp_sum((double)ui[k]);
                const double sd = warp_sum((double)d[k]);
                if (k == n) {
                    vr[k] = (float)(1.0 - 1.0 / sqrt(sd));
                    vi[k] = 0.f;
                } else {
                    vr[k] = (float)(sur / sd);
                    vi[k] = (float)(sui / sd);
                }
            }
            for (int t = 2 * lane; t < Tp; t += 64) {
                const float4 yn = *reinterpret_cast<const float4*>(y + tile_off(N, Tp, n, t));
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    float4 yk = *reinterpret_cast<float4*>(y + tile_off(N, Tp, k, t));
                    // y_k -= v_k y_n
                    yk.x -= vr[k] * yn.x - vi[k] * yn.y;
                    yk.y -= vr[k] * yn.y + vi[k] * yn.x;
                    yk.z -= vr[k] * yn.z - vi[k] * yn.w;
                    yk.w -= vr[k] * yn.w + vi[k] * yn.z;
                    *reinterpret_cast<float4*>(y + tile_off(N, Tp, k, t)) = yk;
                }
            }
        }
        if (p.pw) {
#pragma unroll
            for (int n = 0; n < N; ++n) {
                float s = 0.f;
                for (int t = 2 * lane; t < Tp; t += 64) {
                    const float4 yn = *reinterpret_cast<const float4*>(y + tile_off(N, Tp, n, t));
                    s += fmaf(yn.x, yn.x, yn.y * yn.y) + fmaf(yn.z, yn.z, yn.w * yn.w);
                }
                const double tot = warp_sum((double)s);
                if (lane == 0) p.pw[((size_t)b * N + n) * p.F + f] = tot / (double)p.T;
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_s2g(p.Y + (size_t)item * N * Tp, y, p.tile_bytes);
            bulk_commit();
        }
    }
    if (lane == 0) bulk_wait0();
}

// G[n][c] = (1/T) sum_t y_n conj(x_c): one warp per bin, rows read straight from global memory
template <int C>
__global__ void __launch_bounds__(128) cross_cov_kernel(const cf* Y, const cf* X, double2* G, long long n_bins, int T, int Tp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bf = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (bf >= n_bins) return;
    float gr[C][C], gi[C][C];
#pragma unroll
    for (int n = 0; n < C; ++n)
#pragma unroll
        for (int c = 0; c < C; ++c) gr[n][c] = gi[n][c] = 0.f;
    for (int t = 2 * lane; t < Tp; t += 64) {
        float4 xv[C], yv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            xv[c] = __ldg(reinterpret_cast<const float4*>(X + (size_t)bf * C * Tp + tile_off(C, Tp, c, t)));
            yv[c] = __ldg(reinterpret_cast<const float4*>(Y + (size_t)bf * C * Tp + tile_off(C, Tp, c, t)));
        }
#pragma unroll
        for (int n = 0; n < C; ++n)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                gr[n][c] += fmaf(yv[n].x, xv[c].x, yv[n].y * xv[c].y) + fmaf(yv[n].z, xv[c].z, yv[n].w * xv[c].w);
                gi[n][c] += fmaf(yv[n].y, xv[c].x, -yv[n].x * xv[c].y) + fmaf(yv[n].w, xv[c].z, -yv[n].z * xv[c].w);
            }
    }
#pragma unroll
    for (int n = 0; n < C; ++n)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const double sr = warp_sum((double)gr[n][c]);
            const double si = warp_sum((double)gi[n][c]);
            if (lane == 0) G[(size_t)bf * C * C + n * C + c] = make_double2(sr / (double)T, si / (double)T);
        }
}

// Y[b,f,n,:] *= s with s = 1/aux[b,n] (real, power normalisation) or scale[b,n,f] (complex, projection back);
// the basis follows with |s|^domain.
__global__ void __launch_bounds__(256) scale_y_kernel(cf* Y, float* basis, const double* aux, const double2* scale, int B, int N,
                                                      int F, int Tp, int K, double domain) {
    const long long row = blockIdx.x;   // (b, f, n)
    const int n = (int)(row % N);
    const long long bf = row / N;
    const int f = (int)(bf % F);
    const int b = (int)(bf / F);
    float sx, sy;
    double mag;
    if (aux) {
        const double a = aux[(size_t)b * N + n];
        sx = (float)(1.0 / a);
        sy = 0.f;
        mag = 1.0 / a;
    } else {
        const double2 s = scale[((size_t)b * N + n) * F + f];
        sx = (float)s.x;
        sy = (float)s.y;
        mag = hypot(s.x, s.y);
    }
    cf* tile = Y + (size_t)bf * N * Tp;
    for (int t = threadIdx.x; t < Tp; t += blockDim.x) {
        cf* y = tile + tile_off(N, Tp, n, t);
        const cf v = *y;
        *y = cf_make(v.x * sx - v.y * sy, v.x * sy + v.y * sx);
    }
    if (basis && threadIdx.x < K) {
        const double sc = domain == 2.0 ? mag * mag : pow(mag, domain);
        float* p = basis + (((size_t)b * N + n) * F + f) * K + threadIdx.x;
        *p = (float)((double)*p * sc);
    }
}

// aux[b,n] = max(sqrt(mean_f pw[b,n,f]), eps)
__global__ void __launch_bounds__(256) aux_from_power_kernel(const double* pw, double* aux, int F, double eps) {
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
        double a = sqrt(t / (double)F);
        aux[bn] = a < eps ? eps : a;
    }
}

template <int N>
int launch_iss_t(bss_handle* h, IssParams p) {
    p.tile_bytes = (uint32_t)((size_t)N * p.Tp * 8);
    const uint32_t rinv_bytes = (uint32_t)(((size_t)N * p.Tp * 4 + 127) / 128 * 128);
    // tile_bytes is a multiple of 16 (Tp is even), which is all bulk copies and LDS.128 need
    p.warp_bytes = (128 + rinv_bytes + 2 * p.tile_bytes + 127) / 128 * 128;
    int wpc = (int)(((size_t)h->max_smem - 256) / p.warp_bytes);
    if (wpc > 8) wpc = 8;
    if (wpc < 1) return bss_fail(h, BSS_EINVAL, "ISS: frame tile does not fit in shared memory");
    const size_t smem_bytes = (size_t)wpc * p.warp_bytes;
    static bool attr_done = false;
    if (!attr_done) {
        BSS_CUDA(h, cudaFuncSetAttribute(iss_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem));
        attr_done = true;
    }
    long long grid = cdiv(p.n_items, wpc);
    int ctas = (int)((size_t)h->max_smem / (smem_bytes + 1024));
    if (ctas < 1) ctas = 1;
    if (ctas > 4) ctas = 4;
    if (grid > (long long)h->n_sm * ctas) grid = (long long)h->n_sm * ctas;
    iss_kernel<N><<<(unsigned)grid, wpc * 32, smem_bytes, h->stream>>>(p);
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

// mode 0: ILRMA weights from (basis, act, expo); mode 1: AuxIVA inverse frame weights wfr
int launch_iss(bss_handle* h, cf* Y, int mode, const float* basis, const float* act, const float* wfr, double* pw, int B, int N,
               int F, int T, int Tp, int K, float expo, float eps) {
    IssParams p{};
    p.Y = Y;
    p.basis = basis;
    p.act = act;
    p.wfr = wfr;
    p.pw = pw;
    p.B = B;
    p.F = F;
    p.T = T;
    p.Tp = Tp;
    p.K = K;
    p.mode = mode;
    p.expo = expo;
    p.eps = eps;
    p.n_items = (long long)B * F;
    int rc = BSS_OK;
    BSS_DISPATCH_C(N, (rc = launch_iss_t<CC_>(h, p)))
    return rc;
}

int launch_cross_cov(bss_handle* h, const cf* Y, const cf* X, double2* G, long long n_bins, int C, int T, int Tp) {
    const unsigned grid = (unsigned)cdiv(n_bins, 4);
    BSS_DISPATCH_C(C, (cross_cov_kernel<CC_><<<grid, 128, 0, h->stream>>>(Y, X, G, n_bins, T, Tp)))
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_scale_y(bss_handle* h, cf* Y, float* basis, const double* aux, const double2* scale, int B, int N, int F, int Tp, int K,
                   double domain) {
    const long long rows = (long long)B * F * N;
    scale_y_kernel<<<(unsigned)rows, 256, 0, h->stream>>>(Y, basis, aux, scale, B, N, F, Tp, K, domain);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_aux_from_power(bss_handle* h, const double* pw, double* aux, int B, int N, int F, double eps) {
    aux_from_power_kernel<<<B * N, 256, 0, h->stream>>>(pw, aux, F, eps);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}
