// Sawada's multichannel IS-NMF (src/bss/mnmf.py:116-635, author='Sawada'), fp64.
//
// Model: X_hat[f,t] = sum_n lambda[n,f,t] H[f,n],  lambda[n,f,t] = sum_k Z[n,k] T[f,k] V[k,t]   (mnmf.py:554-562).
// The reference materialises X = x x^H (F,T,C,C), X_hat, its inverse and the product X_hat^-1 X X_hat^-1 for each of
// the four multiplicative updates.  Here a thread owns one (bin, frame): it builds X_hat in registers from the bin's
// N Hermitian matrices (shared memory), inverts it by Cholesky, and -- because X is rank one -- gets everything the
// updates need from q = X_hat^-1 x:  tr(X_hat^-1 X X_hat^-1 H_n) = q^H H_n q,  tr(X_hat^-1 H_n) = <X_hat^-1, H_n>.
// Nothing of size (F,T,C,C) is ever stored.  Hermitian matrices are packed as in handle.h (C real diagonals, then the
// strict lower triangle row by row as (re, im)).
#include "handle.h"
#include "smallmat.cuh"

namespace {

constexpr int TB = 128;       // frames per CTA
constexpr int NMAX = 8;       // sources
constexpr int KGROUP = 8;     // basis functions per pass of the activation reduction

struct SmArgs {
    const cf* X;      // [B][F] bin tiles
    double* H;        // [B][F][N][C*C] packed Hermitian
    double* Z;        // [B][N][K]
    double* T;        // [B][F][K]
    double* V;        // [B][K][Tn]
    int B, N, F, Tn, Tp, K;
    double eps;
};

__host__ __device__ constexpr int tri(int i, int j) { return i * (i - 1) / 2 + j; }   // i > j

template <int C>
struct Herm {
    double d[C];
    double2 o[C * (C - 1) / 2];
};

// S = sum_n lam[n] H_n (shared memory, broadcast reads)
template <int C>
__device__ __forceinline__ void model_covariance(const double* Hs, const double* lam, int N, Herm<C>& S) {
#pragma unroll
    for (int i = 0; i < C; ++i) S.d[i] = 0.0;
#pragma unroll
    for (int e = 0; e < C * (C - 1) / 2; ++e) S.o[e] = make_double2(0.0, 0.0);
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
        if (n < N) {
            const double* p = Hs + n * C * C;
            const double l = lam[n];
#pragma unroll
            for (int i = 0; i < C; ++i) S.d[i] = fma(l, p[i], S.d[i]);
#pragma unroll
            for (int e = 0; e < C * (C - 1) / 2; ++e) {
                S.o[e].x = fma(l, p[C + 2 * e], S.o[e].x);
                S.o[e].y = fma(l, p[C + 2 * e + 1], S.o[e].y);
            }
        }
    }
}

// Inv = S^-1 through S = L L^H; returns log det S.  A matrix that is not positive definite yields NaNs.
template <int C>
__device__ __forceinline__ double herm_inverse(const Herm<C>& S, Herm<C>& Inv) {
    double2 L[C][C];   // lower triangle; the diagonal holds (l, 1/l)
    double logdet = 0.0;
#pragma unroll
    for (int j = 0; j < C; ++j) {
        double d = S.d[j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k].x * L[j][k].x + L[j][k].y * L[j][k].y;
        const double l = sqrt(d), il = 1.0 / l;
        logdet += log(d);
        L[j][j] = make_double2(l, il);
#pragma unroll
        for (int i = j + 1; i < C; ++i) {
            double2 s = S.o[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) {   // s -= L[i][k] conj(L[j][k])
                s.x -= L[i][k].x * L[j][k].x + L[i][k].y * L[j][k].y;
                s.y -= L[i][k].y * L[j][k].x - L[i][k].x * L[j][k].y;
            }
            L[i][j] = make_double2(s.x * il, s.y * il);
        }
    }
    double2 M[C][C];   // M = L^-1, lower triangle, real diagonal
#pragma unroll
    for (int j = 0; j < C; ++j) {
        M[j][j] = make_double2(L[j][j].y, 0.0);
#pragma unroll
        for (int i = j + 1; i < C; ++i) {
            double2 s = make_double2(L[i][j].x * M[j][j].x, L[i][j].y * M[j][j].x);
#pragma unroll
            for (int k = j + 1; k < i; ++k) {
                s.x += L[i][k].x * M[k][j].x - L[i][k].y * M[k][j].y;
                s.y += L[i][k].x * M[k][j].y + L[i][k].y * M[k][j].x;
            }
            M[i][j] = make_double2(-s.x * L[i][i].y, -s.y * L[i][i].y);
        }
    }
    // Inv = M^H M
#pragma unroll
    for (int i = 0; i < C; ++i) {
        double d = 0.0;
#pragma unroll
        for (int k = i; k < C; ++k) d += M[k][i].x * M[k][i].x + M[k][i].y * M[k][i].y;
        Inv.d[i] = d;
#pragma unroll
        for (int j = 0; j < i; ++j) {
            double2 s = make_double2(0.0, 0.0);
#pragma unroll
            for (int k = i; k < C; ++k) {   // conj(M[k][i]) M[k][j]
                s.x += M[k][i].x * M[k][j].x + M[k][i].y * M[k][j].y;
                s.y += M[k][i].x * M[k][j].y - M[k][i].y * M[k][j].x;
            }
            Inv.o[tri(i, j)] = s;
        }
    }
    return logdet;
}

// y = A x for Hermitian A
template <int C>
__device__ __forceinline__ void herm_matvec(const Herm<C>& A, const double2* x, double2* y) {
#pragma unroll
    for (int i = 0; i < C; ++i) {
        double2 s = make_double2(A.d[i] * x[i].x, A.d[i] * x[i].y);
#pragma unroll
        for (int j = 0; j < C; ++j) {
            if (j < i) {
                const double2 a = A.o[tri(i, j)];
                s.x += a.x * x[j].x - a.y * x[j].y;
                s.y += a.x * x[j].y + a.y * x[j].x;
            } else if (j > i) {
                const double2 a = A.o[tri(j, i)];   // conj
                s.x += a.x * x[j].x + a.y * x[j].y;
                s.y += a.x * x[j].y - a.y * x[j].x;
            }
        }
        y[i] = s;
    }
}

// q^H H q and <Inv, H> = tr(Inv H) for a packed Hermitian H in shared memory
template <int C>
__device__ __forceinline__ void traces(const double* p, const Herm<C>& Inv, const double2* q, double& quad, double& tr) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        a = fma(p[i], q[i].x * q[i].x + q[i].y * q[i].y, a);
        b = fma(p[i], Inv.d[i], b);
    }
#pragma unroll
    for (int i = 1; i < C; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) {
            const double hr = p[C + 2 * tri(i, j)], hi = p[C + 2 * tri(i, j) + 1];
            // Re(conj(q_i) H_ij q_j)
            const double gr = q[i].x * q[j].x + q[i].y * q[j].y;    // Re(conj(q_i) q_j)
            const double gi = q[i].x * q[j].y - q[i].y * q[j].x;    // Im(conj(q_i) q_j)
            a += 2.0 * (hr * gr - hi * gi);
            b += 2.0 * (Inv.o[tri(i, j)].x * hr + Inv.o[tri(i, j)].y * hi);
        }
    quad = a;
    tr = b;
}

template <int C>
__device__ __forceinline__ void load_frame(const cf* tile, int Tp, int t, double2* x) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const cf v = tile[tile_off(C, Tp, c, t)];
        x[c] = make_double2((double)v.x, (double)v.y);
    }
}

// shared memory of a bin: Hs[N][C*C], zt[N][K] = Z[n,k] T[f,k]
template <int C>
__device__ __forceinline__ void load_bin(const SmArgs& a, int b, int f, double* Hs, double* zt) {
    const double* Hg = a.H + ((size_t)b * a.F + f) * a.N * C * C;
    for (int i = threadIdx.x; i < a.N * C * C; i += blockDim.x) Hs[i] = Hg[i];
    const double* Zg = a.Z + (size_t)b * a.N * a.K;
    const double* Tg = a.T + ((size_t)b * a.F + f) * a.K;
    for (int i = threadIdx.x; i < a.N * a.K; i += blockDim.x) zt[i] = Zg[i] * Tg[i % a.K];
    __syncthreads();
}

__device__ __forceinline__ void source_power(const SmArgs& a, int b, int t, const double* zt, double* lam) {
#pragma unroll
    for (int n = 0; n < NMAX; ++n) lam[n] = 0.0;
    const double* Vg = a.V + (size_t)b * a.K * a.Tn + t;
    for (int k = 0; k < a.K; ++k) {
        const double v = Vg[(size_t)k * a.Tn];
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
            if (n < a.N) lam[n] = fma(zt[n * a.K + k], v, lam[n]);
    }
}

// ------------------------------------------------------------------------------------------- trace statistics
// num[b][n][f][t] = tr(Xh^-1 X Xh^-1 H_n), den[b][n][f][t] = tr(Xh^-1 H_n)      mnmf.py:392-398 (and :416,:441)
template <int C>
__global__ void __launch_bounds__(TB) smnmf_stats_kernel(SmArgs a, double* num, double* den) {
    extern __shared__ double sm[];
    double* Hs = sm;
    double* zt = Hs + a.N * C * C;
    const int f = blockIdx.y, b = blockIdx.z, t = blockIdx.x * TB + threadIdx.x;
    load_bin<C>(a, b, f, Hs, zt);
    if (t >= a.Tn) return;
    double lam[NMAX];
    source_power(a, b, t, zt, lam);
    Herm<C> S, Inv;
    model_covariance<C>(Hs, lam, a.N, S);
#pragma unroll
    for (int i = 0; i < C; ++i) S.d[i] += a.eps;
    herm_inverse<C>(S, Inv);
    double2 x[C], q[C];
    load_frame<C>(a.X + ((size_t)b * a.F + f) * C * a.Tp, a.Tp, t, x);
    herm_matvec<C>(Inv, x, q);
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
        if (n < a.N) {
            double qa, tr;
            traces<C>(Hs + n * C * C, Inv, q, qa, tr);
            const size_t o = (((size_t)b * a.N + n) * a.F + f) * a.Tn + t;
            num[o] = qa;
            den[o] = tr;
        }
    }
}

// P[b][f][n][k] = sum_t V[k,t] stat[n,f,t] for both statistics (shared by the basis and the latent update)
__global__ void __launch_bounds__(TB) smnmf_nk_kernel(SmArgs a, const double* num, const double* den, double* Pn, double* Pd) {
    const int f = blockIdx.x, b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int idx = warp; idx < a.N * a.K; idx += TB / 32) {
        const int n = idx / a.K, k = idx % a.K;
        const double* v = a.V + ((size_t)b * a.K + k) * a.Tn;
        const size_t o = (((size_t)b * a.N + n) * a.F + f) * a.Tn;
        double s0 = 0.0, s1 = 0.0;
        for (int t = lane; t < a.Tn; t += 32) {
            s0 = fma(v[t], num[o + t], s0);
            s1 = fma(v[t], den[o + t], s1);
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (lane == 0) {
            const size_t p = (((size_t)b * a.F + f) * a.N + n) * a.K + k;
            Pn[p] = s0;
            Pd[p] = s1;
        }
    }
}

// T[f,k] *= sqrt(sum_n Z[n,k] Pn / max(sum_n Z[n,k] Pd, eps))      mnmf.py:394-403
__global__ void smnmf_basis_finish_kernel(SmArgs a, const double* Pn, const double* Pd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)a.B * a.F * a.K) return;
    const int k = (int)(i % a.K);
    const long long bf = i / a.K;
    const int b = (int)(bf / a.F);
    double s0 = 0.0, s1 = 0.0;
    for (int n = 0; n < a.N; ++n) {
        const double z = a.Z[((size_t)b * a.N + n) * a.K + k];
        s0 = fma(z, Pn[((size_t)bf * a.N + n) * a.K + k], s0);
        s1 = fma(z, Pd[((size_t)bf * a.N + n) * a.K + k], s1);
    }
    if (s1 < a.eps) s1 = a.eps;
    a.T[i] *= sqrt(s0 / s1);
}

// Z[n,k] *= sqrt(sum_f T[f,k] Pn[f,n,k] / max(sum_f T[f,k] Pd[f,n,k], eps)); columns renormalised      mnmf.py:440-453
__global__ void __launch_bounds__(256) smnmf_latent_kernel(SmArgs a, const double* Pn, const double* Pd) {
    __shared__ double znew[NMAX * 64];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* Zg = a.Z + (size_t)b * a.N * a.K;
    for (int idx = warp; idx < a.N * a.K; idx += 8) {
        const int n = idx / a.K, k = idx % a.K;
        double s0 = 0.0, s1 = 0.0;
        for (int f = lane; f < a.F; f += 32) {
            const double tf = a.T[((size_t)b * a.F + f) * a.K + k];
            const size_t p = (((size_t)b * a.F + f) * a.N + n) * a.K + k;
            s0 = fma(tf, Pn[p], s0);
            s1 = fma(tf, Pd[p], s1);
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (s1 < a.eps) s1 = a.eps;
        if (lane == 0) znew[idx] = Zg[idx] * sqrt(s0 / s1);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < a.K; k += blockDim.x) {
        double s = 0.0;
        for (int n = 0; n < a.N; ++n) s += znew[n * a.K + k];
        if (s < a.eps) s = a.eps;
        for (int n = 0; n < a.N; ++n) Zg[n * a.K + k] = znew[n * a.K + k] / s;
    }
}

// partial[b][chunk][k][t][2] = sum_{f in chunk} T[f,k] sum_n Z[n,k] stat[n,f,t]      mnmf.py:418-424
__global__ void __launch_bounds__(TB) smnmf_act_partial_kernel(SmArgs a, const double* num, const double* den, double* part,
                                                                int n_chunks, int bins_per_chunk, int n_kgroups) {
    extern __shared__ double sm[];   // zt[bins_per_chunk][N][KGROUP]
    const int t = blockIdx.x * TB + threadIdx.x, chunk = blockIdx.y;
    const int b = blockIdx.z / n_kgroups, k0 = (blockIdx.z % n_kgroups) * KGROUP;
    const int f0 = chunk * bins_per_chunk;
    const int f1 = min(a.F, f0 + bins_per_chunk);
    for (int i = threadIdx.x; i < (f1 - f0) * a.N * KGROUP; i += TB) {
        const int kk = i % KGROUP, n = (i / KGROUP) % a.N, f = f0 + i / (KGROUP * a.N);
        const int k = k0 + kk;
        sm[i] = k < a.K ? a.Z[((size_t)b * a.N + n) * a.K + k] * a.T[((size_t)b * a.F + f) * a.K + k] : 0.0;
    }
    __syncthreads();
    if (t >= a.Tn) return;
    double s0[KGROUP], s1[KGROUP];
#pragma unroll
    for (int kk = 0; kk < KGROUP; ++kk) s0[kk] = s1[kk] = 0.0;
    for (int f = f0; f < f1; ++f)
        for (int n = 0; n < a.N; ++n) {
            const size_t o = (((size_t)b * a.N + n) * a.F + f) * a.Tn + t;
            const double x0 = num[o], x1 = den[o];
            const double* w = sm + ((size_t)(f - f0) * a.N + n) * KGROUP;
#pragma unroll
            for (int kk = 0; kk < KGROUP; ++kk) {
                s0[kk] = fma(w[kk], x0, s0[kk]);
                s1[kk] = fma(w[kk], x1, s1[kk]);
            }
        }
#pragma unroll
    for (int kk = 0; kk < KGROUP; ++kk) {
        const int k = k0 + kk;
        if (k < a.K) {
            const size_t p = ((((size_t)b * n_chunks + chunk) * a.K + k) * a.Tn + t) * 2;
            part[p] = s0[kk];
            part[p + 1] = s1[kk];
        }
    }
}

// V[k,t] *= sqrt(num / max(den, eps))      mnmf.py:424-427
__global__ void smnmf_act_finish_kernel(SmArgs a, const double* part, int n_chunks) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_b = (long long)a.K * a.Tn;
    if (i >= a.B * per_b) return;
    const int b = (int)(i / per_b);
    const long long kt = i % per_b;
    double s0 = 0.0, s1 = 0.0;
    for (int c = 0; c < n_chunks; ++c) {
        const size_t p = (((size_t)b * n_chunks + c) * per_b + kt) * 2;
        s0 += part[p];
        s1 += part[p + 1];
    }
    if (s1 < a.eps) s1 = a.eps;
    a.V[i] *= sqrt(s0 / s1);
}

// ------------------------------------------------------------------------------------------- spatial update
// acc[b][f][n][0] = sum_t lambda[n,f,t] Xh^-1[f,t], acc[b][f][n][1] = sum_t lambda[n,f,t] q q^H  (packed Hermitian).
// mnmf.py:468-475.  Phase 1: a thread per frame writes the 2 C^2 components and the N weights to shared memory; phase 2:
// a thread per output (n, component) accumulates over the frames of the block -- a small GEMM without atomics.
template <int C>
__global__ void __launch_bounds__(TB) smnmf_spatial_acc_kernel(SmArgs a, double* acc) {
    constexpr int CC = C * C, NV = 2 * CC, LD = NV + 1;
    extern __shared__ double sm[];
    double* Hs = sm;
    double* zt = Hs + a.N * CC;
    double* lamS = zt + a.N * a.K;      // [TB][NMAX]
    double* Ms = lamS + TB * NMAX;      // [TB][LD]
    const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    load_bin<C>(a, b, f, Hs, zt);
    const int n_out = a.N * NV;
    double out0 = 0.0, out1 = 0.0;
    const int o0 = tid, o1 = tid + TB;
    for (int t0 = 0; t0 < a.Tn; t0 += TB) {
        const int t = t0 + tid;
        const int cnt = min(TB, a.Tn - t0);
        if (t < a.Tn) {
            double lam[NMAX];
            source_power(a, b, t, zt, lam);
            Herm<C> S, Inv;
            model_covariance<C>(Hs, lam, a.N, S);
#pragma unroll
            for (int i = 0; i < C; ++i) S.d[i] += a.eps;
            herm_inverse<C>(S, Inv);
            double2 x[C], q[C];
            load_frame<C>(a.X + ((size_t)b * a.F + f) * C * a.Tp, a.Tp, t, x);
            herm_matvec<C>(Inv, x, q);
            double* m = Ms + tid * LD;
#pragma unroll
            for (int i = 0; i < C; ++i) {
                m[i] = Inv.d[i];
                m[CC + i] = q[i].x * q[i].x + q[i].y * q[i].y;
            }
#pragma unroll
            for (int i = 1; i < C; ++i)
#pragma unroll
                for (int j = 0; j < i; ++j) {
                    const int e = C + 2 * tri(i, j);
                    m[e] = Inv.o[tri(i, j)].x;
                    m[e + 1] = Inv.o[tri(i, j)].y;
                    // q_i conj(q_j)
                    m[CC + e] = q[i].x * q[j].x + q[i].y * q[j].y;
                    m[CC + e + 1] = q[i].y * q[j].x - q[i].x * q[j].y;
                }
#pragma unroll
            for (int n = 0; n < NMAX; ++n) lamS[tid * NMAX + n] = lam[n];
        }
        __syncthreads();
        if (o0 < n_out) {
            const int n = o0 / NV, comp = o0 % NV;
            for (int tt = 0; tt < cnt; ++tt) out0 = fma(lamS[tt * NMAX + n], Ms[tt * LD + comp], out0);
        }
        if (o1 < n_out) {
            const int n = o1 / NV, comp = o1 % NV;
            for (int tt = 0; tt < cnt; ++tt) out1 = fma(lamS[tt * NMAX + n], Ms[tt * LD + comp], out1);
        }
        __syncthreads();
    }
    double* og = acc + ((size_t)b * a.F + f) * n_out;
    if (o0 < n_out) og[o0] = out0;
    if (o1 < n_out) og[o1] = out1;
}

// H_n <- Riccati(A_n, H_n B_n H_n) + eps I, optionally divided by its trace      mnmf.py:474-483
template <int C>
__global__ void smnmf_riccati_kernel(SmArgs a, const double* acc, int normalize) {
    constexpr int CC = C * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)a.B * a.F * a.N) return;
    Mat<C> A, Bm, H, R, G;
    herm_unpack<C>(acc + (size_t)i * 2 * CC, A);
    herm_unpack<C>(acc + (size_t)i * 2 * CC + CC, Bm);
    herm_unpack<C>(a.H + (size_t)i * CC, H);
    mat_mul(H, Bm, R);
    mat_mul(R, H, G);
    riccati_hermitian<C>(A, G, R);
    double tr = 0.0;
    for (int c = 0; c < C; ++c) {
        R.a[c][c].x += a.eps;
        tr += R.a[c][c].x;
    }
    const double s = normalize ? 1.0 / tr : 1.0;
    double* p = a.H + (size_t)i * CC;
    for (int c = 0; c < C; ++c) p[c] = R.a[c][c].x * s;
    int e = C;
    for (int r = 1; r < C; ++r)
        for (int c = 0; c < r; ++c) {
            p[e] = R.a[r][c].x * s;
            p[e + 1] = R.a[r][c].y * s;
            e += 2;
        }
}

// ------------------------------------------------------------------------------------------- loss / separation
// terms[b][f] = sum_t D_LD(X'', Xh'') with the regularisations of to_PSD (utils_linalg.py:9-31) and mnmf.py:582-583:
// X'' = x x^H + (eps |x|^2 + eps) I,  Xh'' = Xh + (eps tr Xh + eps) I;  divergence.py:96-104.  X'' has the eigenvalues
// c1 (C-1 times) and |x|^2 + c1 with c1 = eps |x|^2 + eps.
template <int C>
__global__ void __launch_bounds__(TB) smnmf_loss_kernel(SmArgs a, double* terms) {
    extern __shared__ double sm[];
    __shared__ double red[TB / 32];
    double* Hs = sm;
    double* zt = Hs + a.N * C * C;
    const int f = blockIdx.x, b = blockIdx.y;
    load_bin<C>(a, b, f, Hs, zt);
    double sum = 0.0;
    for (int t = threadIdx.x; t < a.Tn; t += TB) {
        double lam[NMAX];
        source_power(a, b, t, zt, lam);
        Herm<C> S, Inv;
        model_covariance<C>(Hs, lam, a.N, S);
        double trS = 0.0;
#pragma unroll
        for (int i = 0; i < C; ++i) trS += S.d[i];
        const double add = a.eps * trS + a.eps;
#pragma unroll
        for (int i = 0; i < C; ++i) S.d[i] += add;
        const double logdet_h = herm_inverse<C>(S, Inv);
        double2 x[C], q[C];
        load_frame<C>(a.X + ((size_t)b * a.F + f) * C * a.Tp, a.Tp, t, x);
        herm_matvec<C>(Inv, x, q);
        double x2 = 0.0, xq = 0.0, tri_ = 0.0;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            x2 += x[i].x * x[i].x + x[i].y * x[i].y;
            xq += x[i].x * q[i].x + x[i].y * q[i].y;
            tri_ += Inv.d[i];
        }
        const double c1 = a.eps * x2 + a.eps;
        const double logdet_x = (C - 1) * log(fmax(c1, a.eps)) + log(fmax(x2 + c1, a.eps));
        sum += xq + c1 * tri_ - (logdet_x - logdet_h) - C;
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < TB / 32; ++w) s += red[w];
        terms[(size_t)b * a.F + f] = s;
    }
}

// y[b][n][f][t] = lambda[n,f,t] (H_n Xh^-1 x)[ref]      mnmf.py:609-634
template <int C>
__global__ void __launch_bounds__(TB) smnmf_separate_kernel(SmArgs a, int ref, cf* out) {
    extern __shared__ double sm[];
    double* Hs = sm;
    double* zt = Hs + a.N * C * C;
    const int f = blockIdx.y, b = blockIdx.z, t = blockIdx.x * TB + threadIdx.x;
    load_bin<C>(a, b, f, Hs, zt);
    if (t >= a.Tn) return;
    double lam[NMAX];
    source_power(a, b, t, zt, lam);
    Herm<C> S, Inv;
    model_covariance<C>(Hs, lam, a.N, S);
#pragma unroll
    for (int i = 0; i < C; ++i) S.d[i] += a.eps;
    herm_inverse<C>(S, Inv);
    double2 x[C], q[C];
    load_frame<C>(a.X + ((size_t)b * a.F + f) * C * a.Tp, a.Tp, t, x);
    herm_matvec<C>(Inv, x, q);
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
        if (n < a.N) {
            const double* p = Hs + n * C * C;
            double2 s = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < C; ++j) {
                double hr, hi;
                if (j == ref) {
                    hr = p[j];
                    hi = 0.0;
                } else if (j < ref) {
                    hr = p[C + 2 * tri(ref, j)];
                    hi = p[C + 2 * tri(ref, j) + 1];
                } else {
                    hr = p[C + 2 * tri(j, ref)];
                    hi = -p[C + 2 * tri(j, ref) + 1];
                }
                s.x += hr * q[j].x - hi * q[j].y;
                s.y += hr * q[j].y + hi * q[j].x;
            }
            out[(((size_t)b * a.N + n) * a.F + f) * a.Tn + t] = cf_make((float)(lam[n] * s.x), (float)(lam[n] * s.y));
        }
    }
}

SmArgs make_args(const bss_handle* h) {
    SmArgs a;
    a.X = h->X;
    a.H = h->sH;
    a.Z = h->sZ;
    a.T = h->sT;
    a.V = h->sV;
    a.B = h->B;
    a.N = h->N;
    a.F = h->F;
    a.Tn = h->T;
    a.Tp = h->Tp;
    a.K = h->K;
    a.eps = h->cfg.eps;
    return a;
}

size_t bin_smem(const bss_handle* h) { return ((size_t)h->N * h->C * h->C + (size_t)h->N * h->K) * sizeof(double); }

#define SMNMF_DISPATCH(C, ...)                                \
    switch (C) {                                              \
        case 2: { constexpr int CH = 2; __VA_ARGS__; } break;        \
        case 3: { constexpr int CH = 3; __VA_ARGS__; } break;        \
        case 4: { constexpr int CH = 4; __VA_ARGS__; } break;        \
        default: return bss_fail(h, BSS_EINVAL, "IS-MNMF supports 2 to 4 channels"); \
    }

}  // namespace

// bins per chunk of the activation reduction (bounded so the chunk's Z T products fit in shared memory) and chunk count
void smnmf_act_plan(const bss_handle* h, int* bins, int* chunks) {
    int b = (int)cdiv(h->F, 32);
    if (b > 64) b = 64;
    if (b < 1) b = 1;
    *bins = b;
    *chunks = (int)cdiv(h->F, b);
}

int launch_smnmf_stats(bss_handle* h) {
    const SmArgs a = make_args(h);
    const dim3 grid((unsigned)cdiv(h->T, TB), h->F, h->B);
    double* num = h->sStat;
    double* den = h->sStat + (size_t)h->B * h->N * h->F * h->T;
    SMNMF_DISPATCH(h->C, (smnmf_stats_kernel<CH><<<grid, TB, bin_smem(h), h->stream>>>(a, num, den)))
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

// which: 0 basis, 1 activation, 2 latent -- the contraction of the trace statistics that finishes one MM update
int launch_smnmf_factor(bss_handle* h, int which) {
    const SmArgs a = make_args(h);
    const double* num = h->sStat;
    const double* den = h->sStat + (size_t)h->B * h->N * h->F * h->T;
    if (which == 1) {
        int bins = 0, n_chunks_used = 0;
        smnmf_act_plan(h, &bins, &n_chunks_used);
        const int n_kg = (int)cdiv(h->K, KGROUP);
        const dim3 grid((unsigned)cdiv(h->T, TB), n_chunks_used, h->B * n_kg);
        const size_t smem = (size_t)bins * h->N * KGROUP * sizeof(double);
        smnmf_act_partial_kernel<<<grid, TB, smem, h->stream>>>(a, num, den, h->sPart, n_chunks_used, bins, n_kg);
        h->launches++;
        BSS_CUDA(h, cudaGetLastError());
        const long long n = (long long)h->B * h->K * h->T;
        smnmf_act_finish_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(a, h->sPart, n_chunks_used);
        h->launches++;
        BSS_CUDA(h, cudaGetLastError());
        return BSS_OK;
    }
    double* Pn = h->sPart;
    double* Pd = h->sPart + (size_t)h->B * h->F * h->N * h->K;
    smnmf_nk_kernel<<<dim3(h->F, h->B), TB, 0, h->stream>>>(a, num, den, Pn, Pd);
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    if (which == 0) {
        const long long n = (long long)h->B * h->F * h->K;
        smnmf_basis_finish_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(a, Pn, Pd);
    } else {
        smnmf_latent_kernel<<<h->B, 256, 0, h->stream>>>(a, Pn, Pd);
    }
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_smnmf_spatial(bss_handle* h, int normalize) {
    const SmArgs a = make_args(h);
    const int C = h->C;
    const size_t smem = bin_smem(h) + ((size_t)TB * NMAX + (size_t)TB * (2 * C * C + 1)) * sizeof(double);
    SMNMF_DISPATCH(C, {
        static bool attr_set = false;
        if (!attr_set) {
            BSS_CUDA(h, cudaFuncSetAttribute(smnmf_spatial_acc_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            attr_set = true;
        }
        smnmf_spatial_acc_kernel<CH><<<dim3(h->F, h->B), TB, smem, h->stream>>>(a, h->sAcc);
    })
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    const long long n = (long long)h->B * h->F * h->N;
    SMNMF_DISPATCH(C, (smnmf_riccati_kernel<CH><<<(unsigned)cdiv(n, 64), 64, 0, h->stream>>>(a, h->sAcc, normalize)))
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_smnmf_loss_terms(bss_handle* h, double* terms) {
    const SmArgs a = make_args(h);
    SMNMF_DISPATCH(h->C, (smnmf_loss_kernel<CH><<<dim3(h->F, h->B), TB, bin_smem(h), h->stream>>>(a, terms)))
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_smnmf_separate(bss_handle* h, cf* out) {
    const SmArgs a = make_args(h);
    const dim3 grid((unsigned)cdiv(h->T, TB), h->F, h->B);
    SMNMF_DISPATCH(h->C, (smnmf_separate_kernel<CH><<<grid, TB, bin_smem(h), h->stream>>>(a, h->cfg.reference_id, out)))
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}
