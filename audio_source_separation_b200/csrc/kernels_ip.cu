// Per-bin demixing-matrix updates in fp64 registers: one thread owns one (mixture, bin).
//   - iterative projection sweep         src/bss/ilrma.py:512-530, src/bss/iva.py:500-518, src/bss/mnmf.py:872-886
//   - pairwise (IP2) update              src/bss/ilrma.py:599-626, src/bss/iva.py:566-592
//   - projection-back scale              src/algorithm/projection_back.py:12-21
//   - log|det W|                         src/bss/ilrma.py:675
//   - least-squares filter from (Y, X)   src/bss/ilrma.py:167-173
#include "handle.h"
#include "smallmat.cuh"

namespace {

template <int C>
__device__ __forceinline__ void load_w(const double2* W, Mat<C>& M) {
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const double2 v = W[i * C + j];
            M.a[i][j] = cd_make(v.x, v.y);
        }
}
template <int C>
__device__ __forceinline__ void store_w(double2* W, const Mat<C>& M, cf* Wf = nullptr) {
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            W[i * C + j] = make_double2(M.a[i][j].x, M.a[i][j].y);
            if (Wf) Wf[i * C + j] = cf_make((float)M.a[i][j].x, (float)M.a[i][j].y);
        }
}

// p_n = w_n^H Cx w_n with w_n^H = row n of W  ->  sum_ij W[n][i] Cx[i][j] conj(W[n][j])
template <int C>
__device__ __forceinline__ double row_power(const Mat<C>& W, const Mat<C>& Cx, int n) {
    cd q = cd_make(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < C; ++i) {
        cd s = cd_make(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < C; ++j) cd_fma(s, Cx.a[i][j], cd_conj(W.a[n][j]));
        cd_fma(q, W.a[n][i], s);
    }
    return q.x;
}

// (an L1 prefetch of the later U_n at the top of the sweep did not move the 173 us of the 64-mixture case: profiles/r5f_*)
template <int C>
__global__ void __launch_bounds__(64) ip_sweep_kernel(const IpArgs a) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.B * a.F) return;
    const int b = (int)(idx / a.F), f = (int)(idx - (long long)b * a.F);
    Mat<C> W;
    load_w<C>(a.W + (size_t)idx * C * C, W);
    bool singular = false;
#pragma unroll 1
    for (int n = 0; n < C; ++n) {
        Mat<C> U;
        herm_unpack<C>(a.U + (((size_t)b * C + n) * a.F + f) * C * C, U);
        const int ok = ip_row<C>(W, U, n, a.threshold, a.use_gate != 0, a.floor_den != 0, a.eps, &singular);
        if (a.gate) a.gate[((size_t)b * C + n) * a.F + f] = ok;
    }
    if (singular) atomicAdd(a.flags, 1);
    store_w<C>(a.W + (size_t)idx * C * C, W, a.Wf ? a.Wf + (size_t)idx * C * C : nullptr);
    if (a.pw) {
        Mat<C> Cx;
        herm_unpack<C>(a.Cx + (size_t)idx * C * C, Cx);
#pragma unroll 1
        for (int n = 0; n < C; ++n) a.pw[((size_t)b * C + n) * a.F + f] = row_power<C>(W, Cx, n);
    }
}

// ------------------------------------------------------------------------------------------- cooperative sweep
// The same sweep with G = 2 / 4 / 8 lanes per bin: lane j of a group owns column j of U_n, of A = W U_n and of
// the inverse being built, all in registers; W sits in shared memory; pivot rows and eliminators travel by
// width-G shuffles.  The operations and their order are those of mat_inverse / ip_row (smallmat.cuh), so the
// results agree with the one-thread-per-bin form to rounding -- but a bin costs 1/G of the serial
// latency and needs no local memory (the serial 8 x 8 form spilled 17 KB per thread).
__device__ __forceinline__ cd shfl_cd(cd v, int src, int width) {
    return cd_make(__shfl_sync(BSS_FULL, v.x, src, width), __shfl_sync(BSS_FULL, v.y, src, width));
}
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(BSS_FULL, v, o, G);
    return v;
}

// column j of a packed Hermitian matrix (C diagonals, then strict lower triangle row-major as (re, im))
template <int C>
__device__ __forceinline__ void herm_col(const double* p, int j, cd (&col)[C]) {
#pragma unroll
    for (int i = 0; i < C; ++i) {
        if (i == j) {
            col[i] = cd_make(__ldg(p + i), 0.0);
        } else {
            const int hi = i > j ? i : j, lo = i > j ? j : i;
            const int e = C + 2 * (hi * (hi - 1) / 2 + lo);
            const double re = __ldg(p + e), im = __ldg(p + e + 1);
            col[i] = cd_make(re, i > j ? im : -im);
        }
    }
}

template <int C>
__device__ __noinline__ double cond2_of(const cd* A) {
    Mat<C> M;
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) M.a[i][j] = A[i * C + j];
    return mat_cond2(M);
}

template <int C>
__global__ void __launch_bounds__(128) ip_sweep_group_kernel(const IpArgs a) {
    constexpr int G = C <= 2 ? 2 : (C <= 4 ? 4 : 8);
    constexpr int BPW = 32 / G;                      // bins per warp
    extern __shared__ __align__(16) unsigned char ip_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, j = lane % G;
    const bool col_live = j < C;
    const long long n_bins = (long long)a.B * a.F;
    long long idx = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * BPW + grp;
    const bool bin_live = idx < n_bins;
    if (!bin_live) idx = n_bins - 1;                 // keep the lanes in the shuffles; nothing is written
    const int b = (int)(idx / a.F), f = (int)(idx - (long long)b * a.F);
    cd* Ws = reinterpret_cast<cd*>(ip_smem) + (size_t)(warp * BPW + grp) * 2 * C * C;   // W, then a copy of A
    cd* As = Ws + C * C;
    for (int e = j; e < C * C; e += G) {
        const double2 v = a.W[(size_t)idx * C * C + e];
        Ws[e] = cd_make(v.x, v.y);
    }
    __syncwarp();
    bool singular = false;
#pragma unroll 1
    for (int n = 0; n < C; ++n) {
        cd U[C], A[C], I[C];
        if (col_live) {
            herm_col<C>(a.U + (((size_t)b * C + n) * a.F + f) * C * C, j, U);
        } else {
#pragma unroll
            for (int i = 0; i < C; ++i) U[i] = cd_make(0.0, 0.0);
        }
        // A[:, j] = W U[:, j]
        double fa = 0.0;
#pragma unroll
        for (int i = 0; i < C; ++i) {
            cd s = cd_make(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < C; ++k) cd_fma(s, Ws[i * C + k], U[k]);
            A[i] = s;
            I[i] = cd_make(i == j ? 1.0 : 0.0, 0.0);
            fa += cd_abs2(s);
            if (col_live) As[i * C + j] = s;
        }
        fa = group_sum<G>(fa);
        // Gauss-Jordan with partial pivoting, column k owned by lane k
        bool inv_ok = true;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            int p = k;
            double best = fabs(A[k].x) + fabs(A[k].y);
#pragma unroll
            for (int i = k + 1; i < C; ++i) {
                const double v = fabs(A[i].x) + fabs(A[i].y);
                if (v > best) {
                    best = v;
                    p = i;
                }
            }
            p = __shfl_sync(BSS_FULL, p, k, G);
            best = __shfl_sync(BSS_FULL, best, k, G);
            if (best == 0.0) inv_ok = false;
#pragma unroll
            for (int i = k + 1; i < C; ++i) {
                if (i == p) {
                    cd t = A[k];
                    A[k] = A[i];
                    A[i] = t;
                    t = I[k];
                    I[k] = I[i];
                    I[i] = t;
                }
            }
            const cd piv = cd_div(cd_make(1.0, 0.0), shfl_cd(A[k], k, G));
            A[k] = A[k] * piv;
            I[k] = I[k] * piv;
            cd fct[C];
#pragma unroll
            for (int i = 0; i < C; ++i) fct[i] = shfl_cd(A[i], k, G);
#pragma unroll
            for (int i = 0; i < C; ++i) {
                if (i == k) continue;
                A[i] = A[i] - fct[i] * A[k];
                I[i] = I[i] - fct[i] * I[k];
            }
        }
        double fi = 0.0;
#pragma unroll
        for (int i = 0; i < C; ++i) fi += col_live ? cd_abs2(I[i]) : 0.0;
        fi = group_sum<G>(fi);
        // every shuffle below is executed by all 32 lanes: the groups of a warp may disagree on the branches
        int ok = 0;
        bool need_svd = false;
        if (!inv_ok) {
            singular = true;
        } else if (!a.use_gate) {
            ok = 1;
        } else {
            // kappa_F / C <= cond_2 <= kappa_F: only the band around the threshold needs the SVD (cond_below)
            const double kf = sqrt(fa * fi);
            if (!(kf == kf))
                ok = 0;
            else if (kf < a.threshold)
                ok = 1;
            else if (kf >= a.threshold * C)
                ok = 0;
            else
                need_svd = true;
        }
        if (__any_sync(BSS_FULL, need_svd)) {
            __syncwarp();
            double c2 = 0.0;
            if (need_svd && j == 0) c2 = cond2_of<C>(As);
            c2 = __shfl_sync(BSS_FULL, c2, 0, G);
            if (need_svd) ok = c2 < a.threshold ? 1 : 0;
        }
        {
            // w = column n of the inverse, to every lane of the group
            cd w[C];
#pragma unroll
            for (int i = 0; i < C; ++i) w[i] = shfl_cd(I[i], n, G);
            cd wj = w[0];
#pragma unroll
            for (int i = 1; i < C; ++i)
                if (i == j) wj = w[i];
            // q = w^H U w, lane j contributes (sum_i conj(w_i) U[i][j]) w_j
            cd s = cd_make(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < C; ++i) cd_fma(s, cd_conj(w[i]), U[i]);
            const cd term = col_live ? s * wj : cd_make(0.0, 0.0);
            const cd q = cd_make(group_sum<G>(term.x), group_sum<G>(term.y));
            cd den = cd_sqrt(q);
            if (a.floor_den && cd_less_real(den, a.eps)) den = cd_make(a.eps, 0.0);
            __syncwarp();
            if (inv_ok && ok && col_live) Ws[n * C + j] = cd_div(cd_conj(wj), den);
        }
        __syncwarp();
        if (a.gate && bin_live && j == 0) a.gate[((size_t)b * C + n) * a.F + f] = ok;
    }
    if (singular && bin_live && j == 0) atomicAdd(a.flags, 1);
    if (bin_live) {
        for (int e = j; e < C * C; e += G) {
            const cd v = Ws[e];
            a.W[(size_t)idx * C * C + e] = make_double2(v.x, v.y);
            if (a.Wf) a.Wf[(size_t)idx * C * C + e] = cf_make((float)v.x, (float)v.y);
        }
    }
    if (a.pw) {
        // p_n = w_n^H Cx w_n = sum_j (sum_i W[n][i] Cx[i][j]) conj(W[n][j])
        cd Cx[C];
        if (col_live) {
            herm_col<C>(a.Cx + (size_t)idx * C * C, j, Cx);
        } else {
#pragma unroll
            for (int i = 0; i < C; ++i) Cx[i] = cd_make(0.0, 0.0);
        }
#pragma unroll 1
        for (int n = 0; n < C; ++n) {
            cd s = cd_make(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < C; ++i) cd_fma(s, Ws[n * C + i], Cx[i]);
            const cd t = col_live ? s * cd_conj(Ws[n * C + j]) : cd_make(0.0, 0.0);
            const double pr = group_sum<G>(t.x);
            if (bin_live && j == 0) a.pw[((size_t)b * C + n) * a.F + f] = pr;
        }
    }
}

// 2x2 helper for IP2
struct M2 {
    cd a, b, c, d;   // [[a, b], [c, d]]
};
__device__ __forceinline__ M2 m2_mul(const M2& x, const M2& y) {
    M2 r;
    r.a = x.a * y.a + x.b * y.c;
    r.b = x.a * y.b + x.b * y.d;
    r.c = x.c * y.a + x.d * y.c;
    r.d = x.c * y.b + x.d * y.d;
    return r;
}
__device__ __forceinline__ M2 m2_inv(const M2& x) {
    const cd det = x.a * x.d - x.b * x.c;
    const cd one = cd_make(1.0, 0.0);
    const cd id = cd_div(one, det);
    M2 r;
    r.a = x.d * id;
    r.b = cd_make(-x.b.x, -x.b.y) * id;
    r.c = cd_make(-x.c.x, -x.c.y) * id;
    r.d = x.a * id;
    return r;
}
// unit-norm eigenvector of the 2x2 matrix m for eigenvalue lam, largest component real (LAPACK zgeev convention)
__device__ __forceinline__ void m2_eigvec(const M2& m, cd lam, cd& v0, cd& v1) {
    // (m - lam I) v = 0: two candidate null vectors, take the one built from the larger row
    const cd r0a = m.a - lam, r0b = m.b;
    const cd r1a = m.c, r1b = m.d - lam;
    const double n0 = cd_abs2(r0a) + cd_abs2(r0b), n1 = cd_abs2(r1a) + cd_abs2(r1b);
    if (n0 >= n1) {
        v0 = r0b;
        v1 = cd_make(-r0a.x, -r0a.y);
    } else {
        v0 = r1b;
        v1 = cd_make(-r1a.x, -r1a.y);
    }
    double nn = sqrt(cd_abs2(v0) + cd_abs2(v1));
    if (nn == 0.0) {   // m = lam I
        v0 = cd_make(1.0, 0.0);
        v1 = cd_make(0.0, 0.0);
        nn = 1.0;
    }
    const cd big = cd_abs2(v0) >= cd_abs2(v1) ? v0 : v1;
    const double bn = cd_abs(big);
    const cd ph = cd_make(big.x / bn, -big.y / bn);   // conj(phase of the largest component)
    v0 = (1.0 / nn) * (v0 * ph);
    v1 = (1.0 / nn) * (v1 * ph);
}

// Pairwise update of rows m and n.  order_out[0/1] records which of the two closed-form eigenvalues
// (index 0 = "+" root, 1 = "-" root) went to row m / row n, i.e. argsort(lam)[::-1].
template <int C>
__global__ void __launch_bounds__(64) ip2_kernel(const IpArgs a, int32_t* order_out, double2* eig_out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)a.B * a.F) return;
    const int b = (int)(idx / a.F), f = (int)(idx - (long long)b * a.F);
    const int pm = a.pair_m, pn = a.pair_n;
    Mat<C> W, Um, Un, A, Gm, Gn;
    load_w<C>(a.W + (size_t)idx * C * C, W);
    herm_unpack<C>(a.U + (((size_t)b * C + pm) * a.F + f) * C * C, Um);
    herm_unpack<C>(a.U + (((size_t)b * C + pn) * a.F + f) * C * C, Un);
    bool singular = false;
    mat_mul(W, Um, A);
    const bool inv_m = mat_inverse(A, Gm);
    const bool ok_m = cond_below(A, Gm, inv_m, a.threshold);
    mat_mul(W, Un, A);
    const bool inv_n = mat_inverse(A, Gn);
    const bool ok_n = cond_below(A, Gn, inv_n, a.threshold);
    if (!inv_m || !inv_n) singular = true;

    // P = inverse @ [e_m e_n]  (C x 2): columns pm and pn
    cd Pm[C][2], Pn[C][2];
#pragma unroll
    for (int i = 0; i < C; ++i) {
        Pm[i][0] = Pm[i][1] = Pn[i][0] = Pn[i][1] = cd_make(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < C; ++j) {
            if (j == pm) {
                Pm[i][0] = Gm.a[i][j];
                Pn[i][0] = Gn.a[i][j];
            }
            if (j == pn) {
                Pm[i][1] = Gm.a[i][j];
                Pn[i][1] = Gn.a[i][j];
            }
        }
    }
    // V = P^H U P (2x2)
    M2 Vm, Vn;
    {
        cd t[C][2];
#pragma unroll
        for (int i = 0; i < C; ++i)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                cd s = cd_make(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < C; ++j) cd_fma(s, Um.a[i][j], Pm[j][c]);
                t[i][c] = s;
            }
        cd v[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                cd s = cd_make(0.0, 0.0);
#pragma unroll
                for (int i = 0; i < C; ++i) cd_fma(s, cd_conj(Pm[i][r]), t[i][c]);
                v[r][c] = s;
            }
        Vm.a = v[0][0];
        Vm.b = v[0][1];
        Vm.c = v[1][0];
        Vm.d = v[1][1];
#pragma unroll
        for (int i = 0; i < C; ++i)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                cd s = cd_make(0.0, 0.0);
#pragma unroll
                for (int j = 0; j < C; ++j) cd_fma(s, Un.a[i][j], Pn[j][c]);
                t[i][c] = s;
            }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                cd s = cd_make(0.0, 0.0);
#pragma unroll
                for (int i = 0; i < C; ++i) cd_fma(s, cd_conj(Pn[i][r]), t[i][c]);
                v[r][c] = s;
            }
        Vn.a = v[0][0];
        Vn.b = v[0][1];
        Vn.c = v[1][0];
        Vn.d = v[1][1];
    }
    // eig(Vn^-1 Vm), closed form
    const M2 Mx = m2_mul(m2_inv(Vn), Vm);
    const cd tr = Mx.a + Mx.d;
    const cd det = Mx.a * Mx.d - Mx.b * Mx.c;
    const cd disc = cd_sqrt(tr * tr - 4.0 * det);
    cd lam0 = 0.5 * (tr + disc), lam1 = 0.5 * (tr - disc);
    // descending complex-lexicographic order (np.argsort(lam)[::-1]; a tie gives [1, 0], see cd_lex_greater)
    const bool first_is_max = cd_lex_greater(lam0, lam1);
    if (eig_out) {
        eig_out[idx * 2 + 0] = make_double2(lam0.x, lam0.y);
        eig_out[idx * 2 + 1] = make_double2(lam1.x, lam1.y);
    }
    const cd lmax = first_is_max ? lam0 : lam1, lmin = first_is_max ? lam1 : lam0;
    if (order_out) {
        order_out[idx * 2 + 0] = first_is_max ? 0 : 1;
        order_out[idx * 2 + 1] = first_is_max ? 1 : 0;
    }
    cd vm0, vm1, vn0, vn1;
    m2_eigvec(Mx, lmax, vm0, vm1);
    m2_eigvec(Mx, lmin, vn0, vn1);
    // normalise by sqrt(v^H V v)
    {
        const cd q = cd_conj(vm0) * (Vm.a * vm0 + Vm.b * vm1) + cd_conj(vm1) * (Vm.c * vm0 + Vm.d * vm1);
        const cd s = cd_sqrt(q);
        vm0 = cd_div(vm0, s);
        vm1 = cd_div(vm1, s);
    }
    {
        const cd q = cd_conj(vn0) * (Vn.a * vn0 + Vn.b * vn1) + cd_conj(vn1) * (Vn.c * vn0 + Vn.d * vn1);
        const cd s = cd_sqrt(q);
        vn0 = cd_div(vn0, s);
        vn1 = cd_div(vn1, s);
    }
#pragma unroll
    for (int r = 0; r < C; ++r) {
        if (r == pm && ok_m) {
#pragma unroll
            for (int j = 0; j < C; ++j) W.a[r][j] = cd_conj(Pm[j][0] * vm0 + Pm[j][1] * vm1);
        }
    }
#pragma unroll
    for (int r = 0; r < C; ++r) {
        if (r == pn && ok_n) {
#pragma unroll
            for (int j = 0; j < C; ++j) W.a[r][j] = cd_conj(Pn[j][0] * vn0 + Pn[j][1] * vn1);
        }
    }
    if (a.gate) {
        a.gate[((size_t)b * C + pm) * a.F + f] = ok_m ? 1 : 0;
        a.gate[((size_t)b * C + pn) * a.F + f] = ok_n ? 1 : 0;
    }
    if (singular) atomicAdd(a.flags, 1);
    store_w<C>(a.W + (size_t)idx * C * C, W, a.Wf ? a.Wf + (size_t)idx * C * C : nullptr);
    if (a.pw) {
        Mat<C> Cx;
        herm_unpack<C>(a.Cx + (size_t)idx * C * C, Cx);
#pragma unroll 1
        for (int n = 0; n < C; ++n) a.pw[((size_t)b * C + n) * a.F + f] = row_power<C>(W, Cx, n);
    }
}

// scale[n] = (e_ref^T Cx W^H (W Cx W^H)^-1)[n]   == projection_back(W X, X[ref]) since
// Y Y^H = T W Cx W^H and x_ref Y^H = T (Cx W^H)[ref,:]   (src/algorithm/projection_back.py:15-21)
template <int C>
__global__ void __launch_bounds__(64) pb_scale_kernel(const double2* Wg, const double* Cxg, double2* scale, int B,
                                                     int F, int ref, int32_t* flags) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * F) return;
    const int b = (int)(idx / F), f = (int)(idx - (long long)b * F);
    Mat<C> W, Cx, CW, G, Gi;
    load_w<C>(Wg + (size_t)idx * C * C, W);
    herm_unpack<C>(Cxg + (size_t)idx * C * C, Cx);
    // CW = Cx W^H
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            cd s = cd_make(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < C; ++k) cd_fma(s, Cx.a[i][k], cd_conj(W.a[j][k]));
            CW.a[i][j] = s;
        }
    mat_mul(W, CW, G);
    if (!mat_inverse(G, Gi)) atomicAdd(flags, 1);
#pragma unroll 1
    for (int n = 0; n < C; ++n) {
        cd s = cd_make(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < C; ++k) {
            cd row = CW.a[0][k];
#pragma unroll
            for (int r = 1; r < C; ++r)
                if (r == ref) row = CW.a[r][k];
            cd_fma(s, row, Gi.a[k][n]);
        }
        scale[((size_t)b * C + n) * F + f] = make_double2(s.x, s.y);
    }
}

template <int C>
__global__ void __launch_bounds__(64) logdet_kernel(const double2* Wg, double* out, long long n_bins, int transpose_sq) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_bins) return;
    Mat<C> W;
    load_w<C>(Wg + (size_t)idx * C * C, W);
    if (transpose_sq) {   // FastMNMF loss uses det(Q Q^T), plain transpose (src/bss/mnmf.py:911)
        Mat<C> Wt, P;
#pragma unroll
        for (int i = 0; i < C; ++i)
#pragma unroll
            for (int j = 0; j < C; ++j) Wt.a[i][j] = W.a[j][i];
        mat_mul(W, Wt, P);
        W = P;
    }
    out[idx] = log(cd_abs(mat_det<C>(W)));
}

// W = (Y X^H) (X X^H)^-1 with G = Y X^H / T given as full complex [C][C] and Cx = X X^H / T
template <int C>
__global__ void __launch_bounds__(64) lsq_filter_kernel(const double2* Gg, const double* Cxg, double2* Wg, long long n_bins,
                                                       int32_t* flags) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_bins) return;
    Mat<C> G, Cx, Ci, W;
    load_w<C>(Gg + (size_t)idx * C * C, G);
    herm_unpack<C>(Cxg + (size_t)idx * C * C, Cx);
    if (!mat_inverse(Cx, Ci)) atomicAdd(flags, 1);
    mat_mul(G, Ci, W);
    store_w<C>(Wg + (size_t)idx * C * C, W);
}

// M_f = A_f B_f^H (B_f B_f^H)^-1 for arbitrary complex128 arrays A (Ra,F,T), B (Rb,F,T): projection_back(Y, reference)
// (src/algorithm/projection_back.py:12-21, :25-32) and compute_demix_filter(Y, X) (src/bss/ilrma.py:167-173).
// One CTA per bin; a warp owns one entry of [A B^H ; B B^H] at a time and sums over the frames in a fixed order.
template <int RB>
__global__ void __launch_bounds__(128) lsq_map_kernel(const double2* A, const double2* Bm, double2* out, int Ra, int F, int T,
                                                      int32_t* flags) {
    __shared__ cd G[8 + RB][RB];   // rows [0, Ra): A B^H, rows [8, 8 + RB): B B^H
    const int f = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_entries = (Ra + RB) * RB;
    for (int e = warp; e < n_entries; e += 4) {
        const int r = e / RB, n = e - r * RB;
        const double2* row = r < Ra ? A + ((size_t)r * F + f) * T : Bm + ((size_t)(r - Ra) * F + f) * T;
        const double2* col = Bm + ((size_t)n * F + f) * T;
        double sx = 0.0, sy = 0.0;
        for (int t = lane; t < T; t += 32) {
            const double2 x = row[t], y = col[t];
            sx += x.x * y.x + x.y * y.y;   // x conj(y)
            sy += x.y * y.x - x.x * y.y;
        }
        sx = warp_sum(sx);
        sy = warp_sum(sy);
        if (lane == 0) G[r < Ra ? r : 8 + (r - Ra)][n] = cd_make(sx, sy);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    Mat<RB> H, Hi;
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
        for (int j = 0; j < RB; ++j) H.a[i][j] = G[8 + i][j];
    if (!mat_inverse(H, Hi)) atomicAdd(flags, 1);
    for (int r = 0; r < Ra; ++r)
#pragma unroll
        for (int n = 0; n < RB; ++n) {
            cd s = cd_make(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < RB; ++k) cd_fma(s, G[r][k], Hi.a[k][n]);
            out[((size_t)r * RB + n) * F + f] = make_double2(s.x, s.y);
        }
}

template <int C>
int launch_ip_t(bss_handle* h, const IpArgs& a, int32_t* order_out) {
    const long long n = (long long)a.B * a.F;
    const int threads = 64;
    const unsigned grid = (unsigned)cdiv(n, threads);
    // plenty of bins: one thread per bin has no shuffle traffic and wins (171 vs 212 us for 64 x 2049 bins, C = 4);
    // BSS_OPT_IP_KERNEL overrides the choice (the 8 x 8 thread-per-bin form spills, but it is still the same arithmetic)
    const bool per_thread = a.variant == 1 || (a.variant != 2 && C <= 4 && n >= 32768);
    if (a.pair_m >= 0) {
        ip2_kernel<C><<<grid, threads, 0, h->stream>>>(a, order_out, a.eigval);
        h->last_ip_kernel = 4;
    } else if (per_thread) {
        // (keeping W in shared memory to raise the occupancy -- 200 or 168 registers, five or six CTAs per SM -- measured the
        // same 124 - 126 us as this all-register form once the exact route had been moved out of line: profiles/round2_ab_ip_forms.md)
        ip_sweep_kernel<C><<<grid, threads, 0, h->stream>>>(a);
        h->last_ip_kernel = 1;
    } else {
        h->last_ip_kernel = 2;
        constexpr int G = C <= 2 ? 2 : (C <= 4 ? 4 : 8);
        constexpr int BPW = 32 / G;
        const int warps = 4;
        const size_t smem = (size_t)warps * BPW * 2 * C * C * sizeof(cd);
        ip_sweep_group_kernel<C><<<(unsigned)cdiv(n, (long long)warps * BPW), warps * 32, smem, h->stream>>>(a);
    }
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

}  // namespace

#define BSS_DISPATCH_C(Cval, CALL)                                                   \
    switch (Cval) {                                                                  \
        case 2: { constexpr int CC_ = 2; CALL; } break;                              \
        case 3: { constexpr int CC_ = 3; CALL; } break;                              \
        case 4: { constexpr int CC_ = 4; CALL; } break;                              \
        case 5: { constexpr int CC_ = 5; CALL; } break;                              \
        case 6: { constexpr int CC_ = 6; CALL; } break;                              \
        case 7: { constexpr int CC_ = 7; CALL; } break;                              \
        case 8: { constexpr int CC_ = 8; CALL; } break;                              \
        default: return bss_fail(h, BSS_EINVAL, "n_channels must be between 2 and 8"); \
    }

int launch_ip(bss_handle* h, const IpArgs& a) {
    int rc = BSS_OK;
    BSS_DISPATCH_C(a.C, rc = launch_ip_t<CC_>(h, a, a.order))
    return rc;
}

int launch_pb_scale(bss_handle* h, const double2* W, const double* Cx, double2* scale, int B, int F, int C, int ref) {
    const long long n = (long long)B * F;
    const unsigned grid = (unsigned)cdiv(n, 64);
    BSS_DISPATCH_C(C, (pb_scale_kernel<CC_><<<grid, 64, 0, h->stream>>>(W, Cx, scale, B, F, ref, h->flags)))
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_logdet(bss_handle* h, const double2* W, double* out, long long n_bins, int C, int transpose_sq) {
    const unsigned grid = (unsigned)cdiv(n_bins, 64);
    BSS_DISPATCH_C(C, (logdet_kernel<CC_><<<grid, 64, 0, h->stream>>>(W, out, n_bins, transpose_sq)))
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_lsq_map(bss_handle* h, const double2* A, const double2* Bm, double2* out, int Ra, int Rb, int F, int T) {
    switch (Rb) {
        case 1: lsq_map_kernel<1><<<F, 128, 0, h->stream>>>(A, Bm, out, Ra, F, T, h->flags); break;
        case 2: lsq_map_kernel<2><<<F, 128, 0, h->stream>>>(A, Bm, out, Ra, F, T, h->flags); break;
        case 3: lsq_map_kernel<3><<<F, 128, 0, h->stream>>>(A, Bm, out, Ra, F, T, h->flags); break;
        case 4: lsq_map_kernel<4><<<F, 128, 0, h->stream>>>(A, Bm, out, Ra, F, T, h->flags); break;
        case 5: lsq_map_kernel<5><<<F, 128, 0, h->stream>>>(A, Bm, out, Ra, F, T, h->flags); break;
        case 6: lsq_map_kernel<6><<<F, 128, 0, h->stream>>>(A, Bm, out, Ra, F, T, h->flags); break;
        case 7: lsq_map_kernel<7><<<F, 128, 0, h->stream>>>(A, Bm, out, Ra, F, T, h->flags); break;
        case 8: lsq_map_kernel<8><<<F, 128, 0, h->stream>>>(A, Bm, out, Ra, F, T, h->flags); break;
        default: return bss_fail(h, BSS_EINVAL, "least-squares map: 1 to 8 rows");
    }
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

int launch_lsq_filter(bss_handle* h, const double2* G, const double* Cx, double2* W, long long n_bins, int C) {
    const unsigned grid = (unsigned)cdiv(n_bins, 64);
    BSS_DISPATCH_C(C, (lsq_filter_kernel<CC_><<<grid, 64, 0, h->stream>>>(G, Cx, W, n_bins, h->flags)))
    h->launches++;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}
