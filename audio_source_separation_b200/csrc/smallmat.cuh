// Per-bin dense complex linear algebra in fp64 registers (C x C, C <= 8).
// Everything is __host__ __device__ so the same code is exercised on the CPU by
// tests/hostmath (test harness only; the product calls these from kernels).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SM_HD __host__ __device__ __forceinline__
#else
#define SM_HD inline
#endif

struct cd {
    double x, y;
};
SM_HD cd cd_make(double a, double b) {
    cd r;
    r.x = a;
    r.y = b;
    return r;
}
SM_HD cd operator+(cd a, cd b) { return cd_make(a.x + b.x, a.y + b.y); }
SM_HD cd operator-(cd a, cd b) { return cd_make(a.x - b.x, a.y - b.y); }
SM_HD cd operator*(cd a, cd b) { return cd_make(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SM_HD cd operator*(double s, cd a) { return cd_make(s * a.x, s * a.y); }
SM_HD cd cd_conj(cd a) { return cd_make(a.x, -a.y); }
SM_HD double cd_abs2(cd a) { return a.x * a.x + a.y * a.y; }
SM_HD double cd_abs(cd a) { return hypot(a.x, a.y); }
// a / b, Smith's algorithm (what NumPy uses for complex division)
SM_HD cd cd_div(cd a, cd b) {
    if (fabs(b.x) >= fabs(b.y)) {
        if (b.x == 0.0 && b.y == 0.0) return cd_make(a.x / fabs(b.x), a.y / fabs(b.x));
        const double r = b.y / b.x, s = 1.0 / (b.x + b.y * r);
        return cd_make((a.x + a.y * r) * s, (a.y - a.x * r) * s);
    }
    const double r = b.x / b.y, s = 1.0 / (b.x * r + b.y);
    return cd_make((a.x * r + a.y) * s, (a.y * r - a.x) * s);
}
// principal square root
SM_HD cd cd_sqrt(cd a) {
    if (a.x == 0.0 && a.y == 0.0) return cd_make(0.0, a.y);
    const double m = hypot(a.x, a.y);
    if (a.x >= 0.0) {
        const double t = sqrt(0.5 * (m + a.x));
        return cd_make(t, a.y / (2.0 * t));
    }
    const double t = sqrt(0.5 * (m - a.x));
    return cd_make(fabs(a.y) / (2.0 * t), a.y >= 0.0 ? t : -t);
}
// acc += a * b
SM_HD void cd_fma(cd& acc, cd a, cd b) {
    acc.x += a.x * b.x - a.y * b.y;
    acc.y += a.x * b.y + a.y * b.x;
}
// NumPy orders complex numbers lexicographically (real, then imag): `den[den < eps] = eps`
SM_HD bool cd_less_real(cd a, double e) { return a.x < e || (a.x == e && a.y < 0.0); }

// a > b in NumPy's complex order (real part first, then imaginary).  np.argsort(lam)[::-1] of two eigenvalues puts index 0
// first exactly when lam[0] > lam[1]: on a tie the ascending (stable for two elements) sort keeps [0, 1] and the reversal
// gives [1, 0]  (src/bss/ilrma.py:610, src/bss/iva.py:576).
SM_HD bool cd_lex_greater(cd a, cd b) { return a.x > b.x || (a.x == b.x && a.y > b.y); }

template <int C>
struct Mat {
    cd a[C][C];
};

template <int C>
SM_HD void mat_mul(const Mat<C>& A, const Mat<C>& B, Mat<C>& R) {
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) {
            cd s = cd_make(0.0, 0.0);
#pragma unroll
            for (int k = 0; k < C; ++k) cd_fma(s, A.a[i][k], B.a[k][j]);
            R.a[i][j] = s;
        }
}

template <int C>
SM_HD double mat_fro2(const Mat<C>& A) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) s += cd_abs2(A.a[i][j]);
    return s;
}

// Gauss-Jordan inverse with partial pivoting.  Returns false on an exactly zero pivot
// (LAPACK's `info > 0`, which NumPy turns into LinAlgError("Singular matrix")).
template <int C>
SM_HD bool mat_inverse(const Mat<C>& Ain, Mat<C>& Inv) {
    Mat<C> A = Ain;
#pragma unroll
    for (int i = 0; i < C; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) Inv.a[i][j] = cd_make(i == j ? 1.0 : 0.0, 0.0);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        // pivot search (LAPACK izamax uses |re| + |im|)
        int p = k;
        double best = fabs(A.a[k][k].x) + fabs(A.a[k][k].y);
#pragma unroll
        for (int i = k + 1; i < C; ++i) {
            const double v = fabs(A.a[i][k].x) + fabs(A.a[i][k].y);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        if (best == 0.0) ok = false;
        // row swap k <-> p (compile-time indices only: predicated exchange)
#pragma unroll
        for (int i = k + 1; i < C; ++i) {
            if (i == p) {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    cd t = A.a[k][j];
                    A.a[k][j] = A.a[i][j];
                    A.a[i][j] = t;
                    t = Inv.a[k][j];
                    Inv.a[k][j] = Inv.a[i][j];
                    Inv.a[i][j] = t;
                }
            }
        }
        const cd piv = cd_div(cd_make(1.0, 0.0), A.a[k][k]);
#pragma unroll
        for (int j = 0; j < C; ++j) {
            A.a[k][j] = A.a[k][j] * piv;
            Inv.a[k][j] = Inv.a[k][j] * piv;
        }
#pragma unroll
        for (int i = 0; i < C; ++i) {
            if (i == k) continue;
            const cd f = A.a[i][k];
#pragma unroll
            for (int j = 0; j < C; ++j) {
                A.a[i][j] = A.a[i][j] - f * A.a[k][j];
                Inv.a[i][j] = Inv.a[i][j] - f * Inv.a[k][j];
            }
        }
    }
    return ok;
}

template <int C>
SM_HD cd mat_det(const Mat<C>& Ain) {
    Mat<C> A = Ain;
    cd det = cd_make(1.0, 0.0);
#pragma unroll
    for (int k = 0; k < C; ++k) {
        int p = k;
        double best = fabs(A.a[k][k].x) + fabs(A.a[k][k].y);
#pragma unroll
        for (int i = k + 1; i < C; ++i) {
            const double v = fabs(A.a[i][k].x) + fabs(A.a[i][k].y);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        if (best == 0.0) return cd_make(0.0, 0.0);
#pragma unroll
        for (int i = k + 1; i < C; ++i) {
            if (i == p) {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    cd t = A.a[k][j];
                    A.a[k][j] = A.a[i][j];
                    A.a[i][j] = t;
                }
                det = cd_make(-det.x, -det.y);
            }
        }
        det = det * A.a[k][k];
        const cd piv = cd_div(cd_make(1.0, 0.0), A.a[k][k]);
#pragma unroll
        for (int i = k + 1; i < C; ++i) {
            const cd f = A.a[i][k] * piv;
#pragma unroll
            for (int j = k; j < C; ++j) A.a[i][j] = A.a[i][j] - f * A.a[k][j];
        }
    }
    return det;
}

// 2-norm condition number by one-sided (Hestenes) Jacobi on the columns: accurate for the small
// singular values, unlike an eigen-decomposition of A^H A.  Only used in the narrow band where
// the cheap Frobenius bounds cannot decide the gate.
template <int C>
SM_HD double mat_cond2(const Mat<C>& Ain) {
    Mat<C> A = Ain;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < C - 1; ++p)
            for (int q = p + 1; q < C; ++q) {
                double app = 0.0, aqq = 0.0;
                cd apq = cd_make(0.0, 0.0);
                for (int i = 0; i < C; ++i) {
                    app += cd_abs2(A.a[i][p]);
                    aqq += cd_abs2(A.a[i][q]);
                    cd_fma(apq, cd_conj(A.a[i][p]), A.a[i][q]);
                }
                const double g = cd_abs(apq);
                if (g == 0.0 || g <= 1e-16 * sqrt(app * aqq)) continue;
                off = fmax(off, g / sqrt(app * aqq));
                // rotation that zeroes the (p,q) inner product
                const cd ph = cd_make(apq.x / g, apq.y / g);      // e^{i arg}
                const double zeta = (aqq - app) / (2.0 * g);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < C; ++i) {
                    const cd xp = A.a[i][p];
                    const cd xq = A.a[i][q] * cd_conj(ph);
                    A.a[i][p] = cd_make(c * xp.x - s * xq.x, c * xp.y - s * xq.y);
                    A.a[i][q] = cd_make(s * xp.x + c * xq.x, s * xp.y + c * xq.y);
                }
            }
        if (off < 1e-15) break;
    }
    double smax = 0.0, smin = 1e300;
    for (int j = 0; j < C; ++j) {
        double n2 = 0.0;
        for (int i = 0; i < C; ++i) n2 += cd_abs2(A.a[i][j]);
        const double s = sqrt(n2);
        smax = fmax(smax, s);
        smin = fmin(smin, s);
    }
    return smin > 0.0 ? smax / smin : INFINITY;
}

// Decide `cond_2(A) < threshold` given A and its inverse.  kappa_F = |A|_F |A^-1|_F satisfies
// kappa_F / C <= cond_2 <= kappa_F, so only C-wide band around the threshold needs the SVD.
template <int C>
SM_HD bool cond_below(const Mat<C>& A, const Mat<C>& Ainv, bool invertible, double threshold) {
    if (!invertible) return false;   // numpy: cond = inf
    const double kf = sqrt(mat_fro2(A) * mat_fro2(Ainv));
    if (!(kf == kf)) return false;
    if (kf < threshold) return true;
    if (kf >= threshold * C) return false;
    return mat_cond2(A) < threshold;
}

// unpack a Hermitian matrix stored as C real diagonals followed by the strictly-lower triangle
// (row-major pairs (i,j), i > j) as (re, im)
template <int C>
SM_HD void herm_unpack(const double* p, Mat<C>& U) {
    int e = C;
#pragma unroll
    for (int i = 0; i < C; ++i) U.a[i][i] = cd_make(p[i], 0.0);
#pragma unroll
    for (int i = 1; i < C; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) {
            const cd v = cd_make(p[e], p[e + 1]);
            e += 2;
            U.a[i][j] = v;
            U.a[j][i] = cd_conj(v);
        }
}

// One iterative-projection row update (src/bss/ilrma.py:516-528; floor_den: src/bss/mnmf.py:883,
// src/bss/ilrma.py:981).  W row n is replaced in place when the gate passes.
// use_gate = false follows tILRMA (plain inverse, no condition test, src/bss/ilrma.py:975).
// Returns 1 (updated), 0 (kept by the gate); *singular is set on an exactly singular W U.
// w = A^-1 e_n and |det A|^2 by Gaussian elimination with partial pivoting of [A | e_n] (the pivoting rule of mat_inverse),
// a third of the work of the full inverse.  Returns false on an exactly zero pivot.
template <int C>
SM_HD bool solve_unit_inplace(Mat<C>& A, int n, cd (&w)[C], double* absdet2) {   // A is overwritten by the elimination
    cd b[C], rp[C];
#pragma unroll
    for (int i = 0; i < C; ++i) b[i] = cd_make(i == n ? 1.0 : 0.0, 0.0);
    bool ok = true;
    double ad2 = 1.0;
#pragma unroll
    for (int k = 0; k < C; ++k) {
        int p = k;
        double best = fabs(A.a[k][k].x) + fabs(A.a[k][k].y);
#pragma unroll
        for (int i = k + 1; i < C; ++i) {
            const double v = fabs(A.a[i][k].x) + fabs(A.a[i][k].y);
            if (v > best) {
                best = v;
                p = i;
            }
        }
        if (best == 0.0) ok = false;
#pragma unroll
        for (int i = k + 1; i < C; ++i) {
            if (i == p) {
#pragma unroll
                for (int j = k; j < C; ++j) {
                    const cd t = A.a[k][j];
                    A.a[k][j] = A.a[i][j];
                    A.a[i][j] = t;
                }
                const cd t = b[k];
                b[k] = b[i];
                b[i] = t;
            }
        }
        ad2 *= cd_abs2(A.a[k][k]);
        rp[k] = cd_div(cd_make(1.0, 0.0), A.a[k][k]);
#pragma unroll
        for (int i = k + 1; i < C; ++i) {
            const cd f = A.a[i][k] * rp[k];
#pragma unroll
            for (int j = k + 1; j < C; ++j) A.a[i][j] = A.a[i][j] - f * A.a[k][j];
            b[i] = b[i] - f * b[k];
        }
    }
#pragma unroll
    for (int i = C - 1; i >= 0; --i) {
        cd s = b[i];
#pragma unroll
        for (int j = i + 1; j < C; ++j) s = s - A.a[i][j] * w[j];
        w[i] = s * rp[i];
    }
    *absdet2 = ad2;
    return ok;
}

template <int C>
SM_HD bool solve_unit(const Mat<C>& Ain, int n, cd (&w)[C], double* absdet2) {
    Mat<C> A = Ain;
    return solve_unit_inplace(A, n, w, absdet2);
}

// The exact route of a row update: full inverse, gate decision from cond_2 (Frobenius bounds, Jacobi SVD in the undecided
// band), w = column n of the inverse.  Rare (ill-conditioned, near-threshold, singular or non-finite bins only), so it is kept
// out of line: inlined, its two extra matrices pushed the common path of the thread-per-bin kernel into local memory.
// Returns 1 (update), 0 (kept by the gate) or -1 (exactly singular).
#if defined(__CUDACC__)
#define SM_NOINLINE __host__ __device__ __noinline__
#else
#define SM_NOINLINE inline
#endif
template <int C>
SM_NOINLINE int ip_row_exact(const Mat<C>* Ap, int n, double threshold, bool use_gate, cd* w) {
    const Mat<C>& A = *Ap;
    Mat<C> Ainv;
    const bool inv_ok = mat_inverse(A, Ainv);
    if (!inv_ok) return -1;
    const bool ok = use_gate ? cond_below(A, Ainv, inv_ok, threshold) : true;
    // w = A^-1 e_n : column n of the inverse (compile-time indexed select)
#pragma unroll
    for (int i = 0; i < C; ++i) {
        w[i] = Ainv.a[i][0];
#pragma unroll
        for (int j = 1; j < C; ++j)
            if (j == n) w[i] = Ainv.a[i][j];
    }
    return ok ? 1 : 0;
}

// The row update given a way to form A = W U (`make_A(A)`; called a second time only on the rare exact route, because the
// elimination of the common route overwrites A in place -- a kept copy costs 64 registers): on return `row` holds the new
// row n (conj(w) / sqrt(w^H U w)) when the result is 1.
template <int C, class MakeA>
SM_HD int ip_row_from_product(const MakeA& make_A, const Mat<C>& U, int n, double threshold, bool use_gate, bool floor_den, double eps,
                              bool* singular, cd (&row)[C]) {
    cd w[C];
    bool ok = true;
    bool solved = false;
    if (use_gate) {
        // Almost every bin passes the gate by a wide margin.  cond_2(A) <= 2 (|A|_F / sqrt(C))^C / |det A| (Guggenheimer,
        // Edelman, Johnson) needs only the determinant, which the elimination for w = A^-1 e_n yields anyway: when that bound
        // is already under half the threshold the gate passes for certain and the inverse is never formed.  Everything else
        // (near the threshold, above it, singular, non-finite) takes the exact route below, so the decisions are unchanged.
        Mat<C> A;
        make_A(A);
        double g = mat_fro2(A) / (double)C;   // (|A|_F^2 / C)^C
        double gc = g;
#pragma unroll
        for (int i = 1; i < C; ++i) gc *= g;
        double ad2 = 0.0;
        const bool piv_ok = solve_unit_inplace(A, n, w, &ad2);
        solved = piv_ok && (4.0 * gc < 0.25 * threshold * threshold * ad2);
    }
    if (!solved) {
        Mat<C> Ax;          // the out-of-line call takes addresses: only these objects live in local memory
        make_A(Ax);
        cd wx[C];
        const int r = ip_row_exact<C>(&Ax, n, threshold, use_gate, wx);
        if (r < 0) {
            *singular = true;
            return 0;
        }
        ok = r == 1;
#pragma unroll
        for (int i = 0; i < C; ++i) w[i] = wx[i];
    }
    // q = w^H U w
    cd q = cd_make(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < C; ++i) {
        cd s = cd_make(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < C; ++j) cd_fma(s, U.a[i][j], w[j]);
        cd_fma(q, cd_conj(w[i]), s);
    }
    cd den = cd_sqrt(q);
    if (floor_den && cd_less_real(den, eps)) den = cd_make(eps, 0.0);
#pragma unroll
    for (int j = 0; j < C; ++j) row[j] = cd_div(cd_conj(w[j]), den);
    return ok ? 1 : 0;
}

template <int C>
struct ProductOfRegisters {   // A = W U with both factors in registers (mat_mul)
    const Mat<C>& W;
    const Mat<C>& U;
    SM_HD void operator()(Mat<C>& A) const { mat_mul(W, U, A); }
};

template <int C>
SM_HD int ip_row(Mat<C>& W, const Mat<C>& U, int n, double threshold, bool use_gate, bool floor_den, double eps,
                 bool* singular) {
    cd row[C];
    const ProductOfRegisters<C> make_A{W, U};
    const int ok = ip_row_from_product<C>(make_A, U, n, threshold, use_gate, floor_den, eps, singular, row);
    if (ok) {
#pragma unroll
        for (int r = 0; r < C; ++r) {
            if (r == n) {
#pragma unroll
                for (int j = 0; j < C; ++j) W.a[r][j] = row[j];
            }
        }
    }
    return ok;
}

// ------------------------------------------------------------------------------ Hermitian eigen / Riccati
// Two-sided (classical cyclic) Jacobi for a Hermitian matrix: A = V diag(w) V^H.  Only the Hermitian part of the
// input is used.  A pair (p,q) is first made real by the phase of a_pq, then rotated by the real Jacobi angle.
template <int C>
SM_HD void herm_eig(const Mat<C>& Ain, Mat<C>& V, double* w) {
    Mat<C> A;
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) {
            const cd a = Ain.a[i][j], b = cd_conj(Ain.a[j][i]);
            A.a[i][j] = cd_make(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
            V.a[i][j] = cd_make(i == j ? 1.0 : 0.0, 0.0);
        }
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < C; ++i) {
            diag += A.a[i][i].x * A.a[i][i].x;
            for (int j = 0; j < i; ++j) off += cd_abs2(A.a[i][j]);
        }
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < C - 1; ++p)
            for (int q = p + 1; q < C; ++q) {
                const cd apq = A.a[p][q];
                const double g = cd_abs(apq);
                if (g == 0.0) continue;
                const cd ph = cd_make(apq.x / g, apq.y / g);                 // e^{i phi}
                const double tau = (A.a[q][q].x - A.a[p][p].x) / (2.0 * g);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                // columns: (a_ip, a_iq) <- (c a_ip - s e^{-i phi} a_iq, s a_ip + c e^{-i phi} a_iq); same for V
                for (int i = 0; i < C; ++i) {
                    cd xp = A.a[i][p], xq = A.a[i][q] * cd_conj(ph);
                    A.a[i][p] = cd_make(c * xp.x - s * xq.x, c * xp.y - s * xq.y);
                    A.a[i][q] = cd_make(s * xp.x + c * xq.x, s * xp.y + c * xq.y);
                    xp = V.a[i][p];
                    xq = V.a[i][q] * cd_conj(ph);
                    V.a[i][p] = cd_make(c * xp.x - s * xq.x, c * xp.y - s * xq.y);
                    V.a[i][q] = cd_make(s * xp.x + c * xq.x, s * xp.y + c * xq.y);
                }
                // rows: the conjugate transpose of the same rotation from the left
                for (int j = 0; j < C; ++j) {
                    const cd xp = A.a[p][j], xq = A.a[q][j] * ph;
                    A.a[p][j] = cd_make(c * xp.x - s * xq.x, c * xp.y - s * xq.y);
                    A.a[q][j] = cd_make(s * xp.x + c * xq.x, s * xp.y + c * xq.y);
                }
                A.a[p][q] = cd_make(0.0, 0.0);
                A.a[q][p] = cd_make(0.0, 0.0);
                A.a[p][p].y = 0.0;
                A.a[q][q].y = 0.0;
            }
    }
    for (int i = 0; i < C; ++i) w[i] = A.a[i][i].x;
}

// R = V diag(f) V^H
template <int C>
SM_HD void herm_compose(const Mat<C>& V, const double* f, Mat<C>& R) {
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) {
            cd s = cd_make(0.0, 0.0);
            for (int k = 0; k < C; ++k) cd_fma(s, f[k] * V.a[i][k], cd_conj(V.a[j][k]));
            R.a[i][j] = s;
        }
}

// The positive semi-definite solution of H A H = B for Hermitian positive (semi-)definite A, B:
// H = A^-1/2 (A^1/2 B A^1/2)^1/2 A^-1/2, Hermitian part.  Replaces solve_Riccati (src/algorithm/linalg.py:7-30), which
// takes the same solution from the stable invariant subspace of the 2C x 2C matrix [[0,-A],[-B,0]].
template <int C>
SM_HD void riccati_hermitian(const Mat<C>& A, const Mat<C>& B, Mat<C>& H) {
    Mat<C> V, S, Si, M, R;
    double w[C], f[C];
    herm_eig(A, V, w);
    for (int i = 0; i < C; ++i) f[i] = sqrt(fmax(w[i], 0.0));
    herm_compose(V, f, S);
    for (int i = 0; i < C; ++i) f[i] = 1.0 / f[i];
    herm_compose(V, f, Si);
    mat_mul(S, B, R);
    mat_mul(R, S, M);
    herm_eig(M, V, w);
    for (int i = 0; i < C; ++i) f[i] = sqrt(fmax(w[i], 0.0));
    herm_compose(V, f, M);
    mat_mul(Si, M, R);
    mat_mul(R, Si, M);
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) {
            const cd a = M.a[i][j], b = cd_conj(M.a[j][i]);
            H.a[i][j] = cd_make(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
        }
}
