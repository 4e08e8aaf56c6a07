// Test harness only: compiles the __host__ __device__ per-bin linear algebra of
// audio_source_separation_b200/csrc/smallmat.cuh for the CPU so that the exact code the kernels run
// can be checked against the oracle without a GPU.  Never linked into the product.
#include <cstdint>
#include <cstring>
#include "../../audio_source_separation_b200/csrc/smallmat.cuh"

template <int C>
static void load(const double* src, Mat<C>& M) {
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) M.a[i][j] = cd_make(src[(i * C + j) * 2], src[(i * C + j) * 2 + 1]);
}
template <int C>
static void store(double* dst, const Mat<C>& M) {
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j) {
            dst[(i * C + j) * 2] = M.a[i][j].x;
            dst[(i * C + j) * 2 + 1] = M.a[i][j].y;
        }
}

// W (F,C,C) in/out, U (C,F,C,C) full Hermitian, gate (C,F)
template <int C>
static int sweep(int F, double* W, const double* U, int32_t* gate, double thr, int use_gate, int floor_den, double eps) {
    int n_singular = 0;
    for (int f = 0; f < F; ++f) {
        Mat<C> Wm;
        load<C>(W + (size_t)f * C * C * 2, Wm);
        bool singular = false;
        for (int n = 0; n < C; ++n) {
            Mat<C> Um;
            load<C>(U + ((size_t)n * F + f) * C * C * 2, Um);
            gate[(size_t)n * F + f] = ip_row<C>(Wm, Um, n, thr, use_gate != 0, floor_den != 0, eps, &singular);
        }
        if (singular) ++n_singular;
        store<C>(W + (size_t)f * C * C * 2, Wm);
    }
    return n_singular;
}

template <int C>
static double cond2(const double* A) {
    Mat<C> M;
    load<C>(A, M);
    return mat_cond2<C>(M);
}

template <int C>
static int inverse(const double* A, double* out) {
    Mat<C> M, I;
    load<C>(A, M);
    const bool ok = mat_inverse<C>(M, I);
    store<C>(out, I);
    return ok ? 1 : 0;
}

template <int C>
static void det(const double* A, double* out) {
    Mat<C> M;
    load<C>(A, M);
    const cd d = mat_det<C>(M);
    out[0] = d.x;
    out[1] = d.y;
}

template <int C>
static void riccati(const double* A, const double* B, double* out) {
    Mat<C> Am, Bm, H;
    load<C>(A, Am);
    load<C>(B, Bm);
    riccati_hermitian<C>(Am, Bm, H);
    store<C>(out, H);
}

template <int C>
static void eigh(const double* A, double* vecs, double* vals) {
    Mat<C> Am, V;
    load<C>(A, Am);
    herm_eig<C>(Am, V, vals);
    store<C>(vecs, V);
}

#define DISPATCH(C, EXPR)              \
    switch (C) {                       \
        case 2: { constexpr int K = 2; EXPR; } break; \
        case 3: { constexpr int K = 3; EXPR; } break; \
        case 4: { constexpr int K = 4; EXPR; } break; \
        case 5: { constexpr int K = 5; EXPR; } break; \
        case 6: { constexpr int K = 6; EXPR; } break; \
        case 7: { constexpr int K = 7; EXPR; } break; \
        case 8: { constexpr int K = 8; EXPR; } break; \
        default: return -1;            \
    }

extern "C" {
int hm_ip_sweep(int C, int F, double* W, const double* U, int32_t* gate, double thr, int use_gate, int floor_den, double eps) {
    int r = 0;
    DISPATCH(C, r = sweep<K>(F, W, U, gate, thr, use_gate, floor_den, eps))
    return r;
}
int hm_cond2(int C, const double* A, double* out) {
    DISPATCH(C, *out = cond2<K>(A))
    return 0;
}
int hm_inverse(int C, const double* A, double* out) {
    int r = 0;
    DISPATCH(C, r = inverse<K>(A, out))
    return r;
}
int hm_det(int C, const double* A, double* out) {
    DISPATCH(C, det<K>(A, out))
    return 0;
}
int hm_riccati(int C, const double* A, const double* B, double* out) {
    DISPATCH(C, riccati<K>(A, B, out))
    return 0;
}
int hm_eigh(int C, const double* A, double* vecs, double* vals) {
    DISPATCH(C, eigh<K>(A, vecs, vals))
    return 0;
}
}
