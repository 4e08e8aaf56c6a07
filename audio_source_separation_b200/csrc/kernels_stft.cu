// STFT feed on the GPU: scipy.signal.stft / istft exactly as src/transform/stft.py:4-17 calls them
// (window 'hann' periodic by default, boundary='zeros', padded=True, onesided, scaling='spectrum'):
//   stft : extend by fft/2 zeros on both sides, zero-pad the tail to a whole number of hops, frame, window,
//          real FFT, divide by sum(window)                                      -> (signals, fft/2+1, frames)
//   istft: inverse real FFT of every frame, times sum(window), window again, overlap-add, divide by the
//          overlap-added squared window where it exceeds 1e-10, drop the fft/2 extension at both ends
// One CTA transforms one frame: the real FFT of length N is a complex radix-2 Stockham FFT of length N/2 in
// shared memory (twiddles from a table rounded from float64) with the usual split/merge step.  The forward
// kernel can write straight into a handle's block-interleaved bin tiles, so a mixture is fed to the update
// loop from its waveform without the (twice larger) spectrogram ever crossing PCIe.
#include <cstring>

#include "handle.h"

namespace {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// Shared-memory index with one pad element per 16 (a 128-byte row of float2): the strided stores of the Stockham passes
// (stride radix x Ns elements) would otherwise fall on two banks.
__device__ __forceinline__ int fpad(int i) { return i + (i >> 4); }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// forward DFTs of size 2 / 4 / 8 in registers, outputs in natural order
__device__ __forceinline__ void dft4(float2& u0, float2& u1, float2& u2, float2& u3) {
    const float2 a0 = cadd(u0, u2), a1 = csub(u0, u2), a2 = cadd(u1, u3), a3 = mul_mi(csub(u1, u3));
    u0 = cadd(a0, a2);
    u1 = cadd(a1, a3);
    u2 = csub(a0, a2);
    u3 = csub(a1, a3);
}
__device__ __forceinline__ void dft8(float2 (&u)[8]) {
    float2 e0 = u[0], e1 = u[2], e2 = u[4], e3 = u[6], o0 = u[1], o1 = u[3], o2 = u[5], o3 = u[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    const float h = 0.70710678118654752440f;
    // odd[q] * exp(-2 pi i q / 8)
    o1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));      // (1 - i) / sqrt 2
    o2 = mul_mi(o2);                                             // -i
    o3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));     // (-1 - i) / sqrt 2
    u[0] = cadd(e0, o0);
    u[4] = csub(e0, o0);
    u[1] = cadd(e1, o1);
    u[5] = csub(e1, o1);
    u[2] = cadd(e2, o2);
    u[6] = csub(e2, o2);
    u[3] = cadd(e3, o3);
    u[7] = csub(e3, o3);
}

// One Stockham pass of radix R over M points: a -> b (both padded with fpad).  tw is the full circle exp(-2 pi i m / N),
// m < N = 2 M, so the twiddle exp(-2 pi i r k / (Ns R)) is tw[r k (N / (Ns R))].
template <int R>
__device__ __forceinline__ void stockham_pass(const float2* a, float2* b, const float2* tw, int M, int Ns) {
    const int per = M / R;
    const int tstep = (2 * M) / (Ns * R);
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
        const int k = j & (Ns - 1);
        float2 u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = a[fpad(j + r * per)];
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) u[r] = cmulf(u[r], __ldg(tw + r * k * tstep));
        }
        if constexpr (R == 8) {
            dft8(u);
        } else if constexpr (R == 4) {
            dft4(u[0], u[1], u[2], u[3]);
        } else {
            const float2 x = u[0], y = u[1];
            u[0] = cadd(x, y);
            u[1] = csub(x, y);
        }
        const int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) b[fpad(j0 + r * Ns)] = u[r];
    }
}

// forward FFT of length M = N / 2 (a power of two >= 4) held in `a` (padded layout, see fpad); ping-pongs between a and b
// and returns the buffer holding the result.  Radix-8 passes, then one radix-4 or radix-2 pass for what is left.
__device__ __forceinline__ float2* fft_stockham(float2* a, float2* b, const float2* tw, int N) {
    const int M = N >> 1;
    int Ns = 1;
    while (Ns < M) {
        const int rest = M / Ns;
        int R;
        if (rest >= 8 && rest != 16) {   // 16 = 4 x 4 (two passes either way, radix 4 keeps all threads busy)
            stockham_pass<8>(a, b, tw, M, Ns);
            R = 8;
        } else if (rest >= 4) {
            stockham_pass<4>(a, b, tw, M, Ns);
            R = 4;
        } else {
            stockham_pass<2>(a, b, tw, M, Ns);
            R = 2;
        }
        Ns *= R;
        __syncthreads();
        float2* t = a;
        a = b;
        b = t;
    }
    return a;
}

struct StftParams {
    const float* x;      // [S][n_samples]
    const float* win;    // [N]
    const float2* tw;    // [N]  exp(-2 pi i m / N), the full circle
    int S, n_samples, N, hop, n_frames;
    float scale;         // 1 / sum(window)
    double2* out128;     // [S][N/2+1][n_frames]   (reference layout) or null
    cf* X;               // handle tiles [B][F][C][Tp] block-interleaved, signal s = b C + c, or null
    int C, Tp;
};

__global__ void __launch_bounds__(1024) stft_kernel(const StftParams p) {
    extern __shared__ __align__(16) float2 fft_smem[];
    const int N = p.N, M = N >> 1;
    float2* a = fft_smem;
    float2* b = fft_smem + fpad(M) + 1;
    const int frame = blockIdx.x, s = blockIdx.y;
    const float* x = p.x + (size_t)s * p.n_samples;
    const int start = frame * p.hop - (N >> 1);   // position of the frame in the unextended signal
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const int i0 = start + 2 * j, i1 = i0 + 1;
        const float v0 = (i0 >= 0 && i0 < p.n_samples) ? x[i0] * p.win[2 * j] : 0.f;
        const float v1 = (i1 >= 0 && i1 < p.n_samples) ? x[i1] * p.win[2 * j + 1] : 0.f;
        a[fpad(j)] = make_float2(v0, v1);
    }
    __syncthreads();
    const float2* Z = fft_stockham(a, b, p.tw, N);
    for (int k = threadIdx.x; k <= M; k += blockDim.x) {
        const float2 zk = Z[fpad(k == M ? 0 : k)];
        const float2 zm = Z[fpad(k == 0 ? 0 : M - k)];
        const float2 w = k < M ? __ldg(p.tw + k) : make_float2(-1.f, 0.f);
        // X[k] = (Zk + conj(Zm)) / 2 - (i / 2) w (Zk - conj(Zm))
        const float2 e = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        const float2 o = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
        const float2 wo = cmulf(w, o);
        const float2 v = make_float2((e.x + wo.y) * p.scale, (e.y - wo.x) * p.scale);
        if (p.out128) p.out128[((size_t)s * (M + 1) + k) * p.n_frames + frame] = make_double2((double)v.x, (double)v.y);
        if (p.X) {
            const int bb = s / p.C, c = s - bb * p.C;
            p.X[((size_t)bb * (M + 1) + k) * p.C * p.Tp + tile_off(p.C, p.Tp, c, frame)] = v;
        }
    }
}

template <typename TZ>
struct IstftParams {
    const TZ* z;         // [S][N/2+1][n_frames]  (double2 or float2)
    const float* win;
    const float2* tw;
    int S, N, hop, n_frames;
    float scale;         // sum(window) / (N / 2)   (spectrum scaling and the 1 / M of the inverse transform)
    float* frames;       // [S][n_frames][N]
};

template <typename TZ>
__global__ void __launch_bounds__(1024) istft_frames_kernel(const IstftParams<TZ> p) {
    extern __shared__ __align__(16) float2 fft_smem[];
    const int N = p.N, M = N >> 1;
    float2* a = fft_smem;
    float2* b = fft_smem + fpad(M) + 1;
    const int frame = blockIdx.x, s = blockIdx.y;
    const TZ* z = p.z + (size_t)s * (M + 1) * p.n_frames + frame;
    for (int k = threadIdx.x; k < M; k += blockDim.x) {
        const TZ xk = z[(size_t)k * p.n_frames];
        const TZ xm = z[(size_t)(M - k) * p.n_frames];
        const float2 w = __ldg(p.tw + k);
        // Z[k] = (Xk + conj(Xm)) / 2 + (i / 2) conj(w) (Xk - conj(Xm));  stored conjugated for the inverse transform
        const float2 e = make_float2(0.5f * (float)(xk.x + xm.x), 0.5f * (float)(xk.y - xm.y));
        const float2 o = make_float2(0.5f * (float)(xk.x - xm.x), 0.5f * (float)(xk.y + xm.y));
        const float2 wo = cmulf(make_float2(w.x, -w.y), o);
        a[fpad(k)] = make_float2(e.x - wo.y, -(e.y + wo.x));
    }
    __syncthreads();
    const float2* Z = fft_stockham(a, b, p.tw, N);
    float* out = p.frames + ((size_t)s * p.n_frames + frame) * N;
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const float2 v = Z[fpad(j)];   // conj(v) / M is the time signal pair
        out[2 * j] = v.x * p.scale * p.win[2 * j];
        out[2 * j + 1] = -v.y * p.scale * p.win[2 * j + 1];
    }
}

// overlap-add, squared-window normalisation, boundary trim
template <typename TO>
__global__ void __launch_bounds__(256) istft_ola_kernel(const float* frames, const float* win, TO* out, int S, int N, int hop,
                                                        int n_frames, int out_len) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)S * out_len) return;
    const int i = (int)(idx % out_len);
    const int s = (int)(idx / out_len);
    const int n = i + (N >> 1);
    int lo = (n - N + hop) / hop;   // ceil((n - N + 1) / hop)
    if (n - N + 1 <= 0) lo = 0;
    int hi = n / hop;
    if (hi > n_frames - 1) hi = n_frames - 1;
    float acc = 0.f, norm = 0.f;
    for (int ii = lo; ii <= hi; ++ii) {
        const int r = n - ii * hop;
        acc += frames[((size_t)s * n_frames + ii) * N + r];
        norm += win[r] * win[r];
    }
    out[idx] = (TO)(norm > 1e-10f ? acc / norm : acc);
}

__global__ void __launch_bounds__(256) to_float_kernel(const double* in, float* out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}

__global__ void __launch_bounds__(256) pcm16_to_float_kernel(const short* in, float* out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i] * (1.0f / 32768.0f);
}

bool pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

struct FftTables {
    float* win = nullptr;
    float2* tw = nullptr;
    double win_sum = 0.0;
};

int make_tables(const double* window, int N, cudaStream_t stream, FftTables* t) {
    std::vector<float> w(N);
    std::vector<float2> tw(N);   // the full circle: the radix-8 passes use exp(-2 pi i m / N) up to m = 7 N / 8
    double sum = 0.0;
    for (int i = 0; i < N; ++i) {
        w[i] = (float)window[i];
        sum += window[i];
    }
    const double pi = 3.14159265358979323846;
    for (int m = 0; m < N; ++m) tw[m] = make_float2((float)cos(-2.0 * pi * m / N), (float)sin(-2.0 * pi * m / N));
    if (cudaMalloc(&t->win, N * sizeof(float)) != cudaSuccess) return BSS_ENOMEM;
    if (cudaMalloc(&t->tw, N * sizeof(float2)) != cudaSuccess) return BSS_ENOMEM;
    cudaMemcpyAsync(t->win, w.data(), N * sizeof(float), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(t->tw, tw.data(), N * sizeof(float2), cudaMemcpyHostToDevice, stream);
    cudaStreamSynchronize(stream);   // the host vectors go out of scope
    t->win_sum = sum;
    return BSS_OK;
}

void free_tables(FftTables* t) {
    if (t->win) cudaFree(t->win);
    if (t->tw) cudaFree(t->tw);
}

// one thread per radix-8 butterfly of the M = N / 2 point transform
int fft_threads(int N) {
    int th = N / 16;
    if (th < 32) th = 32;
    if (th > 1024) th = 1024;
    return th;
}
// two padded buffers of M = N / 2 float2 (fpad)
size_t fft_smem_bytes(int N) { return (size_t)(2 * ((N / 2) + (N / 2) / 16 + 2)) * sizeof(float2); }

}  // namespace

extern "C" {

int bss_stft_frames(int n_samples, int fft_size, int hop_size) {
    if (n_samples < 0 || fft_size < 2 || hop_size < 1 || hop_size > fft_size) return BSS_EINVAL;
    const long long ext = (long long)n_samples + 2 * (fft_size / 2);
    long long nadd = (-(ext - fft_size)) % hop_size;
    if (nadd < 0) nadd += hop_size;
    nadd %= fft_size;
    return (int)((ext + nadd - fft_size) / hop_size + 1);
}

int bss_istft_length(int n_frames, int fft_size, int hop_size) {
    if (n_frames < 1 || fft_size < 2 || hop_size < 1) return BSS_EINVAL;
    return fft_size + (n_frames - 1) * hop_size - 2 * (fft_size / 2);
}

static int stft_common(int device, cudaStream_t stream, int S, int n_samples, int fft_size, int hop_size, const double* window,
                       const void* x, int x_dtype, double2* out128_dev, cf* X, int C, int Tp, int n_frames) {
    (void)device;
    FftTables t;
    int rc = make_tables(window, fft_size, stream, &t);
    if (rc != BSS_OK) {
        free_tables(&t);
        return rc;
    }
    const long long n = (long long)S * n_samples;
    float* xf = nullptr;
    void* staged = nullptr;
    if (cudaMalloc(&xf, (n > 0 ? n : 1) * sizeof(float)) != cudaSuccess) {
        free_tables(&t);
        return BSS_ENOMEM;
    }
    if (x_dtype == BSS_F32) {
        cudaMemcpyAsync(xf, x, n * sizeof(float), cudaMemcpyHostToDevice, stream);
    } else {
        if (cudaMalloc(&staged, (n > 0 ? n : 1) * sizeof(double)) != cudaSuccess) {
            cudaFree(xf);
            free_tables(&t);
            return BSS_ENOMEM;
        }
        cudaMemcpyAsync(staged, x, n * sizeof(double), cudaMemcpyHostToDevice, stream);
        to_float_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>((const double*)staged, xf, n);
    }
    StftParams p{};
    p.x = xf;
    p.win = t.win;
    p.tw = t.tw;
    p.S = S;
    p.n_samples = n_samples;
    p.N = fft_size;
    p.hop = hop_size;
    p.n_frames = n_frames;
    p.scale = (float)(1.0 / t.win_sum);
    p.out128 = out128_dev;
    p.X = X;
    p.C = C;
    p.Tp = Tp;
    const size_t smem = fft_smem_bytes(fft_size);
    cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(n_frames, S);
    stft_kernel<<<grid, fft_threads(fft_size), smem, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(xf);
    if (staged) cudaFree(staged);
    free_tables(&t);
    return e == cudaSuccess ? BSS_OK : BSS_ECUDA;
}

int bss_stft(int device, int n_signals, int n_samples, int fft_size, int hop_size, const double* window, const double* x, void* out) {
    if (!window || !x || !out || n_signals < 1 || n_samples < 1) return BSS_EINVAL;
    if (!pow2(fft_size) || fft_size < 8 || fft_size > 16384) return BSS_EUNSUPPORTED;
    const int n_frames = bss_stft_frames(n_samples, fft_size, hop_size);
    if (n_frames < 1) return BSS_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return BSS_ECUDA;
    const size_t n_out = (size_t)n_signals * (fft_size / 2 + 1) * n_frames;
    double2* dev = nullptr;
    if (cudaMalloc(&dev, n_out * sizeof(double2)) != cudaSuccess) return BSS_ENOMEM;
    int rc = stft_common(device, 0, n_signals, n_samples, fft_size, hop_size, window, x, BSS_F64, dev, nullptr, 1, 0, n_frames);
    if (rc == BSS_OK && cudaMemcpy(out, dev, n_out * sizeof(double2), cudaMemcpyDeviceToHost) != cudaSuccess) rc = BSS_ECUDA;
    cudaFree(dev);
    return rc;
}

int bss_istft(int device, int n_signals, int n_frames, int fft_size, int hop_size, const double* window, const void* z, double* out) {
    if (!window || !z || !out || n_signals < 1 || n_frames < 1 || hop_size < 1 || hop_size > fft_size) return BSS_EINVAL;
    if (!pow2(fft_size) || fft_size < 8 || fft_size > 16384) return BSS_EUNSUPPORTED;
    if (cudaSetDevice(device) != cudaSuccess) return BSS_ECUDA;
    const int out_len = bss_istft_length(n_frames, fft_size, hop_size);
    if (out_len < 1) return BSS_EINVAL;
    FftTables t;
    int rc = make_tables(window, fft_size, 0, &t);
    const size_t n_in = (size_t)n_signals * (fft_size / 2 + 1) * n_frames;
    double2* zd = nullptr;
    float* frames = nullptr;
    double* od = nullptr;
    if (rc == BSS_OK && (cudaMalloc(&zd, n_in * sizeof(double2)) != cudaSuccess ||
                         cudaMalloc(&frames, (size_t)n_signals * n_frames * fft_size * sizeof(float)) != cudaSuccess ||
                         cudaMalloc(&od, (size_t)n_signals * out_len * sizeof(double)) != cudaSuccess))
        rc = BSS_ENOMEM;
    if (rc == BSS_OK) {
        cudaMemcpy(zd, z, n_in * sizeof(double2), cudaMemcpyHostToDevice);
        IstftParams<double2> p{};
        p.z = zd;
        p.win = t.win;
        p.tw = t.tw;
        p.S = n_signals;
        p.N = fft_size;
        p.hop = hop_size;
        p.n_frames = n_frames;
        p.scale = (float)(t.win_sum / (double)(fft_size / 2));
        p.frames = frames;
        const size_t smem = fft_smem_bytes(fft_size);
        cudaFuncSetAttribute(istft_frames_kernel<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid(n_frames, n_signals);
        istft_frames_kernel<double2><<<grid, fft_threads(fft_size), smem>>>(p);
        const long long n = (long long)n_signals * out_len;
        istft_ola_kernel<double><<<(unsigned)cdiv(n, 256), 256>>>(frames, t.win, od, n_signals, fft_size, hop_size, n_frames, out_len);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(out, od, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = BSS_ECUDA;
    }
    if (zd) cudaFree(zd);
    if (frames) cudaFree(frames);
    if (od) cudaFree(od);
    free_tables(&t);
    return rc;
}

}  // extern "C"

// STFT tables of a handle: built once per (fft_size, window) and kept, so that the waveform feed of a job-after-job handle
// allocates nothing (cudaMalloc / cudaFree synchronise the whole device and would serialise the pipelined sub-batches)
static int handle_tables(bss_handle* h, const double* window, int N, FftTables* t) {
    if (h->fft_N != N || h->fft_window_host.size() != (size_t)N || memcmp(h->fft_window_host.data(), window, N * sizeof(double)) != 0) {
        FftTables fresh;
        const int rc = make_tables(window, N, h->stream, &fresh);
        if (rc != BSS_OK) {
            free_tables(&fresh);
            return rc;
        }
        if (h->fft_win) cudaFree(h->fft_win);
        if (h->fft_tw) cudaFree(h->fft_tw);
        h->fft_win = fresh.win;
        h->fft_tw = fresh.tw;
        h->fft_win_sum = fresh.win_sum;
        h->fft_N = N;
        h->fft_window_host.assign(window, window + N);
    }
    t->win = h->fft_win;
    t->tw = h->fft_tw;
    t->win_sum = h->fft_win_sum;
    return BSS_OK;
}

// waveform feed of a handle: x (B, C, n_samples) on the host -> h->X tiles, no spectrogram on the host.  Everything is queued
// on the handle's stream; the caller (finish_input) waits before the host buffer is handed back.
int stft_into_handle(bss_handle* h, const void* x, int dtype, int n_samples, int fft_size, int hop_size, const double* window) {
    if (!pow2(fft_size) || fft_size < 8 || fft_size > 16384) return bss_fail(h, BSS_EUNSUPPORTED, "fft_size must be a power of two in [8, 16384]");
    if (dtype != BSS_F32 && dtype != BSS_F64 && dtype != BSS_I16) return bss_fail(h, BSS_EINVAL, "waveforms are float32, float64 or int16 PCM");
    if (fft_size / 2 + 1 != h->F) return bss_fail(h, BSS_EINVAL, "n_bins of the handle must be fft_size / 2 + 1");
    const int n_frames = bss_stft_frames(n_samples, fft_size, hop_size);
    if (n_frames != h->T) return bss_fail(h, BSS_EINVAL, "n_frames of the handle does not match the waveform length");
    FftTables t;
    if (handle_tables(h, window, fft_size, &t) != BSS_OK) return bss_fail(h, BSS_ENOMEM, "stft tables");
    const int S = h->B * h->C;
    const long long n = (long long)S * n_samples;
    // staging: [float waveform | the waveform as it came (float64 or int16 input only)]
    BSS_TRY(ensure_staging(h, (size_t)n * sizeof(float) + (dtype == BSS_F64 ? (size_t)n * sizeof(double) : dtype == BSS_I16 ? (size_t)n * 2 : 0)));
    float* xf = (float*)h->staging;
    if (dtype == BSS_F32) {
        BSS_CUDA(h, cudaMemcpyAsync(xf, x, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    } else if (dtype == BSS_I16) {
        short* staged = (short*)((char*)h->staging + (size_t)n * sizeof(float));
        BSS_CUDA(h, cudaMemcpyAsync(staged, x, n * 2, cudaMemcpyHostToDevice, h->stream));
        pcm16_to_float_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(staged, xf, n);
        h->launches += 1;
    } else {
        double* staged = (double*)((char*)h->staging + (size_t)n * sizeof(float));
        BSS_CUDA(h, cudaMemcpyAsync(staged, x, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        to_float_kernel<<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(staged, xf, n);
        h->launches += 1;
    }
    // pad frames (T odd) must read as zero
    if (h->Tp != h->T) BSS_CUDA(h, cudaMemsetAsync(h->X, 0, (size_t)h->B * h->F * h->C * h->Tp * sizeof(cf), h->stream));
    StftParams p{};
    p.x = xf;
    p.win = t.win;
    p.tw = t.tw;
    p.S = S;
    p.n_samples = n_samples;
    p.N = fft_size;
    p.hop = hop_size;
    p.n_frames = n_frames;
    p.scale = (float)(1.0 / t.win_sum);
    p.out128 = nullptr;
    p.X = h->X;
    p.C = h->C;
    p.Tp = h->Tp;
    const size_t smem = fft_smem_bytes(fft_size);
    BSS_CUDA(h, cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(n_frames, S);
    stft_kernel<<<grid, fft_threads(fft_size), smem, h->stream>>>(p);
    h->launches += 1;
    BSS_CUDA(h, cudaGetLastError());
    return BSS_OK;
}

// time-domain output of a handle: ISTFT of the separated estimates already sitting on the device as (B, N, F, T) complex64;
// y (B, N, out_len) float32/float64 on the host.  Scratch lives on the handle (no allocation per call).
int istft_from_device(bss_handle* h, const cf* z, int n_signals, int fft_size, int hop_size, const double* window, void* y, int dtype,
                      int y_on_device) {
    if (!pow2(fft_size) || fft_size < 8 || fft_size > 16384) return bss_fail(h, BSS_EUNSUPPORTED, "fft_size must be a power of two in [8, 16384]");
    if (dtype != BSS_F32 && dtype != BSS_F64) return bss_fail(h, BSS_EINVAL, "waveforms are float32 or float64");
    if (fft_size / 2 + 1 != h->F) return bss_fail(h, BSS_EINVAL, "n_bins of the handle must be fft_size / 2 + 1");
    const int n_frames = h->T;
    const int out_len = bss_istft_length(n_frames, fft_size, hop_size);
    if (out_len < 1) return bss_fail(h, BSS_EINVAL, "invalid ISTFT geometry");
    FftTables t;
    if (handle_tables(h, window, fft_size, &t) != BSS_OK) return bss_fail(h, BSS_ENOMEM, "stft tables");
    const size_t esz = dtype == BSS_F32 ? 4 : 8;
    const size_t frames_bytes = ((size_t)n_signals * n_frames * fft_size * sizeof(float) + 255) / 256 * 256;
    BSS_TRY(ensure_scratch2(h, frames_bytes + (y_on_device ? 0 : (size_t)n_signals * out_len * esz)));
    float* frames = (float*)h->scratch2;
    void* od = y_on_device ? y : (void*)((char*)h->scratch2 + frames_bytes);   // overlap-add straight into the caller's device buffer
    IstftParams<float2> p{};
    p.z = z;
    p.win = t.win;
    p.tw = t.tw;
    p.S = n_signals;
    p.N = fft_size;
    p.hop = hop_size;
    p.n_frames = n_frames;
    p.scale = (float)(t.win_sum / (double)(fft_size / 2));
    p.frames = frames;
    const size_t smem = fft_smem_bytes(fft_size);
    BSS_CUDA(h, cudaFuncSetAttribute(istft_frames_kernel<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(n_frames, n_signals);
    istft_frames_kernel<float2><<<grid, fft_threads(fft_size), smem, h->stream>>>(p);
    const long long n = (long long)n_signals * out_len;
    if (dtype == BSS_F32)
        istft_ola_kernel<float><<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(frames, t.win, (float*)od, n_signals, fft_size, hop_size,
                                                                              n_frames, out_len);
    else
        istft_ola_kernel<double><<<(unsigned)cdiv(n, 256), 256, 0, h->stream>>>(frames, t.win, (double*)od, n_signals, fft_size,
                                                                               hop_size, n_frames, out_len);
    h->launches += 2;
    BSS_CUDA(h, cudaGetLastError());
    if (y_on_device) return BSS_OK;   // queued on the handle's stream; bss_synchronize when the caller needs it
    BSS_CUDA(h, cudaMemcpyAsync(y, od, (size_t)n * esz, cudaMemcpyDeviceToHost, h->stream));
    BSS_CUDA(h, bss_wait(h));
    return BSS_OK;
}
