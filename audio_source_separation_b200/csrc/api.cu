// C ABI of libbssgpu.so (see include/bssgpu.h): handle life cycle, host <-> device state
// movement and the orchestration of one update_once per method.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <dlfcn.h>

#include "handle.h"
#include "methods.h"

static thread_local std::string g_create_error;

// NCCL is resolved at run time: the library has no link-time dependency on it, and a process that already carries
// libnccl.so.2 (torch does) shares that copy.
namespace {
typedef int (*nccl_group_fn)(void);
typedef int (*nccl_bcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*nccl_err_fn)(int);
struct NcclApi {
    nccl_group_fn group_start = nullptr, group_end = nullptr;
    nccl_bcast_fn broadcast = nullptr;
    nccl_allgather_fn all_gather = nullptr;
    nccl_err_fn error_string = nullptr;
    bool ok = false;
};
const NcclApi& nccl_api() {
    static const NcclApi api = [] {
        NcclApi a;
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return a;
        a.group_start = (nccl_group_fn)dlsym(lib, "ncclGroupStart");
        a.group_end = (nccl_group_fn)dlsym(lib, "ncclGroupEnd");
        a.broadcast = (nccl_bcast_fn)dlsym(lib, "ncclBroadcast");
        a.all_gather = (nccl_allgather_fn)dlsym(lib, "ncclAllGather");
        a.error_string = (nccl_err_fn)dlsym(lib, "ncclGetErrorString");
        a.ok = a.group_start && a.group_end && a.broadcast && a.all_gather;
        return a;
    }();
    return api;
}
}  // namespace

namespace {

bool is_nmf(int m) { return m >= BSS_NMF_EUC && m <= BSS_NMF_CAUCHY; }

template <typename T>
int dev_alloc(bss_handle* h, T** p, size_t n) {
    if (n == 0) n = 1;
    BSS_CUDA(h, cudaMalloc((void**)p, n * sizeof(T)));
    BSS_CUDA(h, cudaMemsetAsync(*p, 0, n * sizeof(T), h->stream));
    return BSS_OK;
}

int validate(const bss_config* c, std::string* why) {
    auto bad = [&](const char* m) {
        *why = m;
        return BSS_EINVAL;
    };
    if (c->n_batch < 1) return bad("n_batch must be >= 1");
    if (c->n_bins < 1 || c->n_frames < 1) return bad("n_bins and n_frames must be >= 1");
    if (c->n_basis < 1 || c->n_basis > 64) return bad("n_basis must be between 1 and 64");
    if (is_nmf(c->method)) {
        if (!(c->domain >= 1.0 && c->domain <= 2.0)) return bad("1 <= `domain` <= 2 is not satisfied.");
        return BSS_OK;
    }
    if (c->n_channels < 2 || c->n_channels > 8) return bad("n_channels must be between 2 and 8");
    if (c->method == BSS_IS_MNMF) {
        if (c->n_channels > 4) return bad("IS-MNMF supports 2 to 4 channels");
        if (c->n_sources < 1 || c->n_sources > 8) return bad("n_sources must be between 1 and 8");
        if (c->reference_id < 0 || c->reference_id >= c->n_channels) return bad("reference_id out of range");
        return BSS_OK;
    }
    if (c->method == BSS_FAST_MNMF) {
        if (c->n_sources < 1 || c->n_sources > 8) return bad("n_sources must be between 1 and 8");
    } else {
        if (c->n_sources != c->n_channels) return bad("determined methods need n_sources == n_channels");
    }
    if (c->method == BSS_GAUSS_ILRMA && !(c->domain >= 1.0 && c->domain <= 2.0))
        return bad("1 <= `domain` <= 2 is not satisfied.");
    if (c->reference_id < 0 || c->reference_id >= c->n_channels) return bad("reference_id out of range");
    if (c->spatial < BSS_SPATIAL_IP || c->spatial > BSS_SPATIAL_IP2) return bad("unknown algorithm_spatial");
    if (c->spatial == BSS_SPATIAL_IP2 && c->n_sources < 2) return bad("IP2 needs at least two sources");
    switch (c->method) {
        case BSS_GAUSS_ILRMA:
        case BSS_T_ILRMA:
        case BSS_AUX_LAPLACE_IVA:
        case BSS_AUX_GAUSS_IVA:
        case BSS_FAST_MNMF: break;
        case BSS_GAUSS_IDLMA:
            if (c->spatial != BSS_SPATIAL_IP) return bad("GaussIDLMA updates the spatial model by IP only");
            if (!(c->domain >= 1.0 && c->domain <= 2.0)) return bad("1 <= `domain` <= 2 is not satisfied.");
            break;
        default: return bad("unknown method");
    }
    return BSS_OK;
}

// surface exactly singular bins (the reference raises LinAlgError there); requires a sync
int check_flags(bss_handle* h) {
    int32_t flag = 0;
    BSS_CUDA(h, cudaMemcpyAsync(&flag, h->flags, sizeof(flag), cudaMemcpyDeviceToHost, h->stream));
    BSS_CUDA(h, bss_wait(h));
    if (flag != 0) {
        cudaMemsetAsync(h->flags, 0, sizeof(int32_t), h->stream);
        return bss_fail(h, BSS_ESINGULAR, "Singular matrix");
    }
    return BSS_OK;
}

}  // namespace

extern "C" {

const char* bss_version(void) { return "bssgpu 0.1 (sm_100a)"; }

const char* bss_last_error(const bss_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t bss_launch_count(const bss_handle* h) { return h ? h->launches : 0; }

int bss_create(const bss_config* cfg, bss_handle** out) {
    if (!cfg || !out) {
        g_create_error = "null argument";
        return BSS_EINVAL;
    }
    *out = nullptr;
    std::string why;
    if (validate(cfg, &why) != BSS_OK) {
        g_create_error = why;
        return BSS_EINVAL;
    }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) +
                         " (libbssgpu has no CPU path)";
        return BSS_ECUDA;
    }
    if (cfg->device < 0 || cfg->device >= n_dev) {
        g_create_error = "device ordinal out of range";
        return BSS_EINVAL;
    }
    bss_handle* h = new (std::nothrow) bss_handle();
    if (!h) {
        g_create_error = "out of host memory";
        return BSS_ENOMEM;
    }
    h->cfg = *cfg;
    h->B = cfg->n_batch;
    h->C = cfg->n_channels;
    h->N = cfg->n_sources;
    h->F = cfg->n_bins;
    h->T = cfg->n_frames;
    h->Tp = round_up(cfg->n_frames, 2);
    h->K = cfg->n_basis;
    auto fail = [&](int rc) {
        g_create_error = h->err;
        bss_destroy(h);
        return rc;
    };
#define CREATE_CUDA(call)                                                   \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) {                                           \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e__);   \
            return fail(BSS_ECUDA);                                         \
        }                                                                   \
    } while (0)
    CREATE_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CREATE_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) {
        h->err = "libbssgpu is built for sm_100a (B200) only";
        return fail(BSS_ECUDA);
    }
    h->n_sm = prop.multiProcessorCount;
    h->max_smem = (int)prop.sharedMemPerBlockOptin;
    {
        int least = 0, greatest = 0;   // numerically: greatest priority <= least priority
        CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        int prio = cfg->stream_priority;
        if (prio < greatest) prio = greatest;
        if (prio > least) prio = least;
        CREATE_CUDA(cudaStreamCreateWithPriority(&h->own_stream, cudaStreamNonBlocking, prio));
    }
    h->stream = h->own_stream;
    CREATE_CUDA(cudaEventCreate(&h->ev0));
    CREATE_CUDA(cudaEventCreate(&h->ev1));
    int rc = dev_alloc(h, &h->flags, 4);
    if (rc == BSS_OK) {
        if (is_nmf(cfg->method))
            rc = nmf_allocate(h);
        else if (cfg->method == BSS_FAST_MNMF)
            rc = mnmf_allocate(h);
        else if (cfg->method == BSS_IS_MNMF)
            rc = smnmf_allocate(h);
        else
            rc = bss_allocate(h);
    }
    if (rc != BSS_OK) return fail(rc);
    CREATE_CUDA(bss_wait(h));
#undef CREATE_CUDA
    *out = h;
    return BSS_OK;
}

void bss_destroy(bss_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) bss_wait(h);
    void* bufs[] = {h->X,   h->Y,    h->W,     h->Wf,    h->basis, h->basis2, h->act,     h->latent, h->U,      h->Cx,
                    h->gate, h->flags, h->pw,   h->scale, h->wfr,   h->wraw,   h->order,   h->logdet, h->aux,    h->G2x,
                    h->part, h->iw, h->P, h->eigval, h->scratch2, h->fft_win, h->fft_tw, h->lossbuf, h->staging, h->G, h->target, h->xt, h->mn_acc, h->mn_acc2, h->latent2,
                    h->nz,   h->nt,   h->nv,    h->npart, h->loss_hist, h->G2, h->beff, h->aeff, h->praw,
                    h->sH,   h->sZ,   h->sT,    h->sV,    h->sStat, h->sPart, h->sAcc};
    for (void* p : bufs)
        if (p) cudaFree(p);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->graph_exec);
    if (h->graph_rec_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->graph_rec_exec);
    if (h->loss_counter) cudaFree(h->loss_counter);
    if (h->ev_block) cudaEventDestroy(h->ev_block);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int bss_set_stream(bss_handle* h, void* cuda_stream) {
    if (!h) return BSS_EINVAL;
    BSS_CUDA(h, bss_wait(h));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return BSS_OK;
}

int bss_synchronize(bss_handle* h) {
    if (!h) return BSS_EINVAL;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    return check_flags(h);
}

int bss_timer_begin(bss_handle* h) {
    if (!h) return BSS_EINVAL;
    BSS_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    return BSS_OK;
}

int bss_timer_end(bss_handle* h, float* elapsed_ms) {
    if (!h || !elapsed_ms) return BSS_EINVAL;
    BSS_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    BSS_CUDA(h, cudaEventSynchronize(h->ev1));
    BSS_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
    return BSS_OK;
}

static int finish_input(bss_handle* h);

int bss_set_input(bss_handle* h, const void* x, int dtype) {
    if (!h || !x) return BSS_EINVAL;
    if (is_nmf(h->cfg.method)) return bss_fail(h, BSS_EINVAL, "NMF takes its target through bss_set_state");
    if (dtype != BSS_C64 && dtype != BSS_C128) return bss_fail(h, BSS_EINVAL, "input must be complex64 or complex128");
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    const size_t elems = (size_t)h->B * h->C * h->F * h->T;
    const size_t bytes = elems * (dtype == BSS_C128 ? 16 : 8);
    BSS_TRY(ensure_staging(h, bytes));
    BSS_CUDA(h, cudaMemcpyAsync(h->staging, x, bytes, cudaMemcpyHostToDevice, h->stream));
    BSS_TRY(launch_import_x(h, h->staging, dtype, h->X, h->B, h->C, h->F, h->T, h->Tp));
    return finish_input(h);
}

int bss_set_input_waveform(bss_handle* h, const void* x, int dtype, int n_samples, int fft_size, int hop_size, const double* window) {
    if (!h || !x || !window) return BSS_EINVAL;
    if (is_nmf(h->cfg.method)) return bss_fail(h, BSS_EINVAL, "NMF takes its target through bss_set_state");
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    BSS_TRY(stft_into_handle(h, x, dtype, n_samples, fft_size, hop_size, window));
    return finish_input(h);
}

static int finish_input(bss_handle* h) {
    // plain covariance mean_t x x^H (algebraic power normalisation / projection back), accumulated in fp64
    if (h->cfg.method != BSS_IS_MNMF) BSS_TRY(launch_plain_covariance(h, h->X, h->Cx, h->B, h->F, h->C, h->T, h->Tp));
    h->has_input = true;
    h->y_valid = false;
    // ISS carries estimates instead of a filter: (re)derive them when the filter came first
    if (h->cfg.spatial == BSS_SPATIAL_ISS && h->cfg.method != BSS_FAST_MNMF && h->has_filter) BSS_TRY(bss_refresh_estimates(h));
    // the caller's buffer may be reused as soon as we return (BSS_OPT_ASYNC_INPUT: the caller keeps it until the next wait)
    if (!h->opt_async_input) BSS_CUDA(h, bss_wait(h));
    return BSS_OK;
}

int bss_reset_spatial(bss_handle* h) {
    if (!h) return BSS_EINVAL;
    if (is_nmf(h->cfg.method)) return BSS_OK;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (h->cfg.method == BSS_FAST_MNMF) return mnmf_reset(h);
    if (h->cfg.method == BSS_IS_MNMF) return smnmf_reset(h);
    return bss_reset_filter(h);
}

int bss_set_update_pair(bss_handle* h, int m, int n) {
    if (!h) return BSS_EINVAL;
    if (m == -1 && n == -1) {   // `update_pair = None`: the next bss_run starts the schedule over at (0, 1)
        h->pair_m = h->pair_n = -1;
        return BSS_OK;
    }
    if (m < 0 || n < 0 || m >= h->N || n >= h->N || m == n) return bss_fail(h, BSS_EINVAL, "invalid update pair");
    h->pair_m = m;
    h->pair_n = n;
    return BSS_OK;
}

int bss_update_once(bss_handle* h) {
    if (!h) return BSS_EINVAL;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (is_nmf(h->cfg.method)) return nmf_update_once(h);
    if (!h->has_input) return bss_fail(h, BSS_ESTATE, "Specify data!");
    switch (h->cfg.method) {
        case BSS_GAUSS_ILRMA: return ilrma_update_once(h);
        case BSS_T_ILRMA: return tilrma_update_once(h);
        case BSS_AUX_LAPLACE_IVA:
        case BSS_AUX_GAUSS_IVA: return auxiva_update_once(h);
        case BSS_GAUSS_IDLMA: return idlma_update_once(h);
        case BSS_FAST_MNMF: return mnmf_update_once(h);
        case BSS_IS_MNMF: return smnmf_update_once(h);
    }
    return bss_fail(h, BSS_EINVAL, "unknown method");
}

int bss_normalize(bss_handle* h) {
    if (!h) return BSS_EINVAL;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (h->cfg.method != BSS_GAUSS_IDLMA)
        return bss_fail(h, BSS_EUNSUPPORTED, "normalisation is part of update_once for this method");
    if (!h->has_input) return bss_fail(h, BSS_ESTATE, "Specify data!");
    return idlma_normalize(h);
}

// n eager iterations (advances the IP2 pair schedule like the reference's __call__, src/bss/ilrma.py:635-646)
static int run_eager(bss_handle* h, int n_iter) {
    for (int i = 0; i < n_iter; ++i) {
        if (!is_nmf(h->cfg.method) && h->cfg.spatial == BSS_SPATIAL_IP2 && h->cfg.method != BSS_FAST_MNMF) {
            if (h->pair_m < 0) {
                h->pair_m = 0;
                h->pair_n = 1;
            } else {
                h->pair_m = (h->pair_m + 1) % h->N;
                h->pair_n = (h->pair_n + 1) % h->N;
            }
        }
        BSS_TRY(bss_update_once(h));
    }
    return BSS_OK;
}

// Long loops are replayed from a CUDA graph: small problems (one mixture, NMF) are bound by launch latency, and a graph
// of two iterations (two, because the basis / spatial buffers are double buffered and swap every iteration) removes the
// host from the loop.  IP2 changes its kernel arguments every iteration (the update pair) and stays eager.
static bool graph_capable(const bss_handle* h) {
    if (getenv("BSSGPU_NO_GRAPH")) return false;
    if (!is_nmf(h->cfg.method) && h->cfg.method != BSS_FAST_MNMF && h->cfg.spatial == BSS_SPATIAL_IP2) return false;
    return true;
}

// FNV-1a over every device pointer (and scratch size) a captured update_once of the determined methods can reference, in
// their current roles (basis / basis2 swap every iteration).  0 = this configuration is re-captured on every call.
static uint64_t graph_signature(const bss_handle* h) {
    if (h->cfg.method > BSS_AUX_GAUSS_IVA || h->cfg.partitioning) return 0;
    const void* ptrs[] = {h->X, h->Y, h->W, h->Wf, h->basis, h->basis2, h->act, h->U, h->Cx, h->gate, h->flags, h->pw, h->scale,
                          h->wfr, h->wraw, h->order, h->logdet, h->aux, h->G2x, h->part, h->iw, h->P, h->eigval, h->scratch2, h->fft_win, h->fft_tw, h->lossbuf, h->staging};
    uint64_t s = 1469598103934665603ull;
    auto mix = [&](uint64_t v) {
        for (int i = 0; i < 8; ++i) {
            s ^= (v >> (8 * i)) & 0xffu;
            s *= 1099511628211ull;
        }
    };
    for (const void* q : ptrs) mix((uint64_t)(uintptr_t)q);
    mix((uint64_t)h->part_elems);
    mix((uint64_t)h->staging_bytes);
    mix((uint64_t)(uintptr_t)h->stream);
    return s ? s : 1;
}

// one loss evaluation queued on the stream; result at lossbuf[B*F .. B*F+B)
static int loss_device(bss_handle* h) {
    if (is_nmf(h->cfg.method)) return nmf_loss(h);
    if (!h->has_input) return bss_fail(h, BSS_ESTATE, "Specify data!");
    if (h->cfg.method == BSS_IS_MNMF) return smnmf_loss(h);
    return h->cfg.method == BSS_FAST_MNMF ? mnmf_loss(h) : bss_loss_device(h);
}

// n iterations, each optionally followed by the loss and its append to the device-side history
static int run_eager_rec(bss_handle* h, int n_iter, bool record) {
    if (!record) return run_eager(h, n_iter);
    for (int i = 0; i < n_iter; ++i) {
        BSS_TRY(run_eager(h, 1));
        BSS_TRY(loss_device(h));
        BSS_TRY(launch_loss_append(h, h->lossbuf + (size_t)h->B * h->F, h->loss_hist, h->loss_counter, h->B,
                                   (int)(h->loss_hist_elems / (size_t)h->B)));
    }
    return BSS_OK;
}

static int run_loop(bss_handle* h, int n_iter, bool record) {
    const int kPerGraph = 2, kWarm = 2, kMinReplays = 4;
    if (n_iter < kWarm + kPerGraph * kMinReplays || !graph_capable(h)) return run_eager_rec(h, n_iter, record);
    // eager warm-up: every scratch buffer reaches its final size and every kernel attribute is set before the capture
    BSS_TRY(run_eager_rec(h, kWarm, record));
    n_iter -= kWarm;
    // A handle that is used for job after job (batch.py keeps its handles) replays the graph it already has as long as
    // the captured kernel arguments still describe the current state: instantiating and destroying an executable graph
    // per call costs host time and synchronises with the other handles' streams, which serialised the pipelined
    // end-to-end job (profiles/r5b_e2e_timeline.txt).
    void*& slot_exec = record ? h->graph_rec_exec : h->graph_exec;
    uint64_t& slot_sig = record ? h->graph_rec_sig : h->graph_sig;
    int64_t& slot_launches = record ? h->graph_rec_launches : h->graph_launches;
    const uint64_t sig = graph_signature(h);
    cudaGraphExec_t exec = nullptr;
    int64_t per_graph = 0;
    if (slot_exec && sig != 0 && sig == slot_sig) {
        exec = (cudaGraphExec_t)slot_exec;
        per_graph = slot_launches;
    } else {
        if (slot_exec) {
            cudaGraphExecDestroy((cudaGraphExec_t)slot_exec);
            slot_exec = nullptr;
            slot_sig = 0;
        }
        cudaGraph_t graph = nullptr;
        const int64_t l0 = h->launches;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            return run_eager_rec(h, n_iter, record);
        }
        const int rc_cap = run_eager_rec(h, kPerGraph, record);
        const cudaError_t e_end = cudaStreamEndCapture(h->stream, &graph);
        per_graph = h->launches - l0;
        h->launches = l0;   // nothing of the capture has executed
        if (rc_cap != BSS_OK || e_end != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            if (rc_cap != BSS_OK) return rc_cap;
            return run_eager_rec(h, n_iter, record);
        }
        cudaGraphDestroy(graph);
        slot_exec = exec;
        slot_launches = per_graph;
        // the capture itself may have grown a scratch buffer: sign the state the kernels were actually recorded with
        slot_sig = graph_signature(h) == sig ? sig : 0;
    }
    const int reps = n_iter / kPerGraph;
    for (int r = 0; r < reps; ++r) BSS_CUDA(h, cudaGraphLaunch(exec, h->stream));
    h->launches += per_graph * reps;
    h->graph_replays += reps;
    return run_eager_rec(h, n_iter - reps * kPerGraph, record);
}

int bss_run(bss_handle* h, int n_iter) {
    if (!h || n_iter < 0) return BSS_EINVAL;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (is_nmf(h->cfg.method)) return nmf_run(h, n_iter, nullptr);   // one cluster launch for the whole loop when it fits
    return run_loop(h, n_iter, false);
}

int bss_run_record(bss_handle* h, int n_iter, double* loss) {
    if (!h || n_iter < 0 || !loss) return BSS_EINVAL;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    const size_t need = (size_t)(n_iter > 0 ? n_iter : 1) * h->B;
    if (need > h->loss_hist_elems) {
        if (h->loss_hist) cudaFree(h->loss_hist);
        h->loss_hist = nullptr;
        h->loss_hist_elems = 0;
        BSS_CUDA(h, cudaMalloc(&h->loss_hist, need * sizeof(double)));
        h->loss_hist_elems = need;
    }
    if (is_nmf(h->cfg.method)) {
        BSS_TRY(nmf_run(h, n_iter, h->loss_hist));
        if (n_iter > 0)
            BSS_CUDA(h, cudaMemcpyAsync(loss, h->loss_hist, sizeof(double) * h->B * n_iter, cudaMemcpyDeviceToHost, h->stream));
        return check_flags(h);
    }
    // every iteration is followed by its loss, appended to the history at a device-side counter: the (update, loss, append)
    // triples replay from a CUDA graph like the plain loop, so the default recordable_loss=True path of the reference
    // (src/bss/ilrma.py:239-241) costs the loss kernels and nothing else
    if (!h->loss_counter) BSS_CUDA(h, cudaMalloc((void**)&h->loss_counter, sizeof(int)));
    BSS_CUDA(h, cudaMemsetAsync(h->loss_counter, 0, sizeof(int), h->stream));
    BSS_TRY(run_loop(h, n_iter, true));
    if (n_iter > 0)
        BSS_CUDA(h, cudaMemcpyAsync(loss, h->loss_hist, sizeof(double) * h->B * n_iter, cudaMemcpyDeviceToHost, h->stream));
    return check_flags(h);
}

int bss_loss(bss_handle* h, double* loss) {
    if (!h || !loss) return BSS_EINVAL;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    int rc;
    if (is_nmf(h->cfg.method))
        rc = nmf_loss(h);
    else {
        if (!h->has_input) return bss_fail(h, BSS_ESTATE, "Specify data!");
        rc = loss_device(h);
    }
    if (rc != BSS_OK) return rc;
    // results sit at lossbuf[B*F .. B*F+B)
    BSS_CUDA(h, cudaMemcpyAsync(loss, h->lossbuf + (size_t)h->B * h->F, sizeof(double) * h->B, cudaMemcpyDeviceToHost,
                                h->stream));
    return check_flags(h);
}

int bss_separate_device(bss_handle* h, void* y_device, int apply_projection_back) {
    if (!h || !y_device) return BSS_EINVAL;
    if (is_nmf(h->cfg.method)) return bss_fail(h, BSS_EINVAL, "NMF has no separate()");
    if (!h->has_input) return bss_fail(h, BSS_ESTATE, "Specify data!");
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (h->cfg.method == BSS_FAST_MNMF) return mnmf_separate(h, (cf*)y_device);
    if (h->cfg.method == BSS_IS_MNMF) return smnmf_separate(h, (cf*)y_device);
    return bss_separate_to(h, (cf*)y_device, apply_projection_back);
}

int bss_separate(bss_handle* h, void* y, int dtype, int apply_projection_back) {
    if (!h || !y) return BSS_EINVAL;
    if (dtype != BSS_C64 && dtype != BSS_C128) return bss_fail(h, BSS_EINVAL, "output must be complex64 or complex128");
    const size_t elems = (size_t)h->B * h->N * h->F * h->T;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (dtype == BSS_C64) {
        BSS_TRY(ensure_staging(h, elems * 8));
        BSS_TRY(bss_separate_device(h, h->staging, apply_projection_back));
        BSS_CUDA(h, cudaMemcpyAsync(y, h->staging, elems * 8, cudaMemcpyDeviceToHost, h->stream));
        return check_flags(h);
    }
    // complex128 for the caller: widen on the device and copy straight into the caller's array (no pinned bounce buffer,
    // no host-side conversion loop)
    BSS_TRY(ensure_staging(h, elems * 24));
    cf* narrow = (cf*)((char*)h->staging + elems * 16);
    BSS_TRY(bss_separate_device(h, narrow, apply_projection_back));
    BSS_TRY(launch_widen(h, narrow, (double2*)h->staging, (long long)elems));
    BSS_CUDA(h, cudaMemcpyAsync(y, h->staging, elems * 16, cudaMemcpyDeviceToHost, h->stream));
    return check_flags(h);
}

int bss_separate_waveform(bss_handle* h, void* y, int dtype, int fft_size, int hop_size, const double* window,
                          int apply_projection_back) {
    if (!h || !y || !window) return BSS_EINVAL;
    const size_t elems = (size_t)h->B * h->N * h->F * h->T;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    BSS_TRY(ensure_staging(h, elems * 8));
    BSS_TRY(bss_separate_device(h, h->staging, apply_projection_back));
    BSS_TRY(istft_from_device(h, (const cf*)h->staging, h->B * h->N, fft_size, hop_size, window, y, dtype));
    return check_flags(h);
}

int bss_separate_waveform_device(bss_handle* h, void* y_device, int dtype, int fft_size, int hop_size, const double* window,
                                 int apply_projection_back) {
    if (!h || !y_device || !window) return BSS_EINVAL;
    const size_t elems = (size_t)h->B * h->N * h->F * h->T;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    BSS_TRY(ensure_staging(h, elems * 8));
    BSS_TRY(bss_separate_device(h, h->staging, apply_projection_back));
    return istft_from_device(h, (const cf*)h->staging, h->B * h->N, fft_size, hop_size, window, y_device, dtype, 1);
}

int bss_gather_outputs(bss_handle* h, void* nccl_comm, int n_ranks, int rank, const void* send_device, void* recv_base_device,
                       size_t bytes, size_t rank_stride_bytes) {
    if (!h || !nccl_comm || !send_device || !recv_base_device || n_ranks < 1 || rank < 0 || rank >= n_ranks) return BSS_EINVAL;
    const NcclApi& nccl = nccl_api();
    if (!nccl.ok) return bss_fail(h, BSS_ENCCL, "libnccl.so.2 not found");
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (bytes == 0) return BSS_OK;
    auto check = [&](int rc, const char* what) -> int {
        if (rc == 0) return BSS_OK;
        return bss_fail(h, BSS_ENCCL, std::string(what) + ": " + (nccl.error_string ? nccl.error_string(rc) : "NCCL error"));
    };
    // One NCCL group of broadcasts, each rank's contribution straight into its final place (no staging copy).
    // BSSGPU_GATHER=allgather selects the other form for A/B measurements: one ncclAllGather into a rank-major scratch buffer
    // plus one strided device copy -- measured slower inside the 8-GPU job (64.0 against 60.6 ms, profiles/round2_scaling.md).
    static const bool by_allgather = getenv("BSSGPU_GATHER") && !strcmp(getenv("BSSGPU_GATHER"), "allgather");
    if (!by_allgather || n_ranks == 1) {
        BSS_TRY(check(nccl.group_start(), "ncclGroupStart"));
        int rc = 0;
        for (int r = 0; r < n_ranks && rc == 0; ++r)
            rc = nccl.broadcast(send_device, (char*)recv_base_device + (size_t)r * rank_stride_bytes, bytes, /*ncclInt8*/ 0, r, nccl_comm, h->stream);
        const int rc_end = nccl.group_end();
        BSS_TRY(check(rc, "ncclBroadcast"));
        return check(rc_end, "ncclGroupEnd");
    }
    if (rank_stride_bytes == bytes) return check(nccl.all_gather(send_device, recv_base_device, bytes, 0, nccl_comm, h->stream), "ncclAllGather");
    BSS_TRY(ensure_scratch2(h, (size_t)n_ranks * bytes));
    BSS_TRY(check(nccl.all_gather(send_device, h->scratch2, bytes, 0, nccl_comm, h->stream), "ncclAllGather"));
    BSS_CUDA(h, cudaMemcpy2DAsync(recv_base_device, rank_stride_bytes, h->scratch2, bytes, bytes, (size_t)n_ranks, cudaMemcpyDeviceToDevice,
                                  h->stream));
    return BSS_OK;
}

int bss_peer_alloc(int device, size_t bytes, void** dptr, void* ipc_handle_64) {
    if (!dptr || !ipc_handle_64 || bytes == 0) return BSS_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    if (cudaSetDevice(device) != cudaSuccess) return BSS_ECUDA;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return BSS_ENOMEM;
    }
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, p) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
        return BSS_ECUDA;
    }
    memcpy(ipc_handle_64, &hd, 64);
    *dptr = p;
    return BSS_OK;
}

int bss_peer_open(int device, const void* ipc_handle_64, void** dptr) {
    if (!dptr || !ipc_handle_64) return BSS_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return BSS_ECUDA;
    cudaIpcMemHandle_t hd;
    memcpy(&hd, ipc_handle_64, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return BSS_ECUDA;
    }
    *dptr = p;
    return BSS_OK;
}

int bss_peer_close(int device, void* dptr) {
    if (!dptr) return BSS_OK;
    if (cudaSetDevice(device) != cudaSuccess) return BSS_ECUDA;
    return cudaIpcCloseMemHandle(dptr) == cudaSuccess ? BSS_OK : BSS_ECUDA;
}

int bss_peer_free(int device, void* dptr) {
    if (!dptr) return BSS_OK;
    if (cudaSetDevice(device) != cudaSuccess) return BSS_ECUDA;
    return cudaFree(dptr) == cudaSuccess ? BSS_OK : BSS_ECUDA;
}

int bss_push_outputs(bss_handle* h, int n_ranks, int rank, void* const* peer_bases, const void* send_device, size_t dst_offset_bytes,
                     size_t bytes) {
    if (!h || !peer_bases || !send_device || n_ranks < 1 || rank < 0 || rank >= n_ranks) return BSS_EINVAL;
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    if (bytes == 0) return BSS_OK;
    // start with the next rank: at any moment the ranks write to different peers
    for (int i = 1; i < n_ranks; ++i) {
        const int r = (rank + i) % n_ranks;
        if (!peer_bases[r]) return bss_fail(h, BSS_EINVAL, "push: a peer buffer is not mapped");
        BSS_CUDA(h, cudaMemcpyAsync((char*)peer_bases[r] + dst_offset_bytes, send_device, bytes, cudaMemcpyDefault, h->stream));
    }
    return BSS_OK;
}

int bss_set_option(bss_handle* h, int option, int value) {
    if (!h) return BSS_EINVAL;
    switch (option) {
        case BSS_OPT_IP_KERNEL:
            if (value < 0 || value > 2) return bss_fail(h, BSS_EINVAL, "BSS_OPT_IP_KERNEL takes 0 (auto), 1 or 2");
            if (value != h->opt_ip_kernel) h->graph_sig = h->graph_rec_sig = 0;   // a kept graph recorded the other kernel
            h->opt_ip_kernel = value;
            return BSS_OK;
        case BSS_OPT_SOURCE_MODEL:
            if (value < 0 || value > 2) return bss_fail(h, BSS_EINVAL, "BSS_OPT_SOURCE_MODEL takes 0 (auto), 1 or 2");
            if (value != h->opt_source_model) h->graph_sig = h->graph_rec_sig = 0;
            h->opt_source_model = value;
            return BSS_OK;
        case BSS_OPT_BLOCKING_SYNC:
            h->opt_blocking_sync = value != 0;
            return BSS_OK;
        case BSS_OPT_ASYNC_INPUT:
            h->opt_async_input = value != 0;
            return BSS_OK;
        case BSS_OPT_ACT_CHUNKS:
            if (value < 0) return bss_fail(h, BSS_EINVAL, "BSS_OPT_ACT_CHUNKS takes 0 (auto) or a positive count");
            if (value != h->opt_act_chunks) h->graph_sig = h->graph_rec_sig = 0;
            h->opt_act_chunks = value;
            return BSS_OK;
    }
    return bss_fail(h, BSS_EINVAL, "unknown option");
}

int bss_get_info(bss_handle* h, int what, int64_t* value) {
    if (!h || !value) return BSS_EINVAL;
    switch (what) {
        case BSS_INFO_IP_KERNEL: *value = h->last_ip_kernel; return BSS_OK;
        case BSS_INFO_GRAPH_REPLAYS: *value = h->graph_replays; return BSS_OK;
        case BSS_INFO_LAUNCHES: *value = h->launches; return BSS_OK;
        case BSS_INFO_ACT_CHUNKS: *value = h->last_act_chunks; return BSS_OK;
        case BSS_INFO_SOURCE_MODEL: *value = h->last_source_model; return BSS_OK;
    }
    return bss_fail(h, BSS_EINVAL, "unknown info");
}

int bss_compute_demix_filter(bss_handle* h) {
    if (!h) return BSS_EINVAL;
    if (is_nmf(h->cfg.method) || h->cfg.method == BSS_FAST_MNMF || h->cfg.method == BSS_IS_MNMF)
        return bss_fail(h, BSS_EINVAL, "no demixing filter");
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    return bss_filter_from_estimates(h);
}

int bss_device_buffer(bss_handle* h, int which, void** dptr, size_t* bytes) {
    if (!h || !dptr) return BSS_EINVAL;
    size_t n = 0;
    void* p = nullptr;
    switch (which) {
        case BSS_STATE_DEMIX_FILTER:
        case BSS_STATE_DIAGONALIZER:
            p = h->W;
            n = (size_t)h->B * h->F * h->C * h->C * 16;
            break;
        case BSS_STATE_ESTIMATION:
            p = h->Y;
            n = (size_t)h->B * h->F * h->N * h->Tp * 8;
            break;
        case BSS_STATE_BASIS: p = h->basis; break;
        case BSS_STATE_ACTIVATION: p = h->act; break;
        case BSS_STATE_COVARIANCE:
            p = h->U;
            n = (size_t)h->B * h->C * h->F * h->C * h->C * 8;
            break;
        default: return bss_fail(h, BSS_EINVAL, "no such device buffer");
    }
    *dptr = p;
    if (bytes) *bytes = n;
    return BSS_OK;
}

int bss_time_covariance(bss_handle* h, int repeat, float* mean_ms) {
    if (!h || !mean_ms || repeat < 1) return BSS_EINVAL;
    if (is_nmf(h->cfg.method) || h->cfg.method == BSS_IS_MNMF || !h->has_input) return bss_fail(h, BSS_ESTATE, "Specify data!");
    BSS_CUDA(h, cudaSetDevice(h->cfg.device));
    BSS_TRY(bss_covariance_only(h));   // warm-up
    BSS_CUDA(h, cudaEventRecord(h->ev0, h->stream));
    for (int i = 0; i < repeat; ++i) BSS_TRY(bss_covariance_only(h));
    BSS_CUDA(h, cudaEventRecord(h->ev1, h->stream));
    BSS_CUDA(h, cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    BSS_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    *mean_ms = ms / (float)repeat;
    return BSS_OK;
}

}  // extern "C"
