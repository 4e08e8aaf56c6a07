// Shared device helpers: complex arithmetic, mbarrier + bulk-copy (TMA, UBLKCP) wrappers,
// warp reductions and the per-warp streaming ring used by the bin-tile kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define BSS_WARP 32
#define BSS_FULL 0xffffffffu

// ------------------------------------------------------------------------------ complex (fp32)
typedef float2 cf;

__device__ __forceinline__ cf cf_make(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float cf_abs2(cf a) { return fmaf(a.x, a.x, a.y * a.y); }
// acc += a * b
__device__ __forceinline__ void cf_fma(cf& acc, cf a, cf b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}

__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------------------ mbarrier / bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t a = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
// global -> shared bulk copy (TMA engine, SASS UBLKCP); bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global bulk copy
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------ warp reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BSS_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(BSS_FULL, v, o);
    return v;
}

// Butterfly reduce-scatter of MP (= 32*Q) per-lane values: afterwards lane L holds in v[0..Q) the
// warp-wide sums of elements Q*L .. Q*L+Q-1.  31*Q shuffles instead of 5*MP.
template <int MP>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[MP], int lane) {
    static_assert(MP % 32 == 0, "pad to a multiple of 32");
#pragma unroll
    for (int lvl = 0; lvl < 5; ++lvl) {
        const int off = 16 >> lvl;
        const int cnt = (MP / 2) >> lvl;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt; ++i) {
            const float lo = v[i], hi = v[i + cnt];
            const float send = up ? lo : hi;
            const float keep = up ? hi : lo;
            v[i] = keep + __shfl_xor_sync(BSS_FULL, send, off);
        }
    }
}

// ------------------------------------------------------------------------------ bin-tile layout
// A bin tile [rows][Tp] (X: rows = channels, Y: rows = sources) is stored as consecutive blocks of
// BSS_XSLAB frames: block s holds `rows` rows of L_s = min(BSS_XSLAB, Tp - s BSS_XSLAB) frames each.  Every
// block (one ring stage) and the whole tile are contiguous in HBM, so a stage is filled by ONE bulk
// copy and a warp that walks a bin reads 8 rows Tp bytes strictly sequentially.  Tp <= BSS_XSLAB
// degenerates to the plain row-major tile.
#define BSS_XSLAB 128
__host__ __device__ __forceinline__ size_t tile_off(int rows, int Tp, int r, int t) {
    const int t0 = t & ~(BSS_XSLAB - 1);
    const int rest = Tp - t0;
    const int L = rest < BSS_XSLAB ? rest : BSS_XSLAB;
    return (size_t)t0 * rows + (size_t)r * L + (t - t0);
}

// ------------------------------------------------------------------------------ streaming ring
// A warp walks a list of jobs (bin tile, frame slab); lane 0 keeps STAGES-1 bulk copies in
// flight into the warp's private shared-memory ring while the whole warp consumes the oldest.
struct TileGeom {
    int n_rows;        // rows per bin tile (channels or sources)
    int row_len;       // padded frames per row (Tp)
    int slab;          // frames per block (BSS_XSLAB, or Tp when the tile is a single block)
    int n_slabs;       // blocks per tile
    int row_stride;    // unused by the kernels (rows of a staged block are frames() apart)
    uint32_t stage_bytes;
};

struct JobCursor {
    int item;          // flat item index (32 bit: the per-slab control path must stay cheap)
    int slab;
    __device__ __forceinline__ void advance(int stride, int n_slabs) {
        if (++slab == n_slabs) {
            slab = 0;
            item += stride;
        }
    }
};

__device__ __forceinline__ int slab_frames(const TileGeom& g, int slab) {
    const int rest = g.row_len - slab * g.slab;
    return rest < g.slab ? rest : g.slab;
}

// issue the copy of one job (one block of the tile, contiguous); `tile` points at the bin tile in global memory
__device__ __forceinline__ void ring_issue(const TileGeom& g, const cf* tile, int slab, unsigned char* stage,
                                           uint64_t* bar) {
    const uint32_t bytes = (uint32_t)g.n_rows * (uint32_t)slab_frames(g, slab) * 8u;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(stage, tile + (size_t)slab * g.slab * g.n_rows, bytes, bar);
}

// Per-warp stream over (item, slab) jobs.  Items owned by a warp are first, first+stride, ... < n_items;
// the bin tile of an item is tile index item / items_per_tile.  Geometry, base pointer and items_per_tile are
// passed to every call (they live in the kernel's constant parameter bank) so that the per-warp state is a
// dozen 32-bit registers: the streaming kernels are register bound.
template <int STAGES>
struct WarpStream {
    JobCursor prod, cons;
    int n_items, stride;
    uint32_t bars_sa;            // shared-space address of this warp's STAGES mbarriers
    uint32_t ring_sa;            // shared-space address of this warp's ring
    const unsigned char* ring;   // the same ring as a generic pointer (consumer loads)
    int pstage, cstage, lane;
    uint32_t cphase;

    __device__ __forceinline__ void issue_next(const TileGeom& g, const cf* base, int items_per_tile) {
        if (prod.item < n_items) {
            if (lane == 0) {
                const uint32_t tile = items_per_tile == 1 ? (uint32_t)prod.item : (uint32_t)prod.item / (uint32_t)items_per_tile;
                const uint32_t bytes = (uint32_t)g.n_rows * (uint32_t)slab_frames(g, prod.slab) * 8u;
                const uint32_t bar = bars_sa + 8u * (uint32_t)pstage;
                const cf* src = base + (size_t)tile * ((uint32_t)g.n_rows * (uint32_t)g.row_len) +
                                (size_t)((uint32_t)prod.slab * (uint32_t)g.slab * (uint32_t)g.n_rows);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 ring_sa + (uint32_t)pstage * g.stage_bytes),
                             "l"(src), "r"(bytes), "r"(bar)
                             : "memory");
            }
            prod.advance(stride, g.n_slabs);
            pstage = (pstage + 1 == STAGES) ? 0 : pstage + 1;
        }
    }
    __device__ __forceinline__ void start(const TileGeom& g, uint64_t* bars_, unsigned char* ring_, const cf* base, int first,
                                          int stride_, int n_items_, int items_per_tile, int lane_) {
        bars_sa = smem_u32(bars_);
        ring_sa = smem_u32(ring_);
        ring = ring_;
        n_items = n_items_;
        stride = stride_;
        lane = lane_;
        prod.item = cons.item = first;
        prod.slab = cons.slab = 0;
        pstage = cstage = 0;
        cphase = 0;
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) mbar_init(&bars_[s], 1);
            mbar_fence_init();
        }
        __syncwarp();
#pragma unroll 1
        for (int s = 0; s < STAGES - 1; ++s) issue_next(g, base, items_per_tile);
    }
    __device__ __forceinline__ bool active() const { return cons.item < n_items; }
    // wait for the current job's block; returns its shared-memory address
    __device__ __forceinline__ const cf* acquire(const TileGeom& g) {
        const uint32_t bar = bars_sa + 8u * (uint32_t)cstage;
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(cphase)
                : "memory");
        }
        return reinterpret_cast<const cf*>(ring + (size_t)((uint32_t)cstage * g.stage_bytes));
    }
    __device__ __forceinline__ int frames(const TileGeom& g) const { return slab_frames(g, cons.slab); }
    __device__ __forceinline__ int frame0(const TileGeom& g) const { return cons.slab * g.slab; }
    __device__ __forceinline__ bool first_slab() const { return cons.slab == 0; }
    __device__ __forceinline__ bool last_slab(const TileGeom& g) const { return cons.slab == g.n_slabs - 1; }
    // all lanes are done reading the current stage
    __device__ __forceinline__ void release(const TileGeom& g) {
        __syncwarp();
        cons.advance(stride, g.n_slabs);
        if (++cstage == STAGES) {
            cstage = 0;
            cphase ^= 1u;
        }
    }
};

// A CTA walks a contiguous range of items (its warps interleave inside it): consecutive bins of one
// mixture, so whatever is shared by the bins of a mixture can be cached per CTA.
__device__ __forceinline__ void cta_item_range(int n_items, int& lo, int& hi) {
    const int per_cta = (n_items + (int)gridDim.x - 1) / (int)gridDim.x;
    lo = (int)blockIdx.x * per_cta;
    hi = min(n_items, lo + per_cta);
}
// Copy the activation rows (floats_per_mix = N K Tp floats per mixture, contiguous over mixtures) of the
// mixtures touched by items [lo, hi) into shared memory; returns the first mixture.  The host guarantees
// (plan_stream_cached) that the range spans at most two mixtures.  Ends with a CTA barrier.
__device__ __forceinline__ int load_act_cache(float* cache, const float* act, int floats_per_mix, int lo, int hi, int items_per_bin,
                                              int n_bins) {
    int b_lo = 0;
    if (lo < hi) {
        b_lo = (lo / items_per_bin) / n_bins;
        const int b_hi = ((hi - 1) / items_per_bin) / n_bins;
        const int n2 = (b_hi - b_lo + 1) * (floats_per_mix >> 1);   // float2 elements (Tp is even)
        const float2* src = reinterpret_cast<const float2*>(act + (size_t)b_lo * floats_per_mix);
        float2* dst = reinterpret_cast<float2*>(cache);
        for (int i = threadIdx.x; i < n2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    return b_lo;
}

// host-side geometry of a [rows][Tp] bin tile: one block when Tp <= BSS_XSLAB, BSS_XSLAB-frame blocks otherwise
static inline TileGeom make_tile_geom(int rows, int Tp, int slab_frames = BSS_XSLAB) {
    slab_frames = BSS_XSLAB;   // fixed by the memory layout
    TileGeom g;
    g.n_rows = rows;
    g.row_len = Tp;
    if (Tp <= slab_frames) {
        g.slab = Tp;
        g.n_slabs = 1;
        g.row_stride = Tp;
    } else {
        g.slab = slab_frames;
        g.n_slabs = (Tp + slab_frames - 1) / slab_frames;
        g.row_stride = slab_frames;
    }
    g.stage_bytes = (uint32_t)(((size_t)rows * g.row_stride * 8 + 127) / 128 * 128);
    return g;
}
