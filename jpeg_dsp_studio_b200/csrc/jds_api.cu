// jds_api.cu - the C ABI of libjds.so (include/jds.h): context, buffers, staging of
// host data, chunking of batches / sweeps, and the launch sequence of the kernels.
// No CPU compute path exists here: every entry point that produces pixels or metrics
// launches the kernels of jds_kernels.cu and fails when no CUDA device is usable.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/jds.h"
#include "jds_host.h"
#include "jds_kernels.cuh"

using namespace jds;

// ------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define JDS_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return fail(JDS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                    \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                    \
    } while (0)

// ------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct jds_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    static constexpr int kTimedChunks = 32;             // chunks per call with per-stage events
    cudaEvent_t evs[5 * kTimedChunks] = {};
    cudaStream_t s_in = nullptr, s_out = nullptr;      // copy streams of the pipelined host path
    cudaStream_t stream2 = nullptr;                    // second compute stream (L2-sized sequences)
    cudaEvent_t ev_setup = nullptr, ev_join = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr},
                ev_out[2] = {nullptr, nullptr};
    int plan_chunk = 1;
    double stage_ms[4] = {0, 0, 0, 0};   // forward, codec, inverse, ssim (accumulated)
    uint64_t stage_launches[4] = {0, 0, 0, 0};
    DevBuf planes, in, recon, coeffs, errs, metrics, tables, selected, payload, alias, fcoef;
    DevBuf ent_sizes, ent_bits, ent_out, band;   // entropy bitstream: size passes, unstuffed bits, stuffed bytes
    void* h_metrics = nullptr;   // pinned
    size_t h_metrics_bytes = 0;
    void* h_tables = nullptr;    // pinned
    size_t h_tables_bytes = 0;
    void* h_selected = nullptr;  // pinned
    uint64_t launches = 0;
    int sm_count = 148;
    size_t l2_bytes = (size_t)96 << 20;
    bool legacy_ssim = false;     // JDS_LEGACY_SSIM=1: use the tile kernel (debug / A-B runs)
    bool no_fused = false;        // JDS_NO_FUSED=1: fast mode through the staged kernels
    bool no_hoist = false;        // JDS_NO_HOIST=1: sweeps redo the forward half per point (A/B runs)
    int l2_chunking = 0;          // JDS_L2_CHUNK=k: launch sequences of k x the frames whose working set fits L2 (0 = off)
    bool stage_timing = false;    // per-kernel CUDA events (jds_ctx_stage_timing)
    // jds_sweep_records returns without synchronising: its table uploads are staged in a ring
    // of pinned slots, each guarded by an event, so back-to-back calls never wait on the stream
    static constexpr int kTableRing = 4;
    void* h_tables_ring[kTableRing] = {};
    size_t h_tables_ring_bytes[kTableRing] = {};
    cudaEvent_t ev_tables[kTableRing] = {};
    bool ev_tables_used[kTableRing] = {};
    int tables_slot = 0;
    size_t scratch_budget = (size_t)1 << 30;
    int pipe_chunk = 0;
    bool no_dual_sweep = false;
    // a jds_roundtrip_batch_begin whose results jds_ctx_finish has not collected yet
    struct Pending {
        bool active = false;
        bool pipelined = false;
        int units = 0;
        jds_metrics* metrics = nullptr;
        uint64_t ssim_count = 0, ncoef = 0, luma_blocks = 0;
    } pend;
};

static int ensure(jds_ctx* c, DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes) return JDS_OK;
    if (b.p) {
        JDS_CUDA(cudaStreamSynchronize(c->stream));
        JDS_CUDA(cudaFree(b.p));
        b.p = nullptr;
        b.bytes = 0;
    }
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.p, bytes);
        want = bytes;
    }
    if (e != cudaSuccess)
        return fail(JDS_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    b.bytes = want;
    return JDS_OK;
}

static int ensure_pinned(void** p, size_t* have, size_t bytes) {
    if (*have >= bytes) return JDS_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *have = 0;
    JDS_CUDA(cudaMallocHost(p, bytes));
    *have = bytes;
    return JDS_OK;
}

extern "C" int jds_abi_version(void) { return JDS_ABI_VERSION; }

extern "C" const char* jds_last_error(void) { return g_last_error.c_str(); }

extern "C" int jds_device_count(int* count) {
    if (!count) return fail(JDS_ERR_INVALID, "count is NULL");
    *count = 0;
    JDS_CUDA(cudaGetDeviceCount(count));
    return JDS_OK;
}

extern "C" int jds_ctx_create(int device, jds_ctx** out) {
    if (!out) return fail(JDS_ERR_INVALID, "ctx out pointer is NULL");
    *out = nullptr;
    int n = 0;
    JDS_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n)
        return fail(JDS_ERR_INVALID, "device %d out of range (%d CUDA devices)", device, n);
    JDS_CUDA(cudaSetDevice(device));
    jds_ctx* c = new jds_ctx();
    c->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    for (int i = 0; i < 5 * jds_ctx::kTimedChunks && e == cudaSuccess; ++i) e = cudaEventCreate(&c->evs[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_setup, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        delete c;
        return fail(JDS_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
    }
    c->own_stream = true;
    // kernel attributes and constant tables of this device (no lazily initialised statics)
    e = fused_configure_device();
    if (e == cudaSuccess) e = ssim_configure_device();
    if (e == cudaSuccess) e = entropy_configure_device();
    if (e == cudaSuccess) e = exact_fused_configure_device();
    if (e != cudaSuccess) {
        jds_ctx_destroy(c);
        return fail(JDS_ERR_CUDA, "kernel setup failed: %s", cudaGetErrorString(e));
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    {
        int l2 = 0;
        if (cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, device) == cudaSuccess && l2 > 0)
            c->l2_bytes = (size_t)l2;
    }
    const char* ls = getenv("JDS_LEGACY_SSIM");
    c->legacy_ssim = ls && atoi(ls) != 0;
    const char* nf = getenv("JDS_NO_FUSED");
    c->no_fused = nf && atoi(nf) != 0;
    const char* nh = getenv("JDS_NO_HOIST");
    c->no_hoist = nh && atoi(nh) != 0;
    const char* l2c = getenv("JDS_L2_CHUNK");
    c->l2_chunking = l2c ? atoi(l2c) : 0;
    if (c->l2_chunking < 0) c->l2_chunking = 0;
    const char* nds = getenv("JDS_NO_DUAL_SWEEP");
    c->no_dual_sweep = nds && atoi(nds) != 0;
    const char* pc = getenv("JDS_PIPE_CHUNK");      // A/B runs: frames per chunk of the host pipeline
    c->pipe_chunk = pc ? atoi(pc) : 0;
    const char* mb = getenv("JDS_SCRATCH_MB");
    if (mb && atol(mb) > 0) c->scratch_budget = (size_t)atol(mb) << 20;
    *out = c;
    return JDS_OK;
}

extern "C" int jds_ctx_destroy(jds_ctx* c) {
    if (!c) return JDS_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->planes, &c->in, &c->recon, &c->coeffs, &c->errs,
                      &c->metrics, &c->tables, &c->selected, &c->payload, &c->alias, &c->fcoef,
                      &c->ent_sizes, &c->ent_bits, &c->ent_out, &c->band};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (c->h_metrics) cudaFreeHost(c->h_metrics);
    if (c->h_tables) cudaFreeHost(c->h_tables);
    for (int i = 0; i < jds_ctx::kTableRing; ++i) {
        if (c->h_tables_ring[i]) cudaFreeHost(c->h_tables_ring[i]);
        if (c->ev_tables[i]) cudaEventDestroy(c->ev_tables[i]);
    }
    if (c->h_selected) cudaFreeHost(c->h_selected);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int i = 0; i < 5 * jds_ctx::kTimedChunks; ++i)
        if (c->evs[i]) cudaEventDestroy(c->evs[i]);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    }
    if (c->ev_setup) cudaEventDestroy(c->ev_setup);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return JDS_OK;
}

extern "C" int jds_ctx_set_stream(jds_ctx* c, void* cuda_stream) {
    if (!c) return fail(JDS_ERR_INVALID, "ctx is NULL");
    JDS_CUDA(cudaSetDevice(c->device));
    JDS_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return JDS_OK;
}

// Order the context's stream after / before work of another stream of the same device (a
// cudaEvent_t handle, e.g. torch.cuda.Event().cuda_event): wait_event makes every later kernel
// of this context wait for the event (inputs produced on the caller's stream), record_event
// records it behind everything enqueued so far (results of the non-synchronising entry points).
extern "C" int jds_ctx_wait_event(jds_ctx* c, void* cuda_event) {
    if (!c || !cuda_event) return fail(JDS_ERR_INVALID, "NULL argument");
    JDS_CUDA(cudaSetDevice(c->device));
    JDS_CUDA(cudaStreamWaitEvent(c->stream, (cudaEvent_t)cuda_event, 0));
    return JDS_OK;
}

extern "C" int jds_ctx_record_event(jds_ctx* c, void* cuda_event) {
    if (!c || !cuda_event) return fail(JDS_ERR_INVALID, "NULL argument");
    JDS_CUDA(cudaSetDevice(c->device));
    JDS_CUDA(cudaEventRecord((cudaEvent_t)cuda_event, c->stream));
    return JDS_OK;
}

extern "C" int jds_ctx_synchronize(jds_ctx* c) {
    if (!c) return fail(JDS_ERR_INVALID, "ctx is NULL");
    JDS_CUDA(cudaSetDevice(c->device));
    JDS_CUDA(cudaStreamSynchronize(c->stream));
    return JDS_OK;
}

extern "C" int jds_ctx_stage_timing(jds_ctx* c, int enable) {
    if (!c) return fail(JDS_ERR_INVALID, "ctx is NULL");
    c->stage_timing = enable != 0;
    return JDS_OK;
}

extern "C" int jds_ctx_stage_times(jds_ctx* c, double ms[4], uint64_t launches[4], int reset) {
    if (!c || !ms || !launches) return fail(JDS_ERR_INVALID, "NULL argument");
    for (int i = 0; i < 4; ++i) {
        ms[i] = c->stage_ms[i];
        launches[i] = c->stage_launches[i];
        if (reset) {
            c->stage_ms[i] = 0;
            c->stage_launches[i] = 0;
        }
    }
    return JDS_OK;
}

extern "C" int jds_ctx_launch_count(jds_ctx* c, uint64_t* launches) {
    if (!c || !launches) return fail(JDS_ERR_INVALID, "NULL argument");
    *launches = c->launches;
    return JDS_OK;
}

// ------------------------------------------------------------------------------
// host-only helpers
// ------------------------------------------------------------------------------
extern "C" int jds_quant_table(int quality, double table[64]) {
    if (!table) return fail(JDS_ERR_INVALID, "table is NULL");
    if (quality < 1 || quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", quality);
    quant_table_host(quality, table);
    return JDS_OK;
}

static int make_geom(int H, int W, int sub, Geom* g) {
    switch (geom_init(H, W, sub, g)) {
        case 0: return JDS_OK;
        case 1: return fail(JDS_ERR_INVALID, "bad frame size %dx%d", H, W);
        case 2: return fail(JDS_ERR_INVALID, "Unknown subsampling mode: %d", sub);
        default: return fail(JDS_ERR_INVALID, "bad frame geometry %dx%d mode %d", H, W, sub);
    }
}

extern "C" int jds_coeff_count(int height, int width, int subsampling, uint64_t* count) {
    if (!count) return fail(JDS_ERR_INVALID, "count is NULL");
    Geom g;
    int rc = make_geom(height, width, subsampling, &g);
    if (rc) return rc;
    *count = 64ull * (uint64_t)(g.nblk_y + 2 * g.nblk_c);
    return JDS_OK;
}

extern "C" int jds_plane_dims(int height, int width, int subsampling, int* ch, int* cw) {
    if (!ch || !cw) return fail(JDS_ERR_INVALID, "NULL argument");
    Geom g;
    int rc = make_geom(height, width, subsampling, &g);
    if (rc) return rc;
    *ch = g.hc;
    *cw = g.wc;
    return JDS_OK;
}

static int check_params(const jds_params* p) {
    if (!p) return fail(JDS_ERR_INVALID, "params is NULL");
    if (p->precision != JDS_EXACT && p->precision != JDS_FAST)
        return fail(JDS_ERR_INVALID, "bad precision %d", p->precision);
    return JDS_OK;
}

// ------------------------------------------------------------------------------
// the launch sequence over `units` units (frames of a batch or sweep points)
// ------------------------------------------------------------------------------
// One fp64 record per unit from the device accumulators (jds_sweep_records): the rows a
// sharded sweep all-gathers, built where the partials already are.  Rows past `units` are
// marked empty (unit = -1) so that every rank contributes the same number of rows.
struct RecordQualities {
    short q[JDS_SWEEP_RECORDS_MAX];
};

__global__ void k_pack_records(const DevMetrics* __restrict__ m, int units, int capacity, int unit0,
                               int unit_step, RecordQualities rq, double ssim_count,
                               double total_coeffs, double luma_blocks, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= capacity) return;
    double* r = out + (size_t)i * JDS_RECORD_FIELDS;
    if (i >= units) {
        r[0] = -1.0;
        for (int k = 1; k < JDS_RECORD_FIELDS; ++k) r[k] = 0.0;
        return;
    }
    const DevMetrics& d = m[i];
    r[0] = (double)(unit0 + i * unit_step);
    r[1] = (double)rq.q[i];
    r[2] = (double)d.sse_rgb;
    r[3] = d.sse_y;
    r[4] = d.ssim_sum[0];
    r[5] = d.ssim_sum[1];
    r[6] = d.ssim_sum[2];
    r[7] = d.ssim_sum[3];
    r[8] = ssim_count;
    r[9] = (double)d.coeff_bits;
    r[10] = (double)d.nnz;
    r[11] = total_coeffs;
    r[12] = luma_blocks;
}

struct UnitJob {
    const jds_params* p;
    Geom g;
    int units;
    const int32_t* qualities;   // per unit (sweep) or NULL (params->quality for all)
    bool shared_input;          // sweep: one frame, forward stage run once
    const uint8_t* rgb;
    int rgb_loc;
    uint8_t* recon;
    int16_t* coeffs;
    double* err_y;
    double* err_rgb;
    int out_loc;
    jds_metrics* metrics;
    // jds_sweep_records: device-resident fp64 records instead of host structs, no synchronisation
    double* d_records;
    int rec_capacity, unit0, unit_step;
    // jds_roundtrip_band: no whole-frame comparison; `tail` enqueues the band's own reductions and
    // copies behind the kernels, before the job's single synchronisation
    bool no_compare;
    bool defer;                 // jds_roundtrip_batch_begin: return without synchronising
    int (*tail)(jds_ctx*, void*, cudaStream_t);
    void* tail_arg;
};

// Kernels of one chunk of `n` units on the compute stream.  Returns which stages ran.
struct ChunkPtrs {
    const uint8_t* d_rgb;
    size_t rgb_stride;
    uint8_t* d_recon;
    int16_t* d_coeffs;
    double* d_ey;
    double* d_ergb;
    const QTables* d_tables;
    DevMetrics* d_metrics;
    bool first_chunk;
    bool hoist;                 // sweep: forward half once per frame (c->fcoef), see launch_chunk
};

static int launch_chunk(jds_ctx* c, const UnitJob& J, const ChunkPtrs& P, int n, cudaStream_t s,
                        int scratch_slot, cudaEvent_t* evs /* 5 events or NULL */, bool ran[4]) {
    const bool timed = evs != nullptr;
    const Geom& g = J.g;
    const jds_params* p = J.p;
    const bool exact = p->precision == JDS_EXACT;
    const size_t esz = exact ? sizeof(double) : sizeof(float);
    const size_t frame_bytes = (size_t)g.H * g.W * 3;
    const size_t planes_elems = (size_t)(g.plane_y + 2 * g.plane_c);
    const size_t ncoef = 64ull * (size_t)(g.nblk_y + 2 * g.nblk_c);
    const bool want_hist = (p->outputs & JDS_OUT_HIST) != 0;
    const bool want_ssim = (p->outputs & JDS_OUT_SSIM) != 0;
    const int tstride = J.qualities ? 1 : 0;
    const int fwd_units = J.shared_input ? 1 : c->plan_chunk;
    char* planes = (char*)c->planes.p;
    void* fwd = planes;
    void* rec = planes + planes_elems * esz * (size_t)fwd_units;
    const size_t fwd_stride = J.shared_input ? 0 : planes_elems;
    const size_t rec_stride = planes_elems;

    if (timed) JDS_CUDA(cudaEventRecord(evs[0], s));
    // fast mode on block-aligned frames runs the fused kernels (jds_fused.cu); exact mode,
    // ragged frames and the fp64 error maps run the staged kernels (jds_kernels.cu)
    const bool fused = !exact && !c->no_fused && !P.d_ey && !P.d_ergb &&
                       fused_supported(g, p->prefilter, P.d_rgb, P.rgb_stride, P.d_recon, frame_bytes) &&
                       ssim_strip_supported(g.H, g.W, P.d_rgb, P.rgb_stride, P.d_recon, frame_bytes);
    if (fused) {
        const size_t cpl_stride = fused_chroma_plane_floats(g);
        // each compute stream has its own chroma-plane scratch
        float* cpl = (float*)c->planes.p + (size_t)scratch_slot * cpl_stride * (size_t)c->plan_chunk;
        // sweeps (one shared frame, metrics only): the quality-independent forward half - colour,
        // decimation / prefilter, forward DCT - ran once per frame into c->fcoef (run_job's pre-pass),
        // every quality point starts at quantisation (gui/worker.py:55-74 redoes it all)
        const bool hoist = P.hoist;
        float* fcoef = hoist ? (float*)c->fcoef.p : nullptr;
        const int stage = hoist ? 2 : 0;
        ran[0] = g.sub != 0;
        if (ran[0]) {
            JDS_CUDA(launch_fused_chroma(g, p->prefilter, P.d_rgb, P.rgb_stride, cpl, cpl_stride, P.d_tables,
                                         tstride, P.d_coeffs, ncoef, P.d_metrics, n, s, stage, fcoef));
            c->launches++;
        }
        if (timed) JDS_CUDA(cudaEventRecord(evs[1], s));
        JDS_CUDA(launch_fused_luma(g, P.d_rgb, P.rgb_stride, cpl, cpl_stride, P.d_tables, tstride,
                                   P.d_coeffs, ncoef, P.d_recon, frame_bytes, P.d_metrics, n, s, stage, fcoef));
        c->launches++;
        if (timed) {
            JDS_CUDA(cudaEventRecord(evs[2], s));
            JDS_CUDA(cudaEventRecord(evs[3], s));
        }
        ran[1] = true;
        ran[2] = false;
        ran[3] = !J.no_compare;      // squared errors always come from the strip kernel here
        if (ran[3]) {
            JDS_CUDA(launch_ssim_strip(g.H, g.W, P.d_rgb, P.rgb_stride, P.d_recon, frame_bytes,
                                       P.d_metrics, n, want_ssim, true, c->sm_count, s));
            c->launches++;
        }
    } else if (exact && !c->no_fused && !P.d_ey && !P.d_ergb && !g.general &&
               fused_supported(g, p->prefilter, P.d_rgb, P.rgb_stride, P.d_recon, frame_bytes) &&
               ssim_strip_supported(g.H, g.W, P.d_rgb, P.rgb_stride, P.d_recon, frame_bytes)) {
        // exact mode on block-aligned frames: chroma through the staged forward / codec kernels
        // (decimated planes only), luma + compose fused (jds_fused_exact.cu) - no fp64 luma planes
        // in HBM; squared errors come from the strip kernel (integer SSE: psnr_rgb stays exact)
        if (exact_chroma_supported(g, p->prefilter)) {
            // decimation + chroma codec in one kernel (stage "forward_colour" in the timings)
            ran[0] = true;
            JDS_CUDA(launch_exact_chroma(g, p->prefilter, P.d_rgb, P.rgb_stride, (double*)rec, rec_stride, P.d_tables,
                                         tstride, P.d_coeffs, ncoef, P.d_metrics, n, s));
            c->launches++;
            if (timed) JDS_CUDA(cudaEventRecord(evs[1], s));
        } else {
            ran[0] = !J.shared_input || P.first_chunk;
            if (ran[0]) {
                launch_forward(true, g, p->prefilter, P.d_rgb, P.rgb_stride, fwd, fwd_stride,
                               J.shared_input ? 1 : n, s, true);
                c->launches++;
            }
            if (timed) JDS_CUDA(cudaEventRecord(evs[1], s));
            launch_codec(true, g, fwd, fwd_stride, rec, rec_stride, P.d_tables, tstride, P.d_coeffs, ncoef,
                         false, P.d_metrics, n, s, true);
            c->launches++;
        }
        JDS_CUDA(launch_exact_luma(g, P.d_rgb, P.rgb_stride, (const double*)rec, rec_stride, P.d_tables,
                                   tstride, P.d_coeffs, ncoef, P.d_recon, frame_bytes, P.d_metrics, n, s));
        c->launches++;
        if (timed) {
            JDS_CUDA(cudaEventRecord(evs[2], s));
            JDS_CUDA(cudaEventRecord(evs[3], s));
        }
        ran[1] = true;
        ran[2] = false;
        ran[3] = !J.no_compare;
        if (ran[3]) {
            JDS_CUDA(launch_ssim_strip(g.H, g.W, P.d_rgb, P.rgb_stride, P.d_recon, frame_bytes,
                                       P.d_metrics, n, want_ssim, true, c->sm_count, s));
            c->launches++;
        }
    } else {
        ran[0] = !J.shared_input || P.first_chunk;
        ran[1] = ran[2] = true;
        if (ran[0]) {
            launch_forward(exact, g, p->prefilter, P.d_rgb, P.rgb_stride, fwd, fwd_stride,
                           J.shared_input ? 1 : n, s);
            c->launches++;
        }
        if (timed) JDS_CUDA(cudaEventRecord(evs[1], s));
        launch_codec(exact, g, fwd, fwd_stride, rec, rec_stride, P.d_tables, tstride, P.d_coeffs,
                     ncoef, false, P.d_metrics, n, s);
        if (timed) JDS_CUDA(cudaEventRecord(evs[2], s));
        launch_inverse(exact, g, P.d_rgb, P.rgb_stride, fwd, fwd_stride, rec, rec_stride, P.d_recon,
                       frame_bytes, P.d_ey, P.d_ergb, P.d_metrics, n, s);
        if (timed) JDS_CUDA(cudaEventRecord(evs[3], s));
        c->launches += 2;
        ran[3] = want_ssim && g.H >= 7 && g.W >= 7;
        if (ran[3]) {
            if (!c->legacy_ssim &&
                ssim_strip_supported(g.H, g.W, P.d_rgb, P.rgb_stride, P.d_recon, frame_bytes)) {
                JDS_CUDA(launch_ssim_strip(g.H, g.W, P.d_rgb, P.rgb_stride, P.d_recon, frame_bytes,
                                           P.d_metrics, n, true, false, c->sm_count, s));
            } else {
                launch_ssim(exact, g.H, g.W, P.d_rgb, P.rgb_stride, P.d_recon, frame_bytes,
                            P.d_metrics, n, s);
            }
            c->launches++;
        }
    }
    if (want_hist && P.d_coeffs) {
        // 50-bin histogram from the coefficient buffer just written (both code paths)
        launch_hist50(P.d_coeffs, ncoef, ncoef, P.d_metrics, n, c->sm_count, s);
        c->launches++;
    }
    if (timed) JDS_CUDA(cudaEventRecord(evs[4], s));
    JDS_CUDA(cudaGetLastError());
    return JDS_OK;
}

// The whole job: chunks of units flow through three streams - host->device copies,
// kernels, device->host copies - with double-buffered staging, so that with host
// buffers the PCIe transfers of neighbouring chunks overlap the kernels and each other
// (full duplex).  With device buffers the copy streams stay idle.
// every entry point that uses the context's pinned result buffers refuses to run while a deferred
// batch (jds_roundtrip_batch_begin) is waiting for jds_ctx_finish
#define JDS_NO_PENDING(c)                                                                        \
    if ((c)->pend.active)                                                                        \
        return fail(JDS_ERR_INVALID, "a deferred batch is pending on this context: call jds_ctx_finish first")

// sync + host-side metric structs of a finished job (shared by run_job and jds_ctx_finish)
static void fill_job_metrics(const jds_ctx* c, jds_metrics* out, int units, uint64_t ssim_count,
                             uint64_t ncoef, uint64_t luma_blocks, double ms_per_unit) {
    const DevMetrics* h_metrics = (const DevMetrics*)c->h_metrics;
    for (int i = 0; i < units; ++i) {
        jds_metrics* m = &out[i];
        const DevMetrics& d = h_metrics[i];
        memset(m, 0, sizeof *m);
        m->sse_rgb = d.sse_rgb;
        m->sse_y = d.sse_y;
        for (int k = 0; k < 4; ++k) m->ssim_sum[k] = d.ssim_sum[k];
        m->ssim_count = ssim_count;
        m->coeff_bits = d.coeff_bits;
        m->nnz = d.nnz;
        m->total_coeffs = ncoef;
        m->luma_blocks = luma_blocks;
        for (int k = 0; k < 50; ++k) m->hist50[k] = (int64_t)d.hist[k];
        m->gpu_ms = ms_per_unit;
    }
}

static int run_job(jds_ctx* c, const UnitJob& J) {
    JDS_NO_PENDING(c);
    const Geom& g = J.g;
    const jds_params* p = J.p;
    const bool exact = p->precision == JDS_EXACT;
    const size_t esz = exact ? sizeof(double) : sizeof(float);
    const size_t frame_bytes = (size_t)g.H * g.W * 3;
    const size_t planes_elems = (size_t)(g.plane_y + 2 * g.plane_c);
    const size_t ncoef = 64ull * (size_t)(g.nblk_y + 2 * g.nblk_c);
    const bool want_coeffs = (p->outputs & JDS_OUT_COEFFS) && J.coeffs;
    const bool want_hist = (p->outputs & JDS_OUT_HIST) != 0;
    const bool want_ssim = (p->outputs & JDS_OUT_SSIM) != 0;
    const bool want_recon = (p->outputs & JDS_OUT_RECON) && J.recon;
    const bool want_ey = (p->outputs & JDS_OUT_ERR_Y) && J.err_y;
    const bool want_ergb = (p->outputs & JDS_OUT_ERR_RGB) && J.err_rgb;
    if ((want_ey || want_ergb) && J.units != 1)
        return fail(JDS_ERR_INVALID, "error maps are single-frame outputs");
    const bool in_host = J.rgb_loc == JDS_HOST;
    const bool out_host = J.out_loc == JDS_HOST;

    JDS_CUDA(cudaSetDevice(c->device));
    // units per chunk: bounded by the scratch budget; with host buffers also small enough
    // that there are several chunks to overlap
    const size_t per_unit = planes_elems * esz * (J.shared_input ? 1 : 2) + frame_bytes * 2 +
                            ((want_coeffs || want_hist) ? ncoef * 2 : 0);
    int chunk = (int)(c->scratch_budget / (per_unit ? per_unit : 1));
    if (chunk < 1) chunk = 1;
    if (chunk > J.units) chunk = J.units;
    if (chunk > 65535) chunk = 65535;
    if (!exact && c->l2_chunking) {
        // JDS_L2_CHUNK=1: keep the working set of one launch sequence - input, output, chroma
        // planes - inside L2 so that DRAM sees each image once in and once out (the fused
        // kernels pass a frame from one to the next through L2).  Off by default: on B200 the
        // path is issue-bound, not DRAM-bound, and one-frame launches lose ~25 % to wave tails
        // and launch gaps (measured: 36.3 vs 47.1 Gpixel/s on 8 x 4K).
        const size_t ws = frame_bytes * 2 + fused_chroma_plane_floats(g) * sizeof(float) +
                          (want_coeffs ? ncoef * 2 : 0);
        int lc = (int)((c->l2_bytes * 6 / 10) / (ws ? ws : 1));
        if (lc < 1) lc = 1;
        lc *= c->l2_chunking;          // k > 1: trade some L2 residency for larger grids
        if (chunk > lc) chunk = lc;
    }
    const bool pipelined = (in_host || (out_host && (want_recon || want_coeffs))) && J.units > 1;
    if (pipelined) {
        // many small chunks hide the pipeline fill / drain (one copy each way), but keep
        // every copy >= 8 MB so PCIe stays efficient
        int target = (J.units + 15) / 16;
        const size_t floor_bytes = (size_t)8 << 20;
        const int min_units = (int)((floor_bytes + frame_bytes - 1) / frame_bytes);
        if (target < min_units) target = min_units;
        if (target < 1) target = 1;
        if (chunk > target) chunk = target;
        if (c->pipe_chunk > 0) chunk = c->pipe_chunk < J.units ? c->pipe_chunk : J.units;
    }
    c->plan_chunk = chunk;
    const int nbuf = pipelined ? 2 : 1;
    // L2-sized launch sequences: alternate chunks over two compute streams so the wave tail
    // and launch gaps of one frame's kernels are filled by the next frame's (fast mode only;
    // off while per-kernel timing is on, which wants kernels one at a time)
    // Only the fused kernels have per-stream scratch (chroma planes per slot, recon / coeff
    // scratch per slot below); the staged fallback shares its fwd / rec planes, so it never
    // runs on two streams (chunk k+1 overwrote what chunk k's SSIM kernel still read).
    const uint8_t* probe_in = in_host ? (const uint8_t*)nullptr : J.rgb;
    const uint8_t* probe_out = (want_recon && !out_host) ? J.recon : (const uint8_t*)nullptr;
    const bool will_fuse = !exact && !c->no_fused && !want_ey && !want_ergb &&
                           fused_supported(g, p->prefilter, probe_in, frame_bytes, probe_out, frame_bytes) &&
                           ssim_strip_supported(g.H, g.W, probe_in, frame_bytes, probe_out, frame_bytes);
    // hoisted sweep: one shared frame, fused kernels, metrics (and optionally recon) only
    const bool hoist = J.shared_input && will_fuse && J.units > 1 && !want_coeffs && !want_hist &&
                       fused_fcoef_floats(g) > 0 && !c->no_hoist;
    // A hoisted sweep's points split into two halves that alternate over the two compute streams
    // (after the pre-pass): the ragged last wave of one half's kernels is filled by the other
    // half's - what limits the strong scaling of a 12 / 13-point share (JDS_NO_DUAL_SWEEP=1: off)
    const bool dual_sweep = hoist && !c->stage_timing && !c->no_dual_sweep && J.units >= 4;
    if (dual_sweep && chunk >= J.units) {
        chunk = (J.units + 1) / 2;
        c->plan_chunk = chunk;
    }
    const bool dual = (c->l2_chunking > 0 && will_fuse && !c->stage_timing && J.units > chunk && !hoist) ||
                      (dual_sweep && J.units > chunk);
    const int nscr = (pipelined || dual) ? 2 : 1;        // recon / coefficient scratch slots

    int rc;
    const int fwd_units = J.shared_input ? 1 : chunk;
    {
        size_t need = planes_elems * esz * (size_t)(fwd_units + chunk);
        const size_t need2 = 2 * fused_chroma_plane_floats(g) * sizeof(float) * (size_t)chunk;
        if (dual && need2 > need) need = need2;
        if ((rc = ensure(c, c->planes, need))) return rc;
    }
    if (hoist && (rc = ensure(c, c->fcoef, fused_fcoef_floats(g) * sizeof(float)))) return rc;
    if ((rc = ensure(c, c->metrics, sizeof(DevMetrics) * (size_t)J.units))) return rc;
    const int n_tables_total = J.qualities ? J.units : 1;
    if ((rc = ensure(c, c->tables, sizeof(QTables) * (size_t)n_tables_total))) return rc;
    if ((rc = ensure_pinned(&c->h_metrics, &c->h_metrics_bytes, sizeof(DevMetrics) * (size_t)J.units)))
        return rc;
    int tslot = -1;
    if (J.d_records) {
        tslot = c->tables_slot;
        c->tables_slot = (tslot + 1) % jds_ctx::kTableRing;
        if (!c->ev_tables[tslot])
            JDS_CUDA(cudaEventCreateWithFlags(&c->ev_tables[tslot], cudaEventDisableTiming));
        if (c->ev_tables_used[tslot]) JDS_CUDA(cudaEventSynchronize(c->ev_tables[tslot]));
        if ((rc = ensure_pinned(&c->h_tables_ring[tslot], &c->h_tables_ring_bytes[tslot],
                                sizeof(QTables) * (size_t)n_tables_total)))
            return rc;
    } else if ((rc = ensure_pinned(&c->h_tables, &c->h_tables_bytes, sizeof(QTables) * (size_t)n_tables_total)))
        return rc;
    const size_t in_slot = frame_bytes * (size_t)fwd_units;
    const size_t recon_slot = frame_bytes * (size_t)chunk;
    const size_t coeff_slot = ncoef * 2 * (size_t)chunk;
    if (in_host && (rc = ensure(c, c->in, in_slot * (J.shared_input ? 1 : nbuf)))) return rc;
    if ((!want_recon || out_host) && (rc = ensure(c, c->recon, recon_slot * nscr))) return rc;
    // the histogram is taken from the coefficient buffer: scratch when the caller wants the
    // histogram but not the coefficients themselves (or wants them on the host)
    const bool coeff_scratch = (want_coeffs && out_host) || (want_hist && !want_coeffs);
    if (coeff_scratch && (rc = ensure(c, c->coeffs, coeff_slot * nscr))) return rc;
    if ((want_ey || want_ergb) && out_host && (rc = ensure(c, c->errs, (size_t)g.H * g.W * 8 * 2)))
        return rc;

    DevMetrics* d_metrics = (DevMetrics*)c->metrics.p;
    QTables* d_tables = (QTables*)c->tables.p;
    QTables* h_tables = (QTables*)(tslot >= 0 ? c->h_tables_ring[tslot] : c->h_tables);
    DevMetrics* h_metrics = (DevMetrics*)c->h_metrics;
    cudaStream_t s = c->stream;
    cudaStream_t s_in = pipelined ? c->s_in : s;
    cudaStream_t s_out = pipelined ? c->s_out : s;

    // quantiser tables of every unit, metric accumulators
    for (int i = 0; i < n_tables_total; ++i)
        fill_tables(J.qualities ? J.qualities[i] : p->quality, &h_tables[i]);
    JDS_CUDA(cudaMemcpyAsync(d_tables, h_tables, sizeof(QTables) * n_tables_total,
                             cudaMemcpyHostToDevice, s));
    if (tslot >= 0) {
        JDS_CUDA(cudaEventRecord(c->ev_tables[tslot], s));
        c->ev_tables_used[tslot] = true;
    }
    JDS_CUDA(cudaMemsetAsync(d_metrics, 0, sizeof(DevMetrics) * J.units, s));
    if (J.shared_input && in_host)
        JDS_CUDA(cudaMemcpyAsync(c->in.p, J.rgb, frame_bytes, cudaMemcpyHostToDevice, s));
    if (hoist) {
        JDS_CUDA(cudaEventRecord(c->ev0, s));              // gpu_ms of a sweep includes its pre-pass
        // pre-pass of the sweep's one frame: colour, prefilter / decimation, forward DCT -> c->fcoef
        const uint8_t* d_frame = in_host ? (const uint8_t*)c->in.p : J.rgb;
        float* fcoef = (float*)c->fcoef.p;
        JDS_CUDA(launch_fused_chroma(g, p->prefilter, d_frame, 0, nullptr, 0, nullptr, 0, nullptr, 0,
                                     nullptr, 1, s, 1, fcoef));
        JDS_CUDA(launch_fused_luma(g, d_frame, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr,
                                   1, s, 1, fcoef));
        c->launches += 2;
    }
    if (dual) {
        JDS_CUDA(cudaEventRecord(c->ev_setup, s));
        JDS_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_setup, 0));
    }

    int chunk_idx = 0;
    bool timed_ran[jds_ctx::kTimedChunks][4];
    int n_timed = 0;
    for (int u0 = 0; u0 < J.units; u0 += chunk, ++chunk_idx) {
        const int n = (J.units - u0 < chunk) ? (J.units - u0) : chunk;
        const int b = chunk_idx % nbuf;
        const int slot = dual ? (chunk_idx & 1) : 0;
        const int sb = pipelined ? b : slot;               // output scratch slot of this chunk
        cudaStream_t cs = slot ? c->stream2 : s;           // compute stream of this chunk
        ChunkPtrs P;
        P.first_chunk = (u0 == 0);
        P.hoist = hoist;
        P.d_tables = d_tables + (J.qualities ? u0 : 0);
        P.d_metrics = d_metrics + u0;
        // --- inputs ---
        P.rgb_stride = J.shared_input ? 0 : frame_bytes;
        if (J.shared_input) {
            P.d_rgb = in_host ? (const uint8_t*)c->in.p : J.rgb;
        } else if (in_host) {
            uint8_t* slot = (uint8_t*)c->in.p + (size_t)b * in_slot;
            if (pipelined && chunk_idx >= nbuf)      // slot last read by the kernels of chunk-2
                JDS_CUDA(cudaStreamWaitEvent(s_in, c->ev_comp[b], 0));
            JDS_CUDA(cudaMemcpyAsync(slot, J.rgb + (size_t)u0 * frame_bytes, frame_bytes * (size_t)n,
                                     cudaMemcpyHostToDevice, s_in));
            if (pipelined) {
                JDS_CUDA(cudaEventRecord(c->ev_in[b], s_in));
                JDS_CUDA(cudaStreamWaitEvent(cs, c->ev_in[b], 0));
            }
            P.d_rgb = slot;
        } else {
            P.d_rgb = J.rgb + (size_t)u0 * frame_bytes;
        }
        // --- outputs ---
        P.d_recon = (want_recon && !out_host) ? J.recon + (size_t)u0 * frame_bytes
                                              : (uint8_t*)c->recon.p + (size_t)sb * recon_slot;
        P.d_coeffs = nullptr;
        if (coeff_scratch)
            P.d_coeffs = (int16_t*)((char*)c->coeffs.p + (size_t)sb * coeff_slot);
        else if (want_coeffs)
            P.d_coeffs = J.coeffs + (size_t)u0 * ncoef;
        P.d_ey = want_ey ? (out_host ? (double*)c->errs.p : J.err_y) : nullptr;
        P.d_ergb = want_ergb ? (out_host ? (double*)c->errs.p + (size_t)g.H * g.W : J.err_rgb) : nullptr;
        if (pipelined && chunk_idx >= nbuf)          // staging slot still being copied out?
            JDS_CUDA(cudaStreamWaitEvent(cs, c->ev_out[b], 0));

        // ev0 .. ev1 bracket the kernels: after the first chunk's input is on its way, before
        // the last chunk's results are copied out
        if (u0 == 0 && !hoist) JDS_CUDA(cudaEventRecord(c->ev0, s));
        // per-stage events for the first kTimedChunks chunks of the call
        cudaEvent_t* evs = (c->stage_timing && !J.d_records && chunk_idx < jds_ctx::kTimedChunks)
                               ? &c->evs[5 * chunk_idx] : nullptr;
        bool ran[4];
        if ((rc = launch_chunk(c, J, P, n, cs, slot, evs, ran))) return rc;
        if (evs) {
            n_timed = chunk_idx + 1;
            for (int k = 0; k < 4; ++k) timed_ran[chunk_idx][k] = ran[k];
        }
        if (u0 + n >= J.units) {
            if (dual) {                              // join the second compute stream
                JDS_CUDA(cudaEventRecord(c->ev_join, c->stream2));
                JDS_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
            }
            JDS_CUDA(cudaEventRecord(c->ev1, s));
        }
        if (pipelined) {
            JDS_CUDA(cudaEventRecord(c->ev_comp[b], cs));
            JDS_CUDA(cudaStreamWaitEvent(s_out, c->ev_comp[b], 0));
        }
        // --- results back ---
        if (out_host) {
            if (want_recon)
                JDS_CUDA(cudaMemcpyAsync(J.recon + (size_t)u0 * frame_bytes, P.d_recon,
                                         frame_bytes * (size_t)n, cudaMemcpyDeviceToHost, s_out));
            if (want_coeffs)
                JDS_CUDA(cudaMemcpyAsync(J.coeffs + (size_t)u0 * ncoef, P.d_coeffs,
                                         ncoef * 2 * (size_t)n, cudaMemcpyDeviceToHost, s_out));
            if (want_ey)
                JDS_CUDA(cudaMemcpyAsync(J.err_y, P.d_ey, (size_t)g.H * g.W * 8,
                                         cudaMemcpyDeviceToHost, s_out));
            if (want_ergb)
                JDS_CUDA(cudaMemcpyAsync(J.err_rgb, P.d_ergb, (size_t)g.H * g.W * 8,
                                         cudaMemcpyDeviceToHost, s_out));
            if (pipelined) JDS_CUDA(cudaEventRecord(c->ev_out[b], s_out));
        }
    }
    if (J.d_records) {
        // device-resident records, no synchronisation: the caller orders its consumers on
        // the context's stream (jds_ctx_set_stream) and synchronises once
        if (pipelined) {
            JDS_CUDA(cudaStreamSynchronize(s_in));
            JDS_CUDA(cudaStreamSynchronize(s_out));
        }
        RecordQualities rq;
        for (int i = 0; i < J.units; ++i) rq.q[i] = (short)(J.qualities ? J.qualities[i] : p->quality);
        const double cnt = (want_ssim && g.H >= 7 && g.W >= 7) ? (double)(g.H - 6) * (double)(g.W - 6) : 0.0;
        const int cap = J.rec_capacity;
        k_pack_records<<<(cap + 127) / 128, 128, 0, s>>>(d_metrics, J.units, cap, J.unit0, J.unit_step, rq,
                                                         cnt, (double)ncoef, (double)g.nblk_y, J.d_records);
        JDS_CUDA(cudaGetLastError());
        c->launches++;
        return JDS_OK;
    }
    if (J.tail && (rc = J.tail(c, J.tail_arg, s))) return rc;
    JDS_CUDA(cudaMemcpyAsync(h_metrics, d_metrics, sizeof(DevMetrics) * J.units,
                             cudaMemcpyDeviceToHost, s));
    const uint64_t ssim_count = (want_ssim && g.H >= 7 && g.W >= 7)
                                    ? (uint64_t)(g.H - 6) * (uint64_t)(g.W - 6) : 0;
    if (J.defer) {
        // everything is enqueued; jds_ctx_finish synchronises and fills the metric structs
        c->pend.active = true;
        c->pend.pipelined = pipelined;
        c->pend.units = J.units;
        c->pend.metrics = J.metrics;
        c->pend.ssim_count = ssim_count;
        c->pend.ncoef = ncoef;
        c->pend.luma_blocks = (uint64_t)g.nblk_y;
        return JDS_OK;
    }
    if (pipelined) {
        JDS_CUDA(cudaStreamSynchronize(s_in));
        JDS_CUDA(cudaStreamSynchronize(s_out));
    }
    JDS_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    JDS_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    for (int ci = 0; ci < n_timed; ++ci)
        for (int k = 0; k < 4; ++k) {
            float t = 0.f;
            JDS_CUDA(cudaEventElapsedTime(&t, c->evs[5 * ci + k], c->evs[5 * ci + k + 1]));
            if (timed_ran[ci][k]) {
                c->stage_ms[k] += t;
                c->stage_launches[k] += 1;
            }
        }
    fill_job_metrics(c, J.metrics, J.units, ssim_count, ncoef, (uint64_t)g.nblk_y, (double)ms / J.units);
    return JDS_OK;
}

// ------------------------------------------------------------------------------
// entry points
// ------------------------------------------------------------------------------
extern "C" int jds_roundtrip(jds_ctx* c, const jds_params* p, const uint8_t* rgb, int rgb_loc,
                             uint8_t* recon, int16_t* coeffs, double* err_y, double* err_rgb,
                             int out_loc, jds_metrics* metrics) {
    if (!c || !rgb || !metrics) return fail(JDS_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    J.p = p;
    J.units = 1;
    J.rgb = rgb;
    J.rgb_loc = rgb_loc;
    J.recon = recon;
    J.coeffs = coeffs;
    J.err_y = err_y;
    J.err_rgb = err_rgb;
    J.out_loc = out_loc;
    J.metrics = metrics;
    return run_job(c, J);
}

extern "C" int jds_roundtrip_batch(jds_ctx* c, const jds_params* p, int n_frames,
                                   const uint8_t* rgb, int rgb_loc, uint8_t* recon,
                                   int16_t* coeffs, int out_loc, jds_metrics* metrics) {
    if (!c || !rgb || !metrics) return fail(JDS_ERR_INVALID, "NULL argument");
    if (n_frames < 1) return fail(JDS_ERR_INVALID, "n_frames must be >= 1, got %d", n_frames);
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    if (p->outputs & (JDS_OUT_ERR_Y | JDS_OUT_ERR_RGB))
        return fail(JDS_ERR_INVALID, "error maps are not produced in batch mode");
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    J.p = p;
    J.units = n_frames;
    J.rgb = rgb;
    J.rgb_loc = rgb_loc;
    J.recon = recon;
    J.coeffs = coeffs;
    J.out_loc = out_loc;
    J.metrics = metrics;
    return run_job(c, J);
}

// The batch call split in two (pipelining consecutive batches over two contexts: while one
// context's last chunks drain over PCIe the other's first chunks are already on their way).
// begin enqueues everything - host->device copies, kernels, device->host copies - and returns;
// the buffers and the metrics array must stay valid until jds_ctx_finish, which synchronises and
// fills the metric structs.  No other call on this context in between.
extern "C" int jds_roundtrip_batch_begin(jds_ctx* c, const jds_params* p, int n_frames,
                                         const uint8_t* rgb, int rgb_loc, uint8_t* recon,
                                         int16_t* coeffs, int out_loc, jds_metrics* metrics) {
    if (!c || !rgb || !metrics) return fail(JDS_ERR_INVALID, "NULL argument");
    if (n_frames < 1) return fail(JDS_ERR_INVALID, "n_frames must be >= 1, got %d", n_frames);
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    if (p->outputs & (JDS_OUT_ERR_Y | JDS_OUT_ERR_RGB))
        return fail(JDS_ERR_INVALID, "error maps are not produced in batch mode");
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    J.p = p;
    J.units = n_frames;
    J.rgb = rgb;
    J.rgb_loc = rgb_loc;
    J.recon = recon;
    J.coeffs = coeffs;
    J.out_loc = out_loc;
    J.metrics = metrics;
    J.defer = true;
    return run_job(c, J);
}

extern "C" int jds_ctx_finish(jds_ctx* c) {
    if (!c) return fail(JDS_ERR_INVALID, "ctx is NULL");
    if (!c->pend.active) return JDS_OK;
    JDS_CUDA(cudaSetDevice(c->device));
    c->pend.active = false;
    if (c->pend.pipelined) {
        JDS_CUDA(cudaStreamSynchronize(c->s_in));
        JDS_CUDA(cudaStreamSynchronize(c->s_out));
    }
    JDS_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    JDS_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    fill_job_metrics(c, c->pend.metrics, c->pend.units, c->pend.ssim_count, c->pend.ncoef,
                     c->pend.luma_blocks, (double)ms / c->pend.units);
    return JDS_OK;
}

extern "C" int jds_sweep(jds_ctx* c, const jds_params* p, const int32_t* qualities, int n_q,
                         const uint8_t* rgb, int rgb_loc, uint8_t* recon, int out_loc,
                         jds_metrics* metrics) {
    if (!c || !rgb || !metrics || !qualities) return fail(JDS_ERR_INVALID, "NULL argument");
    if (n_q < 1) return fail(JDS_ERR_INVALID, "n_q must be >= 1, got %d", n_q);
    int rc = check_params(p);
    if (rc) return rc;
    for (int i = 0; i < n_q; ++i)
        if (qualities[i] < 1 || qualities[i] > 100)
            return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", qualities[i]);
    if (p->outputs & (JDS_OUT_ERR_Y | JDS_OUT_ERR_RGB | JDS_OUT_COEFFS))
        return fail(JDS_ERR_INVALID, "sweeps produce metrics and (optionally) recon only");
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    J.p = p;
    J.units = n_q;
    J.qualities = qualities;
    J.shared_input = true;
    J.rgb = rgb;
    J.rgb_loc = rgb_loc;
    J.recon = recon;
    J.out_loc = out_loc;
    J.metrics = metrics;
    return run_job(c, J);
}

// Sweep whose results stay on the device (sharded sweeps, BASELINE config 4): the kernels of
// jds_sweep, then one fp64 record per quality written to `records` (device memory,
// `capacity` rows of JDS_RECORD_FIELDS doubles; rows n_q..capacity-1 are marked empty with
// unit = -1).  Returns WITHOUT synchronising: everything is ordered on the context's stream, so
// a collective or copy the caller enqueues on the same stream (jds_ctx_set_stream) sees the
// records, and one synchronisation at the very end replaces the per-call one.
extern "C" int jds_sweep_records(jds_ctx* c, const jds_params* p, const int32_t* qualities, int n_q,
                                 const uint8_t* rgb, int rgb_loc, int unit0, int unit_step,
                                 double* records, int capacity) {
    if (!c || !rgb || !records || (!qualities && n_q > 0)) return fail(JDS_ERR_INVALID, "NULL argument");
    if (n_q < 0 || n_q > capacity || capacity < 1 || capacity > JDS_SWEEP_RECORDS_MAX)
        return fail(JDS_ERR_INVALID, "need 0 <= n_q <= capacity <= %d, got n_q %d capacity %d",
                    JDS_SWEEP_RECORDS_MAX, n_q, capacity);
    int rc = check_params(p);
    if (rc) return rc;
    for (int i = 0; i < n_q; ++i)
        if (qualities[i] < 1 || qualities[i] > 100)
            return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", qualities[i]);
    if (p->outputs & (JDS_OUT_ERR_Y | JDS_OUT_ERR_RGB | JDS_OUT_COEFFS | JDS_OUT_RECON | JDS_OUT_HIST))
        return fail(JDS_ERR_INVALID, "jds_sweep_records produces metric records only");
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    if (n_q == 0) {
        // a rank that owns no point still contributes `capacity` empty rows
        JDS_CUDA(cudaSetDevice(c->device));
        RecordQualities rq = {};
        k_pack_records<<<(capacity + 127) / 128, 128, 0, c->stream>>>(nullptr, 0, capacity, unit0, unit_step,
                                                                      rq, 0.0, 0.0, 0.0, records);
        JDS_CUDA(cudaGetLastError());
        c->launches++;
        return JDS_OK;
    }
    J.p = p;
    J.units = n_q;
    J.qualities = qualities;
    J.shared_input = true;
    J.rgb = rgb;
    J.rgb_loc = rgb_loc;
    J.out_loc = JDS_DEVICE;
    J.d_records = records;
    J.rec_capacity = capacity;
    J.unit0 = unit0;
    J.unit_step = unit_step;
    return run_job(c, J);
}

// jds_roundtrip_batch whose metric records stay on the device (same contract as
// jds_sweep_records: rows of JDS_RECORD_FIELDS doubles, no synchronisation).  Frames and the
// optional reconstruction are DEVICE buffers; a caller streaming batches enqueues step after
// step (and, sharded over GPUs, one all-reduce of the rows per step) and synchronises once.
extern "C" int jds_roundtrip_batch_records(jds_ctx* c, const jds_params* p, int n_frames,
                                           const uint8_t* rgb, uint8_t* recon, int unit0,
                                           int unit_step, double* records, int capacity) {
    if (!c || !rgb || !records) return fail(JDS_ERR_INVALID, "NULL argument");
    if (n_frames < 1 || n_frames > capacity || capacity > JDS_SWEEP_RECORDS_MAX)
        return fail(JDS_ERR_INVALID, "need 1 <= n_frames <= capacity <= %d, got n_frames %d capacity %d",
                    JDS_SWEEP_RECORDS_MAX, n_frames, capacity);
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    if (p->outputs & (JDS_OUT_ERR_Y | JDS_OUT_ERR_RGB | JDS_OUT_COEFFS | JDS_OUT_HIST))
        return fail(JDS_ERR_INVALID, "jds_roundtrip_batch_records produces recon and metric records only");
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    J.p = p;
    J.units = n_frames;
    J.rgb = rgb;
    J.rgb_loc = JDS_DEVICE;
    J.recon = recon;
    J.out_loc = JDS_DEVICE;
    J.d_records = records;
    J.rec_capacity = capacity;
    J.unit0 = unit0;
    J.unit_step = unit_step;
    return run_job(c, J);
}

// ------------------------------------------------------------------------------
// chroma-aliasing demo
// ------------------------------------------------------------------------------
// squared errors + SSIM sums of two device-resident uint8 RGB frames into *m (synchronises)
static int compare_device(jds_ctx* c, const uint8_t* d_a, const uint8_t* d_b, int height, int width,
                          jds_metrics* m) {
    const size_t frame_bytes = (size_t)height * width * 3;
    int rc;
    if ((rc = ensure(c, c->metrics, sizeof(DevMetrics)))) return rc;
    if ((rc = ensure_pinned(&c->h_metrics, &c->h_metrics_bytes, sizeof(DevMetrics)))) return rc;
    cudaStream_t s = c->stream;
    DevMetrics* dm = (DevMetrics*)c->metrics.p;
    JDS_CUDA(cudaMemsetAsync(dm, 0, sizeof(DevMetrics), s));
    const bool ssim_ok = height >= 7 && width >= 7;
    launch_sse_u8(d_a, d_b, (long long)height * width, dm, c->sm_count, s);
    c->launches++;
    if (ssim_ok) {
        if (!c->legacy_ssim && ssim_strip_supported(height, width, d_a, frame_bytes, d_b, frame_bytes))
            JDS_CUDA(launch_ssim_strip(height, width, d_a, frame_bytes, d_b, frame_bytes, dm, 1, true,
                                       false, c->sm_count, s));
        else
            launch_ssim(true, height, width, d_a, frame_bytes, d_b, frame_bytes, dm, 1, s);
        c->launches++;
    }
    JDS_CUDA(cudaGetLastError());
    JDS_CUDA(cudaMemcpyAsync(c->h_metrics, dm, sizeof(DevMetrics), cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    const DevMetrics& d = *(const DevMetrics*)c->h_metrics;
    memset(m, 0, sizeof *m);
    m->sse_rgb = d.sse_rgb;
    m->sse_y = d.sse_y;
    for (int k = 0; k < 4; ++k) m->ssim_sum[k] = d.ssim_sum[k];
    m->ssim_count = ssim_ok ? (uint64_t)(height - 6) * (uint64_t)(width - 6) : 0;
    return JDS_OK;
}

// The band's own share of the metrics and outputs, enqueued behind the kernels of its extent
// (UnitJob::tail): squared errors over its rows, SSIM over its window centres, bit counts over its
// blocks, then the copies of its rows / blocks to the caller.
struct BandTail {
    const uint8_t* d_in;
    const uint8_t* d_recon;
    int W, sse_row0, sse_rows, ssim_row0, ssim_rows;
    const int16_t* src[3];
    size_t cnt[3];
    uint8_t* recon_rows;
    int16_t* coeffs_rows;
    cudaMemcpyKind kind;
};

static int band_tail(jds_ctx* c, void* arg, cudaStream_t s) {
    const BandTail& T = *(const BandTail*)arg;
    DevMetrics* dm = (DevMetrics*)c->metrics.p + 1;
    const size_t row_bytes = (size_t)T.W * 3;
    JDS_CUDA(cudaMemsetAsync(dm, 0, sizeof(DevMetrics), s));
    const size_t own = row_bytes * (size_t)T.sse_row0;
    launch_sse_u8(T.d_in + own, T.d_recon + own, (long long)T.sse_rows * T.W, dm, c->sm_count, s);
    c->launches++;
    if (T.ssim_rows) {
        const size_t off = row_bytes * (size_t)T.ssim_row0, fb = row_bytes * (size_t)T.ssim_rows;
        if (!c->legacy_ssim && ssim_strip_supported(T.ssim_rows, T.W, T.d_in + off, fb, T.d_recon + off, fb))
            JDS_CUDA(launch_ssim_strip(T.ssim_rows, T.W, T.d_in + off, fb, T.d_recon + off, fb, dm, 1, true,
                                       false, c->sm_count, s));
        else
            launch_ssim(true, T.ssim_rows, T.W, T.d_in + off, fb, T.d_recon + off, fb, dm, 1, s);
        c->launches++;
    }
    for (int k = 0; k < 3; ++k)
        if (T.cnt[k]) {
            launch_bitcount(T.src[k], T.cnt[k], dm, c->sm_count, s);
            c->launches++;
        }
    JDS_CUDA(cudaGetLastError());
    JDS_CUDA(cudaMemcpyAsync((DevMetrics*)c->h_metrics + 1, dm, sizeof(DevMetrics), cudaMemcpyDeviceToHost, s));
    if (T.recon_rows)
        JDS_CUDA(cudaMemcpyAsync(T.recon_rows, T.d_recon + own, row_bytes * (size_t)T.sse_rows, T.kind, s));
    if (T.coeffs_rows) {
        size_t at = 0;
        for (int k = 0; k < 3; ++k) {
            if (T.cnt[k]) JDS_CUDA(cudaMemcpyAsync(T.coeffs_rows + at, T.src[k], T.cnt[k] * 2, T.kind, s));
            at += T.cnt[k];
        }
    }
    return JDS_OK;
}

// Tile-band sharding of ONE frame: the rows [row0, row1) of a frame as one
// rank's share of engines/pipeline.py:17-167.  The band is run with a halo of whole MCU rows on
// each interior edge - everything a pixel of the band depends on lies inside it:
//   * 8x8 blocks / 16-row MCUs are independent (block_processor.py:19-48);
//   * the bilinear chroma upsample reads one chroma sample beyond the band (color_space.py:64-65)
//     and SSIM's 7x7 windows three reconstructed rows (utils/metrics.py:12-14): 16 rows cover both;
//   * the 3x3 prefilter (color_space.py:39-40) reads one row beyond each chroma block, so the
//     outermost halo MCU row of a prefiltered band is itself inexact - a second one is added.
// The partial metrics are sums over the band's own rows / window centres / blocks, so adding the
// partials of all bands (an all-reduce of a few dozen numbers) gives the frame's jds_metrics.
extern "C" int jds_roundtrip_band(jds_ctx* c, const jds_params* p, const uint8_t* rgb, int rgb_loc,
                                  int row0, int row1, uint8_t* recon_rows, int16_t* coeffs_rows,
                                  int out_loc, jds_metrics* m) {
    if (!c || !rgb || !m) return fail(JDS_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    const int H = p->height, W = p->width, sub = p->subsampling;
    if (row0 < 0 || row1 > H || row0 >= row1 || (row0 % 16) != 0 || ((row1 % 16) != 0 && row1 != H))
        return fail(JDS_ERR_INVALID, "band [%d, %d) of %d rows: bounds must be multiples of 16 (or the last row)",
                    row0, row1, H);
    if (p->outputs & (JDS_OUT_ERR_Y | JDS_OUT_ERR_RGB | JDS_OUT_HIST))
        return fail(JDS_ERR_UNSUPPORTED, "error maps and histogram are whole-frame outputs");
    const bool whole = row0 == 0 && row1 == H;
    if (sub == JDS_SUB_420 && (H & 1) && !whole)
        return fail(JDS_ERR_UNSUPPORTED, "4:2:0 with an odd height: cv2.resize's area taps depend on the "
                                         "whole height (engines/color_space.py:48-49), bands cannot reproduce them");
    Geom gf;
    if ((rc = make_geom(H, W, sub, &gf))) return rc;
    const int halo = (p->prefilter && sub != JDS_SUB_444) ? 32 : 16;
    const int e0 = row0 - halo > 0 ? row0 - halo : 0, e1 = row1 + halo < H ? row1 + halo : H;
    const int he = e1 - e0;
    jds_params pe = *p;
    pe.height = he;
    pe.outputs = JDS_OUT_RECON | JDS_OUT_COEFFS | JDS_OUT_PSNR;
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(he, W, sub, &J.g))) return rc;
    const Geom& g = J.g;
    JDS_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t row_bytes = (size_t)W * 3, ext_bytes = row_bytes * he;
    const size_t ncoef = 64ull * (size_t)(g.nblk_y + 2 * g.nblk_c);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const bool in_host = rgb_loc == JDS_HOST;
    const size_t o_recon = in_host ? up(ext_bytes) : 0, o_coef = o_recon + up(ext_bytes);
    if ((rc = ensure(c, c->band, o_coef + up(ncoef * 2)))) return rc;
    uint8_t* base = (uint8_t*)c->band.p;
    const uint8_t* d_in = rgb + row_bytes * e0;
    if (in_host) {
        JDS_CUDA(cudaMemcpyAsync(base, rgb + row_bytes * e0, ext_bytes, cudaMemcpyHostToDevice, s));
        d_in = base;
    }
    uint8_t* d_recon = base + o_recon;
    int16_t* d_coef = (int16_t*)(base + o_coef);
    // accumulator [1] (device and pinned host) is the band's; [0] is run_job's whole-extent one
    if ((rc = ensure(c, c->metrics, 2 * sizeof(DevMetrics)))) return rc;
    if ((rc = ensure_pinned(&c->h_metrics, &c->h_metrics_bytes, 2 * sizeof(DevMetrics)))) return rc;
    BandTail T;
    T.d_in = d_in;
    T.d_recon = d_recon;
    T.W = W;
    T.sse_row0 = row0 - e0;
    T.sse_rows = row1 - row0;
    // window centres of the frame are rows [3, H-3): this band's are [c0, c1)
    const int c0 = row0 > 3 ? row0 : 3, c1 = row1 < H - 3 ? row1 : H - 3;
    const bool want_ssim = (p->outputs & JDS_OUT_SSIM) != 0 && c1 > c0 && W >= 7;
    T.ssim_row0 = want_ssim ? c0 - 3 - e0 : 0;
    T.ssim_rows = want_ssim ? c1 - c0 + 6 : 0;
    // blocks of the band: block rows [row0/8, row1/8) of Y, [row0/8v, row1/8v) of Cb and Cr
    const int v = sub == JDS_SUB_420 ? 2 : 1;
    const int by0 = row0 / 8, by1 = row1 == H ? gf.nby_y : row1 / 8;
    const int cy0 = row0 / (8 * v), cy1 = row1 == H ? gf.nby_c : row1 / (8 * v);
    const size_t n_y = 64ull * (size_t)(by1 - by0) * gf.nbx_y, n_c = 64ull * (size_t)(cy1 - cy0) * gf.nbx_c;
    T.src[0] = d_coef + 64ull * (size_t)(by0 - e0 / 8) * g.nbx_y;
    T.src[1] = d_coef + 64ull * ((size_t)g.nblk_y + (size_t)(cy0 - e0 / (8 * v)) * g.nbx_c);
    T.src[2] = T.src[1] + 64ull * (size_t)g.nblk_c;
    T.cnt[0] = n_y;
    T.cnt[1] = T.cnt[2] = n_c;
    T.recon_rows = ((p->outputs & JDS_OUT_RECON) && recon_rows) ? recon_rows : nullptr;
    T.coeffs_rows = ((p->outputs & JDS_OUT_COEFFS) && coeffs_rows) ? coeffs_rows : nullptr;
    T.kind = out_loc == JDS_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    jds_metrics ext;
    J.p = &pe;
    J.units = 1;
    J.rgb = d_in;
    J.rgb_loc = JDS_DEVICE;
    J.recon = d_recon;
    J.coeffs = d_coef;
    J.out_loc = JDS_DEVICE;
    J.metrics = &ext;
    J.no_compare = true;
    J.tail = band_tail;
    J.tail_arg = &T;
    if ((rc = run_job(c, J))) return rc;          // kernels, tail, one synchronisation
    const DevMetrics& d = ((const DevMetrics*)c->h_metrics)[1];
    memset(m, 0, sizeof *m);
    m->sse_rgb = d.sse_rgb;
    m->sse_y = d.sse_y;
    for (int k = 0; k < 4; ++k) m->ssim_sum[k] = want_ssim ? d.ssim_sum[k] : 0.0;
    m->ssim_count = want_ssim ? (uint64_t)(c1 - c0) * (uint64_t)(W - 6) : 0;
    m->coeff_bits = d.coeff_bits;
    m->nnz = d.nnz;
    m->total_coeffs = n_y + 2 * n_c;
    m->luma_blocks = (uint64_t)(by1 - by0) * gf.nbx_y;
    m->gpu_ms = ext.gpu_ms;
    return JDS_OK;
}

// One arm of AliasingDemoWorker.run (gui/dialogs/aliasing_demo_dialog.py:98-166):
// _process_with_explicit_subsample(prefilter) - OpenCV float32 YCrCb, optional 5x5 blur,
// [::2, ::2], bilinear re-enlargement, back to uint8 RGB (jds_alias.cu) - then the hot path at
// 4:4:4 on that frame (the staged / fused kernels of jds_roundtrip), compute_metrics
// (RGB and OpenCV's integer luma) against the ORIGINAL frame, and _compute_difference.
extern "C" int jds_aliasing_demo(jds_ctx* c, const uint8_t* rgb, int rgb_loc, int height, int width,
                                 int quality, int prefilter, int precision, uint8_t* subsampled,
                                 uint8_t* recon, uint8_t* diff, int out_loc, jds_metrics* m_rgb,
                                 jds_metrics* m_luma) {
    if (!c || !rgb || !m_rgb || !m_luma) return fail(JDS_ERR_INVALID, "NULL argument");
    if (height < 8 || width < 8)
        return fail(JDS_ERR_INVALID, "aliasing demo needs frames of at least 8x8, got %dx%d", height, width);
    if (quality < 1 || quality > 100) return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", quality);
    if (precision != JDS_EXACT && precision != JDS_FAST) return fail(JDS_ERR_INVALID, "bad precision %d", precision);
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t n_px = (size_t)height * width, frame_bytes = n_px * 3;
    // scratch: float planes | subsampled frame | reconstruction | two luma3 frames | input copy
    const size_t fl_bytes = (alias_scratch_floats(height, width) * sizeof(float) + 255) & ~(size_t)255;
    const size_t fr = (frame_bytes + 255) & ~(size_t)255;
    int rc;
    if ((rc = ensure(c, c->alias, fl_bytes + 5 * fr))) return rc;
    char* base = (char*)c->alias.p;
    float* planes = (float*)base;
    uint8_t* d_sub = (uint8_t*)(base + fl_bytes);
    uint8_t* d_rec = d_sub + fr;
    uint8_t* d_la = d_rec + fr;
    uint8_t* d_lb = d_la + fr;
    uint8_t* d_in = d_lb + fr;
    cudaStream_t s = c->stream;
    const uint8_t* d_rgb = rgb;
    if (rgb_loc == JDS_HOST) {
        JDS_CUDA(cudaMemcpyAsync(d_in, rgb, frame_bytes, cudaMemcpyHostToDevice, s));
        d_rgb = d_in;
    }
    c->launches += (uint64_t)launch_alias_subsample(height, width, prefilter, d_rgb, planes, d_sub, s);
    JDS_CUDA(cudaGetLastError());

    // the hot path at 4:4:4, prefilter off (aliasing_demo_dialog.py:152-158)
    jds_params p;
    memset(&p, 0, sizeof p);
    p.height = height;
    p.width = width;
    p.quality = quality;
    p.subsampling = JDS_SUB_444;
    p.precision = precision;
    p.outputs = JDS_OUT_RECON | JDS_OUT_PSNR;
    jds_metrics unused;
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(height, width, JDS_SUB_444, &J.g))) return rc;
    J.p = &p;
    J.units = 1;
    J.rgb = d_sub;
    J.rgb_loc = JDS_DEVICE;
    J.recon = d_rec;
    J.out_loc = JDS_DEVICE;
    J.metrics = &unused;
    if ((rc = run_job(c, J))) return rc;

    // compute_metrics(original, reconstructed) (:69-83)
    if ((rc = compare_device(c, d_rgb, d_rec, height, width, m_rgb))) return rc;
    launch_alias_luma3(n_px, d_rgb, d_la, s);
    launch_alias_luma3(n_px, d_rec, d_lb, s);
    c->launches += 2;
    if ((rc = compare_device(c, d_la, d_lb, height, width, m_luma))) return rc;

    const cudaMemcpyKind out_kind = out_loc == JDS_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (diff) {
        launch_alias_diff(frame_bytes, d_rgb, d_rec, d_la, s);      // d_la is free again
        c->launches++;
        JDS_CUDA(cudaMemcpyAsync(diff, d_la, frame_bytes, out_kind, s));
    }
    if (subsampled) JDS_CUDA(cudaMemcpyAsync(subsampled, d_sub, frame_bytes, out_kind, s));
    if (recon) JDS_CUDA(cudaMemcpyAsync(recon, d_rec, frame_bytes, out_kind, s));
    JDS_CUDA(cudaGetLastError());
    JDS_CUDA(cudaStreamSynchronize(s));
    return JDS_OK;
}

// compute_metrics(original, reconstructed) of the aliasing demo on its own
// (gui/dialogs/aliasing_demo_dialog.py:69-83): RGB sums and OpenCV-integer-luma sums
extern "C" int jds_aliasing_metrics(jds_ctx* c, const uint8_t* a, const uint8_t* b, int loc,
                                    int height, int width, jds_metrics* m_rgb, jds_metrics* m_luma) {
    if (!c || !a || !b || !m_rgb || !m_luma) return fail(JDS_ERR_INVALID, "NULL argument");
    JDS_NO_PENDING(c);
    if (height < 1 || width < 1) return fail(JDS_ERR_INVALID, "bad frame size %dx%d", height, width);
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t n_px = (size_t)height * width, frame_bytes = n_px * 3;
    const size_t fr = (frame_bytes + 255) & ~(size_t)255;
    int rc;
    if ((rc = ensure(c, c->alias, 4 * fr))) return rc;
    uint8_t* base = (uint8_t*)c->alias.p;
    cudaStream_t s = c->stream;
    const uint8_t *d_a = a, *d_b = b;
    if (loc == JDS_HOST) {
        JDS_CUDA(cudaMemcpyAsync(base + 2 * fr, a, frame_bytes, cudaMemcpyHostToDevice, s));
        JDS_CUDA(cudaMemcpyAsync(base + 3 * fr, b, frame_bytes, cudaMemcpyHostToDevice, s));
        d_a = base + 2 * fr;
        d_b = base + 3 * fr;
    }
    if ((rc = compare_device(c, d_a, d_b, height, width, m_rgb))) return rc;
    launch_alias_luma3(n_px, d_a, base, s);
    launch_alias_luma3(n_px, d_b, base + fr, s);
    c->launches += 2;
    return compare_device(c, base, base + fr, height, width, m_luma);
}

// Exact size of the baseline-JPEG (T.81 Huffman, Annex K tables) scans of a coefficient array
// in the reference's order (engines/pipeline.py:56,99) - SURVEY 8f #4: the on-wire size the
// reference's estimate_bitrate_no_entropy (utils/metrics.py:51-92) only approximates.
extern "C" int jds_entropy_bits(jds_ctx* c, const int16_t* coeffs, int loc, int height, int width,
                                int subsampling, uint64_t scan_bits[3]) {
    if (!c || !coeffs || !scan_bits) return fail(JDS_ERR_INVALID, "NULL argument");
    JDS_NO_PENDING(c);
    Geom g;
    int rc = make_geom(height, width, subsampling, &g);
    if (rc) return rc;
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t ncoef = 64ull * (size_t)(g.nblk_y + 2 * g.nblk_c);
    if ((rc = ensure(c, c->metrics, sizeof(DevMetrics)))) return rc;
    if ((rc = ensure_pinned(&c->h_metrics, &c->h_metrics_bytes, sizeof(DevMetrics)))) return rc;
    cudaStream_t s = c->stream;
    const int16_t* d_c = coeffs;
    if (loc == JDS_HOST) {
        if ((rc = ensure(c, c->coeffs, ncoef * 2))) return rc;
        JDS_CUDA(cudaMemcpyAsync(c->coeffs.p, coeffs, ncoef * 2, cudaMemcpyHostToDevice, s));
        d_c = (const int16_t*)c->coeffs.p;
    }
    unsigned long long* d_bits = (unsigned long long*)c->metrics.p;
    JDS_CUDA(launch_entropy_bits(d_c, g.nblk_y, g.nblk_c, d_bits, s));
    c->launches++;
    JDS_CUDA(cudaMemcpyAsync(c->h_metrics, d_bits, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 3; ++k) scan_bits[k] = ((const unsigned long long*)c->h_metrics)[k];
    return JDS_OK;
}

// The entropy-coded BYTES of the same three scans, produced on the device (jds_entropy.cu).
// On success *d_scans points at the stuffed scans, back to back, in ctx scratch (valid until the
// next entropy call on this context) and the stream has been synchronised.
static int entropy_encode_device(jds_ctx* c, const int16_t* d_c, const Geom& g, const uint8_t** d_scans,
                                 uint64_t scan_bytes[3], uint64_t scan_bits[3]) {
    int rc;
    cudaStream_t s = c->stream;
    const EntropyGrid eg = make_entropy_grid(g.nblk_y, g.nblk_c);
    const size_t nblk = (size_t)(g.nblk_y + 2 * g.nblk_c);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    // [layout][bits per block][sum per CTA][offset per CTA]
    const size_t o_bits = up(sizeof(EntropyLayout)), o_part = o_bits + up(nblk * 4),
                 o_off = o_part + up((size_t)eg.ctas * 4), sizes_bytes = o_off + up((size_t)eg.ctas * 8);
    if ((rc = ensure(c, c->ent_sizes, sizes_bytes))) return rc;
    if ((rc = ensure_pinned(&c->h_metrics, &c->h_metrics_bytes, sizeof(DevMetrics) + sizeof(EntropyLayout))))
        return rc;
    uint8_t* base = (uint8_t*)c->ent_sizes.p;
    EntropyLayout* d_lay = (EntropyLayout*)base;
    uint32_t* d_blk = (uint32_t*)(base + o_bits);
    uint32_t* d_part = (uint32_t*)(base + o_part);
    unsigned long long* d_off = (unsigned long long*)(base + o_off);
    EntropyLayout* h_lay = (EntropyLayout*)c->h_metrics;
    JDS_CUDA(launch_entropy_sizes(d_c, eg, d_blk, d_part, d_off, d_lay, s));
    c->launches += 2;
    JDS_CUDA(cudaMemcpyAsync(h_lay, d_lay, sizeof(EntropyLayout), cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 3; ++k) scan_bits[k] = h_lay->bits[k];
    if (h_lay->invalid)
        return fail(JDS_ERR_INVALID, "coefficients outside the baseline JPEG code tables "
                                     "(DC difference beyond 11 bits or AC value beyond 10 bits)");
    // unstuffed bits: whole 4096-byte chunks; then [count per chunk][offset per chunk]
    const size_t ubytes = ((size_t)h_lay->total_ubytes + 4095) & ~(size_t)4095;
    const size_t chunks = ubytes / 4096;
    const size_t o_cnt = ubytes, o_coff = o_cnt + up(chunks * 4), bits_bytes = o_coff + up(chunks * 8);
    if ((rc = ensure(c, c->ent_bits, bits_bytes))) return rc;
    uint8_t* ub = (uint8_t*)c->ent_bits.p;
    uint32_t* d_ubuf = (uint32_t*)ub;
    uint32_t* d_cnt = (uint32_t*)(ub + o_cnt);
    unsigned long long* d_coff = (unsigned long long*)(ub + o_coff);
    JDS_CUDA(launch_entropy_pack(d_c, eg, d_blk, d_off, d_lay, d_ubuf, ubytes, s));
    JDS_CUDA(launch_stuff_sizes(d_ubuf, ubytes, d_lay, d_cnt, d_coff, s));
    c->launches += chunks ? 3 : 2;
    JDS_CUDA(cudaMemcpyAsync(h_lay, d_lay, sizeof(EntropyLayout), cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    uint64_t total = 0;
    for (int k = 0; k < 3; ++k) total += (scan_bytes[k] = h_lay->ubytes[k] + h_lay->ff[k]);
    if (total != h_lay->stuffed_bytes)
        return fail(JDS_ERR_CUDA, "entropy coder: stuffed size %llu != %llu", (unsigned long long)total,
                    (unsigned long long)h_lay->stuffed_bytes);
    if ((rc = ensure(c, c->ent_out, (size_t)total + 16))) return rc;
    JDS_CUDA(launch_stuff_scatter(d_ubuf, ubytes, d_lay, d_coff, (uint8_t*)c->ent_out.p, s));
    if (chunks) c->launches++;
    JDS_CUDA(cudaStreamSynchronize(s));
    *d_scans = (const uint8_t*)c->ent_out.p;
    return JDS_OK;
}

static int entropy_input(jds_ctx* c, const int16_t* coeffs, int loc, const Geom& g, const int16_t** d_c) {
    *d_c = coeffs;
    if (loc != JDS_HOST) return JDS_OK;
    const size_t ncoef = 64ull * (size_t)(g.nblk_y + 2 * g.nblk_c);
    int rc;
    if ((rc = ensure(c, c->coeffs, ncoef * 2))) return rc;
    JDS_CUDA(cudaMemcpyAsync(c->coeffs.p, coeffs, ncoef * 2, cudaMemcpyHostToDevice, c->stream));
    *d_c = (const int16_t*)c->coeffs.p;
    return JDS_OK;
}

extern "C" int jds_entropy_encode(jds_ctx* c, const int16_t* coeffs, int loc, int height, int width,
                                  int subsampling, uint8_t* out, int out_loc, uint64_t out_capacity,
                                  uint64_t scan_bytes[3], uint64_t scan_bits[3]) {
    if (!c || !coeffs || !scan_bytes || !scan_bits) return fail(JDS_ERR_INVALID, "NULL argument");
    JDS_NO_PENDING(c);
    Geom g;
    int rc = make_geom(height, width, subsampling, &g);
    if (rc) return rc;
    JDS_CUDA(cudaSetDevice(c->device));
    const int16_t* d_c;
    if ((rc = entropy_input(c, coeffs, loc, g, &d_c))) return rc;
    const uint8_t* d_scans;
    if ((rc = entropy_encode_device(c, d_c, g, &d_scans, scan_bytes, scan_bits))) return rc;
    const uint64_t total = scan_bytes[0] + scan_bytes[1] + scan_bytes[2];
    if (!out) return JDS_OK;
    if (out_capacity < total)
        return fail(JDS_ERR_CAPACITY, "output buffer of %llu bytes, the scans need %llu",
                    (unsigned long long)out_capacity, (unsigned long long)total);
    JDS_CUDA(cudaMemcpyAsync(out, d_scans, (size_t)total,
                             out_loc == JDS_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    JDS_CUDA(cudaStreamSynchronize(c->stream));
    return JDS_OK;
}

// A complete baseline JFIF file of the round trip's coefficients: what the reference would have
// to write to put its "compressed" result on disk (it never does; utils/metrics.py:57-61 says
// so).  Headers on the host, the three scans from the device coder above.
extern "C" int jds_jfif_encode(jds_ctx* c, const int16_t* coeffs, int loc, int height, int width,
                               int subsampling, const double qtable[64], uint8_t* out,
                               uint64_t out_capacity, uint64_t* out_bytes, uint64_t scan_bits[3]) {
    if (!c || !coeffs || !qtable || !out_bytes || !scan_bits) return fail(JDS_ERR_INVALID, "NULL argument");
    JDS_NO_PENDING(c);
    Geom g;
    int rc = make_geom(height, width, subsampling, &g);
    if (rc) return rc;
    if (height > 65535 || width > 65535) return fail(JDS_ERR_INVALID, "JPEG frames are at most 65535 x 65535");
    // JPEG sizes a subsampled component with ceil(), the pipeline with floor()
    // (engines/color_space.py:44-49): the two agree on even sizes only
    if (subsampling != JDS_SUB_444 && ((width & 1) || (subsampling == JDS_SUB_420 && (height & 1))))
        return fail(JDS_ERR_INVALID, "odd frame sizes: JPEG's component geometry differs from the pipeline's");
    uint8_t q8[64];
    for (int k = 0; k < 64; ++k) {
        if (!(qtable[k] >= 1.0 && qtable[k] <= 255.0) || qtable[k] != (double)(int)qtable[k])
            return fail(JDS_ERR_INVALID, "quantisation table entries must be integers in 1..255");
        q8[k] = (uint8_t)qtable[k];
    }
    JDS_CUDA(cudaSetDevice(c->device));
    const int16_t* d_c;
    if ((rc = entropy_input(c, coeffs, loc, g, &d_c))) return rc;
    const uint8_t* d_scans;
    uint64_t scan_bytes[3];
    if ((rc = entropy_encode_device(c, d_c, g, &d_scans, scan_bytes, scan_bits))) return rc;
    const size_t head = jfif_write_headers(nullptr, height, width, subsampling, q8);
    const uint64_t total = head + 3 * 10 + scan_bytes[0] + scan_bytes[1] + scan_bytes[2] + 2;
    *out_bytes = total;
    if (!out) return JDS_OK;
    if (out_capacity < total)
        return fail(JDS_ERR_CAPACITY, "output buffer of %llu bytes, the file needs %llu",
                    (unsigned long long)out_capacity, (unsigned long long)total);
    size_t at = jfif_write_headers(out, height, width, subsampling, q8);
    size_t from = 0;
    for (int k = 0; k < 3; ++k) {
        at += jfif_write_sos(out + at, k);
        JDS_CUDA(cudaMemcpyAsync(out + at, d_scans + from, (size_t)scan_bytes[k], cudaMemcpyDeviceToHost, c->stream));
        at += (size_t)scan_bytes[k];
        from += (size_t)scan_bytes[k];
    }
    JDS_CUDA(cudaStreamSynchronize(c->stream));
    out[at++] = 0xFF;
    out[at++] = 0xD9;
    return JDS_OK;
}

// GUI plot payload: the round trip plus, instead of the 25 MB coefficient
// array and the 66 MB fp64 error maps, what the reference's plots draw from them
// (gui/compression_tab.py:653-676 -> gui/widgets/mpl_canvas.py:81-130): the count of every
// coefficient value and the x10 clipped error map(s) as uint8.
extern "C" int jds_plot_payload(jds_ctx* c, const jds_params* p, const uint8_t* rgb, int rgb_loc,
                                uint8_t* recon, uint8_t* heat_y, uint8_t* heat_rgb,
                                int64_t* value_hist, int out_loc, jds_metrics* metrics) {
    if (!c || !rgb || !metrics) return fail(JDS_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    UnitJob J;
    memset(&J, 0, sizeof J);
    if ((rc = make_geom(p->height, p->width, p->subsampling, &J.g))) return rc;
    const Geom& g = J.g;
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t px = (size_t)g.H * g.W;
    const size_t frame_bytes = px * 3;
    const size_t ncoef = 64ull * (size_t)(g.nblk_y + 2 * g.nblk_c);
    const bool out_host = out_loc == JDS_HOST;
    // scratch layout: [err_y f64][err_rgb f64][coeffs i16][hist u64][heat_y u8][heat_rgb u8][recon u8]
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_ey = 0, o_ergb = o_ey + up(px * 8), o_coef = o_ergb + up(heat_rgb ? px * 8 : 0);
    const size_t o_hist = o_coef + up(value_hist ? ncoef * 2 : 0);
    const size_t o_hy = o_hist + up(VALUE_HIST_BINS * 8), o_hrgb = o_hy + up(px);
    const size_t o_recon = o_hrgb + up(px), total = o_recon + up(frame_bytes);
    if ((rc = ensure(c, c->payload, total))) return rc;
    char* base = (char*)c->payload.p;
    const uint8_t* d_rgb = rgb;
    if (rgb_loc == JDS_HOST) {
        if ((rc = ensure(c, c->in, frame_bytes))) return rc;
        JDS_CUDA(cudaMemcpyAsync(c->in.p, rgb, frame_bytes, cudaMemcpyHostToDevice, c->stream));
        d_rgb = (const uint8_t*)c->in.p;
    }
    jds_params q = *p;
    q.outputs = (p->outputs & (JDS_OUT_SSIM | JDS_OUT_PSNR | JDS_OUT_HIST)) | JDS_OUT_RECON;
    if (heat_y) q.outputs |= JDS_OUT_ERR_Y;
    if (heat_rgb) q.outputs |= JDS_OUT_ERR_RGB;
    if (value_hist) q.outputs |= JDS_OUT_COEFFS;
    J.p = &q;
    J.units = 1;
    J.rgb = d_rgb;
    J.rgb_loc = JDS_DEVICE;
    J.recon = (recon && !out_host) ? recon : (uint8_t*)(base + o_recon);
    J.coeffs = value_hist ? (int16_t*)(base + o_coef) : nullptr;
    J.err_y = heat_y ? (double*)(base + o_ey) : nullptr;
    J.err_rgb = heat_rgb ? (double*)(base + o_ergb) : nullptr;
    J.out_loc = JDS_DEVICE;
    J.metrics = metrics;
    if ((rc = run_job(c, J))) return rc;
    cudaStream_t s = c->stream;
    unsigned long long* d_hist = (unsigned long long*)(base + o_hist);
    if (value_hist) {
        JDS_CUDA(cudaMemsetAsync(d_hist, 0, VALUE_HIST_BINS * 8, s));
        launch_value_hist(J.coeffs, ncoef, d_hist, c->sm_count, s);
        c->launches++;
    }
    uint8_t* d_hy = (heat_y && !out_host) ? heat_y : (uint8_t*)(base + o_hy);
    uint8_t* d_hrgb = (heat_rgb && !out_host) ? heat_rgb : (uint8_t*)(base + o_hrgb);
    if (heat_y) { launch_heat_u8(J.err_y, d_hy, px, s); c->launches++; }
    if (heat_rgb) { launch_heat_u8(J.err_rgb, d_hrgb, px, s); c->launches++; }
    JDS_CUDA(cudaGetLastError());
    const cudaMemcpyKind kind = out_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (value_hist) JDS_CUDA(cudaMemcpyAsync(value_hist, d_hist, VALUE_HIST_BINS * 8, kind, s));
    if (out_host) {
        if (heat_y) JDS_CUDA(cudaMemcpyAsync(heat_y, d_hy, px, kind, s));
        if (heat_rgb) JDS_CUDA(cudaMemcpyAsync(heat_rgb, d_hrgb, px, kind, s));
        if (recon) JDS_CUDA(cudaMemcpyAsync(recon, J.recon, frame_bytes, kind, s));
    }
    JDS_CUDA(cudaStreamSynchronize(s));
    return JDS_OK;
}

// Preview downscale in front of the round trip (gui/compression_tab.py:532-552)
extern "C" int jds_preview_size(int height, int width, int target_w, int target_h, int* out_h,
                                int* out_w) {
    if (!out_h || !out_w) return fail(JDS_ERR_INVALID, "NULL argument");
    if (height < 1 || width < 1 || target_w < 1 || target_h < 1)
        return fail(JDS_ERR_INVALID, "bad preview geometry %dx%d -> %dx%d", height, width, target_h, target_w);
    if (width <= target_w && height <= target_h) {       // :541-543: small images are kept
        *out_h = height;
        *out_w = width;
        return JDS_OK;
    }
    // :545-547 in Python floats (IEEE double): scale = min(tw / w, th / h); int() truncates
    volatile double sx = (double)target_w / (double)width, sy = (double)target_h / (double)height;
    volatile double scale = sx < sy ? sx : sy;
    volatile double fw = (double)width * scale, fh = (double)height * scale;
    *out_w = (int)fw;
    *out_h = (int)fh;
    return JDS_OK;
}

extern "C" int jds_resize_area(jds_ctx* c, const uint8_t* rgb, int rgb_loc, int height, int width,
                               uint8_t* out, int out_h, int out_w, int out_loc) {
    if (!c || !rgb || !out) return fail(JDS_ERR_INVALID, "NULL argument");
    if (height < 1 || width < 1 || out_h < 1 || out_w < 1)
        return fail(JDS_ERR_INVALID, "bad resize geometry %dx%d -> %dx%d", height, width, out_h, out_w);
    if (out_h > height || out_w > width)
        return fail(JDS_ERR_UNSUPPORTED, "INTER_AREA enlargement (%dx%d -> %dx%d) is not part of the "
                    "preview path", height, width, out_h, out_w);
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t in_bytes = (size_t)height * width * 3, out_bytes = (size_t)out_h * out_w * 3;
    int rc;
    const uint8_t* d_in = rgb;
    uint8_t* d_out = out;
    cudaStream_t s = c->stream;
    if (rgb_loc == JDS_HOST) {
        if ((rc = ensure(c, c->in, in_bytes))) return rc;
        JDS_CUDA(cudaMemcpyAsync(c->in.p, rgb, in_bytes, cudaMemcpyHostToDevice, s));
        d_in = (const uint8_t*)c->in.p;
    }
    if (out_loc == JDS_HOST) {
        if ((rc = ensure(c, c->recon, out_bytes))) return rc;
        d_out = (uint8_t*)c->recon.p;
    }
    if (launch_resize_area_u8(height, width, out_h, out_w, d_in, d_out, s))
        return fail(JDS_ERR_INVALID, "bad resize geometry");
    c->launches++;
    JDS_CUDA(cudaGetLastError());
    if (out_loc == JDS_HOST)
        JDS_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    return JDS_OK;
}

extern "C" int jds_selected_block(jds_ctx* c, const jds_params* p, const uint8_t* rgb,
                                  int rgb_loc, int block_row, int block_col,
                                  double original[64], double shifted[64], double dct[64],
                                  int16_t quantized[64], double dequantized[64],
                                  double reconstructed[64], int* present) {
    if (!c || !rgb || !present || !original || !shifted || !dct || !quantized || !dequantized ||
        !reconstructed)
        return fail(JDS_ERR_INVALID, "NULL argument");
    JDS_NO_PENDING(c);
    int rc = check_params(p);
    if (rc) return rc;
    if (p->quality < 1 || p->quality > 100)
        return fail(JDS_ERR_INVALID, "Quality must be 1-100, got %d", p->quality);
    Geom g;
    if ((rc = make_geom(p->height, p->width, JDS_SUB_444, &g))) return rc;
    // engines/pipeline.py:134-137: target = row * blocks_per_row + col, valid iff in range
    const long long target = (long long)block_row * g.nbx_y + block_col;
    if (target < 0 || target >= g.nblk_y) {
        *present = 0;
        return JDS_OK;
    }
    const int by = (int)(target / g.nbx_y), bx = (int)(target % g.nbx_y);
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t frame_bytes = (size_t)g.H * g.W * 3;
    const size_t ob = selected_out_bytes();
    if ((rc = ensure(c, c->selected, ob))) return rc;
    if ((rc = ensure(c, c->tables, sizeof(QTables)))) return rc;
    if ((rc = ensure_pinned(&c->h_tables, &c->h_tables_bytes, sizeof(QTables)))) return rc;
    if (!c->h_selected) JDS_CUDA(cudaMallocHost(&c->h_selected, ob));
    const uint8_t* d_rgb = rgb;
    if (rgb_loc == JDS_HOST) {
        // only the 8 (reflect-mapped) rows of the block are needed, but the kernel
        // indexes the frame: stage the whole frame (a few MB, once per GUI click)
        if ((rc = ensure(c, c->in, frame_bytes))) return rc;
        JDS_CUDA(cudaMemcpyAsync(c->in.p, rgb, frame_bytes, cudaMemcpyHostToDevice, c->stream));
        d_rgb = (const uint8_t*)c->in.p;
    }
    fill_tables(p->quality, (QTables*)c->h_tables);
    JDS_CUDA(cudaMemcpyAsync(c->tables.p, c->h_tables, sizeof(QTables), cudaMemcpyHostToDevice,
                             c->stream));
    launch_selected_block(g, d_rgb, bx, by, (const QTables*)c->tables.p, c->selected.p, c->stream);
    c->launches++;
    JDS_CUDA(cudaGetLastError());
    JDS_CUDA(cudaMemcpyAsync(c->h_selected, c->selected.p, ob, cudaMemcpyDeviceToHost, c->stream));
    JDS_CUDA(cudaStreamSynchronize(c->stream));
    const double* h = (const double*)c->h_selected;
    memcpy(original, h, 64 * 8);
    memcpy(shifted, h + 64, 64 * 8);
    memcpy(dct, h + 128, 64 * 8);
    memcpy(dequantized, h + 192, 64 * 8);
    memcpy(reconstructed, h + 256, 64 * 8);
    memcpy(quantized, h + 320, 64 * 2);
    *present = 1;
    return JDS_OK;
}

// Stand-alone block operators (host buffers): engines/dct_engine.py:7-27 and
// engines/quantizer.py:22-29 on n_blocks 8x8 blocks, exact (reference) arithmetic.
extern "C" int jds_block_op(jds_ctx* c, int op, int64_t n_blocks, const double* in,
                            const int16_t* in_q, const double* qtable, double* out,
                            int16_t* out_q) {
    if (!c) return fail(JDS_ERR_INVALID, "ctx is NULL");
    if (n_blocks < 1) return fail(JDS_ERR_INVALID, "n_blocks must be >= 1");
    if (op < JDS_BLOCKOP_DCT2 || op > JDS_BLOCKOP_DEQUANTIZE)
        return fail(JDS_ERR_INVALID, "unknown block op %d", op);
    const bool needs_q = op == JDS_BLOCKOP_QUANTIZE || op == JDS_BLOCKOP_DEQUANTIZE;
    const bool in_is_q = op == JDS_BLOCKOP_DEQUANTIZE, out_is_q = op == JDS_BLOCKOP_QUANTIZE;
    if ((needs_q && !qtable) || (in_is_q ? !in_q : !in) || (out_is_q ? !out_q : !out))
        return fail(JDS_ERR_INVALID, "NULL buffer for block op %d", op);
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t n = (size_t)n_blocks * 64;
    // layout of the scratch: in (fp64 or int16) | out (fp64 or int16) | table
    int rc;
    if ((rc = ensure(c, c->planes, n * 16 + 512))) return rc;
    char* base = (char*)c->planes.p;
    double* d_in = (double*)base;
    double* d_out = (double*)(base + n * 8);
    double* d_tab = (double*)(base + n * 16);
    cudaStream_t s = c->stream;
    if (in_is_q) JDS_CUDA(cudaMemcpyAsync(d_in, in_q, n * 2, cudaMemcpyHostToDevice, s));
    else JDS_CUDA(cudaMemcpyAsync(d_in, in, n * 8, cudaMemcpyHostToDevice, s));
    if (needs_q) JDS_CUDA(cudaMemcpyAsync(d_tab, qtable, 512, cudaMemcpyHostToDevice, s));
    launch_block_ops(op, n_blocks, d_in, (const int16_t*)d_in, d_tab, d_out, (int16_t*)d_out, s);
    c->launches++;
    JDS_CUDA(cudaGetLastError());
    if (out_is_q) JDS_CUDA(cudaMemcpyAsync(out_q, d_out, n * 2, cudaMemcpyDeviceToHost, s));
    else JDS_CUDA(cudaMemcpyAsync(out, d_out, n * 8, cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    return JDS_OK;
}

// ------------------------------------------------------------------------------
// stand-alone stage operators (csrc/jds_ops.cu), host buffers for the fp64 planes
// ------------------------------------------------------------------------------
extern "C" int jds_color_convert(jds_ctx* c, int direction, int64_t n_pixels, const double* in,
                                 double* out) {
    if (!c || !in || !out) return fail(JDS_ERR_INVALID, "NULL argument");
    if (n_pixels < 1 || (direction != 0 && direction != 1))
        return fail(JDS_ERR_INVALID, "bad colour conversion request");
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n_pixels * 24;
    int rc;
    if ((rc = ensure(c, c->planes, 2 * bytes))) return rc;
    double* d_in = (double*)c->planes.p;
    double* d_out = d_in + (size_t)n_pixels * 3;
    cudaStream_t s = c->stream;
    JDS_CUDA(cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, s));
    launch_color_f64(direction, n_pixels, d_in, d_out, s);
    c->launches++;
    JDS_CUDA(cudaGetLastError());
    JDS_CUDA(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    return JDS_OK;
}

extern "C" int jds_subsample_plane(jds_ctx* c, const double* plane, int height, int width,
                                   int subsampling, int prefilter, double* out) {
    if (!c || !plane || !out) return fail(JDS_ERR_INVALID, "NULL argument");
    Geom g;
    int rc = make_geom(height, width, subsampling, &g);
    if (rc) return rc;
    if (subsampling != JDS_SUB_444 && prefilter && (height < 2 || width < 2))
        return fail(JDS_ERR_UNSUPPORTED, "prefilter on a %dx%d plane", height, width);
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t in_bytes = (size_t)height * width * 8, out_bytes = (size_t)g.hc * g.wc * 8;
    if (subsampling == JDS_SUB_444) {                 // engines/color_space.py:35-36: a copy
        memcpy(out, plane, in_bytes);
        return JDS_OK;
    }
    if ((rc = ensure(c, c->planes, in_bytes + out_bytes + 256))) return rc;
    double* d_in = (double*)c->planes.p;
    double* d_out = (double*)((char*)c->planes.p + ((in_bytes + 255) & ~(size_t)255));
    cudaStream_t s = c->stream;
    JDS_CUDA(cudaMemcpyAsync(d_in, plane, in_bytes, cudaMemcpyHostToDevice, s));
    launch_subsample_plane(d_in, height, width, g.hc, g.wc, subsampling, prefilter, d_out, s);
    c->launches++;
    JDS_CUDA(cudaGetLastError());
    JDS_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    return JDS_OK;
}

extern "C" int jds_upsample_plane(jds_ctx* c, const double* plane, int height, int width,
                                  double* out, int out_height, int out_width) {
    if (!c || !plane || !out) return fail(JDS_ERR_INVALID, "NULL argument");
    if (height < 1 || width < 1 || out_height < height || out_width < width)
        return fail(JDS_ERR_UNSUPPORTED, "upsample %dx%d -> %dx%d: only enlarging is restated",
                    height, width, out_height, out_width);
    if ((height < 2 && out_height != height) || (width < 2 && out_width != width))
        return fail(JDS_ERR_UNSUPPORTED, "upsampling a 1-sample axis is not restated");
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t in_bytes = (size_t)height * width * 8, out_bytes = (size_t)out_height * out_width * 8;
    int rc;
    if ((rc = ensure(c, c->planes, in_bytes + out_bytes + 256))) return rc;
    double* d_in = (double*)c->planes.p;
    double* d_out = (double*)((char*)c->planes.p + ((in_bytes + 255) & ~(size_t)255));
    cudaStream_t s = c->stream;
    JDS_CUDA(cudaMemcpyAsync(d_in, plane, in_bytes, cudaMemcpyHostToDevice, s));
    launch_upsample_plane(d_in, height, width, out_height, out_width, d_out, s);
    c->launches++;
    JDS_CUDA(cudaGetLastError());
    JDS_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    return JDS_OK;
}

extern "C" int jds_compare_images(jds_ctx* c, const uint8_t* a, const uint8_t* b, int loc,
                                  int height, int width, jds_metrics* m) {
    if (!c || !a || !b || !m) return fail(JDS_ERR_INVALID, "NULL argument");
    JDS_NO_PENDING(c);
    if (height < 1 || width < 1) return fail(JDS_ERR_INVALID, "bad frame size %dx%d", height, width);
    JDS_CUDA(cudaSetDevice(c->device));
    const size_t frame_bytes = (size_t)height * width * 3;
    int rc;
    if ((rc = ensure(c, c->metrics, sizeof(DevMetrics)))) return rc;
    if ((rc = ensure_pinned(&c->h_metrics, &c->h_metrics_bytes, sizeof(DevMetrics)))) return rc;
    const uint8_t *d_a = a, *d_b = b;
    cudaStream_t s = c->stream;
    if (loc == JDS_HOST) {
        if ((rc = ensure(c, c->in, frame_bytes))) return rc;
        if ((rc = ensure(c, c->recon, frame_bytes))) return rc;
        JDS_CUDA(cudaMemcpyAsync(c->in.p, a, frame_bytes, cudaMemcpyHostToDevice, s));
        JDS_CUDA(cudaMemcpyAsync(c->recon.p, b, frame_bytes, cudaMemcpyHostToDevice, s));
        d_a = (const uint8_t*)c->in.p;
        d_b = (const uint8_t*)c->recon.p;
    }
    return compare_device(c, d_a, d_b, height, width, m);
}

extern "C" int jds_bitrate_partials(jds_ctx* c, const int16_t* coeffs, int loc, uint64_t n,
                                    uint64_t* nnz, uint64_t* coeff_bits) {
    if (!c || !coeffs || !nnz || !coeff_bits) return fail(JDS_ERR_INVALID, "NULL argument");
    JDS_NO_PENDING(c);
    JDS_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure(c, c->metrics, sizeof(DevMetrics)))) return rc;
    if ((rc = ensure_pinned(&c->h_metrics, &c->h_metrics_bytes, sizeof(DevMetrics)))) return rc;
    cudaStream_t s = c->stream;
    const int16_t* d_c = coeffs;
    if (loc == JDS_HOST && n) {
        if ((rc = ensure(c, c->coeffs, (size_t)n * 2))) return rc;
        JDS_CUDA(cudaMemcpyAsync(c->coeffs.p, coeffs, (size_t)n * 2, cudaMemcpyHostToDevice, s));
        d_c = (const int16_t*)c->coeffs.p;
    }
    DevMetrics* dm = (DevMetrics*)c->metrics.p;
    JDS_CUDA(cudaMemsetAsync(dm, 0, sizeof(DevMetrics), s));
    if (n) {
        launch_bitcount(d_c, n, dm, c->sm_count, s);
        c->launches++;
    }
    JDS_CUDA(cudaGetLastError());
    JDS_CUDA(cudaMemcpyAsync(c->h_metrics, dm, sizeof(DevMetrics), cudaMemcpyDeviceToHost, s));
    JDS_CUDA(cudaStreamSynchronize(s));
    *nnz = ((const DevMetrics*)c->h_metrics)->nnz;
    *coeff_bits = ((const DevMetrics*)c->h_metrics)->coeff_bits;
    return JDS_OK;
}
