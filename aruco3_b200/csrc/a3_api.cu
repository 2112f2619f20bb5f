// C ABI of the detection path (include/aruco3_b200.h): detector handle, the batched pipeline and the stage probes.
//
// Pipeline of a3_detect_batch, per chunk of frames (frames are independent, src/aruco.rs:52-121 is a pure
// function of one image, so chunks — and GPUs — never exchange data):
//   H2D frames (skipped for device input) -> K1 (grey + 1-bit mask) -> D2H mask bits
//   -> host threads: border following + quad filters, one frame per task (host_quads.cpp)
//   -> H2D quads -> K2 (warp, otsu, bits, dictionary match) -> D2H decode records -> markers in candidate order.
// Two chunk slots are kept in flight so the host stage of chunk c overlaps the device stages of chunk c+1.
// There is no CPU fallback: without a CUDA device every compute entry point returns A3_ERR_CUDA.
#include <string.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "a3_internal.h"

namespace a3 {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
a3_status fail(a3_status s, const std::string &msg) { g_error = msg; return s; }
a3_status cuda_fail(cudaError_t e, const char *what) {
    g_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    cudaGetLastError();
    return A3_ERR_CUDA;
}

namespace {

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, n * sizeof(T), cudaHostAllocDefault);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_h2d = nullptr, ev_k1 = nullptr, ev_bits = nullptr, ev_k2a = nullptr, ev_k2b = nullptr;
    DevBuf<uint8_t> d_src, d_grey, d_mask, d_patches;
    DevBuf<uint32_t> d_bits, d_quads, d_qframe;
    DevBuf<a3_decode> d_dec;
    PinBuf<uint32_t> h_bits, h_quads, h_qframe;
    PinBuf<a3_decode> h_dec;
};

}  // namespace
}  // namespace a3

struct a3_detector {
    a3_config cfg;
    a3_dictionary dict;
    int device = 0;
    uint32_t host_threads = 1;
    uint32_t mark_size = 0;
    uint32_t max_taps = 0;
    a3::Slot slot[2];
    a3::K1Tuning k1_tuning{};
    bool has_tuning = false;
    uint32_t chunk_override = 0;
    a3::DevBuf<uint64_t> d_codes;
    a3::DevBuf<float> d_taps;
    a3::DevBuf<int> d_meta;
};

namespace a3 {
namespace {

a3_status check_config(const a3_config &c) {
    if (c.threshold_window == 0) return fail(A3_ERR_INVALID_ARGUMENT, "threshold_window must be > 0 (imageproc asserts block_radius > 0)");
    if (c.threshold_window > 16) return fail(A3_ERR_UNSUPPORTED, "threshold_window > 16 is not supported by the CUDA path");
    if (!(c.contour_simplification_epsilon > 0.0))
        return fail(A3_ERR_INVALID_ARGUMENT, "contour_simplification_epsilon must be > 0 (approximate_polygon_dp panics otherwise)");
    if (c.homography_sample_size == 0 || c.homography_sample_size > 256)
        return fail(A3_ERR_UNSUPPORTED, "homography_sample_size must be in 1..256");
    return A3_OK;
}

uint32_t bytes_per_pixel(a3_format f) { return f == A3_FMT_RGB8 ? 3 : (f == A3_FMT_RGBA8 ? 4 : 1); }

// frames per chunk: bounded device footprint, enough CTAs per launch
uint32_t chunk_frames(uint32_t n, uint32_t w, uint32_t h, uint32_t bpp) {
    const size_t per_frame = (size_t)w * h * (bpp + 2);
    size_t c = ((size_t)768 << 20) / (per_frame ? per_frame : 1);
    if (c < 1) c = 1;
    if (c > 64) c = 64;
    if (c > n) c = n;
    return (uint32_t)c;
}

// run fn(i) for i in [0, n) on `threads` host threads
template <typename F>
void parallel_for(uint32_t n, uint32_t threads, F fn) {
    if (threads <= 1 || n <= 1) {
        for (uint32_t i = 0; i < n; i++) fn(i);
        return;
    }
    if (threads > n) threads = n;
    std::atomic<uint32_t> next{0};
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (uint32_t t = 0; t < threads; t++)
        pool.emplace_back([&] {
            for (;;) {
                const uint32_t i = next.fetch_add(1);
                if (i >= n) break;
                fn(i);
            }
        });
    for (auto &th : pool) th.join();
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace
}  // namespace a3

using namespace a3;

extern "C" {

const char *a3_version(void) { return "aruco3_b200 0.1.0 (sm_100a)"; }
const char *a3_last_error(void) { return g_error.c_str(); }
const char *a3_status_string(a3_status s) {
    switch (s) {
        case A3_OK: return "ok";
        case A3_ERR_INVALID_ARGUMENT: return "invalid argument";
        case A3_ERR_UNKNOWN_DICTIONARY: return "unknown dictionary";
        case A3_ERR_CUDA: return "CUDA error";
        case A3_ERR_CAPACITY: return "output capacity too small";
        case A3_ERR_UNSUPPORTED: return "unsupported";
        case A3_ERR_OUT_OF_MEMORY: return "out of memory";
        default: return "unknown status";
    }
}
int32_t a3_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void a3_config_default(a3_config *c) {
    if (!c) return;
    c->threshold_window = 7;
    c->contour_simplification_epsilon = 0.05;
    c->min_side_length_factor = 0.2f;
    c->min_corner_separation_factor = 0.1f;
    c->homography_sample_size = 49;
    c->filter_high_bit_errors = 1;
}

a3_status a3_detector_create(const a3_config *cfg, const a3_dictionary *dict, int32_t device, a3_detector **out) {
    if (!cfg || !dict || !out || !dict->codes) return fail(A3_ERR_INVALID_ARGUMENT, "a3_detector_create: null argument");
    if (a3_status s = check_config(*cfg)) return s;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(A3_ERR_CUDA, "no CUDA device: aruco3_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(A3_ERR_INVALID_ARGUMENT, "a3_detector_create: bad device index");
    A3_CUDA(cudaSetDevice(device));
    a3_detector *d = new a3_detector();
    d->cfg = *cfg;
    d->dict = *dict;
    d->device = device;
    d->mark_size = mark_size_of(dict->num_bits);
    uint32_t hc = std::thread::hardware_concurrency();
    d->host_threads = hc ? (hc > 64 ? 64 : hc) : 1;
    ResizeTaps tp = make_resize_taps(cfg->homography_sample_size, d->mark_size);
    d->max_taps = tp.max_taps;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = d->d_codes.reserve(dict->n_codes ? dict->n_codes : 1);
    if (e == cudaSuccess && dict->n_codes) e = cudaMemcpy(d->d_codes.p, dict->codes, (size_t)dict->n_codes * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = d->d_taps.reserve(tp.weights.size());
    if (e == cudaSuccess) e = cudaMemcpy(d->d_taps.p, tp.weights.data(), tp.weights.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = d->d_meta.reserve(tp.meta.size());
    if (e == cudaSuccess) e = cudaMemcpy(d->d_meta.p, tp.meta.data(), tp.meta.size() * 4, cudaMemcpyHostToDevice);
    for (auto &s : d->slot) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        for (cudaEvent_t *ev : {&s.ev_start, &s.ev_h2d, &s.ev_k1, &s.ev_bits, &s.ev_k2a, &s.ev_k2b})
            if (e == cudaSuccess) e = cudaEventCreate(ev);
    }
    if (e != cudaSuccess) {
        a3_detector_destroy(d);
        return cuda_fail(e, "a3_detector_create");
    }
    *out = d;
    return A3_OK;
}

void a3_detector_destroy(a3_detector *d) {
    if (!d) return;
    cudaSetDevice(d->device);
    for (auto &s : d->slot) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        for (cudaEvent_t ev : {s.ev_start, s.ev_h2d, s.ev_k1, s.ev_bits, s.ev_k2a, s.ev_k2b})
            if (ev) cudaEventDestroy(ev);
        s.d_src.release(); s.d_grey.release(); s.d_mask.release(); s.d_patches.release();
        s.d_bits.release(); s.d_quads.release(); s.d_qframe.release(); s.d_dec.release();
        s.h_bits.release(); s.h_quads.release(); s.h_qframe.release(); s.h_dec.release();
    }
    d->d_codes.release(); d->d_taps.release(); d->d_meta.release();
    delete d;
}

a3_status a3_detector_set_host_threads(a3_detector *d, uint32_t threads) {
    if (!d || threads == 0) return fail(A3_ERR_INVALID_ARGUMENT, "a3_detector_set_host_threads: bad argument");
    d->host_threads = threads > 256 ? 256 : threads;
    return A3_OK;
}

a3_status a3_detector_set_k1_tuning(a3_detector *d, const a3_k1_tuning *t) {
    if (!d) return fail(A3_ERR_INVALID_ARGUMENT, "a3_detector_set_k1_tuning: null detector");
    d->has_tuning = t != nullptr;
    d->k1_tuning = K1Tuning{};
    d->chunk_override = 0;
    if (t) {
        if (t->tma_rows != 0 && t->tma_rows != 1 && t->tma_rows != 2 && t->tma_rows != 4) return fail(A3_ERR_INVALID_ARGUMENT, "tma_rows must be 0, 1, 2 or 4");
        if (t->tma_stages > 8) return fail(A3_ERR_INVALID_ARGUMENT, "tma_stages must be <= 8");
        d->k1_tuning.strip_cols = t->strip_cols; d->k1_tuning.seg_rows = t->seg_rows; d->k1_tuning.force_no_tma = (int)t->force_no_tma;
        d->k1_tuning.force_generic = (int)t->force_generic; d->k1_tuning.tma_rows = t->tma_rows; d->k1_tuning.tma_stages = t->tma_stages;
        d->chunk_override = t->chunk_frames;
    }
    return A3_OK;
}

a3_status a3_gray_threshold_batch(a3_detector *d, const void *frames, a3_format format, a3_mem_kind mem, uint32_t n,
                                  uint32_t w, uint32_t h, size_t pitch, size_t frame_stride, uint8_t *grey, uint8_t *mask,
                                  uint32_t *mask_bits, void *cuda_stream) {
    if (!d || !frames) return fail(A3_ERR_INVALID_ARGUMENT, "a3_gray_threshold_batch: null argument");
    if ((int)format < 0 || (int)format > 2) return fail(A3_ERR_INVALID_ARGUMENT, "bad format");
    if (n == 0 || w == 0 || h == 0) return A3_OK;
    const uint32_t bpp = bytes_per_pixel(format);
    if (pitch < (size_t)w * bpp || frame_stride < pitch * h) return fail(A3_ERR_INVALID_ARGUMENT, "pitch / frame_stride too small");
    A3_CUDA(cudaSetDevice(d->device));
    K1Params p;
    p.format = format; p.n = n; p.w = w; p.h = h; p.pitch = pitch; p.frame_stride = frame_stride; p.radius = d->cfg.threshold_window;
    if (mem == A3_MEM_DEVICE) {
        p.src = static_cast<const uint8_t *>(frames); p.grey = grey; p.mask = mask; p.bits = mask_bits;
        A3_CUDA(k1_gray_threshold(p, d->has_tuning ? &d->k1_tuning : nullptr, static_cast<cudaStream_t>(cuda_stream), nullptr));
        return A3_OK;
    }
    // host pointers: stage through slot 0, chunk by chunk, synchronously
    Slot &s = d->slot[0];
    const size_t px = (size_t)w * h, wpr = (w + 31) / 32;
    const uint32_t chunk = chunk_frames(n, w, h, bpp);
    for (uint32_t f0 = 0; f0 < n; f0 += chunk) {
        const uint32_t c = n - f0 < chunk ? n - f0 : chunk;
        A3_CUDA(s.d_src.reserve((size_t)c * frame_stride));
        A3_CUDA(cudaMemcpyAsync(s.d_src.p, static_cast<const uint8_t *>(frames) + (size_t)f0 * frame_stride, (size_t)c * frame_stride,
                                cudaMemcpyHostToDevice, s.stream));
        if (grey) A3_CUDA(s.d_grey.reserve(c * px));
        if (mask) A3_CUDA(s.d_mask.reserve(c * px));
        if (mask_bits) A3_CUDA(s.d_bits.reserve(c * wpr * h));
        p.src = s.d_src.p; p.n = c;
        p.grey = grey ? s.d_grey.p : nullptr; p.mask = mask ? s.d_mask.p : nullptr; p.bits = mask_bits ? s.d_bits.p : nullptr;
        A3_CUDA(k1_gray_threshold(p, d->has_tuning ? &d->k1_tuning : nullptr, s.stream, nullptr));
        if (grey) A3_CUDA(cudaMemcpyAsync(grey + f0 * px, s.d_grey.p, c * px, cudaMemcpyDeviceToHost, s.stream));
        if (mask) A3_CUDA(cudaMemcpyAsync(mask + f0 * px, s.d_mask.p, c * px, cudaMemcpyDeviceToHost, s.stream));
        if (mask_bits) A3_CUDA(cudaMemcpyAsync(mask_bits + f0 * wpr * h, s.d_bits.p, c * wpr * h * 4, cudaMemcpyDeviceToHost, s.stream));
        A3_CUDA(cudaStreamSynchronize(s.stream));
    }
    return A3_OK;
}

a3_status a3_quads_from_mask(const a3_config *cfg, const uint8_t *mask, uint32_t w, uint32_t h, uint32_t *quads,
                             uint32_t quad_capacity, uint32_t *n_quads, a3_stats *stats) {
    if (!cfg || !mask || !n_quads) return fail(A3_ERR_INVALID_ARGUMENT, "a3_quads_from_mask: null argument");
    if (a3_status s = check_config(*cfg)) return s;
    std::vector<uint32_t> bits, q;
    uint32_t wpr = 0;
    bits_from_mask(mask, w, h, bits, &wpr);
    QuadStats qs;
    quads_from_bits(bits.data(), wpr, w, h, *cfg, q, &qs);
    *n_quads = (uint32_t)(q.size() / 8);
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_frames = 1; stats->n_contours = qs.n_contours; stats->n_contour_points = qs.n_contour_points;
        stats->n_candidates_before_discard = qs.n_before_discard; stats->n_candidates = *n_quads;
    }
    if (*n_quads > quad_capacity || (!quads && *n_quads)) return fail(A3_ERR_CAPACITY, "a3_quads_from_mask: quad_capacity too small");
    if (*n_quads) memcpy(quads, q.data(), q.size() * 4);
    return A3_OK;
}

a3_status a3_decode_candidates(a3_detector *d, const uint8_t *grey, uint32_t n_frames, uint32_t w, uint32_t h,
                               const uint32_t *quads, const uint32_t *quad_frame, uint32_t n_quads, a3_decode *decodes,
                               uint8_t *patches) {
    if (!d || !grey || (!quads && n_quads) || (!decodes && n_quads)) return fail(A3_ERR_INVALID_ARGUMENT, "a3_decode_candidates: null argument");
    if (n_quads == 0) return A3_OK;
    for (uint32_t i = 0; quad_frame && i < n_quads; i++)
        if (quad_frame[i] >= n_frames) return fail(A3_ERR_INVALID_ARGUMENT, "a3_decode_candidates: quad_frame out of range");
    A3_CUDA(cudaSetDevice(d->device));
    Slot &s = d->slot[0];
    const size_t px = (size_t)w * h, np = (size_t)d->cfg.homography_sample_size * d->cfg.homography_sample_size;
    A3_CUDA(s.d_grey.reserve(px * n_frames));
    A3_CUDA(s.d_quads.reserve((size_t)n_quads * 8));
    A3_CUDA(s.d_qframe.reserve(n_quads));
    A3_CUDA(s.d_dec.reserve(n_quads));
    if (patches) A3_CUDA(s.d_patches.reserve(n_quads * np));
    A3_CUDA(cudaMemcpyAsync(s.d_grey.p, grey, px * n_frames, cudaMemcpyHostToDevice, s.stream));
    A3_CUDA(cudaMemcpyAsync(s.d_quads.p, quads, (size_t)n_quads * 32, cudaMemcpyHostToDevice, s.stream));
    if (quad_frame) A3_CUDA(cudaMemcpyAsync(s.d_qframe.p, quad_frame, (size_t)n_quads * 4, cudaMemcpyHostToDevice, s.stream));
    K2Params p;
    p.grey = s.d_grey.p; p.w = w; p.h = h; p.quads = s.d_quads.p; p.quad_frame = quad_frame ? s.d_qframe.p : nullptr;
    p.n_quads = n_quads; p.patch_size = d->cfg.homography_sample_size; p.mark_size = d->mark_size;
    p.codes = d->d_codes.p; p.n_codes = d->dict.n_codes; p.tau = d->dict.tau; p.filter_high_bit_errors = d->cfg.filter_high_bit_errors;
    p.resize_w = d->d_taps.p; p.resize_meta = d->d_meta.p; p.decodes = s.d_dec.p; p.patches = patches ? s.d_patches.p : nullptr;
    A3_CUDA(k2_decode(p, s.stream));
    A3_CUDA(cudaMemcpyAsync(decodes, s.d_dec.p, (size_t)n_quads * sizeof(a3_decode), cudaMemcpyDeviceToHost, s.stream));
    if (patches) A3_CUDA(cudaMemcpyAsync(patches, s.d_patches.p, n_quads * np, cudaMemcpyDeviceToHost, s.stream));
    A3_CUDA(cudaStreamSynchronize(s.stream));
    return A3_OK;
}

a3_status a3_detect_batch(a3_detector *d, const void *frames, a3_format format, a3_mem_kind mem, uint32_t n, uint32_t w,
                          uint32_t h, size_t pitch, size_t frame_stride, a3_marker *markers, uint32_t marker_capacity,
                          uint32_t *n_markers, a3_outputs *outs, a3_stats *stats) {
    if (!d || !frames || !n_markers) return fail(A3_ERR_INVALID_ARGUMENT, "a3_detect_batch: null argument");
    if ((int)format < 0 || (int)format > 2) return fail(A3_ERR_INVALID_ARGUMENT, "bad format");
    *n_markers = 0;
    if (outs) outs->n_candidates = 0;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n == 0 || w == 0 || h == 0) {
        if (outs && outs->frame_marker_offsets) memset(outs->frame_marker_offsets, 0, ((size_t)n + 1) * 4);
        return A3_OK;
    }
    const uint32_t bpp = bytes_per_pixel(format);
    if (pitch < (size_t)w * bpp || frame_stride < pitch * h) return fail(A3_ERR_INVALID_ARGUMENT, "pitch / frame_stride too small");
    A3_CUDA(cudaSetDevice(d->device));
    const double t_begin = now_ms();
    const size_t px = (size_t)w * h, wpr = (w + 31) / 32, np = (size_t)d->cfg.homography_sample_size * d->cfg.homography_sample_size;
    uint32_t chunk = chunk_frames(n, w, h, bpp);
    if (d->chunk_override) chunk = d->chunk_override < n ? d->chunk_override : n;
    const uint32_t nchunks = (n + chunk - 1) / chunk;
    const bool want_mask = outs && outs->mask, want_grey = outs && outs->grey, want_patches = outs && outs->homographies;
    a3_stats st;
    memset(&st, 0, sizeof(st));
    st.n_frames = n;
    st.host_threads = d->host_threads;

    auto issue = [&](uint32_t c) -> a3_status {  // device front end of chunk c
        Slot &s = d->slot[c & 1];
        const uint32_t f0 = c * chunk, cn = n - f0 < chunk ? n - f0 : chunk;
        const uint8_t *src = static_cast<const uint8_t *>(frames) + (size_t)f0 * frame_stride;
        A3_CUDA(cudaEventRecord(s.ev_start, s.stream));
        if (mem == A3_MEM_HOST) {
            A3_CUDA(s.d_src.reserve((size_t)cn * frame_stride));
            A3_CUDA(cudaMemcpyAsync(s.d_src.p, src, (size_t)cn * frame_stride, cudaMemcpyHostToDevice, s.stream));
            src = s.d_src.p;
        }
        A3_CUDA(cudaEventRecord(s.ev_h2d, s.stream));
        A3_CUDA(s.d_grey.reserve(cn * px));
        A3_CUDA(s.d_bits.reserve(cn * wpr * h));
        A3_CUDA(s.h_bits.reserve(cn * wpr * h));
        if (want_mask) A3_CUDA(s.d_mask.reserve(cn * px));
        K1Params p;
        p.src = src; p.format = format; p.n = cn; p.w = w; p.h = h; p.pitch = pitch; p.frame_stride = frame_stride;
        p.grey = s.d_grey.p; p.mask = want_mask ? s.d_mask.p : nullptr; p.bits = s.d_bits.p; p.radius = d->cfg.threshold_window;
        A3_CUDA(k1_gray_threshold(p, d->has_tuning ? &d->k1_tuning : nullptr, s.stream, nullptr));
        st.pixel_kernel_launches++;
        A3_CUDA(cudaEventRecord(s.ev_k1, s.stream));
        A3_CUDA(cudaMemcpyAsync(s.h_bits.p, s.d_bits.p, cn * wpr * h * 4, cudaMemcpyDeviceToHost, s.stream));
        if (want_grey) A3_CUDA(cudaMemcpyAsync(outs->grey + f0 * px, s.d_grey.p, cn * px, cudaMemcpyDeviceToHost, s.stream));
        if (want_mask) A3_CUDA(cudaMemcpyAsync(outs->mask + f0 * px, s.d_mask.p, cn * px, cudaMemcpyDeviceToHost, s.stream));
        A3_CUDA(cudaEventRecord(s.ev_bits, s.stream));
        return A3_OK;
    };

    uint32_t total_markers = 0, total_cands = 0;
    bool overflow = false;
    if (a3_status s0 = issue(0)) return s0;
    std::vector<std::vector<uint32_t>> frame_quads;
    std::vector<QuadStats> frame_stats;
    for (uint32_t c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks)
            if (a3_status s1 = issue(c + 1)) return s1;
        Slot &s = d->slot[c & 1];
        const uint32_t f0 = c * chunk, cn = n - f0 < chunk ? n - f0 : chunk;
        A3_CUDA(cudaEventSynchronize(s.ev_bits));
        float ms = 0;
        cudaEventElapsedTime(&ms, s.ev_start, s.ev_h2d); st.ms_h2d += ms;
        cudaEventElapsedTime(&ms, s.ev_h2d, s.ev_k1); st.ms_pixel_kernel += ms;
        cudaEventElapsedTime(&ms, s.ev_k1, s.ev_bits); st.ms_mask_d2h += ms;
        // ---- host stage: one frame per task ----
        const double th0 = now_ms();
        frame_quads.assign(cn, {});
        frame_stats.assign(cn, QuadStats());
        parallel_for(cn, d->host_threads, [&](uint32_t i) {
            quads_from_bits(s.h_bits.p + (size_t)i * wpr * h, (uint32_t)wpr, w, h, d->cfg, frame_quads[i], &frame_stats[i]);
        });
        uint32_t nq = 0;
        for (uint32_t i = 0; i < cn; i++) {
            nq += (uint32_t)(frame_quads[i].size() / 8);
            st.n_contours += frame_stats[i].n_contours;
            st.n_contour_points += frame_stats[i].n_contour_points;
            st.n_candidates_before_discard += frame_stats[i].n_before_discard;
        }
        st.ms_host_quads += now_ms() - th0;
        st.n_candidates += nq;
        // ---- decode ----
        if (nq) {
            A3_CUDA(s.h_quads.reserve((size_t)nq * 8));
            A3_CUDA(s.h_qframe.reserve(nq));
            A3_CUDA(s.h_dec.reserve(nq));
            A3_CUDA(s.d_quads.reserve((size_t)nq * 8));
            A3_CUDA(s.d_qframe.reserve(nq));
            A3_CUDA(s.d_dec.reserve(nq));
            if (want_patches) A3_CUDA(s.d_patches.reserve(nq * np));
            uint32_t k = 0;
            for (uint32_t i = 0; i < cn; i++) {
                const uint32_t m = (uint32_t)(frame_quads[i].size() / 8);
                if (m) memcpy(s.h_quads.p + (size_t)k * 8, frame_quads[i].data(), (size_t)m * 32);
                for (uint32_t j = 0; j < m; j++) s.h_qframe.p[k + j] = i;
                k += m;
            }
            A3_CUDA(cudaMemcpyAsync(s.d_quads.p, s.h_quads.p, (size_t)nq * 32, cudaMemcpyHostToDevice, s.stream));
            A3_CUDA(cudaMemcpyAsync(s.d_qframe.p, s.h_qframe.p, (size_t)nq * 4, cudaMemcpyHostToDevice, s.stream));
            K2Params p;
            p.grey = s.d_grey.p; p.w = w; p.h = h; p.quads = s.d_quads.p; p.quad_frame = s.d_qframe.p; p.n_quads = nq;
            p.patch_size = d->cfg.homography_sample_size; p.mark_size = d->mark_size; p.codes = d->d_codes.p;
            p.n_codes = d->dict.n_codes; p.tau = d->dict.tau; p.filter_high_bit_errors = d->cfg.filter_high_bit_errors;
            p.resize_w = d->d_taps.p; p.resize_meta = d->d_meta.p; p.decodes = s.d_dec.p;
            p.patches = want_patches ? s.d_patches.p : nullptr;
            A3_CUDA(cudaEventRecord(s.ev_k2a, s.stream));
            A3_CUDA(k2_decode(p, s.stream));
            st.decode_kernel_launches++;
            A3_CUDA(cudaEventRecord(s.ev_k2b, s.stream));
            A3_CUDA(cudaMemcpyAsync(s.h_dec.p, s.d_dec.p, (size_t)nq * sizeof(a3_decode), cudaMemcpyDeviceToHost, s.stream));
            if (want_patches && total_cands < outs->cand_capacity) {
                const uint32_t room = outs->cand_capacity - total_cands, m = nq < room ? nq : room;
                A3_CUDA(cudaMemcpyAsync(outs->homographies + (size_t)total_cands * np, s.d_patches.p, (size_t)m * np,
                                        cudaMemcpyDeviceToHost, s.stream));
            }
            A3_CUDA(cudaStreamSynchronize(s.stream));
            cudaEventElapsedTime(&ms, s.ev_k2a, s.ev_k2b); st.ms_decode_kernel += ms;
        }
        // ---- markers, in candidate order (src/aruco.rs:75-113) ----
        uint32_t k = 0;
        for (uint32_t i = 0; i < cn; i++) {
            if (outs && outs->frame_marker_offsets) outs->frame_marker_offsets[f0 + i] = total_markers;
            const uint32_t m = (uint32_t)(frame_quads[i].size() / 8);
            for (uint32_t j = 0; j < m; j++, k++) {
                const a3_decode &dc = s.h_dec.p[k];
                const uint32_t *q = &frame_quads[i][(size_t)j * 8];
                if (outs && total_cands < outs->cand_capacity) {
                    if (outs->candidates) memcpy(outs->candidates + (size_t)total_cands * 8, q, 32);
                    if (outs->candidate_frame) outs->candidate_frame[total_cands] = f0 + i;
                    if (outs->decodes) outs->decodes[total_cands] = dc;
                } else if (outs && (outs->candidates || outs->decodes || outs->homographies)) {
                    overflow = true;
                }
                total_cands++;
                if (!dc.accepted) continue;
                if (markers && total_markers < marker_capacity) {
                    a3_marker &mk = markers[total_markers];
                    memset(&mk, 0, sizeof(mk));
                    mk.id = dc.id;
                    mk.code = dc.codes[dc.rotation & 3];
                    mk.frame = f0 + i;
                    mk.candidate = j;
                    mk.hamming_distance = dc.hamming_distance;
                    mk.rotation = dc.rotation;
                    for (uint32_t cidx = 0; cidx < 4; cidx++) {  // corners.rotate_left(min_rotation)
                        const uint32_t sidx = (cidx + dc.rotation) & 3;
                        mk.corners[2 * cidx] = q[2 * sidx];
                        mk.corners[2 * cidx + 1] = q[2 * sidx + 1];
                    }
                } else {
                    overflow = true;
                }
                total_markers++;
            }
        }
        if (want_grey || want_mask) A3_CUDA(cudaStreamSynchronize(s.stream));
    }
    if (outs && outs->frame_marker_offsets) outs->frame_marker_offsets[n] = total_markers;
    if (outs) outs->n_candidates = total_cands;
    *n_markers = total_markers;
    st.n_markers = total_markers;
    st.ms_total = now_ms() - t_begin;
    if (stats) *stats = st;
    if (overflow) return fail(A3_ERR_CAPACITY, "a3_detect_batch: output capacity too small (counts are valid)");
    return A3_OK;
}

}  // extern "C"
