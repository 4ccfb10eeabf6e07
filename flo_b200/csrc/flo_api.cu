// flo_api.cu -- host side of the C ABI declared in include/flo_b200.h.
//
// Mirrors the reference's Encoder (libflo/src/lossless/encoder.rs:9-45): a pure
// function (samples, sample_rate, channels, bit_depth, level, metadata) -> bytes.
// All arithmetic of the path runs in the kernels of flo_kernels.cu; this file only
// lays the batch out (frame table, offsets that do not depend on the data), moves
// buffers and launches.  There is no CPU implementation of any stage here.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

#include "flo_internal.h"

using namespace flo;

static thread_local char g_err[512] = "";
static void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char *flo_last_error(void) { return g_err; }
extern "C" const char *flo_version(void) { return "flo_b200 0.1 (sm_100a)"; }

#define CK(expr)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (expr);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            set_err("CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, \
                    cudaGetErrorString(e_));                                               \
            return e_ == cudaErrorMemoryAllocation ? FLO_ERR_NOMEM : FLO_ERR_CUDA;         \
        }                                                                                  \
    } while (0)

extern "C" int flo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void *flo_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void flo_host_free(void *p) { if (p) cudaFreeHost(p); }

namespace {

struct DevBuf {                       // grow-only device arena
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return FLO_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) { set_err("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); p = nullptr; return FLO_ERR_NOMEM; }
        cap = want;
        return FLO_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostBuf {                      // grow-only pinned host arena
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return FLO_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        if (cudaMallocHost(&p, bytes + bytes / 8 + 256) != cudaSuccess) { cudaGetLastError(); set_err("cudaMallocHost(%zu) failed", bytes); p = nullptr; return FLO_ERR_NOMEM; }
        cap = bytes + bytes / 8 + 256;
        return FLO_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// The library switches to the context's device for the duration of a call and puts the caller's current
// device back afterwards (a process that also runs other CUDA code keeps its own current device).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != dev) err = cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
constexpr int MAX_WAVES = 16;

// Host threads that copy pageable caller memory into the pinned staging ring, a slice each (a single thread
// copies at ~10 GB/s, a fifth of what the PCIe link takes).
class CopyPool {
  public:
    explicit CopyPool(int workers) : n_(workers) {
        for (int i = 0; i < n_; i++) th_.emplace_back([this, i] { worker(i); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; gen_++; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    // dst[0, bytes) = src[0, bytes), split over the workers and the calling thread
    void copy(uint8_t *dst, const uint8_t *src, size_t bytes) {
        const size_t parts = (size_t)n_ + 1;
        const size_t piece = align_up((bytes + parts - 1) / parts, 4096);
        if (n_ == 0 || bytes < (1u << 20)) { memcpy(dst, src, bytes); return; }
        { std::lock_guard<std::mutex> lk(m_); dst_ = dst; src_ = src; bytes_ = bytes; piece_ = piece; pending_ = n_; gen_++; }
        cv_.notify_all();
        slice(n_);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
  private:
    static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
    void slice(int i) {
        const size_t lo = std::min(bytes_, piece_ * (size_t)i), hi = std::min(bytes_, piece_ * (size_t)(i + 1));
        if (hi > lo) memcpy(dst_ + lo, src_ + lo, hi - lo);
    }
    void worker(int i) {
        unsigned long long seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> lk(m_); cv_.wait(lk, [&] { return gen_ != seen; }); seen = gen_; if (stop_) return; }
            slice(i);
            { std::lock_guard<std::mutex> lk(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    uint8_t *dst_ = nullptr; const uint8_t *src_ = nullptr; size_t bytes_ = 0, piece_ = 0;
    int pending_ = 0; unsigned long long gen_ = 0; bool stop_ = false;
};
constexpr int STAGE_SLOTS = 4;
constexpr size_t STAGE_SLOT_BYTES = 8u << 20;

inline int copy_workers() {
    cpu_set_t set;
    int cores = 1;
    if (sched_getaffinity(0, sizeof set, &set) == 0) cores = CPU_COUNT(&set);
    if (const char *e = getenv("FLO_B200_COPY_THREADS")) return std::max(0, atoi(e) - 1);
    return std::max(0, std::min(11, cores - 1));          // 12 copying threads: 31.5 ms per 1.27 GB of pageable f32 (8: 37.5 ms, 16: 31.8 ms)
}

// Pinned output blocks.  Large results are copied device -> host straight into one page-locked block
// and handed out as pointers into it (no second host copy); flo_free() on the last pointer of a
// block returns it to a small process-wide pool for the next call.
struct OutBlock {
    uint8_t *base = nullptr;
    size_t cap = 0;
    size_t refs = 0;
};
std::mutex g_blocks_mu;
std::vector<OutBlock *> g_live;       // handed out, refs > 0
std::vector<OutBlock *> g_pool;       // free, reusable
constexpr size_t POOL_MAX_BLOCKS = 4;
constexpr size_t SMALL_OUTPUT = 1u << 20;

OutBlock *take_block(size_t bytes) {
    std::lock_guard<std::mutex> lk(g_blocks_mu);
    size_t best = g_pool.size();
    for (size_t i = 0; i < g_pool.size(); i++)
        if (g_pool[i]->cap >= bytes && (best == g_pool.size() || g_pool[i]->cap < g_pool[best]->cap)) best = i;
    if (best < g_pool.size()) {
        OutBlock *b = g_pool[best];
        g_pool.erase(g_pool.begin() + best);
        return b;
    }
    OutBlock *b = new (std::nothrow) OutBlock();
    if (!b) return nullptr;
    const size_t cap = bytes + bytes / 8 + 4096;
    if (cudaMallocHost((void **)&b->base, cap) != cudaSuccess) { cudaGetLastError(); delete b; return nullptr; }
    b->cap = cap;
    return b;
}
void publish_block(OutBlock *b, size_t refs) {
    std::lock_guard<std::mutex> lk(g_blocks_mu);
    b->refs = refs;
    g_live.push_back(b);
}
void recycle_block_locked(OutBlock *b) {
    if (g_pool.size() < POOL_MAX_BLOCKS) { g_pool.push_back(b); return; }
    cudaFreeHost(b->base);
    delete b;
}
void drop_block(OutBlock *b) {          // never published (error path)
    std::lock_guard<std::mutex> lk(g_blocks_mu);
    recycle_block_locked(b);
}

}  // namespace

extern "C" void flo_free(void *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_blocks_mu);
        for (size_t i = 0; i < g_live.size(); i++) {
            OutBlock *b = g_live[i];
            if ((uint8_t *)p >= b->base && (uint8_t *)p < b->base + b->cap) {
                if (--b->refs == 0) {
                    g_live.erase(g_live.begin() + i);
                    recycle_block_locked(b);
                }
                return;
            }
        }
    }
    free(p);
}

struct flo_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    std::mutex mu;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;          // copy streams of the pipelined host entry
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_in[MAX_WAVES] = {}, ev_k[MAX_WAVES] = {}, ev_fin = nullptr;
    DevBuf in, out, meta, tracks, frames, ctrl, fexcl, fsize, foff, plane, cres, report;
    DevBuf dec_frames, dec_units, dec_base, dec_ctl;      // decoder scratch
    DevBuf conv;                                           // f32 samples of the U8 / S32 ingest pre-pass
    DevBuf peaks;                                          // waveform peaks (analysis metadata)
    DevBuf defer;                                          // per-CTA scratch of frames packed before their offset is known
    DevBuf crc_tab, fcrc;                                  // CRC tables for the encode kernel; raw CRC of every frame
    uint64_t counters[24] = {0};      // [0..7] analysis counters, [8..23] per-phase SM clock sums
    HostBuf h_small, h_out;
    HostBuf stage;                                         // pinned ring for pageable caller buffers
    cudaEvent_t ev_slot[STAGE_SLOTS] = {};
    bool slot_used[STAGE_SLOTS] = {};
    unsigned stage_next = 0;
    CopyPool *pool = nullptr;
    bool report_on = false;
    size_t persist_max = 0, window_max = 0;          // L2 persisting carve-out limits of the device
    size_t persist_prev = (size_t)-1;                // the device's persisting-L2 limit before this context raised it
    void *l2_win_ptr = nullptr; size_t l2_win_bytes = 0; cudaStream_t l2_win_stream = nullptr;
    uint32_t report_frames = 0;
    float ms[6] = {0, 0, 0, 0, 0, 0};
    uint32_t launches = 0;
};

extern "C" int flo_ctx_create(int device, flo_ctx **out) {
    if (!out) { set_err("flo_ctx_create: out is NULL"); return FLO_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_err("no CUDA device available (%s); flo_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return FLO_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_err("device %d out of range (0..%d)", device, n - 1); return FLO_ERR_ARG; }
    DeviceGuard dg(device);
    CK(dg.err);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) { set_err("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return FLO_ERR_CUDA; }
    flo_ctx *c = new (std::nothrow) flo_ctx();
    if (!c) { set_err("out of host memory"); return FLO_ERR_NOMEM; }
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->persist_max = (size_t)std::max(0, prop.persistingL2CacheMaxSize);
    c->window_max = (size_t)std::max(0, prop.accessPolicyMaxWindowSize);
    if (c->persist_max) {
        size_t prev = 0;
        if (cudaDeviceGetLimit(&prev, cudaLimitPersistingL2CacheSize) == cudaSuccess) c->persist_prev = prev;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->persist_max);
        cudaGetLastError();
    }
    // The context's own stream is a BLOCKING stream: it orders itself against the legacy default stream, so
    // device inputs produced there (e.g. by torch's default stream) are complete before the kernels read them.
    // Inputs produced on any other stream: pass that stream with flo_ctx_set_stream (see flo_b200.h).
    cudaError_t e2 = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamDefault);
    if (e2 != cudaSuccess) { set_err("cudaStreamCreate: %s", cudaGetErrorString(e2)); delete c; return FLO_ERR_CUDA; }
    c->stream = c->own_stream;
    for (auto &ev : c->ev) cudaEventCreate(&ev);
    cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    for (auto &ev : c->ev_in) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : c->ev_k) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : c->ev_slot) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_fin, cudaEventDisableTiming);
    upload_crc_tables();
    {
        static uint32_t tab[CRC_TAB_WORDS];
        crc_tables_host(tab);
        if (c->crc_tab.reserve(sizeof tab) || cudaMemcpy(c->crc_tab.p, tab, sizeof tab, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_err("CRC table upload failed"); flo_ctx_destroy(c); return FLO_ERR_CUDA;
        }
    }
    for (int th : {512, 256, 128}) {
        e2 = encode_variant(th).configure(c->smem_optin / encode_variant(th).ctas_per_sm - (encode_variant(th).ctas_per_sm > 1 ? 1024 : 0));
        if (e2 != cudaSuccess) { set_err("cudaFuncSetAttribute(max dynamic smem): %s", cudaGetErrorString(e2)); flo_ctx_destroy(c); return FLO_ERR_CUDA; }
    }
    e2 = cudaDeviceSynchronize();
    if (e2 != cudaSuccess) { set_err("device init: %s", cudaGetErrorString(e2)); flo_ctx_destroy(c); return FLO_ERR_CUDA; }
    *out = c;
    return FLO_OK;
}

extern "C" void flo_ctx_destroy(flo_ctx *c) {
    if (!c) return;
    DeviceGuard dg(c->device);
    cudaDeviceSynchronize();
    if (c->l2_win_ptr) {
        // take the persisting window off the stream it was installed on and give the L2 set-aside back
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        if (c->l2_win_stream == c->own_stream || c->l2_win_stream == c->stream) cudaStreamSetAttribute(c->l2_win_stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
    }
    if (c->persist_prev != (size_t)-1) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, c->persist_prev); cudaGetLastError(); }
    for (DevBuf *b : {&c->in, &c->out, &c->meta, &c->tracks, &c->frames, &c->ctrl, &c->fexcl, &c->fsize,
                      &c->foff, &c->plane, &c->cres, &c->report, &c->crc_tab, &c->fcrc, &c->defer, &c->peaks, &c->dec_frames, &c->dec_units, &c->dec_base, &c->dec_ctl, &c->conv})
        b->release();
    c->h_small.release();
    c->h_out.release();
    c->stage.release();
    delete c->pool;
    for (auto &ev : c->ev_slot) if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->ev_in) if (ev) cudaEventDestroy(ev);
    for (auto &ev : c->ev_k) if (ev) cudaEventDestroy(ev);
    if (c->ev_fin) cudaEventDestroy(c->ev_fin);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" int flo_ctx_set_stream(flo_ctx *c, void *cuda_stream) {
    if (!c) { set_err("ctx is NULL"); return FLO_ERR_ARG; }
    std::lock_guard<std::mutex> lk(c->mu);
    cudaStream_t ns = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    if (ns != c->stream && c->l2_win_ptr && c->l2_win_stream == c->stream) {
        // the persisting access window installed on the stream we leave must not outlive our use of it
        DeviceGuard dg(c->device);
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof av);
        cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
        c->l2_win_ptr = nullptr; c->l2_win_bytes = 0; c->l2_win_stream = nullptr;
    }
    c->stream = ns;
    return FLO_OK;
}

extern "C" int flo_ctx_last_timing(flo_ctx *c, float ms[6], uint32_t *launches) {
    if (!c) { set_err("ctx is NULL"); return FLO_ERR_ARG; }
    std::lock_guard<std::mutex> lk(c->mu);
    if (ms) memcpy(ms, c->ms, sizeof c->ms);
    if (launches) *launches = c->launches;
    return FLO_OK;
}

extern "C" int flo_ctx_last_counters(flo_ctx *c, uint64_t out[24]) {
    if (!c || !out) { set_err("bad argument"); return FLO_ERR_ARG; }
    std::lock_guard<std::mutex> lk(c->mu);
    memcpy(out, c->counters, sizeof c->counters);
    return FLO_OK;
}

extern "C" int flo_ctx_enable_report(flo_ctx *c, int enable) {
    if (!c) { set_err("ctx is NULL"); return FLO_ERR_ARG; }
    std::lock_guard<std::mutex> lk(c->mu);
    c->report_on = enable != 0;
    return FLO_OK;
}

extern "C" int flo_ctx_read_report(flo_ctx *c, uint32_t frame, uint32_t channel, flo_cand_report out[14]) {
    if (!c || !out) { set_err("bad argument"); return FLO_ERR_ARG; }
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->report.p || frame >= c->report_frames || channel >= (uint32_t)REPORT_CH) { set_err("no report for frame %u channel %u", frame, channel); return FLO_ERR_ARG; }
    DeviceGuard dg(c->device);
    CK(dg.err);
    const flo_cand_report *src = (const flo_cand_report *)c->report.p + ((size_t)frame * REPORT_CH + channel) * NCAND;
    CK(cudaMemcpy(out, src, sizeof(flo_cand_report) * NCAND, cudaMemcpyDeviceToHost));
    return FLO_OK;
}

// ---- batch layout -----------------------------------------------------------------
namespace {

struct Layout {
    std::vector<TrackDev> tr;
    uint64_t n_frames = 0;
    uint64_t n_segs = 0;
    uint64_t out_bound = 0;
    uint64_t meta_total = 0;
    uint64_t in_bytes = 0;            // host-input arena bytes
    std::vector<uint64_t> in_off;     // byte offset of each track in the input arena
    uint64_t max_plane_elems = 0;     // C * stride of the largest frame
    uint64_t max_group_bytes = 0;     // bytes of one interleaved sample frame of the widest track with <= 2 channels (staged ingest)
};

// worst-case bytes of one frame: 6 + C * (4 + 1 + 48 + 3 + 2 n)   (types.rs:242-267, raw payload bound)
inline uint64_t frame_bound(uint64_t cl0, uint64_t C) { return 6 + C * (4 + 52 + 2 * cl0); }

inline size_t fmt_size(int format) { return format == FLO_FMT_PCM16 ? 2 : format == FLO_FMT_U8 ? 1 : 4; }

int make_layout(const flo_track *tracks, size_t n_tracks, int format, Layout &L) {
    const size_t esz = fmt_size(format);
    L.tr.resize(n_tracks);
    L.in_off.resize(n_tracks);
    uint64_t stat = 0, frames = 0, segs = 0, bound = 0, meta = 0, inb = 0;
    for (size_t t = 0; t < n_tracks; t++) {
        const flo_track &k = tracks[t];
        if (k.channels == 0) { set_err("track %zu: channels == 0 (the reference panics: division by zero, encoder.rs:48)", t); return FLO_ERR_ARG; }
        if (k.sample_rate == 0) { set_err("track %zu: sample_rate == 0 (the reference panics: division by zero, encoder.rs:50)", t); return FLO_ERR_ARG; }
        if (k.n_interleaved && !k.samples) { set_err("track %zu: samples is NULL", t); return FLO_ERR_ARG; }
        if (k.meta_len && !k.meta) { set_err("track %zu: meta is NULL", t); return FLO_ERR_ARG; }
        const uint64_t C = k.channels, sr = k.sample_rate;
        if (sr * C > (1ull << 28)) { set_err("track %zu: sample_rate * channels too large for one frame", t); return FLO_ERR_ARG; }
        const uint64_t total = k.n_interleaved / C;                   // encoder.rs:48
        const uint64_t nf = (total + sr - 1) / sr;                    // encoder.rs:50
        if (frames + nf > 0xFFFFFFF0ull) { set_err("batch has too many frames"); return FLO_ERR_ARG; }
        TrackDev &d = L.tr[t];
        memset(&d, 0, sizeof d);
        d.n_inter = k.n_interleaved;
        d.static_off = stat;
        d.meta_off = meta;
        d.meta_len = k.meta_len;
        d.sample_rate = k.sample_rate;
        d.channels = k.channels;
        d.bit_depth = k.bit_depth;
        d.first_frame = (uint32_t)frames;
        d.n_frames = (uint32_t)nf;
        d.first_seg = (uint32_t)segs;
        uint64_t dbound = 0;
        if (nf) {
            const uint64_t full = frame_bound(sr, C);
            const uint64_t last_len = k.n_interleaved - (nf - 1) * sr * C;
            const uint64_t last_cl0 = std::min<uint64_t>((last_len + C - 1) / C, sr + 1);
            dbound = (nf - 1) * full + frame_bound(last_cl0, C);
            const uint64_t cl0 = nf > 1 ? sr : last_cl0;
            const uint64_t stride = (cl0 + 15) & ~15ull;
            L.max_plane_elems = std::max(L.max_plane_elems, C * stride);
            if (C <= 2) L.max_group_bytes = std::max<uint64_t>(L.max_group_bytes, C * (format == FLO_FMT_PCM16 ? 2 : 4));
        }
        const uint64_t fixed = FILE_HDR + 4 + 20 * nf + k.meta_len;
        stat += fixed;
        bound += fixed + dbound;
        frames += nf;
        segs += (dbound + CRC_SEG - 1) / CRC_SEG;
        if (segs > 0xFFFFFFF0ull) { set_err("batch too large"); return FLO_ERR_ARG; }
        meta += k.meta_len;
        L.in_off[t] = inb;
        inb += align_up(k.n_interleaved * esz, 256);
    }
    L.n_frames = frames; L.n_segs = segs; L.out_bound = bound + 64; L.meta_total = meta; L.in_bytes = inb;
    return FLO_OK;
}

}  // namespace

extern "C" size_t flo_output_bound(const flo_track *tracks, size_t n_tracks) {
    Layout L;
    if (make_layout(tracks, n_tracks, FLO_FMT_F32, L) != FLO_OK) return 0;
    return (size_t)L.out_bound;
}

// Host -> device copy of caller memory.  Page-locked sources go straight to the copy engine; pageable ones
// (what a Rust &[f32] is) are staged through a ring of pinned slots filled by the copy threads, so the DMA of
// one slot overlaps the host copy of the next instead of the driver's single-threaded bounce buffer.
static cudaError_t h2d_from_caller(flo_ctx *c, void *dst, const void *src, size_t bytes, bool pinned, cudaStream_t st) {
    if (pinned || bytes < (256u << 10) || getenv("FLO_B200_NO_STAGING")) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    if (!c->stage.p) {
        if (c->stage.reserve(STAGE_SLOTS * STAGE_SLOT_BYTES)) return cudaErrorMemoryAllocation;
        if (!c->pool) c->pool = new CopyPool(copy_workers());
    }
    for (size_t off = 0; off < bytes; off += STAGE_SLOT_BYTES) {
        const size_t len = std::min(STAGE_SLOT_BYTES, bytes - off);
        const unsigned slot = c->stage_next++ % STAGE_SLOTS;
        uint8_t *sp = (uint8_t *)c->stage.p + slot * STAGE_SLOT_BYTES;
        cudaError_t e = cudaSuccess;
        if (c->slot_used[slot]) e = cudaEventSynchronize(c->ev_slot[slot]);       // the DMA that last read this slot
        if (e != cudaSuccess) return e;
        c->pool->copy(sp, (const uint8_t *)src + off, len);
        e = cudaMemcpyAsync((uint8_t *)dst + off, sp, len, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev_slot[slot], st);
        if (e != cudaSuccess) return e;
        c->slot_used[slot] = true;
    }
    return cudaSuccess;
}
static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// ---- the batch pass ----------------------------------------------------------------
// `sink` (host entry only): when non-null and the batch is large, the images are copied to a pinned host block
// while later waves are still being uploaded and encoded; *sink receives the block (its data is complete on return).
static int encode_batch_impl(flo_ctx *c, const flo_track *tracks, size_t n_tracks, int format, uint8_t level,
                             bool host_inputs, void *d_out, size_t d_out_cap, uint64_t *offsets, uint64_t *lens,
                             OutBlock **sink = nullptr) {
    if (format != FLO_FMT_F32 && format != FLO_FMT_PCM16 && format != FLO_FMT_U8 && format != FLO_FMT_S32) { set_err("unknown sample format %d", format); return FLO_ERR_ARG; }
    const bool convert = format == FLO_FMT_U8 || format == FLO_FMT_S32;     // pre-pass to f32, then the f32 path
    if (n_tracks == 0) return FLO_OK;
    if (level > 9) level = 9;                                          // with_compression, encoder.rs:26-29
    DeviceGuard dg(c->device);
    CK(dg.err);
    Layout L;
    int rc = make_layout(tracks, n_tracks, format, L);
    if (rc) return rc;
    const size_t esz = fmt_size(format);
    cudaStream_t st = c->stream;
    std::vector<uint64_t> conv_off(convert ? n_tracks : 0);
    if (convert) {
        uint64_t cb = 0;
        for (size_t t = 0; t < n_tracks; t++) { conv_off[t] = cb; cb += align_up(tracks[t].n_interleaved * 4, 256); }
        if ((rc = c->conv.reserve(std::max<uint64_t>(cb, 1)))) return rc;
    }

    uint8_t *out = (uint8_t *)d_out;
    if (!out) {
        if ((rc = c->out.reserve(L.out_bound))) return rc;
        out = (uint8_t *)c->out.p;
    } else {
        if (d_out_cap < L.out_bound) { set_err("d_out_capacity %zu < flo_output_bound %llu", d_out_cap, (unsigned long long)L.out_bound); return FLO_ERR_ARG; }
        if ((uintptr_t)out & 15) { set_err("d_out must be 16-byte aligned"); return FLO_ERR_ARG; }
    }
    const uint32_t NF = (uint32_t)L.n_frames, NSEG = (uint32_t)L.n_segs, NTR = (uint32_t)n_tracks;
    if ((rc = c->tracks.reserve(sizeof(TrackDev) * n_tracks))) return rc;
    if ((rc = c->frames.reserve(sizeof(uint2) * std::max<uint64_t>(NF, 1)))) return rc;
    const size_t ctrl_bytes = 8ull * NF + 256 + 4ull * n_tracks + 4ull * MAX_WAVES;       // status words, ticket, err, counters, phase clocks, track CRCs
    if ((rc = c->ctrl.reserve(ctrl_bytes))) return rc;
    if ((rc = c->fexcl.reserve(8ull * std::max<uint64_t>(NF, 1)))) return rc;
    if ((rc = c->fsize.reserve(4ull * std::max<uint64_t>(NF, 1)))) return rc;
    if ((rc = c->fcrc.reserve(4ull * std::max<uint64_t>(NF, 1)))) return rc;
    if ((rc = c->foff.reserve(16ull * n_tracks))) return rc;
    if ((rc = c->meta.reserve(std::max<uint64_t>(L.meta_total, 1)))) return rc;
    if ((rc = c->h_small.reserve(sizeof(TrackDev) * n_tracks + L.meta_total + 16ull * n_tracks + 256))) return rc;

    // Kernel variant: several small CTAs per SM when the largest frame's 16-bit planes fit the smaller share of
    // shared memory (4 x 128 threads up to ~16 kHz content, 2 x 256 for mono up to 44.1 kHz); larger frames
    // (44.1/48 kHz stereo and up) keep their planes in an L2-persisting scratch and also run as 2 x 256 threads:
    // two independent CTAs per SM overlap each other's barriers and serial sections, which measured 11 % faster
    // than one 512-thread CTA with the frame in shared memory (3.30 vs 3.69 ms per hour of 44.1 kHz stereo).
    // FLO_B200_VARIANT=512|256|128 forces a variant (tests, comparisons).
    const size_t plane_bytes = L.max_plane_elems * 2;
    // Shared memory of a CTA: its fixed state, a work area (ingest stages of the bulk async copies, then the
    // packer's 16 KB staging ring) and, when they fit, the sample planes.  A stage is one step of all threads.
    auto work_min = [&](const EncodeVariant &v) { return (size_t)v.threads * 64; };   // one 2 KB ring per warp
    auto ctas_of = [&](const EncodeVariant &v) { return level < 4 ? v.ctas_fixed : v.ctas_per_sm; };   // levels 0-3: the instantiation without LPC
    auto share_of = [&](const EncodeVariant &v) { return ctas_of(v) > 1 ? c->smem_optin / ctas_of(v) - 1024 : c->smem_optin; };
    auto stage_of = [&](const EncodeVariant &v) { return (size_t)v.threads * (v.threads >= 512 ? 4 : 8) * (size_t)std::max<uint64_t>(L.max_group_bytes, 2); };
    auto fits = [&](const EncodeVariant &v) { return v.static_smem() + plane_bytes + std::max(work_min(v), 2 * stage_of(v)) <= share_of(v); };
    const EncodeVariant *var = &encode_variant(512);
    if (fits(encode_variant(128))) var = &encode_variant(128);
    else if (fits(encode_variant(256))) var = &encode_variant(256);
    else if (host_inputs && fits(encode_variant(512))) var = &encode_variant(512);   // PCIe-bound entry: the shared-memory variant leaves HBM/L2 to the copy engines (25.3 vs 26.8 ms e2e)
    else var = &encode_variant(256);                                       // planes in L2 scratch
    if (const char *e = getenv("FLO_B200_VARIANT")) var = &encode_variant(atoi(e));
    const size_t smem_static = var->static_smem();
    const bool planes_in_smem = fits(*var);
    size_t work = std::max(work_min(*var), 4 * stage_of(*var));
    work = std::min(work, (share_of(*var) - smem_static - (planes_in_smem ? plane_bytes : 0)) & ~(size_t)127);
    if (const char *e = getenv("FLO_B200_WORK_KB")) work = std::max(work_min(*var), (size_t)atoi(e) * 1024);
    size_t dyn, plane_cap;
    if (planes_in_smem) { dyn = share_of(*var); plane_cap = dyn - smem_static - work; }
    else { dyn = smem_static + work; plane_cap = 0; }                     // global (L2) planes
    const int ctas_per_sm = ctas_of(*var);
    if (getenv("FLO_B200_DEBUG_OCC"))
        fprintf(stderr, "flo_b200: variant %d x %d, dyn smem %zu, occupancy %d\n", var->threads, ctas_per_sm, dyn, var->occupancy(dyn));
    int grid = (int)std::min<uint64_t>(std::max<uint64_t>(NF, 1), (uint64_t)c->sm_count * ctas_per_sm);
    if (const char *e = getenv("FLO_B200_GRID")) grid = std::max(1, std::min(grid, atoi(e)));   // experiments
    if ((rc = c->cres.reserve(sizeof(ChanResult) * 256ull * grid))) return rc;
    // Frames of up to ~48 KB are packed into a per-CTA scratch (L2-resident: at most 592 x 48 KB) and copied to
    // their place once the offset is known; larger frames wait for their offset and are packed in place.
    // FLO_B200_DEFER_KB overrides the limit (0: every frame is packed in place).
    size_t defer_bytes = 48u << 10;
    if (const char *e = getenv("FLO_B200_DEFER_KB")) defer_bytes = (size_t)std::max(0, atoi(e)) << 10;
    if (defer_bytes && (rc = c->defer.reserve(defer_bytes * (size_t)grid))) return rc;
    uint64_t plane_elems = 0;
    if (L.max_plane_elems * 2 > plane_cap) {
        plane_elems = align_up(L.max_plane_elems, 64);
        if ((rc = c->plane.reserve(plane_elems * 2 * grid))) return rc;
        // The planes are written once and re-read by every pass: pin them in L2 (persisting access window) so the
        // streaming input and output do not push them out to DRAM.
        if (c->persist_max > 0 && !getenv("FLO_B200_NO_L2_PERSIST")) {
            const size_t bytes = plane_elems * 2 * grid;
            if (c->l2_win_ptr != c->plane.p || c->l2_win_bytes != bytes || c->l2_win_stream != st) {
                const size_t setaside = std::min<size_t>(c->persist_max, bytes);
                cudaStreamAttrValue av;
                memset(&av, 0, sizeof av);
                av.accessPolicyWindow.base_ptr = c->plane.p;
                av.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, c->window_max);
                av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setaside / (double)std::max<size_t>(av.accessPolicyWindow.num_bytes, 1));
                av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
                cudaGetLastError();
                c->l2_win_ptr = c->plane.p; c->l2_win_bytes = bytes; c->l2_win_stream = st;
            }
        }
    }
    if (c->report_on) {
        if ((rc = c->report.reserve(sizeof(flo_cand_report) * NCAND * REPORT_CH * std::max<uint64_t>(NF, 1)))) return rc;
        c->report_frames = NF;
    }

    // ---- wave plan: the host entry uploads and encodes the batch in waves of whole frames so that the
    // H2D copy of wave w+1 overlaps the encode kernel of wave w (frames are independent; the look-back
    // status words carry the running byte offset from one launch to the next).
    int n_waves = 1;
    if (host_inputs && !convert && NF >= 2u * (uint32_t)grid * 2u) n_waves = (int)std::min<uint32_t>(MAX_WAVES, NF / (3u * (uint32_t)grid));
    if (n_waves < 1) n_waves = 1;
    std::vector<uint32_t> wave_end(n_waves);
    for (int w = 0; w < n_waves; w++) wave_end[w] = (uint32_t)((uint64_t)NF * (w + 1) / n_waves);
    // The call ends one wave's encode and one wave's D2H after the last H2D: the last wave is halved again and again
    // (down to about half a frame per CTA) so that this tail is short -- its pieces are still uploaded back to back.
    // 1-hour stereo stream, 4 waves of 900 frames: tail 2.3 ms -> 0.7 ms of a 26.8 ms call.
    if (n_waves > 1 && !getenv("FLO_B200_NO_TAPER")) {
        uint32_t begin = wave_end[n_waves - 2], rest = NF - begin;
        wave_end.pop_back();
        while (rest > (uint32_t)grid / 2u + 1u && (int)wave_end.size() < MAX_WAVES - 1) {
            const uint32_t piece = (rest + 1u) / 2u;
            begin += piece; rest -= piece;
            wave_end.push_back(begin);
        }
        if (rest) wave_end.push_back(NF);
        n_waves = (int)wave_end.size();
    }
    const bool piped = n_waves > 1;
    cudaStream_t s_in = piped ? c->s_in : st;

    // small tables through pinned staging
    CK(cudaEventRecord(c->ev[0], st));
    if (host_inputs) {
        if ((rc = c->in.reserve(std::max<uint64_t>(L.in_bytes, 1)))) return rc;
        for (size_t t = 0; t < n_tracks; t++) L.tr[t].samples = (const uint8_t *)c->in.p + L.in_off[t];
    } else {
        for (size_t t = 0; t < n_tracks; t++) {
            L.tr[t].samples = tracks[t].samples;
            if (((uintptr_t)tracks[t].samples & (esz - 1)) != 0) { set_err("track %zu: device samples pointer not aligned to the sample size", t); return FLO_ERR_ARG; }
        }
    }
    std::vector<const void *> raw_src(convert ? n_tracks : 0);
    if (convert)
        for (size_t t = 0; t < n_tracks; t++) { raw_src[t] = L.tr[t].samples; L.tr[t].samples = (const uint8_t *)c->conv.p + conv_off[t]; }
    uint8_t *hs = (uint8_t *)c->h_small.p;
    memcpy(hs, L.tr.data(), sizeof(TrackDev) * n_tracks);
    uint8_t *hmeta = hs + sizeof(TrackDev) * n_tracks;
    for (size_t t = 0; t < n_tracks; t++)
        if (tracks[t].meta_len) memcpy(hmeta + L.tr[t].meta_off, tracks[t].meta, tracks[t].meta_len);
    CK(cudaMemcpyAsync(c->tracks.p, hs, sizeof(TrackDev) * n_tracks, cudaMemcpyHostToDevice, st));
    if (L.meta_total) CK(cudaMemcpyAsync(c->meta.p, hmeta, L.meta_total, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(c->ctrl.p, 0, ctrl_bytes, st));

    uint32_t launches = 0;
    EncodeParams ep;
    memset(&ep, 0, sizeof ep);
    ep.tracks = (const TrackDev *)c->tracks.p;
    ep.frames = (const uint2 *)c->frames.p;
    ep.n_frames = NF;
    ep.format = convert ? FLO_FMT_F32 : format;
    ep.level = level;
    ep.out = out;
    ep.status = (unsigned long long *)c->ctrl.p;
    uint32_t *ctl = (uint32_t *)((uint8_t *)c->ctrl.p + 8ull * NF);
    ep.err = ctl + 1;
    ep.counters = ctl + 2;
    ep.phase_cycles = (unsigned long long *)(ctl + 16);
    uint32_t *wave_ticket = ctl + 48 + n_tracks;
    ep.frame_excl = (unsigned long long *)c->fexcl.p;
    ep.frame_size = (uint32_t *)c->fsize.p;
    ep.plane_scratch = (int16_t *)c->plane.p;
    ep.plane_scratch_elems = plane_elems;
    ep.cres = (ChanResult *)c->cres.p;
    ep.report = c->report_on ? (flo_cand_report *)c->report.p : nullptr;
    ep.smem_plane_bytes = (uint32_t)plane_cap;
    ep.work_bytes = (uint32_t)work;
    ep.defer_bytes = (uint32_t)defer_bytes;
    ep.defer_scratch = (uint8_t *)c->defer.p;
    ep.crc_tab = (const uint32_t *)c->crc_tab.p;
    ep.frame_crc = (uint32_t *)c->fcrc.p;
    ep.stagger = getenv("FLO_B200_STAGGER") ? (uint32_t)atoi(getenv("FLO_B200_STAGGER")) : 0u;

    FinalParams fp;
    memset(&fp, 0, sizeof fp);
    fp.tracks = ep.tracks; fp.n_tracks = NTR; fp.n_frames = NF; fp.level = level; fp.frames = ep.frames;
    fp.out = out; fp.meta = (const uint8_t *)c->meta.p;
    fp.frame_excl = ep.frame_excl; fp.frame_size = ep.frame_size;
    fp.track_crc = ctl + 48; fp.n_segs = NSEG;
    fp.frame_crc = (const uint32_t *)c->fcrc.p;
    fp.file_off = (unsigned long long *)c->foff.p;
    fp.file_len = fp.file_off + n_tracks;

    CK(launch_setup(ep.tracks, NTR, (uint2 *)c->frames.p, NF, st));
    launches += NF ? 1 : 0;
    CK(cudaEventRecord(c->ev[1], st));
    if (piped) { CK(cudaEventRecord(c->ev_fin, st)); CK(cudaStreamWaitEvent(s_in, c->ev_fin, 0)); }   // arenas are reused across calls

    // per-wave D2H sink (up to 256 tracks: header/TOC/metadata regions are copied one by one at the end)
    OutBlock *blk = nullptr;
    const bool sink_on = sink && piped && n_tracks <= 256 && L.out_bound >= SMALL_OUTPUT;
    struct WaveEnd { unsigned long long excl; uint32_t size; uint32_t pad; };
    WaveEnd *h_wave = nullptr;
    if (sink_on) {
        blk = take_block(L.out_bound);
        if (!blk) { set_err("cudaMallocHost(%llu) failed for the output block", (unsigned long long)L.out_bound); return FLO_ERR_NOMEM; }
        if ((rc = c->h_out.reserve(sizeof(WaveEnd) * MAX_WAVES + 64))) { drop_block(blk); return rc; }
        h_wave = (WaveEnd *)c->h_out.p;
    }
    auto fail = [&](cudaError_t e, const char *what) {
        cudaDeviceSynchronize();
        if (blk) drop_block(blk);
        set_err("%s: %s", what, cudaGetErrorString(e));
        return FLO_ERR_CUDA;
    };

    // track cursor for the per-wave sample copies
    size_t tcur = 0;
    std::vector<bool> pinned_src;         // is the caller's buffer of track t page-locked? (asked once per track)
    size_t pinned_known = 0;
    { cudaError_t e = cudaEventRecord(c->ev[2], st); if (e != cudaSuccess) return fail(e, "event record failed"); }
    uint32_t g0 = 0;
    for (int w = 0; w < n_waves; w++) {
        const uint32_t g1 = wave_end[w];
        if (host_inputs) {
            // samples of frames [g0, g1): per track the frames [max(g0, first) - first, min(g1, first + nf) - first)
            for (size_t t = tcur; t < n_tracks; t++) {
                const TrackDev &d = L.tr[t];
                const uint64_t f_lo = std::max<uint64_t>(g0, d.first_frame), f_hi = std::min<uint64_t>(g1, (uint64_t)d.first_frame + d.n_frames);
                if (d.first_frame >= g1 && d.n_frames) break;
                if (d.n_frames == 0 || f_hi <= f_lo) { if ((uint64_t)d.first_frame + d.n_frames <= g1) tcur = t + 1; continue; }
                const uint64_t spf = (uint64_t)d.sample_rate * d.channels;
                const uint64_t e_lo = (f_lo - d.first_frame) * spf;
                const uint64_t e_hi = (f_hi == (uint64_t)d.first_frame + d.n_frames) ? d.n_inter : (f_hi - d.first_frame) * spf;
                if (t >= pinned_known) { pinned_src.resize(t + 1, false); for (size_t q = pinned_known; q <= t; q++) pinned_src[q] = is_pinned(tracks[q].samples); pinned_known = t + 1; }
                cudaError_t e = h2d_from_caller(c, (uint8_t *)c->in.p + L.in_off[t] + e_lo * esz,
                                                (const uint8_t *)tracks[t].samples + e_lo * esz, (e_hi - e_lo) * esz,
                                                pinned_src[t], s_in);
                if (e != cudaSuccess) return fail(e, "H2D of the samples failed");
                if ((uint64_t)d.first_frame + d.n_frames <= g1) tcur = t + 1;
            }
            if (piped) {
                cudaError_t e = cudaEventRecord(c->ev_in[w], s_in);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(st, c->ev_in[w], 0);
                if (e != cudaSuccess) return fail(e, "stream ordering failed");
            }
        }
        if (convert) {                                     // single wave: the whole batch is resident here
            for (size_t t = 0; t < n_tracks; t++) {
                cudaError_t e = launch_ingest_convert(raw_src[t], (float *)((uint8_t *)c->conv.p + conv_off[t]), tracks[t].n_interleaved, format, st);
                if (e != cudaSuccess) return fail(e, "ingest pre-pass failed");
                launches += tracks[t].n_interleaved ? 1 : 0;
            }
        }
        ep.frame_begin = g0; ep.frame_end = g1; ep.ticket = n_waves > 1 ? wave_ticket + w : ctl;
        {
            cudaError_t e = var->launch(ep, grid, dyn, st);
            if (e != cudaSuccess) return fail(e, "encode kernel launch failed");
            launches += g1 > g0 ? 1 : 0;
        }
        if (sink_on && g1 > g0) {
            cudaError_t e = cudaMemcpyAsync(&h_wave[w].excl, ep.frame_excl + (g1 - 1), 8, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(&h_wave[w].size, ep.frame_size + (g1 - 1), 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaEventRecord(c->ev_k[w], st);
            if (e != cudaSuccess) return fail(e, "wave bookkeeping failed");
        }
        g0 = g1;
    }
    {
        cudaError_t e = cudaEventRecord(c->ev[3], st);
        if (e == cudaSuccess) e = launch_toc(fp, st);
        launches += NF ? 1 : 0;
        if (e == cudaSuccess) e = cudaEventRecord(c->ev[4], st);
        if (e == cudaSuccess) e = launch_crc_frames(fp, st);
        launches += NF ? 1 : 0;
        if (e == cudaSuccess) e = cudaEventRecord(c->ev[5], st);
        if (e == cudaSuccess) e = launch_headers(fp, st);
        launches += 1;
        if (e == cudaSuccess) e = cudaEventRecord(c->ev[6], st);
        if (e != cudaSuccess) return fail(e, "finalise kernels failed");
    }

    if (sink_on) {
        // DATA bytes of each wave leave the device as soon as its kernel is done, on their own stream
        auto data_base = [&](uint32_t g) {           // file position of the DATA chunk of the track that owns frame g
            size_t lo = 0, hi = n_tracks;
            while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (L.tr[mid].first_frame <= g) lo = mid; else hi = mid; }
            return L.tr[lo].static_off + FILE_HDR + 4 + 20ull * L.tr[lo].n_frames;
        };
        uint64_t cur = NF ? data_base(0) : 0;
        uint32_t gp = 0;
        for (int w = 0; w < n_waves; w++) {
            const uint32_t g1 = wave_end[w];
            if (g1 == gp) continue;
            cudaError_t e = cudaEventSynchronize(c->ev_k[w]);
            if (e != cudaSuccess) return fail(e, "wave synchronisation failed");
            const uint64_t end = data_base(g1 - 1) + h_wave[w].excl + h_wave[w].size;
            if (end > cur) {
                e = cudaMemcpyAsync(blk->base + cur, out + cur, end - cur, cudaMemcpyDeviceToHost, c->s_out);
                if (e != cudaSuccess) return fail(e, "D2H of a wave failed");
            }
            cur = end;
            gp = g1;
        }
    }
    // results: per-track offsets/lengths + error flag
    uint64_t *h_off = (uint64_t *)(hs + align_up(sizeof(TrackDev) * n_tracks + L.meta_total, 16));
    uint32_t *h_err = (uint32_t *)(h_off + 2 * n_tracks);           // err + 8 counters
    {
        cudaError_t e = cudaMemcpyAsync(h_off, c->foff.p, 16ull * n_tracks, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_err, ep.err, 4 * 47, cudaMemcpyDeviceToHost, st);   // err, counters[8], pad, phase clocks[16]
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return fail(e, "reading back the batch results failed");
    }
    if (*h_err) { if (blk) { cudaDeviceSynchronize(); drop_block(blk); } set_err("device-side consistency check failed (code 0x%08x)", *h_err); return FLO_ERR_INTERNAL; }
    for (size_t t = 0; t < n_tracks; t++) { offsets[t] = h_off[t]; lens[t] = h_off[n_tracks + t]; }
    for (int i = 0; i < 8; i++) c->counters[i] = h_err[1 + i];
    for (int i = 0; i < 16; i++) memcpy(&c->counters[8 + i], &h_err[15 + 2 * i], 8);
    if (sink_on) {
        // header + TOC in front of each DATA chunk and the metadata behind it were written by the final kernels
        for (size_t t = 0; t < n_tracks; t++) {
            const uint64_t head = FILE_HDR + 4 + 20ull * L.tr[t].n_frames;
            cudaError_t e = cudaMemcpyAsync(blk->base + offsets[t], out + offsets[t], head, cudaMemcpyDeviceToHost, c->s_out);
            if (e == cudaSuccess && L.tr[t].meta_len)
                e = cudaMemcpyAsync(blk->base + offsets[t] + lens[t] - L.tr[t].meta_len, out + offsets[t] + lens[t] - L.tr[t].meta_len,
                                    L.tr[t].meta_len, cudaMemcpyDeviceToHost, c->s_out);
            if (e != cudaSuccess) return fail(e, "D2H of a header failed");
        }
        cudaError_t e = cudaStreamSynchronize(c->s_out);
        if (e != cudaSuccess) return fail(e, "D2H synchronisation failed");
        *sink = blk;
    }

    float t_all = 0, t_enc = 0, t_toc = 0, t_crc = 0, t_hdr = 0, t_setup = 0, t_h2d = 0;
    cudaEventElapsedTime(&t_h2d, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&t_setup, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&t_enc, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&t_toc, c->ev[3], c->ev[4]);
    cudaEventElapsedTime(&t_crc, c->ev[4], c->ev[5]);
    cudaEventElapsedTime(&t_hdr, c->ev[5], c->ev[6]);
    cudaEventElapsedTime(&t_all, c->ev[1], c->ev[6]);
    c->ms[0] = t_all; c->ms[1] = t_enc; c->ms[2] = t_crc; c->ms[3] = t_setup + t_toc + t_hdr; c->ms[4] = t_h2d; c->ms[5] = 0;
    c->launches = launches;
    return FLO_OK;
}

extern "C" int flo_encode_batch_device(flo_ctx *c, const flo_track *tracks, size_t n_tracks, int format, uint8_t level,
                                       void *d_out, size_t d_out_capacity, uint64_t *offsets, uint64_t *lens) {
    if (!c || (n_tracks && (!tracks || !offsets || !lens)) || !d_out) { set_err("bad argument"); return FLO_ERR_ARG; }
    std::lock_guard<std::mutex> lk(c->mu);
    return encode_batch_impl(c, tracks, n_tracks, format, level, false, d_out, d_out_capacity, offsets, lens);
}

extern "C" int flo_encode_batch(flo_ctx *c, const flo_track *tracks, size_t n_tracks, int format, uint8_t level,
                                flo_out *outs) {
    if (!c || (n_tracks && (!tracks || !outs))) { set_err("bad argument"); return FLO_ERR_ARG; }
    std::lock_guard<std::mutex> lk(c->mu);
    for (size_t t = 0; t < n_tracks; t++) { outs[t].data = nullptr; outs[t].len = 0; }
    if (n_tracks == 0) return FLO_OK;
    std::vector<uint64_t> off(n_tracks), len(n_tracks);
    OutBlock *piped = nullptr;
    int rc = encode_batch_impl(c, tracks, n_tracks, format, level, true, nullptr, 0, off.data(), len.data(), &piped);
    if (rc) return rc;
    if (piped) {                       // the images already sit in a pinned block (copied out wave by wave)
        for (size_t t = 0; t < n_tracks; t++) { outs[t].data = piped->base + off[t]; outs[t].len = (size_t)len[t]; }
        publish_block(piped, n_tracks);
        return FLO_OK;
    }
    // one D2H of the compact image region, then split per track
    uint64_t total = 0;
    for (size_t t = 0; t < n_tracks; t++) total = std::max(total, off[t] + len[t]);
    cudaStream_t st = c->stream;
    if (total >= SMALL_OUTPUT) {
        // large result: straight into a pinned block, pointers into it are handed out (flo_free recycles it)
        OutBlock *blk = take_block(total + 16);
        if (!blk) { set_err("cudaMallocHost(%llu) failed for the output block", (unsigned long long)total); return FLO_ERR_NOMEM; }
        cudaError_t e = cudaEventRecord(c->ev[0], st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(blk->base, c->out.p, total, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(c->ev[1], st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { drop_block(blk); set_err("D2H of the result failed: %s", cudaGetErrorString(e)); return FLO_ERR_CUDA; }
        cudaEventElapsedTime(&c->ms[5], c->ev[0], c->ev[1]);
        for (size_t t = 0; t < n_tracks; t++) { outs[t].data = blk->base + off[t]; outs[t].len = (size_t)len[t]; }
        publish_block(blk, n_tracks);
        return FLO_OK;
    }
    if ((rc = c->h_out.reserve(total + 16))) return rc;
    CK(cudaEventRecord(c->ev[0], st));
    CK(cudaMemcpyAsync(c->h_out.p, c->out.p, total, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(c->ev[1], st));
    CK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c->ms[5], c->ev[0], c->ev[1]);
    for (size_t t = 0; t < n_tracks; t++) {
        uint8_t *b = (uint8_t *)malloc(len[t] ? len[t] : 1);
        if (!b) {
            for (size_t u = 0; u < t; u++) { free(outs[u].data); outs[u].data = nullptr; outs[u].len = 0; }
            set_err("out of host memory");
            return FLO_ERR_NOMEM;
        }
        memcpy(b, (const uint8_t *)c->h_out.p + off[t], len[t]);
        outs[t].data = b;
        outs[t].len = (size_t)len[t];
    }
    return FLO_OK;
}

static int encode_one(flo_ctx *c, const void *samples, size_t n, int format, uint32_t sr, uint8_t ch, uint8_t bits,
                      uint8_t level, const uint8_t *meta, size_t meta_len, uint8_t **out, size_t *out_len) {
    if (!c || !out || !out_len) { set_err("bad argument"); return FLO_ERR_ARG; }
    *out = nullptr; *out_len = 0;
    flo_track t;
    t.samples = samples; t.n_interleaved = n; t.sample_rate = sr; t.channels = ch; t.bit_depth = bits;
    t.meta = meta; t.meta_len = meta_len;
    flo_out o = {nullptr, 0};
    int rc = flo_encode_batch(c, &t, 1, format, level, &o);
    if (rc) return rc;
    *out = o.data; *out_len = o.len;
    return FLO_OK;
}

extern "C" int flo_encode(flo_ctx *c, const float *samples, size_t n, uint32_t sr, uint8_t ch, uint8_t bits, uint8_t level,
                          const uint8_t *meta, size_t meta_len, uint8_t **out, size_t *out_len) {
    return encode_one(c, samples, n, FLO_FMT_F32, sr, ch, bits, level, meta, meta_len, out, out_len);
}
extern "C" int flo_encode_pcm16(flo_ctx *c, const int16_t *pcm, size_t n, uint32_t sr, uint8_t ch, uint8_t bits,
                                uint8_t level, const uint8_t *meta, size_t meta_len, uint8_t **out, size_t *out_len) {
    return encode_one(c, pcm, n, FLO_FMT_PCM16, sr, ch, bits, level, meta, meta_len, out, out_len);
}

// ------------------------------------------------------------------------------------------------
// StreamingEncoder::encode_frame_data (libflo/src/streaming/encoder.rs:216-257) for a run of frames
// ------------------------------------------------------------------------------------------------
// The reference encodes one frame with Encoder::encode, reads the one-frame file back (Reader) and writes every
// channel again as [rice_parameter][coefficients as i32 LE][residual bytes] (serialize_channel, :243-257).
// Frames are independent in the encoder (encoder.rs:53-61), so frame g of the file of the whole run holds the
// bytes a one-frame file of second g would hold: the run goes through ONE device pass and the frames are
// re-serialised here, on the host, from that image (what Reader::read_channel_data reads, reader.rs:168-247).
extern "C" int flo_stream_encode_frames(flo_ctx *c, const float *samples, size_t n, uint32_t sr, uint8_t ch, uint8_t bits,
                                        uint8_t level, uint8_t **out, size_t *out_len, uint64_t **frame_off, uint32_t *n_frames) {
    if (!out || !out_len || !frame_off || !n_frames) { set_err("bad argument"); return FLO_ERR_ARG; }
    *out = nullptr; *out_len = 0; *frame_off = nullptr; *n_frames = 0;
    uint8_t *img = nullptr;
    size_t img_len = 0;
    int rc = encode_one(c, samples, n, FLO_FMT_F32, sr, ch, bits, level, nullptr, 0, &img, &img_len);
    if (rc) return rc;
    auto rd32 = [&](size_t p) { return (uint32_t)img[p] | ((uint32_t)img[p + 1] << 8) | ((uint32_t)img[p + 2] << 16) | ((uint32_t)img[p + 3] << 24); };
    const uint32_t nf = img_len >= 74 ? rd32(70) : 0;
    const size_t data0 = 74 + 20ull * nf;
    // the re-serialised frame is never longer than the frame itself
    uint8_t *buf = (uint8_t *)malloc(img_len > data0 ? img_len - data0 + 16 : 16);
    uint64_t *offs = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)nf + 1));
    if (!buf || !offs) { free(buf); free(offs); flo_free(img); set_err("out of host memory"); return FLO_ERR_NOMEM; }
    auto wr32 = [&](uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); };
    size_t w = 0, pos = data0;
    for (uint32_t g = 0; g < nf; g++) {
        offs[g] = w;
        const uint8_t ftype = img[pos];
        const uint32_t nsamp = rd32(pos + 1);
        memcpy(buf + w, img + pos, 6);                       // frame_type, frame_samples, flags (encoder.rs:226-230)
        w += 6; pos += 6;
        for (uint32_t q = 0; q < ch; q++) {
            const uint32_t size = rd32(pos);
            const uint8_t *body = img + pos + 4;
            pos += 4 + (size_t)size;
            uint8_t *lenp = buf + w;
            w += 4;
            size_t cl = 0;
            if (ftype == 254) {                              // Raw: at most 2 * frame_samples bytes (reader.rs:182-188)
                cl = std::min<size_t>(2ull * nsamp, size);
                memcpy(buf + w, body, cl);
            } else if (ftype >= 1 && ftype <= 12 && size >= 3) {   // ALPC: k, coefficients, residual bytes
                const uint32_t order = body[0];
                size_t p = 1 + 4ull * order;
                const uint8_t enc = body[p + 1];
                const uint8_t k = enc == 0 ? body[p + 2] : 0;
                p += enc == 0 ? 3 : 2;
                buf[w] = k;
                memcpy(buf + w + 1, body + 1, 4ull * order);
                memcpy(buf + w + 1 + 4ull * order, body + p, size - p);
                cl = 1 + 4ull * order + (size - p);
            }                                                // Silence and anything else: an empty channel
            wr32(lenp, (uint32_t)cl);
            w += cl;
        }
    }
    offs[nf] = w;
    flo_free(img);
    *out = buf; *out_len = w; *frame_off = offs; *n_frames = nf;
    return FLO_OK;
}

// ------------------------------------------------------------------------------------------------
// Waveform peaks of libflo::encode()'s analysis metadata (include/flo_b200.h: flo_waveform_peaks*)
// ------------------------------------------------------------------------------------------------
namespace {
// analysis.rs:44-54: how many windows, and samples per window in f64.  A zero channel count or sample rate makes
// the reference ask for a Vec of usize::MAX peaks (capacity overflow panic): an argument error here.
int peaks_plan(size_t n, uint32_t sr, uint8_t ch, uint32_t pps, double *spp_out, size_t *total_out) {
    *spp_out = 0.0; *total_out = 0;
    if (ch == 0 || sr == 0) { set_err("waveform peaks: channels and sample_rate must be non-zero"); return FLO_ERR_ARG; }
    if (n == 0) return FLO_OK;                                              // analysis.rs:44-50
    const double spp = (double)sr / (double)pps;                            // +inf for peaks_per_second = 0: no peaks
    const double t = std::ceil((double)n / (spp * (double)ch));
    if (!(t < 1e12)) { set_err("waveform peaks: %g windows requested", t); return FLO_ERR_ARG; }
    size_t total = t > 0 ? (size_t)t : 0;
    // the loop leaves at the first window that starts behind the input (analysis.rs:65-67); starts never decrease
    while (total > 0) {
        const double s = (double)(total - 1) * spp;
        const unsigned long long s0 = s >= 18446744073709551615.0 ? ~0ull : (unsigned long long)s;
        if (s0 > (~0ull) / ch || s0 * ch >= n) total--;
        else break;
    }
    *spp_out = spp; *total_out = total;
    return FLO_OK;
}
int peaks_impl(flo_ctx *c, const float *h_x, const float *d_x, size_t n, uint32_t sr, uint8_t ch, uint32_t pps,
               float *d_peaks_user, size_t cap, float **out, size_t *n_peaks) {
    if (!c || !n_peaks || (n && !h_x && !d_x) || (h_x && !out)) { set_err("bad argument"); return FLO_ERR_ARG; }
    *n_peaks = 0;
    if (out) *out = nullptr;
    double spp;
    size_t total;
    if (int rc = peaks_plan(n, sr, ch, pps, &spp, &total)) return rc;
    if (total == 0) return FLO_OK;
    if (d_peaks_user == nullptr && !out) { set_err("bad argument"); return FLO_ERR_ARG; }
    if (!out && total > cap) { set_err("waveform peaks: %zu peaks, room for %zu", total, cap); return FLO_ERR_ARG; }
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard dg(c->device);
    CK(dg.err);
    cudaStream_t st = c->stream;
    if (h_x) {
        if (int rc = c->in.reserve(n * sizeof(float))) return rc;
        CK(cudaMemcpyAsync(c->in.p, h_x, n * sizeof(float), cudaMemcpyHostToDevice, st));
        d_x = (const float *)c->in.p;
    } else if (((uintptr_t)d_x & 3u) != 0) { set_err("device samples pointer not aligned to the sample size"); return FLO_ERR_ARG; }
    if (int rc = c->peaks.reserve(total * sizeof(float) + 16)) return rc;
    flo::PeakParams pp;
    pp.x = d_x; pp.n = n; pp.spp = spp; pp.channels = ch; pp.n_peaks = total;
    pp.max_bits = (unsigned *)c->peaks.p;
    pp.peaks = out ? (float *)((uint8_t *)c->peaks.p + 16) : d_peaks_user;
    CK(flo::launch_waveform_peaks(pp, st));
    if (out) {
        float *h = (float *)malloc(total * sizeof(float));
        if (!h) { cudaStreamSynchronize(st); set_err("out of host memory"); return FLO_ERR_NOMEM; }
        cudaError_t e = cudaMemcpyAsync(h, pp.peaks, total * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { free(h); cudaGetLastError(); set_err("waveform peaks: %s", cudaGetErrorString(e)); return FLO_ERR_CUDA; }
        *out = h;
    } else CK(cudaStreamSynchronize(st));
    *n_peaks = total;
    return FLO_OK;
}
}  // namespace

extern "C" int flo_waveform_peaks(flo_ctx *c, const float *samples, size_t n, uint32_t sr, uint8_t ch, uint32_t pps, float **peaks, size_t *n_peaks) {
    if (!peaks) { set_err("bad argument"); return FLO_ERR_ARG; }
    return peaks_impl(c, samples, nullptr, n, sr, ch, pps, nullptr, 0, peaks, n_peaks);
}
extern "C" int flo_waveform_peaks_device(flo_ctx *c, const float *d_samples, size_t n, uint32_t sr, uint8_t ch, uint32_t pps,
                                         float *d_peaks, size_t capacity, size_t *n_peaks) {
    return peaks_impl(c, nullptr, d_samples, n, sr, ch, pps, d_peaks, capacity, nullptr, n_peaks);
}
extern "C" size_t flo_waveform_peaks_count(size_t n, uint32_t sr, uint8_t ch, uint32_t pps) {
    double spp;
    size_t total;
    return peaks_plan(n, sr, ch, pps, &spp, &total) ? 0 : total;
}

// ------------------------------------------------------------------------------------------------
// EBU R128 integrated loudness of libflo::encode()'s analysis metadata (include/flo_b200.h: flo_integrated_loudness*)
// ------------------------------------------------------------------------------------------------
namespace {
struct Biquad { double b0, b1, b2, a1, a2, z1, z2; };
inline double biquad_step(Biquad &f, double x) {                         // ebu_r128.rs:43-48
    const double y = f.b0 * x + f.z1;
    f.z1 = f.b1 * x - f.a1 * y + f.z2;
    f.z2 = f.b2 * x - f.a2 * y;
    return y;
}
void kweighting_coeffs(double sr, double co[10]) {                       // KWeighting::new, ebu_r128.rs:58-102
    const double pi = 3.14159265358979323846264338327950288;
    const double f0 = 1681.974450955533, g_db = 3.999843853973347, q = 0.7071752369554196;
    const double k = std::tan(pi * f0 / sr);
    const double vh = std::pow(10.0, g_db / 20.0);
    const double vb = std::pow(vh, 0.4996667741545416);
    const double a0 = 1.0 + k / q + k * k;
    co[0] = (vh + vb * k / q + k * k) / a0;
    co[1] = 2.0 * (k * k - vh) / a0;
    co[2] = (vh - vb * k / q + k * k) / a0;
    co[3] = 2.0 * (k * k - 1.0) / a0;
    co[4] = (1.0 - k / q + k * k) / a0;
    const double f0_hp = 38.13547087602444, q_hp = 0.5003270373238773;
    const double k_hp = std::tan(pi * f0_hp / sr);
    const double a0_hp = 1.0 + k_hp / q_hp + k_hp * k_hp;
    co[5] = 1.0; co[6] = -2.0; co[7] = 1.0;
    co[8] = 2.0 * (k_hp * k_hp - 1.0) / a0_hp;
    co[9] = (1.0 - k_hp / q_hp + k_hp * k_hp) / a0_hp;
}
// state (shelf z1, z2, high-pass z1, z2) after `steps` zero-input samples from each unit state: column c of M
void hop_transition(const double co[10], uint32_t steps, double M[4][4]) {
    for (int c = 0; c < 4; c++) {
        Biquad sh = {co[0], co[1], co[2], co[3], co[4], c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0};
        Biquad hp = {co[5], co[6], co[7], co[8], co[9], c == 2 ? 1.0 : 0.0, c == 3 ? 1.0 : 0.0};
        for (uint32_t i = 0; i < steps; i++) biquad_step(hp, biquad_step(sh, 0.0));
        M[0][c] = sh.z1; M[1][c] = sh.z2; M[2][c] = hp.z1; M[3][c] = hp.z2;
    }
    // what is left of the shelf's state after 100 ms is subnormal; as zero it changes nothing within 1e-290 and keeps
    // the chain below out of the subnormal slow path (9.8 -> 0.6 ms for the 72 000 hops of an hour of stereo)
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) if (std::fabs(M[r][c]) < 1e-290) M[r][c] = 0.0;
}
// the gating of compute_ebu_r128_loudness over the 400 ms block energies, ebu_r128.rs:267-313
double gated_loudness(const std::vector<double> &be) {
    if (be.empty()) return -23.0;
    const double abs_gate = std::pow(10.0, (-70.0 + 0.691) / 10.0);
    double sum_e = 0.0; size_t cnt = 0;
    for (double e : be) if (e >= abs_gate) { sum_e += e; cnt++; }
    if (!cnt) return -23.0;
    const double ungated = -0.691 + 10.0 * std::log10(sum_e / (double)cnt);
    const double rel_gate = std::pow(10.0, (ungated - 10.0 + 0.691) / 10.0);
    double s2 = 0.0; size_t c2 = 0;
    for (double e : be) if (e >= abs_gate && e >= rel_gate) { s2 += e; c2++; }
    return c2 ? -0.691 + 10.0 * std::log10(s2 / (double)c2) : ungated;
}
int loudness_impl(flo_ctx *c, const float *h_x, const float *d_x, size_t n, uint32_t sr, uint8_t ch, double *lufs) {
    if (!c || !lufs || (n && !h_x && !d_x)) { set_err("bad argument"); return FLO_ERR_ARG; }
    *lufs = -23.0;                                                       // ebu_r128.rs:187-194: no samples or no channels
    if (n == 0 || ch == 0) return FLO_OK;
    const double srd = (double)sr;
    const uint32_t hop = (uint32_t)std::llround(srd * 0.1);              // :197
    if (hop == 0) { set_err("loudness: sample_rate %u has no 100 ms hop (the reference never returns)", sr); return FLO_ERR_ARG; }
    const uint64_t frames = n / ch;
    if (frames == 0) return FLO_OK;                                      // no block at all: :267-275
    const uint64_t n_hops = (frames + hop - 1) / hop;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard dg(c->device);
    CK(dg.err);
    cudaStream_t st = c->stream;
    if (h_x) {
        if (int rc = c->in.reserve(n * sizeof(float))) return rc;
        CK(cudaMemcpyAsync(c->in.p, h_x, n * sizeof(float), cudaMemcpyHostToDevice, st));
        d_x = (const float *)c->in.p;
    } else if (((uintptr_t)d_x & 3u) != 0) { set_err("device samples pointer not aligned to the sample size"); return FLO_ERR_ARG; }
    const size_t n_seg = (size_t)n_hops * ch;
    if (int rc = c->peaks.reserve(n_seg * 5 * sizeof(double))) return rc;
    flo::KwParams kp;
    kp.x = d_x; kp.frames = frames; kp.channels = ch; kp.hop = hop; kp.n_hops = n_hops;
    kweighting_coeffs(srd, kp.co);
    // Below ~3.4 kHz the shelf's corner (1682 Hz) lies above Nyquist and the prototype yields an UNSTABLE biquad
    // (|a2| > 1): the reference's output there is an overflow to inf.  Hop states cannot be chained through an
    // unstable filter, and inf is not a loudness: refused.
    for (int q = 0; q < 2; q++) {
        const double a1 = kp.co[5 * q + 3], a2 = kp.co[5 * q + 4];
        if (!(std::fabs(a2) < 1.0 && std::fabs(a1) < 1.0 + a2)) {
            set_err("loudness: the K-weighting filter is unstable at %u Hz (the reference overflows to inf)", sr);
            return FLO_ERR_ARG;
        }
    }
    kp.state = (double *)c->peaks.p;
    kp.hop_sum = kp.state + 4 * n_seg;
    std::vector<double> hs(n_seg * 4);
    const bool dbg = getenv("FLO_B200_DEBUG_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    if (dbg) { cudaStreamSynchronize(st); t0 = now(); }
    CK(flo::launch_kweight(kp, 1, st));
    if (dbg) { cudaStreamSynchronize(st); t1 = now(); }
    CK(cudaMemcpyAsync(hs.data(), kp.state, hs.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (dbg) t2 = now();
    // chain the hops: start state of hop j + 1 = M * start state of hop j + forced response of hop j
    double M[4][4];
    hop_transition(kp.co, hop, M);
    for (uint32_t q = 0; q < ch; q++) {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        double *row = hs.data() + (size_t)q * n_hops * 4;
        for (uint64_t j = 0; j < n_hops; j++) {
            double f[4] = {row[4 * j], row[4 * j + 1], row[4 * j + 2], row[4 * j + 3]};
            for (int r = 0; r < 4; r++) row[4 * j + r] = s[r];
            double nx[4];
            for (int r = 0; r < 4; r++) nx[r] = M[r][0] * s[0] + M[r][1] * s[1] + M[r][2] * s[2] + M[r][3] * s[3] + f[r];
            for (int r = 0; r < 4; r++) s[r] = nx[r];
        }
    }
    if (dbg) t3 = now();
    CK(cudaMemcpyAsync(kp.state, hs.data(), hs.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(flo::launch_kweight(kp, 2, st));
    std::vector<double> sums(n_seg);
    CK(cudaMemcpyAsync(sums.data(), kp.hop_sum, n_seg * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (dbg) { t4 = now(); fprintf(stderr, "flo_b200 loudness: pass 1 %.3f ms, states to host %.3f ms, chain %.3f ms, upload + pass 2 + sums %.3f ms\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3); }
    // 400 ms blocks every 100 ms, the last one ends with the input (ebu_r128.rs:236-265)
    std::vector<double> be;
    const uint64_t block = 4ull * hop;
    for (uint64_t b = 0, start = 0; start < frames; b++, start += hop) {
        const uint64_t end = std::min(start + block, frames);
        const double len = (double)(end - start);
        double energy = 0.0;
        for (uint32_t q = 0; q < ch; q++) {
            double sum_sq = 0.0;
            for (uint64_t j = b; j < n_hops && j < b + 4; j++) sum_sq += sums[(size_t)q * n_hops + j];
            energy += sum_sq / len;
        }
        be.push_back(energy);
        if (end == frames) break;
    }
    *lufs = gated_loudness(be);
    return FLO_OK;
}
}  // namespace

extern "C" int flo_integrated_loudness(flo_ctx *c, const float *samples, size_t n, uint32_t sr, uint8_t ch, double *lufs) {
    return loudness_impl(c, samples, nullptr, n, sr, ch, lufs);
}
extern "C" int flo_integrated_loudness_device(flo_ctx *c, const float *d_samples, size_t n, uint32_t sr, uint8_t ch, double *lufs) {
    if (n && !d_samples) { set_err("bad argument"); return FLO_ERR_ARG; }
    return loudness_impl(c, nullptr, d_samples, n, sr, ch, lufs);
}

// ------------------------------------------------------------------------------------------------
// Lossless decoder (include/flo_b200.h: flo_decode / flo_decode_device)
// ------------------------------------------------------------------------------------------------
namespace {

struct FileHead {
    uint8_t vmaj, channels, bits, level;
    uint32_t sample_rate, crc, n_toc;
    uint64_t total, toc_size, data_size, extra_size, meta_size;
    uint64_t toc_pos, data_start, data_end, meta_off;
    bool tail_eof;
};
inline uint64_t le(const uint8_t *p, int n) { uint64_t v = 0; for (int i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i); return v; }

// Reader::read up to the DATA chunk (reader.rs:16-99), on the first min(len, 74) bytes of the file.
int parse_head(const uint8_t *h, size_t len, FileHead &H) {
    if (len < 4 || memcmp(h, "FLO!", 4) != 0) { set_err("Invalid flo file: bad magic"); return FLO_ERR_ARG; }
    if (len < 70) { set_err("Unexpected end of file"); return FLO_ERR_ARG; }
    H.vmaj = h[4]; H.sample_rate = (uint32_t)le(h + 8, 4); H.channels = h[12]; H.bits = h[13];
    H.total = le(h + 14, 8); H.level = h[22]; H.crc = (uint32_t)le(h + 26, 4);
    H.toc_size = le(h + 38, 8); H.data_size = le(h + 46, 8); H.extra_size = le(h + 54, 8); H.meta_size = le(h + 62, 8);
    H.n_toc = 0; H.toc_pos = 70; H.data_start = 70;
    if (H.toc_size >= 4) {
        if (len < 74) { set_err("Unexpected end of file"); return FLO_ERR_ARG; }
        H.n_toc = (uint32_t)le(h + 70, 4);
        if (H.n_toc > 100000) { set_err("Invalid TOC: too many entries"); return FLO_ERR_ARG; }
        H.toc_pos = 74;
        H.data_start = 74 + 20ull * H.n_toc;
        if (H.data_start > len) { set_err("Unexpected end of file"); return FLO_ERR_ARG; }
    }
    H.data_end = H.data_start + H.data_size;
    if (H.data_end < H.data_start) { set_err("Unexpected end of file"); return FLO_ERR_ARG; }
    // after the frames: pos = data_end; skip(extra) clamps to the file; read_bytes(meta_size) (reader.rs:39-43)
    uint64_t pos = H.data_end + H.extra_size;
    if (pos < H.data_end || pos > len) pos = len;
    H.meta_off = pos;
    H.tail_eof = H.meta_size > len - pos;
    return FLO_OK;
}

int decode_error(uint32_t key) {
    const uint32_t kind = key & 15u, frame = key >> 13;
    switch (kind) {
    case flo::DEC_TOO_MANY: set_err("Invalid frame: too many samples"); break;
    case flo::DEC_BAD_ORDER: set_err("Invalid LPC order"); break;
    case flo::DEC_EOF: set_err("Unexpected end of file"); break;
    case flo::DEC_TRANSFORM: set_err("frame %u is a transform (lossy) frame: not supported by the lossless GPU decoder", frame); break;
    default: set_err("frame %u: unsupported Rice parameter (> 31)", frame); break;
    }
    return FLO_ERR_ARG;
}

// i16: 0 = interleaved f32 (the reference's result), 1 = the integer samples saturated to i16 (flo_decode_i16)
int decode_impl(flo_ctx *c, const uint8_t *h_file, const void *d_file, size_t len, void *d_out_user, size_t cap,
                void **out, size_t *n_out, flo_info *info, int i16) {
    const size_t esz = i16 ? sizeof(int16_t) : sizeof(float);
    if (!c || !n_out || (!h_file && !d_file) || (h_file && !out)) { set_err("bad argument"); return FLO_ERR_ARG; }
    *n_out = 0;
    if (out) *out = nullptr;
    std::lock_guard<std::mutex> lk(c->mu);
    DeviceGuard dg(c->device);
    CK(dg.err);
    cudaStream_t st = c->stream;
    uint8_t head[80] = {0};
    const size_t have = len < 74 ? len : 74;
    if (h_file) memcpy(head, h_file, have);
    else if (have) { CK(cudaMemcpyAsync(head, d_file, have, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); }
    FileHead H;
    if (int rc = parse_head(head, len, H)) return rc;
    const uint32_t C = H.channels;

    const uint8_t *df = (const uint8_t *)d_file;
    CK(cudaEventRecord(c->ev[0], st));
    if (h_file) {
        if (int rc = c->in.reserve(len + 64)) return rc;
        CK(cudaMemcpyAsync(c->in.p, h_file, len, cudaMemcpyHostToDevice, st));
        df = (const uint8_t *)c->in.p;
    }
    CK(cudaEventRecord(c->ev[1], st));
    const size_t nf = H.n_toc ? H.n_toc : 1;
    if (int rc = c->dec_frames.reserve(nf * sizeof(flo::DecFrame))) return rc;
    if (int rc = c->dec_units.reserve(nf * (C ? C : 1) * sizeof(flo::DecUnit))) return rc;
    if (int rc = c->dec_base.reserve(nf * sizeof(uint64_t))) return rc;
    if (int rc = c->dec_ctl.reserve(64)) return rc;
    if (int rc = c->h_small.reserve(256)) return rc;
    uint32_t *hc = (uint32_t *)c->h_small.p;
    hc[0] = H.n_toc; hc[1] = 0xFFFFFFFFu; hc[2] = 0; hc[3] = 0;
    CK(cudaMemcpyAsync(c->dec_ctl.p, hc, 16, cudaMemcpyHostToDevice, st));

    flo::DecodeParams p;
    p.file = df; p.len = len; p.toc_pos = H.toc_pos; p.n_toc = H.n_toc; p.channels = C;
    p.data_start = H.data_start; p.data_end = H.data_end;
    p.frames = (flo::DecFrame *)c->dec_frames.p; p.units = (flo::DecUnit *)c->dec_units.p;
    p.base = (unsigned long long *)c->dec_base.p; p.ctl = (uint32_t *)c->dec_ctl.p; p.out = nullptr; p.out_i16 = i16;
    CK(flo::launch_decode_parse(p, st));
    CK(cudaEventRecord(c->ev[2], st));
    CK(cudaMemcpyAsync(hc + 8, c->dec_ctl.p, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint32_t keep = hc[8] < H.n_toc ? hc[8] : H.n_toc;
    if (hc[9] != 0xFFFFFFFFu && (hc[9] >> 13) < keep) return decode_error(hc[9]);
    const uint64_t total = (uint64_t)hc[10] | ((uint64_t)hc[11] << 32);
    const uint64_t n = total * C;
    *n_out = (size_t)n;

    void *d_out = d_out_user;
    if (h_file) {
        if (int rc = c->out.reserve(n * esz + 64)) return rc;
        d_out = c->out.p;
    } else if (n > cap || (n && !d_out_user)) {
        set_err("flo_decode_device: output capacity %zu samples, %llu needed", cap, (unsigned long long)n);
        return FLO_ERR_ARG;
    }
    p.out = d_out; p.out_i16 = i16;
    p.n_toc = keep;
    CK(cudaEventRecord(c->ev[3], st));
    uint32_t launches = H.n_toc ? 2 : 1;
    if (n) { CK(flo::launch_decode_units(p, st)); launches++; }
    CK(cudaEventRecord(c->ev[4], st));
    CK(cudaMemcpyAsync(hc + 8, c->dec_ctl.p, 16, cudaMemcpyDeviceToHost, st));

    OutBlock *blk = nullptr;
    void *h_out = nullptr;
    struct Guard {                                         // releases the result buffer on every early return
        OutBlock *&blk; void *&h_out; cudaStream_t st; bool armed = true;
        ~Guard() { if (!armed) return; cudaStreamSynchronize(st); cudaGetLastError(); if (blk) drop_block(blk); else free(h_out); }
    } guard{blk, h_out, st};
    if (h_file) {
        const size_t bytes = (size_t)n * esz;
        if (bytes >= SMALL_OUTPUT) {
            blk = take_block(bytes);
            if (!blk) { set_err("pinned output allocation (%zu bytes) failed", bytes); return FLO_ERR_NOMEM; }
            h_out = blk->base;
        } else {
            h_out = malloc(bytes ? bytes : 1);
            if (!h_out) { set_err("malloc(%zu) failed", bytes); return FLO_ERR_NOMEM; }
        }
        if (bytes) CK(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, st));
    }
    cudaEventRecord(c->ev[5], st);
    cudaError_t e = cudaStreamSynchronize(st);
    int rc = FLO_OK;
    if (e != cudaSuccess) { set_err("CUDA error %s in decode: %s", cudaGetErrorName(e), cudaGetErrorString(e)); rc = FLO_ERR_CUDA; }
    else if (hc[9] != 0xFFFFFFFFu && (hc[9] >> 13) < keep) rc = decode_error(hc[9]);
    else if (H.tail_eof) { set_err("Unexpected end of file"); rc = FLO_ERR_ARG; }
    if (rc) {
        *n_out = 0;
        return rc;
    }
    guard.armed = false;
    if (h_file) {
        if (blk) publish_block(blk, 1);
        *out = h_out;
    }
    float t = 0;
    for (float &m : c->ms) m = 0;
    if (cudaEventElapsedTime(&t, c->ev[1], c->ev[4]) == cudaSuccess) c->ms[0] = t;
    if (cudaEventElapsedTime(&t, c->ev[3], c->ev[4]) == cudaSuccess) c->ms[1] = t;
    if (cudaEventElapsedTime(&t, c->ev[1], c->ev[2]) == cudaSuccess) c->ms[3] = t;
    if (cudaEventElapsedTime(&t, c->ev[0], c->ev[1]) == cudaSuccess) c->ms[4] = t;
    if (cudaEventElapsedTime(&t, c->ev[4], c->ev[5]) == cudaSuccess) c->ms[5] = t;
    c->launches = launches;
    if (info) {
        info->sample_rate = H.sample_rate; info->channels = H.channels; info->bit_depth = H.bits; info->level = H.level;
        info->version_major = H.vmaj; info->total_samples = H.total; info->decoded_frames = total; info->n_frames = keep;
        info->data_crc32 = H.crc; info->meta_offset = H.meta_off; info->meta_size = H.meta_size;
    }
    return FLO_OK;
}

}  // namespace

extern "C" int flo_decode(flo_ctx *c, const uint8_t *file, size_t len, float **out, size_t *n_interleaved, flo_info *info) {
    if (!file) { set_err("bad argument"); return FLO_ERR_ARG; }
    return decode_impl(c, file, nullptr, len, nullptr, 0, (void **)out, n_interleaved, info, 0);
}
extern "C" int flo_decode_device(flo_ctx *c, const void *d_file, size_t len, float *d_out, size_t d_out_capacity,
                                 size_t *n_interleaved, flo_info *info) {
    if (!d_file) { set_err("bad argument"); return FLO_ERR_ARG; }
    return decode_impl(c, nullptr, d_file, len, d_out, d_out_capacity, nullptr, n_interleaved, info, 0);
}
extern "C" int flo_decode_i16(flo_ctx *c, const uint8_t *file, size_t len, int16_t **out, size_t *n_interleaved, flo_info *info) {
    if (!file) { set_err("bad argument"); return FLO_ERR_ARG; }
    return decode_impl(c, file, nullptr, len, nullptr, 0, (void **)out, n_interleaved, info, 1);
}
extern "C" int flo_decode_i16_device(flo_ctx *c, const void *d_file, size_t len, int16_t *d_out, size_t d_out_capacity,
                                     size_t *n_interleaved, flo_info *info) {
    if (!d_file) { set_err("bad argument"); return FLO_ERR_ARG; }
    return decode_impl(c, nullptr, d_file, len, d_out, d_out_capacity, nullptr, n_interleaved, info, 1);
}
