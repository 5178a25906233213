// engine.cu — the host-buffer picture pipeline of libsvtav1_b200.so (include/svt_av1_b200.h, "picture engine").
//
// The picture-level entries of the other files take DEVICE pointers and a stream.  The reference's process loops
// (motion_estimation_kernel, dlf_kernel, cdef_kernel; SURVEY.md §8b "Batched entry" / "Memory ownership") own HOST
// pictures and call from many pipeline threads at once, one picture per call.  The engine is the piece in between,
// inside the product library so that a C host gets it by linking:
//   * a residency cache of the three padded luma planes of every picture that ME touches (as source or as reference),
//     keyed by (host object, picture number): a picture is uploaded once and then serves as reference for the pictures
//     that follow it (EbPaReferenceObject lifetime, EbPictureBufferDesc.c:65-78 layout);
//   * per-call slots (stream + device scratch/outputs + pinned staging) so N threads run N pictures at once;
//   * the deblock -> CDEF search -> (host strength decision) -> CDEF apply chain of one picture with the
//     reconstruction resident on the device across the stages.
// Measured on the 128-thread host of the B200 box (profiles/r2_engine_*.txt): what costs in this setting is not the GPU
// (0.3 ms per 1080p picture) but CUDA calls that take process-wide locks while ~20 pipeline threads are inside the
// library — cudaHostRegister of the caller's buffers (4 ms each and it stalls every other thread), cudaMalloc (device
// synchronisation) and pageable cudaMemcpy (holds the context lock for the whole host copy).  So every buffer is
// allocated ONCE, when the first picture fixes the geometry (one device arena, one pinned arena), host pictures are
// packed into the slot's pinned staging with plain memcpy outside any CUDA lock, and each call makes ~10 short
// asynchronous CUDA calls and one stream wait.
// Every entry is synchronous for its caller (the reference's stage returns when its picture is done) and re-entrant.
// There is no CPU fallback: a CUDA failure is returned as SVT_B200_ERR_CUDA and the integration aborts the encode.
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <sys/resource.h>
#include <vector>

#include "common.cuh"

using namespace svtb200;

namespace {

constexpr int kMeSlots = 8;       // concurrent ME pictures
constexpr int kFiltSlots = 8;     // concurrent deblock / CDEF pictures
constexpr int kPlaneEntries = 96; // resident ME pictures (3.3 MB each at 1080p, 13 MB at 2160p); >= kMeSlots * 9 so that
                                  // every in-flight picture can hold its source + 8 references at once (no deadlock)

inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

struct PlaneEntry {
    const void *key = nullptr;
    uint64_t tag = 0;
    uint8_t *d[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ready = nullptr;
    int state = 0; // 0 free, 1 loading, 2 ready
    int users = 0;
    uint64_t stamp = 0;
};

struct MeSlot {
    bool busy = false;
    cudaStream_t st = nullptr;
    cudaEvent_t up_done = nullptr; // the staging buffer has been consumed by its upload
    bool up_pending = false;
    uint8_t *scratch = nullptr;
    uint8_t *dev = nullptr;     // best_sad | best_mv | hme | me_mv | me_cand | total | rc
    uint8_t *pin_up = nullptr;  // full | quarter | sixteenth of one picture
    uint8_t *pin_out = nullptr; // me_mv | me_cand | total | rc
};

struct FiltSlot {
    bool busy = false;
    cudaStream_t st = nullptr;
    SvtB200Frame recon, out, source; // device pictures
    uint8_t *misc = nullptr;         // mi array | skip8 | mse | fb idx
    uint8_t *pin_a = nullptr;        // packed picture (reconstruction up / down)
    uint8_t *pin_b = nullptr;        // packed picture (source); the mode-info array for deblocking
    uint8_t *pin_small = nullptr;    // skip map | mse | fb idx | decision
    uint8_t *pin_mi = nullptr;       // the mode-info array when the call must not reuse pin_b (single-sync path)
};

struct MeGeom {
    size_t nb[3] = {0, 0, 0}, entry = 0, n_sb = 0, scratch = 0;
    size_t b_sad = 0, b_hme = 0, b_mv = 0, b_cand = 0, b_tot = 0, b_rc = 0, o_mv = 0, o_cand = 0, o_tot = 0, o_rc = 0, dev = 0, out = 0;
    bool operator==(const MeGeom &o) const { return nb[0] == o.nb[0] && nb[1] == o.nb[1] && nb[2] == o.nb[2] && n_sb == o.n_sb; }
};

struct FiltGeom {
    int w = 0, h = 0, bd = 0, mi_rows = 0, mi_cols = 0;
    int sy = 0, sc = 0; // device strides (samples)
    size_t by = 0, bc = 0, frame = 0, packed = 0, mi = 0, skip = 0, mse = 0, nfb = 0, misc = 0, small = 0, idx = 0, dec = 0;
    bool same(int w_, int h_, int bd_) const { return w == w_ && h == h_ && bd == bd_; }
};

} // namespace

struct SvtB200Engine {
    int device = 0;
    std::mutex mu;
    std::condition_variable cv;
    uint64_t clock = 0;
    PlaneEntry planes[kPlaneEntries];
    MeSlot me[kMeSlots];
    FiltSlot filt[kFiltSlots];
    // geometry-bound memory, allocated once by the first picture of each kind
    std::mutex init_mu;
    bool me_ready = false, filt_ready = false;
    MeGeom mg;
    FiltGeom fg;
    uint8_t *me_dev = nullptr, *me_pin = nullptr, *filt_dev = nullptr, *filt_pin = nullptr;
    struct {
        std::atomic<uint64_t> me_pictures{0}, dlf_frames{0}, cdef_frames{0}, lr_frames{0}, me_plane_uploads{0}, me_plane_hits{0}, h2d_bytes{0},
            d2h_bytes{0}, pinned_bytes{0}, ns_slot_wait{0}, ns_pin{0}, ns_plane_wait{0}, ns_issue{0}, ns_sync{0}, ns_host_copy{0},
            pin_calls{0};
    } stats;
};

namespace {

inline uint64_t now_ns() {
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec;
}
struct Lap { // accumulates wall time into an engine counter
    std::atomic<uint64_t> &acc;
    uint64_t t0;
    explicit Lap(std::atomic<uint64_t> &a) : acc(a), t0(now_ns()) {}
    ~Lap() { acc += now_ns() - t0; }
};

#define ENG_TRY(expr)                                                                                              \
    do {                                                                                                           \
        cudaError_t e__ = (expr);                                                                                  \
        if (e__ != cudaSuccess) {                                                                                  \
            set_error("engine: %s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);        \
            return SVT_B200_ERR_CUDA;                                                                              \
        }                                                                                                          \
    } while (0)

template <typename Slot, int N>
Slot *acquire(SvtB200Engine *e, Slot (&pool)[N]) {
    Lap lap(e->stats.ns_slot_wait);
    std::unique_lock<std::mutex> lk(e->mu);
    for (;;) {
        for (int i = 0; i < N; i++)
            if (!pool[i].busy) {
                pool[i].busy = true;
                return &pool[i];
            }
        e->cv.wait(lk);
    }
}
template <typename Slot>
void release(SvtB200Engine *e, Slot *s) {
    {
        std::lock_guard<std::mutex> g(e->mu);
        s->busy = false;
    }
    e->cv.notify_all();
}

cudaError_t timed_sync(SvtB200Engine *e, cudaStream_t st) {
    Lap lap(e->stats.ns_sync);
    return cudaStreamSynchronize(st);
}

size_t plane_bytes(const SvtB200Plane &g) { return (size_t)g.stride * (size_t)(g.height + 2 * g.origin_y); }

// ---- one-time allocation for the ME side --------------------------------------------------------------------------
int ensure_me(SvtB200Engine *e, const SvtB200MeParams *p) {
    MeGeom g;
    g.nb[0] = plane_bytes(p->full);
    g.nb[1] = plane_bytes(p->quarter);
    g.nb[2] = plane_bytes(p->sixteenth);
    g.n_sb = (size_t)((p->full.width + 63) / 64) * ((p->full.height + 63) / 64);
    std::lock_guard<std::mutex> lk(e->init_mu);
    if (e->me_ready) {
        if (e->mg == g) return SVT_B200_OK;
        set_error("engine: the ME picture geometry changed (one engine serves one sequence geometry)");
        return SVT_B200_ERR_UNSUPPORTED;
    }
    g.entry = al256(g.nb[0]) + al256(g.nb[1]) + al256(g.nb[2]);
    g.scratch = al256(svt_b200_me_scratch_bytes(p));
    g.b_sad = g.n_sb * 8 * 85 * 4;
    g.b_hme = g.n_sb * 8 * sizeof(SvtB200HmeResult);
    g.b_mv = g.n_sb * 85 * 7 * 4;
    g.b_cand = g.n_sb * 85 * 23;
    g.b_tot = g.n_sb * 85;
    g.b_rc = g.n_sb * 4;
    g.o_mv = 2 * al256(g.b_sad) + al256(g.b_hme);
    g.o_cand = g.o_mv + al256(g.b_mv);
    g.o_tot = g.o_cand + al256(g.b_cand);
    g.o_rc = g.o_tot + al256(g.b_tot);
    g.dev = g.o_rc + al256(g.b_rc);
    g.out = g.dev - g.o_mv;
    const size_t dev_total = (size_t)kPlaneEntries * g.entry + (size_t)kMeSlots * (g.scratch + g.dev);
    const size_t pin_total = (size_t)kMeSlots * (g.entry + al256(g.out));
    ENG_TRY(cudaMalloc((void **)&e->me_dev, dev_total));
    ENG_TRY(cudaMallocHost((void **)&e->me_pin, pin_total));
    e->stats.pinned_bytes += pin_total;
    uint8_t *d = e->me_dev, *h = e->me_pin;
    for (auto &c : e->planes) {
        c.d[0] = d;
        c.d[1] = d + al256(g.nb[0]);
        c.d[2] = c.d[1] + al256(g.nb[1]);
        d += g.entry;
    }
    for (auto &s : e->me) {
        s.scratch = d;
        s.dev = d + g.scratch;
        d += g.scratch + g.dev;
        s.pin_up = h;
        s.pin_out = h + g.entry;
        h += g.entry + al256(g.out);
    }
    e->mg = g;
    e->me_ready = true;
    return SVT_B200_OK;
}

// ---- one-time allocation for the deblock / CDEF side ---------------------------------------------------------------
int ensure_filt(SvtB200Engine *e, int w, int h, int bd, int mi_rows, int mi_cols) {
    std::lock_guard<std::mutex> lk(e->init_mu);
    if (e->filt_ready) {
        if (e->fg.same(w, h, bd)) return SVT_B200_OK;
        set_error("engine: the picture geometry changed (one engine serves one sequence geometry)");
        return SVT_B200_ERR_UNSUPPORTED;
    }
    FiltGeom g;
    g.w = w, g.h = h, g.bd = bd, g.mi_rows = mi_rows, g.mi_cols = mi_cols;
    const int bps = bd > 8 ? 2 : 1, cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    g.sy = (w + 63) & ~63; // rows start 64-byte (8-bit) / 128-byte (16-bit) aligned: 128-bit vector access on every row
    g.sc = (cw + 63) & ~63;
    g.by = al256((size_t)g.sy * h * bps) + 256; // + slack between planes
    g.bc = al256((size_t)g.sc * ch * bps) + 256;
    g.frame = g.by + 2 * g.bc + 1024;
    g.packed = al256(((size_t)w * h + 2 * (size_t)cw * ch) * bps);
    g.mi = al256((size_t)mi_rows * mi_cols * sizeof(SvtB200DlfMi));
    g.nfb = (size_t)((mi_rows + 15) / 16) * ((mi_cols + 15) / 16);
    g.skip = al256((size_t)((mi_rows + 1) / 2) * ((((mi_cols + 1) / 2) + 15) & ~15));
    g.mse = al256(g.nfb * 2 * 64 * 8);
    g.idx = al256(g.nfb);
    g.dec = al256(sizeof(SvtB200CdefDecision)) + al256(g.nfb * (16 + 16 * 64) + 64); // the device decision + cdef_decide's scratch (any table size)
    g.small = g.skip + g.mse + g.idx + g.dec;
    g.misc = g.mi + 4096 + g.small; // mode-info summary | the level search's scratch | skip map, mse, filter-block indices
    const size_t dev_total = (size_t)kFiltSlots * (3 * g.frame + g.misc);
    const size_t pin_b = std::max(g.packed, g.mi);
    const size_t pin_total = (size_t)kFiltSlots * (g.packed + pin_b + g.small + g.mi);
    ENG_TRY(cudaMalloc((void **)&e->filt_dev, dev_total));
    ENG_TRY(cudaMallocHost((void **)&e->filt_pin, pin_total));
    e->stats.pinned_bytes += pin_total;
    uint8_t *d = e->filt_dev, *hp = e->filt_pin;
    for (auto &s : e->filt) {
        SvtB200Frame *fr[3] = {&s.recon, &s.out, &s.source};
        for (SvtB200Frame *f : fr) {
            f->y = d + 512; // margin: the filters' vector loads may touch a few bytes before sample (0,0)
            f->cb = d + 512 + g.by;
            f->cr = d + 512 + g.by + g.bc;
            f->stride_y = g.sy;
            f->stride_c = g.sc;
            f->width = w;
            f->height = h;
            f->bit_depth = bd;
            d += g.frame;
        }
        s.misc = d;
        d += g.misc;
        s.pin_a = hp;
        s.pin_b = hp + g.packed;
        s.pin_small = s.pin_b + pin_b;
        s.pin_mi = s.pin_small + g.small;
        hp += g.packed + pin_b + g.small + g.mi;
    }
    e->fg = g;
    e->filt_ready = true;
    return SVT_B200_OK;
}

// memcpy split over four threads above 4 MB (see run_copies below)
// copies at or below this size stay on the calling thread (SVT_B200_COPY_MIN_KB; thread creation costs ~50 us)
size_t copy_split_bytes() {
    static const size_t v = getenv("SVT_B200_COPY_MIN_KB") ? (size_t)atol(getenv("SVT_B200_COPY_MIN_KB")) << 10 : (size_t)4 << 20;
    return v;
}

void par_memcpy(void *dst, const void *src, size_t n) {
    if (n <= copy_split_bytes()) {
        memcpy(dst, src, n);
        return;
    }
    const size_t q = ((n / 4) + 4095) & ~(size_t)4095;
    std::thread th[3];
    for (int t = 1; t < 4; t++) {
        const size_t o = q * t, len = o >= n ? 0 : std::min(q, n - o);
        th[t - 1] = std::thread([=] {
            if (len) memcpy((uint8_t *)dst + o, (const uint8_t *)src + o, len);
        });
    }
    memcpy(dst, src, std::min(q, n));
    for (auto &t : th) t.join();
}

// Make the three planes of `pic` resident and ordered before later work on the slot's stream.  users++ on return.
int plane_acquire(SvtB200Engine *e, MeSlot *s, const SvtB200MeParams *p, const SvtB200HostMePicture *pic, int filtered_ds,
                  PlaneEntry **out) {
    const MeGeom &g = e->mg;
    PlaneEntry *en = nullptr;
    bool loader = false;
    {
        std::unique_lock<std::mutex> lk(e->mu);
        for (;;) {
            PlaneEntry *victim = nullptr;
            for (auto &c : e->planes) {
                if (c.state && c.key == pic->key && c.tag == pic->tag) {
                    en = &c;
                    break;
                }
                if (c.users == 0 && c.state != 1 &&
                    (!victim || c.state < victim->state || (c.state == victim->state && c.stamp < victim->stamp)))
                    victim = &c;
            }
            if (en) {
                en->users++;
                {
                    Lap lap(e->stats.ns_plane_wait);
                    while (en->state == 1) e->cv.wait(lk);
                }
                if (en->state != 2) { // the loader failed
                    en->users--;
                    set_error("engine: upload of a reference picture failed in another thread");
                    return SVT_B200_ERR_CUDA;
                }
                e->stats.me_plane_hits++;
                break;
            }
            if (victim) {
                en = victim;
                en->key = pic->key;
                en->tag = pic->tag;
                en->state = 1;
                en->users = 1;
                loader = true;
                e->stats.me_plane_uploads++;
                break;
            }
            e->cv.wait(lk); // every entry is in use: wait for a picture to finish
        }
        en->stamp = ++e->clock;
    }
    if (loader) {
        int rc = SVT_B200_OK;
        const bool gen = !pic->quarter || !pic->sixteenth; // decimations derived on the device
        if (s->up_pending) { // the staging buffer still feeds the previous upload of this call
            if (cudaEventSynchronize(s->up_done) != cudaSuccess) rc = SVT_B200_ERR_CUDA;
            s->up_pending = false;
        }
        size_t bytes = 0;
        if (rc == SVT_B200_OK) {
            Lap lap(e->stats.ns_host_copy);
            par_memcpy(s->pin_up, pic->full, g.nb[0]);
            bytes = g.nb[0];
            if (!gen) {
                memcpy(s->pin_up + al256(g.nb[0]), pic->quarter, g.nb[1]);
                memcpy(s->pin_up + al256(g.nb[0]) + al256(g.nb[1]), pic->sixteenth, g.nb[2]);
                bytes = al256(g.nb[0]) + al256(g.nb[1]) + g.nb[2]; // the three planes are laid out alike on both sides
            }
        }
        if (rc == SVT_B200_OK && cudaMemcpyAsync(en->d[0], s->pin_up, bytes, cudaMemcpyHostToDevice, s->st) != cudaSuccess)
            rc = SVT_B200_ERR_CUDA;
        if (rc == SVT_B200_OK && cudaEventRecord(s->up_done, s->st) != cudaSuccess) rc = SVT_B200_ERR_CUDA;
        s->up_pending = rc == SVT_B200_OK;
        e->stats.h2d_bytes += bytes;
        if (rc == SVT_B200_OK && gen) {
            SvtB200MePlanes dp = {en->d[0], en->d[1], en->d[2]};
            rc = svt_b200_me_downsample(&p->full, &p->quarter, &p->sixteenth, &dp, filtered_ds, s->st);
        }
        if (rc == SVT_B200_OK && cudaEventRecord(en->ready, s->st) != cudaSuccess) rc = SVT_B200_ERR_CUDA;
        {
            std::lock_guard<std::mutex> gl(e->mu);
            en->state = rc == SVT_B200_OK ? 2 : 0;
            if (rc != SVT_B200_OK) en->users--;
        }
        e->cv.notify_all();
        if (rc != SVT_B200_OK) {
            if (rc == SVT_B200_ERR_CUDA) set_error("engine: ME plane upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            return rc;
        }
    } else {
        ENG_TRY(cudaStreamWaitEvent(s->st, en->ready, 0));
    }
    *out = en;
    return SVT_B200_OK;
}

void plane_release(SvtB200Engine *e, PlaneEntry **ents, int n) {
    {
        std::lock_guard<std::mutex> g(e->mu);
        for (int i = 0; i < n; i++)
            if (ents[i]) ents[i]->users--;
    }
    e->cv.notify_all();
}

// host picture <-> packed (tight rows: y, cb, cr) pinned staging.  A 2160p 10-bit picture is 25 MB and one pipeline
// thread copies it at 2-3 GB/s (the picture pools of the encoder are spread over both sockets: profiles/r2_encoder_2160p_*),
// so pictures above 4 MB are split by rows over four threads (the host has 128 hardware threads and ~10 busy ones).
struct RowCopy {
    uint8_t *dst;
    const uint8_t *src;
    size_t dst_pitch, src_pitch, bytes;
    int rows;
};
void copy_rows(const RowCopy &c, int r0, int r1) {
    for (int y = r0; y < r1; y++) memcpy(c.dst + (size_t)y * c.dst_pitch, c.src + (size_t)y * c.src_pitch, c.bytes);
}
void run_copies(SvtB200Engine *e, const RowCopy (&c)[3]) {
    Lap lap(e->stats.ns_host_copy);
    size_t total = 0;
    for (const RowCopy &k : c) total += k.bytes * k.rows;
    static const int trace = getenv("SVT_B200_ENGINE_TRACE") ? atoi(getenv("SVT_B200_ENGINE_TRACE")) : 0;
    static const int max_threads = getenv("SVT_B200_COPY_THREADS") ? atoi(getenv("SVT_B200_COPY_THREADS")) : 4;
    const uint64_t t0 = trace ? now_ns() : 0;
    struct Tr {
        uint64_t t0;
        size_t n;
        int to_host;
        rusage r0;
        ~Tr() {
            if (!t0) return;
            rusage r1;
            getrusage(RUSAGE_THREAD, &r1);
            fprintf(stderr, "engine trace: picture copy %s %.1f MB in %.3f ms (minflt %ld majflt %ld nvcsw %ld nivcsw %ld cpu %.3f ms)\n",
                    to_host ? "pinned->host" : "host->pinned", n / 1e6, (now_ns() - t0) / 1e6, r1.ru_minflt - r0.ru_minflt,
                    r1.ru_majflt - r0.ru_majflt, r1.ru_nvcsw - r0.ru_nvcsw, r1.ru_nivcsw - r0.ru_nivcsw,
                    ((r1.ru_utime.tv_sec - r0.ru_utime.tv_sec) + (r1.ru_stime.tv_sec - r0.ru_stime.tv_sec)) * 1e3 +
                        ((r1.ru_utime.tv_usec - r0.ru_utime.tv_usec) + (r1.ru_stime.tv_usec - r0.ru_stime.tv_usec)) / 1e3);
        }
    } tr{t0, total, 0, {}};
    if (trace) {
        getrusage(RUSAGE_THREAD, &tr.r0);
        tr.to_host = c[0].dst_pitch != c[0].bytes; // unpack writes strided host rows
    }
    const int nt = total > copy_split_bytes() ? std::max(1, std::min(8, max_threads)) : 1;
    auto part = [&c, nt](int t) {
        for (const RowCopy &k : c) copy_rows(k, (int)((int64_t)k.rows * t / nt), (int)((int64_t)k.rows * (t + 1) / nt));
    };
    if (nt == 1) {
        part(0);
        return;
    }
    std::thread th[7];
    for (int t = 1; t < nt; t++) th[t - 1] = std::thread(part, t);
    part(0);
    for (int t = 1; t < nt; t++) th[t - 1].join();
}
void pack_frame(SvtB200Engine *e, uint8_t *pin, const SvtB200Frame *h) {
    const int bps = h->bit_depth > 8 ? 2 : 1, cw = (h->width + 1) >> 1, ch = (h->height + 1) >> 1;
    const size_t rw = (size_t)h->width * bps, rc = (size_t)cw * bps;
    const RowCopy c[3] = {{pin, (const uint8_t *)h->y, rw, (size_t)h->stride_y * bps, rw, h->height},
                          {pin + rw * h->height, (const uint8_t *)h->cb, rc, (size_t)h->stride_c * bps, rc, ch},
                          {pin + rw * h->height + rc * ch, (const uint8_t *)h->cr, rc, (size_t)h->stride_c * bps, rc, ch}};
    run_copies(e, c);
}
void unpack_frame(SvtB200Engine *e, const SvtB200Frame *h, const uint8_t *pin) {
    const int bps = h->bit_depth > 8 ? 2 : 1, cw = (h->width + 1) >> 1, ch = (h->height + 1) >> 1;
    const size_t rw = (size_t)h->width * bps, rc = (size_t)cw * bps;
    const RowCopy c[3] = {{(uint8_t *)h->y, pin, (size_t)h->stride_y * bps, rw, rw, h->height},
                          {(uint8_t *)h->cb, pin + rw * h->height, (size_t)h->stride_c * bps, rc, rc, ch},
                          {(uint8_t *)h->cr, pin + rw * h->height + rc * ch, (size_t)h->stride_c * bps, rc, rc, ch}};
    run_copies(e, c);
}
// packed pinned staging <-> device picture (three 2-D copies; both sides page-locked / device: pure DMA, asynchronous)
int copy_packed(SvtB200Engine *e, const SvtB200Frame *dev, uint8_t *pin, cudaMemcpyKind kind, cudaStream_t st) {
    const int bps = dev->bit_depth > 8 ? 2 : 1, cw = (dev->width + 1) >> 1, ch = (dev->height + 1) >> 1;
    const size_t rw = (size_t)dev->width * bps, rc = (size_t)cw * bps;
    uint8_t *py = pin, *pcb = pin + rw * dev->height, *pcr = pcb + rc * ch;
    if (kind == cudaMemcpyHostToDevice) {
        ENG_TRY(cudaMemcpy2DAsync(dev->y, (size_t)dev->stride_y * bps, py, rw, rw, dev->height, kind, st));
        ENG_TRY(cudaMemcpy2DAsync(dev->cb, (size_t)dev->stride_c * bps, pcb, rc, rc, ch, kind, st));
        ENG_TRY(cudaMemcpy2DAsync(dev->cr, (size_t)dev->stride_c * bps, pcr, rc, rc, ch, kind, st));
        e->stats.h2d_bytes += rw * dev->height + 2 * rc * ch;
    } else {
        ENG_TRY(cudaMemcpy2DAsync(py, rw, dev->y, (size_t)dev->stride_y * bps, rw, dev->height, kind, st));
        ENG_TRY(cudaMemcpy2DAsync(pcb, rc, dev->cb, (size_t)dev->stride_c * bps, rc, ch, kind, st));
        ENG_TRY(cudaMemcpy2DAsync(pcr, rc, dev->cr, (size_t)dev->stride_c * bps, rc, ch, kind, st));
        e->stats.d2h_bytes += rw * dev->height + 2 * rc * ch;
    }
    return SVT_B200_OK;
}

struct DeviceGuard { // engine calls come from arbitrary pipeline threads: bind the engine's device for the call
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev)
            cudaSetDevice(dev);
        else
            prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

} // namespace

extern "C" {

int svt_b200_engine_create(int device, SvtB200Engine **out) {
    if (!out) return SVT_B200_ERR_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        set_error("svt_b200_engine_create: CUDA device %d not present (%d devices) - no CPU fallback", device, n);
        return SVT_B200_ERR_CUDA;
    }
    SvtB200Engine *e = new SvtB200Engine();
    e->device = device;
    DeviceGuard g(device);
    for (auto &s : e->me) {
        ENG_TRY(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
        ENG_TRY(cudaEventCreateWithFlags(&s.up_done, cudaEventDisableTiming));
    }
    for (auto &s : e->filt) ENG_TRY(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    for (auto &c : e->planes) ENG_TRY(cudaEventCreateWithFlags(&c.ready, cudaEventDisableTiming));
    *out = e;
    return SVT_B200_OK;
}

void svt_b200_engine_destroy(SvtB200Engine *e) {
    if (!e) return;
    DeviceGuard g(e->device);
    cudaDeviceSynchronize();
    for (auto &s : e->me) {
        if (s.st) cudaStreamDestroy(s.st);
        if (s.up_done) cudaEventDestroy(s.up_done);
    }
    for (auto &s : e->filt)
        if (s.st) cudaStreamDestroy(s.st);
    for (auto &c : e->planes)
        if (c.ready) cudaEventDestroy(c.ready);
    if (e->me_dev) cudaFree(e->me_dev);
    if (e->filt_dev) cudaFree(e->filt_dev);
    if (e->me_pin) cudaFreeHost(e->me_pin);
    if (e->filt_pin) cudaFreeHost(e->filt_pin);
    cudaGetLastError();
    delete e;
}

int svt_b200_engine_get_stats(SvtB200Engine *e, SvtB200EngineStats *out) {
    if (!e || !out) return SVT_B200_ERR_ARG;
    out->me_pictures = e->stats.me_pictures;
    out->dlf_frames = e->stats.dlf_frames;
    out->cdef_frames = e->stats.cdef_frames;
    out->lr_frames = e->stats.lr_frames;
    out->me_plane_uploads = e->stats.me_plane_uploads;
    out->me_plane_hits = e->stats.me_plane_hits;
    out->h2d_bytes = e->stats.h2d_bytes;
    out->d2h_bytes = e->stats.d2h_bytes;
    out->pinned_bytes = e->stats.pinned_bytes;
    out->ns_slot_wait = e->stats.ns_slot_wait;
    out->ns_pin = e->stats.ns_pin;
    out->ns_plane_wait = e->stats.ns_plane_wait;
    out->ns_issue = e->stats.ns_issue;
    out->ns_sync = e->stats.ns_sync;
    out->ns_host_copy = e->stats.ns_host_copy;
    out->pin_calls = e->stats.pin_calls;
    return SVT_B200_OK;
}

int svt_b200_engine_me_picture(SvtB200Engine *e, const SvtB200MeParams *p, const SvtB200HostMePicture *src,
                               const SvtB200HostMePicture refs[SVT_B200_ME_LISTS][SVT_B200_ME_MAX_REFS],
                               int32_t filtered_downsample, int16_t *me_mv, uint8_t *me_cand, uint8_t *total_cand,
                               uint32_t *rc_me_distortion) {
    if (!e || !p || !src || !refs || !me_mv || !me_cand || !total_cand || !rc_me_distortion || !src->full) {
        set_error("svt_b200_engine_me_picture: null argument");
        return SVT_B200_ERR_ARG;
    }
    DeviceGuard dg(e->device);
    int rc = ensure_me(e, p);
    if (rc != SVT_B200_OK) return rc;
    const MeGeom &g = e->mg;
    MeSlot *s = acquire(e, e->me);
    PlaneEntry *ents[1 + SVT_B200_ME_LISTS * SVT_B200_ME_MAX_REFS] = {nullptr};
    int n_ent = 0;
    do {
        SvtB200MePlanes dsrc, drefs[SVT_B200_ME_LISTS][SVT_B200_ME_MAX_REFS];
        memset(drefs, 0, sizeof(drefs));
        {
            Lap lap(e->stats.ns_issue);
            if ((rc = plane_acquire(e, s, p, src, filtered_downsample, &ents[n_ent])) != SVT_B200_OK) break;
            dsrc = {ents[n_ent]->d[0], ents[n_ent]->d[1], ents[n_ent]->d[2]};
            n_ent++;
            for (int l = 0; l < p->num_lists && rc == SVT_B200_OK; l++)
                for (int r = 0; r < p->num_refs[l] && rc == SVT_B200_OK; r++) {
                    if (!refs[l][r].full) {
                        set_error("svt_b200_engine_me_picture: reference [%d][%d] has no planes", l, r);
                        rc = SVT_B200_ERR_ARG;
                        break;
                    }
                    if ((rc = plane_acquire(e, s, p, &refs[l][r], filtered_downsample, &ents[n_ent])) != SVT_B200_OK) break;
                    drefs[l][r] = {ents[n_ent]->d[0], ents[n_ent]->d[1], ents[n_ent]->d[2]};
                    n_ent++;
                }
            if (rc != SVT_B200_OK) break;
            SvtB200MeOutputs o;
            o.best_sad = (uint32_t *)s->dev;
            o.best_mv = (uint32_t *)(s->dev + al256(g.b_sad));
            o.hme = (SvtB200HmeResult *)(s->dev + 2 * al256(g.b_sad));
            o.me_mv = (int16_t *)(s->dev + g.o_mv);
            o.me_cand = s->dev + g.o_cand;
            o.total_cand = s->dev + g.o_tot;
            o.rc_me_distortion = (uint32_t *)(s->dev + g.o_rc);
            if ((rc = svt_b200_me_picture(p, &dsrc, drefs, &o, s->scratch, s->st)) != SVT_B200_OK) break;
            // the four result arrays are contiguous on the device: one read-back
            if (cudaMemcpyAsync(s->pin_out, s->dev + g.o_mv, g.out, cudaMemcpyDeviceToHost, s->st) != cudaSuccess) {
                set_error("engine: ME read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = SVT_B200_ERR_CUDA;
                break;
            }
        }
        if (timed_sync(e, s->st) != cudaSuccess) {
            set_error("engine: ME failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        s->up_pending = false;
        Lap lap(e->stats.ns_host_copy);
        memcpy(me_mv, s->pin_out, g.b_mv);
        memcpy(me_cand, s->pin_out + (g.o_cand - g.o_mv), g.b_cand);
        memcpy(total_cand, s->pin_out + (g.o_tot - g.o_mv), g.b_tot);
        memcpy(rc_me_distortion, s->pin_out + (g.o_rc - g.o_mv), g.b_rc);
        e->stats.me_pictures++;
        e->stats.d2h_bytes += g.out;
    } while (0);
    if (rc != SVT_B200_OK) {
        cudaStreamSynchronize(s->st); // nothing of this picture may still read the planes
        s->up_pending = false;
    }
    plane_release(e, ents, n_ent);
    release(e, s);
    return rc;
}

int svt_b200_engine_dlf_frame(SvtB200Engine *e, const SvtB200DlfParams *p, const SvtB200Frame *frame, const SvtB200DlfMi *mi) {
    if (!e || !p || !frame || !mi || p->mi_stride != p->mi_cols) {
        set_error("svt_b200_engine_dlf_frame: bad argument (mi_stride must equal mi_cols)");
        return SVT_B200_ERR_ARG;
    }
    DeviceGuard dg(e->device);
    int rc = ensure_filt(e, frame->width, frame->height, frame->bit_depth, p->mi_rows, p->mi_cols);
    if (rc != SVT_B200_OK) return rc;
    const FiltGeom &g = e->fg;
    const size_t b_mi = (size_t)p->mi_rows * p->mi_cols * sizeof(SvtB200DlfMi);
    if (b_mi > g.mi) {
        set_error("svt_b200_engine_dlf_frame: mode-info array larger than the sequence geometry");
        return SVT_B200_ERR_ARG;
    }
    FiltSlot *s = acquire(e, e->filt);
    do {
        pack_frame(e, s->pin_a, frame);
        {
            Lap lap(e->stats.ns_host_copy);
            par_memcpy(s->pin_b, mi, b_mi);
        }
        {
            Lap lap(e->stats.ns_issue);
            if (cudaMemcpyAsync(s->misc, s->pin_b, b_mi, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
                set_error("engine: mode-info upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            if ((rc = copy_packed(e, &s->recon, s->pin_a, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
            if ((rc = svt_b200_dlf_frame(p, &s->recon, (const SvtB200DlfMi *)s->misc, s->st)) != SVT_B200_OK) break;
            if ((rc = copy_packed(e, &s->recon, s->pin_a, cudaMemcpyDeviceToHost, s->st)) != SVT_B200_OK) break;
        }
        if (timed_sync(e, s->st) != cudaSuccess) {
            set_error("engine: deblocking failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        unpack_frame(e, frame, s->pin_a);
        e->stats.dlf_frames++;
        e->stats.h2d_bytes += b_mi;
    } while (0);
    if (rc != SVT_B200_OK) cudaStreamSynchronize(s->st);
    release(e, s);
    return rc;
}

int svt_b200_engine_dlf_pick_frame(SvtB200Engine *e, const SvtB200LpfPickParams *pp, const SvtB200Frame *recon,
                                   const SvtB200Frame *source, const SvtB200DlfMi *mi, int32_t levels_out[4]) {
    if (!e || !pp || !recon || !source || !mi || !levels_out || pp->dlf.mi_stride != pp->dlf.mi_cols || recon->width != source->width ||
        recon->height != source->height || recon->bit_depth != source->bit_depth) {
        set_error("svt_b200_engine_dlf_pick_frame: bad argument");
        return SVT_B200_ERR_ARG;
    }
    DeviceGuard dg(e->device);
    int rc = ensure_filt(e, recon->width, recon->height, recon->bit_depth, pp->dlf.mi_rows, pp->dlf.mi_cols);
    if (rc != SVT_B200_OK) return rc;
    const FiltGeom &g = e->fg;
    const size_t b_mi = (size_t)pp->dlf.mi_rows * pp->dlf.mi_cols * sizeof(SvtB200DlfMi);
    if (b_mi > g.mi) {
        set_error("svt_b200_engine_dlf_pick_frame: mode-info array larger than the sequence geometry");
        return SVT_B200_ERR_ARG;
    }
    FiltSlot *s = acquire(e, e->filt);
    do {
        uint8_t *d_scratch = s->misc + g.mi; // 768 B level table + the SSE accumulator of the search, then the final table
        pack_frame(e, s->pin_a, recon);
        {
            Lap lap(e->stats.ns_issue);
            if ((rc = copy_packed(e, &s->recon, s->pin_a, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
        }
        // pin_b holds the mode-info array first, then (after its upload completed) the packed source picture
        {
            Lap lap(e->stats.ns_host_copy);
            par_memcpy(s->pin_b, mi, b_mi);
        }
        if (cudaMemcpyAsync(s->misc, s->pin_b, b_mi, cudaMemcpyHostToDevice, s->st) != cudaSuccess || timed_sync(e, s->st) != cudaSuccess) {
            set_error("engine: mode-info upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        pack_frame(e, s->pin_b, source);
        if ((rc = copy_packed(e, &s->source, s->pin_b, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
        // every trial: one plane pass + SSE + plane restore on the device, one 8-byte read-back (the bisection is host logic)
        if ((rc = svt_b200_pick_filter_level(pp, &s->recon, &s->source, &s->out, (const SvtB200DlfMi *)s->misc, d_scratch, levels_out,
                                             s->st)) != SVT_B200_OK)
            break;
        SvtB200DlfParams dp = pp->dlf;
        dp.sharpness = 0; // svt_av1_pick_filter_level sets lf->sharpness_level = 0 (EbDeblockingFilter.c:1202)
        dp.filter_level[0] = levels_out[0], dp.filter_level[1] = levels_out[1];
        dp.filter_level_u = levels_out[2], dp.filter_level_v = levels_out[3];
        dp.plane_start = 0, dp.plane_end = 3;
        uint8_t(*lut)[2][128] = reinterpret_cast<uint8_t(*)[2][128]>(s->pin_small);
        svt_b200_lf_level_lut(&pp->init, levels_out, lut);
        {
            Lap lap(e->stats.ns_issue);
            if (cudaMemcpyAsync(d_scratch + 1024, s->pin_small, 768, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            if ((rc = svt_b200_dlf_frame_lut(&dp, &s->recon, (const SvtB200DlfMi *)s->misc, d_scratch + 1024, s->st)) != SVT_B200_OK) break;
            if ((rc = copy_packed(e, &s->recon, s->pin_a, cudaMemcpyDeviceToHost, s->st)) != SVT_B200_OK) break;
        }
        if (timed_sync(e, s->st) != cudaSuccess) {
            set_error("engine: deblocking (picked levels) failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        unpack_frame(e, recon, s->pin_a);
        e->stats.dlf_frames++;
        e->stats.h2d_bytes += b_mi;
    } while (0);
    if (rc != SVT_B200_OK) cudaStreamSynchronize(s->st);
    release(e, s);
    return rc;
}

int svt_b200_engine_lr_frame(SvtB200Engine *e, const SvtB200LrFrameParams *p, const int32_t n_units[3], const SvtB200Frame *frame,
                             const SvtB200HostLrLines lines[3]) {
    if (!e || !p || !n_units || !frame || !lines) {
        set_error("svt_b200_engine_lr_frame: null argument");
        return SVT_B200_ERR_ARG;
    }
    DeviceGuard dg(e->device);
    int rc = ensure_filt(e, frame->width, frame->height, frame->bit_depth, (frame->height + 3) / 4, (frame->width + 3) / 4);
    if (rc != SVT_B200_OK) return rc;
    const FiltGeom &g = e->fg;
    const int bps = frame->bit_depth > 8 ? 2 : 1;
    size_t b_units = 0;
    for (int i = 0; i < 3; i++) b_units += al256((size_t)std::max(n_units[i], 0) * sizeof(SvtB200LrUnit));
    // staging of the boundary lines: 4 rows per stripe boundary and plane
    size_t b_lines = 0;
    int nb[3], pw[3], ph[3], SH[3], off[3];
    for (int i = 0; i < 3; i++) {
        const int ss = i ? 1 : 0;
        pw[i] = i ? (frame->width + 1) >> 1 : frame->width;
        ph[i] = i ? (frame->height + 1) >> 1 : frame->height;
        SH[i] = 64 >> ss, off[i] = 8 >> ss;
        const int n_stripes = (ph[i] + off[i] + SH[i] - 1) / SH[i];
        nb[i] = p->plane[i].frame_restoration_type ? n_stripes - 1 : 0;
        b_lines += al256((size_t)nb[i] * 4 * pw[i] * bps);
    }
    if (b_units + 256 > g.misc || b_lines > g.packed) {
        set_error("svt_b200_engine_lr_frame: unit / boundary-line arrays larger than the sequence geometry");
        return SVT_B200_ERR_ARG;
    }
    FiltSlot *s = acquire(e, e->filt);
    do {
        pack_frame(e, s->pin_a, frame);
        // units -> pin_small (host) -> misc (device)
        SvtB200LrFrameParams dp = *p;
        {
            size_t o = 0;
            if (b_units > g.small) {
                set_error("svt_b200_engine_lr_frame: too many restoration units");
                rc = SVT_B200_ERR_ARG;
                break;
            }
            for (int i = 0; i < 3; i++) {
                const size_t n = (size_t)std::max(n_units[i], 0) * sizeof(SvtB200LrUnit);
                if (n) memcpy(s->pin_small + o, p->plane[i].units, n);
                dp.plane[i].units = (const SvtB200LrUnit *)(s->misc + o);
                o += al256(n);
            }
        }
        // boundary lines: [boundary][row -2, -1, 0, +1 around the stripe edge][width], packed per plane in pin_b
        {
            Lap lap(e->stats.ns_host_copy);
            size_t o = 0;
            for (int i = 0; i < 3; i++) {
                const size_t rb = (size_t)pw[i] * bps, ls = (size_t)lines[i].stride * bps;
                for (int b = 0; b < nb[i]; b++) {
                    uint8_t *d = s->pin_b + o + (size_t)b * 4 * rb;
                    const uint8_t *ab = (const uint8_t *)lines[i].above + (size_t)(2 * (b + 1)) * ls; // stripe b+1: rows y0-2, y0-1
                    const uint8_t *be = (const uint8_t *)lines[i].below + (size_t)(2 * b) * ls;       // stripe b: rows y1, y1+1
                    memcpy(d, ab, rb);
                    memcpy(d + rb, ab + ls, rb);
                    memcpy(d + 2 * rb, be, rb);
                    memcpy(d + 3 * rb, be + ls, rb);
                }
                o += al256((size_t)nb[i] * 4 * rb);
            }
        }
        {
            Lap lap(e->stats.ns_issue);
            if (b_units && cudaMemcpyAsync(s->misc, s->pin_small, b_units, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            if ((rc = copy_packed(e, &s->recon, s->pin_a, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
            // scatter the boundary rows into a device picture that is only read at those rows (the deblocked context)
            size_t o = 0;
            bool ok = true;
            for (int i = 0; i < 3 && ok; i++) {
                if (!nb[i]) continue;
                const size_t rb = (size_t)pw[i] * bps;
                uint8_t *plane = (uint8_t *)(i == 0 ? s->source.y : i == 1 ? s->source.cb : s->source.cr);
                const size_t dpitch = (size_t)(i ? s->source.stride_c : s->source.stride_y) * bps;
                for (int k = 0; k < 4 && ok; k++) { // row B - 2 + k of every boundary B = (b + 1) * SH - off
                    int rows = nb[i];
                    if ((nb[i]) * SH[i] - off[i] - 2 + k >= ph[i]) rows--; // the last boundary's row lies below the plane
                    if (rows <= 0) continue;
                    ok = cudaMemcpy2DAsync(plane + (size_t)(SH[i] - off[i] - 2 + k) * dpitch, dpitch * SH[i], s->pin_b + o + (size_t)k * rb,
                                           4 * rb, rb, rows, cudaMemcpyHostToDevice, s->st) == cudaSuccess;
                }
                o += al256((size_t)nb[i] * 4 * rb);
            }
            if (!ok) {
                set_error("engine: boundary-line upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            if ((rc = svt_b200_lr_frame(&dp, &s->recon, &s->source, &s->out, s->st)) != SVT_B200_OK) break;
            if ((rc = copy_packed(e, &s->out, s->pin_a, cudaMemcpyDeviceToHost, s->st)) != SVT_B200_OK) break;
        }
        if (timed_sync(e, s->st) != cudaSuccess) {
            set_error("engine: loop restoration failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        // planes without restoration stay as they are on the host (the reference copies back only filtered planes)
        SvtB200Frame fr = *frame;
        {
            Lap lap(e->stats.ns_host_copy);
            const int cw = pw[1], ch = ph[1];
            const size_t rw = (size_t)frame->width * bps, rcb = (size_t)cw * bps;
            const uint8_t *py = s->pin_a, *pcb = py + rw * frame->height, *pcr = pcb + rcb * ch;
            if (p->plane[0].frame_restoration_type) {
                const RowCopy c[3] = {{(uint8_t *)fr.y, py, (size_t)fr.stride_y * bps, rw, rw, fr.height}, {nullptr, nullptr, 0, 0, 0, 0}, {nullptr, nullptr, 0, 0, 0, 0}};
                run_copies(e, c);
            }
            if (p->plane[1].frame_restoration_type)
                for (int y = 0; y < ch; y++) memcpy((uint8_t *)fr.cb + (size_t)y * fr.stride_c * bps, pcb + y * rcb, rcb);
            if (p->plane[2].frame_restoration_type)
                for (int y = 0; y < ch; y++) memcpy((uint8_t *)fr.cr + (size_t)y * fr.stride_c * bps, pcr + y * rcb, rcb);
        }
        e->stats.lr_frames++;
    } while (0);
    if (rc != SVT_B200_OK) cudaStreamSynchronize(s->st);
    release(e, s);
    return rc;
}

int svt_b200_engine_cdef_frame(SvtB200Engine *e, const SvtB200CdefSearchParams *sp, const SvtB200Frame *recon,
                               const SvtB200Frame *source, const uint8_t *skip8, int32_t skip_stride, uint64_t *mse,
                               SvtB200CdefDecideFn decide, void *user) {
    return svt_b200_engine_dlf_cdef_frame(e, nullptr, nullptr, sp, recon, source, skip8, skip_stride, mse, decide, user);
}

int svt_b200_engine_dlf_cdef_frame(SvtB200Engine *e, const SvtB200DlfParams *dlf, const SvtB200DlfMi *mi,
                                   const SvtB200CdefSearchParams *sp, const SvtB200Frame *recon, const SvtB200Frame *source,
                                   const uint8_t *skip8, int32_t skip_stride, uint64_t *mse, SvtB200CdefDecideFn decide, void *user) {
    if (!e || !sp || !recon || !source || !skip8 || !mse || !decide || recon->width != source->width ||
        recon->height != source->height || recon->bit_depth != source->bit_depth || (dlf && (!mi || dlf->mi_stride != dlf->mi_cols))) {
        set_error("svt_b200_engine_dlf_cdef_frame: bad argument");
        return SVT_B200_ERR_ARG;
    }
    DeviceGuard dg(e->device);
    int rc = ensure_filt(e, recon->width, recon->height, recon->bit_depth, sp->mi_rows, sp->mi_cols);
    if (rc != SVT_B200_OK) return rc;
    const FiltGeom &g = e->fg;
    const int nvfb = (sp->mi_rows + 15) / 16, nhfb = (sp->mi_cols + 15) / 16, nfb = nvfb * nhfb;
    const size_t b_skip = (size_t)((sp->mi_rows + 1) / 2) * skip_stride, b_mse = (size_t)2 * nfb * 64 * 8;
    if (al256(b_skip) > g.skip || (size_t)nfb > g.nfb) {
        set_error("svt_b200_engine_cdef_frame: skip map / filter-block count larger than the sequence geometry");
        return SVT_B200_ERR_ARG;
    }
    const size_t b_mi = dlf ? (size_t)dlf->mi_rows * dlf->mi_cols * sizeof(SvtB200DlfMi) : 0;
    if (b_mi > g.mi) {
        set_error("svt_b200_engine_dlf_cdef_frame: mode-info array larger than the sequence geometry");
        return SVT_B200_ERR_ARG;
    }
    FiltSlot *s = acquire(e, e->filt);
    do {
        uint8_t *d_small = s->misc + g.mi + 4096; // the mode-info summary (deblocking) occupies the head of misc
        uint8_t *d_skip = d_small, *d_mse = d_small + g.skip, *d_idx = d_mse + g.mse;
        uint8_t *h_skip = s->pin_small, *h_mse = h_skip + g.skip, *h_idx = h_mse + g.mse;
        pack_frame(e, s->pin_a, recon);
        if (dlf) { // the deblocking the DLF stage deferred: same upload of the reconstruction serves both filters
            {
                Lap lap(e->stats.ns_host_copy);
                par_memcpy(s->pin_b, mi, b_mi);
            }
            Lap lap(e->stats.ns_issue);
            if (cudaMemcpyAsync(s->misc, s->pin_b, b_mi, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            if ((rc = copy_packed(e, &s->recon, s->pin_a, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
            if ((rc = svt_b200_dlf_frame(dlf, &s->recon, (const SvtB200DlfMi *)s->misc, s->st)) != SVT_B200_OK) break;
            if (timed_sync(e, s->st) != cudaSuccess) { // pin_b is reused for the source picture below
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            e->stats.dlf_frames++;
            e->stats.h2d_bytes += b_mi;
        }
        pack_frame(e, s->pin_b, source);
        memcpy(h_skip, skip8, b_skip);
        {
            Lap lap(e->stats.ns_issue);
            if (cudaMemcpyAsync(d_skip, h_skip, b_skip, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
                set_error("engine: skip-map upload failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            if (!dlf && (rc = copy_packed(e, &s->recon, s->pin_a, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
            if ((rc = copy_packed(e, &s->source, s->pin_b, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
            if ((rc = svt_b200_cdef_search(sp, &s->recon, &s->source, d_skip, skip_stride, (uint64_t *)d_mse, s->st)) != SVT_B200_OK)
                break;
            if (cudaMemcpyAsync(h_mse, d_mse, b_mse, cudaMemcpyDeviceToHost, s->st) != cudaSuccess) {
                rc = SVT_B200_ERR_CUDA;
                break;
            }
        }
        if (timed_sync(e, s->st) != cudaSuccess) {
            set_error("engine: CDEF search failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        memcpy(mse, h_mse, b_mse);
        // the strength decision (finish_cdef_search, EbEncCdef.c:1167) is the host's: it reads mse and fills the apply set
        SvtB200CdefApplyParams ap;
        memset(&ap, 0, sizeof(ap));
        ap.mi_rows = sp->mi_rows;
        ap.mi_cols = sp->mi_cols;
        int8_t *idx = (int8_t *)h_idx;
        memset(idx, -1, (size_t)nfb);
        const int apply = decide(user, mse, &ap, idx);
        e->stats.cdef_frames++;
        e->stats.h2d_bytes += b_skip;
        e->stats.d2h_bytes += b_mse;
        if (apply <= 0) {
            if (apply < 0) {
                set_error("svt_b200_engine_cdef_frame: the strength-decision callback failed");
                rc = SVT_B200_ERR_ARG;
            }
            break;
        }
        {
            Lap lap(e->stats.ns_issue);
            if (cudaMemcpyAsync(d_idx, idx, (size_t)nfb, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            if ((rc = svt_b200_cdef_apply(&ap, &s->recon, &s->out, d_skip, skip_stride, (const int8_t *)d_idx, s->st)) != SVT_B200_OK)
                break;
            if ((rc = copy_packed(e, &s->out, s->pin_a, cudaMemcpyDeviceToHost, s->st)) != SVT_B200_OK) break;
        }
        if (timed_sync(e, s->st) != cudaSuccess) {
            set_error("engine: CDEF apply failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        unpack_frame(e, recon, s->pin_a);
    } while (0);
    if (rc != SVT_B200_OK) cudaStreamSynchronize(s->st);
    release(e, s);
    return rc;
}

// The same picture pass with the strength decision on the DEVICE (svt_b200_cdef_decide replaces the host callback): one
// upload, deblocking, CDEF search, decision, CDEF apply, one download, ONE synchronisation - the reconstruction and the
// decision arrive together.  apply: 0 = search + decision only (EbCdefProcess.c:527-529: nobody reads the filtered picture).
int svt_b200_engine_dlf_cdef_frame_dev(SvtB200Engine *e, const SvtB200DlfParams *dlf, const SvtB200DlfMi *mi,
                                       const SvtB200CdefSearchParams *sp, const SvtB200CdefDecideParams *dp, int32_t damping,
                                       int32_t apply, const SvtB200Frame *recon, const SvtB200Frame *source, const uint8_t *skip8,
                                       int32_t skip_stride, SvtB200CdefDecision *decision, int8_t *fb_strength_idx) {
    if (!e || !sp || !dp || !recon || !source || !skip8 || !decision || !fb_strength_idx || recon->width != source->width ||
        recon->height != source->height || recon->bit_depth != source->bit_depth || (dlf && (!mi || dlf->mi_stride != dlf->mi_cols)) ||
        dp->mi_rows != sp->mi_rows || dp->mi_cols != sp->mi_cols || dp->n_strengths != sp->n_strengths) {
        set_error("svt_b200_engine_dlf_cdef_frame_dev: bad argument");
        return SVT_B200_ERR_ARG;
    }
    DeviceGuard dg(e->device);
    int rc = ensure_filt(e, recon->width, recon->height, recon->bit_depth, sp->mi_rows, sp->mi_cols);
    if (rc != SVT_B200_OK) return rc;
    const FiltGeom &g = e->fg;
    const int nvfb = (sp->mi_rows + 15) / 16, nhfb = (sp->mi_cols + 15) / 16, nfb = nvfb * nhfb;
    const size_t b_skip = (size_t)((sp->mi_rows + 1) / 2) * skip_stride;
    const size_t b_mi = dlf ? (size_t)dlf->mi_rows * dlf->mi_cols * sizeof(SvtB200DlfMi) : 0;
    if (al256(b_skip) > g.skip || (size_t)nfb > g.nfb || b_mi > g.mi) {
        set_error("svt_b200_engine_dlf_cdef_frame_dev: picture larger than the sequence geometry");
        return SVT_B200_ERR_ARG;
    }
    FiltSlot *s = acquire(e, e->filt);
    do {
        uint8_t *d_small = s->misc + g.mi + 4096;
        uint8_t *d_skip = d_small, *d_mse = d_small + g.skip, *d_idx = d_mse + g.mse, *d_dec = d_idx + g.idx;
        uint8_t *d_scr = d_dec + al256(sizeof(SvtB200CdefDecision));
        uint8_t *h_skip = s->pin_small, *h_idx = h_skip + g.skip + g.mse;
        pack_frame(e, s->pin_a, recon);
        {
            Lap lap(e->stats.ns_host_copy);
            if (dlf) par_memcpy(s->pin_mi, mi, b_mi);
            memcpy(h_skip, skip8, b_skip);
        }
        {
            Lap lap(e->stats.ns_issue);
            if (dlf && cudaMemcpyAsync(s->misc, s->pin_mi, b_mi, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            if ((rc = copy_packed(e, &s->recon, s->pin_a, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
            if (dlf && (rc = svt_b200_dlf_frame(dlf, &s->recon, (const SvtB200DlfMi *)s->misc, s->st)) != SVT_B200_OK) break;
        }
        pack_frame(e, s->pin_b, source); // the host packs the source while the device receives and deblocks the reconstruction
        {
            Lap lap(e->stats.ns_issue);
            if (cudaMemcpyAsync(d_skip, h_skip, b_skip, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
                rc = SVT_B200_ERR_CUDA;
                break;
            }
            if ((rc = copy_packed(e, &s->source, s->pin_b, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
            if ((rc = svt_b200_cdef_search(sp, &s->recon, &s->source, d_skip, skip_stride, (uint64_t *)d_mse, s->st)) != SVT_B200_OK)
                break;
            if ((rc = svt_b200_cdef_decide(dp, (const uint64_t *)d_mse, d_skip, skip_stride, (SvtB200CdefDecision *)d_dec, (int8_t *)d_idx,
                                           d_scr, s->st)) != SVT_B200_OK)
                break;
            if (apply) {
                if ((rc = svt_b200_cdef_apply_dev(sp->mi_rows, sp->mi_cols, damping, (const SvtB200CdefDecision *)d_dec, &s->recon, &s->out,
                                                  d_skip, skip_stride, (const int8_t *)d_idx, s->st)) != SVT_B200_OK)
                    break;
                if ((rc = copy_packed(e, &s->out, s->pin_a, cudaMemcpyDeviceToHost, s->st)) != SVT_B200_OK) break;
            }
            // filter-block indices and the decision are adjacent on both sides: one copy
            if (cudaMemcpyAsync(h_idx, d_idx, g.idx + sizeof(SvtB200CdefDecision), cudaMemcpyDeviceToHost, s->st) != cudaSuccess) {
                rc = SVT_B200_ERR_CUDA;
                break;
            }
        }
        if (timed_sync(e, s->st) != cudaSuccess) {
            set_error("engine: deblocking + CDEF failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        memcpy(fb_strength_idx, h_idx, (size_t)nfb);
        memcpy(decision, h_idx + g.idx, sizeof(SvtB200CdefDecision));
        if (apply) unpack_frame(e, recon, s->pin_a);
        if (dlf) e->stats.dlf_frames++;
        e->stats.cdef_frames++;
        e->stats.h2d_bytes += b_skip + b_mi;
        e->stats.d2h_bytes += g.idx + sizeof(SvtB200CdefDecision);
    } while (0);
    if (rc != SVT_B200_OK) cudaStreamSynchronize(s->st);
    release(e, s);
    return rc;
}

} // extern "C"
