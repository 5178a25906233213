// engine.cu — the host-buffer picture pipeline of libsvtav1_b200.so (include/svt_av1_b200.h, "picture engine").
//
// The picture-level entries of the other files take DEVICE pointers and a stream.  The reference's process loops
// (motion_estimation_kernel, dlf_kernel, cdef_kernel; SURVEY.md §8b "Batched entry" / "Memory ownership") own HOST
// pictures and call from many pipeline threads at once, one picture per call.  The engine is the piece in between,
// inside the product library so that a C host gets it by linking:
//   * a residency cache of the three padded luma planes of every picture that ME touches (as source or as reference),
//     keyed by (host object, picture number): a picture is uploaded once and then serves as reference for the pictures
//     that follow it (EbPaReferenceObject lifetime, EbPictureBufferDesc.c:65-78 layout);
//   * per-call slots (stream + device scratch/outputs + pinned result staging) so N threads run N pictures at once;
//   * lazily page-locking (cudaHostRegister) the host planes it is handed, so uploads / read-backs are real DMA;
//   * the deblock -> CDEF search -> (host strength decision) -> CDEF apply chain of one picture with the
//     reconstruction resident on the device across the stages.
// Every entry is synchronous for its caller (the reference's stage returns when its picture is done) and re-entrant.
// There is no CPU fallback: a CUDA failure is returned as SVT_B200_ERR_CUDA and the integration aborts the encode.
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

using namespace svtb200;

namespace {

constexpr int kMeSlots = 8;      // concurrent ME pictures
constexpr int kFiltSlots = 6;    // concurrent deblock / CDEF pictures
constexpr int kPlaneEntries = 96; // resident ME pictures (3.3 MB each at 1080p, 13 MB at 2160p); >= kMeSlots * 9 so that
                                  // every in-flight picture can hold its source + 8 references at once (no deadlock)

struct PlaneEntry {
    const void *key = nullptr;
    uint64_t tag = 0;
    uint8_t *d[3] = {nullptr, nullptr, nullptr};
    size_t cap[3] = {0, 0, 0};
    cudaEvent_t ready = nullptr;
    int state = 0; // 0 free, 1 loading, 2 ready
    int users = 0;
    uint64_t stamp = 0;
};

struct MeSlot {
    bool busy = false;
    cudaStream_t st = nullptr;
    void *scratch = nullptr;
    size_t scratch_cap = 0;
    uint8_t *dev = nullptr; // best_sad | best_mv | hme | me_mv | me_cand | total | rc
    size_t dev_cap = 0;
    uint8_t *pin = nullptr; // me_mv | me_cand | total | rc
    size_t pin_cap = 0;
};

struct DevFrame {
    uint8_t *base = nullptr;
    size_t cap = 0;
    SvtB200Frame f;
};

struct FiltSlot {
    bool busy = false;
    cudaStream_t st = nullptr;
    DevFrame recon, out, source;
    uint8_t *misc = nullptr; // mi array | skip8 | mse | fb idx
    size_t misc_cap = 0;
    uint8_t *pin = nullptr;
    size_t pin_cap = 0;
};

} // namespace

struct SvtB200Engine {
    int device = 0;
    bool pin_host = true;
    std::mutex mu;
    std::condition_variable cv;
    uint64_t clock = 0;
    PlaneEntry planes[kPlaneEntries];
    MeSlot me[kMeSlots];
    FiltSlot filt[kFiltSlots];
    std::mutex pin_mu;
    std::unordered_map<const void *, size_t> pinned; // base -> bytes (0: registration failed, do not retry)
    struct {
        std::atomic<uint64_t> me_pictures{0}, dlf_frames{0}, cdef_frames{0}, me_plane_uploads{0}, me_plane_hits{0}, h2d_bytes{0},
            d2h_bytes{0}, pinned_bytes{0};
    } stats;
};

namespace {

#define ENG_TRY(expr)                                                                                              \
    do {                                                                                                           \
        cudaError_t e__ = (expr);                                                                                  \
        if (e__ != cudaSuccess) {                                                                                  \
            set_error("engine: %s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);        \
            return SVT_B200_ERR_CUDA;                                                                              \
        }                                                                                                          \
    } while (0)

int grow_dev(uint8_t **p, size_t *cap, size_t need) {
    if (need <= *cap) return SVT_B200_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    const size_t n = (need + 0xffff) & ~(size_t)0xffff;
    ENG_TRY(cudaMalloc((void **)p, n));
    *cap = n;
    return SVT_B200_OK;
}
int grow_pin(uint8_t **p, size_t *cap, size_t need) {
    if (need <= *cap) return SVT_B200_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    const size_t n = (need + 0xffff) & ~(size_t)0xffff;
    ENG_TRY(cudaMallocHost((void **)p, n));
    *cap = n;
    return SVT_B200_OK;
}

// Page-lock a host range once (the reference allocates its picture buffers once per encoder and recycles them).
void pin_range(SvtB200Engine *e, const void *base, size_t bytes) {
    if (!e->pin_host || !base || !bytes) return;
    std::lock_guard<std::mutex> g(e->pin_mu);
    auto it = e->pinned.find(base);
    if (it != e->pinned.end() && (it->second == 0 || it->second >= bytes)) return;
    if (it != e->pinned.end()) cudaHostUnregister((void *)base);
    cudaError_t r = cudaHostRegister((void *)base, bytes, cudaHostRegisterDefault);
    if (r != cudaSuccess) {
        cudaGetLastError(); // pageable copies still work; remember not to retry
        e->pinned[base] = 0;
        return;
    }
    e->pinned[base] = bytes;
    e->stats.pinned_bytes += bytes;
}

template <typename Slot, int N>
Slot *acquire(SvtB200Engine *e, Slot (&pool)[N]) {
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

size_t plane_bytes(const SvtB200Plane &g) { return (size_t)g.stride * (size_t)(g.height + 2 * g.origin_y); }

// Make the three planes of `pic` resident and ordered before later work on `st`.  Returns the entry with users++.
int plane_acquire(SvtB200Engine *e, const SvtB200MeParams *p, const SvtB200HostMePicture *pic, int filtered_ds,
                  cudaStream_t st, PlaneEntry **out) {
    const size_t nb[3] = {plane_bytes(p->full), plane_bytes(p->quarter), plane_bytes(p->sixteenth)};
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
                if (c.users == 0 && c.state != 1 && (!victim || c.state < victim->state ||
                                                     (c.state == victim->state && c.stamp < victim->stamp)))
                    victim = &c;
            }
            if (en) {
                en->users++;
                while (en->state == 1) e->cv.wait(lk);
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
        const uint8_t *host[3] = {pic->full, pic->quarter, pic->sixteenth};
        for (int i = 0; i < 3 && rc == SVT_B200_OK; i++) rc = grow_dev(&en->d[i], &en->cap[i], nb[i]);
        if (rc == SVT_B200_OK && !en->ready && cudaEventCreateWithFlags(&en->ready, cudaEventDisableTiming) != cudaSuccess)
            rc = SVT_B200_ERR_CUDA;
        const bool gen = !pic->quarter || !pic->sixteenth; // decimations derived on the device
        for (int i = 0; i < (gen ? 1 : 3) && rc == SVT_B200_OK; i++) {
            pin_range(e, host[i], nb[i]);
            if (cudaMemcpyAsync(en->d[i], host[i], nb[i], cudaMemcpyHostToDevice, st) != cudaSuccess) rc = SVT_B200_ERR_CUDA;
            e->stats.h2d_bytes += nb[i];
        }
        if (rc == SVT_B200_OK && gen) {
            SvtB200MePlanes dp = {en->d[0], en->d[1], en->d[2]};
            rc = svt_b200_me_downsample(&p->full, &p->quarter, &p->sixteenth, &dp, filtered_ds, st);
        }
        if (rc == SVT_B200_OK && cudaEventRecord(en->ready, st) != cudaSuccess) rc = SVT_B200_ERR_CUDA;
        {
            std::lock_guard<std::mutex> g(e->mu);
            en->state = rc == SVT_B200_OK ? 2 : 0;
            if (rc != SVT_B200_OK) en->users--;
        }
        e->cv.notify_all();
        if (rc != SVT_B200_OK) {
            if (rc == SVT_B200_ERR_CUDA) set_error("engine: ME plane upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            return rc;
        }
    } else {
        ENG_TRY(cudaStreamWaitEvent(st, en->ready, 0));
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

size_t frame_plane_bytes(int w, int h, int bd, int *stride) {
    const int bps = bd > 8 ? 2 : 1;
    *stride = (w + 63) & ~63; // samples; 64-sample multiple keeps every row 64-byte (8-bit) / 128-byte aligned
    return (size_t)*stride * h * bps;
}

// Device picture of the geometry of `h` (luma w x h, 4:2:0).
int dev_frame(DevFrame *d, const SvtB200Frame *h) {
    int sy, sc;
    const int cw = (h->width + 1) >> 1, ch = (h->height + 1) >> 1;
    const size_t by = frame_plane_bytes(h->width, h->height, h->bit_depth, &sy);
    const size_t bc = frame_plane_bytes(cw, ch, h->bit_depth, &sc);
    const size_t pad = 256;
    int rc = grow_dev(&d->base, &d->cap, by + 2 * bc + 4 * pad);
    if (rc != SVT_B200_OK) return rc;
    d->f = *h;
    d->f.y = d->base + pad;
    d->f.cb = d->base + 2 * pad + by;
    d->f.cr = d->base + 3 * pad + by + bc;
    d->f.stride_y = sy;
    d->f.stride_c = sc;
    return SVT_B200_OK;
}

int copy_frame(SvtB200Engine *e, const SvtB200Frame *dst, const SvtB200Frame *src, cudaMemcpyKind kind, cudaStream_t st) {
    const int bps = src->bit_depth > 8 ? 2 : 1;
    const int cw = (src->width + 1) >> 1, ch = (src->height + 1) >> 1;
    const SvtB200Frame *host = kind == cudaMemcpyHostToDevice ? src : dst;
    pin_range(e, host->y, ((size_t)host->stride_y * (host->height - 1) + host->width) * bps);
    pin_range(e, host->cb, ((size_t)host->stride_c * (ch - 1) + cw) * bps);
    pin_range(e, host->cr, ((size_t)host->stride_c * (ch - 1) + cw) * bps);
    ENG_TRY(cudaMemcpy2DAsync(dst->y, (size_t)dst->stride_y * bps, src->y, (size_t)src->stride_y * bps, (size_t)src->width * bps,
                              src->height, kind, st));
    ENG_TRY(cudaMemcpy2DAsync(dst->cb, (size_t)dst->stride_c * bps, src->cb, (size_t)src->stride_c * bps, (size_t)cw * bps, ch, kind, st));
    ENG_TRY(cudaMemcpy2DAsync(dst->cr, (size_t)dst->stride_c * bps, src->cr, (size_t)src->stride_c * bps, (size_t)cw * bps, ch, kind, st));
    const size_t n = ((size_t)src->width * src->height + 2 * (size_t)cw * ch) * bps;
    if (kind == cudaMemcpyHostToDevice)
        e->stats.h2d_bytes += n;
    else
        e->stats.d2h_bytes += n;
    return SVT_B200_OK;
}

struct DeviceGuard { // engine calls come from arbitrary pipeline threads: bind the engine's device for the call
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
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
    const char *pin = getenv("SVT_B200_PIN_HOST");
    e->pin_host = !(pin && pin[0] == '0');
    DeviceGuard g(device);
    for (auto &s : e->me) ENG_TRY(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    for (auto &s : e->filt) ENG_TRY(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    *out = e;
    return SVT_B200_OK;
}

void svt_b200_engine_destroy(SvtB200Engine *e) {
    if (!e) return;
    DeviceGuard g(e->device);
    cudaDeviceSynchronize();
    for (auto &s : e->me) {
        if (s.st) cudaStreamDestroy(s.st);
        if (s.scratch) cudaFree(s.scratch);
        if (s.dev) cudaFree(s.dev);
        if (s.pin) cudaFreeHost(s.pin);
    }
    for (auto &s : e->filt) {
        if (s.st) cudaStreamDestroy(s.st);
        for (DevFrame *d : {&s.recon, &s.out, &s.source})
            if (d->base) cudaFree(d->base);
        if (s.misc) cudaFree(s.misc);
        if (s.pin) cudaFreeHost(s.pin);
    }
    for (auto &c : e->planes) {
        for (int i = 0; i < 3; i++)
            if (c.d[i]) cudaFree(c.d[i]);
        if (c.ready) cudaEventDestroy(c.ready);
    }
    for (auto &kv : e->pinned)
        if (kv.second) cudaHostUnregister((void *)kv.first);
    cudaGetLastError();
    delete e;
}

int svt_b200_engine_get_stats(SvtB200Engine *e, SvtB200EngineStats *out) {
    if (!e || !out) return SVT_B200_ERR_ARG;
    out->me_pictures = e->stats.me_pictures;
    out->dlf_frames = e->stats.dlf_frames;
    out->cdef_frames = e->stats.cdef_frames;
    out->me_plane_uploads = e->stats.me_plane_uploads;
    out->me_plane_hits = e->stats.me_plane_hits;
    out->h2d_bytes = e->stats.h2d_bytes;
    out->d2h_bytes = e->stats.d2h_bytes;
    out->pinned_bytes = e->stats.pinned_bytes;
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
    const size_t n_sb = (size_t)((p->full.width + 63) / 64) * ((p->full.height + 63) / 64);
    const size_t b_sad = n_sb * 8 * 85 * 4, b_hme = n_sb * 8 * sizeof(SvtB200HmeResult);
    const size_t b_mv = n_sb * 85 * 7 * 4, b_cand = n_sb * 85 * 23, b_tot = n_sb * 85, b_rc = n_sb * 4;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    MeSlot *s = acquire(e, e->me);
    int rc = SVT_B200_OK;
    PlaneEntry *ents[1 + SVT_B200_ME_LISTS * SVT_B200_ME_MAX_REFS] = {nullptr};
    int n_ent = 0;
    do {
        const size_t need_scratch = svt_b200_me_scratch_bytes(p);
        if ((rc = grow_dev((uint8_t **)&s->scratch, &s->scratch_cap, need_scratch)) != SVT_B200_OK) break;
        const size_t o_mv = 2 * al(b_sad) + al(b_hme), o_cand = o_mv + al(b_mv), o_tot = o_cand + al(b_cand), o_rc = o_tot + al(b_tot);
        if ((rc = grow_dev(&s->dev, &s->dev_cap, o_rc + al(b_rc))) != SVT_B200_OK) break;
        const size_t out_bytes = al(b_mv) + al(b_cand) + al(b_tot) + al(b_rc);
        if ((rc = grow_pin(&s->pin, &s->pin_cap, out_bytes)) != SVT_B200_OK) break;
        SvtB200MePlanes dsrc, drefs[SVT_B200_ME_LISTS][SVT_B200_ME_MAX_REFS];
        memset(drefs, 0, sizeof(drefs));
        if ((rc = plane_acquire(e, p, src, filtered_downsample, s->st, &ents[n_ent])) != SVT_B200_OK) break;
        dsrc = {ents[n_ent]->d[0], ents[n_ent]->d[1], ents[n_ent]->d[2]};
        n_ent++;
        for (int l = 0; l < p->num_lists && rc == SVT_B200_OK; l++)
            for (int r = 0; r < p->num_refs[l] && rc == SVT_B200_OK; r++) {
                if (!refs[l][r].full) {
                    set_error("svt_b200_engine_me_picture: reference [%d][%d] has no planes", l, r);
                    rc = SVT_B200_ERR_ARG;
                    break;
                }
                if ((rc = plane_acquire(e, p, &refs[l][r], filtered_downsample, s->st, &ents[n_ent])) != SVT_B200_OK) break;
                drefs[l][r] = {ents[n_ent]->d[0], ents[n_ent]->d[1], ents[n_ent]->d[2]};
                n_ent++;
            }
        if (rc != SVT_B200_OK) break;
        SvtB200MeOutputs o;
        o.best_sad = (uint32_t *)s->dev;
        o.best_mv = (uint32_t *)(s->dev + al(b_sad));
        o.hme = (SvtB200HmeResult *)(s->dev + 2 * al(b_sad));
        o.me_mv = (int16_t *)(s->dev + o_mv);
        o.me_cand = s->dev + o_cand;
        o.total_cand = s->dev + o_tot;
        o.rc_me_distortion = (uint32_t *)(s->dev + o_rc);
        if ((rc = svt_b200_me_picture(p, &dsrc, drefs, &o, s->scratch, s->st)) != SVT_B200_OK) break;
        // the four result arrays are contiguous on the device: one read-back
        if (cudaMemcpyAsync(s->pin, s->dev + o_mv, out_bytes, cudaMemcpyDeviceToHost, s->st) != cudaSuccess ||
            cudaStreamSynchronize(s->st) != cudaSuccess) {
            set_error("engine: ME read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        memcpy(me_mv, s->pin, b_mv);
        memcpy(me_cand, s->pin + al(b_mv), b_cand);
        memcpy(total_cand, s->pin + al(b_mv) + al(b_cand), b_tot);
        memcpy(rc_me_distortion, s->pin + al(b_mv) + al(b_cand) + al(b_tot), b_rc);
        e->stats.me_pictures++;
        e->stats.d2h_bytes += out_bytes;
    } while (0);
    if (rc != SVT_B200_OK) cudaStreamSynchronize(s->st); // nothing of this picture may still read the planes
    plane_release(e, ents, n_ent);
    release(e, s);
    return rc;
}

int svt_b200_engine_dlf_frame(SvtB200Engine *e, const SvtB200DlfParams *p, const SvtB200Frame *frame, const SvtB200DlfMi *mi) {
    if (!e || !p || !frame || !mi) {
        set_error("svt_b200_engine_dlf_frame: null argument");
        return SVT_B200_ERR_ARG;
    }
    DeviceGuard dg(e->device);
    FiltSlot *s = acquire(e, e->filt);
    int rc;
    do {
        const size_t b_mi = (size_t)p->mi_rows * p->mi_stride * sizeof(SvtB200DlfMi);
        if ((rc = dev_frame(&s->recon, frame)) != SVT_B200_OK) break;
        if ((rc = grow_dev(&s->misc, &s->misc_cap, b_mi)) != SVT_B200_OK) break;
        pin_range(e, mi, b_mi);
        if (cudaMemcpyAsync(s->misc, mi, b_mi, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
            set_error("engine: mode-info upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        if ((rc = copy_frame(e, &s->recon.f, frame, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
        if ((rc = svt_b200_dlf_frame(p, &s->recon.f, (const SvtB200DlfMi *)s->misc, s->st)) != SVT_B200_OK) break;
        if ((rc = copy_frame(e, frame, &s->recon.f, cudaMemcpyDeviceToHost, s->st)) != SVT_B200_OK) break;
        if (cudaStreamSynchronize(s->st) != cudaSuccess) {
            set_error("engine: deblocking failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        e->stats.dlf_frames++;
        e->stats.h2d_bytes += b_mi;
    } while (0);
    if (rc != SVT_B200_OK) cudaStreamSynchronize(s->st);
    release(e, s);
    return rc;
}

int svt_b200_engine_cdef_frame(SvtB200Engine *e, const SvtB200CdefSearchParams *sp, const SvtB200Frame *recon,
                               const SvtB200Frame *source, const uint8_t *skip8, int32_t skip_stride, uint64_t *mse,
                               SvtB200CdefDecideFn decide, void *user) {
    if (!e || !sp || !recon || !source || !skip8 || !mse || !decide) {
        set_error("svt_b200_engine_cdef_frame: null argument");
        return SVT_B200_ERR_ARG;
    }
    DeviceGuard dg(e->device);
    FiltSlot *s = acquire(e, e->filt);
    int rc;
    do {
        const int nvfb = (sp->mi_rows + 15) / 16, nhfb = (sp->mi_cols + 15) / 16, nfb = nvfb * nhfb;
        const size_t b_skip = (size_t)((sp->mi_rows + 1) / 2) * skip_stride, b_mse = (size_t)2 * nfb * 64 * 8;
        auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const size_t o_mse = al(b_skip), o_idx = o_mse + al(b_mse);
        if ((rc = dev_frame(&s->recon, recon)) != SVT_B200_OK) break;
        if ((rc = dev_frame(&s->out, recon)) != SVT_B200_OK) break;
        if ((rc = dev_frame(&s->source, source)) != SVT_B200_OK) break;
        if ((rc = grow_dev(&s->misc, &s->misc_cap, o_idx + al((size_t)nfb))) != SVT_B200_OK) break;
        if ((rc = grow_pin(&s->pin, &s->pin_cap, al(b_mse) + al((size_t)nfb))) != SVT_B200_OK) break;
        if (cudaMemcpyAsync(s->misc, skip8, b_skip, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
            set_error("engine: skip-map upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        if ((rc = copy_frame(e, &s->recon.f, recon, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
        if ((rc = copy_frame(e, &s->source.f, source, cudaMemcpyHostToDevice, s->st)) != SVT_B200_OK) break;
        if ((rc = svt_b200_cdef_search(sp, &s->recon.f, &s->source.f, s->misc, skip_stride, (uint64_t *)(s->misc + o_mse), s->st)) !=
            SVT_B200_OK)
            break;
        if (cudaMemcpyAsync(s->pin, s->misc + o_mse, b_mse, cudaMemcpyDeviceToHost, s->st) != cudaSuccess ||
            cudaStreamSynchronize(s->st) != cudaSuccess) {
            set_error("engine: CDEF search failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        memcpy(mse, s->pin, b_mse);
        // the strength decision (finish_cdef_search, EbEncCdef.c:1167) is the host's: it reads mse and fills the apply set
        SvtB200CdefApplyParams ap;
        memset(&ap, 0, sizeof(ap));
        ap.mi_rows = sp->mi_rows;
        ap.mi_cols = sp->mi_cols;
        int8_t *idx = (int8_t *)(s->pin + al(b_mse));
        memset(idx, -1, (size_t)nfb);
        const int apply = decide(user, mse, &ap, idx);
        e->stats.cdef_frames++;
        e->stats.h2d_bytes += b_skip;
        e->stats.d2h_bytes += b_mse;
        if (apply <= 0) {
            rc = apply < 0 ? SVT_B200_ERR_ARG : SVT_B200_OK;
            break;
        }
        if (cudaMemcpyAsync(s->misc + o_idx, idx, (size_t)nfb, cudaMemcpyHostToDevice, s->st) != cudaSuccess) {
            rc = SVT_B200_ERR_CUDA;
            break;
        }
        if ((rc = svt_b200_cdef_apply(&ap, &s->recon.f, &s->out.f, s->misc, skip_stride, (const int8_t *)(s->misc + o_idx), s->st)) !=
            SVT_B200_OK)
            break;
        if ((rc = copy_frame(e, recon, &s->out.f, cudaMemcpyDeviceToHost, s->st)) != SVT_B200_OK) break;
        if (cudaStreamSynchronize(s->st) != cudaSuccess) {
            set_error("engine: CDEF apply failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = SVT_B200_ERR_CUDA;
            break;
        }
    } while (0);
    if (rc != SVT_B200_OK) cudaStreamSynchronize(s->st);
    release(e, s);
    return rc;
}

} // extern "C"
