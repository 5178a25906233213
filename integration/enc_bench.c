/*
 * enc_bench.c - measurement driver around the reference's public API (Source/API/EbSvtAv1Enc.h): our own code, linked
 * against either the reference's SIMD build (oracle/_ref/simd) or the overlay build with the CUDA backend
 * (integration/_build).  It does what SvtAv1EncApp does - svt_av1_enc_init_handle / set_parameter / init /
 * send_picture / get_packet - but from a clip preloaded in host RAM, and it timestamps every output packet so that
 * bench.py can time EXACTLY the frames of the K timed steps after W warm-up steps (the app only prints one average
 * that includes pipeline fill).  Prints one JSON object.
 *
 *   enc_bench <clip.yuv> <width> <height> <frames> <bits> <preset> <qp> <warm_frames> [ivf_out]
 *
 * "fps_timed" = (frames - warm_frames) / (t(last packet) - t(packet #warm_frames)); "fps_all" = frames / (t(last packet)
 * - t(first send)), the app's "Average Speed" definition (EbAppMain.c:277-303).  md5-able output: the concatenated
 * packet payloads (the IVF framing of the app adds only headers).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "EbSvtAv1Enc.h"

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + t.tv_nsec * 1e-9;
}

typedef struct {
    EbComponentType *handle;
    uint8_t *        clip;
    int              w, h, frames, bps;
    double           t_first_send;
} Sender;

static void *sender_main(void *arg) {
    Sender *           s = (Sender *)arg;
    const size_t       luma = (size_t)s->w * s->h * s->bps, chroma = (size_t)(s->w / 2) * (s->h / 2) * s->bps;
    EbBufferHeaderType hdr;
    EbSvtIOFormat      io;
    memset(&hdr, 0, sizeof(hdr));
    memset(&io, 0, sizeof(io));
    hdr.size     = sizeof(hdr);
    hdr.p_buffer = (uint8_t *)&io;
    for (int n = 0; n < s->frames; n++) {
        uint8_t *f   = s->clip + (size_t)n * (luma + 2 * chroma);
        io.luma      = f;
        io.cb        = f + luma;
        io.cr        = f + luma + chroma;
        io.y_stride  = s->w;
        io.cb_stride = io.cr_stride = s->w / 2;
        io.width                    = s->w;
        io.height                   = s->h;
        hdr.n_filled_len            = (uint32_t)(luma + 2 * chroma);
        hdr.pts                     = n;
        hdr.pic_type                = EB_AV1_INVALID_PICTURE;
        hdr.flags                   = 0;
        hdr.metadata                = NULL;
        if (n == 0) s->t_first_send = now_s();
        svt_av1_enc_send_picture(s->handle, &hdr);
    }
    EbBufferHeaderType eos;
    memset(&eos, 0, sizeof(eos));
    eos.size     = sizeof(eos);
    eos.flags    = EB_BUFFERFLAG_EOS;
    eos.pic_type = EB_AV1_INVALID_PICTURE;
    svt_av1_enc_send_picture(s->handle, &eos);
    return NULL;
}

int main(int argc, char **argv) {
    if (argc < 9) {
        fprintf(stderr, "usage: enc_bench clip.yuv width height frames bits preset qp warm_frames [out.obu]\n");
        return 2;
    }
    const char *path = argv[1];
    const int   w = atoi(argv[2]), h = atoi(argv[3]), frames = atoi(argv[4]), bits = atoi(argv[5]), preset = atoi(argv[6]),
              qp = atoi(argv[7]), warm = atoi(argv[8]);
    const int    bps   = bits > 8 ? 2 : 1;
    const size_t fsize = ((size_t)w * h + 2 * (size_t)(w / 2) * (h / 2)) * bps;
    uint8_t *    clip  = (uint8_t *)malloc(fsize * frames);
    FILE *       f     = fopen(path, "rb");
    if (!clip || !f || fread(clip, fsize, frames, f) != (size_t)frames) {
        fprintf(stderr, "enc_bench: cannot read %d frames from %s\n", frames, path);
        return 2;
    }
    fclose(f);
    FILE *out = argc > 9 ? fopen(argv[9], "wb") : NULL;

    EbComponentType *        handle = NULL;
    EbSvtAv1EncConfiguration cfg;
    memset(&cfg, 0, sizeof(cfg)); /* the app callocs its EbConfig: init_handle does not set every field */
    if (svt_av1_enc_init_handle(&handle, NULL, &cfg) != EB_ErrorNone) return 3;
    cfg.source_width       = w;
    cfg.source_height      = h;
    cfg.frame_rate         = 30 << 16;
    cfg.enc_mode           = (int8_t)preset;
    cfg.rate_control_mode  = 0;
    cfg.qp                 = qp;
    cfg.encoder_bit_depth  = bits;
    cfg.recon_enabled      = 0;
    if (getenv("ENC_BENCH_LP")) cfg.logical_processors = atoi(getenv("ENC_BENCH_LP"));
    if (getenv("ENC_BENCH_SOCKET")) cfg.target_socket = atoi(getenv("ENC_BENCH_SOCKET"));
    if (svt_av1_enc_set_parameter(handle, &cfg) != EB_ErrorNone) return 3;
    const double t_init0 = now_s();
    if (svt_av1_enc_init(handle) != EB_ErrorNone) return 3;
    const double t_init1 = now_s();

    Sender    s = {handle, clip, w, h, frames, bps, 0.0};
    pthread_t th;
    pthread_create(&th, NULL, sender_main, &s);

    double   t_warm = 0, t_last = 0;
    int      n_out = 0, eos = 0;
    uint64_t bytes = 0;
    while (!eos) {
        EbBufferHeaderType *pkt = NULL;
        EbErrorType         st  = svt_av1_enc_get_packet(handle, &pkt, 1 /* blocking */);
        if (st == EB_ErrorMax) {
            fprintf(stderr, "enc_bench: encoder error\n");
            return 4;
        }
        if (st == EB_NoErrorEmptyQueue || !pkt) continue;
        const uint32_t flags = pkt->flags;
        if (!(flags & EB_BUFFERFLAG_IS_ALT_REF)) {
            n_out++;
            t_last = now_s();
            if (n_out == warm) t_warm = t_last;
        }
        bytes += pkt->n_filled_len;
        if (out) fwrite(pkt->p_buffer, 1, pkt->n_filled_len, out);
        eos = (flags & EB_BUFFERFLAG_EOS) != 0;
        svt_av1_enc_release_out_buffer(&pkt);
    }
    pthread_join(th, NULL);
    if (out) fclose(out);
    if (warm <= 0) t_warm = s.t_first_send;
    printf("{\"frames\": %d, \"warm_frames\": %d, \"packets\": %d, \"bytes\": %llu, \"init_s\": %.3f, \"seconds_all\": %.4f, "
           "\"seconds_timed\": %.4f, \"fps_all\": %.3f, \"fps_timed\": %.3f}\n",
           frames, warm, n_out, (unsigned long long)bytes, t_init1 - t_init0, t_last - s.t_first_send, t_last - t_warm,
           frames / (t_last - s.t_first_send), (frames - (warm > 0 ? warm : 0)) / (t_last - t_warm));
    fflush(stdout);
    svt_av1_enc_deinit(handle);
    svt_av1_enc_deinit_handle(handle);
    free(clip);
    return n_out == frames ? 0 : 5;
}
