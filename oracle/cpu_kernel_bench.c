/* cpu_kernel_bench.c - single-thread time of the reference's own kernels, C flavour next to the SIMD flavour the
 * encoder's RTCD table picks on this machine (per-kernel `_c` / `_avx2` columns).  TEST / MEASUREMENT
 * INFRASTRUCTURE: loads oracle/_ref/simd/libSvtAv1EncSimd.so (the unmodified reference, SSE2..AVX-512 intrinsics) with
 * dlopen and calls the exported flavours directly; nothing of the product is involved.
 *   usage: cpu_kernel_bench <path to libSvtAv1EncSimd.so>      -> one JSON object on stdout */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static void *g_lib;
static double now(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec + 1e-9 * t.tv_nsec;
}
static uint32_t rs = 12345;
static uint32_t rnd(void) { return rs = rs * 1664525u + 1013904223u; }

static uint8_t *pic8;   /* 512 x 512 smooth-ish 8-bit picture */
static uint8_t *pic8b;
static uint16_t *pic16; /* same as 16-bit container */
enum { PW = 512, PH = 512 };

typedef void (*Fn)(void);
typedef void (*Runner)(Fn f, int iters);

/* ---- one runner per signature ---- */
static void run_sad_loop(Fn f, int it) {
    typedef void (*T)(uint8_t *, uint32_t, uint8_t *, uint32_t, uint32_t, uint32_t, uint64_t *, int16_t *, int16_t *, uint32_t, int16_t, int16_t);
    uint64_t best;
    int16_t x, y;
    for (int i = 0; i < it; i++) ((T)f)(pic8 + 100 * PW + 100, PW, pic8b + 60 * PW + 60, PW, 64, 64, &best, &x, &y, PW, 64, 32);
}
static void run_ext_all(Fn f, int it) {
    typedef void (*T)(uint8_t *, uint32_t, uint8_t *, uint32_t, uint32_t, uint32_t *, uint32_t *, uint32_t *, uint32_t *, uint32_t (*)[8], uint32_t (*)[8], uint8_t);
    static uint32_t b8[64], b16[16], m8[64], m16[16], e16[16][8], e8[64][8];
    memset(b8, 0xff, sizeof(b8));
    memset(b16, 0xff, sizeof(b16));
    for (int i = 0; i < it; i++) ((T)f)(pic8 + 100 * PW + 100, PW, pic8b + 100 * PW + 96 + (i & 7) * 8, PW, 0, b8, b16, m8, m16, e16, e8, 0);
}
static void run_cdef_dir(Fn f, int it) {
    typedef int32_t (*T)(const uint16_t *, int32_t, int32_t *, int32_t);
    int32_t var;
    volatile int32_t s = 0;
    for (int i = 0; i < it; i++) s += ((T)f)(pic16 + (8 * (i & 31)) * PW + 8 * ((i >> 5) & 31), PW, &var, 0);
}
static void run_cdef_filter(Fn f, int it) {
    typedef void (*T)(uint8_t *, uint16_t *, int32_t, const uint16_t *, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t);
    static uint8_t dst[64 * 64];
    for (int i = 0; i < it; i++)
        ((T)f)(dst, NULL, 64, pic16 + (16 + 8 * (i & 15)) * 144 + 16 + 8 * ((i >> 4) & 7), 4 + (i & 3), 2, i & 7, 5, 5, 3 /* BLOCK_8X8 */, 0);
}
static void run_lpf(Fn f, int it) {
    typedef void (*T)(uint8_t *, int32_t, const uint8_t *, const uint8_t *, const uint8_t *);
    static const uint8_t bl[16] = {40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40, 40}, li[16] = {12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12},
                         th[16] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2};
    for (int i = 0; i < it; i++) ((T)f)(pic8 + (8 + 8 * (i & 31)) * PW + 4 * ((i >> 5) & 63) + 8, PW, bl, li, th);
}
static int16_t resid[64 * 64];
static int32_t coef[64 * 64], qc[64 * 64], dq[64 * 64];
static void run_fwd(Fn f, int it) {
    typedef void (*T)(int16_t *, int32_t *, uint32_t, uint8_t, uint8_t);
    for (int i = 0; i < it; i++) ((T)f)(resid, coef, 64, 0, 8);
}
static void run_inv16(Fn f, int it) {
    typedef void (*T)(const int32_t *, uint16_t *, int32_t, uint16_t *, int32_t, uint8_t, int32_t);
    static uint16_t out[16 * 16];
    for (int i = 0; i < it; i++) ((T)f)(dq, pic16 + 64 * PW + 64, PW, out, 16, 0, 8);
}
static int16_t scan16[256], iscan16[256];
static void run_quant_b(Fn f, int it) {
    typedef void (*T)(const int32_t *, intptr_t, const int16_t *, const int16_t *, const int16_t *, const int16_t *, int32_t *, int32_t *, const int16_t *, uint16_t *,
                      const int16_t *, const int16_t *, const void *, const void *, int32_t);
    static const int16_t zbin[8] = {60, 72, 72, 72, 72, 72, 72, 72}, round_[8] = {30, 36, 36, 36, 36, 36, 36, 36}, quant[8] = {9000, 7000, 7000, 7000, 7000, 7000, 7000, 7000},
                         shift[8] = {16384, 16384, 16384, 16384, 16384, 16384, 16384, 16384}, deq[8] = {80, 96, 96, 96, 96, 96, 96, 96};
    uint16_t eob;
    for (int i = 0; i < it; i++) ((T)f)(coef, 256, zbin, round_, quant, shift, qc, dq, deq, &eob, scan16, iscan16, NULL, NULL, 0);
}
static void run_stats(Fn f, int it) {
    typedef void (*T)(int32_t, const uint8_t *, const uint8_t *, int32_t, int32_t, int32_t, int32_t, int32_t, int32_t, int64_t *, int64_t *);
    static int64_t M[49], H[49 * 49];
    for (int i = 0; i < it; i++) ((T)f)(7, pic8, pic8b, 64, 128, 64, 128, PW, PW, M, H);
}
static void run_resid(Fn f, int it) {
    typedef void (*T)(uint8_t *, uint32_t, uint8_t *, uint32_t, int16_t *, uint32_t, uint32_t, uint32_t);
    for (int i = 0; i < it; i++) ((T)f)(pic8 + 64 * PW + 64, PW, pic8b + 64 * PW + 64, PW, resid, 64, 64, 64);
}
static void run_sgr(Fn f, int it) {
    typedef void (*T)(const uint8_t *, int32_t, int32_t, int32_t, int32_t *, int32_t *, int32_t, int32_t, int32_t, int32_t);
    static int32_t f0[64 * 72], f1[64 * 72];
    for (int i = 0; i < it; i++) ((T)f)(pic8 + 64 * PW + 64, 64, 64, PW, f0, f1, 72, 4, 8, 0);
}

typedef struct {
    const char *what, *unit_note;
    Runner run;
    int iters;
    const char *names[5]; /* [0] = the C flavour, then SIMD flavours in ascending order (last found = what the RTCD picks) */
} Row;

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    g_lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
    if (!g_lib) {
        fprintf(stderr, "%s\n", dlerror());
        return 1;
    }
    pic8 = malloc(PW * PH), pic8b = malloc(PW * PH), pic16 = malloc(PW * PH * 2);
    for (int y = 0; y < PH; y++)
        for (int x = 0; x < PW; x++) {
            int v = 128 + (int)(60 * ((x * 7 + y * 3) % 97) / 97.0) + (int)(rnd() % 9) - 4;
            pic8[y * PW + x] = (uint8_t)v;
            pic8b[y * PW + x] = (uint8_t)(v + (int)(rnd() % 13) - 6);
            pic16[y * PW + x] = (uint16_t)v;
        }
    for (int i = 0; i < 64 * 64; i++) resid[i] = (int16_t)((int)(rnd() % 61) - 30), coef[i] = (int)(rnd() % 2001) - 1000, dq[i] = i < 40 ? (int)(rnd() % 801) - 400 : 0;
    for (int i = 0; i < 256; i++) scan16[i] = (int16_t)i, iscan16[i] = (int16_t)i;
    void (*init_c)(uint64_t) = (void (*)(uint64_t))dlsym(g_lib, "setup_common_rtcd_internal");
    (void)init_c;
    Row rows[] = {
        {"svt_sad_loop_kernel 64x64 block, 64x32 search area (2048 positions)", "", run_sad_loop, 40,
         {"svt_sad_loop_kernel_c", "svt_sad_loop_kernel_sse4_1_intrin", "svt_sad_loop_kernel_avx2_intrin", "svt_sad_loop_kernel_avx512_intrin"}},
        {"svt_ext_all_sad_calculation_8x8_16x16 (one 64x64 SB, 8 positions)", "", run_ext_all, 20000,
         {"svt_ext_all_sad_calculation_8x8_16x16_c", "svt_ext_all_sad_calculation_8x8_16x16_avx2"}},
        {"svt_cdef_find_dir (8x8)", "", run_cdef_dir, 400000, {"svt_cdef_find_dir_c", "svt_cdef_find_dir_avx2"}},
        {"svt_cdef_filter_block 8x8 (8-bit destination)", "", run_cdef_filter, 400000, {"svt_cdef_filter_block_c", "svt_cdef_filter_block_avx2"}},
        {"svt_aom_lpf_horizontal_8 (4 samples of one edge)", "", run_lpf, 2000000, {"svt_aom_lpf_horizontal_8_c", "svt_aom_lpf_horizontal_8_sse2"}},
        {"svt_av1_fwd_txfm2d_16x16 DCT_DCT", "", run_fwd, 200000,
         {"svt_av1_transform_two_d_16x16_c", "svt_av1_fwd_txfm2d_16x16_avx2", "svt_av1_fwd_txfm2d_16x16_avx512"}},
        {"svt_av1_fwd_txfm2d_32x32 DCT_DCT", "", run_fwd, 50000,
         {"svt_av1_transform_two_d_32x32_c", "svt_av1_fwd_txfm2d_32x32_avx2", "svt_av1_fwd_txfm2d_32x32_avx512"}},
        {"svt_av1_inv_txfm2d_add_16x16 DCT_DCT", "", run_inv16, 200000,
         {"svt_av1_inv_txfm2d_add_16x16_c", "svt_av1_inv_txfm2d_add_16x16_sse4_1", "svt_av1_inv_txfm2d_add_16x16_avx2", "svt_av1_inv_txfm2d_add_16x16_avx512"}},
        {"svt_aom_quantize_b 16x16 (256 coefficients)", "", run_quant_b, 400000, {"svt_aom_quantize_b_c_ii", "svt_aom_quantize_b_avx2"}},
        {"svt_av1_compute_stats win 7, one 64x64 unit", "", run_stats, 300,
         {"svt_av1_compute_stats_c", "svt_av1_compute_stats_avx2", "svt_av1_compute_stats_avx512"}},
        {"svt_residual_kernel8bit 64x64", "", run_resid, 400000, {"svt_residual_kernel8bit_c", "svt_residual_kernel8bit_avx2", "svt_residual_kernel8bit_avx512"}},
        {"svt_av1_selfguided_restoration 64x64, params 4 (both radii)", "", run_sgr, 2000,
         {"svt_av1_selfguided_restoration_c", "svt_av1_selfguided_restoration_avx2"}},
    };
    const int avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    printf("{\"threads\": 1, \"avx512\": %s, \"rows\": [\n", avx512 ? "true" : "false");
    const int n = (int)(sizeof(rows) / sizeof(rows[0]));
    for (int r = 0; r < n; r++) {
        printf(" {\"kernel\": \"%s\"", rows[r].what);
        double t_c = 0;
        for (int k = 0; k < 5 && rows[r].names[k]; k++) {
            if (strstr(rows[r].names[k], "avx512") && !avx512) continue;
            Fn f = (Fn)dlsym(g_lib, rows[r].names[k]);
            if (!f) continue;
            int it = rows[r].iters;
            if (k == 0 && it > 2000) it /= 4; /* the C flavours are slow */
            rows[r].run(f, it / 10 + 1);
            double best = 1e30;
            for (int rep = 0; rep < 3; rep++) {
                const double t0 = now();
                rows[r].run(f, it);
                const double dt = (now() - t0) / it;
                if (dt < best) best = dt;
            }
            if (k == 0) t_c = best;
            printf(", \"%s_us\": %.4f", rows[r].names[k], best * 1e6);
            if (k) printf(", \"%s_speedup_vs_c\": %.1f", rows[r].names[k], t_c / best);
        }
        printf("}%s\n", r + 1 < n ? "," : "");
    }
    printf("]}\n");
    return 0;
}
