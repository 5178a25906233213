/* prof_preload.c - a tiny sampling profiler (the image has no perf / gdb / gprof-for-threads).
 *
 *   gcc -O2 -fPIC -shared -o /tmp/libprof.so tools/prof_preload.c -ldl -lpthread -lrt
 *   PROF_OUT=/tmp/prof.txt LD_PRELOAD=/tmp/libprof.so  SvtAv1EncApp...
 *
 * Every thread gets a CPU-time interval timer (CLOCK_THREAD_CPUTIME_ID, 1 ms) that delivers SIGPROF to that thread;
 * the handler records the interrupted program counter.  At exit the PCs are resolved with dladdr and a flat profile
 * (CPU-time share per function, all threads) is written to $PROF_OUT.  Measurement infrastructure only.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <time.h>
#include <ucontext.h>
#include <unistd.h>

#define MAX_SAMPLES (1u << 23)
static uintptr_t *        g_pc;
static volatile uint32_t g_n;

static void on_prof(int sig, siginfo_t *si, void *uc_) {
    (void)sig; (void)si;
    ucontext_t *uc = (ucontext_t *)uc_;
    uint32_t    i  = __sync_fetch_and_add(&g_n, 1);
    if (i < MAX_SAMPLES) g_pc[i] = (uintptr_t)uc->uc_mcontext.gregs[REG_RIP];
}

static void arm_thread_timer(void) {
    struct sigevent sev;
    memset(&sev, 0, sizeof(sev));
    sev.sigev_notify          = SIGEV_THREAD_ID;
    sev.sigev_signo           = SIGPROF;
    sev._sigev_un._tid        = (pid_t)syscall(SYS_gettid);
    timer_t t;
    if (timer_create(CLOCK_THREAD_CPUTIME_ID, &sev, &t) != 0) return;
    struct itimerspec its = {{0, 1000000}, {0, 1000000}};
    timer_settime(t, 0, &its, NULL);
}

struct start { void *(*fn)(void *); void *arg; };
static void *trampoline(void *p) {
    struct start s = *(struct start *)p;
    free(p);
    arm_thread_timer();
    return s.fn(s.arg);
}
int pthread_create(pthread_t *th, const pthread_attr_t *attr, void *(*fn)(void *), void *arg) {
    static int (*real)(pthread_t *, const pthread_attr_t *, void *(*)(void *), void *);
    if (!real) real = dlsym(RTLD_NEXT, "pthread_create");
    struct start *s = malloc(sizeof(*s));
    s->fn = fn; s->arg = arg;
    return real(th, attr, trampoline, s);
}

static void dump(void) {
    const char *path = getenv("PROF_OUT");
    uint32_t    n    = g_n < MAX_SAMPLES ? g_n : MAX_SAMPLES;
    if (n < 100) return; /* wrapper processes (timeout, sh) load the preload too */
    char full[4096];
    if (path) snprintf(full, sizeof(full), "%s.%d", path, (int)getpid());
    FILE *f = path ? fopen(full, "w") : stderr;
    if (!f) return;
    /* raw samples as "object offset": tools/prof_resolve.py maps them to functions through the full symbol table
     * (dladdr only sees exported symbols, and half of an encoder's time is in static functions) */
    fprintf(f, "# %u samples of 1 ms thread CPU time\n", n);
    for (uint32_t i = 0; i < n; i++) {
        Dl_info di;
        if (dladdr((void *)g_pc[i], &di) && di.dli_fname)
            fprintf(f, "%s %lx\n", di.dli_fname, (unsigned long)(g_pc[i] - (uintptr_t)di.dli_fbase));
        else
            fprintf(f, "? %lx\n", (unsigned long)g_pc[i]);
    }
    if (f != stderr) fclose(f);
}

__attribute__((constructor)) static void init(void) {
    g_pc = malloc(sizeof(uintptr_t) * MAX_SAMPLES);
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_sigaction = on_prof;
    sa.sa_flags     = SA_SIGINFO | SA_RESTART;
    sigaction(SIGPROF, &sa, NULL);
    arm_thread_timer();
    atexit(dump);
}
