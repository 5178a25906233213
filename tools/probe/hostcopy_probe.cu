// hostcopy_probe: host memcpy bandwidth between pageable memory and cudaMallocHost memory on the GPU box (diagnostic for the
// picture engine's staging copies).  nvcc -O2 -o hostcopy_probe hostcopy_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void par(void *d, const void *s, size_t n, int nt) {
    std::vector<std::thread> th;
    size_t q = n / nt;
    for (int t = 1; t < nt; t++) th.emplace_back([=] { memcpy((char *)d + q * t, (const char *)s + q * t, t == nt - 1 ? n - q * t : q); });
    memcpy(d, s, nt == 1 ? n : q);
    for (auto &t : th) t.join();
}
int main() {
    const size_t n = 25u << 20;
    char *pg = (char *)malloc(n), *pg2 = (char *)malloc(n), *pin = nullptr, *pinwc = nullptr;
    memset(pg, 1, n);
    memset(pg2, 2, n);
    cudaMallocHost(&pin, n);
    cudaHostAlloc(&pinwc, n, cudaHostAllocWriteCombined);
    memset(pin, 3, n);
    memset(pinwc, 3, n);
    struct { const char *name; char *d; char *s; } cases[] = {{"pageable -> pageable", pg2, pg}, {"pageable -> pinned", pin, pg}, {"pinned -> pageable", pg, pin},
                                                          {"pageable -> pinned(WC)", pinwc, pg}};
    for (auto &c : cases)
        for (int nt : {1, 4}) {
            par(c.d, c.s, n, nt);
            double t0 = now();
            for (int i = 0; i < 5; i++) par(c.d, c.s, n, nt);
            double dt = (now() - t0) / 5;
            printf("%-24s %d thread(s): %6.2f ms  %6.2f GB/s\n", c.name, nt, dt * 1e3, n / dt / 1e9);
        }
    return 0;
}
