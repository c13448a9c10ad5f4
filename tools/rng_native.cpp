#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <chrono>
#include <vector>
#include <dlfcn.h>
typedef int64_t (*fn_t)(void*, int64_t, int32_t, double, int64_t, uint8_t*);
int main() {
    void* h = dlopen("/root/repo/subspace-reg_b200/srb200/libsrb200.so", RTLD_NOW);
    if (!h) { printf("dlopen: %s\n", dlerror()); return 1; }
    fn_t f = (fn_t)dlsym(h, "sr_host_bernoulli");
    struct Blob { uint64_t seed; int32_t left; int32_t seeded; uint64_t next; uint64_t state[624]; double pad[8]; } b;
    memset(&b, 0, sizeof(b));
    b.seed = 1; b.left = 1; b.seeded = 1; b.next = 624;
    for (int i = 0; i < 624; ++i) b.state[i] = (uint32_t)(i * 2654435761u + 17);
    const int64_t n = 22579200;
    std::vector<uint8_t> out(n, 0);
    for (const char* thr : {"0", "2", "4", "6", "8", "12"}) {
        setenv("SRB_RNG_THREADS", thr, 1);
        for (int kind = 0; kind < 3; ++kind) {
            double best = 1e9;
            for (int rep = 0; rep < 4; ++rep) {
                auto t0 = std::chrono::steady_clock::now();
                f(&b, sizeof(b), kind, 0.9, n, out.data());
                auto t1 = std::chrono::steady_clock::now();
                double d = std::chrono::duration<double>(t1 - t0).count();
                if (d < best) best = d;
            }
            printf("threads %s kind %d: %.2f ms  %.3f ns/elem\n", thr, kind, best * 1e3, best * 1e9 / n);
        }
    }
}
