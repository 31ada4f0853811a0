// sr_pack_host.cpp -- host side of the packed PCM transport of sr_recognise_batch (sr_api.cu).
//
// The end-to-end call is bound by the PCIe copy of the u16 PCM (1.05 GB per 65 536 utterances, ~54 GB/s). The
// reference's samples are 12-bit ADC readings (ADC.C: 12-bit right-aligned conversions), so host worker threads
// repack chunks whose samples are all < 4096 into 12 bits per sample (3 bytes per sample pair: a | b << 12,
// little endian) while other chunks travel unpacked; the device expands them again before the first kernel
// (unpack12_kernel). Any chunk holding a sample >= 4096 is sent as it is, so the call stays exact for every u16
// input. Plain C++ (built with g++, not nvcc) so that the SIMD variants can use target attributes.
#include <immintrin.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "sr_pack_host.h"

namespace srk {

// ---- 12-bit packing of n samples (n even); returns the OR of all samples (valid iff (result & 0xF000) == 0) -------
static uint32_t pack12_scalar(const uint16_t *src, size_t n, uint8_t *dst) {
    uint32_t o = 0;
    for (size_t i = 0; i + 2 <= n; i += 2) {
        const uint32_t a = src[i], b = src[i + 1];
        o |= a | b;
        uint8_t *d = dst + (i >> 1) * 3;
        d[0] = (uint8_t)a;
        d[1] = (uint8_t)((a >> 8) | (b << 4));
        d[2] = (uint8_t)(b >> 4);
    }
    return o;
}

__attribute__((target("avx2"))) static uint32_t pack12_avx2(const uint16_t *src, size_t n, uint8_t *dst) {
    const __m256i m1 = _mm256_set1_epi32(0x00000FFF), m2 = _mm256_set1_epi32(0x00FFF000);
    const __m256i shuf = _mm256_setr_epi8(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1,
                                          0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1);
    const __m256i perm = _mm256_setr_epi32(0, 1, 2, 4, 5, 6, 3, 7);
    __m256i acc = _mm256_setzero_si256();
    size_t i = 0;
    for (; i + 16 <= n; i += 16) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + i));
        acc = _mm256_or_si256(acc, v);
        // per 32-bit lane (a | b << 16): 24 bits a | b << 12
        const __m256i p = _mm256_or_si256(_mm256_and_si256(v, m1), _mm256_and_si256(_mm256_srli_epi32(v, 4), m2));
        const __m256i r = _mm256_permutevar8x32_epi32(_mm256_shuffle_epi8(p, shuf), perm);   // 24 bytes at the bottom
        uint8_t *d = dst + (i >> 1) * 3;
        _mm_storeu_si128(reinterpret_cast<__m128i *>(d), _mm256_castsi256_si128(r));
        _mm_storel_epi64(reinterpret_cast<__m128i *>(d + 16), _mm256_extracti128_si256(r, 1));
    }
    uint32_t lanes[8];
    _mm256_storeu_si256(reinterpret_cast<__m256i *>(lanes), acc);
    uint32_t o = 0;
    for (int k = 0; k < 8; ++k) o |= lanes[k];
    o = (o | (o >> 16)) & 0xFFFFu;
    return o | pack12_scalar(src + i, n - i, dst + (i >> 1) * 3);
}

__attribute__((target("avx512f,avx512bw,avx512vbmi"))) static uint32_t pack12_vbmi(const uint16_t *src, size_t n,
                                                                                   uint8_t *dst) {
    alignas(64) static const uint8_t idxb[64] = {0,  1,  2,  4,  5,  6,  8,  9,  10, 12, 13, 14, 16, 17, 18, 20,
                                                 21, 22, 24, 25, 26, 28, 29, 30, 32, 33, 34, 36, 37, 38, 40, 41,
                                                 42, 44, 45, 46, 48, 49, 50, 52, 53, 54, 56, 57, 58, 60, 61, 62,
                                                 0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0};
    const __m512i idx = _mm512_load_si512(idxb);
    const __m512i m1 = _mm512_set1_epi32(0x00000FFF), m2 = _mm512_set1_epi32(0x00FFF000);
    __m512i acc = _mm512_setzero_si512();
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m512i v = _mm512_loadu_si512(src + i);
        acc = _mm512_or_si512(acc, v);
        const __m512i p = _mm512_or_si512(_mm512_and_si512(v, m1), _mm512_and_si512(_mm512_srli_epi32(v, 4), m2));
        _mm512_mask_storeu_epi8(dst + (i >> 1) * 3, 0xFFFFFFFFFFFFull, _mm512_permutexvar_epi8(idx, p));
    }
    uint32_t o = (uint32_t)_mm512_reduce_or_epi32(acc);
    o = (o | (o >> 16)) & 0xFFFFu;
    return o | pack12_scalar(src + i, n - i, dst + (i >> 1) * 3);
}

typedef uint32_t (*pack_fn)(const uint16_t *, size_t, uint8_t *);
static pack_fn pick_pack() {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("avx512bw")) return pack12_vbmi;
    if (__builtin_cpu_supports("avx2")) return pack12_avx2;
    return pack12_scalar;
}
uint32_t pack12(const uint16_t *src, size_t n, uint8_t *dst) {
    static const pack_fn f = pick_pack();
    return f(src, n, dst);
}
uint32_t pack12_variant(int variant, const uint16_t *src, size_t n, uint8_t *dst) {
    __builtin_cpu_init();
    if (variant == 2 && __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("avx512bw")) return pack12_vbmi(src, n, dst);
    if (variant == 1 && __builtin_cpu_supports("avx2")) return pack12_avx2(src, n, dst);
    if (variant == 0) return pack12_scalar(src, n, dst);
    return 0xFFFFFFFFu;   // variant not available on this CPU
}

// CPUs this process may use: affinity mask capped by the cgroup CPU quota (cpu.max: "quota period" or "max")
int usable_cpus() {
    cpu_set_t set;
    int n = 1;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
    FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r");
    if (f) {
        char q[64];
        long long period = 0;
        if (fscanf(f, "%63s %lld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) {
            const long long quota = atoll(q);
            const int cap = (int)((quota + period - 1) / period);
            if (cap >= 1 && cap < n) n = cap;
        }
        fclose(f);
    }
    return n < 1 ? 1 : n;
}

// ---- fork-join pool: run() packs one range with all workers ---------------------------------------------------------
struct PackPool::Impl {
    std::vector<std::thread> th;
    std::mutex m;
    std::condition_variable cv_go, cv_done;
    uint64_t gen = 0;
    int remaining = 0;
    bool stop = false;
    const uint16_t *src = nullptr;
    uint8_t *dst = nullptr;
    size_t n = 0;
    std::atomic<uint32_t> orbits{0};
    int nthreads = 0;

    void slice(int t) {
        // slices are multiples of 32 samples (48 packed bytes): no two workers touch the same output byte
        const size_t groups = n / 32, per = (groups + nthreads - 1) / nthreads;
        const size_t g0 = (size_t)t * per, g1 = g0 + per < groups ? g0 + per : groups;
        uint32_t o = 0;
        if (g0 < g1) o = pack12(src + g0 * 32, (g1 - g0) * 32, dst + g0 * 48);
        if (t == nthreads - 1 && groups * 32 < n) o |= pack12(src + groups * 32, n - groups * 32, dst + groups * 48);
        orbits.fetch_or(o, std::memory_order_relaxed);
    }
    void worker(int t) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m);
                cv_go.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
            }
            slice(t);
            {
                std::lock_guard<std::mutex> lk(m);
                if (--remaining == 0) cv_done.notify_all();
            }
        }
    }
};

PackPool::PackPool(int nthreads) : p(new Impl) {
    p->nthreads = nthreads < 1 ? 1 : nthreads;
    for (int t = 0; t < p->nthreads; ++t) p->th.emplace_back([this, t] { p->worker(t); });
}
PackPool::~PackPool() {
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->stop = true;
    }
    p->cv_go.notify_all();
    for (auto &t : p->th) t.join();
    delete p;
}
int PackPool::threads() const { return p->nthreads; }
uint32_t PackPool::run(const uint16_t *src, size_t n, uint8_t *dst) {
    std::unique_lock<std::mutex> lk(p->m);
    p->src = src; p->dst = dst; p->n = n;
    p->orbits.store(0, std::memory_order_relaxed);
    p->remaining = p->nthreads;
    ++p->gen;
    p->cv_go.notify_all();
    p->cv_done.wait(lk, [&] { return p->remaining == 0; });
    return p->orbits.load(std::memory_order_relaxed);
}

}  // namespace srk
