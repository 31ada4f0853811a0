// sr_synth.cu -- synthetic PCM / feature workload generator (see include/sr_synth.h). Integer-only
// `__host__ __device__` core so that CPU and GPU produce byte-identical buffers.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <mutex>
#include <thread>
#include <vector>
#include "../../include/sr_synth.h"
#include "sr_synth_tables.h"

namespace {

struct Word { uint32_t ws, wl, nh, step, hw[5], hwsum, pk; };
struct Params { uint64_t seed; uint32_t mid, na, nw; Word w[3]; };

__host__ __device__ inline uint64_t splitmix64(uint64_t &state) {
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint32_t hash32(uint64_t key, uint64_t n) {
    uint64_t s = key ^ (n * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    return (uint32_t)(splitmix64(s) >> 32);
}

__host__ __device__ inline void make_params(Params &p, uint64_t seed, uint32_t U, uint32_t nwords) {
    uint64_t st = seed;
    p.seed = seed;
    p.mid = 1900u + (uint32_t)(splitmix64(st) % 301u);
    p.na = 15u + (uint32_t)(splitmix64(st) % 46u);
    uint32_t pos = 2480u + (uint32_t)(splitmix64(st) % 720u);
    p.nw = 0;
    if (nwords > 3) nwords = 3;
    for (uint32_t k = 0; k < nwords; ++k) {
        Word w;
        w.wl = 2000u + (uint32_t)(splitmix64(st) % 1601u);
        w.nh = 3u + (uint32_t)(splitmix64(st) % 3u);
        const uint32_t f0 = 100u + (uint32_t)(splitmix64(st) % 151u);
        w.step = (uint32_t)((((uint64_t)f0) << 32) / 8000u);
        w.hwsum = 0;
        for (int h = 0; h < 5; ++h) {
            w.hw[h] = 77u + (uint32_t)(splitmix64(st) % 180u);
            if ((uint32_t)h < w.nh) w.hwsum += w.hw[h];
        }
        w.pk = 300u + (uint32_t)(splitmix64(st) % 1201u);
        const uint32_t gap = 1600u + (uint32_t)(splitmix64(st) % 2401u);
        w.ws = pos;
        if ((uint64_t)pos + w.wl + 1040u > U) break;
        p.w[p.nw++] = w;
        pos += w.wl + gap;
    }
}

__host__ __device__ inline int32_t sine_q15(const int16_t *tab, uint32_t phase) { return tab[phase >> 22]; }

__host__ __device__ inline uint16_t sample(const Params &p, const int16_t *tab, uint32_t n) {
    const uint32_t blk = n / 80u, r = n % 80u;
    const uint32_t rot = hash32(p.seed ^ 0xB10Cull, blk) % 80u;
    const uint32_t pv = (r * 37u + rot) % 80u;
    const int32_t mag = (int32_t)((p.na * pv) / 80u);
    int32_t v = (int32_t)p.mid + ((hash32(p.seed, n) & 1u) ? mag : -mag);
    for (uint32_t k = 0; k < p.nw; ++k) {
        const Word &w = p.w[k];
        if (n < w.ws || n >= w.ws + w.wl) continue;
        const uint32_t t = n - w.ws;
        const uint32_t rr = t < w.wl - 1u - t ? t : w.wl - 1u - t;
        int32_t env = 32767;
        if (rr < 160u) env = (32768 - (int32_t)tab[(((rr * 512u) / 160u) + 256u) & 1023u]) >> 1;
        int32_t s = 0;
        for (uint32_t h = 0; h < w.nh; ++h) s += (int32_t)w.hw[h] * sine_q15(tab, (h + 1u) * w.step * t);
        const int32_t voiced = s / (int32_t)w.hwsum;
        const int32_t fr = (int32_t)(hash32(p.seed ^ 0xF1C0FFEEull, n) & 0xFFFFu) - 32768;
        const int32_t mix = (voiced * 9 + fr) / 10;
        const int32_t amp = (int32_t)(((int64_t)w.pk * env) >> 15);
        v += (int32_t)(((int64_t)amp * mix) >> 15);
    }
    if (v < 0) v = 0;
    if (v > 4095) v = 4095;
    return (uint16_t)v;
}

__global__ void synth_kernel(uint16_t *pcm, uint32_t U, uint32_t B, uint64_t seed_base, uint32_t nwords,
                             const int16_t *tab_g) {
    __shared__ int16_t tab[1024];
    __shared__ Params p;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = tab_g[i];
    for (uint32_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) make_params(p, seed_base + b, U, nwords);
        __syncthreads();
        uint16_t *out = pcm + (size_t)b * U;
        for (uint32_t n = threadIdx.x; n < U; n += blockDim.x) out[n] = sample(p, tab, n);
    }
}

}  // namespace

extern "C" int sr_synth_pcm_host(uint16_t *pcm, uint32_t U, uint32_t B, uint64_t seed_base, uint32_t nwords) {
    if (!pcm && B) return -1;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 64) nt = 64;
    if (nt > B) nt = B ? B : 1;
    auto work = [=](uint32_t lo, uint32_t hi) {
        for (uint32_t b = lo; b < hi; ++b) {
            Params p;
            make_params(p, seed_base + b, U, nwords);
            uint16_t *out = pcm + (size_t)b * U;
            for (uint32_t n = 0; n < U; ++n) out[n] = sample(p, sr_synth_sine, n);
        }
    };
    std::vector<std::thread> th;
    for (unsigned k = 0; k < nt; ++k)
        th.emplace_back(work, (uint32_t)((uint64_t)B * k / nt), (uint32_t)((uint64_t)B * (k + 1) / nt));
    for (auto &t : th) t.join();
    return 0;
}

extern "C" int sr_synth_pcm_dev(uint16_t *pcm_dev, uint32_t U, uint32_t B, uint64_t seed_base, uint32_t nwords,
                                void *cuda_stream) {
    if (B == 0) return 0;
    static int16_t *tab_dev[64] = {nullptr};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return -1;
    std::lock_guard<std::mutex> lk(mu);
    if (!tab_dev[dev]) {
        if (cudaMalloc(&tab_dev[dev], sizeof(sr_synth_sine)) != cudaSuccess) return -1;
        if (cudaMemcpy(tab_dev[dev], sr_synth_sine, sizeof(sr_synth_sine), cudaMemcpyHostToDevice) != cudaSuccess) return -1;
    }
    uint32_t grid = B < 148u * 8u ? B : 148u * 8u;
    synth_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(pcm_dev, U, B, seed_base, nwords, tab_dev[dev]);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

extern "C" int sr_synth_ftr_host(void *out, uint32_t stride, uint32_t B, uint64_t seed_base, uint32_t fmin, uint32_t fmax) {
    if ((!out && B) || stride < 2860u || fmin > fmax || fmax > 119u) return -1;
    for (uint32_t b = 0; b < B; ++b) {
        unsigned char *s = static_cast<unsigned char *>(out) + (size_t)b * stride;
        uint64_t st = seed_base + b;
        const uint16_t sign = 12345, frm = (uint16_t)(fmin + (uint32_t)(splitmix64(st) % (fmax - fmin + 1u)));
        memcpy(s, &sign, 2); memcpy(s + 2, &frm, 2);
        int16_t *d = reinterpret_cast<int16_t *>(s + 4);
        memset(d, 0, 2856);
        for (uint32_t i = 0; i < (uint32_t)frm * 12u; ++i) {
            // sum of four uniforms in [-600,600] ~ bell shaped, sd ~ 690, clipped to +-3000
            int32_t v = 0;
            const uint64_t r = splitmix64(st);
            for (int k = 0; k < 4; ++k) v += (int32_t)((r >> (16 * k)) & 0xFFFFu) % 1201 - 600;
            if (i % 12u == 0) v += 400;
            if (v > 3000) v = 3000;
            if (v < -3000) v = -3000;
            d[i] = (int16_t)v;
        }
    }
    return 0;
}

// ---- WAV -> 12-bit ADC codes (host) ---------------------------------------------------------------------------
extern "C" long sr_wav_to_adc12(const void *wav, size_t wav_bytes, uint16_t *out, size_t max_samples, uint32_t *sample_rate) {
    const unsigned char *p = static_cast<const unsigned char *>(wav);
    auto rd16 = [&](size_t o) { return (uint32_t)p[o] | ((uint32_t)p[o + 1] << 8); };
    auto rd32 = [&](size_t o) { return rd16(o) | (rd16(o + 2) << 16); };
    if (!p || wav_bytes < 12 || memcmp(p, "RIFF", 4) != 0 || memcmp(p + 8, "WAVE", 4) != 0) return -1;
    uint32_t fmt = 0, ch = 0, rate = 0, bits = 0;
    size_t pos = 12, data_off = 0, data_len = 0;
    while (pos + 8 <= wav_bytes) {
        const uint32_t len = rd32(pos + 4);
        if (memcmp(p + pos, "fmt ", 4) == 0 && pos + 8 + 16 <= wav_bytes) {
            fmt = rd16(pos + 8); ch = rd16(pos + 10); rate = rd32(pos + 12); bits = rd16(pos + 22);
        } else if (memcmp(p + pos, "data", 4) == 0) {
            data_off = pos + 8;
            data_len = len;
            if (data_off + data_len > wav_bytes) data_len = wav_bytes - data_off;
            break;
        }
        pos += 8 + (size_t)len + (len & 1);
    }
    if (fmt != 1 || ch == 0 || (bits != 8 && bits != 16) || data_off == 0) return -1;
    if (sample_rate) *sample_rate = rate;
    const size_t frame = (size_t)ch * (bits / 8), n = data_len / frame;
    size_t k = 0;
    for (; k < n && k < max_samples; ++k) {
        const unsigned char *q = p + data_off + k * frame;
        int v;
        if (bits == 16) v = (int)(int16_t)(q[0] | (q[1] << 8)) / 16 + 2048;
        else v = ((int)q[0] - 128) * 16 + 2048;
        out[k] = (uint16_t)(v < 0 ? 0 : (v > 4095 ? 4095 : v));
    }
    return (long)k;
}
