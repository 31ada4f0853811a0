/* sr_synth.h -- deterministic synthetic 8 kHz / 12-bit PCM for benchmarks and tests.
 * NOT part of the reference's call surface: it stands in for the ADC capture (Src/BSP/ADC.C) so
 * that the CPU baseline and the GPU path can be fed byte-identical inputs. Integer-only, so the
 * host and device generators produce the same bytes.
 *
 * Utterance `id` of a batch is generated from seed = seed_base + id (splitmix64 parameter stream):
 *   DC level mid in [1900,2200]; background noise: every 80-sample block holds a random-signed
 *   permutation of {0..79}*na/80 (na in [15,60]) -- uniform amplitude with block-constant energy;
 *   samples [0,2400) are noise only (the 300 ms calibration window of main.c:258);
 *   `nwords` words (first start in [2480,3200), length 2000..3600 samples = 250..450 ms, >= 200 ms
 *   gaps): 3..5 harmonics of f0 in [100,250] Hz, Q8 weights, 20 ms raised-cosine ramps, peak
 *   300..1500 LSB, 10 % white-noise (fricative-like) admixture; a word is dropped if fewer than 1040
 *   noise-only samples would follow it; clip to [0,4095]. */
#ifndef SR_SYNTH_H_
#define SR_SYNTH_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* host generator (multi-threaded, no GPU needed): pcm[B][U] */
int sr_synth_pcm_host(uint16_t *pcm, uint32_t U, uint32_t B, uint64_t seed_base, uint32_t nwords);
/* device generator: pcm_dev is device memory on the current device; asynchronous on `cuda_stream` */
int sr_synth_pcm_dev(uint16_t *pcm_dev, uint32_t U, uint32_t B, uint64_t seed_base, uint32_t nwords, void *cuda_stream);
/* synthetic feature structs for the DTW-only configuration: frm_num in
 * [fmin,fmax], mfcc ~ clipped +-3000 triangular-ish noise of scale 600 with +400 on c0;
 * out = B structs of `stride` bytes (>= 2860), save_sign = 12345. Host only. */
int sr_synth_ftr_host(void *out, uint32_t stride, uint32_t B, uint64_t seed_base, uint32_t fmin, uint32_t fmax);
/* WAV ingestion: RIFF/WAVE PCM, 8-bit unsigned or 16-bit signed, mono or interleaved (channel 0 is
 * taken) -> the 12-bit unsigned ADC codes the path consumes: 16-bit x -> x/16 + 2048 (C truncation), 8-bit
 * x -> (x-128)*16 + 2048. Host only (input adaptation, like the generator above). Returns the number of samples
 * written (<= max_samples), or -1 on a malformed / unsupported file; *sample_rate receives the file's rate. */
long sr_wav_to_adc12(const void *wav, size_t wav_bytes, uint16_t *out, size_t max_samples, uint32_t *sample_rate);
#ifdef __cplusplus
}
#endif
#endif
