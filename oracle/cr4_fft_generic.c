/* ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * The radix-4 FFT of Src/BSP/cr4_fft_1024_stm32.s:95-281 written for a general power-of-four size N (64, 256, 1024):
 * the same macros (CXMUL_V7 .s:95-102, CXADDA4 .s:105-129, BUTFLY4ZERO_OPT .s:132-178, BUTFLY4_V7 .s:180-205), the same
 * pass structure (bit-reversed first pass, then strides 4, 16, .. N/4 with the twiddle block of each stride), and the
 * twiddle table TableFFT_V7 whose blocks for strides 4, 16, 64 are shared by every size (the table is cumulative:
 * .s:287-290 N=16, :292-307 N=64, :309-372 N=256, :374-629 N=1024).
 *   - N = 1024 must reproduce cr4_fft_restated.c bit for bit (tests/test_oracle.py): that pins the generalisation;
 *   - N = 256 is the FFT of the GEOM_B extension (BASELINE configs[0] "256-pt"). The reference ships no 256-point
 *     routine, so everything built on it is PARITY UNPINNED: this file is its only checker. */
#include <stdint.h>
#include SR_TWIDDLE_HEADER

#define ASR(x, n) ((uint32_t)((int32_t)(x) >> (n)))
#define SX16(x)   ((uint32_t)(int32_t)(int16_t)((x) & 0xFFFFu))

static inline uint32_t pack16(uint32_t re, uint32_t im) { return (re & 0xFFFFu) | (im << 16); }

static inline void cxmul(uint32_t *zr, uint32_t *zi, uint32_t yr, uint32_t yi, uint32_t ka, uint32_t kb) {
    uint32_t t = (yi - yr) * kb;
    uint32_t k2 = ka + (kb << 1);
    *zi = yi * ka + t;
    *zr = yr * k2 + t;
}

static inline void cxadda4(uint32_t *Ar, uint32_t *Ai, uint32_t *Br, uint32_t *Bi,
                           uint32_t *Cr, uint32_t *Ci, uint32_t *Dr, uint32_t *Di, int s) {
    *Cr = *Cr + *Dr;            *Ci = *Ci + *Di;
    *Dr = *Cr - (*Dr << 1);     *Di = *Ci - (*Di << 1);
    *Ar = ASR(*Ar, 2);          *Ai = ASR(*Ai, 2);
    *Ar = *Ar + ASR(*Br, 2 + s); *Ai = *Ai + ASR(*Bi, 2 + s);
    *Br = *Ar - ASR(*Br, 1 + s); *Bi = *Ai - ASR(*Bi, 1 + s);
    *Ar = *Ar + ASR(*Cr, 2 + s); *Ai = *Ai + ASR(*Ci, 2 + s);
    *Cr = *Ar - ASR(*Cr, 1 + s); *Ci = *Ai - ASR(*Ci, 1 + s);
    *Br = *Br + ASR(*Di, 2 + s);
    *Bi = *Bi - ASR(*Dr, 2 + s);
    *Di = *Br - ASR(*Di, 1 + s);
    *Dr = *Bi + ASR(*Dr, 1 + s);
}

/* N in {64, 256, 1024}; in/out: N packed (re | im << 16) words */
int sro_cr4_fft(uint32_t *out, const uint32_t *in, unsigned N) {
    unsigned bits = 0;
    if (N != 64 && N != 256 && N != 1024) return -1;
    while ((4u << bits) < N) ++bits;                       /* log2(N/4) */
    const unsigned Q = N / 4;
    for (unsigned idx = 0; idx < Q; ++idx) {
        unsigned j = 0;
        for (unsigned b = 0; b < bits; ++b) if (idx & (1u << b)) j |= 1u << (bits - 1 - b);
        uint32_t Ar = SX16(in[j]),         Ai = SX16(in[j] >> 16);
        uint32_t Cr = SX16(in[j + Q]),     Ci = SX16(in[j + Q] >> 16);
        uint32_t Br = SX16(in[j + 2 * Q]), Bi = SX16(in[j + 2 * Q] >> 16);
        uint32_t Dr = SX16(in[j + 3 * Q]), Di = SX16(in[j + 3 * Q] >> 16);
        cxadda4(&Ar, &Ai, &Br, &Bi, &Cr, &Ci, &Dr, &Di, 0);
        out[4 * idx + 0] = pack16(Ar, Ai);
        out[4 * idx + 1] = pack16(Br, Bi);
        out[4 * idx + 2] = pack16(Cr, Ci);
        out[4 * idx + 3] = pack16(Di, Dr);
    }
    const int16_t *K = SR_TWIDDLE_NAME;
    for (unsigned s = 4; s <= Q; s <<= 2) {
        for (unsigned base = 0; base < N; base += 4 * s) {
            for (unsigned q = 0; q < s; ++q) {
                const int16_t *k = K + 6 * q;
                uint32_t *p0 = out + base + q, *p1 = p0 + s, *p2 = p1 + s, *p3 = p2 + s;
                uint32_t Ar, Ai, Br, Bi, Cr, Ci, Dr, Di;
                cxmul(&Dr, &Di, SX16(*p3), SX16(*p3 >> 16), (uint32_t)(int32_t)k[0], (uint32_t)(int32_t)k[1]);
                cxmul(&Cr, &Ci, SX16(*p2), SX16(*p2 >> 16), (uint32_t)(int32_t)k[2], (uint32_t)(int32_t)k[3]);
                cxmul(&Br, &Bi, SX16(*p1), SX16(*p1 >> 16), (uint32_t)(int32_t)k[4], (uint32_t)(int32_t)k[5]);
                Ar = SX16(*p0); Ai = SX16(*p0 >> 16);
                cxadda4(&Ar, &Ai, &Br, &Bi, &Cr, &Ci, &Dr, &Di, 14);
                *p0 = pack16(Ar, Ai);
                *p1 = pack16(Br, Bi);
                *p2 = pack16(Cr, Ci);
                *p3 = pack16(Di, Dr);
            }
        }
        K += 6 * s;
    }
    return 0;
}
