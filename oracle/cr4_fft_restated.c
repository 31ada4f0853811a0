/* ORACLE -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * C restatement of the one file of the hot path that cannot be compiled on a host:
 *   Src/BSP/cr4_fft_1024_stm32.s:95-281  (ST MCD DSP-library V2.0.0 radix-4 1024-pt complex FFT,
 *   ARM Thumb-2 assembly), exporting the same symbol and argument order:
 *       void cr4_fft_1024_stm32(void *pssOUT, void *pssIN, u16 Nbin)      (MFCC.C:12)
 * Every register of the asm is modelled as a 32-bit two's-complement value with wrap-around
 * (uint32_t arithmetic), `ASR` is an arithmetic shift, LDRSH sign-extends, STRH keeps 16 bits.
 *
 * The twiddle table comes from the header named by SR_TWIDDLE_HEADER, which must define
 *     static const int16_t SR_TWIDDLE_NAME[2040];
 *   - oracle/_ref build: oracle/_ref/twiddle_ref.h, extracted by oracle/extract_ref_tables.py
 *     from the DCW lines of cr4_fft_1024_stm32.s:285-629 (the reference's own numbers);
 *   - oracle port build: the closed-form table of tools/gen_tables.py.
 */
#include <stdint.h>
#include SR_TWIDDLE_HEADER

#define ASR(x, n) ((uint32_t)((int32_t)(x) >> (n)))
#define SX16(x)   ((uint32_t)(int32_t)(int16_t)((x) & 0xFFFFu))

static inline uint32_t pack16(uint32_t re, uint32_t im) { return (re & 0xFFFFu) | (im << 16); }

/* RBIT + LSR#22 of an 8-bit counter (.s:227,230): element index = 8-bit bit reversal */
static inline unsigned bitrev8(unsigned v) {
    v = ((v & 0xF0u) >> 4) | ((v & 0x0Fu) << 4);
    v = ((v & 0xCCu) >> 2) | ((v & 0x33u) << 2);
    v = ((v & 0xAAu) >> 1) | ((v & 0x55u) << 1);
    return v;
}

/* CXMUL_V7 (.s:95-102): (Zr,Zi) = (Yr,Yi) * conj-twiddle, 32-bit products, no shift */
static inline void cxmul(uint32_t *zr, uint32_t *zi, uint32_t yr, uint32_t yi, uint32_t ka, uint32_t kb) {
    uint32_t t = (yi - yr) * kb;        /* SUB, MUL            */
    uint32_t k2 = ka + (kb << 1);       /* ADD Kr, Ki, LSL#1   */
    *zi = yi * ka + t;                  /* MLA                 */
    *zr = yr * k2 + t;                  /* MLA                 */
}

/* CXADDA4 $s (.s:105-129) and the identical tree of BUTFLY4ZERO_OPT (.s:147-168, s=0).
 * In/out: A,B,C,D register pairs. On return leg-3 lives swapped: real part in *Di, imag in *Dr,
 * which is how the asm stores it ("inversion here", .s:176-177, 203-204). */
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

void cr4_fft_1024_stm32(void *pssOUT, void *pssIN, uint16_t Nbin) {
    uint32_t *out = (uint32_t *)pssOUT;
    const uint32_t *in = (const uint32_t *)pssIN;
    (void)Nbin;                                   /* the routine only handles 1024 (.s:214-215) */

    /* preloop_v7 (.s:226-232): 256 x BUTFLY4ZERO_OPT, bit-reversed gather; note the load order
     * A=x[j], C=x[j+256], B=x[j+512], D=x[j+768] (.s:134-145) */
    for (unsigned idx = 0; idx < 256; ++idx) {
        unsigned j = bitrev8(idx);
        uint32_t Ar = SX16(in[j]),       Ai = SX16(in[j] >> 16);
        uint32_t Cr = SX16(in[j + 256]), Ci = SX16(in[j + 256] >> 16);
        uint32_t Br = SX16(in[j + 512]), Bi = SX16(in[j + 512] >> 16);
        uint32_t Dr = SX16(in[j + 768]), Di = SX16(in[j + 768] >> 16);
        cxadda4(&Ar, &Ai, &Br, &Bi, &Cr, &Ci, &Dr, &Di, 0);
        out[4 * idx + 0] = pack16(Ar, Ai);
        out[4 * idx + 1] = pack16(Br, Bi);
        out[4 * idx + 2] = pack16(Cr, Ci);
        out[4 * idx + 3] = pack16(Di, Dr);        /* inversion here */
    }

    /* passloop_v7 / grouploop_v7 / butterloop_v7 (.s:254-279): strides 4,16,64,256 elements,
     * twiddle block of `s` triples per pass, rewound for every group (.s:271-273) */
    const int16_t *K = SR_TWIDDLE_NAME;
    for (unsigned s = 4; s <= 256; s <<= 2) {
        for (unsigned base = 0; base < 1024; base += 4 * s) {
            for (unsigned q = 0; q < s; ++q) {
                const int16_t *k = K + 6 * q;
                uint32_t *p0 = out + base + q, *p1 = p0 + s, *p2 = p1 + s, *p3 = p2 + s;
                uint32_t Ar, Ai, Br, Bi, Cr, Ci, Dr, Di;
                /* BUTFLY4_V7 (.s:180-205): leg3*K[0..1] -> D, leg2*K[2..3] -> C, leg1*K[4..5] -> B */
                cxmul(&Dr, &Di, SX16(*p3), SX16(*p3 >> 16), (uint32_t)(int32_t)k[0], (uint32_t)(int32_t)k[1]);
                cxmul(&Cr, &Ci, SX16(*p2), SX16(*p2 >> 16), (uint32_t)(int32_t)k[2], (uint32_t)(int32_t)k[3]);
                cxmul(&Br, &Bi, SX16(*p1), SX16(*p1 >> 16), (uint32_t)(int32_t)k[4], (uint32_t)(int32_t)k[5]);
                Ar = SX16(*p0); Ai = SX16(*p0 >> 16);
                cxadda4(&Ar, &Ai, &Br, &Bi, &Cr, &Ci, &Dr, &Di, 14);
                *p0 = pack16(Ar, Ai);
                *p1 = pack16(Br, Bi);
                *p2 = pack16(Cr, Ci);
                *p3 = pack16(Di, Dr);             /* inversion here */
            }
        }
        K += 6 * s;
    }
}
