/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement of the reference's VAD -> MFCC -> DTW path (the .C files of Src/Speech_Recog,
 * Src/BSP/cr4_fft_1024_stm32.s, Src/APP/main.c:249-296). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this; the product library
 * (libspeech_b200.so) never links, loads or calls it.
 * Parity pinning: validated bit-for-bit against oracle/_ref/libref.so (the reference's own,
 * unmodified C sources compiled for the host) by tests/test_oracle.py, and against the golden
 * vectors in tests/golden/ generated from libref.so. */
#ifndef SR_ORACLE_H_
#define SR_ORACLE_H_
#include <stdint.h>

#define SRO_MAX_VC_CON 3          /* VAD.H:4   */
#define SRO_FRAME_LEN 160         /* VAD.H:7   (20 ms @ 8 kHz, ADC.H:7) */
#define SRO_FRAME_MOV 80          /* VAD.H:8   */
#define SRO_FFT_POINT 1024        /* MFCC.H:8  */
#define SRO_FRQ_MAX 512           /* MFCC.H:9  */
#define SRO_TRI_NUM 24            /* MFCC.H:12 */
#define SRO_MFCC_NUM 12           /* MFCC.H:13 */
#define SRO_VV_FRM_MAX 119        /* MFCC.H:15-16 */
#define SRO_DIS_ERR 0xFFFFFFFFu   /* DTW.H:4   */
#define SRO_SAVE_MASK 12345       /* Flash.H:11 */
#define SRO_FTR_PER_COMM 4        /* Flash.H:15 */
#define SRO_NULL 0xFFFFFFFFu      /* offset encoding of a NULL valid_tag pointer */

typedef struct { uint32_t mid_val; uint16_t n_thl; uint16_t z_thl; uint32_t s_thl; } sro_atap; /* VAD.H:10-16 */
#pragma pack(push, 1)
typedef struct { uint16_t save_sign; uint16_t frm_num; int16_t mfcc_dat[SRO_VV_FRM_MAX * SRO_MFCC_NUM]; } sro_ftr; /* MFCC.H:18-25 */
#pragma pack(pop)

void     sro_noise_atap(const uint16_t *noise, uint32_t n_len, sro_atap *atap);
void     sro_vad(const uint16_t *vc, uint32_t buf_len, const sro_atap *atap, uint32_t *seg_off6);
void     sro_fft_raw(const uint32_t *in1024, uint32_t *out1024);
void     sro_fft_mag(const int16_t *frame, uint32_t len, uint32_t *mag512);
void     sro_mfcc(const uint16_t *pcm, uint32_t start, uint32_t end, const sro_atap *atap, sro_ftr *out);
uint32_t sro_get_dis(const int16_t *a, const int16_t *b);
int      sro_dtw_limit(int x, int y, int I, int M);
uint32_t sro_dtw(const sro_ftr *in, const sro_ftr *mdl, uint32_t *cells);
uint32_t sro_get_mdl(const sro_ftr *f1, const sro_ftr *f2, sro_ftr *fm);
uint32_t sro_dtw_band(const sro_ftr *in, const sro_ftr *mdl, int r, uint32_t *cells);
int      sro_recognise(const uint16_t *pcm, uint32_t buf_len, uint32_t n_len, const uint8_t *bank, uint32_t n_slot,
                       uint32_t slot_stride, sro_atap *atap_out, uint32_t *seg_off6, sro_ftr *ftr_out,
                       uint32_t *score, uint32_t *best_idx, uint32_t *best_dis, uint32_t *cmd);
/* batch forms; nthreads>1 splits the shard over pthreads (the port is re-entrant) */
void sro_recognise_batch(const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, const uint8_t *bank,
                         uint32_t n_slot, uint32_t slot_stride, sro_atap *atap, uint32_t *seg_off, sro_ftr *ftr,
                         uint32_t *score, uint32_t *best_idx, uint32_t *best_dis, uint32_t *cmd, uint8_t *status,
                         int nthreads);
void sro_mfcc_batch(const uint16_t *pcm, uint32_t U, uint32_t B, const uint32_t *seg2, const sro_atap *atap,
                    sro_ftr *ftr, int nthreads);
void sro_dtw_batch(const sro_ftr *in, uint32_t B, const uint8_t *bank, uint32_t n_slot, uint32_t slot_stride,
                   int check_sign, int band_r /* <0: greedy reference walk */, uint32_t *score,
                   uint64_t *cells_total, int nthreads);
uint32_t sro_log100(uint32_t v);
#endif
