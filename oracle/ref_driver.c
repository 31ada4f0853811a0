/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Thin batch driver compiled INTO oracle/_ref/libref.so next to the reference's own, unmodified
 * Src/Speech_Recog/{VAD,MFCC,DTW}.C. It adds no arithmetic of its own: it only calls the
 * reference functions in the order their one caller does (Src/APP/main.c:249-296 spch_recg and
 * main.c:121-138 save_mdl) so that Python tests / bench.py can run whole shards through the
 * reference without a ctypes round trip per call, and converts valid_tag pointers to sample
 * offsets (0xFFFFFFFF = NULL). The reference keeps statics (MFCC.C:14-15, DTW.C:65-68): this
 * library is NOT re-entrant -- parallel runs use one process per core. */
#include <stddef.h>
#include <string.h>
#include "stm32f10x.h"
#include "ADC.h"
#include "VAD.H"
#include "MFCC.H"
#include "DTW.H"
#include SR_REF_FLASH_H

u32 *fft(s16 *dat_buf, u16 buf_len);           /* MFCC.C:27 */
u32 get_dis(s16 *frm_ftr1, s16 *frm_ftr2);     /* DTW.C:45  */
u8 dtw_limit(u16 x, u16 y);                    /* DTW.C:76  */

#define REF_NULL 0xFFFFFFFFu

unsigned ref_sizeof_ftr(void) { return (unsigned)sizeof(v_ftr_tag); }
unsigned ref_const(int which) {
    switch (which) {
    case 0: return fs;          case 1: return VcBuf_Len;   case 2: return atap_len;
    case 3: return frame_len;   case 4: return frame_mov;   case 5: return max_vc_con;
    case 6: return fft_point;   case 7: return mfcc_num;    case 8: return vv_frm_max;
    case 9: return save_mask;   case 10: return size_per_ftr; case 11: return ftr_per_comm;
    case 12: return tri_num;
    default: return 0;
    }
}

static void seg_to_off(const u16 *base, const valid_tag *v, u32 *off6) {
    for (int i = 0; i < max_vc_con; ++i) {
        off6[2 * i + 0] = v[i].start ? (u32)(v[i].start - base) : REF_NULL;
        off6[2 * i + 1] = v[i].end ? (u32)(v[i].end - base) : REF_NULL;
    }
}

/* noise_atap + VAD on one buffer; atap is in/out exactly like the reference (noise_atap leaves it
 * untouched when n_len % 240 != 0, VAD.C:33-36) */
void ref_vad(const u16 *pcm, u32 buf_len, u32 n_len, atap_tag *atap, u32 *seg_off6) {
    valid_tag vv[max_vc_con];
    noise_atap(pcm, (u16)n_len, atap);
    VAD(pcm, (u16)buf_len, vv, atap);
    seg_to_off(pcm, vv, seg_off6);
}

void ref_mfcc_seg(const u16 *pcm, u32 start, u32 end, atap_tag *atap, v_ftr_tag *out) {
    valid_tag v;
    v.start = (u16 *)pcm + start;
    v.end = (u16 *)pcm + end;
    get_mfcc(&v, out, atap);
}

/* spch_recg (main.c:249-296) with the buffer length, noise window and template bank passed in
 * instead of the firmware's globals/flash addresses. Returns 0 ok, 1 VAD fail, 2 MFCC fail. */
int ref_recognise(const u16 *pcm, u32 buf_len, u32 n_len, const u8 *bank, u32 n_slot, u32 slot_stride,
                  atap_tag *atap_out, u32 *seg_off6, v_ftr_tag *ftr_out, u32 *score /*[n_slot] or NULL*/,
                  u32 *best_idx, u32 *best_dis, u32 *cmd) {
    atap_tag atap;
    valid_tag vv[max_vc_con];
    static v_ftr_tag ftr;
    memset(&atap, 0, sizeof atap);
    noise_atap(pcm, (u16)n_len, &atap);
    VAD(pcm, (u16)buf_len, vv, &atap);
    if (atap_out) *atap_out = atap;
    if (seg_off6) seg_to_off(pcm, vv, seg_off6);
    *best_idx = 0; *cmd = 0; *best_dis = dis_err;
    if (vv[0].end == NULL) return 1;
    get_mfcc(&vv[0], &ftr, &atap);
    if (ftr_out) memcpy(ftr_out, &ftr, sizeof ftr);
    if (ftr.frm_num == 0) return 2;
    u32 min_dis = dis_max, min_i = 0;
    for (u32 i = 0; i < n_slot; ++i) {
        v_ftr_tag *mdl = (v_ftr_tag *)(bank + (size_t)i * slot_stride);
        u32 cur = (mdl->save_sign == save_mask) ? dtw(&ftr, mdl) : dis_err;
        if (score) score[i] = cur;
        if (cur < min_dis) { min_dis = cur; min_i = i; }
    }
    *best_idx = min_i; *best_dis = min_dis; *cmd = min_i / ftr_per_comm;
    return 0;
}

/* shard loop for timing/parity: utterances [0,B) of stride U samples */
void ref_recognise_batch(const u16 *pcm, u32 U, u32 B, u32 n_len, const u8 *bank, u32 n_slot, u32 slot_stride,
                         u32 *seg_off /*[B][6]*/, v_ftr_tag *ftr /*[B] or NULL*/, u32 *score /*[B][n_slot] or NULL*/,
                         u32 *best_idx, u32 *best_dis, u32 *cmd, u8 *status) {
    for (u32 b = 0; b < B; ++b) {
        int st = ref_recognise(pcm + (size_t)b * U, U, n_len, bank, n_slot, slot_stride, NULL,
                               seg_off ? seg_off + 6 * (size_t)b : NULL, ftr ? ftr + b : NULL,
                               score ? score + (size_t)b * n_slot : NULL, best_idx + b, best_dis + b, cmd + b);
        if (status) status[b] = (u8)st;
    }
}

/* fixed-segment MFCC over a shard (BASELINE config 2: segment [start,end) identical for all) */
void ref_mfcc_batch(const u16 *pcm, u32 U, u32 B, const u32 *seg /*[B][2]*/, const atap_tag *atap /*[B]*/,
                    v_ftr_tag *ftr /*[B]*/) {
    for (u32 b = 0; b < B; ++b) {
        atap_tag a = atap[b];
        ref_mfcc_seg(pcm + (size_t)b * U, seg[2 * b], seg[2 * b + 1], &a, ftr + b);
    }
}

/* all-pairs dtw over a shard of feature structs (BASELINE config 3); also counts get_dis cells */
void ref_dtw_batch(const v_ftr_tag *in, u32 B, const u8 *bank, u32 n_slot, u32 slot_stride, int check_sign,
                   u32 *score /*[B][n_slot]*/) {
    for (u32 b = 0; b < B; ++b)
        for (u32 t = 0; t < n_slot; ++t) {
            v_ftr_tag *mdl = (v_ftr_tag *)(bank + (size_t)t * slot_stride);
            score[(size_t)b * n_slot + t] =
                (!check_sign || mdl->save_sign == save_mask) ? dtw((v_ftr_tag *)(in + b), mdl) : dis_err;
        }
}

/* isolated FFT / magnitude for unit parity: raw asm-restatement output and fft() magnitudes */
void cr4_fft_1024_stm32(void *pssOUT, void *pssIN, u16 Nbin);
void ref_fft_raw(const u32 *in1024, u32 *out1024) { cr4_fft_1024_stm32(out1024, (void *)in1024, 1024); }
void ref_fft_mag(const s16 *frame, u32 len, u32 *mag512) {
    u32 *p = fft((s16 *)frame, (u16)len);
    if (p) memcpy(mag512, p, 512 * sizeof(u32));
}
