/* ORACLE -- TEST INFRASTRUCTURE ONLY (see sr_oracle.h). Plain C restatement, one function per
 * reference function, each citing the file:line it follows. Re-entrant (the reference's statics
 * fft_in/fft_out MFCC.C:14-15 and X1/X2/in_frm_num/mdl_frm_num DTW.C:65-68 become locals). */
#include "sr_oracle.h"
#include "sr_tables.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

void cr4_fft_1024_stm32(void *pssOUT, void *pssIN, uint16_t Nbin);   /* cr4_fft_restated.c */

/* ---- VAD.C:22-71 noise_atap ------------------------------------------------------------------
 * atap_frm_len = (8000/1000)*30 = 240 (VAD.C:13-14); ratio macros evaluated left to right at the
 * use site: s_thl = abs_sum*11/10 (VAD.C:9,69), z_thl = 160*2/160/1 = 2 (VAD.C:10,70). */
void sro_noise_atap(const uint16_t *noise, uint32_t n_len, sro_atap *atap) {
    n_len &= 0xFFFFu;                                   /* u16 parameter */
    if (n_len % 240u != 0 || n_len == 0) return;        /* VAD.C:33-36 (n_len==0 divides by zero in the reference) */
    uint32_t frm_num = n_len / 240u, n_sum = 0, max_sum = 0, abs_sum = 0;
    for (uint32_t i = 0; i < n_len; ++i) n_sum += noise[i];
    uint32_t mid = n_sum / n_len;                       /* VAD.C:45 */
    for (uint32_t i = 0; i < n_len; i += 240u) {        /* VAD.C:48-63 */
        uint32_t n_max = 0;
        for (uint32_t h = 0; h < 240u; ++h) {
            uint32_t v = noise[i + h], a = v > mid ? v - mid : mid - v;
            if (a > n_max) n_max = a;
            abs_sum += a;
        }
        max_sum += n_max;
    }
    abs_sum /= (n_len / SRO_FRAME_LEN);                 /* VAD.C:65 */
    max_sum /= frm_num;                                 /* VAD.C:66 */
    atap->mid_val = mid;
    atap->n_thl = (uint16_t)(max_sum * 1u);             /* VAD.C:68 */
    atap->s_thl = abs_sum * 11u / 10u;                  /* VAD.C:69 */
    atap->z_thl = (uint16_t)(SRO_FRAME_LEN * 2 / 160 / 1); /* VAD.C:70 */
}

/* ---- VAD.C:97-218 VAD ------------------------------------------------------------------------
 * Segments as sample offsets (SRO_NULL = NULL pointer). v_durmin_f = 80/(20-10) = 8 and
 * s_durmax_f = 110/(20-10) = 11 (VAD.C:72-75). last_sig is never reset between frames (VAD.C:99). */
void sro_vad(const uint16_t *vc, uint32_t buf_len, const sro_atap *atap, uint32_t *seg) {
    uint32_t last_sig = 0, cur = 0, front = 0, back = 0, valid_con = 0;
    uint32_t mid = atap->mid_val;
    uint32_t a_thl = mid + atap->n_thl, b_thl = mid - atap->n_thl;   /* VAD.C:112-113 (u32 wrap) */
    buf_len &= 0xFFFFu;
    for (int i = 0; i < 6; ++i) seg[i] = SRO_NULL;                   /* VAD.C:115-119 */
    uint32_t lim = (uint32_t)((int)buf_len - SRO_FRAME_LEN);         /* VAD.C:121: int -> u32 compare */
    for (uint32_t i = 0; i < lim; i += SRO_FRAME_MOV) {
        uint32_t frm_sum = 0, frm_zero = 0;
        for (uint32_t h = 0; h < SRO_FRAME_LEN; ++h) {               /* VAD.C:126-129 */
            uint32_t v = vc[i + h];
            frm_sum += v > mid ? v - mid : mid - v;
        }
        for (uint32_t h = 0; h < SRO_FRAME_LEN - 1; ++h) {           /* VAD.C:132-157 */
            uint32_t v = vc[i + h], w = vc[i + h + 1];
            if (v >= a_thl) last_sig = 2; else if (v < b_thl) last_sig = 1;
            if (w >= a_thl) { if (last_sig == 1) ++frm_zero; }
            else if (w < b_thl) { if (last_sig == 2) ++frm_zero; }
        }
        if (frm_sum > atap->s_thl || frm_zero > atap->z_thl) {       /* VAD.C:164-187 */
            if (cur == 0) { cur = 1; front = 1; }
            else if (cur == 1) {
                if (++front >= 8) { cur = 2; seg[2 * valid_con] = i - 7 * SRO_FRAME_MOV; front = 0; }
            } else if (cur == 3) { back = 0; cur = 2; }
        } else {                                                     /* VAD.C:188-216 */
            if (cur == 2) { cur = 3; back = 1; }
            else if (cur == 3) {
                if (++back >= 11) {
                    cur = 0;
                    seg[2 * valid_con + 1] = i - 11 * SRO_FRAME_MOV + SRO_FRAME_LEN;
                    if (++valid_con == SRO_MAX_VC_CON) return;
                    back = 0;
                }
            } else if (cur == 1) { front = 0; cur = 0; }
        }
    }
}

/* ---- MFCC.C:27-62 fft ------------------------------------------------------------------------ */
void sro_fft_raw(const uint32_t *in1024, uint32_t *out1024) {
    cr4_fft_1024_stm32(out1024, (void *)in1024, 1024);
}
void sro_fft_mag(const int16_t *frame, uint32_t len, uint32_t *mag) {
    uint32_t in[SRO_FFT_POINT], out[SRO_FFT_POINT];
    if (len > SRO_FFT_POINT) return;                                 /* MFCC.C:32-35 */
    for (uint32_t i = 0; i < len; ++i) in[i] = (uint16_t)frame[i];   /* MFCC.C:37-41: im=0, re=low half */
    for (uint32_t i = len; i < SRO_FFT_POINT; ++i) in[i] = 0;        /* MFCC.C:42-45 */
    cr4_fft_1024_stm32(out, in, SRO_FFT_POINT);                      /* MFCC.C:47 */
    for (uint32_t i = 0; i < SRO_FRQ_MAX; ++i) {                     /* MFCC.C:49-60 */
        int32_t re = (int16_t)(out[i]), im = (int16_t)(out[i] >> 16);
        int32_t pw = (int32_t)((uint32_t)(re * re) + (uint32_t)(im * im));
        float f = sqrtf((float)pw) * 10;                             /* float*int -> float multiply */
        mag[i] = (pw < 0) ? 0u : (uint32_t)f;                        /* NaN -> 0 as x86-64 cvttss2si does */
    }
}

/* (u32)(log((double)v)*100), MFCC.C:168. log(0) = -inf converts to 0 on x86-64 and ARM alike. */
uint32_t sro_log100(uint32_t v) {
    if (v == 0) return 0;
    return (uint32_t)(log((double)v) * 100);
}

/* ---- MFCC.C:86-191 get_mfcc ------------------------------------------------------------------
 * start/end are sample offsets into pcm. hp_ratio 95/100 (MFCC.H:7), hamm_top/10 = 1000
 * (MFCC.H:10, MFCC.C:122), tri_top/10 = 100 (MFCC.H:11, MFCC.C:139). */
void sro_mfcc(const uint16_t *pcm, uint32_t start, uint32_t end, const sro_atap *atap, sro_ftr *out) {
    /* MFCC.C:102: ((u32)end-(u32)start)/2 on byte addresses, then u16 truncation */
    uint32_t nbytes = (uint32_t)(2u * end) - (uint32_t)(2u * start);
    uint16_t v_frm_num = (uint16_t)((nbytes / 2u - SRO_FRAME_LEN) / SRO_FRAME_MOV + 1u);
    if (v_frm_num > SRO_VV_FRM_MAX) { out->frm_num = 0; return; }   /* MFCC.C:103-107 */
    int32_t mid = (int32_t)atap->mid_val;
    int16_t *mfcc_p = out->mfcc_dat;
    uint16_t frm_con = 0;
    /* MFCC.C:113: for (vc_dat=start; vc_dat <= end-frame_len; vc_dat += 80) -- signed offsets */
    for (int64_t p = (int64_t)start; p <= (int64_t)end - SRO_FRAME_LEN; p += SRO_FRAME_MOV) {
        int16_t w[SRO_FRAME_LEN];
        uint32_t mag[SRO_FRQ_MAX], pow_spct[SRO_TRI_NUM];
        for (int i = 0; i < SRO_FRAME_LEN; ++i) {                    /* MFCC.C:115-124 */
            int32_t t = ((int32_t)pcm[p + i] - mid) - ((int32_t)pcm[p + i - 1] - mid) * 95 / 100;
            w[i] = (int16_t)(t * (int32_t)sr_tab_hamm[i] / 1000);
        }
        sro_fft_mag(w, SRO_FRAME_LEN, mag);                          /* MFCC.C:126 */
        for (int i = 0; i < SRO_FRQ_MAX; ++i) mag[i] *= mag[i];      /* MFCC.C:128-133 (u32 wrap) */
        const uint16_t *cen = sr_tab_tri_cen;
        pow_spct[0] = 0;                                             /* MFCC.C:136-140 */
        for (uint32_t i = 0; i < cen[1]; ++i) pow_spct[0] += mag[i] * sr_tab_tri_even[i] / 100u;
        for (int h = 2; h < SRO_TRI_NUM; h += 2) {                   /* MFCC.C:141-148 */
            pow_spct[h] = 0;
            for (uint32_t i = cen[h - 1]; i < cen[h + 1]; ++i) pow_spct[h] += mag[i] * sr_tab_tri_even[i] / 100u;
        }
        for (int h = 1; h < SRO_TRI_NUM - 2; h += 2) {               /* MFCC.C:150-157 */
            pow_spct[h] = 0;
            for (uint32_t i = cen[h - 1]; i < cen[h + 1]; ++i) pow_spct[h] += mag[i] * sr_tab_tri_odd[i] / 100u;
        }
        pow_spct[SRO_TRI_NUM - 1] = 0;                               /* MFCC.C:158-162 */
        for (uint32_t i = cen[SRO_TRI_NUM - 2]; i < SRO_FRQ_MAX; ++i)
            pow_spct[SRO_TRI_NUM - 1] += mag[i] * sr_tab_tri_odd[i] / 100u;
        for (int h = 0; h < SRO_TRI_NUM; ++h) pow_spct[h] = sro_log100(pow_spct[h]);   /* MFCC.C:165-170 */
        const int8_t *dct = sr_tab_dct;                              /* MFCC.C:173-183 */
        for (int c = 0; c < SRO_MFCC_NUM; ++c) {
            int16_t acc = 0;
            for (int i = 0; i < SRO_TRI_NUM; ++i)
                acc = (int16_t)(acc + ((int32_t)pow_spct[i]) * ((int32_t)dct[i]) / 100);
            mfcc_p[c] = acc;
            dct += SRO_TRI_NUM;
        }
        mfcc_p += SRO_MFCC_NUM;
        ++frm_con;
    }
    out->frm_num = frm_con;                                          /* MFCC.C:189 */
}

/* ---- GEOM_B extension (BASELINE configs[0]: 25 ms frames / 10 ms hop / 256-point FFT) ---------------------------------
 * get_mfcc (MFCC.C:86-191) with frame_len = 200, fft_point = 256, frq_max = 128 and the tables the reference's Matlab
 * formulas give for that geometry (tools/gen_tables.py; speech_recog.m:217-313); everything else -- pre-emphasis 95/100,
 * hamm/1000, magnitude*10, energy, tri/100, log*100, DCT/100 with s16 accumulation, vv_frm_max -- as in MFCC.C.
 * PARITY UNPINNED: the reference has no 256-point path; this function is the only checker of the CUDA GEOM_B kernel. */
int sro_cr4_fft(uint32_t *out, const uint32_t *in, unsigned N);          /* cr4_fft_generic.c */
#define SRO_B_FRAME_LEN 200
#define SRO_B_FFT_POINT 256
#define SRO_B_FRQ_MAX 128
void sro_fft_raw_n(const uint32_t *in, uint32_t *out, uint32_t N) { sro_cr4_fft(out, in, N); }
void sro_mfcc_geom_b(const uint16_t *pcm, uint32_t start, uint32_t end, const sro_atap *atap, sro_ftr *out) {
    uint32_t nbytes = (uint32_t)(2u * end) - (uint32_t)(2u * start);
    uint16_t v_frm_num = (uint16_t)((nbytes / 2u - SRO_B_FRAME_LEN) / SRO_FRAME_MOV + 1u);
    if (v_frm_num > SRO_VV_FRM_MAX) { out->frm_num = 0; return; }
    int32_t mid = (int32_t)atap->mid_val;
    int16_t *mfcc_p = out->mfcc_dat;
    uint16_t frm_con = 0;
    for (int64_t p = (int64_t)start; p <= (int64_t)end - SRO_B_FRAME_LEN; p += SRO_FRAME_MOV) {
        uint32_t in[SRO_B_FFT_POINT], fo[SRO_B_FFT_POINT], mag[SRO_B_FRQ_MAX], pow_spct[SRO_TRI_NUM];
        for (int i = 0; i < SRO_B_FRAME_LEN; ++i) {
            int32_t t = ((int32_t)pcm[p + i] - mid) - ((int32_t)pcm[p + i - 1] - mid) * 95 / 100;
            in[i] = (uint16_t)(int16_t)(t * (int32_t)sr_tab_b_hamm[i] / 1000);
        }
        for (int i = SRO_B_FRAME_LEN; i < SRO_B_FFT_POINT; ++i) in[i] = 0;
        sro_cr4_fft(fo, in, SRO_B_FFT_POINT);
        for (int i = 0; i < SRO_B_FRQ_MAX; ++i) {
            int32_t re = (int16_t)(fo[i]), im = (int16_t)(fo[i] >> 16);
            int32_t pw = (int32_t)((uint32_t)(re * re) + (uint32_t)(im * im));
            float f = sqrtf((float)pw) * 10;
            mag[i] = (pw < 0) ? 0u : (uint32_t)f;
            mag[i] *= mag[i];
        }
        const uint16_t *cen = sr_tab_b_tri_cen;
        pow_spct[0] = 0;
        for (uint32_t i = 0; i < cen[1]; ++i) pow_spct[0] += mag[i] * sr_tab_b_tri_even[i] / 100u;
        for (int h = 2; h < SRO_TRI_NUM; h += 2) {
            pow_spct[h] = 0;
            for (uint32_t i = cen[h - 1]; i < cen[h + 1]; ++i) pow_spct[h] += mag[i] * sr_tab_b_tri_even[i] / 100u;
        }
        for (int h = 1; h < SRO_TRI_NUM - 2; h += 2) {
            pow_spct[h] = 0;
            for (uint32_t i = cen[h - 1]; i < cen[h + 1]; ++i) pow_spct[h] += mag[i] * sr_tab_b_tri_odd[i] / 100u;
        }
        pow_spct[SRO_TRI_NUM - 1] = 0;
        for (uint32_t i = cen[SRO_TRI_NUM - 2]; i < SRO_B_FRQ_MAX; ++i)
            pow_spct[SRO_TRI_NUM - 1] += mag[i] * sr_tab_b_tri_odd[i] / 100u;
        for (int h = 0; h < SRO_TRI_NUM; ++h) pow_spct[h] = sro_log100(pow_spct[h]);
        const int8_t *dct = sr_tab_dct;
        for (int c = 0; c < SRO_MFCC_NUM; ++c) {
            int16_t acc = 0;
            for (int i = 0; i < SRO_TRI_NUM; ++i)
                acc = (int16_t)(acc + ((int32_t)pow_spct[i]) * ((int32_t)dct[i]) / 100);
            mfcc_p[c] = acc;
            dct += SRO_TRI_NUM;
        }
        mfcc_p += SRO_MFCC_NUM;
        ++frm_con;
    }
    out->frm_num = frm_con;
}
void sro_mfcc_geom_b_batch(const uint16_t *pcm, uint32_t U, uint32_t B, const uint32_t *seg2, const sro_atap *atap, sro_ftr *ftr) {
    for (uint32_t b = 0; b < B; ++b) {
        const uint32_t st = seg2[2 * b], en = seg2[2 * b + 1];
        if (st == SRO_NULL || en == SRO_NULL || en > U || st > en || en - st < SRO_B_FRAME_LEN) { ftr[b].frm_num = 0; continue; }
        sro_mfcc_geom_b(pcm + (size_t)b * U, st, en, atap + b, ftr + b);
    }
}

/* ---- DTW.C:45-62 get_dis --------------------------------------------------------------------- */
uint32_t sro_get_dis(const int16_t *a, const int16_t *b) {
    uint32_t dis = 0;
    for (int i = 0; i < SRO_MFCC_NUM; ++i) {
        int32_t dif = a[i] - b[i];
        dis += (uint32_t)dif * (uint32_t)dif;
    }
    return (uint32_t)sqrtf((float)dis);                              /* DTW.C:59: u32->float->u32 */
}

/* ---- DTW.C:76-109 dtw_limit: 0 = ins, 1 = outs; X1,X2 recomputed from I,M (DTW.C:141-142) ---- */
int sro_dtw_limit(int x, int y, int I, int M) {
    int X1 = (uint16_t)((2 * M - I) / 3), X2 = (uint16_t)((4 * I - 2 * M) / 3);
    x &= 0xFFFF; y &= 0xFFFF;
    if (x < X1) { if (y >= 2 * x + 2) return 1; }
    else { if (2 * y + I - 2 * M >= x + 4) return 1; }
    if (x < X2) { if (2 * y + 2 <= x) return 1; }
    else { if (y + 4 <= 2 * x + M - 2 * I) return 1; }
    return 0;
}

/* ---- DTW.C:120-192 dtw: the greedy parallelogram walk; *cells counts get_dis evaluations ----- */
uint32_t sro_dtw(const sro_ftr *fin, const sro_ftr *fmdl, uint32_t *cells) {
    int I = fin->frm_num, M = fmdl->frm_num;
    uint32_t nc = 0;
    if (cells) *cells = 0;
    if (I > M * 2 || 2 * I < M) return SRO_DIS_ERR;                  /* DTW.C:133-137 */
    const int16_t *in = fin->mfcc_dat, *mdl = fmdl->mfcc_dat;
    uint32_t dis = sro_get_dis(in, mdl);                             /* DTW.C:146 */
    nc = 1;
    uint16_t x = 1, y = 1, step = 1;
    do {                                                             /* DTW.C:150-188 */
        uint32_t up = SRO_DIS_ERR, right = SRO_DIS_ERR, ru = SRO_DIS_ERR;
        if (!sro_dtw_limit(x, y + 1, I, M)) { up = sro_get_dis(mdl + 12, in); ++nc; }
        if (!sro_dtw_limit(x + 1, y, I, M)) { right = sro_get_dis(mdl, in + 12); ++nc; }
        if (!sro_dtw_limit(x + 1, y + 1, I, M)) { ru = sro_get_dis(mdl + 12, in + 12); ++nc; }
        uint32_t mn = ru;
        if (mn > right) mn = right;
        if (mn > up) mn = up;
        dis += mn;
        if (mn == ru) { in += 12; ++x; mdl += 12; ++y; }
        else if (mn == up) { mdl += 12; ++y; }
        else { in += 12; ++x; }
        ++step;
    } while (x < I && y < M);
    if (cells) *cells = nc;
    return dis / step;                                               /* DTW.C:191 */
}

/* ---- DTW.C:195-296 get_mean + get_mdl (dead code in the firmware): averaged template along the walk --------
 * Returns dis/step (SRO_DIS_ERR and mdl untouched if the guard rejects). The reference writes row `step-1` for
 * every visited point without bound; rows >= 119 are dropped here and frm_num clamped (the reference would
 * overrun mfcc_dat). */
uint32_t sro_get_mdl(const sro_ftr *f1, const sro_ftr *f2, sro_ftr *fm) {
    int I = f1->frm_num, M = f2->frm_num;
    if (I > M * 2 || 2 * I < M) return SRO_DIS_ERR;                  /* DTW.C:231-234 */
    const int16_t *in1 = f1->mfcc_dat, *in2 = f2->mfcc_dat;
    int row = 0;
    uint32_t dis = sro_get_dis(in1, in2);                            /* DTW.C:244 */
    for (int i = 0; i < 12; ++i) fm->mfcc_dat[i] = (int16_t)((in1[i] + in2[i]) / 2);   /* DTW.C:201,245 */
    uint16_t x = 1, y = 1, step = 1;
    do {                                                             /* DTW.C:249-291 */
        uint32_t up = !sro_dtw_limit(x, y + 1, I, M) ? sro_get_dis(in2 + 12, in1) : SRO_DIS_ERR;
        uint32_t right = !sro_dtw_limit(x + 1, y, I, M) ? sro_get_dis(in2, in1 + 12) : SRO_DIS_ERR;
        uint32_t ru = !sro_dtw_limit(x + 1, y + 1, I, M) ? sro_get_dis(in2 + 12, in1 + 12) : SRO_DIS_ERR;
        uint32_t mn = ru;
        if (mn > right) mn = right;
        if (mn > up) mn = up;
        dis += mn;
        if (mn == ru) { in1 += 12; ++x; in2 += 12; ++y; }
        else if (mn == up) { in2 += 12; ++y; }
        else { in1 += 12; ++x; }
        ++step;
        ++row;                                                       /* mdl += mfcc_num, DTW.C:286 */
        if (row < SRO_VV_FRM_MAX)
            for (int i = 0; i < 12; ++i) fm->mfcc_dat[row * 12 + i] = (int16_t)((in1[i] + in2[i]) / 2);
    } while (x < I && y < M);
    fm->frm_num = step > SRO_VV_FRM_MAX ? SRO_VV_FRM_MAX : step;     /* DTW.C:293 */
    return dis / step;
}

/* ---- dtw_band: NOT in the reference -- PARITY UNPINNED --------------
 * Classic DP the north_star's config[2] names: D(i,j) = d(i,j) + min(D(i-1,j), D(i,j-1), D(i-1,j-1))
 * over the Sakoe-Chiba band |j - round_down(i*M/I)| <= r (1-based i<=I, j<=M mapped 0-based below),
 * local distance = the reference's get_dis, u32 saturating accumulate, result D(I,M)/(I+M).
 * Same 2:1 length guard as dtw (DTW.C:133). Used only to check our own dtw_band kernel. */
uint32_t sro_dtw_band(const sro_ftr *fin, const sro_ftr *fmdl, int r, uint32_t *cells) {
    int I = fin->frm_num, M = fmdl->frm_num;
    if (cells) *cells = 0;
    if (I == 0 || M == 0 || I > M * 2 || 2 * I < M) return SRO_DIS_ERR;
    uint32_t prev[SRO_VV_FRM_MAX], curr[SRO_VV_FRM_MAX], nc = 0;
    for (int j = 0; j < M; ++j) prev[j] = SRO_DIS_ERR;
    for (int i = 0; i < I; ++i) {
        int c = (int)((int64_t)i * M / I);
        for (int j = 0; j < M; ++j) curr[j] = SRO_DIS_ERR;
        int lo = c - r < 0 ? 0 : c - r, hi = c + r >= M ? M - 1 : c + r;
        for (int j = lo; j <= hi; ++j) {
            uint32_t best;
            if (i == 0 && j == 0) best = 0;
            else {
                best = SRO_DIS_ERR;
                if (i > 0 && prev[j] < best) best = prev[j];
                if (j > 0 && curr[j - 1] < best) best = curr[j - 1];
                if (i > 0 && j > 0 && prev[j - 1] < best) best = prev[j - 1];
                if (best == SRO_DIS_ERR) continue;                   /* unreachable cell */
            }
            uint32_t d = sro_get_dis(fin->mfcc_dat + 12 * i, fmdl->mfcc_dat + 12 * j);
            ++nc;
            uint64_t s = (uint64_t)best + d;
            curr[j] = s >= SRO_DIS_ERR ? SRO_DIS_ERR - 1 : (uint32_t)s;
        }
        memcpy(prev, curr, sizeof(uint32_t) * (size_t)M);
    }
    if (cells) *cells = nc;
    if (prev[M - 1] == SRO_DIS_ERR) return SRO_DIS_ERR;
    return prev[M - 1] / (uint32_t)(I + M);
}

/* ---- main.c:249-296 spch_recg (buffer length, noise window and bank passed in) --------------- */
int sro_recognise(const uint16_t *pcm, uint32_t buf_len, uint32_t n_len, const uint8_t *bank, uint32_t n_slot,
                  uint32_t slot_stride, sro_atap *atap_out, uint32_t *seg_off6, sro_ftr *ftr_out,
                  uint32_t *score, uint32_t *best_idx, uint32_t *best_dis, uint32_t *cmd) {
    sro_atap atap; uint32_t seg[6]; sro_ftr ftr_local, *ftr = ftr_out ? ftr_out : &ftr_local;
    memset(&atap, 0, sizeof atap);
    sro_noise_atap(pcm, n_len, &atap);                               /* main.c:258 */
    sro_vad(pcm, buf_len, &atap, seg);                               /* main.c:260 */
    if (atap_out) *atap_out = atap;
    if (seg_off6) memcpy(seg_off6, seg, sizeof seg);
    *best_idx = 0; *cmd = 0; *best_dis = SRO_DIS_ERR;
    if (seg[1] == SRO_NULL) return 1;                                /* main.c:261-266 */
    sro_mfcc(pcm, seg[0], seg[1], &atap, ftr);                       /* main.c:268 */
    if (ftr->frm_num == 0) return 2;                                 /* main.c:269-274 */
    uint32_t min_dis = SRO_DIS_ERR, min_i = 0;                       /* main.c:276-291 */
    for (uint32_t i = 0; i < n_slot; ++i) {
        const sro_ftr *mdl = (const sro_ftr *)(bank + (size_t)i * slot_stride);
        uint32_t cur = (mdl->save_sign == SRO_SAVE_MASK) ? sro_dtw(ftr, mdl, NULL) : SRO_DIS_ERR;
        if (score) score[i] = cur;
        if (cur < min_dis) { min_dis = cur; min_i = i; }
    }
    *best_idx = min_i; *best_dis = min_dis; *cmd = min_i / SRO_FTR_PER_COMM;   /* main.c:292-294 */
    return 0;
}

/* ---- batch drivers (pthreads over contiguous shards) ----------------------------------------- */
typedef struct {
    int kind; uint32_t lo, hi;
    const uint16_t *pcm; uint32_t U, n_len; const uint8_t *bank; uint32_t n_slot, slot_stride;
    sro_atap *atap; const sro_atap *catap; uint32_t *seg_off; const uint32_t *seg2; sro_ftr *ftr; const sro_ftr *cin;
    uint32_t *score, *best_idx, *best_dis, *cmd; uint8_t *status; int check_sign, band_r; uint64_t cells;
} job_t;

static void *job_run(void *arg) {
    job_t *j = (job_t *)arg;
    for (uint32_t b = j->lo; b < j->hi; ++b) {
        if (j->kind == 0) {
            int st = sro_recognise(j->pcm + (size_t)b * j->U, j->U, j->n_len, j->bank, j->n_slot, j->slot_stride,
                                   j->atap ? j->atap + b : NULL, j->seg_off ? j->seg_off + 6 * (size_t)b : NULL,
                                   j->ftr ? j->ftr + b : NULL, j->score ? j->score + (size_t)b * j->n_slot : NULL,
                                   j->best_idx + b, j->best_dis + b, j->cmd + b);
            if (j->status) j->status[b] = (uint8_t)st;
        } else if (j->kind == 1) {
            sro_mfcc(j->pcm + (size_t)b * j->U, j->seg2[2 * b], j->seg2[2 * b + 1], j->catap + b, j->ftr + b);
        } else {
            for (uint32_t t = 0; t < j->n_slot; ++t) {
                const sro_ftr *mdl = (const sro_ftr *)(j->bank + (size_t)t * j->slot_stride);
                uint32_t c = 0, s = SRO_DIS_ERR;
                if (!j->check_sign || mdl->save_sign == SRO_SAVE_MASK)
                    s = j->band_r < 0 ? sro_dtw(j->cin + b, mdl, &c) : sro_dtw_band(j->cin + b, mdl, j->band_r, &c);
                j->score[(size_t)b * j->n_slot + t] = s;
                j->cells += c;
            }
        }
    }
    return NULL;
}

static uint64_t run_jobs(job_t *proto, uint32_t B, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > B) nthreads = B ? (int)B : 1;
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)nthreads);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int k = 0; k < nthreads; ++k) {
        jobs[k] = *proto;
        jobs[k].lo = (uint32_t)((uint64_t)B * k / nthreads);
        jobs[k].hi = (uint32_t)((uint64_t)B * (k + 1) / nthreads);
        jobs[k].cells = 0;
        if (nthreads > 1) pthread_create(&th[k], NULL, job_run, &jobs[k]);
        else job_run(&jobs[k]);
    }
    uint64_t cells = 0;
    for (int k = 0; k < nthreads; ++k) {
        if (nthreads > 1) pthread_join(th[k], NULL);
        cells += jobs[k].cells;
    }
    free(jobs); free(th);
    return cells;
}

void sro_recognise_batch(const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, const uint8_t *bank,
                         uint32_t n_slot, uint32_t slot_stride, sro_atap *atap, uint32_t *seg_off, sro_ftr *ftr,
                         uint32_t *score, uint32_t *best_idx, uint32_t *best_dis, uint32_t *cmd, uint8_t *status,
                         int nthreads) {
    job_t j; memset(&j, 0, sizeof j);
    j.kind = 0; j.pcm = pcm; j.U = U; j.n_len = n_len; j.bank = bank; j.n_slot = n_slot; j.slot_stride = slot_stride;
    j.atap = atap; j.seg_off = seg_off; j.ftr = ftr; j.score = score; j.best_idx = best_idx; j.best_dis = best_dis;
    j.cmd = cmd; j.status = status;
    run_jobs(&j, B, nthreads);
}

void sro_mfcc_batch(const uint16_t *pcm, uint32_t U, uint32_t B, const uint32_t *seg2, const sro_atap *atap,
                    sro_ftr *ftr, int nthreads) {
    job_t j; memset(&j, 0, sizeof j);
    j.kind = 1; j.pcm = pcm; j.U = U; j.seg2 = seg2; j.catap = atap; j.ftr = ftr;
    run_jobs(&j, B, nthreads);
}

void sro_dtw_batch(const sro_ftr *in, uint32_t B, const uint8_t *bank, uint32_t n_slot, uint32_t slot_stride,
                   int check_sign, int band_r, uint32_t *score, uint64_t *cells_total, int nthreads) {
    job_t j; memset(&j, 0, sizeof j);
    j.kind = 2; j.cin = in; j.bank = bank; j.n_slot = n_slot; j.slot_stride = slot_stride;
    j.check_sign = check_sign; j.band_r = band_r; j.score = score;
    uint64_t c = run_jobs(&j, B, nthreads);
    if (cells_total) *cells_total = c;
}
