/* spch_host.c -- the thin C host above the C-ABI: the two callers of the hot path in the reference,
 *   save_mdl   Src/APP/main.c:121-138   (train: noise_atap -> VAD -> get_mfcc -> save_ftr_mdl)
 *   spch_recg  Src/APP/main.c:249-296   (recognise: ... -> dtw over the bank -> argmin -> command)
 * written against the reference's OWN headers (include/compat/{VAD,MFCC,DTW,Flash}.H) and calling its OWN
 * function names, which libspeech_b200.so implements as batch-of-1 CUDA launches. The on-chip flash of
 * Flash.C is replaced by a RAM bank with the same 80 x 4 KB slot layout. A batched path
 * (sr_recognise_batch) is shown next to it: same results, one call for the whole batch. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "VAD.H"
#include "MFCC.H"
#include "DTW.H"
#include "Flash.H"
#include "sr_synth.h"

#define save_ok   0
#define VAD_fail  1
#define MFCC_fail 2

static u8 ftr_bank[comm_num * ftr_per_comm][size_per_ftr];   /* stands in for flash 0x8030000..0x8080000 */
static atap_tag  atap_arg;
static valid_tag valid_voice[max_vc_con];
static v_ftr_tag ftr;
typedef struct { u8 str[4]; } comm_tag;
static comm_tag commstr[comm_num] = {{"0"},{"1"},{"2"},{"3"},{"4"},{"5"},{"6"},{"7"},{"8"},{"9"},
                                     {"up"},{"dn"},{"fw"},{"bk"},{"lf"},{"rt"},{"bg"},{"sm"},{"x0"},{"x1"}};

/* Flash.C:17-67 save_ftr_mdl with the flash programming replaced by a memcpy into the RAM bank */
static u8 save_ftr_mdl(v_ftr_tag *f, u32 slot) {
    if (slot >= comm_num * ftr_per_comm) return Flash_Fail;
    memset(ftr_bank[slot], 0xFF, size_per_ftr);               /* erased flash */
    v_ftr_tag *dst = (v_ftr_tag *)ftr_bank[slot];
    dst->save_sign = save_mask;
    dst->frm_num = f->frm_num;
    memcpy(dst->mfcc_dat, f->mfcc_dat, 2u * mfcc_num * f->frm_num);
    return Flash_Success;
}

static u8 save_mdl(u16 *v_dat, u32 buf_len, u32 slot) {        /* main.c:121-138 */
    noise_atap(v_dat, atap_len, &atap_arg);
    VAD(v_dat, (u16)buf_len, valid_voice, &atap_arg);
    if (valid_voice[0].end == NULL) return VAD_fail;
    get_mfcc(&valid_voice[0], &ftr, &atap_arg);
    if (ftr.frm_num == 0) return MFCC_fail;
    return save_ftr_mdl(&ftr, slot);
}

static u8 *spch_recg(u16 *v_dat, u32 buf_len, u32 *mtch_dis) { /* main.c:249-296 */
    u16 i = 0, min_comm = 0;
    u32 min_dis = dis_max, cur_dis;
    noise_atap(v_dat, atap_len, &atap_arg);
    VAD(v_dat, (u16)buf_len, valid_voice, &atap_arg);
    if (valid_voice[0].end == NULL) { *mtch_dis = dis_err; return NULL; }
    get_mfcc(&valid_voice[0], &ftr, &atap_arg);
    if (ftr.frm_num == 0) { *mtch_dis = dis_err; return NULL; }
    for (u32 s = 0; s < comm_num * ftr_per_comm; ++s) {
        v_ftr_tag *mdl = (v_ftr_tag *)ftr_bank[s];
        cur_dis = (mdl->save_sign == save_mask) ? dtw(&ftr, mdl) : dis_err;
        if (cur_dis < min_dis) { min_dis = cur_dis; min_comm = i; }
        i++;
    }
    min_comm /= ftr_per_comm;
    *mtch_dis = min_dis;
    return commstr[min_comm].str;
}

int main(int argc, char **argv) {
    const u32 U = 8000, ncmd = 5, B = argc > 1 ? (u32)atoi(argv[1]) : 16;
    if (sr_device_count() == 0) { fprintf(stderr, "no CUDA device: libspeech_b200 has no CPU fallback\n"); return 2; }
    memset(ftr_bank, 0xFF, sizeof ftr_bank);
    u16 *tpl = malloc((size_t)ncmd * ftr_per_comm * U * 2), *pcm = malloc((size_t)B * U * 2);
    /* enrol: 4 redundant templates per command (Flash.C:4-5), synthetic "words" seeded per command */
    sr_synth_pcm_host(tpl, U, ncmd * ftr_per_comm, 0x7E3A0000ull, 1);
    for (u32 s = 0; s < ncmd * ftr_per_comm; ++s) {
        u8 rc = save_mdl(tpl + (size_t)s * U, U, s);
        if (rc != save_ok) { fprintf(stderr, "enrol slot %u failed (%u): %s\n", s, rc, sr_last_error(NULL)); return 1; }
    }
    /* recognise one by one through the reference's call sequence */
    sr_synth_pcm_host(pcm, U, B, 0x7E3A0000ull, 1);            /* same seeds: utterance b is template b */
    u32 *dis1 = malloc(B * 4), *cmd1 = malloc(B * 4);
    for (u32 b = 0; b < B; ++b) {
        u8 *lab = spch_recg(pcm + (size_t)b * U, U, &dis1[b]);
        cmd1[b] = lab ? (u32)((comm_tag *)lab - commstr) : 0xFFFFFFFFu;
    }
    /* the same batch in ONE call */
    sr_handle *h = NULL;
    if (sr_create(0, &h)) { fprintf(stderr, "%s\n", sr_last_error(NULL)); return 1; }
    sr_set_bank(h, ftr_bank, comm_num * ftr_per_comm, size_per_ftr);
    u32 *dis2 = malloc(B * 4), *cmd2 = malloc(B * 4);
    u8 *st = malloc(B);
    sr_recog_out out;
    memset(&out, 0, sizeof out);
    out.best_dis = dis2; out.cmd = cmd2; out.status = st;
    if (sr_recognise_batch(h, pcm, U, B, atap_len, &out)) { fprintf(stderr, "%s\n", sr_last_error(h)); return 1; }
    int bad = 0;
    for (u32 b = 0; b < B; ++b) {
        const u32 c2 = st[b] ? 0xFFFFFFFFu : cmd2[b];
        if (c2 != cmd1[b] || dis2[b] != dis1[b]) ++bad;
        if (b < 8) printf("utt %2u: spch_recg -> cmd %u dis %u | batch -> cmd %u dis %u status %u\n", b, cmd1[b], dis1[b], cmd2[b], dis2[b], st[b]);
    }
    printf("%u utterances, single-call path vs batched path: %d mismatches\n", B, bad);
    sr_destroy(h);
    return bad ? 1 : 0;
}
