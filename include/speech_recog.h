/* speech_recog.h -- C-ABI of libspeech_b200.so, the B200-native drop-in for the reference's
 * VAD -> MFCC -> DTW hot path (gk969/stm32-speech-recognition, Src/Speech_Recog + the FFT asm).
 *
 * Two layers, both plain C (pointers and sizes only, no CUDA/torch types in any signature):
 *
 *  (1) The reference's own entry points, same names / argument order / struct layouts, each a
 *      batch-of-1 launch of the CUDA kernels on a lazily created default handle (device 0):
 *          noise_atap   Src/Speech_Recog/VAD.H:24   (VAD.C:22-71)
 *          VAD          Src/Speech_Recog/VAD.H:25   (VAD.C:97-218)
 *          get_mfcc     Src/Speech_Recog/MFCC.H:27  (MFCC.C:86-191)
 *          dtw          Src/Speech_Recog/DTW.H:7    (DTW.C:120-192)
 *          fft          Src/Speech_Recog/MFCC.C:27  (global, no header)
 *          get_dis      Src/Speech_Recog/DTW.C:45   (global, no header)
 *          dtw_limit    Src/Speech_Recog/DTW.C:76   (global, no header; its file-static state is per thread here)
 *      A host program written against VAD.H / MFCC.H / DTW.H links unchanged (see include/compat/).
 *
 *  (2) Batched, re-entrant forms on an explicit handle (sr_*): B independent utterances per call,
 *      segments as sample OFFSETS instead of pointers (SR_SEG_NULL = NULL), template bank in the
 *      reference's flash-slot layout (Src/BSP/Flash.H:11-20). Host-buffer variants copy
 *      host->device, launch, copy back and synchronise; *_dev variants take device pointers and
 *      are asynchronous on the handle's stream (sr_set_stream / sr_sync).
 *
 * Results are bit-identical to the reference C compiled for a host CPU (VAD boundaries, MFCC
 * s16, DTW u32 distances, best-template index). There is no CPU fallback: without a CUDA
 * device every compute entry point fails (sr_* return non-zero; the reference-named functions
 * write their failure sentinels and set sr_last_error).
 */
#ifndef SPEECH_RECOG_H_
#define SPEECH_RECOG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants visible through the ABI (same values as the reference's macros) ------------- */
#define SR_FS            8000u        /* ADC.H:7   fs                                   */
#define SR_VCBUF_LEN     16000u       /* ADC.H:9   VcBuf_Len (2 s)                      */
#define SR_ATAP_LEN      2400u        /* ADC.H:11  atap_len (300 ms noise window)       */
#define SR_MAX_VC_CON    3u           /* VAD.H:4   max_vc_con                           */
#define SR_FRAME_LEN     160u         /* VAD.H:7   frame_len (20 ms)                    */
#define SR_FRAME_MOV     80u          /* VAD.H:8   frame_mov (10 ms hop)                */
#define SR_FFT_POINT     1024u        /* MFCC.H:8  fft_point                            */
#define SR_FRQ_MAX       512u         /* MFCC.H:9  frq_max                              */
#define SR_TRI_NUM       24u          /* MFCC.H:12 tri_num                              */
#define SR_MFCC_NUM      12u          /* MFCC.H:13 mfcc_num                             */
#define SR_VV_FRM_MAX    119u         /* MFCC.H:15-16 vv_frm_max                        */
#define SR_DIS_ERR       0xFFFFFFFFu  /* DTW.H:4   dis_err                              */
#define SR_DIS_MAX       0xFFFFFFFFu  /* DTW.H:5   dis_max                              */
#define SR_SAVE_MASK     12345u       /* Flash.H:11 save_mask                           */
#define SR_SIZE_PER_FTR  4096u        /* Flash.H:13 size_per_ftr (flash slot)           */
#define SR_FTR_PER_COMM  4u           /* Flash.H:15 ftr_per_comm                        */
#define SR_COMM_NUM      20u          /* Flash.H:17 comm_num                            */
#define SR_SEG_NULL      0xFFFFFFFFu  /* offset encoding of a NULL valid_tag pointer    */

/* per-utterance status of sr_recognise_* (main.c:38-41: save_ok / VAD_fail / MFCC_fail)       */
#define SR_ST_OK         0u
#define SR_ST_VAD_FAIL   1u
#define SR_ST_MFCC_FAIL  2u

/* ---- the reference's types (VAD.H:10-22, MFCC.H:18-25), identical layout -------------------- */
#ifndef SR_NO_REFERENCE_TYPES
typedef struct {
    uint32_t mid_val;   /* DC level of the capture ("signed zero")           */
    uint16_t n_thl;     /* noise band half-width for the band-crossing rate  */
    uint16_t z_thl;     /* band-crossing-rate threshold                      */
    uint32_t s_thl;     /* short-time magnitude threshold                    */
} atap_tag;

typedef struct {
    uint16_t *start;    /* first sample of the segment (into the caller's PCM buffer) */
    uint16_t *end;      /* one past the last sample; NULL = segment never closed      */
} valid_tag;

#pragma pack(push, 1)
typedef struct {
    uint16_t save_sign;                                   /* SR_SAVE_MASK marks a valid flash template */
    uint16_t frm_num;                                     /* number of MFCC frames                     */
    int16_t  mfcc_dat[SR_VV_FRM_MAX * SR_MFCC_NUM];       /* row-major frame x coefficient             */
} v_ftr_tag;                                              /* 2860 bytes                                */
#pragma pack(pop)
#endif

/* ---- (1) reference-named entry points ------------------------------------------------------- */
void      noise_atap(const uint16_t *noise, uint16_t n_len, atap_tag *atap);
void      VAD(const uint16_t *vc, uint16_t buf_len, valid_tag *valid_voice, atap_tag *atap_arg);
void      get_mfcc(valid_tag *valid, v_ftr_tag *v_ftr, atap_tag *atap_arg);
uint32_t  dtw(v_ftr_tag *ftr_in, v_ftr_tag *frt_mdl);
uint32_t *fft(int16_t *dat_buf, uint16_t buf_len);        /* returns a thread-local u32[1024]; [0,512) valid */
uint32_t  get_dis(int16_t *frm_ftr1, int16_t *frm_ftr2);
uint8_t   dtw_limit(uint16_t x, uint16_t y);            /* DTW.C:76: 0 ins / 1 outs, for the (I,M) of this thread's last dtw() */

/* ---- (2) batched handle API -----------------------------------------------------------------
 * A handle owns its device workspaces and stream; use one handle per thread (calls on the same handle must not
 * overlap). Different handles -- on the same or on different GPUs -- are independent. */
typedef struct sr_handle sr_handle;

int         sr_create(int device, sr_handle **out);       /* device ordinal; <0 = current device     */
int         sr_destroy(sr_handle *h);
/* A handle's device scratch (segments, features, scores, the kernels' utterance hand-out counters) serves ONE stream at a
 * time: sr_sync (or order the new stream behind the old one) before switching streams; concurrent streams take one handle
 * each -- handles are cheap, the tables are per device. */
int         sr_set_stream(sr_handle *h, void *cuda_stream /* cudaStream_t used verbatim; NULL = legacy default stream */);
int         sr_use_own_stream(sr_handle *h);              /* back to the handle's private non-blocking stream (the default) */
int         sr_sync(sr_handle *h);
const char *sr_last_error(const sr_handle *h /* NULL: last error of the calling thread */);
int         sr_device_count(void);                        /* 0 when no CUDA device is usable         */
int         sr_abi_version(void);
void       *sr_host_alloc(size_t bytes);                  /* pinned host memory for fast H2D/D2H     */
void        sr_host_free(void *p);                        /* for sr_host_alloc and sr_host_alloc_dev  */
/* NUMA placement of the host side. The end-to-end call is bound by the H2D copy of the caller's PCM (the u16 v_dat
 * buffer of main.c:249), so on a two-socket box the pinned pages and the threads that feed a GPU belong on the
 * socket that GPU hangs off. sr_host_alloc_dev returns pinned memory whose pages live on `device`'s NUMA node
 * (plain sr_host_alloc on single-node boxes); sr_bind_thread_to_device restricts the CALLING thread (and every
 * thread it creates afterwards, e.g. the library's packer pool) to that node's CPUs and returns the node, or -1 when
 * nothing was changed (one node, unknown topology); sr_device_numa_node / sr_host_numa_node report placement. */
void       *sr_host_alloc_dev(int device, size_t bytes);
int         sr_bind_thread_to_device(int device);
int         sr_device_numa_node(int device);              /* -1 = unknown                            */
int         sr_host_numa_node(const void *p);             /* node backing the page at p; -1 = unknown */

/* Frame geometry of get_mfcc for this handle (sr_mfcc_batch*, sr_recognise_batch*, sr_enrol_batch, streaming):
 *   SR_GEOM_REF  160-sample frames / 80 hop / 1024-point FFT -- the reference's (VAD.H:5-8, MFCC.H:8); bit-exact parity
 *   SR_GEOM_B    200-sample frames / 80 hop /  256-point FFT -- BASELINE configs[0]'s "256-pt/25 ms/10 ms"; an EXTENSION:
 *                the reference's algorithm and Matlab table formulas at the other two sizes, checked only against this
 *                repo's own CPU restatement (parity unpinned by the reference). noise_atap / VAD keep the reference's
 *                framing in both; features of the two geometries must not be mixed in one bank. */
#define SR_GEOM_REF 0
#define SR_GEOM_B   1
int sr_set_geometry(sr_handle *h, int geom);
int sr_get_geometry(const sr_handle *h);

/* Template bank: n_slot slots of slot_stride bytes (>= sizeof(v_ftr_tag), multiple of 4), each
 * starting with a v_ftr_tag; the flash layout of Flash.H:11-20 is slot_stride = 4096.
 * sr_set_bank copies host->device; sr_set_bank_dev borrows a device pointer. */
int sr_set_bank(sr_handle *h, const void *bank, uint32_t n_slot, uint32_t slot_stride);
int sr_set_bank_dev(sr_handle *h, const void *bank_dev, uint32_t n_slot, uint32_t slot_stride);

/* pcm: B utterances, utterance b at pcm + b*U (u16 samples, 12-bit ADC codes).
 * noise_atap over the first n_len samples of every utterance (VAD.C:22-71); atap[b] is left
 * untouched when n_len % 240 != 0 exactly like the reference (VAD.C:33-36). */
int sr_noise_atap_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len,
                        atap_tag *atap /* [B] in/out */);
/* VAD over the first buf_len (<= U) samples; seg_off[b][k][0/1] = start/end offset of segment k */
int sr_vad_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t buf_len,
                 const atap_tag *atap /* [B] */, uint32_t *seg_off /* [B][3][2] */);
/* get_mfcc of one segment per utterance: seg[b*seg_stride + 0/1] = start/end sample offsets.
 * Only frm_num and the first frm_num rows of ftr[b] are written (MFCC.C never writes save_sign). */
int sr_mfcc_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, const uint32_t *seg,
                  uint32_t seg_stride, const atap_tag *atap /* [B] */, v_ftr_tag *ftr /* [B] */);
/* dtw of every input against every bank slot. flags bit0: honour save_sign like spch_recg
 * (main.c:283: slots whose save_sign != 12345 score SR_DIS_ERR). score may be NULL.
 * best_idx/best_dis follow main.c:276-291 (strict '<', first wins, start 0 / 0xFFFFFFFF). */
#define SR_DTW_CHECK_SIGN 1u
#define SR_DTW_BAND       2u          /* use the Sakoe-Chiba banded DP (extension, not in the reference) */
int sr_dtw_batch(sr_handle *h, const v_ftr_tag *in, uint32_t B, uint32_t flags, int band_r,
                 uint32_t *score /* [B][n_slot] or NULL */, uint32_t *best_idx /* [B] or NULL */,
                 uint32_t *best_dis /* [B] or NULL */);
/* spch_recg (main.c:249-296) for B utterances: noise_atap(first n_len) -> VAD(U) -> get_mfcc(seg 0)
 * -> dtw against the bank -> argmin -> cmd = idx / SR_FTR_PER_COMM. Any output pointer may be NULL. */
typedef struct {
    atap_tag  *atap;       /* [B]            */
    uint32_t  *seg_off;    /* [B][3][2]      */
    v_ftr_tag *ftr;        /* [B]            */
    uint32_t  *score;      /* [B][n_slot]    */
    uint32_t  *best_idx;   /* [B]            */
    uint32_t  *best_dis;   /* [B] (mtch_dis) */
    uint32_t  *cmd;        /* [B]            */
    uint8_t   *status;     /* [B] SR_ST_*    */
} sr_recog_out;
int sr_recognise_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len,
                       const sr_recog_out *out);

/* Command labels: commstr[] of main.c:25-31, the u8* spch_recg returns (main.c:295). n_labels records of label_stride
 * bytes (the reference: comm_tag {u8 str[3]}), copied. Without a table the reference's own 18 labels are used
 * ("0 ".."9 ", then the GBK codes of up/down/front/back/left/right/big/small). sr_label returns NULL for a command
 * index without a label; sr_labels_batch maps the cmd/status arrays of sr_recognise_batch, NULL where spch_recg
 * returns NULL (VAD or MFCC failed, main.c:261-274). */
int            sr_set_labels(sr_handle *h, const void *labels, uint32_t n_labels, uint32_t label_stride);
const uint8_t *sr_label(const sr_handle *h /* may be NULL: reference table */, uint32_t cmd);
int            sr_labels_batch(const sr_handle *h, const uint32_t *cmd, const uint8_t *status, uint32_t B, const uint8_t **labels_out);

/* The same call spread over several GPUs of one box: contiguous shards, one host thread per handle, results
 * written straight into the caller's host arrays (no collective needed for host outputs). handles[g] must be
 * handles on different devices with the same bank set. */
int sr_recognise_batch_multi(sr_handle *const *handles, uint32_t n_handles, const uint16_t *pcm, uint32_t U, uint32_t B,
                             uint32_t n_len, const sr_recog_out *out);

/* ---- the one exchange step of the multi-GPU form: NCCL all-gather behind the C-ABI ----------------------
 * Utterances are sharded over ranks -- one handle per GPU, one process or one host thread per rank -- with no
 * data-path communication; at the end the per-template scores (and the 8-byte argmin keys) of all shards are
 * all-gathered. NCCL is bound at run time (dlopen of libnccl.so.2, SR_NCCL_LIB overrides), so single-GPU users need
 * no NCCL at all. sr_comm_unique_id is called on ONE rank and its 128 bytes handed to the others by the host's own
 * means (shared memory between threads, a file, MPI, torch.distributed ...); sr_comm_create is collective.
 * The collective runs on a stream of its own, ordered after the kernels that produced its input, and overlaps whatever
 * the handle's stream does next; sr_comm_wait makes the handle's stream (and thus sr_sync) wait for it. Errors:
 * 10000 + ncclResult_t, or -2 when NCCL cannot be loaded. */
#define SR_COMM_ID_BYTES 128
int sr_comm_unique_id(void *id128);
int sr_comm_create(sr_handle *h, int rank, int world, const void *id128);
int sr_comm_destroy(sr_handle *h);
int sr_comm_rank(const sr_handle *h);
int sr_comm_world(const sr_handle *h);
int sr_comm_nccl_version(void);                           /* 0 when NCCL cannot be loaded */
int sr_comm_wait(sr_handle *h);
int sr_allgather_dev(sr_handle *h, const void *send_dev, void *recv_dev, size_t bytes_per_rank);

/* device-pointer variants: every pointer is device memory on the handle's device, calls are
 * asynchronous on the handle's stream. Alignment: pcm 2 bytes, everything else 4 bytes. */
int sr_noise_atap_batch_dev(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, atap_tag *atap);
int sr_vad_batch_dev(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t buf_len,
                     const atap_tag *atap, uint32_t *seg_off);
int sr_mfcc_batch_dev(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, const uint32_t *seg,
                      uint32_t seg_stride, const atap_tag *atap, v_ftr_tag *ftr);
int sr_dtw_batch_dev(sr_handle *h, const v_ftr_tag *in, uint32_t B, uint32_t flags, int band_r,
                     uint32_t *score, uint32_t *best_idx, uint32_t *best_dis);
int sr_recognise_batch_dev(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len,
                           const sr_recog_out *out_dev);
/* sr_recognise_batch_dev on this rank's shard + the exchange step: gathered_score[world*B][n_slot] (rank-major, i.e.
 * global utterance order for equal contiguous shards; needs out_dev->score) and/or gathered_best[world*B] =
 * best_dis << 32 | best_idx, the key of the strict-'<' first-wins argmin (main.c:285-289). Either may be NULL. */
int sr_recognise_batch_dev_allgather(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len,
                                     const sr_recog_out *out_dev, uint32_t *gathered_score, uint64_t *gathered_best);

/* save_mdl (main.c:121-138) for B utterances: noise_atap -> VAD -> get_mfcc(segment 0) -> save_ftr_mdl
 * (Flash.C:17-67) into slot b of a flash-layout bank image bank_out[B][slot_stride] (host memory).
 * status[b] (may be NULL): 0 save_ok / 1 VAD_fail / 2 MFCC_fail; failed slots stay erased (0xFF). */
int sr_enrol_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, void *bank_out,
                   uint32_t slot_stride, uint8_t *status);
/* get_mdl + get_mean (DTW.C:195-296, defined but never called by the firmware): mdl[p] = element-wise mean of
 * in1[p] and in2[p] along their greedy DTW path, dis[p] = step-normalised path distance; rejected pairs
 * (2:1 length guard) return dis_err and leave mdl[p] untouched. Paths longer than 119 points are truncated
 * (the reference writes out of bounds there). */
int sr_get_mdl_batch(sr_handle *h, const v_ftr_tag *in1, const v_ftr_tag *in2, uint32_t n, v_ftr_tag *mdl, uint32_t *dis);

/* ---- streaming front end (stands in for record(), main.c:77-102 / ADC.C:11-103) ----------------------------
 * n_streams concurrent captures of max_samples samples each, fed in chunks -- in lock step (sr_streams_push) or every
 * stream at its own pace (sr_streams_push_ragged). Each push advances noise_atap (once the first n_len samples of a
 * stream are in) and VAD with the reference's carried state, and recognises every segment that closes (get_mfcc + dtw
 * + argmin against the handle's bank): one H2D copy, five kernels, one D2H copy and ONE synchronisation per push.
 * After the last chunk the events equal the batch results on the complete buffers; the reference itself only ever
 * recognises segment 0 (main.c:268), here all <= 3 segments of a stream produce an event.
 * Events are never dropped: what does not fit max_events stays queued (oldest first) and is handed out by the next
 * push or by sr_streams_fetch; 3 * n_streams is always enough for one push. */
typedef struct sr_stream_pool sr_stream_pool;
typedef struct {
    uint32_t stream, segment;   /* which stream, which of its <= 3 segments       */
    uint32_t start, end;        /* sample offsets of the segment                   */
    uint32_t status;            /* SR_ST_OK or SR_ST_MFCC_FAIL                     */
    uint32_t frm_num, best_idx, best_dis, cmd;
} sr_stream_event;
int sr_streams_create(sr_handle *h, uint32_t n_streams, uint32_t max_samples, uint32_t n_len, sr_stream_pool **out);
int sr_streams_destroy(sr_stream_pool *p);
int sr_streams_reset(sr_stream_pool *p);
int sr_streams_push(sr_stream_pool *p, const uint16_t *chunk /* host [n_streams][chunk_stride] */, uint32_t chunk_len,
                    uint32_t chunk_stride, sr_stream_event *events, uint32_t max_events, uint32_t *n_events);
int sr_streams_push_ragged(sr_stream_pool *p, const uint16_t *chunk /* host [n_streams][chunk_stride] */, uint32_t chunk_stride,
                           const uint32_t *lens /* [n_streams] samples for each stream, 0 = none */,
                           sr_stream_event *events, uint32_t max_events, uint32_t *n_events);
int sr_streams_fetch(sr_stream_pool *p, sr_stream_event *events, uint32_t max_events, uint32_t *n_events);
uint32_t sr_streams_pending(const sr_stream_pool *p);
int sr_streams_segments(sr_stream_pool *p, uint32_t *seg_off /* [n_streams][3][2] or NULL */, atap_tag *atap /* or NULL */);

/* The same over several GPUs of one box (BASELINE configs[4]): streams [S*g/G, S*(g+1)/G) live on handles[g]; one
 * persistent host thread per shard (bound to its GPU's NUMA node) runs that shard's push, so the G pushes overlap.
 * chunk / lens / seg_off / atap are indexed by GLOBAL stream number, events carry global stream numbers. */
typedef struct sr_stream_group sr_stream_group;
int sr_stream_group_create(sr_handle *const *handles, uint32_t n_handles, uint32_t n_streams, uint32_t max_samples,
                           uint32_t n_len, sr_stream_group **out);
int sr_stream_group_destroy(sr_stream_group *g);
int sr_stream_group_reset(sr_stream_group *g);
int sr_stream_group_push(sr_stream_group *g, const uint16_t *chunk, uint32_t chunk_len, uint32_t chunk_stride,
                         sr_stream_event *events, uint32_t max_events, uint32_t *n_events);
int sr_stream_group_push_ragged(sr_stream_group *g, const uint16_t *chunk, uint32_t chunk_stride, const uint32_t *lens,
                                sr_stream_event *events, uint32_t max_events, uint32_t *n_events);
int sr_stream_group_segments(sr_stream_group *g, uint32_t *seg_off, atap_tag *atap);

/* secondary globals of the reference, batched: fft magnitudes (MFCC.C:27-62) of n frames of
 * `len` (<=1024) s16 samples each -> u32[n][512]; get_dis (DTW.C:45-62) of n row pairs. */
int sr_fft_mag_batch(sr_handle *h, const int16_t *frames, uint32_t len, uint32_t n, uint32_t *mag);
int sr_get_dis_batch(sr_handle *h, const int16_t *a, const int16_t *b, uint32_t n, uint32_t *dis);
/* dtw_limit (DTW.C:76-109) for n points with explicit frame counts: out[i] = 0 ins / 1 outs */
int sr_dtw_limit_batch(sr_handle *h, const uint16_t *x, const uint16_t *y, const uint16_t *I, const uint16_t *M, uint32_t n,
                       uint8_t *out);
/* raw cr4_fft_1024_stm32 (Src/BSP/cr4_fft_1024_stm32.s:219-281) of n packed inputs (re | im<<16,
 * u32[n][1024]) -> packed outputs; exists so tests can pin the FFT kernel code against the asm restatement
 * on arbitrary complex data */
int sr_fft_raw_batch(sr_handle *h, const uint32_t *in_packed, uint32_t n, uint32_t *out_packed);

/* test hook: count of float bit patterns in [lo_bits, hi_bits) where the kernels' branch-free sqrt differs
 * from the IEEE sqrt.rn.f32 (0 over [1.0f, 2^33), the range the path can produce) */
int sr_debug_sqrt_mismatches(sr_handle *h, uint32_t lo_bits, uint32_t hi_bits, uint64_t *mismatches);

/* Packed PCM transport of sr_recognise_batch (host buffers): chunks whose samples are all < 4096 (the reference's
 * 12-bit ADC range) may cross PCIe as 12 bits per sample, packed by host worker threads and expanded on the
 * device; chunks with any larger sample travel as plain u16, so results never change. The caller's thread keeps
 * sending plain chunks from the front of the batch while the workers pack from the back, so the call is never slower
 * than the plain transport and approaches 3/4 of its PCIe time as the CPU share grows. mode: 0 off, 1 on,
 * -1 automatic = the default: considered when this rank's share of the usable CPUs (affinity capped by the cgroup quota,
 * divided by LOCAL_WORLD_SIZE) is >= 6, no other local rank's GPU hangs off the same NUMA node and the batch has >= 4
 * chunks; the library then MEASURES: one call plain, one packed, afterwards whichever is clearly faster, the other re-probed every 32nd call
 * (packing gains ~16 % with one GPU per socket and loses with four). SR_PACK12=0|1 overrides the automatic choice,
 * SR_PACK_THREADS the worker count (default: CPU share - 3, at most 10). */
int sr_set_transport(sr_handle *h, int mode);
/* transport statistics of the last sr_recognise_batch call: chunks sent packed / plain, bytes copied host -> device */
int sr_transport_stats(const sr_handle *h, uint32_t *packed_chunks, uint32_t *plain_chunks, uint64_t *h2d_bytes);
/* test hooks: the host packer alone (variant 0 scalar, 1 AVX2, 2 AVX-512 VBMI, 3 AVX-512 VBMI with non-temporal stores, -1 best available, 100+N the N-thread worker pool; returns the OR
 * of all samples or 0xFFFFFFFF if the variant is unavailable; no GPU needed) and the device expander alone */
uint32_t sr_debug_pack12_host(int variant, const uint16_t *src, uint64_t n, uint8_t *dst);
int sr_debug_unpack12(sr_handle *h, const uint8_t *packed, uint64_t n, uint16_t *out);

/* Per-kernel device timing: after sr_timing_enable(h, max_records) every kernel launch of this handle is
 * bracketed by a CUDA event pair on the launching stream; sr_timing_collect synchronises the stream and
 * returns (tag, milliseconds) per launch in issue order, then rearms. Tags: 0 noise_atap+VAD, 1 get_mfcc,
 * 2 status, 3 best-init, 4 dtw (greedy), 5 best-final, 6 dtw (banded DP). max_records = 0 disables. */
int sr_timing_enable(sr_handle *h, uint32_t max_records);
int sr_timing_collect(sr_handle *h, uint32_t *tags, float *ms, uint32_t cap, uint32_t *n);

/* Kernel variant of the greedy dtw (both bit-identical): 0 = one lane per (utterance, template) pair for the whole walk,
 * 1 = pairs handed to lanes dynamically from a ring of staged utterances (no lane waits for the longest walk of its
 * warp), -1 = the library default (SR_DTW_VARIANT=0|1 overrides it). */
int sr_set_dtw_variant(sr_handle *h, int variant);

/* number of kernel launches this handle has issued (bench.py reports it as gpu_launches) */
uint64_t sr_launch_count(const sr_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* SPEECH_RECOG_H_ */
