/* spch_host_mgpu.c -- the multi-GPU form of the thin C host (north_star: "a thin C host over a C-ABI ... the batch of
 * independent utterances shards trivially across the GPUs of one box with a single NCCL all-gather of per-template
 * scores at the end"). Plain C + pthreads, no Python, no torch: one host thread per GPU, each with its own handle;
 * the template bank (the flash image of Flash.H:11-20) is replicated; utterances are cut into contiguous shards;
 * every rank runs spch_recg (main.c:249-296) on its shard and the scores + argmin keys of all shards are all-gathered
 * through sr_recognise_batch_dev_allgather (NCCL bound inside libspeech_b200.so at run time).
 * Check: every rank's gathered result equals the single-GPU result of the whole batch, bit for bit.
 *   usage: spch_host_mgpu [n_gpus (default: all, 2 ranks on one GPU are refused by NCCL)] [utterances per rank] */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "speech_recog.h"
#include "sr_synth.h"

#define CK(x) do { int rc__ = (x); if (rc__) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc__, sr_last_error(NULL)); exit(1); } } while (0)
#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__)); exit(1); } } while (0)

enum { U = 8000, T = 12, N_LEN = 2400 };

typedef struct {
    int rank, world;
    uint32_t B;                      /* utterances per rank */
    const uint16_t *pcm_all;         /* host, [world*B][U] */
    const unsigned char *bank;       /* host flash image, T slots */
    unsigned char id[SR_COMM_ID_BYTES];
    uint32_t *g_score;               /* host copies of what this rank gathered */
    uint64_t *g_best;
} rank_arg;

static void *rank_main(void *p) {
    rank_arg *a = (rank_arg *)p;
    const uint32_t B = a->B, W = (uint32_t)a->world;
    sr_handle *h = NULL;
    CK(sr_create(a->rank, &h));
    sr_bind_thread_to_device(a->rank);                       /* feed the GPU from its own socket */
    CK(sr_set_bank(h, a->bank, T, SR_SIZE_PER_FTR));
    CK(sr_comm_create(h, a->rank, a->world, a->id));
    CU(cudaSetDevice(a->rank));
    uint16_t *d_pcm; uint32_t *d_score, *d_gs; uint64_t *d_gb;
    CU(cudaMalloc((void **)&d_pcm, (size_t)B * U * 2));
    CU(cudaMalloc((void **)&d_score, (size_t)B * T * 4));
    CU(cudaMalloc((void **)&d_gs, (size_t)W * B * T * 4));
    CU(cudaMalloc((void **)&d_gb, (size_t)W * B * 8));
    CU(cudaMemcpy(d_pcm, a->pcm_all + (size_t)a->rank * B * U, (size_t)B * U * 2, cudaMemcpyHostToDevice));
    sr_recog_out o;
    memset(&o, 0, sizeof o);
    o.score = d_score;
    for (int rep = 0; rep < 2; ++rep)                        /* twice: the second call orders itself after the first gather */
        CK(sr_recognise_batch_dev_allgather(h, d_pcm, U, B, N_LEN, &o, d_gs, d_gb));
    CK(sr_sync(h));                                          /* covers the collective (sr_comm_wait inside) */
    CU(cudaMemcpy(a->g_score, d_gs, (size_t)W * B * T * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(a->g_best, d_gb, (size_t)W * B * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_pcm); cudaFree(d_score); cudaFree(d_gs); cudaFree(d_gb);
    CK(sr_comm_destroy(h));
    sr_destroy(h);
    return NULL;
}

int main(int argc, char **argv) {
    int ndev = sr_device_count();
    if (ndev == 0) { fprintf(stderr, "no CUDA device: libspeech_b200 has no CPU fallback\n"); return 2; }
    int world = argc > 1 ? atoi(argv[1]) : ndev;
    if (world < 1 || world > ndev) { fprintf(stderr, "need 1..%d ranks (one GPU each)\n", ndev); return 2; }
    const uint32_t B = argc > 2 ? (uint32_t)atoi(argv[2]) : 512, N = B * (uint32_t)world;
    if (sr_comm_nccl_version() == 0) { fprintf(stderr, "NCCL cannot be loaded: %s\n", sr_last_error(NULL)); return 3; }

    /* enrolment on GPU 0: T synthetic words -> flash-layout bank (save_mdl, main.c:121-138) */
    uint16_t *tpl = malloc((size_t)T * U * 2), *pcm = malloc((size_t)N * U * 2);
    unsigned char *bank = malloc((size_t)T * SR_SIZE_PER_FTR);
    sr_synth_pcm_host(tpl, U, T, 0x7E3A0000ull, 1);
    sr_synth_pcm_host(pcm, U, N, 0x5EED0000ull, 1);
    sr_handle *h0 = NULL;
    CK(sr_create(0, &h0));
    uint8_t est[T];
    CK(sr_enrol_batch(h0, tpl, U, T, N_LEN, bank, SR_SIZE_PER_FTR, est));
    CK(sr_set_bank(h0, bank, T, SR_SIZE_PER_FTR));

    /* single-GPU result of the WHOLE batch: the thing every rank must end up with */
    uint32_t *score1 = malloc((size_t)N * T * 4), *idx1 = malloc((size_t)N * 4), *dis1 = malloc((size_t)N * 4), *cmd1 = malloc((size_t)N * 4);
    uint8_t *st1 = malloc(N);
    sr_recog_out o1;
    memset(&o1, 0, sizeof o1);
    o1.score = score1; o1.best_idx = idx1; o1.best_dis = dis1; o1.cmd = cmd1; o1.status = st1;
    CK(sr_recognise_batch(h0, pcm, U, N, N_LEN, &o1));
    sr_destroy(h0);

    unsigned char id[SR_COMM_ID_BYTES];
    CK(sr_comm_unique_id(id));                                /* one "rank" creates it; threads share it through memory */
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)world);
    rank_arg *args = calloc((size_t)world, sizeof(rank_arg));
    for (int r = 0; r < world; ++r) {
        args[r].rank = r; args[r].world = world; args[r].B = B; args[r].pcm_all = pcm; args[r].bank = bank;
        memcpy(args[r].id, id, sizeof id);
        args[r].g_score = malloc((size_t)N * T * 4);
        args[r].g_best = malloc((size_t)N * 8);
        pthread_create(&th[r], NULL, rank_main, &args[r]);
    }
    for (int r = 0; r < world; ++r) pthread_join(th[r], NULL);

    long bad = 0;
    for (int r = 0; r < world; ++r)
        for (uint32_t u = 0; u < N; ++u) {
            /* failed utterances never reach dtw (main.c:261-274): their score rows are undefined, their key is (dis_err, 0) */
            if (st1[u] == SR_ST_OK && memcmp(args[r].g_score + (size_t)u * T, score1 + (size_t)u * T, T * 4) != 0) ++bad;
            const uint64_t want = ((uint64_t)dis1[u] << 32) | idx1[u];
            if (args[r].g_best[u] != want) ++bad;
        }
    uint32_t ok = 0;
    for (uint32_t u = 0; u < N; ++u) ok += st1[u] == SR_ST_OK;
    const uint8_t *lab = sr_label(NULL, cmd1[0]);
    printf("NCCL %d, %d ranks x %u utterances x %d templates: %u recognised, utterance 0 -> command %u label bytes %02x %02x; "
           "gathered scores + argmin keys on every rank vs the single-GPU batch: %ld mismatches\n",
           sr_comm_nccl_version(), world, B, T, ok, cmd1[0], lab ? lab[0] : 0, lab ? lab[1] : 0, bad);
    return bad ? 1 : 0;
}
