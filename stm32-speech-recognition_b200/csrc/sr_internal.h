// sr_internal.h -- private to the library: the handle, workspaces and launch prototypes shared by sr_api.cu
// and sr_stream.cu. Nothing here is part of the C-ABI.
#pragma once
#include <mutex>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>
#include "sr_common.cuh"

namespace srk {
cudaError_t launch_vad(const u16 *pcm, u32 U, u32 B, u32 n_len, u32 buf_len, int do_atap, int do_vad, atap_tag *atap,
                       u32 *seg_off, int num_sms, cudaStream_t st, u32 *work = nullptr);
cudaError_t launch_mfcc(const u16 *pcm, u32 U, u32 B, const u32 *seg, u32 seg_stride, const atap_tag *atap, void *ftr,
                        int num_sms, cudaStream_t st, const u32 *row_map = nullptr, u32 rows_total = 0,
                        const u32 *B_dev = nullptr, u32 *work = nullptr);
cudaError_t launch_mfcc_geomb(const u16 *pcm, u32 U, u32 B, const u32 *seg, u32 seg_stride, const atap_tag *atap, void *ftr,
                              int num_sms, cudaStream_t st, const u32 *row_map = nullptr, const u32 *B_dev = nullptr);
cudaError_t launch_fft_generic(const u32 *in_packed, const s16 *frames, u32 len, u32 n, u32 *raw_out, u32 *mag,
                               cudaStream_t st);
cudaError_t launch_dtw(const void *in_ftr, u32 B, const void *bank, u32 T, u32 slot_stride, u32 flags, u32 *score,
                       u64 *best, const u8 *status, int num_sms, cudaStream_t st, const u32 *B_dev = nullptr,
                       const u32 *perm = nullptr);
cudaError_t launch_dtw_dyn(const void *in_ftr, u32 B, const void *bank, u32 T, u32 slot_stride, u32 flags, u32 *score,
                           u64 *best, const u8 *status, int num_sms, cudaStream_t st, u32 *max_frm_scratch,
                           const u32 *B_dev = nullptr, const u32 *perm = nullptr);
cudaError_t launch_dtw_band(const void *in_ftr, u32 B, const void *bank, u32 T, u32 slot_stride, u32 flags, int band_r,
                            u32 *score, u64 *best, int num_sms, cudaStream_t st);
cudaError_t launch_best_init(u64 *best, u32 B, cudaStream_t st);
cudaError_t launch_best_final(const u64 *best, u32 B, u32 *best_idx, u32 *best_dis, u32 *cmd, const u8 *status,
                              cudaStream_t st);
cudaError_t launch_status(const u32 *seg_off, const void *ftr, u32 B, u8 *status, cudaStream_t st);
cudaError_t launch_get_dis(const s16 *a, const s16 *b, u32 n, u32 *out, cudaStream_t st);
cudaError_t launch_dtw_limit(const u16 *x, const u16 *y, const u16 *I, const u16 *M, u32 n, u8 *out, cudaStream_t st);
cudaError_t launch_get_mdl(const void *in1, const void *in2, void *mdl, u32 n, u32 *dis, cudaStream_t st);
cudaError_t launch_pack_slots(const void *ftr, const u8 *status, u32 B, void *bank, u32 slot_stride, cudaStream_t st);
cudaError_t launch_sqrt_check(u32 lo, u32 hi, unsigned long long *bad_dev, cudaStream_t st);
cudaError_t launch_unpack12(const void *packed, u64 n_samples, u16 *out, cudaStream_t st);
class PackPool;
}  // namespace srk

using namespace srk;

inline thread_local std::string g_tls_error;

struct sr_comm;                                       // sr_comm.cu: NCCL communicator + its stream

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct sr_handle {
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;                 // H2D of the next chunk while the current one computes
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    uint64_t launches = 0;
    std::string err;
    // template bank
    const void *bank = nullptr;
    DevBuf bank_own;
    u32 n_slot = 0, slot_stride = 0;
    // bank slots in ascending frm_num order (only for banks wider than one 32-template tile): an ordering hint for the
    // greedy dtw kernels, recomputed when the bank pointer / geometry changes; correctness never depends on it
    DevBuf bank_perm;
    const u32 *perm = nullptr;
    const void *perm_bank = nullptr;
    u32 perm_n = 0, perm_stride = 0;
    // optional per-kernel timing (sr_timing_enable): event pairs recorded around every launch
    bool timing = false;
    std::vector<cudaEvent_t> ev;
    std::vector<uint32_t> ev_tag;
    size_t ev_used = 0;
    // packed PCM transport (sr_recognise_batch): worker pool, pinned staging slots, device staging
    int transport_mode = -1;                            // 0 off, 1 on, -1 automatic
    uint64_t auto_calls = 0;                            // automatic mode: calls seen, measured ns per PCM byte [plain, packed]
    double auto_ns_per_byte[2] = {0.0, 0.0};
    PackPool *pool = nullptr;
    static constexpr int kStage = 4;
    void *stage[kStage] = {nullptr, nullptr, nullptr, nullptr};
    size_t stage_cap = 0;
    uint32_t last_packed = 0, last_plain = 0, chunk_seq = 0;
    uint64_t last_h2d = 0;
    DevBuf dpacked;
    // command labels (commstr, main.c:25-31): n_labels records of label_stride bytes, NUL-terminated
    std::vector<uint8_t> labels;
    u32 n_labels = 0, label_stride = 0;
    sr_comm *comm = nullptr;                           // the exchange step (sr_comm_create), optional
    int dtw_variant = -1;                              // greedy dtw kernel: 0 static lane = pair (sr_dtw.cu), 1 dynamic pairs (sr_dtw_dyn.cu), -1 default
    DevBuf mfcc_work;                                  // the same for mfcc_kernel (next utterance, CTAs finished)
    DevBuf vad_work;                                   // two words: dynamic utterance hand-out of vad_kernel (zeroed once, self re-arming)
    DevBuf dtw_scratch;                                // one word: max frm_num of the current inputs (dynamic kernel's slot size)
    int geom = 0;                                      // SR_GEOM_REF (160/80/1024) or SR_GEOM_B (200/80/256, extension)
    int numa_node = -1;                                // node the device hangs off (-1 unknown / single node)
    // grow-only device workspaces
    DevBuf pcm, atap, seg, ftr, score, best, best_alt, status, bidx, bdis, cmd, misc0, misc1, misc2;
    int best_sel = 0;                                  // which of best / best_alt the current recognise call uses (alternates when a
                                                       // communicator is attached: the previous call's keys may still be being gathered)
};

inline int fail(sr_handle *h, const char *what, cudaError_t e) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s", what, e == cudaSuccess ? "invalid argument" : cudaGetErrorString(e));
    g_tls_error = buf;
    if (h) h->err = buf;
    return e == cudaSuccess ? -1 : (int)e;
}
#define SR_CK(h, call)                                           \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return fail((h), #call, e__);    \
    } while (0)
#define SR_REQUIRE(h, cond)                                      \
    do {                                                         \
        if (!(cond)) return fail((h), "requirement failed: " #cond, cudaSuccess); \
    } while (0)

inline cudaError_t ensure(DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return cudaSuccess;
    if (b.p) { cudaError_t e = cudaFree(b.p); b.p = nullptr; b.cap = 0; if (e != cudaSuccess) return e; }
    size_t want = bytes + (bytes >> 3) + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { b.p = nullptr; return e; }
    b.cap = want;
    return cudaSuccess;
}

inline u32 *mfcc_work(sr_handle *h) {
    if (!h->mfcc_work.p) {
        if (ensure(h->mfcc_work, 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cudaMemsetAsync(h->mfcc_work.p, 0, 16, h->stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    return static_cast<u32 *>(h->mfcc_work.p);
}

// get_mfcc in the handle's geometry
inline cudaError_t launch_mfcc_h(sr_handle *h, const u16 *pcm, u32 U, u32 B, const u32 *seg, u32 seg_stride, const atap_tag *atap,
                                 void *ftr, const u32 *row_map = nullptr, u32 rows_total = 0, const u32 *B_dev = nullptr) {
    if (h->geom == 1) return launch_mfcc_geomb(pcm, U, B, seg, seg_stride, atap, ftr, h->num_sms, h->stream, row_map, B_dev);
    return launch_mfcc(pcm, U, B, seg, seg_stride, atap, ftr, h->num_sms, h->stream, row_map, rows_total, B_dev, mfcc_work(h));
}

// the handle's work counters for vad_kernel (allocated and zeroed on first use)
inline u32 *vad_work(sr_handle *h) {
    if (!h->vad_work.p) {
        if (ensure(h->vad_work, 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (cudaMemsetAsync(h->vad_work.p, 0, 16, h->stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    return static_cast<u32 *>(h->vad_work.p);
}

// greedy dtw of B inputs against the handle's bank with the handle's kernel variant
#ifndef SR_DTW_VARIANT_DEFAULT
#define SR_DTW_VARIANT_DEFAULT 0
#endif
inline cudaError_t launch_dtw_h(sr_handle *h, const void *in_ftr, u32 B, u32 flags, u32 *score, u64 *best, const u8 *status,
                                const u32 *B_dev = nullptr) {
    int v = h->dtw_variant;
    if (v < 0) {
        static const int env_v = [] { const char *e = getenv("SR_DTW_VARIANT"); return e && *e ? atoi(e) : SR_DTW_VARIANT_DEFAULT; }();
        v = env_v;
    }
    if (v == 1) {
        cudaError_t e = ensure(h->dtw_scratch, 16);
        if (e != cudaSuccess) return e;
        return launch_dtw_dyn(in_ftr, B, h->bank, h->n_slot, h->slot_stride, flags, score, best, status, h->num_sms, h->stream,
                              static_cast<u32 *>(h->dtw_scratch.p), B_dev, h->perm);
    }
    return launch_dtw(in_ftr, B, h->bank, h->n_slot, h->slot_stride, flags, score, best, status, h->num_sms, h->stream, B_dev, h->perm);
}

int comm_wait_before_scan(sr_handle *h, const void *score);   // sr_comm.cu
int recognise_dev_impl(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, const sr_recog_out *o,
                       bool wait_comm);             // sr_api.cu

// records a (start,end) event pair around one kernel launch when timing is enabled
struct TimedLaunch {
    sr_handle *h;
    size_t slot = (size_t)-1;
    TimedLaunch(sr_handle *hh, uint32_t tag) : h(hh) {
        if (h->timing && (h->ev_used + 1) * 2 <= h->ev.size()) {
            slot = h->ev_used++;
            h->ev_tag[slot] = tag;
            cudaEventRecord(h->ev[2 * slot], h->stream);
        }
    }
    ~TimedLaunch() {
        if (slot != (size_t)-1) cudaEventRecord(h->ev[2 * slot + 1], h->stream);
    }
};
enum { TAG_VAD = 0, TAG_MFCC = 1, TAG_STATUS = 2, TAG_BEST_INIT = 3, TAG_DTW = 4, TAG_BEST_FINAL = 5, TAG_DTW_BAND = 6,
       TAG_FFT = 7, TAG_GET_DIS = 8 };

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};


#define H2D(h, dst, src, bytes) SR_CK(h, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, (h)->stream))
#define D2H(h, dst, src, bytes) SR_CK(h, cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, (h)->stream))
