// sr_api.cu -- the C-ABI of libspeech_b200.so (include/speech_recog.h): handle, device workspaces,
// host<->device plumbing and the reference-named batch-of-1 entry points. No arithmetic of the
// recognition path happens on the host: every result is produced by the kernels in sr_vad.cu,
// sr_mfcc.cu and sr_dtw.cu. Without a CUDA device every entry point fails loudly.
#include "sr_internal.h"
#include <chrono>
#include <condition_variable>
#include <deque>
#include <utility>
#include "sr_pack_host.h"
#include "sr_numa.h"
#include <map>
#include <algorithm>

#ifndef SR_TRANSPORT_AUTO_DEFAULT
#define SR_TRANSPORT_AUTO_DEFAULT 1      // what mode -1 (automatic) means: 1 = pack when this rank's share of the CPUs is >= 6
#endif

static int device_numa_node(int device) {
    char id[64] = {0};
    if (cudaDeviceGetPCIBusId(id, (int)sizeof id, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    return numa_node_of_pci(id);
}

extern "C" {

int sr_abi_version(void) { return 2; }

int sr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *sr_last_error(const sr_handle *h) { return h ? h->err.c_str() : g_tls_error.c_str(); }

int sr_create(int device, sr_handle **out) {
    if (!out) return fail(nullptr, "sr_create: out == NULL", cudaSuccess);
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, "sr_create: no CUDA device (libspeech_b200 has no CPU fallback)", e == cudaSuccess ? cudaErrorNoDevice : e);
    if (device < 0) SR_CK(nullptr, cudaGetDevice(&device));
    if (device >= n) return fail(nullptr, "sr_create: device ordinal out of range", cudaErrorInvalidDevice);
    sr_handle *h = new (std::nothrow) sr_handle;
    if (!h) return fail(nullptr, "sr_create: out of host memory", cudaErrorMemoryAllocation);
    h->device = device;
    DeviceGuard g(device);
    if (!g.ok) { delete h; return fail(nullptr, "sr_create: cudaSetDevice", cudaErrorInvalidDevice); }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete h; return fail(nullptr, "cudaGetDeviceProperties", e); }
    if (prop.major < 10) {
        delete h;
        return fail(nullptr, "sr_create: kernels are built for sm_100a (B200) only", cudaErrorInvalidDevice);
    }
    h->num_sms = prop.multiProcessorCount;
    h->numa_node = device_numa_node(device);
    // every failure from here on goes through sr_destroy, which releases whatever has been created so far
    e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    h->stream = h->own_stream;
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming | cudaEventBlockingSync);   // the packed transport's sender sleeps on it
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { sr_destroy(h); return fail(nullptr, "sr_create: stream/event creation", e); }
    if (!dev_tables()) { sr_destroy(h); return fail(nullptr, "sr_create: table upload", cudaErrorInitializationError); }
    *out = h;
    return 0;
}

int sr_destroy(sr_handle *h) {
    if (!h) return 0;
    DeviceGuard g(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    sr_comm_destroy(h);
    delete h->pool;
    for (void *&st : h->stage) if (st) { cudaFreeHost(st); st = nullptr; }
    DevBuf *bufs[] = {&h->dpacked, &h->bank_own, &h->pcm, &h->atap, &h->seg, &h->ftr, &h->score, &h->best, &h->best_alt, &h->status,
                      &h->bidx, &h->bdis, &h->cmd, &h->misc0, &h->misc1, &h->misc2, &h->dtw_scratch, &h->bank_perm, &h->vad_work, &h->mfcc_work};
    for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int i = 0; i < 2; ++i) { if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]); if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]); }
    delete h;
    return 0;
}

int sr_set_stream(sr_handle *h, void *cuda_stream) {
    SR_REQUIRE(h, h != nullptr);
    h->stream = static_cast<cudaStream_t>(cuda_stream);     // used verbatim: NULL is CUDA's legacy default stream
    return 0;
}

int sr_use_own_stream(sr_handle *h) {
    SR_REQUIRE(h, h != nullptr);
    h->stream = h->own_stream;
    return 0;
}

int sr_sync(sr_handle *h) {
    SR_REQUIRE(h, h != nullptr);
    DeviceGuard g(h->device);
    if (h->comm) { const int rc = sr_comm_wait(h); if (rc) return rc; }   // collectives issued so far are covered too
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int sr_set_geometry(sr_handle *h, int geom) {
    SR_REQUIRE(h, h && (geom == SR_GEOM_REF || geom == SR_GEOM_B));
    h->geom = geom;
    return 0;
}
int sr_get_geometry(const sr_handle *h) { return h ? h->geom : SR_GEOM_REF; }

int sr_set_dtw_variant(sr_handle *h, int variant) {
    SR_REQUIRE(h, h && variant >= -1 && variant <= 1);
    h->dtw_variant = variant;
    return 0;
}

uint64_t sr_launch_count(const sr_handle *h) { return h ? h->launches : 0; }

// ---- pinned host memory -------------------------------------------------------------------------------
// sr_host_alloc_dev places the pages on the NUMA node the GPU hangs off (first touch on that node's CPUs, then
// cudaHostRegister), so the H2D copy never crosses the socket interconnect; sr_host_alloc is the plain form.
static std::mutex g_host_mu;
static std::map<void *, size_t> g_node_allocs;           // node_alloc'ed + registered regions -> length

int sr_device_numa_node(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return -1; }
    return device_numa_node(device);
}

int sr_bind_thread_to_device(int device) {
    const int node = sr_device_numa_node(device);
    cpu_set_t want;
    if (node < 0 || numa_node_count() < 2 || !cpus_of_node(node, &want)) return -1;
    if (sched_setaffinity(0, sizeof want, &want) != 0) return -1;
    return node;
}

void *sr_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void *sr_host_alloc_dev(int device, size_t bytes) {
    if (bytes == 0) return nullptr;
    const int node = sr_device_numa_node(device);
    if (node < 0 || numa_node_count() < 2) return sr_host_alloc(bytes);     // single node: nothing to place
    void *p = node_alloc(bytes, node);
    if (!p) return nullptr;
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); node_free(p, bytes); return nullptr; }
    std::lock_guard<std::mutex> lk(g_host_mu);
    g_node_allocs[p] = bytes;
    return p;
}

void sr_host_free(void *p) {
    if (!p) return;
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_host_mu);
        auto it = g_node_allocs.find(p);
        if (it != g_node_allocs.end()) { bytes = it->second; g_node_allocs.erase(it); }
    }
    if (bytes) { cudaHostUnregister(p); node_free(p, bytes); }
    else cudaFreeHost(p);
}

int sr_host_numa_node(const void *p) { return p ? numa_node_of_page(p) : -1; }

// ---- per-kernel timing: event pairs on the launching stream around every kernel -----------------------
int sr_timing_enable(sr_handle *h, uint32_t max_records) {
    SR_REQUIRE(h, h != nullptr);
    DeviceGuard g(h->device);
    for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
    h->ev.clear(); h->ev_tag.clear(); h->ev_used = 0;
    h->timing = max_records > 0;
    for (uint32_t i = 0; i < 2 * max_records; ++i) {
        cudaEvent_t e;
        SR_CK(h, cudaEventCreate(&e));
        h->ev.push_back(e);
    }
    h->ev_tag.assign(max_records, 0);
    return 0;
}
int sr_timing_collect(sr_handle *h, uint32_t *tags, float *ms, uint32_t cap, uint32_t *n) {
    SR_REQUIRE(h, h && n);
    DeviceGuard g(h->device);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    uint32_t k = 0;
    for (size_t i = 0; i < h->ev_used && k < cap; ++i, ++k) {
        float t = 0.f;
        SR_CK(h, cudaEventElapsedTime(&t, h->ev[2 * i], h->ev[2 * i + 1]));
        if (tags) tags[k] = h->ev_tag[i];
        if (ms) ms[k] = t;
    }
    *n = k;
    h->ev_used = 0;
    return 0;
}

// ---- template bank --------------------------------------------------------------------------------
// Banks wider than one 32-template tile are walked in ascending frm_num order, so that the templates sharing a warp have
// similar walk lengths (CPU model: mean/max walk length per tile 0.81 -> 0.88 at T = 200). The order is a hint: scores and
// argmin keys carry the original slot numbers, and a stale order (bank rewritten in place) only costs efficiency.
static int bank_order(sr_handle *h, const unsigned char *hdr_host /* n_slot headers, 4 bytes each, or NULL: fetch */) {
    h->perm = nullptr;
    const u32 T = h->n_slot;
    if (T <= 32 || !h->bank) { h->perm_bank = h->bank; h->perm_n = T; h->perm_stride = h->slot_stride; return 0; }
    std::vector<u32> hdr(T);
    if (hdr_host) memcpy(hdr.data(), hdr_host, (size_t)T * 4);
    else {
        SR_CK(h, cudaMemcpy2DAsync(hdr.data(), 4, h->bank, h->slot_stride, 4, T, cudaMemcpyDeviceToHost, h->stream));
        SR_CK(h, cudaStreamSynchronize(h->stream));
    }
    std::vector<u32> order(T);
    for (u32 i = 0; i < T; ++i) order[i] = i;
    auto key = [&](u32 i) { const u32 f = hdr[i] >> 16; return f > 119u ? 0xFFFFu : f; };   // garbage headers last
    std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return key(a) < key(b); });
    SR_CK(h, ensure(h->bank_perm, (size_t)T * 4));
    SR_CK(h, cudaMemcpyAsync(h->bank_perm.p, order.data(), (size_t)T * 4, cudaMemcpyHostToDevice, h->stream));
    SR_CK(h, cudaStreamSynchronize(h->stream));                     // `order` is a local
    h->perm = static_cast<const u32 *>(h->bank_perm.p);
    h->perm_bank = h->bank; h->perm_n = T; h->perm_stride = h->slot_stride;
    return 0;
}

int sr_set_bank_dev(sr_handle *h, const void *bank_dev, uint32_t n_slot, uint32_t slot_stride) {
    SR_REQUIRE(h, h != nullptr);
    SR_REQUIRE(h, n_slot == 0 || (bank_dev != nullptr && slot_stride >= (uint32_t)kFtrBytes && slot_stride % 4 == 0));
    SR_REQUIRE(h, (reinterpret_cast<uintptr_t>(bank_dev) & 3) == 0);
    const bool same = h->perm_bank == bank_dev && h->perm_n == n_slot && h->perm_stride == slot_stride;
    h->bank = bank_dev; h->n_slot = n_slot; h->slot_stride = slot_stride;
    if (same) return 0;                                   // same buffer as last time (callers re-set it per batch): keep the order
    DeviceGuard g(h->device);
    return bank_order(h, nullptr);
}
int sr_set_bank(sr_handle *h, const void *bank, uint32_t n_slot, uint32_t slot_stride) {
    SR_REQUIRE(h, h != nullptr);
    SR_REQUIRE(h, n_slot == 0 || (bank != nullptr && slot_stride >= (uint32_t)kFtrBytes && slot_stride % 4 == 0));
    DeviceGuard g(h->device);
    const size_t bytes = (size_t)n_slot * slot_stride;
    SR_CK(h, cudaStreamSynchronize(h->stream));
    SR_CK(h, ensure(h->bank_own, bytes + 16));
    if (bytes) SR_CK(h, cudaMemcpyAsync(h->bank_own.p, bank, bytes, cudaMemcpyHostToDevice, h->stream));
    SR_CK(h, cudaStreamSynchronize(h->stream));
    h->bank = h->bank_own.p; h->n_slot = n_slot; h->slot_stride = slot_stride;
    std::vector<u32> hdr(n_slot);
    for (u32 i = 0; i < n_slot; ++i) memcpy(&hdr[i], static_cast<const unsigned char *>(bank) + (size_t)i * slot_stride, 4);
    return bank_order(h, reinterpret_cast<const unsigned char *>(hdr.data()));
}

// ---- command labels: commstr[] of main.c:25-31, what spch_recg returns (main.c:295) ---------------------------------
// default table = the reference's own 18 entries: "0 " .. "9 " and the GBK codes of up/down/front/back/left/right/big/small
static const uint8_t kRefLabels[18][3] = {
    {0x30, 0x20, 0}, {0x31, 0x20, 0}, {0x32, 0x20, 0}, {0x33, 0x20, 0}, {0x34, 0x20, 0}, {0x35, 0x20, 0}, {0x36, 0x20, 0},
    {0x37, 0x20, 0}, {0x38, 0x20, 0}, {0x39, 0x20, 0}, {0xC9, 0xCF, 0}, {0xCF, 0xC2, 0}, {0xC7, 0xB0, 0}, {0xBA, 0xF3, 0},
    {0xD7, 0xF3, 0}, {0xD3, 0xD2, 0}, {0xB4, 0xF3, 0}, {0xD0, 0xA1, 0}};

int sr_set_labels(sr_handle *h, const void *labels, uint32_t n_labels, uint32_t label_stride) {
    SR_REQUIRE(h, h && (n_labels == 0 || (labels && label_stride > 0)));
    h->labels.assign(static_cast<const uint8_t *>(labels), static_cast<const uint8_t *>(labels) + (size_t)n_labels * label_stride);
    h->n_labels = n_labels; h->label_stride = label_stride;
    return 0;
}

const uint8_t *sr_label(const sr_handle *h, uint32_t cmd) {
    if (h && h->label_stride) return cmd < h->n_labels ? h->labels.data() + (size_t)cmd * h->label_stride : nullptr;
    return cmd < 18u ? kRefLabels[cmd] : nullptr;                  // no table set: the reference's
}

int sr_labels_batch(const sr_handle *h, const uint32_t *cmd, const uint8_t *status, uint32_t B, const uint8_t **labels_out) {
    if (!cmd || !labels_out) return -1;
    for (uint32_t b = 0; b < B; ++b)                               // NULL = spch_recg's early returns (main.c:261-274)
        labels_out[b] = (status && status[b] != SR_ST_OK) ? nullptr : sr_label(h, cmd[b]);
    return 0;
}

// ---- device-pointer entry points ------------------------------------------------------------------
int sr_noise_atap_batch_dev(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, atap_tag *atap) {
    SR_REQUIRE(h, h && (B == 0 || (pcm && atap)));
    SR_REQUIRE(h, U <= 65535u && n_len <= 65535u);
    DeviceGuard g(h->device);
    { TimedLaunch tl(h, TAG_VAD); SR_CK(h, launch_vad(pcm, U, B, n_len, 0, 1, 0, atap, nullptr, h->num_sms, h->stream, vad_work(h))); }
    h->launches += B ? 1 : 0;
    return 0;
}

int sr_vad_batch_dev(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t buf_len, const atap_tag *atap,
                     uint32_t *seg_off) {
    SR_REQUIRE(h, h && (B == 0 || (pcm && atap && seg_off)));
    SR_REQUIRE(h, U <= 65535u && buf_len <= U);
    DeviceGuard g(h->device);
    { TimedLaunch tl(h, TAG_VAD); SR_CK(h, launch_vad(pcm, U, B, 0, buf_len, 0, 1, const_cast<atap_tag *>(atap), seg_off, h->num_sms, h->stream, vad_work(h))); }
    h->launches += B ? 1 : 0;
    return 0;
}

int sr_mfcc_batch_dev(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, const uint32_t *seg, uint32_t seg_stride,
                      const atap_tag *atap, v_ftr_tag *ftr) {
    SR_REQUIRE(h, h && (B == 0 || (pcm && seg && atap && ftr)));
    SR_REQUIRE(h, seg_stride >= 2 && (reinterpret_cast<uintptr_t>(ftr) & 3) == 0);
    DeviceGuard g(h->device);
    { TimedLaunch tl(h, TAG_MFCC); SR_CK(h, launch_mfcc_h(h, pcm, U, B, seg, seg_stride, atap, ftr)); }
    h->launches += B ? 1 : 0;
    return 0;
}

static int dtw_dev_impl(sr_handle *h, const v_ftr_tag *in, uint32_t B, uint32_t flags, int band_r, uint32_t *score,
                        uint32_t *best_idx, uint32_t *best_dis, uint32_t *cmd, const u8 *status) {
    SR_REQUIRE(h, h && (B == 0 || in));
    SR_REQUIRE(h, (reinterpret_cast<uintptr_t>(in) & 3) == 0);
    if (B == 0) return 0;
    const bool want_best = best_idx || best_dis || cmd;
    u64 *best = nullptr;
    if (want_best) {
        DevBuf &bb = h->best_sel ? h->best_alt : h->best;
        SR_CK(h, ensure(bb, (size_t)B * 8));
        best = static_cast<u64 *>(bb.p);
        { TimedLaunch tl(h, TAG_BEST_INIT); SR_CK(h, launch_best_init(best, B, h->stream)); }
        ++h->launches;
    }
    if (h->n_slot) {
        if (flags & SR_DTW_BAND) {
            SR_REQUIRE(h, band_r >= 0);
            TimedLaunch tl(h, TAG_DTW_BAND);
            SR_CK(h, launch_dtw_band(in, B, h->bank, h->n_slot, h->slot_stride, flags, band_r, score, best, h->num_sms, h->stream));
        } else {
            TimedLaunch tl(h, TAG_DTW);
            SR_CK(h, launch_dtw_h(h, in, B, flags, score, best, status));
        }
        ++h->launches;
    }
    if (want_best) {
        { TimedLaunch tl(h, TAG_BEST_FINAL); SR_CK(h, launch_best_final(best, B, best_idx, best_dis, cmd, status, h->stream)); }
        ++h->launches;
    }
    return 0;
}

int sr_dtw_batch_dev(sr_handle *h, const v_ftr_tag *in, uint32_t B, uint32_t flags, int band_r, uint32_t *score,
                     uint32_t *best_idx, uint32_t *best_dis) {
    SR_REQUIRE(h, h != nullptr);
    DeviceGuard g(h->device);
    return dtw_dev_impl(h, in, B, flags, band_r, score, best_idx, best_dis, nullptr, nullptr);
}

int sr_recognise_batch_dev(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len,
                           const sr_recog_out *o) {
    return recognise_dev_impl(h, pcm, U, B, n_len, o, false);
}

}  // extern "C"

// wait_comm: order the template scan (the first kernel that rewrites score / best) after the handle's pending collective,
// so that an all-gather of the previous batch overlaps this batch's VAD and MFCC (sr_comm.cu)
int recognise_dev_impl(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, const sr_recog_out *o,
                       bool wait_comm) {
    SR_REQUIRE(h, h && o && (B == 0 || pcm));
    SR_REQUIRE(h, U <= 65535u && n_len <= U);
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    atap_tag *atap = o->atap;
    if (!atap) {
        SR_CK(h, ensure(h->atap, (size_t)B * sizeof(atap_tag)));
        atap = static_cast<atap_tag *>(h->atap.p);
        SR_CK(h, cudaMemsetAsync(atap, 0, (size_t)B * sizeof(atap_tag), h->stream));
    }
    u32 *seg = o->seg_off;
    if (!seg) { SR_CK(h, ensure(h->seg, (size_t)B * 24)); seg = static_cast<u32 *>(h->seg.p); }
    v_ftr_tag *ftr = o->ftr;
    if (!ftr) { SR_CK(h, ensure(h->ftr, (size_t)B * kFtrBytes)); ftr = static_cast<v_ftr_tag *>(h->ftr.p); }
    u8 *status = o->status;
    if (!status) { SR_CK(h, ensure(h->status, (size_t)B)); status = static_cast<u8 *>(h->status.p); }
    // main.c:258-260 noise_atap + VAD (one fused launch on the staged utterance)
    { TimedLaunch tl(h, TAG_VAD); SR_CK(h, launch_vad(pcm, U, B, n_len, U, 1, 1, atap, seg, h->num_sms, h->stream, vad_work(h))); }
    // main.c:268 get_mfcc of segment 0
    { TimedLaunch tl(h, TAG_MFCC); SR_CK(h, launch_mfcc_h(h, pcm, U, B, seg, 6, atap, ftr)); }
    { TimedLaunch tl(h, TAG_STATUS); SR_CK(h, launch_status(seg, ftr, B, status, h->stream)); }
    h->launches += 3;
    if (h->comm) {                                       // collectives of earlier calls may still read score / the key buffer
        const int rc = wait_comm ? comm_wait_before_scan(h, o->score) : sr_comm_wait(h);
        if (rc) return rc;
    }
    // main.c:276-294 template scan, argmin, command index
    return dtw_dev_impl(h, ftr, B, SR_DTW_CHECK_SIGN, 0, o->score, o->best_idx, o->best_dis, o->cmd, status);
}

extern "C" {

// ---- host-buffer entry points ---------------------------------------------------------------------

int sr_noise_atap_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, atap_tag *atap) {
    SR_REQUIRE(h, h && (B == 0 || (pcm && atap)));
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->pcm, (size_t)B * U * 2 + 16));
    SR_CK(h, ensure(h->atap, (size_t)B * sizeof(atap_tag)));
    H2D(h, h->pcm.p, pcm, (size_t)B * U * 2);
    H2D(h, h->atap.p, atap, (size_t)B * sizeof(atap_tag));
    int rc = sr_noise_atap_batch_dev(h, static_cast<const u16 *>(h->pcm.p), U, B, n_len, static_cast<atap_tag *>(h->atap.p));
    if (rc) return rc;
    D2H(h, atap, h->atap.p, (size_t)B * sizeof(atap_tag));
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int sr_vad_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t buf_len, const atap_tag *atap,
                 uint32_t *seg_off) {
    SR_REQUIRE(h, h && (B == 0 || (pcm && atap && seg_off)));
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->pcm, (size_t)B * U * 2 + 16));
    SR_CK(h, ensure(h->atap, (size_t)B * sizeof(atap_tag)));
    SR_CK(h, ensure(h->seg, (size_t)B * 24));
    H2D(h, h->pcm.p, pcm, (size_t)B * U * 2);
    H2D(h, h->atap.p, atap, (size_t)B * sizeof(atap_tag));
    int rc = sr_vad_batch_dev(h, static_cast<const u16 *>(h->pcm.p), U, B, buf_len, static_cast<const atap_tag *>(h->atap.p),
                              static_cast<u32 *>(h->seg.p));
    if (rc) return rc;
    D2H(h, seg_off, h->seg.p, (size_t)B * 24);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int sr_mfcc_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, const uint32_t *seg, uint32_t seg_stride,
                  const atap_tag *atap, v_ftr_tag *ftr) {
    SR_REQUIRE(h, h && (B == 0 || (pcm && seg && atap && ftr)));
    SR_REQUIRE(h, seg_stride >= 2);
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->pcm, (size_t)B * U * 2 + 16));
    SR_CK(h, ensure(h->atap, (size_t)B * sizeof(atap_tag)));
    SR_CK(h, ensure(h->seg, (size_t)B * seg_stride * 4));
    SR_CK(h, ensure(h->ftr, (size_t)B * kFtrBytes));
    H2D(h, h->pcm.p, pcm, (size_t)B * U * 2);
    H2D(h, h->atap.p, atap, (size_t)B * sizeof(atap_tag));
    H2D(h, h->seg.p, seg, (size_t)B * seg_stride * 4);
    int rc = sr_mfcc_batch_dev(h, static_cast<const u16 *>(h->pcm.p), U, B, static_cast<const u32 *>(h->seg.p), seg_stride,
                               static_cast<const atap_tag *>(h->atap.p), static_cast<v_ftr_tag *>(h->ftr.p));
    if (rc) return rc;
    // MFCC.C never writes save_sign: copy back bytes [2, 2860) of every struct only
    SR_CK(h, cudaMemcpy2DAsync(reinterpret_cast<unsigned char *>(ftr) + 2, kFtrBytes,
                               static_cast<unsigned char *>(h->ftr.p) + 2, kFtrBytes, kFtrBytes - 2, B,
                               cudaMemcpyDeviceToHost, h->stream));
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int sr_dtw_batch(sr_handle *h, const v_ftr_tag *in, uint32_t B, uint32_t flags, int band_r, uint32_t *score,
                 uint32_t *best_idx, uint32_t *best_dis) {
    SR_REQUIRE(h, h && (B == 0 || in));
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    const size_t T = h->n_slot;
    SR_CK(h, ensure(h->ftr, (size_t)B * kFtrBytes));
    if (score) SR_CK(h, ensure(h->score, (size_t)B * T * 4 + 4));
    SR_CK(h, ensure(h->bidx, (size_t)B * 4));
    SR_CK(h, ensure(h->bdis, (size_t)B * 4));
    H2D(h, h->ftr.p, in, (size_t)B * kFtrBytes);
    int rc = dtw_dev_impl(h, static_cast<const v_ftr_tag *>(h->ftr.p), B, flags, band_r,
                          score ? static_cast<u32 *>(h->score.p) : nullptr,
                          best_idx ? static_cast<u32 *>(h->bidx.p) : nullptr,
                          best_dis ? static_cast<u32 *>(h->bdis.p) : nullptr, nullptr, nullptr);
    if (rc) return rc;
    if (score && T) D2H(h, score, h->score.p, (size_t)B * T * 4);
    if (best_idx) D2H(h, best_idx, h->bidx.p, (size_t)B * 4);
    if (best_dis) D2H(h, best_dis, h->bdis.p, (size_t)B * 4);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// ranks of this node that share the host (torchrun exports LOCAL_WORLD_SIZE)
static int local_world_size() {
    static const int local_world = [] { const char *e = getenv("LOCAL_WORLD_SIZE"); const int v = e ? atoi(e) : 1; return v > 0 ? v : 1; }();
    return local_world;
}
// CPUs this rank may count on: the process' usable CPUs (affinity capped by the cgroup quota) divided by those ranks
static int rank_cpu_share() { return usable_cpus() / local_world_size(); }

// local ranks whose GPU hangs off the same NUMA node as this handle's (torchrun convention: local rank r drives device r);
// unknown topology counts everybody
static int ranks_on_socket(const sr_handle *h) {
    const int W = local_world_size();
    if (W <= 1) return 1;
    if (h->numa_node < 0) return W;
    int n = 0;
    for (int d = 0; d < W; ++d) if (device_numa_node(d) == h->numa_node) ++n;
    return n < 1 ? 1 : n;
}

// 1 = forced on, 0 = forced off, -1 = automatic (decided per call by transport_auto_pick)
static int transport_mode(const sr_handle *h) {
    int mode = h->transport_mode;
    if (mode < 0) {
        static const int env_mode = [] { const char *e = getenv("SR_PACK12"); return e && *e ? atoi(e) : -1; }();
        mode = env_mode;
    }
    // automatic: needs CPUs to pack with, and the socket's DRAM bandwidth to itself: with several GPUs per socket the DMA
    // reads alone load it (4 x 54 GB/s at four) and packing measured 25.8 vs 19.5 ms at 4 and 8 ranks, 19.8-21.3 vs 19.5 with
    // two ranks on one socket -- and ranks that probe at different moments talk each other into it. One rank per socket only.
    if (mode < 0 && !(SR_TRANSPORT_AUTO_DEFAULT && rank_cpu_share() >= 6 && ranks_on_socket(h) <= 1)) mode = 0;
    return mode;
}

// Automatic mode measures instead of guessing. Whether packing pays depends on what else loads the host's memory system:
// one or two ranks per socket gain ~16 % (16.3 vs 19.4 ms per 1.05 GB), but with four ranks per socket the DMA reads
// alone take ~216 GB/s of that socket's DRAM bandwidth and the packers' extra traffic makes the call SLOWER (25.8 vs
// 19.5 ms, measured at 4 and 8 GPUs). So: the first qualifying call goes plain, the second packed, then the faster of
// the two (ns per byte, exponentially averaged; packing must win by 7 %) is used, with the other re-probed every 32nd call.
// With more than one rank on this GPU's socket the automatic mode stays plain (transport_mode above).
static bool transport_auto_pick(sr_handle *h) {
    const uint64_t n = h->auto_calls++;
    if (h->auto_ns_per_byte[0] <= 0.0) return false;
    if (h->auto_ns_per_byte[1] <= 0.0) return true;
    const bool packed_better = h->auto_ns_per_byte[1] < 0.93 * h->auto_ns_per_byte[0];   // a clear win only (N = 1: 0.84)
    if (n % 32 == 31) return !packed_better;                       // probe the loser now and then: conditions change
    return packed_better;
}
static void transport_auto_record(sr_handle *h, bool packed, double ns_per_byte) {
    double &v = h->auto_ns_per_byte[packed ? 1 : 0];
    v = v <= 0.0 ? ns_per_byte : 0.75 * v + 0.25 * ns_per_byte;
}

int sr_set_transport(sr_handle *h, int mode) {
    SR_REQUIRE(h, h && mode >= -1 && mode <= 1);
    h->transport_mode = mode;
    return 0;
}

int sr_transport_stats(const sr_handle *h, uint32_t *packed_chunks, uint32_t *plain_chunks, uint64_t *h2d_bytes) {
    if (!h) return -1;
    if (packed_chunks) *packed_chunks = h->last_packed;
    if (plain_chunks) *plain_chunks = h->last_plain;
    if (h2d_bytes) *h2d_bytes = h->last_h2d;
    return 0;
}

uint32_t sr_debug_pack12_host(int variant, const uint16_t *src, uint64_t n, uint8_t *dst) {
    if (!src || !dst || (n & 1)) return 0xFFFFFFFFu;
    if (variant >= 100) { PackPool pool(variant - 100); uint32_t o = 0; for (int rep = 0; rep < 3; ++rep) o = pool.run(src, (size_t)n, dst); return o; }   // the worker pool (3 fork-joins)
    return variant < 0 ? pack12(src, (size_t)n, dst) : pack12_variant(variant, src, (size_t)n, dst);
}

int sr_debug_unpack12(sr_handle *h, const uint8_t *packed, uint64_t n, uint16_t *out) {
    SR_REQUIRE(h, h && packed && out && !(n & 1));
    if (n == 0) return 0;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->misc0, (size_t)(n / 2 * 3) + 64));
    SR_CK(h, ensure(h->misc1, (size_t)n * 2 + 64));
    H2D(h, h->misc0.p, packed, (size_t)(n / 2 * 3));
    SR_CK(h, launch_unpack12(h->misc0.p, n, static_cast<u16 *>(h->misc1.p), h->stream));
    ++h->launches;
    D2H(h, out, h->misc1.p, (size_t)n * 2);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// Host-buffer spch_recg for B utterances. Large batches are processed in chunks through two device PCM
// buffers: the H2D copy of chunk c+1 (copy stream) overlaps the kernels of chunk c (compute stream), so
// with pinned host memory the call is bound by max(PCIe, compute) instead of their sum.
int sr_recognise_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, const sr_recog_out *o) {
    SR_REQUIRE(h, h && o && (B == 0 || pcm));
    SR_REQUIRE(h, (B == 0 || U > 0) && U <= 65535u && n_len <= U);
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    const size_t T = h->n_slot;
    // chunk: ~32 MB of PCM (measured best of 16..256 MB on B200; SR_CHUNK_MB overrides), a multiple of 8 utterances (keeps every chunk base 16-byte aligned)
    static const size_t chunk_mb = [] { const char *e = getenv("SR_CHUNK_MB"); const long v = e ? atol(e) : 0; return (size_t)(v > 0 ? v : 32); }();
    uint32_t chunk = (uint32_t)((chunk_mb << 20) / ((size_t)U * 2));
    chunk = chunk < 8 ? 8 : (chunk & ~7u);
    if (chunk > B) chunk = B;
    const uint32_t nchunks = (B + chunk - 1) / chunk;
    const size_t chunk_bytes = (((size_t)chunk * U * 2 + 255) / 256) * 256;
    SR_CK(h, ensure(h->pcm, (nchunks > 1 ? 2 : 1) * chunk_bytes + 16));
    sr_recog_out d;
    memset(&d, 0, sizeof d);
    if (o->atap) { SR_CK(h, ensure(h->atap, (size_t)B * sizeof(atap_tag))); d.atap = static_cast<atap_tag *>(h->atap.p);
                   H2D(h, d.atap, o->atap, (size_t)B * sizeof(atap_tag)); }
    if (o->seg_off) { SR_CK(h, ensure(h->seg, (size_t)B * 24)); d.seg_off = static_cast<u32 *>(h->seg.p); }
    if (o->ftr) { SR_CK(h, ensure(h->ftr, (size_t)B * kFtrBytes)); d.ftr = static_cast<v_ftr_tag *>(h->ftr.p); }
    if (o->score) { SR_CK(h, ensure(h->score, (size_t)B * T * 4 + 4)); d.score = static_cast<u32 *>(h->score.p); }
    if (o->best_idx) { SR_CK(h, ensure(h->bidx, (size_t)B * 4)); d.best_idx = static_cast<u32 *>(h->bidx.p); }
    if (o->best_dis) { SR_CK(h, ensure(h->bdis, (size_t)B * 4)); d.best_dis = static_cast<u32 *>(h->bdis.p); }
    if (o->cmd) { SR_CK(h, ensure(h->cmd, (size_t)B * 4)); d.cmd = static_cast<u32 *>(h->cmd.p); }
    if (o->status) { SR_CK(h, ensure(h->status, (size_t)B)); d.status = static_cast<u8 *>(h->status.p); }
    // one chunk: H2D (plain u16, or 12-bit packed from a pinned staging slot + expansion on the device) -> kernels
    auto issue_chunk = [&](uint32_t c, int buf, const void *packed_src) -> int {
        const uint32_t b0 = c * chunk, nb = (b0 + chunk <= B) ? chunk : B - b0;
        const size_t ns = (size_t)nb * U;
        u16 *dpcm = reinterpret_cast<u16 *>(static_cast<unsigned char *>(h->pcm.p) + (size_t)buf * chunk_bytes);
        cudaStream_t cs = nchunks > 1 ? h->copy_stream : h->stream;
        if (nchunks > 1 && h->chunk_seq >= 2) SR_CK(h, cudaStreamWaitEvent(cs, h->ev_done[buf], 0));      // buffers free again
        if (packed_src) {
            unsigned char *dpk = static_cast<unsigned char *>(h->dpacked.p) + (size_t)buf * h->stage_cap;
            SR_CK(h, cudaMemcpyAsync(dpk, packed_src, ns / 2 * 3, cudaMemcpyHostToDevice, cs));
            h->last_h2d += ns / 2 * 3; ++h->last_packed;
        } else {
            SR_CK(h, cudaMemcpyAsync(dpcm, pcm + (size_t)b0 * U, ns * 2, cudaMemcpyHostToDevice, cs));
            h->last_h2d += ns * 2; ++h->last_plain;
        }
        if (nchunks > 1) {
            SR_CK(h, cudaEventRecord(h->ev_h2d[buf], cs));
            SR_CK(h, cudaStreamWaitEvent(h->stream, h->ev_h2d[buf], 0));
        }
        if (packed_src) {
            SR_CK(h, launch_unpack12(static_cast<unsigned char *>(h->dpacked.p) + (size_t)buf * h->stage_cap, ns, dpcm, h->stream));
            ++h->launches;
        }
        sr_recog_out dc = d;
        if (d.atap) dc.atap = d.atap + b0;
        if (d.seg_off) dc.seg_off = d.seg_off + (size_t)b0 * 6;
        if (d.ftr) dc.ftr = d.ftr + b0;
        if (d.score) dc.score = d.score + (size_t)b0 * T;
        if (d.best_idx) dc.best_idx = d.best_idx + b0;
        if (d.best_dis) dc.best_dis = d.best_dis + b0;
        if (d.cmd) dc.cmd = d.cmd + b0;
        if (d.status) dc.status = d.status + b0;
        int rc = sr_recognise_batch_dev(h, dpcm, U, nb, n_len, &dc);
        if (rc) return rc;
        if (nchunks > 1) SR_CK(h, cudaEventRecord(h->ev_done[buf], h->stream));
        ++h->chunk_seq;
        return 0;
    };
    h->chunk_seq = 0; h->last_packed = 0; h->last_plain = 0; h->last_h2d = 0;

    const int tmode = nchunks >= 4 ? transport_mode(h) : 0;
    const bool tauto = tmode < 0;
    bool packed_transport = tmode > 0 || (tauto && transport_auto_pick(h));
    const auto t_call0 = std::chrono::steady_clock::now();
    bool did_setup = false;                              // this call created the pool / staging: its time is not a measurement
    if (packed_transport) {                              // workers, pinned staging slots, device staging
        const size_t pk = ((((size_t)chunk * U + 1) / 2 * 3 + 64 + 255) / 256) * 256;
        did_setup = !h->pool || h->stage_cap < pk || h->dpacked.cap < 2 * pk;
        ScopedNodeAffinity node_scope(h->numa_node);      // workers inherit it; staging pages are allocated from this node
        if (!h->pool) {
            // packers = this rank's CPU share minus room for the sender, the CUDA runtime's threads and the caller's own work
            static const int env_nt = [] { const char *e = getenv("SR_PACK_THREADS"); return e && *e ? atoi(e) : 0; }();
            // (measured on a 2 x 32-core host, 16-CPU quota: 8..12 packers all land at ~16.3 ms per 1.05 GB step; more only add
            // memory traffic next to the DMA reads, which slows the link: 54 -> 47 GB/s at 14 packers)
            int nt = env_nt > 0 ? env_nt : rank_cpu_share() - 3;
            if (env_nt <= 0 && nt > 10) nt = 10;
            nt = nt > 16 ? 16 : nt;
            if (nt >= 2) h->pool = new (std::nothrow) PackPool(nt);
        }
        if (h->pool && h->stage_cap < pk) {
            for (void *&st : h->stage) if (st) { cudaFreeHost(st); st = nullptr; }
            h->stage_cap = 0;
            bool ok = true;
            // SR_PACK_WC=1: write-combined staging (the packers only ever stream whole cache lines into it and never read it back)
            static const unsigned stage_flags = [] { const char *e = getenv("SR_PACK_WC"); return e && atoi(e) > 0 ? cudaHostAllocWriteCombined : cudaHostAllocDefault; }();
            for (void *&st : h->stage) if (ok && cudaHostAlloc(&st, pk, stage_flags) != cudaSuccess) { st = nullptr; ok = false; }
            if (ok) h->stage_cap = pk;
            else { cudaGetLastError(); for (void *&st : h->stage) if (st) { cudaFreeHost(st); st = nullptr; } }
        }
        if (!h->pool || !h->stage_cap || ensure(h->dpacked, 2 * h->stage_cap) != cudaSuccess) { cudaGetLastError(); packed_transport = false; }
    }

    if (!packed_transport) {
        for (uint32_t c = 0; c < nchunks; ++c) {
            int rc = issue_chunk(c, (int)(c & 1), nullptr);
            if (rc) return rc;
        }
    } else {
        // The caller's thread sends chunks from the front as plain u16, paced by the copy engine; the worker pool packs
        // chunks from the back into the staging slots and those are sent packed as soon as they are ready. The two
        // meet in the middle, so the call is never slower than the plain path and approaches 3/4 of its PCIe time.
        std::mutex m;
        std::condition_variable cv_slot;
        std::deque<std::pair<uint32_t, int>> ready;       // (chunk, staging slot)
        std::vector<uint32_t> retry;                      // chunks that hold a sample >= 4096: sent plain
        int lo = 0, hi = (int)nchunks - 1;                // unclaimed chunks [lo, hi]
        unsigned free_mask = (1u << sr_handle::kStage) - 1u;
        bool abort = false;
        std::thread packer([&] {
            ScopedNodeAffinity bind(h->numa_node);        // slice 0 of every chunk is packed by this thread
            for (;;) {
                int slot;
                uint32_t c;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv_slot.wait(lk, [&] { return abort || lo > hi || free_mask != 0; });
                    if (abort || lo > hi) return;
                    slot = __builtin_ctz(free_mask);
                    free_mask &= ~(1u << slot);
                    c = (uint32_t)hi--;
                }
                const uint32_t b0 = c * chunk, nb = (b0 + chunk <= B) ? chunk : B - b0;
                const size_t ns = (size_t)nb * U;
                const uint32_t orb = (ns & 1) ? 0xFFFFu : h->pool->run(pcm + (size_t)b0 * U, ns, static_cast<uint8_t *>(h->stage[slot]));
                {
                    std::lock_guard<std::mutex> lk(m);
                    if (orb & 0xF000u) { retry.push_back(c); free_mask |= 1u << slot; }
                    else ready.emplace_back(c, slot);
                }
            }
        });
        int rc = 0;
        int slot_of[2] = {-1, -1};                        // staging slot behind the copy last issued on each buffer
        auto release = [&](int &sl) {
            if (sl < 0) return;
            { std::lock_guard<std::mutex> lk(m); free_mask |= 1u << sl; }
            cv_slot.notify_one();
            sl = -1;
        };
        uint32_t sent = 0;
        while (sent < nchunks) {
            int slot = -1;
            long c = -1;
            {
                std::lock_guard<std::mutex> lk(m);
                if (!ready.empty()) { c = ready.front().first; slot = ready.front().second; ready.pop_front(); }
                else if (!retry.empty()) { c = retry.back(); retry.pop_back(); }
                else if (lo <= hi) c = lo++;
            }
            if (c < 0) { std::this_thread::sleep_for(std::chrono::microseconds(50)); continue; }   // every chunk is claimed; the pool is still packing
            const int buf = (int)(sent & 1);
            if (sent >= 2) {                                               // at most two copies in flight: paces this thread
                cudaError_t e = cudaEventSynchronize(h->ev_h2d[buf]);
                if (e != cudaSuccess) { rc = fail(h, "cudaEventSynchronize", e); if (slot >= 0) release(slot); break; }
                release(slot_of[buf]);
            }
            rc = issue_chunk((uint32_t)c, buf, slot >= 0 ? h->stage[slot] : nullptr);
            if (rc) { if (slot >= 0) release(slot); break; }
            slot_of[buf] = slot;
            ++sent;
        }
        { std::lock_guard<std::mutex> lk(m); abort = true; }
        cv_slot.notify_all();
        packer.join();
        if (rc) { cudaStreamSynchronize(h->copy_stream); return rc; }
    }
    if (o->atap) D2H(h, o->atap, d.atap, (size_t)B * sizeof(atap_tag));
    if (o->seg_off) D2H(h, o->seg_off, d.seg_off, (size_t)B * 24);
    if (o->ftr)
        SR_CK(h, cudaMemcpy2DAsync(reinterpret_cast<unsigned char *>(o->ftr) + 2, kFtrBytes,
                                   reinterpret_cast<unsigned char *>(d.ftr) + 2, kFtrBytes, kFtrBytes - 2, B,
                                   cudaMemcpyDeviceToHost, h->stream));
    if (o->score && T) D2H(h, o->score, d.score, (size_t)B * T * 4);
    if (o->best_idx) D2H(h, o->best_idx, d.best_idx, (size_t)B * 4);
    if (o->best_dis) D2H(h, o->best_dis, d.best_dis, (size_t)B * 4);
    if (o->cmd) D2H(h, o->cmd, d.cmd, (size_t)B * 4);
    if (o->status) D2H(h, o->status, d.status, (size_t)B);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    if (tauto && !did_setup)
        transport_auto_record(h, packed_transport && h->last_packed > 0,
                              std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t_call0).count() / ((double)B * U * 2.0));
    return 0;
}

// save_mdl (main.c:121-138) for B utterances: noise_atap -> VAD -> get_mfcc(seg 0) -> save_ftr_mdl into slot b of a
// flash-layout bank image (host memory, B x slot_stride bytes). status[b]: 0 save_ok, 1 VAD_fail, 2 MFCC_fail
// (main.c:38-40); failed slots stay erased (0xFF). The result can be handed to sr_set_bank unchanged.
int sr_enrol_batch(sr_handle *h, const uint16_t *pcm, uint32_t U, uint32_t B, uint32_t n_len, void *bank_out,
                   uint32_t slot_stride, uint8_t *status) {
    SR_REQUIRE(h, h && (B == 0 || (pcm && bank_out)));
    SR_REQUIRE(h, (B == 0 || U > 0) && U <= 65535u && n_len <= U && slot_stride >= (uint32_t)kFtrBytes && slot_stride % 4 == 0);
    if (B == 0) return 0;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->pcm, (size_t)B * U * 2 + 16));
    SR_CK(h, ensure(h->atap, (size_t)B * sizeof(atap_tag)));
    SR_CK(h, ensure(h->seg, (size_t)B * 24));
    SR_CK(h, ensure(h->ftr, (size_t)B * kFtrBytes));
    SR_CK(h, ensure(h->status, (size_t)B));
    SR_CK(h, ensure(h->misc0, (size_t)B * slot_stride));
    H2D(h, h->pcm.p, pcm, (size_t)B * U * 2);
    SR_CK(h, cudaMemsetAsync(h->atap.p, 0, (size_t)B * sizeof(atap_tag), h->stream));
    { TimedLaunch tl(h, TAG_VAD); SR_CK(h, launch_vad(static_cast<const u16 *>(h->pcm.p), U, B, n_len, U, 1, 1, static_cast<atap_tag *>(h->atap.p), static_cast<u32 *>(h->seg.p), h->num_sms, h->stream, vad_work(h))); }
    { TimedLaunch tl(h, TAG_MFCC); SR_CK(h, launch_mfcc_h(h, static_cast<const u16 *>(h->pcm.p), U, B, static_cast<const u32 *>(h->seg.p), 6, static_cast<const atap_tag *>(h->atap.p), h->ftr.p)); }
    SR_CK(h, launch_status(static_cast<const u32 *>(h->seg.p), h->ftr.p, B, static_cast<u8 *>(h->status.p), h->stream));
    SR_CK(h, launch_pack_slots(h->ftr.p, static_cast<const u8 *>(h->status.p), B, h->misc0.p, slot_stride, h->stream));
    h->launches += 4;
    D2H(h, bank_out, h->misc0.p, (size_t)B * slot_stride);
    if (status) D2H(h, status, h->status.p, (size_t)B);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// get_mdl (DTW.C:217-296) for n pairs: mdl[p] = average of in1[p], in2[p] along their greedy DTW path; dis[p] = the
// path's step-normalised distance (dis_err and mdl[p] untouched when the 2:1 length guard rejects the pair).
int sr_get_mdl_batch(sr_handle *h, const v_ftr_tag *in1, const v_ftr_tag *in2, uint32_t n, v_ftr_tag *mdl, uint32_t *dis) {
    SR_REQUIRE(h, h && (n == 0 || (in1 && in2 && mdl)));
    if (n == 0) return 0;
    DeviceGuard g(h->device);
    const size_t bytes = (size_t)n * kFtrBytes;
    SR_CK(h, ensure(h->misc0, bytes));
    SR_CK(h, ensure(h->misc1, bytes));
    SR_CK(h, ensure(h->ftr, bytes));
    SR_CK(h, ensure(h->bdis, (size_t)n * 4));
    H2D(h, h->misc0.p, in1, bytes);
    H2D(h, h->misc1.p, in2, bytes);
    H2D(h, h->ftr.p, mdl, bytes);                                   // rejected pairs leave mdl as the caller passed it
    SR_CK(h, launch_get_mdl(h->misc0.p, h->misc1.p, h->ftr.p, n, static_cast<u32 *>(h->bdis.p), h->stream));
    ++h->launches;
    // like get_mfcc, get_mdl never writes save_sign: copy back bytes [2, 2860) only
    SR_CK(h, cudaMemcpy2DAsync(reinterpret_cast<unsigned char *>(mdl) + 2, kFtrBytes, static_cast<unsigned char *>(h->ftr.p) + 2,
                               kFtrBytes, kFtrBytes - 2, n, cudaMemcpyDeviceToHost, h->stream));
    if (dis) D2H(h, dis, h->bdis.p, (size_t)n * 4);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// One host call, several GPUs: the batch is cut into contiguous shards, shard g runs on handles[g]
// from its own host thread, and every shard writes its results straight into its slice of the caller's host
// arrays -- with host outputs the "gather" is the D2H copies themselves, no collective is needed. (Device-resident
// multi-GPU use is one process per GPU with a NCCL all-gather of the score blocks, see bench.py.)
// All handles must have the same template bank set. Returns the first non-zero shard status.
int sr_recognise_batch_multi(sr_handle *const *handles, uint32_t n_handles, const uint16_t *pcm, uint32_t U, uint32_t B,
                             uint32_t n_len, const sr_recog_out *o) {
    if (!handles || n_handles == 0 || !o) return fail(nullptr, "sr_recognise_batch_multi: bad arguments", cudaSuccess);
    for (uint32_t g = 0; g < n_handles; ++g)
        if (!handles[g] || handles[g]->n_slot != handles[0]->n_slot) return fail(nullptr, "sr_recognise_batch_multi: handles differ", cudaSuccess);
    const size_t T = handles[0]->n_slot;
    std::vector<int> rc(n_handles, 0);
    std::vector<std::thread> th;
    for (uint32_t g = 0; g < n_handles; ++g) {
        const uint32_t lo = (uint32_t)((uint64_t)B * g / n_handles), hi = (uint32_t)((uint64_t)B * (g + 1) / n_handles);
        sr_recog_out s = *o;
        if (s.atap) s.atap += lo;
        if (s.seg_off) s.seg_off += (size_t)lo * 6;
        if (s.ftr) s.ftr += lo;
        if (s.score) s.score += (size_t)lo * T;
        if (s.best_idx) s.best_idx += lo;
        if (s.best_dis) s.best_dis += lo;
        if (s.cmd) s.cmd += lo;
        if (s.status) s.status += lo;
        th.emplace_back([=, &rc]() { rc[g] = sr_recognise_batch(handles[g], pcm + (size_t)lo * U, U, hi - lo, n_len, &s); });
    }
    for (auto &t : th) t.join();
    for (uint32_t g = 0; g < n_handles; ++g) if (rc[g]) return rc[g];
    return 0;
}

int sr_fft_mag_batch(sr_handle *h, const int16_t *frames, uint32_t len, uint32_t n, uint32_t *mag) {
    SR_REQUIRE(h, h && (n == 0 || (frames && mag)));
    SR_REQUIRE(h, len <= SR_FFT_POINT);                                   // MFCC.C:32-35
    if (n == 0) return 0;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->misc0, (size_t)n * len * 2 + 16));
    SR_CK(h, ensure(h->misc1, (size_t)n * 512 * 4));
    if (len) H2D(h, h->misc0.p, frames, (size_t)n * len * 2);
    SR_CK(h, launch_fft_generic(nullptr, static_cast<const s16 *>(h->misc0.p), len, n, nullptr, static_cast<u32 *>(h->misc1.p), h->stream));
    ++h->launches;
    D2H(h, mag, h->misc1.p, (size_t)n * 512 * 4);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// dtw_limit (DTW.C:76-109) for n points, explicit (I, M) per point instead of the reference's file statics
int sr_dtw_limit_batch(sr_handle *h, const uint16_t *x, const uint16_t *y, const uint16_t *I, const uint16_t *M, uint32_t n,
                       uint8_t *out) {
    SR_REQUIRE(h, h && (n == 0 || (x && y && I && M && out)));
    if (n == 0) return 0;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->misc0, (size_t)n * 8 + 16));
    SR_CK(h, ensure(h->misc2, (size_t)n + 16));
    u16 *d = static_cast<u16 *>(h->misc0.p);
    H2D(h, d, x, (size_t)n * 2); H2D(h, d + n, y, (size_t)n * 2); H2D(h, d + 2 * (size_t)n, I, (size_t)n * 2); H2D(h, d + 3 * (size_t)n, M, (size_t)n * 2);
    SR_CK(h, launch_dtw_limit(d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, n, static_cast<u8 *>(h->misc2.p), h->stream));
    ++h->launches;
    D2H(h, out, h->misc2.p, (size_t)n);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// raw FFT of packed (re | im<<16) 1024-point inputs -- test hook for the asm restatement parity
int sr_fft_raw_batch(sr_handle *h, const uint32_t *in_packed, uint32_t n, uint32_t *out_packed) {
    SR_REQUIRE(h, h && (n == 0 || (in_packed && out_packed)));
    if (n == 0) return 0;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->misc0, (size_t)n * 4096));
    SR_CK(h, ensure(h->misc1, (size_t)n * 4096));
    H2D(h, h->misc0.p, in_packed, (size_t)n * 4096);
    SR_CK(h, launch_fft_generic(static_cast<const u32 *>(h->misc0.p), nullptr, 0, n, static_cast<u32 *>(h->misc1.p), nullptr, h->stream));
    ++h->launches;
    D2H(h, out_packed, h->misc1.p, (size_t)n * 4096);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int sr_get_dis_batch(sr_handle *h, const int16_t *a, const int16_t *b, uint32_t n, uint32_t *dis) {
    SR_REQUIRE(h, h && (n == 0 || (a && b && dis)));
    if (n == 0) return 0;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->misc0, (size_t)n * 24));
    SR_CK(h, ensure(h->misc1, (size_t)n * 24));
    SR_CK(h, ensure(h->misc2, (size_t)n * 4));
    H2D(h, h->misc0.p, a, (size_t)n * 24);
    H2D(h, h->misc1.p, b, (size_t)n * 24);
    SR_CK(h, launch_get_dis(static_cast<const s16 *>(h->misc0.p), static_cast<const s16 *>(h->misc1.p), n, static_cast<u32 *>(h->misc2.p), h->stream));
    ++h->launches;
    D2H(h, dis, h->misc2.p, (size_t)n * 4);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// test hook: number of float bit patterns in [lo_bits, hi_bits) for which the branch-free sqrt differs from
// the IEEE intrinsic (must be 0 over [1.0f, 2^33) = the range the kernels feed it)
int sr_debug_sqrt_mismatches(sr_handle *h, uint32_t lo_bits, uint32_t hi_bits, uint64_t *mismatches) {
    SR_REQUIRE(h, h && mismatches);
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->misc2, 16));
    SR_CK(h, cudaMemsetAsync(h->misc2.p, 0, 8, h->stream));
    SR_CK(h, launch_sqrt_check(lo_bits, hi_bits, static_cast<unsigned long long *>(h->misc2.p), h->stream));
    D2H(h, mismatches, h->misc2.p, 8);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- (1) the reference's own entry points: batch-of-1 on a lazily created default handle -----------
static std::mutex g_default_mu;
static sr_handle *g_default = nullptr;
static sr_handle *default_handle() {
    if (!g_default) {
        sr_handle *h = nullptr;
        if (sr_create(0, &h) == 0) g_default = h;
    }
    return g_default;
}

// VAD.H:24 / VAD.C:22-71. On failure (no device) *atap is left untouched and sr_last_error(NULL) is set.
void noise_atap(const uint16_t *noise, uint16_t n_len, atap_tag *atap) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    sr_handle *h = default_handle();
    if (!h || !noise || !atap) return;
    if (n_len == 0) return;
    sr_noise_atap_batch(h, noise, n_len, 1, n_len, atap);
}

// VAD.H:25 / VAD.C:97-218. Segments come back as pointers into the caller's buffer.
void VAD(const uint16_t *vc, uint16_t buf_len, valid_tag *valid_voice, atap_tag *atap_arg) {
    if (valid_voice) for (unsigned i = 0; i < SR_MAX_VC_CON; ++i) { valid_voice[i].start = nullptr; valid_voice[i].end = nullptr; }   // VAD.C:115-119
    std::lock_guard<std::mutex> lk(g_default_mu);
    sr_handle *h = default_handle();
    if (!h || !vc || !valid_voice || !atap_arg || buf_len == 0) return;
    uint32_t seg[6];
    if (sr_vad_batch(h, vc, buf_len, 1, buf_len, atap_arg, seg) != 0) return;
    for (unsigned i = 0; i < SR_MAX_VC_CON; ++i) {
        valid_voice[i].start = seg[2 * i] == SR_SEG_NULL ? nullptr : const_cast<uint16_t *>(vc) + seg[2 * i];
        valid_voice[i].end = seg[2 * i + 1] == SR_SEG_NULL ? nullptr : const_cast<uint16_t *>(vc) + seg[2 * i + 1];
    }
}

// MFCC.H:27 / MFCC.C:86-191. Like the reference this reads valid->start[-1] (MFCC.C:119, i=0).
void get_mfcc(valid_tag *valid, v_ftr_tag *v_ftr, atap_tag *atap_arg) {
    if (!v_ftr) return;
    std::lock_guard<std::mutex> lk(g_default_mu);
    sr_handle *h = default_handle();
    if (!h || !valid || !atap_arg || !valid->start || !valid->end || valid->end < valid->start) { v_ftr->frm_num = 0; return; }
    size_t len = (size_t)(valid->end - valid->start);
    // more than vv_frm_max frames is rejected by the kernel (MFCC.C:103-107); cap what is shipped to the device
    const size_t cap = 120 * 80 + 80;
    if (len > cap) len = cap;
    const uint32_t U = (uint32_t)len + 1;
    const uint32_t seg[2] = {1u, U};
    if (sr_mfcc_batch(h, valid->start - 1, U, 1, seg, 2, atap_arg, v_ftr) != 0) v_ftr->frm_num = 0;
}

// the reference keeps in_frm_num / mdl_frm_num of the last dtw() call in file statics (DTW.C:65-68, set at :130-131);
// dtw_limit() reads them. Here they are per calling thread.
static thread_local uint16_t g_last_I = 0, g_last_M = 0;

// DTW.C:76-109 (global, no header): 0 = "ins", 1 = "outs", for the (I, M) of this thread's last dtw() call
uint8_t dtw_limit(uint16_t x, uint16_t y) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    sr_handle *h = default_handle();
    uint8_t r = 1;
    if (!h) return r;
    if (sr_dtw_limit_batch(h, &x, &y, &g_last_I, &g_last_M, 1, &r) != 0) return 1;
    return r;
}

// DTW.H:7 / DTW.C:120-192
uint32_t dtw(v_ftr_tag *ftr_in, v_ftr_tag *frt_mdl) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    sr_handle *h = default_handle();
    if (!h || !ftr_in || !frt_mdl) return SR_DIS_ERR;
    g_last_I = ftr_in->frm_num; g_last_M = frt_mdl->frm_num;                  // DTW.C:130-131
    const void *sv_bank = h->bank; const u32 sv_n = h->n_slot, sv_s = h->slot_stride;
    const u32 *sv_perm = h->perm;
    uint32_t score = SR_DIS_ERR;
    DeviceGuard g(h->device);
    if (ensure(h->misc2, kFtrBytes + 16) != cudaSuccess) return SR_DIS_ERR;
    if (cudaMemcpyAsync(h->misc2.p, frt_mdl, kFtrBytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return SR_DIS_ERR;
    h->bank = h->misc2.p; h->n_slot = 1; h->slot_stride = kFtrBytes;
    h->perm = nullptr;                                              // the slot order belongs to the handle's real bank
    const int rc = sr_dtw_batch(h, ftr_in, 1, 0, 0, &score, nullptr, nullptr);
    h->bank = sv_bank; h->n_slot = sv_n; h->slot_stride = sv_s; h->perm = sv_perm;
    return rc == 0 ? score : SR_DIS_ERR;
}

// MFCC.C:27-62: returns a pointer to a buffer owned by the library (thread-local instead of the
// reference's single static): [0,512) magnitudes, [512,1024) the raw packed FFT bins like fft_out.
uint32_t *fft(int16_t *dat_buf, uint16_t buf_len) {
    static thread_local uint32_t out[SR_FFT_POINT];
    if (buf_len > SR_FFT_POINT || !dat_buf) return nullptr;          // MFCC.C:32-35
    std::lock_guard<std::mutex> lk(g_default_mu);
    sr_handle *h = default_handle();
    if (!h) return nullptr;
    DeviceGuard g(h->device);
    if (ensure(h->misc0, 4096) != cudaSuccess || ensure(h->misc1, 4096) != cudaSuccess || ensure(h->misc2, 2048) != cudaSuccess) return nullptr;
    if (buf_len && cudaMemcpyAsync(h->misc0.p, dat_buf, (size_t)buf_len * 2, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return nullptr;
    if (launch_fft_generic(nullptr, static_cast<const s16 *>(h->misc0.p), buf_len, 1, static_cast<u32 *>(h->misc1.p),
                           static_cast<u32 *>(h->misc2.p), h->stream) != cudaSuccess) return nullptr;
    ++h->launches;
    if (cudaMemcpyAsync(out, h->misc2.p, 2048, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return nullptr;
    if (cudaMemcpyAsync(out + 512, static_cast<u32 *>(h->misc1.p) + 512, 2048, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return nullptr;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return nullptr;
    return out;
}

// DTW.C:45-62
uint32_t get_dis(int16_t *frm_ftr1, int16_t *frm_ftr2) {
    std::lock_guard<std::mutex> lk(g_default_mu);
    sr_handle *h = default_handle();
    if (!h || !frm_ftr1 || !frm_ftr2) return SR_DIS_ERR;
    uint32_t d = SR_DIS_ERR;
    if (sr_get_dis_batch(h, frm_ftr1, frm_ftr2, 1, &d) != 0) return SR_DIS_ERR;
    return d;
}

}  // extern "C"
