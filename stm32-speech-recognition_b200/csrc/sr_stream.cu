// sr_stream.cu -- K4: streaming front end standing in for the reference's blocking capture loop
// (record(), Src/APP/main.c:77-102 + ADC_DMA_Init, Src/BSP/ADC.C:11-103): S concurrent audio streams are
// fed in lock-step chunks; every pushed chunk advances, per stream, exactly the computation the reference
// would do on the finished buffer -- noise_atap once the 300 ms calibration window is complete
// (main.c:258), then VAD frame by frame with the reference's own carried state (`last_sig`, FSM counters,
// VAD.C:97-218) -- and every segment the FSM closes is recognised at once (get_mfcc + dtw + argmin,
// main.c:268-294) with the batch kernels through a row map. After the last chunk the union of the events
// equals the batch result on the complete buffer (segments of sr_vad_batch; segment 0 = sr_recognise_batch).
// The reference only ever recognises segment 0 (main.c:268); here all <= 3 segments are.
//
// VAD here is the literal sequential algorithm, one thread per stream: a push adds only a frame or a few,
// so there is nothing to parallelise inside a stream and the carried state makes it naturally incremental.
#include "sr_internal.h"

namespace srk {

struct StreamState {            // one per stream, device resident
    atap_tag atap;
    u32 frames_done;            // VAD frames already evaluated
    u32 last_sig, cur, front, back, valid_con;
    u32 seg[6];
    u32 calibrated;
};

struct StreamEventDev {         // compact work list of segments closed by the current push
    u32 stream, segment, start, end;
};

__global__ void stream_reset_kernel(StreamState *st, u32 S) {
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    StreamState z;
    z.atap.mid_val = 0; z.atap.n_thl = 0; z.atap.z_thl = 0; z.atap.s_thl = 0;
    z.frames_done = 0; z.last_sig = 0; z.cur = 0; z.front = 0; z.back = 0; z.valid_con = 0; z.calibrated = 0;
    for (int i = 0; i < 6; ++i) z.seg[i] = SR_SEG_NULL;
    st[s] = z;
}

// one thread per stream: calibrate when possible, then evaluate every frame that became complete
__global__ void stream_vad_step_kernel(const u16 *__restrict__ pcm, u32 L /* row length = final buffer length */,
                                       u32 S, u32 n /* samples received so far */, u32 n_len,
                                       StreamState *__restrict__ state, StreamEventDev *__restrict__ ev,
                                       u32 *__restrict__ seg_ev /*[cap][2]*/, atap_tag *__restrict__ atap_ev,
                                       u32 *__restrict__ map_ev, u32 *__restrict__ n_ev, u32 cap) {
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    StreamState st = state[s];
    const u16 *x = pcm + (size_t)s * L;
    if (!st.calibrated) {
        if (n < n_len) return;
        st.calibrated = 1;
        if (n_len != 0 && n_len % 240u == 0) {                       // noise_atap, VAD.C:22-71 (else atap untouched)
            u32 sum = 0;
            for (u32 i = 0; i < n_len; ++i) sum += x[i];
            const u32 mid = sum / n_len;
            u32 max_sum = 0, abs_sum = 0;
            for (u32 i = 0; i < n_len; i += 240u) {
                u32 mx = 0;
                for (u32 h = 0; h < 240u; ++h) {
                    const u32 v = x[i + h], a = v > mid ? v - mid : mid - v;
                    mx = max(mx, a); abs_sum += a;
                }
                max_sum += mx;
            }
            abs_sum /= (n_len / SR_FRAME_LEN);
            max_sum /= (n_len / 240u);
            st.atap.mid_val = mid; st.atap.n_thl = (u16)max_sum; st.atap.s_thl = abs_sum * 11u / 10u; st.atap.z_thl = 2;
        }
    }
    const u32 mid = st.atap.mid_val, a_thl = mid + st.atap.n_thl, b_thl = mid - st.atap.n_thl;
    // frames i = 80k while i < L-160 (VAD.C:121), as soon as samples [i, i+160) have arrived
    const u32 nfr_total = L > SR_FRAME_LEN ? (L - SR_FRAME_LEN + SR_FRAME_MOV - 1) / SR_FRAME_MOV : 0;
    u32 k = st.frames_done;
    while (k < nfr_total && 80u * k + 160u <= n && st.valid_con < SR_MAX_VC_CON) {
        const u32 i = 80u * k;
        u32 frm_sum = 0, frm_zero = 0, last_sig = st.last_sig;
        for (u32 h = 0; h < SR_FRAME_LEN; ++h) {                     // VAD.C:126-129
            const u32 v = x[i + h];
            frm_sum += v > mid ? v - mid : mid - v;
        }
        for (u32 h = 0; h < SR_FRAME_LEN - 1; ++h) {                 // VAD.C:132-157
            const u32 v = x[i + h], w = x[i + h + 1];
            if (v >= a_thl) last_sig = 2; else if (v < b_thl) last_sig = 1;
            if (w >= a_thl) { if (last_sig == 1) ++frm_zero; }
            else if (w < b_thl) { if (last_sig == 2) ++frm_zero; }
        }
        st.last_sig = last_sig;
        if (frm_sum > st.atap.s_thl || frm_zero > st.atap.z_thl) {   // VAD.C:164-187
            if (st.cur == 0) { st.cur = 1; st.front = 1; }
            else if (st.cur == 1) { if (++st.front >= 8) { st.cur = 2; st.seg[2 * st.valid_con] = i - 7 * 80; st.front = 0; } }
            else if (st.cur == 3) { st.back = 0; st.cur = 2; }
        } else {                                                     // VAD.C:188-216
            if (st.cur == 2) { st.cur = 3; st.back = 1; }
            else if (st.cur == 3) {
                if (++st.back >= 11) {
                    st.cur = 0;
                    const u32 sgi = st.valid_con;
                    st.seg[2 * sgi + 1] = i - 11 * 80 + 160;
                    ++st.valid_con;
                    st.back = 0;
                    const u32 e = atomicAdd(n_ev, 1u);               // segment closed: queue it for recognition
                    if (e < cap) {
                        StreamEventDev d; d.stream = s; d.segment = sgi; d.start = st.seg[2 * sgi]; d.end = st.seg[2 * sgi + 1];
                        ev[e] = d;
                        seg_ev[2 * e] = d.start; seg_ev[2 * e + 1] = d.end;
                        atap_ev[e] = st.atap; map_ev[e] = s;
                    }
                }
            } else if (st.cur == 1) { st.front = 0; st.cur = 0; }
        }
        ++k;
    }
    st.frames_done = k;
    state[s] = st;
}

__global__ void stream_segments_kernel(const StreamState *st, u32 S, u32 *seg_off, atap_tag *atap) {
    const u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (seg_off) for (int i = 0; i < 6; ++i) seg_off[(size_t)s * 6 + i] = st[s].seg[i];
    if (atap) atap[s] = st[s].atap;
}

// status per event from the freshly computed features (MFCC fail = frm_num 0, main.c:269-274)
__global__ void stream_status_kernel(const unsigned char *ftr, u32 n, u8 *status, u32 *frm) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 f = (*reinterpret_cast<const u32 *>(ftr + (size_t)i * kFtrBytes)) >> 16;
    status[i] = f == 0 ? SR_ST_MFCC_FAIL : SR_ST_OK;
    frm[i] = f;
}

}  // namespace srk

struct sr_stream_pool {
    sr_handle *h = nullptr;
    u32 S = 0, L = 0, n_len = 0, n = 0, cap = 0;
    DevBuf pcm, state, ev, seg_ev, atap_ev, map_ev, n_ev, ftr, status, frm, bidx, bdis, cmd;
    u32 *n_ev_host = nullptr;                      // pinned
    unsigned char *ev_host = nullptr;              // pinned staging of the event records
    size_t ev_host_bytes = 0;
};

extern "C" {

int sr_streams_destroy(sr_stream_pool *p) {
    if (!p) return 0;
    DeviceGuard g(p->h->device);
    cudaStreamSynchronize(p->h->stream);
    DevBuf *bufs[] = {&p->pcm, &p->state, &p->ev, &p->seg_ev, &p->atap_ev, &p->map_ev, &p->n_ev, &p->ftr,
                      &p->status, &p->frm, &p->bidx, &p->bdis, &p->cmd};
    for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
    if (p->n_ev_host) cudaFreeHost(p->n_ev_host);
    if (p->ev_host) cudaFreeHost(p->ev_host);
    delete p;
    return 0;
}

int sr_streams_reset(sr_stream_pool *p) {
    SR_REQUIRE(nullptr, p != nullptr);
    sr_handle *h = p->h;
    DeviceGuard g(h->device);
    p->n = 0;
    stream_reset_kernel<<<(p->S + 127) / 128, 128, 0, h->stream>>>(static_cast<StreamState *>(p->state.p), p->S);
    SR_CK(h, cudaGetLastError());
    SR_CK(h, cudaMemsetAsync(p->pcm.p, 0, (size_t)p->S * p->L * 2, h->stream));
    ++h->launches;
    return 0;
}

int sr_streams_create(sr_handle *h, uint32_t n_streams, uint32_t max_samples, uint32_t n_len, sr_stream_pool **out) {
    SR_REQUIRE(h, h && out && n_streams > 0 && max_samples > 0 && max_samples <= 65535u && n_len <= max_samples);
    DeviceGuard g(h->device);
    sr_stream_pool *p = new (std::nothrow) sr_stream_pool;
    SR_REQUIRE(h, p != nullptr);
    p->h = h; p->S = n_streams; p->L = max_samples; p->n_len = n_len; p->cap = 3 * n_streams;
    const size_t cap = p->cap;
    cudaError_t e = cudaSuccess;
    auto need = [&](DevBuf &b, size_t bytes) { if (e == cudaSuccess) e = ensure(b, bytes); };
    need(p->pcm, (size_t)n_streams * max_samples * 2 + 64);
    need(p->state, (size_t)n_streams * sizeof(StreamState));
    need(p->ev, cap * sizeof(StreamEventDev));
    need(p->seg_ev, cap * 8);
    need(p->atap_ev, cap * sizeof(atap_tag));
    need(p->map_ev, cap * 4);
    need(p->n_ev, 16);
    need(p->ftr, cap * kFtrBytes);
    need(p->status, cap);
    need(p->frm, cap * 4);
    need(p->bidx, cap * 4);
    need(p->bdis, cap * 4);
    need(p->cmd, cap * 4);
    p->ev_host_bytes = cap * (sizeof(StreamEventDev) + 4 * 4 + 1) + 64;
    if (e == cudaSuccess) e = cudaMallocHost(&p->n_ev_host, 16);
    if (e == cudaSuccess) e = cudaMallocHost(&p->ev_host, p->ev_host_bytes);
    if (e != cudaSuccess) { sr_streams_destroy(p); return fail(h, "sr_streams_create: allocation", e); }
    *out = p;
    return sr_streams_reset(p);
}

// Append chunk_len samples to every stream (chunk[s*chunk_stride + i], host memory; pinned for best latency),
// advance VAD, recognise every segment that closed. Returns after the results are on the host.
int sr_streams_push(sr_stream_pool *p, const uint16_t *chunk, uint32_t chunk_len, uint32_t chunk_stride,
                    sr_stream_event *events, uint32_t max_events, uint32_t *n_events) {
    SR_REQUIRE(nullptr, p && n_events);
    sr_handle *h = p->h;
    SR_REQUIRE(h, chunk_len == 0 || chunk != nullptr);
    SR_REQUIRE(h, p->n + chunk_len <= p->L && chunk_stride >= chunk_len);
    DeviceGuard g(h->device);
    *n_events = 0;
    if (chunk_len) {
        SR_CK(h, cudaMemcpy2DAsync(static_cast<u16 *>(p->pcm.p) + p->n, (size_t)p->L * 2, chunk, (size_t)chunk_stride * 2,
                                   (size_t)chunk_len * 2, p->S, cudaMemcpyHostToDevice, h->stream));
        p->n += chunk_len;
    }
    SR_CK(h, cudaMemsetAsync(p->n_ev.p, 0, 4, h->stream));
    stream_vad_step_kernel<<<(p->S + 63) / 64, 64, 0, h->stream>>>(
        static_cast<const u16 *>(p->pcm.p), p->L, p->S, p->n, p->n_len, static_cast<StreamState *>(p->state.p),
        static_cast<StreamEventDev *>(p->ev.p), static_cast<u32 *>(p->seg_ev.p), static_cast<atap_tag *>(p->atap_ev.p),
        static_cast<u32 *>(p->map_ev.p), static_cast<u32 *>(p->n_ev.p), p->cap);
    SR_CK(h, cudaGetLastError());
    ++h->launches;
    SR_CK(h, cudaMemcpyAsync(p->n_ev_host, p->n_ev.p, 4, cudaMemcpyDeviceToHost, h->stream));
    SR_CK(h, cudaStreamSynchronize(h->stream));
    u32 ne = *p->n_ev_host;
    if (ne > p->cap) ne = p->cap;
    if (ne == 0) return 0;
    // recognise the closed segments: get_mfcc through the row map, then dtw + argmin against the bank
    SR_CK(h, launch_mfcc(static_cast<const u16 *>(p->pcm.p), p->L, ne, static_cast<const u32 *>(p->seg_ev.p), 2,
                         static_cast<const atap_tag *>(p->atap_ev.p), p->ftr.p, h->num_sms, h->stream,
                         static_cast<const u32 *>(p->map_ev.p), p->S));
    stream_status_kernel<<<(ne + 127) / 128, 128, 0, h->stream>>>(static_cast<const unsigned char *>(p->ftr.p), ne,
                                                                 static_cast<u8 *>(p->status.p), static_cast<u32 *>(p->frm.p));
    SR_CK(h, cudaGetLastError());
    h->launches += 2;
    SR_CK(h, ensure(h->best, (size_t)ne * 8));
    u64 *best = static_cast<u64 *>(h->best.p);
    SR_CK(h, launch_best_init(best, ne, h->stream));
    if (h->n_slot)
        SR_CK(h, launch_dtw(p->ftr.p, ne, h->bank, h->n_slot, h->slot_stride, SR_DTW_CHECK_SIGN, nullptr, best,
                            static_cast<const u8 *>(p->status.p), h->num_sms, h->stream));
    SR_CK(h, launch_best_final(best, ne, static_cast<u32 *>(p->bidx.p), static_cast<u32 *>(p->bdis.p),
                               static_cast<u32 *>(p->cmd.p), static_cast<const u8 *>(p->status.p), h->stream));
    h->launches += 3;
    unsigned char *hp = p->ev_host;
    StreamEventDev *hev = reinterpret_cast<StreamEventDev *>(hp);
    u32 *hfrm = reinterpret_cast<u32 *>(hp + p->cap * sizeof(StreamEventDev));
    u32 *hidx = hfrm + p->cap, *hdis = hidx + p->cap, *hcmd = hdis + p->cap;
    u8 *hst = reinterpret_cast<u8 *>(hcmd + p->cap);
    D2H(h, hev, p->ev.p, (size_t)ne * sizeof(StreamEventDev));
    D2H(h, hfrm, p->frm.p, (size_t)ne * 4);
    D2H(h, hidx, p->bidx.p, (size_t)ne * 4);
    D2H(h, hdis, p->bdis.p, (size_t)ne * 4);
    D2H(h, hcmd, p->cmd.p, (size_t)ne * 4);
    D2H(h, hst, p->status.p, (size_t)ne);
    SR_CK(h, cudaStreamSynchronize(h->stream));
    const u32 nout = ne < max_events ? ne : max_events;
    for (u32 i = 0; i < nout && events; ++i) {
        sr_stream_event &o = events[i];
        o.stream = hev[i].stream; o.segment = hev[i].segment; o.start = hev[i].start; o.end = hev[i].end;
        o.status = hst[i]; o.frm_num = hfrm[i]; o.best_idx = hidx[i]; o.best_dis = hdis[i]; o.cmd = hcmd[i];
    }
    *n_events = ne;
    return 0;
}

// segments (and atap) found so far: same layout as sr_vad_batch's output; host pointers, may be NULL
int sr_streams_segments(sr_stream_pool *p, uint32_t *seg_off, atap_tag *atap) {
    SR_REQUIRE(nullptr, p != nullptr);
    sr_handle *h = p->h;
    DeviceGuard g(h->device);
    SR_CK(h, ensure(h->seg, (size_t)p->S * 24));
    SR_CK(h, ensure(h->atap, (size_t)p->S * sizeof(atap_tag)));
    stream_segments_kernel<<<(p->S + 127) / 128, 128, 0, h->stream>>>(static_cast<const StreamState *>(p->state.p), p->S,
                                                                     static_cast<u32 *>(h->seg.p), static_cast<atap_tag *>(h->atap.p));
    SR_CK(h, cudaGetLastError());
    ++h->launches;
    if (seg_off) D2H(h, seg_off, h->seg.p, (size_t)p->S * 24);
    if (atap) D2H(h, atap, h->atap.p, (size_t)p->S * sizeof(atap_tag));
    SR_CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

}  // extern "C"
