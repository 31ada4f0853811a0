// sr_vad.cu -- K0: batched noise_atap (Src/Speech_Recog/VAD.C:22-71) and VAD (VAD.C:97-218).
//
// One warp per utterance. The utterance's PCM is staged once into shared memory by a 1-D bulk
// async copy (TMA engine) and both functions run on the staged copy, so HBM traffic is the
// algorithmic 2*U bytes per utterance.
//   * noise_atap: lane-strided sums / maxima + warp reductions;
//   * VAD features: one lane per 20 ms frame. frm_sum is order independent; the band-crossing
//     count depends on `last_sig`, which the reference never resets between frames (VAD.C:99), so a
//     frame's result depends on the class of the last out-of-band sample at index <= i_k+78.
//     Only ONE pair per frame can see that carried-in state (the pair ending at the frame's first
//     out-of-band sample), so each lane counts crossings with an "unknown" initial state, records
//     the class/position of its first out-of-band sample and block summaries, and the carried state
//     is applied afterwards in a short serial pass that also runs the 4-state endpoint FSM
//     (VAD.C:164-216).
#include "sr_common.cuh"

namespace srk {

constexpr int kVadMaxWarps = 12;

struct VadWarpView {
    const u16 *x;       // staged samples, x[0] = first sample of the utterance
    bool vec_ok;        // 16-byte aligned -> uint4 shared loads
};

__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ u32 warp_max(u32 v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
    return v;
}

// per-frame record produced by the parallel pass
//   bits 0..7 zc0 (crossings with unknown initial state), 8..9 class of first out-of-band sample,
//   bit 10 that sample is not at position 0, 11..12 last class in [0,78], 13..14 last class in [0,79],
//   bit 15 frm_sum > s_thl
__device__ __forceinline__ u32 frame_scan(const VadWarpView &v, u32 i0, u32 mid, u32 a_thl, u32 b_thl, u32 s_thl) {
    u32 frm_sum = 0, zc = 0, last = 0, first_cls = 0, first_pos1 = 0, la = 0, lf = 0;
    const u16 *p = v.x + i0;
#pragma unroll 1
    for (int c = 0; c < 20; ++c) {
        u32 w[4];
        if (v.vec_ok) {
            const uint4 q = *reinterpret_cast<const uint4 *>(p + 8 * c);
            w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = (u32)p[8 * c + 2 * j] | ((u32)p[8 * c + 2 * j + 1] << 16);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int h = 8 * c + j;
            const u32 s = (j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xFFFFu);
            frm_sum += s > mid ? s - mid : mid - s;                       // VAD.C:126-129
            const u32 cls = s >= a_thl ? 2u : (s < b_thl ? 1u : 0u);      // VAD.C:134-141 / 143-156
            // test of sample h against the state left by samples < h (h >= 1), then update from sample h
            if (h >= 1) {
                zc += (cls != 0 && last != 0 && cls != last) ? 1u : 0u;
                if (cls != 0 && first_cls == 0) { first_cls = cls; first_pos1 = 1; }
            } else if (cls != 0) {
                first_cls = cls; first_pos1 = 0;
            }
            // NOTE: the reference updates from sample h only for h <= 158; sample 159 is tested, never
            // "updated from" inside this frame -- irrelevant here because nothing is tested after it.
            if (cls != 0) last = cls;
            if (h == 78) la = last;
            if (h == 79) lf = last;
        }
    }
    return (zc & 0xFFu) | (first_cls << 8) | (first_pos1 << 10) | (la << 11) | (lf << 13) |
           ((frm_sum > s_thl ? 1u : 0u) << 15);
}

__global__ void __launch_bounds__(kVadMaxWarps * 32)
vad_kernel(const u16 *__restrict__ pcm, u32 U, u32 B, u32 n_len, u32 buf_len, int do_atap, int do_vad,
           atap_tag *__restrict__ atap, u32 *__restrict__ seg_off, u32 buf_bytes, u32 max_frames) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ u64 bars[kVadMaxWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    unsigned char *buf = smem_raw + (size_t)warp * buf_bytes;
    u32 *info = reinterpret_cast<u32 *>(smem_raw + (size_t)nwarps * buf_bytes) + (size_t)warp * max_frames;
    if (lane == 0) mbar_init(&bars[warp], 1);
    if (threadIdx.x == 0) mbar_fence_init();
    __syncthreads();

    const size_t total_bytes = (size_t)B * U * 2;
    const bool base_aligned = (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;
    u32 need = 0;                                                  // samples this call touches per utterance
    if (do_atap) need = n_len;
    if (do_vad && buf_len > need) need = buf_len;
    if (need > U) need = U;

    int it = 0;
    for (u32 b = blockIdx.x * nwarps + warp; b < B; b += gridDim.x * nwarps, ++it) {
        // ---- stage samples [b*U, b*U+need) ----------------------------------------------------
        const size_t lo = (size_t)b * U * 2, hi = lo + (size_t)need * 2;
        int shift = 0;
        if (base_aligned && need > 0) {
            const size_t lo_al = lo & ~(size_t)15;
            size_t hi_al = (hi + 15) & ~(size_t)15;
            const size_t lim = total_bytes & ~(size_t)15;
            if (hi_al > lim) hi_al = lim;
            shift = (int)((lo - lo_al) >> 1);
            if (hi > hi_al) {
                const u16 *g = reinterpret_cast<const u16 *>(reinterpret_cast<const unsigned char *>(pcm) + hi_al);
                u16 *d = reinterpret_cast<u16 *>(buf + (hi_al - lo_al));
                const int n = (int)((hi - hi_al) >> 1);
                if (lane < n) d[lane] = g[lane];
            }
            __syncwarp();
            if (lane == 0) {
                const u32 nbytes = (u32)(hi_al - lo_al);
                mbar_arrive_expect_tx(&bars[warp], nbytes);
                bulk_g2s(buf, reinterpret_cast<const unsigned char *>(pcm) + lo_al, nbytes, &bars[warp]);
            }
            mbar_wait(&bars[warp], it & 1);
        } else {
            const u16 *g = pcm + (size_t)b * U;
            u16 *d = reinterpret_cast<u16 *>(buf);
            for (u32 i = lane; i < need; i += 32) d[i] = g[i];
            __syncwarp();
        }
        VadWarpView v;
        v.x = reinterpret_cast<const u16 *>(buf) + shift;
        v.vec_ok = (shift & 7) == 0;

        // ---- noise_atap, VAD.C:22-71 ------------------------------------------------------------
        atap_tag at = atap[b];
        if (do_atap && n_len != 0 && (n_len % 240u) == 0 && n_len <= U) {     // VAD.C:33-36: else untouched
            u32 s = 0;
            for (u32 i = lane; i < n_len; i += 32) s += v.x[i];
            const u32 mid = warp_sum(s) / n_len;                               // VAD.C:41-45
            u32 max_sum = 0, abs_sum = 0;
            for (u32 i = 0; i < n_len; i += 240u) {                            // VAD.C:48-63
                u32 mx = 0, sm = 0;
                for (u32 h = lane; h < 240u; h += 32) {
                    const u32 x = v.x[i + h], a = x > mid ? x - mid : mid - x;
                    mx = max(mx, a); sm += a;
                }
                max_sum += warp_max(mx);
                abs_sum += sm;
            }
            abs_sum = warp_sum(abs_sum);
            abs_sum /= (n_len / SR_FRAME_LEN);                                 // VAD.C:65
            max_sum /= (n_len / 240u);                                         // VAD.C:66
            at.mid_val = mid;
            at.n_thl = (u16)max_sum;                                           // n_thl_ratio 1, VAD.C:68
            at.s_thl = abs_sum * 11u / 10u;                                    // s_thl_ratio 11/10, VAD.C:69
            at.z_thl = 2;                                                      // 160*2/160/1, VAD.C:70
            if (lane == 0) atap[b] = at;
        }

        // ---- VAD, VAD.C:97-218 ------------------------------------------------------------------
        if (do_vad) {
            const u32 mid = at.mid_val;
            const u32 a_thl = mid + at.n_thl, b_thl = mid - at.n_thl;          // VAD.C:112-113 (u32 wrap)
            // frames i = 0,80,.. while i < buf_len-160 (VAD.C:121); buf_len <= 160 reads past the buffer in
            // the reference (int -> u32 compare) -- here: no frames.
            u32 nfr = buf_len > SR_FRAME_LEN ? (buf_len - SR_FRAME_LEN + SR_FRAME_MOV - 1) / SR_FRAME_MOV : 0;
            if (buf_len > U) nfr = 0;
            for (u32 k = lane; k < nfr; k += 32) info[k] = frame_scan(v, 80u * k, mid, a_thl, b_thl, at.s_thl);
            __syncwarp();
            u32 seg[6] = {SR_SEG_NULL, SR_SEG_NULL, SR_SEG_NULL, SR_SEG_NULL, SR_SEG_NULL, SR_SEG_NULL};
            u32 carry = 0, cur = 0, front = 0, back = 0, valid_con = 0;
            for (u32 k = 0; k < nfr; ++k) {
                const u32 r = info[k];
                const u32 la = (r >> 11) & 3u, lf = (r >> 13) & 3u, fc = (r >> 8) & 3u;
                const u32 init = k == 0 ? 0u : (la ? la : carry);              // class of last out-of-band sample <= i+78
                u32 zc = r & 0xFFu;
                if (((r >> 10) & 1u) && init != 0 && init != fc) ++zc;
                if (lf) carry = lf;
                const bool active = ((r >> 15) & 1u) || zc > at.z_thl;         // VAD.C:164
                const u32 i = 80u * k;
                if (active) {
                    if (cur == 0) { cur = 1; front = 1; }
                    else if (cur == 1) { if (++front >= 8) { cur = 2; seg[2 * valid_con] = i - 7 * 80; front = 0; } }
                    else if (cur == 3) { back = 0; cur = 2; }
                } else {
                    if (cur == 2) { cur = 3; back = 1; }
                    else if (cur == 3) {
                        if (++back >= 11) {
                            cur = 0;
                            seg[2 * valid_con + 1] = i - 11 * 80 + 160;
                            if (++valid_con == SR_MAX_VC_CON) break;
                            back = 0;
                        }
                    } else if (cur == 1) { front = 0; cur = 0; }
                }
            }
            if (lane < 6) {
                u32 val = seg[0];
#pragma unroll
                for (int j = 1; j < 6; ++j) if (lane == j) val = seg[j];
                seg_off[(size_t)b * 6 + lane] = val;
            }
        }
        __syncwarp();
    }
}

cudaError_t launch_vad(const u16 *pcm, u32 U, u32 B, u32 n_len, u32 buf_len, int do_atap, int do_vad,
                       atap_tag *atap, u32 *seg_off, int num_sms, cudaStream_t st) {
    if (B == 0) return cudaSuccess;
    u32 need = 0;
    if (do_atap) need = n_len;
    if (do_vad && buf_len > need) need = buf_len;
    if (need > U) need = U;
    const u32 buf_bytes = ((need * 2 + 32 + 127) / 128) * 128;
    const u32 max_frames = (buf_len > 160 ? (buf_len - 160 + 79) / 80 : 0) + 1;
    const size_t per_warp = (size_t)buf_bytes + (size_t)max_frames * 4;
    int warps = (int)((220 * 1024) / per_warp);
    if (warps < 1) return cudaErrorInvalidValue;
    if (warps > kVadMaxWarps) warps = kVadMaxWarps;
    const size_t smem = per_warp * warps;
    cudaError_t e = cudaFuncSetAttribute(vad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);   // + 96 B static barriers <= 227 KB
    if (e != cudaSuccess) return e;
    u32 grid = (B + warps - 1) / warps;
    const u32 cap = (u32)num_sms * 4;
    if (grid > cap) grid = cap;
    vad_kernel<<<grid, warps * 32, smem, st>>>(pcm, U, B, n_len, buf_len, do_atap, do_vad, atap, seg_off, buf_bytes,
                                              max_frames);
    return cudaGetLastError();
}

}  // namespace srk
