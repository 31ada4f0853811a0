// sr_vad.cu -- K0: batched noise_atap (Src/Speech_Recog/VAD.C:22-71) and VAD (VAD.C:97-218).
//
// One warp per utterance. The utterance's PCM is staged once into shared memory by a 1-D bulk
// async copy (TMA engine) and both functions run on the staged copy, so HBM traffic is the
// algorithmic 2*U bytes per utterance.
//   * noise_atap: lane-strided sums / maxima + warp reductions;
//   * VAD features: frames overlap by 50 %, so the scan works on 80-sample BLOCKS (each sample is
//     touched once) and frame k = block k + block k+1. A block summary is a small monoid element:
//     sum |x-mid|, number of class alternations among its out-of-band samples, class of the last
//     out-of-band sample (the first one follows from the parity of the alternations). The band-
//     crossing count of the reference depends on `last_sig`, which is never reset between frames
//     (VAD.C:99): entering frame k it is the class of the last out-of-band sample at index <= i_k+78.
//     Only ONE pair per frame can see that carried-in state (the pair ending at the frame's first
//     out-of-band sample), so it is applied as a +1 correction; the carry itself is a warp scan.
//   * endpoint FSM (VAD.C:164-216): 8 consecutive active frames open a segment at the first of
//     them, 11 consecutive inactive frames close it at the first of those -- evaluated with bit
//     tricks on the per-frame activity bitmap (one 32-frame word per lane) instead of a serial walk.
#include "sr_common.cuh"

namespace srk {

constexpr int kVadMaxWarps = 32;

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

// per-block summary: bs = sum |x-mid| over the 80 samples; flags = zc (bits 0..6, alternations inside the
// block) | lc << 7 (class of last out-of-band sample, 0 none / 1 below / 2 above) | lcA << 9 (same over the
// first 79 samples) | p0 << 11 (sample 0 is out of band)
// Word-level alternation count. H / L = bitmaps (bit i = sample i) of the samples above / below the band
// (disjoint), `last` = class (0 none / 1 below / 2 above) of the last out-of-band sample before bit 0 of this
// word -- 0 means "unknown" and never counts (the frame-level pass applies the carried-in state separately).
// Returns the number of out-of-band samples whose class differs from the previous out-of-band sample's and
// updates `last`. The previous class of every position comes from a segmented forward fill of H over the
// markers N = H|L (5 doubling steps) instead of a per-sample state machine.
__device__ __forceinline__ u32 word_alternations(u32 H, u32 L, u32 &last) {
    const u32 N = H | L;
    u32 P = H, K = ~N;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { P |= K & (P << d); K &= K << d; }
    // P[i]: the most recent marker at or before i is "above"
    const u32 low = N & (0u - N);                       // lowest marker
    const u32 seen = ~(low | (low - 1u));               // positions strictly above the lowest marker (0 if N == 0)
    const u32 in2 = last == 2u ? 0xFFFFFFFFu : 0u, in1 = last == 1u ? 0xFFFFFFFFu : 0u;
    const u32 prevH = ((P << 1) & seen) | (~seen & in2);
    const u32 prevL = (~(P << 1) & seen) | (~seen & in1);
    const u32 alt = (H & prevL) | (L & prevH);
    if (N) last = ((H >> (31 - __clz(N))) & 1u) ? 2u : 1u;
    return __popc(alt);
}

// Per-block summary, built from bitmaps: per sample only 1 extract + 1 VABSDIFF + 2 compares + 2 predicated ORs.
__device__ __forceinline__ void block_scan(const VadWarpView &v, u32 i0, u32 mid, u32 a_thl, u32 b_thl, u32 &bs_out,
                                           u32 &flags_out) {
    u32 bs = 0;
    u32 H[3] = {0, 0, 0}, L[3] = {0, 0, 0};            // samples 0..31, 32..63, 64..79
    const u16 *p = v.x + i0;
#pragma unroll
    for (int c = 0; c < 10; ++c) {
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
            bs = __usad(s, mid, bs);                                        // VAD.C:126-129
            if (s >= a_thl) H[h >> 5] |= 1u << (h & 31);                    // VAD.C:134-141 / 143-156
            if (s < b_thl) L[h >> 5] |= 1u << (h & 31);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) L[k] &= ~H[k];                              // ">= a" is tested first, "< b" only else
    u32 last = 0, zc = 0;
    zc += word_alternations(H[0], L[0], last);
    zc += word_alternations(H[1], L[1], last);
    u32 lastA = last;                                                       // state after sample 63
    const u32 zc2 = word_alternations(H[2], L[2], last);
    zc += zc2;
    // last class among the first 79 samples: redo the (cheap) "last" update of word 2 without sample 79 (bit 15)
    {
        const u32 Hm = H[2] & 0x7FFFu, Nm = (H[2] | L[2]) & 0x7FFFu;
        if (Nm) lastA = ((Hm >> (31 - __clz(Nm))) & 1u) ? 2u : 1u;
    }
    const u32 p0 = (H[0] | L[0]) & 1u;
    bs_out = bs;
    flags_out = zc | (last << 7) | (lastA << 9) | (p0 << 11);
}

// position of the first set bit at index >= from in a bitmap held one 32-bit word per lane; -1 if none
__device__ __forceinline__ int find_first(u32 word, int lane, int from) {
    const int fw = from >> 5, fb = from & 31;
    u32 m = lane > fw ? word : (lane == fw ? (word & (0xFFFFFFFFu << fb)) : 0u);
    if (from >= 1024) m = 0;
    const u32 bal = __ballot_sync(0xFFFFFFFFu, m != 0);
    if (!bal) return -1;
    const int L = __ffs(bal) - 1;
    const u32 mw = __shfl_sync(0xFFFFFFFFu, m, L);
    return 32 * L + __ffs(mw) - 1;
}
// bitmap shift towards index 0: result[i] = x[i+s], 0 < s < 32
__device__ __forceinline__ u32 bm_shr(u32 x, int s, int lane) {
    u32 nxt = __shfl_down_sync(0xFFFFFFFFu, x, 1);
    if (lane == 31) nxt = 0;
    return (x >> s) | (nxt << (32 - s));
}

// stage samples [first, first+count) of the batch into `buf` (bulk async copy when the batch base is 16-byte
// aligned, plain loads otherwise); returns the sample index of `first` inside buf
__device__ __forceinline__ int stage_chunk(unsigned char *buf, const u16 *pcm, size_t total_bytes, bool base_aligned,
                                           size_t first, u32 count, u64 *bar, u32 &phase, int lane) {
    if (count == 0) return 0;
    const size_t lo = first * 2, hi = lo + (size_t)count * 2;
    if (base_aligned) {
        const size_t lo_al = lo & ~(size_t)15;
        size_t hi_al = (hi + 15) & ~(size_t)15;
        const size_t lim = total_bytes & ~(size_t)15;
        if (hi_al > lim) hi_al = lim;
        const int shift = (int)((lo - lo_al) >> 1);
        if (hi > hi_al) {                                        // tail beyond the last whole 16-byte granule
            const u16 *g = reinterpret_cast<const u16 *>(reinterpret_cast<const unsigned char *>(pcm) + hi_al);
            u16 *d = reinterpret_cast<u16 *>(buf + (hi_al - lo_al));
            const int n = (int)((hi - hi_al) >> 1);
            if (lane < n) d[lane] = g[lane];
        }
        __syncwarp();
        if (lane == 0) {
            const u32 nbytes = (u32)(hi_al - lo_al);
            mbar_arrive_expect_tx(bar, nbytes);
            bulk_g2s(buf, reinterpret_cast<const unsigned char *>(pcm) + lo_al, nbytes, bar);
        }
        mbar_wait(bar, phase & 1u);
        ++phase;
        return shift;
    }
    const u16 *g = pcm + first;
    u16 *d = reinterpret_cast<u16 *>(buf);
    for (u32 i = lane; i < count; i += 32) d[i] = g[i];
    __syncwarp();
    return 0;
}

constexpr u32 kVadChunk = 32 * 80;                               // samples per staged chunk: one 80-sample block per lane

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
    u32 phase = 0;                                                 // completed bulk copies of this warp's barrier

    for (u32 b = blockIdx.x * nwarps + warp; b < B; b += gridDim.x * nwarps) {
        const size_t ubase = (size_t)b * U;
        atap_tag at = atap[b];
        const bool atap_on = do_atap && n_len != 0 && (n_len % 240u) == 0 && n_len <= U;   // VAD.C:33-36: else untouched
        // frames i = 0,80,.. while i < buf_len-160 (VAD.C:121); buf_len <= 160 reads past the buffer in the reference
        // (int -> u32 compare) -- here: no frames.
        u32 nfr = (do_vad && buf_len > SR_FRAME_LEN && buf_len <= U) ? (buf_len - SR_FRAME_LEN + SR_FRAME_MOV - 1) / SR_FRAME_MOV : 0;
        const u32 nblk = nfr ? nfr + 1 : 0;                        // frame k = blocks k, k+1
        const u32 vad_samples = 80u * nblk;                        // <= buf_len

        // ---- noise_atap, VAD.C:22-71: from the first staged chunk when it fits, else straight from global ----------
        const bool atap_staged = atap_on && n_len <= kVadChunk;
        if (atap_on && !atap_staged) {
            const u16 *g = pcm + ubase;
            u32 s = 0;
            for (u32 i = lane; i < n_len; i += 32) s += g[i];
            const u32 mid = warp_sum(s) / n_len;
            u32 max_sum = 0, abs_sum = 0;
            for (u32 i = 0; i < n_len; i += 240u) {
                u32 mx = 0, sm = 0;
                for (u32 h = lane; h < 240u; h += 32) { const u32 x = g[i + h], a = x > mid ? x - mid : mid - x; mx = max(mx, a); sm += a; }
                max_sum += warp_max(mx);
                abs_sum += sm;
            }
            abs_sum = warp_sum(abs_sum) / (n_len / SR_FRAME_LEN);
            max_sum /= (n_len / 240u);
            at.mid_val = mid; at.n_thl = (u16)max_sum; at.s_thl = abs_sum * 11u / 10u; at.z_thl = 2;
            if (lane == 0) atap[b] = at;
        }
        u32 mid = at.mid_val, a_thl = mid + at.n_thl, b_thl = mid - at.n_thl;            // VAD.C:112-113 (u32 wrap)

        // ---- stream the utterance in chunks of 32 blocks: one block per lane, every PCM byte read from HBM once ------
        const u32 total = max(vad_samples, atap_staged ? n_len : 0u);
        for (u32 c0 = 0; c0 < total; c0 += kVadChunk) {
            const u32 cnt = min(kVadChunk, total - c0);
            const int shift = stage_chunk(buf, pcm, total_bytes, base_aligned, ubase + c0, cnt, &bars[warp], phase, lane);
            VadWarpView v;
            v.x = reinterpret_cast<const u16 *>(buf) + shift;
            v.vec_ok = (shift & 7) == 0;
            if (c0 == 0 && atap_staged) {
                u32 s = 0;
                for (u32 i = lane; i < n_len; i += 32) s += v.x[i];
                const u32 m = warp_sum(s) / n_len;                             // VAD.C:41-45
                u32 max_sum = 0, abs_sum = 0;
                for (u32 i = 0; i < n_len; i += 240u) {                        // VAD.C:48-63
                    u32 mx = 0, sm = 0;
                    for (u32 h = lane; h < 240u; h += 32) { const u32 x = v.x[i + h], a = x > m ? x - m : m - x; mx = max(mx, a); sm += a; }
                    max_sum += warp_max(mx);
                    abs_sum += sm;
                }
                abs_sum = warp_sum(abs_sum) / (n_len / SR_FRAME_LEN);          // VAD.C:65
                max_sum /= (n_len / 240u);                                     // VAD.C:66
                at.mid_val = m;
                at.n_thl = (u16)max_sum;                                       // n_thl_ratio 1, VAD.C:68
                at.s_thl = abs_sum * 11u / 10u;                                // s_thl_ratio 11/10, VAD.C:69
                at.z_thl = 2;                                                  // 160*2/160/1, VAD.C:70
                if (lane == 0) atap[b] = at;
                mid = m; a_thl = mid + at.n_thl; b_thl = mid - at.n_thl;
            }
            const u32 blk = c0 / 80u + (u32)lane;
            if (blk < nblk) {
                u32 bs, fl;
                block_scan(v, 80u * (u32)lane, mid, a_thl, b_thl, bs, fl);
                info[2 * blk] = bs; info[2 * blk + 1] = fl;
            }
            __syncwarp();                                                      // buffer is re-staged next iteration
        }

        // ---- VAD, VAD.C:97-218: frames from the block summaries ---------------------------------------------------
        if (do_vad) {
            u32 seg[6] = {SR_SEG_NULL, SR_SEG_NULL, SR_SEG_NULL, SR_SEG_NULL, SR_SEG_NULL, SR_SEG_NULL};
            if (nfr > 0) {
                u32 aw = 0;                                                    // lane j: activity of frames 32j..32j+31
                u32 cin = 0;                                                   // class of last out-of-band sample before this pass
                for (u32 k0 = 0, j = 0; k0 < nfr; k0 += 32, ++j) {
                    const u32 k = k0 + lane;
                    const bool ok = k < nfr;
                    u32 bs0 = 0, f0 = 0, bs1 = 0, f1 = 0;
                    if (ok) { bs0 = info[2 * k]; f0 = info[2 * k + 1]; bs1 = info[2 * k + 2]; f1 = info[2 * k + 3]; }
                    const u32 zc0 = f0 & 127u, lc0 = (f0 >> 7) & 3u, lcA0 = (f0 >> 9) & 3u, p00 = (f0 >> 11) & 1u;
                    const u32 zc1 = f1 & 127u, lc1 = (f1 >> 7) & 3u;
                    const u32 fc0 = lc0 ? ((zc0 & 1u) ? 3u - lc0 : lc0) : 0u;      // first class from last class + parity
                    const u32 fc1 = lc1 ? ((zc1 & 1u) ? 3u - lc1 : lc1) : 0u;
                    // inclusive "last out-of-band class" scan over the blocks of this pass
                    u32 inc = lc0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const u32 up = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                        if (lane >= o && inc == 0) inc = up;
                    }
                    if (inc == 0) inc = cin;
                    u32 prev = __shfl_up_sync(0xFFFFFFFFu, inc, 1);               // carry of block k-1
                    if (lane == 0) prev = cin;
                    cin = __shfl_sync(0xFFFFFFFFu, inc, 31);
                    const u32 init = k == 0 ? 0u : (lcA0 ? lcA0 : prev);          // class of last out-of-band sample <= i+78
                    u32 zc = zc0 + zc1 + ((lc0 && fc1 && lc0 != fc1) ? 1u : 0u);
                    const u32 F = fc0 ? fc0 : fc1;
                    const bool pos1 = fc0 ? (p00 == 0) : (fc1 != 0);              // first out-of-band sample not at position 0
                    if (pos1 && init != 0 && init != F) ++zc;
                    const bool active = ok && ((bs0 + bs1) > at.s_thl || zc > at.z_thl);   // VAD.C:164
                    const u32 word = __ballot_sync(0xFFFFFFFFu, active);
                    if ((u32)lane == j) aw = word;
                }
                // ---- endpoint FSM on the bitmap ------------------------------------------------------
                const u32 fullw = nfr >> 5, rem = nfr & 31u;
                const u32 vmask = (u32)lane < fullw ? 0xFFFFFFFFu : ((u32)lane == fullw ? ((1u << rem) - 1u) : 0u);
                u32 a8 = aw & bm_shr(aw, 1, lane);
                a8 &= bm_shr(a8, 2, lane);
                a8 &= bm_shr(a8, 4, lane);                                     // a8[i]: frames i..i+7 all active
                u32 z = ~aw & vmask;
                u32 z8 = z & bm_shr(z, 1, lane);
                z8 &= bm_shr(z8, 2, lane);
                z8 &= bm_shr(z8, 4, lane);
                const u32 z11 = z8 & bm_shr(z8, 3, lane);                      // z11[i]: frames i..i+10 all inactive
                int cur = 0;
                for (int sgi = 0; sgi < (int)SR_MAX_VC_CON; ++sgi) {
                    const int pfr = find_first(a8, lane, cur);
                    if (pfr < 0) break;
                    seg[2 * sgi] = 80u * (u32)pfr;                             // VAD.C:178: i - 7*80 with i = 80*(pfr+7)
                    const int q = find_first(z11, lane, pfr + 8);
                    if (q < 0) break;                                          // never closes: end stays NULL
                    seg[2 * sgi + 1] = 80u * (u32)q + 80u;                     // VAD.C:201: i - 11*80 + 160 with i = 80*(q+10)
                    cur = q + 11;
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
    const u32 buf_bytes = ((kVadChunk * 2 + 32 + 127) / 128) * 128;   // one 32-block chunk + alignment slack
    const u32 max_frames = 2 * ((buf_len > 160 ? (buf_len - 160 + 79) / 80 : 0) + 2);   // 2 words per 80-sample block
    const size_t per_warp = (size_t)buf_bytes + (size_t)max_frames * 4;
    int warps = (int)((220 * 1024) / per_warp);
    if (warps < 1) return cudaErrorInvalidValue;
    if (warps > kVadMaxWarps) warps = kVadMaxWarps;
    const size_t smem = per_warp * warps;
    cudaError_t e = cudaFuncSetAttribute(vad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);   // + 96 B static barriers <= 227 KB
    if (e != cudaSuccess) return e;
    u32 grid = (B + warps - 1) / warps;
    size_t resident = (224 * 1024) / (smem + 2048);                     // CTAs that fit one SM (smem-limited)
    if (resident < 1) resident = 1;
    const u32 cap = (u32)num_sms * (u32)resident;                       // persistent: one wave, warps stride over utterances
    if (grid > cap) grid = cap;
    vad_kernel<<<grid, warps * 32, smem, st>>>(pcm, U, B, n_len, buf_len, do_atap, do_vad, atap, seg_off, buf_bytes,
                                              max_frames);
    return cudaGetLastError();
}

}  // namespace srk
