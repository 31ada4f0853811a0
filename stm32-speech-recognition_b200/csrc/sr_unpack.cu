// sr_unpack.cu -- device side of the packed PCM transport (sr_pack_host.cpp): 12 bits per sample, 3 bytes per
// sample pair (a | b << 12, little endian) -> u16 samples. One thread per 16 samples (24 packed bytes, three 8-byte
// loads -> two 16-byte stores); the tail (< 16 samples) is expanded by one thread byte-wise.
#include "sr_common.cuh"

namespace srk {

__global__ void __launch_bounds__(256)
unpack12_kernel(const uint2 *in, uint4 *out, u64 groups, const u8 *in_bytes,   // (in, in_bytes) and (out, out16) alias
                u16 *out16, u64 n) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        const uint2 a = in[3 * g], b = in[3 * g + 1], c = in[3 * g + 2];
        const u64 w0 = ((u64)a.y << 32) | a.x, w1 = ((u64)b.y << 32) | b.x, w2 = ((u64)c.y << 32) | c.x;
        // pair k occupies bits [24k, 24k+24) of the 192-bit string w2:w1:w0
        u32 p[8];
        p[0] = (u32)w0 & 0xFFFFFFu;
        p[1] = (u32)(w0 >> 24) & 0xFFFFFFu;
        p[2] = (u32)((w0 >> 48) | (w1 << 16)) & 0xFFFFFFu;
        p[3] = (u32)(w1 >> 8) & 0xFFFFFFu;
        p[4] = (u32)(w1 >> 32) & 0xFFFFFFu;
        p[5] = (u32)((w1 >> 56) | (w2 << 8)) & 0xFFFFFFu;
        p[6] = (u32)(w2 >> 16) & 0xFFFFFFu;
        p[7] = (u32)(w2 >> 40) & 0xFFFFFFu;
        u32 q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) q[k] = (p[k] & 0xFFFu) | ((p[k] >> 12) << 16);      // a | b << 16: two u16 samples
        out[2 * g] = make_uint4(q[0], q[1], q[2], q[3]);
        out[2 * g + 1] = make_uint4(q[4], q[5], q[6], q[7]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (u64 i = groups * 16; i + 2 <= n; i += 2) {
            const u8 *s = in_bytes + i / 2 * 3;
            const u32 t = (u32)s[0] | ((u32)s[1] << 8) | ((u32)s[2] << 16);
            out16[i] = (u16)(t & 0xFFFu);
            out16[i + 1] = (u16)(t >> 12);
        }
    }
}

// packed and out must be 16-byte aligned (cudaMalloc'd bases are); n even
cudaError_t launch_unpack12(const void *packed, u64 n, u16 *out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const u64 groups = n / 16;
    u64 blocks = (groups + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 16) blocks = 148 * 16;
    unpack12_kernel<<<(unsigned)blocks, 256, 0, st>>>(static_cast<const uint2 *>(packed), reinterpret_cast<uint4 *>(out), groups,
                                                      static_cast<const u8 *>(packed), out, n);
    return cudaGetLastError();
}

}  // namespace srk
