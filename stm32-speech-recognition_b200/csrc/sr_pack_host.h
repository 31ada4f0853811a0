// sr_pack_host.h -- host-side 12-bit PCM packer + worker pool (sr_pack_host.cpp); private to the library.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace srk {

// pack n samples (n even) into 3 bytes per pair (a | b << 12, little endian); returns the OR of all samples:
// the packing is lossless iff (result & 0xF000) == 0
uint32_t pack12(const uint16_t *src, size_t n, uint8_t *dst);
// test hook: 0 scalar, 1 AVX2, 2 AVX-512 VBMI; 0xFFFFFFFF if the variant is not available on this CPU
uint32_t pack12_variant(int variant, const uint16_t *src, size_t n, uint8_t *dst);
// CPUs this process may use (affinity capped by the cgroup quota)
int usable_cpus();

class PackPool {
public:
    explicit PackPool(int nthreads);
    ~PackPool();
    PackPool(const PackPool &) = delete;
    PackPool &operator=(const PackPool &) = delete;
    int threads() const;
    // pack [src, src+n) into dst with all workers (blocks until done); returns the OR of all samples
    uint32_t run(const uint16_t *src, size_t n, uint8_t *dst);
private:
    struct Impl;
    Impl *p;
};

}  // namespace srk
