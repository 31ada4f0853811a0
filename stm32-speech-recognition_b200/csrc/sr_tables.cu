// sr_tables.cu -- builds the constant tables of the MFCC path once per device and keeps them in
// global memory (kernels stage what they need into shared memory). Everything derives from the
// closed-form tables of tools/gen_tables.py (sr_tables.h); see there for the reference citations.
#include <mutex>
#include <string.h>
#include "sr_common.cuh"
#include "sr_tables.h"

namespace srk {

static void build_tables(DevTables &t) {
    memset(&t, 0, sizeof t);
    for (int i = 0; i < 340 * 3; ++i) {
        const int ka = sr_tab_twiddle[2 * i], kb = sr_tab_twiddle[2 * i + 1];
        t.tw[i] = make_int2(ka + kb, kb);
    }
    for (int i = 0; i < 2220; ++i) t.log_thr[i] = sr_tab_log_thr[i];
    for (int i = 0; i < 160; ++i) t.hamm[i] = sr_tab_hamm[i];
    for (int i = 0; i < 512; ++i) { t.tri_even[i] = sr_tab_tri_even[i]; t.tri_odd[i] = sr_tab_tri_odd[i]; }
    for (int i = 0; i < 288; ++i) t.dct[i] = sr_tab_dct[i];
    // Filter ranges, MFCC.C:136-162. Even filters h=0,2,..,22 tile [0,cen[1]) [cen[1],cen[3]) .. [cen[21],cen[23])
    // and use tri_even; odd filters h=1,3,..,23 tile [cen[0],cen[2]) .. [cen[20],cen[22]) [cen[22],512), tri_odd.
    const uint16_t *cen = sr_tab_tri_cen;
    int bnd[2][13];
    bnd[0][0] = 0;
    for (int j = 1; j <= 12; ++j) bnd[0][j] = cen[2 * j - 1];
    for (int j = 0; j < 12; ++j) bnd[1][j] = cen[2 * j];
    bnd[1][12] = 512;
    for (int par = 0; par < 2; ++par)
        for (int j = 0; j < 12; ++j) {
            const int h = 2 * j + par, lo = bnd[par][j], hi = bnd[par][j + 1];
            auto e_off = [&](int k) { return k >= 512 ? kFltZero : par * kFltRowWords + flt_word(k >> 4, k & 15); };
            t.flt_lo[h] = (u16)lo; t.flt_hi[h] = (u16)hi;
            t.flt_e_lo[h] = (u16)e_off(lo); t.flt_e_hi[h] = (u16)e_off(hi);
            t.flt_x_lo[h] = (u8)(lo >> 4); t.flt_x_hi[h] = (u8)(hi >> 4);   // 512 >> 4 = 32: the grand total
        }
    // GEOM_B extension: same construction over 128 bins (MFCC.C:136-162 with the regenerated centres)
    for (int i = 0; i < 200; ++i) t.b_hamm[i] = sr_tab_b_hamm[i];
    for (int i = 0; i < 128; ++i) { t.b_tri_even[i] = sr_tab_b_tri_even[i]; t.b_tri_odd[i] = sr_tab_b_tri_odd[i]; }
    const uint16_t *cb = sr_tab_b_tri_cen;
    for (int h = 0; h < 24; ++h) {
        t.b_flt_lo[h] = (u16)(h == 0 ? 0 : cb[h - 1]);
        t.b_flt_hi[h] = (u16)(h == 23 ? 128 : cb[h + 1]);
    }
}

const DevTables *dev_tables() {
    static std::mutex mu;
    static DevTables *ptr[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!ptr[dev]) {
        DevTables *h = new DevTables;
        build_tables(*h);
        DevTables *d = nullptr;
        if (cudaMalloc(&d, sizeof(DevTables)) != cudaSuccess) { delete h; return nullptr; }
        if (cudaMemcpy(d, h, sizeof(DevTables), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); delete h; return nullptr; }
        delete h;
        ptr[dev] = d;
    }
    return ptr[dev];
}

// host-side copy for unit tests of the filter ranges (no GPU needed): bins [lo[h], hi[h]) per filter, the scratch word
// offsets the kernel reads for S(lo), S(hi), and the X indices
extern "C" void sr_debug_filter_ranges(uint16_t *lo, uint16_t *hi, uint16_t *e_lo, uint16_t *e_hi, uint8_t *x_lo, uint8_t *x_hi) {
    DevTables *h = new DevTables;
    build_tables(*h);
    memcpy(lo, h->flt_lo, 48); memcpy(hi, h->flt_hi, 48);
    memcpy(e_lo, h->flt_e_lo, 48); memcpy(e_hi, h->flt_e_hi, 48);
    memcpy(x_lo, h->flt_x_lo, 24); memcpy(x_hi, h->flt_x_hi, 24);
    delete h;
}

}  // namespace srk
