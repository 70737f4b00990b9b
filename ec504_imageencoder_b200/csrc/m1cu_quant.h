// m1cu_quant.h -- host-side construction of the coder's constant tables (inline, shared by
// m1cu_api.cu, m1cu_kernels.cu and the host test harness tests/host/block_host.cu).
#pragma once
#include "m1cu_block.cuh"
#include "m1cu_tables.h"

// Reciprocals and non-zero thresholds of the scaled matrix qm (raster order).  Every claim the
// kernels rely on is checked here for all -2048 <= c <= 2047 (the 12-bit biased range): the multiply-high quotient equals C's
// truncating division, and the unsigned-range test equals "level != 0".
inline bool m1_make_quant(const int32_t qm[64], M1Quant *q)
{
    for (int k = 0; k < 64; ++k) {
        const int m = qm[k];
        if (m < 1 || m > 8192) return false;               // packed non-zero test needs 0x7800 - m > 0 with room
        const uint32_t R = (uint32_t)(((1ull << 31) + (unsigned)m - 1) / (unsigned)m);
        q->rcp[k] = R; q->ta[k] = m - 1; q->tb[k] = 2 * m - 2;
        for (int c = -2048; c <= 2047; ++c) {
            const uint32_t a = (uint32_t)(c < 0 ? -c : c);
            const uint32_t mag = (uint32_t)(((unsigned long long)(2u * a + 1u) * R) >> 32);
            const int got = c < 0 ? -(int)mag : (int)mag;
            if (got != c / m) return false;
            if (((unsigned)(c + q->ta[k]) > (unsigned)q->tb[k]) != (c / m != 0)) return false;
        }
    }
    return true;
}

inline void m1k_nz_keys(const M1Quant &q, M1NzKeys *k)
{
    for (int w = 0; w < 32; ++w) {
        const int zlo = (w & 15) + ((w >> 4) << 5), zhi = zlo + 16;
        const uint32_t mlo = (uint32_t)q.ta[zz_raster(zlo)] + 1u, mhi = (uint32_t)q.ta[zz_raster(zhi)] + 1u;
        k->ka[w] = ((0x7800u - mhi) << 16) | (0x7800u - mlo);
        k->kb[w] = ((0x8800u - mhi) << 16) | (0x8800u - mlo);
    }
}

inline void m1k_fill_tables(M1Tables *t, const M1Quant &q)
{
    for (int i = 0; i < 112; ++i) t->ac[i] = i < M1_AC_ENTRIES ? kM1AcTable[i] : 0u;
    t->ac[0] = 0x02000003u;                                  // (run 0, |level| 1) -> '11' (source/vlc.c:330)
    for (int i = 0; i < 18; ++i) t->dc[i] = kM1DcSize[i];
    for (int r = 0; r < 64; ++r)
        t->acrun[r] = r < 32 ? (uint16_t)(kM1AcFirst[r] | ((kM1AcFirst[r + 1] - kM1AcFirst[r]) << 8)) : (uint16_t)0;
    for (int z = 0; z < 64; ++z) { t->qrcp[z] = q.rcp[zz_raster(z)]; t->zofs[z] = (uint8_t)rec_byte_offset(z); }
    M1NzKeys nk;
    m1k_nz_keys(q, &nk);
    for (int w = 0; w < 32; ++w) { t->ka[w] = nk.ka[w]; t->kb[w] = nk.kb[w]; }
}
