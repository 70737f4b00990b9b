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

// -------------------------------------------------------------------------------------------
// Flat-block bound.  Every AC coefficient of fast_DCT is, up to its truncating shifts, a LINEAR functional of the
// 64 samples that does not see a constant added to all of them (only differences and zero-sum combinations of the
// inputs reach an AC output, source/image_processing.c:210-305).  So for a block whose samples lie within
// [lo, lo + R]:   |c(u,v)| <= ceil(R / 2) * W(u,v) + E(u),  (u,v) != (0,0),  where
//   W(u,v) = A_col(u) * A_row(v), the L1 norms of the column- and row-pass functionals (m1_fdct_linear_1d: the same
//            butterflies in exact real arithmetic, shifts as divisions), and
//   E(u)   = A_col(u) + 2.05 bounds what the truncations add: every row-pass output is off by less than 1, the column
//            functional spreads that with L1 gain A_col(u), and the column pass's own rounding constant and floor add at
//            most 1.5 (0.5 + 1), or 2.05 for u = 3, 5 (floor(t / 256) * 181 / 4096 + 2 inside a floor).
// m1_flat_range returns the largest even R for which that bound stays below the quantiser entry at EVERY AC position
// (then every AC level of such a block is zero, whatever its samples), or -1 when there is none.  The DC coefficient
// of any block is (sum of the 64 samples + 16) >> 3.  Checked against the oracle in tests/test_block_host.py.
// -------------------------------------------------------------------------------------------
inline void m1_fdct_linear_1d(const double x[8], double out[8], bool column)
{
    const double c1 = 1004, s1 = 200, c3 = 851, s3 = 569, r2c6 = 554, r2s6 = 1337, r2 = 181;
    const double d0 = x[0] - x[7], d1 = x[1] - x[6], d2 = x[2] - x[5], d3 = x[3] - x[4];
    const double a0 = x[0] + x[7], a1 = x[1] + x[6], a2 = x[2] + x[5], a3 = x[3] + x[4];
    const double e0 = a0 + a3, e3 = a0 - a3, e1 = a1 + a2, e2 = a1 - a2;
    const double m12 = c1 * (d1 + d2), p2 = (-s1 - c1) * d2 + m12, p1 = (s1 - c1) * d1 + m12;
    const double m03 = c3 * (d0 + d3), p3 = (-s3 - c3) * d3 + m03, p0 = (s3 - c3) * d0 + m03;
    const double m78 = r2c6 * (e2 + e3), t2 = (r2s6 - r2c6) * e3 + m78, t3 = (-r2s6 - r2c6) * e2 + m78;
    const double t5 = p0 + p2, t7 = p0 - p2, t6 = p3 - p1, t4 = p3 + p1;
    const double s_a = column ? 8.0 : 1.0, s_b = column ? 8192.0 : 1024.0, s_c = column ? 1048576.0 : 131072.0;
    out[0] = (e0 + e1) / s_a; out[4] = (e0 - e1) / s_a;
    out[2] = t2 / s_b; out[6] = t3 / s_b; out[1] = (t4 + t5) / s_b; out[7] = (t4 - t5) / s_b;
    out[3] = t6 * r2 / s_c; out[5] = t7 * r2 / s_c;
}
inline void m1_fdct_l1_norms(double a_row[8], double a_col[8])
{
    for (int k = 0; k < 8; ++k) a_row[k] = a_col[k] = 0.0;
    for (int i = 0; i < 8; ++i) {
        double x[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }, r[8], c[8];
        x[i] = 1.0;
        m1_fdct_linear_1d(x, r, false);
        m1_fdct_linear_1d(x, c, true);
        for (int k = 0; k < 8; ++k) { a_row[k] += r[k] < 0 ? -r[k] : r[k]; a_col[k] += c[k] < 0 ? -c[k] : c[k]; }
    }
}
inline int m1_flat_range(const int32_t qm[64])
{
    double a_row[8], a_col[8];
    m1_fdct_l1_norms(a_row, a_col);
    int h = 255;                                           // ceil(R / 2) can never exceed 128
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) {
            if (u == 0 && v == 0) continue;
            const double room = (double)qm[u * 8 + v] - (a_col[u] + 2.05);     // h * W < room
            if (room <= 0.0) return -1;
            int hk = (int)(room / (a_col[u] * a_row[v]));
            while (hk > 0 && (double)hk * a_col[u] * a_row[v] >= room) --hk;   // strict inequality, whatever the division rounded to
            if (hk < h) h = hk;
        }
    return h > 0 ? 2 * h : -1;
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
