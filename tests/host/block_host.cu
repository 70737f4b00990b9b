// block_host.cu -- TEST HARNESS: the kernels' per-block device functions (m1cu_block.cuh: forward
// DCT, pack + non-zero test, quantiser, DC/AC coder, register bit accumulator) compiled for the
// HOST, so the CPU suite can compare them with the oracle block by block (tests/test_block_host.py).
// Not part of libm1cu.so and not a fallback: the product has no host encode path.
#include "../../ec504_imageencoder_b200/csrc/m1cu_block.cuh"
#include "../../ec504_imageencoder_b200/csrc/m1cu_quant.h"
#include "../../ec504_imageencoder_b200/csrc/m1cu_colour.cuh"
#include <string.h>

namespace {
M1Quant g_q;
M1Tables g_tb;
M1NzKeys g_nk;

struct StrSink {                      // every bit of the block as '0'/'1' characters
    char *out; int cap; int n;
    __host__ __device__ void put(uint32_t code, int len)
    {
        for (int i = len - 1; i >= 0; --i) { if (n < cap) out[n] = ((code >> i) & 1u) ? '1' : '0'; ++n; }
    }
};
}  // namespace

extern "C" {

// scaled quantiser matrix (raster order) -> the coder's tables; 0 when every self-check passed
int m1bh_set_matrix(const int32_t qm[64])
{
    if (!m1_make_quant(qm, &g_q)) return -1;
    m1k_fill_tables(&g_tb, g_q);
    m1k_nz_keys(g_q, &g_nk);
    return 0;
}

// One block exactly as a thread of k_encode_chunks handles it.  samples: 64 raster int samples;
// tid/key: record slot and swizzle key.  Outputs: dct[64] raster (bias removed), levels[64] zigzag,
// bits ('0'/'1', returns the count), the register accumulator after finish().
int m1bh_block(const int32_t samples[64], int is_luma, int first_in_mb, int tid, int key,
               int32_t dct[64], int16_t levels[64], char *bits, int cap,
               uint32_t *acc_hi, uint32_t *acc_lo, int *acc_n, unsigned long long *nz_out, int *bad_out)
{
    int v[64];
    for (int i = 0; i < 64; ++i) v[i] = samples[i];
    fdct8x8(v);
    for (int i = 0; i < 64; ++i) dct[i] = v[i] - M1_COEF_BIAS;
    uint32_t pk[32];
    const unsigned long long nz = pack_and_flag(v, pk, g_nk);
    *nz_out = nz;
    static short rec[129 * 128];
    if (tid < 0 || tid > 128) return -1;
    for (int gI = 0; gI < 8; ++gI)                         // the kernel's swizzled 128-bit record stores
        memcpy(rec + tid * 128 + (((gI ^ key) & 7) << 3), pk + 4 * gI, 16);
    for (int z = 0; z < 64; ++z) {
        const int direct = rec[rec_index(tid, z, key)];
        levels[z] = (int16_t)quant_level(direct, z, &g_tb);
    }
    BitAcc acc{0u, 0u, 0};
    if (first_in_mb) { acc.lo = 3u; acc.n = 2; }
    const int bad = code_block(acc, rec, tid, nz, is_luma != 0, &g_tb, key);
    acc.finish();
    *acc_hi = acc.hi; *acc_lo = acc.lo; *acc_n = acc.n; *bad_out = bad;
    StrSink ss{bits, cap, 0};
    if (first_in_mb) ss.put(3u, 2);
    code_block(ss, rec, tid, nz, is_luma != 0, &g_tb, key);
    return ss.n;
}

// The coder alone: zz[64] = wanted quantised levels in zigzag order; the coefficient of position z
// is set to level * m (m = matrix entry), which quantises back to exactly that level.  Needs
// |level * m| <= 2047.  Returns the bit count ('0'/'1' in bits), -1 if a coefficient is out of range.
int m1bh_code_levels(const int32_t zz[64], int is_luma, int tid, int key, char *bits, int cap,
                     uint32_t *acc_hi, uint32_t *acc_lo, int *acc_n, int *bad_out)
{
    int v[64];
    for (int z = 0; z < 64; ++z) {
        const int k = zz_raster(z), m = g_q.ta[k] + 1, c = zz[z] * m;
        if (c < -2048 || c > 2047) return -1;
        v[k] = c + M1_COEF_BIAS;
    }
    uint32_t pk[32];
    const unsigned long long nz = pack_and_flag(v, pk, g_nk);
    for (int z = 0; z < 64; ++z) if (((nz >> z) & 1ull) != (zz[z] != 0 ? 1ull : 0ull)) return -2;
    static short rec[129 * 128];
    if (tid < 0 || tid > 128) return -1;
    for (int gI = 0; gI < 8; ++gI)
        memcpy(rec + tid * 128 + (((gI ^ key) & 7) << 3), pk + 4 * gI, 16);
    BitAcc acc{0u, 0u, 0};
    const int bad = code_block(acc, rec, tid, nz, is_luma != 0, &g_tb, key);
    acc.finish();
    *acc_hi = acc.hi; *acc_lo = acc.lo; *acc_n = acc.n; *bad_out = bad;
    StrSink ss{bits, cap, 0};
    code_block(ss, rec, tid, nz, is_luma != 0, &g_tb, key);
    return ss.n;
}

// The kernels' integer colour path (m1cu_colour.cuh) over ALL 2^24 colours, in each of the four byte
// alignments a 3-byte pixel can have inside 32-bit words (neighbouring bytes filled with `junk`, which
// the zero coefficients must ignore).  The claim the kernel relies on: an UNFLAGGED pixel's three
// quotients equal the reference's double chain.  Returns the number of violations (must be 0).
// stats: [0] flagged colours (alignment 0), [1..3] colours whose double chain differs from the integer
// quotient for Y / Cb / Cr (all of them must be flagged), [4] flagged colours whose quotients were right.
long m1bh_colour_sweep(unsigned junk, long long stats[5])
{
    long bad = 0;
    for (int i = 0; i < 5; ++i) stats[i] = 0;
    const uint32_t J = (junk & 0xffu) * 0x01010101u;
    for (int r = 0; r < 256; ++r)
        for (int g = 0; g < 256; ++g)
            for (int b = 0; b < 256; ++b) {
                int ye, cbe, cre;
                ycbcr_exact_host(r, g, b, ye, cbe, cre);
                for (int sh = 0; sh < 4; ++sh) {
                    // bytes sh, sh+1, sh+2 of the 8-byte window (w0 low word) hold r, g, b
                    unsigned long long win = ((unsigned long long)J << 32) | J;
                    win &= ~(0xffffffull << (8 * sh));
                    win |= ((unsigned long long)r | ((unsigned long long)g << 8) | ((unsigned long long)b << 16)) << (8 * sh);
                    int y, cb, cr;
                    uint32_t fm = 0xffffffffu;
                    colour_int_pixel((uint32_t)win, (uint32_t)(win >> 32), sh, y, cb, cr, fm);
                    const bool flagged = fm < M1_COLOUR_FLAG_LIMIT;
                    const bool same = y == ye && cb == cbe && cr == cre;
                    if (!flagged && !same) ++bad;
                    if (sh == 0) {
                        stats[0] += flagged;
                        stats[1] += y != ye; stats[2] += cb != cbe; stats[3] += cr != cre;
                        stats[4] += flagged && same;
                    }
                }
            }
    return bad;
}

// the harness's own double chain for an array of pixels (checked against the oracle by the test)
void m1bh_ycbcr_exact(const unsigned char *rgb, long n, unsigned char *y, unsigned char *cb, unsigned char *cr)
{
    for (long i = 0; i < n; ++i) {
        int a, b, c;
        ycbcr_exact_host(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], a, b, c);
        y[i] = (unsigned char)a; cb[i] = (unsigned char)b; cr[i] = (unsigned char)c;
    }
}

// The flat-block bound of m1cu_quant.h: largest sample span for which no AC level can be non-zero (-1: none), and the
// L1 norms of the row- / column-pass functionals it is built from.
int m1bh_flat_range(const int32_t qm[64]) { return m1_flat_range(qm); }
void m1bh_l1_norms(double a_row[8], double a_col[8]) { m1_fdct_l1_norms(a_row, a_col); }
// the real-valued (truncation-free) 2-D transform of a block: out[u*8+v]
void m1bh_fdct_linear(const double samples[64], double out[64])
{
    double rows[64], col[8], r[8];
    for (int i = 0; i < 8; ++i) m1_fdct_linear_1d(samples + 8 * i, rows + 8 * i, false);
    for (int j = 0; j < 8; ++j) {
        for (int i = 0; i < 8; ++i) col[i] = rows[8 * i + j];
        m1_fdct_linear_1d(col, r, true);
        for (int u = 0; u < 8; ++u) out[8 * u + j] = r[u];
    }
}
}  // extern "C"
