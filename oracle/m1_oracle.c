/*
 * m1_oracle.c -- plain-C restatement of the reference's per-block MPEG-1 I-frame path.
 * TEST INFRASTRUCTURE ONLY (see m1_oracle.h).  Parity pinned against oracle/_ref/libm1ref.so
 * (the unmodified reference sources) by tests/test_oracle_vs_ref.py and against the committed
 * golden vectors by tests/test_oracle_golden.py.
 *
 * Written from the behaviour of the reference (file:line cited per function), not from its
 * text: one MSB-first bit writer instead of malloc'd bit vectors, packed tables instead of
 * C strings, one pass per picture.  Quirks of the reference are reproduced on purpose.
 */
#include "m1_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * MSB-first bit writer.  Semantics of source/bit_vector.c:13-42 (put_bit), :44-83
 * (put_byte_off) and :100-121 (concat): bit k of the stream is bit 7-(k%8) of byte k/8.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t *buf;
    long     cap_bytes;
    long     nbits;
    int      overflow;
} bitw_t;

static void bw_put(bitw_t *w, uint32_t code, int len)
{
    for (int k = len - 1; k >= 0; --k) {
        long byte = w->nbits >> 3;
        if (byte >= w->cap_bytes) { w->overflow = 1; w->nbits++; continue; }
        int sh = 7 - (int)(w->nbits & 7);
        if (sh == 7) w->buf[byte] = 0;
        w->buf[byte] |= (uint8_t)(((code >> k) & 1u) << sh);
        w->nbits++;
    }
}

static void bw_put_str(bitw_t *w, const char *s)
{
    for (; *s; ++s) bw_put(w, (uint32_t)(*s == '1'), 1);
}

/* ------------------------------------------------------------------------------------------
 * Tables.
 * ---------------------------------------------------------------------------------------- */

/* Default intra matrix, source/image_processing.c:17-26. */
static const int kIntraQ[64] = {
     8, 16, 19, 22, 26, 27, 29, 34,   16, 16, 22, 24, 27, 29, 34, 37,
    19, 22, 26, 27, 29, 34, 34, 38,   22, 22, 26, 27, 29, 34, 37, 40,
    22, 26, 27, 29, 32, 35, 40, 48,   26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69,   27, 29, 35, 38, 46, 56, 69, 83 };

/* Zigzag rank of raster position i*8+j, source/image_processing.c:28-37. */
static const uint8_t kZigzagRank[64] = {
     0,  1,  5,  6, 14, 15, 27, 28,    2,  4,  7, 13, 16, 26, 29, 42,
     3,  8, 12, 17, 25, 30, 41, 43,    9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54,   20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61,   35, 36, 48, 49, 57, 58, 62, 63 };

/* dct_dc_size codes, source/vlc.c:121-144, index = size 0..8. */
static const char *const kDcSizeLuma[9]   = { "100", "00", "01", "101", "110", "1110", "11110",
                                              "111110", "1111110" };
static const char *const kDcSizeChroma[9] = { "00", "01", "10", "110", "1110", "11110", "111110",
                                              "1111110", "11111110" };

/* AC run/level codes without sign bit, source/vlc.c:176-288, one string per run (0..31), codes
 * separated by blanks, in ascending level order.  Run 0 starts at level 2 (source/vlc.c:177);
 * the lookup below indexes it with |level|-1 exactly as source/vlc.c:337-340 does, which is
 * what shifts run-0 codes by one level.  The (16,2) entry has 15 bits in the reference
 * (source/vlc.c:270), one fewer than ISO 11172-2 B.5c; kept as the reference has it. */
static const char *const kAcCodes[32] = {
    /* 0*/ "0100 00101 0000110 00100110 00100001 0000001010 000000011101 000000011000 "
           "000000010011 000000010000 0000000011010 0000000011001 0000000011000 0000000010111 "
           "00000000011111 00000000011110 00000000011101 00000000011100 00000000011011 "
           "00000000011010 00000000011001 00000000011000 00000000010111 00000000010110 "
           "00000000010101 00000000010100 00000000010011 00000000010010 00000000010001 "
           "00000000010000 000000000011000 000000000010111 000000000010110 000000000010101 "
           "000000000010100 000000000010011 000000000010010 000000000010001 000000000010000",
    /* 1*/ "011 000110 00100101 0000001100 000000011011 0000000010110 0000000010101 "
           "000000000011111 000000000011110 000000000011101 000000000011100 000000000011011 "
           "000000000011010 000000000011001 0000000000010011 0000000000010010 0000000000010001 "
           "0000000000010000",
    /* 2*/ "0101 0000100 0000001011 000000010100 0000000010100",
    /* 3*/ "00111 00100100 000000011100 0000000010011",
    /* 4*/ "00110 0000001111 000000010010",
    /* 5*/ "000111 0000001001 0000000010010",
    /* 6*/ "000101 000000011110 0000000000010100",
    /* 7*/ "000100 000000010101",
    /* 8*/ "0000111 000000010001",
    /* 9*/ "0000101 0000000010001",
    /*10*/ "00100111 0000000010000",
    /*11*/ "00100011 0000000000011010",
    /*12*/ "00100010 0000000000011001",
    /*13*/ "00100000 0000000000011000",
    /*14*/ "0000001110 0000000000010111",
    /*15*/ "0000001101 0000000000010110",
    /*16*/ "0000001000 000000000010101",
    /*17*/ "000000011111", /*18*/ "000000011010", /*19*/ "000000011001", /*20*/ "000000010111",
    /*21*/ "000000010110", /*22*/ "0000000011111", /*23*/ "0000000011110", /*24*/ "0000000011101",
    /*25*/ "0000000011100", /*26*/ "0000000011011", /*27*/ "0000000000011111",
    /*28*/ "0000000000011110", /*29*/ "0000000000011101", /*30*/ "0000000000011100",
    /*31*/ "0000000000011011" };

/* Returns the idx-th blank-separated token of run r's list into tok, or 0 when there is none. */
static int ac_table_code(int r, int idx, char tok[24])
{
    const char *p = kAcCodes[r];
    for (int k = 0; *p; ++k) {
        while (*p == ' ') ++p;
        const char *e = p;
        while (*e && *e != ' ') ++e;
        if (e == p) break;
        if (k == idx) {
            int n = (int)(e - p);
            memcpy(tok, p, (size_t)n);
            tok[n] = 0;
            return n;
        }
        p = e;
    }
    return 0;
}

/* Exposed for the table cross-check in tests: code string for (run r, table index a) or "". */
int m1o_ac_table_entry(int r, int a, char *out24)
{
    out24[0] = 0;
    if (r < 0 || r > 31 || a < 0) return 0;
    return ac_table_code(r, a, out24);
}

/* ------------------------------------------------------------------------------------------
 * Stage functions.
 * ---------------------------------------------------------------------------------------- */

/* source/image_processing.c:314-343.  The scale factor is a float; the product int*float is a
 * float (rounded to single), the division by 100.0 is done in double, then round() half away. */
void m1o_qmatrix(int q, int32_t out[64])
{
    if (q < 1) q = 1;
    if (q > 100) q = 100;
    float sf = (q < 50) ? (float)(5000.0 / q) : (float)(200.0 - 2 * q);
    for (int k = 0; k < 64; ++k) {
        float prod = (float)kIntraQ[k] * sf;
        int v = (int)round((double)prod / 100.0);
        out[k] = v < 1 ? 1 : v;
    }
}

/* source/image_processing.c:92-107.  Double arithmetic, left to right, no fused multiply-add
 * (this file must be built with -ffp-contract=off), truncation toward zero. */
void m1o_rgb_to_ycbcr(const uint8_t *rgb, int channels, long npix,
                      uint8_t *Y, uint8_t *Cb, uint8_t *Cr)
{
    for (long p = 0; p < npix; ++p) {
        const uint8_t *px = rgb + p * channels;
        volatile double r = px[0], g = px[1], b = px[2];
        volatile double y  = 0.299 * r;      y  = y + 0.587 * g;       y  = y + 0.114 * b;
        volatile double cb = 0.168736 * r;   cb = 128 - cb;            cb = cb - 0.331264 * g;
        cb = cb + 0.5 * b;
        volatile double cr = 0.5 * r;        cr = 128 + cr;            cr = cr - 0.418688 * g;
        cr = cr - 0.081312 * b;
        Y[p]  = (uint8_t)(int)y;
        Cb[p] = (uint8_t)(int)cb;
        Cr[p] = (uint8_t)(int)cr;
    }
}

/* source/image_processing.c:114-133: truncating mean of each 2x2. */
void m1o_subsample_420(const uint8_t *pl, int W, int H, uint8_t *out)
{
    int sw = W / 2;
    for (int y = 0; y + 1 < H; y += 2) {
        for (int x = 0; x + 1 < W; x += 2) {
            int s = pl[y * W + x] + pl[y * W + x + 1] + pl[(y + 1) * W + x] + pl[(y + 1) * W + x + 1];
            out[(y / 2) * sw + x / 2] = (uint8_t)(s / 4);
        }
    }
}

/* One 8-point pass of source/image_processing.c:210-238 (identical for rows and columns up to
 * the output stage).  in/out hold the eight lanes; on return
 * t[0]=k0 sum, t[1]=k4 diff, t[2]=k2 pre-shift, t[3]=k6 pre-shift, t[4]=x2, t[5]=x5, t[6]=x3, t[7]=x0. */
static void fdct_core(const int32_t in[8], int32_t t[8])
{
    const int32_t c1 = 1004, s1 = 200, c3 = 851, s3 = 569, r2c6 = 554, r2s6 = 1337;
    int32_t a0 = in[0] + in[7], d0 = in[0] - in[7];
    int32_t a1 = in[1] + in[6], d1 = in[1] - in[6];
    int32_t a2 = in[2] + in[5], d2 = in[2] - in[5];
    int32_t a3 = in[3] + in[4], d3 = in[3] - in[4];

    int32_t e0 = a0 + a3, e3 = a0 - a3;          /* x4, x8 of the reference */
    int32_t e1 = a1 + a2, e2 = a1 - a2;          /* x5, x7 */

    int32_t m12 = c1 * (d1 + d2);
    int32_t p2  = (-s1 - c1) * d2 + m12;         /* x2 */
    int32_t p1  = (s1 - c1) * d1 + m12;          /* x1 */
    int32_t m03 = c3 * (d0 + d3);
    int32_t p3  = (-s3 - c3) * d3 + m03;         /* x3 */
    int32_t p0  = (s3 - c3) * d0 + m03;          /* x0 */

    int32_t m78 = r2c6 * (e2 + e3);
    t[0] = e0 + e1;
    t[1] = e0 - e1;
    t[2] = (r2s6 - r2c6) * e3 + m78;             /* x8 */
    t[3] = (-r2s6 - r2c6) * e2 + m78;            /* x7 */
    t[5] = p0 + p2;                              /* x5 */
    t[7] = p0 - p2;                              /* x0 */
    t[4] = p3 + p1;                              /* x2 */
    t[6] = p3 - p1;                              /* x3 */
}

/* source/image_processing.c:192-307.  >> on negative int32 is arithmetic (gcc, x86-64). */
void m1o_fdct8x8(const uint8_t blk[64], int32_t out[64])
{
    const int32_t r2 = 181;
    int32_t rows[64], in[8], t[8];
    for (int i = 0; i < 8; ++i) {
        for (int j = 0; j < 8; ++j) in[j] = blk[i * 8 + j];
        fdct_core(in, t);
        rows[i * 8 + 0] = t[0];
        rows[i * 8 + 4] = t[1];
        rows[i * 8 + 2] = t[2] >> 10;
        rows[i * 8 + 6] = t[3] >> 10;
        rows[i * 8 + 7] = (t[4] - t[5]) >> 10;
        rows[i * 8 + 1] = (t[4] + t[5]) >> 10;
        rows[i * 8 + 3] = (t[6] * r2) >> 17;
        rows[i * 8 + 5] = (t[7] * r2) >> 17;
    }
    for (int j = 0; j < 8; ++j) {
        for (int i = 0; i < 8; ++i) in[i] = rows[i * 8 + j];
        fdct_core(in, t);
        out[0 * 8 + j] = (t[0] + 16) >> 3;
        out[4 * 8 + j] = (t[1] + 16) >> 3;
        out[2 * 8 + j] = (t[2] + 16384) >> 13;
        out[6 * 8 + j] = (t[3] + 16384) >> 13;
        out[7 * 8 + j] = (t[4] - t[5] + 16384) >> 13;
        out[1 * 8 + j] = (t[4] + t[5] + 16384) >> 13;
        out[3 * 8 + j] = ((t[6] >> 8) * r2 + 8192) >> 12;
        out[5 * 8 + j] = ((t[7] >> 8) * r2 + 8192) >> 12;
    }
}

/* source/image_processing.c:365-369: (int)(round(dct)/m) with dct integral == C truncating
 * division; :373-381: zz[rank(i,j)] = q[i][j]. */
void m1o_quant_zigzag(const int32_t dct[64], const int32_t qm[64], int32_t zz[64])
{
    for (int k = 0; k < 64; ++k) zz[kZigzagRank[k]] = dct[k] / qm[k];
}

/* source/vlc.c:315-385 with first == 0 (source/image_processing.c:409-414 always passes 0).
 * z = zeros since the previous non-zero (>= 1 here), L = level.  Returns 0, or -2 when the
 * reference would return NULL (|L| >= 256). */
static int put_ac(bitw_t *w, int z, int L)
{
    int r = z - 1, mag = L < 0 ? -L : L, a = mag - 1;
    char tok[24];
    if (r == 0 && a == 0) { bw_put_str(w, "11"); return 0; }
    if (r <= 31 && ac_table_code(r, a, tok)) { bw_put_str(w, tok); return 0; }
    if (mag >= 256 || r >= 64) return -2;
    bw_put_str(w, "000001");
    bw_put(w, (uint32_t)r & 0x3f, 6);
    if (mag < 128) {
        bw_put(w, (uint32_t)L & 0xff, 8);                 /* 8-bit two's complement */
    } else {
        bw_put(w, L < 0 ? 0x80u : 0x00u, 8);
        bw_put(w, (uint32_t)L & 0xff, 8);                 /* low byte of two's complement */
    }
    return 0;
}

/* source/mpeg1_blk.c:67-117 + source/image_processing.c:400-433 + :703-751. */
static int put_block(bitw_t *w, const int32_t zz[64], int is_luma)
{
    int k = 0, prev = -1;
    /* first non-zero at position 0 <=> RLE pair (level, run 0) leads (source/mpeg1_blk.c:73) */
    if (zz[0] != 0) {
        int v = zz[0], c = v < 0 ? -v : v, sz = 1;
        for (int i = 1; i <= 8; ++i) if (c & (1 << (i - 1))) sz = i;    /* :77-83, bits 0..7 only */
        bw_put_str(w, is_luma ? kDcSizeLuma[sz] : kDcSizeChroma[sz]);    /* source/vlc.c:146-157 */
        if (v < 0) c ^= 1 << (sz - 1);                                   /* :87-89 */
        bw_put(w, (uint32_t)c & ((1u << sz) - 1u), sz);                  /* :91 */
        prev = 0;
        k = 1;
    } else {
        bw_put_str(w, is_luma ? "100" : "00");                           /* :98-102 */
    }
    for (; k < 64; ++k) {
        if (zz[k] == 0) continue;
        int z = k - prev - 1;
        if (z == 0) break;                       /* source/image_processing.c:421-423 */
        int rc = put_ac(w, z, zz[k]);
        if (rc) return rc;
        prev = k;
    }
    bw_put_str(w, "10");                         /* source/mpeg1_blk.c:115-117 */
    return 0;
}

int m1o_block_bits(const int32_t zz[64], int is_luma, char *bits, int cap)
{
    uint8_t buf[160];
    bitw_t w = { buf, (long)sizeof buf, 0, 0 };
    int rc = put_block(&w, zz, is_luma);
    if (rc) return rc;
    if (w.nbits + 1 > cap) return -1;
    for (long k = 0; k < w.nbits; ++k) bits[k] = (char)('0' + ((buf[k >> 3] >> (7 - (k & 7))) & 1));
    bits[w.nbits] = 0;
    return (int)w.nbits;
}

/* ------------------------------------------------------------------------------------------
 * Picture level.
 * ---------------------------------------------------------------------------------------- */

long m1o_picture_macroblocks(int W, int H, int mode)
{
    if (mode == M1O_MODE_REF_COMPAT) return 6 * 9;            /* include/encoder.h:238,248 */
    return (long)((W + 15) / 16) * ((H + 15) / 16);
}

static int code_block(bitw_t *w, const uint8_t blk[64], int is_luma, const int32_t qm[64],
                      int16_t **levels)
{
    int32_t dct[64], zz[64];
    m1o_fdct8x8(blk, dct);
    m1o_quant_zigzag(dct, qm, zz);
    if (*levels) {
        for (int k = 0; k < 64; ++k) (*levels)[k] = (int16_t)zz[k];
        *levels += 64;
    }
    return put_block(w, zz, is_luma);
}

/* source/mpeg1_blk.c:12-20: start code 000001, (vertical_pos+1) & 0xff, 5-bit quant_scale=1, 0. */
static void put_slice_header(bitw_t *w, int vertical_pos)
{
    bw_put(w, 0x000001u, 24);
    bw_put(w, (uint32_t)((vertical_pos & 0xff) + 1) & 0xffu, 8);
    bw_put(w, 1u, 5);
    bw_put(w, 0u, 1);
}

static void gather8x8(const uint8_t *pl, long stride, long x0, long y0, uint8_t blk[64])
{
    /* source/image_processing.c:138-150 */
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) blk[i * 8 + j] = pl[(y0 + i) * stride + x0 + j];
}

long m1o_encode_picture(const uint8_t *rgb, int W, int H, int channels, int mode,
                        const int32_t qm[64], uint8_t *out, long cap, int16_t *levels)
{
    if (!rgb || !out || W <= 0 || H <= 0 || channels < 3) return -3;
    bitw_t w = { out, cap, 0, 0 };
    uint8_t blk[64];
    int rc = 0;

    if (mode == M1O_MODE_REF_COMPAT) {
        /* Literal traversal of include/encoder.h:238-443: the "slice" variable walks columns,
         * the macroblock variable walks rows, chroma blocks are cut from the full-resolution
         * planes with stride W/2 (:347-348). */
        long half = W / 2;
        if (W < 96 || H < 144 || (64 + 7) * half + 40 + 7 >= (long)W * H) return -3;
        long npix = (long)W * H;
        uint8_t *Y = malloc((size_t)npix), *Cb = malloc((size_t)npix), *Cr = malloc((size_t)npix);
        if (!Y || !Cb || !Cr) { free(Y); free(Cb); free(Cr); return -3; }
        m1o_rgb_to_ycbcr(rgb, channels, npix, Y, Cb, Cr);
        int vpos = 0;
        for (int x = 0; x < 96 && !rc; x += 16) {
            put_slice_header(&w, vpos++);
            for (int y = 0; y < 144 && !rc; y += 16) {
                bw_put_str(&w, "11");                         /* source/mpeg1_blk.c:38-58 */
                for (int b = 0; b < 4 && !rc; ++b) {
                    gather8x8(Y, W, x + (b % 2) * 8, y + (b / 2) * 8, blk);
                    rc = code_block(&w, blk, 1, qm, &levels);
                }
                if (!rc) { gather8x8(Cb, half, x / 2, y / 2, blk); rc = code_block(&w, blk, 0, qm, &levels); }
                if (!rc) { gather8x8(Cr, half, x / 2, y / 2, blk); rc = code_block(&w, blk, 0, qm, &levels); }
            }
            while (w.nbits & 7) bw_put(&w, 0, 1);             /* include/encoder.h:442-443 */
        }
        free(Y); free(Cb); free(Cr);
    } else {
        /* FULL: same per-block functions, raster macroblocks over the coded frame (dimensions
         * rounded up to 16 by edge replication), chroma from the 2x2 mean. */
        int Wc = (W + 15) & ~15, Hc = (H + 15) & ~15;
        long npix = (long)Wc * Hc;
        uint8_t *pad = malloc((size_t)npix * 3);
        uint8_t *Y = malloc((size_t)npix), *Cb = malloc((size_t)npix), *Cr = malloc((size_t)npix);
        uint8_t *Cbs = malloc((size_t)npix / 4), *Crs = malloc((size_t)npix / 4);
        if (!pad || !Y || !Cb || !Cr || !Cbs || !Crs) {
            free(pad); free(Y); free(Cb); free(Cr); free(Cbs); free(Crs);
            return -3;
        }
        for (int y = 0; y < Hc; ++y) {
            int sy = y < H ? y : H - 1;
            for (int x = 0; x < Wc; ++x) {
                int sx = x < W ? x : W - 1;
                memcpy(pad + ((long)y * Wc + x) * 3, rgb + ((long)sy * W + sx) * channels, 3);
            }
        }
        m1o_rgb_to_ycbcr(pad, 3, npix, Y, Cb, Cr);
        m1o_subsample_420(Cb, Wc, Hc, Cbs);
        m1o_subsample_420(Cr, Wc, Hc, Crs);
        for (int my = 0; my < Hc / 16 && !rc; ++my) {
            put_slice_header(&w, my);
            for (int mx = 0; mx < Wc / 16 && !rc; ++mx) {
                bw_put_str(&w, "11");
                for (int b = 0; b < 4 && !rc; ++b) {
                    gather8x8(Y, Wc, mx * 16 + (b % 2) * 8, my * 16 + (b / 2) * 8, blk);
                    rc = code_block(&w, blk, 1, qm, &levels);
                }
                if (!rc) { gather8x8(Cbs, Wc / 2, mx * 8, my * 8, blk); rc = code_block(&w, blk, 0, qm, &levels); }
                if (!rc) { gather8x8(Crs, Wc / 2, mx * 8, my * 8, blk); rc = code_block(&w, blk, 0, qm, &levels); }
            }
            while (w.nbits & 7) bw_put(&w, 0, 1);
        }
        free(pad); free(Y); free(Cb); free(Cr); free(Cbs); free(Crs);
    }
    if (rc) return rc;
    if (w.overflow) return -1;
    return w.nbits >> 3;
}

/* ------------------------------------------------------------------------------------------
 * Stream assembly.
 * ---------------------------------------------------------------------------------------- */

static void put_mux_rate(uint32_t rate, uint8_t *o3)
{
    /* source/mpeg1_enc.c:14-20 / :29-35 (little-endian byte picks of ((rate|1<<22)<<1)|1) */
    uint32_t v = (((rate & 0x3fffffu) | 0x400000u) << 1) | 1u;
    o3[0] = (uint8_t)(v >> 16); o3[1] = (uint8_t)(v >> 8); o3[2] = (uint8_t)v;
}

int m1o_file_prologue(uint8_t out[27])
{
    static const uint8_t pack[9] = { 0x00, 0x00, 0x01, 0xba, 0x21, 0x00, 0x01, 0x00, 0x01 };
    memcpy(out, pack, 9);
    put_mux_rate(2202035u, out + 9);                                  /* include/encoder.h:86 */
    static const uint8_t sys0[6] = { 0x00, 0x00, 0x01, 0xbb, 0x00, 0x09 };
    memcpy(out + 12, sys0, 6);
    put_mux_rate(2202035u, out + 18);                                 /* include/encoder.h:88 */
    out[21] = 0x00; out[22] = 0x21; out[23] = 0xff; out[24] = 0xe0; out[25] = 0xe0; out[26] = 0xe6;
    return 27;
}

static void put_stamp(uint8_t lead, uint32_t v, uint8_t *o5)
{
    /* source/mpeg1_enc.c:59-71 */
    o5[0] = (uint8_t)(lead | ((v & 0xe0000000u) >> 28));
    o5[1] = (uint8_t)((v & 0x1fe00000u) >> 21);
    o5[2] = (uint8_t)(0x01 | ((v & 0x001fc000u) >> 13));
    o5[3] = (uint8_t)((v & 0x00003fc0u) >> 6);
    o5[4] = (uint8_t)(0x01 | ((v & 0x0000003fu) << 1));
}

int m1o_frame_prefix(long frame_index, int W, int H, int mode, long payload_bytes, uint8_t out[44])
{
    /* include/encoder.h:475-484: minute % 60 == 0 always holds, so hour++ after every frame and
     * second = minute = 0; hour is a uint8_t (:42). */
    uint32_t hour = (uint32_t)(frame_index & 0xff);
    uint32_t t = 1u + 3600u * hour;                                   /* include/encoder.h:200 */
    uint32_t dts = (uint32_t)((double)t * 1.2) + 0xbeefu;             /* source/mpeg1_enc.c:57-58 */
    uint32_t pts = dts - 0xbeefu;
    uint16_t len = (uint16_t)(44 + payload_bytes - 8);                /* include/encoder.h:448-454 */
    uint8_t *o = out;
    *o++ = 0; *o++ = 0; *o++ = 1; *o++ = 0xe0;
    *o++ = (uint8_t)(len >> 8); *o++ = (uint8_t)len;
    put_stamp(0x31, dts, o); o += 5;
    put_stamp(0x11, pts, o); o += 5;
    /* sequence header, source/mpeg1_enc.c:81-94; REF_COMPAT truncates the dimensions to uint8_t
     * first (include/encoder.h:186-187) */
    uint32_t w = (uint32_t)W, h = (uint32_t)H;
    if (mode == M1O_MODE_REF_COMPAT) { w &= 0xff; h &= 0xff; }
    w &= 0xffff; h &= 0xffff;
    *o++ = 0; *o++ = 0; *o++ = 1; *o++ = 0xb3;
    *o++ = (uint8_t)((w & 0xff0) >> 4);
    *o++ = (uint8_t)(((w & 0xf) << 4) | ((h & 0xf00) >> 8));
    *o++ = (uint8_t)(h & 0xff);
    *o++ = (uint8_t)(((1 & 0xf) << 4) | (4 & 0xf));                   /* aspect 1, frame rate 4 */
    *o++ = 0xff; *o++ = 0xff; *o++ = 0xe0; *o++ = (uint8_t)((3 & 0x1f) << 3);
    /* GOP, source/mpeg1_enc.c:103-113: drop 0, minute 0, second 0, num_pic 0, closed 1 */
    *o++ = 0; *o++ = 0; *o++ = 1; *o++ = 0xb8;
    *o++ = (uint8_t)((hour & 0x1f) << 2); *o++ = 0x08; *o++ = 0x00; *o++ = 0x40;
    /* picture, source/mpeg1_enc.c:120-137: temporal_ref 0, type 1, vbv_delay 0xffff */
    *o++ = 0; *o++ = 0; *o++ = 1; *o++ = 0x00;
    *o++ = 0x00; *o++ = 0x0f; *o++ = 0xff; *o++ = 0xf8;
    return 44;
}

long m1o_encode_stream(const uint8_t *frames, long n_frames, int W, int H, int channels, int mode,
                       int quality_factor, uint8_t *out, long cap)
{
    int32_t qm[64];
    m1o_qmatrix(quality_factor, qm);
    if (cap < 27) return -1;
    long pos = m1o_file_prologue(out);
    for (long f = 0; f < n_frames; ++f) {
        if (cap - pos < 44 + 4) return -1;
        long n = m1o_encode_picture(frames + f * (long)W * H * channels, W, H, channels, mode, qm,
                                    out + pos + 44, cap - pos - 44 - 4, NULL);
        if (n < 0) return n;
        m1o_frame_prefix(f, W, H, mode, n, out + pos);
        pos += 44 + n;
        /* include/encoder.h:456-458 writes 4 uninitialised bytes here; the evident intent
         * (source/mpeg1_enc.c:96-98) is the sequence end code.  Masked in comparisons. */
        out[pos++] = 0; out[pos++] = 0; out[pos++] = 1; out[pos++] = 0xb7;
    }
    return pos;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic input (ours).
 * ---------------------------------------------------------------------------------------- */
static uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

void m1o_synth_rgb(uint32_t seed, long frame_index, int W, int H, int kind, uint8_t *rgb)
{
    uint32_t f = (uint32_t)frame_index;
    uint32_t fkey = mix32(seed * 0x85ebca6bu + f * 0x9e3779b9u + 0x165667b1u);
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            uint32_t n = mix32(fkey ^ ((uint32_t)y * (uint32_t)W + (uint32_t)x));
            uint8_t *p = rgb + ((long)y * W + x) * 3;
            /* M1O_SYNTH_SCATTERED: the natural picture with a quarter of its 8x8 pixel tiles (chosen by a hash of the tile's
             * position and the frame) replaced by noise -- busy and flat blocks side by side inside every warp of the encoder */
            int noisy = kind == M1O_SYNTH_NOISE;
            if (kind == M1O_SYNTH_SCATTERED)
                noisy = (mix32(fkey ^ (0x51ed270bu + ((uint32_t)(y >> 3) * 0x9e3779b1u) + (uint32_t)(x >> 3))) & 3u) == 0u;
            if (noisy) {
                p[0] = (uint8_t)n; p[1] = (uint8_t)(n >> 8); p[2] = (uint8_t)(n >> 16);
            } else {
                p[0] = (uint8_t)((255u * (uint32_t)x / (uint32_t)W + (n & 15u) + f) & 255u);
                p[1] = (uint8_t)((255u * (uint32_t)y / (uint32_t)H + ((n >> 4) & 15u)) & 255u);
                p[2] = (uint8_t)((((uint32_t)x + (uint32_t)y) / 8u + ((n >> 8) & 15u) + 2u * f) & 255u);
                /* worst cases of an integer colour path (every pixel an exact-quotient exception of
                 * source/image_processing.c:104-106): grey pictures, and pictures with r == g */
                if (kind == M1O_SYNTH_GREY) { p[1] = p[0]; p[2] = p[0]; }
                else if (kind == M1O_SYNTH_RG_EQUAL) { p[1] = p[0]; }
            }
        }
    }
}
