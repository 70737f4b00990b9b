// m1cu_block.cuh -- the per-8x8-block arithmetic of the sm_100a encode kernels: forward DCT,
// coefficient packing + non-zero test, quantisation of the coded coefficients, DC/AC VLC.
//
// Everything here is `__host__ __device__`: the kernels inline it, and tests/host/block_host.cu
// compiles the very same functions for the host so the CPU test suite can compare them with the
// oracle block by block (tests/test_block_host.py).  That host build is a test harness, not a
// fallback: nothing in libm1cu.so calls these functions on the host.
#pragma once
#include "m1cu_common.cuh"

#define M1_HD __host__ __device__ __forceinline__

// ---- intrinsics with host stand-ins -------------------------------------------------------------
M1_HD int m1_clz(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
M1_HD int m1_ffs(uint32_t x)                 // 1-based index of the lowest set bit, 0 if none
{
#ifdef __CUDA_ARCH__
    return __ffs((int)x);
#else
    return __builtin_ffs((int)x);
#endif
}
M1_HD uint32_t m1_umulhi(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((unsigned long long)a * b) >> 32);
#endif
}
M1_HD uint32_t m1_funnel_l(uint32_t lo, uint32_t hi, uint32_t s)   // high word of (hi:lo) << (s & 31)
{
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}
// lo | hi << 16 for two values below 2^16 as ONE byte permute (the compiler cannot know the range and would shift + or)
M1_HD uint32_t m1_pack16(uint32_t lo, uint32_t hi)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(lo, hi, 0x5410);
#else
    return (lo & 0xffffu) | (hi << 16);
#endif
}
// a + b + c and a + b - c as ONE IADD3 each.  Written as two PTX adds with a private temporary: in C
// the compiler would share (a + b) between the sum and the difference and spend three instructions
// on the pair instead of two.
M1_HD int m1_add3(int a, int b, int c)
{
#ifdef __CUDA_ARCH__
    int d;
    asm("{\n\t.reg .s32 t;\n\tadd.s32 t, %1, %2;\n\tadd.s32 %0, t, %3;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return a + b + c;
#endif
}
M1_HD int m1_addsub3(int a, int b, int c)
{
#ifdef __CUDA_ARCH__
    int d;
    asm("{\n\t.reg .s32 t;\n\tadd.s32 t, %1, %2;\n\tsub.s32 %0, t, %3;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return a + b - c;
#endif
}

// -------------------------------------------------------------------------------------------
// 8-point integer butterfly of fast_DCT (source/image_processing.c:210-238).  On return
//   t0 = k0 sum, t1 = k4 diff, t2 = k2 (unshifted), t3 = k6 (unshifted),
//   k1 = p31 + t5, k7 = p31 - t5 with p31 = p3 + p1 left to the caller as the pair (p3, p1) so the
//   sum and the difference are one three-input add each; t6 -> k3, t7 -> k5 (before the r2 scaling).
// 41 instructions for a row, 42 for a column (11 multiplies); the stage-1 sums a0 = x0 + x7 and
// a1 = x1 + x6 are never formed, they only occur inside three-input adds.
// -------------------------------------------------------------------------------------------
struct Fdct8 { int t0, t1, t2, t3, p3, p1, t5, t6, t7; };

M1_HD Fdct8 fdct_core(int x0, int x1, int x2, int x3, int x4, int x5, int x6, int x7)
{
    const int c1 = 1004, s1 = 200, c3 = 851, s3 = 569, r2c6 = 554, r2s6 = 1337;
    const int d0 = x0 - x7, d1 = x1 - x6;
    const int a2 = x2 + x5, d2 = x2 - x5;
    const int a3 = x3 + x4, d3 = x3 - x4;
    const int e0 = m1_add3(x0, x7, a3), e3 = m1_addsub3(x0, x7, a3);
    const int e1 = m1_add3(x1, x6, a2), e2 = m1_addsub3(x1, x6, a2);
    const int m12 = c1 * (d1 + d2);
    const int p2 = (-s1 - c1) * d2 + m12;
    const int p1 = (s1 - c1) * d1 + m12;
    const int m03 = c3 * (d0 + d3);
    const int p3 = (-s3 - c3) * d3 + m03;
    const int p0 = (s3 - c3) * d0 + m03;
    const int m78 = r2c6 * (e2 + e3);
    Fdct8 o;
    o.t0 = e0 + e1;
    o.t1 = e0 - e1;
    o.t2 = (r2s6 - r2c6) * e3 + m78;
    o.t3 = (-r2s6 - r2c6) * e2 + m78;
    o.t5 = p0 + p2;
    o.t7 = p0 - p2;
    o.p3 = p3;
    o.p1 = p1;
    o.t6 = p3 - p1;
    return o;
}

// Forward DCT of one 8x8 block held in registers: v[i*8+j] in, dct[u*8+v] + M1_COEF_BIAS out (in
// place).  Row pass source/image_processing.c:198-250, column pass :253-305.  The bias (2048, folded
// into the rounding constants of the final shifts, so it is free and exact: (a + b*2^s) >> s ==
// (a >> s) + b) makes every output a non-negative 12-bit number, which lets two coefficients share a
// 32-bit word without sign trouble (see pack_and_flag).
// (Measured on the SASS and dropped: (x * r2) >> 17 as one IMAD.HI -- ptxas then folds the consumers'
// adds into the IMAD.HI's 64-bit addend and pays three moves per fold; the column pass's last
// multiply-shift as IMAD.WIDE -- register pairs cost more moves than the shifts they save.)
#define M1_COEF_BIAS 2048
// One row of the row pass: x0..x7 in, the eight row-pass outputs out[0..7] (source/image_processing.c:198-250).
M1_HD void fdct_row(int x0, int x1, int x2, int x3, int x4, int x5, int x6, int x7, int (&out)[8])
{
    const int r2 = 181;
    Fdct8 o = fdct_core(x0, x1, x2, x3, x4, x5, x6, x7);
    out[0] = o.t0;
    out[4] = o.t1;
    out[2] = o.t2 >> 10;
    out[6] = o.t3 >> 10;
    out[7] = m1_addsub3(o.p3, o.p1, o.t5) >> 10;
    out[1] = m1_add3(o.p3, o.p1, o.t5) >> 10;
    out[3] = (o.t6 * r2) >> 17;
    out[5] = (o.t7 * r2) >> 17;
}
// One column of the column pass: the row-pass outputs of one column in, the BIASED coefficients out[u] of that column
// (source/image_processing.c:253-305).
M1_HD void fdct_col(int x0, int x1, int x2, int x3, int x4, int x5, int x6, int x7, int (&out)[8])
{
    const int r2 = 181;
    Fdct8 o = fdct_core(x0, x1, x2, x3, x4, x5, x6, x7);
    const int t4 = o.p3 + o.p1;
    out[0] = (o.t0 + (16 + (M1_COEF_BIAS << 3))) >> 3;
    out[4] = (o.t1 + (16 + (M1_COEF_BIAS << 3))) >> 3;
    out[2] = (o.t2 + (16384 + (M1_COEF_BIAS << 13))) >> 13;
    out[6] = (o.t3 + (16384 + (M1_COEF_BIAS << 13))) >> 13;
    out[7] = (t4 - o.t5 + (16384 + (M1_COEF_BIAS << 13))) >> 13;
    out[1] = (t4 + o.t5 + (16384 + (M1_COEF_BIAS << 13))) >> 13;
    out[3] = ((o.t6 >> 8) * r2 + (8192 + (M1_COEF_BIAS << 12))) >> 12;
    out[5] = ((o.t7 >> 8) * r2 + (8192 + (M1_COEF_BIAS << 12))) >> 12;
}
M1_HD void fdct8x8(int (&v)[64])
{
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int o[8];
        fdct_row(v[i * 8 + 0], v[i * 8 + 1], v[i * 8 + 2], v[i * 8 + 3], v[i * 8 + 4], v[i * 8 + 5], v[i * 8 + 6], v[i * 8 + 7], o);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[i * 8 + k] = o[k];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int o[8];
        fdct_col(v[0 * 8 + j], v[1 * 8 + j], v[2 * 8 + j], v[3 * 8 + j], v[4 * 8 + j], v[5 * 8 + j], v[6 * 8 + j], v[7 * 8 + j], o);
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u * 8 + j] = o[u];
    }
}

// zigzag rank of raster position k (source/image_processing.c:28-37), compile-time table
__host__ __device__ constexpr int zz_rank(int k)
{
    constexpr int t[64] = { 0,  1,  5,  6, 14, 15, 27, 28,  2,  4,  7, 13, 16, 26, 29, 42,
                            3,  8, 12, 17, 25, 30, 41, 43,  9, 11, 18, 24, 31, 40, 44, 53,
                           10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
                           21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63 };
    return t[k];
}
__host__ __device__ constexpr int zz_raster(int z)
{
    for (int k = 0; k < 64; ++k) if (zz_rank(k) == z) return k;
    return 0;
}

// Coefficient record of thread t: 32 words at byte offset t*256, each holding two BIASED
// coefficients as 16-bit lanes: word w = (z & 15) + 16*(z >> 5) carries zigzag position z in its low
// lane when bit 4 of z is clear, in its high lane otherwise (pairs (z, z+16)).  16-byte groups are
// XOR-swizzled by `key` (k_encode_chunks: the owning thread's index, see blk_key; the experimental
// kernels: t itself).  rec_index returns the index in shorts.
// kStride = distance between records in shorts (128 when the record aliases the block's plane
// memory, 64 for the dense record array of the warp-specialised kernel).
template <int kStride = 128>
M1_HD int rec_index(int t, int z, int key)
{
    const int w = (z & 15) + ((z >> 5) << 4);
    return t * kStride + ((((w >> 2) ^ key) & 7) << 3) + ((w & 3) << 1) + ((z >> 4) & 1);
}
template <int kStride = 128>
M1_HD int rec_index(int t, int z) { return rec_index<kStride>(t, z, t); }

// Pack pairs (z, z+16) of biased coefficients and test both lanes at once:
//   lane + (0x7800 - m) has bit 15 set  <=>  c >=  m
//   (0x8800 - m) - lane has bit 15 set  <=>  c <= -m          (m = scaled matrix entry, no
// carries cross the lanes: every lane value stays inside [0, 0xffff]), so level != 0 <=> either.
// Shifting the accumulator right once per word lands word i's flags on bits i and 16+i.
// pk[w] = the 32 record words; returns the 64-bit zigzag-order non-zero mask.  Keys: any type with
// ka[32] / kb[32] (the kernel-parameter constant bank in k_encode_chunks, shared tables elsewhere).
template <class Keys>
M1_HD unsigned long long pack_and_flag(const int (&v)[64], uint32_t (&pk)[32], const Keys &nk)
{
    uint32_t half[2];
#pragma unroll
    for (int hblk = 0; hblk < 2; ++hblk) {
        uint32_t fl = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int w = hblk * 16 + i, z = hblk * 32 + i;
            const uint32_t p = m1_pack16((uint32_t)v[zz_raster(z)], (uint32_t)v[zz_raster(z + 16)]);
            pk[w] = p;
            const uint32_t f = ((p + nk.ka[w]) | (nk.kb[w] - p)) & 0x80008000u;
            fl = f + (fl >> 1);
        }
        half[hblk] = fl;
    }
    return ((unsigned long long)half[1] << 32) | half[0];
}

// -------------------------------------------------------------------------------------------
// Bit sink for the block coder: the first 64 bits of a block in a register pair (typical blocks are
// 5..40 bits).  put() appends at the LOW end (shift left, or in the code: four instructions, no
// branches); finish() left-aligns once per block, after which hi:lo hold the block's bits MSB-first.
// n keeps counting past 64 so the length is always exact; the bits are only valid while n <= 64.
// len is 1..31.
// -------------------------------------------------------------------------------------------
struct BitAcc {
    uint32_t hi, lo;
    int n;
    M1_HD void put(uint32_t code, int len)
    {
        hi = m1_funnel_l(lo, hi, (uint32_t)len);
        lo = (lo << len) | code;
        n += len;
    }
    M1_HD void finish()
    {
        const uint32_t s = (uint32_t)(64 - n);                 // only meaningful for 0 < n <= 64
        const uint32_t th = m1_funnel_l(lo, hi, s), tl = lo << (s & 31);
        if (s & 32) { hi = tl; lo = 0; } else { hi = th; lo = tl; }
    }
};

// Quantised level of zigzag position z from the biased DCT coefficient: C truncating division by the
// scaled matrix entry (source/image_processing.c:367).  |level| = floor(|c| / m) is the high word of
// (2|c| + 1) * ceil(2^31 / m) (checked exhaustively for |c| <= 2047 when the context is created,
// m1cu_quant.h).
M1_HD uint32_t quant_mag_biased(int biased, int z, const M1Tables *tb)
{
    const int t = 2 * biased - 2 * M1_COEF_BIAS;               // 2c
    return m1_umulhi((uint32_t)(t < 0 ? -t : t) + 1u, tb->qrcp[z]);
}
M1_HD int quant_level(int biased, int z, const M1Tables *tb)
{
    const int mag = (int)quant_mag_biased(biased, z, tb);
    return biased < M1_COEF_BIAS ? -mag : mag;
}

// Byte offset of zigzag position z inside an unswizzled coefficient record (rec_index with key 0,
// times two); a swizzled record XORs it with key << 4.  M1Tables.zofs holds this table so that the
// coder's dynamic index is one byte load and one XOR.
__host__ __device__ constexpr int rec_byte_offset(int z)
{
    return 2 * (((((z & 15) + ((z >> 5) << 4)) >> 2) << 3) + ((z & 3) << 1) + ((z >> 4) & 1));
}

// One block's bits: DC (source/mpeg1_blk.c:67-113), AC walk (source/image_processing.c:400-433,
// source/vlc.c:315-385), end of block (source/mpeg1_blk.c:115-117).  rec: the shared coefficient
// records; nz: bit z set <=> quantised level z is non-zero.  Returns non-zero when a coded AC
// level is outside the reference's encodable range.
// The AC walk is one short loop body with a single put(): tb->acrun[r] = first table entry | entries
// << 8 of run r (0 entries for r >= 32, so those escape), and entry 0 of run 0 -- which the
// reference never reaches, it special-cases (run 0, |level| 1) as '11' (source/vlc.c:330) -- holds
// that '11'.
template <int kStride = 128, class Sink>
M1_HD int code_block(Sink &s, const short *rec, int tid, unsigned long long nz,
                     bool is_luma, const M1Tables *tb, int key = -1)
{
    if (key < 0) key = tid;                                   // record swizzled by its own index
    int bad = 0;
    int prev = -1;
    unsigned long long m = nz;
    // this thread's record, and the swizzle applied to the byte offsets of tb->zofs
    const unsigned char *myrec = (const unsigned char *)(rec + tid * kStride);
    const uint32_t ksw = (uint32_t)(key & 7) << 4;
    if (nz & 1ull) {
        const int b0 = *(const unsigned short *)(myrec + (tb->zofs[0] ^ ksw));
        const int d = b0 - M1_COEF_BIAS;
        int c = (int)quant_mag_biased(b0, 0, tb);
        const int low = c & 0xff;
        const int sz = low ? 32 - m1_clz((uint32_t)low) : 1;   // highest set bit of bits 0..7, default 1
        const uint32_t e = tb->dc[sz + (is_luma ? 0 : 9)];
        if (d < 0) c ^= 1 << (sz - 1);
        const uint32_t val = (uint32_t)c & ((1u << sz) - 1u);
        s.put(((e & 0xffffffu) << sz) | val, (int)(e >> 24) + sz);
        prev = 0;
        m &= ~1ull;
    } else {
        if (is_luma) s.put(4u, 3); else s.put(0u, 2);
    }
    if ((nz >> 1) == 0ull) {                                  // no AC level at all (every flat block): end of block
        s.put(2u, 2);
        return 0;
    }
    // coding stops at the first non-zero whose predecessor position is also non-zero
    const unsigned long long adj = nz & (nz << 1);
    if (adj) m &= (adj & (0ull - adj)) - 1ull;
#pragma unroll 1
    for (int zb = 0; zb < 64; zb += 32) {
        uint32_t mm = zb ? (uint32_t)(m >> 32) : (uint32_t)m;
#pragma unroll 1
        while (mm) {
            const int k = zb + m1_ffs(mm) - 1;
            mm &= mm - 1u;
            const int b = *(const unsigned short *)(myrec + (tb->zofs[k] ^ ksw));
            const int d = b - M1_COEF_BIAS;
            const uint32_t mag = quant_mag_biased(b, k, tb);
            const int r = k - prev - 2;                          // run - 1 of source/vlc.c:326, 0..61
            prev = k;
            const uint32_t fw = tb->acrun[r];
            uint32_t code;
            int len;
            if (mag - 1u < (fw >> 8)) {
                const uint32_t e = tb->ac[(fw & 0xffu) + mag - 1u];
                code = e & 0xffffffu;
                len = (int)(e >> 24);
            } else {
                if (mag >= 256u) bad = 1;                        // reference: NULL -> crash
                const uint32_t L = d < 0 ? 0u - mag : mag;       // two's complement level
                const uint32_t head = (1u << 6) | (uint32_t)(r & 0x3f);    // 000001 rrrrrr
                if (mag < 128u) { code = (head << 8) | (L & 0xffu); len = 20; }
                else            { code = (head << 16) | (d < 0 ? 0x8000u : 0u) | (L & 0xffu); len = 28; }
            }
            s.put(code, len);
        }
    }
    s.put(2u, 2);
    return bad;
}
