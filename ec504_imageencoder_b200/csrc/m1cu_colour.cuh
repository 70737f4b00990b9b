// m1cu_colour.cuh -- the colour arithmetic of the sm_100a encode kernels: the reference's double
// chain restated bit for bit (exact path) and the integer fast path with its exception flag.
//
// `__host__ __device__` like m1cu_block.cuh: tests/host/block_host.cu compiles the same functions
// for the host and tests/test_block_host.py sweeps all 2^24 colours (every pixel alignment of the
// packed input) against the double chain.  The host build is a test harness, not a fallback.
//
// Reference (source/image_processing.c:104-106), evaluated in IEEE double, left to right, products
// rounded separately (gcc -O0, SSE2, no FMA), then truncated:
//   Y  = (uchar)(0.299 r + 0.587 g + 0.114 b)
//   Cb = (uchar)(128 - 0.168736 r - 0.331264 g + 0.5 b)
//   Cr = (uchar)(128 + 0.5 r - 0.418688 g - 0.081312 b)
//
// Integer fast path.  The exact values are rationals,
//   Y  = t  / 1000,   t  = 299 r + 587 g + 114 b                       (0 .. 255000)
//   Cb = Nb / 31250,  Nb = 4000000 - 5273 r - 10352 g + 15625 b        (15625 .. 7984375)
//   Cr = Nr / 31250,  Nr = 4000000 + 15625 r - 13084 g - 2541 b        (15625 .. 7984375)
// and the double chain's accumulated rounding error (< 2^-40) is far smaller than the distance of a
// non-integer quotient to the next integer (>= 1/31250), so the reference's truncation equals the
// floor of the exact quotient EXCEPT where the quotient is an integer: there the double sum may land
// one ulp low (3464 / 942 / 2706 of the 2^24 colours).
// Each numerator is two IDP.2A straight from the packed pixel bytes (16-bit coefficients x unsigned
// bytes), each quotient one IMAD.WIDE by ceil(2^32 / D): the high word is the floor unless the
// fraction is within eps = N * (ceil(2^32/D) * D - 2^32) / (D * 2^32) of the next integer, and the LOW
// word is the fraction scaled by 2^32.  A low word below M1_COLOUR_FLAG_LIMIT therefore marks every
// case in which the high word may differ from the reference: exact multiples (low word = q * 704 or
// q * 1454 <= 370770) and the fractions eps-close to 1 that the short reciprocal rounds up (they
// wrap to a small low word as well).  Unflagged quotients are exact; flagged 2x2 pixel quads are
// recomputed with the double chain by the kernel's fix-up pass.
#pragma once
#include "m1cu_common.cuh"

#ifndef M1_HD
#define M1_HD __host__ __device__ __forceinline__
#endif

#define M1_COLOUR_FLAG_LIMIT 0x80000u          // 2^19 > 370770 (see above)
#define M1_RCP_1000  4294968u                  // ceil(2^32 / 1000):  * 1000  - 2^32 = 704
#define M1_RCP_31250 137439u                   // ceil(2^32 / 31250): * 31250 - 2^32 = 1454

// ---- exact path: the reference's double arithmetic ---------------------------------------------
#ifdef __CUDACC__
// The seven non-trivial coefficients live in constant memory so the FP64 instructions take them as
// constant-bank operands instead of re-materialising 64-bit immediates through uniform registers.
__constant__ double kYcc[7] = { 0.299, 0.587, 0.114, 0.168736, 0.331264, 0.418688, 0.081312 };
// int -> double and double -> int go through the 2^52 mantissa trick instead of I2F/F2I (those run
// on the 16-lane/SM XU pipe): 2^52 + v holds v in its low word, and adding 2^52 with round-toward-
// zero leaves trunc(x) there (x >= 0 always holds here: Y >= 0, Cb, Cr >= 0.5).  0.5*x is exact, so
// fma(0.5, x, acc) rounds once exactly like the reference's separate multiply and add.
__device__ __forceinline__ double u8_to_double(int v)
{
    return __dsub_rn(__hiloint2double(0x43300000, v), 4503599627370496.0);
}
__device__ __forceinline__ int trunc_nonneg(double x)
{
    return __double2loint(__dadd_rz(x, 4503599627370496.0));
}
__device__ __forceinline__ int luma_from_doubles(double rd, double gd, double bd)
{
    double t = __dadd_rn(__dmul_rn(kYcc[0], rd), __dmul_rn(kYcc[1], gd));
    t = __dadd_rn(t, __dmul_rn(kYcc[2], bd));
    return trunc_nonneg(t);
}
__device__ __forceinline__ void chroma_from_doubles(double rd, double gd, double bd, int &cb, int &cr)
{
    double u = __dsub_rn(128.0, __dmul_rn(kYcc[3], rd));
    u = __dsub_rn(u, __dmul_rn(kYcc[4], gd));
    u = __fma_rn(0.5, bd, u);
    cb = trunc_nonneg(u);
    double v = __fma_rn(0.5, rd, 128.0);
    v = __dsub_rn(v, __dmul_rn(kYcc[5], gd));
    v = __dsub_rn(v, __dmul_rn(kYcc[6], bd));
    cr = trunc_nonneg(v);
}
// Chroma of one pixel ADDED to running sums kept in the 2^52 form (see convert_half_tile_exact): accb / accr are
// 2^52 + (sum so far); RZ(u + acc) = acc + trunc(u) exactly because u >= 0 and the result is an integer-spaced double.
__device__ __forceinline__ void chroma_accumulate(double rd, double gd, double bd, double &accb, double &accr)
{
    double u = __dsub_rn(128.0, __dmul_rn(kYcc[3], rd));
    u = __dsub_rn(u, __dmul_rn(kYcc[4], gd));
    u = __fma_rn(0.5, bd, u);
    accb = __dadd_rz(u, accb);
    double v = __fma_rn(0.5, rd, 128.0);
    v = __dsub_rn(v, __dmul_rn(kYcc[5], gd));
    v = __dsub_rn(v, __dmul_rn(kYcc[6], bd));
    accr = __dadd_rz(v, accr);
}
__device__ __forceinline__ void ycbcr_from_doubles(double rd, double gd, double bd, int &y, int &cb, int &cr)
{
    y = luma_from_doubles(rd, gd, bd);
    chroma_from_doubles(rd, gd, bd, cb, cr);
}
__device__ __forceinline__ void ycbcr_exact(int r, int g, int b, int &y, int &cb, int &cr)
{
    ycbcr_from_doubles(u8_to_double(r), u8_to_double(g), u8_to_double(b), y, cb, cr);
}
// byte B of w as a double: one I2F.F64.U8 with a byte selector (XU pipe), no extraction ALU op
__device__ __forceinline__ double byte_to_double(uint32_t w, int b)   // b is a constant after unrolling
{
    double d;
    const uint32_t s = w >> (8 * b);
    asm("cvt.rn.f64.u8 %0, %1;" : "=d"(d) : "r"(s));
    return d;
}
#endif
// host restatement for the test harness (tests/host/block_host.cu): the same chain in plain doubles,
// every intermediate forced through memory so that no FMA contraction or excess precision can occur
inline void ycbcr_exact_host(int r, int g, int b, int &y, int &cb, int &cr)
{
    volatile double p0 = 0.299 * r, p1 = 0.587 * g, p2 = 0.114 * b;
    volatile double t = p0 + p1; t = t + p2;
    y = (int)(unsigned char)t;
    volatile double q0 = 0.168736 * r, q1 = 0.331264 * g, q2 = 0.5 * b;
    volatile double u = 128 - q0; u = u - q1; u = u + q2;
    cb = (int)(unsigned char)u;
    volatile double s0 = 0.5 * r, s1 = 0.418688 * g, s2 = 0.081312 * b;
    volatile double v = 128 + s0; v = v - s1; v = v - s2;
    cr = (int)(unsigned char)v;
}

// ---- integer fast path ---------------------------------------------------------------------------
// dp2a: d = c + a.h0 * b.byte[sel] + a.h1 * b.byte[sel + 1], a = two signed 16-bit halves, b = four
// unsigned bytes, sel = 0 (lo) or 2 (hi).
M1_HD int m1_dp2a_lo(uint32_t a, uint32_t b, int c)
{
#ifdef __CUDA_ARCH__
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return c + (int)(short)(a & 0xffffu) * (int)(b & 0xffu) + (int)(short)(a >> 16) * (int)((b >> 8) & 0xffu);
#endif
}
M1_HD int m1_dp2a_hi(uint32_t a, uint32_t b, int c)
{
#ifdef __CUDA_ARCH__
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return c + (int)(short)(a & 0xffffu) * (int)((b >> 16) & 0xffu) + (int)(short)(a >> 16) * (int)(b >> 24);
#endif
}
M1_HD void m1_mul_wide(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
#ifdef __CUDA_ARCH__
    unsigned long long p;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(b));
    hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
#else
    const unsigned long long p = (unsigned long long)a * b;
    hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
#endif
}

#define M1_PK16(h0, h1) ((uint32_t)(((uint32_t)(uint16_t)(short)(h1) << 16) | (uint32_t)(uint16_t)(short)(h0)))

// Coefficient pairs of component comp (0 = Y, 1 = Cb, 2 = Cr) for the four ways a 3-byte pixel can
// sit in 32-bit words: [0] (cr, cg)  [1] (cb, 0)  [2] (0, cr)  [3] (cg, cb).
struct M1ColourK { uint32_t k[3][4]; int base[3]; };
__host__ __device__ constexpr M1ColourK m1_colour_k()
{
    return M1ColourK{ {
        { M1_PK16(299, 587),      M1_PK16(114, 0),    M1_PK16(0, 299),    M1_PK16(587, 114) },
        { M1_PK16(-5273, -10352), M1_PK16(15625, 0),  M1_PK16(0, -5273),  M1_PK16(-10352, 15625) },
        { M1_PK16(15625, -13084), M1_PK16(-2541, 0),  M1_PK16(0, 15625),  M1_PK16(-13084, -2541) } },
        { 0, 4000000, 4000000 } };
}

// Numerator of component comp for the pixel whose first byte is byte `sh` (0..3) of word w0; w1 is
// the following word (only read when the pixel straddles: sh >= 2).  kCh = 3: any sh; kCh = 4: sh = 0.
template <int comp>
M1_HD int colour_numer(uint32_t w0, uint32_t w1, int sh)
{
    constexpr M1ColourK K = m1_colour_k();
    switch (sh) {
    case 0:  return m1_dp2a_hi(K.k[comp][1], w0, m1_dp2a_lo(K.k[comp][0], w0, K.base[comp]));   // r g b .
    case 1:  return m1_dp2a_hi(K.k[comp][3], w0, m1_dp2a_lo(K.k[comp][2], w0, K.base[comp]));   // . r g b
    case 2:  return m1_dp2a_lo(K.k[comp][1], w1, m1_dp2a_hi(K.k[comp][0], w0, K.base[comp]));   // . . r g | b
    default: return m1_dp2a_lo(K.k[comp][3], w1, m1_dp2a_hi(K.k[comp][2], w0, K.base[comp]));   // . . . r | g b
    }
}

// One pixel: the three truncated quotients and the smallest of the three low words (the exception
// flag: *flag_min < M1_COLOUR_FLAG_LIMIT <=> some component must be recomputed exactly).
M1_HD void colour_int_pixel(uint32_t w0, uint32_t w1, int sh, int &y, int &cb, int &cr, uint32_t &flag_min)
{
    uint32_t hy, ly, hb, lb, hr, lr;
    m1_mul_wide((uint32_t)colour_numer<0>(w0, w1, sh), M1_RCP_1000, hy, ly);
    m1_mul_wide((uint32_t)colour_numer<1>(w0, w1, sh), M1_RCP_31250, hb, lb);
    m1_mul_wide((uint32_t)colour_numer<2>(w0, w1, sh), M1_RCP_31250, hr, lr);
    y = (int)hy; cb = (int)hb; cr = (int)hr;
    const uint32_t m = ly < lb ? ly : lb;
    const uint32_t m2 = m < lr ? m : lr;
    flag_min = flag_min < m2 ? flag_min : m2;
}
