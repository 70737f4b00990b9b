// m1cu_kernels.cu -- hand-written sm_100a kernels for the MPEG-1 I-frame per-block path.
// See m1cu_common.cuh for the pipeline and include/m1cu.h for the reference interfaces replaced.
#include "m1cu_common.cuh"
#include "m1cu_kernels.h"
#include "m1cu_block.cuh"
#include "m1cu_colour.cuh"
#include "m1cu_quant.h"
#ifdef M1_EXPERIMENTS
#include "../../tools/experiments/m1x_env.h"
#include "../../tools/experiments/m1x_trace.cuh"
#else
#define M1X_MARK(slot) do { } while (0)
#endif

// -------------------------------------------------------------------------------------------
// Shared-memory plane layout.  Samples are int32, block-major: block `blk` owns 64 words; its
// sixteen 16-byte chunks (chunk i = row*2 + (col>>2)) are XOR-swizzled with a per-block key so
// that both the producers (8x2-pixel colour half-tiles, 128-bit stores) and the consumers (one thread
// per block, 128-bit loads) are bank-conflict free.
//   luma blocks:   blk = by * 2C + bc      (by = block row 0/1 of the macroblock row, bc = 8-pixel column)
//   chroma blocks: blk = 4C + mb (Cb), 5C + mb (Cr)                       C = chunk_mbs
// Exactly one consumer thread reads each block.  After it has pulled the block into registers it
// reuses the first 128 bytes of the same 256 bytes for the block's DCT-coefficient record.
// The key of a block is the low three bits of its position in CODING order (t = 6 * mb + block, the
// consumer's thread index in k_encode_chunks): eight consecutive consumer threads -- one 128-bit
// shared-memory wavefront -- then hold eight different keys, and so do eight neighbouring colour
// strips (block columns 8j .. 8j+7 map to t & 7 = 0,1,6,7,4,5,2,3 (+2 for the lower block row)).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ int blk_key(int blk, int C)
{
    if (blk < 4 * C) {
        const int by = blk >= 2 * C, bc = blk - by * 2 * C;
        return (6 * (bc >> 1) + 2 * by + (bc & 1)) & 7;
    }
    const int cr = blk >= 5 * C;
    return (6 * (blk - (4 + cr) * C) + 4 + cr) & 7;
}
__device__ __forceinline__ int chunk_word_keyed(int blk, int i, int key)      // first word of chunk i of block blk
{
    return blk * 64 + (((i & 8) | ((i & 7) ^ key)) << 2);
}
__device__ __forceinline__ int chunk_word(int blk, int i, int C) { return chunk_word_keyed(blk, i, blk_key(blk, C)); }
__device__ __forceinline__ int plane_word(int blk, int r, int c, int C)
{
    return chunk_word(blk, r * 2 + (c >> 2), C) + (c & 3);
}
// -------------------------------------------------------------------------------------------
// Shared-memory bit sink for the block coder (the register sink BitAcc is in m1cu_block.cuh).
// -------------------------------------------------------------------------------------------
// Streams MSB-first into a shared-memory window of logical 32-bit words covering chunk bits
// [w0, w0 + 32*M1_WIN_WORDS); bits outside the window are dropped (another pass takes them).
struct WindowWriter {
    uint32_t *win;
    int pos;   // absolute bit position in the chunk
    int w0;
    int nw;    // window size in words (g.win_words)
    __device__ __forceinline__ void put(uint32_t code, int len)
    {
        const int p = pos - w0;
        pos += len;
        if (p + len <= 0 || p >= 32 * nw) return;
        const int word = p >> 5, o = p & 31;
        const unsigned long long sh = ((unsigned long long)code << (64 - len)) >> o;
        const uint32_t hi = (uint32_t)(sh >> 32), lo = (uint32_t)sh;
        if (hi && word >= 0 && word < nw) atomicOr(&win[word], hi);
        if (lo && word + 1 >= 0 && word + 1 < nw) atomicOr(&win[word + 1], lo);
    }
};

// A block longer than 64 bits, or one that straddles the end of the window: coded again straight into the window.
// Rare, and deliberately not inlined so that its set-up stays out of the kernel's common path.
__device__ __noinline__ void recode_into_window(uint32_t *win, int my_off, int w0, int nw, bool first_of_mb, const short *rec,
                                                int pb, unsigned long long nz, bool is_luma, const M1Tables *tb, int key)
{
    WindowWriter ww{win, my_off, w0, nw};
    if (first_of_mb) ww.put(3u, 2);                         // address increment '1' + macroblock_type '1'
    code_block(ww, rec, pb, nz, is_luma, tb, key);
}

// -------------------------------------------------------------------------------------------
// Colour tiles.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ const uint8_t *px_ptr(const uint8_t *frame, const M1Geom &g, int x, int y)
{
    x = min(x, g.W - 1);
    y = min(y, g.H - 1);
    return frame + ((size_t)y * g.W + x) * g.channels;
}

// Fast half-tile: 8 pixels x 2 rows = the 2x8 strip of luma block column `bc` of the chunk in rows
// 2*qy, 2*qy+1 of the macroblock row, plus the four 2x2 chroma means under it; everything in range
// and aligned.  64/128-bit loads, then the integer colour path of m1cu_colour.cuh straight from the
// packed bytes: per pixel six IDP.2A (numerators), three IMAD.WIDE (quotient in the high word, scaled
// fraction in the low word) and the running minimum of the low words of its 2x2 quad.  128-bit
// swizzled stores of int32 samples.  Returns the four quad minima folded into one flag word: bit q
// set <=> quad q (pixels 2q, 2q+1 of both rows) holds a pixel whose quotient the integer path cannot
// vouch for; the caller queues those quads for the exact (double) recomputation.
// Called from a runtime loop so the body exists once in the instruction stream.
template <int CH>
struct HalfTilePixels { uint32_t w[2][2 * CH]; };   // the raw bytes of 8 pixels x 2 rows, in registers

template <int CH>
__device__ __forceinline__ HalfTilePixels<CH> load_half_tile(const uint8_t *__restrict__ row0, size_t pitch)
{
    HalfTilePixels<CH> px;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const uint8_t *p = row0 + dy * pitch;
        if (CH == 4) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const uint4 a = __ldg((const uint4 *)p + i);
                px.w[dy][4 * i] = a.x; px.w[dy][4 * i + 1] = a.y; px.w[dy][4 * i + 2] = a.z; px.w[dy][4 * i + 3] = a.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const uint2 a = __ldg((const uint2 *)p + i);
                px.w[dy][2 * i] = a.x; px.w[dy][2 * i + 1] = a.y;
            }
        }
    }
    return px;
}

// Which pixels of a half-tile take which arithmetic (compile time; every setting is bit-exact):
//   2 (product): every pixel by the reference's double chain -- FP64 + conversion pipes, content independent.
//   0: every pixel by the integer path of m1cu_colour.cuh (IDP.2A / IMAD.WIDE on the FMA pipe, 18 instead of 28
//      issue slots per pixel), flagged 2x2 quads queued and recomputed exactly after the colour barrier.
//   1: row 0 integer, row 1 double chain (both pipe groups busy, half the flags).
// Measured on a B200 (profiles/r2_colour_variants.txt, 1080p x 300, k_encode_chunks ms): 2: 1.37, 0: 1.41, 1: 1.44
// on the default content; 0 doubles to 2.75 ms on grey or r == g pictures (every quad flagged).  The kernel is bound
// by latency at 7 resident CTAs per SM, not by the colour arithmetic, so the shorter
// instruction stream buys nothing and its fix-up pass costs; 0 and 1 are kept for tools/build_experiments.sh
// and tests/test_gpu_variants.py only.
#ifndef M1_COLOUR_SPLIT
#define M1_COLOUR_SPLIT 2
#endif
template <int CH>
__device__ __forceinline__ uint32_t convert_half_tile(const HalfTilePixels<CH> px, int bc, int qy, int C, int *__restrict__ planes)
{
    const uint32_t (&w)[2][2 * CH] = px.w;
    int sb[4], sr[4];
    uint32_t fmin[4] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu };
    // The four 16-byte luma groups of this strip are chunks i0 .. i0+3 of one block (i0 = 4*(qy & 3)):
    // chunk_word(blk, i0 + n) == a1 ^ (n << 2), one LOP3 per store instead of the full swizzle.
    const int by = qy >> 2, blk = by * 2 * C + bc, i0 = (qy & 3) << 2;
    const int a1 = chunk_word_keyed(blk, i0, (6 * (bc >> 1) + 2 * by + (bc & 1)) & 7);   // == blk_key(blk, C)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {                 // 4-pixel groups = 16-byte chunks
            int yv[4], cbv[4], crv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * j + e;              // pixel 0..7 of the row
                const int byte0 = CH * i, wi = byte0 >> 2, sh = byte0 & 3;
                int &cb = cbv[e], &cr = crv[e];
                if (M1_COLOUR_SPLIT == 2 || (M1_COLOUR_SPLIT == 1 && dy == 1))
                    ycbcr_from_doubles(byte_to_double(w[dy][byte0 >> 2], byte0 & 3), byte_to_double(w[dy][(byte0 + 1) >> 2], (byte0 + 1) & 3),
                                       byte_to_double(w[dy][(byte0 + 2) >> 2], (byte0 + 2) & 3), yv[e], cb, cr);
                else
                    colour_int_pixel(w[dy][wi], w[dy][wi + 1 < 2 * CH ? wi + 1 : wi], sh, yv[e], cb, cr, fmin[i >> 1]);
            }
            // 2x2 chroma sums as three-input adds (ALU pipe; the FMA pipe is the integer path's bottleneck)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int q = 2 * j + h2;
                if (dy == 0) { sb[q] = cbv[2 * h2] + cbv[2 * h2 + 1]; sr[q] = crv[2 * h2] + crv[2 * h2 + 1]; }
                else         { sb[q] = m1_add3(sb[q], cbv[2 * h2], cbv[2 * h2 + 1]); sr[q] = m1_add3(sr[q], crv[2 * h2], crv[2 * h2 + 1]); }
            }
            *(int4 *)(planes + (a1 ^ ((2 * dy + j) << 2))) = make_int4(yv[0], yv[1], yv[2], yv[3]);
        }
    }
    const int k = bc >> 1, h = bc & 1;                // macroblock, left/right half of its chroma row
    const int kc = (6 * k + 4) & 7;                   // key of the Cb block; the Cr block (thread + 1) has kc ^ 1
    *(int4 *)(planes + chunk_word_keyed(4 * C + k, qy * 2 + h, kc)) = make_int4(sb[0] >> 2, sb[1] >> 2, sb[2] >> 2, sb[3] >> 2);
    *(int4 *)(planes + chunk_word_keyed(5 * C + k, qy * 2 + h, kc ^ 1)) = make_int4(sr[0] >> 2, sr[1] >> 2, sr[2] >> 2, sr[3] >> 2);
    uint32_t flags = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) flags |= (fmin[q] < M1_COLOUR_FLAG_LIMIT ? 1u : 0u) << q;
    return flags;
}

template <int CH>
__device__ __forceinline__ uint32_t color_half_tile(const uint8_t *__restrict__ row0, size_t pitch, int bc, int qy,
                                                    int C, int *__restrict__ planes)
{
    return convert_half_tile<CH>(load_half_tile<CH>(row0, pitch), bc, qy, C, planes);
}

// The product's half-tile (M1_COLOUR_SPLIT == 2): every pixel by the reference's double chain.  Two things keep
// the non-FP64 instruction count down:
//  * the 2x2 chroma sums are formed by the truncating adds themselves: trunc(x) is read from the low word of
//    RZ(x + 2^52), and RZ(x' + (2^52 + s)) = 2^52 + s + trunc(x') exactly (the sum is an integer-spaced double,
//    x' >= 0), so chaining the four pixels of a quad through one accumulator leaves cb0 + cb1 + cb2 + cb3 in the
//    low word with no integer add at all;
//  * the three swizzled shared-memory addresses come in precomputed (a1: the strip's first luma chunk of this
//    half-tile, acb / acr: its chroma chunk): the caller derives the second half-tile's from the first by one XOR.
template <int CH>
__device__ __forceinline__ void convert_half_tile_exact(const HalfTilePixels<CH> px, int a1, int acb, int acr,
                                                        int *__restrict__ planes)
{
    const uint32_t (&w)[2][2 * CH] = px.w;
    int sb[4], sr[4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {                     // 4-pixel groups = 16-byte luma chunks, two 2x2 quads each
        int yv[2][4];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            double accb = 4503599627370496.0, accr = 4503599627370496.0;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int e = 2 * h2 + dx, byte0 = CH * (4 * j + e);
                    const double rd = byte_to_double(w[dy][byte0 >> 2], byte0 & 3);
                    const double gd = byte_to_double(w[dy][(byte0 + 1) >> 2], (byte0 + 1) & 3);
                    const double bd = byte_to_double(w[dy][(byte0 + 2) >> 2], (byte0 + 2) & 3);
                    yv[dy][e] = luma_from_doubles(rd, gd, bd);
                    chroma_accumulate(rd, gd, bd, accb, accr);
                }
            sb[2 * j + h2] = __double2loint(accb) >> 2;
            sr[2 * j + h2] = __double2loint(accr) >> 2;
        }
        *(int4 *)(planes + (a1 ^ (j << 2))) = make_int4(yv[0][0], yv[0][1], yv[0][2], yv[0][3]);
        *(int4 *)(planes + (a1 ^ ((2 + j) << 2))) = make_int4(yv[1][0], yv[1][1], yv[1][2], yv[1][3]);
    }
    *(int4 *)(planes + acb) = make_int4(sb[0], sb[1], sb[2], sb[3]);
    *(int4 *)(planes + acr) = make_int4(sr[0], sr[1], sr[2], sr[3]);
}

// One 2x2 pixel quad by the exact double chain: any alignment / channel count, coordinates clamped to
// the picture (= edge replication up to the coded size).  (x0, y0): top-left pixel of the half-tile,
// qx: quad 0..3 inside it.  Used by the generic half-tile and by the fix-up pass of the fast path.
__device__ __forceinline__ void color_quad_exact(const uint8_t *__restrict__ fr, const M1Geom &g, int x0, int y0,
                                                 int bc, int qy, int qx, int C, int *__restrict__ planes)
{
    int sb = 0, sr = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const uint8_t *p = px_ptr(fr, g, x0 + 2 * qx + dx, y0 + dy);
            int yy, cb, cr;
            ycbcr_exact(p[0], p[1], p[2], yy, cb, cr);
            const int rr = 2 * qy + dy, cc = 2 * qx + dx;
            planes[plane_word((rr >> 3) * 2 * C + bc, rr & 7, cc, C)] = yy;
            sb += cb; sr += cr;
        }
    planes[plane_word(4 * C + (bc >> 1), qy, 4 * (bc & 1) + qx, C)] = sb >> 2;
    planes[plane_word(5 * C + (bc >> 1), qy, 4 * (bc & 1) + qx, C)] = sr >> 2;
}

// Generic half-tile: any alignment / channel count, coordinates clamped to the picture (= edge
// replication up to the coded size).
__device__ __noinline__ void color_half_tile_generic(const uint8_t *__restrict__ fr, const M1Geom &g, int x0, int y0,
                                                     int bc, int qy, int C, int *__restrict__ planes)
{
    for (int qx = 0; qx < 4; ++qx) color_quad_exact(fr, g, x0, y0, bc, qy, qx, C, planes);
}

// -------------------------------------------------------------------------------------------
// k_encode_chunks: one CTA per (chunk, slice, picture).
// -------------------------------------------------------------------------------------------
// kLoad: 3 / 4 = FULL mode with aligned 3- / 4-byte pixels (fast half-tiles), 0 = FULL mode generic
// loads only, -1 = REF_COMPAT.  One instantiation per input format keeps each kernel's code small.
#ifndef M1_LANES8_MAX
#define M1_LANES8_MAX 24    // a warp with at most this many blocks that need a DCT (and at least one that does not) uses eight lanes per block
                            // (4 / 8 / 12 / 16 / 24 measured on scattered content: 1.544 / 1.569 / 1.593 / 1.521 / 1.518 ms, profiles/r2_variants_ab.txt)
#endif
#ifndef M1_ENC_MIN_CTAS
#define M1_ENC_MIN_CTAS 8   // 64 registers, 8 CTAs/SM (= what the shared memory allows).  The spills land in the one-thread-per-block
                            // DCT, which most warps no longer run: 8 beats 7 (72 registers) by 1.3 % on the default content and by 0.5 %
                            // on noise (profiles/r2_cta8_ab.txt); before the flat-block shortcut 8 lost 4 % to 7
#endif
// kFlat: the quality has a flat range (g.flat_range >= 0): 64 registers / 8 CTAs per SM.  Without one every block runs the
// one-thread DCT, which wants its 72 registers: 7 CTAs per SM as before.
template <int kLoad, bool kLevels, bool kFlat>
__global__ void __launch_bounds__(128, kFlat ? M1_ENC_MIN_CTAS : M1_ENC_MIN_CTAS - 1)
k_encode_chunks(const __grid_constant__ M1Geom g, const __grid_constant__ M1NzKeys nk,
                const uint8_t *__restrict__ rgb, const M1Tables *__restrict__ gtab,
                uint32_t *__restrict__ staging, uint32_t *__restrict__ chunk_bits,
                short *__restrict__ levels, int *__restrict__ err)
{
    extern __shared__ __align__(256) unsigned char smem[];   // 256: the eight-lane DCT places words by XOR on the byte address
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int chunk = blockIdx.x, slice = blockIdx.y, frame = blockIdx.z;
    const int C = g.chunk_mbs;
    const int mb0 = chunk * C;
    // Paired last chunks (g.pair_tails = T macroblocks each): the CTA of an even slice also encodes the last chunk
    // of the next slice -- macroblock slots 0 .. T-1 are slice `slice`, slots T .. 2T-1 slice `slice + 1`, each
    // with its own staging record; the odd slice's own CTA has nothing to do.
    const bool last_chunk = chunk == g.chunks_per_slice - 1;
    const int T = (kLoad >= 0 && last_chunk) ? g.pair_tails : 0;
    if (T && (slice & 1)) return;
    const bool paired = T && slice + 1 < g.slices;
    const int nmb = paired ? 2 * T : min(C, g.mbs_per_slice - mb0);
#ifdef M1_EXPERIMENTS
    const size_t m1x_lin = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (m1x_stagger_ns) {                                        // the k-th CTA to start on an SM waits k * stagger (first wave only)
        if (threadIdx.x == 0) {
            const unsigned k = atomicAdd(&m1x_sm_arrivals[m1x_smid()], 1u);
            if (k < (unsigned)m1x_stagger_ctas) {
                const unsigned long long t0 = m1x_now(), d = (unsigned long long)k * m1x_stagger_ns;
                while (m1x_now() - t0 < d) __nanosleep(200);
            }
        }
        __syncthreads();
    }
    M1X_MARK(0);
#endif

    // shared memory carve-up
    int *planes = (int *)smem;                                   // [6C blocks][64] int32, swizzled
    short *rec = (short *)smem;                                  // aliases planes (see layout note)
    uint32_t *win = (uint32_t *)(planes + 6 * C * 64);           // [M1_WIN_WORDS + 4]
    M1Tables *tb = (M1Tables *)(win + M1_WIN_WORDS + 4);         // 16-byte aligned
    int *wtot = (int *)(tb + 1);                                 // [64]: [0..3] bits per warp, [4..7] fix-up queue lengths, [8] pair split,
                                                                 // [16..63] per warp: (plane block, lane) of the blocks that need a DCT; 16-byte aligned

    // Per-CTA prologue, kept short: the coder's tables up to zofs[] in 128-bit pieces (the non-zero
    // keys come from the constant bank), and only the first blockDim.x window words zeroed -- a chunk
    // that needs more (> 32 * blockDim.x bits) zeroes the rest once its size is known.
    constexpr int kTabVecs = ((int)offsetof(M1Tables, ka) + 15) / 16;
#pragma unroll 1
    for (int i = tid; i < kTabVecs; i += nthr) ((uint4 *)tb)[i] = __ldg((const uint4 *)gtab + i);
    // Until the block phase needs it as the bit window, `win` holds the fix-up queues of the integer
    // colour path: one 16-bit entry per flagged 2x2 pixel quad, one queue of 256 entries per warp (a
    // warp's 32 strips hold 256 quads; 4 queues = the window's 512 words), counted in wtot[4 + warp].
    // Queue and counter are private to the warp until the barrier, so zeroing the counter needs no
    // CTA-wide barrier of its own.
#if M1_COLOUR_SPLIT != 2
    unsigned short *fixq = (unsigned short *)win + 256 * (tid >> 5);
    int *fix_cnt = wtot + 4 + (tid >> 5);
    if ((tid & 31) == 0) *fix_cnt = 0;
    if (tid < 4 && tid >= (nthr >> 5)) wtot[4 + tid] = 0;
    __syncwarp();
#endif
    if (tid < 4 && tid >= (nthr >> 5)) wtot[tid] = 0;           // bit totals of the warps this CTA does not have (the others write theirs)

    const uint8_t *fr = rgb + (size_t)frame * g.frame_stride;

    // ---- phase 1: colour conversion into the chunk's planes --------------------------------
#ifdef M1_EXPERIMENTS
    if (g.debug_skip & 1) {
        // profiling only (tools/ build): leave the planes as they are
    } else
#endif
    if (kLoad >= 0) {
        // FULL: slice = macroblock row.  A thread owns the 8-pixel x 4-row strip (block column bc of the
        // chunk, row quad q4) = two half-tiles, so the index split and the pixel address are computed once
        // and stepped by rows; 4 * nbc = 8 * nmb strips <= blockDim.x.
        const size_t pitch = (size_t)g.W * g.channels;
        const int nbc = 2 * nmb;
        constexpr int kCh = kLoad > 0 ? kLoad : 3;
        const unsigned inv = paired ? g.inv_nbc[2] : last_chunk ? g.inv_nbc[1] : g.inv_nbc[0];
        for (int st = tid; st < 4 * nbc; st += nthr) {
            const int q4 = (int)(((unsigned)st * inv) >> 16), bc = st - q4 * nbc;
            const int sub = (paired && bc >= 2 * T) ? 1 : 0;  // second slice of a pair
            const int x0 = 16 * mb0 + 8 * (bc - sub * 2 * T);
            int y = 16 * (slice + sub) + 4 * q4;
            if (kLoad > 0 && x0 + 8 <= g.W) {
                // rows below the picture replicate its last row (edge replication up to the coded size):
                // clamp the first row, then step by the pitch only while the next row exists
                const int last = g.H - 1;
                const uint8_t *row = fr + (size_t)min(y, last) * pitch + (size_t)x0 * kCh;
#if M1_COLOUR_SPLIT == 2
                // swizzled addresses of half-tile 0 (qy = 2 * q4): luma chunk i0 = 8 * (q4 & 1) of block (q4 >> 1, bc),
                // chroma chunk 4 * q4 + (bc & 1) of the macroblock's Cb / Cr block; half-tile 1 toggles chunk bit 2
                // (luma) / bit 1 (chroma), i.e. one XOR on the word address
                const int by = q4 >> 1, k = bc >> 1;
                const int a1 = chunk_word_keyed(by * 2 * C + bc, (q4 & 1) << 3, (6 * k + 2 * by + (bc & 1)) & 7);
                const int kc = (6 * k + 4) & 7;       // key of the Cb block; the Cr block (thread + 1) has kc ^ 1
                const int acb = chunk_word_keyed(4 * C + k, 4 * q4 + (bc & 1), kc);
                const int acr = chunk_word_keyed(5 * C + k, 4 * q4 + (bc & 1), kc ^ 1);
#pragma unroll 1
                for (int h = 0; h < 2; ++h, y += 2) {
                    const size_t rp = (y + 1 <= last) ? pitch : 0;
                    convert_half_tile_exact<kCh>(load_half_tile<kCh>(row, rp), a1 ^ (h << 4), acb ^ (h << 3), acr ^ (h << 3), planes);
                    row += rp + ((y + 2 <= last) ? pitch : 0);
                }
#else
#pragma unroll 1
                for (int h = 0; h < 2; ++h, y += 2) {
                    const size_t rp = (y + 1 <= last) ? pitch : 0;
                    const uint32_t fl = color_half_tile<kCh>(row, rp, bc, 2 * q4 + h, C, planes);
                    if (fl) {                                // rare: queue the flagged quads of this half-tile
                        const int slot = atomicAdd(fix_cnt, __popc(fl));
                        int k = 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (fl & (1u << q)) fixq[slot + k++] = (unsigned short)((st << 3) | (h << 2) | q);
                    }
                    row += rp + ((y + 2 <= last) ? pitch : 0);
                }
#endif
            } else {
                color_half_tile_generic(fr, g, x0, y, bc, 2 * q4, C, planes);
                color_half_tile_generic(fr, g, x0, y + 2, bc, 2 * q4 + 1, C, planes);
            }
        }
    } else {
        // REF_COMPAT (include/encoder.h:238-348): "slice" s is the 16-pixel column x = 16*s, the
        // macroblocks walk down rows y = 16*mb; chroma blocks are 8 consecutive bytes of the
        // FULL-resolution planes at linear offset (y/2+i)*(W/2) + x/2 + j.
        const int x0 = 16 * slice;
        for (int i = tid; i < 256 * nmb; i += nthr) {
            const int mb = i >> 8, r = (i >> 4) & 15, c = i & 15;
            const uint8_t *p = fr + ((size_t)(16 * (mb0 + mb) + r) * g.W + x0 + c) * g.channels;
            int yy, cb, cr;
            ycbcr_exact(p[0], p[1], p[2], yy, cb, cr);
            planes[plane_word((r >> 3) * 2 * C + 2 * mb + (c >> 3), r & 7, c & 7, C)] = yy;
        }
        const int half = g.W / 2;
        for (int i = tid; i < 64 * nmb; i += nthr) {
            const int mb = i >> 6, r = (i >> 3) & 7, c = i & 7;
            const size_t off = (size_t)(8 * (mb0 + mb) + r) * half + x0 / 2 + c;
            const uint8_t *p = fr + off * g.channels;      // pixel `off` of the row-major picture
            int yy, cb, cr;
            ycbcr_exact(p[0], p[1], p[2], yy, cb, cr);
            planes[plane_word(4 * C + mb, r, c, C)] = cb;
            planes[plane_word(5 * C + mb, r, c, C)] = cr;
        }
    }
    __syncthreads();
    M1X_MARK(1);
    if (M1_COLOUR_SPLIT != 2 && kLoad > 0) {
        // Fix-up pass of the integer colour path: the queued quads again, by the reference's double chain
        // (their pixels are L1/L2 hits).  One quad per thread, so the cost follows the NUMBER of flagged
        // quads (about 2 % of them on noise; r == g or g == b pixels are the common causes), not the number of
        // warps that happen to contain one.
        const int4 fc = *(const int4 *)(wtot + 4);          // queue lengths of the (at most four) warps
        const int nfix = fc.x + fc.y + fc.z + fc.w;
        if (nfix) {
            const int nbc = 2 * nmb;
            const unsigned inv = paired ? g.inv_nbc[2] : last_chunk ? g.inv_nbc[1] : g.inv_nbc[0];
            for (int e = tid; e < nfix; e += nthr) {
                int qi = e, qw = 0;                         // entry qi of warp qw's queue
                if (qi >= fc.x) { qi -= fc.x; qw = 1; if (qi >= fc.y) { qi -= fc.y; qw = 2; if (qi >= fc.z) { qi -= fc.z; qw = 3; } } }
                const int code = ((const unsigned short *)win)[256 * qw + qi], st = code >> 3, h = (code >> 2) & 1, q = code & 3;
                const int q4 = (int)(((unsigned)st * inv) >> 16), bc = st - q4 * nbc;
                const int sub = (paired && bc >= 2 * T) ? 1 : 0;
                color_quad_exact(fr, g, 16 * mb0 + 8 * (bc - sub * 2 * T), 16 * (slice + sub) + 4 * q4 + 2 * h, bc, 2 * q4 + h, q, C, planes);
            }
            __syncthreads();
        }
    }
    win[tid] = 0;                                           // the queue is consumed: `win` becomes the bit window
#ifdef M1_EXPERIMENTS
    if (g.debug_skip & 2) return;                           // profiling only (tools/ build)
#endif

    // ---- phase 2: one thread per 8x8 block, threads in CODING order (t = 6*mb + blk), so the
    // bit offsets are a plain scan over the thread index.  pb = the thread's plane block.
    //   2a  every block thread reads its block for min / max / sum.  A block whose samples span at most g.flat_range
    //       cannot have a non-zero AC level at this quality (m1_flat_range, m1cu_quant.h): its DC coefficient is
    //       (sum + 16) >> 3 and it is done.
    //   2b  the blocks that do need a DCT, per warp (no CTA barrier anywhere in this phase):
    //       - all or nearly all of the warp's blocks (busy content, or no flat range at this quality): one thread per
    //         block, the whole block in registers, exactly as before;
    //       - a few of them: EIGHT lanes per block (a row each, then a column each, exchanged through the block's own
    //         256 bytes), four blocks per round -- cost and latency follow the number of blocks that need a DCT, not
    //         the mere presence of one in the warp; each owner then reads its coefficients back.
    //   2c  the owner packs and tests the coefficients and writes the record.
    //   3   every block thread codes its own block from its record and non-zero mask.
    const int mb = tid / 6, blk = tid - mb * 6;             // macroblock in chunk, block 0..5
    const bool active = mb < nmb;
    const bool is_luma = blk < 4;
    const int pb = is_luma ? (blk >> 1) * 2 * C + 2 * mb + (blk & 1) : blk * C + mb;
    const int lane = tid & 31, warp = tid >> 5;
    unsigned long long nz = 0;
    BitAcc acc{0u, 0u, 0};
    {
        bool need = active;                                  // this block needs a DCT
        const unsigned amask = __ballot_sync(0xffffffffu, active);
        if (kFlat && active) {
            const int key4 = (tid & 7) << 2;                 // == blk_key(pb, C): threads are in coding order
            const int *src = planes + pb * 64;
            int mn = 255, mx = 0, sum = 0;
            // two rows per step (rows i and i + 4 of the block); straight-line code, so the eight addresses are immediates
            auto rows = [&](int i) {
                const int o = ((i << 2) ^ key4);
                const int4 a = *(const int4 *)(src + o);
                const int4 b = *(const int4 *)(src + o + 32);
                mn = __vimin3_s32(mn, a.x, a.y); mn = __vimin3_s32(mn, a.z, a.w); mn = __vimin3_s32(mn, b.x, b.y); mn = __vimin3_s32(mn, b.z, b.w);
                mx = __vimax3_s32(mx, a.x, a.y); mx = __vimax3_s32(mx, a.z, a.w); mx = __vimax3_s32(mx, b.x, b.y); mx = __vimax3_s32(mx, b.z, b.w);
                sum = m1_add3(sum, a.x, a.y); sum = m1_add3(sum, a.z, a.w); sum = m1_add3(sum, b.x, b.y); sum = m1_add3(sum, b.z, b.w);
            };
            // busy content: stop reading as soon as no block of this warp can pass any more
            rows(0); rows(1);
            if (!__all_sync(amask, mx - mn > g.flat_range)) {
                rows(2); rows(3);
                if (!__all_sync(amask, mx - mn > g.flat_range)) { rows(4); rows(5); rows(6); rows(7); }
            }
            if (mx - mn <= g.flat_range) {
                need = false;
                // source/image_processing.c:296: dct[0][0] = (x6 + 16) >> 3, x6 = sum of the eight row sums
                const int dcb = (sum + (16 + (M1_COEF_BIAS << 3))) >> 3;
                const int m0 = 0x7800 - (int)(nk.ka[0] & 0xffffu), c = dcb - M1_COEF_BIAS;
                nz = (c >= m0 || c <= -m0) ? 1ull : 0ull;
                if (kLevels) {                               // the level dump reads every position of the record
#pragma unroll
                    for (int gI = 0; gI < 8; ++gI)
                        *(uint4 *)(rec + pb * 128 + (gI << 3)) = make_uint4(0x08000800u, 0x08000800u, 0x08000800u, 0x08000800u);
                }
                rec[rec_index(pb, 0, tid & 7)] = (short)dcb;
            }
        }
        const unsigned nmask = __ballot_sync(0xffffffffu, need);
        const int n_need = __popc(nmask);
        if (nmask) {                                          // warp-uniform
            int v[64];
            if (nmask == amask || n_need > M1_LANES8_MAX) {
                // one thread per block (six rounds of the eight-lane form would cost as much)
                if (need) {
                    const int key4 = (tid & 7) << 2;
                    const int *src = planes + pb * 64;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int o = ((i << 2) ^ key4);
                        const int4 a = *(const int4 *)(src + o);
                        const int4 b = *(const int4 *)(src + o + 32);
                        v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
                        v[32 + 4 * i] = b.x; v[32 + 4 * i + 1] = b.y; v[32 + 4 * i + 2] = b.z; v[32 + 4 * i + 3] = b.w;
                    }
                    fdct8x8(v);
                }
            } else {
                // eight lanes per block.  wlist[k] = lane of the warp's k-th block that needs a DCT.  Scratch = the block's
                // own 64 words: T[k][r] (row-pass output k of row r) at word ((k ^ s) * 8 + r), s = slot 0..3 of the round
                // (its four blocks then hit four different bank groups); then the coefficients C[u][j] at
                // ((u ^ (owner lane & 7)) * 8 + j), which the owner reads back row by row.
                unsigned short *wlist = (unsigned short *)(wtot + 16) + 24 * warp;   // [<= M1_LANES8_MAX] (plane block << 8) | lane
                if (need) wlist[__popc(nmask & ((1u << lane) - 1u))] = (unsigned short)((pb << 8) | lane);
                __syncwarp();
#pragma unroll 1
                for (int base = 0; base < n_need; base += 4) {
                    const int s = lane >> 3, r = lane & 7, q = base + s;
                    const bool valid = q < n_need;
                    const int ent = valid ? wlist[q] : 0, owner = ent & 31;
                    int *blkw = planes + (ent >> 8) * 64;
                    int x[8], o[8];
                    // shared-space byte address of word r of the block; the block is 256-byte aligned, so a word index
                    // ((k ^ s) << 3) + r is the XOR of (k << 5) into (that address ^ (s << 5)): one LOP3 per store
                    const unsigned sb = (unsigned)__cvta_generic_to_shared(blkw) + 4u * r;
                    if (valid) {
                        const int key = owner & 7, i0 = 2 * r;    // (the owner's coding index & 7; warps start at multiples of 8) row r = chunks 2r, 2r + 1
                        const int4 a = *(const int4 *)(blkw + (((i0 & 8) | ((i0 & 7) ^ key)) << 2));
                        const int4 b = *(const int4 *)(blkw + ((((i0 + 1) & 8) | (((i0 + 1) & 7) ^ key)) << 2));
                        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                    }
                    __syncwarp();                             // every row is in registers before T overwrites the samples
                    if (valid) {
                        fdct_row(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], o);
                        const unsigned aT = sb ^ ((unsigned)s << 5);
#pragma unroll
                        for (int k = 0; k < 8; ++k) asm volatile("st.shared.b32 [%0], %1;" :: "r"(aT ^ (unsigned)(k << 5)), "r"(o[k]) : "memory");
                    }
                    __syncwarp();
                    if (valid) {
                        const int4 a = *(const int4 *)(blkw + ((r ^ s) << 3)), b = *(const int4 *)(blkw + ((r ^ s) << 3) + 4);
                        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                    }
                    __syncwarp();                             // every column is in registers before C overwrites T
                    if (valid) {
                        fdct_col(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7], o);
                        const unsigned aC = sb ^ ((unsigned)(owner & 7) << 5);
#pragma unroll
                        for (int u = 0; u < 8; ++u) asm volatile("st.shared.b32 [%0], %1;" :: "r"(aC ^ (unsigned)(u << 5)), "r"(o[u]) : "memory");
                    }
                }
                __syncwarp();
                if (need) {
                    // row u of the coefficients sits at word (u ^ (lane & 7)) << 3 of the block: XOR on the byte address again
                    const unsigned rb = ((unsigned)__cvta_generic_to_shared(planes + pb * 64)) ^ ((unsigned)(lane & 7) << 5);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        int4 a, b;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(rb ^ (unsigned)(u << 5)) : "memory");
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4+16];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(rb ^ (unsigned)(u << 5)) : "memory");
                        v[8 * u] = a.x; v[8 * u + 1] = a.y; v[8 * u + 2] = a.z; v[8 * u + 3] = a.w;
                        v[8 * u + 4] = b.x; v[8 * u + 5] = b.y; v[8 * u + 6] = b.z; v[8 * u + 7] = b.w;
                    }
                }
            }
            if (need) {
                // packed coefficient pairs + the zigzag-order non-zero mask (m1cu_block.cuh)
                uint32_t pk[32];
                nz = pack_and_flag(v, pk, nk);
                // the block's coefficients are in registers now: the first half of its 256 bytes of plane becomes the record
#pragma unroll
                for (int gI = 0; gI < 8; ++gI)
                    *(uint4 *)(rec + pb * 128 + (((gI ^ tid) & 7) << 3)) =
                        make_uint4(pk[4 * gI], pk[4 * gI + 1], pk[4 * gI + 2], pk[4 * gI + 3]);
            }
        }
    }
    if (active) {
        // ---- phase 3: code the block into registers ------------------------------------------
        if (blk == 0) { acc.lo = 3u; acc.n = 2; }           // address increment '1' + macroblock_type '1'
        if (code_block(acc, rec, pb, nz, is_luma, tb, tid & 7)) atomicOr(err, M1_ERRBIT_LEVEL);
        acc.finish();
    }

    // scan of the block lengths in thread (= coding) order
    const int my_bits = acc.n;
    int incl = my_bits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)                        // the shuffle's own predicate says "lane >= d"
        asm volatile("{\n\t.reg .pred p;\n\t.reg .s32 t;\n\tshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t@p add.s32 %0, %0, t;\n\t}"
                     : "+r"(incl) : "r"(d));
    if (lane == 31) wtot[warp] = incl;
    const int t2 = 6 * T;                                    // first thread of the second record of a pair
    if (paired && tid == t2) wtot[8] = incl - my_bits;       // its exclusive prefix inside its warp
    __syncthreads();
    M1X_MARK(2);

    if (kLevels) {
        // debug output: quantised zigzag levels in coding order, [picture][macroblock][6][64]
        short *dst = levels + ((size_t)frame * g.mbs_per_frame + (size_t)slice * g.mbs_per_slice + mb0) * 384;
        for (int i = tid; i < nmb * 384; i += nthr) {
            const int p = i >> 6, z = i & 63, m = p / 6, b = p - m * 6;
            const int t = b < 4 ? (b >> 1) * 2 * C + 2 * m + (b & 1) : b * C + m;
            // slot m of a pair's second half is macroblock m - T of the NEXT slice
            const size_t o = (paired && m >= T) ? (size_t)(g.mbs_per_slice - T) * 384 : 0;
            dst[o + i] = (short)quant_level(rec[rec_index(t, z, p & 7)], z, tb);   // p = 6 * m + b = the block's thread
        }
    }

    const int hdr_bits = chunk == 0 ? M1_SLICE_HDR_BITS : 0;
    const int4 wt = *(const int4 *)wtot;                    // blockDim.x <= 128: at most four warps
    const int real_bits = hdr_bits + wt.x + wt.y + wt.z + wt.w;
    const int base = hdr_bits + (warp > 0 ? wt.x : 0) + (warp > 1 ? wt.y : 0) + (warp > 2 ? wt.z : 0);
    // A pair's second record starts on a fresh 32-bit word of the (virtual) bit string the window passes walk:
    // bits0 = bits of the first record, gap = the padding in front of the second.
    int bits0 = real_bits, gap = 0;
    if (paired) {
        const int w2 = t2 >> 5;
        bits0 = (w2 > 0 ? wt.x : 0) + (w2 > 1 ? wt.y : 0) + (w2 > 2 ? wt.z : 0) + wtot[8];   // (a last chunk has no slice header)
        gap = (32 - (bits0 & 31)) & 31;
    }
    const int total_bits = real_bits + gap;
    const int my_off = base + incl - my_bits + ((paired && tid >= t2) ? gap : 0);

    // record index and word offsets in 32 bits: the staging of one launch round is bounded to 2 GiB (m1cu_create)
    const unsigned rec0 = (unsigned)frame * (unsigned)g.chunks_per_frame + (unsigned)(slice * g.chunks_per_slice + chunk);
    const unsigned rec_words = g.chunk_stride >> 2;
    uint32_t *out = staging + (size_t)(rec0 * rec_words);
    // word v of the virtual bit string -> its place in the staging records
    const int vsplit = paired ? (bits0 + gap) >> 5 : 0x7fffffff;
    uint32_t *out2 = staging + (size_t)((rec0 + (unsigned)g.chunks_per_slice) * rec_words - (unsigned)(paired ? vsplit : 0));
    const int WW = g.win_words;                             // <= M1_WIN_WORDS (smaller only in tests)
    if (min(WW, (total_bits + 31) >> 5) + 2 > nthr) {       // uniform; rare at typical qualities
#pragma unroll 1
        for (int i = nthr + tid; i < WW + 2; i += nthr) win[i] = 0;
        __syncthreads();
    }
    for (int w0 = 0;; w0 += 32 * WW) {            // the window words in use are zero here
        if (tid == 0 && hdr_bits && w0 == 0) {
            // source/mpeg1_blk.c:12-20: 000001 | (vertical_pos+1)&0xff | quant_scale(5)=1 | 0 = 38 bits at bit 0 of
            // the chunk: word 0 = 0x000001 vv, word 1 starts 00001 0
            atomicOr(&win[0], 0x100u | (((uint32_t)(slice & 0xff) + 1u) & 0xffu));
            atomicOr(&win[1], 0x08000000u);
        }
        if (active && my_off < w0 + 32 * WW && my_off + my_bits > w0) {
            if (my_bits <= 64 && my_off >= w0 && my_off + my_bits <= w0 + 32 * WW) {
                const int p = my_off - w0, word = p >> 5, o = p & 31;
                const uint32_t a = acc.hi >> o;
                const uint32_t b = __funnelshift_r(acc.lo, acc.hi, o);
                const uint32_t c = __funnelshift_r(0u, acc.lo, o);
                if (a) atomicOr(&win[word], a);
                if (b) atomicOr(&win[word + 1], b);
                if (c) atomicOr(&win[word + 2], c);
            } else {
                recode_into_window(win, my_off, w0, WW, blk == 0, rec, pb, nz, is_luma, tb, tid & 7);
            }
        }
        __syncthreads();
        const int nwords = min(WW, (total_bits - w0 + 31) >> 5);
#pragma unroll 1
        for (int i = tid; i < nwords; i += nthr) {
            const int v = (w0 >> 5) + i;
            (v < vsplit ? out : out2)[v] = win[i];
        }
        if (w0 + 32 * WW >= total_bits) break;
        __syncthreads();                                    // rare: the chunk needs another window pass
#pragma unroll 1
        for (int i = tid; i < WW + 2; i += nthr) win[i] = 0;
        __syncthreads();
    }
    if (tid == 0) {
        chunk_bits[rec0] = (uint32_t)bits0;
        if (paired) chunk_bits[rec0 + g.chunks_per_slice] = (uint32_t)(real_bits - bits0);
    }
    M1X_MARK(3);
}

// -------------------------------------------------------------------------------------------
// k_layout: one CTA per picture.  Slice s starts at a byte boundary (include/encoder.h:442-443);
// chunk c of slice s starts at slice start + bits of the slice's earlier chunks.  The last CTA
// to finish turns the picture sizes into 16-byte-aligned offsets, continuing from *running.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_layout(const __grid_constant__ M1Geom g, int n_frames, const uint32_t *__restrict__ chunk_bits,
         uint32_t *__restrict__ chunk_dst, uint32_t *__restrict__ frame_bytes,
         unsigned long long *__restrict__ frame_off, unsigned long long *running,
         unsigned int *done_counter, unsigned long long out_cap, int *err)
{
    __shared__ unsigned int wsum[8];
    __shared__ unsigned int carry;
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    const uint32_t *cb = chunk_bits + (size_t)f * g.chunks_per_frame;
    uint32_t *cd = chunk_dst + (size_t)f * g.chunks_per_frame;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int s0 = 0; s0 < g.slices; s0 += 256) {
        const int s = s0 + tid;
        unsigned int bits = 0;
        if (s < g.slices) {
            for (int c = 0; c < g.chunks_per_slice; ++c) bits += cb[s * g.chunks_per_slice + c];
            bits = (bits + 7u) & ~7u;
        }
        unsigned int incl = bits;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned int base = carry;
        for (int w = 0; w < warp; ++w) base += wsum[w];
        if (s < g.slices) {
            unsigned int pos = base + incl - bits;
            for (int c = 0; c < g.chunks_per_slice; ++c) {
                cd[s * g.chunks_per_slice + c] = pos;
                pos += cb[s * g.chunks_per_slice + c];
            }
        }
        __syncthreads();
        if (tid == 255) carry = base + incl;
        __syncthreads();
    }
    if (tid == 0) {
        frame_bytes[f] = carry >> 3;
        __threadfence();
        const unsigned int prev = atomicAdd(done_counter, 1u);
        is_last = (prev == (unsigned int)n_frames - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // serial-by-tile exclusive scan of the aligned picture sizes (n_frames is small)
    if (warp == 0) {
        unsigned long long base = *running;
        for (int f0 = 0; f0 < n_frames; f0 += 32) {
            const int ff = f0 + lane;
            const unsigned long long sz = ff < n_frames ? (((unsigned long long)((volatile uint32_t *)frame_bytes)[ff] + 15ull) & ~15ull) : 0ull;
            unsigned long long incl = sz;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (ff < n_frames) frame_off[ff] = base + incl - sz;
            base += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) {
            frame_off[n_frames] = base;
            *running = base;
            *done_counter = 0;
            if (base > out_cap) atomicOr(err, M1_ERRBIT_CAPACITY);
        }
    }
}

// -------------------------------------------------------------------------------------------
// k_stitch: every thread produces 32-bit words of the final payload.  blockIdx.y = picture.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t read_bits(const uint32_t *__restrict__ src, unsigned int bit, int n)
{
    // n in 1..32 bits starting at `bit` of a logical MSB-first word array
    const unsigned int w = bit >> 5, o = bit & 31;
    unsigned long long v = (unsigned long long)src[w] << 32;
    if (o + n > 32) v |= src[w + 1];
    return (uint32_t)((v << o) >> (64 - n));
}

__global__ void __launch_bounds__(256)
k_stitch(const __grid_constant__ M1Geom g, const uint32_t *__restrict__ staging,
         const uint32_t *__restrict__ chunk_bits, const uint32_t *__restrict__ chunk_dst,
         const uint32_t *__restrict__ frame_bytes, const unsigned long long *__restrict__ frame_off,
         uint8_t *__restrict__ out, unsigned long long out_cap)
{
    const int f = blockIdx.y;
    const unsigned int fbytes = frame_bytes[f];
    const unsigned long long foff = frame_off[f];
    if (foff + ((fbytes + 15ull) & ~15ull) > out_cap) return;      // flagged by k_layout
    const unsigned int nwords = (fbytes + 3u) >> 2;
    const uint32_t *cb = chunk_bits + (size_t)f * g.chunks_per_frame;
    const uint32_t *cd = chunk_dst + (size_t)f * g.chunks_per_frame;
    const int nc = g.chunks_per_frame;
    uint32_t *dst = (uint32_t *)(out + foff);
    // One warp per chunk, no search: chunk c OWNS the output words whose first bit lies in
    // [cd[c], cd[c+1]) (its own bits plus any slice padding behind it); a word that runs past the
    // chunk's end pulls the rest from the following chunk(s).
    const int lane = threadIdx.x & 31;
    const int warps_per_row = (gridDim.x * blockDim.x) >> 5;
    for (int c0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c0 < nc; c0 += warps_per_row) {
        const unsigned int own_lo = cd[c0];
        const unsigned int own_hi = c0 + 1 < nc ? cd[c0 + 1] : nwords * 32u;
        const unsigned int w_lo = (own_lo + 31u) >> 5, w_hi = min((own_hi + 31u) >> 5, nwords);
        for (unsigned int w = w_lo + lane; w < w_hi; w += 32) {
            unsigned int b = w * 32u;
            const unsigned int end = b + 32u;
            int c = c0;
            uint32_t acc = 0;
            while (b < end && c < nc) {
                const unsigned int cs = cd[c], ce = cs + cb[c];
                if (b < ce) {
                    const int take = (int)(min(end, ce) - b);
                    const uint32_t *src = staging + (size_t)((size_t)f * nc + c) * (g.chunk_stride / 4);
                    const uint32_t bits = read_bits(src, b - cs, take);
                    acc |= bits << (end - b - take);
                    b += take;
                }
                if (b >= ce) {
                    ++c;
                    if (c < nc) { const unsigned int ns = cd[c]; if (ns > b) b = min(end, ns); }  // slice padding zeros
                    else b = end;
                }
            }
            dst[w] = __byte_perm(acc, 0, 0x0123);
        }
    }
}

// -------------------------------------------------------------------------------------------
// Full-resolution planes (the .bit side files, source/image_processing.c:753-787) and the
// synthetic generator (the test oracle restates the same integer formula).
// -------------------------------------------------------------------------------------------
// blockIdx.y = picture: picture f's planes are Y, Cb, Cr + f * plane_stride, its pixels rgb + f * frame_stride.
__global__ void k_ycbcr_planes(const uint8_t *__restrict__ rgb, int channels, size_t npix, size_t frame_stride,
                               size_t plane_stride, uint8_t *__restrict__ Y, uint8_t *__restrict__ Cb, uint8_t *__restrict__ Cr)
{
    const size_t f = blockIdx.y;
    rgb += f * frame_stride; Y += f * plane_stride; Cb += f * plane_stride; Cr += f * plane_stride;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t *p = rgb + i * channels;
        int y, cb, cr;
        ycbcr_exact(p[0], p[1], p[2], y, cb, cr);
        Y[i] = (uint8_t)y; Cb[i] = (uint8_t)cb; Cr[i] = (uint8_t)cr;
    }
}

// Copies the payload bytes [0, *end) of one encode call (end = frame_offsets[n], a multiple of 16) to
// `dst`, which may be another GPU's memory mapped over NVLink (distributed.PeerGather, staged mode).
// A few CTAs on a high-priority side stream: they slip in as the next step's encode CTAs retire.
__global__ void __launch_bounds__(256)
k_push_bytes(uint4 *__restrict__ dst, const uint4 *__restrict__ src, const unsigned long long *__restrict__ end,
             unsigned long long cap)
{
    const unsigned long long n16 = min(*end, cap) >> 4;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n16;
         i += (unsigned long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// -------------------------------------------------------------------------------------------
// Final stream image on the device: [prologue] + per picture { 44-byte prefix with the
// packet length patched in (include/encoder.h:448-454), payload, 4-byte trailer }.  The prefix bytes
// themselves are built on the host by the reference-API header writers (include/mpeg1_enc.h) and passed
// in as 256 templates indexed by (picture index & 255) -- the reference's clock is a uint8_t hour.
//   k_stream_offsets  one CTA: exclusive scan of (48 + payload bytes) -> start of every picture's segment
//   k_stream_copy     blockIdx.y = picture; aligned 32-bit words of the destination are assembled from two
//                     source words, the ragged first / last bytes of a segment are written bytewise
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_stream_offsets(int n_frames, const uint32_t *__restrict__ frame_bytes, unsigned long long base,
                 unsigned long long *__restrict__ seg_off, unsigned long long cap, int *err)
{
    __shared__ unsigned long long wsum[32];
    __shared__ unsigned long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = base;
    __syncthreads();
    for (int f0 = 0; f0 < n_frames; f0 += 1024) {
        const int f = f0 + tid;
        const unsigned long long sz = f < n_frames ? 48ull + frame_bytes[f] : 0ull;
        unsigned long long incl = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned long long before = carry;
        for (int w = 0; w < warp; ++w) before += wsum[w];
        if (f < n_frames) seg_off[f] = before + incl - sz;
        __syncthreads();
        if (tid == 1023) carry = before + incl;
        __syncthreads();
    }
    if (tid == 0) {
        seg_off[n_frames] = carry;
        if (carry > cap) atomicOr(err, M1_ERRBIT_CAPACITY);
    }
}

// byte k of picture f's segment
__device__ __forceinline__ uint32_t stream_byte(unsigned int k, unsigned int n, const uint8_t *__restrict__ prefix,
                                                const uint8_t *__restrict__ payload, uint32_t trailer_be)
{
    if (k < 44u) {
        const unsigned int len = (44u + n - 8u) & 0xffffu;          // unsigned short arithmetic of the reference
        return k == 4u ? len >> 8 : k == 5u ? len & 0xffu : prefix[k];
    }
    if (k < 44u + n) return payload[k - 44u];
    return (trailer_be >> (8u * (3u - (k - 44u - n)))) & 0xffu;
}

__global__ void __launch_bounds__(256)
k_stream_copy(const uint8_t *__restrict__ payloads, const uint32_t *__restrict__ frame_bytes,
              const unsigned long long *__restrict__ frame_off, const unsigned long long *__restrict__ seg_off,
              long first_index, const uint8_t *__restrict__ prefix256, uint32_t trailer_be,
              uint8_t *__restrict__ out, unsigned long long cap)
{
    const int f = blockIdx.y;
    const unsigned long long s0 = seg_off[f], s1 = seg_off[f + 1];
    if (s1 > cap) return;                                            // flagged by k_stream_offsets
    const unsigned int n = frame_bytes[f];
    const uint8_t *payload = payloads + frame_off[f];                // 16-byte aligned
    const uint8_t *prefix = prefix256 + 44u * (unsigned int)((first_index + f) & 255);
    // destination words fully inside the payload part: [a0, a1) in bytes, both multiples of 4
    const unsigned long long p0 = s0 + 44ull, p1 = p0 + n;
    const unsigned long long a0 = (p0 + 3ull) & ~3ull, a1 = p1 & ~3ull;
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (a1 > a0) {
        const unsigned int sh = (unsigned int)(a0 - p0);             // payload byte that lands on word a0: 0..3
        const uint32_t *src = (const uint32_t *)payload;
        uint32_t *dst = (uint32_t *)(out + a0);
        const unsigned int nw = (unsigned int)((a1 - a0) >> 2);
        for (unsigned int i = tid; i < nw; i += nthr) {
            const uint32_t lo = src[i], hi = sh ? src[i + 1] : 0u;  // payload bytes 4i+sh .. 4i+sh+3; i+1 stays inside the
            dst[i] = __funnelshift_r(lo, hi, 8u * sh);               // 16-byte padded picture because 4i+sh+3 < n
        }
    }
    // the ragged ends: prefix + bytes up to a0, and bytes from a1 to the end of the trailer
    const unsigned long long e0 = a1 > a0 ? a0 : s1, b1 = a1 > a0 ? a1 : s1;
    for (unsigned long long g = s0 + tid; g < e0; g += nthr) out[g] = (uint8_t)stream_byte((unsigned int)(g - s0), n, prefix, payload, trailer_be);
    for (unsigned long long g = b1 + tid; g < s1; g += nthr) out[g] = (uint8_t)stream_byte((unsigned int)(g - s0), n, prefix, payload, trailer_be);
}

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void k_synth_rgb(uint32_t seed, long first_frame, int n_frames, int W, int H, int kind,
                            uint8_t *__restrict__ rgb)
{
    const size_t per = (size_t)W * H;
    const size_t total = per * n_frames;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t fi = i / per, p = i - fi * per;
        const uint32_t f = (uint32_t)(first_frame + (long)fi);
        const uint32_t y = (uint32_t)(p / W), x = (uint32_t)(p - (size_t)y * W);
        const uint32_t fkey = mix32(seed * 0x85ebca6bu + f * 0x9e3779b9u + 0x165667b1u);
        const uint32_t n = mix32(fkey ^ (y * (uint32_t)W + x));
        uint8_t *o = rgb + i * 3;
        bool noisy = kind == 1;
        if (kind == 4)                                              // scattered: a quarter of the 8x8 pixel tiles are noise
            noisy = (mix32(fkey ^ (0x51ed270bu + ((y >> 3) * 0x9e3779b1u) + (x >> 3))) & 3u) == 0u;
        if (noisy) {
            o[0] = (uint8_t)n; o[1] = (uint8_t)(n >> 8); o[2] = (uint8_t)(n >> 16);
        } else {
            o[0] = (uint8_t)((255u * x / (uint32_t)W + (n & 15u) + f) & 255u);
            o[1] = (uint8_t)((255u * y / (uint32_t)H + ((n >> 4) & 15u)) & 255u);
            o[2] = (uint8_t)(((x + y) / 8u + ((n >> 8) & 15u) + 2u * f) & 255u);
            if (kind == 2) { o[1] = o[0]; o[2] = o[0]; }          // grey
            else if (kind == 3) { o[1] = o[0]; }                   // r == g
        }
    }
}

// -------------------------------------------------------------------------------------------
// Launchers (called from m1cu_api.cu).
// -------------------------------------------------------------------------------------------
size_t m1k_encode_smem_bytes(const M1Geom &g, int threads)
{
    (void)threads;
#ifdef M1_EXPERIMENTS
    static const size_t pad = (size_t)m1x_env_int("M1_PAD_SMEM");   // occupancy experiments
#else
    const size_t pad = 0;
#endif
    return (size_t)6 * g.chunk_mbs * 256 + (size_t)(M1_WIN_WORDS + 4) * 4 + sizeof(M1Tables) + 64 * sizeof(int) + 16 + pad;
}

int m1k_encode_threads(const M1Geom &g) { return (8 * g.chunk_mbs + 31) & ~31; }   // one colour tile per thread; 6C of them own a block

typedef void (*encode_kernel_t)(const M1Geom, const M1NzKeys, const uint8_t *, const M1Tables *, uint32_t *, uint32_t *,
                                short *, int *);

template <int kLoad>
static encode_kernel_t pick_encode_kernel_for(bool levels, bool flat)
{
    if (flat) return levels ? k_encode_chunks<kLoad, true, true> : k_encode_chunks<kLoad, false, true>;
    return levels ? k_encode_chunks<kLoad, true, false> : k_encode_chunks<kLoad, false, false>;
}
static encode_kernel_t pick_encode_kernel(const M1Geom &g, bool levels)
{
    const bool flat = g.flat_range >= 0;
    if (g.mode != 0) return pick_encode_kernel_for<-1>(levels, flat);
    if (g.fast_load == 3) return pick_encode_kernel_for<3>(levels, flat);
    if (g.fast_load == 4) return pick_encode_kernel_for<4>(levels, flat);
    return pick_encode_kernel_for<0>(levels, flat);
}

cudaError_t m1k_prepare(const M1Geom &g)
{
    const size_t smem = m1k_encode_smem_bytes(g, m1k_encode_threads(g));
    M1Geom v = g;
    // every variant this geometry can be launched with (an unaligned input pointer falls back to 0)
    for (int fl : { g.fast_load, 0 }) {
        v.fast_load = fl;
        for (bool lv : { false, true }) {
            cudaError_t e = cudaFuncSetAttribute(pick_encode_kernel(v, lv), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaSuccess;
}

cudaError_t m1k_launch_encode(const M1Geom &g, const M1Quant &q, const uint8_t *rgb, int n_frames,
                              const M1Tables *tables, uint32_t *staging, uint32_t *chunk_bits,
                              short *levels, int *err, cudaStream_t st)
{
    const int threads = m1k_encode_threads(g);
    const size_t smem = m1k_encode_smem_bytes(g, threads);
    dim3 grid(g.chunks_per_slice, g.slices, n_frames);
    M1NzKeys nk;
    m1k_nz_keys(q, &nk);
#ifdef M1_EXPERIMENTS
    m1x_before_launch(grid, st);
#endif
    pick_encode_kernel(g, levels != nullptr)<<<grid, threads, smem, st>>>(g, nk, rgb, tables, staging, chunk_bits, levels, err);
#ifdef M1_EXPERIMENTS
    m1x_after_launch(grid, st);
#endif
    return cudaGetLastError();
}

cudaError_t m1k_launch_layout(const M1Geom &g, int n_frames, const uint32_t *chunk_bits, uint32_t *chunk_dst,
                              uint32_t *frame_bytes, unsigned long long *frame_off,
                              unsigned long long *running, unsigned int *done_counter,
                              unsigned long long out_cap, int *err, cudaStream_t st)
{
    k_layout<<<n_frames, 256, 0, st>>>(g, n_frames, chunk_bits, chunk_dst, frame_bytes, frame_off, running,
                                       done_counter, out_cap, err);
    return cudaGetLastError();
}

cudaError_t m1k_launch_stitch(const M1Geom &g, int n_frames, int blocks_x, const uint32_t *staging,
                              const uint32_t *chunk_bits, const uint32_t *chunk_dst,
                              const uint32_t *frame_bytes, const unsigned long long *frame_off,
                              uint8_t *out, unsigned long long out_cap, cudaStream_t st)
{
    dim3 grid(blocks_x, n_frames);
    k_stitch<<<grid, 256, 0, st>>>(g, staging, chunk_bits, chunk_dst, frame_bytes, frame_off, out, out_cap);
    return cudaGetLastError();
}

cudaError_t m1k_launch_planes(const uint8_t *rgb, int channels, size_t npix, int n_frames, size_t frame_stride,
                              size_t plane_stride, uint8_t *Y, uint8_t *Cb, uint8_t *Cr, cudaStream_t st)
{
    const int blocks = (int)((npix + 255) / 256 < 148 * 8 ? (npix + 255) / 256 : 148 * 8);
    dim3 grid(blocks > 0 ? blocks : 1, n_frames);
    k_ycbcr_planes<<<grid, 256, 0, st>>>(rgb, channels, npix, frame_stride, plane_stride, Y, Cb, Cr);
    return cudaGetLastError();
}

cudaError_t m1k_launch_push(uint8_t *dst, const uint8_t *src, const unsigned long long *end, unsigned long long cap,
                            cudaStream_t st)
{
    k_push_bytes<<<64, 256, 0, st>>>((uint4 *)dst, (const uint4 *)src, end, cap);
    return cudaGetLastError();
}

cudaError_t m1k_launch_stream(const uint8_t *payloads, const uint32_t *frame_bytes, const unsigned long long *frame_off,
                              int n_frames, long first_index, const uint8_t *prefix256, uint32_t trailer_be,
                              unsigned long long base, unsigned long long *seg_off, uint8_t *out, unsigned long long cap,
                              int *err, cudaStream_t st)
{
    k_stream_offsets<<<1, 1024, 0, st>>>(n_frames, frame_bytes, base, seg_off, cap, err);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 grid(8, n_frames);
    k_stream_copy<<<grid, 256, 0, st>>>(payloads, frame_bytes, frame_off, seg_off, first_index, prefix256, trailer_be, out, cap);
    return cudaGetLastError();
}

cudaError_t m1k_launch_synth(uint32_t seed, long first_frame, int n_frames, int W, int H, int kind,
                             uint8_t *rgb, cudaStream_t st)
{
    k_synth_rgb<<<148 * 8, 256, 0, st>>>(seed, first_frame, n_frames, W, H, kind, rgb);
    return cudaGetLastError();
}
