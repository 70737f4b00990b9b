// m1cu_common.cuh -- shared declarations for the sm_100a MPEG-1 I-frame path.
//
// Pipeline (one launch sequence per batch of pictures):
//   k_encode_chunks   RGB -> exact YCbCr -> 4:2:0 -> 8x8 int32 DCT -> quantise -> zigzag ->
//                     DC/AC VLC -> bits of one "chunk" (<= chunk_mbs macroblocks of one slice)
//                     packed MSB-first in shared memory -> chunk staging + chunk bit count
//   k_layout          per picture: slice/chunk bit offsets (slices byte-aligned), picture sizes,
//                     then (last CTA) 16-byte-aligned picture offsets
//   k_stitch          funnel-shifts the chunk bit strings into the final payload bytes
//
// Reference behaviour reproduced (paths relative to the reference checkout): see include/m1cu.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define M1_MAX_CHUNK_MBS 16          // upper bound of macroblocks per chunk (CTA of 128 threads)
#define M1_DEFAULT_CHUNK_MBS 16      // default target: 8*16 = 128 colour-tile threads, 6*16 = 96 block threads per CTA
#define M1_BLOCK_MAX_BITS 901
#define M1_MB_MAX_BITS (2 + 6 * M1_BLOCK_MAX_BITS)
#define M1_SLICE_HDR_BITS 38
#define M1_WIN_WORDS 512             // shared-memory bit window per chunk pass (2 KiB = 16384 bits)

enum { M1_ERRBIT_CAPACITY = 1, M1_ERRBIT_LEVEL = 2 };

struct M1Geom {
    int W, H, channels;
    int mode;                  // M1CU_MODE_*
    int fast_load;             // 3 / 4: channels with 16-byte-aligned 16-pixel tiles, 0: generic loads only
    int win_words;             // bit-window words actually used (= M1_WIN_WORDS; tests shrink it via M1_WIN_WORDS)
    int debug_skip;            // profiling only (env M1_DEBUG_SKIP): bit 0 skips the colour phase, bit 1 the block phases
    int slices;                // per picture
    int mbs_per_slice;
    int chunk_mbs;             // macroblocks per chunk (last chunk of a slice may hold fewer)
    int chunks_per_slice;
    int chunks_per_frame;
    int mbs_per_frame;
    int pair_tails;            // > 0: the last chunk of a slice holds at most chunk_mbs / 2 macroblocks (= this value), and
                               // ONE CTA encodes the last chunks of slices 2k and 2k+1 together (FULL mode only)
    unsigned inv_nbc[3];       // ceil(2^16 / block columns) of a full chunk [0], of a slice's last chunk [1] and of a
                               // pair of last chunks [2]: x / nbc == (x * inv) >> 16 for x < 2048
    int flat_range;            // >= 0: a block whose samples span at most this much has no non-zero AC level at this quality
                               // (m1_flat_range, m1cu_quant.h) and skips DCT + quantiser test; -1: no such range
    unsigned chunk_stride;     // staging bytes per chunk (multiple of 16)
    unsigned long long frame_stride;   // input bytes per picture
};

// Quantiser constants for raster position k (m1cu_quant.h fills and checks them):
//   |level| = floor(|c| / m) = high word of (2|c| + 1) * rcp, rcp = ceil(2^31 / m)   (C truncating division
//   c / m of source/image_processing.c:367 once the sign is put back), exact for |c| <= 2047;
//   level != 0  <=>  (unsigned)(c + ta) > tb.
struct M1Quant {
    uint32_t rcp[64];
    int ta[64];                // m - 1
    int tb[64];                // 2m - 2
};

// Non-zero test constants, one word per coefficient pair (see pack_and_flag): lanes 0x7800 - m and
// 0x8800 - m.  Passed by value as a kernel parameter so that the statically indexed uses read them
// straight from the constant bank.
struct M1NzKeys {
    uint32_t ka[32], kb[32];
};

struct alignas(16) M1Tables {   // size is a multiple of 16 (copied to shared memory in 128-bit pieces)
    uint32_t ac[112];          // (len << 24) | code; entry 0 = '11' (run 0, |level| 1), see code_block
    uint32_t dc[18];
    uint16_t acrun[64];        // run r (= zeros - 1): first entry | entries << 8; 0 entries from r = 32 up
    uint32_t qrcp[64];         // quantiser reciprocals in ZIGZAG order (for the coder's dynamic index)
    uint8_t  zofs[64];         // byte offset of zigzag position z in an unswizzled coefficient record
    uint32_t ka[32], kb[32];   // per coefficient-pair word: lanes 0x7800 - m and 0x8800 - m (non-zero test)
};
