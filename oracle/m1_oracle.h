/*
 * m1_oracle.h -- CPU restatement of the reference's per-block MPEG-1 I-frame path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The shipped path is the CUDA library (ec504_imageencoder_b200/csrc, include/m1cu.h) and it
 * never calls into this file.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_vs_ref.py, run in
 * the dev container where /root/reference exists) against the reference's own code compiled
 * unmodified into oracle/_ref/libm1ref.so, and (tests/test_oracle_golden.py, runs anywhere)
 * against golden vectors generated from that library by tests/golden/make_golden.py.
 *
 * All file:line citations are relative to the reference checkout (/root/reference).
 */
#ifndef M1_ORACLE_H
#define M1_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { M1O_MODE_FULL = 0, M1O_MODE_REF_COMPAT = 1 };
enum { M1O_SYNTH_NATURAL = 0, M1O_SYNTH_NOISE = 1, M1O_SYNTH_GREY = 2, M1O_SYNTH_RG_EQUAL = 3, M1O_SYNTH_SCATTERED = 4 };

/* source/image_processing.c:314-343 (scale_quantization_matrix); out is row-major [i*8+j]. */
void m1o_qmatrix(int quality_factor, int32_t out[64]);

/* source/image_processing.c:68-110 (convert_rgb_to_ycbcr); planes are caller-allocated. */
void m1o_rgb_to_ycbcr(const uint8_t *rgb, int channels, long npix,
                      uint8_t *Y, uint8_t *Cb, uint8_t *Cr);

/* source/image_processing.c:114-133 (subsampling_420) for one plane; out is (W/2)*(H/2). */
void m1o_subsample_420(const uint8_t *plane, int W, int H, uint8_t *out);

/* source/image_processing.c:192-307 (fast_DCT): integer butterflies, out[u*8+v]. */
void m1o_fdct8x8(const uint8_t blk[64], int32_t out[64]);

/* source/image_processing.c:349-381 (quantization + zigzag_scanning). */
void m1o_quant_zigzag(const int32_t dct[64], const int32_t qm[64], int32_t zz[64]);

/* source/mpeg1_blk.c:67-117 + source/image_processing.c:400-433 + source/vlc.c:315-385.
 * Writes one char ('0'/'1') per bit into bits (capacity cap, NUL-terminated) and returns the
 * bit count, or -2 when a coded AC level has |L| >= 256 (the reference dereferences NULL there). */
int m1o_block_bits(const int32_t zz[64], int is_luma, char *bits, int cap);

/* One picture's slice payload (include/encoder.h:216-445 generalised, see oracle/README.md).
 * rgb: interleaved, `channels` bytes per pixel (>= 3), W x H.  Returns payload bytes, -1 if out
 * is too small, -2 on an unencodable level, -3 on bad arguments.  levels (optional) receives the
 * zigzag-ordered quantised levels, int16, [macroblock][6][64] in coding order. */
long m1o_encode_picture(const uint8_t *rgb, int W, int H, int channels, int mode,
                        const int32_t qm[64], uint8_t *out, long cap, int16_t *levels);

/* Number of macroblocks m1o_encode_picture codes for this geometry (levels sizing). */
long m1o_picture_macroblocks(int W, int H, int mode);

/* source/mpeg1_enc.c:7-45 as driven by include/encoder.h:86-89: the 27-byte file prologue. */
int m1o_file_prologue(uint8_t out[27]);

/* source/mpeg1_enc.c:47-137 as driven by include/encoder.h:196-231,448-454: the 44 bytes in
 * front of picture `frame_index`'s payload (packet, sequence, GOP, picture headers). */
int m1o_frame_prefix(long frame_index, int W, int H, int mode, long payload_bytes, uint8_t out[44]);

/* Whole stream image: prologue + per frame (prefix + payload + 4-byte trailer 00 00 01 b7).
 * frames = n_frames consecutive W*H*channels images.  Returns total bytes or <0. */
long m1o_encode_stream(const uint8_t *frames, long n_frames, int W, int H, int channels, int mode,
                       int quality_factor, uint8_t *out, long cap);

/* Seeded synthetic RGB (ours).  Same integer formula as the CUDA
 * generator m1cu_synth_rgb so both sides see identical bytes. */
void m1o_synth_rgb(uint32_t seed, long frame_index, int W, int H, int kind, uint8_t *rgb);

#ifdef __cplusplus
}
#endif
#endif
