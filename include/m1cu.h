/*
 * m1cu.h -- thin C ABI between the host C encoder and the sm_100a CUDA path.
 *
 * This is the drop-in boundary for the reference's per-picture loop body,
 * include/encoder.h:216-445 (convert_rgb_to_ycbcr -> subsampling_420 -> per macroblock
 * extract_8x8_block / fast_DCT / quantization / zigzag_scanning / run_length_encode /
 * encode_block_header_i / encode_block_end -> slice padding).  Everything above that loop
 * (file I/O, JPEG decode, pack/system/packet/sequence/GOP/picture headers, packet-length patch)
 * stays host C and calls in here through these entry points only.
 *
 * Conventions: extern "C", plain pointers and sizes, int return codes (M1CU_OK == 0, negative on
 * failure, same spirit as the reference's 0 / -1), no exceptions, one context per GPU.  Work is
 * stream-ordered on the context's stream.  There is NO CPU fallback: without a CUDA device every
 * entry point that computes fails with M1CU_ERR_CUDA.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference checkout):
 *   m1cu_qmatrix            scale_quantization_matrix   source/image_processing.c:314-343
 *   m1cu_encode_device/host the loop body               include/encoder.h:216-445, i.e.
 *                           convert_rgb_to_ycbcr        source/image_processing.c:68-110
 *                           subsampling_420             source/image_processing.c:114-133
 *                           extract_8x8_block           source/image_processing.c:138-150
 *                           fast_DCT                    source/image_processing.c:192-307
 *                           quantization                source/image_processing.c:349-370
 *                           zigzag_scanning             source/image_processing.c:373-381
 *                           run_length_encode           source/image_processing.c:703-751
 *                           VLC_encode                  source/image_processing.c:400-433
 *                           mpeg1_slice                 source/mpeg1_blk.c:12-20
 *                           encode_macroblock_header_i  source/mpeg1_blk.c:38-58
 *                           encode_block_header_i       source/mpeg1_blk.c:67-113
 *                           encode_block_end            source/mpeg1_blk.c:115-117
 *                           encode_blk_coeff            source/vlc.c:315-385
 *                           encode_coeff_sz_fast        source/vlc.c:146-157
 *                           bitvector_* (bit order)     source/bit_vector.c:13-121
 *   m1cu_ycbcr_planes       write_to_bitstream's input  source/image_processing.c:753-787
 */
#ifndef M1CU_H
#define M1CU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M1CU_ABI_VERSION 1

enum m1cu_mode {
    M1CU_MODE_FULL       = 0, /* raster macroblocks over the coded frame, 4:2:0 chroma          */
    M1CU_MODE_REF_COMPAT = 1  /* the literal traversal of include/encoder.h:238-443 (96x144 ROI) */
};

enum m1cu_status {
    M1CU_OK            =  0,
    M1CU_ERR_ARG       = -1, /* bad argument / geometry                                          */
    M1CU_ERR_CUDA      = -2, /* CUDA runtime failure or no device (see m1cu_last_error)          */
    M1CU_ERR_CAPACITY  = -3, /* caller's output buffer too small                                 */
    M1CU_ERR_LEVEL     = -4  /* a coded AC level had |L| >= 256: the reference returns NULL from
                                encode_blk_coeff (source/vlc.c:383) and crashes; we report it   */
};

enum m1cu_synth_kind { M1CU_SYNTH_NATURAL = 0, M1CU_SYNTH_NOISE = 1,
                       M1CU_SYNTH_GREY = 2,      /* natural's R channel in all three: every pixel is an exact-quotient case */
                       M1CU_SYNTH_RG_EQUAL = 3,  /* natural with G := R                                                   */
                       M1CU_SYNTH_SCATTERED = 4  /* natural with a quarter of its 8x8 pixel tiles replaced by noise      */ };

typedef struct m1cu_ctx m1cu_ctx;

/* ---- host-only helpers (no device needed) ------------------------------------------------ */
int         m1cu_abi_version(void);
int         m1cu_device_count(void);                       /* 0 when there is no usable GPU     */
int         m1cu_qmatrix(int quality_factor, int32_t out[64]);
const char *m1cu_last_error(const m1cu_ctx *ctx);          /* ctx may be NULL: last global error */

/* ---- context ----------------------------------------------------------------------------- */
/* width/height: picture size in pixels; channels: bytes per pixel of the interleaved input
 * (>= 3; R,G,B are the first three, as source/image_processing.c:94-97 indexes them).
 * max_frames: the largest n_frames a single encode call will be given (sizes the staging). */
int  m1cu_create(m1cu_ctx **out, int device, int width, int height, int channels,
                 int mode, int quality_factor, int max_frames);
/* Same, with the kernel's work partition chosen by the caller (tests and tuning sweeps; NULL or zero
 * fields = defaults).  None of the fields changes a result byte. */
typedef struct m1cu_tuning {
    int chunk_mbs;   /* macroblocks per CTA, 1..16 (default 16: full chunks + one shorter tail per slice) */
    int chunk_even;  /* != 0: equal chunks per slice instead                                            */
    int win_words;   /* shared-memory bit-window words per pass, 4..512 (default 512)                   */
    int no_tail_pairing; /* != 0: every slice's short last chunk gets its own CTA (default: two of them share one
                            when they fit a chunk together, e.g. 1080p: 120 macroblocks = 7 x 16 + 8)          */
    int batch_frames;/* pictures per launch round, > 0 lowers the default (about 2 GiB of staging); a call
                        with more pictures runs several rounds, layout + stitch of one beside the encode
                        of the next                                                                     */
    int no_flat_skip;/* != 0: every block goes through DCT + quantiser test (default: a block whose samples span
                        so little that no AC level can be non-zero at this quality -- a rigorous bound, see
                        csrc/m1cu_quant.h -- only gets its DC coefficient computed)                            */
} m1cu_tuning;
int  m1cu_create_ex(m1cu_ctx **out, int device, int width, int height, int channels,
                    int mode, int quality_factor, int max_frames, const m1cu_tuning *tuning);
int  m1cu_destroy(m1cu_ctx *ctx);
int  m1cu_set_stream(m1cu_ctx *ctx, void *cuda_stream);    /* cudaStream_t; NULL = own stream   */
int  m1cu_synchronize(m1cu_ctx *ctx);

/* geometry / sizing */
int    m1cu_macroblocks_per_frame(const m1cu_ctx *ctx);
int    m1cu_flat_range(const m1cu_ctx *ctx);               /* largest sample span of an 8x8 block that proves every AC level
                                                              zero at this quality (such blocks skip the DCT); -1: none, below
                                                              6 grey levels (not worth the test), or switched off            */
size_t m1cu_frame_bytes_in(const m1cu_ctx *ctx);           /* width*height*channels             */
size_t m1cu_payload_bound(const m1cu_ctx *ctx);            /* worst-case payload bytes per frame,
                                                              multiple of 16                    */
size_t m1cu_typical_out_bytes(const m1cu_ctx *ctx, int n_frames); /* a comfortable out_cap      */

/* ---- the hot path, device-resident ------------------------------------------------------- */
/* d_rgb:           n_frames pictures, device memory.
 * d_out/out_cap:   device buffer receiving the payloads; frame f occupies
 *                  [d_frame_offsets[f], d_frame_offsets[f] + d_frame_bytes[f]); offsets are
 *                  16-byte aligned and ascending, d_frame_offsets[n_frames] = end of data.
 * d_frame_bytes:   n_frames uint32, device.   d_frame_offsets: n_frames + 1 uint64, device.
 * d_levels:        optional (NULL in production): int16 [n_frames][macroblocks][6][64], the
 *                  zigzag-ordered quantised levels in coding order (the parity metric).
 * Asynchronous on the context's stream.  Faults detected on the device (capacity, level) are
 * reported by the next m1cu_check(). */
int m1cu_encode_device(m1cu_ctx *ctx, const uint8_t *d_rgb, int n_frames,
                       uint8_t *d_out, size_t out_cap,
                       uint32_t *d_frame_bytes, uint64_t *d_frame_offsets, int16_t *d_levels);

/* Synchronises the stream and returns the sticky device-side status (then clears it). */
int m1cu_check(m1cu_ctx *ctx);

/* ---- the hot path, host buffers (what mpeg_encode_procedure calls) ------------------------- */
/* h_rgb: n_frames pictures in host memory (pinned memory makes the copies asynchronous).
 * h_out receives the payloads back to back WITHOUT padding: frame f starts at
 * sum(h_frame_bytes[0..f)).  h_levels optional as above.  Synchronous.  Returns M1CU_OK or an
 * error; *total_bytes (optional) = sum of the payload sizes. */
int m1cu_encode_host(m1cu_ctx *ctx, const uint8_t *h_rgb, int n_frames,
                     uint8_t *h_out, size_t out_cap, uint32_t *h_frame_bytes,
                     int16_t *h_levels, size_t *total_bytes);

/* Full-resolution Y, Cb, Cr planes of one picture (the content of the reference's .bit side
 * files, source/image_processing.c:753-787).  Device pointers, width*height bytes each. */
int m1cu_ycbcr_planes(m1cu_ctx *ctx, const uint8_t *d_rgb, uint8_t *d_y, uint8_t *d_cb, uint8_t *d_cr);

/* The same planes for the first n_frames pictures of the LAST m1cu_encode_host / m1cu_encode_host_stream call,
 * computed from the copy of the input that call left on the device: one launch and one download per batch,
 * no second upload.  h_planes (capacity cap >= 3*width*height*n_frames) receives per picture Y, Cb, Cr.
 * Synchronous.  What mpeg_encode_procedure uses for the image_%d.bit files (include/encoder.h:460-465). */
int m1cu_host_batch_planes(m1cu_ctx *ctx, int n_frames, uint8_t *h_planes, size_t cap);

/* ---- utilities used by tests and bench.py -------------------------------------------------- */
/* Seeded synthetic RGB written straight into device memory (3 bytes per pixel), frames
 * first_frame .. first_frame + n_frames - 1.  Same integer formula as oracle/m1_oracle.c. */
int m1cu_synth_rgb(m1cu_ctx *ctx, uint32_t seed, long first_frame, int n_frames, int kind, uint8_t *d_rgb);

/* counters: kernels launched by this context since creation (for bench.py's gpu_launches) */
unsigned long long m1cu_launch_count(const m1cu_ctx *ctx);

/* Per-kernel device timing for the roofline report.  While enabled, every launch of the three
 * pipeline kernels is bracketed by CUDA events on the context's stream.  m1cu_kernel_times
 * synchronises, adds the elapsed milliseconds of all launches recorded since the last call into
 * ms[0..2] (0 = k_encode_chunks, 1 = k_layout, 2 = k_stitch) and their counts into n[0..2],
 * then forgets them. */
int m1cu_enable_timing(m1cu_ctx *ctx, int on);
int m1cu_kernel_times(m1cu_ctx *ctx, double ms[3], unsigned long long n[3]);

/* raw device/pinned memory for C callers that have no CUDA headers (the host encoder) */
void *m1cu_device_alloc(size_t bytes);
void  m1cu_device_free(void *p);
void *m1cu_pinned_alloc(size_t bytes);
void  m1cu_pinned_free(void *p);
void *m1cu_pinned_alloc_wc(size_t bytes);  /* write-combined pinned memory (fill it with writes only); free with m1cu_pinned_free */
int   m1cu_memcpy_h2d(void *dst, const void *src, size_t bytes);
int   m1cu_memcpy_d2h(void *dst, const void *src, size_t bytes);

/* Peer memory for frame-range sharding (one process per GPU).  The rank that
 * assembles the stream (rank 0) allocates one receive region per rank with m1cu_device_alloc and
 * exports it; every other rank opens the handle (peer access over NVLink is enabled on first use)
 * and passes its region as d_out of m1cu_encode_device, so k_stitch writes the finished payload
 * bytes straight into rank 0's memory -- no separate gather of the compressed segments.  The
 * handle is CUDA's 64-byte IPC handle; it is only meaningful to other processes on the same box
 * (CUDA refuses to open it in the exporting process). */
#define M1CU_IPC_HANDLE_BYTES 64
int m1cu_ipc_export(void *d_ptr, unsigned char handle[M1CU_IPC_HANDLE_BYTES]);
int m1cu_ipc_open(int device, const unsigned char handle[M1CU_IPC_HANDLE_BYTES], void **d_ptr);
int m1cu_ipc_close(int device, void *d_ptr);

/* Staged variant of the same exchange: copies the payload bytes [0, d_frame_offsets[n_frames]) of a
 * finished m1cu_encode_device call from d_src to dst (normally a peer mapping from m1cu_ipc_open) with a
 * small kernel on `stream` (NULL = the context's stream).  On a high-priority side stream it overlaps
 * the next encode, which a stitch that writes remotely cannot.  The caller orders `stream` after the
 * encode (event) and before the reuse of d_src. */
int m1cu_push_payloads(m1cu_ctx *ctx, void *stream, uint8_t *dst, size_t dst_cap, const uint8_t *d_src,
                       const uint64_t *d_frame_offsets, int n_frames);

/* ---- final stream image on the device (optional) --------------------------- */
/* Turns the result of m1cu_encode_device into the bytes the reference's driver writes to the .mpeg file
 * (include/encoder.h:196-231, :448-458): [h_prologue, 27 bytes, or NULL] then per picture the 44-byte
 * prefix, the payload and the 4-byte trailer.  The headers stay the host's business: h_prefix256 holds 256
 * prefixes of 44 bytes, indexed by (picture index & 255), built with the include/mpeg1_enc.h writers exactly
 * as for a picture with an empty payload; the device only patches the 16-bit packet length
 * (44 + payload - 8, unsigned short arithmetic) into bytes 4..5.  Picture f of this call has index
 * first_frame_index + f.  d_stream (16-byte aligned, stream_cap bytes) receives the image starting at byte
 * stream_offset (any value: a later batch, or another rank's pictures, continue where the previous call
 * ended), *d_stream_bytes (device) the offset one past the last byte written.  Asynchronous on the context's stream; a too small stream_cap is
 * reported by the next m1cu_check() (M1CU_ERR_CAPACITY).  The host copies of the templates are cached:
 * they are uploaded again only when their bytes change. */
int m1cu_assemble_stream(m1cu_ctx *ctx, const uint8_t *d_payloads, const uint32_t *d_frame_bytes,
                         const uint64_t *d_frame_offsets, int n_frames, long first_frame_index,
                         const uint8_t *h_prefix256, const uint8_t *h_prologue, const uint8_t h_trailer[4],
                         uint8_t *d_stream, size_t stream_cap, size_t stream_offset, uint64_t *d_stream_bytes);

/* Host-buffer form of the two calls above: upload, encode, assemble on the device, ONE download of the
 * finished bytes into h_stream (capacity stream_cap); *stream_bytes = their count.  Synchronous.  What
 * mpeg_encode_procedure uses when M1_DEVICE_STREAM=1. */
int m1cu_encode_host_stream(m1cu_ctx *ctx, const uint8_t *h_rgb, int n_frames, long first_frame_index,
                            const uint8_t *h_prefix256, const uint8_t *h_prologue, const uint8_t h_trailer[4],
                            uint8_t *h_stream, size_t stream_cap, size_t *stream_bytes);

#ifdef __cplusplus
}
#endif
#endif /* M1CU_H */
