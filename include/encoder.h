/*
 * encoder.h -- the driver entry point, same name, signature and return codes as the function the
 * reference DEFINES in its include/encoder.h:20-498:
 *
 *     int mpeg_encode_procedure(const char *images_folder, const char *bitstream_folder,
 *                               const char *video_path, int quality_factor);
 *
 *   0   success (also after creating a missing images folder, reference :111-116)
 *   1   video_path cannot be opened for writing (:77-80)
 *  -1   images folder unreadable, no / mismatching pictures, allocation or GPU failure (:121-135,:175-183)
 *
 * Here the function lives in libencoder (csrc/host/m1_driver.c) and this header only declares it,
 * so main.c / encoder_jni.c keep compiling unchanged (they may still define
 * STB_IMAGE_IMPLEMENTATION before including this file; it is ignored).  The per-picture loop body
 * (reference :216-445) runs on the GPU through include/m1cu.h; file I/O, JPEG decode (stb_image)
 * and every header stay here on the host.
 *
 * Environment (so the C signature stays the reference's):
 *   M1_MODE   "ref_compat" (default: byte-identical to the reference binary, 96x144 region,
 *             uint8 header fields) or "full" (whole coded picture, raster macroblocks, 4:2:0 chroma)
 *   M1_DEVICE CUDA device index (default 0)
 *   M1_BIT_FILES  "0" disables the image_%d.bit side files (default: written, as the reference does)
 */
#ifndef M1_COMPAT_ENCODER_H
#define M1_COMPAT_ENCODER_H

#include <dirent.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/stat.h>
#include <unistd.h>

#include "bit_vector.h"
#include "image_processing.h"
#include "jpeg_handler.h"
#include "mpeg1.h"

#ifdef __cplusplus
extern "C" {
#endif

int mpeg_encode_procedure(const char *images_folder, const char *bitstream_folder,
                          const char *video_path, int quality_factor);

/* Same stream from pictures already in memory (n_frames x height x width x channels bytes);
 * mode: 0 = full, 1 = ref_compat.  Used by the tests and by callers that decode elsewhere. */
int m1_encode_frames_to_file(const char *video_path, const unsigned char *frames, int n_frames,
                             int width, int height, int channels, int quality_factor, int mode);

/* Stream image in memory: returns bytes written to out (capacity cap), or a negative m1cu status. */
long m1_encode_frames_to_memory(const unsigned char *frames, int n_frames, int width, int height,
                                int channels, int quality_factor, int mode, unsigned char *out, long cap);

/* The header bytes m1cu_assemble_stream (include/m1cu.h) needs to build the stream image on the device:
 * prefix256 = 256 x 44 bytes, entry i = the packet / sequence / GOP / picture headers in front of picture
 * index i (the reference's clock is a uint8_t hour, include/encoder.h:42, so they repeat every 256 pictures;
 * the packet length is left for an empty payload), prologue = the 27 file bytes of include/encoder.h:86-88,
 * trailer = the 4 bytes after every picture.  All written by the include/mpeg1_enc.h functions. */
void m1_stream_templates(int width, int height, int mode, unsigned char prefix256[256 * 44],
                         unsigned char prologue[27], unsigned char trailer[4]);

#ifdef __cplusplus
}
#endif
#endif
