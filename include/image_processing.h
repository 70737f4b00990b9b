/*
 * image_processing.h -- per-stage functions of the block pipeline, same prototypes as the
 * reference's include/image_processing.h:8-30 (what `make sharedlib` exports).
 *
 * These are host-C compatibility entry points for callers that drive the stages one block at a
 * time, as the reference's own driver does.  They are NOT the accelerated path: whole pictures
 * go through the CUDA library (include/m1cu.h), which mpeg_encode_procedure uses.
 * Ownership follows the reference: convert_rgb_to_ycbcr / subsampling_420 malloc their outputs
 * (caller frees); everything else fills caller buffers.
 */
#ifndef M1_COMPAT_IMAGE_PROCESSING_H
#define M1_COMPAT_IMAGE_PROCESSING_H

#include "bit_vector.h"
#include "jpeg_handler.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference source/image_processing.c:68-110 */
void convert_rgb_to_ycbcr(Image *img, unsigned char **Y, unsigned char **Cb, unsigned char **Cr);
/* :114-133 */
void subsampling_420(unsigned char *Cb, unsigned char *Cr, int width, int height,
                     unsigned char **Cb_sub, unsigned char **Cr_sub);
/* :138-150 */
void extract_8x8_block(unsigned char *channel, int image_width, int start_x, int start_y, unsigned char block[8][8]);
/* :192-307 -- integer butterflies; outputs are exact integers stored as double */
void fast_DCT(const unsigned char block[8][8], double dct_block[8][8]);
/* :314-343 */
void scale_quantization_matrix(int scaled_q_matrix[8][8], int quality_factor);
/* :349-370 */
void quantization(double dct_block[8][8], int quantized_block[8][8], int quality_factor);
/* :373-381 */
void zigzag_scanning(int quantized_block[8][8], int zigzag_array[64]);
/* :385-398 (identity) */
void equalize_coefficients(int zigzag_array[64], int equalized_array[64]);
/* :703-751 -- (level, zeros-before) pairs, terminated by two zeros; needs room for 130 ints */
int *run_length_encode(int array[64], int encode_array[128]);
/* :400-433 */
void VLC_encode(int RLE_array[128], BITVECTOR *temp_dest_bv);
/* :753-787 -- int32 width, int32 height, then the Y, Cb, Cr planes */
void write_to_bitstream(const char *filename, unsigned char *Y, unsigned char *Cb, unsigned char *Cr, int width, int height);
/* :695-700 */
void print_array(int array[], int size);

#ifdef __cplusplus
}
#endif
#endif
