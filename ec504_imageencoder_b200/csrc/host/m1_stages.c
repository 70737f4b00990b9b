/*
 * m1_stages.c -- the per-stage functions of include/image_processing.h, one block / one picture
 * at a time on the host, for callers that use the reference's `make sharedlib` API directly.
 * NOT the accelerated path (that is csrc/m1cu_kernels.cu behind include/m1cu.h).  Built with
 * -ffp-contract=off: the colour conversion must round every product and sum separately.
 */
#include "global_variables.h"
#include "image_processing.h"
#include "mpeg1.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* reference source/image_processing.c:17-26, :28-37 */
const int Q_MATRIX[8][8] = {
    {  8, 16, 19, 22, 26, 27, 29, 34 }, { 16, 16, 22, 24, 27, 29, 34, 37 },
    { 19, 22, 26, 27, 29, 34, 34, 38 }, { 22, 22, 26, 27, 29, 34, 37, 40 },
    { 22, 26, 27, 29, 32, 35, 40, 48 }, { 26, 27, 29, 32, 35, 40, 48, 58 },
    { 26, 27, 29, 34, 38, 46, 56, 69 }, { 27, 29, 35, 38, 46, 56, 69, 83 } };
const int ZIGZAG_ORDER[8][8] = {
    {  0,  1,  5,  6, 14, 15, 27, 28 }, {  2,  4,  7, 13, 16, 26, 29, 42 },
    {  3,  8, 12, 17, 25, 30, 41, 43 }, {  9, 11, 18, 24, 31, 40, 44, 53 },
    { 10, 19, 23, 32, 39, 45, 52, 54 }, { 20, 22, 33, 38, 46, 51, 55, 60 },
    { 21, 34, 37, 47, 50, 56, 59, 61 }, { 35, 36, 48, 49, 57, 58, 62, 63 } };
/* reference source/global_variables.c:3-4 (overflowing char constants, unused) */
const char START_FILE = (char)0x000001ba;
const char START_PICTURE = (char)0x00000100;

/* :48-66 */
int check_dimensions(Image *images[], int count)
{
    if (count == 0) { printf("No images found in directory.\n"); return 0; }
    for (int i = 1; i < count; ++i)
        if (images[i]->width != images[0]->width || images[i]->height != images[0]->height) {
            printf("Error: Image dimensions do not match\n");
            return 0;
        }
    printf("Images have matching dimensions of width = %d and height = %d\n", images[0]->width, images[0]->height);
    return 1;
}

/* :68-110 -- IEEE double, left to right, truncation; outputs malloc'd, caller frees */
void convert_rgb_to_ycbcr(Image *img, unsigned char **Y, unsigned char **Cb, unsigned char **Cr)
{
    if (img->channels < 3) {
        printf("Error: Image does not have correct color channels for RBG to YCbCr conversion.\n");
        return;
    }
    const long n = (long)img->width * img->height;
    *Y = (unsigned char *)malloc((size_t)n);
    *Cb = (unsigned char *)malloc((size_t)n);
    *Cr = (unsigned char *)malloc((size_t)n);
    if (!*Y || !*Cb || !*Cr) {
        printf("Error: memory allocation failed for YCbCr components.\n");
        free(*Y); free(*Cb); free(*Cr);
        return;
    }
    for (long i = 0; i < n; ++i) {
        const unsigned char *p = img->data + i * img->channels;
        const double r = p[0], g = p[1], b = p[2];
        (*Y)[i]  = (unsigned char)(0.299 * r + 0.587 * g + 0.114 * b);
        (*Cb)[i] = (unsigned char)(128 - 0.168736 * r - 0.331264 * g + 0.5 * b);
        (*Cr)[i] = (unsigned char)(128 + 0.5 * r - 0.418688 * g - 0.081312 * b);
    }
}

/* :114-133 -- truncating mean of each 2x2 */
void subsampling_420(unsigned char *Cb, unsigned char *Cr, int width, int height,
                     unsigned char **Cb_sub, unsigned char **Cr_sub)
{
    const int sw = width / 2, sh = height / 2;
    *Cb_sub = (unsigned char *)malloc((size_t)sw * sh);
    *Cr_sub = (unsigned char *)malloc((size_t)sw * sh);
    for (int y = 0; y + 1 < height; y += 2)
        for (int x = 0; x + 1 < width; x += 2) {
            const int a = y * width + x, b = a + width, o = (y / 2) * sw + x / 2;
            (*Cb_sub)[o] = (unsigned char)((Cb[a] + Cb[a + 1] + Cb[b] + Cb[b + 1]) / 4);
            (*Cr_sub)[o] = (unsigned char)((Cr[a] + Cr[a + 1] + Cr[b] + Cr[b + 1]) / 4);
        }
}

/* :138-150 */
void extract_8x8_block(unsigned char *channel, int image_width, int start_x, int start_y, unsigned char block[8][8])
{
    for (int i = 0; i < 8; ++i) memcpy(block[i], channel + (long)(start_y + i) * image_width + start_x, 8);
}

/* one 8-point pass of :210-238; o[] = k0, k4, k2, k6 (unshifted), x2, x5, x3, x0 */
static void butterfly8(const int x[8], int o[8])
{
    enum { c1 = 1004, s1 = 200, c3 = 851, s3 = 569, r2c6 = 554, r2s6 = 1337 };
    const int s07 = x[0] + x[7], d07 = x[0] - x[7], s16 = x[1] + x[6], d16 = x[1] - x[6];
    const int s25 = x[2] + x[5], d25 = x[2] - x[5], s34 = x[3] + x[4], d34 = x[3] - x[4];
    const int ee = s07 + s34, eo = s07 - s34, oe = s16 + s25, oo = s16 - s25;
    const int ta = c1 * (d16 + d25), tb = c3 * (d07 + d34), tc = r2c6 * (oo + eo);
    const int y2 = (-s1 - c1) * d25 + ta, y1 = (s1 - c1) * d16 + ta;
    const int y3 = (-s3 - c3) * d34 + tb, y0 = (s3 - c3) * d07 + tb;
    o[0] = ee + oe; o[1] = ee - oe;
    o[2] = (r2s6 - r2c6) * eo + tc; o[3] = (-r2s6 - r2c6) * oo + tc;
    o[4] = y3 + y1; o[5] = y0 + y2; o[6] = y3 - y1; o[7] = y0 - y2;
}

/* :192-307 */
void fast_DCT(const unsigned char block[8][8], double dct_block[8][8])
{
    enum { r2 = 181 };
    int rows[8][8], in[8], o[8];
    for (int i = 0; i < 8; ++i) {
        for (int j = 0; j < 8; ++j) in[j] = block[i][j];
        butterfly8(in, o);
        rows[i][0] = o[0]; rows[i][4] = o[1]; rows[i][2] = o[2] >> 10; rows[i][6] = o[3] >> 10;
        rows[i][7] = (o[4] - o[5]) >> 10; rows[i][1] = (o[4] + o[5]) >> 10;
        rows[i][3] = (o[6] * r2) >> 17; rows[i][5] = (o[7] * r2) >> 17;
    }
    for (int j = 0; j < 8; ++j) {
        for (int i = 0; i < 8; ++i) in[i] = rows[i][j];
        butterfly8(in, o);
        dct_block[0][j] = (double)((o[0] + 16) >> 3);
        dct_block[4][j] = (double)((o[1] + 16) >> 3);
        dct_block[2][j] = (double)((o[2] + 16384) >> 13);
        dct_block[6][j] = (double)((o[3] + 16384) >> 13);
        dct_block[7][j] = (double)((o[4] - o[5] + 16384) >> 13);
        dct_block[1][j] = (double)((o[4] + o[5] + 16384) >> 13);
        dct_block[3][j] = (double)(((o[6] >> 8) * r2 + 8192) >> 12);
        dct_block[5][j] = (double)(((o[7] >> 8) * r2 + 8192) >> 12);
    }
}

/* :314-343 -- float scale factor, float product, double division, round half away, floor of 1 */
void scale_quantization_matrix(int scaled_q_matrix[8][8], int quality_factor)
{
    if (quality_factor < 1) quality_factor = 1;
    if (quality_factor > 100) quality_factor = 100;
    const float scale = quality_factor < 50 ? (float)(5000.0 / quality_factor) : (float)(200.0 - 2 * quality_factor);
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            const float prod = (float)Q_MATRIX[i][j] * scale;
            const int v = (int)round((double)prod / 100.0);
            scaled_q_matrix[i][j] = v < 1 ? 1 : v;
        }
}

/* :349-370 */
void quantization(double dct_block[8][8], int quantized_block[8][8], int quality_factor)
{
    int m[8][8];
    scale_quantization_matrix(m, quality_factor);
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) quantized_block[i][j] = (int)(round(dct_block[i][j]) / m[i][j]);
}

/* :373-381 */
void zigzag_scanning(int quantized_block[8][8], int zigzag_array[64])
{
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) zigzag_array[ZIGZAG_ORDER[i][j]] = quantized_block[i][j];
}

/* :385-398 -- identity (the +-1 adjustment is commented out in the reference) */
void equalize_coefficients(int zigzag_array[64], int equalized_array[64])
{
    memcpy(equalized_array, zigzag_array, 64 * sizeof(int));
}

/* :703-751 -- (level, zeros since the previous non-zero) pairs + two terminating zeros */
int *run_length_encode(int zigzag_block[64], int encoded_array[128])
{
    int n = 0, zeros = 0;
    for (int i = 0; i < 64; ++i) {
        if (zigzag_block[i] != 0) { encoded_array[n++] = zigzag_block[i]; encoded_array[n++] = zeros; zeros = 0; }
        else ++zeros;
    }
    if (n < 128) encoded_array[n] = 0;
    if (n + 1 < 128) encoded_array[n + 1] = 0;
    return encoded_array;
}

/* :400-433 -- stops at the first pair whose run or level is zero */
void VLC_encode(int RLE_array[128], BITVECTOR *temp_dest_bv)
{
    for (int k = 0; k < 64; ++k) {
        const int level = RLE_array[2 * k], run = RLE_array[2 * k + 1];
        if (run == 0 || level == 0) break;
        BITVECTOR *t = encode_blk_coeff(run, level, 0);
        if (!t) break;      /* |level| >= 256: the reference dereferences NULL here; we stop the block */
        bitvector_concat(temp_dest_bv, t);
        free(t->value);
        free(t);
    }
}

/* :753-787 */
void write_to_bitstream(const char *filename, unsigned char *Y, unsigned char *Cb, unsigned char *Cr, int width, int height)
{
    FILE *f = fopen(filename, "wb");
    if (!f) { printf("Error: Could not open bitstream file.\n"); return; }
    const size_t n = (size_t)width * height;
    fwrite(&width, sizeof(int), 1, f);
    fwrite(&height, sizeof(int), 1, f);
    fwrite(Y, 1, n, f); fwrite(Cb, 1, n, f); fwrite(Cr, 1, n, f);
    fclose(f);
}

/* :695-700 */
void print_array(int arr[], int size)
{
    for (int i = 0; i < size; ++i) printf("%d ", arr[i]);
    printf("\n");
}
