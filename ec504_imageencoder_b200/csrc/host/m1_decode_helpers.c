/*
 * m1_decode_helpers.c -- the decoder-side / reference-only helpers that the reference's
 * `make sharedlib` exports but its driver never calls.  Exported for
 * ABI completeness: plain host C, no GPU involvement.  Arithmetic follows the reference function
 * cited at each definition, including where that function is not a true inverse.
 */
#include "image_processing.h"
#include "global_variables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define M1_PI 3.14159265358979323846

void DCT(const unsigned char block[64], float dct_block[64]);
void IDCT(const float dct_block[64], unsigned char block[64]);
void dequantization(int quantized_block[8][8], double dct_block[8][8]);
void fast_IDCT(const double dct_block[8][8], unsigned char block[8][8]);
void upsampling(unsigned char *Cb_sub, unsigned char *Cr_sub, int width, int height, unsigned char **Cb, unsigned char **Cr);
void insert_8x8_block(unsigned char *channel, int image_width, int start_x, int start_y, unsigned char block[8][8]);
void convert_ycbcr_to_rgb(unsigned char *Y, unsigned char *Cb, unsigned char *Cr, Image *img);
char *concat_char(char *array1, char *array2);

static float norm(int k) { return (float)(k == 0 ? sqrt(1.0 / 8) : sqrt(2.0 / 8)); }

/* reference source/image_processing.c:157-183 -- direct O(n^4) float DCT, block[y*8+x], out[v*8+u] */
void DCT(const unsigned char block[64], float dct_block[64])
{
    for (int u = 0; u < 8; ++u)
        for (int v = 0; v < 8; ++v) {
            float sum = 0.0f;
            const float cu = norm(u), cv = norm(v);
            for (int x = 0; x < 8; ++x)
                for (int y = 0; y < 8; ++y) {
                    const float pixel = (float)block[y * 8 + x];
                    sum += pixel * cos((2 * x + 1) * u * M1_PI / (2.0 * 8)) * cos((2 * y + 1) * v * M1_PI / (2.0 * 8));
                }
            dct_block[v * 8 + u] = cu * cv * sum;
        }
}

/* :452-481 -- direct float inverse with rounding and clamping */
void IDCT(const float dct_block[64], unsigned char block[64])
{
    for (int x = 0; x < 8; ++x)
        for (int y = 0; y < 8; ++y) {
            float sum = 0.0f;
            for (int u = 0; u < 8; ++u)
                for (int v = 0; v < 8; ++v) {
                    const float cu = norm(u), cv = norm(v);
                    sum += cu * cv * dct_block[v * 8 + u] * cos((2 * x + 1) * u * M1_PI / (2.0 * 8)) *
                           cos((2 * y + 1) * v * M1_PI / (2.0 * 8));
                }
            const int p = (int)round(sum);
            block[y * 8 + x] = (unsigned char)(p < 0 ? 0 : (p > 255 ? 255 : p));
        }
}

/* :438-446 -- multiplies by the UNSCALED default matrix */
void dequantization(int quantized_block[8][8], double dct_block[8][8])
{
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) dct_block[i][j] = quantized_block[i][j] * Q_MATRIX[i][j];
}

/* The reference's int arithmetic overflows for large coefficients (undefined behaviour that every gcc build resolves as
 * two's-complement wrap-around): reproduced with unsigned arithmetic, which wraps by definition. */
static inline int w_add(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
static inline int w_sub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
static inline int w_mul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }

static void pass8(const int x[8], int o[8])
{
    /* the same butterfly as fast_DCT (the reference's fast_IDCT reuses the forward stages) */
    enum { c1 = 1004, s1 = 200, c3 = 851, s3 = 569, r2c6 = 554, r2s6 = 1337 };
    const int s07 = w_add(x[0], x[7]), d07 = w_sub(x[0], x[7]), s16 = w_add(x[1], x[6]), d16 = w_sub(x[1], x[6]);
    const int s25 = w_add(x[2], x[5]), d25 = w_sub(x[2], x[5]), s34 = w_add(x[3], x[4]), d34 = w_sub(x[3], x[4]);
    const int ee = w_add(s07, s34), eo = w_sub(s07, s34), oe = w_add(s16, s25), oo = w_sub(s16, s25);
    const int ta = w_mul(c1, w_add(d16, d25)), tb = w_mul(c3, w_add(d07, d34)), tc = w_mul(r2c6, w_add(oo, eo));
    const int y2 = w_add(w_mul(-s1 - c1, d25), ta), y1 = w_add(w_mul(s1 - c1, d16), ta);
    const int y3 = w_add(w_mul(-s3 - c3, d34), tb), y0 = w_add(w_mul(s3 - c3, d07), tb);
    o[0] = w_add(ee, oe); o[1] = w_sub(ee, oe);
    o[2] = w_add(w_mul(r2s6 - r2c6, eo), tc); o[3] = w_add(w_mul(-r2s6 - r2c6, oo), tc);
    o[4] = w_add(y3, y1); o[5] = w_add(y0, y2); o[6] = w_sub(y3, y1); o[7] = w_sub(y0, y2);
}

/* :492-601 -- columns then rows through the FORWARD butterfly, then clamps / shifts as the
 * reference writes them (so this is not the inverse of fast_DCT; reproduced, not corrected) */
void fast_IDCT(const double dct_block[8][8], unsigned char block[8][8])
{
    enum { r2 = 181 };
    int cols[8][8], in[8], o[8];
    for (int i = 0; i < 8; ++i) {
        for (int k = 0; k < 8; ++k) in[k] = (int)dct_block[k][i];
        pass8(in, o);
        cols[0][i] = o[0]; cols[4][i] = o[1]; cols[2][i] = o[2] >> 10; cols[6][i] = o[3] >> 10;
        cols[7][i] = w_sub(o[4], o[5]) >> 10; cols[1][i] = w_add(o[4], o[5]) >> 10;
        cols[3][i] = w_mul(o[6], r2) >> 17; cols[5][i] = w_mul(o[7], r2) >> 17;
    }
    for (int i = 0; i < 8; ++i) {
        pass8(cols[i], o);
        block[i][0] = (unsigned char)(o[0] < 0 ? 0 : (o[0] > 255 ? 255 : o[0]));
        block[i][4] = (unsigned char)(o[1] < 0 ? 0 : (o[1] > 255 ? 255 : o[1]));
        block[i][2] = (unsigned char)(o[2] >> 10);
        block[i][6] = (unsigned char)(o[3] >> 10);
        block[i][7] = (unsigned char)(w_sub(o[4], o[5]) >> 10);
        block[i][1] = (unsigned char)(w_add(o[4], o[5]) >> 10);
        block[i][3] = (unsigned char)(w_mul(o[6], r2) >> 17);
        block[i][5] = (unsigned char)(w_mul(o[7], r2) >> 17);
    }
}

/* :607-638 -- each subsampled value fills its 2x2 */
void upsampling(unsigned char *Cb_sub, unsigned char *Cr_sub, int width, int height, unsigned char **Cb, unsigned char **Cr)
{
    const int sw = width / 2, sh = height / 2;
    *Cb = (unsigned char *)malloc((size_t)width * height);
    *Cr = (unsigned char *)malloc((size_t)width * height);
    for (int y = 0; y < sh; ++y)
        for (int x = 0; x < sw; ++x)
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    const size_t o = (size_t)(2 * y + dy) * width + 2 * x + dx;
                    (*Cb)[o] = Cb_sub[y * sw + x];
                    (*Cr)[o] = Cr_sub[y * sw + x];
                }
}

/* :641-647 */
void insert_8x8_block(unsigned char *channel, int image_width, int start_x, int start_y, unsigned char block[8][8])
{
    for (int i = 0; i < 8; ++i) memcpy(channel + (size_t)(start_y + i) * image_width + start_x, block[i], 8);
}

/* :650-692.  The reference reads y, cb and cr from img->data right after allocating it (:661-673),
 * i.e. from uninitialised memory, and never looks at its Y/Cb/Cr arguments; there is no defined
 * result to reproduce.  This implements the conversion its comments describe, from the arguments. */
void convert_ycbcr_to_rgb(unsigned char *Y, unsigned char *Cb, unsigned char *Cr, Image *img)
{
    if (img->channels < 3) {
        printf("Error: Image does not have correct color channels for YCbCr to RGB conversion.\n");
        return;
    }
    const size_t n = (size_t)img->width * img->height;
    img->data = (unsigned char *)malloc(n * img->channels);
    if (!img->data) { printf("Error: Failed to allocate memory for RGB image.\n"); return; }
    for (size_t i = 0; i < n; ++i) {
        const int y = Y[i], cb = Cb[i], cr = Cr[i];
        int r = (int)(y + 1.402 * (cr - 128));
        int g = (int)(y - 0.344136 * (cb - 128) - 0.714136 * (cr - 128));
        int b = (int)(y + 1.772 * (cb - 128));
        r = r < 0 ? 0 : (r > 255 ? 255 : r);
        g = g < 0 ? 0 : (g > 255 ? 255 : g);
        b = b < 0 ? 0 : (b > 255 ? 255 : b);
        unsigned char *p = img->data + i * img->channels;
        p[0] = (unsigned char)r; p[1] = (unsigned char)g; p[2] = (unsigned char)b;
        for (int c = 3; c < img->channels; ++c) p[c] = 255;
    }
}

/* reference source/mpeg1_enc.c:145-157 returns the address of a local array (undefined behaviour,
 * never called).  Kept as a symbol; returns a malloc'd concatenation of the two C strings. */
char *concat_char(char *array1, char *array2)
{
    const size_t a = array1 ? strlen(array1) : 0, b = array2 ? strlen(array2) : 0;
    char *out = (char *)malloc(a + b + 1);
    if (!out) return NULL;
    if (a) memcpy(out, array1, a);
    if (b) memcpy(out + a, array2, b);
    out[a + b] = 0;
    return out;
}
