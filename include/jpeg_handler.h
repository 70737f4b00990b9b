/*
 * jpeg_handler.h -- the decoded-picture descriptor.  Replaces reference include/jpeg_handler.h:6-15.
 * data is interleaved 8-bit samples, row-major, `channels` bytes per pixel, R,G,B first
 * (reference source/image_processing.c:94-97).  This is the input layout of the accelerated path.
 */
#ifndef M1_COMPAT_JPEG_HANDLER_H
#define M1_COMPAT_JPEG_HANDLER_H

typedef struct {
    int width;
    int height;
    int channels;
    unsigned char *data;
} Image;

#ifdef __cplusplus
extern "C" {
#endif
Image *read_jpeg(const char *filename);            /* stb_image decode into a malloc'd Image, NULL on failure */
int check_dimensions(Image *images[], int count);  /* 1 when all pictures share one size, else 0            */
void free_image(Image *img);
#ifdef __cplusplus
}
#endif
#endif
