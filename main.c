/* main.c -- command-line entry point, same behaviour as the reference's main.c:15-17:
 * encode the JPEGs of ./images into bitstreams/awesome_video.mpeg at quality factor 12. */
#define STB_IMAGE_IMPLEMENTATION   /* kept for source compatibility; the decoder lives in libencoder */
#include "encoder.h"

int main(void)
{
    return mpeg_encode_procedure("images/", "bitstreams", "bitstreams/awesome_video.mpeg", 12);
}
