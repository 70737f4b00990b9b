/* vlc.h -- as the reference's include/vlc.h:1-9. */
#ifndef M1_COMPAT_VLC_H
#define M1_COMPAT_VLC_H
#include "bit_vector.h"
#include "jpeg_handler.h"
#ifdef __cplusplus
extern "C" {
#endif
BITVECTOR *encode_macblk_address_value(int value);
#ifdef __cplusplus
}
#endif
#endif
