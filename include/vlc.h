/* vlc.h -- drop-in for the reference's include/vlc.h (lines 1-9): the one function that header
 * declares.  Implemented in ec504_imageencoder_b200/csrc/host/m1_vlc.c as a host compatibility entry
 * point; the GPU path never calls it (k_encode_chunks emits the two macroblock-header bits itself).
 *
 * encode_macblk_address_value(value): a freshly allocated BITVECTOR holding the MPEG-1
 * macroblock_address_increment code of `value` (reference source/vlc.c:77-85, table :33-70): the
 * 11-bit escape 00000001000 once per 33 above 33, then the VLC of the remainder.  The caller owns the
 * result (the reference never frees it).  The driver only ever asks for value 1, the single bit '1'.
 */
#ifndef M1_COMPAT_VLC_H
#define M1_COMPAT_VLC_H

#include "bit_vector.h"      /* BITVECTOR */
#include "jpeg_handler.h"    /* pulled in by the reference's header as well; kept for source compatibility */

#ifdef __cplusplus
extern "C" {
#endif

BITVECTOR *encode_macblk_address_value(int value);

#ifdef __cplusplus
}
#endif

#endif /* M1_COMPAT_VLC_H */
