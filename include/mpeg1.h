/*
 * mpeg1.h -- VLC table entry types and lookup functions, as the reference's include/mpeg1.h:27-51.
 */
#ifndef M1_COMPAT_MPEG1_H
#define M1_COMPAT_MPEG1_H

#include <stdint.h>
#include <stdio.h>
#include "bit_vector.h"
#include "mpeg1_blk.h"
#include "mpeg1_enc.h"

struct vlc_macroblock { const char *binstring; unsigned bit_len; };
struct vlc_block      { const char *binstring; unsigned bit_len; };

#ifdef __cplusplus
extern "C" {
#endif
BITVECTOR *encode_macblk_address_value(int value);            /* reference source/vlc.c:77-85   */
BITVECTOR *encode_macblk_encoding_value(int value);           /* :108-118 (motion vectors)      */
BITVECTOR *encode_blk_coeff(int run, int level, int first);   /* :315-385; NULL when |level| >= 256 */
void encode_coeff_sz_fast(BITVECTOR *output, char value, char is_luma);   /* :146-157          */
#ifdef __cplusplus
}
#endif
#endif
