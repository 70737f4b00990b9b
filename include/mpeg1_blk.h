/*
 * mpeg1_blk.h -- bit-level slice / macroblock / block syntax, same prototypes as the reference's
 * include/mpeg1_blk.h:6-12.  Host-C compatibility entry points; the accelerated path emits the
 * same bits on the GPU.
 */
#ifndef M1_COMPAT_MPEG1_BLK_H
#define M1_COMPAT_MPEG1_BLK_H

#include <stdint.h>
#include <stdlib.h>
#include "bit_vector.h"

#ifdef __cplusplus
extern "C" {
#endif
/* reference source/mpeg1_blk.c:12-20: 000001, (vertical_pos+1)&0xff, 5-bit quant_scale, 0 */
void mpeg1_slice(uint8_t quant_scale, uint8_t vertical_pos, BITVECTOR *out);
/* :38-58: address-increment VLC (escape per 33) + macroblock_type "1" */
void encode_macroblock_header_i(unsigned address, short quant_scale, BITVECTOR *output);
/* :60-62 */
void encode_macroblock_end(BITVECTOR *output);
/* :67-113: DC size + DC bits (or "100"/"00"), then the AC walk */
void encode_block_header_i(unsigned char is_luma, int coeff[128], BITVECTOR *output);
/* :115-117: "10" */
void encode_block_end(BITVECTOR *output);
#ifdef __cplusplus
}
#endif
#endif
