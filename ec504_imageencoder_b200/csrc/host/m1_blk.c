/*
 * m1_blk.c -- bit-level slice / macroblock / block syntax of include/mpeg1_blk.h.
 * Host compatibility layer written from SURVEY.md appendix A; the accelerated path produces the
 * same bits in csrc/m1cu_kernels.cu (code_block, k_encode_chunks).
 */
#include "image_processing.h"
#include "mpeg1.h"

#include <stdint.h>
#include <stdlib.h>

struct bitvector slice_start_code = { "\x00\x00\x01", 24, 24, 24 };   /* reference source/mpeg1_blk.c:10 */

static void append_and_free(BITVECTOR *out, BITVECTOR *t)
{
    if (!t) return;
    bitvector_concat(out, t);
    free(t->value);
    free(t);
}

/* reference source/mpeg1_blk.c:12-20 (the reference also prints the whole buffer here) */
void mpeg1_slice(uint8_t quant_scale, uint8_t vertical_pos, BITVECTOR *out)
{
    bitvector_concat(out, &slice_start_code);
    bitvector_put_byte_ent(out, (char)(vertical_pos + 1));
    bitvector_put_byte_off(out, (unsigned char)(quant_scale & 0x1f), 5, 3);
    bitvector_put_bit(out, 0);
}

/* :38-58 -- one escape per 33 of address, the increment code, then macroblock_type "1" */
void encode_macroblock_header_i(unsigned address, short quant_scale, BITVECTOR *output)
{
    (void)quant_scale;
    while (address > 33) { append_and_free(output, encode_macblk_address_value(35)); address -= 33; }
    append_and_free(output, encode_macblk_address_value((int)address));
    bitvector_put_bit(output, 1);
}

/* :60-62 */
void encode_macroblock_end(BITVECTOR *output) { bitvector_put_bit(output, 1); }

/* :67-113 -- coeff holds run_length_encode's (level, zeros-before) pairs */
void encode_block_header_i(unsigned char is_luma, int coeff[128], BITVECTOR *output)
{
    if (coeff[0] != 0 && coeff[1] == 0) {
        int mag = coeff[0] < 0 ? -coeff[0] : coeff[0];
        int sz = 1;
        for (int i = 1; i <= 8; ++i) if (mag & (1 << (i - 1))) sz = i;      /* bits 0..7 only, default 1 */
        encode_coeff_sz_fast(output, (char)sz, (char)is_luma);
        if (coeff[0] < 0) mag ^= 1 << (sz - 1);                             /* clears the top bit        */
        bitvector_put_byte_off(output, (unsigned char)(mag & 0xff), (char)sz, (char)(8 - sz));
        VLC_encode(coeff + 2, output);
    } else {
        bitvector_put_binstring(output, is_luma ? "100" : "00");
        VLC_encode(coeff, output);
    }
}

/* :115-117 */
void encode_block_end(BITVECTOR *output) { bitvector_put_binstring(output, "10"); }
