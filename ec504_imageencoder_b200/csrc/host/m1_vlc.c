/*
 * m1_vlc.c -- VLC tables and lookups of include/mpeg1.h, with the data symbols `make sharedlib`
 * exports in the reference (source/vlc.c).  Host compatibility layer; the GPU path uses the packed
 * tables in csrc/m1cu_tables.h generated from the same listing (tools/gen_vlc_tables.py).
 */
#include "mpeg1.h"

#include <stdio.h>
#include <stdlib.h>

#define V(s) { s, sizeof(s) }

/* macroblock_address_increment 1..33, then stuffing and escape (reference source/vlc.c:33-70) */
struct vlc_macroblock encoding_table[36] = {
    { NULL, 0 },
    V("1"), V("011"), V("010"), V("0011"), V("0010"), V("00011"), V("00010"), V("0000111"), V("0000110"),
    V("00001011"), V("00001010"), V("00001001"), V("00001000"), V("00000111"), V("00000110"),
    V("0000010111"), V("0000010110"), V("0000010101"), V("0000010100"), V("0000010011"), V("0000010010"),
    V("00000100011"), V("00000100010"), V("00000100001"), V("00000100000"), V("00000011111"), V("00000011110"),
    V("00000011101"), V("00000011100"), V("00000011011"), V("00000011010"), V("00000011001"), V("00000011000"),
    V("00000001111"), V("00000001000")
};

/* motion vector magnitudes 0..16 (:87-105); unused by the I-frame path */
struct vlc_macroblock mv_encoding_table[17] = {
    V("1"), V("010"), V("0010"), V("00010"), V("0000110"), V("00001010"), V("00001000"), V("00000110"),
    V("0000010110"), V("0000010100"), V("0000010010"), V("00000100010"), V("00000100000"), V("00000011110"),
    V("00000011100"), V("00000011010"), V("00000011000")
};

/* dct_dc_size_luminance / _chrominance, sizes 0..8 (:121-144) */
struct vlc_macroblock dc_sz_luma_table[9] = {
    V("100"), V("00"), V("01"), V("101"), V("110"), V("1110"), V("11110"), V("111110"), V("1111110") };
struct vlc_macroblock dc_sz_chroma_table[9] = {
    V("00"), V("01"), V("10"), V("110"), V("1110"), V("11110"), V("111110"), V("1111110"), V("11111110") };

/* first row of each run in blk_rle_table (:172-174) */
unsigned int blk_rle_lookup[33] = {
    0, 39, 57, 62, 66, 69, 72, 75, 77, 79, 81, 83, 85, 87, 89, 91, 93, 95, 96, 97, 98, 99, 100, 101, 102, 103, 104,
    105, 106, 107, 108, 109, 110 };

struct vlc_block_rle { unsigned run; unsigned level; struct vlc_block code; };
struct vlc_block_rle blk_rle_table[] = {
#include "m1_vlc_data.inc"
};

struct vlc_block blk_coeff_1_f = { "1", 2 };
struct vlc_block blk_coeff_1_n = { "11", 3 };
struct vlc_block blk_coeff_end = { "10", 2 };

/* :77-85 */
BITVECTOR *encode_macblk_address_value(int value)
{
    if (value < 1 || value > 35) return NULL;
    return bitvector_new(encoding_table[value].binstring, encoding_table[value].bit_len);
}

/* :108-118 */
BITVECTOR *encode_macblk_encoding_value(int value)
{
    if (value < -16 || value > 16) return NULL;
    const int mag = value < 0 ? -value : value;
    BITVECTOR *res = bitvector_new(mv_encoding_table[mag].binstring, mv_encoding_table[mag].bit_len);
    if (value < 0) { bitvector_pos(res, -1); bitvector_put_bit(res, 1); }
    return res;
}

/* :146-157 */
void encode_coeff_sz_fast(BITVECTOR *output, char value, char is_luma)
{
    if (value > 8) { printf("[ERROR] Incorrect coeff size found!!\n"); exit(1); }
    const struct vlc_macroblock *e = is_luma ? &dc_sz_luma_table[(int)value] : &dc_sz_chroma_table[(int)value];
    BITVECTOR *t = bitvector_new(e->binstring, e->bit_len);
    bitvector_concat(output, t);
    free(t->value); free(t);
}

/* :315-385.  run = zeros before the coefficient PLUS ONE (the caller passes the RLE count, >= 1),
 * level = the coefficient.  Table lookups index with |level| - 1, so the run-0 list (which starts
 * at level 2) is shifted by one level -- kept.  No sign bit is appended (:344 is commented out in
 * the reference).  Escape: 000001, 6-bit run, 8- or 16-bit level; NULL when |level| >= 256. */
BITVECTOR *encode_blk_coeff(int run, int level, int first)
{
    if (level == 0) return NULL;
    const int negative = level < 0;
    int mag = negative ? -level : level;
    const int a = mag - 1, r = run - 1;
    const struct vlc_block *code = NULL;
    if (r == 0 && a == 0) code = first ? &blk_coeff_1_f : &blk_coeff_1_n;
    else if (r >= 0 && r <= 31 && (unsigned)a < blk_rle_lookup[r + 1] - blk_rle_lookup[r])
        code = &blk_rle_table[blk_rle_lookup[r] + a].code;
    if (code && code->binstring) return bitvector_new(code->binstring, code->bit_len);
    if (mag >= 256 || r >= 64 || r < 0) return NULL;
    BITVECTOR *res = bitvector_new("000001", 24);
    bitvector_put_byte_off(res, (unsigned char)(r & 0x3f), 6, 2);
    if (mag < 128) {
        unsigned char e = (unsigned char)(mag & 0x7f);
        if (negative) e = (unsigned char)(~e + 1);
        bitvector_put_byte_ent(res, (char)e);
    } else {
        unsigned char e = (unsigned char)mag;
        if (negative) e = (unsigned char)(~e + 1);
        bitvector_put_byte_ent(res, (char)(negative ? 0x80 : 0x00));
        bitvector_put_byte_ent(res, (char)e);
    }
    return res;
}
