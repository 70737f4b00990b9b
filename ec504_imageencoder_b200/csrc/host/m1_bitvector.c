/*
 * m1_bitvector.c -- the growable MSB-first bit buffer of include/bit_vector.h.
 * Behaviour follows reference source/bit_vector.c (cited per function); the storage management is
 * our own (grow-to-fit instead of the reference's single doubling), which no caller can observe
 * through the stream contents.  Host compatibility layer, not on the accelerated path.
 */
#include "bit_vector.h"

#include <string.h>

static void reserve_bits(BITVECTOR *bv, long long int need_bits)
{
    if (need_bits < bv->bits) return;
    long long int nb = bv->bits > 0 ? bv->bits : 8;
    while (nb <= need_bits) nb <<= 1;
    char *p = (char *)realloc(bv->value, (size_t)(nb >> 3) + 1);
    if (!p) { printf("REALLOC FAILED"); return; }
    memset(p + (bv->bits >> 3) + 1, 0, (size_t)((nb >> 3) - (bv->bits >> 3)));
    bv->value = p;
    bv->bits = nb;
}

/* source/bit_vector.c:7-11 */
void bitvector_init(BITVECTOR *bv, long long int size)
{
    bv->cap = bv->cursor = 0;
    bv->bits = size;
    bv->value = (char *)calloc((size_t)(size >> 3) + 1, 1);
}

/* :148-170 -- capacity doubles */
void bitvector_expand_size(BITVECTOR *bv, long long int speculative)
{
    (void)speculative;
    reserve_bits(bv, bv->bits);
}

/* :13-27 */
void bitvector_put_bit(BITVECTOR *bv, char bit)
{
    reserve_bits(bv, bv->cursor + 1);
    const long long int byte = bv->cursor >> 3;
    const int sh = 7 - (int)(bv->cursor & 7);
    if (bit) bv->value[byte] = (char)(bv->value[byte] | (1 << sh));
    else     bv->value[byte] = (char)(bv->value[byte] & ~(1 << sh));
    bv->cursor++;
    if (bv->cap < bv->cursor) bv->cap = bv->cursor;
}

/* :29-40 -- '1' (or the byte value 1) is a one, anything else a zero */
void bitvector_put_binstring(BITVECTOR *bv, const char *bitstring)
{
    for (const char *p = bitstring; *p; ++p) bitvector_put_bit(bv, (char)(*p == '1' || *p == 1));
    if (bv->cap < bv->cursor) bv->cap = bv->cursor;
}

/* :44-83 -- appends `bits` bits of val, starting `offset` bits below its MSB */
void bitvector_put_byte_off(BITVECTOR *bv, unsigned char val, char bits, char offset)
{
    const unsigned v = ((unsigned)val >> (8 - offset - bits)) & ((1u << bits) - 1u);
    for (int k = bits - 1; k >= 0; --k) bitvector_put_bit(bv, (char)((v >> k) & 1u));
}

void bitvector_put_byte(BITVECTOR *bv, char val, char bits) { bitvector_put_byte_off(bv, (unsigned char)val, bits, 0); } /* :87-89 */
void bitvector_put_byte_ent(BITVECTOR *bv, char val) { bitvector_put_byte_off(bv, (unsigned char)val, 8, 0); }           /* :91-93 */

/* :93-97 */
long long int bitvector_pos(BITVECTOR *bv, long long int off)
{
    bv->cursor += off;
    if (bv->cap < bv->cursor) bv->cap = bv->cursor;
    return bv->cursor;
}

/* :100-121 -- append bits [0, src->cap) of src at dest's cursor */
void bitvector_concat(BITVECTOR *dest, BITVECTOR *src)
{
    for (long long int k = 0; k < src->cap; ++k)
        bitvector_put_bit(dest, (char)((src->value[k >> 3] >> (7 - (k & 7))) & 1));
}

/* :124-134 -- note the reference masks the LOW bits of the partial byte; kept */
int bitvector_toarray(BITVECTOR *bv, char *output)
{
    int total = (int)(bv->cap >> 3);
    memcpy(output, bv->value, (size_t)total);
    if (bv->cap & 7) {
        const int n = (int)(bv->cap & 7);
        output[total] = (char)(bv->value[total] & ~((1 << n) - 1));
        total++;
    }
    return total;
}

/* :136-146 -- whole bytes, then (when the length is not a multiple of 8) the reference writes the
 * FIRST byte of the buffer once more; kept.  Returns cap >> 3 like the reference. */
int bitvector_fwrite(BITVECTOR *bv, FILE *file)
{
    const int total = (int)(bv->cap >> 3);
    fwrite(bv->value, sizeof(char), (size_t)total, file);
    if (bv->cap & 7) fwrite(bv->value, sizeof(char), 1, file);
    return total;
}

/* :172-178 */
BITVECTOR *bitvector_clone(BITVECTOR *bv)
{
    BITVECTOR *n = (BITVECTOR *)malloc(sizeof(BITVECTOR));
    bitvector_init(n, bv->bits);
    n->cap = n->cursor = bv->cap;
    memcpy(n->value, bv->value, (size_t)(bv->bits / 8) + (bv->bits % 8 ? 1 : 0));
    return n;
}

/* :180-185 -- `size` is a capacity hint, the length is strlen(binstring) */
BITVECTOR *bitvector_new(const char *binstring, long long int size)
{
    BITVECTOR *n = (BITVECTOR *)malloc(sizeof(BITVECTOR));
    bitvector_init(n, size);
    bitvector_put_binstring(n, binstring);
    return n;
}

/* :187-195 */
void bitvector_print(BITVECTOR *bv)
{
    for (long long int k = 0; k < bv->cap; ++k) putchar('0' + ((bv->value[k >> 3] >> (7 - (k & 7))) & 1));
    putchar('\n');
}
