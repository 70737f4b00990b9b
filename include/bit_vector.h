/*
 * bit_vector.h -- the reference's growable MSB-first bit buffer, ABI-compatible.
 * Replaces reference include/bit_vector.h:9-42 (struct layout and every prototype kept, so code
 * compiled against the reference header links against this library unchanged).
 * Bit k of the stream is bit 7-(k%8) of byte k/8 (reference source/bit_vector.c:20,:47).
 * Host-side compatibility layer: the accelerated path packs bits on the GPU (csrc/m1cu_kernels.cu).
 */
#ifndef M1_COMPAT_BIT_VECTOR_H
#define M1_COMPAT_BIT_VECTOR_H

#include <stdio.h>
#include <stdlib.h>

#define BITVECTOR struct bitvector

struct bitvector {
    char *value;            /* storage, (bits >> 3) + 1 bytes            */
    long long int bits;     /* capacity in bits                          */
    long long int cursor;   /* next write position                       */
    long long int cap;      /* number of valid bits (high-water mark)    */
};

#ifdef __cplusplus
extern "C" {
#endif

void bitvector_init(BITVECTOR *bv, long long int size);
BITVECTOR *bitvector_new(const char *binstring, long long int size);   /* size = capacity; length = strlen */
BITVECTOR *bitvector_clone(BITVECTOR *bv);
void bitvector_expand_size(BITVECTOR *bv, long long int speculative);

void bitvector_put_bit(BITVECTOR *bv, char bit);
void bitvector_put_binstring(BITVECTOR *bv, const char *bitstring);
void bitvector_put_byte_off(BITVECTOR *bv, unsigned char val, char bits, char offset);
void bitvector_put_byte(BITVECTOR *bv, char val, char bits);
void bitvector_put_byte_ent(BITVECTOR *bv, char val);
void bitvector_concat(BITVECTOR *dest, BITVECTOR *src);
long long int bitvector_pos(BITVECTOR *bv, long long int off);

int bitvector_toarray(BITVECTOR *bv, char *output);
int bitvector_fwrite(BITVECTOR *bv, FILE *file);
void bitvector_print(BITVECTOR *bv);

#ifdef __cplusplus
}
#endif
#endif
