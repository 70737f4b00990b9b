/*
 * m1_stream.c -- byte-level MPEG-1 system / video headers (include/mpeg1_enc.h), written from the
 * byte formulas in SURVEY.md appendix C; each function cites the reference lines it reproduces.
 * Host side by design (north_star: "header assembly on the host").
 */
#include "mpeg1_enc.h"

#include <stdio.h>
#include <string.h>

static void put_start(uint8_t *o, uint8_t code) { o[0] = 0x00; o[1] = 0x00; o[2] = 0x01; o[3] = code; }

/* 22-bit mux rate between marker bits, big-endian 3 bytes (reference source/mpeg1_enc.c:14-20) */
static void put_rate(uint32_t rate, uint8_t *o)
{
    const uint32_t v = (((rate & 0x3fffffu) | 0x400000u) << 1) | 1u;
    o[0] = (uint8_t)(v >> 16); o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)v;
}

/* pack header, reference source/mpeg1_enc.c:7-22 */
void mpeg1_file_header(uint32_t multiplex_rate, uint8_t out[12])
{
    put_start(out, 0xba);
    out[4] = 0x21; out[5] = 0x00; out[6] = 0x01; out[7] = 0x00; out[8] = 0x01;
    put_rate(multiplex_rate, out + 9);
}

/* system header, :24-45 */
void mpeg1_sys_header(uint32_t multiplex_rate, uint8_t packet_num, uint8_t out[15])
{
    put_start(out, 0xbb);
    out[4] = 0x00; out[5] = 0x09;
    put_rate(multiplex_rate, out + 6);
    out[9] = 0x00; out[10] = 0x21; out[11] = 0xff; out[12] = 0xe0; out[13] = 0xe0;
    out[14] = packet_num;
}

static uint8_t *put_stamp(uint8_t *o, uint8_t lead, uint32_t v)
{
    *o++ = (uint8_t)(lead | ((v & 0xe0000000u) >> 28));
    *o++ = (uint8_t)((v & 0x1fe00000u) >> 21);
    *o++ = (uint8_t)(0x01 | ((v & 0x001fc000u) >> 13));
    *o++ = (uint8_t)((v & 0x00003fc0u) >> 6);
    *o++ = (uint8_t)(0x01 | ((v & 0x0000003fu) << 1));
    return o;
}

/* packet header, :47-77: 16 bytes with DTS/PTS when the argument is non-zero, 7 otherwise;
 * the two length bytes are left zero for the caller to patch (include/encoder.h:448-454) */
void mpeg1_packet_header(uint32_t t, uint8_t *out)
{
    put_start(out, 0xe0);
    out[4] = 0x00; out[5] = 0x00;
    if (t) {
        uint32_t dts = (uint32_t)((double)t * 1.2);          /* "dts_optinal *= 1.2" */
        dts += 0xbeefu;
        uint8_t *o = put_stamp(out + 6, 0x31, dts);
        put_stamp(o, 0x11, dts - 0xbeefu);
    } else {
        out[6] = 0x3f;
    }
}

/* sequence header, :81-94 */
void mpeg1_sequence_header(uint16_t width, uint16_t height, uint8_t aspect_ratio, uint8_t frame_rate,
                           uint8_t yby_size, uint8_t *out)
{
    put_start(out, 0xb3);
    out[4] = (uint8_t)((width & 0xff0) >> 4);
    out[5] = (uint8_t)(((width & 0xf) << 4) | ((height & 0xf00) >> 8));
    out[6] = (uint8_t)(height & 0xff);
    out[7] = (uint8_t)(((aspect_ratio & 0xf) << 4) | (frame_rate & 0xf));
    out[8] = 0xff; out[9] = 0xff; out[10] = 0xe0;
    out[11] = (uint8_t)((yby_size & 0x1f) << 3);
}

/* :96-98 */
void mpeg1_sequence_end(uint8_t out[4]) { put_start(out, 0xb7); }

/* group of pictures, :103-113 */
void mpeg1_gop(uint8_t drop_frame, uint8_t hour, uint8_t minute, uint8_t second, uint8_t num_pic,
               uint8_t closed, uint8_t broken, uint8_t *out)
{
    put_start(out, 0xb8);
    out[4] = (uint8_t)((drop_frame << 7) | ((hour & 0x1f) << 2) | ((minute & 0x30) >> 4));
    out[5] = (uint8_t)(((minute & 0xf) << 4) | 0x8 | ((second & 0x38) >> 3));
    out[6] = (uint8_t)(((second & 0x7) << 5) | ((num_pic & 0xfc) >> 1));
    out[7] = (uint8_t)(((num_pic & 1) << 7) | ((closed & 1) << 6) | ((broken & 1) << 5));
}

/* picture header, :120-137: 8 bytes for I pictures, a 9th for P/B */
void mpeg1_picture_header(uint16_t temporal_ref, uint8_t picture_type, uint16_t vbv_delay,
                          uint8_t *bidir_vector, uint8_t *out)
{
    put_start(out, 0x00);
    out[4] = (uint8_t)((temporal_ref & 0x3fc) >> 2);
    out[5] = (uint8_t)(((temporal_ref & 0x3) << 6) | ((picture_type & 0x7) << 3) | ((vbv_delay & 0xe000) >> 13));
    out[6] = (uint8_t)((vbv_delay & 0x1fe0) >> 5);
    out[7] = (uint8_t)((vbv_delay & 0x1f) << 3);
    if (picture_type == 2 || picture_type == 3) {
        out[7] |= (uint8_t)(((bidir_vector[0] & 1) << 2) | ((bidir_vector[1] & 6) >> 1));
        out[8] = (uint8_t)((bidir_vector[1] & 1) << 7);
        if (picture_type == 3) out[8] |= (uint8_t)(((bidir_vector[2] & 1) << 6) | ((bidir_vector[3] & 7) << 3));
    }
}

/* :139-143 */
void display_u8arr(uint8_t *buf, int32_t size)
{
    for (int32_t i = 0; i < size; ++i) printf("0x%02x ", buf[i]);
    printf("\n");
}
