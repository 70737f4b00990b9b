/*
 * mpeg1_enc.h -- byte-level system / sequence / GOP / picture headers, same prototypes as the
 * reference's include/mpeg1_enc.h:8-17.  Header assembly stays on the host (north_star).
 */
#ifndef M1_COMPAT_MPEG1_ENC_H
#define M1_COMPAT_MPEG1_ENC_H

#include <stdint.h>
#include <stdlib.h>
#include "bit_vector.h"
#include "mpeg1_blk.h"

#ifdef __cplusplus
extern "C" {
#endif
void mpeg1_file_header(uint32_t multiplex_rate, uint8_t out[12]);                     /* source/mpeg1_enc.c:7-22   */
void mpeg1_sys_header(uint32_t multiplex_rate, uint8_t packet_num, uint8_t out[15]);  /* :24-45                    */
void mpeg1_packet_header(uint32_t pts_optinal, uint8_t *out);                         /* :47-77, 16 bytes when != 0 */
void mpeg1_sequence_header(uint16_t width, uint16_t height, uint8_t aspect_ratio,
                           uint8_t frame_rate, uint8_t yby_size, uint8_t *out);       /* :81-94, 12 bytes          */
void mpeg1_sequence_end(uint8_t out[4]);                                              /* :96-98                    */
void mpeg1_gop(uint8_t drop_frame, uint8_t hour, uint8_t minute, uint8_t second, uint8_t num_pic,
               uint8_t closed, uint8_t broken, uint8_t *out);                         /* :103-113, 8 bytes         */
void mpeg1_picture_header(uint16_t temporal_ref, uint8_t picture_type, uint16_t vbv_delay,
                          uint8_t *bidir_vector, uint8_t *out);                       /* :120-137, 8 (I) or 9 bytes */
void display_u8arr(uint8_t *buf, int32_t size);                                       /* :139-143                  */
#ifdef __cplusplus
}
#endif
#endif
