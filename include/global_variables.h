/* global_variables.h -- as the reference's include/global_variables.h:4-7. */
#ifndef M1_COMPAT_GLOBAL_VARIABLES_H
#define M1_COMPAT_GLOBAL_VARIABLES_H
extern const int Q_MATRIX[8][8];       /* default intra matrix, reference source/image_processing.c:17-26 */
extern const int ZIGZAG_ORDER[8][8];   /* zigzag rank of (row, col), :28-37                                */
extern const char START_FILE;
extern const char START_PICTURE;
#endif
