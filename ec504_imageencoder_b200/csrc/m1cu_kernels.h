// m1cu_kernels.h -- host-callable launchers of m1cu_kernels.cu (internal to libm1cu.so).
#pragma once
#include "m1cu_common.cuh"

size_t      m1k_encode_smem_bytes(const M1Geom &g, int threads);
int         m1k_encode_threads(const M1Geom &g);
cudaError_t m1k_prepare(const M1Geom &g);
// (m1k_fill_tables / m1_make_quant: inline in m1cu_quant.h)

cudaError_t m1k_launch_encode(const M1Geom &g, const M1Quant &q, const uint8_t *rgb, int n_frames,
                              const M1Tables *tables, uint32_t *staging, uint32_t *chunk_bits,
                              short *levels, int *err, cudaStream_t st);
cudaError_t m1k_launch_layout(const M1Geom &g, int n_frames, const uint32_t *chunk_bits, uint32_t *chunk_dst,
                              uint32_t *frame_bytes, unsigned long long *frame_off,
                              unsigned long long *running, unsigned int *done_counter,
                              unsigned long long out_cap, int *err, cudaStream_t st);
cudaError_t m1k_launch_stitch(const M1Geom &g, int n_frames, int blocks_x, const uint32_t *staging,
                              const uint32_t *chunk_bits, const uint32_t *chunk_dst,
                              const uint32_t *frame_bytes, const unsigned long long *frame_off,
                              uint8_t *out, unsigned long long out_cap, cudaStream_t st);
cudaError_t m1k_launch_planes(const uint8_t *rgb, int channels, size_t npix, int n_frames, size_t frame_stride,
                              size_t plane_stride, uint8_t *Y, uint8_t *Cb, uint8_t *Cr, cudaStream_t st);
cudaError_t m1k_launch_synth(uint32_t seed, long first_frame, int n_frames, int W, int H, int kind,
                             uint8_t *rgb, cudaStream_t st);
cudaError_t m1k_launch_push(uint8_t *dst, const uint8_t *src, const unsigned long long *end, unsigned long long cap,
                            cudaStream_t st);
cudaError_t m1k_launch_stream(const uint8_t *payloads, const uint32_t *frame_bytes, const unsigned long long *frame_off,
                              int n_frames, long first_index, const uint8_t *prefix256, uint32_t trailer_be,
                              unsigned long long base, unsigned long long *seg_off, uint8_t *out, unsigned long long cap,
                              int *err, cudaStream_t st);
