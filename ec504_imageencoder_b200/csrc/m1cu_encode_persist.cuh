// m1cu_encode_persist.cuh -- k_encode_persist: the fused encode kernel as a PERSISTENT CTA with the
// next chunk's pixels prefetched by cp.async.  Included by m1cu_kernels.cu.
//
// Same phases and the same arithmetic as k_encode_chunks (colour -> planes, one thread per 8x8
// block, register bit accumulator, window), but
//   * a CTA loops over chunks cid = blockIdx.x, blockIdx.x + gridDim.x, ... (grid = 5 CTAs per SM),
//     so tables, window and barriers are set up once per CTA instead of once per chunk;
//   * while chunk i is in its block phases, the raw RGB bytes of chunk i+1 travel global -> shared
//     with cp.async (LDGSTS, no registers held), so the colour phase starts from shared memory and
//     the pixel loads' DRAM latency (the top stall of the colour phase, profiles/
//     r1_ncu_colour_phase_only.txt) is off the critical path.
// FULL mode, aligned 3- or 4-byte pixels only; everything else stays on k_encode_chunks.
#pragma once

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct PChunk { int frame, slice, chunk, mb0, nmb; };
__device__ __forceinline__ PChunk p_decode(const M1Geom &g, int cid)
{
    PChunk c;
    c.frame = cid / g.chunks_per_frame;
    const int rem = cid - c.frame * g.chunks_per_frame;
    c.slice = rem / g.chunks_per_slice;
    c.chunk = rem - c.slice * g.chunks_per_slice;
    c.mb0 = c.chunk * g.chunk_mbs;
    c.nmb = min(g.chunk_mbs, g.mbs_per_slice - c.mb0);
    return c;
}

// Stage layout: 8-byte units, unit u of half-tile slot h of thread t at stage[(h * 2*CH + u) * nthr + t]
// (u = dy * CH + i: row dy, i-th 8 bytes of the 8-pixel row) -- consecutive threads, consecutive
// addresses: conflict-free for both the cp.async writes and the 64-bit reads.
template <int CH>
__device__ __forceinline__ void prefetch_chunk(const M1Geom &g, const uint8_t *__restrict__ rgb, int cid, int tid, int nthr,
                                               uint2 *__restrict__ stage)
{
    const PChunk c = p_decode(g, cid);
    const uint8_t *fr = rgb + (size_t)c.frame * g.frame_stride;
    const size_t pitch = (size_t)g.W * CH;
    const int nbc = 2 * c.nmb;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ht = tid + h * nthr;
        if (ht < 8 * nbc) {
            const int qy = ht / nbc, bc = ht - qy * nbc;
            const int x0 = 16 * c.mb0 + 8 * bc, y0 = 16 * c.slice + 2 * qy;
            if (x0 + 8 <= g.W) {
                const int ry = min(y0, g.H - 1);                       // rows below the picture replicate its last row
                const size_t rp = (y0 + 1 <= g.H - 1) ? pitch : 0;
                const uint8_t *row0 = fr + (size_t)ry * pitch + (size_t)x0 * CH;
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                    for (int i = 0; i < CH; ++i)
                        cp_async8(stage + (size_t)(h * 2 * CH + dy * CH + i) * nthr + tid, row0 + dy * rp + 8 * i);
            }
        }
    }
    cp_async_commit();
}

template <int CH, bool kLevels>
__global__ void __launch_bounds__(128, 5)
k_encode_persist(const __grid_constant__ M1Geom g, const uint8_t *__restrict__ rgb, const M1Tables *__restrict__ gtab,
                 int n_chunks, uint32_t *__restrict__ staging, uint32_t *__restrict__ chunk_bits,
                 short *__restrict__ levels, int *__restrict__ err)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int C = g.chunk_mbs;
    const int WW = g.win_words;
    const int lane = tid & 31, warp = tid >> 5;

    // shared memory carve-up
    int *planes = (int *)smem;                                   // [6C blocks][64] int32, swizzled
    short *rec = (short *)smem;                                  // aliases planes (block-private reuse)
    uint2 *stage = (uint2 *)(planes + 6 * C * 64);               // [2 slots][2*CH units][nthr] raw pixels of the NEXT chunk
    uint32_t *win = (uint32_t *)(stage + (size_t)4 * CH * nthr); // [M1_WIN_WORDS + 2]
    M1Tables *tb = (M1Tables *)(win + M1_WIN_WORDS + 2);
    int *wtot = (int *)(tb + 1);                                 // [8] bits per warp

    for (int i = tid; i < (int)(sizeof(M1Tables) / 4); i += nthr) ((uint32_t *)tb)[i] = ((const uint32_t *)gtab)[i];
    for (int i = tid; i < M1_WIN_WORDS + 2; i += nthr) win[i] = 0;

    // block-thread identity (coding order t = 6*mb + blk) does not change from chunk to chunk
    const int mb = tid / 6, blk = tid - mb * 6;
    const bool is_luma = blk < 4;
    const int pb = is_luma ? (blk >> 1) * 2 * C + 2 * mb + (blk & 1) : blk * C + mb;

    if ((int)blockIdx.x < n_chunks) prefetch_chunk<CH>(g, rgb, blockIdx.x, tid, nthr, stage);
    __syncthreads();

    for (int cid = blockIdx.x; cid < n_chunks; cid += gridDim.x) {
        const PChunk c = p_decode(g, cid);
        const int nmb = c.nmb;
        const uint8_t *fr = rgb + (size_t)c.frame * g.frame_stride;

        // ---- phase 1: colour conversion, pixels come from the stage filled during the previous chunk
        cp_async_wait_all();                                     // each thread reads only what it copied itself
        {
            const int nbc = 2 * nmb;
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int ht = tid + h * nthr;
                if (ht >= 8 * nbc) break;
                const int qy = ht / nbc, bc = ht - qy * nbc;
                const int x0 = 16 * c.mb0 + 8 * bc, y0 = 16 * c.slice + 2 * qy;
                if (x0 + 8 <= g.W) {
                    HalfTilePixels<CH> px;
#pragma unroll
                    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            const uint2 u = stage[(size_t)(h * 2 * CH + dy * CH + i) * nthr + tid];
                            px.w[dy][2 * i] = u.x; px.w[dy][2 * i + 1] = u.y;
                        }
                    convert_half_tile<CH>(px, bc, qy, C, planes);
                } else {
                    color_half_tile_generic(fr, g, x0, y0, bc, qy, C, planes);
                }
            }
        }
        __syncthreads();                                         // planes complete, stage free

        // the next chunk's pixels start travelling now and have the whole block phase to arrive
        if (cid + (int)gridDim.x < n_chunks) prefetch_chunk<CH>(g, rgb, cid + gridDim.x, tid, nthr, stage);

        // ---- phase 2: one thread per 8x8 block: DCT, non-zero mask, coefficient record ---------------
        const bool active = mb < nmb;
        unsigned long long nz = 0;
        BitAcc acc{0u, 0u, 0};
        if (active) {
            int v[64];
            {
                const int key4 = blk_key(pb, C) << 2;
                const int *src = planes + pb * 64;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int o = ((i << 2) ^ key4);
                    const int4 a = *(const int4 *)(src + o);
                    const int4 b = *(const int4 *)(src + o + 32);
                    v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
                    v[32 + 4 * i] = b.x; v[32 + 4 * i + 1] = b.y; v[32 + 4 * i + 2] = b.z; v[32 + 4 * i + 3] = b.w;
                }
            }
            fdct8x8(v);
            uint32_t pk[32];
                nz = pack_and_flag(v, pk, *tb);
#pragma unroll
            for (int gI = 0; gI < 8; ++gI)
                *(uint4 *)(rec + pb * 128 + (((gI ^ pb) & 7) << 3)) =
                    make_uint4(pk[4 * gI], pk[4 * gI + 1], pk[4 * gI + 2], pk[4 * gI + 3]);
            // ---- phase 3: code the block into registers
            if (blk == 0) acc.put(3u, 2);
            if (code_block(acc, rec, pb, nz, is_luma, tb)) atomicOr(err, M1_ERRBIT_LEVEL);
            acc.finish();
        }

        // scan of the block lengths in thread (= coding) order
        const int my_bits = acc.n;
        int incl = my_bits;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();

        if (kLevels) {
            short *dst = levels + ((size_t)c.frame * g.mbs_per_frame + (size_t)c.slice * g.mbs_per_slice + c.mb0) * 384;
            for (int i = tid; i < nmb * 384; i += nthr) {
                const int p = i >> 6, z = i & 63, m = p / 6, b = p - m * 6;
                const int t = b < 4 ? (b >> 1) * 2 * C + 2 * m + (b & 1) : b * C + m;
                dst[i] = (short)quant_level(rec[rec_index(t, z)], z, tb);
            }
        }

        const int hdr_bits = c.chunk == 0 ? M1_SLICE_HDR_BITS : 0;
        int base = hdr_bits, total_bits = hdr_bits;
        {
            const int nw = nthr >> 5;
            for (int w = 0; w < nw; ++w) { const int t = wtot[w]; total_bits += t; if (w < warp) base += t; }
        }
        const int my_off = base + incl - my_bits;

        uint32_t *out = staging + (size_t)cid * (g.chunk_stride / 4);
        for (int w0 = 0;; w0 += 32 * WW) {                       // the window is all zero here
            if (tid == 0 && hdr_bits && w0 == 0) {
                // source/mpeg1_blk.c:12-20: 000001 | (vertical_pos+1)&0xff | quant_scale(5)=1 | 0
                WindowWriter ww{win, 0, 0, WW};
                ww.put(1u, 24);
                ww.put(((((uint32_t)(c.slice & 0xff) + 1u) & 0xffu) << 6) | (1u << 1), 14);
            }
            if (active && my_off < w0 + 32 * WW && my_off + my_bits > w0) {
                if (my_bits <= 64 && my_off >= w0 && my_off + my_bits <= w0 + 32 * WW) {
                    const int p = my_off - w0, word = p >> 5, o = p & 31;
                    const uint32_t a = acc.hi >> o;
                    const uint32_t b = __funnelshift_r(acc.lo, acc.hi, o);
                    const uint32_t cc = __funnelshift_r(0u, acc.lo, o);
                    if (a) atomicOr(&win[word], a);
                    if (b) atomicOr(&win[word + 1], b);
                    if (cc) atomicOr(&win[word + 2], cc);
                } else {
                    WindowWriter ww{win, my_off, w0, WW};        // long block, or one straddling the window
                    if (blk == 0) ww.put(3u, 2);
                    code_block(ww, rec, pb, nz, is_luma, tb);
                }
            }
            __syncthreads();                                     // bits complete; rec / planes no longer needed
            // copy out and clear in one sweep (the thread that copies a word clears it)
            const int nwords = min(WW, (total_bits - w0 + 31) >> 5);
            for (int i = tid; i < nwords + 2; i += nthr) {
                if (i < nwords) out[(w0 >> 5) + i] = win[i];
                win[i] = 0;
            }
            if (w0 + 32 * WW >= total_bits) break;
            __syncthreads();                                     // rare: the chunk needs another window pass
        }
        if (tid == 0) chunk_bits[cid] = (uint32_t)total_bits;
        // no barrier needed here: the next iteration writes planes (free since the barrier above) and
        // touches the window only after two more barriers
    }
}

typedef void (*persist_kernel_t)(const M1Geom, const uint8_t *, const M1Tables *, int, uint32_t *, uint32_t *, short *, int *);

// Measured on B200 (1080p, 300 frames): 149 k fps against 169 k for k_encode_chunks -- the prefetch
// removes the pixel-load stalls, but the per-chunk bookkeeping (two chunk decodes, 12 cp.async and
// 12 shared loads per thread) and 96 registers (5 CTAs per SM) cost more than they buy.  Parity-green
// (tests/test_gpu_variants.py), opt-in with M1_PERSIST=1.
static persist_kernel_t persist_kernel_for(const M1Geom &g, bool levels)
{
    static const bool enabled = getenv("M1_PERSIST") && atoi(getenv("M1_PERSIST")) != 0;
    if (!enabled || g.mode != 0 || g.debug_skip || g.chunk_mbs > 16) return nullptr;
    if (g.fast_load == 3) return levels ? k_encode_persist<3, true> : k_encode_persist<3, false>;
    if (g.fast_load == 4) return levels ? k_encode_persist<4, true> : k_encode_persist<4, false>;
    return nullptr;
}

static size_t m1k_persist_smem_bytes(const M1Geom &g, int threads)
{
    return (size_t)6 * g.chunk_mbs * 256 + (size_t)4 * g.fast_load * threads * 8 + (size_t)(M1_WIN_WORDS + 2) * 4
           + sizeof(M1Tables) + 8 * sizeof(int) + 16;
}
