// m1cu_encode_ws.cuh -- k_encode_ws: the persistent, warp-specialised form of the encode kernel.
// Included by m1cu_kernels.cu (uses its device functions: colour half-tiles, fdct8x8, code_block, ...).
//
// Why: measured on B200 (profiles/r1_phase_split.txt) the colour phase alone (FP64 + XU pipes) and
// the block phases alone (integer FMA + ALU pipes) each take about half of the fused kernel's time
// and do not overlap when every warp of a CTA walks through the phases together.  Here the roles
// run concurrently on every SM sub-partition, connected by mbarrier-guarded shared-memory rings:
//
//   colour warps 0-3 : RGB -> exact YCbCr -> 4:2:0, int32 samples into planes[stage]      (DP / XU)
//        | full[stage] / empty[stage]                       (2-stage ring of chunk planes)
//   block warps 4-6  : 8x8 DCT, non-zero mask, VLC into registers, scan, bits into win[w]  (IMAD / ALU)
//        | wfull[w] / wempty[w]                             (2 bit windows)
//   writer warp 7    : window -> chunk staging record, chunk bit count, re-zero the window (LSU)
//
// One chunk = up to 16 macroblocks of one slice = 256 colour half-tiles (2 per colour thread) and
// 96 blocks (1 per block thread).  CTAs are persistent: CTA b handles chunks b, b + grid, ...
#pragma once

#define M1_WS_CHUNK 16
#define M1_WS_STAGES 2
#define M1_WS_COLOUR_THREADS 128
#define M1_WS_BLOCK_THREADS 96
#define M1_WS_THREADS 256

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_addr(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, uint32_t parity)
{
    uint32_t ok = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_addr(b)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void block_bar() { asm volatile("bar.sync 1, 96;" ::: "memory"); }

struct WsChunk { int frame, slice, chunk, mb0, nmb; };
__device__ __forceinline__ WsChunk ws_decode(const M1Geom &g, int cid)
{
    WsChunk c;
    c.frame = cid / g.chunks_per_frame;
    const int rem = cid - c.frame * g.chunks_per_frame;
    c.slice = rem / g.chunks_per_slice;
    c.chunk = rem - c.slice * g.chunks_per_slice;
    c.mb0 = c.chunk * g.chunk_mbs;
    c.nmb = min(g.chunk_mbs, g.mbs_per_slice - c.mb0);
    return c;
}

template <int CH, bool kLevels>
__global__ void __launch_bounds__(M1_WS_THREADS, 3)
k_encode_ws(const __grid_constant__ M1Geom g, const uint8_t *__restrict__ rgb, const M1Tables *__restrict__ gtab,
            int n_chunks, uint32_t *__restrict__ staging, uint32_t *__restrict__ chunk_bits,
            short *__restrict__ levels, int *__restrict__ err)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    const int C = g.chunk_mbs;                                   // <= M1_WS_CHUNK
    const int WW = g.win_words;                                  // <= M1_WIN_WORDS (smaller only in tests)

    // shared memory carve-up
    int *planes0 = (int *)smem;                                                  // [STAGES][6*16 blocks][64]
    short *rec = (short *)(planes0 + M1_WS_STAGES * 6 * M1_WS_CHUNK * 64);       // [96][64] shorts, dense records
    uint32_t *win0 = (uint32_t *)(rec + M1_WS_BLOCK_THREADS * 64);              // [2][M1_WIN_WORDS + 2]
    M1Tables *tb = (M1Tables *)(win0 + 2 * (M1_WIN_WORDS + 2));
    int *wtot = (int *)(tb + 1);                                                 // [2][4] warp totals
    int *wtotal = wtot + 8;                                                      // [2] chunk bit totals for the writer
    unsigned long long *bars = (unsigned long long *)(((uintptr_t)(wtotal + 2) + 7) & ~(uintptr_t)7);
    unsigned long long *full = bars, *empty = bars + M1_WS_STAGES, *wfull = bars + 2 * M1_WS_STAGES, *wempty = wfull + 2;

    for (int i = tid; i < (int)(sizeof(M1Tables) / 4); i += M1_WS_THREADS) ((uint32_t *)tb)[i] = ((const uint32_t *)gtab)[i];
    for (int i = tid; i < 2 * (M1_WIN_WORDS + 2); i += M1_WS_THREADS) win0[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < M1_WS_STAGES; ++s) { mbar_init(&full[s], M1_WS_COLOUR_THREADS); mbar_init(&empty[s], M1_WS_BLOCK_THREADS); }
        for (int w = 0; w < 2; ++w) { mbar_init(&wfull[w], M1_WS_BLOCK_THREADS); mbar_init(&wempty[w], 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // roles rotate with the CTA's residency slot so the block warps of co-resident CTAs land on
    // different SM sub-partitions (warp w runs on sub-partition w % 4)
    const int rot = (blockIdx.x / 148) & 3;
    const int vw = ((tid >> 5) + 8 - rot) & 7;                   // virtual warp: 0-3 colour, 4-6 block, 7 writer
    const int lane = tid & 31;
    const size_t pitch = (size_t)g.W * g.channels;

    if (vw < 4) {
        // ================================ colour warps =========================================
        const int ct = vw * 32 + lane;
        int it = 0;
        for (int cid = blockIdx.x; cid < n_chunks; cid += gridDim.x, ++it) {
            const int s = it % M1_WS_STAGES;
            const uint32_t ph = (uint32_t)(it / M1_WS_STAGES) & 1u;
            const WsChunk c = ws_decode(g, cid);
            int *planes = planes0 + s * 6 * M1_WS_CHUNK * 64;
            const uint8_t *fr = rgb + (size_t)c.frame * g.frame_stride;
            mbar_wait(&empty[s], ph ^ 1u);                       // the block warps have drained this stage
            const int nbc = 2 * c.nmb;
#pragma unroll 1
            for (int ht = ct; ht < 8 * nbc; ht += M1_WS_COLOUR_THREADS) {
                const int qy = ht / nbc, bc = ht - qy * nbc;
                const int x0 = 16 * c.mb0 + 8 * bc, y0 = 16 * c.slice + 2 * qy;
                const int ry = min(y0, g.H - 1);                 // rows below the picture replicate its last row
                const size_t rp = (y0 + 1 <= g.H - 1) ? pitch : 0;
                if (x0 + 8 <= g.W)
                    color_half_tile<CH>(fr + (size_t)ry * pitch + (size_t)x0 * CH, rp, bc, qy, C, planes);
                else
                    color_half_tile_generic(fr, g, x0, y0, bc, qy, C, planes);
            }
            mbar_arrive(&full[s]);
        }
    } else if (vw < 7) {
        // ================================ block warps ==========================================
        const int bw = vw - 4, bt = bw * 32 + lane;              // 0..95, coding order: bt = 6*mb + blk
        const int mb = bt / 6, blk = bt - mb * 6;
        const bool is_luma = blk < 4;
        const int pb = is_luma ? (blk >> 1) * 2 * C + 2 * mb + (blk & 1) : blk * C + mb;
        int it = 0;
        for (int cid = blockIdx.x; cid < n_chunks; cid += gridDim.x, ++it) {
            const int s = it % M1_WS_STAGES;
            const uint32_t ph = (uint32_t)(it / M1_WS_STAGES) & 1u;
            const int w = it & 1;
            const uint32_t wph = (uint32_t)(it >> 1) & 1u;
            const WsChunk c = ws_decode(g, cid);
            const bool active = mb < c.nmb;
            const int *planes = planes0 + s * 6 * M1_WS_CHUNK * 64;
            uint32_t *win = win0 + w * (M1_WIN_WORDS + 2);

            int v[64];
            mbar_wait(&full[s], ph);
            if (active) {
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
            mbar_arrive(&empty[s]);                              // samples are in registers: release the stage

            unsigned long long nz = 0;
            BitAcc acc{0u, 0u, 0};
            if (active) {
                fdct8x8(v);
                uint32_t pk[32];
                nz = pack_and_flag(v, pk, *tb);
#pragma unroll
                for (int gI = 0; gI < 8; ++gI)
                    *(uint4 *)(rec + bt * 64 + (((gI ^ bt) & 7) << 3)) =
                        make_uint4(pk[4 * gI], pk[4 * gI + 1], pk[4 * gI + 2], pk[4 * gI + 3]);
                if (blk == 0) acc.put(3u, 2);                    // address increment '1' + macroblock_type '1'
                if (code_block<64>(acc, rec, bt, nz, is_luma, tb)) atomicOr(err, M1_ERRBIT_LEVEL);
                acc.finish();
            }

            // scan of the block lengths over the 96 block threads (coding order)
            const int my_bits = acc.n;
            int incl = my_bits;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (lane == 31) wtot[w * 4 + bw] = incl;
            block_bar();
            const int hdr_bits = c.chunk == 0 ? M1_SLICE_HDR_BITS : 0;
            const int t0 = wtot[w * 4], t1 = wtot[w * 4 + 1], t2 = wtot[w * 4 + 2];
            const int total_bits = hdr_bits + t0 + t1 + t2;
            const int my_off = hdr_bits + (bw > 0 ? t0 : 0) + (bw > 1 ? t1 : 0) + incl - my_bits;

            if (kLevels) {
                short *dst = levels + ((size_t)c.frame * g.mbs_per_frame + (size_t)c.slice * g.mbs_per_slice + c.mb0) * 384;
                for (int i = bt; i < c.nmb * 384; i += M1_WS_BLOCK_THREADS)
                    dst[i] = (short)quant_level(rec[rec_index<64>(i >> 6, i & 63)], i & 63, tb);
            }

            mbar_wait(&wempty[w], wph ^ 1u);                     // the writer has emptied and re-zeroed this window
            uint32_t *out = staging + (size_t)cid * (g.chunk_stride / 4);
            for (int w0 = 0;; w0 += 32 * WW) {
                if (bt == 0 && hdr_bits && w0 == 0) {
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
                        code_block<64>(ww, rec, bt, nz, is_luma, tb);
                    }
                }
                if (w0 + 32 * WW >= total_bits) break;
                // rare: the chunk needs more than one window; the block warps flush and clear it themselves
                block_bar();
                for (int i = bt; i < WW; i += M1_WS_BLOCK_THREADS) out[(w0 >> 5) + i] = win[i];
                block_bar();
                for (int i = bt; i < WW + 2; i += M1_WS_BLOCK_THREADS) win[i] = 0;
                block_bar();
            }
            if (bt == 0) wtotal[w] = total_bits;
            if (kLevels) block_bar();                            // rec is read cooperatively above: keep it stable
            mbar_arrive(&wfull[w]);                              // release: window + total visible to the writer
        }
    } else {
        // ================================ writer warp ==========================================
        int it = 0;
        for (int cid = blockIdx.x; cid < n_chunks; cid += gridDim.x, ++it) {
            const int w = it & 1;
            const uint32_t wph = (uint32_t)(it >> 1) & 1u;
            uint32_t *win = win0 + w * (M1_WIN_WORDS + 2);
            uint32_t *out = staging + (size_t)cid * (g.chunk_stride / 4);
            mbar_wait(&wfull[w], wph);
            const int total_bits = wtotal[w];
            const int w0 = ((total_bits - 1) / (32 * WW)) * (32 * WW);   // last window's first bit
            const int nwords = (total_bits - w0 + 31) >> 5;
            for (int i = lane; i < nwords; i += 32) { out[(w0 >> 5) + i] = win[i]; }
            for (int i = lane; i < nwords + 2 && i < M1_WIN_WORDS + 2; i += 32) win[i] = 0;
            if (lane == 0) chunk_bits[cid] = (uint32_t)total_bits;
            mbar_arrive(&wempty[w]);
        }
    }
}

typedef void (*ws_kernel_t)(const M1Geom, const uint8_t *, const M1Tables *, int, uint32_t *, uint32_t *, short *, int *);

// The warp-specialised kernel can serve FULL mode with aligned 3- or 4-byte pixels and chunks of at
// most 16 macroblocks.  It is parity-green but, as measured on B200, slower than
// k_encode_chunks: with 80 registers per block thread only 9 block warps fit per SM and the DCT code
// needs more resident warps than that to hide its dependent-issue latency, so the colour warps end up
// waiting on empty[].  It is therefore opt-in (M1_WS=1) and kept for the next round's tuning.
static ws_kernel_t ws_kernel_for(const M1Geom &g, bool levels)
{
    static const bool enabled = getenv("M1_WS") && atoi(getenv("M1_WS")) != 0;
    if (!enabled || g.mode != 0 || g.chunk_mbs > M1_WS_CHUNK || g.debug_skip) return nullptr;
    if (g.fast_load == 3) return levels ? k_encode_ws<3, true> : k_encode_ws<3, false>;
    if (g.fast_load == 4) return levels ? k_encode_ws<4, true> : k_encode_ws<4, false>;
    return nullptr;
}

static size_t m1k_ws_smem_bytes()
{
    return (size_t)M1_WS_STAGES * 6 * M1_WS_CHUNK * 256 + (size_t)M1_WS_BLOCK_THREADS * 128
           + 2 * (size_t)(M1_WIN_WORDS + 2) * 4 + sizeof(M1Tables) + 10 * sizeof(int) + 8 * sizeof(unsigned long long) + 32;
}
