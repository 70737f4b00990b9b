// m1cu_encode_groups.cuh -- k_encode_groups: the same chunk encoder as k_encode_chunks, restructured so that a
// chunk (<= 16 macroblocks of one slice) is the work of ONE WARP and nothing in the kernel waits for another warp.
//
// Why: k_encode_chunks gives a chunk to a CTA of four warps that meet at four barriers.  128 threads convert pixels
// but only 96 own a block, so a quarter of the resident warps idles through the block phase, the others wait for the
// slowest warp of their CTA, and the CTA holds its registers and shared memory until its last warp is done: about a
// third of the warp slots of an SM sit at a barrier at any time (ncu: 2.5 of 6.6 warps per sub-partition), and the
// kernel cannot cover its own latency.  Here a warp owns 16 macroblocks = 4096 pixels = 96
// blocks, which fills all 32 lanes in every phase:
//   round A  lane = one luma block of macroblocks 0..7 : colour conversion of ITS 8x8 pixels -> DCT -> code
//   round B  the same for macroblocks 8..15
//   round C  lane = one chroma block (Cb / Cr of the 16 macroblocks), gathered in shared memory by rounds A and B
//   then a warp scan over the 96 bit lengths in coding order, the bits OR-ed into the warp's window, copied out.
// Only __syncwarp between the phases; the four warps of a CTA share nothing but the coder's tables.
//
// Samples travel through shared memory as 16-bit values (128 bytes per block: a lane's luma staging doubles as its
// coefficient record after the DCT), 9 KB per warp, so that 24 independent warps fit an SM.
//
// Scope: FULL mode, 3- or 4-byte pixels, 16-pixel-aligned width and 16-byte-aligned input (the geometry of every
// BASELINE configuration); everything else, and any chunk with a block longer than 64 bits or more bits than the
// window holds (rare: 20 bits per block is typical), is encoded by k_encode_chunks / k_encode_redo.
#pragma once

#define M1G_WIN_WORDS 240                   // bit window per warp: 7680 bits (a chunk has 96 blocks); sized so that six CTAs fit an SM
#define M1G_WARP_WORDS (1024 + 1024 + M1G_WIN_WORDS + 8)   // luma staging + chroma blocks + window (+ slack), 32-bit words
#define M1G_WARPS 4

// 16-byte piece `c` (0..7) of 128-byte block `b`, swizzled so that eight neighbouring blocks hit eight different
// bank groups (128-byte stride would put them all on the same four banks)
__device__ __forceinline__ int g_piece(int b, int c) { return b * 32 + (((c ^ b) & 7) << 2); }

// One half-tile (8 pixels x 2 rows) of the lane's own luma block: rows 2*qy, 2*qy+1 of block `lb` (= lane) in the
// luma staging, and the four 2x2 chroma means into chroma blocks cb_blk / cb_blk + 1 at row crow, columns 4*h .. 4*h+3.
template <int CH>
__device__ __forceinline__ void g_convert_half_tile(const HalfTilePixels<CH> px, int lb, int qy, int cb_blk, int crow, int h,
                                                    uint32_t *__restrict__ lum, uint32_t *__restrict__ chr)
{
    const uint32_t (&w)[2][2 * CH] = px.w;
    int sb[4], sr[4];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        int yv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int byte0 = CH * i;
            int cb, cr;
            ycbcr_from_doubles(byte_to_double(w[dy][byte0 >> 2], byte0 & 3), byte_to_double(w[dy][(byte0 + 1) >> 2], (byte0 + 1) & 3),
                               byte_to_double(w[dy][(byte0 + 2) >> 2], (byte0 + 2) & 3), yv[i], cb, cr);
            if (dy == 0 && (i & 1) == 0) { sb[i >> 1] = cb; sr[i >> 1] = cr; }
            else                         { sb[i >> 1] += cb; sr[i >> 1] += cr; }
        }
        *(uint4 *)(lum + g_piece(lb, 2 * qy + dy)) =
            make_uint4(__byte_perm(yv[0], yv[1], 0x5410), __byte_perm(yv[2], yv[3], 0x5410),
                       __byte_perm(yv[4], yv[5], 0x5410), __byte_perm(yv[6], yv[7], 0x5410));
    }
    *(uint2 *)(chr + g_piece(cb_blk, crow) + 2 * h) =
        make_uint2(__byte_perm(sb[0] >> 2, sb[1] >> 2, 0x5410), __byte_perm(sb[2] >> 2, sb[3] >> 2, 0x5410));
    *(uint2 *)(chr + g_piece(cb_blk + 1, crow) + 2 * h) =
        make_uint2(__byte_perm(sr[0] >> 2, sr[1] >> 2, 0x5410), __byte_perm(sr[2] >> 2, sr[3] >> 2, 0x5410));
}

// The lane's block `b` of `area` (16-bit samples) -> DCT -> coefficient record in the same 128 bytes -> bits.
// Returns the register accumulator; *bad |= 1 when a coded level is outside the reference's range.
template <bool kLevels>
__device__ __forceinline__ BitAcc g_block(uint32_t *__restrict__ area, int b, bool is_luma, bool first_in_mb,
                                          const M1NzKeys &nk, const M1Tables *__restrict__ tb, int *bad,
                                          short *__restrict__ lev_dst)
{
    int v[64];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint4 a = *(const uint4 *)(area + g_piece(b, r));
        v[8 * r + 0] = (int)(a.x & 0xffffu); v[8 * r + 1] = (int)(a.x >> 16);
        v[8 * r + 2] = (int)(a.y & 0xffffu); v[8 * r + 3] = (int)(a.y >> 16);
        v[8 * r + 4] = (int)(a.z & 0xffffu); v[8 * r + 5] = (int)(a.z >> 16);
        v[8 * r + 6] = (int)(a.w & 0xffffu); v[8 * r + 7] = (int)(a.w >> 16);
    }
    fdct8x8(v);
    uint32_t pk[32];
    const unsigned long long nz = pack_and_flag(v, pk, nk);
    short *rec = (short *)area;
    const int key = b & 7;
#pragma unroll
    for (int gI = 0; gI < 8; ++gI)
        *(uint4 *)(rec + b * 64 + (((gI ^ key) & 7) << 3)) = make_uint4(pk[4 * gI], pk[4 * gI + 1], pk[4 * gI + 2], pk[4 * gI + 3]);
    if (kLevels) {
        // debug output: quantised zigzag levels of this block (the record is private to the lane: no barrier needed)
        for (int z = 0; z < 64; ++z) lev_dst[z] = (short)quant_level(rec[rec_index<64>(b, z, key)], z, tb);
    }
    BitAcc acc{0u, 0u, 0};
    if (first_in_mb) { acc.lo = 3u; acc.n = 2; }            // address increment '1' + macroblock_type '1'
    if (code_block<64>(acc, rec, b, nz, is_luma, tb, key)) *bad |= 1;
    acc.finish();
    return acc;
}

__device__ __forceinline__ void g_or_bits(uint32_t *__restrict__ win, int pos, const BitAcc &acc)
{
    if (acc.n == 0) return;
    const int word = pos >> 5, o = pos & 31;
    const uint32_t a = acc.hi >> o;
    const uint32_t b = __funnelshift_r(acc.lo, acc.hi, o);
    const uint32_t c = __funnelshift_r(0u, acc.lo, o);
    if (a) atomicOr(&win[word], a);
    if (b) atomicOr(&win[word + 1], b);
    if (c) atomicOr(&win[word + 2], c);
}

#ifndef M1G_MIN_CTAS
#define M1G_MIN_CTAS 6
#endif
template <int CH, bool kLevels>
__global__ void __launch_bounds__(32 * M1G_WARPS, M1G_MIN_CTAS)
k_encode_groups(const __grid_constant__ M1Geom g, const __grid_constant__ M1NzKeys nk,
                const uint8_t *__restrict__ rgb, const M1Tables *__restrict__ gtab, int n_groups,
                uint32_t *__restrict__ staging, uint32_t *__restrict__ chunk_bits,
                short *__restrict__ levels, int *__restrict__ err,
                unsigned int *__restrict__ redo_count, unsigned int *__restrict__ redo_list)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *wsm = (uint32_t *)smem + warp * M1G_WARP_WORDS;
    uint32_t *lum = wsm, *chr = wsm + 1024, *win = wsm + 2048;
    M1Tables *tb = (M1Tables *)((uint32_t *)smem + M1G_WARPS * M1G_WARP_WORDS);   // only the part up to ka[] is resident

    // the only CTA-wide step: the coder's tables (up to zofs[]; the non-zero keys come from the constant bank)
    constexpr int kTabVecs = ((int)offsetof(M1Tables, ka) + 15) / 16;
#pragma unroll 1
    for (int i = tid; i < kTabVecs; i += 32 * M1G_WARPS) ((uint4 *)tb)[i] = __ldg((const uint4 *)gtab + i);
#pragma unroll 1
    for (int i = lane; i < M1G_WIN_WORDS + 8; i += 32) win[i] = 0;
    __syncthreads();

    const int grp = blockIdx.x * M1G_WARPS + warp;          // = frame * chunks_per_frame + slice * chunks_per_slice + chunk
    if (grp >= n_groups) return;
    const int frame = grp / g.chunks_per_frame, rem = grp - frame * g.chunks_per_frame;
    const int slice = rem / g.chunks_per_slice, chunk = rem - slice * g.chunks_per_slice;
    const int mb0 = chunk * g.chunk_mbs;
    const int nmb = min(g.chunk_mbs, g.mbs_per_slice - mb0);
    const uint8_t *fr = rgb + (size_t)frame * g.frame_stride;
    const size_t pitch = (size_t)g.W * CH;
    const int last = g.H - 1;
    int bad = 0;

    // ---- rounds A and B: the lane's luma block of macroblocks 8*rd .. 8*rd+7 ---------------------------------
    // lanes 0..15 hold the upper block row (block columns 0..15 of the eight macroblocks), lanes 16..31 the lower
    const int by = lane >> 4, bc = lane & 15;
    const int mbl = bc >> 1;                                 // macroblock within the round
    const int blk = 2 * by + (bc & 1);                       // luma block 0..3 in coding order
    BitAcc accA{0u, 0u, 0}, accB{0u, 0u, 0};
#pragma unroll 1
    for (int rd = 0; rd < 2; ++rd) {
        const int mb = 8 * rd + mbl;
        if (mb < nmb) {
            const int x0 = 16 * (mb0 + mb) + 8 * (bc & 1);
            int y = 16 * slice + 8 * by;
            // rows below the picture replicate its last row (edge replication up to the coded size)
            const uint8_t *row = fr + (size_t)min(y, last) * pitch + (size_t)x0 * CH;
#pragma unroll 1
            for (int qy = 0; qy < 4; ++qy, y += 2) {
                const size_t rp = (y + 1 <= last) ? pitch : 0;
                g_convert_half_tile<CH>(load_half_tile<CH>(row, rp), lane, qy, 2 * mb, 4 * by + qy, bc & 1, lum, chr);
                row += rp + ((y + 2 <= last) ? pitch : 0);
            }
            short *lev = kLevels ? levels + (((size_t)frame * g.mbs_per_frame + (size_t)slice * g.mbs_per_slice + mb0 + mb) * 6 + blk) * 64
                                 : nullptr;
            const BitAcc a = g_block<kLevels>(lum, lane, true, blk == 0, nk, tb, &bad, lev);
            if (rd == 0) accA = a; else accB = a;            // (no runtime-indexed array: that would live in local memory)
        }
    }
    __syncwarp();                                            // the chroma blocks are complete

    // ---- round C: lane = chroma block 2*mb + c ------------------------------------------------------------------
    BitAcc accC{0u, 0u, 0};
    {
        const int mb = lane >> 1, c = lane & 1;
        if (mb < nmb) {
            short *lev = kLevels ? levels + (((size_t)frame * g.mbs_per_frame + (size_t)slice * g.mbs_per_slice + mb0 + mb) * 6 + 4 + c) * 64
                                 : nullptr;
            accC = g_block<kLevels>(chr, lane, false, false, nk, tb, &bad, lev);
        }
    }
    if (bad) atomicOr(err, M1_ERRBIT_LEVEL);

    // ---- bit offsets: exclusive scan over the 96 lengths in coding order (position 6*mb + block) ----------------
    // lum is free now (its records are coded): lens[96] lives there
    __syncwarp();
    int *lens = (int *)lum;
    const int pA = 6 * mbl + blk, pB = 6 * (8 + mbl) + blk, pC = 6 * (lane >> 1) + 4 + (lane & 1);
    lens[pA] = accA.n; lens[pB] = accB.n; lens[pC] = accC.n;
    __syncwarp();
    const int l0 = lens[3 * lane], l1 = lens[3 * lane + 1], l2 = lens[3 * lane + 2];
    const int mine = l0 + l1 + l2;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const int hdr_bits = chunk == 0 ? M1_SLICE_HDR_BITS : 0;
    const int total_bits = hdr_bits + __shfl_sync(0xffffffffu, incl, 31);
    const int base = hdr_bits + incl - mine;
    __syncwarp();
    lens[3 * lane] = base; lens[3 * lane + 1] = base + l0; lens[3 * lane + 2] = base + l0 + l1;
    __syncwarp();
    const int oA = lens[pA], oB = lens[pB], oC = lens[pC];

    const size_t gi = (size_t)grp;
    const bool too_long = max(max(accA.n, accB.n), accC.n) > 64;
    if (__any_sync(0xffffffffu, too_long) || total_bits > 32 * M1G_WIN_WORDS) {
        // hand the chunk to k_encode_redo (CTA-per-chunk path: re-codes long blocks straight into its window)
        if (lane == 0) {
            chunk_bits[gi] = 0u;
            redo_list[atomicAdd(redo_count, 1u)] = (unsigned int)grp;
        }
        return;
    }
    if (lane == 0 && hdr_bits) {
        // source/mpeg1_blk.c:12-20: 000001 | (vertical_pos+1)&0xff | quant_scale(5)=1 | 0   (38 bits)
        win[0] = 1u << 8 | ((((uint32_t)(slice & 0xff) + 1u) & 0xffu));          // 24 bits 000001, 8 bits position
        win[1] = 1u << 27;                                                        // 00001 0, then the blocks
    }
    __syncwarp();
    g_or_bits(win, oA, accA);
    g_or_bits(win, oB, accB);
    g_or_bits(win, oC, accC);
    __syncwarp();
    uint32_t *out = staging + gi * (g.chunk_stride / 4);
    const int nwords = (total_bits + 31) >> 5;
#pragma unroll 1
    for (int i = lane; i < nwords; i += 32) out[i] = win[i];
    if (lane == 0) chunk_bits[gi] = (uint32_t)total_bits;
}
