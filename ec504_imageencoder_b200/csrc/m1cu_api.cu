// m1cu_api.cu -- the C ABI of include/m1cu.h: context, sizing, stream-ordered launches and the
// host-buffer convenience path.  No CPU fallback lives here: every compute entry point needs a
// CUDA device and reports M1CU_ERR_CUDA otherwise.
#include "../../include/m1cu.h"
#include "m1cu_common.cuh"
#include "m1cu_kernels.h"
#include "m1cu_quant.h"

#ifdef M1_EXPERIMENTS
#include "../../tools/experiments/m1x_env.h"
#endif

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

char g_last_error[256] = "";

// Default intra matrix (reference source/image_processing.c:17-26), raster order.
const int kIntra[64] = {
     8, 16, 19, 22, 26, 27, 29, 34,   16, 16, 22, 24, 27, 29, 34, 37,
    19, 22, 26, 27, 29, 34, 34, 38,   22, 22, 26, 27, 29, 34, 37, 40,
    22, 26, 27, 29, 32, 35, 40, 48,   26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69,   27, 29, 35, 38, 46, 56, 69, 83 };

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

struct m1cu_ctx {
    int device = 0;
    int quality = 0;
    int max_frames = 0;
    int batch_frames = 0;           // pictures per launch round (bounds the staging memory)
    M1Geom g{};
    M1Quant q{};
    int32_t qm[64]{};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // device state
    uint32_t *d_staging = nullptr, *d_chunk_bits = nullptr, *d_chunk_dst = nullptr;
    // second set + side stream: calls longer than batch_frames run the layout + stitch of one launch round
    // beside the chunk encoder of the next (allocated by the first such call)
    uint32_t *d_staging2 = nullptr, *d_chunk_bits2 = nullptr, *d_chunk_dst2 = nullptr;
    cudaStream_t tail_stream = nullptr;
    cudaEvent_t enc_done[2] = { nullptr, nullptr }, tail_done[2] = { nullptr, nullptr };
    M1Tables *d_tables = nullptr;
    // m1cu_assemble_stream: prefix templates (256 x 44) + prologue (27) on the device, the host copy they came from
    uint8_t *d_stream_tmpl = nullptr;
    unsigned long long *d_seg_off = nullptr; int seg_frames = 0;
    uint8_t *d_stream = nullptr; size_t d_stream_cap = 0;     // m1cu_encode_host_stream
    unsigned long long *d_stream_end = nullptr;
    std::vector<uint8_t> h_stream_tmpl;
    int *d_err = nullptr;
    unsigned long long *d_running = nullptr;
    unsigned int *d_done = nullptr;
    // host-path scratch
    uint8_t *d_in = nullptr;   size_t d_in_cap = 0;  int d_in_frames = 0;   // pictures the last host-buffer call uploaded
    uint8_t *d_planes = nullptr; size_t d_planes_cap = 0;                      // m1cu_host_batch_planes
    uint8_t *d_out = nullptr;  size_t d_out_cap = 0;
    uint32_t *d_fbytes = nullptr; unsigned long long *d_foff = nullptr; int d_meta_frames = 0;
    int16_t *d_levels = nullptr; size_t d_levels_cap = 0;
    uint8_t *h_out = nullptr;  size_t h_out_cap = 0;       // pinned bounce buffer
    uint32_t *h_fbytes = nullptr; unsigned long long *h_foff = nullptr; int h_meta_frames = 0;
    // pipelined host path (encode_host_pipelined)
    cudaStream_t copy_stream = nullptr, back_stream = nullptr;
    std::vector<cudaEvent_t> pipe_events;
    uint32_t *p_fbytes = nullptr, *ph_fbytes = nullptr;
    unsigned long long *p_foff = nullptr, *ph_foff = nullptr;
    int pipe_sub = 0, pipe_frames = 0;
    unsigned long long launches = 0;
    // optional per-kernel timing (m1cu_enable_timing)
    bool timing = false;
    struct Span { cudaEvent_t a, b; int kind; };
    std::vector<Span> spans;          // recorded, not yet read
    std::vector<cudaEvent_t> pool;    // recycled events
    char err[256] = "";
};

namespace {

int fail(m1cu_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    char buf[256];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    if (c) memcpy(c->err, buf, sizeof buf);
    memcpy(g_last_error, buf, sizeof buf);
    return code;
}

#define CU(call)                                                              \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return fail(ctx, M1CU_ERR_CUDA, #call, e_);    \
    } while (0)

cudaEvent_t take_event(m1cu_ctx *ctx)
{
    cudaEvent_t e = nullptr;
    if (!ctx->pool.empty()) { e = ctx->pool.back(); ctx->pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

// brackets one launch with events when timing is on
struct Timed {
    m1cu_ctx *c; int kind; cudaStream_t s; cudaEvent_t a = nullptr;
    Timed(m1cu_ctx *ctx, int k, cudaStream_t st = nullptr) : c(ctx), kind(k), s(st ? st : ctx->stream)
    { if (c->timing) { a = take_event(c); cudaEventRecord(a, s); } }
    ~Timed() { if (a) { cudaEvent_t b = take_event(c); cudaEventRecord(b, s); c->spans.push_back({a, b, kind}); } }
};

int ensure(m1cu_ctx *ctx, void **p, size_t *cap, size_t need, bool pinned = false)
{
    if (*cap >= need && *p) return M1CU_OK;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cap = 0; }
    const size_t want = align_up(need + need / 8, 256);
    CU(pinned ? cudaMallocHost(p, want) : cudaMalloc(p, want));
    *cap = want;
    return M1CU_OK;
}

}  // namespace

extern "C" {

int m1cu_abi_version(void) { return M1CU_ABI_VERSION; }

int m1cu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *m1cu_last_error(const m1cu_ctx *ctx) { return ctx ? ctx->err : g_last_error; }

// scale_quantization_matrix, reference source/image_processing.c:314-343: float scale factor,
// int*float product in float, /100.0 in double, round half away, floor of 1.
int m1cu_qmatrix(int quality, int32_t out[64])
{
    if (!out) return M1CU_ERR_ARG;
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    const float sf = quality < 50 ? (float)(5000.0 / quality) : (float)(200.0 - 2 * quality);
    for (int k = 0; k < 64; ++k) {
        const float prod = (float)kIntra[k] * sf;
        const int v = (int)round((double)prod / 100.0);
        out[k] = v < 1 ? 1 : v;
    }
    return M1CU_OK;
}

int m1cu_create(m1cu_ctx **out, int device, int width, int height, int channels, int mode,
                int quality, int max_frames)
{
    return m1cu_create_ex(out, device, width, height, channels, mode, quality, max_frames, nullptr);
}

int m1cu_create_ex(m1cu_ctx **out, int device, int width, int height, int channels, int mode,
                   int quality, int max_frames, const m1cu_tuning *tuning)
{
    m1cu_ctx *ctx = nullptr;
    if (!out) return fail(nullptr, M1CU_ERR_ARG, "m1cu_create: out is NULL");
    *out = nullptr;
    if (width <= 0 || height <= 0 || channels < 3 || max_frames <= 0 ||
        (mode != M1CU_MODE_FULL && mode != M1CU_MODE_REF_COMPAT) || width > 65535 || height > 65535)
        return fail(nullptr, M1CU_ERR_ARG, "m1cu_create: bad geometry");
    if (mode == M1CU_MODE_REF_COMPAT) {
        // the literal loops read columns 0..95, rows 0..143 and chroma offsets up to 71*(W/2)+47
        if (width < 96 || height < 144 || (size_t)71 * (width / 2) + 47 >= (size_t)width * height)
            return fail(nullptr, M1CU_ERR_ARG, "m1cu_create: REF_COMPAT needs at least 96x144");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(nullptr, M1CU_ERR_CUDA, "no CUDA device (this library has no CPU fallback)"); }
    if (device < 0 || device >= ndev) return fail(nullptr, M1CU_ERR_ARG, "m1cu_create: bad device index");

    ctx = new m1cu_ctx();
    ctx->device = device; ctx->quality = quality; ctx->max_frames = max_frames;
    M1Geom &g = ctx->g;
    g.W = width; g.H = height; g.channels = channels; g.mode = mode;
    if (mode == M1CU_MODE_FULL) { g.slices = (height + 15) / 16; g.mbs_per_slice = (width + 15) / 16; }
    else                        { g.slices = 6; g.mbs_per_slice = 9; }
    int max_chunk = M1_DEFAULT_CHUNK_MBS;
    if (tuning && tuning->chunk_mbs) {
        if (tuning->chunk_mbs < 1 || tuning->chunk_mbs > M1_MAX_CHUNK_MBS) {
            delete ctx;
            return fail(nullptr, M1CU_ERR_ARG, "m1cu_create_ex: chunk_mbs out of range");
        }
        max_chunk = tuning->chunk_mbs;
    }
    // Full chunks plus one shorter tail per slice (1080p: 7 x 16 + 8 macroblocks): a full chunk fills its
    // three block warps and four colour warps completely, which beats equal chunks of 15 by 1.7 %.
    // tuning->chunk_even restores the equal split.
    g.chunk_mbs = max_chunk < g.mbs_per_slice ? max_chunk : g.mbs_per_slice;
    if (tuning && tuning->chunk_even) {
        g.chunks_per_slice = (g.mbs_per_slice + max_chunk - 1) / max_chunk;
        g.chunk_mbs = (g.mbs_per_slice + g.chunks_per_slice - 1) / g.chunks_per_slice;
    }
    g.chunks_per_slice = (g.mbs_per_slice + g.chunk_mbs - 1) / g.chunk_mbs;
    g.chunks_per_frame = g.chunks_per_slice * g.slices;
    g.mbs_per_frame = g.mbs_per_slice * g.slices;
    {
        const int last_mbs = g.mbs_per_slice - (g.chunks_per_slice - 1) * g.chunk_mbs;
        g.inv_nbc[0] = (65536u + 2u * g.chunk_mbs - 1u) / (2u * g.chunk_mbs);
        g.inv_nbc[1] = (65536u + 2u * last_mbs - 1u) / (2u * last_mbs);
        // A short last chunk (1080p: 120 = 7 x 16 + 8) would hold a whole CTA slot for half a CTA's work: when two
        // of them fit one chunk, the last chunks of slices 2k and 2k+1 share a CTA (k_encode_chunks).
        g.pair_tails = (mode == M1CU_MODE_FULL && g.chunks_per_slice >= 2 && g.slices >= 2 && 2 * last_mbs <= g.chunk_mbs &&
                        !(tuning && tuning->no_tail_pairing)) ? last_mbs : 0;
        g.inv_nbc[2] = (65536u + 4u * last_mbs - 1u) / (4u * last_mbs);
    }
    g.chunk_stride = (unsigned)align_up(((size_t)M1_SLICE_HDR_BITS + (size_t)g.chunk_mbs * M1_MB_MAX_BITS + 7) / 8 + 8, 16);
    g.frame_stride = (unsigned long long)width * height * channels;
    // 128-bit tile loads need 16-pixel tiles to start on 16-byte boundaries in every row and picture
    g.fast_load = (mode == M1CU_MODE_FULL && (channels == 3 || channels == 4) && width % 16 == 0) ? channels : 0;
#ifdef M1_EXPERIMENTS
    g.debug_skip = m1x_env_int("M1_DEBUG_SKIP");           // tools/ profiling build only
#else
    g.debug_skip = 0;
#endif
    g.win_words = M1_WIN_WORDS;
    if (tuning && tuning->win_words) {                     // tests: force the multi-window path
        if (tuning->win_words < 4 || tuning->win_words > M1_WIN_WORDS) {
            delete ctx;
            return fail(nullptr, M1CU_ERR_ARG, "m1cu_create_ex: win_words out of range");
        }
        g.win_words = tuning->win_words;
    }

    m1cu_qmatrix(quality, ctx->qm);
    if (!m1_make_quant(ctx->qm, &ctx->q)) { delete ctx; return fail(nullptr, M1CU_ERR_ARG, "m1cu_create: quantiser constants failed self-check"); }
    // blocks whose samples span at most flat_range have no non-zero AC level at this quality: they skip DCT + test
    // (a range below 6 grey levels -- quality 40 and up -- is not worth the test: sensor noise alone spans more, and the test
    // plus the 64-register build cost 3 % where nothing qualifies, profiles/r2_flat_skip_ab.txt)
    g.flat_range = (tuning && tuning->no_flat_skip) ? -1 : m1_flat_range(ctx->qm);
    if (g.flat_range < 6) g.flat_range = -1;

#define CUC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { int rc_ = fail(ctx, M1CU_ERR_CUDA, #call, e_); memcpy(g_last_error, ctx->err, sizeof g_last_error); m1cu_destroy(ctx); return rc_; } } while (0)
    CUC(cudaSetDevice(device));
    CUC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
    // bound the staging memory: at most ~2 GiB of (worst-case sized, sparsely written) chunk records
    // per launch round; 300 frames of 1080p fit in one round
    const size_t per_frame = (size_t)g.chunks_per_frame * g.chunk_stride;
    size_t batch = ((size_t)2 << 30) / per_frame;
    if (batch < 1) batch = 1;
    if (batch > (size_t)max_frames) batch = (size_t)max_frames;
    if (batch > 65535) batch = 65535;                       // grid.z of k_encode_chunks / grid.y of k_stitch
    if (tuning && tuning->batch_frames > 0 && (size_t)tuning->batch_frames < batch) batch = (size_t)tuning->batch_frames;
    ctx->batch_frames = (int)batch;
    CUC(cudaMalloc(&ctx->d_staging, per_frame * batch));
    CUC(cudaMalloc(&ctx->d_chunk_bits, sizeof(uint32_t) * g.chunks_per_frame * batch));
    CUC(cudaMalloc(&ctx->d_chunk_dst, sizeof(uint32_t) * g.chunks_per_frame * batch));
    CUC(cudaMalloc(&ctx->d_tables, sizeof(M1Tables)));
    CUC(cudaMalloc(&ctx->d_err, sizeof(int)));
    CUC(cudaMalloc(&ctx->d_running, sizeof(unsigned long long)));
    CUC(cudaMalloc(&ctx->d_done, sizeof(unsigned int)));
    // Everything below is ordered on the context's own (non-blocking) stream and waited for here: the
    // legacy default stream does not order against it.
    CUC(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    CUC(cudaMemsetAsync(ctx->d_done, 0, sizeof(unsigned int), ctx->stream));
    CUC(cudaMemsetAsync(ctx->d_running, 0, sizeof(unsigned long long), ctx->stream));
    M1Tables ht;
    m1k_fill_tables(&ht, ctx->q);
    CUC(cudaMemcpyAsync(ctx->d_tables, &ht, sizeof ht, cudaMemcpyHostToDevice, ctx->stream));
    CUC(cudaStreamSynchronize(ctx->stream));                // `ht` is a stack object
    CUC(m1k_prepare(g));
#undef CUC
    *out = ctx;
    return M1CU_OK;
}

int m1cu_destroy(m1cu_ctx *ctx)
{
    if (!ctx) return M1CU_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->tail_stream) cudaStreamSynchronize(ctx->tail_stream);
    cudaFree(ctx->d_staging); cudaFree(ctx->d_chunk_bits); cudaFree(ctx->d_chunk_dst);
    cudaFree(ctx->d_staging2); cudaFree(ctx->d_chunk_bits2); cudaFree(ctx->d_chunk_dst2);
    for (int i = 0; i < 2; ++i) { if (ctx->enc_done[i]) cudaEventDestroy(ctx->enc_done[i]); if (ctx->tail_done[i]) cudaEventDestroy(ctx->tail_done[i]); }
    if (ctx->tail_stream) cudaStreamDestroy(ctx->tail_stream);
    cudaFree(ctx->d_tables); cudaFree(ctx->d_stream_tmpl); cudaFree(ctx->d_seg_off); cudaFree(ctx->d_stream); cudaFree(ctx->d_stream_end); cudaFree(ctx->d_err); cudaFree(ctx->d_running); cudaFree(ctx->d_done);
    cudaFree(ctx->d_in); cudaFree(ctx->d_out); cudaFree(ctx->d_fbytes); cudaFree(ctx->d_foff);
    cudaFree(ctx->d_levels); cudaFree(ctx->d_planes);
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->h_fbytes) cudaFreeHost(ctx->h_fbytes);
    if (ctx->h_foff) cudaFreeHost(ctx->h_foff);
    cudaFree(ctx->p_fbytes); cudaFree(ctx->p_foff);
    if (ctx->ph_fbytes) cudaFreeHost(ctx->ph_fbytes);
    if (ctx->ph_foff) cudaFreeHost(ctx->ph_foff);
    for (auto e : ctx->pipe_events) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->back_stream) cudaStreamDestroy(ctx->back_stream);
    for (auto &s : ctx->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : ctx->pool) cudaEventDestroy(e);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return M1CU_OK;
}

int m1cu_set_stream(m1cu_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return M1CU_ERR_ARG;
    if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    ctx->own_stream = false;
    ctx->stream = (cudaStream_t)cuda_stream;
    if (!cuda_stream) {
        CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return M1CU_OK;
}

int m1cu_synchronize(m1cu_ctx *ctx)
{
    if (!ctx) return M1CU_ERR_ARG;
    CU(cudaStreamSynchronize(ctx->stream));
    return M1CU_OK;
}

int m1cu_macroblocks_per_frame(const m1cu_ctx *ctx) { return ctx ? ctx->g.mbs_per_frame : 0; }
int m1cu_flat_range(const m1cu_ctx *ctx) { return ctx ? ctx->g.flat_range : -1; }
size_t m1cu_frame_bytes_in(const m1cu_ctx *ctx) { return ctx ? (size_t)ctx->g.frame_stride : 0; }

size_t m1cu_payload_bound(const m1cu_ctx *ctx)
{
    if (!ctx) return 0;
    const M1Geom &g = ctx->g;
    const size_t slice_bits = M1_SLICE_HDR_BITS + (size_t)g.mbs_per_slice * M1_MB_MAX_BITS + 7;
    return align_up((slice_bits / 8) * g.slices, 16);
}

size_t m1cu_typical_out_bytes(const m1cu_ctx *ctx, int n_frames)
{
    if (!ctx || n_frames <= 0) return 0;
    // a quarter of the coded 4:2:0 sample count per picture: > 10x what natural content needs
    const size_t per = align_up((size_t)ctx->g.mbs_per_frame * 96 + 256, 16);
    const size_t bound = m1cu_payload_bound(ctx);
    return (per < bound ? per : bound) * (size_t)n_frames;
}

int m1cu_encode_device(m1cu_ctx *ctx, const uint8_t *d_rgb, int n_frames, uint8_t *d_out, size_t out_cap,
                       uint32_t *d_frame_bytes, uint64_t *d_frame_offsets, int16_t *d_levels)
{
    if (!ctx || !d_rgb || !d_out || !d_frame_bytes || !d_frame_offsets || n_frames <= 0)
        return fail(ctx, M1CU_ERR_ARG, "m1cu_encode_device: bad argument");
    if (n_frames > ctx->max_frames) return fail(ctx, M1CU_ERR_ARG, "m1cu_encode_device: n_frames > max_frames");
    if (((uintptr_t)d_out & 15) != 0) return fail(ctx, M1CU_ERR_ARG, "m1cu_encode_device: d_out must be 16-byte aligned");
    M1Geom g = ctx->g;
    if (((uintptr_t)d_rgb & 15) != 0) g.fast_load = 0;          // unaligned input: generic loads
    cudaStream_t st = ctx->stream;
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(ctx->d_running, 0, sizeof(unsigned long long), st));
    int bx = (g.chunks_per_frame + 7) / 8;                     // k_stitch: one warp per chunk, 8 warps per CTA
    if (bx > 512) bx = 512;
    {   // many pictures per launch: cap the grid at 32 CTAs per SM and let every warp loop over chunks (20 400 CTAs of
        // one chunk per warp for 300 pictures of 1080p are bound by CTA launches: 51.5 -> 47.2 us, profiles/r2_stitch_grid.txt)
        const int nb0 = n_frames < ctx->batch_frames ? n_frames : ctx->batch_frames;
        const int want = (148 * 32 + nb0 - 1) / nb0;
        if (want < bx) bx = want;
    }
    const int rounds = (n_frames + ctx->batch_frames - 1) / ctx->batch_frames;
    // More than one launch round (the staging memory is bounded): the layout + stitch of round r run on a side
    // stream beside the chunk encoder of round r + 1, with two sets of staging / chunk arrays.  Everything is
    // still ordered on the context's stream for the caller: it waits for the last tail before the call returns
    // control of the stream.
    const bool overlap = rounds > 1;
    if (overlap && !ctx->d_staging2) {
        const size_t per_frame = (size_t)g.chunks_per_frame * g.chunk_stride;
        CU(cudaMalloc(&ctx->d_staging2, per_frame * (size_t)ctx->batch_frames));
        CU(cudaMalloc(&ctx->d_chunk_bits2, sizeof(uint32_t) * g.chunks_per_frame * (size_t)ctx->batch_frames));
        CU(cudaMalloc(&ctx->d_chunk_dst2, sizeof(uint32_t) * g.chunks_per_frame * (size_t)ctx->batch_frames));
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&ctx->tail_stream, cudaStreamNonBlocking, hi));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&ctx->enc_done[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->tail_done[i], cudaEventDisableTiming));
        }
    }
    int r = 0;
    for (int f0 = 0; f0 < n_frames; f0 += ctx->batch_frames, ++r) {
        const int nb = n_frames - f0 < ctx->batch_frames ? n_frames - f0 : ctx->batch_frames;
        const int b = overlap ? (r & 1) : 0;
        uint32_t *staging = b ? ctx->d_staging2 : ctx->d_staging;
        uint32_t *cbits = b ? ctx->d_chunk_bits2 : ctx->d_chunk_bits;
        uint32_t *cdst = b ? ctx->d_chunk_dst2 : ctx->d_chunk_dst;
        cudaStream_t tail = overlap ? ctx->tail_stream : st;
        if (overlap && r >= 2) CU(cudaStreamWaitEvent(st, ctx->tail_done[b], 0));   // set b has been stitched
        {
            Timed t(ctx, 0);
            CU(m1k_launch_encode(g, ctx->q, d_rgb + (size_t)f0 * g.frame_stride, nb, ctx->d_tables, staging,
                                 cbits, d_levels ? d_levels + (size_t)f0 * g.mbs_per_frame * 384 : nullptr,
                                 ctx->d_err, st));
        }
        if (overlap) {
            CU(cudaEventRecord(ctx->enc_done[b], st));
            CU(cudaStreamWaitEvent(tail, ctx->enc_done[b], 0));
        }
        {
            Timed t(ctx, 1, tail);
            CU(m1k_launch_layout(g, nb, cbits, cdst, d_frame_bytes + f0,
                                 (unsigned long long *)d_frame_offsets + f0, ctx->d_running, ctx->d_done,
                                 (unsigned long long)out_cap, ctx->d_err, tail));
        }
        {
            Timed t(ctx, 2, tail);
            CU(m1k_launch_stitch(g, nb, bx, staging, cbits, cdst, d_frame_bytes + f0,
                                 (const unsigned long long *)d_frame_offsets + f0, d_out, (unsigned long long)out_cap, tail));
        }
        if (overlap) CU(cudaEventRecord(ctx->tail_done[b], tail));
        ctx->launches += 3;
    }
    if (overlap) CU(cudaStreamWaitEvent(st, ctx->tail_done[(rounds - 1) & 1], 0));     // the side stream is in order: this covers every round
    return M1CU_OK;
}

int m1cu_check(m1cu_ctx *ctx)
{
    if (!ctx) return M1CU_ERR_ARG;
    int flags = 0;
    CU(cudaMemcpyAsync(&flags, ctx->d_err, sizeof flags, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (flags) {
        CU(cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (flags & M1_ERRBIT_CAPACITY) return fail(ctx, M1CU_ERR_CAPACITY, "output buffer too small for the encoded payloads");
    if (flags & M1_ERRBIT_LEVEL) return fail(ctx, M1CU_ERR_LEVEL, "coded AC level with |L| >= 256 (outside the reference's encodable range)");
    return M1CU_OK;
}

// Host-buffer path for long sequences: the input is uploaded in sub-batches on a copy stream while
// earlier sub-batches are encoded on the compute stream and their payloads come back on a third
// stream, so the call costs little more than the host->device copy of the RGB bytes (PCIe).
// Every sub-batch owns a fixed region of d_out, so nothing has to be known on the host before the
// next launch.  Returns M1CU_ERR_CAPACITY when a region overflows (the caller then takes the simple
// path with worst-case sizing).
enum { M1_INTERNAL_REGION_OVERFLOW = -100 };   // never leaves this file: a sub-batch's payload region was too small

static int encode_host_pipelined_impl(m1cu_ctx *ctx, const uint8_t *h_rgb, int n_frames, uint8_t *h_out, size_t out_cap,
                                      uint32_t *h_frame_bytes, size_t *total_bytes, int S);

// Every exit leaves no copy in flight on the side streams: the caller may reuse or free h_rgb / h_out.
static int encode_host_pipelined(m1cu_ctx *ctx, const uint8_t *h_rgb, int n_frames, uint8_t *h_out, size_t out_cap,
                                 uint32_t *h_frame_bytes, size_t *total_bytes, int S)
{
    const int rc = encode_host_pipelined_impl(ctx, h_rgb, n_frames, h_out, out_cap, h_frame_bytes, total_bytes, S);
    if (rc != M1CU_OK) {
        if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->back_stream) cudaStreamSynchronize(ctx->back_stream);
    }
    return rc;
}

static int encode_host_pipelined_impl(m1cu_ctx *ctx, const uint8_t *h_rgb, int n_frames, uint8_t *h_out, size_t out_cap,
                                 uint32_t *h_frame_bytes, size_t *total_bytes, int S)
{
    const M1Geom &g = ctx->g;
    const int nsub = (n_frames + S - 1) / S;
    const size_t region = m1cu_typical_out_bytes(ctx, S);
    int rc;
    ctx->d_in_frames = 0;
    if ((rc = ensure(ctx, (void **)&ctx->d_in, &ctx->d_in_cap, (size_t)g.frame_stride * n_frames))) return rc;
    ctx->d_in_frames = n_frames;
    if ((rc = ensure(ctx, (void **)&ctx->d_out, &ctx->d_out_cap, region * nsub))) return rc;
    if ((rc = ensure(ctx, (void **)&ctx->h_out, &ctx->h_out_cap, region * nsub, true))) return rc;
    if (ctx->pipe_sub < nsub || ctx->pipe_frames < n_frames) {
        cudaFree(ctx->p_fbytes); cudaFree(ctx->p_foff);
        if (ctx->ph_fbytes) cudaFreeHost(ctx->ph_fbytes);
        if (ctx->ph_foff) cudaFreeHost(ctx->ph_foff);
        ctx->p_fbytes = nullptr; ctx->p_foff = nullptr; ctx->ph_fbytes = nullptr; ctx->ph_foff = nullptr;
        CU(cudaMalloc(&ctx->p_fbytes, sizeof(uint32_t) * n_frames));
        CU(cudaMalloc(&ctx->p_foff, sizeof(unsigned long long) * (size_t)nsub * (S + 1)));
        CU(cudaMallocHost(&ctx->ph_fbytes, sizeof(uint32_t) * n_frames));
        CU(cudaMallocHost(&ctx->ph_foff, sizeof(unsigned long long) * (size_t)nsub * (S + 1)));
        ctx->pipe_sub = nsub; ctx->pipe_frames = n_frames;
    }
    if (!ctx->copy_stream) CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->back_stream) CU(cudaStreamCreateWithFlags(&ctx->back_stream, cudaStreamNonBlocking));
    while ((int)ctx->pipe_events.size() < 3 * nsub) {
        cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_events.push_back(e);
    }
    cudaStream_t st = ctx->stream;
    // the copy stream must not start before work already queued on the compute stream (e.g. a
    // previous call's kernels still reading d_in)
    cudaEvent_t start = ctx->pipe_events[0];
    CU(cudaEventRecord(start, st));
    CU(cudaStreamWaitEvent(ctx->copy_stream, start, 0));
    // 1. enqueue everything that needs no host knowledge: uploads, encodes, metadata downloads
    for (int s = 0; s < nsub; ++s) {
        const int f0 = s * S, ns = n_frames - f0 < S ? n_frames - f0 : S;
        cudaEvent_t up = ctx->pipe_events[3 * s + 1], enc = ctx->pipe_events[3 * s + 2];
        CU(cudaMemcpyAsync(ctx->d_in + (size_t)f0 * g.frame_stride, h_rgb + (size_t)f0 * g.frame_stride,
                           (size_t)ns * g.frame_stride, cudaMemcpyHostToDevice, ctx->copy_stream));
        CU(cudaEventRecord(up, ctx->copy_stream));
        CU(cudaStreamWaitEvent(st, up, 0));
        rc = m1cu_encode_device(ctx, ctx->d_in + (size_t)f0 * g.frame_stride, ns, ctx->d_out + (size_t)s * region, region,
                                ctx->p_fbytes + f0, (uint64_t *)(ctx->p_foff + (size_t)s * (S + 1)), nullptr);
        if (rc) return rc;
        CU(cudaMemcpyAsync(ctx->ph_fbytes + f0, ctx->p_fbytes + f0, sizeof(uint32_t) * ns, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->ph_foff + (size_t)s * (S + 1), ctx->p_foff + (size_t)s * (S + 1),
                           sizeof(unsigned long long) * (ns + 1), cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(enc, st));
    }
    // 2. as each sub-batch finishes: fetch exactly its payload bytes, then compact the previous one
    size_t pos = 0;
    int flags = 0;
    auto compact = [&](int s) -> int {
        const int f0 = s * S, ns = n_frames - f0 < S ? n_frames - f0 : S;
        const unsigned long long *off = ctx->ph_foff + (size_t)s * (S + 1);
        for (int i = 0; i < ns; ++i) {
            const size_t nb = ctx->ph_fbytes[f0 + i];
            if (pos + nb > out_cap) return M1CU_ERR_CAPACITY;
            memcpy(h_out + pos, ctx->h_out + (size_t)s * region + off[i], nb);
            h_frame_bytes[f0 + i] = (uint32_t)nb;
            pos += nb;
        }
        return M1CU_OK;
    };
    for (int s = 0; s < nsub; ++s) {
        const int f0 = s * S, ns = n_frames - f0 < S ? n_frames - f0 : S;
        CU(cudaEventSynchronize(ctx->pipe_events[3 * s + 2]));
        const size_t end = (size_t)ctx->ph_foff[(size_t)s * (S + 1) + ns];
        if (end > region) flags |= M1_ERRBIT_CAPACITY;
        else if (end) CU(cudaMemcpyAsync(ctx->h_out + (size_t)s * region, ctx->d_out + (size_t)s * region, end,
                                         cudaMemcpyDeviceToHost, ctx->back_stream));
        if (s > 0 && !flags) {
            // payload s-1 was requested one iteration ago; wait for it, then copy it out while s streams
            CU(cudaStreamSynchronize(ctx->back_stream));      // covers s-1 and (shortly) s
            if ((rc = compact(s - 1))) return fail(ctx, rc, "m1cu_encode_host: h_out too small");
        }
    }
    rc = m1cu_check(ctx);                                      // device-side flags of all sub-batches
    if (rc == M1CU_ERR_CAPACITY) return fail(ctx, M1_INTERNAL_REGION_OVERFLOW, "payload region of a sub-batch overflowed");
    if (rc) return rc;
    if (flags) return fail(ctx, M1_INTERNAL_REGION_OVERFLOW, "payload region of a sub-batch overflowed");
    CU(cudaStreamSynchronize(ctx->back_stream));
    if ((rc = compact(nsub - 1))) return fail(ctx, rc, "m1cu_encode_host: h_out too small");
    if (total_bytes) *total_bytes = pos;
    return M1CU_OK;
}

int m1cu_encode_host(m1cu_ctx *ctx, const uint8_t *h_rgb, int n_frames, uint8_t *h_out, size_t out_cap,
                     uint32_t *h_frame_bytes, int16_t *h_levels, size_t *total_bytes)
{
    if (!ctx || !h_rgb || !h_out || !h_frame_bytes || n_frames <= 0)
        return fail(ctx, M1CU_ERR_ARG, "m1cu_encode_host: bad argument");
    if (n_frames > ctx->max_frames) return fail(ctx, M1CU_ERR_ARG, "m1cu_encode_host: n_frames > max_frames");
    const M1Geom &g = ctx->g;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    {
        // long sequences: overlap upload, encode and download (sub-batches of ~100 MB of input)
        int S = (int)(((size_t)100 << 20) / (size_t)g.frame_stride);
        if (S < 1) S = 1;
        if (!h_levels && n_frames >= 3 * S) {
            const int prc = encode_host_pipelined(ctx, h_rgb, n_frames, h_out, out_cap, h_frame_bytes, total_bytes, S);
            if (prc != M1_INTERNAL_REGION_OVERFLOW) return prc; // includes M1CU_ERR_CAPACITY: the caller's h_out is too small
            m1cu_check(ctx);                                   // rare: clear the flag, fall through to worst-case sizing below
        }
    }
    const size_t in_bytes = (size_t)g.frame_stride * n_frames;
    int rc;
    ctx->d_in_frames = 0;
    if ((rc = ensure(ctx, (void **)&ctx->d_in, &ctx->d_in_cap, in_bytes))) return rc;
    ctx->d_in_frames = n_frames;
    if (ctx->d_meta_frames < n_frames) {
        cudaFree(ctx->d_fbytes); cudaFree(ctx->d_foff); ctx->d_fbytes = nullptr; ctx->d_foff = nullptr;
        if (ctx->h_fbytes) cudaFreeHost(ctx->h_fbytes);
        if (ctx->h_foff) cudaFreeHost(ctx->h_foff);
        ctx->h_fbytes = nullptr; ctx->h_foff = nullptr; ctx->d_meta_frames = 0;
        CU(cudaMalloc(&ctx->d_fbytes, sizeof(uint32_t) * n_frames));
        CU(cudaMalloc(&ctx->d_foff, sizeof(unsigned long long) * (n_frames + 1)));
        CU(cudaMallocHost(&ctx->h_fbytes, sizeof(uint32_t) * n_frames));
        CU(cudaMallocHost(&ctx->h_foff, sizeof(unsigned long long) * (n_frames + 1)));
        ctx->d_meta_frames = n_frames;
    }
    if (h_levels) {
        const size_t lb = (size_t)n_frames * g.mbs_per_frame * 384 * sizeof(int16_t);
        if ((rc = ensure(ctx, (void **)&ctx->d_levels, &ctx->d_levels_cap, lb))) return rc;
    }
    CU(cudaMemcpyAsync(ctx->d_in, h_rgb, in_bytes, cudaMemcpyHostToDevice, st));

    size_t want = m1cu_typical_out_bytes(ctx, n_frames);
    for (int attempt = 0; attempt < 2; ++attempt) {
        if ((rc = ensure(ctx, (void **)&ctx->d_out, &ctx->d_out_cap, want))) return rc;
        rc = m1cu_encode_device(ctx, ctx->d_in, n_frames, ctx->d_out, ctx->d_out_cap, ctx->d_fbytes,
                                (uint64_t *)ctx->d_foff, h_levels ? ctx->d_levels : nullptr);
        if (rc) return rc;
        CU(cudaMemcpyAsync(ctx->h_fbytes, ctx->d_fbytes, sizeof(uint32_t) * n_frames, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(ctx->h_foff, ctx->d_foff, sizeof(unsigned long long) * (n_frames + 1), cudaMemcpyDeviceToHost, st));
        rc = m1cu_check(ctx);
        if (rc == M1CU_ERR_CAPACITY && attempt == 0) { want = m1cu_payload_bound(ctx) * (size_t)n_frames; continue; }
        if (rc) return rc;
        break;
    }
    const size_t dev_bytes = (size_t)ctx->h_foff[n_frames];
    if ((rc = ensure(ctx, (void **)&ctx->h_out, &ctx->h_out_cap, dev_bytes ? dev_bytes : 16, true))) return rc;
    CU(cudaMemcpyAsync(ctx->h_out, ctx->d_out, dev_bytes, cudaMemcpyDeviceToHost, st));
    if (h_levels)
        CU(cudaMemcpyAsync(h_levels, ctx->d_levels, (size_t)n_frames * g.mbs_per_frame * 384 * sizeof(int16_t),
                           cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    size_t pos = 0;
    for (int f = 0; f < n_frames; ++f) {
        const size_t n = ctx->h_fbytes[f];
        if (pos + n > out_cap) return fail(ctx, M1CU_ERR_CAPACITY, "m1cu_encode_host: h_out too small");
        memcpy(h_out + pos, ctx->h_out + ctx->h_foff[f], n);
        h_frame_bytes[f] = (uint32_t)n;
        pos += n;
    }
    if (total_bytes) *total_bytes = pos;
    return M1CU_OK;
}

int m1cu_ycbcr_planes(m1cu_ctx *ctx, const uint8_t *d_rgb, uint8_t *d_y, uint8_t *d_cb, uint8_t *d_cr)
{
    if (!ctx || !d_rgb || !d_y || !d_cb || !d_cr) return fail(ctx, M1CU_ERR_ARG, "m1cu_ycbcr_planes: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(m1k_launch_planes(d_rgb, ctx->g.channels, (size_t)ctx->g.W * ctx->g.H, 1, 0, 0, d_y, d_cb, d_cr, ctx->stream));
    ctx->launches += 1;
    return M1CU_OK;
}

// The .bit side files of a whole batch (source/image_processing.c:753-787, include/encoder.h:460-465)
// from the pictures the last host-buffer encode call left resident on the device: one launch, one
// download, no second upload.  h_planes receives per picture Y, Cb, Cr (width*height bytes each).
int m1cu_host_batch_planes(m1cu_ctx *ctx, int n_frames, uint8_t *h_planes, size_t cap)
{
    if (!ctx || !h_planes || n_frames <= 0) return fail(ctx, M1CU_ERR_ARG, "m1cu_host_batch_planes: bad argument");
    if (n_frames > ctx->d_in_frames || !ctx->d_in)
        return fail(ctx, M1CU_ERR_ARG, "m1cu_host_batch_planes: more pictures than the last host-buffer encode call uploaded");
    const size_t npix = (size_t)ctx->g.W * ctx->g.H, need = 3 * npix * (size_t)n_frames;
    if (cap < need) return fail(ctx, M1CU_ERR_CAPACITY, "m1cu_host_batch_planes: h_planes too small");
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, (void **)&ctx->d_planes, &ctx->d_planes_cap, need))) return rc;
    cudaStream_t st = ctx->stream;
    CU(m1k_launch_planes(ctx->d_in, ctx->g.channels, npix, n_frames, (size_t)ctx->g.frame_stride, 3 * npix,
                         ctx->d_planes, ctx->d_planes + npix, ctx->d_planes + 2 * npix, st));
    ctx->launches += 1;
    CU(cudaMemcpyAsync(h_planes, ctx->d_planes, need, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return M1CU_OK;
}

int m1cu_synth_rgb(m1cu_ctx *ctx, uint32_t seed, long first_frame, int n_frames, int kind, uint8_t *d_rgb)
{
    if (!ctx || !d_rgb || n_frames <= 0 || ctx->g.channels != 3)
        return fail(ctx, M1CU_ERR_ARG, "m1cu_synth_rgb: bad argument (needs a 3-channel context)");
    CU(cudaSetDevice(ctx->device));
    CU(m1k_launch_synth(seed, first_frame, n_frames, ctx->g.W, ctx->g.H, kind, d_rgb, ctx->stream));
    ctx->launches += 1;
    return M1CU_OK;
}

unsigned long long m1cu_launch_count(const m1cu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int m1cu_enable_timing(m1cu_ctx *ctx, int on)
{
    if (!ctx) return M1CU_ERR_ARG;
    ctx->timing = on != 0;
    return M1CU_OK;
}

int m1cu_kernel_times(m1cu_ctx *ctx, double ms[3], unsigned long long n[3])
{
    if (!ctx || !ms || !n) return M1CU_ERR_ARG;
    CU(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 3; ++k) { ms[k] = 0.0; n[k] = 0; }
    for (auto &s : ctx->spans) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) { ms[s.kind] += t; n[s.kind] += 1; }
        ctx->pool.push_back(s.a); ctx->pool.push_back(s.b);
    }
    ctx->spans.clear();
    return M1CU_OK;
}

void *m1cu_device_alloc(size_t bytes) { void *p = nullptr; if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void  m1cu_device_free(void *p) { if (p) cudaFree(p); }
void *m1cu_pinned_alloc(size_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void  m1cu_pinned_free(void *p) { if (p) cudaFreeHost(p); }
// write-combined pinned memory: the CPU must only WRITE it (reads are uncached and very slow); DMA reads skip the
// cache snoop, which can raise host->device throughput when several GPUs pull from one host memory system
void *m1cu_pinned_alloc_wc(size_t bytes) { void *p = nullptr; if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
// A pageable source is only staged when cudaMemcpy returns; the wait on the legacy stream makes the
// bytes visible to work submitted afterwards on the contexts' non-blocking streams.
int   m1cu_memcpy_h2d(void *dst, const void *src, size_t bytes)
{
    if (cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return M1CU_ERR_CUDA;
    return cudaStreamSynchronize(cudaStreamLegacy) == cudaSuccess ? M1CU_OK : M1CU_ERR_CUDA;
}
int   m1cu_memcpy_d2h(void *dst, const void *src, size_t bytes) { return cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? M1CU_OK : M1CU_ERR_CUDA; }

int m1cu_ipc_export(void *d_ptr, unsigned char handle[M1CU_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == M1CU_IPC_HANDLE_BYTES, "IPC handle size");
    if (!d_ptr || !handle) return fail(nullptr, M1CU_ERR_ARG, "m1cu_ipc_export: bad argument");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, d_ptr);
    if (e != cudaSuccess) return fail(nullptr, M1CU_ERR_CUDA, "cudaIpcGetMemHandle", e);
    memcpy(handle, &h, sizeof h);
    return M1CU_OK;
}

int m1cu_ipc_open(int device, const unsigned char handle[M1CU_IPC_HANDLE_BYTES], void **d_ptr)
{
    if (!handle || !d_ptr) return fail(nullptr, M1CU_ERR_ARG, "m1cu_ipc_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(nullptr, M1CU_ERR_CUDA, "cudaIpcOpenMemHandle", e);
    return M1CU_OK;
}

int m1cu_ipc_close(int device, void *d_ptr)
{
    if (!d_ptr) return M1CU_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaIpcCloseMemHandle(d_ptr);
    if (e != cudaSuccess) return fail(nullptr, M1CU_ERR_CUDA, "cudaIpcCloseMemHandle", e);
    return M1CU_OK;
}

int m1cu_push_payloads(m1cu_ctx *ctx, void *stream, uint8_t *dst, size_t dst_cap, const uint8_t *d_src,
                       const uint64_t *d_frame_offsets, int n_frames)
{
    if (!ctx || !dst || !d_src || !d_frame_offsets || n_frames <= 0 || (((uintptr_t)dst | (uintptr_t)d_src) & 15))
        return fail(ctx, M1CU_ERR_ARG, "m1cu_push_payloads: bad argument");
    CU(cudaSetDevice(ctx->device));
    CU(m1k_launch_push(dst, d_src, (const unsigned long long *)d_frame_offsets + n_frames, (unsigned long long)dst_cap,
                       stream ? (cudaStream_t)stream : ctx->stream));
    ctx->launches += 1;
    return M1CU_OK;
}

int m1cu_assemble_stream(m1cu_ctx *ctx, const uint8_t *d_payloads, const uint32_t *d_frame_bytes,
                         const uint64_t *d_frame_offsets, int n_frames, long first_frame_index,
                         const uint8_t *h_prefix256, const uint8_t *h_prologue, const uint8_t h_trailer[4],
                         uint8_t *d_stream, size_t stream_cap, size_t stream_offset, uint64_t *d_stream_bytes)
{
    if (!ctx || !d_payloads || !d_frame_bytes || !d_frame_offsets || n_frames <= 0 || !h_prefix256 || !h_trailer ||
        !d_stream || !d_stream_bytes || ((uintptr_t)d_stream & 15) || ((uintptr_t)d_payloads & 15))
        return fail(ctx, M1CU_ERR_ARG, "m1cu_assemble_stream: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    constexpr size_t kTmpl = 256 * 44, kAll = kTmpl + 32;           // templates, then the prologue (27 of 32 bytes)
    if (!ctx->d_stream_tmpl) CU(cudaMalloc(&ctx->d_stream_tmpl, kAll));
    std::vector<uint8_t> want(kAll, 0);
    memcpy(want.data(), h_prefix256, kTmpl);
    if (h_prologue) memcpy(want.data() + kTmpl, h_prologue, 27);
    if (want != ctx->h_stream_tmpl) {                                // first call, or the caller changed the headers
        CU(cudaStreamSynchronize(st));                               // nobody reads the old templates any more
        CU(cudaMemcpyAsync(ctx->d_stream_tmpl, want.data(), kAll, cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));                               // pageable source: landed before `want` goes away
        ctx->h_stream_tmpl.swap(want);
    }
    if (ctx->seg_frames < n_frames + 1) {
        CU(cudaStreamSynchronize(st));
        cudaFree(ctx->d_seg_off); ctx->d_seg_off = nullptr; ctx->seg_frames = 0;
        CU(cudaMalloc(&ctx->d_seg_off, sizeof(unsigned long long) * (size_t)(n_frames + 1)));
        ctx->seg_frames = n_frames + 1;
    }
    unsigned long long base = stream_offset;
    if (h_prologue) {
        if (stream_cap < stream_offset + 27) return fail(ctx, M1CU_ERR_CAPACITY, "m1cu_assemble_stream: stream_cap too small");
        CU(cudaMemcpyAsync(d_stream + stream_offset, ctx->d_stream_tmpl + kTmpl, 27, cudaMemcpyDeviceToDevice, st));
        base += 27;
    }
    const uint32_t trailer_be = ((uint32_t)h_trailer[0] << 24) | ((uint32_t)h_trailer[1] << 16) | ((uint32_t)h_trailer[2] << 8) | h_trailer[3];
    CU(m1k_launch_stream(d_payloads, d_frame_bytes, (const unsigned long long *)d_frame_offsets, n_frames, first_frame_index,
                         ctx->d_stream_tmpl, trailer_be, base, ctx->d_seg_off, d_stream, (unsigned long long)stream_cap,
                         ctx->d_err, st));
    CU(cudaMemcpyAsync(d_stream_bytes, ctx->d_seg_off + n_frames, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    ctx->launches += 2;
    return M1CU_OK;
}

int m1cu_encode_host_stream(m1cu_ctx *ctx, const uint8_t *h_rgb, int n_frames, long first_frame_index,
                            const uint8_t *h_prefix256, const uint8_t *h_prologue, const uint8_t h_trailer[4],
                            uint8_t *h_stream, size_t stream_cap, size_t *stream_bytes)
{
    if (!ctx || !h_rgb || !h_stream || !stream_bytes || n_frames <= 0 || !h_prefix256 || !h_trailer)
        return fail(ctx, M1CU_ERR_ARG, "m1cu_encode_host_stream: bad argument");
    if (n_frames > ctx->max_frames) return fail(ctx, M1CU_ERR_ARG, "m1cu_encode_host_stream: n_frames > max_frames");
    const M1Geom &g = ctx->g;
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t in_bytes = (size_t)g.frame_stride * n_frames;
    int rc;
    ctx->d_in_frames = 0;
    if ((rc = ensure(ctx, (void **)&ctx->d_in, &ctx->d_in_cap, in_bytes))) return rc;
    ctx->d_in_frames = n_frames;
    if (ctx->d_meta_frames < n_frames) {
        cudaFree(ctx->d_fbytes); cudaFree(ctx->d_foff); ctx->d_fbytes = nullptr; ctx->d_foff = nullptr;
        if (ctx->h_fbytes) cudaFreeHost(ctx->h_fbytes);
        if (ctx->h_foff) cudaFreeHost(ctx->h_foff);
        ctx->h_fbytes = nullptr; ctx->h_foff = nullptr; ctx->d_meta_frames = 0;
        CU(cudaMalloc(&ctx->d_fbytes, sizeof(uint32_t) * n_frames));
        CU(cudaMalloc(&ctx->d_foff, sizeof(unsigned long long) * (n_frames + 1)));
        CU(cudaMallocHost(&ctx->h_fbytes, sizeof(uint32_t) * n_frames));
        CU(cudaMallocHost(&ctx->h_foff, sizeof(unsigned long long) * (n_frames + 1)));
        ctx->d_meta_frames = n_frames;
    }
    if (!ctx->d_stream_end) CU(cudaMalloc(&ctx->d_stream_end, sizeof(unsigned long long)));
    CU(cudaMemcpyAsync(ctx->d_in, h_rgb, in_bytes, cudaMemcpyHostToDevice, st));
    size_t want = m1cu_typical_out_bytes(ctx, n_frames);
    unsigned long long end = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if ((rc = ensure(ctx, (void **)&ctx->d_out, &ctx->d_out_cap, want))) return rc;
        if ((rc = ensure(ctx, (void **)&ctx->d_stream, &ctx->d_stream_cap, want + 48 * (size_t)n_frames + 32))) return rc;
        if ((rc = m1cu_encode_device(ctx, ctx->d_in, n_frames, ctx->d_out, ctx->d_out_cap, ctx->d_fbytes,
                                     (uint64_t *)ctx->d_foff, nullptr))) return rc;
        if ((rc = m1cu_assemble_stream(ctx, ctx->d_out, ctx->d_fbytes, (const uint64_t *)ctx->d_foff, n_frames, first_frame_index,
                                       h_prefix256, h_prologue, h_trailer, ctx->d_stream, ctx->d_stream_cap, 0,
                                       (uint64_t *)ctx->d_stream_end))) return rc;
        CU(cudaMemcpyAsync(&end, ctx->d_stream_end, sizeof end, cudaMemcpyDeviceToHost, st));
        rc = m1cu_check(ctx);                                        // synchronises
        if (rc == M1CU_ERR_CAPACITY && attempt == 0) { want = m1cu_payload_bound(ctx) * (size_t)n_frames; continue; }
        if (rc) return rc;
        break;
    }
    if (end > stream_cap) return fail(ctx, M1CU_ERR_CAPACITY, "m1cu_encode_host_stream: h_stream too small");
    CU(cudaMemcpyAsync(h_stream, ctx->d_stream, (size_t)end, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *stream_bytes = (size_t)end;
    return M1CU_OK;
}

}  // extern "C"
