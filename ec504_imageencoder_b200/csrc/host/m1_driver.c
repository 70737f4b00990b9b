/*
 * m1_driver.c -- mpeg_encode_procedure: the reference's driver (include/encoder.h:20-498) with its
 * per-picture loop body (:216-445) replaced by one batched call into the CUDA library.
 *
 * Host side (this file): output file, pack + system header, folder scan in readdir order, stb_image
 * JPEG decode, per picture the packet / sequence / GOP / picture headers, packet-length patch,
 * 4-byte trailer, image_%d.bit side files, the reference's time bookkeeping and return codes.
 * Device side (include/m1cu.h): everything between the RGB bytes and the slice payload bytes.
 * There is no CPU encode path here: without a GPU the function reports the failure and returns -1.
 */
#include "encoder.h"
#include "m1cu.h"

#include <errno.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

/* stb_image v2.30 (public domain) is compiled into its own object by the Makefile from the copy
 * the reference vendors; only these entry points are used (reference include/encoder.h:162). */
extern unsigned char *stbi_load(char const *filename, int *x, int *y, int *channels_in_file, int desired_channels);
extern void stbi_image_free(void *retval_from_stbi_load);
extern const char *stbi_failure_reason(void);
extern int stbi_info(char const *filename, int *x, int *y, int *comp);


/* Pictures per GPU call: about 384 MB of input, so that 1080p batches (64 pictures) reach the overlapped
 * upload / encode / download path of m1cu_encode_host, at most 256 and at least 1. */
static int batch_size(size_t frame_bytes, int count)
{
    size_t b = ((size_t)384 << 20) / (frame_bytes ? frame_bytes : 1);
    if (b < 1) b = 1;
    if (b > 256) b = 256;
    return (int)b < count ? (int)b : count;
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

static int env_mode(void)
{
    const char *v = getenv("M1_MODE");
    if (v && (!strcmp(v, "full") || !strcmp(v, "FULL") || !strcmp(v, "0"))) return M1CU_MODE_FULL;
    return M1CU_MODE_REF_COMPAT;
}

Image *read_jpeg(const char *filename)
{
    Image *img = (Image *)malloc(sizeof(Image));
    if (!img) return NULL;
    img->data = stbi_load(filename, &img->width, &img->height, &img->channels, 0);
    if (!img->data) { free(img); return NULL; }
    return img;
}

void free_image(Image *img)
{
    if (!img) return;
    if (img->data) stbi_image_free(img->data);
    free(img);
}

/* The 44 bytes in front of picture `index`'s payload (reference include/encoder.h:196-231 and the
 * length patch :448-454), driven by the reference's clock: hour++ after every picture because
 * `minute % 60 == 0` always holds (:475-484), all three uint8_t. */
static void frame_prefix(long index, int width, int height, int mode, long payload, uint8_t out[44])
{
    const uint8_t hour = (uint8_t)index, minute = 0, second = 0;
    uint8_t bidir[4] = { 0, 0, 0, 0 };
    mpeg1_packet_header((uint32_t)(1 + second + minute * 60 + hour * 60 * 60), out);
    if (mode == M1CU_MODE_REF_COMPAT)      /* uint8_t width/height, include/encoder.h:186-187 */
        mpeg1_sequence_header((uint8_t)width, (uint8_t)height, 1, 4, 3, out + 16);
    else
        mpeg1_sequence_header((uint16_t)width, (uint16_t)height, 1, 4, 3, out + 16);
    mpeg1_gop(0, hour, minute, second, 0, 1, 0, out + 28);
    mpeg1_picture_header(0, 1, 0xffff, bidir, out + 36);
    const unsigned short fwd = (unsigned short)(44 + payload - 4 - 4);
    out[4] = (uint8_t)((fwd & 0xff00) >> 8);
    out[5] = (uint8_t)(fwd & 0xff);
}

void m1_stream_templates(int width, int height, int mode, unsigned char prefix256[256 * 44],
                         unsigned char prologue[27], unsigned char trailer[4])
{
    for (long i = 0; i < 256; ++i) frame_prefix(i, width, height, mode, 0, prefix256 + 44 * i);
    mpeg1_file_header(2202035, prologue);                 /* include/encoder.h:86 */
    mpeg1_sys_header(2202035, 0xe6, prologue + 12);       /* :88 */
    mpeg1_sequence_end(trailer);                          /* see encode_batch */
}

typedef int (*sink_fn)(void *cookie, const void *data, size_t n);

static int sink_file(void *cookie, const void *data, size_t n) { return fwrite(data, 1, n, (FILE *)cookie) == n ? 0 : -1; }

struct mem_sink { unsigned char *out; long cap, pos; };
static int sink_mem(void *cookie, const void *data, size_t n)
{
    struct mem_sink *m = (struct mem_sink *)cookie;
    if (m->pos + (long)n > m->cap) return -1;
    memcpy(m->out + m->pos, data, n);
    m->pos += (long)n;
    return 0;
}

/* Encodes pictures [first, first + n) (already packed back to back in `batch`) and emits
 * prefix + payload + trailer for each. */
static int encode_batch(m1cu_ctx *ctx, const unsigned char *batch, int n, long first, int width, int height,
                        int mode, unsigned char *payloads, size_t payload_cap, uint32_t *sizes,
                        sink_fn sink, void *cookie)
{
    if (env_int("M1_DEVICE_STREAM", 0)) {
        /* the GPU also places the header, payload and trailer bytes: one download, one write per batch */
        static unsigned char prefix256[256 * 44];
        unsigned char prologue[27], trailer[4];
        size_t bytes = 0;
        m1_stream_templates(width, height, mode, prefix256, prologue, trailer);
        const int src = m1cu_encode_host_stream(ctx, batch, n, first, prefix256, NULL, trailer, payloads, payload_cap, &bytes);
        if (src != M1CU_OK) {
            printf("Error: GPU encode failed (%d): %s\n", src, m1cu_last_error(ctx));
            return src;
        }
        return sink(cookie, payloads, bytes) ? M1CU_ERR_CAPACITY : M1CU_OK;
    }
    size_t total = 0;
    const int rc = m1cu_encode_host(ctx, batch, n, payloads, payload_cap, sizes, NULL, &total);
    if (rc != M1CU_OK) {
        printf("Error: GPU encode failed (%d): %s\n", rc, m1cu_last_error(ctx));
        return rc;
    }
    size_t pos = 0;
    for (int i = 0; i < n; ++i) {
        uint8_t prefix[44];
        /* the reference writes 4 uninitialised bytes after each picture (include/encoder.h:456-458,
         * the mpeg1_sequence_end call is commented out); we write the evidently intended end code */
        uint8_t trailer[4];
        frame_prefix(first + i, width, height, mode, (long)sizes[i], prefix);
        mpeg1_sequence_end(trailer);
        if (sink(cookie, prefix, 44) || sink(cookie, payloads + pos, sizes[i]) || sink(cookie, trailer, 4)) return M1CU_ERR_CAPACITY;
        pos += sizes[i];
    }
    return M1CU_OK;
}

static int encode_frames(const unsigned char *frames, int n_frames, int width, int height, int channels,
                         int quality, int mode, sink_fn sink, void *cookie)
{
    if (!frames || n_frames <= 0) return M1CU_ERR_ARG;
    m1cu_ctx *ctx = NULL;
    const int batch = batch_size((size_t)width * height * channels, n_frames);
    int rc = m1cu_create(&ctx, env_int("M1_DEVICE", 0), width, height, channels, mode, quality, batch);
    if (rc != M1CU_OK) { printf("Error: cannot create the GPU encoder (%d): %s\n", rc, m1cu_last_error(NULL)); return rc; }
    const size_t cap = (m1cu_payload_bound(ctx) + 48) * (size_t)batch;   /* + headers and trailer when the GPU assembles the stream */
    unsigned char *payloads = (unsigned char *)malloc(cap);
    uint32_t *sizes = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)batch);
    uint8_t prologue[27];
    mpeg1_file_header(2202035, prologue);                 /* include/encoder.h:86 */
    mpeg1_sys_header(2202035, 0xe6, prologue + 12);       /* :88 */
    rc = (payloads && sizes) ? (sink(cookie, prologue, 27) ? M1CU_ERR_CAPACITY : M1CU_OK) : M1CU_ERR_ARG;
    const size_t fsz = (size_t)width * height * channels;
    for (int f0 = 0; rc == M1CU_OK && f0 < n_frames; f0 += batch) {
        const int n = n_frames - f0 < batch ? n_frames - f0 : batch;
        rc = encode_batch(ctx, frames + (size_t)f0 * fsz, n, f0, width, height, mode, payloads, cap, sizes, sink, cookie);
    }
    free(payloads); free(sizes);
    m1cu_destroy(ctx);
    return rc;
}

int m1_encode_frames_to_file(const char *video_path, const unsigned char *frames, int n_frames,
                             int width, int height, int channels, int quality_factor, int mode)
{
    FILE *fp = fopen(video_path, "wb");
    if (!fp) { perror("Error opening mpeg file"); return 1; }
    const int rc = encode_frames(frames, n_frames, width, height, channels, quality_factor, mode, sink_file, fp);
    fclose(fp);
    return rc == M1CU_OK ? 0 : -1;
}

long m1_encode_frames_to_memory(const unsigned char *frames, int n_frames, int width, int height,
                                int channels, int quality_factor, int mode, unsigned char *out, long cap)
{
    struct mem_sink m = { out, cap, 0 };
    const int rc = encode_frames(frames, n_frames, width, height, channels, quality_factor, mode, sink_mem, &m);
    return rc == M1CU_OK ? m.pos : (long)rc;
}

/* image_%d.bit (include/encoder.h:460-465) for the n pictures of the batch the GPU just encoded: the
 * full-resolution planes come from the copy of the batch that is still resident on the device (one
 * launch and one download per batch, m1cu_host_batch_planes). */
static void write_bit_files(m1cu_ctx *ctx, const char *folder, long first_index, int n, int width, int height,
                            unsigned char *h_planes, size_t cap)
{
    const size_t np = (size_t)width * height;
    if (m1cu_host_batch_planes(ctx, n, h_planes, cap) != M1CU_OK) return;
    for (int i = 0; i < n; ++i) {
        char name[1100];
        unsigned char *p = h_planes + 3 * np * (size_t)i;
        snprintf(name, sizeof name, "%s/image_%ld.bit", folder, first_index + i + 1);
        write_to_bitstream(name, p, p + np, p + 2 * np, width, height);
    }
}

/* ---- decode pipeline ---------------------------------------------------------------------------------
 * The reference decodes every JPEG before it encodes the first picture (include/encoder.h:118-172).  Here
 * worker threads decode straight into a ring of two pinned batch buffers while the calling thread feeds
 * finished batches to the GPU and writes their bytes, so JPEG decode, PCIe upload, encode and file output
 * overlap from the folder scan on.  Picture order stays readdir order. */
typedef struct {
    const char *folder;
    char **names;
    int n_names;
    int width, height, channels;
    size_t fsz;
    int batch;
    unsigned char *ring[2];
    unsigned char *ok;              /* per file: 1 decoded into its slot, 0 skipped (as the reference skips it) */
    pthread_mutex_t mu;
    pthread_cond_t cv;
    int next;                       /* next file index to claim */
    int armed[2];                   /* batch number a ring slot currently accepts, -1 = none */
    int done[2];                    /* files of that batch finished (decoded or skipped) */
    int mismatch;                   /* a decoded picture did not have the announced geometry */
} decode_pipe;

static void *decode_worker(void *arg)
{
    decode_pipe *p = (decode_pipe *)arg;
    char path[1100];
    for (;;) {
        pthread_mutex_lock(&p->mu);
        int i = -1, slot = 0;
        while (p->next < p->n_names) {
            const int b = p->next / p->batch;
            slot = b & 1;
            if (p->armed[slot] == b) { i = p->next++; break; }
            pthread_cond_wait(&p->cv, &p->mu);             /* the slot still holds an older batch */
        }
        pthread_mutex_unlock(&p->mu);
        if (i < 0) return NULL;
        snprintf(path, sizeof path, "%s/%s", p->folder, p->names[i]);
        int w = 0, h = 0, c = 0, good = 0, bad_geometry = 0;
        unsigned char *px = stbi_load(path, &w, &h, &c, 0);
        if (px) {
            if (w == p->width && h == p->height && c == p->channels) {
                memcpy(p->ring[slot] + (size_t)(i % p->batch) * p->fsz, px, p->fsz);
                good = 1;
            } else bad_geometry = 1;
            stbi_image_free(px);
        }
        pthread_mutex_lock(&p->mu);
        p->ok[i] = (unsigned char)good;
        if (bad_geometry) p->mismatch = 1;
        p->done[slot]++;
        pthread_cond_broadcast(&p->cv);
        pthread_mutex_unlock(&p->mu);
    }
}

/* All pictures already decoded (the reference's order of work); used when the header pre-scan cannot
 * vouch for the folder and when M1_DECODE_THREADS=0. */
static int encode_loaded(Image **images, int count, const char *bitstream_folder, int quality_factor, FILE *fp)
{
    const int width = images[0]->width, height = images[0]->height, channels = images[0]->channels;
    const int mode = env_mode();
    const size_t fsz = (size_t)width * height * channels;
    const int batch = batch_size(fsz, count);
    m1cu_ctx *ctx = NULL;
    int rc = m1cu_create(&ctx, env_int("M1_DEVICE", 0), width, height, channels, mode, quality_factor, batch);
    if (rc != M1CU_OK) {
        printf("Error: cannot create the GPU encoder (%d): %s\n", rc, m1cu_last_error(NULL));
        return -1;
    }
    const size_t cap = (m1cu_payload_bound(ctx) + 48) * (size_t)batch;   /* + headers and trailer when the GPU assembles the stream */
    const int bit_files = env_int("M1_BIT_FILES", 1);
    const size_t pcap = bit_files ? (size_t)3 * width * height * batch : 0;
    unsigned char *staging = (unsigned char *)m1cu_pinned_alloc(fsz * (size_t)batch);
    unsigned char *payloads = (unsigned char *)malloc(cap);
    uint32_t *sizes = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)batch);
    unsigned char *h_planes = bit_files ? (unsigned char *)m1cu_pinned_alloc(pcap) : NULL;
    rc = (staging && payloads && sizes) ? 0 : -1;
    for (int f0 = 0; rc == 0 && f0 < count; f0 += batch) {
        const int n = count - f0 < batch ? count - f0 : batch;
        for (int i = 0; i < n; ++i) memcpy(staging + (size_t)i * fsz, images[f0 + i]->data, fsz);
        if (encode_batch(ctx, staging, n, f0, width, height, mode, payloads, cap, sizes, sink_file, fp) != M1CU_OK) rc = -1;
        if (rc == 0 && h_planes) write_bit_files(ctx, bitstream_folder, f0, n, width, height, h_planes, pcap);
    }
    m1cu_pinned_free(staging); free(payloads); free(sizes); m1cu_pinned_free(h_planes);
    m1cu_destroy(ctx);
    return rc;
}

/* Decode -> GPU pipeline over the file names (all headers agreed on width x height x channels). */
static int encode_streaming(const char *images_folder, char **names, int n_names, int width, int height, int channels,
                            const char *bitstream_folder, int quality_factor, int threads, FILE *fp)
{
    const int mode = env_mode();
    decode_pipe p;
    memset(&p, 0, sizeof p);
    p.folder = images_folder; p.names = names; p.n_names = n_names;
    p.width = width; p.height = height; p.channels = channels;
    p.fsz = (size_t)width * height * channels;
    p.batch = batch_size(p.fsz, n_names);
    m1cu_ctx *ctx = NULL;
    int rc = m1cu_create(&ctx, env_int("M1_DEVICE", 0), width, height, channels, mode, quality_factor, p.batch);
    if (rc != M1CU_OK) {
        printf("Error: cannot create the GPU encoder (%d): %s\n", rc, m1cu_last_error(NULL));
        return -1;
    }
    const size_t cap = (m1cu_payload_bound(ctx) + 48) * (size_t)p.batch;
    const int bit_files = env_int("M1_BIT_FILES", 1);
    const size_t pcap = bit_files ? (size_t)3 * width * height * p.batch : 0;
    const int n_batches = (n_names + p.batch - 1) / p.batch;
    p.ring[0] = (unsigned char *)m1cu_pinned_alloc(p.fsz * (size_t)p.batch);
    p.ring[1] = n_batches > 1 ? (unsigned char *)m1cu_pinned_alloc(p.fsz * (size_t)p.batch) : p.ring[0];
    p.ok = (unsigned char *)calloc((size_t)n_names, 1);
    unsigned char *payloads = (unsigned char *)malloc(cap);
    uint32_t *sizes = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)p.batch);
    unsigned char *h_planes = bit_files ? (unsigned char *)m1cu_pinned_alloc(pcap) : NULL;
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    int started = 0;
    rc = (p.ring[0] && p.ring[1] && p.ok && payloads && sizes && tid) ? 0 : -1;
    if (rc == 0) {
        pthread_mutex_init(&p.mu, NULL);
        pthread_cond_init(&p.cv, NULL);
        p.armed[0] = 0; p.armed[1] = n_batches > 1 ? 1 : -1;
        for (; started < threads; ++started)
            if (pthread_create(&tid[started], NULL, decode_worker, &p) != 0) break;
        if (started == 0) rc = -1;
    }
    long emitted = 0;                                             /* pictures written so far = index of the next one */
    for (int b = 0; rc == 0 && b < n_batches; ++b) {
        const int slot = b & 1, i0 = b * p.batch;
        const int nb = n_names - i0 < p.batch ? n_names - i0 : p.batch;
        pthread_mutex_lock(&p.mu);
        while (p.done[slot] < nb) pthread_cond_wait(&p.cv, &p.mu);
        const int mismatch = p.mismatch;
        pthread_mutex_unlock(&p.mu);
        if (mismatch) { printf("Image dimensions do not match.\n"); rc = -1; break; }
        /* compact the slot over the files stb could not decode (the reference skips them too) */
        int n = 0;
        for (int i = 0; i < nb; ++i) {
            if (!p.ok[i0 + i]) { printf("Error loading image %s/%s\n", images_folder, names[i0 + i]); continue; }
            if (n != i) memmove(p.ring[slot] + (size_t)n * p.fsz, p.ring[slot] + (size_t)i * p.fsz, p.fsz);
            printf("Loaded image: %s (Width: %d, Height: %d)\n", names[i0 + i], width, height);
            ++n;
        }
        if (n > 0) {
            if (encode_batch(ctx, p.ring[slot], n, emitted, width, height, mode, payloads, cap, sizes, sink_file, fp) != M1CU_OK) rc = -1;
            if (rc == 0 && h_planes) write_bit_files(ctx, bitstream_folder, emitted, n, width, height, h_planes, pcap);
            emitted += n;
        }
        pthread_mutex_lock(&p.mu);                                /* hand the slot to batch b + 2 */
        p.done[slot] = 0;
        p.armed[slot] = b + 2 < n_batches ? b + 2 : -1;
        pthread_cond_broadcast(&p.cv);
        pthread_mutex_unlock(&p.mu);
    }
    if (started) {
        pthread_mutex_lock(&p.mu);                                /* error exit: release workers still waiting for a slot */
        p.next = p.n_names;
        pthread_cond_broadcast(&p.cv);
        pthread_mutex_unlock(&p.mu);
        for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
        pthread_mutex_destroy(&p.mu);
        pthread_cond_destroy(&p.cv);
    }
    if (rc == 0 && emitted == 0) { printf("Image dimensions do not match.\n"); rc = -1; }   /* check_dimensions(…, 0) == 0 */
    if (p.ring[1] != p.ring[0]) m1cu_pinned_free(p.ring[1]);
    m1cu_pinned_free(p.ring[0]); free(p.ok); free(payloads); free(sizes); free(tid); m1cu_pinned_free(h_planes);
    m1cu_destroy(ctx);
    return rc;
}

int mpeg_encode_procedure(const char *images_folder, const char *bitstream_folder, const char *video_path,
                          int quality_factor)
{
    FILE *fp = fopen(video_path, "wb");                                   /* include/encoder.h:75-80 */
    if (fp == NULL) { perror("Error opening mpeg file"); return 1; }
    {
        uint8_t prologue[27];                                             /* :85-89: written before anything can fail */
        mpeg1_file_header(2202035, prologue);
        mpeg1_sys_header(2202035, 0xe6, prologue + 12);
        fwrite(prologue, 1, 27, fp);
    }

    struct stat st;
    if (stat(bitstream_folder, &st) == -1) {                              /* :104-108 */
        mkdir(bitstream_folder, 0700);
        printf("Created directory for bitstreams: %s\n", bitstream_folder);
    }
    if (stat(images_folder, &st) == -1) {                                 /* :111-116 */
        fclose(fp);
        mkdir(images_folder, 0700);
        printf("Created directory for images: %s\n", images_folder);
        printf("Please add your .jpg images in the '%s' folder and rerun the program.\n", images_folder);
        return 0;
    }
    DIR *dir = opendir(images_folder);                                    /* :119-124 */
    if (!dir) { printf("Error: Could not open images directory.\n"); fclose(fp); return -1; }

    /* the file names in readdir order (:140-171), and what their headers announce */
    int n_names = 0, capacity = 100;
    char **names = (char **)malloc(sizeof(char *) * (size_t)capacity);
    if (!names) { printf("Error: Memory allocation failed for images array.\n"); closedir(dir); fclose(fp); return -1; }
    struct dirent *entry;
    char path[1100];
    while ((entry = readdir(dir)) != NULL) {
        if (!strstr(entry->d_name, ".jpg") && !strstr(entry->d_name, ".jpeg")) continue;
        if (n_names >= capacity) {
            capacity *= 2;
            char **grown = (char **)realloc(names, sizeof(char *) * (size_t)capacity);
            if (!grown) { printf("Error: Memory reallocation failed for images array.\n"); closedir(dir); fclose(fp); return -1; }
            names = grown;
        }
        names[n_names] = strdup(entry->d_name);
        if (names[n_names]) ++n_names;
    }
    closedir(dir);

    int rc = 0;
    int threads = env_int("M1_DECODE_THREADS", -1);
    if (threads < 0) {
        const long cores = sysconf(_SC_NPROCESSORS_ONLN);
        threads = cores > 16 ? 16 : (cores < 1 ? 1 : (int)cores);
    }
    /* Streaming needs the geometry before the first decode: every header must parse and agree.  Otherwise
     * (or with M1_DECODE_THREADS=0) do what the reference does: decode everything, then check, then encode. */
    int width = 0, height = 0, channels = 0, uniform = threads > 0 && n_names > 0;
    for (int i = 0; uniform && i < n_names; ++i) {
        int w = 0, h = 0, c = 0;
        snprintf(path, sizeof path, "%s/%s", images_folder, names[i]);
        if (!stbi_info(path, &w, &h, &c)) uniform = 0;
        else if (i == 0) { width = w; height = h; channels = c; }
        else if (w != width || h != height || c != channels) uniform = 0;
    }
    if (uniform && channels >= 3) {
        rc = encode_streaming(images_folder, names, n_names, width, height, channels, bitstream_folder, quality_factor,
                              threads, fp);
    } else {
        int count = 0;
        Image **images = (Image **)malloc(sizeof(Image *) * (size_t)(n_names > 0 ? n_names : 1));
        if (!images) { printf("Error: Memory allocation failed for images array.\n"); rc = -1; }
        for (int i = 0; images && i < n_names; ++i) {
            snprintf(path, sizeof path, "%s/%s", images_folder, names[i]);
            Image *img = read_jpeg(path);
            if (!img) { printf("Error loading image %s: %s\n", path, stbi_failure_reason()); continue; }
            images[count++] = img;
            printf("Loaded image: %s (Width: %d, Height: %d)\n", names[i], img->width, img->height);
        }
        if (images) {
            int ok = check_dimensions(images, count);                     /* :175-183 */
            for (int i = 1; ok && i < count; ++i) ok = images[i]->channels == images[0]->channels;
            if (!ok || images[0]->channels < 3) {
                printf("Image dimensions do not match.\n");
                rc = -1;
            } else {
                rc = encode_loaded(images, count, bitstream_folder, quality_factor, fp);
            }
            for (int i = 0; i < count; ++i) free_image(images[i]);
            free(images);
        }
    }
    for (int i = 0; i < n_names; ++i) free(names[i]);
    free(names);
    fclose(fp);
    if (rc == 0) printf("Image processing finished.\n");
    return rc;
}
