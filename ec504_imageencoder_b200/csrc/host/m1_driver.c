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
#include <stdint.h>
#include <string.h>

/* stb_image v2.30 (public domain) is compiled into its own object by the Makefile from the copy
 * the reference vendors; only these three entry points are used (reference include/encoder.h:162). */
extern unsigned char *stbi_load(char const *filename, int *x, int *y, int *channels_in_file, int desired_channels);
extern void stbi_image_free(void *retval_from_stbi_load);
extern const char *stbi_failure_reason(void);

#define M1_BATCH 32      /* pictures per GPU call */

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
    const int batch = n_frames < M1_BATCH ? n_frames : M1_BATCH;
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

/* image_%d.bit (include/encoder.h:460-465): full-resolution planes from the device conversion */
static void write_bit_file(m1cu_ctx *ctx, const char *folder, long index, const unsigned char *rgb,
                           int width, int height, int channels, void *d_rgb, void *d_planes, unsigned char *h_planes)
{
    const size_t n = (size_t)width * height;
    unsigned char *dp = (unsigned char *)d_planes;
    if (m1cu_memcpy_h2d(d_rgb, rgb, n * channels) != M1CU_OK) return;
    if (m1cu_ycbcr_planes(ctx, (const uint8_t *)d_rgb, dp, dp + n, dp + 2 * n) != M1CU_OK) return;
    if (m1cu_synchronize(ctx) != M1CU_OK) return;
    if (m1cu_memcpy_d2h(h_planes, d_planes, 3 * n) != M1CU_OK) return;
    char name[512];
    snprintf(name, sizeof name, "%s/image_%ld.bit", folder, index + 1);
    write_to_bitstream(name, h_planes, h_planes + n, h_planes + 2 * n, width, height);
}

int mpeg_encode_procedure(const char *images_folder, const char *bitstream_folder, const char *video_path,
                          int quality_factor)
{
    FILE *fp = fopen(video_path, "wb");                                   /* include/encoder.h:75-80 */
    if (fp == NULL) { perror("Error opening mpeg file"); return 1; }

    struct stat st;
    if (stat(bitstream_folder, &st) == -1) {                              /* :104-108 */
        mkdir(bitstream_folder, 0700);
        printf("Created directory for bitstreams: %s\n", bitstream_folder);
    }
    if (stat(images_folder, &st) == -1) {                                 /* :111-116 */
        uint8_t prologue[27];
        mpeg1_file_header(2202035, prologue);
        mpeg1_sys_header(2202035, 0xe6, prologue + 12);
        fwrite(prologue, 1, 27, fp);                                      /* the reference has written these by now */
        fclose(fp);
        mkdir(images_folder, 0700);
        printf("Created directory for images: %s\n", images_folder);
        printf("Please add your .jpg images in the '%s' folder and rerun the program.\n", images_folder);
        return 0;
    }
    DIR *dir = opendir(images_folder);                                    /* :119-124 */
    if (!dir) { printf("Error: Could not open images directory.\n"); fclose(fp); return -1; }

    int count = 0, capacity = 100;
    Image **images = (Image **)malloc(sizeof(Image *) * (size_t)capacity);
    if (!images) { printf("Error: Memory allocation failed for images array.\n"); closedir(dir); fclose(fp); return -1; }
    struct dirent *entry;
    char path[1024];
    while ((entry = readdir(dir)) != NULL) {                              /* :140-171, readdir order */
        if (!strstr(entry->d_name, ".jpg") && !strstr(entry->d_name, ".jpeg")) continue;
        if (count >= capacity) {
            capacity *= 2;
            Image **grown = (Image **)realloc(images, sizeof(Image *) * (size_t)capacity);
            if (!grown) { printf("Error: Memory reallocation failed for images array.\n"); closedir(dir); fclose(fp); return -1; }
            images = grown;
        }
        snprintf(path, sizeof path, "%s/%s", images_folder, entry->d_name);
        Image *img = read_jpeg(path);
        if (!img) { printf("Error loading image %s: %s\n", path, stbi_failure_reason()); continue; }
        images[count++] = img;
        printf("Loaded image: %s (Width: %d, Height: %d)\n", entry->d_name, img->width, img->height);
    }
    closedir(dir);

    int rc = 0;
    int ok = check_dimensions(images, count);                             /* :175-183 */
    for (int i = 1; ok && i < count; ++i) ok = images[i]->channels == images[0]->channels;
    if (!ok || images[0]->channels < 3) {
        printf("Image dimensions do not match.\n");
        rc = -1;
    } else {
        const int width = images[0]->width, height = images[0]->height, channels = images[0]->channels;
        const int mode = env_mode();
        const size_t fsz = (size_t)width * height * channels;
        const int batch = count < M1_BATCH ? count : M1_BATCH;
        m1cu_ctx *ctx = NULL;
        rc = m1cu_create(&ctx, env_int("M1_DEVICE", 0), width, height, channels, mode, quality_factor, batch);
        if (rc != M1CU_OK) {
            printf("Error: cannot create the GPU encoder (%d): %s\n", rc, m1cu_last_error(NULL));
            rc = -1;
        } else {
            const size_t cap = (m1cu_payload_bound(ctx) + 48) * (size_t)batch;   /* + headers and trailer when the GPU assembles the stream */
            unsigned char *staging = (unsigned char *)m1cu_pinned_alloc(fsz * (size_t)batch);
            unsigned char *payloads = (unsigned char *)malloc(cap);
            uint32_t *sizes = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)batch);
            const int bit_files = env_int("M1_BIT_FILES", 1);
            void *d_rgb = bit_files ? m1cu_device_alloc(fsz) : NULL;
            void *d_planes = bit_files ? m1cu_device_alloc((size_t)width * height * 3) : NULL;
            unsigned char *h_planes = bit_files ? (unsigned char *)malloc((size_t)width * height * 3) : NULL;
            uint8_t prologue[27];
            mpeg1_file_header(2202035, prologue);
            mpeg1_sys_header(2202035, 0xe6, prologue + 12);
            fwrite(prologue, 1, 27, fp);
            if (!staging || !payloads || !sizes) rc = -1;
            for (int f0 = 0; rc == 0 && f0 < count; f0 += batch) {
                const int n = count - f0 < batch ? count - f0 : batch;
                for (int i = 0; i < n; ++i) memcpy(staging + (size_t)i * fsz, images[f0 + i]->data, fsz);
                if (encode_batch(ctx, staging, n, f0, width, height, mode, payloads, cap, sizes, sink_file, fp) != M1CU_OK) rc = -1;
                for (int i = 0; rc == 0 && bit_files && d_rgb && d_planes && h_planes && i < n; ++i)
                    write_bit_file(ctx, bitstream_folder, f0 + i, images[f0 + i]->data, width, height, channels,
                                   d_rgb, d_planes, h_planes);
            }
            m1cu_pinned_free(staging); free(payloads); free(sizes); free(h_planes);
            m1cu_device_free(d_rgb); m1cu_device_free(d_planes);
            m1cu_destroy(ctx);
        }
    }
    for (int i = 0; i < count; ++i) free_image(images[i]);
    free(images);
    fclose(fp);
    if (rc == 0) printf("Image processing finished.\n");
    return rc;
}
