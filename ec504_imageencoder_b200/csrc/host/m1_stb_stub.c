/* Weak stand-ins used only when the library is built without stb_image (no JPEG decode): the
 * folder-based mpeg_encode_procedure then finds no loadable picture and returns -1, the in-memory
 * entry points (m1_encode_frames_to_*) are unaffected. */
#include <stddef.h>
__attribute__((weak)) unsigned char *stbi_load(char const *f, int *x, int *y, int *c, int d)
{ (void)f; (void)x; (void)y; (void)c; (void)d; return NULL; }
__attribute__((weak)) void stbi_image_free(void *p) { (void)p; }
__attribute__((weak)) const char *stbi_failure_reason(void) { return "libencoder was built without stb_image.h"; }
__attribute__((weak)) int stbi_info(char const *f, int *x, int *y, int *c) { (void)f; (void)x; (void)y; (void)c; return 0; }
