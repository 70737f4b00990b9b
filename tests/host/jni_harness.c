/* jni_harness.c -- TEST HARNESS: a three-string fake JVM around the JNI shim (encoder_jni.c compiled against
 * tests/host/jni.h).  Java strings are modelled as objects holding a C string; GetStringUTFChars hands out a
 * fresh copy (as a JVM may), ReleaseStringUTFChars checks that exactly that copy comes back for that string. */
#include "jni.h"
#include <stdlib.h>
#include <string.h>

struct m1_fake_object { const char *utf; char *lent; int gets, releases, bad_release; };

static const char *get_chars(JNIEnv *env, jstring s, jboolean *is_copy)
{
    (void)env;
    if (is_copy) *is_copy = 1;
    s->gets++;
    s->lent = strdup(s->utf);
    return s->lent;
}

static void release_chars(JNIEnv *env, jstring s, const char *chars)
{
    (void)env;
    s->releases++;
    if (chars != s->lent || !s->lent) { s->bad_release++; return; }
    free(s->lent);
    s->lent = NULL;
}

extern jint Java_com_example_Encoder_mpegEncodeProcedure(JNIEnv *, jobject, jstring, jstring, jstring, jint);

/* counts[0..2] = GetStringUTFChars calls per string, [3..5] = releases, [6] = releases of a wrong pointer,
 * [7] = copies still outstanding after the call */
int m1jni_call(const char *images, const char *streams, const char *video, int quality, int counts[8])
{
    static struct JNINativeInterface_ table;
    table.GetStringUTFChars = get_chars;
    table.ReleaseStringUTFChars = release_chars;
    JNIEnv env = &table;
    struct m1_fake_object s[3] = { { images, 0, 0, 0, 0 }, { streams, 0, 0, 0, 0 }, { video, 0, 0, 0, 0 } };
    const jint rc = Java_com_example_Encoder_mpegEncodeProcedure(&env, 0, &s[0], &s[1], &s[2], quality);
    counts[6] = counts[7] = 0;
    for (int i = 0; i < 3; ++i) {
        counts[i] = s[i].gets; counts[3 + i] = s[i].releases;
        counts[6] += s[i].bad_release; counts[7] += s[i].lent != 0;
    }
    return rc;
}
