/* jni.h -- TEST STAND-IN for the JDK header (this image has no JDK): just enough of the JNI C interface to
 * compile encoder_jni.c unchanged and call Java_com_example_Encoder_mpegEncodeProcedure from a test.
 * The function table keeps the real slot numbers of the two calls the shim makes (JNI specification,
 * "Interface Function Table": GetStringUTFChars = 169, ReleaseStringUTFChars = 170), so the object code of
 * the shim is the same a JDK build would produce.  Reference: /root/reference/encoder_jni.c:5-22. */
#ifndef M1_TEST_JNI_H
#define M1_TEST_JNI_H

typedef int jint;
typedef unsigned char jboolean;
typedef struct m1_fake_object *jobject;
typedef jobject jstring;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_ *JNIEnv;

struct JNINativeInterface_ {
    void *slots_0_168[169];
    const char *(*GetStringUTFChars)(JNIEnv *env, jstring str, jboolean *is_copy);
    void (*ReleaseStringUTFChars)(JNIEnv *env, jstring str, const char *chars);
};

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

#endif
