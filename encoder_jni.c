/* encoder_jni.c -- JNI shim for com.example.Encoder.mpegEncodeProcedure(String, String, String, int),
 * the same native method the reference exports (encoder_jni.c:5-22).  Built by `make jni` where a
 * JDK provides <jni.h> (not available in this image). */
#include <jni.h>
#include "include/encoder.h"

JNIEXPORT jint JNICALL Java_com_example_Encoder_mpegEncodeProcedure(JNIEnv *env, jobject self, jstring images_folder,
                                                                    jstring bitstream_folder, jstring video_path,
                                                                    jint quality)
{
    (void)self;
    const char *images = (*env)->GetStringUTFChars(env, images_folder, NULL);
    const char *streams = (*env)->GetStringUTFChars(env, bitstream_folder, NULL);
    const char *video = (*env)->GetStringUTFChars(env, video_path, NULL);
    jint rc = -1;
    if (images && streams && video) rc = (jint)mpeg_encode_procedure(images, streams, video, (int)quality);
    if (video) (*env)->ReleaseStringUTFChars(env, video_path, video);
    if (streams) (*env)->ReleaseStringUTFChars(env, bitstream_folder, streams);
    if (images) (*env)->ReleaseStringUTFChars(env, images_folder, images);
    return rc;
}
