/*
 * TEST-ONLY stand-in for a JDK's jni.h: just enough of the JNI types and of the JNIEnv function table (by name and
 * signature, not by slot order) for `gcc -fsyntax-only jni/vce_jni.c` to type-check the binding in an image without a JDK.
 * Never used to build anything that runs.
 */
#ifndef VCE_TEST_JNI_STUB_H
#define VCE_TEST_JNI_STUB_H
#include <stdint.h>

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef double jdouble;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jbyteArray;
typedef jarray jdoubleArray;
typedef jarray jobjectArray;
struct _jfieldID;
typedef struct _jfieldID* jfieldID;

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_COMMIT 1

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;
struct JNINativeInterface_ {
  jclass (*GetObjectClass)(JNIEnv*, jobject);
  jfieldID (*GetFieldID)(JNIEnv*, jclass, const char*, const char*);
  jobject (*GetObjectField)(JNIEnv*, jobject, jfieldID);
  void (*SetObjectField)(JNIEnv*, jobject, jfieldID, jobject);
  jint (*GetIntField)(JNIEnv*, jobject, jfieldID);
  void (*SetIntField)(JNIEnv*, jobject, jfieldID, jint);
  jsize (*GetArrayLength)(JNIEnv*, jarray);
  jobject (*GetObjectArrayElement)(JNIEnv*, jobjectArray, jsize);
  void* (*GetPrimitiveArrayCritical)(JNIEnv*, jarray, jboolean*);
  void (*ReleasePrimitiveArrayCritical)(JNIEnv*, jarray, void*, jint);
  const char* (*GetStringUTFChars)(JNIEnv*, jstring, jboolean*);
  void (*ReleaseStringUTFChars)(JNIEnv*, jstring, const char*);
  jstring (*NewStringUTF)(JNIEnv*, const char*);
  jintArray (*NewIntArray)(JNIEnv*, jsize);
  jlongArray (*NewLongArray)(JNIEnv*, jsize);
  jbyteArray (*NewByteArray)(JNIEnv*, jsize);
  jdoubleArray (*NewDoubleArray)(JNIEnv*, jsize);
  void (*SetDoubleArrayRegion)(JNIEnv*, jdoubleArray, jsize, jsize, const jdouble*);
  jint* (*GetIntArrayElements)(JNIEnv*, jintArray, jboolean*);
  void (*ReleaseIntArrayElements)(JNIEnv*, jintArray, jint*, jint);
  void (*GetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, jint*);
  void (*SetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, const jint*);
  void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
  void (*GetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*);
  void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
};
#endif
