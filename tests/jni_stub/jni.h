// Minimal stand-in for <jni.h> (no JDK in the development image): just enough declarations to syntax-check csrc/jni_shim.cpp.
#pragma once
#include <cstdint>
#define JNIEXPORT
#define JNICALL
#define JNI_ABORT 2
typedef int32_t jint; typedef int64_t jlong; typedef int8_t jbyte; typedef uint8_t jboolean; typedef jint jsize;
struct _jobject {}; typedef _jobject *jobject; typedef jobject jclass; typedef jobject jstring; typedef jobject jarray;
typedef jarray jlongArray; typedef jarray jintArray; typedef jarray jbyteArray; typedef jarray jbooleanArray;
struct JNIEnv {
  jclass FindClass(const char*); jint ThrowNew(jclass, const char*);
  const char* GetStringUTFChars(jstring, jboolean*); void ReleaseStringUTFChars(jstring, const char*);
  jstring NewStringUTF(const char*);
  jlongArray NewLongArray(jsize); void SetLongArrayRegion(jlongArray, jsize, jsize, const jlong*);
  jintArray NewIntArray(jsize); void SetIntArrayRegion(jintArray, jsize, jsize, const jint*);
  jsize GetArrayLength(jarray);
  jlong* GetLongArrayElements(jlongArray, jboolean*); void ReleaseLongArrayElements(jlongArray, jlong*, jint);
  jint* GetIntArrayElements(jintArray, jboolean*); void ReleaseIntArrayElements(jintArray, jint*, jint);
  void SetByteArrayRegion(jbyteArray, jsize, jsize, const jbyte*);
  jbyte* GetByteArrayElements(jbyteArray, jboolean*); void ReleaseByteArrayElements(jbyteArray, jbyte*, jint);
  jboolean* GetBooleanArrayElements(jbooleanArray, jboolean*); void ReleaseBooleanArrayElements(jbooleanArray, jboolean*, jint);
};
