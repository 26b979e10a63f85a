// TEST INFRASTRUCTURE (oracle): a minimal stand-in for the Android NDK <jni.h>, just large enough
// that the reference's UNMODIFIED JNI translation units (JniDecoding.cpp, ReformatBitmap.cpp,
// SizeScaler.cpp, Support.cpp, JxlAnimatedDecoderCoordinator.cpp, NativeColorSpace.cpp,
// JniExceptions.cpp under /root/reference/jxlcoder/src/main/cpp) compile and run on a glibc host.
// Objects are plain C++ structs; "Bitmap" is a heap buffer (see fakejni.cpp).  Never shipped.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstddef>

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef int32_t jsize;
typedef float jfloat;
typedef double jdouble;

struct FakeObject;  // defined in fakejni.h
typedef FakeObject *jobject;
typedef jobject jclass;
typedef jobject jbyteArray;
typedef jobject jstring;
typedef jobject jthrowable;
struct FakeMember;
typedef FakeMember *jmethodID;
typedef FakeMember *jfieldID;

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_TRUE 1
#define JNI_FALSE 0

extern "C" int android_get_device_api_level();

struct JNIEnv {
  jclass FindClass(const char *name);
  jint ThrowNew(jclass cls, const char *msg);
  jsize GetArrayLength(jbyteArray arr);
  void GetByteArrayRegion(jbyteArray arr, jsize start, jsize len, jbyte *buf);
  void *GetDirectBufferAddress(jobject buf);
  jlong GetDirectBufferCapacity(jobject buf);
  jmethodID GetMethodID(jclass cls, const char *name, const char *sig);
  jmethodID GetStaticMethodID(jclass cls, const char *name, const char *sig);
  jfieldID GetStaticFieldID(jclass cls, const char *name, const char *sig);
  jobject GetStaticObjectField(jclass cls, jfieldID f);
  jobject CallStaticObjectMethod(jclass cls, jmethodID m, ...);
  jobject NewObject(jclass cls, jmethodID m, ...);
  // state
  void *impl;
};
