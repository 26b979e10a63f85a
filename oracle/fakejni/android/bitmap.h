// TEST INFRASTRUCTURE (oracle): stand-in for the NDK <android/bitmap.h>; see ../jni.h.
#pragma once
#include <jni.h>
#ifndef _Nullable
#define _Nullable
#endif
#ifndef _Nonnull
#define _Nonnull
#endif
typedef struct {
  uint32_t width;
  uint32_t height;
  uint32_t stride;
  int32_t format;
  uint32_t flags;
} AndroidBitmapInfo;
int AndroidBitmap_getInfo(JNIEnv *env, jobject jbitmap, AndroidBitmapInfo *info);
int AndroidBitmap_lockPixels(JNIEnv *env, jobject jbitmap, void **addrPtr);
int AndroidBitmap_unlockPixels(JNIEnv *env, jobject jbitmap);
extern "C" int android_get_device_api_level();
