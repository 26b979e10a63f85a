// TEST INFRASTRUCTURE (oracle) — never linked into the product.
// Implements the handful of JNI / NDK entry points the reference's JNI translation units call, so that
// the UNMODIFIED reference code path  Java_com_awxkee_jxlcoder_JxlCoder_decodeSampledImpl → decodeSampledImageImpl
// (/root/reference/jxlcoder/src/main/cpp/JniDecoding.cpp:45-359) runs on a glibc host and hands back a "Bitmap"
// that is just a heap buffer.
#include "fakejni.h"

#include <cstdio>
#include <cstring>

static int g_api_level = 34;
extern "C" int android_get_device_api_level() { return g_api_level; }
extern "C" void fakejni_set_api_level(int v) { g_api_level = v; }

static FakeEnvState *st(JNIEnv *e) { return reinterpret_cast<FakeEnvState *>(e->impl); }

FakeObject *FakeEnvState::make(FakeObject::Kind k, const std::string &name) {
  auto *o = new FakeObject();
  o->kind = k;
  o->name = name;
  objects.push_back(o);
  return o;
}
FakeMember *FakeEnvState::member(const std::string &cls, const std::string &name, const std::string &sig) {
  auto *m = new FakeMember{cls, name, sig};
  members.push_back(m);
  return m;
}
FakeEnvState::~FakeEnvState() {
  for (auto *o : objects) delete o;
  for (auto *m : members) delete m;
}

jclass JNIEnv::FindClass(const char *name) { return st(this)->make(FakeObject::Class, name); }
jint JNIEnv::ThrowNew(jclass cls, const char *msg) {
  st(this)->exception_class = cls ? cls->name : "";
  st(this)->exception_msg = msg ? msg : "";
  st(this)->has_exception = true;
  return 0;
}
jsize JNIEnv::GetArrayLength(jbyteArray arr) { return (jsize) arr->len; }
void JNIEnv::GetByteArrayRegion(jbyteArray arr, jsize start, jsize len, jbyte *buf) {
  memcpy(buf, arr->data + start, (size_t) len);
}
void *JNIEnv::GetDirectBufferAddress(jobject buf) {
  return buf->kind == FakeObject::DirectBuffer ? const_cast<uint8_t *>(buf->data) : nullptr;
}
jlong JNIEnv::GetDirectBufferCapacity(jobject buf) {
  return buf->kind == FakeObject::DirectBuffer ? (jlong) buf->len : -1;
}
jmethodID JNIEnv::GetMethodID(jclass cls, const char *name, const char *sig) {
  return st(this)->member(cls->name, name, sig);
}
jmethodID JNIEnv::GetStaticMethodID(jclass cls, const char *name, const char *sig) {
  return st(this)->member(cls->name, name, sig);
}
jfieldID JNIEnv::GetStaticFieldID(jclass cls, const char *name, const char *sig) {
  return st(this)->member(cls->name, name, sig);
}
jobject JNIEnv::GetStaticObjectField(jclass cls, jfieldID f) {
  // Bitmap$Config.<name> and ColorSpace$Named.<name> are the only static fields the reference reads.
  return st(this)->make(FakeObject::EnumConst, f->name);
}

static int bytes_per_pixel(const std::string &config) {
  if (config == "RGBA_F16") return 8;
  if (config == "RGB_565") return 2;
  return 4;  // ARGB_8888, RGBA_1010102
}

jobject JNIEnv::CallStaticObjectMethod(jclass cls, jmethodID m, ...) {
  va_list ap;
  va_start(ap, m);
  jobject result = nullptr;
  if (cls->name == "android/graphics/ColorSpace" && m->name == "get") {
    jobject named = va_arg(ap, jobject);
    result = st(this)->make(FakeObject::ColorSpace, named ? named->name : "");
  } else if (cls->name == "android/graphics/Bitmap" && m->name == "createBitmap") {
    int w = va_arg(ap, int);
    int h = va_arg(ap, int);
    jobject cfg = va_arg(ap, jobject);
    jobject cs = nullptr;
    if (m->sig.find("ZLandroid/graphics/ColorSpace;") != std::string::npos) {
      (void) va_arg(ap, int);  // hasAlpha (bool promoted to int)
      cs = va_arg(ap, jobject);
    }
    auto *b = st(this)->make(FakeObject::Bitmap, "Bitmap");
    b->width = w;
    b->height = h;
    b->config = cfg ? cfg->name : "";
    b->colorspace = cs ? cs->name : "";
    b->stride = w * bytes_per_pixel(b->config);
    b->pixels.assign((size_t) b->stride * (size_t) h, 0);
    result = b;
  } else {
    // wrapHardwareBuffer etc.: not available off-device.
    result = nullptr;
  }
  va_end(ap);
  return result;
}

jobject JNIEnv::NewObject(jclass cls, jmethodID m, ...) {
  va_list ap;
  va_start(ap, m);
  jobject result = nullptr;
  if (cls->name == "android/util/Size") {
    auto *s = st(this)->make(FakeObject::Size, "Size");
    s->width = va_arg(ap, int);
    s->height = va_arg(ap, int);
    result = s;
  }
  va_end(ap);
  return result;
}

int AndroidBitmap_getInfo(JNIEnv *, jobject b, AndroidBitmapInfo *info) {
  if (!b || b->kind != FakeObject::Bitmap) return -1;
  info->width = (uint32_t) b->width;
  info->height = (uint32_t) b->height;
  info->stride = (uint32_t) b->stride;
  info->format = 0;
  info->flags = 0;
  return 0;
}
int AndroidBitmap_lockPixels(JNIEnv *, jobject b, void **addr) {
  if (!b || b->kind != FakeObject::Bitmap) return -1;
  *addr = b->pixels.data();
  return 0;
}
int AndroidBitmap_unlockPixels(JNIEnv *, jobject) { return 0; }

extern "C" int __android_log_print(int, const char *tag, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[%s] ", tag ? tag : "");
  vfprintf(stderr, fmt, ap);
  fputc('\n', stderr);
  va_end(ap);
  return 0;
}

// HardwareBuffersCompat.cpp is Android-only (dlopen of libandroid.so); the HARDWARE colour config is out of
// scope.  Provide its symbols so ReformatBitmap.cpp links; the load reports failure,
// which the reference turns into "Cannot load hardware buffers API".
#include <HardwareBuffersCompat.h>
bool loadAHardwareBuffersAPI() { return false; }
AHardwareBufferAllocateFunc AHardwareBuffer_allocate_compat = nullptr;
AHardwareBufferIsSupportedFunc AHardwareBuffer_isSupported_compat = nullptr;
AHardwareBufferUnlockFunc AHardwareBuffer_unlock_compat = nullptr;
AHardwareBufferReleaseFunc AHardwareBuffer_release_compat = nullptr;
AHardwareBufferLockFunc AHardwareBuffer_lock_compat = nullptr;
AHardwareBufferToHardwareBufferFunc AHardwareBuffer_toHardwareBuffer_compat = nullptr;
AHardwareBufferDescribeFunc AHardwareBuffer_describe_compat = nullptr;
