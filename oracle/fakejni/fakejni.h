// TEST INFRASTRUCTURE (oracle): object model behind the fake JNI (see jni.h).
#pragma once
#include <jni.h>
#include <android/bitmap.h>
#include <string>
#include <vector>

struct FakeObject {
  enum Kind { Class, ByteArray, DirectBuffer, HeapBuffer, EnumConst, ColorSpace, Bitmap, Size } kind = Class;
  std::string name;
  const uint8_t *data = nullptr;  // ByteArray / buffers (borrowed)
  size_t len = 0;
  // Bitmap / Size
  int width = 0, height = 0, stride = 0;
  std::string config;      // Bitmap.Config name
  std::string colorspace;  // ColorSpace.Named name ("" = none passed)
  std::vector<uint8_t> pixels;
};

struct FakeMember {
  std::string cls, name, sig;
};

struct FakeEnvState {
  std::vector<FakeObject *> objects;
  std::vector<FakeMember *> members;
  bool has_exception = false;
  std::string exception_class, exception_msg;
  FakeObject *make(FakeObject::Kind k, const std::string &name);
  FakeMember *member(const std::string &cls, const std::string &name, const std::string &sig);
  ~FakeEnvState();
};

extern "C" void fakejni_set_api_level(int v);
