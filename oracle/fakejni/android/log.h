// TEST INFRASTRUCTURE (oracle): stand-in for the NDK <android/log.h>.
#pragma once
#define ANDROID_LOG_ERROR 6
extern "C" int __android_log_print(int prio, const char *tag, const char *fmt, ...);
