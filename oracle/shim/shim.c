#define _GNU_SOURCE
#include <stdio.h>
#include <stdarg.h>
#include <unistd.h>
#include <stdint.h>
/* bionic LP64 sizeof(FILE)==152; stdin/stdout/stderr are &__sF[0..2] */
char shim___sF[3 * 152] __attribute__((aligned(16)));
__asm__(".symver shim___sF, __sF@LIBC");
static FILE *mapf(void *f) {
  uintptr_t d = (uintptr_t)f - (uintptr_t)shim___sF;
  if (d < sizeof(shim___sF)) { int i = (int)(d / 152); return i == 0 ? stdin : (i == 1 ? stdout : stderr); }
  return (FILE *)f;
}
int shim_fprintf(void *f, const char *fmt, ...) { va_list ap; va_start(ap, fmt); int r = vfprintf(mapf(f), fmt, ap); va_end(ap); return r; }
__asm__(".symver shim_fprintf, fprintf@LIBC");
int shim_vfprintf(void *f, const char *fmt, va_list ap) { return vfprintf(mapf(f), fmt, ap); }
__asm__(".symver shim_vfprintf, vfprintf@LIBC");
int shim_fputc(int c, void *f) { return fputc(c, mapf(f)); }
__asm__(".symver shim_fputc, fputc@LIBC");
int shim_fflush(void *f) { return fflush(f ? mapf(f) : NULL); }
__asm__(".symver shim_fflush, fflush@LIBC");
size_t shim_fwrite(const void *p, size_t s, size_t n, void *f) { return fwrite(p, s, n, mapf(f)); }
__asm__(".symver shim_fwrite, fwrite@LIBC");
long shim_sysconf(int name) {
  switch (name) {  /* bionic _SC_* numbering */
    case 0x27: case 0x28: return sysconf(_SC_PAGESIZE);
    case 0x60: return sysconf(_SC_NPROCESSORS_CONF);
    case 0x61: return sysconf(_SC_NPROCESSORS_ONLN);
    default: return -1;
  }
}
__asm__(".symver shim_sysconf, sysconf@LIBC");
void shim_android_set_abort_message(const char *m) { if (m) fprintf(stderr, "abort message: %s\n", m); }
__asm__(".symver shim_android_set_abort_message, android_set_abort_message@LIBC");
int __android_log_write(int prio, const char *tag, const char *text) { fprintf(stderr, "[%d] %s: %s\n", prio, tag ? tag : "", text ? text : ""); return 0; }
int __android_log_print(int prio, const char *tag, const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); return 0; }
