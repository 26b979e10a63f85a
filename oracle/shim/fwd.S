.text
.macro FWD name
.globl shim_\name
.type shim_\name,@function
shim_\name:
  jmp \name@PLT
.symver shim_\name, \name@LIBC
.endm
.macro FWD2 name target
.globl shim_\name
.type shim_\name,@function
shim_\name:
  jmp \target@PLT
.symver shim_\name, \name@LIBC
.endm
FWD __cxa_atexit
FWD __cxa_finalize
FWD __memcpy_chk
FWD __memset_chk
FWD __stack_chk_fail
FWD abort
FWD atan2
FWD cbrtf
FWD closelog
FWD cos
FWD dl_iterate_phdr
FWD exp
FWD exit
FWD fmod
FWD free
FWD hypot
FWD hypotf
FWD ldexp
FWD ldexpf
FWD llroundf
FWD log
FWD log1p
FWD log1pf
FWD log2
FWD log2f
FWD logf
FWD lroundf
FWD malloc
FWD memchr
FWD memcmp
FWD memcpy
FWD memmove
FWD memset
FWD modff
FWD openlog
FWD posix_memalign
FWD pow
FWD powf
FWD pthread_cond_broadcast
FWD pthread_cond_destroy
FWD pthread_cond_signal
FWD pthread_cond_wait
FWD pthread_create
FWD pthread_getspecific
FWD pthread_join
FWD pthread_key_create
FWD pthread_key_delete
FWD pthread_mutex_destroy
FWD pthread_mutex_lock
FWD pthread_mutex_unlock
FWD pthread_once
FWD pthread_rwlock_rdlock
FWD pthread_rwlock_unlock
FWD pthread_rwlock_wrlock
FWD pthread_setspecific
FWD realloc
FWD remainder
FWD sin
FWD snprintf
FWD sqrt
FWD sqrtf
FWD strcmp
FWD strlen
FWD syscall
FWD syslog
FWD vasprintf
FWD vsnprintf
FWD wmemchr
FWD2 __errno __errno_location
FWD2 strerror_r __xpg_strerror_r
.section .note.GNU-stack,"",@progbits
