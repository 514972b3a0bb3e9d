/*
 * oracle/path_redirect.c -- TEST INFRASTRUCTURE (LD_PRELOAD helper), not product code.
 *
 * The reference hard-codes absolute paths under /home/srujan_d/RISS/code/btrapz/src/
 * (trp_wrapper.cpp:23,288; cub_wrapper.cpp:22,268; trp_wrapper.py:45,100).  To run the
 * SHIPPED libtrp.so/libcub.so (and the recompiled oracle/_ref libraries) without
 * creating anything outside the repo, this interposer rewrites that prefix to the
 * directory named by $SPECTRAL_IO_DIR for fopen/fopen64/open/open64.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char kPrefix[] = "/home/srujan_d/RISS/code/btrapz/src/";

static const char *remap(const char *path, char *buf, size_t cap) {
  const char *dir = getenv("SPECTRAL_IO_DIR");
  if (!dir || !path || strncmp(path, kPrefix, sizeof(kPrefix) - 1) != 0) return path;
  snprintf(buf, cap, "%s/%s", dir, path + sizeof(kPrefix) - 1);
  return buf;
}

FILE *fopen(const char *path, const char *mode) {
  static FILE *(*real)(const char *, const char *);
  if (!real) real = (FILE * (*)(const char *, const char *)) dlsym(RTLD_NEXT, "fopen");
  char buf[4096];
  return real(remap(path, buf, sizeof buf), mode);
}

FILE *fopen64(const char *path, const char *mode) {
  static FILE *(*real)(const char *, const char *);
  if (!real) real = (FILE * (*)(const char *, const char *)) dlsym(RTLD_NEXT, "fopen64");
  char buf[4096];
  return real(remap(path, buf, sizeof buf), mode);
}

int open(const char *path, int flags, ...) {
  static int (*real)(const char *, int, ...);
  if (!real) real = (int (*)(const char *, int, ...))dlsym(RTLD_NEXT, "open");
  mode_t mode = 0;
  if (flags & (O_CREAT | O_TMPFILE)) {
    va_list ap;
    va_start(ap, flags);
    mode = va_arg(ap, mode_t);
    va_end(ap);
  }
  char buf[4096];
  return real(remap(path, buf, sizeof buf), flags, mode);
}

int open64(const char *path, int flags, ...) {
  static int (*real)(const char *, int, ...);
  if (!real) real = (int (*)(const char *, int, ...))dlsym(RTLD_NEXT, "open64");
  mode_t mode = 0;
  if (flags & (O_CREAT | O_TMPFILE)) {
    va_list ap;
    va_start(ap, flags);
    mode = va_arg(ap, mode_t);
    va_end(ap);
  }
  char buf[4096];
  return real(remap(path, buf, sizeof buf), flags, mode);
}
