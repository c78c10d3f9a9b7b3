/* config.h for compiling the liquid-dsp reference sources (as they lie under /root/reference)
 * into oracle/_ref/.  Written by hand for this repo: autoconf is not available, and the
 * reference's own build system is not run.  Internal FFT is forced (no fftw3.h in the image,
 * and LIQUID_FFTOVERRIDE for good measure); the x86 SSE dot products are selected by the
 * file list in oracle/Makefile, as configure.ac:156-167 would on x86-64. */
#ifndef __LIQUID_CONFIG_H__
#define __LIQUID_CONFIG_H__
#define HAVE_COMPLEX_H 1
#define HAVE_FLOAT_H 1
#define HAVE_GETOPT_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_LIMITS_H 1
#define HAVE_STDIO_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_SYS_RESOURCE_H 1
#define HAVE_UNISTD_H 1
#define HAVE_MMINTRIN_H 1
#define HAVE_XMMINTRIN_H 1
#define HAVE_EMMINTRIN_H 1
#define HAVE_PMMINTRIN_H 1
#define HAVE_TMMINTRIN_H 1
#define HAVE_SMMINTRIN_H 1
#define HAVE_IMMINTRIN_H 1
#define SIZEOF_INT 4
#define SIZEOF_UNSIGNED_INT 4
#define LIQUID_FFTOVERRIDE 1
#endif
