/* Local config.h for building the reference subset as the CPU oracle.
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md). Mirrors what spandsp's
 * configure would define on x86-64/glibc (configure.ac:276-510). */
#define HAVE_MATH_H 1
#define HAVE_TGMATH_H 1
#define HAVE_STDBOOL_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_STDINT_H 1
#define HAVE_SINF 1
#define HAVE_COSF 1
#define HAVE_TANF 1
#define HAVE_ASINF 1
#define HAVE_ACOSF 1
#define HAVE_ATANF 1
#define HAVE_ATAN2F 1
#define HAVE_CEILF 1
#define HAVE_FLOORF 1
#define HAVE_POWF 1
#define HAVE_EXPF 1
#define HAVE_LOGF 1
#define HAVE_LOG10F 1
#define HAVE_LONG_DOUBLE 1
#define HAVE_ALIGNED_ALLOC 1
#define HAVE_POSIX_MEMALIGN 1
#define HAVE_OPEN_MEMSTREAM 1
#define HAVE_DRAND48 1
#define SPANDSP_USE_EXPORT_CAPABILITY 1
#if defined(ORACLE_REF_FLAGS)
/* the reference's x86-64 default (configure.ac:276,509-510) */
#define SPANDSP_USE_SSE2 1
#endif
