/*
 * TEST INFRASTRUCTURE ONLY (oracle build).  Declaration-only stand-in for <zstd.h>.
 *
 * The reference's umbrella header include/albatross/Common:18 includes <zstd.h>, which is not
 * installed in this image.  zstd is only used by the reference's serialization code
 * (include/albatross/src/utils/compress.hpp:24-132), which the oracle never calls, so these
 * declarations are never linked.
 */
#ifndef AB_ORACLE_ZSTD_STUB_H
#define AB_ORACLE_ZSTD_STUB_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
#define ZSTD_CLEVEL_DEFAULT 3
#define ZSTD_CONTENTSIZE_UNKNOWN (0ULL - 1)
#define ZSTD_CONTENTSIZE_ERROR (0ULL - 2)
size_t ZSTD_compressBound(size_t srcSize);
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src, size_t srcSize,
                     int compressionLevel);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
unsigned long long ZSTD_getFrameContentSize(const void *src, size_t srcSize);
unsigned ZSTD_isError(size_t code);
const char *ZSTD_getErrorName(size_t code);
#ifdef __cplusplus
}
#endif
#endif
