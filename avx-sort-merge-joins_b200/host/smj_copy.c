/* smj_copy.c -- host-side copy loops of the pageable staging pipeline (csrc/smj_runtime.cu).
 *
 * The staging threads copy 4 MiB chunks of the caller's (pageable) relation into pinned bounce
 * buffers that the DMA engine reads right afterwards.  glibc's memcpy switches to non-temporal
 * stores above a size threshold; whether the bounce data should bypass the cache (the DMA engine
 * then reads it from DRAM) or stay in it (the ring of 16 x 4 MiB about fits the last-level cache) is
 * a property of the host, so both loops exist and $SMJ_STAGE_COPY picks one (see smj_stage_copy).
 */
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

__attribute__((target("avx2"))) static void copy_cached_avx2(char *dst, const char *src, size_t n) {
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        _mm_prefetch(src + i + 1024, _MM_HINT_NTA);
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64));
        const __m256i d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
        _mm256_storeu_si256((__m256i *)(dst + i), a);
        _mm256_storeu_si256((__m256i *)(dst + i + 32), b);
        _mm256_storeu_si256((__m256i *)(dst + i + 64), c);
        _mm256_storeu_si256((__m256i *)(dst + i + 96), d);
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
}

__attribute__((target("avx2"))) static void copy_stream_avx2(char *dst, const char *src, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) {
        dst[i] = src[i];
        i++;
    }
    for (; i + 128 <= n; i += 128) {
        _mm_prefetch(src + i + 1024, _MM_HINT_NTA);
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 64));
        const __m256i d = _mm256_loadu_si256((const __m256i *)(src + i + 96));
        _mm256_stream_si256((__m256i *)(dst + i), a);
        _mm256_stream_si256((__m256i *)(dst + i + 32), b);
        _mm256_stream_si256((__m256i *)(dst + i + 64), c);
        _mm256_stream_si256((__m256i *)(dst + i + 96), d);
    }
    _mm_sfence();
    if (i < n) memcpy(dst + i, src + i, n - i);
}

/* mode: 0 = memcpy, 1 = cached AVX2 stores, 2 = non-temporal AVX2 stores ($SMJ_STAGE_COPY = memcpy |
 * cached | stream, read once) */
void smj_stage_copy(void *dst, const void *src, size_t n) {
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("SMJ_STAGE_COPY");
        int m = 0;
        if (e && e[0] == 'c') m = 1;
        if (e && e[0] == 's') m = 2;
        if (m && !__builtin_cpu_supports("avx2")) m = 0;
        mode = m;
    }
    if (mode == 1)
        copy_cached_avx2((char *)dst, (const char *)src, n);
    else if (mode == 2)
        copy_stream_avx2((char *)dst, (const char *)src, n);
    else
        memcpy(dst, src, n);
}
