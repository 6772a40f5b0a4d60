// Host-side streaming copy for the staging of pageable caller buffers (handle.cuh: CopyTeam).
//
// glibc's memcpy switches to non-temporal stores only above a threshold derived from the shared cache size, far above the
// few MB a staging thread copies at a time; with ordinary stores every destination line is first read (write allocate), so
// a copy moves three bytes through the memory controllers per byte copied - beside the DMA engine that reads the pinned
// ring at the link rate.  Non-temporal stores make it two.
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>

__attribute__((target("avx2"))) static void stream_copy_avx2(uint8_t *d, const uint8_t *s, size_t n) {
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(d) & 31)) & 31;
    if (head) { const size_t h = head < n ? head : n; memcpy(d, s, h); d += h; s += h; n -= h; }
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        _mm_prefetch(reinterpret_cast<const char *>(s + i + 1024), _MM_HINT_NTA);
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 64));
        const __m256i e = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(s + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i *>(d + i + 96), e);
    }
    _mm_sfence();
    if (i < n) memcpy(d + i, s + i, n - i);
}
#endif

extern "C" void rl_stream_copy(void *dst, const void *src, size_t n) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096 && getenv("RLSIM_NO_STREAM_COPY") == nullptr) { stream_copy_avx2(static_cast<uint8_t *>(dst), static_cast<const uint8_t *>(src), n); return; }
#endif
    memcpy(dst, src, n);
}
