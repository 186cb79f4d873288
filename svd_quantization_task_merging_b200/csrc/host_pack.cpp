// Inner loop of svdq_host_pack_mask(_batch): torch.bool bytes -> bits, output bytes [lo, hi) of one mask.
// AVX2 path (runtime-dispatched): compare 32 bytes against zero and take the byte sign bits with one movemask; the
// portable path folds 8 bytes into one with a multiply.  Any non-zero byte counts as set.
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

void pack_portable(const uint8_t* src, int64_t n, uint8_t* dst, int64_t lo, int64_t hi) {
    for (int64_t k = lo; k < hi; ++k) {
        const int64_t e = k * 8;
        uint64_t x = 0;
        if (e + 8 <= n) std::memcpy(&x, src + e, 8);
        else std::memcpy(&x, src + e, (size_t)(n - e));
        x |= x >> 4; x |= x >> 2; x |= x >> 1;
        x &= 0x0101010101010101ull;
        dst[k] = (uint8_t)((x * 0x0102040810204080ull) >> 56);
    }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void pack_avx2(const uint8_t* src, int64_t n, uint8_t* dst, int64_t lo, int64_t hi) {
    int64_t k = lo;
    const __m256i zero = _mm256_setzero_si256();
    for (; k + 4 <= hi && (k + 4) * 8 <= n; k += 4) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + k * 8));
        const uint32_t m = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, zero));
        std::memcpy(dst + k, &m, 4);
    }
    if (k < hi) pack_portable(src, n, dst, k, hi);
}
#endif

}  // namespace

extern "C" void svdq_host_pack_range(const uint8_t* src, int64_t n, uint8_t* dst, int64_t lo, int64_t hi) {
#if defined(__x86_64__)
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2) { pack_avx2(src, n, dst, lo, hi); return; }
#endif
    pack_portable(src, n, dst, lo, hi);
}
