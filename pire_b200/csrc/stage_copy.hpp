// stage_copy.hpp -- the copy that stages pageable input into the library's pinned slots (capi_host.cu).
// Header-only so that tests/cpp/stage_copy_check.cpp can run it on the CPU against memcpy.
#pragma once

#include <cstdint>
#include <cstring>
#if defined(__x86_64__) && defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace pire_b200 {

// One slice of a staging copy.  The destination is a pinned slot that the GPU reads by DMA and no CPU reads again, so
// on x86 the bytes are written with non-temporal stores: no read-for-ownership of the destination lines (a third of the
// memory traffic of an ordinary copy) and the caller's cache keeps its contents.
inline void StageCopy(uint8_t* dst, const uint8_t* src, size_t bytes)
{
#if defined(__x86_64__) && defined(__SSE2__)
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
    if (bytes >= 256 + head) {
        std::memcpy(dst, src, head);
        dst += head;
        src += head;
        bytes -= head;
        const size_t blocks = bytes / 64;
        for (size_t i = 0; i < blocks; ++i) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 48), d);
            src += 64;
            dst += 64;
        }
        _mm_sfence();                       // the stores are globally visible before the slice is reported done
        bytes -= blocks * 64;
    }
#endif
    std::memcpy(dst, src, bytes);
}

} // namespace pire_b200
