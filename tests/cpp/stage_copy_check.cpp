// StageCopy (pire_b200/csrc/stage_copy.hpp) against memcpy: every alignment of source and destination, lengths around
// the 64-byte blocks of the non-temporal loop and around its threshold, guard bytes behind the destination.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "stage_copy.hpp"

int main()
{
    std::vector<uint8_t> src(1 << 22), dst(1 << 22);
    for (size_t i = 0; i < src.size(); ++i)
        src[i] = (uint8_t) ((i * 2654435761u) >> 13);
    long bad = 0, cases = 0;
    auto one = [&](size_t so, size_t dof, size_t n) {
        std::memset(&dst[dof ? dof - 1 : 0], 0xEE, n + 66);
        pire_b200::StageCopy(&dst[dof], &src[so], n);
        bad += std::memcmp(&dst[dof], &src[so], n) != 0;
        bad += dst[dof + n] != 0xEE || dst[dof + n + 63] != 0xEE;          // nothing written behind the slice
        if (dof)
            bad += dst[dof - 1] != 0xEE;                                     // nor in front of it
        ++cases;
    };
    for (size_t so = 0; so < 33; ++so)
        for (size_t dof = 0; dof < 33; ++dof)
            for (size_t n : {0u, 1u, 15u, 16u, 63u, 64u, 255u, 256u, 257u, 270u, 271u, 272u, 319u, 320u, 321u, 1000u, 4096u, 4099u})
                one(so, dof, n);
    srand(7);
    for (int it = 0; it < 300; ++it)
        one(rand() % 4096, rand() % 4096, (2u << 20) - 100 + rand() % 200);     // the pool's slices are 2 MiB
    std::printf("%ld cases, %ld bad\n", cases, bad);
    return bad != 0;
}
