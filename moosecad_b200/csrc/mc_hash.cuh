// Order- and numbering-independent mesh checksums (mc_mesh_checksum; numpy mirror in tests/canon.py).
//   vertex: H(x, y, z) over the bit patterns (-0.0 counts as +0.0)
//   tet:    the four vertex hashes, rotated by the EVEN permutation that makes the sequence
//           lexicographically smallest (orientation is preserved, the labelling is not), hashed
//   mesh:   sums mod 2^64 — independent of the order and of the vertex numbering, additive over slabs
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mc {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

__device__ __forceinline__ uint64_t coord_bits(double v) {
    if (v == 0.0) v = 0.0;  // -0.0 and +0.0 are the same position
    return (uint64_t)__double_as_longlong(v);
}

__device__ __forceinline__ uint64_t vertex_hash(double x, double y, double z) {
    uint64_t h = 0x9e3779b97f4a7c15ull;
    h = mix64(h + coord_bits(x));
    h = mix64(h + coord_bits(y));
    h = mix64(h + coord_bits(z));
    return h;
}

__device__ __forceinline__ uint64_t tet_hash(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    // the 12 even permutations of (0, 1, 2, 3)
    constexpr int P[12][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 0, 3, 2}, {1, 2, 0, 3}, {1, 3, 2, 0},
                              {2, 0, 1, 3}, {2, 1, 3, 0}, {2, 3, 0, 1}, {3, 0, 2, 1}, {3, 1, 0, 2}, {3, 2, 1, 0}};
    const uint64_t h[4] = {a, b, c, d};
    uint64_t best[4] = {a, b, c, d};
#pragma unroll
    for (int p = 1; p < 12; ++p) {
        const uint64_t s[4] = {h[P[p][0]], h[P[p][1]], h[P[p][2]], h[P[p][3]]};
        bool less = false, decided = false;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (!decided && s[k] != best[k]) { less = s[k] < best[k]; decided = true; }
        if (less) { best[0] = s[0]; best[1] = s[1]; best[2] = s[2]; best[3] = s[3]; }
    }
    uint64_t t = 0xd1b54a32d192ed03ull;
#pragma unroll
    for (int k = 0; k < 4; ++k) t = mix64(t + best[k]);
    return t;
}

}  // namespace mc
