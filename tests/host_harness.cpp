// tests/host_harness.cpp -- compiles the product's device header
// (tcam_wsol_video_b200/csrc/lattice.cuh) for the HOST so that the CPU test
// suite can check the exact embedding / packed-key code the CUDA kernels run,
// against the oracle, without a GPU.  Test infrastructure only; built by
// tests/conftest.py with g++ -O2 -ffp-contract=off.
#include <cstdint>
#include <cstring>

#include "../tcam_wsol_video_b200/csrc/lattice.cuh"

using namespace tcamcrf;

template <int D>
static int embed_all(const float *feat, int n, const float *scale, int16_t *coords, float *bary,
                     unsigned long long *packed)
{
    EmbedConsts ec;
    for (int i = 0; i < kMaxD; i++) ec.scale[i] = i < D ? scale[i] : 0.f;
    int bad = 0;
    for (int p = 0; p < n; p++) {
        float f[D];
        for (int i = 0; i < D; i++) f[i] = feat[(size_t)p * D + i];
        int z[D + 1], rank[D + 1];
        float b[D + 1];
        if (!embed_point<D>(f, ec, z, rank, b)) bad++;
        unsigned long long keys[D + 1];
        KeyCodec<D>::pack_simplex(z, rank, keys);   // the routine build_kernel uses
        for (int r = 0; r <= D; r++) {
            int q[D];
            for (int i = 0; i < D; i++) q[i] = z[i] - ((rank[i] + r > D) ? 1 : 0);
            const unsigned long long key = keys[r];
            if (key != KeyCodec<D>::pack(q, r)) bad += 1000000;   // incremental packing must equal field packing
            packed[(size_t)p * (D + 1) + r] = key;
            for (int i = 0; i < D; i++)
                coords[((size_t)p * (D + 1) + r) * D + i] = (int16_t)KeyCodec<D>::coord(key, i);
            bary[(size_t)p * (D + 1) + r] = b[r];
        }
    }
    return bad;
}

template <int D>
static void neighbours_all(const unsigned long long *keys, int m, int16_t *n1, int16_t *n2)
{
    for (int j = 0; j <= D; j++)
        for (int v = 0; v < m; v++) {
            unsigned long long a, b;
            KeyCodec<D>::neighbour_keys(keys[v], j, a, b);
            for (int i = 0; i < D; i++) {
                n1[(((size_t)j * m) + v) * D + i] = (int16_t)KeyCodec<D>::coord(a, i);
                n2[(((size_t)j * m) + v) * D + i] = (int16_t)KeyCodec<D>::coord(b, i);
            }
        }
}

extern "C" {

// feat [n][D]; coords [n][D+1][D]; bary [n][D+1]; packed [n][D+1]. Returns #points out of key range, -1 on bad D.
int harness_embed(int D, const float *feat, int n, const float *scale, int16_t *coords, float *bary,
                  unsigned long long *packed)
{
    switch (D) {
    case 1: return embed_all<1>(feat, n, scale, coords, bary, packed);
    case 2: return embed_all<2>(feat, n, scale, coords, bary, packed);
    case 3: return embed_all<3>(feat, n, scale, coords, bary, packed);
    case 4: return embed_all<4>(feat, n, scale, coords, bary, packed);
    case 5: return embed_all<5>(feat, n, scale, coords, bary, packed);
    case 6: return embed_all<6>(feat, n, scale, coords, bary, packed);
    }
    return -1;
}

// keys [m]; n1/n2 [D+1][m][D] decoded coordinates of the two neighbours along each axis.
int harness_neighbours(int D, const unsigned long long *keys, int m, int16_t *n1, int16_t *n2)
{
    switch (D) {
    case 1: neighbours_all<1>(keys, m, n1, n2); return 0;
    case 2: neighbours_all<2>(keys, m, n1, n2); return 0;
    case 3: neighbours_all<3>(keys, m, n1, n2); return 0;
    case 4: neighbours_all<4>(keys, m, n1, n2); return 0;
    case 5: neighbours_all<5>(keys, m, n1, n2); return 0;
    case 6: neighbours_all<6>(keys, m, n1, n2); return 0;
    }
    return -1;
}

int harness_field_bits(int D)
{
    switch (D) {
    case 1: return KeyCodec<1>::kFieldBits;
    case 2: return KeyCodec<2>::kFieldBits;
    case 3: return KeyCodec<3>::kFieldBits;
    case 4: return KeyCodec<4>::kFieldBits;
    case 5: return KeyCodec<5>::kFieldBits;
    case 6: return KeyCodec<6>::kFieldBits;
    }
    return -1;
}
}
