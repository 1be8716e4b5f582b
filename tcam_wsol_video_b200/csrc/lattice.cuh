// lattice.cuh -- device-side building blocks of the permutohedral lattice:
// packed 64-bit vertex keys, the per-pixel embedding, and the open-addressing
// key table.  sm_100a only.
//
// What it replaces in the reference (dlib/crf/crfwrapper/bilateralfilter/):
//   embed_point<D>   <- the per-pixel body of Permutohedral::init, SSE path
//                       (permutohedral.cpp:168-252)
//   KeyCodec<D>      <- the `short key[d]` arrays + canonical simplex table
//                       (permutohedral.cpp:144-153,245-248)
//   table_insert /   <- HashTable::find(k, create=true/false)
//   table_lookup        (permutohedral.cpp:67-95)
//
// Design notes (B200-first, not a translation):
//   * A lattice vertex has d coordinates that are all congruent to the same
//     remainder r (mod d+1).  Instead of d int16 values the key stores the d
//     quotients q_i = floor(k_i/(d+1)) in B-bit biased fields plus r in 3 bits,
//     which fits one 64-bit word for every d <= 6.  One 64-bit atomicCAS then
//     inserts a vertex, and the +-1 / -+d neighbour steps of the blur become
//     integer adds on the packed word (neighbour_keys()).
//   * fp32 operations whose rounding decides which simplex a pixel falls in are
//     written with explicit round-to-nearest intrinsics so nvcc cannot contract
//     them into FMAs: the reference's SSE2 code rounds after every multiply.
#pragma once
#include <cstdint>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#else
// Host emulation, used only by tests/host_harness.cpp to run the very same
// embedding and key arithmetic on the CPU (compile with -ffp-contract=off).
#include <cmath>
#define __device__
#define __host__
#define __forceinline__ inline
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __float2int_rn(float v) { return (int)lrintf(v); }
static inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((unsigned long long)a * b) >> 32); }
#endif

namespace tcamcrf {

constexpr int kMaxD = 6;
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

template <int D>
struct KeyCodec {
    static_assert(D >= 1 && D <= kMaxD, "lattice dimension out of range");
    static constexpr int kRemBits = 3;
    static constexpr int kFieldBits = ((64 - kRemBits) / D) < 20 ? ((64 - kRemBits) / D) : 20;
    static constexpr int kBias = 1 << (kFieldBits - 1);
    static constexpr unsigned long long kFieldMask = (1ull << kFieldBits) - 1ull;
    // smallest / largest quotient that can still take a +-1 neighbour step
    static constexpr int kQMin = -kBias + 1;
    static constexpr int kQMax = kBias - 2;

    // 1 in every quotient field
    __host__ __device__ static constexpr unsigned long long ones()
    {
        unsigned long long v = 0;
        for (int i = 0; i < D; i++) v |= 1ull << (kRemBits + i * kFieldBits);
        return v;
    }
    __host__ __device__ static constexpr unsigned long long unit(int i)
    {
        return 1ull << (kRemBits + i * kFieldBits);
    }
    __device__ static unsigned long long pack(const int (&q)[D], int rem)
    {
        unsigned long long k = (unsigned long long)rem;
#pragma unroll
        for (int i = 0; i < D; i++)
            k |= ((unsigned long long)(unsigned)(q[i] + kBias) & kFieldMask) << (kRemBits + i * kFieldBits);
        return k;
    }
    // The d+1 vertices of the simplex of a point: remainder r has quotients z[i] - [rank[i] + r > D]
    // (the canonical simplex, permutohedral.cpp:148-153,247).  Coordinate i crosses at r = D + 1 - rank[i] and the
    // ranks are a permutation of 0..D, so from one remainder to the next exactly one coordinate steps down:
    //   key[r] = key[r-1] + 1 - unit(i with rank[i] == D + 1 - r)      (nothing to subtract when that i is D,
    // the implicit coordinate).  The inverse permutation is kept as nibbles of one word.
    __device__ static void pack_simplex(const int (&z)[D + 1], const int (&rank)[D + 1],
                                        unsigned long long (&key)[D + 1])
    {
        int q0[D];
#pragma unroll
        for (int i = 0; i < D; i++) q0[i] = z[i];
        key[0] = pack(q0, 0);   // rank[i] + 0 > D never holds
        unsigned int inv = 0;   // nibble t = the coordinate whose rank is t
#pragma unroll
        for (int i = 0; i <= D; i++) inv |= (unsigned int)i << (4 * rank[i]);
#pragma unroll
        for (int r = 1; r <= D; r++) {
            const unsigned int i = (inv >> (4 * (D + 1 - r))) & 15u;
            const unsigned long long u = i < (unsigned int)D ? (1ull << (kRemBits + i * kFieldBits)) : 0ull;
            key[r] = key[r - 1] + 1ull - u;
        }
    }
    // lattice coordinate i of a packed key (for debugging / tests)
    __host__ __device__ static int coord(unsigned long long key, int i)
    {
        int rem = (int)(key & 7ull);
        int q = (int)((key >> (kRemBits + i * kFieldBits)) & kFieldMask) - kBias;
        return q * (D + 1) + rem;
    }

    // Keys of the two blur neighbours along axis j (0..D) of vertex `key`:
    //   n1 = key - 1 on every coordinate, + (D+1) on coordinate j
    //   n2 = key + 1 on every coordinate, - (D+1) on coordinate j
    // (permutohedral.cpp:285-290; axis D is the implicit coordinate, so only the
    // all-coordinates step remains).  In (q, rem) form a -1 step lowers rem, and
    // wraps every quotient down when rem was 0; the +(D+1) on axis j is q_j + 1.
    __device__ static void neighbour_keys(unsigned long long key, int j, unsigned long long &n1,
                                          unsigned long long &n2)
    {
        const int rem = (int)(key & 7ull);
        const unsigned long long uj = (j < D) ? unit(j) : 0ull;
        if (rem > 0)
            n1 = key - 1ull + uj;              // rem-1, q_j+1
        else
            n1 = key + (unsigned long long)D - ones() + uj;  // rem=D, all q-1, then q_j back up
        if (rem < D)
            n2 = key + 1ull - uj;              // rem+1, q_j-1
        else
            n2 = key - (unsigned long long)D + ones() - uj;  // rem=0, all q+1, then q_j back down
    }
};

// Multiply-shift (Fibonacci) hashing: slot = top log2(slots) bits of key * odd constant.  The packed keys are
// sums of small integers in fixed bit fields; on them this spreads better than the murmur3 finaliser used before
// (sequential linear-probing inserts of one 224x224 frame into 262 144 slots: 1.058 probes per key against 1.135
// on iid-noise frames, 1.000 against 1.008 on natural frames) and costs a third of the instructions: only the
// upper 32 bits of the 64-bit product are formed.  `shift` = 32 - log2(slots), slots <= 2^30.
__device__ __forceinline__ unsigned int hash_slot(unsigned long long k, unsigned long long c, unsigned int shift)
{
    const unsigned int klo = (unsigned int)k, khi = (unsigned int)(k >> 32);
    const unsigned int clo = (unsigned int)c, chi = (unsigned int)(c >> 32);
    const unsigned int upper = __umulhi(klo, clo) + klo * chi + khi * clo;   // bits 32..63 of k * c
    return upper >> shift;
}
constexpr unsigned long long kHashMul1 = 0x9E3779B97F4A7C15ull;   // 2^64 / golden ratio: primary tier
constexpr unsigned long long kHashMul2 = 0xC2B2AE3D27D4EB4Full;   // overflow tier (independent of the first)

struct EmbedConsts {
    float scale[kMaxD];  // diagonal of E, double-evaluated on the host (permutohedral.cpp:156-159)
};

// Embeds one feature vector.  Outputs: z[i] (rem0[i] = z[i]*(D+1)), rank[i], bary[r].
// Returns false when a quotient leaves the packed-key range.
template <int D>
__device__ __forceinline__ bool embed_point(const float (&f)[D], const EmbedConsts &ec, int (&z)[D + 1],
                                            int (&rank)[D + 1], float (&bary)[D + 1])
{
    constexpr float inv_dp1 = 1.0f / (D + 1);
    constexpr float dp1 = (float)(D + 1);
    float el[D + 1];

    // y = E p (permutohedral.cpp:177-184)
    float sm = 0.0f;
#pragma unroll
    for (int j = D; j > 0; j--) {
        const float cf = __fmul_rn(f[j - 1], ec.scale[j - 1]);
        el[j] = __fsub_rn(sm, __fmul_rn((float)j, cf));
        sm = __fadd_rn(sm, cf);
    }
    el[0] = sm;

    // nearest 0-coloured vertex; cvtps2dq rounds half to even (:187-197)
    int sum = 0;
    float rem0[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++) {
        const int vi = __float2int_rn(__fmul_rn(inv_dp1, el[i]));
        z[i] = vi;
        rem0[i] = __fmul_rn((float)vi, dp1);
        sum += vi;
    }

    // rank of the residuals, fp32 compares (:200-210)
    float res[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++) {
        res[i] = __fsub_rn(el[i], rem0[i]);
        rank[i] = 0;
    }
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = i + 1; j <= D; j++) {
            const int c = res[i] < res[j] ? 1 : 0;
            rank[i] += c;
            rank[j] += 1 - c;
        }

    // back onto the plane sum == 0 (:213-219)
    bool in_range = true;
#pragma unroll
    for (int i = 0; i <= D; i++) {
        rank[i] += sum;
        if (rank[i] < 0) {
            rank[i] += D + 1;
            z[i] += 1;
        } else if (rank[i] > D) {
            rank[i] -= D + 1;
            z[i] -= 1;
        }
        // q = z or z-1 must stay inside the field, with room for a neighbour step
        in_range = in_range && (z[i] - 1 >= KeyCodec<D>::kQMin) && (z[i] <= KeyCodec<D>::kQMax);
        // a rank outside 0..D can only come from a non-finite feature
        in_range = in_range && (rank[i] >= 0) && (rank[i] <= D);
    }

    // barycentric weights (:222-240): with v sorted by rank,
    //   b[t] = v[rank = D-t] - v[rank = D-t+1],  b[0] += 1 + b[D+1]
    float v[D + 1], vs[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++) v[i] = __fmul_rn(__fsub_rn(el[i], __fmul_rn((float)z[i], dp1)), inv_dp1);
#pragma unroll
    for (int t = 0; t <= D; t++) {
        float sel = 0.0f;
#pragma unroll
        for (int i = 0; i <= D; i++) sel = (rank[i] == t) ? v[i] : sel;
        vs[t] = sel;
    }
#pragma unroll
    for (int t = 1; t <= D; t++) bary[t] = __fsub_rn(vs[D - t], vs[D - t + 1]);
    // b[D+1] = -v[rank 0];  b[0] = v[rank D] + (1 + b[D+1])
    bary[0] = __fadd_rn(vs[D], __fadd_rn(1.0f, -vs[0]));
    return in_range;
}

#if defined(__CUDACC__)
// 1 << s as a 64-bit word for s in [0, 64); 0 for s >= 64.  PTX shl.b32 clamps the shift amount at the register
// width, so the low word vanishes for s >= 32 and the high word for s < 32 (s - 32 wraps to a huge amount) and s = 64.
__device__ __forceinline__ unsigned long long one_shifted(unsigned int s)
{
    unsigned int lo, hi;
    asm("shl.b32 %0, 1, %1;" : "=r"(lo) : "r"(s));
    asm("shl.b32 %0, 1, %1;" : "=r"(hi) : "r"(s - 32u));
    return ((unsigned long long)hi << 32) | lo;
}

// embed_point + KeyCodec::pack_simplex in one pass, for kernels that can spare a per-thread scratch column in shared
// memory: `sf` (floats) and `si` (16-bit words) point at the calling thread's column of a [D+1][stride] array.  The
// same arithmetic, operation for operation, as the two functions above (tests compare the kernels that use either);
// what changes is how "the element with rank t" is found: the generic code selects it with (D+1)^2 compares (and
// decodes the inverse permutation from nibbles with 64-bit variable shifts), here every element is stored at the row
// its rank names and read back in order -- conflict-free (the column is the thread), ~110 instructions less at D = 5.
// Returns false when a quotient leaves the packed-key range (keys and weights are then well-formed but meaningless).
template <int D>
__device__ __forceinline__ bool embed_simplex_scratch(const float (&f)[D], const EmbedConsts &ec, float *sf,
                                                      unsigned short *si, int stride, unsigned long long (&key)[D + 1],
                                                      float (&bary)[D + 1])
{
    using Codec = KeyCodec<D>;
    constexpr float inv_dp1 = 1.0f / (D + 1);
    constexpr float dp1 = (float)(D + 1);
    float el[D + 1];
    float sm = 0.0f;
#pragma unroll
    for (int j = D; j > 0; j--) {
        const float cf = __fmul_rn(f[j - 1], ec.scale[j - 1]);
        el[j] = __fsub_rn(sm, __fmul_rn((float)j, cf));
        sm = __fadd_rn(sm, cf);
    }
    el[0] = sm;

    int z[D + 1], rank[D + 1];
    int sum = 0;
    float res[D + 1];
#pragma unroll
    for (int i = 0; i <= D; i++) {
        const int vi = __float2int_rn(__fmul_rn(inv_dp1, el[i]));
        z[i] = vi;
        res[i] = __fsub_rn(el[i], __fmul_rn((float)vi, dp1));
        sum += vi;
    }
    // rank[j] = #{i < j : res[i] >= res[j]} + #{k > j : res[j] < res[k]}  (the pairwise counts of embed_point)
#pragma unroll
    for (int i = 0; i <= D; i++) rank[i] = i;
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = i + 1; j <= D; j++) {
            const int c = res[i] < res[j] ? 1 : 0;
            rank[i] += c;
            rank[j] -= c;
        }
    int zmin = 0x7fffffff, zmax = -0x7fffffff - 1, rmin = 0x7fffffff, rmax = -0x7fffffff - 1;
#pragma unroll
    for (int i = 0; i <= D; i++) {
        rank[i] += sum;
        if (rank[i] < 0) {
            rank[i] += D + 1;
            z[i] += 1;
        } else if (rank[i] > D) {
            rank[i] -= D + 1;
            z[i] -= 1;
        }
        zmin = min(zmin, z[i]);
        zmax = max(zmax, z[i]);
        rmin = min(rmin, rank[i]);
        rmax = max(rmax, rank[i]);
    }
    // q = z or z-1 must stay inside the field, with room for a neighbour step; a rank outside 0..D can only come
    // from a non-finite feature
    const bool in_range = (zmin - 1 >= Codec::kQMin) && (zmax <= Codec::kQMax) && rmin >= 0 && rmax <= D;
    if (!in_range) {
#pragma unroll
        for (int i = 0; i <= D; i++) {
            z[i] = 0;
            rank[i] = i;
        }
    }
#pragma unroll
    for (int i = 0; i <= D; i++) {
        sf[rank[i] * stride] = __fmul_rn(__fsub_rn(el[i], __fmul_rn((float)z[i], dp1)), inv_dp1);
        // where coordinate i sits in the packed key (the implicit coordinate D has no field: shift 64 -> unit 0)
        si[rank[i] * stride] = (unsigned short)(i < D ? Codec::kRemBits + i * Codec::kFieldBits : 64);
    }
    float vs[D + 1];
#pragma unroll
    for (int t = 0; t <= D; t++) vs[t] = sf[t * stride];
#pragma unroll
    for (int t = 1; t <= D; t++) bary[t] = __fsub_rn(vs[D - t], vs[D - t + 1]);
    bary[0] = __fadd_rn(vs[D], __fadd_rn(1.0f, -vs[0]));

    int q0[D];
#pragma unroll
    for (int i = 0; i < D; i++) q0[i] = z[i];
    key[0] = Codec::pack(q0, 0);
#pragma unroll
    for (int r = 1; r <= D; r++) key[r] = key[r - 1] + 1ull - one_shifted(si[(D + 1 - r) * stride]);
    return in_range;
}

// ---------------------------------------------------------------------------
// Two-tier open-addressing table (one per frame), 16-byte entries {key, id}.
//
//   primary tier : `slots1` entries, sized for the lattice sizes real frames produce (a few
//                  vertices per pixel at most) so that the tables of the frames in flight stay in
//                  the 126 MB L2.  A key probes at most `window` consecutive slots here.
//   overflow tier: `slots2` entries, sized for the worst case (every pixel contributes d+1 distinct
//                  vertices).  Only keys whose primary window is full of other keys go here, so
//                  it is normally never touched -- and never needs clearing (see prepare_kernel).
//
// Occupied slots never change, so "window full of other keys" is a stable property: every thread
// that handles the same key takes the same decision and finds the same entry.
// Replaces HashTable::find / grow (permutohedral.cpp:18-41,67-95): no rehash is ever needed.
// ---------------------------------------------------------------------------
struct __align__(16) Entry {
    unsigned long long key;
    int id;   // dense vertex id, written by the thread that created the entry
    int pad;
};

struct TableGeom {
    unsigned int slots1, slots2;  // powers of two, 2^8 .. 2^30
    unsigned int window;          // max probes in the primary tier
    unsigned int shift1, shift2;  // 32 - log2(slots): hash_slot() keeps the top log2(slots) bits
    unsigned int base2;           // where the overflow tier starts inside a frame's table (the allocated size of
                                  // the primary tier; slots1 may be smaller: see effective_geom)
};

// The part of the primary tier a call really uses.  The allocation (base2 entries) is sized for iid-noise frames,
// ~1.1 vertices per pixel; real frames produce 10-20 times fewer, and clearing + probing a table of that size is
// pure overhead for them.  prepare_kernel therefore picks slots1 = the power of two >= 4 x the largest per-frame
// vertex count of the PREVIOUS call on this workspace (never above the allocation, and the allocation itself on
// first use or after a call that spilled into the overflow tier), and the kernels of the call read it back from the
// workspace.  A guess that turns out too small costs speed only: the window fills up and keys go to the overflow
// tier, which is sized for the worst case.  The result never depends on the size.
__device__ __forceinline__ TableGeom effective_geom(TableGeom g, unsigned int eff_slots1)
{
    g.slots1 = eff_slots1;
    g.shift1 = 1u + (unsigned int)__clz((int)eff_slots1);   // 32 - log2(eff) for a power of two
    g.window = eff_slots1 < g.window ? eff_slots1 : g.window;
    return g;
}
__device__ __forceinline__ unsigned int hash_primary(unsigned long long k, const TableGeom &g)
{
    return hash_slot(k, kHashMul1, g.shift1);
}
__device__ __forceinline__ unsigned int hash_overflow(unsigned long long k, const TableGeom &g)
{
    return hash_slot(k, kHashMul2, g.shift2);
}

__device__ __forceinline__ unsigned long long load_key_cg(const Entry *e)
{
    return __ldcg(&e->key);  // bypass L1: other SMs insert concurrently
}

// One probe: key and id of an entry with a single 16-byte load (L2; other SMs insert concurrently).  The id is
// written by the entry's creator with a plain 4-byte store some time after its compare-and-swap, so it may still
// read -1; it never changes once it is >= 0.
__device__ __forceinline__ void load_entry_cg(const Entry *e, unsigned long long &key, int &id)
{
    const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(e));
    key = ((unsigned long long)v.y << 32) | v.x;
    id = (int)v.z;
}

// Continues the insertion of `key` into the frame table `tab` (primary tier at [0, slots1), overflow at
// [slots1, slots1+slots2)) from primary slot `h`, after `probes` primary slots have already been rejected;
// `cur` holds the key read from slot `h`.  Callers run the first probes of all their keys in lock step
// (build_kernel) so that the round trips overlap, and come here only for the stragglers.
// Returns the entry index, or -1 when both tiers are full.  `won`: this call created the entry.
// `spilled`: the overflow tier was used.
__device__ __forceinline__ int table_insert_from(Entry *tab, const TableGeom g, unsigned long long key,
                                                 unsigned int h, unsigned int probes, unsigned long long cur,
                                                 bool &won, bool &spilled)
{
    won = false;
    spilled = false;
    const unsigned int mask1 = g.slots1 - 1;
    for (; probes < g.window; probes++) {
        if (cur == key) return (int)h;
        if (cur == kEmptyKey) {
            cur = atomicCAS(&tab[h].key, kEmptyKey, key);
            if (cur == kEmptyKey) {
                won = true;
                return (int)h;
            }
            if (cur == key) return (int)h;
        }
        h = (h + 1) & mask1;
        cur = load_key_cg(tab + h);
    }
    spilled = true;
    Entry *ov = tab + g.base2;
    const unsigned int mask2 = g.slots2 - 1;
    h = hash_overflow(key, g);
    for (probes = 0; probes <= mask2; probes++) {
        cur = load_key_cg(ov + h);
        if (cur == key) return (int)(g.base2 + h);
        if (cur == kEmptyKey) {
            cur = atomicCAS(&ov[h].key, kEmptyKey, key);
            if (cur == kEmptyKey) {
                won = true;
                return (int)(g.base2 + h);
            }
            if (cur == key) return (int)(g.base2 + h);
        }
        h = (h + 1) & mask2;
    }
    return -1;
}

// Continues a lookup from primary slot `h` whose entry `e` has already been loaded (`probes` slots rejected
// before it).  Returns the vertex id or -1.
__device__ __forceinline__ int table_lookup_from(const Entry *__restrict__ tab, const TableGeom g,
                                                 unsigned long long key, unsigned int h, unsigned int probes,
                                                 uint4 e)
{
    const unsigned int mask1 = g.slots1 - 1;
    for (; probes < g.window; probes++) {
        const unsigned long long cur = ((unsigned long long)e.y << 32) | e.x;
        if (cur == key) return (int)e.z;
        if (cur == kEmptyKey) return -1;
        h = (h + 1) & mask1;
        e = __ldg(reinterpret_cast<const uint4 *>(tab + h));
    }
    const Entry *ov = tab + g.base2;
    const unsigned int mask2 = g.slots2 - 1;
    h = hash_overflow(key, g);
    for (probes = 0; probes <= mask2; probes++) {
        e = __ldg(reinterpret_cast<const uint4 *>(ov + h));
        const unsigned long long cur = ((unsigned long long)e.y << 32) | e.x;
        if (cur == key) return (int)e.z;
        if (cur == kEmptyKey) return -1;
        h = (h + 1) & mask2;
    }
    return -1;
}

// Finds `key` (table complete, read-only); returns its vertex id or -1.  One 16-byte load per probe
// returns both the key and the id.
__device__ __forceinline__ int table_lookup(const Entry *__restrict__ tab, const TableGeom g,
                                            unsigned long long key)
{
    const unsigned int h = hash_primary(key, g);
    const uint4 e = __ldg(reinterpret_cast<const uint4 *>(tab + h));
    return table_lookup_from(tab, g, key, h, 0, e);
}
#endif  // __CUDACC__

}  // namespace tcamcrf
