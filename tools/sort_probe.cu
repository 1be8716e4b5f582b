// Measurement probe (not on the product path): what would the sort-based lattice build VERDICT r1 item 2 names cost?
//
// A sort-based build of one step of the headline input (32 frames x 224 x 224, d = 5) has to order
// N * (d+1) * P = 9 633 792 packed 64-bit simplex keys (frame id in the top bits, so one sort is the per-frame
// segmented sort), carrying the 32-bit point index each key came from, then run-length encode them into dense vertex
// ids and scatter the ids back.  This program times the first two of those three steps with the library radix sort
// (cub::DeviceRadixSort, one-sweep; the fastest sort this image has) restricted to the key bits in use, on keys with the
// multiplicity of the noise input (about 1.1 distinct keys per pixel out of 6), so that the number can stand next to
// the hash build's 0.206 ms for the same step.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/sort_probe tools/sort_probe.cu
//   tools/bin/sort_probe [frames] [pixels] [key_bits]
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            std::fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));    \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

// splitmix64: keys with a chosen number of distinct values per frame
__global__ void make_keys(unsigned long long *keys, unsigned int *vals, long long n, long long per_frame, long long distinct, int key_bits)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long frame = i / per_frame;
    unsigned long long z = (unsigned long long)(i % per_frame);
    z = (z * 0x9E3779B97F4A7C15ull) >> 11;
    z %= (unsigned long long)distinct;              // which of the frame's vertices this simplex corner is
    z = (z + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    const int frame_bits = 5;
    z &= (1ull << (key_bits - frame_bits)) - 1ull;
    keys[i] = ((unsigned long long)frame << (key_bits - frame_bits)) | z;
    vals[i] = (unsigned int)i;
}

int main(int argc, char **argv)
{
    const long long frames = argc > 1 ? std::atoll(argv[1]) : 32;
    const long long pixels = argc > 2 ? std::atoll(argv[2]) : 224 * 224;
    const int key_bits = argc > 3 ? std::atoi(argv[3]) : 56;   // 5 x 10-bit coordinates + frame id, rounded up to a byte
    const int corners = 6;
    const long long n = frames * pixels * corners;
    const long long distinct = (long long)(1.1 * (double)pixels);

    unsigned long long *k_in, *k_out;
    unsigned int *v_in, *v_out, *rle_len;
    unsigned long long *rle_key;
    long long *rle_runs;
    CK(cudaMalloc(&k_in, n * 8));
    CK(cudaMalloc(&k_out, n * 8));
    CK(cudaMalloc(&v_in, n * 4));
    CK(cudaMalloc(&v_out, n * 4));
    CK(cudaMalloc(&rle_key, n * 8));
    CK(cudaMalloc(&rle_len, n * 4));
    CK(cudaMalloc(&rle_runs, 8));
    make_keys<<<(unsigned int)((n + 255) / 256), 256>>>(k_in, v_in, n, pixels * corners, distinct, key_bits);
    CK(cudaDeviceSynchronize());

    size_t tmp_pairs = 0, tmp_keys = 0, tmp_rle = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_pairs, k_in, k_out, v_in, v_out, n, 0, key_bits));
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_keys, k_in, k_out, n, 0, key_bits));
    CK(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_rle, k_out, rle_key, rle_len, rle_runs, n));
    size_t tmp_bytes = tmp_pairs > tmp_keys ? tmp_pairs : tmp_keys;
    if (tmp_rle > tmp_bytes) tmp_bytes = tmp_rle;
    void *tmp;
    CK(cudaMalloc(&tmp, tmp_bytes));

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int reps = 50;
    float ms_pairs = 0.f, ms_keys = 0.f, ms_rle = 0.f;
    for (int w = 0; w < 5; w++) CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_pairs, k_in, k_out, v_in, v_out, n, 0, key_bits));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_pairs, k_in, k_out, v_in, v_out, n, 0, key_bits));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms_pairs, e0, e1));
    for (int w = 0; w < 5; w++) CK(cub::DeviceRadixSort::SortKeys(tmp, tmp_keys, k_in, k_out, n, 0, key_bits));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) CK(cub::DeviceRadixSort::SortKeys(tmp, tmp_keys, k_in, k_out, n, 0, key_bits));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms_keys, e0, e1));
    CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_pairs, k_in, k_out, v_in, v_out, n, 0, key_bits));
    for (int w = 0; w < 5; w++) CK(cub::DeviceRunLengthEncode::Encode(tmp, tmp_rle, k_out, rle_key, rle_len, rle_runs, n));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) CK(cub::DeviceRunLengthEncode::Encode(tmp, tmp_rle, k_out, rle_key, rle_len, rle_runs, n));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms_rle, e0, e1));
    long long runs = 0;
    CK(cudaMemcpy(&runs, rle_runs, 8, cudaMemcpyDeviceToHost));

    std::printf("{\"keys\": %lld, \"key_bits\": %d, \"distinct\": %lld, \"sort_pairs_ms\": %.4f, \"sort_keys_ms\": %.4f, "
                "\"run_length_encode_ms\": %.4f, \"sort_pairs_gkeys_per_s\": %.2f, \"what\": \"cub::DeviceRadixSort of the "
                "packed simplex keys of one step (+ 32-bit point index), then cub::DeviceRunLengthEncode: the first two "
                "steps of a sort-based lattice build\"}\n",
                n, key_bits, runs, ms_pairs / reps, ms_keys / reps, ms_rle / reps, (double)n / (ms_pairs / reps) * 1e-6);
    return 0;
}
