// Epoch shuffle of the training set on the device.
//
// Reference: GPUShuffler::ReShuffle / DistShuffler::ReShuffle (cuda_shuffler.cc:75-126,
// dist_shuffler.cc:98-151) run a host Fisher-Yates over the whole train set every
// epoch (std::default_random_engine; seed = wall clock single-GPU, = epoch in the
// multi-process engine) and copy it H2D.  On B200 that O(n_train) host loop would be
// a fifth of a papers100M epoch, so the permutation is drawn on the GPU instead:
// key_i = Philox(seed, epoch)[i], stable radix sort of (key_i, train_set[i]).  Like
// the reference's dist shuffler it is a pure function of (seed, epoch), so every
// sampler GPU derives the same permutation without communicating.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace fgnn {
namespace {

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__global__ void __launch_bounds__(kBlock)
shuffle_keys_kernel(uint32_t *keys, size_t n, RngKey key) {
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride)
    keys[i] = rand_u32(key, (uint32_t)i, 0);
}

size_t sort_bytes(size_t n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)n, 0, 32,
                                  (cudaStream_t)0);
  return bytes;
}

}  // namespace
}  // namespace fgnn

using namespace fgnn;

extern "C" size_t fgnn_k_shuffle_workspace_bytes(size_t n) {
  if (n == 0) n = 1;
  return 2 * align256(n * 4) + align256(sort_bytes(n));
}

extern "C" int fgnn_k_shuffle(const uint32_t *train_set, size_t n, uint64_t seed, uint64_t epoch,
                              uint32_t *out, void *workspace, size_t workspace_bytes,
                              fgnn_stream_t stream) {
  if (n == 0) return 0;
  if (!train_set || !out || !workspace || n > 0xFFFFFFFFull) return FGNN_ERR_BAD_ARG;
  if (workspace_bytes < fgnn_k_shuffle_workspace_bytes(n)) return FGNN_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t *k0 = (uint32_t *)workspace;
  uint32_t *k1 = (uint32_t *)((char *)workspace + align256(n * 4));
  void *temp = (char *)workspace + 2 * align256(n * 4);
  size_t temp_bytes = workspace_bytes - 2 * align256(n * 4);
  fgnn_rng r;
  r.seed = seed;
  r.batch_key = epoch;
  r.tag = FGNN_SHUFFLE_TAG;
  shuffle_keys_kernel<<<persistent_grid(n, 4 * kBlock, 8, false), kBlock, 0, st>>>(k0, n, make_rng_key(r));
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, k0, k1, train_set, out, (int64_t)n,
                                                  0, 32, st);
  if (e != cudaSuccess) return (int)e;
  note_launch();
  return check_last();
}
