// Row-sharded 1:N search from a C++ host, no Python: one fr_ctx + gallery shard per GPU of this
// process, one NCCL communicator per GPU (ncclCommInitAll), one host thread per rank calling
// fr_gallery_search_sharded (local fused GEMM + top-k -> one all-gather of packed
// records over NVLink -> rank merge).  Checks every rank's result against a single-GPU search
// of the whole gallery on device 0: indices and scores must be identical (rank-count invariance).
//
//   sharded_search_demo [world] [rows_per_rank] [queries]
//
// The reference has no multi-GPU path (src/main.cpp:264-319 is a single-process CLI); this is the
// multi-GPU entry a C++ deployment of the 1:N extension would use.
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "../../include/fr_capi.h"

#include <cuda_runtime_api.h>

#define CHECK(x)                                                                   \
  do {                                                                             \
    const int s_ = (x);                                                            \
    if (s_ != 0) {                                                                 \
      std::fprintf(stderr, "%s failed with %d (%s:%d)\n", #x, s_, __FILE__, __LINE__); \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

int main(int argc, char** argv) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != 0 || ndev <= 0) {
    std::fprintf(stderr, "no CUDA device\n");
    return 2;
  }
  const int world = argc > 1 ? std::min(std::atoi(argv[1]), ndev) : std::min(ndev, 8);
  const int64_t rows_per_rank = argc > 2 ? std::atoll(argv[2]) : 30011;
  const int nq = argc > 3 ? std::atoi(argv[3]) : 257;
  const int k = 10;
  const uint64_t seed = 1000;

  std::vector<int> devs(world);
  for (int r = 0; r < world; ++r) devs[r] = r;
  std::vector<ncclComm_t> comms(world);
  CHECK(ncclCommInitAll(comms.data(), world, devs.data()));

  std::vector<fr_ctx*> ctx(world, nullptr);
  std::vector<fr_gallery*> shard(world, nullptr);
  for (int r = 0; r < world; ++r) {
    CHECK(fr_create(&ctx[r], r, nullptr, nullptr));
    CHECK(fr_gallery_create(ctx[r], &shard[r], rows_per_rank, r * rows_per_rank));
    CHECK(fr_gallery_fill_synthetic(shard[r], rows_per_rank, seed));
  }
  // the whole gallery on device 0 (same seed, global row index in the hash -> the same rows)
  fr_gallery* whole = nullptr;
  CHECK(fr_gallery_create(ctx[0], &whole, rows_per_rank * world, 0));
  CHECK(fr_gallery_fill_synthetic(whole, rows_per_rank * world, seed));

  // queries: noisy copies of gallery rows spread over all shards (+ one exact copy), unit norm
  std::vector<float> q((size_t)nq * FR_FEAT_DIM);
  std::mt19937 rng(7);
  std::normal_distribution<float> noise(0.f, 0.02f);
  std::vector<float> row(FR_FEAT_DIM);
  for (int i = 0; i < nq; ++i) {
    const int64_t g = (int64_t)((uint64_t)rng() % (uint64_t)(rows_per_rank * world));
    CHECK(fr_gallery_get_rows(whole, g, 1, row.data()));
    double ss = 0;
    for (int d = 0; d < FR_FEAT_DIM; ++d) {
      const float v = row[d] + (i == 0 ? 0.f : noise(rng));
      q[(size_t)i * FR_FEAT_DIM + d] = v;
      ss += (double)v * v;
    }
    const float inv = (float)(1.0 / std::sqrt(ss));
    for (int d = 0; d < FR_FEAT_DIM; ++d) q[(size_t)i * FR_FEAT_DIM + d] *= inv;
  }

  std::vector<float> ref_s((size_t)nq * k);
  std::vector<int64_t> ref_i((size_t)nq * k);
  CHECK(fr_gallery_search(whole, q.data(), nq, k, FR_MEM_HOST, ref_s.data(), ref_i.data()));

  std::vector<std::vector<float>> out_s(world, std::vector<float>((size_t)nq * k));
  std::vector<std::vector<int64_t>> out_i(world, std::vector<int64_t>((size_t)nq * k));
  std::vector<int> status(world, 0);
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r)
    th.emplace_back([&, r] {
      for (int rep = 0; rep < 3 && status[r] == 0; ++rep)   // repeated calls reuse the exchange buffers
        status[r] = fr_gallery_search_sharded(shard[r], comms[r], world, q.data(), nq, k, FR_MEM_HOST, out_s[r].data(),
                                              out_i[r].data());
    });
  for (auto& t : th) t.join();
  int bad = 0;
  for (int r = 0; r < world; ++r) {
    if (status[r] != 0) {
      std::fprintf(stderr, "rank %d: fr_gallery_search_sharded -> %d (%s)\n", r, status[r], fr_last_error(ctx[r]));
      return 1;
    }
    if (std::memcmp(out_i[r].data(), ref_i.data(), ref_i.size() * sizeof(int64_t)) != 0) ++bad;
    if (std::memcmp(out_s[r].data(), ref_s.data(), ref_s.size() * sizeof(float)) != 0) ++bad;
  }
  std::printf("query 0: top-1 global row %lld score %.6f\n", (long long)ref_i[0], ref_s[0]);
  for (int r = 0; r < world; ++r) {
    fr_gallery_destroy(shard[r]);
  }
  fr_gallery_destroy(whole);
  for (int r = 0; r < world; ++r) {
    fr_destroy(ctx[r]);
    ncclCommDestroy(comms[r]);
  }
  if (bad) {
    std::fprintf(stderr, "sharded search differs from the single-GPU search on %d outputs\n", bad);
    return 1;
  }
  std::printf("sharded_search_demo ok: world=%d rows=%lld queries=%d\n", world, (long long)(rows_per_rank * world), nq);
  return 0;
}
