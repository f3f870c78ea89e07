// tests/cpp/stats_allreduce.cpp -- a C++ host (the language north_star fixes for the host side) reducing the statistics of an
// ensemble sharded over the GPUs of one box with rbis_batch_stats_allreduce(handle, ncclComm_t, ...).
//
// One process, one thread and one NCCL rank per GPU (ncclCommInitAll); every rank owns a contiguous shard of N_TOTAL
// filters, runs the same short IMU + leg-odometry program on it and calls the collective.  Checks:
//   * every rank receives the same totals, bit for bit;
//   * they equal the single-GPU reduction of the whole ensemble (rbis_batch_stats + rbis_stats_reduce_chunks), bit for bit.
// With one visible GPU the communicator has one rank (the code path is the same); usage: stats_allreduce [n_gpus]
#include <cuda_runtime.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/rbis_batch.h"

#define CHECK(x)                                                                    \
  do {                                                                              \
    if ((x) != 0) {                                                                 \
      std::fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, rbis_last_error(), __FILE__, __LINE__); \
      std::exit(2);                                                                 \
    }                                                                               \
  } while (0)

static const int64_t N_TOTAL = 8192, T = 12;
static const int CHUNK = 256;

// deterministic pseudo-random inputs, a function of the GLOBAL filter index only
static double u01(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; x ^= x >> 31;
  return (double)(x >> 11) / 9007199254740992.0 - 0.5;
}

struct Shard {
  int64_t lo, n;
  std::vector<double> vec, quat, cov, imu, z;
};
static Shard make_shard(int64_t lo, int64_t n) {
  Shard s{lo, n, std::vector<double>(21 * n, 0.0), std::vector<double>(4 * n, 0.0), std::vector<double>(441 * n, 0.0),
          std::vector<double>((size_t)T * 6 * n), std::vector<double>((size_t)(T / 2) * 3 * n)};
  for (int64_t k = 0; k < n; k++) {
    const uint64_t g = (uint64_t)(lo + k);
    for (int i = 3; i < 6; i++) s.vec[i * n + k] = 0.2 * u01(g * 131 + i);
    s.vec[11 * n + k] = 0.85;
    s.quat[k] = 1.0;
    const double sig[21] = {0, 0, 0, .1, .1, .1, .05, .05, .05, .1, .1, .1, 0, 0, 0, .002, .002, .002, .1, .1, .1};
    for (int i = 0; i < 21; i++) s.cov[(size_t)(i + 21 * i) * n + k] = sig[i] * sig[i];
    for (int64_t t = 0; t < T; t++) {
      for (int c = 0; c < 6; c++) s.imu[((size_t)t * 6 + c) * n + k] = 0.05 * u01(g * 977 + t * 13 + c) + (c == 5 ? 9.8 : 0.0);
      if (t % 2 == 0)
        for (int c = 0; c < 3; c++) s.z[((size_t)(t / 2) * 3 + c) * n + k] = 0.1 * u01(g * 7919 + t * 17 + c);
    }
  }
  return s;
}
static void run_program(rbis_batch_t* h, const Shard& s) {
  CHECK(rbis_batch_set_process_noise(h, 7.6e-5, 1e-2, 3e-10, 1e-6));
  CHECK(rbis_batch_set_state(h, s.vec.data(), s.quat.data(), s.cov.data(), nullptr, 0, RBIS_MEM_HOST));
  std::vector<rbis_op_t> ops;
  for (int64_t t = 0; t < T; t++) {
    ops.push_back({RBIS_OP_IMU, 0, t, (t + 1) * 1000, 1e-3});
    if (t % 2 == 0) ops.push_back({RBIS_OP_MEAS, 0, t / 2, (t + 1) * 1000, 0.0});
  }
  const double R[9] = {0.01, 0, 0, 0, 0.01, 0, 0, 0, 0.01};
  rbis_stream_t st;
  std::memset(&st, 0, sizeof(st));
  st.m = 3; st.idx[0] = 3; st.idx[1] = 4; st.idx[2] = 5; st.r_mode = RBIS_R_SHARED_FULL;
  st.z = s.z.data(); st.R = R; st.rows = T / 2;
  CHECK(rbis_batch_run_fused(h, (int64_t)ops.size(), ops.data(), s.imu.data(), T, 1, &st, RBIS_MEM_HOST));
  CHECK(rbis_batch_synchronize(h));
}

int main(int argc, char** argv) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { std::fprintf(stderr, "no CUDA device\n"); return 5; }
  int world = argc > 1 ? std::atoi(argv[1]) : (ndev >= 2 ? 2 : 1);
  if (world > ndev) world = ndev;
  std::vector<int> devs(world);
  for (int r = 0; r < world; r++) devs[r] = r;
  std::vector<ncclComm_t> comms(world);
  if (ncclCommInitAll(comms.data(), world, devs.data()) != ncclSuccess) { std::fprintf(stderr, "ncclCommInitAll failed\n"); return 3; }
  const double truth_vec[21] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0.85, 0, 0, 0, 0, 0, 0, 0, 0, 0}, truth_quat[4] = {1, 0, 0, 0};
  const int64_t total_chunks = N_TOTAL / CHUNK, per = N_TOTAL / world;
  std::vector<std::vector<double>> totals(world, std::vector<double>(RBIS_NUM_STATS)), tables(world, std::vector<double>(total_chunks * RBIS_NUM_STATS));
  std::vector<std::thread> th;
  for (int r = 0; r < world; r++)
    th.emplace_back([&, r]() {
      rbis_batch_config_t cfg;
      rbis_default_config(&cfg);
      cfg.device = r;
      rbis_batch_t* h = nullptr;
      CHECK(rbis_batch_create(&h, per, &cfg));
      const Shard s = make_shard(r * per, per);
      run_program(h, s);
      CHECK(rbis_batch_stats_allreduce(h, (void*)comms[r], truth_vec, truth_quat, CHUNK, r * per / CHUNK, total_chunks,
                                       totals[r].data(), tables[r].data()));
      CHECK(rbis_batch_destroy(h));
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < world; r++) ncclCommDestroy(comms[r]);
  // single-GPU reference over the whole ensemble
  rbis_batch_config_t cfg;
  rbis_default_config(&cfg);
  cfg.device = 0;
  rbis_batch_t* h = nullptr;
  CHECK(rbis_batch_create(&h, N_TOTAL, &cfg));
  run_program(h, make_shard(0, N_TOTAL));
  std::vector<double> chunks(total_chunks * RBIS_NUM_STATS), ref(RBIS_NUM_STATS);
  int64_t nch = 0;
  CHECK(rbis_batch_stats(h, truth_vec, truth_quat, 0, CHUNK, chunks.data(), &nch, nullptr, RBIS_MEM_HOST));
  CHECK(rbis_stats_reduce_chunks(chunks.data(), nch, ref.data()));
  CHECK(rbis_batch_destroy(h));
  int bad = 0;
  for (int r = 0; r < world; r++) {
    if (std::memcmp(totals[r].data(), ref.data(), sizeof(double) * RBIS_NUM_STATS)) { std::fprintf(stderr, "rank %d totals differ from the single-GPU reduction\n", r); bad++; }
    if (std::memcmp(tables[r].data(), chunks.data(), sizeof(double) * chunks.size())) { std::fprintf(stderr, "rank %d chunk table differs\n", r); bad++; }
  }
  std::printf("stats_allreduce: world=%d filters=%lld chunks=%lld count=%.0f non_finite=%.0f sum_nees=%.17g -> %s\n", world, (long long)N_TOTAL,
              (long long)total_chunks, ref[46], ref[45], ref[42], bad ? "MISMATCH" : "bit-identical on every rank and to the single-GPU reduction");
  return bad ? 1 : 0;
}
