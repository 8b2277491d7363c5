// engine.h - internal state of a mamcts_engine (one CUDA device, one stream, scratch, NCCL comm).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>

#include "mamcts_b200.h"

struct mamcts_engine {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::mutex mu;  // calls on one engine are serialised (SURVEY.md §8b threading)
  std::mutex err_mu;   // guards last_error alone (fail() runs with or without `mu` held)
  std::string last_error;
  uint64_t launches = 0;
  // staging / scratch (grow-only)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* reduce_buf = nullptr;  // ticket + partials of the reductions (reduce.cu); separate from `scratch`,
  size_t reduce_bytes = 0;     // which holds the per-rollout values a moments reduction may be reading
  void* staging = nullptr;  // device staging for host-buffer calls
  size_t staging_bytes = 0;
  void* pinned = nullptr;   // pinned host staging
  size_t pinned_bytes = 0;
  // work-list cursor of the persistent rollout kernel (one uint32, zeroed per launch)
  unsigned int* work_cursor = nullptr;
  // discount / probability tables of the rollout (see ensure_discount_table)
  double* disc = nullptr;
  uint32_t disc_len = 0;
  double disc_gamma = 0.0, disc_inv = 0.0;
  // NCCL (loaded lazily with dlopen; see comm.cu)
  void* nccl_comm = nullptr;
  int world_size = 1;
  int rank = 0;
};

namespace mamcts {

void set_global_error(const std::string& msg);
int fail(mamcts_engine* e, int code, const std::string& msg);
int cuda_fail(mamcts_engine* e, cudaError_t err, const char* what);

#define MAMCTS_CUDA(e, call)                                       \
  do {                                                             \
    cudaError_t err__ = (call);                                    \
    if (err__ != cudaSuccess) return ::mamcts::cuda_fail((e), err__, #call); \
  } while (0)

// grow-only device scratch / staging; contents undefined after the call
int ensure_scratch(mamcts_engine* e, size_t bytes);
int ensure_reduce_buf(mamcts_engine* e, size_t bytes);
// {sum, sum^2} per agent + count over device arrays (reduce.cu), asynchronous on the engine stream, all-reduced
// over the communicator when one is attached; d_dst: 2 * (num_others + 1) + 1 doubles on the device. No lock taken.
int rollout_moments_device(mamcts_engine* e, const double* d_ego, const double* d_other, uint32_t num_others,
                           uint32_t n, double* d_dst);
int ensure_staging(mamcts_engine* e, size_t bytes);
int ensure_pinned(mamcts_engine* e, size_t bytes);

// Bump allocator over the staging buffers used to mirror HOST arguments on the device.
struct Stager {
  mamcts_engine* e;
  size_t offset = 0;
  size_t capacity = 0;
  explicit Stager(mamcts_engine* eng) : e(eng) {}
  static size_t align(size_t n) { return (n + 255) & ~size_t(255); }
};

// The rollout's discount sequence depends only on (gamma, 1/nA_ego, step index): m1[k] = the weight
// the k-th step's reward is multiplied with (random_heuristic.h:71-72,85) and prob[k] = (1/nA_ego)^k
// (:88), both produced by the reference's own chain of individually rounded multiplications. They are
// tabulated once per parameter set by a one-thread kernel, so that the step loops do no FP64 work
// unless a reward is non-zero. Layout: tab[0, len) = m1, tab[len, 2 len) = prob; len = max steps + 1.
// Writes the device pointer to *tab (owned by the engine, valid until the next call with other
// parameters on the same engine stream).
int ensure_discount_table(mamcts_engine* e, double gamma, double inv_n_ego, uint32_t max_steps,
                          const double** tab, uint32_t* len);
// Same table into caller-owned device memory of 2 * (max_steps + 1) doubles.
int fill_discount_table(mamcts_engine* e, double gamma, double inv_n_ego, uint32_t max_steps, double* dst);

inline int grid_for(const mamcts_engine* e, int blocks_per_sm) { return e->num_sms * blocks_per_sm; }

}  // namespace mamcts
