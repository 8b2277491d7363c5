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
  std::string last_error;
  uint64_t launches = 0;
  // staging / scratch (grow-only)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  void* staging = nullptr;  // device staging for host-buffer calls
  size_t staging_bytes = 0;
  void* pinned = nullptr;   // pinned host staging
  size_t pinned_bytes = 0;
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

inline int grid_for(const mamcts_engine* e, int blocks_per_sm) { return e->num_sms * blocks_per_sm; }

}  // namespace mamcts
