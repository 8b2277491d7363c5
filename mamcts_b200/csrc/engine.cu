// engine.cu - engine lifetime, memory, stream, parameter defaults and the host-side reference tables.
#include "engine.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

namespace mamcts {

static std::mutex g_err_mu;
static std::string g_last_error;

void set_global_error(const std::string& msg) {
  std::lock_guard<std::mutex> lock(g_err_mu);
  g_last_error = msg;
}

int fail(mamcts_engine* e, int code, const std::string& msg) {
  if (e) {
    std::lock_guard<std::mutex> lock(e->err_mu);
    e->last_error = msg;
  }
  set_global_error(msg);
  return code;
}

int cuda_fail(mamcts_engine* e, cudaError_t err, const char* what) {
  std::string msg = std::string("CUDA error: ") + cudaGetErrorString(err) + " in " + what;
  return fail(e, MAMCTS_ERR_CUDA, msg);
}

static int grow(mamcts_engine* e, void** buf, size_t* have, size_t need, bool pinned) {
  if (*have >= need) return MAMCTS_OK;
  // everything queued may still read the old buffer
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  if (*buf) {
    if (pinned) cudaFreeHost(*buf); else cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
  }
  size_t bytes = std::max<size_t>(need + need / 4, size_t(1) << 20);
  if (pinned) MAMCTS_CUDA(e, cudaMallocHost(buf, bytes));
  else MAMCTS_CUDA(e, cudaMalloc(buf, bytes));
  *have = bytes;
  return MAMCTS_OK;
}

int ensure_scratch(mamcts_engine* e, size_t bytes) { return grow(e, &e->scratch, &e->scratch_bytes, bytes, false); }
int ensure_reduce_buf(mamcts_engine* e, size_t bytes) { return grow(e, &e->reduce_buf, &e->reduce_bytes, bytes, false); }
int ensure_staging(mamcts_engine* e, size_t bytes) { return grow(e, &e->staging, &e->staging_bytes, bytes, false); }
int ensure_pinned(mamcts_engine* e, size_t bytes) { return grow(e, &e->pinned, &e->pinned_bytes, bytes, true); }

// One thread: the sequence is a chain of dependent, individually rounded multiplications
// (reference random_heuristic.h:46,71,85,88), a few microseconds for the default cap of 1000.
__global__ void discount_table_kernel(double gamma, double inv_n_ego, uint32_t len, double* tab) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double m = gamma, prob = 1.0;      // :46, :48
  for (uint32_t k = 0; k < len; ++k) {
    const double m1 = __dmul_rn(gamma, m);   // :71
    tab[k] = m1;
    tab[len + k] = prob;                     // prob after k steps
    m = __dmul_rn(m1, gamma);                // :85
    prob = __dmul_rn(prob, inv_n_ego);       // :88
  }
}

int fill_discount_table(mamcts_engine* e, double gamma, double inv_n_ego, uint32_t max_steps, double* dst) {
  discount_table_kernel<<<1, 32, 0, e->stream>>>(gamma, inv_n_ego, max_steps + 1, dst);
  e->launches++;
  MAMCTS_CUDA(e, cudaGetLastError());
  return MAMCTS_OK;
}

int ensure_discount_table(mamcts_engine* e, double gamma, double inv_n_ego, uint32_t max_steps,
                          const double** tab, uint32_t* len) {
  const uint32_t need = max_steps + 1;
  const bool same = e->disc && e->disc_len == need &&
                    std::memcmp(&e->disc_gamma, &gamma, sizeof(double)) == 0 &&
                    std::memcmp(&e->disc_inv, &inv_n_ego, sizeof(double)) == 0;
  if (!same) {
    if (e->disc_len < need || !e->disc) {
      // the old table may still be read by kernels in flight on the stream
      if (e->disc) { MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream)); MAMCTS_CUDA(e, cudaFree(e->disc)); e->disc = nullptr; }
      MAMCTS_CUDA(e, cudaMalloc(reinterpret_cast<void**>(&e->disc), sizeof(double) * 2 * (size_t)need));
    }
    int rc = fill_discount_table(e, gamma, inv_n_ego, max_steps, e->disc);
    if (rc != MAMCTS_OK) return rc;
    e->disc_len = need;
    e->disc_gamma = gamma;
    e->disc_inv = inv_n_ego;
  }
  *tab = e->disc;
  *len = need;
  return MAMCTS_OK;
}

}  // namespace mamcts

using namespace mamcts;

extern "C" {

int mamcts_abi_version(void) { return MAMCTS_ABI_VERSION; }

int mamcts_create(const mamcts_config* cfg, mamcts_engine** out) {
  if (!cfg || !out) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_create: null argument");
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0) {
    return fail(nullptr, MAMCTS_ERR_CUDA,
                std::string("mamcts_create: no CUDA device (") + cudaGetErrorString(err) +
                    "); the engine has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= count)
    return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_create: device ordinal out of range");
  mamcts_engine* e = new mamcts_engine();
  e->device = cfg->device;
  cudaDeviceProp prop;
  err = cudaSetDevice(e->device);
  if (err == cudaSuccess) err = cudaGetDeviceProperties(&prop, e->device);
  if (err != cudaSuccess) { int rc = cuda_fail(nullptr, err, "mamcts_create: cudaSetDevice / cudaGetDeviceProperties"); delete e; return rc; }
  e->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    std::string msg = "mamcts_create: device " + std::string(prop.name) +
                      " is not sm_100-class; this library is compiled for sm_100a only";
    delete e;
    return fail(nullptr, MAMCTS_ERR_UNSUPPORTED, msg);
  }
  err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (err != cudaSuccess) { int rc = cuda_fail(nullptr, err, "cudaStreamCreate"); delete e; return rc; }
  e->own_stream = true;
  *out = e;
  return MAMCTS_OK;
}

void mamcts_destroy(mamcts_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  mamcts_comm_destroy(e);
  if (e->scratch) cudaFree(e->scratch);
  if (e->reduce_buf) cudaFree(e->reduce_buf);
  if (e->staging) cudaFree(e->staging);
  if (e->pinned) cudaFreeHost(e->pinned);
  if (e->disc) cudaFree(e->disc);
  if (e->work_cursor) cudaFree(e->work_cursor);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

const char* mamcts_last_error(const mamcts_engine* e) {
  // a per-thread copy: the engine may be shared by several host threads, one of which may fail while another reads
  static thread_local std::string copy;
  if (e) {
    mamcts_engine* m = const_cast<mamcts_engine*>(e);
    std::lock_guard<std::mutex> lock(m->err_mu);
    copy = m->last_error;
    return copy.c_str();
  }
  std::lock_guard<std::mutex> lock(g_err_mu);
  copy = g_last_error;
  return copy.c_str();
}

int mamcts_set_stream(mamcts_engine* e, void* cuda_stream) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_set_stream: null engine");
  std::lock_guard<std::mutex> lock(e->mu);
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  e->stream = static_cast<cudaStream_t>(cuda_stream);
  e->own_stream = false;
  return MAMCTS_OK;
}

void* mamcts_get_stream(mamcts_engine* e) { return e ? static_cast<void*>(e->stream) : nullptr; }

int mamcts_sync(mamcts_engine* e) {
  if (!e) return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_sync: null engine");
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  MAMCTS_CUDA(e, cudaGetLastError());
  return MAMCTS_OK;
}

uint64_t mamcts_launch_count(const mamcts_engine* e) { return e ? e->launches : 0; }

int mamcts_device_alloc(mamcts_engine* e, size_t bytes, void** dptr) {
  if (!e || !dptr) return fail(e, MAMCTS_ERR_INVALID, "mamcts_device_alloc: null argument");
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  MAMCTS_CUDA(e, cudaMalloc(dptr, bytes ? bytes : 1));
  return MAMCTS_OK;
}

int mamcts_device_free(mamcts_engine* e, void* dptr) {
  if (!e) return fail(e, MAMCTS_ERR_INVALID, "mamcts_device_free: null engine");
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  MAMCTS_CUDA(e, cudaFree(dptr));
  return MAMCTS_OK;
}

int mamcts_memcpy_h2d(mamcts_engine* e, void* dst, const void* src, size_t bytes) {
  if (!e) return fail(e, MAMCTS_ERR_INVALID, "mamcts_memcpy_h2d: null engine");
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  MAMCTS_CUDA(e, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  return MAMCTS_OK;
}

int mamcts_memcpy_d2h(mamcts_engine* e, void* dst, const void* src, size_t bytes) {
  if (!e) return fail(e, MAMCTS_ERR_INVALID, "mamcts_memcpy_d2h: null engine");
  MAMCTS_CUDA(e, cudaSetDevice(e->device));
  MAMCTS_CUDA(e, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e->stream));
  MAMCTS_CUDA(e, cudaStreamSynchronize(e->stream));
  return MAMCTS_OK;
}

// mcts_default_parameters(), reference mcts/mcts_parameters.h:97-135
void mamcts_default_params(mamcts_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->discount_factor = 0.9;
  p->random_seed = 1000;
  p->max_number_of_iterations = 10000;
  p->max_number_of_nodes = 10000;
  p->max_search_depth = 1000;
  p->use_bound_estimation = 1;
  p->num_parallel_mcts = 1;
  p->rh_max_number_of_iterations = 1000;
  p->uct_lower_bound = -1000;
  p->uct_upper_bound = 100;
  p->uct_exploration_constant = 0.7;
  p->uct_pw_k = 4.0;
  p->uct_pw_alpha = 0.25;
}

// default_crossing_state_parameters<int>(), reference environments/crossing_state_parameters.h:36-59
void mamcts_default_crossing_params_f32(mamcts_crossing_params_f32* p) {
  // default_crossing_state_parameters<float>(), environments/crossing_state_parameters.h:36-59
  if (!p) return;
  p->num_other_agents = 2;
  p->cost_only_collision = 0;
  p->max_velocity_other = 3.0f;
  p->min_velocity_other = -3.0f;
  p->max_velocity_ego = 2.0f;
  p->min_velocity_ego = -1.0f;
  p->chain_length = 21.0f;
  p->ego_goal_pos = 12.0f;
  p->num_other_actions = 30;
  p->reward_collision = -1000.0;
  p->reward_goal_reached = 100.0;
  p->reward_step = 0.0;
  p->reserved = 0;
}

void mamcts_default_crossing_params(mamcts_crossing_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->num_other_agents = 2;
  p->cost_only_collision = 0;
  p->max_velocity_other = 3;
  p->min_velocity_other = -3;
  p->max_velocity_ego = 2;
  p->min_velocity_ego = -1;
  p->chain_length = 21;
  p->ego_goal_pos = 12;
  p->num_other_actions = 7;
  p->reward_collision = -1000.0;
  p->reward_goal_reached = 100.0;
  p->reward_step = 0.0;
}

// SimpleState(length): winning 10, loosing -1, reference test/uct/simple_state.h:16
void mamcts_default_simple_params(mamcts_simple_params* p) {
  p->winning_state_length = 10;
  p->loosing_state_length = -1;
}

// UctStatistic's widening order (reference mcts/statistics/uct_statistic.h:61-69): every statistic
// owns std::mt19937(RANDOM_SEED) (mcts/random_generator.h:16) and draws
// uniform_int_distribution<ActionIdx>(0, unexpanded.size()-1), erasing the drawn slot. The
// distribution's algorithm is implementation-defined, so it is evaluated here with the host's
// libstdc++ (the same library the reference is built against) and never re-derived on the device.
int mamcts_reference_expand_order(uint32_t random_seed, uint32_t num_actions, uint32_t* order) {
  if (!order || num_actions == 0 || num_actions > MAMCTS_MAX_ACTIONS)
    return fail(nullptr, MAMCTS_ERR_INVALID, "mamcts_reference_expand_order: bad num_actions");
  std::mt19937 gen(random_seed);
  std::vector<size_t> unexpanded(num_actions);
  std::iota(unexpanded.begin(), unexpanded.end(), 0);
  for (uint32_t k = 0; k < num_actions; ++k) {
    std::uniform_int_distribution<size_t> pick(0, unexpanded.size() - 1);
    const size_t idx = pick(gen);
    order[k] = static_cast<uint32_t>(unexpanded[idx]);
    unexpanded.erase(unexpanded.begin() + idx);
  }
  return MAMCTS_OK;
}

}  // extern "C"
