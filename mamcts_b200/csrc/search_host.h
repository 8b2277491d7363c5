// search_host.h - the type-erased handle behind the mamcts_search_* entry points of the C ABI, shared by the
// UCT searches (search.cu) and the hypothesis-MCTS search (search_hyp.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "engine.h"
#include "root_merge.h"
#include "search_kernel.cuh"

// Type-erased handle behind the C ABI; one concrete SearchT per compiled (environment, action-count,
// child-index) combination.
struct mamcts_search {
  mamcts_engine* e = nullptr;
  mamcts_search_config cfg;
  mamcts::SearchParams p;
  mamcts::SearchTables tab;  // host copy
  uint32_t A = 0, state_ints = 0;
  void* d_nodes = nullptr;
  uint32_t* d_hash = nullptr;
  uint32_t* d_tree_nodes = nullptr;
  uint32_t* d_tree_iters = nullptr;
  mamcts::SearchTables* d_tab = nullptr;
  double* d_log = nullptr;
  double* d_disc = nullptr;
  unsigned long long* d_counters = nullptr;
  int32_t* d_root_x = nullptr;
  // compact (two-tier) layout, search_compact.cuh
  void* d_inner = nullptr;
  void* d_leaves = nullptr;
  uint32_t* d_tree_inner = nullptr;
  uint32_t* d_tree_leaves = nullptr;
  uint32_t max_inner = 0;
  uint32_t iters_done = 0;
  int block_threads = mamcts::kSearchThreads;

  virtual ~mamcts_search() {}
  virtual size_t node_bytes() const = 0;     // bytes of the record that carries the statistics
  virtual size_t leaf_bytes() const { return 0; }
  virtual bool uses_hash() const = 0;
  virtual bool is_compact() const { return false; }
  // first record (the root) of a tree and the distance between consecutive trees' roots
  virtual const char* root_record(uint32_t tree) const {
    return static_cast<const char*>(d_nodes) + node_bytes() * (size_t)tree * p.max_nodes;
  }
  // canonical depth-first export of one tree (mamcts_search_dump_tree)
  virtual int dump(uint32_t tree, double* records, uint32_t capacity_records, uint32_t* num_records) = 0;
  virtual int launch_reset() = 0;
  virtual int launch_run(uint32_t iterations) = 0;
  // host-side views of a node record copied back from the device
  virtual void read_header(const void* node, uint32_t* parent, uint32_t* depth, uint32_t* visits,
                           uint8_t* ja) const = 0;
  virtual void read_stat(const void* node, uint32_t agent, double* value, double* latest,
                         uint32_t* visits, uint32_t* nexp, double* q, uint32_t* n) const = 0;
  virtual void root_gather(mamcts::RootGather* g) const = 0;
  // new decision state; the default copies root_x into d_root_x (UCT searches ignore the last actions)
  virtual int upload_root(const int32_t* root_x, const int32_t* root_last_action);
  // device-side failures a read-back must report instead of returning statistics of an unfinished search
  virtual int health();
  // buffers a subclass owns beyond the ones below
  virtual int alloc_extra(uint32_t /*max_nodes*/) { return MAMCTS_OK; }
  virtual void free_extra() {}
};

namespace mamcts {
void free_search(mamcts_search* s);
}  // namespace mamcts

