// root_merge.h - shared between backup.cu (mamcts_root_merge) and search.cu (device-resident trees).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "engine.h"

namespace mamcts {

// acc layout (doubles): [0,nA) sum_q, [nA,2nA) sum_n, [2nA,3nA) occurrences, [3nA] sum_visits.
// A single block reduces in a fixed order (thread-strided partials, then a shared-memory tree), so
// the result does not depend on scheduling.
struct RootGather {
  const double* q;        // action values
  const void* n;          // action counts (present <=> n > 0 at iteration boundaries), count_bytes wide
  const void* visits;
  uint32_t count_bytes;   // 4 (uint32) or 2 (uint16, compact layout of short searches)
  const uint32_t* root_slot;  // may be null: strided layout below (device-resident trees)
  size_t stride_q;            // doubles between consecutive trees' q rows when root_slot is null
  size_t stride_n;            // uint32s between consecutive trees' n rows
  size_t stride_v;            // uint32s between consecutive trees' visit counters
  uint32_t max_actions;
};


int root_merge_device(mamcts_engine* e, const RootGather& g, uint32_t n_trees, uint32_t nA,
                      double* merged_q, uint64_t* merged_n, uint32_t* merged_occ,
                      uint64_t* merged_total_visits, double* merged_value, uint32_t* best_action);

}  // namespace mamcts
