// staging.h - mirrors caller HOST buffers on the device for one C-ABI call.
//
// MAMCTS_MEM_DEVICE arguments pass through untouched (no copies, fully asynchronous).
// MAMCTS_MEM_HOST arguments are copied host->device on the engine stream before the kernels and
// device->host after them; the call then synchronises so results are ready on return (the
// reference's calls are synchronous). Pinned caller buffers make both copies true DMA transfers.
#pragma once
#include <vector>

#include "engine.h"

namespace mamcts {

class ArgStager {
 public:
  explicit ArgStager(mamcts_engine* e) : e_(e) {}

  // Plans an input (copied in) or output (copied out) of `bytes`; returns a ticket.
  int plan(const void* host_or_dev, size_t bytes, int mem, bool is_output) {
    Item it;
    it.user = const_cast<void*>(host_or_dev);
    it.bytes = bytes;
    it.mem = mem;
    it.out = is_output;
    it.offset = 0;
    if (it.user && mem == MAMCTS_MEM_HOST) {
      it.offset = total_;
      total_ += (bytes + 255) & ~size_t(255);
    }
    items_.push_back(it);
    return static_cast<int>(items_.size()) - 1;
  }

  // Allocates the staging area and issues the host->device copies of the inputs.
  int commit() {
    if (total_ == 0) return MAMCTS_OK;
    int rc = ensure_staging(e_, total_);
    if (rc != MAMCTS_OK) return rc;
    for (const Item& it : items_) {
      if (!it.user || it.mem != MAMCTS_MEM_HOST || it.out) continue;
      MAMCTS_CUDA(e_, cudaMemcpyAsync(static_cast<char*>(e_->staging) + it.offset, it.user, it.bytes,
                                      cudaMemcpyHostToDevice, e_->stream));
      h2d_bytes_ += it.bytes;
    }
    return MAMCTS_OK;
  }

  template <class T>
  T* dev(int ticket) const {
    const Item& it = items_[ticket];
    if (!it.user) return nullptr;
    if (it.mem == MAMCTS_MEM_HOST) return reinterpret_cast<T*>(static_cast<char*>(e_->staging) + it.offset);
    return reinterpret_cast<T*>(it.user);
  }

  // Issues the device->host copies of the outputs and, if any host buffer was involved, waits.
  int finish() {
    bool any_host = false;
    for (const Item& it : items_) {
      if (!it.user || it.mem != MAMCTS_MEM_HOST) continue;
      any_host = true;
      if (!it.out) continue;
      MAMCTS_CUDA(e_, cudaMemcpyAsync(it.user, static_cast<char*>(e_->staging) + it.offset, it.bytes,
                                      cudaMemcpyDeviceToHost, e_->stream));
      d2h_bytes_ += it.bytes;
    }
    if (any_host) MAMCTS_CUDA(e_, cudaStreamSynchronize(e_->stream));
    return MAMCTS_OK;
  }

  size_t h2d_bytes() const { return h2d_bytes_; }
  size_t d2h_bytes() const { return d2h_bytes_; }

 private:
  struct Item {
    void* user;
    size_t bytes;
    int mem;
    bool out;
    size_t offset;
  };
  mamcts_engine* e_;
  std::vector<Item> items_;
  size_t total_ = 0;
  size_t h2d_bytes_ = 0;
  size_t d2h_bytes_ = 0;
};

}  // namespace mamcts
