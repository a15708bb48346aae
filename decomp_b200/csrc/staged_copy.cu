// Host-side staging of pageable input arrays (the reference API receives ordinary numpy arrays, decomp/lasso.py:19,
// decomp/nmf.py:16): a copy from pageable memory through the driver's own bounce buffer runs on the calling thread
// at ~11 GB/s on the B200 boxes.  Here a few threads memcpy pieces of the array into a ring of page-locked slots the
// caller supplies while the calling thread hands finished slots to the copy engine (cudaMemcpyAsync on the caller's
// stream), so the link sees close to what the host's memcpy threads deliver.  Nothing is allocated for the caller;
// no Python object is touched, so a Python host can call this from a helper thread with the GIL released.
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "common.h"

using namespace dcp;

extern "C" {

int decomp_staged_upload(void* dst_device, const void* src_host, size_t bytes, void* pinned_ring, size_t slot_bytes,
                         int32_t slots, int32_t threads, void* stream) {
  if (bytes == 0) return DECOMP_OK;
  if (dst_device == nullptr || src_host == nullptr || pinned_ring == nullptr || slot_bytes == 0 || slots < 2 ||
      threads < 1) {
    set_error("decomp_staged_upload: invalid argument");
    return DECOMP_ERR_INVALID;
  }
  cudaStream_t st = as_stream(stream);
  const size_t pieces = (bytes + slot_bytes - 1) / slot_bytes;
  if ((size_t)threads > pieces) threads = (int32_t)pieces;
  if ((size_t)slots > pieces) slots = (int32_t)pieces;

  std::vector<cudaEvent_t> drained(slots);      // the async copy out of a slot has finished
  for (int s = 0; s < slots; ++s) {
    cudaError_t e = cudaEventCreateWithFlags(&drained[s], cudaEventDisableTiming);
    if (e != cudaSuccess) {
      for (int t = 0; t < s; ++t) cudaEventDestroy(drained[t]);
      return check_cuda(e, "decomp_staged_upload: cudaEventCreate");
    }
  }
  // filled[p]: piece p sits in its slot; issued: pieces whose copy (and drained event) has been enqueued
  std::vector<std::atomic<int>> filled(pieces);
  for (size_t p = 0; p < pieces; ++p) filled[p].store(0, std::memory_order_relaxed);
  std::atomic<size_t> next(0), issued(0);
  std::atomic<int> failed(0);
  int device = 0;
  cudaGetDevice(&device);

  auto worker = [&]() {
    cudaSetDevice(device);
    for (;;) {
      const size_t p = next.fetch_add(1);
      if (p >= pieces || failed.load()) return;
      const int slot = (int)(p % (size_t)slots);
      if (p >= (size_t)slots) {
        // the slot is free once the copy of piece p - slots has been enqueued and has drained
        while (issued.load(std::memory_order_acquire) <= p - (size_t)slots) {
          if (failed.load()) return;
          std::this_thread::yield();
        }
        if (cudaEventSynchronize(drained[slot]) != cudaSuccess) {
          failed.store(1);
          return;
        }
      }
      const size_t off = p * slot_bytes;
      const size_t n = bytes - off < slot_bytes ? bytes - off : slot_bytes;
      memcpy(static_cast<char*>(pinned_ring) + (size_t)slot * slot_bytes, static_cast<const char*>(src_host) + off, n);
      filled[p].store(1, std::memory_order_release);
    }
  };
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t) pool.emplace_back(worker);

  cudaError_t err = cudaSuccess;
  for (size_t p = 0; p < pieces && err == cudaSuccess; ++p) {
    while (!filled[p].load(std::memory_order_acquire)) {
      if (failed.load()) break;
      std::this_thread::yield();
    }
    if (failed.load()) break;
    const int slot = (int)(p % (size_t)slots);
    const size_t off = p * slot_bytes;
    const size_t n = bytes - off < slot_bytes ? bytes - off : slot_bytes;
    err = cudaMemcpyAsync(static_cast<char*>(dst_device) + off, static_cast<char*>(pinned_ring) + (size_t)slot * slot_bytes,
                          n, cudaMemcpyHostToDevice, st);
    if (err == cudaSuccess) err = cudaEventRecord(drained[slot], st);
    issued.store(p + 1, std::memory_order_release);
  }
  if (err != cudaSuccess) failed.store(1);
  for (auto& t : pool) t.join();
  // the ring is the caller's and may be reused right away: wait until the last copies have drained
  for (int s = 0; s < slots; ++s) {
    if (err == cudaSuccess && !failed.load()) {
      cudaError_t e = cudaEventSynchronize(drained[s]);
      if (e != cudaSuccess) err = e;
    }
    cudaEventDestroy(drained[s]);
  }
  if (err != cudaSuccess) return check_cuda(err, "decomp_staged_upload");
  if (failed.load()) {
    set_error("decomp_staged_upload: a staging thread failed");
    return DECOMP_ERR_CUDA;
  }
  return DECOMP_OK;
}

}  // extern "C"
