// emu_runtime.cpp -- TEST INFRASTRUCTURE: runs the CUDA kernels of tiberate_fhe_b200/csrc on host
// threads (one std::thread per CUDA thread of a block, blocks executed one after another) so the
// kernels' indexing logic can be checked against the oracle in a container without a GPU.
// Never loaded by the product; see tests/emu/README in tests/test_emu_*.py docstrings.
#include <barrier>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "tb200_platform.h"

thread_local emu_uint3 threadIdx;
thread_local emu_uint3 blockIdx;
thread_local dim3 blockDim;
thread_local dim3 gridDim;

static thread_local std::barrier<>* t_bar = nullptr;

void emu_syncthreads() { t_bar->arrive_and_wait(); }

void emu_launch(dim3 grid, dim3 block, const std::function<void()>& body) {
  const unsigned nt = block.x * block.y * block.z;
  std::barrier<> bar((std::ptrdiff_t)nt);
  std::vector<std::thread> pool;
  pool.reserve(nt);
  for (unsigned t = 0; t < nt; ++t) {
    pool.emplace_back([&, t]() {
      t_bar = &bar;
      blockDim = block;
      gridDim = grid;
      threadIdx.x = t % block.x;
      threadIdx.y = (t / block.x) % block.y;
      threadIdx.z = t / (block.x * block.y);
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx.x = bx;
            blockIdx.y = by;
            blockIdx.z = bz;
            body();
            bar.arrive_and_wait();  // block boundary: static "shared" arrays are reused
          }
    });
  }
  for (auto& th : pool) th.join();
}
