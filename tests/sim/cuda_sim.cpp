// cuda_sim.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_sim.h).
#include "cuda_sim.h"

#include <semaphore>

thread_local psk_sim_uint3 threadIdx;
psk_sim_uint3 blockIdx, blockDim, gridDim;

namespace psk_sim {

BlockCtx* g_ctx = nullptr;
unsigned char* g_dyn_smem = nullptr;

namespace {

// A persistent pool: worker i runs CUDA thread i of the current block.  Each worker sleeps on
// its own semaphore, so a 32-thread block wakes 32 workers, not the whole pool.
class Pool {
 public:
  void run_block(unsigned nthreads, const std::function<void()>& fn) {
    grow(nthreads);
    fn_ = &fn;
    remaining_.store((int)nthreads, std::memory_order_release);
    for (unsigned i = 0; i < nthreads; ++i) workers_[i]->go.release();
    std::unique_lock<std::mutex> l(mu_);
    cv_done_.wait(l, [&] { return remaining_.load(std::memory_order_acquire) == 0; });
  }

 private:
  struct Worker {
    std::binary_semaphore go{0};
    std::thread th;
  };

  void grow(unsigned n) {
    while (workers_.size() < n) {
      unsigned id = (unsigned)workers_.size();
      workers_.emplace_back(std::make_unique<Worker>());
      Worker* w = workers_.back().get();
      w->th = std::thread([this, id, w] { worker(id, w); });
      w->th.detach();
    }
  }

  void worker(unsigned id, Worker* w) {
    for (;;) {
      w->go.acquire();
      threadIdx.x = id;
      threadIdx.y = threadIdx.z = 0;
      (*fn_)();
      // a CUDA thread that has exited no longer takes part in barriers
      g_ctx->warp_bar[id >> 5]->arrive_and_drop();
      g_ctx->block_bar->arrive_and_drop();
      if (remaining_.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        std::lock_guard<std::mutex> l(mu_);
        cv_done_.notify_all();
      }
    }
  }

  std::vector<std::unique_ptr<Worker>> workers_;
  std::mutex mu_;
  std::condition_variable cv_done_;
  std::atomic<int> remaining_{0};
  const std::function<void()>* fn_ = nullptr;
};

Pool& pool() {
  static Pool* p = new Pool();  // leaked on purpose: workers outlive static destruction
  return *p;
}

std::mutex g_launch_mu;

}  // namespace

void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& fn) {
  std::lock_guard<std::mutex> guard(g_launch_mu);
  std::vector<unsigned char> dyn(smem + 64);
  g_dyn_smem = dyn.data() + (64 - (reinterpret_cast<uintptr_t>(dyn.data()) & 63)) % 64;
  gridDim.x = grid;
  gridDim.y = gridDim.z = 1;
  blockDim.x = block;
  blockDim.y = blockDim.z = 1;
  for (unsigned b = 0; b < grid; b++) {
    BlockCtx ctx;
    ctx.block_bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)block);
    unsigned nwarps = (block + 31) / 32;
    for (unsigned w = 0; w < nwarps; w++) {
      unsigned n = std::min(32u, block - w * 32);
      ctx.warp_bar.emplace_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)n));
    }
    ctx.xchg.assign(block, 0);
    g_ctx = &ctx;
    blockIdx.x = b;
    blockIdx.y = blockIdx.z = 0;
    pool().run_block(block, fn);
    g_ctx = nullptr;
  }
  g_dyn_smem = nullptr;
}

}  // namespace psk_sim
