// cuda_sim.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_sim.h).
#include "cuda_sim.h"

thread_local psk_sim_uint3 threadIdx;
psk_sim_uint3 blockIdx, blockDim, gridDim;

namespace psk_sim {

BlockCtx* g_ctx = nullptr;
unsigned char* g_dyn_smem = nullptr;

namespace {

// A persistent pool: worker i runs CUDA thread i of the current block.
class Pool {
 public:
  ~Pool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }

  void run_block(unsigned nthreads, const std::function<void()>& fn) {
    grow(nthreads);
    {
      std::lock_guard<std::mutex> l(mu_);
      fn_ = &fn;
      active_ = nthreads;
      done_ = 0;
      gen_++;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> l(mu_);
    cv_done_.wait(l, [&] { return done_ == th_.size(); });
  }

 private:
  void grow(unsigned n) {
    while (th_.size() < n) {
      unsigned id = (unsigned)th_.size();
      uint64_t start_gen;
      {
        std::lock_guard<std::mutex> l(mu_);
        start_gen = gen_;
      }
      th_.emplace_back([this, id, start_gen] { worker(id, start_gen); });
    }
  }

  void worker(unsigned id, uint64_t seen) {
    for (;;) {
      const std::function<void()>* fn;
      unsigned active;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        fn = fn_;
        active = active_;
      }
      if (id < active) {
        threadIdx.x = id;
        threadIdx.y = threadIdx.z = 0;
        (*fn)();
        // a CUDA thread that has exited no longer takes part in barriers
        g_ctx->warp_bar[id >> 5]->arrive_and_drop();
        g_ctx->block_bar->arrive_and_drop();
      }
      {
        std::lock_guard<std::mutex> l(mu_);
        done_++;
        if (done_ == th_.size()) cv_done_.notify_all();
      }
    }
  }

  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, cv_done_;
  uint64_t gen_ = 0;
  bool stop_ = false;
  const std::function<void()>* fn_ = nullptr;
  unsigned active_ = 0;
  size_t done_ = 0;
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
