// TEST INFRASTRUCTURE: runtime of the CUDA-on-CPU shim (see cuda_emu.h).
#include "cuda_emu.h"

namespace emu {
thread_local emu_idx t_idx;
thread_local int t_lin;
emu_idx b_idx;
dim3 b_dim, g_dim;
Block* blk;
unsigned char* dyn_smem;
std::mutex atomic_lock;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nthr = (int)(block.x * block.y * block.z);
  const int nwarp = (nthr + 31) / 32;
  std::vector<unsigned char> sm(smem + 64, 0);
  dyn_smem = sm.data();
  b_dim = block; g_dim = grid;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        b_idx = emu_idx{bx, by, bz};
        Block B;
        B.bar = std::make_unique<std::barrier<>>(nthr);
        for (int w = 0; w < nwarp; ++w) B.wbar.push_back(std::make_unique<std::barrier<>>(std::min(32, nthr - 32 * w)));
        B.xch.assign(nthr, 0);
        B.frag.assign((size_t)nthr * 6, 0);
        blk = &B;
        std::vector<std::thread> ths;
        ths.reserve(nthr);
        for (int t = 0; t < nthr; ++t)
          ths.emplace_back([&, t] {
            t_lin = t;
            t_idx = emu_idx{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
            body();
            // a thread that has left the kernel no longer takes part in barriers
            B.wbar[t >> 5]->arrive_and_drop();
            B.bar->arrive_and_drop();
          });
        for (auto& th : ths) th.join();
      }
}
}  // namespace emu

// plumbing the kernels' host wrappers expect from api.cu
#include <string>
namespace ftc {
static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }
void count_launch() {}
int pdl_enabled() { return 0; }
}  // namespace ftc
extern "C" const char* ftc_last_error(void) { return ftc::g_err.c_str(); }
