// pb_simt_emu.cpp — TEST INFRASTRUCTURE ONLY (see pb_simt_emu.h).
#include "pb_simt_emu.h"
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

namespace pb_emu {

BlockCtx* g_blk = nullptr;
thread_local uint3_ t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    const long nblocks = (long)grid.x * grid.y * grid.z;
    if (nthreads <= 0 || nblocks <= 0) return;
    BlockCtx ctx;
    ctx.nthreads = nthreads;
    ctx.block_bar.reset(nthreads);
    const int nwarps = (nthreads + 31) / 32;
    for (int w = 0; w < nwarps; w++) {
        int width = std::min(32, nthreads - w * 32);
        ctx.warp_bar.push_back(new Barrier(width));
    }
    ctx.xchg.assign((size_t)nwarps * 32, 0);
    std::vector<char> smem(dyn_smem_bytes + 64);
    ctx.dyn_smem = (char*)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
    g_blk = &ctx;

    Barrier all(nthreads);   // keeps the threads of a block in step between blocks
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (int t = 0; t < nthreads; t++) {
        pool.emplace_back([&, t] {
            t_blockDim = block; t_gridDim = grid;
            t_threadIdx = uint3_{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
            for (long b = 0; b < nblocks; b++) {
                t_blockIdx = uint3_{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long)grid.x * grid.y))};
                body();
                all.wait();
            }
        });
    }
    for (auto& th : pool) th.join();
    for (auto* b : ctx.warp_bar) delete b;
    for (auto* b : ctx.named_bar) delete b;
    g_blk = nullptr;
}

}  // namespace pb_emu
