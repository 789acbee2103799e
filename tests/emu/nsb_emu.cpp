/* tests/emu/nsb_emu.cpp -- TEST INFRASTRUCTURE ONLY. See nsb_emu.h. */
#include "nsb_emu.h"

namespace nsbemu {

State g;

static std::vector<Fiber> pool;
static const size_t kStack = 256 * 1024;

static void trampoline() {
    (*g.body)();
    g.cur->done = true;
    swapcontext(&g.cur->ctx, &g.sched);
}

void barrier() { swapcontext(&g.cur->ctx, &g.sched); }

void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> body) {
    const unsigned nthreads = block.x * block.y * block.z;
    if (pool.size() < nthreads) {
        size_t old = pool.size();
        pool.resize(nthreads);
        for (size_t k = old; k < nthreads; ++k) pool[k].stack = (char *)malloc(kStack);
    }
    static const bool reverse = getenv("NSB_EMU_REVERSE") && atoi(getenv("NSB_EMU_REVERSE"));
    std::vector<unsigned char> smem(smem_bytes + 64);
    g.smem = (unsigned char *)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
    g.bdim = block;
    g.gdim = grid;
    g.body = &body;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g.bid = {bx, by, bz};
                unsigned t = 0;
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx, ++t) {
                            Fiber &f = pool[t];
                            f.tid = {tx, ty, tz};
                            f.done = false;
                            getcontext(&f.ctx);
                            f.ctx.uc_stack.ss_sp = f.stack;
                            f.ctx.uc_stack.ss_size = kStack;
                            f.ctx.uc_link = nullptr;
                            makecontext(&f.ctx, trampoline, 0);
                        }
                unsigned remaining = nthreads;
                while (remaining) {
                    for (unsigned k = 0; k < nthreads; ++k) {
                        Fiber &f = pool[reverse ? nthreads - 1 - k : k];
                        if (f.done) continue;
                        g.cur = &f;
                        swapcontext(&g.sched, &f.ctx);
                        if (f.done) --remaining;
                    }
                }
            }
}

}  // namespace nsbemu
