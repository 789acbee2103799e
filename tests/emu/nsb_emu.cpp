/* tests/emu/nsb_emu.cpp -- TEST INFRASTRUCTURE ONLY. See nsb_emu.h. */
#include "nsb_emu.h"

namespace nsbemu {

State g;

static std::vector<Fiber> pool;
static const size_t kStack = 256 * 1024;

void yield() { swapcontext(&g.cur->ctx, &g.sched); }

static void trampoline() {
    (*g.body)();
    Fiber *f = g.cur;
    f->done = true;
    /* exited threads no longer take part in barriers */
    g.live--;
    g.warp_live[f->warp]--;
    if (g.live > 0 && g.block_arrived == g.live) { g.block_arrived = 0; g.block_gen++; }
    if (g.warp_live[f->warp] > 0 && g.warp_arrived[f->warp] == g.warp_live[f->warp]) {
        g.warp_arrived[f->warp] = 0;
        g.warp_gen[f->warp]++;
    }
    swapcontext(&f->ctx, &g.sched);
}

void barrier() {
    const unsigned gen = g.block_gen;
    if (++g.block_arrived == g.live) { g.block_arrived = 0; g.block_gen++; return; }
    while (g.block_gen == gen) yield();
}

/* bar.sync id, nthreads: every participant names the same count; exited threads do not take part */
void named_barrier(int id, int nthreads) {
    const unsigned gen = g.named_gen[id];
    if (++g.named_arrived[id] == nthreads) { g.named_arrived[id] = 0; g.named_gen[id]++; return; }
    while (g.named_gen[id] == gen) yield();
}

void warp_barrier() {
    const int w = g.cur->warp;
    const unsigned gen = g.warp_gen[w];
    if (++g.warp_arrived[w] == g.warp_live[w]) { g.warp_arrived[w] = 0; g.warp_gen[w]++; return; }
    while (g.warp_gen[w] == gen) yield();
}

double shfl(double v, int src_lane) {
    const int w = g.cur->warp, lane = g.cur->lane;
    g.warp_buf[w * 32 + lane] = v;
    warp_barrier();
    const double r = g.warp_buf[w * 32 + (src_lane & 31)];
    warp_barrier();
    return r;
}

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
    const unsigned nwarps = (nthreads + 31) / 32;
    g.warp_live.assign(nwarps, 0);
    g.warp_arrived.assign(nwarps, 0);
    g.warp_gen.assign(nwarps, 0);
    g.warp_buf.assign((size_t)nwarps * 32, 0.0);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g.bid = {bx, by, bz};
                unsigned t = 0;
                std::fill(g.warp_live.begin(), g.warp_live.end(), 0);
                std::fill(g.warp_arrived.begin(), g.warp_arrived.end(), 0);
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx, ++t) {
                            Fiber &f = pool[t];
                            f.tid = {tx, ty, tz};
                            f.done = false;
                            f.warp = (int)(t / 32);
                            f.lane = (int)(t % 32);
                            g.warp_live[f.warp]++;
                            getcontext(&f.ctx);
                            f.ctx.uc_stack.ss_sp = f.stack;
                            f.ctx.uc_stack.ss_size = kStack;
                            f.ctx.uc_link = nullptr;
                            makecontext(&f.ctx, trampoline, 0);
                        }
                g.live = (int)nthreads;
                g.block_arrived = 0;
                for (int k = 0; k < 16; ++k) { g.named_arrived[k] = 0; g.named_gen[k] = 0; }
                unsigned guard = 0;
                while (g.live > 0) {
                    for (unsigned k = 0; k < nthreads; ++k) {
                        Fiber &f = pool[reverse ? nthreads - 1 - k : k];
                        if (f.done) continue;
                        g.cur = &f;
                        swapcontext(&g.sched, &f.ctx);
                    }
                    if (++guard > 100000000u) { fprintf(stderr, "nsbemu: barrier deadlock\n"); abort(); }
                }
            }
}

}  // namespace nsbemu
