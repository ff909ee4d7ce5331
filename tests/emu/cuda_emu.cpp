// TEST INFRASTRUCTURE ONLY — fiber scheduler behind tests/emu/cuda_emu.h.
#include "cuda_emu.h"

namespace emu {

uint3_emu threadIdx_, blockIdx_;
dim3 blockDim_, gridDim_;
char* dyn_smem = nullptr;

namespace {
const size_t STACK = 256 * 1024;
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
    int wait = 0;       // 0 runnable, 1 block barrier, 2 warp barrier
    unsigned wait_gen = 0;
    uint3_emu tid;
};
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
int cur = -1;
int nthreads = 0;
unsigned block_gen = 0, block_cnt = 0, block_done = 0;
struct WarpState {
    unsigned gen = 0, cnt = 0, done = 0, size = 0;
    uint32_t buf[32];
    int pred[32];
};
std::vector<WarpState> warps;
const std::function<void()>* body_ptr = nullptr;

void yield_to_sched() {
    int me = cur;
    swapcontext(&fibers[me].ctx, &sched_ctx);
    threadIdx_ = fibers[me].tid;
}
void fiber_main() {
    (*body_ptr)();
    Fiber& f = fibers[cur];
    f.done = true;
    // a finished thread no longer takes part in barriers
    block_done++;
    WarpState& w = warps[cur / 32];
    w.done++;
    if (block_cnt && block_cnt == (unsigned)nthreads - block_done) {
        block_cnt = 0;
        block_gen++;
    }
    if (w.cnt && w.cnt == w.size - w.done) {
        w.cnt = 0;
        w.gen++;
    }
    swapcontext(&f.ctx, &sched_ctx);
}
}  // namespace

void sync_block() {
    block_cnt++;
    if (block_cnt == (unsigned)nthreads - block_done) {
        block_cnt = 0;
        block_gen++;
        return;
    }
    Fiber& f = fibers[cur];
    f.wait = 1;
    f.wait_gen = block_gen;
    yield_to_sched();
}
void sync_warp() {
    WarpState& w = warps[cur / 32];
    w.cnt++;
    if (w.cnt == w.size - w.done) {
        w.cnt = 0;
        w.gen++;
        return;
    }
    Fiber& f = fibers[cur];
    f.wait = 2;
    f.wait_gen = w.gen;
    yield_to_sched();
}
uint32_t shfl(uint32_t v, int src_lane) {
    WarpState& w = warps[cur / 32];
    int lane = cur % 32;
    w.buf[lane] = v;
    sync_warp();
    uint32_t r = w.buf[src_lane < (int)w.size ? src_lane : lane];
    sync_warp();
    return r;
}
uint32_t ballot(int pred) {
    WarpState& w = warps[cur / 32];
    int lane = cur % 32;
    w.pred[lane] = pred;
    sync_warp();
    uint32_t r = 0;
    for (unsigned i = 0; i < w.size; i++)
        if (w.pred[i]) r |= 1u << i;
    sync_warp();
    return r;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    nthreads = block.x * block.y * block.z;
    if ((int)fibers.size() < nthreads) {
        size_t old = fibers.size();
        fibers.resize(nthreads);
        for (size_t i = old; i < fibers.size(); i++) fibers[i].stack = (char*)malloc(STACK);
    }
    std::vector<char> smem_buf(smem + 64);
    dyn_smem = smem_buf.data();
    blockDim_ = block;
    gridDim_ = grid;
    body_ptr = &body;
    int nwarps = (nthreads + 31) / 32;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx_ = uint3_emu{bx, by, bz};
                block_gen = block_cnt = block_done = 0;
                warps.assign(nwarps, WarpState());
                for (int w = 0; w < nwarps; w++) warps[w].size = std::min(32, nthreads - 32 * w);
                for (int t = 0; t < nthreads; t++) {
                    Fiber& f = fibers[t];
                    f.done = false;
                    f.wait = 0;
                    f.tid = uint3_emu{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = f.stack;
                    f.ctx.uc_stack.ss_size = STACK;
                    f.ctx.uc_link = &sched_ctx;
                    makecontext(&f.ctx, fiber_main, 0);
                }
                int remaining = nthreads;
                while (remaining > 0) {
                    bool progressed = false;
                    for (int t = 0; t < nthreads; t++) {
                        Fiber& f = fibers[t];
                        if (f.done) continue;
                        if (f.wait == 1 && f.wait_gen == block_gen) continue;
                        if (f.wait == 2 && f.wait_gen == warps[t / 32].gen) continue;
                        f.wait = 0;
                        cur = t;
                        threadIdx_ = f.tid;
                        swapcontext(&sched_ctx, &f.ctx);
                        progressed = true;
                        if (f.done) remaining--;
                    }
                    if (!progressed) {
                        fprintf(stderr, "emu: deadlock in block (%u,%u,%u)\n", bx, by, bz);
                        abort();
                    }
                }
            }
    cur = -1;
    dyn_smem = nullptr;
}

}  // namespace emu
