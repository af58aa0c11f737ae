// TEST INFRASTRUCTURE -- not part of the product.  See include/cuda_runtime.h.
//
// Cooperative execution of one CUDA thread block: every thread is a ucontext fiber on its own
// stack; a fiber runs until it finishes or reaches a scheduling point (__syncthreads*, a warp
// collective).  A collective completes when every LIVE lane named in its mask has arrived with
// the same mask; the scheduler then computes each lane's result, so lanes never read each
// other's operands after they may have moved on.  Exited threads count as arrived (what the
// hardware does for exited lanes).  No runnable fiber and nothing to release = a divergent
// barrier in the kernel: abort with a message.
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <ucontext.h>

#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace emu {
namespace {

constexpr size_t kStackBytes = 256 << 10;
enum Wait { kRun, kBlock, kWarp };

struct Lane {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = true;
    Wait wait = kRun;
    // pending collective
    Op op;
    unsigned mask;
    uint64_t value;
    int param;
    uint64_t result;
    // pending block barrier
    int pred, mode;
    uint3 tid;
};

std::vector<Lane*> g_lanes;
ucontext_t g_sched;
Lane* g_cur = nullptr;
const std::function<void()>* g_body = nullptr;
int g_block_result = 0;

void lane_entry()
{
    (*g_body)();
    g_cur->done = true;
    swapcontext(&g_cur->ctx, &g_sched);
    abort();    // a finished fiber is never resumed
}

void yield()
{
    Lane* me = g_cur;
    swapcontext(&me->ctx, &g_sched);
    threadIdx = me->tid;
}

[[noreturn]] void die(const char* what)
{
    fprintf(stderr, "cuda-emu: %s (block %u,%u,%u)\n", what, blockIdx.x, blockIdx.y, blockIdx.z);
    abort();
}

// Completes the warp collectives whose participants have all arrived.  Returns true if any lane
// was released.
bool resolve_warps(int n_threads)
{
    bool released = false;
    for (int w0 = 0; w0 < n_threads; w0 += 32) {
        const int wn = n_threads - w0 < 32 ? n_threads - w0 : 32;
        unsigned live = 0, waiting = 0;
        for (int l = 0; l < wn; ++l) {
            Lane* L = g_lanes[w0 + l];
            if (!L->done) live |= 1u << l;
            if (!L->done && L->wait == kWarp) waiting |= 1u << l;
        }
        unsigned todo = waiting;
        while (todo) {
            const int lead = __builtin_ctz(todo);
            Lane* A = g_lanes[w0 + lead];
            const unsigned group = A->mask & live;
            todo &= ~group;
            if ((group & waiting) != group) continue;            // someone has not arrived yet
            bool same = true;
            for (int l = 0; l < wn; ++l)
                if (group >> l & 1u) {
                    Lane* L = g_lanes[w0 + l];
                    if (L->mask != A->mask || L->op != A->op) same = false;
                }
            if (!same) continue;                                  // sub-groups with other masks: wait for them
            // results
            uint64_t ball = 0, all = 1, any = 0, red = 0;
            bool first = true;
            for (int l = 0; l < wn; ++l)
                if (group >> l & 1u) {
                    Lane* L = g_lanes[w0 + l];
                    if (L->value) { ball |= 1ull << l; any = 1; } else all = 0;
                    if (A->op == kReduceMax || A->op == kReduceMin) {
                        const bool sgn = L->param != 0;
                        const bool better = first ||
                            (A->op == kReduceMax ? (sgn ? (int64_t)L->value > (int64_t)red : L->value > red)
                                                 : (sgn ? (int64_t)L->value < (int64_t)red : L->value < red));
                        if (better) red = L->value;
                    } else if (A->op == kReduceAdd) {
                        red += L->value;
                    }
                    first = false;
                }
            for (int l = 0; l < wn; ++l) {
                if (!(group >> l & 1u)) continue;
                Lane* L = g_lanes[w0 + l];
                int src = l;
                switch (L->op) {
                case kShflIdx: src = L->param & 31; break;
                case kShflXor: src = l ^ L->param; break;
                case kShflUp: src = l - L->param; break;
                case kShflDown: src = l + L->param; break;
                default: break;
                }
                switch (L->op) {
                case kShflIdx: case kShflXor: case kShflUp: case kShflDown:
                    // a source lane outside the group (or the warp) returns the caller's own value
                    L->result = (src >= 0 && src < wn && (group >> src & 1u)) ? g_lanes[w0 + src]->value : L->value;
                    break;
                case kMatchAny: {
                    unsigned m = 0;
                    for (int k = 0; k < wn; ++k)
                        if ((group >> k & 1u) && g_lanes[w0 + k]->value == L->value) m |= 1u << k;
                    L->result = m;
                    break;
                }
                case kAll: L->result = all; break;
                case kAny: L->result = any; break;
                case kBallot: L->result = ball; break;
                case kReduceMax: case kReduceMin: case kReduceAdd: L->result = red; break;
                case kSyncWarp: L->result = 0; break;
                }
            }
            for (int l = 0; l < wn; ++l)
                if (group >> l & 1u) g_lanes[w0 + l]->wait = kRun;
            released = true;
        }
    }
    return released;
}

bool resolve_block(int n_threads)
{
    int live = 0, waiting = 0, acc_or = 0, acc_and = 1, count = 0, mode = -1;
    for (int t = 0; t < n_threads; ++t) {
        Lane* L = g_lanes[t];
        if (L->done) continue;
        ++live;
        if (L->wait == kBlock) {
            ++waiting;
            acc_or |= L->pred != 0;
            acc_and &= L->pred != 0;
            count += L->pred != 0;
            if (mode >= 0 && mode != L->mode) die("threads of a block wait at different kinds of barrier");
            mode = L->mode;
        }
    }
    if (live == 0 || waiting != live) return false;
    g_block_result = mode == 1 ? acc_or : mode == 2 ? acc_and : mode == 3 ? count : 0;
    for (int t = 0; t < n_threads; ++t)
        if (!g_lanes[t]->done) g_lanes[t]->wait = kRun;
    return true;
}

void run_block(int n_threads)
{
    while ((int)g_lanes.size() < n_threads) {
        Lane* L = new Lane;
        void* s = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK, -1, 0);
        if (s == MAP_FAILED) die("cannot map a fiber stack");
        L->stack = (char*)s;
        g_lanes.push_back(L);
    }
    for (int t = 0; t < n_threads; ++t) {
        Lane* L = g_lanes[t];
        getcontext(&L->ctx);
        L->ctx.uc_stack.ss_sp = L->stack;
        L->ctx.uc_stack.ss_size = kStackBytes;
        L->ctx.uc_link = nullptr;
        makecontext(&L->ctx, lane_entry, 0);
        L->done = false;
        L->wait = kRun;
        L->tid.x = (unsigned)t % blockDim.x;
        L->tid.y = (unsigned)t / blockDim.x % blockDim.y;
        L->tid.z = (unsigned)t / (blockDim.x * blockDim.y);
    }
    int remaining = n_threads;
    while (remaining > 0) {
        bool progressed = false;
        for (int t = 0; t < n_threads; ++t) {
            Lane* L = g_lanes[t];
            if (L->done || L->wait != kRun) continue;
            g_cur = L;
            threadIdx = L->tid;
            swapcontext(&g_sched, &L->ctx);
            progressed = true;
            if (L->done) --remaining;
        }
        const bool w = resolve_warps(n_threads);
        const bool b = resolve_block(n_threads);
        if (!progressed && !w && !b && remaining > 0)
            die("deadlock: a barrier or warp collective that not every participating thread reaches");
    }
    g_cur = nullptr;
}

}  // namespace

void launch(dim3 grid, dim3 block, const std::function<void()>& body)
{
    if (g_cur) die("nested launch");
    const int n_threads = (int)(block.x * block.y * block.z);
    if (n_threads <= 0 || n_threads > 1024) die("bad block size");
    gridDim = grid;
    blockDim = block;
    g_body = &body;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
                run_block(n_threads);
            }
    g_body = nullptr;
}

static unsigned long long g_collective_calls[16];

uint64_t collective(Op op, unsigned mask, uint64_t value, int param)
{
    Lane* me = g_cur;
    ++g_collective_calls[(int)op & 15];
    const unsigned lane = (me->tid.x + blockDim.x * (me->tid.y + blockDim.y * me->tid.z)) & 31u;
    if (!(mask >> lane & 1u)) die("a lane calls a warp collective with a mask that does not name it");
    me->op = op; me->mask = mask; me->value = value; me->param = param;
    me->wait = kWarp;
    yield();
    return me->result;
}

int block_barrier(int pred, int mode)
{
    Lane* me = g_cur;
    me->pred = pred; me->mode = mode;
    me->wait = kBlock;
    yield();
    return g_block_result;
}

void spin()
{
    Lane* me = g_cur;
    me->wait = kRun;        // still runnable: resumed on the scheduler's next round, after the others
    yield();
}

void misaligned(const void* p, size_t a)
{
    fprintf(stderr, "cuda-emu: misaligned access %p (needs %zu-byte alignment)\n", p, a);
    abort();
}

}  // namespace emu

// per-op count of warp collectives executed so far (tests use it to see that a code path was taken)
extern "C" unsigned long long emu_collective_calls(int op) { return emu::g_collective_calls[op & 15]; }
