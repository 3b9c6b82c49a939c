/* mmg_emu.cpp -- TEST-ONLY SIMT emulator runtime (see mmg_emu.h). */
#include "mmg_emu.h"
#include <stdio.h>
#include <vector>
#include <thread>
#include <atomic>
#include <chrono>
#include <mutex>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
thread_local unsigned char *emu_dyn_smem;
thread_local const char *emu_kernel_name = "?";

/* ---------------- host runtime shim ---------------- */
struct emu_stream_st { int dummy; };
struct emu_event_st { std::chrono::steady_clock::time_point t; };
static const char *emu_err = "no error";
/* MMG_EMU_DEVICES emulated devices (default 1): they share the host's memory, so multi-device host logic can be tested */
cudaError_t cudaGetDeviceCount(int *n) { const char *e = getenv("MMG_EMU_DEVICES"); *n = e ? atoi(e) : 1; return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceDisablePeerAccess(int) { return cudaSuccess; }
cudaError_t cudaMemcpyPeer(void *d, int, const void *s, int, size_t n) { memmove(d, s, n); return cudaSuccess; }
static thread_local int emu_cur_dev = 0;
cudaError_t cudaSetDevice(int d) { emu_cur_dev = d; return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = emu_cur_dev; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { memset(p, 0, sizeof(*p)); const char *e = getenv("MMG_EMU_SMS"); p->multiProcessorCount = e ? atoi(e) : 4; p->totalGlobalMem = (size_t)8 << 30; p->sharedMemPerBlockOptin = 227 * 1024; strcpy(p->name, "mmg-emu"); p->major = 10; return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emu_stream_st(); return cudaSuccess; }
cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new emu_stream_st(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event_st(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event_st(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return emu_err; }

/* ---------------- fibers ---------------- */
extern "C" void emu_swap(void **save_sp, void *load_sp);
asm(R"(
.text
.globl emu_swap
.type emu_swap,@function
emu_swap:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_swap,.-emu_swap
)");

#define EMU_STACK (256 * 1024)

struct Fiber {
	void *sp;
	unsigned char *stack;
	int state; /* 0 runnable, 1 waiting warp, 2 waiting block, 3 done */
	unsigned tid;
	uint64_t coll_result;
};

struct WarpState {
	unsigned arrived;     /* lanes that deposited for the current collective */
	uint64_t vals[32];
	int args[32];
	int op, width;
	unsigned mask;
	unsigned done_lanes;  /* exited lanes */
};

struct BlockCtx {
	~BlockCtx() { for (auto &f : fibers) free(f.stack); }
	std::vector<Fiber> fibers;
	std::vector<WarpState> warps;
	void *sched_sp;
	int cur;
	unsigned n_threads, n_block_wait, n_done;
	const std::function<void()> *body;
};

static thread_local BlockCtx *g_blk;

static void fiber_yield_to_sched(void)
{
	BlockCtx *b = g_blk;
	Fiber *f = &b->fibers[b->cur];
	emu_swap(&f->sp, b->sched_sp);
}

static void fiber_entry(void)
{
	BlockCtx *b = g_blk;
	Fiber *f = &b->fibers[b->cur];
	(*b->body)();
	f = &g_blk->fibers[g_blk->cur];
	f->state = 3;
	++b->n_done;
	b->warps[f->tid >> 5].done_lanes |= 1u << (f->tid & 31);
	/* a pending collective may now be complete without this lane: handled by the scheduler */
	fiber_yield_to_sched();
	abort();
}

static void warp_complete(BlockCtx *b, int w)
{
	WarpState *ws = &b->warps[w];
	unsigned part = ws->arrived;
	uint64_t res[32];
	int op = ws->op;
	unsigned ballot = 0;
	uint64_t acc = 0;
	bool first = true;
	for (int l = 0; l < 32; ++l) if (part >> l & 1) {
		uint64_t v = ws->vals[l];
		if (op == EMU_BALLOT) { if (v) ballot |= 1u << l; }
		else if (op == EMU_RED_OR) acc = first ? v : (acc | v);
		else if (op == EMU_RED_AND) acc = first ? v : (acc & v);
		else if (op == EMU_RED_ADD) acc = first ? v : (uint32_t)(acc + v);
		else if (op == EMU_RED_MAX_S) acc = first ? v : ((int64_t)v > (int64_t)acc ? v : acc);
		else if (op == EMU_RED_MIN_S) acc = first ? v : ((int64_t)v < (int64_t)acc ? v : acc);
		else if (op == EMU_RED_MAX_U) acc = first ? v : (v > acc ? v : acc);
		else if (op == EMU_RED_MIN_U) acc = first ? v : (v < acc ? v : acc);
		first = false;
	}
	for (int l = 0; l < 32; ++l) if (part >> l & 1) {
		int width = ws->width, base = l / width * width, src;
		switch (op) {
		case EMU_SHFL_IDX: src = base + (ws->args[l] & (width - 1)); res[l] = (part >> src & 1) ? ws->vals[src] : ws->vals[l]; break;
		case EMU_SHFL_UP: src = l - ws->args[l]; res[l] = (src >= base && (part >> src & 1)) ? ws->vals[src] : ws->vals[l]; break;
		case EMU_SHFL_DOWN: src = l + ws->args[l]; res[l] = (src < base + width && (part >> src & 1)) ? ws->vals[src] : ws->vals[l]; break;
		case EMU_SHFL_XOR: src = l ^ ws->args[l]; res[l] = (src < base + width && src >= base && (part >> src & 1)) ? ws->vals[src] : ws->vals[l]; break;
		case EMU_BALLOT: res[l] = ballot & ws->mask; break;
		case EMU_MATCH_ANY: { unsigned m = 0; for (int q = 0; q < 32; ++q) if ((part >> q & 1) && ws->vals[q] == ws->vals[l]) m |= 1u << q; res[l] = m; break; }
		default: res[l] = acc; break;
		}
	}
	for (int l = 0; l < 32; ++l) if (part >> l & 1) {
		Fiber *f = &b->fibers[w * 32 + l];
		f->coll_result = res[l];
		f->state = 0;
	}
	ws->arrived = 0;
}

uint64_t emu_collective(int op, unsigned mask, uint64_t v, int arg, int width)
{
	BlockCtx *b = g_blk;
	Fiber *f = &b->fibers[b->cur];
	int w = f->tid >> 5, l = f->tid & 31;
	WarpState *ws = &b->warps[w];
	if (ws->arrived == 0) ws->op = op, ws->width = width, ws->mask = mask;
	else if (ws->op != op) { fprintf(stderr, "[emu] divergent collective in %s, block %u warp %d: op %d vs %d\n", emu_kernel_name, blockIdx.x, w, ws->op, op); abort(); }
	ws->vals[l] = v;
	ws->args[l] = arg;
	ws->arrived |= 1u << l;
	f->state = 1;
	{
		unsigned need = ws->mask & ~ws->done_lanes;
		unsigned nthr_mask = (b->n_threads - w * 32 >= 32) ? 0xffffffffu : ((1u << (b->n_threads - w * 32)) - 1);
		need &= nthr_mask;
		if ((ws->arrived & need) == need) { warp_complete(b, w); return f->coll_result; }
	}
	fiber_yield_to_sched();
	return g_blk->fibers[g_blk->cur].coll_result;
}

void __syncwarp(unsigned mask) { (void)emu_collective(EMU_BALLOT, mask, 0, 0, 32); }

void __syncthreads(void)
{
	BlockCtx *b = g_blk;
	Fiber *f = &b->fibers[b->cur];
	f->state = 2;
	++b->n_block_wait;
	if (b->n_block_wait + b->n_done == b->n_threads) {
		for (unsigned i = 0; i < b->n_threads; ++i) if (b->fibers[i].state == 2) b->fibers[i].state = 0;
		b->n_block_wait = 0;
		return;
	}
	fiber_yield_to_sched();
}

static void run_block(BlockCtx *b, unsigned bx, dim3 grid, dim3 block, unsigned char *smem)
{
	unsigned T = block.x;
	g_blk = b;
	b->n_threads = T; b->n_block_wait = 0; b->n_done = 0;
	b->warps.assign((T + 31) / 32, WarpState());
	for (auto &w : b->warps) w.arrived = 0, w.done_lanes = 0;
	if (b->fibers.size() < T) {
		size_t old = b->fibers.size();
		b->fibers.resize(T);
		for (size_t i = old; i < T; ++i) b->fibers[i].stack = (unsigned char*)aligned_alloc(64, EMU_STACK);
	}
	for (unsigned i = 0; i < T; ++i) {
		Fiber *f = &b->fibers[i];
		f->tid = i; f->state = 0;
		uintptr_t top = ((uintptr_t)f->stack + EMU_STACK) & ~(uintptr_t)15;
		void **sp = (void**)top;
		*--sp = 0;                    /* fake return address of fiber_entry's caller */
		*--sp = (void*)fiber_entry;   /* popped by ret */
		for (int k = 0; k < 6; ++k) *--sp = 0;
		f->sp = sp;
	}
	blockIdx.x = bx; blockIdx.y = blockIdx.z = 0;
	blockDim = block; gridDim = grid;
	emu_dyn_smem = smem;
	while (b->n_done < T) {
		bool progressed = false;
		for (unsigned i = 0; i < T; ++i) {
			Fiber *f = &b->fibers[i];
			if (f->state != 0) continue;
			progressed = true;
			b->cur = i;
			threadIdx.x = i; threadIdx.y = threadIdx.z = 0;
			emu_swap(&b->sched_sp, f->sp);
			/* back in the scheduler: if the fiber exited, pending collectives of its warp may be complete */
			if (f->state == 3) {
				int w = i >> 5;
				WarpState *ws = &b->warps[w];
				unsigned nthr_mask = (T - w * 32 >= 32) ? 0xffffffffu : ((1u << (T - w * 32)) - 1);
				unsigned need = ws->mask & ~ws->done_lanes & nthr_mask;
				if (ws->arrived && (ws->arrived & need) == need) {
					warp_complete(b, w);
				}
				if (b->n_block_wait && b->n_block_wait + b->n_done == T) {
					for (unsigned k = 0; k < T; ++k) if (b->fibers[k].state == 2) b->fibers[k].state = 0;
					b->n_block_wait = 0;
				}
			}
		}
		if (!progressed) { fprintf(stderr, "[emu] deadlock in block %u (done %u/%u, block_wait %u)\n", bx, b->n_done, T, b->n_block_wait); abort(); }
	}
}

void emu_launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body)
{
	static int n_os_threads = -1;
	if (n_os_threads < 0) { const char *e = getenv("MMG_EMU_THREADS"); n_os_threads = e ? atoi(e) : 1; if (n_os_threads < 1) n_os_threads = 1; }
	std::atomic<unsigned> next(0);
	auto worker = [&]() {
		static thread_local BlockCtx ctx;
		ctx.body = &body;
		std::vector<unsigned char> smem(dyn_smem + 64);
		for (;;) {
			unsigned bx = next.fetch_add(1);
			if (bx >= grid.x) break;
			run_block(&ctx, bx, grid, block, smem.data());
		}
	};
	if (n_os_threads == 1 || grid.x == 1) worker();
	else {
		std::vector<std::thread> th;
		int n = n_os_threads < (int)grid.x ? n_os_threads : (int)grid.x;
		for (int i = 0; i < n; ++i) th.emplace_back(worker);
		for (auto &t : th) t.join();
	}
}
