/* mmg_emu.h -- TEST-ONLY SIMT emulator.
 *
 * Lets the CPU test-suite compile the product's CUDA kernel source
 * (mappy-rs_b200/csrc/*.cu) with g++ and run it lane-for-lane: every CUDA
 * thread is a fiber; warp collectives and __syncthreads() are rendezvous
 * points.  It exists so kernel logic can be checked against the oracle in a
 * container without a GPU.  It is NOT a CPU fallback: the product library
 * (libmmg.so, built by nvcc) contains no CPU mapping path and fails loudly
 * without a device; only tests/ builds and loads the emulated library.
 */
#ifndef MMG_EMU_H
#define MMG_EMU_H

#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <functional>

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
static const int warpSize = 32;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static thread_local
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

typedef int cudaError_t;
typedef struct emu_stream_st *cudaStream_t;
typedef struct emu_event_st *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
struct cudaDeviceProp { int multiProcessorCount; size_t totalGlobalMem; size_t sharedMemPerBlockOptin; char name[256]; int major, minor; };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaDeviceCanAccessPeer(int *can, int dev, int peer);
cudaError_t cudaDeviceEnablePeerAccess(int peer, unsigned flags);
cudaError_t cudaDeviceDisablePeerAccess(int peer);
cudaError_t cudaMemcpyPeer(void *d, int ddev, const void *s, int sdev, size_t n);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t n);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaHostRegister(void *p, size_t n, unsigned flags);
cudaError_t cudaHostUnregister(void *p);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st);
cudaError_t cudaMemset(void *d, int v, size_t n);
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned f);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned f);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetLastError(void);
cudaError_t cudaPeekAtLastError(void);
const char *cudaGetErrorString(cudaError_t e);
template<typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

/* ---- launch ---- */
void emu_launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);
extern thread_local const char *emu_kernel_name;   /* name of the kernel being emulated (diagnostics) */
extern thread_local unsigned char *emu_dyn_smem;

/* ---- rendezvous primitives ---- */
void __syncthreads(void);
void __syncwarp(unsigned mask = 0xffffffffu);
enum { EMU_SHFL_IDX, EMU_SHFL_UP, EMU_SHFL_DOWN, EMU_SHFL_XOR, EMU_BALLOT, EMU_RED_OR, EMU_RED_AND, EMU_RED_MAX_S, EMU_RED_MIN_S, EMU_RED_MAX_U, EMU_RED_MIN_U, EMU_RED_ADD, EMU_MATCH_ANY };
uint64_t emu_collective(int op, unsigned mask, uint64_t v, int arg, int width);

template<typename T> static inline uint64_t emu_to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template<typename T> static inline T emu_from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template<typename T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) { return emu_from_bits<T>(emu_collective(EMU_SHFL_IDX, m, emu_to_bits(v), src, width)); }
template<typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) { return emu_from_bits<T>(emu_collective(EMU_SHFL_UP, m, emu_to_bits(v), (int)d, width)); }
template<typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) { return emu_from_bits<T>(emu_collective(EMU_SHFL_DOWN, m, emu_to_bits(v), (int)d, width)); }
template<typename T> static inline T __shfl_xor_sync(unsigned m, T v, int lm, int width = 32) { return emu_from_bits<T>(emu_collective(EMU_SHFL_XOR, m, emu_to_bits(v), lm, width)); }
static inline unsigned __ballot_sync(unsigned m, int pred) { return (unsigned)emu_collective(EMU_BALLOT, m, pred ? 1 : 0, 0, 32); }
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return (__ballot_sync(m, pred) & m) == m; }
static inline unsigned __reduce_or_sync(unsigned m, unsigned v) { return (unsigned)emu_collective(EMU_RED_OR, m, v, 0, 32); }
static inline unsigned __reduce_and_sync(unsigned m, unsigned v) { return (unsigned)emu_collective(EMU_RED_AND, m, v, 0, 32); }
static inline unsigned __reduce_add_sync(unsigned m, unsigned v) { return (unsigned)emu_collective(EMU_RED_ADD, m, v, 0, 32); }
static inline int __reduce_add_sync(unsigned m, int v) { return (int)(unsigned)emu_collective(EMU_RED_ADD, m, (unsigned)v, 0, 32); }
static inline int __reduce_max_sync(unsigned m, int v) { return (int)(int64_t)emu_collective(EMU_RED_MAX_S, m, (uint64_t)(int64_t)v, 0, 32); }
static inline int __reduce_min_sync(unsigned m, int v) { return (int)(int64_t)emu_collective(EMU_RED_MIN_S, m, (uint64_t)(int64_t)v, 0, 32); }
static inline unsigned __reduce_max_sync(unsigned m, unsigned v) { return (unsigned)emu_collective(EMU_RED_MAX_U, m, v, 0, 32); }
static inline unsigned __reduce_min_sync(unsigned m, unsigned v) { return (unsigned)emu_collective(EMU_RED_MIN_U, m, v, 0, 32); }
template<typename T> static inline unsigned __match_any_sync(unsigned m, T v) { return (unsigned)emu_collective(EMU_MATCH_ANY, m, emu_to_bits(v), 0, 32); }
static inline unsigned __activemask(void) { return 0xffffffffu; }

/* ---- scalar intrinsics ---- */
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((x >> i) & 1u) << (31 - i); return r; }
static inline unsigned long long __brevll(unsigned long long x) { unsigned long long r = 0; for (int i = 0; i < 64; ++i) r |= ((x >> i) & 1ull) << (63 - i); return r; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline float __int_as_float(int x) { float f; memcpy(&f, &x, 4); return f; }
static inline int __float_as_int(float f) { int x; memcpy(&x, &f, 4); return x; }
static inline float __uint_as_float(unsigned x) { float f; memcpy(&f, &x, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned x; memcpy(&x, &f, 4); return x; }
static inline double __longlong_as_double(long long x) { double d; memcpy(&d, &x, 8); return d; }
static inline long long __double_as_longlong(double d) { long long x; memcpy(&x, &d, 8); return x; }
/* the emulated library is compiled with -ffp-contract=off, so plain operators are unfused */
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __int2float_rn(int x) { return (float)x; }
static inline float __double2float_rn(double x) { return (float)x; }
static inline int __float2int_rz(float x) { return (int)x; }
template<typename T> static inline T __ldg(const T *p) { return *p; }
template<typename T> static inline T __ldcg(const T *p) { return *(const volatile T*)p; }
template<typename T> static inline T __ldcs(const T *p) { return *p; }
template<typename T> static inline void __stcg(T *p, T v) { *(volatile T*)p = v; }
template<typename T> static inline void __stcs(T *p, T v) { *p = v; }
static inline void __threadfence(void) { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block(void) { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

/* ---- atomics (real atomics: blocks may run on several OS threads) ---- */
template<typename T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template<typename T> static inline T atomicSub(T *p, T v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
template<typename T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template<typename T> static inline T atomicAnd(T *p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template<typename T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template<typename T> static inline T atomicCAS(T *p, T cmp, T v) { __atomic_compare_exchange_n(p, &cmp, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
template<typename T> static inline T atomicMax(T *p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o < v && !__atomic_compare_exchange_n(p, &o, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template<typename T> static inline T atomicMin(T *p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o > v && !__atomic_compare_exchange_n(p, &o, v, 0, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }

#endif
