// TEST INFRASTRUCTURE ONLY -- a tiny single-process SIMT emulator.
//
// Lets the product's .cu sources be compiled with plain g++ (-DDQA_EMUL) so that the
// *logic* of every kernel (indexing, barriers, warp shuffles, reductions) can be
// exercised in the CPU test-suite of a box without a GPU.  Each CUDA thread of a block
// is a ucontext fiber; blocks run one after another; __syncthreads / __syncwarp /
// __shfl_*_sync are cooperative barriers between fibers.  It is never loaded by the
// product package (perceptron_dqa_b200/_lib.py only loads the nvcc-built libdqa.so and
// raises if it is missing) -- it is not a CPU fallback, it is a debugger.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct int2 { int x, y; };

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __constant__ static

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };

namespace emul {

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  bool done = true;
};

struct State {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  int cur = -1;
  int nthreads = 0;
  // block barrier
  int bar_arrived = 0;
  unsigned bar_gen = 0;
  // warp barriers
  std::vector<int> w_arrived;
  std::vector<unsigned> w_gen;
  std::vector<double2> w_buf;  // 32 slots per warp
  char* smem = nullptr;
  size_t smem_cap = 0;
  std::function<void()> body;
  long progress = 0;
};

inline State& st() {
  static State s;
  return s;
}

inline uint3 threadIdx_;
inline uint3 blockIdx_;
inline dim3 blockDim_;
inline dim3 gridDim_;
inline char* dyn_smem_;

inline void set_tid(int t) {
  threadIdx_.x = t % blockDim_.x;
  threadIdx_.y = (t / blockDim_.x) % blockDim_.y;
  threadIdx_.z = t / (blockDim_.x * blockDim_.y);
}

inline void yield() {
  State& s = st();
  int me = s.cur;
  swapcontext(&s.fibers[me].ctx, &s.sched);
  set_tid(me);
}

inline void syncthreads() {
  State& s = st();
  unsigned g = s.bar_gen;
  s.progress++;
  if (++s.bar_arrived == s.nthreads) {
    s.bar_arrived = 0;
    s.bar_gen++;
  } else {
    while (s.bar_gen == g) yield();
  }
}

inline int warp_of(int t) { return t / 32; }
inline int lanes_in_warp(int w) {
  State& s = st();
  int lo = w * 32, hi = lo + 32;
  if (hi > s.nthreads) hi = s.nthreads;
  return hi - lo;
}

inline void syncwarp() {
  State& s = st();
  int w = warp_of(s.cur);
  unsigned g = s.w_gen[w];
  s.progress++;
  if (++s.w_arrived[w] == lanes_in_warp(w)) {
    s.w_arrived[w] = 0;
    s.w_gen[w]++;
  } else {
    while (s.w_gen[w] == g) yield();
  }
}

template <class T>
inline T shfl_generic(T v, int src_lane) {
  static_assert(sizeof(T) <= sizeof(double2), "shuffle payload too large");
  State& s = st();
  int t = s.cur, w = warp_of(t), lane = t % 32;
  std::memcpy(&s.w_buf[w * 32 + lane], &v, sizeof(T));
  syncwarp();
  T out;
  int sl = src_lane & 31;
  if (sl >= lanes_in_warp(w)) sl = lane;
  std::memcpy(&out, &s.w_buf[w * 32 + sl], sizeof(T));
  syncwarp();
  return out;
}

// mma.sync.m8n8k4.f64 with the PTX fragment layout (a=A[g][t], b=B[t][g], c=C[g][2t..2t+1])
inline void dmma884(double& c0, double& c1, double a, double b) {
  State& s = st();
  int t0 = s.cur, w = warp_of(t0), lane = t0 % 32;
  double2 mine;
  mine.x = a;
  mine.y = b;
  s.w_buf[w * 32 + lane] = mine;
  syncwarp();
  const int g = lane >> 2, t = lane & 3;
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < 4; ++k) {
    const double av = s.w_buf[w * 32 + g * 4 + k].x;             // A[g][k]
    s0 += av * s.w_buf[w * 32 + (2 * t) * 4 + k].y;              // B[k][2t]
    s1 += av * s.w_buf[w * 32 + (2 * t + 1) * 4 + k].y;          // B[k][2t+1]
  }
  syncwarp();
  c0 += s0;
  c1 += s1;
}

inline void fiber_entry() {
  State& s = st();
  int me = s.cur;
  set_tid(me);
  s.body();
  s.fibers[me].done = true;
}

inline void run_block() {
  State& s = st();
  const size_t STK = 256 * 1024;
  if ((int)s.fibers.size() < s.nthreads) s.fibers.resize(s.nthreads);
  for (int t = 0; t < s.nthreads; ++t) {
    Fiber& f = s.fibers[t];
    if (!f.stack) f.stack = (char*)std::malloc(STK);
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = STK;
    f.ctx.uc_link = &s.sched;
    f.done = false;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  s.bar_arrived = 0;
  int nw = (s.nthreads + 31) / 32;
  s.w_arrived.assign(nw, 0);
  s.w_gen.assign(nw, 0);
  s.w_buf.resize(nw * 32);
  int remaining = s.nthreads;
  // Scheduling: by default every pass runs all fibers in thread order (every warp advances one synchronisation point per
  // pass, in lock step).  With DQA_EMUL_SEED=<n> the order of the WARPS is shuffled every pass and, now and then, a warp
  // is held back for up to 255 passes while the others run until they block on it -- an adversarial schedule for
  // kernels whose warps hand data to each other without a block barrier (a warp racing a whole round ahead of its
  // neighbour is exactly what real hardware does and lock-step emulation hides).
  static const char* seed_env = std::getenv("DQA_EMUL_SEED");
  static unsigned long long rng = seed_env ? 0x9E3779B97F4A7C15ull * (unsigned long long)(std::atoll(seed_env) + 1) : 0ull;
  auto next_rand = [&]() {
    rng ^= rng << 13;
    rng ^= rng >> 7;
    rng ^= rng << 17;
    return rng;
  };
  std::vector<int> worder(nw), hold(nw, 0);
  for (int w = 0; w < nw; ++w) worder[w] = w;
  int idle_passes = 0;
  while (remaining > 0) {
    long before = s.progress;
    int ran = 0;
    if (seed_env)
      for (int w = nw - 1; w > 0; --w) std::swap(worder[w], worder[next_rand() % (unsigned)(w + 1)]);
    for (int wi = 0; wi < nw; ++wi) {
      const int w = worder[wi];
      if (seed_env) {
        if (hold[w] > 0) {
          hold[w]--;
          continue;
        }
        if ((next_rand() & 63) == 0) {
          hold[w] = (int)(next_rand() & 255);
          continue;
        }
      }
      for (int t = w * 32; t < s.nthreads && t < w * 32 + 32; ++t) {
        if (s.fibers[t].done) continue;
        s.cur = t;
        set_tid(t);
        swapcontext(&s.sched, &s.fibers[t].ctx);
        ran++;
        if (s.fibers[t].done) remaining--, s.progress++;
      }
    }
    if (ran > 0 && s.progress == before) {
      // a warp spinning on a flag makes no "progress" by itself: only call it a deadlock when nothing moves for a long time
      if (++idle_passes > (seed_env ? 4096 : 0)) {
        std::fprintf(stderr, "cuda_emul: deadlock (divergent barrier/shuffle) in block (%u,%u)\n", blockIdx_.x, blockIdx_.y);
        std::abort();
      }
    } else if (ran > 0) {
      idle_passes = 0;
    }
  }
}

inline void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
  State& s = st();
  s.body = std::move(body);
  s.nthreads = block.x * block.y * block.z;
  blockDim_ = block;
  gridDim_ = grid;
  if (smem + 64 > s.smem_cap) {
    std::free(s.smem);
    s.smem_cap = smem + 64;
    s.smem = (char*)std::aligned_alloc(128, ((s.smem_cap + 127) / 128) * 128);
  }
  dyn_smem_ = s.smem;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx_ = uint3{bx, by, bz};
        std::memset(s.smem, 0xCD, s.smem_cap);  // poison: catches reads of uninitialised smem
        run_block();
      }
}

}  // namespace emul

#define threadIdx (::emul::threadIdx_)
#define blockIdx (::emul::blockIdx_)
#define blockDim (::emul::blockDim_)
#define gridDim (::emul::gridDim_)

static inline void __syncthreads() { emul::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::syncwarp(); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lm) {
  return emul::shfl_generic(v, (emul::st().cur % 32) ^ lm);
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
  return emul::shfl_generic(v, src);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d) {
  int lane = emul::st().cur % 32;
  return emul::shfl_generic(v, lane + d < 32 ? lane + d : lane);
}
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p += v; return o; }
static inline int atomicOr(int* p, int v) { int o = *p; *p |= v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline long long clock64() { return 0; }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline unsigned __ballot_sync(unsigned, int pred) {
  // every lane contributes one bit
  unsigned mine = pred ? 1u : 0u;
  unsigned out = 0;
  for (int l = 0; l < 32; ++l) {
    unsigned b = emul::shfl_generic(mine, l);
    int t = emul::st().cur;
    if (l < emul::lanes_in_warp(emul::warp_of(t))) out |= (b & 1u) << l;
  }
  return out;
}

static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }

// ---- host runtime subset ----------------------------------------------------------
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
enum cudaDeviceAttr { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
// CUDA graphs: not emulated (the plan never enables them in this build); the calls only have to compile
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
enum cudaStreamCaptureMode { cudaStreamCaptureModeThreadLocal = 1 };
static const unsigned cudaStreamNonBlocking = 1;
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return 1; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return 1; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return 1; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return 0; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return 0; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return 1; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 227 * 1024; return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emul"; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static const unsigned cudaEventDisableTiming = 2;
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
