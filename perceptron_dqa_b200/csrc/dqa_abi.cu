// dqa_abi.cu -- the C ABI of libdqa.so (include/dqa.h): plan, workspace carving, launch order.
// Host code only orchestrates; every arithmetic step of the hot path is a kernel in
// dqa_kernels.cuh.  There is no CPU fallback: with no CUDA device every call that touches
// the device returns the cudaError_t it got.
#include "../../include/dqa.h"
#include "dqa_kernels.cuh"

#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#ifdef DQA_EMUL
#define OPS_THREADS 64
#else
#define OPS_THREADS 1024
#endif

// template instances with a comma in their argument list, named so they can pass through macros
static void (*const KJ_0S)(DqaDev, int, int, int) = k_jacobi<true>;
static void (*const KJ_0G)(DqaDev, int, int, int) = k_jacobi<false>;
static void (*const KE_1)(DqaDev) = k_energy<false>;
static void (*const KE_2)(DqaDev) = k_energy<true>;
static void (*const KJ_Y2)(DqaDev, int, int, int) = k_jacobi_sys<2>;
static void (*const KJ_Y3)(DqaDev, int, int, int) = k_jacobi_sys<3>;
static void (*const KJ_Y4)(DqaDev, int, int, int) = k_jacobi_sys<4>;
static void (*const KJ_Y6)(DqaDev, int, int, int) = k_jacobi_sys<6>;
static void (*const KJ_Y8)(DqaDev, int, int, int) = k_jacobi_sys<8>;
static void (*const KJ_Y10)(DqaDev, int, int, int) = k_jacobi_sys<10>;  // chi <= 40 .. 64: one CTA per SM, the factor still in registers
static void (*const KJ_Y12)(DqaDev, int, int, int) = k_jacobi_sys<12>;
#ifndef DQA_EMUL
static void (*const KJ_C16)(DqaDev, int, int, int) = k_jacobi_sys_cl<16, 2, 1, 8>;  // 56 < chi <= 64: the systolic array on a cluster of two CTAs
#endif

struct dqa_handle {
  dqa_config cfg;
  DqaDev dev;
  cudaStream_t stream;
  void* ws;
  size_t ws_bytes;
  int8_t* d_xi;   // [batch][n_xi][N]
  int32_t* d_e;   // scratch for dqa_get_index_tensors
  int32_t* d_q;
  int pp;
  bool have_patterns, have_fourier, have_state;
  int t_qr, t_theta, t_svd, t_energy, t_energy_env = 0;  // block sizes
  int t_lq;
  cudaStream_t cap_stream = nullptr;   // capture stream of the step graph (the handle's own stream may be the legacy one)
  cudaGraphExec_t step_graph = nullptr;  // one dQA step with the step counter read from the device
  bool use_graph = false, capturing = false;
  // the captured step runs the batch as n_split independent sub-batches on parallel branches of the graph (DQA_SPLIT): every
  // launch of the step ends in a tail where the last CTAs of the batch finish on otherwise idle SMs; the branches fill
  // each other's tails (measured at batch 888: +3.3 % realisation-steps/s with two branches, +4.0 % with three, +3.7 % with four
  // at chi = 32; +4.6 % with three at chi = 10 and 20; +3 % with two at batch 148 and 296; +-0 at batch 128)
  int n_split = 1;
  cudaStream_t split_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  long long graph_nodes = 0;
  bool jac_sys;
  size_t jac_pad = 0;  // DQA_JAC_SMEM_PAD: extra dynamic shared memory of the systolic Jacobi (occupancy experiments: >= 50 KB leaves one CTA per SM at chi = 32)
  int jac_sys_big = 2;  // systolic Jacobi above chi = 32 (DQA_JAC_SYS_BIG): 0 off; 1: chi <= 48 (one CTA per SM, 166 registers); 2: also 56 < chi <= 64 on a cluster of two CTAs
  char err[512];
  // profiling
  bool profiling;
  struct Ev { int fam; cudaEvent_t a, b; };
  std::vector<Ev> evs;
  size_t ev_used;
  long long launches;
};

// A handle lives on cfg.device: every entry point makes that device current for its own duration and
// restores the caller's (and torch's) current device on exit.
struct DevGuard {
  int prev = -1;
  bool switched = false;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DevGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

static int fail(const dqa_t* h, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(const_cast<dqa_t*>(h)->err, sizeof(h->err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return fail(h, (int)e_, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

#define HL(...) do { h->launches++; DQA_LAUNCH(__VA_ARGS__); } while (0)

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// one table drives both the size query and the carving
struct Carve {
  char* base;
  size_t off;
  template <class T>
  T* take(size_t n) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off = align_up(off + n * sizeof(T));
    return p;
  }
};

struct Aux {
  int8_t* d_xi;
  int32_t* d_e;
  int32_t* d_q;
};

static void carve(const dqa_config& c, DqaDev& d, Aux& aux, char* base, size_t* total) {
  const size_t B = c.batch, N = c.N, D = c.N + 1, chi = c.chi, P = c.P, nx = c.n_xi;
  Carve cv{base, 0};
  d.mps = cv.take<cplx>(B * N * 2 * chi * chi);
  d.bond = cv.take<int>(B * (N + 1));
  d.hbit = cv.take<uint8_t>(B * nx * N);
  d.uz = cv.take<cplx>((size_t)d.ntab * P * D);
  d.tab = cv.take<int>(B);
  d.dtv = cv.take<double>(B);
  d.pdev = cv.take<int>(1);
  d.mixtab = cv.take<double>(2 * P);
  d.gk = cv.take<cplx>(D);
  d.phase = cv.take<cplx>(D);
  d.rbuf = cv.take<cplx>(B * 2 * d.smax * chi * chi);
  d.qbuf = cv.take<cplx>(B * (size_t)d.nsec_total * 2 * chi * chi);
  d.qdim = cv.take<int>(B * (N + 1) * d.smax);
  d.theta = cv.take<cplx>(B * 2 * 2 * chi * (size_t)d.ld_theta);
  d.work = cv.take<cplx>(B * 2 * chi * (size_t)d.ld_theta);
  d.lbuf = cv.take<cplx>(B * 4 * chi * chi);
  d.scale = cv.take<double>(B);
  d.schmidt = cv.take<double>(B * N * 2 * chi);
  d.nsv = cv.take<int>(B * N);
  d.tmk = cv.take<cplx>(B * nx * d.kh);
  d.eps = cv.take<cplx>(B * (P + 1));
  d.escratch = cv.take<cplx>(B);
  d.discw = cv.take<double>(B * N);
  d.bdmon = cv.take<int>(B * P * nx);
  d.status = cv.take<int>(B);
  d.counters = cv.take<unsigned long long>(B * DQA_NCNT);
  aux.d_xi = cv.take<int8_t>(B * nx * N);
  aux.d_e = cv.take<int32_t>(nx * N * 2);
  aux.d_q = cv.take<int32_t>(nx * N * 2 * D);
  *total = cv.off;
}

// The plan restricted to instances [i0, i0 + nb): every per-instance buffer advanced by i0 strides (same table as carve()).
static DqaDev view_of(const DqaDev& f, const dqa_config& c, size_t i0, int nb) {
  DqaDev d = f;
  const size_t N = c.N, chi = c.chi, P = c.P, nx = c.n_xi;
  d.batch = nb;
  d.mps += i0 * N * 2 * chi * chi;
  d.bond += i0 * (N + 1);
  d.hbit += i0 * nx * N;
  d.tab += i0;
  d.dtv += i0;
  d.rbuf += i0 * 2 * f.smax * chi * chi;
  d.qbuf += i0 * (size_t)f.nsec_total * 2 * chi * chi;
  d.qdim += i0 * (N + 1) * f.smax;
  d.theta += i0 * 2 * 2 * chi * (size_t)f.ld_theta;
  d.work += i0 * 2 * chi * (size_t)f.ld_theta;
  d.lbuf += i0 * 4 * chi * chi;
  d.scale += i0;
  d.schmidt += i0 * N * 2 * chi;
  d.nsv += i0 * N;
  d.tmk += i0 * nx * f.kh;
  d.eps += i0 * (P + 1);
  d.escratch += i0;
  d.discw += i0 * N;
  d.bdmon += i0 * P * nx;
  d.status += i0;
  d.counters += i0 * DQA_NCNT;
  return d;
}

extern "C" const char* dqa_version(void) { return "dqa-b200 0.1 (sector path, complex128)"; }

extern "C" int dqa_create(dqa_t** out, const dqa_config* cfg) {
  if (!out || !cfg) return DQA_EINVAL;
  *out = nullptr;
  dqa_t* h = new dqa_t();
  h->err[0] = 0;
  h->cfg = *cfg;
  dqa_config& c = h->cfg;
  if (c.max_sweeps <= 0) c.max_sweeps = 30;
  if (c.cutoff == 0.0) c.cutoff = 1e-10;
  if (c.cutoff < 0.0) c.cutoff = 0.0;
  if (!c.policy_set) {  // the reference's hard-coded call site, lib/tenn.py:238
    c.cutoff_mode = 4;
    c.renorm = 2;
    c.absorb = 1;
  }
  if (c.n_tables <= 0) c.n_tables = 1;
  *out = h;  // returned even on error so the message can be read; caller destroys it
  if (c.N < 2 || c.n_xi < 1 || c.P < 1 || c.chi < 1 || c.batch < 1) return fail(h, DQA_EINVAL, "N>=2, n_xi>=1, P>=1, chi>=1, batch>=1 required");
  if (c.N > 254) return fail(h, DQA_ELIMIT, "N=%d > 254 not supported", c.N);
  if (sector_qr_smem(c.chi) > 227 * 1024 || svd_smem(c.chi) > 227 * 1024 || energy_smem(c.chi) > 227 * 1024 ||
      false)
    return fail(h, DQA_ELIMIT, "chi=%d at N=%d needs more than 227 KB of shared memory per CTA in this build (chi <= 64, and chi*(N+1) bounded by the LQ panel)", c.chi, c.N);
  if (c.chi > 64) return fail(h, DQA_ELIMIT, "chi=%d > 64 not supported", c.chi);
  if (c.cutoff_mode < 1 || c.cutoff_mode > 6) return fail(h, DQA_EINVAL, "cutoff_mode must be 1..6 (lib/tenn.py:317-345)");
  if (c.renorm < 0) return fail(h, DQA_EINVAL, "renorm must be >= 0");
  if (c.renorm > 0 && c.cutoff_mode < 3) return fail(h, DQA_EINVAL, "renorm needs a cumulative cutoff mode (3..6), as in lib/tenn.py:347-354");
  if (c.absorb != 1)
    return fail(h, DQA_ELIMIT, "absorb=%d: the left-to-right sweep of compress_svd_normalized absorbs to the right (lib/tenn.py:238); "
                "other values are served by the op-level dqa_op_svd_truncated", c.absorb);
  if (c.n_tables > c.batch) return fail(h, DQA_EINVAL, "n_tables > batch");
  h->stream = (cudaStream_t)c.stream;
  h->ws = nullptr;
  h->ws_bytes = 0;
  h->pp = 0;
  h->have_patterns = h->have_fourier = h->have_state = false;
  h->profiling = false;
  h->ev_used = 0;
  h->launches = 0;
  DqaDev& d = h->dev;
  memset(&d, 0, sizeof(d));
  d.N = c.N;
  d.D = c.N + 1;
  d.n_xi = c.n_xi;
  d.chi = c.chi;
  d.batch = c.batch;
  d.P = c.P;
  d.smax = c.N + 1;
  d.ld_theta = ((c.chi * (c.N + 1)) + 1) & ~1;
  d.nsec_total = dqa_secbase(c.N, c.N);
  {
    // bond n (1..N-1) of the input MPS is at most min(chi, 2^n, 2^(N-n)); it carries N-n+1 sectors
    long qm = 1;
    for (int n = 1; n < c.N; ++n) {
      long b = c.chi;
      if (n < 30 && (1L << n) < b) b = 1L << n;
      if (c.N - n < 30 && (1L << (c.N - n)) < b) b = 1L << (c.N - n);
      const long v = b * (c.N - n + 1);
      if (v > qm) qm = v;
    }
    d.qmax = (int)qm;
    d.lq_pb = lq_pick_pb(d.qmax, c.chi);
    if (lq_smem(d.qmax, c.chi, d.lq_pb) > 227 * 1024)
      return fail(h, DQA_ELIMIT, "chi=%d, N=%d: one LQ panel row (%d columns) does not fit in shared memory", c.chi, c.N, d.qmax);
  }
  d.kh = d.D / 2 + 1;
  d.max_sweeps = c.max_sweeps;
  d.cutoff = c.cutoff;
  d.cutoff_mode = c.cutoff_mode;
  d.renorm = c.renorm;
  d.ntab = c.n_tables;
  {
    const char* env = getenv("DQA_JAC_QUAD");  // relative threshold, default 1e-10; 0 disables the shortcut
    const double q = env ? atof(env) : 1e-10;
    d.jac_quad2 = q * q;
  }
  // block sizes: warps scale with chi; the emulator build keeps them small
#ifdef DQA_EMUL
  h->t_qr = 32 * ((8 * c.chi + 31) / 32); h->t_theta = 64; h->t_svd = 64; h->t_energy = 64; h->t_lq = 64;
  if (h->t_qr < 64) h->t_qr = 64;
#else
  h->t_qr = 32 * ((8 * c.chi + 31) / 32);  // one 8-lane group per column of the sector block
  if (h->t_qr < 64) h->t_qr = 64;
  // measured (tools/block_size_sweep.py, batch 888): the tile counts of the two contraction kernels grow as (chi/8)^2, and
  // CTAs no larger than their task count win by a wide margin at small bonds (chi=10: theta 34 -> 14 ms per step)
  h->t_theta = c.chi <= 12 ? 32 : (c.chi <= 24 ? 64 : (c.chi <= 32 ? 128 : 256));
  {
    const char* env = getenv("DQA_THETA_THREADS");
    if (env && atoi(env) >= 32 && atoi(env) <= 256) h->t_theta = atoi(env) & ~31;
    env = getenv("DQA_ENERGY_THREADS");
    if (env && atoi(env) >= 32 && atoi(env) <= 256) h->t_energy_env = atoi(env) & ~31;
  }
  {
    // Jacobi: few warps per CTA, each owning ~8 column pairs per round; several CTAs share an SM
    int warps = (c.chi + 3) / 4;  // 4 column pairs per warp and round: one pass per round
    if (warps < 2) warps = 2;
    if (warps > 16) warps = 16;
    const char* env = getenv("DQA_JAC_WARPS");
    if (env && atoi(env) >= 1 && atoi(env) <= 16) warps = atoi(env);
    if (warps < (c.chi + 31) / 32) warps = (c.chi + 31) / 32;  // at most 32 column pairs per warp
    h->t_svd = 32 * warps;
    h->t_lq = c.chi <= 8 ? 128 : (c.chi <= 24 ? 256 : 512);
    env = getenv("DQA_LQ_THREADS");
    if (env && atoi(env) >= 32 && atoi(env) <= 32 * LQ_MAXW) h->t_lq = atoi(env) & ~31;
  }
  h->t_energy = c.chi <= 16 ? 64 : (c.chi <= 24 ? 128 : 256);
  if (h->t_energy_env) h->t_energy = h->t_energy_env;
#endif
  {
    const char* env = getenv("DQA_GRAPH");  // 0: plain launches for dqa_step (default: one captured graph per step)
    h->use_graph = !(env && atoi(env) == 0);
#ifdef DQA_EMUL
    h->use_graph = false;
#endif
    env = getenv("DQA_SPLIT");  // sub-batches on parallel graph branches (measured, N=21: never slower; nothing to gain below a wave)
    h->n_split = env ? atoi(env) : (c.batch >= 444 ? 3 : (c.batch >= 148 ? 2 : 1));
    if (h->n_split < 1) h->n_split = 1;
    if (h->n_split > 4) h->n_split = 4;
    if (h->n_split > c.batch) h->n_split = c.batch;
  }
  {
    const char* env = getenv("DQA_JAC_SYS");  // 0: use the shared-memory-resident Jacobi for every chi
    h->jac_sys = !(env && atoi(env) == 0);
    env = getenv("DQA_JAC_SYS_BIG");
    if (env) h->jac_sys_big = atoi(env);
    if (!h->jac_sys) h->jac_sys_big = 0;
    env = getenv("DQA_JAC_SMEM_PAD");
    if (env && atol(env) > 0) h->jac_pad = (size_t)atol(env);
  }
  DevGuard guard(c.device);
  {
    int cur = -1;
    CK(cudaGetDevice(&cur));
    if (cur != c.device) CK(cudaSetDevice(c.device));  // surfaces an invalid ordinal; the guard restores the caller's device
  }
  // The opt-in limit is a property of the (process-global) kernel, not of this handle: always raise it to the device
  // maximum, so a later plan with a smaller chi can never lower it under a live plan's launches.
  int smem_optin = 0;
  CK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device));
#define DQA_OPTIN(f) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin))
  DQA_OPTIN(k_sector_qr<2>); DQA_OPTIN(k_sector_qr<3>); DQA_OPTIN(k_sector_qr<4>); DQA_OPTIN(k_sector_qr<6>);
  DQA_OPTIN(k_sector_qr<8>); DQA_OPTIN(k_sector_qr<10>); DQA_OPTIN(k_sector_qr<12>); DQA_OPTIN(k_sector_qr<16>);
  DQA_OPTIN(KJ_0S); DQA_OPTIN(KJ_0G);
  DQA_OPTIN(KJ_Y2); DQA_OPTIN(KJ_Y3); DQA_OPTIN(KJ_Y4); DQA_OPTIN(KJ_Y6); DQA_OPTIN(KJ_Y8);
  DQA_OPTIN(KJ_Y10); DQA_OPTIN(KJ_Y12);
#ifndef DQA_EMUL
  DQA_OPTIN(KJ_C16);
#endif
  DQA_OPTIN(k_lq); DQA_OPTIN(KE_1); DQA_OPTIN(KE_2); DQA_OPTIN(k_theta);
#undef DQA_OPTIN
  return DQA_OK;
}

extern "C" void dqa_destroy(dqa_t* h) {
  if (!h) return;
  DevGuard guard_(h->cfg.device);
  for (auto& e : h->evs) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
#ifndef DQA_EMUL
  if (h->step_graph) cudaGraphExecDestroy(h->step_graph);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
#endif
  for (int i = 0; i < 3; ++i) {
    if (h->split_stream[i]) cudaStreamDestroy(h->split_stream[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  delete h;
}

extern "C" const char* dqa_last_error(const dqa_t* h) { return h ? h->err : "null handle"; }

extern "C" int dqa_workspace_bytes(const dqa_t* h, size_t* bytes) {
  if (!h || !bytes) return DQA_EINVAL;
  DqaDev tmp = h->dev;  // carve() only writes pointers; work on a copy
  Aux aux;
  carve(h->cfg, tmp, aux, nullptr, bytes);
  return DQA_OK;
}

extern "C" int dqa_bind_workspace(dqa_t* h, void* p, size_t bytes) {
  if (!h || !p) return DQA_EINVAL;
  DevGuard guard_(h->cfg.device);
  size_t need = 0;
  DqaDev tmp = h->dev;
  Aux aux;
  carve(h->cfg, tmp, aux, nullptr, &need);
  if (bytes < need) return fail(h, DQA_EINVAL, "workspace too small: %zu < %zu", bytes, need);
  if ((uintptr_t)p & 255) return fail(h, DQA_EINVAL, "workspace must be 256-byte aligned");
  carve(h->cfg, h->dev, aux, (char*)p, &need);
  h->d_xi = aux.d_xi;
  h->d_e = aux.d_e;
  h->d_q = aux.d_q;
  h->ws = p;
  h->ws_bytes = bytes;
  CK(cudaMemsetAsync(p, 0, need, h->stream));
  h->have_patterns = h->have_fourier = h->have_state = false;
  return DQA_OK;
}

#define NEED_WS() \
  if (!h) return DQA_EINVAL; \
  DevGuard guard_(h->cfg.device); \
  if (!h->ws) return fail(h, DQA_ESTATE, "bind a workspace first")

extern "C" int dqa_set_patterns(dqa_t* h, const int8_t* xi) {
  NEED_WS();
  if (!xi) return DQA_EINVAL;
  const size_t n = (size_t)h->cfg.batch * h->cfg.n_xi * h->cfg.N;
  for (size_t i = 0; i < n; ++i)
    if (xi[i] != 1 && xi[i] != -1) return fail(h, DQA_EINVAL, "pattern entry %zu is %d, not +-1", i, (int)xi[i]);
  CK(cudaMemcpyAsync(h->d_xi, xi, n, cudaMemcpyHostToDevice, h->stream));
  const int blocks = (int)((n + 255) / 256);
  HL(k_hamming_tables, dim3(blocks < 1024 ? blocks : 1024), dim3(256), 0, h->stream, (const int8_t*)h->d_xi, (int)n,
             h->dev.D, const_cast<uint8_t*>(h->dev.hbit), (int32_t*)nullptr, (int32_t*)nullptr);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  h->have_patterns = true;
  return DQA_OK;
}

extern "C" int dqa_get_index_tensors(dqa_t* h, int inst, uint8_t* hbit, int32_t* e, int32_t* q) {
  NEED_WS();
  if (!h->have_patterns) return fail(h, DQA_ESTATE, "set patterns first");
  if (inst < 0 || inst >= h->cfg.batch) return fail(h, DQA_EINVAL, "instance out of range");
  const size_t n = (size_t)h->cfg.n_xi * h->cfg.N;
  const int D = h->dev.D;
  const int8_t* xi = h->d_xi + (size_t)inst * n;
  const int blocks = (int)((n + 127) / 128);
  // hbit is recomputed into the tail of d_q's scratch? no: reuse the plan's table for this instance
  HL(k_hamming_tables, dim3(blocks), dim3(128), 0, h->stream, xi, (int)n, D, (uint8_t*)nullptr, h->d_e, h->d_q);
  CK(cudaGetLastError());
  if (hbit) CK(cudaMemcpyAsync(hbit, h->dev.hbit + (size_t)inst * n, n, cudaMemcpyDeviceToHost, h->stream));
  if (e) CK(cudaMemcpyAsync(e, h->d_e, n * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  if (q) CK(cudaMemcpyAsync(q, h->d_q, n * 2 * D * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

// u_p(x) = (1/sqrt D) sum_k Uk_FT[k,p] exp(2 pi i k x / D): the diagonal of the Fourier MPO U_z
// (lib/dQA_utils.py:46-59, product over the N sites of the per-site coefficient and phases).
static void uz_from_ukft(const std::complex<double>* uk, int D, int P, std::complex<double>* uz) {
  typedef std::complex<double> cd;
  std::vector<cd> w(D);
  for (int k = 0; k < D; ++k) {
    const double ang = 2.0 * M_PI * k / D;
    w[k] = cd(std::cos(ang), std::sin(ang));
  }
  const double rs = 1.0 / std::sqrt((double)D);
  for (int p = 0; p < P; ++p)
    for (int x = 0; x < D; ++x) {
      cd acc(0.0, 0.0);
      for (int k = 0; k < D; ++k) acc += uk[(size_t)k * P + p] * w[(int)(((long)k * x) % D)];
      uz[(size_t)p * D + x] = acc * rs;
    }
}

static int set_fourier_impl(dqa_t* h, int ntab, const double* uk_ft, const double* fxft, const int32_t* tab, const double* dts) {
  typedef std::complex<double> cd;
  const int D = h->dev.D, P = h->cfg.P, B = h->cfg.batch;
  const cd* uk = reinterpret_cast<const cd*>(uk_ft);
  const cd* fx = reinterpret_cast<const cd*>(fxft);
  std::vector<cd> uz((size_t)ntab * P * D), gk(D), ph(D);
  for (int k = 0; k < D; ++k) {
    // phase table of the energy sweep: exp(i*(pi/D)*k*e) with e = 2 (lib/dQA_utils.py:106)
    const double a2 = (M_PI / D) * k * 2.0;
    ph[k] = cd(std::cos(a2), std::sin(a2));
    gk[k] = fx[k] / std::sqrt((double)D);
  }
  for (int t = 0; t < ntab; ++t) uz_from_ukft(uk + (size_t)t * D * P, D, P, uz.data() + (size_t)t * P * D);
  std::vector<double> mix((size_t)2 * P);
  for (int p = 0; p < P; ++p) {
    const double beta = (1.0 - (double)(p + 1) / P) * h->cfg.dt;  // dQA_mps.py:88-89
    mix[2 * p] = std::cos(beta);
    mix[2 * p + 1] = std::sin(beta);
  }
  CK(cudaMemcpyAsync(const_cast<double*>(h->dev.mixtab), mix.data(), sizeof(double) * mix.size(), cudaMemcpyHostToDevice, h->stream));
  std::vector<int> tb(B, 0);
  std::vector<double> dv(B, h->cfg.dt);
  for (int b = 0; b < B; ++b) {
    if (tab) {
      if (tab[b] < 0 || tab[b] >= ntab) return fail(h, DQA_EINVAL, "instance %d: table %d out of range", b, tab[b]);
      tb[b] = tab[b];
    }
    if (dts) dv[b] = dts[b];
  }
  CK(cudaMemcpyAsync(const_cast<cplx*>(h->dev.uz), uz.data(), sizeof(cd) * uz.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(const_cast<cplx*>(h->dev.gk), gk.data(), sizeof(cd) * D, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(const_cast<cplx*>(h->dev.phase), ph.data(), sizeof(cd) * D, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(const_cast<int*>(h->dev.tab), tb.data(), sizeof(int) * B, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(const_cast<double*>(h->dev.dtv), dv.data(), sizeof(double) * B, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->have_fourier = true;
  return DQA_OK;
}

extern "C" int dqa_set_fourier(dqa_t* h, const double* uk_ft, const double* fxft) {
  NEED_WS();
  if (!uk_ft || !fxft) return DQA_EINVAL;
  if (h->dev.ntab != 1) return fail(h, DQA_ESTATE, "plan was created with n_tables=%d: use dqa_set_fourier_tables", h->dev.ntab);
  return set_fourier_impl(h, 1, uk_ft, fxft, nullptr, nullptr);
}

extern "C" int dqa_set_fourier_tables(dqa_t* h, int n_tables, const double* uk_ft, const double* fxft, const int32_t* table_of_instance,
                                      const double* dt_of_instance) {
  NEED_WS();
  if (!uk_ft || !fxft || !table_of_instance || !dt_of_instance) return DQA_EINVAL;
  if (n_tables != h->dev.ntab) return fail(h, DQA_EINVAL, "n_tables=%d but the plan was created for %d", n_tables, h->dev.ntab);
  return set_fourier_impl(h, n_tables, uk_ft, fxft, table_of_instance, dt_of_instance);
}

// ---- profiling helpers ---------------------------------------------------------------
static void prof_a(dqa_t* h, int fam) {
  if (!h->profiling) return;
  if (h->ev_used == h->evs.size()) {
    dqa_handle::Ev e;
    e.fam = fam;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    h->evs.push_back(e);
  }
  h->evs[h->ev_used].fam = fam;
  cudaEventRecord(h->evs[h->ev_used].a, h->stream);
}
static void prof_b(dqa_t* h) {
  if (!h->profiling) return;
  cudaEventRecord(h->evs[h->ev_used].b, h->stream);
  h->ev_used++;
}

extern "C" int dqa_profile_begin(dqa_t* h) {
  if (!h) return DQA_EINVAL;
  h->profiling = true;
  h->ev_used = 0;
  return DQA_OK;
}

extern "C" int dqa_profile_end(dqa_t* h, double* ms, int64_t* launches) {
  if (!h || !ms || !launches) return DQA_EINVAL;
  DevGuard guard_(h->cfg.device);
  CK(cudaStreamSynchronize(h->stream));
  for (int f = 0; f < DQA_NFAM; ++f) ms[f] = 0.0, launches[f] = 0;
  for (size_t i = 0; i < h->ev_used; ++i) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, h->evs[i].a, h->evs[i].b));
    ms[h->evs[i].fam] += t;
    launches[h->evs[i].fam]++;
  }
  h->profiling = false;
  h->ev_used = 0;
  return DQA_OK;
}

// ---- launch sequences ------------------------------------------------------------------
static int launch_canonize(dqa_t* h, int mu) {
  const DqaDev& d = h->dev;
  HL(k_sector_init, dim3(d.batch), dim3(32), 0, h->stream, d);
  for (int n = d.N - 1; n >= 1; --n) {
    prof_a(h, 0);
    const dim3 grid(d.N - n + 1, d.batch), block(h->t_qr);
    const size_t sm = sector_qr_smem(d.chi);
    if (d.chi <= 8) HL(k_sector_qr<2>, grid, block, sm, h->stream, d, n, mu);
    else if (d.chi <= 12) HL(k_sector_qr<3>, grid, block, sm, h->stream, d, n, mu);
    else if (d.chi <= 16) HL(k_sector_qr<4>, grid, block, sm, h->stream, d, n, mu);
    else if (d.chi <= 24) HL(k_sector_qr<6>, grid, block, sm, h->stream, d, n, mu);
    else if (d.chi <= 32) HL(k_sector_qr<8>, grid, block, sm, h->stream, d, n, mu);
    else if (d.chi <= 40) HL(k_sector_qr<10>, grid, block, sm, h->stream, d, n, mu);
    else if (d.chi <= 48) HL(k_sector_qr<12>, grid, block, sm, h->stream, d, n, mu);
    else HL(k_sector_qr<16>, grid, block, sm, h->stream, d, n, mu);
    prof_b(h);
  }
  CK(cudaGetLastError());
  return DQA_OK;
}

static int launch_truncate(dqa_t* h, int p, int mu) {
  const DqaDev& d = h->dev;
  prof_a(h, 1);
  HL(k_theta0, dim3(d.N, d.batch), dim3(64), 0, h->stream, d, mu, p);
  prof_b(h);
  for (int l = 0; l < d.N; ++l) {
    if (l > 0) {
      prof_a(h, 1);
      HL(k_theta, dim3(d.N - l + 1, d.batch), dim3(h->t_theta), theta_smem(d.chi), h->stream, d, l, mu);
      prof_b(h);
    }
    prof_a(h, 2);
    HL(k_lq, dim3(d.batch), dim3(h->t_lq), lq_smem(d.qmax, d.chi, d.lq_pb), h->stream, d, l);
    prof_b(h);
    prof_a(h, 5);
    {
      // chi <= 32: the systolic kernel (2chi <= 8 RS rows per column); above: L in shared memory (chi <= 56) or in L2
      const int t_sys = 32 * ((d.chi + 3) / 4);  // one 8-lane group per column pair
      if (h->jac_sys && d.chi <= 8) HL(KJ_Y2, dim3(d.batch), dim3(t_sys), svd_smem(d.chi), h->stream, d, l, mu, p);
      else if (h->jac_sys && d.chi <= 12) HL(KJ_Y3, dim3(d.batch), dim3(t_sys), svd_smem(d.chi), h->stream, d, l, mu, p);
      else if (h->jac_sys && d.chi <= 16) HL(KJ_Y4, dim3(d.batch), dim3(t_sys), svd_smem(d.chi), h->stream, d, l, mu, p);
      else if (h->jac_sys && d.chi <= 24) HL(KJ_Y6, dim3(d.batch), dim3(t_sys), svd_smem(d.chi), h->stream, d, l, mu, p);
      else if (h->jac_sys && d.chi <= 32) HL(KJ_Y8, dim3(d.batch), dim3(t_sys), svd_smem(d.chi) + h->jac_pad, h->stream, d, l, mu, p);
      else if (h->jac_sys_big && d.chi <= 40) HL(KJ_Y10, dim3(d.batch), dim3(t_sys), svd_smem(d.chi), h->stream, d, l, mu, p);
      else if (h->jac_sys_big && d.chi <= 48) HL(KJ_Y12, dim3(d.batch), dim3(t_sys), svd_smem(d.chi), h->stream, d, l, mu, p);
#ifndef DQA_EMUL
      // 48 < chi <= 56 stays on the shared-memory kernel: L fits there, and the cluster build was measured 2 % slower at chi = 56
      else if (h->jac_sys_big > 1 && d.chi > 56 && d.chi <= 64) HL(KJ_C16, dim3(2 * d.batch), dim3(32 * ((8 * ((d.chi + 1) / 2) + 31) / 32)), svd_cl_smem(d.chi, 2), h->stream, d, l, mu, p);
#endif
      else if (svd_l_in_smem(d.chi)) HL(KJ_0S, dim3(d.batch), dim3(h->t_svd), svd_smem(d.chi), h->stream, d, l, mu, p);
      else HL(KJ_0G, dim3(d.batch), dim3(h->t_svd), svd_smem(d.chi), h->stream, d, l, mu, p);
    }
    prof_b(h);
  }
  CK(cudaGetLastError());
  return DQA_OK;
}

// slot >= 0: eps trace slot; -1: the scratch vector read by dqa_energy; -2: trace slot p+1 with p read on the device
static int launch_energy(dqa_t* h, int slot) {
  const DqaDev& d = h->dev;
  prof_a(h, 4);
  if (d.chi > 16) HL(KE_2, dim3(d.kh, d.n_xi, d.batch), dim3(h->t_energy), energy_smem(d.chi), h->stream, d);
  else HL(KE_1, dim3(d.kh, d.n_xi, d.batch), dim3(h->t_energy), energy_smem(d.chi), h->stream, d);
  if (slot >= 0) HL(k_energy_reduce, dim3(d.batch), dim3(32), 0, h->stream, d, d.eps + slot, (size_t)(d.P + 1));
  else if (slot == -1) HL(k_energy_reduce, dim3(d.batch), dim3(32), 0, h->stream, d, d.escratch, (size_t)1);
  else HL(k_energy_reduce, dim3(d.batch), dim3(32), 0, h->stream, d, (cplx*)nullptr, (size_t)(d.P + 1));  // slot from the device step counter
  prof_b(h);
  CK(cudaGetLastError());
  return DQA_OK;
}

#define NEED_READY() \
  NEED_WS(); \
  if (!h->have_patterns || !h->have_fourier) return fail(h, DQA_ESTATE, "set patterns and Fourier tables first"); \
  if (!h->have_state) return fail(h, DQA_ESTATE, "no state: call dqa_reset_plus or dqa_set_mps first")

extern "C" int dqa_reset_plus(dqa_t* h) {
  NEED_WS();
  if (!h->have_patterns || !h->have_fourier) return fail(h, DQA_ESTATE, "set patterns and Fourier tables first");
  const DqaDev& d = h->dev;
  HL(k_init_plus, dim3(d.batch), dim3(256), 0, h->stream, d);
  CK(cudaMemsetAsync(d.counters, 0, sizeof(unsigned long long) * DQA_NCNT * d.batch, h->stream));
  h->pp = 0;
  h->have_state = true;
  return launch_energy(h, 0);
}

extern "C" int dqa_get_step(const dqa_t* h, int* pp) {
  if (!h || !pp) return DQA_EINVAL;
  *pp = h->pp;
  return DQA_OK;
}
extern "C" int dqa_set_step(dqa_t* h, int pp) {
  if (!h) return DQA_EINVAL;
  if (pp < 0 || pp > h->cfg.P) return fail(h, DQA_EINVAL, "step out of range");
  h->pp = pp;
  return DQA_OK;
}

extern "C" int dqa_canonize(dqa_t* h, int p, int mu) {
  NEED_READY();
  if (p < 0 || p >= h->cfg.P || mu < 0 || mu >= h->cfg.n_xi) return fail(h, DQA_EINVAL, "p or mu out of range");
  return launch_canonize(h, mu);
}
extern "C" int dqa_truncate(dqa_t* h, int p, int mu) {
  NEED_READY();
  if (p < 0 || p >= h->cfg.P || mu < 0 || mu >= h->cfg.n_xi) return fail(h, DQA_EINVAL, "p or mu out of range");
  return launch_truncate(h, p, mu);
}
extern "C" int dqa_apply_pattern(dqa_t* h, int p, int mu) {
  int rc = dqa_canonize(h, p, mu);
  if (rc) return rc;
  return launch_truncate(h, p, mu);
}

static int launch_mixer(dqa_t* h, double beta, double one_minus_s, int mode) {
  const DqaDev& d = h->dev;
  prof_a(h, 3);
  HL(k_mixer, dim3(d.N, d.batch), dim3(128), 0, h->stream, d, std::cos(beta), std::sin(beta), one_minus_s, mode);
  prof_b(h);
  CK(cudaGetLastError());
  return DQA_OK;
}

extern "C" int dqa_apply_ux(dqa_t* h, double beta) {
  NEED_READY();
  return launch_mixer(h, beta, 0.0, 0);
}

static int check_status(dqa_t* h) {
  std::vector<int> st(h->cfg.batch);
  CK(cudaMemcpyAsync(st.data(), h->dev.status, sizeof(int) * st.size(), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (size_t i = 0; i < st.size(); ++i)
    if (st[i]) return fail(h, DQA_ENUMERIC, "instance %zu has status 0x%x (1 non-finite, 2 zero norm, 4 Jacobi not converged)", i, st[i]);
  return DQA_OK;
}

static int step_impl(dqa_t* h, int nsteps, bool check);
extern "C" int dqa_step(dqa_t* h, int nsteps) { return step_impl(h, nsteps, true); }
// same launches, no host synchronisation and no status check: lets several handles on different streams overlap
extern "C" int dqa_step_async(dqa_t* h, int nsteps) { return step_impl(h, nsteps, false); }
// One whole step (n_xi pattern sub-steps, mixer, energy) on h->stream.  p >= 0: step index known on the host; p < 0: the
// graph form, every kernel reads the step index from device memory.  With n_split > 1 the batch runs as n_split independent
// sub-batches: branch 0 on h->stream, the others on the handle's own streams, forked and joined with events -- inside a
// stream capture this makes parallel branches of the graph.  Profiling runs keep one branch (per-family event timers).
static int launch_step(dqa_t* h, int p) {
  const int ns = h->profiling ? 1 : h->n_split;
  const DqaDev full = h->dev;
  cudaStream_t main_stream = h->stream;
  struct Restore {
    dqa_t* h;
    const DqaDev& full;
    cudaStream_t st;
    ~Restore() {
      h->dev = full;
      h->stream = st;
    }
  } restore_{h, full, main_stream};
  if (ns > 1) {
    if (!h->ev_fork) CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (int s = 1; s < ns; ++s) {
      if (!h->split_stream[s - 1]) CK(cudaStreamCreateWithFlags(&h->split_stream[s - 1], cudaStreamNonBlocking));
      if (!h->ev_join[s - 1]) CK(cudaEventCreateWithFlags(&h->ev_join[s - 1], cudaEventDisableTiming));
    }
    CK(cudaEventRecord(h->ev_fork, main_stream));
  }
  for (int s = 0; s < ns; ++s) {
    if (ns > 1) {
      const size_t i0 = (size_t)full.batch * s / ns, i1 = (size_t)full.batch * (s + 1) / ns;
      h->dev = view_of(full, h->cfg, i0, (int)(i1 - i0));
      if (s > 0) {
        h->stream = h->split_stream[s - 1];
        CK(cudaStreamWaitEvent(h->stream, h->ev_fork, 0));
      }
    }
    for (int mu = 0; mu < h->cfg.n_xi; ++mu) {
      int rc = launch_canonize(h, mu);
      if (rc) return rc;
      rc = launch_truncate(h, p < 0 ? -1 : p, mu);
      if (rc) return rc;
    }
    int rc;
    if (p < 0) {
      rc = launch_mixer(h, 0.0, 0.0, 2);
    } else {
      const double s_p = (double)(p + 1) / h->cfg.P;
      const double beta = (1.0 - s_p) * h->cfg.dt;  // dQA_mps.py:88-89
      rc = launch_mixer(h, beta, 1.0 - s_p, full.ntab > 1 ? 1 : 0);  // several tables: beta = (1 - s_p) * dt of each instance
    }
    if (rc) return rc;
    rc = launch_energy(h, p < 0 ? -2 : p + 1);
    if (rc) return rc;
    if (s > 0) CK(cudaEventRecord(h->ev_join[s - 1], h->stream));
    h->stream = main_stream;
  }
  for (int s = 1; s < ns; ++s) CK(cudaStreamWaitEvent(main_stream, h->ev_join[s - 1], 0));
  return DQA_OK;
}

// One dQA step captured as a CUDA graph: the launch grid depends only on (N, n_xi, chi, batch), never on data, and every
// kernel that needs the step index reads it from device memory (p < 0), so the same executable graph replays for every
// step.  Pays off when the kernels are short (a single realisation, small bonds): 1 431 launches per step at N=21.
static int build_step_graph(dqa_t* h) {
  if (!h->cap_stream) CK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
  cudaStream_t user = h->stream;
  const long long l0 = h->launches;
  h->stream = h->cap_stream;
  h->capturing = true;
  cudaError_t e = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal);
  int rc = DQA_OK;
  if (e == cudaSuccess) {
    rc = launch_step(h, -1);
    if (rc == DQA_OK) HL(k_advance, dim3(1), dim3(32), 0, h->stream, h->dev);
    cudaGraph_t g = nullptr;
    e = cudaStreamEndCapture(h->cap_stream, &g);
    if (e == cudaSuccess && rc == DQA_OK) e = cudaGraphInstantiate(&h->step_graph, g, 0);
    if (g) cudaGraphDestroy(g);
  }
  h->stream = user;
  h->capturing = false;
  h->graph_nodes = h->launches - l0;
  h->launches = l0;
  if (e != cudaSuccess || rc != DQA_OK) {
    cudaGetLastError();
    h->step_graph = nullptr;
    h->use_graph = false;  // fall back to plain launches, which report their own errors
  }
  return DQA_OK;
}

static int step_impl(dqa_t* h, int nsteps, bool check) {
  NEED_READY();
  if (nsteps < 0 || h->pp + nsteps > h->cfg.P) return fail(h, DQA_EINVAL, "step %d + %d exceeds P=%d", h->pp, nsteps, h->cfg.P);
  if (h->use_graph && !h->profiling && nsteps > 0) {
    if (!h->step_graph) build_step_graph(h);
    if (h->step_graph) {
      const int p0 = h->pp;
      CK(cudaMemcpyAsync(h->dev.pdev, &p0, sizeof(int), cudaMemcpyHostToDevice, h->stream));
      for (int it = 0; it < nsteps; ++it) {
        CK(cudaGraphLaunch(h->step_graph, h->stream));
        h->launches += h->graph_nodes;
        h->pp++;
      }
      return check ? check_status(h) : DQA_OK;
    }
  }
  for (int it = 0; it < nsteps; ++it) {
    const int rc = launch_step(h, h->pp);
    if (rc) return rc;
    h->pp++;
  }
  return check ? check_status(h) : DQA_OK;
}

extern "C" int dqa_energy(dqa_t* h, double* eps) {
  NEED_READY();
  if (!eps) return DQA_EINVAL;
  // reduced into a scratch vector of its own (the eps trace is never touched), fetched with one contiguous copy
  int rc = launch_energy(h, -1);
  if (rc) return rc;
  CK(cudaMemcpyAsync(eps, h->dev.escratch, sizeof(cplx) * (size_t)h->dev.batch, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

// ---- state access ----------------------------------------------------------------------
extern "C" int dqa_get_bonds(dqa_t* h, int inst, int32_t* bonds) {
  NEED_WS();
  if (!bonds || inst < 0 || inst >= h->cfg.batch) return DQA_EINVAL;
  CK(cudaMemcpyAsync(bonds, h->dev.bond + (size_t)inst * (h->cfg.N + 1), sizeof(int) * (h->cfg.N + 1), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

extern "C" int dqa_get_mps(dqa_t* h, int inst, int site, double* buf, int* l, int* r) {
  NEED_WS();
  const int N = h->cfg.N, chi = h->cfg.chi;
  if (!buf || !l || !r || inst < 0 || inst >= h->cfg.batch || site < 0 || site >= N) return DQA_EINVAL;
  int b[2];
  CK(cudaMemcpyAsync(b, h->dev.bond + (size_t)inst * (N + 1) + site, sizeof(int) * 2, cudaMemcpyDeviceToHost, h->stream));
  std::vector<std::complex<double>> pad((size_t)2 * chi * chi);
  CK(cudaMemcpyAsync(pad.data(), h->dev.mps + ((size_t)inst * N + site) * 2 * chi * chi, sizeof(cplx) * pad.size(), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  *l = b[0];
  *r = b[1];
  std::complex<double>* out = reinterpret_cast<std::complex<double>*>(buf);
  for (int a = 0; a < b[0]; ++a)
    for (int s = 0; s < 2; ++s)
      for (int c = 0; c < b[1]; ++c) out[((size_t)a * 2 + s) * b[1] + c] = pad[((size_t)a * 2 + s) * chi + c];
  return DQA_OK;
}

extern "C" int dqa_set_mps(dqa_t* h, int inst, int site, const double* buf, int l, int r) {
  NEED_WS();
  const int N = h->cfg.N, chi = h->cfg.chi;
  if (!buf || inst < 0 || inst >= h->cfg.batch || site < 0 || site >= N) return DQA_EINVAL;
  if (l < 1 || r < 1 || l > chi || r > chi) return fail(h, DQA_ELIMIT, "site %d bonds (%d,%d) exceed chi=%d", site, l, r, chi);
  if ((site == 0 && l != 1) || (site == N - 1 && r != 1)) return fail(h, DQA_EINVAL, "boundary bonds must be 1");
  {
    // the sector kernels size their shared memory for bonds <= min(chi, 2^n, 2^(N-n)) (anything larger is
    // rank-deficient padding; compress such a state first)
    const int bn[2] = {site, site + 1}, bv[2] = {l, r};
    for (int e = 0; e < 2; ++e) {
      long cap = chi;
      if (bn[e] < 30 && (1L << bn[e]) < cap) cap = 1L << bn[e];
      if (N - bn[e] < 30 && (1L << (N - bn[e])) < cap) cap = 1L << (N - bn[e]);
      if (bv[e] > cap) return fail(h, DQA_ELIMIT, "bond %d has dimension %d > min(chi, 2^n, 2^(N-n)) = %ld", bn[e], bv[e], cap);
    }
  }
  std::vector<std::complex<double>> pad((size_t)2 * chi * chi, std::complex<double>(0.0, 0.0));
  const std::complex<double>* in = reinterpret_cast<const std::complex<double>*>(buf);
  for (int a = 0; a < l; ++a)
    for (int s = 0; s < 2; ++s)
      for (int c = 0; c < r; ++c) pad[((size_t)a * 2 + s) * chi + c] = in[((size_t)a * 2 + s) * r + c];
  int b[2] = {l, r};
  CK(cudaMemcpyAsync(h->dev.mps + ((size_t)inst * N + site) * 2 * chi * chi, pad.data(), sizeof(cplx) * pad.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->dev.bond + (size_t)inst * (N + 1) + site, b, sizeof(int) * 2, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->dev.status + inst, 0, sizeof(int), h->stream));  // a freshly loaded state starts healthy
  CK(cudaStreamSynchronize(h->stream));
  h->have_state = true;
  return DQA_OK;
}

extern "C" int dqa_get_schmidt(dqa_t* h, int inst, int bond, double* s, int* n) {
  NEED_WS();
  const int N = h->cfg.N, chi = h->cfg.chi;
  if (!s || !n || inst < 0 || inst >= h->cfg.batch || bond < 0 || bond >= N - 1) return DQA_EINVAL;
  CK(cudaMemcpyAsync(n, h->dev.nsv + (size_t)inst * N + bond, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(s, h->dev.schmidt + ((size_t)inst * N + bond) * 2 * chi, sizeof(double) * 2 * chi, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

extern "C" int dqa_get_eps(dqa_t* h, double* eps) {
  NEED_WS();
  if (!eps) return DQA_EINVAL;
  CK(cudaMemcpyAsync(eps, h->dev.eps, sizeof(cplx) * (size_t)h->cfg.batch * (h->cfg.P + 1), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}
extern "C" int dqa_get_bd_monitor(dqa_t* h, int32_t* bd) {
  NEED_WS();
  if (!bd) return DQA_EINVAL;
  CK(cudaMemcpyAsync(bd, h->dev.bdmon, sizeof(int) * (size_t)h->cfg.batch * h->cfg.P * h->cfg.n_xi, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}
extern "C" int dqa_get_status(dqa_t* h, int32_t* st) {
  NEED_WS();
  if (!st) return DQA_EINVAL;
  CK(cudaMemcpyAsync(st, h->dev.status, sizeof(int) * (size_t)h->cfg.batch, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}
extern "C" int dqa_clear_status(dqa_t* h) {
  NEED_WS();
  CK(cudaMemsetAsync(h->dev.status, 0, sizeof(int) * (size_t)h->cfg.batch, h->stream));
  return DQA_OK;
}
extern "C" int dqa_get_truncation_error(dqa_t* h, int inst, double* w) {
  NEED_WS();
  if (!w || inst < 0 || inst >= h->cfg.batch) return DQA_EINVAL;
  CK(cudaMemcpyAsync(w, h->dev.discw + (size_t)inst * h->cfg.N, sizeof(double) * (size_t)(h->cfg.N - 1), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}
extern "C" int dqa_get_counters(dqa_t* h, uint64_t* c, int reset) {
  NEED_WS();
  if (!c) return DQA_EINVAL;
  CK(cudaMemcpyAsync(c, h->dev.counters, sizeof(uint64_t) * DQA_NCNT * (size_t)h->cfg.batch, cudaMemcpyDeviceToHost, h->stream));
  if (reset) CK(cudaMemsetAsync(h->dev.counters, 0, sizeof(uint64_t) * DQA_NCNT * (size_t)h->cfg.batch, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

extern "C" int dqa_state_bytes(const dqa_t* h, size_t* bytes) {
  if (!h || !bytes) return DQA_EINVAL;
  const size_t B = h->cfg.batch, N = h->cfg.N, chi = h->cfg.chi;
  *bytes = B * N * 2 * chi * chi * sizeof(cplx) + B * (N + 1) * sizeof(int);
  return DQA_OK;
}
extern "C" int dqa_download_state(dqa_t* h, void* host) {
  NEED_WS();
  if (!host) return DQA_EINVAL;
  const size_t B = h->cfg.batch, N = h->cfg.N, chi = h->cfg.chi;
  const size_t nm = B * N * 2 * chi * chi * sizeof(cplx), nb = B * (N + 1) * sizeof(int);
  CK(cudaMemcpyAsync(host, h->dev.mps, nm, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync((char*)host + nm, h->dev.bond, nb, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}
extern "C" int dqa_upload_state(dqa_t* h, const void* host) {
  NEED_WS();
  if (!host) return DQA_EINVAL;
  const size_t B = h->cfg.batch, N = h->cfg.N, chi = h->cfg.chi;
  const size_t nm = B * N * 2 * chi * chi * sizeof(cplx), nb = B * (N + 1) * sizeof(int);
  CK(cudaMemcpyAsync(h->dev.mps, host, nm, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->dev.bond, (const char*)host + nm, nb, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->dev.status, 0, sizeof(int) * B, h->stream));  // a freshly loaded state starts healthy
  h->have_state = true;
  return DQA_OK;
}
extern "C" int dqa_get_launch_count(const dqa_t* h, int64_t* n) {
  if (!h || !n) return DQA_EINVAL;
  *n = h->launches;
  return DQA_OK;
}
extern "C" int dqa_sync(dqa_t* h) {
  if (!h) return DQA_EINVAL;
  DevGuard guard_(h->cfg.device);
  CK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

// ---- ensemble moments (input of the one NCCL all-reduce) -------------------------------
__global__ void k_ensemble_moments(DqaDev d, double* out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > d.P) return;
  double m1 = 0.0, m2 = 0.0, ok = 0.0;
  for (int b = 0; b < d.batch; ++b) {
    if (d.status[b]) continue;
    const double v = d.eps[(size_t)b * (d.P + 1) + s].x;
    m1 += v;
    m2 += v * v;
    ok += 1.0;
  }
  out[s] = m1;
  out[(d.P + 1) + s] = m2;
  out[2 * (d.P + 1) + s] = ok;
}

extern "C" int dqa_ensemble_moments(dqa_t* h, double* device_out) {
  NEED_WS();
  if (!device_out) return DQA_EINVAL;
  const int n = h->cfg.P + 1;
  HL(k_ensemble_moments, dim3((n + 127) / 128), dim3(128), 0, h->stream, h->dev, device_out);
  CK(cudaGetLastError());
  return DQA_OK;
}

// ---- FP64 peak microbenchmark ----------------------------------------------------------
__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// same probe on the FP64 tensor path: 8 independent m8n8k4 accumulator tiles per warp
__global__ void k_dmma_peak(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-9 + i;
  const double a = 1.0000001, b = 1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dmma884(c[2 * j], c[2 * j + 1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static int fp64_probe(dqa_t* h, double* tflops, bool tensor) {
  NEED_WS();
  if (!tflops) return DQA_EINVAL;
  // scratch: the LQ work buffer (>= 2*chi*ld_theta*batch complex)
  const size_t cap = (size_t)h->cfg.batch * 2 * h->cfg.chi * h->dev.ld_theta * 2;  // doubles
  int blocks = 148 * 8, threads = 256;
  while ((size_t)blocks * threads > cap && blocks > 1) blocks /= 2;
  if ((size_t)blocks * threads > cap) return fail(h, DQA_ELIMIT, "workspace too small for the FP64 probe");
#ifdef DQA_EMUL
  *tflops = 0.0;
  return DQA_OK;
#endif
  const int iters = 1 << 16;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a, h->stream));
    if (tensor) HL(k_dmma_peak, dim3(blocks), dim3(threads), 0, h->stream, (double*)h->dev.work, iters);
    else HL(k_fp64_peak, dim3(blocks), dim3(threads), 0, h->stream, (double*)h->dev.work, iters);
    CK(cudaEventRecord(b, h->stream));
    CK(cudaEventSynchronize(b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    // DFMA: 8 fma per thread and iteration; DMMA: 8 tiles of 8x8x4 fma per warp and iteration
    const double fl = tensor ? 2.0 * 8.0 * 256.0 * iters * (double)blocks * (threads / 32) : 2.0 * 8.0 * iters * (double)blocks * threads;
    if (ms > 0.f && fl / (ms * 1e-3) / 1e12 > best) best = fl / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *tflops = best;
  return DQA_OK;
}

extern "C" int dqa_measure_fp64_peak(dqa_t* h, double* tflops) { return fp64_probe(h, tflops, false); }
extern "C" int dqa_measure_dmma_peak(dqa_t* h, double* tflops) { return fp64_probe(h, tflops, true); }

#include "dqa_ops_abi.cuh"
