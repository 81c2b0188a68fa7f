// dqa_common.cuh -- shared definitions for the sm_100a dQA kernels.
//
// Built two ways: nvcc (the product, libdqa.so) and, with -DDQA_EMUL, g++ on top of
// tests/emul/cuda_emul.h (a fiber SIMT emulator used only by the CPU test-suite to
// debug kernel logic; never loaded by the package).
#pragma once

#ifdef DQA_EMUL
#include "cuda_emul.h"
#define DQA_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::emul::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define DQA_DYN_SMEM(name) char* name = ::emul::dyn_smem_
#else
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define DQA_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DQA_DYN_SMEM(name) extern __shared__ __align__(16) char name[]
#endif

#include <stdint.h>

typedef double2 cplx;

#define DQA_FULL 0xffffffffu

// ---------------------------------------------------------------------------------
// complex128 helpers (interleaved re,im == the reference's numpy complex128 layout)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ cplx cmk(double x, double y) { return make_double2(x, y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return cmk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return cmk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cmk(a.x * s, a.y * s); }
__device__ __forceinline__ cplx cconj(cplx a) { return cmk(a.x, -a.y); }
// acc + a*b
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx acc) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
  return acc;
}
// acc + a*conj(b)
__device__ __forceinline__ cplx cfmac(cplx a, cplx b, cplx acc) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y);
  acc.y = fma(-a.x, b.y, acc.y);
  return acc;
}
// acc - a*b
__device__ __forceinline__ cplx cfms(cplx a, cplx b, cplx acc) {
  acc.x = fma(-a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
  return acc;
}
__device__ __forceinline__ double cabs2(cplx a) { return fma(a.x, a.x, a.y * a.y); }
__device__ __forceinline__ cplx crecip(cplx a) {
  double d = 1.0 / cabs2(a);
  return cmk(a.x * d, -a.y * d);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DQA_FULL, v, o);
  return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(DQA_FULL, v.x, o);
    v.y += __shfl_xor_sync(DQA_FULL, v.y, o);
  }
  return v;
}

// Sum 16 independent values across the warp with 16 shuffles instead of 80: a butterfly that halves the
// payload at every stage.  On return lane l holds the warp-wide total of slot (l >> 1) (both lanes of a pair).
__device__ __forceinline__ double warp_sum16(const double (&v)[16]) {
  const int lane = threadIdx.x & 31;
  double a[8], b[4], c[2], dd;
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    const double send = h4 ? v[s] : v[s + 8];
    const double keep = h4 ? v[s + 8] : v[s];
    a[s] = keep + __shfl_xor_sync(DQA_FULL, send, 16);
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const double send = h3 ? a[s] : a[s + 4];
    const double keep = h3 ? a[s + 4] : a[s];
    b[s] = keep + __shfl_xor_sync(DQA_FULL, send, 8);
  }
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const double send = h2 ? b[s] : b[s + 2];
    const double keep = h2 ? b[s + 2] : b[s];
    c[s] = keep + __shfl_xor_sync(DQA_FULL, send, 4);
  }
  {
    const double send = h1 ? c[0] : c[1];
    const double keep = h1 ? c[1] : c[0];
    dd = keep + __shfl_xor_sync(DQA_FULL, send, 2);
  }
  dd += __shfl_xor_sync(DQA_FULL, dd, 1);
  return dd;  // slot index = 8*h4 + 4*h3 + 2*h2 + h1 = lane >> 1
}

// ---------------------------------------------------------------------------------
// FP64 tensor-core tile: D(8x8) += A(8x4) * B(4x8), mma.sync.m8n8k4.f64 (DMMA).
// Fragment ownership (PTX ISA, .f64 m8n8k4): g = lane>>2, t = lane&3
//   a = A[g][t],  b = B[t][g],  (c0,c1) = C[g][2t], C[g][2t+1].
// Complex128 tiles are four real DMMAs on the (re,im) parts of the interleaved operands.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
#ifdef DQA_EMUL
  ::emul::dmma884(c0, c1, a, b);
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
#endif
}

struct ZTile {  // one lane's share of an 8x8 complex accumulator tile
  double r0, r1, i0, i1;
};
__device__ __forceinline__ ZTile ztile_zero() {
  ZTile z;
  z.r0 = z.r1 = z.i0 = z.i1 = 0.0;
  return z;
}
// acc += A(8 x K) * B(K x 8); a_at(row, k) and b_at(k, col) return complex values (0 outside
// the operand), K is rounded up to a multiple of 4 by the caller's accessors returning 0.
template <class FA, class FB>
__device__ __forceinline__ void warp_zgemm_8x8(ZTile& acc, int K, FA a_at, FB b_at) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 4
  for (int k0 = 0; k0 < K; k0 += 4) {
    const cplx a = a_at(g, k0 + t);
    const cplx b = b_at(k0 + t, g);
    dmma884(acc.r0, acc.r1, a.x, b.x);
    dmma884(acc.r0, acc.r1, -a.y, b.y);
    dmma884(acc.i0, acc.i1, a.x, b.y);
    dmma884(acc.i0, acc.i1, a.y, b.x);
  }
}

// NT row tiles against the same B operand: acc[i] += A_i(8 x K) * B(K x 8).  The B fragment is loaded once per
// k-step and feeds 4*NT DMMAs (the single-tile form loads two fragments per 4 DMMAs and saturates the L1 data pipe).
template <int NT, class FA, class FB>
__device__ __forceinline__ void warp_zgemm_8x8_rows(ZTile (&acc)[NT], int K, FA a_at, FB b_at) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll(NT == 1 ? 4 : 2)
  for (int k0 = 0; k0 < K; k0 += 4) {
    const cplx b = b_at(k0 + t, g);
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const cplx a = a_at(i, g, k0 + t);
      dmma884(acc[i].r0, acc[i].r1, a.x, b.x);
      dmma884(acc[i].r0, acc[i].r1, -a.y, b.y);
      dmma884(acc[i].i0, acc[i].i1, a.x, b.y);
      dmma884(acc[i].i0, acc[i].i1, a.y, b.x);
    }
  }
}

// ---------------------------------------------------------------------------------
// device-side view of one plan (all pointers are caller-owned device memory, carved
// out of the workspace bound with dqa_bind_workspace)
// ---------------------------------------------------------------------------------
struct DqaDev {
  int N, D, n_xi, chi, batch, P;
  int smax;        // N+1 sectors at most per bond
  int ld_theta;    // row stride of Theta (>= chi*(N+1))
  int lq_pb;       // rows per LQ panel (8, or fewer when a panel of qmax columns would not fit in shared memory)
  int qmax;        // bound on the live columns of Theta: max_n min(chi,2^n,2^(N-n)) * (N-n+1)
  int nsec_total;  // number of (site, sector) Q blocks per instance
  int kh;          // D/2+1: independent Fourier modes of the energy sweep
  int max_sweeps;
  int cutoff_mode;  // 1 abs, 2 rel, 3 sum2, 4 rsum2, 5 sum1, 6 rsum1 (lib/tenn.py:317-345); the reference hard-codes 4
  int renorm;       // 0: kept values not rescaled; > 0: rescaled so their p-norm is preserved (lib/tenn.py:347-354); reference: 2
  int ntab;         // number of Fourier tables (1 unless instances run different dt)
  double cutoff;
  double jac_quad2;  // (relative off-diagonal)^2 below which a rotated pair counts as already converged (0: always run the checking sweep)
  cplx* mps;              // [batch][N][chi*2*chi]   site (a,s,c) at (a*2+s)*chi + c
  int* bond;              // [batch][N+1]
  const uint8_t* hbit;    // [batch][n_xi][N]        (1 - xi)/2
  const cplx* uz;         // [ntab][P][D]            u_p(x) = exp(-i gamma_p f(x)), one table per distinct dt
  const int* tab;         // [batch]                 Fourier table of each instance
  const double* dtv;      // [batch]                 dt of each instance (mixer angle), only read when ntab > 1
  int* pdev;              // [1]                     the step counter p on the device: kernels launched with p < 0 read it, so a
                          //                         captured CUDA graph of one step replays for every step
  const double* mixtab;   // [P][2]                  cos, sin of the mixer angle (1 - s_p) dt, host-computed (single-dt plans)
  const cplx* gk;         // [D]                     fxft[k]/sqrt(D)
  const cplx* phase;      // [D]                     exp(2 pi i k / D) = exp(i pi q / D), q = 2k
  cplx* rbuf;             // [batch][2][smax][chi*chi]     R_{n,y} row-major (q x b_n), ld chi
  cplx* qbuf;             // [batch][nsec_total][2chi*chi] Q_{n,y} col-major (m x q),  ld 2chi
  int* qdim;              // [batch][N+1][smax]            q_{n,y}
  cplx* theta;            // [batch][2][2chi*ld_theta]     Theta_l row-major, ping-pong on l
  cplx* work;             // [batch][2chi*ld_theta]        LQ scratch
  cplx* lbuf;             // [batch][2chi*2chi]            L factor of Theta_l, col-major ld 2chi
  double* scale;          // [batch]                       renorm factor of the last truncation
  double* schmidt;        // [batch][N][2chi]              singular values seen at cut l
  int* nsv;               // [batch][N]
  cplx* tmk;              // [batch][n_xi][kh]             <psi|Phi_{mu,k}|psi>
  cplx* eps;              // [batch][P+1]
  cplx* escratch;         // [batch]                       dqa_energy's result (never aliases the trace)
  double* discw;          // [batch][N]                    relative discarded weight of the last truncation at cut l
  int* bdmon;             // [batch][P*n_xi]
  int* status;            // [batch]  bit0: non-finite, bit1: zero norm, bit2: Jacobi not converged
  unsigned long long* counters;  // [batch][DQA_NCNT] rotations, sweeps, svds, qr columns, flops: LQ, Jacobi, sector QR, theta GEMMs; thread-0 cycles of sector QR phases (GEMM+load, factor, R/Q out) and #CTAs; of LQ phases (panel, T, trailing) and #CTAs
};

__host__ __device__ __forceinline__ int dqa_secbase(int N, int n) {
  // offset of site n's first Q block; site n (1..N-1) has sectors y = 0..N-n
  return (n - 1) * (N + 1) - (n - 1) * n / 2;
}

#define DQA_NCNT 16  // counters per instance

enum {
  DQA_ST_NONFINITE = 1,
  DQA_ST_ZERONORM = 2,
  DQA_ST_NOCONV = 4,
  DQA_ST_LIMIT = 8,
};
