// dqa_sv.cuh -- exact (untruncated) state-vector dQA on the GPU: SURVEY 8f-3, the gauge-free validator.
// Same dynamics as the MPS path with chi = infinity: for each step p
//     psi(sigma) <- exp(-i gamma_p sum_mu f(x_mu(sigma))) psi(sigma)          (all U_z^mu at once: they commute)
//     psi <- prod_i exp(+i beta_p sigma^x_i) psi                              (RX on every qubit)
//     eps = <psi| sum_mu f(x_mu) |psi> / N
// Conventions (lib/HammingEvolution.py:133-148): site 0 is the most significant bit, s=0 <-> sigma=+1,
// x_mu = popcount(sigma XOR H_mu) with H_mu holding hbit[mu,i] at bit N-1-i.  HBM-bound: every kernel
// streams the 16*2^N-byte vector; the mixer handles up to 10 qubits per pass in shared-memory tiles.
#pragma once
#include "dqa_common.cuh"

__device__ __forceinline__ double sv_cost(unsigned long long sigma, const unsigned long long* pat, int n_xi, const double* fx) {
  double h = 0.0;
  for (int mu = 0; mu < n_xi; ++mu) h += fx[__popcll(sigma ^ pat[mu])];
  return h;
}

__global__ void k_sv_init(cplx* psi, unsigned long long dim, double amp) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += (unsigned long long)gridDim.x * blockDim.x)
    psi[i] = cmk(amp, 0.0);
}

// diagonal phase; pat (n_xi <= 64 patterns) and fx (N+1 <= 65 values) staged in shared memory
__global__ void k_sv_phase(cplx* psi, unsigned long long dim, const unsigned long long* pat, int n_xi, const double* fx, int nfx, double gamma) {
  __shared__ unsigned long long sp[64];
  __shared__ double sf[72];
  if (threadIdx.x < n_xi) sp[threadIdx.x] = pat[threadIdx.x];
  if (threadIdx.x < nfx) sf[threadIdx.x] = fx[threadIdx.x];
  __syncthreads();
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += (unsigned long long)gridDim.x * blockDim.x) {
    const double ang = -gamma * sv_cost(i, sp, n_xi, sf);
    double sn, cs;
    sincos(ang, &sn, &cs);
    psi[i] = cmul(psi[i], cmk(cs, sn));
  }
}

// RX(beta) = [[c, is],[is, c]] on bit positions q0 .. q0+nq-1 (bit 0 = least significant = site N-1).
// A tile is 2^nq target values x `low` contiguous amplitudes (low = min(2^q0, 32)), staged in shared memory.
__global__ void k_sv_rx(cplx* psi, int nbits, int q0, int nq, int lowbits, double c, double s) {
  DQA_DYN_SMEM(smraw);
  cplx* tile = (cplx*)smraw;
  const int tsz = 1 << nq, low = 1 << lowbits, tot = tsz * low;
  // enumerate the bits that are neither target nor low: tile index -> base address
  const unsigned long long ntiles = 1ull << (nbits - nq - lowbits);
  for (unsigned long long tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
    // scatter tix's bits into positions [lowbits, q0) and [q0+nq, nbits)
    const int midbits = q0 - lowbits;
    const unsigned long long mid = tix & ((1ull << midbits) - 1ull), hi = tix >> midbits;
    const unsigned long long base = (hi << (q0 + nq)) | (mid << lowbits);
    for (int i = threadIdx.x; i < tot; i += blockDim.x) {
      const int t = i / low, j = i % low;
      tile[i] = psi[base | ((unsigned long long)t << q0) | (unsigned long long)j];
    }
    __syncthreads();
    for (int b = 0; b < nq; ++b) {
      for (int i = threadIdx.x; i < tot / 2; i += blockDim.x) {
        const int j = i % low, tp = i / low;                       // tp enumerates target values with bit b clear
        const int t0 = ((tp >> b) << (b + 1)) | (tp & ((1 << b) - 1));
        const int i0 = t0 * low + j, i1 = (t0 | (1 << b)) * low + j;
        const cplx a0 = tile[i0], a1 = tile[i1];
        tile[i0] = cmk(c * a0.x - s * a1.y, c * a0.y + s * a1.x);   // c a0 + i s a1
        tile[i1] = cmk(c * a1.x - s * a0.y, c * a1.y + s * a0.x);   // i s a0 + c a1
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < tot; i += blockDim.x) {
      const int t = i / low, j = i % low;
      psi[base | ((unsigned long long)t << q0) | (unsigned long long)j] = tile[i];
    }
    __syncthreads();
  }
}

// partial sums of |psi|^2 * H(sigma): one value per block (deterministic two-stage reduction)
__global__ void k_sv_energy(const cplx* psi, unsigned long long dim, const unsigned long long* pat, int n_xi, const double* fx, int nfx, double* partial) {
  __shared__ unsigned long long sp[64];
  __shared__ double sf[72];
  __shared__ double red[64];
  if (threadIdx.x < n_xi) sp[threadIdx.x] = pat[threadIdx.x];
  if (threadIdx.x < nfx) sf[threadIdx.x] = fx[threadIdx.x];
  __syncthreads();
  double acc = 0.0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += (unsigned long long)gridDim.x * blockDim.x)
    acc += cabs2(psi[i]) * sv_cost(i, sp, n_xi, sf);
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void k_sv_energy_final(const double* partial, int n, double inv_n, double* out) {
  __shared__ double red[64];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = acc * inv_n;
}
