// dqa_kernels.cuh -- hand-written sm_100a kernels of the dQA MPS step (fused "sector" path).
//
// One dQA sub-step for pattern mu (reference: dQA_mps.py:92-106) is
//     psi <- truncate_chi( right_canonise( U_z^mu psi ) )
// with U_z^mu = exp(-i gamma_p f(X_mu)) (lib/dQA_utils.py:32-75 builds it as a bond-(N+1)
// Fourier MPO; X_mu = Hamming distance to pattern mu).  Instead of the reference's
// bond-(chi*(N+1)) direct-sum tensors we label the bond between sites n-1 and n by
// (c, y): c = MPS bond index, y = number of mismatches on sites >= n.  Different y are
// orthogonal, so:
//   k_sector_qr   right-canonisation (lib/tenn.py:152-190) = independent Householder QRs,
//                 one CTA per (site, sector y, instance), operand (q_y + q_{y-1}) x b in SMEM;
//   k_theta0/k_theta  build the centre matrix Theta_l (2c_l x sum_y q_y) of the left-to-right
//                 sweep from the sector Q blocks (lib/tenn.py:248 tensordot, block-sparse);
//   k_svd         per instance: Householder LQ of Theta_l, one-sided Jacobi SVD of the
//                 (2c x 2c) factor with warp-shuffle reductions, the reference's trim rule
//                 (cutoff 1e-10 on relative cumulative s^2, clamp to chi, renormalise;
//                 lib/tenn.py:317-354) and the new left-orthonormal site tensor.
//   k_mixer       exp(+i beta sigma^x) on every physical leg (lib/dQA_utils.py:26-30).
//   k_energy      transfer-matrix sweep for <psi|Phi_{mu,k}|psi> (dQA_mps.py:68-81,
//                 lib/tenn.py:93-109), one CTA per (k, mu, instance), using
//                 T_{mu,D-k} = conj(T_{mu,k}).
// All arithmetic is complex128; only gauge-invariant results are promised to match the
// reference.
#pragma once
#include "dqa_common.cuh"

// ---------------------------------------------------------------------------------
// block-wide sum, result broadcast to every thread (red: >= 33 doubles of shared memory)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect red[] against the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nwarp; ++w) t += red[w];
  return t;
}

// Householder generator (LAPACK zlarfg semantics without the safmin loop): given alpha and
// ss = |x[1:]|^2 returns beta (real), tau, and scal = 1/(alpha-beta); H^H [alpha;x] = [beta;0],
// H = I - tau v v^H, v = [1; x*scal].
__device__ __forceinline__ void householder_gen(cplx alpha, double ss, double& beta, cplx& tau, cplx& scal) {
  if (ss == 0.0 && alpha.y == 0.0) {
    beta = alpha.x;
    tau = cmk(0.0, 0.0);
    scal = cmk(0.0, 0.0);
  } else {
    double nrm = sqrt(cabs2(alpha) + ss);
    beta = (alpha.x >= 0.0) ? -nrm : nrm;
    tau = cmk((beta - alpha.x) / beta, -alpha.y / beta);
    scal = crecip(cmk(alpha.x - beta, alpha.y));
  }
}

// Same as householder_gen but with one rsqrt and one reciprocal on the critical path instead of sqrt + 3 divisions.
__device__ __forceinline__ void householder_gen_fast(cplx alpha, double ss, double& beta, cplx& tau, cplx& scal) {
  if (ss == 0.0 && alpha.y == 0.0) {
    beta = alpha.x;
    tau = cmk(0.0, 0.0);
    scal = cmk(0.0, 0.0);
  } else {
    const double s2 = cabs2(alpha) + ss;
    const double rs = rsqrt(s2);
    const double nrm = s2 * rs;
    const double ib = (alpha.x >= 0.0) ? -rs : rs;  // 1/beta
    beta = (alpha.x >= 0.0) ? -nrm : nrm;
    tau = cmk(1.0 - alpha.x * ib, -alpha.y * ib);
    const double dx = alpha.x - beta;
    const double inv = 1.0 / fma(dx, dx, alpha.y * alpha.y);
    scal = cmk(dx * inv, -alpha.y * inv);
  }
}

// ---------------------------------------------------------------------------------
// integer pattern tables (implicit in lib/dQA_utils.py:50-53,106)
//   hbit = (1-xi)/2, e = 1 - xi*(-1)^s = 2*(hbit^s), q = k*e.  Pure integer arithmetic.
// ---------------------------------------------------------------------------------
__global__ void k_hamming_tables(const int8_t* xi, int n_items, int D, uint8_t* hbit, int32_t* e, int32_t* q) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += gridDim.x * blockDim.x) {
    const int x = xi[i];
    const int hb = (1 - x) / 2;
    if (hbit) hbit[i] = (uint8_t)hb;
    for (int s = 0; s < 2; ++s) {
      const int es = 1 - x * (s ? -1 : 1);
      if (e) e[i * 2 + s] = es;
      if (q)
        for (int k = 0; k < D; ++k) q[(i * 2 + s) * D + k] = k * es;
    }
  }
}

// |+>^N, bond 1 everywhere (dQA_mps.py:139); also resets the per-instance status words.
__global__ void k_init_plus(DqaDev d) {
  const int inst = blockIdx.x;
  const int site_sz = d.chi * 2 * d.chi;
  cplx* m = d.mps + (size_t)inst * d.N * site_sz;
  for (int i = threadIdx.x; i < d.N * site_sz; i += blockDim.x) {
    const int r = i % site_sz;
    m[i] = (r == 0 || r == d.chi) ? cmk(0.70710678118654757, 0.0) : cmk(0.0, 0.0);
  }
  for (int i = threadIdx.x; i <= d.N; i += blockDim.x) d.bond[inst * (d.N + 1) + i] = 1;
  if (threadIdx.x == 0) {
    d.status[inst] = 0;
    d.scale[inst] = 1.0;
  }
}

// ---------------------------------------------------------------------------------
// k_sector_qr: one (site n, sector y, instance) block of the right-canonisation sweep.
//   M[(s,j), a] = sum_c R_{n+1, y-nbit(s)}[j,c] * A_n[a,s,c]      (m x b_n, m = q0+q1)  -- DMMA tiles
//   M = Q R (Householder; R diagonal made real >= 0 as qr_stabilized does, with sgn(0):=+1)
// Mapping: one 8-lane group per column; the group keeps its column in registers for the whole
// factorisation (rows li, li+8, ...), reflector k is published to shared memory by the group
// that owns column k as soon as it has applied reflector k-1 (one __syncthreads per column),
// dot products are 3-stage __shfl_xor reductions.  Q = H_0...H_{q-1} I is then formed column by
// column with no block barrier at all (column j only needs reflectors j..0).
// shared memory: M (2chi x chi, col-major; column k ends up holding reflector v_k) | tau | beta
// ---------------------------------------------------------------------------------
#ifndef SQR_TWO_UPTO
#define SQR_TWO_UPTO 12  // largest slot count compiled for two CTAs per SM (12 slots spill at 85 registers and still gain 13 %)
#endif
#ifndef SQR_MINB
#define SQR_MINB 3  // CTAs per SM the 8-slot sector-QR kernel is compiled for
#endif
__host__ __device__ inline int sector_qr_maxs(int chi) { return chi <= 8 ? 2 : (chi <= 12 ? 3 : (chi <= 16 ? 4 : (chi <= 24 ? 6 : (chi <= 32 ? 8 : (chi <= 40 ? 10 : (chi <= 48 ? 12 : 16)))))); }
__host__ __device__ inline size_t sector_qr_smem(int chi) {  // M has 8*MAXS rows so every register slot has a row
  return sizeof(cplx) * ((size_t)8 * sector_qr_maxs(chi) * chi + chi) + sizeof(double) * chi;
}

// Reflector k is stored in shared memory as the FULL vector over the live register slots
// (0 above row k, 1 at row k, the scaled entries below, 0 beyond row m-1), so applying it needs no
// per-row conditions: w = sum conj(v) col, col -= tau w v.  Slots t < k>>3 hold only finished rows and
// are skipped by jumping into the unrolled slot sequence (fall-through switch: one copy of the code).
template <int MAXS>
__device__ __forceinline__ void sqr_publish(cplx (&col)[MAXS], int k, int m, int j, int li, cplx* Mk, cplx* tau, double* betas) {
  // executed by every lane of the warp that contains group k; only group k == j stores
  double ss = 0.0;
  cplx alpha = cmk(0.0, 0.0);
#pragma unroll
  for (int t = 0; t < MAXS; ++t) {
    const int row = li + 8 * t;
    if (row > k && row < m) ss += cabs2(col[t]);
    if (row == k) alpha = col[t];
  }
  // the three reductions are independent: issue them together (alpha lives in lane k&7 of the group)
  const int src = (threadIdx.x & 24) | (k & 7);
  alpha.x = __shfl_sync(DQA_FULL, alpha.x, src);
  alpha.y = __shfl_sync(DQA_FULL, alpha.y, src);
  ss += __shfl_xor_sync(DQA_FULL, ss, 4);
  ss += __shfl_xor_sync(DQA_FULL, ss, 2);
  ss += __shfl_xor_sync(DQA_FULL, ss, 1);
  if (j != k) return;
  // zlarfg without divisions on the critical path: 1/beta = -+rsqrt(|alpha|^2+ss), one reciprocal for 1/(alpha-beta)
  double beta;
  cplx t_, scal;
  if (ss == 0.0 && alpha.y == 0.0) {
    beta = alpha.x;
    t_ = cmk(0.0, 0.0);
    scal = cmk(0.0, 0.0);
  } else {
    const double s2 = cabs2(alpha) + ss;
    const double rs = rsqrt(s2);
    const double nrm = s2 * rs;
    const double ib = (alpha.x >= 0.0) ? -rs : rs;  // 1/beta
    beta = (alpha.x >= 0.0) ? -nrm : nrm;
    t_ = cmk(1.0 - alpha.x * ib, -alpha.y * ib);
    const double dx = alpha.x - beta;
    const double inv = 1.0 / fma(dx, dx, alpha.y * alpha.y);
    scal = cmk(dx * inv, -alpha.y * inv);
  }
#pragma unroll
  for (int t = 0; t < MAXS; ++t) {
    const int row = li + 8 * t;
    cplx v = cmk(0.0, 0.0);
    if (row > k && row < m) {
      col[t] = cmul(col[t], scal);
      v = col[t];
    }
    if (row == k) {
      col[t] = cmk(beta, 0.0);
      v = cmk(1.0, 0.0);
    }
    if (row >= (k & ~7)) Mk[row] = v;
  }
  if (li == 0) {
    tau[k] = t_;
    betas[k] = beta;
  }
}

// The reflector is read again for the update instead of being kept in registers across the reduction
// (measured: 80 registers without spills, -9 % kernel time).  Splitting the dot product over two
// accumulators was measured too: no gain.
#define SQR_DOT(T) \
  case T:          \
    if (T < MAXS) w = cfmac(col[T < MAXS ? T : 0], vk[li + 8 * T], w);
#define SQR_UPD(T) \
  case T:          \
    if (T < MAXS) col[T < MAXS ? T : 0] = cfms(tw, vk[li + 8 * T], col[T < MAXS ? T : 0]);

// col <- (I - tk v v^H) col for the group's column; t0 = first live slot of reflector v
template <int MAXS>
__device__ __forceinline__ void sqr_apply(cplx (&col)[MAXS], const cplx* vk, cplx tk, int t0, int li, bool act) {
  cplx w = cmk(0.0, 0.0);
  switch (t0) {  // fall through on purpose
    SQR_DOT(0) SQR_DOT(1) SQR_DOT(2) SQR_DOT(3) SQR_DOT(4) SQR_DOT(5)
    SQR_DOT(6) SQR_DOT(7) SQR_DOT(8) SQR_DOT(9) SQR_DOT(10) SQR_DOT(11)
    SQR_DOT(12) SQR_DOT(13) SQR_DOT(14) SQR_DOT(15)
    default: break;
  }
  w.x += __shfl_xor_sync(DQA_FULL, w.x, 4);
  w.y += __shfl_xor_sync(DQA_FULL, w.y, 4);
  w.x += __shfl_xor_sync(DQA_FULL, w.x, 2);
  w.y += __shfl_xor_sync(DQA_FULL, w.y, 2);
  w.x += __shfl_xor_sync(DQA_FULL, w.x, 1);
  w.y += __shfl_xor_sync(DQA_FULL, w.y, 1);
  if (act) {
    const cplx tw = cmul(tk, w);
    switch (t0) {  // fall through on purpose
      SQR_UPD(0) SQR_UPD(1) SQR_UPD(2) SQR_UPD(3) SQR_UPD(4) SQR_UPD(5)
      SQR_UPD(6) SQR_UPD(7) SQR_UPD(8) SQR_UPD(9) SQR_UPD(10) SQR_UPD(11)
      SQR_UPD(12) SQR_UPD(13) SQR_UPD(14) SQR_UPD(15)
      default: break;
    }
  }
}

// Measured and rejected (round 2): reflectors in UNSCALED form (u = [alpha - beta; x], H^H = I - t u u^H, t = -1/(beta d)) with
// every consumer group deriving (beta, d, t) for itself from the published (alpha, |x|^2), so that neither the square root nor
// the reciprocal nor a scaling pass sits on the owner's path between two block barriers: 3-10 % SLOWER than the scaled form
// above at every bond dimension (chi=32: 310 vs 300 ms per step; 40.3k vs 37.9k cycles in the factor loop).  At three CTAs per
// SM the ~45 extra FP64-path instructions per warp and reflector cost more than the shorter hand-off saves.
template <int MAXS>
__global__ void __launch_bounds__(MAXS <= 8 ? 256 : (MAXS == 10 ? 320 : (MAXS == 12 ? 384 : 512)), MAXS <= 4 ? 4 : (MAXS <= 8 ? SQR_MINB : (MAXS <= SQR_TWO_UPTO ? 2 : 1))) k_sector_qr(DqaDev d, int n, int mu) {
  DQA_DYN_SMEM(smraw);
  const int y = blockIdx.x, inst = blockIdx.y;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const int j = tid >> 3, li = tid & 7;  // column owned by this 8-lane group, lane in group
  const int chi = d.chi, ldm = 8 * MAXS, ldq = 2 * chi;
  cplx* M = (cplx*)smraw;
  cplx* tau = M + (size_t)ldm * chi;
  double* betas = (double*)(tau + chi);

  const int* bond = d.bond + inst * (d.N + 1);
  const int bl = bond[n], br = bond[n + 1];
  const int hb = d.hbit[((size_t)inst * d.n_xi + mu) * d.N + n];
  const int ymax_in = d.N - n - 1;  // sectors of bond n+1 are 0..N-n-1
  const int* qd_in = d.qdim + ((size_t)inst * (d.N + 1) + (n + 1)) * d.smax;
  const int yp0 = y - hb, yp1 = y - (hb ^ 1);
  const int q0 = (yp0 >= 0 && yp0 <= ymax_in) ? qd_in[yp0] : 0;
  const int q1 = (yp1 >= 0 && yp1 <= ymax_in) ? qd_in[yp1] : 0;
  const int m = q0 + q1;
  const int q = m < bl ? m : bl;
  const size_t rsz = (size_t)chi * chi;
  const cplx* rin = d.rbuf + ((size_t)inst * 2 + ((n + 1) & 1)) * d.smax * rsz;
  cplx* rout = d.rbuf + (((size_t)inst * 2 + (n & 1)) * d.smax + y) * rsz;
  const cplx* a_site = d.mps + ((size_t)inst * d.N + n) * (2 * rsz);

  const long long tq0 = clock64();
  // ---- M = [R_{yp0} A_0^T ; R_{yp1} A_1^T] with FP64 tensor-core tiles ----
  {
    const int g = lane >> 2, t = lane & 3;
    const int tr0 = (q0 + 7) >> 3, tr1 = (q1 + 7) >> 3, tc = (bl + 7) >> 3;
    for (int tile = warp; tile < (tr0 + tr1) * tc; tile += nwarp) {
      const int rt = tile / tc, a0 = (tile % tc) * 8;
      const int s = rt < tr0 ? 0 : 1;
      const int r0 = (s ? rt - tr0 : rt) * 8, qs = s ? q1 : q0, rowoff = s ? q0 : 0;
      const cplx* R = rin + (size_t)(s ? yp1 : yp0) * rsz;
      ZTile acc = ztile_zero();
      warp_zgemm_8x8(
          acc, br,
          [&](int row, int kk) { return (r0 + row < qs && kk < br) ? R[(r0 + row) * chi + kk] : cmk(0.0, 0.0); },
          [&](int kk, int col) { return (kk < br && a0 + col < bl) ? a_site[((a0 + col) * 2 + s) * chi + kk] : cmk(0.0, 0.0); });
      const int row = r0 + g, a = a0 + 2 * t;
      if (row < qs) {
        if (a < bl) M[rowoff + row + a * ldm] = cmk(acc.r0, acc.i0);
        if (a + 1 < bl) M[rowoff + row + (a + 1) * ldm] = cmk(acc.r1, acc.i1);
      }
    }
  }
  __syncthreads();

  // ---- Householder QR with register-resident columns ----
  const bool has = j < bl;
  cplx col[MAXS];
#pragma unroll
  for (int t = 0; t < MAXS; ++t) {
    const int row = li + 8 * t;
    col[t] = (has && row < m) ? M[row + j * ldm] : cmk(0.0, 0.0);
  }
  __syncthreads();  // M is reused for the reflectors from here on
  const long long tq1 = clock64();
  if (warp == 0 && q > 0) sqr_publish<MAXS>(col, 0, m, j, li, M, tau, betas);
  __syncthreads();
  for (int k = 0; k < q; ++k) {
    if (4 * warp + 3 > k) {  // warps whose four columns are all finished (j <= k) have nothing to do
      const cplx tk = cconj(tau[k]);  // H^H is applied
      const bool act = has && j > k && (tk.x != 0.0 || tk.y != 0.0);
      sqr_apply<MAXS>(col, M + k * ldm, tk, k >> 3, li, act);
      if (k + 1 < q && warp == ((k + 1) >> 2)) sqr_publish<MAXS>(col, k + 1, m, j, li, M + (k + 1) * ldm, tau, betas);
    }
    __syncthreads();
  }

  const long long tq2 = clock64();
  // ---- R out (row-major q x bl, ld chi), diagonal made non-negative ----
  if (has) {
#pragma unroll
    for (int t = 0; t < MAXS; ++t) {
      const int row = li + 8 * t;
      if (row < q) {
        cplx v = cmk(0.0, 0.0);
        if (row <= j) {
          v = col[t];
          if (betas[row] < 0.0) v = cmk(-v.x, -v.y);
        }
        rout[row * chi + j] = v;
      }
    }
  }

  // ---- Q = H_0 ... H_{q-1} [I; 0]: column j needs reflectors j, j-1, ..., 0; no block barrier ----
  {
    cplx qc[MAXS];
#pragma unroll
    for (int t = 0; t < MAXS; ++t) qc[t] = (li + 8 * t == j && j < q) ? cmk(1.0, 0.0) : cmk(0.0, 0.0);
    int imax = 4 * warp + 3;
    if (imax > q - 1) imax = q - 1;
    for (int i = imax; i >= 0; --i) {  // warp-uniform trip count
      const cplx ti = tau[i];
      const bool act = (j < q) && (i <= j) && (ti.x != 0.0 || ti.y != 0.0);
      sqr_apply<MAXS>(qc, M + i * ldm, ti, i >> 3, li, act);
    }
    if (j < q) {
      cplx* qout = d.qbuf + ((size_t)inst * d.nsec_total + dqa_secbase(d.N, n) + y) * (2 * rsz);
      const double sg = betas[j] < 0.0 ? -1.0 : 1.0;
#pragma unroll
      for (int t = 0; t < MAXS; ++t) {
        const int row = li + 8 * t;
        if (row < m) qout[row + j * ldq] = cscale(qc[t], sg);
      }
    }
  }
  if (tid == 0) {
    d.qdim[((size_t)inst * (d.N + 1) + n) * d.smax + y] = q;
    atomicAdd(&d.counters[inst * DQA_NCNT + 3], (unsigned long long)q);
    unsigned long long fl = (unsigned long long)m * bl * br * 8ull;  // M = R A^T
    for (int k = 0; k < q; ++k) {
      fl += 10ull * (m - k - 1) + 16ull * (unsigned long long)(bl - k - 1) * (m - k);   // zgeqr2 column k
      fl += 6ull * (m - k - 1) + 16ull * (unsigned long long)(q - k - 1) * (m - k - 1);  // zung2r column k
    }
    atomicAdd(&d.counters[inst * DQA_NCNT + 6], fl);
    const long long tq3 = clock64();
    atomicAdd(&d.counters[inst * DQA_NCNT + 8], (unsigned long long)(tq1 - tq0));
    atomicAdd(&d.counters[inst * DQA_NCNT + 9], (unsigned long long)(tq2 - tq1));
    atomicAdd(&d.counters[inst * DQA_NCNT + 10], (unsigned long long)(tq3 - tq2));
    atomicAdd(&d.counters[inst * DQA_NCNT + 11], 1ull);
  }
}

// bond N has one sector (y=0) of dimension 1 with R = [1]
__global__ void k_sector_init(DqaDev d) {
  const int inst = blockIdx.x;
  if (threadIdx.x == 0) {
    d.qdim[((size_t)inst * (d.N + 1) + d.N) * d.smax + 0] = 1;
    d.rbuf[(((size_t)inst * 2 + (d.N & 1)) * d.smax + 0) * d.chi * d.chi] = cmk(1.0, 0.0);
  }
}

__device__ __forceinline__ int qdim_prefix(const int* qd, int y) {
  int o = 0;
  for (int i = 0; i < y; ++i) o += qd[i];
  return o;
}

// ---------------------------------------------------------------------------------
// k_theta0: Theta_0[(s), (j,y')] = u_p(y' + nbit_0(s)) * sum_c R_{1,y'}[j,c] A_0[0,s,c]
// grid (N sectors of bond 1, batch)
// ---------------------------------------------------------------------------------
__global__ void k_theta0(DqaDev d, int mu, int p_arg) {
  const int yp = blockIdx.x, inst = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  const int p = p_arg >= 0 ? p_arg : *d.pdev;
  const int chi = d.chi;
  const int* qd = d.qdim + ((size_t)inst * (d.N + 1) + 1) * d.smax;
  const int qn = qd[yp], off = qdim_prefix(qd, yp);
  const int br = d.bond[inst * (d.N + 1) + 1];
  const int hb = d.hbit[((size_t)inst * d.n_xi + mu) * d.N + 0];
  const size_t rsz = (size_t)chi * chi;
  const cplx* r1 = d.rbuf + (((size_t)inst * 2 + 1) * d.smax + yp) * rsz;
  const cplx* a0 = d.mps + ((size_t)inst * d.N) * (2 * rsz);
  cplx* th = d.theta + ((size_t)inst * 2 + 0) * (size_t)(2 * chi) * d.ld_theta;
  for (int i = tid; i < 2 * qn; i += nt) {
    const int j = i % qn, s = i / qn;
    cplx acc = cmk(0.0, 0.0);
    for (int c = 0; c < br; ++c) acc = cfma(r1[j * chi + c], a0[s * chi + c], acc);
    const cplx u = d.uz[((size_t)d.tab[inst] * d.P + p) * d.D + yp + (hb ^ s)];
    th[(size_t)s * d.ld_theta + off + j] = cmul(u, acc);
  }
}

// ---------------------------------------------------------------------------------
// k_theta (l >= 1): absorb the previous cut's S V^H and contract with the sector Q blocks.
//   C_y[beta,j]   = scale * sum_r conj(U[r,beta]) Theta_{l-1}[r, off_y + j]      (U = new site l-1)
//   Theta_l[(beta,s), off'_{y'} + j'] = sum_j C_y[beta,j] Q_{l,y}[(rowoff_s + j'), j],  y' = y - nbit_l(s)
// grid (N-l+1 sectors of bond l, batch); shared: C (chi x chi)
// ---------------------------------------------------------------------------------
// The two GEMMs of k_theta with NT row tiles per warp task sharing the B fragment (see warp_zgemm_8x8_rows).
template <int NT>
__device__ __forceinline__ void theta_gemm1(cplx* C, int ldc, const cplx* U, const cplx* thp, int ld, int offy, int chi, int cl, int qy,
                                            int mrp, double scale) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
  const int tb = (cl + 7) >> 3, tj = (qy + 7) >> 3, tbg = (tb + NT - 1) / NT;
  for (int task = warp; task < tj * tbg; task += nwarp) {
    const int j0 = (task % tj) * 8, bt0 = (task / tj) * NT;
    ZTile acc[NT];
#pragma unroll
    for (int q = 0; q < NT; ++q) acc[q] = ztile_zero();
    warp_zgemm_8x8_rows<NT>(
        acc, mrp,
        [&](int q, int row, int kk) {
          const int b = (bt0 + q) * 8 + row;
          return (b < cl && kk < mrp) ? cconj(U[kk * chi + b]) : cmk(0.0, 0.0);
        },
        [&](int kk, int col) { return (kk < mrp && j0 + col < qy) ? thp[(size_t)kk * ld + offy + j0 + col] : cmk(0.0, 0.0); });
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      const int b = (bt0 + q) * 8 + g, j = j0 + 2 * t;
      if (b < cl) {
        if (j < qy) C[b * ldc + j] = cmk(acc[q].r0 * scale, acc[q].i0 * scale);
        if (j + 1 < qy) C[b * ldc + j + 1] = cmk(acc[q].r1 * scale, acc[q].i1 * scale);
      }
    }
  }
}
template <int NT>
__device__ __forceinline__ void theta_gemm2(cplx* thc, int ld, int offo, int s, const cplx* C, int ldc, const cplx* Qs, int ldq, int cl,
                                            int qy, int qo) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
  const int tb = (cl + 7) >> 3, to = (qo + 7) >> 3, tbg = (tb + NT - 1) / NT;
  for (int task = warp; task < to * tbg; task += nwarp) {
    const int p0 = (task % to) * 8, bt0 = (task / to) * NT;
    ZTile acc[NT];
#pragma unroll
    for (int q = 0; q < NT; ++q) acc[q] = ztile_zero();
    warp_zgemm_8x8_rows<NT>(
        acc, qy,
        [&](int q, int row, int kk) {
          const int b = (bt0 + q) * 8 + row;
          return (b < cl && kk < qy) ? C[b * ldc + kk] : cmk(0.0, 0.0);
        },
        [&](int kk, int col) { return (kk < qy && p0 + col < qo) ? Qs[(p0 + col) + kk * ldq] : cmk(0.0, 0.0); });
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      const int b = (bt0 + q) * 8 + g, jp = p0 + 2 * t;
      if (b < cl) {
        cplx* dst = thc + (size_t)(b * 2 + s) * ld + offo;
        if (jp < qo) dst[jp] = cmk(acc[q].r0, acc[q].i0);
        if (jp + 1 < qo) dst[jp + 1] = cmk(acc[q].r1, acc[q].i1);
      }
    }
  }
}

// Measured and rejected (round 2): 2 x 2 register blocking of the two GEMMs (one task = 2 row tiles x 2 column tiles, NT + MT
// fragment loads per 16 DMMAs instead of 3 per 8, aimed at the 82 % busy L1 data pipe).  98 registers instead of 56 cut the
// resident CTAs per SM from 9 to 5 and the kernel got SLOWER: 92.6 vs 71.6 ms per step at chi=32, 107.6 vs 76.4 at chi=48
// (batch 888 / 296).  The loads from global memory are hidden by warps in flight, not by fewer loads.
// Measured and rejected (round 2; branch exp/theta-bulk-staging): k_theta_bulk -- all three operands staged in shared memory by
// the TMA engine (cp.async.bulk: one copy for the site tensor, one for the Q block, one per row of the Theta block; two
// mbarriers counting bytes; UBLKCP / SYNCS in the SASS, 58 registers), GEMM-1 accumulators held in registers across a barrier so
// that C takes the Theta block's place, every DMMA fragment read from shared memory with conflict-free strides.  Parity-green
// on the GPU, and SLOWER: 113 ms per step with 128-thread CTAs, 103 ms with 256, against 73 ms at chi=32 (45 vs 38 at chi=20,
// equal at chi=10; batch 888).  96 KB of operands leave two CTAs per SM instead of nine; nine CTAs' worth of warps hide the
// L2 latency of the direct fragment loads better than two CTAs hide their staging phase.
// row stride of C in shared memory: = 4 (mod 8) complex, so the two rows a quarter-warp reads in one DMMA fragment load
// (4 consecutive k each) do not share banks
__host__ __device__ inline int theta_ldc(int chi) { return ((chi + 7) & ~7) + 4; }
__host__ __device__ inline size_t theta_smem(int chi) { return sizeof(cplx) * (size_t)chi * theta_ldc(chi); }
__global__ void __launch_bounds__(256) k_theta(DqaDev d, int l, int mu) {
  DQA_DYN_SMEM(smraw);
  cplx* C = (cplx*)smraw;
  const int y = blockIdx.x, inst = blockIdx.y, tid = threadIdx.x, nwarp = blockDim.x >> 5;
  const int chi = d.chi, ldq = 2 * chi, ld = d.ld_theta;
  const size_t rsz = (size_t)chi * chi;
  const int* bond = d.bond + inst * (d.N + 1);
  const int cl = bond[l], mrp = 2 * bond[l - 1];
  const int* qd = d.qdim + ((size_t)inst * (d.N + 1) + l) * d.smax;
  const int* qd1 = qd + d.smax;
  const int qy = qd[y], offy = qdim_prefix(qd, y);
  const int hb = d.hbit[((size_t)inst * d.n_xi + mu) * d.N + l];
  const int ymax_out = d.N - l - 1;
  const cplx* thp = d.theta + ((size_t)inst * 2 + ((l - 1) & 1)) * (size_t)(2 * chi) * ld;
  cplx* thc = d.theta + ((size_t)inst * 2 + (l & 1)) * (size_t)(2 * chi) * ld;
  const cplx* U = d.mps + ((size_t)inst * d.N + (l - 1)) * (2 * rsz);
  const double scale = d.scale[inst];
  const int tb = (cl + 7) >> 3, tj = (qy + 7) >> 3;
  const int ldc = theta_ldc(chi);
  // GEMM 1 (DMMA): C[b,j] = scale * sum_r conj(U[r,b]) Theta_{l-1}[r, offy+j]; two row tiles share the Theta fragment
  // when that still leaves a task for every warp
  if (tj * ((tb + 1) >> 1) >= nwarp) theta_gemm1<2>(C, ldc, U, thp, ld, offy, chi, cl, qy, mrp, scale);
  else theta_gemm1<1>(C, ldc, U, thp, ld, offy, chi, cl, qy, mrp, scale);
  __syncthreads();
  const cplx* Qm = d.qbuf + ((size_t)inst * d.nsec_total + dqa_secbase(d.N, l) + y) * (2 * rsz);
  const int yp0 = y - hb, yp1 = y - (hb ^ 1);
  const int v0 = (yp0 >= 0 && yp0 <= ymax_out), v1 = (yp1 >= 0 && yp1 <= ymax_out);
  const int q0 = v0 ? qd1[yp0] : 0;
  // GEMM 2 (DMMA): Theta_l[(b,s), off' + j'] = sum_j C[b,j] Q[(rowoff_s + j'), j]
  for (int s = 0; s < 2; ++s) {
    const int yp = s ? yp1 : yp0;
    if (!(s ? v1 : v0)) continue;
    const int qo = qd1[yp], offo = qdim_prefix(qd1, yp), rowoff = s ? q0 : 0;
    const int to = (qo + 7) >> 3;
    if (to * ((tb + 1) >> 1) >= nwarp) theta_gemm2<2>(thc, ld, offo, s, C, ldc, Qm + rowoff, ldq, cl, qy, qo);
    else theta_gemm2<1>(thc, ld, offo, s, C, ldc, Qm + rowoff, ldq, cl, qy, qo);
    if (tid == 0) atomicAdd(&d.counters[inst * DQA_NCNT + 7], 8ull * cl * qo * qy);
  }
  if (tid == 0) atomicAdd(&d.counters[inst * DQA_NCNT + 7], 8ull * cl * qy * mrp);
}

// ---------------------------------------------------------------------------------
// k_lq: Householder LQ of Theta_l (mr x Q, row-major in global/L2) for one instance:
//   Theta = L Qh, L (mr x k) lower-trapezoidal, k = min(mr, Q); only L is needed.
// Blocked (compact WY) algorithm: a panel of LQ_PB rows is factored in shared memory, then
// every trailing row is updated once per panel, Y <- Y - ((Y V) T) V^H, one warp per row,
// lanes over columns; 2*PB complex FMAs per element and panel.
// shared: Vp (PB x npad) | Tm (PB x PB) | Sg (PB x PB) | taus (PB)
// ---------------------------------------------------------------------------------
#define LQ_PB 8
#define LQ_MAXW 16  // partial-W tiles kept in shared memory (>= row tiles, <= warps per CTA)
__host__ __device__ inline int lq_npad(int ncols) { return ((ncols + 7) & ~7) + 4; }
// qmax = largest number of Theta columns any cut can have (host-computed bound), chi sets the row tiles
__host__ __device__ inline size_t lq_smem(int qmax, int chi, int pb) {
  return sizeof(cplx) * ((size_t)pb * lq_npad(qmax) + 2 * LQ_PB * LQ_PB + LQ_PB + (size_t)LQ_MAXW * 64 + (size_t)((2 * chi + 7) / 8) * 64);
}
// panel height: 8 rows while such a panel (pb x qmax) leaves room for two CTAs per SM; wider Theta (N=64, chi=64: 3 776
// columns) takes the tallest panel that still fits one CTA (3 rows there; the power-of-two search of round 1 gave 2)
__host__ __device__ inline int lq_pick_pb(int qmax, int chi) {
  if (lq_smem(qmax, chi, LQ_PB) <= 180 * 1024) return LQ_PB;
  int pb = LQ_PB;
  while (pb > 1 && lq_smem(qmax, chi, pb) > 224 * 1024) pb -= 1;
  return pb;
}

__global__ void __launch_bounds__(1024) k_lq(DqaDev d, int l) {
  DQA_DYN_SMEM(smraw);
  const int inst = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const int chi = d.chi, ld = d.ld_theta, npad = lq_npad(d.qmax);
  cplx* Vp = (cplx*)smraw;
  const int pbmax = d.lq_pb;
  cplx* Tm = Vp + (size_t)pbmax * npad;
  cplx* Sg = Tm + LQ_PB * LQ_PB;
  cplx* taus = Sg + LQ_PB * LQ_PB;
  cplx* Wp = taus + LQ_PB;            // LQ_MAXW partial 8x8 tiles
  cplx* Zs = Wp + (size_t)LQ_MAXW * 64;  // up to 16 row tiles (2chi <= 128 rows)
  const int* bond = d.bond + inst * (d.N + 1);
  const int mr = 2 * bond[l];
  const int* qd1 = d.qdim + ((size_t)inst * (d.N + 1) + (l + 1)) * d.smax;
  const int Q = qdim_prefix(qd1, d.N - l);
  const int k = mr < Q ? mr : Q;
  if (Q > d.qmax) {  // cannot happen for states accepted by dqa_set_mps; never overrun shared memory
    if (tid == 0) atomicOr(&d.status[inst], DQA_ST_LIMIT);
    return;
  }
  const cplx* th = d.theta + ((size_t)inst * 2 + (l & 1)) * (size_t)(2 * chi) * ld;
  cplx* wk = d.work + (size_t)inst * (size_t)(2 * chi) * ld;
  cplx* lout = d.lbuf + (size_t)inst * (size_t)(2 * chi) * (2 * chi);
  const int ldl = 2 * chi;

  long long tl_panel = 0, tl_t = 0, tl_trail = 0;
  for (int p0 = 0; p0 < k; p0 += pbmax) {
    const long long tl0 = clock64();
    const int pb = (k - p0) < pbmax ? (k - p0) : pbmax;
    const int n = Q - p0;  // panel columns p0..Q-1
    const cplx* src = (p0 == 0) ? th : wk;
    for (int c = tid; c < n; c += nt) {  // row by row: independent loads, no integer division
#pragma unroll
      for (int r = 0; r < LQ_PB; ++r)
        if (r < pb) Vp[r * npad + c] = src[(size_t)(p0 + r) * ld + p0 + c];
    }
    __syncthreads();
    tl_t += clock64() - tl0;  // panel load (reported as the "tmat" counter)
    // ---- factor the panel (rows p0..p0+pb-1), unblocked, in shared memory ----
    // (a column-distributed variant with one fused reduction per reflector was tried and measured slower
    //  under load: 203 vs 187 ms per step at batch 888)
    // One block barrier per reflector: while the other warps apply reflector i to their rows, warp 0 (which always
    // owns row i+1) accumulates the norm of its updated row inside the update loop, then generates and scales
    // reflector i+1 straight away.
    auto gen_row = [&](int i, double ss) {  // warp 0; ss = sum_{c>i} |row_i[c]|^2, already reduced
      cplx* row = Vp + i * npad;
      const cplx alpha = row[i];
      double beta;
      cplx t, scal;
      householder_gen_fast(cconj(alpha), ss, beta, t, scal);
      const cplx cs = cconj(scal);  // stored vector is conj(v)
      __syncwarp();
#pragma unroll 4
      for (int c = i + 1 + lane; c < n; c += 32) row[c] = cmul(row[c], cs);
      if (lane == 0) {
        taus[i] = t;
        row[i] = cmk(beta, 0.0);
      }
    };
    if (warp == 0) {
      double ss0 = 0.0, ss1 = 0.0;
      for (int c = 1 + lane; c < n; c += 64) {
        ss0 += cabs2(Vp[c]);
        if (c + 32 < n) ss1 += cabs2(Vp[c + 32]);
      }
      gen_row(0, warp_sum(ss0 + ss1));
    }
    __syncthreads();
    for (int i = 0; i < pb; ++i) {
      const cplx* __restrict__ row = Vp + i * npad;
      const cplx t = taus[i];
      const bool rot = (t.x != 0.0 || t.y != 0.0);
      double ssn = 0.0;
      for (int r = i + 1 + warp; r < pb; r += nwarp) {
        cplx* __restrict__ yr = Vp + r * npad;  // a different row than `row`: loads may be hoisted above the stores
        const bool next = (r == i + 1);  // warp 0's first row: the next reflector
        if (rot) {
          cplx w0 = (lane == 0) ? yr[i] : cmk(0.0, 0.0), w1 = cmk(0.0, 0.0);
#pragma unroll 2
          for (int c = i + 1 + lane; c < n; c += 64) {
            w0 = cfmac(yr[c], row[c], w0);
            if (c + 32 < n) w1 = cfmac(yr[c + 32], row[c + 32], w1);
          }
          const cplx w = warp_sum(cadd(w0, w1));
          const cplx tw = cmul(t, w);
          if (lane == 0) yr[i] = csub(yr[i], tw);
          if (next) {
#pragma unroll 4
            for (int c = i + 1 + lane; c < n; c += 32) {
              const cplx yv = cfms(tw, row[c], yr[c]);
              yr[c] = yv;
              if (c > i + 1) ssn += cabs2(yv);
            }
          } else {
#pragma unroll 4
            for (int c = i + 1 + lane; c < n; c += 32) yr[c] = cfms(tw, row[c], yr[c]);
          }
        } else if (next) {
          for (int c = i + 2 + lane; c < n; c += 32) ssn += cabs2(yr[c]);
        }
      }
      if (warp == 0 && i + 1 < pb) {
        __syncwarp();
        gen_row(i + 1, warp_sum(ssn));
      }
      __syncthreads();
    }
    // ---- L entries of the panel rows go back to the work buffer; Vp becomes the clean V ----
    for (int i = tid; i < pb * pb; i += nt) {
      const int c = i % pb, r = i / pb;
      if (c <= r) {
        wk[(size_t)(p0 + r) * ld + p0 + c] = Vp[r * npad + c];
      }
    }
    __syncthreads();
    for (int i = tid; i < pb * pb; i += nt) {
      const int c = i % pb, r = i / pb;
      if (c < r) Vp[r * npad + c] = cmk(0.0, 0.0);
      if (c == r) Vp[r * npad + c] = cmk(1.0, 0.0);
    }
    __syncthreads();
    const int r_first = p0 + pb;
    const long long tl1 = clock64();
    tl_panel += tl1 - tl0;
    if (r_first < mr) {
      const long long tl2 = clock64();
      // ---- trailing rows with FP64 tensor-core tiles --------------------------------------------------
      // pass 1: W = Y V for every 8-row tile of trailing rows, plus one extra tile with Y = the panel
      //         itself, which yields S = V^H V (the Gram matrix of the reflectors); K split over warps.
      // Z = W T without ever forming T: T^{-1} = diag(1/tau) + strict_upper(S), so every row solves
      //         z_i = tau_i (w_i - sum_{j<i} z_j S[j][i]) independently (forward substitution).
      // pass 2: Y -= Z V^H.
      const int g = lane >> 2, t4 = lane & 3;
      const int nrows = mr - r_first, ntr = (nrows + 7) >> 3, ntile1 = ntr + 1;
      int ks = nwarp / ntile1;  // K-splits per tile
      if (ks < 1) ks = 1;
      if (ks > LQ_MAXW / ntile1) ks = LQ_MAXW / ntile1 > 0 ? LQ_MAXW / ntile1 : 1;
      const int kchunk = (((n + ks - 1) / ks) + 3) & ~3;
      for (int task = warp; task < ntile1 * ks; task += nwarp) {
        const int rt = task / ks, kpart = task % ks;
        const int kbeg = kpart * kchunk;
        int kend = kbeg + kchunk;
        if (kend > n) kend = n;
        ZTile acc = ztile_zero();
        if (kbeg < kend) {
          const cplx* vsrc = Vp + kbeg;
          const int klen = kend - kbeg;
          if (rt < ntr) {
            const int r0 = r_first + rt * 8;
            const cplx* ysrc = src + p0 + kbeg;
            warp_zgemm_8x8(
                acc, klen,
                [&](int row, int kk) { return (r0 + row < mr && kk < klen) ? ysrc[(size_t)(r0 + row) * ld + kk] : cmk(0.0, 0.0); },
                [&](int kk, int col) { return (col < pb && kk < klen) ? cconj(vsrc[col * npad + kk]) : cmk(0.0, 0.0); });
          } else {  // the panel against itself: acc[j][i] = sum_c vbar_j[c] conj(vbar_i[c]) = v_j^H v_i = S[j][i]
            warp_zgemm_8x8(
                acc, klen,
                [&](int row, int kk) { return (row < pb && kk < klen) ? vsrc[row * npad + kk] : cmk(0.0, 0.0); },
                [&](int kk, int col) { return (col < pb && kk < klen) ? cconj(vsrc[col * npad + kk]) : cmk(0.0, 0.0); });
          }
        }
        cplx* wp = Wp + (size_t)task * 64;
        wp[g * 8 + 2 * t4] = cmk(acc.r0, acc.i0);
        wp[g * 8 + 2 * t4 + 1] = cmk(acc.r1, acc.i1);
      }
      __syncthreads();
      for (int o = tid; o < ntr * 8; o += nt) {  // one thread per trailing row
        const int rt = o >> 3, row = o & 7;
        cplx z[LQ_PB];
#pragma unroll
        for (int i = 0; i < LQ_PB; ++i) {
          cplx acc = cmk(0.0, 0.0);
          if (i < pb) {
            for (int kp2 = 0; kp2 < ks; ++kp2) acc = cadd(acc, Wp[(size_t)(rt * ks + kp2) * 64 + row * 8 + i]);
#pragma unroll
            for (int jj = 0; jj < LQ_PB; ++jj) {
              if (jj < i) {
                cplx sji = cmk(0.0, 0.0);
                for (int kp2 = 0; kp2 < ks; ++kp2) sji = cadd(sji, Wp[(size_t)(ntr * ks + kp2) * 64 + jj * 8 + i]);
                acc = cfms(z[jj], sji, acc);
              }
            }
            acc = cmul(taus[i], acc);
          }
          z[i] = acc;
          Zs[o * 8 + i] = acc;
        }
      }
      __syncthreads();
      const int ntc = (n + 7) >> 3;
      // the Y entries of the NEXT tile are requested before the current tile's small GEMM, so their L2/HBM latency
      // overlaps it (one tile of look-ahead is what the 64-register budget allows)
      cplx yn0 = cmk(0.0, 0.0), yn1 = cmk(0.0, 0.0);
      {
        const int tile = warp;
        if (tile < ntr * ntc) {
          const int r = r_first + (tile / ntc) * 8 + g, c = (tile % ntc) * 8 + 2 * t4;
          if (r < mr) {
            const cplx* ys = src + (size_t)r * ld + p0;
            if (c < n) yn0 = ys[c];
            if (c + 1 < n) yn1 = ys[c + 1];
          }
        }
      }
      for (int tile = warp; tile < ntr * ntc; tile += nwarp) {
        const int rt = tile / ntc, c0 = (tile % ntc) * 8;
        const int r0 = r_first + rt * 8;
        const cplx y0 = yn0, y1 = yn1;
        {
          const int tn = tile + nwarp;
          if (tn < ntr * ntc) {
            const int rn = r_first + (tn / ntc) * 8 + g, cn = (tn % ntc) * 8 + 2 * t4;
            if (rn < mr) {
              const cplx* ys = src + (size_t)rn * ld + p0;
              if (cn < n) yn0 = ys[cn];
              if (cn + 1 < n) yn1 = ys[cn + 1];
            }
          }
        }
        ZTile acc = ztile_zero();
        const cplx* zt = Zs + rt * 64;
        warp_zgemm_8x8(
            acc, 8, [&](int row, int kk) { return zt[row * 8 + kk]; },
            [&](int kk, int col) { return (kk < pb && c0 + col < n) ? Vp[kk * npad + c0 + col] : cmk(0.0, 0.0); });
        const int r = r0 + g, c = c0 + 2 * t4;
        if (r < mr) {
          cplx* yd = wk + (size_t)r * ld + p0;
          if (c < n) yd[c] = cmk(y0.x - acc.r0, y0.y - acc.i0);
          if (c + 1 < n) yd[c + 1] = cmk(y1.x - acc.r1, y1.y - acc.i1);
        }
      }
      tl_trail += clock64() - tl2;
    }
    __syncthreads();
  }
  for (int i = tid; i < mr * k; i += nt) {
    const int r = i % mr, c = i / mr;
    lout[r + c * ldl] = (c <= r) ? wk[(size_t)r * ld + c] : cmk(0.0, 0.0);
  }
  if (tid == 0) {
    unsigned long long lq = 0;
    for (int kk = 0; kk < k; ++kk) lq += 10ull * (Q - kk - 1) + 16ull * (unsigned long long)(mr - kk - 1) * (Q - kk);
    atomicAdd(&d.counters[inst * DQA_NCNT + 4], lq);
    atomicAdd(&d.counters[inst * DQA_NCNT + 12], (unsigned long long)tl_panel);
    atomicAdd(&d.counters[inst * DQA_NCNT + 13], (unsigned long long)tl_t);
    atomicAdd(&d.counters[inst * DQA_NCNT + 14], (unsigned long long)tl_trail);
    atomicAdd(&d.counters[inst * DQA_NCNT + 15], 1ull);
  }
}

// Measured and rejected (round 2; branch exp/lq-panel-pivot): PANEL-LEVEL PIVOTING -- every panel takes the 8 rows of largest
// residual norm (norms
// computed once, downdated by the L entries each trailing update produces; rows addressed through a position -> row map, the
// Jacobi tail un-permuting the rows of the new site tensor), which grades the factor a la Drmac-Veselic without a second
// factorisation.  Parity-green on the GPU; the Jacobi needs 6.08 instead of 6.51 sweeps per SVD (5 908 vs 6 414 rotations;
// tools/jacobi_study.py PIVOT=1 predicts 7.8 vs 8.6 on the mid-chain factors) and takes 394 instead of 418 ms per step at
// chi=32 -- and the LQ takes 192 instead of 178 ms (one more pass over Theta for the norms, a rank sort per panel, indirect
// row addressing): 1.006 s per step against 1.005 s.  A wash; left out.
// Measured and rejected (round 2; branch exp/lq-row-teams): ROW TEAMS in the panel factorisation -- the warps form R = pb-1-i
// teams of nwarp / R warps,
// one team per remaining row, the row's columns split over the team's lanes, partial dots and partial norms meeting in shared
// memory behind a named barrier of the team (bar.sync id, 32 T), every lane of the next reflector's team deriving the
// scalars for itself -- so that all 16 warps work instead of at most 7 and the next reflector's warp no longer walks all n
// columns three times.  Parity-green on the GPU, and SLOWER at every bond: LQ 221 vs 178 ms per step at chi=32 (panel phase
// 223k vs 146k cycles per CTA), 87 vs 72 at chi=20, 41 vs 32 at chi=10, batch 888.  The extra warp reductions and barrier
// traffic of 16 active warps cost more than the shorter critical path saves: with a second CTA on the SM in its DMMA phase the
// panel phase is not waited for, it is paid for in issue slots.
// Measured and rejected (round 2): the panel of k_lq held in REGISTERS, one column per thread (544 threads), with one
// block-wide reduction per reflector (a 16-slot shuffle butterfly per warp for |x|^2 and the dots with all later panel
// rows, one shared-memory hand-off, one barrier) and every thread deriving the reflector scalars for itself.  It removes
// the 975 KB of shared-memory traffic per panel, and was SLOWER: 219 vs 178 ms per step at batch 888, chi=32 (panel phase
// 177k vs 146k cycles per CTA; alone on the GPU 124k vs 110k): the block-wide reduction of 15 values costs ~1 900 cycles
// per reflector, more than the four passes over the panel it replaces.
// ---------------------------------------------------------------------------------
// k_jacobi (bonds above 32; k_jacobi_sys below handles chi <= 32): SVD of L (mr x k) by one-sided Jacobi + the reference's trim rule
// (lib/tenn.py:317-354) + the new left-orthonormal site tensor (lib/tenn.py:229-252).
// Few warps per CTA (several CTAs share an SM): every warp owns ppw column pairs per round;
// the complex dot products are reduced with warp shuffles; the rotation parameters of the
// warp's pairs are computed lane-parallel (lane u <-> pair u) and broadcast with shuffles,
// so the scalar FP64 chain is paid once per warp and round, not once per pair.
// Squared column norms are cached (exact recomputation every sweep).
// shared: L (2chi x 2chi col-major, ld = mr) | nrm2/sig, ssorted (2chi doubles) | order | sh_i
// ---------------------------------------------------------------------------------
// L lives in shared memory while (2chi)^2 complex fit next to two more CTAs' worth of headroom; beyond
// that (chi > 56) the kernel iterates on the L2-resident copy in `lbuf` (same code, generic pointers).
__host__ __device__ inline bool svd_l_in_smem(int chi) { return sizeof(cplx) * (size_t)(2 * chi) * (2 * chi) <= 200 * 1024; }
// When L does not fit (chi > 56) its first svd_hyb_cols(chi) columns still live in shared memory and only the rest is
// iterated on in L2 (chi = 64: 108 of 128 columns, so 71 % of the column pairs never leave shared memory).
__host__ __device__ inline int svd_hyb_cols(int chi) {
  return svd_l_in_smem(chi) ? 0 : (int)(((size_t)216 * 1024) / (sizeof(cplx) * (size_t)(2 * chi))) & ~3;
}
__host__ __device__ inline size_t svd_smem(int chi) {
  return (svd_l_in_smem(chi) ? sizeof(cplx) * (size_t)(2 * chi) * (2 * chi) : sizeof(cplx) * (size_t)svd_hyb_cols(chi) * (2 * chi)) +
         sizeof(double) * (4 * chi + 8) + sizeof(int) * (2 * chi + 8);
}

__device__ __forceinline__ void rr_pair(int n, int round, int idx, int& p, int& q) {
  // circle method on n (even) players: round in [0,n-2], idx in [0,n/2)
  if (idx == 0) {
    p = n - 1;
    q = round;
  } else {
    p = round + idx;
    if (p >= n - 1) p -= n - 1;
    q = round - idx;
    if (q < 0) q += n - 1;
  }
  if (p > q) {
    const int t = p;
    p = q;
    q = t;
  }
}

// The trimming decision of _trim_and_renorm_svd_result (lib/tenn.py:317-354) with the plan's policy; the reference's call
// site (lib/tenn.py:238) fixes cutoff 1e-10, mode 4 (relative cumulative s^2), renorm 2, absorb right.  One
// thread of the (out-of-line) tail.
__device__ __forceinline__ void trim_policy(const double* ssorted, int k, int chi, double cutoff, int mode, int renorm, int converged,
                                         int* n_chi_out, double* sc_out, double* discw_out, int* st_out) {
  const bool pw2 = (mode == 3 || mode == 4 || mode < 3);
  double tot = 0.0, tot2 = 0.0;
  for (int i = 0; i < k; ++i) {
    tot2 += ssorted[i] * ssorted[i];
    tot += pw2 ? ssorted[i] * ssorted[i] : ssorted[i];
  }
  double csp = 0.0, ckeep = 0.0, ckeep2 = 0.0;
  int cnt = 0, n_chi;
  if (cutoff > 0.0 || renorm > 0) {
    if (mode == 1) {
      for (int i = 0; i < k; ++i) cnt += ssorted[i] > cutoff;
      n_chi = cnt;
    } else if (mode == 2) {
      for (int i = 0; i < k; ++i) cnt += ssorted[i] > cutoff * ssorted[0];
      n_chi = cnt;
    } else {
      const double thr = (1.0 - cutoff) * tot;
      for (int i = 0; i < k; ++i) {
        csp += pw2 ? ssorted[i] * ssorted[i] : ssorted[i];
        if (mode == 4 || mode == 6) cnt += csp < thr;
        else cnt += (tot - csp) > cutoff;
      }
      n_chi = cnt + 1;
    }
    if (n_chi < 1) n_chi = 1;
    if (n_chi > chi) n_chi = chi;
  } else {
    n_chi = chi;
  }
  if (n_chi > k) n_chi = k;
  for (int i = 0; i < n_chi; ++i) {
    ckeep += pw2 ? ssorted[i] * ssorted[i] : ssorted[i];
    ckeep2 += ssorted[i] * ssorted[i];
  }
  double sc = 1.0;
  int st = 0;
  if (!(tot2 == tot2) || tot2 > 1e300) st |= DQA_ST_NONFINITE;
  if (tot2 == 0.0) st |= DQA_ST_ZERONORM;
  if (!converged) st |= DQA_ST_NOCONV;
  if (n_chi < k && ckeep > 0.0 && renorm > 0 && mode >= 3) sc = pw2 ? sqrt(tot / ckeep) : tot / ckeep;
  *n_chi_out = n_chi;
  *sc_out = sc;
  *discw_out = tot2 > 0.0 ? (tot2 - ckeep2) / tot2 : 0.0;
  *st_out = st;
}

// Shared tail of the Jacobi kernels: singular values of the kc stored columns (k of them real, kc - k zero padding),
// ordering, the reference's trim rule (lib/tenn.py:317-354), counters, and the new left-orthonormal site tensor.
// what the tail needs of the plan, passed BY VALUE to an out-of-line function: taking the address of the kernel's DqaDev
// parameter would force a 288-byte stack copy of it in every kernel that calls the tail (measured +9 % on the chi=10 Jacobi),
// while inlining the tail perturbs the allocation of the round loops (+3.5 % on the chi=32 one)
struct JacTailArgs {
  cplx* mps;
  double* schmidt;
  int* nsv;
  double* scale;
  int* status;
  unsigned long long* counters;
  int* bdmon;
  double* discw;
  const int* pdev;
  double cutoff;
  int N, P, n_xi, chi, cutoff_mode, renorm;
};
__device__ __forceinline__ JacTailArgs jac_tail_args(const DqaDev& d) {
  JacTailArgs a;
  a.mps = d.mps; a.schmidt = d.schmidt; a.nsv = d.nsv; a.scale = d.scale; a.status = d.status; a.counters = d.counters;
  a.bdmon = d.bdmon; a.discw = d.discw; a.pdev = d.pdev; a.cutoff = d.cutoff;
  a.N = d.N; a.P = d.P; a.n_xi = d.n_xi; a.chi = d.chi; a.cutoff_mode = d.cutoff_mode; a.renorm = d.renorm;
  return a;
}
__device__ __noinline__ void jacobi_tail(const JacTailArgs d, int inst, int l, int mu, int p_step, const cplx* L, int ldl, int mr, int k,
                                            int kc, double* sig, double* ssorted, int* order, int* sh_i, int* bond, int converged,
                                            int sweeps, unsigned long long nrot_total) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const int chi = d.chi;
  const size_t rsz = (size_t)chi * chi;
  for (int j = warp; j < kc; j += nwarp) {
    double a = 0.0;
    for (int r = lane; r < mr; r += 32) a += cabs2(L[r + j * ldl]);
    a = warp_sum(a);
    if (lane == 0) sig[j] = sqrt(a);
  }
  for (int j = tid; j < kc; j += nt) order[j] = j;
  __syncthreads();
  for (int j = tid; j < kc; j += nt) {
    const double sj = sig[j];
    int rank = 0;
    for (int i = 0; i < kc; ++i) {
      const double si = sig[i];
      rank += (si > sj) || (si == sj && i < j);
    }
    if (sj == sj) {  // NaN keeps the identity slot
      order[rank] = j;
      ssorted[rank] = sj;
    } else {
      ssorted[j] = sj;
    }
  }
  __syncthreads();
  if (tid == 0) {
    int n_chi, st;
    double sc, dw;
    trim_policy(ssorted, k, chi, d.cutoff, d.cutoff_mode, d.renorm, converged, &n_chi, &sc, &dw, &st);
    d.discw[(size_t)inst * d.N + l] = dw;
    if (st) atomicOr(&d.status[inst], st);
    sh_i[1] = n_chi;
    d.scale[inst] = sc;
    bond[l + 1] = n_chi;
    d.nsv[inst * d.N + l] = k;
    atomicAdd(&d.counters[inst * DQA_NCNT + 0], nrot_total);
    atomicAdd(&d.counters[inst * DQA_NCNT + 1], (unsigned long long)sweeps);
    atomicAdd(&d.counters[inst * DQA_NCNT + 2], 1ull);
    atomicAdd(&d.counters[inst * DQA_NCNT + 5],
              (unsigned long long)sweeps * (unsigned long long)(k * (k - 1) / 2) * 8ull * mr + nrot_total * 20ull * mr +
                  (unsigned long long)sweeps * 4ull * mr * k);
    if (l == d.N - 1) d.bdmon[(size_t)inst * d.P * d.n_xi + (size_t)(p_step >= 0 ? p_step : *d.pdev) * d.n_xi + mu] = bond[d.N / 2];
  }
  __syncthreads();
  const int n_chi = sh_i[1];
  for (int i = tid; i < k; i += nt) d.schmidt[((size_t)inst * d.N + l) * (2 * chi) + i] = ssorted[i];
  cplx* site = d.mps + ((size_t)inst * d.N + l) * (2 * rsz);
  for (int i = tid; i < mr * n_chi; i += nt) {
    const int b = i % n_chi, r = i / n_chi;
    const double sv = ssorted[b];
    const double inv = sv > 0.0 ? 1.0 / sv : 0.0;
    site[r * chi + b] = cscale(L[r + order[b] * ldl], inv);
  }
}

template <bool LSM>  // LSM: L lives in shared memory (compile-time, so its accesses are LDS/STS)
__global__ void __launch_bounds__(512) k_jacobi(DqaDev d, int l, int mu, int p_step) {
  DQA_DYN_SMEM(smraw);
  const int inst = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const int chi = d.chi;
    constexpr bool lsm = LSM;
  cplx* lglob = d.lbuf + (size_t)inst * (size_t)(2 * chi) * (2 * chi);
  cplx* L;
  if (LSM) L = (cplx*)smraw; else L = lglob;
  const int cs = LSM ? 0 : svd_hyb_cols(chi);  // hybrid (not LSM): columns < cs are iterated on in shared memory, ld 2chi
  cplx* Ls = (cplx*)smraw;
  double* sig = (double*)(smraw + (lsm ? sizeof(cplx) * (size_t)(2 * chi) * (2 * chi) : sizeof(cplx) * (size_t)cs * (2 * chi)));
  double* ssorted = sig + 2 * chi;
  int* order = (int*)(ssorted + 2 * chi + 8);
  int* sh_i = order + 2 * chi;  // [0]=rotation count, [1]=n_chi, [2]=some rotated pair was above the quadratic threshold
  int* bond = d.bond + inst * (d.N + 1);
  const int mr = 2 * bond[l];
  const int* qd1 = d.qdim + ((size_t)inst * (d.N + 1) + (l + 1)) * d.smax;
  const int Q = qdim_prefix(qd1, d.N - l);
  const int k = mr < Q ? mr : Q;
  const int ldl = lsm ? mr : 2 * chi;
  if (lsm) {
    for (int i = tid; i < mr * k; i += nt) {
      const int r = i % mr, c = i / mr;
      L[r + c * ldl] = lglob[r + c * (2 * chi)];
    }
  } else {
    const int kc = k < cs ? k : cs;
    for (int i = tid; i < mr * kc; i += nt) {
      const int r = i % mr, c = i / mr;
      Ls[r + c * ldl] = lglob[r + c * ldl];
    }
  }
  __syncthreads();
  auto colp = [&](int j) -> cplx* { return (LSM ? L : (j < cs ? Ls : lglob)) + (size_t)j * ldl; };

  // Work decomposition: a warp is 4 groups of 8 lanes; one group owns one column pair per
  // pass (8 lanes x mr/8 rows each), so a warp orthogonalises 4 pairs per pass with 3-stage
  // shuffle reductions that all four groups share.
  const int kp = k + (k & 1), half = kp >> 1;
  const int ppw = (half + nwarp - 1) / nwarp;  // pairs per warp and round
  const int grp = lane >> 3, li = lane & 7;
  const int npass = (ppw + 3) >> 2;  // passes of 4 pairs per warp and round (<= 8)
  const double tol = sqrt((double)mr) * 2.220446049250313e-16, tol2 = tol * tol;
  double* nrm2 = sig;
  int sweeps = 0, converged = (k <= 1);
  unsigned long long nrot_total = 0;
  for (int sweep = 0; sweep < d.max_sweeps && !converged; ++sweep) {
    for (int j0 = warp * 4; j0 < k; j0 += nwarp * 4) {  // warp-uniform trip count
      const int j = j0 + grp;
      double a = 0.0;
      if (j < k)
        for (int r = li; r < mr; r += 8) a += cabs2(colp(j)[r]);
      a += __shfl_xor_sync(DQA_FULL, a, 4);
      a += __shfl_xor_sync(DQA_FULL, a, 2);
      a += __shfl_xor_sync(DQA_FULL, a, 1);
      if (li == 0 && j < k) nrm2[j] = a;
    }
    if (tid == 0) {
      sh_i[0] = 0;
      sh_i[2] = 0;
    }
    __syncthreads();
    int myrot = 0, mybig = 0;
    for (int round = 0; round < kp - 1; ++round) {
      // ---- phase 1: conj(x_p).x_q for every pair of this warp; lane li==pass of a group keeps it ----
      cplx myg = cmk(0.0, 0.0);
      int myp = 0, myq = 0;
      bool myact = false;
      double mya = 0.0, myb = 0.0;
      const bool onepass = npass == 1;  // then every lane of a group keeps the pair's data: no broadcasts needed
      for (int pass = 0; pass < npass; ++pass) {
        const int u = pass * 4 + grp, pr = warp * ppw + u;
        int p = 0, q = 0;
        bool act = (u < ppw) && (pr < half);
        if (act) {
          rr_pair(kp, round, pr, p, q);
          act = q < k;  // padded dummy column
        }
        cplx g = cmk(0.0, 0.0);
        double a0 = 0.0, b0 = 0.0;
        if (act) {  // cached squared norms, read before the warp-wide reduction (they are rewritten in phase 2)
          a0 = nrm2[p];
          b0 = nrm2[q];
        }
        if (act) {
          const cplx* xp = colp(p);
          const cplx* xq = colp(q);
          for (int r = li; r < mr; r += 8) g = cfmac(xq[r], xp[r], g);  // conj(xp) * xq
        }
        g.x += __shfl_xor_sync(DQA_FULL, g.x, 4);
        g.y += __shfl_xor_sync(DQA_FULL, g.y, 4);
        g.x += __shfl_xor_sync(DQA_FULL, g.x, 2);
        g.y += __shfl_xor_sync(DQA_FULL, g.y, 2);
        g.x += __shfl_xor_sync(DQA_FULL, g.x, 1);
        g.y += __shfl_xor_sync(DQA_FULL, g.y, 1);
        if (li == pass || onepass) {
          myg = g;
          myact = act;
          mya = a0;
          myb = b0;
        }
        if (li == pass || onepass) {
          myp = p;
          myq = q;
        }
      }
      // ---- phase 2: rotation parameters, one pair per lane (the FP64 chain is paid once per round) ----
      double rc = 1.0, rsn = 0.0;
      cplx rsph = cmk(0.0, 0.0), rcph = cmk(1.0, 0.0);
      int rot = 0;
      if (myact) {
        const double a = mya, b = myb;
        const double ag2 = cabs2(myg);
        if (ag2 > tol2 * a * b && ag2 > 0.0) {
          const double rinv = rsqrt(ag2);
          const double ag = ag2 * rinv;
          const double zeta = 0.5 * (b - a) * rinv;
          const double az = fabs(zeta);
          double tt;
          if (az < 1e100) {
            const double w = fma(zeta, zeta, 1.0);
            tt = 1.0 / (az + w * rsqrt(w));
          } else {
            tt = 0.5 / az;
          }
          tt = copysign(tt, zeta);
          rc = rsqrt(fma(tt, tt, 1.0));
          rsn = rc * tt;
          const cplx ph = cmk(myg.x * rinv, -myg.y * rinv);  // e^{-i phi}
          rsph = cscale(ph, rsn);
          rcph = cscale(ph, rc);
          rot = 1;
          if (ag2 > d.jac_quad2 * a * b) mybig = 1;  // an off-diagonal above the quadratic-convergence threshold
          if (!onepass || li == 0) {
            const double na = a - tt * ag, nb = b + tt * ag;
            nrm2[myp] = na > 0.0 ? na : 0.0;
            nrm2[myq] = nb > 0.0 ? nb : 0.0;
            myrot++;
          }
        }
      }
      // ---- phase 3: every group applies its pair's rotation, pass by pass ----
      for (int pass = 0; pass < npass; ++pass) {
        int dr = rot, p = myp, q = myq;
        double c = rc, sn = rsn;
        cplx sph = rsph, cph = rcph;
        if (!onepass) {
          const int src = (lane & ~7) | pass;
          dr = __shfl_sync(DQA_FULL, rot, src);
          if (!__any_sync(DQA_FULL, dr)) continue;  // warp-uniform
          p = __shfl_sync(DQA_FULL, myp, src);
          q = __shfl_sync(DQA_FULL, myq, src);
          c = __shfl_sync(DQA_FULL, rc, src);
          sn = __shfl_sync(DQA_FULL, rsn, src);
          sph = cmk(__shfl_sync(DQA_FULL, rsph.x, src), __shfl_sync(DQA_FULL, rsph.y, src));
          cph = cmk(__shfl_sync(DQA_FULL, rcph.x, src), __shfl_sync(DQA_FULL, rcph.y, src));
        }
        if (dr) {
          cplx* xp = colp(p);
          cplx* xq = colp(q);
          for (int r = li; r < mr; r += 8) {
            const cplx a_ = xp[r], b_ = xq[r];
            xp[r] = cfms(sph, b_, cscale(a_, c));   // c*xp - s e^{-i phi} xq
            xq[r] = cfma(cph, b_, cscale(a_, sn));  // s*xp + c e^{-i phi} xq
          }
        }
      }
      __syncthreads();
    }
    if (myrot) atomicAdd(&sh_i[0], myrot);
    if (mybig) atomicOr(&sh_i[2], 1);
    __syncthreads();
    sweeps++;
    nrot_total += (unsigned long long)sh_i[0];
    // Converged when a sweep rotated nothing, or when every pair it rotated was already orthogonal to
    // jac_quad = 1e-10 (relative): cyclic Jacobi converges quadratically, so after such a sweep all
    // off-diagonals are O(n * 1e-20 / relative gap) << sqrt(m)*eps and the checking sweep is skipped.
    converged = (sh_i[0] == 0) || (sh_i[2] == 0);
    __syncthreads();
  }

  if (!LSM && cs > 0) {  // the tail reads the factor from one place: put the shared-memory columns back
    const int kc = k < cs ? k : cs;
    for (int i = tid; i < mr * kc; i += nt) {
      const int r = i % mr, c = i / mr;
      lglob[r + c * ldl] = Ls[r + c * ldl];
    }
    __syncthreads();
  }
  jacobi_tail(jac_tail_args(d), inst, l, mu, p_step, L, ldl, mr, k, k, sig, ssorted, order, sh_i, bond, converged, sweeps, nrot_total);
}

// ---------------------------------------------------------------------------------
// k_jacobi_sys: the same SVD + trim as k_jacobi for 2chi <= 8*RS rows, as a systolic array.
// An 8-lane group owns two columns for the WHOLE iteration, in registers (RS rows per lane
// and column).  Pairing follows the odd-even transposition ordering: in even rounds a group
// orthogonalises its own two columns and passes one to the right-hand neighbour, in odd
// rounds it orthogonalises the received column against the one it kept and passes one back
// to the left; the columns at the two ends of the line idle in odd rounds (group 0 holds
// both and only exchanges their roles), which is what makes every pair meet exactly once in
// a sweep of 2*groups rounds.  Per pair and round one column crosses shared memory (one
// store + one load) instead of the three loads + one store per column of k_jacobi, which is
// shared-memory-bandwidth bound.  Cached squared norms travel with the columns.  Rotation
// parameters without a division:  delta=(b-a)/2, r=hypot(delta,|g|), u=r+|delta|,
// q=1/sqrt(2ru):  c=uq, s e^{-i phi}=sign(delta) q conj(g), t|g|=2|g|^2 r q^2; the second
// column of a pair is left with the phase e^{i phi} (a gauge of the singular vector).
// shared: exchange buffers 2 x (groups x mr), reused for the final columns | sig | xn/ssorted | order | sh_i
// ---------------------------------------------------------------------------------
#ifndef JSYS_MINB8
#define JSYS_MINB8 2  // CTAs per SM the 8-row-slot systolic Jacobi is compiled for
#endif
template <int RS>
__device__ __forceinline__ void jacobi_sys_body(const DqaDev& d, int l, int mu, int p_step) {
  DQA_DYN_SMEM(smraw);
  const int inst = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int chi = d.chi;
  const cplx* lglob = d.lbuf + (size_t)inst * (size_t)(2 * chi) * (2 * chi);
  cplx* X = (cplx*)smraw;
  double* sig = (double*)(smraw + sizeof(cplx) * (size_t)(2 * chi) * (2 * chi));
  double* ssorted = sig + 2 * chi;
  int* order = (int*)(ssorted + 2 * chi + 8);
  int* sh_i = order + 2 * chi;  // [0]=rotation count, [1]=n_chi, [2]=some rotated pair was above the quadratic threshold
  int* bond = d.bond + inst * (d.N + 1);
  const int mr = 2 * bond[l];
  const int* qd1 = d.qdim + ((size_t)inst * (d.N + 1) + (l + 1)) * d.smax;
  const int Q = qdim_prefix(qd1, d.N - l);
  const int k = mr < Q ? mr : Q;
  const int kp = k + (k & 1), h = kp >> 1;  // h groups hold the kp columns (one zero column when k is odd)
  const int g = tid >> 3, li = lane & 7;
  const bool act = g < h;
  double* xn = ssorted;  // [2][chi] squared norms in flight (ssorted is only used by the tail)

  // A = column 2g+1, B = column 2g, straight from the L2-resident factor
  cplx A[RS], B[RS];
#pragma unroll
  for (int t = 0; t < RS; ++t) {
    const int r = li + 8 * t;
    const bool in = act && r < mr;
    B[t] = (in && 2 * g < k) ? lglob[r + (size_t)(2 * g) * (2 * chi)] : cmk(0.0, 0.0);
    A[t] = (in && 2 * g + 1 < k) ? lglob[r + (size_t)(2 * g + 1) * (2 * chi)] : cmk(0.0, 0.0);
  }
  const int gr = (g + 1 < h) ? g + 1 : 0, gl = (g > 0) ? g - 1 : h - 1;  // ring neighbours (only the parked column wraps)
  const size_t bufsz = (size_t)h * mr;
  const double tol = sqrt((double)mr) * 2.220446049250313e-16, tol2 = tol * tol;
  int sweeps = 0, converged = (k <= 1);
  unsigned long long nrot_total = 0;
  for (int sweep = 0; sweep < d.max_sweeps && !converged; ++sweep) {
    double nA = 0.0, nB = 0.0;  // exact squared norms once per sweep, updated by the rotation formulas in between
#pragma unroll
    for (int t = 0; t < RS; ++t) {
      nA += cabs2(A[t]);
      nB += cabs2(B[t]);
    }
    nA += __shfl_xor_sync(DQA_FULL, nA, 4);
    nB += __shfl_xor_sync(DQA_FULL, nB, 4);
    nA += __shfl_xor_sync(DQA_FULL, nA, 2);
    nB += __shfl_xor_sync(DQA_FULL, nB, 2);
    nA += __shfl_xor_sync(DQA_FULL, nA, 1);
    nB += __shfl_xor_sync(DQA_FULL, nB, 1);
    if (tid == 0) {
      sh_i[0] = 0;
      sh_i[2] = 0;
    }
    int myrot = 0, mybig = 0;
    for (int round = 0; round < kp; ++round) {
      const bool straddle = round & 1;
      const bool idle = straddle && g == 0;  // both end columns of the line: exchange roles, no rotation
      // ---- conj(A).B with independent partial sums ----
      cplx g0 = cmk(0.0, 0.0), g1 = cmk(0.0, 0.0);
#pragma unroll
      for (int t = 0; t < RS; t += 2) {
        g0 = cfmac(B[t], A[t], g0);
        if (t + 1 < RS) g1 = cfmac(B[t + 1], A[t + 1], g1);
      }
      cplx gd = cadd(g0, g1);
      gd.x += __shfl_xor_sync(DQA_FULL, gd.x, 4);
      gd.y += __shfl_xor_sync(DQA_FULL, gd.y, 4);
      gd.x += __shfl_xor_sync(DQA_FULL, gd.x, 2);
      gd.y += __shfl_xor_sync(DQA_FULL, gd.y, 2);
      gd.x += __shfl_xor_sync(DQA_FULL, gd.x, 1);
      gd.y += __shfl_xor_sync(DQA_FULL, gd.y, 1);
      // ---- rotation parameters (every lane of the group, redundantly): A' = c A + z1 B, B' = c B + z2 A ----
      double c = 1.0;
      cplx z1 = cmk(0.0, 0.0), z2 = cmk(0.0, 0.0);
      int rot = 0;
      if (idle) {
        c = 0.0;
        z1 = cmk(1.0, 0.0);
        z2 = cmk(1.0, 0.0);
        const double tn = nA;
        nA = nB;
        nB = tn;
        rot = 1;
      } else if (act) {
        const double ag2 = cabs2(gd);
        if (ag2 > tol2 * nA * nB && ag2 > 0.0) {
          const double dl = 0.5 * (nB - nA), ad = fabs(dl);
          const double r2 = fma(dl, dl, ag2);
          const double r = r2 * rsqrt(r2);
          const double u = r + ad;
          const double q = rsqrt(2.0 * r * u);
          c = u * q;
          const double sq = copysign(q, dl);
          z1 = cmk(-sq * gd.x, sq * gd.y);  // -s e^{-i phi}
          z2 = cmk(sq * gd.x, sq * gd.y);   //  s e^{+i phi}
          const double tg = copysign(2.0 * ag2 * r * q * q, dl);
          if (ag2 > d.jac_quad2 * nA * nB) mybig = 1;  // an off-diagonal above the quadratic-convergence threshold
          nA -= tg;
          nB += tg;
          nA = nA > 0.0 ? nA : 0.0;
          nB = nB > 0.0 ? nB : 0.0;
          rot = 1;
          if (li == 0) myrot++;
        }
      }
      if (__any_sync(DQA_FULL, rot)) {
#pragma unroll
        for (int t = 0; t < RS; ++t) {
          const cplx a_ = A[t], b_ = B[t];
          A[t] = cfma(z1, b_, cscale(a_, c));
          B[t] = cfma(z2, a_, cscale(b_, c));
        }
      }
      // ---- pass one column on: B to the right after an even round, A back to the left after an odd one ----
      cplx* xb = X + (straddle ? bufsz : 0);
      double* xnb = xn + (straddle ? chi : 0);
      if (!straddle) {
        if (act) {
          cplx* dst = xb + (size_t)gr * mr;
#pragma unroll
          for (int t = 0; t < RS; ++t)
            if (li + 8 * t < mr) dst[li + 8 * t] = B[t];
          if (li == 0) xnb[gr] = nB;
        }
        __syncthreads();
        if (act) {
          const cplx* src = xb + (size_t)g * mr;
#pragma unroll
          for (int t = 0; t < RS; ++t)
            if (li + 8 * t < mr) B[t] = src[li + 8 * t];
          nB = xnb[g];
        }
      } else {
        if (act) {
          cplx* dst = xb + (size_t)gl * mr;
#pragma unroll
          for (int t = 0; t < RS; ++t)
            if (li + 8 * t < mr) dst[li + 8 * t] = A[t];
          if (li == 0) xnb[gl] = nA;
        }
        __syncthreads();
        if (act) {
          const cplx* src = xb + (size_t)g * mr;
#pragma unroll
          for (int t = 0; t < RS; ++t)
            if (li + 8 * t < mr) A[t] = src[li + 8 * t];
          nA = xnb[g];
        }
      }
    }
    if (myrot) atomicAdd(&sh_i[0], myrot);
    if (mybig) atomicOr(&sh_i[2], 1);
    __syncthreads();
    sweeps++;
    nrot_total += (unsigned long long)sh_i[0];
    converged = (sh_i[0] == 0) || (sh_i[2] == 0);  // same rule as k_jacobi
    __syncthreads();
  }
  // ---- park the columns in shared memory (mr x kp, any order: the tail sorts by singular value) ----
  __syncthreads();
  if (act) {
#pragma unroll
    for (int t = 0; t < RS; ++t) {
      const int r = li + 8 * t;
      if (r < mr) {
        X[r + (size_t)(2 * g) * mr] = B[t];
        X[r + (size_t)(2 * g + 1) * mr] = A[t];
      }
    }
  }
  __syncthreads();
  jacobi_tail(jac_tail_args(d), inst, l, mu, p_step, X, mr, mr, k, kp, sig, ssorted, order, sh_i, bond, converged, sweeps, nrot_total);
}

template <int RS>
__global__ void __launch_bounds__(RS * 32, RS > 8 ? 1 : (RS == 8 ? JSYS_MINB8 : (RS >= 6 ? 3 : (RS >= 4 ? 6 : 8)))) k_jacobi_sys(DqaDev d, int l, int mu, int p_step) {
  jacobi_sys_body<RS>(d, l, mu, p_step);
}
// (A 14-row-slot build for 48 < chi <= 56 does not exist: 448 threads are 14 warps, four on two of the SM's sub-partitions, and a
// sub-partition's 16 K registers hold four warps only up to 128 registers per thread -- 112 of which the column data would take.
// It spills at 128 (1 142 vs 982 ms per step for the shared-memory kernel at chi = 56) and cannot launch at 144.)



// Measured and rejected (round 2; commit bdc5484, branch exp/jacobi-row-per-lane): k_jacobi_blk, the same SVD with the columns
// laid out
// ROW-PER-LANE -- a warp owns 8 columns (two half-blocks of 4) for a whole block-round, a rotation of two of its columns
// is lane-local, the 4 pair dots of a sub-round are summed by a reduce-scatter butterfly (12 double shuffles) that leaves
// pair k in lanes 8k..8k+7, which compute the 4 parameter sets SIMT-parallel and publish them through 32 bytes of
// warp-private shared memory; only every 4th sub-round a half-block crosses shared memory behind a block barrier (odd-even
// transposition over half-blocks; 13 % fewer sweeps than the column ordering on the L factors of a chi=32 run).  Parity-green
// on the GPU (37 teacher-forced sub-steps), and SLOWER at every bond: chi=32 481 vs 418 ms per step, chi=24 324 vs 210,
// chi=16 98 vs 82, chi=10 62 vs 41 (batch 888).  Three of four barriers and exchanges are gone, but the two extra butterfly
// stages, the operand selects of the reduce-scatter and the parameter broadcast put as much latency back into every
// sub-round as the exchange took out of it, at 128 registers with spills (155 without the cap): the round is a latency
// chain (dot -> reduce -> two dependent rsqrt -> rotate), and the layout does not shorten it.

#ifndef DQA_EMUL
// ---------------------------------------------------------------------------------
// k_jacobi_sys_cl: the systolic array of k_jacobi_sys spread over a CLUSTER OF TWO CTAs (two SMs), for 2chi = 112 / 128
// rows (chi = 56, 64): at RS = 14 / 16 rows per lane the two register-resident columns of a group take 112 / 128
// registers, so the 56 / 64 groups of one SVD do not fit the register file of one SM (the one-CTA build spills and was
// measured slower than the shared-memory kernel).  Each CTA runs half of the line of groups; the column a group passes
// to its ring neighbour is stored straight into the neighbour's exchange slot -- in this CTA's shared memory or, for the
// two groups at the seam and the wrap-around, in the partner's (distributed shared memory) -- and the per-round
// block barrier becomes a cluster barrier.  The converged columns are parked in global memory (lbuf, L2-resident) and
// CTA 0 runs the common tail (ordering, the reference's trim rule, new site tensor).
// shared (per CTA): exchange buffers 2 x (groups/2 x mr) | sig | xn/ssorted | order | sh_i
// ---------------------------------------------------------------------------------
__host__ __device__ inline size_t svd_cl_smem(int chi, int cs) {
  const int gl = (chi + cs - 1) / cs;  // groups per CTA
  return sizeof(cplx) * (size_t)2 * gl * (2 * chi) + sizeof(double) * (4 * chi + 8) + sizeof(int) * (2 * chi + 8);
}

// CS CTAs per cluster (one SVD on CS SMs), MINB CTAs (of different SVDs) resident per SM, LPG lanes per group: a group
// holds RS * LPG rows of its two columns.  Instantiated as <16, 2, 1, 8> for 56 < chi <= 64.  Measured and dropped: LPG = 16
// (<8, 2, 1, 16>: twice the warps at half the rows per lane, one more shuffle stage) 1 386 vs 1 310 ms per step at chi = 64,
// batch 296; four CTAs per SVD with two SVDs interleaved per SM (<16, 4, 2, 8>) 1 322 ms; two-CTA clusters at two per SM for
// chi = 48 / 40 / 32 (<12|10|8, 2, 2|2|4, 8>) 504 vs 438, 363 vs 302, 507 vs 427 ms: below chi = 64 the one-CTA kernels win.
template <int RS, int CS, int MINB, int LPG>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(RS * LPG * LPG / (2 * CS), MINB) k_jacobi_sys_cl(DqaDev d, int l, int mu, int p_step) {
  DQA_DYN_SMEM(smraw);
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int inst = blockIdx.x / CS, tid = threadIdx.x, lane = tid & 31;
  const int chi = d.chi;
  const int GL = (chi + CS - 1) / CS;  // groups per CTA (blockDim.x = LPG * GL rounded up to a warp)
  cplx* lglob = d.lbuf + (size_t)inst * (size_t)(2 * chi) * (2 * chi);
  cplx* X = (cplx*)smraw;
  double* sig = (double*)(smraw + sizeof(cplx) * (size_t)2 * GL * (2 * chi));
  double* ssorted = sig + 2 * chi;
  int* order = (int*)(ssorted + 2 * chi + 8);
  int* sh_i = order + 2 * chi;  // [0]=rotation count, [1]=n_chi, [2]=some rotated pair was above the quadratic threshold
  int* bond = d.bond + inst * (d.N + 1);
  const int mr = 2 * bond[l];
  const int* qd1 = d.qdim + ((size_t)inst * (d.N + 1) + (l + 1)) * d.smax;
  const int Q = qdim_prefix(qd1, d.N - l);
  const int k = mr < Q ? mr : Q;
  const int kp = k + (k & 1), h = kp >> 1;  // h groups hold the kp columns
  const int gloc = tid / LPG, li = lane & (LPG - 1);
  const int g = crank * GL + gloc;           // position in the line of groups
  const bool act = gloc < GL && g < h;
  double* xn = ssorted;  // [2][GL] squared norms in flight

  cplx A[RS], B[RS];
#pragma unroll
  for (int t = 0; t < RS; ++t) {
    const int r = li + LPG * t;
    const bool in = act && r < mr;
    B[t] = (in && 2 * g < k) ? lglob[r + (size_t)(2 * g) * (2 * chi)] : cmk(0.0, 0.0);
    A[t] = (in && 2 * g + 1 < k) ? lglob[r + (size_t)(2 * g + 1) * (2 * chi)] : cmk(0.0, 0.0);
  }
  // ring neighbours and where their exchange slots live
  const int gr = (g + 1 < h) ? g + 1 : 0, gl = (g > 0) ? g - 1 : h - 1;
  cplx* Xr = cluster.map_shared_rank(X, gr / GL);
  cplx* Xl = cluster.map_shared_rank(X, gl / GL);
  double* xnr = cluster.map_shared_rank(xn, gr / GL);
  double* xnl = cluster.map_shared_rank(xn, gl / GL);
  const int sr = gr % GL, sl = gl % GL;
  const size_t bufsz = (size_t)GL * mr;
  const double tol = sqrt((double)mr) * 2.220446049250313e-16, tol2 = tol * tol;
  int sweeps = 0, converged = (k <= 1);
  unsigned long long nrot_total = 0;
  cluster.sync();  // all CTAs of the cluster are resident before anybody stores into the partner's shared memory
  for (int sweep = 0; sweep < d.max_sweeps && !converged; ++sweep) {
    double nA = 0.0, nB = 0.0;
#pragma unroll
    for (int t = 0; t < RS; ++t) {
      nA += cabs2(A[t]);
      nB += cabs2(B[t]);
    }
    if (LPG == 16) {
      nA += __shfl_xor_sync(DQA_FULL, nA, 8);
      nB += __shfl_xor_sync(DQA_FULL, nB, 8);
    }
    nA += __shfl_xor_sync(DQA_FULL, nA, 4);
    nB += __shfl_xor_sync(DQA_FULL, nB, 4);
    nA += __shfl_xor_sync(DQA_FULL, nA, 2);
    nB += __shfl_xor_sync(DQA_FULL, nB, 2);
    nA += __shfl_xor_sync(DQA_FULL, nA, 1);
    nB += __shfl_xor_sync(DQA_FULL, nB, 1);
    if (tid == 0) {
      sh_i[0] = 0;
      sh_i[2] = 0;
    }
    int myrot = 0, mybig = 0;
    for (int round = 0; round < kp; ++round) {
      const bool straddle = round & 1;
      const bool idle = straddle && g == 0;
      cplx g0 = cmk(0.0, 0.0), g1 = cmk(0.0, 0.0);
#pragma unroll
      for (int t = 0; t < RS; t += 2) {
        g0 = cfmac(B[t], A[t], g0);
        if (t + 1 < RS) g1 = cfmac(B[t + 1], A[t + 1], g1);
      }
      cplx gd = cadd(g0, g1);
      if (LPG == 16) {
        gd.x += __shfl_xor_sync(DQA_FULL, gd.x, 8);
        gd.y += __shfl_xor_sync(DQA_FULL, gd.y, 8);
      }
      gd.x += __shfl_xor_sync(DQA_FULL, gd.x, 4);
      gd.y += __shfl_xor_sync(DQA_FULL, gd.y, 4);
      gd.x += __shfl_xor_sync(DQA_FULL, gd.x, 2);
      gd.y += __shfl_xor_sync(DQA_FULL, gd.y, 2);
      gd.x += __shfl_xor_sync(DQA_FULL, gd.x, 1);
      gd.y += __shfl_xor_sync(DQA_FULL, gd.y, 1);
      double c = 1.0;
      cplx z1 = cmk(0.0, 0.0), z2 = cmk(0.0, 0.0);
      int rot = 0;
      if (idle && act) {
        c = 0.0;
        z1 = cmk(1.0, 0.0);
        z2 = cmk(1.0, 0.0);
        const double tn = nA;
        nA = nB;
        nB = tn;
        rot = 1;
      } else if (act) {
        const double ag2 = cabs2(gd);
        if (ag2 > tol2 * nA * nB && ag2 > 0.0) {
          const double dl = 0.5 * (nB - nA), ad = fabs(dl);
          const double r2 = fma(dl, dl, ag2);
          const double r = r2 * rsqrt(r2);
          const double u = r + ad;
          const double q = rsqrt(2.0 * r * u);
          c = u * q;
          const double sq = copysign(q, dl);
          z1 = cmk(-sq * gd.x, sq * gd.y);
          z2 = cmk(sq * gd.x, sq * gd.y);
          const double tg = copysign(2.0 * ag2 * r * q * q, dl);
          if (ag2 > d.jac_quad2 * nA * nB) mybig = 1;
          nA -= tg;
          nB += tg;
          nA = nA > 0.0 ? nA : 0.0;
          nB = nB > 0.0 ? nB : 0.0;
          rot = 1;
          if (li == 0) myrot++;
        }
      }
      if (__any_sync(DQA_FULL, rot)) {
#pragma unroll
        for (int t = 0; t < RS; ++t) {
          const cplx a_ = A[t], b_ = B[t];
          A[t] = cfma(z1, b_, cscale(a_, c));
          B[t] = cfma(z2, a_, cscale(b_, c));
        }
      }
      // ---- pass one column on: B to the right after an even round, A back to the left after an odd one ----
      const size_t boff = straddle ? bufsz : 0;
      const int noff = straddle ? GL : 0;
      if (!straddle) {
        if (act) {
          cplx* dst = Xr + boff + (size_t)sr * mr;
#pragma unroll
          for (int t = 0; t < RS; ++t)
            if (li + LPG * t < mr) dst[li + LPG * t] = B[t];
          if (li == 0) xnr[noff + sr] = nB;
        }
        cluster.sync();
        if (act) {
          const cplx* src = X + boff + (size_t)gloc * mr;
#pragma unroll
          for (int t = 0; t < RS; ++t)
            if (li + LPG * t < mr) B[t] = src[li + LPG * t];
          nB = xn[noff + gloc];
        }
      } else {
        if (act) {
          cplx* dst = Xl + boff + (size_t)sl * mr;
#pragma unroll
          for (int t = 0; t < RS; ++t)
            if (li + LPG * t < mr) dst[li + LPG * t] = A[t];
          if (li == 0) xnl[noff + sl] = nA;
        }
        cluster.sync();
        if (act) {
          const cplx* src = X + boff + (size_t)gloc * mr;
#pragma unroll
          for (int t = 0; t < RS; ++t)
            if (li + LPG * t < mr) A[t] = src[li + LPG * t];
          nA = xn[noff + gloc];
        }
      }
    }
    if (myrot) atomicAdd(&sh_i[0], myrot);
    if (mybig) atomicOr(&sh_i[2], 1);
    cluster.sync();
    sweeps++;
    int rot_all = 0, big_all = 0;
#pragma unroll
    for (int rk = 0; rk < CS; ++rk) {
      const int* sh_r = cluster.map_shared_rank(sh_i, rk);
      rot_all += sh_r[0];
      big_all |= sh_r[2];
    }
    nrot_total += (unsigned long long)rot_all;
    converged = (rot_all == 0) || (big_all == 0);  // same rule as k_jacobi
    cluster.sync();  // the partner has read this CTA's counters before they are reset
  }
  // ---- park the columns in global memory (any order: the tail sorts by singular value) ----
  if (act) {
#pragma unroll
    for (int t = 0; t < RS; ++t) {
      const int r = li + LPG * t;
      if (r < mr) {
        lglob[r + (size_t)(2 * g) * (2 * chi)] = B[t];
        lglob[r + (size_t)(2 * g + 1) * (2 * chi)] = A[t];
      }
    }
  }
  __threadfence();
  cluster.sync();
  if (crank != 0) return;  // no distributed-shared-memory access happens after the barrier above
  jacobi_tail(jac_tail_args(d), inst, l, mu, p_step, lglob, 2 * chi, mr, k, kp, sig, ssorted, order, sh_i, bond, converged, sweeps, nrot_total);
}
#endif  // DQA_EMUL

// ---------------------------------------------------------------------------------
// k_mixer: A[a,s',c] <- sum_s u[s',s] A[a,s,c], u = [[cos b, i sin b],[i sin b, cos b]]
// (lib/dQA_utils.py:26-30 + lib/tenn.py:67-81 for a bond-1 MPO), all sites in one launch.
// grid (N, batch)
// ---------------------------------------------------------------------------------
// mode 0: cb, sb as given; 1: per-instance dt, (1 - s_p) given; 2: step counter on the device (graph replay)
__global__ void k_mixer(DqaDev d, double cb, double sb, double one_minus_s, int mode) {
  const int site = blockIdx.x, inst = blockIdx.y;
  if (mode == 2) {
    const int p = *d.pdev;
    if (d.ntab > 1) {
      one_minus_s = 1.0 - (double)(p + 1) / d.P;
      mode = 1;
    } else {
      cb = d.mixtab[2 * p];
      sb = d.mixtab[2 * p + 1];
    }
  }
  if (mode == 1) {  // instances with their own dt: beta = (1 - s_p) * dt_inst (dQA_mps.py:88-89)
    const double beta = one_minus_s * d.dtv[inst];
    cb = cos(beta);
    sb = sin(beta);
  }
  const int chi = d.chi;
  const int* bond = d.bond + inst * (d.N + 1);
  const int bl = bond[site], br = bond[site + 1];
  cplx* a = d.mps + ((size_t)inst * d.N + site) * (size_t)(2 * chi * chi);
  for (int i = threadIdx.x; i < bl * br; i += blockDim.x) {
    const int c = i % br, r = i / br;
    cplx* p0 = a + (r * 2 + 0) * chi + c;
    cplx* p1 = a + (r * 2 + 1) * chi + c;
    const cplx v0 = *p0, v1 = *p1;
    // i*sb*v = (-sb*v.y, sb*v.x)
    *p0 = cmk(cb * v0.x - sb * v1.y, cb * v0.y + sb * v1.x);
    *p1 = cmk(cb * v1.x - sb * v0.y, cb * v1.y + sb * v0.x);
  }
}

// ---------------------------------------------------------------------------------
// k_energy: T_{mu,k} = <psi| prod_i diag_s(exp(i pi k e[mu,i,s]/D)) |psi>, k = 0..D/2
//   E'[b,d] = sum_s ph_s sum_{a,c} E[a,c] A[a,s,b] conj(A[c,s,d])   (lib/tenn.py:102-107)
// grid (kh, n_xi, batch); shared: E, E2, W, As (chi*chi each)
// n_phys_sites < N_sites supports the ancilla variant (lib/dQA_loss.py:28-46): trailing
// sites carry the identity.
// ---------------------------------------------------------------------------------
// E, E2 and one W buffer per spin value (4 chi^2) while that fits; beyond chi = 56 the two spin values
// share one W buffer and are processed one after the other (3 chi^2).
// Row stride of the shared-memory environments: = 2 (mod 8) complex, so the four k-rows an 8-lane quarter-warp
// touches in one DMMA fragment load fall into different bank groups (a stride that is a multiple of 8 complex = 128 B
// makes every fragment load a 4-way bank conflict: measured 75 % of the kernel's shared-memory wavefronts).
__host__ __device__ inline int energy_ld(int chi) { return ((chi + 7) & ~7) + 2; }
__host__ __device__ inline bool energy_two_w(int chi) { return sizeof(cplx) * (size_t)4 * chi * energy_ld(chi) <= 200 * 1024; }
__host__ __device__ inline size_t energy_smem(int chi) { return sizeof(cplx) * (size_t)(energy_two_w(chi) ? 4 : 3) * chi * energy_ld(chi); }

// where the site tensors of one MPS live: the plan's padded layout or a packed external MPS
struct EnergyView {
  const cplx* base;   // first site
  const long* off;    // per-site element offsets (NULL: site i at i * 2*chi*chi, row stride chi)
  const int* bond;    // n_sites+1 bond dimensions
  int n_sites;        // sites swept
  int n_phys;         // sites that carry the phase operator; trailing ones (ancillas) carry the identity
  int chi;            // capacity of the shared-memory environments
};

// T = <psi| prod_{i<n_phys} diag_s(ph_i(s)) |psi> for one (mu,k): transfer-matrix sweep with
// FP64 tensor-core tiles.  smem: E, E2, W0, W1 (chi*chi each).
// The two GEMM phases of one site of the transfer-matrix sweep, NT row tiles per warp task: the tiles of a task share
// the site-tensor fragment (global memory through L1), which is loaded once per k-step for 4*NT DMMAs.
// phase 1: W_s[c,b] = sum_a E[a,c] A[a,s,b]
template <int NT>
__device__ __forceinline__ void energy_phase1(const cplx* E, cplx* W0, cplx* W1, bool two_w, const cplx* a, int lda, int lde, int bl,
                                              int br, int s_lo, int s_hi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
  const int tr = (bl + 7) >> 3, tc = (br + 7) >> 3, trg = (tr + NT - 1) / NT;
  for (int task = warp; task < (s_hi - s_lo) * tc * trg; task += nwarp) {
    const int s = s_lo + task / (tc * trg), rem = task % (tc * trg), b0 = (rem % tc) * 8, rt0 = (rem / tc) * NT;
    cplx* W = (two_w && s) ? W1 : W0;
    ZTile acc[NT];
#pragma unroll
    for (int q = 0; q < NT; ++q) acc[q] = ztile_zero();
    warp_zgemm_8x8_rows<NT>(
        acc, bl,
        [&](int q, int row, int kk) {
          const int c = (rt0 + q) * 8 + row;
          return (c < bl && kk < bl) ? E[kk * lde + c] : cmk(0.0, 0.0);
        },
        [&](int kk, int col) { return (kk < bl && b0 + col < br) ? a[((size_t)kk * 2 + s) * lda + b0 + col] : cmk(0.0, 0.0); });
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      const int c = (rt0 + q) * 8 + g, b = b0 + 2 * t;
      if (c < bl) {
        if (b < br) W[c * lde + b] = cmk(acc[q].r0, acc[q].i0);
        if (b + 1 < br) W[c * lde + b + 1] = cmk(acc[q].r1, acc[q].i1);
      }
    }
  }
}
// phase 2: E2[b,dd] (+)= sum_s ph_s sum_c W_s[c,b] conj(A[c,s,dd])
template <int NT>
__device__ __forceinline__ void energy_phase2(cplx* E2, const cplx* W0, const cplx* W1, bool two_w, const cplx* a, int lda, int lde,
                                              int bl, int br, int s_lo, int s_hi, cplx ph0, cplx ph1, bool first) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
  const int tc = (br + 7) >> 3, tcg = (tc + NT - 1) / NT;
  for (int task = warp; task < tc * tcg; task += nwarp) {
    const int d0 = (task % tc) * 8, bt0 = (task / tc) * NT;
    cplx out[NT][2];
#pragma unroll
    for (int q = 0; q < NT; ++q) out[q][0] = out[q][1] = cmk(0.0, 0.0);
    for (int s = s_lo; s < s_hi; ++s) {
      const cplx* W = (two_w && s) ? W1 : W0;
      ZTile acc[NT];
#pragma unroll
      for (int q = 0; q < NT; ++q) acc[q] = ztile_zero();
      warp_zgemm_8x8_rows<NT>(
          acc, bl,
          [&](int q, int row, int kk) {
            const int b = (bt0 + q) * 8 + row;
            return (b < br && kk < bl) ? W[kk * lde + b] : cmk(0.0, 0.0);
          },
          [&](int kk, int col) { return (kk < bl && d0 + col < br) ? cconj(a[((size_t)kk * 2 + s) * lda + d0 + col]) : cmk(0.0, 0.0); });
      const cplx ph = s ? ph1 : ph0;
#pragma unroll
      for (int q = 0; q < NT; ++q) {
        out[q][0] = cfma(ph, cmk(acc[q].r0, acc[q].i0), out[q][0]);
        out[q][1] = cfma(ph, cmk(acc[q].r1, acc[q].i1), out[q][1]);
      }
    }
#pragma unroll
    for (int q = 0; q < NT; ++q) {
      const int b = (bt0 + q) * 8 + g, dd = d0 + 2 * t;
      if (b < br) {
        if (dd < br) E2[b * lde + dd] = first ? out[q][0] : cadd(E2[b * lde + dd], out[q][0]);
        if (dd + 1 < br) E2[b * lde + dd + 1] = first ? out[q][1] : cadd(E2[b * lde + dd + 1], out[q][1]);
      }
    }
  }
}

template <bool PAIR>  // PAIR: row tiles may be paired (bonds > 16); the unpaired build keeps 60 registers for small bonds
__device__ __forceinline__ cplx energy_sweep(const EnergyView& v, const uint8_t* hb, cplx wk, cplx* smem) {
  const int tid = threadIdx.x, nwarp = blockDim.x >> 5;
  const int chi = v.chi, lde = energy_ld(chi);
  const size_t rsz = (size_t)chi * chi, esz = (size_t)chi * lde;
  const bool two_w = energy_two_w(chi);
  cplx* E = smem;
  cplx* E2 = E + esz;
  cplx* W0 = E2 + esz;
  cplx* W1 = W0 + esz;  // only touched when two_w
  if (tid == 0) E[0] = cmk(1.0, 0.0);
  __syncthreads();
  for (int i = 0; i < v.n_sites; ++i) {
    const int bl = v.bond[i], br = v.bond[i + 1];
    const cplx* a = v.off ? v.base + v.off[i] : v.base + (size_t)i * (2 * rsz);
    const int lda = v.off ? br : chi;
    const int tr = (bl + 7) >> 3, tc = (br + 7) >> 3;
    const cplx ph0 = (i < v.n_phys && (hb[i] ^ 0)) ? wk : cmk(1.0, 0.0), ph1 = (i < v.n_phys && (hb[i] ^ 1)) ? wk : cmk(1.0, 0.0);
    // two_w: phase 1 computes W_0 and W_1 together, phase 2 combines them (2 barriers per site);
    // otherwise the spin values take turns in the single W buffer (4 barriers per site)
    const int ngrp = two_w ? 1 : 2;
    for (int grp_s = 0; grp_s < ngrp; ++grp_s) {
      const int s_lo = two_w ? 0 : grp_s, s_hi = two_w ? 2 : grp_s + 1;
      // pair up row tiles only when every warp still gets a task
      if (PAIR && (s_hi - s_lo) * tc * ((tr + 1) >> 1) >= nwarp) energy_phase1<2>(E, W0, W1, two_w, a, lda, lde, bl, br, s_lo, s_hi);
      else energy_phase1<1>(E, W0, W1, two_w, a, lda, lde, bl, br, s_lo, s_hi);
      __syncthreads();
      if (PAIR && tc * ((tc + 1) >> 1) >= nwarp) energy_phase2<2>(E2, W0, W1, two_w, a, lda, lde, bl, br, s_lo, s_hi, ph0, ph1, grp_s == 0);
      else energy_phase2<1>(E2, W0, W1, two_w, a, lda, lde, bl, br, s_lo, s_hi, ph0, ph1, grp_s == 0);
      __syncthreads();
    }
    cplx* t2 = E;
    E = E2;
    E2 = t2;
  }
  return E[0];
}

template <bool PAIR>
__global__ void __launch_bounds__(256, PAIR ? 3 : 4) k_energy(DqaDev d) {
  DQA_DYN_SMEM(smraw);
  const int k = blockIdx.x, mu = blockIdx.y, inst = blockIdx.z;
  EnergyView v;
  v.base = d.mps + (size_t)inst * d.N * (size_t)(2 * d.chi * d.chi);
  v.off = nullptr;
  v.bond = d.bond + inst * (d.N + 1);
  v.n_sites = d.N;
  v.n_phys = d.N;
  v.chi = d.chi;
  const cplx r = energy_sweep<PAIR>(v, d.hbit + ((size_t)inst * d.n_xi + mu) * d.N, d.phase[k], (cplx*)smraw);
  if (threadIdx.x == 0) d.tmk[((size_t)inst * d.n_xi + mu) * d.kh + k] = r;
}

// mydQA_ancilla.compute_loss (lib/dQA_loss.py:67-78): an external, packed MPS of n_sites >= n_phys sites
// (trailing ancilla sites carry the identity), bonds <= chi.  grid (kh, n_xi); tmk [n_xi][kh].
template <bool PAIR>
__global__ void __launch_bounds__(256) k_energy_ext(const cplx* mps, const long* off, const int* bond, int n_sites, int n_phys,
                                                    int chi, const uint8_t* hbit, const cplx* phase, int kh, cplx* tmk) {
  DQA_DYN_SMEM(smraw);
  const int k = blockIdx.x, mu = blockIdx.y;
  EnergyView v;
  v.base = mps;
  v.off = off;
  v.bond = bond;
  v.n_sites = n_sites;
  v.n_phys = n_phys;
  v.chi = chi;
  const cplx r = energy_sweep<PAIR>(v, hbit + (size_t)mu * n_phys, phase[k], (cplx*)smraw);
  if (threadIdx.x == 0) tmk[(size_t)mu * kh + k] = r;
}

// eps = (1/n_phys) sum_mu [...] for the external-MPS variant (same combination as k_energy_reduce)
__global__ void k_energy_ext_reduce(const cplx* tmk, const cplx* gk, int n_xi, int kh, int D, int n_phys, cplx* out) {
  const int lane = threadIdx.x;
  cplx acc = cmk(0.0, 0.0);
  const int n = n_xi * kh;
  for (int i = lane; i < n; i += 32) {
    const int k = i % kh;
    const cplx t = tmk[i];
    acc = cfma(gk[k], t, acc);
    if (k > 0 && 2 * k != D) acc = cfmac(gk[D - k], t, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[0] = cscale(acc, 1.0 / n_phys);
}

// eps = (1/N) sum_mu [ sum_{k<=D/2} g_k T_k + sum_{0<k<D/2} g_{D-k} conj(T_k) ]; one warp per instance
__global__ void k_energy_reduce(DqaDev d, cplx* out, size_t stride) {
  const int inst = blockIdx.x, lane = threadIdx.x;
  if (!out) out = d.eps + (*d.pdev + 1);  // graph replay: the slot of the step in flight
  cplx acc = cmk(0.0, 0.0);
  const int n = d.n_xi * d.kh;
  for (int i = lane; i < n; i += 32) {
    const int k = i % d.kh;
    const cplx t = d.tmk[(size_t)inst * n + i];
    acc = cfma(d.gk[k], t, acc);
    if (k > 0 && 2 * k != d.D) acc = cfmac(d.gk[d.D - k], t, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[(size_t)inst * stride] = cscale(acc, 1.0 / d.N);
}


// end of a replayed step: advance the device step counter
__global__ void k_advance(DqaDev d) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *d.pdev += 1;
}
