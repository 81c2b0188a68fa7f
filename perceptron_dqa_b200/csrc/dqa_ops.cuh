// dqa_ops.cuh -- op-level kernels behind the reference's lib/tenn.py functions, for arbitrary
// (dense) MPS/MPO site tensors: apply_mpsmpo, braket, qr_stabilized / right_canonize_twosites,
// svd_truncated / compress_normalize_onesite.  The driver's hot loop does NOT use these (it runs
// the fused sector kernels of dqa_kernels.cuh); they exist so that callers of the reference's
// op-level API find the same functions, computed on the GPU.  One CTA per matrix, operands in
// global memory (L2), sizes unrestricted; tuned for the sizes the reference's own tests use.
#pragma once
#include "dqa_kernels.cuh"

// ---------------------------------------------------------------------------------
// contract_mps_mpo_single_site (lib/tenn.py:67-81):
//   out[(a*wl+j), l, (c*wr+k)] = sum_b A[a,b,c] * W[j,k,l,b]
// A (al,2,ar), W (wl,wr,2,2), out (al*wl, 2, ar*wr); grid-stride over outputs.
// ---------------------------------------------------------------------------------
__global__ void k_g_contract_site(const cplx* A, int al, int ar, const cplx* W, int wl, int wr, cplx* out) {
  const long total = (long)al * wl * 2 * ar * wr;
  const int orr = ar * wr;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int col = (int)(i % orr);
    const long rest = i / orr;
    const int l = (int)(rest % 2);
    const int row = (int)(rest / 2);
    const int c = col / wr, k = col % wr, a = row / wl, j = row % wl;
    cplx acc = cmk(0.0, 0.0);
    for (int b = 0; b < 2; ++b) acc = cfma(A[((size_t)a * 2 + b) * ar + c], W[(((size_t)j * wr + k) * 2 + l) * 2 + b], acc);
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------------
// one site of braket (lib/tenn.py:93-109): env'[b,d] = sum_{a,c,x} env[a,c] B[a,x,b] conj(K[c,x,d])
// env (bl x kl) or NULL for the first site (uses a=c=0), B (bl,2,br), K (kl,2,kr); tmp (kl*2*br).
// Single CTA, two phases.
// ---------------------------------------------------------------------------------
__global__ void k_g_braket_site(const cplx* env, const cplx* B, int bl, int br, const cplx* K, int kl, int kr, cplx* tmp, cplx* out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  // tmp[c,x,b] = sum_a env[a,c] B[a,x,b]
  for (int i = tid; i < kl * 2 * br; i += nt) {
    const int b = i % br, x = (i / br) % 2, c = i / (2 * br);
    cplx acc = cmk(0.0, 0.0);
    if (env) {
      for (int a = 0; a < bl; ++a) acc = cfma(env[(size_t)a * kl + c], B[((size_t)a * 2 + x) * br + b], acc);
    } else {
      acc = (c == 0) ? B[(size_t)x * br + b] : cmk(0.0, 0.0);  // prev = mm[0,0,:,:]
    }
    tmp[i] = acc;
  }
  __syncthreads();
  const int kl_eff = env ? kl : 1;
  for (int i = tid; i < br * kr; i += nt) {
    const int dd = i % kr, b = i / kr;
    cplx acc = cmk(0.0, 0.0);
    for (int c = 0; c < kl_eff; ++c)
      for (int x = 0; x < 2; ++x) acc = cfmac(tmp[((size_t)c * 2 + x) * br + b], K[((size_t)c * 2 + x) * kr + dd], acc);
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------------
// Householder QR of a dense m x n matrix held column-major in global memory (ld = m), one CTA.
//   on exit: Mw holds R in its upper triangle (diagonal sign NOT yet fixed) and the reflectors
//   below; tau[k], beta[k]; then Q (m x kq col-major, ld m) is formed in Qw.  kq = min(m,n).
// This is zgeqr2 + zung2r; qr_stabilized (lib/tenn.py:294-304) multiplies by sign(diag R) with
// sgn(0) := +1 when the caller extracts Q and R.
// ---------------------------------------------------------------------------------
__device__ void g_householder_qr(cplx* Mw, int m, int n, cplx* tau, double* beta, cplx* Qw, double* red) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const int kq = m < n ? m : n;
  for (int k = 0; k < kq; ++k) {
    cplx* colk = Mw + (size_t)k * m;
    double ss = 0.0;
    for (int i = k + 1 + tid; i < m; i += nt) ss += cabs2(colk[i]);
    const cplx alpha = colk[k];
    ss = block_sum(ss, red);
    double b_;
    cplx t_, scal;
    householder_gen(alpha, ss, b_, t_, scal);
    for (int i = k + 1 + tid; i < m; i += nt) colk[i] = cmul(colk[i], scal);
    __syncthreads();
    if (tid == 0) {
      tau[k] = t_;
      beta[k] = b_;
      colk[k] = cmk(b_, 0.0);
    }
    const cplx tk = cconj(t_);
    if (tk.x != 0.0 || tk.y != 0.0) {
      for (int j = k + 1 + warp; j < n; j += nwarp) {
        cplx* colj = Mw + (size_t)j * m;
        cplx w = (lane == 0) ? colj[k] : cmk(0.0, 0.0);
        for (int i = k + 1 + lane; i < m; i += 32) w = cfmac(colj[i], colk[i], w);
        w = warp_sum(w);
        const cplx tw = cmul(tk, w);
        if (lane == 0) colj[k] = csub(colj[k], tw);
        for (int i = k + 1 + lane; i < m; i += 32) colj[i] = cfms(tw, colk[i], colj[i]);
      }
    }
    __syncthreads();
  }
  // Q = H_0 ... H_{kq-1} [I;0]
  for (int i = tid; i < m * kq; i += nt) {
    const int r = i % m, c = i / m;
    Qw[i] = (r == c) ? cmk(1.0, 0.0) : cmk(0.0, 0.0);
  }
  __syncthreads();
  for (int i = kq - 1; i >= 0; --i) {
    const cplx ti = tau[i];
    const cplx* vi = Mw + (size_t)i * m;
    if (ti.x != 0.0 || ti.y != 0.0) {
      for (int j = i + warp; j < kq; j += nwarp) {
        cplx* qj = Qw + (size_t)j * m;
        cplx w = (lane == 0) ? qj[i] : cmk(0.0, 0.0);
        for (int r = i + 1 + lane; r < m; r += 32) w = cfmac(qj[r], vi[r], w);
        w = warp_sum(w);
        const cplx tw = cmul(ti, w);
        if (lane == 0) qj[i] = csub(qj[i], tw);
        for (int r = i + 1 + lane; r < m; r += 32) qj[r] = cfms(tw, vi[r], qj[r]);
      }
    }
    __syncthreads();
  }
}

// qr_stabilized(x) (lib/tenn.py:294-304): X (m x n) row-major -> Q (m x kq) row-major, R (kq x n) row-major.
// work: Mw (m*n) | Qw (m*kq) | tau (kq) cplx, beta (kq) doubles
__global__ void __launch_bounds__(1024) k_g_qr_stabilized(const cplx* X, int m, int n, cplx* Q, cplx* R, cplx* work) {
  __shared__ double red[64];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int kq = m < n ? m : n;
  cplx* Mw = work;
  cplx* Qw = Mw + (size_t)m * n;
  cplx* tau = Qw + (size_t)m * kq;
  double* beta = (double*)(tau + kq);
  for (int i = tid; i < m * n; i += nt) {
    const int r = i / n, c = i % n;
    Mw[r + (size_t)c * m] = X[i];
  }
  __syncthreads();
  g_householder_qr(Mw, m, n, tau, beta, Qw, red);
  for (int i = tid; i < m * kq; i += nt) {
    const int r = i / kq, c = i % kq;
    const cplx v = Qw[r + (size_t)c * m];
    Q[i] = beta[c] < 0.0 ? cmk(-v.x, -v.y) : v;
  }
  for (int i = tid; i < kq * n; i += nt) {
    const int r = i / n, c = i % n;
    cplx v = cmk(0.0, 0.0);
    if (c >= r) {
      v = Mw[r + (size_t)c * m];
      if (beta[r] < 0.0) v = cmk(-v.x, -v.y);
    }
    R[i] = v;
  }
}

// right_canonize_twosites(A, B) (lib/tenn.py:152-176):
//   M[(s,r), l] = B[l,s,r];  M = Q R (stabilised);  B' [q,s,r] = Q[(s,r), q];
//   A'[a,s,d] = sum_c A[a,s,c] R[d,c];  A' /= |A'|_F.
// A (al,2,bl) , B (bl,2,br) -> Anew (al,2,kq), Bnew (kq,2,br), kq = min(2*br, bl)
__global__ void __launch_bounds__(1024) k_g_right_canonize_twosites(const cplx* A, int al, const cplx* B, int bl, int br, cplx* Anew, cplx* Bnew, cplx* work) {
  __shared__ double red[64];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int m = 2 * br, n = bl, kq = m < n ? m : n;
  cplx* Mw = work;
  cplx* Qw = Mw + (size_t)m * n;
  cplx* tau = Qw + (size_t)m * kq;
  double* beta = (double*)(tau + kq);
  for (int i = tid; i < m * n; i += nt) {
    const int l = i % n, sr = i / n;  // sr = s*br + r
    Mw[sr + (size_t)l * m] = B[(size_t)l * m + sr];
  }
  __syncthreads();
  g_householder_qr(Mw, m, n, tau, beta, Qw, red);
  for (int i = tid; i < kq * m; i += nt) {
    const int sr = i % m, qq = i / m;
    const cplx v = Qw[sr + (size_t)qq * m];
    Bnew[(size_t)qq * m + sr] = beta[qq] < 0.0 ? cmk(-v.x, -v.y) : v;
  }
  // A' = A R^T, then Frobenius normalisation
  double ss = 0.0;
  for (int i = tid; i < al * 2 * kq; i += nt) {
    const int dd = i % kq, as = i / kq;
    cplx acc = cmk(0.0, 0.0);
    for (int c = dd; c < n; ++c) acc = cfma(A[(size_t)as * n + c], Mw[dd + (size_t)c * m], acc);  // R[dd,c], c >= dd
    if (beta[dd] < 0.0) acc = cmk(-acc.x, -acc.y);
    Anew[i] = acc;
    ss += cabs2(acc);
  }
  ss = block_sum(ss, red);
  const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
  for (int i = tid; i < al * 2 * kq; i += nt) Anew[i] = cscale(Anew[i], inv);
}

// ---------------------------------------------------------------------------------
// thin SVD of a dense m x n matrix by one-sided Jacobi on the columns of the taller of X, X^H
// (global-memory operands, one CTA, one warp per column pair), with accumulated right rotations.
//   work: Y (Mr x Nc col-major) | V (Nc x Nc col-major) | sig (Nc) | order (Nc ints)
//   out: U (m x k) row-major, s (k), Vh (k x n) row-major, k = min(m,n), descending order.
// np.linalg.svd at lib/tenn.py:402.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_g_svd(const cplx* X, int m, int n, cplx* U, double* S, cplx* Vh, cplx* work, int max_sweeps, int* status) {
  __shared__ int s_rot;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const bool tall = m >= n;
  const int Mr = tall ? m : n, Nc = tall ? n : m;
  cplx* Y = work;
  cplx* V = Y + (size_t)Mr * Nc;
  double* sig = (double*)(V + (size_t)Nc * Nc);
  int* order = (int*)(sig + Nc);
  for (int i = tid; i < Mr * Nc; i += nt) {
    const int r = i % Mr, c = i / Mr;
    Y[i] = tall ? X[(size_t)r * n + c] : cconj(X[(size_t)c * n + r]);
  }
  for (int i = tid; i < Nc * Nc; i += nt) V[i] = (i % Nc == i / Nc) ? cmk(1.0, 0.0) : cmk(0.0, 0.0);
  __syncthreads();
  const int kp = Nc + (Nc & 1), half = kp >> 1;
  const double tol = sqrt((double)Mr) * 2.220446049250313e-16, tol2 = tol * tol;
  int converged = (Nc <= 1);
  for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
    if (tid == 0) s_rot = 0;
    __syncthreads();
    for (int round = 0; round < kp - 1; ++round) {
      for (int pr = warp; pr < half; pr += nwarp) {
        int p, q;
        rr_pair(kp, round, pr, p, q);
        if (q >= Nc) continue;
        cplx* xp = Y + (size_t)p * Mr;
        cplx* xq = Y + (size_t)q * Mr;
        double a = 0.0, b = 0.0;
        cplx g = cmk(0.0, 0.0);
        for (int r = lane; r < Mr; r += 32) {
          const cplx vp = xp[r], vq = xq[r];
          a += cabs2(vp);
          b += cabs2(vq);
          g = cfmac(vq, vp, g);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        g = warp_sum(g);
        const double ag2 = cabs2(g);
        if (ag2 > tol2 * a * b && ag2 > 0.0) {
          const double ag = sqrt(ag2);
          const double zeta = (b - a) / (2.0 * ag);
          const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + tt * tt), sn = c * tt;
          const cplx ph = cmk(g.x / ag, -g.y / ag);
          const cplx sph = cscale(ph, sn), cph = cscale(ph, c);
          for (int r = lane; r < Mr; r += 32) {
            const cplx vp = xp[r], vq = xq[r];
            xp[r] = cfms(sph, vq, cscale(vp, c));
            xq[r] = cfma(cph, vq, cscale(vp, sn));
          }
          cplx* vpc = V + (size_t)p * Nc;
          cplx* vqc = V + (size_t)q * Nc;
          for (int r = lane; r < Nc; r += 32) {
            const cplx vp = vpc[r], vq = vqc[r];
            vpc[r] = cfms(sph, vq, cscale(vp, c));
            vqc[r] = cfma(cph, vq, cscale(vp, sn));
          }
          if (lane == 0) atomicAdd(&s_rot, 1);
        }
      }
      __syncthreads();
    }
    converged = (s_rot == 0);
    __syncthreads();
  }
  for (int j = warp; j < Nc; j += nwarp) {
    double a = 0.0;
    const cplx* xj = Y + (size_t)j * Mr;
    for (int r = lane; r < Mr; r += 32) a += cabs2(xj[r]);
    a = warp_sum(a);
    if (lane == 0) sig[j] = sqrt(a);
  }
  for (int j = tid; j < Nc; j += nt) order[j] = j;
  __syncthreads();
  for (int j = tid; j < Nc; j += nt) {
    const double sj = sig[j];
    int rank = 0;
    for (int i = 0; i < Nc; ++i) rank += (sig[i] > sj) || (sig[i] == sj && i < j);
    if (sj == sj) order[rank] = j;
  }
  __syncthreads();
  if (tid == 0 && !converged) atomicOr(status, DQA_ST_NOCONV);
  const int k = Nc;
  for (int i = tid; i < k; i += nt) S[i] = sig[order[i]];
  // left/right factors: Y[:,j]/sig = left vectors of the tall matrix; V = its right vectors
  for (int i = tid; i < m * k; i += nt) {
    const int r = i / k, c = i % k, j = order[c];
    const double sv = sig[j], inv = sv > 0.0 ? 1.0 / sv : 0.0;
    U[i] = tall ? cscale(Y[r + (size_t)j * Mr], inv) : V[r + (size_t)j * Nc];
  }
  for (int i = tid; i < k * n; i += nt) {
    const int r = i / n, c = i % n, j = order[r];
    const double sv = sig[j], inv = sv > 0.0 ? 1.0 / sv : 0.0;
    Vh[i] = tall ? cconj(V[c + (size_t)j * Nc]) : cconj(cscale(Y[c + (size_t)j * Mr], inv));
  }
}

// ---------------------------------------------------------------------------------
// _trim_and_renorm_svd_result (lib/tenn.py:307-368): all six cutoff modes, max_bond, renorm,
// absorb in {-1, 0, 1, 2=None}.  One thread decides n_chi; the block rescales in place.
// U (m x k) row-major, S (k), Vh (k x n) row-major; info[0] = n_chi on exit.
// The caller then reads the leading n_chi columns of U / rows of Vh.
// ---------------------------------------------------------------------------------
__global__ void k_g_trim(cplx* U, double* S, cplx* Vh, int m, int k, int n, double cutoff, int cutoff_mode, int max_bond, int absorb, int renorm, int* info) {
  __shared__ int s_nchi;
  __shared__ double s_scale;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) {
    int n_chi = k;
    double tot = 0.0, keep = 0.0;
    int pw = 2;
    if (cutoff > 0.0 || renorm > 0) {
      if (cutoff_mode == 1) {
        n_chi = 0;
        for (int i = 0; i < k; ++i) n_chi += S[i] > cutoff;
      } else if (cutoff_mode == 2) {
        n_chi = 0;
        for (int i = 0; i < k; ++i) n_chi += S[i] > cutoff * S[0];
      } else {
        pw = (cutoff_mode == 3 || cutoff_mode == 4) ? 2 : 1;
        for (int i = 0; i < k; ++i) tot += pw == 2 ? S[i] * S[i] : S[i];
        double csp = 0.0;
        int cnt = 0;
        for (int i = 0; i < k; ++i) {
          csp += pw == 2 ? S[i] * S[i] : S[i];
          if (cutoff_mode == 4 || cutoff_mode == 6) cnt += csp < (1.0 - cutoff) * tot;
          else cnt += (tot - csp) > cutoff;
        }
        n_chi = cnt + 1;
      }
      if (n_chi < 1) n_chi = 1;
      if (max_bond > 0 && n_chi > max_bond) n_chi = max_bond;
    } else if (max_bond > 0) {
      n_chi = max_bond;
    }
    if (n_chi > k) n_chi = k;
    double sc = 1.0;
    if (n_chi < k && renorm > 0 && cutoff_mode >= 3) {
      for (int i = 0; i < n_chi; ++i) keep += pw == 2 ? S[i] * S[i] : S[i];
      sc = pw == 2 ? sqrt(tot / keep) : tot / keep;
    }
    s_nchi = n_chi;
    s_scale = sc;
    info[0] = n_chi;
  }
  __syncthreads();
  const int n_chi = s_nchi;
  const double sc = s_scale;
  __syncthreads();
  for (int i = tid; i < n_chi; i += nt) S[i] *= sc;
  __syncthreads();
  if (absorb == -1 || absorb == 0) {
    for (int i = tid; i < m * n_chi; i += nt) {
      const int r = i / n_chi, c = i % n_chi;
      const double f = absorb == -1 ? S[c] : sqrt(S[c]);
      U[(size_t)r * k + c] = cscale(U[(size_t)r * k + c], f);
    }
  }
  if (absorb == 1 || absorb == 0) {
    for (int i = tid; i < n_chi * n; i += nt) {
      const int r = i / n;
      const double f = absorb == 1 ? S[r] : sqrt(S[r]);
      Vh[i] = cscale(Vh[i], f);
    }
  }
}

// A_next' = V (nchi x q) . A_next (q, 2*r)   (jnp.tensordot at lib/tenn.py:248); V = leading rows of Vh (ld n)
__global__ void k_g_absorb_next(const cplx* Vh, int nchi, int q, const cplx* Anext, int cols, cplx* out) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)nchi * cols; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols), r = (int)(i / cols);
    cplx acc = cmk(0.0, 0.0);
    for (int j = 0; j < q; ++j) acc = cfma(Vh[(size_t)r * q + j], Anext[(size_t)j * cols + c], acc);
    out[i] = acc;
  }
}

// x /= |x|_F  (lib/tenn.py:175,259), single CTA
__global__ void k_g_normalize(cplx* x, long n) {
  __shared__ double red[64];
  double ss = 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x) ss += cabs2(x[i]);
  ss = block_sum(ss, red);
  const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
  for (long i = threadIdx.x; i < n; i += blockDim.x) x[i] = cscale(x[i], inv);
}
