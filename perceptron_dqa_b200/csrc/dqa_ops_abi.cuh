// dqa_ops_abi.cuh -- C ABI of the op-level functions (include/dqa.h, "op-level API" block).
// Host buffers in, host buffers out; the device scratch is caller-owned (dqa_ops_create).
#pragma once
#include "dqa_ops.cuh"
#include "dqa_sv.cuh"

#include <complex>
#include <vector>

struct dqa_ops {
  cudaStream_t stream;
  char* ws;
  size_t ws_bytes;
  int device;
  char err[384];
};

static int ofail(dqa_ops* h, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h->err, sizeof(h->err), fmt, ap);
  va_end(ap);
  return code;
}
#define OCK(call)                                                                                     \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) return ofail(h, (int)e_, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

struct OCarve {
  dqa_ops* h;
  size_t off;
  bool ok;
  template <class T>
  T* take(size_t n) {
    const size_t a = (off + 255) & ~(size_t)255;
    if (a + n * sizeof(T) > h->ws_bytes) {
      ok = false;
      off = a + n * sizeof(T);
      return nullptr;
    }
    off = a + n * sizeof(T);
    return reinterpret_cast<T*>(h->ws + a);
  }
};
#define OCARVE_CHECK(cv) \
  if (!cv.ok) return ofail(h, DQA_ELIMIT, "op needs %zu bytes of scratch, context has %zu", cv.off, h->ws_bytes)

extern "C" int dqa_ops_create(dqa_ops_t** out, void* device_scratch, size_t bytes, int device, void* stream) {
  if (!out || !device_scratch) return DQA_EINVAL;
  dqa_ops* h = new dqa_ops();
  h->stream = (cudaStream_t)stream;
  h->ws = (char*)device_scratch;
  h->ws_bytes = bytes;
  h->device = device;
  h->err[0] = 0;
  *out = h;
  OCK(cudaSetDevice(device));
  OCK(cudaFuncSetAttribute(k_energy_ext<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  OCK(cudaFuncSetAttribute(k_energy_ext<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return DQA_OK;
}
extern "C" void dqa_ops_destroy(dqa_ops_t* h) { delete h; }
extern "C" const char* dqa_ops_last_error(const dqa_ops_t* h) { return h ? h->err : "null ops context"; }

extern "C" int dqa_op_contract_site(dqa_ops_t* h, const double* A, int al, int ar, const double* W, int wl, int wr, double* out) {
  if (!h || !A || !W || !out || al < 1 || ar < 1 || wl < 1 || wr < 1) return DQA_EINVAL;
  const size_t na = (size_t)al * 2 * ar, nw = (size_t)wl * wr * 4, no = (size_t)al * wl * 2 * ar * wr;
  OCarve cv{h, 0, true};
  cplx* dA = cv.take<cplx>(na);
  cplx* dW = cv.take<cplx>(nw);
  cplx* dO = cv.take<cplx>(no);
  OCARVE_CHECK(cv);
  OCK(cudaMemcpyAsync(dA, A, na * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dW, W, nw * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  const int blocks = (int)((no + 255) / 256 < 2048 ? (no + 255) / 256 : 2048);
  DQA_LAUNCH(k_g_contract_site, dim3(blocks), dim3(256), 0, h->stream, (const cplx*)dA, al, ar, (const cplx*)dW, wl, wr, dO);
  OCK(cudaGetLastError());
  OCK(cudaMemcpyAsync(out, dO, no * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

// braket(bra, ket) over n sites (lib/tenn.py:93-109); bonds arrays have n+1 entries; flat = packed (l,2,r) tensors
extern "C" int dqa_op_braket(dqa_ops_t* h, int n, const int32_t* bbond, const double* bra, const int32_t* kbond, const double* ket, double* env_out) {
  if (!h || n < 1 || !bbond || !bra || !kbond || !ket || !env_out) return DQA_EINVAL;
  size_t nb = 0, nk = 0, maxenv = 1, maxtmp = 1;
  for (int i = 0; i < n; ++i) {
    nb += (size_t)bbond[i] * 2 * bbond[i + 1];
    nk += (size_t)kbond[i] * 2 * kbond[i + 1];
    maxenv = std::max(maxenv, (size_t)bbond[i + 1] * kbond[i + 1]);
    maxtmp = std::max(maxtmp, (size_t)kbond[i] * 2 * bbond[i + 1]);
  }
  OCarve cv{h, 0, true};
  cplx* dB = cv.take<cplx>(nb);
  cplx* dK = cv.take<cplx>(nk);
  cplx* e0 = cv.take<cplx>(maxenv);
  cplx* e1 = cv.take<cplx>(maxenv);
  cplx* tmp = cv.take<cplx>(maxtmp);
  OCARVE_CHECK(cv);
  OCK(cudaMemcpyAsync(dB, bra, nb * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dK, ket, nk * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  size_t ob = 0, ok = 0;
  cplx* cur = nullptr;
  for (int i = 0; i < n; ++i) {
    cplx* nxt = (cur == e0) ? e1 : e0;
    DQA_LAUNCH(k_g_braket_site, dim3(1), dim3(256), 0, h->stream, (const cplx*)cur, (const cplx*)(dB + ob), (int)bbond[i], (int)bbond[i + 1],
               (const cplx*)(dK + ok), (int)kbond[i], (int)kbond[i + 1], tmp, nxt);
    ob += (size_t)bbond[i] * 2 * bbond[i + 1];
    ok += (size_t)kbond[i] * 2 * kbond[i + 1];
    cur = nxt;
  }
  OCK(cudaGetLastError());
  OCK(cudaMemcpyAsync(env_out, cur, (size_t)bbond[n] * kbond[n] * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

extern "C" int dqa_op_qr_stabilized(dqa_ops_t* h, const double* X, int m, int n, double* Q, double* R) {
  if (!h || !X || !Q || !R || m < 1 || n < 1) return DQA_EINVAL;
  const int kq = m < n ? m : n;
  OCarve cv{h, 0, true};
  cplx* dX = cv.take<cplx>((size_t)m * n);
  cplx* dQ = cv.take<cplx>((size_t)m * kq);
  cplx* dR = cv.take<cplx>((size_t)kq * n);
  cplx* wk = cv.take<cplx>((size_t)m * n + (size_t)m * kq + 2 * (size_t)kq + 8);
  OCARVE_CHECK(cv);
  OCK(cudaMemcpyAsync(dX, X, (size_t)m * n * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  DQA_LAUNCH(k_g_qr_stabilized, dim3(1), dim3(OPS_THREADS), 0, h->stream, (const cplx*)dX, m, n, dQ, dR, wk);
  OCK(cudaGetLastError());
  OCK(cudaMemcpyAsync(Q, dQ, (size_t)m * kq * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaMemcpyAsync(R, dR, (size_t)kq * n * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

extern "C" int dqa_op_right_canonize_twosites(dqa_ops_t* h, const double* A, int al, const double* B, int bl, int br, double* Anew, double* Bnew,
                                              int* kq_out) {
  if (!h || !A || !B || !Anew || !Bnew || !kq_out || al < 1 || bl < 1 || br < 1) return DQA_EINVAL;
  const int m = 2 * br, n = bl, kq = m < n ? m : n;
  OCarve cv{h, 0, true};
  cplx* dA = cv.take<cplx>((size_t)al * 2 * bl);
  cplx* dB = cv.take<cplx>((size_t)bl * 2 * br);
  cplx* dAn = cv.take<cplx>((size_t)al * 2 * kq);
  cplx* dBn = cv.take<cplx>((size_t)kq * 2 * br);
  cplx* wk = cv.take<cplx>((size_t)m * n + (size_t)m * kq + 2 * (size_t)kq + 8);
  OCARVE_CHECK(cv);
  OCK(cudaMemcpyAsync(dA, A, (size_t)al * 2 * bl * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dB, B, (size_t)bl * 2 * br * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  DQA_LAUNCH(k_g_right_canonize_twosites, dim3(1), dim3(OPS_THREADS), 0, h->stream, (const cplx*)dA, al, (const cplx*)dB, bl, br, dAn, dBn, wk);
  OCK(cudaGetLastError());
  OCK(cudaMemcpyAsync(Anew, dAn, (size_t)al * 2 * kq * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaMemcpyAsync(Bnew, dBn, (size_t)kq * 2 * br * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  *kq_out = kq;
  return DQA_OK;
}

// svd_truncated(x, cutoff, cutoff_mode, max_bond, absorb, renorm) (lib/tenn.py:372-405): U (m x k), S (k), Vh (k x n) are
// written in full (k = min(m,n)); the leading n_chi columns/rows are the reference's return value.  absorb: -1,0,1 or 2 = None.
extern "C" int dqa_op_svd_truncated(dqa_ops_t* h, const double* X, int m, int n, double cutoff, int cutoff_mode, int max_bond, int absorb,
                                    int renorm, double* U, double* S, double* Vh, int* n_chi) {
  if (!h || !X || !U || !S || !Vh || !n_chi || m < 1 || n < 1) return DQA_EINVAL;
  if (cutoff_mode < 1 || cutoff_mode > 6) return ofail(h, DQA_EINVAL, "cutoff_mode must be 1..6");
  const int k = m < n ? m : n, Mr = m >= n ? m : n;
  OCarve cv{h, 0, true};
  cplx* dX = cv.take<cplx>((size_t)m * n);
  cplx* dU = cv.take<cplx>((size_t)m * k);
  double* dS = cv.take<double>(k);
  cplx* dV = cv.take<cplx>((size_t)k * n);
  int* info = cv.take<int>(4);
  cplx* wk = cv.take<cplx>((size_t)Mr * k + (size_t)k * k + 2 * (size_t)k + 8);
  OCARVE_CHECK(cv);
  OCK(cudaMemcpyAsync(dX, X, (size_t)m * n * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemsetAsync(info, 0, 4 * sizeof(int), h->stream));
  DQA_LAUNCH(k_g_svd, dim3(1), dim3(OPS_THREADS), 0, h->stream, (const cplx*)dX, m, n, dU, dS, dV, wk, 60, info + 1);
  DQA_LAUNCH(k_g_trim, dim3(1), dim3(256), 0, h->stream, dU, dS, dV, m, k, n, cutoff, cutoff_mode, max_bond, absorb, renorm, info);
  OCK(cudaGetLastError());
  int hinfo[4];
  OCK(cudaMemcpyAsync(U, dU, (size_t)m * k * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaMemcpyAsync(S, dS, (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaMemcpyAsync(Vh, dV, (size_t)k * n * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaMemcpyAsync(hinfo, info, sizeof(hinfo), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  *n_chi = hinfo[0];
  if (hinfo[1]) return ofail(h, DQA_ENUMERIC, "Jacobi SVD did not converge");
  return DQA_OK;
}

// compress_normalize_onesite(A, A_next, max_bd) (lib/tenn.py:229-252): policy cutoff=1e-10, mode 4, renorm 2, absorb right
extern "C" int dqa_op_compress_onesite(dqa_ops_t* h, const double* A, int l, int r, const double* Anext, int r2, int max_bd, double* Aout,
                                       double* Anext_out, int* n_chi) {
  if (!h || !A || !Anext || !Aout || !Anext_out || !n_chi || l < 1 || r < 1 || r2 < 1) return DQA_EINVAL;
  const int m = 2 * l, n = r, k = m < n ? m : n, Mr = m >= n ? m : n, cols = 2 * r2;
  OCarve cv{h, 0, true};
  cplx* dX = cv.take<cplx>((size_t)m * n);
  cplx* dN = cv.take<cplx>((size_t)r * cols);
  cplx* dU = cv.take<cplx>((size_t)m * k);
  double* dS = cv.take<double>(k);
  cplx* dV = cv.take<cplx>((size_t)k * n);
  cplx* dO = cv.take<cplx>((size_t)k * cols);
  int* info = cv.take<int>(4);
  cplx* wk = cv.take<cplx>((size_t)Mr * k + (size_t)k * k + 2 * (size_t)k + 8);
  OCARVE_CHECK(cv);
  OCK(cudaMemcpyAsync(dX, A, (size_t)m * n * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dN, Anext, (size_t)r * cols * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemsetAsync(info, 0, 4 * sizeof(int), h->stream));
  DQA_LAUNCH(k_g_svd, dim3(1), dim3(OPS_THREADS), 0, h->stream, (const cplx*)dX, m, n, dU, dS, dV, wk, 60, info + 1);
  DQA_LAUNCH(k_g_trim, dim3(1), dim3(256), 0, h->stream, dU, dS, dV, m, k, n, 1e-10, 4, max_bd, 1, 2, info);
  int hinfo[4];
  OCK(cudaMemcpyAsync(hinfo, info, sizeof(hinfo), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  const int nc = hinfo[0];
  const int blocks = (int)(((size_t)nc * cols + 255) / 256);
  DQA_LAUNCH(k_g_absorb_next, dim3(blocks < 1024 ? blocks : 1024), dim3(256), 0, h->stream, (const cplx*)dV, nc, n, (const cplx*)dN, cols, dO);
  OCK(cudaGetLastError());
  // A' = U[:, :nc] reshaped (l,2,nc): strided copy out of the (m x k) buffer
  std::vector<std::complex<double>> hu((size_t)m * k);
  OCK(cudaMemcpyAsync(hu.data(), dU, (size_t)m * k * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaMemcpyAsync(Anext_out, dO, (size_t)nc * cols * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  std::complex<double>* ao = reinterpret_cast<std::complex<double>*>(Aout);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < nc; ++j) ao[(size_t)i * nc + j] = hu[(size_t)i * k + j];
  *n_chi = nc;
  if (hinfo[1]) return ofail(h, DQA_ENUMERIC, "Jacobi SVD did not converge");
  return DQA_OK;
}

extern "C" int dqa_op_normalize(dqa_ops_t* h, double* x, long n) {
  if (!h || !x || n < 1) return DQA_EINVAL;
  OCarve cv{h, 0, true};
  cplx* dX = cv.take<cplx>((size_t)n);
  OCARVE_CHECK(cv);
  OCK(cudaMemcpyAsync(dX, x, (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  DQA_LAUNCH(k_g_normalize, dim3(1), dim3(256), 0, h->stream, dX, n);
  OCK(cudaGetLastError());
  OCK(cudaMemcpyAsync(x, dX, (size_t)n * sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

// mydQA_ancilla.compute_loss (lib/dQA_loss.py:67-78) / mydQA.compute_loss (dQA_mps.py:68-81) for an external packed MPS:
// n_sites >= n_phys sites (the trailing n_sites-n_phys ancilla sites carry the identity), xi int8 [n_xi][n_phys], fxft [n_phys+1].
extern "C" int dqa_op_energy(dqa_ops_t* h, int n_sites, int n_phys, int n_xi, const int32_t* bonds, const double* mps, const int8_t* xi,
                             const double* fxft, double* eps) {
  if (!h || !bonds || !mps || !xi || !fxft || !eps || n_phys < 2 || n_sites < n_phys || n_xi < 1) return DQA_EINVAL;
  typedef std::complex<double> cd;
  const int D = n_phys + 1, kh = D / 2 + 1;
  int chi = 1;
  std::vector<long> off(n_sites);
  size_t tot = 0;
  for (int i = 0; i < n_sites; ++i) {
    off[i] = (long)tot;
    tot += (size_t)bonds[i] * 2 * bonds[i + 1];
    chi = std::max(chi, (int)std::max(bonds[i], bonds[i + 1]));
  }
  if (bonds[0] != 1 || bonds[n_sites] != 1) return ofail(h, DQA_EINVAL, "boundary bonds must be 1");
  if (energy_smem(chi) > 200 * 1024) return ofail(h, DQA_ELIMIT, "bond %d too large for the energy kernel (max 64)", chi);
  std::vector<uint8_t> hb((size_t)n_xi * n_phys);
  for (size_t i = 0; i < hb.size(); ++i) {
    if (xi[i] != 1 && xi[i] != -1) return ofail(h, DQA_EINVAL, "pattern entry %zu is not +-1", i);
    hb[i] = (uint8_t)((1 - xi[i]) / 2);
  }
  std::vector<cd> gk(D), ph(D);
  const cd* fx = reinterpret_cast<const cd*>(fxft);
  for (int k = 0; k < D; ++k) {
    gk[k] = fx[k] / std::sqrt((double)D);
    const double a2 = (M_PI / D) * k * 2.0;
    ph[k] = cd(std::cos(a2), std::sin(a2));
  }
  OCarve cv{h, 0, true};
  cplx* dM = cv.take<cplx>(tot);
  long* dOff = cv.take<long>(n_sites);
  int* dBond = cv.take<int>(n_sites + 1);
  uint8_t* dHb = cv.take<uint8_t>(hb.size());
  cplx* dGk = cv.take<cplx>(D);
  cplx* dPh = cv.take<cplx>(D);
  cplx* dT = cv.take<cplx>((size_t)n_xi * kh);
  cplx* dE = cv.take<cplx>(1);
  OCARVE_CHECK(cv);
  OCK(cudaMemcpyAsync(dM, mps, tot * sizeof(cplx), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dOff, off.data(), sizeof(long) * n_sites, cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dBond, bonds, sizeof(int) * (n_sites + 1), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dHb, hb.data(), hb.size(), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dGk, gk.data(), sizeof(cd) * D, cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dPh, ph.data(), sizeof(cd) * D, cudaMemcpyHostToDevice, h->stream));
  if (chi > 16)
    DQA_LAUNCH(k_energy_ext<true>, dim3(kh, n_xi), dim3(256), energy_smem(chi), h->stream, (const cplx*)dM, (const long*)dOff,
               (const int*)dBond, n_sites, n_phys, chi, (const uint8_t*)dHb, (const cplx*)dPh, kh, dT);
  else
    DQA_LAUNCH(k_energy_ext<false>, dim3(kh, n_xi), dim3(chi <= 8 ? 64 : 256), energy_smem(chi), h->stream, (const cplx*)dM,
               (const long*)dOff, (const int*)dBond, n_sites, n_phys, chi, (const uint8_t*)dHb, (const cplx*)dPh, kh, dT);
  DQA_LAUNCH(k_energy_ext_reduce, dim3(1), dim3(32), 0, h->stream, (const cplx*)dT, (const cplx*)dGk, n_xi, kh, D, n_phys, dE);
  OCK(cudaGetLastError());
  OCK(cudaMemcpyAsync(eps, dE, sizeof(cplx), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  return DQA_OK;
}

// Exact state-vector dQA: runs `nsteps` Trotter steps from |+>^N on the device and returns
// eps(s) for s = 0, 1/P, ..., nsteps/P (nsteps+1 doubles).  Needs 16*2^N bytes of scratch; N <= 40 here.
// ms_out (optional, 3 doubles): device time spent in the phase / mixer / energy kernels.
extern "C" int dqa_sv_run(dqa_ops_t* h, int N, int n_xi, const int8_t* xi, int P, double dt, int nsteps, double* eps, double* ms_out) {
  if (!h || !xi || !eps || N < 2 || N > 40 || n_xi < 1 || n_xi > 64 || P < 1 || nsteps < 0 || nsteps > P) return DQA_EINVAL;
  const unsigned long long dim = 1ull << N;
  std::vector<unsigned long long> pat(n_xi, 0ull);
  for (int mu = 0; mu < n_xi; ++mu)
    for (int i = 0; i < N; ++i) {
      const int v = xi[(size_t)mu * N + i];
      if (v != 1 && v != -1) return ofail(h, DQA_EINVAL, "pattern entry is not +-1");
      if (v == -1) pat[mu] |= 1ull << (N - 1 - i);
    }
  std::vector<double> fx(N + 1);
  for (int x = 0; x <= N; ++x) {
    const int m = N - 2 * x;  // f(x) = h(N-2x), h(m) = 0 for m >= 0 else -m/sqrt(N)  (lib/dQA_utils.py:78-95)
    fx[x] = m >= 0 ? 0.0 : -(double)m / std::sqrt((double)N);
  }
#ifdef DQA_EMUL
  const int nblk = 2;
#else
  const int nblk = 148 * 8;
#endif
  OCarve cv{h, 0, true};
  cplx* psi = cv.take<cplx>(dim);
  unsigned long long* dpat = cv.take<unsigned long long>(n_xi);
  double* dfx = cv.take<double>(N + 1);
  double* part = cv.take<double>(nblk);
  double* dout = cv.take<double>(nsteps + 1);
  OCARVE_CHECK(cv);
  OCK(cudaMemcpyAsync(dpat, pat.data(), sizeof(unsigned long long) * n_xi, cudaMemcpyHostToDevice, h->stream));
  OCK(cudaMemcpyAsync(dfx, fx.data(), sizeof(double) * (N + 1), cudaMemcpyHostToDevice, h->stream));
  OCK(cudaFuncSetAttribute(k_sv_rx, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  cudaEvent_t ev[4];
  float acc_ms[3] = {0.f, 0.f, 0.f};
  for (auto& e : ev) OCK(cudaEventCreate(&e));
  DQA_LAUNCH(k_sv_init, dim3(nblk), dim3(256), 0, h->stream, psi, dim, std::pow(2.0, -0.5 * N));
  DQA_LAUNCH(k_sv_energy, dim3(nblk), dim3(256), 0, h->stream, (const cplx*)psi, dim, (const unsigned long long*)dpat, n_xi, (const double*)dfx, N + 1, part);
  DQA_LAUNCH(k_sv_energy_final, dim3(1), dim3(256), 0, h->stream, (const double*)part, nblk, 1.0 / N, dout);
  for (int p = 0; p < nsteps; ++p) {
    const double sp = (double)(p + 1) / P, gamma = sp * dt, beta = (1.0 - sp) * dt;
    OCK(cudaEventRecord(ev[0], h->stream));
    DQA_LAUNCH(k_sv_phase, dim3(nblk), dim3(256), 0, h->stream, psi, dim, (const unsigned long long*)dpat, n_xi, (const double*)dfx, N + 1, gamma);
    OCK(cudaEventRecord(ev[1], h->stream));
    // mixer: low 10 bits in one contiguous-tile pass, then groups of up to 7 bits with 32-wide coalesced rows
    int q0 = 0;
    while (q0 < N) {
      const int lowbits = q0 == 0 ? 0 : (q0 < 5 ? q0 : 5);
      int nq = q0 == 0 ? 10 : 7;
      if (nq > N - q0) nq = N - q0;
      const size_t sm = sizeof(cplx) * ((size_t)1 << (nq + lowbits));
      const unsigned long long ntiles = 1ull << (N - nq - lowbits);
      DQA_LAUNCH(k_sv_rx, dim3((unsigned)(ntiles < 148 * 16 ? ntiles : 148 * 16)), dim3(256), sm, h->stream, psi, N, q0, nq, lowbits, std::cos(beta), std::sin(beta));
      q0 += nq;
    }
    OCK(cudaEventRecord(ev[2], h->stream));
    DQA_LAUNCH(k_sv_energy, dim3(nblk), dim3(256), 0, h->stream, (const cplx*)psi, dim, (const unsigned long long*)dpat, n_xi, (const double*)dfx, N + 1, part);
    DQA_LAUNCH(k_sv_energy_final, dim3(1), dim3(256), 0, h->stream, (const double*)part, nblk, 1.0 / N, dout + p + 1);
    OCK(cudaEventRecord(ev[3], h->stream));
    if (ms_out) {
      OCK(cudaEventSynchronize(ev[3]));
      for (int j = 0; j < 3; ++j) {
        float t = 0.f;
        OCK(cudaEventElapsedTime(&t, ev[j], ev[j + 1]));
        acc_ms[j] += t;
      }
    }
  }
  OCK(cudaGetLastError());
  OCK(cudaMemcpyAsync(eps, dout, sizeof(double) * (nsteps + 1), cudaMemcpyDeviceToHost, h->stream));
  OCK(cudaStreamSynchronize(h->stream));
  for (auto& e : ev) cudaEventDestroy(e);
  if (ms_out)
    for (int j = 0; j < 3; ++j) ms_out[j] = acc_ms[j];
  return DQA_OK;
}
