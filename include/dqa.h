/* dqa.h -- C ABI of libdqa.so, the B200-native engine for the digitized-Quantum-Annealing
 * MPS step of baronefr/perceptron-dqa.
 *
 * The reference has no FFI: its boundary is the Python class `mydQA` (dQA_mps.py:38-185)
 * and the op-level functions of lib/tenn.py.  Each entry point below names the reference
 * interface it stands behind; INTEGRATION.md shows the ctypes binding a maintainer adds to
 * dQA_mps.py to switch the hot path over.
 *
 * Conventions
 *  - C linkage, no exceptions, no C++/torch types.  Every function returns 0 on success,
 *    a negative DQA_E* code for argument/state errors, a positive value = cudaError_t for
 *    CUDA failures; dqa_last_error() gives the message.
 *  - The caller owns all device memory: ask dqa_workspace_bytes(), allocate (e.g. a torch
 *    uint8 tensor), hand the pointer to dqa_bind_workspace().  Nothing is allocated in
 *    dqa_step().  Host buffers are caller-allocated.
 *  - complex128 is interleaved (re, im) doubles, the numpy layout.  MPS site tensors cross
 *    the ABI in the reference's (l, phys=2, r) row-major layout (lib/tenn.py:72-75).
 *  - A handle is bound to one device and one stream and is not thread-safe; calls are
 *    stream-ordered, the get/energy calls synchronise the stream.  dqa_step() may fan the
 *    batch out over streams of its own (sub-batches on parallel branches of the step graph);
 *    they are joined into the handle's stream before the step counts as done there.
 *  - There is no CPU fallback anywhere behind this ABI.
 */
#ifndef DQA_H_
#define DQA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dqa_handle dqa_t;

enum {
  DQA_OK = 0,
  DQA_EINVAL = -1,   /* bad argument */
  DQA_ESTATE = -2,   /* call order (workspace / patterns / tables not set) */
  DQA_ELIMIT = -3,   /* size outside what the kernels support */
  DQA_ENUMERIC = -4  /* some instance went non-finite / zero norm / Jacobi did not converge */
};

/* mydQA.__init__(dataset, P, dt, max_bond) (dQA_mps.py:40-60) plus the truncation policy
 * hard-coded at lib/tenn.py:238 (cutoff=1e-10, cutoff_mode=4 'rsum2', renorm=2, absorb right),
 * exposed as plan fields defaulting to those values, and the ensemble size. */
typedef struct dqa_config {
  int32_t N;           /* sites = pattern length                          */
  int32_t n_xi;        /* patterns per realisation                        */
  int32_t P;           /* annealing steps                                 */
  int32_t chi;         /* max_bond                                        */
  int32_t batch;       /* independent realisations evolved by this handle */
  int32_t device;      /* CUDA device ordinal                             */
  int32_t max_sweeps;  /* Jacobi sweep cap (0 -> 30)                      */
  int32_t n_tables;    /* Fourier tables (distinct dt values) in the batch; 0 -> 1 */
  double dt;           /* dt of every instance (n_tables == 1)            */
  double cutoff;       /* 0 -> 1e-10 (lib/tenn.py:238); <0 -> no cutoff    */
  void* stream;        /* cudaStream_t, NULL = legacy default stream      */
  /* truncation policy of svd_truncated / _trim_and_renorm_svd_result (lib/tenn.py:307-405).  policy_set = 0 keeps
   * the values the reference hard-codes at its call site (lib/tenn.py:238): mode 4, renorm 2, absorb 1. */
  int32_t policy_set;  /* 0: reference defaults; 1: take the three fields below verbatim */
  int32_t cutoff_mode; /* 1 abs, 2 rel, 3 sum2, 4 rsum2, 5 sum1, 6 rsum1 (lib/tenn.py:317-345) */
  int32_t renorm;      /* 0: kept values as they are; >0: rescaled to preserve their p-norm (lib/tenn.py:347-354) */
  int32_t absorb;      /* where S goes: 1 = right, the only value compress_svd_normalized's sweep can use (DQA_ELIMIT otherwise) */
} dqa_config;

/* life cycle ---------------------------------------------------------------------- */
int dqa_create(dqa_t** h, const dqa_config* cfg);
void dqa_destroy(dqa_t* h);
const char* dqa_last_error(const dqa_t* h);
int dqa_workspace_bytes(const dqa_t* h, size_t* bytes);
int dqa_bind_workspace(dqa_t* h, void* device_ptr, size_t bytes);

/* inputs -------------------------------------------------------------------------- */
/* xi: host int8 [batch][n_xi][N], every entry +-1 (self.dataset, dQA_mps.py:46-52). */
int dqa_set_patterns(dqa_t* h, const int8_t* xi);
/* Tables of mydQA.init_fourier (dQA_mps.py:62-66), computed by the caller exactly as the
 * reference does: Uk_FT host complex128 [D][P], fxft host complex128 [D], D = N+1. */
int dqa_set_fourier(dqa_t* h, const double* Uk_FT, const double* fxft);
/* The same for a batch whose instances run different dt (the dt grid of benchmarker-mps.py:50-57 mapped onto the
 * ensemble dimension): Uk_FT host complex128 [n_tables][D][P], one init_fourier table per distinct dt; fxft [D]
 * (independent of dt); table_of_instance [batch]; dt_of_instance [batch] (the mixer angle (1-s_p)*dt, dQA_mps.py:88-89).
 * n_tables must equal dqa_config.n_tables. */
int dqa_set_fourier_tables(dqa_t* h, int n_tables, const double* Uk_FT, const double* fxft, const int32_t* table_of_instance,
                           const double* dt_of_instance);
/* Integer pattern-to-MPO tables, computed ON THE DEVICE, for the bit-exact check
 * (lib/dQA_utils.py:50-53,106): hbit [n_xi][N] u8, e [n_xi][N][2] i32, q [n_xi][N][2][D] i32.
 * Any output pointer may be NULL. */
int dqa_get_index_tensors(dqa_t* h, int inst, uint8_t* hbit, int32_t* e, int32_t* q);

/* the driver path: mydQA.run / single_step (dQA_mps.py:84-174) ------------------------ */
int dqa_reset_plus(dqa_t* h);            /* |+>^N, pp=0, eps[0] = eps(0)               */
int dqa_step(dqa_t* h, int nsteps);      /* nsteps x single_step(); DQA_ENUMERIC if any
                                            instance's status word is non-zero          */
int dqa_step_async(dqa_t* h, int nsteps); /* same, but no host sync / status check (overlap handles on streams) */
int dqa_get_step(const dqa_t* h, int* pp);
int dqa_set_step(dqa_t* h, int pp);

/* sub-steps, for teacher-forced parity ---------------------------- */
/* U_z^mu at step p + right_canonize + compress_svd_normalized (dQA_mps.py:94-103). */
int dqa_apply_pattern(dqa_t* h, int p, int mu);
/* only the right-canonisation half (lib/tenn.py:179-190), leaving the sector factors in the
 * workspace; dqa_truncate then runs the left-to-right truncation (lib/tenn.py:255-262). */
int dqa_canonize(dqa_t* h, int p, int mu);
int dqa_truncate(dqa_t* h, int p, int mu);
/* make_Ux + apply_mpsmpo (lib/dQA_utils.py:26-30, dQA_mps.py:108-109). */
int dqa_apply_ux(dqa_t* h, double beta);
/* mydQA.compute_loss on the current state (dQA_mps.py:68-81): eps host complex128 [batch]. */
int dqa_energy(dqa_t* h, double* eps);

/* state access: checkpoint / resume / teacher forcing -------------------------------- */
int dqa_get_bonds(dqa_t* h, int inst, int32_t* bonds /* [N+1] */);
int dqa_get_mps(dqa_t* h, int inst, int site, double* buf /* l*2*r complex */, int* l, int* r);
int dqa_set_mps(dqa_t* h, int inst, int site, const double* buf, int l, int r);
/* whole-batch state in the device layout: [batch][N][chi*2*chi] complex128 (site (a,s,c) at
 * (a*2+s)*chi+c) followed by [batch][N+1] int32 bond dimensions.  For checkpoint/resume and
 * for the end-to-end benchmark (pinned host buffers). */
int dqa_state_bytes(const dqa_t* h, size_t* bytes);
int dqa_download_state(dqa_t* h, void* host);
int dqa_upload_state(dqa_t* h, const void* host);
int dqa_sync(dqa_t* h);
/* kernels launched by this handle so far */
int dqa_get_launch_count(const dqa_t* h, int64_t* n);
/* singular values seen at cut `bond` (between sites bond and bond+1) by the last truncation. */
int dqa_get_schmidt(dqa_t* h, int inst, int bond, double* s /* [2*chi] */, int* n);

/* traces: self.loss / self.bd_monitor (dQA_mps.py:105-106,116,147) ------------------- */
int dqa_get_eps(dqa_t* h, double* eps /* [batch][P+1] complex */);
int dqa_get_bd_monitor(dqa_t* h, int32_t* bd /* [batch][P*n_xi] */);
int dqa_get_status(dqa_t* h, int32_t* status /* [batch] */);
/* status words are sticky until a new state is loaded (dqa_reset_plus / dqa_set_mps / dqa_upload_state) or cleared here */
int dqa_clear_status(dqa_t* h);
/* relative discarded weight (sum of dropped s^2 / sum of all s^2) of the last truncation at every cut: w [N-1]
 * (what _trim_and_renorm_svd_result, lib/tenn.py:347-354, renormalises away) */
int dqa_get_truncation_error(dqa_t* h, int inst, double* w);
/* [batch][16] u64: Jacobi rotations, Jacobi sweeps, SVDs, QR columns, and the FP64 flops executed by
 * the Householder LQ, the Jacobi iteration, the sector QRs and the theta GEMMs (exact loop counts),
 * then thread-0 clock cycles per phase and CTA counts of k_sector_qr (8..11) and k_lq (12..15). */
int dqa_get_counters(dqa_t* h, uint64_t* counters, int reset);

/* ensemble moments for the multi-GPU reduction: writes to DEVICE memory (caller-owned,
 * 3*(P+1) doubles): sum_r Re eps_r(s), sum_r (Re eps_r(s))^2, number of healthy instances;
 * the caller all-reduces that buffer (one ncclAllReduce of 24 KB at P=1000). */
int dqa_ensemble_moments(dqa_t* h, double* device_out);

/* profiling: per kernel-family device time, measured with CUDA events on the handle's
 * stream.  families: 0 sector_qr, 1 theta, 2 lq, 3 mixer, 4 energy, 5 jacobi. */
#define DQA_NFAM 6
int dqa_profile_begin(dqa_t* h);
int dqa_profile_end(dqa_t* h, double* ms /* [DQA_NFAM] */, int64_t* launches /* [DQA_NFAM] */);

/* FP64 FMA throughput of this device (TFLOP/s), measured by a register-resident DFMA loop;
 * the roofline denominator for the FP64-bound kernels (MEASURED_PEAKS.json has no FP64 entry). */
int dqa_measure_fp64_peak(dqa_t* h, double* tflops);
/* the same on the FP64 tensor path (mma.sync m8n8k4 f64 loop) */
int dqa_measure_dmma_peak(dqa_t* h, double* tflops);

const char* dqa_version(void);

/* ---- op-level API: the reference's lib/tenn.py functions on dense site tensors --------------------
 * Host buffers in and out (complex128 interleaved, row-major, the reference's array layouts); the device
 * scratch is caller-owned and handed over once (dqa_ops_create).  An op that needs more scratch than the
 * context has returns DQA_ELIMIT and says how much.  The driver's hot loop does not use these. */
typedef struct dqa_ops dqa_ops_t;
int dqa_ops_create(dqa_ops_t** ctx, void* device_scratch, size_t bytes, int device, void* stream);
void dqa_ops_destroy(dqa_ops_t* ctx);
const char* dqa_ops_last_error(const dqa_ops_t* ctx);
/* contract_mps_mpo_single_site (lib/tenn.py:67-81): A (al,2,ar) x W (wl,wr,2,2) -> out (al*wl, 2, ar*wr) */
int dqa_op_contract_site(dqa_ops_t* ctx, const double* A, int al, int ar, const double* W, int wl, int wr, double* out);
/* braket(bra, ket) (lib/tenn.py:93-109; conjugates ket): packed site tensors, bond arrays of n+1 entries -> env (rb x rk) */
int dqa_op_braket(dqa_ops_t* ctx, int n, const int32_t* bra_bonds, const double* bra, const int32_t* ket_bonds, const double* ket, double* env);
/* qr_stabilized(x) (lib/tenn.py:294-304), sgn(0):=+1: X (m x n) -> Q (m x min), R (min x n) */
int dqa_op_qr_stabilized(dqa_ops_t* ctx, const double* X, int m, int n, double* Q, double* R);
/* right_canonize_twosites(A, B) (lib/tenn.py:152-176): A (al,2,bl), B (bl,2,br) -> A' (al,2,kq) normalised, B' (kq,2,br) */
int dqa_op_right_canonize_twosites(dqa_ops_t* ctx, const double* A, int al, const double* B, int bl, int br, double* Anew, double* Bnew, int* kq);
/* svd_truncated (lib/tenn.py:307-405): cutoff modes 1..6, max_bond (<=0: none), absorb -1/0/1/2(None), renorm; buffers sized for
 * k = min(m,n); the leading n_chi columns of U (ld k) / rows of Vh are the result */
int dqa_op_svd_truncated(dqa_ops_t* ctx, const double* X, int m, int n, double cutoff, int cutoff_mode, int max_bond, int absorb, int renorm,
                         double* U, double* S, double* Vh, int* n_chi);
/* compress_normalize_onesite(A, A_next, max_bd) (lib/tenn.py:229-252): A (l,2,r), A_next (r,2,r2) -> A' (l,2,n_chi), A_next' (n_chi,2,r2) */
int dqa_op_compress_onesite(dqa_ops_t* ctx, const double* A, int l, int r, const double* Anext, int r2, int max_bd, double* Aout,
                            double* Anext_out, int* n_chi);
/* x /= |x|_F (lib/tenn.py:259) */
int dqa_op_normalize(dqa_ops_t* ctx, double* x, long n);
/* compute_loss of an external packed MPS with n_sites - n_phys trailing ancilla sites (lib/dQA_loss.py:28-78, dQA_mps.py:68-81) */
int dqa_op_energy(dqa_ops_t* ctx, int n_sites, int n_phys, int n_xi, const int32_t* bonds, const double* mps, const int8_t* xi,
                  const double* fxft, double* eps);

/* Exact (untruncated) state-vector dQA on the GPU -- the gauge-free validator of SURVEY 8f-3, same conventions as the MPS
 * path (site 0 = most significant bit, s=0 <-> sigma=+1): nsteps Trotter steps from |+>^N, eps[0..nsteps].  Scratch: 16*2^N B.
 * ms (optional, 3 doubles): device milliseconds in the diagonal-phase, mixer and energy kernels (all HBM-streaming). */
int dqa_sv_run(dqa_ops_t* ctx, int N, int n_xi, const int8_t* xi, int P, double dt, int nsteps, double* eps, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* DQA_H_ */
