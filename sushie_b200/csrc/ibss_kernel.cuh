// ibss_kernel.cuh -- the persistent, device-resident IBSS loop for sm_100a.
//
// One *team* of thread blocks (a thread-block cluster with a hardware barrier, or a
// co-resident group with a global-memory barrier) owns one locus from start to
// convergence; teams pull loci from a device-side queue.  Inside a locus the whole
// reference loop runs on device:
//
//   reference                                    here
//   ------------------------------------------   -----------------------------------------
//   infer.py:498-548 host loop + ELBO sync       iteration loop + on-device stop rule
//   infer.py:618 lax.fori_loop over L            effect loop
//   infer.py:640,653,676 three einsums over X    ONE streamed traversal of X per effect
//                                                (traverse(): X b_new, residual update and
//                                                 X^T r_next from the same shared-memory band)
//   infer.py:692-745 _compute_posterior (x2)     posterior passes A / B1 / B2 (per-SNP k x k
//                                                closed form, online log-sum-exp softmax)
//   infer.py:128-159 _EMOptFunc/_NoopOptFunc     prior update between pass A and pass B
//   infer.py:786-808 _eloglike/_erss             cached norms, no extra pass over X
//   infer_ss.py:437-575                          same kernel, SS=true: rows are LD rows,
//                                                no X^T r sweep, ERSS from cached vectors
//
// Genotype / LD bands are brought into shared memory with bulk asynchronous copies
// (cp.async.bulk + mbarrier, i.e. the TMA engine, SASS UBLKCP) through a multi-stage ring
// that keeps prefetching across the posterior phases, because the traversal order of X is
// static.  All reductions are fixed-order (no floating-point atomics), so results are
// bit-reproducible for a given team size.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include <type_traits>

#ifndef SB_I8_KEEP
#define SB_I8_KEEP 1
#endif
#ifndef SB_I8_BLOCKS_PER_SM
#define SB_I8_BLOCKS_PER_SM 1
#endif

namespace sb {

constexpr int KMAX = 6;  // ancestries per fit (the kernels are instantiated per count, 1..KMAX)
constexpr int NT = 256;  // threads per block
constexpr int NW = NT / 32;
constexpr int NT_B2 = 256;  // threads of the bit-plane kernel (512 measured slower: 230 vs 204 us per update at p = 2000)
constexpr int NW_MAX = NT_B2 / 32;
constexpr int GMAX = 32;  // max rows per band
constexpr int NSTAGE_MAX = 6;
constexpr int I8_STATE_ROWS = 2048;  // hard-call path: band rows whose (r, xb) state is staged in shared memory at a time
constexpr int SS_CHUNK_ROWS = 256;  // summary-level register-tile path: rows whose partial sums are parked in shared memory
constexpr int NSYM_MAX = KMAX * (KMAX + 1) / 2;
constexpr int FAST_GLOBAL_VECTORS = 9;  // ProbDev::fast: generic traversal with b / accumulator in global memory
// bit-plane hard-call storage (XT = uint8_t, "B2"): genotypes 0/1/2 as two bit planes (g >= 1, g == 2), stored twice:
// row-major bits for the X b sweep and column-major bits for the X^T r sweep (see traverse_b2)
constexpr int B2_TILE_COLS = 256;       // columns per chunk / tile = 32 byte-groups of 8 columns
constexpr int B2_STAGE_BYTES = 32768;   // one ring stage of the bit stream
constexpr int B2_NSTAGE = 2;             // one stage below the table sets, one above (shared-memory alignment)
constexpr int B2_ROUND_ROWS = 512;      // rows per sweep-1 round of one block (64 per warp)
constexpr int B2_TABLE_BYTES = 65536;   // 256 entries x 32 groups x 8 bytes; two sets, aligned to 64 KB in shared memory
constexpr int B2_MAX_ROWS = 1024;       // rows per ancestry (their next residual is staged in shared memory)
constexpr int B2_MAX_SUB = 4;           // column tiles of one block per sweep-2 piece: 8 x B2_MAX_SUB >= p / 256
constexpr int B2_QMAX = 8;              // columns of the covariate basis Q (intercept included) the bit-plane path handles
constexpr int STAT_STRIDE = 32;  // doubles per team-reduction record (>= 2 + NSYM_MAX + 2 + KMAX)
constexpr int TRAV_STRIDE = 2 * KMAX;

struct OutBuf {
  double* alpha;  // [L][p]
  double* pm;     // [L][p][K]
  double* pmsq;   // [L][p][K][K]
  double* lbf;    // [L][p]
  double* wsc;    // [L][K][K]
  double* kl;     // [L]
  double* ecov;   // [L][K][K]
  double* rv;     // [K]
  double* xtxb2;  // [L][K]   sum_p XtX * E[b^2]  (ERSS term)
  double* lse;    // [L][2]   softmax normaliser of each effect: (max, sum exp(w - max))
};

struct ProbDev {
  int K, p, L, p_pad;
  int n[KMAX];         // rows per ancestry (summary level: p)
  int row0[KMAX + 1];  // offsets into the concatenated row space
  int band0[KMAX + 1]; // band index offsets per ancestry
  int G, nstage, nseg; // rows per band, ring stages, warps per row in sweep 1
  int em;              // 1 = EM prior update, 0 = times/plus adjustor
  int skip_first_pass; // adjustor.times == 0: the first posterior pass cannot matter
  int fast;            // traversal variant: 0 generic; register tile 1: 4 vec x 4 rows, 2: 8 x 2, 3: 12 x 1, 4: 16 x 1 (summary level)
  const void* X[KMAX]; // [n_k][p_pad] storage type
  double nscale[KMAX]; // summary level: n_k (XtX = n LD); individual level: 1
  double nsamp[KMAX];  // n_k as double for the ELBO
  const double* y;     // [nrows]   (summary level: Xty [K][p] concatenated)
  double* logpi;       // [p]
  const double* pi_raw; // [p] user prior weights, nullptr = uniform 1/p (only read by the prologue)
  const double* xtx;   // [K][p_pad] diagonal of XtX
  const double* ecov0; // [L][K][K]
  const double* rv0;   // [K]
  const double* adj_times;  // [K][K]
  const double* adj_plus;   // [K][K]
  double* r;      // [nrows] current residual (with the current effect added back)
  double* xb;     // [L][nrows] cached X b_l  (summary level: XtX b_l)
  double* xty;    // [K][p_pad] reduced X^T r                    (individual level)
  double* rfull;  // [nrows] full residual at iteration end      (summary level)
  const double* cmean;  // [K][p_pad] column means      (hard-call int8 storage only; pad = 0)
  const double* cisd;   // [K][p_pad] 1 / column std    (hard-call int8 storage only; pad = 0)
  // bit-plane storage: rbits[k] = [chunk][row quad][plane][32 groups][4 rows] bytes, bit c = column 8 g + c;
  //                    cbits[k] = [row chunk][column quad][plane][32 row groups][4 columns] bytes, bit r = row 8 rg + r
  const unsigned char* rbits[KMAX];
  const unsigned char* cbits[KMAX];
  int ntiles;           // ceil(p / 256)
  int pad_b2_;
  // genotypes residualised on covariates (infer.py:319-335) without leaving hard-call storage: with Q the
  // orthonormal basis of [1, covariates] and M = G^T Q,  X_std b = (I - Q Q^T) G b'  and
  // X_std^T r = (G^T r - M Q^T r) / sd  -- rank-(c+1) corrections on the n- and p-vectors.  qc = 0: no covariates
  // (centring through cmean).
  const double* Qb[KMAX];  // [n_k][qc]
  const double* Mq[KMAX];  // [qc][p_pad]
  int qc[KMAX];
  double* bvec;   // [L][K][p_pad] posterior-mean effects b_l = alpha_l * mu_l of the current iteration (pad = 0)
  OutBuf out[2];
  double* elbo;   // [max_iter + 1]
  int* meta;      // [0]=n_iter [1]=elbo_increase [2]=final buffer (-1: initial state) [3]=n_ser
                  // [4]=status: 0 not started, 1 suspended at an iteration boundary, 2 finished
                  // [5]=1 when the run stopped on a NaN ELBO (reported as sb_result.status)
};

struct LaunchParams {
  const ProbDev* probs;
  const int* prob_ids;  // problems of this launch (same K / mode / storage)
  int n_ids;
  int* queue;          // device work queue head
  int* team_slot;      // [n_teams]
  unsigned* team_bar;  // [n_teams][64] software barrier words
  int team_size;
  int hw_cluster;
  int max_iter;
  double min_tol;
  // per-block scratch (index = blockIdx.x): [part: K*p_pad_max][3 x STAT_STRIDE][trav: L_max*TRAV_STRIDE]
  double* scratch;
  long long scratch_stride;  // doubles per block
  int off_stat, off_trav;    // offsets (doubles) inside a block's scratch
  // phase control: ctl[0] = teams that found the queue empty, ctl[1] = drain request
  int* ctl;
  int n_teams;
  int allow_drain;  // a later phase with larger teams exists: suspend loci when half the teams idle
};

// ---------------------------------------------------------------------------------------
// PTX helpers

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// Bulk asynchronous global->shared copy (TMA engine), completion counted on an mbarrier.  The genotype /
// LD stream is read once per single-effect update and the per-SM working set (one locus per block) is far
// larger than L2, so it is tagged evict-first: the small per-locus vectors (residual, cached X b, effect
// vectors, partial X^T r, priors) then stay L2-resident instead of being flushed every ~20 us.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

struct Team {
  int size, rank, id;
  int hw;
  unsigned* bar;  // software barrier: bar[0] = arrivals, bar[32] = generation
};

__device__ __forceinline__ void team_sync(const Team& t) {
  if (t.size == 1) {
    __syncthreads();
    return;
  }
  if (t.hw) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    return;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned gen = ld_acquire_u32(t.bar + 32);
    __threadfence();
    const unsigned prev = atomicAdd(t.bar, 1u);
    if (prev == static_cast<unsigned>(t.size) - 1u) {
      t.bar[0] = 0u;
      __threadfence();
      st_release_u32(t.bar + 32, gen + 1u);
    } else {
      while (ld_acquire_u32(t.bar + 32) == gen) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// storage-type helpers: one 16-byte shared-memory vector of genotypes -> doubles

template <typename XT>
struct Vec;
template <>
struct Vec<double> {
  static constexpr int W = 2;
  __device__ static __forceinline__ void load(const double* base, int v, double (&o)[2]) {
    const double2 t = reinterpret_cast<const double2*>(base)[v];
    o[0] = t.x;
    o[1] = t.y;
  }
  __device__ static __forceinline__ void load_raw(const double* base, int v, double (&o)[2]) { load(base, v, o); }
};
template <>
struct Vec<float> {
  static constexpr int W = 4;
  __device__ static __forceinline__ void load(const float* base, int v, double (&o)[4]) {
    const float4 t = reinterpret_cast<const float4*>(base)[v];
    o[0] = t.x;
    o[1] = t.y;
    o[2] = t.z;
    o[3] = t.w;
  }
  // storage-typed copy: the register tile keeps fp32 genotypes as fp32 and widens at the FMA
  __device__ static __forceinline__ void load_raw(const float* base, int v, float (&o)[4]) {
    const float4 t = reinterpret_cast<const float4*>(base)[v];
    o[0] = t.x;
    o[1] = t.y;
    o[2] = t.z;
    o[3] = t.w;
  }
};

template <>
struct Vec<uint8_t> {  // bit-plane hard-call storage has its own traversal (traverse_b2)
  static constexpr int W = 16;
  template <typename A, typename B>
  __device__ static void load(A, int, B&) {}
  template <typename A, typename B>
  __device__ static void load_raw(A, int, B&) {}
};
template <>
struct Vec<int8_t> {  // hard-call storage has its own traversal (traverse_i8); only the width is shared
  static constexpr int W = 16;
  // name-lookup stubs for the dense traversals, which are never instantiated for this storage type
  template <typename A, typename B>
  __device__ static void load(A, int, B&) {}
  template <typename A, typename B>
  __device__ static void load_raw(A, int, B&) {}
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}


// Hard-call genotype bytes (0..255, in practice 0/1/2) -> double without the conversion pipe: byte `B` of
// `w` is zero-extended (prmt) into the low word under the exponent of 2^52, then 2^52 is subtracted.
// Two instructions, exact.
template <int B>
__device__ __forceinline__ double u8_to_f64(uint32_t w) {
  constexpr uint32_t sel = 0x4440u | B;  // bytes 1..3 <- byte 0 of the zero operand
  uint32_t lo;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(lo) : "r"(w), "r"(0u), "r"(sel));
  return __hiloint2double(0x43300000, static_cast<int>(lo)) - 4503599627370496.0;  // 2^52
}

// Transposed butterfly: every thread brings GR partial sums; afterwards lane l holds the warp total
// of row l / (32 / GR).  GR + log2(32 / GR) - 1 double shuffles instead of 5 * GR.
template <int GR>
__device__ __forceinline__ double warp_rows_sum(double (&ps)[GR], int lane) {
  static_assert(GR == 1 || GR == 2 || GR == 4 || GR == 8 || GR == 16 || GR == 32, "rows per band");
  int off = 16;
#pragma unroll
  for (int cnt = GR; cnt > 1; cnt >>= 1) {
    const bool hi = lane & off;
#pragma unroll
    for (int i = 0; i < cnt / 2; ++i) {
      const double keep = hi ? ps[i + cnt / 2] : ps[i];
      const double send = hi ? ps[i] : ps[i + cnt / 2];
      ps[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  double t = ps[0];
#pragma unroll
  for (; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
  return t;
}

// Shared-memory header (fixed part of the dynamic allocation).
struct SmemHdr {
  uint64_t mbar[NSTAGE_MAX];  // "stage filled" (bulk-copy completion)
  uint64_t ebar[NSTAGE_MAX];  // "stage drained" (one arrival per warp; summary-level register-tile path)
  ProbDev pb;                // the locus being processed
  alignas(16) double partsum[2][4 * GMAX];  // sweep-1 partial dot products [band parity][g * nseg + seg]
  double rn[GMAX];           // next residual of the band's rows
  double red[NW_MAX][STAT_STRIDE];
  double tot[STAT_STRIDE];   // block / team totals
  double cst[2 * KMAX * KMAX + 8];  // Cinv, misc per-pass constants
  double trow[GMAX][2];      // per-row scalars of a traversal
  int decision;
  int pad_[3];
  unsigned drain[NSTAGE_MAX];  // hard-call row-team traversal: warps that have finished with the band in stage s (mod NW)
  unsigned pad2_[2];
  double qv[2 * 8];          // bit-plane traversal: centring terms of the current ancestry (Q^T G b' | Q^T r_next)
  double tsum[NW_MAX];       // per row team: sum of the next residual over the rows it processed (current ancestry)
};

// Block-wide sum of NV values per thread; totals land in H->tot[0..NV) (valid for all
// threads after return).  Fixed order => deterministic.
template <int NV, int NWARPS = NW>
__device__ __forceinline__ void block_sum(SmemHdr* H, double (&v)[NV]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();  // protect red/tot from the previous use
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) H->red[w][i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int ww = 0; ww < NWARPS; ++ww) s += H->red[ww][threadIdx.x];
    H->tot[threadIdx.x] = s;
  }
  __syncthreads();
}

template <int NWARPS = NW>
__device__ __forceinline__ double block_max(SmemHdr* H, double v) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) H->red[w][0] = v;
  __syncthreads();
  double m = H->red[0][0];
#pragma unroll
  for (int ww = 1; ww < NWARPS; ++ww) m = fmax(m, H->red[ww][0]);
  return m;  // red is protected by the leading barrier of the next block_* call
}

// ---------------------------------------------------------------------------------------
// k x k algebra (compile-time K, fully unrolled => registers)

// General inverse + log|det| by Gauss-Jordan with partial pivoting (prior covariance; the
// reference uses LU-based inv/slogdet, infer.py:704,764,778).  Run by one thread per pass.
template <int K>
__device__ void inv_logdet(const double* __restrict__ C, double* __restrict__ Cinv, double& logabsdet) {
  double a[K][2 * K];
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j) {
      a[i][j] = C[i * K + j];
      a[i][K + j] = (i == j) ? 1.0 : 0.0;
    }
  double lad = 0.0;
  for (int c = 0; c < K; ++c) {
    int piv = c;
    double best = fabs(a[c][c]);
    for (int r = c + 1; r < K; ++r)
      if (fabs(a[r][c]) > best) {
        best = fabs(a[r][c]);
        piv = r;
      }
    if (piv != c)
      for (int j = 0; j < 2 * K; ++j) {
        const double t = a[c][j];
        a[c][j] = a[piv][j];
        a[piv][j] = t;
      }
    const double d = a[c][c];
    lad += log(fabs(d));
    const double inv = 1.0 / d;
    for (int j = 0; j < 2 * K; ++j) a[c][j] *= inv;
    for (int r = 0; r < K; ++r)
      if (r != c) {
        const double f = a[r][c];
        for (int j = 0; j < 2 * K; ++j) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j) Cinv[i * K + j] = a[i][K + j];
  logabsdet = lad;
}

// Per-SNP posterior (infer.py:704-719): S = (diag(d) + Cinv)^-1, mu = S r,
// logdetA = log|diag(d) + Cinv| = -log|S|.  LDL^T without square roots; a non-positive
// pivot (A not PD) yields NaN like the reference's Cholesky-based logpdf.
template <int K>
__device__ __forceinline__ void snp_posterior(const double (&Cinv)[K][K], const double (&d)[K],
                                              const double (&r)[K], double (&S)[K][K], double (&mu)[K],
                                              double& logdetA) {
  double Lm[K][K];  // unit lower factor (strict lower part used)
  double D[K], Dinv[K];
  double det = 1.0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    double dj = Cinv[j][j] + d[j];
#pragma unroll
    for (int t = 0; t < j; ++t) dj -= Lm[j][t] * Lm[j][t] * D[t];
    if (!(dj > 0.0)) dj = CUDART_NAN;
    D[j] = dj;
    Dinv[j] = 1.0 / dj;
    det *= dj;
#pragma unroll
    for (int i = j + 1; i < K; ++i) {
      double s = Cinv[i][j];
#pragma unroll
      for (int t = 0; t < j; ++t) s -= Lm[i][t] * Lm[j][t] * D[t];
      Lm[i][j] = s * Dinv[j];
    }
  }
  logdetA = log(det);
  // Li = Lm^-1 (unit lower)
  double Li[K][K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
#pragma unroll
    for (int j = 0; j < K; ++j) Li[i][j] = (i == j) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
#pragma unroll
    for (int i = j + 1; i < K; ++i) {
      double s = 0.0;
#pragma unroll
      for (int t = j; t < i; ++t) s -= Lm[i][t] * Li[t][j];
      Li[i][j] = s;
    }
  }
  // S = Li^T D^-1 Li
#pragma unroll
  for (int a = 0; a < K; ++a) {
#pragma unroll
    for (int b = a; b < K; ++b) {
      double s = 0.0;
#pragma unroll
      for (int t = b; t < K; ++t) s += Li[t][a] * Dinv[t] * Li[t][b];
      S[a][b] = s;
      S[b][a] = s;
    }
  }
#pragma unroll
  for (int a = 0; a < K; ++a) {
    double s = 0.0;
#pragma unroll
    for (int b = 0; b < K; ++b) s += S[a][b] * r[b];
    mu[a] = s;
  }
}

constexpr double LOG_2PI = 1.8378770664093454835606594728112;

// ---------------------------------------------------------------------------------------
// the kernel

template <int N>
using IntC = std::integral_constant<int, N>;

// V: traversal variant fixed at compile time (hard-call storage: 1 = 8 columns x 8 rows per thread, 2 = 16 x 4,
// 3 = 32 x 2, 4 = 4 x 8 for narrow loci, so each variant gets its own register allocation); 0 = chosen per
// locus at run time.
template <int K, typename XT, bool SS, int V = 0>
__global__ void __launch_bounds__((std::is_same<XT, uint8_t>::value ? NT_B2 : NT),
                                  (std::is_same<XT, int8_t>::value ? SB_I8_BLOCKS_PER_SM : 1)) ibss_kernel(const LaunchParams P) {
  // threads / warps per block of THIS instantiation (shadow the namespace constants inside the kernel body)
  constexpr int NT = std::is_same<XT, uint8_t>::value ? sb::NT_B2 : sb::NT;
  constexpr int NW = NT / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int NSYM = K * (K + 1) / 2;
  constexpr int VW = Vec<XT>::W;
  constexpr bool B2 = std::is_same<XT, uint8_t>::value;  // bit-plane storage: X^T r is complete per block
  SmemHdr* H = reinterpret_cast<SmemHdr*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  Team team;
  team.size = P.team_size;
  team.id = blockIdx.x / P.team_size;
  team.rank = blockIdx.x % P.team_size;
  team.hw = P.hw_cluster;
  team.bar = P.team_bar + static_cast<size_t>(team.id) * 64;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE_MAX; ++s) {
      mbar_init(&H->mbar[s], 1);
      mbar_init(&H->ebar[s], NW);
      H->drain[s] = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const uint64_t stream_policy = l2_evict_first_policy();
  uint32_t phase_bits = 0;   // bit s = mbarrier parity of the next fill to consume in stage s
  uint32_t ephase_bits = 0;  // bit s = parity of the next "drained" phase of stage s (producer thread only)

  for (;;) {
    // ---- pull the next locus from the device queue
    if (team.rank == 0 && tid == 0) {
      const int v = atomicAdd(P.queue, 1);
      __stcg(&P.team_slot[team.id], v);
    }
    team_sync(team);
    const int qi = __ldcg(&P.team_slot[team.id]);
    if (qi >= P.n_ids) {
      if (team.rank == 0 && tid == 0 && P.allow_drain) {
        const int idle = atomicAdd(&P.ctl[0], 1) + 1;
        if (2 * idle >= P.n_teams) atomicExch(&P.ctl[1], 1);
      }
      break;
    }
    {
      const int* src = reinterpret_cast<const int*>(&P.probs[P.prob_ids[qi]]);
      int* dst = reinterpret_cast<int*>(&H->pb);
      for (int i = tid; i < static_cast<int>(sizeof(ProbDev) / sizeof(int)); i += NT) dst[i] = src[i];
    }
    __syncthreads();
    const ProbDev& pb = H->pb;

    const int p = pb.p, p_pad = pb.p_pad, L = pb.L;
    const int G = pb.G, NS = pb.nstage, NSEG = pb.nseg;
    const int NB = pb.band0[K];
    const int nrows = pb.row0[K];
    const int C = team.size, rank = team.rank;
    const int nb_own = (NB > rank) ? (NB - rank + C - 1) / C : 0;  // bands owned by this block
    const int pv = p_pad / VW;                                       // 16-byte vectors per row
    const size_t row_bytes = static_cast<size_t>(p_pad) * sizeof(XT);
    const size_t stage_bytes = row_bytes * G;
    const int fast = pb.fast;  // register-tile traversal variant (0 = generic shared-memory path)
    // scratch of team member c (partials, reduction records, per-traversal scalars)
    auto scr = [&](int c) -> double* {
      return P.scratch + static_cast<size_t>(team.id * team.size + c) * static_cast<size_t>(P.scratch_stride);
    };
    double* const my_scr = scr(team.rank);
    const int status0 = __ldcg(&pb.meta[4]);  // 0 fresh, 1 resumed from a suspended state

    // dynamic shared-memory carve-up: [header][b (generic path)][acc (generic, individual)][ring]
    double* bsm_sm = reinterpret_cast<double*>(smem_raw + ((sizeof(SmemHdr) + 127) / 128) * 128);
    double* acc_sm = bsm_sm + p_pad;
    unsigned char* ring =
        fast ? reinterpret_cast<unsigned char*>(bsm_sm)
             : reinterpret_cast<unsigned char*>(SS ? (bsm_sm + p_pad) : (acc_sm + p_pad));

    // SNP slice of this block for the posterior phases
    const int js = static_cast<int>((static_cast<long long>(p) * rank) / C);
    const int je = static_cast<int>((static_cast<long long>(p) * (rank + 1)) / C);

    // band decode: own position -> (ancestry, first row, rows)
    auto band_of = [&](int pos, int& k, int& g0, int& rows) {
      const int q = rank + pos * C;
      k = 0;
      int b0 = 0, nk = pb.n[0];
#pragma unroll
      for (int kk = 1; kk < K; ++kk)
        if (q >= pb.band0[kk]) {
          k = kk;
          b0 = pb.band0[kk];
          nk = pb.n[kk];
        }
      g0 = (q - b0) * G;
      rows = min(G, nk - g0);
    };
    // producer: one bulk copy of own band `pos` (rows are contiguous) into ring stage `st`
    auto issue_fill = [&](int pos, int st) {
      int k, g0, rows;
      band_of(pos, k, g0, rows);
      const unsigned char* src =
          reinterpret_cast<const unsigned char*>(pb.X[k]) + static_cast<size_t>(g0) * row_bytes;
      unsigned char* dst = ring + static_cast<size_t>(st) * stage_bytes;
      const uint32_t bytes = static_cast<uint32_t>(row_bytes * rows);
      mbar_expect_tx(&H->mbar[st], bytes);
      bulk_g2s(dst, src, bytes, &H->mbar[st], stream_policy);
    };

    // ---- locus initialisation (state owned by this block: rows of its bands)
    for (int pos = 0; status0 == 0 && pos < nb_own; ++pos) {
      int k, g0, rows;
      band_of(pos, k, g0, rows);
      for (int t = tid; t < rows; t += NT) {
        const int row = pb.row0[k] + g0 + t;
        pb.r[row] = pb.y[row];
        for (int l = 0; l < L; ++l) pb.xb[static_cast<size_t>(l) * nrows + row] = 0.0;
      }
    }

    // ring prologue: keep NS fills in flight; positions wrap (the traversal order is static)
    // the thread that issues bulk copies (owns prod_pos): lane 0 of the last warp in the register-tile
    // traversals, thread 0 in the generic one (also when its vectors live in global memory)
    const int producer_tid = (fast && fast != FAST_GLOBAL_VECTORS) ? NT - 32 : 0;
    int prod_pos = 0;  // next own band position to issue (producer thread only)
    int cons_seq = 0;  // bands consumed so far in this locus (stage = cons_seq % NS)
    __syncthreads();
    if (tid == producer_tid && nb_own > 0) {
      for (int s = 0; s < NS; ++s) {
        issue_fill(prod_pos, s);
        prod_pos = (prod_pos + 1 == nb_own) ? 0 : prod_pos + 1;
      }
    }

    // b_l = softmax(w) * mu (infer.py:721-723) is materialised once per single-effect update by the
    // team (end of posterior()); the traversals read it back with coalesced loads.
    auto effect_vec = [&](int l, int k) -> const double* {
      return pb.bvec + (static_cast<size_t>(l) * K + k) * p_pad;
    };

    // per-traversal scalars: fixed-order serial sum over the designated threads
    auto store_trav = [&](int l, bool first, const double (&vsq)[K], const double (&rsq)[K]) {
      __syncthreads();
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (tid < GMAX) {
          H->trow[tid][0] = vsq[k];
          H->trow[tid][1] = rsq[k];
        }
        __syncthreads();
        if (tid == 0 && !first) {
          double a = 0.0, b = 0.0;
          for (int g = 0; g < GMAX; ++g) {
            a += H->trow[g][0];
            b += H->trow[g][1];
          }
          double* dst = my_scr + P.off_trav + l * TRAV_STRIDE;
          __stcg(&dst[k], a);
          __stcg(&dst[KMAX + k], b);
        }
        __syncthreads();
      }
    };
    auto zero_unowned_partials = [&]() {
      for (int k = 0; k < K; ++k) {
        const int b0 = pb.band0[k], b1 = pb.band0[k + 1];
        const int first_q = b0 + (((rank - b0) % C) + C) % C;  // first band index == rank (mod C)
        if (!(first_q < b1)) {
          double* dst = my_scr + static_cast<size_t>(k) * p_pad;
          for (int j = tid; j < p_pad; j += NT) __stcg(&dst[j], 0.0);
        }
      }
    };

    // ------------------------------------------------------------------------------
    // one streamed traversal of X (individual) / LD (summary) for effect l:
    //   v = X b_l ; xb[l] = v ; R = r - v ; r_next = R + xb[lnext] ; acc += X^T r_next
    // first == true (individual level only): b = 0, produces X^T y.
    //
    // Register-tile variant: every thread owns CPT 16-byte column vectors and keeps them for
    // all GR rows of the band in registers, so the band is read from shared memory once and
    // both sweeps run from registers; b and the X^T r accumulators live in registers too.
    auto traverse_fast = [&](auto cpt_c, auto gr_c, int l, bool first, bool last_effect) {
      constexpr int CPT = decltype(cpt_c)::value;
      constexpr int GR = decltype(gr_c)::value;
      constexpr int PRODUCER_TID = NT - 32;  // lane 0 of the last warp issues the bulk copies
      const int lnext = first ? 0 : ((l + 1 == L) ? 0 : l + 1);
      const bool self_next = !first && lnext == l;
      double vsq[K], rsq[K];  // meaningful in threads < GR only
#pragma unroll
      for (int k = 0; k < K; ++k) vsq[k] = rsq[k] = 0.0;
      double bq[CPT][VW], aq[CPT][VW];
      XT x[GR][CPT][VW];  // the band slice of this thread (storage type); columns past the row end stay 0
      uint32_t vmask = 0;     // bit c: this thread owns a c-th column vector
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if (tid + c * NT < pv) vmask |= (1u << c);
#pragma unroll
        for (int u = 0; u < VW; ++u) {
          bq[c][u] = aq[c][u] = 0.0;
#pragma unroll
          for (int g = 0; g < GR; ++g) x[g][c][u] = XT(0);
        }
      }
      int cur_k = -1;
      int st = cons_seq % NS;
      auto flush_acc = [&](int k) {
        double* dst = my_scr + static_cast<size_t>(k) * p_pad;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          if (vmask & (1u << c)) {
            const int v = tid + c * NT;
#pragma unroll
            for (int u = 0; u < VW; u += 2)
              __stcg(reinterpret_cast<double2*>(dst + v * VW + u), make_double2(aq[c][u], aq[c][u + 1]));
          }
        }
      };

      for (int pos = 0; pos < nb_own; ++pos) {
        int k, g0, rows;
        band_of(pos, k, g0, rows);
        if (k != cur_k) {
          if (!SS && cur_k >= 0) flush_acc(cur_k);
          const double* bl = effect_vec(l, k);
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const int v = tid + c * NT;
            const bool on = !first && (vmask & (1u << c));
#pragma unroll
            for (int u = 0; u < VW; u += 2) {
              const double2 t2 = on ? __ldcg(reinterpret_cast<const double2*>(bl + v * VW + u)) : make_double2(0.0, 0.0);
              aq[c][u] = aq[c][u + 1] = 0.0;
              bq[c][u] = t2.x;
              bq[c][u + 1] = t2.y;
            }
          }
          cur_k = k;
        }
        const XT* tile = reinterpret_cast<const XT*>(ring + static_cast<size_t>(st) * stage_bytes);
        const int row_base = pb.row0[k] + g0;
        const double nsc = pb.nscale[k];
        const bool full_rows = (rows == GR);

        // per-row state (same addresses for every thread: broadcast loads), in flight while the band lands
        double r_old[GR], xb_next[GR];
        {
          const double* rp = pb.r + row_base;
          const double* xp = pb.xb + static_cast<size_t>(lnext) * nrows + row_base;
#pragma unroll
          for (int g = 0; g < GR; ++g) {
            const bool ok = full_rows || g < rows;
            r_old[g] = ok ? rp[g] : 0.0;
            xb_next[g] = ok ? xp[g] : 0.0;
          }
        }
        mbar_wait(&H->mbar[st], (phase_bits >> st) & 1u);

        // the band -> registers (the only shared-memory read of the genotypes)
        if (full_rows) {
#pragma unroll
          for (int g = 0; g < GR; ++g)
#pragma unroll
            for (int c = 0; c < CPT; ++c)
              if (vmask & (1u << c)) Vec<XT>::load_raw(tile + static_cast<size_t>(g) * p_pad, tid + c * NT, x[g][c]);
        } else {
#pragma unroll
          for (int g = 0; g < GR; ++g)
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              if (g < rows) {
                if (vmask & (1u << c)) Vec<XT>::load_raw(tile + static_cast<size_t>(g) * p_pad, tid + c * NT, x[g][c]);
              } else {
#pragma unroll
                for (int u = 0; u < VW; ++u) x[g][c][u] = XT(0);
              }
            }
        }

        // sweep 1: per-thread partial dot products, transposed butterfly reduction over the warp
        const int pbuf = cons_seq & 1;
        if (!first) {
          double ps[GR];
#pragma unroll
          for (int g = 0; g < GR; ++g) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int c = 0; c < CPT; ++c)
#pragma unroll
              for (int u = 0; u < VW; u += 2) {
                s0 = fma(static_cast<double>(x[g][c][u]), bq[c][u], s0);
                s1 = fma(static_cast<double>(x[g][c][u + 1]), bq[c][u + 1], s1);
              }
            ps[g] = s0 + s1;
          }
          // after the butterfly, lanes with (lane / (32 / GR)) == g hold the warp total of row g
          double t;
          if constexpr (GR == 4) {
            const bool hi = lane & 16;
            double k0 = hi ? ps[2] : ps[0], k1 = hi ? ps[3] : ps[1];
            const double s0 = hi ? ps[0] : ps[2], s1 = hi ? ps[1] : ps[3];
            k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            const bool hi2 = lane & 8;
            t = hi2 ? k1 : k0;
            const double s = hi2 ? k0 : k1;
            t += __shfl_xor_sync(0xffffffffu, s, 8);
            t += __shfl_xor_sync(0xffffffffu, t, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
          } else if constexpr (GR == 2) {
            const bool hi = lane & 16;
            t = hi ? ps[1] : ps[0];
            const double s = hi ? ps[0] : ps[1];
            t += __shfl_xor_sync(0xffffffffu, s, 16);
            t += __shfl_xor_sync(0xffffffffu, t, 8);
            t += __shfl_xor_sync(0xffffffffu, t, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
          } else if constexpr (GR == 8) {
            t = warp_rows_sum<8>(ps, lane);
          } else {
            t = warp_sum(ps[0]);
          }
          if ((lane & (32 / GR - 1)) == 0) H->partsum[pbuf][(lane / (32 / GR)) * NW + warp] = t;
        }
        __syncthreads();  // (A) partial sums visible; every thread holds its tile in registers
        if (tid == PRODUCER_TID) {  // the stage is free again: refill it with the band NS positions ahead
          issue_fill(prod_pos, st);
          prod_pos = (prod_pos + 1 == nb_own) ? 0 : prod_pos + 1;
        }
        // every thread finishes the row sums itself (no second barrier); fixed pairwise order
        double rn[GR];
#pragma unroll
        for (int g = 0; g < GR; ++g) {
          double v = 0.0;
          if (!first) {
            const double2* ps2 = reinterpret_cast<const double2*>(&H->partsum[pbuf][g * NW]);
            static_assert(NW == 8 || B2, "pairwise tree below assumes 8 warps");
            const double2 q0 = ps2[0], q1 = ps2[1], q2 = ps2[2], q3 = ps2[3];
            v = (((q0.x + q0.y) + (q1.x + q1.y)) + ((q2.x + q2.y) + (q3.x + q3.y))) * nsc;
          }
          const double Rfull = r_old[g] - v;
          rn[g] = Rfull + (self_next ? v : xb_next[g]);  // L == 1: the effect added back is the one just computed
          if (tid == g && g < rows) {  // designated thread of row g: publish the row state
            const int row = row_base + g;
            if (!first) pb.xb[static_cast<size_t>(l) * nrows + row] = v;
            pb.r[row] = rn[g];
            double vterm = v * v;  // |X b_l|^2 (infer.py:806)
            if (SS) {
              vterm = __ldcg(&effect_vec(l, k)[g0 + g]) * v;  // b_l^T XtX b_l (infer_ss.py:572)
              if (last_effect) pb.rfull[row] = Rfull;
            }
#pragma unroll
            for (int kk = 0; kk < K; ++kk)
              if (kk == k) {
                vsq[kk] += vterm;
                rsq[kk] += Rfull * Rfull;  // |y - X sum_l b_l|^2 at the last effect (infer.py:804)
              }
          }
        }
        // sweep 2: X^T r_next accumulated in registers
        if (!SS) {
#pragma unroll
          for (int g = 0; g < GR; ++g)
#pragma unroll
            for (int c = 0; c < CPT; ++c)
#pragma unroll
              for (int u = 0; u < VW; ++u) aq[c][u] = fma(static_cast<double>(x[g][c][u]), rn[g], aq[c][u]);
        }
        phase_bits ^= (1u << st);
        cons_seq += 1;
        st = (st + 1 == NS) ? 0 : st + 1;
      }
      if (!SS) {
        if (cur_k >= 0) flush_acc(cur_k);
        zero_unowned_partials();
      }
      store_trav(l, first, vsq, rsq);
    };

    // Hard-call variant (XT = int8_t): the band holds raw genotype bytes (1 byte per entry: an
    // eighth of the fp64 traffic); centring / scaling is folded into the vectors,
    //   X_std b = G (b / sd) - sum_j mean_j b_j / sd_j,   X_std^T r = (G^T r - mean * sum_i r_i) / sd,
    // so the traversal is G-only.  Every thread owns CPT 8-byte column vectors, widens its slice of
    // the GR band rows to fp64 once (integer splice + one DADD per entry) and runs both sweeps from
    // registers.
    auto traverse_i8 = [&](auto cpt_c, auto gr_c, auto keep_c, auto vb_c, int l, bool first, bool last_effect) {
      constexpr int CPT = decltype(cpt_c)::value;
      constexpr int VB = decltype(vb_c)::value;  // genotype bytes (= columns) per vector: 8, or 4 for narrow loci
      constexpr int VWD = VB / 4;                // 32-bit words per vector
      constexpr bool KEEP = decltype(keep_c)::value != 0;  // widen once and keep the fp64 tile in registers
      constexpr int GR = decltype(gr_c)::value;
      constexpr int PRODUCER_TID = NT - 32;
      constexpr int LG = 32 / GR;  // lanes sharing one band row after the butterfly
      const int pv8 = p_pad / VB;  // vectors per row
      const int lnext = first ? 0 : ((l + 1 == L) ? 0 : l + 1);
      const bool self_next = !first && lnext == l;
      double vsq[K], rsq[K];  // meaningful in the row-writer lanes of warp 0 only
#pragma unroll
      for (int k = 0; k < K; ++k) vsq[k] = rsq[k] = 0.0;
      double bq[CPT][VB], aq[CPT][VB];
      uint32_t x[GR][CPT][VWD];  // raw genotype bytes of the band slice (1 register per 4 entries)
      double xk[KEEP ? GR : 1][KEEP ? CPT : 1][VB];  // KEEP: the widened tile, shared by both sweeps
      uint32_t vmask = 0;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if (tid + c * NT < pv8) vmask |= (1u << c);
#pragma unroll
        for (int u = 0; u < VB; ++u) {
          bq[c][u] = aq[c][u] = 0.0;
        }
      }
      int cur_k = -1;
      int st = cons_seq % NS;
      double cshift = 0.0;  // sum_j mean_j b_j / sd_j of the current ancestry
      double rsum = 0.0;    // sum of r_next over the rows of the current ancestry seen by this block
      // row state (r, xb[lnext]) of up to I8_STATE_ROWS band rows is staged in shared memory ahead of the
      // bands that use it: a band lasts well under one L2 round trip here, so per-band global loads would
      // sit on the critical path of the block barrier
      double* rs_sm = reinterpret_cast<double*>(ring + static_cast<size_t>(NS) * stage_bytes);
      double* xs_sm = rs_sm + I8_STATE_ROWS;
      constexpr int CHUNK_BANDS = I8_STATE_ROWS / GR;
      auto stage_rows = [&](int pos0, int pos1) {
        __syncthreads();  // previous chunk fully consumed
        for (int i = tid; i < (pos1 - pos0) * GR; i += NT) {
          int k, g0, rows;
          band_of(pos0 + i / GR, k, g0, rows);
          const int g = i % GR;
          double rv = 0.0, xv = 0.0;
          if (g < rows) {
            const int row = pb.row0[k] + g0 + g;
            rv = pb.r[row];
            xv = pb.xb[static_cast<size_t>(lnext) * nrows + row];
          }
          rs_sm[i] = rv;
          xs_sm[i] = xv;
        }
        __syncthreads();
      };

      // partial X_std^T r of ancestry k: (G^T r - mean * sum r) / sd
      auto flush_acc = [&](int k) {
        double* dst = my_scr + static_cast<size_t>(k) * p_pad;
        const double* cm = pb.cmean + static_cast<size_t>(k) * p_pad;
        const double* cs = pb.cisd + static_cast<size_t>(k) * p_pad;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          if (vmask & (1u << c)) {
            const int j0 = (tid + c * NT) * VB;
#pragma unroll
            for (int u = 0; u < VB; u += 2) {
              const double2 m2 = *reinterpret_cast<const double2*>(cm + j0 + u);
              const double2 s2 = *reinterpret_cast<const double2*>(cs + j0 + u);
              __stcg(reinterpret_cast<double2*>(dst + j0 + u),
                     make_double2((aq[c][u] - m2.x * rsum) * s2.x, (aq[c][u + 1] - m2.y * rsum) * s2.y));
            }
          }
        }
      };

      int chunk0 = 0;
      for (int pos = 0; pos < nb_own; ++pos) {
        if (pos == 0 || pos == chunk0 + CHUNK_BANDS) {
          chunk0 = pos;
          stage_rows(chunk0, min(nb_own, chunk0 + CHUNK_BANDS));
        }
        int k, g0, rows;
        band_of(pos, k, g0, rows);
        if (k != cur_k) {
          if (cur_k >= 0) flush_acc(cur_k);
          const double* bl = effect_vec(l, k);
          const double* cm = pb.cmean + static_cast<size_t>(k) * p_pad;
          const double* cs = pb.cisd + static_cast<size_t>(k) * p_pad;
          double part[1] = {0.0};
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const int j0 = (tid + c * NT) * VB;
            const bool on = !first && (vmask & (1u << c));
#pragma unroll
            for (int u = 0; u < VB; u += 2) {
              double2 b2 = make_double2(0.0, 0.0);
              if (on) {
                const double2 t2 = __ldcg(reinterpret_cast<const double2*>(bl + j0 + u));
                const double2 s2 = *reinterpret_cast<const double2*>(cs + j0 + u);
                const double2 m2 = *reinterpret_cast<const double2*>(cm + j0 + u);
                b2 = make_double2(t2.x * s2.x, t2.y * s2.y);
                part[0] += m2.x * b2.x + m2.y * b2.y;
              }
              aq[c][u] = aq[c][u + 1] = 0.0;
              bq[c][u] = b2.x;
              bq[c][u + 1] = b2.y;
            }
          }
          block_sum<1, NW>(H, part);
          cshift = H->tot[0];
          rsum = 0.0;
          cur_k = k;
        }
        const unsigned char* tile = ring + static_cast<size_t>(st) * stage_bytes;
        const int row_base = pb.row0[k] + g0;
        mbar_wait(&H->mbar[st], (phase_bits >> st) & 1u);

        // the band -> registers (the only shared-memory read of the genotypes)
#pragma unroll
        for (int g = 0; g < GR; ++g)
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
#pragma unroll
            for (int w = 0; w < VWD; ++w) x[g][c][w] = 0u;
            if (g < rows && (vmask & (1u << c))) {
              const unsigned char* src = tile + static_cast<size_t>(g) * p_pad + (tid + c * NT) * VB;
              if constexpr (VB == 8) {
                const uint2 t2 = *reinterpret_cast<const uint2*>(src);
                x[g][c][0] = t2.x;
                x[g][c][1] = t2.y;
              } else {
                x[g][c][0] = *reinterpret_cast<const uint32_t*>(src);
              }
            }
          }
        auto widen = [](const uint32_t (&w)[VWD], double (&o)[VB]) {
#pragma unroll
          for (int q4 = 0; q4 < VWD; ++q4) {
            o[4 * q4 + 0] = u8_to_f64<0>(w[q4]);
            o[4 * q4 + 1] = u8_to_f64<1>(w[q4]);
            o[4 * q4 + 2] = u8_to_f64<2>(w[q4]);
            o[4 * q4 + 3] = u8_to_f64<3>(w[q4]);
          }
        };

        if constexpr (KEEP) {
#pragma unroll
          for (int g = 0; g < GR; ++g)
#pragma unroll
            for (int c = 0; c < CPT; ++c) widen(x[g][c], xk[g][c]);
        }
        // sweep 1: G b' per row, transposed butterfly over the warp, one partial per (row, warp)
        const int pbuf = cons_seq & 1;
        if (!first) {
          double ps[GR];
#pragma unroll
          for (int g = 0; g < GR; ++g) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              double xd[VB];
              if constexpr (!KEEP) widen(x[g][c], xd);
              const double(&xv)[VB] = *(KEEP ? &xk[KEEP ? g : 0][KEEP ? c : 0] : &xd);
#pragma unroll
              for (int u = 0; u < VB; u += 2) {
                s0 = fma(xv[u], bq[c][u], s0);
                s1 = fma(xv[u + 1], bq[c][u + 1], s1);
              }
            }
            ps[g] = s0 + s1;
          }
          const double t = warp_rows_sum<GR>(ps, lane);
          if ((lane & (LG - 1)) == 0) H->partsum[pbuf][(lane / LG) * NW + warp] = t;
        }
        __syncthreads();  // (A) partial sums visible; every thread holds its tile in registers
        if (tid == PRODUCER_TID) {
          issue_fill(prod_pos, st);
          prod_pos = (prod_pos + 1 == nb_own) ? 0 : prod_pos + 1;
        }
        // lane group g = lane / LG finishes band row g (fixed order), then the rows are exchanged by shuffle
        const int g_me = lane / LG, q = lane & (LG - 1);
        double v = 0.0;
        if (!first) {
          double t = 0.0;
          if constexpr (LG <= NW) {
            constexpr int PER = NW / LG;
#pragma unroll
            for (int i = 0; i < PER; ++i) t += H->partsum[pbuf][g_me * NW + q * PER + i];
          } else {
            if (q < NW) t = H->partsum[pbuf][g_me * NW + q];
          }
#pragma unroll
          for (int o = LG / 2; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          v = t - cshift;
        }
        const bool live = g_me < rows;
        const int srow = (pos - chunk0) * GR + g_me;
        const double Rfull = live ? rs_sm[srow] - v : 0.0;
        const double rn_me = live ? Rfull + (self_next ? v : xs_sm[srow]) : 0.0;  // L == 1: add back the new effect
        if (warp == 0 && q == 0 && live) {  // row writer
          const int row = row_base + g_me;
          if (!first) pb.xb[static_cast<size_t>(l) * nrows + row] = v;
          pb.r[row] = rn_me;
#pragma unroll
          for (int kk = 0; kk < K; ++kk)
            if (kk == k) {
              vsq[kk] += v * v;          // |X b_l|^2 (infer.py:806)
              rsq[kk] += Rfull * Rfull;  // |y - X sum_l b_l|^2 at the last effect (infer.py:804)
            }
        }
        // sweep 2: G^T r_next accumulated in registers.  The raw bytes are laundered through an empty
        // asm so the compiler widens them again here instead of keeping 8 * GR * CPT doubles alive
        // across the barrier (which spills).
        if constexpr (!KEEP) {
#pragma unroll
          for (int g = 0; g < GR; ++g)
#pragma unroll
            for (int c = 0; c < CPT; ++c)
#pragma unroll
              for (int w = 0; w < VWD; ++w) asm volatile("" : "+r"(x[g][c][w]));
        }
#pragma unroll
        for (int g = 0; g < GR; ++g) {
          const double rg = __shfl_sync(0xffffffffu, rn_me, g * LG);
          rsum += rg;
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            double xd[VB];
            if constexpr (!KEEP) widen(x[g][c], xd);
            const double(&xv)[VB] = *(KEEP ? &xk[KEEP ? g : 0][KEEP ? c : 0] : &xd);
#pragma unroll
            for (int u = 0; u < VB; ++u) aq[c][u] = fma(xv[u], rg, aq[c][u]);
          }
        }
        phase_bits ^= (1u << st);
        cons_seq += 1;
        st = (st + 1 == NS) ? 0 : st + 1;
      }
      if (cur_k >= 0) flush_acc(cur_k);
      zero_unowned_partials();
      // store_trav sums trow[0..GMAX): the row writers are lanes 0, LG, 2 LG, ... of warp 0
      store_trav(l, first, vsq, rsq);
    };


    // Hard-call traversal by ROW TEAMS (XT = int8_t).  A row team = T warps that own one band row at a time: lane
    // q of the team holds COLS genotype bytes of the row (NV 16-byte vectors, strided so that shared-memory reads
    // are conflict-free), widens them once and uses the widened values for both sweeps of that row:
    //   v = <g_row, b'> - cshift  (warp butterfly, then T partials through shared memory and one named barrier
    //   of the team)  ->  R = r - v  ->  r_next = R + xb[lnext]  ->  acc[cols] += g_row * r_next.
    // Nothing is block-wide inside a band: the 8 / T teams of a block run rows of the same band independently,
    // each warp hands a band back through a counter, and the LAST warp to finish re-arms the stage with the band
    // NS positions ahead.  Per genotype this costs the widening (3 issue slots) + 2 DFMA and ~1 slot of per-row
    // overhead, against ~4.4 slots of per-band overhead per genotype in the column-tile traversal (transposed
    // butterflies, block barrier, row-sum finishing), and the warps no longer stall on each other.
    // Centring / scaling stay folded into the vectors:
    //   X_std b = G (b / sd) - sum_j mean_j b_j / sd_j,   X_std^T r = (G^T r - mean * sum_i r_i) / sd.
    auto traverse_i8_rows = [&](auto cols_c, auto t_c, int l, bool first, bool last_effect) {
      constexpr int COLS = decltype(cols_c)::value;  // genotype bytes (= columns) per lane
      constexpr int T = decltype(t_c)::value;        // warps per row team
      constexpr int NV = COLS / 16;                  // 16-byte vectors per lane
      constexpr int NTEAMS = NW / T;
      static_assert(COLS % 16 == 0 && NW % T == 0, "row-team geometry");
      (void)last_effect;
      const int team_w = warp / T, tw = warp % T;
      const int q = tw * 32 + lane;  // lane index inside the row team
      const int pv16 = p_pad / 16;
      const int lnext = first ? 0 : ((l + 1 == L) ? 0 : l + 1);
      const bool self_next = !first && lnext == l;
      const bool leader = (tw == 0 && lane == 0);  // publishes the row state of the team's rows
      double vsq[K], rsq[K];  // meaningful in the team leaders only
#pragma unroll
      for (int k = 0; k < K; ++k) vsq[k] = rsq[k] = 0.0;
      double bq[NV][16], aq[NV][16];
      uint32_t vmask = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (q + i * 32 * T < pv16) vmask |= (1u << i);
#pragma unroll
        for (int u = 0; u < 16; ++u) bq[i][u] = aq[i][u] = 0.0;
      }
      int cur_k = -1;
      int st = cons_seq % NS;
      double cshift = 0.0;  // sum_j mean_j b_j / sd_j of the current ancestry
      double rsum = 0.0;    // sum of r_next over the rows of the current ancestry processed by this team
      unsigned rowcnt = 0;  // rows processed by this team (parity picks the partial-sum buffer)
      // shared memory behind the ring: staged row state, then the cross-team exchange area
      double* rs_sm = reinterpret_cast<double*>(ring + static_cast<size_t>(NS) * stage_bytes);
      double* xs_sm = rs_sm + I8_STATE_ROWS;
      double* xteam = xs_sm + I8_STATE_ROWS;  // [NTEAMS][p_pad]; also the staging area of b'
      const int chunk_bands = I8_STATE_ROWS / G;
      auto stage_rows = [&](int pos0, int pos1) {
        __syncthreads();  // previous chunk fully consumed
        for (int i = tid; i < (pos1 - pos0) * G; i += NT) {
          int k, g0, rows;
          band_of(pos0 + i / G, k, g0, rows);
          const int g = i % G;
          double rv = 0.0, xv = 0.0;
          if (g < rows) {
            const int row = pb.row0[k] + g0 + g;
            rv = pb.r[row];
            xv = pb.xb[static_cast<size_t>(lnext) * nrows + row];
          }
          rs_sm[i] = rv;
          xs_sm[i] = xv;
        }
        __syncthreads();
      };
      // column j = 16 v + u of a team's partial sits at ((u / 2) * pv16 + v) * 2 + (u & 1): conflict-free stores
      auto flush_acc = [&](int k) {
        double* xt = xteam + static_cast<size_t>(team_w) * p_pad;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          if (vmask & (1u << i)) {
            const int v = q + i * 32 * T;
#pragma unroll
            for (int u2 = 0; u2 < 8; ++u2)
              *reinterpret_cast<double2*>(xt + (static_cast<size_t>(u2) * pv16 + v) * 2) =
                  make_double2(aq[i][2 * u2], aq[i][2 * u2 + 1]);
          }
        }
        if (q == 0) H->tsum[team_w] = rsum;
        __syncthreads();
        double rs = 0.0;
#pragma unroll
        for (int t = 0; t < NTEAMS; ++t) rs += H->tsum[t];
        double* dst = my_scr + static_cast<size_t>(k) * p_pad;
        const double* cm = pb.cmean + static_cast<size_t>(k) * p_pad;
        const double* cs = pb.cisd + static_cast<size_t>(k) * p_pad;
        for (int j2 = tid; j2 < p_pad / 2; j2 += NT) {
          const int j = 2 * j2, v = j >> 4, u2 = (j & 15) >> 1;
          const double* src = xteam + (static_cast<size_t>(u2) * pv16 + v) * 2;
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int t = 0; t < NTEAMS; ++t) {
            const double2 t2 = *reinterpret_cast<const double2*>(src + static_cast<size_t>(t) * p_pad);
            a0 += t2.x;
            a1 += t2.y;
          }
          const double2 m2 = *reinterpret_cast<const double2*>(cm + j);
          const double2 s2 = *reinterpret_cast<const double2*>(cs + j);
          __stcg(reinterpret_cast<double2*>(dst + j), make_double2((a0 - m2.x * rs) * s2.x, (a1 - m2.y * rs) * s2.y));
        }
        __syncthreads();  // the exchange area is reused
      };

      int chunk0 = 0;
      for (int pos = 0; pos < nb_own; ++pos) {
        if (pos == 0 || pos == chunk0 + chunk_bands) {
          chunk0 = pos;
          stage_rows(chunk0, min(nb_own, chunk0 + chunk_bands));
        }
        int k, g0, rows;
        band_of(pos, k, g0, rows);
        if (k != cur_k) {
          if (cur_k >= 0) flush_acc(cur_k);
          // b' = b_l / sd staged through shared memory with coalesced loads; cshift = sum_j mean_j b'_j
          const double* bl = effect_vec(l, k);
          const double* cm = pb.cmean + static_cast<size_t>(k) * p_pad;
          const double* cs = pb.cisd + static_cast<size_t>(k) * p_pad;
          double part[1] = {0.0};
          for (int j = tid; j < p_pad; j += NT) {
            const double b = first ? 0.0 : __ldcg(&bl[j]) * cs[j];
            xteam[j] = b;
            part[0] += cm[j] * b;
          }
          block_sum<1, NW>(H, part);
          cshift = H->tot[0];
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int v = q + i * 32 * T;
            const bool on = vmask & (1u << i);
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
              const double2 b2 = on ? *reinterpret_cast<const double2*>(xteam + v * 16 + u) : make_double2(0.0, 0.0);
              bq[i][u] = b2.x;
              bq[i][u + 1] = b2.y;
              aq[i][u] = aq[i][u + 1] = 0.0;
            }
          }
          rsum = 0.0;
          cur_k = k;
          __syncthreads();  // the staging area is free again
        }
        const unsigned char* tile = ring + static_cast<size_t>(st) * stage_bytes;
        const int row_base = pb.row0[k] + g0;
        mbar_wait(&H->mbar[st], (phase_bits >> st) & 1u);

        for (int g = team_w; g < rows; g += NTEAMS) {
          const unsigned char* src = tile + static_cast<size_t>(g) * p_pad;
          double xk[NV][16];
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            uint4 w4 = make_uint4(0u, 0u, 0u, 0u);
            if (vmask & (1u << i)) w4 = *reinterpret_cast<const uint4*>(src + (q + i * 32 * T) * 16);
            const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              xk[i][4 * q4 + 0] = u8_to_f64<0>(w[q4]);
              xk[i][4 * q4 + 1] = u8_to_f64<1>(w[q4]);
              xk[i][4 * q4 + 2] = u8_to_f64<2>(w[q4]);
              xk[i][4 * q4 + 3] = u8_to_f64<3>(w[q4]);
            }
          }
          double v = 0.0;
          if (!first) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int i = 0; i < NV; ++i)
#pragma unroll
              for (int u = 0; u < 16; u += 4) {
                s0 = fma(xk[i][u], bq[i][u], s0);
                s1 = fma(xk[i][u + 1], bq[i][u + 1], s1);
                s2 = fma(xk[i][u + 2], bq[i][u + 2], s2);
                s3 = fma(xk[i][u + 3], bq[i][u + 3], s3);
              }
            double s = warp_sum((s0 + s1) + (s2 + s3));  // xor butterfly: every lane holds the warp total
            if constexpr (T > 1) {
              const int par = rowcnt & 1u;
              if (lane == 0) H->partsum[par][warp] = s;
              asm volatile("bar.sync %0, %1;" ::"r"(1 + team_w), "r"(32 * T) : "memory");
              s = 0.0;
#pragma unroll
              for (int t = 0; t < T; ++t) s += H->partsum[par][team_w * T + t];
            }
            v = s - cshift;
          }
          const int srow = (pos - chunk0) * G + g;
          const double Rfull = rs_sm[srow] - v;
          const double rn = Rfull + (self_next ? v : xs_sm[srow]);  // L == 1: add back the new effect
          if (leader) {
            const int row = row_base + g;
            if (!first) pb.xb[static_cast<size_t>(l) * nrows + row] = v;
            pb.r[row] = rn;
#pragma unroll
            for (int kk = 0; kk < K; ++kk)
              if (kk == k) {
                vsq[kk] += v * v;          // |X b_l|^2 (infer.py:806)
                rsq[kk] += Rfull * Rfull;  // |y - X sum_l b_l|^2 at the last effect (infer.py:804)
              }
          }
#pragma unroll
          for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int u = 0; u < 16; ++u) aq[i][u] = fma(xk[i][u], rn, aq[i][u]);
          rsum += rn;
          rowcnt += 1u;
        }
        // hand the band back; the last warp to finish re-arms the stage with the band NS positions ahead
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          const unsigned old = atomicAdd(&H->drain[st], 1u);
          if ((old & (NW - 1)) == NW - 1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_fill((cons_seq + NS) % nb_own, st);
          }
        }
        phase_bits ^= (1u << st);
        cons_seq += 1;
        st = (st + 1 == NS) ? 0 : st + 1;
      }
      if (cur_k >= 0) flush_acc(cur_k);
      zero_unowned_partials();
      prod_pos = (cons_seq + NS) % max(nb_own, 1);  // ring invariant, for the traversals that track it
      // per-traversal scalars: fixed-order block sum of the team leaders' partials
      {
        double tv[2 * K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          tv[k] = leader ? vsq[k] : 0.0;
          tv[K + k] = leader ? rsq[k] : 0.0;
        }
        block_sum<2 * K, NW>(H, tv);
        if (tid < K && !first) {
          double* dst = my_scr + P.off_trav + l * TRAV_STRIDE;
          __stcg(&dst[tid], H->tot[tid]);
          __stcg(&dst[KMAX + tid], H->tot[K + tid]);
        }
        __syncthreads();
      }
    };


    // Bit-plane traversal (XT = uint8_t): hard calls 0/1/2 are two bit planes, B1 = (g >= 1) and B2 = (g == 2), so
    // g = B1 + B2 and both products are SUBSET SUMS that come out of lookup tables instead of fp64 multiply-adds
    // on widened genotypes (the column-tile traversal spends 3 of its 5 issue slots per genotype on widening):
    //   sweep 1   v_i  = sum_groups T[g][bits of row i in column group g]     T[g][e]  = sum_{c in e} b'_{8 g + c}
    //   sweep 2   a_j  = sum_groups T[rg][bits of column j in row group rg]   T[rg][e] = sum_{r in e} r_next_{8 rg + r}
    // i.e. one table read + one add per 8 genotypes and plane.  The two sweeps are the same code on the two bit
    // layouts (row-major / column-major).  A pass handles 256 "inner" indices (32 groups of 8: one 64 KB table set
    // in shared memory, entry-major so that the 32 lanes of a warp -- lane = group -- always read 32 different
    // banks) against up to 512 "outer" indices (a warp owns 64 of them, 64 lane-private sums in registers); after
    // the last pass of a round a transposed butterfly turns the lane-private sums into totals.  The look-up address
    // is built by ONE prmt: the table set is aligned to 64 KB, so (base | bits << 8 | lane << 3) is a byte splice.
    // The bit stream arrives through a 3-stage ring of bulk copies; the last warp to finish a piece re-arms its
    // stage.  Team members split the rows in sweep 1 and the columns (their own SNP slice) in sweep 2, so X^T r is
    // complete per block and needs no reduction.
    auto traverse_b2 = [&](auto /*instantiate on use*/, int l, bool first) {
      const int lnext = first ? 0 : ((l + 1 == L) ? 0 : l + 1);
      const bool self_next = !first && lnext == l;
      const int ntiles = pb.ntiles;
      constexpr int RQ = B2_ROUND_ROWS / 4;  // outer quads per round
      // ---- shared memory: [header][small vectors][ring stage 0] | 64 KB boundary | [table set 0][table set 1][ring stage 1]
      const uint32_t smem_base = smem_u32(smem_raw);
      const uint32_t hdr_end = smem_base + static_cast<uint32_t>(((sizeof(SmemHdr) + 127) / 128) * 128);
      const uint32_t low_end = hdr_end + 8192u + B2_MAX_ROWS * 8u + B2_STAGE_BYTES;
      const uint32_t tab_addr = (low_end + B2_TABLE_BYTES - 1u) & ~static_cast<uint32_t>(B2_TABLE_BYTES - 1);
      double* const tab0 = reinterpret_cast<double*>(smem_raw + (tab_addr - smem_base));
      double* const bstage = reinterpret_cast<double*>(smem_raw + (hdr_end - smem_base));  // [2][256] b' of a chunk
      double* const vsum = bstage + 2 * B2_TILE_COLS;                                        // [512] row sums of a round
      double* const rn_sm = vsum + B2_ROUND_ROWS;                                            // [B2_MAX_ROWS] next residual
      unsigned char* const stage_lo = reinterpret_cast<unsigned char*>(rn_sm + B2_MAX_ROWS);
      unsigned char* const stage_hi = reinterpret_cast<unsigned char*>(tab0) + 2 * B2_TABLE_BYTES;
      auto stage_ptr = [&](int stg) { return stg == 0 ? stage_lo : stage_hi; };
      auto table_ptr = [&](int t) { return tab0 + static_cast<size_t>(t) * (B2_TABLE_BYTES / 8); };
      const uint32_t lane_const0 = tab_addr | (static_cast<uint32_t>(lane) << 3);
      auto lane_const_of = [&](int t) { return lane_const0 + static_cast<uint32_t>(t) * B2_TABLE_BYTES; };
      // ---- work split inside the team
      int q0[K], q1[K], rounds1[K], rounds2[K], nch2[K];
      int npa = 0, np_total;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int nq = (pb.n[k] + 3) >> 2;  // row quads, split over the team
        q0[k] = static_cast<int>((static_cast<long long>(nq) * rank) / C);
        q1[k] = static_cast<int>((static_cast<long long>(nq) * (rank + 1)) / C);
        rounds1[k] = first ? 0 : (q1[k] - q0[k] + RQ - 1) / RQ;
        npa += rounds1[k] * ntiles;
      }
      const int t0 = js / B2_TILE_COLS, t1 = (je > js) ? (je + B2_TILE_COLS - 1) / B2_TILE_COLS : t0;  // own column tiles
      const int cq0 = 64 * t0, cq1 = 64 * t1;  // own column quads
      np_total = npa;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        nch2[k] = (pb.n[k] + 255) >> 8;  // 256-row chunks
        rounds2[k] = (cq1 - cq0 + RQ - 1) / RQ;
        np_total += rounds2[k] * nch2[k];
      }
      // piece s of the static stream -> (source, bytes)
      auto issue_piece = [&](int s, int stg) {
        const unsigned char* src = nullptr;
        uint32_t bytes = 0;
        int r = s;
        bool found = false;
        if (r < npa) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const int cnt = rounds1[k] * ntiles;
            if (!found && r < cnt) {
              const int rd = r / ntiles, c = r - rd * ntiles;
              const int nq = (pb.n[k] + 3) >> 2;
              const int qa = q0[k] + rd * RQ;
              src = pb.rbits[k] + (static_cast<size_t>(c) * nq + qa) * 256;
              bytes = static_cast<uint32_t>(min(RQ, q1[k] - qa)) * 256u;
              found = true;
            }
            if (!found) r -= cnt;
          }
        } else {
          r -= npa;
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const int cnt = rounds2[k] * nch2[k];
            if (!found && r < cnt) {
              const int rd = r / nch2[k], d = r - rd * nch2[k];
              const int qa = cq0 + rd * RQ;
              src = pb.cbits[k] + (static_cast<size_t>(d) * (64 * ntiles) + qa) * 256;
              bytes = static_cast<uint32_t>(min(RQ, cq1 - qa)) * 256u;
              found = true;
            }
            if (!found) r -= cnt;
          }
        }
        mbar_expect_tx(&H->mbar[stg], bytes);
        bulk_g2s(stage_ptr(stg), src, bytes, &H->mbar[stg], stream_policy);
      };
      int pseq = 0;  // pieces consumed in this traversal
      auto piece_wait = [&]() -> const unsigned char* {
        const int stg = cons_seq % B2_NSTAGE;
        mbar_wait(&H->mbar[stg], (phase_bits >> stg) & 1u);
        return stage_ptr(stg);
      };
      auto piece_done = [&]() {
        const int stg = cons_seq % B2_NSTAGE;
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          const unsigned old = atomicAdd(&H->drain[stg], 1u);
          if ((old & (NW - 1)) == NW - 1 && pseq + B2_NSTAGE < np_total) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_piece(pseq + B2_NSTAGE, stg);
          }
        }
        phase_bits ^= (1u << stg);
        cons_seq += 1;
        pseq += 1;
      };
      // table set of 32 groups from 256 weights: thread (g, q) writes entries e = (256 / NW) q + r of group g
      // (fixed association: low nibble sum + high nibble sum)
      auto build_tables = [&](double* tab, const double* w256) {
        const int g = tid & 31, qq = tid >> 5;
        const double2* b8 = reinterpret_cast<const double2*>(w256 + 8 * g);
        const double2 b01 = b8[0], b23 = b8[1], b45 = b8[2], b67 = b8[3];
        double lo[16];
        lo[0] = 0.0;
        lo[1] = b01.x;
        lo[2] = b01.y;
        lo[3] = b01.x + b01.y;
        lo[4] = b23.x;
        lo[5] = b01.x + b23.x;
        lo[6] = b01.y + b23.x;
        lo[7] = lo[3] + b23.x;
        lo[8] = b23.y;
        lo[9] = b01.x + b23.y;
        lo[10] = b01.y + b23.y;
        lo[11] = lo[3] + b23.y;
        lo[12] = b23.x + b23.y;
        lo[13] = lo[5] + b23.y;
        lo[14] = lo[6] + b23.y;
        lo[15] = lo[7] + b23.y;
        constexpr int EPT = 256 / NW;  // entries per thread: 32 q .. or 16 q .. of group g
        double hi2[EPT / 16];
#pragma unroll
        for (int h = 0; h < EPT / 16; ++h) {
          const int nib = (EPT / 16) * qq + h;
          double t = 0.0;
          if (nib & 1) t = b45.x;
          if (nib & 2) t = (nib & 1) ? t + b45.y : b45.y;
          if (nib & 4) t = (nib & 3) ? t + b67.x : b67.x;
          if (nib & 8) t = (nib & 7) ? t + b67.y : b67.y;
          hi2[h] = t;
        }
        double* dst = tab + static_cast<size_t>(EPT * qq) * 32 + g;
#pragma unroll
        for (int r = 0; r < EPT; ++r) dst[r * 32] = lo[r & 15] + hi2[r >> 4];
      };
      // one pass: the warp's OQ outer quads (OW outer indices) of the piece against the resident table set
      constexpr int OW = B2_ROUND_ROWS / NW, OQ = OW / 4;
      auto lookup_pass = [&](const unsigned char* piece, int qn, uint32_t lane_const, double (&acc)[OW]) {
        const uint32_t* words = reinterpret_cast<const uint32_t*>(piece) + (OQ * warp) * 64 + lane;
#pragma unroll
        for (int oc = 0; oc < OQ / 4; ++oc) {  // four quads at a time (uniform skip past the end of the round)
          if (OQ * warp + 4 * oc < qn) {
#pragma unroll
            for (int m4 = 0; m4 < 4; ++m4) {
              const int m = 4 * oc + m4;
              if (OQ * warp + m < qn) {
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
                  const uint32_t w = words[(m * 2 + pl) * 32];
                  acc[4 * m + 0] += lds_f64(prmt(w, lane_const, 0x7604u));
                  acc[4 * m + 1] += lds_f64(prmt(w, lane_const, 0x7614u));
                  acc[4 * m + 2] += lds_f64(prmt(w, lane_const, 0x7624u));
                  acc[4 * m + 3] += lds_f64(prmt(w, lane_const, 0x7634u));
                }
              }
            }
          }
        }
      };
      // lane-private sums -> totals: lane i of the warp gets outer indices OW w + 32 s + i (s < OW / 32)
      auto totals = [&](double (&acc)[OW], double (&tot)[OW / 32]) {
#pragma unroll
        for (int sset = 0; sset < OW / 32; ++sset) {
          double part[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) part[i] = acc[32 * sset + i];
          tot[sset] = warp_rows_sum<32>(part, lane);
        }
      };
      __syncthreads();  // every warp has left the previous phase: the ring and the tables are free
      if (tid == 0)
        for (int s = 0; s < B2_NSTAGE && s < np_total; ++s) issue_piece(s, (cons_seq + s) % B2_NSTAGE);

      double vsq[K], rsq[K];
#pragma unroll
      for (int k = 0; k < K; ++k) vsq[k] = rsq[k] = 0.0;

      if (first) {
        // locus initialisation (state owned by this block: its rows): r = y, cached X b_l = 0
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int ia = 4 * q0[k], ib = min(pb.n[k], 4 * q1[k]);
          for (int i = ia + tid; i < ib; i += NT) {
            const int row = pb.row0[k] + i;
            pb.r[row] = pb.y[row];
            for (int ll = 0; ll < L; ++ll) pb.xb[static_cast<size_t>(ll) * nrows + row] = 0.0;
          }
        }
      } else {
        // ---------------- sweep 1: v = G b' - cshift for the rows of this block, then the row state
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
          const double* bl = effect_vec(l, k);
          const double* cm = pb.cmean + static_cast<size_t>(k) * p_pad;
          const double* cs = pb.cisd + static_cast<size_t>(k) * p_pad;
          // centring weights of column j: its mean, or -- genotypes residualised on covariates -- row j of M = G^T Q
          const int qc = pb.qc[k];
          const double* Mq = pb.Mq[k];
          const double* Qb = pb.Qb[k];
          auto load_b = [&](int c) -> double {  // b' of column 256 c + tid
            const int j = c * B2_TILE_COLS + tid;
            return (tid < B2_TILE_COLS && c < ntiles && j < p_pad) ? __ldcg(&bl[j]) * cs[j] : 0.0;
          };
          // centring terms of the ancestry, once, outside the chunk pipeline (its registers are the look-up sums):
          // sum_j mean_j b'_j, or Q^T G b' = M^T b' with covariates; parked in shared memory until the rows finish
          {
            double part[B2_QMAX];
#pragma unroll
            for (int t = 0; t < B2_QMAX; ++t) part[t] = 0.0;
            for (int j = tid; j < p_pad; j += NT) {
              const double b = __ldcg(&bl[j]) * cs[j];
              if (qc == 0) {
                part[0] += cm[j] * b;
              } else {
#pragma unroll
                for (int t = 0; t < B2_QMAX; ++t)
                  if (t < qc) part[t] += Mq[static_cast<size_t>(t) * p_pad + j] * b;
              }
            }
            if (qc == 0) {  // (block-uniform) one value to reduce instead of B2_QMAX
              double one[1] = {part[0]};
              block_sum<1, NW>(H, one);
            } else {
              block_sum<B2_QMAX, NW>(H, part);
            }
            if (tid < B2_QMAX) H->qv[tid] = H->tot[tid];
            __syncthreads();
          }
#pragma unroll 1
          for (int rd = 0; rd < rounds1[k]; ++rd) {
            const int qa = q0[k] + rd * RQ;
            const int qn = min(RQ, q1[k] - qa);  // row quads of this round
            double acc[OW];
#pragma unroll
            for (int i = 0; i < OW; ++i) acc[i] = 0.0;
            // chunk pipeline: b' of chunk c + 2 is loaded and staged, the table set of chunk c + 1 is built and the
            // look-ups of chunk c run between two block barriers (two table sets, two staging halves)
            double b_reg = load_b(0);
            auto stage_b = [&](int c) {  // stage b' of chunk c (held in b_reg), start loading chunk c + 1
              if (tid < B2_TILE_COLS) bstage[(c & 1) * B2_TILE_COLS + tid] = b_reg;
              b_reg = load_b(c + 1);
            };
            stage_b(0);
            __syncthreads();
            build_tables(table_ptr(0), bstage);
            if (ntiles > 1) stage_b(1);
            __syncthreads();
#pragma unroll 1
            for (int c = 0; c < ntiles; ++c) {
              const unsigned char* piece = piece_wait();
              lookup_pass(piece, qn, lane_const_of(c & 1), acc);
              piece_done();
              if (c + 1 < ntiles) build_tables(table_ptr((c + 1) & 1), bstage + ((c + 1) & 1) * B2_TILE_COLS);
              if (c + 2 < ntiles) stage_b(c + 2);
              __syncthreads();
            }
            {
              double tot[OW / 32];
              totals(acc, tot);
#pragma unroll
              for (int sset = 0; sset < OW / 32; ++sset) vsum[OW * warp + 32 * sset + lane] = tot[sset];
            }
            __syncthreads();  // vsum visible
            double tq[B2_QMAX];  // sum_j mean_j b'_j, or Q^T G b'
#pragma unroll
            for (int t = 0; t < B2_QMAX; ++t) tq[t] = H->qv[t];
            // finish the rows of the round
            for (int i = tid; i < 4 * qn; i += NT) {
              const int irow = 4 * qa + i;
              if (irow < pb.n[k]) {
                const int row = pb.row0[k] + irow;
                double shift = tq[0];
                if (qc > 0) {
                  shift = 0.0;
#pragma unroll
                  for (int t = 0; t < B2_QMAX; ++t)
                    if (t < qc) shift += Qb[static_cast<size_t>(irow) * qc + t] * tq[t];
                }
                const double v = vsum[i] - shift;
                const double Rfull = pb.r[row] - v;
                const double rn = Rfull + (self_next ? v : pb.xb[static_cast<size_t>(lnext) * nrows + row]);
                pb.xb[static_cast<size_t>(l) * nrows + row] = v;
                pb.r[row] = rn;
#pragma unroll
                for (int kk = 0; kk < K; ++kk)
                  if (kk == k) {
                    vsq[kk] += v * v;          // |X b_l|^2 (infer.py:806)
                    rsq[kk] += Rfull * Rfull;  // |y - X sum_l b_l|^2 at the last effect (infer.py:804)
                  }
              }
            }
            __syncthreads();  // vsum / bstage are reused by the next round
          }
        }
      }
      team_sync(team);  // the next residual of every row (all team members) is in global memory

      // ---------------- sweep 2: X_std^T r_next for the own column tiles (same passes on the column-major bits)
#pragma unroll 1
      for (int k = 0; k < K; ++k) {
        const int nk = pb.n[k];
        const int qc = pb.qc[k];
        const double* Mq = pb.Mq[k];
        const double* Qb = pb.Qb[k];
        double rpart[B2_QMAX];  // sum_i r_i, or Q^T r
#pragma unroll
        for (int t = 0; t < B2_QMAX; ++t) rpart[t] = 0.0;
        for (int i = tid; i < (nch2[k] << 8); i += NT) {
          const double rv = (i < nk) ? __ldcg(&pb.r[pb.row0[k] + i]) : 0.0;
          rn_sm[i] = rv;
          if (qc == 0) {
            rpart[0] += rv;
          } else if (i < nk) {
#pragma unroll
            for (int t = 0; t < B2_QMAX; ++t)
              if (t < qc) rpart[t] += Qb[static_cast<size_t>(i) * qc + t] * rv;
          }
        }
        if (qc == 0) {  // (block-uniform) one value to reduce instead of B2_QMAX
          double one[1] = {rpart[0]};
          block_sum<1, NW>(H, one);  // (barriers inside: rn_sm visible)
        } else {
          block_sum<B2_QMAX, NW>(H, rpart);
        }
        if (tid < B2_QMAX) H->qv[B2_QMAX + tid] = H->tot[tid];  // parked: the rounds' registers are the look-up sums
        __syncthreads();
        const double* cm = pb.cmean + static_cast<size_t>(k) * p_pad;
        const double* cs = pb.cisd + static_cast<size_t>(k) * p_pad;
        double* dst = pb.xty + static_cast<size_t>(k) * p_pad;
        // up to two 256-row chunks: both table sets stay resident for every round of the ancestry (no block barrier
        // inside the rounds); more rows: one table set is rebuilt per pass
        const bool resident = nch2[k] <= 2;
        if (resident) {
          for (int d = 0; d < nch2[k]; ++d) build_tables(table_ptr(d), rn_sm + 256 * d);
          __syncthreads();
        }
#pragma unroll 1
        for (int rd = 0; rd < rounds2[k]; ++rd) {
          const int qa = cq0 + rd * RQ;
          const int qn = min(RQ, cq1 - qa);  // column quads of this round
          double acc[OW];
#pragma unroll
          for (int i = 0; i < OW; ++i) acc[i] = 0.0;
#pragma unroll 1
          for (int d = 0; d < nch2[k]; ++d) {
            if (!resident) {
              __syncthreads();  // the previous table set is no longer read
              build_tables(table_ptr(0), rn_sm + 256 * d);
              __syncthreads();
            }
            const unsigned char* piece = piece_wait();
            lookup_pass(piece, qn, lane_const_of(resident ? d : 0), acc);
            piece_done();
          }
          // lane-private column sums -> totals; X_std^T r = (G^T r - mean * sum r) / sd for the own slice
          double tot[OW / 32];
          totals(acc, tot);
#pragma unroll
          for (int sset = 0; sset < OW / 32; ++sset) {
            const int j = 4 * qa + OW * warp + 32 * sset + lane;
            if (j >= js && j < je) {
              double corr = cm[j] * H->qv[B2_QMAX];
              if (qc > 0) {
                corr = 0.0;
                for (int t = 0; t < qc; ++t) corr += Mq[static_cast<size_t>(t) * p_pad + j] * H->qv[B2_QMAX + t];
              }
              dst[j] = (tot[sset] - corr) * cs[j];
            }
          }
        }
        __syncthreads();  // rn_sm is reloaded for the next ancestry
      }
      // per-traversal scalars: fixed-order block sums
      {
        double tv[2 * K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          tv[k] = vsq[k];
          tv[K + k] = rsq[k];
        }
        block_sum<2 * K, NW>(H, tv);
        if (tid < K && !first) {
          double* dst = my_scr + P.off_trav + l * TRAV_STRIDE;
          __stcg(&dst[tid], H->tot[tid]);
          __stcg(&dst[KMAX + tid], H->tot[K + tid]);
        }
        __syncthreads();
      }
    };

    // Summary-level register-tile variant: rows are LD rows, there is no X^T r sweep, so a warp
    // never needs another warp's result while it streams.  Every thread keeps its CPT column
    // vectors of b in registers, takes its slice of each row straight from the shared-memory
    // band, and the per-warp partial dot products of up to SS_CHUNK_ROWS rows are parked in
    // shared memory.  Warps run free of block-wide barriers; a stage is handed back to the
    // bulk-copy producer through a "drained" mbarrier (one arrival per warp).  Once per chunk
    // the block finishes the rows in parallel (row sums in fixed order, residual update,
    // cached XtX b_l, ERSS terms).
    auto traverse_ss = [&](auto cpt_c, auto gr_c, int l, bool last_effect) {
      constexpr int CPT = decltype(cpt_c)::value;
      constexpr int GR = decltype(gr_c)::value;
      constexpr int PRODUCER_TID = NT - 32;
      constexpr int CHUNK_BANDS = SS_CHUNK_ROWS / GR;
      static_assert(SS_CHUNK_ROWS <= NT, "one finishing thread per parked row");
      const int lnext = (l + 1 == L) ? 0 : l + 1;
      double* psum = reinterpret_cast<double*>(ring + static_cast<size_t>(NS) * stage_bytes);  // [rows][NW]
      double tv[2 * K];  // per-thread ERSS terms: b_l^T XtX b_l, |full residual|^2 per ancestry
#pragma unroll
      for (int i = 0; i < 2 * K; ++i) tv[i] = 0.0;
      double bq[CPT][VW];
      uint32_t vmask = 0;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if (tid + c * NT < pv) vmask |= (1u << c);
#pragma unroll
        for (int u = 0; u < VW; ++u) bq[c][u] = 0.0;
      }
      int cur_k = -1;
      int st = cons_seq % NS;

      for (int pos0 = 0; pos0 < nb_own; pos0 += CHUNK_BANDS) {
        const int pos1 = min(nb_own, pos0 + CHUNK_BANDS);
        for (int pos = pos0; pos < pos1; ++pos) {
          int k, g0, rows;
          band_of(pos, k, g0, rows);
          if (k != cur_k) {
            const double* bl = effect_vec(l, k);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              const int v = tid + c * NT;
              const bool on = vmask & (1u << c);
#pragma unroll
              for (int u = 0; u < VW; u += 2) {
                const double2 t2 = on ? __ldcg(reinterpret_cast<const double2*>(bl + v * VW + u)) : make_double2(0.0, 0.0);
                bq[c][u] = t2.x;
                bq[c][u + 1] = t2.y;
              }
            }
            cur_k = k;
          }
          const XT* tile = reinterpret_cast<const XT*>(ring + static_cast<size_t>(st) * stage_bytes);
          mbar_wait(&H->mbar[st], (phase_bits >> st) & 1u);
          // rows past the end of a short band hold stale (finite or not) data: their sums are never read
          double ps[GR];
#pragma unroll
          for (int g = 0; g < GR; ++g) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              if (vmask & (1u << c)) {
                double x[VW];
                Vec<XT>::load(tile + static_cast<size_t>(g) * p_pad, tid + c * NT, x);
                if constexpr (VW == 4) {
                  s0 = fma(x[0], bq[c][0], s0);
                  s1 = fma(x[1], bq[c][1], s1);
                  s2 = fma(x[2], bq[c][2], s2);
                  s3 = fma(x[3], bq[c][3], s3);
                } else {
                  if (c & 1) {
                    s2 = fma(x[0], bq[c][0], s2);
                    s3 = fma(x[1], bq[c][1], s3);
                  } else {
                    s0 = fma(x[0], bq[c][0], s0);
                    s1 = fma(x[1], bq[c][1], s1);
                  }
                }
              }
            }
            ps[g] = (s0 + s1) + (s2 + s3);
          }
          double t;
          if constexpr (GR == 4) {
            const bool hi = lane & 16;
            double k0 = hi ? ps[2] : ps[0], k1 = hi ? ps[3] : ps[1];
            const double q0 = hi ? ps[0] : ps[2], q1 = hi ? ps[1] : ps[3];
            k0 += __shfl_xor_sync(0xffffffffu, q0, 16);
            k1 += __shfl_xor_sync(0xffffffffu, q1, 16);
            const bool hi2 = lane & 8;
            t = hi2 ? k1 : k0;
            const double q = hi2 ? k0 : k1;
            t += __shfl_xor_sync(0xffffffffu, q, 8);
            t += __shfl_xor_sync(0xffffffffu, t, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
          } else if constexpr (GR == 2) {
            const bool hi = lane & 16;
            t = hi ? ps[1] : ps[0];
            const double q = hi ? ps[0] : ps[1];
            t += __shfl_xor_sync(0xffffffffu, q, 16);
            t += __shfl_xor_sync(0xffffffffu, t, 8);
            t += __shfl_xor_sync(0xffffffffu, t, 4);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 1);
          } else {
            t = warp_sum(ps[0]);
          }
          // the shuffles above are a warp-wide convergence point: every lane has finished reading the stage
          if ((lane & (32 / GR - 1)) == 0) psum[((pos - pos0) * GR + lane / (32 / GR)) * NW + warp] = t;
          if (lane == 0) mbar_arrive(&H->ebar[st]);
          if (tid == PRODUCER_TID) {  // refill the stage as soon as all warps have drained it
            mbar_wait(&H->ebar[st], (ephase_bits >> st) & 1u);
            ephase_bits ^= (1u << st);
            issue_fill(prod_pos, st);
            prod_pos = (prod_pos + 1 == nb_own) ? 0 : prod_pos + 1;
          }
          phase_bits ^= (1u << st);
          cons_seq += 1;
          st = (st + 1 == NS) ? 0 : st + 1;
        }
        __syncthreads();  // partial sums of the chunk visible
        {
          const int pos = pos0 + tid / GR, g = tid % GR;
          if (tid < SS_CHUNK_ROWS && pos < pos1) {
            int k, g0, rows;
            band_of(pos, k, g0, rows);
            if (g < rows) {
              const int row = pb.row0[k] + g0 + g;
              const double2* ps2 = reinterpret_cast<const double2*>(&psum[tid * NW]);
              static_assert(NW == 8 || B2, "pairwise tree below assumes 8 warps");
              const double2 q0 = ps2[0], q1 = ps2[1], q2 = ps2[2], q3 = ps2[3];
              const double v = (((q0.x + q0.y) + (q1.x + q1.y)) + ((q2.x + q2.y) + (q3.x + q3.y))) * pb.nscale[k];
              const double Rfull = pb.r[row] - v;
              pb.r[row] = Rfull + (lnext == l ? v : pb.xb[static_cast<size_t>(lnext) * nrows + row]);  // L == 1
              pb.xb[static_cast<size_t>(l) * nrows + row] = v;
              if (last_effect) pb.rfull[row] = Rfull;
              const double vterm = __ldcg(&effect_vec(l, k)[g0 + g]) * v;  // b_l^T XtX b_l (infer_ss.py:572)
#pragma unroll
              for (int kk = 0; kk < K; ++kk)
                if (kk == k) {
                  tv[kk] += vterm;
                  tv[K + kk] += Rfull * Rfull;
                }
            }
          }
        }
        __syncthreads();  // the parking area is reused by the next chunk
      }
      block_sum<2 * K, NW>(H, tv);
      if (tid < K) {
        double* dst = my_scr + P.off_trav + l * TRAV_STRIDE;
        __stcg(&dst[tid], H->tot[tid]);
        __stcg(&dst[KMAX + tid], H->tot[K + tid]);
      }
      __syncthreads();
    };

    // Generic variant (any p that fits, any G <= GMAX): b and the accumulators in shared memory,
    // two sweeps over the shared-memory band separated by a barrier.
    auto traverse_generic = [&](auto /*instantiate on use*/, int l, bool first, bool last_effect) {
      const int lnext = first ? 0 : ((l + 1 == L) ? 0 : l + 1);
      double vsq[K], rsq[K];  // meaningful in threads < G only
#pragma unroll
      for (int k = 0; k < K; ++k) vsq[k] = rsq[k] = 0.0;
      int cur_k = -1;
      int pending_refill = -1;  // stage to refill once every thread has left its sweep 2
      // loci too wide for b / the X^T r accumulator to sit in shared memory next to two band stages keep both
      // in global memory (b is read from the materialised effect vector, the accumulator is the block's
      // partial X^T r itself); slower (L2 traffic per row) but any row that fits the ring twice works
      const bool big = (fast == FAST_GLOBAL_VECTORS);
      double* bsm = big ? nullptr : bsm_sm;
      double* acc = big ? nullptr : acc_sm;

      for (int pos = 0; pos < nb_own; ++pos) {
        int k, g0, rows;
        band_of(pos, k, g0, rows);
        if (k != cur_k) {
          // ancestry change: flush X^T r partial, load the new b, clear the accumulator
          __syncthreads();
          if (!SS && cur_k >= 0 && !big) {
            double* dst = my_scr + static_cast<size_t>(cur_k) * p_pad;
            for (int j = tid; j < p_pad; j += NT) __stcg(&dst[j], acc[j]);
          }
          if (big) {
            bsm = const_cast<double*>(effect_vec(l, k));  // read-only; not touched when first
            acc = my_scr + static_cast<size_t>(k) * p_pad;
          } else if (!first) {
            for (int j = tid; j < p_pad; j += NT) bsm[j] = __ldcg(&effect_vec(l, k)[j]);
          }
          if (!SS)
            for (int j = tid; j < p_pad; j += NT) acc[j] = 0.0;
          cur_k = k;
          __syncthreads();
        }
        const int st = cons_seq % NS;
        const XT* tile = reinterpret_cast<const XT*>(ring + static_cast<size_t>(st) * stage_bytes);

        // designated threads prefetch the per-row state while the band lands
        double r_old = 0.0, xb_next = 0.0;
        const int row = pb.row0[k] + g0 + tid;
        if (tid < rows) {
          r_old = pb.r[row];
          xb_next = pb.xb[static_cast<size_t>(lnext) * nrows + row];
        }
        mbar_wait(&H->mbar[st], (phase_bits >> st) & 1u);

        // sweep 1: v[g] = <X[g,:], b>
        if (!first) {
          const int nworkers = NW / NSEG;  // row workers of NSEG warps each
          const int wk = warp / NSEG, seg = warp % NSEG;
          if (wk < nworkers) {
            for (int g = wk; g < rows; g += nworkers) {
              const XT* xr = tile + static_cast<size_t>(g) * p_pad;
              double a0 = 0.0, a1 = 0.0;
              for (int v = seg * 32 + lane; v < pv; v += NSEG * 32) {
                double x[VW];
                Vec<XT>::load(xr, v, x);
                const double* bb = bsm + v * VW;
#pragma unroll
                for (int u = 0; u < VW; u += 2) {
                  // (global-memory b is rewritten every iteration: read it past L1)
                  const double2 b2 = big ? __ldcg(reinterpret_cast<const double2*>(bb + u))
                                         : *reinterpret_cast<const double2*>(bb + u);
                  a0 = fma(x[u], b2.x, a0);
                  a1 = fma(x[u + 1], b2.y, a1);
                }
              }
              const double s = warp_sum(a0 + a1);
              if (lane == 0) H->partsum[cons_seq & 1][g * NSEG + seg] = s;
            }
          }
        }
        __syncthreads();  // (A) partial sums visible; every thread has left the previous sweep 2
        if (tid == 0 && pending_refill >= 0) {
          issue_fill(prod_pos, pending_refill);
          prod_pos = (prod_pos + 1 == nb_own) ? 0 : prod_pos + 1;
        }
        if (tid < rows) {
          double v = 0.0;
          if (!first) {
            for (int s = 0; s < NSEG; ++s) v += H->partsum[cons_seq & 1][tid * NSEG + s];
            v *= pb.nscale[k];
            pb.xb[static_cast<size_t>(l) * nrows + row] = v;
          }
          const double Rfull = r_old - v;
          const double rnext = Rfull + ((lnext == l && !first) ? v : xb_next);  // L == 1: add back the new effect
          pb.r[row] = rnext;
          H->rn[tid] = rnext;
          // b_l^T XtX b_l (infer_ss.py:572)  |  |X b_l|^2 (infer.py:806)
          const double vterm = SS ? (big ? __ldcg(&bsm[g0 + tid]) : bsm[g0 + tid]) * v : v * v;
          if (SS && last_effect) pb.rfull[row] = Rfull;
#pragma unroll
          for (int kk = 0; kk < K; ++kk)
            if (kk == k) {
              vsq[kk] += vterm;
              rsq[kk] += Rfull * Rfull;  // |y - X sum_l b_l|^2 at the last effect (infer.py:804)
            }
        }
        if (!SS) {
          __syncthreads();  // (B) r_next visible
          // sweep 2: acc[j] += sum_g X[g, j] * r_next[g]
          for (int gc = 0; gc < rows; gc += 8) {
            double rr[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) rr[u] = (gc + u < rows) ? H->rn[gc + u] : 0.0;
            for (int v = tid; v < pv; v += NT) {
              double a[VW];
#pragma unroll
              for (int u = 0; u < VW; u += 2) {
                const double2 t = *reinterpret_cast<const double2*>(acc + v * VW + u);
                a[u] = t.x;
                a[u + 1] = t.y;
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                if (gc + u < rows) {
                  double x[VW];
                  Vec<XT>::load(tile + static_cast<size_t>(gc + u) * p_pad, v, x);
#pragma unroll
                  for (int e = 0; e < VW; ++e) a[e] = fma(x[e], rr[u], a[e]);
                }
              }
#pragma unroll
              for (int u = 0; u < VW; u += 2)
                *reinterpret_cast<double2*>(acc + v * VW + u) = make_double2(a[u], a[u + 1]);
            }
          }
        }
        phase_bits ^= (1u << st);
        pending_refill = st;
        cons_seq += 1;
      }
      // tail: flush the last ancestry, refill the last consumed stage
      __syncthreads();
      if (tid == 0 && pending_refill >= 0) {
        issue_fill(prod_pos, pending_refill);
        prod_pos = (prod_pos + 1 == nb_own) ? 0 : prod_pos + 1;
      }
      if (!SS) {
        if (cur_k >= 0 && !big) {
          double* dst = my_scr + static_cast<size_t>(cur_k) * p_pad;
          for (int j = tid; j < p_pad; j += NT) __stcg(&dst[j], acc[j]);
        }
        zero_unowned_partials();
      }
      store_trav(l, first, vsq, rsq);
    };

    auto traverse = [&](int l, bool first, bool last_effect) {
      if constexpr (std::is_same<XT, uint8_t>::value) {
        traverse_b2(IntC<0>{}, l, first);
      } else if constexpr (std::is_same<XT, int8_t>::value && V == 5) {
        // narrow loci (p_pad <= 512): one warp per row, 16 columns per lane -- measured 111 / 117 us per update at
        // p = 300 / 500 against 154 / 157 for the column tile; wider loci are faster with the column tile
        // (p = 2000: 290 vs 357 us), profiles/r02b_shapes_i8_*.jsonl
        traverse_i8_rows(IntC<16>{}, IntC<1>{}, l, first, last_effect);
      } else if constexpr (std::is_same<XT, int8_t>::value) {
        if constexpr (V == 1)
          traverse_i8(IntC<1>{}, IntC<8>{}, IntC<SB_I8_KEEP>{}, IntC<8>{}, l, first, last_effect);
        else if constexpr (V == 2)
          traverse_i8(IntC<2>{}, IntC<4>{}, IntC<0>{}, IntC<8>{}, l, first, last_effect);
        else if constexpr (V == 3)
          traverse_i8(IntC<4>{}, IntC<2>{}, IntC<0>{}, IntC<8>{}, l, first, last_effect);
        else  // narrow loci (p_pad <= 1024): 4 columns x 16 rows per thread so every thread has work and bands are tall
          traverse_i8(IntC<1>{}, IntC<16>{}, IntC<1>{}, IntC<4>{}, l, first, last_effect);
      } else if constexpr (SS) {
        if (fast == 1)
          traverse_ss(IntC<4>{}, IntC<4>{}, l, last_effect);
        else if (fast == 2)
          traverse_ss(IntC<8>{}, IntC<2>{}, l, last_effect);
        else if (fast == 3)
          traverse_ss(IntC<12>{}, IntC<1>{}, l, last_effect);
        else if (fast == 4)
          traverse_ss(IntC<16>{}, IntC<1>{}, l, last_effect);
        else
          traverse_generic(IntC<0>{}, l, first, last_effect);
      } else {
        // the wide fp64 tile (12 vectors x 1 row) lives in its own kernel (V == 3): compiled next to the
        // others it wrecks the code generated for the common 4 x 4 tile (3x slower, measured)
        if constexpr (V == 3 && sizeof(XT) == 8) {
          traverse_fast(IntC<12>{}, IntC<1>{}, l, first, last_effect);
        } else if constexpr (V == 1 && sizeof(XT) == 8) {
          traverse_fast(IntC<4>{}, IntC<4>{}, l, first, last_effect);
        } else if constexpr (V == 5 && sizeof(XT) == 8) {
          traverse_fast(IntC<1>{}, IntC<8>{}, l, first, last_effect);  // narrow loci: p_pad <= 512
        } else if constexpr (V == 6 && sizeof(XT) == 8) {
          traverse_fast(IntC<2>{}, IntC<8>{}, l, first, last_effect);  // p_pad <= 1024
        } else {
          if (fast == 1)
            traverse_fast(IntC<4>{}, IntC<4>{}, l, first, last_effect);
          else if (fast == 2)
            traverse_fast(IntC<8>{}, IntC<2>{}, l, first, last_effect);
          else
            traverse_generic(IntC<0>{}, l, first, last_effect);
        }
      }
    };

    // ------------------------------------------------------------------------------
    // team combine of per-block (max, scaled sums) records written to `stat`:
    // rec[0] = m, rec[1..nv] = sums scaled by exp(-m).  Result in H->tot[0..nv] (tot[0] = m).
    auto team_combine = [&](int which, auto nv_c) {
      constexpr int NV = decltype(nv_c)::value;
      // every block holds its own record in H->tot already (tot[0]=m, tot[1..NV])
      if (C == 1) return;
      const int so = P.off_stat + which * STAT_STRIDE;
      if (tid <= NV) __stcg(&my_scr[so + tid], H->tot[tid]);
      team_sync(team);
      // one thread per team member (C <= NT): all records are fetched in one round trip
      double mc = -CUDART_INF, v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = 0.0;
      if (tid < C) {
        const double* rec = scr(tid) + so;
        mc = __ldcg(rec);
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = __ldcg(rec + 1 + i);
      }
      const double m = block_max<NW>(H, mc);
      if (tid < C) {
        // NaN anywhere must poison the result like the reference's softmax would
        const double sc = (mc == -CUDART_INF) ? 0.0 : exp(mc - m);
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = (mc != mc) ? mc : v[i] * sc;
      }
      block_sum<NV, NW>(H, v);
      if (tid == 0) {
        for (int i = NV - 1; i >= 0; --i) H->tot[i + 1] = H->tot[i];
        H->tot[0] = m;
      }
      __syncthreads();
    };

    // ------------------------------------------------------------------------------
    // posterior phases of effect l in iteration `it`: writes the un-normalised rows of out[buf]
    // (alpha slot = log weights, pm = mu_j, pmsq = S_j + mu_j mu_j^T, lbf) and the per-effect
    // scalars; the softmax normaliser (m, sum) is returned through H->tot[0], H->tot[1].
    // P1 (individual level): reduce the per-block partial X^T r over the team, own SNP slice
    auto reduce_partials = [&]() {
      double* __restrict__ xty_w = pb.xty;
      for (int j = js + tid; j < je; j += NT) {
        double s[K];
#pragma unroll
        for (int k = 0; k < K; ++k) s[k] = 0.0;
        for (int c = 0; c < C; ++c) {
          const double* __restrict__ part = scr(c);
#pragma unroll
          for (int k = 0; k < K; ++k) s[k] += __ldcg(&part[static_cast<size_t>(k) * p_pad + j]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) xty_w[k * p_pad + j] = s[k];
      }
    };

    auto posterior = [&](int l, int it, bool have_xty) {
      const int buf = it & 1;
      const OutBuf& ob = pb.out[buf];
      const double* ecov_prev = (it == 0) ? pb.ecov0 : pb.out[buf ^ 1].ecov;
      const double* rv_prev = (it == 0) ? pb.rv0 : pb.out[buf ^ 1].rv;
      double rv[K];
#pragma unroll
      for (int k = 0; k < K; ++k) rv[k] = __ldcg(&rv_prev[k]);

      if (!SS && !B2 && !have_xty) reduce_partials();
      // raw per-SNP inputs (X^T r, XtX, log pi); fetched one SNP ahead of the arithmetic
      constexpr int NRAW = 2 * K + 1;
      const double* __restrict__ xty_r = pb.xty;
      const double* __restrict__ xtx_r = pb.xtx;
      const double* __restrict__ lpi_r = pb.logpi;
      auto fetch = [&](int j, double (&raw)[NRAW]) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          raw[k] = SS ? __ldcg(&pb.r[pb.row0[k] + j]) : xty_r[k * p_pad + j];
          raw[K + k] = xtx_r[k * p_pad + j];
        }
        raw[2 * K] = lpi_r[j];
      };
      auto unpack = [&](const double (&raw)[NRAW], double (&d)[K], double (&r)[K]) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          r[k] = raw[k] / rv[k];      // rTZDinv   (infer.py:677)
          d[k] = raw[K + k] / rv[k];  // inv_shat2 (infer.py:680)
        }
      };
      auto load_cinv = [&](const double* Cmat, double (&Cinv)[K][K], double& logdetC) {
        __syncthreads();
        if (tid == 0) {
          double ci[K * K], lad;
          inv_logdet<K>(Cmat, ci, lad);
          for (int i = 0; i < K * K; ++i) H->cst[i] = ci[i];
          H->cst[K * K] = lad;
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
          for (int b = 0; b < K; ++b) Cinv[a][b] = H->cst[a * K + b];
        logdetC = H->cst[K * K];
      };

      double Cnew[K * K];
      // ---- pass A: posterior under the old prior, only to update the prior (infer.py:138,153)
      if (!pb.skip_first_pass) {
        double Cold[K * K];
#pragma unroll
        for (int i = 0; i < K * K; ++i) Cold[i] = __ldcg(&ecov_prev[l * K * K + i]);
        double Cinv[K][K], logdetC;
        load_cinv(Cold, Cinv, logdetC);
        double m = -CUDART_INF;
        double sv[1 + NSYM];
#pragma unroll
        for (int i = 0; i < 1 + NSYM; ++i) sv[i] = 0.0;
        double nxt[NRAW];
        if (js + tid < je) fetch(js + tid, nxt);
        for (int j = js + tid; j < je; j += NT) {
          double raw[NRAW];
#pragma unroll
          for (int i = 0; i < NRAW; ++i) raw[i] = nxt[i];
          if (j + NT < je) fetch(j + NT, nxt);
          double d[K], r[K], S[K][K], mu[K], lda;
          unpack(raw, d, r);
          snp_posterior<K>(Cinv, d, r, S, mu, lda);
          double q = 0.0;
#pragma unroll
          for (int k = 0; k < K; ++k) q += r[k] * mu[k];
          const double lbf = 0.5 * (K * LOG_2PI - lda + q);  // infer.py:712-719
          const double w = raw[2 * K] + lbf;
          double e;
          if (w > m) {  // online softmax: rescale the running sums
            const double sc = exp(m - w);
#pragma unroll
            for (int i = 0; i < 1 + NSYM; ++i) sv[i] *= sc;
            m = w;
            e = 1.0;
          } else {
            e = (w == -CUDART_INF) ? 0.0 : exp(w - m);  // NaN w propagates
          }
          sv[0] += e;
          int t = 1;
#pragma unroll
          for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = a; b < K; ++b) sv[t++] += e * (S[a][b] + mu[a] * mu[b]);
        }
        const double mb = block_max<NW>(H, m);
        const double sc = (m == -CUDART_INF) ? 0.0 : exp(m - mb);
#pragma unroll
        for (int i = 0; i < 1 + NSYM; ++i) sv[i] *= sc;
        block_sum<1 + NSYM, NW>(H, sv);
        if (tid == 0) {
          for (int i = NSYM; i >= 0; --i) H->tot[i + 1] = H->tot[i];
          H->tot[0] = mb;
        }
        __syncthreads();
        team_combine(0, IntC<1 + NSYM>{});
        const double ssum = H->tot[1];
        int t = 2;
#pragma unroll
        for (int a = 0; a < K; ++a)
#pragma unroll
          for (int b = a; b < K; ++b) {
            const double wv = H->tot[t++] / ssum;  // weighted_sum_covar (infer.py:726)
            Cnew[a * K + b] = wv;
            Cnew[b * K + a] = wv;
          }
        if (!pb.em) {
#pragma unroll
          for (int i = 0; i < K * K; ++i) Cnew[i] = Cnew[i] * pb.adj_times[i] + pb.adj_plus[i];  // infer.py:155-157
        }
      } else {
#pragma unroll
        for (int i = 0; i < K * K; ++i) Cnew[i] = pb.adj_plus[i];
      }

      // ---- pass B: posterior under the new prior; un-normalised rows + online statistics
      constexpr int NVB = 1 + NSYM + 2 + K;  // s, W, sum e*kl_beta, sum e*lbf, sum e*xtx*Mkk
      {
        double Cinv[K][K], logdetC;
        load_cinv(Cnew, Cinv, logdetC);
        double m = -CUDART_INF;
        double sv[NVB];
#pragma unroll
        for (int i = 0; i < NVB; ++i) sv[i] = 0.0;
        double nxt[NRAW];
        if (js + tid < je) fetch(js + tid, nxt);
        for (int j = js + tid; j < je; j += NT) {
          double raw[NRAW];
#pragma unroll
          for (int i = 0; i < NRAW; ++i) raw[i] = nxt[i];
          if (j + NT < je) fetch(j + NT, nxt);
          double d[K], r[K], S[K][K], mu[K], lda;
          unpack(raw, d, r);
          snp_posterior<K>(Cinv, d, r, S, mu, lda);
          double q = 0.0;
#pragma unroll
          for (int k = 0; k < K; ++k) q += r[k] * mu[k];
          const double lbf = 0.5 * (K * LOG_2PI - lda + q);
          const double w = raw[2 * K] + lbf;
          // KL(N(mu,S) || N(0,C))  (infer.py:755-783)
          double tr = 0.0, quad = 0.0;
#pragma unroll
          for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = 0; b < K; ++b) {
              tr += Cinv[a][b] * S[b][a];
              quad += mu[a] * Cinv[a][b] * mu[b];
            }
          const double klb = 0.5 * (tr - K + quad + logdetC + lda);
          // un-normalised rows (scaled by alpha once, when the locus is finished)
          const size_t lj = static_cast<size_t>(l) * p + j;
          __stcg(&ob.lbf[lj], lbf);
          __stcg(&ob.alpha[lj], w);
#pragma unroll
          for (int a = 0; a < K; ++a) {
            __stcg(&ob.pm[lj * K + a], mu[a]);
#pragma unroll
            for (int b = 0; b < K; ++b) __stcg(&ob.pmsq[(lj * K + a) * K + b], S[a][b] + mu[a] * mu[b]);
          }
          double e;
          if (w > m) {
            const double sc = exp(m - w);
#pragma unroll
            for (int i = 0; i < NVB; ++i) sv[i] *= sc;
            m = w;
            e = 1.0;
          } else {
            e = (w == -CUDART_INF) ? 0.0 : exp(w - m);
          }
          sv[0] += e;
          int t = 1;
#pragma unroll
          for (int a = 0; a < K; ++a)
#pragma unroll
            for (int b = a; b < K; ++b) sv[t++] += e * (S[a][b] + mu[a] * mu[b]);
          sv[t++] += e * klb;
          sv[t++] += e * lbf;
#pragma unroll
          for (int k = 0; k < K; ++k) sv[t++] += e * raw[K + k] * (S[k][k] + mu[k] * mu[k]);
        }
        const double mb = block_max<NW>(H, m);
        const double sc = (m == -CUDART_INF) ? 0.0 : exp(m - mb);
#pragma unroll
        for (int i = 0; i < NVB; ++i) sv[i] *= sc;
        block_sum<NVB, NW>(H, sv);
        if (tid == 0) {
          for (int i = NVB - 1; i >= 0; --i) H->tot[i + 1] = H->tot[i];
          H->tot[0] = mb;
        }
        __syncthreads();
        team_combine(1, IntC<NVB>{});
      }
      // per-effect scalars (one writer)
      {
        const double mt = H->tot[0], ssum = H->tot[1];
        if (rank == 0 && tid == 0) {
          int t = 2;
          for (int a = 0; a < K; ++a)
            for (int b = a; b < K; ++b) {
              const double wv = H->tot[t++] / ssum;
              ob.wsc[(l * K + a) * K + b] = wv;
              ob.wsc[(l * K + b) * K + a] = wv;
            }
          const double klb = H->tot[t++] / ssum;
          const double elbf = H->tot[t++] / ssum;
          // KL(alpha || pi) = sum alpha (log alpha - log pi) = E[lbf] - logsumexp  (infer.py:748-752)
          __stcg(&ob.kl[l], (elbf - (mt + log(ssum))) + klb);
          for (int k = 0; k < K; ++k) __stcg(&ob.xtxb2[l * K + k], H->tot[t++] / ssum);
          for (int i = 0; i < K * K; ++i) __stcg(&ob.ecov[l * K * K + i], Cnew[i]);
          __stcg(&ob.lse[2 * l], mt);
          __stcg(&ob.lse[2 * l + 1], ssum);
        }
        // b_l = softmax(w) * mu for the own SNP slice (infer.py:721-723), read back by the traversal
        const double inv_s = 1.0 / ssum;
        double* bl = pb.bvec + static_cast<size_t>(l) * K * p_pad;
        for (int j = js + tid; j < je; j += NT) {
          const size_t lj = static_cast<size_t>(l) * p + j;
          const double e = exp(__ldcg(&ob.alpha[lj]) - mt) * inv_s;
#pragma unroll
          for (int k = 0; k < K; ++k) __stcg(&bl[static_cast<size_t>(k) * p_pad + j], e * __ldcg(&ob.pm[lj * K + k]));
        }
      }
    };

    // ------------------------------------------------------------------------------
    // the IBSS loop
    int n_iter = 0, inc_flag = 1, final_buf = -1, n_ser = 0, it0 = 0;
    double elbo_last = -CUDART_INF;
    bool suspended = false;
    if (status0 == 0) {
      if (!SS) traverse(0, true, false);  // X^T y for the first effect
      if (rank == 0 && tid == 0) pb.elbo[0] = -CUDART_INF;
    } else {
      // resume a locus suspended at an iteration boundary: everything lives in global memory
      it0 = __ldcg(&pb.meta[0]);
      n_ser = __ldcg(&pb.meta[3]);
      elbo_last = __ldcg(&pb.elbo[it0]);
      n_iter = it0;
      final_buf = (it0 - 1) & 1;
    }
    team_sync(team);

    for (int it = it0; it < P.max_iter; ++it) {
      const int buf = it & 1;
      for (int l = 0; l < L; ++l) {
        posterior(l, it, status0 != 0 && it == it0 && l == 0);
        team_sync(team);  // b_l of every SNP slice visible to the whole team
        traverse(l, false, l + 1 == L);
        team_sync(team);
        n_ser += 1;
      }
      // ---- ELBO, residual variance, stop rule (infer.py:622-633, 517-548)
      const OutBuf& ob = pb.out[buf];
      const double* rv_prev = (it == 0) ? pb.rv0 : pb.out[buf ^ 1].rv;
      double* erss_sm = H->cst;  // [K]
      if (SS) {
        // terms 2 and 3 of infer_ss.py:555-575 from E[beta], Xty and the full residual
        double tv[2 * K];
#pragma unroll
        for (int i = 0; i < 2 * K; ++i) tv[i] = 0.0;
        for (int j = js + tid; j < je; j += NT) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            double eb = 0.0;
            for (int l = 0; l < L; ++l) eb += __ldcg(&effect_vec(l, k)[j]);
            const double xty = pb.y[pb.row0[k] + j];
            tv[k] += eb * xty;                                         // E[b]^T Xty
            tv[K + k] += eb * (xty - __ldcg(&pb.rfull[pb.row0[k] + j]));  // E[b]^T XtX E[b]
          }
        }
        block_sum<2 * K, NW>(H, tv);
        if (C > 1) {
          const int so = P.off_stat + 2 * STAT_STRIDE;
          if (tid < 2 * K) __stcg(&my_scr[so + tid], H->tot[tid]);
          team_sync(team);
#pragma unroll
          for (int i = 0; i < 2 * K; ++i) tv[i] = (tid < C) ? __ldcg(&scr(tid)[so + i]) : 0.0;
          block_sum<2 * K, NW>(H, tv);
        }
      }
      double ssv[2 * K];  // summary level: E[b]^T Xty, E[b]^T XtX E[b] per ancestry
#pragma unroll
      for (int i = 0; i < 2 * K; ++i) ssv[i] = SS ? H->tot[i] : 0.0;
      // per (l, k) traversal scalars summed over the team: one thread per team member, fixed order
      {
        double tr[2 * K];
#pragma unroll
        for (int i = 0; i < 2 * K; ++i) tr[i] = 0.0;
        if (tid < C) {
          const double* rec = scr(tid) + P.off_trav;
          for (int l = 0; l < L; ++l) {
#pragma unroll
            for (int k = 0; k < K; ++k) tr[k] += __ldcg(&rec[l * TRAV_STRIDE + k]);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) tr[K + k] = __ldcg(&rec[(L - 1) * TRAV_STRIDE + KMAX + k]);
        }
        block_sum<2 * K, NW>(H, tr);
      }
      if (tid < K) {
        const int k = tid;
        double t5 = 0.0;  // sum_l sum_p XtX E[b_l^2]
        for (int l = 0; l < L; ++l) t5 += __ldcg(&ob.xtxb2[l * K + k]);
        double e;
        if (SS)
          e = pb.nsamp[k] - 2.0 * ssv[k] + ssv[K + k] - H->tot[k] + t5;  // infer_ss.py:555-575
        else
          e = H->tot[K + k] + (t5 - H->tot[k]);  // infer.py:799-808
        erss_sm[k] = e;
      }
      __syncthreads();
      if (tid == 0) {
        double exp_ll = 0.0;
        for (int k = 0; k < K; ++k) {
          const double s2 = __ldcg(&rv_prev[k]);
          exp_ll += -(0.5 * pb.nsamp[k]) * log(2.0 * CUDART_PI * s2) - (0.5 / s2) * erss_sm[k];
        }
        double kls = 0.0;
        for (int l = 0; l < L; ++l) kls += __ldcg(&ob.kl[l]);
        const double cur = exp_ll - kls;
        // elbo_increase = cur >= last or isclose(cur, last, rtol=1e-5, atol=1e-8)
        const bool inc = (cur >= elbo_last) || (fabs(cur - elbo_last) <= (1e-8 + 1e-5 * fabs(elbo_last)));
        int dec = 0;  // 0 continue, 1 converged, 2 ELBO decreased -> revert
        if (!inc)
          dec = 2;
        else if (fabs(cur - elbo_last) < P.min_tol)
          dec = 1;
        H->decision = dec;
        H->cst[KMAX] = cur;
        if (rank == 0) {
          pb.elbo[it + 1] = cur;
          for (int k = 0; k < K; ++k) __stcg(&ob.rv[k], erss_sm[k] / pb.nsamp[k]);
        }
      }
      __syncthreads();
      const int dec = H->decision;
      elbo_last = H->cst[KMAX];
      n_iter = it + 1;
      if (dec == 2) {
        inc_flag = 0;
        final_buf = (it == 0) ? -1 : (buf ^ 1);
        break;
      }
      final_buf = buf;
      if (dec == 1) break;
      // drain request (half of the teams idle, larger teams available): one consistent read per team
      if (rank == 0 && tid == 0)
        __stcg(&P.team_slot[P.n_teams + team.id], (P.allow_drain && it + 1 < P.max_iter) ? atomicAdd(&P.ctl[1], 0) : 0);
      team_sync(team);  // new residual variances (and the drain decision) visible to every block
      if (__ldcg(&P.team_slot[P.n_teams + team.id])) {
        // suspend: reduce the pending X^T r now so the next phase may use any team size
        if (!SS && !B2) reduce_partials();
        suspended = true;
        break;
      }
    }
    if (rank == 0 && tid == 0) {
      pb.meta[0] = n_iter;
      pb.meta[1] = inc_flag;
      pb.meta[2] = final_buf;
      pb.meta[3] = n_ser;
      pb.meta[4] = suspended ? 1 : 2;
      pb.meta[5] = (!inc_flag && !(elbo_last == elbo_last)) ? 1 : 0;  // SB_STATUS_NONFINITE: the ELBO became NaN
    }
    // ---- normalise the returned buffer once: alpha = softmax(w), alpha-weighted mean / second moment
    if (!suspended && final_buf >= 0) {
      const OutBuf& ob = pb.out[final_buf];
      for (int l = 0; l < L; ++l) {
        const double mt = __ldcg(&ob.lse[2 * l]), ssum = __ldcg(&ob.lse[2 * l + 1]);
        for (int j = js + tid; j < je; j += NT) {
          const size_t lj = static_cast<size_t>(l) * p + j;
          const double al = exp(__ldcg(&ob.alpha[lj]) - mt) / ssum;  // softmax (infer.py:721)
          ob.alpha[lj] = al;
#pragma unroll
          for (int a = 0; a < K; ++a) {
            ob.pm[lj * K + a] = __ldcg(&ob.pm[lj * K + a]) * al;
#pragma unroll
            for (int b = 0; b < K; ++b)
              ob.pmsq[(lj * K + a) * K + b] = __ldcg(&ob.pmsq[(lj * K + a) * K + b]) * al;
          }
        }
      }
    }
    // drain the speculative prefetches so the ring can be re-pointed at the next locus
    if (nb_own > 0) {
      for (int s = 0; s < NS; ++s) {
        const int st = (cons_seq + s) % NS;
        mbar_wait(&H->mbar[st], (phase_bits >> st) & 1u);
        phase_bits ^= (1u << st);
      }
    }
    team_sync(team);
  }
}

// ---------------------------------------------------------------------------------------
// prologue kernels

// Column statistics + standardisation: in [n][p] (any input type) -> out [n][p_pad] storage,
// xtx[j] = sum_n out^2.  One thread per column (coalesced across the warp).
// Restates infer.py:326-333 (centre, scale by the population std) and infer.py:494.
template <typename IT, typename XT>
__global__ void prep_columns_kernel(const IT* in, int in_ld, int n, int p, int p_pad, int flags, XT* out,
                                    double* __restrict__ xtx) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p_pad) return;
  if (j >= p) {
    for (int i = 0; i < n; ++i) out[static_cast<size_t>(i) * p_pad + j] = XT(0);
    xtx[j] = 0.0;
    return;
  }
  double mean = 0.0, sd = 1.0;
  if (flags & 1) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += static_cast<double>(in[static_cast<size_t>(i) * in_ld + j]);
    mean = s / n;
  }
  if (flags & 2) {
    // numpy.std of the centred column: sqrt(mean(|x - mean(x)|^2))
    double s1 = 0.0;
    for (int i = 0; i < n; ++i) s1 += static_cast<double>(in[static_cast<size_t>(i) * in_ld + j]) - mean;
    const double m2 = s1 / n;
    double s2 = 0.0;
    for (int i = 0; i < n; ++i) {
      const double c = (static_cast<double>(in[static_cast<size_t>(i) * in_ld + j]) - mean) - m2;
      s2 += c * c;
    }
    sd = sqrt(s2 / n);
  }
  double ss = 0.0;
  for (int i = 0; i < n; ++i) {
    double v = static_cast<double>(in[static_cast<size_t>(i) * in_ld + j]) - mean;
    if (flags & 2) v /= sd;
    const XT sv = static_cast<XT>(v);
    out[static_cast<size_t>(i) * p_pad + j] = sv;
    const double back = static_cast<double>(sv);
    ss += back * back;
  }
  xtx[j] = ss;
}

// Batched column prologue: one launch standardises many (locus, ancestry) matrices (blockIdx.y picks the
// descriptor).  Per-matrix launches of a 16-block grid were the largest host-visible cost of the end-to-end
// path at batch throughput.  Same arithmetic as prep_columns_kernel / prep_columns_i8_kernel.
struct PrepDesc {
  const void* in;   // [n][in_ld] input type IT
  void* out;        // [n][p_pad] storage type XT
  double* xtx;      // [p_pad]
  double* cmean;    // [p_pad] (hard-call storage only)
  double* cisd;     // [p_pad] (hard-call storage only)
  int in_ld, n, p, p_pad, flags, qc;
  const double* Q;  // [n][qc] covariate basis (hard-call storage, qc > 0) or nullptr
  double* M;        // [qc][p_pad] out: G^T Q
  const int* rows;  // row view: row i of this matrix is row rows[i] of `in` (cross-validation folds share the
                    // staged copy of their gene's matrix); nullptr = rows 0..n-1
};

template <typename IT, typename XT>
__global__ void prep_columns_batched_kernel(const PrepDesc* __restrict__ descs, int* __restrict__ bad_flag) {
  const PrepDesc d = descs[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d.p_pad) return;
  constexpr bool HARD = std::is_same<XT, int8_t>::value;
  const IT* in = static_cast<const IT*>(d.in);
  XT* out = static_cast<XT*>(d.out);
  const int n = d.n, in_ld = d.in_ld, p_pad = d.p_pad;
  const int* __restrict__ rows = d.rows;
  auto src = [&](int i) -> IT { return in[static_cast<size_t>(rows ? rows[i] : i) * in_ld + j]; };
  if (j >= d.p) {
    for (int i = 0; i < n; ++i) out[static_cast<size_t>(i) * p_pad + j] = XT(0);
    d.xtx[j] = 0.0;
    if (HARD) {
      d.cmean[j] = 0.0;
      d.cisd[j] = 0.0;
      for (int t = 0; t < d.qc; ++t) d.M[static_cast<size_t>(t) * p_pad + j] = 0.0;
    }
    return;
  }
  if (HARD && d.qc > 0) {
    // covariate-residualised column (utils.py:114-136 then infer.py:326-333): x = g - Q (Q^T g); the intercept is a
    // column of Q, so x is centred; sd = numpy.std(x), XtX = sum (x / sd)^2
    const int qc = d.qc;
    double m[B2_QMAX];
    for (int t = 0; t < B2_QMAX; ++t) m[t] = 0.0;
    bool bad = false;
    for (int i = 0; i < n; ++i) {
      const IT g = src(i);
      out[static_cast<size_t>(i) * p_pad + j] = static_cast<XT>(g);
      if (g < IT(0)) bad = true;
      const double gd = static_cast<double>(g);
      for (int t = 0; t < qc; ++t) m[t] += gd * d.Q[static_cast<size_t>(i) * qc + t];
    }
    for (int t = 0; t < qc; ++t) d.M[static_cast<size_t>(t) * p_pad + j] = m[t];
    double s1 = 0.0;
    for (int i = 0; i < n; ++i) {
      double x = static_cast<double>(src(i));
      for (int t = 0; t < qc; ++t) x -= d.Q[static_cast<size_t>(i) * qc + t] * m[t];
      s1 += x;
    }
    const double mu = s1 / n;  // rounding-level mean the reference subtracts again (infer.py:327)
    double sd = 1.0;
    if (d.flags & 2) {
      double s2 = 0.0;
      for (int i = 0; i < n; ++i) {
        double x = static_cast<double>(src(i));
        for (int t = 0; t < qc; ++t) x -= d.Q[static_cast<size_t>(i) * qc + t] * m[t];
        const double c = x - mu;
        s2 += c * c;
      }
      sd = sqrt(s2 / n);
    }
    double ss = 0.0;
    for (int i = 0; i < n; ++i) {
      double x = static_cast<double>(src(i));
      for (int t = 0; t < qc; ++t) x -= d.Q[static_cast<size_t>(i) * qc + t] * m[t];
      const double v = (x - mu) / sd;
      ss += v * v;
    }
    d.xtx[j] = ss;
    d.cmean[j] = 0.0;
    d.cisd[j] = 1.0 / sd;
    if (bad) atomicOr(bad_flag, 1);
    return;
  }
  double mean = 0.0, sd = 1.0;
  if (d.flags & 1) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += static_cast<double>(src(i));
    mean = s / n;
  }
  if (d.flags & 2) {
    // numpy.std of the centred column: sqrt(mean(|x - mean(x)|^2))
    double s1 = 0.0;
    for (int i = 0; i < n; ++i) s1 += static_cast<double>(src(i)) - mean;
    const double m2 = s1 / n;
    double s2 = 0.0;
    for (int i = 0; i < n; ++i) {
      const double c = (static_cast<double>(src(i)) - mean) - m2;
      s2 += c * c;
    }
    sd = sqrt(s2 / n);
  }
  double ss = 0.0;
  bool bad = false;
  for (int i = 0; i < n; ++i) {
    const IT g = src(i);
    double v = static_cast<double>(g) - mean;
    if (d.flags & 2) v /= sd;
    if (HARD) {
      out[static_cast<size_t>(i) * p_pad + j] = static_cast<XT>(g);
      if (g < IT(0)) bad = true;  // the traversal widens the bytes as unsigned
      ss += v * v;
    } else {
      const XT sv = static_cast<XT>(v);
      out[static_cast<size_t>(i) * p_pad + j] = sv;
      const double back = static_cast<double>(sv);
      ss += back * back;
    }
  }
  d.xtx[j] = ss;
  if (HARD) {
    d.cmean[j] = mean;
    d.cisd[j] = 1.0 / sd;
    if (bad) atomicOr(bad_flag, 1);
  }
}

// Bit-plane packing of hard calls 0/1/2 (see traverse_b2): from the resident one-byte matrix [n][p_pad] to
//   rbits = [chunk][row quad][plane][32 groups][4 rows]           bit c of a byte = column 256 chunk + 8 group + c
//   cbits = [row chunk][column quad][plane][32 groups][4 columns]   bit r of a byte = row 256 chunk + 8 group + r
// plane 0 = (g >= 1), plane 1 = (g == 2); rows / columns past the end are zero.
struct PackDesc {
  const signed char* X;  // [n][p_pad]
  unsigned char* rbits;
  unsigned char* cbits;
  int n, p_pad, ntiles, pad_;
};

static __global__ void pack_b2_kernel(const PackDesc* __restrict__ descs, int* __restrict__ bad_flag) {
  const PackDesc d = descs[blockIdx.y];
  const int nq = (d.n + 3) >> 2, nch = (d.n + 255) >> 8, ncq = 64 * d.ntiles;
  const long long n_r = static_cast<long long>(d.ntiles) * nq * 32;
  const long long n_c = static_cast<long long>(nch) * ncq * 32;
  bool bad = false;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n_r + n_c;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (idx < n_r) {
      const int g = static_cast<int>(idx & 31);
      const long long t = idx >> 5;
      const int quad = static_cast<int>(t % nq), c = static_cast<int>(t / nq);
      const int col0 = c * B2_TILE_COLS + 8 * g;
      uint32_t w0 = 0u, w1 = 0u;
      if (col0 < d.p_pad) {
        for (int rr = 0; rr < 4; ++rr) {
          const int row = 4 * quad + rr;
          if (row < d.n) {
            const uint2 v8 = *reinterpret_cast<const uint2*>(d.X + static_cast<size_t>(row) * d.p_pad + col0);
            for (int cc = 0; cc < 8; ++cc) {
              const uint32_t v = (((cc < 4) ? v8.x : v8.y) >> (8 * (cc & 3))) & 0xffu;
              if (v > 2u) bad = true;
              w0 |= static_cast<uint32_t>(v >= 1u) << (8 * rr + cc);
              w1 |= static_cast<uint32_t>(v == 2u) << (8 * rr + cc);
            }
          }
        }
      }
      uint32_t* out = reinterpret_cast<uint32_t*>(d.rbits) + ((static_cast<size_t>(c) * nq + quad) * 2) * 32 + g;
      out[0] = w0;
      out[32] = w1;
    } else {
      // thread = (row chunk, column quad, row group): 8 rows x 4 columns -> one word per plane
      const long long i2 = idx - n_r;
      const int rg = static_cast<int>(i2 & 31);
      const long long t = i2 >> 5;
      const int cq = static_cast<int>(t % ncq), ch = static_cast<int>(t / ncq);
      const int j0 = 4 * cq, row0 = 256 * ch + 8 * rg;
      uint32_t w0 = 0u, w1 = 0u;
      if (j0 < d.p_pad) {
        for (int r = 0; r < 8; ++r) {
          const int row = row0 + r;
          if (row < d.n) {
            const uint32_t v4 = *reinterpret_cast<const uint32_t*>(d.X + static_cast<size_t>(row) * d.p_pad + j0);
            for (int cc = 0; cc < 4; ++cc) {
              const uint32_t v = (v4 >> (8 * cc)) & 0xffu;
              w0 |= static_cast<uint32_t>(v >= 1u) << (8 * cc + r);
              w1 |= static_cast<uint32_t>(v == 2u) << (8 * cc + r);
            }
          }
        }
      }
      uint32_t* out = reinterpret_cast<uint32_t*>(d.cbits) + ((static_cast<size_t>(ch) * ncq + cq) * 2) * 32 + rg;
      out[0] = w0;
      out[32] = w1;
    }
  }
  if (bad) atomicOr(bad_flag, 2);
}

// Summary-level prologue (infer_ss.py:340-344): Xty = sqrt(n) sqrt(n/(n+z^2)) z,
// diag(XtX) = n * LD_jj, log pi.
template <typename XT>
__global__ void prep_ss_kernel(const XT* __restrict__ ld, const double* __restrict__ z, double ns, int p,
                               int p_pad, double* __restrict__ xty, double* __restrict__ xtx) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p_pad) return;
  if (j >= p) {
    xtx[j] = 0.0;
    return;
  }
  const double zz = z[j];
  const double s2 = ns / (ns + zz * zz);
  xty[j] = sqrt(ns) * sqrt(s2) * zz;
  xtx[j] = ns * static_cast<double>(ld[static_cast<size_t>(j) * p_pad + j]);
}

// log pi (infer.py:721) of every problem of a batch (blockIdx.y = problem); no user weights means the uniform
// default 1/p (infer.py:302-303)
static __global__ void log_pi_batched_kernel(const ProbDev* __restrict__ probs) {
  const ProbDev& P = probs[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.p) P.logpi[i] = log(P.pi_raw ? P.pi_raw[i] : 1.0 / static_cast<double>(P.p));
}

// Pad-copy rows: in [n][p] -> out [n][p_pad] with zero tail, optional type conversion.
template <typename IT, typename XT>
__global__ void pad_rows_kernel(const IT* __restrict__ in, int n, int p, int p_pad, XT* __restrict__ out) {
  const size_t total = static_cast<size_t>(n) * p_pad;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t i = idx / p_pad;
    const int j = static_cast<int>(idx - i * p_pad);
    out[idx] = (j < p) ? static_cast<XT>(in[i * p + j]) : XT(0);
  }
}

// Purity (infer.py:946-956): min |X_sel^T X_sel| / n over all pairs of a credible set.
// Credible-set purity (infer.py:936-965) for the sets of many loci in one launch: block (set, ancestry) returns
// min |X_sel^T X_sel| / n over the m x m Gram matrix of the set's columns (summary level: min |LD_sel|).
// The selected columns are gathered -- and, for hard-call storage, standardised -- into shared memory `rows`
// rows at a time (32, fewer for sets too large for that); every thread owns 4 x 4 tiles of the upper triangle
// and accumulates them in registers over the rows in row order (so the sums do not depend on the tiling).
constexpr int PUR_ROWS = 32;

// min that propagates NaN like jnp.min (infer.py:954-963): a NaN correlation (monomorphic column, NaN LD entry)
// must make the purity NaN so that `purity > threshold` is false and the set is dropped, as in the reference.
__device__ __forceinline__ double nan_min(double a, double b) { return (a != a || b != b) ? CUDART_NAN : fmin(a, b); }

__host__ __device__ inline size_t purity_smem_bytes(int m, int rows) {
  const size_t m_pad = static_cast<size_t>((m + 3) / 4) * 4;
  return m_pad * (sizeof(double) * (rows + 2 + B2_QMAX) + sizeof(int));
}

template <typename XT, bool SS>
__global__ void __launch_bounds__(256) purity_kernel(const ProbDev* __restrict__ probs,
                                                     const int* __restrict__ set_prob,
                                                     const int* __restrict__ idx, const int* __restrict__ offsets,
                                                     int m_pad_max, int rows, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char pur_smem[];
  const int set = blockIdx.x, k = blockIdx.y;
  const ProbDev& P = probs[set_prob[set]];
  if (k >= P.K) return;
  const int o0 = offsets[set], m = offsets[set + 1] - o0;
  const int* sel = idx + o0;
  const XT* X = reinterpret_cast<const XT*>(P.X[k]);
  const int n = P.n[k], p_pad = P.p_pad;
  double best = CUDART_INF;
  if (SS) {
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int a = e / m, b = e % m;
      best = nan_min(best, fabs(static_cast<double>(X[static_cast<size_t>(sel[a]) * p_pad + sel[b]])));
    }
  } else {
    constexpr bool I8 = sizeof(XT) == 1;
    const int m4 = (m + 3) / 4, m_pad = m4 * 4;
    double* tile = reinterpret_cast<double*>(pur_smem);           // [rows][m_pad]
    double* c_mean = tile + static_cast<size_t>(rows) * m_pad_max;  // [m_pad]
    double* c_isd = c_mean + m_pad_max;                            // [m_pad]
    double* c_mq = c_isd + m_pad_max;                              // [B2_QMAX][m_pad] rows of M = G^T Q (covariates)
    int* c_sel = reinterpret_cast<int*>(c_mq + static_cast<size_t>(B2_QMAX) * m_pad_max);  // [m_pad]
    const int qc = I8 ? P.qc[k] : 0;
    const double* Qb = I8 ? P.Qb[k] : nullptr;
    for (int c = threadIdx.x; c < m_pad; c += blockDim.x) {
      const int j = c < m ? sel[c] : 0;
      c_sel[c] = j;
      for (int t = 0; t < qc; ++t) c_mq[static_cast<size_t>(t) * m_pad_max + c] = (c < m) ? P.Mq[k][static_cast<size_t>(t) * p_pad + j] : 0.0;
      c_mean[c] = (I8 && c < m) ? P.cmean[static_cast<size_t>(k) * p_pad + j] : 0.0;
      c_isd[c] = (I8 && c < m) ? P.cisd[static_cast<size_t>(k) * p_pad + j] : (c < m ? 1.0 : 0.0);
    }
    const int ntiles = m4 * (m4 + 1) / 2;
    for (int base = 0; base < ntiles; base += blockDim.x) {
      const int t = base + threadIdx.x;
      const bool valid = t < ntiles;
      int ta = 0, tb = 0;
      if (valid) {
        // row ta of the upper triangle of the m4 x m4 tile grid: first index with ta*(2*m4-ta+1)/2 <= t
        ta = static_cast<int>((2.0 * m4 + 1.0 - sqrt((2.0 * m4 + 1.0) * (2.0 * m4 + 1.0) - 8.0 * t)) * 0.5);
        while (ta * (2 * m4 - ta + 1) / 2 > t) --ta;
        while ((ta + 1) * (2 * m4 - ta) / 2 <= t) ++ta;
        tb = ta + (t - ta * (2 * m4 - ta + 1) / 2);
      }
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
      for (int r0 = 0; r0 < n; r0 += rows) {
        const int rc = min(rows, n - r0);
        __syncthreads();
        for (int e = threadIdx.x; e < rc * m_pad; e += blockDim.x) {
          const int rr = e / m_pad, c = e - rr * m_pad;
          double val = static_cast<double>(X[static_cast<size_t>(r0 + rr) * p_pad + c_sel[c]]);
          if (I8 && qc > 0) {  // hard calls residualised on covariates: x = g - Q (Q^T g)
            for (int t = 0; t < qc; ++t) val -= Qb[static_cast<size_t>(r0 + rr) * qc + t] * c_mq[static_cast<size_t>(t) * m_pad_max + c];
            val *= c_isd[c];
          } else if (I8) val = (val - c_mean[c]) * c_isd[c];  // hard-call storage: standardise on the fly
          else val *= c_isd[c];                        // zero for the padding columns
          tile[rr * m_pad + c] = val;
        }
        __syncthreads();
        if (valid) {
          const double* ra = tile + 4 * ta;
          const double* rb = tile + 4 * tb;
          for (int rr = 0; rr < rc; ++rr) {
            const double2 a01 = *reinterpret_cast<const double2*>(ra + rr * m_pad);
            const double2 a23 = *reinterpret_cast<const double2*>(ra + rr * m_pad + 2);
            const double2 b01 = *reinterpret_cast<const double2*>(rb + rr * m_pad);
            const double2 b23 = *reinterpret_cast<const double2*>(rb + rr * m_pad + 2);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y};
            const double bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
          }
        }
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (4 * ta + i < m && 4 * tb + j < m) best = nan_min(best, fabs(acc[i][j] / n));
      }
    }
  }
  // block min
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) best = nan_min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = red[0];
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) b = nan_min(b, red[w]);
    out[set * KMAX + k] = b;
  }
}


// ---------------------------------------------------------------------------------------------------------
// Purity for hard-call storage by EXACT integer Gram matrices (infer.py:946-965).  The m selected one-byte
// columns of block (set, ancestry) are gathered ONCE into shared memory as words of four consecutive rows,
// laid out [row quad][column] so that the four columns of a 4 x 4 tile are one 128-bit load; every pair of
// columns then costs n / 4 `dp4a` instructions instead of n fp64 multiply-adds on widened genotypes (and the
// columns are no longer re-gathered from global memory once per pass over the tiles).  With S_ab = sum_i g_ia g_ib
// and s_a = sum_i g_ia exact,
//     sum_i x_ia x_ib = (S_ab - mu_a s_b - mu_b s_a + n mu_a mu_b) / (sd_a sd_b)          x = (g - mu) / sd
//     sum_i x_ia x_ib = (S_ab - sum_t M_ta M_tb) / (sd_a sd_b)       x = (g - Q Q^T g) / sd, M = G^T Q (covariates)
// which is the reference's X_sel^T X_sel on the standardised columns up to fp64 rounding of the last step.
__host__ __device__ inline size_t purity_i8_smem_bytes(int m_pad, int nq) {
  return static_cast<size_t>(m_pad) * nq * 4 + static_cast<size_t>(m_pad) * (sizeof(double) * (3 + B2_QMAX) + sizeof(int));
}

static __global__ void __launch_bounds__(256) purity_i8_kernel(const ProbDev* __restrict__ probs,
                                                        const int* __restrict__ set_prob,
                                                        const int* __restrict__ idx, const int* __restrict__ offsets,
                                                        int m_pad_max, int nq_max, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char pur_smem[];
  const int set = blockIdx.x, k = blockIdx.y;
  const ProbDev& P = probs[set_prob[set]];
  if (k >= P.K) return;
  const int o0 = offsets[set], m = offsets[set + 1] - o0;
  const int* sel = idx + o0;
  const signed char* X = reinterpret_cast<const signed char*>(P.X[k]);
  const int n = P.n[k], p_pad = P.p_pad;
  const int m4 = (m + 3) / 4, m_pad = m4 * 4, nq = (n + 3) / 4;
  unsigned* W = reinterpret_cast<unsigned*>(pur_smem);  // [nq][m_pad]
  double* c_mean = reinterpret_cast<double*>(pur_smem + static_cast<size_t>(m_pad_max) * nq_max * 4);  // [m_pad]
  double* c_isd = c_mean + m_pad_max;
  double* c_sum = c_isd + m_pad_max;
  double* c_mq = c_sum + m_pad_max;  // [B2_QMAX][m_pad_max]
  int* c_sel = reinterpret_cast<int*>(c_mq + static_cast<size_t>(B2_QMAX) * m_pad_max);
  const int qc = P.qc[k];
  for (int c = threadIdx.x; c < m_pad; c += blockDim.x) {
    const int j = c < m ? sel[c] : 0;
    c_sel[c] = j;
    for (int t = 0; t < qc; ++t) c_mq[static_cast<size_t>(t) * m_pad_max + c] = (c < m) ? P.Mq[k][static_cast<size_t>(t) * p_pad + j] : 0.0;
    c_mean[c] = (c < m) ? P.cmean[static_cast<size_t>(k) * p_pad + j] : 0.0;
    c_isd[c] = (c < m) ? P.cisd[static_cast<size_t>(k) * p_pad + j] : 0.0;
  }
  __syncthreads();
  // gather: word (q, c) = rows 4q .. 4q+3 of column sel[c]; adjacent threads read neighbouring columns of one row
  for (int e = threadIdx.x; e < nq * m_pad; e += blockDim.x) {
    const int q = e / m_pad, c = e - q * m_pad;
    unsigned w = 0;
    if (c < m) {
      const signed char* col = X + c_sel[c];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int r = 4 * q + t;
        if (r < n) w |= static_cast<unsigned>(static_cast<unsigned char>(col[static_cast<size_t>(r) * p_pad])) << (8 * t);
      }
    }
    W[e] = w;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < m_pad; c += blockDim.x) {  // exact column sums
    int sum = 0;
    for (int q = 0; q < nq; ++q) sum = __dp4a(static_cast<int>(W[q * m_pad + c]), 0x01010101, sum);
    c_sum[c] = static_cast<double>(sum);
  }
  __syncthreads();
  double best = CUDART_INF;
  const int ntiles = m4 * (m4 + 1) / 2;
  const double nd = static_cast<double>(n);
  for (int t = threadIdx.x; t < ntiles; t += blockDim.x) {
    // row ta of the upper triangle of the m4 x m4 tile grid: first index with ta*(2*m4-ta+1)/2 <= t
    int ta = static_cast<int>((2.0 * m4 + 1.0 - sqrt((2.0 * m4 + 1.0) * (2.0 * m4 + 1.0) - 8.0 * t)) * 0.5);
    while (ta * (2 * m4 - ta + 1) / 2 > t) --ta;
    while ((ta + 1) * (2 * m4 - ta) / 2 <= t) ++ta;
    const int tb = ta + (t - ta * (2 * m4 - ta + 1) / 2);
    int acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    const uint4* wa = reinterpret_cast<const uint4*>(W) + ta;
    const uint4* wb = reinterpret_cast<const uint4*>(W) + tb;
#pragma unroll 4
    for (int q = 0; q < nq; ++q) {
      const uint4 a = wa[q * m4], b = wb[q * m4];
      const int av[4] = {static_cast<int>(a.x), static_cast<int>(a.y), static_cast<int>(a.z), static_cast<int>(a.w)};
      const int bv[4] = {static_cast<int>(b.x), static_cast<int>(b.y), static_cast<int>(b.z), static_cast<int>(b.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __dp4a(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ca = 4 * ta + i, cb = 4 * tb + j;
        if (ca < m && cb < m) {
          double g = static_cast<double>(acc[i][j]);
          if (qc > 0) {
            for (int tq = 0; tq < qc; ++tq)
              g -= c_mq[static_cast<size_t>(tq) * m_pad_max + ca] * c_mq[static_cast<size_t>(tq) * m_pad_max + cb];
          } else {
            g = g - c_mean[ca] * c_sum[cb] - c_mean[cb] * c_sum[ca] + nd * c_mean[ca] * c_mean[cb];
          }
          // a monomorphic column has sd = 0: the reference's standardised column is 0 / 0 = NaN and so is every
          // correlation with it (here 1 / sd = inf must not meet a rounding-level g != 0 and turn into +-inf)
          const double sa = c_isd[ca], sc = c_isd[cb];
          const bool degenerate = isinf(sa) || isinf(sc) || sa != sa || sc != sc;
          best = nan_min(best, degenerate ? CUDART_NAN : fabs(g * sa * sc / nd));
        }
      }
  }
  __shared__ double red_i8[32];
  for (int o = 16; o > 0; o >>= 1) best = nan_min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) red_i8[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = red_i8[0];
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) b = nan_min(b, red_i8[w]);
    out[set * KMAX + k] = b;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Credible-set selection on the device (reference make_cs, infer.py:899-915): block (effect l, problem) sorts the
// p posterior inclusion probabilities alpha_l in descending order (ties: ascending SNP index) with a bitonic
// network in shared memory (keys = the bit patterns of the non-negative doubles), then ONE thread accumulates the
// running sum left to right -- the order of additions of pandas' cumsum, so the cut `csum < threshold` falls where
// the reference's does -- and the block writes order / csum / set size.  `scratch` != nullptr: loci too wide for
// shared memory sort in global memory (one slot of n_pow2 keys + indices per block).
constexpr int SEL_THREADS = 512;

struct SelectDesc {
  int* order;    // [L][p]
  double* csum;  // [L][p]
  int* sizes;    // [L]
};

__device__ __forceinline__ bool sel_before(unsigned long long ka, int ia, unsigned long long kb, int ib) {
  return ka > kb || (ka == kb && ia < ib);
}

static __global__ void __launch_bounds__(SEL_THREADS) select_sets_kernel(const ProbDev* __restrict__ probs,
                                                                  const SelectDesc* __restrict__ descs,
                                                                  const int* __restrict__ prob_ids,
                                                                  const double* __restrict__ thresholds,
                                                                  unsigned char* __restrict__ scratch,
                                                                  long long scratch_stride) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  __shared__ int s_special, s_count;
  const int l = blockIdx.x, pi = prob_ids[blockIdx.y];
  const ProbDev& P = probs[pi];
  if (l >= P.L) return;
  const SelectDesc D = descs[pi];
  const double threshold = thresholds[pi];
  const int p = P.p;
  int N = 1;
  while (N < p) N <<= 1;
  unsigned char* base = scratch ? scratch + (static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * scratch_stride
                                : sel_smem;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base);
  int* ids = reinterpret_cast<int*>(base + sizeof(unsigned long long) * N);
  const int fb = P.meta[2];  // final output buffer; -1: the initial state is returned (alpha = 0)
  const double* alpha = fb >= 0 ? P.out[fb].alpha + static_cast<size_t>(l) * p : nullptr;
  if (threadIdx.x == 0) {
    s_special = 0;
    s_count = 0;
  }
  __syncthreads();
  bool special = false;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    unsigned long long key = 0ull;
    int id = 0x7fffffff;  // padding sorts behind every real entry
    if (i < p) {
      const double a = alpha ? alpha[i] : 0.0;
      id = i;
      if (a > 0.0) key = static_cast<unsigned long long>(__double_as_longlong(a));
      else if (a != 0.0) special = true;  // NaN or negative: sorted as the smallest, summed with its true value
    }
    keys[i] = key;
    ids[i] = id;
  }
  if (special) s_special = 1;
  __syncthreads();
  for (int kk = 2; kk <= N; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int q = i | j;
        const unsigned long long ka = keys[i], kb = keys[q];
        const int ia = ids[i], ib = ids[q];
        const bool up = (i & kk) == 0;
        const bool swap = up ? sel_before(kb, ib, ka, ia) : sel_before(ka, ia, kb, ib);
        if (swap) {
          keys[i] = kb;
          keys[q] = ka;
          ids[i] = ib;
          ids[q] = ia;
        }
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    double c = 0.0;
    if (s_special) {
      for (int i = 0; i < p; ++i) {
        c += alpha[ids[i]];
        keys[i] = static_cast<unsigned long long>(__double_as_longlong(c));
      }
    } else {
#pragma unroll 8
      for (int i = 0; i < p; ++i) {
        c += __longlong_as_double(static_cast<long long>(keys[i]));
        keys[i] = static_cast<unsigned long long>(__double_as_longlong(c));
      }
    }
  }
  __syncthreads();
  int below = 0;
  for (int i = threadIdx.x; i < p; i += blockDim.x) {
    const double c = __longlong_as_double(static_cast<long long>(keys[i]));
    if (D.csum) D.csum[static_cast<size_t>(l) * p + i] = c;
    if (D.order) D.order[static_cast<size_t>(l) * p + i] = ids[i];
    below += (c < threshold) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
  if ((threadIdx.x & 31) == 0 && below) atomicAdd(&s_count, below);
  __syncthreads();
  if (threadIdx.x == 0 && D.sizes) D.sizes[l] = (s_count == p) ? s_count : s_count + 1;
}

}  // namespace sb
