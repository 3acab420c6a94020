// capi.cu -- host side of the C ABI declared in include/sushie_b200.h.
//
// Owns device memory (one arena per batch), uploads problems, runs the device-side
// prologue (standardisation / XtX / Xty, reference infer.py:326-333,494 and
// infer_ss.py:340-344), launches the persistent IBSS kernel (ibss_kernel.cuh) with a
// thread-block-cluster or cooperative launch, and downloads results.  No CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sushie_b200.h"
#include "ibss_kernel.cuh"

using sb::KMAX;
using sb::LaunchParams;
using sb::OutBuf;
using sb::ProbDev;

static thread_local std::string g_init_error;

struct sb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int num_sms = 0;
  int smem_optin = 0;
  int cluster_ok = 0;
  int coop_ok = 0;
  int pool_ok = 0;  // stream-ordered allocator available (cudaMallocAsync)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_block = nullptr;  // blocking-sync event: host waits sleep instead of spinning on a core
  std::string err;
  // pinned staging for result downloads: device -> pinned at full PCIe rate, then host memcpy into the caller's
  // (pageable) buffers; pageable destinations alone ran at ~1 GB/s for the ~8 small fields of each fit
  unsigned char* pinned = nullptr;
  size_t pinned_bytes = 0;
  // pinned buffer for the small control transfers (iteration counts, index lists, purity results).  A copy to or
  // from PAGEABLE memory blocks inside the call until the stream reaches it, and the driver runs such copies one at
  // a time per process: one worker thread waiting for its device -> pageable copy behind a seconds-long IBSS kernel
  // stalled every pageable copy of the other workers (their uploads and downloads) for as long.
  unsigned char* ctl = nullptr;
  size_t ctl_bytes = 0;
};

// `bytes` of the context's pinned control buffer (grown on demand; contents are not preserved)
static unsigned char* ctl_buffer(sb_ctx* ctx, size_t bytes) {
  if (ctx->ctl_bytes >= bytes) return ctx->ctl;
  if (ctx->ctl) cudaFreeHost(ctx->ctl);
  ctx->ctl = nullptr;
  ctx->ctl_bytes = 0;
  const size_t want = std::max<size_t>(bytes + bytes / 2, size_t(4) << 20);
  if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->ctl), want, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  ctx->ctl_bytes = want;
  return ctx->ctl;
}

struct HostProb {
  int K = 0, p = 0, L = 0, p_pad = 0, nrows = 0;
  int n[KMAX] = {};
  size_t out_off[2] = {0, 0};  // offsets of the two output blocks inside the arena
  size_t elbo_off = 0, meta_off = 0;
  double bytes_per_ser = 0.0;  // algorithmic: one traversal of X / LD at storage precision
  std::vector<double> ecov0, rv0;
  void* d_X[KMAX] = {};
  const double* d_cmean[KMAX] = {};  // hard-call storage: column means, 1/std
  const double* d_cisd[KMAX] = {};
};

struct sb_batch {
  bool ss = false;
  int storage = SB_F64;
  int n_probs = 0;
  sb_opts opts{};
  int team_size = 1;
  int sms = 0;  // SMs available to the IBSS kernel (device SMs minus sb_opts.reserve_sms)
  int hw_cluster = 0;
  int n_teams_cap = 0;
  int smem_bytes = 0;
  int G_max = 0, NS_max = 0;
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  ProbDev* d_probs = nullptr;
  int* d_ids = nullptr;     // problem ids grouped by K
  int* d_queue = nullptr;   // device work-queue head of the running launch
  int* d_slots = nullptr;
  unsigned* d_bar = nullptr;
  size_t metas_off = 0;
  int* d_ctl = nullptr;
  double* d_scratch = nullptr;
  long long scratch_stride = 0;
  int off_stat = 0, off_trav = 0;
  int phases = 0;
  std::vector<int> ids_host;  // problem ids ordered by launch group
  struct Group {
    int K, variant, start, count;
  };
  std::vector<Group> groups;  // one kernel (K, traversal variant) per group
  std::vector<HostProb> hp;
  std::vector<ProbDev> pd;
  double prologue_ms = 0.0;
  int n_launches_create = 0;
  bool ran = false;
};

#define SB_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
      return (e_ == cudaErrorMemoryAllocation) ? SB_ENOMEM : SB_ECUDA;                      \
    }                                                                                       \
  } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------
// kernel table

// The kernel instantiations live in ibss_kernels.cu, one translation unit per ancestry count (built in parallel).
extern "C" const void* sb_kernel_table_k1(int storage, int ss, int variant);
extern "C" const void* sb_kernel_table_k2(int storage, int ss, int variant);
extern "C" const void* sb_kernel_table_k3(int storage, int ss, int variant);
extern "C" const void* sb_kernel_table_k4(int storage, int ss, int variant);
extern "C" const void* sb_kernel_table_k5(int storage, int ss, int variant);
extern "C" const void* sb_kernel_table_k6(int storage, int ss, int variant);
static_assert(KMAX == 6 && SB_MAX_ANCESTRIES == KMAX, "one kernel translation unit per ancestry count");

// variant: traversal variant compiled into the kernel -- hard-call storage: Geometry::fast (1..4); dense fp64
// individual level: 3 for the wide 12 x 1 tile; everything else 0 (variant chosen per locus at run time)
static int launch_variant(int storage, bool ss, int fast) {
  if (storage == SB_I8) return fast;
  if (storage == SB_F64 && !ss && (fast == 1 || fast == 3 || fast == 5 || fast == 6)) return fast;
  return 0;
}
static const void* pick_kernel(int K, int storage, bool ss, int variant) {
#ifdef SB_ONLY_K  // development builds: a single ancestry count is compiled and linked
  if (K != SB_ONLY_K) return nullptr;
#define SB_TABLE2(k) sb_kernel_table_k##k
#define SB_TABLE(k) SB_TABLE2(k)
  return SB_TABLE(SB_ONLY_K)(storage, ss ? 1 : 0, variant);
#else
  switch (K) {
    case 1: return sb_kernel_table_k1(storage, ss ? 1 : 0, variant);
    case 2: return sb_kernel_table_k2(storage, ss ? 1 : 0, variant);
    case 3: return sb_kernel_table_k3(storage, ss ? 1 : 0, variant);
    case 4: return sb_kernel_table_k4(storage, ss ? 1 : 0, variant);
    case 5: return sb_kernel_table_k5(storage, ss ? 1 : 0, variant);
    case 6: return sb_kernel_table_k6(storage, ss ? 1 : 0, variant);
  }
  return nullptr;
#endif
}

// ------------------------------------------------------------------------------------------
// context

extern "C" int sb_version(void) { return SB_VERSION; }

extern "C" int sb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" const char* sb_last_error(const sb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

extern "C" int sb_init(int device, sb_ctx** out) {
  if (!out) return SB_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_init_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    return SB_ENODEV;
  }
  if (device < 0 || device >= n) {
    g_init_error = "device index out of range";
    return SB_EINVAL;
  }
  sb_ctx* ctx = new sb_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_init_error = std::string("cudaSetDevice/GetDeviceProperties: ") + cudaGetErrorString(e);
    delete ctx;
    return SB_ECUDA;
  }
  if (prop.major < 9) {
    g_init_error = "sushie_b200 needs bulk-async-copy hardware (built for sm_100a); found sm_" +
                   std::to_string(prop.major) + std::to_string(prop.minor);
    delete ctx;
    return SB_ENODEV;
  }
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrClusterLaunch, device);
  ctx->cluster_ok = v;
  cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, device);
  ctx->coop_ok = v;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_init_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
    delete ctx;
    return SB_ECUDA;
  }
  for (auto& ev : ctx->ev) cudaEventCreate(&ev);
  cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
  cudaDeviceGetAttribute(&v, cudaDevAttrMemoryPoolsSupported, device);
  if (v && std::getenv("SUSHIE_B200_NO_POOL") == nullptr) {
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;  // never hand memory back between chunks
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      ctx->pool_ok = 1;
    }
    cudaGetLastError();
  }
  *out = ctx;
  return SB_OK;
}

extern "C" int sb_destroy(sb_ctx* ctx) {
  if (!ctx) return SB_OK;
  cudaSetDevice(ctx->device);
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->ctl) cudaFreeHost(ctx->ctl);
  if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SB_OK;
}

extern "C" int sb_set_stream(sb_ctx* ctx, void* cuda_stream) {
  if (!ctx) return SB_EINVAL;
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = false;
  return SB_OK;
}

// Wait for the context's stream WITHOUT spinning: several worker threads (and, under torchrun, several processes)
// wait on their GPUs at the same time, and a spinning cudaStreamSynchronize per waiter takes the host cores away
// from the threads doing the host half of other chunks (8 ranks on a 32-core box: config-5 throughput halved).
static cudaError_t sync_stream(sb_ctx* ctx) {
  if (!ctx->ev_block) return cudaStreamSynchronize(ctx->stream);
  cudaError_t e = cudaEventRecord(ctx->ev_block, ctx->stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ctx->ev_block);
}

// Small device -> host read through the pinned control buffer at `ctl_off`: enqueue, wait for the stream's event,
// copy out.  (Never a device -> pageable copy: see sb_ctx::ctl.)
static cudaError_t read_small(sb_ctx* ctx, void* dst, const void* d_src, size_t bytes, size_t ctl_off = 0) {
  unsigned char* buf = ctl_buffer(ctx, ctl_off + bytes);
  if (!buf) {
    cudaError_t e = cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    return e != cudaSuccess ? e : sync_stream(ctx);
  }
  cudaError_t e = cudaMemcpyAsync(buf + ctl_off, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (e != cudaSuccess) return e;
  if ((e = sync_stream(ctx)) != cudaSuccess) return e;
  std::memcpy(dst, buf + ctl_off, bytes);
  return cudaSuccess;
}

// Device memory comes from the device's stream-ordered pool: cudaFree (and often cudaMalloc) synchronise the WHOLE
// device, so with two worker threads per GPU the one releasing its finished batch used to wait for the other's IBSS
// kernel.  Allocation and release are ordered on the context's stream instead; the pool keeps what it has been
// given (release threshold raised in sb_init), so steady-state chunks allocate without a driver call.
static cudaError_t dev_alloc(sb_ctx* ctx, void** ptr, size_t bytes) {
  return ctx->pool_ok ? cudaMallocAsync(ptr, bytes, ctx->stream) : cudaMalloc(ptr, bytes);
}
static void dev_free(sb_ctx* ctx, void* ptr) {
  if (!ptr) return;
  if (ctx && ctx->pool_ok) cudaFreeAsync(ptr, ctx->stream);
  else cudaFree(ptr);
}

// ------------------------------------------------------------------------------------------
// geometry: rows per band / stages / sweep-1 segmentation for one problem

struct Geometry {
  int p_pad, G, NS, NSEG, fast;
  size_t smem;
};

static bool plan_geometry(int p, int n_max, int es, bool ss, int smem_budget, Geometry& g) {
  g.p_pad = static_cast<int>(align_up(static_cast<size_t>(p), 32));
  const size_t hdr = align_up(sizeof(sb::SmemHdr), 128);
  const size_t row_bytes = static_cast<size_t>(g.p_pad) * es;
  const int pv = static_cast<int>(row_bytes / 16);  // 16-byte vectors per row
  g.fast = 0;
  if (es == 1 && g.p_pad <= 512) {
    // hard-call storage, narrow locus: row-team traversal (traverse_i8_rows<16, 1>): one warp owns one row at a
    // time, every lane 16 genotype bytes of it; the 8 warps of a block work on different rows of the same band
    g.fast = 5;
    const int n_teams = sb::NW;
    const size_t state = 2 * sizeof(double) * sb::I8_STATE_ROWS;       // staged (r, xb) rows behind the ring
    const size_t xteam = sizeof(double) * n_teams * row_bytes;         // cross-team exchange of the partial G^T r
    const size_t ring_budget = static_cast<size_t>(smem_budget) - hdr - state - xteam;
    int rows = static_cast<int>(std::min<size_t>(sb::GMAX, (32 * 1024) / row_bytes));
    rows = std::max(n_teams, rows / n_teams * n_teams);
    g.G = rows;
    g.NS = static_cast<int>(std::min<size_t>(sb::NSTAGE_MAX, ring_budget / (rows * row_bytes)));
    g.NSEG = 1;
    g.smem = hdr + static_cast<size_t>(g.NS) * g.G * row_bytes + state + xteam;
    return true;
  }
  if (es == 1) {
    // hard-call storage: every thread owns CPT 8-byte vectors of GR rows (traverse_i8)
    const int pv8 = g.p_pad / 8;
    int gr;
    if (g.p_pad / 4 <= sb::NT) {  // narrow locus: 4-byte vectors keep every thread busy
      g.fast = 4;
      gr = 16;
    } else if (pv8 <= sb::NT) {
      g.fast = 1;
      gr = 8;
    } else if (pv8 <= 2 * sb::NT) {
      g.fast = 2;
      gr = 4;
    } else if (pv8 <= 4 * sb::NT) {
      g.fast = 3;
      gr = 2;
    } else {
      return false;
    }
    const size_t state = 2 * sizeof(double) * sb::I8_STATE_ROWS;  // staged (r, xb) rows behind the ring
    const int budget = smem_budget / SB_I8_BLOCKS_PER_SM - (SB_I8_BLOCKS_PER_SM > 1 ? 1024 : 0);  // co-resident blocks
    const int rows_fit = static_cast<int>((budget - hdr - state) / row_bytes);
    g.G = gr;
    g.NS = std::max(2, std::min(sb::NSTAGE_MAX, rows_fit / gr));
    g.NSEG = 1;
    g.smem = hdr + static_cast<size_t>(g.NS) * g.G * row_bytes + state;
    return true;
  }
  // register-tile traversal (individual level): every thread owns CPT vectors of GR rows
  {
    int gr = 0;
    if (!ss && es == 8 && pv >= sb::NT / 2 && pv <= sb::NT) {  // narrow fp64 loci: 1 or 2 vectors x 8 rows (own kernels)
      g.fast = 5;
      gr = 8;
    } else if (!ss && es == 8 && pv > sb::NT && pv <= 2 * sb::NT) {
      g.fast = 6;
      gr = 8;
    } else if (pv > sb::NT && pv <= 4 * sb::NT) {
      g.fast = 1;
      gr = 4;
    } else if (pv > 4 * sb::NT && pv <= 8 * sb::NT) {
      g.fast = 2;
      gr = 2;
    } else if ((ss || es == 8) && pv > 8 * sb::NT && pv <= 12 * sb::NT) {  // (fp32 individual level: generic path)
      g.fast = 3;
      gr = 1;
    } else if (ss && pv > 12 * sb::NT && pv <= 16 * sb::NT) {
      g.fast = 4;
      gr = 1;
    }
    if (g.fast) {
      // summary level: partial row sums of a chunk are parked behind the ring
      const size_t park = ss ? sizeof(double) * sb::SS_CHUNK_ROWS * sb::NW : 0;
      const int rows_fit = static_cast<int>((smem_budget - hdr - park) / row_bytes);
      if (rows_fit >= 2 * gr) {
        g.G = gr;
        g.NS = std::min(sb::NSTAGE_MAX, rows_fit / gr);
        g.NSEG = 1;
        g.smem = hdr + static_cast<size_t>(g.NS) * g.G * row_bytes + park;
        return true;
      }
      g.fast = 0;
    }
  }
  size_t fixed = hdr + static_cast<size_t>(g.p_pad) * sizeof(double) * (ss ? 1 : 2);
  if (fixed + 2 * row_bytes > static_cast<size_t>(smem_budget)) {
    // very wide locus: keep b / the accumulator in global memory, all shared memory goes to the ring
    fixed = hdr;
    if (fixed + 2 * row_bytes > static_cast<size_t>(smem_budget)) return false;
    g.fast = sb::FAST_GLOBAL_VECTORS;
  }
  const int rows_fit = static_cast<int>((smem_budget - fixed) / row_bytes);
  g.NS = rows_fit >= 3 ? 3 : 2;
  g.G = std::max(1, std::min({sb::GMAX, rows_fit / g.NS, std::max(1, n_max)}));
  // spend left-over shared memory on a 4th stage when bands are small
  if (g.NS == 3 && rows_fit >= 4 * g.G && g.G * row_bytes < 48 * 1024) g.NS = 4;
  int nseg = 1;
  while (nseg * 2 <= sb::NW / g.G) nseg *= 2;
  g.NSEG = nseg;
  g.smem = fixed + static_cast<size_t>(g.NS) * g.G * row_bytes;
  return true;
}

static int choose_team(const sb_ctx* ctx, int S, int n_probs, double bytes_per_locus, int requested, int& hw_cluster) {
  int c;
  if (requested > 0) {
    c = std::min(requested, S);
  } else {
    const int avail = std::max(1, S / std::max(1, n_probs));
    const int want = std::max(1, static_cast<int>(bytes_per_locus / (256.0 * 1024.0)));
    c = std::min(want, avail);
    if (n_probs >= S) c = 1;  // big batches: one block per locus, the phased schedule handles the tail
  }
  if (c <= 16) {
    int p2 = 1;
    while (p2 * 2 <= c) p2 *= 2;
    c = p2;
    hw_cluster = (c > 1 && ctx->cluster_ok) ? 1 : 0;
    if (c > 1 && !ctx->cluster_ok) hw_cluster = 0;
  } else {
    hw_cluster = 0;
  }
  if (c > 1 && !hw_cluster && !ctx->coop_ok) c = 1;
  return c;
}

// ------------------------------------------------------------------------------------------
// arena bookkeeping

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes, size_t align = 256) {
    off = align_up(off, align);
    const size_t o = off;
    off += bytes;
    return o;
  }
};

struct ProbLayout {  // offsets inside the arena
  size_t rbits[KMAX], cbits[KMAX], Qb[KMAX], Mq[KMAX];
  size_t X[KMAX], y, logpi, pi_raw, xtx, cmean, cisd, ecov0, rv0, adjt, adjp, r, xb, xty, rfull, bvec, out[2], elbo, meta, z;
  size_t out_fields[2][10];
};

static size_t out_block_bytes(int L, int p, int K) {
  const size_t Lp = static_cast<size_t>(L) * p;
  return sizeof(double) * (Lp + Lp * K + Lp * K * K + Lp + static_cast<size_t>(L) * K * K + L +
                           static_cast<size_t>(L) * K * K + K + static_cast<size_t>(L) * K + 2 * static_cast<size_t>(L) + 8);
}

static void out_block_fields(int L, int p, int K, size_t base, size_t (&f)[10]) {
  const size_t Lp = static_cast<size_t>(L) * p;
  size_t o = base;
  f[0] = o; o += Lp * 8;                               // alpha
  f[1] = o; o += Lp * K * 8;                           // pm
  f[2] = o; o += Lp * K * K * 8;                       // pmsq
  f[3] = o; o += Lp * 8;                               // lbf
  f[4] = o; o += static_cast<size_t>(L) * K * K * 8;   // wsc
  f[5] = o; o += static_cast<size_t>(L) * 8;           // kl
  f[6] = o; o += static_cast<size_t>(L) * K * K * 8;   // ecov
  f[7] = o; o += static_cast<size_t>(K) * 8;           // rv
  f[8] = o; o += static_cast<size_t>(L) * K * 8;       // xtxb2
  f[9] = o;                                            // lse
}

template <typename P>
static int validate_common(sb_ctx* ctx, const P& pr, int i) {
  if (pr.k < 1 || pr.k > KMAX) {
    ctx->err = "problem " + std::to_string(i) + ": k=" + std::to_string(pr.k) + " unsupported (1.." +
               std::to_string(KMAX) + ")";
    return SB_EINVAL;
  }
  if (pr.p < 1 || pr.L < 1) {
    ctx->err = "problem " + std::to_string(i) + ": p and L must be positive";
    return SB_EINVAL;
  }
  if (!pr.resid_var || !pr.effect_covar) {
    ctx->err = "problem " + std::to_string(i) + ": resid_var / effect_covar missing";
    return SB_EINVAL;
  }
  if (!pr.em_update && (!pr.adj_times || !pr.adj_plus)) {
    ctx->err = "problem " + std::to_string(i) + ": adj_times / adj_plus required when em_update == 0";
    return SB_EINVAL;
  }
  return SB_OK;
}

static int elem_size(int dtype) { return dtype == SB_F64 ? 8 : (dtype == SB_F32 ? 4 : 1); }
// hard-call storages keep the raw one-byte matrix resident (SB_B2 adds the bit planes next to it)
static bool hard_storage(int storage) { return storage == SB_I8 || storage == SB_B2; }

template <typename IT, typename XT>
static void launch_prep_columns(const void* in, int in_ld, int n, int p, int p_pad, int flags, void* out,
                                double* xtx, cudaStream_t st) {
  const int T = 128;
  sb::prep_columns_kernel<IT, XT><<<(p_pad + T - 1) / T, T, 0, st>>>(
      static_cast<const IT*>(in), in_ld, n, p, p_pad, flags, static_cast<XT*>(out), xtx);
}

static void prep_columns_dispatch(int in_dtype, int storage, const void* in, int in_ld, int n, int p, int p_pad,
                                  int flags, void* out, double* xtx, cudaStream_t st) {
  if (storage == SB_F64) {
    if (in_dtype == SB_F64) launch_prep_columns<double, double>(in, in_ld, n, p, p_pad, flags, out, xtx, st);
    if (in_dtype == SB_F32) launch_prep_columns<float, double>(in, in_ld, n, p, p_pad, flags, out, xtx, st);
    if (in_dtype == SB_I8) launch_prep_columns<int8_t, double>(in, in_ld, n, p, p_pad, flags, out, xtx, st);
  } else {
    if (in_dtype == SB_F64) launch_prep_columns<double, float>(in, in_ld, n, p, p_pad, flags, out, xtx, st);
    if (in_dtype == SB_F32) launch_prep_columns<float, float>(in, in_ld, n, p, p_pad, flags, out, xtx, st);
    if (in_dtype == SB_I8) launch_prep_columns<int8_t, float>(in, in_ld, n, p, p_pad, flags, out, xtx, st);
  }
}

template <typename IT, typename XT>
static void launch_pad_rows(const void* in, int n, int p, int p_pad, void* out, cudaStream_t st) {
  const size_t total = static_cast<size_t>(n) * p_pad;
  const int T = 256;
  const int blocks = static_cast<int>(std::min<size_t>((total + T - 1) / T, 148 * 16));
  sb::pad_rows_kernel<IT, XT><<<blocks, T, 0, st>>>(static_cast<const IT*>(in), n, p, p_pad, static_cast<XT*>(out));
}

// ------------------------------------------------------------------------------------------
// batch creation (shared by individual / summary level)

struct GenericProblem {
  int K, p, L, em;
  int n[KMAX];
  const void* mat[KMAX];
  int mat_dtype;
  int standardize;
  const double* vec[KMAX];  // y (individual) or z (summary)
  double ns[KMAX];
  const double *pi, *resid_var, *effect_covar, *adj_times, *adj_plus;
  const double* Q[KMAX] = {};  // covariate basis per ancestry (SB_B2 only)
  int qc[KMAX] = {};
  const int32_t* rows[KMAX] = {};  // row view: this problem's rows inside the source matrix mat[k] (nullptr: all rows)
  int src_rows[KMAX] = {};         // rows of the source matrix when rows[k] is given
  // bytes of staging one (problem, ancestry) needs: the source matrix plus, for a row view, its row list
  size_t stage_need(int k) const {
    const size_t es_in = mat_dtype == SB_F64 ? 8 : (mat_dtype == SB_F32 ? 4 : 1);
    const size_t src = static_cast<size_t>(rows[k] ? src_rows[k] : n[k]) * p * es_in;
    return (src + 255) / 256 * 256 + (rows[k] ? (sizeof(int32_t) * n[k] + 255) / 256 * 256 : 0);
  }
};

static int batch_create(sb_ctx* ctx, const std::vector<GenericProblem>& gp, bool ss, const sb_opts* opts_in,
                        sb_batch** out) {
  if (!ctx || !out) return SB_EINVAL;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  const int n_probs = static_cast<int>(gp.size());
  if (n_probs < 1) {
    ctx->err = "empty batch";
    return SB_EINVAL;
  }
  sb_opts opts{};
  opts.max_iter = 500;
  opts.min_tol = 1e-4;
  opts.storage = SB_F64;
  if (opts_in) opts = *opts_in;
  if (opts.max_iter < 1) {
    ctx->err = "max_iter must be >= 1";
    return SB_EINVAL;
  }
  if (opts.storage != SB_F64 && opts.storage != SB_F32 && opts.storage != SB_I8 && opts.storage != SB_B2) {
    ctx->err = "storage must be SB_F64, SB_F32, SB_I8 or SB_B2";
    return SB_EINVAL;
  }
  if (hard_storage(opts.storage)) {
    if (ss) {
      ctx->err = "SB_I8 / SB_B2 storage is for individual-level hard-call genotypes only";
      return SB_EINVAL;
    }
    for (int i = 0; i < n_probs; ++i)
      if (gp[i].mat_dtype != SB_I8) {
        ctx->err = "problem " + std::to_string(i) + ": SB_I8 / SB_B2 storage needs x_dtype == SB_I8";
        return SB_EINVAL;
      }
  }
  const int es = elem_size(opts.storage);
  const int smem_budget = ctx->smem_optin - 1024;  // leave room for static shared memory

  std::unique_ptr<sb_batch> b(new sb_batch());
  b->ss = ss;
  b->storage = opts.storage;
  b->n_probs = n_probs;
  b->opts = opts;
  b->hp.resize(n_probs);
  b->pd.resize(n_probs);

  // geometry + team size
  std::vector<Geometry> geo(n_probs);
  double bytes_mean = 0.0;
  size_t smem_max = 0;
  for (int i = 0; i < n_probs; ++i) {
    const GenericProblem& g = gp[i];
    int n_max = 0;
    double bytes = 0.0;
    for (int k = 0; k < g.K; ++k) {
      if (g.n[k] < 1) {
        ctx->err = "problem " + std::to_string(i) + ": empty ancestry " + std::to_string(k);
        return SB_EINVAL;
      }
      n_max = std::max(n_max, g.n[k]);
      bytes += static_cast<double>(g.n[k]) * g.p * es;
    }
    if (opts.storage == SB_B2) {
      Geometry& ge = geo[i];
      if (g.p > 8 * sb::B2_MAX_SUB * sb::B2_TILE_COLS || n_max > sb::B2_MAX_ROWS) {
        ctx->err = "problem " + std::to_string(i) + ": SB_B2 storage handles p <= " +
                   std::to_string(8 * sb::B2_MAX_SUB * sb::B2_TILE_COLS) + " and n <= " + std::to_string(sb::B2_MAX_ROWS) +
                   " per ancestry";
        return SB_EINVAL;
      }
      ge.p_pad = static_cast<int>(align_up(static_cast<size_t>(g.p), 32));
      ge.G = 8;  // (no row bands: the bit stream has its own ring)
      ge.NS = sb::B2_NSTAGE;
      ge.NSEG = 1;
      ge.fast = 0;
      ge.smem = static_cast<size_t>(smem_budget);  // table set aligned to 64 KB inside the block's window
      bytes *= 0.5;                                // two layouts x two bit planes per genotype
    } else if (!plan_geometry(g.p, n_max, es, ss, smem_budget, geo[i])) {
      ctx->err = "problem " + std::to_string(i) + ": p=" + std::to_string(g.p) +
                 " does not fit the shared-memory band pipeline at this storage precision";
      return SB_EINVAL;
    }
    b->hp[i].bytes_per_ser = bytes;
    bytes_mean += bytes / n_probs;
    smem_max = std::max(smem_max, geo[i].smem);
    b->G_max = std::max(b->G_max, geo[i].G);
    b->NS_max = std::max(b->NS_max, geo[i].NS);
  }
  b->smem_bytes = static_cast<int>(smem_max);
  // SMs the IBSS launches of this batch may occupy (sb_opts.reserve_sms are left to other streams' small kernels)
  b->sms = std::max(1, ctx->num_sms - std::max(0, std::min(opts.reserve_sms, ctx->num_sms / 2)));
  b->team_size = choose_team(ctx, b->sms, n_probs, bytes_mean, opts.team_size, b->hw_cluster);
  const int C = b->team_size;

  // arena layout
  Bump bump, small;
  std::vector<ProbLayout> lay(n_probs);
  const size_t probs_off = bump.take(sizeof(ProbDev) * n_probs);
  const size_t ids_off = bump.take(sizeof(int) * n_probs);
  const size_t queue_off = bump.take(sizeof(int) * 64);
  const size_t slots_off = bump.take(sizeof(int) * 2 * (2 * ctx->num_sms + 8));
  const size_t ctl_off = bump.take(sizeof(int) * 64);
  const size_t bar_off = bump.take(sizeof(unsigned) * 64 * (2 * ctx->num_sms + 8));
  const size_t metas_off = bump.take(sizeof(int) * 8 * n_probs);
  const size_t elbos_off = bump.take(sizeof(double) * (opts.max_iter + 1) * static_cast<size_t>(n_probs));
  size_t staging_bytes = 0;
  size_t Kp_max = 0;
  int L_max = 1;
  for (int i = 0; i < n_probs; ++i) {
    const GenericProblem& g = gp[i];
    const Geometry& ge = geo[i];
    ProbLayout& l = lay[i];
    HostProb& h = b->hp[i];
    h.K = g.K;
    h.p = g.p;
    h.L = g.L;
    h.p_pad = ge.p_pad;
    int nrows = 0;
    for (int k = 0; k < g.K; ++k) {
      h.n[k] = g.n[k];
      nrows += g.n[k];
      l.X[k] = bump.take(static_cast<size_t>(g.n[k]) * ge.p_pad * es);
      const bool direct = (g.mat_dtype == opts.storage) && g.standardize == 0 && !hard_storage(opts.storage) && !g.rows[k];
      l.Qb[k] = l.Mq[k] = 0;
      if (g.qc[k] > 0) {
        l.Qb[k] = bump.take(sizeof(double) * g.n[k] * g.qc[k]);  // (uploaded before the column prologue that reads it)
        l.Mq[k] = bump.take(sizeof(double) * g.qc[k] * ge.p_pad);
      }
      if (opts.storage == SB_B2) {
        const size_t ntiles = (static_cast<size_t>(g.p) + sb::B2_TILE_COLS - 1) / sb::B2_TILE_COLS;
        l.rbits[k] = bump.take(ntiles * ((g.n[k] + 3) / 4) * 256);
        l.cbits[k] = bump.take(((g.n[k] + 255) / 256) * (64 * ntiles) * 256);
      }
      if (!direct) staging_bytes = std::max(staging_bytes, g.stage_need(k));
    }
    h.nrows = nrows;
    const size_t Kp = static_cast<size_t>(g.K) * ge.p_pad;
    // host-filled small vectors live in one contiguous region (offsets relative to its base for now): the whole
    // batch's priors, phenotypes / z-scores and prior weights go up in ONE copy instead of ~10 per locus
    (ss ? l.z : l.y) = small.take(sizeof(double) * nrows, 16);
    if (ss) l.y = bump.take(sizeof(double) * nrows);  // Xty, formed on device from z
    l.pi_raw = g.pi ? small.take(sizeof(double) * g.p, 16) : 0;
    l.ecov0 = small.take(sizeof(double) * g.L * g.K * g.K, 16);
    l.rv0 = small.take(sizeof(double) * g.K, 16);
    l.adjt = small.take(sizeof(double) * g.K * g.K, 16);
    l.adjp = small.take(sizeof(double) * g.K * g.K, 16);
    l.logpi = bump.take(sizeof(double) * g.p);
    l.xtx = bump.take(sizeof(double) * Kp);
    l.cmean = hard_storage(opts.storage) ? bump.take(sizeof(double) * Kp) : 0;
    l.cisd = hard_storage(opts.storage) ? bump.take(sizeof(double) * Kp) : 0;
    l.r = bump.take(sizeof(double) * nrows);
    l.xb = bump.take(sizeof(double) * nrows * g.L);
    l.xty = ss ? 0 : bump.take(sizeof(double) * Kp);
    l.rfull = ss ? bump.take(sizeof(double) * nrows) : 0;
    l.bvec = bump.take(sizeof(double) * g.L * Kp);
    Kp_max = std::max(Kp_max, Kp);
    L_max = std::max(L_max, g.L);
    for (int q = 0; q < 2; ++q) {
      l.out[q] = bump.take(out_block_bytes(g.L, g.p, g.K));
      out_block_fields(g.L, g.p, g.K, l.out[q], l.out_fields[q]);
      h.out_off[q] = l.out[q];
    }
    l.elbo = elbos_off + sizeof(double) * (opts.max_iter + 1) * static_cast<size_t>(i);
    l.meta = metas_off + sizeof(int) * 8 * static_cast<size_t>(i);
    h.elbo_off = l.elbo;
    h.meta_off = l.meta;
  }
  const size_t small_bytes = align_up(small.off, 256);
  const size_t small_base = bump.take(small_bytes);
  for (int i = 0; i < n_probs; ++i) {
    ProbLayout& l = lay[i];
    (ss ? l.z : l.y) += small_base;
    if (gp[i].pi) l.pi_raw += small_base;
    l.ecov0 += small_base;
    l.rv0 += small_base;
    l.adjt += small_base;
    l.adjp += small_base;
  }
  // staging for the batched column prologue: as many input matrices as fit under the cap share one launch
  size_t staging_total = 0;
  for (int i = 0; i < n_probs; ++i)
    for (int k = 0; k < gp[i].K; ++k)
      staging_total += gp[i].stage_need(k);
  const size_t staging_cap = std::max(staging_bytes + 512, std::min<size_t>(staging_total + 512, size_t(1) << 30));
  const size_t staging_off = bump.take(staging_bytes ? staging_cap : 16);
  const size_t descs_off = bump.take(sizeof(sb::PrepDesc) * static_cast<size_t>(n_probs) * KMAX);
  const size_t pack_off = bump.take(sizeof(sb::PackDesc) * static_cast<size_t>(n_probs) * KMAX);
  // per-block scratch: partial X^T r, three reduction records, per-traversal scalars
  const int off_stat = static_cast<int>(align_up(Kp_max, 16));
  const int off_trav = off_stat + 3 * sb::STAT_STRIDE + 4;
  const size_t scratch_stride = align_up(static_cast<size_t>(off_trav) + static_cast<size_t>(L_max) * sb::TRAV_STRIDE, 32);
  const size_t scratch_off = bump.take(sizeof(double) * scratch_stride * (2 * ctx->num_sms + 8));
  b->arena_bytes = align_up(bump.off, 256);
  {
    cudaError_t e = dev_alloc(ctx, reinterpret_cast<void**>(&b->arena), b->arena_bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      ctx->err = "device allocation (" + std::to_string(b->arena_bytes) + " bytes): " + cudaGetErrorString(e);
      b->arena = nullptr;
      return SB_ENOMEM;
    }
  }
  unsigned char* A = b->arena;
  b->d_probs = reinterpret_cast<ProbDev*>(A + probs_off);
  b->d_ids = reinterpret_cast<int*>(A + ids_off);
  b->d_queue = reinterpret_cast<int*>(A + queue_off);
  b->d_slots = reinterpret_cast<int*>(A + slots_off);
  b->d_bar = reinterpret_cast<unsigned*>(A + bar_off);
  b->metas_off = metas_off;
  b->d_ctl = reinterpret_cast<int*>(A + ctl_off);
  b->d_scratch = reinterpret_cast<double*>(A + scratch_off);
  b->scratch_stride = static_cast<long long>(scratch_stride);
  b->off_stat = off_stat;
  b->off_trav = off_trav;
  cudaStream_t st = ctx->stream;
  auto fail = [&](int code) {
    sync_stream(ctx);
    dev_free(ctx, b->arena);
    b->arena = nullptr;
    return code;
  };
#define SB_CUDA_B(call)                                                          \
  do {                                                                           \
    cudaError_t e_ = (call);                                                     \
    if (e_ != cudaSuccess) {                                                     \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);             \
      return fail(SB_ECUDA);                                                     \
    }                                                                            \
  } while (0)

  SB_CUDA_B(cudaMemsetAsync(b->d_ctl, 0, sizeof(int) * 64, st));
  SB_CUDA_B(cudaEventRecord(ctx->ev[2], st));
  int launches = 0;
  std::vector<unsigned char> small_host(small_bytes, 0);
  struct SsPrep {
    const void* ld;
    const double* z;
    double ns;
    int p, p_pad;
    double *xty, *xtx;
  };
  std::vector<SsPrep> ss_prep;
  std::vector<sb::PackDesc> pack_descs;  // bit-plane storage: one packing job per (problem, ancestry)
  int p_max_all = 0;
  // batched column prologue: matrices are staged back to back, described, and standardised by one launch per
  // staging fill (and per input type)
  std::vector<sb::PrepDesc> pending;
  std::map<const void*, std::pair<size_t, size_t>> staged_src;  // host matrix -> (staging offset, bytes):
  // source matrices in the current staging fill (row views share them)
  size_t staged = 0, descs_used = 0;
  int pending_dtype = -1, pending_pmax = 0;
  auto flush_prep = [&]() -> cudaError_t {
    if (pending.empty()) return cudaSuccess;
    sb::PrepDesc* d_descs = reinterpret_cast<sb::PrepDesc*>(A + descs_off) + descs_used;
    cudaError_t e = cudaMemcpyAsync(d_descs, pending.data(), sizeof(sb::PrepDesc) * pending.size(),
                                    cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    const dim3 grid((pending_pmax + 127) / 128, static_cast<unsigned>(pending.size()));
    int* flag = b->d_ctl + 32;
#define SB_PREP(IT, XT) sb::prep_columns_batched_kernel<IT, XT><<<grid, 128, 0, st>>>(d_descs, flag)
    if (hard_storage(opts.storage)) {
      SB_PREP(int8_t, int8_t);
    } else if (opts.storage == SB_F64) {
      if (pending_dtype == SB_F64) SB_PREP(double, double);
      if (pending_dtype == SB_F32) SB_PREP(float, double);
      if (pending_dtype == SB_I8) SB_PREP(int8_t, double);
    } else {
      if (pending_dtype == SB_F64) SB_PREP(double, float);
      if (pending_dtype == SB_F32) SB_PREP(float, float);
      if (pending_dtype == SB_I8) SB_PREP(int8_t, float);
    }
#undef SB_PREP
    ++launches;
    descs_used += pending.size();
    pending.clear();
    staged_src.clear();
    staged = 0;
    pending_pmax = 0;
    return cudaGetLastError();
  };
  auto enqueue_prep = [&](const void* host_mat, int dtype, int n, int p, int p_pad, int flags, void* out, double* xtx,
                          double* cmean, double* cisd, const double* d_Q = nullptr, double* d_M = nullptr,
                          int qc = 0, const int32_t* rows_host = nullptr, int src_rows = 0) -> cudaError_t {
    // the source matrix: `n` rows, or -- for a row view -- `src_rows` rows of which the problem uses rows_host[0..n)
    const size_t bytes = static_cast<size_t>(rows_host ? src_rows : n) * p * elem_size(dtype);
    const size_t rows_bytes = rows_host ? align_up(sizeof(int32_t) * n, 256) : 0;
    if (!pending.empty() && dtype != pending_dtype) {
      cudaError_t e = flush_prep();
      if (e != cudaSuccess) return e;
    }
    auto hit = staged_src.find(host_mat);
    if (hit != staged_src.end() && hit->second.second != bytes) hit = staged_src.end();  // same pointer, other extent
    if (hit == staged_src.end() ? staged + align_up(bytes, 256) + rows_bytes > staging_cap
                                : staged + rows_bytes > staging_cap) {
      cudaError_t e = flush_prep();  // (also forgets the staged sources)
      if (e != cudaSuccess) return e;
      hit = staged_src.end();
    }
    cudaError_t e;
    size_t src_off;
    if (hit != staged_src.end()) {
      src_off = hit->second.first;  // a problem of this staging fill has uploaded the same matrix already
    } else {
      src_off = staged;
      e = cudaMemcpyAsync(A + staging_off + staged, host_mat, bytes, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return e;
      staged_src[host_mat] = std::make_pair(src_off, bytes);
      staged += align_up(bytes, 256);
    }
    const int* d_rows = nullptr;
    if (rows_host) {
      e = cudaMemcpyAsync(A + staging_off + staged, rows_host, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) return e;
      d_rows = reinterpret_cast<const int*>(A + staging_off + staged);
      staged += rows_bytes;
    }
    sb::PrepDesc pd{};
    pd.in = A + staging_off + src_off;
    pd.rows = d_rows;
    pd.out = out;
    pd.xtx = xtx;
    pd.cmean = cmean;
    pd.cisd = cisd;
    pd.in_ld = p;
    pd.n = n;
    pd.p = p;
    pd.p_pad = p_pad;
    pd.flags = flags;
    pd.qc = qc;
    pd.Q = d_Q;
    pd.M = d_M;
    pending.push_back(pd);
    pending_dtype = dtype;
    pending_pmax = std::max(pending_pmax, p_pad);
    return cudaSuccess;
  };
  for (int i = 0; i < n_probs; ++i) {
    const GenericProblem& g = gp[i];
    const Geometry& ge = geo[i];
    const ProbLayout& l = lay[i];
    HostProb& h = b->hp[i];
    ProbDev& d = b->pd[i];
    std::memset(&d, 0, sizeof(d));
    d.K = g.K;
    d.p = g.p;
    d.L = g.L;
    d.p_pad = ge.p_pad;
    d.G = ge.G;
    d.nstage = ge.NS;
    d.nseg = ge.NSEG;
    d.fast = ge.fast;
    d.em = g.em;
    int row = 0, band = 0;
    for (int k = 0; k < g.K; ++k) {
      d.n[k] = g.n[k];
      d.row0[k] = row;
      d.band0[k] = band;
      row += g.n[k];
      band += (opts.storage == SB_B2) ? 0 : (g.n[k] + ge.G - 1) / ge.G;  // bit-plane storage: no row bands
      d.X[k] = A + l.X[k];
      if (opts.storage == SB_B2) {
        d.rbits[k] = A + l.rbits[k];
        d.cbits[k] = A + l.cbits[k];
        sb::PackDesc pk{};
        pk.X = reinterpret_cast<const signed char*>(A + l.X[k]);
        pk.rbits = A + l.rbits[k];
        pk.cbits = A + l.cbits[k];
        pk.n = g.n[k];
        pk.p_pad = ge.p_pad;
        pk.ntiles = (g.p + sb::B2_TILE_COLS - 1) / sb::B2_TILE_COLS;
        pack_descs.push_back(pk);
      }
      h.d_X[k] = A + l.X[k];
      d.nscale[k] = ss ? g.ns[k] : 1.0;
      d.nsamp[k] = ss ? g.ns[k] : static_cast<double>(g.n[k]);
    }
    for (int k = g.K; k <= KMAX; ++k) {
      d.row0[k] = row;
      d.band0[k] = band;
    }
    d.y = reinterpret_cast<double*>(A + l.y);
    d.logpi = reinterpret_cast<double*>(A + l.logpi);
    d.xtx = reinterpret_cast<double*>(A + l.xtx);
    d.cmean = hard_storage(opts.storage) ? reinterpret_cast<double*>(A + l.cmean) : nullptr;
    d.cisd = hard_storage(opts.storage) ? reinterpret_cast<double*>(A + l.cisd) : nullptr;
    d.ntiles = (g.p + sb::B2_TILE_COLS - 1) / sb::B2_TILE_COLS;
    d.ecov0 = reinterpret_cast<double*>(A + l.ecov0);
    d.rv0 = reinterpret_cast<double*>(A + l.rv0);
    d.adj_times = reinterpret_cast<double*>(A + l.adjt);
    d.adj_plus = reinterpret_cast<double*>(A + l.adjp);
    d.r = reinterpret_cast<double*>(A + l.r);
    d.xb = reinterpret_cast<double*>(A + l.xb);
    d.xty = ss ? nullptr : reinterpret_cast<double*>(A + l.xty);
    d.rfull = ss ? reinterpret_cast<double*>(A + l.rfull) : nullptr;
    d.bvec = reinterpret_cast<double*>(A + l.bvec);
    SB_CUDA_B(cudaMemsetAsync(A + l.bvec, 0, sizeof(double) * g.L * g.K * ge.p_pad, st));  // pad columns stay 0
    for (int q = 0; q < 2; ++q) {
      OutBuf& o = d.out[q];
      const size_t* f = l.out_fields[q];
      o.alpha = reinterpret_cast<double*>(A + f[0]);
      o.pm = reinterpret_cast<double*>(A + f[1]);
      o.pmsq = reinterpret_cast<double*>(A + f[2]);
      o.lbf = reinterpret_cast<double*>(A + f[3]);
      o.wsc = reinterpret_cast<double*>(A + f[4]);
      o.kl = reinterpret_cast<double*>(A + f[5]);
      o.ecov = reinterpret_cast<double*>(A + f[6]);
      o.rv = reinterpret_cast<double*>(A + f[7]);
      o.xtxb2 = reinterpret_cast<double*>(A + f[8]);
      o.lse = reinterpret_cast<double*>(A + f[9]);
    }
    d.elbo = reinterpret_cast<double*>(A + l.elbo);
    d.meta = reinterpret_cast<int*>(A + l.meta);

    // small vectors
    const int KK = g.K * g.K;
    h.ecov0.assign(g.effect_covar, g.effect_covar + static_cast<size_t>(g.L) * KK);
    h.rv0.assign(g.resid_var, g.resid_var + g.K);
    auto host_at = [&](size_t arena_off) { return reinterpret_cast<double*>(small_host.data() + (arena_off - small_base)); };
    std::memcpy(host_at(l.ecov0), g.effect_covar, sizeof(double) * g.L * KK);
    std::memcpy(host_at(l.rv0), g.resid_var, sizeof(double) * g.K);
    bool times_zero = true;
    for (int t = 0; t < KK; ++t) {
      host_at(l.adjt)[t] = g.em ? 1.0 : g.adj_times[t];
      host_at(l.adjp)[t] = g.em ? 0.0 : g.adj_plus[t];
      if (host_at(l.adjt)[t] != 0.0) times_zero = false;
    }
    d.skip_first_pass = (!g.em && times_zero) ? 1 : 0;
    if (g.pi) std::memcpy(host_at(l.pi_raw), g.pi, sizeof(double) * g.p);
    d.pi_raw = g.pi ? reinterpret_cast<const double*>(A + l.pi_raw) : nullptr;
    p_max_all = std::max(p_max_all, g.p);

    // matrices and per-row vectors
    for (int k = 0; k < g.K; ++k) {
      const int n = g.n[k];
      double* xtx_k = reinterpret_cast<double*>(A + l.xtx) + static_cast<size_t>(k) * ge.p_pad;
      const bool direct = (g.mat_dtype == opts.storage) && g.standardize == 0 && !hard_storage(opts.storage) && !g.rows[k];
      if (hard_storage(opts.storage)) {
        double* cmean_k = reinterpret_cast<double*>(A + l.cmean) + static_cast<size_t>(k) * ge.p_pad;
        double* cisd_k = reinterpret_cast<double*>(A + l.cisd) + static_cast<size_t>(k) * ge.p_pad;
        h.d_cmean[k] = cmean_k;
        h.d_cisd[k] = cisd_k;
        const double* d_Q = nullptr;
        double* d_M = nullptr;
        if (g.qc[k] > 0) {
          d_Q = reinterpret_cast<const double*>(A + l.Qb[k]);
          d_M = reinterpret_cast<double*>(A + l.Mq[k]);
          SB_CUDA_B(cudaMemcpyAsync(A + l.Qb[k], g.Q[k], sizeof(double) * n * g.qc[k], cudaMemcpyHostToDevice, st));
        }
        d.Qb[k] = d_Q;
        d.Mq[k] = d_M;
        d.qc[k] = g.qc[k];
        SB_CUDA_B(enqueue_prep(g.mat[k], SB_I8, n, g.p, ge.p_pad, g.standardize, A + l.X[k], xtx_k, cmean_k, cisd_k, d_Q,
                               d_M, g.qc[k], g.rows[k], g.src_rows[k]));
      } else if (direct) {
        SB_CUDA_B(cudaMemcpy2DAsync(A + l.X[k], static_cast<size_t>(ge.p_pad) * es, g.mat[k],
                                    static_cast<size_t>(g.p) * es, static_cast<size_t>(g.p) * es, n,
                                    cudaMemcpyHostToDevice, st));
        if (!ss) {
          prep_columns_dispatch(opts.storage, opts.storage, A + l.X[k], ge.p_pad, n, g.p, ge.p_pad, 0, A + l.X[k],
                                xtx_k, st);
        } else if (ge.p_pad > g.p) {
          SB_CUDA_B(cudaMemset2DAsync(A + l.X[k] + static_cast<size_t>(g.p) * es, static_cast<size_t>(ge.p_pad) * es,
                                      0, static_cast<size_t>(ge.p_pad - g.p) * es, n, st));
        }
      } else {
        if (!ss) {
          SB_CUDA_B(enqueue_prep(g.mat[k], g.mat_dtype, n, g.p, ge.p_pad, g.standardize, A + l.X[k], xtx_k, nullptr,
                                 nullptr, nullptr, nullptr, 0, g.rows[k], g.src_rows[k]));
        } else {
          SB_CUDA_B(cudaMemcpyAsync(A + staging_off, g.mat[k], static_cast<size_t>(n) * g.p * elem_size(g.mat_dtype),
                                    cudaMemcpyHostToDevice, st));
          if (g.mat_dtype == SB_F64 && opts.storage == SB_F32)
            launch_pad_rows<double, float>(A + staging_off, n, g.p, ge.p_pad, A + l.X[k], st);
          else if (g.mat_dtype == SB_F32 && opts.storage == SB_F64)
            launch_pad_rows<float, double>(A + staging_off, n, g.p, ge.p_pad, A + l.X[k], st);
          else {
            ctx->err = "unsupported LD dtype / storage combination";
            return fail(SB_EINVAL);
          }
        }
      }
      if (ss || direct) ++launches;
      std::memcpy(host_at((ss ? l.z : l.y) + sizeof(double) * d.row0[k]), g.vec[k], sizeof(double) * n);
      if (ss) {
        double* xty_k = reinterpret_cast<double*>(A + l.y) + d.row0[k];
        const double* z_k = reinterpret_cast<const double*>(A + l.z) + d.row0[k];
        ss_prep.push_back({A + l.X[k], z_k, g.ns[k], g.p, ge.p_pad, xty_k, xtx_k});
      }
    }
  }
  SB_CUDA_B(flush_prep());
  if (!pack_descs.empty()) {
    // bit planes of every (problem, ancestry) from the one-byte matrices the column prologue has just written
    sb::PackDesc* d_pack = reinterpret_cast<sb::PackDesc*>(A + pack_off);
    SB_CUDA_B(cudaMemcpyAsync(d_pack, pack_descs.data(), sizeof(sb::PackDesc) * pack_descs.size(), cudaMemcpyHostToDevice, st));
    for (size_t first_desc = 0; first_desc < pack_descs.size(); first_desc += 32768) {  // grid.y limit
      const unsigned count = static_cast<unsigned>(std::min<size_t>(32768, pack_descs.size() - first_desc));
      sb::pack_b2_kernel<<<dim3(64, count), 256, 0, st>>>(d_pack + first_desc, b->d_ctl + 32);
      ++launches;
    }
    SB_CUDA_B(cudaGetLastError());
  }
  // the batch's small vectors: one copy (pageable source: staged by the runtime before the call returns)
  SB_CUDA_B(cudaMemcpyAsync(A + small_base, small_host.data(), small_bytes, cudaMemcpyHostToDevice, st));
  for (const SsPrep& sp : ss_prep) {  // Xty / diag(XtX) need the z-scores just uploaded
    if (opts.storage == SB_F64)
      sb::prep_ss_kernel<double><<<(sp.p_pad + 127) / 128, 128, 0, st>>>(static_cast<const double*>(sp.ld), sp.z, sp.ns, sp.p,
                                                                       sp.p_pad, sp.xty, sp.xtx);
    else
      sb::prep_ss_kernel<float><<<(sp.p_pad + 127) / 128, 128, 0, st>>>(static_cast<const float*>(sp.ld), sp.z, sp.ns, sp.p,
                                                                      sp.p_pad, sp.xty, sp.xtx);
    ++launches;
  }
  SB_CUDA_B(cudaGetLastError());
  // problem table, ids grouped by K
  {
    std::vector<int> ids;
    ids.reserve(n_probs);
    for (int K = 1; K <= KMAX; ++K)
      for (int variant = 0; variant <= 6; ++variant) {
        const int start = static_cast<int>(ids.size());
        for (int i = 0; i < n_probs; ++i)
          if (gp[i].K == K && launch_variant(opts.storage, ss, geo[i].fast) == variant) ids.push_back(i);
        const int count = static_cast<int>(ids.size()) - start;
        if (count > 0) b->groups.push_back({K, variant, start, count});
      }
    b->ids_host = ids;
    SB_CUDA_B(cudaMemcpyAsync(b->d_ids, ids.data(), sizeof(int) * n_probs, cudaMemcpyHostToDevice, st));
    SB_CUDA_B(cudaMemcpyAsync(b->d_probs, b->pd.data(), sizeof(ProbDev) * n_probs, cudaMemcpyHostToDevice, st));
    // log pi of every problem in one launch (was one launch per locus)
    sb::log_pi_batched_kernel<<<dim3((p_max_all + 255) / 256, n_probs), 256, 0, st>>>(b->d_probs);
    ++launches;
    SB_CUDA_B(cudaGetLastError());
    SB_CUDA_B(cudaEventRecord(ctx->ev[3], st));
    SB_CUDA_B(sync_stream(ctx));
  }
  if (hard_storage(opts.storage)) {
    int bad = 0;
    SB_CUDA_B(read_small(ctx, &bad, b->d_ctl + 32, sizeof(int)));
    if (bad) {
      ctx->err = (bad & 2) ? "SB_B2 storage needs hard-call genotypes 0 / 1 / 2"
                           : "SB_I8 storage needs hard-call genotypes in 0..127 (negative entry found)";
      return fail(SB_EINVAL);
    }
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]);
  b->prologue_ms = ms;
  b->n_launches_create = launches;
  *out = b.release();
  return SB_OK;
#undef SB_CUDA_B
}

// ------------------------------------------------------------------------------------------
// public batch API

extern "C" int sb_batch_create_ind(sb_ctx* ctx, const sb_ind_problem* probs, int n_probs, const sb_opts* opts,
                                   sb_batch** out) {
  if (!ctx || !probs || !out) return SB_EINVAL;
  std::vector<GenericProblem> gp(std::max(0, n_probs));
  for (int i = 0; i < n_probs; ++i) {
    const sb_ind_problem& pr = probs[i];
    int rc = validate_common(ctx, pr, i);
    if (rc != SB_OK) return rc;
    if (!pr.n || !pr.X || !pr.y) {
      ctx->err = "problem " + std::to_string(i) + ": n / X / y missing";
      return SB_EINVAL;
    }
    if (pr.x_dtype != SB_F64 && pr.x_dtype != SB_F32 && pr.x_dtype != SB_I8) {
      ctx->err = "problem " + std::to_string(i) + ": bad x_dtype";
      return SB_EINVAL;
    }
    GenericProblem& g = gp[i];
    g.K = pr.k;
    g.p = pr.p;
    g.L = pr.L;
    g.em = pr.em_update ? 1 : 0;
    g.mat_dtype = pr.x_dtype;
    g.standardize = pr.standardize & 3;
    if (pr.x_rows && !pr.x_src_rows) {
      ctx->err = "problem " + std::to_string(i) + ": x_rows needs x_src_rows";
      return SB_EINVAL;
    }
    for (int k = 0; k < pr.k; ++k) {
      g.n[k] = pr.n[k];
      g.mat[k] = pr.X[k];
      g.vec[k] = pr.y[k];
      g.ns[k] = pr.n[k];
      if (!pr.X[k] || !pr.y[k]) {
        ctx->err = "problem " + std::to_string(i) + ": null X / y pointer";
        return SB_EINVAL;
      }
      if (pr.x_rows && pr.x_rows[k]) {
        g.rows[k] = pr.x_rows[k];
        g.src_rows[k] = pr.x_src_rows[k];
        for (int r = 0; r < pr.n[k]; ++r)
          if (pr.x_rows[k][r] < 0 || pr.x_rows[k][r] >= pr.x_src_rows[k]) {
            ctx->err = "problem " + std::to_string(i) + ": x_rows entry outside the source matrix";
            return SB_EINVAL;
          }
      }
    }
    g.pi = pr.pi;
    g.resid_var = pr.resid_var;
    g.effect_covar = pr.effect_covar;
    g.adj_times = pr.adj_times;
    g.adj_plus = pr.adj_plus;
    if (pr.covar_q) {
      if (!pr.covar_q_cols || !opts || opts->storage != SB_B2) {
        ctx->err = "problem " + std::to_string(i) + ": covar_q needs covar_q_cols and SB_B2 storage";
        return SB_EINVAL;
      }
      for (int k = 0; k < pr.k; ++k) {
        if (!pr.covar_q[k] || pr.covar_q_cols[k] < 1 || pr.covar_q_cols[k] > SB_MAX_COVAR_BASIS) {
          ctx->err = "problem " + std::to_string(i) + ": covar_q_cols must be 1.." + std::to_string(SB_MAX_COVAR_BASIS);
          return SB_EINVAL;
        }
        g.Q[k] = pr.covar_q[k];
        g.qc[k] = pr.covar_q_cols[k];
      }
    }
  }
  return batch_create(ctx, gp, false, opts, out);
}

extern "C" int sb_batch_create_ss(sb_ctx* ctx, const sb_ss_problem* probs, int n_probs, const sb_opts* opts,
                                  sb_batch** out) {
  if (!ctx || !probs || !out) return SB_EINVAL;
  std::vector<GenericProblem> gp(std::max(0, n_probs));
  for (int i = 0; i < n_probs; ++i) {
    const sb_ss_problem& pr = probs[i];
    int rc = validate_common(ctx, pr, i);
    if (rc != SB_OK) return rc;
    if (!pr.ns || !pr.ld || !pr.z) {
      ctx->err = "problem " + std::to_string(i) + ": ns / ld / z missing";
      return SB_EINVAL;
    }
    if (pr.ld_dtype != SB_F64 && pr.ld_dtype != SB_F32) {
      ctx->err = "problem " + std::to_string(i) + ": bad ld_dtype";
      return SB_EINVAL;
    }
    GenericProblem& g = gp[i];
    g.K = pr.k;
    g.p = pr.p;
    g.L = pr.L;
    g.em = pr.em_update ? 1 : 0;
    g.mat_dtype = pr.ld_dtype;
    g.standardize = 0;
    for (int k = 0; k < pr.k; ++k) {
      g.n[k] = pr.p;  // rows of the LD matrix
      g.mat[k] = pr.ld[k];
      g.vec[k] = pr.z[k];
      g.ns[k] = pr.ns[k];
      if (!pr.ld[k] || !pr.z[k]) {
        ctx->err = "problem " + std::to_string(i) + ": null ld / z pointer";
        return SB_EINVAL;
      }
    }
    g.pi = pr.pi;
    g.resid_var = pr.resid_var;
    g.effect_covar = pr.effect_covar;
    g.adj_times = pr.adj_times;
    g.adj_plus = pr.adj_plus;
  }
  return batch_create(ctx, gp, true, opts, out);
}

// Team-size ladder for the phased schedule: hardware clusters up to 16 blocks, then
// co-operative (global-barrier) teams up to the whole device.
static int next_team_size(const sb_ctx* ctx, int S, int remaining, int current) {
  const int ladder[] = {1, 2, 4, 8, 16, 32, 64, S};
  int best = current;
  for (int c : ladder) {
    if (c <= current) continue;
    if (c > 16 && !ctx->coop_ok) break;
    if (c > 1 && c <= 16 && !ctx->cluster_ok) continue;
    if (S / c >= remaining) best = c;
  }
  return best;
}

extern "C" int sb_batch_run(sb_ctx* ctx, sb_batch* b, sb_run_stats* stats) {
  if (!ctx || !b || !b->arena) return SB_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  SB_CUDA(cudaMemsetAsync(b->arena + b->metas_off, 0, sizeof(int) * 8 * b->n_probs, st));
  int launches = 0, teams_used = 0, phases = 0;
  std::vector<int> meta(static_cast<size_t>(b->n_probs) * 8);
  SB_CUDA(cudaEventRecord(ctx->ev[0], st));
  for (const sb_batch::Group& grp : b->groups) {
    const int K = grp.K, g0 = grp.start, g1 = grp.start + grp.count;
    const void* fn = pick_kernel(K, b->storage, b->ss, grp.variant);
    if (!fn) {
      ctx->err = "no kernel for K=" + std::to_string(K);
      return SB_EINVAL;
    }
    SB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, b->smem_bytes));
    SB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));

    // Phased schedule: run with the current team size until half of the teams have run out of
    // work, suspend the loci still in flight at their next iteration boundary, and continue
    // them with larger teams, so the tail of the batch keeps every SM busy.
    std::vector<int> ids(b->ids_host.begin() + g0, b->ids_host.begin() + g1);
    int C = b->team_size;
    int hw = b->hw_cluster;
      while (!ids.empty()) {
      // (re)upload the id list of this phase: a previous run leaves the last phase's list behind
      {
        // (through the pinned control buffer; its first n_probs ints are the id list, the meta table follows)
        unsigned char* ctl = ctl_buffer(ctx, sizeof(int) * 9 * static_cast<size_t>(b->n_probs));
        const void* src = ids.data();
        if (ctl) {
          // (this group's own slice: the asynchronous copy of another group's list may still be reading its slice)
          std::memcpy(ctl + sizeof(int) * g0, ids.data(), sizeof(int) * ids.size());
          src = ctl + sizeof(int) * g0;
        }
        SB_CUDA(cudaMemcpyAsync(b->d_ids + g0, src, sizeof(int) * ids.size(), cudaMemcpyHostToDevice, st));
      }
      cudaLaunchConfig_t cfg{};
      const int block_threads = (b->storage == SB_B2) ? sb::NT_B2 : sb::NT;
      cfg.blockDim = dim3(block_threads);
      cfg.dynamicSmemBytes = b->smem_bytes;
      cfg.stream = st;
      cudaLaunchAttribute attrs[2];
      int na = 0;
      int max_teams;
      if (C == 1) {
        int per_sm = 0;
        SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block_threads, b->smem_bytes));
        max_teams = per_sm * b->sms;
      } else if (hw) {
        attrs[na].id = cudaLaunchAttributeClusterDimension;
        attrs[na].val.clusterDim.x = C;
        attrs[na].val.clusterDim.y = 1;
        attrs[na].val.clusterDim.z = 1;
        ++na;
        cfg.attrs = attrs;
        cfg.numAttrs = na;
        cfg.gridDim = dim3(C);
        int ncl = 0;
        SB_CUDA(cudaOccupancyMaxActiveClusters(&ncl, fn, &cfg));
        max_teams = ncl;
      } else {
        int per_sm = 0;
        SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block_threads, b->smem_bytes));
        max_teams = (per_sm * b->sms) / C;
        attrs[na].id = cudaLaunchAttributeCooperative;
        attrs[na].val.cooperative = 1;
        ++na;
      }
      // one block per SM, except single-block teams of the hard-call kernel, which is built to
      // co-reside (two loci per SM overlap each other's barriers and posterior phases)
      const int per_sm_cap = (b->storage == SB_I8 && C == 1) ? SB_I8_BLOCKS_PER_SM : 1;
      (void)0;
      max_teams = std::min(max_teams, per_sm_cap * b->sms / C);
      if (max_teams < 1) {
        ctx->err = "kernel does not fit on the device (team_size=" + std::to_string(C) + ", smem=" +
                   std::to_string(b->smem_bytes) + ")";
        return SB_EINVAL;
      }
      const int n_ids = static_cast<int>(ids.size());
      const int n_teams = std::min(max_teams, n_ids);
      teams_used = std::max(teams_used, n_teams);
      cfg.gridDim = dim3(n_teams * C);
      cfg.attrs = attrs;
      cfg.numAttrs = na;

      const int bigger = (b->opts.flags & SB_FLAG_SINGLE_PHASE) ? C : next_team_size(ctx, b->sms, 1, C);
      SB_CUDA(cudaMemsetAsync(b->d_queue, 0, sizeof(int) * 64, st));
      SB_CUDA(cudaMemsetAsync(b->d_ctl, 0, sizeof(int) * 64, st));
      SB_CUDA(cudaMemsetAsync(b->d_bar, 0, sizeof(unsigned) * 64 * (2 * ctx->num_sms + 8), st));

      LaunchParams lp{};
      lp.probs = b->d_probs;
      lp.prob_ids = b->d_ids + g0;
      lp.n_ids = n_ids;
      lp.queue = b->d_queue;  // reset before every launch
      lp.team_slot = b->d_slots;
      lp.team_bar = b->d_bar;
      lp.team_size = C;
      lp.hw_cluster = hw;
      lp.max_iter = b->opts.max_iter;
      lp.min_tol = b->opts.min_tol;
      lp.scratch = b->d_scratch;
      lp.scratch_stride = b->scratch_stride;
      lp.off_stat = b->off_stat;
      lp.off_trav = b->off_trav;
      lp.ctl = b->d_ctl;
      lp.n_teams = n_teams;
      lp.allow_drain = (bigger > C) ? 1 : 0;
      void* args[] = {&lp};
      static const bool trace = std::getenv("SUSHIE_B200_TRACE_PHASES") != nullptr;
      if (trace) {  // development aid: wall time of the previous phase ends where this line is printed
        sync_stream(ctx);
        std::fprintf(stderr, "[sushie_b200] phase %d: team_size=%d hw=%d teams=%d loci=%d\n", phases, C, hw, n_teams, n_ids);
      }
      SB_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
      if (trace) {
        cudaEvent_t e0 = ctx->ev[2];
        (void)e0;
        auto t0 = std::chrono::steady_clock::now();
        sync_stream(ctx);
        const double ms_phase = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::fprintf(stderr, "[sushie_b200] phase %d took %.1f ms\n", phases, ms_phase);
      }
      ++launches;
      ++phases;
      if (!lp.allow_drain) break;  // final phase: everything runs to completion
      // which loci were suspended?
      SB_CUDA(read_small(ctx, meta.data(), b->arena + b->metas_off, sizeof(int) * 8 * b->n_probs,
                         sizeof(int) * static_cast<size_t>(b->n_probs)));
      std::vector<int> rest;
      for (int id : ids)
        if (meta[static_cast<size_t>(id) * 8 + 4] != 2) rest.push_back(id);
      ids.swap(rest);
      if (ids.empty()) break;
      const int c_next = next_team_size(ctx, b->sms, static_cast<int>(ids.size()), C);
      if (c_next != C) {
        C = c_next;
        hw = (C > 1 && C <= 16 && ctx->cluster_ok) ? 1 : 0;
      }
    }
  }
  SB_CUDA(cudaEventRecord(ctx->ev[1], st));
  SB_CUDA(sync_stream(ctx));
  SB_CUDA(cudaGetLastError());
  b->ran = true;
  b->phases = phases;
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    stats->device_ms = ms;
    stats->prologue_ms = b->prologue_ms;
    stats->n_launches = launches + b->n_launches_create;
    stats->team_size = b->team_size;
    stats->n_teams = teams_used;
    stats->rows_per_band = b->G_max;
    stats->n_stages = b->NS_max;
    stats->smem_bytes = b->smem_bytes;
    stats->n_run_launches = launches;
    SB_CUDA(read_small(ctx, meta.data(), b->arena + b->metas_off, sizeof(int) * 8 * b->n_probs,
                       sizeof(int) * static_cast<size_t>(b->n_probs)));
    for (int i = 0; i < b->n_probs; ++i) {
      stats->n_iter += meta[i * 8 + 0];
      stats->n_ser += meta[i * 8 + 3];
      stats->algorithmic_bytes += static_cast<double>(meta[i * 8 + 3]) * b->hp[i].bytes_per_ser;
    }
  }
  return SB_OK;
}

// One staged piece of a download: `bytes` at `stage_off` of the pinned buffer belong at `dst` (caller memory).
struct Piece {
  void* dst;
  size_t stage_off, bytes;
};

// pinned -> caller copies of a staging fill.  The destinations are freshly allocated result arrays, so the copy is
// bound by page faults of one thread (~4 GB/s, 4x slower than the PCIe transfer that fed the pinned buffer): large
// fills are split over a few short-lived threads (SUSHIE_B200_COPY_THREADS, default 4; 1 = the calling thread only).
static void copy_pieces(const unsigned char* pinned, const std::vector<Piece>& pieces) {
  static const int max_threads = [] {
    const char* env = std::getenv("SUSHIE_B200_COPY_THREADS");
    const int hw = static_cast<int>(std::thread::hardware_concurrency());
    int n = env ? std::atoi(env) : std::min(4, std::max(1, hw / 8));
    return std::max(1, std::min(n, 16));
  }();
  size_t total = 0;
  for (const Piece& pc : pieces) total += pc.bytes;
  const int nt = static_cast<int>(std::min<size_t>(max_threads, total / (size_t(8) << 20) + 1));
  if (nt <= 1) {
    for (const Piece& pc : pieces) std::memcpy(pc.dst, pinned + pc.stage_off, pc.bytes);
    return;
  }
  // contiguous runs of pieces of about equal bytes per thread
  std::vector<size_t> cut(nt + 1, pieces.size());
  cut[0] = 0;
  size_t acc = 0;
  int t = 1;
  for (size_t i = 0; i < pieces.size() && t < nt; ++i) {
    acc += pieces[i].bytes;
    if (acc >= total * t / nt) cut[t++] = i + 1;
  }
  auto work = [&](int w) {
    for (size_t i = cut[w]; i < cut[w + 1]; ++i) std::memcpy(pieces[i].dst, pinned + pieces[i].stage_off, pieces[i].bytes);
  };
  std::vector<std::thread> pool;
  for (int w = 1; w < nt; ++w) pool.emplace_back(work, w);
  work(0);
  for (std::thread& th : pool) th.join();
}

// pinned staging for downloads (allocated on first use, kept by the context)
static void ensure_pinned(sb_ctx* ctx) {
  const size_t want = size_t(256) << 20;
  if (ctx->pinned_bytes >= want) return;
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr;
  ctx->pinned_bytes = 0;
  if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->pinned), want, cudaHostAllocDefault) == cudaSuccess)
    ctx->pinned_bytes = want;
  else
    cudaGetLastError();  // callers fall back to direct copies
}

extern "C" int sb_batch_fetch(sb_ctx* ctx, sb_batch* b, sb_result* results) {
  if (!ctx || !b || !results || !b->arena) return SB_EINVAL;
  if (!b->ran) {
    ctx->err = "sb_batch_fetch before sb_batch_run";
    return SB_EINVAL;
  }
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  std::vector<int> meta(static_cast<size_t>(b->n_probs) * 8);
  SB_CUDA(read_small(ctx, meta.data(), b->arena + b->metas_off, sizeof(int) * 8 * b->n_probs));
  ensure_pinned(ctx);
  std::vector<Piece> pieces;
  size_t staged = 0;
  auto flush = [&]() -> cudaError_t {
    cudaError_t e = sync_stream(ctx);
    if (e != cudaSuccess) return e;
    copy_pieces(ctx->pinned, pieces);
    pieces.clear();
    staged = 0;
    return cudaSuccess;
  };
  bool direct = false;  // the caller's buffers of the current problem are page-locked already: copy straight into them
  // per-effect arrays: `stride` doubles per effect, written to dst in the caller's effect order (sb_result.effect_order)
  const int32_t* order = nullptr;
  int order_L = 0;
  auto get = [&](double* dst, size_t off, size_t count, size_t stride = 0) -> cudaError_t {
    if (!dst || count == 0) return cudaSuccess;
    const size_t bytes = sizeof(double) * count;
    const bool permute = order && stride;  // effect l of the output = the engine's effect order[l]
    if (direct || !ctx->pinned || bytes > ctx->pinned_bytes) {
      if (!permute) return cudaMemcpyAsync(dst, b->arena + off, bytes, cudaMemcpyDeviceToHost, st);
      cudaError_t e = cudaSuccess;
      for (int l = 0; l < order_L && e == cudaSuccess; ++l)
        e = cudaMemcpyAsync(dst + static_cast<size_t>(l) * stride,
                            b->arena + off + sizeof(double) * stride * static_cast<size_t>(order[l]),
                            sizeof(double) * stride, cudaMemcpyDeviceToHost, st);
      return e;
    }
    if (staged + bytes > ctx->pinned_bytes) {
      cudaError_t e = flush();
      if (e != cudaSuccess) return e;
    }
    // one contiguous device -> pinned copy; the permutation costs nothing extra in the pinned -> caller memcpy
    cudaError_t e = cudaMemcpyAsync(ctx->pinned + staged, b->arena + off, bytes, cudaMemcpyDeviceToHost, st);
    if (!permute) {
      pieces.push_back({reinterpret_cast<unsigned char*>(dst), staged, bytes});
    } else {
      for (int l = 0; l < order_L; ++l)
        pieces.push_back({reinterpret_cast<unsigned char*>(dst + static_cast<size_t>(l) * stride),
                          staged + sizeof(double) * stride * static_cast<size_t>(order[l]), sizeof(double) * stride});
    }
    staged += align_up(bytes, 64);
    return e;
  };
  for (int i = 0; i < b->n_probs; ++i) {
    const HostProb& h = b->hp[i];
    sb_result& r = results[i];
    const int n_iter = meta[i * 8 + 0], inc = meta[i * 8 + 1], fb = meta[i * 8 + 2];
    r.n_iter = n_iter;
    r.elbo_increase = inc;
    r.status = meta[i * 8 + 5] ? SB_STATUS_NONFINITE : SB_STATUS_OK;
    const size_t Lp = static_cast<size_t>(h.L) * h.p;
    const int K = h.K, KK = K * K;
    {
      cudaPointerAttributes attr{};
      const void* probe = r.post_mean_sq ? static_cast<const void*>(r.post_mean_sq) : static_cast<const void*>(r.alpha);
      direct = probe && cudaPointerGetAttributes(&attr, probe) == cudaSuccess && attr.type == cudaMemoryTypeHost;
      cudaGetLastError();
    }
    order = nullptr;
    order_L = h.L;
    if (r.effect_order) {
      std::vector<char> seen(h.L, 0);
      for (int l = 0; l < h.L; ++l) {
        const int e = r.effect_order[l];
        if (e < 0 || e >= h.L || seen[e]) {
          ctx->err = "problem " + std::to_string(i) + ": effect_order is not a permutation of 0..L-1";
          return SB_EINVAL;
        }
        seen[e] = 1;
      }
      order = r.effect_order;
    }
    SB_CUDA(get(r.elbo, h.elbo_off, static_cast<size_t>(n_iter + 1)));
    if (fb < 0) {
      // ELBO failed to improve in the very first iteration: the reference returns the
      // initial state (zero posteriors, initial priors), infer.py:531-532.
      if (r.alpha) std::memset(r.alpha, 0, sizeof(double) * Lp);
      if (r.post_mean) std::memset(r.post_mean, 0, sizeof(double) * Lp * K);
      if (r.post_mean_sq) std::memset(r.post_mean_sq, 0, sizeof(double) * Lp * KK);
      if (r.log_bf) std::memset(r.log_bf, 0, sizeof(double) * Lp);
      if (r.weighted_sum_covar) std::memset(r.weighted_sum_covar, 0, sizeof(double) * h.L * KK);
      if (r.kl) std::memset(r.kl, 0, sizeof(double) * h.L);
      if (r.effect_covar) {
        for (int l = 0; l < h.L; ++l)
          std::memcpy(r.effect_covar + static_cast<size_t>(l) * KK,
                      h.ecov0.data() + static_cast<size_t>(order ? order[l] : l) * KK, sizeof(double) * KK);
      }
      if (r.resid_var) std::memcpy(r.resid_var, h.rv0.data(), sizeof(double) * K);
      continue;
    }
    size_t f[10];
    out_block_fields(h.L, h.p, K, h.out_off[fb], f);
    const size_t pp = static_cast<size_t>(h.p);
    SB_CUDA(get(r.alpha, f[0], Lp, pp));
    SB_CUDA(get(r.post_mean, f[1], Lp * K, pp * K));
    SB_CUDA(get(r.post_mean_sq, f[2], Lp * KK, pp * KK));
    SB_CUDA(get(r.log_bf, f[3], Lp, pp));
    SB_CUDA(get(r.weighted_sum_covar, f[4], static_cast<size_t>(h.L) * KK, KK));
    SB_CUDA(get(r.kl, f[5], h.L, 1));
    SB_CUDA(get(r.effect_covar, f[6], static_cast<size_t>(h.L) * KK, KK));
    SB_CUDA(get(r.resid_var, f[7], K));
  }
  SB_CUDA(flush());
  return SB_OK;
}

extern "C" int sb_batch_destroy(sb_ctx* ctx, sb_batch* b) {
  if (!b) return SB_OK;
  if (ctx) cudaSetDevice(ctx->device);
  if (b->arena) dev_free(ctx, b->arena);
  delete b;
  return SB_OK;
}

extern "C" int sb_batch_storage(sb_batch* b, int prob, int ancestry, void** d_ptr, int64_t* ld_elems) {
  if (!b || prob < 0 || prob >= b->n_probs || ancestry < 0 || ancestry >= b->hp[prob].K) return SB_EINVAL;
  if (d_ptr) *d_ptr = b->hp[prob].d_X[ancestry];
  if (ld_elems) *ld_elems = b->hp[prob].p_pad;
  return SB_OK;
}

// Credible-set selection on the device (make_cs, infer.py:899-915): sort, running sum and cut of every effect of
// every problem in one pass of launches (one per power-of-two width class), then one staged download.
extern "C" int sb_batch_select(sb_ctx* ctx, sb_batch* b, const double* thresholds, sb_selection* out) {
  if (!ctx || !b || !out || !thresholds || !b->arena) return SB_EINVAL;
  if (!b->ran) {
    ctx->err = "sb_batch_select before sb_batch_run";
    return SB_EINVAL;
  }
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int n = b->n_probs;
  // device layout: descriptors | per problem: csum [L][p] | order [L][p] | sizes [L]
  std::vector<sb::SelectDesc> descs(n);
  std::vector<size_t> off_csum(n), off_order(n), off_sizes(n);
  Bump bump;
  const size_t descs_off = bump.take(sizeof(sb::SelectDesc) * n);
  const size_t ids_off = bump.take(sizeof(int) * n);
  const size_t thr_off = bump.take(sizeof(double) * n);
  for (int i = 0; i < n; ++i) {
    const size_t Lp = static_cast<size_t>(b->hp[i].L) * b->hp[i].p;
    off_csum[i] = bump.take(sizeof(double) * Lp);
    off_order[i] = bump.take(sizeof(int) * Lp);
    off_sizes[i] = bump.take(sizeof(int) * b->hp[i].L);
  }
  // width classes: problems sorted by the power of two their p pads to (shared memory is sized per launch)
  auto pow2 = [](int p) {
    int N = 1;
    while (N < p) N <<= 1;
    return N;
  };
  std::vector<int> ids(n);
  for (int i = 0; i < n; ++i) ids[i] = i;
  std::stable_sort(ids.begin(), ids.end(), [&](int a, int c) { return pow2(b->hp[a].p) < pow2(b->hp[c].p); });
  const size_t smem_cap = static_cast<size_t>(ctx->smem_optin) - 1024;
  size_t scratch_bytes = 0;
  for (int i = 0; i < n; ++i) {
    const size_t need = static_cast<size_t>(pow2(b->hp[i].p)) * 12;
    if (need > smem_cap) scratch_bytes = std::max(scratch_bytes, align_up(need, 256) * b->hp[i].L);
  }
  const size_t scratch_off = bump.take(scratch_bytes ? scratch_bytes : 16);
  unsigned char* d_buf = nullptr;
  SB_CUDA(dev_alloc(ctx, reinterpret_cast<void**>(&d_buf), align_up(bump.off, 256)));
  auto bail = [&](cudaError_t e, const char* what) {
    ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
    sync_stream(ctx);
    dev_free(ctx, d_buf);
    return SB_ECUDA;
  };
  for (int i = 0; i < n; ++i) {
    descs[i].csum = out[i].csum ? reinterpret_cast<double*>(d_buf + off_csum[i]) : nullptr;
    descs[i].order = out[i].order ? reinterpret_cast<int*>(d_buf + off_order[i]) : nullptr;
    descs[i].sizes = out[i].sizes ? reinterpret_cast<int*>(d_buf + off_sizes[i]) : nullptr;
  }
  cudaError_t e;
  {
    // the three small tables are contiguous at the head of d_buf: one copy from the pinned control buffer
    const size_t head = thr_off + sizeof(double) * n;
    unsigned char* ctl = ctl_buffer(ctx, head);
    std::vector<unsigned char> fallback;
    if (!ctl) {
      fallback.resize(head);
      ctl = fallback.data();
    }
    std::memcpy(ctl + descs_off, descs.data(), sizeof(sb::SelectDesc) * n);
    std::memcpy(ctl + ids_off, ids.data(), sizeof(int) * n);
    std::memcpy(ctl + thr_off, thresholds, sizeof(double) * n);
    if ((e = cudaMemcpyAsync(d_buf, ctl, head, cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "select H2D tables");
    if (!fallback.empty() && (e = sync_stream(ctx)) != cudaSuccess) return bail(e, "select H2D sync");
  }
  const double* d_thr = reinterpret_cast<const double*>(d_buf + thr_off);
  const sb::SelectDesc* d_descs = reinterpret_cast<const sb::SelectDesc*>(d_buf + descs_off);
  const int* d_ids = reinterpret_cast<const int*>(d_buf + ids_off);
  for (int first = 0; first < n;) {
    const int N = pow2(b->hp[ids[first]].p);
    int last = first, L_max = 1;
    while (last < n && pow2(b->hp[ids[last]].p) == N && last - first < 32768) L_max = std::max(L_max, b->hp[ids[last++]].L);
    const size_t smem = static_cast<size_t>(N) * 12;
    if (smem <= smem_cap) {
      // (static + dynamic shared memory above 48 KB needs the opt-in: N = 4096 is 48 KB of keys + the two counters)
      cudaFuncSetAttribute(sb::select_sets_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      sb::select_sets_kernel<<<dim3(L_max, last - first), sb::SEL_THREADS, smem, st>>>(b->d_probs, d_descs, d_ids + first,
                                                                                   d_thr, nullptr, 0);
    } else {
      // too wide for shared memory: one problem per launch, keys and indices in global scratch (one slot per effect)
      for (int t = first; t < last; ++t)
        sb::select_sets_kernel<<<dim3(b->hp[ids[t]].L, 1), sb::SEL_THREADS, 0, st>>>(
            b->d_probs, d_descs, d_ids + t, d_thr, d_buf + scratch_off, static_cast<long long>(align_up(smem, 256)));
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "select launch");
    first = last;
  }
  // download through the pinned staging buffer
  ensure_pinned(ctx);
  std::vector<Piece> pieces;
  size_t staged = 0;
  auto flush = [&]() -> cudaError_t {
    cudaError_t fe = sync_stream(ctx);
    if (fe != cudaSuccess) return fe;
    copy_pieces(ctx->pinned, pieces);
    pieces.clear();
    staged = 0;
    return cudaSuccess;
  };
  auto get = [&](void* dst, size_t off, size_t bytes) -> cudaError_t {
    if (!dst || bytes == 0) return cudaSuccess;
    if (!ctx->pinned || bytes > ctx->pinned_bytes) return cudaMemcpyAsync(dst, d_buf + off, bytes, cudaMemcpyDeviceToHost, st);
    if (staged + bytes > ctx->pinned_bytes) {
      cudaError_t fe = flush();
      if (fe != cudaSuccess) return fe;
    }
    cudaError_t ge = cudaMemcpyAsync(ctx->pinned + staged, d_buf + off, bytes, cudaMemcpyDeviceToHost, st);
    pieces.push_back({dst, staged, bytes});
    staged += align_up(bytes, 64);
    return ge;
  };
  for (int i = 0; i < n; ++i) {
    const size_t Lp = static_cast<size_t>(b->hp[i].L) * b->hp[i].p;
    if ((e = get(out[i].csum, off_csum[i], sizeof(double) * Lp)) != cudaSuccess) return bail(e, "select D2H csum");
    if ((e = get(out[i].order, off_order[i], sizeof(int) * Lp)) != cudaSuccess) return bail(e, "select D2H order");
    if ((e = get(out[i].sizes, off_sizes[i], sizeof(int) * b->hp[i].L)) != cudaSuccess) return bail(e, "select D2H sizes");
  }
  if ((e = flush()) != cudaSuccess) return bail(e, "select sync");
  dev_free(ctx, d_buf);
  return SB_OK;
}

extern "C" int sb_batch_purity_many(sb_ctx* ctx, sb_batch* b, int n_sets, const int32_t* set_prob,
                                    const int32_t* idx, const int32_t* offsets, double* out_min_abs) {
  if (!ctx || !b || n_sets < 0 || (n_sets > 0 && (!set_prob || !idx || !offsets || !out_min_abs))) return SB_EINVAL;
  if (n_sets == 0) return SB_OK;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  if (offsets[0] != 0) {
    ctx->err = "offsets must start at 0";
    return SB_EINVAL;
  }
  const int total = offsets[n_sets];
  int m_max = 0;
  for (int s = 0; s < n_sets; ++s) {
    const int m = offsets[s + 1] - offsets[s];
    if (m < 1) {
      ctx->err = "empty index set";
      return SB_EINVAL;
    }
    if (set_prob[s] < 0 || set_prob[s] >= b->n_probs) {
      ctx->err = "problem index out of range";
      return SB_EINVAL;
    }
    const int p = b->hp[set_prob[s]].p;
    for (int t = offsets[s]; t < offsets[s + 1]; ++t)
      if (idx[t] < 0 || idx[t] >= p) {
        ctx->err = "SNP index out of range";
        return SB_EINVAL;
      }
    m_max = std::max(m_max, m);
  }
  const int m_pad_max = (m_max + 3) / 4 * 4;
  // rows of the set's columns staged per pass: 32, halved until the largest set fits in shared memory
  int rows = sb::PUR_ROWS;
  const size_t smem_cap = static_cast<size_t>(ctx->smem_optin) - 1024;
  while (!b->ss && rows > 1 && sb::purity_smem_bytes(m_max, rows) > smem_cap) rows >>= 1;
  const size_t smem = b->ss ? 0 : sb::purity_smem_bytes(m_max, rows);
  if (smem > smem_cap) {
    ctx->err = "credible set too large for the purity kernel (" + std::to_string(m_max) + " SNPs)";
    return SB_EINVAL;
  }
  unsigned char* d_buf = nullptr;
  const size_t idx_b = align_up(sizeof(int) * total, 256), off_b = align_up(sizeof(int) * (n_sets + 1), 256);
  const size_t prob_b = align_up(sizeof(int) * n_sets, 256);
  const size_t out_n = static_cast<size_t>(n_sets) * sb::KMAX;
  SB_CUDA(dev_alloc(ctx, reinterpret_cast<void**>(&d_buf), idx_b + off_b + prob_b + align_up(sizeof(double) * out_n, 256)));
  int* d_idx = reinterpret_cast<int*>(d_buf);
  int* d_off = reinterpret_cast<int*>(d_buf + idx_b);
  int* d_prob = reinterpret_cast<int*>(d_buf + idx_b + off_b);
  double* d_out = reinterpret_cast<double*>(d_buf + idx_b + off_b + prob_b);
  auto bail = [&](cudaError_t e, const char* what) {
    ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
    sync_stream(ctx);
    dev_free(ctx, d_buf);
    return SB_ECUDA;
  };
  cudaError_t e;
  // index lists up and results down through the pinned control buffer (see sb_ctx::ctl): [idx | offsets | ids | out]
  const size_t out_ctl = idx_b + off_b + prob_b;
  unsigned char* ctl = ctl_buffer(ctx, out_ctl + sizeof(double) * out_n);
  const void *h_idx = idx, *h_off = offsets, *h_prob = set_prob;
  if (ctl) {
    std::memcpy(ctl, idx, sizeof(int) * total);
    std::memcpy(ctl + idx_b, offsets, sizeof(int) * (n_sets + 1));
    std::memcpy(ctl + idx_b + off_b, set_prob, sizeof(int) * n_sets);
    h_idx = ctl;
    h_off = ctl + idx_b;
    h_prob = ctl + idx_b + off_b;
  }
  if ((e = cudaMemcpyAsync(d_idx, h_idx, sizeof(int) * total, cudaMemcpyHostToDevice, st)) != cudaSuccess)
    return bail(e, "purity H2D idx");
  if ((e = cudaMemcpyAsync(d_off, h_off, sizeof(int) * (n_sets + 1), cudaMemcpyHostToDevice, st)) != cudaSuccess)
    return bail(e, "purity H2D offsets");
  if ((e = cudaMemcpyAsync(d_prob, h_prob, sizeof(int) * n_sets, cudaMemcpyHostToDevice, st)) != cudaSuccess)
    return bail(e, "purity H2D problem ids");
  if ((e = cudaMemsetAsync(d_out, 0, sizeof(double) * out_n, st)) != cudaSuccess) return bail(e, "purity memset");
  // grid.x is limited only by 2^31 - 1; blocks of ancestries a problem does not have exit at once
  dim3 grid(n_sets, sb::KMAX);
#define SB_PUR(XT, SSV)                                                                                        \
  do {                                                                                                         \
    auto kf = sb::purity_kernel<XT, SSV>;                                                                      \
    if (smem > 48 * 1024) cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    kf<<<grid, 256, smem, st>>>(b->d_probs, d_prob, d_idx, d_off, m_pad_max, rows, d_out);                     \
  } while (0)
  int n_max_rows = 1;
  if (hard_storage(b->storage))
    for (int s2 = 0; s2 < n_sets; ++s2) {
      const HostProb& h = b->hp[set_prob[s2]];
      for (int k = 0; k < h.K; ++k) n_max_rows = std::max(n_max_rows, h.n[k]);
    }
  const int nq_max = (n_max_rows + 3) / 4;
  const size_t smem_i8 = sb::purity_i8_smem_bytes(m_pad_max, nq_max);
  static const bool no_dp4a = std::getenv("SUSHIE_B200_PURITY_FP64") != nullptr;  // development: the fp64 kernel
  if (hard_storage(b->storage) && smem_i8 <= smem_cap && !no_dp4a) {
    // exact integer Gram matrices from columns gathered once (dp4a)
    cudaFuncSetAttribute(sb::purity_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_i8));
    sb::purity_i8_kernel<<<grid, 256, smem_i8, st>>>(b->d_probs, d_prob, d_idx, d_off, m_pad_max, nq_max, d_out);
  } else if (hard_storage(b->storage)) {
    SB_PUR(int8_t, false);
  } else if (b->storage == SB_F64) {
    if (b->ss)
      SB_PUR(double, true);
    else
      SB_PUR(double, false);
  } else {
    if (b->ss)
      SB_PUR(float, true);
    else
      SB_PUR(float, false);
  }
#undef SB_PUR
  if ((e = cudaGetLastError()) != cudaSuccess) return bail(e, "purity launch");
  if ((e = cudaMemcpyAsync(ctl ? static_cast<void*>(ctl + out_ctl) : static_cast<void*>(out_min_abs), d_out,
                           sizeof(double) * out_n, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
    return bail(e, "purity D2H");
  if ((e = sync_stream(ctx)) != cudaSuccess) return bail(e, "purity sync");
  if (ctl) std::memcpy(out_min_abs, ctl + out_ctl, sizeof(double) * out_n);
  dev_free(ctx, d_buf);
  return SB_OK;
}

extern "C" int sb_batch_purity(sb_ctx* ctx, sb_batch* b, int prob, const int32_t* idx, const int32_t* offsets,
                               int n_sets, double* out_min_abs) {
  if (!ctx || !b || !idx || !offsets || !out_min_abs || prob < 0 || prob >= b->n_probs || n_sets < 0)
    return SB_EINVAL;
  if (n_sets == 0) return SB_OK;
  const int K = b->hp[prob].K;
  std::vector<int32_t> set_prob(static_cast<size_t>(n_sets), prob);
  std::vector<double> wide(static_cast<size_t>(n_sets) * sb::KMAX);
  const int rc = sb_batch_purity_many(ctx, b, n_sets, set_prob.data(), idx, offsets, wide.data());
  if (rc != SB_OK) return rc;
  for (int s = 0; s < n_sets; ++s)
    for (int k = 0; k < K; ++k) out_min_abs[s * K + k] = wide[static_cast<size_t>(s) * sb::KMAX + k];
  return SB_OK;
}

// ------------------------------------------------------------------------------------------
// one-shot entry points

template <typename P, typename CreateFn>
static int one_shot(sb_ctx* ctx, const P* probs, int n, const sb_opts* opts, sb_result* results, sb_run_stats* stats,
                    CreateFn create) {
  sb_batch* b = nullptr;
  int rc = create(ctx, probs, n, opts, &b);
  if (rc != SB_OK) return rc;
  rc = sb_batch_run(ctx, b, stats);
  if (rc == SB_OK) rc = sb_batch_fetch(ctx, b, results);
  sb_batch_destroy(ctx, b);
  return rc;
}

extern "C" int sb_ibss_ind(sb_ctx* ctx, const sb_ind_problem* probs, int n_probs, const sb_opts* opts,
                           sb_result* results, sb_run_stats* stats) {
  if (!ctx || !probs || !results) return SB_EINVAL;
  return one_shot(ctx, probs, n_probs, opts, results, stats, sb_batch_create_ind);
}

extern "C" int sb_ibss_ss(sb_ctx* ctx, const sb_ss_problem* probs, int n_probs, const sb_opts* opts,
                          sb_result* results, sb_run_stats* stats) {
  if (!ctx || !probs || !results) return SB_EINVAL;
  return one_shot(ctx, probs, n_probs, opts, results, stats, sb_batch_create_ss);
}
