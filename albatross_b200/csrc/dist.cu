// One process per GPU: distributed exact GP over NCCL (NVLink 5 / NVSwitch).
//
// The reference has no distributed (or even multi-process) path; its only parallelism is a pthread
// pool (third_party/ThreadPool) over Gram column blocks (covariance_functions/callers.hpp:134-166)
// and over CV groups (utils/async_utils.hpp:75-189).  The analogue built here:
//
//   * ab_dist_gp_fit   - K is generated directly in block-column-cyclic layout (block column j on
//                        rank j % world, every rank holds all features: no Gram collective), then a
//                        right-looking blocked Cholesky over three streams per rank: the panel
//                        pipeline (high priority) applies every arriving panel to the next block
//                        column the rank owns and factors it the moment it is complete (potrf of
//                        the diagonal block + TRSM of the rows below, packed); the broadcasts
//                        (ncclBroadcast) run on their own stream into a ring of packed-panel
//                        buffers; the update stream applies panels in PAIRS to all the other block
//                        columns of the rank in one DSYRK/DGEMM launch of depth 2 nb on the FP64
//                        tensor pipe.  On NVSwitch every GPU reaches every peer at full bandwidth,
//                        so the 1 x P process grid (a block-cyclic layout with P_r = 1) moves
//                        N^2/2 * 8 bytes per GPU in total — 69 GB at N = 131 072, ~0.2 s at
//                        measured broadcast rates against >3 s of DMMA work — and keeps every
//                        trailing update a single tall GEMM; a P_r > 1 grid would only pay off
//                        across nodes.  AB_DIST_SCHEDULE=lookahead1 keeps the round-1 schedule
//                        (two buffers, the next owner prepares its panel inside the step).
//   * forward / backward block substitution for information = K^-1 y, one small reduce / broadcast
//     per block column; log|K| and y^T K^-1 y by all-reduce of two doubles.
//   * ab_dist_gram_rows - row-block sharded Gram build.
//   * ab_dist_gp_cv     - leave-one-group-out folds sharded over ranks.
//
// libnccl.so.2 is opened lazily (dlopen) so that single-GPU users have no NCCL dependency.
#include "internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

struct ab_dist_factor_s {
  int64_t n = 0;
  int64_t nb = 0;
  int64_t nblk = 0;
  int rank = 0;
  int world = 1;
  ab_matrix_s *A = nullptr; // n x (nloc * nb): the block columns owned by this rank, L after fit
  double *dinv = nullptr;   // per local block column: explicit inverses of its LEAF diagonal blocks
  size_t dinv_bytes = 0;
  int64_t dinv_stride = 0; // doubles per block column
  int64_t bad_pivot = -1;
};

namespace ab {
namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *) = nullptr; // optional
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr; // optional
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int,
                         ncclComm_t, cudaStream_t) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.lib != nullptr) {
    return AB_OK;
  }
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (lib == nullptr) {
    lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  }
  if (lib == nullptr) {
    set_error("cannot load libnccl.so.2: %s", dlerror());
    return AB_ERR_NCCL;
  }
#define AB_NCCL_SYM(field, name)                                                                \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));                    \
  if (g_nccl.field == nullptr) {                                                                \
    set_error("libnccl lacks %s", name);                                                        \
    return AB_ERR_NCCL;                                                                         \
  }
  AB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  AB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  AB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  AB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  AB_NCCL_SYM(AllReduce, "ncclAllReduce")
  AB_NCCL_SYM(Broadcast, "ncclBroadcast")
  AB_NCCL_SYM(Reduce, "ncclReduce")
#undef AB_NCCL_SYM
  g_nccl.CommInitRankConfig =
      reinterpret_cast<decltype(g_nccl.CommInitRankConfig)>(dlsym(lib, "ncclCommInitRankConfig"));
  g_nccl.CommSplit = reinterpret_cast<decltype(g_nccl.CommSplit)>(dlsym(lib, "ncclCommSplit"));
  g_nccl.lib = lib;
  return AB_OK;
}

#define AB_NCCL(expr)                                                                           \
  do {                                                                                          \
    ncclResult_t _r = (expr);                                                                   \
    if (_r != ncclSuccess) {                                                                    \
      ab::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));   \
      return AB_ERR_NCCL;                                                                       \
    }                                                                                           \
  } while (0)

ncclComm_t comm_of(const ab_handle_s *h) { return static_cast<ncclComm_t>(h->comm); }

// Second communicator of the same ranks WITHOUT the CTA cap (ab_dist_init): used for the panel broadcasts of
// the chain-bound tail of the factorisation, where the broadcast is on the critical path and the update stream
// has SMs to spare.  Keyed by handle; absent when NCCL has no ncclCommSplit or the cap is off.
std::map<const ab_handle_s *, ncclComm_t> g_tail_comm;
std::mutex g_tail_comm_mu;
ncclComm_t tail_comm_of(const ab_handle_s *h) {
  std::lock_guard<std::mutex> lock(g_tail_comm_mu);
  auto it = g_tail_comm.find(h);
  return it == g_tail_comm.end() ? nullptr : it->second;
}

// out[i] = a[i] - b[i]
__global__ void sub_kernel(const double *a, const double *b, int64_t n, double *out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) {
    out[i] = a[i] - b[i];
  }
}

__global__ void __launch_bounds__(1024) sum_sq_and_logdiag_kernel(const double *z, int64_t n,
                                                                  const double *logs,
                                                                  int64_t nlogs, double *out) {
  __shared__ double red[2][32];
  double a = 0., b = 0.;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    a = fma(z[i], z[i], a);
  }
  for (int64_t i = threadIdx.x; i < nlogs; i += 1024) {
    b += logs[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = a;
    red[1][threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0., tb = 0.;
    for (int w = 0; w < 32; ++w) {
      ta += red[0][w];
      tb += red[1][w];
    }
    out[0] = ta;
    out[1] = tb;
  }
}

void free_dist_factor(ab_handle_s *h, ab_dist_factor_s *f) {
  if (f == nullptr) {
    return;
  }
  matrix_delete(h, f->A);
  dev_release(h, f->dinv, f->dinv_bytes);
  delete f;
}

} // namespace

int dist_rank(const ab_handle_s *h) { return h->rank; }
int dist_world(const ab_handle_s *h) { return h->world; }

int dist_allreduce_sum(ab_handle_s *h, double *d_buf, int64_t count) {
  if (h->world <= 1 || count <= 0) {
    return AB_OK;
  }
  AB_REQUIRE(h->comm != nullptr, "distributed group not initialised");
  AB_NCCL(g_nccl.AllReduce(d_buf, d_buf, static_cast<size_t>(count), ncclDouble, ncclSum,
                           comm_of(h), h->stream));
  return AB_OK;
}

int64_t dist_total(ab_handle_s *h, int64_t local) {
  if (h->world <= 1) {
    return local;
  }
  h->h_scalars[32] = static_cast<double>(local);
  cudaMemcpyAsync(h->d_scalars + 32, h->h_scalars + 32, sizeof(double), cudaMemcpyHostToDevice,
                  h->stream);
  if (dist_allreduce_sum(h, h->d_scalars + 32, 1) != AB_OK) {
    return local;
  }
  cudaMemcpyAsync(h->h_scalars + 32, h->d_scalars + 32, sizeof(double), cudaMemcpyDeviceToHost,
                  h->stream);
  cudaStreamSynchronize(h->stream);
  return static_cast<int64_t>(h->h_scalars[32] + 0.5);
}

namespace {

// The broadcast stream and the events that tie it to the update / panel streams (created on first use: a
// world-1 group runs the same schedule without NCCL).
int ensure_dist_streams(ab_handle_s *h) {
  if (h->comm_stream != nullptr) {
    return AB_OK;
  }
  int lo = 0, hi = 0;
  AB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  AB_CUDA(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi));
  AB_CUDA(cudaEventCreateWithFlags(&h->ev_bcast[0], cudaEventDisableTiming));
  AB_CUDA(cudaEventCreateWithFlags(&h->ev_bcast[1], cudaEventDisableTiming));
  AB_CUDA(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  AB_CUDA(cudaEventCreateWithFlags(&h->ev_free, cudaEventDisableTiming));
  return AB_OK;
}

// Launch helpers issue on h->stream; the panel chain borrows it for a scope.
struct StreamSwap {
  StreamSwap(ab_handle_s *h, cudaStream_t s) : h_(h), saved_(h->stream) { h->stream = s; }
  ~StreamSwap() { h_->stream = saved_; }
  ab_handle_s *h_;
  cudaStream_t saved_;
};

// Per-step timing events of the distributed factorisation (grown on demand, owned by the process).
struct DistEvents {
  std::vector<cudaEvent_t> wait_begin, wait_end, panel_begin, panel_end;
  std::vector<cudaEvent_t> arrived, bulkdone, pdone; // per panel: broadcast landed / S done with it / PS done
  std::vector<cudaEvent_t> coldone; // per panel: S's single-column update with it is done (paired schedule)
  std::vector<char> owned; // panel k was factored by this rank in the most recent fit
  cudaEvent_t factor_end = nullptr;
  int64_t steps = 0;
};

std::map<ab_handle_s *, DistEvents> g_dist_events;
std::mutex g_dist_events_mu;

DistEvents &dist_events_existing(ab_handle_s *h) {
  std::lock_guard<std::mutex> lock(g_dist_events_mu);
  return g_dist_events[h];
}

DistEvents &dist_events(ab_handle_s *h, int64_t nblk) {
  std::lock_guard<std::mutex> lock(g_dist_events_mu);
  DistEvents &ev = g_dist_events[h];
  if (ev.factor_end == nullptr) {
    cudaEventCreate(&ev.factor_end);
  }
  while (static_cast<int64_t>(ev.wait_begin.size()) < nblk) {
    cudaEvent_t e[4];
    for (auto &x : e) {
      cudaEventCreate(&x);
    }
    ev.wait_begin.push_back(e[0]);
    ev.wait_end.push_back(e[1]);
    ev.panel_begin.push_back(e[2]);
    ev.panel_end.push_back(e[3]);
    cudaEvent_t q[4];
    for (auto &x : q) {
      cudaEventCreateWithFlags(&x, cudaEventDisableTiming);
    }
    ev.arrived.push_back(q[0]);
    ev.bulkdone.push_back(q[1]);
    ev.pdone.push_back(q[2]);
    ev.coldone.push_back(q[3]);
  }
  ev.steps = nblk;
  ev.owned.assign(static_cast<size_t>(nblk), 0);
  return ev;
}

// The distributed fit proper.  F: dim x n device features (replicated); d_y, d_yvar: device vectors.
int dist_fit_impl(ab_handle_s *h, const DevProg &P, const ab_matrix_s *F, const double *d_y,
                  const double *d_yvar, int64_t n, int64_t nb, ab_dist_factor_s *fac,
                  double *d_info, double *nll_out) {
  Scope sc(h);
  const int W = h->world, me = h->rank;
  const int dim = static_cast<int>(F->rows);
  const int64_t nblk = (n + nb - 1) / nb;
  const int64_t nloc = (nblk - me + W - 1) / W; // block columns j with j % W == me
  fac->n = n;
  fac->nb = nb;
  fac->nblk = nblk;
  fac->rank = me;
  fac->world = W;
  AB_TRY(matrix_new(h, n, std::max<int64_t>(nloc, 1) * nb, &fac->A));
  fac->dinv_stride = (nb / LEAF) * LEAF * LEAF;
  fac->dinv_bytes = static_cast<size_t>(std::max<int64_t>(nloc, 1) * fac->dinv_stride) * sizeof(double);
  {
    void *p = nullptr;
    AB_TRY(dev_alloc(h, fac->dinv_bytes, &p));
    fac->dinv = static_cast<double *>(p);
  }
  const MatView A = view(fac->A);
  auto width = [&](int64_t j) { return std::min(nb, n - j * nb); };
  auto colblk = [&](int64_t j) { return A.sub(0, (j / W) * nb); };
  auto dinv_of = [&](int64_t j) { return fac->dinv + (j / W) * fac->dinv_stride; };

  // ---- Gram, generated in place in the cyclic layout: rows >= j*nb of block column j -----------
  phase_begin(h, PH_GRAM);
  for (int64_t j = me; j < nblk; j += W) {
    const int64_t r0 = j * nb;
    const MatView dst = colblk(j).sub(r0, 0);
    AB_TRY(gram_into(h, P, dim, false, F->d + r0 * F->ld, F->ld, n - r0, F->d + r0 * F->ld, F->ld,
                     width(j), dst.p, dst.ld, 0u));
    if (d_yvar != nullptr) {
      AB_TRY(add_diag(h, dst, width(j), d_yvar + r0));
    }
  }
  // pivot floors from the ORIGINAL diagonal (linalg.cu potrf): a block's diagonal is ~0 after the trailing
  // updates when the matrix is singular, so the floor must be taken now
  void *d_floor = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(std::max<int64_t>(nloc, 1) * nb) * sizeof(double), &d_floor));
  for (int64_t j = me; j < nblk; j += W) {
    AB_TRY(pivot_floor(h, colblk(j).sub(j * nb, 0), width(j), static_cast<double *>(d_floor) + (j / W) * nb));
  }
  phase_end(h, PH_GRAM);

  // ---- factorisation --------------------------------------------------------------------------
  // Three streams per rank: S (h->stream) runs the trailing updates, PS (high priority) the panel chain of
  // the block column this rank owns next — leaf factorisations and small GEMMs that use 1-2 % of the SMs —
  // and CS (high priority) the broadcasts.  Look-ahead 1: at step k the owner of block column k+1 updates
  // that column first (on S), then factors and packs it on PS while S goes on with the rest of update k,
  // so that neither the panel chain nor the broadcast is ever on the critical path of the DMMA work.
  phase_begin(h, PH_FACTOR);
  AB_TRY(ensure_panel_stream(h));
  AB_TRY(ensure_dist_streams(h));
  cudaStream_t S = h->stream, PS = h->panel_stream, CS = h->comm_stream;
  const int64_t ldp_max = round_up(n, 2);
  // AB_DIST_SCHEDULE=lookahead1 selects the round-1 style schedule (two panel buffers, the owner of column
  // k+1 prepares it inside step k on the update stream) for comparison; default: the decoupled pipeline
  bool pipelined = true;
  if (const char *e = std::getenv("AB_DIST_SCHEDULE")) {
    pipelined = std::strcmp(e, "lookahead1") != 0;
  }
  // packed-panel buffers: 2 on two ranks (measured best: 1458 vs 1558 ms with 4 at N = 65 536), 4 beyond
  // (3099 vs 3126 with 2 at N = 131 072 on 8 GPUs; 8 buffers: 3345 — a chain far ahead of the updates competes
  // with them for HBM); profiles/r02i_*, r02j_*, r02k_*
  int64_t NBUF = pipelined ? (W <= 2 ? 2 : 4) : 2;
  // Paired updates (default in the pipelined schedule; AB_DIST_PAIR=0 for the one-panel-per-launch form):
  // the update stream applies panels 2q and 2q+1 in ONE launch of k-depth 2 nb.  The DMMA kernel pays a
  // fixed cost per output tile (pipeline fill, read-modify-write of C), so a trailing update of depth 512
  // runs 3-4 % below one of depth 1024 (world-1 runs of this routine at N = 65 536: 32.3 vs 33.5 TFLOP/s,
  // profiles/r02p_dist_w1.txt) while the panel CHAIN wants narrow panels (it is sequential across the ranks).
  // Pairing keeps the chain at nb and gives the bulk update 2 nb.  The two panels of a pair lie side by side
  // in one buffer with a common leading dimension and row origin (block row 2q), so that [P_2q | P_2q+1] is
  // one operand; panel 2q+1 leaves its first nb rows unused.
  bool pairing = pipelined;
  if (const char *e = std::getenv("AB_DIST_PAIR")) {
    pairing = pairing && e[0] != '0';
  }
  if (const char *e = std::getenv("AB_DIST_PCOL")) {
    pairing = pairing && e[0] != '0';
  }
  if (pairing) {
    NBUF = W <= 2 ? 4 : 6; // in panels: 2 / 3 pair buffers
  }
  if (const char *e = std::getenv("AB_DIST_NBUF")) {
    NBUF = std::max<int64_t>(2, std::min<int64_t>(16, std::atoll(e)));
  }
  if (!pipelined) {
    NBUF = 2;
  }
  if (pairing) {
    NBUF = std::max<int64_t>(4, NBUF + (NBUF & 1)); // whole pair buffers, at least two of them
  }
  const int64_t nslots = pairing ? NBUF / 2 : NBUF;
  std::vector<void *> pb(static_cast<size_t>(nslots), nullptr);
  const size_t pbytes =
      static_cast<size_t>(ldp_max) * static_cast<size_t>(nb) * sizeof(double) * (pairing ? 2 : 1);
  for (auto &b : pb) {
    AB_TRY(sc.alloc(pbytes, &b));
  }
  void *d_bad = nullptr;
  AB_TRY(sc.alloc(static_cast<size_t>(nblk) * sizeof(int), &d_bad));
  {
    std::vector<int> init(static_cast<size_t>(nblk), INT_MAX);
    AB_CUDA(cudaMemcpyAsync(d_bad, init.data(), init.size() * sizeof(int), cudaMemcpyHostToDevice, S));
    AB_CUDA(cudaStreamSynchronize(S)); // `init` is a stack temporary
  }
  // accounting (ab_dist_fit_breakdown): time S spends waiting for a panel, time of the panel chains
  DistEvents &ev = dist_events(h, nblk);
  // the packed panel k as a view whose row 0 is global row origin(k): k * nb, or the pair's first row
  auto origin = [&](int64_t k) { return (pairing ? (k & ~int64_t(1)) : k) * nb; };
  auto panel = [&](int64_t k) {
    const int64_t ld = round_up(n - origin(k), 2);
    if (!pairing) {
      return MatView{static_cast<double *>(pb[static_cast<size_t>(k % NBUF)]), ld};
    }
    return MatView{static_cast<double *>(pb[static_cast<size_t>((k / 2) % nslots)]) + (k & 1) * nb * ld, ld};
  };
  // panels k and k+1 (k even) go to the bulk update as one operand of depth 2 nb
  auto paired = [&](int64_t k) {
    const int64_t head = k & ~int64_t(1);
    return pairing && head + 1 < nblk && width(head + 1) == nb;
  };
  // owner only, on PS: factor block column k (diagonal potrf + TRSM of the rows below) and pack it
  auto factor_and_pack = [&](int64_t k) -> int {
    StreamSwap swap(h, PS);
    const int64_t r0 = k * nb, wk = width(k), hk = n - r0;
    const MatView D = colblk(k).sub(r0, 0);
    AB_CUDA(cudaEventRecord(ev.panel_begin[k], PS));
    ev.owned[static_cast<size_t>(k)] = 1;
    AB_TRY(potrf(h, D, wk, dinv_of(k), static_cast<int *>(d_bad) + k,
                 static_cast<double *>(d_floor) + (k / W) * nb));
    AB_TRY(trsm_right_lower_T(h, D, dinv_of(k), wk, D.sub(wk, 0), hk - wk));
    const MatView Pk = panel(k).sub(r0 - origin(k), 0);
    AB_CUDA(cudaMemcpy2DAsync(Pk.p, Pk.ld * sizeof(double), D.p, D.ld * sizeof(double),
                              static_cast<size_t>(hk) * sizeof(double), static_cast<size_t>(wk),
                              cudaMemcpyDeviceToDevice, PS));
    AB_CUDA(cudaEventRecord(ev.panel_end[k], PS));
    return AB_OK;
  };
  // A[j*nb:, block j] -= L[j*nb:, block k] L[block j rows, block k]^T   (j > k, j owned by me).
  // One column at a time (the look-ahead column, and the fallback for shapes the TMA kernel does not take):
  auto update = [&](int64_t j, int64_t k) -> int {
    const MatView Pk = panel(k);
    const int64_t off = j * nb - origin(k);
    return gemm(h, GEMM_TRANS_B, n - j * nb, width(j), width(k), -1., Pk.sub(off, 0), Pk.sub(off, 0), 1.,
                colblk(j).sub(j * nb, 0));
  };
  // ... and ALL owned block columns j >= jfirst (jfirst owned by me) in ONE launch: they are contiguous in
  // the local matrix, their B operands are the row blocks jfirst-k, jfirst-k+W, ... of the packed panel
  // (CyclicB), and tiles above the stretched diagonal are skipped.  Per-column launches cost a partial last
  // wave each (measured at N = 65 536 on 2 GPUs: 0.87 of the 1-GPU rate, 214 ms of tails in 1.6 s; the
  // launches of one step carry 1/W of a full trailing update, so the loss grows with W).
  // count = -1: every owned column from jfirst on; count = 1: block column jfirst alone.
  // npanels = 2: panels k and k+1 of a pair (k even) in one launch of depth 2 nb.
  auto update_cols = [&](int64_t jfirst, int64_t count, int64_t k, int64_t npanels = 1) -> int {
    if (jfirst >= nblk) {
      return AB_OK;
    }
    const MatView Pk = panel(k);
    const int64_t R0 = jfirst * nb, m = n - R0, l0 = jfirst / W;
    const int64_t llast = count < 0 ? nloc - 1 : std::min<int64_t>(l0 + count - 1, nloc - 1);
    const int64_t jlast = me + llast * W;
    const int64_t ncols = (llast - l0) * nb + width(jlast);
    CyclicB cyc;
    cyc.blk = nb;
    cyc.stride = static_cast<int64_t>(W) * nb;
    cyc.row0 = R0 - origin(k);
    cyc.rows = n - origin(k);
    int st = AB_ERR_UNSUPPORTED;
    if (gemm_tma_enabled() && width(k + npanels - 1) == nb) {
      st = gemm_nt_tma(h, true, m, ncols, npanels * nb, -1., Pk.sub(R0 - origin(k), 0), Pk, 1.,
                       A.sub(R0, l0 * nb), &cyc);
    }
    if (st != AB_ERR_UNSUPPORTED) {
      return st;
    }
    for (int64_t kk = k; kk < k + npanels; ++kk) {
      for (int64_t j = jfirst; j <= jlast; j += W) {
        AB_TRY(update(j, kk));
      }
    }
    return AB_OK;
  };
  auto update_from = [&](int64_t jfirst, int64_t k) -> int { return update_cols(jfirst, -1, k); };
  // enqueue the broadcast of panel k on CS; afterwards ev_bcast[k % 2] says "panel k is in pb[k % 2]"
  auto bcast = [&](int64_t k) -> int {
    const MatView Pk = panel(k);
    const int root = static_cast<int>(k % W);
    if (root == me) {
      AB_CUDA(cudaStreamWaitEvent(CS, ev.panel_end[k], 0));
    } else {
      // the buffer was last read by the updates of step k-2, all enqueued on S by now
      AB_CUDA(cudaEventRecord(h->ev_free, S));
      AB_CUDA(cudaStreamWaitEvent(CS, h->ev_free, 0));
    }
    if (W > 1) {
      AB_NCCL(g_nccl.Broadcast(Pk.p, Pk.p, static_cast<size_t>(Pk.ld * width(k)), ncclDouble, root,
                               comm_of(h), CS));
    }
    AB_CUDA(cudaEventRecord(h->ev_bcast[k % 2], CS));
    return AB_OK;
  };

  if (pipelined) {
    // ---- decoupled panel pipeline ---------------------------------------------------------------------
    // The chain "apply panel k to the next column I own, factor it when it is complete, broadcast it" is
    // sequential across the ranks and costs ~24 ms per step at N = 131 072 while a rank's share of the
    // trailing update shrinks quadratically (43 ms at step 0 on 8 GPUs): from step ~56 of 128 on the chain,
    // not the DMMA work, sets the pace (measured: 155 ms of 3.2 s waiting, profiles/r02g_*).  So the chain
    // gets its own high-priority stream PS on every rank and never waits for the bulk: for every panel k,
    // PS updates the first owned column after k (jstar) the moment the panel arrives and factors it as soon
    // as it is complete, while S applies the panel to the owned columns beyond jstar in one launch.  S may
    // lag up to NBUF - 1 panels behind (NBUF packed-panel buffers), which also lets ranks whose share of a
    // step is one block column larger (4 % of the work at W = 8) drift instead of stalling the others.
    std::vector<cudaEvent_t> &arrived = ev.arrived, &bulkdone = ev.bulkdone, &pdone = ev.pdone;
    // Chain-bound tail: with r rows left a rank's share of the bulk update takes r^2 nb / (W R) (R = 33 TFLOP/s)
    // and one chain step ~0.6 ms + r * 5.9e-8 s (column update 2 nb^2 / 31 TFLOP/s + TRSM nb^2 / 12 TFLOP/s +
    // broadcast 8 nb / 200 GB/s per row at nb = 512, potrf 0.56 ms; tools/chain_bench.py,
    // profiles/r02w_chain_bench.txt).  From the step where the chain is the longer of the two, the broadcasts go
    // through the uncapped communicator (AB_DIST_TAIL_COMM=0: never).  N = 131 072 on 8 GPUs: 3003 -> 2985 ms
    // (profiles/r02z_*).  Measured and dropped in the same run: factoring the diagonal block while a second
    // panel stream updates the rows below it (no change: 3002.9 vs 3003.4 ms).
    ncclComm_t tail_comm = W > 1 ? tail_comm_of(h) : nullptr;
    if (const char *e = std::getenv("AB_DIST_TAIL_COMM")) {
      if (e[0] == '0') {
        tail_comm = nullptr;
      }
    }
    int64_t tail_from = nblk;
    for (int64_t k = 0; k < nblk; ++k) {
      const double r = static_cast<double>(n - k * nb);
      const double bulk = r * r * static_cast<double>(nb) / (W * 33e12);
      const double chain = 0.6e-3 + r * 5.9e-8 * (static_cast<double>(nb) / 512.);
      if (bulk < chain) {
        tail_from = k;
        break;
      }
    }

    auto first_owned_after = [&](int64_t k) { return k + 1 + ((me - (k + 1)) % W + W) % W; };
    // C stream: broadcast of panel k into buffer k % NBUF
    auto bcast_p = [&](int64_t k) -> int {
      const MatView Pk = panel(k);
      const int root = static_cast<int>(k % W);
      const int64_t prev = k - NBUF; // the panel this buffer held before
      if (root == me) {
        AB_CUDA(cudaStreamWaitEvent(CS, ev.panel_end[k], 0)); // packed (and with it: buffer was free)
      } else if (prev >= 0) {
        AB_CUDA(cudaStreamWaitEvent(CS, bulkdone[prev], 0));
        AB_CUDA(cudaStreamWaitEvent(CS, pdone[prev], 0));
      }
      if (W > 1) {
        // from the panel's own first row (k * nb) to the end of its last column
        double *first = Pk.p + (k * nb - origin(k));
        const size_t count = static_cast<size_t>((width(k) - 1) * Pk.ld + (n - k * nb));
        AB_NCCL(g_nccl.Broadcast(first, first, count, ncclDouble, root,
                                 k >= tail_from && tail_comm != nullptr ? tail_comm : comm_of(h), CS));
      }
      AB_CUDA(cudaEventRecord(arrived[k], CS));
      return AB_OK;
    };
    // PS: factor + pack block column k (all its updates are on PS before this point)
    auto factor_and_pack_p = [&](int64_t k) -> int {
      const int64_t prev = k - NBUF;
      if (prev >= 0) { // the pack overwrites buffer k % NBUF: its previous panel must be consumed
        AB_CUDA(cudaStreamWaitEvent(PS, bulkdone[prev], 0)); // (PS's own use of it is stream-ordered)
      }
      return factor_and_pack(k);
    };
    if (nblk > 0) {
      AB_CUDA(cudaEventRecord(h->ev_ready, S)); // PS starts behind the Gram build on S
      AB_CUDA(cudaStreamWaitEvent(PS, h->ev_ready, 0));
      if (me == 0) {
        AB_TRY(factor_and_pack_p(0));
      }
      AB_TRY(bcast_p(0));
    }
    // AB_DIST_PCOL=0 (experiment): PS takes the next owned column only in the step that completes it
    // (jstar == k + 1); in the other steps it stays in the bulk launch
    bool pcol_always = true;
    if (const char *e = std::getenv("AB_DIST_PCOL")) {
      pcol_always = e[0] != '0';
    }
    for (int64_t k = 0; k < nblk; ++k) {
      int64_t jstar = first_owned_after(k);
      const bool p_takes = jstar < nblk && (pcol_always || jstar == k + 1);
      // ---- PS: the next column I own
      AB_CUDA(cudaStreamWaitEvent(PS, arrived[k], 0));
      if (p_takes) {
        // S updated this column with the panels before my previous column (jstar - W); it must be done
        // with them before PS takes the column over
        const int64_t handover = pcol_always ? jstar - W - 1 : k - 1;
        if ((!pcol_always || k == std::max<int64_t>(jstar - W, 0)) && handover >= 0) {
          // paired schedule: the head of a pair reaches this column in a launch of its own (below)
          const bool head = paired(handover) && (handover & 1) == 0;
          AB_CUDA(cudaStreamWaitEvent(PS, head ? ev.coldone[handover] : bulkdone[handover], 0));
        }
        {
          StreamSwap swap(h, PS);
          AB_TRY(update_cols(jstar, 1, k));
        }
      }
      AB_CUDA(cudaEventRecord(pdone[k], PS));
      if (jstar == k + 1 && jstar < nblk) {
        AB_TRY(factor_and_pack_p(jstar));
      }
      // ---- S: everything I own beyond jstar
      AB_CUDA(cudaEventRecord(ev.wait_begin[k], S));
      AB_CUDA(cudaStreamWaitEvent(S, arrived[k], 0));
      AB_CUDA(cudaEventRecord(ev.wait_end[k], S));
      if (!paired(k)) {
        AB_TRY(update_cols(p_takes ? jstar + W : jstar, -1, k));
        AB_CUDA(cudaEventRecord(bulkdone[k], S));
      } else if ((k & 1) == 0) {
        // head of a pair: wait for its partner — except for the one column that PS takes over with the
        // partner already (I own k + 1, so PS applies panels k + 1 .. to my NEXT column k + 1 + W)
        if (jstar == k + 1) {
          AB_TRY(update_cols(jstar + W, 1, k));
        }
        AB_CUDA(cudaEventRecord(ev.coldone[k], S));
      } else {
        // tail: panels k - 1 and k to every owned column beyond the one PS holds, depth 2 nb
        AB_TRY(update_cols(jstar + W, -1, k - 1, 2));
        AB_CUDA(cudaEventRecord(bulkdone[k - 1], S));
        AB_CUDA(cudaEventRecord(bulkdone[k], S));
      }
      // ---- CS: the next panel
      if (k + 1 < nblk) {
        AB_TRY(bcast_p(k + 1));
      }
    }
    // S continues (the solves read every block column) only after PS is done
    AB_CUDA(cudaEventRecord(h->ev_ready, PS));
    AB_CUDA(cudaStreamWaitEvent(S, h->ev_ready, 0));
  } else {
  if (nblk > 0) {
      if (me == 0) {
        // PS starts behind the Gram build on S
        AB_CUDA(cudaEventRecord(h->ev_ready, S));
        AB_CUDA(cudaStreamWaitEvent(PS, h->ev_ready, 0));
        AB_TRY(factor_and_pack(0));
      }
      AB_TRY(bcast(0));
    }
    for (int64_t k = 0; k < nblk; ++k) {
      AB_CUDA(cudaEventRecord(ev.wait_begin[k], S));
      AB_CUDA(cudaStreamWaitEvent(S, h->ev_bcast[k % 2], 0)); // panel k has arrived
      AB_CUDA(cudaEventRecord(ev.wait_end[k], S));
      const int64_t next = k + 1;
      if (next < nblk) {
        if (next % W == me) {
          AB_TRY(update(next, k));
          // PS: behind this update (and with it behind every earlier reader of pb[next % 2])
          AB_CUDA(cudaEventRecord(h->ev_ready, S));
          AB_CUDA(cudaStreamWaitEvent(PS, h->ev_ready, 0));
          AB_TRY(factor_and_pack(next));
        }
        AB_TRY(bcast(next));
      }
      // my first block column after k that is still to be updated (the look-ahead column already is)
      int64_t jfirst = k + 1 + ((me - (k + 1)) % W + W) % W;
      if (jfirst == next && next % W == me) {
        jfirst += W;
      }
      AB_TRY(update_from(jfirst, k));
    }
    // S continues (the solves read every block column) only after the last panel chain
    if (nblk > 0 && (nblk - 1) % W == me) {
      AB_CUDA(cudaStreamWaitEvent(S, ev.panel_end[nblk - 1], 0));
    }
}
  AB_CUDA(cudaEventRecord(ev.factor_end, S));
  phase_end(h, PH_FACTOR);

  // ---- information = K^-1 y by block substitution ----------------------------------------------
  // Left-looking in both directions, so that every step is one short-and-fat matrix-vector product that ALL
  // ranks perform concurrently on their own block columns (the row-block slice of L they own), one tiny
  // collective and one 1024 x 1024 block solve on the owner:
  //   forward   z_j = L_jj^-1 (y_j - sum_r p_r),  p_r = L[block j rows, rank r's columns < j] z[those columns]
  //   backward  x_j = L_jj^-T (z_j - t_j),  then every rank: t[own columns < j] += L[block j rows, .]^T x_j
  // (round 1 was right-looking: the owner of block j streamed its whole panel below the diagonal inside
  // step j, so the 69 GB of L at N = 131 072 were read one GPU at a time: 180 ms on 2, 4 and 8 GPUs alike.)
  phase_begin(h, PH_SOLVE);
  const size_t nbytes = static_cast<size_t>(round_up(n, 2)) * sizeof(double);
  const int64_t nloc1 = std::max<int64_t>(nloc, 1);
  const size_t lbytes = static_cast<size_t>(nloc1 * nb) * sizeof(double);
  void *d_z = nullptr, *d_zl = nullptr, *d_tl = nullptr, *d_p = nullptr, *d_t = nullptr, *d_logs = nullptr;
  AB_TRY(sc.alloc(nbytes, &d_z));
  AB_TRY(sc.alloc(lbytes, &d_zl));
  AB_TRY(sc.alloc(lbytes, &d_tl));
  AB_TRY(sc.alloc(static_cast<size_t>(nb) * sizeof(double), &d_p));
  AB_TRY(sc.alloc(static_cast<size_t>(nb) * sizeof(double), &d_t));
  AB_TRY(sc.alloc(static_cast<size_t>(nblk) * sizeof(double), &d_logs));
  AB_CUDA(cudaMemsetAsync(d_z, 0, nbytes, h->stream));
  AB_CUDA(cudaMemsetAsync(d_zl, 0, lbytes, h->stream));
  AB_CUDA(cudaMemsetAsync(d_tl, 0, lbytes, h->stream));
  AB_CUDA(cudaMemsetAsync(d_logs, 0, static_cast<size_t>(nblk) * sizeof(double), h->stream));
  double *z = static_cast<double *>(d_z), *zl = static_cast<double *>(d_zl);
  double *tl = static_cast<double *>(d_tl), *pbuf = static_cast<double *>(d_p);
  double *t = static_cast<double *>(d_t);
  // number of this rank's block columns with global index < j
  auto owned_before = [&](int64_t j) { return j > me ? (j - me + W - 1) / W : int64_t(0); };
  for (int64_t j = 0; j < nblk; ++j) {
    const int64_t r0 = j * nb, wj = width(j);
    const int root = static_cast<int>(j % W);
    const int64_t kcols = owned_before(j) * nb;
    if (kcols > 0) {
      AB_TRY(gemv_n(h, wj, kcols, 1., A.sub(r0, 0), zl, 0., pbuf));
    } else {
      AB_CUDA(cudaMemsetAsync(pbuf, 0, static_cast<size_t>(wj) * sizeof(double), h->stream));
    }
    if (W > 1) {
      AB_NCCL(g_nccl.Reduce(pbuf, t, static_cast<size_t>(wj), ncclDouble, ncclSum, root, comm_of(h),
                            h->stream));
    }
    if (root == me) {
      double *zj = zl + (j / W) * nb;
      sub_kernel<<<static_cast<unsigned>((wj + 255) / 256), 256, 0, h->stream>>>(
          d_y + r0, W > 1 ? t : pbuf, wj, zj);
      AB_LAUNCHED(h);
      const MatView D = colblk(j).sub(r0, 0);
      AB_TRY(trsv_lower(h, D, dinv_of(j), wj, zj)); // 512-row sub-blocks: one CTA solving 1024 rows costs ~1 ms
      AB_CUDA(cudaMemcpyAsync(z + r0, zj, static_cast<size_t>(wj) * sizeof(double),
                              cudaMemcpyDeviceToDevice, h->stream));
      AB_TRY(logdet_chol(h, D, wj, static_cast<double *>(d_logs) + j));
    }
  }
  // [0] = |z|^2 (own blocks; others are zero), [1] = sum of own log-determinants
  sum_sq_and_logdiag_kernel<<<1, 1024, 0, h->stream>>>(z, n, static_cast<double *>(d_logs), nblk,
                                                       h->d_scalars + 8);
  AB_LAUNCHED(h);
  AB_TRY(dist_allreduce_sum(h, h->d_scalars + 8, 2));
  double *x = d_info;
  for (int64_t j = nblk - 1; j >= 0; --j) {
    const int64_t r0 = j * nb, wj = width(j);
    const int root = static_cast<int>(j % W);
    if (root == me) {
      const int64_t l = j / W;
      sub_kernel<<<static_cast<unsigned>((wj + 255) / 256), 256, 0, h->stream>>>(
          zl + l * nb, tl + l * nb, wj, x + r0);
      AB_LAUNCHED(h);
      AB_TRY(trsv_lower_T(h, colblk(j).sub(r0, 0), dinv_of(j), wj, x + r0));
    }
    if (W > 1) {
      AB_NCCL(g_nccl.Broadcast(x + r0, x + r0, static_cast<size_t>(wj), ncclDouble, root, comm_of(h),
                               h->stream));
    }
    const int64_t kcols = owned_before(j) * nb;
    if (kcols > 0) {
      AB_TRY(gemv_t(h, wj, kcols, 1., A.sub(r0, 0), x + r0, 1., tl));
    }
  }
  phase_end(h, PH_SOLVE);

  // ---- scalars ----------------------------------------------------------------------------------
  std::vector<int> bad(static_cast<size_t>(nblk));
  AB_TRY(download_bytes(h, d_bad, bad.size() * sizeof(int), bad.data()));
  AB_TRY(download_bytes(h, h->d_scalars + 8, 2 * sizeof(double), h->h_scalars + 8));
  int64_t first_bad = -1;
  for (int64_t k = 0; k < nblk && first_bad < 0; ++k) {
    if (bad[static_cast<size_t>(k)] != INT_MAX) {
      first_bad = k * nb + bad[static_cast<size_t>(k)];
    }
  }
  // every rank must agree on failure: the smallest bad pivot over ranks (max of negated values)
  fac->bad_pivot = first_bad;
  if (W > 1) {
    h->h_scalars[34] = first_bad < 0 ? 0. : 1.;
    AB_CUDA(cudaMemcpyAsync(h->d_scalars + 34, h->h_scalars + 34, sizeof(double),
                            cudaMemcpyHostToDevice, h->stream));
    AB_TRY(dist_allreduce_sum(h, h->d_scalars + 34, 1));
    AB_TRY(download_bytes(h, h->d_scalars + 34, sizeof(double), h->h_scalars + 34));
    if (h->h_scalars[34] > 0. && first_bad < 0) {
      fac->bad_pivot = n; // another rank met a non-positive pivot
    }
  }
  if (fac->bad_pivot >= 0) {
    set_error("distributed factorisation: matrix is not positive definite");
    return AB_ERR_NOT_PD;
  }
  if (nll_out != nullptr) {
    const double zz = h->h_scalars[8], log_det = h->h_scalars[9];
    *nll_out = 0.5 * (log_det + zz + static_cast<double>(n) * std::log(2 * M_PI));
  }
  return AB_OK;
}

} // namespace
} // namespace ab

using namespace ab;

extern "C" {

int ab_dist_unique_id(void *id_out) {
  AB_REQUIRE(id_out != nullptr, "null");
  AB_TRY(load_nccl());
  static_assert(sizeof(ncclUniqueId) == AB_DIST_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  AB_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return AB_OK;
}

int ab_dist_init(ab_handle h, int rank, int world, const void *id) {
  AB_REQUIRE(h != nullptr && world >= 1 && rank >= 0 && rank < world, "rank / world");
  Lock lock(h);
  AB_REQUIRE(h->comm == nullptr, "distributed group already initialised");
  h->rank = rank;
  h->world = world;
  if (world == 1) {
    return AB_OK;
  }
  AB_REQUIRE(id != nullptr, "null id");
  AB_TRY(load_nccl());
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  // The panel broadcasts run WHILE the DMMA updates fill the machine: every NCCL CTA takes an SM's worth of
  // shared memory from them.  Eight CTAs carry a 0.5-1 GB panel fast enough (N = 131 072 on 8 GPUs, same box:
  // factorisation 3125 ms with NCCL's default, 3045 ms with 8 channels, 3094 ms with 4 where the chain starts to
  // wait; profiles/r02o_*), so the communicator is capped at 8 (AB_DIST_NCCL_MAX_CTAS overrides; 0 = NCCL default).
  int max_ctas = 8;
  if (const char *e = std::getenv("AB_DIST_NCCL_MAX_CTAS")) {
    max_ctas = std::atoi(e);
  }
  if (g_nccl.CommInitRankConfig != nullptr && max_ctas > 0) {
    ncclConfig_t config = NCCL_CONFIG_INITIALIZER;
    config.maxCTAs = max_ctas;
    AB_NCCL(g_nccl.CommInitRankConfig(&comm, world, uid, rank, &config));
  } else {
    AB_NCCL(g_nccl.CommInitRank(&comm, world, uid, rank));
  }
  h->comm = comm;
  if (g_nccl.CommSplit != nullptr && g_nccl.CommInitRankConfig != nullptr && max_ctas > 0) {
    ncclConfig_t config = NCCL_CONFIG_INITIALIZER;
    config.maxCTAs = 32;
    ncclComm_t tail = nullptr;
    AB_NCCL(g_nccl.CommSplit(comm, 0, rank, &tail, &config));
    std::lock_guard<std::mutex> lock2(g_tail_comm_mu);
    g_tail_comm[h] = tail;
  }
  return ensure_dist_streams(h);
}

int ab_dist_finalize(ab_handle h) {
  AB_REQUIRE(h != nullptr, "null handle");
  Lock lock(h);
  if (h->comm != nullptr) {
    cudaStreamSynchronize(h->stream);
    if (h->comm_stream != nullptr) {
      cudaStreamSynchronize(h->comm_stream);
    }
    if (ncclComm_t tail = tail_comm_of(h)) {
      g_nccl.CommDestroy(tail);
      std::lock_guard<std::mutex> lock2(g_tail_comm_mu);
      g_tail_comm.erase(h);
    }
    g_nccl.CommDestroy(comm_of(h));
    h->comm = nullptr;
  }
  if (h->comm_stream != nullptr) { // also created by a world-1 ab_dist_gp_fit
    cudaStreamSynchronize(h->comm_stream);
    cudaStreamDestroy(h->comm_stream);
    h->comm_stream = nullptr;
    cudaEventDestroy(h->ev_bcast[0]);
    cudaEventDestroy(h->ev_bcast[1]);
    cudaEventDestroy(h->ev_ready);
    cudaEventDestroy(h->ev_free);
    h->ev_bcast[0] = h->ev_bcast[1] = h->ev_ready = h->ev_free = nullptr;
  }
  h->rank = 0;
  h->world = 1;
  return AB_OK;
}

int ab_dist_info(ab_handle h, int *rank, int *world) {
  AB_REQUIRE(h != nullptr, "null handle");
  if (rank != nullptr) {
    *rank = h->rank;
  }
  if (world != nullptr) {
    *world = h->world;
  }
  return AB_OK;
}

int ab_dist_block_owner(int64_t block, int world, int *rank, int64_t *local_block) {
  AB_REQUIRE(block >= 0 && world >= 1, "block / world");
  if (rank != nullptr) {
    *rank = static_cast<int>(block % world);
  }
  if (local_block != nullptr) {
    *local_block = block / world;
  }
  return AB_OK;
}

int ab_dist_gp_fit(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n, int dim,
                   const double *y, const double *yvar, int64_t nb, ab_dist_factor *factor,
                   double *information, double *nll) {
  AB_REQUIRE(h != nullptr && factor != nullptr && n >= 1 && feats != nullptr && y != nullptr, "null");
  if (nb <= 0) {
    // The panel chain (apply panel -> potrf -> TRSM -> pack -> broadcast) is sequential across the ranks and
    // its cost per step grows with nb^2 while a rank's share of the trailing update shrinks with 1/world:
    // narrower block columns from 4 ranks on (N = 131 072 on 8 GPUs: nb 1024 / 768 / 640 / 512 = 3165 / 3138 /
    // 3127 / 3126 ms, waiting for a panel 143 -> 66 ms; profiles/r02m_*), 1024 below (deeper DMMA k-loop)
    nb = h->world >= 4 ? 512 : 1024;
    if (const char *e = std::getenv("AB_DIST_NB")) { // experiment hook (tools/bench_configs_dist.py --schedules)
      nb = std::atoll(e);
    }
  }
  AB_REQUIRE(nb % 128 == 0, "block size must be a multiple of 128");
  Lock lock(h);
  AB_REQUIRE(h->world == 1 || h->comm != nullptr, "distributed group not initialised");
  DevProg P;
  AB_TRY(compile_program(prog, nops, &P));
  Scope sc(h);
  timings_reset(h);
  *factor = nullptr;
  ab_matrix_s *F = nullptr;
  phase_begin(h, PH_H2D);
  AB_TRY(upload_features(h, feats, n, dim, &F));
  sc.own(F);
  void *d_y = nullptr, *d_yvar = nullptr, *d_info = nullptr;
  const size_t nbytes = static_cast<size_t>(n) * sizeof(double);
  AB_TRY(upload_bytes(h, sc, y, nbytes, &d_y));
  if (yvar != nullptr) {
    AB_TRY(upload_bytes(h, sc, yvar, nbytes, &d_yvar));
  }
  phase_end(h, PH_H2D);
  AB_TRY(sc.alloc(static_cast<size_t>(round_up(n, 2)) * sizeof(double), &d_info));
  AB_CUDA(cudaMemsetAsync(d_info, 0, static_cast<size_t>(round_up(n, 2)) * sizeof(double),
                          h->stream));
  auto *fac = new ab_dist_factor_s();
  int s = dist_fit_impl(h, P, F, static_cast<double *>(d_y), static_cast<double *>(d_yvar), n, nb,
                        fac, static_cast<double *>(d_info), nll);
  cudaEventRecord(h->ev_total_end, h->stream);
  if (s != AB_OK) {
    cudaStreamSynchronize(h->stream);
    if (h->comm_stream != nullptr) {
      cudaStreamSynchronize(h->comm_stream);
    }
    free_dist_factor(h, fac);
    return s;
  }
  if (information != nullptr) {
    s = download_bytes(h, d_info, nbytes, information);
  } else {
    cudaStreamSynchronize(h->stream);
  }
  *factor = fac;
  return s;
}

int ab_dist_fit_breakdown(ab_handle h, double *wait_ms, double *panel_ms, int64_t *steps) {
  AB_REQUIRE(h != nullptr, "null handle");
  Lock lock(h);
  AB_CUDA(cudaStreamSynchronize(h->stream));
  DistEvents &ev = dist_events_existing(h);
  double w = 0., p = 0.;
  for (int64_t k = 0; k < ev.steps; ++k) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.wait_begin[k], ev.wait_end[k]) == cudaSuccess) {
      w += ms;
    }
    if (ev.owned[static_cast<size_t>(k)] &&
        cudaEventElapsedTime(&ms, ev.panel_begin[k], ev.panel_end[k]) == cudaSuccess) {
      p += ms;
    }
  }
  cudaGetLastError(); // events of a fit that never ran report an error: not ours to keep
  if (wait_ms != nullptr) {
    *wait_ms = w;
  }
  if (panel_ms != nullptr) {
    *panel_ms = p;
  }
  if (steps != nullptr) {
    *steps = ev.steps;
  }
  return AB_OK;
}

int ab_dist_factor_free(ab_handle h, ab_dist_factor f) {
  AB_REQUIRE(h != nullptr, "null handle");
  Lock lock(h);
  free_dist_factor(h, f);
  return AB_OK;
}

int ab_dist_gram_rows(ab_handle h, const ab_op *prog, int nops, const double *feats, int64_t n,
                      int dim, int64_t *row0, int64_t *rows, ab_matrix *out) {
  AB_REQUIRE(h != nullptr && out != nullptr && n >= 0 && (n == 0 || feats != nullptr), "null");
  Lock lock(h);
  DevProg P;
  AB_TRY(compile_program(prog, nops, &P));
  Scope sc(h);
  timings_reset(h);
  // balanced row blocks, multiples of the 64-row Gram tile
  const int64_t per = round_up((n + h->world - 1) / h->world, 64);
  const int64_t r0 = std::min<int64_t>(n, per * h->rank);
  const int64_t nr = std::min<int64_t>(per, n - r0);
  ab_matrix_s *F = nullptr, *K = nullptr;
  AB_TRY(upload_features(h, feats, n, dim, &F));
  sc.own(F);
  AB_TRY(matrix_new(h, nr, n, &K));
  phase_begin(h, PH_GRAM);
  int s = nr > 0 ? gram_into(h, P, dim, false, F->d + r0 * F->ld, F->ld, nr, F->d, F->ld, n, K->d,
                             K->ld, 0u)
                 : AB_OK;
  phase_end(h, PH_GRAM);
  cudaEventRecord(h->ev_total_end, h->stream);
  cudaStreamSynchronize(h->stream);
  if (s != AB_OK) {
    matrix_delete(h, K);
    return s;
  }
  if (row0 != nullptr) {
    *row0 = r0;
  }
  if (rows != nullptr) {
    *rows = nr;
  }
  *out = K;
  return AB_OK;
}

int ab_dist_factor_broadcast(ab_handle h, ab_factor *factor, int root) {
  AB_REQUIRE(h != nullptr && factor != nullptr, "null");
  Lock lock(h);
  AB_REQUIRE(root >= 0 && root < h->world, "root rank");
  if (h->world <= 1) {
    return AB_OK;
  }
  AB_REQUIRE(h->comm != nullptr, "distributed group not initialised");
  const bool is_root = h->rank == root;
  AB_REQUIRE(is_root ? *factor != nullptr : *factor == nullptr,
             "the root passes its factor, every other rank a null handle");
  // header: n and the first bad pivot (a factor that is not usable is replicated as such)
  h->h_scalars[40] = is_root ? static_cast<double>((*factor)->n) : 0.;
  h->h_scalars[41] = is_root ? static_cast<double>((*factor)->bad_pivot) : 0.;
  AB_CUDA(cudaMemcpyAsync(h->d_scalars + 40, h->h_scalars + 40, 2 * sizeof(double),
                          cudaMemcpyHostToDevice, h->stream));
  AB_NCCL(g_nccl.Broadcast(h->d_scalars + 40, h->d_scalars + 40, 2, ncclDouble, root, comm_of(h),
                           h->stream));
  AB_TRY(download_bytes(h, h->d_scalars + 40, 2 * sizeof(double), h->h_scalars + 40));
  const int64_t n = static_cast<int64_t>(h->h_scalars[40]);
  ab_factor_s *f = is_root ? *factor : nullptr;
  if (!is_root) {
    ab_matrix_s *m = nullptr;
    AB_TRY(matrix_new(h, n, n, &m));
    int s = new_factor(h, m, &f);
    if (s != AB_OK) {
      matrix_delete(h, m);
      return s;
    }
    f->bad_pivot = static_cast<int64_t>(h->h_scalars[41]);
  }
  // one broadcast of the factor matrix (identical leading dimension on every rank: padded_ld(n)) and one
  // of the explicit leaf inverses
  timings_reset(h);
  phase_begin(h, PH_H2D);
  int status = AB_OK;
  if (n > 0) {
    // nothing else runs during this transfer: the uncapped communicator when there is one (the 8-CTA cap of
    // the main one costs 8.6 GB of L 69 ms instead of 19 ms on 8 GPUs, profiles/r02t_bench_8gpu_n131072.json)
    ncclComm_t wide = tail_comm_of(h) != nullptr ? tail_comm_of(h) : comm_of(h);
    if (g_nccl.Broadcast(f->m->d, f->m->d, static_cast<size_t>(f->m->ld) * static_cast<size_t>(n),
                         ncclDouble, root, wide, h->stream) != ncclSuccess ||
        g_nccl.Broadcast(f->dinv, f->dinv, f->dinv_bytes / sizeof(double), ncclDouble, root, wide,
                         h->stream) != ncclSuccess) {
      set_error("ncclBroadcast of the factor failed");
      status = AB_ERR_NCCL;
    }
  }
  phase_end(h, PH_H2D);
  cudaEventRecord(h->ev_total_end, h->stream);
  if (cudaStreamSynchronize(h->stream) != cudaSuccess && status == AB_OK) {
    set_error("factor broadcast failed");
    status = AB_ERR_CUDA;
  }
  if (status != AB_OK) {
    if (!is_root) {
      delete_factor(h, f);
    }
    return status;
  }
  *factor = f;
  return AB_OK;
}

int ab_dist_gp_cv(ab_handle h, ab_factor factor, const double *y, const double *information,
                  const int64_t *indices, const int64_t *offsets, int64_t ngroups, int what,
                  double *mean, double *var, double *score) {
  AB_REQUIRE(h != nullptr && y != nullptr && information != nullptr && indices != nullptr &&
                 offsets != nullptr && mean != nullptr && ngroups >= 0,
             "null");
  AB_REQUIRE(what == AB_PREDICT_MEAN || (what == AB_PREDICT_MARGINAL && var != nullptr),
             "ab_dist_gp_cv returns means or marginals");
  Lock lock(h);
  AB_TRY(require_usable(factor));
  timings_reset(h);
  const int64_t n = factor->n;
  double local_score = 0.;
  // a rank-local failure (a held-out block that is not positive definite, an allocation) must not skip
  // the assembling all-reduce: the other ranks would wait in NCCL for ever.  The status travels with
  // the payload and every rank returns the agreed result after the same sequence of collectives.
  const int local_status =
      gp_cv_impl(h, factor, y, information, indices, offsets, ngroups, what, h->rank, h->world, mean,
                 var, nullptr, score != nullptr ? &local_score : nullptr, nullptr);
  if (h->world <= 1 || n == 0) {
    if (local_status != AB_OK) {
      return local_status;
    }
  } else {
    // assemble: every observation was written by exactly one rank, the others hold zero
    Scope sc(h);
    void *d = nullptr;
    const int64_t cnt = 2 * n + 3; // means, variances, score, #ranks not PD, #ranks failed otherwise
    AB_TRY(sc.alloc(static_cast<size_t>(cnt) * sizeof(double), &d));
    double *buf = static_cast<double *>(d);
    std::vector<double> host(static_cast<size_t>(cnt), 0.);
    if (local_status == AB_OK) {
      std::copy(mean, mean + n, host.begin());
      if (what == AB_PREDICT_MARGINAL) {
        std::copy(var, var + n, host.begin() + n);
      }
      host[static_cast<size_t>(2 * n)] = local_score;
    } else {
      host[static_cast<size_t>(2 * n + (local_status == AB_ERR_NOT_PD ? 1 : 2))] = 1.;
    }
    AB_CUDA(cudaMemcpyAsync(buf, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice,
                            h->stream));
    AB_TRY(dist_allreduce_sum(h, buf, cnt));
    AB_TRY(download_bytes(h, buf, host.size() * sizeof(double), host.data()));
    if (local_status != AB_OK) {
      return local_status; // ab_last_error() already describes it
    }
    if (host[static_cast<size_t>(2 * n + 1)] > 0.) {
      set_error("a held-out block of the inverse covariance is not positive definite on another rank");
      return AB_ERR_NOT_PD;
    }
    if (host[static_cast<size_t>(2 * n + 2)] > 0.) {
      set_error("ab_dist_gp_cv failed on another rank");
      return AB_ERR_CUDA;
    }
    std::copy(host.begin(), host.begin() + n, mean);
    if (what == AB_PREDICT_MARGINAL) {
      std::copy(host.begin() + n, host.begin() + 2 * n, var);
    }
    local_score = host[static_cast<size_t>(2 * n)];
  }
  if (score != nullptr) {
    *score = local_score;
  }
  return AB_OK;
}

} // extern "C"
