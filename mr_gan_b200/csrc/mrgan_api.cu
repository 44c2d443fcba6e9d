// mrgan_api.cu -- handle, memory layout, launch sequences, CUDA-graph epoch and the C-ABI
// declared in include/mrgan.h.  Reference interface replaced: the K.function callables of
// mr_gan.py:169-171, the epoch loop mr_gan.py:183-230 and mr_nn.py:114-118.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

#include "../../include/mrgan.h"
#include "kernels_simt.cuh"
#ifdef MRGAN_WITH_TC
#include "kernels_tc.cuh"
#endif
#include "kernels_dp.cuh"

namespace {

thread_local std::string g_last_error;

// ------------------------------------------------------------------ op table
enum Op {
  OP_G1, OP_G2, OP_G3D, OP_G3G,
  OP_D1, OP_D2, OP_D3, OP_D4, OP_D5, OP_D6,
  OP_DW1, OP_DW2, OP_DW3, OP_DW4, OP_DW5, OP_DW6,
  OP_DX2, OP_DX3, OP_DX4, OP_DX5, OP_DX6,     // OP_DXl: dZ[l] -> dZ[l-1]
  OP_DX1G,                                     // dZ[1] -> dFake (G step only)
  OP_D1G, OP_D2G, OP_D3G, OP_D4G, OP_D5G,      // G step: D forward over 2B rows [fake | real] (own TMA boxes / MMA-N)
  OP_DX2G, OP_DX3G, OP_DX4G, OP_DX5G,          // G step: dX chain over the B fake rows
  OP_GW1, OP_GW2, OP_GW3, OP_GX2, OP_GX3,
  OP_E1, OP_E1S, OP_E2, OP_E3, OP_E4, OP_E5, OP_E6,
  NUM_OPS
};

struct OpInfo { bool at = false, bt = false, used = false; int maxM = 0, maxN = 0; };

struct Arena {
  char* base = nullptr; size_t off = 0;
  template <typename T> T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

struct TensorLayout { int rows, cols, pitch; long long off; };   // inside a net's flat block

struct NetLayout {
  std::vector<TensorLayout> t;   // D: 6 augmented matrices; G: W1aug, gamma, beta, W2aug, W3aug
  long long n = 0;               // padded length (multiple of 32)
  long long n_ref = 0;           // length in the reference's packed order
  long long off = 0;             // offset of this net in the handle's flat buffers
};

struct FoldBuffers {
  float *a[6], *hb[6], *dz[6], *lg, *dlg, *dfake, *zb, *h1g, *xhat, *istd, *u, *h2g, *dz2g, *du, *dz1g;
  float *stage_x, *stage_z, *ex_stage, *xte, *eh[6], *elg, *xtr, *eval_out, *eval_out_s;
  int *stage_y, *labels_cur, *ey_stage, *ytr, *yte, *idx;
  int *lab_rows, *unl_rows;      // device-side epoch permutations: labeled / unlabeled row subsets (mrgan_set_epoch_rows)
  int n_lab = 0, n_unl = 0;
  int lda[6], ldz[6];   // pitches of a/h (with ones column) and dz
  bool loaded = false;
};

}  // namespace

struct mrgan_handle;
#ifdef MRGAN_WITH_TC
namespace {
int tc_setup(mrgan_handle* h);
void tc_teardown(mrgan_handle* h);
bool tc_launch_gemm(mrgan_handle* h, int op, int f0, int nfl, int rows_override, cudaStream_t st);
void tc_launch_dw_adam(mrgan_handle* h, int op, int nlayers, int f0, int nfl, cudaStream_t st);
bool tc_dw_merge(const mrgan_handle* h);
void tc_params_changed(mrgan_handle* h, int fold, int net);
int tc_debug_gemm(mrgan_handle* h, int mode, const GemmDesc& g, int esz);
}
#endif

constexpr int kMaxChains = 16;
struct mrgan_handle {
  mrgan_config cfg;
  int nf = 0, R = 0, NE = 0, n_train = 0;
  std::vector<mrgan_fold_shape> shapes;
  std::vector<NetLayout> net[2];
  std::vector<FoldBuffers> fb;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;        // dW (+Adam) kernels run here, concurrently with the latency-bound dX chain
  cudaEvent_t ev_pool[16] = {nullptr}; int ev_next = 0;
  bool use_pdl = false;               // programmatic dependent launch between the kernels of a step: opt-in (MRGAN_PDL=1);
                                      // measured neutral at 74 folds/GPU and -17 % at 12 (early-resident dependents hold SM resources)
  bool side_open = false;             // work forked to `side` since the last join (joins are no-ops otherwise)
  // the folds of a group are independent, so an epoch is captured as `nchains` parallel chains of kernels (disjoint fold
  // ranges, own main + side stream): one chain's latency-bound small kernels fill the SMs another chain leaves idle
  int nchains = 1;
  cudaStream_t cmain[kMaxChains] = {nullptr}, cside[kMaxChains] = {nullptr};
  char* arena = nullptr; size_t arena_bytes = 0;
  __half* harena = nullptr;           // f16 mode: one __half per float of the arena (operand copies at the same element index)
  OperandMode om = {0, 1.0f, nullptr, nullptr};
  float *P = nullptr, *Mo = nullptr, *Vo = nullptr, *Gr = nullptr; long long n_flat = 0;
  FoldState* d_folds = nullptr; std::vector<FoldState> h_folds;
  GemmDesc* d_descs = nullptr; std::vector<GemmDesc> h_descs;
  PermDesc* d_perm = nullptr;
  BnDesc* d_bn = nullptr; LossDesc* d_loss = nullptr; EvalDesc *d_eval = nullptr, *d_eval_s = nullptr;
  AdamRange* d_ranges[2] = {nullptr, nullptr};
  OpInfo ops[NUM_OPS];
  float *d_step_stats = nullptr, *d_epoch_stats = nullptr, *h_epoch_stats = nullptr;
  int* h_idx_pinned[2] = {nullptr, nullptr}; cudaEvent_t idx_free[2] = {nullptr, nullptr}; int idx_slot = 0;
  float* h_scratch = nullptr;   // pinned scalars for the step API
  std::map<int, cudaGraphExec_t> graphs;   // key: nb (GAN) or n_idx (NN)
  std::map<int, long long> graph_nodes;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool epoch_pending = false; double last_ms = 0.0;
  long long launches = 0;
  std::string err;
  int sticky = 0;                     // first failure of a call that cannot return a status itself (kernel launch, NCCL enqueue):
                                      // surfaced by the next public entry point through take_sticky()
  AdamHyper hp;
  // data-parallel mode (mrgan_dp_init): NCCL communicator + buffers of the all-reduced batch statistics
  // device-resident raw datasets (mrgan_load_dataset) and scratch of the device-side fold preparation
  struct Dataset { float* x = nullptr; int* y = nullptr; int n = 0, D = 0, ld = 0; } datasets[8];
  double* d_prep_stats = nullptr; int* d_prep_rows = nullptr; int prep_rows_cap = 0, prep_stats_cap = 0;
  int dp_world = 1, dp_rank = 0;
  bool dp_virtual = false;            // the handle's folds play the ranks (mrgan_dp_init_virtual): collectives are local sums
  // fused gradient exchange (kernels_dp.cuh): peer mappings of every rank's arena (cudaIpc), flag blocks, per-net pointer sets
  bool dp_fused = false;
  unsigned* d_dpflags = nullptr;      // DP_MAX_RANKS x DP_FLAG_WORDS words inside the arena (zero at creation)
  char* peer_arena[DP_MAX_RANKS] = {nullptr}; __half* peer_harena[DP_MAX_RANKS] = {nullptr};
  DpPeers dp_peers[2];                // [net]
  void* nccl_comm = nullptr;
  float* d_dpmem = nullptr; DpBufs* d_dpbufs = nullptr;
  float *dp_bnf = nullptr, *dp_bnb = nullptr, *dp_fm = nullptr;   // [nf][2*500], [nf][2*500], [nf][2*250]
  float* dp_losspart = nullptr; unsigned* dp_lossctr = nullptr;    // k_loss_disc block partials [nf][4*64] and arrival counters [nf]
#ifdef MRGAN_WITH_TC
  TcOp* d_tcops = nullptr;            // [NUM_OPS][nf] tensor maps + epilogue descriptors
  int tc_bn[NUM_OPS] = {0}, tc_maxME[NUM_OPS] = {0}, tc_maxNE[NUM_OPS] = {0};
  int tc_ksplit[NUM_OPS] = {0};       // large-batch dW: contraction slices per tile (deterministic split-K), 0/1 = off
  bool tc_mt1[NUM_OPS] = {false};     // large-batch forward / dX whose 256 x 256 tiling would leave SMs idle: 128-feature CTAs (tc_pick_tile)
  float* d_tcws = nullptr;            // ... and their partial-product workspace
  bool tc_fused_adam = true;          // dW epilogue applies Adam in place (no gradient round trip)
  bool tc_mt2 = true;                 // forward / dX: 256 features per CTA where the layer is wide enough (MRGAN_MT2=0 disables)
  bool tc_adam_tma = true;            // ... with W/m/v staged through smem by TMA (k_dw_adam_tc) instead of the LSU
  bool tc_dw_merge = true;            // dW+Adam of the narrow layers (D 3..6; G 1, 2) in one launch each; MRGAN_DW_MERGE=0: one launch per layer
  bool tc_dw_small = true;            // dW+Adam: 16 KB operand stages (32 fp16 / 16 fp32 batch rows) -> three CTAs per SM; MRGAN_DW_SMALL=0: two
  int tc_heads = 0;                   // losses / feature matching / BatchNorm fused into GEMM epilogues (set by tc_setup), bit mask:
                                      // 1 = HEAD_DISC (no k_loss_disc, no k_adam of D), 2 = HEAD_FM (no k_fm), 4 = HEAD_BN (no k_bn_fwd),
                                      // 8 = HEAD_BN_BWD (no k_bn_bwd); 2 and 8 together also retire the generator's k_adam
  TcAdamOp* d_tcadam = nullptr;       // [NUM_OPS][nf]
  AdamRange* d_ranges_tc[2] = {nullptr, nullptr};   // what is left for k_adam: BN gamma/beta (G), nothing (D)
#endif
};

namespace {

int take_sticky(mrgan_handle* h);

int fail(mrgan_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_last_error = msg;
  return code;
}

#define CKS()                                                                                      \
  do {                                                                                             \
    CK(cudaGetLastError());                                                                        \
    const int _s = take_sticky(h);                                                                 \
    if (_s) return _s;                                                                             \
  } while (0)

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(h, MRGAN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
  } while (0)

const int kDW[5] = {1000, 500, 250, 250, 250};   // mr_gan.py:119-127
const int kGH = 500;                               // mr_gan.py:111,113

long long align32(long long x) { return (x + 31) & ~31LL; }

NetLayout layout_disc(int D, int K) {
  NetLayout L;
  int dims[7] = {D, kDW[0], kDW[1], kDW[2], kDW[3], kDW[4], K};
  for (int l = 0; l < 6; ++l) {
    TensorLayout t{dims[l] + 1, dims[l + 1], pitch8(dims[l + 1]), L.n};
    L.t.push_back(t);
    L.n = align32(L.n + (long long)t.rows * t.pitch);
    L.n_ref += (long long)(dims[l] + 1) * dims[l + 1];
  }
  return L;
}

NetLayout layout_gen(int D, int nd) {
  NetLayout L;
  auto add = [&](int rows, int cols) {
    TensorLayout t{rows, cols, pitch8(cols), L.n};
    L.t.push_back(t);
    L.n = align32(L.n + (long long)rows * t.pitch);
    L.n_ref += (long long)rows * cols;
  };
  add(nd + 1, kGH);   // W1aug
  add(1, kGH);        // gamma
  add(1, kGH);        // beta
  add(kGH + 1, kGH);  // W2aug
  add(kGH + 1, D);    // W3aug
  return L;
}

__global__ void k_round_tf32(float* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = rna_tf32(p[i]);
}

__global__ void k_set_col(float* p, int ld, int rows, int col, float v, OperandMode om) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) {
    p[(size_t)r * ld + col] = v;
    if (om.mode == 2) om.hbase[p + (size_t)r * ld + col - om.fbase] = __float2half_rn(v);
  }
}

__global__ void k_scale_buf(float* p, size_t n, float mul) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] *= mul;
}

// f16 mode, test hook: operand-only buffers exist as fp16 copies only; expand one (times `mul`) into a float scratch
__global__ void k_from_half(const float* src, float* dst, size_t n, float mul, OperandMode om) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __half2float(om.hbase[src + i - om.fbase]) * mul;
}

// ---- buffer layout (run twice: sizing pass with base == nullptr, then for real) ----
void layout_buffers(mrgan_handle* h, Arena& ar) {
  const mrgan_config& c = h->cfg;
  const int B = c.batch, R = h->R, NE = h->NE, K = c.n_classes, nd = c.noise_dim, nf = h->nf;
  const bool gan = c.model == MRGAN_MODEL_GAN;
  h->P = ar.take<float>(h->n_flat);
  h->Mo = ar.take<float>(h->n_flat);
  h->Vo = ar.take<float>(h->n_flat);
  h->Gr = ar.take<float>(h->n_flat);
  h->d_folds = ar.take<FoldState>(nf);
  h->d_descs = ar.take<GemmDesc>((size_t)NUM_OPS * nf);
  h->d_perm = ar.take<PermDesc>(nf);
  h->d_bn = ar.take<BnDesc>(nf);
  h->d_loss = ar.take<LossDesc>(nf);
  h->d_eval = ar.take<EvalDesc>(nf);
  h->d_eval_s = ar.take<EvalDesc>(nf);
  h->d_ranges[0] = ar.take<AdamRange>(nf);
  h->d_ranges[1] = ar.take<AdamRange>(nf);
  h->d_step_stats = ar.take<float>((size_t)(h->n_train / B + 1) * nf * 4);
  h->d_dpflags = ar.take<unsigned>((size_t)DP_MAX_RANKS * DP_FLAG_WORDS);
  h->d_epoch_stats = ar.take<float>((size_t)nf * 8);
  for (int f = 0; f < nf; ++f) {
    FoldBuffers& b = h->fb[f];
    const int D = h->shapes[f].D, ntr = h->shapes[f].n_train, nte = h->shapes[f].n_test;
    const int ldx = pitch8(D);
    int win[6] = {D, kDW[0], kDW[1], kDW[2], kDW[3], kDW[4]};
    for (int l = 0; l < 6; ++l) { b.lda[l] = pitch8(win[l] + 1); b.ldz[l] = pitch8(win[l]); }
    b.a[0] = ar.take<float>((size_t)R * b.lda[0]);
    b.hb[0] = nullptr; b.dz[0] = nullptr;
    for (int l = 1; l <= 5; ++l) {
      b.hb[l] = ar.take<float>((size_t)R * b.lda[l]);
      b.a[l] = (l < 5) ? ar.take<float>((size_t)R * b.lda[l]) : b.hb[l];
      b.dz[l] = ar.take<float>((size_t)R * b.ldz[l]);
    }
    b.lg = ar.take<float>((size_t)R * pitch8(K));
    b.dlg = ar.take<float>((size_t)R * pitch8(K));
    b.labels_cur = ar.take<int>(R);
    b.stage_x = ar.take<float>((size_t)R * ldx);
    b.stage_y = ar.take<int>(R);
    if (gan) {
      b.dfake = ar.take<float>((size_t)B * ldx);
      b.zb = ar.take<float>((size_t)B * pitch8(nd + 1));
      const int ldg = pitch8(kGH);
      b.h1g = ar.take<float>((size_t)B * ldg);
      b.xhat = ar.take<float>((size_t)B * ldg);
      b.istd = ar.take<float>(kGH);
      b.u = ar.take<float>((size_t)B * pitch8(kGH + 1));
      b.h2g = ar.take<float>((size_t)B * pitch8(kGH + 1));
      b.dz2g = ar.take<float>((size_t)B * ldg);
      b.du = ar.take<float>((size_t)B * ldg);
      b.dz1g = ar.take<float>((size_t)B * ldg);
      b.stage_z = ar.take<float>((size_t)B * nd);
    }
    b.ex_stage = ar.take<float>((size_t)NE * b.lda[0]);
    b.xte = ar.take<float>((size_t)nte * b.lda[0]);
    for (int l = 1; l <= 5; ++l) b.eh[l] = ar.take<float>((size_t)NE * b.lda[l]);
    b.elg = ar.take<float>((size_t)NE * pitch8(K));
    b.ey_stage = ar.take<int>(NE);
    b.eval_out = ar.take<float>(4);
    b.eval_out_s = ar.take<float>(4);
    b.xtr = ar.take<float>((size_t)ntr * ldx);
    b.ytr = ar.take<int>(ntr);
    b.yte = ar.take<int>(nte);
    b.idx = ar.take<int>((size_t)3 * ntr);
    b.lab_rows = ar.take<int>(ntr);
    b.unl_rows = ar.take<int>(ntr);
  }
}

GemmDesc make_desc(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N, int K, int epi,
                   int act, int fold) {
  GemmDesc d;
  memset(&d, 0, sizeof(d));
  d.A = A; d.lda = lda; d.B = Bm; d.ldb = ldb; d.C = C; d.ldc = ldc; d.M = M; d.N = N; d.K = K;
  d.epi = epi; d.act = act; d.fold = fold;
  return d;
}

int build_descs(mrgan_handle* h) {
  const mrgan_config& c = h->cfg;
  const int B = c.batch, R = h->R, K = c.n_classes, nd = c.noise_dim, nf = h->nf;
  const bool gan = c.model == MRGAN_MODEL_GAN;
  h->h_descs.assign((size_t)NUM_OPS * nf, GemmDesc{});
  h->h_folds.assign(nf, FoldState{});
  std::vector<BnDesc> bn(nf); std::vector<LossDesc> ls(nf); std::vector<EvalDesc> ev(nf), evs(nf);
  std::vector<AdamRange> rg0(nf), rg1(nf);
  const bool tensor = c.precision != MRGAN_PREC_FP32;      // tf32 / f16: which outputs feed a tensor-core GEMM as operands
  auto setop = [&](int op, int f, GemmDesc d, bool at, bool bt, int rnd = 0) {
    d.rnd = tensor ? rnd : 0;
    h->h_descs[(size_t)op * nf + f] = d;
    OpInfo& oi = h->ops[op];
    oi.at = at; oi.bt = bt; oi.used = true;
    if (d.M > oi.maxM) oi.maxM = d.M;
    if (d.N > oi.maxN) oi.maxN = d.N;
  };
  for (int f = 0; f < nf; ++f) {
    FoldBuffers& b = h->fb[f];
    const int D = h->shapes[f].D;
    const NetLayout& LD = h->net[0][f];
    float* PD = h->P + LD.off; float* GD = h->Gr + LD.off;
    int win[6] = {D, kDW[0], kDW[1], kDW[2], kDW[3], kDW[4]};
    int wout[6] = {kDW[0], kDW[1], kDW[2], kDW[3], kDW[4], K};
    const float sig[5] = {c.sigma_in, c.sigma_hidden, c.sigma_hidden, c.sigma_hidden, c.sigma_hidden};
    const int hact = c.hidden_act == MRGAN_ACT_LEAKY_RELU ? ACT_LEAKY : ACT_RELU;      // mr_gan.py:119-127 / wganlpctsemi.py:169
    const bool dropout = c.dropout > 0.0f;
    // ---- discriminator forward (train): layer l (1-based) reads a[l-1], writes h[l] (+ noisy a[l])
    for (int l = 1; l <= 6; ++l) {
      const TensorLayout& W = LD.t[l - 1];
      float* C = (l <= 5) ? b.hb[l] : b.lg;
      const int ldc = (l <= 5) ? b.lda[l] : pitch8(K);
      GemmDesc d = make_desc(b.a[l - 1], b.lda[l - 1], PD + W.off, W.pitch, C, ldc, R, wout[l - 1], win[l - 1] + 1,
                             EPI_FWD, l <= 5 ? hact : ACT_NONE, f);
      if (l <= 4) { d.C2 = b.a[l]; d.ldc2 = b.lda[l]; d.sigma = sig[l]; d.tid = l; d.row0 = 0; }
      setop(OP_D1 + l - 1, f, d, false, false, (l == 5 ? 1 : 0) | 2);
      if (gan && l <= 5) { GemmDesc dg = d; dg.M = 2 * B; setop(OP_D1G + l - 1, f, dg, false, false, (l == 5 ? 1 : 0) | 2); }
      // eval twin: no noise, reads the clean activations
      const float* EA = (l == 1) ? b.xte : b.eh[l - 1];
      float* EC = (l <= 5) ? b.eh[l] : b.elg;
      GemmDesc e = make_desc(EA, b.lda[l - 1], PD + W.off, W.pitch, EC, ldc, h->shapes[f].n_test, wout[l - 1],
                             win[l - 1] + 1, EPI_FWD, l <= 5 ? hact : ACT_NONE, f);
      setop(l == 1 ? OP_E1 : OP_E2 + l - 2, f, e, false, false, l <= 5 ? 1 : 0);
      if (l == 1) { e.A = b.ex_stage; setop(OP_E1S, f, e, false, false, 1); }
      // dW (+db): grad(Waug_l) = a[l-1]^T @ dZ[l]
      const float* dZ = (l <= 5) ? b.dz[l] : b.dlg;
      const int lddz = (l <= 5) ? b.ldz[l] : pitch8(K);
      GemmDesc w = make_desc(b.a[l - 1], b.lda[l - 1], dZ, lddz, GD + W.off, W.pitch, win[l - 1] + 1, wout[l - 1], R,
                             EPI_STORE, ACT_NONE, f);
      setop(OP_DW1 + l - 1, f, w, true, false);
      // dX: dZ[l-1] = (dZ[l] @ W_l^T) * relu'(h[l-1])
      if (l >= 2) {
        GemmDesc x = make_desc(dZ, lddz, PD + W.off, W.pitch, b.dz[l - 1], b.ldz[l - 1], R, win[l - 1], wout[l - 1],
                               EPI_DX, hact, f);
        // act'(h[l-1]) needs the layer's clean output -- or, when Dropout follows the activation, the dropped output a[l-1],
        // which carries the keep mask and the sign of h at once (dx_rows); tid >= 1 marks "behind a hidden-layer transform"
        x.aux = (dropout && l - 1 <= 4) ? b.a[l - 1] : b.hb[l - 1]; x.ldaux = b.lda[l - 1]; x.tid = (l - 1 <= 4) ? l - 1 : 0;
        setop(OP_DX2 + l - 2, f, x, false, true, 1);
        if (gan && l <= 5) { GemmDesc xg = x; xg.M = B; setop(OP_DX2G + l - 2, f, xg, false, true, 1); }
      }
    }
    rg0[f] = AdamRange{LD.off, LD.n};
    ls[f] = LossDesc{b.lg, b.dlg, pitch8(K), b.labels_cur, b.hb[5], b.dz[5], b.lda[5], b.ldz[5], kDW[4]};
    ev[f] = EvalDesc{b.elg, pitch8(K), b.yte, h->shapes[f].n_test, (h->shapes[f].n_test / B) * B, b.eval_out};
    evs[f] = EvalDesc{b.elg, pitch8(K), b.ey_stage, h->shapes[f].n_test, 0, b.eval_out_s};
    if (gan) {
      const NetLayout& LG = h->net[1][f];
      float* PG = h->P + LG.off; float* GG = h->Gr + LG.off;
      const TensorLayout &W1 = LG.t[0], &Tg = LG.t[1], &Tb = LG.t[2], &W2 = LG.t[3], &W3 = LG.t[4];
      const int ldzb = pitch8(nd + 1), ldu = pitch8(kGH + 1), ldx = pitch8(D), ldg = pitch8(kGH);
      setop(OP_G1, f, make_desc(b.zb, ldzb, PG + W1.off, W1.pitch, b.h1g, ldg, B, kGH, nd + 1, EPI_FWD, ACT_SOFTPLUS, f), false, false);
      setop(OP_G2, f, make_desc(b.u, ldu, PG + W2.off, W2.pitch, b.h2g, ldu, B, kGH, kGH + 1, EPI_FWD, ACT_SOFTPLUS, f), false, false, 1);
      GemmDesc g3 = make_desc(b.h2g, ldu, PG + W3.off, W3.pitch, nullptr, 0, B, D, kGH + 1, EPI_FWD, ACT_NONE, f);
      g3.C2 = b.a[0] + (size_t)2 * B * b.lda[0]; g3.ldc2 = b.lda[0]; g3.sigma = c.sigma_in; g3.tid = 0; g3.row0 = 2 * B;
      setop(OP_G3D, f, g3, false, false, 2);
      g3.C2 = b.a[0]; g3.row0 = 0;
      setop(OP_G3G, f, g3, false, false, 2);
      // dFake = dZ[1] @ W1^T  (no activation derivative: fake is G's linear output)
      const TensorLayout& DW1 = LD.t[0];
      setop(OP_DX1G, f, make_desc(b.dz[1], b.ldz[1], PD + DW1.off, DW1.pitch, b.dfake, ldx, B, D, kDW[0], EPI_DX, ACT_NONE, f), false, true, 1);
      setop(OP_GW3, f, make_desc(b.h2g, ldu, b.dfake, ldx, GG + W3.off, W3.pitch, kGH + 1, D, B, EPI_STORE, ACT_NONE, f), true, false);
      GemmDesc gx3 = make_desc(b.dfake, ldx, PG + W3.off, W3.pitch, b.dz2g, ldg, B, kGH, D, EPI_DX, ACT_SOFTPLUS, f);
      gx3.aux = b.h2g; gx3.ldaux = ldu;
      setop(OP_GX3, f, gx3, false, true, 1);
      setop(OP_GW2, f, make_desc(b.u, ldu, b.dz2g, ldg, GG + W2.off, W2.pitch, kGH + 1, kGH, B, EPI_STORE, ACT_NONE, f), true, false);
      setop(OP_GX2, f, make_desc(b.dz2g, ldg, PG + W2.off, W2.pitch, b.du, ldg, B, kGH, kGH, EPI_DX, ACT_NONE, f), false, true);
      setop(OP_GW1, f, make_desc(b.zb, ldzb, b.dz1g, ldg, GG + W1.off, W1.pitch, nd + 1, kGH, B, EPI_STORE, ACT_NONE, f), true, false);
      bn[f] = BnDesc{b.h1g, b.xhat, b.u, b.istd, PG + Tg.off, PG + Tb.off, b.du, b.dz1g, GG + Tg.off, GG + Tb.off,
                     ldg, ldu, B, kGH};
      rg1[f] = AdamRange{LG.off, LG.n};
    }
    FoldState& fs = h->h_folds[f];
    fs.key0 = (uint32_t)(h->shapes[f].seed & 0xFFFFFFFFull);
    fs.key1 = (uint32_t)(h->shapes[f].seed >> 32);
    fs.D = D; fs.n_train = h->shapes[f].n_train; fs.n_test = h->shapes[f].n_test; fs.ldx = pitch8(D);
    fs.x_train = b.xtr; fs.y_train = b.ytr; fs.y_test = b.yte;
    for (int s = 0; s < 3; ++s) fs.idx[s] = b.idx + (size_t)s * h->shapes[f].n_train;
    fs.stage_x = b.stage_x; fs.stage_y = b.stage_y; fs.stage_z = gan ? b.stage_z : nullptr;
    fs.a0 = b.a[0]; fs.lda0 = b.lda[0]; fs.z = gan ? b.zb : nullptr; fs.ldz = pitch8(nd + 1);
    fs.labels_cur = b.labels_cur;
  }
  CK(cudaMemcpy(h->d_descs, h->h_descs.data(), h->h_descs.size() * sizeof(GemmDesc), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_folds, h->h_folds.data(), nf * sizeof(FoldState), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_loss, ls.data(), nf * sizeof(LossDesc), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_eval, ev.data(), nf * sizeof(EvalDesc), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_eval_s, evs.data(), nf * sizeof(EvalDesc), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_ranges[0], rg0.data(), nf * sizeof(AdamRange), cudaMemcpyHostToDevice));
  if (gan) {
    CK(cudaMemcpy(h->d_bn, bn.data(), nf * sizeof(BnDesc), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_ranges[1], rg1.data(), nf * sizeof(AdamRange), cudaMemcpyHostToDevice));
  }
  return MRGAN_OK;
}

void set_ones(mrgan_handle* h, float* p, int ld, int rows, int col) {
  k_set_col<<<(rows + 127) / 128, 128, 0, h->stream>>>(p, ld, rows, col, 1.0f, h->om);
}

void init_ones(mrgan_handle* h) {
  const mrgan_config& c = h->cfg;
  const bool gan = c.model == MRGAN_MODEL_GAN;
  for (int f = 0; f < h->nf; ++f) {
    FoldBuffers& b = h->fb[f];
    const int D = h->shapes[f].D;
    int win[6] = {D, kDW[0], kDW[1], kDW[2], kDW[3], kDW[4]};
    set_ones(h, b.a[0], b.lda[0], h->R, D);
    set_ones(h, b.ex_stage, b.lda[0], h->NE, D);
    set_ones(h, b.xte, b.lda[0], h->shapes[f].n_test, D);
    for (int l = 1; l <= 5; ++l) {
      set_ones(h, b.hb[l], b.lda[l], h->R, win[l]);
      if (l < 5) set_ones(h, b.a[l], b.lda[l], h->R, win[l]);
      set_ones(h, b.eh[l], b.lda[l], h->NE, win[l]);
    }
    if (gan) {
      set_ones(h, b.zb, pitch8(c.noise_dim + 1), c.batch, c.noise_dim);
      set_ones(h, b.u, pitch8(kGH + 1), c.batch, kGH);
      set_ones(h, b.h2g, pitch8(kGH + 1), c.batch, kGH);
    }
  }
}


// ------------------------------------------------------------------ kernel launch with programmatic stream serialization
template <typename... KArgs, typename... Args>
void launch_k(mrgan_handle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = h->use_pdl ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (e != cudaSuccess && !h->sticky) { h->sticky = MRGAN_ERR_CUDA; h->err = std::string("kernel launch: ") + cudaGetErrorString(e); }
  h->launches++;
}

// status of everything enqueued since the last check (launch_k / dp_allreduce record the first failure)
int take_sticky(mrgan_handle* h) {
  if (!h->sticky) return MRGAN_OK;
  const int code = h->sticky;
  h->sticky = 0;
  g_last_error = h->err;
  return code;
}

// ------------------------------------------------------------------ NCCL (resolved at run time: no link dependency)
struct NcclId { char b[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

bool nccl_load() {
  if (g_nccl.lib) return true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return false;
  g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) return false;
  g_nccl.lib = lib;
  return true;
}

// Virtual-rank stand-in for the collective: `buf` holds W equal segments (one per fold = rank); every segment becomes the
// element-wise sum over the segments, added in rank order (what ncclAllReduce(sum) leaves on every rank).
__global__ void k_virtual_allreduce(float* __restrict__ buf, size_t seg, int W) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < seg; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < W; ++w) s += buf[(size_t)w * seg + i];
    for (int w = 0; w < W; ++w) buf[(size_t)w * seg + i] = s;
  }
}

// in-stream sum all-reduce of fp32 (ncclFloat32 = 7, ncclSum = 0); no-op outside the data-parallel mode
void dp_allreduce(mrgan_handle* h, float* buf, size_t n) {
  if (h->dp_world <= 1 || n == 0) return;
  if (h->dp_virtual) {
    const size_t seg = n / (size_t)h->dp_world;
    const int blocks = (int)std::min<size_t>((seg + 255) / 256, 1184);
    k_virtual_allreduce<<<blocks, 256, 0, h->stream>>>(buf, seg, h->dp_world);
    h->launches++;
    return;
  }
  const int rc = g_nccl.AllReduce(buf, buf, n, 7, 0, h->nccl_comm, h->stream);
  if (rc != 0 && !h->sticky) {
    h->sticky = MRGAN_ERR_CUDA;
    h->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
  }
}

void dp_allreduce_grads(mrgan_handle* h, int net) {
  if (h->dp_world <= 1) return;
  long long n = 0;
  for (int f = 0; f < h->nf; ++f) n += h->net[net][f].n;
  dp_allreduce(h, h->Gr + h->net[net][0].off, (size_t)n);      // the nets of all folds are contiguous in the flat buffer
}

void launch_adam(mrgan_handle* h, int f0, int nfl, int net, bool counters_only);

// Pointer sets of the fused exchange.  Real ranks: the same offsets inside every rank's (IPC-mapped) arena.  Virtual ranks:
// fold r's net inside this handle's own flat buffers.
void dp_build_peers(mrgan_handle* h) {
  const float* abase = reinterpret_cast<const float*>(h->arena);
  for (int net = 0; net < 2; ++net) {
    DpPeers& pp = h->dp_peers[net];
    memset(&pp, 0, sizeof(pp));
    if (h->net[net].empty()) continue;
    for (int p = 0; p < h->dp_world; ++p) {
      if (h->dp_virtual) {
        const long long d = h->net[net][p].off - h->net[net][0].off;
        pp.P[p] = h->P + d; pp.Gr[p] = h->Gr + d; pp.Mo[p] = h->Mo + d; pp.Vo[p] = h->Vo + d;
        pp.Ph[p] = h->harena ? h->harena + (h->P - abase) + d : nullptr;
        pp.flags[p] = h->d_dpflags + (size_t)p * DP_FLAG_WORDS;
      } else {
        float* pb = reinterpret_cast<float*>(h->peer_arena[p]);
        pp.P[p] = pb + (h->P - abase); pp.Gr[p] = pb + (h->Gr - abase); pp.Mo[p] = pb + (h->Mo - abase); pp.Vo[p] = pb + (h->Vo - abase);
        pp.Ph[p] = h->peer_harena[p] ? h->peer_harena[p] + (h->P - abase) : nullptr;
        pp.flags[p] = reinterpret_cast<unsigned*>(h->peer_arena[p] + (reinterpret_cast<char*>(h->d_dpflags) - h->arena));
      }
    }
  }
}

// One kernel per rank: reduce-scatter of the net's flat gradient by peer loads, Adam on the owned shard, all-gather of the
// updated weights by peer stores, step counters (kernels_dp.cuh).  Replaces ncclAllReduce + the full-size k_adam.
void dp_exchange(mrgan_handle* h, int net) {
  long long len = 0;
  if (h->dp_virtual) len = h->net[net][0].n;
  else for (int f = 0; f < h->nf; ++f) len += h->net[net][f].n;
  const long long off = h->net[net][0].off;
  const int W = h->dp_world;
  const int bx = h->dp_virtual ? std::max(1, 296 / W) : 296;      // all CTAs co-resident (2 per SM): they wait on one another
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(bx, h->dp_virtual ? W : 1, 1); cfg.blockDim = dim3(256); cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const float ginv = h->om.mode == 2 ? 1.0f / h->om.gscale : 1.0f;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, k_dp_exchange, h->dp_peers[net], W, h->dp_virtual ? -1 : h->dp_rank, off, len, h->d_folds,
                                           h->nf, net, h->hp, ginv);
  if (e != cudaSuccess && !h->sticky) { h->sticky = MRGAN_ERR_CUDA; h->err = std::string("k_dp_exchange launch: ") + cudaGetErrorString(e); }
  h->launches++;
}

// gradient exchange + optimizer of one net in the data-parallel mode
void dp_update(mrgan_handle* h, int f0, int nfl, int net) {
  if (h->dp_fused) { dp_exchange(h, net); return; }
  dp_allreduce_grads(h, net);
  launch_adam(h, f0, nfl, net, false);
}

// Scratch of the split (statistics -> [all-reduce] -> apply) BatchNorm / feature-matching kernels: used by the data-parallel
// mode and, on one GPU, by the large-batch regime where one CTA per fold would serialise thousands of rows.
int alloc_split_bufs(mrgan_handle* h) {
  if (h->d_dpbufs) return MRGAN_OK;
  const int nf = h->nf;
  // per fold: the three statistics blocks, the row-chunk partials of the statistics kernels, the block partials of
  // k_loss_disc, and the arrival counters (16 column blocks + 1 for the loss kernel, as 32 words)
  const size_t stats = 2 * kGH + 2 * kGH + 2 * kDW[4];
  const size_t per = stats + (size_t)SPLIT_MAX_Y * 2 * SPLIT_PART_W + 4 * LOSS_MAX_BLOCKS + 32;
  CK(cudaMalloc(&h->d_dpmem, per * nf * sizeof(float)));
  CK(cudaMemset(h->d_dpmem, 0, per * nf * sizeof(float)));
  h->dp_bnf = h->d_dpmem; h->dp_bnb = h->dp_bnf + (size_t)nf * 2 * kGH; h->dp_fm = h->dp_bnb + (size_t)nf * 2 * kGH;
  float* const part = h->dp_fm + (size_t)nf * 2 * kDW[4];
  h->dp_losspart = part + (size_t)nf * SPLIT_MAX_Y * 2 * SPLIT_PART_W;
  unsigned* const ctr = reinterpret_cast<unsigned*>(h->dp_losspart + (size_t)nf * 4 * LOSS_MAX_BLOCKS);
  h->dp_lossctr = ctr + (size_t)nf * 16;
  std::vector<DpBufs> bufs(nf);
  for (int f = 0; f < nf; ++f)
    bufs[f] = DpBufs{h->dp_bnf + (size_t)f * 2 * kGH, h->dp_bnb + (size_t)f * 2 * kGH, h->dp_fm + (size_t)f * 2 * kDW[4],
                     part + (size_t)f * SPLIT_MAX_Y * 2 * SPLIT_PART_W, ctr + (size_t)f * 16};
  CK(cudaMalloc(&h->d_dpbufs, nf * sizeof(DpBufs)));
  CK(cudaMemcpy(h->d_dpbufs, bufs.data(), nf * sizeof(DpBufs), cudaMemcpyHostToDevice));
  return MRGAN_OK;
}

// Row chunks (grid.y) of the split statistics / apply kernels: one chunk per 256 rows (8 rows per thread of a 1024-thread
// block), so that a large batch spreads over the chip instead of 8 - 16 CTAs.
int split_chunks(int rows) {
  const int y = (rows + 255) / 256;
  return y < 1 ? 1 : (y > SPLIT_MAX_Y ? SPLIT_MAX_Y : y);
}

// ------------------------------------------------------------------ launch sequences
void launch_gemm(mrgan_handle* h, int op, int f0, int nfl, int rows_override, cudaStream_t st = nullptr) {
  if (!st) st = h->stream;
  const OpInfo& oi = h->ops[op];
  int M = oi.maxM;
  if (rows_override > 0 && !oi.at) M = rows_override;
  const GemmDesc* d = h->d_descs + (size_t)op * h->nf + f0;
#ifdef MRGAN_WITH_TC
  if (h->cfg.precision != MRGAN_PREC_FP32 && tc_launch_gemm(h, op, f0, nfl, rows_override, st)) return;
#endif
  dim3 grid((oi.maxN + 63) / 64, (M + 63) / 64, nfl);
  if (!oi.at && !oi.bt) launch_k(h, k_gemm_simt<false, false>, grid, dim3(256), 0, st, d, (const FoldState*)h->d_folds, rows_override, h->hp);
  else if (!oi.at && oi.bt) launch_k(h, k_gemm_simt<false, true>, grid, dim3(256), 0, st, d, (const FoldState*)h->d_folds, rows_override, h->hp);
  else launch_k(h, k_gemm_simt<true, false>, grid, dim3(256), 0, st, d, (const FoldState*)h->d_folds, rows_override, h->hp);
}

// side stream: fork = "side waits for everything enqueued on main so far", join = the reverse.
// Works both eagerly and under stream capture (where it becomes graph edges).
void fork_side(mrgan_handle* h) {
  cudaEvent_t e = h->ev_pool[h->ev_next++ & 15];
  cudaEventRecord(e, h->stream);
  cudaStreamWaitEvent(h->side, e, 0);
  h->side_open = true;
}
void join_side(mrgan_handle* h) {
  if (!h->side_open) return;
  h->side_open = false;
  cudaEvent_t e = h->ev_pool[h->ev_next++ & 15];
  cudaEventRecord(e, h->side);
  cudaStreamWaitEvent(h->stream, e, 0);
}

int max_D(const mrgan_handle* h, int f0, int nfl) {
  int m = 0;
  for (int f = f0; f < f0 + nfl; ++f) m = h->shapes[f].D > m ? h->shapes[f].D : m;
  return m;
}

void launch_prep(mrgan_handle* h, int f0, int nfl, int mode, int from_stage, int t, int nrows) {
  const mrgan_config& c = h->cfg;
  int cols = max_D(h, f0, nfl);
  if (c.noise_dim > cols) cols = c.noise_dim;
  // rows per thread: 8 at the reference batch, 16 in the large-batch regime (kernels_simt.cuh: k_prep; 8 / 16 / 32 rows per
  // thread measured the same once the gathers were batched: 224 / 227 / 227 step pairs/s at D = 12032, B = 8192)
  const int pg = nrows > 512 ? 4 : 2;
  dim3 grid((cols + 127) / 128, (nrows + 4 * pg - 1) / (4 * pg), nfl);
#define LAUNCH_PREP(PG) launch_k(h, k_prep<PG>, grid, dim3(128), 0, h->stream, h->d_folds, f0, mode, from_stage, t, c.batch, nrows, c.noise_dim, \
                                 c.sigma_in, h->hp, h->om)
  if (pg == 4) LAUNCH_PREP(4); else LAUNCH_PREP(2);
#undef LAUNCH_PREP
}

void launch_adam(mrgan_handle* h, int f0, int nfl, int net, bool counters_only) {
  long long nmax = 0;
  for (int f = f0; f < f0 + nfl; ++f) nmax = h->net[net][f].n > nmax ? h->net[net][f].n : nmax;
  int blocks = (int)((nmax / 4 + 255) / 256);
  if (blocks > 2048) blocks = 2048;
  if (blocks < 1) blocks = 1;
  const AdamRange* ranges = h->d_ranges[net];
#ifdef MRGAN_WITH_TC
  if (h->cfg.precision != MRGAN_PREC_FP32 && h->tc_fused_adam) { ranges = h->d_ranges_tc[counters_only ? 0 : net]; blocks = 1; }
#endif
  // f16 mode: the gradient buffer carries the loss scale, and every parameter update refreshes the fp16 operand copy
  const float ginv = h->om.mode == 2 ? 1.0f / h->om.gscale : 1.0f;
  __half* Ph = h->om.mode == 2 ? h->harena + (h->P - reinterpret_cast<float*>(h->arena)) : nullptr;
  launch_k(h, k_adam, dim3(blocks, nfl), dim3(256), 0, h->stream, h->P, h->Mo, h->Vo, (const float*)h->Gr, ranges, h->d_folds, f0, net, h->hp,
           ginv, Ph);
}

// The side stream's dW kernels read the step's activation buffers (a[l], dZ[l]), which the NEXT step's batch assembly
// and forward pass overwrite: every step therefore joins the side stream before it ends (measured: deferring the join
// past the next step's generator forward gains nothing once several fold chains run concurrently, and would need
// double-buffered activations).
bool deferred_join(const mrgan_handle*) { return false; }

int heads_on(const mrgan_handle* h) {
#ifdef MRGAN_WITH_TC
  return h->cfg.precision != MRGAN_PREC_FP32 ? h->tc_heads : 0;
#else
  return 0;
#endif
}

void enqueue_gen_fwd(mrgan_handle* h, int f0, int nfl, int op_g3) {
  if (deferred_join(h)) join_side(h);          // generator weights of the previous G step (GW1..3 on the side stream)
  launch_gemm(h, OP_G1, f0, nfl, 0);
  if (heads_on(h) & 4) {                       // BatchNorm ran in G1's epilogue
    launch_gemm(h, OP_G2, f0, nfl, 0);
    launch_gemm(h, op_g3, f0, nfl, 0);
    return;
  }
  const dim3 bnf((kGH + BN_COLS - 1) / BN_COLS, h->d_dpbufs ? split_chunks(h->cfg.batch) : 1, nfl);
  const dim3 bnt(h->cfg.batch > 256 ? 1024 : 256);     // 32 columns x 8 (reference batch) or 32 (large batch) row slices
  const OperandMode tf32 = h->om;
  if (h->d_dpbufs) {          // batch statistics over the GLOBAL batch: local sums -> NVLink all-reduce -> apply
    k_bn_stats<<<bnf, bnt, 0, h->stream>>>(h->d_bn + f0, h->d_dpbufs + f0);
    dp_allreduce(h, h->dp_bnf + (size_t)f0 * 2 * kGH, (size_t)nfl * 2 * kGH);
    k_bn_apply<<<bnf, bnt, 0, h->stream>>>(h->d_bn + f0, h->d_dpbufs + f0, h->cfg.bn_eps, tf32, h->hp.dp_bg);
    h->launches += 2;
  } else {
    launch_k(h, k_bn_fwd, bnf, bnt, 0, h->stream, (const BnDesc*)(h->d_bn + f0), h->cfg.bn_eps, tf32);
  }
  launch_gemm(h, OP_G2, f0, nfl, 0);
  launch_gemm(h, op_g3, f0, nfl, 0);
}

// train_batch_disc (mr_gan.py:169)
void enqueue_disc_step(mrgan_handle* h, int f0, int nfl, int t, int from_stage) {
  const mrgan_config& c = h->cfg;
  const int B = c.batch;
  const bool heads = (heads_on(h) & 1) != 0;
  h->hp.t = t;
  launch_prep(h, f0, nfl, 0, from_stage, t, 2 * B);
  enqueue_gen_fwd(h, f0, nfl, OP_G3D);
  for (int l = 0; l < 6; ++l) launch_gemm(h, OP_D1 + l, f0, nfl, 0);
  if (!heads) {      // otherwise the logit layer's epilogue computed the losses, their gradients and advanced the counters
    // large batch: the 3B rows spread over up to 64 blocks (needs the split scratch for the ordered partial sums)
    int lb = h->dp_losspart ? (3 * B + 511) / 512 : 1;
    lb = lb < 1 ? 1 : (lb > LOSS_MAX_BLOCKS ? LOSS_MAX_BLOCKS : lb);
    launch_k(h, k_loss_disc, dim3(lb, 1, nfl), dim3(256), 0, h->stream, (const LossDesc*)(h->d_loss + f0), h->d_step_stats, f0, h->nf, t, B,
             c.n_classes, c.unlabeled_weight, h->om, h->hp.dp_bg, h->dp_losspart, h->dp_lossctr);
  }
#ifdef MRGAN_WITH_TC
  const bool merge = tc_dw_merge(h);
#else
  const bool merge = false;
#endif
  for (int l = 6; l >= 1; --l) {      // dX first: it reads W_l, which the fused-Adam dW epilogue overwrites
    if (l >= 2) launch_gemm(h, OP_DX2 + l - 2, f0, nfl, 0);
    if (merge && l > 3) continue;     // the narrow layers 3..6 share ONE dW+Adam launch, after the last dX that reads their weights
    fork_side(h);                     // dW_l (+Adam) streams HBM on the side while main continues the dX chain
#ifdef MRGAN_WITH_TC
    if (merge && l == 3) { tc_launch_dw_adam(h, OP_DW3, 4, f0, nfl, h->side); continue; }
#endif
    launch_gemm(h, OP_DW1 + l - 1, f0, nfl, 0, h->side);
  }
  if (!deferred_join(h) || from_stage) join_side(h);   // deferred: dW1+Adam overlaps the next G step's generator forward
  if (h->dp_world > 1) dp_update(h, f0, nfl, 0);
  else if (!heads) launch_adam(h, f0, nfl, 0, false);
  if (from_stage) dp_allreduce(h, h->d_step_stats + ((size_t)t * h->nf + f0) * 4, (size_t)nfl * 4);
}

// train_batch_gen (mr_gan.py:170)
void enqueue_gen_step(mrgan_handle* h, int f0, int nfl, int t, int from_stage) {
  const mrgan_config& c = h->cfg;
  const int B = c.batch;
  const int hm = heads_on(h);
  h->hp.t = t;
  launch_prep(h, f0, nfl, 1, from_stage, t, 2 * B);
  enqueue_gen_fwd(h, f0, nfl, OP_G3G);
  if (deferred_join(h)) join_side(h);          // discriminator weights updated by the D step's dW+Adam kernels
  for (int l = 0; l < 5; ++l) launch_gemm(h, OP_D1G + l, f0, nfl, 0);
  if (h->d_dpbufs) {
    const dim3 fmg((kDW[4] + BN_COLS - 1) / BN_COLS, split_chunks(B), nfl), fmt(B > 256 ? 1024 : 256);
    k_fm_stats<<<fmg, fmt, 0, h->stream>>>(h->d_loss + f0, h->d_dpbufs + f0, B);
    dp_allreduce(h, h->dp_fm + (size_t)f0 * 2 * kDW[4], (size_t)nfl * 2 * kDW[4]);
    k_fm_apply<<<fmg, fmt, 0, h->stream>>>(h->d_loss + f0, h->d_dpbufs + f0, h->d_step_stats, f0, h->nf, t, B,
                                           h->om, h->hp.dp_bg, h->dp_world, h->hp.alpha);
    h->launches += 2;
  } else if (!(hm & 2)) {   // otherwise D layer 5's epilogue computed the feature-matching loss, its gradient, and advanced the counters
    launch_k(h, k_fm, dim3(1, 1, nfl), dim3(1024), 0, h->stream, (const LossDesc*)(h->d_loss + f0), h->d_step_stats, f0, h->nf, t, B,
             h->om, h->hp.alpha);
  }
  for (int l = 5; l >= 2; --l) launch_gemm(h, OP_DX2G + l - 2, f0, nfl, 0);
  launch_gemm(h, OP_DX1G, f0, nfl, 0);
  launch_gemm(h, OP_GX3, f0, nfl, 0);
  fork_side(h);
  launch_gemm(h, OP_GW3, f0, nfl, 0, h->side);
  launch_gemm(h, OP_GX2, f0, nfl, 0);
#ifdef MRGAN_WITH_TC
  const bool merge = tc_dw_merge(h);
#else
  const bool merge = false;
#endif
  if (!merge) {
    fork_side(h);
    launch_gemm(h, OP_GW2, f0, nfl, 0, h->side);
  }
  const dim3 bng((kGH + BN_COLS - 1) / BN_COLS, h->d_dpbufs ? split_chunks(B) : 1, nfl), bnt(B > 256 ? 1024 : 256);
  if (h->d_dpbufs) {
    k_bn_bwd_stats<<<bng, bnt, 0, h->stream>>>(h->d_bn + f0, h->d_dpbufs + f0, h->om);
    dp_allreduce(h, h->dp_bnb + (size_t)f0 * 2 * kGH, (size_t)nfl * 2 * kGH);
    k_bn_bwd_apply<<<bng, bnt, 0, h->stream>>>(h->d_bn + f0, h->d_dpbufs + f0, h->om, h->hp.dp_bg);
    h->launches += 2;
  } else if (!(hm & 8)) {   // otherwise GX2's epilogue did BatchNorm backward and the Adam update of gamma / beta
    launch_k(h, k_bn_bwd, bng, bnt, 0, h->stream, (const BnDesc*)(h->d_bn + f0), h->om);
  }
  fork_side(h);
#ifdef MRGAN_WITH_TC
  if (merge) tc_launch_dw_adam(h, OP_GW1, 2, f0, nfl, h->side);      // G layers 1 and 2 share one dW+Adam launch
  else
#endif
  launch_gemm(h, OP_GW1, f0, nfl, 0, h->side);
  if (!deferred_join(h) || from_stage) join_side(h);
  if (h->dp_world > 1) dp_update(h, f0, nfl, 1);
  else if ((hm & 10) != 10) launch_adam(h, f0, nfl, 1, (hm & 8) != 0);      // gamma / beta and the counters, unless heads took both over
  dp_allreduce(h, h->d_step_stats + ((size_t)t * h->nf + f0) * 4, (size_t)nfl * 4);
}

// one model.fit batch of mr_nn (mr_nn.py:117)
void enqueue_nn_step(mrgan_handle* h, int f0, int nfl, int t, int from_stage, int n) {
  const mrgan_config& c = h->cfg;
  // a ragged batch (n < batch) runs at full height: k_loss_mse zeroes the gradient rows >= n, so the
  // stale rows contribute nothing to dW/db and every GEMM keeps its static shape
  h->hp.t = t;
  launch_prep(h, f0, nfl, 2, from_stage, t, n);
  for (int l = 0; l < 6; ++l) launch_gemm(h, OP_D1 + l, f0, nfl, 0);
  launch_k(h, k_loss_mse, dim3(1, 1, nfl), dim3(256), 0, h->stream, (const LossDesc*)(h->d_loss + f0), h->d_step_stats, f0, h->nf, t, n, h->R,
           c.n_classes, h->om);
  for (int l = 6; l >= 1; --l) {
    if (l >= 2) launch_gemm(h, OP_DX2 + l - 2, f0, nfl, 0);
    fork_side(h);
    launch_gemm(h, OP_DW1 + l - 1, f0, nfl, 0, h->side);
  }
  join_side(h);
  launch_adam(h, f0, nfl, 0, false);
}

// test_batch (mr_gan.py:171): phase 0, no noise
void enqueue_eval(mrgan_handle* h, int f0, int nfl, bool staged, int n_override) {
  join_side(h);
  launch_gemm(h, staged ? OP_E1S : OP_E1, f0, nfl, n_override);
  for (int l = 1; l < 6; ++l) launch_gemm(h, OP_E2 + l - 1, f0, nfl, n_override);
  launch_k(h, k_argmax_err, dim3(1, 1, nfl), dim3(256), 0, h->stream, (const EvalDesc*)((staged ? h->d_eval_s : h->d_eval) + f0), n_override,
           h->cfg.n_classes);
}

int check_fold(mrgan_handle* h, int fold) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (fold < 0 || fold >= h->nf) return fail(h, MRGAN_ERR_ARG, "fold index out of range");
  return MRGAN_OK;
}

int finish_pending(mrgan_handle* h) {
  if (h->epoch_pending) {
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms;
    h->epoch_pending = false;
  }
  return MRGAN_OK;
}

// reference-order <-> internal (augmented, padded) parameter layout
void pack_params(const mrgan_handle* h, int fold, int net, const float* src, std::vector<float>& dst, bool to_internal,
                 float* out_ref) {
  const NetLayout& L = h->net[net][fold];
  long long o = 0;
  auto mat = [&](const TensorLayout& t, int rows_w) {     // W[rows_w, cols] then b[cols]
    for (int r = 0; r < rows_w + 1; ++r)
      for (int cidx = 0; cidx < t.cols; ++cidx) {
        const long long ii = t.off + (long long)r * t.pitch + cidx;
        if (to_internal) dst[ii] = src[o]; else out_ref[o] = dst[ii];
        ++o;
      }
  };
  auto vec = [&](const TensorLayout& t) {
    for (int cidx = 0; cidx < t.cols; ++cidx) {
      if (to_internal) dst[t.off + cidx] = src[o]; else out_ref[o] = dst[t.off + cidx];
      ++o;
    }
  };
  if (net == 0) {
    for (int l = 0; l < 6; ++l) mat(L.t[l], L.t[l].rows - 1);
  } else {
    mat(L.t[0], L.t[0].rows - 1); vec(L.t[1]); vec(L.t[2]); mat(L.t[3], L.t[3].rows - 1); mat(L.t[4], L.t[4].rows - 1);
  }
}

int build_graph(mrgan_handle* h, int key, int nb, int n_idx_nn) {
  if (h->graphs.count(key)) return MRGAN_OK;
  const mrgan_config& c = h->cfg;
  const long long before = h->launches;
  cudaGraph_t g = nullptr;
  CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  {
    cudaStream_t origin = h->stream, origin_side = h->side;
    const int nch = h->dp_world > 1 ? 1 : h->nchains;      // collectives: one chain, the same order on every rank
    cudaEvent_t e = h->ev_pool[h->ev_next++ & 15];
    cudaEventRecord(e, origin);
    for (int ch = 1; ch < nch; ++ch) cudaStreamWaitEvent(h->cmain[ch], e, 0);
    for (int ch = 0; ch < nch; ++ch) {
      const int f0 = (int)((long long)h->nf * ch / nch), f1 = (int)((long long)h->nf * (ch + 1) / nch);
      if (f1 <= f0) continue;
      h->stream = h->cmain[ch]; h->side = h->cside[ch];
      h->side_open = false;
      if (c.model == MRGAN_MODEL_GAN) {
        for (int t = 0; t < nb; ++t) {
          enqueue_disc_step(h, f0, f1 - f0, t, 0);
          enqueue_gen_step(h, f0, f1 - f0, t, 0);
        }
        if (c.eval_each_epoch) enqueue_eval(h, f0, f1 - f0, false, 0);
        join_side(h);                             // every forked stream must rejoin before the capture ends
      } else {
        for (int t = 0; t < nb; ++t) enqueue_nn_step(h, f0, f1 - f0, t, 0, c.batch);
        join_side(h);
      }
    }
    h->stream = origin; h->side = origin_side;
    for (int ch = 1; ch < nch; ++ch) {
      cudaEvent_t j = h->ev_pool[h->ev_next++ & 15];
      cudaEventRecord(j, h->cmain[ch]);
      cudaStreamWaitEvent(origin, j, 0);
    }
  }
  k_epoch_reduce<<<h->nf, 32, 0, h->stream>>>(h->d_step_stats, h->d_eval, h->d_epoch_stats, h->nf, nb,
                                              c.model == MRGAN_MODEL_GAN && c.eval_each_epoch);
  h->launches++;
  CK(cudaMemcpyAsync(h->h_epoch_stats, h->d_epoch_stats, (size_t)h->nf * 8 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamEndCapture(h->stream, &g));
  { const int sk = take_sticky(h); if (sk) { cudaGraphDestroy(g); return sk; } }
  cudaGraphExec_t ge = nullptr;
  CK(cudaGraphInstantiate(&ge, g, 0));
  CK(cudaGraphDestroy(g));
  h->graphs[key] = ge;
  h->graph_nodes[key] = h->launches - before;
  h->launches = before;
  (void)n_idx_nn;
  return MRGAN_OK;
}

int upload_indices(mrgan_handle* h, const int32_t* const* streams, int n_streams, int n_idx) {
  // host [nf][n_idx] per stream -> pinned staging -> device idx[f][s][n_train]
  const int slot = h->idx_slot;
  h->idx_slot ^= 1;
  CK(cudaEventSynchronize(h->idx_free[slot]));
  int* pin = h->h_idx_pinned[slot];
  for (int f = 0; f < h->nf; ++f)
    for (int s = 0; s < n_streams; ++s)
      memcpy(pin + ((size_t)f * 3 + s) * h->n_train, streams[s] + (size_t)f * n_idx, (size_t)n_idx * sizeof(int));
  for (int f = 0; f < h->nf; ++f)
    CK(cudaMemcpyAsync(h->fb[f].idx, pin + (size_t)f * 3 * h->n_train, (size_t)3 * h->n_train * sizeof(int),
                       cudaMemcpyHostToDevice, h->stream));
  CK(cudaEventRecord(h->idx_free[slot], h->stream));
  return MRGAN_OK;
}

}  // namespace


#ifdef MRGAN_WITH_TC
// ------------------------------------------------------------------ tcgen05 path: host side
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2D fp32 tensor [rows, cols] with `pitch` floats per row; box = 32 columns (one 128B swizzle row) x box_rows
// L2 promotion of the TMA loads: every box row is one 128-byte line of a row-major matrix, so neighbouring boxes /
// k-blocks touch the adjacent 128 bytes of the same DRAM page; fetching 256 B per miss halves the DRAM activations.
CUtensorMapL2promotion g_l2_promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;

// esz = 4: fp32 elements (read as tf32); esz = 2: fp16 elements (`base` then points at __half data, pitch in elements).
// The swizzled box is always one 128-byte row wide (32 fp32 / 64 fp16 elements); MN-major 32-bit operands need the
// 32-byte-atom variant of the 128B swizzle, 16-bit ones the plain 128B swizzle.
bool make_map(EncodeTiledFn fn, CUtensorMap* m, const void* base, int cols, int rows, int pitch, int box_rows,
              bool mn_major, int box_cols = 0, int esz = 4) {
  const int row_elems = 128 / esz;
  if (box_cols == 0) box_cols = row_elems;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * (cuuint64_t)esz};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1u, 1u};
  return fn(m, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
            strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            box_cols != row_elems ? CU_TENSOR_MAP_SWIZZLE_NONE
                                  : ((mn_major && esz == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
            g_l2_promo,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

// forward / dX: 4-stage ring, up to 256 accumulator columns, 1 CTA per SM
// dW (+ fused Adam): 2-stage ring (the contraction is only ~150 rows = 5 blocks), 128 columns, 2 CTAs per SM
#define TC_FWD_STAGES 4
#define TC_DW_STAGES 2
#define TC_FWD_THREADS (64 + 32 * 8)
#define TC_DW_THREADS (64 + 32 * 4)
// kernel instantiations, indexed by the operand format F (false: fp32 operands as tf32, true: fp16 operand copies)
#define K_TC_FWD(F, V) k_gemm_tc<true, false, TC_FWD_STAGES, 256, 1, 8, 1, F, V>
#define K_TC_FWD_N(F, V) k_gemm_tc<true, false, TC_FWD_STAGES, 512, 1, 8, 1, F, V>   // all 512 TMEM columns: room to park the epilogue's noise
#define K_TC_DX(F, V) k_gemm_tc<false, false, TC_FWD_STAGES, 256, 1, 8, 1, F, V>
#define K_TC_FWD2(F, V) k_gemm_tc<true, false, TC_FWD_STAGES, 512, 1, 8, 2, F, V>     // 256 features per CTA (layers >= 500 wide)
#define K_TC_DX_N(F, V) k_gemm_tc<false, false, TC_FWD_STAGES, 512, 1, 8, 1, F, V>
#define K_TC_DX2(F, V) k_gemm_tc<false, false, TC_FWD_STAGES, 512, 1, 8, 2, F, V>
#define K_TC_DW(F, V) k_gemm_tc<true, true, TC_DW_STAGES, 128, 2, 4, 1, F, V>
// large-batch regime (more than 256 batch rows: data-parallel config 5): 256 x 256 tiles, 3-stage ring of 64 KB stages
#define TC_BIG_STAGES 3
#define K_TC_FWD_BIG(F, V) k_gemm_tc<true, false, TC_BIG_STAGES, 512, 1, 8, 2, F, V>
#define K_TC_DX_BIG(F, V) k_gemm_tc<false, false, TC_BIG_STAGES, 512, 1, 8, 2, F, V>
#define K_TC_DW_BIG(F, V) k_gemm_tc<true, true, TC_BIG_STAGES, 512, 1, 8, 2, F, V>
// launches K(operand format, variant) with identical arguments: f16 = fp16 operand copies, var = the handle uses the
// LeakyReLU / Dropout discriminator variants (their epilogue code lives in separate instantiations)
#define TC_LAUNCH(h, f16, K, grid, block, smem, st, ...)                          \
  do {                                                                            \
    const bool var_ = (h)->cfg.hidden_act != MRGAN_ACT_RELU || (h)->cfg.dropout > 0.0f;                \
    if (f16 && var_) launch_k(h, K(true, true), grid, block, smem, st, __VA_ARGS__);                   \
    else if (f16) launch_k(h, K(true, false), grid, block, smem, st, __VA_ARGS__);                     \
    else if (var_) launch_k(h, K(false, true), grid, block, smem, st, __VA_ARGS__);                    \
    else launch_k(h, K(false, false), grid, block, smem, st, __VA_ARGS__);                             \
  } while (0)
size_t tc_smem_bytes(int bn, int stages, int mt = 1) { return 1024 + (size_t)stages * ((size_t)mt * 128 * 128 + (size_t)bn * 128) + 256; }
// two feature sub-tiles per CTA when the layer is wide enough and the 4-stage ring still fits in 227 KB
bool tc_use_mt2(int maxME, int bn) { return maxME >= 500 && tc_smem_bytes(bn, TC_FWD_STAGES, 2) <= 227 * 1024; }

int tc_setup(mrgan_handle* h);

EncodeTiledFn tc_encoder() {
  EncodeTiledFn fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) return nullptr;
  return fn;
}

template <bool F, bool V> void tc_set_smem_attr_fmt() {
  cudaFuncSetAttribute(K_TC_FWD(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(256, TC_FWD_STAGES));
  cudaFuncSetAttribute(K_TC_FWD_N(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(256, TC_FWD_STAGES));
  cudaFuncSetAttribute(K_TC_DX(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(256, TC_FWD_STAGES));
  cudaFuncSetAttribute(K_TC_DX_N(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(256, TC_FWD_STAGES));
  cudaFuncSetAttribute(K_TC_FWD_BIG(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(K_TC_DX_BIG(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(K_TC_DW_BIG(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(K_TC_FWD2(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(K_TC_DX2(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(K_TC_DW(F, V), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(128, TC_DW_STAGES));
  if (!V) {
    cudaFuncSetAttribute(k_dw_adam_tc<F, (F ? 64 : 32)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcAdamCfg<F, (F ? 64 : 32)>::SMEM);
    cudaFuncSetAttribute(k_dw_adam_tc<F, (F ? 32 : 16)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcAdamCfg<F, (F ? 32 : 16)>::SMEM);
  }
}
void tc_set_smem_attr() {
  tc_set_smem_attr_fmt<false, false>(); tc_set_smem_attr_fmt<true, false>();
  tc_set_smem_attr_fmt<false, true>(); tc_set_smem_attr_fmt<true, true>();
}

// fills the tcgen05 view of one GEMM (shapes in the fp32 path's convention); mode 0 fwd, 1 dX, 2 dW
// esz = 2: g.A / g.B point at fp16 operand copies (pitches in elements); the epilogue side of g is unchanged.
// Large-batch forward / dX (more than 256 stacked rows): batch rows per tile and feature sub-tiles per CTA.  256 x 256
// tiles are the efficient shape, but a narrow layer or a rank's slice of the batch leaves most of the 148 SMs without a
// tile (D layer 1 at 3072 local rows: 4 x 12 = 48 CTAs; the generator's dX over K = 12032: 2 x 4 = 8 CTAs, each alone on
// a 188-step contraction): fall back to 128-feature CTAs and then to narrower row tiles until the grid fills the chip.
void tc_pick_tile(int rows, int feats, int nf, int* bn, bool* mt1) {
  *bn = 256; *mt1 = false;
  auto tiles = [&](int b, bool one) { return ((feats + (one ? 127 : 255)) / (one ? 128 : 256)) * ((rows + b - 1) / b) * nf; };
  if (tiles(256, false) >= 120) return;
  *mt1 = true;
  for (int b : {256, 128, 64}) { *bn = b; if (tiles(b, true) >= 120) return; }
}

bool tc_fill_op(EncodeTiledFn fn, TcOp& t, const GemmDesc& g, int mode, int esz = 4, int bn_big = 256) {
  t.g = g;
  t.esz = esz;
  t.ME = g.N; t.NE = g.M; t.KE = g.K;
  const int kb = 128 / esz;   // contraction rows of an MN-major box = elements of one 128-byte row
  if (mode == 0) {            // forward: C[M rows, N feats] = act[M, K] @ W[K, N]
    t.epi = EPI_FWD;
    t.bn = g.M <= 256 ? round_up(g.M, 16) : bn_big;
    return make_map(fn, &t.mapA, g.B, g.N, g.K, g.ldb, kb, true, 0, esz) && make_map(fn, &t.mapB, g.A, g.K, g.M, g.lda, t.bn, false, 0, esz);
  }
  if (mode == 1) {            // dX: C[M rows, N in-feats] = dZ[M, K] @ W[N, K]^T
    t.epi = EPI_DX;
    t.bn = g.M <= 256 ? round_up(g.M, 16) : bn_big;
    return make_map(fn, &t.mapA, g.B, g.K, g.N, g.ldb, 128, false, 0, esz) && make_map(fn, &t.mapB, g.A, g.K, g.M, g.lda, t.bn, false, 0, esz);
  }
  t.epi = EPI_STORE;          // dW: C[M in-feats(+1), N out-feats] = act[K rows, M]^T @ dZ[K rows, N]
  t.bn = g.K > 512 ? 256 : 128;   // a long contraction (large batch) is a regular big GEMM: 256 x 256 tiles
  return make_map(fn, &t.mapA, g.B, g.N, g.K, g.ldb, kb, true, 0, esz) && make_map(fn, &t.mapB, g.A, g.M, g.K, g.lda, kb, true, 0, esz);
}

int tc_debug_gemm(mrgan_handle* h, int mode, const GemmDesc& g, int esz) {
  EncodeTiledFn fn = tc_encoder();
  if (!fn) return fail(h, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  TcOp t; memset(&t, 0, sizeof(t));
  if (!tc_fill_op(fn, t, g, mode, esz)) return fail(h, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  TcOp* d = nullptr;
  CK(cudaMalloc(&d, sizeof(TcOp)));
  CK(cudaMemcpyAsync(d, &t, sizeof(t), cudaMemcpyHostToDevice, h->stream));
  tc_set_smem_attr();
  dim3 grid((t.ME + 127) / 128, (t.NE + t.bn - 1) / t.bn, 1);
  const OperandMode om0{0, 1.0f, nullptr, nullptr};
  const bool f16 = esz == 2;
  if (mode == 0) TC_LAUNCH(h, f16, K_TC_FWD, grid, dim3(TC_FWD_THREADS), tc_smem_bytes(t.bn, TC_FWD_STAGES), h->stream, (const TcOp*)d, h->d_folds, 0, h->hp, om0);
  else if (mode == 1) TC_LAUNCH(h, f16, K_TC_DX, grid, dim3(TC_FWD_THREADS), tc_smem_bytes(t.bn, TC_FWD_STAGES), h->stream, (const TcOp*)d, h->d_folds, 0, h->hp, om0);
  else TC_LAUNCH(h, f16, K_TC_DW, grid, dim3(TC_DW_THREADS), tc_smem_bytes(t.bn, TC_DW_STAGES), h->stream, (const TcOp*)d, h->d_folds, 0, h->hp, om0);
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(d);
  return MRGAN_OK;
}

int tc_setup(mrgan_handle* h) {
  EncodeTiledFn fn = tc_encoder();
  if (!fn) return fail(nullptr, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  if (h->R > 512) h->tc_fused_adam = false;   // large batch: dW is tensor-bound -> big-tile GEMM + flat Adam instead of the fused kernel
  if (const char* pr = getenv("MRGAN_L2PROMO"))
    g_l2_promo = atoi(pr) == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : (atoi(pr) == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                 : (atoi(pr) == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
  const int nf = h->nf;
  // reductions fused into GEMM epilogues: the reference batch regime (one batch tile, batch statistics local to the GPU)
  h->tc_heads = (h->cfg.model == MRGAN_MODEL_GAN && h->tc_fused_adam && !h->d_dpbufs && h->R <= 256 && h->cfg.n_classes <= 32) ? 7 : 0;
  // default: loss, feature-matching and BatchNorm-forward heads.  HEAD_BN_BWD (bit 8) is implemented and parity-tested but
  // measured 1-2 % slower than the stand-alone k_bn_bwd (a thread per feature walks all 50 rows twice behind a GEMM whose
  // own work is tiny), so it is opt-in: MRGAN_HEADS=15.  MRGAN_HEADS=0 runs every reduction as its own kernel (A/B).
  if (const char* hv = getenv("MRGAN_HEADS")) h->tc_heads = h->tc_heads ? (atoi(hv) & 15) : 0;
  if (const char* kr = getenv("MRGAN_DW_SMALL")) h->tc_dw_small = atoi(kr) != 0;
  if (const char* kr = getenv("MRGAN_DW_MERGE")) h->tc_dw_merge = atoi(kr) != 0;
  std::vector<TcOp> ops((size_t)NUM_OPS * nf);
  memset(ops.data(), 0, ops.size() * sizeof(TcOp));
  for (int op = 0; op < NUM_OPS; ++op) {
    const OpInfo& oi = h->ops[op];
    if (!oi.used) continue;
    for (int f = 0; f < nf; ++f) {
      const GemmDesc& g = h->h_descs[(size_t)op * nf + f];
      TcOp& t = ops[(size_t)op * nf + f];
      const int mode = (!oi.at && !oi.bt) ? 0 : ((!oi.at && oi.bt) ? 1 : 2);
      int bn_big = 256;
      if (mode != 2 && oi.maxM > 256) { bool m1; tc_pick_tile(oi.maxM, oi.maxN, nf, &bn_big, &m1); h->tc_mt1[op] = m1; }
      if (h->om.mode == 2) {    // operands come from the fp16 copies (same element offsets and pitches as the fp32 buffers)
        GemmDesc gh = g;
        gh.A = reinterpret_cast<const float*>(h->harena + (g.A - h->om.fbase));
        gh.B = reinterpret_cast<const float*>(h->harena + (g.B - h->om.fbase));
        if (!tc_fill_op(fn, t, gh, mode, 2, bn_big)) return fail(nullptr, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled (fp16 operands) failed");
        t.g = g;                // the epilogue keeps addressing the fp32 buffers (and derives the copies' addresses itself)
      } else if (!tc_fill_op(fn, t, g, mode, 4, bn_big)) return fail(nullptr, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled failed");
      t.net = (op == OP_GW1 || op == OP_GW2 || op == OP_GW3) ? 1 : 0;
      if (h->tc_heads) {
        t.stats = h->d_step_stats + (size_t)f * 4;
        t.stats_stride = nf * 4;
        t.mo_off = h->Mo - h->P; t.vo_off = h->Vo - h->P;
        if (op == OP_D6 && (h->tc_heads & 1)) { t.head = HEAD_DISC; t.hd = h->d_loss + f; t.advance = 1; }
        else if (op == OP_D5G && (h->tc_heads & 2)) { t.head = HEAD_FM; t.hd = h->d_loss + f; t.advance = (h->tc_heads & 8) ? 1 : 0; }
        else if (op == OP_G1 && (h->tc_heads & 4)) { t.head = HEAD_BN; t.hd = h->d_bn + f; }
        else if (op == OP_GX2 && (h->tc_heads & 8)) { t.head = HEAD_BN_BWD; t.hd = h->d_bn + f; }
      }
      if (mode == 2) {
        if (h->tc_fused_adam) t.epi = EPI_ADAM;
        const size_t off = (size_t)(g.C - h->Gr);
        t.P = h->P + off; t.Mo = h->Mo + off; t.Vo = h->Vo + off;
      }
      if (t.bn > h->tc_bn[op]) h->tc_bn[op] = t.bn;
      if (t.ME > h->tc_maxME[op]) h->tc_maxME[op] = t.ME;
      if (t.NE > h->tc_maxNE[op]) h->tc_maxNE[op] = t.NE;
    }
  }
  // Large-batch dW without the fused Adam: the narrow layers' gradients are a handful of 256x256 tiles with a contraction of
  // thousands of rows -> split the contraction over ~2 CTAs per SM, partial products to a workspace, summed in slice order.
  if (!h->tc_fused_adam) {
    size_t ws_floats = 0;
    for (int op = 0; op < NUM_OPS; ++op) {
      const OpInfo& oi = h->ops[op];
      if (!oi.used || !oi.at || h->tc_bn[op] != 256) continue;
      const int kblk = h->om.mode == 2 ? 64 : TC_KBLK;      // contraction rows per stage (one 128-byte row of operand elements)
      const int nkb = (ops[(size_t)op * nf].KE + kblk - 1) / kblk;
      const int tiles = ((h->tc_maxME[op] + 255) / 256) * ((h->tc_maxNE[op] + 255) / 256) * nf;
      int ks = std::min(std::min(2 * 148 / tiles, nkb / 8), 32);
      if (ks < 2) continue;
      const int per = (nkb + ks - 1) / ks;
      ks = (nkb + per - 1) / per;                       // no empty slice
      h->tc_ksplit[op] = ks;
      for (int f = 0; f < nf; ++f) {
        TcOp& t = ops[(size_t)op * nf + f];
        t.ws_stride = t.NE * t.g.ldc;
        t.ws = reinterpret_cast<float*>(ws_floats);     // offset now, pointer once the workspace exists
        ws_floats += (size_t)ks * t.ws_stride;
      }
    }
    if (ws_floats) {
      if (cudaMalloc(&h->d_tcws, ws_floats * sizeof(float)) != cudaSuccess) return fail(nullptr, MRGAN_ERR_CUDA, "cudaMalloc split-K workspace");
      cudaMemset(h->d_tcws, 0, ws_floats * sizeof(float));      // pad columns are never written and must read as zero
      for (int op = 0; op < NUM_OPS; ++op)
        if (h->tc_ksplit[op] > 1)
          for (int f = 0; f < nf; ++f) {
            TcOp& t = ops[(size_t)op * nf + f];
            t.ws = h->d_tcws + reinterpret_cast<size_t>(t.ws);
          }
    }
  }
  if (cudaMalloc(&h->d_tcops, ops.size() * sizeof(TcOp)) != cudaSuccess) return fail(nullptr, MRGAN_ERR_CUDA, "cudaMalloc tc ops");
  if (cudaMemcpy(h->d_tcops, ops.data(), ops.size() * sizeof(TcOp), cudaMemcpyHostToDevice) != cudaSuccess) return fail(nullptr, MRGAN_ERR_CUDA, "cudaMemcpy tc ops");
  if (h->tc_fused_adam && h->tc_adam_tma) {
    std::vector<TcAdamOp> aops((size_t)NUM_OPS * nf);
    memset(aops.data(), 0, aops.size() * sizeof(TcAdamOp));
    for (int op = 0; op < NUM_OPS; ++op) {
      const OpInfo& oi = h->ops[op];
      if (!oi.used || !oi.at) continue;
      for (int f = 0; f < nf; ++f) {
        const TcOp& t = ops[(size_t)op * nf + f];
        TcAdamOp& a = aops[(size_t)op * nf + f];
        a.mapA = t.mapA; a.mapB = t.mapB; a.ME = t.ME; a.NE = t.NE; a.KE = t.KE; a.fold = t.g.fold; a.net = t.net;
        if (h->tc_dw_small) {      // operand boxes of 32 (fp16) / 16 (fp32) batch rows: 16 KB stages
          const GemmDesc& g = t.g;
          bool ok;
          if (h->om.mode == 2) {
            const __half* hA = h->harena + (g.A - h->om.fbase);
            const __half* hB = h->harena + (g.B - h->om.fbase);
            ok = make_map(fn, &a.mapA, hB, g.N, g.K, g.ldb, 32, true, 0, 2) && make_map(fn, &a.mapB, hA, g.M, g.K, g.lda, 32, true, 0, 2);
          } else {
            ok = make_map(fn, &a.mapA, g.B, g.N, g.K, g.ldb, 16, true) && make_map(fn, &a.mapB, g.A, g.M, g.K, g.lda, 16, true);
          }
          if (!ok) return fail(nullptr, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled (dW operands) failed");
        }
        a.ginv = h->om.mode == 2 ? 1.0f / h->om.gscale : 1.0f;
        // fp16 operand copy of W (refreshed with every update), same geometry as the fp32 tensor
        if (h->om.mode == 2 && !make_map(fn, &a.mapH, h->harena + (t.P - h->om.fbase), t.ME, t.NE, t.g.ldc, TCA_KC, false, 128, 2))
          return fail(nullptr, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled (fp16 weight copy) failed");
        // W / m / v as [rows = in+1, cols = out] with the tensor's pitch; box 128 cols x KC rows, clipped at the logical extents
        if (!make_map(fn, &a.mapP, t.P, t.ME, t.NE, t.g.ldc, TCA_KC, false, 128) ||
            !make_map(fn, &a.mapM, t.Mo, t.ME, t.NE, t.g.ldc, TCA_KC, false, 128) ||
            !make_map(fn, &a.mapV, t.Vo, t.ME, t.NE, t.g.ldc, TCA_KC, false, 128))
          return fail(nullptr, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled (optimizer state) failed");
      }
    }
    if (cudaMalloc(&h->d_tcadam, aops.size() * sizeof(TcAdamOp)) != cudaSuccess) return fail(nullptr, MRGAN_ERR_CUDA, "cudaMalloc tc adam ops");
    if (cudaMemcpy(h->d_tcadam, aops.data(), aops.size() * sizeof(TcAdamOp), cudaMemcpyHostToDevice) != cudaSuccess) return fail(nullptr, MRGAN_ERR_CUDA, "cudaMemcpy tc adam ops");
  }
  // what is left for the flat Adam kernel when dW applies Adam itself
  std::vector<AdamRange> r0(nf), r1(nf);
  for (int f = 0; f < nf; ++f) {
    r0[f] = AdamRange{h->net[0][f].off, 0};
    if (h->cfg.model == MRGAN_MODEL_GAN) {
      const NetLayout& LG = h->net[1][f];
      r1[f] = AdamRange{LG.off + LG.t[1].off, LG.t[3].off - LG.t[1].off};     // gamma, beta (contiguous, padded)
    }
  }
  for (int n = 0; n < 2; ++n) {
    if (cudaMalloc(&h->d_ranges_tc[n], nf * sizeof(AdamRange)) != cudaSuccess) return fail(nullptr, MRGAN_ERR_CUDA, "cudaMalloc");
    if (cudaMemcpy(h->d_ranges_tc[n], (n == 0 ? r0 : r1).data(), nf * sizeof(AdamRange), cudaMemcpyHostToDevice) != cudaSuccess) return fail(nullptr, MRGAN_ERR_CUDA, "cudaMemcpy adam ranges");
  }
  tc_set_smem_attr();
  return MRGAN_OK;
}

void tc_teardown(mrgan_handle* h) {
  if (h->d_tcops) cudaFree(h->d_tcops);
  if (h->d_tcadam) cudaFree(h->d_tcadam);
  if (h->d_tcws) cudaFree(h->d_tcws);
  h->d_tcws = nullptr;
  for (int i = 0; i < NUM_OPS; ++i) h->tc_ksplit[i] = 0;
  for (int n = 0; n < 2; ++n) { if (h->d_ranges_tc[n]) cudaFree(h->d_ranges_tc[n]); h->d_ranges_tc[n] = nullptr; }
  for (int i = 0; i < NUM_OPS; ++i) { h->tc_bn[i] = 0; h->tc_maxME[i] = 0; h->tc_maxNE[i] = 0; h->tc_mt1[i] = false; }
  h->d_tcadam = nullptr;
  h->d_tcops = nullptr;
}

void tc_params_changed(mrgan_handle*, int, int) {}   // fp32 master weights are the MMA operands: nothing to refresh

// dW + fused Adam of `nlayers` consecutive ops (op, op + 1, ...) of folds [f0, f0 + nfl) in ONE launch (grid.z = layer x fold;
// CTAs beyond a layer's extents exit at once)
void tc_launch_dw_adam(mrgan_handle* h, int op, int nlayers, int f0, int nfl, cudaStream_t st) {
  const bool f16 = h->om.mode == 2;
  int gx = 0, gy = 0;
  for (int l = 0; l < nlayers; ++l) {
    gx = std::max(gx, (h->tc_maxME[op + l] + 127) / 128);
    gy = std::max(gy, (h->tc_maxNE[op + l] + 127) / 128);
  }
  const dim3 grid(gx, gy, nlayers * nfl);
  const TcAdamOp* ad = h->d_tcadam + (size_t)op * h->nf + f0;
  // small operand stages: three CTAs per SM (default); MRGAN_DW_SMALL=0: the two-CTA configuration (A/B)
  if (f16 && h->tc_dw_small) launch_k(h, k_dw_adam_tc<true, 32>, grid, dim3(192), TcAdamCfg<true, 32>::SMEM, st, ad, h->d_folds, h->hp, nfl, h->nf);
  else if (f16) launch_k(h, k_dw_adam_tc<true, 64>, grid, dim3(192), TcAdamCfg<true, 64>::SMEM, st, ad, h->d_folds, h->hp, nfl, h->nf);
  else if (h->tc_dw_small) launch_k(h, k_dw_adam_tc<false, 16>, grid, dim3(192), TcAdamCfg<false, 16>::SMEM, st, ad, h->d_folds, h->hp, nfl, h->nf);
  else launch_k(h, k_dw_adam_tc<false, 32>, grid, dim3(192), TcAdamCfg<false, 32>::SMEM, st, ad, h->d_folds, h->hp, nfl, h->nf);
}

// the fused-Adam dW path is active (tcgen05 precision, reference batch regime): narrow layers may share launches
bool tc_dw_merge(const mrgan_handle* h) { return h->cfg.precision != MRGAN_PREC_FP32 && h->d_tcadam != nullptr && h->tc_dw_merge; }

bool tc_launch_gemm(mrgan_handle* h, int op, int f0, int nfl, int rows_override, cudaStream_t st) {
  const OpInfo& oi = h->ops[op];
  const bool f16 = h->om.mode == 2;
  const int bn = h->tc_bn[op];
  int NE = h->tc_maxNE[op];
  if (rows_override > 0 && !oi.at) NE = rows_override;
  dim3 grid((h->tc_maxME[op] + 127) / 128, (NE + bn - 1) / bn, nfl);
  const TcOp* d = h->d_tcops + (size_t)op * h->nf + f0;
  if (bn == 256 && (oi.at || h->tc_maxME[op] >= 500) && !h->tc_mt1[op]) {     // large-batch regime
    grid.x = (h->tc_maxME[op] + 255) / 256;
    const size_t smem = tc_smem_bytes(256, TC_BIG_STAGES, 2);
    if (oi.at) {
      const int ks = h->tc_ksplit[op] > 1 ? h->tc_ksplit[op] : 1;
      grid.z = nfl * ks;
      TC_LAUNCH(h, f16, K_TC_DW_BIG, grid, dim3(TC_FWD_THREADS), smem, st, d, h->d_folds, ks, h->hp, h->om);
      if (ks > 1) {
        const int n4 = h->tc_maxNE[op] * pitch8(h->tc_maxME[op]) / 4;
        launch_k(h, k_splitk_reduce, dim3(std::min((n4 + 255) / 256, 1184), nfl, 1), dim3(256), 0, st, d, ks);
      }
    } else if (!oi.bt) TC_LAUNCH(h, f16, K_TC_FWD_BIG, grid, dim3(TC_FWD_THREADS), smem, st, d, h->d_folds, rows_override, h->hp, h->om);
    else TC_LAUNCH(h, f16, K_TC_DX_BIG, grid, dim3(TC_FWD_THREADS), smem, st, d, h->d_folds, rows_override, h->hp, h->om);
    return true;
  }
  const bool head_mt2 = (h->tc_heads & 2) && op == OP_D5G;      // feature matching: one CTA sums the loss over all 250 features
  if (!oi.at && !h->tc_mt1[op] && ((h->tc_mt2 && tc_use_mt2(h->tc_maxME[op], bn)) || head_mt2)) {
    grid.x = (h->tc_maxME[op] + 255) / 256;
    const size_t smem = tc_smem_bytes(bn, TC_FWD_STAGES, 2);
    if (!oi.bt) TC_LAUNCH(h, f16, K_TC_FWD2, grid, dim3(TC_FWD_THREADS), smem, st, d, h->d_folds, rows_override, h->hp, h->om);
    else TC_LAUNCH(h, f16, K_TC_DX2, grid, dim3(TC_FWD_THREADS), smem, st, d, h->d_folds, rows_override, h->hp, h->om);
    return true;
  }
  if (!oi.at && !oi.bt) {
    // one CTA per SM anyway once the stages exceed half the shared memory: take the whole TMEM then (noise parking)
    if (2 * tc_smem_bytes(bn, TC_FWD_STAGES) > 227 * 1024)
      TC_LAUNCH(h, f16, K_TC_FWD_N, grid, dim3(TC_FWD_THREADS), tc_smem_bytes(bn, TC_FWD_STAGES), st, d, h->d_folds, rows_override, h->hp, h->om);
    else
      TC_LAUNCH(h, f16, K_TC_FWD, grid, dim3(TC_FWD_THREADS), tc_smem_bytes(bn, TC_FWD_STAGES), st, d, h->d_folds, rows_override, h->hp, h->om);
  }
  else if (!oi.at && oi.bt) {
    if (2 * tc_smem_bytes(bn, TC_FWD_STAGES) > 227 * 1024)
      TC_LAUNCH(h, f16, K_TC_DX_N, grid, dim3(TC_FWD_THREADS), tc_smem_bytes(bn, TC_FWD_STAGES), st, d, h->d_folds, rows_override, h->hp, h->om);
    else
      TC_LAUNCH(h, f16, K_TC_DX, grid, dim3(TC_FWD_THREADS), tc_smem_bytes(bn, TC_FWD_STAGES), st, d, h->d_folds, rows_override, h->hp, h->om);
  }
  else if (h->d_tcadam) { tc_launch_dw_adam(h, op, 1, f0, nfl, st); return true; }
  else TC_LAUNCH(h, f16, K_TC_DW, grid, dim3(TC_DW_THREADS), tc_smem_bytes(bn, TC_DW_STAGES), st, d, h->d_folds, 0, h->hp, h->om);
  return true;
}

}  // namespace
#endif  // MRGAN_WITH_TC

// ====================================================================== C-ABI
extern "C" {

const char* mrgan_version(void) { return "mrgan-b200 0.2 (sm_100a)"; }

int mrgan_abi_info(int out[4]) {
  if (!out) return fail(nullptr, MRGAN_ERR_ARG, "abi_info: null pointer");
  out[0] = MRGAN_ABI_VERSION; out[1] = (int)sizeof(mrgan_config); out[2] = (int)sizeof(mrgan_fold_shape); out[3] = (int)sizeof(mrgan_epoch_stats);
  return MRGAN_OK;
}

int mrgan_default_config(int model, mrgan_config* cfg) {
  if (!cfg) return fail(nullptr, MRGAN_ERR_ARG, "cfg is null");
  memset(cfg, 0, sizeof(*cfg));
  cfg->model = model;
  cfg->n_folds = 1;
  cfg->n_classes = 6;
  cfg->noise_dim = 100;
  cfg->precision = MRGAN_PREC_FP32;
  cfg->shared_t = 1;
  cfg->eval_each_epoch = 1;
  cfg->device = 0;
  cfg->beta2 = 0.999f; cfg->adam_eps = 1e-8f; cfg->bn_eps = 2e-5f;
  cfg->unlabeled_weight = 1.0f; cfg->sigma_in = 0.3f; cfg->sigma_hidden = 0.5f;
  cfg->hidden_act = MRGAN_ACT_RELU; cfg->leaky_alpha = 0.3f; cfg->dropout = 0.0f;
  if (model == MRGAN_MODEL_NN) { cfg->batch = 20; cfg->lr = 1e-3f; cfg->beta1 = 0.9f; }
  else { cfg->batch = 50; cfg->lr = 6e-4f; cfg->beta1 = 0.5f; }
  return MRGAN_OK;
}

const char* mrgan_last_error(const mrgan_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int mrgan_create(const mrgan_config* cfg, const mrgan_fold_shape* folds, mrgan_handle** out) {
  mrgan_handle* h = nullptr;
  if (!cfg || !folds || !out) return fail(nullptr, MRGAN_ERR_ARG, "null argument");
  *out = nullptr;
  if (cfg->n_folds < 1 || cfg->batch < 1 || cfg->n_classes < 2 || cfg->n_classes > 64 || cfg->noise_dim < 1)
    return fail(nullptr, MRGAN_ERR_ARG, "bad config (n_folds/batch/n_classes/noise_dim)");
  if (cfg->model != MRGAN_MODEL_GAN && cfg->model != MRGAN_MODEL_NN) return fail(nullptr, MRGAN_ERR_ARG, "bad model");
  if (cfg->batch > (1 << 16)) return fail(nullptr, MRGAN_ERR_ARG, "batch too large");
  if ((cfg->hidden_act != MRGAN_ACT_RELU && cfg->hidden_act != MRGAN_ACT_LEAKY_RELU) || !(cfg->dropout >= 0.0f && cfg->dropout < 1.0f))
    return fail(nullptr, MRGAN_ERR_ARG, "bad config (hidden_act / dropout)");
  for (int f = 0; f < cfg->n_folds; ++f) {
    if (folds[f].D < 1 || folds[f].n_train < cfg->batch || folds[f].n_test < 1)
      return fail(nullptr, MRGAN_ERR_ARG, "bad fold shape");
    if (folds[f].n_train != folds[0].n_train)
      return fail(nullptr, MRGAN_ERR_ARG, "all folds of a handle must have the same n_train");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device >= ndev)
    return fail(nullptr, MRGAN_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10 || prop.minor != 0)
    return fail(nullptr, MRGAN_ERR_NO_DEVICE, "device is not sm_100 (B200): the library holds sm_100a code only and has no CPU fallback");
  if (cfg->precision < MRGAN_PREC_FP32 || cfg->precision > MRGAN_PREC_F16) return fail(nullptr, MRGAN_ERR_ARG, "bad config (precision)");
#ifndef MRGAN_WITH_TC
  if (cfg->precision != MRGAN_PREC_FP32) return fail(nullptr, MRGAN_ERR_ARG, "library built without the tcgen05 kernels");
#endif
  h = new mrgan_handle();
  h->cfg = *cfg;
  h->nf = cfg->n_folds;
  h->R = cfg->model == MRGAN_MODEL_GAN ? 3 * cfg->batch : cfg->batch;
  h->n_train = folds[0].n_train;
  h->shapes.assign(folds, folds + h->nf);
  h->fb.resize(h->nf);
  h->hp = AdamHyper{cfg->lr, cfg->beta1, cfg->beta2, cfg->adam_eps, cfg->shared_t, cfg->batch, cfg->batch, 0,
                    cfg->unlabeled_weight, cfg->bn_eps, cfg->n_classes, 0,
                    cfg->hidden_act == MRGAN_ACT_LEAKY_RELU ? cfg->leaky_alpha : 0.0f, cfg->dropout, 1.0f / (1.0f - cfg->dropout)};
  int ne = h->R;
  for (int f = 0; f < h->nf; ++f) ne = folds[f].n_test > ne ? folds[f].n_test : ne;
  h->NE = ne;
  auto cleanup = [&](int code) { mrgan_destroy(h); return code; };
  if (cudaSetDevice(cfg->device) != cudaSuccess) return cleanup(fail(nullptr, MRGAN_ERR_CUDA, "cudaSetDevice failed"));
  // flat parameter space: all D nets, then all G nets
  long long off = 0;
  for (int n = 0; n < (cfg->model == MRGAN_MODEL_GAN ? 2 : 1); ++n) {
    h->net[n].resize(h->nf);
    for (int f = 0; f < h->nf; ++f) {
      NetLayout L = n == 0 ? layout_disc(folds[f].D, cfg->n_classes) : layout_gen(folds[f].D, cfg->noise_dim);
      L.off = off;
      off += L.n;
      h->net[n][f] = L;
    }
  }
  h->n_flat = off;
  Arena sizing;
  layout_buffers(h, sizing);
  h->arena_bytes = sizing.off + 256;
  cudaError_t e = cudaMalloc(&h->arena, h->arena_bytes);
  if (e != cudaSuccess) return cleanup(fail(nullptr, MRGAN_ERR_CUDA, std::string("cudaMalloc arena: ") + cudaGetErrorString(e)));
  e = cudaMemset(h->arena, 0, h->arena_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cleanup(fail(nullptr, MRGAN_ERR_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e)));
  Arena real; real.base = h->arena;
  layout_buffers(h, real);
  h->om = OperandMode{cfg->precision, 1.0f, reinterpret_cast<const float*>(h->arena), nullptr};
  if (cfg->precision == MRGAN_PREC_F16) {      // operand copies: one __half per float of the arena, zero like the arena
    e = cudaMalloc(&h->harena, h->arena_bytes / 2);
    if (e == cudaSuccess) e = cudaMemset(h->harena, 0, h->arena_bytes / 2);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return cleanup(fail(nullptr, MRGAN_ERR_CUDA, std::string("cudaMalloc fp16 operand arena: ") + cudaGetErrorString(e)));
    h->om.hbase = h->harena;
    // gradient-side operands: fp16 times a loss scale (saturating; see OperandMode).  MRGAN_LOSS_SCALE overrides the default.
    h->om.gscale = MRGAN_F16_LOSS_SCALE;
    if (const char* ls = getenv("MRGAN_LOSS_SCALE")) { const float v = (float)atof(ls); if (v >= 1.0f) h->om.gscale = v; }
  }
  bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 16 && ok; ++i) ok = cudaEventCreateWithFlags(&h->ev_pool[i], cudaEventDisableTiming) == cudaSuccess;
  {
#ifdef MRGAN_WITH_TC
    if (const char* mt2 = getenv("MRGAN_MT2")) h->tc_mt2 = atoi(mt2) != 0;
    if (const char* at = getenv("MRGAN_ADAM_TMA")) h->tc_adam_tma = atoi(at) != 0;    // 0: LSU epilogue of k_gemm_tc (A/B switch)
#endif
    const char* pdl = getenv("MRGAN_PDL");
    if (pdl) h->use_pdl = atoi(pdl) != 0;
    const char* env = getenv("MRGAN_CHAINS");
    int nch = env ? atoi(env) : (h->nf >= 32 ? 4 : (h->nf >= 8 ? 2 : 1));
    if (nch < 1) nch = 1;
    if (nch > kMaxChains) nch = kMaxChains;
    if (nch > h->nf) nch = h->nf;
    h->nchains = nch;
    h->cmain[0] = h->stream; h->cside[0] = h->side;
    for (int ch = 1; ch < nch && ok; ++ch) {
      ok = cudaStreamCreateWithFlags(&h->cmain[ch], cudaStreamNonBlocking) == cudaSuccess &&
           cudaStreamCreateWithFlags(&h->cside[ch], cudaStreamNonBlocking) == cudaSuccess;
    }
  }
  ok = ok && cudaEventCreate(&h->ev0) == cudaSuccess && cudaEventCreate(&h->ev1) == cudaSuccess;
  ok = ok && cudaMallocHost(&h->h_epoch_stats, (size_t)h->nf * 8 * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMallocHost(&h->h_scratch, 64 * sizeof(float)) == cudaSuccess;
  for (int s = 0; s < 2 && ok; ++s) {
    ok = ok && cudaMallocHost(&h->h_idx_pinned[s], (size_t)h->nf * 3 * h->n_train * sizeof(int)) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->idx_free[s], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) return cleanup(fail(nullptr, MRGAN_ERR_CUDA, "stream/event/pinned allocation failed"));
  { const int rc = build_descs(h); if (rc != MRGAN_OK) return cleanup(rc); }
  init_ones(h);
  if (cfg->model == MRGAN_MODEL_GAN && cfg->batch > 256) {   // large batch: row-parallel split BatchNorm / feature-matching kernels
    int rc = alloc_split_bufs(h);
    if (rc != MRGAN_OK) return cleanup(rc);
  }
#ifdef MRGAN_WITH_TC
  if (cfg->precision != MRGAN_PREC_FP32) {
    int rc = tc_setup(h);
    if (rc != MRGAN_OK) return cleanup(rc);
  }
#endif
  e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cleanup(fail(nullptr, MRGAN_ERR_CUDA, std::string("create: ") + cudaGetErrorString(e)));
  *out = h;
  return MRGAN_OK;
}

int mrgan_destroy(mrgan_handle* h) {
  if (!h) return MRGAN_OK;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  for (auto& ds : h->datasets) { if (ds.x) cudaFree(ds.x); if (ds.y) cudaFree(ds.y); }
  if (h->d_prep_stats) cudaFree(h->d_prep_stats);
  if (h->d_prep_rows) cudaFree(h->d_prep_rows);
  if (h->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->nccl_comm);
  if (!h->dp_virtual)
    for (int p = 0; p < DP_MAX_RANKS; ++p) {
      if (h->peer_arena[p] && h->peer_arena[p] != h->arena) cudaIpcCloseMemHandle(h->peer_arena[p]);
      if (h->peer_harena[p] && h->peer_harena[p] != h->harena) cudaIpcCloseMemHandle(h->peer_harena[p]);
    }
  if (h->d_dpmem) cudaFree(h->d_dpmem);
  if (h->d_dpbufs) cudaFree(h->d_dpbufs);
#ifdef MRGAN_WITH_TC
  tc_teardown(h);
#endif
  if (h->arena) cudaFree(h->arena);
  if (h->harena) cudaFree(h->harena);
  if (h->h_epoch_stats) cudaFreeHost(h->h_epoch_stats);
  if (h->h_scratch) cudaFreeHost(h->h_scratch);
  for (int s = 0; s < 2; ++s) {
    if (h->h_idx_pinned[s]) cudaFreeHost(h->h_idx_pinned[s]);
    if (h->idx_free[s]) cudaEventDestroy(h->idx_free[s]);
  }
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int i = 0; i < 16; ++i) if (h->ev_pool[i]) cudaEventDestroy(h->ev_pool[i]);
  for (int ch = 1; ch < kMaxChains; ++ch) { if (h->cmain[ch]) cudaStreamDestroy(h->cmain[ch]); if (h->cside[ch]) cudaStreamDestroy(h->cside[ch]); }
  if (h->side) cudaStreamDestroy(h->side);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return MRGAN_OK;
}

int mrgan_sync(mrgan_handle* h) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h);
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return MRGAN_OK;
}

int64_t mrgan_num_params(const mrgan_handle* h, int fold, int net) {
  if (!h || fold < 0 || fold >= h->nf || net < 0 || net > 1 || h->net[net].empty()) return -1;
  return h->net[net][fold].n_ref;
}

int mrgan_set_params(mrgan_handle* h, int fold, int net, const float* src, int64_t n) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (net < 0 || net > 1 || h->net[net].empty()) return fail(h, MRGAN_ERR_ARG, "no such net in this model");
  const NetLayout& L = h->net[net][fold];
  if (!src || n != L.n_ref) return fail(h, MRGAN_ERR_ARG, "set_params: length does not match mrgan_num_params");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  std::vector<float> buf((size_t)L.n, 0.f);
  pack_params(h, fold, net, src, buf, true, nullptr);
  CK(cudaMemcpyAsync(h->P + L.off, buf.data(), (size_t)L.n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  if (h->om.mode == 2) {      // refresh the fp16 operand copy of the uploaded weights
    k_to_half<<<256, 256, 0, h->stream>>>(h->P + L.off, (size_t)L.n, h->om);
    h->launches++;
  }
  CK(cudaStreamSynchronize(h->stream));
#ifdef MRGAN_WITH_TC
  if (h->cfg.precision != MRGAN_PREC_FP32) tc_params_changed(h, fold, net);
#endif
  return MRGAN_OK;
}

static int get_flat(mrgan_handle* h, int fold, int net, const float* dev, float* dst, int64_t n) {
  const NetLayout& L = h->net[net][fold];
  if (!dst || n != L.n_ref) return fail(h, MRGAN_ERR_ARG, "length does not match mrgan_num_params");
  std::vector<float> buf((size_t)L.n);
  CK(cudaMemcpyAsync(buf.data(), dev + L.off, (size_t)L.n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  pack_params(h, fold, net, nullptr, buf, false, dst);
  return MRGAN_OK;
}

int mrgan_get_params(mrgan_handle* h, int fold, int net, float* dst, int64_t n) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (net < 0 || net > 1 || h->net[net].empty()) return fail(h, MRGAN_ERR_ARG, "no such net in this model");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  return get_flat(h, fold, net, h->P, dst, n);
}

int mrgan_get_adam(mrgan_handle* h, int fold, int net, float* m, float* v, int64_t n) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (net < 0 || net > 1 || h->net[net].empty()) return fail(h, MRGAN_ERR_ARG, "no such net in this model");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  rc = get_flat(h, fold, net, h->Mo, m, n); if (rc) return rc;
  return get_flat(h, fold, net, h->Vo, v, n);
}

int mrgan_get_counters(mrgan_handle* h, int fold, int* iterations, int* rng_step) {
  int rc = check_fold(h, fold); if (rc) return rc;
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  FoldState fs;
  CK(cudaMemcpyAsync(&fs, h->d_folds + fold, sizeof(fs), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (iterations) *iterations = h->cfg.shared_t ? fs.iterations : fs.it_net[0];
  if (rng_step) *rng_step = fs.rng_step;
  return MRGAN_OK;
}

int mrgan_load_fold(mrgan_handle* h, int fold, const float* x_train, const int32_t* y_train, const float* x_test,
                    const int32_t* y_test) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (!x_train || !y_train || !x_test || !y_test) return fail(h, MRGAN_ERR_ARG, "load_fold: null pointer");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  const mrgan_fold_shape& s = h->shapes[fold];
  for (int i = 0; i < s.n_train; ++i)
    if (y_train[i] < 0 || y_train[i] >= h->cfg.n_classes) return fail(h, MRGAN_ERR_ARG, "load_fold: y_train label out of range");
  for (int i = 0; i < s.n_test; ++i)
    if (y_test[i] < 0 || y_test[i] >= h->cfg.n_classes) return fail(h, MRGAN_ERR_ARG, "load_fold: y_test label out of range");
  FoldBuffers& b = h->fb[fold];
  const size_t w = (size_t)s.D * sizeof(float);
  CK(cudaMemcpy2DAsync(b.xtr, (size_t)pitch8(s.D) * sizeof(float), x_train, w, w, s.n_train, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpy2DAsync(b.xte, (size_t)b.lda[0] * sizeof(float), x_test, w, w, s.n_test, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(b.ytr, y_train, (size_t)s.n_train * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(b.yte, y_test, (size_t)s.n_test * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (h->cfg.precision == MRGAN_PREC_TF32) {   // X_test feeds the first eval MMA directly: put it on the tf32 grid (RN) once
    k_round_tf32<<<256, 256, 0, h->stream>>>(b.xte, (size_t)s.n_test * b.lda[0]);
    h->launches++;
  } else if (h->om.mode == 2) {                // ... or make its fp16 operand copy (ones column included)
    k_to_half<<<256, 256, 0, h->stream>>>(b.xte, (size_t)s.n_test * b.lda[0], h->om);
    h->launches++;
  }
  CK(cudaStreamSynchronize(h->stream));
  b.loaded = true;
  return MRGAN_OK;
}

int mrgan_load_dataset(mrgan_handle* h, int slot, const float* x, const int32_t* y, int n_rows, int D) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (slot < 0 || slot >= 8 || !x || !y || n_rows < 1 || D < 1) return fail(h, MRGAN_ERR_ARG, "load_dataset: bad argument");
  for (int i = 0; i < n_rows; ++i)
    if (y[i] < 0 || y[i] >= h->cfg.n_classes) return fail(h, MRGAN_ERR_ARG, "load_dataset: label out of range");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  mrgan_handle::Dataset& ds = h->datasets[slot];
  if (ds.x && (ds.n != n_rows || ds.D != D)) { cudaFree(ds.x); cudaFree(ds.y); ds.x = nullptr; ds.y = nullptr; }
  ds.n = n_rows; ds.D = D; ds.ld = pitch8(D);
  if (!ds.x) {       // a re-upload of the same shape reuses the buffers (cudaFree / cudaMalloc synchronise the device)
    CK(cudaMalloc(&ds.x, (size_t)n_rows * ds.ld * sizeof(float)));
    CK(cudaMalloc(&ds.y, (size_t)n_rows * sizeof(int)));
  }
  CK(cudaMemcpy2DAsync(ds.x, (size_t)ds.ld * sizeof(float), x, (size_t)D * sizeof(float), (size_t)D * sizeof(float), n_rows,
                       cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(ds.y, y, (size_t)n_rows * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MRGAN_OK;
}

int mrgan_prepare_fold(mrgan_handle* h, int fold, int slot, const int32_t* train_rows, const int32_t* test_rows) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (slot < 0 || slot >= 8 || !h->datasets[slot].x) return fail(h, MRGAN_ERR_STATE, "prepare_fold: dataset slot is empty");
  if (!train_rows || !test_rows) return fail(h, MRGAN_ERR_ARG, "prepare_fold: null pointer");
  const mrgan_handle::Dataset& ds = h->datasets[slot];
  const mrgan_fold_shape& s = h->shapes[fold];
  if (ds.D != s.D) return fail(h, MRGAN_ERR_ARG, "prepare_fold: dataset width differs from the fold's D");
  for (int i = 0; i < s.n_train; ++i) if ((unsigned)train_rows[i] >= (unsigned)ds.n) return fail(h, MRGAN_ERR_ARG, "prepare_fold: train row out of range");
  for (int i = 0; i < s.n_test; ++i) if ((unsigned)test_rows[i] >= (unsigned)ds.n) return fail(h, MRGAN_ERR_ARG, "prepare_fold: test row out of range");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  const int nrows = s.n_train + s.n_test;
  if (nrows > h->prep_rows_cap) {
    if (h->d_prep_rows) cudaFree(h->d_prep_rows);
    CK(cudaMalloc(&h->d_prep_rows, (size_t)nrows * sizeof(int)));
    h->prep_rows_cap = nrows;
  }
  if (2 * s.D > h->prep_stats_cap) {      // per-slice partial column sums: [PREP_SLICES][2 * D] doubles
    if (h->d_prep_stats) cudaFree(h->d_prep_stats);
    CK(cudaMalloc(&h->d_prep_stats, (size_t)PREP_SLICES * 2 * s.D * sizeof(double)));
    h->prep_stats_cap = 2 * s.D;
  }
  FoldBuffers& b = h->fb[fold];
  CK(cudaMemcpyAsync(h->d_prep_rows, train_rows, (size_t)s.n_train * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_prep_rows + s.n_train, test_rows, (size_t)s.n_test * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  const int gx = (s.D + 127) / 128;
  k_col_stats<<<dim3(gx, PREP_SLICES), 128, 0, h->stream>>>(ds.x, ds.ld, h->d_prep_rows, s.n_train, s.D, h->d_prep_stats);
  k_scale_gather<<<dim3(gx, 64), 128, 0, h->stream>>>(ds.x, ds.ld, h->d_prep_rows, s.n_train, s.D, h->d_prep_stats, PREP_SLICES, s.n_train,
                                                     b.xtr, pitch8(s.D), ds.y, b.ytr, OperandMode{0, 1.0f, nullptr, nullptr});
  k_scale_gather<<<dim3(gx, 64), 128, 0, h->stream>>>(ds.x, ds.ld, h->d_prep_rows + s.n_train, s.n_test, s.D, h->d_prep_stats, PREP_SLICES, s.n_train,
                                                     b.xte, b.lda[0], ds.y, b.yte, h->om);
  h->launches += 3;
  CK(cudaStreamSynchronize(h->stream));     // the pageable index arrays are borrowed only for the call
  CKS();
  b.loaded = true;
  return MRGAN_OK;
}

int mrgan_disc_step(mrgan_handle* h, int fold, const float* x_lab, const int32_t* labels, const float* x_unl,
                    const float* z, float out[3]) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (h->cfg.model != MRGAN_MODEL_GAN) return fail(h, MRGAN_ERR_STATE, "disc_step needs a GAN handle");
  if (!x_lab || !labels || !x_unl || !z || !out) return fail(h, MRGAN_ERR_ARG, "disc_step: null pointer");
  if (h->dp_virtual) return fail(h, MRGAN_ERR_STATE, "disc_step: virtual ranks step together: use mrgan_train_epoch");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  const int B = h->cfg.batch, D = h->shapes[fold].D, nd = h->cfg.noise_dim;
  for (int i = 0; i < B; ++i)
    if (labels[i] < 0 || labels[i] >= h->cfg.n_classes) return fail(h, MRGAN_ERR_ARG, "disc_step: label out of range");
  FoldBuffers& b = h->fb[fold];
  const size_t w = (size_t)D * sizeof(float), ld = (size_t)pitch8(D) * sizeof(float);
  CK(cudaMemcpy2DAsync(b.stage_x, ld, x_lab, w, w, B, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpy2DAsync(b.stage_x + (size_t)B * pitch8(D), ld, x_unl, w, w, B, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(b.stage_y, labels, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(b.stage_z, z, (size_t)B * nd * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  enqueue_disc_step(h, fold, 1, 0, 1);
  CK(cudaMemcpyAsync(h->h_scratch, h->d_step_stats + (size_t)fold * 4, 4 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKS();
  out[0] = h->h_scratch[0]; out[1] = h->h_scratch[1]; out[2] = h->h_scratch[2];
  return MRGAN_OK;
}

int mrgan_gen_step(mrgan_handle* h, int fold, const float* x_unl, const float* z, float out[1]) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (h->cfg.model != MRGAN_MODEL_GAN) return fail(h, MRGAN_ERR_STATE, "gen_step needs a GAN handle");
  if (!x_unl || !z || !out) return fail(h, MRGAN_ERR_ARG, "gen_step: null pointer");
  if (h->dp_virtual) return fail(h, MRGAN_ERR_STATE, "gen_step: virtual ranks step together: use mrgan_train_epoch");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  const int B = h->cfg.batch, D = h->shapes[fold].D, nd = h->cfg.noise_dim;
  FoldBuffers& b = h->fb[fold];
  const size_t w = (size_t)D * sizeof(float), ld = (size_t)pitch8(D) * sizeof(float);
  CK(cudaMemcpy2DAsync(b.stage_x + (size_t)B * pitch8(D), ld, x_unl, w, w, B, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(b.stage_z, z, (size_t)B * nd * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  enqueue_gen_step(h, fold, 1, 0, 1);
  CK(cudaMemcpyAsync(h->h_scratch, h->d_step_stats + (size_t)fold * 4, 4 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKS();
  out[0] = h->h_scratch[3];
  return MRGAN_OK;
}

int mrgan_test_batch(mrgan_handle* h, int fold, const float* x, const int32_t* y, int n, float* err) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (!x || !y || !err) return fail(h, MRGAN_ERR_ARG, "test_batch: null pointer");
  if (n < 1 || n > h->NE) return fail(h, MRGAN_ERR_ARG, "test_batch: n exceeds max(n_test, rows of a train step)");
  for (int i = 0; i < n; ++i)
    if (y[i] < 0 || y[i] >= h->cfg.n_classes) return fail(h, MRGAN_ERR_ARG, "test_batch: label out of range");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  FoldBuffers& b = h->fb[fold];
  const int D = h->shapes[fold].D;
  const size_t w = (size_t)D * sizeof(float);
  CK(cudaMemcpy2DAsync(b.ex_stage, (size_t)b.lda[0] * sizeof(float), x, w, w, n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(b.ey_stage, y, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (h->cfg.precision == MRGAN_PREC_TF32) {
    k_round_tf32<<<64, 256, 0, h->stream>>>(b.ex_stage, (size_t)n * b.lda[0]);
    h->launches++;
  } else if (h->om.mode == 2) {
    k_to_half<<<64, 256, 0, h->stream>>>(b.ex_stage, (size_t)n * b.lda[0], h->om);
    h->launches++;
  }
  enqueue_eval(h, fold, 1, true, n);
  CK(cudaMemcpyAsync(h->h_scratch, b.eval_out_s, 4 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKS();
  err[0] = h->h_scratch[1];
  return MRGAN_OK;
}

int mrgan_eval(mrgan_handle* h, int fold, float* err) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (!err) return fail(h, MRGAN_ERR_ARG, "eval: null pointer");
  if (!h->fb[fold].loaded) return fail(h, MRGAN_ERR_STATE, "eval before mrgan_load_fold");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  enqueue_eval(h, fold, 1, false, 0);
  CK(cudaMemcpyAsync(h->h_scratch, h->fb[fold].eval_out, 4 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKS();
  err[0] = h->h_scratch[1];
  return MRGAN_OK;
}

int mrgan_epoch_result(mrgan_handle* h, mrgan_epoch_stats* stats) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  CKS();
  if (stats)
    for (int f = 0; f < h->nf; ++f) {
      const float* s = h->h_epoch_stats + (size_t)f * 8;
      stats[f] = mrgan_epoch_stats{s[0], s[1], s[2], s[3], s[4]};
    }
  return MRGAN_OK;
}

int mrgan_set_epoch_rows(mrgan_handle* h, int fold, const int32_t* lab_rows, int n_lab, const int32_t* unl_rows, int n_unl) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (h->cfg.model != MRGAN_MODEL_GAN) return fail(h, MRGAN_ERR_STATE, "set_epoch_rows needs a GAN handle");
  const int N = h->n_train;
  if (!lab_rows || n_lab < 1 || n_lab > N || n_lab > PERM_MAX) return fail(h, MRGAN_ERR_ARG, "set_epoch_rows: 1 <= n_lab <= min(n_train, 8192)");
  if (unl_rows ? (n_unl < 1 || n_unl > N || n_unl > PERM_MAX) : (n_unl != 0 || N > PERM_MAX))
    return fail(h, MRGAN_ERR_ARG, "set_epoch_rows: unlabeled subset must have 1..min(n_train, 8192) rows (none: n_train <= 8192)");
  for (int i = 0; i < n_lab; ++i) if ((unsigned)lab_rows[i] >= (unsigned)N) return fail(h, MRGAN_ERR_ARG, "set_epoch_rows: labeled row out of range");
  for (int i = 0; i < n_unl; ++i) if ((unsigned)unl_rows[i] >= (unsigned)N) return fail(h, MRGAN_ERR_ARG, "set_epoch_rows: unlabeled row out of range");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  FoldBuffers& b = h->fb[fold];
  CK(cudaMemcpyAsync(b.lab_rows, lab_rows, (size_t)n_lab * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  if (unl_rows) CK(cudaMemcpyAsync(b.unl_rows, unl_rows, (size_t)n_unl * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  b.n_lab = n_lab; b.n_unl = unl_rows ? n_unl : 0;
  PermDesc pd{b.lab_rows, unl_rows ? b.unl_rows : nullptr, b.idx, n_lab, b.n_unl, N,
              (uint32_t)(h->shapes[fold].seed & 0xFFFFFFFFull), (uint32_t)(h->shapes[fold].seed >> 32)};
  CK(cudaMemcpyAsync(h->d_perm + fold, &pd, sizeof(pd), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));     // the pageable row arrays are borrowed only for the call
  return MRGAN_OK;
}

static int launch_epoch(mrgan_handle* h, int nb, mrgan_epoch_stats* stats);

int mrgan_train_epoch_seeded(mrgan_handle* h, uint32_t epoch, mrgan_epoch_stats* stats) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (h->cfg.model != MRGAN_MODEL_GAN) return fail(h, MRGAN_ERR_STATE, "train_epoch_seeded needs a GAN handle");
  for (int f = 0; f < h->nf; ++f) {
    if (!h->fb[f].loaded) return fail(h, MRGAN_ERR_STATE, "train_epoch_seeded before mrgan_load_fold");
    if (h->fb[f].n_lab < 1) return fail(h, MRGAN_ERR_STATE, "train_epoch_seeded before mrgan_set_epoch_rows");
  }
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_epoch_perm, cudaFuncAttributeMaxDynamicSharedMemorySize, PERM_MAX * 8); attr_set = true; }
  k_epoch_perm<<<dim3(3, h->nf), 1024, PERM_MAX * 8, h->stream>>>(h->d_perm, epoch);
  h->launches++;
  return launch_epoch(h, h->n_train / h->cfg.batch, stats);
}

int mrgan_debug_epoch_indices(mrgan_handle* h, int fold, int32_t* dst) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (!dst) return fail(h, MRGAN_ERR_ARG, "debug_epoch_indices: null pointer");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  CK(cudaMemcpyAsync(dst, h->fb[fold].idx, (size_t)3 * h->n_train * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MRGAN_OK;
}

int mrgan_train_epoch(mrgan_handle* h, const int32_t* idx_lab, const int32_t* idx_unl, const int32_t* idx_unl2,
                      mrgan_epoch_stats* stats) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (h->cfg.model != MRGAN_MODEL_GAN) return fail(h, MRGAN_ERR_STATE, "train_epoch needs a GAN handle");
  if (!idx_lab || !idx_unl || !idx_unl2) return fail(h, MRGAN_ERR_ARG, "train_epoch: null index array");
  for (int f = 0; f < h->nf; ++f)
    if (!h->fb[f].loaded) return fail(h, MRGAN_ERR_STATE, "train_epoch before mrgan_load_fold");
  const int nb = h->n_train / h->cfg.batch;
  const size_t tot = (size_t)h->nf * h->n_train;
  for (size_t i = 0; i < tot; ++i)
    if ((unsigned)idx_lab[i] >= (unsigned)h->n_train || (unsigned)idx_unl[i] >= (unsigned)h->n_train ||
        (unsigned)idx_unl2[i] >= (unsigned)h->n_train)
      return fail(h, MRGAN_ERR_ARG, "train_epoch: row index out of range");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  const int32_t* streams[3] = {idx_lab, idx_unl, idx_unl2};
  rc = upload_indices(h, streams, 3, h->n_train); if (rc) return rc;
  return launch_epoch(h, nb, stats);
}

// the epoch itself (index arrays already on their way on the handle's stream): one graph launch, or step by step
static int launch_epoch(mrgan_handle* h, int nb, mrgan_epoch_stats* stats) {
  // data-parallel epochs are captured like the others (NCCL collectives and the fused exchange are stream-ordered graph
  // nodes; one chain, so every rank issues them in the same order); MRGAN_DP_GRAPH=0 enqueues them step by step instead
  static const bool dp_graph = !(getenv("MRGAN_DP_GRAPH") && atoi(getenv("MRGAN_DP_GRAPH")) == 0);
  const bool eager = h->dp_world > 1 && !dp_graph;
  int rc = MRGAN_OK;
  if (!eager) { rc = build_graph(h, nb, nb, 0); if (rc) return rc; }
  CK(cudaEventRecord(h->ev0, h->stream));
  if (eager) {
    for (int t = 0; t < nb; ++t) { enqueue_disc_step(h, 0, h->nf, t, 0); enqueue_gen_step(h, 0, h->nf, t, 0); }
    if (h->cfg.eval_each_epoch) enqueue_eval(h, 0, h->nf, false, 0);
    k_epoch_reduce<<<h->nf, 32, 0, h->stream>>>(h->d_step_stats, h->d_eval, h->d_epoch_stats, h->nf, nb, h->cfg.eval_each_epoch);
    h->launches++;
    CK(cudaMemcpyAsync(h->h_epoch_stats, h->d_epoch_stats, (size_t)h->nf * 8 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  } else {
    CK(cudaGraphLaunch(h->graphs[nb], h->stream));
    h->launches += h->graph_nodes[nb];
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  h->epoch_pending = true;
  { const int sk = take_sticky(h); if (sk) return sk; }
  if (stats) return mrgan_epoch_result(h, stats);
  return MRGAN_OK;
}

int mrnn_step(mrgan_handle* h, int fold, const float* x, const int32_t* labels, int n, float out[2]) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (h->cfg.model != MRGAN_MODEL_NN) return fail(h, MRGAN_ERR_STATE, "mrnn_step needs an NN handle");
  if (!x || !labels || !out) return fail(h, MRGAN_ERR_ARG, "mrnn_step: null pointer");
  if (n < 1 || n > h->cfg.batch) return fail(h, MRGAN_ERR_ARG, "mrnn_step: n must be in [1, batch]");
  for (int i = 0; i < n; ++i)
    if (labels[i] < 0 || labels[i] >= h->cfg.n_classes) return fail(h, MRGAN_ERR_ARG, "mrnn_step: label out of range");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  FoldBuffers& b = h->fb[fold];
  const int D = h->shapes[fold].D;
  const size_t w = (size_t)D * sizeof(float);
  CK(cudaMemcpy2DAsync(b.stage_x, (size_t)pitch8(D) * sizeof(float), x, w, w, n, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(b.stage_y, labels, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  enqueue_nn_step(h, fold, 1, 0, 1, n);
  CK(cudaMemcpyAsync(h->h_scratch, h->d_step_stats + (size_t)fold * 4, 4 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKS();
  out[0] = h->h_scratch[0]; out[1] = h->h_scratch[1];
  return MRGAN_OK;
}

int mrnn_train_epoch(mrgan_handle* h, const int32_t* idx, int n_idx, float* loss_acc) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (h->cfg.model != MRGAN_MODEL_NN) return fail(h, MRGAN_ERR_STATE, "mrnn_train_epoch needs an NN handle");
  if (!idx || n_idx < h->cfg.batch || n_idx > h->n_train || n_idx % h->cfg.batch)
    return fail(h, MRGAN_ERR_ARG, "mrnn_train_epoch: n_idx must be a multiple of batch in [batch, n_train]");
  for (int f = 0; f < h->nf; ++f)
    if (!h->fb[f].loaded) return fail(h, MRGAN_ERR_STATE, "mrnn_train_epoch before mrgan_load_fold");
  for (size_t i = 0; i < (size_t)h->nf * n_idx; ++i)
    if ((unsigned)idx[i] >= (unsigned)h->n_train) return fail(h, MRGAN_ERR_ARG, "mrnn_train_epoch: row index out of range");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  const int nb = n_idx / h->cfg.batch;
  rc = build_graph(h, nb, nb, n_idx); if (rc) return rc;
  const int32_t* streams[1] = {idx};
  rc = upload_indices(h, streams, 1, n_idx); if (rc) return rc;
  CK(cudaEventRecord(h->ev0, h->stream));
  CK(cudaGraphLaunch(h->graphs[nb], h->stream));
  CK(cudaEventRecord(h->ev1, h->stream));
  h->launches += h->graph_nodes[nb];
  h->epoch_pending = true;
  if (loss_acc) {
    rc = finish_pending(h); if (rc) return rc;
    CKS();
    for (int f = 0; f < h->nf; ++f) { loss_acc[2 * f] = h->h_epoch_stats[f * 8]; loss_acc[2 * f + 1] = h->h_epoch_stats[f * 8 + 1]; }
  }
  return MRGAN_OK;
}

int mrnn_evaluate(mrgan_handle* h, int fold, float out[2]) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (!out) return fail(h, MRGAN_ERR_ARG, "evaluate: null pointer");
  if (!h->fb[fold].loaded) return fail(h, MRGAN_ERR_STATE, "evaluate before mrgan_load_fold");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  enqueue_eval(h, fold, 1, false, 0);
  CK(cudaMemcpyAsync(h->h_scratch, h->fb[fold].eval_out, 4 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKS();
  out[0] = h->h_scratch[2]; out[1] = 1.0f - h->h_scratch[1];
  return MRGAN_OK;
}

int mrgan_fill_normal(mrgan_handle* h, int fold, int step, int tensor_id, int rows, int cols, int row0, float* dst) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (!dst || rows < 1 || cols < 1) return fail(h, MRGAN_ERR_ARG, "fill_normal: bad argument");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  float* d = nullptr;
  const size_t n = (size_t)rows * cols;
  CK(cudaMalloc(&d, n * sizeof(float)));
  k_fill_normal<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(d, h->d_folds, fold, step, tensor_id, rows, cols, row0);
  h->launches++;
  cudaError_t e = cudaMemcpyAsync(dst, d, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(h, MRGAN_ERR_CUDA, cudaGetErrorString(e));
  return MRGAN_OK;
}

int mrgan_adam_flat(mrgan_handle* h, float* p, float* m, float* v, const float* g, int64_t n, int t) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (!p || !m || !v || !g || n < 1 || t < 1) return fail(h, MRGAN_ERR_ARG, "adam_flat: bad argument");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  const int64_t n4 = (n + 3) / 4;
  float* d = nullptr;
  CK(cudaMalloc(&d, (size_t)n4 * 16 * 4));
  CK(cudaMemsetAsync(d, 0, (size_t)n4 * 16 * 4, h->stream));
  float *dp = d, *dm = d + n4 * 4, *dv = d + n4 * 8, *dg = d + n4 * 12;
  cudaMemcpyAsync(dp, p, n * 4, cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(dm, m, n * 4, cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(dv, v, n * 4, cudaMemcpyHostToDevice, h->stream);
  cudaMemcpyAsync(dg, g, n * 4, cudaMemcpyHostToDevice, h->stream);
  const double b1t = pow((double)h->hp.b1, (double)t), b2t = pow((double)h->hp.b2, (double)t);
  const float lr_t = (float)((double)h->hp.lr * sqrt(1.0 - b2t) / (1.0 - b1t));
  int blocks = (int)((n4 + 255) / 256); if (blocks > 4096) blocks = 4096;
  k_adam_plain<<<blocks, 256, 0, h->stream>>>((float4*)dp, (float4*)dm, (float4*)dv, (const float4*)dg, n4, lr_t,
                                              h->hp.b1, h->hp.b2, h->hp.eps);
  h->launches++;
  cudaMemcpyAsync(p, dp, n * 4, cudaMemcpyDeviceToHost, h->stream);
  cudaMemcpyAsync(m, dm, n * 4, cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e = cudaMemcpyAsync(v, dv, n * 4, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(d);
  if (e != cudaSuccess) return fail(h, MRGAN_ERR_CUDA, cudaGetErrorString(e));
  return MRGAN_OK;
}

int mrgan_time_op(mrgan_handle* h, int which, int reps, float* ms_avg) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (!ms_avg || reps < 1) return fail(h, MRGAN_ERR_ARG, "time_op: bad argument");
  const bool gan = h->cfg.model == MRGAN_MODEL_GAN;
  if (!gan && (which == MRGAN_TIME_ADAM_G || which == MRGAN_TIME_GEN_STEP || which == MRGAN_TIME_DX1)) return fail(h, MRGAN_ERR_STATE, "time_op: no generator in this model");
  for (int f = 0; f < h->nf; ++f)
    if (!h->fb[f].loaded) return fail(h, MRGAN_ERR_STATE, "time_op before mrgan_load_fold");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  auto once = [&]() {
    switch (which) {
      case MRGAN_TIME_ADAM_D: launch_adam(h, 0, h->nf, 0, false); break;
      case MRGAN_TIME_ADAM_G: launch_adam(h, 0, h->nf, 1, false); break;
      case MRGAN_TIME_DW1: launch_gemm(h, OP_DW1, 0, h->nf, 0); break;
      case MRGAN_TIME_FWD1: launch_gemm(h, OP_D1, 0, h->nf, 0); break;
      case MRGAN_TIME_DISC_STEP: if (gan) enqueue_disc_step(h, 0, h->nf, 0, 0); else enqueue_nn_step(h, 0, h->nf, 0, 0, h->cfg.batch); break;
      case MRGAN_TIME_GEN_STEP: enqueue_gen_step(h, 0, h->nf, 0, 0); break;
      case MRGAN_TIME_DX1: launch_gemm(h, OP_DX1G, 0, h->nf, 0); break;
      default: break;
    }
  };
  if (which < 0 || which > MRGAN_TIME_DX1) return fail(h, MRGAN_ERR_ARG, "time_op: unknown op");
  once();                                   // warm-up
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < reps; ++i) once();
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKS();
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  *ms_avg = ms / reps;
  return MRGAN_OK;
}

int mrgan_nccl_unique_id(void* id128) {
  if (!id128) return fail(nullptr, MRGAN_ERR_ARG, "nccl_unique_id: null pointer");
  if (!nccl_load()) return fail(nullptr, MRGAN_ERR_STATE, "libnccl.so.2 not found");
  NcclId id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail(nullptr, MRGAN_ERR_CUDA, "ncclGetUniqueId failed");
  memcpy(id128, id.b, 128);
  return MRGAN_OK;
}

int mrgan_dp_init(mrgan_handle* h, int rank, int world, const void* id128) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (world < 1 || rank < 0 || rank >= world || !id128) return fail(h, MRGAN_ERR_ARG, "dp_init: bad rank / world / id");
  if (world > DP_MAX_RANKS) return fail(h, MRGAN_ERR_ARG, "dp_init: at most 8 ranks (one NVSwitch domain)");
  if (h->cfg.model != MRGAN_MODEL_GAN) return fail(h, MRGAN_ERR_STATE, "dp_init: data-parallel mode is implemented for the GAN model");
  if (h->dp_world > 1 || h->nccl_comm) return fail(h, MRGAN_ERR_STATE, "dp_init called twice");
  if (world == 1) return MRGAN_OK;
  if (h->cfg.batch % 4) return fail(h, MRGAN_ERR_ARG, "dp_init: the local batch must be a multiple of 4 (noise row groups)");
  if (!nccl_load()) return fail(h, MRGAN_ERR_STATE, "libnccl.so.2 not found");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  NcclId id; memcpy(id.b, id128, 128);
  const int nrc = g_nccl.CommInitRank(&h->nccl_comm, world, id, rank);
  if (nrc != 0) return fail(h, MRGAN_ERR_CUDA, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error"));
  rc = alloc_split_bufs(h); if (rc) return rc;
  h->dp_world = world; h->dp_rank = rank;
  h->hp.dp_bloc = h->cfg.batch; h->hp.dp_bg = h->cfg.batch * world; h->hp.dp_rank = rank;
#ifdef MRGAN_WITH_TC
  if (h->cfg.precision != MRGAN_PREC_FP32) {   // gradients must be all-reduced before Adam: dW stores them, k_adam applies
    tc_teardown(h);
    h->tc_fused_adam = false;
    rc = tc_setup(h); if (rc) return rc;
  }
#endif
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  h->graphs.clear(); h->graph_nodes.clear();
  CK(cudaDeviceSynchronize());
  return MRGAN_OK;
}

int mrgan_dp_ipc_export(mrgan_handle* h, void* out128) {
  if (!h || !out128) return fail(h, MRGAN_ERR_ARG, "dp_ipc_export: null argument");
  CK(cudaSetDevice(h->cfg.device));
  memset(out128, 0, 128);
  cudaIpcMemHandle_t a;
  CK(cudaIpcGetMemHandle(&a, h->arena));
  memcpy(out128, &a, sizeof(a));
  if (h->harena) {
    CK(cudaIpcGetMemHandle(&a, h->harena));
    memcpy(static_cast<char*>(out128) + 64, &a, sizeof(a));
  }
  return MRGAN_OK;
}

int mrgan_dp_ipc_open(mrgan_handle* h, const void* handles, int world) {
  if (!h || !handles) return fail(h, MRGAN_ERR_ARG, "dp_ipc_open: null argument");
  if (h->dp_world != world || world < 2 || world > DP_MAX_RANKS || h->dp_virtual)
    return fail(h, MRGAN_ERR_STATE, "dp_ipc_open: call after mrgan_dp_init with the same world size (at most 8 ranks)");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  const char* hb = static_cast<const char*>(handles);
  for (int p = 0; p < world; ++p) {
    if (p == h->dp_rank) { h->peer_arena[p] = h->arena; h->peer_harena[p] = h->harena; continue; }
    cudaIpcMemHandle_t a;
    memcpy(&a, hb + (size_t)p * 128, sizeof(a));
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, a, cudaIpcMemLazyEnablePeerAccess));
    h->peer_arena[p] = static_cast<char*>(ptr);
    if (h->harena) {
      memcpy(&a, hb + (size_t)p * 128 + 64, sizeof(a));
      CK(cudaIpcOpenMemHandle(&ptr, a, cudaIpcMemLazyEnablePeerAccess));
      h->peer_harena[p] = static_cast<__half*>(ptr);
    }
  }
  h->dp_fused = !(getenv("MRGAN_DP_FUSED") && atoi(getenv("MRGAN_DP_FUSED")) == 0);
  dp_build_peers(h);
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  h->graphs.clear(); h->graph_nodes.clear();
  return MRGAN_OK;
}

int mrgan_dp_init_virtual(mrgan_handle* h, int world) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (h->cfg.model != MRGAN_MODEL_GAN) return fail(h, MRGAN_ERR_STATE, "dp_init_virtual: data-parallel mode is implemented for the GAN model");
  if (h->dp_world > 1 || h->nccl_comm) return fail(h, MRGAN_ERR_STATE, "dp_init called twice");
  if (world < 2 || world != h->nf) return fail(h, MRGAN_ERR_ARG, "dp_init_virtual: the handle must hold exactly `world` folds (one per virtual rank)");
  if (h->cfg.batch % 4) return fail(h, MRGAN_ERR_ARG, "dp_init_virtual: the local batch must be a multiple of 4 (noise row groups)");
  for (int f = 1; f < h->nf; ++f)
    if (h->shapes[f].D != h->shapes[0].D || h->shapes[f].seed != h->shapes[0].seed)
      return fail(h, MRGAN_ERR_ARG, "dp_init_virtual: virtual ranks are replicas: same input width and noise key");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  rc = alloc_split_bufs(h); if (rc) return rc;
  if (world > DP_MAX_RANKS) return fail(h, MRGAN_ERR_ARG, "dp_init_virtual: at most 8 ranks");
  h->dp_world = world; h->dp_rank = 0; h->dp_virtual = true;
  h->hp.dp_bloc = h->cfg.batch; h->hp.dp_bg = h->cfg.batch * world; h->hp.dp_rank = -1;     // -1: rank = fold index (common.cuh)
  h->dp_fused = !(getenv("MRGAN_DP_FUSED") && atoi(getenv("MRGAN_DP_FUSED")) == 0);
  dp_build_peers(h);
#ifdef MRGAN_WITH_TC
  if (h->cfg.precision != MRGAN_PREC_FP32) {   // gradients must be all-reduced before Adam: dW stores them, k_adam applies
    tc_teardown(h);
    h->tc_fused_adam = false;
    rc = tc_setup(h); if (rc) return rc;
  }
#endif
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  h->graphs.clear(); h->graph_nodes.clear();
  CK(cudaDeviceSynchronize());
  return MRGAN_OK;
}

int mrgan_debug_buffer(mrgan_handle* h, int fold, int which, float* dst, int rows, int cols) {
  int rc = check_fold(h, fold); if (rc) return rc;
  if (!dst || rows < 1 || cols < 1) return fail(h, MRGAN_ERR_ARG, "debug_buffer: bad argument");
  CK(cudaSetDevice(h->cfg.device));
  rc = finish_pending(h); if (rc) return rc;
  const FoldBuffers& b = h->fb[fold];
  const int K = h->cfg.n_classes, D = h->shapes[fold].D, nd = h->cfg.noise_dim;
  const bool gan = h->cfg.model == MRGAN_MODEL_GAN;
  const float* src = nullptr; int ld = 0;
  if (which >= 0 && which <= 5) { src = b.a[which]; ld = b.lda[which]; }
  else if (which >= 11 && which <= 15) { src = b.hb[which - 10]; ld = b.lda[which - 10]; }
  else if (which >= 21 && which <= 25) { src = b.dz[which - 20]; ld = b.ldz[which - 20]; }
  else if (which == 30) { src = b.lg; ld = pitch8(K); }
  else if (which == 31) { src = b.dlg; ld = pitch8(K); }
  else if (which == 50) { src = b.xtr; ld = pitch8(D); }
  else if (which == 51) { src = b.xte; ld = b.lda[0]; }
  else if (gan && which == 32) { src = b.dfake; ld = pitch8(D); }
  else if (gan && which == 40) { src = b.zb; ld = pitch8(nd + 1); }
  else if (gan && which == 41) { src = b.h1g; ld = pitch8(kGH); }
  else if (gan && which == 42) { src = b.u; ld = pitch8(kGH + 1); }
  else if (gan && which == 43) { src = b.h2g; ld = pitch8(kGH + 1); }
  else if (gan && which == 44) { src = b.dz2g; ld = pitch8(kGH); }
  else if (gan && which == 45) { src = b.du; ld = pitch8(kGH); }
  else if (gan && which == 46) { src = b.dz1g; ld = pitch8(kGH); }
  if (!src || cols > ld) return fail(h, MRGAN_ERR_ARG, "debug_buffer: unknown buffer or too many columns");
  float* tmp = nullptr;
  if (h->om.mode == 2) {
    // f16 mode: noisy activations / inputs and every dZ exist only as fp16 operand copies (dZ times the loss scale)
    const bool act_only = (which >= 0 && which <= 4) || which == 40 || which == 42;
    const bool grad_only = (which >= 21 && which <= 25) || which == 31 || which == 32 || which == 44 || which == 46;
    if (act_only || grad_only) {
      CK(cudaMalloc(&tmp, (size_t)rows * ld * sizeof(float)));
      k_from_half<<<256, 256, 0, h->stream>>>(src, tmp, (size_t)rows * ld, grad_only ? 1.0f / h->om.gscale : 1.0f, h->om);
      src = tmp;
    } else if (which == 45) {   // du stays fp32 but carries the loss scale
      CK(cudaMalloc(&tmp, (size_t)rows * ld * sizeof(float)));
      CK(cudaMemcpyAsync(tmp, src, (size_t)rows * ld * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
      k_scale_buf<<<256, 256, 0, h->stream>>>(tmp, (size_t)rows * ld, 1.0f / h->om.gscale);
      src = tmp;
    }
  }
  cudaError_t ce = cudaMemcpy2DAsync(dst, (size_t)cols * sizeof(float), src, (size_t)ld * sizeof(float), (size_t)cols * sizeof(float), rows,
                                     cudaMemcpyDeviceToHost, h->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
  if (tmp) cudaFree(tmp);
  if (ce != cudaSuccess) return fail(h, MRGAN_ERR_CUDA, std::string("debug_buffer: ") + cudaGetErrorString(ce));
  return MRGAN_OK;
}

int mrgan_debug_gemm(mrgan_handle* h, int mode, int M, int N, int K, const float* A, const float* B, float* C, int use_tc) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (mode < 0 || mode > 2 || M < 1 || N < 1 || K < 1 || !A || !B || !C) return fail(h, MRGAN_ERR_ARG, "debug_gemm: bad argument");
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  // operand shapes (rows, cols) as stored
  const int ar = mode == 2 ? K : M, ac = mode == 2 ? M : K;
  const int br = mode == 1 ? N : K, bc = mode == 1 ? K : N;
  const int lda = pitch8(ac), ldb = pitch8(bc), ldc = pitch8(N);
  float *dA = nullptr, *dB = nullptr, *dC = nullptr; GemmDesc* dd = nullptr;
  CK(cudaMalloc(&dA, (size_t)ar * lda * 4)); CK(cudaMalloc(&dB, (size_t)br * ldb * 4)); CK(cudaMalloc(&dC, (size_t)M * ldc * 4));
  CK(cudaMalloc(&dd, sizeof(GemmDesc)));
  CK(cudaMemsetAsync(dA, 0, (size_t)ar * lda * 4, h->stream)); CK(cudaMemsetAsync(dB, 0, (size_t)br * ldb * 4, h->stream));
  CK(cudaMemsetAsync(dC, 0, (size_t)M * ldc * 4, h->stream));
  CK(cudaMemcpy2DAsync(dA, (size_t)lda * 4, A, (size_t)ac * 4, (size_t)ac * 4, ar, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpy2DAsync(dB, (size_t)ldb * 4, B, (size_t)bc * 4, (size_t)bc * 4, br, cudaMemcpyHostToDevice, h->stream));
  GemmDesc g = make_desc(dA, lda, dB, ldb, dC, ldc, M, N, K, mode == 0 ? EPI_FWD : (mode == 1 ? EPI_DX : EPI_STORE), ACT_NONE, 0);
  CK(cudaMemcpyAsync(dd, &g, sizeof(g), cudaMemcpyHostToDevice, h->stream));
  bool done = false;
#ifdef MRGAN_WITH_TC
  __half *hA = nullptr, *hB = nullptr;
  if (use_tc == 2) {          // fp16 operand copies (kind::f16): host conversion, pitch rounded to 8 elements (16-byte TMA strides)
    const int lha = (ac + 7) & ~7, lhb = (bc + 7) & ~7;
    std::vector<__half> ha((size_t)ar * lha, __float2half(0.f)), hb((size_t)br * lhb, __float2half(0.f));
    for (int r = 0; r < ar; ++r) for (int c2 = 0; c2 < ac; ++c2) ha[(size_t)r * lha + c2] = __float2half_rn(A[(size_t)r * ac + c2]);
    for (int r = 0; r < br; ++r) for (int c2 = 0; c2 < bc; ++c2) hb[(size_t)r * lhb + c2] = __float2half_rn(B[(size_t)r * bc + c2]);
    CK(cudaMalloc(&hA, ha.size() * sizeof(__half))); CK(cudaMalloc(&hB, hb.size() * sizeof(__half)));
    CK(cudaMemcpy(hA, ha.data(), ha.size() * sizeof(__half), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(hB, hb.data(), hb.size() * sizeof(__half), cudaMemcpyHostToDevice));
    GemmDesc gh = g;
    gh.A = reinterpret_cast<const float*>(hA); gh.lda = lha; gh.B = reinterpret_cast<const float*>(hB); gh.ldb = lhb;
    rc = tc_debug_gemm(h, mode, gh, 2);
    cudaFree(hA); cudaFree(hB);
    if (rc) return rc;
    done = true;
  } else if (use_tc) {
    rc = tc_debug_gemm(h, mode, g, 4);
    if (rc) return rc;
    done = true;
  }
#endif
  if (use_tc && !done) return fail(h, MRGAN_ERR_ARG, "library built without the tcgen05 kernels");
  if (!done) {
    dim3 grid((N + 63) / 64, (M + 63) / 64, 1);
    if (mode == 0) k_gemm_simt<false, false><<<grid, 256, 0, h->stream>>>(dd, h->d_folds, 0, h->hp);
    else if (mode == 1) k_gemm_simt<false, true><<<grid, 256, 0, h->stream>>>(dd, h->d_folds, 0, h->hp);
    else k_gemm_simt<true, false><<<grid, 256, 0, h->stream>>>(dd, h->d_folds, 0, h->hp);
    h->launches++;
  }
  cudaError_t e = cudaMemcpy2DAsync(C, (size_t)N * 4, dC, (size_t)ldc * 4, (size_t)N * 4, M, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dd);
  if (e != cudaSuccess) return fail(h, MRGAN_ERR_CUDA, std::string("debug_gemm: ") + cudaGetErrorString(e));
  return MRGAN_OK;
}

int mrgan_debug_gemm_time(mrgan_handle* h, int mode, int M, int N, int K, int groups, int reps, float* ms) {
  if (!h) return fail(nullptr, MRGAN_ERR_ARG, "null handle");
  if (mode < 0 || mode > 2 || M < 1 || N < 1 || K < 1 || groups < 1 || reps < 1 || !ms) return fail(h, MRGAN_ERR_ARG, "debug_gemm_time: bad argument");
#ifndef MRGAN_WITH_TC
  return fail(h, MRGAN_ERR_ARG, "library built without the tcgen05 kernels");
#else
  CK(cudaSetDevice(h->cfg.device));
  int rc = finish_pending(h); if (rc) return rc;
  const int ar = mode == 2 ? K : M, ac = mode == 2 ? M : K;
  const int br = mode == 1 ? N : K, bc = mode == 1 ? K : N;
  const int lda = pitch8(ac), ldb = pitch8(bc), ldc = pitch8(N);
  const size_t sa = (size_t)ar * lda, sb = (size_t)br * ldb, sc = (size_t)M * ldc;
  float* buf = nullptr;
  CK(cudaMalloc(&buf, (sa + sb + sc) * groups * 4));
  CK(cudaMemsetAsync(buf, 0, (sa + sb + sc) * groups * 4, h->stream));
  EncodeTiledFn fn = tc_encoder();
  if (!fn) return fail(h, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  std::vector<TcOp> ops(groups);
  memset(ops.data(), 0, sizeof(TcOp) * groups);
  for (int g = 0; g < groups; ++g) {
    float* A = buf + (sa + sb + sc) * g; float* B = A + sa; float* C = B + sb;
    GemmDesc d = make_desc(A, lda, B, ldb, C, ldc, M, N, K, mode == 0 ? EPI_FWD : (mode == 1 ? EPI_DX : EPI_STORE), ACT_NONE, 0);
    if (!tc_fill_op(fn, ops[g], d, mode)) return fail(h, MRGAN_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  }
  TcOp* dops = nullptr;
  CK(cudaMalloc(&dops, sizeof(TcOp) * groups));
  CK(cudaMemcpyAsync(dops, ops.data(), sizeof(TcOp) * groups, cudaMemcpyHostToDevice, h->stream));
  tc_set_smem_attr();
  const TcOp& t = ops[0];
  dim3 grid((t.ME + 127) / 128, (t.NE + t.bn - 1) / t.bn, groups);
  auto once = [&]() {
    const OperandMode om0{0, 1.0f, nullptr, nullptr};
    if (mode == 0) K_TC_FWD(false, false)<<<grid, TC_FWD_THREADS, tc_smem_bytes(t.bn, TC_FWD_STAGES), h->stream>>>(dops, h->d_folds, 0, h->hp, om0);
    else if (mode == 1) K_TC_DX(false, false)<<<grid, TC_FWD_THREADS, tc_smem_bytes(t.bn, TC_FWD_STAGES), h->stream>>>(dops, h->d_folds, 0, h->hp, om0);
    else K_TC_DW(false, false)<<<grid, TC_DW_THREADS, tc_smem_bytes(t.bn, TC_DW_STAGES), h->stream>>>(dops, h->d_folds, 0, h->hp, om0);
  };
  once();
  CK(cudaEventRecord(h->ev0, h->stream));
  for (int i = 0; i < reps; ++i) once();
  CK(cudaEventRecord(h->ev1, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CKS();
  float t_ms = 0.f;
  CK(cudaEventElapsedTime(&t_ms, h->ev0, h->ev1));
  *ms = t_ms / reps;
  h->launches += reps + 1;
  cudaFree(buf); cudaFree(dops);
  return MRGAN_OK;
#endif
}

int64_t mrgan_kernel_launches(const mrgan_handle* h) { return h ? h->launches : -1; }
double mrgan_last_device_ms(const mrgan_handle* h) { return h ? h->last_ms : -1.0; }

}  // extern "C"
