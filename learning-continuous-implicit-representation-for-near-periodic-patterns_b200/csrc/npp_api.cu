// Host side of libnpp_b200.so: the plan (layer graph, arena layout, workspace, TMA descriptors)
// and the extern "C" entry points declared in include/npp_b200.h.
#include "../../include/npp_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "gemm_sm100.cuh"
#include "simt_kernels.cuh"

using namespace npp;

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const std::string& m) {
  g_err = m;
  return 1;
}
#define CK(expr)                                                                                         \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fail(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" +   \
                  std::to_string(__LINE__) + ")");                                                       \
  } while (0)
#define CKI(expr)            \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

// ------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode_fn() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr || q != cudaDriverEntryPointSuccess)
    return fail("cuTensorMapEncodeTiled is not available from the CUDA driver (need a Hopper/Blackwell driver)");
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

// fp16 row-major [rows, cols] tensor with pitch `ld` elements; box = {64 cols, box_rows}, 128B swizzle.
static int make_map(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
  CKI(load_encode_fn());
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (rows=" +
                std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld) + ")");
  return 0;
}

// ------------------------------------------------------------------------------- plan
struct Buf {
  std::string name;
  int width = 0;  // elements per row == pitch
  __half* ptr = nullptr;
};

struct Seg {
  int buf;        // activation buffer feeding this K segment
  int ref_lo;     // first reference input column
  int ref_w;      // reference columns
  int pad_off;    // first padded column inside the shadow weight
  int pad_w;      // padded width (multiple of 64) == buffer width
  int producer;   // layer whose output this is, -1 for the encoding
  int wt_row0;    // first row inside the transposed shadow, -1 if no dgrad flows here
};

struct Layer {
  std::string name;  // reference module path
  int out = 0, in_ref = 0, kpad = 0;   // out: out_features as the GEMMs see them (out_ref rounded up to the 256 tile)
  int out_ref = 0;                     // reference out_features (arena shape); padded units have zero weights
  std::vector<Seg> segs;
  int act = 0;  // 1 = snake
  int buf_h = -1, buf_d = -1, buf_delta = -1;
  long long w_off = 0, b_off = 0;  // arena
  long long pg_off = 0, bg_off = 0;
  __half* wf = nullptr;
  __half* wt = nullptr;
  int wt_rows = 0;
  CUtensorMap map_wf, map_wt;
};

struct DgradOp {
  int producer;  // layer whose delta is produced
  struct Src {
    int layer, seg;
  } src[2];
  int nsrc;
};

enum { PROF_ENCODE = 0, PROF_GEMM_FWD, PROF_HEAD_LOSS, PROF_GEMM_DGRAD, PROF_GEMM_WGRAD, PROF_FINALIZE, PROF_ADAM, NPP_PROF_CLASSES };

constexpr long long WG_ROWS_PER_SPLIT = 32768;  // rows one split-K accumulator of the wgrad kernel may see

struct NppPlan {
  NppConfig cfg;
  EncTable enc;
  int E = 0, Ep = 0, A = 0, Ap = 0;  // encoding widths (top-1 / aux), reference and padded
  std::vector<Buf> bufs;
  std::vector<Layer> layers;
  std::vector<DgradOp> dgrads;
  std::vector<NppTensorInfo> tensors;
  long long arena_total = 0, arena_trained = 0;
  long long rgb_w_off = 0, rgb_b_off = 0;
  int buf_enc1 = -1, buf_enca = -1;
  int buf_enc1_alt = -1, buf_enca_alt = -1;   // second encoding set: filled by npp_encode_prefetch while a step runs
  int enc_set = 0;                            // encoding set the last forward used (its backward reads it again)
  struct PrefRec {                            // an encoding written ahead of its step by npp_encode_prefetch
    const float* coords = nullptr;            //   identified by the coords pointer + row count of the step that uses it
    long long n = 0;
    bool valid = false;
    cudaEvent_t done = nullptr;
    int tag = -1;                             // unrolled iteration it was encoded for (npp_multi_fit_run), else -1
  } pref[2];                                  // one per encoding set
  int iter_tag = -1;                          // unrolled iteration being captured (matches PrefRec::tag), else -1
  int enc_step_off = 0;                       // re-launched step graph: the encode kernel reads batch *d_step + enc_step_off
  cudaStream_t enc_stream2 = nullptr;         // capture branch of the encodings pipelined inside npp_multi_fit_run
  cudaEvent_t enc_fork = nullptr;
  cudaStream_t side_stream = nullptr;
  cudaEvent_t pref_fork = nullptr;
  int head_width = 0;  // W/2

  // device memory owned by the plan
  void* workspace = nullptr;
  void* shadow_mem = nullptr;
  float* partial = nullptr;
  long long slab_stride = 0;
  float* acc = nullptr;  // [bias accumulators | head accumulators | amax | loss scratch]
  long long acc_floats = 0, headacc_off = 0, amax_off = 0;
  long long acc_zero_floats = 0;   // [0, acc_zero_floats) are the per-step accumulators; behind them: state that survives a step
  long long ring_off = 0;          // max|g| ring of the fused step (3 slots) and its loss accumulator (slot 3)
  long long step_seq = 0;          // fused steps run so far (selects the ring slots)
  bool acc_clean = true;           // the per-step accumulators are all zero (the fused step's update kernel leaves them so)
  int pdl = 1;                     // programmatic dependent launch between the kernels of a fused step
  int pdl_edges = 7;               // which launches carry the attribute: 1 chain (after update), 2 wgrad (after chain), 4 update (after wgrad)
  bool fused_step = true;          // npp_train_step runs forward + head + backward as one chain (NPP_SPLIT_STEP=1: r01 path)
  float* g_buf = nullptr;       // [max_rows,3] grad wrt logits (fused path)
  unsigned int* d_barrier = nullptr;  // {arrivals, generation} of the fused head kernel's grid barrier
  int head_fused_blocks = 0;          // co-resident blocks of npp_head_fused_kernel (0: not usable)
  float* logits_buf = nullptr;  // [max_rows,3] (fused path)
  FinalizeLayer* d_fin = nullptr;
  ShadowLayer* d_shadow = nullptr;
  UpdateLayer* d_update = nullptr;
  UpdateTable update_table;      // the same table by value (kernel parameter of the fused update)
  KmajorParams* d_fwd_ops = nullptr;    // forward chain (one op per dense layer)
  KmajorParams* d_fwd_ops_alt = nullptr;  // same, reading the second encoding set
  KmajorParams* d_dgrad_ops = nullptr;  // dgrad chain
  KmajorParams* d_step_ops = nullptr;      // fused train step: forward ops (last one with the head epilogue) + dgrad ops
  KmajorParams* d_step_ops_alt = nullptr;  // same, reading the second encoding set
  WgUnit* d_units = nullptr;
  WgUnit* d_tile_units = nullptr;        // tile_units on the device: the unit table of a launch without row splits
  int splits_override = 0;               // > 0: row splits of the next prepare() instead of splits_fill (npp_multi_fit_run)
  WgUnit* d_units_bal = nullptr;         // balanced schedule of the current row count (see prepare), 4 * tiles slots
  bool wg_balanced = false;
  int n_units = 0;
  std::vector<WgUnit> tile_units;        // one entry per (layer, m-tile, n-tile) with split 0: template of every unit table
  std::vector<int> tile_begin;           // first entry of layer i inside tile_units (size layers + 1)
  struct GroupTable {                    // unit table of one layer range for npp_step_wgrad (built on first use)
    int lb, le, splits;
    WgUnit* d_units;
    int n_units;
  };
  std::vector<GroupTable> groups;
  int slabs_alloc = 0;                   // split-K slabs allocated (>= splits_max: layer-range launches split further)
  int splits_max = 0;       // slabs allocated
  int splits_fill = 1;      // splits that give every CTA pair a unit (or the caller's explicit choice)
  bool splits_auto = false;
  int num_sms = 148;
  int cluster = 2;   // 2: the chain kernel runs on CTA pairs with cta_group::2 UMMAs; 1: single-CTA UMMAs
  int wg_cluster = 2;  // same choice for the weight-gradient kernel (256 x 256 tile per CTA pair)

  int* d_step = nullptr;                // step index read by the kernels of npp_fit_run's re-launched step graph
  AdamScalars* d_ad_table = nullptr;    // Adam scalars per step of the current npp_fit_run
  long long ad_table_cap = 0;
  bool step_mode = false;               // launches take their batch / scalars through d_step
  cudaEvent_t tables_evt = nullptr;     // recorded after the last kernels that read the device op tables
  cudaEvent_t coop_evt = nullptr;       // recorded after this plan's last cooperative head launch (see coop_head_allowed)
  cudaEvent_t step_evt = nullptr;       // recorded after this plan's last training launch (see PairStepScope)
  bool step_pending = false;            // step_evt has been recorded and not yet been seen complete (g_pair_mu held)
  bool capturing = false;               // npp_fit_run is recording into side_stream
  cudaGraphExec_t fit_exec = nullptr;   // the last npp_fit_run, captured as one graph (kept until the next run / destroy)
  cudaStream_t fit_stream = nullptr;    // stream it was launched on

  bool keep_grads = false;  // fused train step also writes the gradient arena (tests)
  // bound arenas
  float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr;

  // per-row-count state
  long long prepared_n = -1;
  std::vector<CUtensorMap> map_a;   // per buffer, box {64,128}: K-major A operand
  std::vector<CUtensorMap> map_mn;  // per buffer, box {64,64}: MN-major wgrad operand
  std::vector<CUtensorMap> map_ep;  // per buffer, box {64,32}: per-warp epilogue store / load
  std::vector<KmajorParams> fwd_params, fwd_params_alt, dgrad_params, step_params, step_params_alt;
  WgradParams wg_params, wg_params_alt;
  std::vector<int> wg_src_bufs;  // buffer ids behind WgradParams::maps[nl + k]
  int launches = 0;
  int fwd_subs = 0, dgrad_subs = 0;  // output sub-tiles per stripe of each chain

  // optional per-kernel-class timing (CUDA events on the launching stream)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  size_t ev_used = 0;
  double cls_ms[NPP_PROF_CLASSES] = {0};
  long long cls_launches[NPP_PROF_CLASSES] = {0};
};

struct ProfScope {
  NppPlan* p; cudaStream_t st; int cls; int n; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(NppPlan* p_, cudaStream_t st_, int cls_, int n_launches) : p(p_), st(st_), cls(cls_), n(n_launches) {
    if (!p->profiling) return;
    while (p->ev_pool.size() < p->ev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); p->ev_pool.push_back(e); }
    a = p->ev_pool[p->ev_used++]; b = p->ev_pool[p->ev_used++];
    cudaEventRecord(a, st);
  }
  ~ProfScope() {
    if (!a) return;
    cudaEventRecord(b, st);
    NppPlan::Span s; s.cls = cls; s.a = a; s.b = b;
    p->spans.push_back(s);
    p->cls_launches[cls] += n;
  }
};

static int add_buf(NppPlan* p, const std::string& name, int width) {
  Buf b;
  b.name = name;
  b.width = width;
  p->bufs.push_back(b);
  return (int)p->bufs.size() - 1;
}

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

static void add_tensor(NppPlan* p, const std::string& name, int rows, int cols, int is_bias, int trained,
                       long long* off_out) {
  NppTensorInfo t;
  memset(&t, 0, sizeof(t));
  snprintf(t.name, sizeof(t.name), "%s", name.c_str());
  t.offset = p->arena_total;
  t.rows = rows;
  t.cols = cols;
  t.is_bias = is_bias;
  t.trained = trained;
  p->tensors.push_back(t);
  if (off_out) *off_out = t.offset;
  long long cnt = (long long)rows * cols;
  p->arena_total += (cnt + 3) / 4 * 4;
}

// Adds a dense layer reading `srcs` (buffer, reference width, producer layer) in reference order.
static int add_layer(NppPlan* p, const std::string& name, int out, int act,
                     const std::vector<std::pair<int, std::pair<int, int>>>& srcs) {
  Layer L;
  L.name = name;
  L.out_ref = out;
  out = round_up(out, BN);   // pos_linears.0 of NPP_Net_light at W = 256 has 128 units: the GEMMs run on a zero-padded tile
  L.out = out;
  L.act = act;
  int ref = 0, pad = 0;
  for (auto& s : srcs) {
    Seg g;
    g.buf = s.first;
    g.ref_lo = ref;
    g.ref_w = s.second.first;
    g.pad_off = pad;
    g.pad_w = p->bufs[g.buf].width;
    g.producer = s.second.second;
    g.wt_row0 = -1;
    if (g.producer >= 0) {
      g.wt_row0 = L.wt_rows;
      L.wt_rows += g.ref_w;
    }
    ref += g.ref_w;
    pad += g.pad_w;
    L.segs.push_back(g);
  }
  L.in_ref = ref;
  L.kpad = pad;
  const int idx = (int)p->layers.size();
  const std::string tag = std::to_string(idx);
  L.buf_h = add_buf(p, "h" + tag, out);
  if (act) L.buf_d = add_buf(p, "d" + tag, out);
  L.buf_delta = add_buf(p, "delta" + tag, out);
  p->layers.push_back(L);
  return idx;
}

static int build_graph(NppPlan* p) {
  const NppConfig& c = p->cfg;
  const int W = c.width, D = c.depth;
  const int B = 2 * (c.include_input + 2 * c.n_aug);
  const int F = 1 + 2 * c.n_freq;
  p->E = B * F;
  p->Ep = round_up(p->E, 256);
  p->A = p->E * (c.topk - 1);
  p->Ap = round_up(p->A, 64);
  if (c.model == NPP_MODEL_LIGHT) {
    // search mode: the periodic encoding has no Fourier expansion (4*n_aug columns) and the second operand buffer
    // holds the 2-D positional encoding that pos_linears.0 concatenates after feature1 (networks.py:249)
    p->E = B;
    p->Ep = round_up(p->E, 64);
    p->A = 2 + 4 * c.n_freq;
    p->Ap = round_up(p->A, 64);
  }
  const bool second = c.model != NPP_MODEL_TOP1;
  p->head_width = W / 2;
  p->buf_enc1 = add_buf(p, "enc1", p->Ep);
  if (second) p->buf_enca = add_buf(p, "enc_aux", p->Ap);
  p->buf_enc1_alt = add_buf(p, "enc1_alt", p->Ep);
  if (second) p->buf_enca_alt = add_buf(p, "enc_aux_alt", p->Ap);

  typedef std::pair<int, std::pair<int, int>> S;
  auto src = [](int buf, int w, int prod) { return S(buf, std::make_pair(w, prod)); };
  // periodic_linears (networks.py:42-43, loop at :63-71)
  int prev = -1;
  for (int i = 0; i < D; ++i) {
    std::vector<S> in;
    if (i == 0) {
      in.push_back(src(p->buf_enc1, p->E, -1));
    } else if (i - 1 == c.skip_layer) {
      in.push_back(src(p->buf_enc1, p->E, -1));  // torch.cat([input_periodic, h], -1)
      in.push_back(src(p->layers[prev].buf_h, W, prev));
    } else {
      in.push_back(src(p->layers[prev].buf_h, W, prev));
    }
    prev = add_layer(p, "periodic_linears." + std::to_string(i), W, 1, in);
  }
  const int f1 = add_layer(p, "feature_linear1", W, 0, {src(p->layers[prev].buf_h, W, prev)});  // :73
  int last;
  if (c.model == NPP_MODEL_TOPK) {
    const int s0 = add_layer(p, "scale_linears.0", W, 1,
                             {src(p->layers[f1].buf_h, W, f1), src(p->buf_enca, p->A, -1)});        // :76-82
    const int f2 = add_layer(p, "feature_linear2", W, 0, {src(p->layers[s0].buf_h, W, s0)});        // :84
    last = add_layer(p, "pos_linears.0", W / 2, 1,
                     {src(p->layers[f1].buf_h, W, f1), src(p->layers[f2].buf_h, W, f2)});           // :85-92
  } else if (c.model == NPP_MODEL_LIGHT) {
    last = add_layer(p, "pos_linears.0", W / 2, 1,
                     {src(p->layers[f1].buf_h, W, f1), src(p->buf_enca, p->A, -1)});                // :249-258
  } else {
    last = add_layer(p, "pos_linears.0", W / 2, 1, {src(p->layers[f1].buf_h, W, f1)});              // :161-170
  }
  (void)last;

  // arena: trained tensors first, in reference module order within that group
  for (auto& L : p->layers) {
    add_tensor(p, L.name + ".weight", L.out_ref, L.in_ref, 0, 1, &L.w_off);
    add_tensor(p, L.name + ".bias", 1, L.out_ref, 1, 1, &L.b_off);
  }
  add_tensor(p, "rgb_linear.weight", 3, W / 2, 0, 1, &p->rgb_w_off);
  add_tensor(p, "rgb_linear.bias", 1, 3, 1, 1, &p->rgb_b_off);
  p->arena_trained = p->arena_total;
  if (c.model == NPP_MODEL_TOP1) {  // allocated but unused by NPP_Net_top1.forward (networks.py:135)
    add_tensor(p, "feature_linear2.weight", W, W, 0, 0, nullptr);
    add_tensor(p, "feature_linear2.bias", 1, W, 1, 0, nullptr);
  }
  if (c.model == NPP_MODEL_LIGHT) {  // the scale branch is skipped when len(freq_scales) == 1 (networks.py:236-250)
    add_tensor(p, "scale_linears.0.weight", W, W, 0, 0, nullptr);   // scale_dim == 0: Linear(W, W), networks.py:204
    add_tensor(p, "scale_linears.0.bias", 1, W, 1, 0, nullptr);
    add_tensor(p, "feature_linear2.weight", W, W, 0, 0, nullptr);
    add_tensor(p, "feature_linear2.bias", 1, W, 1, 0, nullptr);
  }
  add_tensor(p, "alpha_linear.weight", 1, W, 0, 0, nullptr);  // networks.py:48, never used
  add_tensor(p, "alpha_linear.bias", 1, 1, 1, 0, nullptr);

  // backward schedule: producers in reverse order; the last layer's delta comes from the head.
  const int nl = (int)p->layers.size();
  for (int P = nl - 2; P >= 0; --P) {
    DgradOp op;
    op.producer = P;
    op.nsrc = 0;
    for (int cidx = P + 1; cidx < nl; ++cidx)
      for (int s = 0; s < (int)p->layers[cidx].segs.size(); ++s)
        if (p->layers[cidx].segs[s].producer == P) {
          if (op.nsrc == 2) return fail("layer graph: more than two consumers of one activation");
          op.src[op.nsrc].layer = cidx;
          op.src[op.nsrc].seg = s;
          ++op.nsrc;
        }
    if (op.nsrc == 0) return fail("layer graph: activation without consumer");
    p->dgrads.push_back(op);
  }
  return 0;
}

static int alloc_plan_memory(NppPlan* p) {
  const long long R = p->cfg.max_rows;
  // activation workspace
  size_t bytes = 0;
  std::vector<size_t> offs;
  for (auto& b : p->bufs) {
    offs.push_back(bytes);
    bytes += ((size_t)R * b.width * 2 + 1023) / 1024 * 1024;
  }
  CK(cudaMalloc(&p->workspace, bytes));
  CK(cudaMemset(p->workspace, 0, bytes));  // encoding pad columns stay zero forever
  for (size_t i = 0; i < p->bufs.size(); ++i) p->bufs[i].ptr = reinterpret_cast<__half*>((char*)p->workspace + offs[i]);

  // shadow weights
  size_t sbytes = 0;
  std::vector<size_t> wf_off, wt_off;
  for (auto& L : p->layers) {
    wf_off.push_back(sbytes);
    sbytes += ((size_t)L.out * L.kpad * 2 + 1023) / 1024 * 1024;
    wt_off.push_back(sbytes);
    sbytes += ((size_t)L.wt_rows * L.out * 2 + 1023) / 1024 * 1024;
  }
  CK(cudaMalloc(&p->shadow_mem, sbytes));
  CK(cudaMemset(p->shadow_mem, 0, sbytes));  // K padding columns stay zero forever
  long long pg = 0, bg = 0;
  for (size_t i = 0; i < p->layers.size(); ++i) {
    Layer& L = p->layers[i];
    L.wf = reinterpret_cast<__half*>((char*)p->shadow_mem + wf_off[i]);
    L.wt = L.wt_rows ? reinterpret_cast<__half*>((char*)p->shadow_mem + wt_off[i]) : nullptr;
    CKI(make_map(&L.map_wf, L.wf, L.out, L.kpad, L.kpad, BN / p->cluster));
    if (L.wt) CKI(make_map(&L.map_wt, L.wt, L.wt_rows, L.out, L.out, BN / p->cluster));
    L.pg_off = pg;
    pg += (long long)L.out * L.kpad;
    L.bg_off = bg;
    bg += L.out;
  }
  p->slab_stride = pg;

  // split-K factor of the grouped weight-gradient kernel
  int tiles = 0;
  const int wg_bm = BM * p->wg_cluster;   // out-features per work unit (a CTA pair covers 256)
  for (auto& L : p->layers) tiles += (L.out / wg_bm) * ((L.kpad + BN - 1) / BN);
  int S = p->cfg.wgrad_splits;
  if (const char* e = getenv("NPP_WG_SPLITS")) S = atoi(e);   // tuning experiments
  if (S <= 0) {
    // Measured on B200 (16 k rows, 66 tiles on 74 CTA pairs): the kernel runs at the operand-ingest limit of the
    // 2-CTA mainloop whatever the split factor, and every extra split adds a slab of fp32 partials to write here and
    // to read back in the update kernel (S = 1: 0.150 ms for both, S = 6: 0.181 ms).  Split only as far as it takes
    // to give every CTA pair a unit.
    const int slots = p->num_sms / p->wg_cluster;
    S = tiles > 0 ? (slots + tiles / 2) / tiles : 1;
    if (S < 1) S = 1;
    // search-stage fits (7 tiles, 2048 rows): 103 us per step with 4 splits against 107 with the 11 that would fill the
    // GPU, and a quarter of the CTAs, which matters because several candidates are fitted side by side
    if (p->cfg.model == NPP_MODEL_LIGHT && S > 4) S = 4;
    p->splits_auto = true;
  }
  p->splits_fill = S;
  // one fp32 accumulator should not see more than WG_ROWS_PER_SPLIT rows: beyond that the rounding of the running
  // sum shows (2^18 rows in one split: 1.1e-4 relative difference between a batch and its two halves)
  if (p->splits_auto) S = std::max(S, (int)((p->cfg.max_rows + WG_ROWS_PER_SPLIT - 1) / WG_ROWS_PER_SPLIT));
  if (S > NPP_MAX_SPLITS) S = NPP_MAX_SPLITS;
  p->splits_max = S;
  p->slabs_alloc = std::max(S, 4);   // npp_step_wgrad launches a few layers at a time and splits their rows to fill the GPU
  CK(cudaMalloc(&p->partial, (size_t)p->slabs_alloc * p->slab_stride * sizeof(float)));

  // accumulators: [bias acc (bg) | head acc (3*hw+3) | amax | loss]
  p->headacc_off = (bg + 3) / 4 * 4;
  p->amax_off = p->headacc_off + (3 * p->head_width + 3 + 3) / 4 * 4;
  p->acc_zero_floats = p->amax_off + 4;
  p->ring_off = p->acc_zero_floats;
  p->acc_floats = p->ring_off + 4;
  CK(cudaMalloc(&p->acc, p->acc_floats * sizeof(float)));
  CK(cudaMemset(p->acc, 0, p->acc_floats * sizeof(float)));
  CK(cudaMalloc(&p->g_buf, (size_t)R * 3 * sizeof(float)));
  CK(cudaMalloc(&p->d_step, sizeof(int)));
  CK(cudaMemset(p->d_step, 0, sizeof(int)));
  CK(cudaMalloc(&p->d_barrier, 2 * sizeof(unsigned int)));
  CK(cudaMemset(p->d_barrier, 0, 2 * sizeof(unsigned int)));
  {
    int coop = 0, per_sm = 0, dev = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    int per_sm_reg = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, npp_head_fused_kernel, 256, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_reg, npp_head_fused_reg_kernel, 256, 0));
    if (per_sm_reg < per_sm) per_sm = per_sm_reg;
    if (per_sm > 2) per_sm = 2;
    p->head_fused_blocks = coop && p->head_width <= 256 ? per_sm * p->num_sms : 0;
  }
  CK(cudaMalloc(&p->logits_buf, (size_t)R * 3 * sizeof(float)));

  // device tables
  std::vector<FinalizeLayer> fin;
  std::vector<ShadowLayer> sh;
  for (auto& L : p->layers) {
    FinalizeLayer f;
    f.w_off = L.w_off;
    f.b_off = L.b_off;
    f.pg_off = L.pg_off;
    f.bg_off = L.bg_off;
    f.out = L.out_ref;
    f.in_ref = L.in_ref;
    f.kpad = L.kpad;
    f.split_col = L.segs.size() > 1 ? L.segs[1].ref_lo : L.in_ref;
    f.off0 = L.segs[0].pad_off;
    f.off1 = L.segs.size() > 1 ? L.segs[1].pad_off : 0;
    fin.push_back(f);
    ShadowLayer s;
    memset(&s, 0, sizeof(s));
    s.w_off = L.w_off;
    s.out = L.out_ref;
    s.wt_ld = L.out;
    s.in_ref = L.in_ref;
    s.kpad = L.kpad;
    s.split_col = f.split_col;
    s.off0 = f.off0;
    s.off1 = f.off1;
    s.wf = L.wf;
    s.wt = L.wt;
    s.t_lo = s.t_hi = s.t_lo2 = s.t_hi2 = -1;
    int nt = 0;
    for (auto& g : L.segs)
      if (g.wt_row0 >= 0) {
        if (nt == 0) {
          s.t_lo = g.ref_lo;
          s.t_hi = g.ref_lo + g.ref_w;
          s.t_row0 = g.wt_row0;
        } else {
          s.t_lo2 = g.ref_lo;
          s.t_hi2 = g.ref_lo + g.ref_w;
          s.t_row02 = g.wt_row0;
        }
        ++nt;
      }
    sh.push_back(s);
  }
  CK(cudaMalloc(&p->d_fin, fin.size() * sizeof(FinalizeLayer)));
  CK(cudaMemcpy(p->d_fin, fin.data(), fin.size() * sizeof(FinalizeLayer), cudaMemcpyHostToDevice));
  {
    std::vector<UpdateLayer> up;
    for (size_t i = 0; i < fin.size(); ++i) {
      UpdateLayer u;
      u.w_off = fin[i].w_off; u.b_off = fin[i].b_off; u.pg_off = fin[i].pg_off; u.bg_off = fin[i].bg_off;
      u.out = fin[i].out; u.in_ref = fin[i].in_ref; u.kpad = fin[i].kpad;
      u.split_col = fin[i].split_col; u.off0 = fin[i].off0; u.off1 = fin[i].off1;
      u.wf = sh[i].wf; u.wt = sh[i].wt; u.wt_ld = sh[i].wt_ld;
      u.t_lo = sh[i].t_lo; u.t_hi = sh[i].t_hi; u.t_row0 = sh[i].t_row0;
      u.t_lo2 = sh[i].t_lo2; u.t_hi2 = sh[i].t_hi2; u.t_row02 = sh[i].t_row02;
      up.push_back(u);
    }
    if (up.size() > (size_t)NPP_MAX_UPDATE_LAYERS) return fail("too many layers for the fused update table");
    memset(&p->update_table, 0, sizeof(p->update_table));
    p->update_table.n_layers = (int)up.size();
    for (size_t i = 0; i < up.size(); ++i) {
      p->update_table.L[i] = up[i];
      p->update_table.tile_begin[i + 1] = p->update_table.tile_begin[i] + (up[i].out / 32) * ((up[i].kpad + 127) / 128);
    }
    CK(cudaMalloc(&p->d_update, up.size() * sizeof(UpdateLayer)));
    CK(cudaMemcpy(p->d_update, up.data(), up.size() * sizeof(UpdateLayer), cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&p->d_fwd_ops, p->layers.size() * sizeof(KmajorParams)));
  CK(cudaMalloc(&p->d_fwd_ops_alt, p->layers.size() * sizeof(KmajorParams)));
  CK(cudaStreamCreateWithFlags(&p->side_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&p->pref_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&p->tables_evt, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&p->coop_evt, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&p->pref[i].done, cudaEventDisableTiming));
  CK(cudaMalloc(&p->d_dgrad_ops, (p->dgrads.size() + 1) * sizeof(KmajorParams)));
  CK(cudaMalloc(&p->d_step_ops, (p->layers.size() + p->dgrads.size()) * sizeof(KmajorParams)));
  CK(cudaMalloc(&p->d_step_ops_alt, (p->layers.size() + p->dgrads.size()) * sizeof(KmajorParams)));
  CK(cudaMalloc(&p->d_shadow, sh.size() * sizeof(ShadowLayer)));
  CK(cudaMemcpy(p->d_shadow, sh.data(), sh.size() * sizeof(ShadowLayer), cudaMemcpyHostToDevice));

  // wgrad work units: layer-major, then split, then (m, n) tiles so that concurrently running CTAs
  // share the same row range of the same delta/activation tensors in L2.
  // map index convention inside WgradParams::maps: [0, nl) = delta of layer i; nl + k = k-th distinct source buffer
  std::vector<int> src_bufs;
  auto src_index = [&](int buf) {
    for (size_t i = 0; i < src_bufs.size(); ++i)
      if (src_bufs[i] == buf) return (int)i;
    src_bufs.push_back(buf);
    return (int)src_bufs.size() - 1;
  };
  std::vector<WgUnit> units;
  const int nl = (int)p->layers.size();
  p->tile_begin.assign(1, 0);
  for (int li = 0; li < nl; ++li) {
    const Layer& L = p->layers[li];
    for (int s = 0; s < S; ++s)
      for (int m0 = 0; m0 < L.out; m0 += wg_bm)
        for (int n0 = 0; n0 < L.kpad; n0 += BN) {
          const Seg* sg = nullptr;
          for (auto& g : L.segs)
            if (n0 >= g.pad_off && n0 < g.pad_off + g.pad_w) sg = &g;
          if (!sg) return fail("wgrad tiling: column without segment");
          if ((sg->pad_off % BN) != 0) return fail("wgrad tiling: segment boundary not aligned to 256 columns");
          WgUnit u;
          u.a_map = (short)li;
          u.b_map = (short)(nl + src_index(sg->buf));
          u.a_m0 = m0;
          u.b_n0 = n0 - sg->pad_off;
          u.split = s;
          u.out_off = (int)(L.pg_off + (long long)m0 * L.kpad + n0);
          u.ld = L.kpad;
          u.ncols_left = sg->pad_off + sg->pad_w - n0;
          u.bias_off = n0 == 0 ? (int)(L.bg_off + m0) : -1;   // one unit per (layer, split, m-tile) sums the delta columns
          units.push_back(u);
          if (s == 0) p->tile_units.push_back(u);
        }
    p->tile_begin.push_back((int)p->tile_units.size());
  }
  if (nl + (int)src_bufs.size() > WG_MAX_MAPS) return fail("wgrad: too many tensor maps");
  if (p->slab_stride > 0x7fffffffLL) return fail("wgrad: slab too large for 32-bit offsets");
  p->n_units = (int)units.size();
  CK(cudaMalloc(&p->d_units, units.size() * sizeof(WgUnit)));
  CK(cudaMemcpy(p->d_units, units.data(), units.size() * sizeof(WgUnit), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&p->d_tile_units, p->tile_units.size() * sizeof(WgUnit)));
  CK(cudaMemcpy(p->d_tile_units, p->tile_units.data(), p->tile_units.size() * sizeof(WgUnit), cudaMemcpyHostToDevice));
  p->wg_src_bufs = src_bufs;
  return 0;
}

// Completes a chain's op list: stages per tile, and on-chip forwarding between consecutive ops where the next op
// reads this op's whole output (<= 512 columns) as one K segment.  That segment becomes segment 0 of the next op
// (the order of K segments only changes the fp32 summation order).
static void finish_chain_ops(std::vector<KmajorParams>& ops, int cluster) {
  for (auto& k : ops) {
    k.kb_per_tile = 0;
    for (int s = 0; s < k.nseg; ++s) k.kb_per_tile += k.kblocks[s];
    k.fwd_in = k.fwd_out = 0;
  }
  // Needs the 4-stage ring of the pair kernel: the four blocks of a tile must land in four different stages whose
  // previous occupants belong to UMMAs that do not wait for this very epilogue (with 3 stages block 3 would wait
  // for the next op's first UMMA, which waits for this tile's accumulator: a cycle).
  if (cluster != 2 || getenv("NPP_NO_FWD")) return;
  static_assert(PAIR_STAGES == 4, "on-chip forwarding is laid out for a 4-stage ring");
  for (size_t i = 0; i + 1 < ops.size(); ++i) {
    KmajorParams& cur = ops[i];
    KmajorParams& nxt = ops[i + 1];
    if (cur.tiles_n < 1 || cur.tiles_n > 2) continue;
    int sseg = -1;
    for (int s = 0; s < nxt.nseg; ++s)
      if (nxt.a_src[s] == (int)i && nxt.a_k0[s] == 0 && nxt.kblocks[s] == cur.tiles_n * (BN / BK)) sseg = s;
    if (sseg < 0) continue;
    if (sseg == 1) {
      std::swap(nxt.tmA[0], nxt.tmA[1]);
      std::swap(nxt.tmB[0], nxt.tmB[1]);
      std::swap(nxt.kblocks[0], nxt.kblocks[1]);
      std::swap(nxt.a_k0[0], nxt.a_k0[1]);
      std::swap(nxt.b_k0[0], nxt.b_k0[1]);
      std::swap(nxt.b_row0[0], nxt.b_row0[1]);
      std::swap(nxt.a_src[0], nxt.a_src[1]);
    }
    cur.fwd_out = 1;
    nxt.fwd_in = 1;
  }
}

// ------------------------------------------------------------------- per-row-count setup
static int prepare(NppPlan* p, long long n) {
  if (n <= 0 || n > p->cfg.max_rows)
    return fail("row count " + std::to_string(n) + " outside (0, max_rows=" + std::to_string(p->cfg.max_rows) + "]");
  if (p->prepared_n == n) return 0;
  // The op tables on the device are about to be overwritten: kernels of an earlier call (any stream, possibly a
  // non-blocking one that the copies below would not wait for) may still be reading them.  Wait for the plan's own
  // work only (a device-wide synchronisation is also illegal while another thread captures a graph).
  CK(cudaEventSynchronize(p->tables_evt));
  for (int i = 0; i < 2; ++i) CK(cudaEventSynchronize(p->pref[i].done));
  const int nb = (int)p->bufs.size();
  p->map_a.assign(nb, CUtensorMap());
  p->map_mn.assign(nb, CUtensorMap());
  p->map_ep.assign(nb, CUtensorMap());
  for (int i = 0; i < nb; ++i) {
    const Buf& b = p->bufs[i];
    // rows == n exactly: TMA zero-fills rows >= n, which keeps stale workspace rows out of the
    // row-contraction of the weight-gradient GEMM.
    CKI(make_map(&p->map_a[i], b.ptr, n, b.width, b.width, BM));
    CKI(make_map(&p->map_mn[i], b.ptr, n, b.width, b.width, 64));
    CKI(make_map(&p->map_ep[i], b.ptr, n, b.width, b.width, 32));
  }
  const int tiles_m = (int)((n + BM - 1) / BM);
  p->fwd_params.clear();
  int fwd_subs = 0;
  for (auto& L : p->layers) {
    KmajorParams k;
    memset(&k, 0, sizeof(k));
    k.nseg = (int)L.segs.size();
    for (int s = 0; s < k.nseg; ++s) {
      k.tmA[s] = p->map_a[L.segs[s].buf];
      k.tmB[s] = L.map_wf;
      k.kblocks[s] = L.segs[s].pad_w / BK;
      k.a_k0[s] = 0;
      k.b_k0[s] = L.segs[s].pad_off;
      k.b_row0[s] = 0;
      k.a_src[s] = L.segs[s].producer;   // forward op index == layer index
    }
    k.sub_base = fwd_subs;
    fwd_subs += (L.out / BN) * (BN / EPI_COLS);
    k.M = (int)n;
    k.tiles_m = tiles_m;
    k.tiles_n = L.out / BN;
    k.bias = p->params + L.b_off;
    k.epi = L.act ? EPI_SNAKE : EPI_LINEAR;
    k.out0 = p->bufs[L.buf_h].ptr;
    k.ld0 = L.out;
    k.tmOut0 = p->map_ep[L.buf_h];
    if (L.act) {
      k.out1 = p->bufs[L.buf_d].ptr;
      k.ld1 = L.out;
      k.tmOut1 = p->map_ep[L.buf_d];
    }
    p->fwd_params.push_back(k);
  }
  // the same chain reading the second encoding set
  auto alt_buf = [&](int b) { return b == p->buf_enc1 ? p->buf_enc1_alt : (b == p->buf_enca ? p->buf_enca_alt : b); };
  p->fwd_params_alt = p->fwd_params;
  for (size_t li = 0; li < p->layers.size(); ++li)
    for (size_t s = 0; s < p->layers[li].segs.size(); ++s)
      p->fwd_params_alt[li].tmA[s] = p->map_a[alt_buf(p->layers[li].segs[s].buf)];
  p->dgrad_params.clear();
  int dg_subs = 0;
  for (auto& op : p->dgrads) {
    const Layer& P = p->layers[op.producer];
    KmajorParams k;
    memset(&k, 0, sizeof(k));
    k.nseg = op.nsrc;
    for (int s = 0; s < op.nsrc; ++s) {
      const Layer& C = p->layers[op.src[s].layer];
      k.tmA[s] = p->map_a[C.buf_delta];
      k.tmB[s] = C.map_wt;
      k.kblocks[s] = C.out / BK;
      k.a_k0[s] = 0;
      k.b_k0[s] = 0;
      k.b_row0[s] = C.segs[op.src[s].seg].wt_row0;
      // the delta of consumer layer C is written by the dgrad op whose producer == C (none for the last layer,
      // whose delta comes from the head-backward kernel)
      k.a_src[s] = -1;
      for (size_t q = 0; q < p->dgrads.size(); ++q)
        if (p->dgrads[q].producer == op.src[s].layer) k.a_src[s] = (int)q;
    }
    k.sub_base = dg_subs;
    dg_subs += (P.out / BN) * (BN / EPI_COLS);
    k.M = (int)n;
    k.tiles_m = tiles_m;
    k.tiles_n = P.out / BN;
    k.out0 = p->bufs[P.buf_delta].ptr;
    k.ld0 = P.out;
    k.tmOut0 = p->map_ep[P.buf_delta];
    if (P.act) {
      k.mul = p->bufs[P.buf_d].ptr;
      k.ldm = P.out;
      k.tmMul = p->map_ep[P.buf_d];
    }
    k.colsum = nullptr;   // bias gradients are summed by the weight-gradient kernel (WgUnit::bias_off)
    k.epi = P.act ? EPI_DGRAD_MUL : EPI_DGRAD;
    p->dgrad_params.push_back(k);
  }
  p->fwd_subs = fwd_subs;
  p->dgrad_subs = dg_subs;
  // Fused train step: ONE chain per stripe = the forward ops, whose last one carries the RGB head + loss + head
  // backward in its epilogue (EPI_SNAKE_HEAD) and writes the last layer's delta, followed by the dgrad ops.
  {
    const int nlf = (int)p->layers.size();
    const Layer& last = p->layers.back();
    auto build = [&](const std::vector<KmajorParams>& fwd, std::vector<KmajorParams>& out) {
      out = fwd;
      KmajorParams& h = out.back();
      h.epi = EPI_SNAKE_HEAD;
      h.out0 = p->bufs[last.buf_delta].ptr;
      h.ld0 = last.out;
      h.tmOut0 = p->map_ep[last.buf_delta];
      h.out1 = nullptr;
      h.colsum = nullptr;
      for (auto k : p->dgrad_params) {
        for (int sg = 0; sg < k.nseg; ++sg) k.a_src[sg] = k.a_src[sg] < 0 ? nlf - 1 : nlf + k.a_src[sg];
        k.sub_base += fwd_subs;
        out.push_back(k);
      }
    };
    build(p->fwd_params, p->step_params);
    build(p->fwd_params_alt, p->step_params_alt);
  }
  finish_chain_ops(p->fwd_params, p->cluster);
  finish_chain_ops(p->fwd_params_alt, p->cluster);
  finish_chain_ops(p->dgrad_params, p->cluster);
  finish_chain_ops(p->step_params, p->cluster);
  finish_chain_ops(p->step_params_alt, p->cluster);
  CK(cudaMemcpy(p->d_step_ops, p->step_params.data(), p->step_params.size() * sizeof(KmajorParams), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->d_step_ops_alt, p->step_params_alt.data(), p->step_params_alt.size() * sizeof(KmajorParams),
                cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->d_fwd_ops, p->fwd_params.data(), p->fwd_params.size() * sizeof(KmajorParams), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->d_fwd_ops_alt, p->fwd_params_alt.data(), p->fwd_params_alt.size() * sizeof(KmajorParams),
                cudaMemcpyHostToDevice));
  CK(cudaMemcpy(p->d_dgrad_ops, p->dgrad_params.data(), p->dgrad_params.size() * sizeof(KmajorParams),
                cudaMemcpyHostToDevice));
  WgradParams& w = p->wg_params;
  const int nl = (int)p->layers.size();
  for (int i = 0; i < nl; ++i) w.maps[i] = p->map_mn[p->layers[i].buf_delta];
  for (size_t i = 0; i < p->wg_src_bufs.size(); ++i) w.maps[nl + i] = p->map_mn[p->wg_src_bufs[i]];
  w.units = p->d_units;
  w.n_units = p->n_units;
  w.rows = (int)n;
  const int kb_total = (int)((n + BK - 1) / BK);
  int want_splits = p->splits_override > 0 ? p->splits_override : p->splits_fill;
  if (p->splits_auto) want_splits = std::max(want_splits, (int)((n + WG_ROWS_PER_SPLIT - 1) / WG_ROWS_PER_SPLIT));
  if (want_splits > p->splits_max) want_splits = p->splits_max;
  w.kb_per_split = (kb_total + want_splits - 1) / want_splits;
  w.n_splits = (kb_total + w.kb_per_split - 1) / w.kb_per_split;
  w.grid_pairs = 0;
  if (w.n_splits == 1 && p->splits_fill > 1) {   // fewer row splits than the unit table was laid out for: no idle CTAs
    w.units = p->d_tile_units;
    w.n_units = (int)p->tile_units.size();
  }
  // Balanced schedule.  With one unit per tile and more CTA pairs than tiles but fewer than twice as many (NPP_Net K=3:
  // 60 tiles, 74 pairs), the kernel takes the time of one whole tile while 19 % of the SMs idle.  Instead, every g
  // consecutive tiles are contracted by g + 1 pairs, g = ceil(T / (P - T)): pair c of a group takes the first
  // (g - c) / (g + 1) of tile c's rows (partial sums to slab 0), then the last c / (g + 1) of tile c - 1's (slab 1).
  // Every tile is cut exactly once, so the update kernel sums two slabs (S = 2); all pairs walk their "own" tile from
  // row 0 at the same time, which keeps the operand tiles the pairs of one layer share in L2 as before.
  p->wg_balanced = false;
  {
    const int T = (int)p->tile_units.size(), P = p->num_sms / p->wg_cluster;
    // Measured (B200, cfg2, steady state under the power cap, tests/diag_step_time.py): weight-gradient kernel 136 -> 130 us,
    // update 36 -> 39 us (second slab), chain 320 -> 324 us, step 485.3 -> 485.9 us: the SMs it puts to work draw the power
    // the other kernels then lack.  Off by default (NPP_WG_BALANCE=1 switches it on).
    static const bool balance = getenv("NPP_WG_BALANCE") != nullptr && atoi(getenv("NPP_WG_BALANCE")) != 0;
    if (balance && p->wg_cluster == 2 && p->splits_auto && w.n_splits == 1 && T < P && 2 * T > P && p->slabs_alloc >= 2) {
      const int g = (T + (P - T) - 1) / (P - T);
      if (kb_total >= 8 * (g + 1)) {
        WgUnit empty{};
        empty.kb0 = empty.kb1 = -1;
        empty.bias_off = -1;
        std::vector<WgUnit> head, tail;
        for (int t0 = 0; t0 < T; t0 += g) {
          const int r = std::min(g, T - t0);
          const long long total = (long long)r * kb_total;
          for (int c = 0; c <= r; ++c) {
            const long long lo = c * total / (r + 1), hi = (c + 1) * total / (r + 1);
            WgUnit h = empty, t = empty;
            if (c < r) {
              h = p->tile_units[t0 + c];
              h.split = 0;
              h.kb0 = 0;
              h.kb1 = (int)(hi - (long long)c * kb_total);
            }
            if (c >= 1) {
              t = p->tile_units[t0 + c - 1];
              t.split = 1;
              t.kb0 = (int)(lo - (long long)(c - 1) * kb_total);
              t.kb1 = kb_total;
            }
            head.push_back(h);
            tail.push_back(t);
          }
        }
        const int chunks = (int)head.size();
        if (chunks <= P) {
          head.insert(head.end(), tail.begin(), tail.end());
          if (p->d_units_bal == nullptr) CK(cudaMalloc(&p->d_units_bal, (size_t)4 * T * sizeof(WgUnit)));
          CK(cudaMemcpy(p->d_units_bal, head.data(), head.size() * sizeof(WgUnit), cudaMemcpyHostToDevice));
          w.units = p->d_units_bal;
          w.n_units = (int)head.size();
          w.grid_pairs = chunks;      // pair c: unit c (head), then unit c + chunks (tail)
          w.n_splits = 2;             // slabs the update kernel sums
          p->wg_balanced = true;
        }
      }
    }
  }
  w.partial = p->partial;
  w.slab_stride = p->slab_stride;
  w.bias_acc = getenv("NPP_WG_NOBIAS") ? nullptr : p->acc;   // (timing experiments only: the bias gradients are then missing)
  p->wg_params_alt = w;
  for (size_t i = 0; i < p->wg_src_bufs.size(); ++i) p->wg_params_alt.maps[nl + i] = p->map_mn[alt_buf(p->wg_src_bufs[i])];
  p->prepared_n = n;
  p->pref[0].valid = p->pref[1].valid = false;   // prefetched encodings belong to the previous row count
  return 0;
}

static long long* g_chain_dbg = nullptr;  // set by npp_debug_gemm_bench when NPP_DEBUG_STAMPS is given
static int g_smem_attr_done = 0;
static int set_smem_attrs() {
  if (g_smem_attr_done) return 0;
  CK(cudaFuncSetAttribute(npp_gemm_kmajor<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CK(cudaFuncSetAttribute(npp_gemm_kmajor<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
  CK(cudaFuncSetAttribute(npp_gemm_kmajor<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CK(cudaFuncSetAttribute(npp_gemm_kmajor<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
  CK(cudaFuncSetAttribute(npp_gemm_wgrad<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, WGRAD_SMEM_BYTES));
  CK(cudaFuncSetAttribute(npp_gemm_wgrad<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WGRAD_PAIR_LAUNCH_SMEM_BYTES));
  g_smem_attr_done = 1;
  return 0;
}

// Runs ops[0..n_ops) (device array) as one persistent chain: CTA b owns row stripes b, b+grid, ...
// cluster == 2: CTA pairs (thread-block clusters of 2) run cta_group::2 UMMAs, each CTA holding half of every weight tile.
// -DNPP_HANG_DEBUG builds: wait-state buffer of the plan whose entry point is running on this host thread (else nullptr)
static thread_local unsigned long long* tl_dbg_state = nullptr;

static int launch_chain(const KmajorParams* d_ops, const KmajorParams* h_ops, int n_ops, int M, int num_sms,
                        cudaStream_t st, int subs_per_stripe, int cluster = 1, float* zero_a = nullptr, int zero_n = 0,
                        float* zero_b = nullptr, int relu = 0, const HeadArgs* head = nullptr, int pdl = 0) {
  CKI(set_smem_attrs());
  if (n_ops > MAX_CHAIN_OPS) return fail("chain longer than MAX_CHAIN_OPS");
  ChainParams cp;
  memset(&cp, 0, sizeof(cp));
  for (int i = 0; i < n_ops; ++i) {
    const KmajorParams& k = h_ops[i];
    OpScalars& sc = cp.sc[i];
    sc.nseg = k.nseg; sc.tiles_n = k.tiles_n; sc.epi = k.epi;
    sc.fwd_in = k.fwd_in; sc.fwd_out = k.fwd_out; sc.kb_per_tile = k.kb_per_tile;
    for (int s = 0; s < 2; ++s) {
      sc.kblocks[s] = k.kblocks[s]; sc.a_k0[s] = k.a_k0[s]; sc.b_k0[s] = k.b_k0[s];
      sc.b_row0[s] = k.b_row0[s]; sc.a_src[s] = s < k.nseg ? k.a_src[s] : -1;
      const int src = sc.a_src[s];
      sc.src_sub_base[s] = src >= 0 ? h_ops[src].sub_base : 0;
      sc.src_tiles_n[s] = src >= 0 ? h_ops[src].tiles_n : 0;
    }
    sc.ldf = k.ldf; sc.k_adv = k.k_adv; sc.bias = k.bias; sc.colsum = k.colsum; sc.out_f32 = k.out_f32;
    sc.desc_hi = k.desc_hi;
  }
  cp.ops = d_ops;
  cp.zero_a = zero_a;
  cp.zero_n = zero_n;
  cp.zero_b = zero_b;
  cp.relu = relu;
  if (head != nullptr) cp.head = *head;
  cp.pdl = pdl;
  cp.dbg_state = tl_dbg_state;
  cp.n_ops = n_ops;
  cp.M = M;
  cp.tiles_m = (M + BM - 1) / BM;
  cp.subs_per_stripe = subs_per_stripe;
  cp.dbg = g_chain_dbg;
  cp.dbg_block = getenv("NPP_DEBUG_STAMP_BLOCK") ? atoi(getenv("NPP_DEBUG_STAMP_BLOCK")) : 0;
  cp.dbg_warp = getenv("NPP_DEBUG_STAMP_WARP") ? atoi(getenv("NPP_DEBUG_STAMP_WARP")) : 2;
  int grid = cp.tiles_m < num_sms ? cp.tiles_m : num_sms;
  if (cluster == 1) {
    if (relu) npp_gemm_kmajor<1, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(cp);
    else npp_gemm_kmajor<1><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(cp);
    CK(cudaGetLastError());
    return 0;
  }
  grid = (grid + cluster - 1) / cluster * cluster;
  if (grid > num_sms) grid = num_sms / cluster * cluster;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = PAIR_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  // NPP_DEBUG_MODEL_STAMPS=<k>: the k-th chain launch of the process prints the per-tile clock stamps of one CTA
  static int model_stamp_calls = 0;
  const char* ms_env = getenv("NPP_DEBUG_MODEL_STAMPS");
  if (ms_env != nullptr && cp.dbg == nullptr && ++model_stamp_calls == atoi(ms_env)) {
    const int max_tiles = 256;
    long long* d_dbg = nullptr;
    CK(cudaMalloc(&d_dbg, (size_t)max_tiles * 8 * sizeof(long long)));
    CK(cudaMemset(d_dbg, 0, (size_t)max_tiles * 8 * sizeof(long long)));
    cp.dbg = d_dbg;
    CK(cudaLaunchKernelEx(&cfg, npp_gemm_kmajor<2>, cp));
    CK(cudaStreamSynchronize(st));
    std::vector<long long> h((size_t)max_tiles * 8);
    CK(cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    std::vector<KmajorParams> ops((size_t)n_ops);
    CK(cudaMemcpy(ops.data(), d_ops, ops.size() * sizeof(KmajorParams), cudaMemcpyDeviceToHost));
    int tile = 0;
    const long long t0 = h[0];
    long long prev_enter = t0;
    for (int oi = 0; oi < n_ops && tile < max_tiles; ++oi)
      for (int nt = 0; nt < ops[oi].tiles_n && tile < max_tiles; ++nt, ++tile) {
        const long long* q = &h[(size_t)tile * 8];
        printf("  op %2d tile %d (epi %d kb %2d fwd_in %d): enter %7lld (+%6lld) ready %7lld subs %7lld done %7lld released %7lld | "
               "mma free %7lld kb0 %7lld kbN %7lld\n",
               oi, nt, ops[oi].epi, ops[oi].kb_per_tile, ops[oi].fwd_in, q[0] - t0, q[0] - prev_enter, q[1] - t0, q[2] - t0,
               q[3] - t0, q[4] - t0, q[5] - t0, q[6] - t0, q[7] - t0);
        prev_enter = q[0];
      }
    fflush(stdout);
    cudaFree(d_dbg);
    return 0;
  }
  if (relu) CK(cudaLaunchKernelEx(&cfg, npp_gemm_kmajor<2, true>, cp));
  else CK(cudaLaunchKernelEx(&cfg, npp_gemm_kmajor<2>, cp));
  return 0;
}

static int launch_wgrad(const WgradParams& w_in, int num_sms, cudaStream_t st, int cluster, int pdl = 0) {
  CKI(set_smem_attrs());
  WgradParams w = w_in;
  w.dbg_state = tl_dbg_state;
  if (cluster == 1) {
    const int grid = w.n_units < num_sms ? w.n_units : num_sms;
    npp_gemm_wgrad<1><<<grid, WGRAD_THREADS, WGRAD_SMEM_BYTES, st>>>(w);
    CK(cudaGetLastError());
    return 0;
  }
  int grid = 2 * w.n_units < num_sms ? 2 * w.n_units : num_sms / 2 * 2;
  if (w.grid_pairs > 0) grid = 2 * w.grid_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(WGRAD_THREADS);
  // The pair kernel asks for ALL the shared memory of an SM although its ring needs 161.5 KB: a CTA of a pair must not
  // share its SM with blocks of other kernels.  With the 64 KB it used to leave free, update / encode blocks of OTHER
  // plans could sit on one SM of the pair when it was launched, and fits of several plans running side by side
  // dead-locked on the device (3 x 8192-row fits: 10 of 12 runs; with the SM to itself 0 of 14, 8 x 2048 rows 0 of 3, nine
  // NPP_Net_light fits 0 of 4; profiles/r02/concurrent_plans.txt).  The single-CTA instantiation shares its SM with such
  // blocks without harm (0 of 32), and the chain kernel fills the SM anyway.  NPP_WG_SMEM_SHARE=1 restores the old
  // request (to reproduce the dead-lock).
  static const bool smem_share = getenv("NPP_WG_SMEM_SHARE") != nullptr && atoi(getenv("NPP_WG_SMEM_SHARE")) != 0;
  cfg.dynamicSmemBytes = smem_share ? WGRAD_PAIR_SMEM_BYTES : WGRAD_PAIR_LAUNCH_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  CK(cudaLaunchKernelEx(&cfg, npp_gemm_wgrad<2>, w));
  return 0;
}

// Optional serialisation of training launches of plans that use the CTA-pair weight-gradient kernel
// (NPP_PAIR_SERIALISE=1; off by default).  It was the first containment of the dead-lock described in launch_wgrad
// (fits of several full-size plans overlapping on the device), before the cause was narrowed down to other blocks
// sharing an SM with a pair CTA; it is kept as a switch because it needs nothing from the kernels: every training
// entry point of a pair plan holds this mutex while it enqueues, makes its stream wait for the last training launch of
// every other pair plan, and records its own event, so steps of different plans never overlap (3 x 8192 rows: 50 ms
// per search instead of 33 ms; 8 x 2048 rows: 277 ms instead of 89 ms).
static bool pair_serialise() {
  static const bool on = getenv("NPP_PAIR_SERIALISE") != nullptr && atoi(getenv("NPP_PAIR_SERIALISE")) != 0;
  return on;
}
static std::mutex g_pair_mu;
static std::vector<NppPlan*> g_pair_plans;   // live plans with wg_cluster == 2
class PairStepScope {
 public:
  PairStepScope(NppPlan* p, cudaStream_t st) : p_(p), st_(st) {
    on_ = pair_serialise() && p->wg_cluster == 2 && !p->capturing && p->step_evt != nullptr;
    if (on_) g_pair_mu.lock();
  }
  ~PairStepScope() {
    if (on_) g_pair_mu.unlock();
  }
  int begin() {   // this plan's stream waits for the training launches other pair plans have in flight
    if (!on_) return 0;
    for (NppPlan* q : g_pair_plans) {
      if (q == p_ || !q->step_pending) continue;
      if (cudaEventQuery(q->step_evt) == cudaSuccess) {
        q->step_pending = false;
      } else {
        cudaGetLastError();   // cudaErrorNotReady is not sticky, but keep the error state clean
        CK(cudaStreamWaitEvent(st_, q->step_evt, 0));
      }
    }
    return 0;
  }
  int end() {
    if (!on_) return 0;
    CK(cudaEventRecord(p_->step_evt, st_));
    p_->step_pending = true;
    return 0;
  }
 private:
  NppPlan* p_;
  cudaStream_t st_;
  bool on_ = false;
};

// The fused head is a cooperative launch with a grid barrier.  Two such grids in flight at the same time (two plans
// stepping on two streams) can each hold part of the SMs and wait for the rest -- observed as a hang with nine
// concurrent fits.  One cooperative head may be in flight per process: a plan that finds another plan's head still
// running takes the two-kernel head for that step.  Steps of ONE plan are ordered by construction (they share the
// activation workspace), so the common single-plan case never leaves the cooperative path.
static std::mutex g_coop_mu;
static NppPlan* g_coop_plan = nullptr;   // plan that launched the last cooperative head
static bool coop_head_allowed(NppPlan* p) {   // g_coop_mu held
  if (g_coop_plan == nullptr || g_coop_plan == p) return true;
  const cudaError_t q = cudaEventQuery(g_coop_plan->coop_evt);
  if (q == cudaSuccess) return true;
  cudaGetLastError();   // cudaErrorNotReady is not sticky, but keep the error state clean
  return false;
}

// After the last kernel of a call that reads the device op tables (see prepare()).
static int mark_busy(NppPlan* p, cudaStream_t st) {
  if (!p->capturing) CK(cudaEventRecord(p->tables_evt, st));
  return 0;
}

// Encodes `coords` into encoding set `set` (0: enc1 / enc_aux, 1: their alternates).  zero_loss != nullptr: the kernel
// also clears the step accumulators and *zero_loss (first kernel of a fused train step).
static int launch_encode(NppPlan* p, const float* coords, long long n, int set, cudaStream_t st, float* zero_loss) {
  if (p->cfg.model == NPP_MODEL_LIGHT) {
    const int b1 = set ? p->buf_enc1_alt : p->buf_enc1;
    const int bp = set ? p->buf_enca_alt : p->buf_enca;
    const long long items = n * (2 * p->cfg.n_aug + 1 + p->cfg.n_freq);
    unsigned blocks = (unsigned)std::min<long long>((items + 255) / 256, (long long)p->num_sms * 8);
    npp_encode_search_kernel<<<blocks, 256, 0, st>>>(coords, (int)n, p->enc, p->bufs[b1].ptr, p->Ep, p->bufs[bp].ptr,
                                                     p->Ap, zero_loss ? p->acc : nullptr,
                                                     zero_loss ? (int)p->acc_zero_floats : 0, zero_loss,
                                                     p->step_mode ? p->d_step : nullptr, p->enc_step_off);
    CK(cudaGetLastError());
    ++p->launches;
    return 0;
  }
  const int width = p->E;
  const int B = 2 * (p->cfg.include_input + 2 * p->cfg.n_aug);
  const int rows = std::max(1, ENC_THREADS / (B / 2));   // one thread per (row, base-feature pair)
  dim3 grid((unsigned)((n + rows - 1) / rows), p->cfg.topk);
  const size_t smem = (size_t)enc_base_bytes(rows, B) + (size_t)rows * enc_row_stride(width) * sizeof(__half);
  const int b1 = set ? p->buf_enc1_alt : p->buf_enc1;
  const int ba = set ? p->buf_enca_alt : p->buf_enca;
  __half* enca = ba >= 0 ? p->bufs[ba].ptr : nullptr;
  npp_encode_kernel<<<grid, ENC_THREADS, smem, st>>>(coords, (int)n, p->enc, p->bufs[b1].ptr, p->Ep, enca, p->Ap, rows,
                                                     zero_loss ? p->acc : nullptr, zero_loss ? (int)p->acc_zero_floats : 0,
                                                     zero_loss);
  CK(cudaGetLastError());
  ++p->launches;
  return 0;
}

// Makes the encoding of `coords` available in one of the two encoding sets (p->enc_set on return): picks up a
// prefetched one (*prefetched = true) or encodes in line.
static int select_encoding(NppPlan* p, const float* coords, long long n, cudaStream_t st, float* zero_loss,
                           bool* prefetched) {
  int hit = -1;
  for (int i = 0; i < 2; ++i)
    if (p->pref[i].valid && p->pref[i].coords == coords && p->pref[i].n == n && p->pref[i].tag == p->iter_tag) hit = i;
  if (hit >= 0) {
    // npp_encode_prefetch already wrote this batch's encoding into set `hit` (on the side stream, while earlier
    // work was running): wait for it, switch sets, and let the chain kernel clear the step accumulators
    CK(cudaStreamWaitEvent(st, p->pref[hit].done, 0));
    p->pref[hit].valid = false;
    p->enc_set = hit;
    *prefetched = true;
    ++p->launches;   // the encode launch belongs to this step
    return 0;
  }
  // encode in line, into a set nobody has prefetched into (or, failing that, over a prefetch that is now stale)
  int set = p->enc_set;
  if (p->pref[set].valid && !p->pref[1 - set].valid) set = 1 - set;
  if (p->pref[set].valid) {
    CK(cudaStreamWaitEvent(st, p->pref[set].done, 0));
    p->pref[set].valid = false;
  }
  p->enc_set = set;
  ProfScope ps(p, st, PROF_ENCODE, 1);
  CKI(launch_encode(p, coords, n, set, st, zero_loss));
  return 0;
}

static int run_forward(NppPlan* p, const float* coords, long long n, float* logits, cudaStream_t st, bool with_head = true,
                       const float* enc_f32 = nullptr, float* zero_loss = nullptr) {
  if (!p->params) return fail("npp_plan_bind has not been called");
  CKI(prepare(p, n));
  bool prefetched = false;
  if (enc_f32 != nullptr) {
    ProfScope ps(p, st, PROF_ENCODE, 1);
    if (p->pref[0].valid) {   // a pending prefetch into set 0 is overwritten: order after it, then forget it
      CK(cudaStreamWaitEvent(st, p->pref[0].done, 0));
      p->pref[0].valid = false;
    }
    p->enc_set = 0;
    __half* enca = p->buf_enca >= 0 ? p->bufs[p->buf_enca].ptr : nullptr;
    npp_load_encoding_kernel<<<p->num_sms * 8, 256, 0, st>>>(enc_f32, (int)n, npp_plan_encoding_width(p), p->E,
                                                             p->bufs[p->buf_enc1].ptr, p->Ep, enca, p->Ap);
    CK(cudaGetLastError());
    ++p->launches;
  } else {
    CKI(select_encoding(p, coords, n, st, zero_loss, &prefetched));
  }
  {
    ProfScope ps(p, st, PROF_GEMM_FWD, 1);
    const bool alt = p->enc_set != 0;
    CKI(launch_chain(alt ? p->d_fwd_ops_alt : p->d_fwd_ops, alt ? p->fwd_params_alt.data() : p->fwd_params.data(),
                     (int)p->layers.size(), (int)n, p->num_sms, st, p->fwd_subs, p->cluster,
                     prefetched && zero_loss ? p->acc : nullptr, prefetched && zero_loss ? (int)p->acc_zero_floats : 0,
                     prefetched ? zero_loss : nullptr, p->cfg.activation));
    ++p->launches;
  }
  CKI(mark_busy(p, st));
  if (!with_head) return 0;
  const Layer& last = p->layers.back();
  ProfScope ps_head(p, st, PROF_HEAD_LOSS, 1);
  npp_head_fwd_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(
      p->bufs[last.buf_h].ptr, last.out, p->head_width, (int)n, p->params + p->rgb_w_off, p->params + p->rgb_b_off,
      logits);
  CK(cudaGetLastError());
  ++p->launches;
  return 0;
}

// grad_logits must already be reflected in acc[amax] (loss kernel or amax kernel).
static int run_backward(NppPlan* p, long long n, const float* g, cudaStream_t st, bool finalize = true,
                        bool head_done = false) {
  if (!p->grads) return fail("npp_plan_bind was called without a gradient arena");
  CKI(prepare(p, n));
  CKI(set_smem_attrs());
  p->acc_clean = false;
  const Layer& last = p->layers.back();
  unsigned int* amax = reinterpret_cast<unsigned int*>(p->acc + p->amax_off);
  if (!head_done) {
  ProfScope ps(p, st, PROF_HEAD_LOSS, 1);
  npp_head_bwd_kernel<<<(unsigned)((n + HEAD_BWD_ROWS - 1) / HEAD_BWD_ROWS), 256, 0, st>>>(
      g, p->bufs[last.buf_h].ptr, p->bufs[last.buf_d].ptr, last.out, p->head_width, (int)n, p->params + p->rgb_w_off,
      amax, p->bufs[last.buf_delta].ptr, last.out, p->acc + p->headacc_off, nullptr);
  CK(cudaGetLastError());
  ++p->launches;
  }
  {
    ProfScope ps(p, st, PROF_GEMM_DGRAD, 1);
    CKI(launch_chain(p->d_dgrad_ops, p->dgrad_params.data(), (int)p->dgrads.size(), (int)n, p->num_sms, st,
                     p->dgrad_subs, p->cluster));
    ++p->launches;
  }
  {
    ProfScope ps(p, st, PROF_GEMM_WGRAD, 1);
    CKI(launch_wgrad(p->enc_set ? p->wg_params_alt : p->wg_params, p->num_sms, st, p->wg_cluster));
    ++p->launches;
  }
  CKI(mark_busy(p, st));
  if (finalize) {
    ProfScope ps(p, st, PROF_FINALIZE, 3);
    dim3 grid(128, (unsigned)p->layers.size());
    npp_grad_finalize_kernel<<<grid, 256, 0, st>>>(p->d_fin, p->partial, p->wg_params.n_splits, p->slab_stride, p->acc,
                                                   amax, p->grads, 0.f);
    CK(cudaGetLastError());
    const int hn = 3 * p->head_width + 3;
    // rgb_linear weight [3, W/2] and bias [3] are contiguous in the head accumulator and (up to the
    // 4-float arena padding) in the arena: copy them separately.
    npp_copy_kernel<<<2, 256, 0, st>>>(p->acc + p->headacc_off, p->grads + p->rgb_w_off, 3 * p->head_width);
    npp_copy_kernel<<<1, 32, 0, st>>>(p->acc + p->headacc_off + 3 * p->head_width, p->grads + p->rgb_b_off, 3);
    (void)hn;
    CK(cudaGetLastError());
    p->launches += 3;
  }
  return 0;
}

static int zero_acc(NppPlan* p, cudaStream_t st) {
  CK(cudaMemsetAsync(p->acc, 0, p->acc_zero_floats * sizeof(float), st));
  return 0;
}

static int run_adam(NppPlan* p, float lr, float beta1, float beta2, float eps, long long step, cudaStream_t st) {
  if (!p->params || !p->grads || !p->m || !p->v) return fail("npp_adam_step needs params, grads, exp_avg and exp_avg_sq bound");
  if (step < 1) return fail("Adam step must be >= 1");
  const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
  const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
  ProfScope ps(p, st, PROF_ADAM, 1);
  npp_adam_kernel<<<p->num_sms * 4, 256, 0, st>>>(p->params, p->grads, p->m, p->v, p->arena_trained, beta1, beta2,
                                                  step_size, inv_sqrt_bc2, eps, 1);
  CK(cudaGetLastError());
  ++p->launches;
  return 0;
}

static int run_shadow(NppPlan* p, cudaStream_t st) {
  ProfScope ps(p, st, PROF_ADAM, 1);
  dim3 grid(96, (unsigned)p->layers.size());
  npp_shadow_kernel<<<grid, 256, 0, st>>>(p->d_shadow, p->params);
  CK(cudaGetLastError());
  ++p->launches;
  return 0;
}

// ------------------------------------------------------------------------ entry points
extern "C" {

const char* npp_last_error(void) { return g_err.c_str(); }
int npp_abi_version(void) { return 1; }

int npp_plan_create(const NppConfig* cfg, NppPlan** out) {
  if (!cfg || !out) return fail("npp_plan_create: null argument");
  *out = nullptr;
  if (cfg->model != NPP_MODEL_TOPK && cfg->model != NPP_MODEL_TOP1 && cfg->model != NPP_MODEL_LIGHT)
    return fail("unknown model kind");
  if (cfg->model == NPP_MODEL_LIGHT) {
    if (cfg->topk != 1) return fail("NPP_Net_light takes one proposal (topk == 1)");
    if (cfg->include_input != 0) return fail("NPP_Net_light uses the search-mode encoders (include_input == 0)");
    if (cfg->n_freq < 1) return fail("NPP_Net_light needs the positional frequencies (n_freq >= 1)");
  }
  if (cfg->model == NPP_MODEL_TOPK && cfg->topk < 2) return fail("NPP_Net (top-K) needs topk >= 2");
  if (cfg->model == NPP_MODEL_TOP1 && cfg->topk != 1) return fail("NPP_Net_top1 needs topk == 1");
  if (cfg->topk > MAX_TOPK) return fail("topk exceeds MAX_TOPK=8");
  if (cfg->width != 256 && cfg->width != 512)
    return fail("netwidth must be 512 (the reference default) or 256 (the class default and the search-stage default)");
  if (cfg->activation != NPP_ACT_SNAKE && cfg->activation != NPP_ACT_RELU) return fail("unknown activation");
  if (cfg->depth < 2 || cfg->depth > 16) return fail("netdepth must be in [2,16]");
  if (cfg->skip_layer >= cfg->depth - 1) return fail("skip layer must be < depth-1");
  if (cfg->n_aug < 1 || cfg->n_aug > MAX_AUG) return fail("n_aug out of range");
  if (cfg->n_freq < 0 || cfg->n_freq > MAX_FREQ) return fail("n_freq out of range");
  if (cfg->max_rows < 1 || cfg->max_rows > (1LL << 24)) return fail("max_rows out of range");
  if (!cfg->cos_t || !cfg->sin_t || !cfg->period || (cfg->n_freq > 0 && !cfg->freq))
    return fail("encoder tables missing");
  int dev = 0;
  CK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(std::string("libnpp_b200 requires an sm_100 (B200) device, found sm_") + std::to_string(prop.major) +
                std::to_string(prop.minor) + "; there is no fallback path");
  NppPlan* p = new NppPlan();
  p->cfg = *cfg;
  p->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("NPP_CLUSTER")) p->cluster = atoi(e) == 1 ? 1 : 2;
  // Search-stage fits run many plans side by side on one GPU, and a pair CTA keeps its SM to itself (see launch_wgrad):
  // NPP_Net_light takes the single-CTA weight-gradient kernel, which shares SMs with the other candidates' small kernels
  // and whose L2 traffic does not matter at 2048 rows (same time per search).
  if (cfg->model == NPP_MODEL_LIGHT) p->wg_cluster = 1;
  if (const char* e = getenv("NPP_WG_CLUSTER")) p->wg_cluster = atoi(e) == 1 ? 1 : 2;
  if (const char* e = getenv("NPP_PDL")) p->pdl = atoi(e) != 0;
  if (const char* e = getenv("NPP_PDL_EDGES")) p->pdl_edges = atoi(e);
  if (const char* e = getenv("NPP_SPLIT_STEP")) p->fused_step = atoi(e) == 0;
  memset(&p->enc, 0, sizeof(p->enc));
  p->enc.topk = cfg->topk;
  p->enc.n_aug = cfg->n_aug;
  p->enc.n_freq = cfg->n_freq;
  p->enc.include_input = cfg->include_input;
  p->enc.res_h = (float)cfg->res_h;
  p->enc.res_w = (float)cfg->res_w;
  for (int j = 0; j < cfg->topk; ++j)
    for (int d = 0; d < 2; ++d)
      for (int a = 0; a < cfg->n_aug; ++a) {
        const int i = (j * 2 + d) * cfg->n_aug + a;
        p->enc.cos_t[j][d][a] = cfg->cos_t[i];
        p->enc.sin_t[j][d][a] = cfg->sin_t[i];
        p->enc.period[j][d][a] = cfg->period[i];
      }
  for (int k = 0; k < cfg->n_freq; ++k) p->enc.freq[k] = cfg->freq[k];
  p->cfg.cos_t = p->cfg.sin_t = p->cfg.period = p->cfg.freq = nullptr;
  int r = build_graph(p);
  if (r == 0) r = alloc_plan_memory(p);
  if (r != 0) {
    npp_plan_destroy(p);
    return r;
  }
  if (p->wg_cluster == 2) {   // see PairStepScope
    if (cudaEventCreateWithFlags(&p->step_evt, cudaEventDisableTiming) != cudaSuccess) {
      npp_plan_destroy(p);
      return fail("cudaEventCreate failed");
    }
    std::lock_guard<std::mutex> g(g_pair_mu);
    g_pair_plans.push_back(p);
  }
  *out = p;
  return 0;
}

int npp_plan_destroy(NppPlan* p) {
  if (!p) return 0;
  {
    std::lock_guard<std::mutex> g(g_pair_mu);
    for (size_t i = 0; i < g_pair_plans.size(); ++i)
      if (g_pair_plans[i] == p) {
        g_pair_plans.erase(g_pair_plans.begin() + i);
        break;
      }
    if (p->step_evt) {
      cudaEventSynchronize(p->step_evt);   // another plan's stream may still be waiting for it
      cudaEventDestroy(p->step_evt);
      p->step_evt = nullptr;
    }
  }
  {
    std::lock_guard<std::mutex> g(g_coop_mu);
    if (g_coop_plan == p) {
      cudaEventSynchronize(p->coop_evt);
      g_coop_plan = nullptr;
    }
  }
  if (p->fit_exec) {
    cudaStreamSynchronize(p->fit_stream);
    cudaGraphExecDestroy(p->fit_exec);
    p->fit_exec = nullptr;
  }
  cudaFree(p->workspace);
  cudaFree(p->shadow_mem);
  cudaFree(p->partial);
  cudaFree(p->acc);
  cudaFree(p->g_buf);
  cudaFree(p->d_barrier);
  cudaFree(p->d_step);
  cudaFree(p->d_ad_table);
  cudaFree(p->logits_buf);
  cudaFree(p->d_fin);
  cudaFree(p->d_shadow);
  cudaFree(p->d_update);
  cudaFree(p->d_fwd_ops);
  cudaFree(p->d_fwd_ops_alt);
  if (p->side_stream) cudaStreamDestroy(p->side_stream);
  if (p->pref_fork) cudaEventDestroy(p->pref_fork);
  if (p->enc_fork) cudaEventDestroy(p->enc_fork);
  if (p->enc_stream2) cudaStreamDestroy(p->enc_stream2);
  if (p->tables_evt) cudaEventDestroy(p->tables_evt);
  if (p->coop_evt) cudaEventDestroy(p->coop_evt);
  for (int i = 0; i < 2; ++i) if (p->pref[i].done) cudaEventDestroy(p->pref[i].done);
  cudaFree(p->d_dgrad_ops);
  cudaFree(p->d_step_ops);
  cudaFree(p->d_step_ops_alt);
  cudaFree(p->d_units);
  cudaFree(p->d_units_bal);
  cudaFree(p->d_tile_units);
  for (auto& g : p->groups) cudaFree(g.d_units);
  for (auto e : p->ev_pool) cudaEventDestroy(e);
  delete p;
  return 0;
}

int npp_plan_arena_floats(const NppPlan* p, int64_t* total, int64_t* trained) {
  if (!p) return fail("null plan");
  if (total) *total = p->arena_total;
  if (trained) *trained = p->arena_trained;
  return 0;
}
int npp_plan_tensor_count(const NppPlan* p) { return p ? (int)p->tensors.size() : 0; }
int npp_plan_tensor_info(const NppPlan* p, int i, NppTensorInfo* info) {
  if (!p || !info || i < 0 || i >= (int)p->tensors.size()) return fail("tensor index out of range");
  *info = p->tensors[i];
  return 0;
}
int npp_plan_encoding_width(const NppPlan* p) {
  if (!p) return 0;
  return p->cfg.model == NPP_MODEL_LIGHT ? p->E + p->A : p->E * p->cfg.topk;
}

int npp_plan_bind(NppPlan* p, float* params, float* grads, float* m, float* v) {
  if (!p || !params) return fail("npp_plan_bind: params arena is required");
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)m | (uintptr_t)v) & 15) return fail("arenas must be 16-byte aligned");
  p->params = params;
  p->grads = grads;
  p->m = m;
  p->v = v;
  p->prepared_n = -1;  // bias pointers live in the cached kernel parameters
  return 0;
}

int npp_sync_weights(NppPlan* p, void* stream) {
  if (!p || !p->params) return fail("npp_sync_weights: no parameter arena bound");
  return run_shadow(p, (cudaStream_t)stream);
}

int npp_encode(NppPlan* p, const float* coords, int64_t n, float* out, void* stream) {
  if (!p || !coords || !out) return fail("npp_encode: null argument");
  if (n <= 0) return 0;
  if (p->cfg.model == NPP_MODEL_LIGHT)
    npp_encode_search_f32_kernel<<<p->num_sms * 8, 256, 0, (cudaStream_t)stream>>>(coords, (int)n, p->enc, out);
  else
    npp_encode_f32_kernel<<<p->num_sms * 8, 256, 0, (cudaStream_t)stream>>>(coords, (int)n, p->enc, out);
  CK(cudaGetLastError());
  return 0;
}

int npp_encode_prefetch(NppPlan* p, const float* coords, int64_t n, void* stream) {
  if (!p || !coords) return fail("npp_encode_prefetch: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CKI(prepare(p, n));
  // The target set was last read by the step BEFORE the one that is running / about to be enqueued; everything
  // already enqueued on `stream` (that step, and whatever produces `coords`) is waited for, nothing enqueued later.
  // Target: the set the NEXT step will not use.  That step takes a pending prefetch if there is one, else the set of
  // the last step.  The target was last read by a step that is already enqueued on `stream`: everything enqueued
  // there so far (that step, and whatever produces `coords`) is waited for, nothing enqueued later -- call this
  // BEFORE enqueueing the step that should overlap with the encoding.
  int next_set = p->enc_set;
  for (int i = 0; i < 2; ++i)
    if (p->pref[i].valid) next_set = i;
  const int set = 1 - next_set;
  CK(cudaEventRecord(p->pref_fork, st));
  CK(cudaStreamWaitEvent(p->side_stream, p->pref_fork, 0));
  if (p->pref[set].valid) CK(cudaStreamWaitEvent(p->side_stream, p->pref[set].done, 0));
  const int launches = p->launches;
  CKI(launch_encode(p, coords, n, set, p->side_stream, nullptr));
  p->launches = launches;   // counted by the step that consumes it
  CK(cudaEventRecord(p->pref[set].done, p->side_stream));
  p->pref[set].coords = coords;
  p->pref[set].n = n;
  p->pref[set].valid = true;
  p->pref[set].tag = -1;
  return 0;
}

// Inside the capture of npp_multi_fit_run (branch stream bs, unrolled iteration p->iter_tag about to be captured): encodes
// the batch of the NEXT iteration (*d_step + 1) into the encoding set this iteration does not use, on a second capture
// stream forked from bs, so that it runs beside this iteration's chain kernel instead of in front of the next one.
static int capture_prefetch_next(NppPlan* p, const float* coords, long long n, cudaStream_t bs, int* set_out) {
  if (p->enc_stream2 == nullptr) {
    CK(cudaStreamCreateWithFlags(&p->enc_stream2, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&p->enc_fork, cudaEventDisableTiming));
  }
  int this_set = p->enc_set;   // the set the iteration about to be captured reads: its own prefetch, else an in-line encode
  for (int i = 0; i < 2; ++i)
    if (p->pref[i].valid && p->pref[i].tag == p->iter_tag) this_set = i;
  const int set = 1 - this_set;
  if (p->pref[set].valid) return fail("capture_prefetch_next: both encoding sets are taken");
  CK(cudaEventRecord(p->enc_fork, bs));
  CK(cudaStreamWaitEvent(p->enc_stream2, p->enc_fork, 0));
  const int launches = p->launches;
  p->enc_step_off = 1;
  const int rc = launch_encode(p, coords, n, set, p->enc_stream2, nullptr);
  p->enc_step_off = 0;
  if (rc != 0) return rc;
  p->launches = launches;   // counted by the iteration that consumes it
  CK(cudaEventRecord(p->pref[set].done, p->enc_stream2));
  p->pref[set].coords = coords;
  p->pref[set].n = n;
  p->pref[set].valid = true;
  p->pref[set].tag = p->iter_tag + 1;
  *set_out = set;
  return 0;
}

int npp_forward(NppPlan* p, const float* coords, int64_t n, float* logits, void* stream) {
  if (!p || !coords || !logits) return fail("npp_forward: null argument");
  p->launches = 0;
  return run_forward(p, coords, n, logits, (cudaStream_t)stream);
}

int npp_render_into(NppPlan* p, const float* coords, int64_t n, float* image, int32_t img_h, int32_t img_w,
                    int32_t normalize_type, void* stream) {
  if (!p || !coords || !image) return fail("npp_render_into: null argument");
  if (img_h <= 0 || img_w <= 0) return fail("npp_render_into: image size must be positive");
  if (normalize_type != 1 && normalize_type != 2) return fail("npp_render_into: normalize_type must be 1 (sigmoid) or 2 (tanh)");
  cudaStream_t st = (cudaStream_t)stream;
  int launches = 0;
  const Layer& last = p->layers.back();
  // Any number of pixels, in chunks that fit the workspace.  A chunk is a whole number of waves of the persistent chain
  // kernel (one 128-row stripe per SM: 18 944 rows on 148 SMs) when the workspace allows it -- 32 768-row chunks would
  // run a full wave and a 73 % full one.  The last chunk is moved back so that it has the same row count (the pixels it
  // shares with its predecessor are simply written twice): the per-row-count setup is not redone for a ragged tail.
  const int64_t wave = (int64_t)(p->num_sms / p->cluster * p->cluster) * BM;
  int64_t chunk = p->cfg.max_rows / wave * wave;
  if (chunk == 0) chunk = p->cfg.max_rows;
  if (chunk > n) chunk = n;
  for (int64_t done = 0; done < n; done += chunk) {
    const int64_t r0 = std::min<int64_t>(done, n - chunk);
    const int64_t rows = chunk;
    p->launches = 0;
    CKI(run_forward(p, coords + 2 * r0, rows, nullptr, st, /*with_head=*/false));
    npp_head_render_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(
        p->bufs[last.buf_h].ptr, last.out, p->head_width, (int)rows, p->params + p->rgb_w_off, p->params + p->rgb_b_off,
        coords + 2 * r0, image, img_h, img_w, normalize_type);
    CK(cudaGetLastError());
    launches += p->launches + 1;
  }
  p->launches = launches;
  return 0;
}

int npp_forward_encoded(NppPlan* p, const float* enc, int64_t n, float* logits, void* stream) {
  if (!p || !enc || !logits) return fail("npp_forward_encoded: null argument");
  p->launches = 0;
  return run_forward(p, nullptr, n, logits, (cudaStream_t)stream, true, enc);
}

int npp_backward(NppPlan* p, int64_t n, const float* g, void* stream) {
  if (!p || !g) return fail("npp_backward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  p->launches = 0;
  PairStepScope scope(p, st);
  CKI(scope.begin());
  CKI(zero_acc(p, st));
  npp_amax_kernel<<<64, 256, 0, st>>>(g, (int)(n * 3), reinterpret_cast<unsigned int*>(p->acc + p->amax_off));
  CK(cudaGetLastError());
  ++p->launches;
  CKI(run_backward(p, n, g, st));
  return scope.end();
}

int npp_mse_fwd_bwd(NppPlan* p, const float* logits, const float* target, const float* mask, int64_t n, int64_t n_norm,
                    float* pred, float* g, float* loss, void* stream) {
  if (!p || !logits || !target || !g || !loss) return fail("npp_mse_fwd_bwd: null argument");
  if (n_norm <= 0) return fail("n_norm must be positive");
  // a private amax slot is not needed: npp_backward recomputes it from g.
  unsigned int* scratch = reinterpret_cast<unsigned int*>(p->acc + p->amax_off + 1);
  const float inv_count = 1.0f / (3.0f * (float)n_norm);
  npp_mse_kernel<<<128, 256, 0, (cudaStream_t)stream>>>(logits, target, mask, (int)n, inv_count, pred, g, loss, scratch);
  CK(cudaGetLastError());
  return 0;
}

int npp_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, int64_t step, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq) return fail("npp_adam_flat: null argument");
  if (n <= 0) return 0;
  if (step < 1) return fail("Adam step must be >= 1");
  const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
  const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
  long long blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 592) blocks = 592;
  // views into larger tensors (storage offsets) need not be 16-byte aligned: those take the scalar loop
  const int vec4 = ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0) ? 1 : 0;
  npp_adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, (long long)n, beta1,
                                                                    beta2, (float)((double)lr / bc1),
                                                                    (float)(1.0 / std::sqrt(bc2)), eps, vec4);
  CK(cudaGetLastError());
  return 0;
}

int npp_gather_windows(const float* img, int32_t img_h, int32_t img_w, int32_t channels, const int64_t* rows,
                       const int64_t* cols, int64_t m, int32_t h, int32_t w, float* out, void* stream) {
  if (!img || !rows || !cols || !out) return fail("npp_gather_windows: null argument");
  if (img_h <= 0 || img_w <= 0 || channels <= 0 || h <= 0 || w <= 0) return fail("npp_gather_windows: sizes must be positive");
  if (m <= 0) return 0;
  const long long total = (long long)m * channels * h * w;
  if (m > (1LL << 24) || total > (1LL << 40)) return fail("npp_gather_windows: too many windows");
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  static_assert(sizeof(long long) == sizeof(int64_t), "index tables are int64");
  npp_gather_windows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      img, img_h, img_w, channels, reinterpret_cast<const long long*>(rows), reinterpret_cast<const long long*>(cols),
      (int)m, h, w, out);
  CK(cudaGetLastError());
  return 0;
}

int npp_sampler_candidates(const int64_t* sat, int32_t img_h, int32_t img_w, const int64_t* centroids, int32_t n_samples,
                           const int64_t* shifts4, int32_t half_h, int32_t half_w, float max_unknown, uint8_t* keep,
                           void* stream) {
  if (!sat || !centroids || !shifts4 || !keep) return fail("npp_sampler_candidates: null argument");
  if (img_h <= 0 || img_w <= 0 || half_h <= 0 || half_w <= 0) return fail("npp_sampler_candidates: sizes must be positive");
  if (n_samples <= 0) return 0;
  if (n_samples > (1 << 20)) return fail("npp_sampler_candidates: too many samples");
  static_assert(sizeof(long long) == sizeof(int64_t), "index tables are int64");
  const int total = n_samples * 400;
  npp_sampler_candidates_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(sat), img_h, img_w, reinterpret_cast<const long long*>(centroids), n_samples,
      shifts4[0], shifts4[1], shifts4[2], shifts4[3], half_h, half_w, max_unknown, keep);
  CK(cudaGetLastError());
  return 0;
}

int npp_l2_fwd_bwd(const float* x, const float* y, const float* mask, int64_t n, float* loss, float* grad_x,
                   void* stream) {
  if (!x || !y || !loss || !grad_x) return fail("npp_l2_fwd_bwd: null argument");
  if (n <= 0 || n > (1LL << 28)) return fail("npp_l2_fwd_bwd: row count out of range");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(loss, 0, sizeof(float), st));
  int blocks = (int)((3 * n + 255) / 256);
  if (blocks > 592) blocks = 592;
  npp_l2_kernel<<<blocks, 256, 0, st>>>(x, y, mask, (int)n, 1.0f / (3.0f * (float)n), grad_x, loss);
  CK(cudaGetLastError());
  return 0;
}

int npp_robust_adaptive_fwd_bwd(const float* x, const float* y, const float* mask, int64_t n, const float* latent_alpha,
                                const float* latent_scale, const float* cfg4, const float* logz_values,
                                const float* logz_derivs, int n_knots, float alpha_max, float* scratch9, float* out7,
                                float* grad_x, void* stream) {
  if (!x || !y || !latent_alpha || !latent_scale || !cfg4 || !logz_values || !logz_derivs || !scratch9 || !out7 || !grad_x)
    return fail("npp_robust_adaptive_fwd_bwd: null argument");
  if (n <= 0 || n > (1LL << 28)) return fail("npp_robust_adaptive_fwd_bwd: row count out of range");
  if (n_knots < 4) return fail("npp_robust_adaptive_fwd_bwd: log-partition table too small");
  cudaStream_t st = (cudaStream_t)stream;
  RobustCfg cfg;
  cfg.alpha_lo = cfg4[0]; cfg.alpha_hi = cfg4[1]; cfg.scale_lo = cfg4[2]; cfg.scale_ref = cfg4[3];
  cfg.n_knots = n_knots; cfg.alpha_max = alpha_max;
  if (!(cfg.alpha_lo > 0.f) || !(cfg.alpha_hi < 2.f) || !(cfg.alpha_lo < cfg.alpha_hi) || !(cfg.scale_lo < cfg.scale_ref))
    return fail("npp_robust_adaptive_fwd_bwd: needs 0 < alpha_lo < alpha_hi < 2 and scale_lo < scale_ref");
  CK(cudaMemsetAsync(scratch9, 0, 9 * sizeof(float), st));
  const float inv_count = 1.0f / (3.0f * (float)n);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  npp_robust_loss_kernel<<<blocks, 256, 0, st>>>(x, y, mask, (int)n, latent_alpha, latent_scale, cfg, logz_values,
                                                 logz_derivs, inv_count, grad_x, scratch9);
  CK(cudaGetLastError());
  npp_robust_finalize_kernel<<<1, 32, 0, st>>>(scratch9, latent_alpha, latent_scale, cfg, logz_values, logz_derivs,
                                               inv_count, out7);
  CK(cudaGetLastError());
  return 0;
}

int npp_adam_step(NppPlan* p, float lr, float beta1, float beta2, float eps, int64_t step, void* stream) {
  if (!p) return fail("null plan");
  p->launches = 0;
  CKI(run_adam(p, lr, beta1, beta2, eps, step, (cudaStream_t)stream));
  return run_shadow(p, (cudaStream_t)stream);
}

// The update kernel of a train step: split-K slabs -> gradient -> Adam -> fp32 master + fp16 shadows.
static int launch_update(NppPlan* p, const AdamScalars& ad, const unsigned int* amax_bits, const StepReset& rs, int pdl,
                         cudaStream_t st) {
  const unsigned blocks = (unsigned)p->update_table.tile_begin[p->update_table.n_layers] + 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  const float* partial = p->partial;
  float* bias_acc = p->acc;
  float* head_acc = p->acc + p->headacc_off;
  float* grads = p->keep_grads ? p->grads : nullptr;
  const AdamScalars* ad_table = p->step_mode ? p->d_ad_table : nullptr;
  const int* d_step = p->d_step;
#define NPP_UPDATE_CASE(S_)                                                                                            \
  case S_:                                                                                                             \
    CK(cudaLaunchKernelEx(&cfg, npp_fused_update_kernel<S_>, p->update_table, partial, p->slab_stride, bias_acc,       \
                          head_acc, p->rgb_w_off, p->rgb_b_off, p->head_width, amax_bits, p->params, grads, p->m,      \
                          p->v, ad, ad_table, d_step, rs));                                                            \
    break;
  switch (p->wg_params.n_splits) {
    NPP_UPDATE_CASE(1) NPP_UPDATE_CASE(2) NPP_UPDATE_CASE(3) NPP_UPDATE_CASE(4) NPP_UPDATE_CASE(5) NPP_UPDATE_CASE(6)
    NPP_UPDATE_CASE(7) NPP_UPDATE_CASE(8) NPP_UPDATE_CASE(9) NPP_UPDATE_CASE(10) NPP_UPDATE_CASE(11)
    NPP_UPDATE_CASE(12)
    default: return fail("unsupported split-K factor in the fused update");
  }
#undef NPP_UPDATE_CASE
  ++p->launches;
  return 0;
}

static AdamScalars adam_scalars(float lr, float beta1, float beta2, float eps, long long step) {
  AdamScalars ad;
  const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
  const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
  ad.beta1 = beta1;
  ad.beta2 = beta2;
  ad.step_size = (float)((double)lr / bc1);
  ad.inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
  ad.eps = eps;
  return ad;
}

// Fused train step, three launches: [forward chain -> RGB head + loss + head backward -> dgrad chain] as ONE persistent
// kernel (every 128-row stripe runs all of it without a grid-wide dependency, because the fp16 delta scale comes from the
// previous step's max|g|: npp_step_amax), the grouped weight-gradient GEMM, and the update kernel.  Consecutive
// launches are programmatic dependents: a kernel's prologue overlaps its predecessor's tail.
struct StepSlots {
  unsigned int *prev, *cur, *clr;
  float* loss_acc;
  float inv_count;
};
static StepSlots step_slots(NppPlan* p, long long n_norm) {
  unsigned int* ring = reinterpret_cast<unsigned int*>(p->acc + p->ring_off);
  StepSlots s;
  s.cur = ring + (int)(p->step_seq % 3);
  s.prev = ring + (int)((p->step_seq + 2) % 3);
  s.clr = ring + (int)((p->step_seq + 1) % 3);
  s.loss_acc = p->acc + p->ring_off + 3;
  s.inv_count = 1.0f / (3.0f * (float)n_norm);
  return s;
}

// First launch of a fused step: forward chain + RGB head + loss + head backward + dgrad chain (one persistent kernel).
static int launch_step_chain(NppPlan* p, const float* coords, const float* target, const float* mask, long long n,
                             const StepSlots& sl, cudaStream_t st) {
  if (!p->params) return fail("npp_plan_bind has not been called");
  CKI(prepare(p, n));
  CKI(set_smem_attrs());
  if (!p->acc_clean) {   // a call of the unfused path, or an abandoned three-phase step, left its sums behind
    CK(cudaMemsetAsync(p->acc, 0, p->acc_zero_floats * sizeof(float), st));
    CK(cudaMemsetAsync(p->acc + p->ring_off + 3, 0, sizeof(float), st));   // the loss accumulator lives behind the ring
    p->acc_clean = true;
  }
  bool prefetched = false;
  CKI(select_encoding(p, coords, n, st, nullptr, &prefetched));
  const Layer& last = p->layers.back();
  if (last.out != BN) return fail("fused step: the last dense layer must be one 256-column tile");
  HeadArgs hd;
  memset(&hd, 0, sizeof(hd));
  hd.w = p->params + p->rgb_w_off;
  hd.b = p->params + p->rgb_b_off;
  hd.target = target;
  hd.mask = mask;
  hd.logits = nullptr;
  hd.head_acc = p->acc + p->headacc_off;
  hd.loss_acc = sl.loss_acc;
  hd.amax_prev = sl.prev;
  hd.amax_next = sl.cur;
  hd.inv_count = sl.inv_count;
  hd.width = p->head_width;
  if (p->step_mode) {   // one captured step re-launched many times: batch and ring slots follow the device-side step index
    hd.step = p->d_step;
    hd.ring = reinterpret_cast<unsigned int*>(p->acc + p->ring_off);
    hd.seq0 = (int)(p->step_seq % 3);
  }
  const bool alt = p->enc_set != 0;
  ProfScope ps(p, st, PROF_GEMM_FWD, 1);
  const std::vector<KmajorParams>& ops = alt ? p->step_params_alt : p->step_params;
  CKI(launch_chain(alt ? p->d_step_ops_alt : p->d_step_ops, ops.data(), (int)ops.size(), (int)n, p->num_sms, st,
                   p->fwd_subs + p->dgrad_subs, p->cluster, nullptr, 0, nullptr, p->cfg.activation, &hd, p->pdl && (p->pdl_edges & 1)));
  ++p->launches;
  return 0;
}

static int run_fused_step(NppPlan* p, const float* coords, const float* target, const float* mask, long long n,
                          long long n_norm, const AdamScalars& ad, float* loss, cudaStream_t st) {
  if (!p->m || !p->v) return fail("npp_train_step needs exp_avg and exp_avg_sq bound");
  const StepSlots sl = step_slots(p, n_norm);
  CKI(launch_step_chain(p, coords, target, mask, n, sl, st));
  const bool alt = p->enc_set != 0;
  {
    ProfScope ps(p, st, PROF_GEMM_WGRAD, 1);
    CKI(launch_wgrad(alt ? p->wg_params_alt : p->wg_params, p->num_sms, st, p->wg_cluster, p->pdl && (p->pdl_edges & 2)));
    ++p->launches;
  }
  CKI(mark_busy(p, st));
  {
    ProfScope ps(p, st, PROF_ADAM, 1);
    StepReset rs;
    rs.on = 1;
    rs.inv_count = sl.inv_count;
    rs.loss_acc = sl.loss_acc;
    rs.loss_out = loss;
    rs.amax_clear = sl.clr;
    rs.ring = reinterpret_cast<unsigned int*>(p->acc + p->ring_off);
    rs.seq0 = (int)(p->step_seq % 3);
    rs.by_step = p->step_mode ? 1 : 0;
    CKI(launch_update(p, ad, sl.prev, rs, p->pdl && (p->pdl_edges & 4), st));
  }
  if (!p->step_mode) ++p->step_seq;   // a re-launched step graph advances the ring on the device; its caller adds `iters`
  return 0;
}

static int train_step_impl(NppPlan* p, const float* coords, const float* target, const float* mask, int64_t n,
                           int64_t n_norm, float lr, float beta1, float beta2, float eps, int64_t step, float* loss,
                           void* stream);

int npp_train_step(NppPlan* p, const float* coords, const float* target, const float* mask, int64_t n, int64_t n_norm,
                   float lr, float beta1, float beta2, float eps, int64_t step, float* loss, void* stream) {
  if (!p || !coords || !target || !loss) return fail("npp_train_step: null argument");
  if (n_norm <= 0) return fail("n_norm must be positive");
  if (p->splits_override != 0 && !p->capturing) {   // left over from a grouped fit (npp_multi_fit_run)
    p->splits_override = 0;
    p->prepared_n = -1;
  }
  PairStepScope scope(p, (cudaStream_t)stream);
  CKI(scope.begin());
  CKI(train_step_impl(p, coords, target, mask, n, n_norm, lr, beta1, beta2, eps, step, loss, stream));
  return scope.end();
}

static int train_step_impl(NppPlan* p, const float* coords, const float* target, const float* mask, int64_t n,
                           int64_t n_norm, float lr, float beta1, float beta2, float eps, int64_t step, float* loss,
                           void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  p->launches = 0;
  if (p->fused_step && (!p->step_mode || p->cfg.model == NPP_MODEL_LIGHT)) {
    if (step < 1) return fail("Adam step must be >= 1");
    return run_fused_step(p, coords, target, mask, n, n_norm, adam_scalars(lr, beta1, beta2, eps, step), loss, st);
  }
  p->acc_clean = false;
  // the encode kernel clears the accumulators and the loss scalar of this step
  CKI(run_forward(p, coords, n, p->logits_buf, st, /*with_head=*/false, nullptr, loss));
  const float inv_count = 1.0f / (3.0f * (float)n_norm);
  // The fused head is a cooperative launch with a grid barrier.  Two such grids running at the same time (two plans
  // on two streams) can each hold part of the SMs and wait for the rest: search-stage fits are meant to run several
  // candidates side by side (NPP_proposal/search.py:85 loops over up to 9), so NPP_Net_light takes the two-kernel head.
  bool fused_head = p->head_fused_blocks > 0 && p->cfg.model != NPP_MODEL_LIGHT && !getenv("NPP_SPLIT_HEAD");
  std::unique_lock<std::mutex> coop_lock(g_coop_mu, std::defer_lock);
  if (fused_head) {
    coop_lock.lock();
    if (!coop_head_allowed(p)) {
      fused_head = false;
      coop_lock.unlock();
    }
  }
  if (fused_head) {
    // head forward + loss + head backward in one cooperative launch (grid barrier around the max|g| reduction)
    CKI(prepare(p, n));
    ProfScope ps(p, st, PROF_HEAD_LOSS, 1);
    const Layer& last = p->layers.back();
    const __half* hp = p->bufs[last.buf_h].ptr;
    const __half* dp = p->bufs[last.buf_d].ptr;
    int ld = last.out, width = p->head_width, ni = (int)n, ldd = last.out;
    const float* w = p->params + p->rgb_w_off;
    const float* b = p->params + p->rgb_b_off;
    float ic = inv_count;
    float* logits = p->logits_buf;
    float* g = p->g_buf;
    unsigned int* amax = reinterpret_cast<unsigned int*>(p->acc + p->amax_off);
    __half* delta = p->bufs[last.buf_delta].ptr;
    float* head_acc = p->acc + p->headacc_off;
    float* bias_acc = nullptr;   // the weight-gradient kernel sums the bias gradients (WgUnit::bias_off)
    unsigned int* bar = p->d_barrier;
    void* args[] = {&hp, &dp, &ld, &width, &ni, &w, &b, &target, &mask, &ic, &logits, &g, &loss, &amax,
                    &delta, &ldd, &head_acc, &bias_acc, &bar};
    int blocks = p->head_fused_blocks;
    const int want = (int)((n + 7) / 8);   // at least one row per warp
    if (blocks > want) blocks = want;
    int rows_per_block = (int)((n + blocks - 1) / blocks);
    if (rows_per_block <= 8 * HEAD_MAXR && width == 256 && !getenv("NPP_HEAD_GENERIC")) {
      void* rargs[] = {&hp, &dp, &ld, &ni, &w, &b, &target, &mask, &ic, &logits, &g, &loss, &amax,
                       &delta, &ldd, &head_acc, &bias_acc, &bar, &rows_per_block};
      CK(cudaLaunchCooperativeKernel((const void*)npp_head_fused_reg_kernel, dim3((unsigned)blocks), dim3(256), rargs, 0,
                                     st));
    } else {
      CK(cudaLaunchCooperativeKernel((const void*)npp_head_fused_kernel, dim3((unsigned)blocks), dim3(256), args, 0,
                                     st));
    }
    ++p->launches;
    CK(cudaEventRecord(p->coop_evt, st));
    g_coop_plan = p;
    coop_lock.unlock();
  } else {
    ProfScope ps(p, st, PROF_HEAD_LOSS, 1);
    const Layer& last = p->layers.back();
    int blocks = (int)((n + 31) / 32);   // 8 warps x 4 rows in flight each
    if (blocks > p->num_sms * 8) blocks = p->num_sms * 8;
    npp_head_loss_kernel<<<blocks, 256, 0, st>>>(p->bufs[last.buf_h].ptr, last.out, p->head_width, (int)n,
                                                 p->params + p->rgb_w_off, p->params + p->rgb_b_off, target, mask,
                                                 inv_count, p->logits_buf, p->g_buf, loss,
                                                 reinterpret_cast<unsigned int*>(p->acc + p->amax_off),
                                                 p->step_mode ? p->d_step : nullptr);
    CK(cudaGetLastError());
    ++p->launches;
  }
  CKI(run_backward(p, n, p->g_buf, st, /*finalize=*/false, /*head_done=*/fused_head));
  {
    if (!p->m || !p->v) return fail("npp_train_step needs exp_avg and exp_avg_sq bound");
    if (step < 1) return fail("Adam step must be >= 1");
    ProfScope ps(p, st, PROF_ADAM, 1);
    StepReset rs;
    memset(&rs, 0, sizeof(rs));
    CKI(launch_update(p, adam_scalars(lr, beta1, beta2, eps, step), reinterpret_cast<unsigned int*>(p->acc + p->amax_off), rs,
                      0, st));
  }
  return 0;
}

// ---- the fused step in three phases (data parallelism inside one image: the gradient arena must exist between the
//      backward pass and Adam, and the all-reduce of one layer group should overlap the weight-gradient GEMMs of the next)
int npp_step_forward_backward(NppPlan* p, const float* coords, const float* target, const float* mask, int64_t n,
                              int64_t n_norm, void* stream) {
  if (!p || !coords || !target) return fail("npp_step_forward_backward: null argument");
  if (n_norm <= 0) return fail("n_norm must be positive");
  if (!p->grads) return fail("npp_plan_bind was called without a gradient arena");
  p->launches = 0;
  PairStepScope scope(p, (cudaStream_t)stream);   // waits only; npp_step_finish records the event
  CKI(scope.begin());
  const StepSlots sl = step_slots(p, n_norm);
  CKI(launch_step_chain(p, coords, target, mask, n, sl, (cudaStream_t)stream));
  // until npp_step_finish has run, the step accumulators hold this step's sums: a caller that abandons the step here
  // must not have them added to the next one (launch_step_chain clears them when it finds them dirty)
  p->acc_clean = false;
  return mark_busy(p, (cudaStream_t)stream);
}

int npp_plan_layer_count(const NppPlan* p) { return p ? (int)p->layers.size() : 0; }

int npp_plan_layer_grad_range(const NppPlan* p, int32_t layer_begin, int32_t layer_end, int64_t* offset, int64_t* count) {
  if (!p || !offset || !count) return fail("npp_plan_layer_grad_range: null argument");
  const int nl = (int)p->layers.size();
  if (layer_begin < 0 || layer_end > nl || layer_begin >= layer_end) return fail("layer range out of bounds");
  *offset = p->layers[layer_begin].w_off;
  // weights and biases are laid out layer by layer, rgb_linear right behind the last dense layer
  const long long end = layer_end == nl ? p->arena_trained : p->layers[layer_end].w_off;
  *count = end - *offset;
  return 0;
}

int npp_step_wgrad(NppPlan* p, int32_t layer_begin, int32_t layer_end, int64_t n, int64_t n_norm, void* stream) {
  if (!p) return fail("null plan");
  const int nl = (int)p->layers.size();
  if (layer_begin < 0 || layer_end > nl || layer_begin >= layer_end) return fail("npp_step_wgrad: layer range out of bounds");
  if (p->prepared_n != n) return fail("npp_step_wgrad: call npp_step_forward_backward with the same row count first");
  cudaStream_t st = (cudaStream_t)stream;
  // split factor of this launch: enough row splits to give every CTA pair a unit, at least one per 32768 rows
  const int tiles = p->tile_begin[layer_end] - p->tile_begin[layer_begin];
  const int slots = p->num_sms / p->wg_cluster;
  int S = std::max(1, (slots + tiles / 2) / tiles);
  S = std::max(S, (int)((n + WG_ROWS_PER_SPLIT - 1) / WG_ROWS_PER_SPLIT));
  S = std::min(S, std::min(p->slabs_alloc, (int)NPP_MAX_SPLITS));
  const int kb_total = (int)((n + BK - 1) / BK);
  if (S > kb_total) S = kb_total;
  NppPlan::GroupTable* gt = nullptr;
  for (auto& g : p->groups)
    if (g.lb == layer_begin && g.le == layer_end && g.splits == S) gt = &g;
  if (gt == nullptr) {
    std::vector<WgUnit> units;
    for (int sidx = 0; sidx < S; ++sidx)
      for (int t = p->tile_begin[layer_begin]; t < p->tile_begin[layer_end]; ++t) {
        WgUnit u = p->tile_units[t];
        u.split = sidx;
        units.push_back(u);
      }
    NppPlan::GroupTable g;
    g.lb = layer_begin; g.le = layer_end; g.splits = S; g.n_units = (int)units.size(); g.d_units = nullptr;
    CK(cudaMalloc(&g.d_units, units.size() * sizeof(WgUnit)));
    CK(cudaMemcpy(g.d_units, units.data(), units.size() * sizeof(WgUnit), cudaMemcpyHostToDevice));
    p->groups.push_back(g);
    gt = &p->groups.back();
  }
  WgradParams w = p->enc_set ? p->wg_params_alt : p->wg_params;
  w.units = gt->d_units;
  w.n_units = gt->n_units;
  w.grid_pairs = 0;
  w.kb_per_split = (kb_total + S - 1) / S;
  w.n_splits = (kb_total + w.kb_per_split - 1) / w.kb_per_split;
  {
    ProfScope ps(p, st, PROF_GEMM_WGRAD, 1);
    CKI(launch_wgrad(w, p->num_sms, st, p->wg_cluster, 0));
    ++p->launches;
  }
  {
    ProfScope ps(p, st, PROF_FINALIZE, 1);
    const StepSlots sl = step_slots(p, n_norm);
    dim3 grid(128, (unsigned)(layer_end - layer_begin));
    npp_grad_finalize_kernel<<<grid, 256, 0, st>>>(p->d_fin + layer_begin, p->partial, w.n_splits, p->slab_stride, p->acc,
                                                   sl.prev, p->grads, sl.inv_count);
    CK(cudaGetLastError());
    ++p->launches;
    if (layer_end == nl) {   // rgb_linear: unscaled fp32 sums written by the head epilogue
      npp_copy_kernel<<<2, 256, 0, st>>>(p->acc + p->headacc_off, p->grads + p->rgb_w_off, 3 * p->head_width);
      npp_copy_kernel<<<1, 32, 0, st>>>(p->acc + p->headacc_off + 3 * p->head_width, p->grads + p->rgb_b_off, 3);
      CK(cudaGetLastError());
      p->launches += 2;
    }
  }
  return mark_busy(p, st);
}

int npp_step_finish(NppPlan* p, int64_t n_norm, float lr, float beta1, float beta2, float eps, int64_t step, float* loss,
                    void* stream) {
  if (!p) return fail("null plan");
  if (n_norm <= 0) return fail("n_norm must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  CKI(run_adam(p, lr, beta1, beta2, eps, step, st));
  CKI(run_shadow(p, st));
  const StepSlots sl = step_slots(p, n_norm);
  npp_step_finish_kernel<<<8, 256, 0, st>>>(p->acc, (int)p->acc_zero_floats, sl.loss_acc, loss, sl.clr);
  CK(cudaGetLastError());
  ++p->launches;
  ++p->step_seq;
  p->acc_clean = true;
  PairStepScope scope(p, st);
  return scope.end();
}

int npp_fit_run(NppPlan* p, const float* coords_all, const float* target_all, const float* mask_all, int64_t n,
                int64_t iters, float lrate, float decay_rate, float decay_steps, float beta1, float beta2, float eps,
                int64_t first_step, float* losses, void* stream) {
  if (!p || !coords_all || !target_all || !losses) return fail("npp_fit_run: null argument");
  if (iters < 0 || first_step < 1) return fail("npp_fit_run: iters must be >= 0 and first_step >= 1");
  if (!(decay_steps > 0.f) || !(decay_rate > 0.f)) return fail("npp_fit_run: decay_rate and decay_steps must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  auto lr_of = [&](int64_t k) {   // Adam step k (1-based) of the scripts' loop, see the header
    const double expo = (double)(k > 2 ? k - 2 : 0) / (double)decay_steps;
    return (float)((double)lrate * std::pow((double)decay_rate, expo));
  };
  // NPP_FIT_GRAPH: 0 = plain launches (default), 1 = the whole run captured into one graph, 2 = one step captured and
  // re-launched `iters` times (per-step scalars through device memory).  Measured on B200, NPP_Net_light, 300
  // iterations of 2048 rows (profiles/r01/search_fit_timing.txt): one fit 32.9 / 39.6 / 29.7 ms in modes 0 / 1 / 2,
  // but nine fits from nine host threads 95 / 140 / 131 ms: graph launches from different streams did not overlap the
  // way plain launches do, and a one-shot 2100-node graph costs ~7 ms to build.  Hence plain launches by default.
  const char* genv = getenv("NPP_FIT_GRAPH");
  int mode = genv ? atoi(genv) : 0;
  if (p->profiling || iters < 4 || p->cfg.model != NPP_MODEL_LIGHT) mode = 0;   // (the cooperative head cannot be captured)
  int launches = 0;
  if (mode == 0) {
    for (int64_t i = 0; i < iters; ++i) {
      CKI(npp_train_step(p, coords_all + i * n * 2, target_all + i * n * 3, mask_all ? mask_all + i * n : nullptr, n, n,
                         lr_of(first_step + i), beta1, beta2, eps, first_step + i, losses + i, st));
      launches += p->launches;
    }
    p->launches = launches;
    return 0;
  }
  if (!p->params) return fail("npp_plan_bind has not been called");
  CKI(prepare(p, n));          // synchronises and copies op tables: must happen outside the capture
  CKI(set_smem_attrs());
  if (p->fit_exec) {           // the previous run's graph may still be executing
    CK(cudaStreamSynchronize(p->fit_stream));
    CK(cudaGraphExecDestroy(p->fit_exec));
    p->fit_exec = nullptr;
  }
  // The capture records into the plan's own non-blocking stream (nothing executes there): a capture on the caller's
  // stream would be invalidated by any activity on the legacy default stream, which blocking streams synchronise with.
  p->pref[0].valid = p->pref[1].valid = false;   // a pending prefetch lives on the capture stream's real timeline
  CK(cudaStreamSynchronize(p->side_stream));
  if (mode == 2) {
    // Per-step scalars go through device memory so that ONE captured step can be launched `iters` times: the kernels
    // read the batch index from d_step (advanced by the graph's last node) and Adam's scalars from a table.
    if (p->ad_table_cap < iters) {
      cudaFree(p->d_ad_table);
      p->d_ad_table = nullptr;
      CK(cudaMalloc(&p->d_ad_table, (size_t)iters * sizeof(AdamScalars)));
      p->ad_table_cap = iters;
    }
    std::vector<AdamScalars> tab((size_t)iters);
    for (int64_t i = 0; i < iters; ++i) {
      const int64_t k = first_step + i;
      const double bc1 = 1.0 - std::pow((double)beta1, (double)k);
      const double bc2 = 1.0 - std::pow((double)beta2, (double)k);
      tab[i].beta1 = beta1;
      tab[i].beta2 = beta2;
      tab[i].step_size = (float)((double)lr_of(k) / bc1);
      tab[i].inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
      tab[i].eps = eps;
    }
    // pageable source: the call returns once the table sits in the driver's staging buffer
    CK(cudaMemcpyAsync(p->d_ad_table, tab.data(), tab.size() * sizeof(AdamScalars), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  }
  CK(cudaStreamBeginCapture(p->side_stream, cudaStreamCaptureModeThreadLocal));
  p->capturing = true;
  int rc = 0;
  if (mode == 2) {
    p->step_mode = true;
    rc = npp_train_step(p, coords_all, target_all, mask_all, n, n, lrate, beta1, beta2, eps, first_step, losses,
                        p->side_stream);
    if (rc == 0) {
      npp_step_advance_kernel<<<1, 1, 0, p->side_stream>>>(p->d_step);
      ++p->launches;
    }
    launches = p->launches * (int)iters;
    p->step_mode = false;
  } else {
    for (int64_t i = 0; i < iters && rc == 0; ++i) {
      rc = npp_train_step(p, coords_all + i * n * 2, target_all + i * n * 3, mask_all ? mask_all + i * n : nullptr, n, n,
                          lr_of(first_step + i), beta1, beta2, eps, first_step + i, losses + i, p->side_stream);
      launches += p->launches;
    }
  }
  p->capturing = false;
  cudaGraph_t g = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(p->side_stream, &g);
  if (rc != 0) {
    if (g) cudaGraphDestroy(g);
    return rc;
  }
  if (ce != cudaSuccess) return fail(std::string("npp_fit_run: stream capture failed: ") + cudaGetErrorString(ce));
  const cudaError_t ie = cudaGraphInstantiate(&p->fit_exec, g, 0ULL);
  cudaGraphDestroy(g);
  if (ie != cudaSuccess) {
    p->fit_exec = nullptr;
    return fail(std::string("npp_fit_run: cudaGraphInstantiate failed: ") + cudaGetErrorString(ie));
  }
  p->fit_stream = st;
  const int64_t graph_launches = mode == 2 ? iters : 1;
  for (int64_t i = 0; i < graph_launches; ++i) CK(cudaGraphLaunch(p->fit_exec, st));
  if (mode == 2 && p->fused_step) p->step_seq += iters;
  CKI(mark_busy(p, st));
  p->launches = launches;
  return 0;
}

// Several fits advanced in lock step by ONE re-launched CUDA graph: the candidate loop of NPP_proposal/search.py:85-148 with
// every candidate fitted on the same sequence of batches.  One train step of every plan is captured as a parallel branch
// of a single graph (fork / join through events, each plan on its own capture stream), with the batch index, Adam's
// scalars and the max-gradient ring following a device-side step counter per plan; the graph is then launched `iters`
// times from the caller's stream.  Replaces k * iters * 4 kernel launches from k host threads (the run was bound by
// the launch rate) by `iters` graph launches whose branches run side by side on the GPU.
int npp_multi_fit_run(NppPlan* const* plans, int32_t k, const float* const* coords_all, const float* const* target_all,
                      const float* const* mask_all, int64_t n, int64_t iters, float lrate, float decay_rate,
                      float decay_steps, float beta1, float beta2, float eps, const int64_t* first_steps,
                      float* const* losses, void* stream) {
  if (!plans || k <= 0 || !coords_all || !target_all || !losses || !first_steps) return fail("npp_multi_fit_run: null argument");
  if (iters < 1) return fail("npp_multi_fit_run: iters must be >= 1");
  if (!(decay_steps > 0.f) || !(decay_rate > 0.f)) return fail("npp_multi_fit_run: decay_rate and decay_steps must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < k; ++i) {
    NppPlan* p = plans[i];
    if (!p || !coords_all[i] || !target_all[i] || !losses[i]) return fail("npp_multi_fit_run: null argument");
    if (p->cfg.model != NPP_MODEL_LIGHT) return fail("npp_multi_fit_run: built for the search-stage network (NPP_MODEL_LIGHT)");
    if (!p->fused_step) return fail("npp_multi_fit_run needs the fused train step (NPP_SPLIT_STEP is set)");
    if (first_steps[i] < 1) return fail("npp_multi_fit_run: first_step must be >= 1");
    if (!p->params || !p->m || !p->v) return fail("npp_multi_fit_run: parameter and Adam arenas must be bound");
    for (int j = 0; j < i; ++j)
      if (plans[j] == p) return fail("npp_multi_fit_run: the same plan is listed twice");
  }
  NppPlan* lead = plans[0];
  if (lead->fit_exec) {           // the previous run's graph may still be executing
    CK(cudaStreamSynchronize(lead->fit_stream));
    CK(cudaGraphExecDestroy(lead->fit_exec));
    lead->fit_exec = nullptr;
  }
  CKI(set_smem_attrs());
  for (int i = 0; i < k; ++i) {
    NppPlan* p = plans[i];
    // Row splits of the weight-gradient GEMM: a fit on its own is fastest with 4 (103 us per step against 109 with 1), but
    // side by side the fits compete for SMs and every split is another CTA with its prologue: nine candidates x 300
    // iterations take 46.1 / 44.1 / 41.8 / 39.9 ms with 4 / 3 / 2 / 1 splits (tools/bench_search_fits.py --mode grouped).
    const int want_override = (k >= 3 && p->cfg.model == NPP_MODEL_LIGHT && p->splits_auto) ? 1 : 0;
    if (p->splits_override != want_override) {
      p->splits_override = want_override;
      p->prepared_n = -1;
    }
    CKI(prepare(p, n));           // synchronises and copies op tables: must happen outside the capture
    p->pref[0].valid = p->pref[1].valid = false;
    CK(cudaStreamSynchronize(p->side_stream));
    if (p->ad_table_cap < iters) {
      cudaFree(p->d_ad_table);
      p->d_ad_table = nullptr;
      CK(cudaMalloc(&p->d_ad_table, (size_t)iters * sizeof(AdamScalars)));
      p->ad_table_cap = iters;
    }
    std::vector<AdamScalars> tab((size_t)iters);
    for (int64_t it = 0; it < iters; ++it) {
      const int64_t step = first_steps[i] + it;
      const double expo = (double)(step > 2 ? step - 2 : 0) / (double)decay_steps;   // the scripts' LR rewrite, see npp_fit_run
      tab[it] = adam_scalars((float)((double)lrate * std::pow((double)decay_rate, expo)), beta1, beta2, eps, step);
    }
    CK(cudaMemcpyAsync(p->d_ad_table, tab.data(), tab.size() * sizeof(AdamScalars), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(p->d_step, 0, sizeof(int), st));
  }
  // Iterations per graph launch: the branches of one launch join before the next launch starts, so every unrolled
  // iteration is one fork / join (and one graph launch) less, at the price of a larger graph to instantiate per call.
  // Nine candidates x 300 iterations: 39.8 / 38.8 / 38.4 / 39.6 / 42.0 ms with 1 / 2 / 5 / 10 / 30 iterations per launch.
  int unroll = 5;
  if (const char* e = getenv("NPP_FIT_UNROLL")) unroll = std::max(1, atoi(e));
  while (unroll > 1 && iters % unroll != 0) --unroll;
  // NPP_FIT_PIPELINE=1 (off by default): inside a launch the encoding of iteration u + 1 runs beside the chain kernel of
  // iteration u instead of in front of its own.  Measured slower for nine candidates (42.6 against 38.6 ms per search:
  // the extra fork / join per iteration costs more than the 8 us encode kernel it hides).
  bool pipeline = false;
  if (const char* e = getenv("NPP_FIT_PIPELINE")) pipeline = unroll > 1 && atoi(e) != 0;
  for (int i = 0; i < k; ++i) pipeline = pipeline && plans[i]->cfg.model == NPP_MODEL_LIGHT;
  // capture: the lead plan's side stream is the origin, every other plan's side stream a branch forked from it
  cudaStream_t origin = lead->side_stream;
  CK(cudaStreamBeginCapture(origin, cudaStreamCaptureModeThreadLocal));
  int rc = 0;
  int launches = 0;
  // NPP_PAIR_SERIALISE=1 (see PairStepScope): branches of pair plans are chained one behind the other instead of forked
  bool serial = false;
  for (int i = 0; i < k; ++i) serial = serial || (pair_serialise() && plans[i]->wg_cluster == 2);
  cudaError_t ce = cudaEventRecord(lead->pref_fork, origin);
  for (int i = 0; i < k && rc == 0 && ce == cudaSuccess; ++i) {
    NppPlan* p = plans[i];
    cudaStream_t bs = p->side_stream;
    if (i > 0) ce = cudaStreamWaitEvent(bs, serial ? plans[i - 1]->tables_evt : lead->pref_fork, 0);
    if (ce != cudaSuccess) break;
    p->capturing = true;
    p->step_mode = true;
    for (int u = 0; u < unroll && rc == 0; ++u) {
      p->iter_tag = u;
      int next_set = -1;
      if (pipeline && u + 1 < unroll) rc = capture_prefetch_next(p, coords_all[i], n, bs, &next_set);
      if (rc == 0)
        rc = npp_train_step(p, coords_all[i], target_all[i], mask_all ? mask_all[i] : nullptr, n, n, lrate, beta1, beta2,
                            eps, first_steps[i], losses[i], bs);
      // the next iteration's encoding has read the step counter before it advances
      if (rc == 0 && next_set >= 0 && cudaStreamWaitEvent(bs, p->pref[next_set].done, 0) != cudaSuccess)
        rc = fail("npp_multi_fit_run: cudaStreamWaitEvent failed");
      if (rc == 0) {
        npp_step_advance_kernel<<<1, 1, 0, bs>>>(p->d_step);
        launches += p->launches + 1;
      }
    }
    p->iter_tag = -1;
    p->pref[0].valid = p->pref[1].valid = false;
    p->step_mode = false;
    p->capturing = false;
    if (rc == 0 && (i > 0 || serial)) {       // join the branch (serial: also the hand-over to the next one)
      ce = cudaEventRecord(p->tables_evt, bs);
      if (ce == cudaSuccess && i > 0) ce = cudaStreamWaitEvent(origin, p->tables_evt, 0);
    }
  }
  cudaGraph_t gr = nullptr;
  const cudaError_t ee = cudaStreamEndCapture(origin, &gr);
  if (rc != 0) {
    if (gr) cudaGraphDestroy(gr);
    return rc;
  }
  if (ce != cudaSuccess || ee != cudaSuccess) {
    if (gr) cudaGraphDestroy(gr);
    return fail(std::string("npp_multi_fit_run: stream capture failed: ") + cudaGetErrorString(ce != cudaSuccess ? ce : ee));
  }
  const cudaError_t ie = cudaGraphInstantiate(&lead->fit_exec, gr, 0ULL);
  cudaGraphDestroy(gr);
  if (ie != cudaSuccess) {
    lead->fit_exec = nullptr;
    return fail(std::string("npp_multi_fit_run: cudaGraphInstantiate failed: ") + cudaGetErrorString(ie));
  }
  lead->fit_stream = st;
  for (int64_t it = 0; it < iters; it += unroll) CK(cudaGraphLaunch(lead->fit_exec, st));
  for (int i = 0; i < k; ++i) {
    plans[i]->step_seq += iters;
    plans[i]->launches = launches / k * (int)(iters / unroll);
    CK(cudaEventRecord(plans[i]->tables_evt, st));
  }
  return 0;
}

int npp_last_launch_count(const NppPlan* p) { return p ? p->launches : 0; }

// Debugging aid for a stuck stream: with npp_profile_enable(plan, 1) every kernel class is bracketed by events; this
// returns the class (PROF_* index) of the first span whose end event has not completed, its ordinal among the spans
// recorded since profiling was enabled, or -1.  Meant to be called from another host thread while the plan's own thread
// is blocked in a CUDA call.
int npp_debug_pending_class(NppPlan* p, int* ordinal, int* total) {
  if (!p) return -1;
  const size_t n = p->spans.size();
  if (total) *total = (int)n;
  for (size_t i = 0; i < n; ++i) {
    const cudaError_t q = cudaEventQuery(p->spans[i].b);
    if (q == cudaErrorNotReady) {
      cudaGetLastError();
      if (ordinal) *ordinal = (int)i;
      return p->spans[i].cls;
    }
  }
  return -1;
}

#ifdef NPP_HANG_DEBUG
// Debug build only: device buffer of 8 plans x 4096 wait-state words (see npp_state_record in ptx_sm100.cuh); plan
// `ordinal` (0..7) of the calling host thread is selected with npp_debug_state_select before its training calls, and
// npp_debug_state_dump copies the whole buffer out on a private stream (it works while kernels hang).
static unsigned long long* g_state_dev = nullptr;
static cudaStream_t g_state_stream = nullptr;
int npp_debug_state_select(int ordinal) {
  if (g_state_dev == nullptr) {
    CK(cudaMalloc(&g_state_dev, 8 * 4096 * sizeof(unsigned long long)));
    CK(cudaMemset(g_state_dev, 0, 8 * 4096 * sizeof(unsigned long long)));
    CK(cudaStreamCreateWithFlags(&g_state_stream, cudaStreamNonBlocking));
    CK(cudaDeviceSynchronize());
  }
  tl_dbg_state = ordinal >= 0 ? g_state_dev + (size_t)(ordinal & 7) * 4096 : nullptr;
  return 0;
}
int npp_debug_state_dump(unsigned long long* host_out) {
  if (g_state_dev == nullptr || host_out == nullptr) return fail("no state buffer");
  CK(cudaMemcpyAsync(host_out, g_state_dev, 8 * 4096 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, g_state_stream));
  CK(cudaStreamSynchronize(g_state_stream));
  return 0;
}
#endif

int npp_set_keep_grads(NppPlan* p, int on) {
  if (!p) return fail("null plan");
  p->keep_grads = on != 0;
  return 0;
}

int npp_profile_enable(NppPlan* p, int on) {
  if (!p) return fail("null plan");
  p->profiling = on != 0;
  p->spans.clear();
  p->ev_used = 0;
  for (int i = 0; i < NPP_PROF_CLASSES; ++i) { p->cls_ms[i] = 0; p->cls_launches[i] = 0; }
  return 0;
}

int npp_profile_read(NppPlan* p, int n_classes, double* ms, int64_t* launches) {
  if (!p || !ms || !launches) return fail("npp_profile_read: null argument");
  CK(cudaDeviceSynchronize());
  for (auto& s : p->spans) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, s.a, s.b));
    p->cls_ms[s.cls] += t;
  }
  p->spans.clear();
  p->ev_used = 0;
  for (int i = 0; i < n_classes && i < NPP_PROF_CLASSES; ++i) { ms[i] = p->cls_ms[i]; launches[i] = p->cls_launches[i]; }
  return 0;
}

__global__ void npp_half_to_float_kernel(const __half* __restrict__ src, int ld, int width, long long n,
                                         float* __restrict__ dst) {
  const long long total = n * width;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / width;
    const int c = (int)(i - r * width);
    dst[i] = __half2float(src[r * ld + c]);
  }
}

static int find_buf(NppPlan* p, const char* name) {
  if (!p || !name) return -1;
  std::string s(name);
  // aliases in reference vocabulary
  const int nl = (int)p->layers.size();
  if (s == "hp") s = "h" + std::to_string(nl - 1);
  for (size_t i = 0; i < p->bufs.size(); ++i)
    if (p->bufs[i].name == s) return (int)i;
  return -1;
}

int npp_debug_width(NppPlan* p, const char* name) {
  const int b = find_buf(p, name);
  return b < 0 ? -1 : p->bufs[b].width;
}

int npp_debug_copy(NppPlan* p, const char* name, int64_t n, float* out, void* stream) {
  const int b = find_buf(p, name);
  if (b < 0) return fail(std::string("unknown buffer ") + (name ? name : "(null)"));
  npp_half_to_float_kernel<<<p->num_sms * 4, 256, 0, (cudaStream_t)stream>>>(p->bufs[b].ptr, p->bufs[b].width,
                                                                            p->bufs[b].width, n, out);
  CK(cudaGetLastError());
  return 0;
}

float npp_debug_grad_scale(NppPlan* p, void* stream) {
  float amax = 0.f;
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaMemcpy(&amax, p->acc + p->amax_off, sizeof(float), cudaMemcpyDeviceToHost);
  if (!(amax > 0.0f) || !std::isfinite(amax)) return 1.0f;
  int e;
  std::frexp(amax, &e);
  int k = 10 - e;
  k = k < -60 ? -60 : (k > 60 ? 60 : k);
  return std::ldexp(1.0f, k);
}

int npp_debug_gemm(const void* a, const void* b, float* c, int m, int n, int k, void* stream) {
  if (n % BN != 0 || k % BK != 0) return fail("npp_debug_gemm: n % 256 == 0 and k % 64 == 0 required");
  KmajorParams kp;
  memset(&kp, 0, sizeof(kp));
  CKI(make_map(&kp.tmA[0], a, m, k, k, BM));
  CKI(make_map(&kp.tmB[0], b, n, k, k, 256));
  kp.nseg = 1;
  kp.kblocks[0] = k / BK;
  kp.M = m;
  kp.tiles_m = (m + BM - 1) / BM;
  kp.tiles_n = n / BN;
  kp.out_f32 = c;
  kp.ldf = n;
  if (const char* e = getenv("NPP_DEBUG_KM_LBO")) {
    const char* s2 = getenv("NPP_DEBUG_KM_SBO");
    kp.desc_hi = umma_desc_hi((uint32_t)atoi(e), s2 ? (uint32_t)atoi(s2) : 1024u);
  }
  if (const char* e = getenv("NPP_DEBUG_KM_KADV")) kp.k_adv = atoi(e);
  // out0 unused: EPI_LINEAR writes fp16 to out0, so route it to a scratch buffer
  __half* scratch = nullptr;
  CK(cudaMalloc(&scratch, (size_t)m * n * sizeof(__half)));
  kp.out0 = scratch;
  kp.ld0 = n;
  CKI(make_map(&kp.tmOut0, scratch, m, n, n, 32));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  kp.epi = EPI_LINEAR;
  KmajorParams* d_op = nullptr;
  CK(cudaMalloc(&d_op, sizeof(KmajorParams)));
  CK(cudaMemcpy(d_op, &kp, sizeof(KmajorParams), cudaMemcpyHostToDevice));
  kp.a_src[0] = kp.a_src[1] = -1;
  {
    std::vector<KmajorParams> one(1, kp);
    finish_chain_ops(one, 1);
    kp = one[0];
  }
  CK(cudaMemcpy(d_op, &kp, sizeof(KmajorParams), cudaMemcpyHostToDevice));
  int r = launch_chain(d_op, &kp, 1, m, sms, (cudaStream_t)stream, (n / BN) * (BN / EPI_COLS));
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(scratch);
  cudaFree(d_op);
  if (r) return r;
  CK(e);
  return 0;
}

// Times `iters` back-to-back launches of the K-major GEMM (epi 0 linear / 1 snake) with CUDA events.
int npp_debug_gemm_bench(const void* a, const void* b, void* out0, void* out1, int m, int n, int k, int epi, int iters,
                         float* ms_out) {
  if (n % BN != 0 || k % BK != 0) return fail("npp_debug_gemm_bench: n % 256 == 0 and k % 64 == 0 required");
  KmajorParams kp;
  memset(&kp, 0, sizeof(kp));
  CKI(make_map(&kp.tmA[0], a, m, k, k, BM));
  int cluster = 1;
  if (const char* e = getenv("NPP_CLUSTER")) cluster = atoi(e) == 2 ? 2 : 1;
  CKI(make_map(&kp.tmB[0], b, n, k, k, 256 / cluster));
  CKI(make_map(&kp.tmOut0, out0, m, n, n, 32));
  if (out1) CKI(make_map(&kp.tmOut1, out1, m, n, n, 32));
  kp.nseg = 1;
  kp.kblocks[0] = k / BK;
  kp.M = m;
  kp.tiles_m = (m + BM - 1) / BM;
  kp.tiles_n = n / BN;
  kp.out0 = (__half*)out0;
  kp.ld0 = n;
  kp.out1 = (__half*)out1;
  kp.ld1 = n;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (const char* e = getenv("NPP_DEBUG_GRID")) sms = atoi(e);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  kp.epi = epi ? EPI_SNAKE : EPI_LINEAR;
  int chain = 1;
  if (const char* e = getenv("NPP_DEBUG_CHAIN")) chain = atoi(e);   // same op repeated `chain` times in one launch
  kp.a_src[0] = kp.a_src[1] = -1;
  std::vector<KmajorParams> ops((size_t)chain, kp);
  const int subs_op = (n / BN) * (BN / EPI_COLS);
  if (getenv("NPP_DEBUG_CHAIN_DEP") && n == k)   // op i reads op i-1's output (ping-pong between out0 and out1 buffers)
    for (int i = 0; i < chain; ++i) {
      ops[i].sub_base = i * subs_op;
      if (i > 0) ops[i].a_src[0] = i - 1;
    }
  else
    for (int i = 0; i < chain; ++i) ops[i].sub_base = i * subs_op;
  finish_chain_ops(ops, cluster);
  KmajorParams* d_op = nullptr;
  CK(cudaMalloc(&d_op, ops.size() * sizeof(KmajorParams)));
  CK(cudaMemcpy(d_op, ops.data(), ops.size() * sizeof(KmajorParams), cudaMemcpyHostToDevice));
  long long* d_dbg = nullptr;
  const int n_stamp_tiles = chain * (n / BN) * ((m + BM - 1) / BM / sms + 1);
  if (getenv("NPP_DEBUG_STAMPS")) {
    CK(cudaMalloc(&d_dbg, (size_t)n_stamp_tiles * 8 * sizeof(long long)));
    CK(cudaMemset(d_dbg, 0, (size_t)n_stamp_tiles * 8 * sizeof(long long)));
  }
  for (int i = 0; i < 3; ++i) CKI(launch_chain(d_op, ops.data(), chain, m, sms, 0, chain * subs_op, cluster));
  CK(cudaEventRecord(e0, 0));
  for (int i = 0; i < iters; ++i) CKI(launch_chain(d_op, ops.data(), chain, m, sms, 0, chain * subs_op, cluster));
  CK(cudaEventRecord(e1, 0));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(ms_out, e0, e1));
  *ms_out /= (float)chain;
  if (d_dbg) {   // one extra stamped launch, printed as per-tile phase durations in clocks
    g_chain_dbg = d_dbg;
    int r2 = launch_chain(d_op, ops.data(), chain, m, sms, 0, chain * subs_op, cluster);
    g_chain_dbg = nullptr;
    cudaDeviceSynchronize();
    if (r2 == 0) {
      std::vector<long long> h((size_t)n_stamp_tiles * 8);
      cudaMemcpy(h.data(), d_dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      const int shown = chain * (n / BN) < n_stamp_tiles ? chain * (n / BN) : n_stamp_tiles;
      const long long t0 = h[0];
      // epilogue warp 2 of CTA 0: enter, accumulator ready, sub-tiles done, tile done, accumulator released;
      // UMMA warp of CTA 0: accumulator free, first K block landed, last K block landed (all clocks since the first stamp)
      for (int i = 0; i < shown; ++i) {
        const long long* q = &h[(size_t)i * 8];
        printf("  tile %2d: epi enter %7lld ready %7lld subs %7lld done %7lld released %7lld | mma free %7lld kb0 %7lld kbN %7lld\n",
               i, q[0] - t0, q[1] - t0, q[2] - t0, q[3] - t0, q[4] - t0, q[5] - t0, q[6] - t0, q[7] - t0);
      }
      fflush(stdout);
    }
    cudaFree(d_dbg);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_op);
  return 0;
}

int npp_debug_wgrad(const void* a, const void* b, float* c, int rows, int m, int n, int splits, void* stream) {
  int wg_cluster = 2;
  if (const char* e = getenv("NPP_WG_CLUSTER")) wg_cluster = atoi(e) == 1 ? 1 : 2;
  if (m % (BM * wg_cluster) != 0 || n % BN != 0 || splits < 1)
    return fail("npp_debug_wgrad: m % 256 == 0 (128 with NPP_WG_CLUSTER=1), n % 256 == 0 required");
  CKI(set_smem_attrs());
  WgradParams w;
  memset(&w, 0, sizeof(w));
  CKI(make_map(&w.maps[0], a, rows, m, m, 64));
  CKI(make_map(&w.maps[1], b, rows, n, n, 64));
  std::vector<WgUnit> units;
  for (int s = 0; s < splits; ++s)
    for (int m0 = 0; m0 < m; m0 += BM * wg_cluster)
      for (int n0 = 0; n0 < n; n0 += BN) {
        WgUnit u;
        u.a_map = 0;
        u.b_map = 1;
        u.a_m0 = m0;
        u.b_n0 = n0;
        u.split = s;
        u.out_off = m0 * n + n0;
        u.ld = n;
        u.ncols_left = n - n0;
        u.bias_off = -1;
        units.push_back(u);
      }
  WgUnit* d_units = nullptr;
  CK(cudaMalloc(&d_units, units.size() * sizeof(WgUnit)));
  CK(cudaMemcpy(d_units, units.data(), units.size() * sizeof(WgUnit), cudaMemcpyHostToDevice));
  const int kb_total = (rows + BK - 1) / BK;
  w.units = d_units;
  w.n_units = (int)units.size();
  w.rows = rows;
  w.kb_per_split = (kb_total + splits - 1) / splits;
  w.n_splits = (kb_total + w.kb_per_split - 1) / w.kb_per_split;
  w.partial = c;
  w.slab_stride = (long long)m * n;
  if (const char* e = getenv("NPP_DEBUG_MN_LBO")) {
    const char* s2 = getenv("NPP_DEBUG_MN_SBO");
    w.desc_hi = umma_desc_hi((uint32_t)atoi(e), s2 ? (uint32_t)atoi(s2) : 1024u);
  }
  if (const char* e = getenv("NPP_DEBUG_MN_KADV")) w.k_adv = atoi(e);
  if (w.n_splits < splits) CK(cudaMemsetAsync(c, 0, (size_t)splits * m * n * sizeof(float), (cudaStream_t)stream));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int r = launch_wgrad(w, sms, (cudaStream_t)stream, wg_cluster);
  cudaError_t e1 = r ? cudaErrorUnknown : cudaGetLastError();
  cudaError_t e2 = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(d_units);
  CK(e1);
  CK(e2);
  return 0;
}

}  // extern "C"
