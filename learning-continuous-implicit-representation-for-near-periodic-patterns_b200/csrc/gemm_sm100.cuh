// tcgen05 / TMEM / TMA GEMM kernels for the NPP-Net coordinate MLP (sm_100a only).
//
// Two persistent, warp-specialised kernels share one pipeline skeleton
//   warp 0      : TMA producer      (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      : UMMA issuer       (tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   warps 2..5  : epilogue          (tcgen05.ld -> registers -> fused math -> HBM)
// with a 4-stage smem ring (full/empty mbarriers) and a 2-stage TMEM accumulator
// ring (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of tile t+1.
//
//  * npp_gemm_kmajor<EPI> : C[M, 256*tiles_n] = [A0 | A1] . [B0 | B1]^T, all operands
//    K-major fp16.  Used for every dense layer of the forward pass
//    (reference models/networks.py:63-94, concat inputs become a second K segment so
//    torch.cat at networks.py:71,76,85 never materialises) and for every dgrad.
//  * npp_gemm_wgrad : dW[out, in] = delta^T . a, contraction over the coordinate rows,
//    both operands MN-major straight out of the row-major activation buffers,
//    split-K over row ranges with deterministic fp32 partial slabs.
#pragma once
#include "ptx_sm100.cuh"
#include "grad_scale.cuh"

namespace npp {

constexpr int BM = 128;      // UMMA M  (rows of the coordinate batch / out-features for wgrad)
constexpr int BN = 256;      // UMMA N
constexpr int BK = 64;       // K per pipeline stage = one 128-byte swizzle atom of halves
constexpr int UMMA_K = 16;
constexpr int STAGES = 2;     // single-CTA fallback of the chain kernel: 2 x 48 KB ring + 96 KB epilogue staging
constexpr int WG_STAGES = 4;  // wgrad kernel: no staging, deeper ring
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int GEMM_THREADS = 320;    // chain kernel: TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quadrant)
constexpr int WGRAD_THREADS = 192;   // wgrad kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;
constexpr int EPI_COLS = 64;                         // epilogue sub-tile: 32 rows x 64 halves = one 4 KB SW128 box per warp
constexpr int EPI_BUF_BYTES = 32 * EPI_COLS * 2;     // 4 KB
constexpr int EPI_OUT_BYTES = 4 * 4 * EPI_BUF_BYTES;          // output staging: 4 sub-tiles x 4 quadrants = one 128 x 256 tile, 64 KB
constexpr int EPI_AUX_BYTES = EPI_WARPS * EPI_BUF_BYTES;      // one box per epilogue warp (snake derivative out / dgrad multiplier in), 32 KB
constexpr int EPI_STAGE_BYTES = EPI_OUT_BYTES + EPI_AUX_BYTES;
constexpr int GEMM_SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + EPI_STAGE_BYTES + BN * 4 + 512 + 1024;
constexpr int STRIPE_GROUP = 1;  // stripes a CTA interleaves op by op (2 was measured neutral: large batches are
                                 // bound by the snake epilogue's MUFU rate, not by dependency bubbles)
constexpr int PAIR_STAGES = 4;  // cta_group::2: a stage is A 16 KB + half of B 16 KB per CTA
constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES / 2) + EPI_STAGE_BYTES + BN * 4 + 512 + 1024;
static_assert(PAIR_SMEM_BYTES <= 232448, "chain kernel exceeds 227 KB of shared memory");
constexpr int WGRAD_SMEM_BYTES = WG_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 512 + 1024;

enum : int {
  EPI_LINEAR = 0,     // out0 = acc + bias                              (feature_linear1/2)
  EPI_SNAKE = 1,      // z = acc + bias; out0 = z + sin^2 z; out1 = 1 + sin 2z   (activations.py:34-35)
  EPI_DGRAD_MUL = 2,  // out0 = acc * mul  (mul = stored snake derivative); colsum += out0
  EPI_DGRAD = 3,      // out0 = acc; colsum += out0
  // Last forward layer of a fused train step: snake, then the RGB head, sigmoid + masked MSE and the head backward
  // without leaving the CTA; out0 = delta of this layer (fp16, scaled), nothing else is written per row.
  EPI_SNAKE_HEAD = 4,
};

// One dense layer (forward) or one dgrad GEMM of the chain; lives in GLOBAL memory (array of ops).
struct alignas(64) KmajorParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB[2];   // weight tiles: box {64, 256/cluster} rows
  CUtensorMap tmOut0;  // box {64, 32} store maps of out0 / out1, load map of mul
  CUtensorMap tmOut1;
  CUtensorMap tmMul;
  int nseg;
  int kblocks[2];  // K blocks (of 64) per segment
  int a_k0[2];     // first K element of the segment inside A's tensor
  int b_k0[2];     // first K element inside B's tensor
  int b_row0[2];   // first row of B (weight shadow) for output column 0
  int a_src[2];    // index of the op of this chain that writes the segment's A tensor, -1 = produced earlier
  int sub_base;    // 64-column output sub-tiles written per stripe by the ops before this one
  int M;           // valid rows
  int tiles_m, tiles_n;
  int epi;         // EPI_*
  const float* bias;
  __half* out0;
  int ld0;
  __half* out1;
  int ld1;
  const __half* mul;
  int ldm;
  float* colsum;
  float* out_f32;  // optional raw fp32 copy of the epilogue value (tests)
  int ldf;
  unsigned long long desc_hi;  // UMMA smem descriptor bits [16,64) (0 = default K-major SW128)
  int k_adv;                   // byte advance per UMMA_K slice (0 = default 32)
  // On-chip forwarding between consecutive ops of a chain: the epilogue of op i writes each 128 x 64 output block
  // straight into the pipeline stage where tile 0 / segment 0 of op i+1 expects that K block of its A operand
  // (same 128-byte-swizzled K-major layout), so the first tile of the next op starts without waiting for
  // TMA store -> L2 -> TMA load.  The copy in global memory is still written (tile 1 of op i+1, the backward pass).
  int fwd_out;      // this op's epilogue forwards into the stages of the next op (needs tiles_n <= 2)
  int fwd_in;       // tile 0 / segment 0 of this op receives its A blocks from the previous op's epilogue
  int kb_per_tile;  // sum of kblocks[] (pipeline stages one tile of this op consumes)
};

// ring position p of a forwarded K block holds block fwd_perm(p): the epilogue finishes sub-tiles in the order
// 0, 2, 1, 3 (two warps per TMEM lane quadrant, two sub-tiles each), so the UMMAs consume them in that order
__host__ __device__ constexpr int fwd_perm(int p) { return (p & ~3) | ((p & 1) << 1) | ((p >> 1) & 1); }

// The scalar part of an op, copied into the kernel parameters (constant bank): the roles read it at every op
// boundary, and a dependent chain of global loads there (~1500 clocks) is exactly where the pipeline has no slack.
constexpr int MAX_CHAIN_OPS = 24;   // forward (12) + dgrad (11) of the joint model in one launch
struct OpScalars {
  int nseg, tiles_n, epi, fwd_in, fwd_out, kb_per_tile;
  int kblocks[2], a_k0[2], b_k0[2], b_row0[2], a_src[2];
  int src_sub_base[2], src_tiles_n[2];  // sub_base / tiles_n of op a_src[s]
  int ldf, k_adv;
  const float* bias;
  float* colsum;
  float* out_f32;
  unsigned long long desc_hi;
};

// Arguments of the EPI_SNAKE_HEAD epilogue (one head per chain).  Reference: rgb_linear (models/networks.py:94),
// sigmoid (models/helpers.py:55-56), img2mse 'l2' with mask (models/mse_calculator.py:13-27) and their autograd.
struct HeadArgs {
  const float* w;         // rgb_linear.weight [3, width] fp32 (master copy)
  const float* b;         // rgb_linear.bias [3]
  const float* target;    // [M, 3]
  const float* mask;      // [M] or nullptr
  float* logits;          // optional [M, 3]
  float* head_acc;        // [3 * width + 3]: dW_rgb, db_rgb (unscaled fp32 atomics)
  float* loss_acc;        // scalar: sum of the squared masked residuals * inv_count
  const unsigned int* amax_prev;  // max |dL/dlogit| / inv_count of the PREVIOUS step (float bits, 0 = unknown)
  unsigned int* amax_next;        // same of this step (atomicMax)
  float inv_count;        // 1 / (3 * n_norm)
  int width;              // reference in_features of rgb_linear (128 or 256)
  // re-launched step graph (npp_multi_fit_run): step index s = *step lives in device memory; this launch works on batch
  // s of [iters, M, 3] targets / [iters, M] masks and on ring slots (seq0 + s) % 3 (amax_prev / amax_next then unused)
  const int* step;
  unsigned int* ring;
  int seq0;
};

// A chain = ops executed in order for every 128-row stripe; op i may read what ops < i wrote for the
// same rows (rows are independent in forward and dgrad), so a CTA that owns a stripe needs no grid sync.
struct ChainParams {
  OpScalars sc[MAX_CHAIN_OPS];
  const KmajorParams* ops;  // device array (tensor maps)
  int n_ops;
  int M;
  int tiles_m;
  int subs_per_stripe;  // output sub-tiles all ops write per stripe
  long long* dbg;       // optional (tests): per-tile clock64 stamps of one CTA (epilogue warp dbg_warp and the UMMA warp of its pair)
  int dbg_block, dbg_warp;
  float* zero_a;        // optional: accumulators (and *zero_b) that CTA 0 clears before the chain starts (fused train step
  int zero_n;           //   whose encoding was prefetched: the encode kernel, which normally does this, did not run)
  float* zero_b;
  int relu;             // host side only: selects the npp_gemm_kmajor<CLUSTER, true> instantiation
  HeadArgs head;        // used by the EPI_SNAKE_HEAD op, if the chain has one
  int pdl;              // launched with programmatic stream serialization: wait for the previous kernel after the prologue
  unsigned long long* dbg_state;   // -DNPP_HANG_DEBUG builds: this plan's wait-state buffer (npp_state_record), else nullptr
};

struct WgUnit {
  short a_map, b_map;  // indices into WgradParams::maps
  int a_m0;            // first out-feature (inner coordinate of the delta tensor)
  int b_n0;            // first in-feature inside the activation tensor
  int split;           // which row range
  int out_off;         // float offset of (m0, col0) inside one partial slab
  int ld;              // row pitch (floats) of this layer inside the slab
  int ncols_left;      // valid columns from col0 to the padded layer width
  int bias_off;        // >= 0: this unit also sums its delta columns over its rows (bias gradient of out-features
                       //   [a_m0, a_m0 + 256)) into WgradParams::bias_acc + bias_off; -1: another unit of the layer does
  int kb0 = 0, kb1 = 0;  // balanced schedule: explicit range of 64-row blocks [kb0, kb1), the partial sums go to slab
                         //   `split`; kb1 == 0: the range of row split `split` (kb_per_split blocks); kb1 < 0: empty slot
};

constexpr int WG_MAX_MAPS = 28;
struct alignas(64) WgradParams {
  CUtensorMap maps[WG_MAX_MAPS];
  const WgUnit* units;
  int n_units;
  int rows;           // contraction length (coordinate rows)
  int kb_per_split;   // 64-row blocks per split
  int n_splits;       // splits actually used for this row count
  float* partial;     // [n_splits][slab_stride]
  long long slab_stride;
  unsigned long long desc_hi;  // 0 = default MN-major SW128 (LBO 8192, SBO 1024)
  int k_adv;                   // 0 = default 2048
  float* bias_acc;             // bias-gradient accumulators (scaled like the deltas), nullptr: nobody wants them
  int grid_pairs;              // > 0: launch exactly this many CTA pairs (CTAs): the unit table is laid out for that stride
  unsigned long long* dbg_state;   // -DNPP_HANG_DEBUG builds: this plan's wait-state buffer, else nullptr
};

template <int NSTAGES>
struct PipeStateT {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance() {
    if (++stage == NSTAGES) {
      stage = 0;
      phase ^= 1;
    }
  }
};
using PipeState = PipeStateT<STAGES>;

struct GemmSmem {
  uint8_t* a;
  uint8_t* b;
  uint8_t* epi;  // 4 warps x 4 x 4 KB staging, 1024-B aligned
  float* bias;   // BN floats, spare (the bias now lives in registers; kept so the barrier block stays 1 KB aligned)
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tfull;
  uint64_t* tempty;
  uint64_t* epi_bar;  // [8] one per epilogue warp (TMA loads of the dgrad multiplier)
  uint64_t* fempty;   // [4] forwarding: staging block b has been consumed by the next op's UMMAs
  uint64_t* afull;    // [8] forwarding: A block k of the next op (sub-tile k & 3 of tile k >> 2) is in the staging
  uint32_t* prog;     // [8] per epilogue warp: output sub-tiles (cumulative) whose TMA stores have completed
  uint32_t* tmem_ptr;
};

template <int NSTAGES, int EPI_BYTES, int A_BYTES = A_STAGE_BYTES, int B_BYTES = B_STAGE_BYTES>
__device__ __forceinline__ GemmSmem carve_smem_t(uint8_t* raw) {
  uint32_t addr = smem_u32(raw);
  uint8_t* base = raw + ((1024u - (addr & 1023u)) & 1023u);
  GemmSmem s;
  s.a = base;
  s.b = base + NSTAGES * A_BYTES;
  s.epi = s.b + NSTAGES * B_BYTES;
  s.bias = reinterpret_cast<float*>(s.epi + EPI_BYTES);
  s.full = reinterpret_cast<uint64_t*>(s.bias + (EPI_BYTES ? BN : 0));
  s.empty = s.full + NSTAGES;
  s.tfull = s.empty + NSTAGES;
  s.tempty = s.tfull + 2;
  s.epi_bar = s.tempty + 2;
  s.fempty = s.epi_bar + 8;
  s.afull = s.fempty + 4;
  s.prog = reinterpret_cast<uint32_t*>(s.afull + 8);
  s.tmem_ptr = s.prog + 8;
  return s;
}

template <int NSTAGES, int CLUSTER = 1, int NEPI = 4, int FULL_COUNT = 1, int EMPTY_COUNT = 1, int AFULL_COUNT = 0>
__device__ __forceinline__ uint32_t gemm_prologue_t(const GemmSmem& s, int warp) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGES; ++i) {
      mbar_init(&s.full[i], FULL_COUNT);
      mbar_init(&s.empty[i], EMPTY_COUNT);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s.tfull[i], 1);
      mbar_init(&s.tempty[i], NEPI * CLUSTER);  // one arrive per epilogue warp of every CTA feeding this accumulator
    }
    for (int i = 0; i < 8; ++i) mbar_init(&s.epi_bar[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&s.fempty[i], 1);
    for (int i = 0; i < 8; ++i)   // chain kernel: 4 quadrant warps of every CTA
      mbar_init(&s.afull[i], AFULL_COUNT ? AFULL_COUNT : (NEPI / 2) * CLUSTER);
    for (int i = 0; i < 8; ++i) s.prog[i] = 0;
    fence_mbar_init();
  }
  if (warp == 1) {
    if (CLUSTER == 1) tmem_alloc(s.tmem_ptr, TMEM_COLS);
    else tmem_alloc_pair(s.tmem_ptr, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();  // the peer's barriers exist before anything can signal them
  tc_fence_after();
  return *s.tmem_ptr;
}

template <int CLUSTER = 1>
__device__ __forceinline__ void gemm_teardown(uint32_t tmem_base, int warp) {
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();  // no CTA may exit while its peer can still signal its barriers / read its smem
  if (warp == 1) {
    tc_fence_after();
    if (CLUSTER == 1) tmem_dealloc(tmem_base, TMEM_COLS);
    else tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// 32x32 transpose-reduce: on return lane j holds sum over the 32 lanes of v[j].
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      float send = up ? v[i] : v[i + s];
      float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// same, clamped to +-65504 instead of overflowing to infinity (scaled deltas: one clipped step beats NaN weights)
__device__ __forceinline__ uint32_t pack_h2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// Progress counters between the epilogue warps (writers of an op's output) and the TMA producer (reader of it
// as the next op's A operand).  The value is the cumulative number of 64-column sub-tiles whose bulk stores have
// COMPLETED; release/acquire at CTA scope plus an async-proxy fence on both sides order TMA store -> TMA load.
__device__ __forceinline__ void publish_progress(uint32_t* slot, uint32_t value) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(value) : "memory");
}
__device__ __forceinline__ uint32_t wait_progress(const uint32_t* slot, uint32_t need) {
  uint32_t v;
#ifdef NPP_HANG_DEBUG
  npp_state_record(9000, smem_u32(slot), need, 0);
#endif
  do {
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(slot)) : "memory");
#ifdef NPP_SPIN_SLEEP_NS
    if (static_cast<int32_t>(v - need) < 0) __nanosleep(NPP_SPIN_SLEEP_NS);
#endif
  } while (static_cast<int32_t>(v - need) < 0);
#ifdef NPP_HANG_DEBUG
  npp_state_record(9000, smem_u32(slot), need, 1);
#endif
  return v;
}

// ---------------------------------------------------------------------------------
// Epilogue of one 128x256 accumulator tile.  Each warp owns 32 accumulator rows (its TMEM lane quadrant).
// Results are staged in a per-warp 128B-swizzled 32x64 smem box and written with TMA stores (full 128-byte
// lines); the dgrad multiplier (stored snake derivative) arrives the same way through a TMA load.
// scalar fields of an op, copied to registers once per op (the descriptor itself lives in global memory)
struct EpiArgs {
  float4 bias4;     // bias of the warp's 128 output columns, 4 per lane (columns 128 e + 4 lane ..), prefetched a tile ahead
  float* colsum;
  float* out_f32;
  int ldf;
  const CUtensorMap* next_mul;  // dgrad * d: multiplier map of the tile that follows in this stripe (nullptr: none)
  int next_mul_col;             // its first column
  int fwd_out;      // the next op's first tile reads this op's output staging as its A blocks
  int nt;           // N-tile index inside the op
  bool leader;      // this CTA owns the UMMA-side barriers (always true without clusters)
};

// RELU: EPI_SNAKE ops apply F.relu instead (activation != 'snake', models/networks.py:51-54,66-69).  A template
// parameter of the kernel, not a run-time flag: a warp-uniform branch inside the snake epilogue made ptxas spill
// (200 B of stores per thread) and cost the snake path 9 % of its forward time.
template <int EPI, bool RELU>
__device__ __forceinline__ void epilogue_tile(const KmajorParams& p, const EpiArgs ea, const GemmSmem& s,
                                              uint32_t tmem_acc, int m0,
                                              int n0, int M, int warp, int lane, uint32_t& ld_phase,
                                              uint64_t* tfull, uint32_t acc_phase, uint32_t& seq,
                                              bool last_tile_of_op, uint32_t& deferred_seq, long long* dbg,
                                              uint32_t& fwd_phase, bool& mul_ready) {
  constexpr int NSUB = BN / EPI_COLS;  // 4 sub-tiles of 64 columns
  const bool stamp = dbg != nullptr && lane == 0;
  if (stamp) dbg[0] = clock64();
  // Two warps share each TMEM lane quadrant (warps 2..5 take sub-tiles 0 and 1, warps 6..9 take 2 and 3), so the
  // latency chain of one sub-tile (TMEM load -> math -> fence -> TMA store) overlaps the other warp's work.
  const int q = warp & 3;
  const int e = (warp - 2) >> 2;
  const int lane_base = q * 32;
  // Output staging: block `sub` (16 KB) holds the 128 x 64 output sub-tile `sub` of the CTA in the 128-byte-swizzled
  // K-major layout of a UMMA A operand (quadrant q = rows [32 q, 32 q + 32) at + 4 KB q).  The TMA store reads it
  // from there, and when the op forwards (fwd_out) the next op's first tile reads it as its A block, no copy.
  // Aux: one 4 KB box per warp: the snake derivative on its way out, or the dgrad multiplier on its way in.
  uint8_t* aux = s.epi + EPI_OUT_BYTES + (warp - 2) * EPI_BUF_BYTES;
  uint64_t* ebar = &s.epi_bar[warp - 2];
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
  const uint32_t row_off = static_cast<uint32_t>(lane) * 128u;
  const int row0 = m0 + lane_base;
  const int row = row0 + lane;
  const bool row_ok = row < M;
  const bool warp_ok = row0 < M;
  (void)ea.colsum;   // bias gradients are summed by the weight-gradient kernel (WgUnit::bias_off)
  float* out_f32 = ea.out_f32;
  if (EPI == EPI_DGRAD_MUL) {
    // Multiplier of this warp's first sub-tile.  Normally the previous tile already asked for it (right after it had
    // read its own last multiplier out of the aux box), so that the load has a whole sub-tile of work to land.
    if (!mul_ready && lane == 0) {
      mbar_expect_tx(ebar, EPI_BUF_BYTES);
      tma_load_2d(aux, &p.tmMul, ebar, n0 + 2 * e * EPI_COLS, row0);
    }
  }
  mul_ready = false;
  mbar_wait(tfull, acc_phase);
  tc_fence_after();
  if (stamp) dbg[1] = clock64();
#pragma unroll 1
  for (int sub = 2 * e; sub < 2 * e + 2; ++sub) {
    const int col = n0 + sub * EPI_COLS;
    uint8_t* obuf = s.epi + sub * (4 * EPI_BUF_BYTES) + q * EPI_BUF_BYTES;
    // both 32-column halves of the sub-tile are fetched from TMEM before either is consumed
    uint32_t raw0[32], raw1[32];
    tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(lane_base) << 16) + sub * EPI_COLS, raw0);
    tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(lane_base) << 16) + sub * EPI_COLS + 32, raw1);
    tmem_ld_wait();
    if (EPI == EPI_DGRAD_MUL) {
      mbar_wait(ebar, ld_phase);
      ld_phase ^= 1;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(half == 0 ? raw0[i] : raw1[i]);
      const int hcol = col + half * 32;

      if (EPI == EPI_LINEAR || EPI == EPI_SNAKE) {
        // The warp's 128 bias values live in registers, four per lane; a column's value comes by shuffle.  (Staging
        // the slice in shared memory needed two 256-thread barriers per tile, which made every epilogue warp wait
        // for the slowest one: 1300-1900 clocks per tile in the snake layers.)
        const int src0 = (sub - 2 * e) * 16 + half * 8;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          v[i] += __shfl_sync(0xffffffffu, ea.bias4.x, src0 + (i >> 2));
          v[i + 1] += __shfl_sync(0xffffffffu, ea.bias4.y, src0 + (i >> 2));
          v[i + 2] += __shfl_sync(0xffffffffu, ea.bias4.z, src0 + (i >> 2));
          v[i + 3] += __shfl_sync(0xffffffffu, ea.bias4.w, src0 + (i >> 2));
        }
      }
      if (out_f32 != nullptr && row_ok) {
        float4* o = reinterpret_cast<float4*>(out_f32 + (size_t)row * ea.ldf + hcol);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      uint32_t hd[16], dd[16];
      if (EPI == EPI_SNAKE && RELU) {
        // F.relu(h) (networks.py:66-67): out0 = max(z, 0), out1 = d/dz = [z > 0] (torch's relu backward)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          hd[i >> 1] = pack_h2(fmaxf(v[i], 0.f), fmaxf(v[i + 1], 0.f));
          dd[i >> 1] = pack_h2(v[i] > 0.f ? 1.f : 0.f, v[i + 1] > 0.f ? 1.f : 0.f);
        }
      } else if (EPI == EPI_SNAKE) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float h2[2], d2[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const float z = v[i + k];
            const float w = z + z;
            const float sn = __sinf(w);
            const float cs = __cosf(w);
            h2[k] = fmaf(-0.5f, cs, z + 0.5f);  // z + sin^2 z = z + (1 - cos 2z)/2
            d2[k] = 1.0f + sn;                  // d/dz = 1 + sin 2z
          }
          hd[i >> 1] = pack_h2(h2[0], h2[1]);
          dd[i >> 1] = pack_h2(d2[0], d2[1]);
        }
      } else {
        if (EPI == EPI_DGRAD_MUL) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t c = static_cast<uint32_t>(half * 4 + j);
            const uint4 m4 = *reinterpret_cast<const uint4*>(aux + row_off + ((c ^ sw) << 4));
            const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 m2 = unpack_h2(mw[t]);
              v[8 * j + 2 * t] *= m2.x;
              v[8 * j + 2 * t + 1] *= m2.y;
            }
          }
          if (half == 1) {
            // the whole box has been read: fetch the next multiplier into it while this sub-tile is packed and stored
            // (the second sub-tile of this tile, or the first one of the tile that follows)
            __syncwarp();
            if (sub == 2 * e) {
              if (lane == 0) {
                mbar_expect_tx(ebar, EPI_BUF_BYTES);
                tma_load_2d(aux, &p.tmMul, ebar, col + EPI_COLS, row0);
              }
            } else if (ea.next_mul != nullptr) {
              if (lane == 0) {
                mbar_expect_tx(ebar, EPI_BUF_BYTES);
                tma_load_2d(aux, ea.next_mul, ebar, ea.next_mul_col + 2 * e * EPI_COLS, row0);
              }
              mul_ready = true;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
          hd[i] = (EPI == EPI_DGRAD || EPI == EPI_DGRAD_MUL) ? pack_h2_sat(v[2 * i], v[2 * i + 1]) : pack_h2(v[2 * i], v[2 * i + 1]);
      }
      if (half == 0) {
        // Every sub-tile is one bulk group (out0 from staging block `sub`, snake: plus the derivative from the aux
        // box).  Before overwriting a buffer the group that read it must have finished reading: block `sub` was read
        // two groups ago (same sub-tile of the previous tile), the aux box by the previous group.  Asked here, a whole
        // sub-tile of math after those stores were issued.
        if (lane == 0) {
          if (EPI == EPI_SNAKE) bulk_wait_read0();
          else bulk_wait_read1();
        }
        __syncwarp();
        if (ea.fwd_out && ea.nt > 0) {
          // Staging block `sub` still holds the previous tile's sub-tile, which the NEXT op's first tile reads as an
          // A block: wait until its UMMAs have consumed it (one commit per block on the class barrier, so this warp
          // sees every phase of the two blocks it owns).  Tile 0 needs no such wait: the previous occupant fed the
          // tile whose accumulator this warp is draining.
          mbar_wait(&s.fempty[sub], (fwd_phase >> (sub & 1)) & 1u);
          fwd_phase ^= 1u << (sub & 1);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t c = static_cast<uint32_t>(half * 4 + j);
        *reinterpret_cast<uint4*>(obuf + row_off + ((c ^ sw) << 4)) =
            make_uint4(hd[4 * j], hd[4 * j + 1], hd[4 * j + 2], hd[4 * j + 3]);
        if (EPI == EPI_SNAKE)
          *reinterpret_cast<uint4*>(aux + row_off + ((c ^ sw) << 4)) =
              make_uint4(dd[4 * j], dd[4 * j + 1], dd[4 * j + 2], dd[4 * j + 3]);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (ea.fwd_out) {  // this warp's 32 x 64 part of A block (4 nt + sub) of the next op is in place
        uint64_t* afull = &s.afull[4 * ea.nt + sub];
        if (ea.leader) mbar_arrive(afull);
        else mbar_arrive_remote_relaxed(afull, 0);
      }
      if (warp_ok) {
        tma_store_2d(&p.tmOut0, obuf, col, row0);
        if (EPI == EPI_SNAKE) tma_store_2d(&p.tmOut1, aux, col, row0);
        bulk_commit();
      }
      if (deferred_seq != 0) {
        // the previous tile of this op skipped its drain: by now its stores (every group but the one just
        // committed) have long completed, so publishing them costs nothing
        if (warp_ok) bulk_wait1();
        publish_progress(&s.prog[warp - 2], deferred_seq);
      }
    }
    deferred_seq = 0;
  }
  seq += NSUB;  // progress is published per tile: all sub-tiles (of both warps of the quadrant) up to here
  if (stamp) dbg[2] = clock64();
  if (last_tile_of_op) {
    // the next op's second tile reloads this op's output from global memory: drain the stores (~1 us) and publish
    if (lane == 0) {
      bulk_wait0();
      publish_progress(&s.prog[warp - 2], seq);
    }
    __syncwarp();
  } else {
    deferred_seq = seq;  // published from inside the next tile of the same op (nobody is waiting for it yet)
  }
  if (stamp) dbg[3] = clock64();
}

// ---------------------------------------------------------------------------------
// EPI_SNAKE_HEAD: epilogue of the LAST dense layer (pos_linears.0, one 256-column tile) of a fused train step.  The
// accumulator never leaves the CTA as an activation: the snake output feeds the RGB head (models/networks.py:94), the
// sigmoid (models/helpers.py:55-56) and the masked 'l2' loss (models/mse_calculator.py:13-27) right here, and the
// backward of all three produces this layer's delta, which is staged exactly like any other op's output (TMA store
// to the delta buffer + on-chip forwarding into the first dgrad GEMM).  Two passes over the TMEM accumulator:
//   1. h = snake(z) rounded to fp16 (what the stand-alone head kernel would read back), partial logits over this
//      warp's 128 columns; the two warps of a lane quadrant swap their partials through their aux boxes;
//      loss, dL/dlogit, max |dL/dlogit|, db_rgb;
//   2. h and snake'(z) again, delta = (g . W_rgb) * snake'(z) * scale (fp16), dW_rgb and the layer's bias gradient
//      as column sums.
// Neither h_P nor its derivative is written to global memory (the stand-alone path wrote and re-read 2 x 8 MB).
template <bool RELU>
__device__ __forceinline__ void epilogue_head_tile(const KmajorParams& p, const EpiArgs ea, const HeadArgs& hd,
                                                   const GemmSmem& s, uint32_t tmem_acc, int m0, int M, int warp,
                                                   int lane, uint64_t* tfull, uint32_t acc_phase, uint32_t& seq,
                                                   uint32_t& deferred_seq) {
  constexpr int NSUB = BN / EPI_COLS;
  const int q = warp & 3;
  const int e = (warp - 2) >> 2;
  const int lane_base = q * 32;
  uint8_t* aux = s.epi + EPI_OUT_BYTES + (warp - 2) * EPI_BUF_BYTES;
  const uint8_t* aux_peer = s.epi + EPI_OUT_BYTES + ((warp - 2) ^ 4) * EPI_BUF_BYTES;
  const uint32_t sw = static_cast<uint32_t>(lane & 7);
  const uint32_t row_off = static_cast<uint32_t>(lane) * 128u;
  const int row0 = m0 + lane_base;
  const int row = row0 + lane;
  const bool row_ok = row < M;
  const bool warp_ok = row0 < M;
  const int width = hd.width;
  const uint32_t tlane = tmem_acc + (static_cast<uint32_t>(lane_base) << 16);
  const float* target = hd.target;
  const float* maskp = hd.mask;
  const unsigned int* amax_prev = hd.amax_prev;
  unsigned int* amax_next = hd.amax_next;
  if (hd.step != nullptr) {
    const int sidx = *hd.step;
    target += (size_t)sidx * M * 3;
    if (maskp != nullptr) maskp += (size_t)sidx * M;
    const int slot = (hd.seq0 + sidx) % 3;
    amax_next = hd.ring + slot;
    amax_prev = hd.ring + (slot + 2) % 3;
  }
  // inputs of the loss: requested before the accumulator is waited for
  float tg[3] = {0.f, 0.f, 0.f};
  float mk = 1.0f;
  if (row_ok) {
    tg[0] = __ldg(target + 3 * (size_t)row);
    tg[1] = __ldg(target + 3 * (size_t)row + 1);
    tg[2] = __ldg(target + 3 * (size_t)row + 2);
    if (maskp != nullptr) mk = __ldg(maskp + row);
  }
  const float rb0 = __ldg(hd.b), rb1 = __ldg(hd.b + 1), rb2 = __ldg(hd.b + 2);
  const float scale = npp_grad_scale(npp_step_amax(amax_prev, hd.inv_count));
  // This warp's aux box carries its partial logits to the other warp of the quadrant: every bulk store that read the
  // box (the snake derivative of an earlier layer) and every earlier store from the output staging has finished reading.
  if (lane == 0) bulk_wait_read0();
  __syncwarp();
  mbar_wait(tfull, acc_phase);
  tc_fence_after();

  // ---- pass 1: partial logits over columns [128 e, 128 e + 128)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 1
  for (int blk = 0; blk < 4; ++blk) {
    const int col = e * 128 + blk * 32;
    if (col >= width) break;   // padded half of a 128-wide layer: zero weights, nothing to add
    uint32_t raw[32];
    tmem_ld_32x32(tlane + col, raw);
    tmem_ld_wait();
    const float4* w0p = reinterpret_cast<const float4*>(hd.w + col);
    const float4* w1p = reinterpret_cast<const float4*>(hd.w + width + col);
    const float4* w2p = reinterpret_cast<const float4*>(hd.w + 2 * width + col);
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 w0 = __ldg(w0p + (i >> 2)), w1 = __ldg(w1p + (i >> 2)), w2 = __ldg(w2p + (i >> 2));
      float z[4];
      z[0] = __uint_as_float(raw[i]) + __shfl_sync(0xffffffffu, ea.bias4.x, blk * 8 + (i >> 2));
      z[1] = __uint_as_float(raw[i + 1]) + __shfl_sync(0xffffffffu, ea.bias4.y, blk * 8 + (i >> 2));
      z[2] = __uint_as_float(raw[i + 2]) + __shfl_sync(0xffffffffu, ea.bias4.z, blk * 8 + (i >> 2));
      z[3] = __uint_as_float(raw[i + 3]) + __shfl_sync(0xffffffffu, ea.bias4.w, blk * 8 + (i >> 2));
      float h[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (RELU) h[k] = fmaxf(z[k], 0.f);
        else h[k] = fmaf(-0.5f, __cosf(z[k] + z[k]), z[k] + 0.5f);
      }
      const float2 f0 = unpack_h2(pack_h2(h[0], h[1])), f1 = unpack_h2(pack_h2(h[2], h[3]));
      a0 = fmaf(f0.x, w0.x, fmaf(f0.y, w0.y, fmaf(f1.x, w0.z, fmaf(f1.y, w0.w, a0))));
      a1 = fmaf(f0.x, w1.x, fmaf(f0.y, w1.y, fmaf(f1.x, w1.z, fmaf(f1.y, w1.w, a1))));
      a2 = fmaf(f0.x, w2.x, fmaf(f0.y, w2.y, fmaf(f1.x, w2.z, fmaf(f1.y, w2.w, a2))));
    }
  }
  *reinterpret_cast<float4*>(aux + lane * 16) = make_float4(a0, a1, a2, 0.f);
  named_bar_sync(1 + q, 64);
  const float4 o = *reinterpret_cast<const float4*>(aux_peer + lane * 16);
  named_bar_sync(1 + q, 64);   // both boxes have been read before either is reused
  // (fp32 addition commutes: both warps of the quadrant hold bit-identical logits)
  const float z0 = (a0 + o.x) + rb0, z1 = (a1 + o.y) + rb1, z2 = (a2 + o.z) + rb2;
  float g0 = 0.f, g1 = 0.f, g2 = 0.f;
  {
    const float wgt = mk + (1.0f - mk) * 0.3f;
    const float y0 = 1.0f / (1.0f + expf(-z0)), y1 = 1.0f / (1.0f + expf(-z1)), y2 = 1.0f / (1.0f + expf(-z2));
    const float d0 = (y0 - tg[0]) * wgt, d1 = (y1 - tg[1]) * wgt, d2 = (y2 - tg[2]) * wgt;
    float lsum = 0.f;
    if (row_ok) {
      g0 = 2.0f * d0 * wgt * hd.inv_count * y0 * (1.0f - y0);
      g1 = 2.0f * d1 * wgt * hd.inv_count * y1 * (1.0f - y1);
      g2 = 2.0f * d2 * wgt * hd.inv_count * y2 * (1.0f - y2);
      lsum = d0 * d0 + d1 * d1 + d2 * d2;
    }
    if (e == 0) {   // one warp of the quadrant accounts for the row
      if (hd.logits != nullptr && row_ok) {
        hd.logits[3 * (size_t)row] = z0;
        hd.logits[3 * (size_t)row + 1] = z1;
        hd.logits[3 * (size_t)row + 2] = z2;
      }
      float lmax = fmaxf(fabsf(g0), fmaxf(fabsf(g1), fabsf(g2)));
      float s0 = g0, s1 = g1, s2 = g2;
#pragma unroll
      for (int sft = 16; sft >= 1; sft >>= 1) {
        lsum += __shfl_xor_sync(0xffffffffu, lsum, sft);
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, sft));
        s0 += __shfl_xor_sync(0xffffffffu, s0, sft);
        s1 += __shfl_xor_sync(0xffffffffu, s1, sft);
        s2 += __shfl_xor_sync(0xffffffffu, s2, sft);
      }
      if (lane == 0 && warp_ok) {
        atomicAdd(hd.loss_acc, lsum * hd.inv_count);
        if (lmax > 0.f) atomicMax(amax_next, __float_as_uint(lmax / hd.inv_count));
        atomicAdd(hd.head_acc + 3 * width, s0);      // db_rgb
        atomicAdd(hd.head_acc + 3 * width + 1, s1);
        atomicAdd(hd.head_acc + 3 * width + 2, s2);
      }
    }
  }

  // ---- pass 2: delta of this layer, dW_rgb, bias gradient
#pragma unroll 1
  for (int sub = 2 * e; sub < 2 * e + 2; ++sub) {
    uint8_t* obuf = s.epi + sub * (4 * EPI_BUF_BYTES) + q * EPI_BUF_BYTES;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int col = sub * EPI_COLS + half * 32;
      const bool col_ok = col < width;
      uint32_t raw[32];
      tmem_ld_32x32(tlane + col, raw);
      tmem_ld_wait();
      const int src0 = (sub - 2 * e) * 16 + half * 8;
      const float4* w0p = reinterpret_cast<const float4*>(hd.w + col);
      const float4* w1p = reinterpret_cast<const float4*>(hd.w + width + col);
      const float4* w2p = reinterpret_cast<const float4*>(hd.w + 2 * width + col);
      uint32_t hp[16], dl[16];
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 w0 = col_ok ? __ldg(w0p + (i >> 2)) : zero4;
        const float4 w1 = col_ok ? __ldg(w1p + (i >> 2)) : zero4;
        const float4 w2 = col_ok ? __ldg(w2p + (i >> 2)) : zero4;
        float z[4];
        z[0] = __uint_as_float(raw[i]) + __shfl_sync(0xffffffffu, ea.bias4.x, src0 + (i >> 2));
        z[1] = __uint_as_float(raw[i + 1]) + __shfl_sync(0xffffffffu, ea.bias4.y, src0 + (i >> 2));
        z[2] = __uint_as_float(raw[i + 2]) + __shfl_sync(0xffffffffu, ea.bias4.z, src0 + (i >> 2));
        z[3] = __uint_as_float(raw[i + 3]) + __shfl_sync(0xffffffffu, ea.bias4.w, src0 + (i >> 2));
        float h[4], d[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (RELU) {
            h[k] = fmaxf(z[k], 0.f);
            d[k] = z[k] > 0.f ? 1.f : 0.f;
          } else {
            const float w = z[k] + z[k];
            h[k] = fmaf(-0.5f, __cosf(w), z[k] + 0.5f);
            d[k] = 1.0f + __sinf(w);
          }
        }
        hp[i >> 1] = pack_h2(h[0], h[1]);
        hp[(i >> 1) + 1] = pack_h2(h[2], h[3]);
        // the derivative rounded to fp16, as the stand-alone path stores it between forward and backward
        const float2 dd0 = unpack_h2(pack_h2(d[0], d[1])), dd1 = unpack_h2(pack_h2(d[2], d[3]));
        const float t0 = fmaf(g0, w0.x, fmaf(g1, w1.x, g2 * w2.x));
        const float t1 = fmaf(g0, w0.y, fmaf(g1, w1.y, g2 * w2.y));
        const float t2 = fmaf(g0, w0.z, fmaf(g1, w1.z, g2 * w2.z));
        const float t3 = fmaf(g0, w0.w, fmaf(g1, w1.w, g2 * w2.w));
        dl[i >> 1] = pack_h2_sat(t0 * dd0.x * scale, t1 * dd0.y * scale);
        dl[(i >> 1) + 1] = pack_h2_sat(t2 * dd1.x * scale, t3 * dd1.y * scale);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t c = static_cast<uint32_t>(half * 4 + j);
        *reinterpret_cast<uint4*>(obuf + row_off + ((c ^ sw) << 4)) =
            make_uint4(dl[4 * j], dl[4 * j + 1], dl[4 * j + 2], dl[4 * j + 3]);
      }
      float v[32];
      if (ea.colsum != nullptr) {   // bias gradient of this layer: column sums of the fp16-rounded deltas
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 r = unpack_h2(dl[i]);
          v[2 * i] = r.x;
          v[2 * i + 1] = r.y;
        }
        const float cs = warp_colsum32(v, lane);
        if (warp_ok) atomicAdd(ea.colsum + col + lane, cs);
      }
      if (col_ok) {                 // dW_rgb[c, col + j] = sum over rows of g_c * h[row, col + j]
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          const float gc = c == 0 ? g0 : (c == 1 ? g1 : g2);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 r = unpack_h2(hp[i]);
            v[2 * i] = gc * r.x;
            v[2 * i + 1] = gc * r.y;
          }
          const float cs = warp_colsum32(v, lane);
          if (warp_ok) atomicAdd(hd.head_acc + c * width + col + lane, cs);
        }
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (ea.fwd_out) {
        uint64_t* afull = &s.afull[sub];
        if (ea.leader) mbar_arrive(afull);
        else mbar_arrive_remote_relaxed(afull, 0);
      }
      if (warp_ok) {
        tma_store_2d(&p.tmOut0, obuf, sub * EPI_COLS, row0);
        bulk_commit();
      }
      if (deferred_seq != 0) {
        if (warp_ok) bulk_wait1();
        publish_progress(&s.prog[warp - 2], deferred_seq);
      }
    }
    deferred_seq = 0;
  }
  seq += NSUB;
  if (lane == 0) {   // the first dgrad op's second tile (if any) reloads this delta from global memory
    bulk_wait0();
    publish_progress(&s.prog[warp - 2], seq);
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------
// CLUSTER == 1: every CTA is on its own (UMMA cta_group::1, M = 128).
// CLUSTER == 2: CTA pairs (thread-block cluster of 2 = one TPC) run cta_group::2 UMMAs with M = 256: CTA r owns
//   row stripe 2p + r, keeps its own A tile and HALF of every weight tile in its shared memory, and the leader
//   (rank 0) issues the MMAs for both.  Per SM that halves the B bytes written by TMA and read by the tensor core,
//   which is what lifts the shared-memory-port ceiling of the single-CTA SS-mode mainloop.
template <int CLUSTER, bool RELU = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) npp_gemm_kmajor(const __grid_constant__ ChainParams cp) {
  constexpr int NST = CLUSTER == 1 ? STAGES : PAIR_STAGES;
  constexpr int B_BYTES = B_STAGE_BYTES / CLUSTER;       // this CTA's share of a weight tile
  constexpr int B_ROWS = BN / CLUSTER;
  extern __shared__ uint8_t smem_raw[];
  const GemmSmem s = carve_smem_t<NST, EPI_STAGE_BYTES, A_STAGE_BYTES, B_BYTES>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef NPP_PDL_WAIT_FIRST
  grid_dependency_wait();   // experiment: an early-launched CTA holds no TMEM / barrier while it waits
#endif
#ifdef NPP_HANG_DEBUG
  if (threadIdx.x == 0) *npp_state_slot() = cp.dbg_state;   // visible to all warps after the prologue's barrier
#endif
  const uint32_t tmem_base = gemm_prologue_t<NST, CLUSTER, EPI_WARPS>(s, warp);
  // Programmatic dependent launch: everything above (barriers, TMEM, cluster handshake) may overlap the tail of the
  // previous kernel in the stream; nothing below may (a no-op for an ordinary launch).
  grid_dependency_wait();
  if (blockIdx.x == 0 && cp.zero_n > 0) {
    for (int i = threadIdx.x; i < cp.zero_n; i += blockDim.x) cp.zero_a[i] = 0.f;
    if (threadIdx.x == 0 && cp.zero_b != nullptr) *cp.zero_b = 0.f;
  }
  // every CTA of a pair runs the same number of stripe iterations (phantom stripes load zeros, store nothing)
  const int stripe_iters = (cp.tiles_m + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t crank = CLUSTER > 1 ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (every CTA)
    PipeStateT<NST> ps;
    uint32_t seen = 0;        // cumulative output sub-tiles known to be complete (minimum over the epilogue warps)
    uint32_t group_base = 0;  // sub-tiles published by the stripe groups before the current one
    for (int si = 0; si < stripe_iters; si += STRIPE_GROUP) {
      // A CTA that owns several stripes walks them two at a time, op by op (op -> stripe -> tile): while one
      // stripe waits for its previous op's last tile to be stored and reloaded, the other one's MMAs run.
      const int gi = min(STRIPE_GROUP, stripe_iters - si);
      for (int oi = 0; oi < cp.n_ops; ++oi) {
        const KmajorParams& p = cp.ops[oi];
        // the next kernel of the stream may start its prologue while this CTA works on its last op
        if (oi == cp.n_ops - 1 && si + gi >= stripe_iters && lane == 0) grid_launch_dependents();
        if (lane == 0) {
          for (int i = 0; i < cp.sc[oi].nseg; ++i) {
            tma_prefetch_desc(&p.tmA[i]);
            tma_prefetch_desc(&p.tmB[i]);
          }
        }
        const OpScalars& sc = cp.sc[oi];
        const int nseg = sc.nseg;
        const int op_tiles_n = sc.tiles_n;
        const bool fwd_in = sc.fwd_in != 0;
        int r_kbs[2], r_ak0[2], r_bk0[2], r_br0[2], r_src[2];
        uint32_t r_srcbase[2], r_srctiles[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          r_kbs[q] = q < nseg ? sc.kblocks[q] : 0;
          r_ak0[q] = sc.a_k0[q];
          r_bk0[q] = sc.b_k0[q];
          r_br0[q] = sc.b_row0[q];
          r_src[q] = q < nseg ? sc.a_src[q] : -1;
          r_srcbase[q] = r_src[q] >= 0 ? (uint32_t)sc.src_sub_base[q] : 0u;
          r_srctiles[q] = r_src[q] >= 0 ? (uint32_t)sc.src_tiles_n[q] : 0u;
        }
        for (int sl = 0; sl < gi; ++sl) {
          const int m0 = ((si + sl) * (int)gridDim.x + (int)blockIdx.x) * BM;
          for (int nt = 0; nt < op_tiles_n; ++nt) {
            const int n0 = nt * BN;
#pragma unroll
            for (int seg = 0; seg < 2; ++seg) {
              if (seg >= nseg) break;
              const int kbs = r_kbs[seg], ak0 = r_ak0[seg], bk0 = r_bk0[seg], br0 = r_br0[seg];
              const int src = r_src[seg];
              // K block kb of this segment is the 64-column block (ak0/64 + kb) that op `src` wrote for the same
              // stripe; in processing order that op's tiles come after gi * sub_base[src] sub-tiles of this group
              // and after the sl earlier stripes of the op
              uint32_t need0 = 0;
              if (src >= 0)
                need0 = group_base + (uint32_t)gi * r_srcbase[seg] +
                        (uint32_t)sl * (r_srctiles[seg] * (uint32_t)(BN / EPI_COLS)) + ak0 / BK + 1;
              // tile 0 / segment 0 of a forwarded op: the UMMAs read the A blocks from the previous op's output
              // staging (ring position kb pairs with K block fwd_perm(kb)); only the weight half is loaded here
              const bool fwd_blk = fwd_in && nt == 0 && seg == 0;
              for (int kb = 0; kb < kbs; ++kb) {
                mbar_wait(&s.empty[ps.stage], ps.phase ^ 1);
                const int kbb = fwd_blk ? fwd_perm(kb) : kb;
                const uint32_t a_bytes = fwd_blk ? 0u : (uint32_t)A_STAGE_BYTES;
                if (lane == 0) {  // the weight tile never depends on this chain: fetch it while waiting for A
                  if (CLUSTER == 1) {
                    mbar_expect_tx(&s.full[ps.stage], a_bytes + B_BYTES);
                    tma_load_2d(s.b + ps.stage * B_BYTES, &p.tmB[seg], &s.full[ps.stage], bk0 + kbb * BK, br0 + n0);
                  } else {
                    // the leader's barrier collects the bytes of BOTH CTAs (A + half B each)
                    if (leader) {
                      mbar_expect_tx(&s.full[ps.stage], CLUSTER * (a_bytes + B_BYTES));
                      }
                    tma_load_2d_pair(s.b + ps.stage * B_BYTES, &p.tmB[seg], &s.full[ps.stage], bk0 + kbb * BK,
                                     br0 + n0 + (int)crank * B_ROWS);
                  }
                }
                if (fwd_blk) {
                  __syncwarp();
                  ps.advance();
                  continue;
                }
                if (src >= 0 && static_cast<int32_t>(seen - (need0 + kb)) < 0) {
                  // wait until all eight epilogue warps have published this block, remember how far they are
                  // (later K blocks usually need no second look) and order the coming TMA loads after the acquire
                  const uint32_t need = need0 + kb;
                  uint32_t ahead = 0x7fffffffu;
                  if (lane < EPI_WARPS) ahead = wait_progress(&s.prog[lane], need) - need;
#pragma unroll
                  for (int o = 4; o >= 1; o >>= 1) ahead = min(ahead, __shfl_xor_sync(0xffffffffu, ahead, o));
                  seen = need + __shfl_sync(0xffffffffu, ahead, 0);
                  // No proxy fence: the data was written by the async proxy (TMA stores, complete and visible once
                  // their wait_group returned, before the release above) and is read by the async proxy (TMA load).
                  __syncwarp();
                }
                if (lane == 0) {
                  if (CLUSTER == 1)
                    tma_load_2d(s.a + ps.stage * A_STAGE_BYTES, &p.tmA[seg], &s.full[ps.stage], ak0 + kb * BK, m0);
                  else
                    tma_load_2d_pair(s.a + ps.stage * A_STAGE_BYTES, &p.tmA[seg], &s.full[ps.stage], ak0 + kb * BK,
                                     m0);
                }
                __syncwarp();
                ps.advance();
              }
            }
          }
        }
      }
      group_base += (uint32_t)gi * (uint32_t)cp.subs_per_stripe;
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer (pair mode: leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_f16(BM * CLUSTER, BN, 0, 0);
      constexpr uint16_t pair_mask = 0x3;
      PipeStateT<NST> ps;
      int acc = 0;
      uint32_t acc_phase = 0;
      int mma_tile_counter = 0;
      uint32_t afull_phase = 0;  // one phase bit per forwarded A block barrier
      for (int si = 0; si < stripe_iters; si += STRIPE_GROUP) {
        const int gi = min(STRIPE_GROUP, stripe_iters - si);
        for (int oi = 0; oi < cp.n_ops; ++oi) {
          const KmajorParams& p = cp.ops[oi];
          const OpScalars& sc = cp.sc[oi];
          const uint64_t dhi = sc.desc_hi ? sc.desc_hi : umma_desc_hi(16, 1024);
          const int kadv = sc.k_adv ? sc.k_adv : UMMA_K * 2;
          const int total_kb = sc.kb_per_tile;
          const int op_tiles = gi * sc.tiles_n;
          const bool fwd_in = sc.fwd_in != 0;
          const int kb_seg0 = sc.kblocks[0];
          for (int tile = 0; tile < op_tiles; ++tile) {
            long long* mdbg = (cp.dbg && (int)blockIdx.x == (cp.dbg_block & ~(CLUSTER - 1)) && lane == 0) ? cp.dbg + 8 * mma_tile_counter : nullptr;
            ++mma_tile_counter;
            mbar_wait(&s.tempty[acc], acc_phase ^ 1);
            tc_fence_after();
            if (mdbg) mdbg[5] = clock64();
            const uint32_t d_tmem = tmem_base + acc * BN;
            for (int it = 0; it < total_kb; ++it) {
              mbar_wait(&s.full[ps.stage], ps.phase);
              tc_fence_after();
              if (mdbg && it == 0) mdbg[6] = clock64();
              if (mdbg && it == total_kb - 1) mdbg[7] = clock64();
              // forwarded K block: A comes straight from the previous op's output staging (no copy, no TMA)
              const bool fwd_blk = fwd_in && tile == 0 && it < kb_seg0;
              const int fblk = fwd_perm(it);
              if (fwd_blk) {
                if (CLUSTER == 1) mbar_wait(&s.afull[fblk], (afull_phase >> fblk) & 1u);
                else mbar_wait_cluster(&s.afull[fblk], (afull_phase >> fblk) & 1u);  // the peer's warps arrive too
                afull_phase ^= 1u << fblk;
                tc_fence_after();
              }
              if (lane == 0) {
                const uint32_t a_addr = fwd_blk ? smem_u32(s.epi + (fblk & 3) * (4 * EPI_BUF_BYTES))
                                                : smem_u32(s.a + ps.stage * A_STAGE_BYTES);
                const uint32_t b_addr = smem_u32(s.b + ps.stage * B_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  if (CLUSTER == 1)
                    umma_f16(d_tmem, umma_desc(a_addr + k * kadv, dhi), umma_desc(b_addr + k * kadv, dhi), idesc,
                             (it | k) != 0);
                  else
                    umma_f16_pair(d_tmem, umma_desc(a_addr + k * kadv, dhi), umma_desc(b_addr + k * kadv, dhi), idesc,
                                  (it | k) != 0);
                }
                if (CLUSTER == 1) umma_commit(&s.empty[ps.stage]);
                else umma_commit_pair(&s.empty[ps.stage], pair_mask);   // frees the stage in both CTAs
                // staging block fblk held sub-tile fblk of the previous op's FIRST tile; once these UMMAs are done its
                // second tile may overwrite it (only if there is a second tile: one wait per commit)
                if (fwd_blk && it < 4 && kb_seg0 > 4) {
                  if (CLUSTER == 1) umma_commit(&s.fempty[fblk]);
                  else umma_commit_pair(&s.fempty[fblk], pair_mask);
                }
              }
              __syncwarp();
              ps.advance();
            }
            if (lane == 0) {
              if (CLUSTER == 1) umma_commit(&s.tfull[acc]);
              else umma_commit_pair(&s.tfull[acc], pair_mask);          // both CTAs' epilogues may read their half
            }
            __syncwarp();
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (every CTA, own 128 rows)
    uint32_t ld_phase = 0;
    uint32_t seq = 0;
    uint32_t deferred = 0;  // progress value of a tile whose drain + publish was deferred into the next tile
    int tile_counter = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t fwd_phase = 0;  // phase bits of the two forwarding class barriers this warp waits on
    bool mul_ready = false;  // the dgrad multiplier of the coming tile's first sub-tile is already on its way
    const int et_idx = ((warp - 2) >> 2) * 128 + 4 * lane;  // first of this lane's four bias columns inside a tile
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 bias_pref = zero4;
    bool first_tile = true;
    static_assert(STRIPE_GROUP == 1, "on-chip forwarding assumes one stripe at a time");
    for (int si = 0; si < stripe_iters; si += STRIPE_GROUP) {
      const int gi = min(STRIPE_GROUP, stripe_iters - si);
      const bool final_group = si + gi >= stripe_iters;
      for (int oi = 0; oi < cp.n_ops; ++oi) {
        const KmajorParams& p = cp.ops[oi];
        if (lane == 0) {
          tma_prefetch_desc(&p.tmOut0);
          if (cp.sc[oi].epi == EPI_SNAKE) tma_prefetch_desc(&p.tmOut1);
          if (cp.sc[oi].epi == EPI_DGRAD_MUL) tma_prefetch_desc(&p.tmMul);
        }
        const OpScalars& sc = cp.sc[oi];
        const int epi = sc.epi;
        const int op_tiles_n = sc.tiles_n;
        EpiArgs ea;
        const float* op_bias = sc.bias;
        // the bias row of the op that follows in processing order (its first tile is prefetched during this op's last)
        const float* next_op_bias = nullptr;
        if (oi + 1 < cp.n_ops) next_op_bias = cp.sc[oi + 1].bias;
        else if (si + gi < stripe_iters) next_op_bias = cp.sc[0].bias;
        if (first_tile) {
          bias_pref = op_bias != nullptr ? __ldg(reinterpret_cast<const float4*>(op_bias + et_idx)) : zero4;
          first_tile = false;
        }
        ea.colsum = sc.colsum;
        ea.out_f32 = sc.out_f32;
        ea.ldf = sc.ldf;
        ea.fwd_out = sc.fwd_out;
        ea.leader = CLUSTER == 1 || leader;
        for (int sl = 0; sl < gi; ++sl) {
          const int m0 = ((si + sl) * (int)gridDim.x + (int)blockIdx.x) * BM;
          for (int nt = 0; nt < op_tiles_n; ++nt) {
            ea.nt = nt;
            ea.next_mul = nullptr;
            ea.next_mul_col = 0;
            if (epi == EPI_DGRAD_MUL && gi == 1) {  // the tile that follows in this stripe, if it multiplies too
              if (nt + 1 < op_tiles_n) {
                ea.next_mul = &p.tmMul;
                ea.next_mul_col = (nt + 1) * BN;
              } else if (oi + 1 < cp.n_ops && cp.sc[oi + 1].epi == EPI_DGRAD_MUL) {
                ea.next_mul = &cp.ops[oi + 1].tmMul;
              }
            }
            ea.bias4 = bias_pref;
            {  // bias of the NEXT tile: in flight while this tile is drained
              const float* nb = nt + 1 < op_tiles_n ? op_bias : next_op_bias;
              const int ncol = nt + 1 < op_tiles_n ? (nt + 1) * BN : 0;
              bias_pref = nb != nullptr ? __ldg(reinterpret_cast<const float4*>(nb + ncol + et_idx)) : zero4;
            }
            const uint32_t tacc = tmem_base + acc * BN;
            const int n0 = nt * BN;
            // Publish right away only when somebody is about to wait for it: a lone stripe's last tile of an op
            // (its next op is next in line) and the very last tile of the kernel.  Everything else is published
            // from inside the following tile (the consumer of an interleaved stripe comes a whole stripe later).
            long long* dbgp = (cp.dbg && (int)blockIdx.x == cp.dbg_block && warp == cp.dbg_warp) ? cp.dbg + 8 * tile_counter : nullptr;
            ++tile_counter;
            const bool last_of_stripe_op = nt == op_tiles_n - 1;
            const bool last = (gi == 1 && last_of_stripe_op) ||
                              (final_group && oi == cp.n_ops - 1 && sl == gi - 1 && last_of_stripe_op);
            switch (epi) {
              case EPI_LINEAR:
                epilogue_tile<EPI_LINEAR, false>(p, ea, s, tacc, m0, n0, cp.M, warp, lane, ld_phase, &s.tfull[acc], acc_phase, seq,
                                          last, deferred, dbgp, fwd_phase, mul_ready);
                break;
              case EPI_SNAKE:
                epilogue_tile<EPI_SNAKE, RELU>(p, ea, s, tacc, m0, n0, cp.M, warp, lane, ld_phase, &s.tfull[acc], acc_phase, seq,
                                         last, deferred, dbgp, fwd_phase, mul_ready);
                break;
              case EPI_SNAKE_HEAD:
                epilogue_head_tile<RELU>(p, ea, cp.head, s, tacc, m0, cp.M, warp, lane, &s.tfull[acc], acc_phase, seq,
                                         deferred);
                break;
              case EPI_DGRAD_MUL:
                epilogue_tile<EPI_DGRAD_MUL, false>(p, ea, s, tacc, m0, n0, cp.M, warp, lane, ld_phase, &s.tfull[acc], acc_phase,
                                             seq, last, deferred, dbgp, fwd_phase, mul_ready);
                break;
              default:
                epilogue_tile<EPI_DGRAD, false>(p, ea, s, tacc, m0, n0, cp.M, warp, lane, ld_phase, &s.tfull[acc], acc_phase, seq,
                                         last, deferred, dbgp, fwd_phase, mul_ready);
                break;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              // the accumulator stage is released on the LEADER's barrier (it gates the leader's next UMMAs)
              if (CLUSTER == 1 || leader) mbar_arrive(&s.tempty[acc]);
              else mbar_arrive_remote_relaxed(&s.tempty[acc], 0);
              if (dbgp != nullptr) dbgp[4] = clock64();
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
          }
        }
      }
    }
  }
  gemm_teardown<CLUSTER>(tmem_base, warp);
}

// ---------------------------------------------------------------------------------
// CLUSTER == 1: one CTA per 128(out) x 256(in) tile.
// CLUSTER == 2: a CTA pair per 256(out) x 256(in) tile with cta_group::2 UMMAs: CTA r owns out rows
//   [m0 + 128 r, +128) (its own delta columns as A) and loads only half of the activation tile (B columns
//   [n0 + 128 r, +128)), so a pair moves 64 KB per K block for twice the MMA work (the single-CTA kernel is L2
//   bandwidth bound: 1.5 GB of operand reads per step at 16 k rows).
#ifndef NPP_WG_PAIR_STAGES
#define NPP_WG_PAIR_STAGES 5
#endif
constexpr int WG_PAIR_STAGES = NPP_WG_PAIR_STAGES;
constexpr int WGRAD_PAIR_SMEM_BYTES = WG_PAIR_STAGES * (A_STAGE_BYTES + B_STAGE_BYTES / 2) + 512 + 1024;
// what the pair kernel is LAUNCHED with: the whole 227 KB, so that a CTA of a pair has its SM to itself (see launch_wgrad)
constexpr int WGRAD_PAIR_LAUNCH_SMEM_BYTES = 232448;
static_assert(WGRAD_PAIR_SMEM_BYTES <= WGRAD_PAIR_LAUNCH_SMEM_BYTES, "wgrad ring exceeds 227 KB of shared memory");

// 64-row blocks [kb0, kb1) a unit contracts over; false: nothing to do (every warp role skips the unit alike)
__device__ __forceinline__ bool wg_unit_range(const WgUnit& un, const WgradParams& p, int kb_total, int& kb0, int& kb1) {
  if (un.kb1 != 0) {
    kb0 = un.kb0;
    kb1 = min(un.kb1, kb_total);
    return kb0 < kb1;
  }
  if (un.split >= p.n_splits) return false;
  kb0 = un.split * p.kb_per_split;
  kb1 = min(kb0 + p.kb_per_split, kb_total);
  return kb0 < kb1;
}

template <int CLUSTER>
__global__ void __launch_bounds__(WGRAD_THREADS, 1) npp_gemm_wgrad(const __grid_constant__ WgradParams p) {
  constexpr int NST = CLUSTER == 1 ? WG_STAGES : WG_PAIR_STAGES;
  constexpr int B_BYTES = B_STAGE_BYTES / CLUSTER;
  constexpr int B_COLS = BN / CLUSTER;
  extern __shared__ uint8_t smem_raw[];
  const GemmSmem s = carve_smem_t<NST, 0, A_STAGE_BYTES, B_BYTES>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  static_assert(NST <= 8, "the per-stage `alocal` barriers reuse the eight afull slots");
  // Bias gradients (see the epilogue warps below) need to know when THIS CTA's delta tile of a stage has landed.  In
  // pair mode every TMA load used to complete on the leader's full barrier, which the peer CTA cannot wait on; now the
  // peer's delta tile completes on a peer-local barrier (`alocal`, the afull slots), and the peer's otherwise idle warp
  // 1 forwards that completion to the leader's full barrier (second pending arrival there).  A stage is free again when
  // its UMMAs have completed (one commit) and the four epilogue warps of the CTA have arrived.
  uint64_t* const alocal = s.afull;
#ifdef NPP_PDL_WAIT_FIRST
  grid_dependency_wait();
#endif
#ifdef NPP_HANG_DEBUG
  if (threadIdx.x == 0) *npp_state_slot() = p.dbg_state;
#endif
  const uint32_t tmem_base = gemm_prologue_t<NST, CLUSTER, 4, CLUSTER, 1 + 4, 1>(s, warp);
  grid_dependency_wait();   // programmatic dependent launch: the prologue above overlaps the previous kernel's tail
  const int kb_total = (p.rows + BK - 1) / BK;
  const uint32_t crank = CLUSTER > 1 ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  const int unit0 = (int)blockIdx.x / CLUSTER;      // both CTAs of a pair walk the same unit list
  const int ustride = (int)gridDim.x / CLUSTER;

  if (warp == 0) {
    PipeStateT<NST> ps;
    for (int u = unit0; u < p.n_units; u += ustride) {
      const WgUnit un = p.units[u];
      int kb0, kb1;
      if (!wg_unit_range(un, p, kb_total, kb0, kb1)) continue;
      const CUtensorMap* ma = &p.maps[un.a_map];
      const CUtensorMap* mb = &p.maps[un.b_map];
      const int am0 = un.a_m0 + (int)crank * BM;
      const int bn0 = un.b_n0 + (int)crank * B_COLS;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&s.empty[ps.stage], ps.phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = s.a + ps.stage * A_STAGE_BYTES;
          uint8_t* sb = s.b + ps.stage * B_BYTES;
          if (CLUSTER == 1) {
            mbar_expect_tx(&s.full[ps.stage], A_STAGE_BYTES + B_BYTES);
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, ma, &s.full[ps.stage], am0 + j * 64, kb * BK);
#pragma unroll
            for (int j = 0; j < B_COLS / 64; ++j) tma_load_2d(sb + j * 8192, mb, &s.full[ps.stage], bn0 + j * 64, kb * BK);
          } else if (leader) {
            mbar_expect_tx(&s.full[ps.stage], A_STAGE_BYTES + CLUSTER * B_BYTES);   // own delta tile + both activation halves
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, ma, &s.full[ps.stage], am0 + j * 64, kb * BK);
#pragma unroll
            for (int j = 0; j < B_COLS / 64; ++j)
              tma_load_2d_pair(sb + j * 8192, mb, &s.full[ps.stage], bn0 + j * 64, kb * BK);
          } else {
            mbar_expect_tx(&alocal[ps.stage], A_STAGE_BYTES);                        // the peer's delta tile: local barrier
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, ma, &alocal[ps.stage], am0 + j * 64, kb * BK);
#pragma unroll
            for (int j = 0; j < B_COLS / 64; ++j)
              tma_load_2d_pair(sb + j * 8192, mb, &s.full[ps.stage], bn0 + j * 64, kb * BK);
          }
        }
        __syncwarp();
        ps.advance();
      }
    }
    if (lane == 0) grid_launch_dependents();   // all loads of this CTA are issued: the next kernel may start its prologue
  } else if (warp == 1) {
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_f16(BM * CLUSTER, BN, 1, 1);
      constexpr uint16_t pair_mask = 0x3;
      const uint64_t dhi = p.desc_hi ? p.desc_hi : umma_desc_hi(8192, 1024);
      const int kadv = p.k_adv ? p.k_adv : 2048;
      PipeStateT<NST> ps;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = unit0; u < p.n_units; u += ustride) {
        const WgUnit un = p.units[u];
        int kb0, kb1;
        if (!wg_unit_range(un, p, kb_total, kb0, kb1)) continue;
        mbar_wait(&s.tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&s.full[ps.stage], ps.phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_u32(s.a + ps.stage * A_STAGE_BYTES);
            const uint32_t b_addr = smem_u32(s.b + ps.stage * B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
              if (CLUSTER == 1)
                umma_f16(d_tmem, umma_desc(a_addr + k * kadv, dhi), umma_desc(b_addr + k * kadv, dhi), idesc, accum);
              else
                umma_f16_pair(d_tmem, umma_desc(a_addr + k * kadv, dhi), umma_desc(b_addr + k * kadv, dhi), idesc,
                              accum);
            }
            if (CLUSTER == 1) umma_commit(&s.empty[ps.stage]);
            else umma_commit_pair(&s.empty[ps.stage], pair_mask);
          }
          __syncwarp();
          ps.advance();
        }
        if (lane == 0) {
          if (CLUSTER == 1) umma_commit(&s.tfull[acc]);
          else umma_commit_pair(&s.tfull[acc], pair_mask);
        }
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    } else {
      // peer CTA: tell the leader's full barrier that this CTA's delta tile of the stage has landed.  Relaxed arrive: the
      // tile was written by the async proxy and is read by the tensor core, no generic-proxy data is published.
      PipeStateT<NST> ps;
      for (int u = unit0; u < p.n_units; u += ustride) {
        const WgUnit un = p.units[u];
        int kb0, kb1;
        if (!wg_unit_range(un, p, kb_total, kb0, kb1)) continue;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&alocal[ps.stage], ps.phase);
          if (lane == 0) mbar_arrive_remote_relaxed(&s.full[ps.stage], 0);
          __syncwarp();
          ps.advance();
        }
      }
    }
  } else {
    const int lane_base = (warp & 3) * 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    PipeStateT<NST> ps;
    // this CTA's delta tile of a stage has landed: the full barrier in the leader (and without pairs), alocal in the peer
    uint64_t* const mine = (CLUSTER == 1 || leader) ? s.full : alocal;
    // Bias gradients: db[o] = sum over rows of delta[row, o] (fp16-rounded, scaled: exactly the values the tensor core
    // consumes).  The delta tile of every K block passes through this CTA's shared memory anyway (A operand, 64 rows x
    // 128 out-features, MN-major: row r = 128 bytes per 64-feature atom, 16-byte chunks XOR-swizzled with r & 7), so the
    // epilogue warps, idle during the mainloop, add it up: thread t owns chunk column t & 15 (8 features) of the rows
    // r = t >> 4 (mod 8).  This used to be a 32 x 32 shuffle transpose-reduce per epilogue block of the dgrad chain
    // (a quarter of that epilogue's instructions, on the kernel's critical resource).
    const int et = (warp - 2) * 32 + lane;          // 0..127
    const int cc = et & 15, rg = et >> 4;
    const uint32_t rd_off = (uint32_t)(cc >> 3) * 8192u + (uint32_t)rg * 128u + (uint32_t)(((cc & 7) ^ rg) << 4);
    for (int u = unit0; u < p.n_units; u += ustride) {
      const WgUnit un = p.units[u];
      int kb0, kb1;
      if (!wg_unit_range(un, p, kb_total, kb0, kb1)) continue;
      float* out = p.partial + (size_t)un.split * p.slab_stride + un.out_off +
                   (size_t)((int)crank * BM + lane_base + lane) * un.ld;
      {
        const bool want = un.bias_off >= 0 && p.bias_acc != nullptr;
        float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int kb = kb0; kb < kb1; ++kb) {
          // Units without a bias sum arrive as soon as the stage's previous round is over (never on the critical path);
          // the others once they have read the tile, which lands long before the stage's UMMAs complete.
          if (want) mbar_wait(&mine[ps.stage], ps.phase);
          else mbar_wait(&s.empty[ps.stage], ps.phase ^ 1);
          if (want) {
            const uint8_t* sa = s.a + ps.stage * A_STAGE_BYTES + rd_off;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 q4 = *reinterpret_cast<const uint4*>(sa + i * 1024);
              const float2 f0 = unpack_h2(q4.x), f1 = unpack_h2(q4.y), f2 = unpack_h2(q4.z), f3 = unpack_h2(q4.w);
              bs[0] += f0.x; bs[1] += f0.y; bs[2] += f1.x; bs[3] += f1.y;
              bs[4] += f2.x; bs[5] += f2.y; bs[6] += f3.x; bs[7] += f3.y;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&s.empty[ps.stage]);
          ps.advance();
        }
        if (want) {
#pragma unroll
          for (int i = 0; i < 8; ++i) bs[i] += __shfl_xor_sync(0xffffffffu, bs[i], 16);   // rows rg and rg + 1 of this warp
          if (lane < 16) {
            float* dst = p.bias_acc + un.bias_off + (int)crank * BM + cc * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) atomicAdd(dst + i, bs[i]);
          }
        }
      }
      mbar_wait(&s.tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(lane_base) << 16) + acc * BN + chunk * 32, raw);
        tmem_ld_wait();
        if (chunk * 32 < un.ncols_left) {
          float4* o = reinterpret_cast<float4*>(out + chunk * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o[i] = make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]),
                               __uint_as_float(raw[4 * i + 2]), __uint_as_float(raw[4 * i + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CLUSTER == 1 || leader) mbar_arrive(&s.tempty[acc]);
        else mbar_arrive_remote(&s.tempty[acc], 0);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  gemm_teardown<CLUSTER>(tmem_base, warp);
}

}  // namespace npp
