// Fused attention forward for sm_100a:  QK^T -> softmax_1 / clipped softmax(_1) [-> gate] -> PV
// (bert_attention.py:222-337, opt_attention.py:204-322, quantized_{bert,opt}.py taps), written
// directly against tcgen05 / TMEM / TMA with inline PTX (sm100_ptx.cuh).
//
// Data flow of one CTA (128 threads, one 128-row Q tile of one (batch, head)):
//   TMA (SWIZZLE_128B, 4-D tensor maps built from the caller's strides, OOB rows zero-filled)
//     Q tile [128 x 64]  -> smem once;   K/V tiles [128 x 64] -> 2-stage smem rings
//   tcgen05.mma  S[128x128] (fp32, TMEM cols 0..127)   = Q (K-major smem) . K^T (K-major smem)
//   tcgen05.ld   each thread owns one S row (lane == row): scale, qdq, mask, online softmax with the
//                softmax_1 "+1" carried as exp2(-m) in the running-max frame (vutils/softmax_1.py:18-20)
//   tcgen05.st   P (fp16/bf16, packed 2 per column) back into TMEM cols 0..63, aliasing S
//   tcgen05.mma  O[128x64] (fp32, TMEM cols 128..191) += P (A operand from TMEM) . V (MN-major smem)
//   epilogue     O / (l [+ exp2(-m)]) -> [qdq] -> * sigmoid(gate)*scale -> [qdq] -> 16-byte stores into
//                the caller's (B,T,H*d) or (B,H,T,d) layout.
// Clipped variants and the probs fake-quant need the *normalised* probability before PV, so they run
// a stats pass (QK^T only, running (m,l)) and then a PV pass that recomputes QK^T; with a single K/V
// tile (Tk <= 128, the BERT eval shape) both passes collapse into one.
// Two CTAs are co-resident per SM (2 x 256 TMEM columns, 2 x 81 KB smem) so one CTA's softmax
// overlaps the other's MMAs.  Roofline: tensor pipe, 4*B*H*Tq*Tk*d flop per call (causal: unmasked half).
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>

#include "attn_common.cuh"
#include "attn_fwd_mid.cuh"
#include "attn_fwd_one2.cuh"
#include "attn_fwd_ws.cuh"
#include "attn_fwd_wsp.cuh"

namespace oeh {
namespace {

using namespace attn;

constexpr int BM = 128;                 // query rows per CTA  (== threads per CTA, lane == row)
constexpr int kSmemBytes = 5 * kTileBytes + 1024;
constexpr uint32_t kTmemCols = 256;
constexpr uint32_t kColS = 0, kColP = 0, kColO = 128;

// kGeneric = false: no fake-quant taps, no mask tensor, no clamp, no debug dumps -- the tight hot loop.
// kGeneric = true : everything; each optional feature is its own loop behind one uniform branch so the
//                   instruction stream actually executed stays small (a per-element branch version of this
//                   kernel was 310 KB of SASS and ran 3-4x slower from instruction-cache misses).
template <bool kBF16, bool kGeneric>
__global__ void __launch_bounds__(BM, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                const __grid_constant__ CUtensorMap tmv, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[6];      // q, k0, k1, v0, v1, mma
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const int row = q0 + tid;                       // this thread's query row

  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-B alignment
  const uint32_t sQ = smem0;
  const uint32_t sK[2] = {smem0 + 1 * kTileBytes, smem0 + 2 * kTileBytes};
  const uint32_t sV[2] = {smem0 + 3 * kTileBytes, smem0 + 4 * kTileBytes};
  const uint32_t bar_q = ptx::smem_u32(&bars[0]);
  const uint32_t bar_k[2] = {ptx::smem_u32(&bars[1]), ptx::smem_u32(&bars[2])};
  const uint32_t bar_v[2] = {ptx::smem_u32(&bars[3]), ptx::smem_u32(&bars[4])};
  const uint32_t bar_mma = ptx::smem_u32(&bars[5]);

  if (tid == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
#pragma unroll
    for (int i = 0; i < 6; ++i) ptx::mbar_init(ptx::smem_u32(&bars[i]), 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(&tmem_base_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  const uint32_t t_lane = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32-lane slice

  // ---- iteration plan -----------------------------------------------------------------------------
  int n_kv = (p.Tk + BN - 1) / BN;
  if (p.causal) {
    const int last = q0 + BM - 1 + (p.Tk - p.Tq);               // last key any row of this tile may see
    const int lim = last / BN + 1;
    n_kv = lim < n_kv ? lim : n_kv;
  }
  const bool softmax1 = (p.kind == OEH_SOFTMAX_1 || p.kind == OEH_SOFTMAX_CLIPPED1);
  const bool clipped = p.kind >= OEH_SOFTMAX_CLIPPED;
  const bool normalized = clipped || p.qdq_probs != nullptr;   // P must be normalised before PV
  const bool one_tile = normalized && n_kv == 1;
  const int nA = (normalized && !one_tile) ? n_kv : 0;         // stats-pass iterations
  const int n_iter = nA + n_kv;

  constexpr uint32_t kIdescQK = ptx::make_idesc_f16(kBF16, BM, BN, false, false);
  constexpr uint32_t kIdescPV = ptx::make_idesc_f16(kBF16, BM, HD, false, true);

  auto issue_loads = [&](int it) {
    const int tile = it < nA ? it : it - nA;
    const int ks = it & 1;
    ptx::mbar_expect_tx(bar_k[ks], kTileBytes);
    tma_tile(sK[ks], &tmk, bar_k[ks], p.ck, tile * BN, h, b);
    if (it >= nA) {
      const int vs = (it - nA) & 1;
      ptx::mbar_expect_tx(bar_v[vs], kTileBytes);
      tma_tile(sV[vs], &tmv, bar_v[vs], p.cv, tile * BN, h, b);
    }
  };
  auto issue_qk = [&](int it) {
    ptx::mbar_wait(bar_k[it & 1], (uint32_t)(it >> 1) & 1u);
    ptx::tc_fence_after();
#pragma unroll
    for (int k = 0; k < HD / 16; ++k) {
      const uint64_t da = ptx::make_smem_desc_sw128(sQ + k * 32, 16, 1024);
      const uint64_t db = ptx::make_smem_desc_sw128(sK[it & 1] + k * 32, 16, 1024);
      ptx::mma_ss(tmem + kColS, da, db, kIdescQK, k > 0);
    }
  };
  auto issue_pv = [&](int it) {
    const int vi = it - nA;
    ptx::mbar_wait(bar_v[vi & 1], (uint32_t)(vi >> 1) & 1u);
    ptx::tc_fence_after();
#pragma unroll
    for (int k = 0; k < BN / 16; ++k) {
      const uint64_t db = ptx::make_smem_desc_sw128(sV[vi & 1] + k * 2048, 1024, 1024);
      ptx::mma_ts(tmem + kColO, tmem + kColP + k * 8, db, kIdescPV, (vi > 0 || k > 0) ? 1u : 0u);
    }
  };

  // one ELECTED lane of warp 0 issues TMA / tcgen05 (see ptx::elect_one)
  if (warp == 0 && ptx::elect_one()) {
    ptx::mbar_expect_tx(bar_q, kTileBytes);
    tma_tile(sQ, &tmq, bar_q, p.cq, q0, h, b);
    issue_loads(0);
    if (n_iter > 1) issue_loads(1);
    ptx::mbar_wait(bar_q, 0);
    issue_qk(0);
    ptx::mma_commit(bar_mma);
  }
  __syncwarp();

  // ---- per-row state ------------------------------------------------------------------------------
  const bool has_qs = kGeneric && p.qdq_scores != nullptr;
  const bool has_qp = kGeneric && p.qdq_probs != nullptr;
  const bool has_mask = kGeneric && p.mask != nullptr;
  Qdq qd_s{1.f, 1.f, 0.f, p.int_max}, qd_p{1.f, 1.f, 0.f, p.int_max};
  if (has_qs) qd_s = load_qdq(p.qdq_scores, p.int_max);
  if (has_qp) qd_p = load_qdq(p.qdq_probs, p.int_max);
  const int64_t mask_row = p.mask ? ((int64_t)b * p.ms0 + (int64_t)h * p.ms1 + (int64_t)(row < p.Tq ? row : 0) * p.ms2) : 0;
  const int causal_lim = p.causal ? row + (p.Tk - p.Tq) : 0x7fffffff;   // keys j <= causal_lim are visible

  float m_run = -INFINITY, l_run = 0.f, inv_den = 1.f;
  uint32_t mma_phase = 0;

  for (int it = 0; it < n_iter; ++it) {
    ptx::mbar_wait(bar_mma, mma_phase);
    mma_phase ^= 1u;
    ptx::tc_fence_after();
    // S(it) is complete, and so is every MMA issued before it: the stage used by iteration it-1 is free
    if (warp == 0 && it >= 1 && it + 1 < n_iter && ptx::elect_one()) issue_loads(it + 1);
    __syncwarp();

    const bool pv_iter = it >= nA;
    const int tile = pv_iter ? it - nA : it;
    const int k0 = tile * BN;

    float s[BN];
    {
      uint32_t* sr = reinterpret_cast<uint32_t*>(s);
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) ptx::tmem_ld32(t_lane + kColS + c * 32, sr + c * 32);
      ptx::tmem_wait_ld();
    }
    if (kGeneric && p.dbg_s && it == 0 && row < p.Tq) {
      float* d = p.dbg_s + (((int64_t)b * p.H + h) * p.Tq + row) * BN;
#pragma unroll
      for (int j = 0; j < BN; ++j) d[j] = s[j];
    }

    const int valid = min(p.Tk - k0, BN);            // columns < valid exist
    const int vis = min(valid - 1, causal_lim - k0); // columns <= vis are visible to this row
    const float mul = transform_scores<BN, kGeneric>(s, p, qd_s, has_qs, has_mask, mask_row, k0, valid, vis);
    const float tile_max = row_max(s) * mul;         // mul > 0

    const bool online = !normalized;
    if (!pv_iter || one_tile || online) {
      // running (m, l) update
      const float m_new = fmaxf(m_run, tile_max);
      const float m_safe = m_new == -INFINITY ? 0.f : m_new;
      const float alpha = ptx::ex2(m_run - m_safe);   // m_run = -inf -> 0
      l_run = fmaf(l_run, alpha, exp2_and_sum(s, mul, m_safe));
      m_run = m_new;
      if (!pv_iter) {
        ptx::tc_fence_before();
        __syncthreads();                              // every thread has read S(it)
        if (warp == 0 && ptx::elect_one()) {
          ptx::tc_fence_after();
          issue_qk(it + 1);                           // a stats iteration is never the last one
          ptx::mma_commit(bar_mma);
        }
        __syncwarp();
        continue;
      }
      if (one_tile) {
        const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
        inv_den = __frcp_rn(den);
      } else if (it > nA && !__all_sync(0xffffffffu, alpha == 1.f)) {
        // correction: O *= alpha (TMEM round trip; skipped when no row of this warp moved its max)
        uint32_t o[HD];
        ptx::tmem_ld32(t_lane + kColO, o);
        ptx::tmem_ld32(t_lane + kColO + 32, o + 32);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < HD; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
        ptx::tmem_st32(t_lane + kColO, o);
        ptx::tmem_st32(t_lane + kColO + 32, o + 32);
      }
    } else {
      // PV pass of the two-pass (normalised) mode: m, l are final
      const float m_safe = m_run == -INFINITY ? 0.f : m_run;
      if (it == nA) {
        const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
        inv_den = __frcp_rn(den);
      }
#pragma unroll
      for (int j = 0; j < BN; ++j) s[j] = ptx::ex2(fmaf(s[j], mul, -m_safe));
    }

    if (normalized) normalize_probs<BN, kGeneric, kBF16>(s, p, inv_den, clipped, has_qp, qd_p);

    // P -> TMEM (packed 16-bit pairs), aliasing the S columns this thread has already consumed
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = pack2<kBF16>(s[c * 32 + 2 * j], s[c * 32 + 2 * j + 1]);
      ptx::tmem_st16(t_lane + kColP + c * 16, pk);
    }
    ptx::tmem_wait_st();
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0 && ptx::elect_one()) {
      ptx::tc_fence_after();
      issue_pv(it);
      if (it + 1 < n_iter) issue_qk(it + 1);
      ptx::mma_commit(bar_mma);
    }
    __syncwarp();
  }

  // ---- epilogue -------------------------------------------------------------------------------------
  ptx::mbar_wait(bar_mma, mma_phase);
  ptx::tc_fence_after();
  float o[HD];
  {
    uint32_t* orr = reinterpret_cast<uint32_t*>(o);
    ptx::tmem_ld32(t_lane + kColO, orr);
    ptx::tmem_ld32(t_lane + kColO + 32, orr + 32);
    ptx::tmem_wait_ld();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<kTmemCols>(tmem);

  if (row >= p.Tq) return;
  float mul = 1.f;
  if (!normalized) {
    const float m_safe = m_run == -INFINITY ? 0.f : m_run;
    const float den = softmax1 ? l_run + ptx::ex2(-m_safe) : l_run;
    mul = __frcp_rn(den);
  }
  store_row<kBF16, kGeneric>(o, p, mul, b, h, row);
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 4-D tensor map over a (B,H,T,64) strided view whose head dim is contiguous: dim0 = d, dims 1..3 =
// (token, head, batch) ordered by increasing stride; box = [64, box_rows tokens, 1, 1]; 128-byte swizzle.
int make_tmap(CUtensorMap* tm, TmapCoord* coord, const void* base, int dtype, int B, int H, int T,
              const int64_t* strides /* (b,h,t) elements */, int box_rows, const char* what) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return set_error(OEH_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (reinterpret_cast<uintptr_t>(base) & 15)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_attn_fwd: %s base pointer must be 16-byte aligned", what);
  struct Dim { int64_t size, stride; int logical; } dims[3] = {{T, strides[2], 0}, {H, strides[1], 1}, {B, strides[0], 2}};
  for (int i = 0; i < 3; ++i) {
    if (dims[i].size > 1 && (dims[i].stride <= 0 || (dims[i].stride * 2) % 16 != 0))
      return set_error(OEH_ERR_UNSUPPORTED, "oeh_attn_fwd: %s stride %lld (elements) must be a positive multiple of 8",
                       what, (long long)dims[i].stride);
  }
  // size-1 dims: their stride is irrelevant; give them a harmless large stride so they sort last
  int64_t span = HD;
  for (int i = 0; i < 3; ++i)
    if (dims[i].size > 1) span = span > dims[i].stride * dims[i].size ? span : dims[i].stride * dims[i].size;
  for (int i = 0; i < 3; ++i)
    if (dims[i].size <= 1) { dims[i].size = 1; dims[i].stride = (span + 7) / 8 * 8; }
  for (int i = 0; i < 3; ++i)       // sort by stride, ascending
    for (int j = i + 1; j < 3; ++j)
      if (dims[j].stride < dims[i].stride) { Dim t = dims[i]; dims[i] = dims[j]; dims[j] = t; }
  cuuint64_t gdim[4] = {HD, (cuuint64_t)dims[0].size, (cuuint64_t)dims[1].size, (cuuint64_t)dims[2].size};
  cuuint64_t gstr[3] = {(cuuint64_t)dims[0].stride * 2, (cuuint64_t)dims[1].stride * 2, (cuuint64_t)dims[2].stride * 2};
  cuuint32_t box[4] = {HD, 1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i) {
    coord->perm[i] = (int8_t)dims[i].logical;
    if (dims[i].logical == 0) box[i + 1] = (cuuint32_t)box_rows;
  }
  CUresult r = enc(tm, dtype == OEH_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                   const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_attn_fwd: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return OEH_OK;
}

// ---- dispatch overrides (tests / tuning) ---------------------------------------------------------------------------
// The kernel-choice rules below are measured defaults.  They can be overridden per process through oeh_set_override()
// (thread-safe; what the tests use) or, for command-line experiments, through environment variables of the same names --
// read ONCE, at the first call, never on the launch path.  Value < 0 = no override.
enum TuneKey { kTuneAttnKernel = 0, kTuneWsPersistent, kTuneWsTiles, kTuneCausalChunk, kTuneOneRegs, kTuneMidKernel, kTuneOneSplit, kTuneKeys };
const char* const kTuneNames[kTuneKeys] = {"OEH_ATTN_KERNEL", "OEH_WS_PERSISTENT", "OEH_WS_TILES", "OEH_CAUSAL_CHUNK",
                                           "OEH_ONE_REGS", "OEH_MID_KERNEL", "OEH_ONE_SPLIT"};
std::atomic<long long> g_tune[kTuneKeys];
std::once_flag g_tune_once;

long long parse_tune(int key, const char* v) {
  if (key == kTuneAttnKernel) return !strcmp(v, "tile") ? 0 : (!strcmp(v, "ws") ? 1 : -1);   // 0 = tile, 1 = ws
  return atoll(v);
}
void tune_init() {
  std::call_once(g_tune_once, [] {
    for (int k = 0; k < kTuneKeys; ++k) {
      const char* v = getenv(kTuneNames[k]);
      g_tune[k].store(v ? parse_tune(k, v) : -1, std::memory_order_relaxed);
    }
  });
}
long long tune(int key) {
  tune_init();
  return g_tune[key].load(std::memory_order_relaxed);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel): function attributes belong to the device's
// context, and two host threads may race to the first launch.
int ensure_smem(const void* fn, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  OEH_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({dev, fn})) return OEH_OK;
  OEH_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert({dev, fn});
  return OEH_OK;
}

// ---- fp32 activations ------------------------------------------------------------------------------------------------
// The reference evaluates in fp32 (accelerate_configs/1gpu_no_mp.yaml:14).  tcgen05 has no fp32 MMA; the fused path for fp32
// tensors rounds Q / K / V to fp16 OPERANDS (saturating at +-65504), runs the same kernels -- fp32 accumulation, fp32
// softmax, taps on fp32 values -- and widens the context back to fp32: the accuracy of the fp16 path (2e-3 max-abs against
// the fp32 oracle), not of an fp32 GEMM, which is why the Python modules use it only on request (``fused_fp32``).
__global__ void __launch_bounds__(256) cvt_rows_f32_to_f16(const float* __restrict__ src, __half* __restrict__ dst, int H, int T,
                                                           int64_t s0, int64_t s1, int64_t s2, int64_t rows) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;      // one 16-byte output vector (8 elements) per thread
  const int64_t row = i >> 3;
  if (row >= rows) return;
  const int c = (int)(i & 7) * 8;
  const int t = (int)(row % T), h = (int)((row / T) % H);
  const int64_t b = row / ((int64_t)T * H);
  const float4* sp = reinterpret_cast<const float4*>(src + b * s0 + h * s1 + t * s2 + c);
  const float4 x = __ldg(sp), y = __ldg(sp + 1);
  auto sat = [](float v) { return fminf(fmaxf(v, -65504.f), 65504.f); };
  __half2 o[4] = {__floats2half2_rn(sat(x.x), sat(x.y)), __floats2half2_rn(sat(x.z), sat(x.w)),
                  __floats2half2_rn(sat(y.x), sat(y.y)), __floats2half2_rn(sat(y.z), sat(y.w))};
  *reinterpret_cast<uint4*>(dst + row * HD + c) = *reinterpret_cast<uint4*>(o);
}
__global__ void __launch_bounds__(256) cvt_rows_f16_to_f32(const __half* __restrict__ src, float* __restrict__ dst, int H, int T,
                                                           int64_t s0, int64_t s1, int64_t s2, int64_t rows) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t row = i >> 3;
  if (row >= rows) return;
  const int c = (int)(i & 7) * 8;
  const int t = (int)(row % T), h = (int)((row / T) % H);
  const int64_t b = row / ((int64_t)T * H);
  const uint4 v = *reinterpret_cast<const uint4*>(src + row * HD + c);
  const __half2* hp = reinterpret_cast<const __half2*>(&v);
  const float2 a = __half22float2(hp[0]), bq = __half22float2(hp[1]), cq = __half22float2(hp[2]), d = __half22float2(hp[3]);
  float4* dp = reinterpret_cast<float4*>(dst + b * s0 + h * s1 + t * s2 + c);
  dp[0] = make_float4(a.x, a.y, bq.x, bq.y);
  dp[1] = make_float4(cq.x, cq.y, d.x, d.y);
}
inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }
size_t fp32_workspace_bytes(const OehAttnArgs* a) {
  const size_t q = align256((size_t)a->B * a->H * a->Tq * HD * 2), kv = align256((size_t)a->B * a->H * a->Tk * HD * 2);
  return 2 * q + 2 * kv;
}

int validate(const OehAttnArgs* a) {
  OEH_CHECK_ARG(a, "oeh_attn_fwd: null args");
  OEH_CHECK_ARG(a->B >= 0 && a->H >= 0 && a->Tq >= 0 && a->Tk >= 0, "oeh_attn_fwd: negative dimension");
  OEH_CHECK_ARG(a->kind >= OEH_SOFTMAX_VANILLA && a->kind <= OEH_SOFTMAX_CLIPPED1, "oeh_attn_fwd: unknown softmax kind %d",
                a->kind);
  if (a->d != HD) return set_error(OEH_ERR_UNSUPPORTED, "oeh_attn_fwd: head dim %d (only %d is built)", a->d, HD);
  if (a->dtype != OEH_F16 && a->dtype != OEH_BF16 && a->dtype != OEH_F32)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_attn_fwd: dtype %d", a->dtype);
  if (a->causal && a->Tk < a->Tq)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_attn_fwd: causal with Tk < Tq");
  if ((a->qdq_scores || a->qdq_probs || a->qdq_ctx) && (a->n_bits < 1 || a->n_bits > 16))
    return set_error(OEH_ERR_BAD_ARG, "oeh_attn_fwd: n_bits=%d outside [1,16]", a->n_bits);
  return OEH_OK;
}

}  // namespace
}  // namespace oeh

using namespace oeh;

extern "C" size_t oeh_attn_workspace_bytes(const OehAttnArgs* args) {
  // row statistics of the two-pass modes stay in registers: 16-bit tensors need no workspace; fp32 tensors need room for
  // the fp16 copies of q, k, v and the fp16 context
  if (args && args->dtype == OEH_F32 && args->B > 0 && args->H > 0 && args->Tq > 0 && args->Tk > 0) return fp32_workspace_bytes(args);
  return 0;
}

#ifdef OEH_TIMELINE
extern "C" int oeh_set_wsp_timeline(unsigned long long* buf) {
  return cudaMemcpyToSymbol(attn_wsp::g_wsp_timeline, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
extern "C" int oeh_set_ws_timeline(unsigned long long* buf) {
  return cudaMemcpyToSymbol(attn_ws::g_ws_timeline, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
extern "C" int oeh_set_mid_timeline(unsigned long long* buf) {
  return cudaMemcpyToSymbol(attn_mid::g_mid_timeline, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
extern "C" int oeh_set_timeline(unsigned long long* buf) {
  return cudaMemcpyToSymbol(attn_one::g_timeline, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int oeh_attn_fwd_debug(const OehAttnArgs* a, float* dbg_s, float* dbg_o) {
  int rc = validate(a);
  if (rc) return rc;
  if (a->B == 0 || a->H == 0 || a->Tq == 0) return OEH_OK;
  OEH_CHECK_ARG(a->Tk > 0, "oeh_attn_fwd: Tk == 0 (softmax over an empty dimension)");
  OEH_CHECK_ARG(a->q && a->k && a->v && a->out, "oeh_attn_fwd: null tensor pointer");
  OEH_CHECK_ARG(a->H <= 65535 && a->B <= 65535, "oeh_attn_fwd: B and H must be <= 65535");
  const int vec = a->dtype == OEH_F32 ? 4 : 8;                   // elements per 16 bytes
  if ((reinterpret_cast<uintptr_t>(a->out) & 15) || a->out_strides[0] % vec || a->out_strides[1] % vec || a->out_strides[2] % vec)
    return set_error(OEH_ERR_UNSUPPORTED, "oeh_attn_fwd: out must be 16-byte aligned with strides that are multiples of %d", vec);

  if (a->dtype == OEH_F32) {
    for (int i = 0; i < 3; ++i)
      if (a->q_strides[i] % 4 || a->k_strides[i] % 4 || a->v_strides[i] % 4 || a->out_strides[i] % 4)
        return set_error(OEH_ERR_UNSUPPORTED, "oeh_attn_fwd: fp32 tensors need strides that are multiples of 4 elements");
    if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v)) & 15)
      return set_error(OEH_ERR_UNSUPPORTED, "oeh_attn_fwd: fp32 tensors must be 16-byte aligned");
    const size_t need = fp32_workspace_bytes(a);
    if (!a->workspace || a->workspace_bytes < need || (reinterpret_cast<uintptr_t>(a->workspace) & 255))
      return set_error(OEH_ERR_WORKSPACE, "oeh_attn_fwd: fp32 tensors need a 256-byte aligned workspace of %zu bytes", need);
    const size_t qb = align256((size_t)a->B * a->H * a->Tq * HD * 2), kvb = align256((size_t)a->B * a->H * a->Tk * HD * 2);
    uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
    __half *q16 = reinterpret_cast<__half*>(ws), *k16 = reinterpret_cast<__half*>(ws + qb),
           *v16 = reinterpret_cast<__half*>(ws + qb + kvb), *o16 = reinterpret_cast<__half*>(ws + qb + 2 * kvb);
    cudaStream_t st = (cudaStream_t)a->stream;
    const int64_t rq = (int64_t)a->B * a->H * a->Tq, rk = (int64_t)a->B * a->H * a->Tk;
    auto blocks = [](int64_t rows) { return (unsigned)((rows * 8 + 255) / 256); };
    cvt_rows_f32_to_f16<<<blocks(rq), 256, 0, st>>>((const float*)a->q, q16, a->H, a->Tq, a->q_strides[0], a->q_strides[1], a->q_strides[2], rq);
    cvt_rows_f32_to_f16<<<blocks(rk), 256, 0, st>>>((const float*)a->k, k16, a->H, a->Tk, a->k_strides[0], a->k_strides[1], a->k_strides[2], rk);
    cvt_rows_f32_to_f16<<<blocks(rk), 256, 0, st>>>((const float*)a->v, v16, a->H, a->Tk, a->v_strides[0], a->v_strides[1], a->v_strides[2], rk);
    if ((rc = check_launch("cvt_rows_f32_to_f16"))) return rc;
    OehAttnArgs h = *a;
    h.dtype = OEH_F16;
    h.q = q16; h.k = k16; h.v = v16; h.out = o16;
    const int64_t sq[3] = {(int64_t)a->H * a->Tq * HD, (int64_t)a->Tq * HD, HD}, sk[3] = {(int64_t)a->H * a->Tk * HD, (int64_t)a->Tk * HD, HD};
    for (int i = 0; i < 3; ++i) { h.q_strides[i] = sq[i]; h.out_strides[i] = sq[i]; h.k_strides[i] = sk[i]; h.v_strides[i] = sk[i]; }
    h.workspace = nullptr;
    h.workspace_bytes = 0;
    h.clamp_min = a->clamp_min ? 2 : 0;           // 2: clamp at finfo(float32).min, the dtype the caller's scores have
    if ((rc = oeh_attn_fwd_debug(&h, dbg_s, dbg_o))) return rc;
    cvt_rows_f16_to_f32<<<blocks(rq), 256, 0, st>>>(o16, (float*)a->out, a->H, a->Tq, a->out_strides[0], a->out_strides[1], a->out_strides[2], rq);
    return check_launch("cvt_rows_f16_to_f32");
  }

  // Kernel choice: sequences longer than one Q tile use the warp-specialised two-tile kernel (64-key sub-tiles);
  // up to 128 query rows there is nothing to ping-pong with, and the single-tile kernel (2 CTAs / SM) has the
  // lower latency.  OEH_ATTN_KERNEL=tile|ws overrides (tests exercise both on the same shapes).
  // When all keys fit one K/V tile (Tk <= 128, the BERT eval shape) the small one-tile kernel (4 CTAs / SM) runs.
  bool use_ws = a->Tq > BM, use_one = a->Tk <= BN;
  const bool norm_mode = a->kind >= OEH_SOFTMAX_CLIPPED || a->qdq_probs != nullptr;   // two-pass variants
  if (const long long f = tune(kTuneAttnKernel); f >= 0) {
    if (f == 0) use_ws = use_one = false;
    if (f == 1) { use_ws = true; use_one = false; }
  }
  if (use_one) use_ws = false;
  // 256 < Tk <= 512: the whole score row block fits TMEM, so the single-pass kernel of attn_fwd_mid.cuh can replace the
  // two-pass one.  Measured (profiles/r02_mid_kernel_4perrow.jsonl, H=12, four threads per row), mid vs the default: with
  // the INT8 taps B=4 T=512 18.3 vs 22.2 us causal (the config-4 shape) and 25.2 vs 28.3 us non-causal, B=16 T=384 43.5 vs
  // 43.7, but B=8 T=256 16.8 vs 16.1 and B=32 T=512 causal 110 vs 101 us; clipped without taps 15.0 vs 14.8 us; the
  // one-pass kinds are faster on the persistent kernel (11.6 vs 15.0 us).  One CTA per SM (512 TMEM columns) costs it the
  // overlap the two-pass kernel gets from two co-resident CTAs, so it is the default only where it wins: tapped softmax,
  // more than two key tiles, at most four waves of CTAs.  OEH_MID_KERNEL=0 disables it, =1 sends every kind with
  // Tk <= 512 there (=3: the two-threads-per-row shape).
  const long long row_blocks = (long long)a->B * a->H * ((a->Tq + BM - 1) / BM);
  // Since the kernel walks several items per CTA (no second wave of CTAs, the next item's loads under the current one) it is
  // ahead for every normalised kind at every measured size (profiles/r02_mid_kernel_z.jsonl: tapped B=32 T=512 causal 77 vs 91 us,
  // B=16 T=384 33 vs 40, B=8 T=256 13.7 vs 14.4; clippedsoftmax1 52.6 vs 61.5, 23.4 vs 26.5, 10.8 vs 10.8); the one-pass kinds stay
  // on the warp-specialised kernel (10.5 vs 11.7 us at B=4 T=512 causal).
  (void)row_blocks;
  bool use_mid = norm_mode && !use_one && a->Tk > BN && a->Tk <= attn_mid::kMaxTiles * BN;
  if (const long long f = tune(kTuneMidKernel); f >= 0) use_mid = f != 0 && !use_one && a->Tk <= attn_mid::kMaxTiles * BN;
  if (tune(kTuneAttnKernel) >= 0) use_mid = false;
  if (use_mid) use_ws = false;
  const int kv_box = use_ws ? attn_ws::BNW : BN;

  CUtensorMap tmq, tmk, tmv;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  if ((rc = make_tmap(&tmq, &p.cq, a->q, a->dtype, a->B, a->H, a->Tq, a->q_strides, BM, "q"))) return rc;
  if ((rc = make_tmap(&tmk, &p.ck, a->k, a->dtype, a->B, a->H, a->Tk, a->k_strides, kv_box, "k"))) return rc;
  if ((rc = make_tmap(&tmv, &p.cv, a->v, a->dtype, a->B, a->H, a->Tk, a->v_strides, kv_box, "v"))) return rc;
  p.B = a->B; p.H = a->H; p.Tq = a->Tq; p.Tk = a->Tk;
  p.scale = a->scale;
  p.kind = a->kind;
  p.stretch = (float)(a->eta - a->gamma);
  p.gamma = (float)a->gamma;
  p.causal = a->causal;
  p.mask = a->mask;
  p.mask_dtype = a->mask_dtype;
  if (a->mask) { p.ms0 = a->mask_strides[0]; p.ms1 = a->mask_strides[1]; p.ms2 = a->mask_strides[2]; }
  p.clamp_min = a->clamp_min;
  p.fmin = (a->dtype == OEH_F16 && a->clamp_min != 2) ? -65504.f : -3.3895313892515355e38f;   // bf16 and fp32 share the exponent range
  p.gate_logits = a->gate_logits;
  p.gate_scale = a->gate_scale;
  p.qdq_scores = a->qdq_scores;
  p.qdq_probs = a->qdq_probs;
  p.qdq_ctx = a->qdq_ctx;
  p.int_max = (float)((1 << (a->n_bits > 0 ? a->n_bits : 8)) - 1);
  p.ctx_qdq_before_gate = a->ctx_qdq_before_gate;
  p.out = a->out;
  p.os0 = a->out_strides[0]; p.os1 = a->out_strides[1]; p.os2 = a->out_strides[2];
  p.dbg_s = dbg_s;
  p.dbg_o = dbg_o;

  cudaStream_t st = (cudaStream_t)a->stream;
  const bool generic = a->qdq_scores || a->qdq_probs || a->qdq_ctx || a->mask || a->clamp_min || dbg_s || dbg_o ||
                       !(a->scale > 0.f);
  using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const AttnParams);
  const int ki = (a->dtype == OEH_BF16 ? 2 : 0) + (generic ? 1 : 0);
  if (use_mid) {
    // threads per query row: 4 (16 softmax warps) halves a row's serial work once more; OEH_MID_KERNEL=3 selects 2
    static const KernelFn kernels[2][4] = {
        {attn_mid::attn_fwd_mid_kernel<false, false, 2>, attn_mid::attn_fwd_mid_kernel<false, true, 2>,
         attn_mid::attn_fwd_mid_kernel<true, false, 2>, attn_mid::attn_fwd_mid_kernel<true, true, 2>},
        {attn_mid::attn_fwd_mid_kernel<false, false, 4>, attn_mid::attn_fwd_mid_kernel<false, true, 4>,
         attn_mid::attn_fwd_mid_kernel<true, false, 4>, attn_mid::attn_fwd_mid_kernel<true, true, 4>}};
    const int four = tune(kTuneMidKernel) == 3 ? 0 : 1;
    KernelFn fn = kernels[four][ki];
    if ((rc = ensure_smem(reinterpret_cast<const void*>(fn), attn_mid::kSmemBytes))) return rc;
    p.n_qt = (int)((a->Tq + BM - 1) / BM);
    p.hb_chunk = a->H * a->B;                      // causal: ALL heaviest Q tiles first (K / V of a launch fit L2 many times over)
    // one CTA per SM at most (512 TMEM columns each); a CTA walks its share of the (Q tile, head, batch) items
    const long long n_items = (long long)p.n_qt * a->H * a->B;
    const dim3 grid((unsigned)(n_items < num_sms() ? n_items : num_sms()));
    fn<<<grid, four ? attn_mid::Shape<4>::kThreads : attn_mid::Shape<2>::kThreads, attn_mid::kSmemBytes, st>>>(tmq, tmk, tmv, p);
    return check_launch("attn_fwd_mid_kernel");
  }
  if (use_ws) {
    const int kw = ki + (norm_mode ? 4 : 0);
    // Q tiles per CTA: 2 = one 640-thread CTA per SM, 1 = two independent 384-thread CTAs per SM (attn_fwd_ws.cuh, Cfg).
    // Two small CTAs hide each other's prologue / epilogue (~5.9 us per CTA) but load K / V once per tile and run their
    // steady state ~10 % slower; measured crossover (profiles/r01_ws_tiles_per_cta_l.txt): the two-pass modes are
    // faster with one tile at every length, the one-pass modes up to ~24 key steps per query tile on average.
    const double steps = (double)((a->Tk + attn_ws::BNW - 1) / attn_ws::BNW) * (a->causal ? 0.5 : 1.0);
    // Persistent variant (attn_fwd_wsp.cuh): hides the per-CTA prologue / epilogue.  Measured
    // (profiles/r01_ws_persistent_u.txt): one-pass non-causal +22 % at T = 512, +13 % at 1024, +7 % at 2048, +3 % at 4096,
    // equal at 8192.  Causal one-pass (profiles/r01_causal_order_w.txt; items sorted heaviest-first and walked in snake
    // order, see causal_item / item_of): +19 % at T = 512 and +12 % at 1024 as one-tile CTAs, +6 % at 2048 and +1 % at
    // 4096 as two-tile CTAs; at 8192 the per-item launch is 3 % ahead (the hardware's dynamic CTA scheduling balances the
    // long items better than any static walk).  The two-pass modes are faster as two one-tile CTAs per SM launched per
    // item at every length, so those keep the non-persistent kernel.
    int persistent = !norm_mode && (!a->causal || steps <= 32.0);
    // ... unless the whole launch is ONE wave of one-tile CTAs (two per SM): then a persistent CTA has nothing to hide and the
    // per-item form is 4 % ahead (B=4, H=12, T=512: 10.4 vs 10.8 us causal, 11.8 vs 12.4 us; profiles/r02_mid_kernel_u.jsonl)
    const long long one_tile_items = (long long)a->B * a->H * ((a->Tq + BM - 1) / BM);
    if (one_tile_items <= 2LL * num_sms()) persistent = 0;
#ifdef OEH_WS_PERSISTENT
    persistent = OEH_WS_PERSISTENT;
#endif
    if (const long long f = tune(kTuneWsPersistent); f >= 0) persistent = f != 0;
    int tiles = persistent ? ((a->causal && steps <= 8.0) ? 1 : 2) : ((norm_mode || steps <= 24.0) ? 1 : 2);
#ifdef OEH_WS_TILES
    tiles = OEH_WS_TILES;
#endif
    if (const long long f = tune(kTuneWsTiles); f >= 0) tiles = f == 1 ? 1 : 2;
    static const KernelFn kernels[2][8] = {
        {attn_ws::attn_fwd_ws_kernel<false, false, false, 1>, attn_ws::attn_fwd_ws_kernel<false, true, false, 1>,
         attn_ws::attn_fwd_ws_kernel<true, false, false, 1>,  attn_ws::attn_fwd_ws_kernel<true, true, false, 1>,
         attn_ws::attn_fwd_ws_kernel<false, false, true, 1>,  attn_ws::attn_fwd_ws_kernel<false, true, true, 1>,
         attn_ws::attn_fwd_ws_kernel<true, false, true, 1>,   attn_ws::attn_fwd_ws_kernel<true, true, true, 1>},
        {attn_ws::attn_fwd_ws_kernel<false, false, false, 2>, attn_ws::attn_fwd_ws_kernel<false, true, false, 2>,
         attn_ws::attn_fwd_ws_kernel<true, false, false, 2>,  attn_ws::attn_fwd_ws_kernel<true, true, false, 2>,
         attn_ws::attn_fwd_ws_kernel<false, false, true, 2>,  attn_ws::attn_fwd_ws_kernel<false, true, true, 2>,
         attn_ws::attn_fwd_ws_kernel<true, false, true, 2>,   attn_ws::attn_fwd_ws_kernel<true, true, true, 2>}};
    const int smem = tiles == 1 ? attn_ws::Cfg<1>::kSmemBytes : attn_ws::Cfg<2>::kSmemBytes;
    const int threads = tiles == 1 ? attn_ws::Cfg<1>::kThreads : attn_ws::Cfg<2>::kThreads;
    const int qrows = 128 * tiles;
    KernelFn fn = kernels[tiles - 1][kw];
    if ((rc = ensure_smem(reinterpret_cast<const void*>(fn), smem))) return rc;
    dim3 grid((unsigned)((a->Tq + qrows - 1) / qrows), (unsigned)a->H, (unsigned)a->B);
    p.n_qt = (int)grid.x;
    if (a->causal) {
      // causal work order (attn_common.cuh, causal_item): heaviest Q tiles first over chunks of (batch, head) pairs whose
      // K / V (Tk x 64 x 2 B x 2 per pair) stay within ~32 MB of L2.  OEH_CAUSAL_CHUNK=0 restores the per-head order.
      long long chunk = (32ll << 20) / ((long long)a->Tk * HD * 2 * 2);
      const long long hb = (long long)a->H * a->B;
      chunk = chunk < 1 ? 1 : (chunk > hb ? hb : chunk);
      if (const long long f = tune(kTuneCausalChunk); f >= 0) chunk = f > hb ? hb : f;
      p.hb_chunk = (int)chunk;
    }
    if (persistent) {
      // persistent variant (attn_fwd_wsp.cuh): one CTA per SM slot walks the work items with stride gridDim.x
      static const KernelFn pkernels[2][8] = {
          {attn_wsp::attn_fwd_wsp_kernel<false, false, false, 1>, attn_wsp::attn_fwd_wsp_kernel<false, true, false, 1>,
           attn_wsp::attn_fwd_wsp_kernel<true, false, false, 1>,  attn_wsp::attn_fwd_wsp_kernel<true, true, false, 1>,
           attn_wsp::attn_fwd_wsp_kernel<false, false, true, 1>,  attn_wsp::attn_fwd_wsp_kernel<false, true, true, 1>,
           attn_wsp::attn_fwd_wsp_kernel<true, false, true, 1>,   attn_wsp::attn_fwd_wsp_kernel<true, true, true, 1>},
          {attn_wsp::attn_fwd_wsp_kernel<false, false, false, 2>, attn_wsp::attn_fwd_wsp_kernel<false, true, false, 2>,
           attn_wsp::attn_fwd_wsp_kernel<true, false, false, 2>,  attn_wsp::attn_fwd_wsp_kernel<true, true, false, 2>,
           attn_wsp::attn_fwd_wsp_kernel<false, false, true, 2>,  attn_wsp::attn_fwd_wsp_kernel<false, true, true, 2>,
           attn_wsp::attn_fwd_wsp_kernel<true, false, true, 2>,   attn_wsp::attn_fwd_wsp_kernel<true, true, true, 2>}};
      const int psmem = tiles == 1 ? attn_wsp::Cfg<1>::kSmemBytes : attn_wsp::Cfg<2>::kSmemBytes;
      KernelFn pfn = pkernels[tiles - 1][kw];
      if ((rc = ensure_smem(reinterpret_cast<const void*>(pfn), psmem))) return rc;
      const long long items = (long long)grid.x * grid.y * grid.z;
      const long long slots = (long long)num_sms() * (tiles == 1 ? 2 : 1);
      const unsigned pgrid = (unsigned)(items < slots ? items : slots);
      pfn<<<pgrid, threads, psmem, st>>>(tmq, tmk, tmv, p);
      return check_launch("attn_fwd_wsp_kernel");
    }
    fn<<<grid, threads, smem, st>>>(tmq, tmk, tmv, p);
    return check_launch("attn_fwd_ws_kernel");
  }
  if (use_one) {
    // Two threads per query row (attn_fwd_one2.cuh): halves the serial softmax of a CTA.  Measured
    // (profiles/r02_one_tile_two_per_row.jsonl, H=12 T=128): with the three INT8 taps 12.4 vs 13.2 us at B=32 and 7.8 vs 9.0 us
    // at B=8, but 34.7 vs 33.1 us at B=128; without taps the register-resident one-thread kernel is as fast (8.57 vs 8.58 us
    // softmax1, 8.97 vs 9.10 us clipped) -- there the launch is bound by the HBM phase, not by the softmax chain.  Default:
    // tapped / masked launches of at most one wave of CTAs.  OEH_ONE_SPLIT=0 / 1 forces the choice.
    const long long one_items = (long long)((a->Tq + BM - 1) / BM) * a->H * a->B;
    int split = generic && one_items <= 3LL * num_sms();
    if (const long long f = tune(kTuneOneSplit); f >= 0) split = f != 0;
    if (tune(kTuneOneRegs) >= 0) split = 0;
    if (split) {
      static const KernelFn k2[4] = {attn_one2::attn_fwd_one2_kernel<false, false>, attn_one2::attn_fwd_one2_kernel<false, true>,
                                     attn_one2::attn_fwd_one2_kernel<true, false>, attn_one2::attn_fwd_one2_kernel<true, true>};
      if ((rc = ensure_smem(reinterpret_cast<const void*>(k2[ki]), attn_one2::kSmemBytes))) return rc;
      dim3 grid((unsigned)((a->Tq + BM - 1) / BM), (unsigned)a->H, (unsigned)a->B);
      k2[ki]<<<grid, attn_one2::kThreads, attn_one2::kSmemBytes, st>>>(tmq, tmk, tmv, p);
      return check_launch("attn_fwd_one2_kernel");
    }
    // [register-resident S row][bf16 | generic]
    static const KernelFn kernels[2][4] = {
        {attn_one::attn_fwd_one_kernel<false, false, false>, attn_one::attn_fwd_one_kernel<false, true, false>,
         attn_one::attn_fwd_one_kernel<true, false, false>, attn_one::attn_fwd_one_kernel<true, true, false>},
        {attn_one::attn_fwd_one_kernel<false, false, true>, attn_one::attn_fwd_one_kernel<false, true, true>,
         attn_one::attn_fwd_one_kernel<true, false, true>, attn_one::attn_fwd_one_kernel<true, true, true>}};
    // S row in registers (3 CTAs / SM, one TMEM read) vs chunked TMEM passes (4 CTAs / SM): 9.37 vs 9.88 us per launch at
    // the C-BERT shape for the clipped kinds, 9.29 vs 9.52 us for softmax1 (profiles/r01_one_tile_regs_y.txt; round 2, with the
    // coalesced context store: 9.09 / 8.55 us).  A persistent form -- one CTA per SM, three 128-thread item slots, one TMEM
    // allocation, all first-round loads in flight at once -- was built and measured in round 2 and is NOT faster (8.97 vs
    // 8.55 us, profiles/r02_one_tile_persistent_vs_item.jsonl): the launch is HBM-bound while its ~19 MB land, and the
    // per-item grid's natural stagger (the three CTAs of an SM enter ~0.9 us apart) lets the first CTA's softmax run under
    // the others' loads, while simultaneous slots all receive their last byte together (profiles/r02_timeline_one.txt).  The generic
    // instantiation (taps / masks) does not fit 168 registers without spilling and keeps the chunked passes.
    int regs = generic ? 0 : 1;
    if (const long long f = tune(kTuneOneRegs); f >= 0) regs = f != 0;
    if ((rc = ensure_smem(reinterpret_cast<const void*>(kernels[regs][ki]), attn_one::kSmemBytes))) return rc;
    dim3 grid((unsigned)((a->Tq + BM - 1) / BM), (unsigned)a->H, (unsigned)a->B);
    kernels[regs][ki]<<<grid, attn_one::BM, attn_one::kSmemBytes, st>>>(tmq, tmk, tmv, p);
    return check_launch("attn_fwd_one_kernel");
  }
  static const KernelFn kernels[4] = {attn_fwd_kernel<false, false>, attn_fwd_kernel<false, true>,
                                      attn_fwd_kernel<true, false>, attn_fwd_kernel<true, true>};
  if ((rc = ensure_smem(reinterpret_cast<const void*>(kernels[ki]), kSmemBytes))) return rc;
  dim3 grid((unsigned)((a->Tq + BM - 1) / BM), (unsigned)a->H, (unsigned)a->B);
  kernels[ki]<<<grid, BM, kSmemBytes, st>>>(tmq, tmk, tmv, p);
  return check_launch("attn_fwd_kernel");
}

extern "C" int oeh_attn_fwd(const OehAttnArgs* args) { return oeh_attn_fwd_debug(args, nullptr, nullptr); }

extern "C" int oeh_set_override(const char* name, long long value) {
  OEH_CHECK_ARG(name, "oeh_set_override: null name");
  tune_init();
  for (int k = 0; k < kTuneKeys; ++k)
    if (!strcmp(name, kTuneNames[k])) {
      g_tune[k].store(value < 0 ? -1 : value, std::memory_order_relaxed);
      return OEH_OK;
    }
  return set_error(OEH_ERR_BAD_ARG, "oeh_set_override: unknown key '%s'", name);
}
