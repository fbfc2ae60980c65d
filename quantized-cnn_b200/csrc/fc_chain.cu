// Latency path of the PQ fully-connected layers (batch 1..4): ONE persistent launch for a whole run of consecutive FC
// layers (AlexNet: fc6 -> ReLU -> fc7 -> ReLU -> fc8), bound by the HBM stream of the uint8 assignment matrices.
// Replaces CaffeEva::CalcFeatMap_FCntAprx + GetInPdMat (reference src/CaffeEva.cc:968-1025, 1261-1296) for each layer:
//     LUT[s][k] = sum_j x[s*d+j] * ctrd[s][k][j]            (ascending j, rounded mul then rounded add)
//     dst[o]    = bias[o] + sum_s LUT[s][asmt[s][o]]
//
// Design (one CTA per SM, cooperative launch so that all CTAs are co-resident):
//   * the subspaces (rows of the device assignment table [S][DoutPad]) of EVERY layer are split evenly over the CTAs,
//     so a CTA's share of a layer is one contiguous block of bytes.  A producer warp streams these blocks -- all
//     layers back to back, it never waits for the arithmetic -- into a shared-memory ring with cp.async.bulk
//     (mbarrier complete_tx); with ~150 KB of ring per SM the whole 17.7 MB of AlexNet's three FC layers is in flight
//     from the first cycle of the kernel, so the HBM stream is not interrupted by the layer-to-layer dependency.
//   * a CTA builds only the LUT rows of its own subspaces (its 1/148 of the codebook arrives by one bulk copy), then
//     512 consumer threads gather: thread = 4/8/16 consecutive output channels, one 32/64/128-bit shared-memory load
//     of assignment bytes per row, LUT row = one bank sweep for K <= 32 (conflict-free by construction).
//   * a CTA's result is a partial sum over ITS subspaces for ALL output channels.  It is published to an L2-resident
//     buffer [CTA][DoutPad]; the word itself is the ready flag (the buffer holds a NaN pattern no arithmetic produces
//     until it is written, and the one consumer of a word resets it), so there is no grid barrier, no fence and no
//     atomic between the layers: the consumer of the next layer polls exactly the 4-float groups it needs
//     (its own subspaces' inputs), lanes over the producing CTAs, and reduces them with a fixed shuffle tree
//     (deterministic; bias and ReLU are applied there).  The last layer's partials are reduced the same way, eight
//     output channels per CTA.
// Accumulation order: per CTA s ascending, then CTAs in a fixed tree -- a re-association of the reference's
// bias + s-ascending sum (fp32 adds only; tolerance in tests/test_gpu_layers.py).  fc_nsplit = 1 keeps the bit-exact kernel.
#include "qcnn_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int kConsumers = 512;
constexpr int kConsumerWarps = kConsumers / 32;
constexpr int kThreads = kConsumers + 32;
constexpr int kMaxChain = 4;
constexpr int kMaxParts = 160;       // 5 polls per lane
constexpr int kMaxStage = 32;
constexpr int kChunkBytes = 16384;
constexpr int kBarBytes = 1024;      // mbarrier block at the start of shared memory
constexpr uint32_t kSentinel = 0xFFFFFFFFu;

struct ChainLayer {
  const float* ctrd;      // [S][K][d]
  const uint8_t* asmt;    // [S][DoutPad], stored byte = idx << kshift
  const float* bias;
  const int* srcoff;      // first layer only: flattened feature -> source element offset (NULL: identity)
  float* partial;         // [G][DoutPad]
  int Din, Dout, DoutPad, S, K, d;
  int pre;                // stored byte = byte offset inside a LUT row (K <= 64)
  int lutPitch;           // bytes per LUT row in shared memory (256 when pre, else 4 K)
  int unit, unitsBase, unitsRem;   // subspace split in units of `unit` rows
  int cpt, tpr, rg;       // channels per thread, threads per row, row groups
  int rpc;                // rows per ring chunk
  int relu;               // ReLU on this layer's output
  int ctrdOff;            // float offset of the codebook slice in shared memory
};

struct ChainArgs {
  ChainLayer L[kMaxChain];
  const float* src;
  float* dst;
  unsigned long long* dbg;   // optional [2*G] globaltimer stamps (first / last instruction of every CTA)
  int nLayers, G, nStage;
  int xFloats, lutFloats, ctrdFloats, rgFloats, ringOff;
};

__device__ __forceinline__ uint32_t SmemU32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void MbarInit(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemU32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarArrive(uint64_t* b) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(SmemU32(b)) : "memory");
}
__device__ __forceinline__ void MbarExpectTx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemU32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(SmemU32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void BulkLoad(void* smemDst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(SmemU32(smemDst)), "l"(gsrc), "r"(bytes), "r"(SmemU32(bar)) : "memory");
}
__device__ __forceinline__ void ConsumerSync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }
__device__ __forceinline__ unsigned long long GlobalTimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint4 LdRelaxed4(const float* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void StRelaxed4(float* p, float a, float b, float c, float d) {
  asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ bool Pending(const uint4& v) {
  return v.x == kSentinel || v.y == kSentinel || v.z == kSentinel || v.w == kSentinel;
}

// subspaces [r0, r1) of layer L owned by CTA c
__device__ __forceinline__ void RowRange(const ChainLayer& L, int c, int* r0, int* r1) {
  const int u0 = c * L.unitsBase + min(c, L.unitsRem);
  const int nu = L.unitsBase + (c < L.unitsRem ? 1 : 0);
  *r0 = min(L.S, u0 * L.unit);
  *r1 = min(L.S, (u0 + nu) * L.unit);
}

// sum over the G producing CTAs of partial[p][f .. f+3] (waits for every word, then resets it); the same value in all lanes
__device__ __forceinline__ float4 CollectParts(float* partial, int pitch, int G, int f, int lane) {
  constexpr int NP = kMaxParts / 32;
  uint4 v[NP];
#pragma unroll
  for (int j = 0; j < NP; j++) {
    const int p = lane + 32 * j;
    if (p < G) v[j] = LdRelaxed4(partial + static_cast<size_t>(p) * pitch + f);
    else v[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
  for (int j = 0; j < NP; j++) {
    const int p = lane + 32 * j;
    if (p < G) {
      float* addr = partial + static_cast<size_t>(p) * pitch + f;
      while (Pending(v[j])) v[j] = LdRelaxed4(addr);
      *reinterpret_cast<uint4*>(addr) = make_uint4(kSentinel, kSentinel, kSentinel, kSentinel);   // armed for the next call
      s.x += __uint_as_float(v[j].x); s.y += __uint_as_float(v[j].y);
      s.z += __uint_as_float(v[j].z); s.w += __uint_as_float(v[j].w);
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    s.x += __shfl_xor_sync(0xFFFFFFFFu, s.x, m);
    s.y += __shfl_xor_sync(0xFFFFFFFFu, s.y, m);
    s.z += __shfl_xor_sync(0xFFFFFFFFu, s.z, m);
    s.w += __shfl_xor_sync(0xFFFFFFFFu, s.w, m);
  }
  return s;
}

template <int CPT> struct Idx;
template <> struct Idx<4>  { using type = uint32_t; };
template <> struct Idx<8>  { using type = uint2; };
template <> struct Idx<16> { using type = uint4; };
__device__ __forceinline__ uint32_t IdxWord(const uint32_t& v, int) { return v; }
__device__ __forceinline__ uint32_t IdxWord(const uint2& v, int i) { return i == 0 ? v.x : v.y; }
__device__ __forceinline__ uint32_t IdxWord(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

__device__ __forceinline__ float LdShared(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Steps (3) and (4) of a layer for one consumer thread: gather-accumulate the CTA's assignment rows as their ring chunks
// arrive (thread = CPT consecutive channels of the rows rgIdx, rgIdx + rg, ...), then publish the partial sums.
// PRE: the stored byte is already the byte offset inside a LUT row (K <= 64) and LUT rows are 256-byte aligned, so ONE
// byte-permute both extracts the index and forms the shared-memory address: 3 instructions per lookup (PRMT, LDS, FADD).
template <int CPT, bool PRE>
__device__ __forceinline__ void GatherAndPublish(const ChainLayer& L, int rows, int nStage, uint64_t* fullB, uint64_t* emptyB,
                                                 const uint8_t* ring, const float* lut, float* rgred, float* out, int tid,
                                                 int lane, int* itp) {
  using IV = typename Idx<CPT>::type;
  float acc[CPT];
#pragma unroll
  for (int c = 0; c < CPT; c++) acc[c] = 0.0f;
  const int rgIdx = tid / L.tpr, cgIdx = tid - rgIdx * L.tpr;
  const bool active = rgIdx < L.rg;
  const int rowBytes = L.lutPitch, pitch = L.DoutPad, rg = L.rg;
  const uint32_t lutBase = SmemU32(lut);
  int it = *itp;
  for (int r = 0; r < rows; r += L.rpc, it++) {
    const int stage = it % nStage;
    MbarWait(fullB + stage, (it / nStage) & 1);
    if (active) {
      const uint8_t* chunk = ring + static_cast<size_t>(stage) * kChunkBytes + cgIdx * CPT;
      const int n = min(L.rpc, rows - r);
#pragma unroll 4
      for (int q = rgIdx; q < n; q += rg) {
        const IV w = *reinterpret_cast<const IV*>(chunk + static_cast<size_t>(q) * pitch);
        const uint32_t lr = lutBase + static_cast<uint32_t>(r + q) * rowBytes;
#pragma unroll
        for (int c = 0; c < CPT; c++) {
          uint32_t addr;
          if (PRE) addr = __byte_perm(IdxWord(w, c >> 2), lr, 0x7650u | (c & 3));          // (lr & ~0xFF) | byte
          else addr = lr + (__byte_perm(IdxWord(w, c >> 2), 0u, 0x4440u | (c & 3)) << 2);
          acc[c] += LdShared(addr);
        }
      }
    }
    __syncwarp();
    if (lane == 0) MbarArrive(emptyB + stage);
  }
  *itp = it;
  if (rg > 1) {
    if (active) {
      float* mine = rgred + static_cast<size_t>(rgIdx) * pitch + cgIdx * CPT;
#pragma unroll
      for (int c = 0; c < CPT; c += 4) *reinterpret_cast<float4*>(mine + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
    }
    ConsumerSync();
    for (int q = tid; q < (pitch >> 2); q += kConsumers) {
      float4 s = *reinterpret_cast<const float4*>(rgred + 4 * q);
      for (int g = 1; g < rg; g++) {
        const float4 t = *reinterpret_cast<const float4*>(rgred + static_cast<size_t>(g) * pitch + 4 * q);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      StRelaxed4(out + 4 * q, s.x, s.y, s.z, s.w);
    }
  } else if (active) {
    float* o = out + cgIdx * CPT;
#pragma unroll
    for (int c = 0; c < CPT; c += 4) StRelaxed4(o + c, acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
  }
}

__global__ void __launch_bounds__(kThreads, 1) fc_chain_kernel(const __grid_constant__ ChainArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* fullB = reinterpret_cast<uint64_t*>(sm);
  uint64_t* emptyB = fullB + kMaxStage;
  uint64_t* ctrdB = emptyB + kMaxStage;
  float* xs = reinterpret_cast<float*>(sm + kBarBytes);
  float* lut = xs + a.xFloats;
  float* ctrdS = lut + a.lutFloats;
  float* rgred = ctrdS + a.ctrdFloats;
  uint8_t* ring = sm + a.ringOff;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, cta = blockIdx.x;
  if (a.dbg && tid == 0) a.dbg[2 * cta] = GlobalTimer();
  if (tid == 0) {
    for (int i = 0; i < a.nStage; i++) { MbarInit(fullB + i, 1); MbarInit(emptyB + i, kConsumerWarps); }
    for (int l = 0; l < a.nLayers; l++) MbarInit(ctrdB + l, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ---- producer: codebook slices of every layer, then the assignment rows of every layer through the ring ----
    if (lane == 0) {
      for (int l = 0; l < a.nLayers; l++) {
        const ChainLayer& L = a.L[l];
        int r0, r1;
        RowRange(L, cta, &r0, &r1);
        if (r1 > r0) {
          const uint32_t bytes = static_cast<uint32_t>(r1 - r0) * L.K * L.d * 4u;
          MbarExpectTx(ctrdB + l, bytes);
          BulkLoad(ctrdS + L.ctrdOff, L.ctrd + static_cast<size_t>(r0) * L.K * L.d, bytes, ctrdB + l);
        }
      }
      int it = 0;
      for (int l = 0; l < a.nLayers; l++) {
        const ChainLayer& L = a.L[l];
        int r0, r1;
        RowRange(L, cta, &r0, &r1);
        for (int r = r0; r < r1; r += L.rpc, it++) {
          const int stage = it % a.nStage, round = it / a.nStage;
          if (round > 0) MbarWait(emptyB + stage, (round - 1) & 1);
          const uint32_t bytes = static_cast<uint32_t>(min(L.rpc, r1 - r)) * L.DoutPad;
          MbarExpectTx(fullB + stage, bytes);
          BulkLoad(ring + static_cast<size_t>(stage) * kChunkBytes, L.asmt + static_cast<size_t>(r) * L.DoutPad, bytes, fullB + stage);
        }
      }
    }
    return;
  }

  // ---- consumers ----
  int it = 0;
  for (int l = 0; l < a.nLayers; l++) {
    const ChainLayer& L = a.L[l];
    int r0, r1;
    RowRange(L, cta, &r0, &r1);
    const int rows = r1 - r0;
    const int f0 = r0 * L.d, nf = rows * L.d;
    // (1) this CTA's slice of the layer input
    if (l == 0) {
      for (int i = tid; i < nf; i += kConsumers) {
        const int f = f0 + i;
        xs[i] = f < L.Din ? __ldg(a.src + (L.srcoff ? __ldg(L.srcoff + f) : f)) : 0.0f;
      }
    } else {
      const ChainLayer& P = a.L[l - 1];
      const int groups = (nf + 3) >> 2;
      for (int g = warp; g < groups; g += kConsumerWarps) {
        const int f = f0 + 4 * g;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (f < L.Din) {   // Din == P.Dout <= P.DoutPad (multiple of 16): the whole group lies inside a partial row
          v = CollectParts(P.partial, P.DoutPad, a.G, f, lane);
          float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            o[j] = (f + j < L.Din) ? o[j] + __ldg(P.bias + f + j) : 0.0f;
            if (P.relu) o[j] = fmaxf(o[j], 0.0f);
          }
          v = make_float4(o[0], o[1], o[2], o[3]);
        }
        if (lane == 0) *reinterpret_cast<float4*>(xs + 4 * g) = v;
      }
    }
    ConsumerSync();
    // (2) LUT rows of the CTA's subspaces (reference GetInPdMat: ascending j, separate multiply and add)
    if (rows > 0) {
      MbarWait(ctrdB + l, 0);
      const float* cs = ctrdS + L.ctrdOff;
      const int d = L.d;
      for (int e = tid; e < rows * L.K; e += kConsumers) {
        const int r = e / L.K;
        const int dims = min(d, L.Din - (r0 + r) * d);
        const float* c = cs + static_cast<size_t>(e) * d;
        const float* x = xs + r * d;
        float v = 0.0f;
        if (d == 4 && dims == 4) {
          const float4 cv = *reinterpret_cast<const float4*>(c);
          const float4 xv = *reinterpret_cast<const float4*>(x);
          v = __fmul_rn(xv.x, cv.x);
          v = __fadd_rn(v, __fmul_rn(xv.y, cv.y));
          v = __fadd_rn(v, __fmul_rn(xv.z, cv.z));
          v = __fadd_rn(v, __fmul_rn(xv.w, cv.w));
        } else {
          for (int j = 0; j < dims; j++) v = __fadd_rn(v, __fmul_rn(x[j], c[j]));
        }
        lut[r * (L.lutPitch >> 2) + (e - r * L.K)] = v;
      }
    }
    ConsumerSync();
    // (3) + (4) gather-accumulate the assignment rows as they arrive, publish the partial sums of all DoutPad channels
    float* out = L.partial + static_cast<size_t>(cta) * L.DoutPad;
#define QCNN_GP(C, P) GatherAndPublish<C, P>(L, rows, a.nStage, fullB, emptyB, ring, lut, rgred, out, tid, lane, &it)
    if (L.pre) { if (L.cpt == 8) QCNN_GP(8, true); else if (L.cpt == 4) QCNN_GP(4, true); else QCNN_GP(16, true); }
    else { if (L.cpt == 8) QCNN_GP(8, false); else if (L.cpt == 4) QCNN_GP(4, false); else QCNN_GP(16, false); }
#undef QCNN_GP
  }

  // ---- output of the last layer: a multiple of four channels per CTA ----
  {
    const ChainLayer& P = a.L[a.nLayers - 1];
    const int cpc = ((P.Dout + a.G - 1) / a.G + 3) & ~3;
    const int o0 = cta * cpc;
    const int n = max(0, min(cpc, P.Dout - o0));
    for (int g = warp; g < ((n + 3) >> 2); g += kConsumerWarps) {
      const int o = o0 + 4 * g;
      const float4 v = CollectParts(P.partial, P.DoutPad, a.G, o, lane);
      if (lane == 0) {
        const float r[4] = {v.x, v.y, v.z, v.w};
        for (int j = 0; j < 4 && o + j < P.Dout; j++) {
          const float t = r[j] + __ldg(P.bias + o + j);
          a.dst[o + j] = P.relu ? fmaxf(t, 0.0f) : t;
        }
      }
    }
  }
  if (a.dbg && tid == 0) a.dbg[2 * cta + 1] = GlobalTimer();
}

int Gcd(int a, int b) { return b ? Gcd(b, a % b) : a; }

}  // namespace

namespace qcnn {

// Largest batch the chain kernel is launched for (once per image; the assignment bytes of the later images come from L2)
constexpr int kChainMaxN = 4;

static bool ChainEnabled() {
  static const bool on = !(getenv("QCNN_FC_CHAIN") && getenv("QCNN_FC_CHAIN")[0] == '0');
  return on;
}

// Fills the kernel arguments for a run of FC layers; false when the kernel does not apply (the caller falls back to
// the per-layer kernels of fc_aprx.cu).
static bool PlanChain(qcnn_ctx* ctx, qcnn_layer* const* layers, const int* relu, int n, ChainArgs* out, size_t* smemBytes) {
  if (!ChainEnabled() || n < 1 || n > kMaxChain) return false;
  const int G = ctx->sm_count;
  if (G < 1 || G > kMaxParts) return false;
  ChainArgs& a = *out;
  memset(&a, 0, sizeof(a));
  a.nLayers = n; a.G = G;
  int xF = 64, lutF = 64, ctrdF = 0, rgF = 0;
  for (int l = 0; l < n; l++) {
    const qcnn_layer* Q = layers[l];
    if (Q->kind != QCNN_KIND_FC || Q->opt_fc_nsplit || Q->opt_fc_tn) return false;
    if (l > 0 && (Q->Din != layers[l - 1]->Dout || Q->d_srcoff)) return false;
    if (Q->DoutPad > kChunkBytes || Q->K * Q->d % 4 != 0) return false;
    ChainLayer& L = a.L[l];
    L.ctrd = Q->d_ctrd; L.asmt = Q->d_asmt; L.bias = Q->d_bias; L.srcoff = Q->d_srcoff; L.partial = Q->d_cpart;
    L.Din = Q->Din; L.Dout = Q->Dout; L.DoutPad = Q->DoutPad; L.S = Q->S; L.K = Q->K; L.d = Q->d;
    L.pre = Q->kshift == 2 ? 1 : 0;
    L.lutPitch = L.pre ? 256 : Q->K * 4;
    L.relu = relu[l];
    L.cpt = (Q->DoutPad / 4 <= kConsumers) ? 4 : ((Q->DoutPad / 8 <= kConsumers) ? 8 : 16);
    if (Q->DoutPad / L.cpt > kConsumers) return false;
    L.tpr = Q->DoutPad / L.cpt;
    L.rg = std::min(kConsumers / L.tpr, 16);
    L.rpc = std::max(1, kChunkBytes / Q->DoutPad);
    L.unit = 4 / Gcd(Q->d, 4);
    const int units = CeilDiv(Q->S, L.unit);
    L.unitsBase = units / G; L.unitsRem = units % G;
    const int rowsMax = std::min(Q->S, (L.unitsBase + (L.unitsRem ? 1 : 0)) * L.unit);
    L.ctrdOff = ctrdF;
    ctrdF += rowsMax * Q->K * Q->d;
    xF = std::max(xF, RoundUp(rowsMax * Q->d, 64));       // keeps the LUT 256-byte aligned
    lutF = std::max(lutF, RoundUp(rowsMax * (L.lutPitch / 4), 64));
    if (L.rg > 1) rgF = std::max(rgF, L.rg * Q->DoutPad);
  }
  a.xFloats = xF; a.lutFloats = lutF; a.ctrdFloats = ctrdF; a.rgFloats = rgF;
  const size_t fixed = kBarBytes + sizeof(float) * (static_cast<size_t>(xF) + lutF + ctrdF + rgF);
  a.ringOff = static_cast<int>((fixed + 127) & ~static_cast<size_t>(127));
  const size_t cap = ctx->smem_optin ? ctx->smem_optin : 227 * 1024;
  if (static_cast<size_t>(a.ringOff) + 2 * kChunkBytes > cap) return false;
  a.nStage = static_cast<int>(std::min<size_t>(kMaxStage, (cap - a.ringOff) / kChunkBytes));
  *smemBytes = a.ringOff + static_cast<size_t>(a.nStage) * kChunkBytes;
  return true;
}

bool FcChainEligible(qcnn_ctx* ctx, qcnn_layer* const* layers, const int* relu, int n, int N) {
  if (N < 1 || N > kChainMaxN) return false;
  ChainArgs a;
  size_t smem;
  return PlanChain(ctx, layers, relu, n, &a, &smem);
}

void DescribeFcChain(qcnn_ctx* ctx, qcnn_layer* const* layers, const int* relu, int n, char* buf, size_t cap) {
  ChainArgs a;
  size_t smem = 0;
  buf[0] = 0;
  if (!PlanChain(ctx, layers, relu, n, &a, &smem)) return;
  snprintf(buf, cap, "fc_chain(persistent, %d layer%s, grid=%d, ring=%dx%dB, smem=%zuB)", n, n > 1 ? "s" : "", a.G, a.nStage,
           kChunkBytes, smem);
}

// src: [N][Din of the first layer] (or the NHWC map the first layer's srcoff folds); dst: [N][Dout of the last layer]
int LaunchFcChain(qcnn_ctx* ctx, qcnn_layer* const* layers, const int* relu, int n, const float* src, int N, float* dst,
                  cudaStream_t st, unsigned long long* dbg, bool* handled) {
  *handled = false;
  if (N < 1 || N > kChainMaxN) return 0;
  // partial-sum buffers: one per layer, [G][DoutPad], every word armed with the "not yet written" pattern
  for (int l = 0; l < n; l++) {
    qcnn_layer* Q = layers[l];
    if (Q->kind != QCNN_KIND_FC) return 0;
    const size_t need = sizeof(float) * static_cast<size_t>(ctx->sm_count) * Q->DoutPad;
    if (Q->cpart_bytes < need) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      if (st) cudaStreamIsCapturing(st, &cap);
      if (cap != cudaStreamCaptureStatusNone) return 0;   // cannot allocate inside a capture: per-layer kernels this time
      if (Q->d_cpart) QCNN_CUDA(cudaFree(Q->d_cpart));
      Q->d_cpart = nullptr; Q->cpart_bytes = 0;
      QCNN_CUDA(cudaMalloc(&Q->d_cpart, need));
      QCNN_CUDA(cudaMemsetAsync(Q->d_cpart, 0xFF, need, st));
      Q->cpart_bytes = need;
      ctx->alloc_epoch++;
    }
  }
  ChainArgs a;
  size_t smem = 0;
  if (!PlanChain(ctx, layers, relu, n, &a, &smem)) return 0;
  static bool attrSet[64] = {false};
  if (ctx->device < 64 && !attrSet[ctx->device]) {
    QCNN_CUDA(cudaFuncSetAttribute(fc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(ctx->smem_optin ? ctx->smem_optin : 227 * 1024)));
    attrSet[ctx->device] = true;
  }
  a.dbg = dbg;
  const size_t srcImg = layers[0]->Din, dstImg = layers[n - 1]->Dout;
  for (int i = 0; i < N; i++) {
    a.src = src + i * srcImg;
    a.dst = dst + i * dstImg;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(a.G); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident (they wait for each other's partial sums)
    attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    QCNN_CUDA(cudaLaunchKernelEx(&cfg, fc_chain_kernel, a));
    ctx->launches++;
  }
  *handled = true;
  return 0;
}

}  // namespace qcnn
