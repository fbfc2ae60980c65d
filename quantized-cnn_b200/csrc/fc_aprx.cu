// Fused PQ fully-connected layer for sm_100a: LUT stage + uint8 gather-accumulate in ONE kernel.
// Replaces CaffeEva::CalcFeatMap_FCntAprx + GetInPdMat (reference src/CaffeEva.cc:968-1025, 1261-1296):
//     LUT[n][s][k] = sum_j x[n][s*d+j] * ctrd[s][k][j]        (ascending j, rounded mul then rounded add)
//     dst[n][o]    = bias[o] + sum_s LUT[n][s][asmt[s][o]]    (ascending s)
//
// Mapping (HBM/L2 stream of the assignment matrix is the only large operand):
//   * lane = output channel.  A thread owns CPT consecutive channels and TN images -> TN*CPT accumulators.
//   * the device assignment table is [S][DoutPad] bytes, so a warp reads 32*CPT contiguous bytes per subspace
//     row (128-bit loads at CPT=16); rows of one chunk (PF subspaces) are prefetched into registers BEFORE the
//     chunk's LUT slice is built, so the HBM stream overlaps the LUT arithmetic.
//   * the LUT slice of the chunk ([TN][PF][K] floats) lives in shared memory.  With K <= 32 one LUT row is
//     <= 128 B = one bank sweep: distinct codewords hit distinct banks and equal codewords broadcast, so the
//     lane-dependent gather is bank-conflict free by construction.
//   * assignments are stored pre-multiplied by 4 (byte offsets) when K <= 64, saving the shift per lookup.
//   * S can be split over blockIdx.z (needed at small N to fill 148 SMs); partial sums are reduced in a fixed
//     order by fc_reduce_kernel (deterministic; no atomics).  With nsplit == 1 the accumulation order is
//     exactly the reference's (bias, then s ascending) and the result is bit-identical to the CPU path.
#include "qcnn_internal.h"

namespace {

constexpr int kFcThreads = 256;

template <int CPT> struct AsmtVec;
template <> struct AsmtVec<4>  { using type = uint32_t; };
template <> struct AsmtVec<8>  { using type = uint2; };
template <> struct AsmtVec<16> { using type = uint4; };

__device__ __forceinline__ uint32_t LoadStream(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 LoadStream(const uint2* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 LoadStream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void Zero(uint32_t& v) { v = 0u; }
__device__ __forceinline__ void Zero(uint2& v) { v = make_uint2(0u, 0u); }
__device__ __forceinline__ void Zero(uint4& v) { v = make_uint4(0u, 0u, 0u, 0u); }
__device__ __forceinline__ uint32_t Word(const uint32_t& v, int) { return v; }
__device__ __forceinline__ uint32_t Word(const uint2& v, int i) { return i == 0 ? v.x : v.y; }
__device__ __forceinline__ uint32_t Word(const uint4& v, int i) {
  return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// K: codewords per subspace; CPT: channels per thread; TN: images per CTA; PF: subspaces per chunk;
// PRE: assignments stored as byte offsets (idx*4)
template <int K, int CPT, int TN, int PF, bool PRE>
__global__ void __launch_bounds__(kFcThreads, 2) fc_aprx_kernel(const FcArgs a) {
  extern __shared__ __align__(16) float lut[];  // [TN][PF][K]
  using AV = typename AsmtVec<CPT>::type;
  const int tid = threadIdx.x;
  const int o0 = (blockIdx.x * kFcThreads + tid) * CPT;
  const int n0 = blockIdx.y * TN;
  const int split = blockIdx.z;
  const int s_begin = split * a.s_per_split;
  const int s_end = min(a.S, s_begin + a.s_per_split);
  const bool live = o0 < a.DoutPad;

  float acc[TN][CPT];
#pragma unroll
  for (int c = 0; c < CPT; c++) {
    const float b = (split == 0 && o0 + c < a.Dout) ? __ldg(a.bias + o0 + c) : 0.0f;
#pragma unroll
    for (int nl = 0; nl < TN; nl++) acc[nl][c] = b;
  }

  for (int sc = s_begin; sc < s_end; sc += PF) {
    // (1) start the assignment stream for this chunk
    AV areg[PF];
#pragma unroll
    for (int r = 0; r < PF; r++) {
      const int s = sc + r;
      if (live && s < s_end) {
        areg[r] = LoadStream(reinterpret_cast<const AV*>(a.asmt + static_cast<size_t>(s) * a.DoutPad + o0));
      } else {
        Zero(areg[r]);
      }
    }
    // (2) build the LUT slice of this chunk.  A thread owns (subspace r, codeword k) pairs: it loads the pair's
    //     codebook row and source offsets once and sweeps the TN images (x loads are warp-broadcast).
    __syncthreads();  // the previous chunk's gather is done with `lut`
    for (int pr = tid; pr < PF * K; pr += kFcThreads) {
      const int k = pr % K;
      const int r = pr / K;
      const int s = sc + r;
      const int f0 = s * a.d;
      const int sel = (s < s_end) ? min(a.Din - f0, a.d) : 0;
      const float* crow = a.ctrd + (static_cast<size_t>(s) * K + k) * a.d;
      if (a.d <= 8) {
        float c[8];
        int off[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          c[j] = 0.0f;
          off[j] = 0;
          if (j < sel) {
            const int f = f0 + j;
            off[j] = a.hw ? ((f % a.hw) * a.ch + f / a.hw) : f;  // NHWC source read in NCHW-flatten order
            c[j] = __ldg(crow + j);
          }
        }
#pragma unroll
        for (int nl = 0; nl < TN; nl++) {
          const int n = n0 + nl;
          float v = 0.0f;
          if (n < a.N) {
            const float* x = a.src + static_cast<size_t>(n) * a.Din;
#pragma unroll
            for (int j = 0; j < 8; j++)
              if (j < sel) v = __fadd_rn(v, __fmul_rn(__ldg(x + off[j]), c[j]));
          }
          lut[(nl * PF + r) * K + k] = v;
        }
      } else {
        for (int nl = 0; nl < TN; nl++) {
          const int n = n0 + nl;
          float v = 0.0f;
          if (n < a.N) {
            const float* x = a.src + static_cast<size_t>(n) * a.Din;
            for (int j = 0; j < sel; j++) {
              const int f = f0 + j;
              const int off = a.hw ? ((f % a.hw) * a.ch + f / a.hw) : f;
              v = __fadd_rn(v, __fmul_rn(__ldg(x + off), __ldg(crow + j)));
            }
          }
          lut[(nl * PF + r) * K + k] = v;
        }
      }
    }
    __syncthreads();
    // (3) gather-accumulate.  Rows past s_end hold zeros and index 0, so no tail guard is needed.
    const char* lutb = reinterpret_cast<const char*>(lut);
#pragma unroll
    for (int r = 0; r < PF; r++) {
#pragma unroll
      for (int c = 0; c < CPT; c++) {
        const uint32_t w = Word(areg[r], c >> 2);
        uint32_t off = (w >> (8 * (c & 3))) & 0xFFu;
        if (!PRE) off <<= 2;
#pragma unroll
        for (int nl = 0; nl < TN; nl++) {
          acc[nl][c] += *reinterpret_cast<const float*>(lutb + (nl * PF + r) * K * 4 + off);
        }
      }
    }
  }

  if (!live) return;
#pragma unroll
  for (int nl = 0; nl < TN; nl++) {
    const int n = n0 + nl;
    if (n >= a.N) break;
    if (a.nsplit == 1) {
      float* out = a.dst + static_cast<size_t>(n) * a.Dout;
#pragma unroll
      for (int c = 0; c < CPT; c++) {
        if (o0 + c < a.Dout) out[o0 + c] = a.relu ? fmaxf(acc[nl][c], 0.0f) : acc[nl][c];
      }
    } else {
      float* out = a.partial + (static_cast<size_t>(split) * a.N + n) * a.DoutPad;
#pragma unroll
      for (int c = 0; c < CPT; c++) out[o0 + c] = acc[nl][c];
    }
  }
}

// dst[n][o] = sum over splits (ascending; split 0 already carries the bias)
__global__ void fc_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dst, int N, int Dout,
                                 int DoutPad, int nsplit, int relu) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Dout) return;
  const int n = i / Dout, o = i % Dout;
  float v = 0.0f;
  for (int sp = 0; sp < nsplit; sp++) v += partial[(static_cast<size_t>(sp) * N + n) * DoutPad + o];
  dst[i] = relu ? fmaxf(v, 0.0f) : v;
}

template <int K, int CPT, int TN, int PF, bool PRE>
int Launch(const FcArgs& a, dim3 grid, cudaStream_t st) {
  const size_t smem = sizeof(float) * TN * PF * K;
  auto kern = fc_aprx_kernel<K, CPT, TN, PF, PRE>;
  if (smem > 48 * 1024) QCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kFcThreads, smem, st>>>(a);
  QCNN_CUDA(cudaGetLastError());
  return 0;
}

template <int K, bool PRE>
int LaunchK(const FcArgs& a, int tn, dim3 grid, cudaStream_t st) {
  switch (tn) {
    case 1: return Launch<K, 16, 1, 16, PRE>(a, grid, st);
    case 4: return Launch<K, 8, 4, (K <= 64 ? 16 : 8), PRE>(a, grid, st);
    default: return Launch<K, 8, 8, 8, PRE>(a, grid, st);
  }
}

}  // namespace

namespace qcnn {

// chunk length (subspaces) of the instantiation chosen for (K, tn) -- must mirror LaunchK above
static int ChunkLen(int K, int tn) {
  if (tn == 1) return 16;
  if (tn == 4) return K <= 64 ? 16 : 8;
  return 8;
}

int LaunchFc(qcnn_layer* L, const float* src, int N, float* dst, int relu, cudaStream_t st) {
  QCNN_CHECK(L->kind == QCNN_KIND_FC, "qcnn_fc_aprx_forward: layer is not fully-connected");
  QCNN_CHECK(N >= 1, "qcnn_fc_aprx_forward: N must be >= 1");
  qcnn_ctx* ctx = L->ctx;
  FcArgs a;
  a.src = src; a.dst = dst; a.partial = nullptr;
  a.ctrd = L->d_ctrd; a.asmt = L->d_asmt; a.bias = L->d_bias;
  a.N = N; a.Din = L->Din; a.Dout = L->Dout; a.DoutPad = L->DoutPad; a.S = L->S; a.K = L->K; a.d = L->d;
  a.hw = L->src_h * L->src_w; a.ch = L->src_c;
  a.relu = relu;

  // batch tile: 1 (latency path, 16 channels/thread, 128-bit loads), 4 or 8 images per CTA
  int tn = L->opt_fc_tn ? L->opt_fc_tn : (N >= 8 ? 8 : (N >= 4 ? 4 : 1));
  if (tn != 1 && tn != 4) tn = 8;
  const int cpt = (tn == 1) ? 16 : 8;
  const int pf = ChunkLen(L->K, tn);
  const int gx = CeilDiv(L->DoutPad, kFcThreads * cpt);
  const int gy = CeilDiv(N, tn);
  const int chunks = CeilDiv(L->S, pf);
  int nsplit = L->opt_fc_nsplit;
  if (nsplit <= 0) {
    // aim for >= 3 CTAs of 8 warps per SM; never more splits than chunks
    const int want = 3 * ctx->sm_count;
    nsplit = CeilDiv(want, gx * gy);
  }
  nsplit = std::max(1, std::min(nsplit, chunks));
  const int chunks_per_split = CeilDiv(chunks, nsplit);
  nsplit = CeilDiv(chunks, chunks_per_split);
  a.s_per_split = chunks_per_split * pf;
  a.nsplit = nsplit;
  if (nsplit > 1) {
    const size_t need = sizeof(float) * static_cast<size_t>(nsplit) * N * L->DoutPad;
    if (need > L->partial_bytes) {
      if (L->d_partial) QCNN_CUDA(cudaFree(L->d_partial));
      L->d_partial = nullptr;
      L->partial_bytes = 0;
      QCNN_CUDA(cudaMalloc(&L->d_partial, need));
      L->partial_bytes = need;
    }
    a.partial = L->d_partial;
  }
  dim3 grid(gx, gy, nsplit);
  int rc;
  const bool pre = L->kshift == 2;
  switch (L->K) {
    case 16:  rc = pre ? LaunchK<16, true>(a, tn, grid, st) : LaunchK<16, false>(a, tn, grid, st); break;
    case 32:  rc = pre ? LaunchK<32, true>(a, tn, grid, st) : LaunchK<32, false>(a, tn, grid, st); break;
    case 64:  rc = pre ? LaunchK<64, true>(a, tn, grid, st) : LaunchK<64, false>(a, tn, grid, st); break;
    case 128: rc = LaunchK<128, false>(a, tn, grid, st); break;
    case 256: rc = LaunchK<256, false>(a, tn, grid, st); break;
    default:
      SetError("qcnn_fc_aprx_forward: unsupported codebook size K=%d (supported: 16, 32, 64, 128, 256)", L->K);
      return 1;
  }
  if (rc) return rc;
  ctx->launches++;
  if (nsplit > 1) {
    const int total = N * L->Dout;
    fc_reduce_kernel<<<CeilDiv(total, 256), 256, 0, st>>>(a.partial, dst, N, L->Dout, L->DoutPad, nsplit, relu);
    QCNN_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  return 0;
}

}  // namespace qcnn
