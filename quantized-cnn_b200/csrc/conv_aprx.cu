// Fused PQ convolution for sm_100a: per-subspace LUT (codebook x input inner products) built in shared memory
// and consumed by the uint8 gather-accumulate of the same CTA -- the LUT never touches HBM.
// Replaces CaffeEva::CalcFeatMap_ConvAprx + GetInPdMat (reference src/CaffeEva.cc:760-868, 1261-1296):
//     LUT[n][hi][wi][s][k] = sum_j src[n][hi][wi][g*Cg + s*d + j] * ctrd[s][k][j]
//     dst[n][ho][wo][g*Kg+c] = bias + sum_{kh,kw in bounds} sum_s LUT[n][ho*st-pad+kh][wo*st-pad+kw][s][asmt[kh][kw][s][g*Kg+c]]
//
// Design (hard parts 1/4): the binding resource is shared-memory gather bandwidth
// (32 four-byte lookups / clk / SM), so the gather is laid out to be bank-conflict free and issue-light:
//   * lane = output pixel, LUT stored TRANSPOSED as lut[k][position]; the codeword index of a (tap, s, channel)
//     is warp-uniform, so the 32 lanes read 32 consecutive floats (one wavefront, no conflicts);
//   * zero padding is realised by zero LUT columns (padded flat grid), so no per-lane bounds predicate exists in
//     the inner loop -- out-of-image taps add +0.0f, which is what the reference's tap skipping amounts to;
//   * a thread owns J positions x CPT channels (<= 64 accumulators, held as float2 pairs so two positions are
//     accumulated by one FADD2); per 4 channels it issues one 128-bit broadcast load of 4 pre-multiplied LUT row
//     offsets, then 4 x (1 IADD + J LDS + J/2 FADD2);
//   * the per-(s) assignment slice is staged in shared memory already multiplied by the LUT row pitch.
// The LUT stage is a [K x d] x [d x positions] contraction per subspace.  It runs on the FP32 pipe as a
// register-blocked mini-GEMM: 8 codewords x 4 positions per thread, packed FFMA2 (two FMAs per issue slot), codebook
// rows broadcast from shared memory, 128-bit LUT stores.  (fp32, ascending j: parity with the reference's saxpy
// order up to FMA contraction; tensor-core TF32 would break the fp32 tolerance -- SURVEY.md 7.5.)
// Two kernels share both stages:
//   conv_s1_kernel   stride 1 (conv2..conv5, sweep): a CTA owns (image, row strip, group, channel tile) and the
//                    flat padded grid of the strip; tap (kh,kw) is a constant shift kh*PW+kw of the position.
//   conv_roll_kernel any stride (conv1: 11x11 / 4): a CTA walks the input rows of its strip once; each input
//                    row's LUT is built once and used by the <= ceil(k/stride) output rows that see it, each
//                    owned by a different "row group" of warps; columns are de-interleaved by stride phase so
//                    that a tap again is a constant shift.
#include "qcnn_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace {

constexpr int kMaxThreads = 512;

// ------------------------------------------------------------------------------------------------------------
// Operand staging.  Everything the LUT stage of subspace s+1 needs from global memory (input pixels of the
// subspace's channels, its codebook rows, its assignment slice) is copied ASYNCHRONOUSLY (cp.async, no registers)
// into shared-memory staging buffers right before the gather of subspace s starts, and picked up after it ends,
// so no global-load latency is exposed between the two stages.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void CpAsync4(void* smemDst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smemDst));
  const int sz = valid ? 4 : 0;  // src-size 0: nothing is read, the 4 destination bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void CpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// raw codebook chunk: cbN[k*8 + jj] <- ctrd[s][k][jc + jj]   (zero beyond the subspace's real dims)
__device__ __forceinline__ void FetchCodebook(float* __restrict__ cbN, const float* __restrict__ ctrd, int s, int K,
                                              int d, int jc, int nj, int tid, int T) {
  const float* base = ctrd + static_cast<size_t>(s) * K * d + jc;
  for (int e = tid; e < K * 8; e += T) {
    const int k = e >> 3, jj = e & 7;
    CpAsync4(cbN + e, base + (jj < nj ? k * d + jj : 0), jj < nj);
  }
}
// cb2[jj][k] = (c, c)
__device__ __forceinline__ void CommitCodebook(const float* __restrict__ cbN, float2* __restrict__ cb2, int K, int tid,
                                               int T) {
  for (int e = tid; e < K * 8; e += T) {
    const float c = cbN[e];
    cb2[(e & 7) * K + (e >> 3)] = make_float2(c, c);
  }
}

// raw assignment slice of (group, s, channel tile): idN[tap*CT + c] (bytes), fetched as 4-channel words
__device__ __forceinline__ void FetchOffsets(uint8_t* __restrict__ idN, const uint8_t* __restrict__ ap, int taps,
                                             int CT, int KgPad, int tid, int T) {
  const int wpt = CT >> 2;
  for (int w = tid; w < taps * wpt; w += T) {
    const int tap = w / wpt, c4 = w - tap * wpt;
    CpAsync4(idN + 4 * w, ap + static_cast<size_t>(tap) * KgPad + 4 * c4, true);
  }
}
// idx[tap][c] = asmt * (PP * 4): byte offset of LUT row k
__device__ __forceinline__ void CommitOffsets(const uint8_t* __restrict__ idN, uint32_t* __restrict__ idx, int taps,
                                              int CT, uint32_t rowBytes, int tid, int T) {
  const int wpt = CT >> 2;
  for (int w = tid; w < taps * wpt; w += T) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(idN + 4 * w);
    *reinterpret_cast<uint4*>(idx + 4 * w) = make_uint4((v & 0xFFu) * rowBytes, ((v >> 8) & 0xFFu) * rowBytes,
                                                        ((v >> 16) & 0xFFu) * rowBytes, (v >> 24) * rowBytes);
  }
}

// input pixels: xs[jj][pos] <- channel ch0+jj of the pixel behind LUT position pos; posoff[pos] = element offset
// of the pixel's channel 0 (or -1 for zero padding); chStride = 1 for NHWC, H*W for NCHW
__device__ __forceinline__ void FetchPixels(float* __restrict__ xs, const float* __restrict__ src,
                                            const int* __restrict__ posoff, int PP, int ch0, size_t chStride, int nj,
                                            int tid, int T) {
  for (int pos = tid; pos < PP; pos += T) {
    const int off = posoff[pos];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const bool ok = off >= 0 && jj < nj;
      CpAsync4(xs + jj * PP + pos, src + (ok ? off + (ch0 + jj) * chStride : 0), ok);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// LUT stage: lut[k][p] (+)= sum_{jj<nj} cb[jj][k] * xs[jj][p], tile of 8 codewords x 4 positions per thread.
// Lanes walk the position tiles (128-bit LUT stores and x loads are conflict-free, codebook loads broadcast).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void LutStep(float2 (&acc)[8][2], const float2* __restrict__ cb2, const float* __restrict__ xs,
                                        int K, int PP, int j, int k0, int p0) {
  const float4 xv = *reinterpret_cast<const float4*>(xs + j * PP + p0);
  const float2 x01 = make_float2(xv.x, xv.y), x23 = make_float2(xv.z, xv.w);
  const float4* cp = reinterpret_cast<const float4*>(cb2 + j * K + k0);
#pragma unroll
  for (int kk = 0; kk < 4; kk++) {
    const float4 cc = cp[kk];  // (c[2kk], c[2kk], c[2kk+1], c[2kk+1])
    const float2 ca = make_float2(cc.x, cc.y), cb = make_float2(cc.z, cc.w);
    acc[2 * kk][0] = __ffma2_rn(ca, x01, acc[2 * kk][0]);
    acc[2 * kk][1] = __ffma2_rn(ca, x23, acc[2 * kk][1]);
    acc[2 * kk + 1][0] = __ffma2_rn(cb, x01, acc[2 * kk + 1][0]);
    acc[2 * kk + 1][1] = __ffma2_rn(cb, x23, acc[2 * kk + 1][1]);
  }
}

__device__ __forceinline__ void BuildLut(float* __restrict__ lut, const float2* __restrict__ cb2,
                                         const float* __restrict__ xs, int K, int PP, int nj, bool first, int tid,
                                         int T) {
  const int npt = PP >> 2;
  const int ntiles = npt * (K >> 3);
  for (int t = tid; t < ntiles; t += T) {
    const int kt = t / npt;
    const int p0 = (t - kt * npt) << 2;
    const int k0 = kt << 3;
    float2 acc[8][2];
    if (first) {
#pragma unroll
      for (int k = 0; k < 8; k++) { acc[k][0] = make_float2(0.f, 0.f); acc[k][1] = make_float2(0.f, 0.f); }
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const float4 v = *reinterpret_cast<const float4*>(lut + (k0 + k) * PP + p0);
        acc[k][0] = make_float2(v.x, v.y);
        acc[k][1] = make_float2(v.z, v.w);
      }
    }
    if (nj == 8) {  // two steps in flight: enough ILP to cover the LDS latency without blowing the register file
#pragma unroll 2
      for (int j = 0; j < 8; j++) LutStep(acc, cb2, xs, K, PP, j, k0, p0);
    } else {
#pragma unroll 1
      for (int j = 0; j < nj; j++) LutStep(acc, cb2, xs, K, PP, j, k0, p0);
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
      *reinterpret_cast<float4*>(lut + (k0 + k) * PP + p0) = make_float4(acc[k][0].x, acc[k][0].y, acc[k][1].x, acc[k][1].y);
  }
}

// ------------------------------------------------------------------------------------------------------------
// gather stage
// ------------------------------------------------------------------------------------------------------------
// Accumulators of one thread: J positions x CPT channels.  For even J two positions share a float2 so that one
// FADD2 retires two lookups.
template <int CPT, int J>
struct Acc {
  static constexpr int JP = (J % 2 == 0) ? J / 2 : J;
  static constexpr bool kPaired = (J % 2 == 0);
  float2 v[JP][CPT];  // paired: (pos 2jp, pos 2jp+1); unpaired: .x only

  __device__ __forceinline__ void Fill(const float (&b)[CPT]) {
#pragma unroll
    for (int jp = 0; jp < JP; jp++)
#pragma unroll
      for (int c = 0; c < CPT; c++) v[jp][c] = make_float2(b[c], b[c]);
  }
  __device__ __forceinline__ float Get(int j, int c) const {
    if (kPaired) return (j & 1) ? v[j >> 1][c].y : v[j >> 1][c].x;
    return v[j][c].x;
  }
};

// one tap: acc[j][c] += lut[idx[c]][q_j + shift]; `base` = byte address of lut[0][q_0 + shift]; positions of a
// thread are 32 apart (j * 128 bytes), so the J loads of a channel share one address register.
template <int CPT, int J>
__device__ __forceinline__ void GatherTap(Acc<CPT, J>& acc, const char* base, const uint32_t* __restrict__ ip) {
#pragma unroll
  for (int c4 = 0; c4 < CPT; c4 += 4) {
    const uint4 o = *reinterpret_cast<const uint4*>(ip + c4);
    const char* b[4] = {base + o.x, base + o.y, base + o.z, base + o.w};
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (Acc<CPT, J>::kPaired) {
#pragma unroll
        for (int jp = 0; jp < J / 2; jp++) {
          const float2 val = make_float2(*reinterpret_cast<const float*>(b[u] + (2 * jp) * 128),
                                         *reinterpret_cast<const float*>(b[u] + (2 * jp + 1) * 128));
          acc.v[jp][c4 + u] = __fadd2_rn(acc.v[jp][c4 + u], val);
        }
      } else {
#pragma unroll
        for (int j = 0; j < J; j++) acc.v[j][c4 + u].x += *reinterpret_cast<const float*>(b[u] + j * 128);
      }
    }
  }
}

template <int CPT, int J>
__device__ __forceinline__ void StoreChannels(float* __restrict__ out, const Acc<CPT, J>& acc, int j, int relu) {
#pragma unroll
  for (int c = 0; c < CPT; c += 4) {
    float4 o = make_float4(acc.Get(j, c), acc.Get(j, c + 1), acc.Get(j, c + 2), acc.Get(j, c + 3));
    if (relu) {
      o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f);
    }
    *reinterpret_cast<float4*>(out + c) = o;
  }
}

struct SmemLayout {
  float* lut;      // [K][PP]
  uint32_t* idx;   // [taps][CT]   LUT row byte offsets of the current subspace
  float2* cb2;     // [8][K]       duplicated codebook chunk of the current subspace
  float* xs[2];    // [8][PP]      input pixels, double-buffered (current / next subspace)
  int* posoff;     // [PP]
  float* cbN;      // [K*8]        raw codebook chunk of the next subspace (cp.async target)
  uint8_t* idN;    // [taps*CT]    raw assignment slice of the next subspace (cp.async target)
};

__device__ __forceinline__ SmemLayout Carve(unsigned char* smem, int K, int PP, int taps, int CT) {
  SmemLayout s;
  s.lut = reinterpret_cast<float*>(smem);
  s.idx = reinterpret_cast<uint32_t*>(s.lut + static_cast<size_t>(K) * PP);
  s.cb2 = reinterpret_cast<float2*>(s.idx + taps * CT);
  s.xs[0] = reinterpret_cast<float*>(s.cb2 + 8 * K);
  s.xs[1] = s.xs[0] + 8 * PP;
  s.posoff = reinterpret_cast<int*>(s.xs[1] + 8 * PP);
  s.cbN = reinterpret_cast<float*>(s.posoff + PP);
  s.idN = reinterpret_cast<uint8_t*>(s.cbN + 8 * K);
  return s;
}

// stage one (s, jc) chunk with exposed latency: only for the rare jc > 0 chunks of subspaces wider than 8 dims
__device__ __forceinline__ void StageChunkDirect(const SmemLayout& sm, float* xs, const ConvArgs& a, const float* src,
                                                 int s, int jc, int nj, int ch0, size_t chStride, int tid, int T) {
  FetchCodebook(sm.cbN, a.ctrd, s, a.K, a.d, jc, nj, tid, T);
  FetchPixels(xs, src, sm.posoff, a.PP, ch0, chStride, nj, tid, T);
  CpAsyncCommit();
  CpAsyncWaitAll();
  __syncthreads();
  CommitCodebook(sm.cbN, sm.cb2, a.K, tid, T);
}

// ------------------------------------------------------------------------------------------------------------
// stride-1 kernel
// ------------------------------------------------------------------------------------------------------------
// MAXT: launch bound (256 / 384 / 512 threads) -> register cap 255 / 168 / 128, so small CTAs never spill
template <int CPT, int J, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) conv_s1_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int taps = a.ksz * a.ksz;
  const SmemLayout sm = Carve(smem, a.K, a.PP, taps, a.CT);

  const int tid = threadIdx.x, T = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int pw = warp % a.pwarps, cw = warp / a.pwarps;
  int b = blockIdx.x;
  const int strip = b % a.nstrips; b /= a.nstrips;
  const int ct = b % a.nct;
  const int g = b / a.nct;
  const int n = blockIdx.y;
  const int r0 = strip * a.R;            // first output row of the strip
  const int hin0 = r0 - a.pad;           // input row held at ri = 0
  const int q0 = pw * 32 * J + lane;     // first flat position of this thread (others at +32 j)
  const int cbase = ct * a.CT + cw * CPT;  // first channel (within the group) of this thread
  const float* src = a.src + static_cast<size_t>(n) * a.Hi * a.Wi * a.Cin + g * a.Cg;
  const uint8_t* asmtG = a.asmt + static_cast<size_t>(g) * a.S * taps * a.KgPad + ct * a.CT;

  // LUT position -> source pixel (NHWC element offset within this image/group), -1 = zero padding
  for (int pos = tid; pos < a.PP; pos += T) {
    int off = -1;
    const int pp = pos - a.pad;
    if (pp >= 0) {
      const int ri = pp / a.PW, wi = pp - ri * a.PW;
      const int hi = hin0 + ri;
      if (ri < a.RI && wi < a.Wi && hi >= 0 && hi < a.Hi) off = (hi * a.Wi + wi) * a.Cin;
    }
    sm.posoff[pos] = off;
  }

  Acc<CPT, J> acc;
  {
    float bv[CPT];
#pragma unroll
    for (int c = 0; c < CPT; c++) bv[c] = __ldg(a.bias + g * a.Kg + cbase + c);
    acc.Fill(bv);
  }
  __syncthreads();  // posoff visible

  const uint32_t rowBytes = static_cast<uint32_t>(a.PP) * 4u;
  {
    const int nj0 = max(0, min(8, min(a.Cg, a.d)));
    FetchCodebook(sm.cbN, a.ctrd, 0, a.K, a.d, 0, nj0, tid, T);
    FetchOffsets(sm.idN, asmtG, taps, a.CT, a.KgPad, tid, T);
    FetchPixels(sm.xs[0], src, sm.posoff, a.PP, 0, 1, nj0, tid, T);
    CpAsyncCommit();
  }
  for (int s = 0; s < a.S; s++) {
    const int dsel = min(a.Cg - s * a.d, a.d);  // dims of this subspace that exist (reference CaffeEva.cc:1277)
    float* xs = sm.xs[s & 1];
    // ---- LUT stage ----
    CpAsyncWaitAll();
    __syncthreads();  // (A) staged operands of s have landed; the previous gather is done with lut / idx
    CommitCodebook(sm.cbN, sm.cb2, a.K, tid, T);
    CommitOffsets(sm.idN, sm.idx, taps, a.CT, rowBytes, tid, T);
    __syncthreads();  // (B)
    BuildLut(sm.lut, sm.cb2, xs, a.K, a.PP, max(0, min(8, dsel)), true, tid, T);
    for (int jc = 8; jc < dsel; jc += 8) {
      const int nj = min(8, dsel - jc);
      __syncthreads();
      StageChunkDirect(sm, xs, a, src, s, jc, nj, s * a.d + jc, 1, tid, T);
      __syncthreads();
      BuildLut(sm.lut, sm.cb2, xs, a.K, a.PP, nj, false, tid, T);
    }
    __syncthreads();  // (C) lut complete; cbN / idN / the other xs buffer are free
    if (s + 1 < a.S) {  // operands of the next subspace fly in while this one is gathered
      const int njn = max(0, min(8, min(a.Cg - (s + 1) * a.d, a.d)));
      FetchCodebook(sm.cbN, a.ctrd, s + 1, a.K, a.d, 0, njn, tid, T);
      FetchOffsets(sm.idN, asmtG + static_cast<size_t>(s + 1) * taps * a.KgPad, taps, a.CT, a.KgPad, tid, T);
      FetchPixels(sm.xs[(s + 1) & 1], src, sm.posoff, a.PP, (s + 1) * a.d, 1, njn, tid, T);
      CpAsyncCommit();
    }
    // ---- gather stage ----
    const char* lutq = reinterpret_cast<const char*>(sm.lut) + q0 * 4;
    const uint32_t* ip = sm.idx + cw * CPT;
    for (int kh = 0; kh < a.ksz; kh++) {
      for (int kw = 0; kw < a.ksz; kw++) {
        GatherTap<CPT, J>(acc, lutq + (kh * a.PW + kw) * 4, ip);
        ip += a.CT;
      }
    }
  }

#pragma unroll
  for (int j = 0; j < J; j++) {
    const int q = q0 + 32 * j;
    const int r = q / a.PW, wo = q - r * a.PW;
    const int ho = r0 + r;
    if (r < a.R && wo < a.Wo && ho < a.Ho) {
      float* out = a.dst + ((static_cast<size_t>(n) * a.Ho + ho) * a.Wo + wo) * a.Cout + g * a.Kg + cbase;
      StoreChannels<CPT, J>(out, acc, j, a.relu);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// stride-1 kernel, tensor-core LUT stage (K = 128 codewords, subspaces of <= 8 dims)
//
// The LUT of a subspace is the dense contraction  LUT[k][p] = sum_j ctrd[s][k][j] * x[p][j]  = C (128 x 8) * X^T (8 x N).
// Here it runs on the 5th-generation tensor cores: ONE thread issues three tcgen05.mma kind::tf32 (M = 128, N <= 256,
// K = 8) per subspace -- the 3xTF32 split  C_hi*X_hi + C_hi*X_lo + C_lo*X_hi  keeps fp32-level accuracy (measured
// 7e-7 of the LUT magnitude, tools/tc_lut_test.cu) -- with the accumulator in TMEM (lane = codeword, column =
// position).  The MMA of subspace s+1 executes asynchronously WHILE the CUDA cores gather subspace s out of shared
// memory; at the stage boundary the finished accumulator is drained TMEM -> registers -> lut[k][p] (tcgen05.ld +
// 128-bit stores, bank-conflict free because the row pitch is 4 mod 32 words).  Operands reach the canonical K-major
// no-swizzle UMMA tiles through cp.async (raw) + a hi/lo split pass.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t SmemU32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// canonical K-major / SWIZZLE_NONE operand tile (tf32): rows in groups of 8, per group K-chunk 0 (8 rows x 16 B) then
// K-chunk 1; element (row r, k) -> float index below.  Descriptor: LBO = 128 B (K chunks), SBO = 256 B (row groups).
__device__ __forceinline__ int UmmaIdx(int r, int k) { return (r >> 3) * 64 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }
__device__ __forceinline__ uint64_t UmmaDesc(const void* smem) {
  uint64_t d = (static_cast<uint64_t>(SmemU32(smem)) >> 4) & 0x3FFF;
  d |= static_cast<uint64_t>(8) << 16;    // leading byte offset 128 B >> 4
  d |= static_cast<uint64_t>(16) << 32;   // stride byte offset 256 B >> 4
  d |= static_cast<uint64_t>(1) << 46;    // descriptor version: Blackwell
  return d;
}
__device__ __forceinline__ void UmmaTf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemD), "l"(descA), "l"(descB),
               "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void MbarWait(uint64_t* mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(SmemU32(mbar)), "r"(parity) : "memory");
  }
}
// hi = value truncated to tf32 (what the tensor core reads), lo = exact remainder
__device__ __forceinline__ void SplitTf32(float v, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = v - hi;
}

struct SmemLayoutTc {
  float* lut;       // [128][PP]
  uint32_t* idx;    // [taps][CT]
  float* aHi;       // [128 x 8] canonical
  float* aLo;
  float* bHi;       // [NM x 8] canonical
  float* bLo;
  float* xr;        // [NM][8] raw pixels of the next subspace (cp.async target)
  float* cr;        // [128][8] raw codebook of the next subspace
  uint8_t* idN;     // [taps*CT] raw assignment slice
  int* posoff;      // [NM]
  uint64_t* mbar;
  uint32_t* tmemBase;
};
__device__ __forceinline__ SmemLayoutTc CarveTc(unsigned char* smem, int PP, int NM, int taps, int CT) {
  SmemLayoutTc s;
  s.lut = reinterpret_cast<float*>(smem);
  s.idx = reinterpret_cast<uint32_t*>(s.lut + 128 * PP);
  s.aHi = reinterpret_cast<float*>(s.idx + taps * CT);
  s.aLo = s.aHi + 1024;
  s.bHi = s.aLo + 1024;
  s.bLo = s.bHi + NM * 8;
  s.xr = s.bLo + NM * 8;
  s.cr = s.xr + NM * 8;
  s.posoff = reinterpret_cast<int*>(s.cr + 1024);
  s.mbar = reinterpret_cast<uint64_t*>(s.posoff + NM);
  s.tmemBase = reinterpret_cast<uint32_t*>(s.mbar + 1);
  s.idN = reinterpret_cast<uint8_t*>(s.tmemBase + 2);
  return s;
}

template <int CPT, int J, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) conv_s1_tc_kernel(const ConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int taps = a.ksz * a.ksz;
  const int NM = a.RI;  // MMA N: LUT columns produced per subspace (multiple of 32, <= 256); a.RI is reused for it
  const SmemLayoutTc sm = CarveTc(smem, a.PP, NM, taps, a.CT);

  const int tid = threadIdx.x, T = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
  const int pw = warp % a.pwarps, cw = warp / a.pwarps;
  int b = blockIdx.x;
  const int strip = b % a.nstrips; b /= a.nstrips;
  const int ct = b % a.nct;
  const int g = b / a.nct;
  const int n = blockIdx.y;
  const int r0 = strip * a.R;
  const int hin0 = r0 - a.pad;
  const int rowsIn = a.R + a.ksz - 1;
  const int q0 = pw * 32 * J + lane;
  const int cbase = ct * a.CT + cw * CPT;
  const float* src = a.src + static_cast<size_t>(n) * a.Hi * a.Wi * a.Cin + g * a.Cg;
  const uint8_t* asmtG = a.asmt + static_cast<size_t>(g) * a.S * taps * a.KgPad + ct * a.CT;

  for (int pos = tid; pos < NM; pos += T) {
    int off = -1;
    const int pp = pos - a.pad;
    if (pp >= 0) {
      const int ri = pp / a.PW, wi = pp - ri * a.PW;
      const int hi = hin0 + ri;
      if (ri < rowsIn && wi < a.Wi && hi >= 0 && hi < a.Hi) off = (hi * a.Wi + wi) * a.Cin;
    }
    sm.posoff[pos] = off;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemU32(sm.tmemBase)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemU32(sm.mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  Acc<CPT, J> acc;
  {
    float bv[CPT];
#pragma unroll
    for (int c = 0; c < CPT; c++) bv[c] = __ldg(a.bias + g * a.Kg + cbase + c);
    acc.Fill(bv);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();  // posoff, mbarrier, TMEM base visible
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmemD = *sm.tmemBase;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(NM >> 3) << 17) | (8u << 24);
  const uint32_t rowBytes = static_cast<uint32_t>(a.PP) * 4u;

  // raw operands of subspace s: pixels [NM][8] and codebook [128][8], zero beyond the subspace's real dims
  auto fetchRaw = [&](int s) {
    const int nj = max(0, min(8, min(a.Cg - s * a.d, a.d)));
    const float* cg = a.ctrd + static_cast<size_t>(s) * 128 * a.d;
    for (int e = tid; e < 1024; e += T) {
      const int k = e >> 3, jj = e & 7;
      CpAsync4(sm.cr + e, cg + (jj < nj ? k * a.d + jj : 0), jj < nj);
    }
    const int ch0 = s * a.d;
    for (int e = tid; e < NM * 8; e += T) {
      const int pos = e >> 3, jj = e & 7;
      const int off = sm.posoff[pos];
      const bool ok = off >= 0 && jj < nj;
      CpAsync4(sm.xr + e, src + (ok ? off + ch0 + jj : 0), ok);
    }
  };
  // raw -> hi/lo canonical UMMA tiles (generic-proxy writes, published to the async proxy by the caller's fence)
  auto splitOperands = [&]() {
    for (int e = tid; e < 1024; e += T) {
      float hi, lo;
      SplitTf32(sm.cr[e], hi, lo);
      const int i = UmmaIdx(e >> 3, e & 7);
      sm.aHi[i] = hi;
      sm.aLo[i] = lo;
    }
    for (int e = tid; e < NM * 8; e += T) {
      float hi, lo;
      SplitTf32(sm.xr[e], hi, lo);
      const int i = UmmaIdx(e >> 3, e & 7);
      sm.bHi[i] = hi;
      sm.bLo[i] = lo;
    }
  };
  auto issueMma = [&]() {  // one thread: D = Ah*Bh + Ah*Bl + Al*Bh, then signal the mbarrier when all three retire
    const uint64_t dAh = UmmaDesc(sm.aHi), dAl = UmmaDesc(sm.aLo), dBh = UmmaDesc(sm.bHi), dBl = UmmaDesc(sm.bLo);
    UmmaTf32(tmemD, dAh, dBh, idesc, 0);
    UmmaTf32(tmemD, dAh, dBl, idesc, 1);
    UmmaTf32(tmemD, dAl, dBh, idesc, 1);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemU32(sm.mbar)) : "memory");
  };

  // prologue: LUT(0) on the tensor core, operands of subspace 1 + assignment slice 0 in flight
  fetchRaw(0);
  CpAsyncCommit();
  CpAsyncWaitAll();
  __syncthreads();
  splitOperands();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) issueMma();
  if (a.S > 1) fetchRaw(1);
  FetchOffsets(sm.idN, asmtG, taps, a.CT, a.KgPad, tid, T);
  CpAsyncCommit();

  for (int s = 0; s < a.S; s++) {
    MbarWait(sm.mbar, s & 1);   // LUT(s) is complete in TMEM; the operand tiles are free again
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    CpAsyncWaitAll();
    __syncthreads();            // (A) gather(s-1) is done with lut/idx; raw operands of s+1 and idN(s) are visible
    // drain TMEM -> lut: warp w owns TMEM lanes [32 (w%4), +32) = codewords; 32 columns (positions) per load
    {
      const int q = warp & 3;
      const int k = q * 32 + lane;
      float* row = sm.lut + k * a.PP;
      for (int c = warp >> 2; c < (NM >> 5); c += (nwarps + 3 - q) >> 2) {
        uint32_t r[32];
        const uint32_t taddr = tmemD + (static_cast<uint32_t>(q * 32) << 16) + c * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                       "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
                       "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
                       "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(row + c * 32 + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
      }
    }
    CommitOffsets(sm.idN, sm.idx, taps, a.CT, rowBytes, tid, T);
    if (s + 1 < a.S) splitOperands();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();            // (B) lut + idx ready, TMEM drained, operand tiles of s+1 published
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (s + 1 < a.S) {
      if (tid == 0) issueMma();                  // LUT(s+1) is computed while subspace s is gathered
      if (s + 2 < a.S) fetchRaw(s + 2);
      FetchOffsets(sm.idN, asmtG + static_cast<size_t>(s + 1) * taps * a.KgPad, taps, a.CT, a.KgPad, tid, T);
      CpAsyncCommit();
    }
    // ---- gather stage ----
    const char* lutq = reinterpret_cast<const char*>(sm.lut) + q0 * 4;
    const uint32_t* ip = sm.idx + cw * CPT;
    for (int kh = 0; kh < a.ksz; kh++) {
      for (int kw = 0; kw < a.ksz; kw++) {
        GatherTap<CPT, J>(acc, lutq + (kh * a.PW + kw) * 4, ip);
        ip += a.CT;
      }
    }
  }

#pragma unroll
  for (int j = 0; j < J; j++) {
    const int q = q0 + 32 * j;
    const int r = q / a.PW, wo = q - r * a.PW;
    const int ho = r0 + r;
    if (r < a.R && wo < a.Wo && ho < a.Ho) {
      float* out = a.dst + ((static_cast<size_t>(n) * a.Ho + ho) * a.Wo + wo) * a.Cout + g * a.Kg + cbase;
      StoreChannels<CPT, J>(out, acc, j, a.relu);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemD), "r"(256) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// rolling-row kernel (any stride)
// ------------------------------------------------------------------------------------------------------------
template <int CPT, int J, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) conv_roll_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int taps = a.ksz * a.ksz;
  const SmemLayout sm = Carve(smem, a.K, a.PP, taps, a.CT);  // PP = stride * PH (rounded)

  const int tid = threadIdx.x, T = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int rg = warp % a.rgroups;
  const int rest = warp / a.rgroups;
  const int pw = rest % a.pwarps, cw = rest / a.pwarps;
  int b = blockIdx.x;
  const int strip = b % a.nstrips; b /= a.nstrips;
  const int ct = b % a.nct;
  const int g = b / a.nct;
  const int n = blockIdx.y;
  const int ho0 = strip * a.R;
  const int ho_end = min(a.Ho, ho0 + a.R);
  const int PH = a.PW;
  const int wo0 = pw * 32 * J + lane;
  const int cbase = ct * a.CT + cw * CPT;
  // source addressing: pixel (hi, wi), channel ch  ->  base + hi*rowStride + wi*pixStride + ch*chStride
  const size_t chStride = a.src_nchw ? static_cast<size_t>(a.Hi) * a.Wi : 1;
  const int pixStride = a.src_nchw ? 1 : a.Cin;
  const size_t rowStride = static_cast<size_t>(a.Wi) * pixStride;
  const float* src = a.src + static_cast<size_t>(n) * a.Hi * a.Wi * a.Cin + static_cast<size_t>(g) * a.Cg * chStride;
  const uint8_t* asmtG = a.asmt + static_cast<size_t>(g) * a.S * taps * a.KgPad + ct * a.CT;

  // LUT position (phase-major de-interleaved column) -> element offset of the pixel inside its input row
  for (int pos = tid; pos < a.PP; pos += T) {
    const int phase = pos / PH, i = pos - phase * PH;
    const int wi = i * a.stride + phase - a.pad;
    sm.posoff[pos] = (phase < a.stride && wi >= 0 && wi < a.Wi) ? wi * pixStride : -1;
  }

  float bias[CPT];
#pragma unroll
  for (int c = 0; c < CPT; c++) bias[c] = __ldg(a.bias + g * a.Kg + cbase + c);
  Acc<CPT, J> acc;
  acc.Fill(bias);
  int ho_cur = ho0 + rg;  // output row this warp is accumulating
  __syncthreads();        // posoff visible

  const uint32_t rowBytes = static_cast<uint32_t>(a.PP) * 4u;
  const int hi_begin = ho0 * a.stride - a.pad;
  const int hi_end = (ho_end - 1) * a.stride - a.pad + a.ksz;
  const int hi_lo = max(hi_begin, 0), hi_hi = min(hi_end, a.Hi);  // rows outside the image contribute nothing
  const int steps = (hi_hi - hi_lo) * a.S;                         // (input row, subspace) pairs, row-major

  if (steps > 0) {
    const int nj0 = max(0, min(8, min(a.Cg, a.d)));
    FetchCodebook(sm.cbN, a.ctrd, 0, a.K, a.d, 0, nj0, tid, T);
    FetchOffsets(sm.idN, asmtG, taps, a.CT, a.KgPad, tid, T);
    FetchPixels(sm.xs[0], src + hi_lo * rowStride, sm.posoff, a.PP, 0, chStride, nj0, tid, T);
    CpAsyncCommit();
  }
  int hi = hi_lo, s = 0;
  for (int it = 0; it < steps; it++) {
    const int dsel = min(a.Cg - s * a.d, a.d);
    const int kh = hi + a.pad - ho_cur * a.stride;  // warp-uniform
    const bool active = ho_cur < ho_end && kh >= 0 && kh < a.ksz;
    float* xs = sm.xs[it & 1];
    CpAsyncWaitAll();
    __syncthreads();  // (A)
    CommitCodebook(sm.cbN, sm.cb2, a.K, tid, T);
    if (a.S > 1 || it == 0) CommitOffsets(sm.idN, sm.idx, taps, a.CT, rowBytes, tid, T);
    __syncthreads();  // (B)
    BuildLut(sm.lut, sm.cb2, xs, a.K, a.PP, max(0, min(8, dsel)), true, tid, T);
    for (int jc = 8; jc < dsel; jc += 8) {
      const int nj = min(8, dsel - jc);
      __syncthreads();
      StageChunkDirect(sm, xs, a, src + hi * rowStride, s, jc, nj, s * a.d + jc, chStride, tid, T);
      __syncthreads();
      BuildLut(sm.lut, sm.cb2, xs, a.K, a.PP, nj, false, tid, T);
    }
    __syncthreads();  // (C)
    // next (row, subspace)
    int hin = hi, sn = s + 1;
    if (sn == a.S) { sn = 0; hin = hi + 1; }
    if (it + 1 < steps) {
      const int njn = max(0, min(8, min(a.Cg - sn * a.d, a.d)));
      FetchCodebook(sm.cbN, a.ctrd, sn, a.K, a.d, 0, njn, tid, T);
      if (a.S > 1) FetchOffsets(sm.idN, asmtG + static_cast<size_t>(sn) * taps * a.KgPad, taps, a.CT, a.KgPad, tid, T);
      FetchPixels(sm.xs[(it + 1) & 1], src + hin * rowStride, sm.posoff, a.PP, sn * a.d, chStride, njn, tid, T);
      CpAsyncCommit();
    }
    if (active) {
      const char* lutq = reinterpret_cast<const char*>(sm.lut) + wo0 * 4;
      const uint32_t* ip = sm.idx + (kh * a.ksz) * a.CT + cw * CPT;
      for (int kw = 0; kw < a.ksz; kw++) {
        const int phase = kw % a.stride, sh = kw / a.stride;
        GatherTap<CPT, J>(acc, lutq + (phase * PH + sh) * 4, ip);
        ip += a.CT;
      }
    }
    // this input row was the last one the current output row needs: emit it and move on
    if (s == a.S - 1 && ho_cur < ho_end && kh == a.ksz - 1) {
#pragma unroll
      for (int j = 0; j < J; j++) {
        const int wo = wo0 + 32 * j;
        if (wo < a.Wo) {
          float* out = a.dst + ((static_cast<size_t>(n) * a.Ho + ho_cur) * a.Wo + wo) * a.Cout + g * a.Kg + cbase;
          StoreChannels<CPT, J>(out, acc, j, a.relu);
        }
      }
      acc.Fill(bias);
      ho_cur += a.rgroups;
    }
    hi = hin; s = sn;
  }
  // output rows whose window ends below the image (padding rows were skipped above)
  for (int h = max(hi_hi, hi_begin); h < hi_end; h++) {
    const int kh = h + a.pad - ho_cur * a.stride;
    if (ho_cur < ho_end && kh == a.ksz - 1) {
#pragma unroll
      for (int j = 0; j < J; j++) {
        const int wo = wo0 + 32 * j;
        if (wo < a.Wo) {
          float* out = a.dst + ((static_cast<size_t>(n) * a.Ho + ho_cur) * a.Wo + wo) * a.Cout + g * a.Kg + cbase;
          StoreChannels<CPT, J>(out, acc, j, a.relu);
        }
      }
      acc.Fill(bias);
      ho_cur += a.rgroups;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// rolling-row kernel, tensor-core LUT stage: same row-group structure as conv_roll_kernel; the LUT of the NEXT
// (input row, subspace) step is produced by tcgen05.mma into TMEM while the CUDA cores gather the current one.
// ------------------------------------------------------------------------------------------------------------
template <int CPT, int J, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) conv_roll_tc_kernel(const ConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int taps = a.ksz * a.ksz;
  const int NM = a.RI;  // MMA N (multiple of 32, <= 256)
  const SmemLayoutTc sm = CarveTc(smem, a.PP, NM, taps, a.CT);

  const int tid = threadIdx.x, T = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = T >> 5;
  const int rg = warp % a.rgroups;
  const int rest = warp / a.rgroups;
  const int pw = rest % a.pwarps, cw = rest / a.pwarps;
  int b = blockIdx.x;
  const int strip = b % a.nstrips; b /= a.nstrips;
  const int ct = b % a.nct;
  const int g = b / a.nct;
  const int n = blockIdx.y;
  const int ho0 = strip * a.R;
  const int ho_end = min(a.Ho, ho0 + a.R);
  const int PH = a.PW;
  const int wo0 = pw * 32 * J + lane;
  const int cbase = ct * a.CT + cw * CPT;
  const size_t chStride = a.src_nchw ? static_cast<size_t>(a.Hi) * a.Wi : 1;
  const int pixStride = a.src_nchw ? 1 : a.Cin;
  const size_t rowStride = static_cast<size_t>(a.Wi) * pixStride;
  const float* src = a.src + static_cast<size_t>(n) * a.Hi * a.Wi * a.Cin + static_cast<size_t>(g) * a.Cg * chStride;
  const uint8_t* asmtG = a.asmt + static_cast<size_t>(g) * a.S * taps * a.KgPad + ct * a.CT;

  for (int pos = tid; pos < NM; pos += T) {
    const int phase = pos / PH, i = pos - phase * PH;
    const int wi = i * a.stride + phase - a.pad;
    sm.posoff[pos] = (phase < a.stride && wi >= 0 && wi < a.Wi) ? wi * pixStride : -1;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemU32(sm.tmemBase)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemU32(sm.mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float bias[CPT];
#pragma unroll
  for (int c = 0; c < CPT; c++) bias[c] = __ldg(a.bias + g * a.Kg + cbase + c);
  Acc<CPT, J> acc;
  acc.Fill(bias);
  int ho_cur = ho0 + rg;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmemD = *sm.tmemBase;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(NM >> 3) << 17) | (8u << 24);
  const uint32_t rowBytes = static_cast<uint32_t>(a.PP) * 4u;

  const int hi_begin = ho0 * a.stride - a.pad;
  const int hi_end = (ho_end - 1) * a.stride - a.pad + a.ksz;
  const int hi_lo = max(hi_begin, 0), hi_hi = min(hi_end, a.Hi);
  const int steps = (hi_hi - hi_lo) * a.S;

  auto fetchRaw = [&](int step) {  // step = (hi - hi_lo) * S + s
    const int hi = hi_lo + step / a.S, s = step % a.S;
    const int nj = max(0, min(8, min(a.Cg - s * a.d, a.d)));
    const float* cg = a.ctrd + static_cast<size_t>(s) * 128 * a.d;
    for (int e = tid; e < 1024; e += T) {
      const int k = e >> 3, jj = e & 7;
      CpAsync4(sm.cr + e, cg + (jj < nj ? k * a.d + jj : 0), jj < nj);
    }
    const float* rowp = src + hi * rowStride;
    const int ch0 = s * a.d;
    for (int e = tid; e < NM * 8; e += T) {
      const int pos = e >> 3, jj = e & 7;
      const int off = sm.posoff[pos];
      const bool ok = off >= 0 && jj < nj;
      CpAsync4(sm.xr + e, rowp + (ok ? off + (ch0 + jj) * chStride : 0), ok);
    }
  };
  auto splitOperands = [&]() {
    for (int e = tid; e < 1024; e += T) {
      float hi, lo;
      SplitTf32(sm.cr[e], hi, lo);
      const int i = UmmaIdx(e >> 3, e & 7);
      sm.aHi[i] = hi;
      sm.aLo[i] = lo;
    }
    for (int e = tid; e < NM * 8; e += T) {
      float hi, lo;
      SplitTf32(sm.xr[e], hi, lo);
      const int i = UmmaIdx(e >> 3, e & 7);
      sm.bHi[i] = hi;
      sm.bLo[i] = lo;
    }
  };
  auto issueMma = [&]() {
    const uint64_t dAh = UmmaDesc(sm.aHi), dAl = UmmaDesc(sm.aLo), dBh = UmmaDesc(sm.bHi), dBl = UmmaDesc(sm.bLo);
    UmmaTf32(tmemD, dAh, dBh, idesc, 0);
    UmmaTf32(tmemD, dAh, dBl, idesc, 1);
    UmmaTf32(tmemD, dAl, dBh, idesc, 1);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemU32(sm.mbar)) : "memory");
  };

  if (steps > 0) {
    fetchRaw(0);
    CpAsyncCommit();
    CpAsyncWaitAll();
    __syncthreads();
    splitOperands();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) issueMma();
    if (steps > 1) fetchRaw(1);
    FetchOffsets(sm.idN, asmtG, taps, a.CT, a.KgPad, tid, T);
    CpAsyncCommit();
  }
  int hi = hi_lo, s = 0;
  for (int it = 0; it < steps; it++) {
    const int kh = hi + a.pad - ho_cur * a.stride;  // warp-uniform
    const bool active = ho_cur < ho_end && kh >= 0 && kh < a.ksz;
    MbarWait(sm.mbar, it & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    CpAsyncWaitAll();
    __syncthreads();  // (A)
    {
      const int q = warp & 3;
      const int k = q * 32 + lane;
      float* row = sm.lut + k * a.PP;
      for (int c = warp >> 2; c < (NM >> 5); c += (nwarps + 3 - q) >> 2) {
        uint32_t r[32];
        const uint32_t taddr = tmemD + (static_cast<uint32_t>(q * 32) << 16) + c * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                       "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
                       "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
                       "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(row + c * 32 + i) = make_uint4(r[i], r[i + 1], r[i + 2], r[i + 3]);
      }
    }
    if (a.S > 1 || it == 0) CommitOffsets(sm.idN, sm.idx, taps, a.CT, rowBytes, tid, T);
    if (it + 1 < steps) splitOperands();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // (B)
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int hin = hi, sn = s + 1;
    if (sn == a.S) { sn = 0; hin = hi + 1; }
    if (it + 1 < steps) {
      if (tid == 0) issueMma();
      if (it + 2 < steps) fetchRaw(it + 2);
      if (a.S > 1) FetchOffsets(sm.idN, asmtG + static_cast<size_t>(sn) * taps * a.KgPad, taps, a.CT, a.KgPad, tid, T);
      CpAsyncCommit();
    }
    if (active) {
      const char* lutq = reinterpret_cast<const char*>(sm.lut) + wo0 * 4;
      const uint32_t* ip = sm.idx + (kh * a.ksz) * a.CT + cw * CPT;
      for (int kw = 0; kw < a.ksz; kw++) {
        const int phase = kw % a.stride, sh = kw / a.stride;
        GatherTap<CPT, J>(acc, lutq + (phase * PH + sh) * 4, ip);
        ip += a.CT;
      }
    }
    if (s == a.S - 1 && ho_cur < ho_end && kh == a.ksz - 1) {
#pragma unroll
      for (int j = 0; j < J; j++) {
        const int wo = wo0 + 32 * j;
        if (wo < a.Wo) {
          float* out = a.dst + ((static_cast<size_t>(n) * a.Ho + ho_cur) * a.Wo + wo) * a.Cout + g * a.Kg + cbase;
          StoreChannels<CPT, J>(out, acc, j, a.relu);
        }
      }
      acc.Fill(bias);
      ho_cur += a.rgroups;
    }
    hi = hin; s = sn;
  }
  for (int h = max(hi_hi, hi_begin); h < hi_end; h++) {
    const int kh = h + a.pad - ho_cur * a.stride;
    if (ho_cur < ho_end && kh == a.ksz - 1) {
#pragma unroll
      for (int j = 0; j < J; j++) {
        const int wo = wo0 + 32 * j;
        if (wo < a.Wo) {
          float* out = a.dst + ((static_cast<size_t>(n) * a.Ho + ho_cur) * a.Wo + wo) * a.Cout + g * a.Kg + cbase;
          StoreChannels<CPT, J>(out, acc, j, a.relu);
        }
      }
      acc.Fill(bias);
      ho_cur += a.rgroups;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemD), "r"(256) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// single-subspace, narrow-subspace layers (AlexNet conv1: S = 1, 3 of the 8 codebook dims exist).
// With S = 1 every LUT entry  LUT[pixel][k] = <x_pixel, c_k>  would be read only ~ (k/stride)^2 * Cout / K ~ 5 times,
// so tabulating costs more than it saves.  This kernel keeps the layer in PQ form (codebook + uint8 assignments) but
// evaluates the inner product where it is consumed:  acc[p][c] += <x[p + tap], c_{asmt[tap][c]}>  -- the same sum,
// the codeword (not the partial product) is what the warp-uniform assignment index gathers (128-bit broadcast load).
// A thread owns 2 output rows x 2 columns x CPT channels (float2 pairs, FFMA2); the strip's input rows live in shared
// memory, de-interleaved by stride phase, so a tap is a constant shift and there are no barriers after the tile load.
// ------------------------------------------------------------------------------------------------------------
template <int CPT, int NJ, int ROWS>
__global__ void __launch_bounds__(384, 1) conv_direct_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int taps = a.ksz * a.ksz;
  const int PH = a.PW, PITCH = a.PP;                // phase length, padded row pitch (floats)
  const int rowsIn = (a.R - 1) * a.stride + a.ksz;  // input rows of the strip
  float* xin = reinterpret_cast<float*>(smem);                                   // [rowsIn][NJ][PITCH]
  float2* cb = reinterpret_cast<float2*>(xin + static_cast<size_t>(rowsIn) * NJ * PITCH + 64);  // [K][4] (c,c) pairs
  uint32_t* idxa = reinterpret_cast<uint32_t*>(cb + a.K * 4);                    // [taps][CT] smem address of cb[k]

  const int tid = threadIdx.x, T = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int cw = warp % a.cwarps, rw = warp / a.cwarps;   // channel warp, row-pair warp
  int b = blockIdx.x;
  const int strip = b % a.nstrips; b /= a.nstrips;
  const int ct = b % a.nct;
  const int g = b / a.nct;
  const int n = blockIdx.y;
  const int ho0 = strip * a.R;
  const int ho_end = min(a.Ho, ho0 + a.R);
  const int cbase = ct * a.CT + cw * CPT;
  const size_t chStride = a.src_nchw ? static_cast<size_t>(a.Hi) * a.Wi : 1;
  const int pixStride = a.src_nchw ? 1 : a.Cin;
  const size_t rowStride = static_cast<size_t>(a.Wi) * pixStride;
  const float* src = a.src + static_cast<size_t>(n) * a.Hi * a.Wi * a.Cin + static_cast<size_t>(g) * a.Cg * chStride;
  const int nj = min(NJ, min(a.Cg, a.d));
  const int hi0 = ho0 * a.stride - a.pad;

  // input tile: element (row r, dim jj, column position) <- pixel (hi0 + r, wi), zero outside the image
  for (int e = tid; e < rowsIn * NJ * PITCH; e += T) {
    const int pos = e % PITCH;
    const int rj = e / PITCH;
    const int jj = rj % NJ, r = rj / NJ;
    const int phase = pos / PH, i = pos - phase * PH;
    const int wi = i * a.stride + phase - a.pad;
    const int hi = hi0 + r;
    const bool ok = phase < a.stride && wi >= 0 && wi < a.Wi && hi >= 0 && hi < a.Hi && jj < nj;
    CpAsync4(xin + e, src + (ok ? hi * rowStride + static_cast<size_t>(wi) * pixStride + jj * chStride : 0), ok);
  }
  CpAsyncCommit();
  for (int e = tid; e < a.K * 4; e += T) {
    const int k = e >> 2, jj = e & 3;
    const float c = (jj < nj) ? __ldg(a.ctrd + static_cast<size_t>(k) * a.d + jj) : 0.0f;
    cb[e] = make_float2(c, c);
  }
  {
    const uint8_t* ap = a.asmt + static_cast<size_t>(g) * taps * a.KgPad + ct * a.CT;  // S == 1
    const uint32_t cbBase = static_cast<uint32_t>(__cvta_generic_to_shared(cb));
    for (int e = tid; e < taps * a.CT; e += T) {
      const int tap = e / a.CT, c = e - tap * a.CT;
      idxa[e] = cbBase + static_cast<uint32_t>(__ldg(ap + tap * a.KgPad + c)) * 32u;
    }
  }
  float bias[CPT];
#pragma unroll
  for (int c = 0; c < CPT; c++) bias[c] = __ldg(a.bias + g * a.Kg + cbase + c);
  CpAsyncWaitAll();
  __syncthreads();

  const int groups = (ho_end - ho0 + ROWS - 1) / ROWS;
  for (int gr = rw; gr < groups; gr += a.rgroups) {
    const int ra = ROWS * gr;                    // strip-local output rows ra .. ra + ROWS - 1
    float2 acc[ROWS][CPT];
#pragma unroll
    for (int r = 0; r < ROWS; r++)
#pragma unroll
      for (int c = 0; c < CPT; c++) acc[r][c] = make_float2(bias[c], bias[c]);
    int rowOff[ROWS];                            // rows past the strip re-read row ra (results discarded)
#pragma unroll
    for (int r = 0; r < ROWS; r++) rowOff[r] = (ho0 + ra + r < ho_end ? r * a.stride : 0) * NJ * PITCH;
    for (int kh = 0; kh < a.ksz; kh++) {
      const float* rowA = xin + static_cast<size_t>(ra * a.stride + kh) * NJ * PITCH + lane;
      const uint32_t* ip = idxa + (kh * a.ksz) * a.CT + cw * CPT;
      for (int kw = 0; kw < a.ksz; kw++) {
        const int shift = (kw % a.stride) * PH + kw / a.stride;
        float2 x[ROWS][NJ];
#pragma unroll
        for (int r = 0; r < ROWS; r++)
#pragma unroll
          for (int jj = 0; jj < NJ; jj++)
            x[r][jj] = make_float2(rowA[rowOff[r] + jj * PITCH + shift], rowA[rowOff[r] + jj * PITCH + shift + 32]);
#pragma unroll
        for (int c4 = 0; c4 < CPT; c4 += 4) {
          const uint4 ad = *reinterpret_cast<const uint4*>(ip + c4);
          const uint32_t adr[4] = {ad.x, ad.y, ad.z, ad.w};
#pragma unroll
          for (int u = 0; u < 4; u++) {
            float4 c01;
            float2 c2 = make_float2(0.f, 0.f), c3 = make_float2(0.f, 0.f);
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c01.x), "=f"(c01.y), "=f"(c01.z), "=f"(c01.w) : "r"(adr[u]));
            if (NJ > 2) {
              if (NJ > 3) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(c2.x), "=f"(c2.y), "=f"(c3.x), "=f"(c3.y) : "r"(adr[u]));
              else asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+16];" : "=f"(c2.x), "=f"(c2.y) : "r"(adr[u]));
            }
            const int c = c4 + u;
#pragma unroll
            for (int r = 0; r < ROWS; r++) {
              acc[r][c] = __ffma2_rn(make_float2(c01.x, c01.y), x[r][0], acc[r][c]);
              if (NJ > 1) acc[r][c] = __ffma2_rn(make_float2(c01.z, c01.w), x[r][1 % NJ], acc[r][c]);
              if (NJ > 2) acc[r][c] = __ffma2_rn(c2, x[r][2 % NJ], acc[r][c]);
              if (NJ > 3) acc[r][c] = __ffma2_rn(c3, x[r][3 % NJ], acc[r][c]);
            }
          }
        }
        ip += a.CT;
      }
    }
    // emit: lanes hold columns lane and lane + 32 of each row
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
      if (ho0 + ra + r >= ho_end) break;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int wo = lane + 32 * half;
        if (wo < a.Wo) {
          float* out = a.dst + ((static_cast<size_t>(n) * a.Ho + ho0 + ra + r) * a.Wo + wo) * a.Cout + g * a.Kg + cbase;
#pragma unroll
          for (int c = 0; c < CPT; c += 4) {
            float4 o = half ? make_float4(acc[r][c].y, acc[r][c + 1].y, acc[r][c + 2].y, acc[r][c + 3].y)
                            : make_float4(acc[r][c].x, acc[r][c + 1].x, acc[r][c + 2].x, acc[r][c + 3].x);
            if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(out + c) = o;
          }
        }
      }
    }
  }
}

template <int CPT, int J, int MAXT>
int LaunchBound(const ConvPlan& p, const ConvArgs& a, cudaStream_t st) {
  dim3 grid(a.G * a.nct * a.nstrips, a.N);
  if (p.kernel == 0) {
    auto kern = conv_s1_kernel<CPT, J, MAXT>;
    QCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<grid, p.threads, p.smem, st>>>(a);
  } else if (p.kernel == 2) {
    auto kern = conv_s1_tc_kernel<CPT, J, MAXT>;
    QCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<grid, p.threads, p.smem, st>>>(a);
  } else if (p.kernel == 3) {
    auto kern = conv_roll_tc_kernel<CPT, J, MAXT>;
    QCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<grid, p.threads, p.smem, st>>>(a);
  } else {
    auto kern = conv_roll_kernel<CPT, J, MAXT>;
    QCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<grid, p.threads, p.smem, st>>>(a);
  }
  QCNN_CUDA(cudaGetLastError());
  return 0;
}

template <int CPT, int J>
int LaunchOne(const ConvPlan& p, const ConvArgs& a, cudaStream_t st) {
  if (p.threads <= 256) return LaunchBound<CPT, J, 256>(p, a, st);
  if (p.threads <= 384) return LaunchBound<CPT, J, 384>(p, a, st);
  return LaunchBound<CPT, J, 512>(p, a, st);
}

}  // namespace

namespace qcnn {

constexpr size_t kMaxCand = 18;

// Chooses the tiling of a conv layer for batch size N.  Cost model in SM-cycles per CTA, times the number of
// waves the whole batch needs on this GPU (so small batches trade LUT rebuilds for parallelism):
//   gather  = shared-memory wavefronts: one per (warp, tap, s, channel, position-slot) plus one 128-bit offset load
//             per 4 channels -> (1 + 1/(4J)) per lookup slot, 1 wavefront / clk / SM
//   build   = LUT entries built (incl. halo rows, garbage columns and per-channel-tile rebuilds) x
//             (0.66 * dims + 0.3) issue slots / 32 lanes / ~2.5 warp-instructions per clk
//   sync    = ~300 clk per (s [, input row]) phase pair
// The two stages do not overlap inside a CTA (one CTA per SM), so the costs add.
int PlanConv(qcnn_layer* L, int N) {
  if (L->plan_N == N) return 0;
  if (L->tunedPlans)  // a batch size seen (and timed) before: reuse its plan
    for (const std::pair<int, ConvPlan>& e : *L->tunedPlans)
      if (e.first == N) { L->plan = e.second; L->plan_N = N; L->tuned = 1; return 0; }
  std::vector<std::pair<double, ConvPlan>> cands;
  bool found = false;
  const int G = L->grp, Cg = L->Cin / G, Kg = L->Cout / G;
  const int taps = L->ksz * L->ksz;
  const size_t smemMax = L->ctx->smem_optin ? L->ctx->smem_optin : 227 * 1024;
  double dimsSum = 0;  // sum over subspaces of the dims that exist
  for (int s = 0; s < L->S; s++) dimsSum += std::max(0, std::min(Cg - s * L->d, L->d));
  const double ipe = 0.66 * (dimsSum / L->S) + 0.3;  // issue slots per LUT entry
  // tensor-core LUT stage: 128 codewords (MMA M), every subspace <= 8 dims (one K = 8 tf32 step); env QCNN_NO_TC=1 disables
  const bool tcOk = L->K == 128 && L->d <= 8 && getenv("QCNN_NO_TC") == nullptr;
  const int cpts[2] = {32, 16};
  for (int ci = 0; ci < 2; ci++) {
    const int CPT = cpts[ci];
    if (Kg % CPT != 0) continue;
    for (int J = 1; J * CPT <= 64 && J <= 4; J++) {
      for (int nct = 1; nct <= Kg / CPT; nct++) {
        if (Kg % nct != 0) continue;
        const int CT = Kg / nct;
        if (CT % CPT != 0 || CT % 16 != 0) continue;
        const int cwarps = CT / CPT;
        for (int R = 1; R <= L->Ho; R++) {
          for (int tc = 0; tc <= (tcOk ? 1 : 0); tc++) {
            ConvPlan p;
            memset(&p, 0, sizeof(p));
            ConvArgs& a = p.a;
            a.R = R;
            a.nstrips = CeilDiv(L->Ho, R);
            a.CT = CT; a.nct = nct; a.cwarps = cwarps;
            a.ksplit = 1;
            double gatherSlots, builtEntries, phases;
            int warps;
            size_t smemFloats;
            if (L->stride == 1) {
              p.kernel = tc ? 2 : 0;
              a.PW = L->Win + L->pad;
              a.RI = R + L->ksz - 1;
              a.pwarps = CeilDiv(R * a.PW, 32 * J);
              a.rgroups = 1;
              const int need1 = a.RI * a.PW + std::max(L->pad, L->ksz - 1) + 1;
              const int need2 = a.pwarps * 32 * J + (L->ksz - 1) * a.PW + (L->ksz - 1) + 1;
              warps = a.pwarps * cwarps;
              if (tc) {
                const int NM = RoundUp(std::max(need1, need2), 32);  // MMA N = LUT columns per subspace
                if (NM > 256 || warps < 4) continue;
                a.RI = NM;          // the tensor-core kernel receives NM in this field
                a.PP = NM + 4;      // row pitch = 4 (mod 32) words: conflict-free 128-bit drains from TMEM
                smemFloats = 128 * static_cast<size_t>(a.PP) + static_cast<size_t>(taps) * CT + 2048 + 25 * static_cast<size_t>(NM) + 1024 + 8;
              } else {
                a.PP = RoundUp(std::max(need1, need2), 4);
                smemFloats = static_cast<size_t>(L->K) * a.PP + static_cast<size_t>(taps) * CT + 24 * static_cast<size_t>(L->K) + 17 * static_cast<size_t>(a.PP);
              }
              gatherSlots = static_cast<double>(a.pwarps) * J * taps * L->S * CT;   // wavefronts of LUT reads
              builtEntries = static_cast<double>(a.PP) * L->K * L->S;
              phases = L->S;
            } else {
              p.kernel = tc ? 3 : 1;
              a.rgroups = CeilDiv(L->ksz, L->stride);
              a.pwarps = CeilDiv(L->Wo, 32 * J);
              const int ph1 = CeilDiv(L->Win + 2 * L->pad, L->stride);
              const int ph2 = a.pwarps * 32 * J + (L->ksz - 1) / L->stride + 1;
              warps = a.pwarps * cwarps * a.rgroups;
              const double rowsIn = (R - 1) * L->stride + L->ksz;
              if (tc) {
                // phase length = the real columns only; idle lanes beyond Wo then read into the next phase's columns
                // (valid shared memory, results discarded), so the row must hold (stride-1)*PH + lanes + shift entries
                a.PW = ph1;
                const int NM = RoundUp(std::max(ph1 * L->stride, (L->stride - 1) * ph1 + ph2), 32);
                if (NM > 256 || warps < 4) continue;
                a.RI = NM;
                a.PP = NM + 4;
                smemFloats = 128 * static_cast<size_t>(a.PP) + static_cast<size_t>(taps) * CT + 2048 + 25 * static_cast<size_t>(NM) + 1024 + 8;
              } else {
                a.PW = std::max(ph1, ph2);           // phase length
                a.PP = RoundUp(a.PW * L->stride, 4);
                a.RI = 0;
                smemFloats = static_cast<size_t>(L->K) * a.PP + static_cast<size_t>(taps) * CT + 24 * static_cast<size_t>(L->K) + 17 * static_cast<size_t>(a.PP);
              }
              gatherSlots = static_cast<double>(a.pwarps) * J * R * taps * L->S * CT;
              builtEntries = static_cast<double>(a.PP) * L->K * L->S * rowsIn;
              phases = L->S * rowsIn;
            }
            p.CPT = CPT; p.J = J;
            p.threads = warps * 32;
            if (p.threads > kMaxThreads || p.threads < 64) continue;
            // FFMA path: lut + idx + cb2 + xs[2] + posoff + cbN; tensor path: lut + idx + A/B tiles + raw staging
            p.smem = sizeof(float) * smemFloats + RoundUp(taps * CT, 16);
            if (p.smem > smemMax) continue;
            // measured on B200 (profiles/): ~1.3 clk per gather wavefront; FFMA LUT stage ~0.12 clk per entry at d = 8;
            // tensor-core LUT stage: the MMA is hidden behind the gather, what remains is the TMEM drain + operand split
            // gather latency hiding improves with resident warps (9 warps: 1.37 clk/wavefront measured, 12: 1.28)
            const double gather = gatherSlots * (1.0 + 1.0 / (4.0 * J)) * (1.15 + 1.6 / warps);
            const double build = tc ? builtEntries * 0.02 + 300.0 * phases : builtEntries * ipe / 5.5 * 0.12;
            const double perCta = gather + build + 400.0 * phases;
            // few resident warps cannot keep ~30 LDS in flight per SM nor feed the FMA pipe during the LUT stage
            // (measured: 2-warp CTAs run the build at ~0.5 IPC); 64 accumulators at 512 threads spill
            double occPenalty = warps < 4 ? 2.5 : (warps < 6 ? 1.6 : (warps < 8 ? 1.25 : 1.0));
            if (p.threads > 384 && J * CPT >= 64) occPenalty *= 1.1;
            const double ctas = static_cast<double>(G) * nct * a.nstrips * N;
            const double waves = std::ceil(ctas / L->ctx->sm_count);
            const double cost = perCta * waves * occPenalty;
            cands.emplace_back(cost, p);
            found = true;
          }
        }
      }
    }
  }
  // single subspace of <= 4 existing dims: evaluate <x, codeword> at the point of use (conv_direct_kernel)
  if (L->S == 1 && std::min(Cg, L->d) <= 4 && L->Wo <= 64 && getenv("QCNN_NO_DIRECT") == nullptr) {
    const int NJ = std::min(Cg, L->d) <= 3 ? 3 : 4;
    const int dcpts[2] = {16, 8};
    for (int ci = 0; ci < 2; ci++) {
      const int CPT = dcpts[ci];
      if (Kg % CPT != 0 || Kg % 16 != 0) continue;
      for (int nct = 1; nct <= 2; nct++) {
        if (Kg % nct != 0 || (Kg / nct) % 16 != 0 || (Kg / nct) % CPT != 0) continue;
        const int CT = Kg / nct, cwarps = CT / CPT;
        const int Rs[5] = {2, 4, 6, 8, 14};
        for (int ri = 0; ri < 5; ri++) {
          const int R = Rs[ri];
          if (R > L->Ho + 1) continue;
          for (int rgroups = 1; rgroups <= 2; rgroups++)
          for (int rowsPer = 2; rowsPer <= (CPT == 8 ? 4 : 2); rowsPer += 2) {
            if (rgroups > R / rowsPer) continue;
            ConvPlan p;
            memset(&p, 0, sizeof(p));
            ConvArgs& a = p.a;
            p.kernel = 4; p.CPT = CPT; p.J = rowsPer;
            a.R = std::min(R, L->Ho + (L->Ho & 1)); a.nstrips = CeilDiv(L->Ho, a.R);
            a.CT = CT; a.nct = nct; a.cwarps = cwarps; a.pwarps = 1; a.rgroups = rgroups; a.ksplit = 1;
            a.PW = CeilDiv(L->Win + 2 * L->pad, L->stride);       // phase length
            a.PP = RoundUp(a.PW * L->stride, 4);                   // row pitch
            p.threads = 32 * cwarps * rgroups;
            if (p.threads > 384 || p.threads < 128) continue;
            const int rowsIn = (a.R - 1) * L->stride + L->ksz;
            p.smem = sizeof(float) * (static_cast<size_t>(rowsIn) * NJ * a.PP + 64 + 8 * static_cast<size_t>(L->K) +
                                      static_cast<size_t>(taps) * CT);
            if (p.smem > smemMax) continue;
            const int pairs = CeilDiv(a.R, rowsPer);
            // issue slots: per (row pair, tap, channel) 2 loads + 2*NJ FFMA2, plus 4*NJ pixel loads per tap
            const double perWarp = static_cast<double>(CeilDiv(pairs, rgroups)) * taps * (CPT * (2.0 + rowsPer * NJ) + 2.0 * rowsPer * NJ + CPT / 4.0);
            const double perCta = perWarp * cwarps * rgroups / 2.6 + rowsIn * NJ * a.PP * 0.5;
            const double ctas = static_cast<double>(G) * nct * a.nstrips * N;
            const double waves = std::ceil(ctas / L->ctx->sm_count);
            cands.emplace_back(perCta * waves, p);
            found = true;
          }
        }
      }
    }
  }
  // decode-at-use GEMMs on the tensor cores (pq_gemm_tc.cu).  The kernel FAMILY is a setting, not a timing result:
  // with tensor_core = 1 (default) an eligible layer always runs pq_gemm_tc (3xTF32; tolerance in DESIGN.md 2), with
  // tensor_core = 0 always the LUT + gather kernels; on-device timing only chooses among tilings of that family, so the
  // numerical path of a (layer, batch size) is the same in every process.
  if (!L->opt_no_tc) {
    std::vector<std::pair<double, ConvPlan>> tc;
    PlanPqGemm(L, N, &tc);
    if (!tc.empty()) { cands.swap(tc); found = true; }
  }
  QCNN_CHECK(found, "qcnn_conv_layer_create: no tiling fits (Cout/grp=%d must be a multiple of 16; K=%d must be a "
             "multiple of 8; k=%d, W=%d)", Kg, L->K, L->ksz, L->Win);
  // keep the kMaxCand cheapest tilings (by the model); LaunchConv times them on the device the first time a batch size
  // is seen (QCNN_AUTOTUNE=0 keeps the model's first choice)
  std::stable_sort(cands.begin(), cands.end(), [](const std::pair<double, ConvPlan>& x, const std::pair<double, ConvPlan>& y) { return x.first < y.first; });
  if (!L->cands) L->cands = new std::vector<ConvPlan>();
  L->cands->clear();
  // layer parameter "force_kernel" / QCNN_FORCE_KERNEL=<0 s1 | 1 roll | 2 s1_tc | 3 roll_tc | 4 direct | 6 pq_gemm_tc>
  // restricts the choice (tests pin the kernel they check; fails when that kernel has no tiling for the layer)
  // Within the LUT + gather family the kernels come in two numerical classes -- fp32 LUT / decode (s1, roll, direct) and
  // 3xTF32 tensor-core LUT stage (s1_tc, roll_tc; ~5e-6 apart) -- and on-device timing must not pick between classes
  // either: the class of the cost model's favourite is kept, timing chooses among its tilings only.
  if (!cands.empty() && cands[0].second.kernel != 6 && !L->opt_force_kernel && getenv("QCNN_FORCE_KERNEL") == nullptr) {
    auto cls = [](int k) { return (k == 2 || k == 3) ? 1 : 0; };
    const int want = cls(cands[0].second.kernel);
    std::vector<std::pair<double, ConvPlan>> kept;
    for (const auto& c : cands) if (c.second.kernel != 6 && cls(c.second.kernel) == want) kept.push_back(c);
    cands.swap(kept);
  }
  if (L->opt_gemm_nt) {   // layer parameter "gemm_nt": only pq_gemm_tc tilings with that many positions per CTA
    std::vector<std::pair<double, ConvPlan>> kept;
    for (const auto& c : cands) if (c.second.kernel != 6 || c.second.g.NT == L->opt_gemm_nt) kept.push_back(c);
    QCNN_CHECK(!kept.empty(), "qcnn_conv_aprx_forward: gemm_nt=%d has no tiling for this layer at batch %d", L->opt_gemm_nt, N);
    cands.swap(kept);
  }
  const char* force = getenv("QCNN_FORCE_KERNEL");
  if (L->opt_force_kernel || force) {
    const int want = L->opt_force_kernel ? L->opt_force_kernel - 1 : atoi(force);
    std::vector<std::pair<double, ConvPlan>> kept;
    for (const auto& c : cands) if (c.second.kernel == want) kept.push_back(c);
    QCNN_CHECK(!kept.empty() || !L->opt_force_kernel, "qcnn_conv_aprx_forward: force_kernel=%d has no tiling for this layer at batch %d", want, N);
    if (!kept.empty()) cands.swap(kept);
  }
  // every kernel family that has a feasible tiling gets at least two seats among the candidates
  for (int pass = 0; pass < 2; pass++) {
    int perKernel[7] = {0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < cands.size() && L->cands->size() < kMaxCand; i++) {
      ConvPlan c = cands[i].second;
      if (pass == 0 && perKernel[c.kernel] >= (c.kernel == 4 ? 8 : (c.kernel == 6 ? 6 : 2))) continue;
      ConvArgs& a = c.a;
      a.Hi = L->Hin; a.Wi = L->Win; a.Cin = L->Cin; a.Ho = L->Ho; a.Wo = L->Wo; a.Cout = L->Cout;
      a.ksz = L->ksz; a.pad = L->pad; a.stride = L->stride; a.G = G; a.Cg = Cg; a.Kg = Kg;
      a.KgPad = RoundUp(Kg, 16);
      a.S = L->S; a.K = L->K; a.d = L->d;
      // skip near-duplicates: same kernel/CPT/J/channel tiling and a strip count already present
      bool dup = false;
      for (const ConvPlan& e : *L->cands)
        if (e.kernel == c.kernel && e.CPT == c.CPT && e.J == c.J && e.a.nct == c.a.nct && e.a.nstrips == c.a.nstrips &&
            e.a.rgroups == c.a.rgroups && e.a.R == c.a.R && (e.kernel != 4 || e.J == c.J)) dup = true;
      if (dup) continue;
      L->cands->push_back(c);
      perKernel[c.kernel]++;
    }
  }
  // the model's overall favourite first (it is the plan used when autotuning is off)
  for (size_t i = 1; i < L->cands->size(); i++)
    if ((*L->cands)[i].kernel == cands[0].second.kernel && (*L->cands)[i].CPT == cands[0].second.CPT &&
        (*L->cands)[i].J == cands[0].second.J && (*L->cands)[i].a.nct == cands[0].second.a.nct &&
        (*L->cands)[i].a.nstrips == cands[0].second.a.nstrips && (*L->cands)[i].a.rgroups == cands[0].second.a.rgroups) {
      std::swap((*L->cands)[0], (*L->cands)[i]);
      break;
    }
  const ConvPlan best = (*L->cands)[0];
  L->plan = best;
  L->plan_N = N;
  L->tuned = 0;
  return 0;
}

template <int CPT, int NJ, int ROWS>
static int LaunchDirect(const ConvPlan& p, const ConvArgs& a, cudaStream_t st) {
  dim3 grid(a.G * a.nct * a.nstrips, a.N);
  auto kern = conv_direct_kernel<CPT, NJ, ROWS>;
  QCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  kern<<<grid, p.threads, p.smem, st>>>(a);
  QCNN_CUDA(cudaGetLastError());
  return 0;
}

static int LaunchPlan(qcnn_layer* L, const ConvPlan& p, const float* src, int N, float* dst, int relu, cudaStream_t st) {
  ConvArgs a = p.a;
  a.src = src; a.dst = dst; a.ctrd = L->d_ctrd; a.asmt = L->d_asmt; a.bias = L->d_bias;
  a.N = N; a.relu = relu; a.src_nchw = L->src_nchw;
  if (p.kernel == 6) return LaunchPqGemm(L, p, src, N, dst, relu, st);
  if (p.kernel == 4) {
    const int nj = std::min(a.Cg, a.d);
    // p.J carries the output rows per thread (2 or 4)
    if (p.CPT == 16) return nj <= 3 ? LaunchDirect<16, 3, 2>(p, a, st) : LaunchDirect<16, 4, 2>(p, a, st);
    if (p.J == 4) return nj <= 3 ? LaunchDirect<8, 3, 4>(p, a, st) : LaunchDirect<8, 4, 4>(p, a, st);
    return nj <= 3 ? LaunchDirect<8, 3, 2>(p, a, st) : LaunchDirect<8, 4, 2>(p, a, st);
  }
  QCNN_CHECK(!(a.src_nchw && (p.kernel == 0 || p.kernel == 2)), "qcnn_conv_aprx_forward: NCHW source is only supported by the strided kernel");
  int rc = 1;
#define QCNN_DISPATCH(C, JJ) if (p.CPT == C && p.J == JJ) rc = LaunchOne<C, JJ>(p, a, st); else
  QCNN_DISPATCH(32, 1) QCNN_DISPATCH(32, 2) QCNN_DISPATCH(16, 1) QCNN_DISPATCH(16, 2) QCNN_DISPATCH(16, 3)
  QCNN_DISPATCH(16, 4) { SetError("internal: no conv instantiation for CPT=%d J=%d", p.CPT, p.J); return 1; }
#undef QCNN_DISPATCH
  return rc;
}

int LaunchConv(qcnn_layer* L, const float* src, int N, float* dst, int relu, cudaStream_t st) {
  QCNN_CHECK(L->kind == QCNN_KIND_CONV, "qcnn_conv_aprx_forward: layer is not convolutional");
  QCNN_CHECK(N >= 1 && N <= 65535, "qcnn_conv_aprx_forward: N must be in [1, 65535] (got %d)", N);
  if (int prc = PlanConv(L, N)) return prc;
  // empirical choice among the model's best tilings: time each once on this device for this batch size
  static const bool autotune = !(getenv("QCNN_AUTOTUNE") && getenv("QCNN_AUTOTUNE")[0] == '0');
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  if (autotune && !L->opt_no_autotune && !L->tuned && L->cands && L->cands->size() > 1 && cap == cudaStreamCaptureStatusNone) {
    cudaEvent_t e0, e1;
    QCNN_CUDA(cudaEventCreate(&e0));
    if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return CudaFail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__); }
    float bestMs = 1e30f;
    size_t bestI = 0;
    for (size_t i = 0; i < L->cands->size(); i++) {
      const ConvPlan& c = (*L->cands)[i];
      if (LaunchPlan(L, c, src, N, dst, relu, st)) continue;   // warm-up (also sets the smem attribute)
      // best of two timed samples (a sample = 1 launch at large batches, 3 at small ones): single samples made the choice
      // between near-equal candidates vary from run to run
      const int reps = N >= 64 ? 1 : 3;
      float ms = 1e30f;
      bool ok = true;
      for (int sample = 0; sample < 2 && ok; sample++) {
        if (cudaEventRecord(e0, st) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        for (int r = 0; r < reps; r++) LaunchPlan(L, c, src, N, dst, relu, st);
        if (cudaEventRecord(e1, st) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        float t = 0.0f;
        cudaEventElapsedTime(&t, e0, e1);
        ms = std::min(ms, t);
      }
      if (!ok) continue;
      static const bool tuneLog = getenv("QCNN_AUTOTUNE_LOG") != nullptr;
      if (tuneLog)
        fprintf(stderr, "[qcnn autotune] N=%d Cout=%d k=%d cand %zu: kernel=%d NT/CPT=%d GT/J=%d lite=%d wide=%d bf=%d nsplit=%d smem=%zu -> %.4f ms\n",
                N, L->Cout, L->ksz, i, c.kernel, c.CPT, c.kernel == 6 ? c.g.GT : c.J, c.g.lite, c.g.wide, c.g.bf, c.g.nsplit, c.smem, ms);
      if (ms < bestMs) { bestMs = ms; bestI = i; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    L->plan = (*L->cands)[bestI];
    L->tuned = 1;
    if (!L->tunedPlans) L->tunedPlans = new std::vector<std::pair<int, ConvPlan>>();
    L->tunedPlans->emplace_back(N, L->plan);
  }
  const int rc = LaunchPlan(L, L->plan, src, N, dst, relu, st);
  if (rc == 0) L->ctx->launches++;
  return rc;
}

// human-readable tiling of the plan for batch N (bench / DESIGN bookkeeping)
int DescribeConv(qcnn_layer* L, int N, char* buf, size_t cap) {
  if (int prc = PlanConv(L, N)) return prc;
  const ConvPlan& p = L->plan;
  if (p.kernel == 6) {
    int ksteps = 0;
    for (int c = 0; c < p.g.nChunks; c++) ksteps += p.g.chunkCount[p.g.mode == 1 ? c : 0];
    snprintf(buf, cap, "pq_gemm_tc(tcgen05, weights decoded into TMEM%s) mode=%d NT=%d GT=%d slots=%d smem=%zuB grid=%d NPOS=%d "
             "chunks=%d ksteps=%d nsplit=%d", p.g.bf ? (p.g.lite ? ", bf16x2, 2 CTAs/SM" : (p.g.wide ? ", bf16x2, 8 decoder warps" : (p.g.xl ? ", bf16x2, 16 warps" : ", bf16x2")))
                    : (p.g.lite ? ", 2 CTAs/SM" : (p.g.wide ? ", 8 decoder warps" : "")), p.g.mode, p.g.NT, p.g.GT, p.g.NSLOT, p.smem,
             CeilDiv(N * p.g.IB, p.g.NT) * L->grp * p.g.nct * std::max(1, p.g.nsplit), p.g.NPOS, p.g.nChunks,
             ksteps / std::max(1, p.g.nsplit), std::max(1, p.g.nsplit));
    return 0;
  }
  snprintf(buf, cap, "%s CPT=%d J=%d threads=%d smem=%zuB grid=(%d,%d) R=%d strips=%d CT=%d nct=%d PP=%d pwarps=%d "
           "cwarps=%d rgroups=%d", p.kernel == 0 ? "conv_s1" : (p.kernel == 2 ? "conv_s1_tc(tcgen05 LUT)" : (p.kernel == 3 ? "conv_roll_tc(tcgen05 LUT)" : (p.kernel == 4 ? "conv_direct" : "conv_roll"))), p.CPT, p.J, p.threads, p.smem,
           p.a.G * p.a.nct * p.a.nstrips, N, p.a.R, p.a.nstrips, p.a.CT, p.a.nct, p.a.PP, p.a.pwarps, p.a.cwarps,
           p.a.rgroups);
  return 0;
}

}  // namespace qcnn
