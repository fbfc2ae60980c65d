// PQ layers as decode-at-use GEMMs on the 5th-generation tensor cores, weights decoded straight into tensor memory.
//
//     dst[q][c] = bias[c] + sum_{k-steps} sum_{j<8}  W_kstep[c][j] * X_kstep[q][j]
//
// One k-step is 8 input values of every position q (two 4-float halves); W_kstep[c][half] is the 4-float piece of the
// codeword that channel c's uint8 assignment index names for that half -- the same codebook + index arrays the LUT /
// gather kernels use (reference CaffeEva::CalcFeatMap_ConvAprx / _FCntAprx + GetInPdMat, src/CaffeEva.cc:760-868,
// 968-1025, 1261-1296, evaluated as x . (decoded w) instead of gather(LUT(x)); same sums, different association).
//
// Roles inside a CTA (warp-specialised, mbarrier pipelines, no CTA-wide barrier in the main loop):
//   warps 0-3  decoders: thread = output channel (= TMEM lane).  Per k-step: index byte(s) -> codeword piece(s) from the
//   (+ 8-11)   staged codebook slice -> tcgen05.st into the A ring in TMEM (16 columns per k-step).  Decoded weights never
//              touch shared memory.  bf16x2 operands (default): the codebook was split into bf16 pieces {w1, w2} once at
//              layer creation, the decoders convert nothing; conv layers: all tap indices of a chunk sit in registers
//              (channel-major index block, one or two 128-bit loads per chunk), so a k-step costs ONE dependent
//              shared-memory round trip (the codeword gather) -- under the MMAs' operand fetch a round trip is ~450 clk.
//   warp 4     one elected thread issues tcgen05.mma (M = 128 channels, N = NT positions) with A in TMEM and B = the
//              position planes in shared memory: two kind::f16 MMAs of K = 16 per k-step (bf16x2:
//              [w1|w1].[x1|x2] + [w2|w2].[x1|x2]) or three kind::tf32 MMAs of K = 8 (3xTF32: Ah*Bh + Ah*Bl + Al*Bh).
//   warps 5-7  stagers: the next chunk's positions go global -> registers -> planes (K-major, SWIZZLE_NONE core matrices:
//   (+ 12-15)  8 positions x 16 B, SBO = 128 B, LBO = distance between the two halves / pieces), buffered against the MMAs
//              of the current chunk (3x3 tiles: two register sets, loads two chunks ahead); a chunk's codebook slices and
//              index blocks arrive by cp.async.bulk counted on the chunk's mbarrier (3xTF32: cp.async), four chunks deep.
//              FC layers: the planes were pre-split by fc_prep_kernel and arrive by one cp.async.bulk per chunk.
//   all warps  epilogue: TMEM (lane = channel, column = position) -> + bias, ReLU -> NHWC stores (a warp writes 32
//              consecutive channels of one position: 128 B).
// Instantiations: 8 warps (FC tiles, 3xTF32), LITE = 8 warps at <= 128 registers, two CTAs per SM (tiles of <= 128
// positions), WIDE = 12 warps (second decoder group; conv2-5), XL = 16 warps (second decoder group + seven stager warps;
// conv1).  What paces the kernel is the latency of shared memory under load (MMA operand fetch + codeword gathers keep the
// pipe 65-70 % busy), not a throughput limit: profiles/README.md, "Role cycle counters".
//
// Layer geometries are expressed as a table of k-steps (KStep: B start / half distance inside the staged planes, index
// row and codebook slot of either half), so one main loop serves
//   mode 0  stride-1 convolutions: all images form one flat padded grid (image block = (Hi+pad) rows of pitch Wi+pad;
//           the leading pad rows / columns are zero and double as the previous row's / image's trailing padding), a tap
//           is a start-address shift of the B descriptor, a chunk is 8 input channels x all taps;
//   mode 1  strided convolutions with <= 4 input channels (conv1: 11x11 / 4): input de-interleaved into stride x stride
//           phase planes so that a tap again is a shift; a k-step pairs two taps (4 floats each: channels + zero pad);
//           a chunk is one phase row;
//   mode 2  fully-connected layers: position = image, a chunk is a run of k-steps over consecutive input features.
#include "qcnn_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>

namespace {

constexpr int kThreads = 256;
constexpr int kDecoders = 128;   // warps 0-3
constexpr int kStagers = 96;     // warps 5-7
constexpr int kStager0 = 160;
constexpr int kMaxSlots = 5;
constexpr int kMaxGT = 8;        // k-steps per stage (TMEM ring: NSLOT * GT * 16 columns <= 256)
constexpr int kRegPos = 16;     // float4 a stager thread holds in registers (convolution modes: planeF4 <= 16 * 96)
// "lite" instantiation: TWO CTAs per SM (<= 128 registers per thread, 256 TMEM columns and ~110 KB of shared memory each),
// so that one CTA's set-up, first-plane wait and epilogue overlap the other's MMA phase -- with one CTA per SM the tensor
// pipe idles 35-45 % of a CTA's life (profiles/README.md).  Tiles of <= 128 positions, shorter register windows.
constexpr int kRegPosLite = 11;
constexpr int kMaxGTLite = 4;
constexpr int kCbBufs = 4;      // codebook / index buffers: the decoders run ahead of the position planes

__device__ __forceinline__ uint32_t SmemU32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void CpAsync16(void* smemDst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0: nothing is read, the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(SmemU32(smemDst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void CpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void MbarInit(uint64_t* mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemU32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarArrive(uint64_t* mbar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(SmemU32(mbar)) : "memory");
}
// bounded spin: a protocol error traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void MbarWait(uint64_t* mbar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; spins++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(SmemU32(mbar)), "r"(parity) : "memory");
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void UmmaCommit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemU32(mbar)) : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool ElectOne() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void UmmaTf32Ts(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmemD), "r"(tmemA), "l"(descB),
               "r"(idesc), "r"(accumulate) : "memory");
}
// 3xTF32 operand split: hi = v truncated to tf32 (what the tensor core reads), lo = exact remainder (the tensor core
// truncates it to tf32 in turn).  Rounding both pieces to nearest instead was measured to change the end error by < 10 %:
// the error of this path (1e-5 .. 2e-5 of max|out| per layer, ~n * 2^-24) comes from the tensor core's fp32
// ACCUMULATION over the n = 3 * k-steps chained MMAs, not from the operand representation (tools/accuracy.py).
__device__ __forceinline__ void SplitTf32x4(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
}

// one k-step of decoded weights (two 4-float halves) -> hi | lo -> 16 TMEM columns of this thread's lane
__device__ __forceinline__ void StoreWeights(uint32_t taddr, const float4 w0, const float4 w1) {
  float4 h0, l0, h1, l1;
  SplitTf32x4(w0, h0, l0);
  SplitTf32x4(w1, h1, l1);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr),
                  "r"(__float_as_uint(h0.x)), "r"(__float_as_uint(h0.y)), "r"(__float_as_uint(h0.z)), "r"(__float_as_uint(h0.w)),
                  "r"(__float_as_uint(h1.x)), "r"(__float_as_uint(h1.y)), "r"(__float_as_uint(h1.z)), "r"(__float_as_uint(h1.w)),
                  "r"(__float_as_uint(l0.x)), "r"(__float_as_uint(l0.y)), "r"(__float_as_uint(l0.z)), "r"(__float_as_uint(l0.w)),
                  "r"(__float_as_uint(l1.x)), "r"(__float_as_uint(l1.y)), "r"(__float_as_uint(l1.z)), "r"(__float_as_uint(l1.w))
               : "memory");
}

// ---- bf16x2 operands: v = v1 + v2 with v1 = bf16(v), v2 = bf16(v - v1) (|v - v1 - v2| <= 2^-18 |v|).  One k-step of
// eight values is TWO kind::f16 MMAs of K = 16: [w1 | w1] . [x1 | x2] + [w2 | w2] . [x1 | x2] = (w1 + w2)(x1 + x2) --
// all four cross terms, at the fp16/bf16 rate (twice tf32's), against three K = 8 tf32 MMAs for 3xTF32.
__device__ __forceinline__ uint32_t PackBf16x2(float lo, float hi) {   // low 16 bits = bf16(lo), high 16 bits = bf16(hi)
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void SplitBf16x4(const float4 v, uint2& p1, uint2& p2) {
  p1.x = PackBf16x2(v.x, v.y);
  p1.y = PackBf16x2(v.z, v.w);
  // residuals against the rounded pieces (exact in fp32)
  const float rx = v.x - __uint_as_float(p1.x << 16), ry = v.y - __uint_as_float(p1.x & 0xFFFF0000u);
  const float rz = v.z - __uint_as_float(p1.y << 16), rw = v.w - __uint_as_float(p1.y & 0xFFFF0000u);
  p2.x = PackBf16x2(rx, ry);
  p2.y = PackBf16x2(rz, rw);
}
// one k-step of decoded weights -> [w1 | w1 | w2 | w2] (8 bf16 = 4 columns each) -> 16 TMEM columns of this thread's lane.
// PRE: the staged codebook already holds {w1, w2} per 4-float piece (qcnn_layer::d_ctrd_bf, split once at layer creation):
// the decoders convert nothing; otherwise split here.
template <bool PRE>
__device__ __forceinline__ void StoreWeightsBf(uint32_t taddr, const float4 w0, const float4 w1) {
  uint2 a1, a2, b1, b2;
  if (PRE) {
    a1 = make_uint2(__float_as_uint(w0.x), __float_as_uint(w0.y)); a2 = make_uint2(__float_as_uint(w0.z), __float_as_uint(w0.w));
    b1 = make_uint2(__float_as_uint(w1.x), __float_as_uint(w1.y)); b2 = make_uint2(__float_as_uint(w1.z), __float_as_uint(w1.w));
  } else {
    SplitBf16x4(w0, a1, a2);
    SplitBf16x4(w1, b1, b2);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr),
                  "r"(a1.x), "r"(a1.y), "r"(b1.x), "r"(b1.y), "r"(a1.x), "r"(a1.y), "r"(b1.x), "r"(b1.y),
                  "r"(a2.x), "r"(a2.y), "r"(b2.x), "r"(b2.y), "r"(a2.x), "r"(a2.y), "r"(b2.x), "r"(b2.y)
               : "memory");
}
__device__ __forceinline__ void UmmaF16Ts(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmemD), "r"(tmemA), "l"(descB),
               "r"(idesc), "r"(accumulate) : "memory");
}

// fp32 codebook [S][K][P pieces of 4 floats] -> [S][P][K] x {w1, w2}: the K pieces of one (subspace, piece) are contiguous
__global__ void split_codebook_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, int S, int K, int P) {
  const size_t n = static_cast<size_t>(S) * K * P;
  const size_t gs = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += gs) {
    const int pc = static_cast<int>(i % P);
    const size_t sk = i / P;
    const int k = static_cast<int>(sk % K);
    const size_t sidx = sk / K;
    uint2 p1, p2;
    SplitBf16x4(src[i], p1, p2);
    dst[(sidx * P + pc) * K + k] = make_uint4(p1.x, p1.y, p2.x, p2.y);
  }
}

// byte t (< 32) of the 32 bytes {lo, hi} held in registers (t is warp-uniform: selects, no local memory)
__device__ __forceinline__ int ByteOf32(const uint4 lo, const uint4 hi, int t) {
  const uint32_t x = (t & 16) ? hi.x : lo.x, y = (t & 16) ? hi.y : lo.y, z = (t & 16) ? hi.z : lo.z, w = (t & 16) ? hi.w : lo.w;
  const uint32_t xy = (t & 4) ? y : x, zw = (t & 4) ? w : z;
  const uint32_t v = (t & 8) ? zw : xy;
  return static_cast<int>((v >> ((t & 3) * 8)) & 0xFFu);
}

// byte t (< 48) of the 48 bytes {q0, q1, q2}
__device__ __forceinline__ int ByteOf48(const uint4 q0, const uint4 q1, const uint4 q2, int t) {
  const int hi = t >> 4;
  const uint4 q = hi == 0 ? q0 : (hi == 1 ? q1 : q2);
  const uint32_t xy = (t & 4) ? q.y : q.x, zw = (t & 4) ? q.w : q.z;
  const uint32_t v = (t & 8) ? zw : xy;
  return static_cast<int>((v >> ((t & 3) * 8)) & 0xFFu);
}

struct SmemMap {  // byte offsets inside the dynamic shared memory
  int planes, cbs, ids, tab, posoff, posrow, posdst, outoff, bias, bars, tmem, total;
};
__host__ __device__ inline SmemMap MapSmem(const GemmArgs& a) {
  SmemMap m;
  int o = 0;
  m.planes = o; o += a.nPB * 2 * a.planeRows * 16;      // [buf][hi,lo | x1,x2][planeRows] 16-byte rows
  m.cbs = o;    o += kCbBufs * a.cbSlots * a.cbF4 * 16; // [cbuf][slot][cbF4] codeword pieces (raw fp32)
  m.ids = o;    o += kCbBufs * a.idRows * 128;          // [cbuf][row][128 channels] assignment indices
  m.tab = o;    o += a.ntab * 16;                       // k-step table for the decoders that do not keep their indices in registers
  m.posoff = o; o += a.planeF4 * 4;                     // source element offset of every staged float4 (-1: zero)
  m.posrow = o; o += a.mode == 1 ? a.planeF4 * 4 : 0;   // mode 1: first input row of the position (phase row 0)
  m.posdst = o; o += (a.bf && a.mode != 2) ? a.planeF4 * 4 : 0;   // bf16x2: byte offset of every staged float4 inside a plane
  m.outoff = o; o += 256 * 4;                           // destination element offset of every position (-1: none)
  m.bias = o;   o += 128 * 4;
  m.bars = o;   o += 8 * (2 * kMaxSlots + 2 * kCbBufs + 9);
  m.tmem = o;   o += 16;
  m.total = o;
  return m;
}

// DBG: cycle counters of the three roles (QCNN_GEMM_DBG=1) -- a separate instantiation, the production kernel reads no clocks
// WIDE: a SECOND group of four decoder warps (warps 8-11; 384 threads): the two groups take the stages alternately.  The
// decoders -- one warp per SM sub-partition, a chain of dependent shared-memory reads, the split and a tcgen05.st per
// k-step -- are what paces the kernel (~400 clk per k-step against 384 clk of 3xTF32 MMAs or 256 clk of bf16x2 MMAs).
// XL: sixteen warps (512 threads, <= 128 registers): two decoder groups AND seven stager warps (5-7, 12-15) -- conv1's tiles
// (16-18 k-steps per phase row, four phase columns of 3-channel pixels to stage per row) need both to keep the MMAs fed.
template <bool DBG, bool LITE, bool BF, bool WIDE, bool XL = false>
__global__ void __launch_bounds__(XL ? 512 : (WIDE ? 384 : kThreads), LITE ? 2 : 1) pq_gemm_tc_kernel(const GemmArgs a) {
  constexpr int NTHR = XL ? 512 : (WIDE ? 384 : kThreads);
  constexpr int NSTG = XL ? kStagers + 128 : kStagers;      // stager threads
  constexpr bool TWOG = WIDE || XL;                          // two decoder groups (warps 0-3 and 8-11)
  constexpr int RP = XL ? 7 : ((LITE || WIDE) ? kRegPosLite : kRegPos);
  constexpr int MG = (LITE || XL) ? kMaxGTLite : (WIDE ? 5 : kMaxGT);
  constexpr uint32_t kTmemCols = LITE ? 256u : 512u;
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemMap sm = MapSmem(a);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = a.K, NT = a.NT, GT = a.GT, NSLOT = a.NSLOT;

  float4* planes = reinterpret_cast<float4*>(smem + sm.planes);
  float4* cbs = reinterpret_cast<float4*>(smem + sm.cbs);
  uint8_t* ids = smem + sm.ids;
  KStep* tabS = reinterpret_cast<KStep*>(smem + sm.tab);   // (FC tiles: shared memory is lightly loaded there and an LDS beats an indexed LDC)
  int* posoff = reinterpret_cast<int*>(smem + sm.posoff);
  int* posrow = reinterpret_cast<int*>(smem + sm.posrow);
  int* posdst = reinterpret_cast<int*>(smem + sm.posdst);
  int* outoff = reinterpret_cast<int*>(smem + sm.outoff);
  float* biasS = reinterpret_cast<float*>(smem + sm.bias);
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem + sm.bars);   // [kMaxSlots] decoders -> issuer
  uint64_t* emptyA = fullA + kMaxSlots;                            // [kMaxSlots] MMAs retired -> decoders
  uint64_t* fullB = emptyA + kMaxSlots;                            // [<=4] position planes: stagers -> issuer
  uint64_t* emptyB = fullB + 4;                                    // [<=4] chunk's MMAs retired -> stagers
  uint64_t* fullC = emptyB + 4;                                    // [kCbBufs] codebook + indices: stagers -> decoders
  uint64_t* emptyC = fullC + kCbBufs;                              // [kCbBufs] decoders done with the chunk -> stagers
  uint64_t* doneBar = emptyC + kCbBufs;
  uint32_t* tmemBase = reinterpret_cast<uint32_t*>(smem + sm.tmem);

  int b = blockIdx.x;
  const int ct = b % a.nct; b /= a.nct;
  int g = 0, split = 0;
  if (a.nsplit > 1) { split = b % a.nsplit; b /= a.nsplit; }
  if (a.mode != 2) { g = b % a.G; b /= a.G; }
  const int tile = b;
  const int ch0 = ct * 128;                            // first channel of the tile inside the group
  const int CTv = min(128, a.Kg - ch0);                // valid channels
  const int Q0 = tile * NT;                            // first flat position
  const int i0 = Q0 / a.IB;                            // first image touched
  const float* srcBase = a.src + static_cast<size_t>(i0) * a.srcImg;
  float* dstBase = a.dst + static_cast<size_t>(i0) * a.dstImg + g * a.Kg + ch0;
  // mode 2: this CTA's k-step range [k0, k0 + kTotal) of the layer; partial sums go to their own plane
  const int k0 = split * a.kPerSplit;
  const int kTotal = a.mode == 2 ? min(a.kPerSplit, a.kAll - k0) : 0;
  // convolution modes: K split by chunks [kcBase, kcBase + nChunks) (small batches: more CTAs than position tiles)
  const int kcBase = a.mode == 2 ? 0 : split * a.kPerSplit;
  const int nChunks = a.mode == 2 ? (kTotal + a.chunkCount[0] - 1) / a.chunkCount[0]
                                  : (a.nsplit > 1 ? min(a.kPerSplit, a.nChunks - kcBase) : a.nChunks);
  if (a.mode == 2 && a.nsplit > 1) dstBase = a.partial + (static_cast<size_t>(split) * a.N + Q0) * a.dstRow + ch0;
  if (a.mode != 2 && a.nsplit > 1)
    dstBase = a.partial + (static_cast<size_t>(split) * a.N + i0) * a.dstImg + g * a.Kg + ch0;

  // ---- set-up (all threads) ----
  for (int e = tid; e < a.ntab; e += NTHR) tabS[e] = a.tab[e];
  for (int p = tid; p < a.planeF4; p += NTHR) {
    int off = -1;
    if (a.mode == 0) {
      // float4 p: half = p / NPOS, position = p % NPOS; the half only selects the channel offset (added per chunk)
      const int pos = p % a.NPOS;
      const int F = Q0 + pos;
      const int i = F / a.IB, rem = F - i * a.IB;
      const int r = rem / a.PW, c = rem - r * a.PW;
      if (i < a.N && r >= a.pad && c >= a.pad) off = (((i - i0) * a.Hi + (r - a.pad)) * a.Wi + (c - a.pad)) * a.Cin;
    } else if (a.mode == 1) {
      // float4 p: phase column pw = p / NPOS, position = p % NPOS -> pixel (r*stride + ph - pad, c*stride + pw - pad);
      // the phase row ph is the chunk, so the row part is resolved per chunk from posrow
      const int pw = p / a.NPOS, pos = p - pw * a.NPOS;
      const int F = Q0 + pos;
      const int i = F / a.IB, rem = F - i * a.IB;
      const int r = rem / a.PW, c = rem - r * a.PW;
      const int wi = c * a.stride + pw - a.pad;
      const bool colOk = i < a.N && wi >= 0 && wi < a.Wi;
      off = static_cast<int>((i - i0) * a.srcImg) + (r * a.stride - a.pad) * a.rowStride + wi * a.colStride;
      posrow[p] = colOk ? r * a.stride - a.pad : -(1 << 28);
    }
    posoff[p] = off;   // (mode 2 stages its planes by bulk copy: no offsets)
    if (BF && a.mode != 2) {
      // bf16x2 planes hold 16-byte rows of eight values: mode 0 row = position, the half selects bytes 0-7 / 8-15;
      // mode 1 row = (phase-column pair, position), the column's parity selects the half
      const int grp = p / a.NPOS, pos = p - grp * a.NPOS;
      posdst[p] = a.mode == 0 ? pos * 16 + grp * 8 : ((grp >> 1) * a.NPOS + pos) * 16 + (grp & 1) * 8;
    }
  }
  for (int p = tid; p < 256; p += NTHR) {
    int off = -1;
    if (p < NT) {
      const int Q = Q0 + p;
      const int i = Q / a.IB, rem = Q - i * a.IB;
      const int ho = rem / a.PW, wo = rem - ho * a.PW;
      if (a.mode == 2) { if (Q < a.N) off = p * a.dstRow; }
      else if (i < a.N && ho < a.Ho && wo < a.Wo) off = (((i - i0) * a.Ho + ho) * a.Wo + wo) * a.Cout;
    }
    outoff[p] = off;
  }
  if (a.mode == 1 && a.bulkC)      // slot 1 (unpaired taps) stays zero; slot 0 arrives by bulk copy
    for (int e = tid; e < kCbBufs * K; e += NTHR) cbs[(e / K) * a.cbSlots * K + K + (e % K)] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  for (int c = tid; c < 128; c += NTHR) biasS[c] = (c < CTv && split == 0) ? __ldg(a.bias + g * a.Kg + ch0 + c) : 0.0f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemU32(tmemBase)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < kMaxSlots; i++) { MbarInit(fullA + i, kDecoders); MbarInit(emptyA + i, 1); }
    for (int i = 0; i < 4; i++) { MbarInit(fullB + i, NSTG); MbarInit(emptyB + i, 1); }
    for (int i = 0; i < kCbBufs; i++) { MbarInit(fullC + i, a.bulkC ? 1 : NSTG); MbarInit(emptyC + i, TWOG ? 2 * kDecoders : kDecoders); }
    MbarInit(doneBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmemD = *tmemBase;          // D: columns [0, NT)
  const uint32_t tmemA = tmemD + static_cast<uint32_t>(a.aOff);   // A ring: columns [aOff, aOff + NSLOT*GT*16)

  const int warpU = __shfl_sync(0xffffffffu, warp, 0);   // provably warp-uniform role selector
  if ((warpU >= 5 && warpU < 8) || (XL && warpU >= 12)) {
    // =========================== stagers ===========================
    const int st = warpU < 8 ? tid - kStager0 : kStagers + tid - 384;
    const int sw = st >> 5;                      // stager warp 0 .. NSTG/32 - 1
    // 4-float piece j0/4 of codeword k of subspace s: fp32 codebook [S][K][d] or the pre-split one [S][d/4][K][4 words]
    auto piece = [&](int s, int j0, int k) -> const float* {
      return a.cbPre ? a.ctrd + ((static_cast<size_t>(s) * (a.d >> 2) + (j0 >> 2)) * K + k) * 4
                     : a.ctrd + (static_cast<size_t>(s) * K + k) * a.d + j0;
    };
    // chunk kc: codebook slices + index rows by cp.async (one group); the positions travel through registers
    auto fetchChunk = [&](int kc) {
      const int cbuf = kc % kCbBufs;
      if (a.mode == 0) {
        // per-chunk scalars: the two 4-channel halves, their subspaces and the offsets inside the codewords
        const int taps = a.ksz * a.ksz;
        const int chA = (kcBase + kc) * 8, chB = chA + 4;
        const bool okA = chA < a.Cg, okB = chB < a.Cg;
        const int sA = okA ? chA / a.d : 0, jA = okA ? chA - sA * a.d : 0;
        const int sB = okB ? chB / a.d : 0, jB = okB ? chB - sB * a.d : 0;
        // (positions travel through registers: loadPos / storePos)
        // codebook slices: slot = half
        float4* cdst = cbs + cbuf * a.cbSlots * K;
        if (a.bulkC) {
          // four bulk copies (two contiguous codebook slices, two contiguous index blocks), completion counted on fullC;
          // a half beyond the group's channels stages valid data (its positions are staged as zeros)
          if (st == 0) {
            const uint32_t cbBytes = static_cast<uint32_t>(K) * 16u, idBytes = static_cast<uint32_t>(CTv * a.tapsPad);
            const uint32_t bar = SmemU32(fullC + cbuf);
            uint8_t* idst = ids + cbuf * a.idRows * 128;
            const uint8_t* tA = a.asmtT + (static_cast<size_t>(g * a.S + sA) * a.KgPad + ch0) * a.tapsPad;
            const uint8_t* tB = a.asmtT + (static_cast<size_t>(g * a.S + sB) * a.KgPad + ch0) * a.tapsPad;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2u * cbBytes + 2u * idBytes) : "memory");
#define QCNN_BULK(dst, src, bytes) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" \
                                                ::"r"(SmemU32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory")
            QCNN_BULK(cdst, piece(sA, jA, 0), cbBytes);
            QCNN_BULK(cdst + K, piece(sB, jB, 0), cbBytes);
            QCNN_BULK(idst, tA, idBytes);
            QCNN_BULK(idst + 128 * a.tapsPad, tB, idBytes);
#undef QCNN_BULK
          }
          return;
        }
        for (int k = st; k < K; k += NSTG) {
          CpAsync16(cdst + k, piece(sA, jA, k), okA);
          CpAsync16(cdst + K + k, piece(sB, jB, k), okB);
        }
        if (a.idxT) {
          // channel-major index block [half][channel][tapsPad]: contiguous CTv * tapsPad bytes per half
          const int gran = (CTv * a.tapsPad) >> 4;
          const uint8_t* tA = a.asmtT + (static_cast<size_t>(g * a.S + sA) * a.KgPad + ch0) * a.tapsPad;
          const uint8_t* tB = a.asmtT + (static_cast<size_t>(g * a.S + sB) * a.KgPad + ch0) * a.tapsPad;
          uint8_t* idst = ids + cbuf * a.idRows * 128;
          for (int e = st; e < 2 * gran; e += NSTG) {
            const bool hb = e >= gran;
            const int q = hb ? e - gran : e;
            CpAsync16(idst + (hb ? 128 * a.tapsPad : 0) + (q << 4), (hb ? tB : tA) + (static_cast<size_t>(q) << 4), true);
          }
        }
        // index rows [half][tap]: stager warp w takes rows w, w+3, ...; lane < gran copies one 16-byte granule
        const int gran = CTv >> 4;
        if (!a.idxT && lane < gran) {
          const uint8_t* asA = a.asmt + static_cast<size_t>(g * a.S + sA) * taps * a.KgPad + ch0 + (lane << 4);
          const uint8_t* asB = a.asmt + static_cast<size_t>(g * a.S + sB) * taps * a.KgPad + ch0 + (lane << 4);
          uint8_t* idst = ids + cbuf * a.idRows * 128 + (lane << 4);
          for (int row = sw; row < 2 * taps; row += NSTG / 32) {
            const bool hb = row >= taps;
            const int tap = hb ? row - taps : row;
            CpAsync16(idst + row * 128, (hb ? asB : asA) + static_cast<size_t>(tap) * a.KgPad, true);
          }
        }
      }
      if (a.mode == 1) {
        // chunk = phase row ph: pixels of input rows r*stride + ph - pad, one 4-byte copy per channel
        const int ph = kcBase + kc;
        // (positions travel through registers: loadPos / storePos)
        // codebook: slot 0 = the first 4 floats of every codeword of subspace 0, slot 1 = zeros (unpaired taps)
        float4* cdst = cbs + cbuf * a.cbSlots * K;
        if (a.bulkC) {
          if (st == 0) {
            const uint32_t cbBytes = static_cast<uint32_t>(K) * 16u, idBytes = static_cast<uint32_t>(CTv * a.tapsPad);
            const uint32_t bar = SmemU32(fullC + cbuf);
            const uint8_t* tA = a.asmtT + (static_cast<size_t>(g * a.stride + ph) * a.KgPad + ch0) * a.tapsPad;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cbBytes + idBytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(SmemU32(cdst)), "l"(piece(0, 0, 0)), "r"(cbBytes), "r"(bar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(SmemU32(ids + cbuf * a.idRows * 128)), "l"(tA), "r"(idBytes), "r"(bar) : "memory");
          }
          return;
        }
        for (int k = st; k < K; k += NSTG) {
          CpAsync16(cdst + k, piece(0, 0, k), true);
          CpAsync16(cdst + K + k, a.ctrd, false);
        }
        // index rows of the taps with kh % stride == ph, in (kh, kw) order
        const int gran = CTv >> 4;
        if (lane < gran) {
          const uint8_t* as0 = a.asmt + static_cast<size_t>(g * a.S) * a.ksz * a.ksz * a.KgPad + ch0 + (lane << 4);
          uint8_t* idst = ids + cbuf * a.idRows * 128 + (lane << 4);
          const int nrow = ((a.ksz - ph + a.stride - 1) / a.stride) * a.ksz;
          for (int row = sw; row < nrow; row += NSTG / 32) {
            const int kh = ph + (row / a.ksz) * a.stride, kw = row % a.ksz;
            CpAsync16(idst + row * 128, as0 + static_cast<size_t>(kh * a.ksz + kw) * a.KgPad, true);
          }
        }
      }
      if (a.mode == 2) {
        // chunk = KS consecutive k-steps of this CTA's range: features [f0, f0 + 8 ne)
        const int KS = a.chunkCount[0];
        const int ne = min(KS, kTotal - kc * KS);
        const int f0 = (k0 + kc * KS) * 8;
        const int gran = (CTv + 15) >> 4;
        if (a.d == 1) {
          // slot q = subspace f0 + q: its K scalar codewords; index row q
          const int kq = K >> 2;
          float4* cdst = cbs + cbuf * a.cbSlots * a.cbF4;
          for (int e = st; e < 8 * ne * kq; e += NSTG) CpAsync16(cdst + e, a.ctrd + static_cast<size_t>(f0) * K + e * 4, true);
          if (lane < gran) {
            const uint8_t* as0 = a.asmt + static_cast<size_t>(f0) * a.KgPad + ch0 + (lane << 4);
            uint8_t* idst = ids + cbuf * a.idRows * 128 + (lane << 4);
            for (int row = sw; row < 8 * ne; row += NSTG / 32) CpAsync16(idst + row * 128, as0 + static_cast<size_t>(row) * a.KgPad, true);
          }
        } else {
          // slot q = features [f0 + 4q, +4): piece j0 of every codeword of subspace s; index row q = row s of the layer
          float4* cdst = cbs + cbuf * a.cbSlots * a.cbF4;
          for (int e = st; e < 2 * ne * K; e += NSTG) {
            const int q = e / K, k = e - q * K;
            const int f = f0 + 4 * q;
            const int s = f / a.d, j0 = f - s * a.d;
            CpAsync16(cdst + e, piece(s, j0, k), true);
          }
          if (lane < gran) {
            uint8_t* idst = ids + cbuf * a.idRows * 128 + (lane << 4);
            for (int row = sw; row < 2 * ne; row += NSTG / 32) {
              const int s = (f0 + 4 * row) / a.d;
              CpAsync16(idst + row * 128, a.asmt + static_cast<size_t>(s) * a.KgPad + ch0 + (lane << 4), true);
            }
          }
        }
      }
      CpAsyncCommit();
    };
    // Positions of the convolution modes: global -> registers -> hi/lo planes.  (16-byte cp.async of scattered pieces
    // costs one shared-memory wavefront per THREAD -- profiles/README.md, "staging" -- whereas a 128-bit store of 32
    // consecutive float4 costs four per warp.)  A thread owns float4 st, st+96, ...: at most kRegPos of them.
    float4 rg[RP];
    int poffR[RP];        // chunk-invariant source offset of the thread's float4 (mode 0: incl. the half's +4)
    int prowR[RP];        // mode 1: first input row (phase row 0), very negative when the column is outside
    uint32_t pvalid = 0;       // mode 0: bit i = the float4 exists and its position is inside an image
    uint32_t phalf = 0;        // mode 0: bit i = second half (channels 4..7 of the chunk)
    if (a.mode != 2) {
#pragma unroll
      for (int i = 0; i < RP; i++) {
        const int p = st + i * NSTG;
        poffR[i] = 0; prowR[i] = -(1 << 28);
        if (p < a.planeF4) {
          const int off = posoff[p];
          if (a.mode == 0) {
            const bool hb = p >= a.NPOS;
            poffR[i] = off + (hb ? 4 : 0);
            if (off >= 0) pvalid |= 1u << i;
            if (hb) phalf |= 1u << i;
          } else {
            poffR[i] = off;
            prowR[i] = posrow[p];
          }
        }
      }
    }
    auto loadPos = [&](int kc) {
      if (a.mode == 0) {
        const int chA = (kcBase + kc) * 8;
        uint32_t m = pvalid;
        if (chA >= a.Cg) m = 0;
        else if (chA + 4 >= a.Cg) m &= ~phalf;
        const float* srcG = srcBase + g * a.Cg + chA;
#pragma unroll
        for (int i = 0; i < RP; i++) {
          rg[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if ((m >> i) & 1u) rg[i] = __ldg(reinterpret_cast<const float4*>(srcG + poffR[i]));
        }
      } else {
        const int ph = kcBase + kc;
        const float* srcG = srcBase + static_cast<size_t>(g) * a.Cg * a.chStride + ph * a.rowStride;
#pragma unroll
        for (int i = 0; i < RP; i++) {
          rg[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (static_cast<unsigned>(prowR[i] + ph) < static_cast<unsigned>(a.Hi)) {
            const float* px = srcG + poffR[i];
            rg[i].x = __ldg(px);
            if (a.Cg > 1) rg[i].y = __ldg(px + a.chStride);
            if (a.Cg > 2) rg[i].z = __ldg(px + 2 * a.chStride);
            if (a.Cg > 3) rg[i].w = __ldg(px + 3 * a.chStride);
          }
        }
      }
    };
    long long sCp = 0, sEB = 0, sEC = 0, sT0 = (DBG ? clock64() : 0ll);
    // Twelve-warp tiles of 3x3 layers (<= 6 float4 per stager thread, chunks of 9 k-steps): TWO register sets, the loads of
    // chunk kc+2 are issued right after the stores of chunk kc, so a global-load latency (~1-1.5 k clk under load) overlaps
    // a whole chunk instead of being waited for in every iteration -- with one set the stagers, not the MMAs, paced these
    // layers (role counters: 3.8 k clk per chunk against 2.4 k clk of MMAs).
    constexpr int H = 6;
    const bool dbl = WIDE && a.mode == 0 && a.planeF4 <= H * NSTG;
    if (WIDE && dbl) {
      float4 rg2[WIDE ? H : 1];
      int pdst[WIDE ? H : 1];           // bf16x2: byte offset of the element inside a plane (chunk-invariant)
#pragma unroll
      for (int i = 0; i < H; i++) {
        const int p = st + i * NSTG;
        const int grp = p >= a.NPOS ? 1 : 0;
        pdst[i] = (p - grp * a.NPOS) * 16 + grp * 8;
      }
      auto loadH = [&](int kc, auto& r) {
        const int chA = (kcBase + kc) * 8;
        uint32_t m = pvalid;
        if (chA >= a.Cg) m = 0;
        else if (chA + 4 >= a.Cg) m &= ~phalf;
        const float* srcG = srcBase + g * a.Cg + chA;
#pragma unroll
        for (int i = 0; i < H; i++) {
          r[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if ((m >> i) & 1u) r[i] = __ldg(reinterpret_cast<const float4*>(srcG + poffR[i]));
        }
      };
      auto iter = [&](int kc, auto& r) {
        const int buf = kc % a.nPB;
        long long c0 = (DBG ? clock64() : 0ll);
        if (!a.bulkC) {
          if (kc + 1 < nChunks) asm volatile("cp.async.wait_group 1;" ::: "memory");
          else CpAsyncWaitAll();
          sCp += (DBG ? clock64() : 0ll) - c0;
          MbarArrive(fullC + kc % kCbBufs);
        }
        c0 = (DBG ? clock64() : 0ll);
        if (kc >= a.nPB) MbarWait(emptyB + buf, ((kc / a.nPB) - 1) & 1);
        sEB += (DBG ? clock64() : 0ll) - c0;
        float4* pHi = planes + (buf * 2 + 0) * a.planeRows;
        float4* pLo = planes + (buf * 2 + 1) * a.planeRows;
#pragma unroll
        for (int i = 0; i < H; i++) {
          const int p = st + i * NSTG;
          if (p < a.planeF4) {
            if (BF) {
              uint2 p1, p2;
              SplitBf16x4(r[i], p1, p2);
              *reinterpret_cast<uint2*>(reinterpret_cast<char*>(pHi) + pdst[i]) = p1;
              *reinterpret_cast<uint2*>(reinterpret_cast<char*>(pLo) + pdst[i]) = p2;
            } else {
              float4 hi, lo;
              SplitTf32x4(r[i], hi, lo);
              pHi[p] = hi;
              pLo[p] = lo;
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        MbarArrive(fullB + buf);
        if (kc + 2 < nChunks) {
          loadH(kc + 2, r);
          const int nb = (kc + 2) % kCbBufs;
          c0 = (DBG ? clock64() : 0ll);
          if (kc + 2 >= kCbBufs) MbarWait(emptyC + nb, (((kc + 2) / kCbBufs) - 1) & 1);
          sEC += (DBG ? clock64() : 0ll) - c0;
          fetchChunk(kc + 2);
        }
      };
      fetchChunk(0);
      loadH(0, rg);
      if (nChunks > 1) { fetchChunk(1); loadH(1, rg2); }
      for (int kc = 0; kc < nChunks; kc += 2) {
        iter(kc, rg);
        if (kc + 1 < nChunks) iter(kc + 1, rg2);
      }
    } else {
    fetchChunk(0);
    if (a.mode != 2) loadPos(0);
    if (nChunks > 1) fetchChunk(1);
    for (int kc = 0; kc < nChunks; kc++) {
      const int buf = kc % a.nPB;
      long long c0 = (DBG ? clock64() : 0ll);
      if (!a.bulkC) {
        if (kc + 1 < nChunks) asm volatile("cp.async.wait_group 1;" ::: "memory");   // chunk kc landed, kc+1 may be in flight
        else CpAsyncWaitAll();
        // (each stager arrives after its own copies have landed; the barrier completes when all 96 have)
        sCp += (DBG ? clock64() : 0ll) - c0;
        MbarArrive(fullC + kc % kCbBufs);                  // the decoders may start on chunk kc
      }
      c0 = (DBG ? clock64() : 0ll);
      if (kc >= a.nPB) MbarWait(emptyB + buf, ((kc / a.nPB) - 1) & 1);   // planes last read by the MMAs of chunk kc-nPB
      sEB += (DBG ? clock64() : 0ll) - c0;
      float4* pHi = planes + (buf * 2 + 0) * a.planeRows;
      float4* pLo = planes + (buf * 2 + 1) * a.planeRows;
      if (a.mode == 2) {
        // the plane image of (tile, chunk) was written by fc_prep_kernel: one bulk copy (hi + lo, contiguous) straight
        // into the planes, completion counted on the same mbarrier the stagers arrive on
        if (st == 0) {
          const uint32_t bytes = static_cast<uint32_t>(a.planeRows) * 32u;
          const float* gsrc = a.xprep + (static_cast<size_t>(tile) * a.nChunksAll + (k0 / a.chunkCount[0] + kc)) * a.planeRows * 8;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemU32(fullB + buf)), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(SmemU32(pHi)), "l"(gsrc), "r"(bytes), "r"(SmemU32(fullB + buf)) : "memory");
        } else {
          MbarArrive(fullB + buf);
        }
      } else {
#pragma unroll
        for (int i = 0; i < RP; i++) {
          const int p = st + i * NSTG;
          if (p < a.planeF4) {
            if (BF) {
              uint2 p1, p2;
              SplitBf16x4(rg[i], p1, p2);
              const int db = posdst[p];
              *reinterpret_cast<uint2*>(reinterpret_cast<char*>(pHi) + db) = p1;
              *reinterpret_cast<uint2*>(reinterpret_cast<char*>(pLo) + db) = p2;
            } else {
              float4 hi, lo;
              SplitTf32x4(rg[i], hi, lo);
              pHi[p] = hi;
              pLo[p] = lo;
            }
          }
        }
      }
      if (a.mode != 2) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // planes are read by the tensor core
        MbarArrive(fullB + buf);
        if (kc + 1 < nChunks) loadPos(kc + 1);
      }
      if (kc + 2 < nChunks) {
        const int nb = (kc + 2) % kCbBufs;
        c0 = (DBG ? clock64() : 0ll);
        if (kc + 2 >= kCbBufs) MbarWait(emptyC + nb, (((kc + 2) / kCbBufs) - 1) & 1);
        sEC += (DBG ? clock64() : 0ll) - c0;
        fetchChunk(kc + 2);
      }
    }
    }
    if (DBG && a.dbg && st == 0) {
      atomicAdd(a.dbg + 8, static_cast<unsigned long long>((DBG ? clock64() : 0ll) - sT0));
      atomicAdd(a.dbg + 9, static_cast<unsigned long long>(sCp));
      atomicAdd(a.dbg + 10, static_cast<unsigned long long>(sEB));
      atomicAdd(a.dbg + 11, static_cast<unsigned long long>(sEC));
    }
  } else if (warpU == 4) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the loop on warp-uniform values (k-step table read from the kernel parameters, i.e. the
    // constant bank, so descriptors stay in uniform registers); one elected lane issues the tcgen05 instructions.
    // (Issued from a single divergent thread, every UTCHMMA is wrapped in an ELECT / BRA.U.ANY loop with R2UR moves; in
    //  this form the SASS issues them back to back -- profiles/README.md, "MMA issue".)
    {
      // instruction descriptor: D = F32, A = B = TF32, K-major, N = NT, M = 128
      // (bf16x2: A = B = BF16, format code 1; same N / M fields)
      const uint32_t idesc = BF ? ((1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(NT >> 3) << 17) | (8u << 24))
                                : ((1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(NT >> 3) << 17) | (8u << 24));
      const uint64_t descFixed = (static_cast<uint64_t>(8) << 32) | (static_cast<uint64_t>(1) << 46);  // SBO = 128 B
      const uint32_t planes0 = SmemU32(planes);
      int t = 0, islot = 0, iround = 0;
      uint32_t acc = 0;
      // NT <= 128 leaves room for a second accumulator: the hi*hi products go to D, the two cross terms to D + NT, and
      // the epilogue adds them -- the tensor core's accumulation error grows with the number of chained MMAs per
      // accumulator, and the chain of the dominant term is three times shorter this way
      const uint32_t corrOff = a.corr ? static_cast<uint32_t>(NT) : 0u;
      uint32_t accCorr = 0;
      long long wBC = 0, wA = 0, tStart = (DBG ? clock64() : 0ll);
      for (int kc = 0; kc < nChunks; kc++) {
        const int buf = kc % a.nPB;
        long long c0 = (DBG ? clock64() : 0ll);
        MbarWait(fullB + buf, (kc / a.nPB) & 1);
        wBC += (DBG ? clock64() : 0ll) - c0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dHi = descFixed | (((planes0 + static_cast<uint32_t>((buf * 2 + 0) * a.planeRows) * 16u) >> 4) & 0x3FFFu);
        const uint64_t dLo = descFixed | (((planes0 + static_cast<uint32_t>((buf * 2 + 1) * a.planeRows) * 16u) >> 4) & 0x3FFFu);
        const int e0 = a.chunkFirst[a.mode == 1 ? kcBase + kc : 0];
        const int ne = a.mode == 2 ? min(a.chunkCount[0], kTotal - kc * a.chunkCount[0]) : a.chunkCount[a.mode == 1 ? kcBase + kc : 0];
        for (int s0 = 0; s0 < ne; s0 += GT, t++) {
          const int slot = islot;                         // = t % NSLOT, ring round = t / NSLOT (kept incrementally)
          c0 = (DBG ? clock64() : 0ll);
          MbarWait(fullA + slot, iround & 1);
          if (++islot == NSLOT) { islot = 0; iround++; }
          wA += (DBG ? clock64() : 0ll) - c0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int n = min(GT, ne - s0);
          if (ElectOne()) {
            for (int i = 0; i < n; i++) {
              const KStep ks = a.tab[e0 + s0 + i];
              const uint64_t off = static_cast<uint64_t>(static_cast<uint32_t>(ks.bStart)) |
                                   (static_cast<uint64_t>(static_cast<uint32_t>(ks.lbo)) << 16);
              const uint32_t aHi = tmemA + static_cast<uint32_t>((slot * GT + i) * 16), aLo = aHi + 8;
              if (BF) {
                // B = [x1 | x2]: K-core-matrix 0 in the first plane, 1 in the second (the table's lbo is the plane distance);
                // A = [w1 | w1] then [w2 | w2] (columns 0-7 / 8-15 of the slot)
                UmmaF16Ts(tmemD, aHi, dHi + off, idesc, acc);
                UmmaF16Ts(tmemD + corrOff, aLo, dHi + off, idesc, corrOff ? accCorr : 1u);
              } else {
                UmmaTf32Ts(tmemD, aHi, dHi + off, idesc, acc);
                UmmaTf32Ts(tmemD + corrOff, aHi, dLo + off, idesc, corrOff ? accCorr : 1u);
                UmmaTf32Ts(tmemD + corrOff, aLo, dHi + off, idesc, 1u);
              }
              acc = 1u;
              accCorr = 1u;
            }
            // commits are issued by the thread that issued the MMAs they track
            UmmaCommit(emptyA + slot);
            if (s0 + GT >= ne) {
              UmmaCommit(emptyB + buf);
              if (kc == nChunks - 1) UmmaCommit(doneBar);
            }
          }
          acc = 1u;
          accCorr = 1u;
          __syncwarp();
        }
      }
      if (DBG && a.dbg && lane == 0) {
        const long long tIssue = (DBG ? clock64() : 0ll) - tStart;
        long long c0 = (DBG ? clock64() : 0ll);
        MbarWait(doneBar, 0);
        atomicAdd(a.dbg + 0, static_cast<unsigned long long>(tIssue));
        atomicAdd(a.dbg + 1, static_cast<unsigned long long>(wBC));
        atomicAdd(a.dbg + 2, static_cast<unsigned long long>(wA));
        atomicAdd(a.dbg + 3, static_cast<unsigned long long>((DBG ? clock64() : 0ll) - c0));
        atomicAdd(a.dbg + 5, 1ull);
      }
    }
  } else {
    // =========================== decoders ===========================
    const int c = (warp & 3) * 32 + lane;           // channel row = TMEM lane (a warp reaches the lanes of its sub-partition)
    const int cc = min(c, CTv - 1);                 // rows beyond the valid channels decode a copy (never stored)
    const uint32_t laneBase = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int grp = (TWOG && warp >= 8) ? 1 : 0;    // WIDE: group 0 decodes the even stages, group 1 the odd ones
    const bool pre = BF && a.cbPre != 0;
    int t = 0, dslot = 0, dround = 0;
    long long dP1 = 0, dP2 = 0, dP3 = 0, dP4 = 0, dP3s = 0;   // DBG: index loads / codeword loads / split + tcgen05.st issue / wait::st
    long long dFC = 0, dEA = 0, dT0 = (DBG ? clock64() : 0ll);
    for (int kc = 0; kc < nChunks; kc++) {
      const int cbuf = kc % kCbBufs;
      long long c0 = (DBG ? clock64() : 0ll);
      MbarWait(fullC + cbuf, (kc / kCbBufs) & 1);
      dFC += (DBG ? clock64() : 0ll) - c0;
      const uint8_t* idb = ids + cbuf * a.idRows * 128 + cc;
      const float4* cb = cbs + cbuf * a.cbSlots * a.cbF4;
      const float* cbf = reinterpret_cast<const float*>(cb);
      const int e0 = a.chunkFirst[a.mode == 1 ? kcBase + kc : 0];
      const int ne = a.mode == 2 ? min(a.chunkCount[0], kTotal - kc * a.chunkCount[0]) : a.chunkCount[a.mode == 1 ? kcBase + kc : 0];
      // idxT: every tap index of this channel for the chunk's two halves in registers (one or two 128-bit loads each): the
      // per-k-step chain loses two dependent shared-memory round trips (table entry, index byte) -- under the load of the
      // MMAs' operand fetch a round trip costs ~450 clk (role counters, profiles/README.md)
      uint4 iA0 = make_uint4(0, 0, 0, 0), iA1 = iA0, iB0 = iA0, iB1 = iA0;
      if (a.idxT) {
        const uint4* ip = reinterpret_cast<const uint4*>(ids + cbuf * a.idRows * 128);
        const int W = a.tapsPad >> 4;
        iA0 = ip[cc * W];
        if (a.mode == 1) {          // one block of <= 48 rows: iA0, iA1, iB0
          if (W > 1) iA1 = ip[cc * W + 1];
          if (W > 2) iB0 = ip[cc * W + 2];
        } else {
          iB0 = ip[(128 + cc) * W];
          if (W > 1) { iA1 = ip[cc * W + 1]; iB1 = ip[(128 + cc) * W + 1]; }
        }
      }
      for (int s0 = 0; s0 < ne; s0 += GT, t++) {
        const int slot = dslot, round = dround;           // = t % NSLOT, t / NSLOT (kept incrementally: no divisions)
        if (++dslot == NSLOT) { dslot = 0; dround++; }
        if (TWOG && (t & 1) != grp) continue;
        if (round > 0) {
          c0 = (DBG ? clock64() : 0ll);
          MbarWait(emptyA + slot, (round - 1) & 1);
          dEA += (DBG ? clock64() : 0ll) - c0;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const int n = min(GT, ne - s0);
        if (a.d == 1) {
          for (int i = 0; i < n; i++) {
            // scalar codewords: every feature is its own subspace (rows / slots idx0 .. idx0+3 and idx1 .. idx1+3)
            const KStep ks = tabS[e0 + s0 + i];
            const uint8_t* r0 = idb + ks.idx0 * 128;
            const uint8_t* r1 = idb + ks.idx1 * 128;
            const float* c0p = cbf + ks.cb0 * K;
            const float* c1p = cbf + ks.cb1 * K;
            float4 w0, w1;
            w0.x = c0p[r0[0] >> a.kshift];           w0.y = c0p[K + (r0[128] >> a.kshift)];
            w0.z = c0p[2 * K + (r0[256] >> a.kshift)]; w0.w = c0p[3 * K + (r0[384] >> a.kshift)];
            w1.x = c1p[r1[0] >> a.kshift];           w1.y = c1p[K + (r1[128] >> a.kshift)];
            w1.z = c1p[2 * K + (r1[256] >> a.kshift)]; w1.w = c1p[3 * K + (r1[384] >> a.kshift)];
            if (BF) StoreWeightsBf<false>(tmemA + laneBase + static_cast<uint32_t>((slot * GT + i) * 16), w0, w1);
            else StoreWeights(tmemA + laneBase + static_cast<uint32_t>((slot * GT + i) * 16), w0, w1);
          }
        } else {
          // the stage's k-steps are independent: table entries, index bytes and codeword pieces of all of them are in
          // flight together (one decoder warp per SM sub-partition: latency, not bandwidth, paces this role)
          int i0x[MG], i1x[MG];
          long long p0 = (DBG ? clock64() : 0ll);
          // (the operand-source choice is hoisted out of the unrolled loops: inside them it would serialise the loads)
          if (a.idxT && a.mode == 1) {     // rows and codebook slots from the k-step table (constant bank), bytes from registers
#pragma unroll
            for (int i = 0; i < MG; i++) {
              if (i < n) {
                const KStep ks = a.tab[e0 + s0 + i];
                i0x[i] = ks.cb0 * K + ByteOf48(iA0, iA1, iB0, ks.idx0);
                i1x[i] = ks.cb1 * K + ByteOf48(iA0, iA1, iB0, ks.idx1);
              }
            }
          } else if (a.idxT) {             // mode 0: k-step = tap s0 + i, codebook slots 0 / 1
#pragma unroll
            for (int i = 0; i < MG; i++) {
              if (i < n) {
                i0x[i] = ByteOf32(iA0, iA1, s0 + i);
                i1x[i] = K + ByteOf32(iB0, iB1, s0 + i);
              }
            }
          } else {                         // table entries and index bytes of all k-steps in flight together
            KStep ks[MG];
#pragma unroll
            for (int i = 0; i < MG; i++)
              if (i < n) ks[i] = tabS[e0 + s0 + i];
#pragma unroll
            for (int i = 0; i < MG; i++) {
              if (i < n) {
                i0x[i] = ks[i].cb0 * K + (idb[ks[i].idx0 * 128] >> a.kshift);
                i1x[i] = ks[i].cb1 * K + (idb[ks[i].idx1 * 128] >> a.kshift);
              }
            }
          }
          if (DBG) { asm volatile("" :: "r"(i0x[0]), "r"(i1x[0]) : "memory"); dP1 += clock64() - p0; p0 = clock64(); }
          float4 w0[MG], w1[MG];
#pragma unroll
          for (int i = 0; i < MG; i++) {
            if (i < n) { w0[i] = cb[i0x[i]]; w1[i] = cb[i1x[i]]; }
          }
          if (DBG) { asm volatile("" :: "f"(w0[0].x), "f"(w1[0].x) : "memory"); dP2 += clock64() - p0; dP3s = clock64(); }
#pragma unroll
          for (int i = 0; i < MG; i++) {
            if (i < n) {
              if (BF && pre) StoreWeightsBf<true>(tmemA + laneBase + static_cast<uint32_t>((slot * GT + i) * 16), w0[i], w1[i]);
              else if (BF) StoreWeightsBf<false>(tmemA + laneBase + static_cast<uint32_t>((slot * GT + i) * 16), w0[i], w1[i]);
              else StoreWeights(tmemA + laneBase + static_cast<uint32_t>((slot * GT + i) * 16), w0[i], w1[i]);
            }
          }
        }
        if (DBG) { dP3 += clock64() - dP3s; dP3s = clock64(); }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (DBG) { dP4 += clock64() - dP3s; }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        MbarArrive(fullA + slot);
      }
      MbarArrive(emptyC + cbuf);
    }
    if (DBG && a.dbg && tid == 0) {
      atomicAdd(a.dbg + 12, static_cast<unsigned long long>((DBG ? clock64() : 0ll) - dT0));
      atomicAdd(a.dbg + 13, static_cast<unsigned long long>(dFC));
      atomicAdd(a.dbg + 14, static_cast<unsigned long long>(dEA));
      atomicAdd(a.dbg + 4, static_cast<unsigned long long>(dP1));
      atomicAdd(a.dbg + 6, static_cast<unsigned long long>(dP2));
      atomicAdd(a.dbg + 7, static_cast<unsigned long long>(dP3));
      atomicAdd(a.dbg + 15, static_cast<unsigned long long>(dP4));
    }
  }

  // ---- epilogue: TMEM (lane = channel, column = position) -> + bias, ReLU -> NHWC global ----
  MbarWait(doneBar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int q = warp & 3, hsel = warp >> 2;     // (warps 4-7 and, WIDE, 8-11 take the other position blocks)
    const int c = q * 32 + lane;
    const bool chOk = c < CTv;
    const float bv = biasS[c];
    const int nblk = NT >> 4;   // 16-position blocks
    for (int blk = hsel; blk < nblk; blk += NTHR / 128) {
      uint32_t r[16];
      const uint32_t taddr = tmemD + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(blk * 16);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                     "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (a.corr) {   // second accumulator (cross terms of the 3xTF32 split)
        uint32_t r2[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7]),
                       "=r"(r2[8]), "=r"(r2[9]), "=r"(r2[10]), "=r"(r2[11]), "=r"(r2[12]), "=r"(r2[13]), "=r"(r2[14]), "=r"(r2[15])
                     : "r"(taddr + static_cast<uint32_t>(NT)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      }
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const int off = outoff[blk * 16 + j];
        if (off >= 0 && chOk) {
          float v = __uint_as_float(r[j]) + bv;
          if (a.relu) v = fmaxf(v, 0.0f);   // (never set together with split-K partial sums)
          dstBase[off + c] = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemD), "r"(kTmemCols) : "memory");
}

}  // namespace

namespace qcnn {

// Candidate tilings for batch N (cost in SM-cycles, comparable with PlanConv's model): 3 MMAs of NT/2 clk per k-step.
void AddLitePqGemm(const qcnn_layer* L, std::vector<std::pair<double, ConvPlan>>* cands, size_t first);
void AddWidePqGemm(const qcnn_layer* L, std::vector<std::pair<double, ConvPlan>>* cands, size_t first);
void AddXlPqGemm(const qcnn_layer* L, std::vector<std::pair<double, ConvPlan>>* cands, size_t first);
static void PlanPqGemmFull(const qcnn_layer* L, int N, std::vector<std::pair<double, ConvPlan>>* cands);
void PlanPqGemm(const qcnn_layer* L, int N, std::vector<std::pair<double, ConvPlan>>* cands) {
  const size_t first = cands->size();
  PlanPqGemmFull(L, N, cands);
  static const bool lite = !(getenv("QCNN_GEMM_LITE") && getenv("QCNN_GEMM_LITE")[0] == '0');
  static const bool wide = !(getenv("QCNN_GEMM_WIDE") && getenv("QCNN_GEMM_WIDE")[0] == '0');
  const size_t mid = cands->size();
  static const bool xl = !(getenv("QCNN_GEMM_XL") && getenv("QCNN_GEMM_XL")[0] == '0');
  if (xl) AddXlPqGemm(L, cands, first);
  if (wide) AddWidePqGemm(L, cands, first);
  if (lite) {
    const size_t end = cands->size();
    AddLitePqGemm(L, cands, first);
    (void)mid; (void)end;
  }
}
static void PlanPqGemmFull(const qcnn_layer* L, int N, std::vector<std::pair<double, ConvPlan>>* cands) {
  const int G = L->grp, Cg = L->Cin / G, Kg = L->Cout / G, taps = L->ksz * L->ksz;
  if (L->K < 1 || L->K > 256 || Kg % 16 != 0) return;
  const size_t smemMax0 = L->ctx->smem_optin ? L->ctx->smem_optin : 227 * 1024;
  if (L->stride > 1) {
    // mode 1: strided convolution with one subspace over <= 4 input channels (conv1), input on stride x stride phase planes
    const int st = L->stride, rowsPer = CeilDiv(L->ksz, st);
    if (Cg > 4 || L->S != 1 || L->d < Cg || st > 8 || L->Wo > CeilDiv(L->Win + 2 * L->pad, st)) return;
    const int pairs = (L->ksz + 1) / 2;
    if (rowsPer * pairs > 32 || L->ksz * pairs > kMaxKSteps) return;
    const int PWp = CeilDiv(L->Win + 2 * L->pad, st), PHp = CeilDiv(L->Hin + 2 * L->pad, st);
    const int IB = PHp * PWp;
    if (static_cast<double>(N) * IB > 2.0e9 || static_cast<double>(L->Hin) * L->Win * L->Cin * 4 > 2.0e9) return;
    const int nts1[2] = {256, 128};
    for (int ni = 0; ni < 2; ni++) {
      const int NT = nts1[ni];
      const int gts1[3] = {6, 4, 8};
      for (int gi = 0; gi < 3; gi++) {
        const int GT = gts1[gi];
        ConvPlan p;
        memset(&p, 0, sizeof(p));
        p.kernel = 6; p.CPT = NT; p.J = GT; p.threads = kThreads;
        GemmArgs& ga = p.g;
        ga.mode = 1;
        ga.PW = PWp; ga.IB = IB; ga.NT = NT; ga.GT = GT; ga.NSLOT = std::min(kMaxSlots, 16 / GT);
        ga.NPOS = RoundUp(NT + ((L->ksz - 1) / st) * (PWp + 1), 8);
        ga.planeF4 = st * ga.NPOS;
        if (ga.planeF4 > kRegPos * kStagers) continue;
        ga.aOff = 256; ga.corr = NT <= 128 ? 1 : 0; ga.lite = 0;
        ga.bf = L->opt_tc_bf ? 1 : 0;
        if (ga.bf && (st & 1)) continue;       // bf16x2 rows hold a PAIR of phase columns
        ga.planeRows = ga.bf ? (st / 2) * ga.NPOS : ga.planeF4;
        ga.cbSlots = 2; ga.idRows = rowsPer * L->ksz;
        ga.tapsPad = RoundUp(rowsPer * L->ksz, 16);
        ga.idxT = (ga.tapsPad <= 48 && L->d_asmt_t && L->asmt_t_mode == 1) ? 1 : 0;
        if (ga.idxT) ga.idRows = ga.tapsPad;       // [128 channels][tapsPad] bytes
        ga.nChunks = st;
        ga.K = L->K; ga.cbF4 = L->K; ga.nPB = 2;
        // bf16x2 planes are half the size: keep EVERY phase row of the tile resident (no plane reuse), so that the stagers
        // never wait for the MMAs and the few k-steps of a phase row (conv1: 16-18) are not what hides a global-load latency
        if (ga.bf && st <= 4) ga.nPB = st;
        ga.nct = CeilDiv(Kg, 128);
        int ne = 0;
        for (int ph = 0; ph < st; ph++) {
          ga.chunkFirst[ph] = ne;
          for (int kh = ph; kh < L->ksz; kh += st)
            for (int kw = 0; kw < L->ksz; kw += 2, ne++) {
              if (ne >= kMaxKSteps) { ne = kMaxKSteps + 1000; break; }
              KStep& ks = ga.tab[ne];
              // partner in the next phase column, same shift (bf16x2: the partner must be the odd column of the same pair)
              const bool paired = kw + 1 < L->ksz && (kw % st) + 1 < st && !(ga.bf && ((kw % st) & 1));
              ks.bStart = (kw % st) * ga.NPOS + (kh / st) * PWp + kw / st;
              ks.lbo = paired ? ga.NPOS : 1;
              if (ga.bf) {   // row = (pair of phase columns, position); K-core-matrix 1 = the x2 plane
                ks.bStart = ((kw % st) >> 1) * ga.NPOS + (kh / st) * PWp + kw / st;
                ks.lbo = ga.planeRows;
              }
              ks.idx0 = static_cast<short>((kh - ph) / st * L->ksz + kw);
              ks.idx1 = static_cast<short>(paired ? ks.idx0 + 1 : ks.idx0);
              ks.cb0 = 0; ks.cb1 = paired ? 0 : 1;
              if (!paired && kw + 1 < L->ksz) kw--;   // the partner starts its own k-step
            }
          ga.chunkCount[ph] = ne - ga.chunkFirst[ph];
        }
        if (ne > kMaxKSteps) continue;
        ga.ntab = ne;
        p.smem = static_cast<size_t>(MapSmem(ga).total);
        if (p.smem > smemMax0) continue;
        p.a.CT = 128; p.a.nct = ga.nct; p.a.R = NT; p.a.nstrips = GT;
        const double ctas0 = static_cast<double>(CeilDiv(N * IB, NT)) * G * ga.nct;
        // small batches: split the phase rows over CTAs (partial sums reduced by LaunchSplitReduce)
        for (int nsplit = 1; nsplit <= st; nsplit *= 2) {
          if (nsplit > 1 && ctas0 * nsplit > 1.5 * L->ctx->sm_count) break;
          ga.kPerSplit = CeilDiv(st, nsplit);
          ga.nsplit = CeilDiv(st, ga.kPerSplit);
          if (ga.nsplit != nsplit) continue;
          p.a.rgroups = nsplit;
          const double perCta = ne * 3.0 * (NT / 2.0) * 1.15 / nsplit + 120.0 * ne / GT / nsplit + 9000.0 + NT * 24.0 +
                                (nsplit > 1 ? 6000.0 : 0.0);
          cands->emplace_back(perCta * std::ceil(ctas0 * nsplit / L->ctx->sm_count), p);
        }
      }
    }
    return;
  }
  if (L->src_nchw) return;
  if (L->d % 4 != 0 || Cg % 4 != 0 || L->Cin % 4 != 0 || L->S * L->d < Cg || taps > kMaxKSteps) return;
  const int PW = L->Win + L->pad, IB = (L->Hin + L->pad) * PW;
  if (L->Ho > L->Hin + L->pad || L->Wo > PW) return;
  if (static_cast<double>(N) * IB > 2.0e9) return;
  const size_t smemMax = L->ctx->smem_optin ? L->ctx->smem_optin : 227 * 1024;
  const int nts[3] = {256, 128, 64};
  for (int ni = 0; ni < 3; ni++) {
    const int NT = nts[ni];
    const int gts[3] = {4, 5, 8};
    for (int gi = 0; gi < 3; gi++) {
      const int GT = gts[gi];
      ConvPlan p;
      memset(&p, 0, sizeof(p));
      p.kernel = 6; p.CPT = NT; p.J = GT; p.threads = kThreads;
      GemmArgs& ga = p.g;
      ga.mode = 0;
      ga.PW = PW; ga.IB = IB; ga.NT = NT; ga.GT = GT; ga.NSLOT = std::min(kMaxSlots, 16 / GT);
      ga.NPOS = RoundUp(NT + (L->ksz - 1) * (PW + 1), 8);
      if (ga.NPOS > 16000) continue;
      ga.planeF4 = 2 * ga.NPOS;
      if (ga.planeF4 > kRegPos * kStagers) continue;
      ga.aOff = 256; ga.corr = NT <= 128 ? 1 : 0; ga.lite = 0;
      ga.bf = L->opt_tc_bf ? 1 : 0;
      ga.planeRows = ga.bf ? ga.NPOS : ga.planeF4;   // (bf16x2: lbo = NPOS rows is the x1 -> x2 plane distance as well)
      ga.cbSlots = 2; ga.idRows = 2 * taps;
      ga.tapsPad = RoundUp(taps, 16);
      ga.idxT = (ga.tapsPad <= 32 && L->d_asmt_t && L->asmt_t_mode == 0) ? 1 : 0;
      if (ga.idxT) ga.idRows = 2 * ga.tapsPad;      // [half][128 channels][tapsPad] bytes = 2 * tapsPad rows of 128
      ga.nChunks = CeilDiv(Cg, 8);
      ga.ntab = taps;
      ga.chunkFirst[0] = 0; ga.chunkCount[0] = taps;
      for (int t = 0; t < taps; t++) {
        KStep& ks = ga.tab[t];
        ks.bStart = (t / L->ksz) * PW + (t % L->ksz);
        ks.lbo = ga.NPOS;
        ks.idx0 = static_cast<short>(t); ks.idx1 = static_cast<short>(taps + t);
        ks.cb0 = 0; ks.cb1 = 1;
      }
      ga.K = L->K; ga.cbF4 = L->K; ga.nPB = 2;
      ga.nct = CeilDiv(Kg, 128);
      p.smem = static_cast<size_t>(MapSmem(ga).total);
      if (p.smem > smemMax) continue;
      p.a.CT = 128; p.a.nct = ga.nct; p.a.R = NT; p.a.nstrips = GT;   // (candidate de-duplication keys)
      const double ksteps = static_cast<double>(ga.nChunks) * taps;
      const double ctas0 = static_cast<double>(CeilDiv(N * IB, NT)) * G * ga.nct;
      // small batches: split the 8-channel chunks over CTAs (partial sums reduced by LaunchSplitReduce)
      for (int nsplit = 1; nsplit <= ga.nChunks; nsplit *= 2) {
        if (nsplit > 1 && ctas0 * nsplit > 1.5 * L->ctx->sm_count) break;
        ga.kPerSplit = CeilDiv(ga.nChunks, nsplit);
        ga.nsplit = CeilDiv(ga.nChunks, ga.kPerSplit);
        if (ga.nsplit != nsplit) continue;
        p.a.rgroups = nsplit;
        const double perCta = ksteps * 3.0 * (NT / 2.0) * 1.1 / nsplit + 120.0 * ksteps / GT / nsplit + 4000.0 + NT * 24.0 +
                              (nsplit > 1 ? 6000.0 : 0.0);
        cands->emplace_back(perCta * std::ceil(ctas0 * nsplit / L->ctx->sm_count), p);
      }
    }
  }
}

// Two-CTAs-per-SM variants of the candidates above (see kRegPosLite): tiles of <= 128 positions whose register windows,
// TMEM columns (accumulators + weight ring <= 256) and shared memory (two CTAs + their 1 KB reservations per SM) fit
// twice.  The fixed part of the cost model is halved: it overlaps the neighbour CTA's MMA phase.
// Twelve-warp variants (second decoder group): stages of <= 5 k-steps, register windows of the lite size.
void AddWidePqGemm(const qcnn_layer* L, std::vector<std::pair<double, ConvPlan>>* cands, size_t first) {
  (void)L;
  const size_t n = cands->size();
  for (size_t i = first; i < n; i++) {
    ConvPlan p = (*cands)[i].second;
    GemmArgs& ga = p.g;
    if (p.kernel != 6 || ga.lite || ga.xl || ga.GT > 5 || ga.planeF4 > kRegPosLite * kStagers || ga.NSLOT < 2) continue;
    ga.wide = 1;
    p.threads = 384;
    p.J = ga.GT + 200;     // (candidate de-duplication key)
    cands->emplace_back((*cands)[i].first * 0.85, p);
  }
}

// Sixteen-warp variants for mode 1 (conv1): tiles of 256 positions, stages of <= 4 k-steps, 224 stager threads
void AddXlPqGemm(const qcnn_layer* L, std::vector<std::pair<double, ConvPlan>>* cands, size_t first) {
  (void)L;
  const size_t n = cands->size();
  for (size_t i = first; i < n; i++) {
    ConvPlan p = (*cands)[i].second;
    GemmArgs& ga = p.g;
    if (p.kernel != 6 || ga.lite || ga.wide || ga.mode != 1 || !ga.bf || ga.GT > kMaxGTLite || ga.NSLOT < 2 ||
        ga.planeF4 > 7 * (kStagers + 128)) continue;
    ga.xl = 1;
    p.threads = 512;
    p.J = ga.GT + 300;     // (candidate de-duplication key)
    cands->emplace_back((*cands)[i].first * 0.7, p);
  }
}

void AddLitePqGemm(const qcnn_layer* L, std::vector<std::pair<double, ConvPlan>>* cands, size_t first) {
  const size_t smemMax = L->ctx->smem_optin ? L->ctx->smem_optin : 227 * 1024;
  const size_t perCtaSmem = (smemMax + 1024) / 2 - 1024 - 512;
  const size_t n = cands->size();
  for (size_t i = first; i < n; i++) {
    ConvPlan p = (*cands)[i].second;
    GemmArgs& ga = p.g;
    if (p.kernel != 6 || ga.wide || ga.xl || ga.NT > 128 || ga.GT > kMaxGTLite || ga.planeF4 > kRegPosLite * kStagers || p.smem > perCtaSmem) continue;
    const int ring = ga.NSLOT * ga.GT * 16;
    ga.NSLOT = std::min(ga.NSLOT, (256 - ga.NT) / (ga.GT * 16));
    if (ga.NSLOT < 2) continue;
    (void)ring;
    ga.corr = (2 * ga.NT + ga.NSLOT * ga.GT * 16 <= 256) ? 1 : 0;
    ga.aOff = ga.corr ? 2 * ga.NT : ga.NT;
    ga.lite = 1;
    p.J = ga.GT + 100;     // (candidate de-duplication key: distinct from the one-CTA-per-SM twin)
    cands->emplace_back((*cands)[i].first * 0.8, p);
  }
}

// dynamic shared-memory limit of both instantiations: raised once per device to the opt-in maximum (not per launch,
// which would make the limit follow whichever layer ran last -- fragile under graph capture / concurrent streams)
static int SetSmemLimitOnce(qcnn_ctx* ctx) {
  static bool done[64] = {false};
  if (ctx->device >= 0 && ctx->device < 64 && done[ctx->device]) return 0;
  const int lim = static_cast<int>(ctx->smem_optin ? ctx->smem_optin : 227 * 1024);
#define QCNN_SMEM_ATTR(...) QCNN_CUDA(cudaFuncSetAttribute(pq_gemm_tc_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim))
  QCNN_SMEM_ATTR(false, false, false, false); QCNN_SMEM_ATTR(true, false, false, false); QCNN_SMEM_ATTR(false, true, false, false);
  QCNN_SMEM_ATTR(false, false, true, false);  QCNN_SMEM_ATTR(true, false, true, false);  QCNN_SMEM_ATTR(false, true, true, false);
  QCNN_SMEM_ATTR(false, false, false, true);  QCNN_SMEM_ATTR(false, false, true, true);  QCNN_SMEM_ATTR(true, false, true, true);
  QCNN_SMEM_ATTR(false, false, true, false, true);  QCNN_SMEM_ATTR(true, false, true, false, true);
#undef QCNN_SMEM_ATTR
  // the lite kernels' CTAs must really pair up: ask for the largest shared-memory carve-out
  QCNN_CUDA(cudaFuncSetAttribute(pq_gemm_tc_kernel<false, true, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  QCNN_CUDA(cudaFuncSetAttribute(pq_gemm_tc_kernel<false, true, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  if (ctx->device >= 0 && ctx->device < 64) done[ctx->device] = true;
  return 0;
}

// bf16x2 form of the codebook, made once per layer: every aligned 4-float piece becomes {w1 (4 x bf16), w2 (4 x bf16)},
// stored [S][d/4][K] so that the K pieces a chunk stages are one contiguous run (bulk-copyable)
int BuildCtrdBf(qcnn_layer* L) {
  if (L->d % 4 != 0 || L->d_ctrd_bf) return 0;
  const size_t n = static_cast<size_t>(L->S) * L->K * L->d / 4;
  QCNN_CUDA(cudaMalloc(&L->d_ctrd_bf, n * sizeof(uint4)));
  split_codebook_kernel<<<static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 1024)), 256>>>(
      reinterpret_cast<const float4*>(L->d_ctrd), reinterpret_cast<uint4*>(L->d_ctrd_bf), L->S, L->K, L->d / 4);
  QCNN_CUDA(cudaGetLastError());
  QCNN_CUDA(cudaDeviceSynchronize());
  return 0;
}

int LaunchPqGemm(qcnn_layer* L, const ConvPlan& p, const float* src, int N, float* dst, int relu, cudaStream_t st) {
  GemmArgs a = p.g;
  a.src = src; a.dst = dst; a.ctrd = L->d_ctrd; a.asmt = L->d_asmt; a.bias = L->d_bias;
  if (a.bf && L->d_ctrd_bf) { a.ctrd = reinterpret_cast<const float*>(L->d_ctrd_bf); a.cbPre = 1; }
  a.asmtT = L->d_asmt_t;
  a.bulkC = (a.mode != 2 && a.idxT && a.cbPre) ? 1 : 0;
  if (a.mode == 1 && !a.bulkC) a.idxT = 0;   // (3xTF32 operands: index rows by cp.async in tap-major order; the planned buffer is large enough)
  a.N = N; a.relu = a.nsplit > 1 ? 0 : relu;
  if (a.nsplit < 1) a.nsplit = 1;
  a.Hi = L->Hin; a.Wi = L->Win; a.Cin = L->Cin; a.Ho = L->Ho; a.Wo = L->Wo; a.Cout = L->Cout;
  a.ksz = L->ksz; a.pad = L->pad; a.stride = L->stride; a.G = L->grp; a.Cg = L->Cin / L->grp; a.Kg = L->Cout / L->grp;
  a.KgPad = RoundUp(a.Kg, 16); a.S = L->S; a.K = L->K; a.d = L->d;
  a.srcImg = static_cast<long long>(a.Hi) * a.Wi * a.Cin;
  if (L->src_nchw) { a.rowStride = a.Wi; a.colStride = 1; a.chStride = a.Hi * a.Wi; }
  else { a.rowStride = a.Wi * a.Cin; a.colStride = a.Cin; a.chStride = 1; }
  a.dstImg = static_cast<long long>(a.Ho) * a.Wo * a.Cout;
  const long long blocks = static_cast<long long>(CeilDiv(N * a.IB, a.NT)) * a.G * a.nct * a.nsplit;
  QCNN_CHECK(blocks <= 2147483647LL, "qcnn_conv_aprx_forward: batch too large for the tensor-core tiling");
  if (a.nsplit > 1) {
    const size_t need = sizeof(float) * static_cast<size_t>(a.nsplit) * N * a.dstImg;
    if (need > L->partial_bytes) {
      if (L->d_partial) QCNN_CUDA(cudaFree(L->d_partial));
      L->d_partial = nullptr; L->partial_bytes = 0;
      QCNN_CUDA(cudaMalloc(&L->d_partial, need));
      L->ctx->alloc_epoch++;
      L->partial_bytes = need;
    }
    a.partial = L->d_partial;
  }
  if (int rc = SetSmemLimitOnce(L->ctx)) return rc;
  static const bool dbg = getenv("QCNN_GEMM_DBG") != nullptr;
  if (dbg) {
    QCNN_CUDA(cudaMalloc(&a.dbg, 128));
    QCNN_CUDA(cudaMemsetAsync(a.dbg, 0, 128, st));
  }
  const unsigned nb = static_cast<unsigned>(blocks);
  if (a.xl) {
    if (dbg) pq_gemm_tc_kernel<true, false, true, false, true><<<nb, 512, p.smem, st>>>(a);
    else pq_gemm_tc_kernel<false, false, true, false, true><<<nb, 512, p.smem, st>>>(a);
  } else if (a.wide) {
    if (a.bf) { if (dbg) pq_gemm_tc_kernel<true, false, true, true><<<nb, 384, p.smem, st>>>(a); else pq_gemm_tc_kernel<false, false, true, true><<<nb, 384, p.smem, st>>>(a); }
    else pq_gemm_tc_kernel<false, false, false, true><<<nb, 384, p.smem, st>>>(a);
  } else if (a.bf) {
    if (dbg && !a.lite) pq_gemm_tc_kernel<true, false, true, false><<<nb, kThreads, p.smem, st>>>(a);
    else if (a.lite) pq_gemm_tc_kernel<false, true, true, false><<<nb, kThreads, p.smem, st>>>(a);
    else pq_gemm_tc_kernel<false, false, true, false><<<nb, kThreads, p.smem, st>>>(a);
  } else {
    if (dbg && !a.lite) pq_gemm_tc_kernel<true, false, false, false><<<nb, kThreads, p.smem, st>>>(a);
    else if (a.lite) pq_gemm_tc_kernel<false, true, false, false><<<nb, kThreads, p.smem, st>>>(a);
    else pq_gemm_tc_kernel<false, false, false, false><<<nb, kThreads, p.smem, st>>>(a);
  }
  QCNN_CUDA(cudaGetLastError());
  if (dbg) {
    unsigned long long h[16];
    QCNN_CUDA(cudaStreamSynchronize(st));
    QCNN_CUDA(cudaMemcpy(h, a.dbg, 128, cudaMemcpyDeviceToHost));
    const double n = h[5] ? static_cast<double>(h[5]) : 1.0;
    fprintf(stderr, "[pq_gemm dbg] ctas=%llu NT=%d GT=%d chunks=%d ksteps=%d | per CTA: issue-loop %.0f clk (wait planes %.0f, wait weights %.0f), "
            "drain %.0f | stager loop %.0f (wait copies %.0f, wait planes free %.0f, wait cb free %.0f) | decoder loop %.0f (wait cb %.0f, "
            "wait ring %.0f; index loads %.0f, codeword loads %.0f, split + tcgen05.st %.0f, wait::st %.0f)\n", h[5], a.NT, a.GT, a.nChunks,
            a.chunkCount[0], h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[8] / n, h[9] / n,
            h[10] / n, h[11] / n, h[12] / n, h[13] / n, h[14] / n, h[4] / n, h[6] / n, h[7] / n, h[15] / n);
    cudaFree(a.dbg);
  }
  if (a.nsplit > 1)
    return LaunchSplitReduce(L->ctx, a.partial, dst, N * a.Ho * a.Wo, a.Cout, a.Cout, a.nsplit, relu, st);
  return 0;
}

// generic entry for callers that fill GemmArgs themselves (the fully-connected path in fc_aprx.cu)
size_t PqGemmSmemBytes(const GemmArgs& a) { return static_cast<size_t>(MapSmem(a).total); }
int LaunchPqGemmArgs(qcnn_ctx* ctx, const GemmArgs& a, long long blocks, cudaStream_t st) {
  const size_t smem = PqGemmSmemBytes(a);
  QCNN_CHECK(blocks >= 1 && blocks <= 2147483647LL, "pq_gemm_tc: bad grid");
  if (int rc = SetSmemLimitOnce(ctx)) return rc;
  if (a.bf) pq_gemm_tc_kernel<false, false, true, false><<<static_cast<unsigned>(blocks), kThreads, smem, st>>>(a);
  else pq_gemm_tc_kernel<false, false, false, false><<<static_cast<unsigned>(blocks), kThreads, smem, st>>>(a);
  QCNN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace qcnn
