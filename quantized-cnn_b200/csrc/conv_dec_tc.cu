// PQ convolution as a decode-at-use implicit GEMM on the 5th-generation tensor cores (sm_100a, tcgen05 + TMEM).
//
// Same function as conv_aprx.cu (reference CaffeEva::CalcFeatMap_ConvAprx + GetInPdMat, src/CaffeEva.cc:760-868,
// 1261-1296), evaluated in the other association order:
//     dst[n][ho][wo][c] = bias[c] + sum_{kh,kw} sum_s sum_j x[n][ho-pad+kh][wo-pad+kw][s*d+j] * ctrd[s][asmt[kh][kw][s][c]][j]
// i.e. the uint8 assignment index gathers the CODEWORD (d floats) instead of a LUT entry, and the inner products run on
// the tensor core.  The weights stay product-quantised in HBM (codebook + uint8 indices, exactly the layer state of the
// gather kernels); a CTA decodes the [CT channels x 8 input channels] tile of one (tap, 8-channel chunk) into shared
// memory right before the MMA that consumes it.  For large batches this replaces ~10^8 shared-memory LUT reads per
// image (the bound of the gather kernels: 32 reads / clk / SM) by tensor-core work that the B200 has to spare; for small
// batches the gather kernels win (too few 128-position tiles to fill the GPU) and the autotuner keeps them.
//
// Numerics: 3xTF32 -- every operand is split into hi (tf32) + lo (exact remainder) and x*w is accumulated as
// xh*wh + xh*wl + xl*wh in fp32 in TMEM; the dropped xl*wl term is < 2^-22 of the product (same scheme and measured
// accuracy as the LUT stage of conv_s1_tc_kernel, tools/tc_lut_test.cu).
//
// Geometry: stride-1 layers.  All images of the batch form ONE flat padded grid: image block = (Hi + pad) rows of
// pitch PW = Wi + pad; the `pad` leading rows / columns of every block are zero and double as the bottom / right
// padding of the previous row / image, so tap (kh, kw) of output position Q reads input position Q + kh*PW + kw.
// An M tile is 128 consecutive flat positions (TMEM lanes); the staged input planes are laid out so that a tap shift
// is just a different start address of the A descriptor (no im2col copy):
//     A_{hi,lo}[half][position][4 floats]   K-major / SWIZZLE_NONE core matrices: 8 positions x 16 B contiguous,
//                                           SBO = 128 B (next 8 positions), LBO = plane stride (channels 4..7)
//     B_{hi,lo}[tap][(c/8)][half][c%8][4]   canonical tile of the decoded weights, LBO = 128 B, SBO = 256 B
// Pipeline per 8-channel chunk kc: taps are processed in stages of GT taps; stage t+1 is decoded by all threads while
// the MMAs of stage t run (two weight-tile buffers, one mbarrier each, tcgen05.commit signals buffer reuse); the
// input planes / codebook slices / assignment slices of chunk kc+1 arrive by cp.async during chunk kc.
#include "qcnn_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int kThreads = 256;
constexpr int kWorkers = 224;   // decode threads (warps 0..6); thread kWorkers issues the MMAs

__device__ __forceinline__ uint32_t SmemU32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void CpAsync16(void* smemDst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0: nothing is read, the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(SmemU32(smemDst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void CpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE: start address, leading (K) / stride (M,N) byte offsets
__device__ __forceinline__ uint64_t Desc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes) {
  uint64_t d = (saddr >> 4) & 0x3FFFu;
  d |= static_cast<uint64_t>((lboBytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sboBytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version: Blackwell
  return d;
}
__device__ __forceinline__ void UmmaTf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmemD), "l"(descA), "l"(descB),
               "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void UmmaCommit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemU32(mbar)) : "memory");
}
// bounded spin: a protocol error traps (launch failure) instead of hanging the GPU
__device__ __forceinline__ void MbarWait(uint64_t* mbar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; spins++) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(SmemU32(mbar)), "r"(parity) : "memory");
    if (spins > (1u << 24)) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void SplitTf32x4(const float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
}

size_t DecSmemBytes(int NPOS, int GT, int CT, int K, int taps) {
  return static_cast<size_t>(128) * NPOS + 32 * static_cast<size_t>(NPOS) + static_cast<size_t>(128) * GT * CT +
         160 * static_cast<size_t>(K) + 4 * static_cast<size_t>(taps) * CT + 4 * static_cast<size_t>(NPOS) +
         4 * static_cast<size_t>(CT) + 64;
}

__global__ void __launch_bounds__(kThreads, 1) conv_dec_tc_kernel(const ConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int taps = a.ksz * a.ksz, CT = a.CT, MT = a.MT, GT = a.GT, K = a.K, NPOS = a.NPOS, NKC = a.NKC;
  const int NST = (taps + GT - 1) / GT;

  float4* Abuf = reinterpret_cast<float4*>(smem);               // [buf 2][hi,lo][half 2][NPOS]
  float4* rawA = Abuf + 8 * NPOS;                                // [NPOS][half 2]         cp.async target
  float4* Bbuf = rawA + 2 * NPOS;                                // [buf 2][GT][hi,lo][CT*2]
  float4* cbs = Bbuf + 8 * GT * CT;                              // [buf 2][hi,lo][half 2][K]
  float4* rawC = cbs + 8 * K;                                    // [half 2][K]            cp.async target
  uint8_t* idN = reinterpret_cast<uint8_t*>(rawC + 2 * K);       // [buf 2][half 2][taps*CT]
  int* posoff = reinterpret_cast<int*>(idN + 4 * taps * CT);     // [NPOS]
  float* biasS = reinterpret_cast<float*>(posoff + NPOS);        // [CT]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(biasS + CT);      // [0,1]: weight-tile buffers, [2]: chunk complete
  uint32_t* tmemBase = reinterpret_cast<uint32_t*>(mbar + 3);

  int b = blockIdx.x;
  const int gc = b % (a.G * a.nct);
  const int tile = b / (a.G * a.nct);
  const int ct = gc % a.nct, g = gc / a.nct;
  const int Q0 = tile * MT * 128;
  const int i0 = Q0 / a.IB;
  const float* srcBase = a.src + static_cast<size_t>(i0) * a.Hi * a.Wi * a.Cin + g * a.Cg;

  for (int p = tid; p < NPOS; p += kThreads) {
    const int F = Q0 + p;
    const int i = F / a.IB, rem = F - i * a.IB;
    const int r = rem / a.PW, c = rem - r * a.PW;
    const bool ok = i < a.N && r >= a.pad && c >= a.pad;
    posoff[p] = ok ? (((i - i0) * a.Hi + (r - a.pad)) * a.Wi + (c - a.pad)) * a.Cin : -1;
  }
  for (int c = tid; c < CT; c += kThreads) biasS[c] = __ldg(a.bias + g * a.Kg + ct * CT + c);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemU32(tmemBase)), "r"(a.tmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (int i = 0; i < 3; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemU32(mbar + i)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmemD = *tmemBase;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = CT, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(CT >> 3) << 17) | (8u << 24);

  // ---- operand staging of one 8-channel chunk ----
  auto fetchRaw = [&](int kc) {
    for (int e = tid; e < NPOS * 2; e += kThreads) {
      const int off = posoff[e >> 1];
      const int ch = kc * 8 + (e & 1) * 4;
      const bool ok = off >= 0 && ch < a.Cg;
      CpAsync16(rawA + e, srcBase + (ok ? off + ch : 0), ok);
    }
    for (int e = tid; e < 2 * K; e += kThreads) {
      const int half = e >= K ? 1 : 0, k = e - half * K;
      const int ch = kc * 8 + half * 4;
      const bool ok = ch < a.Cg;
      const int s = ok ? ch / a.d : 0, j0 = ok ? ch - s * a.d : 0;
      CpAsync16(rawC + e, a.ctrd + (static_cast<size_t>(s) * K + k) * a.d + j0, ok);
    }
    const int nG = (taps * CT) >> 4;  // 16-byte granules per assignment slice
    for (int e = tid; e < 2 * nG; e += kThreads) {
      const int half = e >= nG ? 1 : 0, q = e - half * nG;
      const int ch = kc * 8 + half * 4;
      const int s = ch < a.Cg ? ch / a.d : 0;
      const int tap = (q << 4) / CT, cc = (q << 4) - tap * CT;
      const uint8_t* gsrc = a.asmt + (static_cast<size_t>(g * a.S + s) * taps + tap) * a.KgPad + ct * CT + cc;
      CpAsync16(idN + ((kc & 1) * 2 + half) * taps * CT + (q << 4), gsrc, true);
    }
  };
  auto splitOperands = [&](int kc) {
    const int ab = kc & 1;
    float4* aHi = Abuf + (ab * 2 + 0) * 2 * NPOS;
    float4* aLo = Abuf + (ab * 2 + 1) * 2 * NPOS;
    for (int e = tid; e < NPOS * 2; e += kThreads) {
      float4 hi, lo;
      SplitTf32x4(rawA[e], hi, lo);
      const int o = (e & 1) * NPOS + (e >> 1);
      aHi[o] = hi;
      aLo[o] = lo;
    }
    float4* cHi = cbs + (ab * 2 + 0) * 2 * K;
    float4* cLo = cbs + (ab * 2 + 1) * 2 * K;
    for (int e = tid; e < 2 * K; e += kThreads) {
      float4 hi, lo;
      SplitTf32x4(rawC[e], hi, lo);
      cHi[e] = hi;
      cLo[e] = lo;
    }
  };
  // ---- weight tiles of taps [t0, t0+nt): gather the codeword halves the assignment indices name ----
  // warp 7 only issues MMAs; warps 0..6 decode, so the issue of stage t overlaps the decode of stage t+1
  auto decodeStage = [&](int kc, int t0, int nt, int buf) {
    if (tid >= kWorkers) return;
    const int ab = kc & 1;
    const float4* cHi = cbs + (ab * 2 + 0) * 2 * K;
    const float4* cLo = cbs + (ab * 2 + 1) * 2 * K;
    const int per = 2 * CT;
    for (int tapi = 0; tapi < nt; tapi++) {
      const uint8_t* idb = idN + ab * 2 * taps * CT + (t0 + tapi) * CT;
      float4* Bt = Bbuf + (buf * GT + tapi) * 4 * CT;
      for (int r = tid; r < per; r += kWorkers) {
        const int half = r >= CT ? 1 : 0, c = r - half * CT;
        const int idx = idb[half * taps * CT + c];
        const int o = (c >> 3) * 16 + half * 8 + (c & 7);
        Bt[o] = cHi[half * K + idx];
        Bt[2 * CT + o] = cLo[half * K + idx];
      }
    }
  };
  // descriptors advance by (bytes >> 4) in their low word: A by the tap shift / 128 positions per M tile, B per tap
  auto issueStage = [&](int kc, int t0, int nt, int buf) {  // one thread
    const int ab = kc & 1;
    const uint32_t lboA = static_cast<uint32_t>(NPOS) * 16u;
    const uint64_t dAh0 = Desc(SmemU32(Abuf + (ab * 2 + 0) * 2 * NPOS), lboA, 128);
    const uint64_t dAl0 = Desc(SmemU32(Abuf + (ab * 2 + 1) * 2 * NPOS), lboA, 128);
    uint64_t dBh = Desc(SmemU32(Bbuf + buf * GT * 4 * CT), 128, 256);
    const uint64_t bLo = static_cast<uint64_t>(2 * CT), bTap = static_cast<uint64_t>(4 * CT);
    int kh = t0 / a.ksz, kw = t0 - kh * a.ksz;
    uint32_t acc = (kc | t0) != 0 ? 1u : 0u;
    for (int tapi = 0; tapi < nt; tapi++) {
      const uint64_t shift = static_cast<uint64_t>(kh * a.PW + kw);
      uint64_t dAh = dAh0 + shift, dAl = dAl0 + shift;
      uint32_t d = tmemD;
      for (int m = 0; m < MT; m++) {
        UmmaTf32(d, dAh, dBh, idesc, acc);
        UmmaTf32(d, dAh, dBh + bLo, idesc, 1u);
        UmmaTf32(d, dAl, dBh, idesc, 1u);
        dAh += 128; dAl += 128; d += static_cast<uint32_t>(CT);
      }
      acc = 1u;
      dBh += bTap;
      if (++kw == a.ksz) { kw = 0; kh++; }
    }
  };

  // prologue: chunk 0 staged and split, chunk 1 in flight
  fetchRaw(0);
  CpAsyncCommit();
  CpAsyncWaitAll();
  __syncthreads();
  splitOperands(0);
  __syncthreads();
  if (NKC > 1) { fetchRaw(1); CpAsyncCommit(); }

  int t = 0;
  for (int kc = 0; kc < NKC; kc++) {
    for (int st = 0; st < NST; st++, t++) {
      const int buf = t & 1;
      const bool last = st == NST - 1;
      // every phase of the chunk barrier is observed in order (parity waits must not skip a phase)
      if (last && kc >= 1) MbarWait(mbar + 2, (kc - 1) & 1);
      if (last && kc + 1 < NKC) {
        // operands of chunk kc+1: their planes were last read by the MMAs of chunk kc-1 (complete, see above)
        CpAsyncWaitAll();
        __syncthreads();
        splitOperands(kc + 1);
      }
      if (t >= 2) MbarWait(mbar + buf, ((t >> 1) - 1) & 1);   // the MMAs of stage t-2 released this weight buffer
      const int t0 = st * GT, nt = min(GT, taps - t0);
      decodeStage(kc, t0, nt, buf);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (last && kc + 2 < NKC) { fetchRaw(kc + 2); CpAsyncCommit(); }
      if (tid == kWorkers) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issueStage(kc, t0, nt, buf);
        UmmaCommit(mbar + buf);
        if (last) UmmaCommit(mbar + 2);
      }
    }
  }

  // ---- epilogue: TMEM (lane = position, column = channel) -> + bias, ReLU -> NHWC global ----
  MbarWait(mbar + 2, (NKC - 1) & 1);
  {
    const int quarter = warp & 3, wsel = warp >> 2;
    const int nch = CT >> 4;
    for (int m = 0; m < MT; m++) {
      const int Q = Q0 + m * 128 + quarter * 32 + lane;
      const int i = Q / a.IB, rem = Q - i * a.IB;
      const int ho = rem / a.PW, wo = rem - ho * a.PW;
      const bool valid = i < a.N && ho < a.Ho && wo < a.Wo;
      float* out = a.dst + ((static_cast<size_t>(i) * a.Ho + ho) * a.Wo + wo) * a.Cout + g * a.Kg + ct * CT;
      for (int cc = wsel; cc < nch; cc += kThreads / 128) {
        uint32_t r[16];
        const uint32_t taddr = tmemD + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(m * CT + cc * 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (valid) {
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(biasS + cc * 16 + q);
            float4 o = make_float4(__uint_as_float(r[q]) + bv.x, __uint_as_float(r[q + 1]) + bv.y,
                                   __uint_as_float(r[q + 2]) + bv.z, __uint_as_float(r[q + 3]) + bv.w);
            if (a.relu) { o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f); }
            *reinterpret_cast<float4*>(out + cc * 16 + q) = o;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemD), "r"(a.tmemCols) : "memory");
}

}  // namespace

namespace qcnn {

// Candidate tilings of the decode-at-use kernel for batch N (cost in SM-cycles, comparable with PlanConv's model):
//   MMA    = chunks * taps * MT * 3 passes * CT/2 clk   (kind::tf32, M = 128: 128 x CT x 8 MACs every CT/2 clk)
//   stage  = ~350 clk of barrier / issue latency per weight-tile stage that the queued MMAs do not cover
//   tail   = TMEM drain + stores, prologue
// env QCNN_NO_DECTC=1 removes the family (the LUT + gather kernels remain).
void PlanConvDec(const qcnn_layer* L, int N, std::vector<std::pair<double, ConvPlan>>* cands) {
  if (getenv("QCNN_NO_DECTC") != nullptr) return;
  const int G = L->grp, Cg = L->Cin / G, Kg = L->Cout / G, taps = L->ksz * L->ksz;
  if (L->stride != 1 || L->src_nchw) return;
  if (L->d % 4 != 0 || Cg % 4 != 0 || L->Cin % 4 != 0 || L->S * L->d < Cg) return;
  if (L->K < 1 || L->K > 256 || Kg % 16 != 0 || L->Cout % 4 != 0) return;
  const int PW = L->Win + L->pad, IB = (L->Hin + L->pad) * PW;
  if (L->Ho > L->Hin + L->pad || L->Wo > PW) return;
  if (static_cast<double>(N) * IB > 2.0e9) return;
  const size_t smemMax = L->ctx->smem_optin ? L->ctx->smem_optin : 227 * 1024;
  const int NKC = CeilDiv(Cg, 8);
  for (int CT = 16; CT <= std::min(Kg, 256); CT += 16) {
    if (Kg % CT != 0) continue;
    for (int MT = 1; MT <= 4; MT *= 2) {
      if (MT * CT > 512) continue;
      int tmemCols = 32;
      while (tmemCols < MT * CT) tmemCols *= 2;
      const int NPOS = RoundUp(MT * 128 + (L->ksz - 1) * (PW + 1), 8);
      if (NPOS > 16000) continue;
      int gts[5] = {L->ksz, (taps + 1) / 2, 3, 2, 1};
      for (int gi = 0; gi < 5; gi++) {
        const int GT = gts[gi];
        if (GT < 1 || (taps > 1 && GT >= taps)) continue;
        bool dup = false;
        for (int gj = 0; gj < gi; gj++) dup = dup || gts[gj] == GT;
        if (dup) continue;
        const size_t smem = DecSmemBytes(NPOS, GT, CT, L->K, taps);
        if (smem > smemMax) continue;
        ConvPlan p;
        memset(&p, 0, sizeof(p));
        p.kernel = 5; p.CPT = CT; p.J = MT; p.threads = kThreads; p.smem = smem;
        ConvArgs& a = p.a;
        a.PW = PW; a.IB = IB; a.MT = MT; a.NPOS = NPOS; a.GT = GT; a.NKC = NKC; a.tmemCols = tmemCols;
        a.CT = CT; a.nct = Kg / CT; a.R = MT; a.nstrips = GT; a.rgroups = 1; a.pwarps = 4; a.cwarps = 2;
        const int NST = CeilDiv(taps, GT);
        const double mma = static_cast<double>(NKC) * taps * MT * 3.0 * (CT / 2.0);
        const double decode = static_cast<double>(NKC) * taps * CT * 2.0 * 0.8;   // smem cycles, overlapped with the MMAs
        const double perCta = std::max(mma, decode) * 1.1 + 350.0 * NKC * NST + 3000.0 + MT * 128.0 * CT / 24.0;
        const double ctas = static_cast<double>(CeilDiv(N * IB, MT * 128)) * G * a.nct;
        const double waves = std::ceil(ctas / L->ctx->sm_count);
        cands->emplace_back(perCta * waves, p);
      }
    }
  }
}

int LaunchConvDec(const ConvPlan& p, const ConvArgs& a, cudaStream_t st) {
  const int ntiles = CeilDiv(a.N * a.IB, a.MT * 128);
  const long long blocks = static_cast<long long>(ntiles) * a.G * a.nct;
  QCNN_CHECK(blocks <= 2147483647LL, "qcnn_conv_aprx_forward: batch too large for the tensor-core tiling");
  QCNN_CUDA(cudaFuncSetAttribute(conv_dec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  conv_dec_tc_kernel<<<static_cast<unsigned>(blocks), kThreads, p.smem, st>>>(a);
  QCNN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace qcnn
