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
//   * a thread owns J positions x CPT channels (<= 64 accumulators); per 4 channels it issues one 128-bit
//     broadcast load of 4 pre-multiplied LUT row offsets, then 4 x (1 IADD + J LDS + J FADD);
//   * the per-(s) assignment slice is staged in shared memory already multiplied by the LUT row pitch.
// Two kernels share that inner loop:
//   conv_s1_kernel   stride 1 (conv2..conv5, sweep): a CTA owns (image, row strip, group, channel tile) and the
//                    flat padded grid of the strip; tap (kh,kw) is a constant shift kh*PW+kw of the position.
//   conv_roll_kernel any stride (conv1: 11x11 / 4): a CTA walks the input rows of its strip once; each input
//                    row's LUT is built once and used by the <= ceil(k/stride) output rows that see it, each
//                    owned by a different "row group" of warps; columns are de-interleaved by stride phase so
//                    that a tap again is a constant shift.
#include "qcnn_internal.h"

#include <algorithm>
#include <cmath>

namespace {

constexpr int kMaxThreads = 512;

// ------------------------------------------------------------------------------------------------------------
// LUT slice build shared by both kernels: lut[k][pos] (+)= sum_{jj<8} x[jj] * cb[k][jj]  for k in [kbeg,kend)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void BuildColumn(float* __restrict__ lut, const float* __restrict__ cb, int PP, int pos,
                                            int kbeg, int kend, const float (&x)[8], bool first, bool narrow) {
  if (narrow) {  // only x[0..3] are non-zero (d <= 4 or last partial chunk): half the FMAs
#pragma unroll 4
    for (int k = kbeg; k < kend; k++) {
      const float4 c0 = *reinterpret_cast<const float4*>(cb + k * 8);
      float v = first ? 0.0f : lut[k * PP + pos];
      v = fmaf(x[0], c0.x, v);
      v = fmaf(x[1], c0.y, v);
      v = fmaf(x[2], c0.z, v);
      v = fmaf(x[3], c0.w, v);
      lut[k * PP + pos] = v;
    }
  } else {
#pragma unroll 4
    for (int k = kbeg; k < kend; k++) {
      const float4 c0 = *reinterpret_cast<const float4*>(cb + k * 8);
      const float4 c1 = *reinterpret_cast<const float4*>(cb + k * 8 + 4);
      float v = first ? 0.0f : lut[k * PP + pos];
      v = fmaf(x[0], c0.x, v);
      v = fmaf(x[1], c0.y, v);
      v = fmaf(x[2], c0.z, v);
      v = fmaf(x[3], c0.w, v);
      v = fmaf(x[4], c1.x, v);
      v = fmaf(x[5], c1.y, v);
      v = fmaf(x[6], c1.z, v);
      v = fmaf(x[7], c1.w, v);
      lut[k * PP + pos] = v;
    }
  }
}

// gather of one tap: acc[j][c] += lut[idx[c]][q_j + shift]; `base` = byte address of lut[0][q_0 + shift]
template <int CPT, int J>
__device__ __forceinline__ void GatherTap(float (&acc)[J][CPT], const char* base, const uint32_t* __restrict__ ip) {
#pragma unroll
  for (int c4 = 0; c4 < CPT; c4 += 4) {
    const uint4 o = *reinterpret_cast<const uint4*>(ip + c4);
    const char* b0 = base + o.x;
    const char* b1 = base + o.y;
    const char* b2 = base + o.z;
    const char* b3 = base + o.w;
#pragma unroll
    for (int j = 0; j < J; j++) {
      acc[j][c4 + 0] += *reinterpret_cast<const float*>(b0 + j * 128);
      acc[j][c4 + 1] += *reinterpret_cast<const float*>(b1 + j * 128);
      acc[j][c4 + 2] += *reinterpret_cast<const float*>(b2 + j * 128);
      acc[j][c4 + 3] += *reinterpret_cast<const float*>(b3 + j * 128);
    }
  }
}

template <int CPT>
__device__ __forceinline__ void StoreChannels(float* __restrict__ out, const float (&v)[CPT], int relu) {
#pragma unroll
  for (int c = 0; c < CPT; c += 4) {
    float4 o = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
    if (relu) {
      o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f);
    }
    *reinterpret_cast<float4*>(out + c) = o;
  }
}

// ------------------------------------------------------------------------------------------------------------
// stride-1 kernel
// ------------------------------------------------------------------------------------------------------------
template <int CPT, int J>
__global__ void __launch_bounds__(kMaxThreads, 1) conv_s1_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int taps = a.ksz * a.ksz;
  float* lut = reinterpret_cast<float*>(smem);                      // [K][PP]
  uint32_t* idx = reinterpret_cast<uint32_t*>(lut + a.K * a.PP);    // [taps][CT] byte offsets k*PP*4
  float* cb = reinterpret_cast<float*>(idx + taps * a.CT);          // [K][8]

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

  float acc[J][CPT];
#pragma unroll
  for (int c = 0; c < CPT; c++) {
    const float bv = __ldg(a.bias + g * a.Kg + cbase + c);
#pragma unroll
    for (int j = 0; j < J; j++) acc[j][c] = bv;
  }

  const int kper = a.K / a.ksplit;
  const uint32_t rowBytes = static_cast<uint32_t>(a.PP) * 4u;
  for (int s = 0; s < a.S; s++) {
    const int dsel = min(a.Cg - s * a.d, a.d);  // dims of this subspace that exist (reference CaffeEva.cc:1277)
    // ---- LUT stage ----
    for (int jc = 0; jc == 0 || jc < dsel; jc += 8) {
      __syncthreads();
      for (int e = tid; e < a.K * 8; e += T) {
        const int k = e >> 3, j = jc + (e & 7);
        cb[e] = (j < dsel) ? __ldg(a.ctrd + (static_cast<size_t>(s) * a.K + k) * a.d + j) : 0.0f;
      }
      if (jc == 0) {
        const uint8_t* ap = a.asmt + (static_cast<size_t>(g) * a.S + s) * taps * a.KgPad + ct * a.CT;
        for (int e = tid; e < taps * a.CT; e += T) {
          const int tap = e / a.CT, c = e - tap * a.CT;
          idx[e] = static_cast<uint32_t>(__ldg(ap + tap * a.KgPad + c)) * rowBytes;
        }
      }
      __syncthreads();
      const bool narrow = (dsel - jc) <= 4;
      for (int e = tid; e < a.PP * a.ksplit; e += T) {
        const int kq = e / a.PP;
        const int pos = e - kq * a.PP;
        float x[8];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) x[jj] = 0.0f;
        const int pp = pos - a.pad;
        if (pp >= 0) {
          const int ri = pp / a.PW, wi = pp - ri * a.PW;
          const int hi = hin0 + ri;
          if (ri < a.RI && wi < a.Wi && hi >= 0 && hi < a.Hi) {
            const float* xp = a.src + ((static_cast<size_t>(n) * a.Hi + hi) * a.Wi + wi) * a.Cin + g * a.Cg + s * a.d + jc;
#pragma unroll
            for (int jj = 0; jj < 8; jj++)
              if (jc + jj < dsel) x[jj] = __ldg(xp + jj);
          }
        }
        BuildColumn(lut, cb, a.PP, pos, kq * kper, (kq + 1) * kper, x, jc == 0, narrow);
      }
    }
    __syncthreads();
    // ---- gather stage ----
    const char* lutq = reinterpret_cast<const char*>(lut) + q0 * 4;
    const uint32_t* ip = idx + cw * CPT;
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
      StoreChannels<CPT>(out, acc[j], a.relu);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// rolling-row kernel (any stride)
// ------------------------------------------------------------------------------------------------------------
template <int CPT, int J>
__global__ void __launch_bounds__(kMaxThreads, 1) conv_roll_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int taps = a.ksz * a.ksz;
  float* lut = reinterpret_cast<float*>(smem);                      // [K][PP], PP = stride * PH
  uint32_t* idx = reinterpret_cast<uint32_t*>(lut + a.K * a.PP);    // [taps][CT]
  float* cb = reinterpret_cast<float*>(idx + taps * a.CT);          // [K][8]

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

  float bias[CPT];
  float acc[J][CPT];
#pragma unroll
  for (int c = 0; c < CPT; c++) {
    bias[c] = __ldg(a.bias + g * a.Kg + cbase + c);
#pragma unroll
    for (int j = 0; j < J; j++) acc[j][c] = bias[c];
  }
  int ho_cur = ho0 + rg;  // output row this warp is accumulating

  const int kper = a.K / a.ksplit;
  const uint32_t rowBytes = static_cast<uint32_t>(a.PP) * 4u;
  const int hi_begin = ho0 * a.stride - a.pad;
  const int hi_end = (ho_end - 1) * a.stride - a.pad + a.ksz;
  bool idx_ready = false;
  for (int hi = hi_begin; hi < hi_end; hi++) {
    const bool row_valid = hi >= 0 && hi < a.Hi;   // CTA-uniform; rows outside the image contribute nothing
    const int kh = hi + a.pad - ho_cur * a.stride;  // warp-uniform
    const bool active = row_valid && ho_cur < ho_end && kh >= 0 && kh < a.ksz;
    if (row_valid) {
      for (int s = 0; s < a.S; s++) {
        const int dsel = min(a.Cg - s * a.d, a.d);
        for (int jc = 0; jc == 0 || jc < dsel; jc += 8) {
          __syncthreads();
          for (int e = tid; e < a.K * 8; e += T) {
            const int k = e >> 3, j = jc + (e & 7);
            cb[e] = (j < dsel) ? __ldg(a.ctrd + (static_cast<size_t>(s) * a.K + k) * a.d + j) : 0.0f;
          }
          if (jc == 0 && (a.S > 1 || !idx_ready)) {
            const uint8_t* ap = a.asmt + (static_cast<size_t>(g) * a.S + s) * taps * a.KgPad + ct * a.CT;
            for (int e = tid; e < taps * a.CT; e += T) {
              const int tap = e / a.CT, c = e - tap * a.CT;
              idx[e] = static_cast<uint32_t>(__ldg(ap + tap * a.KgPad + c)) * rowBytes;
            }
            idx_ready = true;
          }
          __syncthreads();
          const bool narrow = (dsel - jc) <= 4;
          for (int e = tid; e < a.PP * a.ksplit; e += T) {
            const int kq = e / a.PP;
            const int pos = e - kq * a.PP;
            const int phase = pos / PH, i = pos - phase * PH;
            const int wi = i * a.stride + phase - a.pad;
            float x[8];
#pragma unroll
            for (int jj = 0; jj < 8; jj++) x[jj] = 0.0f;
            if (wi >= 0 && wi < a.Wi) {
              const int ch0 = g * a.Cg + s * a.d + jc;
              if (a.src_nchw) {
                const float* xp = a.src + ((static_cast<size_t>(n) * a.Cin + ch0) * a.Hi + hi) * a.Wi + wi;
                const size_t plane = static_cast<size_t>(a.Hi) * a.Wi;
#pragma unroll
                for (int jj = 0; jj < 8; jj++)
                  if (jc + jj < dsel) x[jj] = __ldg(xp + jj * plane);
              } else {
                const float* xp = a.src + ((static_cast<size_t>(n) * a.Hi + hi) * a.Wi + wi) * a.Cin + ch0;
#pragma unroll
                for (int jj = 0; jj < 8; jj++)
                  if (jc + jj < dsel) x[jj] = __ldg(xp + jj);
              }
            }
            BuildColumn(lut, cb, a.PP, pos, kq * kper, (kq + 1) * kper, x, jc == 0, narrow);
          }
        }
        __syncthreads();
        if (active) {
          const char* lutq = reinterpret_cast<const char*>(lut) + wo0 * 4;
          const uint32_t* ip = idx + (kh * a.ksz) * a.CT + cw * CPT;
          for (int kw = 0; kw < a.ksz; kw++) {
            const int phase = kw % a.stride, sh = kw / a.stride;
            GatherTap<CPT, J>(acc, lutq + (phase * PH + sh) * 4, ip);
            ip += a.CT;
          }
        }
      }
    }
    // this input row was the last one the current output row needs: emit it and move on
    if (ho_cur < ho_end && kh == a.ksz - 1) {
#pragma unroll
      for (int j = 0; j < J; j++) {
        const int wo = wo0 + 32 * j;
        if (wo < a.Wo) {
          float* out = a.dst + ((static_cast<size_t>(n) * a.Ho + ho_cur) * a.Wo + wo) * a.Cout + g * a.Kg + cbase;
          StoreChannels<CPT>(out, acc[j], a.relu);
        }
#pragma unroll
        for (int c = 0; c < CPT; c++) acc[j][c] = bias[c];
      }
      ho_cur += a.rgroups;
    }
  }
}

template <int CPT, int J>
int LaunchOne(const ConvPlan& p, const ConvArgs& a, cudaStream_t st) {
  dim3 grid(a.G * a.nct * a.nstrips, a.N);
  if (p.kernel == 0) {
    auto kern = conv_s1_kernel<CPT, J>;
    QCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<grid, p.threads, p.smem, st>>>(a);
  } else {
    auto kern = conv_roll_kernel<CPT, J>;
    QCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    kern<<<grid, p.threads, p.smem, st>>>(a);
  }
  QCNN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

namespace qcnn {

// Chooses the tiling of a conv layer.  Cost model (SM-cycles per image, lower is better):
//   gather = lookups issued (incl. idle lanes / garbage columns) / 32 per clk
//   build  = LUT entries built (incl. halo rows and per-channel-tile rebuilds) * (d_eff + 2) / 128 per clk
// divided over the CTAs of one image and multiplied by the number of waves the whole batch needs on this GPU,
// so small batches trade LUT rebuilds for parallelism and large batches do not.
int PlanConv(qcnn_layer* L, int N) {
  if (L->plan_N == N) return 0;
  ConvPlan best;
  double bestCost = 1e300;
  bool found = false;
  const int G = L->grp, Cg = L->Cin / G, Kg = L->Cout / G;
  const int taps = L->ksz * L->ksz;
  const size_t smemMax = L->ctx->smem_optin ? L->ctx->smem_optin : 227 * 1024;
  const double dEff = std::min(L->d, Cg) <= 4 ? 4.0 : 8.0 * CeilDiv(std::min(L->d, Cg), 8);
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
          ConvPlan p;
          memset(&p, 0, sizeof(p));
          ConvArgs& a = p.a;
          a.R = R;
          a.nstrips = CeilDiv(L->Ho, R);
          a.CT = CT; a.nct = nct; a.cwarps = cwarps;
          double lanesPerRow, builtPerStrip;
          int warps;
          if (L->stride == 1) {
            p.kernel = 0;
            a.PW = L->Win + L->pad;
            a.RI = R + L->ksz - 1;
            a.pwarps = CeilDiv(R * a.PW, 32 * J);
            a.rgroups = 1;
            const int need1 = a.RI * a.PW + std::max(L->pad, L->ksz - 1) + 1;
            const int need2 = a.pwarps * 32 * J + (L->ksz - 1) * a.PW + (L->ksz - 1) + 1;
            a.PP = RoundUp(std::max(need1, need2), 4);
            warps = a.pwarps * cwarps;
            lanesPerRow = static_cast<double>(a.pwarps) * 32 * J / R;  // lanes issued per output row
            builtPerStrip = a.PP;
          } else {
            p.kernel = 1;
            a.rgroups = CeilDiv(L->ksz, L->stride);
            a.pwarps = CeilDiv(L->Wo, 32 * J);
            const int ph1 = CeilDiv(L->Win + 2 * L->pad, L->stride);
            const int ph2 = a.pwarps * 32 * J + (L->ksz - 1) / L->stride + 1;
            a.PW = std::max(ph1, ph2);           // phase length
            a.PP = RoundUp(a.PW * L->stride, 4);
            a.RI = 0;
            warps = a.pwarps * cwarps * a.rgroups;
            lanesPerRow = static_cast<double>(a.pwarps) * 32 * J * (static_cast<double>(a.rgroups) * L->stride / L->ksz);
            builtPerStrip = static_cast<double>(a.PP) * ((R - 1) * L->stride + L->ksz);  // positions x input rows
          }
          p.CPT = CPT; p.J = J;
          p.threads = warps * 32;
          if (p.threads > kMaxThreads || p.threads < 64) continue;
          p.smem = sizeof(float) * (static_cast<size_t>(L->K) * a.PP + static_cast<size_t>(taps) * CT + L->K * 8);
          if (p.smem > smemMax) continue;
          a.ksplit = 1;
          while (a.ksplit * 2 <= L->K && a.PP * (a.ksplit * 2) <= p.threads && L->K % (a.ksplit * 2) == 0) a.ksplit *= 2;
          // cost per image
          const double rowsIssued = static_cast<double>(a.nstrips) * R;
          const double gather = rowsIssued * lanesPerRow * taps * L->S * Kg * G / 32.0;
          const double build = a.nstrips * builtPerStrip * L->K * L->S * (dEff + 2.0) * nct * G / 128.0;
          // mild preference for more resident warps (latency hiding) and fewer, larger CTAs
          const double occPenalty = warps < 8 ? 1.25 : (warps < 12 ? 1.08 : 1.0);
          const double ctasPerImg = static_cast<double>(G) * nct * a.nstrips;
          const double waves = std::ceil(ctasPerImg * N / L->ctx->sm_count);
          const double cost = (gather + build) / ctasPerImg * waves * occPenalty;
          if (cost < bestCost) { bestCost = cost; best = p; found = true; }
        }
      }
    }
  }
  QCNN_CHECK(found, "qcnn_conv_layer_create: no tiling fits (Cout/grp=%d must be a multiple of 16; K=%d, k=%d, W=%d)",
             Kg, L->K, L->ksz, L->Win);
  ConvArgs& a = best.a;
  a.Hi = L->Hin; a.Wi = L->Win; a.Cin = L->Cin; a.Ho = L->Ho; a.Wo = L->Wo; a.Cout = L->Cout;
  a.ksz = L->ksz; a.pad = L->pad; a.stride = L->stride; a.G = G; a.Cg = Cg; a.Kg = Kg;
  a.KgPad = RoundUp(Kg, 16);
  a.S = L->S; a.K = L->K; a.d = L->d;
  L->plan = best;
  L->plan_N = N;
  return 0;
}

int LaunchConv(qcnn_layer* L, const float* src, int N, float* dst, int relu, cudaStream_t st) {
  QCNN_CHECK(L->kind == QCNN_KIND_CONV, "qcnn_conv_aprx_forward: layer is not convolutional");
  QCNN_CHECK(N >= 1 && N <= 65535, "qcnn_conv_aprx_forward: N must be in [1, 65535] (got %d)", N);
  if (int prc = PlanConv(L, N)) return prc;
  ConvPlan& p = L->plan;
  ConvArgs a = p.a;
  a.src = src; a.dst = dst; a.ctrd = L->d_ctrd; a.asmt = L->d_asmt; a.bias = L->d_bias;
  a.N = N; a.relu = relu; a.src_nchw = L->src_nchw;
  QCNN_CHECK(!(a.src_nchw && p.kernel == 0), "qcnn_conv_aprx_forward: NCHW source is only supported by the strided kernel");
  int rc = 1;
#define QCNN_DISPATCH(C, JJ) if (p.CPT == C && p.J == JJ) rc = LaunchOne<C, JJ>(p, a, st); else
  QCNN_DISPATCH(32, 1) QCNN_DISPATCH(32, 2) QCNN_DISPATCH(16, 1) QCNN_DISPATCH(16, 2) QCNN_DISPATCH(16, 3)
  QCNN_DISPATCH(16, 4) { SetError("internal: no conv instantiation for CPT=%d J=%d", p.CPT, p.J); return 1; }
#undef QCNN_DISPATCH
  if (rc == 0) L->ctx->launches++;
  return rc;
}

}  // namespace qcnn
