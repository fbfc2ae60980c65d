// Stand-alone check of the tcgen05 LUT stage:  D[k][n] = sum_j C[k][j] * X[n][j]   (k < 128, n < N, j < 8)
// computed as 3xTF32 (hi*hi + hi*lo + lo*hi) by tcgen05.mma kind::tf32 with the accumulator in TMEM, read back with
// tcgen05.ld and compared against an fp32 FMA reference on the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_lut_test tools/tc_lut_test.cu && ./tc_lut_test
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

constexpr int kM = 128;   // codewords
constexpr int kN = 224;   // positions (multiple of 16, <= 256)
constexpr int kCols = 256;  // TMEM columns allocated (power of two >= kN)

__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// canonical K-major, no-swizzle operand tile: rows in groups of 8; per group chunk0 (8 rows x 16 B) then chunk1
// element (row r, k) -> byte (r/8)*256 + (k/4)*128 + (r%8)*16 + (k%4)*4
__device__ __forceinline__ int CanonIdx(int r, int k) { return (r >> 3) * 64 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t MakeDesc(const void* smem) {
  const uint64_t addr = SmemAddr(smem);
  uint64_t d = (addr >> 4) & 0x3FFF;
  d |= static_cast<uint64_t>(8) << 16;    // leading byte offset  = 128 B  (between the two K chunks)
  d |= static_cast<uint64_t>(16) << 32;   // stride byte offset   = 256 B  (between 8-row groups)
  d |= static_cast<uint64_t>(1) << 46;    // descriptor version (Blackwell)
  return d;                               // layout type 0 = SWIZZLE_NONE, base offset 0
}

__global__ void __launch_bounds__(128, 1) tc_lut_kernel(const float* __restrict__ C, const float* __restrict__ X,
                                                        float* __restrict__ D) {
  __shared__ __align__(128) float aHi[kM * 8], aLo[kM * 8], bHi[kN * 8], bLo[kN * 8];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmemBase;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int e = tid; e < kM * 8; e += 128) {
    const int r = e >> 3, k = e & 7;
    const float v = C[e];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    aHi[CanonIdx(r, k)] = hi;
    aLo[CanonIdx(r, k)] = v - hi;
  }
  for (int e = tid; e < kN * 8; e += 128) {
    const int r = e >> 3, k = e & 7;
    const float v = X[e];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    bHi[CanonIdx(r, k)] = hi;
    bLo[CanonIdx(r, k)] = v - hi;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(&tmemBase)), "r"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(&mbar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  // operand tiles were written through the generic proxy; the tensor core reads them through the async proxy
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = tmemBase;

  if (tid == 0) {
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kN >> 3) << 17) |
                           (static_cast<uint32_t>(kM >> 4) << 24);
    const uint64_t dAh = MakeDesc(aHi), dAl = MakeDesc(aLo), dBh = MakeDesc(bHi), dBl = MakeDesc(bLo);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase), "l"(dAh), "l"(dBh),
                 "r"(idesc), "r"(0));
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase), "l"(dAh), "l"(dBl),
                 "r"(idesc), "r"(1));
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase), "l"(dAl), "l"(dBh),
                 "r"(idesc), "r"(1));
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemAddr(&mbar)));
  }
  // wait for the MMAs (phase 0)
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(SmemAddr(&mbar)), "r"(0));
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");

  // warp w reads TMEM lanes [32w, 32w+32): thread t holds row k = 32w + t, 32 consecutive columns per load
  const int k = warp * 32 + lane;
  for (int n0 = 0; n0 < kN; n0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tbase + (static_cast<uint32_t>(warp * 32) << 16) + n0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int i = 0; i < 32; i++)
      if (n0 + i < kN) D[k * kN + n0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(kCols));
}

int main() {
  std::vector<float> C(kM * 8), X(kN * 8), D(kM * kN), R(kM * kN);
  srand(1);
  for (auto& v : C) v = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
  for (auto& v : X) v = (rand() / (float)RAND_MAX) * 100.0f;
  for (int n = 200; n < kN; n++) for (int j = 0; j < 8; j++) X[n * 8 + j] = 0.0f;  // padding columns must give exact zeros
  for (int k = 0; k < kM; k++)
    for (int n = 0; n < kN; n++) {
      float v = 0.0f;
      for (int j = 0; j < 8; j++) v = fmaf(X[n * 8 + j], C[k * 8 + j], v);
      R[k * kN + n] = v;
    }
  float *dC, *dX, *dD;
  cudaMalloc(&dC, C.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, D.size() * 4);
  tc_lut_kernel<<<1, 128>>>(dC, dX, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxAbs = 0, maxRel = 0, maxRef = 0;
  int bad = 0;
  for (int i = 0; i < kM * kN; i++) {
    const double d = fabs((double)D[i] - R[i]);
    maxAbs = fmax(maxAbs, d);
    maxRef = fmax(maxRef, fabs((double)R[i]));
    maxRel = fmax(maxRel, d / fmax(1.0, fabs((double)R[i])));
    if (!(d <= 1e-3)) bad++;
  }
  printf("N=%d max|ref|=%.3f max abs err=%.3e max rel err=%.3e bad=%d\n", kN, maxRef, maxAbs, maxRel, bad);
  printf("sample D[0][0..3] = %f %f %f %f | ref %f %f %f %f\n", D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
  printf("sample D[77][150] = %f ref %f ; D[5][210] (padding) = %g\n", D[77 * kN + 150], R[77 * kN + 150], D[5 * kN + 210]);
  return bad == 0 ? 0 : 1;
}
