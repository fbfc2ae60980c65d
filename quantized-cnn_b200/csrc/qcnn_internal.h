// Internal declarations shared by the .cu / .cc files of libqcnn_b200.so (not part of the C ABI).
#ifndef QCNN_INTERNAL_H_
#define QCNN_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/qcnn.h"

namespace qcnn {

// thread-local error text behind qcnn_last_error()
void SetError(const char* fmt, ...);
int CudaFail(cudaError_t e, const char* what, const char* file, int line);

#define QCNN_CUDA(call)                                                      \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) return ::qcnn::CudaFail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define QCNN_CHECK(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ::qcnn::SetError(__VA_ARGS__);     \
      return 1;                          \
    }                                    \
  } while (0)

// Every entry point that needs a particular GPU makes it current for the duration of the call and puts the caller's
// device back on return: the CUDA "current device" is process-visible state (PyTorch's default device IS the runtime's),
// and a library that leaves it changed silently moves the caller's next allocation to another GPU.
class DeviceGuard {
 public:
  explicit DeviceGuard(int device) : prev_(-1), err_(cudaSuccess) {
    if (cudaGetDevice(&prev_) != cudaSuccess) prev_ = -1;
    if (prev_ != device) err_ = cudaSetDevice(device); else prev_ = -1;   // nothing to restore when already current
  }
  ~DeviceGuard() { if (prev_ >= 0) cudaSetDevice(prev_); }
  cudaError_t error() const { return err_; }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
 private:
  int prev_;
  cudaError_t err_;
};
#define QCNN_ON_DEVICE(dev)                 \
  ::qcnn::DeviceGuard device_guard__(dev);  \
  QCNN_CUDA(device_guard__.error())

inline int CeilDiv(int a, int b) { return (a + b - 1) / b; }
inline int RoundUp(int a, int b) { return CeilDiv(a, b) * b; }

}  // namespace qcnn

struct qcnn_ctx {
  int device;
  int sm_count;
  int cc_major, cc_minor;
  size_t smem_optin;  // max dynamic shared memory per block (opt-in)
  unsigned long long launches;  // kernels launched through this ctx (monotonic)
  unsigned long long alloc_epoch;  // bumped whenever a device scratch buffer is (re)allocated: captured graphs go stale
};

enum { QCNN_KIND_CONV = 0, QCNN_KIND_FC = 1 };

// ---- kernel argument blocks --------------------------------------------------------------------------------
struct FcArgs {
  const float* src;
  float* dst;
  float* partial;      // [nsplit][N][DoutPad] when nsplit > 1
  const float* ctrd;   // [S][K][d] (file order)
  const uint8_t* asmt; // [S][DoutPad], value = idx << kshift
  const float* bias;
  int N, Din, Dout, DoutPad, S, K, d;
  const int* srcoff;   // [Din] element offset of flattened feature f inside one source image (NULL: identity)
  int s_per_split, nsplit;
  int relu;
};

struct ConvArgs {
  const float* src;
  float* dst;
  const float* ctrd;    // [S][K][d] (file order), shared by all groups
  const uint8_t* asmt;  // [G][S][taps][KgPad]
  const float* bias;
  int N, Hi, Wi, Cin, Ho, Wo, Cout, ksz, pad, stride, G, Cg, Kg, KgPad, S, K, d;
  int src_nchw;
  // tiling plan
  int R, nstrips;       // output rows per CTA / strips per image
  int PW;               // s1: row pitch of the flat padded grid; roll: phase length PH
  int RI;               // s1: input rows held per strip
  int PP;               // LUT pitch in floats (positions per codeword row), multiple of 4
  int CT, nct;          // output channels per CTA (per group) / channel tiles per group
  int pwarps, cwarps, rgroups;
  int ksplit;           // LUT build: K is split over ksplit thread groups
  int relu;
};

// ---- decode-at-use tensor-core GEMM (pq_gemm_tc.cu) ----
constexpr int kMaxKSteps = 96;
struct KStep {           // one K = 8 step of the GEMM: two 4-float halves
  int bStart;            // B operand: first float4 of half 0 inside the staged planes (position 0 of the tile)
  int lbo;               // distance (in float4) from half 0 to half 1
  short idx0, idx1;      // assignment-index row of either half inside the chunk's staged rows
  short cb0, cb1;        // codebook slot of either half inside the chunk's staged slices
};
struct GemmArgs {
  const float* src;
  float* dst;
  const float* ctrd;
  const uint8_t* asmt;
  const float* bias;
  unsigned long long* dbg;    // optional cycle counters of the MMA issuer (QCNN_GEMM_DBG=1)
  long long srcImg, dstImg;   // elements per source / destination image
  int N, Hi, Wi, Cin, Ho, Wo, Cout, ksz, pad, stride, G, Cg, Kg, KgPad, S, K, d;
  int mode;                   // 0 stride-1 conv, 1 strided conv on phase planes, 2 fully connected
  int PW, IB;                 // flat grid: row pitch / positions per image
  int rowStride, colStride, chStride;   // mode 1: source element strides of an input row / column / channel
  int nsplit, kAll, kPerSplit;  // mode 2: K splits, k-steps of the layer / per split (partial sums -> fc_reduce)
  int kshift;                   // stored assignment byte = index << kshift
  float* partial;               // mode 2, nsplit > 1: [nsplit][N][dstRow]
  int dstRow;                   // mode 2: floats per destination row
  int nPB;                      // position-plane buffers (2; 3 for mode 2, whose planes arrive by bulk copy)
  const float* xprep;           // mode 2: activations pre-split into the plane image (fc_prep_kernel), [tile][chunk][hi,lo][2KS][NT][4]
  int nChunksAll;               // mode 2: chunks of the whole layer (xprep indexing)
  int cbF4;                     // float4 per codebook slot (K for d % 4 == 0 pieces, K/4 for d == 1 scalars)
  int NT;                     // positions per CTA = MMA N (multiple of 16, <= 256)
  int aOff;                   // first TMEM column of the decoded-weight ring (after the accumulator(s))
  int corr;                   // 1: the 3xTF32 cross terms have their own accumulator at column NT (added in the epilogue)
  int lite;                   // 1: the two-CTAs-per-SM instantiation (256 TMEM columns, short register windows)
  int wide;                   // 1: the twelve-warp instantiation (second decoder warp group)
  int xl;                     // 1: the sixteen-warp instantiation (second decoder group + 224 stager threads; mode 1)
  int bf;                     // 1: bf16x2 operands (x = x1 + x2, w = w1 + w2 as bf16 pieces; TWO kind::f16 MMAs of K = 16 per
                              //    k-step: [w1|w1].[x1|x2] + [w2|w2].[x1|x2]) instead of 3xTF32 (three kind::tf32 MMAs of K = 8)
  const uint8_t* asmtT;       // qcnn_layer::d_asmt_t (mode 0 with idxT)
  int idxT, tapsPad;          // 1: the chunk's indices are staged channel-major ([half][128 channels][tapsPad bytes])
  int bulkC;                  // mode 0 with idxT and cbPre: codebook slices + index blocks arrive by four cp.async.bulk per chunk
  int cbPre;                  // 1: `ctrd` is the pre-split codebook (qcnn_layer::d_ctrd_bf): the decoders convert nothing
  int planeRows;              // 16-byte rows per staged plane (3xTF32: planeF4; bf16x2: positions x k-step groups)
  int NPOS;                   // staged positions per plane (NT + halo)
  int planeF4;                // float4 per staged plane set (one of hi / lo, one buffer)
  int cbSlots, idRows;        // codebook slices / index rows staged per chunk
  int nChunks;
  int GT, NSLOT;              // k-steps per stage, stages in the TMEM weight ring
  int nct;                    // 128-channel tiles per group
  int relu;
  int ntab;
  int chunkFirst[8], chunkCount[8];   // k-step table range of a chunk (mode 1: per phase row; else entry 0)
  KStep tab[kMaxKSteps];
};

struct ConvPlan {
  int kernel;           // 0 s1, 1 roll, 2 s1_tc, 3 roll_tc, 4 direct, 6 pq_gemm_tc (decode-at-use GEMMs on tcgen05); 5 retired
  int CPT, J;
  int threads;
  size_t smem;
  ConvArgs a;           // geometry + tiling (pointers/N filled at launch)
  GemmArgs g;           // kernel 6 (pq_gemm_tc)
};

struct qcnn_layer {
  qcnn_ctx* ctx;
  int kind;
  // geometry
  int Cin, Hin, Win, Cout, ksz, pad, stride, grp, S, K, d;
  int Ho, Wo;
  int Din, Dout, DoutPad;
  // FC source permutation (NHWC map) / conv NCHW source
  int src_h, src_w, src_c;
  int* d_srcoff;        // FC: NCHW-flatten index -> NHWC element offset (NULL when the source is flat)
  int src_nchw;
  // device parameters
  float* d_ctrd;
  uint8_t* d_asmt;
  int asmt_t_mode;        // layout of d_asmt_t: 0 = [grp][S][KgPad][tapsPad] (stride 1), 1 = [grp][stride phase rows][KgPad][rowsPad] (strided, S = 1)
  uint8_t* d_asmt_t;      // conv layers: the same indices as [grp][S][KgPad][tapsPad] (tapsPad = taps rounded up to 16): a decoder thread
                          // (= channel) reads all taps of a chunk with one or two 128-bit loads (pq_gemm_tc.cu, mode 0)
  float* d_bias;
  size_t asmt_bytes;
  int kshift;           // FC: stored assignment = idx << kshift
  // plans
  ConvPlan plan;
  int plan_N;            // batch size the cached plan was made for (0 = none)
  int tuned;             // the plan for plan_N was confirmed by on-device timing
  std::vector<ConvPlan>* cands;  // the model's best tilings for plan_N (autotune candidates)
  std::vector<std::pair<int, ConvPlan>>* tunedPlans;  // timed winners per batch size
  // FC scratch
  float* d_partial;
  size_t partial_bytes;
  float* d_flat;         // tensor-core FC path: source pre-split into hi/lo plane images (fc_prep_kernel)
  size_t flat_bytes;
  void* d_ctrd_bf;        // codebook pre-split into bf16 pieces (pq_gemm_tc.cu, bf16x2 operands): [S][d/4 pieces][K] x {w1 (4 x bf16), w2 (4 x bf16)}: a (subspace, piece) slice is contiguous
  float* d_cpart;        // chain kernel (fc_chain.cu): per-CTA partial sums [sm_count][DoutPad], words double as ready flags
  size_t cpart_bytes;
  // tuning overrides (0 = automatic)
  int opt_fc_nsplit;
  int opt_fc_tn;
  int opt_no_tc;         // 1: never use the decode-at-use tensor-core kernels for this layer
  int opt_tc_bf;         // 1: the tensor-core kernels take bf16x2 operands (tensor_core = 2) instead of 3xTF32
  int opt_force_kernel;  // conv: 1 + kernel id the plan is restricted to (0 = none); tests pin the kernel they check
  int opt_no_autotune;   // conv: 1 = keep the cost model's first tiling (no on-device timing)
  int opt_gemm_nt;       // conv: restrict pq_gemm_tc tilings to this many positions per CTA (0 = any)
};

namespace qcnn {

int PlanConv(qcnn_layer* L, int N);
int DescribeConv(qcnn_layer* L, int N, char* buf, size_t cap);
int LaunchConv(qcnn_layer* L, const float* src, int N, float* dst, int relu, cudaStream_t st);
int LaunchFc(qcnn_layer* L, const float* src, int N, float* dst, int relu, cudaStream_t st);
// fc_chain.cu: one persistent launch for a run of consecutive FC layers at batch <= 4 (relu[l]: ReLU after layer l)
bool FcChainEligible(qcnn_ctx* ctx, qcnn_layer* const* layers, const int* relu, int n, int N);
void DescribeFcChain(qcnn_ctx* ctx, qcnn_layer* const* layers, const int* relu, int n, char* buf, size_t cap);
int LaunchFcChain(qcnn_ctx* ctx, qcnn_layer* const* layers, const int* relu, int n, const float* src, int N, float* dst,
                  cudaStream_t st, unsigned long long* dbg, bool* handled);
// pq_gemm_tc.cu
void PlanPqGemm(const qcnn_layer* L, int N, std::vector<std::pair<double, ConvPlan>>* cands);
size_t PqGemmSmemBytes(const GemmArgs& a);
int LaunchPqGemmArgs(qcnn_ctx* ctx, const GemmArgs& a, long long blocks, cudaStream_t st);
int BuildCtrdBf(qcnn_layer* L);   // layer creation: the bf16x2 form of the codebook (d % 4 == 0)
int LaunchSplitReduce(qcnn_ctx* ctx, const float* partial, float* dst, int rows, int cols, int colsPad, int nsplit, int relu,
                      cudaStream_t st);
void DescribeFcTc(const qcnn_layer* L, int N, char* buf, size_t cap);
int LaunchFcTc(qcnn_layer* L, const float* src, int N, float* dst, int relu, cudaStream_t st, bool* handled);
int LaunchPqGemm(qcnn_layer* L, const ConvPlan& p, const float* src, int N, float* dst, int relu, cudaStream_t st);

int LaunchRelu(qcnn_ctx* ctx, const float* src, float* dst, size_t n, cudaStream_t st);
int LaunchLrn(qcnn_ctx* ctx, const float* src, float* dst, size_t pixels, int C, int size, float alpha, float beta,
              float k, cudaStream_t st);
int LaunchMaxPool(qcnn_ctx* ctx, const float* src, float* dst, int N, int H, int W, int C, int ksz, int pad,
                  int stride, cudaStream_t st);
int LaunchLrnMaxPool(qcnn_ctx* ctx, const float* src, float* dst, int N, int H, int W, int C, int size, float alpha,
                     float beta, float k, int ksz, int pad, int stride, cudaStream_t st);
int LaunchSoftmax(qcnn_ctx* ctx, const float* src, float* dst, int N, int C, cudaStream_t st);
int LaunchNchwToNhwc(qcnn_ctx* ctx, const float* src, float* dst, int N, int C, int H, int W, cudaStream_t st);
int LaunchNhwcToNchw(qcnn_ctx* ctx, const float* src, float* dst, int N, int H, int W, int C, cudaStream_t st);

// preproc.cu
int LaunchU8ToF32(qcnn_ctx* ctx, const uint8_t* src, const float* mean, float* dst, int N, int C, int HW, cudaStream_t st);
int LaunchTopK(qcnn_ctx* ctx, const float* prob, int N, int C, int k, int mode, int* idx, float* val, cudaStream_t st);

inline int PoolOut(int in, int pad, int ksz, int stride) {
  // ceil((in + 2p - k) / s) + 1   (reference src/CaffeEva.cc:365-372)
  int num = in + 2 * pad - ksz;
  int q = num >= 0 ? (num + stride - 1) / stride : -((-num) / stride);
  return q + 1;
}

}  // namespace qcnn

#endif  // QCNN_INTERNAL_H_
