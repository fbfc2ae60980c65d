// See CaffeEva.h.  Behaviour follows /root/reference/src/CaffeEva.cc (cited per function); every computation goes
// through the C ABI (include/qcnn.h).
#include "CaffeEva.h"

#include <cfloat>
#include <cstdio>
#include <cstring>

#include "FileIO.h"

namespace {
const int kLablCntPerData = 5;   // predicted labels per image (reference CaffeEva.cc:25)
const int kEvalCntDefault = 100;  // the reference scores kDataCntInBatch * kBatchCntProc = 100 images (CaffeEva.cc:23-24,274)
}  // namespace

CaffeEva::CaffeEva(void)
    : enblAprx(true), device(0), deviceCnt(1), evalCnt(kEvalCntDefault), evalBatch(256), ctx(nullptr), net(nullptr), preproc(nullptr), multi(nullptr),
      msAllLayers(0.0) {}

CaffeEva::~CaffeEva(void) {
  if (preproc) qcnn_preproc_destroy(preproc);
  if (multi) qcnn_multi_destroy(multi);
  if (net) qcnn_net_destroy(net);
  if (ctx) qcnn_ctx_destroy(ctx);
}

bool CaffeEva::Fail(const std::string& what) {
  errorMsg = what + ": " + qcnn_last_error();
  printf("[ERROR] %s\n", errorMsg.c_str());
  return false;
}

void CaffeEva::Init(const bool enblAprxSrc) {
  enblAprx = enblAprxSrc;
  msAllLayers = 0.0;
  std::fill(msIndvLayerLst.begin(), msIndvLayerLst.end(), 0.0);
}

void CaffeEva::SetModelName(const std::string& modelNameSrc) { modelName = modelNameSrc; }

void CaffeEva::SetModelPath(const std::string& dirPathMainSrc, const std::string& fileNamePfxSrc) {
  dirPathMain = dirPathMainSrc;
  fileNamePfx = fileNamePfxSrc;
}

// reference src/CaffeEva.cc:83-107
bool CaffeEva::LoadDataset(const std::string& dirPathData) {
  if (!FileIO::ReadBinFile(dirPathData + "/dataMatTst.single.bin", &dataLst)) return false;
  if (!FileIO::ReadBinFile(dirPathData + "/lablVecTst.uint16.bin", &lablVecGrth)) return false;
  return true;
}

// reference src/CaffeEva.cc:109-149: layer table by model name, parameters from <dir>/<pfx>.*; the buffer
// preparation steps (PrepFeatMap/PrepFeatBuf/PrepCtrdBuf/PrepAsmtBuf) happen on the device inside qcnn_net_create
bool CaffeEva::LoadCaffePara(void) {
  if (!enblAprx) {
    errorMsg = "only the approximate (product-quantized) path is implemented: call Init(true)";
    printf("[ERROR] %s\n", errorMsg.c_str());
    return false;
  }
  caffeParaObj.Init(dirPathMain, fileNamePfx);
  if (!caffeParaObj.ConfigLayer_ByName(modelName)) {
    printf("[ERROR] unrecognized caffe model name: %s\n", modelName.c_str());
    return false;
  }
  if (!caffeParaObj.LoadLayerPara(enblAprx, ENUM_AsmtEnc::Compact)) return false;  // host copy, as the reference keeps
  if (!ctx && qcnn_ctx_create(device, &ctx) != 0) return Fail("qcnn_ctx_create");
  if (net) { qcnn_net_destroy(net); net = nullptr; }
  // the parameters were read ONCE, above; the device network is built from these host matrices (the re-layout and
  // upload PrepCtrdBuf / PrepAsmtBuf do on the CPU, reference src/CaffeEva.cc:534-623, happen inside the call)
  std::vector<qcnn_layer_info> infos(caffeParaObj.layerCnt);
  std::vector<qcnn_layer_para> paras(caffeParaObj.layerCnt);
  for (int l = 0; l < caffeParaObj.layerCnt; l++) {
    const LayerInfo& li = caffeParaObj.layerInfoLst[l];
    qcnn_layer_info& o = infos[l];
    o.type = static_cast<int>(li.type);
    o.padSiz = li.padSiz; o.knlSiz = li.knlSiz; o.knlCnt = li.knlCnt; o.grpCnt = li.grpCnt; o.stride = li.stride;
    o.nodCnt = li.nodCnt; o.lrnSiz = li.lrnSiz; o.lrnAlp = li.lrnAlp; o.lrnBet = li.lrnBet; o.lrnIni = li.lrnIni; o.drpRat = li.drpRat;
    qcnn_layer_para& q = paras[l];
    q.ctrd = nullptr; q.asmt = nullptr; q.bias = nullptr; q.S = q.K = q.d = 0;
    if (li.type != ENUM_LyrType::Conv && li.type != ENUM_LyrType::FCnt) continue;
    const LayerPara& lp = caffeParaObj.layerParaLst[l];
    const bool conv = li.type == ENUM_LyrType::Conv;
    const int S = lp.ctrdLst.GetDimLen(0);
    const bool ok = lp.ctrdLst.GetDimCnt() == 3 && lp.asmtLst.GetDimCnt() == (conv ? 4 : 2) &&
                    (conv ? (lp.asmtLst.GetDimLen(0) == li.knlCnt && lp.asmtLst.GetDimLen(1) == li.knlSiz &&
                             lp.asmtLst.GetDimLen(2) == li.knlSiz && lp.asmtLst.GetDimLen(3) == S && lp.biasVec.GetEleCnt() == li.knlCnt)
                          : (lp.asmtLst.GetDimLen(0) == li.nodCnt && lp.asmtLst.GetDimLen(1) == S && lp.biasVec.GetEleCnt() == li.nodCnt));
    if (!ok) {
      errorMsg = "layer " + std::to_string(l + 1) + ": parameter shapes do not match the layer table";
      printf("[ERROR] %s\n", errorMsg.c_str());
      return false;
    }
    q.ctrd = lp.ctrdLst.GetDataPtr(); q.asmt = lp.asmtLst.GetDataPtr(); q.bias = lp.biasVec.GetDataPtr();
    q.S = S; q.K = lp.ctrdLst.GetDimLen(1); q.d = lp.ctrdLst.GetDimLen(2);
  }
  if (qcnn_net_create_from_para(ctx, caffeParaObj.layerCnt, infos.data(), paras.data(), caffeParaObj.imgChnIn,
                                caffeParaObj.imgHeiIn, caffeParaObj.imgWidIn, &net) != 0)
    return Fail("qcnn_net_create_from_para");
  if (multi) { qcnn_multi_destroy(multi); multi = nullptr; }
  if (deviceCnt > 1 &&
      qcnn_multi_create_from_para(deviceCnt, nullptr, caffeParaObj.layerCnt, infos.data(), paras.data(), caffeParaObj.imgChnIn,
                                  caffeParaObj.imgHeiIn, caffeParaObj.imgWidIn, &multi) != 0)
    return Fail("qcnn_multi_create_from_para");
  qcnn_net_set_profiling(net, 1);
  msIndvLayerLst.assign(caffeParaObj.layerCnt, 0.0);
  return true;
}

bool CaffeEva::SetPreproc(const int reszType, const int meanType, const int heiFull, const int widFull, const int heiCrop,
                          const int widCrop, const Matrix<float>& meanImg) {
  if (!ctx && qcnn_ctx_create(device, &ctx) != 0) return Fail("qcnn_ctx_create");
  if (preproc) { qcnn_preproc_destroy(preproc); preproc = nullptr; }
  if (qcnn_preproc_create(ctx, reszType, meanType, heiFull, widFull, heiCrop, widCrop, meanImg.GetDataPtr(), meanImg.GetDimLen(1),
                          meanImg.GetDimLen(2), &preproc) != 0)
    return Fail("qcnn_preproc_create");
  return true;
}

bool CaffeEva::ClassifyPixels(const unsigned char* pixels, const int* hei, const int* wid, const int imgCnt, const int topk,
                              std::vector<int>* pClsIdxLst, std::vector<float>* pClsProbLst) {
  if (!net || !preproc) { errorMsg = "ClassifyPixels: LoadCaffePara() and SetPreproc() must have succeeded"; return false; }
  pClsIdxLst->clear();
  pClsProbLst->clear();
  if (imgCnt < 1 || topk < 1) return true;
  std::vector<long long> off(imgCnt);
  size_t total = 0;
  for (int i = 0; i < imgCnt; i++) { off[i] = static_cast<long long>(total); total += static_cast<size_t>(hei[i]) * wid[i] * 3; }
  const int outLen = qcnn_net_out_len(net);
  const size_t imgLen = static_cast<size_t>(caffeParaObj.imgChnIn) * caffeParaObj.imgHeiIn * caffeParaObj.imgWidIn;
  void *dPix = nullptr, *dImg = nullptr, *dProb = nullptr, *dIdx = nullptr, *dVal = nullptr;
  bool ok = qcnn_dev_alloc(ctx, total, &dPix) == 0 && qcnn_dev_alloc(ctx, sizeof(float) * imgCnt * imgLen, &dImg) == 0 &&
            qcnn_dev_alloc(ctx, sizeof(float) * imgCnt * outLen, &dProb) == 0 &&
            qcnn_dev_alloc(ctx, sizeof(int) * imgCnt * topk, &dIdx) == 0 && qcnn_dev_alloc(ctx, sizeof(float) * imgCnt * topk, &dVal) == 0;
  pClsIdxLst->assign(static_cast<size_t>(imgCnt) * topk, 0);
  pClsProbLst->assign(static_cast<size_t>(imgCnt) * topk, 0.0f);
  ok = ok && qcnn_copy_h2d(ctx, dPix, pixels, total, nullptr) == 0 &&
       qcnn_preproc_run(preproc, static_cast<const uint8_t*>(dPix), off.data(), hei, wid, imgCnt, static_cast<float*>(dImg), nullptr) == 0 &&
       qcnn_net_forward(net, static_cast<const float*>(dImg), imgCnt, static_cast<float*>(dProb), nullptr, nullptr) == 0 &&
       qcnn_topk(ctx, static_cast<const float*>(dProb), imgCnt, outLen, topk, 0, static_cast<int*>(dIdx), static_cast<float*>(dVal), nullptr) == 0 &&
       qcnn_copy_d2h(ctx, pClsIdxLst->data(), dIdx, sizeof(int) * imgCnt * topk, nullptr) == 0 &&
       qcnn_copy_d2h(ctx, pClsProbLst->data(), dVal, sizeof(float) * imgCnt * topk, nullptr) == 0 && qcnn_stream_sync(ctx, nullptr) == 0;
  if (!ok) Fail("ClassifyPixels");
  else AccumulateTimes();
  qcnn_dev_free(ctx, dPix); qcnn_dev_free(ctx, dImg); qcnn_dev_free(ctx, dProb); qcnn_dev_free(ctx, dIdx); qcnn_dev_free(ctx, dVal);
  return ok;
}

void CaffeEva::AccumulateTimes(void) {
  for (int l = 0; l < caffeParaObj.layerCnt; l++) {
    float ms = 0.0f;
    if (qcnn_net_layer_time_ms(net, l, &ms) == 0) {
      msIndvLayerLst[l] += ms;
      msAllLayers += ms;
    }
  }
}

// reference src/CaffeEva.cc:213-261: imgDataIn [N,C,H,W] (NCHW) -> probabilities [N * classes]
void CaffeEva::ExecForwardPass(const Matrix<float>& imgDataIn, Matrix<float>* pProbVecOut) {
  if (!net) { printf("[ERROR] LoadCaffePara() has not succeeded\n"); return; }
  const int dataCnt = imgDataIn.GetDimLen(0);
  const int outLen = qcnn_net_out_len(net);
  pProbVecOut->Resize(dataCnt * outLen);
  if (multi) {   // batch sharded over the GPUs of this process
    if (qcnn_multi_forward_h(multi, imgDataIn.GetDataPtr(), dataCnt, pProbVecOut->GetDataPtr()) != 0) Fail("qcnn_multi_forward_h");
    return;
  }
  if (qcnn_net_forward_h(net, imgDataIn.GetDataPtr(), dataCnt, pProbVecOut->GetDataPtr(), nullptr) != 0) {
    Fail("qcnn_net_forward_h");
    return;
  }
  AccumulateTimes();
}

// reference src/CaffeEva.cc:151-211: forward pass over the loaded dataset + top-5 labels
void CaffeEva::ExecForwardPass(void) {
  if (!net) { printf("[ERROR] LoadCaffePara() has not succeeded\n"); return; }
  const int dataCnt = std::min(dataLst.GetDimLen(0), evalCnt);
  const int outLen = qcnn_net_out_len(net);
  const size_t imgLen = static_cast<size_t>(dataLst.GetDimStp(0));
  lablVecPred.Create(dataLst.GetDimLen(0), kLablCntPerData, 1, 1);
  std::vector<float> prob(static_cast<size_t>(evalBatch) * outLen);
  for (int beg = 0; beg < dataCnt; beg += evalBatch) {
    const int cnt = std::min(evalBatch, dataCnt - beg);
    if (multi) {
      if (qcnn_multi_forward_h(multi, dataLst.GetDataPtr() + beg * imgLen, cnt, prob.data()) != 0) {
        Fail("qcnn_multi_forward_h");
        return;
      }
    } else {
      if (qcnn_net_forward_h(net, dataLst.GetDataPtr() + beg * imgLen, cnt, prob.data(), nullptr) != 0) {
        Fail("qcnn_net_forward_h");
        return;
      }
      AccumulateTimes();
    }
    CvtFeatMapToLablVec(beg, beg + cnt - 1, prob.data(), outLen);
  }
}

// reference src/CaffeEva.cc:1162-1190: k-fold arg-max from a FLT_MIN start, winner zeroed
void CaffeEva::CvtFeatMapToLablVec(const int dataIndL, const int dataIndU, const float* probs, const int probVecLen) {
  std::vector<float> p(probVecLen);
  for (int i = dataIndL; i <= dataIndU; i++) {
    std::copy(probs + static_cast<size_t>(i - dataIndL) * probVecLen,
              probs + static_cast<size_t>(i - dataIndL + 1) * probVecLen, p.begin());
    for (int r = 0; r < kLablCntPerData; r++) {
      float best = FLT_MIN;
      uint16_t bi = 0;
      for (int c = 0; c < probVecLen; c++)
        if (best < p[c]) { best = p[c]; bi = static_cast<uint16_t>(c); }
      p[bi] = 0.0f;
      lablVecPred.SetEleAt(bi, i, r, 0, 0);
    }
  }
}

// reference src/CaffeEva.cc:263-295
void CaffeEva::CalcPredAccu(void) {
  const int dataCnt = std::min(dataLst.GetDimLen(0), evalCnt);
  uint32_t hit[kLablCntPerData] = {0, 0, 0, 0, 0};
  for (int i = 0; i < dataCnt; i++)
    for (int r = 0; r < kLablCntPerData; r++)
      if (lablVecGrth.GetDataPtr()[i] == lablVecPred.GetEleAt(i, r, 0, 0)) hit[r]++;
  for (int r = 1; r < kLablCntPerData; r++) hit[r] += hit[r - 1];
  for (int r = 0; r < kLablCntPerData; r++)
    printf("ACCURACY@%d: %d, %.2f%%\n", r + 1, hit[r], 100.0 * hit[r] / std::max(dataCnt, 1));
}

// reference src/CaffeEva.cc:297-326: prints and resets the timers; returns the total in seconds
float CaffeEva::DispElpsTime(void) {
  const float total = static_cast<float>(msAllLayers * 1e-3);
  double byType[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int l = 0; l < caffeParaObj.layerCnt; l++)
    byType[static_cast<int>(caffeParaObj.layerInfoLst[l].type)] += msIndvLayerLst[l] * 1e-3;
  printf("swAllLayers: %.4f (s)\n", total);
  printf("swConvLayer: %.4f (s)\n", byType[0]);   // includes the fused ReLU epilogue
  printf("swPoolLayer: %.4f (s)\n", byType[1]);
  printf("swFCntLayer: %.4f (s)\n", byType[2]);
  printf("swReLuLayer: %.4f (s)\n", byType[3]);
  printf("swLoRNLayer: %.4f (s)\n", byType[4]);   // includes the fused max-pool that follows it
  printf("swDrptLayer: %.4f (s)\n", byType[5]);
  printf("swSMaxLayer: %.4f (s)\n", byType[6]);
  for (int l = 0; l < caffeParaObj.layerCnt; l++) printf("swIndvLayerLst #%2d: %.4f (s)\n", l + 1, msIndvLayerLst[l] * 1e-3);
  Init(enblAprx);
  return total;
}

// ---- per-layer kernels ---------------------------------------------------------------------------------------
template <typename Op>
void CaffeEva::RunOnDevice(const Matrix<float>& src, Matrix<float>* dst, Op op) {
  void *dSrc = nullptr, *dDst = nullptr;
  const size_t sb = sizeof(float) * src.GetEleCnt(), db = sizeof(float) * dst->GetEleCnt();
  bool ok = qcnn_dev_alloc(ctx, sb, &dSrc) == 0 && qcnn_dev_alloc(ctx, db, &dDst) == 0 &&
            qcnn_copy_h2d(ctx, dSrc, src.GetDataPtr(), sb, nullptr) == 0 &&
            op(static_cast<const float*>(dSrc), static_cast<float*>(dDst)) == 0 &&
            qcnn_copy_d2h(ctx, dst->GetDataPtr(), dDst, db, nullptr) == 0 && qcnn_stream_sync(ctx, nullptr) == 0;
  if (!ok) Fail("CalcFeatMap");
  qcnn_dev_free(ctx, dSrc);
  qcnn_dev_free(ctx, dDst);
}

// dispatcher == reference src/CaffeEva.cc:625-670
void CaffeEva::CalcFeatMap(const Matrix<float>& featMapSrc, const int layerInd, Matrix<float>* pFeatMapDst) {
  switch (caffeParaObj.layerInfoLst[layerInd].type) {
    case ENUM_LyrType::Conv: CalcFeatMap_ConvAprx(featMapSrc, layerInd, pFeatMapDst); break;
    case ENUM_LyrType::Pool: CalcFeatMap_Pool(featMapSrc, layerInd, pFeatMapDst); break;
    case ENUM_LyrType::FCnt: CalcFeatMap_FCntAprx(featMapSrc, layerInd, pFeatMapDst); break;
    case ENUM_LyrType::ReLU: CalcFeatMap_ReLu(featMapSrc, layerInd, pFeatMapDst); break;
    case ENUM_LyrType::LoRN: CalcFeatMap_LoRN(featMapSrc, layerInd, pFeatMapDst); break;
    case ENUM_LyrType::Drpt: CalcFeatMap_Drpt(featMapSrc, layerInd, pFeatMapDst); break;
    case ENUM_LyrType::SMax: CalcFeatMap_SMax(featMapSrc, layerInd, pFeatMapDst); break;
    default: printf("[ERROR] invalid layer type\n");
  }
}

// reference src/CaffeEva.cc:760-868: src [N,Hi,Wi,Cin] NHWC -> dst [N,Ho,Wo,Cout] (bias included, no ReLU)
void CaffeEva::CalcFeatMap_ConvAprx(const Matrix<float>& src, const int layerInd, Matrix<float>* dst) {
  qcnn_layer* L = qcnn_net_pq_layer(net, layerInd);
  int od[3];
  if (!L || qcnn_layer_out_dims(L, od) != 0) { Fail("CalcFeatMap_ConvAprx"); return; }
  const int N = src.GetDimLen(0);
  dst->Resize(N, od[0], od[1], od[2]);
  RunOnDevice(src, dst, [&](const float* s, float* d) {
    qcnn_conv_layer_set_src_nchw(L, 0);  // per-layer entry point takes NHWC like the reference
    const int rc = qcnn_conv_aprx_forward(L, s, N, d, 0, nullptr);
    if (layerInd == 0 && caffeParaObj.layerInfoLst[0].stride > 1) qcnn_conv_layer_set_src_nchw(L, 1);
    return rc;
  });
}

// reference src/CaffeEva.cc:968-1025: src [N, Din] (already NCHW-flattened, as the reference's caller leaves it)
void CaffeEva::CalcFeatMap_FCntAprx(const Matrix<float>& src, const int layerInd, Matrix<float>* dst) {
  qcnn_layer* L = qcnn_net_pq_layer(net, layerInd);
  int od[3];
  if (!L || qcnn_layer_out_dims(L, od) != 0) { Fail("CalcFeatMap_FCntAprx"); return; }
  const int N = src.GetDimLen(0);
  dst->Resize(N, 1, 1, od[2]);
  // the net folds the NHWC->NCHW permute into this layer's addressing; this entry point takes the flat vector
  RunOnDevice(src, dst, [&](const float* s, float* d) { return qcnn_fc_aprx_forward_flat(L, s, N, d, 0, nullptr); });
}

void CaffeEva::CalcFeatMap_Pool(const Matrix<float>& src, const int layerInd, Matrix<float>* dst) {
  const LayerInfo& li = caffeParaObj.layerInfoLst[layerInd];
  const int N = src.GetDimLen(0), H = src.GetDimLen(1), W = src.GetDimLen(2), C = src.GetDimLen(3);
  const int Ho = (H + 2 * li.padSiz - li.knlSiz + li.stride - 1) / li.stride + 1;
  const int Wo = (W + 2 * li.padSiz - li.knlSiz + li.stride - 1) / li.stride + 1;
  dst->Resize(N, Ho, Wo, C);
  RunOnDevice(src, dst, [&](const float* s, float* d) {
    return qcnn_maxpool(ctx, s, d, N, H, W, C, li.knlSiz, li.padSiz, li.stride, nullptr);
  });
}

void CaffeEva::CalcFeatMap_ReLu(const Matrix<float>& src, const int, Matrix<float>* dst) {
  *dst = src;
  RunOnDevice(src, dst, [&](const float* s, float* d) { return qcnn_relu(ctx, s, d, src.GetEleCnt(), nullptr); });
}

void CaffeEva::CalcFeatMap_LoRN(const Matrix<float>& src, const int layerInd, Matrix<float>* dst) {
  const LayerInfo& li = caffeParaObj.layerInfoLst[layerInd];
  *dst = src;
  const int C = src.GetDimLen(src.GetDimCnt() - 1);
  RunOnDevice(src, dst, [&](const float* s, float* d) {
    return qcnn_lrn(ctx, s, d, src.GetEleCnt() / C, C, li.lrnSiz, li.lrnAlp, li.lrnBet, li.lrnIni, nullptr);
  });
}

// reference src/CaffeEva.cc:1091-1096: identity at test time
void CaffeEva::CalcFeatMap_Drpt(const Matrix<float>& src, const int, Matrix<float>* dst) { *dst = src; }

void CaffeEva::CalcFeatMap_SMax(const Matrix<float>& src, const int, Matrix<float>* dst) {
  *dst = src;
  const int N = src.GetDimLen(0);
  RunOnDevice(src, dst, [&](const float* s, float* d) { return qcnn_softmax(ctx, s, d, N, src.GetEleCnt() / N, nullptr); });
}
