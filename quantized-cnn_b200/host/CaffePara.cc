// See CaffePara.h.  Behaviour follows /root/reference/src/CaffePara.cc (cited per function).
#include "CaffePara.h"

#include <cstdio>
#include <cstring>

#include "FileIO.h"

namespace {

// Tiny builder so an architecture reads as one chain: Net(...).C(pad,k,cnt,grp,stride).R().L(...).P(...)...
class Net {
 public:
  explicit Net(CaffePara* p, int chn, int hei, int wid) : p_(p) {
    p_->imgChnIn = chn;
    p_->imgHeiIn = hei;
    p_->imgWidIn = wid;
    p_->layerInfoLst.clear();
  }
  ~Net() { p_->layerCnt = static_cast<int>(p_->layerInfoLst.size()); }
  Net& C(int pad, int k, int cnt, int grp, int stride) {
    LayerInfo& li = Add(ENUM_LyrType::Conv);
    li.padSiz = pad; li.knlSiz = k; li.knlCnt = cnt; li.grpCnt = grp; li.stride = stride;
    return *this;
  }
  Net& P(int pad, int k, int stride) {
    LayerInfo& li = Add(ENUM_LyrType::Pool);
    li.padSiz = pad; li.knlSiz = k; li.stride = stride;
    return *this;
  }
  Net& F(int nod) { Add(ENUM_LyrType::FCnt).nodCnt = nod; return *this; }
  Net& R() { Add(ENUM_LyrType::ReLU); return *this; }
  Net& L(int siz, float alp, float bet, float ini) {
    LayerInfo& li = Add(ENUM_LyrType::LoRN);
    li.lrnSiz = siz; li.lrnAlp = alp; li.lrnBet = bet; li.lrnIni = ini;
    return *this;
  }
  Net& D(float rat) { Add(ENUM_LyrType::Drpt).drpRat = rat; return *this; }
  Net& S() { Add(ENUM_LyrType::SMax); return *this; }
  // conv -> relu pairs, the VGG idiom
  Net& CR(int cnt, int times) { for (int i = 0; i < times; i++) C(1, 3, cnt, 1, 1).R(); return *this; }
  // fc -> relu -> dropout
  Net& FRD(int nod, float rat) { return F(nod).R().D(rat); }

 private:
  LayerInfo& Add(ENUM_LyrType t) {
    LayerInfo li;
    memset(&li, 0, sizeof(li));
    li.type = t;
    p_->layerInfoLst.push_back(li);
    return p_->layerInfoLst.back();
  }
  CaffePara* p_;
};

// CaffeNet and its fine-tuned variants differ only in dropout ratio and class count
void CaffeNetFamily(CaffePara* p, float drp, int classes) {
  Net(p, 3, 227, 227)
      .C(0, 11, 96, 1, 4).R().P(0, 3, 2).L(5, 0.0001f, 0.75f, 1.0f)
      .C(2, 5, 256, 2, 1).R().P(0, 3, 2).L(5, 0.0001f, 0.75f, 1.0f)
      .C(1, 3, 384, 1, 1).R().C(1, 3, 384, 2, 1).R().C(1, 3, 256, 2, 1).R().P(0, 3, 2)
      .FRD(4096, drp).FRD(4096, drp).F(classes).S();
}

}  // namespace

void CaffePara::Init(const std::string& dirPathSrc, const std::string& filePfxSrc) {
  dirPath = dirPathSrc;
  filePfx = filePfxSrc;
}

// reference src/CaffePara.cc:20-52 (23 layers; LRN before pooling)
void CaffePara::ConfigLayer_AlexNet(void) {
  Net(this, 3, 227, 227)
      .C(0, 11, 96, 1, 4).R().L(5, 0.0001f, 0.75f, 1.0f).P(0, 3, 2)
      .C(2, 5, 256, 2, 1).R().L(5, 0.0001f, 0.75f, 1.0f).P(0, 3, 2)
      .C(1, 3, 384, 1, 1).R().C(1, 3, 384, 2, 1).R().C(1, 3, 256, 2, 1).R().P(0, 3, 2)
      .FRD(4096, 0.50f).FRD(4096, 0.50f).F(1000).S();
}
// reference src/CaffePara.cc:54-86 (pooling before LRN)
void CaffePara::ConfigLayer_CaffeNet(void) { CaffeNetFamily(this, 0.50f, 1000); }
// reference src/CaffePara.cc:88-119
void CaffePara::ConfigLayer_VggCnnS(void) {
  Net(this, 3, 224, 224)
      .C(0, 7, 96, 1, 2).R().L(5, 0.0005f, 0.75f, 2.0f).P(0, 3, 3)
      .C(1, 5, 256, 1, 1).R().P(0, 2, 2)
      .C(1, 3, 512, 1, 1).R().C(1, 3, 512, 1, 1).R().C(1, 3, 512, 1, 1).R().P(0, 3, 3)
      .FRD(4096, 0.50f).FRD(4096, 0.50f).F(1000).S();
}
// reference src/CaffePara.cc:121-169
void CaffePara::ConfigLayer_VGG16(void) {
  Net(this, 3, 224, 224)
      .CR(64, 2).P(0, 2, 2).CR(128, 2).P(0, 2, 2).CR(256, 3).P(0, 2, 2).CR(512, 3).P(0, 2, 2).CR(512, 3).P(0, 2, 2)
      .FRD(4096, 0.50f).FRD(4096, 0.50f).F(1000).S();
}
// reference src/CaffePara.cc:171-203 / 205-237
void CaffePara::ConfigLayer_CaffeNetFGB(void) { CaffeNetFamily(this, 0.70f, 518); }
void CaffePara::ConfigLayer_CaffeNetFGD(void) { CaffeNetFamily(this, 0.50f, 200); }

bool CaffePara::ConfigLayer_ByName(const std::string& modelName) {
  if (modelName == "AlexNet") ConfigLayer_AlexNet();
  else if (modelName == "CaffeNet") ConfigLayer_CaffeNet();
  else if (modelName == "VggCnnS") ConfigLayer_VggCnnS();
  else if (modelName == "VGG16") ConfigLayer_VGG16();
  else if (modelName == "CaffeNetFGB") ConfigLayer_CaffeNetFGB();
  else if (modelName == "CaffeNetFGD") ConfigLayer_CaffeNetFGD();
  else return false;
  return true;
}

std::string CaffePara::ParaPath(const char* kind, int layerInd, const char* ext) const {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s/%s.%s.%02d.%s", dirPath.c_str(), filePfx.c_str(), kind, layerInd + 1, ext);
  return std::string(buf);
}

// reference src/CaffePara.cc:239-306.  File index = layerInd + 1; assignments become 0-based (:285-288).
bool CaffePara::LoadLayerPara(const bool enblAprx, const ENUM_AsmtEnc asmtEnc) {
  bool ok = true;
  layerParaLst.clear();
  layerParaLst.resize(layerCnt);
  for (int l = 0; l < layerCnt; l++) {
    const ENUM_LyrType t = layerInfoLst[l].type;
    if (t != ENUM_LyrType::Conv && t != ENUM_LyrType::FCnt) continue;
    LayerPara& para = layerParaLst[l];
    ok &= FileIO::ReadBinFile(ParaPath("biasVec", l, "bin"), &para.biasVec);
    if (enblAprx) {
      ok &= FileIO::ReadBinFile(ParaPath("ctrdLst", l, "bin"), &para.ctrdLst);
      if (asmtEnc == ENUM_AsmtEnc::Raw) ok &= FileIO::ReadBinFile(ParaPath("asmtLst", l, "bin"), &para.asmtLst);
      else ok &= FileIO::ReadCbnFile(ParaPath("asmtLst", l, "cbn"), &para.asmtLst);
      uint8_t* a = para.asmtLst.GetDataPtr();
      for (int i = 0, n = para.asmtLst.GetEleCnt(); i < n; i++) a[i] = static_cast<uint8_t>(a[i] - 1);
    } else if (t == ENUM_LyrType::Conv) {
      ok &= FileIO::ReadBinFile(ParaPath("convKnl", l, "bin"), &para.convKnlLst);
    } else {
      ok &= FileIO::ReadBinFile(ParaPath("fcntWei", l, "bin"), &para.fcntWeiMat);
    }
  }
  return ok;
}

// reference src/CaffePara.cc:308-358: Raw (.bin, 1-based u8) <-> Compact (.cbn) for every conv/FC layer
bool CaffePara::CvtAsmtEnc(const ENUM_AsmtEnc asmtEncSrc, const ENUM_AsmtEnc asmtEncDst) {
  if (asmtEncSrc == asmtEncDst) {
    printf("[INFO] no encoding conversion is required\n");
    return true;
  }
  bool ok = true;
  Matrix<uint8_t> asmt;
  for (int l = 0; l < layerCnt; l++) {
    const ENUM_LyrType t = layerInfoLst[l].type;
    if (t != ENUM_LyrType::Conv && t != ENUM_LyrType::FCnt) continue;
    const std::string binPath = ParaPath("asmtLst", l, "bin");
    const std::string cbnPath = ParaPath("asmtLst", l, "cbn");
    if (asmtEncSrc == ENUM_AsmtEnc::Raw) {
      ok &= FileIO::ReadBinFile(binPath, &asmt);
      const int bits = CalcBitCntPerEle(asmt);
      printf("layer #%d: bitCntPerEle = %d\n", l + 1, bits);
      ok &= FileIO::WriteCbnFile(cbnPath, asmt, bits);
    } else {
      ok &= FileIO::ReadCbnFile(cbnPath, &asmt);
      ok &= FileIO::WriteBinFile(binPath, asmt);
    }
  }
  return ok;
}

// reference src/CaffePara.cc:360-378: bits needed for (max 1-based value - 1)
int CaffePara::CalcBitCntPerEle(const Matrix<uint8_t>& asmtLst) {
  uint8_t mx = 0;
  const uint8_t* a = asmtLst.GetDataPtr();
  for (int i = 0, n = asmtLst.GetEleCnt(); i < n; i++) mx = std::max(mx, a[i]);
  mx = static_cast<uint8_t>(mx - 1);
  int bits = 0;
  for (; mx != 0; mx /= 2) bits++;
  return bits;
}
