// See CaffeEvaWrapper.h.  Behaviour follows /root/reference/src/CaffeEvaWrapper.cc (cited per function).
#include "CaffeEvaWrapper.h"

#include <cstdio>
#include <fstream>

CaffeEvaWrapper::CaffeEvaWrapper(void) {}

// reference src/CaffeEvaWrapper.cc:15-42
bool CaffeEvaWrapper::SetPath(const std::string& mainDirPathSrc, const std::string& clsNameFilePath,
                              const std::string& imgLablFilePath) {
  mainDirPath = mainDirPathSrc;
  if (!LoadClsName(clsNameFilePath)) {
    errorMsg = "[CaffeEvaWrapper::SetPath] could not open file: " + clsNameFilePath;
    return false;
  }
  if (!imgLablFilePath.empty() && !LoadImgLabl(imgLablFilePath)) {
    errorMsg = "[CaffeEvaWrapper::SetPath] could not open file: " + imgLablFilePath;
    return false;
  }
  return true;
}

// reference src/CaffeEvaWrapper.cc:44-151: per-model preprocessing recipe + parameter location
bool CaffeEvaWrapper::SetModel(const ENUM_CaffeModel& caffeModelSrc, const ENUM_CompMethod& compMethodSrc) {
  struct Recipe { const char* name; const char* pfx; ENUM_ReszType resz; ENUM_MeanType mean; int crop; };
  Recipe r;
  switch (caffeModelSrc) {
    case ENUM_CaffeModel::AlexNet:     r = {"AlexNet", "bvlc_alexnet_aCaF", ENUM_ReszType::Strict, ENUM_MeanType::Full, 227}; break;
    case ENUM_CaffeModel::CaffeNet:    r = {"CaffeNet", "bvlc_caffenet_aCaF", ENUM_ReszType::Strict, ENUM_MeanType::Full, 227}; break;
    case ENUM_CaffeModel::CaffeNetFGB: r = {"CaffeNetFGB", "bvlc_caffenetfgb_aCaF", ENUM_ReszType::Strict, ENUM_MeanType::Full, 227}; break;
    case ENUM_CaffeModel::CaffeNetFGD: r = {"CaffeNetFGD", "bvlc_caffenetfgd_aCaF", ENUM_ReszType::Strict, ENUM_MeanType::Full, 227}; break;
    case ENUM_CaffeModel::VggCnnS:     r = {"VggCnnS", "vgg_cnn_s_aCaF", ENUM_ReszType::Relaxed, ENUM_MeanType::Crop, 224}; break;
    case ENUM_CaffeModel::VGG16:
      printf("[FATAL ERROR] VGG-16 is not supported (for now)\n");
      errorMsg = "[CaffeEvaWrapper::SetModel] unsupported caffe model name";
      return false;
    default:
      errorMsg = "[CaffeEvaWrapper::SetModel] unrecognized caffe model name";
      return false;
  }
  BmpImgIOPara p;
  p.reszType = r.resz;
  p.meanType = r.mean;
  p.imgHeiFull = p.imgWidFull = 256;
  p.imgHeiCrop = p.imgWidCrop = r.crop;
  p.filePathMean = mainDirPath + "/" + r.name + "/imagenet_mean.single.bin";
  if (!bmpImgIOObj.Init(p)) {
    errorMsg = "[CaffeEvaWrapper::SetModel] could not open the mean image file";
    return false;
  }
  caffeEvaObj.Init(compMethodSrc == ENUM_CompMethod::Aprx);
  caffeEvaObj.SetModelName(r.name);
  caffeEvaObj.SetModelPath(mainDirPath + "/" + r.name + "/Bin.Files", r.pfx);
  if (!caffeEvaObj.LoadCaffePara()) {
    errorMsg = "[CaffeEvaWrapper::SetModel] could not load model files";
    return false;
  }
  // the recipe above, on the device: Proc ships the decoded pixels and gets the top-k back
  if (!caffeEvaObj.SetPreproc(r.resz == ENUM_ReszType::Strict ? 0 : 1, r.mean == ENUM_MeanType::Full ? 0 : 1, p.imgHeiFull,
                              p.imgWidFull, p.imgHeiCrop, p.imgWidCrop, bmpImgIOObj.GetMeanImg())) {
    errorMsg = "[CaffeEvaWrapper::SetModel] could not set up the device preprocessing: " + caffeEvaObj.GetErrorMsg();
    return false;
  }
  return true;
}

// reference src/CaffeEvaWrapper.cc:153-209
bool CaffeEvaWrapper::Proc(const std::string& filePathProcImg, CaffeEvaRslt* pCaffeEvaRslt) {
  // BMP decode on the host; resize, mean subtraction, crop (BmpImgIO::Load, :40-71), the forward pass and the top-k by
  // repeated arg-max (first maximum wins, winner zeroed, :188-206) on the device
  std::vector<unsigned char> pixels;
  int hei = 0, wid = 0;
  if (!BmpImgIO::DecodeBmp(filePathProcImg, &pixels, &hei, &wid)) {
    errorMsg = "[CaffeEvaWrapper::Proc] could open the BMP file";
    return false;
  }
  std::vector<int> topIdx;
  std::vector<float> topProb;
  if (!caffeEvaObj.ClassifyPixels(pixels.data(), &hei, &wid, 1, pCaffeEvaRslt->clsCntPred, &topIdx, &topProb)) {
    errorMsg = "[CaffeEvaWrapper::Proc] " + caffeEvaObj.GetErrorMsg();
    return false;
  }
  pCaffeEvaRslt->timeTotal = caffeEvaObj.DispElpsTime();

  const std::string fileName = ExtrFileName(filePathProcImg);
  pCaffeEvaRslt->hasGrthClsName = false;
  for (const GrthEntry& e : clsNameGrthLst) {
    if (e.fileName == fileName) {
      pCaffeEvaRslt->hasGrthClsName = true;
      pCaffeEvaRslt->clsNameGrth = e.clsName;
      break;
    }
  }
  pCaffeEvaRslt->clsIdxLst.clear();
  pCaffeEvaRslt->clsProbLst.clear();
  pCaffeEvaRslt->clsNameLst.clear();
  for (size_t rank = 0; rank < topIdx.size(); rank++) {
    const int best = topIdx[rank];
    pCaffeEvaRslt->clsIdxLst.push_back(best);
    pCaffeEvaRslt->clsProbLst.push_back(topProb[rank]);
    pCaffeEvaRslt->clsNameLst.push_back(best < static_cast<int>(clsNameLst.size()) ? clsNameLst[best] : std::string());
  }
  return true;
}

std::string CaffeEvaWrapper::GetErrorMsg(void) { return errorMsg; }
void CaffeEvaWrapper::ClrErrorMsg(void) { errorMsg = ""; }

// one class name per line (reference :219-249)
bool CaffeEvaWrapper::LoadClsName(const std::string& filePath) {
  std::ifstream in(filePath);
  if (!in) return false;
  clsNameLst.clear();
  for (std::string line; std::getline(in, line);) clsNameLst.push_back(line);
  return true;
}

// "<file> <0-based class index>" per line (reference :251-284)
bool CaffeEvaWrapper::LoadImgLabl(const std::string& filePath) {
  std::ifstream in(filePath);
  if (!in) return false;
  clsNameGrthLst.clear();
  std::string name;
  int idx;
  while (in >> name >> idx) {
    if (idx < 0 || idx >= static_cast<int>(clsNameLst.size())) continue;
    clsNameGrthLst.push_back({ExtrFileName(name), clsNameLst[idx]});
  }
  return true;
}

// file name without directory and extension (reference :286-320)
std::string CaffeEvaWrapper::ExtrFileName(const std::string& filePath) {
  const size_t slash = filePath.find_last_of("/\\");
  const size_t beg = (slash == std::string::npos) ? 0 : slash + 1;
  const size_t dot = filePath.find_last_of('.');
  const size_t end = (dot == std::string::npos || dot < beg) ? filePath.size() : dot;
  return filePath.substr(beg, end - beg);
}
