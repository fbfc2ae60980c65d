// See BmpImgIO.h.  Behaviour follows /root/reference/src/BmpImgIO.cc (cited per function).
#include "BmpImgIO.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "FileIO.h"

namespace {
const int kImgChn = 3;

uint32_t Le32(const uint8_t* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24); }
uint16_t Le16(const uint8_t* p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }
}  // namespace

bool BmpImgIO::Init(const BmpImgIOPara& bmpImgIOPara) {
  para_ = bmpImgIOPara;
  if (!FileIO::ReadBinFile(para_.filePathMean, &imgDataMean)) return false;  // [3][H][W] BGR
  imgHeiMean = imgDataMean.GetDimLen(1);
  imgWidMean = imgDataMean.GetDimLen(2);
  return true;
}

// reference src/BmpImgIO.cc:40-71
bool BmpImgIO::Load(const std::string& filePath, Matrix<float>* pImgDataFnal) {
  Matrix<float> orgn, full;
  if (!LoadBmpImg(filePath, &orgn)) return false;
  ReszImg(orgn, &full, para_.reszType, para_.imgHeiFull, para_.imgWidFull);
  if (para_.meanType == ENUM_MeanType::Full) {
    RmMeanImg(imgDataMean, &full);
    CropImg(full, pImgDataFnal, para_.imgHeiCrop, para_.imgWidCrop);
  } else {
    CropImg(full, pImgDataFnal, para_.imgHeiCrop, para_.imgWidCrop);
    RmMeanImg(imgDataMean, pImgDataFnal);
  }
  return true;
}

// 24-bpp uncompressed BMP -> interleaved B, G, R bytes (the file's own channel order), row 0 = top of the picture
bool BmpImgIO::DecodeBmp(const std::string& filePath, std::vector<unsigned char>* pPixels, int* pHei, int* pWid) {
  FILE* f = fopen(filePath.c_str(), "rb");
  if (f == nullptr) {
    printf("[ERROR] cannot open the BMP image at %s\n", filePath.c_str());
    return false;
  }
  uint8_t hdr[54];
  bool ok = fread(hdr, 1, 54, f) == 54 && hdr[0] == 'B' && hdr[1] == 'M';
  const uint32_t dataOff = ok ? Le32(hdr + 10) : 0;
  const int32_t wid = ok ? static_cast<int32_t>(Le32(hdr + 18)) : 0;
  const int32_t heiRaw = ok ? static_cast<int32_t>(Le32(hdr + 22)) : 0;
  ok = ok && Le16(hdr + 28) == 24 && Le32(hdr + 30) == 0 && wid > 0 && heiRaw != 0;
  if (!ok) {
    printf("[ERROR] unsupported BMP (need 24-bpp uncompressed) at %s\n", filePath.c_str());
    fclose(f);
    return false;
  }
  const bool bottomUp = heiRaw > 0;
  const int hei = bottomUp ? heiRaw : -heiRaw;
  const size_t rowBytes = (static_cast<size_t>(wid) * 3 + 3) & ~static_cast<size_t>(3);
  std::vector<uint8_t> row(rowBytes);
  pPixels->resize(static_cast<size_t>(hei) * wid * 3);
  fseek(f, dataOff, SEEK_SET);
  for (int r = 0; r < hei; r++) {
    if (fread(row.data(), 1, rowBytes, f) != rowBytes) {
      printf("[ERROR] truncated BMP at %s\n", filePath.c_str());
      fclose(f);
      return false;
    }
    const int y = bottomUp ? hei - 1 - r : r;
    std::copy(row.begin(), row.begin() + static_cast<size_t>(wid) * 3, pPixels->begin() + static_cast<size_t>(y) * wid * 3);
  }
  fclose(f);
  *pHei = hei;
  *pWid = wid;
  return true;
}

// 24-bpp uncompressed BMP -> [1][3][H][W] in B, G, R plane order (reference src/BmpImgIO.cc:73-103 stores
// B at channel 0, G at 1, R at 2; row 0 is the top of the picture)
bool BmpImgIO::LoadBmpImg(const std::string& filePath, Matrix<float>* pImgData) {
  std::vector<unsigned char> pix;
  int hei = 0, wid = 0;
  if (!DecodeBmp(filePath, &pix, &hei, &wid)) return false;
  pImgData->Create(1, kImgChn, hei, wid);
  for (int y = 0; y < hei; y++)
    for (int x = 0; x < wid; x++)
      for (int c = 0; c < kImgChn; c++) pImgData->SetEleAt(pix[(static_cast<size_t>(y) * wid + x) * 3 + c], 0, c, y, x);
  return true;
}

// reference src/BmpImgIO.cc:105-178: scale = (src-1)/(dst-1), weight-normalised bilinear interpolation
void BmpImgIO::ReszImg(const Matrix<float>& src, Matrix<float>* pDst, const ENUM_ReszType type, const int heiPst,
                       const int widPst) {
  const int hs = src.GetDimLen(2), ws = src.GetDimLen(3);
  float sh = static_cast<float>(hs - 1) / (heiPst - 1);
  float sw = static_cast<float>(ws - 1) / (widPst - 1);
  int hd = heiPst, wd = widPst;
  if (type == ENUM_ReszType::Relaxed) {
    sh = std::min(sh, sw);
    sw = std::min(sh, sw);
    hd = static_cast<int>((hs - 1) / sh + 0.0000001) + 1;
    wd = static_cast<int>((ws - 1) / sw + 0.0000001) + 1;
  }
  pDst->Resize(1, kImgChn, hd, wd);
  for (int y = 0; y < hd; y++) {
    const float yc = sh * y;
    const int yl = std::max(0, static_cast<int>(yc));
    const int yh = std::min(hs - 1, yl + 1);
    const float wyl = 1.0 - (yc - yl), wyh = 1.0 - (yh - yc);
    for (int x = 0; x < wd; x++) {
      const float xc = sw * x;
      const int xl = std::max(0, static_cast<int>(xc));
      const int xh = std::min(ws - 1, xl + 1);
      const float wxl = 1.0 - (xc - xl), wxh = 1.0 - (xh - xc);
      const float wLT = wyl * wxl, wRT = wyl * wxh, wLB = wyh * wxl, wRB = wyh * wxh;
      const float wSum = wLT + wRT + wLB + wRB;
      for (int c = 0; c < kImgChn; c++) {
        const float v = src.GetEleAt(0, c, yl, xl) * wLT + src.GetEleAt(0, c, yl, xh) * wRT +
                        src.GetEleAt(0, c, yh, xl) * wLB + src.GetEleAt(0, c, yh, xh) * wRB;
        pDst->SetEleAt(v / wSum, 0, c, y, x);
      }
    }
  }
}

// centre crop (reference src/BmpImgIO.cc:180-201)
void BmpImgIO::CropImg(const Matrix<float>& src, Matrix<float>* pDst, const int heiDst, const int widDst) {
  const int yo = (src.GetDimLen(2) - heiDst) / 2, xo = (src.GetDimLen(3) - widDst) / 2;
  pDst->Resize(1, kImgChn, heiDst, widDst);
  for (int c = 0; c < kImgChn; c++)
    for (int y = 0; y < heiDst; y++)
      std::copy(src.GetDataPtr(0, c, y + yo, xo), src.GetDataPtr(0, c, y + yo, xo) + widDst, pDst->GetDataPtr(0, c, y, 0));
}

// reference src/BmpImgIO.cc:203-224
void BmpImgIO::RmMeanImg(const Matrix<float>& mean, Matrix<float>* pProc) {
  if (pProc->GetDimLen(2) != imgHeiMean || pProc->GetDimLen(3) != imgWidMean) {
    printf("[ERROR] mismatch in the image size\n");
    return;
  }
  float* p = pProc->GetDataPtr();
  const float* m = mean.GetDataPtr();
  for (int i = 0, n = pProc->GetEleCnt(); i < n; i++) p[i] -= m[i];
}
