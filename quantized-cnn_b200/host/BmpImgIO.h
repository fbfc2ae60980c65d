// BMP -> network input tensor, same interface as the reference's BmpImgIO (/root/reference/include/BmpImgIO.h):
// decode a 24-bpp BMP to BGR CHW float, bilinear-resize, mean-subtract, centre-crop.  Host-side preprocessing that
// sits immediately BEFORE the PQ path; own 24-bit BMP decoder instead of the vendored
// bitmap_image.hpp.
#ifndef QCNN_HOST_BMPIMGIO_H_
#define QCNN_HOST_BMPIMGIO_H_

#include <string>
#include <vector>

#include "Matrix.h"

enum class ENUM_ReszType { Strict, Relaxed };  // Strict: exactly HxW; Relaxed: keep aspect, cover HxW
enum class ENUM_MeanType { Full, Crop };       // mean image has the full / the cropped size

typedef struct {
  ENUM_ReszType reszType;
  ENUM_MeanType meanType;
  int imgHeiFull;
  int imgWidFull;
  int imgHeiCrop;
  int imgWidCrop;
  std::string filePathMean;
} BmpImgIOPara;

class BmpImgIO {
 public:
  bool Init(const BmpImgIOPara& bmpImgIOPara);
  bool Load(const std::string& filePath, Matrix<float>* pImgDataFnal);

  // extensions (not in the reference): the file decode alone -- interleaved B, G, R bytes, top row first -- for callers
  // that run ReszImg / RmMeanImg / CropImg on the GPU (qcnn_preproc_*), and read access to the recipe
  static bool DecodeBmp(const std::string& filePath, std::vector<unsigned char>* pPixels, int* pHei, int* pWid);
  const BmpImgIOPara& GetPara(void) const { return para_; }
  const Matrix<float>& GetMeanImg(void) const { return imgDataMean; }

 private:
  BmpImgIOPara para_;
  Matrix<float> imgDataMean;
  int imgHeiMean;
  int imgWidMean;

  bool LoadBmpImg(const std::string& filePath, Matrix<float>* pImgData);
  void ReszImg(const Matrix<float>& imgDataSrc, Matrix<float>* pImgDataDst, const ENUM_ReszType type,
               const int imgHeiDstPst, const int imgWidDstPst);
  void CropImg(const Matrix<float>& imgDataSrc, Matrix<float>* pImgDataDst, const int imgHeiDst, const int imgWidDst);
  void RmMeanImg(const Matrix<float>& imgDataMean, Matrix<float>* pImgDataProc);
};

#endif  // QCNN_HOST_BMPIMGIO_H_
