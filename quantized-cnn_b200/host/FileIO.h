// On-disk formats of the reference, kept bit-compatible (/root/reference/include/FileIO.h):
//   .bin  int32 dimCnt | int32 dimLen[dimCnt] | raw T elements, row-major          (ReadBinFile  :56-107)
//   .cbn  int32 dimCnt | int32 dimLen[]       | int32 bitCntPerEle | 4096-byte blocks, each holding
//         floor(32768/bits) values packed MSB-first; block tails unused            (ReadCbnFile  :110-178)
// Same static-template interface as the reference's FileIO class (FileIO.h:23-51); fresh implementation
// (whole-block bit cursor instead of the reference's byte-walking state machine).  Like the reference reader,
// ReadCbnFile returns MATLAB-style 1-based values (FileIO.h:165); CaffePara::LoadLayerPara subtracts the 1.
#ifndef QCNN_HOST_FILEIO_H_
#define QCNN_HOST_FILEIO_H_

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "Matrix.h"

class FileIO {
 public:
  template <typename T>
  static bool ReadBinFile(const std::string& filePath, Matrix<T>* pDataLst);
  template <typename T>
  static bool ReadCbnFile(const std::string& filePath, Matrix<T>* pDataLst);
  template <typename T>
  static bool WriteBinFile(const std::string& filePath, const Matrix<T>& dataLst);
  template <typename T>
  static bool WriteCbnFile(const std::string& filePath, const Matrix<T>& dataLst, const int bitCntPerEle);
  // bit width stored in a .cbn header (0 if unreadable)
  static int PeekCbnBits(const std::string& filePath);
  // set false to silence the [INFO] lines (the reference always prints them)
  static bool& Verbose() { static bool v = false; return v; }

 private:
  static const int kBlockBytes = 4096;
  struct Closer {
    FILE* f;
    ~Closer() { if (f) fclose(f); }
  };
  static bool ReadHeader(FILE* f, int* dimCnt, int* dims) {
    int32_t dc = 0;
    if (fread(&dc, sizeof(int32_t), 1, f) != 1 || dc < 1 || dc > kMatDimCntMax) return false;
    int32_t dl[kMatDimCntMax];
    if (fread(dl, sizeof(int32_t), dc, f) != static_cast<size_t>(dc)) return false;
    for (int i = 0; i < dc; i++) {
      if (dl[i] < 0) return false;
      dims[i] = dl[i];
    }
    *dimCnt = dc;
    return true;
  }
  template <typename T>
  static bool WriteHeader(FILE* f, const Matrix<T>& m) {
    int32_t dc = m.GetDimCnt();
    if (fwrite(&dc, sizeof(int32_t), 1, f) != 1) return false;
    for (int i = 0; i < dc; i++) {
      int32_t dl = m.GetDimLen(i);
      if (fwrite(&dl, sizeof(int32_t), 1, f) != 1) return false;
    }
    return true;
  }
};

template <typename T>
bool FileIO::ReadBinFile(const std::string& filePath, Matrix<T>* pDataLst) {
  Closer c{fopen(filePath.c_str(), "rb")};
  if (c.f == nullptr) {
    printf("[ERROR] could not open file at %s\n", filePath.c_str());
    return false;
  }
  int dimCnt, dims[kMatDimCntMax];
  if (!ReadHeader(c.f, &dimCnt, dims)) {
    printf("[ERROR] malformed header in %s\n", filePath.c_str());
    return false;
  }
  pDataLst->Create(dimCnt, dims);
  const size_t n = static_cast<size_t>(pDataLst->GetEleCnt());
  if (fread(pDataLst->GetDataPtr(), sizeof(T), n, c.f) != n) {
    printf("[ERROR] truncated data in %s\n", filePath.c_str());
    return false;
  }
  if (Verbose()) { printf("[INFO] read %s: ", filePath.c_str()); pDataLst->DispSizInfo(); }
  return true;
}

template <typename T>
bool FileIO::ReadCbnFile(const std::string& filePath, Matrix<T>* pDataLst) {
  Closer c{fopen(filePath.c_str(), "rb")};
  if (c.f == nullptr) {
    printf("[ERROR] could not open file at %s\n", filePath.c_str());
    return false;
  }
  int dimCnt, dims[kMatDimCntMax];
  int32_t bits = 0;
  if (!ReadHeader(c.f, &dimCnt, dims) || fread(&bits, sizeof(int32_t), 1, c.f) != 1 || bits < 1 || bits > 8) {
    printf("[ERROR] malformed header in %s\n", filePath.c_str());
    return false;
  }
  pDataLst->Create(dimCnt, dims);
  const long total = pDataLst->GetEleCnt();
  const long perBlock = static_cast<long>(kBlockBytes) * 8 / bits;
  T* out = pDataLst->GetDataPtr();
  uint8_t block[kBlockBytes];
  for (long base = 0; base < total; base += perBlock) {
    if (fread(block, 1, kBlockBytes, c.f) != static_cast<size_t>(kBlockBytes)) {
      printf("[ERROR] truncated block in %s\n", filePath.c_str());
      return false;
    }
    const long cnt = std::min(perBlock, total - base);
    // bit cursor over the block: keep up to 32 unread bits in `window`
    uint32_t window = 0;
    int have = 0;
    int bytePos = 0;
    const uint32_t mask = (1u << bits) - 1u;
    for (long i = 0; i < cnt; i++) {
      while (have < bits) {
        window = (window << 8) | block[bytePos++];
        have += 8;
      }
      have -= bits;
      const uint32_t v = (window >> have) & mask;
      out[base + i] = static_cast<T>(v + 1);  // 1-based, as the reference reader returns it
    }
  }
  if (Verbose()) { printf("[INFO] read %s (%d bits): ", filePath.c_str(), bits); pDataLst->DispSizInfo(); }
  return true;
}

template <typename T>
bool FileIO::WriteBinFile(const std::string& filePath, const Matrix<T>& dataLst) {
  Closer c{fopen(filePath.c_str(), "wb")};
  if (c.f == nullptr) {
    printf("[ERROR] could not open file at %s\n", filePath.c_str());
    return false;
  }
  const size_t n = static_cast<size_t>(dataLst.GetEleCnt());
  return WriteHeader(c.f, dataLst) && fwrite(dataLst.GetDataPtr(), sizeof(T), n, c.f) == n;
}

// input values are 1-based (value - 1 is stored), exactly like the reference writer (FileIO.h:281-350)
template <typename T>
bool FileIO::WriteCbnFile(const std::string& filePath, const Matrix<T>& dataLst, const int bitCntPerEle) {
  if (bitCntPerEle < 1 || bitCntPerEle > 8) return false;
  Closer c{fopen(filePath.c_str(), "wb")};
  if (c.f == nullptr) {
    printf("[ERROR] could not open file at %s\n", filePath.c_str());
    return false;
  }
  int32_t bits = bitCntPerEle;
  if (!WriteHeader(c.f, dataLst) || fwrite(&bits, sizeof(int32_t), 1, c.f) != 1) return false;
  const long total = dataLst.GetEleCnt();
  const long perBlock = static_cast<long>(kBlockBytes) * 8 / bits;
  const T* in = dataLst.GetDataPtr();
  const uint32_t mask = (1u << bits) - 1u;
  uint8_t block[kBlockBytes];
  for (long base = 0; base < total; base += perBlock) {
    std::fill(block, block + kBlockBytes, 0);
    const long cnt = std::min(perBlock, total - base);
    uint64_t window = 0;  // pending bits, right-aligned
    int have = 0;
    int bytePos = 0;
    for (long i = 0; i < cnt; i++) {
      const uint32_t v = (static_cast<uint32_t>(in[base + i]) - 1u) & mask;
      window = (window << bits) | v;
      have += bits;
      while (have >= 8) {
        have -= 8;
        block[bytePos++] = static_cast<uint8_t>((window >> have) & 0xFFu);
      }
    }
    if (have > 0) block[bytePos] = static_cast<uint8_t>((window << (8 - have)) & 0xFFu);
    if (fwrite(block, 1, kBlockBytes, c.f) != static_cast<size_t>(kBlockBytes)) return false;
  }
  return true;
}

inline int FileIO::PeekCbnBits(const std::string& filePath) {
  Closer c{fopen(filePath.c_str(), "rb")};
  if (c.f == nullptr) return 0;
  int dimCnt, dims[kMatDimCntMax];
  int32_t bits = 0;
  if (!ReadHeader(c.f, &dimCnt, dims) || fread(&bits, sizeof(int32_t), 1, c.f) != 1) return 0;
  return bits;
}

#endif  // QCNN_HOST_FILEIO_H_
