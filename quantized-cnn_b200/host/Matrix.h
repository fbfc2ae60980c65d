// Host tensor type kept at the API boundary: same public interface as the reference's Matrix<T>
// (/root/reference/include/Matrix.h:18-97: owning, dense, row-major, 1..4 dims), re-implemented from scratch
// around a dims[] array and std::vector storage.  Not a GPU target ("KEPT").
#ifndef QCNN_HOST_MATRIX_H_
#define QCNN_HOST_MATRIX_H_

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

const int kMatDimCntMax = 4;

template <typename T>
class Matrix {
 public:
  Matrix() : rank_(0) { SetDims(0, nullptr); }
  explicit Matrix(const int m) { const int d[1] = {m}; Create(1, d); }
  Matrix(const int m, const int n) { const int d[2] = {m, n}; Create(2, d); }
  Matrix(const int m, const int n, const int p) { const int d[3] = {m, n, p}; Create(3, d); }
  Matrix(const int m, const int n, const int p, const int q) { const int d[4] = {m, n, p, q}; Create(4, d); }
  Matrix(const int dimCnt, const int* dimLenLst) { Create(dimCnt, dimLenLst); }
  // value semantics (the reference hand-writes a copy-ctor with a static scratch array, Matrix.h:108-119;
  // std::vector gives the same behaviour re-entrantly)
  Matrix(const Matrix<T>&) = default;
  Matrix<T>& operator=(const Matrix<T>&) = default;

  void Create(const int m) { const int d[1] = {m}; Create(1, d); }
  void Create(const int m, const int n) { const int d[2] = {m, n}; Create(2, d); }
  void Create(const int m, const int n, const int p) { const int d[3] = {m, n, p}; Create(3, d); }
  void Create(const int m, const int n, const int p, const int q) { const int d[4] = {m, n, p, q}; Create(4, d); }
  void Create(const int dimCnt, const int* dimLenLst) {
    SetDims(dimCnt, dimLenLst);
    buf_.assign(static_cast<size_t>(GetEleCnt()), T());
  }
  void Destroy() { buf_.clear(); buf_.shrink_to_fit(); SetDims(0, nullptr); }

  T* GetDataPtr() const { return const_cast<T*>(buf_.data()); }
  T* GetDataPtr(const int im) const { return GetDataPtr() + Offset(im, 0, 0, 0); }
  T* GetDataPtr(const int im, const int in) const { return GetDataPtr() + Offset(im, in, 0, 0); }
  T* GetDataPtr(const int im, const int in, const int ip) const { return GetDataPtr() + Offset(im, in, ip, 0); }
  T* GetDataPtr(const int im, const int in, const int ip, const int iq) const {
    return GetDataPtr() + Offset(im, in, ip, iq);
  }

  int GetDimCnt() const { return rank_; }
  int GetDimLen(const int dimIdx) const {
    if (dimIdx < 0 || dimIdx >= kMatDimCntMax) {
      printf("[ERROR] invalid index of dimension: %d\n", dimIdx);
      return -1;
    }
    return dims_[dimIdx];
  }
  // elements spanned by one step along dimIdx (product of the trailing dims)
  int GetDimStp(const int dimIdx) const {
    int stp = 1;
    for (int i = rank_ - 1; i > dimIdx; i--) stp *= dims_[i];
    return stp;
  }
  int GetEleCnt() const {
    if (rank_ == 0) return 0;
    int cnt = 1;
    for (int i = 0; i < rank_; i++) cnt *= dims_[i];
    return cnt;
  }
  void DispSizInfo() const {
    printf("[INFO] matrix size:");
    for (int i = 0; i < rank_; i++) printf(i == 0 ? " %d" : " x %d", dims_[i]);
    printf("\n");
  }

  void SetEleAt(const T val, const int im) { buf_[Offset(im, 0, 0, 0)] = val; }
  void SetEleAt(const T val, const int im, const int in) { buf_[Offset(im, in, 0, 0)] = val; }
  void SetEleAt(const T val, const int im, const int in, const int ip) { buf_[Offset(im, in, ip, 0)] = val; }
  void SetEleAt(const T val, const int im, const int in, const int ip, const int iq) {
    buf_[Offset(im, in, ip, iq)] = val;
  }
  T GetEleAt(const int im) const { return buf_[Offset(im, 0, 0, 0)]; }
  T GetEleAt(const int im, const int in) const { return buf_[Offset(im, in, 0, 0)]; }
  T GetEleAt(const int im, const int in, const int ip) const { return buf_[Offset(im, in, ip, 0)]; }
  T GetEleAt(const int im, const int in, const int ip, const int iq) const { return buf_[Offset(im, in, ip, iq)]; }

  // Resize == reshape when the element count is unchanged, else re-create (reference Matrix.h:381-426)
  void Resize(const int m) { const int d[1] = {m}; Reshape(1, d); }
  void Resize(const int m, const int n) { const int d[2] = {m, n}; Reshape(2, d); }
  void Resize(const int m, const int n, const int p) { const int d[3] = {m, n, p}; Reshape(3, d); }
  void Resize(const int m, const int n, const int p, const int q) { const int d[4] = {m, n, p, q}; Reshape(4, d); }

  // new dim i := old dim perm[i]; data physically re-ordered (reference Matrix.h:429-553)
  void Permute(const int mSdx, const int nSdx) { const int p[2] = {mSdx, nSdx}; PermuteN(2, p); }
  void Permute(const int mSdx, const int nSdx, const int pSdx) { const int p[3] = {mSdx, nSdx, pSdx}; PermuteN(3, p); }
  void Permute(const int mSdx, const int nSdx, const int pSdx, const int qSdx) {
    const int p[4] = {mSdx, nSdx, pSdx, qSdx};
    PermuteN(4, p);
  }

  // copies the window of *this starting at the given indices into pMatDst (whose dims define the window);
  // parts of the window outside *this are zero (reference Matrix.h:556-650)
  void GetSubMat(const int imBeg, Matrix<T>* pMatDst) const { const int b[4] = {imBeg, 0, 0, 0}; SubMatN(b, pMatDst); }
  void GetSubMat(const int imBeg, const int inBeg, Matrix<T>* pMatDst) const {
    const int b[4] = {imBeg, inBeg, 0, 0};
    SubMatN(b, pMatDst);
  }
  void GetSubMat(const int imBeg, const int inBeg, const int ipBeg, Matrix<T>* pMatDst) const {
    const int b[4] = {imBeg, inBeg, ipBeg, 0};
    SubMatN(b, pMatDst);
  }
  void GetSubMat(const int imBeg, const int inBeg, const int ipBeg, const int iqBeg, Matrix<T>* pMatDst) const {
    const int b[4] = {imBeg, inBeg, ipBeg, iqBeg};
    SubMatN(b, pMatDst);
  }

 private:
  int rank_;
  int dims_[kMatDimCntMax];
  std::vector<T> buf_;

  void SetDims(const int dimCnt, const int* dimLenLst) {
    rank_ = dimCnt;
    for (int i = 0; i < kMatDimCntMax; i++) dims_[i] = (i < dimCnt) ? dimLenLst[i] : 1;
  }
  size_t Offset(const int im, const int in, const int ip, const int iq) const {
    const int idx[4] = {im, in, ip, iq};
    size_t off = 0;
    for (int i = 0; i < rank_; i++) off = off * dims_[i] + idx[i];
    return off;
  }
  void Reshape(const int dimCnt, const int* d) {
    long cnt = 1;
    for (int i = 0; i < dimCnt; i++) cnt *= d[i];
    if (cnt != GetEleCnt()) Create(dimCnt, d);
    else SetDims(dimCnt, d);
  }
  void PermuteN(const int n, const int* perm) {
    int nd[kMatDimCntMax] = {1, 1, 1, 1}, ostp[kMatDimCntMax] = {0, 0, 0, 0};
    for (int i = 0; i < n; i++) { nd[i] = dims_[perm[i]]; ostp[i] = GetDimStp(perm[i]); }
    std::vector<T> out(buf_.size());
    size_t o = 0;
    for (int a = 0; a < nd[0]; a++)
      for (int b = 0; b < nd[1]; b++)
        for (int c = 0; c < nd[2]; c++)
          for (int d = 0; d < nd[3]; d++)
            out[o++] = buf_[static_cast<size_t>(a) * ostp[0] + static_cast<size_t>(b) * ostp[1] +
                            static_cast<size_t>(c) * ostp[2] + static_cast<size_t>(d) * ostp[3]];
    buf_.swap(out);
    SetDims(n, nd);
  }
  void SubMatN(const int* beg, Matrix<T>* dst) const {
    std::fill(dst->buf_.begin(), dst->buf_.end(), T());
    int lo[4], hi[4];
    for (int i = 0; i < 4; i++) {
      lo[i] = std::max(0, -beg[i]);
      hi[i] = std::min(dst->dims_[i] - 1, dims_[i] - 1 - beg[i]);
      if (hi[i] < lo[i]) return;
    }
    for (int a = lo[0]; a <= hi[0]; a++)
      for (int b = lo[1]; b <= hi[1]; b++)
        for (int c = lo[2]; c <= hi[2]; c++) {
          const int idxS[4] = {a + beg[0], b + beg[1], c + beg[2], lo[3] + beg[3]};
          const int idxD[4] = {a, b, c, lo[3]};
          size_t os = 0, od = 0;
          for (int i = 0; i < 4; i++) { os = os * dims_[i] + idxS[i]; od = od * dst->dims_[i] + idxD[i]; }
          std::copy(buf_.begin() + os, buf_.begin() + os + (hi[3] - lo[3] + 1), dst->buf_.begin() + od);
        }
  }
};

#endif  // QCNN_HOST_MATRIX_H_
