// cu_base.h — CuDevice / CuMatrix / CuVector / CuMath / CuRand: the C++ call surface of the reference's
// CuBaseLib (src/CuBaseLib/{cudevice,cumatrix,cuvector,cumath,curand}.h) re-hosted on the C ABI of
// libtnetb200.so.  Same class and method names, argument meaning and error behaviour (exceptions), so that
// component code written against the reference reads the same here.
//
// What differs by design: nothing synchronises after each op (the reference calls cudaThreadSynchronize()
// in cuSafeCall, cucommon.h:13-22); ops are enqueued on the context's stream and the host only blocks in
// CopyTo().  Buffers keep their capacity across Init() calls instead of cudaFree/cudaMallocPitch churn.
#ifndef TNETB200_CU_BASE_H_
#define TNETB200_CU_BASE_H_

#include <algorithm>
#include <map>

#include "tnet_b200.h"
#include "tnet_base.h"

namespace TNet {

#define TNB_CHECK(call)                                                                          \
  do {                                                                                           \
    int rc__ = (call);                                                                           \
    if (rc__ != TNB_OK) {                                                                        \
      std::ostringstream os__;                                                                   \
      os__ << "CUDA ERROR #" << rc__ << " " << __FILE__ << ":" << __LINE__ << " '" << #call << "' " << tnb_last_error(); \
      throw ::TNet::MyException(os__.str());                                                     \
    }                                                                                            \
  } while (0)

/// The device singleton of the reference (cudevice.h:15-73), one PER THREAD: a thread that never starts another one sees exactly the
/// reference's process-wide object; the loader thread and the per-GPU worker threads of bin/TNetCu each get their own context (own
/// streams, own GPU) behind the same Cx().  Lazily creates the TnbContext.
class CuDevice {
 public:
  static CuDevice &Instantiate() { static thread_local CuDevice dev; return dev; }
  /// a new thread starts from the arithmetic mode (and, unless gpu_id >= 0, the GPU) of the thread that created it
  void InheritFrom(CuDevice &parent, int gpu_id = -1) {
    mMath = parent.Math();
    int dev = gpu_id;
    if (dev < 0) TNB_CHECK(tnb_ctx_device(parent.Ctx(), &dev));
    SelectGPU(dev);
  }
  int Device() { int d = -1; TNB_CHECK(tnb_ctx_device(Ctx(), &d)); return d; }
  ~CuDevice() {
    if (mVerbose && mCtx) PrintProfile();
    if (mCtx) tnb_ctx_destroy(mCtx);
  }
  /// --GPUSELECT: must be called before the first device op (TNetCu.cc:242-243)
  void SelectGPU(int gpu_id) {
    if (mCtx) {
      int cur = -1;
      tnb_ctx_device(mCtx, &cur);
      if (cur == gpu_id) return;
      tnb_ctx_destroy(mCtx);
      mCtx = NULL;
    }
    mSelected = gpu_id;
  }
  void Verbose(bool v) { mVerbose = v; }
  TnbContext *Ctx() {
    if (!mCtx) {
      TNB_CHECK(tnb_ctx_create(&mCtx, mSelected));
      if (mMath >= 0) TNB_CHECK(tnb_ctx_set_math(mCtx, mMath));
    }
    return mCtx;
  }
  void SetMath(int mode) { mMath = mode; if (mCtx) TNB_CHECK(tnb_ctx_set_math(mCtx, mode)); }
  int Math() { int m = 0; TNB_CHECK(tnb_ctx_get_math(Ctx(), &m)); return m; }
  void Sync() { TNB_CHECK(tnb_ctx_sync(Ctx())); }
  void AccuProfile(const std::string &key, double time) { mProfile[key] += time; }
  void PrintProfile() {
    std::cout << "[cudevice profile]\n";
    for (std::map<std::string, double>::iterator it = mProfile.begin(); it != mProfile.end(); ++it)
      std::cout << it->first << "\t" << it->second << "s\n";
  }
  std::string GetFreeMemory() {
    size_t fr = 0, tot = 0;
    TNB_CHECK(tnb_ctx_free_memory(Ctx(), &fr, &tot));
    std::ostringstream os;
    os << "free: " << fr / (1024 * 1024) << "M total: " << tot / (1024 * 1024) << "M ratio: " << (double)fr / (double)tot;
    return os.str();
  }
  unsigned long long Launches() { unsigned long long n = 0; TNB_CHECK(tnb_ctx_launch_count(Ctx(), &n)); return n; }

 private:
  CuDevice() : mCtx(NULL), mSelected(-1), mMath(-1), mVerbose(false) {
    const char *e = getenv("TNB_DEVICE");
    if (e) mSelected = atoi(e);
    const char *m = getenv("TNB_MATH");
    if (m) mMath = !strcmp(m, "tf32") ? TNB_MATH_TF32 : (!strcmp(m, "simt") ? TNB_MATH_FP32_SIMT : (!strcmp(m, "bf16") ? TNB_MATH_BF16 : TNB_MATH_3XTF32));
  }
  CuDevice(const CuDevice &);
  TnbContext *mCtx;
  int mSelected, mMath;
  bool mVerbose;
  std::map<std::string, double> mProfile;
};

inline TnbContext *Cx() { return CuDevice::Instantiate().Ctx(); }

template <typename T> class CuVector;

/// Pitched device matrix (reference: cumatrix.h:19-181).  4-byte element types only.
template <typename T>
class CuMatrix {
 public:
  CuMatrix() : mRows(0), mCols(0), mStride(0), mCap(0), mpCUData(NULL), mpTwin(NULL), mTwinCap(0), mTwinStride(0), mTwinValid(false), mExported(false) {}
  CuMatrix(size_t rows, size_t cols)
      : mRows(0), mCols(0), mStride(0), mCap(0), mpCUData(NULL), mpTwin(NULL), mTwinCap(0), mTwinStride(0), mTwinValid(false), mExported(false) {
    Init(rows, cols);
  }
  ~CuMatrix() { Destroy(); }

  size_t Rows() const { return mRows; }
  size_t Cols() const { return mCols; }
  size_t Stride() const { return mStride; }
  TnbMatrixDim Dim() const { TnbMatrixDim d = {(int)mRows, (int)mCols, (int)mStride}; return d; }
  const T *pCUData() const { return mpCUData; }
  T *pCUData() { mTwinValid = false; return mpCUData; }  // a mutable pointer may be written through: the bf16 twin goes stale
  const T *pCURowData(size_t r) const { assert(r < mRows); return mpCUData + r * mStride; }
  T *pCURowData(size_t r) { assert(r < mRows); mTwinValid = false; return mpCUData + r * mStride; }
  size_t MSize() const { return mRows * mStride * sizeof(T); }

  // ---- bf16 twin (TNB_MATH_BF16): a 16-bit copy of this matrix that the tensor-core GEMMs read instead of converting the fp32
  // array on every call.  Every mutating method and every mutable-pointer access marks it stale; Twin() refreshes a stale
  // twin with one conversion kernel; the fused GEMM epilogues that write the fp32 values and their twin together call
  // TwinForWrite() AFTER taking the mutable fp32 pointer.
  const uint16_t *Twin() const {
    AllocTwin();
    if (!mTwinValid) {
      TNB_CHECK(tnb_to_bf16(Cx(), mpTwin, (int)mTwinStride, (const float *)mpCUData, Dim()));
      mTwinValid = true;
    }
    return mpTwin;
  }
  uint16_t *TwinForWrite() { AllocTwin(); mTwinValid = true; return mpTwin; }
  int TwinStride() const { return (int)mTwinStride; }
  bool TwinValid() const { return mTwinValid; }

  /// (re)allocate; contents are zeroed only when the dimensions change (cumatrix.tcc:16-34)
  CuMatrix<T> &Init(size_t rows, size_t cols) {
    static_assert(sizeof(T) == 4, "4-byte elements");
    if (mRows == rows && mCols == cols) return *this;
    mTwinValid = false;
    size_t stride = ((cols + 31) / 32) * 32;
    if (stride == 0) stride = 32;
    size_t need = (rows ? rows : 1) * stride;
    if (need > mCap) {
      if (mExported) Error("CuMatrix::Init would move a buffer that other ranks have mapped (peer-memory data parallel)");
      Destroy();
      void *p = NULL;
      int st = 0;
      TNB_CHECK(tnb_malloc_pitch(Cx(), &p, &st, (int)rows, (int)cols));  // zero-filled
      mpCUData = (T *)p;
      mCap = need;
      assert((size_t)st == stride);
    } else {
      TNB_CHECK(tnb_memset(Cx(), mpCUData, 0, need * sizeof(T)));
    }
    mRows = rows; mCols = cols; mStride = stride;
    return *this;
  }
  /// keep the logical dimensions and the contents, but make the allocation large enough for `rows` rows (the extra rows are
  /// zero): the data-parallel update shards the weight rows over the ranks in equal blocks (tnb_dp_update)
  void ReserveRows(size_t rows) {
    const size_t need = rows * mStride;
    if (need <= mCap) return;
    if (mExported) Error("CuMatrix::ReserveRows would move a buffer that other ranks have mapped (peer-memory data parallel)");
    void *p = NULL;
    int st = 0;
    TNB_CHECK(tnb_malloc_pitch(Cx(), &p, &st, (int)rows, (int)mCols));  // zero-filled
    assert((size_t)st == mStride);
    if (mpCUData) {
      TNB_CHECK(tnb_memcpy(Cx(), p, mpCUData, mRows * mStride * sizeof(T), 2));
      tnb_free(Cx(), mpCUData);
    }
    mpCUData = (T *)p;
    mCap = need;
  }
  /// the allocation has been mapped into other processes (tnb_peer_map): freeing it while a peer still has it open is undefined
  /// (cudaIpcOpenMemHandle), and the peers close at their own pace — it is left to the end of the process instead
  void MarkExported() { mExported = true; }
  /// exchange the buffers of two matrices (no device work)
  void Swap(CuMatrix<T> &o) {
    std::swap(mRows, o.mRows); std::swap(mCols, o.mCols); std::swap(mStride, o.mStride); std::swap(mCap, o.mCap);
    std::swap(mpCUData, o.mpCUData); std::swap(mpTwin, o.mpTwin); std::swap(mTwinCap, o.mTwinCap); std::swap(mTwinStride, o.mTwinStride);
    std::swap(mTwinValid, o.mTwinValid); std::swap(mExported, o.mExported);
  }
  void Destroy() {
    if (mpCUData && !mExported) tnb_free(Cx(), mpCUData);
    mExported = false;
    if (mpTwin) tnb_free(Cx(), mpTwin);
    mpCUData = NULL;
    mpTwin = NULL;
    mRows = mCols = mStride = mCap = mTwinCap = mTwinStride = 0;
    mTwinValid = false;
  }

  CuMatrix<T> &CopyFrom(const CuMatrix<T> &src) {
    Init(src.Rows(), src.Cols());
    mTwinValid = false;

    TNB_CHECK(tnb_memcpy2d(Cx(), mpCUData, mStride * sizeof(T), src.pCUData(), src.Stride() * sizeof(T), src.Cols() * sizeof(T),
                           src.Rows(), 2));
    return *this;
  }
  CuMatrix<T> &CopyFrom(const Matrix<T> &src) {
    Init(src.Rows(), src.Cols());
    mTwinValid = false;

    TNB_CHECK(tnb_memcpy2d(Cx(), mpCUData, mStride * sizeof(T), src.pData(), src.Stride() * sizeof(T), src.Cols() * sizeof(T),
                           src.Rows(), 0));
    CuDevice::Instantiate().Sync();  // pageable source must stay valid until the copy is done
    return *this;
  }
  Matrix<T> &CopyTo(Matrix<T> &dst) const {
    if (dst.Rows() != mRows || dst.Cols() != mCols) dst.Init(mRows, mCols);
    TNB_CHECK(tnb_memcpy2d(Cx(), dst.pData(), dst.Stride() * sizeof(T), mpCUData, mStride * sizeof(T), mCols * sizeof(T), mRows, 1));
    return dst;
  }
  /// rowCnt rows of src starting at srcOri -> this starting at dstOri (cumatrix.tcc:120-143)
  void CopyRows(size_t rowCnt, size_t srcOri, const CuMatrix<T> &src, size_t dstOri) {
    assert(rowCnt + srcOri <= src.Rows());
    assert(rowCnt + dstOri <= Rows());
    assert(Cols() == src.Cols());
    mTwinValid = false;
    TNB_CHECK(tnb_memcpy2d(Cx(), mpCUData + dstOri * mStride, mStride * sizeof(T), src.pCUData() + srcOri * src.Stride(),
                           src.Stride() * sizeof(T), src.Cols() * sizeof(T), rowCnt, 2));
  }
  void CopyCols(size_t colCnt, size_t srcOri, const CuMatrix<T> &src, size_t dstOri) {
    assert(colCnt + srcOri <= src.Cols());
    assert(colCnt + dstOri <= Cols());
    assert(Rows() == src.Rows());
    mTwinValid = false;
    TNB_CHECK(tnb_memcpy2d(Cx(), mpCUData + dstOri, mStride * sizeof(T), src.pCUData() + srcOri, src.Stride() * sizeof(T),
                           colCnt * sizeof(T), Rows(), 2));
  }
  void SetZero() { mTwinValid = false; if (mpCUData) TNB_CHECK(tnb_memset(Cx(), mpCUData, 0, MSize())); }

  // ---- math (float only, as in the reference's specialisations cumatrix.tcc:194-420) ----
  void SetConst(T v) { mTwinValid = false; TNB_CHECK(tnb_set_const(Cx(), mpCUData, v, Dim())); }
  void ApplyLog() { mTwinValid = false; TNB_CHECK(tnb_apply_log(Cx(), mpCUData, Dim())); }
  void ScaleCols(const CuVector<T> &scale);
  void ScaleRows(const CuVector<T> &scale);
  void AddScaled(T alpha, const CuMatrix<T> &A, T beta) {
    assert(A.Rows() == Rows() && A.Cols() == Cols() && A.Stride() == Stride());
    mTwinValid = false;
    TNB_CHECK(tnb_add_scaled(Cx(), alpha, A.pCUData(), beta, mpCUData, Dim()));
  }
  void AddScaledRow(T alpha, const CuVector<T> &row, T beta);
  /// C = alpha*op(A)*op(B) + beta*C  (cumatrix.tcc:335-370)
  void Gemm(char transa, char transb, T alpha, const CuMatrix<T> &A, const CuMatrix<T> &B, T beta) {
    size_t m = (transa == 'T' || transa == 't') ? A.Cols() : A.Rows();
    size_t k = (transa == 'T' || transa == 't') ? A.Rows() : A.Cols();
    size_t n = (transb == 'T' || transb == 't') ? B.Rows() : B.Cols();
    size_t k1 = (transb == 'T' || transb == 't') ? B.Cols() : B.Rows();
    if (m != Rows() || n != Cols() || k != k1) Error("Gemm: non-matching dimensions");
    mTwinValid = false;
    TNB_CHECK(tnb_gemm(Cx(), transa, transb, (int)m, (int)n, (int)k, alpha, A.pCUData(), (int)A.Stride(), B.pCUData(),
                       (int)B.Stride(), beta, mpCUData, (int)mStride));
  }
  void BlasGer(T alpha, const CuVector<T> &x, const CuVector<T> &y);
  void MulElem(const CuMatrix<T> &A) {
    assert(A.Rows() == Rows() && A.Cols() == Cols() && A.Stride() == Stride());
    mTwinValid = false;
    TNB_CHECK(tnb_mul_elem(Cx(), mpCUData, A.pCUData(), Dim()));
  }
  void LogElem() { mTwinValid = false; TNB_CHECK(tnb_log_elem(Cx(), mpCUData, Dim())); }
  void Print() const { Matrix<T> m; CopyTo(m); std::cout << m; }
  void CheckData() const { Matrix<T> m; CopyTo(m); m.CheckData(); }

 private:
  CuMatrix(const CuMatrix<T> &);
  CuMatrix<T> &operator=(const CuMatrix<T> &);
  void AllocTwin() const {
    static_assert(sizeof(T) == 4, "4-byte elements");
    const size_t st = ((mCols + 63) / 64) * 64, need = (mRows ? mRows : 1) * (st ? st : 64);
    if (need > mTwinCap || st != mTwinStride) {
      if (mpTwin) tnb_free(Cx(), mpTwin);
      mpTwin = NULL;
      void *p = NULL;
      int got = 0;
      TNB_CHECK(tnb_malloc_pitch16(Cx(), &p, &got, (int)mRows, (int)mCols));
      mpTwin = (uint16_t *)p;
      mTwinCap = need;
      mTwinStride = (size_t)got;
      mTwinValid = false;
    }
  }
  size_t mRows, mCols, mStride, mCap;
  T *mpCUData;
  mutable uint16_t *mpTwin;
  mutable size_t mTwinCap, mTwinStride;
  mutable bool mTwinValid;
  bool mExported;
};

/// Device vector (reference: cuvector.h:14-85)
template <typename T>
class CuVector {
 public:
  CuVector() : mDim(0), mCap(0), mpCUData(NULL) {}
  explicit CuVector(size_t dim) : mDim(0), mCap(0), mpCUData(NULL) { Init(dim); }
  ~CuVector() { Destroy(); }
  size_t Dim() const { return mDim; }
  const T *pCUData() const { return mpCUData; }
  T *pCUData() { return mpCUData; }
  CuVector<T> &Init(size_t dim) {
    if (dim == mDim) return *this;
    if (dim > mCap) {
      Destroy();
      void *p = NULL;
      TNB_CHECK(tnb_malloc(Cx(), &p, (dim ? dim : 1) * sizeof(T)));
      mpCUData = (T *)p;
      mCap = dim ? dim : 1;
    } else {
      TNB_CHECK(tnb_memset(Cx(), mpCUData, 0, dim * sizeof(T)));
    }
    mDim = dim;
    return *this;
  }
  void Destroy() {
    if (mpCUData) tnb_free(Cx(), mpCUData);
    mpCUData = NULL;
    mDim = mCap = 0;
  }
  CuVector<T> &CopyFrom(const CuVector<T> &src) {
    Init(src.Dim());
    TNB_CHECK(tnb_memcpy(Cx(), mpCUData, src.pCUData(), mDim * sizeof(T), 2));
    return *this;
  }
  CuVector<T> &CopyFrom(const Vector<T> &src) {
    Init(src.Dim());
    TNB_CHECK(tnb_memcpy(Cx(), mpCUData, src.pData(), mDim * sizeof(T), 0));
    CuDevice::Instantiate().Sync();
    return *this;
  }
  Vector<T> &CopyTo(Vector<T> &dst) const {
    if (dst.Dim() != mDim) dst.Init(mDim);
    TNB_CHECK(tnb_memcpy(Cx(), dst.pData(), mpCUData, mDim * sizeof(T), 1));
    return dst;
  }
  void SetZero() { if (mpCUData) TNB_CHECK(tnb_memset(Cx(), mpCUData, 0, mDim * sizeof(T))); }
  void SetConst(T v) { TnbMatrixDim d = {1, (int)mDim, (int)mDim}; TNB_CHECK(tnb_set_const(Cx(), (float *)mpCUData, v, d)); }
  void AddScaled(T alpha, const CuVector<T> &vec, T beta) {
    assert(vec.Dim() == Dim());
    TnbMatrixDim d = {1, (int)mDim, (int)mDim};
    TNB_CHECK(tnb_add_scaled(Cx(), alpha, vec.pCUData(), beta, mpCUData, d));
  }
  /// this = alpha*colsum(mat) + beta*this  (cuvector.tcc:164-191)
  void AddColSum(T alpha, const CuMatrix<T> &mat, T beta) {
    assert(mat.Cols() == Dim());
    TNB_CHECK(tnb_add_col_sum(Cx(), alpha, mat.pCUData(), beta, mpCUData, mat.Dim()));
  }
  void Print() const { Vector<T> v; CopyTo(v); std::cout << v << "\n"; }

 private:
  CuVector(const CuVector<T> &);
  CuVector<T> &operator=(const CuVector<T> &);
  size_t mDim, mCap;
  T *mpCUData;
};

template <typename T>
inline void CuMatrix<T>::ScaleCols(const CuVector<T> &scale) {
  assert(scale.Dim() == Cols());
  mTwinValid = false;
  TNB_CHECK(tnb_scale_cols(Cx(), mpCUData, scale.pCUData(), Dim()));
}
template <typename T>
inline void CuMatrix<T>::ScaleRows(const CuVector<T> &scale) {
  assert(scale.Dim() == Rows());
  mTwinValid = false;
  TNB_CHECK(tnb_scale_rows(Cx(), mpCUData, scale.pCUData(), Dim()));
}
template <typename T>
inline void CuMatrix<T>::AddScaledRow(T alpha, const CuVector<T> &row, T beta) {
  if (row.Dim() != Cols()) {
    std::ostringstream os;
    os << "Non matching dimensions: Cols:" << Cols() << " VectorDim:" << row.Dim();
    Error(os.str());
  }
  mTwinValid = false;
  TNB_CHECK(tnb_add_scaled_row(Cx(), alpha, row.pCUData(), beta, mpCUData, Dim()));
}
template <typename T>
inline void CuMatrix<T>::BlasGer(T alpha, const CuVector<T> &x, const CuVector<T> &y) {
  assert(x.Dim() == Rows() && y.Dim() == Cols());
  mTwinValid = false;
  TNB_CHECK(tnb_ger(Cx(), alpha, x.pCUData(), (int)x.Dim(), y.pCUData(), (int)y.Dim(), mpCUData, Dim()));
}

template <typename T>
inline std::ostream &operator<<(std::ostream &out, const CuMatrix<T> &mat) {
  out << "[CUMATRIX R" << mat.Rows() << " C" << mat.Cols() << " S" << mat.Stride() << " PTR" << (const void *)mat.pCUData() << "]";
  return out;
}

/// Math helpers of the NN training (reference: cumath.h:16-71)
template <typename T>
class CuMath {
 public:
  static void Sigmoid(CuMatrix<T> &Y, const CuMatrix<T> &X) { TNB_CHECK(tnb_sigmoid(Cx(), Y.pCUData(), X.pCUData(), X.Dim())); }
  static void DiffSigmoid(CuMatrix<T> &Eout, const CuMatrix<T> &Ein, const CuMatrix<T> &Y) {
    TNB_CHECK(tnb_diff_sigmoid(Cx(), Eout.pCUData(), Ein.pCUData(), Y.pCUData(), Eout.Dim()));
  }
  static void Softmax(CuMatrix<T> &Y, const CuMatrix<T> &X) { TNB_CHECK(tnb_softmax(Cx(), Y.pCUData(), X.pCUData(), X.Dim())); }
  /// per-band Y[:, i*m:(i+1)*m] = X[:, i*k:(i+1)*k] * block_transf   (cumath.cc:77-113)
  static void BlockLinearity(CuMatrix<T> &Y, const CuMatrix<T> &X, const CuMatrix<T> &block_transf) {
    assert(Y.Rows() == X.Rows());
    assert((X.Cols() % block_transf.Rows()) == 0 && (Y.Cols() % block_transf.Cols()) == 0);
    int blocks = (int)(X.Cols() / block_transf.Rows());
    int m = (int)block_transf.Cols(), k = (int)block_transf.Rows(), n = (int)X.Rows();
    for (int i = 0; i < blocks; i++)
      TNB_CHECK(tnb_gemm(Cx(), 'N', 'N', n, m, k, 1.0f, X.pCUData() + i * k, (int)X.Stride(), block_transf.pCUData(),
                         (int)block_transf.Stride(), 0.0f, Y.pCUData() + i * m, (int)Y.Stride()));
  }
  /// C[:, offC:] = alpha * op(A[:, offA:]) * op(B[:, offB:]) + beta * C[:, offC:] on sub-blocks addressed by a column offset into
  /// the operands' first rows (cumath.cc:210-244).  The extents are those of B clipped to C: k = min(op(B) rows, op(A) cols),
  /// columns = min(op(B) cols, C cols), rows = min(op(A) rows, C rows) — the callers rely on the clipping (<sharedlinearity>
  /// passes the whole input matrix with the weight block's extents).  Upstream's row clip reads `n<C.Rows() ? m : C.Rows()`,
  /// a slip that never triggers in its callers (op(A) always has at least C's rows); the plain minimum is taken here.
  static void OffsetGemm(char transA, char transB, T alpha, const CuMatrix<T> &A, const CuMatrix<T> &B, T beta, CuMatrix<T> &C, int offA,
                         int offB, int offC) {
    const bool ta = (transA == 'T' || transA == 't'), tb = (transB == 'T' || transB == 't');
    size_t cols = tb ? B.Rows() : B.Cols();
    size_t rows = ta ? A.Cols() : A.Rows();
    size_t k = tb ? B.Cols() : B.Rows();
    const size_t k1 = ta ? A.Rows() : A.Cols();
    k = std::min(k, k1);
    cols = std::min(cols, C.Cols());
    rows = std::min(rows, C.Rows());
    TNB_CHECK(tnb_gemm(Cx(), transA, transB, (int)rows, (int)cols, (int)k, alpha, A.pCUData() + offA, (int)A.Stride(), B.pCUData() + offB,
                       (int)B.Stride(), beta, C.pCUData() + offC, (int)C.Stride()));
  }
  /// out = [in in in ...] (cumath.cc:366-384: the splice kernel on a one-row matrix with all-zero offsets)
  static void VecExpand(const CuVector<T> &in, CuVector<T> &out) {
    assert(in.Dim() > 0 && out.Dim() % in.Dim() == 0);
    CuVector<int> offsets(out.Dim() / in.Dim());
    offsets.SetZero();
    VecExpand(in, out, offsets);
  }
  /// the same with a caller-kept all-zero offset vector of out.Dim()/in.Dim() entries (no allocation per call)
  static void VecExpand(const CuVector<T> &in, CuVector<T> &out, const CuVector<int> &offsets) {
    assert(in.Dim() > 0 && out.Dim() == in.Dim() * offsets.Dim());
    TnbMatrixDim din = {1, (int)in.Dim(), (int)in.Dim()}, dout = {1, (int)out.Dim(), (int)out.Dim()};
    TNB_CHECK(tnb_expand(Cx(), out.pCUData(), in.pCUData(), offsets.pCUData(), dout, din));
  }
  /// out = alpha * (sum of the out.Dim()-long pieces of in) + beta * out (cumath.cc:388-404: `in` read as [pieces x out.Dim()])
  static void VecAddColSum(T alpha, const CuVector<T> &in, T beta, CuVector<T> &out) {
    assert(out.Dim() > 0 && in.Dim() % out.Dim() == 0);
    TnbMatrixDim d = {(int)(in.Dim() / out.Dim()), (int)out.Dim(), (int)out.Dim()};
    TNB_CHECK(tnb_add_col_sum(Cx(), alpha, in.pCUData(), beta, out.pCUData(), d));
  }
  static void Expand(CuMatrix<T> &Y, const CuMatrix<T> &X, const CuVector<int> &frameOffsets) {
    assert(Y.Rows() == X.Rows() && X.Cols() * frameOffsets.Dim() == Y.Cols());
    TNB_CHECK(tnb_expand(Cx(), Y.pCUData(), X.pCUData(), frameOffsets.pCUData(), Y.Dim(), X.Dim()));
  }
  static void Rearrange(CuMatrix<T> &Y, const CuMatrix<T> &X, const CuVector<int> &copyFrom) {
    assert(copyFrom.Dim() == Y.Cols() && Y.Rows() == X.Rows());
    TNB_CHECK(tnb_rearrange(Cx(), Y.pCUData(), X.pCUData(), copyFrom.pCUData(), Y.Dim(), X.Dim()));
  }
  static void Randomize(CuMatrix<T> &Y, const CuMatrix<T> &X, const CuVector<int> &copyFrom) {
    assert(X.Cols() == Y.Cols() && X.Rows() == Y.Rows() && copyFrom.Dim() <= Y.Rows());
    TnbMatrixDim dx = X.Dim(), dy = Y.Dim();
    dx.rows = dy.rows = (int)copyFrom.Dim();
    TNB_CHECK(tnb_randomize(Cx(), Y.pCUData(), X.pCUData(), copyFrom.pCUData(), dy, dx));
  }
  static void CheckClass(const CuMatrix<T> &out, const CuMatrix<T> &des, CuVector<int> &match) {
    assert(out.Cols() == des.Cols() && out.Rows() == des.Rows() && out.Stride() == des.Stride() && match.Dim() == out.Rows());
    TNB_CHECK(tnb_check_class(Cx(), out.pCUData(), des.pCUData(), match.pCUData(), out.Dim()));
  }
  static void OffsetGemv(char trans, T alpha, const CuMatrix<T> &A, const T *x, size_t dimX, T beta, T *y, size_t dimY, size_t offsetY) {
    TNB_CHECK(tnb_offset_gemv(Cx(), trans, alpha, A.pCUData(), A.Dim(), x, (int)dimX, beta, y, (int)dimY, (int)offsetY));
  }
  static void BlasGer(T alpha, const T *x, size_t dimX, const T *y, size_t dimY, CuMatrix<T> &A) {
    TNB_CHECK(tnb_ger(Cx(), alpha, x, (int)dimX, y, (int)dimY, A.pCUData(), A.Dim()));
  }
};

/// Per-element Hybrid-Taus generator (reference: curand.h:11-32, curand.tcc:13-155)
template <typename T>
class CuRand {
 public:
  CuRand(size_t rows, size_t cols) { SeedGpu(rows, cols); }
  /// consumes 4*rows*cols (+rejections) lrand48() draws, matrix after matrix, row-major (curand.tcc:13-49)
  void SeedGpu(size_t rows, size_t cols) {
    Matrix<unsigned> mat(rows, cols);
    SeedRandom(mat); z1.CopyFrom(mat);
    SeedRandom(mat); z2.CopyFrom(mat);
    SeedRandom(mat); z3.CopyFrom(mat);
    SeedRandom(mat); z4.CopyFrom(mat);
    tmp.Init(rows, cols);
  }
  void Rand(CuMatrix<T> &tgt) {
    tgt.Init(z1.Rows(), z1.Cols());
    TNB_CHECK(tnb_rand(Cx(), tgt.pCUData(), z1.pCUData(), z2.pCUData(), z3.pCUData(), z4.pCUData(), tgt.Dim()));
  }
  void GaussRand(CuMatrix<T> &tgt) {
    tgt.Init(z1.Rows(), z1.Cols());
    TNB_CHECK(tnb_gauss_rand(Cx(), tgt.pCUData(), z1.pCUData(), z2.pCUData(), z3.pCUData(), z4.pCUData(), tgt.Dim()));
  }
  /// states = probs > rand ? 1 : 0 — fused, no tmp round trip (curand.tcc:136-155)
  void BinarizeProbs(const CuMatrix<T> &probs, CuMatrix<T> &states) {
    if (probs.Rows() != z1.Rows() || probs.Cols() != z1.Cols()) Error("Non matching dims!!");
    states.Init(z1.Rows(), z1.Cols());
    TNB_CHECK(tnb_rand_binarize(Cx(), states.pCUData(), probs.pCUData(), z1.pCUData(), z2.pCUData(), z3.pCUData(), z4.pCUData(),
                                states.Dim()));
  }
  void AddGaussNoise(CuMatrix<T> &tgt, T gscale = 1.0) {
    if (tgt.Rows() != z1.Rows() || tgt.Cols() != z1.Cols()) Error("Non matching dims!!");
    TNB_CHECK(tnb_add_gauss_noise(Cx(), tgt.pCUData(), gscale, z1.pCUData(), z2.pCUData(), z3.pCUData(), z4.pCUData(), tgt.Dim()));
  }

 private:
  static void SeedRandom(Matrix<unsigned> &mat) {
    for (size_t j = 0; j < mat.Rows(); j++)
      for (size_t i = 0; i < mat.Cols(); i++) {
        unsigned value = 0;
        while (value <= 128) value = (unsigned)lrand48();
        mat(j, i) = value;
      }
  }
  CuMatrix<unsigned> z1, z2, z3, z4;
  CuMatrix<T> tmp;
};

}  // namespace TNet
#endif
