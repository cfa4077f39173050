// cu_nnet.h — the reference's CuTNetLib call surface (CuComponent / CuUpdatableComponent, the layer classes,
// CuNetwork, CuCache, CuObjectiveFunction) hosted on libtnetb200.so.
//
// Reference files mirrored (paths under src/CuTNetLib): cuComponent.h:27-175, cuBiasedLinearity.{h,cc},
// cuActivation.{h,cc}, cuCRBEDctFeat.h, cuRbm.{h,cc}, cuRecurrent.{h,cc}, cuNetwork.{h,cc}, cuCache.{h,cc},
// cuObjectiveFunction.{h,cc}.  Class names, virtuals, ownership (a component owns mOutput/mErrorOutput, its
// inputs are borrowed pointers) and error behaviour (exceptions) are the reference's.
//
// B200-first differences, all inside CuNetwork's traversal and invisible to component-level callers:
//   * <biasedlinearity> followed by <sigmoid> runs as ONE tcgen05 GEMM whose epilogue adds the bias and applies the
//     sigmoid, writing the sigmoid component's output directly (the affine pre-activation is never materialised);
//   * dX of an affine layer whose predecessor is a <sigmoid> multiplies by y(1-y) in the GEMM epilogue;
//   * the weight gradient GEMM applies momentum / learning rate / L2 in its epilogue (single GPU), or is followed by
//     one kernel that sums the ranks' gradients over NVLink peer memory, updates this rank's rows and hands them to every
//     rank (data parallel; NCCL all-reduce / reduce-scatter schedules selectable);
//   * <softmax>'s identity backward copy and the output copy of Propagate() are pointer re-wiring, not copies;
//   * Xent / correct / frames accumulate on the device and are read when Report()/GetError() is called.
// SetFusion(false) (or TNB_FUSE=0) restores the component-by-component traversal for debugging and parity tests.
#ifndef TNETB200_CU_NNET_H_
#define TNETB200_CU_NNET_H_

#include <algorithm>
#include <list>

#include "cu_base.h"

namespace TNet {

// =====================================================================================================
// CuComponent / CuUpdatableComponent
// =====================================================================================================
class CuComponent {
 public:
  typedef enum {
    UPDATABLE_COMPONENT = 0x0100, BIASED_LINEARITY, DISCRETE_LINEARITY, SHARED_LINEARITY, SPARSE_LINEARITY, RBM, RBM_SPARSE, RECURRENT,
    ACT_FUN = 0x0200, SOFTMAX, SIGMOID,
    OTHER = 0x0400, EXPAND, COPY, TRANSPOSE, BLOCK_LINEARITY, WINDOW, BIAS, LOG, BLOCK_ARRAY, CLUSTER_LINEARITY
  } ComponentType;

  CuComponent(size_t nInputs, size_t nOutputs, CuComponent *pPred)
      : mNInputs(nInputs), mNOutputs(nOutputs), mpInput(NULL), mpErrorInput(NULL) {
    if (pPred != NULL) {  // double link with the predecessor
      SetInput(pPred->GetOutput());
      pPred->SetErrorInput(GetErrorOutput());
    }
  }
  virtual ~CuComponent() {}

  virtual ComponentType GetType() const = 0;
  virtual const char *GetName() const = 0;
  virtual bool IsUpdatable() const { return false; }

  size_t GetNInputs() const { return mNInputs; }
  size_t GetNOutputs() const { return mNOutputs; }

  const CuMatrix<BaseFloat> &GetInput() const { if (NULL == mpInput) Error("mpInput is NULL"); return *mpInput; }
  const CuMatrix<BaseFloat> &GetOutput() const { return mOutput; }
  const CuMatrix<BaseFloat> &GetErrorInput() const { if (NULL == mpErrorInput) Error("mpErrorInput is NULL"); return *mpErrorInput; }
  const CuMatrix<BaseFloat> &GetErrorOutput() const { return mErrorOutput; }

  void SetInput(const CuMatrix<BaseFloat> &rInput) { mpInput = &rInput; }
  void SetErrorInput(const CuMatrix<BaseFloat> &rErrorInput) { mpErrorInput = &rErrorInput; }

  /// forward pass Input -> Output (cuComponent.h:205-218)
  void Propagate() {
    mOutput.Init(GetInput().Rows(), GetNOutputs());
    if (GetNInputs() != GetInput().Cols())
      KALDI_ERR << "Non-matching INPUT dim!!! Network dim: " << GetNInputs() << " Data dim: " << GetInput().Cols();
    PropagateFnc(GetInput(), mOutput);
  }
  /// backward pass ErrorInput -> ErrorOutput (cuComponent.h:221-235)
  void Backpropagate() {
    mErrorOutput.Init(GetErrorInput().Rows(), GetNInputs());
    if (GetErrorInput().Cols() != mNOutputs) Error("Backpropagate: non-matching error dim");
    BackpropagateFnc(GetErrorInput(), mErrorOutput);
  }

  virtual void ReadFromStream(std::istream &) {}
  virtual void WriteToStream(std::ostream &) {}

  /// network-internal: lets CuNetwork write a fused result straight into this component's buffers
  CuMatrix<BaseFloat> &MutableOutput() { return mOutput; }
  CuMatrix<BaseFloat> &MutableErrorOutput() { return mErrorOutput; }

 protected:
  virtual void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) = 0;
  virtual void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) = 0;

  size_t mNInputs, mNOutputs;
  const CuMatrix<BaseFloat> *mpInput;       ///< NOT owned
  const CuMatrix<BaseFloat> *mpErrorInput;  ///< NOT owned
  CuMatrix<BaseFloat> mOutput;              ///< owned
  CuMatrix<BaseFloat> mErrorOutput;         ///< owned
};

class CuUpdatableComponent : public CuComponent {
 public:
  CuUpdatableComponent(size_t nInputs, size_t nOutputs, CuComponent *pPred)
      : CuComponent(nInputs, nOutputs, pPred), mLearningRate(0.0), mMomentum(0), mWeightcost(0), mGradDivFrm(true) {}
  virtual bool IsUpdatable() const { return true; }
  virtual void Update() = 0;
  void LearnRate(BaseFloat rate) { mLearningRate = rate; }
  BaseFloat LearnRate() { return mLearningRate; }
  void Momentum(BaseFloat mmt) { mMomentum = mmt; }
  BaseFloat Momentum() { return mMomentum; }
  void Weightcost(BaseFloat cost) { mWeightcost = cost; }
  BaseFloat Weightcost() { return mWeightcost; }
  void GradDivFrm(bool div) { mGradDivFrm = div; }
  bool GradDivFrm() { return mGradDivFrm; }

 protected:
  BaseFloat mLearningRate, mMomentum, mWeightcost;
  bool mGradDivFrm;
};

// =====================================================================================================
// Activations (cuActivation.{h,cc})
// =====================================================================================================
class CuSigmoid : public CuComponent {
 public:
  CuSigmoid(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuComponent(nInputs, nOutputs, pPred) {}
  ComponentType GetType() const { return SIGMOID; }
  const char *GetName() const { return "<sigmoid>"; }
 protected:
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { CuMath<BaseFloat>::Sigmoid(Y, X); }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { CuMath<BaseFloat>::DiffSigmoid(Y, X, mOutput); }
};

class CuSoftmax : public CuComponent {
 public:
  CuSoftmax(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuComponent(nInputs, nOutputs, pPred) {}
  ComponentType GetType() const { return SOFTMAX; }
  const char *GetName() const { return "<softmax>"; }
 protected:
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { CuMath<BaseFloat>::Softmax(Y, X); }
  /// X is already dE/d(softmax input) (cuActivation.cc:35-41)
  void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { Y.CopyFrom(X); }
};

// =====================================================================================================
// CuBiasedLinearity (cuBiasedLinearity.{h,cc})
// =====================================================================================================
class CuBiasedLinearity : public CuUpdatableComponent {
 public:
  CuBiasedLinearity(size_t nInputs, size_t nOutputs, CuComponent *pPred)
      : CuUpdatableComponent(nInputs, nOutputs, pPred), mLinearity(nInputs, nOutputs), mBias(nOutputs),
        mLinearityCorrection(nInputs, nOutputs), mBiasCorrection(nOutputs), mDpFrames(0), mRowsPad(0), mEvE(NULL), mEvB(NULL), mEvAR(NULL), mEvDone(NULL), mEvG(NULL), mEvPush(NULL), mPushMode(-1), mDpPending(false), mPeerMapped(false) {}
  ~CuBiasedLinearity() {
    if (mEvE) { tnb_event_destroy(Cx(), mEvE); tnb_event_destroy(Cx(), mEvB); tnb_event_destroy(Cx(), mEvAR); tnb_event_destroy(Cx(), mEvDone); }
    if (mEvG) { tnb_event_destroy(Cx(), mEvG); tnb_event_destroy(Cx(), mEvPush); }
    if (mPeerMapped) { tnb_peer_unmap(Cx(), mGradPeers); tnb_peer_unmap(Cx(), mWPeers); }
  }
  ComponentType GetType() const { return BIASED_LINEARITY; }
  const char *GetName() const { return "<biasedlinearity>"; }

  // In TNB_MATH_BF16 every GEMM below reads the bf16 twins of its operands (CuMatrix::Twin(): resident, refreshed only when
  // stale) and writes the twin of its result from the epilogue, so that no conversion pass runs between the layers.
  static bool Bf16() { return CuDevice::Instantiate().Math() == TNB_MATH_BF16; }
  void Forward(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y, int act) {
    WaitDataParallel();
    const CuMatrix<BaseFloat> &W = mLinearity;
    if (Bf16()) {
      const uint16_t *x16 = X.Twin(), *w16 = W.Twin();
      float *y = Y.pCUData();
      uint16_t *y16 = Y.TwinForWrite();
      TNB_CHECK(tnb_affine_fwd_bf16(Cx(), x16, X.TwinStride(), X.Dim(), w16, W.TwinStride(), W.Dim(), mBias.pCUData(), y, Y.Dim(), y16,
                                    Y.TwinStride(), act));
    } else {
      TNB_CHECK(tnb_affine_fwd(Cx(), X.pCUData(), X.Dim(), W.pCUData(), W.Dim(), mBias.pCUData(), Y.pCUData(), Y.Dim(), act));
    }
  }
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { Forward(X, Y, TNB_ACT_NONE); }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { BackpropagateDiffSigmoid(X, NULL, Y); }
  /// bias + GEMM + sigmoid in one kernel; Y is the following <sigmoid>'s output buffer
  void PropagateSigmoid(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { Forward(X, Y, TNB_ACT_SIGMOID); }
  /// dX fused with the diff-sigmoid of the <sigmoid> below: Eprev = (E*W^T) .* Yprev .* (1-Yprev)
  void BackpropagateDiffSigmoid(const CuMatrix<BaseFloat> &E, const CuMatrix<BaseFloat> &Yprev, CuMatrix<BaseFloat> &Eprev) {
    BackpropagateDiffSigmoid(E, &Yprev, Eprev);
  }
  void BackpropagateDiffSigmoid(const CuMatrix<BaseFloat> &E, const CuMatrix<BaseFloat> *pYprev, CuMatrix<BaseFloat> &Eprev) {
    WaitDataParallel();
    const CuMatrix<BaseFloat> &W = mLinearity;
    TnbMatrixDim none = {0, 0, 0};
    const float *yp = pYprev ? pYprev->pCUData() : NULL;
    const TnbMatrixDim dyp = pYprev ? pYprev->Dim() : none;
    if (Bf16()) {
      const uint16_t *e16 = E.Twin(), *w16 = W.Twin();
      float *ep = Eprev.pCUData();
      uint16_t *ep16 = Eprev.TwinForWrite();
      TNB_CHECK(tnb_affine_bwd_dx_bf16(Cx(), e16, E.TwinStride(), E.Dim(), w16, W.TwinStride(), W.Dim(), yp, dyp, ep, Eprev.Dim(), ep16,
                                       Eprev.TwinStride()));
    } else {
      TNB_CHECK(tnb_affine_bwd_dx(Cx(), E.pCUData(), E.Dim(), W.pCUData(), W.Dim(), yp, dyp, Eprev.pCUData(), Eprev.Dim()));
    }
  }
  /// cuBiasedLinearity.cc:44-64, fused into the dW GEMM epilogue
  void Update() { Update(NULL); }
  /// pDeferredBias != NULL: only the weight half runs now; the bias half (which nothing reads before the next forward pass) is
  /// described in *pDeferredBias for CuNetwork to apply together with the other layers' (tnb_bias_update_batch)
  void Update(TnbBiasJob *pDeferredBias) {
    WaitDataParallel();
    const CuMatrix<BaseFloat> &X = GetInput(), &E = GetErrorInput();
    float *bias = pDeferredBias ? NULL : mBias.pCUData(), *corrb = pDeferredBias ? NULL : mBiasCorrection.pCUData();
    if (pDeferredBias) {
      TnbBiasJob j = {E.pCUData(), E.Dim(), mBias.pCUData(), mBiasCorrection.pCUData(), mLearningRate, mMomentum, mGradDivFrm ? 1 : 0, 0};
      *pDeferredBias = j;
    }
    if (Bf16()) {
      const uint16_t *x16 = X.Twin(), *e16 = E.Twin();
      mLinearity.Twin();  // allocate (and fill on first use) so that the epilogue can keep it current from here on
      float *w = mLinearity.pCUData();
      uint16_t *w16 = mLinearity.TwinForWrite();
      TNB_CHECK(tnb_affine_update_bf16(Cx(), x16, X.TwinStride(), X.Dim(), e16, E.TwinStride(), E.pCUData(), E.Dim(), w, mLinearity.Dim(),
                                       w16, mLinearity.TwinStride(), bias, mLinearityCorrection.pCUData(), corrb, mLearningRate, mMomentum,
                                       mWeightcost, mGradDivFrm ? 1 : 0, 0));
      return;
    }
    TNB_CHECK(tnb_affine_update(Cx(), X.pCUData(), X.Dim(), E.pCUData(), E.Dim(), mLinearity.pCUData(), mLinearity.Dim(), bias,
                                mLinearityCorrection.pCUData(), corrb, mLearningRate, mMomentum, mWeightcost, mGradDivFrm ? 1 : 0, 0));
  }
  // ---- the same two GEMMs as JOBS of a batched launch (tnb_gemm_batch: independent GEMMs of the backward pass share one grid) ----
  /// dX (+ diff-sigmoid of the <sigmoid> below when pYprev != NULL) as a job; false if the batch kernel does not take the shape
  bool MakeBackpropJob(const CuMatrix<BaseFloat> &E, const CuMatrix<BaseFloat> *pYprev, CuMatrix<BaseFloat> &Eprev, TnbGemmJob *pJob) {
    WaitDataParallel();
    const CuMatrix<BaseFloat> &W = mLinearity;
    TnbMatrixDim none = {0, 0, 0};
    TNB_CHECK(tnb_job_affine_bwd_dx(pJob, E.pCUData(), E.Dim(), W.pCUData(), W.Dim(), pYprev ? pYprev->pCUData() : NULL,
                                    pYprev ? pYprev->Dim() : none, Eprev.pCUData(), Eprev.Dim()));
    if (Bf16()) {
      const uint16_t *e16 = E.Twin(), *w16 = W.Twin();
      uint16_t *ep16 = Eprev.TwinForWrite();  // (allocates on first use: take it BEFORE TwinStride() is read in the argument list below)
      TNB_CHECK(tnb_job_set_twins(pJob, e16, E.TwinStride(), w16, W.TwinStride(), ep16, Eprev.TwinStride(), NULL, 0));
    }
    if (tnb_gemm_batch_ok(Cx(), pJob)) return true;
    if (Bf16()) (void)Eprev.pCUData();  // not taken: the twin was not written after all
    return false;
  }
  /// the weight half of Update() as a job (the bias half goes to *pDeferredBias as in Update(TnbBiasJob *)); false if not batchable
  bool MakeUpdateJob(TnbGemmJob *pJob, TnbBiasJob *pDeferredBias) {
    WaitDataParallel();
    const CuMatrix<BaseFloat> &X = GetInput(), &E = GetErrorInput();
    const CuMatrix<BaseFloat> &Wc = mLinearity;
    TNB_CHECK(tnb_job_affine_update(pJob, X.pCUData(), X.Dim(), E.pCUData(), E.Dim(), const_cast<float *>(Wc.pCUData()), Wc.Dim(),
                                    const_cast<float *>(static_cast<const CuMatrix<BaseFloat> &>(mLinearityCorrection).pCUData()), mLearningRate,
                                    mMomentum, mWeightcost, mGradDivFrm ? 1 : 0, 0));
    if (Bf16()) {
      const uint16_t *x16 = X.Twin(), *e16 = E.Twin();
      uint16_t *w16 = const_cast<uint16_t *>(Wc.Twin());  // allocate (and fill on first use); the epilogue keeps it current from here on
      TNB_CHECK(tnb_job_set_twins(pJob, x16, X.TwinStride(), e16, E.TwinStride(), NULL, 0, w16, Wc.TwinStride()));
    }
    if (!tnb_gemm_batch_ok(Cx(), pJob)) return false;
    (void)mLinearityCorrection.pCUData();  // both are rewritten by the job ...
    (void)mLinearity.pCUData();
    if (Bf16()) (void)mLinearity.TwinForWrite();  // ... and the epilogue writes the weights' twin from the value it stores
    TnbBiasJob j = {E.pCUData(), E.Dim(), mBias.pCUData(), mBiasCorrection.pCUData(), mLearningRate, mMomentum, mGradDivFrm ? 1 : 0, 0};
    *pDeferredBias = j;
    return true;
  }
  // ---- data-parallel halves of Update(): local gradient, (all-reduce by the network), apply ----
  /// rows of the weight matrix rounded up to a multiple of the world size: the ranks own equal blocks of rows in the update
  void PrepareDataParallel(int world) {
    mRowsPad = ((mNInputs + world - 1) / world) * world;
    mLinearity.ReserveRows(mRowsPad);
    mLinearityCorrection.ReserveRows(mRowsPad);
    mGrad.Init(mRowsPad + 1, mNOutputs);  // [dW (padded rows stay zero) ; db]
  }
  void ComputeGradient(bool with_bias = true) {
    const CuMatrix<BaseFloat> &X = GetInput(), &E = GetErrorInput();
    if (mRowsPad == 0) PrepareDataParallel(1);
    TnbMatrixDim dG = mLinearity.Dim();
    float *gb = with_bias ? mGrad.pCURowData(mRowsPad) : NULL;
    if (Bf16()) {
      const uint16_t *x16 = X.Twin(), *e16 = E.Twin();
      TNB_CHECK(tnb_affine_grad_bf16(Cx(), x16, X.TwinStride(), X.Dim(), e16, E.TwinStride(), E.pCUData(), E.Dim(), mGrad.pCUData(), dG, gb));
      return;
    }
    TNB_CHECK(tnb_affine_grad(Cx(), X.pCUData(), X.Dim(), E.pCUData(), E.Dim(), mGrad.pCUData(), dG, gb));
  }
  /// reduce-scatter + this rank's block of the update + all-gather, on the communication stream (tnb_dp_update)
  void DataParallelUpdate(int n_frames_global) {
    TNB_CHECK(tnb_dp_update(Cx(), mGrad.pCUData(), mLinearity.pCUData(), mLinearityCorrection.pCUData(), mLinearity.Dim(), (int)mRowsPad,
                            mGrad.pCURowData(mRowsPad), mBias.pCUData(), mBiasCorrection.pCUData(), mLearningRate, mMomentum, mWeightcost,
                            mGradDivFrm ? 1 : 0, n_frames_global));
  }
  /// all-reduce schedule: description of this layer's update from the summed gradient, for tnb_sgd_update_batch
  TnbSgdJob GradientJob(int n_frames_global) {
    TnbSgdJob j;
    memset(&j, 0, sizeof(j));
    j.G = mGrad.pCUData(); j.dW = mLinearity.Dim(); j.gb = mGrad.pCURowData(mRowsPad);
    j.corrW = mLinearityCorrection.pCUData(); j.bias = mBias.pCUData(); j.corrb = mBiasCorrection.pCUData();
    j.lr = mLearningRate; j.mmt = mMomentum; j.wc = mWeightcost; j.grad_div_frm = mGradDivFrm ? 1 : 0; j.n_frames = n_frames_global;
    if (Bf16()) mLinearity.Twin();           // allocate / fill once; kept current by the update kernel from here on
    j.W = mLinearity.pCUData();               // (marks the twin stale ...)
    if (Bf16()) { j.W16 = mLinearity.TwinForWrite(); j.ldw16 = mLinearity.TwinStride(); }  // (... and this hands it to the kernel)
    return j;
  }
  /// all-reduce schedule, one layer, first half: the bias gradient (column sums of E) on a side stream, the weight gradient
  /// GEMM on the compute stream.  Only the GEMM stays on the compute stream's critical path.
  /// evEReady: an event of the compute stream behind which this layer's error input is complete (the gradient event of the layer
  /// above, CuNetwork::Backpropagate); NULL = record one now.  Peer-memory schedule: the only event recorded on the compute stream
  /// per layer is mEvG behind the gradient GEMM, so the dX GEMM -> gradient GEMM pair keeps its programmatic-dependent-launch overlap.
  void DataParallelGradient(void *evEReady = NULL) {
    if (mRowsPad == 0) PrepareDataParallel(1);
    if (!mEvE) {
      TNB_CHECK(tnb_event_create(Cx(), &mEvE)); TNB_CHECK(tnb_event_create(Cx(), &mEvB));
      TNB_CHECK(tnb_event_create(Cx(), &mEvAR)); TNB_CHECK(tnb_event_create(Cx(), &mEvDone));
    }
    if (!mEvG) { TNB_CHECK(tnb_event_create(Cx(), &mEvG)); TNB_CHECK(tnb_event_create(Cx(), &mEvPush)); }
    const CuMatrix<BaseFloat> &E = GetErrorInput();
    if (!evEReady) {
      TNB_CHECK(tnb_event_record(Cx(), mEvE, TNB_STREAM_COMPUTE));      // E (and the previous bunch's use of the buffers) is done
      evEReady = mEvE;
    }
    TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_AUX, evEReady));
    TnbBiasJob bj = {E.pCUData(), E.Dim(), NULL, mGrad.pCURowData(mRowsPad), 0.0f, 0.0f, 0, 0};  // gradient only
    TNB_CHECK(tnb_bias_update_batch_on(Cx(), TNB_STREAM_AUX, &bj, 1));
    TNB_CHECK(tnb_event_record(Cx(), mEvB, TNB_STREAM_AUX));
    const int push = mPeerMapped ? PushMode() : 0;
    if (push == 1) {
      // GEMM -> reduce-scatter in one kernel: the tiles of dW go straight to the owning ranks' staging slices over NVLink
      const CuMatrix<BaseFloat> &X = GetInput();
      int rank = 0, world = 1;
      TNB_CHECK(tnb_comm_world(Cx(), &rank, &world));
      const bool bf = Bf16();
      const uint16_t *x16 = bf ? X.Twin() : NULL, *e16 = bf ? E.Twin() : NULL;
      TNB_CHECK(tnb_affine_grad_scatter(Cx(), X.pCUData(), X.Dim(), E.pCUData(), E.Dim(), x16, bf ? X.TwinStride() : 0, e16, bf ? E.TwinStride() : 0,
                                        (float *const *)mGradPeers, world, rank, mLinearity.Dim(), (int)mRowsPad));
    } else if (push == 2) {
      // the GEMM keeps its stores local (mGradLocal); the copy engines carry the row blocks to their owners' staging slices on a
      // side stream while the compute stream goes on with the next layer's GEMMs (no SM waits for NVLink)
      const CuMatrix<BaseFloat> &X = GetInput();
      TnbMatrixDim dG = mLinearity.Dim();
      if (Bf16()) TNB_CHECK(tnb_affine_grad_bf16(Cx(), X.Twin(), X.TwinStride(), X.Dim(), E.Twin(), E.TwinStride(), E.pCUData(), E.Dim(), mGradLocal.pCUData(), dG, NULL));
      else TNB_CHECK(tnb_affine_grad(Cx(), X.pCUData(), X.Dim(), E.pCUData(), E.Dim(), mGradLocal.pCUData(), dG, NULL));
    } else {
      ComputeGradient(false);
    }
    if (mPeerMapped) {
      TNB_CHECK(tnb_event_record(Cx(), mEvG, TNB_STREAM_COMPUTE));
      if (push == 2) {
        int rank = 0, world = 1;
        TNB_CHECK(tnb_comm_world(Cx(), &rank, &world));
        TNB_CHECK(tnb_peer_push_blocks(Cx(), TNB_STREAM_AUX2, mGradLocal.pCUData(), (float *const *)mGradPeers, world, rank, mLinearity.Dim(), (int)mRowsPad,
                                       mEvG, mEvPush));
      }
    }
  }
  /// this layer's way of moving its gradient blocks (TNB_DP_PUSH values); CuNetwork gives the lowest layers, whose exchange is the
  /// exposed tail of the step, the GEMM-epilogue push (no copy-engine latency between the GEMM and the update kernel)
  int PushMode() const { return mPushMode >= 0 ? mPushMode : DpPush(); }
  void SetPushMode(int m) { if (mPeerMapped) Error("SetPushMode after PreparePeer"); mPushMode = m; }
  /// the event behind this layer's gradient GEMM on the compute stream (peer-memory schedule), NULL otherwise
  void *GradientEvent() { return mPeerMapped ? mEvG : NULL; }
  /// TNB_DP_PUSH: how a rank's gradient blocks reach their owners in the peer-memory schedule — 2 (default): copy engines behind a
  /// local gradient GEMM; 1: peer stores of the gradient GEMM's epilogue; 0: the owner's update kernel pulls them with peer loads
  static int DpPush() { static int v = -1; if (v < 0) { const char *e = getenv("TNB_DP_PUSH"); v = e ? atoi(e) : 2; if (v < 0 || v > 2) v = 2; } return v; }
  /// second half: the all-reduce of [dW ; db] on the communication stream behind both, the update on a second side stream behind
  /// the all-reduce.  Nothing waits for it here: whoever touches the parameters next does (WaitDataParallel), which lets the
  /// exchange of this bunch run into the forward pass of the next one.
  void DataParallelReduceUpdate(int n_frames_global) {
    TNB_CHECK(tnb_allreduce_sum_ev(Cx(), mGrad.pCUData(), GradCount(), mEvB, mEvAR));
    TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_AUX2, mEvAR));
    TnbSgdJob job = GradientJob(n_frames_global);
    TNB_CHECK(tnb_sgd_update_batch_on(Cx(), TNB_STREAM_AUX2, &job, 1));
    TNB_CHECK(tnb_event_record(Cx(), mEvDone, TNB_STREAM_AUX2));
    mDpPending = true;
  }
  /// peer-memory schedule (TNB_DP_MODE=peer): map every rank's gradient buffer and weights into this process.  Collective — every
  /// rank calls it for its layers in the same order; the buffers must not be reallocated afterwards (Init() with unchanged
  /// dimensions, CopyFrom() and SetParams() keep them)
  void PreparePeer() {
    if (mPeerMapped) return;
    if (mRowsPad == 0) Error("PreparePeer before PrepareDataParallel");
    TNB_CHECK(tnb_peer_map(Cx(), mGrad.pCUData(), mGradPeers));
    TNB_CHECK(tnb_peer_map(Cx(), mLinearity.pCUData(), mWPeers));
    mGrad.MarkExported();
    mLinearity.MarkExported();
    if (PushMode() == 2) mGradLocal.Init(mRowsPad, mNOutputs);  // padded rows stay zero
    mPeerMapped = true;
  }
  /// second half, peer-memory schedule: ONE kernel on the communication stream sums this rank's block of rows over all ranks'
  /// gradients, updates it and stores the new weights into every rank's matrix (tnb_dp_peer_update)
  void DataParallelPeerUpdate(int n_frames_global) {
    if (!mPeerMapped) Error("DataParallelPeerUpdate before PreparePeer");
    TnbPeerJob j;
    memset(&j, 0, sizeof(j));
    for (int r = 0; r < TNB_MAX_PEERS; r++) { j.G[r] = (float *)mGradPeers[r]; j.W[r] = (float *)mWPeers[r]; }
    (void)mLinearity.pCUData();  // the kernels (this rank's and its peers') rewrite the weights: the bf16 twin goes stale
    j.corrW = mLinearityCorrection.pCUData(); j.bias = mBias.pCUData(); j.corrb = mBiasCorrection.pCUData();
    j.dW = mLinearity.Dim(); j.rows_pad = (int)mRowsPad;
    j.lr = mLearningRate; j.mmt = mMomentum; j.wc = mWeightcost; j.grad_div_frm = mGradDivFrm ? 1 : 0; j.n_frames = n_frames_global;
    j.pushed = PushMode() ? 1 : 0;
    void *waits[2] = {PushMode() == 2 ? mEvPush : mEvG, mEvB};  // the gradient (pushed, or complete where it is) and the bias gradient
    TNB_CHECK(tnb_dp_peer_update_after(Cx(), &j, waits, 2, mEvDone));
    mDpPending = true;
  }
  /// pieces of the second half for a GROUP of layers exchanged in one NCCL launch (CuNetwork's deferred layers)
  void *BiasGradientEvent() { return mEvB; }
  void MarkDataParallelUpdateEnqueued() {  // the group's batched update has just been enqueued on the update stream
    TNB_CHECK(tnb_event_record(Cx(), mEvDone, TNB_STREAM_AUX2));
    mDpPending = true;
  }
  /// order the compute stream behind this layer's outstanding data-parallel update (no host synchronisation)
  void WaitDataParallel() const {
    if (!mDpPending) return;
    TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_COMPUTE, mEvDone));
    mDpPending = false;
  }
  /// events for running this layer's update GEMM on a side stream (created on first use)
  void SideStreamEvents(void **pDx, void **pUpd) {
    if (!mEvE) {
      TNB_CHECK(tnb_event_create(Cx(), &mEvE)); TNB_CHECK(tnb_event_create(Cx(), &mEvB));
      TNB_CHECK(tnb_event_create(Cx(), &mEvAR)); TNB_CHECK(tnb_event_create(Cx(), &mEvDone));
    }
    *pDx = mEvE; *pUpd = mEvDone;
  }
  float *GradBuffer() { return mGrad.pCUData(); }
  size_t GradCount() const { return mGrad.Rows() * mGrad.Stride(); }
  void ApplyGradient(int n_frames_global) {
    TNB_CHECK(tnb_sgd_update(Cx(), mGrad.pCUData(), mLinearity.pCUData(), mLinearityCorrection.pCUData(), mLinearity.Dim(),
                             mGrad.pCURowData(mRowsPad), mBias.pCUData(), mBiasCorrection.pCUData(), mLearningRate, mMomentum, mWeightcost,
                             mGradDivFrm ? 1 : 0, n_frames_global));
  }

  void ReadFromStream(std::istream &rIn) {
    WaitDataParallel();
    BfMatrix transpose;  // stored transposed [out x in] (cuBiasedLinearity.cc:70-78)
    rIn >> transpose;
    BfVector bias;
    rIn >> bias;
    if (transpose.Cols() * transpose.Rows() == 0) Error("Missing linearity matrix in network file");
    if (bias.Dim() == 0) Error("Missing bias vector in network file");
    if (transpose.Rows() != GetNOutputs() || transpose.Cols() != GetNInputs() || bias.Dim() != GetNOutputs()) {
      std::ostringstream os;
      os << "Wrong dimensionalities of matrix/vector in network file\n"
         << "Inputs:" << GetNInputs() << "Outputs:" << GetNOutputs() << "\n"
         << "linearityCols:" << transpose.Rows() << "linearityRows:" << transpose.Cols() << "biasDims:" << bias.Dim() << "\n";
      Error(os.str());
    }
    mLinearity.CopyFrom(BfMatrix(transpose, TRANS));
    mBias.CopyFrom(bias);
  }
  void WriteToStream(std::ostream &rOut) {
    WaitDataParallel();
    BfMatrix tmp;
    mLinearity.CopyTo(tmp);
    rOut << BfMatrix(tmp, TRANS);
    BfVector vec;
    mBias.CopyTo(vec);
    rOut << vec;
    rOut << std::endl;
  }
  /// set the parameters from host memory; Wt is the on-disk layout [nOutputs x nInputs]
  void SetParams(const BfMatrix &Wt, const BfVector &bias) {
    WaitDataParallel();
    if (Wt.Rows() != GetNOutputs() || Wt.Cols() != GetNInputs() || bias.Dim() != GetNOutputs()) Error("SetParams: wrong dimensions");
    mLinearity.CopyFrom(BfMatrix(Wt, TRANS));
    mBias.CopyFrom(bias);
  }
  const CuMatrix<BaseFloat> &Linearity() const { WaitDataParallel(); return mLinearity; }
  const CuVector<BaseFloat> &Bias() const { WaitDataParallel(); return mBias; }

 protected:
  CuMatrix<BaseFloat> mLinearity;  ///< [nInputs x nOutputs]
  CuVector<BaseFloat> mBias;
  CuMatrix<BaseFloat> mLinearityCorrection;
  CuVector<BaseFloat> mBiasCorrection;
  CuMatrix<BaseFloat> mGrad;  ///< data-parallel only: [dW ; db]
  CuMatrix<BaseFloat> mGradLocal;  ///< peer-memory schedule with copy-engine pushes: this rank's full dW before its row blocks travel to their owners' mGrad
  int mDpFrames;
  size_t mRowsPad;
  void *mEvE, *mEvB, *mEvAR, *mEvDone;  ///< data-parallel stream ordering (created on first use)
  void *mEvG, *mEvPush;                 ///< peer-memory schedule: behind the gradient GEMM (compute stream) / behind the copy engines' pushes
  int mPushMode;                        ///< -1 = DpPush()
  mutable bool mDpPending;              ///< an update of this layer is in flight on the side streams
  bool mPeerMapped;                     ///< peer-memory schedule: the tables below are filled
  void *mGradPeers[TNB_MAX_PEERS], *mWPeers[TNB_MAX_PEERS];  ///< every rank's mGrad / mLinearity as mapped into this process
};

// =====================================================================================================
// Feature-transform components (cuCRBEDctFeat.h)
// =====================================================================================================
class CuExpand : public CuComponent {
 public:
  CuExpand(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuComponent(nInputs, nOutputs, pPred) {}
  ComponentType GetType() const { return EXPAND; }
  const char *GetName() const { return "<expand>"; }
  void ReadFromStream(std::istream &rIn) { Vector<int> vec; rIn >> vec; mFrameOffset.CopyFrom(vec); }
  void WriteToStream(std::ostream &rOut) { Vector<int> vec; mFrameOffset.CopyTo(vec); rOut << vec; }
 protected:
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    if (X.Cols() * mFrameOffset.Dim() != Y.Cols()) Error("<expand>: output dim must be input dim x number of offsets");
    CuMath<BaseFloat>::Expand(Y, X, mFrameOffset);
  }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &, CuMatrix<BaseFloat> &) { Error("BackpropagateFnc Nonsense"); }
  CuVector<int> mFrameOffset;
};

class CuCopy : public CuComponent {
 public:
  CuCopy(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuComponent(nInputs, nOutputs, pPred) {}
  ComponentType GetType() const { return COPY; }
  const char *GetName() const { return "<copy>"; }
  void ReadFromStream(std::istream &rIn) { Vector<int> vec; rIn >> vec; vec.Add(-1); mCopyFromIndices.CopyFrom(vec); }  // 1-based on disk
  void WriteToStream(std::ostream &rOut) { Vector<int> vec; mCopyFromIndices.CopyTo(vec); vec.Add(1); rOut << vec; }
 protected:
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { CuMath<BaseFloat>::Rearrange(Y, X, mCopyFromIndices); }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &, CuMatrix<BaseFloat> &) { Error("BackpropagateFnc Nonsense"); }
  CuVector<int> mCopyFromIndices;
};

class CuTranspose : public CuComponent {
 public:
  CuTranspose(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuComponent(nInputs, nOutputs, pPred), mContext(0) {}
  ComponentType GetType() const { return TRANSPOSE; }
  const char *GetName() const { return "<transpose>"; }
  void ReadFromStream(std::istream &rIn) {
    rIn >> std::ws >> mContext;
    if (GetNInputs() != GetNOutputs()) Error("Input dim must be same as output dim");
    if (mContext <= 0 || GetNInputs() % mContext != 0) Error("Number of inputs must be divisible by context length");
    Vector<int> vec(GetNInputs());
    int channels = (int)GetNInputs() / mContext;
    for (int i = 0, ch = 0; ch < channels; ch++)
      for (int idx = ch; idx < (int)GetNInputs(); idx += channels, i++) vec[i] = idx;
    mCopyFromIndices.CopyFrom(vec);
  }
  void WriteToStream(std::ostream &rOut) { rOut << " " << mContext << "\n"; }
 protected:
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { CuMath<BaseFloat>::Rearrange(Y, X, mCopyFromIndices); }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &, CuMatrix<BaseFloat> &) { Error("BackpropagateFnc Nonsense"); }
  int mContext;
  CuVector<int> mCopyFromIndices;
};

class CuBlockLinearity : public CuComponent {
 public:
  CuBlockLinearity(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuComponent(nInputs, nOutputs, pPred) {}
  ComponentType GetType() const { return BLOCK_LINEARITY; }
  const char *GetName() const { return "<blocklinearity>"; }
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { CuMath<BaseFloat>::BlockLinearity(Y, X, mBlockLinearity); }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &, CuMatrix<BaseFloat> &) { Error("BackpropagateFnc Not implemented"); }
  void ReadFromStream(std::istream &rIn) {
    Matrix<BaseFloat> mat;
    rIn >> mat;
    mBlockLinearity.CopyFrom(Matrix<BaseFloat>(mat, TRANS));
    if ((GetNOutputs() % mBlockLinearity.Cols() != 0) || (GetNInputs() % mBlockLinearity.Rows() != 0) ||
        ((GetNOutputs() / mBlockLinearity.Cols()) != (GetNInputs() / mBlockLinearity.Rows())))
      Error("BlockLinearity matrix dimensions must divide IO dims");
  }
  void WriteToStream(std::ostream &rOut) { Matrix<BaseFloat> mat; mBlockLinearity.CopyTo(mat); rOut << Matrix<BaseFloat>(mat, TRANS); }
 private:
  CuMatrix<BaseFloat> mBlockLinearity;
};

class CuBias : public CuComponent {
 public:
  CuBias(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuComponent(nInputs, nOutputs, pPred) {}
  ComponentType GetType() const { return BIAS; }
  const char *GetName() const { return "<bias>"; }
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { Y.CopyFrom(X); Y.AddScaledRow(1.0, mBias, 1.0); }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { Y.CopyFrom(X); }
  void ReadFromStream(std::istream &rIn) { Vector<BaseFloat> vec; rIn >> vec; mBias.CopyFrom(vec); }
  void WriteToStream(std::ostream &rOut) { Vector<BaseFloat> vec; mBias.CopyTo(vec); rOut << vec; }
 private:
  CuVector<BaseFloat> mBias;
};

class CuWindow : public CuComponent {
 public:
  CuWindow(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuComponent(nInputs, nOutputs, pPred) {}
  ComponentType GetType() const { return WINDOW; }
  const char *GetName() const { return "<window>"; }
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { Y.CopyFrom(X); Y.ScaleCols(mWindow); }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &, CuMatrix<BaseFloat> &) { Error("BackpropagateFnc Not implemented"); }
  void ReadFromStream(std::istream &rIn) { Vector<BaseFloat> vec; rIn >> vec; mWindow.CopyFrom(vec); }
  void WriteToStream(std::ostream &rOut) { Vector<BaseFloat> vec; mWindow.CopyTo(vec); rOut << vec; }
 private:
  CuVector<BaseFloat> mWindow;
};

class CuLog : public CuComponent {
 public:
  CuLog(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuComponent(nInputs, nOutputs, pPred) {}
  ComponentType GetType() const { return LOG; }
  const char *GetName() const { return "<log>"; }
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) { Y.CopyFrom(X); Y.ApplyLog(); }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &, CuMatrix<BaseFloat> &) { Error("BackpropagateFnc Not implemented"); }
};

// =====================================================================================================
// CuRbm (cuRbm.{h,cc}) — CD-1 building blocks used by TRbmCu
// =====================================================================================================
/// what TRbmCu's loop needs from an RBM layer (cuRbm.h:15-45): implemented by CuRbm and CuRbmSparse
class CuRbmBase : public CuUpdatableComponent {
 public:
  typedef enum { BERNOULLI, GAUSSIAN } RbmUnitType;
  CuRbmBase(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuUpdatableComponent(nInputs, nOutputs, pPred) {}
  virtual void Propagate(const CuMatrix<BaseFloat> &visProbs, CuMatrix<BaseFloat> &hidProbs) = 0;
  virtual void Reconstruct(const CuMatrix<BaseFloat> &hidState, CuMatrix<BaseFloat> &visProbs) = 0;
  virtual void RbmUpdate(const CuMatrix<BaseFloat> &pos_vis, const CuMatrix<BaseFloat> &pos_hid, const CuMatrix<BaseFloat> &neg_vis,
                         const CuMatrix<BaseFloat> &neg_hid) = 0;
  virtual RbmUnitType VisType() = 0;
  virtual RbmUnitType HidType() = 0;
  using CuComponent::Propagate;  // the no-argument network-internal form stays visible next to the RBM one
};

class CuRbm : public CuRbmBase {
 public:
  CuRbm(size_t nInputs, size_t nOutputs, CuComponent *pPred)
      : CuRbmBase(nInputs, nOutputs, pPred), mVisHid(nInputs, nOutputs), mVisBias(nInputs), mHidBias(nOutputs),
        mVisHidCorrection(nInputs, nOutputs), mVisBiasCorrection(nInputs), mHidBiasCorrection(nOutputs), mVisType(BERNOULLI),
        mHidType(BERNOULLI) {}
  ComponentType GetType() const { return RBM; }
  const char *GetName() const { return "<rbm>"; }
  RbmUnitType VisType() { return mVisType; }
  RbmUnitType HidType() { return mHidType; }

  /// h = (sigmoid)(v*W + hidbias) — one fused GEMM (cuRbm.cc:15-23)
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    TNB_CHECK(tnb_affine_fwd(Cx(), X.pCUData(), X.Dim(), mVisHid.pCUData(), mVisHid.Dim(), mHidBias.pCUData(), Y.pCUData(), Y.Dim(),
                             mHidType == BERNOULLI ? TNB_ACT_SIGMOID : TNB_ACT_NONE));
  }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    if (mHidType == BERNOULLI) {
      mBackpropErrBuf.Init(X.Rows(), X.Cols());
      CuMath<BaseFloat>::DiffSigmoid(mBackpropErrBuf, X, GetOutput());
    } else {
      mBackpropErrBuf.CopyFrom(X);
    }
    Y.Gemm('N', 'T', 1.0, mBackpropErrBuf, mVisHid, 0.0);
  }
  /// backprop-style update (cuRbm.cc:40-100)
  void Update() {
    if (mHidType == BERNOULLI) {
      mBackpropErrBuf.Init(GetErrorInput().Rows(), GetErrorInput().Cols());
      CuMath<BaseFloat>::DiffSigmoid(mBackpropErrBuf, GetErrorInput(), GetOutput());
    } else {
      mBackpropErrBuf.CopyFrom(GetErrorInput());
    }
    BaseFloat N = 1;
    if (mGradDivFrm) N = static_cast<BaseFloat>(GetInput().Rows());
    BaseFloat mmt_gain = static_cast<BaseFloat>(1.0 / (1.0 - mMomentum));
    N *= mmt_gain;
    mVisHidCorrection.Gemm('T', 'N', 1.0, GetInput(), mBackpropErrBuf, mMomentum);
    mHidBiasCorrection.AddColSum(1.0, mBackpropErrBuf, mMomentum);
    mVisHid.AddScaled(-mLearningRate / N, mVisHidCorrection, 1.0);
    mHidBias.AddScaled(-mLearningRate / N, mHidBiasCorrection, 1.0);
    mVisHid.AddScaled(-mLearningRate * mWeightcost, mVisHid, 1.0);
  }
  void Propagate(const CuMatrix<BaseFloat> &visProbs, CuMatrix<BaseFloat> &hidProbs) {
    if (visProbs.Cols() != GetNInputs()) {
      std::ostringstream os;
      os << " Nonmatching input dim, needs:" << GetNInputs() << " got:" << visProbs.Cols() << "\n";
      Error(os.str());
    }
    hidProbs.Init(visProbs.Rows(), GetNOutputs());
    PropagateFnc(visProbs, hidProbs);
  }
  /// v' = (sigmoid)(h*W^T + visbias) (cuRbm.cc:118-128): W is read as the K-major operand, no transposed copy
  void Reconstruct(const CuMatrix<BaseFloat> &hidState, CuMatrix<BaseFloat> &visProbs) {
    visProbs.Init(hidState.Rows(), mNInputs);
    visProbs.AddScaledRow(1.0, mVisBias, 0.0);
    visProbs.Gemm('N', 'T', 1.0, hidState, mVisHid, 1.0);
    if (mVisType == BERNOULLI) CuMath<BaseFloat>::Sigmoid(visProbs, visProbs);
  }
  /// CD-1 update (cuRbm.cc:131-174)
  void RbmUpdate(const CuMatrix<BaseFloat> &pos_vis, const CuMatrix<BaseFloat> &pos_hid, const CuMatrix<BaseFloat> &neg_vis,
                 const CuMatrix<BaseFloat> &neg_hid) {
    if (!(pos_vis.Rows() == pos_hid.Rows() && pos_vis.Rows() == neg_vis.Rows() && pos_vis.Rows() == neg_hid.Rows() &&
          pos_vis.Cols() == neg_vis.Cols() && pos_hid.Cols() == neg_hid.Cols() && pos_vis.Cols() == mNInputs && pos_hid.Cols() == mNOutputs))
      Error("RbmUpdate: non-matching dimensions");
    // weights and both biases through one fused entry point: the second statistics GEMM applies the weight-cost and the weight
    // update in its epilogue, one kernel per bias does its column sums and updates (tnb_rbm_cd1_update has the formulas)
    TNB_CHECK(tnb_rbm_cd1_update(Cx(), pos_vis.pCUData(), neg_vis.pCUData(), pos_vis.Dim(), pos_hid.pCUData(), neg_hid.pCUData(), pos_hid.Dim(),
                                 mVisHid.pCUData(), mVisHid.Dim(), mVisHidCorrection.pCUData(), mVisBias.pCUData(), mVisBiasCorrection.pCUData(),
                                 mHidBias.pCUData(), mHidBiasCorrection.pCUData(), mLearningRate, mMomentum, mWeightcost));
  }
  void ReadFromStream(std::istream &rIn) {
    std::string str;
    rIn >> std::ws >> str;
    if (str == "bern") mVisType = BERNOULLI; else if (str == "gauss") mVisType = GAUSSIAN; else Error(std::string("Invalid unit type: ") + str);
    rIn >> std::ws >> str;
    if (str == "bern") mHidType = BERNOULLI; else if (str == "gauss") mHidType = GAUSSIAN; else Error(std::string("Invalid unit type: ") + str);
    BfMatrix transpose;
    rIn >> transpose;
    if (transpose.Rows() != GetNOutputs() || transpose.Cols() != GetNInputs()) Error("<rbm>: wrong weight matrix dimensions");
    mVisHid.CopyFrom(BfMatrix(transpose, TRANS));
    BfVector bias;
    rIn >> bias; mVisBias.CopyFrom(bias);
    rIn >> bias; mHidBias.CopyFrom(bias);
  }
  void WriteToStream(std::ostream &rOut) {
    rOut << (mVisType == BERNOULLI ? " bern " : " gauss ");
    rOut << (mHidType == BERNOULLI ? " bern\n" : " gauss\n");
    BfMatrix tmp;
    mVisHid.CopyTo(tmp);
    rOut << BfMatrix(tmp, TRANS);
    BfVector vec;
    mVisBias.CopyTo(vec); rOut << vec; rOut << std::endl;
    mHidBias.CopyTo(vec); rOut << vec; rOut << std::endl;
  }
 protected:
  CuMatrix<BaseFloat> mVisHid;
  CuVector<BaseFloat> mVisBias, mHidBias;
  CuMatrix<BaseFloat> mVisHidCorrection;
  CuVector<BaseFloat> mVisBiasCorrection, mHidBiasCorrection;
  CuMatrix<BaseFloat> mBackpropErrBuf;
  RbmUnitType mVisType, mHidType;
};

}  // namespace TNet
#include "cu_nnet_ext.h"  // CuSharedLinearity, CuDiscreteLinearity, CuRbmSparse
namespace TNet {

// =====================================================================================================
// CuRecurrent (cuRecurrent.{h,cc}) — Elman layer, one frame per call, truncated BPTT in Update()
// =====================================================================================================
class CuRecurrent : public CuUpdatableComponent {
 public:
  CuRecurrent(size_t nInputs, size_t nOutputs, CuComponent *pPred)
      : CuUpdatableComponent(nInputs, nOutputs, pPred), mLinearity(nInputs + nOutputs, nOutputs), mBias(nOutputs),
        mLinearityCorrection(nInputs + nOutputs, nOutputs), mBiasCorrection(nOutputs), mBpttOrder(0) {}
  ComponentType GetType() const { return RECURRENT; }
  const char *GetName() const { return "<recurrent>"; }
  void BpttOrder(int ord) {
    mBpttOrder = ord;
    mInputHistory.Init(ord + 1, GetNInputs() + GetNOutputs());
    mHistTmp.Init(ord + 1, GetNInputs() + GetNOutputs());
    mDiffSigm.Init(1, GetNOutputs()); mErrPrev.Init(1, GetNOutputs());
    mDiffs.Init(ord + 1, GetNOutputs());
  }
  void ClearHistory() {
    mInputHistory.SetConst(0.0);
    if (mOutput.MSize() > 0) mOutput.SetConst(0.0);
  }
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    if (X.Rows() != 1 || Y.Rows() != 1) Error("<recurrent> processes one frame per call");
    if (mInputHistory.Rows() == 0) Error("Bptt order was not set");
    const size_t H = mInputHistory.Rows(), K = mInputHistory.Cols();
    // shift the history down by one row (persistent scratch: no per-frame allocation)
    mHistTmp.CopyRows(H - 1, 0, mInputHistory, 0);
    mInputHistory.CopyRows(H - 1, 0, mHistTmp, 1);
    // row 0 = [x_t ; y_{t-1}]
    TNB_CHECK(tnb_memcpy(Cx(), mInputHistory.pCUData(), X.pCUData(), sizeof(BaseFloat) * X.Cols(), 2));
    TNB_CHECK(tnb_memcpy(Cx(), mInputHistory.pCUData() + X.Cols(), Y.pCUData(), sizeof(BaseFloat) * Y.Cols(), 2));
    Y.AddScaledRow(1.0, mBias, 0.0);
    CuMath<BaseFloat>::OffsetGemv('T', 1.0, mLinearity, mInputHistory.pCUData(), K, 1.0, Y.pCUData(), Y.Cols(), 0);
    CuMath<BaseFloat>::Sigmoid(Y, Y);
  }
  void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    if (X.Rows() != 1 || Y.Rows() != 1) Error("<recurrent> processes one frame per call");
    CuMath<BaseFloat>::DiffSigmoid(mDiffSigm, X, GetOutput());
    // accumulates into Y with beta = 1 as the reference does (cuRecurrent.cc:81)
    CuMath<BaseFloat>::OffsetGemv('N', 1.0, mLinearity, mDiffSigm.pCUData(), mDiffSigm.Cols(), 1.0, Y.pCUData(), Y.Cols(), 0);
  }
  /// cuRecurrent.cc:92-153.  Same arithmetic per element as the reference's sequence of calls; the BPTT chain runs one fused
  /// launch per step (tnb_rnn_bptt_step) and the rank-1 updates of all steps are applied in one pass over W (tnb_rnn_apply).
  void Update() {
    const size_t H = GetNOutputs(), nin = GetInput().Cols();
    // step 0: d_0 = diffsigmoid(error, output) ; bcorr = -lr*d_0 + mmt*bcorr
    TnbMatrixDim drow = {1, (int)H, (int)mDiffs.Stride()};
    TNB_CHECK(tnb_diff_sigmoid(Cx(), mDiffs.pCUData(), GetErrorInput().pCUData(), GetOutput().pCUData(), drow));
    TNB_CHECK(tnb_add_col_sum(Cx(), -mLearningRate, mDiffs.pCUData(), mMomentum, mBiasCorrection.pCUData(), drow));
    for (int i = 1; i <= mBpttOrder; i++)   // d_i from d_{i-1} and the activations of history frame i-1 (columns nin.. of its row)
      TNB_CHECK(tnb_rnn_bptt_step(Cx(), mLinearity.pCUData(), mLinearity.Dim(), (int)nin, mDiffs.pCURowData(i - 1),
                                  mInputHistory.pCURowData(i - 1) + nin, mDiffs.pCURowData(i), mBiasCorrection.pCUData(), mLearningRate));
    TNB_CHECK(tnb_rnn_apply(Cx(), mLinearity.pCUData(), mLinearity.Dim(), mInputHistory.pCUData(), (int)mInputHistory.Stride(), mDiffs.pCUData(),
                            (int)mDiffs.Stride(), mBpttOrder + 1, mLearningRate, mWeightcost));
    mBias.AddScaled(1.0, mBiasCorrection, 1.0);
  }
  void ReadFromStream(std::istream &rIn) {
    BfMatrix transpose;
    rIn >> transpose;
    if (transpose.Rows() != GetNOutputs() || transpose.Cols() != GetNInputs() + GetNOutputs()) Error("<recurrent>: wrong weight matrix dimensions");
    mLinearity.CopyFrom(BfMatrix(transpose, TRANS));
    BfVector bias;
    rIn >> bias;
    mBias.CopyFrom(bias);
  }
  void WriteToStream(std::ostream &rOut) {
    BfMatrix tmp;
    mLinearity.CopyTo(tmp);
    rOut << BfMatrix(tmp, TRANS);
    BfVector vec;
    mBias.CopyTo(vec);
    rOut << vec;
    rOut << std::endl;
  }
 protected:
  CuMatrix<BaseFloat> mLinearity;
  CuVector<BaseFloat> mBias;
  CuMatrix<BaseFloat> mLinearityCorrection;
  CuVector<BaseFloat> mBiasCorrection;
  CuMatrix<BaseFloat> mInputHistory, mHistTmp, mDiffSigm, mErrPrev;
  CuMatrix<BaseFloat> mDiffs;  ///< [(bptt + 1) x nOutputs]: the back-propagated error of every BPTT step of the current frame
  int mBpttOrder;
};

// =====================================================================================================
// Objective functions (cuObjectiveFunction.{h,cc})
// =====================================================================================================
class CuObjectiveFunction {
 public:
  typedef enum { OBJ_FUN_I = 0x0300, MEAN_SQUARE_ERROR, CROSS_ENTROPY } ObjFunType;
  static CuObjectiveFunction *Factory(ObjFunType type);
  CuObjectiveFunction() : mpStats(NULL) {
    void *p = NULL;
    TNB_CHECK(tnb_malloc(Cx(), &p, sizeof(TnbObjStats)));  // zero-filled
    mpStats = (TnbObjStats *)p;
  }
  virtual ~CuObjectiveFunction() { if (mpStats) tnb_free(Cx(), mpStats); }
  virtual ObjFunType GetTypeId() = 0;
  virtual const char *GetTypeLabel() = 0;
  virtual void Evaluate(const CuMatrix<BaseFloat> &rNetOutput, const CuMatrix<BaseFloat> &rDesired, CuMatrix<BaseFloat> &rNetError) = 0;
  double GetError() { return Read().error; }
  size_t GetFrames() { return (size_t)Read().frames; }
  size_t GetCorrect() { return (size_t)Read().correct; }
  virtual std::string Report() = 0;
  TnbObjStats *DeviceStats() { return mpStats; }
  /// data parallel: fold the other ranks' totals into this one (cf. MergeStats, src/TNetLib/ObjFun.cc:214-228)
  void AddStats(double error, long long frames, long long correct) {
    TnbObjStats st = Read();
    st.error += error; st.frames += frames; st.correct += correct;
    TNB_CHECK(tnb_memcpy(Cx(), mpStats, &st, sizeof(st), 0));
    CuDevice::Instantiate().Sync();
  }
 protected:
  TnbObjStats Read() {
    TnbObjStats st;
    TNB_CHECK(tnb_memcpy(Cx(), &st, mpStats, sizeof(st), 1));
    return st;
  }
  TnbObjStats *mpStats;  ///< device-resident accumulators
};

/// the objective kernels index the targets with the network output's rows and pitch (one TnbMatrixDim for both): refuse a target
/// matrix of another shape or pitch instead of reading it out of bounds
inline void CheckTargetLayout(const CuMatrix<BaseFloat> &rOut, const CuMatrix<BaseFloat> &rDesired) {
  if (rDesired.Rows() != rOut.Rows() || rDesired.Stride() != rOut.Stride()) {
    std::ostringstream os;
    os << "Non-matching dimensions of network output with training targets!!! Netoutput rows:" << rOut.Rows() << " pitch:" << rOut.Stride()
       << " Targets rows:" << rDesired.Rows() << " pitch:" << rDesired.Stride();
    Error(os.str());
  }
}

class CuMeanSquareError : public CuObjectiveFunction {
 public:
  ObjFunType GetTypeId() { return MEAN_SQUARE_ERROR; }
  const char *GetTypeLabel() { return "<mean_square_error>"; }
  void Evaluate(const CuMatrix<BaseFloat> &rNetOutput, const CuMatrix<BaseFloat> &rDesired, CuMatrix<BaseFloat> &rNetError) {
    if (rDesired.Cols() != rNetOutput.Cols() || rDesired.Rows() != rNetOutput.Rows()) Error("Non-matching dimensions of network output with training targets!!!");
    CheckTargetLayout(rNetOutput, rDesired);
    rNetError.Init(rNetOutput.Rows(), rNetOutput.Cols());
    TNB_CHECK(tnb_mse_eval(Cx(), rNetOutput.pCUData(), rDesired.pCUData(), rNetError.pCUData(), rNetOutput.Dim(), mpStats));
  }
  std::string Report() {
    TnbObjStats st = Read();
    std::ostringstream ss;
    ss << "Mse:" << st.error << " frames:" << (size_t)st.frames << " err/frm:" << st.error / st.frames << "\n";
    return ss.str();
  }
};

class CuCrossEntropy : public CuObjectiveFunction {
 public:
  ObjFunType GetTypeId() { return CROSS_ENTROPY; }
  const char *GetTypeLabel() { return "<cross_entropy>"; }
  void Evaluate(const CuMatrix<BaseFloat> &rNetOutput, const CuMatrix<BaseFloat> &rDesired, CuMatrix<BaseFloat> &rNetError) {
    if (rDesired.Cols() != rNetOutput.Cols()) {
      std::ostringstream os;
      os << "Non-matching dimensions of network output with training targets!!!" << " Netoutput:" << rNetOutput.Cols()
         << " Targets:" << rDesired.Cols();
      Error(os.str());
    }
    CheckTargetLayout(rNetOutput, rDesired);
    rNetError.Init(rNetOutput.Rows(), rNetOutput.Cols());
    TNB_CHECK(tnb_xent_eval(Cx(), rNetOutput.pCUData(), rDesired.pCUData(), rNetError.pCUData(), rNetOutput.Dim(), mpStats));
  }
  /// softmax + Evaluate in one kernel, from the pre-softmax activations (used by CuNetwork::PropagateEvaluate)
  void EvaluateFromActivations(const CuMatrix<BaseFloat> &rAct, const CuMatrix<BaseFloat> &rDesired, CuMatrix<BaseFloat> &rSoftmaxOut,
                               CuMatrix<BaseFloat> &rNetError) {
    if (rDesired.Cols() != rAct.Cols()) Error("Non-matching dimensions of network output with training targets!!!");
    CheckTargetLayout(rAct, rDesired);
    rSoftmaxOut.Init(rAct.Rows(), rAct.Cols());
    rNetError.Init(rAct.Rows(), rAct.Cols());
    TNB_CHECK(tnb_softmax_xent(Cx(), rAct.pCUData(), rDesired.pCUData(), rSoftmaxOut.pCUData(), rNetError.pCUData(), rAct.Dim(), mpStats));
  }
  /// Targets as ONE CLASS ID PER FRAME: rLabels is a [rows x 1] matrix whose 4-byte elements hold int32 ids (the form CuCache keeps
  /// them in when it is fed ids: 128 bytes per frame of pitch instead of 4 * nOutputs).  Same results, bit for bit, as Evaluate on the
  /// one-hot matrix of those ids (tnb_xent_eval_labels).
  void EvaluateLabels(const CuMatrix<BaseFloat> &rNetOutput, const CuMatrix<BaseFloat> &rLabels, CuMatrix<BaseFloat> &rNetError) {
    if (rLabels.Cols() != 1 || rLabels.Rows() != rNetOutput.Rows()) Error("EvaluateLabels: the labels must be a [frames x 1] matrix of class ids");
    EvaluateIds(rNetOutput, (const int *)rLabels.pCUData(), (int)rLabels.Stride(), rNetError);
  }
  void EvaluateFromActivationsLabels(const CuMatrix<BaseFloat> &rAct, const CuMatrix<BaseFloat> &rLabels, CuMatrix<BaseFloat> &rSoftmaxOut,
                                     CuMatrix<BaseFloat> &rNetError) {
    if (rLabels.Cols() != 1 || rLabels.Rows() != rAct.Rows()) Error("EvaluateFromActivationsLabels: the labels must be a [frames x 1] matrix of class ids");
    EvaluateFromActivationsIds(rAct, (const int *)rLabels.pCUData(), (int)rLabels.Stride(), rSoftmaxOut, rNetError);
  }
  /// the same on a device array of ids, `stride` ints apart (1 = packed), one per row of the network output
  void EvaluateIds(const CuMatrix<BaseFloat> &rNetOutput, const int *pIds, int stride, CuMatrix<BaseFloat> &rNetError) {
    rNetError.Init(rNetOutput.Rows(), rNetOutput.Cols());
    TNB_CHECK(tnb_xent_eval_labels(Cx(), rNetOutput.pCUData(), pIds, stride, rNetError.pCUData(), rNetOutput.Dim(), mpStats));
  }
  void EvaluateFromActivationsIds(const CuMatrix<BaseFloat> &rAct, const int *pIds, int stride, CuMatrix<BaseFloat> &rSoftmaxOut,
                                  CuMatrix<BaseFloat> &rNetError) {
    rSoftmaxOut.Init(rAct.Rows(), rAct.Cols());
    rNetError.Init(rAct.Rows(), rAct.Cols());
    TNB_CHECK(tnb_softmax_xent_labels(Cx(), rAct.pCUData(), pIds, stride, rSoftmaxOut.pCUData(), rNetError.pCUData(), rAct.Dim(), mpStats));
  }
  /// [frames x 1] label matrix from a vector of ids (a strided device-to-device copy; the bits of the ints are kept)
  static void LabelsFromIds(const CuVector<int> &rIds, size_t first, size_t rows, CuMatrix<BaseFloat> &rLabels) {
    if (first + rows > rIds.Dim()) Error("LabelsFromIds: range");
    rLabels.Init(rows, 1);
    if (rows == 0) return;
    TNB_CHECK(tnb_memcpy2d(Cx(), rLabels.pCUData(), rLabels.Stride() * sizeof(BaseFloat), rIds.pCUData() + first, sizeof(int), sizeof(int), rows, 2));
  }
  /// byte-compatible with cuObjectiveFunction.h:132-144 (tools/train/training_scheduler.sh greps it)
  std::string Report() {
    TnbObjStats st = Read();
    std::ostringstream ss;
    ss << "Xent:" << st.error << " frames:" << (size_t)st.frames << " err/frm:" << st.error / st.frames << " correct["
       << 100.0 * st.correct / st.frames << "%]" << "\n";
    return ss.str();
  }
};

inline CuObjectiveFunction *CuObjectiveFunction::Factory(ObjFunType type) {
  switch (type) {
    case MEAN_SQUARE_ERROR: return new CuMeanSquareError;
    case CROSS_ENTROPY: return new CuCrossEntropy;
    default: Error("Unknown ObjFun type");
  }
  return NULL;
}

// =====================================================================================================
// CuNetwork (cuNetwork.{h,cc})
// =====================================================================================================
class CuNetwork {
  typedef std::vector<CuComponent *> LayeredType;
 public:
  CuNetwork() : mpPropagErrorStopper(NULL), mGlobLearnRate(0.0), mpLearnRateFactors(NULL), mpTempBasisDir(NULL), mFuse(true), mBatch(BatchDefault()), mBatchPolicy(BatchPolicyDefault()), mBatchFlushTiles(BatchFlushDefault()), mWorld(1), mEvGroup(NULL), mDpGroup(false), mBwdStreams(BwdStreamsDefault()), mDpDeferBegin(0), mDpDeferEnd(0), mDpShard(false), mDpPeer(false) {
    const char *e = getenv("TNB_FUSE");
    if (e && atoi(e) == 0) mFuse = false;
  }
  explicit CuNetwork(std::istream &rIn)
      : mpPropagErrorStopper(NULL), mGlobLearnRate(0.0), mpLearnRateFactors(NULL), mpTempBasisDir(NULL), mFuse(true), mBatch(BatchDefault()), mBatchPolicy(BatchPolicyDefault()), mBatchFlushTiles(BatchFlushDefault()), mWorld(1), mEvGroup(NULL), mDpGroup(false), mBwdStreams(BwdStreamsDefault()), mDpDeferBegin(0), mDpDeferEnd(0), mDpShard(false), mDpPeer(false) {
    ReadNetwork(rIn);
  }
  ~CuNetwork() {
    if (mEvGroup) tnb_event_destroy(Cx(), mEvGroup);
    for (LayeredType::iterator it = mNetComponents.begin(); it != mNetComponents.end(); ++it) delete *it;
    mNetComponents.clear();
  }
  void AddLayer(CuComponent *layer) {
    if (mNetComponents.size() > 0) {
      if (GetNOutputs() != layer->GetNInputs()) Error("Nonmatching dims");
      layer->SetInput(mNetComponents.back()->GetOutput());
      mNetComponents.back()->SetErrorInput(layer->GetErrorOutput());
    }
    mNetComponents.push_back(layer);
  }
  int Layers() { return (int)mNetComponents.size(); }
  CuComponent &Layer(int i) { return *mNetComponents[i]; }

  static int BwdStreamsDefault() { const char *e = getenv("TNB_BWD_STREAMS"); return e ? atoi(e) : 1; }
  /// -1 = by measurement (profiles/r02_batch_kernel.md): on in bf16 mode (0.465 against 0.474 ms per bunch on config C), off in 3xTF32, whose
  /// mainloop already saturates the shared-memory port, so that an epilogue running under it slows both (0.988 against 0.944 ms)
  static int BatchDefault() { const char *e = getenv("TNB_GEMM_BATCH"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }
  void SetBatching(bool on) { mBatch = on ? 1 : 0; }
  static int BatchPolicyDefault() { const char *e = getenv("TNB_BATCH_POLICY"); return e ? atoi(e) : 1; }
  static int BatchFlushDefault() { const char *e = getenv("TNB_BATCH_FLUSH_TILES"); return e ? atoi(e) : 148; }
  void SetBatchPolicy(int p) { mBatchPolicy = p; }
  void SetFusion(bool on) { mFuse = on; }
  /// order the compute stream behind every outstanding data-parallel update (stream order only; callers that time or end a
  /// run of bunches use it so that the last bunch's exchange is inside what they measure)
  void WaitDataParallel() {
    for (size_t i = 0; i < mNetComponents.size(); i++)
      if (mNetComponents[i]->GetType() == CuComponent::BIASED_LINEARITY) static_cast<CuBiasedLinearity *>(mNetComponents[i])->WaitDataParallel();
  }
  /// data parallel over `world` ranks: Update() becomes gradient -> exchange -> apply (N uses rows*world)
  void SetDataParallel(int world) {
    // only <biasedlinearity> has a gradient exchange; any other trainable layer would silently train on its rank's rows alone
    if (world > 1)
      for (size_t i = 0; i < mNetComponents.size(); i++)
        if (mNetComponents[i]->IsUpdatable() && mNetComponents[i]->GetType() != CuComponent::BIASED_LINEARITY)
          Error(std::string("data-parallel training exchanges the gradients of <biasedlinearity> layers only; the network contains ") +
                mNetComponents[i]->GetName());
    mWorld = world;
    const char *e = getenv("TNB_DP_MODE");
    mDpShard = e && !strcmp(e, "shard");
    // default with more than one rank: the all-reduce schedule's order with the NCCL call + update kernel of every layer replaced by
    // ONE peer-memory kernel (csrc/peer.cu).  Measured on config C, ms per bunch: 2 B200 1.081 against 1.262 with NCCL all-reduce,
    // 8 B200 1.152 against 1.493.  TNB_DP_MODE=allreduce|shard select the NCCL schedules.
    mDpPeer = world > 1 && (!e || !strcmp(e, "peer"));
    if (e && strcmp(e, "peer") && strcmp(e, "shard") && strcmp(e, "allreduce")) Error(std::string("TNB_DP_MODE must be peer, allreduce or shard, not ") + e);
    // default: with L >= 6 updatable layers the top one and the bottom two exchange at once, the middle ones late (measured on
    // 2 B200, config C, ms per bunch: no deferral 1.311, 1:5 1.244, 2:5 1.261, 3:5 1.288, 2:6 1.305);
    // TNB_DP_DEFER=begin:end overrides (0:0 = plain top-to-bottom order)
    int nupd = 0;
    for (size_t i = 0; i < mNetComponents.size(); i++) nupd += mNetComponents[i]->GetType() == CuComponent::BIASED_LINEARITY;
    mDpDeferBegin = mDpDeferEnd = 0;
    if (nupd >= 6) { mDpDeferBegin = 1; mDpDeferEnd = nupd - 2; }
    // one NCCL launch for the deferred layers pays a call's fixed latency once (8 ranks: ~50 us of a 16.8 MB layer's 101 us) but
    // holds their updates until the whole group is through.  Measured slower on both 2 ranks (1.340 vs 1.237 ms per bunch) and
    // 8 ranks (1.560 vs 1.496 ms): off unless TNB_DP_GROUP=1
    mDpGroup = false;
    const char *gr = getenv("TNB_DP_GROUP");
    if (gr) mDpGroup = atoi(gr) != 0;
    const char *d = getenv("TNB_DP_DEFER");
    if (d) { int a = 0, b = 0; if (sscanf(d, "%d:%d", &a, &b) == 2) { mDpDeferBegin = a; mDpDeferEnd = b; } }
    for (size_t i = 0; i < mNetComponents.size(); i++)
      if (mNetComponents[i]->GetType() == CuComponent::BIASED_LINEARITY) static_cast<CuBiasedLinearity *>(mNetComponents[i])->PrepareDataParallel(world);
    if (mDpPeer) {
      // TNB_DP_PUSH_TAIL=n (default 2): the n lowest layers push from the GEMM epilogue when the others use the copy engines
      int tail = 2;
      const char *t = getenv("TNB_DP_PUSH_TAIL");
      if (t) tail = atoi(t);
      for (size_t i = 0; i < mNetComponents.size() && tail > 0; i++)
        if (mNetComponents[i]->GetType() == CuComponent::BIASED_LINEARITY) {
          CuBiasedLinearity *lin = static_cast<CuBiasedLinearity *>(mNetComponents[i]);
          if (lin->PushMode() == 2) lin->SetPushMode(1);
          tail--;
        }
      for (size_t i = 0; i < mNetComponents.size(); i++)
        if (mNetComponents[i]->GetType() == CuComponent::BIASED_LINEARITY) static_cast<CuBiasedLinearity *>(mNetComponents[i])->PreparePeer();
    }
  }

  /// forward the data to the output (cuNetwork.h:137-165)
  void Propagate(const CuMatrix<BaseFloat> &in, CuMatrix<BaseFloat> &out) {
    if (mNetComponents.size() == 0) { out.CopyFrom(in); return; }
    ForwardTo(in, mNetComponents.size());
    out.CopyFrom(mNetComponents.back()->GetOutput());
  }
  /// Propagate + CuCrossEntropy::Evaluate with the final <softmax> fused into the objective kernel; the network
  /// output stays in Layer(last).GetOutput() (no copy).  Falls back to Propagate+Evaluate for any other tail.
  void PropagateEvaluate(const CuMatrix<BaseFloat> &in, const CuMatrix<BaseFloat> &desired, CuObjectiveFunction &obj,
                         CuMatrix<BaseFloat> &globerr) {
    size_t n = mNetComponents.size();
    CuCrossEntropy *xent = dynamic_cast<CuCrossEntropy *>(&obj);
    if (mFuse && xent && n >= 2 && mNetComponents[n - 1]->GetType() == CuComponent::SOFTMAX) {
      ForwardTo(in, n - 1);
      xent->EvaluateFromActivations(mNetComponents[n - 2]->GetOutput(), desired, mNetComponents[n - 1]->MutableOutput(), globerr);
    } else {
      if (n == 0) Error("PropagateEvaluate on an empty network");
      ForwardTo(in, n);
      obj.Evaluate(mNetComponents.back()->GetOutput(), desired, globerr);
    }
  }

  /// PropagateEvaluate with class ids ([frames x 1] matrix of int32 bits, see CuCrossEntropy::EvaluateLabels) instead of dense targets
  void PropagateEvaluateLabels(const CuMatrix<BaseFloat> &in, const CuMatrix<BaseFloat> &labels, CuCrossEntropy &xent, CuMatrix<BaseFloat> &globerr) {
    if (labels.Cols() != 1 || labels.Rows() != in.Rows()) Error("PropagateEvaluateLabels: the labels must be a [frames x 1] matrix of class ids");
    PropagateEvaluateIds(in, (const int *)labels.pCUData(), (int)labels.Stride(), xent, globerr);
  }
  /// ... or a device array of ids `stride` ints apart (1 = packed), one per row of `in`
  void PropagateEvaluateIds(const CuMatrix<BaseFloat> &in, const int *pIds, int stride, CuCrossEntropy &xent, CuMatrix<BaseFloat> &globerr) {
    size_t n = mNetComponents.size();
    if (n == 0) Error("PropagateEvaluate on an empty network");
    if (mFuse && n >= 2 && mNetComponents[n - 1]->GetType() == CuComponent::SOFTMAX) {
      ForwardTo(in, n - 1);
      xent.EvaluateFromActivationsIds(mNetComponents[n - 2]->GetOutput(), pIds, stride, mNetComponents[n - 1]->MutableOutput(), globerr);
    } else {
      ForwardTo(in, n);
      xent.EvaluateIds(mNetComponents.back()->GetOutput(), pIds, stride, globerr);
    }
  }

  /// backpropagate the error while updating weights (cuNetwork.h:170-194)
  void Backpropagate(const CuMatrix<BaseFloat> &globerr) {
    if (mNetComponents.size() == 0) return;
    const int n = (int)mNetComponents.size();
    mNetComponents.back()->SetErrorInput(globerr);
    std::vector<CuBiasedLinearity *> pending;  // data parallel: layers whose gradient is in flight
    std::vector<CuBiasedLinearity *> deferred; // data parallel, all-reduce schedule: layers whose exchange is issued after the lowest layer's
    std::vector<void *> side_done;             // fused schedule on two streams: one event per layer, behind its update GEMM
    std::vector<TnbBiasJob> bias_jobs;         // fused schedule: bias halves of the updates, applied together after the last layer
    int sig_done_by_batch = -1;                // batched schedule: index of the <sigmoid> whose backward the last dX job included
    void *dp_prev_grad_event = NULL;           // peer-memory schedule: event behind the gradient GEMM of the updatable layer above (its dX precedes it: this layer's error is complete)
    // Batched schedule (single GPU, fused): the dX GEMM of a layer shares ONE persistent launch with tiles of the weight-gradient
    // GEMMs of the layers ABOVE it (tnb_gemm_batch).  A layer's update job enters the pool only after its own dX has been
    // launched (the update rewrites W in place, dX reads it); what is left in the pool runs in a last launch.
    // mBatchPolicy 2: the dX chain keeps its own split-K launches (it is the critical path and wants all SMs); only the weight
    // updates are pooled, and the pool runs as launches of two tiles per CTA pair whenever that many have accumulated.
    std::vector<TnbGemmJob> pool;
    const bool batch = mFuse && (mBatch == 1 || (mBatch < 0 && CuBiasedLinearity::Bf16())) && mWorld == 1 && mBwdStreams != 2;
    const bool batch_dx = batch && mBatchPolicy == 1;
    for (int i = n - 1; i >= 0; i--) {
      CuComponent *c = mNetComponents[i];
      if (c != mpPropagErrorStopper) {
        bool done = false;
        if (batch_dx && c->GetType() == CuComponent::BIASED_LINEARITY) {
          // dX of this layer as a job, with the diff-sigmoid of a <sigmoid> right below fused as in the unbatched schedule
          CuBiasedLinearity *lin = static_cast<CuBiasedLinearity *>(c);
          const bool sig_below = i > 0 && mNetComponents[i - 1]->GetType() == CuComponent::SIGMOID && mNetComponents[i - 1] != mpPropagErrorStopper;
          CuMatrix<BaseFloat> &eprev = sig_below ? mNetComponents[i - 1]->MutableErrorOutput() : c->MutableErrorOutput();
          eprev.Init(c->GetErrorInput().Rows(), c->GetNInputs());
          TnbGemmJob job;
          // alone in a launch, a GEMM with fewer tiles than half the pairs is better off on the split-K tiles of the single-GEMM path
          const bool pool_has_work = !pool.empty();
          if (lin->MakeBackpropJob(c->GetErrorInput(), sig_below ? &mNetComponents[i - 1]->GetOutput() : NULL, eprev, &job) &&
              (pool_has_work || tnb_gemm_job_tiles(&job) >= 48)) {
            TNB_CHECK(tnb_gemm_batch(Cx(), &job, 1, pool.empty() ? NULL : &pool[0], (int)pool.size()));
            while (!pool.empty() && pool.front().tile_count == 0) pool.erase(pool.begin());
            done = true;
            if (sig_below) sig_done_by_batch = i - 1;
          }
        }
        if (!done && batch && c->GetType() == CuComponent::SIGMOID && i == sig_done_by_batch) done = true;
        if (!done && mFuse) {
          // <softmax> backward is the identity: hand globerr to the layer below instead of copying it
          if (c->GetType() == CuComponent::SOFTMAX && i > 0) {
            mNetComponents[i - 1]->SetErrorInput(c->GetErrorInput());
            done = true;
          }
          // affine dX fused with the diff-sigmoid of the <sigmoid> right below it
          if (!done && c->GetType() == CuComponent::BIASED_LINEARITY && i > 0 && mNetComponents[i - 1]->GetType() == CuComponent::SIGMOID &&
              mNetComponents[i - 1] != mpPropagErrorStopper) {
            CuComponent *sig = mNetComponents[i - 1];
            CuMatrix<BaseFloat> &eprev = sig->MutableErrorOutput();
            eprev.Init(c->GetErrorInput().Rows(), sig->GetNInputs());
            static_cast<CuBiasedLinearity *>(c)->BackpropagateDiffSigmoid(c->GetErrorInput(), sig->GetOutput(), eprev);
            done = true;
          }
          // ... in which case the <sigmoid> itself has nothing left to do
          if (!done && c->GetType() == CuComponent::SIGMOID && i + 1 < n && mNetComponents[i + 1]->GetType() == CuComponent::BIASED_LINEARITY &&
              mNetComponents[i + 1] != mpPropagErrorStopper)
            done = true;
        }
        if (!done) {
          c->Backpropagate();
          if (c->GetType() != CuComponent::BIASED_LINEARITY) dp_prev_grad_event = NULL;  // this kernel, not the gradient GEMM above, completes the error below
        }
      }
      if (c->IsUpdatable()) {
        CuUpdatableComponent &rComp = dynamic_cast<CuUpdatableComponent &>(*c);
        if (rComp.LearnRate() > 0.0f) {
          if (mWorld > 1 && c->GetType() == CuComponent::BIASED_LINEARITY) {
            CuBiasedLinearity *lin = static_cast<CuBiasedLinearity *>(c);
            if (mDpShard) lin->ComputeGradient();
            if (mDpShard) lin->DataParallelUpdate((int)lin->GetInput().Rows() * mWorld);  // reduce-scatter / update / all-gather
            else {
              // Gradient now; exchange + update now or deferred.  The collectives are the step's critical path and the lowest
              // layer's gradient, which the next forward pass needs FIRST, is the last one to exist: the middle layers'
              // exchanges are therefore issued after the lowest layers', in forward order, and run into the next bunch's
              // forward pass (every layer's forward waits only for its own update, CuBiasedLinearity::WaitDataParallel).
              lin->DataParallelGradient(dp_prev_grad_event);
              dp_prev_grad_event = lin->GradientEvent();
              const int k = (int)pending.size();
              if (k >= mDpDeferBegin && k < mDpDeferEnd) deferred.push_back(lin);
              else if (mDpPeer) lin->DataParallelPeerUpdate((int)lin->GetInput().Rows() * mWorld);
              else lin->DataParallelReduceUpdate((int)lin->GetInput().Rows() * mWorld);
            }
            pending.push_back(lin);
          } else if (mFuse && c->GetType() == CuComponent::BIASED_LINEARITY) {
            bias_jobs.push_back(TnbBiasJob());
            CuBiasedLinearity *lin = static_cast<CuBiasedLinearity *>(c);
            TnbGemmJob ujob;
            if (batch && (int)pool.size() < 4 && lin->MakeUpdateJob(&ujob, &bias_jobs.back())) {
              pool.push_back(ujob);  // runs next to the dX GEMMs of the layers below (or in the closing launch)
              if (!batch_dx) {       // policy 2: a full launch (two tiles on every pair) as soon as the pool holds one
                int tiles = 0;
                for (size_t k = 0; k < pool.size(); k++) tiles += pool[k].tile_count;
                if (tiles >= mBatchFlushTiles || (int)pool.size() >= 4) {
                  TNB_CHECK(tnb_gemm_batch(Cx(), NULL, 0, &pool[0], (int)pool.size()));
                  pool.clear();
                }
              }
            } else if (mBwdStreams == 2) {
              // the weight-gradient GEMM (+ fused update) of this layer only needs E and X, which exist: it runs on a side stream next
              // to the dX GEMMs of the layers below (their CTAs fill the SMs the other kernel leaves free or has finished with)
              void *ev_dx = NULL, *ev_upd = NULL;
              lin->SideStreamEvents(&ev_dx, &ev_upd);
              TNB_CHECK(tnb_event_record(Cx(), ev_dx, TNB_STREAM_COMPUTE));   // dX of this layer has read W; E of this layer exists
              TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_AUX, ev_dx));
              TNB_CHECK(tnb_ctx_use_stream(Cx(), TNB_STREAM_AUX));
              lin->Update(&bias_jobs.back());
              TNB_CHECK(tnb_ctx_use_stream(Cx(), TNB_STREAM_COMPUTE));
              TNB_CHECK(tnb_event_record(Cx(), ev_upd, TNB_STREAM_AUX));
              side_done.push_back(ev_upd);
            } else {
              lin->Update(&bias_jobs.back());
            }
          } else {
            rComp.Update();
          }
        }
      }
      if (mpPropagErrorStopper == c) break;
    }
    if (!pool.empty()) TNB_CHECK(tnb_gemm_batch(Cx(), NULL, 0, &pool[0], (int)pool.size()));  // closing launch: every update still in the pool
    for (size_t k = 0; k < side_done.size(); k++) TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_COMPUTE, side_done[k]));
    for (size_t k = 0; k < bias_jobs.size(); k += TNB_MAX_BIAS_JOBS)
      TNB_CHECK(tnb_bias_update_batch(Cx(), &bias_jobs[k], (int)std::min<size_t>(TNB_MAX_BIAS_JOBS, bias_jobs.size() - k)));
    if (mDpPeer) {
      for (size_t k = deferred.size(); k-- > 0;) deferred[k]->DataParallelPeerUpdate((int)deferred[k]->GetInput().Rows() * mWorld);
    } else if (!mDpGroup) {
      for (size_t k = deferred.size(); k-- > 0;)  // bottom-most deferred layer first: the order the next forward pass needs them
        deferred[k]->DataParallelReduceUpdate((int)deferred[k]->GetInput().Rows() * mWorld);
    } else if (!deferred.empty()) {
      // the deferred layers go out together: one NCCL launch for all of them (a call's fixed latency is paid once), one update
      // launch behind it, bottom-most layer first
      if (!mEvGroup) TNB_CHECK(tnb_event_create(Cx(), &mEvGroup));
      std::vector<float *> bufs;
      std::vector<size_t> counts;
      std::vector<void *> evs;
      std::vector<TnbSgdJob> jobs;
      for (size_t k = deferred.size(); k-- > 0;) {
        bufs.push_back(deferred[k]->GradBuffer());
        counts.push_back(deferred[k]->GradCount());
        evs.push_back(deferred[k]->BiasGradientEvent());
      }
      TNB_CHECK(tnb_allreduce_sum_multi(Cx(), bufs.data(), counts.data(), (int)bufs.size(), evs.data(), (int)evs.size(), mEvGroup));
      TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_AUX2, mEvGroup));
      for (size_t k = deferred.size(); k-- > 0;) jobs.push_back(deferred[k]->GradientJob((int)deferred[k]->GetInput().Rows() * mWorld));
      for (size_t k = 0; k < jobs.size(); k += TNB_MAX_BIAS_JOBS)
        TNB_CHECK(tnb_sgd_update_batch_on(Cx(), TNB_STREAM_AUX2, &jobs[k], (int)std::min<size_t>(TNB_MAX_BIAS_JOBS, jobs.size() - k)));
      for (size_t k = 0; k < deferred.size(); k++) deferred[k]->MarkDataParallelUpdateEnqueued();
    }
    if (!pending.empty() && mDpShard) TNB_CHECK(tnb_comm_wait(Cx()));  // the next forward pass reads the gathered weights
    // restore the wiring the fused softmax step changed
    if (mFuse && n >= 2 && mNetComponents[n - 1]->GetType() == CuComponent::SOFTMAX)
      mNetComponents[n - 2]->SetErrorInput(mNetComponents[n - 1]->GetErrorOutput());
  }

  void ReadNetwork(const char *pSrc) {
    std::ifstream in(pSrc);
    if (!in.good()) Error(std::string("Error, cannot read model: ") + pSrc);
    ReadNetwork(in);
    in.close();
  }
  void WriteNetwork(const char *pDst) {
    std::ofstream out(pDst);
    if (!out.good()) Error(std::string("Error, cannot write model: ") + pDst);
    WriteNetwork(out);
    out.close();
  }
  void ReadNetwork(std::istream &rIn) {
    CuComponent *pComp;
    while (NULL != (pComp = ComponentFactory(rIn))) mNetComponents.push_back(pComp);
  }
  void WriteNetwork(std::ostream &rOut) {
    for (LayeredType::iterator it = mNetComponents.begin(); it != mNetComponents.end(); ++it) ComponentDumper(rOut, **it);
  }

  size_t GetNInputs() const { return mNetComponents.empty() ? 0 : mNetComponents.front()->GetNInputs(); }
  size_t GetNOutputs() const { return mNetComponents.empty() ? 0 : mNetComponents.back()->GetNOutputs(); }

  /// learn rate with per-layer factors "a:b:c" or "a,b,c"; also selects the backprop stopper (cuNetwork.cc:80-135)
  void SetLearnRate(BaseFloat learnRate, const char *pLearnRateFactors = NULL) {
    std::list<BaseFloat> lr_factors;
    if (NULL != pLearnRateFactors) {
      std::string str(pLearnRateFactors);
      for (size_t p = 0; p < str.size(); p++) if (str[p] == ':' || str[p] == ',') str[p] = ' ';
      std::istringstream is(str);
      BaseFloat f;
      while (is >> f) lr_factors.push_back(f);
    }
    BaseFloat scale = 1.0f;
    mGlobLearnRate = learnRate;
    mpLearnRateFactors = pLearnRateFactors;
    mpPropagErrorStopper = NULL;
    bool stopper_given = false;
    for (LayeredType::iterator it = mNetComponents.begin(); it != mNetComponents.end(); ++it) {
      if ((*it)->IsUpdatable()) {
        if (NULL != pLearnRateFactors) {
          if (!(lr_factors.size() > 0)) Error("Too few learninig rate scale factors");
          scale = lr_factors.front();
          lr_factors.pop_front();
        }
        dynamic_cast<CuUpdatableComponent *>(*it)->LearnRate(learnRate * scale);
        if (!stopper_given && (learnRate * scale > 0.0)) { mpPropagErrorStopper = *it; stopper_given = true; }
      }
    }
    if (lr_factors.size() > 0) Error("Too much learninig rate scale factors");
  }
  BaseFloat GetLearnRate() { return mGlobLearnRate; }
  void PrintLearnRate() {
    std::cout << "Learning rate: global " << mGlobLearnRate;
    std::cout << " components' ";
    for (size_t i = 0; i < mNetComponents.size(); i++)
      if (mNetComponents[i]->IsUpdatable()) std::cout << " " << dynamic_cast<CuUpdatableComponent *>(mNetComponents[i])->LearnRate();
    std::cout << "\n" << std::flush;
  }
  void SetMomentum(BaseFloat momentum) { ForUpdatable(&CuUpdatableComponent::Momentum, momentum); }
  void SetWeightcost(BaseFloat weightcost) { ForUpdatable(&CuUpdatableComponent::Weightcost, weightcost); }
  void SetL1(BaseFloat) {}  ///< only <sparselinearity> uses it (cuNetwork.cc:186-196); the flag is accepted
  void SetGradDivFrm(bool div) {
    for (LayeredType::iterator it = mNetComponents.begin(); it != mNetComponents.end(); ++it)
      if ((*it)->IsUpdatable()) dynamic_cast<CuUpdatableComponent *>(*it)->GradDivFrm(div);
  }
  void SetTempBasisDir(const char *pDir) { mpTempBasisDir = pDir; }

 private:
  CuNetwork(CuNetwork &);
  CuNetwork &operator=(CuNetwork &);
  void ForUpdatable(void (CuUpdatableComponent::*setter)(BaseFloat), BaseFloat v) {
    for (LayeredType::iterator it = mNetComponents.begin(); it != mNetComponents.end(); ++it)
      if ((*it)->IsUpdatable()) (dynamic_cast<CuUpdatableComponent *>(*it)->*setter)(v);
  }
  /// run layers [0, upto)
  void ForwardTo(const CuMatrix<BaseFloat> &in, size_t upto) {
    if (in.Cols() != GetNInputs()) {
      std::ostringstream os;
      os << "Nonmatching dims" << " data dim is: " << in.Cols() << " network needs: " << GetNInputs();
      Error(os.str());
    }
    mNetComponents.front()->SetInput(in);
    for (size_t i = 0; i < upto; i++) {
      CuComponent *c = mNetComponents[i];
      if (mFuse && c->GetType() == CuComponent::BIASED_LINEARITY && i + 1 < upto && mNetComponents[i + 1]->GetType() == CuComponent::SIGMOID) {
        CuComponent *sig = mNetComponents[i + 1];
        if (c->GetNInputs() != c->GetInput().Cols())
          KALDI_ERR << "Non-matching INPUT dim!!! Network dim: " << c->GetNInputs() << " Data dim: " << c->GetInput().Cols();
        CuMatrix<BaseFloat> &y = sig->MutableOutput();
        y.Init(c->GetInput().Rows(), sig->GetNOutputs());
        static_cast<CuBiasedLinearity *>(c)->PropagateSigmoid(c->GetInput(), y);
        i++;  // the <sigmoid> is done
      } else {
        c->Propagate();
      }
    }
  }
  CuComponent *ComponentFactory(std::istream &rIn);
  void ComponentDumper(std::ostream &rOut, CuComponent &rComp);

  LayeredType mNetComponents;
  CuComponent *mpPropagErrorStopper;
  BaseFloat mGlobLearnRate;
  const char *mpLearnRateFactors;
  const char *mpTempBasisDir;
  bool mFuse;
  int mBatch;                      ///< fused schedule: independent backward GEMMs share persistent launches (tnb_gemm_batch; TNB_GEMM_BATCH=0 disables)
  int mBatchPolicy;                ///< 1: dX jobs share launches with pooled updates; 2: dX on its own split-K launches, updates pooled into full launches
  int mBatchFlushTiles;            ///< policy 2: pool size (tiles) that triggers a launch
  int mWorld;
  void *mEvGroup;                  ///< data parallel: behind the grouped all-reduce of the deferred layers
  bool mDpGroup;                   ///< deferred layers in one NCCL launch (TNB_DP_GROUP=1; measured slower, off by default)
  int mBwdStreams;                 ///< fused single-GPU schedule: 2 = weight-gradient GEMMs on a side stream (TNB_BWD_STREAMS)
  int mDpDeferBegin, mDpDeferEnd;  ///< peer-memory and all-reduce schedules: updatable layers [begin, end), counted from the top, exchange late
  bool mDpShard;  ///< TNB_DP_MODE=shard: reduce-scatter / block update / all-gather per layer with NCCL (tnb_dp_update)
  bool mDpPeer;   ///< default with several ranks: one peer-memory kernel per layer (tnb_dp_peer_update); both false = NCCL all-reduce + update
};

// =====================================================================================================
// CuCache (cuCache.{h,cc})
// =====================================================================================================
class CuCache {
  typedef enum { EMPTY, INTAKE, FULL, EXHAUST } State;
 public:
  CuCache() : mState(EMPTY), mIntakePos(0), mExhaustPos(0), mCachesize(0), mBunchsize(0), mDiscarded(0), mRandomized(false), mTrace(0) {}
  void Init(size_t cachesize, size_t bunchsize) {
    if (bunchsize == 0 || (cachesize % bunchsize) != 0) Error("Non divisible cachesize by bunchsize");
    mCachesize = cachesize; mBunchsize = bunchsize;
    mState = EMPTY; mIntakePos = 0; mExhaustPos = 0; mRandomized = false;
  }
  void AddData(const CuMatrix<BaseFloat> &rFeatures, const CuMatrix<BaseFloat> &rDesired);
  void Randomize();
  void GetBunch(CuMatrix<BaseFloat> &rFeatures, CuMatrix<BaseFloat> &rDesired);
  /// hand the rows that did not fit into this fill to `dst` (the cache object of the NEXT fill): two cache objects filled
  /// alternately then hold exactly what the reference's single cache holds fill after fill (cuCache.cc:60-76,102-108)
  void MoveLeftoverTo(CuCache &dst) {
    dst.mFeaturesLeftover.Swap(mFeaturesLeftover);
    dst.mDesiredLeftover.Swap(mDesiredLeftover);
    mFeaturesLeftover.Destroy();
    mDesiredLeftover.Destroy();
  }
  bool Full() { return (mState == FULL); }
  bool Empty() { return (mState == EMPTY || mIntakePos < mBunchsize); }
  int Discarded() { return mDiscarded; }
  void Trace(int trace) { mTrace = trace; }
  /// the permutation of the last Randomize() (host copy; for parity tests)
  const Vector<int> &LastPermutation() const { return mLastPerm; }
  size_t IntakePos() const { return mIntakePos; }

 private:
  static long int GenerateRandom(int max) { return lrand48() % max; }
  State mState;
  size_t mIntakePos, mExhaustPos, mCachesize, mBunchsize;
  int mDiscarded;
  CuMatrix<BaseFloat> mFeatures, mFeaturesRandom, mFeaturesLeftover;
  CuMatrix<BaseFloat> mDesired, mDesiredRandom, mDesiredLeftover;
  CuVector<int> mCuRandMask;
  Vector<int> mLastPerm;
  bool mRandomized;
  int mTrace;
};

}  // namespace TNet
#endif
