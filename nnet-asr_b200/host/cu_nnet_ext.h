// cu_nnet_ext.h — the remaining updatable layers of the reference's component factory: the users of
// CuMath::OffsetGemm and the sparse RBM.  Included by cu_nnet.h (after CuRbmBase); not a stand-alone header.
//
// Reference files mirrored (paths under src/CuTNetLib): cuSharedLinearity.{h,cc}, cuDiscreteLinearity.{h,cc}, cuRbmSparse.{h,cc}.
// All three run CuNetwork's component-by-component schedule (nothing is fused around them): their GEMMs address column blocks of
// the activation matrices, which tnb_gemm takes through the tcgen05 kernel when the block starts on a 16-byte boundary with a
// 16-byte-multiple pitch and through the fp32 FMA kernel otherwise (the <blocklinearity> case).
#ifndef TNETB200_CU_NNET_EXT_H_
#define TNETB200_CU_NNET_EXT_H_

namespace TNet {

// =====================================================================================================
// CuSharedLinearity (cuSharedLinearity.{h,cc}) — ONE [nin/K x nout/K] weight block applied to K consecutive column blocks of
// the input ("instances", e.g. the frames of a spliced window), one shared bias
// =====================================================================================================
class CuSharedLinearity : public CuUpdatableComponent {
 public:
  CuSharedLinearity(size_t nInputs, size_t nOutputs, CuComponent *pPred) : CuUpdatableComponent(nInputs, nOutputs, pPred), mNInstances(0) {}
  ComponentType GetType() const { return SHARED_LINEARITY; }
  const char *GetName() const { return "<sharedlinearity>"; }

  /// Y = [b b ... b]; Y[:, i-th block] += X[:, i-th block] * W (cuSharedLinearity.cc:9-25)
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    CuMath<BaseFloat>::VecExpand(mBias, mBiasExpand, mZeroOffsets);
    Y.AddScaledRow(1.0, mBiasExpand, 0.0);
    for (int i = 0; i < mNInstances; i++)
      CuMath<BaseFloat>::OffsetGemm('N', 'N', 1.0, X, mLinearity, 1.0, Y, i * (int)mLinearity.Rows(), 0, i * (int)mLinearity.Cols());
  }
  /// Eprev[:, i-th block] = E[:, i-th block] * W^T (cuSharedLinearity.cc:28-36)
  void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    for (int i = 0; i < mNInstances; i++)
      CuMath<BaseFloat>::OffsetGemm('N', 'T', 1.0, X, mLinearity, 0.0, Y, i * (int)mLinearity.Cols(), 0, i * (int)mLinearity.Rows());
  }
  /// the "#if 1 new implementation" (cuSharedLinearity.cc:62-93): the K instances' gradients are summed into one correction
  /// matrix (momentum applied with the first), N additionally scaled by K
  void Update() {
    BaseFloat N = 1;
    if (mGradDivFrm) N = static_cast<BaseFloat>(GetInput().Rows());
    BaseFloat mmt_gain = static_cast<BaseFloat>(1.0 / (1.0 - mMomentum));
    N *= mmt_gain;
    N *= static_cast<BaseFloat>(mNInstances);
    for (int i = 0; i < mNInstances; i++)
      CuMath<BaseFloat>::OffsetGemm('T', 'N', 1.0, GetInput(), GetErrorInput(), ((i == 0) ? mMomentum : 1.0f), mLinearityCorrection,
                                    i * (int)mLinearity.Rows(), i * (int)mLinearity.Cols(), 0);
    mBiasCorrectionExpand.AddColSum(1.0, GetErrorInput(), 0.0);
    CuMath<BaseFloat>::VecAddColSum(1.0, mBiasCorrectionExpand, mMomentum, mBiasCorrection);
    mLinearity.AddScaled(-mLearningRate / N, mLinearityCorrection, 1.0);
    mBias.AddScaled(-mLearningRate / N, mBiasCorrection, 1.0);
    mLinearity.AddScaled(-mLearningRate * mWeightcost, mLinearity, 1.0);
  }

  void ReadFromStream(std::istream &rIn) {
    rIn >> std::ws >> mNInstances;
    if (rIn.fail() || mNInstances < 1) {
      std::ostringstream os;
      os << "Bad number of instances:" << mNInstances;
      Error(os.str());
    }
    if (GetNInputs() % mNInstances != 0 || GetNOutputs() % mNInstances != 0) {
      std::ostringstream os;
      os << "Number of Inputs/Outputs must be divisible by number of instances" << " Inputs:" << GetNInputs() << " Outputs" << GetNOutputs()
         << " Intances:" << mNInstances;
      Error(os.str());
    }
    BfMatrix transpose;  // stored transposed, as in <biasedlinearity>
    rIn >> transpose;
    BfVector bias;
    rIn >> bias;
    if (transpose.Cols() * transpose.Rows() == 0) Error("Missing linearity matrix in network file");
    if (bias.Dim() == 0) Error("Missing bias vector in network file");
    if (transpose.Rows() != GetNOutputs() / mNInstances || transpose.Cols() != GetNInputs() / mNInstances ||
        bias.Dim() != GetNOutputs() / mNInstances) {
      std::ostringstream os;
      os << "Wrong dimensionalities of matrix/vector in network file\n" << "Inputs:" << GetNInputs() << "Outputs:" << GetNOutputs() << "\n"
         << "linearityCols:" << transpose.Rows() << "linearityRows:" << transpose.Cols() << "biasDims:" << bias.Dim() << "\n";
      Error(os.str());
    }
    mLinearity.CopyFrom(BfMatrix(transpose, TRANS));
    mBias.CopyFrom(bias);
    mLinearityCorrection.Init(mLinearity.Rows(), mLinearity.Cols());
    mLinearityCorrection.SetZero();
    mBiasCorrection.Init(mBias.Dim());
    mBiasCorrection.SetZero();
    mBiasExpand.Init(mBias.Dim() * mNInstances);
    mBiasCorrectionExpand.Init(mBias.Dim() * mNInstances);
    mZeroOffsets.Init(mNInstances);
    mZeroOffsets.SetZero();
  }
  void WriteToStream(std::ostream &rOut) {
    rOut << mNInstances << std::endl;
    BfMatrix tmp;
    mLinearity.CopyTo(tmp);
    rOut << BfMatrix(tmp, TRANS);
    BfVector vec;
    mBias.CopyTo(vec);
    rOut << vec;
    rOut << std::endl;
  }
  int NInstances() const { return mNInstances; }

 protected:
  CuMatrix<BaseFloat> mLinearity;  ///< [nin/K x nout/K]
  CuVector<BaseFloat> mBias;       ///< [nout/K]
  CuMatrix<BaseFloat> mLinearityCorrection;
  CuVector<BaseFloat> mBiasCorrection;
  int mNInstances;
  CuVector<BaseFloat> mBiasExpand, mBiasCorrectionExpand;  ///< [nout]
  CuVector<int> mZeroOffsets;                               ///< K zeros: the "offsets" of the bias tiling (CuMath::VecExpand)
};

// =====================================================================================================
// CuDiscreteLinearity (cuDiscreteLinearity.{h,cc}) — block-diagonal affine layer: block i maps its own slice of the input
// columns to its own slice of the output columns with its own weights; one bias over all outputs
// =====================================================================================================
class CuDiscreteLinearity : public CuUpdatableComponent {
 public:
  CuDiscreteLinearity(size_t nInputs, size_t nOutputs, CuComponent *pPred)
      : CuUpdatableComponent(nInputs, nOutputs, pPred), mBias(nOutputs), mBiasCorrection(nOutputs), mNBlocks(0) {}
  ~CuDiscreteLinearity() {
    for (size_t i = 0; i < mLinearity.size(); i++) { delete mLinearity[i]; delete mLinearityCorrection[i]; }
  }
  ComponentType GetType() const { return DISCRETE_LINEARITY; }
  const char *GetName() const { return "<discretelinearity>"; }

  /// cuDiscreteLinearity.cc:7-25
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    Y.AddScaledRow(1.0, mBias, 0.0);
    int offset_in = 0, offset_out = 0;
    for (int i = 0; i < mNBlocks; i++) {
      CuMath<BaseFloat>::OffsetGemm('N', 'N', 1.0, X, *mLinearity[i], 1.0, Y, offset_in, 0, offset_out);
      offset_in += (int)mLinearity[i]->Rows();
      offset_out += (int)mLinearity[i]->Cols();
    }
  }
  /// cuDiscreteLinearity.cc:28-41
  void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    int offset_in = 0, offset_out = 0;
    for (int i = 0; i < mNBlocks; i++) {
      CuMath<BaseFloat>::OffsetGemm('N', 'T', 1.0, X, *mLinearity[i], 0.0, Y, offset_in, 0, offset_out);
      offset_in += (int)mLinearity[i]->Cols();
      offset_out += (int)mLinearity[i]->Rows();
    }
  }
  /// cuDiscreteLinearity.cc:44-76
  void Update() {
    BaseFloat N = 1;
    if (mGradDivFrm) N = static_cast<BaseFloat>(GetInput().Rows());
    BaseFloat mmt_gain = static_cast<BaseFloat>(1.0 / (1.0 - mMomentum));
    N *= mmt_gain;
    int offset_in = 0, offset_out = 0;
    for (int i = 0; i < mNBlocks; i++) {
      CuMath<BaseFloat>::OffsetGemm('T', 'N', 1.0, GetInput(), GetErrorInput(), mMomentum, *mLinearityCorrection[i], offset_in, offset_out, 0);
      offset_in += (int)mLinearity[i]->Rows();
      offset_out += (int)mLinearity[i]->Cols();
    }
    for (int i = 0; i < mNBlocks; i++) {
      mLinearity[i]->AddScaled(-mLearningRate / N, *mLinearityCorrection[i], 1.0);
      mLinearity[i]->AddScaled(-mLearningRate * mWeightcost, *mLinearity[i], 1.0);
    }
    mBiasCorrection.AddColSum(1.0, GetErrorInput(), mMomentum);
    mBias.AddScaled(-mLearningRate / N, mBiasCorrection, 1.0);
  }

  void ReadFromStream(std::istream &rIn) {
    rIn >> std::ws >> mNBlocks;
    if (rIn.fail() || mNBlocks < 1) KALDI_ERR << "Bad number of blocks:" << mNBlocks;
    size_t in_dim = 0, out_dim = 0;
    for (int i = 0; i < mNBlocks; i++) {
      BfMatrix transpose;  // every block stored transposed
      rIn >> transpose;
      if (transpose.Cols() * transpose.Rows() == 0) Error("Missing linearity matrix in network file");
      mLinearity.push_back(new CuMatrix<BaseFloat>());
      mLinearityCorrection.push_back(new CuMatrix<BaseFloat>());
      mLinearity.back()->CopyFrom(BfMatrix(transpose, TRANS));
      mLinearityCorrection.back()->Init(transpose.Cols(), transpose.Rows());
      mLinearityCorrection.back()->SetZero();
      in_dim += transpose.Cols();
      out_dim += transpose.Rows();
    }
    BfVector bias;
    rIn >> bias;
    if (bias.Dim() == 0) Error("Missing bias vector in network file");
    if (out_dim != GetNOutputs() || in_dim != GetNInputs() || bias.Dim() != GetNOutputs()) {
      std::ostringstream os;
      os << "Wrong dimensionalities of matrix/vector in network file\n" << "Inputs:" << GetNInputs() << "Outputs:" << GetNOutputs() << "\n"
         << "linearityCols:" << in_dim << "linearityRows:" << out_dim << "biasDims:" << bias.Dim() << "\n";
      Error(os.str());
    }
    mBias.CopyFrom(bias);
    mBiasCorrection.Init(mBias.Dim());
    mBiasCorrection.SetZero();
  }
  void WriteToStream(std::ostream &rOut) {
    rOut << mNBlocks << "\n";
    for (int i = 0; i < mNBlocks; i++) {
      BfMatrix tmp;
      mLinearity[i]->CopyTo(tmp);
      rOut << BfMatrix(tmp, TRANS);
    }
    BfVector vec;
    mBias.CopyTo(vec);
    rOut << vec;
    rOut << std::endl;
  }
  int NBlocks() const { return mNBlocks; }

 protected:
  std::vector<CuMatrix<BaseFloat> *> mLinearity;  ///< block i: [in_i x out_i] (CuMatrix is not copyable: held by pointer)
  CuVector<BaseFloat> mBias;
  std::vector<CuMatrix<BaseFloat> *> mLinearityCorrection;
  CuVector<BaseFloat> mBiasCorrection;
  int mNBlocks;
};

// =====================================================================================================
// CuRbmSparse (cuRbmSparse.{h,cc}) — CuRbm with a sparsity penalty on the hidden units in the CD-1 update: q (running mean
// activity per hidden unit, decay lambda) is pulled towards the prior p by  dW -= cost * mean(v) (q-p)^T,  dhb -= cost * (q-p)
// =====================================================================================================
class CuRbmSparse : public CuRbmBase {
 public:
  CuRbmSparse(size_t nInputs, size_t nOutputs, CuComponent *pPred)
      : CuRbmBase(nInputs, nOutputs, pPred), mVisHid(nInputs, nOutputs), mVisBias(nInputs), mHidBias(nOutputs),
        mVisHidCorrection(nInputs, nOutputs), mVisBiasCorrection(nInputs), mHidBiasCorrection(nOutputs), mVisType(BERNOULLI),
        mHidType(BERNOULLI), mSparsityPrior(0.0001), mLambda(0.95), mSparsityCost(1e-7), mSparsityQ(nOutputs), mSparsityQCurrent(nOutputs),
        mVisMean(nInputs) {
    mVisHidCorrection.SetZero();
    mVisBiasCorrection.SetZero();
    mHidBiasCorrection.SetZero();
    mSparsityQ.SetConst(mSparsityPrior);
    mSparsityQCurrent.SetZero();
    mVisMean.SetZero();
  }
  ComponentType GetType() const { return RBM_SPARSE; }
  const char *GetName() const { return "<rbmsparse>"; }
  RbmUnitType VisType() { return mVisType; }
  RbmUnitType HidType() { return mHidType; }

  /// h = (sigmoid)(v*W + hidbias) — one fused GEMM (cuRbmSparse.cc:13-23)
  void PropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    TNB_CHECK(tnb_affine_fwd(Cx(), X.pCUData(), X.Dim(), mVisHid.pCUData(), mVisHid.Dim(), mHidBias.pCUData(), Y.pCUData(), Y.Dim(),
                             mHidType == BERNOULLI ? TNB_ACT_SIGMOID : TNB_ACT_NONE));
  }
  /// cuRbmSparse.cc:26-38
  void BackpropagateFnc(const CuMatrix<BaseFloat> &X, CuMatrix<BaseFloat> &Y) {
    DiffHidden(X);
    Y.Gemm('N', 'T', 1.0, mBackpropErrBuf, mVisHid, 0.0);
  }
  /// backprop-style update (cuRbmSparse.cc:41-91, the "#if 1" branch)
  void Update() {
    DiffHidden(GetErrorInput());
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
  /// v' = (sigmoid)(h*W^T + visbias) (cuRbmSparse.cc:112-122)
  void Reconstruct(const CuMatrix<BaseFloat> &hidState, CuMatrix<BaseFloat> &visProbs) {
    visProbs.Init(hidState.Rows(), mNInputs);
    visProbs.AddScaledRow(1.0, mVisBias, 0.0);
    visProbs.Gemm('N', 'T', 1.0, hidState, mVisHid, 1.0);
    if (mVisType == BERNOULLI) CuMath<BaseFloat>::Sigmoid(visProbs, visProbs);
  }
  /// CD-1 update with the sparsity terms (cuRbmSparse.cc:125-168)
  void RbmUpdate(const CuMatrix<BaseFloat> &pos_vis, const CuMatrix<BaseFloat> &pos_hid, const CuMatrix<BaseFloat> &neg_vis,
                 const CuMatrix<BaseFloat> &neg_hid) {
    if (!(pos_vis.Rows() == pos_hid.Rows() && pos_vis.Rows() == neg_vis.Rows() && pos_vis.Rows() == neg_hid.Rows() &&
          pos_vis.Cols() == neg_vis.Cols() && pos_hid.Cols() == neg_hid.Cols() && pos_vis.Cols() == mNInputs && pos_hid.Cols() == mNOutputs))
      Error("RbmUpdate: non-matching dimensions");
    if (mHidType == BERNOULLI) {
      mSparsityQCurrent.AddColSum(1.0 / pos_hid.Rows(), pos_hid, 0.0);
      mSparsityQ.AddScaled(1.0 - mLambda, mSparsityQCurrent, mLambda);
      mSparsityQCurrent.SetConst(-mSparsityPrior);
      mSparsityQCurrent.AddScaled(1.0, mSparsityQ, 1.0);
      mVisMean.AddColSum(1.0 / pos_vis.Rows(), pos_vis, 0.0);
    }
    BaseFloat N = static_cast<BaseFloat>(pos_vis.Rows());
    mVisHidCorrection.Gemm('T', 'N', -mLearningRate / N, neg_vis, neg_hid, mMomentum);
    mVisHidCorrection.Gemm('T', 'N', +mLearningRate / N, pos_vis, pos_hid, 1.0);
    mVisHidCorrection.AddScaled(-mLearningRate * mWeightcost, mVisHid, 1.0);
    if (mHidType == BERNOULLI) mVisHidCorrection.BlasGer(-mSparsityCost, mVisMean, mSparsityQCurrent);
    mVisHid.AddScaled(1.0, mVisHidCorrection, 1.0);
    mVisBiasCorrection.AddColSum(-mLearningRate / N, neg_vis, mMomentum);
    mVisBiasCorrection.AddColSum(+mLearningRate / N, pos_vis, 1.0);
    mVisBias.AddScaled(1.0, mVisBiasCorrection, 1.0);
    mHidBiasCorrection.AddColSum(-mLearningRate / N, neg_hid, mMomentum);
    mHidBiasCorrection.AddColSum(+mLearningRate / N, pos_hid, 1.0);
    if (mHidType == BERNOULLI) mHidBiasCorrection.AddScaled(-mSparsityCost, mSparsityQCurrent, 1.0);
    mHidBias.AddScaled(1.0, mHidBiasCorrection, 1.0);
  }
  void ReadFromStream(std::istream &rIn) {
    std::string str;
    rIn >> std::ws >> str;
    if (str == "bern") mVisType = BERNOULLI; else if (str == "gauss") mVisType = GAUSSIAN; else Error(std::string("Invalid unit type: ") + str);
    rIn >> std::ws >> str;
    if (str == "bern") mHidType = BERNOULLI; else if (str == "gauss") mHidType = GAUSSIAN; else Error(std::string("Invalid unit type: ") + str);
    BfMatrix transpose;
    rIn >> transpose;
    if (transpose.Rows() != GetNOutputs() || transpose.Cols() != GetNInputs()) Error("<rbmsparse>: wrong weight matrix dimensions");
    mVisHid.CopyFrom(BfMatrix(transpose, TRANS));
    BfVector bias;
    rIn >> bias; mVisBias.CopyFrom(bias);
    rIn >> bias; mHidBias.CopyFrom(bias);
    rIn >> std::ws >> mSparsityCost;
    std::cout << "RBM::mSparsityCost=" << mSparsityCost;
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
    rOut << mSparsityCost << std::endl;
  }

 protected:
  void DiffHidden(const CuMatrix<BaseFloat> &E) {
    if (mHidType == BERNOULLI) {
      mBackpropErrBuf.Init(E.Rows(), E.Cols());
      CuMath<BaseFloat>::DiffSigmoid(mBackpropErrBuf, E, GetOutput());
    } else {
      mBackpropErrBuf.CopyFrom(E);
    }
  }
  CuMatrix<BaseFloat> mVisHid;
  CuVector<BaseFloat> mVisBias, mHidBias;
  CuMatrix<BaseFloat> mVisHidCorrection;
  CuVector<BaseFloat> mVisBiasCorrection, mHidBiasCorrection;
  CuMatrix<BaseFloat> mBackpropErrBuf;
  RbmUnitType mVisType, mHidType;
  BaseFloat mSparsityPrior;  ///< target activity of a hidden unit
  BaseFloat mLambda;         ///< decay of the running activity estimate q
  BaseFloat mSparsityCost;
  CuVector<BaseFloat> mSparsityQ, mSparsityQCurrent, mVisMean;
};

}  // namespace TNet
#endif
