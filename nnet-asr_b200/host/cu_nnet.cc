// cu_nnet.cc — out-of-line parts of CuNetwork (component factory / dumper) and CuCache.
// Reference: src/CuTNetLib/cuNetwork.cc:213-387, src/CuTNetLib/cuCache.cc:21-200.
#include "cu_nnet.h"

namespace TNet {

namespace {
struct TagEntry {
  const char *tag;
  CuComponent::ComponentType type;
};
// tags the reader understands; the unsupported ones are named so that the error says why
const TagEntry kTags[] = {
    {"<biasedlinearity>", CuComponent::BIASED_LINEARITY}, {"<rbm>", CuComponent::RBM},           {"<recurrent>", CuComponent::RECURRENT},
    {"<softmax>", CuComponent::SOFTMAX},                  {"<sigmoid>", CuComponent::SIGMOID},   {"<expand>", CuComponent::EXPAND},
    {"<copy>", CuComponent::COPY},                        {"<transpose>", CuComponent::TRANSPOSE}, {"<blocklinearity>", CuComponent::BLOCK_LINEARITY},
    {"<bias>", CuComponent::BIAS},                        {"<window>", CuComponent::WINDOW},     {"<log>", CuComponent::LOG},
    {"<sharedlinearity>", CuComponent::SHARED_LINEARITY}, {"<discretelinearity>", CuComponent::DISCRETE_LINEARITY}, {"<rbmsparse>", CuComponent::RBM_SPARSE},
};
const char *kOutOfScope[] = {"<sparselinearity>", "<blockarray>", "<clusterlinearity>"};
}  // namespace

CuComponent *CuNetwork::ComponentFactory(std::istream &rIn) {
  rIn >> std::ws;
  if (rIn.eof()) return NULL;
  std::string componentTag;
  rIn >> componentTag;
  if (componentTag == "") return NULL;
  std::transform(componentTag.begin(), componentTag.end(), componentTag.begin(), ::tolower);
  if (componentTag[0] != '<' || componentTag[componentTag.size() - 1] != '>') Error(std::string("Invalid component tag:") + componentTag);
  if (componentTag == "<endblock>") return NULL;

  size_t nInputs = 0, nOutputs = 0;
  rIn >> std::ws >> nOutputs >> std::ws >> nInputs;
  if (rIn.fail() || nInputs == 0 || nOutputs == 0) Error(std::string("Missing dimensions after component tag:") + componentTag);

  CuComponent *pPred = mNetComponents.empty() ? NULL : mNetComponents.back();
  CuComponent *pRet = NULL;
  int found = -1;
  for (size_t i = 0; i < sizeof(kTags) / sizeof(kTags[0]); i++)
    if (componentTag == kTags[i].tag) found = (int)i;
  if (found < 0) {
    for (size_t i = 0; i < sizeof(kOutOfScope) / sizeof(kOutOfScope[0]); i++)
      if (componentTag == kOutOfScope[i]) Error(std::string("Component not built into the B200 hot path (see DESIGN.md, out of scope): ") + componentTag);
    Error(std::string("Unknown Component tag:") + componentTag);
  }
  switch (kTags[found].type) {
    case CuComponent::BIASED_LINEARITY: pRet = new CuBiasedLinearity(nInputs, nOutputs, pPred); break;
    case CuComponent::RBM: pRet = new CuRbm(nInputs, nOutputs, pPred); break;
    case CuComponent::RECURRENT: pRet = new CuRecurrent(nInputs, nOutputs, pPred); break;
    case CuComponent::SOFTMAX: pRet = new CuSoftmax(nInputs, nOutputs, pPred); break;
    case CuComponent::SIGMOID: pRet = new CuSigmoid(nInputs, nOutputs, pPred); break;
    case CuComponent::EXPAND: pRet = new CuExpand(nInputs, nOutputs, pPred); break;
    case CuComponent::COPY: pRet = new CuCopy(nInputs, nOutputs, pPred); break;
    case CuComponent::TRANSPOSE: pRet = new CuTranspose(nInputs, nOutputs, pPred); break;
    case CuComponent::BLOCK_LINEARITY: pRet = new CuBlockLinearity(nInputs, nOutputs, pPred); break;
    case CuComponent::BIAS: pRet = new CuBias(nInputs, nOutputs, pPred); break;
    case CuComponent::WINDOW: pRet = new CuWindow(nInputs, nOutputs, pPred); break;
    case CuComponent::LOG: pRet = new CuLog(nInputs, nOutputs, pPred); break;
    case CuComponent::SHARED_LINEARITY: pRet = new CuSharedLinearity(nInputs, nOutputs, pPred); break;
    case CuComponent::DISCRETE_LINEARITY: pRet = new CuDiscreteLinearity(nInputs, nOutputs, pPred); break;
    case CuComponent::RBM_SPARSE: pRet = new CuRbmSparse(nInputs, nOutputs, pPred); break;
    default: Error(std::string("Unknown Component tag:") + componentTag);
  }
  try {
    pRet->ReadFromStream(rIn);
  } catch (...) {
    delete pRet;
    throw;
  }
  return pRet;
}

void CuNetwork::ComponentDumper(std::ostream &rOut, CuComponent &rComp) {
  const char *tag = NULL;
  for (size_t i = 0; i < sizeof(kTags) / sizeof(kTags[0]); i++)
    if (kTags[i].type == rComp.GetType()) tag = kTags[i].tag;
  if (!tag) Error("Unknown ComponentType");
  rOut << tag << " " << rComp.GetNOutputs() << " " << rComp.GetNInputs() << std::endl;
  rComp.WriteToStream(rOut);
}

// ------------------------------------------------------------------------------------------- CuCache
void CuCache::AddData(const CuMatrix<BaseFloat> &rFeatures, const CuMatrix<BaseFloat> &rDesired) {
  if (rFeatures.Rows() != rDesired.Rows()) Error("CuCache::AddData: features and targets differ in length");
  if (mFeatures.Rows() != mCachesize) {  // lazy allocation
    mFeatures.Init(mCachesize, rFeatures.Cols());
    mDesired.Init(mCachesize, rDesired.Cols());
  }
  if (rFeatures.Rows() > mCachesize / 2) {
    std::ostringstream os;
    os << "Too long segment and small feature cache! " << " cachesize: " << mCachesize << " segmentsize: " << rFeatures.Rows();
    Warning(os.str());
  }
  if (mState == EMPTY) {
    if (mTrace & 3) std::cout << "/" << std::flush;
    mState = INTAKE;
    mIntakePos = 0;
    size_t leftover = mFeaturesLeftover.Rows();
    if (leftover > mCachesize) {
      std::ostringstream os;
      os << "Too small feature cache: " << mCachesize << ", truncating: " << leftover - mCachesize << " frames from previous segment leftover";
      Warning(os.str());
      leftover = mCachesize;
    }
    if (leftover > 0) {
      mFeatures.CopyRows(leftover, 0, mFeaturesLeftover, 0);
      mDesired.CopyRows(leftover, 0, mDesiredLeftover, 0);
      mFeaturesLeftover.Destroy();
      mDesiredLeftover.Destroy();
      mIntakePos += leftover;
    }
  }
  if (mState != INTAKE) Error("CuCache::AddData on a cache that is not taking data");
  if (mTrace & 2) std::cout << "F" << std::flush;

  const size_t cache_space = mCachesize - mIntakePos;
  const size_t feature_length = rFeatures.Rows();
  const size_t fill_rows = cache_space < feature_length ? cache_space : feature_length;
  const size_t leftover = feature_length - fill_rows;
  if (cache_space == 0) Error("CuCache::AddData on a full cache");

  mFeatures.CopyRows(fill_rows, 0, rFeatures, mIntakePos);
  mDesired.CopyRows(fill_rows, 0, rDesired, mIntakePos);
  if (leftover > 0) {
    mFeaturesLeftover.Init(leftover, mFeatures.Cols());
    mDesiredLeftover.Init(leftover, mDesired.Cols());
    mFeaturesLeftover.CopyRows(leftover, fill_rows, rFeatures, 0);
    mDesiredLeftover.CopyRows(leftover, fill_rows, rDesired, 0);
  }
  mIntakePos += fill_rows;
  if (mIntakePos == mCachesize) {
    if (mTrace & 3) std::cout << "\\" << std::flush;
    mState = FULL;
  }
}

void CuCache::Randomize() {
  if (!(mState == FULL || mState == INTAKE)) Error("CuCache::Randomize on an empty cache");
  if (mTrace & 3) std::cout << "R" << std::flush;
  mFeaturesRandom.Init(mCachesize, mFeatures.Cols());
  mDesiredRandom.Init(mCachesize, mDesired.Cols());
  // Permutation: bit-exact with the reference, which calls std::random_shuffle(p, p+n, GenerateRandom) with
  // GenerateRandom(k) = lrand48() % k (cuCache.cc:136-141, cuCache.h:47-48).  libstdc++'s algorithm is spelled out
  // instead of calling std::random_shuffle (removed in C++17): for i in 1..n-1: swap(a[i], a[rng(i+1)]).
  const size_t n = mIntakePos;
  mLastPerm.Init(n);
  int *p = mLastPerm.pData();
  for (size_t i = 0; i < n; i++) p[i] = (int)i;
  for (size_t i = 1; i < n; i++) {
    size_t j = (size_t)GenerateRandom((int)(i + 1));
    if (i != j) std::swap(p[i], p[j]);
  }
  mCuRandMask.CopyFrom(mLastPerm);
  CuMath<BaseFloat>::Randomize(mFeaturesRandom, mFeatures, mCuRandMask);
  CuMath<BaseFloat>::Randomize(mDesiredRandom, mDesired, mCuRandMask);
  mRandomized = true;
}

void CuCache::GetBunch(CuMatrix<BaseFloat> &rFeatures, CuMatrix<BaseFloat> &rDesired) {
  if (mState == EMPTY) Error("GetBunch on empty cache!!!");
  if (mState == FULL) {
    if (mTrace & 3) std::cout << "\\" << std::flush;
    mState = EXHAUST;
    mExhaustPos = 0;
  }
  if (mState == INTAKE) {  // the final cache is not completely filled
    if (mTrace & 3) std::cout << "\\-LAST\n" << std::flush;
    mState = EXHAUST;
    mExhaustPos = 0;
  }
  rFeatures.Init(mBunchsize, mFeatures.Cols());
  rDesired.Init(mBunchsize, mDesired.Cols());
  if (mRandomized) {
    rFeatures.CopyRows(mBunchsize, mExhaustPos, mFeaturesRandom, 0);
    rDesired.CopyRows(mBunchsize, mExhaustPos, mDesiredRandom, 0);
  } else {
    rFeatures.CopyRows(mBunchsize, mExhaustPos, mFeatures, 0);
    rDesired.CopyRows(mBunchsize, mExhaustPos, mDesired, 0);
  }
  mExhaustPos += mBunchsize;
  // same unsigned arithmetic as the reference (cuCache.cc:194): no more complete bunches -> count the rest as discarded
  if (mExhaustPos > mIntakePos - mBunchsize) {
    mDiscarded += (int)(mIntakePos - mExhaustPos);
    mState = EMPTY;
  }
}

}  // namespace TNet
