// main_common.h — option retrieval shared by the three trainer mains (feature-side parameters of
// UserInterface::GetFeatureParams, reference: src/KaldiLib/UserInterface.cc:361-462).
#ifndef TNETB200_MAIN_COMMON_H_
#define TNETB200_MAIN_COMMON_H_

#include <sys/time.h>

#include "cu_nnet.h"
#include "io.h"

namespace TNet {

struct FeatureParams {
  bool swap_features;
  int start_frm_ext, end_frm_ext, target_kind, deriv_order;
  const char *cmn_mask, *cvn_mask, *cvg_file;
};

inline FeatureParams GetFeatureParams(UserInterface &ui, const char *sname) {
  FeatureParams p;
  std::string s(sname);
  p.swap_features = !ui.GetBool((s + ":NATURALREADORDER").c_str(), IsBigEndian());
  p.start_frm_ext = ui.GetInt((s + ":STARTFRMEXT").c_str(), 0);
  p.end_frm_ext = ui.GetInt((s + ":ENDFRMEXT").c_str(), 0);
  p.cmn_mask = ui.GetStr((s + ":CMEANMASK").c_str(), NULL);
  ui.GetStr((s + ":CMEANDIR").c_str(), NULL);
  p.cvn_mask = ui.GetStr((s + ":VARSCALEMASK").c_str(), NULL);
  ui.GetStr((s + ":VARSCALEDIR").c_str(), NULL);
  p.cvg_file = ui.GetStr((s + ":VARSCALEFN").c_str(), NULL);
  const char *tk = ui.GetStr((s + ":TARGETKIND").c_str(), "ANON");
  p.target_kind = FeatureRepository::ReadParmKind(tk, false);
  if (p.target_kind == -1) throw std::runtime_error(std::string("Invalid TARGETKIND = '") + tk + "'");
  p.deriv_order = (p.target_kind & 0100000) ? 3 : (p.target_kind & 01000) ? 2 : (p.target_kind & 0400) ? 1 : 0;
  ui.GetInt((s + ":DELTAWINDOW").c_str(), 2);
  ui.GetInt((s + ":ACCWINDOW").c_str(), 2);
  ui.GetInt((s + ":THIRDWINDOW").c_str(), 2);
  ui.GetStr((s + ":DERIVWINDOWS").c_str(), NULL);
  return p;
}

inline void SelectMath(UserInterface &ui, const char *sname) {
  const char *math = ui.GetStr((std::string(sname) + ":MATH").c_str(), "3xtf32");
  if (!strcasecmp(math, "3xtf32")) CuDevice::Instantiate().SetMath(TNB_MATH_3XTF32);
  else if (!strcasecmp(math, "tf32")) CuDevice::Instantiate().SetMath(TNB_MATH_TF32);
  else if (!strcasecmp(math, "simt")) CuDevice::Instantiate().SetMath(TNB_MATH_FP32_SIMT);
  else if (!strcasecmp(math, "bf16")) CuDevice::Instantiate().SetMath(TNB_MATH_BF16);
  else throw std::runtime_error(std::string("Invalid MATH '") + math + "' (3xtf32, tf32, bf16, simt)");
}

inline long SeedOrTime(long seed) {
  if (seed == 0) {
    struct timeval tv;
    if (gettimeofday(&tv, 0) == -1) Error("gettimeofday does not work.");
    seed = (int)(tv.tv_sec) + (int)tv.tv_usec;
  }
  return seed;
}

/// read the current utterance, run the feature transform, trim the frame extension (TNetCu.cc:383-393)
inline void ReadTransformTrim(FeatureRepository &repo, CuNetwork &transform, const FeatureParams &fp, CuMatrix<BaseFloat> &original,
                              CuMatrix<BaseFloat> &expanded, CuMatrix<BaseFloat> &trimmed, bool check) {
  Matrix<BaseFloat> feats_host;
  repo.ReadFullMatrix(feats_host);
  if (check) feats_host.CheckData(repo.Current().Logical());
  original.CopyFrom(feats_host);
  transform.Propagate(original, expanded);
  int rows = (int)expanded.Rows() - fp.start_frm_ext - fp.end_frm_ext;
  if (rows < 1) Error(std::string("Utterance shorter than the frame extension: ") + repo.Current().Logical());
  trimmed.Init(rows, expanded.Cols());
  trimmed.CopyRows(rows, fp.start_frm_ext, expanded, 0);
}

}  // namespace TNet
#endif
