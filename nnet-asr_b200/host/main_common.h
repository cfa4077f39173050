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
  std::vector<int> deriv_win;       // window lengths of the derivatives (empty = none given)
  bool has_cmn_path, has_cvn_path;  // a mask was given: the path is "<dir>/" (or "" without a directory), as the reference builds it
  std::string cmn_path, cvn_path;
  const char *cmn_mask, *cvn_mask, *cvg_file;
  int *DerivWin() { return deriv_win.empty() ? NULL : &deriv_win[0]; }
  const char *CmnPath() const { return has_cmn_path ? cmn_path.c_str() : NULL; }
  const char *CvnPath() const { return has_cvn_path ? cvn_path.c_str() : NULL; }
};

inline FeatureParams GetFeatureParams(UserInterface &ui, const char *sname) {
  FeatureParams p;
  std::string s(sname);
  p.swap_features = !ui.GetBool((s + ":NATURALREADORDER").c_str(), IsBigEndian());
  p.start_frm_ext = ui.GetInt((s + ":STARTFRMEXT").c_str(), 0);
  p.end_frm_ext = ui.GetInt((s + ":ENDFRMEXT").c_str(), 0);
  const char *cmn_dir = ui.GetStr((s + ":CMEANDIR").c_str(), NULL);
  p.cmn_mask = ui.GetStr((s + ":CMEANMASK").c_str(), NULL);
  p.has_cmn_path = p.cmn_mask != NULL;
  if (p.has_cmn_path && cmn_dir) p.cmn_path = std::string(cmn_dir) + "/";
  const char *cvn_dir = ui.GetStr((s + ":VARSCALEDIR").c_str(), NULL);
  p.cvn_mask = ui.GetStr((s + ":VARSCALEMASK").c_str(), NULL);
  p.has_cvn_path = p.cvn_mask != NULL;
  if (p.has_cvn_path && cvn_dir) p.cvn_path = std::string(cvn_dir) + "/";
  p.cvg_file = ui.GetStr((s + ":VARSCALEFN").c_str(), NULL);
  const char *tk = ui.GetStr((s + ":TARGETKIND").c_str(), "ANON");
  p.target_kind = FeatureRepository::ReadParmKind(tk, false);
  if (p.target_kind == -1) throw std::runtime_error(std::string("Invalid TARGETKIND = '") + tk + "'");
  // DERIVWINDOWS ("2_2_2") sets the number of derivatives and their windows and leaves the three single-window parameters unread
  // (UserInterface.cc:421-443); otherwise the order comes from TARGETKIND's _D/_A/_T and the windows from DELTAWINDOW / ACCWINDOW /
  // THIRDWINDOW (default 2); a plain ANON target takes whatever the first file has (order -1)
  const char *dw = ui.GetStr((s + ":DERIVWINDOWS").c_str(), NULL);
  if (dw) {
    std::string str(dw);
    size_t pos = 0;
    while ((pos = str.find_first_not_of(" \t_", pos)) != std::string::npos) {
      size_t end = str.find_first_of(" \t_", pos);
      const std::string tok = str.substr(pos, end == std::string::npos ? std::string::npos : end - pos);
      char *ep;
      long v = strtol(tok.c_str(), &ep, 0);
      if (tok.empty() || *ep) throw std::runtime_error("Integers separated by '_' expected for parameter DERIVWINDOWS");
      p.deriv_win.push_back((int)v);
      pos = end == std::string::npos ? str.size() : end;
    }
    p.deriv_order = (int)p.deriv_win.size();
    return p;
  }
  p.deriv_order = (p.target_kind & 0100000) ? 3 : (p.target_kind & 01000) ? 2 : (p.target_kind & 0400) ? 1 : 0;
  if (p.deriv_order || p.target_kind != 12 /* ANON */) {
    p.deriv_win.push_back(ui.GetInt((s + ":DELTAWINDOW").c_str(), 2));
    p.deriv_win.push_back(ui.GetInt((s + ":ACCWINDOW").c_str(), 2));
    p.deriv_win.push_back(ui.GetInt((s + ":THIRDWINDOW").c_str(), 2));
    return p;
  }
  p.deriv_order = -1;
  return p;
}

/// FeatureRepository::Init with the parameters above (the trainers' call, TNetCu.cc:288-296)
inline void InitFeatureRepository(FeatureRepository &repo, FeatureParams &fp) {
  // the repository copies `deriv_order` window lengths: with DELTAWINDOW / ACCWINDOW / THIRDWINDOW all three exist whatever the order
  std::vector<int> wins = fp.deriv_win;
  repo.Init(fp.swap_features, fp.start_frm_ext, fp.end_frm_ext, fp.target_kind, fp.deriv_order, wins.empty() ? NULL : &wins[0], fp.CmnPath(), fp.cmn_mask,
            fp.CvnPath(), fp.cvn_mask, fp.cvg_file);
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
