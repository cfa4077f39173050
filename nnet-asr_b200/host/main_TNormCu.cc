// TNormCu — drop-in replacement of the reference's global mean/variance normalisation estimator (reference: src/TNormCu.cc:75-345):
// run every utterance through the feature transform given with -H, accumulate first and second order statistics of the output
// over all frames, and write the normalisation as a <bias> (negative mean) + <window> (1/sqrt(variance)) transform.
// The statistics stay on the device (tnb_accum_moments, double accumulators) instead of a D2H copy of every transformed utterance.
#include <math.h>

#include "main_common.h"

using namespace TNet;
#define SNAME "TNORM"

int main(int argc, char *argv[]) try {
  const char *p_option_string =
      " -D n   PRINTCONFIG=TRUE"
      " -H l   SOURCEMMF"
      " -S l   SCRIPT"
      " -T r   TRACE"
      " -V n   PRINTVERSION=TRUE";
  if (argc == 1) {
    fprintf(stderr, "\nUSAGE: %s [options] DataFiles...\n -H mmf (the transform whose output is normalised)  -S scp  -T trace  -D  -V  -A  -C cf\n"
                    "NATURALREADORDER PRINTCONFIG PRINTVERSION SCRIPT SOURCEMMF TARGETMMF TRACE GPUSELECT\nSTARTFRMEXT ENDFRMEXT TARGETKIND ...\n\n", argv[0]);
    return 1;
  }
  UserInterface ui;
  FeatureRepository features;
  Timer timer;
  int args_parsed = ui.ParseOptions(argc, argv, p_option_string, SNAME);
  FeatureParams fp = GetFeatureParams(ui, SNAME);
  const char *p_source_mmf_file = ui.GetStr(SNAME ":SOURCEMMF", NULL);
  const char *p_targetmmf = ui.GetStr(SNAME ":TARGETMMF", NULL);
  const char *p_script = ui.GetStr(SNAME ":SCRIPT", NULL);
  int trace = ui.GetInt(SNAME ":TRACE", 0);
  int gpu_select = ui.GetInt(SNAME ":GPUSELECT", -1);
  if (gpu_select >= 0) CuDevice::Instantiate().SelectGPU(gpu_select);
  if (ui.GetBool(SNAME ":PRINTCONFIG", false)) { std::cout << std::endl; ui.PrintConfig(std::cout); std::cout << std::endl; }
  if (ui.GetBool(SNAME ":PRINTVERSION", false)) std::cout << std::endl << "======= TNormCu (B200) =======" << std::endl << std::endl;
  ui.CheckCommandLineParamUse();
  for (; args_parsed < argc; args_parsed++) features.AddFile(argv[args_parsed]);

  CuNetwork network;
  if (NULL != p_source_mmf_file) {
    if (trace & 1) TraceLog(std::string("Reading network: ") + p_source_mmf_file);
    network.ReadNetwork(p_source_mmf_file);
  } else {
    Error("Source MMF must be specified [-H]");
  }
  if (NULL == p_targetmmf) Error("forgot to specify --TARGETMMF argument");
  InitFeatureRepository(features, fp);
  if (NULL != p_script) features.AddFileList(p_script);
  else Warning("WARNING: The script file is missing [-S]");

  timer.Start();
  std::cout << "===== TNormCu STARTED =====" << std::endl;
  const size_t dim = network.GetNOutputs();
  void *p_first = NULL, *p_second = NULL;
  TNB_CHECK(tnb_malloc(Cx(), &p_first, sizeof(double) * dim));   // zero-filled
  TNB_CHECK(tnb_malloc(Cx(), &p_second, sizeof(double) * dim));
  unsigned long framesN = 0;
  size_t cnt = 0, step = features.QueueSize() / 100;
  if (step == 0) step = 1;
  Matrix<BaseFloat> feats_host;
  CuMatrix<BaseFloat> feats, feats_expanded;
  for (features.Rewind(); !features.EndOfList(); features.MoveNext()) {
    features.ReadFullMatrix(feats_host);
    feats.CopyFrom(feats_host);
    network.Propagate(feats, feats_expanded);
    const int rows = (int)feats_expanded.Rows() - fp.start_frm_ext - fp.end_frm_ext;
    if (rows < 1) Error(std::string("Utterance shorter than the frame extension: ") + features.Current().Logical());
    TnbMatrixDim d = {rows, (int)feats_expanded.Cols(), (int)feats_expanded.Stride()};
    TNB_CHECK(tnb_accum_moments(Cx(), feats_expanded.pCURowData(fp.start_frm_ext), d, (double *)p_first, (double *)p_second));
    framesN += feats_host.Rows();  // as the reference: counts the replicated extension rows too (TNormCu.cc:292)
    if ((cnt++ % step) == 0) std::cout << 100 * cnt / features.QueueSize() << "%, " << std::flush;
  }
  std::vector<double> h1(dim), h2(dim);
  TNB_CHECK(tnb_memcpy(Cx(), h1.data(), p_first, sizeof(double) * dim, 1));
  TNB_CHECK(tnb_memcpy(Cx(), h2.data(), p_second, sizeof(double) * dim, 1));
  tnb_free(Cx(), p_first);
  tnb_free(Cx(), p_second);
  for (size_t i = 0; i < dim; i++)
    if (std::isnan(h1[i]) || std::isnan(h2[i]) || std::isinf(h1[i]) || std::isinf(h2[i])) Error("nan/inf in accumulators");

  // mean / variance -> <bias> = -mean, <window> = 1/sqrt(variance)   (TNormCu.cc:300-328)
  Vector<double> bias(dim), window(dim);
  double max_bias = -1e300, max_window = -1e300, min_window = 1e300;
  for (size_t i = 0; i < dim; i++) {
    const double mean = h1[i] * (1.0 / framesN);
    double variance = h2[i] * (1.0 / framesN);
    variance -= mean * mean;
    bias[i] = mean * -1.0;
    window[i] = 1.0 / sqrt(variance);
    if (bias[i] > max_bias) max_bias = bias[i];
    if (window[i] > max_window) max_window = window[i];
    if (window[i] < min_window) min_window = window[i];
  }
  std::ofstream os(p_targetmmf);
  if (!os.good()) Error(std::string("Cannot open file for writing: ") + p_targetmmf);
  os << "<bias> " << dim << " " << dim << "\n" << bias << "\n\n" << "<window> " << dim << " " << dim << "\n" << window << "\n\n";
  os.close();
  timer.End();
  std::cout << "\n\n===== TNormCu FINISHED ( " << timer.Val() << "s ) " << "[FPS:" << framesN / timer.Val() << ",RT:" << 1.0f / (framesN / timer.Val() / 100.0f)
            << "] =====" << std::endl;
  std::cout << "frames: " << framesN << ", max_bias: " << max_bias << ", max_window: " << max_window << ", min_window: " << min_window << "\n";
  return 0;
} catch (std::exception &rExc) {
  std::cerr << "Exception thrown" << std::endl;
  std::cerr << rExc.what() << std::endl;
  return 1;
}
