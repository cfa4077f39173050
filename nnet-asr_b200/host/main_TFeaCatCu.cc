// TFeaCatCu — drop-in replacement of the reference's forward-only tool (reference: src/TFeaCatCu.cc:75-300): read each feature
// file, run the feature transform and the network forward on the GPU, trim the frame extension, optionally map posteriors to
// log / GMM-bypass domain, and write an HTK USER parameter file per utterance.  Used by decode.sh and by layer-wise RBM
// stacking (tools/train/rbm_train.sh:62-86).  The forward pass is the training path's: fused bias+sigmoid GEMM epilogues,
// softmax through the row kernel.
#include <math.h>

#include "main_common.h"

using namespace TNet;
#define SNAME "TFEACAT"

int main(int argc, char *argv[]) try {
  const char *p_option_string =
      " -l r   TARGETPARAMDIR"
      " -y r   TARGETPARAMEXT"
      " -D n   PRINTCONFIG=TRUE"
      " -H l   SOURCEMMF"
      " -S l   SCRIPT"
      " -T r   TRACE"
      " -V n   PRINTVERSION=TRUE";
  if (argc == 1) {
    fprintf(stderr, "\nUSAGE: %s [options] DataFiles...\n -l dir target directory  -y ext target extension (fea)  -H mmf  -S scp  -T trace  -D  -V  -A  -C cf\n"
                    "FEATURETRANSFORM GMMBYPASS LOGPOSTERIOR NATURALREADORDER PRINTCONFIG PRINTVERSION SCRIPT SOURCEMMF TARGETPARAMDIR TARGETPARAMEXT "
                    "TRACE GPUSELECT MATH\nSTARTFRMEXT ENDFRMEXT TARGETKIND ...\n\n", argv[0]);
    return 1;
  }
  UserInterface ui;
  FeatureRepository feature_repo;
  Timer tim;
  int args_parsed = ui.ParseOptions(argc, argv, p_option_string, SNAME);
  FeatureParams fp = GetFeatureParams(ui, SNAME);
  const char *p_source_mmf_file = ui.GetStr(SNAME ":SOURCEMMF", NULL);
  const char *p_input_transform = ui.GetStr(SNAME ":FEATURETRANSFORM", NULL);
  const char *p_script = ui.GetStr(SNAME ":SCRIPT", NULL);
  const char *p_target_fea_dir = ui.GetStr(SNAME ":TARGETPARAMDIR", NULL);
  const char *p_target_fea_ext = ui.GetStr(SNAME ":TARGETPARAMEXT", "fea");
  bool gmm_bypass = ui.GetBool(SNAME ":GMMBYPASS", false);
  bool log_posterior = ui.GetBool(SNAME ":LOGPOSTERIOR", false);
  int trace = ui.GetInt(SNAME ":TRACE", 0);
  if (trace & 1) CuDevice::Instantiate().Verbose(true);
  int gpu_select = ui.GetInt(SNAME ":GPUSELECT", -1);
  if (gpu_select >= 0) CuDevice::Instantiate().SelectGPU(gpu_select);
  SelectMath(ui, SNAME);
  if (ui.GetBool(SNAME ":PRINTVERSION", false)) std::cout << "Version: TFeaCatCu (B200)" << std::endl;
  if (ui.GetBool(SNAME ":PRINTCONFIG", false)) { std::cout << std::endl; ui.PrintConfig(std::cout); std::cout << std::endl; }
  ui.CheckCommandLineParamUse();
  for (; args_parsed < argc; args_parsed++) feature_repo.AddFile(argv[args_parsed]);

  CuNetwork network, transform_network;
  if (NULL != p_input_transform) {
    if (trace & 1) TraceLog(std::string("Reading input transform network: ") + p_input_transform);
    transform_network.ReadNetwork(p_input_transform);
  }
  if (NULL != p_source_mmf_file) {
    if (trace & 1) TraceLog(std::string("Reading network: ") + p_source_mmf_file);
    network.ReadNetwork(p_source_mmf_file);
  } else {
    Error("Source MMF must be specified [-H]");
  }
  InitFeatureRepository(feature_repo, fp);
  if (NULL != p_script) feature_repo.AddFileList(p_script);
  if (feature_repo.QueueSize() <= 0) KALDI_ERR << "No input features specified,\n" << " try [-S SCP] or positional argument";

  size_t cnt = 0, step = feature_repo.QueueSize() / 100;
  if (step == 0) step = 1;
  tim.Start();
  Matrix<BaseFloat> feats_in, feats_out;
  CuMatrix<BaseFloat> feats_in_cu, feats_transf_cu, feats_out_cu, feats_trim_cu;
  char p_target_fea[4096];
  for (feature_repo.Rewind(); !feature_repo.EndOfList(); feature_repo.MoveNext()) {
    feature_repo.ReadFullMatrix(feats_in);
    feats_in_cu.CopyFrom(feats_in);
    transform_network.Propagate(feats_in_cu, feats_transf_cu);  // even when empty (a copy)
    network.Propagate(feats_transf_cu, feats_out_cu);
    int rows = (int)feats_out_cu.Rows() - fp.start_frm_ext - fp.end_frm_ext;
    if (rows < 1) Error(std::string("Utterance shorter than the frame extension: ") + feature_repo.Current().Logical());
    feats_trim_cu.Init(rows, feats_out_cu.Cols());
    feats_trim_cu.CopyRows(rows, fp.start_frm_ext, feats_out_cu, 0);
    feats_trim_cu.CopyTo(feats_out);
    if (gmm_bypass)  // posteriors as features for HVite (TFeaCatCu.cc:263-269)
      for (size_t i = 0; i < feats_out.Rows(); i++)
        for (size_t j = 0; j < feats_out.Cols(); j++) feats_out(i, j) = static_cast<BaseFloat>(sqrt(-2.0 * log(feats_out(i, j))));
    if (log_posterior)
      for (size_t i = 0; i < feats_out.Rows(); i++)
        for (size_t j = 0; j < feats_out.Cols(); j++) feats_out(i, j) = static_cast<BaseFloat>(log(feats_out(i, j)));
    MakeHtkFileName(p_target_fea, feature_repo.Current().Logical().c_str(), p_target_fea_dir, p_target_fea_ext);
    feature_repo.WriteFeatureMatrix(feats_out, p_target_fea, 9 /* PARAMKIND_USER */, feature_repo.CurrentHeader().mSamplePeriod);
    if (trace & 1)
      if ((cnt++ % step) == 0) std::cout << 100 * cnt / feature_repo.QueueSize() << "%, " << std::flush;
  }
  if (trace & 1) {
    tim.End();
    std::cout << "TFeaCat finished: " << tim.Val() << "s" << std::endl;
  }
  return 0;
} catch (std::exception &rExc) {
  std::cerr << "Exception thrown" << std::endl;
  std::cerr << rExc.what() << std::endl;
  return 1;
}
