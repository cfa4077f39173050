// TRbmCu — drop-in replacement of the reference's RBM CD-1 pre-trainer (reference: src/TRbmCu.cc:99-396).
// One <rbm> layer (Bernoulli or Gaussian units), cache + shuffle as in TNetCu, per bunch: propagate, sample the hidden
// layer with the per-element Hybrid-Taus generator, reconstruct, propagate again, CD-1 update, MSE report.
#include "main_common.h"

using namespace TNet;
#define SNAME "TRBM"

int main(int argc, char *argv[]) try {
  const char *p_option_string =
      " -n r   LEARNINGRATE"
      " -D n   PRINTCONFIG=TRUE"
      " -H l   SOURCEMMF"
      " -S l   SCRIPT"
      " -T r   TRACE"
      " -V n   PRINTVERSION=TRUE";
  if (argc == 1) {
    fprintf(stderr, "\nUSAGE: %s [options] DataFiles...\n -n f learning rate (0.10)  -H mmf  -S scp  -T trace  -D  -V  -A  -C cf\n"
                    "BUNCHSIZE CACHESIZE FEATURETRANSFORM GPUSELECT LEARNINGRATE MOMENTUM NATURALREADORDER PRINTCONFIG PRINTVERSION RANDOMIZE "
                    "SCRIPT SEED SOURCEMMF TARGETMMF TRACE WEIGHTCOST MATH\nSTARTFRMEXT ENDFRMEXT TARGETKIND ...\n\n", argv[0]);
    return 1;
  }
  UserInterface ui;
  FeatureRepository feature_repo;
  Timer timer, timer_frontend;
  double time_frontend = 0.0;
  int args_parsed = ui.ParseOptions(argc, argv, p_option_string, SNAME);
  FeatureParams fp = GetFeatureParams(ui, SNAME);
  const char *p_source_mmf_file = ui.GetStr(SNAME ":SOURCEMMF", NULL);
  const char *p_input_transform = ui.GetStr(SNAME ":FEATURETRANSFORM", NULL);
  const char *p_targetmmf = ui.GetStr(SNAME ":TARGETMMF", NULL);
  const char *p_script = ui.GetStr(SNAME ":SCRIPT", NULL);
  BaseFloat learning_rate = ui.GetFlt(SNAME ":LEARNINGRATE", 0.10f);   // defaults: TRbmCu.cc:163-181
  BaseFloat momentum = ui.GetFlt(SNAME ":MOMENTUM", 0.50f);
  BaseFloat weightcost = ui.GetFlt(SNAME ":WEIGHTCOST", 0.0002f);
  int bunch_size = ui.GetInt(SNAME ":BUNCHSIZE", 256);
  int cache_size = ui.GetInt(SNAME ":CACHESIZE", 12800);
  bool randomize = ui.GetBool(SNAME ":RANDOMIZE", true);
  long int seed = ui.GetInt(SNAME ":SEED", 0);
  int trace = ui.GetInt(SNAME ":TRACE", 0);
  if (trace & 4) CuDevice::Instantiate().Verbose(true);
  int gpu_select = ui.GetInt(SNAME ":GPUSELECT", -1);
  if (gpu_select >= 0) CuDevice::Instantiate().SelectGPU(gpu_select);
  SelectMath(ui, SNAME);
  if (ui.GetBool(SNAME ":PRINTCONFIG", false)) { std::cout << std::endl; ui.PrintConfig(std::cout); std::cout << std::endl; }
  if (ui.GetBool(SNAME ":PRINTVERSION", false)) std::cout << std::endl << "======= TRbmCu (B200) =======" << std::endl << std::endl;
  ui.CheckCommandLineParamUse();
  for (; args_parsed < argc; args_parsed++) feature_repo.AddFile(argv[args_parsed]);

  CuNetwork network, transform_network;
  if (NULL != p_input_transform) transform_network.ReadNetwork(p_input_transform);
  if (NULL != p_source_mmf_file) network.ReadNetwork(p_source_mmf_file);
  else Error("Source MMF must be specified [-H]");
  if (network.Layers() != 1) Error(std::string("Number of layers must be 1") + p_source_mmf_file);
  if (network.Layer(0).GetType() != CuComponent::RBM && network.Layer(0).GetType() != CuComponent::RBM_SPARSE)
    Error(std::string("Layer must be RBM") + p_source_mmf_file);
  CuRbmBase &rbm = dynamic_cast<CuRbmBase &>(network.Layer(0));

  InitFeatureRepository(feature_repo, fp);
  if (NULL != p_script) feature_repo.AddFileList(p_script);
  else Warning("WARNING: The script file is missing [-S]");
  feature_repo.Trace(trace);

  rbm.LearnRate(learning_rate);
  rbm.Momentum(momentum);
  rbm.Weightcost(weightcost);
  srand48(SeedOrTime(seed));
  // the generator state is drawn from lrand48() here: after srand48, before the first cache shuffle (TRbmCu.cc:261-264)
  CuRand<BaseFloat> cu_rand(bunch_size, rbm.GetNOutputs());
  CuMeanSquareError mse;

  timer.Start();
  cache_size = (cache_size / bunch_size) * bunch_size;
  CuCache cache;
  cache.Init(cache_size, bunch_size);
  cache.Trace(trace);
  feature_repo.Rewind();

  CuMatrix<BaseFloat> pos_vis, pos_hid, neg_vis, neg_hid, dummy_labs, dummy_err;
  CuMatrix<BaseFloat> feats_original, feats_expanded, feats_trim, labs_cu;
  while (!feature_repo.EndOfList()) {
    timer_frontend.Start();
    while (!cache.Full() && !feature_repo.EndOfList()) {
      ReadTransformTrim(feature_repo, transform_network, fp, feats_original, feats_expanded, feats_trim, false);
      labs_cu.Init(feats_trim.Rows(), 1);  // "fake the labels" (TRbmCu.cc:313)
      cache.AddData(feats_trim, labs_cu);
      feature_repo.MoveNext();
    }
    timer_frontend.End();
    time_frontend += timer_frontend.Val();
    if (randomize) cache.Randomize();
    bool any = false;
    while (!cache.Empty()) {
      cache.GetBunch(pos_vis, dummy_labs);
      rbm.Propagate(pos_vis, pos_hid);
      if (rbm.HidType() == CuRbmBase::BERNOULLI) {
        cu_rand.BinarizeProbs(pos_hid, neg_hid);
      } else {
        neg_hid.CopyFrom(pos_hid);
        cu_rand.AddGaussNoise(neg_hid);
      }
      rbm.Reconstruct(neg_hid, neg_vis);
      rbm.Propagate(neg_vis, neg_hid);
      rbm.RbmUpdate(pos_vis, pos_hid, neg_vis, neg_hid);
      mse.Evaluate(neg_vis, pos_vis, dummy_err);
      any = true;
      if (trace & 2) std::cout << "." << std::flush;
    }
    if (any) pos_hid.CheckData();  // NaN/Inf guard once per cache (TRbmCu.cc:356)
  }
  if (trace & 1) TraceLog("Training finished");
  if (NULL != p_targetmmf) network.WriteNetwork(p_targetmmf);
  else Error("missing argument --TARGETMMF");
  size_t frames = mse.GetFrames();
  timer.End();
  std::cout << "===== TRbmCu FINISHED ( " << timer.Val() << "s ) " << "[FPS:" << frames / timer.Val() << ",RT:"
            << 1.0f / (frames / timer.Val() / 100.0f) << "] =====" << std::endl;
  std::cout << mse.Report();
  if (trace & 4) std::cout << "\n== PROFILE ==\nT-fe: " << time_frontend << std::endl;
  return 0;
} catch (std::exception &rExc) {
  std::cerr << "Exception thrown" << std::endl;
  std::cerr << rExc.what() << std::endl;
  return 1;
}
