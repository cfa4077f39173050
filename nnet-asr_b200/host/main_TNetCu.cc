// TNetCu — drop-in replacement of the reference's GPU MLP trainer (reference: src/TNetCu.cc:99-488).
// Same flags and files, same stdout report lines (the newbob scheduler greps "correct[..%]",
// tools/train/training_scheduler.sh:48), same cache / shuffle / bunch semantics; the bunch step runs on the
// sm_100a kernels behind libtnetb200.so.
//
// Extensions (ignored by reference scripts): --MATH=3xtf32|tf32|simt selects the GEMM arithmetic,
// --FUSE=TRUE|FALSE toggles the fused network traversal.
#include <sys/time.h>

#include "cu_nnet.h"
#include "io.h"

using namespace TNet;

#define SNAME "TNET"
static const char *kVersion = "1.8-b200";

static void usage(const char *progname) {
  const char *tchrptr;
  if ((tchrptr = strrchr(progname, '/')) != NULL) progname = tchrptr + 1;
  fprintf(stderr,
          "\n%s version %s (B200)\n"
          "\nUSAGE: %s [options] DataFiles...\n\n"
          " Option                                                     Default\n\n"
          " -c         Enable crossvalidation                          off\n"
          " -m file    Set label map of NN outputs                     \n"
          " -n f       Set learning rate to f                          0.06\n"
          " -o ext     Set target model ext                            None\n"
          " -A         Print command line arguments                    Off\n"
          " -C cf      Set config file to cf                           Default\n"
          " -D         Display configuration variables                 Off\n"
          " -H mmf     Load NN macro file                              \n"
          " -I mlf     Load master label file mlf                      \n"
          " -L dir     Set input label (or net) dir                    Current\n"
          " -M dir     Dir to write NN macro files                     Current\n"
          " -O fn      Objective function [mse,xent]                   xent\n"
          " -S file    Set script file                                 None\n"
          " -T N       Set trace flags to N                            0\n"
          " -V         Print version information                       Off\n"
          " -X ext     Set input label file ext                        lab\n"
          "\n"
          "BUNCHSIZE CACHESIZE CROSSVALIDATE FEATURETRANSFORM GPUSELECT GRADDIVFRM L1 LEARNINGRATE LEARNRATEFACTORS MLFTRANSC MOMENTUM "
          "NATURALREADORDER OBJECTIVEFUNCTION OUTPUTLABELMAP PRINTCONFIG PRINTVERSION RANDOMIZE SCRIPT SEED SOURCEMLF SOURCEMMF "
          "SOURCETRANSCDIR SOURCETRANSCEXT TARGETMMF TARGETMODELDIR TARGETMODELEXT TRACE WEIGHTCOST MATH FUSE\n"
          "\n"
          "STARTFRMEXT ENDFRMEXT CMEANDIR CMEANMASK VARSCALEDIR VARSCALEMASK VARSCALEFN TARGETKIND DERIVWINDOWS DELTAWINDOW ACCWINDOW "
          "THIRDWINDOW TEMPBASISFOLDER\n\n",
          progname, kVersion, progname);
  exit(-1);
}

int main(int argc, char *argv[]) try {
  const char *p_option_string =
      " -c n   CROSSVALIDATE=TRUE"
      " -m r   OUTPUTLABELMAP"
      " -n r   LEARNINGRATE"
      " -o r   TARGETMODELEXT"
      " -D n   PRINTCONFIG=TRUE"
      " -H l   SOURCEMMF"
      " -I r   SOURCEMLF"
      " -L r   SOURCETRANSCDIR"
      " -M r   TARGETMODELDIR"
      " -O r   OBJECTIVEFUNCTION"
      " -S l   SCRIPT"
      " -T r   TRACE"
      " -V n   PRINTVERSION=TRUE"
      " -X r   SOURCETRANSCEXT";

  UserInterface ui;
  FeatureRepository feature_repo;
  LabelRepository label_repo;
  Timer timer, timer_frontend;
  double time_frontend = 0.0, time_read = 0.0, time_xform = 0.0, time_labels = 0.0, time_cache = 0.0, time_write = 0.0;

  if (argc == 1) usage(argv[0]);
  int args_parsed = ui.ParseOptions(argc, argv, p_option_string, SNAME);

  // ---- option retrieval (defaults: TNetCu.cc:192-248) ----
  bool swap_features = !ui.GetBool(SNAME ":NATURALREADORDER", IsBigEndian());
  int start_frm_ext = ui.GetInt(SNAME ":STARTFRMEXT", 0);
  int end_frm_ext = ui.GetInt(SNAME ":ENDFRMEXT", 0);
  const char *cmn_mask = ui.GetStr(SNAME ":CMEANMASK", NULL);
  ui.GetStr(SNAME ":CMEANDIR", NULL);
  const char *cvn_mask = ui.GetStr(SNAME ":VARSCALEMASK", NULL);
  ui.GetStr(SNAME ":VARSCALEDIR", NULL);
  const char *cvg_file = ui.GetStr(SNAME ":VARSCALEFN", NULL);
  const char *target_kind_str = ui.GetStr(SNAME ":TARGETKIND", "ANON");
  int target_kind = FeatureRepository::ReadParmKind(target_kind_str, false);
  if (target_kind == -1) throw std::runtime_error(std::string("Invalid TARGETKIND = '") + target_kind_str + "'");
  int deriv_order = (target_kind & 0100000) ? 3 : (target_kind & 01000) ? 2 : (target_kind & 0400) ? 1 : 0;
  ui.GetInt(SNAME ":DELTAWINDOW", 2); ui.GetInt(SNAME ":ACCWINDOW", 2); ui.GetInt(SNAME ":THIRDWINDOW", 2);
  ui.GetStr(SNAME ":DERIVWINDOWS", NULL);

  const char *p_source_mmf_file = ui.GetStr(SNAME ":SOURCEMMF", NULL);
  const char *p_input_transform = ui.GetStr(SNAME ":FEATURETRANSFORM", NULL);
  const char *p_targetmmf = ui.GetStr(SNAME ":TARGETMMF", NULL);
  const char *p_trg_mmf_dir = ui.GetStr(SNAME ":TARGETMODELDIR", "");
  const char *p_trg_mmf_ext = ui.GetStr(SNAME ":TARGETMODELEXT", "");
  const char *p_script = ui.GetStr(SNAME ":SCRIPT", NULL);
  const char *p_output_label_map = ui.GetStr(SNAME ":OUTPUTLABELMAP", NULL);
  BaseFloat learning_rate = ui.GetFlt(SNAME ":LEARNINGRATE", 0.06f);
  const char *learning_rate_factors = ui.GetStr(SNAME ":LEARNRATEFACTORS", NULL);
  BaseFloat momentum = ui.GetFlt(SNAME ":MOMENTUM", 0.0);
  BaseFloat weightcost = ui.GetFlt(SNAME ":WEIGHTCOST", 0.0);
  BaseFloat l1 = ui.GetFlt(SNAME ":L1", 0.0);
  bool grad_div_frm = ui.GetBool(SNAME ":GRADDIVFRM", true);
  const char *objfun = ui.GetStr(SNAME ":OBJECTIVEFUNCTION", "xent");
  CuObjectiveFunction::ObjFunType obj_fun_id;
  if (!strcmp(objfun, "xent")) obj_fun_id = CuObjectiveFunction::CROSS_ENTROPY;
  else if (!strcmp(objfun, "mse")) obj_fun_id = CuObjectiveFunction::MEAN_SQUARE_ERROR;
  else throw std::runtime_error(std::string("Invalid OBJECTIVEFUNCTION '") + objfun + "' (xent, mse)");
  const char *p_source_mlf_file = ui.GetStr(SNAME ":SOURCEMLF", NULL);
  const char *p_src_lbl_dir = ui.GetStr(SNAME ":SOURCETRANSCDIR", NULL);
  const char *p_src_lbl_ext = ui.GetStr(SNAME ":SOURCETRANSCEXT", "lab");
  bool mlf_transc = ui.GetBool(SNAME ":MLFTRANSC", true);
  int bunch_size = ui.GetInt(SNAME ":BUNCHSIZE", 256);
  int cache_size = ui.GetInt(SNAME ":CACHESIZE", 12800);
  bool randomize = ui.GetBool(SNAME ":RANDOMIZE", true);
  long int seed = ui.GetInt(SNAME ":SEED", 0);
  bool cross_validate = ui.GetBool(SNAME ":CROSSVALIDATE", false);
  int trace = ui.GetInt(SNAME ":TRACE", 0);
  if (trace & 4) CuDevice::Instantiate().Verbose(true);
  int gpu_select = ui.GetInt(SNAME ":GPUSELECT", -1);
  if (gpu_select >= 0) CuDevice::Instantiate().SelectGPU(gpu_select);
  ui.GetStr(SNAME ":TEMPBASISFOLDER", NULL);  // <clusterlinearity> only; accepted
  const char *math = ui.GetStr(SNAME ":MATH", "3xtf32");
  if (!strcasecmp(math, "3xtf32")) CuDevice::Instantiate().SetMath(TNB_MATH_3XTF32);
  else if (!strcasecmp(math, "tf32")) CuDevice::Instantiate().SetMath(TNB_MATH_TF32);
  else if (!strcasecmp(math, "simt")) CuDevice::Instantiate().SetMath(TNB_MATH_FP32_SIMT);
  else if (!strcasecmp(math, "bf16")) CuDevice::Instantiate().SetMath(TNB_MATH_BF16);
  else throw std::runtime_error(std::string("Invalid MATH '") + math + "' (3xtf32, tf32, bf16, simt)");
  bool fuse = ui.GetBool(SNAME ":FUSE", true);

  if (ui.GetBool(SNAME ":PRINTCONFIG", false)) { std::cout << std::endl; ui.PrintConfig(std::cout); std::cout << std::endl; }
  if (ui.GetBool(SNAME ":PRINTVERSION", false)) std::cout << std::endl << "======= TNET v" << kVersion << " =======" << std::endl << std::endl;
  ui.CheckCommandLineParamUse();

  for (; args_parsed < argc; args_parsed++) feature_repo.AddFile(argv[args_parsed]);

  // ---- networks ----
  CuNetwork network, transform_network;
  network.SetFusion(fuse);
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

  feature_repo.Init(swap_features, start_frm_ext, end_frm_ext, target_kind, deriv_order, NULL, NULL, cmn_mask, NULL, cvn_mask, cvg_file);
  feature_repo.Trace(trace);
  if (NULL != p_script) feature_repo.AddFileList(p_script);
  else Warning("WARNING: The script file is missing [-S]");

  if (mlf_transc) {
    if (NULL == p_source_mlf_file) Error("Source mlf file file is missing [-I]");
    if (NULL == p_output_label_map) Error("Output label map is missing [-m]");
    if (trace & 1) TraceLog(std::string("Indexing labels: ") + p_source_mlf_file);
    label_repo.Init(p_source_mlf_file, p_output_label_map, p_src_lbl_dir, p_src_lbl_ext);
    label_repo.Trace(trace);
  } else {
    Error("MLFTRANSC=FALSE (targets from HTK matrix files) is not built into the B200 hot path");
  }

  CuObjectiveFunction *p_obj_function = CuObjectiveFunction::Factory(obj_fun_id);
  network.SetLearnRate(learning_rate, learning_rate_factors);
  network.SetMomentum(momentum);
  network.SetWeightcost(weightcost);
  network.SetL1(l1);
  network.SetGradDivFrm(grad_div_frm);

  if (seed == 0) {
    struct timeval tv;
    if (gettimeofday(&tv, 0) == -1) Error("gettimeofday does not work.");
    seed = (int)(tv.tv_sec) + (int)tv.tv_usec;
  }
  srand48(seed);

  // ---- training ----
  timer.Start();
  std::cout << "===== TNET " << (cross_validate ? "CROSSVALIDATION" : "TRAINING") << " STARTED =====" << std::endl;
  std::cout << "Objective function: " << p_obj_function->GetTypeLabel() << std::endl;
  if (!cross_validate) {
    network.PrintLearnRate();
    std::cout << "momentum: " << momentum << " weightcost: " << weightcost << std::endl;
    std::cout << "using seed: " << seed << std::endl;
  }
  cache_size = (cache_size / bunch_size) * bunch_size;
  std::cout << "Bunchsize:" << bunch_size << " Cachesize:" << cache_size << "\n";

  CuCache cache;
  cache.Init(cache_size, bunch_size);
  cache.Trace(trace);
  feature_repo.Rewind();

  CuMatrix<BaseFloat> feats, labs, globerr;
  CuMatrix<BaseFloat> feats_original, feats_expanded, feats_trim, labs_cu;
  CuVector<int> label_ids;
  while (!feature_repo.EndOfList()) {
    timer_frontend.Start();
    while (!cache.Full() && !feature_repo.EndOfList()) {
      Matrix<BaseFloat> feats_host;
      Timer t_part;
      t_part.Start();
      feature_repo.ReadFullMatrix(feats_host);
      feats_host.CheckData(feature_repo.Current().Logical());
      t_part.End(); time_read += t_part.Val(); t_part.Start();
      feats_original.CopyFrom(feats_host);
      transform_network.Propagate(feats_original, feats_expanded);
      int rows = (int)feats_expanded.Rows() - start_frm_ext - end_frm_ext;
      if (rows < 1) Error(std::string("Utterance shorter than the frame extension: ") + feature_repo.Current().Logical());
      feats_trim.Init(rows, feats_expanded.Cols());
      feats_trim.CopyRows(rows, start_frm_ext, feats_expanded, 0);
      t_part.End(); time_xform += t_part.Val(); t_part.Start();
      // labels: class ids go to the device (4 bytes/frame instead of 4*nOutputs) and are expanded to one-hot rows there
      std::vector<int> ids;
      label_repo.GenLabelIds(ids, rows, feature_repo.CurrentHeader().mSamplePeriod, feature_repo.Current().Logical().c_str());
      Vector<int> ids_host(rows);
      for (int i = 0; i < rows; i++) ids_host[i] = ids[i];
      t_part.End(); time_labels += t_part.Val(); t_part.Start();
      label_ids.CopyFrom(ids_host);
      labs_cu.Init(rows, label_repo.NOutputs());
      TNB_CHECK(tnb_onehot(Cx(), labs_cu.pCUData(), label_ids.pCUData(), labs_cu.Dim()));
      if (labs_cu.Cols() != network.GetNOutputs() && obj_fun_id == CuObjectiveFunction::CROSS_ENTROPY) {
        std::ostringstream os;
        os << "Non-matching dimensions of network output with training targets!!!" << " Netoutput:" << network.GetNOutputs()
           << " Targets:" << labs_cu.Cols();
        Error(os.str());
      }
      cache.AddData(feats_trim, labs_cu);
      t_part.End(); time_cache += t_part.Val();
      feature_repo.MoveNext();
    }
    timer_frontend.End();
    time_frontend += timer_frontend.Val();

    if (randomize) cache.Randomize();

    while (!cache.Empty()) {
      cache.GetBunch(feats, labs);
      network.PropagateEvaluate(feats, labs, *p_obj_function, globerr);
      if (!cross_validate) network.Backpropagate(globerr);
      if (trace & 2) std::cout << "." << std::flush;
    }
  }
  if (trace & 1) TraceLog("Training finished");

  Timer t_write;
  t_write.Start();
  if (!cross_validate) {
    char p_trg_mmf_file[4096];
    if (NULL != p_targetmmf) {
      if (trace & 1) TraceLog(std::string("Writing network: ") + p_targetmmf);
      network.WriteNetwork(p_targetmmf);
    } else {
      MakeHtkFileName(p_trg_mmf_file, p_source_mmf_file, p_trg_mmf_dir, p_trg_mmf_ext);
      if (trace & 1) TraceLog(std::string("Writing network: ") + p_trg_mmf_file);
      network.WriteNetwork(p_trg_mmf_file);
    }
  }
  t_write.End();
  time_write = t_write.Val();
  size_t frames = p_obj_function->GetFrames();  // reads the device accumulators (synchronises)
  timer.End();
  std::cout << "===== TNET " << (cross_validate ? "CROSSVALIDATION" : "TRAINING") << " FINISHED ( " << timer.Val() << "s ) "
            << "[FPS:" << frames / timer.Val() << ",RT:" << 1.0f / (frames / timer.Val() / 100.0f) << "] =====" << std::endl;
  std::cout << "-- " << (cross_validate ? "CV " : "TR ") << p_obj_function->Report();
  if (trace & 4) std::cout << "\n== PROFILE ==\nT-fe: " << time_frontend << " (read+check " << time_read << ", H2D+transform " << time_xform << ", labels "
                           << time_labels << ", one-hot+cache " << time_cache << ")\nT-write(sync+network file): " << time_write << "\nkernel launches: " << CuDevice::Instantiate().Launches() << std::endl;
  delete p_obj_function;
  return 0;
} catch (std::exception &rExc) {
  std::cerr << "Exception thrown" << std::endl;
  std::cerr << rExc.what() << std::endl;
  return 1;
}
