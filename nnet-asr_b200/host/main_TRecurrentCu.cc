// TRecurrentCu — drop-in replacement of the reference's recurrent-layer trainer (reference: src/TRecurrentCu.cc:99-420).
// Frame-serial: per utterance the history is cleared, then every frame is propagated, evaluated and back-propagated with
// truncated BPTT inside CuRecurrent::Update().  BUNCHSIZE/CACHESIZE/RANDOMIZE/SEED are accepted and ignored, as upstream.
#include "main_common.h"

using namespace TNet;
#define SNAME "TNET"

int main(int argc, char *argv[]) try {
  const char *p_option_string =
      " -c n   CROSSVALIDATE=TRUE"
      " -m r   OUTPUTLABELMAP"
      " -n r   LEARNINGRATE"
      " -D n   PRINTCONFIG=TRUE"
      " -H l   SOURCEMMF"
      " -I r   SOURCEMLF"
      " -L r   SOURCETRANSCDIR"
      " -O r   OBJECTIVEFUNCTION"
      " -S l   SCRIPT"
      " -T r   TRACE"
      " -V n   PRINTVERSION=TRUE"
      " -X r   SOURCETRANSCEXT";
  if (argc == 1) {
    fprintf(stderr, "\nUSAGE: %s [options] DataFiles...\n -c crossvalidate -m labelmap -n lr -H mmf -I mlf -L dir -X ext -S scp -O xent|mse -T trace\n"
                    "BPTT BUNCHSIZE CACHESIZE CROSSVALIDATE FEATURETRANSFORM GPUSELECT LEARNINGRATE LEARNRATEFACTORS MOMENTUM NATURALREADORDER "
                    "OBJECTIVEFUNCTION OUTPUTLABELMAP RANDOMIZE SCRIPT SEED SOURCEMLF SOURCEMMF TARGETMMF TRACE WEIGHTCOST MATH\n\n", argv[0]);
    return 1;
  }
  UserInterface ui;
  FeatureRepository feature_repo;
  LabelRepository label_repo;
  Timer timer, timer_frontend;
  double time_frontend = 0.0;
  int args_parsed = ui.ParseOptions(argc, argv, p_option_string, SNAME);
  FeatureParams fp = GetFeatureParams(ui, SNAME);
  const char *p_source_mmf_file = ui.GetStr(SNAME ":SOURCEMMF", NULL);
  const char *p_input_transform = ui.GetStr(SNAME ":FEATURETRANSFORM", NULL);
  const char *p_targetmmf = ui.GetStr(SNAME ":TARGETMMF", NULL);
  const char *p_script = ui.GetStr(SNAME ":SCRIPT", NULL);
  const char *p_output_label_map = ui.GetStr(SNAME ":OUTPUTLABELMAP", NULL);
  BaseFloat learning_rate = ui.GetFlt(SNAME ":LEARNINGRATE", 0.06f);
  const char *learning_rate_factors = ui.GetStr(SNAME ":LEARNRATEFACTORS", NULL);
  BaseFloat momentum = ui.GetFlt(SNAME ":MOMENTUM", 0.0);
  BaseFloat weightcost = ui.GetFlt(SNAME ":WEIGHTCOST", 0.0);
  const char *objfun = ui.GetStr(SNAME ":OBJECTIVEFUNCTION", "xent");
  CuObjectiveFunction::ObjFunType obj_fun_id = !strcmp(objfun, "mse") ? CuObjectiveFunction::MEAN_SQUARE_ERROR : CuObjectiveFunction::CROSS_ENTROPY;
  if (strcmp(objfun, "mse") && strcmp(objfun, "xent")) throw std::runtime_error(std::string("Invalid OBJECTIVEFUNCTION '") + objfun + "'");
  const char *p_source_mlf_file = ui.GetStr(SNAME ":SOURCEMLF", NULL);
  const char *p_src_lbl_dir = ui.GetStr(SNAME ":SOURCETRANSCDIR", NULL);
  const char *p_src_lbl_ext = ui.GetStr(SNAME ":SOURCETRANSCEXT", "lab");
  int bptt = ui.GetInt(SNAME ":BPTT", 4);
  ui.GetInt(SNAME ":BUNCHSIZE", 256); ui.GetInt(SNAME ":CACHESIZE", 12800);   // parsed, unused (TRecurrentCu.cc:217-220)
  ui.GetBool(SNAME ":RANDOMIZE", true); ui.GetInt(SNAME ":SEED", 0);
  bool cross_validate = ui.GetBool(SNAME ":CROSSVALIDATE", false);
  int trace = ui.GetInt(SNAME ":TRACE", 0);
  if (trace & 4) CuDevice::Instantiate().Verbose(true);
  int gpu_select = ui.GetInt(SNAME ":GPUSELECT", -1);
  if (gpu_select >= 0) CuDevice::Instantiate().SelectGPU(gpu_select);
  SelectMath(ui, SNAME);
  bool use_graph = ui.GetBool(SNAME ":GRAPH", true);  // replay the per-frame launch sequence as a CUDA graph
  if (ui.GetBool(SNAME ":PRINTCONFIG", false)) { std::cout << std::endl; ui.PrintConfig(std::cout); std::cout << std::endl; }
  if (ui.GetBool(SNAME ":PRINTVERSION", false)) std::cout << std::endl << "======= TRecurrentCu (B200) =======" << std::endl << std::endl;
  ui.CheckCommandLineParamUse();
  for (; args_parsed < argc; args_parsed++) feature_repo.AddFile(argv[args_parsed]);

  CuNetwork network, transform_network;
  if (NULL != p_input_transform) transform_network.ReadNetwork(p_input_transform);
  if (NULL != p_source_mmf_file) network.ReadNetwork(p_source_mmf_file);
  else Error("Source MMF must be specified [-H]");
  InitFeatureRepository(feature_repo, fp);
  if (NULL != p_script) feature_repo.AddFileList(p_script);
  else Warning("WARNING: The script file is missing [-S]");
  if (NULL == p_source_mlf_file) Error("Source mlf file file is missing [-I]");
  if (NULL == p_output_label_map) Error("Output label map is missing [-m]");
  label_repo.Init(p_source_mlf_file, p_output_label_map, p_src_lbl_dir, p_src_lbl_ext);

  CuObjectiveFunction *p_obj_function = CuObjectiveFunction::Factory(obj_fun_id);
  network.SetLearnRate(learning_rate, learning_rate_factors);
  network.SetMomentum(momentum);
  network.SetWeightcost(weightcost);
  for (int i = 0; i < network.Layers(); i++)
    if (network.Layer(i).GetType() == CuComponent::RECURRENT) dynamic_cast<CuRecurrent &>(network.Layer(i)).BpttOrder(bptt);

  timer.Start();
  std::cout << (cross_validate ? "===== TRecurrentCu CROSSVAL STARTED =====" : "===== TRecurrentCu TRAINING STARTED =====") << std::endl;
  int frames = 0, eager_frames = 0;
  void *graph = NULL;
  CuMatrix<BaseFloat> feats, targets, feats_original, feats_expanded;
  CuMatrix<BaseFloat> input_row, output_row, target_row, error_row;
  CuVector<int> label_ids;
  for (feature_repo.Rewind(); !feature_repo.EndOfList(); feature_repo.MoveNext()) {
    timer_frontend.Start();
    ReadTransformTrim(feature_repo, transform_network, fp, feats_original, feats_expanded, feats, false);
    timer_frontend.End();
    time_frontend += timer_frontend.Val();
    std::vector<int> ids;
    label_repo.GenLabelIds(ids, feats.Rows(), feature_repo.CurrentHeader().mSamplePeriod, feature_repo.Current().Logical().c_str());
    Vector<int> ids_host(ids.size());
    for (size_t i = 0; i < ids.size(); i++) ids_host[i] = ids[i];
    label_ids.CopyFrom(ids_host);
    targets.Init(feats.Rows(), label_repo.NOutputs());
    TNB_CHECK(tnb_onehot(Cx(), targets.pCUData(), label_ids.pCUData(), targets.Dim()));
    for (int i = 0; i < network.Layers(); i++)
      if (network.Layer(i).GetType() == CuComponent::RECURRENT) dynamic_cast<CuRecurrent &>(network.Layer(i)).ClearHistory();
    input_row.Init(1, feats.Cols()); output_row.Init(1, network.GetNOutputs());
    target_row.Init(1, network.GetNOutputs()); error_row.Init(1, network.GetNOutputs());
    for (size_t frm = 0; frm < feats.Rows(); frm++) {
      input_row.CopyRows(1, frm, feats, 0);
      target_row.CopyRows(1, frm, targets, 0);
      // One frame = ~115 small dependent launches on fixed buffers (forward GEMV, objective, BPTT chain of GEMV + rank-1 updates,
      // weight update): after two eager frames have created every buffer, scratch and TMA descriptor, the sequence is recorded
      // into a CUDA graph once and replayed for every further frame of the run (--GRAPH=FALSE keeps the eager schedule).
      if (graph) {
        TNB_CHECK(tnb_graph_launch(Cx(), graph));
      } else if (use_graph && eager_frames >= 2) {
        TNB_CHECK(tnb_graph_begin(Cx()));
        network.Propagate(input_row, output_row);
        p_obj_function->Evaluate(output_row, target_row, error_row);
        if (!cross_validate) network.Backpropagate(error_row);
        TNB_CHECK(tnb_graph_end(Cx(), &graph));
        TNB_CHECK(tnb_graph_launch(Cx(), graph));
      } else {
        network.Propagate(input_row, output_row);
        p_obj_function->Evaluate(output_row, target_row, error_row);
        if (!cross_validate) network.Backpropagate(error_row);
        eager_frames++;
      }
    }
    frames += (int)feats.Rows();
    std::cout << "." << std::flush;
  }
  if (trace & 1) TraceLog(cross_validate ? "Crossval finished" : "Training finished");
  if (!cross_validate) {
    if (NULL != p_targetmmf) network.WriteNetwork(p_targetmmf);
    else Error("forgot to specify --TARGETMMF argument");
  }
  CuDevice::Instantiate().Sync();
  if (graph) tnb_graph_destroy(Cx(), graph);
  timer.End();
  std::cout << std::endl;
  std::cout << "===== TRecurrentCu FINISHED ( " << timer.Val() << "s ) " << "[FPS:" << float(frames) / timer.Val() << ",RT:"
            << 1.0f / (float(frames) / timer.Val() / 100.0f) << "] =====" << std::endl;
  std::cout << "-- " << (cross_validate ? "CV" : "TR") << p_obj_function->Report();
  std::cout << "T-fe: " << time_frontend << std::endl;
  delete p_obj_function;
  return 0;
} catch (std::exception &rExc) {
  std::cerr << "Exception thrown" << std::endl;
  std::cerr << rExc.what() << std::endl;
  return 1;
}
