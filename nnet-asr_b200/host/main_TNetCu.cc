// TNetCu — drop-in replacement of the reference's GPU MLP trainer (reference: src/TNetCu.cc:99-488).
// Same flags and files, same stdout report lines (the newbob scheduler greps "correct[..%]",
// tools/train/training_scheduler.sh:48), same cache / shuffle / bunch semantics; the bunch step runs on the
// sm_100a kernels behind libtnetb200.so.
//
// Extensions (ignored by reference scripts): --MATH=3xtf32|tf32|bf16|simt selects the GEMM arithmetic, --FUSE=TRUE|FALSE toggles the
// fused network traversal, --GPUS=N trains data-parallel on N GPUs of this box, --LOADER=FALSE reads and trains in turns.
//
// How the loop of TNetCu.cc:377-441 is laid out here:
//   * a LOADER THREAD (own context = own streams on GPU 0) reads utterances, runs the feature transform, turns the transcription into
//     class ids and fills cache k while the trainer exhausts cache 1-k; the rows that did not fit are handed to the next fill, so the
//     two cache objects hold, fill after fill, exactly what the reference's single cache holds, and the shuffles draw lrand48() in the
//     same order (only the loader calls it);
//   * the cache keeps ONE CLASS ID per frame (a [rows x 1] column) instead of a dense one-hot row (Labels.cc:66,156 / cuCache.cc:124-153):
//     at 3000 classes a fill of 131 072 frames moves 0.26 GB instead of 3.4 GB, and the objective kernel needs no target read
//     (dense rows remain for the MSE objective and for MLFTRANSC=FALSE, where the targets are HTK matrix files);
//   * --GPUS=N: ONE cache and ONE permutation (bit-identical to the single-GPU run); GPU g trains rows [g*B/N, (g+1)*B/N) of every
//     bunch (the reference CPU trainer's bunchsize_/num_thr, TNetLib/Platform.h:159) in its own thread; the gradients are summed and
//     the update applied by the peer-memory kernel (csrc/peer.cu) with N = the whole bunch, so the weights equal the single-GPU run's.
#include <sys/time.h>
#include <unistd.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "cu_nnet.h"
#include "io.h"
#include "main_common.h"

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
          "SOURCETRANSCDIR SOURCETRANSCEXT TARGETMMF TARGETMODELDIR TARGETMODELEXT TRACE WEIGHTCOST MATH FUSE GPUS LOADER\n"
          "\n"
          "STARTFRMEXT ENDFRMEXT CMEANDIR CMEANMASK VARSCALEDIR VARSCALEMASK VARSCALEFN TARGETKIND DERIVWINDOWS DELTAWINDOW ACCWINDOW "
          "THIRDWINDOW TEMPBASISFOLDER\n\n",
          progname, kVersion, progname);
  exit(-1);
}

// blocking queue of small messages between the loader / trainer / worker threads
template <typename T>
class MsgQueue {
 public:
  void Push(const T &v) { { std::lock_guard<std::mutex> lk(mu_); q_.push_back(v); } cv_.notify_one(); }
  T Pop() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !q_.empty(); });
    T v = q_.front();
    q_.pop_front();
    return v;
  }
 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<T> q_;
};

// what a worker rank (--GPUS=N, ranks 1..N-1) shares with rank 0
struct WorkerShared {
  int rank = 0, gpu = 0;
  // two input slots on the worker's GPU, filled by rank 0 with peer copies
  float *feats[2] = {NULL, NULL};
  float *labels[2] = {NULL, NULL};
  size_t feats_stride = 0, labels_stride = 0;
  void *ready[2] = {NULL, NULL};   // recorded by rank 0 behind its copies
  void *used[2] = {NULL, NULL};    // recorded by the worker behind the step that read the slot
  MsgQueue<int> inbox;             // slot index to train on; -1 = stop
  MsgQueue<int> up;                // 1 once the worker is set up (or -1 on failure), then 1 when it has finished
  double error = 0.0;
  long long frames = 0, correct = 0;
  std::string failure;
};

// Threads are heap-allocated and deleted only after a join: an exception that unwinds main() past a running thread must reach the
// handler below (which prints it) instead of std::terminate in a std::thread destructor; the handler then leaves with _exit.
static bool g_threads_started = false;

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
  FeatureParams fp = GetFeatureParams(ui, SNAME);   // STARTFRMEXT .. DERIVWINDOWS (UserInterface.cc:361-462)
  const int start_frm_ext = fp.start_frm_ext, end_frm_ext = fp.end_frm_ext;

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
  int n_gpus = ui.GetInt(SNAME ":GPUS", 1);
  bool use_loader = ui.GetBool(SNAME ":LOADER", true);

  if (ui.GetBool(SNAME ":PRINTCONFIG", false)) { std::cout << std::endl; ui.PrintConfig(std::cout); std::cout << std::endl; }
  if (ui.GetBool(SNAME ":PRINTVERSION", false)) std::cout << std::endl << "======= TNET v" << kVersion << " =======" << std::endl << std::endl;
  ui.CheckCommandLineParamUse();

  for (; args_parsed < argc; args_parsed++) feature_repo.AddFile(argv[args_parsed]);

  // ---- networks ----
  CuNetwork network;
  network.SetFusion(fuse);
  if (NULL != p_source_mmf_file) {
    if (trace & 1) TraceLog(std::string("Reading network: ") + p_source_mmf_file);
    network.ReadNetwork(p_source_mmf_file);
  } else {
    Error("Source MMF must be specified [-H]");
  }

  InitFeatureRepository(feature_repo, fp);
  feature_repo.Trace(trace);
  if (NULL != p_script) feature_repo.AddFileList(p_script);
  else Warning("WARNING: The script file is missing [-S]");

  if (mlf_transc) {
    if (NULL == p_source_mlf_file) Error("Source mlf file file is missing [-I]");
    if (NULL == p_output_label_map) Error("Output label map is missing [-m]");
    if (trace & 1) TraceLog(std::string("Indexing labels: ") + p_source_mlf_file);
    label_repo.Init(p_source_mlf_file, p_output_label_map, p_src_lbl_dir, p_src_lbl_ext);
    label_repo.Trace(trace);
  }

  CuObjectiveFunction *p_obj_function = CuObjectiveFunction::Factory(obj_fun_id);
  CuCrossEntropy *p_xent = dynamic_cast<CuCrossEntropy *>(p_obj_function);
  // cross-entropy on MLF labels takes class ids; MSE, and targets read as matrices (MLFTRANSC=FALSE), need the dense rows
  const bool id_targets = p_xent != NULL && mlf_transc;
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

  // ---- data parallel over the GPUs of this box (--GPUS=N): rank 0 is this thread, ranks 1..N-1 are worker threads ----
  if (n_gpus < 1) Error("GPUS must be at least 1");
  if (n_gpus > 1) {
    if (!id_targets) Error("GPUS > 1 is built for the cross-entropy objective (targets travel as class ids)");
    if (bunch_size % n_gpus != 0) Error("BUNCHSIZE must be a multiple of GPUS (GPU g trains rows [g*B/N, (g+1)*B/N) of every bunch)");
    int have = 0;
    TNB_CHECK(tnb_device_count(&have));
    if (have < n_gpus) { std::ostringstream os; os << "GPUS=" << n_gpus << " but only " << have << " GPUs are visible"; Error(os.str()); }
    std::cout << "Data parallel: " << n_gpus << " GPUs, " << bunch_size / n_gpus << " frames of every bunch each\n";
  }
  CuDevice &main_dev = CuDevice::Instantiate();
  const int gpu0 = main_dev.Device();
  const int rows_per_rank = bunch_size / n_gpus;
  TnbLocalGroup *group = NULL;
  std::vector<WorkerShared *> workers;
  std::vector<std::thread *> worker_threads;
  if (n_gpus > 1) {
    TNB_CHECK(tnb_local_group_create(&group, n_gpus));
    for (int r = 1; r < n_gpus; r++) {
      WorkerShared *w = new WorkerShared();
      w->rank = r;
      w->gpu = (gpu0 + r) % n_gpus;
      workers.push_back(w);
      g_threads_started = true;
      worker_threads.push_back(new std::thread([&, w]() {
        try {
          CuDevice::Instantiate().InheritFrom(main_dev, w->gpu);
          TNB_CHECK(tnb_comm_init_local(Cx(), group, w->rank));
          CuNetwork net;
          net.SetFusion(fuse);
          net.ReadNetwork(p_source_mmf_file);
          net.SetLearnRate(learning_rate, learning_rate_factors);
          net.SetMomentum(momentum);
          net.SetWeightcost(weightcost);
          net.SetL1(l1);
          net.SetGradDivFrm(grad_div_frm);
          net.SetDataParallel(n_gpus);  // collective with the other ranks
          CuCrossEntropy xent;
          CuMatrix<BaseFloat> f[2], l[2], err;
          for (int sl = 0; sl < 2; sl++) {
            f[sl].Init(rows_per_rank, net.GetNInputs());
            l[sl].Init(rows_per_rank, 1);
            w->feats[sl] = f[sl].pCUData();
            w->labels[sl] = l[sl].pCUData();
            TNB_CHECK(tnb_event_create(Cx(), &w->used[sl]));
            TNB_CHECK(tnb_event_record(Cx(), w->used[sl], TNB_STREAM_COMPUTE));
          }
          w->feats_stride = f[0].Stride();
          w->labels_stride = l[0].Stride();
          CuDevice::Instantiate().Sync();
          w->up.Push(1);
          for (;;) {
            const int sl = w->inbox.Pop();
            if (sl < 0) break;
            TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_COMPUTE, w->ready[sl]));  // rank 0's peer copies into the slot
            net.PropagateEvaluateLabels(f[sl], l[sl], xent, err);
            if (!cross_validate) net.Backpropagate(err);
            TNB_CHECK(tnb_event_record(Cx(), w->used[sl], TNB_STREAM_COMPUTE));
          }
          net.WaitDataParallel();
          CuDevice::Instantiate().Sync();
          w->error = xent.GetError(); w->frames = (long long)xent.GetFrames(); w->correct = (long long)xent.GetCorrect();
          w->up.Push(1);
        } catch (std::exception &e) {
          w->failure = e.what();
          w->up.Push(-1);
        }
      }));
    }
    TNB_CHECK(tnb_comm_init_local(Cx(), group, 0));
    network.SetDataParallel(n_gpus);
    for (size_t k = 0; k < workers.size(); k++) {
      if (workers[k]->up.Pop() < 0) Error(std::string("worker rank failed to start: ") + workers[k]->failure);
      for (int sl = 0; sl < 2; sl++) TNB_CHECK(tnb_event_create(Cx(), &workers[k]->ready[sl]));
    }
  }

  // ---- the front end: fills cache[k] while the trainer exhausts cache[1 - k] ----
  CuCache cache[2], carry;
  for (int k = 0; k < 2; k++) { cache[k].Init(cache_size, bunch_size); cache[k].Trace(trace); }
  carry.Init(cache_size, bunch_size);
  feature_repo.Rewind();
  MsgQueue<int> filled, freed;     // cache indices: loader -> trainer (-1 = end of data, -2 = failure), trainer -> loader
  std::string loader_failure;
  auto fill_one = [&](CuNetwork &transform_network, CuCache &c, CuMatrix<BaseFloat> &feats_original, CuMatrix<BaseFloat> &feats_expanded,
                      CuMatrix<BaseFloat> &feats_trim, CuMatrix<BaseFloat> &labs_cu, CuVector<int> &label_ids) {
    timer_frontend.Start();
    carry.MoveLeftoverTo(c);  // rows of the previous fill's last utterance that did not fit
    while (!c.Full() && !feature_repo.EndOfList()) {
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
      if (!mlf_transc) {
        // targets are an HTK matrix file next to (or named after) the features (TNetCu.cc:402-413)
        std::vector<char> lbl_file(strlen(feature_repo.Current().Logical().c_str()) + (p_src_lbl_dir ? strlen(p_src_lbl_dir) : 0) +
                                   (p_src_lbl_ext ? strlen(p_src_lbl_ext) : 0) + 8);
        MakeHtkFileName(&lbl_file[0], feature_repo.Current().Logical().c_str(), p_src_lbl_dir, p_src_lbl_ext);
        Matrix<BaseFloat> labs_host;
        LoadHtkMatrix(&lbl_file[0], labs_host);
        t_part.End(); time_labels += t_part.Val(); t_part.Start();
        labs_cu.CopyFrom(labs_host);
        if ((int)labs_cu.Rows() != rows) Error(std::string("Nonmatching number number of input/target examples") + feature_repo.Current().Logical());
        c.AddData(feats_trim, labs_cu);
        t_part.End(); time_cache += t_part.Val();
        feature_repo.MoveNext();
        continue;
      }
      // labels: class ids go to the device (4 bytes per frame instead of 4*nOutputs) ...
      std::vector<int> ids;
      label_repo.GenLabelIds(ids, rows, feature_repo.CurrentHeader().mSamplePeriod, feature_repo.Current().Logical().c_str());
      Vector<int> ids_host(rows);
      for (int i = 0; i < rows; i++) ids_host[i] = ids[i];
      t_part.End(); time_labels += t_part.Val(); t_part.Start();
      label_ids.CopyFrom(ids_host);
      if ((size_t)label_repo.NOutputs() != network.GetNOutputs() && obj_fun_id == CuObjectiveFunction::CROSS_ENTROPY) {
        std::ostringstream os;
        os << "Non-matching dimensions of network output with training targets!!!" << " Netoutput:" << network.GetNOutputs()
           << " Targets:" << label_repo.NOutputs();
        Error(os.str());
      }
      if (id_targets) {
        // ... and stay ids: the cache holds them as a [rows x 1] column, the objective kernel takes them as they are
        CuCrossEntropy::LabelsFromIds(label_ids, 0, rows, labs_cu);
      } else {
        labs_cu.Init(rows, label_repo.NOutputs());  // MSE: dense one-hot rows, expanded on the device
        TNB_CHECK(tnb_onehot(Cx(), labs_cu.pCUData(), label_ids.pCUData(), labs_cu.Dim()));
      }
      c.AddData(feats_trim, labs_cu);
      t_part.End(); time_cache += t_part.Val();
      feature_repo.MoveNext();
    }
    if (randomize) c.Randomize();
    c.MoveLeftoverTo(carry);
    CuDevice::Instantiate().Sync();  // everything the filling thread enqueued has landed before the cache changes hands
    timer_frontend.End();
    time_frontend += timer_frontend.Val();
  };
  auto loader_body = [&]() {
    try {
      CuDevice::Instantiate().InheritFrom(main_dev);
      CuNetwork transform_network;
      if (NULL != p_input_transform) {
        if (trace & 1) TraceLog(std::string("Reading input transform network: ") + p_input_transform);
        transform_network.ReadNetwork(p_input_transform);
      }
      CuMatrix<BaseFloat> feats_original, feats_expanded, feats_trim, labs_cu;
      CuVector<int> label_ids;
      int k = 0, out = 0;
      while (!feature_repo.EndOfList()) {
        if (out >= 2) (void)freed.Pop();  // both caches are with the trainer: wait for one to come back
        fill_one(transform_network, cache[k], feats_original, feats_expanded, feats_trim, labs_cu, label_ids);
        filled.Push(k);
        out++;
        k ^= 1;
      }
      filled.Push(-1);
    } catch (std::exception &e) {
      loader_failure = e.what();
      filled.Push(-2);
    }
  };
  std::thread *loader_thread = NULL;
  CuNetwork serial_transform;      // --LOADER=FALSE: read and train in turns on this thread, as the reference does
  CuMatrix<BaseFloat> s_orig, s_exp, s_trim, s_labs;
  CuVector<int> s_ids;
  int serial_k = 0;
  if (use_loader) {
    g_threads_started = true;
    loader_thread = new std::thread(loader_body);
  } else if (NULL != p_input_transform) {
    if (trace & 1) TraceLog(std::string("Reading input transform network: ") + p_input_transform);
    serial_transform.ReadNetwork(p_input_transform);
  }

  CuMatrix<BaseFloat> feats, labs, globerr, feats0, labs0;
  unsigned long bunch_no = 0;
  for (;;) {
    int k;
    if (use_loader) {
      k = filled.Pop();
      if (k == -2) { loader_thread->join(); delete loader_thread; loader_thread = NULL; Error(loader_failure); }
      if (k < 0) break;
    } else {
      if (feature_repo.EndOfList()) break;
      k = serial_k;
      serial_k ^= 1;
      fill_one(serial_transform, cache[k], s_orig, s_exp, s_trim, s_labs, s_ids);
    }
    CuCache &c = cache[k];
    while (!c.Empty()) {
      c.GetBunch(feats, labs);
      if (n_gpus == 1) {
        if (id_targets) network.PropagateEvaluateLabels(feats, labs, *p_xent, globerr);
        else network.PropagateEvaluate(feats, labs, *p_obj_function, globerr);
        if (!cross_validate) network.Backpropagate(globerr);
      } else {
        const int sl = (int)(bunch_no & 1);
        for (size_t q = 0; q < workers.size(); q++) {  // rows [r*B/N, (r+1)*B/N) of the bunch to rank r's slot, over NVLink
          WorkerShared *w = workers[q];
          const size_t r0 = (size_t)w->rank * rows_per_rank;
          TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_COMPUTE, w->used[sl]));  // the step that last read the slot is done
          TNB_CHECK(tnb_memcpy2d(Cx(), w->feats[sl], w->feats_stride * sizeof(BaseFloat), feats.pCURowData(r0), feats.Stride() * sizeof(BaseFloat),
                                 feats.Cols() * sizeof(BaseFloat), rows_per_rank, 2));
          TNB_CHECK(tnb_memcpy2d(Cx(), w->labels[sl], w->labels_stride * sizeof(BaseFloat), labs.pCURowData(r0), labs.Stride() * sizeof(BaseFloat),
                                 sizeof(BaseFloat), rows_per_rank, 2));
          TNB_CHECK(tnb_event_record(Cx(), w->ready[sl], TNB_STREAM_COMPUTE));
          w->inbox.Push(sl);
        }
        feats0.Init(rows_per_rank, feats.Cols());
        feats0.CopyRows(rows_per_rank, 0, feats, 0);
        labs0.Init(rows_per_rank, 1);
        labs0.CopyRows(rows_per_rank, 0, labs, 0);
        network.PropagateEvaluateLabels(feats0, labs0, *p_xent, globerr);
        if (!cross_validate) network.Backpropagate(globerr);
      }
      bunch_no++;
      if (trace & 2) std::cout << "." << std::flush;
    }
    if (use_loader) {
      CuDevice::Instantiate().Sync();  // the trainer's reads of this cache are done before the loader refills it
      freed.Push(k);
    }
  }
  if (use_loader && loader_thread) { loader_thread->join(); delete loader_thread; loader_thread = NULL; }
  if (n_gpus > 1) {
    network.WaitDataParallel();
    for (size_t q = 0; q < workers.size(); q++) workers[q]->inbox.Push(-1);
    CuDevice::Instantiate().Sync();
    for (size_t q = 0; q < workers.size(); q++) {
      const int rc = workers[q]->up.Pop();
      worker_threads[q]->join();
      delete worker_threads[q];
      if (rc < 0) Error(std::string("worker rank failed: ") + workers[q]->failure);
      p_obj_function->AddStats(workers[q]->error, workers[q]->frames, workers[q]->correct);  // cf. MergeStats, TNetLib/ObjFun.cc:214-228
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
  for (size_t q = 0; q < workers.size(); q++) delete workers[q];
  if (group) { tnb_comm_destroy(Cx()); tnb_local_group_destroy(group); }
  return 0;
} catch (std::exception &rExc) {
  std::cerr << "Exception thrown" << std::endl;
  std::cerr << rExc.what() << std::endl;
  if (g_threads_started) { std::cout.flush(); std::cerr.flush(); _exit(1); }  // other threads may still be running: no static destructors
  return 1;
}
