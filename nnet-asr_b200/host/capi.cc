// capi.cc — extern "C" handles over the C++ classes of cu_nnet.h (declared in include/tnet_b200_host.h).
// Each handle bundles exactly the objects one reference main() keeps alive:
//   TnhNet  : CuNetwork + CuObjectiveFunction + the feats/labs/globerr matrices of TNetCu.cc:370-441
//   TnhRbm  : CuNetwork(<rbm>) + CuRand + CuMeanSquareError + the four CD-1 matrices of TRbmCu.cc:296-356
//   TnhRnn  : CuNetwork with <recurrent> + Xent + the 1-row matrices of TRecurrentCu.cc:358-372
#include "tnet_b200_host.h"

#include <random>

#include "cu_nnet.h"

using namespace TNet;

static thread_local std::string g_host_err;

#define TNH_TRY try {
#define TNH_CATCH                                   \
  }                                                 \
  catch (std::exception & e) {                      \
    g_host_err = e.what();                          \
    return 1;                                       \
  }                                                 \
  catch (...) {                                     \
    g_host_err = "unknown C++ exception";           \
    return 1;                                       \
  }                                                 \
  return 0;

struct TnhNet_ {
  CuNetwork net;
  CuObjectiveFunction *obj;
  CuMatrix<BaseFloat> feats, labs, output, globerr;
  CuVector<int> labels;
  // resident training set
  CuMatrix<BaseFloat> res_feats;
  CuVector<int> res_labels;
  // pipelined submission (tnh_net_submit_bunch_labels / tnh_net_collect): NSLOT input slots (as many submissions may be in flight)
  enum { NSLOT = 4 };
  struct Slot {
    CuMatrix<BaseFloat> feats;
    CuMatrix<BaseFloat> stage;  // the bunch as it lies in host memory ([rows x nin] without a pitch, stored as ONE row): target of the H2D copy
    CuVector<int> labels;
    void *ready, *used, *done;  // H2D landed (copy stream) / step no longer reads the slot / statistics copied back
    TnbObjStats *stats_host;    // pinned
    Slot() : ready(NULL), used(NULL), done(NULL), stats_host(NULL) {}
  } slot[NSLOT];
  unsigned long long submitted, collected;
  TnhNet_() : obj(NULL), submitted(0), collected(0) {}
  ~TnhNet_() {
    delete obj;
    for (int i = 0; i < NSLOT; i++) {
      if (slot[i].ready) { tnb_event_destroy(Cx(), slot[i].ready); tnb_event_destroy(Cx(), slot[i].used); tnb_event_destroy(Cx(), slot[i].done); }
      if (slot[i].stats_host) tnb_host_free(slot[i].stats_host);
    }
  }
  void Step(bool cv) {  // TNetCu.cc:431-438 with the softmax fused into the objective
    net.PropagateEvaluate(feats, labs, *obj, globerr);
    if (!cv) net.Backpropagate(globerr);
  }
  /// the same step with the targets as class ids on the device (`stride` ints apart): no one-hot matrix exists anywhere.
  /// Objectives other than cross-entropy need the dense targets and get them expanded here.
  void StepIds(const CuMatrix<BaseFloat> &x, const int *ids, int stride, bool cv) {
    CuCrossEntropy *xent = dynamic_cast<CuCrossEntropy *>(obj);
    if (xent) {
      net.PropagateEvaluateIds(x, ids, stride, *xent, globerr);
    } else {
      labs.Init(x.Rows(), net.GetNOutputs());
      TNB_CHECK(tnb_onehot_strided(Cx(), labs.pCUData(), ids, stride, labs.Dim()));
      net.PropagateEvaluate(x, labs, *obj, globerr);
    }
    if (!cv) net.Backpropagate(globerr);
  }
};

struct TnhCache_ {
  CuCache cache;
  CuMatrix<BaseFloat> f, d;
  int fdim, ddim, bunch;
  TnhCache_() : fdim(0), ddim(0), bunch(0) {}
};

struct TnhRbm_ {
  CuNetwork net;
  CuRbmBase *rbm;
  CuRand<BaseFloat> *rnd;
  CuMeanSquareError mse;
  CuMatrix<BaseFloat> pos_vis, pos_hid, neg_vis, neg_hid, dummy_labs, dummy_err;
  TnhRbm_() : rbm(NULL), rnd(NULL) {}
  ~TnhRbm_() { delete rnd; }
  void Step() {  // TRbmCu.cc:326-354
    rbm->Propagate(pos_vis, pos_hid);
    if (rbm->HidType() == CuRbmBase::BERNOULLI) {
      rnd->BinarizeProbs(pos_hid, neg_hid);
    } else {
      neg_hid.CopyFrom(pos_hid);
      rnd->AddGaussNoise(neg_hid);
    }
    rbm->Reconstruct(neg_hid, neg_vis);
    rbm->Propagate(neg_vis, neg_hid);
    rbm->RbmUpdate(pos_vis, pos_hid, neg_vis, neg_hid);
    mse.Evaluate(neg_vis, pos_vis, dummy_err);
  }
};

struct TnhRnn_ {
  CuNetwork net;
  CuCrossEntropy xent;
  CuMatrix<BaseFloat> feats, targets, input_row, output_row, target_row, error_row;
  CuVector<int> labels;
};

static void upload(CuMatrix<BaseFloat> &dst, const float *host, int rows, int cols) {
  dst.Init(rows, cols);
  if (rows == 0 || cols == 0) return;
  TNB_CHECK(tnb_memcpy2d(Cx(), dst.pCUData(), dst.Stride() * sizeof(float), host, (size_t)cols * sizeof(float), (size_t)cols * sizeof(float),
                         rows, 0));
}
static void download(const CuMatrix<BaseFloat> &src, float *host, int rows, int cols) {
  if ((int)src.Rows() < rows || (int)src.Cols() != cols) Error("download: dimension mismatch");
  if (rows == 0 || cols == 0) return;
  TNB_CHECK(tnb_memcpy2d(Cx(), host, (size_t)cols * sizeof(float), src.pCUData(), src.Stride() * sizeof(float), (size_t)cols * sizeof(float),
                         rows, 1));
}

extern "C" {

const char *tnh_last_error(void) { return g_host_err.c_str(); }
int tnh_select_gpu(int device) { TNH_TRY CuDevice::Instantiate().SelectGPU(device); TNH_CATCH }
int tnh_set_math(int mode) { TNH_TRY CuDevice::Instantiate().SetMath(mode); TNH_CATCH }
int tnh_ctx(TnbContext **ctx) { TNH_TRY *ctx = Cx(); TNH_CATCH }
int tnh_sync(void) { TNH_TRY CuDevice::Instantiate().Sync(); TNH_CATCH }
int tnh_launch_count(unsigned long long *n) { TNH_TRY *n = CuDevice::Instantiate().Launches(); TNH_CATCH }
void tnh_srand48(long seed) { srand48(seed); }

// ------------------------------------------------------------------------------------------- MLP
int tnh_net_read(TnhNet **out, const char *file, int objective) {
  TNH_TRY
  *out = NULL;
  TnhNet *h = new TnhNet_();
  try {
    h->net.ReadNetwork(file);
    h->obj = CuObjectiveFunction::Factory(objective == 1 ? CuObjectiveFunction::MEAN_SQUARE_ERROR : CuObjectiveFunction::CROSS_ENTROPY);
  } catch (...) { delete h; throw; }
  *out = h;
  TNH_CATCH
}
int tnh_net_new_mlp(TnhNet **out, const int *dims, int n_dims, unsigned seed, int objective) {
  TNH_TRY
  *out = NULL;
  if (n_dims < 2) Error("an MLP needs at least two widths");
  TnhNet *h = new TnhNet_();
  try {
    std::mt19937 gen(seed);
    std::normal_distribution<float> gauss(0.0f, 1.0f);
    std::uniform_real_distribution<float> uni(0.0f, 1.0f);
    for (int l = 0; l + 1 < n_dims; l++) {
      const int nin = dims[l], nout = dims[l + 1];
      const bool last = (l == n_dims - 2);
      BfMatrix Wt(nout, nin);
      BfVector b(nout);
      for (int o = 0; o < nout; o++)
        for (int i = 0; i < nin; i++) Wt(o, i) = 0.1f * gauss(gen);
      for (int o = 0; o < nout; o++) b[o] = last ? 0.0f : uni(gen) / 5.0f - 4.1f;
      CuBiasedLinearity *lin = new CuBiasedLinearity(nin, nout, NULL);
      lin->SetParams(Wt, b);
      h->net.AddLayer(lin);
      if (last) h->net.AddLayer(new CuSoftmax(nout, nout, NULL)); else h->net.AddLayer(new CuSigmoid(nout, nout, NULL));
    }
    h->obj = CuObjectiveFunction::Factory(objective == 1 ? CuObjectiveFunction::MEAN_SQUARE_ERROR : CuObjectiveFunction::CROSS_ENTROPY);
  } catch (...) { delete h; throw; }
  *out = h;
  TNH_CATCH
}
int tnh_net_free(TnhNet *h) { TNH_TRY delete h; TNH_CATCH }
int tnh_net_write(TnhNet *h, const char *file) { TNH_TRY h->net.WriteNetwork(file); TNH_CATCH }
int tnh_net_set_hyper(TnhNet *h, float lr, const char *factors, float mmt, float wc, int gdf) {
  TNH_TRY
  h->net.SetLearnRate(lr, (factors && *factors) ? factors : NULL);
  h->net.SetMomentum(mmt);
  h->net.SetWeightcost(wc);
  h->net.SetGradDivFrm(gdf != 0);
  TNH_CATCH
}
int tnh_net_set_fusion(TnhNet *h, int on) { TNH_TRY h->net.SetFusion(on != 0); TNH_CATCH }
int tnh_net_set_batching(TnhNet *h, int on) { TNH_TRY h->net.SetBatching(on != 0); TNH_CATCH }
int tnh_net_get_affine(TnhNet *h, int layer, float *W_host, float *bias_host, int *nin, int *nout) {
  TNH_TRY
  if (layer < 0 || layer >= h->net.Layers()) Error("layer index");
  if (h->net.Layer(layer).GetType() != CuComponent::BIASED_LINEARITY) Error("layer is not a <biasedlinearity>");
  CuBiasedLinearity &lin = static_cast<CuBiasedLinearity &>(h->net.Layer(layer));
  if (nin) *nin = (int)lin.GetNInputs();
  if (nout) *nout = (int)lin.GetNOutputs();
  if (W_host) download(lin.Linearity(), W_host, (int)lin.GetNInputs(), (int)lin.GetNOutputs());
  if (bias_host) TNB_CHECK(tnb_memcpy(Cx(), bias_host, lin.Bias().pCUData(), sizeof(float) * lin.GetNOutputs(), 1));
  TNH_CATCH
}
int tnh_net_set_data_parallel(TnhNet *h, int world) { TNH_TRY h->net.SetDataParallel(world); TNH_CATCH }
int tnh_net_dims(TnhNet *h, int *nin, int *nout, int *nl) {
  TNH_TRY
  *nin = (int)h->net.GetNInputs(); *nout = (int)h->net.GetNOutputs(); *nl = h->net.Layers();
  TNH_CATCH
}
int tnh_net_propagate(TnhNet *h, const float *x, int rows, float *out) {
  TNH_TRY
  upload(h->feats, x, rows, (int)h->net.GetNInputs());
  h->net.Propagate(h->feats, h->output);
  download(h->output, out, rows, (int)h->net.GetNOutputs());
  TNH_CATCH
}
int tnh_net_train_bunch(TnhNet *h, const float *x, const float *t, int rows, int cv) {
  TNH_TRY
  upload(h->feats, x, rows, (int)h->net.GetNInputs());
  upload(h->labs, t, rows, (int)h->net.GetNOutputs());
  h->Step(cv != 0);
  CuDevice::Instantiate().Sync();  // x/t may be pageable host memory
  TNH_CATCH
}
int tnh_net_train_bunch_labels(TnhNet *h, const float *x, const int *lab, int rows, int cv) {
  TNH_TRY
  upload(h->feats, x, rows, (int)h->net.GetNInputs());
  h->labels.Init(rows);
  TNB_CHECK(tnb_memcpy(Cx(), h->labels.pCUData(), lab, sizeof(int) * (size_t)rows, 0));
  h->StepIds(h->feats, h->labels.pCUData(), 1, cv != 0);
  TNH_CATCH
}
int tnh_net_submit_bunch_labels(TnhNet *h, const float *x, const int *lab, int rows, int cv) {
  TNH_TRY
  if (h->submitted - h->collected >= TnhNet_::NSLOT) Error("four submissions are already in flight: collect one first");
  TnhNet_::Slot &s = h->slot[h->submitted % TnhNet_::NSLOT];
  const int nin = (int)h->net.GetNInputs();
  if (!s.ready) {
    TNB_CHECK(tnb_event_create(Cx(), &s.ready));
    TNB_CHECK(tnb_event_create(Cx(), &s.used));
    TNB_CHECK(tnb_event_create(Cx(), &s.done));
    void *p = NULL;
    TNB_CHECK(tnb_host_alloc(&p, sizeof(TnbObjStats)));
    s.stats_host = (TnbObjStats *)p;
    TNB_CHECK(tnb_event_record(Cx(), s.used, TNB_STREAM_COMPUTE));
  }
  if ((int)s.feats.Rows() != rows || (int)s.feats.Cols() != nin) {
    s.feats.Init(rows, nin);   // (re)allocation and zero-fill run on the compute stream: the copy must not overtake them
    s.stage.Init(1, (size_t)rows * nin);
    s.labels.Init(rows);
    TNB_CHECK(tnb_event_record(Cx(), s.used, TNB_STREAM_COMPUTE));
  }
  TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_COPY, s.used));  // the step that last read this slot (NSLOT submissions ago) is done
  // ONE contiguous H2D copy (a copy engine at PCIe rate), then the pitched layout on the device at the start of the step.  A pitched
  // H2D copy (cudaMemcpy2DAsync, 1716-byte rows into a 1792-byte pitch for 429 inputs) cost 0.14 ms per bunch as soon as the
  // peer-memory kernels of a data-parallel step occupied the SMs the GEMMs leave free (profiles/r02_dp_timeline.md).
  TNB_CHECK(tnb_memcpy_on(Cx(), TNB_STREAM_COPY, s.stage.pCUData(), x, (size_t)rows * nin * sizeof(float), 0));
  TNB_CHECK(tnb_memcpy_on(Cx(), TNB_STREAM_COPY, s.labels.pCUData(), lab, sizeof(int) * (size_t)rows, 0));
  TNB_CHECK(tnb_event_record(Cx(), s.ready, TNB_STREAM_COPY));
  TNB_CHECK(tnb_stream_wait_event(Cx(), TNB_STREAM_COMPUTE, s.ready));
  TNB_CHECK(tnb_memcpy2d_on(Cx(), TNB_STREAM_COMPUTE, s.feats.pCUData(), s.feats.Stride() * sizeof(float), s.stage.pCUData(), (size_t)nin * sizeof(float),
                            (size_t)nin * sizeof(float), rows, 2));
  h->StepIds(s.feats, s.labels.pCUData(), 1, cv != 0);
  TNB_CHECK(tnb_event_record(Cx(), s.used, TNB_STREAM_COMPUTE));
  TNB_CHECK(tnb_memcpy_on(Cx(), TNB_STREAM_COMPUTE, s.stats_host, h->obj->DeviceStats(), sizeof(TnbObjStats), 1));
  TNB_CHECK(tnb_event_record(Cx(), s.done, TNB_STREAM_COMPUTE));
  h->submitted++;
  TNH_CATCH
}
int tnh_net_collect(TnhNet *h, double *e, long long *f, long long *c) {
  TNH_TRY
  if (h->collected >= h->submitted) Error("nothing in flight");
  TnhNet_::Slot &s = h->slot[h->collected % TnhNet_::NSLOT];
  TNB_CHECK(tnb_event_sync(Cx(), s.done));
  *e = s.stats_host->error; *f = s.stats_host->frames; *c = s.stats_host->correct;
  h->collected++;
  TNH_CATCH
}
int tnh_net_stats(TnhNet *h, double *e, long long *f, long long *c) {
  TNH_TRY
  *e = h->obj->GetError(); *f = (long long)h->obj->GetFrames(); *c = (long long)h->obj->GetCorrect();
  TNH_CATCH
}
int tnh_net_add_stats(TnhNet *h, double e, long long f, long long c) { TNH_TRY h->obj->AddStats(e, f, c); TNH_CATCH }
int tnh_net_layer_output(TnhNet *h, int layer, float *out, int rows, int cols) {
  TNH_TRY
  if (layer < 0 || layer >= h->net.Layers()) Error("layer index");
  download(h->net.Layer(layer).GetOutput(), out, rows, cols);
  TNH_CATCH
}
int tnh_net_layer_error_output(TnhNet *h, int layer, float *out, int rows, int cols) {
  TNH_TRY
  if (layer < 0 || layer >= h->net.Layers()) Error("layer index");
  download(h->net.Layer(layer).GetErrorOutput(), out, rows, cols);
  TNH_CATCH
}
int tnh_net_global_error(TnhNet *h, float *out, int rows, int cols) { TNH_TRY download(h->globerr, out, rows, cols); TNH_CATCH }

int tnh_net_load_resident(TnhNet *h, const float *x, const int *lab, int rows) {
  TNH_TRY
  upload(h->res_feats, x, rows, (int)h->net.GetNInputs());
  h->res_labels.Init(rows);
  TNB_CHECK(tnb_memcpy(Cx(), h->res_labels.pCUData(), lab, sizeof(int) * (size_t)rows, 0));
  CuDevice::Instantiate().Sync();
  TNH_CATCH
}
int tnh_net_train_resident(TnhNet *h, int bunch, int first, int n, int cv) {
  TNH_TRY
  const int total = (int)h->res_feats.Rows() / bunch;
  if (total <= 0) Error("resident set smaller than one bunch");
  h->feats.Init(bunch, h->net.GetNInputs());
  for (int b = 0; b < n; b++) {
    const size_t r0 = (size_t)((first + b) % total) * bunch;
    h->feats.CopyRows(bunch, r0, h->res_feats, 0);  // CuCache::GetBunch's D2D row window; the ids of the window are read in place
    h->StepIds(h->feats, h->res_labels.pCUData() + r0, 1, cv != 0);
  }
  h->net.WaitDataParallel();  // the last bunch's exchange belongs to this call (stream order; the host does not block)
  TNH_CATCH
}

// ------------------------------------------------------------------------------------------- cache
int tnh_cache_new(TnhCache **out, int cachesize, int bunchsize) {
  TNH_TRY
  *out = NULL;
  TnhCache *c = new TnhCache_();
  try { c->cache.Init(cachesize, bunchsize); c->bunch = bunchsize; } catch (...) { delete c; throw; }
  *out = c;
  TNH_CATCH
}
int tnh_cache_free(TnhCache *c) { TNH_TRY delete c; TNH_CATCH }
int tnh_cache_add(TnhCache *c, const float *f, const float *d, int rows, int fdim, int ddim) {
  TNH_TRY
  upload(c->f, f, rows, fdim);
  upload(c->d, d, rows, ddim);
  CuDevice::Instantiate().Sync();
  c->fdim = fdim; c->ddim = ddim;
  c->cache.AddData(c->f, c->d);
  TNH_CATCH
}
int tnh_cache_full(TnhCache *c) { return c->cache.Full() ? 1 : 0; }
int tnh_cache_empty(TnhCache *c) { return c->cache.Empty() ? 1 : 0; }
int tnh_cache_discarded(TnhCache *c) { return c->cache.Discarded(); }
int tnh_cache_randomize(TnhCache *c, int *perm_out, int *perm_len) {
  TNH_TRY
  c->cache.Randomize();
  const Vector<int> &p = c->cache.LastPermutation();
  if (perm_len) *perm_len = (int)p.Dim();
  if (perm_out) memcpy(perm_out, p.pData(), sizeof(int) * p.Dim());
  TNH_CATCH
}
int tnh_cache_get_bunch(TnhCache *c, float *f, float *d) {
  TNH_TRY
  c->cache.GetBunch(c->f, c->d);
  download(c->f, f, c->bunch, c->fdim);
  download(c->d, d, c->bunch, c->ddim);
  TNH_CATCH
}
int tnh_net_train_from_cache(TnhNet *h, TnhCache *c, int cv, int *n_bunches) {
  TNH_TRY
  int n = 0;
  while (!c->cache.Empty()) {
    c->cache.GetBunch(h->feats, h->labs);
    h->Step(cv != 0);
    n++;
  }
  if (n_bunches) *n_bunches = n;
  TNH_CATCH
}

// ------------------------------------------------------------------------------------------- RBM
int tnh_rbm_read(TnhRbm **out, const char *file, int bunchsize, float lr, float mmt, float wc) {
  TNH_TRY
  *out = NULL;
  TnhRbm *h = new TnhRbm_();
  try {
    h->net.ReadNetwork(file);
    if (h->net.Layers() != 1) Error(std::string("Number of layers must be 1") + file);
    if (h->net.Layer(0).GetType() != CuComponent::RBM && h->net.Layer(0).GetType() != CuComponent::RBM_SPARSE)
      Error(std::string("Layer must be RBM") + file);
    h->rbm = dynamic_cast<CuRbmBase *>(&h->net.Layer(0));
    h->rbm->LearnRate(lr); h->rbm->Momentum(mmt); h->rbm->Weightcost(wc);
    // the generator is seeded from lrand48() right here, i.e. after srand48(seed) and before any cache shuffle
    h->rnd = new CuRand<BaseFloat>(bunchsize, h->rbm->GetNOutputs());
  } catch (...) { delete h; throw; }
  *out = h;
  TNH_CATCH
}
int tnh_rbm_free(TnhRbm *h) { TNH_TRY delete h; TNH_CATCH }
int tnh_rbm_write(TnhRbm *h, const char *file) { TNH_TRY h->net.WriteNetwork(file); TNH_CATCH }
int tnh_rbm_dims(TnhRbm *h, int *nvis, int *nhid) { TNH_TRY *nvis = (int)h->rbm->GetNInputs(); *nhid = (int)h->rbm->GetNOutputs(); TNH_CATCH }
int tnh_rbm_cd1_bunch(TnhRbm *h, const float *pos_vis, int rows) {
  TNH_TRY
  upload(h->pos_vis, pos_vis, rows, (int)h->rbm->GetNInputs());
  h->Step();
  CuDevice::Instantiate().Sync();
  TNH_CATCH
}
int tnh_rbm_cd1_from_cache(TnhRbm *h, TnhCache *c, int *n_bunches) {
  TNH_TRY
  int n = 0;
  while (!c->cache.Empty()) {
    c->cache.GetBunch(h->pos_vis, h->dummy_labs);
    h->Step();
    n++;
  }
  if (n > 0) h->pos_hid.CheckData();  // TRbmCu.cc:356
  if (n_bunches) *n_bunches = n;
  TNH_CATCH
}
int tnh_rbm_stats(TnhRbm *h, double *e, long long *f) { TNH_TRY *e = h->mse.GetError(); *f = (long long)h->mse.GetFrames(); TNH_CATCH }
int tnh_rbm_last(TnhRbm *h, float *pos_hid, float *neg_hid, float *neg_vis) {
  TNH_TRY
  if (pos_hid) download(h->pos_hid, pos_hid, (int)h->pos_hid.Rows(), (int)h->pos_hid.Cols());
  if (neg_hid) download(h->neg_hid, neg_hid, (int)h->neg_hid.Rows(), (int)h->neg_hid.Cols());
  if (neg_vis) download(h->neg_vis, neg_vis, (int)h->neg_vis.Rows(), (int)h->neg_vis.Cols());
  TNH_CATCH
}

// ------------------------------------------------------------------------------------------- recurrent
int tnh_rnn_read(TnhRnn **out, const char *file, int bptt, float lr, float mmt, float wc) {
  TNH_TRY
  *out = NULL;
  TnhRnn *h = new TnhRnn_();
  try {
    h->net.ReadNetwork(file);
    h->net.SetLearnRate(lr, NULL);
    h->net.SetMomentum(mmt);
    h->net.SetWeightcost(wc);
    for (int i = 0; i < h->net.Layers(); i++)
      if (h->net.Layer(i).GetType() == CuComponent::RECURRENT) dynamic_cast<CuRecurrent &>(h->net.Layer(i)).BpttOrder(bptt);
  } catch (...) { delete h; throw; }
  *out = h;
  TNH_CATCH
}
int tnh_rnn_free(TnhRnn *h) { TNH_TRY delete h; TNH_CATCH }
int tnh_rnn_write(TnhRnn *h, const char *file) { TNH_TRY h->net.WriteNetwork(file); TNH_CATCH }
int tnh_rnn_train_utterance(TnhRnn *h, const float *x, const int *lab, int rows, int cv) {
  TNH_TRY
  const int nin = (int)h->net.GetNInputs(), nout = (int)h->net.GetNOutputs();
  upload(h->feats, x, rows, nin);
  h->labels.Init(rows);
  TNB_CHECK(tnb_memcpy(Cx(), h->labels.pCUData(), lab, sizeof(int) * (size_t)rows, 0));
  h->targets.Init(rows, nout);
  TNB_CHECK(tnb_onehot(Cx(), h->targets.pCUData(), h->labels.pCUData(), h->targets.Dim()));
  CuDevice::Instantiate().Sync();
  for (int i = 0; i < h->net.Layers(); i++)
    if (h->net.Layer(i).GetType() == CuComponent::RECURRENT) dynamic_cast<CuRecurrent &>(h->net.Layer(i)).ClearHistory();
  h->input_row.Init(1, nin); h->output_row.Init(1, nout); h->target_row.Init(1, nout); h->error_row.Init(1, nout);
  for (int frm = 0; frm < rows; frm++) {  // TRecurrentCu.cc:356-371
    h->input_row.CopyRows(1, frm, h->feats, 0);
    h->target_row.CopyRows(1, frm, h->targets, 0);
    h->net.Propagate(h->input_row, h->output_row);
    h->xent.Evaluate(h->output_row, h->target_row, h->error_row);
    if (!cv) h->net.Backpropagate(h->error_row);
  }
  TNH_CATCH
}
int tnh_rnn_stats(TnhRnn *h, double *e, long long *f, long long *c) {
  TNH_TRY
  *e = h->xent.GetError(); *f = (long long)h->xent.GetFrames(); *c = (long long)h->xent.GetCorrect();
  TNH_CATCH
}

}  // extern "C"
