/*
 * tnet_b200_host.h — C handles over the C++ mirror of the reference's CuTNetLib interface
 * (nnet-asr_b200/host: CuNetwork, CuCache, CuObjectiveFunction, CuRbm + CuRand, CuRecurrent), exported by
 * libtnetb200_host.so.  This is the surface tests/ and bench.py drive (through ctypes); the drop-in binaries
 * TNetCu / TRbmCu / TRecurrentCu use the C++ classes directly.  Everything below bottoms out in the kernel ABI of
 * include/tnet_b200.h — there is no CPU path.
 *
 * Each call returns 0 on success; on failure the C++ exception text is available from tnh_last_error().
 * "host" pointers are dense row-major host arrays (leading dimension = number of columns).
 */
#ifndef TNET_B200_HOST_H_
#define TNET_B200_HOST_H_

#include "tnet_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TnhNet_ TnhNet;       /* CuNetwork + CuObjectiveFunction + bunch buffers (one TNetCu main loop) */
typedef struct TnhCache_ TnhCache;   /* CuCache */
typedef struct TnhRbm_ TnhRbm;       /* single <rbm> network + CuRand + MSE (TRbmCu main loop) */
typedef struct TnhRnn_ TnhRnn;       /* network with <recurrent> layers + Xent (TRecurrentCu main loop) */

const char *tnh_last_error(void);
int tnh_select_gpu(int device);                 /* CuDevice::SelectGPU — before the first device op (TNetCu.cc:242-243) */
int tnh_set_math(int tnb_math_mode);            /* TNB_MATH_* for every GEMM issued by this process */
int tnh_ctx(TnbContext **ctx);                  /* the process-wide kernel-ABI context (for tnb_comm_init etc.) */
int tnh_sync(void);
int tnh_launch_count(unsigned long long *n);
void tnh_srand48(long seed);                    /* srand48(seed) as TNetCu.cc:330-338 */

/* ---- MLP trainer: TNetCu.cc:275-287,321-328,427-441 ------------------------------------------------- */
int tnh_net_read(TnhNet **net, const char *network_file, int objective /*0 xent, 1 mse*/);
/* tools/init/gen_mlp_init.py --gauss --negbias equivalent without the text file: dims[0..n) = layer widths,
 * W ~ 0.1*N(0,1), hidden bias ~ U[-4.1,-3.9], output bias 0, sigmoid hidden layers, softmax output; mt19937(seed). */
int tnh_net_new_mlp(TnhNet **net, const int *dims, int n_dims, unsigned seed, int objective);
int tnh_net_free(TnhNet *net);
int tnh_net_write(TnhNet *net, const char *network_file);
int tnh_net_set_hyper(TnhNet *net, float learn_rate, const char *learn_rate_factors /*NULL or "a:b:c"*/, float momentum,
                      float weightcost, int grad_div_frm);
int tnh_net_set_fusion(TnhNet *net, int on);
/* 0: every GEMM of the backward pass in its own launch (the round-1 schedule); 1 (default): independent ones share persistent
 * launches (tnb_gemm_batch) */
int tnh_net_set_batching(TnhNet *net, int on);
/* parameters of a <biasedlinearity> layer as held on the device, without the text round trip: W [nin x nout] row-major (the file
 * stores its transpose, cuBiasedLinearity.cc:70-78) and bias [nout]; either pointer may be NULL.  *nin / *nout return the dimensions. */
int tnh_net_get_affine(TnhNet *net, int layer, float *W_host, float *bias_host, int *nin, int *nout);
int tnh_net_set_data_parallel(TnhNet *net, int world);
int tnh_net_dims(TnhNet *net, int *n_inputs, int *n_outputs, int *n_layers);
int tnh_net_propagate(TnhNet *net, const float *x_host, int rows, float *out_host);
/* one bunch with dense targets: Propagate, Evaluate, Backpropagate (skipped when cross_validate) */
int tnh_net_train_bunch(TnhNet *net, const float *x_host, const float *t_host, int rows, int cross_validate);
/* the same with class ids; x_host/labels_host should be pinned (tnb_host_alloc) for an asynchronous copy */
int tnh_net_train_bunch_labels(TnhNet *net, const float *x_host, const int *labels_host, int rows, int cross_validate);
int tnh_net_stats(TnhNet *net, double *error, long long *frames, long long *correct);
/* Pipelined form of the pair above, for a host that feeds bunches from PINNED memory (what TNetCu's reader thread would do):
 * submit() enqueues the H2D copy of this bunch on the copy stream into one of four device buffers (at most four submissions in flight), the training step behind it on
 * the compute stream, and an asynchronous D2H copy of the running objective statistics; it returns without waiting.
 * collect() blocks until the OLDEST uncollected submission has finished and returns the statistics as of that bunch.  With one
 * submission kept in flight (submit k+1, then collect k) the copy of bunch k+1 overlaps the step of bunch k; keeping more in flight
 * absorbs host-side jitter (bench.py keeps three with several ranks, where a late rank stalls all of them).  x_host/labels_host
 * must stay unchanged until the submission has been collected. */
int tnh_net_submit_bunch_labels(TnhNet *net, const float *x_host_pinned, const int *labels_host_pinned, int rows, int cross_validate);
int tnh_net_collect(TnhNet *net, double *error, long long *frames, long long *correct);
int tnh_net_add_stats(TnhNet *net, double error, long long frames, long long correct);
int tnh_net_layer_output(TnhNet *net, int layer, float *out_host, int rows, int cols);
int tnh_net_layer_error_output(TnhNet *net, int layer, float *out_host, int rows, int cols);
int tnh_net_global_error(TnhNet *net, float *out_host, int rows, int cols);
/* device-resident training set (the synthetic-data bench): frames stay in HBM, bunches are row windows */
int tnh_net_load_resident(TnhNet *net, const float *x_host, const int *labels_host, int rows);
int tnh_net_train_resident(TnhNet *net, int bunch, int first_bunch, int n_bunches, int cross_validate);

/* ---- CuCache: cuCache.h:13-48 ------------------------------------------------------------------------ */
int tnh_cache_new(TnhCache **cache, int cachesize, int bunchsize);
int tnh_cache_free(TnhCache *cache);
int tnh_cache_add(TnhCache *cache, const float *f_host, const float *d_host, int rows, int fdim, int ddim);
int tnh_cache_full(TnhCache *cache);
int tnh_cache_empty(TnhCache *cache);
int tnh_cache_discarded(TnhCache *cache);
int tnh_cache_randomize(TnhCache *cache, int *perm_out /*NULL or >= cachesize ints*/, int *perm_len);
int tnh_cache_get_bunch(TnhCache *cache, float *f_host, float *d_host);
/* train straight from the cache without the host round trip (what TNetCu does) */
int tnh_net_train_from_cache(TnhNet *net, TnhCache *cache, int cross_validate, int *n_bunches);

/* ---- RBM CD-1: TRbmCu.cc:226-264,326-356 -------------------------------------------------------------- */
int tnh_rbm_read(TnhRbm **rbm, const char *network_file, int bunchsize, float learn_rate, float momentum, float weightcost);
int tnh_rbm_free(TnhRbm *rbm);
int tnh_rbm_write(TnhRbm *rbm, const char *network_file);
int tnh_rbm_dims(TnhRbm *rbm, int *n_vis, int *n_hid);
int tnh_rbm_cd1_bunch(TnhRbm *rbm, const float *pos_vis_host, int rows);
int tnh_rbm_cd1_from_cache(TnhRbm *rbm, TnhCache *cache, int *n_bunches);
int tnh_rbm_stats(TnhRbm *rbm, double *error, long long *frames);
int tnh_rbm_last(TnhRbm *rbm, float *pos_hid_host, float *neg_hid_host, float *neg_vis_host);

/* ---- recurrent: TRecurrentCu.cc:319-375 ---------------------------------------------------------------- */
int tnh_rnn_read(TnhRnn **rnn, const char *network_file, int bptt, float learn_rate, float momentum, float weightcost);
int tnh_rnn_free(TnhRnn *rnn);
int tnh_rnn_write(TnhRnn *rnn, const char *network_file);
int tnh_rnn_train_utterance(TnhRnn *rnn, const float *x_host, const int *labels_host, int rows, int cross_validate);
int tnh_rnn_stats(TnhRnn *rnn, double *error, long long *frames, long long *correct);

#ifdef __cplusplus
}
#endif
#endif
