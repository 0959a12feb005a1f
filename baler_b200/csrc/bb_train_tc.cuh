// Interface between bb_train.cu (C-ABI entry points, fp32 reference kernels) and bb_train_tc.cu (tensor-core trainer).
#pragma once
#include "bb_common.cuh"

struct TcTrainer;  // opaque: packed weight images, transposed activation / gradient scratch, barrier state

// what one launch of the persistent training kernel does per step
enum : int {
  TC_P1 = 1,         // forward + loss + backward of every 16-row tile
  TC_DW = 2,         // weight-gradient tiles
  TC_GRADS = 4,      // ... stored to the flat gradient vector (+ loss slot)
  TC_ADAM = 8,       // ... and applied (Adam) in the tile epilogue, packed weight images refreshed
  TC_FWD_ONLY = 16,  // validation: forward + loss only
  TC_XCHG = 32,      // data parallel: sum the gradient tiles (and the loss) of all ranks over peer memory before Adam
};

struct TcHyper {
  float beta1, beta2, eps;
  double lr, b1d, b2d;  // per-step bias corrections are derived on the host from the step number
};

// kind 0: AE (8 Linears); kind 1: AE_Dropout_BN (models.py:256-313), whose flat parameter vector continues after the
// Linears with gamma / beta of the four decoder BatchNorms at bn_g_off / bn_b_off (n_params_total values in all)
int bb_tc_train_create(bb_ctx* ctx, const int* dims /*9*/, const int* acts /*8*/, int max_batch, float* params, float* m,
                       float* v, float* grads, int kind, const int* bn_g_off /*4*/, const int* bn_b_off /*4*/,
                       int n_params_total, TcTrainer** out);
// AE_Dropout_BN: running statistics (concatenated over the four BatchNorms) and num_batches_tracked[4], device memory
int bb_tc_train_set_bn(TcTrainer* t, float* running_mean, float* running_var, long long* batches_tracked);
// dropout stream: Philox seed, or four injected keep-mask arrays [global batch][width] (uint8, device) for parity tests
int bb_tc_train_set_dropout(TcTrainer* t, unsigned long long seed, const unsigned char* const* masks_dev);
// train = 1: dropout on, batch statistics (updates the running statistics); 0: eval.  drop_step: dropout stream position
// of the first step of the next launch (one position per batch)
void bb_tc_train_set_mode(TcTrainer* t, int train, unsigned long long drop_step);
void bb_tc_train_destroy(TcTrainer* t);
// rebuild the packed fp16 hi / lo weight images from the flat fp32 parameters (after creation, or after the fp32
// kernels moved the parameters)
int bb_tc_train_repack(TcTrainer* t, cudaStream_t s);
// n_steps consecutive batches of `batch` rows starting at x (the last one may be ragged: n_rows total).  `first_step` is
// the 1-based Adam step number of the first batch.  loss_accum (double, device) receives the sum of batch losses when
// TC_ADAM or TC_FWD_ONLY is set.
// dp_slice (data parallel, after bb_tc_train_dp_connect): x is the full table and `batch` the global batch; every rank
// takes its contiguous share of each global batch.  0: x / batch are this rank's own rows.
int bb_tc_train_run(TcTrainer* t, const float* x, int64_t n_rows, int batch, int flags, const TcHyper* h, long long first_step,
                    double* loss_accum, int dp_slice, cudaStream_t s);
// Adam from the flat gradient vector (data-parallel callers all-reduce it in between); adds the loss slot to loss_accum
int bb_tc_train_adam_flat(TcTrainer* t, const TcHyper* h, long long step, double* loss_accum, cudaStream_t s);
// mean over the rows of the last forward pass of the inputs of layers 1,2,3,5,6,7 (hidden activations), NaN padded
int bb_tc_train_activation_means(TcTrainer* t, int rows, double* out_6x200);
int bb_tc_train_range_flag(TcTrainer* t, int reset, int* out);
// diagnostics: the feature-major scratch of the last step as fp32 [features][rows]: which = 0 the input of `layer`
// (K + 1 features, the last one the bias column of ones), 1 its pre-activation gradient (N features); returns #features
int bb_tc_train_debug_layer(TcTrainer* t, int which, int layer, int rows, float* out, int capacity);

// diagnostics: arm clock64 stamps of CTA 0 for step `step` of the following launches (step < 0: off) and read back the
// stamps of the previous armed launch: [0] step start, [1] phase 1 done, [2] barrier passed, [3] phase 2 done,
// [4] second barrier passed, [8 + p] / [32 + p] / [48 + p] start / end of MMAs / end of epilogue of layer pass p.
int bb_tc_train_profile(TcTrainer* t, int step, long long* out_128);

// data parallel over NVLink peer memory: every rank exports one cudaMalloc'd block (flags + exchange buffers) as a CUDA
// IPC handle, the host side all-gathers the handles, connect maps the peers' blocks.  From then on the weight-gradient
// phase pushes every gradient tile into all ranks' buffers and sums the world's tiles in rank order before Adam.
int bb_tc_train_dp_export(TcTrainer* t, int world, unsigned char* handle_out_64);
int bb_tc_train_dp_connect(TcTrainer* t, int rank, int world, const unsigned char* handles);
int bb_tc_train_dp_world(const TcTrainer* t);
