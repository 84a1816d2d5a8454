/* iins_b200.h -- C ABI of the B200-native IIns-VAE hot path (libiins_b200.so).
 *
 * The reference (JadeLilyx/IIns-VAE) is pure Python/PyTorch and has no FFI of its own: its
 * boundary for this path is the nn.Module API of models.py.  Each entry point below therefore
 * replaces one reference *method*; the Python classes in iins_vae_b200/models.py keep the
 * reference's constructor / forward signatures and state_dict keys and bind these symbols with
 * ctypes (see INTEGRATION.md for the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless stated; tensors are contiguous;
 *   - "params" / "grads" are arrays (on the HOST) of device pointers, one per parameter tensor, in
 *     the module's named_parameters() order (== state_dict order without the AdaIN dummy buffers);
 *   - no allocation inside: the caller passes workspaces sized by the *_floats() queries;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), nothing synchronises;
 *   - gradients are ACCUMULATED into `grads` (zero them first) -- weight gradients are reduced with
 *     atomics; gradient outputs w.r.t. module inputs are overwritten unless `accumulate` is set;
 *   - return value: 0 on success, negative iins_status otherwise; iins_last_error() gives text.
 */
#ifndef IINS_B200_H
#define IINS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* iins_stream_t;      /* cudaStream_t */

typedef enum iins_status {
    IINS_OK = 0,
    IINS_ERR_BAD_CONFIG = -1,     /* unsupported shape options (fails loudly, never falls back) */
    IINS_ERR_NULL = -2,
    IINS_ERR_CUDA = -3
} iins_status;

/* Shape options of the 1-D path: models.py:33 (Encoder), :68 (Decoder), :97 (Restorer),
 * :118 (Classifier); call sites train_semi.py:77-82. */
typedef struct iins_config {
    int batch;          /* B                                              */
    int cir_len;        /* 157 (zenodo) / 152 (ewine): train_semi.py:44   */
    int dim;            /* Encoder/Decoder dim, 4                         */
    int n_residual;     /* 3                                              */
    int n_downsample;   /* 4 (== Decoder n_upsample)                      */
    int env_dim;        /* style_dim, 16                                  */
    int range_dim;      /* out_dim, 2                                     */
    int num_classes;    /* Classifier num_classes                         */
    int cls_filters;    /* Classifier filters, 16                         */
    int conv_type;      /* 1 (or 0): Conv1d path; 2: the Conv2d variant with expand = True (models.py:33-61, 179-346, 474-539) */
} iins_config;

int iins_abi_version(void);

/* ---- Context: everything the library keeps between calls (compute mode, tuning switches read from the environment once at
 * creation, the helper streams / events it forks a caller's stream into) lives in an explicit handle.  A context is bound to
 * the device that was current when it was created and may be shared by host threads.  iins_ctx_make_current() selects the
 * calling THREAD's context for all entry points below; a thread without one uses the process's default context, so callers
 * that never touch this section behave as before.  (The launch counter / in-process profile at the end of this header are
 * process-wide diagnostics.) */
typedef struct iins_ctx iins_ctx;
iins_ctx* iins_ctx_create(void);
void iins_ctx_destroy(iins_ctx* ctx);
int iins_ctx_make_current(iins_ctx* ctx);            /* NULL: back to the default context */
iins_ctx* iins_ctx_get_current(void);
int iins_ctx_set_compute_mode(iins_ctx* ctx, int mode);
int iins_ctx_get_compute_mode(const iins_ctx* ctx);
int iins_ctx_set_stream_concurrency(iins_ctx* ctx, int enable);

/* Arithmetic of the conv / linear GEMMs (of the calling thread's current context):
 *   0 (default) tcgen05 tensor cores, fp32-grade: every fp32 operand is split into three bf16 pieces and the
 *               six significant piece products are accumulated in fp32 in TMEM ("bf16x3", error ~2^-23);
 *   1           tcgen05 tensor cores, plain bf16 operands, fp32 accumulation (BASELINE configs[2]);
 *   2           fp32 SIMT (FFMA) kernels: the bring-up path, kept as an on-device cross-check. */
int iins_set_compute_mode(int mode);
int iins_get_compute_mode(void);
const char* iins_last_error(void);
int iins_validate_config(const iins_config* cfg);

/* ---- Encoder: models.py:49-61 (RangeEncoder1d :140-176, EnvEncoder1d :258-298) ------------------
 * x (B,cir_len) -> range_code (B,range_dim,code_len) NCL, env_code (B,env_dim) [= cat],
 * env_code_rv (B,env_dim/2) [= noise*exp(log_sigma)+mu], kl (scalar, zeroed then accumulated).
 * noise: explicit (B,env_dim/2) standard normals, or NULL -> Philox4x32-10(seed, offset).
 * ws: iins_encoder_ws_floats() floats holding the activations saved for backward. */
int iins_encoder_num_params(const iins_config* cfg);
size_t iins_encoder_ws_floats(const iins_config* cfg);
size_t iins_encoder_scratch_floats(const iins_config* cfg);
int iins_encoder_forward(const iins_config* cfg, const float* const* params, const float* x,
                         const float* noise, uint64_t seed, uint64_t offset,
                         float* range_code, float* env_code, float* env_code_rv, float* kl,
                         float* ws, iins_stream_t stream);
/* backward of the above; any of d_range_code / d_env_code / d_env_code_rv / d_kl may be NULL (== 0).
 * d_kl is a device scalar.  range_code / env_code are the forward outputs. */
int iins_encoder_backward(const iins_config* cfg, const float* const* params, const float* noise,
                          uint64_t seed, uint64_t offset, const float* range_code, const float* env_code,
                          const float* ws, const float* d_range_code, const float* d_env_code,
                          const float* d_env_code_rv, const float* d_kl, float* const* grads,
                          float* scratch, iins_stream_t stream);

/* ---- Decoder: models.py:81-91 (Decoder1d :405-471, MLP :951-962, AdaIN :1048-1076, LN :965-985) ---
 * (range_code NCL, env_code) -> x_recon (B,cir_len). */
int iins_decoder_num_params(const iins_config* cfg);
size_t iins_decoder_ws_floats(const iins_config* cfg);
size_t iins_decoder_scratch_floats(const iins_config* cfg);
int iins_decoder_forward(const iins_config* cfg, const float* const* params, const float* range_code,
                         const float* env_code, float* x_recon, float* ws, iins_stream_t stream);
int iins_decoder_backward(const iins_config* cfg, const float* const* params, const float* range_code,
                          const float* env_code, const float* ws, const float* d_x_recon,
                          float* const* grads, float* d_range_code, float* d_env_code, int accumulate,
                          float* scratch, iins_stream_t stream);

/* ---- 2-D variant, SURVEY.md 8(f) row 3 (conv_type = 2, expand = True; models.py:179-215 RangeEncoder2d, :304-346 EnvEncoder2d,
 * :474-539 Decoder2d, :1008-1025 ResidualBlock2d, :1082-1113 AdaptiveInstanceNorm2d).  Same argument meaning as the 1-D entry
 * points; range_code is (B, range_dim, 8, 8) NCHW; cfg->conv_type must be 2.  The Restorer / Classifier entry points take the
 * same cfg (the Restorer's first Linear then has range_dim * 64 inputs). */
size_t iins_encoder2d_ws_floats(const iins_config* cfg);
size_t iins_encoder2d_scratch_floats(const iins_config* cfg);
int iins_encoder2d_forward(const iins_config* cfg, const float* const* params, const float* x,
                           const float* noise, uint64_t seed, uint64_t offset,
                           float* range_code, float* env_code, float* env_code_rv, float* kl,
                           float* ws, iins_stream_t stream);
int iins_encoder2d_backward(const iins_config* cfg, const float* const* params, const float* noise,
                            uint64_t seed, uint64_t offset, const float* range_code, const float* env_code,
                            const float* ws, const float* d_range_code, const float* d_env_code,
                            const float* d_env_code_rv, const float* d_kl, float* const* grads,
                            float* scratch, iins_stream_t stream);
size_t iins_decoder2d_ws_floats(const iins_config* cfg);
size_t iins_decoder2d_scratch_floats(const iins_config* cfg);
int iins_decoder2d_forward(const iins_config* cfg, const float* const* params, const float* range_code,
                           const float* env_code, float* x_recon, float* ws, iins_stream_t stream);
int iins_decoder2d_backward(const iins_config* cfg, const float* const* params, const float* range_code,
                            const float* env_code, const float* ws, const float* d_x_recon,
                            float* const* grads, float* d_range_code, float* d_env_code, int accumulate,
                            float* scratch, iins_stream_t stream);

/* ---- Restorer (RestorerLinear, soft=False): models.py:642-658.  range_code -> err_est (B,1).
 * params: 10 tensors (layers.{0,2,4}, linear_layer1, linear_layer2); linear_layer2 is never touched
 * and its gradient slots are left as they are (grad None in the reference). */
int iins_restorer_num_params(const iins_config* cfg);
size_t iins_restorer_ws_floats(const iins_config* cfg);
size_t iins_restorer_scratch_floats(const iins_config* cfg);
int iins_restorer_forward(const iins_config* cfg, const float* const* params, const float* range_code,
                          float* err_est, float* ws, iins_stream_t stream);
int iins_restorer_backward(const iins_config* cfg, const float* const* params, const float* range_code,
                           const float* ws, const float* d_err_est, float* const* grads,
                           float* d_range_code, int accumulate, float* scratch, iins_stream_t stream);

/* ---- Classifier (ClassifierLinear): models.py:858-862.  env_code -> logits (B,num_classes). */
int iins_classifier_num_params(const iins_config* cfg);
size_t iins_classifier_ws_floats(const iins_config* cfg);
size_t iins_classifier_scratch_floats(const iins_config* cfg);
int iins_classifier_forward(const iins_config* cfg, const float* const* params, const float* env_code,
                            float* logits, float* ws, iins_stream_t stream);
int iins_classifier_backward(const iins_config* cfg, const float* const* params, const float* env_code,
                             const float* ws, const float* d_logits, float* const* grads,
                             float* d_env_code, int accumulate, float* scratch, iins_stream_t stream);

/* ---- soft Restorer (RestorerLinear, soft=True): models.py:634-655 --------------------------------------------------------
 * The trunk ends in linear_layer2 (256 -> 2: mu, logvar); z = noise * exp(logvar / 2) + mu with noise of shape (B, 1) and
 * mu / logvar of shape (B,), which torch BROADCASTS to (B, B): z[i][j] = noise[i] * std[j] + mu[j] -- the reference's
 * behaviour, reproduced as is.  noise: (B,) standard normals (the reference draws np.random.normal on the host).
 * params / grads: the Restorer's 10 tensors; linear_layer1 is not touched (grad None in the reference when soft=True). */
size_t iins_restorer_soft_ws_floats(const iins_config* cfg);
size_t iins_restorer_soft_scratch_floats(const iins_config* cfg);
int iins_restorer_soft_forward(const iins_config* cfg, const float* const* params, const float* range_code, const float* noise,
                               float* z, float* ws, iins_stream_t stream);
int iins_restorer_soft_backward(const iins_config* cfg, const float* const* params, const float* range_code, const float* noise,
                                const float* ws, const float* d_z, float* const* grads, float* d_range_code, int accumulate,
                                float* scratch, iins_stream_t stream);

/* ---- Conv1d heads (net_type='Conv1d'): RestorerConv1d models.py:661-716, ClassifierConv1d models.py:865-902 --------------
 * Conv1d + LeakyReLU(0.2) + Dropout(0.25) blocks, BatchNorm1d(eps = 0.8: the second positional argument of
 * nn.BatchNorm1d(c, 0.8) is eps) on the second block, Linear output (+ LeakyReLU(0.2) on the classifier's logits).
 * params (named_parameters order): conv_blocks.0.{weight,bias}, conv_blocks.3.{weight,bias}, conv_blocks.6.{weight,bias}
 * (BatchNorm), then linear_layer1.{weight,bias} (+ linear_layer2.0.{weight,bias}, never touched) / linear.0.{weight,bias}. */
typedef struct iins_head_state {
    int training;              /* 1: dropout active, batch statistics, running buffers updated (nn.Module.train()); 0: eval */
    const float* mask1;        /* explicit dropout keep-masks (0 / 1) in the reference's (B, C, L) element order, or NULL: */
    const float* mask2;        /*   Philox4x32-10(seed, offset) decides, reproducibly in forward and backward            */
    uint64_t seed, offset;
    float* running_mean;       /* BatchNorm buffers [C] (device); updated with momentum 0.1 in training, read in eval      */
    float* running_var;
    int64_t* num_batches_tracked;
    double* bn_stats;          /* device double[4 * C] scratch: [0,2C) forward sums (x, x^2), [2C,4C) backward sums (dy, dy*xhat) */
    int phase;                 /* 0: whole call.  1: stop after the LOCAL batch sums are in bn_stats; 2: continue from bn_stats --    */
                               /*    a data-parallel caller all-reduces the 2 * C doubles in between (SyncBN)                          */
    double count_scale;        /* number of ranks whose sums were added into bn_stats (1 without SyncBN)                              */
    int64_t sample_offset;     /* index of this rank's first sample in the global batch: Philox counters use GLOBAL sample indices,  */
                               /*    so N ranks x B/N samples draw the masks one rank x B would                                        */
    const int32_t* offset_dev; /* optional DEVICE counter added to `offset` when the Philox masks are drawn: a captured CUDA graph */
                               /*    replays the same kernel arguments, so the per-step offset has to come from device memory         */
} iins_head_state;
size_t iins_restorer_conv_ws_floats(const iins_config* cfg);
size_t iins_restorer_conv_scratch_floats(const iins_config* cfg);
int iins_restorer_conv_forward(const iins_config* cfg, const float* const* params, const float* range_code, float* err_est,
                               float* ws, const iins_head_state* state, iins_stream_t stream);
int iins_restorer_conv_backward(const iins_config* cfg, const float* const* params, const float* range_code, const float* ws,
                                const float* d_err_est, float* const* grads, float* d_range_code, int accumulate,
                                float* scratch, const iins_head_state* state, iins_stream_t stream);
size_t iins_classifier_conv_ws_floats(const iins_config* cfg);
size_t iins_classifier_conv_scratch_floats(const iins_config* cfg);
int iins_classifier_conv_forward(const iins_config* cfg, const float* const* params, const float* env_code, float* logits,
                                 float* ws, const iins_head_state* state, iins_stream_t stream);
int iins_classifier_conv_backward(const iins_config* cfg, const float* const* params, const float* env_code, const float* ws,
                                  const float* d_logits, float* const* grads, float* d_env_code, int accumulate,
                                  float* scratch, const iins_head_state* state, iins_stream_t stream);

/* ---- Fused loss + seed gradients + metrics: train_semi.py:199-225, train.py:87-91, :104-115 -------
 * out[8] (zeroed inside): [0] mean|x-x_recon|  [1] mean|err-err_est|  [2] mean CE  [3] lam-weighted sum
 * of [0..2] (the KL term lives in the encoder)  [4] mean (err_est-err)^2  [5] #correct argmax
 * [6] number of labels outside [0, num_classes) after the offset (must be 0: torch's CrossEntropyLoss device-asserts).
 * x/x_recon may be NULL (train.py variant), err/err_est/logits/label may be NULL (unsupervised batch).
 * label: fp32 holding integers (dataset.py:122) or, if label_i64 != NULL, int64.
 * label_offset: class index = label - label_offset.  train_semi.py:217-222 feeds CrossEntropyLoss `label_gt - 1` and
 * scores `argmax + 1` for every dataset_env except 'room_full' (labels 1..NC): pass 1 there, 0 for room_full / train.py. */
int iins_loss_forward_backward(int batch, int cir_len, int num_classes,
                               const float* x, const float* x_recon, const float* err, const float* err_est,
                               const float* logits, const float* label, const int64_t* label_i64, int label_offset,
                               float lam_ae, float lam_res, float lam_env, float* out,
                               float* d_x_recon, float* d_err_est, float* d_logits, int32_t* pred,
                               iins_stream_t stream);

/* ---- Fused Adam over a flat parameter buffer: torch.optim.Adam as used at train_semi.py:118-122 -----
 * groups: up to 7 half-open element ranges [begin,end) with an `active` flag; an inactive group is
 * skipped entirely (grad None in the reference).  steps: device int32[n_groups + 1]: completed updates per
 * group, advanced by this call for active groups (by the last CTA of the ONE launch), plus a ticket word
 * that must be zero when the call is enqueued.  lr: device scalar.
 * grad_scale: gradients are multiplied by it first (1/world_size after a SUM all-reduce = the data-parallel mean);
 * zero_grads != 0: consumed gradient elements are zeroed (the backward passes accumulate: no separate memset). */
int iins_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                   const int64_t* group_begin, const int64_t* group_end, const int32_t* group_active,
                   int n_groups, int32_t* steps, const float* lr, double beta1, double beta2, float eps,
                   float grad_scale, int zero_grads, iins_stream_t stream);

/* ---- small helpers used at the module boundary ----------------------------------------------------- */
/* AdaptiveAvgPool1d over (B,Lin)->(B,Lout) (models.py:146, 436) and its backward */
int iins_adaptive_pool_forward(const float* x, float* y, int batch, int lin, int lout, iins_stream_t stream);
int iins_adaptive_pool_backward(const float* dy, float* dx, int batch, int lin, int lout, iins_stream_t stream);

/* dst1[0..n1) += src1, dst2[0..n2) += src2 in ONE launch.  The reference sums the gradients of range_code / env_code
 * arriving from the decoder and from the two heads inside autograd (train_semi.py:225-227: one backward() through
 * Dec, Res and Cls); the engine runs the heads concurrently with the decoder and adds their contributions here. */
int iins_accumulate2(float* dst1, const float* src1, size_t n1, float* dst2, const float* src2, size_t n2,
                     iins_stream_t stream);

/* Inside one module pass the library overlaps independent work on helper streams (weight gradients next to the
 * data-gradient chain, env encoder next to the range encoder); 0 serialises everything on the caller's stream --
 * used when individual kernels are timed with events (bench.py roofline, tools/step_profile.py). Default 1. */
int iins_set_stream_concurrency(int enable);

/* Deferred joins.  By default a backward entry point returns with the caller's stream ordered after ALL its work, weight
 * gradients included (what an autograd caller needs).  With iins_set_deferred_join(1) (current context) the stream is only
 * ordered after the data-gradient outputs; the weight gradients of that pass keep running on the context's helper stream of
 * that caller stream, next to whatever the caller enqueues next (the following module's backward).  The caller must then
 * call iins_join_helpers(producer, waiter, 0) before the gradients are consumed: `waiter` (NULL = `producer` itself) waits for
 * the outstanding weight-gradient work of `producer`.  keep_pending = 1 adds a waiter without settling the join (a
 * communication stream that reduces one bucket early; the final join still has to follow). */
int iins_set_deferred_join(int enable);
int iins_join_helpers(iins_stream_t producer, iins_stream_t waiter, int keep_pending);

/* ---- launch accounting / in-process kernel timing (used by bench.py; not a profiler replacement) ------ */
unsigned long long iins_launch_count(void);           /* kernels launched by this library so far */
int iins_profile_begin(void);                          /* record a CUDA-event pair around every launch */
int iins_profile_collect(const char** names, float* ms, double* flops, int cap);
int iins_profile_bytes(double* bytes, int cap);       /* after collect: algorithmic HBM bytes of each launch (0 = not modelled) */
int iins_profile_shapes(int* shapes_mnk, int cap);   /* after collect: (M,N,K) of each GEMM launch, zeros otherwise */
        /* sync; per launch: kernel name, milliseconds, algorithmic FLOPs (2*M*N*K for the GEMM kernels, else 0) */

#ifdef __cplusplus
}
#endif
#endif /* IINS_B200_H */
