/* rcu_b200 — C-ABI of the B200-native stochastic-inference + uncertainty hot path.
 *
 * The reference (alainjungo/reliability-challenges-uncertainty) is pure Python and has no FFI; its seams on
 * this path are Python protocols.  Each entry point below names the reference interface it stands behind
 * (path:line relative to the reference root).  The Python drop-ins in
 * reliability-challenges-uncertainty_b200/ bind these symbols with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative rcu_status; rcu_last_error() gives the message of
 *     the calling thread's last failure;
 *   - all data pointers are DEVICE pointers owned by the caller unless marked "host"; small parameter
 *     tables (edges, thresholds, break tables) are host pointers read before the call returns;
 *   - every call is asynchronous and ordered on the given stream (a cudaStream_t passed as void*);
 *   - the handle's only device allocations are its folded weights (rcu_unet_create); the activation workspace and the
 *     metric workspace are caller-provided buffers;
 *   - one rcu_unet handle per (device, weight set); a handle is not thread-safe, distinct handles are.
 */
#ifndef RCU_B200_H
#define RCU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCU_ABI_VERSION 3

typedef enum rcu_status {
  RCU_OK = 0,
  RCU_EINVAL = -1,   /* bad argument (the Python side turns these into the reference's ValueError) */
  RCU_ECUDA = -2,    /* CUDA runtime / driver error */
  RCU_ENOTSUP = -3,  /* configuration outside the supported hot path */
  RCU_ENOMEM = -4,   /* workspace too small */
  RCU_ENCCL = -5     /* NCCL not loadable, or an NCCL call failed */
} rcu_status;

#define RCU_MAX_BINS 32        /* ECE reliability bins per call */
#define RCU_MAX_UE_CLASSES 32  /* uncertainty intervals (= thresholds + 1) per call */
#define RCU_MAX_BREAKS 96      /* float32 break points of the p -> interval table */

int rcu_abi_version(void);
const char* rcu_last_error(void);
/* 0 if a usable sm_100 device is present, RCU_ECUDA otherwise (message says why). */
int rcu_device_check(int device);

/* ------------------------------------------------------------------------------------------------
 * Calibration / uncertainty-error histograms
 * ------------------------------------------------------------------------------------------------ */

/* Workspace (bytes) the metric calls below need for `n_subjects` subjects per launch. */
size_t rcu_metrics_workspace_bytes(int n_subjects);
/* Must be called once on a fresh workspace (zeroes the per-subject tickets and accumulator tables, which every launch
 * leaves at zero again); stream-ordered.  Pass the same workspace_bytes to every call that uses the workspace (tickets
 * and accumulators live at its end).  Calls that share a workspace must be ordered on one stream. */
int rcu_metrics_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

/* ECE reliability-bin tables.
 * Stands behind np_fn._binary_calibration (common/evalutation/numpyfunctions.py:51-69), reached from
 * EceBinaryNumpy.__call__ (common/evalutation/eval.py:129-142) -> ece_binary (:6-23) -> binary_calibration (:26-48).
 *   p         float32[n_subjects * voxels_per_subject]  foreground probability (prob[..., 1])
 *   target    uint8  [same]                            ground truth {0,1}
 *   mask      uint8  [same] or NULL                    voxels with mask==0 are skipped (BraTS T2>0 brain mask)
 *   edges_f32 host float32[n_bins+1]                   smallest float32 >= each float64 np.linspace(0, 1+1e-8, n_bins+1) edge
 *   range_lo/hi                                        `threshold_range` (exclusive both sides); pass NaN,NaN for None
 * Outputs, per subject, n_bins+1 slots each (slot n_bins counts values no bin takes: p<0, p>=last edge, NaN):
 *   count     uint64[n_subjects][n_bins+1]             == np.bincount(binids)            (bit-exact)
 *   positives uint64[n_subjects][n_bins+1]             == np.bincount(binids, weights=target) (exact integers)
 *   conf_sum  float64[n_subjects][n_bins+1]            == np.bincount(binids, weights=p) up to fp64 summation order;
 *                                                      deterministic run to run (fixed reduction tree)
 */
int rcu_calib_hist(const float* p, const uint8_t* target, const uint8_t* mask, int64_t voxels_per_subject,
                   int n_subjects, const float* edges_f32, int n_bins, float range_lo, float range_hi,
                   uint64_t* count, uint64_t* positives, double* conf_sum, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Joint histogram (confusion class x uncertainty interval), all sweep thresholds in one pass.
 * Stands behind np_fn.uncertainty (common/evalutation/numpyfunctions.py:86-107) as called, once per threshold,
 * by UncertaintyAndCorrectionEvalNumpy.__call__ (common/evalutation/eval.py:182-226) and
 * UncertaintyErrorDiceNumpy.__call__ (:156-173) from the CorrectionAction sweep (bin-eval/eval_uncertainty.py:195-202).
 *
 * The interval index of a voxel is j = seg_class[#{b : breaks[b] <= v}] (ordinary float comparison) where
 *   value_kind 0: v is the float32 foreground probability p; the host derives `breaks_f32` (every float32 value
 *                 at which some predicate u(p) > th_k flips, sorted) and `seg_class` (how many thresholds are
 *                 exceeded between consecutive breaks) from the reference's own float arithmetic
 *                 u(p) = H([1-p, p]) / ln 2 (numpyfunctions.py:166-168, analysis.py:201), so counts are bit-exact;
 *   value_kind 1: v is a float32 uncertainty map, breaks_f32[k] = smallest float32 > th_k, seg_class = identity;
 *   value_kind 2: v is a float64 uncertainty map (what ToEntropy produces); breaks_f64 = the thresholds and the
 *                 comparison is strict (b < v), seg_class = identity.
 * Breaks must be sorted ascending.  Output ue_counts uint64[n_subjects][4][n_classes]: rows tp, tn, fp, fn
 * (prediction/target as booleans), column j = voxels whose uncertainty exceeds exactly j of the thresholds.
 * Thresholded counts are suffix sums: tpu(th_k) = sum_{j>k} ue_counts[tp][j]; tp = sum_j ue_counts[tp][j].
 * `invalid` (uint64[n_subjects], may be NULL) counts values that are negative or NaN (kinds 0 and 1).
 * `mask` (may be NULL) restricts the voxels (UncertaintyErrorDiceNumpy(with_mask=True): mask = ~target_boarder).
 */
int rcu_ue_hist(const void* values, int value_kind, const uint8_t* prediction, const uint8_t* target,
                const uint8_t* mask, int64_t voxels_per_subject, int n_subjects, const float* breaks_f32,
                const double* breaks_f64, int n_breaks, const uint8_t* seg_class, int n_classes,
                uint64_t* ue_counts, uint64_t* invalid, void* workspace, size_t workspace_bytes, void* stream);

/* Both of the above in one pass over (p, prediction, target, mask): 7 bytes per voxel.
 * `mask` applies to the calibration tables only (as in EceAction, bin-eval/eval_uncertainty.py:151-152 vs
 * CorrectionAction :195-202 which is unmasked).
 * For 16-byte aligned p and 4-byte aligned byte maps (and voxels_per_subject % 4 == 0 when n_subjects > 1) this runs the
 * shared-atomic kernel (csrc/eval_atom.cuh): integer tables bit-exact as always; conf_sum of bins 1 .. n_bins-1 is the
 * EXACT sum (integers q = p * 2^(23+K) accumulated exactly, converted once: independent of the launch shape), bin 0 a
 * float64 sum in a fixed order; conf_sum of the overflow slot n_bins is 0 there (a sum over NaN / out-of-range values
 * carries no meaning; `count` and `invalid` still report them). */
int rcu_eval_fused(const float* p, const uint8_t* prediction, const uint8_t* target, const uint8_t* mask,
                   int64_t voxels_per_subject, int n_subjects, const float* edges_f32, int n_bins,
                   const float* breaks_f32, int n_breaks, const uint8_t* seg_class, int n_classes,
                   uint64_t* count, uint64_t* positives, double* conf_sum, uint64_t* ue_counts, uint64_t* invalid,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Confusion counts with pymia 0.2.1 `ConfusionMatrix` semantics (prediction == 1 / == 0 vs label == 1 / == 0), the
 * arithmetic behind np_fn.dice / confusion_matrx / accuracy (common/evalutation/numpyfunctions.py:128-151).
 * counts uint64[n_subjects][4] = tp, tn, fp, fn (zeroed by the call). */
int rcu_confusion(const uint8_t* prediction, const uint8_t* target, int64_t voxels_per_subject, int n_subjects,
                  uint64_t* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-pixel aggregation over samples / ensemble members
 * ------------------------------------------------------------------------------------------------ */

/* Stands behind MultiPredictionSummary.__call__ (rechun/dl/customsteps.py:50-71) + th.entropy
 * (common/utils/torchhelper.py:53-54), with the per-sample F.softmax(logits, 1) of McPredictStep
 * (rechun/dl/customsteps.py:31-36) / EnsemblePredictionStep (bin-dl/brats_test_ensemble.py:85-93) folded in.
 *   input_kind 0: logits, pixel-interleaved  float32[n_samples][n_images][hw][2]   (what rcu_unet_forward emits)
 *   input_kind 1: probabilities, planar      float32[n_samples][n_images][2][hw]   (a reference-made
 *                 `multi_probabilities` tensor; no softmax applied)
 *   input_kind 2: logits, planar             float32[n_samples][n_images][2][hw]
 *   input_kind 3: logit differences l0 - l1  float32[n_samples][n_images][hw]      (rcu_unet_outputs.logit_diff; the softmax
 *                 of two classes is a function of the difference alone: same outputs as kind 0, half the bytes)
 * Outputs (any may be NULL except mean):
 *   mean       float32[n_images][2][hw]   sum_t p_t / n_samples  (sequential fp32 sum in t, like torch)
 *   entropy    float32[n_images][hw]      -sum_c where(p>0, p ln p, 0) of the mean
 *   mutual_info float32[n_images][hw]     entropy - mean_t H(p_t)
 *   variance   float32[n_images][hw]      mean_c var_t(p_t,c), unbiased
 *   prediction uint8  [n_images][hw]      argmax_c mean (ties -> class 0, like np.argmax)
 *   foreground float32[n_images][hw]      mean[:, 1] as a dense map — what WriteHook saves as `*_probabilities.nii.gz`
 *              (bin-dl/brats_test_default.py:98) and the metric kernels read
 *   multi_out  float32[n_samples][n_images][2][hw] planar per-sample probabilities (the reference's
 *              `multi_probabilities`), only written when non-NULL
 */
int rcu_aggregate(const float* input, int input_kind, int n_samples, int64_t n_images, int64_t hw, float* mean,
                  float* entropy, float* mutual_info, float* variance, uint8_t* prediction, float* foreground,
                  float* multi_out, void* stream);

/* McPredictStep + MultiPredictionSummary in ONE pass (rechun/dl/customsteps.py:23-25 and :50-71): as rcu_aggregate on the
 * interleaved per-sample logits (input_kind 0), and in the same launch the deterministic weight-scaling sample
 *   ws_logits float32[n_images][hw][2]  ->  ws_probabilities float32[n_images][2][hw] = softmax  (`ws_probabilities`). */
int rcu_aggregate_ws(const float* logits, int n_samples, int64_t n_images, int64_t hw, const float* ws_logits,
                     float* ws_probabilities, float* mean, float* entropy, float* mutual_info, float* variance,
                     uint8_t* prediction, float* foreground, void* stream);
/* The same on logit DIFFERENCES l0 - l1 (input_kind 3 of rcu_aggregate; what rcu_unet_outputs.logit_diff holds):
 *   logit_diff float32[n_samples][n_images][hw], ws_logit_diff float32[n_images][hw].
 * F.softmax over two classes (customsteps.py:33) depends on the difference alone — p_max = 1 / (1 + exp(-|d|)) — so every
 * output is bit-identical to rcu_aggregate_ws on the logit pairs, for half the bytes read (4 instead of 8 per voxel-sample). */
int rcu_aggregate_ws_diff(const float* logit_diff, int n_samples, int64_t n_images, int64_t hw, const float* ws_logit_diff,
                          float* ws_probabilities, float* mean, float* entropy, float* mutual_info, float* variance,
                          uint8_t* prediction, float* foreground, void* stream);

/* Partial aggregation for sample-sharded runs: writes the raw fp32 sums so ranks can allreduce them.
 *   sums float32[n_images][K][hw] with K = 2 (sum p0, sum p1) [+1: sum_t H(p_t) if want_mi] [+2: sum p0^2, sum p1^2 if want_var]
 * and the finisher that turns allreduced sums into the same outputs as rcu_aggregate. */
int rcu_aggregate_partial(const float* input, int input_kind, int n_samples, int64_t n_images, int64_t hw,
                          int want_mi, int want_var, float* sums, void* stream);
int rcu_aggregate_finish(const float* sums, int total_samples, int64_t n_images, int64_t hw, int has_mi, int has_var,
                         float* mean, float* entropy, float* mutual_info, float* variance, uint8_t* prediction,
                         float* foreground, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Philox4x32-10 Dropout2d keep masks
 * ------------------------------------------------------------------------------------------------ */

/* Materialises the keep-scale table the conv epilogues otherwise generate in registers (same stream):
 *   scale[(sample * n_slices + slice) * total_channels + site_offset[s] + c]
 *       = philox4x32_10(ctr=(c/4, s, slice_index0 + slice, sample0 + sample), key=seed)[c%4] >= thr ? 1/(1-p) : 0
 * Stands behind nn.Dropout2d in train mode inside Conv2dBnRelu (common/model/unet.py:14-15) as toggled by
 * th.set_dropout_mode (common/utils/torchhelper.py:44-50).  site_channels is a host array.
 * rcu_philox_masks_host computes the identical table on the host (no GPU needed) for injection into the reference. */
int rcu_philox_masks(uint64_t seed, float p_drop, const int* site_channels, int n_sites, int64_t slice_index0,
                     int64_t n_slices, int sample0, int n_samples, float* scale, void* stream);
int rcu_philox_masks_host(uint64_t seed, float p_drop, const int* site_channels, int n_sites, int64_t slice_index0,
                          int64_t n_slices, int sample0, int n_samples, float* scale_host);

/* ------------------------------------------------------------------------------------------------
 * Evaluation-side preparations (between the saved maps and the metric kernels)
 * ------------------------------------------------------------------------------------------------ */

/* common/utils/labelhelper.py:12-20 boarder_mask(binary_label_map, distance_in, distance_out) as used by
 * rechun/eval/analysis.py:109-116 (distance_in = distance_out = 1): mask[v] = (dist_in[v] <= distance_in) *
 * (dist_out[v] <= distance_out) with the Euclidean distance transforms of the map and its complement, unit spacing.
 * label / mask: uint8[d0][d1][d2] device arrays (a 2-D map is d0 = 1).  Maps without any voxel of the opposite class
 * give an all-zero mask (scipy's transform is not defined there). */
int rcu_border_mask(const uint8_t* label, int d0, int d1, int d2, int distance_in, int distance_out, uint8_t* mask, void* stream);

/* entry_np.min() / entry_np.max() of RescaleSubjectMinMax (rechun/eval/analysis.py:169-178).  out3 is a DEVICE array of
 * three uint32: order-preserving keys of min and max (decode with rcu_minmax_decode) and the number of NaNs skipped. */
int rcu_minmax(const float* values, int64_t n, uint32_t* out3, void* stream);
float rcu_minmax_decode(uint32_t key);

/* rechun/eval/helper.py:18-21 rescale_uncertainties followed by helper.py:7-15 uncertainty_to_foreground_probabilities:
 *   p = ((u - lo) / range) * scale + epsilon      (only when rescale != 0; each operation rounded to float32 like numpy)
 *   foreground = prediction == 1 ? 1 - p * 0.5 : p * 0.5
 * invalid2 (device, 2 x uint64): [0] values of p outside [0, 1] or NaN, [1] predictions > 1 — the two conditions the
 * reference raises ValueError for (helper.py:10-12,30-46). */
int rcu_confidence_to_foreground(const float* uncertainty, const uint8_t* prediction, int64_t n, int rescale, float lo, float range,
                                 float scale, float epsilon, float* foreground, uint64_t* invalid2, void* stream);

/* ------------------------------------------------------------------------------------------------
 * U-Net forward (tcgen05 implicit-GEMM convolutions)
 * ------------------------------------------------------------------------------------------------ */

typedef struct rcu_unet rcu_unet;

/* One 3x3 'conv -> [Dropout2d] -> BN(eval) -> ReLU' unit (Conv2dBnRelu, common/model/unet.py:8-23), host fp32. */
typedef struct rcu_conv_unit {
  const float* weight;       /* [c_out][c_in][3][3] as in nn.Conv2d.weight */
  const float* bias;         /* [c_out] */
  const float* bn_weight;    /* [c_out] gamma, or NULL when the unit has no BN (up-path `upconv`, head 1x1) */
  const float* bn_bias;      /* beta */
  const float* bn_mean;      /* running_mean */
  const float* bn_var;       /* running_var */
  int c_in, c_out;
  int has_dropout;           /* a Dropout2d sits between conv and BN */
} rcu_conv_unit;

/* Topology and weights of common/model/unet.py:123-186 UNet(nb_classes=2, in_channels, depth, start_filters,
 * dropout, dropout_center[, sigma_out]) with residual=False, bn=True.  Arrays are host pointers.
 *   units   : 2*depth + 2 + 2*depth + 1 Conv2dBnRelu units in forward order (down blocks, bottom, up blocks, conv_cls.0)
 *   upconvs : depth plain 3x3 convs `up_convs.i.upconv.1` (bias only, applied after nearest x2), bn_* = NULL
 *   head    : `conv_cls.1`, 1x1, c_in = start_filters, c_out = 2 (weight [2][c_in][1][1])
 */
typedef struct rcu_unet_desc {
  int in_channels, depth, start_filters, nb_classes;
  float p_drop;              /* Dropout2d p (unused if no unit has_dropout) */
  float bn_eps;
  const rcu_conv_unit* units;
  int n_units;
  const rcu_conv_unit* upconvs;
  int n_upconvs;
  rcu_conv_unit head;
  /* sigma_out=True nets (unet.py:162-164, config/train_brats_aleatoric.yaml:12): `conv_sigma.0` (a Conv2dBnRelu on the
   * same features as conv_cls.0) and `conv_sigma.1` (1x1 -> 2).  Both NULL for sigma_out=False. */
  const rcu_conv_unit* sigma_unit;
  const rcu_conv_unit* sigma_head;
  /* residual=True nets (ConvResidualBlock, unet.py:42-60): the 2 * depth + 1 `residual` 1x1 convolutions (weight [c_out][c_in],
   * bias; bn_* NULL) in forward order — down_convs.0 ... down_convs.{depth-1}, bottom_convs, up_convs.0 ... — each added onto
   * its block's output, whose last Conv2dBnRelu then has no ReLU.  NULL / 0 for residual=False. */
  const rcu_conv_unit* residuals;
  int n_residuals;
} rcu_unet_desc;

/* Folds BN into per-channel scale/shift, converts weights to the device layout, uploads them. */
int rcu_unet_create(const rcu_unet_desc* desc, int device, rcu_unet** out);
void rcu_unet_destroy(rcu_unet* net);

/* Fixes the spatial size and the number of images (slices x samples) processed per internal chunk and reports the
 * size of the activation workspace that configuration needs.  Nothing is allocated: the CALLER owns the workspace
 *  and hands it over with rcu_unet_bind_workspace — a 1024-byte aligned device buffer of at least
 * `workspace_bytes`, which may be shared by several handles of one device as long as their forwards are ordered on one
 * stream (the M members of an ensemble then need one arena, not M).  Binding lays the activations out in the buffer and
 * encodes the TMA descriptors, so re-bind after every rcu_unet_plan and whenever the buffer moves. */
int rcu_unet_plan(rcu_unet* net, int height, int width, int max_images_per_chunk, size_t* workspace_bytes);
int rcu_unet_bind_workspace(rcu_unet* net, void* workspace, size_t workspace_bytes);

/* Runs `n_samples` forwards of each of `n_slices` input slices (the samples are folded into the batch, weights
 * are read once per tile) and writes pixel-interleaved logits float32[n_samples][n_slices][H*W][2].
 * Stands behind `context.model(images)` = UNet.forward (common/model/unet.py:166-186) as called by
 * SegmentationPredictStep (common/trainloop/steps.py:84), McPredictStep (rechun/dl/customsteps.py:23,32) and
 * EnsemblePredictionStep (bin-dl/brats_test_ensemble.py:85,89).
 *   images        float32[n_slices][in_channels][H][W]   NCHW, as the reference hands them to the model
 *   dropout_mode  0: every Dropout2d in eval mode (identity) for all samples
 *                 1: sample index `sample0 + t` draws its masks from Philox(seed) — free-running MC dropout
 *                 2: masks come from `scale` (float32[n_stochastic_samples][n_slices][total_dropout_channels],
 *                    values 0 or 1/(1-p); n_stochastic_samples = n_samples - (det_first != 0)), e.g.
 *                    reference-supplied decisions
 *   det_first     when non-zero, sample 0 of the folded batch is the deterministic "weight scaling" pass of
 *                 McPredictStep (customsteps.py:23-25) and the remaining n_samples-1 are stochastic
 */
int rcu_unet_forward(rcu_unet* net, const float* images, int64_t n_slices, int n_samples, int dropout_mode,
                     int det_first, uint64_t seed, int64_t slice_index0, int sample0, const float* scale,
                     float* logits, void* stream);

/* PostNet (common/model/postnet.py:6-17): n_units 1x1 Conv2dBnRelu units (c_in = c_out = 32, weight [32][32][1][1],
 * eval mode) followed by a 1x1 logits conv 32 -> 2; the auxiliary-feature method runs it on `UNet.features`
 * (bin-dl/brats_test_auxiliary_feat.py:61-80). */
typedef struct rcu_postnet rcu_postnet;
int rcu_postnet_create(const rcu_conv_unit* units, int n_units, const rcu_conv_unit* head, float bn_eps, int device, rcu_postnet** out);
void rcu_postnet_destroy(rcu_postnet* post);
/* Stand-alone call on a materialised feature tensor: float32[n_images][32][hw] (NCHW) -> pixel-interleaved logits
 * float32[n_images][hw][2].  Stands behind `context.model(self.test_model.features)` (brats_test_auxiliary_feat.py:77). */
int rcu_postnet_forward(const rcu_postnet* post, const float* features, int64_t n_images, int64_t hw, float* logits, void* stream);

/* Optional outputs of rcu_unet_forward_ex; NULL members are skipped (and cost nothing).
 *   logits          as rcu_unet_forward (required)
 *   sigma           `out_sigma` of a sigma_out net (unet.py:185-186), raw conv_sigma output, same layout as logits;
 *                   AleatoricPredictStep's abs()/exp() (bin-dl/brats_test_aleatoric.py:62-66) is left to the caller
 *   features        `UNet.features` (unet.py:178-179): float32[n_samples][n_slices][start_filters][H][W]
 *   postnet/postnet_logits   run `postnet` on the features while they are still in the workspace (bf16, never
 *                   materialised as float32) and write its logits, same layout as `logits` */
typedef struct rcu_unet_outputs {
  float* logits;
  float* sigma;
  float* features;
  const rcu_postnet* postnet;
  float* postnet_logits;
  float* logit_diff;   /* alternative to `logits` (exactly one of the two is non-NULL): float32[n_samples][n_slices][H][W],
                        * l0 - l1 of the class head.  The stochastic-inference steps only ever take the softmax of the pair,
                        * which is a function of the difference: the head then writes 4 instead of 8 bytes per voxel-sample
                        * and rcu_aggregate (input_kind 3) / rcu_aggregate_ws_diff read half as much */
} rcu_unet_outputs;
int rcu_unet_forward_ex(rcu_unet* net, const float* images, int64_t n_slices, int n_samples, int dropout_mode,
                        int det_first, uint64_t seed, int64_t slice_index0, int sample0, const float* scale,
                        const rcu_unet_outputs* out, void* stream);

/* Introspection used by tests and the benchmark. */
int rcu_unet_total_dropout_channels(const rcu_unet* net);
/* Copies a named internal activation of the LAST chunk of the last forward to `out` as fp32 NHWC (debug/parity). */
int rcu_unet_debug_activation(rcu_unet* net, int index, float* out, size_t out_elems, void* stream);
/* Selects the convolution implementation: 0 = tcgen05/TMEM/TMA implicit GEMM (default, the product path: halo-tile
 * kernel for the thin high-resolution layers, per-tap kernel for the rest), 1 = straightforward CUDA-core
 * fp32-accumulate kernel over the same bf16 data (on-device cross-check only), 2 = tcgen05 per-tap kernel everywhere,
 * 3 = as 0 but the c_out = 32 layers run the pixel-row halo kernel instead of the pixel-pair one (A/B and parity). */
int rcu_unet_set_conv_impl(rcu_unet* net, int impl);
/* Debug: bit i of `mask` lets conv i (execution order, the first conv excluded) use the halo-tile kernel when it is
 * eligible; cleared bits fall back to the per-tap kernel.  Default: all ones. */
int rcu_unet_set_halo_mask(rcu_unet* net, uint64_t mask);
/* First-layer dedup (default on): nn.Dropout2d sits between the first conv + bias and its BatchNorm (common/model/unet.py:13-19),
 * so the first unit's output for a (sample, slice) is the slice's "every channel kept" image with the dropped channels
 * replaced by constants; it is stored once per slice (two variants: deterministic pass / all kept) and the consuming
 * convolution patches the dropped channels into its tiles.  Bit-identical results; 0 materialises it per (sample, slice). */
int rcu_unet_set_first_layer_dedup(rcu_unet* net, int enable);
/* Optional per-op device timing: when enabled, every kernel launch of rcu_unet_forward is bracketed by CUDA events
 * on the launch stream.  rcu_unet_read_timing synchronises those events and returns, per op of the schedule (see
 * rcu_unet_op_info), the accumulated milliseconds and launch count since the last read.  ms/launches hold n_ops
 * entries; entry n_ops-1 is the per-chunk coefficient-table kernel. */
int rcu_unet_enable_timing(rcu_unet* net, int enable);
int rcu_unet_num_ops(const rcu_unet* net);
/* kind: 0 first conv (CUDA cores), 1 tcgen05 conv, 2 max-pool, 3 coefficient table.  macs_per_image: algorithmic
 * multiply-accumulates of the reference layer this op computes (0 for non-conv ops). */
int rcu_unet_op_info(const rcu_unet* net, int op, int* kind, int64_t* macs_per_image, int* c_in, int* c_out, int* h, int* w);
/* MACs the schedule actually EXECUTES for that op per image: equal to the algorithmic figure of rcu_unet_op_info except
 * for the up-path convs, which run nearest-x2 + conv3x3 as four pre-summed 2x2-tap phase convolutions (2.25x fewer). */
int rcu_unet_op_executed_macs(const rcu_unet* net, int op, int64_t* macs_per_image);
int rcu_unet_read_timing(rcu_unet* net, float* ms, int64_t* launches, int n_ops);
/* Number of kernel launches issued by the last rcu_unet_forward on this handle. */
int64_t rcu_unet_last_launch_count(const rcu_unet* net);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU exchange steps (one process per GPU on one box).  The reference has none of this: its only
 * multi-GPU mechanism is nn.DataParallel in training (common/trainloop/context.py:223-233).
 * ------------------------------------------------------------------------------------------------ */
#define RCU_COMM_ID_BYTES 128    /* sizeof(ncclUniqueId) */
#define RCU_IPC_HANDLE_BYTES 64  /* sizeof(cudaIpcMemHandle_t) */

typedef struct rcu_comm rcu_comm;   /* an NCCL communicator (NCCL is resolved at run time from the process's libnccl.so.2) */

/* Rank 0 makes the id (host buffer of RCU_COMM_ID_BYTES), hands it to the other ranks by any means (the Python side uses
 * torch.distributed), every rank then creates its communicator — collective, like ncclCommInitRank. */
int rcu_comm_unique_id(void* id_out);
int rcu_comm_create(const void* id, int n_ranks, int rank, int device, rcu_comm** out);
void rcu_comm_destroy(rcu_comm* comm);

/* In-place fp32 sum over ranks of the probability-sum planes rcu_aggregate_partial wrote (MC samples or ensemble
 * members split over ranks, rechun/dl/customsteps.py:31-36 / bin-dl/brats_test_ensemble.py:78-94); follow it with
 * rcu_aggregate_finish on every rank. */
int rcu_allreduce_probsum(rcu_comm* comm, float* sums, int64_t n, void* stream);
/* In-place sum over ranks of the per-subject metric tables of subjects whose slices span ranks: the uint64 count tables
 * (count | positives | ue_counts | invalid, any concatenation) and the float64 confidence sums, as one grouped NCCL call. */
int rcu_allreduce_counts(rcu_comm* comm, uint64_t* counts, int64_t n_counts, double* conf_sums, int64_t n_conf, void* stream);

/* Fused "reduce, then finish" over NVLink peer memory — the product path of the probability-sum exchange.
 * Every rank owns one exchange REGION (device memory, CUDA-IPC mapped into every other rank's process) laid out as
 * rcu_peer_layout says: signal words, the partial sums [n_images][planes][hw] (where rcu_aggregate_partial writes), and the
 * outputs.  rcu_aggregate_finish_peer(region_ptrs[n_ranks], ...) on rank r:
 *   flag barrier (all partial sums complete)  ->  ONE kernel that reads pixel slice r of every rank's sums through the
 *   peer mappings, adds them in rank order, applies MultiPredictionSummary's arithmetic (rechun/dl/customsteps.py:57-71)
 *   and stores mean / entropy / [mutual_info] / [variance] / foreground / prediction into EVERY rank's region  ->  flag barrier.
 * Afterwards each region holds the complete outputs, bit-identical on all ranks.  `epoch` must be non-zero and strictly
 * increasing from call to call on a set of regions; the call is collective (every rank must make it). */
typedef struct rcu_peer_layout_t {
  int planes;            /* 2 [+1 if has_mi] [+2 if has_var] */
  size_t off_sums, off_mean, off_entropy, off_mi, off_var, off_foreground, off_prediction;
  size_t bytes;          /* size of one region */
} rcu_peer_layout_t;
int rcu_peer_layout(int64_t n_images, int64_t hw, int has_mi, int has_var, rcu_peer_layout_t* out);
/* Allocates a zeroed region on `device` and exports its IPC handle (host buffer of RCU_IPC_HANDLE_BYTES); the other ranks
 * map it with rcu_peer_region_open.  (Exchange regions are the one place the library allocates on the caller's behalf:
 * cudaIpcGetMemHandle needs the base of a cudaMalloc allocation.) */
int rcu_peer_region_alloc(size_t bytes, int device, void** ptr, void* ipc_handle_out);
int rcu_peer_region_open(const void* ipc_handle, int device, void** ptr);
int rcu_peer_region_close(void* ptr);
int rcu_peer_region_free(void* ptr);
int rcu_aggregate_finish_peer(void* const* region_ptrs, int n_ranks, int rank, uint32_t epoch, int total_samples,
                              int64_t n_images, int64_t hw, int has_mi, int has_var, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RCU_B200_H */
