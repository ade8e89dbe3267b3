/* sib200.h — C ABI of libsib200.so, the sm_100a (B200) kernels behind sota_imagenet_b200.
 *
 * The reference (bonlime/sota_imagenet) ships no native code: every kernel of its training
 * step is reached through torch (cuDNN / cuBLAS / ATen), NVIDIA DALI or pytorch_tools.  This
 * header is therefore the boundary a maintainer would bind *instead of* those library calls;
 * each entry point cites the reference call site whose kernel it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - activations are NHWC bf16 (channel count a multiple of 8, 16-byte aligned base);
 *   - conv filters are [Cout][R][S][Cin] bf16 ("KRSC" = torch channels_last memory format);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value 0 = success; otherwise sib_last_error() describes the failure.  Nothing
 *     falls back to the CPU: without an sm_100a device every compute call fails.
 */
#ifndef SIB200_H_
#define SIB200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define SIB_ABI_VERSION 2

/* activation codes (pytorch_tools ABN `activation=`; BResNet50_encoder.yaml:49 leaky_relu) */
#define SIB_ACT_NONE 0
#define SIB_ACT_RELU 1
#define SIB_ACT_LEAKY 2

/* margin kinds of the fused head */
#define SIB_MARGIN_NONE 0
#define SIB_MARGIN_ARC 1 /* AdditiveAngularMarginLoss, angular_losses.py:128-146 */
#define SIB_MARGIN_COS 2 /* CosFace, angular_losses.py:186-198 and :332-333      */
#define SIB_MARGIN_ARC_PURE 3 /* cos(theta + m) without fallback, angular_losses.py:78-83 */
#define SIB_MARGIN_ARCCOS 4 /* z = -(acos(clamp(cos)) + m[target]) * s: ArcCosSoftmax :572-576, AdaCos arc_logits :323-330 */

/* conv flags */
#define SIB_FLAG_FORCE_IM2COL 1 /* use the im2col TMA path even for plain 1x1 (testing) */
#define SIB_FLAG_TILE_N128 2    /* cap the N tile at 128 columns (tuning / testing)       */
#define SIB_FLAG_NO_2CTA 4      /* never use the cta_group::2 kernel (tuning / testing)   */
#define SIB_FLAG_STATS_ZEROED 8 /* `stats` / `sums` already hold zeros: skip the memset    */
#define SIB_FLAG_FORCE_HALO 16  /* 3x3/s1/p1: use the halo-reuse kernel whenever it applies */
#define SIB_FLAG_NO_HALO 32     /* never use the halo-reuse kernel (tuning / testing)       */
#define SIB_ACT_FLAG_PREZEROED 0x100 /* or-ed into sib_bn_bwd_reduce's `act`: `sums` already zero */

const char* sib_last_error(void);
/* CRC-32C of a HOST buffer (TFRecord framing, records.py); result in *out_host. */
int sib_crc32c_host(const void* data_host, unsigned long long n, unsigned int* out_host);
int sib_abi_version(void);
int sib_device_check(void); /* 0 iff the current device is sm_100 */

/* ---- convolution: replaces F.conv2d fwd / autograd dgrad / wgrad (cuDNN) reached from
 *      pytorch_tools.models.resnet50, reference train.py:64; in-repo call model.py:106 ---- */

/* y[N][OH][OW][K] = conv(x[N][H][W][C], w[K][R][S][C]); optional bias[K];
 * stats (optional, [2][K] fp32) receives per-channel sum and sum-of-squares of y (BN fusion). */
int sib_conv2d_fprop(const void* x, const void* w, void* y, int N, int H, int W, int C, int K,
                     int R, int S, int stride, int pad_h, int pad_w, int OH, int OW,
                     const float* bias, float* stats, int flags, void* stream);

/* y = conv(act(BN(x)), w): BatchNorm (+ activation) of the PRODUCER layer fused into the conv's
 * operand prologue -- x is the raw (pre-BN) output of the previous conv and the normalised
 * activation is never written to memory (north_star (2); replaces the cuDNN BatchNorm forward +
 * ReLU that pytorch_tools' ABN runs between two convs, reference train.py:64,76).
 * Training (bn_stats != NULL: [2][C] sum / sum of squares of x, already all-reduced under SyncBN):
 * the kernel also finalises that BatchNorm -- mean_invstd [2][C], scale_shift [2][C] are written,
 * running_mean / running_var updated with `momentum` (nn.BatchNorm2d semantics, `count` = elements
 * per channel).  Eval (bn_stats == NULL): scale_shift is an input.  Zero padding applies to the
 * ACTIVATED tensor (padding taps contribute 0).  C % 64 == 0, C <= 512. */
int sib_conv2d_fprop_bnact(const void* x, const void* w, void* y, int N, int H, int W, int C, int K,
                           int R, int S, int stride, int pad_h, int pad_w, int OH, int OW,
                           float* stats, int flags, const float* bn_stats, const float* gamma,
                           const float* beta, float* running_mean, float* running_var,
                           float* mean_invstd, float* scale_shift, double count, float eps,
                           float momentum, int act, float slope, void* stream);

/* dx[N][H][W][C] = dgrad(dy[N][OH][OW][K]) [+ residual]; w_dgrad is the tap-flipped transposed
 * filter [C][R][S][K] produced by sib_pack_dgrad_weights.  `residual` (optional, geometry of dx,
 * may alias dx) is added in the epilogue (identity-shortcut gradient).  stride > 1 needs
 * `workspace`: RxS -> N*((OH-1)s+1)*((OW-1)s+1)*K bf16 (zero-inserted dy); 1x1 -> N*OH*OW*C bf16
 * (compact result, then ACCUMULATED into dx at every stride-th pixel; residual must be NULL). */
int sib_conv2d_dgrad(const void* dy, const void* w_dgrad, void* dx, const void* residual,
                     void* workspace, int N, int H, int W, int C, int K, int R, int S, int stride,
                     int pad, int flags, void* stream);
/* dgrad with the BatchNorm-backward reduction of the BN this gradient flows into fused in the
 * epilogue (replaces autograd's separate ReLU-backward + BN-backward reduction passes):
 *   g  = (dgrad(dy) [+ residual]) * act'(z)   -> dx (bf16, already masked)
 *   z  = fmaf(mask_src, scale, shift) with mask_ss = [2][C] (the forward's own arithmetic on the
 *        BN input), or z = mask_src when mask_ss is NULL (mask_src = stored post-activation output)
 *   sums[0][C] = sum g, sums[1][C] = sum g * xhat, xhat = (xhat_src - mean) * invstd
 *   (xhat_src NULL -> mask_src).  Not available for strided 1x1 filters. */
int sib_conv2d_dgrad_bnbwd(const void* dy, const void* w_dgrad, void* dx, const void* residual,
                           void* workspace, int N, int H, int W, int C, int K, int R, int S,
                           int stride, int pad, int flags, const void* mask_src,
                           const float* mask_ss, const void* xhat_src, const float* mean_invstd,
                           int act, float slope, float* sums, void* stream);
/* 3x3 / stride 2 / pad 1 dgrad by row parity (no zero insertion): two stride-1 GEMMs over dy with
 * the sub-filters of sib_pack_dgrad_s2, each storing one row parity of dx[N][H][W][C] in place.
 * H, W even, W/2 <= 32.  mask_src non-NULL fuses the BatchNorm-backward reduction of the BN whose
 * INPUT is mask_src (geometry of dx), as in sib_conv2d_dgrad_bnbwd with mask_ss. */
int sib_conv2d_dgrad_s2(const void* dy, const void* w_sub0, const void* w_sub1, void* dx, int N,
                        int H, int W, int C, int K, int flags, const void* mask_src,
                        const float* mask_ss, const float* mean_invstd, int act, float slope,
                        float* sums, void* stream);
/* w_dgrad [C][3][3][K] (sib_pack_dgrad_weights) -> w_sub0 [2C][1][2][K], w_sub1 [2C][2][2][K] */
int sib_pack_dgrad_s2(const void* w_dgrad, void* w_sub0, void* w_sub1, int C, int K, void* stream);
int sib_scatter_add_strided(const void* src, void* dst, int N, int OH, int OW, int C, int H, int W,
                            int stride, void* stream);
/* dw[K][R][S][C] (fp32) += wgrad(x, dy).  dw must be initialised (zero_grad). */
int sib_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int K,
                     int R, int S, int stride, int pad_h, int pad_w, int OH, int OW, int flags,
                     void* stream);
int sib_upsample_zero(const void* dy, void* up, int N, int OH, int OW, int C, int UH, int UW,
                      int stride, void* stream);

/* ---- BatchNorm (+ReLU, +residual): replaces nn.BatchNorm2d / pytorch_tools ABN kernels
 *      (cuDNN BN + ATen relu/add), momentum patched by reference train.py:76 ---- */
int sib_bn_stats(const void* x, long M, int C, float* stats, void* stream);
int sib_bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float* mean_invstd, float* scale_shift, int C,
                    double count, float eps, float momentum, void* stream);
int sib_bn_eval_scale(const float* gamma, const float* beta, const float* running_mean,
                      const float* running_var, float* scale_shift, int C, float eps, void* stream);
/* y = act(x*scale + shift [+ res] [res*scale2 + shift2]) */
int sib_bn_apply(const void* x, const float* scale_shift, const void* res,
                 const float* scale_shift2, void* y, long M, int C, int act, float slope,
                 void* stream);
/* finalize + apply fused (training): y = act(BN1(x) [+ res | + BN2(res)]); publishes
 * mean_invstd / scale_shift [2][C] for the backward pass and updates the running statistics */
int sib_bn_finalize_apply(const void* x, const float* stats, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float* mean_invstd,
                          float* scale_shift, const void* res, const float* stats2,
                          const float* gamma2, const float* beta2, float* running_mean2,
                          float* running_var2, float* mean_invstd2, float* scale_shift2, void* y,
                          long M, int C, double count, float eps, float momentum, int act,
                          float slope, void* stream);
/* sums[0..1][C] = (sum g, sum g*xhat), g = dy * act'(.); with x2: sums[2..3] for the 2nd BN.
 * The activation mask is taken from `out` (stored forward output) when given, otherwise it is
 * recomputed bit-exactly from x and the forward's scale/shift (`mask_ss`, [2][C]). */
int sib_bn_bwd_reduce(const void* dy, const void* out, const float* mask_ss, const void* x,
                      const float* mean_invstd, const void* x2, const float* mean_invstd2, long M,
                      int C, int act, float slope, float* sums, void* stream);
/* dx = BN backward (dx2 for a second BN sharing g); dgamma / dbeta (optional) += sums * pgrad_scale.
 * pgrad_scale = 1 normally; 1 / world under SyncBN, where `sums` hold totals over ALL ranks and the
 * data-parallel wrapper averages parameter gradients over the ranks afterwards (torch's SyncBatchNorm
 * uses the rank-local sums for the same reason, torch/nn/modules/_functions.py:150-160). */
int sib_bn_bwd_apply(const void* dy, const void* out, const float* mask_ss, const void* x,
                     const float* mean_invstd, const float* gamma, const float* sums,
                     const void* x2, const float* mean_invstd2, const float* gamma2, void* dx,
                     void* dx2, void* gout, float* dgamma, float* dbeta, float* dgamma2,
                     float* dbeta2, long M, int C, double count, int act, float slope,
                     float pgrad_scale, void* stream);
/* sib_bn_bwd_apply for a plain BatchNorm (+ activation) that ALSO re-materialises the forward
 * activation a_out = fwd_act(x * act_ss.scale + act_ss.shift): with the fused conv prologue
 * (sib_conv2d_fprop_bnact) the activation is never stored in forward, and the weight gradient of
 * the consumer conv needs it in backward.  x is in registers anyway: +2 B/element written. */
int sib_bn_bwd_apply_remat(const void* dy, const float* mask_ss, const void* x,
                           const float* mean_invstd, const float* gamma, const float* sums, void* dx,
                           float* dgamma, float* dbeta, const float* act_ss, int fwd_act,
                           float fwd_slope, void* a_out, long M, int C, double count, int act,
                           float slope, float pgrad_scale, void* stream);

/* (dgamma/dbeta[/2], optional: the affine-parameter gradients are ACCUMULATED into them) */
int sib_bn_param_grad(const float* sums, float* dgamma, float* dbeta, int C, int accumulate,
                      void* stream);

/* ---- pooling (stem max-pool, head global average pool) ---- */
int sib_maxpool3x3s2_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C,
                         void* stream);
int sib_maxpool3x3s2_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C,
                         void* stream);
/* stem: BatchNorm finalize (from the conv epilogue's statistics) + apply + activation + 3x3/s2
 * max pool in one pass; the normalised activation is never materialised.  y / idx are
 * bit-identical to sib_bn_finalize_apply followed by sib_maxpool3x3s2_fwd. */
int sib_bn_act_maxpool3x3s2_fwd(const void* x, const float* stats, const float* gamma,
                                const float* beta, float* running_mean, float* running_var,
                                float* mean_invstd, float* scale_shift, void* y, void* idx, int N,
                                int H, int W, int C, double count, float eps, float momentum,
                                int act, float slope, void* stream);
int sib_gap_fwd(const void* x, void* y, int N, int HW, int C, void* stream);
int sib_gap_bwd(const void* dy, void* dx, int N, int HW, int C, void* stream);

/* ---- BResNet-50 extras (BResNet50_encoder.yaml:44-51): anti-alias BlurPool, AvgPool shortcut,
 *      ECA attention, per-(sample,channel) scaling (ECA / drop-connect), add+act ---- */
int sib_blurpool_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream);
int sib_blurpool_bwd(const void* dy, void* dx, int N, int H, int W, int C, void* stream);
int sib_avgpool2_fwd(const void* x, void* y, int N, int H, int W, int C, void* stream);
int sib_avgpool2_bwd(const void* dy, void* dx, int N, int H, int W, int C, void* stream);
int sib_maxpool3x3s1_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C,
                         void* stream);
int sib_maxpool3x3s1_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C,
                         void* stream);
/* out[n][c] = scale * sum_hw a[n][hw][c] * (b ? b[n][hw][c] : 1) */
int sib_chan_reduce(const void* a, const void* b, float* out, int N, int HW, int C, float scale,
                    void* stream);
/* y[n][hw][c] = x[n][hw][c] * mul[n][c] (+ add[n][c]) */
int sib_scale_nc(const void* x, const float* mul, const float* add, void* y, int N, int HW, int C,
                 void* stream);
/* y = act(x * mul[n][c] (+ add[n][c]) + res): BResNet block tail (ECA gate, drop-connect, shortcut,
 * activation).  With `add` (optional) x may be the RAW output of the block's last conv: mul / add
 * then carry its BatchNorm scale / shift times the gate, and the normalised tensor is never written. */
int sib_scale_add_act(const void* x, const float* mul, const float* add, const void* res, void* y,
                      int N, int HW, int C, int act, float slope, void* stream);
/* Block-tail backward in one pass: g = dy * act'(y) (written; also the shortcut gradient) and per
 * (sample, channel) s1[N][C] = sum_hw g, s2[N][C] = sum_hw g * x.  Replaces act_bwd + chan_reduce(g, x)
 * (the ECA gate's backward input, reference model graph: pytorch_tools ECA inside Bottleneck). */
int sib_act_bwd_reduce(const void* dy, const void* y, const void* x, void* g, float* s1, float* s2,
                       int N, int HW, int C, int act, float slope, void* stream);
/* The [N][C]-sized algebra of the fused BResNet block tail, one launch per direction.
 * forward : p = pc * scale + shift, gate = sigmoid(conv1d_k3(p)), gate_k = gate * keep[n] (keep optional),
 *           mul = gate_k * scale, add = gate_k * shift  (inputs of sib_scale_add_act).
 * backward: from s1 / s2 of sib_act_bwd_reduce: the ECA filter gradient dw[3] (+=), the pooled-path term
 *           add_nc = dp / HW and the BatchNorm-backward sums[2][C] (+=) of d = g * gate_k + add_nc. */
int sib_eca_tail_fwd(const float* pc, const float* scale_shift, const float* w, const float* keep, float* p,
                     float* gate, float* gate_k, float* mul, float* add, int N, int C, void* stream);
int sib_eca_tail_bwd(const float* s1, const float* s2, const float* scale_shift, const float* mean_invstd,
                     const float* pc, const float* p, const float* gate, const float* gate_k,
                     const float* keep, const float* w, float* add_nc, float* sums, float* dw, int N, int C,
                     float hw, void* stream);
/* sib_bn_bwd_apply for an activation-free BatchNorm whose incoming gradient is
 * dy * dy_mul[n][c] + dy_add[n][c]: the gate scale of the block tail folded into the BN backward. */
int sib_bn_bwd_apply_scaled(const void* dy, const float* dy_mul, const float* dy_add, const void* x,
                            const float* mean_invstd, const float* gamma, const float* sums, void* dx,
                            float* dgamma, float* dbeta, int N, int HW, int C, double count,
                            float pgrad_scale, void* stream);
int sib_eca_gate_fwd(const float* p, const float* w, float* s, int N, int C, void* stream);
int sib_eca_gate_bwd(const float* ds, const float* s, const float* p, const float* w, float* dp,
                     float* dw, int N, int C, void* stream);
int sib_add_act(const void* a, const void* b, void* y, long n, int act, float slope, void* stream);
int sib_act_bwd(const void* dy, const void* y, void* g, long n, int act, float slope, void* stream);

/* ---- heads: pytorch_tools.losses.smooth.CrossEntropyLoss (arg_parser.py:140-142),
 *      angular_losses.py AdditiveAngularMarginLoss / CosFace / SphereLinearLayer ---- */
int sib_ce_fwd_bwd(const void* logits, int logits_fp32, const long* labels,
                   const float* dense_targets, int B, int C, int ld, float smoothing,
                   float temperature, int margin_kind, float s, float m, float* loss_rows,
                   float* loss_mean, void* dlogits, float grad_scale, void* stream);
int sib_sphere_linear_fwd(const float* x, const float* w, float* cosv, float* xn, float* wn,
                          float* xnorm, float* wnorm, int B, int C, int D, int normalize_x,
                          void* stream);
/* scratch: (B + C) * D floats (d(xn) and d(wn) come out of ONE tensor-core launch) */
int sib_sphere_linear_bwd(const float* dcos, const float* x_or_xn, const float* wn,
                          const float* xnorm, const float* wnorm, float* dx, float* dw,
                          float* scratch, int B, int C, int D, int normalize_x, void* stream);

/* ---- data-parallel: one-shot all-reduce of small fp32 vectors over NVLink peer memory (SyncBN
 *      statistics; replaces the per-layer NCCL all-reduce of SyncBatchNorm under DDP, reference
 *      train.py:113-114).  Every rank allocates a mailbox + flag buffer with sib_ipc_alloc, the
 *      64-byte handles are exchanged by the host, peers map them with sib_ipc_open.
 *      table_dev: device struct { u64* mailbox[8]; void* unused[8]; } indexed by rank.
 *      Mailbox layout [2 parities][parity_stride 8-byte words] ({value, epoch} per float); call
 *      `slot` uses words [slot_off, slot_off + world*n); epochs_dev [num_slots] u32 (local).
 *      All ranks must issue the same sequence of calls.  data is summed in place, bitwise
 *      identical on every rank. ---- */
int sib_ipc_alloc(unsigned long long bytes, void** ptr, unsigned char* handle64);
int sib_ipc_open(const unsigned char* handle64, void** ptr);
int sib_ipc_close(void* ptr);
int sib_ipc_free(void* ptr);
int sib_peer_allreduce(float* data, int n, const void* table_dev, long slot_off, long parity_stride,
                       int slot, int num_slots, void* epochs_dev, int rank, int world, void* stream);

/* ---- optimizer: torch.optim._multi_tensor.SGD (arg_parser.py:136-138), ModelEma
 *      (train.py:112), weight standardisation (train.py:66-67, model.py:91-100) ---- */
/* segs_dev: nseg x {long end; float lr, weight_decay, momentum, dampening; int nesterov, pad} */
int sib_sgd_step(float* params, const float* grads, float* momentum_buf, void* params_bf16,
                 float* ema, float ema_decay, const void* segs_dev, int nseg, long n,
                 int first_step, void* stream);
/* MyNovograd (reference sota_imagenet/optimizers.py:35-161): per-tensor (or per-output-unit,
 * unitwise_norm=True) running norm of the WEIGHTS (:133-140), first moment of the gradient
 * (:143-147), p = (p - lr*ema_grad/(sqrt(ema_norm)+eps)) * (1 - lr*wd) (:155-158).
 * table_dev: ntensors x {long begin, end; int unit_len, ngroups, norm_base; float lr, decay,
 * beta1, one_minus_beta1, beta2, one_minus_beta2; int pad}; group_* are [total groups] fp32. */
int sib_novograd_step(float* params, const float* grads, float* ema_grad, void* params_bf16,
                      float* ema, float ema_decay, const void* table_dev, int ntensors,
                      float* group_sumsq, float* group_ema_norm, float* group_denom, long n,
                      float eps, int unitwise, void* stream);
int sib_cast_bf16(const float* src, void* dst, long n, void* stream);
/* table_dev: nent x {long src_off, dst_off; int K, RS, C, block_begin} */
int sib_pack_dgrad_weights(const void* w_bf16, void* w_dgrad, const void* table_dev, int nent,
                           int total_blocks, void* stream);
int sib_weight_standardize(const float* w, const float* gain, void* out_bf16, float* mean_invstd,
                           int out_channels, int fan, float eps, void* stream);
int sib_weight_standardize_bwd(const float* w, const float* gain, const float* mean_invstd,
                               const float* g, float* dw, int out_channels, int fan, void* stream);
/* The same two maps for EVERY standardised filter of a parameter arena in one launch each (52 tensors
 * in BResNet-50): params / shadow / grads are the flat arenas, table_dev an array of n records
 * {long off; long mi_off; int K; int fan; int block_base; int pad} (one block per output channel),
 * mean_invstd the flat [sum K][2] buffer; the backward rewrites `grads` in place. */
int sib_weight_standardize_batch(const float* params, void* shadow_bf16, float* mean_invstd,
                                 const void* table_dev, int n, int total_blocks, float eps, void* stream);
int sib_weight_standardize_bwd_batch(const float* params, const float* mean_invstd, float* grads,
                                     const void* table_dev, int n, int total_blocks, void* stream);

/* ---- data: DALI train pipeline (dali_dataloader.py:65-74,113-123) on synthetic uint8 ---- */
int sib_rrc_boxes(int* boxes_dev, int B, int H, int W, double min_area, double max_area,
                  unsigned long long seed, unsigned long long first_sample, int do_flip,
                  void* stream);
void sib_rrc_box_host(int H, int W, double min_area, double max_area, unsigned long long seed,
                      unsigned long long sample, int* box5_host);
/* out_mode 0: NHWC bf16, 4 channels (4th zero); 1: NCHW fp32 (the reference layout) */
int sib_augment(const void* src_u8, const int* boxes_dev, void* out, int B, int SH, int SW, int S,
                float mean, float std, int out_mode, void* stream);
/* validation pipeline (dali_dataloader.py:146-160): resize the shorter side to `resize_shorter`
 * (triangular), centre crop S x S, normalise; out_mode as sib_augment.  g4_host receives
 * {resized H, resized W, crop origin y, crop origin x}. */
int sib_val_transform(const void* src_u8, void* out, int B, int SH, int SW, int S,
                      int resize_shorter, float mean, float std, int out_mode, void* stream);
void sib_val_geometry_host(int SH, int SW, int S, int resize_shorter, int* g4_host);
/* ---- hybrid JPEG decode (reference dali_dataloader.py:65-72,140-145: fn.decoders.image*(device="mixed"),
 *      i.e. nvJPEG's hybrid back end): Huffman decoding on the host, dequantisation + IDCT + chroma
 *      upsampling + YCbCr->RGB on the device, written straight into the packed ragged uint8 buffer.
 *      Integer arithmetic of the IJG / libjpeg-turbo decoder (islow IDCT, fancy upsampling):
 *      bit-identical to PIL on the same stream. ---- */
#define SIB_JPEG_OK 0
#define SIB_JPEG_NOT_JPEG 1
#define SIB_JPEG_CORRUPT 2
#define SIB_JPEG_UNSUPPORTED_PROCESS 3     /* progressive / arithmetic / lossless */
#define SIB_JPEG_UNSUPPORTED_PRECISION 4   /* not 8-bit */
#define SIB_JPEG_UNSUPPORTED_COLORSPACE 5  /* CMYK, Adobe RGB, ... */
#define SIB_JPEG_UNSUPPORTED_SAMPLING 6    /* anything but 4:4:4 / 4:2:2 / 4:2:0 */
#define SIB_JPEG_UNSUPPORTED_SCANS 7       /* more than one scan */
typedef struct sib_jpeg_info {
  long coef_count;               /* int16 coefficients of all components (multiple of 64) */
  int status;                    /* SIB_JPEG_*: anything but OK means "decode this one elsewhere" */
  int width, height, ncomp;      /* ncomp 1 (grey) or 3 (YCbCr) */
  int hs[3], vs[3];              /* sampling factors per component */
  int hmax, vmax;
  int mcus_x, mcus_y;
  int blocks_w[3], blocks_h[3];  /* 8x8 blocks per component, padded to whole MCUs */
  int restart_interval;
  unsigned short quant[3][64];   /* quantisation table per component, natural (row-major) order */
} sib_jpeg_info;
/* one record per image of a batch (device array): where its coefficients, scratch planes and output live */
typedef struct sib_jpeg_image {
  long coef_off[3];              /* first coefficient of component c in the batch buffer (int16 elements) */
  long plane_off[3];             /* first byte of component c's plane [blocks_h*8][blocks_w*8] in the scratch */
  long out_off;                  /* first byte of the RGB image [height][width][3] in the output buffer */
  int width, height, ncomp;
  int hmax, vmax;
  int blocks_w[3], blocks_h[3];
  int mcu_rows;                  /* > 0: only the first mcu_rows MCU rows were decoded (row-limited decode) */
  unsigned short quant[3][64];
} sib_jpeg_image;
/* host, thread-safe, no CUDA: */
int sib_jpeg_parse(const unsigned char* data, long size, sib_jpeg_info* info);
/* coef: info.coef_count int16 values, component after component, blocks in raster order, natural order
 * inside a block, not dequantised */
int sib_jpeg_decode_coefficients(const unsigned char* data, long size, short* coef);
/* the same, stopping after the first mcu_rows rows of MCUs (0 = all): what a random crop that ends above the
 * bottom of the image needs (ROI decoding of fn.decoders.image_random_crop, dali_dataloader.py:65-72) */
int sib_jpeg_decode_coefficients_rows(const unsigned char* data, long size, short* coef, int mcu_rows);
/* device: max_blocks = largest sum of blocks over the components of one image, max_pixels = largest
 * width*height; planes_dev = scratch for all component planes of the batch */
int sib_jpeg_idct_rgb(const short* coef_dev, const sib_jpeg_image* images_dev, int B, int max_blocks,
                      int max_pixels, unsigned char* planes_dev, unsigned char* out_dev, void* stream);

/* ragged batches of real images (records.pack_batch): packed uint8 buffer + per-image byte offsets
 * [B] + {H, W} [B][2]; same arithmetic as sib_rrc_boxes / sib_augment / sib_val_transform with the
 * image base and extent looked up per sample (fn.decoders.image_random_crop + fn.resize on images of
 * different sizes, dali_dataloader.py:65-74, :144-148). */
int sib_rrc_boxes_ragged(int* boxes_dev, const int* dims_dev, int B, double min_area, double max_area,
                         unsigned long long seed, unsigned long long first_sample, int do_flip,
                         void* stream);
int sib_augment_ragged(const void* packed_u8, const long* offsets_dev, const int* dims_dev,
                       const int* boxes_dev, void* out, int B, int S, float mean, float std,
                       int out_mode, void* stream);
int sib_val_transform_ragged(const void* packed_u8, const long* offsets_dev, const int* dims_dev,
                             void* out, int B, int S, int resize_shorter, float mean, float std,
                             int out_mode, void* stream);
int sib_one_hot(const long* labels, float* out, int B, int C, void* stream);
/* Per-sample photometric augmentations of the resident batch (the output of sib_augment): replaces
 * fn.color_twist / fn.hsv(saturation=0) / fn.erase of the reference's train_pipeline
 * (dali_dataloader.py:86-111).  In place.  params[N][16 + 4 * nboxes] fp32 per sample:
 *   [0..8] 3x3 colour matrix, [9..11] offset (both already expressed in NORMALISED space),
 *   [12] lower / [13] upper clamp (the normalised images of 0 and 255), [14] grayscale flag,
 *   [15] erase fill value, then nboxes x (h1, w1, h2, w2) erase boxes in UN-mirrored pixel
 *   coordinates (an empty box erases nothing).  crop_boxes (optional): the int32 [N][5] boxes of
 *   sib_rrc_boxes; samples with the flip flag get their erase boxes mirrored.
 * layout 0: bf16 NHWC, 4 channels (4th stays 0); layout 1: fp32 NCHW, 3 channels. */
int sib_pixel_ops(void* x, const float* params, const int* crop_boxes, int N, int H, int W, int layout,
                  int nboxes, void* stream);
/* fn.gaussian_blur(window_size=11, sigma) of dali_dataloader.py:81-83 on the resident batch, out of
 * place; sigma[N] fp32 per sample, sigma <= 0 copies the sample through.  Border: reflect-101. */
int sib_gaussian_blur(const void* x, void* y, const float* sigma, int N, int H, int W, int layout,
                      void* stream);

/* batch-level mixing on the resident batch: pt_clb.Mixup / pt_clb.Cutmix as combined by
 * CutmixMixup (sota_imagenet/callbacks.py:232-247).  layout 0: NHWC bf16 [N][H][W][C], 1: NCHW
 * fp32; mode 0: out = lam*x + one_minus_lam*prev[perm[n]]; mode 1: out = prev[perm[n]] inside
 * rows [h1,h2) x columns [w1,w2), x elsewhere.  prev is [N][PH][PW][C] / [N][C][PH][PW]. */
int sib_mix_batch(const void* x, const void* prev, const int* perm_dev, void* out, int N, int H,
                  int W, int C, int PH, int PW, int layout, int mode, float lam,
                  float one_minus_lam, int h1, int w1, int h2, int w2, void* stream);
/* out[n][c] = w_self*t[n][c] + w_prev*prev_t[perm[n]][c] (soft targets of the mixed batch) */
int sib_mix_targets(const float* t, const float* prev_t, const int* perm_dev, float* out, int N,
                    int C, float w_self, float w_prev, void* stream);
/* stem packing (see csrc/augment.cu): src_mode 0 = NHWC4 bf16, 1 = NCHW fp32 */
int sib_stem_pack(const void* src, void* xq, int N, int H, int W, int KW, int pad_w, int src_mode,
                  void* stream);
int sib_stem_pack_weight(const float* w_oihw, void* wq, int K, int KH, int KW, int NA, int off,
                         void* stream);
int sib_stem_unpack_wgrad(const float* dwq, float* dw_oihw, int K, int KH, int KW, int NA, int off,
                          int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIB200_H_ */
