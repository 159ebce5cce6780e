/*
 * dfu_b200.h -- C ABI of libdfu_b200.so: the sm_100a kernels behind the late-fusion DFU classifier step.
 *
 * The reference (ShreenathKR2000/DFU-Multimodal) has no FFI of its own: its hot path is the PyTorch
 * nn.Module protocol (`model(rgb, thermal)`, `loss.backward()`, `optimizer.step()`;
 * notebooks/train_multimodal_fusion.py:368-388) dispatching into ATen/cuDNN/cuBLAS.  Every entry point
 * below replaces the ATen call named in its comment.  Conventions:
 *   - plain pointers + sizes, no torch types; the CALLER owns every buffer, workspace included;
 *   - all pointers are device pointers unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - no hidden allocation, no host synchronisation, re-entrant across streams;
 *   - returns 0 on success; nonzero -> dfu_last_error() holds the reason (host side raises RuntimeError);
 *   - there is no CPU fallback: without an sm_100 device every compute entry point fails.
 * Activations are bf16, NHWC for images ([B*H*W, C] row-major matrices) and [B*197, 768] for tokens;
 * parameters, statistics, gradients of parameters and optimizer state are fp32.
 */
#ifndef DFU_B200_H
#define DFU_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* dfu_last_error(void);
int dfu_version(void);
/* 0 if the current CUDA device is sm_100 class and the driver TMA entry point resolved. */
int dfu_device_ok(void);
int dfu_sm_count(void);
/* Grid-size budget (in SMs) for the kernels launched next (process-wide); 0 = the whole device.  Lets two streams
 * run side by side on disjoint SM shares (fusion model: ViT branch next to the ResNet branch). */
int dfu_set_sm_budget(int32_t n_sms);

/* ---------------------------------------------------------------------------------------------
 * GEMM: D[z][m][n] = act(alpha * sum_k A[z][m][k] * B[z][n][k] + bias[n]) + res[z][m][n]
 * Replaces aten::addmm / aten::mm / aten::bmm (nn.Linear in timm ViT blocks and the fusion head,
 * train_multimodal_fusion.py:305-313) and cudnn convolution fwd/dgrad/wgrad after im2col
 * (torchvision ResNet, train_multimodal_fusion.py:294).
 *   a_major/b_major: 0 = K-major (A is [M,K] row-major with row stride lda; B is [N,K] row-major),
 *                    1 = MN-major (A stored [K,M] row-major with row stride lda; B stored [K,N]).
 *   z = z1*nb0 + z0, element strides s?0 / s?1 per batch dim.  All strides in ELEMENTS.
 *   d_dtype: 0 = bf16, 1 = fp32.  accumulate=1 (fp32 only): D += result (atomic, enables split-K).
 *   aux (bf16, same geometry as D) receives the pre-activation value when non-null.
 *   act: 0 none, 1 relu, 6 = relu(alpha*acc + bias + res) (ReLU AFTER the residual add: torchvision Bottleneck tail with the
 *   eval-mode BatchNorm folded into weights and bias), 2 gelu(erf), 3 = D = (alpha*acc + bias) * gelu'(res) (res holds the pre-activation; fuses the
 *   GELU backward into the dgrad GEMM); 4 = gelu(erf) with aux receiving gelu'(pre-activation) instead of the
 *   pre-activation (aux required); 5 = D = (alpha*acc + bias) * res (res = the derivative stored by a forward act=4).
 *   impl: 0 = tcgen05/TMEM/TMA kernel, 1 = SIMT fp32 check kernel.
 *   colstats (fp32, caller zeroes; tcgen05 kernel only, bf16 output, N x mode <= 4096, unbatched): += per-column sums of the
 *   stored (bf16-rounded) D over the M rows
 *   (mode 1: sums[N]) and sums of squares (mode 2: sums[N], sumsq[N]) -- BatchNorm batch statistics straight from the
 *   convolution GEMM, bias gradients straight from the dgrad GEMM.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t M, N, K, nb0, nb1;
    const void* A; int32_t a_major; int64_t lda, sa0, sa1;
    const void* B; int32_t b_major; int64_t ldb, sb0, sb1;
    void* D; int32_t d_dtype; int64_t ldd, sd0, sd1;
    int32_t accumulate;
    void* aux;
    const float* bias;
    const void* res; int64_t ldr, sr0, sr1;
    float alpha;
    int32_t act;
    int32_t impl;
    int32_t max_ctas; /* 0 = one per SM */
    float* colstats;
    int32_t colstats_mode;
} dfu_gemm_args;
int dfu_gemm(const dfu_gemm_args* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on NHWC bf16 (channel counts multiples of 64; stride 1 or 2 -- at stride 2 the gathered
 * boxes of modes 1 and 3 walk the input with TMA element strides and mode 2 runs one launch per parity class of the dx
 * pixels, each writing a strided view of dx): the same tcgen05 kernel, with the
 * operand tiles gathered by TMA from the activation tensor as boxes of pixels shifted by the filter tap (padding =
 * TMA out-of-bounds zero fill).  No im2col / col2im matrix exists.  Replaces cudnn convolution forward / backward
 * data / backward filter for the 3x3 and strided 1x1 convolutions of torchvision's Bottleneck
 * (train_multimodal_fusion.py:294).
 *   mode 1: out = y  [B,Ho,Wo,Cout] bf16  from x [B,H,W,C] and w [Cout, ldw] (columns ordered (r, s, c))
 *   mode 2: out = dx [B,H,W,C]      bf16  from dy [B,Ho,Wo,Cout] and w
 *   mode 3: out = dW [Cout, ldo]    fp32  from dy and x; accumulate=1: dW += (caller zeroes; enables split-K)
 * Returns 3 (unsupported) for other strides / channel counts: the caller then uses dfu_im2col_nhwc + dfu_gemm.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t mode;
    int32_t B, H, W, C;
    int32_t Cout, R, S, stride, pad;
    const void* x;
    const void* w; int64_t ldw;
    const void* dy;
    void* out; int64_t ldo;
    int32_t accumulate;
    int32_t max_ctas;
    float* colstats;   /* mode 1 only, may be null: [2*Cout] fp32 (caller zeroes) += per-channel sum and sum of squares of y */
    const float* bias; /* mode 1 only, may be null: y = act(conv + bias[Cout]) -- eval-mode BatchNorm folded into the convolution */
    int32_t act;       /* mode 1 only: 0 none, 1 ReLU */
} dfu_conv_args;
int dfu_conv_gemm(const dfu_conv_args* a, void* stream);

/* Eval-mode BatchNorm folded into the convolution that feeds it (extended_metrics.py:630-640 runs the model in eval()):
 * w_out[co, k] = bf16(w[co, k] * gamma[co] / sqrt(var[co] + eps)) (row stride ldo >= K, padding zeroed),
 * bias_out[co] = beta[co] - mean[co] * gamma[co] / sqrt(var[co] + eps).  Replaces aten::native_batch_norm(training=False)
 * after aten::convolution by one GEMM epilogue (dfu_gemm bias/act 1|6, dfu_conv_gemm bias/act). */
int dfu_fold_bn_weights(const float* w, int32_t Cout, int32_t K, const float* gamma, const float* beta, const float* mean,
                        const float* var, float eps, void* w_out, int64_t ldo, float* bias_out, void* stream);

/* ------------------------------- layout / preprocessing ------------------------------------ */
/* a9: ToTensor+Normalize (train_multimodal_fusion.py:184-188,201-205) on uint8 HWC images.
 * out NHWC bf16: out[b,h,w,c] = (in/255 - mean[c]) / std[c]. */
int dfu_preprocess_u8_nhwc(const uint8_t* in, void* out, int64_t n_pixels, const float* mean3_host,
                           const float* std3_host, void* stream);
/* Same normalisation fused with ViT patch extraction: out[(b*196+p), c*256+i*16+j] (timm PatchEmbed order). */
int dfu_preprocess_u8_patches(const uint8_t* in, void* out, int32_t B, const float* mean3_host,
                              const float* std3_host, void* stream);
/* fp32 NCHW (the nn.Module contract) -> bf16 NHWC, and its adjoint (grad wrt the input image). */
int dfu_nchw_f32_to_nhwc_bf16(const float* in, void* out, int32_t B, int32_t C, int32_t H, int32_t W, void* stream);
int dfu_nhwc_bf16_to_nchw_f32(const void* in, float* out, int32_t B, int32_t C, int32_t H, int32_t W, void* stream);
/* fp32 NCHW 224x224x3 -> ViT patches bf16 [B*196,768], and adjoint. */
int dfu_patchify_f32(const float* in, void* out, int32_t B, void* stream);
int dfu_unpatchify_f32(const void* dpatches, float* dimg, int32_t B, void* stream);
/* im2col on NHWC bf16: col[(b,ho,wo)][(r*S+s)*C+c], row stride ldc (>= R*S*C, pad columns zeroed). */
int dfu_im2col_nhwc(const void* x, void* col, int32_t B, int32_t H, int32_t W, int32_t C, int32_t R, int32_t S,
                    int32_t stride, int32_t pad, int32_t Ho, int32_t Wo, int64_t ldc, void* stream);
/* adjoint of im2col (gather form, deterministic). */
int dfu_col2im_nhwc(const void* col, void* dx, int32_t B, int32_t H, int32_t W, int32_t C, int32_t R, int32_t S,
                    int32_t stride, int32_t pad, int32_t Ho, int32_t Wo, int64_t ldc, void* stream);

/* ------------------------------------ BatchNorm (a3) --------------------------------------- */
/* Per-channel sum / sum of squares of y[M,C] (bf16) accumulated into sums[2*C] (fp32, caller zeroes). */
int dfu_bn_stats(const void* y, float* sums, int64_t M, int32_t C, void* stream);
/* mean/invstd from sums; running stats update (momentum, unbiased var) as torch BatchNorm2d in train(). */
int dfu_bn_finalize(const float* sums, float* mean_invstd, float* running_mean, float* running_var,
                    int64_t* num_batches_tracked, int64_t M, int32_t C, float eps, float momentum, void* stream);
/* eval(): mean/invstd from running stats. */
int dfu_bn_eval_stats(const float* running_mean, const float* running_var, float* mean_invstd, int32_t C,
                      float eps, void* stream);
/* z = [relu]( gamma*(y-mean)*invstd + beta [+ res] ) */
/* zmask (may be null; u8 [M, C/8]): bit i of byte (m, c/8) = (z[m, c/8*8+i] > 0) -- the ReLU mask the backward pass needs
 * when the block has a residual input, 1 bit per element instead of re-reading z. */
int dfu_bn_apply(const void* y, const float* mean_invstd, const float* gamma, const float* beta, const void* res,
                 void* z, uint8_t* zmask, int64_t M, int32_t C, int32_t relu, void* stream);
/* train(): dfu_bn_finalize and dfu_bn_apply in one launch -- mean/invstd from the batch sums (stored to mean_invstd for the
 * backward pass), running statistics and num_batches_tracked updated, then z = [relu](gamma*xhat + beta [+ res]). */
int dfu_bn_apply_train(const void* y, const float* sums, const float* gamma, const float* beta, const void* res, void* z,
                       uint8_t* zmask, float* mean_invstd, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                       int64_t M, int32_t C, int32_t relu, float eps, float momentum, void* stream);
/* Backward pass 1: sums[0:C] += sum(dpre), sums[C:2C] += sum(dpre*xhat), dpre = dz*(z>0 if relu).
 * The ReLU mask comes from zmask (see dfu_bn_apply) when given, else from z; both may be null when the block had no residual input: the mask is then recomputed from y as gamma*xhat+beta > 0
 * (gamma, beta required), which saves reading z. */
int dfu_bn_bwd_reduce(const void* dz, const void* z, const uint8_t* zmask, const void* y, const float* mean_invstd,
                      const float* gamma, const float* beta, float* sums, int64_t M, int32_t C, int32_t relu, void* stream);
/* Backward pass 2: dy = gamma*invstd*(dpre - sum_dpre/M - xhat*sum_dpre_xhat/M); dres = dpre (optional);
 * train=0 (eval BN): dy = gamma*invstd*dpre. dgamma/dbeta (fp32 [C]) written when non-null. */
int dfu_bn_bwd_apply(const void* dz, const void* z, const uint8_t* zmask, const void* y, const float* mean_invstd, const float* gamma,
                     const float* beta, const float* sums, void* dy, void* dres, float* dgamma, float* dbeta, int64_t M,
                     int32_t C, int32_t relu, int32_t train, void* stream);

/* ------------------------------------ pooling / pointwise ---------------------------------- */
/* MaxPool2d(3, stride 2, pad 1) on NHWC bf16; idx (u8, same shape as y, may be null) records the winning tap
 * (first maximum in row-major scan order, as ATen) for the backward gather. */
int dfu_maxpool3x3s2(const void* x, void* y, uint8_t* idx, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
int dfu_maxpool3x3s2_bwd(const uint8_t* idx, const void* dy, void* dx, int32_t B, int32_t H, int32_t W, int32_t C,
                         void* stream);
/* global average pool [B,HW,C] -> [B, C] written with row stride ldo (lets it land inside the concat buffer) */
int dfu_avgpool(const void* x, void* out, int32_t B, int32_t HW, int32_t C, int64_t ldo, void* stream);
int dfu_avgpool_bwd(const void* dout, int64_t ldo, void* dx, int32_t B, int32_t HW, int32_t C, void* stream);
/* zero `bytes` bytes at p (aten::zeros / zero_ for accumulators: BN/LN sums, split-K wgrad targets, bias grads). */
int dfu_zero(void* p, int64_t bytes, void* stream);
int dfu_relu(const void* x, void* y, int64_t n, void* stream);
int dfu_gelu(const void* x, void* y, int64_t n, void* stream);
int dfu_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, void* stream);
int dfu_add(const void* a, const void* b, void* out, int64_t n, void* stream);
/* dx = dy * gelu'(h) */
int dfu_gelu_bwd(const void* dy, const void* h, void* dx, int64_t n, void* stream);
/* dx = dy * gelu'(h) over a dense [M,N] matrix (N % 8 == 0) and colsum[N] (fp32, caller zeroes) += column sums of dx:
 * the GELU backward and fc1's bias gradient (timm Mlp) in one pass. */
int dfu_gelu_bwd_colsum(const void* dy, const void* h, void* dx, float* colsum, int64_t M, int32_t N, void* stream);
/* column sums of a bf16 [M,N] matrix (row stride ld) added into out[N] fp32 (caller zeroes): bias gradients. */
int dfu_colsum(const void* x, float* out, int64_t M, int32_t N, int64_t ld, void* stream);
/* strided 2-D copy bf16 (rows x cols), used for concat / cls-token selection. */
int dfu_copy2d(const void* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int32_t cols, void* stream);
/* fp32 [rows, cols] (row stride cols) -> bf16 [rows, ldo], pad columns zeroed: weight shadow copies. */
int dfu_cast_f32_bf16(const float* src, void* dst, int64_t rows, int32_t cols, int64_t ldo, void* stream);
int dfu_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream);
/* Bernoulli keep-mask dropout (counter-based hash RNG): y = x*mask/(1-p); mask u8 saved for backward.
 * The seed lives in device memory (advanced by dfu_tick_u64) so a captured CUDA graph draws fresh masks. */
int dfu_dropout(const void* x, void* y, uint8_t* mask, int64_t n, float p, const uint64_t* seed_dev,
                uint64_t stream_id, void* stream);
int dfu_dropout_bwd(const void* dy, const uint8_t* mask, void* dx, int64_t n, float p, void* stream);
int dfu_tick_u64(uint64_t* counter, void* stream);

/* ------------------------------------ ViT pieces (a4) -------------------------------------- */
/* tokens[b,0,:] = cls + pos[0]; tokens[b,1+p,:] = patch_out[b*196+p,:] + pos[1+p]  (timm _pos_embed) */
int dfu_vit_assemble(const void* patch_out, const float* cls, const float* pos, void* tokens, int32_t B,
                     int32_t T, int32_t D, void* stream);
/* adjoint: dpatch = dtokens[:,1:], dcls += sum_b dtokens[b,0], dpos += sum_b dtokens[b] (fp32, caller zeroes) */
int dfu_vit_assemble_bwd(const void* dtokens, void* dpatch, float* dcls, float* dpos, int32_t B, int32_t T,
                         int32_t D, void* stream);
/* LayerNorm over the last dim (eps 1e-6 in timm): y bf16, stats[2*rows] = (mean, rstd) fp32. */
int dfu_layernorm(const void* x, const float* gamma, const float* beta, void* y, float* stats, int64_t rows,
                  int32_t D, float eps, void* stream);
/* dx (bf16) = LN'(dy) [+ dres: the skip-connection gradient of the pre-norm residual block, may be null];
 * dgamma/dbeta += (fp32, caller zeroes). */
int dfu_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* stats, const void* dres,
                      void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t D, void* stream);
/* Row softmax of S[rows, ld] over the first n columns (fp32 math), pad columns written as 0. In place allowed. */
int dfu_softmax_rows(const void* s, void* p, int64_t rows, int32_t n, int64_t ld, void* stream);
/* dS = P * (dP - rowsum(dP*P)) * scale */
int dfu_softmax_rows_bwd(const void* p, const void* dp, void* ds, int64_t rows, int32_t n, int64_t ld, float scale,
                         void* stream);
/* Fused softmax attention over one head's whole sequence (timm Attention.forward: softmax(q k^T * scale) v; replaces
 * F.scaled_dot_product_attention / the bmm-softmax-bmm chain).  head_dim is 64, 128 < T <= 208 tokens.
 *   qkv  [B*T, 3*H*64] bf16, timm order [3][H][64] along the row;  out [B*T, H*64] bf16;
 *   lse  [B, H, T] fp32 = log sum_j exp(scale * q.k_j)  (saved for the backward pass instead of the [T,T] matrix).
 * tcgen05 MMAs with S/O (forward) and S^T/dP^T/dV/dK/dQ (backward) accumulators in TMEM; P never leaves the SM.
 * Every buffer 16-byte aligned (they are read and written through TMA tensor maps). */
int dfu_attention_fwd(const void* qkv, void* out, float* lse, int32_t B, int32_t T, int32_t H, float scale, void* stream);
/* dqkv [B*T, 3*H*64] bf16 (every element written). */
int dfu_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int32_t B,
                      int32_t T, int32_t H, float scale, void* stream);
/* ------------------------------------ loss (a6) -------------------------------------------- */
/* Weighted CrossEntropyLoss (mean reduction: sum w[y]*nll / sum w[y]); weight may be null.
 * logits fp32 [B,C]; labels int64; loss fp32 scalar; dlogits fp32 [B,C] = d loss / d logits. */
int dfu_cross_entropy(const float* logits, const int64_t* labels, const float* weight, float* loss,
                      float* dlogits, int32_t B, int32_t C, void* stream);
/* softmax[:,1] and argmax (extended_metrics.py:636-637) */
int dfu_predict(const float* logits, float* prob1, int64_t* pred, int32_t B, int32_t C, void* stream);
/* out = a * (*scalar_dev): applies the upstream scalar gradient of the loss to dlogits. */
int dfu_scale_by_scalar(const float* a, const float* scalar_dev, float* out, int64_t n, void* stream);

/* ------------------------------------ optimizer (a8) --------------------------------------- */
/* Multi-tensor AdamW (torch.optim.AdamW semantics, one group). `table` is a device array of n_tensors
 * records {p, g, m, v, shadow_bf16(or 0), numel}; g may be fp32 or bf16 (g_bf16 flag);
 * step is a device fp32 scalar holding the step count AFTER increment (kernel increments it itself
 * via dfu_adamw_tick so the whole update is CUDA-graph capturable). */
typedef struct {
    float* p; const void* g; float* m; float* v; void* shadow; int64_t numel;
} dfu_adamw_tensor;
int dfu_adamw_tick(float* step, void* stream);
int dfu_adamw(const dfu_adamw_tensor* table, int32_t n_tensors, const int32_t* chunk_tensor,
              const int32_t* chunk_offset, int32_t n_chunks, int32_t chunk_size, const float* step, float lr,
              float beta1, float beta2, float eps, float weight_decay, float grad_scale, int32_t g_bf16,
              void* stream);
/* Multi-tensor fp32 -> bf16 cast in storage order (same chunk tables as dfu_adamw): the fp32 gradients of one data-parallel
 * bucket are written into the bucket's bf16 wire buffer with one launch (replaces one aten::_to_copy per parameter). */
typedef struct {
    const float* src; void* dst; int64_t numel;
} dfu_cast_tensor;
int dfu_cast_multi_f32_bf16(const dfu_cast_tensor* table, int32_t n_tensors, const int32_t* chunk_tensor,
                            const int32_t* chunk_offset, int32_t n_chunks, int32_t chunk_size, void* stream);
/* Flat variant over one contiguous fp32 parameter arena. */
int dfu_adamw_flat(float* p, const void* g, float* m, float* v, void* shadow, int64_t n, const float* step,
                   float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                   int32_t g_bf16, void* stream);

/* ------------------------------------ Grad-CAM (a11) / TTA (a12) --------------------------- */
/* weights w_c = mean_hw G[c]; cam = relu(sum_{c<nc} w_c A[c]) / max (if max>0).
 * A: [HW, CA] bf16 NHWC activations of one image, G: [HW, CG] bf16; cam fp32 [HW]. */
int dfu_gradcam(const void* A, int32_t CA, const void* G, int32_t CG, int32_t HW, int32_t nc, float* cam,
                void* stream);
/* |dimg|.mean(C) / max  over a fp32 [3,H,W] input gradient -> [H*W] */
int dfu_saliency(const float* dimg, float* out, float* max_scratch, int32_t C, int32_t HW, void* stream);
/* TTA vote: probs [n, V] -> mean prob and majority vote (mean(p>0.5) > 0.5). */
int dfu_tta_vote(const float* probs, float* mean_prob, int64_t* pred, int32_t n, int32_t V, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-side evaluation metrics.  Replaces, per batch, the device->host syncs of the reference loops
 * (train_multimodal_fusion.py:383-386; extended_metrics.py:636-640) and, per epoch, sklearn's confusion_matrix / precision /
 * recall / f1 / roc_auc_score / precision_recall_curve + auc / matthews_corrcoef / cohen_kappa_score
 * (extended_metrics.py:374-446).
 *   state [6] int64 (caller zeroes): n, tn, fp, fn, tp, loss terms;  loss_acc [1] fp64 (caller zeroes);
 *   probs / y_true / y_pred [capacity]: softmax(logits)[:,1] (sigmoid for C = 1), label, argmax of every sample seen.
 *   dfu_metrics_update: one launch per batch, no sync; `loss` (device scalar) may be null.
 *   dfu_metrics_finalize: out24 (fp64) = tn fp fn tp accuracy precision recall f1 sensitivity specificity ppv npv fpr fnr
 *   auc_roc auc_pr mcc kappa balanced_accuracy f_harmonic mean_loss n overflowed reserved;  capacity <= 16384.
 * ------------------------------------------------------------------------------------------- */
int dfu_metrics_update(const float* logits, const int64_t* labels, const float* loss, int32_t B, int32_t C, int64_t* state,
                       double* loss_acc, float* probs, int32_t* y_true, int32_t* y_pred, int64_t capacity, void* stream);
int dfu_metrics_finalize(const int64_t* state, const double* loss_acc, const float* probs, const int32_t* y_true, int64_t capacity,
                         double* out24, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Train-time augmentation on the device, fused with ToTensor + Normalize: replaces the PIL pipeline
 * RandomHorizontalFlip -> RandomVerticalFlip -> RandomRotation(30) -> [ColorJitter] -> [RandomAffine] of
 * train_multimodal_fusion.py:173-199 (torchvision.transforms.functional hflip / vflip / rotate / adjust_brightness /
 * adjust_contrast / adjust_saturation / affine on PIL images).  The HOST draws the parameters and prepares, per image, the
 * inverse maps exactly as torchvision + Pillow do (dfu_multimodal_b200/augment.py); given them the output is bit-identical.
 *   flags: bit 0 horizontal flip, bit 1 vertical flip.
 *   rot_kind / aff_kind: 0 identity, 1 Pillow affine_fixed with the six 16.16 coefficients in rot[] / aff[], 2 / 3 / 4 rotation by
 *   180 / 90 / 270 degrees (Pillow's transpose fast paths), 5 (aff only) Pillow ImagingScaleAffine lookup tables aff_xtab / aff_ytab.
 *   color_op[k] in application order: 0 brightness, 1 contrast, 2 saturation, factor color_f[k]; n_color of them.
 *   src [B,H,W,3] uint8 (H, W <= 224); luma_sums [B] uint64 scratch; out_mode as dfu_resize_u8 (0 uint8 / 1 bf16 NHWC / 2 patches).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t flags;
    int32_t rot_kind;
    int32_t rot[6];
    int32_t aff_kind;
    int32_t aff[6];
    int32_t n_color;
    int32_t color_op[3];
    float color_f[3];
    int16_t aff_xtab[224];
    int16_t aff_ytab[224];
} dfu_augment_params;
int dfu_augment_u8(const uint8_t* src, int32_t B, int32_t H, int32_t W, const dfu_augment_params* params, uint64_t* luma_sums,
                   int32_t out_mode, void* out, const float* mean3_host, const float* std3_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a9, resize half: transforms.Resize((224,224)) of the reference pipelines (train_multimodal_fusion.py:174,186,195,204;
 * extended_metrics.py:290-300) fused with ToTensor + Normalize.  Replaces torchvision F.resize -> PIL.Image.resize(BILINEAR)
 * (Pillow ImagingResample: antialiased triangle filter, 22-bit fixed-point coefficients, horizontal then vertical pass, each
 * rounded to uint8) and is BIT-IDENTICAL to it; out_mode 0 returns the resized uint8 image itself.
 *   src: B uint8 HWC images of one size [Hs, Ws, 3], image b at src + b * img_stride_bytes.
 *   out_mode 0: uint8 [B, OH, OW, 3];  1: normalised bf16 NHWC [B, OH, OW, 3];  2: normalised bf16 ViT patch matrix
 *   [B*196, 768] (OH = OW = 224).  workspace: dfu_resize_workspace_ints(Hs, Ws, OH, OW) int32 values (coefficient tables,
 *   written by the call).  mean3/std3 are HOST pointers (unused for mode 0).
 * ------------------------------------------------------------------------------------------- */
int dfu_resize_workspace_ints(int32_t Hs, int32_t Ws, int32_t OH, int32_t OW);
int dfu_resize_u8(const uint8_t* src, int64_t img_stride_bytes, int32_t B, int32_t Hs, int32_t Ws, int32_t OH, int32_t OW,
                  int32_t out_mode, void* out, int32_t* workspace, const float* mean3_host, const float* std3_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * fp32 VALIDATION MODE (north_star: logits within 1e-4 of the fp32 reference "in fp32-accum validation mode").
 * The inference path once more with fp32 activations (NHWC rows / token rows), the fp32 master weights and fp32
 * accumulation (fp64 for statistics), as plain SIMT kernels; dfu_multimodal_b200/validation.py drives them.  Each entry
 * replaces the same ATen call as its bf16 sibling above (aten::convolution via im2col + mm, aten::native_batch_norm,
 * aten::max_pool2d, aten::mean, aten::native_layer_norm, aten::_softmax, aten::addmm / bmm, aten::gelu, aten::cat).
 * dfu_v32_gemm: D[z][m][n] = act(alpha * sum_k A[z][m][k] * B[z][n][k] + bias[n]) + res[z][m][n], A and B addressed
 * by element strides (row stride, column stride, two batch strides), act 0 none / 1 relu / 2 gelu(erf).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t M, N, K, nb0, nb1;
    const float* A; int64_t a_row_stride, a_col_stride, sa0, sa1;
    const float* B; int64_t b_row_stride, b_col_stride, sb0, sb1;
    float* D; int64_t ldd, sd0, sd1;
    const float* bias;
    const float* res; int64_t ldr, sr0, sr1;
    float alpha;
    int32_t act;
} dfu_v32_gemm_args;
int dfu_v32_gemm(const dfu_v32_gemm_args* a, void* stream);
/* uint8 NHWC -> fp32 NHWC, (x/255 - mean)/std; mean3/std3 are HOST pointers to 3 floats. */
int dfu_v32_preprocess_u8(const uint8_t* img_nhwc, float* out_nhwc, int64_t npix, const float* mean3_host, const float* std3_host, void* stream);
int dfu_v32_nchw_to_nhwc(const float* x, float* y, int32_t B, int32_t C, int32_t H, int32_t W, void* stream);
/* col[(b,ho,wo), (r,s,c)] of an NHWC tensor, zero padded. */
int dfu_v32_im2col(const float* x, float* col, int32_t B, int32_t H, int32_t W, int32_t C, int32_t R, int32_t S, int32_t stride,
                   int32_t pad, int32_t Ho, int32_t Wo, void* stream);
/* timm PatchEmbed unfold of fp32 NHWC [B,224,224,3] -> [B*196, 768], columns ordered (c,i,j). */
int dfu_v32_patchify(const float* x_nhwc, float* out, int32_t B, void* stream);
/* x <- float(bf16(x)) in place: emulates bf16 STORAGE of an activation / weight inside the fp32 path (validation.py uses it to
 * show that the fast path's distance from the fp32 reference is storage rounding, not arithmetic). */
int dfu_v32_round_bf16(float* x, int64_t n, void* stream);
int dfu_v32_copy2d(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int32_t cols, void* stream);
int dfu_v32_vit_assemble(const float* patch_out, const float* cls, const float* pos, float* tokens, int32_t B, int32_t T, int32_t D,
                         void* stream);
/* per-channel batch mean / biased variance of y [M, C] (train-mode BatchNorm), fp64 accumulation. */
int dfu_v32_bn_batch_stats(const float* y, float* mean, float* var, int64_t M, int32_t C, void* stream);
/* out = [relu]((y - mean)/sqrt(var + eps) * gamma + beta [+ res]). */
int dfu_v32_bn_apply(const float* y, const float* mean, const float* var, const float* gamma, const float* beta, const float* res,
                     float* out, int64_t M, int32_t C, int32_t relu, float eps, void* stream);
int dfu_v32_maxpool3x3s2(const float* x, float* y, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
int dfu_v32_avgpool(const float* x, float* y, int32_t B, int32_t HW, int32_t C, void* stream);
int dfu_v32_layernorm(const float* x, const float* gamma, const float* beta, float* y, int64_t rows, int32_t D, float eps, void* stream);
/* in place over rows of length T with row stride ld. */
int dfu_v32_softmax_rows(float* x, int64_t rows, int32_t T, int64_t ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif
