// Device-side evaluation metrics: what notebooks/extended_metrics.py:374-446
// (MedicalMetricsCalculator: sklearn confusion_matrix / precision / recall / f1 / roc_auc_score / precision_recall_curve + auc /
// matthews_corrcoef / cohen_kappa_score) and the per-step host reads of the training loop
// (train_multimodal_fusion.py:383-386: loss.item(), predicted.cpu(), labels.cpu()) compute on the host after a device->host
// sync per batch.  Here every batch appends (softmax[:,1], label, argmax) to device buffers and bumps the confusion counts and
// the loss sum with ONE small kernel and no synchronisation; dfu_metrics_finalize turns the buffers into the full metric
// vector (fp64, one block: bitonic sort by score, then one pass over the distinct thresholds for the ROC and PR curves), which
// the host reads once per epoch.
#include "common.cuh"
#include "../../include/dfu_b200.h"

namespace {

// state: [0] n appended, [1..4] tn fp fn tp (as the reference's cm.ravel()), [5] number of loss terms; lossacc: running loss sum
__global__ void __launch_bounds__(256) metrics_update_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                             const float* __restrict__ loss, int B, int C, long long* __restrict__ state,
                                                             double* __restrict__ lossacc, float* __restrict__ probs,
                                                             int* __restrict__ ytrue, int* __restrict__ ypred, long long capacity) {
    pdl_sync();
    __shared__ int conf[4];
    if (threadIdx.x < 4) conf[threadIdx.x] = 0;
    __syncthreads();
    const long long n0 = state[0];
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        const float* l = logits + (long long)i * C;
        float p1;
        int pred;
        if (C == 1) {                                     // sigmoid head (test_time_augmentation.py:109-132)
            p1 = 1.f / (1.f + expf(-l[0]));
            pred = p1 > 0.5f;
        } else {
            float mx = l[0];
            pred = 0;
            for (int c = 1; c < C; ++c) if (l[c] > mx) { mx = l[c]; pred = c; }
            float den = 0.f;
            for (int c = 0; c < C; ++c) den += expf(l[c] - mx);
            p1 = expf(l[1] - mx) / den;                   // torch.softmax(outputs, dim=1)[:, 1]
        }
        const int y = (int)labels[i];
        if (n0 + i < capacity) {
            probs[n0 + i] = p1;
            ytrue[n0 + i] = y;
            ypred[n0 + i] = pred;
        }
        atomicAdd(&conf[(y != 0) * 2 + (pred != 0)], 1);
    }
    __syncthreads();
    if (threadIdx.x < 4) state[1 + threadIdx.x] += conf[threadIdx.x];
    if (threadIdx.x == 0) {
        state[0] = n0 + B;
        if (loss) { *lossacc += (double)*loss; state[5] += 1; }
    }
}

struct ScoreLabel { float s; int y; };

// out[24] (fp64): tn fp fn tp accuracy precision recall f1 sensitivity specificity ppv npv fpr fnr auc_roc auc_pr mcc kappa
//                 balanced_accuracy f_harmonic mean_loss n overflowed reserved
__global__ void __launch_bounds__(1024) metrics_finalize_kernel(const long long* __restrict__ state, const double* __restrict__ lossacc,
                                                                const float* __restrict__ probs, const int* __restrict__ ytrue,
                                                                long long capacity, double* __restrict__ out) {
    pdl_sync();
    extern __shared__ ScoreLabel sl[];
    const long long ntot = state[0];
    const int n = (int)(ntot < capacity ? ntot : capacity);
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        ScoreLabel e;
        e.s = i < n ? probs[i] : -INFINITY;               // padding sorts to the end (descending order)
        e.y = i < n ? (ytrue[i] != 0) : 0;
        sl[i] = e;
    }
    __syncthreads();
    for (int k = 2; k <= np2; k <<= 1)                      // bitonic sort, descending by score
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool desc = (i & k) == 0;
                    const ScoreLabel a = sl[i], b = sl[ixj];
                    if (desc ? (a.s < b.s) : (a.s > b.s)) { sl[i] = b; sl[ixj] = a; }
                }
            }
            __syncthreads();
        }
    if (threadIdx.x != 0) return;
    const double tn = (double)state[1], fp = (double)state[2], fn = (double)state[3], tp = (double)state[4];
    const double tot = tn + fp + fn + tp;
    auto ratio = [](double a, double b) { return b > 0.0 ? a / b : 0.0; };
    const double precision = ratio(tp, tp + fp), recall = ratio(tp, tp + fn);
    const double sens = recall, spec = ratio(tn, tn + fp);
    out[0] = tn; out[1] = fp; out[2] = fn; out[3] = tp;
    out[4] = ratio(tp + tn, tot);
    out[5] = precision;
    out[6] = recall;
    out[7] = (precision + recall) > 0.0 ? 2.0 * precision * recall / (precision + recall) : 0.0;
    out[8] = sens;
    out[9] = spec;
    out[10] = precision;                                    // ppv
    out[11] = ratio(tn, tn + fn);
    out[12] = ratio(fp, fp + tn);
    out[13] = ratio(fn, fn + tp);
    // ROC and PR curves over the distinct thresholds, scores descending (sklearn _binary_clf_curve); trapezoids as sklearn's auc()
    double P = 0.0;
    for (int i = 0; i < n; ++i) P += sl[i].y;
    const double N = (double)n - P;
    double ctp = 0.0, cfp = 0.0, roc = 0.0, pr = 0.0;
    double prev_fpr = 0.0, prev_tpr = 0.0, prev_rec = 0.0, prev_prec = 1.0;
    for (int i = 0; i < n; ++i) {
        ctp += sl[i].y;
        cfp += 1 - sl[i].y;
        if (i + 1 < n && sl[i + 1].s == sl[i].s) continue;   // ties share one threshold
        const double fpr = N > 0.0 ? cfp / N : 0.0, tpr = P > 0.0 ? ctp / P : 0.0;
        roc += (fpr - prev_fpr) * (tpr + prev_tpr) * 0.5;
        const double prec = ctp / (ctp + cfp), rec = tpr;
        pr += (rec - prev_rec) * (prec + prev_prec) * 0.5;
        prev_fpr = fpr; prev_tpr = tpr; prev_rec = rec; prev_prec = prec;
    }
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    out[14] = (P > 0.0 && N > 0.0) ? roc : nan;             // sklearn raises when only one class is present
    out[15] = P > 0.0 ? pr : nan;
    const double den = sqrt((tp + fp) * (tp + fn) * (tn + fp) * (tn + fn));
    out[16] = den > 0.0 ? (tp * tn - fp * fn) / den : 0.0;
    const double po = ratio(tp + tn, tot);
    const double pe = tot > 0.0 ? ((tp + fp) * (tp + fn) + (tn + fn) * (tn + fp)) / (tot * tot) : 0.0;
    out[17] = (1.0 - pe) != 0.0 ? (po - pe) / (1.0 - pe) : nan;
    out[18] = 0.5 * (sens + spec);
    out[19] = (sens + spec) > 0.0 ? 2.0 * sens * spec / (sens + spec) : 0.0;
    out[20] = state[5] > 0 ? *lossacc / (double)state[5] : nan;
    out[21] = (double)ntot;
    out[22] = ntot > capacity ? 1.0 : 0.0;
    out[23] = 0.0;
}

}  // namespace

extern "C" int dfu_metrics_update(const float* logits, const int64_t* labels, const float* loss, int32_t B, int32_t C, int64_t* state,
                                  double* loss_acc, float* probs, int32_t* y_true, int32_t* y_pred, int64_t capacity, void* stream) {
    DFU_CHECK_ARG(logits && labels && state && loss_acc && probs && y_true && y_pred && B > 0 && C > 0 && capacity > 0,
                  "dfu_metrics_update: bad args");
    dfu_launch(metrics_update_kernel, 1, 256, 0, (cudaStream_t)stream, logits, (const long long*)labels, loss, B, C, (long long*)state,
               loss_acc, probs, y_true, y_pred, (long long)capacity);
    DFU_CHECK_LAUNCH("dfu_metrics_update");
    return DFU_OK;
}

extern "C" int dfu_metrics_finalize(const int64_t* state, const double* loss_acc, const float* probs, const int32_t* y_true,
                                    int64_t capacity, double* out24, void* stream) {
    DFU_CHECK_ARG(state && loss_acc && probs && y_true && out24 && capacity > 0 && capacity <= 16384,
                  "dfu_metrics_finalize: capacity must be in 1..16384 (one block sorts the scores in shared memory)");
    long long np2 = 1;
    while (np2 < capacity) np2 <<= 1;
    const size_t smem = (size_t)np2 * sizeof(ScoreLabel);
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(metrics_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * (int)sizeof(ScoreLabel));
        attr = true;
    }
    dfu_launch(metrics_finalize_kernel, 1, 1024, smem, (cudaStream_t)stream, (const long long*)state, loss_acc, probs, y_true,
               (long long)capacity, out24);
    DFU_CHECK_LAUNCH("dfu_metrics_finalize");
    return DFU_OK;
}
