// loss.cu -- the training loss of bin/train_flipflop.py:163-182 as ONE call:
//     loss_b = crf_cost_b + logZ_b / nblk,   d loss_b / d scores
// The label-constrained chains (crf_chain_kernel) and the partition-function
// chains (logz_chain_kernel) are independent, latency-bound and small (2N CTAs
// and 2N warps), so they run CONCURRENTLY: the logZ chains are forked onto a
// library-owned side stream and joined before the posterior kernels.  The CRF
// posterior writes -G/nblk once; the logZ posterior adds P/nblk into the same
// rows, so the combined gradient tensor is produced without an extra pass.
#include <cstdlib>
#include <mutex>

#include "common.cuh"

extern "C" int ty_flipflop_logz_phase(const float *scores, int ld, int nblk, int nbatch,
                                      int nbase, float logz_scale, float *logz_out,
                                      float grad_scale, float *grad_out, int ld_grad,
                                      int accumulate, void *workspace, size_t workspace_bytes,
                                      int phases, void *stream);

namespace ty {

struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};

static SideStream *side_for_current_device() {
    static std::mutex mu;
    static SideStream per_device[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    SideStream &s = per_device[dev];
    if (!s.stream) {
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
            s.stream = nullptr;
            return nullptr;
        }
    }
    return &s;
}

}  // namespace ty

using namespace ty;

extern "C" size_t ty_flipflop_train_loss_workspace_bytes(int ntrans, int nblk, int nbatch,
                                                         int max_seqlen, int want_grad) {
    const size_t a = ty_crf_flipflop_workspace_bytes(ntrans, nblk, nbatch, max_seqlen, want_grad);
    const size_t b = ty_flipflop_logz_workspace_bytes(4, nblk, nbatch);
    return (a + 255) / 256 * 256 + b;
}

extern "C" int ty_flipflop_train_loss(const float *scores, int ntrans, int nblk, int nbatch,
                                      const int32_t *moveidx, const int32_t *stayidx,
                                      const int32_t *modmoveidx, const float *modmovefact,
                                      const int32_t *seqlen, int max_seqlen, float sharp,
                                      int ncan, float *cost_out, float *logz_out,
                                      float *grad_out, void *workspace, size_t workspace_bytes,
                                      void *stream) {
    if (!scores || !cost_out || !logz_out || nblk <= 0 || nbatch <= 0) {
        set_error("ty_flipflop_train_loss: bad argument");
        return TY_EINVAL;
    }
    if (ncan != 40) {
        set_error("ty_flipflop_train_loss: only 4-base flip-flop (40 transitions) is implemented");
        return TY_EINVAL;
    }
    const int want_grad = grad_out != nullptr;
    const size_t crf_bytes =
        (ty_crf_flipflop_workspace_bytes(ntrans, nblk, nbatch, max_seqlen, want_grad) + 255) /
        256 * 256;
    const size_t z_bytes = ty_flipflop_logz_workspace_bytes(4, nblk, nbatch);
    if (!workspace || workspace_bytes < crf_bytes + z_bytes) {
        set_error("ty_flipflop_train_loss: workspace %zu < %zu bytes", workspace_bytes,
                  crf_bytes + z_bytes);
        return TY_EWORKSPACE;
    }
    char *ws = static_cast<char *>(workspace);
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    // TY_LOSS_SERIAL=1: partition-function chains on the caller's stream (A/B timing)
    static const bool serial = [] { const char *e = getenv("TY_LOSS_SERIAL"); return e && e[0] == '1'; }();
    SideStream *side = serial ? nullptr : side_for_current_device();
    const float inv = 1.0f / (float)nblk;
    int rc;
    // ---- fork: partition-function chains on the side stream ----
    cudaStream_t zs = user;
    if (side) {
        cudaEventRecord(side->fork, user);
        cudaStreamWaitEvent(side->stream, side->fork, 0);
        zs = side->stream;
    }
    // TY_LOGZ_PACK=0: one chain per CTA next to the CRF chains (A/B timing)
    static const bool unpacked = [] { const char *e = getenv("TY_LOGZ_PACK"); return e && e[0] == '0'; }();
    const int own_sms = side && !unpacked ? 4 : 0;
    rc = ty_flipflop_logz_phase(scores, ntrans, nblk, nbatch, 4, inv, logz_out, inv, grad_out,
                                ntrans, 1, ws + crf_bytes, z_bytes, 1 | own_sms, zs);
    if (side) cudaEventRecord(side->join, side->stream);
    if (rc) {
        if (side) cudaStreamWaitEvent(user, side->join, 0);
        return rc;
    }
    // ---- label-constrained chains + posterior on the caller's stream ----
    // cost = -score / nblk / sharp ; gradient = -G / nblk   (ctc.pyx:66,113,145)
    rc = ty_crf_flipflop(scores, ntrans, nblk, nbatch, moveidx, stayidx, modmoveidx, modmovefact,
                         seqlen, max_seqlen, sharp, ncan, -inv / sharp, cost_out, -inv, grad_out,
                         ws, crf_bytes, user);
    // ---- join, then add the logZ posterior into the same gradient rows ----
    if (side) cudaStreamWaitEvent(user, side->join, 0);
    if (rc || !want_grad) return rc;
    return ty_flipflop_logz_phase(scores, ntrans, nblk, nbatch, 4, inv, logz_out, inv, grad_out,
                                  ntrans, 1, ws + crf_bytes, z_bytes, 2, user);
}
