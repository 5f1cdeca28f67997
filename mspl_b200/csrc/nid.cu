// NIDLoss (next-row component): loss_fns/segmentation_loss.py:54-144 (NIDLoss + SoftArgMax), the optional
// `--use-nid` training term (call site uest_seg_multi_os.py:1027-1030, constructed at :514 as
// NIDLoss(image_bin=args.nid_bin, label_bin=args.classes)).
//
// The reference makes image_bin sequential passes of sigmoid windows over all pixels, materialises K x num_pixel and
// C x num_pixel matrices and multiplies them.  Here ONE pass per direction:
//   forward : per pixel position, sum over the batch of the K intensity windows and the Lb label windows (the reference sums
//             over the batch BEFORE forming the joint histogram -- P_c is K x num_pixel -- so images at the same position are
//             coupled; kept), accumulate the joint / marginal histograms, and let the last CTA evaluate
//             NID = 1 - I/H, loss = (NID - 0.95) * 20, together with the gradient of the loss w.r.t. the histograms;
//   backward: per position, recompute the windows and chain through d(window)/d(soft label) and the soft-argmax
//             (beta = 500) to the label logits.  The camera image gets no gradient (the reference feeds it un-tracked).
#include "pixel_math.cuh"

namespace mspl {

constexpr int kNidThreads = 256;
constexpr int kNidMaxImageBins = 32;
constexpr int kNidMaxLabelBins = 8;
constexpr int kNidMaxBlocks = 1024;
constexpr int kNidJoint = kNidMaxImageBins * kNidMaxLabelBins;
constexpr int kNidSlots = kNidJoint + kNidMaxImageBins + kNidMaxLabelBins;      // J | a | l per block

struct NidWorkspace {                       // caller-zeroed once; every forward leaves `ticket` at 0
    unsigned int ticket;
    unsigned int pad[3];
    double partial[kNidMaxBlocks][kNidSlots];
};

struct NidState {                           // forward -> backward hand-over (device)
    float loss;
    float pad[3];
    float GJ[kNidJoint];                    // d loss / d J[k][c]   (J = un-normalised joint histogram)
    float Gl[kNidMaxLabelBins];             // d loss / d l[c]      (l = un-normalised label marginal)
};

// 1 / (1 + e^-x) through MUFU.EX2 + MUFU.RCP (relative error ~2^-22; saturates cleanly to 0 / 1 for large |x|)
MSPL_DEVINL float sigmoidf(float x) { return rcp_fast(1.0f + ex2_approx(-x * kLog2e)); }

// soft label of one pixel: sum_i i * e_i / (sum e + 1e-12), e_i = exp((A_i - max A) * 500)        (SoftArgMax, :124-141)
MSPL_DEVINL float soft_label(const float* __restrict__ px, int C, int64_t hw, float* sum_e_out) {
    float mx = -INFINITY;
    for (int i = 0; i < C; ++i) mx = fmaxf(mx, __ldg(px + i * hw));
    float se = 0.f, sie = 0.f;
    for (int i = 0; i < C; ++i) {
        const float e = ex2_approx((__ldg(px + i * hw) - mx) * (500.0f * kLog2e));
        se += e;
        sie = fmaf((float)i, e, sie);
    }
    if (sum_e_out) *sum_e_out = se;
    return sie / (se + 1e-12f);
}

__global__ void __launch_bounds__(kNidThreads) nid_forward_kernel(const float* __restrict__ camera, const float* __restrict__ label,
                                                                  int B, int C, int64_t hw, int K, int Lb, float bw_c, float bw_l,
                                                                  NidWorkspace* ws, NidState* state) {
    __shared__ double s_acc[kNidSlots];
    __shared__ bool s_last;
    for (int i = threadIdx.x; i < kNidSlots; i += kNidThreads) s_acc[i] = 0.0;
    __syncthreads();
    const float Lc = 1.0f / (float)K, inv_bwc = 1.0f / bw_c, inv_bwl = 1.0f / bw_l;
    const int lane = threadIdx.x & 31;
    for (int64_t base = (int64_t)blockIdx.x * kNidThreads; base < hw; base += (int64_t)gridDim.x * kNidThreads) {
        const int64_t pix = base + threadIdx.x;
        const bool active = pix < hw;
        float Pc[kNidMaxImageBins], Pl[kNidMaxLabelBins];
#pragma unroll
        for (int k = 0; k < kNidMaxImageBins; ++k) Pc[k] = 0.f;
#pragma unroll
        for (int c = 0; c < kNidMaxLabelBins; ++c) Pl[c] = 0.f;
        if (active) {
            for (int b = 0; b < B; ++b) {
                const float* cam = camera + ((int64_t)b * 3) * hw + pix;
                const float g = (__ldg(cam) + __ldg(cam + hw) + __ldg(cam + 2 * hw)) / 3.0f;        // get_grayscale (:65-66)
                const float lab = soft_label(label + ((int64_t)b * C) * hw + pix, C, hw, nullptr);
                // neighbouring windows share an edge: PI[k] = s((v - k L)/bw) - s((v - (k+1) L)/bw), K+1 sigmoids instead of 2K
                float lo = sigmoidf(g * inv_bwc);
#pragma unroll
                for (int k = 0; k < kNidMaxImageBins; ++k) {
                    if (k < K) {
                        const float hi = sigmoidf((g - Lc * (float)(k + 1)) * inv_bwc);
                        Pc[k] += lo - hi;
                        lo = hi;
                    }
                }
                lo = sigmoidf((lab + 0.5f) * inv_bwl);
#pragma unroll
                for (int c = 0; c < kNidMaxLabelBins; ++c) {
                    if (c < Lb) {
                        const float hi = sigmoidf((lab - 0.5f - (float)c) * inv_bwl);
                        Pl[c] += lo - hi;
                        lo = hi;
                    }
                }
            }
        }
        // warp-reduce the outer product and the marginals, one shared-memory add per warp and entry
#pragma unroll
        for (int k = 0; k < kNidMaxImageBins; ++k) {
            if (k < K) {
#pragma unroll
                for (int c = 0; c < kNidMaxLabelBins; ++c) {
                    if (c < Lb) {
                        float v = Pc[k] * Pl[c];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        if (lane == 0) atomicAdd(&s_acc[k * kNidMaxLabelBins + c], (double)v);
                    }
                }
                float v = Pc[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) atomicAdd(&s_acc[kNidJoint + k], (double)v);
            }
        }
#pragma unroll
        for (int c = 0; c < kNidMaxLabelBins; ++c) {
            if (c < Lb) {
                float v = Pl[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) atomicAdd(&s_acc[kNidJoint + kNidMaxImageBins + c], (double)v);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kNidSlots; i += kNidThreads) ws->partial[blockIdx.x][i] = s_acc[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- last CTA: total histograms (blocks added in index order), NID and its gradient w.r.t. the histograms ----
    for (int i = threadIdx.x; i < kNidSlots; i += kNidThreads) {
        double t = 0.0;
        for (int b = 0; b < (int)gridDim.x; ++b) t += __ldcg(&ws->partial[b][i]);
        s_acc[i] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double eps = 1e-7;
        double* J = s_acc;
        double* a = s_acc + kNidJoint;
        double* l = s_acc + kNidJoint + kNidMaxImageBins;
        double SJ = 0, Sa = 0, Sl = 0;
        for (int k = 0; k < K; ++k) { Sa += a[k]; for (int c = 0; c < Lb; ++c) SJ += J[k * kNidMaxLabelBins + c]; }
        for (int c = 0; c < Lb; ++c) Sl += l[c];
        double I = 0, H = 0;
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < Lb; ++c) {
                const double p = J[k * kNidMaxLabelBins + c] / SJ, q = (a[k] / Sa) * (l[c] / Sl);
                I += p * (log(p + eps) - log(q + eps));
                H -= p * log(p + eps);
            }
        state->loss = (float)(((1.0 - I / H) - 0.95) * 20.0);                               // (:113-118)
        // gradients: loss = 20 (1 - I/H - 0.95)
        const double dI = -20.0 / H, dH = 20.0 * I / (H * H);
        double gP[kNidJoint], gpl[kNidMaxLabelBins];
        double dotP = 0, dotl = 0;
        for (int c = 0; c < Lb; ++c) gpl[c] = 0;
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < Lb; ++c) {
                const double p = J[k * kNidMaxLabelBins + c] / SJ, pc = a[k] / Sa, pl = l[c] / Sl, q = pc * pl;
                const double lp = log(p + eps), r = p / (p + eps);
                gP[k * kNidMaxLabelBins + c] = dI * (lp - log(q + eps) + r) + dH * (-lp - r);
                dotP += gP[k * kNidMaxLabelBins + c] * p;
                gpl[c] += dI * (-p / (q + eps)) * pc;                                       // via q = p_c p_l
            }
        for (int c = 0; c < Lb; ++c) dotl += gpl[c] * (l[c] / Sl);
        for (int k = 0; k < K; ++k)
            for (int c = 0; c < Lb; ++c) state->GJ[k * kNidMaxLabelBins + c] = (float)((gP[k * kNidMaxLabelBins + c] - dotP) / SJ);
        for (int c = 0; c < Lb; ++c) state->Gl[c] = (float)((gpl[c] - dotl) / Sl);
        ws->ticket = 0;
    }
}

__global__ void __launch_bounds__(kNidThreads) nid_backward_kernel(const float* __restrict__ camera, const float* __restrict__ label,
                                                                   const float* __restrict__ grad_loss, int B, int C, int64_t hw, int K,
                                                                   int Lb, float bw_c, float bw_l, const NidState* __restrict__ state,
                                                                   float* __restrict__ d_label) {
    __shared__ float s_GJ[kNidJoint], s_Gl[kNidMaxLabelBins];
    for (int i = threadIdx.x; i < kNidJoint; i += kNidThreads) s_GJ[i] = state->GJ[i];
    if (threadIdx.x < kNidMaxLabelBins) s_Gl[threadIdx.x] = state->Gl[threadIdx.x];
    __syncthreads();
    const float up = grad_loss[0];
    const float Lc = 1.0f / (float)K, inv_bwc = 1.0f / bw_c, inv_bwl = 1.0f / bw_l;
    for (int64_t pix = (int64_t)blockIdx.x * kNidThreads + threadIdx.x; pix < hw; pix += (int64_t)gridDim.x * kNidThreads) {
        float W[kNidMaxLabelBins];                    // d loss / d P_l[c, pix] = sum_k GJ[k][c] P_c[k, pix] + Gl[c]
#pragma unroll
        for (int c = 0; c < kNidMaxLabelBins; ++c) W[c] = c < Lb ? s_Gl[c] : 0.f;
        float Pc[kNidMaxImageBins];
#pragma unroll
        for (int k = 0; k < kNidMaxImageBins; ++k) Pc[k] = 0.f;
        for (int b = 0; b < B; ++b) {
            const float* cam = camera + ((int64_t)b * 3) * hw + pix;
            const float g = (__ldg(cam) + __ldg(cam + hw) + __ldg(cam + 2 * hw)) / 3.0f;
            float lo = sigmoidf(g * inv_bwc);
#pragma unroll
            for (int k = 0; k < kNidMaxImageBins; ++k) {
                if (k < K) {
                    const float hi = sigmoidf((g - Lc * (float)(k + 1)) * inv_bwc);
                    Pc[k] += lo - hi;
                    lo = hi;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kNidMaxImageBins; ++k)
            if (k < K) {
#pragma unroll
                for (int c = 0; c < kNidMaxLabelBins; ++c)
                    if (c < Lb) W[c] = fmaf(s_GJ[k * kNidMaxLabelBins + c], Pc[k], W[c]);
            }
        for (int b = 0; b < B; ++b) {
            const float* px = label + ((int64_t)b * C) * hw + pix;
            float se;
            const float lab = soft_label(px, C, hw, &se);
            float dlab = 0.f;                         // d loss / d lab
#pragma unroll
            for (int c = 0; c < kNidMaxLabelBins; ++c) {
                if (c < Lb) {
                    const float sa = sigmoidf((lab - (float)c + 0.5f) * inv_bwl), sb = sigmoidf((lab - (float)c - 0.5f) * inv_bwl);
                    dlab = fmaf(W[c], (sa * (1.0f - sa) - sb * (1.0f - sb)) * inv_bwl, dlab);
                }
            }
            dlab *= up;
            float mx = -INFINITY;
            for (int i = 0; i < C; ++i) mx = fmaxf(mx, __ldg(px + i * hw));
            const float inv = 1.0f / (se + 1e-12f);
            for (int j = 0; j < C; ++j) {             // d lab / d A_j = beta * s_j * (j - lab)
                const float sj = ex2_approx((__ldg(px + j * hw) - mx) * (500.0f * kLog2e)) * inv;
                d_label[((int64_t)b * C + j) * hw + pix] = dlab * 500.0f * sj * ((float)j - lab);
            }
        }
    }
}

}  // namespace mspl

using namespace mspl;

extern "C" size_t mspl_nid_workspace_bytes(void) { return sizeof(NidWorkspace); }
extern "C" size_t mspl_nid_state_bytes(void) { return sizeof(NidState); }

static bool nid_bad(const float* camera, const float* label, int64_t b, int c, int64_t hw, int k, int lb, float bwc, float bwl) {
    return !camera || !label || b < 1 || b > (1 << 20) || c < 1 || hw < 1 || k < 1 || k > kNidMaxImageBins || lb < 1 ||
           lb > kNidMaxLabelBins || !(bwc > 0.f) || !(bwl > 0.f);
}

extern "C" int mspl_nid_fwd(const float* camera, const float* label, int64_t batch, int num_classes, int64_t pixels_per_image,
                            int image_bins, int label_bins, float bw_camera, float bw_label, void* workspace, size_t workspace_bytes,
                            void* state, void* stream) {
    if (nid_bad(camera, label, batch, num_classes, pixels_per_image, image_bins, label_bins, bw_camera, bw_label) || !workspace || !state)
        return MSPL_ERR_BAD_ARG;
    if (workspace_bytes < sizeof(NidWorkspace)) return MSPL_ERR_WORKSPACE;
    if (!aligned_to(workspace, 8) || !aligned_to(state, 4) || !aligned_to(camera, 4) || !aligned_to(label, 4)) return MSPL_ERR_ALIGN;
    int64_t grid = (pixels_per_image + kNidThreads - 1) / kNidThreads;
    grid = grid < kNidMaxBlocks ? grid : kNidMaxBlocks;
    nid_forward_kernel<<<(unsigned)grid, kNidThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        camera, label, (int)batch, num_classes, pixels_per_image, image_bins, label_bins, bw_camera, bw_label,
        static_cast<NidWorkspace*>(workspace), static_cast<NidState*>(state));
    return launch_status();
}

extern "C" int mspl_nid_bwd(const float* camera, const float* label, const float* grad_loss, int64_t batch, int num_classes,
                            int64_t pixels_per_image, int image_bins, int label_bins, float bw_camera, float bw_label,
                            const void* state, float* d_label, void* stream) {
    if (nid_bad(camera, label, batch, num_classes, pixels_per_image, image_bins, label_bins, bw_camera, bw_label) || !grad_loss || !state ||
        !d_label)
        return MSPL_ERR_BAD_ARG;
    int64_t grid = (pixels_per_image + kNidThreads - 1) / kNidThreads;
    grid = grid < kNumSMs * 8 ? grid : kNumSMs * 8;
    nid_backward_kernel<<<(unsigned)grid, kNidThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        camera, label, grad_loss, (int)batch, num_classes, pixels_per_image, image_bins, label_bins, bw_camera, bw_label,
        static_cast<const NidState*>(state), d_label);
    return launch_status();
}
