// GPU MIOU.get_iou (utilities/metrics/segmentation_miou.py:13-44): per-class intersection / prediction / mask pixel
// counts of a batch, without the reference's round trip to the CPU and its three torch.histc passes.
//
// Reference semantics reproduced exactly (integers, so bit-exact): pred and target are cast to uint8 and shifted by one
// "so that 255 is 0"; pixels whose shifted target is 0 are dropped; a class id >= num_classes falls outside histc's
// [1, num_classes] range and is not counted; area_union = area_pred + area_mask - area_inter + 1e-6 (host side).
#include "common.cuh"

namespace mspl {

constexpr int kMiouThreads = 256;
constexpr int kMiouMaxClasses = 256;

// pred comes either as logits (n, c, hw) -> first-max argmax over classes (torch.max(output, 1), :19) or as labels.
template <typename PredT, bool FROM_LOGITS>
__global__ void __launch_bounds__(kMiouThreads) miou_kernel(const float* __restrict__ logits, const PredT* __restrict__ pred_lab,
                                                            const int64_t* __restrict__ target, int64_t n, int c, int64_t hw,
                                                            int num_classes, unsigned long long* __restrict__ out) {
    __shared__ uint32_t s_cnt[3 * kMiouMaxClasses];      // [inter | pred | mask]
    for (int i = threadIdx.x; i < 3 * num_classes; i += kMiouThreads) s_cnt[i] = 0;
    __syncthreads();
    const int64_t npix = n * hw;
    for (int64_t i = blockIdx.x * (int64_t)kMiouThreads + threadIdx.x; i < npix; i += (int64_t)gridDim.x * kMiouThreads) {
        uint32_t p;
        if (FROM_LOGITS) {
            const int64_t img = i / hw, off = i - img * hw;
            const float* px = logits + img * c * hw + off;
            float best = __ldcs(px);
            p = 0;
            for (int k = 1; k < c; ++k) {
                const float v = __ldcs(px + k * hw);
                if (v > best) { best = v; p = k; }       // strict >: first maximal index
            }
        } else {
            p = (uint32_t)pred_lab[i];
        }
        const uint32_t ps = (p + 1u) & 0xffu;                              // ByteTensor cast, += 1 (:30-35)
        const uint32_t ts = ((uint32_t)__ldcs(target + i) + 1u) & 0xffu;
        const bool valid = ts > 0;                                         // pred * (target > 0) (:37)
        // warp-aggregated shared-memory counting: lanes holding the same class elect one adder
        const uint32_t kp = (valid && ps >= 1 && ps <= (uint32_t)num_classes) ? ps : 0;
        const uint32_t kt = (ts >= 1 && ts <= (uint32_t)num_classes) ? ts : 0;
        const uint32_t ki = (kp && kp == ts) ? kp : 0;
        const uint32_t keys[3] = {ki, kp, kt};
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const uint32_t peers = __match_any_sync(__activemask(), keys[h]);
            if (keys[h] && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1))
                atomicAdd(&s_cnt[h * num_classes + keys[h] - 1], (uint32_t)__popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * num_classes; i += kMiouThreads)
        if (s_cnt[i]) atomicAdd(out + i, (unsigned long long)s_cnt[i]);
}

}  // namespace mspl

using namespace mspl;

static int64_t miou_grid(int64_t npix) {
    int64_t b = (npix + kMiouThreads - 1) / kMiouThreads;
    return b < 1 ? 1 : (b < kNumSMs * 8 ? b : kNumSMs * 8);
}

extern "C" int mspl_miou_from_logits(const float* logits, const int64_t* target, int64_t n, int c, int64_t pixels_per_image,
                                     int num_classes, unsigned long long* counts, void* stream) {
    if (!logits || !target || !counts || n < 0 || c < 1 || c > kMiouMaxClasses || pixels_per_image < 1) return MSPL_ERR_BAD_ARG;
    if (num_classes < 1 || num_classes > kMiouMaxClasses - 1) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(logits, 4) || !aligned_to(target, 8) || !aligned_to(counts, 8)) return MSPL_ERR_ALIGN;
    if (n == 0) return MSPL_OK;
    miou_kernel<uint8_t, true><<<(unsigned)miou_grid(n * pixels_per_image), kMiouThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        logits, nullptr, target, n, c, pixels_per_image, num_classes, counts);
    return launch_status();
}

extern "C" int mspl_miou_from_labels(const void* pred, int pred_is_int64, const int64_t* target, int64_t num_pixels, int num_classes,
                                     unsigned long long* counts, void* stream) {
    if (!pred || !target || !counts || num_pixels < 0) return MSPL_ERR_BAD_ARG;
    if (num_classes < 1 || num_classes > kMiouMaxClasses - 1) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(target, 8) || !aligned_to(counts, 8) || (pred_is_int64 && !aligned_to(pred, 8))) return MSPL_ERR_ALIGN;
    if (num_pixels == 0) return MSPL_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)miou_grid(num_pixels);
    if (pred_is_int64)
        miou_kernel<int64_t, false><<<grid, kMiouThreads, 0, st>>>(nullptr, static_cast<const int64_t*>(pred), target, 1, 1, num_pixels,
                                                                  num_classes, counts);
    else
        miou_kernel<uint8_t, false><<<grid, kMiouThreads, 0, st>>>(nullptr, static_cast<const uint8_t*>(pred), target, 1, 1, num_pixels,
                                                                  num_classes, counts);
    return launch_status();
}
