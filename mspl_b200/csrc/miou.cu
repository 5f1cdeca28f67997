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
constexpr int kMiouFastClasses = 32;         // per-thread shared-memory counters up to here (32 KB per CTA)
constexpr int kMiouFieldBits = 10;
constexpr int kMiouFlushPixels = (1 << kMiouFieldBits) - 8;

// The (inter, pred, mask) contribution of one pixel, keys in [1, num_classes] or 0 = not counted.
struct MiouKeys { uint32_t ki, kp, kt; };
MSPL_DEVINL MiouKeys miou_keys(uint32_t p, uint32_t t, uint32_t num_classes) {
    const uint32_t ps = (p + 1u) & 0xffu;                              // ByteTensor cast, += 1 (:30-35)
    const uint32_t ts = (t + 1u) & 0xffu;
    const bool valid = ts > 0;                                         // pred * (target > 0) (:37)
    MiouKeys k;
    k.kp = (valid && ps >= 1 && ps <= num_classes) ? ps : 0;
    k.kt = (ts >= 1 && ts <= num_classes) ? ts : 0;
    k.ki = (k.kp && k.kp == ts) ? k.kp : 0;
    return k;
}

// pred comes either as logits (n, c, hw) -> first-max argmax over classes (torch.max(output, 1), :19) or as labels.
// P pixels per thread (vector loads of the class planes when rows are 16-byte aligned).
// FAST (num_classes <= kMiouFastClasses): per-thread counters in shared memory, one word per class and thread
// ([class][thread]: conflict-free), three 10-bit fields [inter | pred | mask], folded into the CTA totals before a field
// can overflow -- no atomics or warp votes per pixel.  Otherwise warp-aggregated shared-memory atomics.
template <typename PredT, bool FROM_LOGITS, int P, bool FAST>
__global__ void __launch_bounds__(kMiouThreads) miou_kernel(const float* __restrict__ logits, const PredT* __restrict__ pred_lab,
                                                            const int64_t* __restrict__ target, int64_t n, int c, int64_t hw,
                                                            int num_classes, unsigned long long* __restrict__ out) {
    __shared__ uint32_t s_tot[3 * kMiouMaxClasses];      // [inter | pred | mask]
    extern __shared__ uint32_t s_cnt[];                  // FAST: [num_classes][kMiouThreads]
    for (int i = threadIdx.x; i < 3 * num_classes; i += kMiouThreads) s_tot[i] = 0;
    uint32_t* const my_cnt = s_cnt + threadIdx.x;
    if (FAST)
        for (int k = 0; k < num_classes; ++k) my_cnt[k * kMiouThreads] = 0;
    __syncthreads();
    constexpr uint32_t kMask = (1u << kMiouFieldBits) - 1u;
    auto flush = [&]() {
        for (int k = 0; k < num_classes; ++k) {
            const uint32_t word = my_cnt[k * kMiouThreads];
            if (word == 0) continue;
            my_cnt[k * kMiouThreads] = 0;
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const uint32_t v = (word >> (f * kMiouFieldBits)) & kMask;      // f: 0 mask, 1 pred, 2 inter
                if (v) atomicAdd(&s_tot[(2 - f) * num_classes + k], v);
            }
        }
    };
    const int64_t gpi = hw / P, n_groups = n * gpi;
    int pending = 0;
    for (int64_t g = blockIdx.x * (int64_t)kMiouThreads + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * kMiouThreads) {
        const int64_t img = g / gpi, off = (g - img * gpi) * P;
        const int64_t i0 = img * hw + off;
        uint32_t p[P];
        if (FROM_LOGITS) {
            const float* px = logits + img * c * hw + off;
            float best[P], v[P];
            PixVec<P>::load(px, best);
#pragma unroll
            for (int q = 0; q < P; ++q) p[q] = 0;
#pragma unroll 4
            for (int k = 1; k < c; ++k) {
                PixVec<P>::load(px + k * hw, v);
#pragma unroll
                for (int q = 0; q < P; ++q)
                    if (v[q] > best[q]) { best[q] = v[q]; p[q] = k; }       // strict >: first maximal index
            }
        } else {
#pragma unroll
            for (int q = 0; q < P; ++q) p[q] = (uint32_t)pred_lab[i0 + q];
        }
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const MiouKeys key = miou_keys(p[q], (uint32_t)__ldcs(target + i0 + q), (uint32_t)num_classes);
            if (FAST) {
                if (key.kp) my_cnt[(key.kp - 1) * kMiouThreads] += 1u << kMiouFieldBits;
                if (key.kt) my_cnt[(key.kt - 1) * kMiouThreads] += 1u + (key.ki ? (1u << (2 * kMiouFieldBits)) : 0u);
            } else {
                // warp-aggregated shared-memory counting: lanes holding the same class elect one adder
                const uint32_t keys[3] = {key.ki, key.kp, key.kt};
#pragma unroll
                for (int h = 0; h < 3; ++h) {
                    const uint32_t peers = __match_any_sync(__activemask(), keys[h]);
                    if (keys[h] && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1))
                        atomicAdd(&s_tot[h * num_classes + keys[h] - 1], (uint32_t)__popc(peers));
                }
            }
        }
        if (FAST) {
            pending += P;
            if (pending >= kMiouFlushPixels) { flush(); pending = 0; }
        }
    }
    if (FAST) flush();
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * num_classes; i += kMiouThreads)
        if (s_tot[i]) atomicAdd(out + i, (unsigned long long)s_tot[i]);
}

}  // namespace mspl

using namespace mspl;

static int64_t miou_grid(int64_t n_groups) {
    int64_t b = (n_groups + kMiouThreads - 1) / kMiouThreads;
    return b < 1 ? 1 : (b < kNumSMs * 6 ? b : kNumSMs * 6);
}

template <typename PredT, bool FROM_LOGITS, int P>
static int miou_launch(const float* logits, const PredT* pred, const int64_t* target, int64_t n, int c, int64_t hw, int num_classes,
                       unsigned long long* counts, cudaStream_t st) {
    const unsigned grid = (unsigned)miou_grid(n * (hw / P));
    if (num_classes <= kMiouFastClasses)
        miou_kernel<PredT, FROM_LOGITS, P, true><<<grid, kMiouThreads, (size_t)num_classes * kMiouThreads * sizeof(uint32_t), st>>>(
            logits, pred, target, n, c, hw, num_classes, counts);
    else
        miou_kernel<PredT, FROM_LOGITS, P, false><<<grid, kMiouThreads, 0, st>>>(logits, pred, target, n, c, hw, num_classes, counts);
    return launch_status();
}

extern "C" int mspl_miou_from_logits(const float* logits, const int64_t* target, int64_t n, int c, int64_t pixels_per_image,
                                     int num_classes, unsigned long long* counts, void* stream) {
    if (!logits || !target || !counts || n < 0 || c < 1 || c > kMiouMaxClasses || pixels_per_image < 1) return MSPL_ERR_BAD_ARG;
    if (num_classes < 1 || num_classes > kMiouMaxClasses - 1) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(logits, 4) || !aligned_to(target, 8) || !aligned_to(counts, 8)) return MSPL_ERR_ALIGN;
    if (n == 0) return MSPL_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pick_vec(pixels_per_image, {logits}) == 4)
        return miou_launch<uint8_t, true, 4>(logits, nullptr, target, n, c, pixels_per_image, num_classes, counts, st);
    return miou_launch<uint8_t, true, 1>(logits, nullptr, target, n, c, pixels_per_image, num_classes, counts, st);
}

extern "C" int mspl_miou_from_labels(const void* pred, int pred_is_int64, const int64_t* target, int64_t num_pixels, int num_classes,
                                     unsigned long long* counts, void* stream) {
    if (!pred || !target || !counts || num_pixels < 0) return MSPL_ERR_BAD_ARG;
    if (num_classes < 1 || num_classes > kMiouMaxClasses - 1) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(target, 8) || !aligned_to(counts, 8) || (pred_is_int64 && !aligned_to(pred, 8))) return MSPL_ERR_ALIGN;
    if (num_pixels == 0) return MSPL_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pred_is_int64)
        return miou_launch<int64_t, false, 1>(nullptr, static_cast<const int64_t*>(pred), target, 1, 1, num_pixels, num_classes, counts, st);
    return miou_launch<uint8_t, false, 1>(nullptr, static_cast<const uint8_t*>(pred), target, 1, 1, num_pixels, num_classes, counts, st);
}
