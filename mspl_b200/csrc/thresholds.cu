// K2 (class-balanced thresholds: exact per-class order statistics without a sort) and K3 (threshold / ignore-mask
// application).  [NEW] stages: the reference only keeps the CBST/CRST flags (uest_seg_multi_os.py:88-107, 216-219); their
// definition is SURVEY.md section 8 A4'' (restated in DESIGN.md).
//
// Two protocols over the same definition:
//  * bracketed (production): K1 already accumulated a LINEAR 2,048-bin histogram of conf per class.  bracket_select finds,
//    per class, the bin holding the j-th largest conf; ONE pass over (label, conf) then settles every pixel outside that
//    bin (keep / ignore), writes the final map and appends the few pixels inside it to a candidate list; a 3-pass radix
//    select over the candidates alone gives the exact threshold and cand_apply patches their labels.  6 B/pixel.
//  * generic radix (mspl_radix_*): 3 full passes over (label, conf) on the order-preserving key + mspl_apply_thresholds,
//    16 B/pixel; kept as the independent cross-check of the bracketed protocol and for callers with their own state.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mspl {

struct RadixState {            // one per target class, caller-zeroed before pass 0
    unsigned long long rank;   // 1-based rank from the top still to be resolved inside the current prefix
    unsigned long long count;  // kept pixels of the class (n_k)
    unsigned long long kept_above;   // candidate passes: keys seen so far that are certainly above the threshold
    uint32_t prefix;           // key bits resolved so far
    uint32_t done;             // kInactive: takes no part in the passes (threshold final or never resolved);
                               // kFixedOne: threshold is 1.0 (floor(n_k*portion) == 0), the candidate passes only COUNT the
                               // keys >= key(1.0) by following that key's digits; 0: select by rank
};
constexpr uint32_t kInactive = 1, kFixedOne = 2;

constexpr int kHistThreads = 256;

// VEC consecutive (label, conf) pairs per load: VEC = 4 -> uchar4 + float4 (a warp reads 128 B of labels and 512 B of conf,
// both contiguous), VEC = 1 -> scalars.  Kernels keep kUnroll such groups in flight per thread, one block-stride apart.
#ifndef MSPL_THR_UNROLL
#define MSPL_THR_UNROLL 4
#endif
constexpr int kUnroll = MSPL_THR_UNROLL;

template <int VEC>
MSPL_DEVINL void load_label_conf(const uint8_t* __restrict__ label, const float* __restrict__ conf, int64_t i0, uint8_t (&l)[VEC],
                                 float (&c)[VEC]) {
    if (VEC == 4) {
        const uchar4 lv = __ldcs(reinterpret_cast<const uchar4*>(label + i0));
        const float4 cv = __ldcs(reinterpret_cast<const float4*>(conf + i0));
        l[0] = lv.x; l[1 % VEC] = lv.y; l[2 % VEC] = lv.z; l[3 % VEC] = lv.w;
        c[0] = cv.x; c[1 % VEC] = cv.y; c[2 % VEC] = cv.z; c[3 % VEC] = cv.w;
    } else {
        l[0] = label[i0];
        c[0] = conf[i0];
    }
}

template <int VEC>
MSPL_DEVINL void store_bytes(uint8_t* __restrict__ dst, int64_t i0, const uint8_t (&v)[VEC]) {
    if (VEC == 4) *reinterpret_cast<uchar4*>(dst + i0) = make_uchar4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
    else dst[i0] = v[0];
}

constexpr int kLinearPass = 3;     // PASS value of the linear-bin histogram (conf_bin) that starts the bracketed protocol

template <int PASS> MSPL_DEVINL uint32_t pass_digit(uint32_t key, float conf) {
    return PASS == kLinearPass ? conf_bin(conf) : radix_digit(key, PASS);
}
template <int PASS> MSPL_DEVINL uint32_t pass_prefix(uint32_t key) { return PASS == kLinearPass ? 0u : radix_prefix(key, PASS); }

// Histogram of the PASS-th digit of the conf keys whose higher bits equal the class's resolved prefix
// (PASS == kLinearPass: the linear conf_bin histogram of every pixel of classes [0,K); `state` unused).
// Instruction-lean inner loop (5 B/pixel leaves ~25 issue slots per pixel at full HBM rate): one shared-memory lookup
// per pixel returns the class's prefix, or an impossible value for classes that are done / out of range.
template <int VEC, int PASS>
__global__ void __launch_bounds__(kHistThreads) radix_hist_kernel(const uint8_t* __restrict__ label, const float* __restrict__ conf,
                                                                  int64_t npix, int64_t hw, int K,
                                                                  const RadixState* __restrict__ state,
                                                                  unsigned long long* __restrict__ hist, int ds_rate) {
    extern __shared__ uint32_t s_hist[];
    __shared__ uint32_t s_prefix[MSPL_MAX_CLASSES + 1];
    const int nbins = K * MSPL_RADIX_BINS;
    for (int i = threadIdx.x; i < nbins; i += kHistThreads) s_hist[i] = 0;
    if (threadIdx.x <= MSPL_MAX_CLASSES) {
        const int k = threadIdx.x;
        if (PASS == kLinearPass) s_prefix[k] = k < K ? 0u : 0xffffffffu;
        else s_prefix[k] = (k < K && state[k].done != kInactive) ? state[k].prefix : 0xffffffffu;    // no key prefix has all 32 bits set
    }
    __syncthreads();
    // conf == +0 is by far the most common duplicate (every ignore-labelled pixel of the vote policies): those are counted
    // in registers (8 bits per class, spilled before overflow) instead of hammering one shared-memory address
    constexpr uint32_t zero_key = 0x80000000u;       // float_to_key(+0.f)
    unsigned long long zpacked = 0;
    uint32_t zcnt[MSPL_MAX_CLASSES] = {};
    int pending = 0;
    const int64_t n_groups = (npix + VEC - 1) / VEC;
    for (int64_t g0 = blockIdx.x * (int64_t)(kHistThreads * kUnroll) + threadIdx.x; g0 < n_groups;
         g0 += (int64_t)gridDim.x * kHistThreads * kUnroll) {
        uint8_t l[kUnroll][VEC];
        float c[kUnroll][VEC];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t g = g0 + u * kHistThreads;
            if (g < n_groups) load_label_conf<VEC>(label, conf, g * VEC, l[u], c[u]);
            else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) l[u][v] = 255;
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const uint32_t lab = min((uint32_t)l[u][v], (uint32_t)MSPL_MAX_CLASSES);
                const uint32_t key = float_to_key(c[u][v]);
                bool match = pass_prefix<PASS>(key) == s_prefix[lab];
                if (ds_rate > 1) match = match && (((g0 + u * kHistThreads) * VEC + v) % hw) % ds_rate == 0;
                if (match) {
                    if (key == zero_key) zpacked += 1ull << (8 * lab);
                    else atomicAdd(&s_hist[lab * MSPL_RADIX_BINS + pass_digit<PASS>(key, c[u][v])], 1u);
                }
            }
        }
        if ((pending += VEC * kUnroll) > 255 - VEC * kUnroll) {
#pragma unroll
            for (int k = 0; k < MSPL_MAX_CLASSES; ++k) zcnt[k] += (uint32_t)(zpacked >> (8 * k)) & 0xffu;
            zpacked = 0;
            pending = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < MSPL_MAX_CLASSES; ++k) {
        zcnt[k] += (uint32_t)(zpacked >> (8 * k)) & 0xffu;
        const uint32_t w = __reduce_add_sync(0xffffffffu, zcnt[k]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_hist[k * MSPL_RADIX_BINS + pass_digit<PASS>(zero_key, 0.f)], w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += kHistThreads)
        if (s_hist[i]) atomicAdd(hist + i, (unsigned long long)s_hist[i]);
}

// Suffix scan helper of the select kernels: the CTA's 256 threads each own nb/256 consecutive bins of s_h (already in shared
// memory); returns through (above, mine, total) the count in the bins above this thread's range, in its range, and overall.
MSPL_DEVINL void suffix_counts(const unsigned long long* s_h, int per, unsigned long long* s_warp, unsigned long long& above,
                               unsigned long long& mine, unsigned long long& total) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    mine = 0;
    for (int i = 0; i < per; ++i) mine += s_h[t * per + i];
    unsigned long long incl = mine;                       // inclusive suffix sum inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += v;
    }
    if (lane == 0) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long higher = 0;
    total = 0;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) {
        total += s_warp[w2];
        if (w2 > warp) higher += s_warp[w2];
    }
    above = higher + incl - mine;
}

// One CTA (256 threads) per class: locate the bin holding the rank-th largest key, extend the prefix, zero the histogram
// row.  INIT: derive the rank from the row total (the first pass of the generic protocol).
template <bool INIT>
__global__ void __launch_bounds__(256) radix_select_kernel(unsigned long long* __restrict__ hist, int pass, double portion,
                                                           RadixState* __restrict__ state, float* __restrict__ thresh,
                                                           unsigned long long* __restrict__ kept_count,
                                                           unsigned long long* __restrict__ final_hist, int ignore) {
    __shared__ unsigned long long s_h[MSPL_RADIX_BINS];
    __shared__ unsigned long long s_warp[8];
    __shared__ RadixState s_st;
    const int k = blockIdx.x;
    unsigned long long* h = hist + (size_t)k * MSPL_RADIX_BINS;
    const int nb = pass == 2 ? 1024 : MSPL_RADIX_BINS;
    const int bits = pass == 2 ? 10 : 11;
    const int per = nb / 256;
    for (int i = threadIdx.x; i < MSPL_RADIX_BINS; i += 256) {
        s_h[i] = h[i];
        h[i] = 0;
    }
    if (threadIdx.x == 0) s_st = state[k];
    __syncthreads();
    unsigned long long above, mine, total;
    suffix_counts(s_h, per, s_warp, above, mine, total);
    RadixState st = s_st;
    if (INIT) {
        st.count = total;
        unsigned long long j = (unsigned long long)((double)total * portion);   // floor(n_k * p), as int(n*p)
        if (j > total) j = total;
        st.rank = j;
        st.prefix = 0;
        st.kept_above = 0;
        st.done = (j == 0) ? kInactive : 0;
        if (threadIdx.x == 0) {
            if (st.done) {
                thresh[k] = 1.0f;
                state[k] = st;
            }
            if (kept_count) kept_count[k] = total;
        }
    }
    if (st.done == kInactive) return;
    // the digit this pass settles: by rank, or (kFixedOne) the digit of key(1.0); exactly one thread owns it
    const bool fixed = st.done == kFixedOne;
    const int fixed_d = (int)radix_digit(float_to_key(1.0f), pass);
    const bool owner = fixed ? (fixed_d / per == (int)threadIdx.x) : (above < st.rank && st.rank <= above + mine);
    if (owner) {
        unsigned long long acc = above;
        int d = threadIdx.x * per + per - 1;
        for (; d > threadIdx.x * per; --d) {
            if (fixed ? d == fixed_d : acc + s_h[d] >= st.rank) break;
            acc += s_h[d];
        }
        if (!fixed) st.rank -= acc;
        st.kept_above += acc;
        st.prefix = (st.prefix << bits) | (uint32_t)d;
        state[k] = st;
        if (pass == 2) {
            if (!fixed) thresh[k] = key_to_float(st.prefix);
            // the candidates that reach the threshold: everything above the selected key plus its duplicates.  They had been
            // counted as ignored; with the histograms all-reduced this patch is the GLOBAL one, identical on every rank.
            const unsigned long long n = st.kept_above + s_h[d];
            if (final_hist && n) {
                atomicAdd(final_hist + k, n);
                if (ignore >= 0) atomicAdd(final_hist + ignore, 0ull - n);
            }
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(256) apply_thresholds_kernel(const uint8_t* __restrict__ label, const float* __restrict__ conf,
                                                               const float* __restrict__ thresh, int64_t npix, int K, int ignore,
                                                               uint8_t* __restrict__ final_label, uint8_t* __restrict__ ignore_mask,
                                                               unsigned long long* __restrict__ final_hist) {
    __shared__ float s_thresh[MSPL_MAX_CLASSES + 1];    // +inf for the ignore class and for labels outside [0,K): never kept
    __shared__ uint32_t s_cls[MSPL_MAX_CLASSES];
    if (threadIdx.x <= MSPL_MAX_CLASSES) {
        const int k = threadIdx.x;
        s_thresh[k] = (k < K && k != ignore) ? thresh[k] : INFINITY;
        if (k < MSPL_MAX_CLASSES) s_cls[k] = 0;
    }
    __syncthreads();
    uint32_t cnt[MSPL_MAX_CLASSES] = {};
    unsigned long long packed = 0;          // 8 bits per class, spilled into cnt[] before it can overflow
    int pending = 0;
    const int64_t n_groups = (npix + VEC - 1) / VEC;
    for (int64_t g0 = blockIdx.x * (int64_t)(256 * kUnroll) + threadIdx.x; g0 < n_groups; g0 += (int64_t)gridDim.x * 256 * kUnroll) {
        uint8_t l[kUnroll][VEC];
        float c[kUnroll][VEC];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t g = g0 + u * 256;
            if (g < n_groups) load_label_conf<VEC>(label, conf, g * VEC, l[u], c[u]);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t g = g0 + u * 256;
            if (g >= n_groups) break;
            uint8_t f[VEC], mk[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const bool keep = c[u][v] >= s_thresh[min((uint32_t)l[u][v], (uint32_t)MSPL_MAX_CLASSES)];
                f[v] = keep ? l[u][v] : (uint8_t)ignore;
                mk[v] = keep ? 0 : 1;
                packed += 1ull << (8 * f[v]);
            }
            store_bytes<VEC>(final_label, g * VEC, f);
            if (ignore_mask) store_bytes<VEC>(ignore_mask, g * VEC, mk);
        }
        if ((pending += VEC * kUnroll) > 255 - VEC * kUnroll) {
#pragma unroll
            for (int k = 0; k < MSPL_MAX_CLASSES; ++k) cnt[k] += (uint32_t)(packed >> (8 * k)) & 0xffu;
            packed = 0;
            pending = 0;
        }
    }
    if (final_hist) {
#pragma unroll
        for (int k = 0; k < MSPL_MAX_CLASSES; ++k) {
            cnt[k] += (uint32_t)(packed >> (8 * k)) & 0xffu;
            const uint32_t w = __reduce_add_sync(0xffffffffu, cnt[k]);
            if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_cls[k], w);
        }
        __syncthreads();
        if (threadIdx.x < K && s_cls[threadIdx.x]) atomicAdd(final_hist + threadIdx.x, (unsigned long long)s_cls[threadIdx.x]);
    }
}

// ---- bracketed protocol ---------------------------------------------------------------------------------------------------
// One CTA per class on the (all-reduced) linear histogram: j = floor(n_k * portion); the bin b holding the j-th largest conf
// brackets the threshold: conf >= hi = (b+1)/2048 is certainly kept, conf < lo = b/2048 certainly dropped, and the pixels
// in between are the candidates among which the (j - #above)-th largest is the threshold.  Leaves `state` ready for the
// candidate radix passes and zeroes the histogram row.
__global__ void __launch_bounds__(256) bracket_select_kernel(unsigned long long* __restrict__ hist, double portion, int ignore,
                                                             RadixState* __restrict__ state, float2* __restrict__ bracket,
                                                             float* __restrict__ thresh, unsigned long long* __restrict__ kept_count,
                                                             const unsigned long long* __restrict__ local_hist,
                                                             unsigned long long* __restrict__ final_hist) {
    __shared__ unsigned long long s_h[MSPL_RADIX_BINS];
    __shared__ unsigned long long s_warp[8];
    __shared__ int s_bin;
    const int k = blockIdx.x;
    unsigned long long* h = hist + (size_t)k * MSPL_RADIX_BINS;
    constexpr int per = MSPL_RADIX_BINS / 256;
    for (int i = threadIdx.x; i < MSPL_RADIX_BINS; i += 256) {
        s_h[i] = h[i];
        h[i] = 0;
    }
    if (threadIdx.x == 0) s_bin = MSPL_RADIX_BINS;      // "no bin": nothing of this class is kept outright
    __syncthreads();
    unsigned long long above, mine, total;
    suffix_counts(s_h, per, s_warp, above, mine, total);
    unsigned long long j = (unsigned long long)((double)total * portion);       // floor(n_k * p), as int(n*p)
    if (j > total) j = total;
    if (k == ignore) {
        // never selected: threshold left unresolved (+inf) -- with the vote policies all its pixels share conf == 0 and would
        // all be candidates
        if (threadIdx.x == 0) {
            RadixState st;
            st.rank = 0; st.count = total; st.kept_above = 0; st.prefix = 0; st.done = kInactive;
            state[k] = st;
            thresh[k] = INFINITY;
            bracket[k] = make_float2(INFINITY, INFINITY);
            if (kept_count) kept_count[k] = total;
        }
    } else if (j == 0) {
        // threshold 1.0, final already: the pixels that reach it sit in the top bin, which becomes the bracket so that the
        // candidate patch keeps exactly those with conf >= 1.0
        if (threadIdx.x == 0) {
            RadixState st;
            st.rank = 0; st.count = total; st.kept_above = 0; st.prefix = 0; st.done = kFixedOne;   // passes only count conf >= 1.0
            state[k] = st;
            thresh[k] = 1.0f;
            bracket[k] = make_float2((float)(MSPL_RADIX_BINS - 1) * (1.0f / MSPL_RADIX_BINS), INFINITY);
            if (kept_count) kept_count[k] = total;
            s_bin = MSPL_RADIX_BINS - 1;
        }
    } else if (above < j && j <= above + mine) {        // exactly one thread
        unsigned long long acc = above;
        int b = threadIdx.x * per + per - 1;
        for (; b > threadIdx.x * per; --b) {
            if (acc + s_h[b] >= j) break;
            acc += s_h[b];
        }
        RadixState st;
        st.rank = j - acc; st.count = total; st.kept_above = 0; st.prefix = 0; st.done = 0;
        state[k] = st;
        bracket[k] = make_float2(b == 0 ? -INFINITY : (float)b * (1.0f / MSPL_RADIX_BINS),
                                 b == MSPL_RADIX_BINS - 1 ? INFINITY : (float)(b + 1) * (1.0f / MSPL_RADIX_BINS));
        if (kept_count) kept_count[k] = total;
        s_bin = b;
    }
    if (!final_hist) return;
    // final class counts straight from the histogram of THIS rank's pixels (valid when it was accumulated with ds_rate 1 over
    // exactly the pixels the classify pass will see): everything above the bracket bin keeps its label, the rest of the
    // class is ignored until cand_apply patches the candidates that reach the threshold
    __syncthreads();
    const int b = s_bin;
    const unsigned long long* lh = local_hist ? local_hist + (size_t)k * MSPL_RADIX_BINS : nullptr;
    unsigned long long n_all = 0, n_keep = 0;
    for (int i = threadIdx.x; i < MSPL_RADIX_BINS; i += 256) {
        const unsigned long long c = lh ? lh[i] : s_h[i];
        n_all += c;
        if (i > b) n_keep += c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_all += __shfl_xor_sync(0xffffffffu, n_all, o);
        n_keep += __shfl_xor_sync(0xffffffffu, n_keep, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_keep) atomicAdd(final_hist + k, n_keep);
        if (ignore >= 0 && n_all - n_keep) atomicAdd(final_hist + ignore, n_all - n_keep);
    }
}

#ifndef MSPL_CLASSIFY_UNR
#define MSPL_CLASSIFY_UNR 6        // (label word, conf float4) groups a thread of the classify pass keeps in flight (4: 0.293 ms, 6: 0.285 ms, 8: 0.294 ms at 245.76 Mpix)
#endif
#ifndef MSPL_CLASSIFY_MINB
#define MSPL_CLASSIFY_MINB 4
#endif
constexpr int kCandBuf = 256 + 128 * (MSPL_CLASSIFY_UNR > 4 ? MSPL_CLASSIFY_UNR : 4);   // per-warp staging entries (768 at UNR 4)

// The one full pass of the bracketed protocol: settle every pixel outside its class's bracket, stage the candidates.
// Candidates get the ignore label for now (and count as ignored); cand_apply patches the ones that reach the threshold.
// Appends are staged per warp in shared memory and flushed with one global atomic per ~kCandFlushAt entries, so the list
// costs nothing when candidates are rare and stays correct (just slower) when every pixel is one.
// Generic form (any alignment, VEC = 1 or 4, counts the final classes itself); see bracket_classify_words_kernel below.
template <int VEC>
__global__ void __launch_bounds__(256, 4) bracket_classify_kernel(const uint8_t* __restrict__ label, const float* __restrict__ conf,
                                                               const float2* __restrict__ bracket, int64_t npix, int K, int ignore,
                                                               uint8_t* __restrict__ final_label, uint8_t* __restrict__ ignore_mask,
                                                               unsigned long long* __restrict__ final_hist,
                                                               uint32_t* __restrict__ cand_index,
                                                               unsigned long long* __restrict__ cand_count) {
    __shared__ float2 s_br[MSPL_MAX_CLASSES + 1];       // (+inf,+inf) for the ignore class and labels outside [0,K): never kept
    __shared__ uint32_t s_cls[MSPL_MAX_CLASSES];
    __shared__ uint32_t s_buf[8][kCandBuf];
    __shared__ uint32_t s_fill[8];
    constexpr int kCandFlushAt = kCandBuf - 32 * VEC * kUnroll;    // a warp adds at most 32 lanes x VEC x kUnroll per iteration
    static_assert(kCandFlushAt > 0, "per-warp candidate staging must hold one iteration's worst case");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x <= MSPL_MAX_CLASSES) {
        const int k = threadIdx.x;
        s_br[k] = (k < K && k != ignore) ? bracket[k] : make_float2(INFINITY, INFINITY);
        if (k < MSPL_MAX_CLASSES) s_cls[k] = 0;
    }
    if (threadIdx.x < 8) s_fill[threadIdx.x] = 0;
    __syncthreads();
    uint32_t cnt[MSPL_MAX_CLASSES] = {};
    unsigned long long packed = 0;          // 8 bits per class, spilled into cnt[] before it can overflow
    int pending = 0;
    auto flush = [&](uint32_t fill) {       // whole warp; fill is warp-uniform
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cand_count, (unsigned long long)fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t i = lane; i < fill; i += 32) cand_index[base + i] = s_buf[warp][i];
        __syncwarp();
        if (lane == 0) s_fill[warp] = 0;
        __syncwarp();
    };
    const int64_t n_groups = (npix + VEC - 1) / VEC;
    // warp-uniform trip count (the staging flush is a warp-collective): iterate on the warp's first group
    for (int64_t w0 = blockIdx.x * (int64_t)(256 * kUnroll) + warp * 32; w0 < n_groups; w0 += (int64_t)gridDim.x * 256 * kUnroll) {
        const int64_t g0 = w0 + lane;
        uint8_t l[kUnroll][VEC];
        float c[kUnroll][VEC];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t g = g0 + u * 256;
            if (g < n_groups) load_label_conf<VEC>(label, conf, g * VEC, l[u], c[u]);
            else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { l[u][v] = 255; c[u][v] = 0.f; }     // class "never kept, never a candidate"
            }
        }
        uint32_t cmask = 0;             // bit u*VEC+v: pixel (u, v) lies inside its class's bracket
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t g = g0 + u * 256;
            uint8_t f[VEC], mk[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float2 br = s_br[min((uint32_t)l[u][v], (uint32_t)MSPL_MAX_CLASSES)];
                const bool keep = c[u][v] >= br.y;
                cmask |= (uint32_t)(c[u][v] >= br.x && !keep) << (u * VEC + v);
                f[v] = keep ? l[u][v] : (uint8_t)ignore;
                mk[v] = keep ? 0 : 1;
                if (final_hist && g < n_groups) packed += 1ull << (8 * f[v]);
            }
            if (g < n_groups) {
                if (final_label) store_bytes<VEC>(final_label, g * VEC, f);
                if (ignore_mask) store_bytes<VEC>(ignore_mask, g * VEC, mk);
            }
        }
        while (cmask) {                 // rare: one divergent region per iteration instead of one per pixel
            const int bit = __ffs(cmask) - 1;
            cmask &= cmask - 1;
            const uint32_t slot = atomicAdd(&s_fill[warp], 1u);
            s_buf[warp][slot] = (uint32_t)((g0 + (bit / VEC) * 256) * VEC + bit % VEC);
        }
        if ((pending += VEC * kUnroll) > 255 - VEC * kUnroll) {
#pragma unroll
            for (int k = 0; k < MSPL_MAX_CLASSES; ++k) cnt[k] += (uint32_t)(packed >> (8 * k)) & 0xffu;
            packed = 0;
            pending = 0;
        }
        __syncwarp();
        const uint32_t fill = s_fill[warp];
        if (fill > (uint32_t)kCandFlushAt) flush(fill);
    }
    __syncwarp();
    const uint32_t fill = s_fill[warp];
    if (fill) flush(fill);
    if (final_hist) {
#pragma unroll
        for (int k = 0; k < MSPL_MAX_CLASSES; ++k) {
            cnt[k] += (uint32_t)(packed >> (8 * k)) & 0xffu;
            const uint32_t w = __reduce_add_sync(0xffffffffu, cnt[k]);
            if (lane == 0 && w) atomicAdd(&s_cls[k], w);
        }
        __syncthreads();
        if (threadIdx.x < K && s_cls[threadIdx.x]) atomicAdd(final_hist + threadIdx.x, (unsigned long long)s_cls[threadIdx.x]);
    }
}

// Production form of the classify pass for 4-aligned maps: labels stay packed four to a 32-bit word, the per-class bracket
// comes from a 256-entry shared table indexed by the label byte (no clamping), the final word is one bitwise select, and the
// class counts are left to bracket_select / cand_apply (COUNT = false) -- ~3x fewer integer instructions per pixel than the
// generic kernel above, which matters because these passes are bound by the half-rate integer pipe, not by HBM.
// Pixels of labels outside [0,K) are written as ignored.
template <int UNR, bool COUNT>
__global__ void __launch_bounds__(256, MSPL_CLASSIFY_MINB) bracket_classify_words_kernel(const uint32_t* __restrict__ label4,
                                                                      const float4* __restrict__ conf4,
                                                                      const float2* __restrict__ bracket, uint32_t n_groups, int K,
                                                                      int ignore, uint32_t* __restrict__ final4,
                                                                      uint32_t* __restrict__ mask4,
                                                                      unsigned long long* __restrict__ final_hist,
                                                                      uint32_t* __restrict__ cand_index,
                                                                      unsigned long long* __restrict__ cand_count) {
    constexpr int kCandFlushAt = kCandBuf - 32 * 4 * UNR;
    static_assert(kCandFlushAt > 0, "per-warp candidate staging must hold one iteration's worst case");
    __shared__ float2 s_tab[256];
    __shared__ uint32_t s_cls[MSPL_MAX_CLASSES];
    __shared__ uint32_t s_buf[8][kCandBuf];
    __shared__ uint32_t s_fill[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const int k = threadIdx.x;
        s_tab[k] = (k < K && k != ignore) ? bracket[k] : make_float2(INFINITY, INFINITY);
        if (k < MSPL_MAX_CLASSES) s_cls[k] = 0;
        if (k < 8) s_fill[k] = 0;
    }
    __syncthreads();
    const uint32_t ign4 = (uint32_t)(ignore & 0xff) * 0x01010101u;
    uint32_t cnt[MSPL_MAX_CLASSES] = {};
    unsigned long long packed = 0;
    int pending = 0;
    auto flush = [&](uint32_t fill) {       // whole warp; fill is warp-uniform
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cand_count, (unsigned long long)fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (uint32_t i = lane; i < fill; i += 32) cand_index[base + i] = s_buf[warp][i];
        __syncwarp();
        if (lane == 0) s_fill[warp] = 0;
        __syncwarp();
    };
    // group counts fit 32 bits (num_pixels < 2^32); the loop index runs in 64 bits only where it could wrap
    for (uint64_t w0 = (uint64_t)blockIdx.x * (256 * UNR) + warp * 32; w0 < n_groups; w0 += (uint64_t)gridDim.x * 256 * UNR) {
        const uint32_t g0 = (uint32_t)w0 + lane;
        uint32_t lw[UNR];
        float4 cv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const uint32_t g = g0 + u * 256;
            if (g < n_groups) {
                lw[u] = __ldcs(label4 + g);
                cv[u] = __ldcs(conf4 + g);
            } else {
                lw[u] = 0xffffffffu;        // table entry 255: never kept, never a candidate
                cv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        uint32_t cmask = 0;             // bit 4u+v: pixel (u, v) lies inside its class's bracket
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const uint32_t g = g0 + u * 256;
            const float c[4] = {cv[u].x, cv[u].y, cv[u].z, cv[u].w};
            uint32_t m = 0;             // 0xff in the bytes that keep their label
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const float2 br = s_tab[__byte_perm(lw[u], 0, 0x4440 + v)];
                const bool keep = c[v] >= br.y;
                m |= keep ? (0xffu << (8 * v)) : 0u;
                cmask |= (c[v] >= br.x && !keep) ? (1u << (4 * u + v)) : 0u;
            }
            const uint32_t fw = (lw[u] & m) | (ign4 & ~m);
            if (g < n_groups) {
                if (final4) __stcs(final4 + g, fw);
                if (mask4) __stcs(mask4 + g, ~m & 0x01010101u);
                if (COUNT) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) packed += 1ull << (8 * ((fw >> (8 * v)) & 7u));
                }
            }
        }
        while (cmask) {                 // rare: one divergent region per iteration instead of one per pixel
            const int bit = __ffs(cmask) - 1;
            cmask &= cmask - 1;
            const uint32_t slot = atomicAdd(&s_fill[warp], 1u);
            s_buf[warp][slot] = (g0 + (bit >> 2) * 256) * 4 + (bit & 3);
        }
        if (COUNT && (pending += 4 * UNR) > 255 - 4 * UNR) {
#pragma unroll
            for (int k = 0; k < MSPL_MAX_CLASSES; ++k) cnt[k] += (uint32_t)(packed >> (8 * k)) & 0xffu;
            packed = 0;
            pending = 0;
        }
        __syncwarp();
        const uint32_t fill = s_fill[warp];
        if (fill > (uint32_t)kCandFlushAt) flush(fill);
    }
    __syncwarp();
    const uint32_t fill = s_fill[warp];
    if (fill) flush(fill);
    if (COUNT) {
#pragma unroll
        for (int k = 0; k < MSPL_MAX_CLASSES; ++k) {
            cnt[k] += (uint32_t)(packed >> (8 * k)) & 0xffu;
            const uint32_t w = __reduce_add_sync(0xffffffffu, cnt[k]);
            if (lane == 0 && w) atomicAdd(&s_cls[k], w);
        }
        __syncthreads();
        if (threadIdx.x < K && s_cls[threadIdx.x]) atomicAdd(final_hist + threadIdx.x, (unsigned long long)s_cls[threadIdx.x]);
    }
}

// Radix pass over the candidate list only (gathers label/conf through the index): same digits, prefixes and state as
// radix_hist_kernel, so radix_select_kernel<false> continues from the rank bracket_select left in `state`.
template <int PASS>
__global__ void __launch_bounds__(kHistThreads) cand_hist_kernel(const uint8_t* __restrict__ label, const float* __restrict__ conf,
                                                                 const uint32_t* __restrict__ cand_index,
                                                                 const unsigned long long* __restrict__ cand_count, int64_t hw, int K,
                                                                 const RadixState* __restrict__ state,
                                                                 unsigned long long* __restrict__ hist, int ds_rate) {
    extern __shared__ uint32_t s_hist[];
    __shared__ uint32_t s_prefix[MSPL_MAX_CLASSES + 1];
    const unsigned long long n = *cand_count;
    if ((unsigned long long)blockIdx.x * kHistThreads >= n) return;
    const int nbins = K * MSPL_RADIX_BINS;
    for (int i = threadIdx.x; i < nbins; i += kHistThreads) s_hist[i] = 0;
    if (threadIdx.x <= MSPL_MAX_CLASSES) {
        const int k = threadIdx.x;
        s_prefix[k] = (k < K && state[k].done != kInactive) ? state[k].prefix : 0xffffffffu;
    }
    __syncthreads();
    for (unsigned long long i = (unsigned long long)blockIdx.x * kHistThreads + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * kHistThreads) {
        const uint32_t idx = cand_index[i];
        const uint32_t lab = min((uint32_t)label[idx], (uint32_t)MSPL_MAX_CLASSES);
        const uint32_t key = float_to_key(conf[idx]);
        bool match = radix_prefix(key, PASS) == s_prefix[lab];
        if (ds_rate > 1) match = match && ((int64_t)idx % hw) % ds_rate == 0;
        if (match) atomicAdd(&s_hist[lab * MSPL_RADIX_BINS + radix_digit(key, PASS)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += kHistThreads)
        if (s_hist[i]) atomicAdd(hist + i, (unsigned long long)s_hist[i]);
}

// Patch the candidates that reach their class's threshold: they were written (and counted) as ignored by the classify pass.
__global__ void __launch_bounds__(256) cand_apply_kernel(const uint8_t* __restrict__ label, const float* __restrict__ conf,
                                                         const float* __restrict__ thresh, const uint32_t* __restrict__ cand_index,
                                                         const unsigned long long* __restrict__ cand_count, int K, int ignore,
                                                         uint8_t* __restrict__ final_label, uint8_t* __restrict__ ignore_mask,
                                                         unsigned long long* __restrict__ final_hist) {
    __shared__ float s_thresh[MSPL_MAX_CLASSES + 1];
    __shared__ uint32_t s_cls[MSPL_MAX_CLASSES];
    const unsigned long long n = *cand_count;
    if ((unsigned long long)blockIdx.x * 256 >= n) return;
    if (threadIdx.x <= MSPL_MAX_CLASSES) {
        const int k = threadIdx.x;
        s_thresh[k] = (k < K && k != ignore) ? thresh[k] : INFINITY;
        if (k < MSPL_MAX_CLASSES) s_cls[k] = 0;
    }
    __syncthreads();
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256) {
        const uint32_t idx = cand_index[i];
        const uint32_t lab = min((uint32_t)label[idx], (uint32_t)MSPL_MAX_CLASSES);
        if (conf[idx] >= s_thresh[lab]) {
            if (final_label) final_label[idx] = (uint8_t)lab;
            if (ignore_mask) ignore_mask[idx] = 0;
            atomicAdd(&s_cls[lab], 1u);
        }
    }
    __syncthreads();
    if (final_hist && threadIdx.x < K && s_cls[threadIdx.x]) {
        atomicAdd(final_hist + threadIdx.x, (unsigned long long)s_cls[threadIdx.x]);
        if (ignore < K) atomicAdd(final_hist + ignore, 0ull - (unsigned long long)s_cls[threadIdx.x]);   // had been counted as ignored
    }
}

// ---- single-rank tail of the bracketed protocol in ONE launch ---------------------------------------------------------------
// Replaces 3 x (cand_hist + cand_select) + cand_apply (7 launches of ~10 us for a few ten-thousand candidates) when no
// all-reduce has to run between the passes.  One thread-block CLUSTER of kResolveCtas CTAs per class: each CTA gathers its
// slice of the candidate list once, keeps the (key, index) pairs of its class in shared memory, and per radix pass
//   local shared-memory histogram  ->  added into cluster rank 0's total through distributed shared memory
//   ->  rank 0 selects the digit  ->  every CTA reads the verdict back over DSMEM
// with three cluster barriers per pass; then each CTA patches the labels of its own candidates.  A CTA whose slice holds more
// candidates of the class than its cache re-gathers them from global memory in every pass (correct, slower).
constexpr int kResolveCtas = 8;             // cluster size (portable maximum)
constexpr int kResolveThreads = 1024;
constexpr int kResolveCap = 12288;          // cached (key, index) pairs per CTA: 96 KB

struct ResolveVerdict { uint32_t digit, count_at; unsigned long long above; };

__global__ void __launch_bounds__(kResolveThreads, 1) cand_resolve_kernel(const uint8_t* __restrict__ label, const float* __restrict__ conf,
                                                                          const uint32_t* __restrict__ cand_index,
                                                                          const unsigned long long* __restrict__ cand_count,
                                                                          int64_t hw, int ignore, int ds_rate,
                                                                          RadixState* __restrict__ state, float* __restrict__ thresh,
                                                                          uint8_t* __restrict__ final_label,
                                                                          uint8_t* __restrict__ ignore_mask,
                                                                          unsigned long long* __restrict__ final_hist) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(rs_smem);                  // this CTA's histogram of the current pass
    uint32_t* s_total = s_hist + MSPL_RADIX_BINS;                             // rank 0: the cluster's histogram
    uint32_t* s_key = s_total + MSPL_RADIX_BINS;
    uint32_t* s_idx = s_key + kResolveCap;
    __shared__ uint32_t s_fill, s_kept;
    __shared__ unsigned long long s_warp[kResolveThreads / 32];
    __shared__ ResolveVerdict s_verdict;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int k = blockIdx.x / kResolveCtas;
    const int t = threadIdx.x;
    const RadixState st0 = state[k];
    if (st0.done == kInactive) return;                    // uniform over the cluster: nothing to resolve for this class
    const unsigned long long n = *cand_count;
    // this CTA's slice of the candidate list
    const unsigned long long per = (n + kResolveCtas - 1) / kResolveCtas;
    const unsigned long long lo = per * rank < n ? per * rank : n, hi = lo + per < n ? lo + per : n;
    if (t == 0) { s_fill = 0; s_kept = 0; }
    __syncthreads();
    // gather once: candidates of class k in the slice
    for (unsigned long long i = lo + t; i < hi; i += kResolveThreads) {
        const uint32_t idx = cand_index[i];
        if (label[idx] == k) {
            const uint32_t slot = atomicAdd(&s_fill, 1u);
            if (slot < (uint32_t)kResolveCap) { s_key[slot] = float_to_key(conf[idx]); s_idx[slot] = idx; }
        }
    }
    __syncthreads();
    const uint32_t mine_n = s_fill;
    const bool cached = mine_n <= (uint32_t)kResolveCap;
    // visit (key, idx) of every candidate of class k in the slice
    auto for_each = [&](auto&& f) {
        if (cached) {
            for (uint32_t i = t; i < mine_n; i += kResolveThreads) f(s_key[i], s_idx[i]);
        } else {
            for (unsigned long long i = lo + t; i < hi; i += kResolveThreads) {
                const uint32_t idx = cand_index[i];
                if (label[idx] == k) f(float_to_key(conf[idx]), idx);
            }
        }
    };
    const bool fixed = st0.done == kFixedOne;
    const uint32_t key_one = float_to_key(1.0f);
    uint32_t prefix = 0;
    unsigned long long rank_left = st0.rank, kept_above = 0;
    for (int pass = 0; pass < MSPL_RADIX_PASSES; ++pass) {
        const int nb = pass == 2 ? 1024 : MSPL_RADIX_BINS;
        const int bits = pass == 2 ? 10 : 11;
        for (int i = t; i < MSPL_RADIX_BINS; i += kResolveThreads) { s_hist[i] = 0; s_total[i] = 0; }
        cluster.sync();                                   // rank 0's total is zeroed before anyone adds to it
        for_each([&](uint32_t key, uint32_t idx) {
            const bool part = ds_rate <= 1 || ((int64_t)idx % hw) % ds_rate == 0;
            if (part && radix_prefix(key, pass) == prefix) atomicAdd(&s_hist[radix_digit(key, pass)], 1u);
        });
        __syncthreads();
        uint32_t* total0 = cluster.map_shared_rank(s_total, 0);
        for (int i = t; i < nb; i += kResolveThreads)
            if (s_hist[i]) atomicAdd(total0 + i, s_hist[i]);
        cluster.sync();                                   // every CTA's counts have landed in rank 0
        if (rank == 0) {
            // suffix counts over the nb bins: thread t owns bins [t*per_t, (t+1)*per_t)
            const int per_t = nb / kResolveThreads;       // 2 or 1
            unsigned long long mine = 0;
            for (int i = 0; i < per_t; ++i) mine += s_total[t * per_t + i];
            unsigned long long incl = mine;
            const int lane = t & 31, warp = t >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_down_sync(0xffffffffu, incl, o);
                if (lane + o < 32) incl += v;
            }
            if (lane == 0) s_warp[warp] = incl;
            if (t == 0) s_verdict = ResolveVerdict{0u, 0u, 0ull};
            __syncthreads();
            unsigned long long higher = 0;
            for (int w2 = warp + 1; w2 < kResolveThreads / 32; ++w2) higher += s_warp[w2];
            const unsigned long long above = higher + incl - mine;
            const int fixed_d = (int)radix_digit(key_one, pass);
            const bool owner = fixed ? (fixed_d / per_t == t) : (above < rank_left && rank_left <= above + mine);
            if (owner) {
                unsigned long long acc = above;
                int d = t * per_t + per_t - 1;
                for (; d > t * per_t; --d) {
                    if (fixed ? d == fixed_d : acc + s_total[d] >= rank_left) break;
                    acc += s_total[d];
                }
                s_verdict = ResolveVerdict{(uint32_t)d, s_total[d], acc};
            }
        }
        cluster.sync();                                   // the verdict is published
        const ResolveVerdict v = *cluster.map_shared_rank(&s_verdict, 0);
        prefix = (prefix << bits) | v.digit;
        if (!fixed) rank_left -= v.above;
        kept_above += v.above;
    }
    const float th = fixed ? 1.0f : key_to_float(prefix);
    if (rank == 0 && t == 0) {
        RadixState st = st0;
        st.prefix = prefix; st.rank = rank_left; st.kept_above = kept_above;
        state[k] = st;
        if (!fixed) thresh[k] = th;
    }
    // patch: the candidates that reach the threshold get their label back (they were written and counted as ignored)
    uint32_t kept = 0;
    for_each([&](uint32_t key, uint32_t idx) {
        if (key_to_float(key) >= th) {
            if (final_label) final_label[idx] = (uint8_t)k;
            if (ignore_mask) ignore_mask[idx] = 0;
            ++kept;
        }
    });
    kept = __reduce_add_sync(0xffffffffu, kept);
    if ((t & 31) == 0 && kept) atomicAdd(&s_kept, kept);
    __syncthreads();
    if (t == 0 && final_hist && s_kept) {
        atomicAdd(final_hist + k, (unsigned long long)s_kept);
        if (ignore >= 0) atomicAdd(final_hist + ignore, 0ull - (unsigned long long)s_kept);
    }
    cluster.sync();                                       // no CTA may exit while others still read its shared memory
}

// Persistent grid of exactly the CTAs that can be resident at once (these kernels are latency-bound: a partial second
// wave would run at a fraction of the occupancy).
template <typename Kern>
static int64_t resident_grid(Kern kern, int64_t n_groups, int threads, size_t smem) {
    int dev = 0, sms = kNumSMs, per_sm = 1;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    const int64_t blocks = (n_groups + threads - 1) / threads;
    const int64_t cap = (int64_t)sms * per_sm;
    return blocks < cap ? (blocks < 1 ? 1 : blocks) : cap;
}

}  // namespace mspl

using namespace mspl;

extern "C" size_t mspl_radix_state_bytes(int num_target_classes) {
    return num_target_classes < 1 ? 0 : sizeof(RadixState) * (size_t)num_target_classes;
}

extern "C" int mspl_radix_hist_pass(const uint8_t* label, const float* conf, int64_t num_pixels, int64_t pixels_per_image,
                                    int num_target_classes, int pass, const void* state, unsigned long long* hist,
                                    int ds_rate, void* stream) {
    const int K = num_target_classes;
    if (!label || !conf || !state || !hist || num_pixels < 0 || pixels_per_image < 1) return MSPL_ERR_BAD_ARG;
    if (K < 1 || K > MSPL_MAX_CLASSES || pass < 0 || pass >= MSPL_RADIX_PASSES || ds_rate < 1) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(conf, 4) || !aligned_to(hist, 8) || !aligned_to(state, 8)) return MSPL_ERR_ALIGN;
    if (num_pixels == 0) return MSPL_OK;
    const size_t smem = sizeof(uint32_t) * (size_t)K * MSPL_RADIX_BINS;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int vec = (num_pixels % 4 == 0 && aligned_to(label, 4) && aligned_to(conf, 16)) ? 4 : 1;
    auto kern = vec == 4 ? (pass == 0 ? radix_hist_kernel<4, 0> : pass == 1 ? radix_hist_kernel<4, 1> : radix_hist_kernel<4, 2>)
                         : (pass == 0 ? radix_hist_kernel<1, 0> : pass == 1 ? radix_hist_kernel<1, 1> : radix_hist_kernel<1, 2>);
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return MSPL_ERR_CUDA;
    }
    const int64_t grid = resident_grid(kern, num_pixels / vec / kUnroll + 1, kHistThreads, smem);
    kern<<<(unsigned)grid, kHistThreads, smem, st>>>(label, conf, num_pixels, pixels_per_image, K,
                                                     static_cast<const RadixState*>(state), hist, ds_rate);
    return launch_status();
}

extern "C" int mspl_radix_select(unsigned long long* hist, int num_target_classes, int pass, double portion, void* state,
                                 float* thresh, unsigned long long* kept_count, void* stream) {
    const int K = num_target_classes;
    if (!hist || !state || !thresh || K < 1 || K > MSPL_MAX_CLASSES || pass < 0 || pass >= MSPL_RADIX_PASSES) return MSPL_ERR_BAD_ARG;
    if (!(portion >= 0.0)) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(hist, 8) || !aligned_to(state, 8) || !aligned_to(thresh, 4)) return MSPL_ERR_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pass == 0)
        radix_select_kernel<true><<<K, 256, 0, st>>>(hist, pass, portion, static_cast<RadixState*>(state), thresh, kept_count, nullptr, -1);
    else
        radix_select_kernel<false><<<K, 256, 0, st>>>(hist, pass, portion, static_cast<RadixState*>(state), thresh, kept_count, nullptr, -1);
    return launch_status();
}

extern "C" int mspl_apply_thresholds(const uint8_t* label, const float* conf, const float* thresh, int64_t num_pixels,
                                     int num_target_classes, int ignore_label, uint8_t* final_label, uint8_t* ignore_mask,
                                     unsigned long long* final_hist, void* stream) {
    const int K = num_target_classes;
    if (!label || !conf || !thresh || !final_label || num_pixels < 0) return MSPL_ERR_BAD_ARG;
    if (K < 1 || K > MSPL_MAX_CLASSES || ignore_label < 0 || ignore_label >= MSPL_MAX_CLASSES) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(conf, 4) || !aligned_to(thresh, 4) || (final_hist && !aligned_to(final_hist, 8))) return MSPL_ERR_ALIGN;
    if (num_pixels == 0) return MSPL_OK;
    auto ok = [&](size_t a) {
        return aligned_to(label, a) && aligned_to(conf, 16) && aligned_to(final_label, a) && (!ignore_mask || aligned_to(ignore_mask, a));
    };
    const int vec = (num_pixels % 4 == 0 && ok(4)) ? 4 : 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t grid = vec == 4 ? resident_grid(apply_thresholds_kernel<4>, num_pixels / vec / kUnroll + 1, 256, 0)
                                  : resident_grid(apply_thresholds_kernel<1>, num_pixels / vec / kUnroll + 1, 256, 0);
    if (vec == 4)
        apply_thresholds_kernel<4><<<(unsigned)grid, 256, 0, st>>>(label, conf, thresh, num_pixels, K, ignore_label, final_label,
                                                                  ignore_mask, final_hist);
    else
        apply_thresholds_kernel<1><<<(unsigned)grid, 256, 0, st>>>(label, conf, thresh, num_pixels, K, ignore_label, final_label,
                                                                  ignore_mask, final_hist);
    return launch_status();
}

// ---- bracketed protocol entry points ----------------------------------------------------------------------------------
extern "C" int mspl_conf_hist(const uint8_t* label, const float* conf, int64_t num_pixels, int64_t pixels_per_image,
                              int num_target_classes, unsigned long long* hist, int ds_rate, void* stream) {
    const int K = num_target_classes;
    if (!label || !conf || !hist || num_pixels < 0 || pixels_per_image < 1) return MSPL_ERR_BAD_ARG;
    if (K < 1 || K > MSPL_MAX_CLASSES || ds_rate < 1) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(conf, 4) || !aligned_to(hist, 8)) return MSPL_ERR_ALIGN;
    if (num_pixels == 0) return MSPL_OK;
    const size_t smem = sizeof(uint32_t) * (size_t)K * MSPL_RADIX_BINS;
    const int vec = (num_pixels % 4 == 0 && aligned_to(label, 4) && aligned_to(conf, 16)) ? 4 : 1;
    auto kern = vec == 4 ? radix_hist_kernel<4, kLinearPass> : radix_hist_kernel<1, kLinearPass>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return MSPL_ERR_CUDA;
    }
    const int64_t grid = resident_grid(kern, num_pixels / vec / kUnroll + 1, kHistThreads, smem);
    kern<<<(unsigned)grid, kHistThreads, smem, static_cast<cudaStream_t>(stream)>>>(label, conf, num_pixels, pixels_per_image, K,
                                                                                    nullptr, hist, ds_rate);
    return launch_status();
}

extern "C" int mspl_bracket_select(unsigned long long* hist, int num_target_classes, double portion, int ignore_label, void* state,
                                   float* bracket, float* thresh, unsigned long long* kept_count,
                                   const unsigned long long* local_hist, unsigned long long* final_hist, void* stream) {
    const int K = num_target_classes;
    if (!hist || !state || !bracket || !thresh || K < 1 || K > MSPL_MAX_CLASSES || !(portion >= 0.0)) return MSPL_ERR_BAD_ARG;
    if (ignore_label < -1 || ignore_label >= MSPL_MAX_CLASSES) return MSPL_ERR_BAD_ARG;
    if (final_hist && (ignore_label < 0 || ignore_label >= K)) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(hist, 8) || !aligned_to(state, 8) || !aligned_to(bracket, 8) || !aligned_to(thresh, 4) ||
        !aligned_to(local_hist, 8) || !aligned_to(final_hist, 8))
        return MSPL_ERR_ALIGN;
    bracket_select_kernel<<<K, 256, 0, static_cast<cudaStream_t>(stream)>>>(hist, portion, ignore_label, static_cast<RadixState*>(state),
                                                                            reinterpret_cast<float2*>(bracket), thresh, kept_count,
                                                                            local_hist, final_hist);
    return launch_status();
}

extern "C" int mspl_bracket_classify(const uint8_t* label, const float* conf, const float* bracket, int64_t num_pixels,
                                     int num_target_classes, int ignore_label, uint8_t* final_label, uint8_t* ignore_mask,
                                     unsigned long long* final_hist, uint32_t* cand_index, unsigned long long* cand_count,
                                     void* stream) {
    const int K = num_target_classes;
    if (!label || !conf || !bracket || !cand_index || !cand_count || num_pixels < 0) return MSPL_ERR_BAD_ARG;
    if (K < 1 || K > MSPL_MAX_CLASSES || ignore_label < -1 || ignore_label >= MSPL_MAX_CLASSES) return MSPL_ERR_BAD_ARG;
    if (ignore_label < 0 && (final_label || ignore_mask || final_hist)) return MSPL_ERR_BAD_ARG;   // thresholds-only mode
    if (num_pixels > 0xffffffffll) return MSPL_ERR_UNSUPPORTED;          // candidate indices are 32-bit
    if (!aligned_to(conf, 4) || !aligned_to(bracket, 8) || !aligned_to(cand_index, 4) || !aligned_to(cand_count, 8) ||
        (final_hist && !aligned_to(final_hist, 8)))
        return MSPL_ERR_ALIGN;
    if (num_pixels == 0) return MSPL_OK;
    auto ok = [&](size_t a) {
        return aligned_to(label, a) && aligned_to(conf, 4 * a) && (!final_label || aligned_to(final_label, a)) &&
               (!ignore_mask || aligned_to(ignore_mask, a));
    };
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float2* br = reinterpret_cast<const float2*>(bracket);
    if (num_pixels % 4 == 0 && ok(4)) {
        constexpr int kU = MSPL_CLASSIFY_UNR;
        const uint32_t n_groups = (uint32_t)(num_pixels / 4);
        auto launch = [&](auto kern) {
            const int64_t grid = resident_grid(kern, n_groups / kU + 1, 256, 0);
            kern<<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(label), reinterpret_cast<const float4*>(conf), br,
                                                 n_groups, K, ignore_label, reinterpret_cast<uint32_t*>(final_label),
                                                 reinterpret_cast<uint32_t*>(ignore_mask), final_hist, cand_index, cand_count);
        };
        if (final_hist) launch(bracket_classify_words_kernel<kU, true>);
        else launch(bracket_classify_words_kernel<kU, false>);
    } else {
        const int64_t grid = resident_grid(bracket_classify_kernel<1>, num_pixels / kUnroll + 1, 256, 0);
        bracket_classify_kernel<1><<<(unsigned)grid, 256, 0, st>>>(label, conf, br, num_pixels, K, ignore_label, final_label, ignore_mask,
                                                                  final_hist, cand_index, cand_count);
    }
    return launch_status();
}

extern "C" int mspl_cand_hist_pass(const uint8_t* label, const float* conf, const uint32_t* cand_index,
                                   const unsigned long long* cand_count, int64_t pixels_per_image, int num_target_classes, int pass,
                                   const void* state, unsigned long long* hist, int ds_rate, void* stream) {
    const int K = num_target_classes;
    if (!label || !conf || !cand_index || !cand_count || !state || !hist || pixels_per_image < 1) return MSPL_ERR_BAD_ARG;
    if (K < 1 || K > MSPL_MAX_CLASSES || pass < 0 || pass >= MSPL_RADIX_PASSES || ds_rate < 1) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(conf, 4) || !aligned_to(cand_index, 4) || !aligned_to(cand_count, 8) || !aligned_to(hist, 8) || !aligned_to(state, 8))
        return MSPL_ERR_ALIGN;
    const size_t smem = sizeof(uint32_t) * (size_t)K * MSPL_RADIX_BINS;
    auto kern = pass == 0 ? cand_hist_kernel<0> : pass == 1 ? cand_hist_kernel<1> : cand_hist_kernel<2>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return MSPL_ERR_CUDA;
    }
    kern<<<2 * kNumSMs, kHistThreads, smem, static_cast<cudaStream_t>(stream)>>>(label, conf, cand_index, cand_count, pixels_per_image,
                                                                               K, static_cast<const RadixState*>(state), hist, ds_rate);
    return launch_status();
}

extern "C" int mspl_cand_select(unsigned long long* hist, int num_target_classes, int pass, void* state, float* thresh,
                                unsigned long long* final_hist, int ignore_label, void* stream) {
    const int K = num_target_classes;
    if (!hist || !state || !thresh || K < 1 || K > MSPL_MAX_CLASSES || pass < 0 || pass >= MSPL_RADIX_PASSES) return MSPL_ERR_BAD_ARG;
    if (final_hist && (ignore_label < 0 || ignore_label >= K)) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(hist, 8) || !aligned_to(state, 8) || !aligned_to(thresh, 4) || !aligned_to(final_hist, 8)) return MSPL_ERR_ALIGN;
    radix_select_kernel<false><<<K, 256, 0, static_cast<cudaStream_t>(stream)>>>(hist, pass, 0.0, static_cast<RadixState*>(state),
                                                                                 thresh, nullptr, final_hist, ignore_label);
    return launch_status();
}

extern "C" int mspl_cand_apply(const uint8_t* label, const float* conf, const float* thresh, const uint32_t* cand_index,
                               const unsigned long long* cand_count, int num_target_classes, int ignore_label, uint8_t* final_label,
                               uint8_t* ignore_mask, unsigned long long* final_hist, void* stream) {
    const int K = num_target_classes;
    if (!label || !conf || !thresh || !cand_index || !cand_count) return MSPL_ERR_BAD_ARG;
    if (K < 1 || K > MSPL_MAX_CLASSES || ignore_label < 0 || ignore_label >= MSPL_MAX_CLASSES) return MSPL_ERR_BAD_ARG;
    if (!final_label && !ignore_mask && !final_hist) return MSPL_OK;
    if (!aligned_to(conf, 4) || !aligned_to(thresh, 4) || !aligned_to(cand_index, 4) || !aligned_to(cand_count, 8) ||
        (final_hist && !aligned_to(final_hist, 8)))
        return MSPL_ERR_ALIGN;
    cand_apply_kernel<<<2 * kNumSMs, 256, 0, static_cast<cudaStream_t>(stream)>>>(label, conf, thresh, cand_index, cand_count, K,
                                                                                 ignore_label, final_label, ignore_mask, final_hist);
    return launch_status();
}

extern "C" int mspl_cand_resolve(const uint8_t* label, const float* conf, const uint32_t* cand_index, const unsigned long long* cand_count,
                                 int64_t pixels_per_image, int num_target_classes, int ignore_label, int ds_rate, void* state,
                                 float* thresh, uint8_t* final_label, uint8_t* ignore_mask, unsigned long long* final_hist,
                                 void* stream) {
    const int K = num_target_classes;
    if (!label || !conf || !cand_index || !cand_count || !state || !thresh || pixels_per_image < 1) return MSPL_ERR_BAD_ARG;
    if (K < 1 || K > MSPL_MAX_CLASSES || ds_rate < 1 || ignore_label < -1 || ignore_label >= MSPL_MAX_CLASSES) return MSPL_ERR_BAD_ARG;
    if ((final_label || ignore_mask || final_hist) && (ignore_label < 0 || ignore_label >= K)) return MSPL_ERR_BAD_ARG;
    if (!aligned_to(conf, 4) || !aligned_to(cand_index, 4) || !aligned_to(cand_count, 8) || !aligned_to(state, 8) ||
        !aligned_to(thresh, 4) || !aligned_to(final_hist, 8))
        return MSPL_ERR_ALIGN;
    const size_t smem = sizeof(uint32_t) * (2 * (size_t)MSPL_RADIX_BINS + 2 * (size_t)kResolveCap);
    if (cudaFuncSetAttribute(cand_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return MSPL_ERR_CUDA;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(K * kResolveCtas));
    cfg.blockDim = dim3(kResolveThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kResolveCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;       // e.g. a partitioned GPU without 8 free SMs in one GPC: the caller runs the multi-launch passes instead
    if (cudaOccupancyMaxActiveClusters(&clusters, cand_resolve_kernel, &cfg) != cudaSuccess || clusters < 1) {
        cudaGetLastError();
        return MSPL_ERR_UNSUPPORTED;
    }
    if (cudaLaunchKernelEx(&cfg, cand_resolve_kernel, label, conf, cand_index, cand_count, pixels_per_image, ignore_label, ds_rate,
                           static_cast<RadixState*>(state), thresh, final_label, ignore_mask, final_hist) != cudaSuccess) {
        cudaGetLastError();
        return MSPL_ERR_CUDA;
    }
    return launch_status();
}
