// K1 -- fused multi-source pseudo-label generation (direct-load variant).
//
// One pass over HBM: every (main, aux) logit of every source is read exactly once with 128-bit streaming
// loads; per pixel the kernel produces the voted / fused label (u8), its confidence, the mean main-vs-aux KLD,
// and accumulates the class histogram plus radix pass 0 of the per-class confidence histogram in shared
// memory.  Replaces uest_seg_multi_os.py:897-921 (+ :669-718) -- see include/mspl_b200.h.
#pragma once
#include "pixel_math.cuh"

namespace mspl {

struct FuseParams {
    const float* main[MSPL_MAX_SOURCES];
    const float* aux[MSPL_MAX_SOURCES];
    float* kld[MSPL_MAX_SOURCES];
    int C[MSPL_MAX_SOURCES];
    uint8_t lut[MSPL_MAX_SOURCES][MSPL_MAX_SRC_CLASSES];
    int S, K, policy, vote_t, ignore, ds_rate;
    int64_t n_img, hw;
    uint8_t* label;
    float* conf;
    float* unc;
    unsigned long long* class_hist;
    unsigned long long* conf_hist;
    unsigned long long* marginal;
};

constexpr int kFuseThreads = 256;

// Dynamic shared memory: [K*2048 u32 conf histogram][8 u32 class counts][S*256 B class tables]
inline size_t fuse_smem_bytes(int K) {
    return sizeof(uint32_t) * ((size_t)K * MSPL_RADIX_BINS + 8) + MSPL_MAX_SOURCES * MSPL_MAX_SRC_CLASSES;
}

// Per-thread class counters packed 8 bits per class, spilled to full counters before they can overflow.
template <int K>
struct ClassCounter {
    unsigned long long packed = 0;
    uint32_t full[K] = {};
    int pending = 0;
    MSPL_DEVINL void add(int label, bool on) { packed += (unsigned long long)on << (8 * label); }
    MSPL_DEVINL void spill() {
#pragma unroll
        for (int k = 0; k < K; ++k) full[k] += (uint32_t)(packed >> (8 * k)) & 0xffu;
        packed = 0;
        pending = 0;
    }
    template <int P> MSPL_DEVINL void tick() { if ((pending += P) > 255 - P) spill(); }
};

// P: pixels per thread (vector width), CH: classes per chunk, KT: compile-time bound on the target classes
// (prm.K <= KT; classes in [prm.K, KT) simply never receive votes or probability),
// GK: per-target-class probabilities needed (policy 'prob', or a vote threshold below S),
// TOP2: near-tie accounting.
template <int P, int CH, int KT, bool GK, bool TOP2, int MINB>
__global__ void __launch_bounds__(kFuseThreads, MINB) fuse_sources_kernel(const __grid_constant__ FuseParams prm) {
    constexpr int K = KT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbins = prm.K * MSPL_RADIX_BINS;
    uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* s_cls = s_hist + nbins;
    uint8_t* s_lut = reinterpret_cast<uint8_t*>(s_cls + 8);

    const bool want_hist = prm.conf_hist != nullptr;
    if (want_hist)
        for (int i = threadIdx.x; i < nbins; i += kFuseThreads) s_hist[i] = 0;
    if (threadIdx.x < 8) s_cls[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < prm.S * MSPL_MAX_SRC_CLASSES; i += kFuseThreads)
        s_lut[i] = prm.lut[i / MSPL_MAX_SRC_CLASSES][i % MSPL_MAX_SRC_CLASSES];
    __syncthreads();

    const int S = prm.S;
    const float fS = (float)S;
    const int64_t hw = prm.hw;
    const int64_t gpi = hw / P;                       // pixel groups per image
    const int64_t n_groups = prm.n_img * gpi;
    const int64_t n_tiles = (n_groups + kFuseThreads - 1) / kFuseThreads;
    const bool prob_policy = prm.policy == MSPL_POLICY_PROB;
    const int ignore = prm.ignore;

    ClassCounter<K> cls;
    uint32_t n_marginal = 0, n_ignore_zero = 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t g = tile * kFuseThreads + threadIdx.x;
        const bool active = g < n_groups;
        g = active ? g : n_groups - 1;
        const int64_t n = g / gpi;
        const int64_t off = (g - n * gpi) * P;       // first pixel of the group inside its image

        float usum[P], csum[P], Fk[K][P];
        uint32_t votes[P];
        bool marg[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            usum[p] = csum[p] = 0.f;
            votes[p] = 0;
            marg[p] = false;
#pragma unroll
            for (int k = 0; k < K; ++k) Fk[k][p] = 0.f;
        }

        for (int s = 0; s < S; ++s) {
            const int C = prm.C[s];
            const float* pm = prm.main[s] + (n * C) * hw + off;
            const float* pa = prm.aux[s] + (n * C) * hw + off;
            SourceStats<P> st;
            st.reset();
            float zk[K][P];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int p = 0; p < P; ++p) zk[k][p] = -INFINITY;

            for (int c0 = 0; c0 < C; c0 += CH) {
                const int cn = min(CH, C - c0);
                float m[CH][P], a[CH][P];
                load_chunk<P, CH>(pm + c0 * hw, pa + c0 * hw, hw, cn, m, a);
                fold_chunk<P, CH, TOP2, GK, K>(st, m, a, c0, cn, c0 == 0, s_lut + s * MSPL_MAX_SRC_CLASSES, zk);
            }

            float d[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                d[p] = kld_of<P>(st, p);
                usum[p] += d[p];
                const float pmax = 1.0f / st.Sz[p];
                if (TOP2) marg[p] |= (1.0f - exp_neg(st.z2[p] - st.Mz[p])) * pmax < kNearTieMargin;
                const int lab = s_lut[s * MSPL_MAX_SRC_CLASSES + st.amax[p]];
                votes[p] += 1u << (4 * lab);
                if (GK) {
#pragma unroll
                    for (int k = 1; k < K; ++k)      // G[0] = 0 as transfer_output_to_greenhouse (:1340)
                        Fk[k][p] += exp_neg(zk[k][p] - st.Mz[p]) * pmax;
                } else {
                    csum[p] += pmax;                 // every source voted for the label: G_s[label] = max prob
                }
            }
            if (prm.kld[s] != nullptr && active) PixVec<P>::store(prm.kld[s] + n * hw + off, d);
        }

        int label[P];
        float conf[P], unc[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            unc[p] = usum[p] / fS;
            if (prob_policy) {
                float best = -1.f, second = -1.f;
                int bk = 0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float f = GK ? Fk[k][p] / fS : 0.f;
                    second = fmaxf(second, fminf(best, f));
                    bk = (f > best) ? k : bk;
                    best = fmaxf(best, f);
                }
                label[p] = bk;
                conf[p] = best;
                if (TOP2) marg[p] |= (best - second) < kNearTieMargin;
            } else {
                int bk = 0;
                uint32_t bc = votes[p] & 15u;
#pragma unroll
                for (int k = 1; k < K; ++k) {
                    const uint32_t c = (votes[p] >> (4 * k)) & 15u;
                    if (c > bc) { bc = c; bk = k; }
                }
                label[p] = ((int)bc < prm.vote_t) ? ignore : bk;
                float f = csum[p];
                if (GK) {
                    f = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) f = (label[p] == k) ? Fk[k][p] : f;
                }
                conf[p] = (label[p] == ignore) ? 0.f : f / fS;
            }
        }

        if (active) {
            const int64_t o = n * hw + off;
            store_labels<P>(prm.label + o, label);
            if (prm.conf) PixVec<P>::store(prm.conf + o, conf);
            if (prm.unc) PixVec<P>::store(prm.unc + o, unc);
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            cls.add(label[p], active);
            if (TOP2) n_marginal += (marg[p] && active);
            if (want_hist) {
                const bool keep = active && (prm.ds_rate <= 1 || ((off + p) % prm.ds_rate) == 0);
                if (!prob_policy && label[p] == ignore) n_ignore_zero += keep;   // conf == 0: one known bin
                else if (keep) atomicAdd(&s_hist[label[p] * MSPL_RADIX_BINS + (float_to_key(conf[p]) >> 21)], 1u);
            }
        }
        cls.template tick<P>();
    }

    // ---- flush per-thread counters -> shared -> global ------------------------------------------------
    cls.spill();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, cls.full[k]);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_cls[k], w);
    }
    if (want_hist) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, n_ignore_zero);
        if ((threadIdx.x & 31) == 0 && w)
            atomicAdd(&s_hist[ignore * MSPL_RADIX_BINS + (float_to_key(0.f) >> 21)], w);
    }
    if (TOP2 && prm.marginal) {
        const uint32_t w = __reduce_add_sync(0xffffffffu, n_marginal);
        if ((threadIdx.x & 31) == 0 && w) atomicAdd(prm.marginal, (unsigned long long)w);
    }
    __syncthreads();
    if (threadIdx.x < prm.K && s_cls[threadIdx.x]) atomicAdd(prm.class_hist + threadIdx.x, (unsigned long long)s_cls[threadIdx.x]);
    if (want_hist)
        for (int i = threadIdx.x; i < nbins; i += kFuseThreads)
            if (s_hist[i]) atomicAdd(prm.conf_hist + i, (unsigned long long)s_hist[i]);
}

}  // namespace mspl
