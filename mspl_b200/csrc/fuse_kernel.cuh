// K1 -- fused multi-source pseudo-label generation.
//
// One pass over HBM: every (main, aux) logit of every source is read exactly once; per pixel the kernel produces the
// voted / fused label (u8), its confidence, the mean main-vs-aux KLD, and accumulates the class histogram plus radix
// pass 0 of the per-class confidence histogram in shared memory.  Replaces uest_seg_multi_os.py:897-921 (+ :669-718)
// -- see include/mspl_b200.h.
//
// Two load mechanisms share all the per-pixel math below:
//   fuse_sources_direct_kernel : every thread streams its own pixels with 128-bit ld.global.nc (any shape/alignment
//                                when P=1; needs hw % 4 == 0 and 16-byte bases when P=4)
//   fuse_sources_tma_kernel    : a producer warp stages [classes-chunk x tile-of-pixels] boxes into a shared-memory
//                                ring with bulk async copies (TMA engine, cp.async.bulk + mbarrier complete_tx),
//                                consumer warps compute from shared memory; bytes in flight no longer cost registers
#pragma once
#include "pixel_math.cuh"
#include "bilinear.cuh"

namespace mspl {

// Geometry of the fused-upsample variant (K1-lowres): the sources hand over their logits BEFORE the network's final
// F.interpolate(..., mode='bilinear', align_corners=True) (model/segmentation/espdnet_ue.py:301-302); the kernel
// interpolates on the fly.  All zero for the plain kernels.
struct LowresGeom {
    int H, W;                            // output (full) resolution; hw == H * W
    int hm[MSPL_MAX_SOURCES], wm[MSPL_MAX_SOURCES];   // main head resolution per source
    int ha[MSPL_MAX_SOURCES], wa[MSPL_MAX_SOURCES];   // aux head resolution per source
    int main_cls_stride, aux_cls_stride, aux_base;    // float offsets inside one ring stage
    int stage_floats;
};

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
    LowresGeom lr;
};

// Shared-memory bookkeeping common to both kernels: [K*2048 u32 conf histogram][8 u32 class counts][S*256 B tables]
inline size_t fuse_tally_smem_bytes(int K) {
    // [K*2048 hist][16 u32: class counts + spare][tables] (+16: padded-class table reads)
    return sizeof(uint32_t) * ((size_t)K * MSPL_RADIX_BINS + 16) + MSPL_MAX_SOURCES * MSPL_MAX_SRC_CLASSES + 16;
}

// ---- per-pixel accumulation across sources --------------------------------------------------------------------------
// GK: per-target-class probabilities needed (policy 'prob', or a vote threshold below S); otherwise every source voted
// for the winning label and G_s[label] is simply that source's max probability.
template <int P, int K, bool GK, bool TOP2>
struct PixelFusion {
    float usum[P], csum[P], Fk[GK ? K : 1][P];
    uint32_t votes[P];
    int last_lab[P];        // label proposed by the most recent source (unanimity shortcut in finish())
    bool marg[P];

    MSPL_DEVINL void reset() {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            usum[p] = csum[p] = 0.f;
            votes[p] = 0;
            marg[p] = false;
#pragma unroll
            for (int k = 0; k < (GK ? K : 1); ++k) Fk[k][p] = 0.f;
        }
    }

    // fold one finished source in; d receives its KLD map values.  gm/ga: this thread's first pixel, class 0, of the
    // source's two heads in global memory (only touched on the degenerate slow path).
    MSPL_DEVINL void add_source(const SourceStats<P>& st, const float (&zk)[K][P], const uint8_t* s_lut_s, float (&d)[P],
                                const float* __restrict__ gm, const float* __restrict__ ga, int C, int64_t hw) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            SourceResult r = finish_source<P>(st, p);
            if (r.degenerate) {
                r.rz = st.Mz[p];
                r.inv_sz = r.pmax = recompute_pmax(gm + p, ga + p, C, hw, st.Mz[p]);
            }
            d[p] = r.kld;
            usum[p] += r.kld;
            if (TOP2) marg[p] |= (1.0f - exp_neg(st.z2[p] - st.Mz[p])) * r.pmax < kNearTieMargin;
            const int lab = s_lut_s[st.amax[p]];
            votes[p] += 1u << (4 * lab);
            last_lab[p] = lab;
            if (GK) {
#pragma unroll
                for (int k = 1; k < K; ++k)      // G[0] = 0 as transfer_output_to_greenhouse (uest_seg_multi_os.py:1340)
                    Fk[k][p] += fminf(exp_neg(zk[k][p] - r.rz) * r.inv_sz, 1.0f);
            } else {
                csum[p] += r.pmax;
            }
        }
    }

    // K1-lowres has no full-resolution logits to recompute from: a degenerate pixel (head maxima more than 16 logit units above the fused maximum)
    // falls back to the exact-but-underflow-prone shared-exponential value, clamped to a valid probability.
    MSPL_DEVINL void add_source_lowres(const SourceStats<P>& st, const float (&zk)[K][P], const uint8_t* s_lut_s, float (&d)[P]) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            SourceResult r = finish_source<P>(st, p);
            if (r.degenerate) {
                r.pmax = fminf(fmaxf(r.pmax, 0.f), 1.f);
                if (!(r.pmax == r.pmax)) r.pmax = 1.f;
            }
            d[p] = r.kld;
            usum[p] += r.kld;
            if (TOP2) marg[p] |= (1.0f - exp_neg(st.z2[p] - st.Mz[p])) * r.pmax < kNearTieMargin;
            const int lab = s_lut_s[st.amax[p]];
            votes[p] += 1u << (4 * lab);
            last_lab[p] = lab;
            if (GK) {
#pragma unroll
                for (int k = 1; k < K; ++k) Fk[k][p] += fminf(exp_neg(zk[k][p] - r.rz) * r.inv_sz, 1.0f);
            } else {
                csum[p] += r.pmax;
            }
        }
    }

    // inv_s = 1/S: the averages over sources are formed as sum * (1/S) (one rounding away from the sum / S the
    // definition states; far inside the 1e-5 tolerance) to keep IEEE division sequences out of the hot loop
    MSPL_DEVINL void finish(const FuseParams& prm, float inv_s, int (&label)[P], float (&conf)[P], float (&unc)[P]) {
        const int ignore = prm.ignore;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            unc[p] = usum[p] * inv_s;
            if (prm.policy == MSPL_POLICY_PROB) {
                float best = -1.f, second = -1.f;
                int bk = 0;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float f = GK ? Fk[GK ? k : 0][p] * inv_s : 0.f;
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
                if (!GK && prm.vote_t == prm.S) {
                    // unanimity required ('all'): the label survives only if every source proposed it, so looking at the
                    // last proposal's count is enough -- no scan over the classes
                    bk = last_lab[p];
                    bc = (votes[p] >> (4 * bk)) & 15u;
                } else {
#pragma unroll
                    for (int k = 1; k < K; ++k) {        // merge_outputs: most votes, lowest class on ties (:713)
                        const uint32_t c = (votes[p] >> (4 * k)) & 15u;
                        if (c > bc) { bc = c; bk = k; }
                    }
                }
                label[p] = ((int)bc < prm.vote_t) ? ignore : bk;     // (:716)
                // every source voted for `label`, so G_s[label] is that source's max probability -- except for target class 0,
                // which transfer_output_to_greenhouse never fills (G[0] = 0, uest_seg_multi_os.py:1340)
                float f = (label[p] == 0) ? 0.f : csum[p];
                if (GK) {
                    f = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) f = (label[p] == k) ? Fk[GK ? k : 0][p] : f;
                }
                conf[p] = (label[p] == ignore) ? 0.f : f * inv_s;
            }
        }
    }
};

// ---- per-thread tallies: class counts (packed 8 bits per class, spilled before overflow), near-ties, histogram ----------
template <int K>
struct Tally {
    unsigned long long packed = 0;
    uint32_t full[K] = {};
    uint32_t n_marginal = 0, n_ignore_zero = 0;
    int pending = 0;

    MSPL_DEVINL void spill() {
#pragma unroll
        for (int k = 0; k < K; ++k) full[k] += (uint32_t)(packed >> (8 * k)) & 0xffu;
        packed = 0;
        pending = 0;
    }

    template <int P>
    MSPL_DEVINL void add(const FuseParams& prm, uint32_t* s_hist, const int (&label)[P], const float (&conf)[P],
                         const bool (&marg)[P], int64_t off, bool active) {
        const bool want_hist = prm.conf_hist != nullptr;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            packed += (unsigned long long)active << (8 * label[p]);
            n_marginal += (marg[p] && active);
            if (want_hist) {
                const bool keep = active && (prm.ds_rate <= 1 || ((off + p) % prm.ds_rate) == 0);
                // vote policies give ignore-labelled pixels conf == 0: one known bin, counted without atomics
                if (prm.policy != MSPL_POLICY_PROB && label[p] == prm.ignore) n_ignore_zero += keep;
                else if (keep) atomicAdd(&s_hist[label[p] * MSPL_RADIX_BINS + conf_bin(conf[p])], 1u);
            }
        }
        if ((pending += P) > 255 - P) spill();
    }

    // whole CTA must call this (it syncs); nthreads = blockDim.x
    MSPL_DEVINL void flush(const FuseParams& prm, uint32_t* s_hist, uint32_t* s_cls, int nthreads) {
        spill();
        const bool leader = (threadIdx.x & 31) == 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t w = __reduce_add_sync(0xffffffffu, full[k]);
            if (leader && w) atomicAdd(&s_cls[k], w);
        }
        const uint32_t wz = __reduce_add_sync(0xffffffffu, n_ignore_zero);
        if (leader && wz) atomicAdd(&s_hist[prm.ignore * MSPL_RADIX_BINS + conf_bin(0.f)], wz);
        const uint32_t wm = __reduce_add_sync(0xffffffffu, n_marginal);
        if (leader && wm && prm.marginal) atomicAdd(prm.marginal, (unsigned long long)wm);
        __syncthreads();
        if ((int)threadIdx.x < prm.K && s_cls[threadIdx.x]) atomicAdd(prm.class_hist + threadIdx.x, (unsigned long long)s_cls[threadIdx.x]);
        if (prm.conf_hist)
            for (int i = threadIdx.x; i < prm.K * MSPL_RADIX_BINS; i += nthreads)
                if (s_hist[i]) atomicAdd(prm.conf_hist + i, (unsigned long long)s_hist[i]);
    }
};

MSPL_DEVINL void tally_smem_init(const FuseParams& prm, unsigned char* smem, uint32_t*& s_hist, uint32_t*& s_cls, uint8_t*& s_lut,
                                 int nthreads) {
    const int nbins = prm.K * MSPL_RADIX_BINS;
    s_hist = reinterpret_cast<uint32_t*>(smem);
    s_cls = s_hist + nbins;
    s_lut = reinterpret_cast<uint8_t*>(s_cls + 16);
    for (int i = threadIdx.x; i < nbins; i += nthreads) s_hist[i] = 0;
    if (threadIdx.x < 16) s_cls[threadIdx.x] = 0;
    for (int i = threadIdx.x; i < prm.S * MSPL_MAX_SRC_CLASSES; i += nthreads)
        s_lut[i] = prm.lut[i / MSPL_MAX_SRC_CLASSES][i % MSPL_MAX_SRC_CLASSES];
}

template <int K, int P>
MSPL_DEVINL void reset_zk(float (&zk)[K][P]) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int p = 0; p < P; ++p) zk[k][p] = -INFINITY;
}

// ======================================================================================================================
// Direct-load kernel.  P: pixels per thread (vector width), CH: classes per chunk, KT: compile-time bound on the target
// classes (prm.K <= KT; classes in [prm.K, KT) never receive votes or probability).
// ======================================================================================================================
template <int P, int CH, int KT, bool GK, bool TOP2, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fuse_sources_direct_kernel(const __grid_constant__ FuseParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *s_hist, *s_cls;
    uint8_t* s_lut;
    tally_smem_init(prm, smem_raw, s_hist, s_cls, s_lut, THREADS);
    __syncthreads();

    const int S = prm.S;
    const float fS = 1.0f / (float)S;
    const int64_t hw = prm.hw;
    const int64_t gpi = hw / P;                       // pixel groups per image
    const int64_t n_groups = prm.n_img * gpi;
    const int64_t n_tiles = (n_groups + THREADS - 1) / THREADS;
    Tally<KT> tally;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t g = tile * THREADS + threadIdx.x;
        const bool active = g < n_groups;
        g = active ? g : n_groups - 1;
        const int64_t n = g / gpi;
        const int64_t off = (g - n * gpi) * P;       // first pixel of the group inside its image

        PixelFusion<P, KT, GK, TOP2> fus;
        fus.reset();
        for (int s = 0; s < S; ++s) {
            const int C = prm.C[s];
            const float* pm = prm.main[s] + (n * C) * hw + off;
            const float* pa = prm.aux[s] + (n * C) * hw + off;
            SourceStats<P> st;
            st.reset();
            float zk[KT][P];
            reset_zk<KT, P>(zk);
            for (int c0 = 0; c0 < C; c0 += CH) {
                const int cn = min(CH, C - c0);
                float m[CH][P], a[CH][P];
                load_chunk<P, CH>(pm + c0 * hw, pa + c0 * hw, hw, cn, m, a);
                fold_chunk<P, CH, TOP2, GK, KT>(st, m, a, c0, c0 == 0, s_lut + s * MSPL_MAX_SRC_CLASSES, zk);
            }
            float d[P];
            fus.add_source(st, zk, s_lut + s * MSPL_MAX_SRC_CLASSES, d, pm, pa, C, hw);
            if (prm.kld[s] != nullptr && active) PixVec<P>::store(prm.kld[s] + n * hw + off, d);
        }
        int label[P];
        float conf[P], unc[P];
        fus.finish(prm, fS, label, conf, unc);
        if (active) {
            const int64_t o = n * hw + off;
            store_labels<P>(prm.label + o, label);
            if (prm.conf) PixVec<P>::store(prm.conf + o, conf);
            if (prm.unc) PixVec<P>::store(prm.unc + o, unc);
        }
        tally.template add<P>(prm, s_hist, label, conf, fus.marg, off, active);
    }
    tally.flush(prm, s_hist, s_cls, THREADS);
}

// ======================================================================================================================
// TMA (bulk async copy) staged kernel.
//   CTA = NCW consumer warps + 1 producer warp.  A tile is up to TP = NCW*32*P consecutive pixels of one image (the
//   last tile of an image may be shorter; hw % 4 == 0 keeps every row copy a multiple of 16 bytes).
//   A stage holds one chunk of up to CH classes of both heads for the tile: [2][CH][TP] floats; the producer fills stages
//   in the order (tile, source, chunk) with one cp.async.bulk per class row, completion signalled on the stage's "full"
//   mbarrier (complete_tx::bytes); a consumer warp copies its pixels' values to registers, releases the stage ("empty"
//   mbarrier, one arrival per consumer warp) and only then does the math, so the ring turns over at load speed.
// ======================================================================================================================
namespace tma {

MSPL_DEVINL uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

MSPL_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
MSPL_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
MSPL_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
MSPL_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`; streaming data: L2 evict-first hint
MSPL_DEVINL void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_addr(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}
MSPL_DEVINL uint64_t evict_first_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

template <int P> MSPL_DEVINL void lds(const float* p, float (&v)[P]);
template <> MSPL_DEVINL void lds<1>(const float* p, float (&v)[1]) { v[0] = *p; }
template <> MSPL_DEVINL void lds<2>(const float* p, float (&v)[2]) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
}
template <> MSPL_DEVINL void lds<4>(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

}  // namespace tma

// 1: the first chunk of a source skips the rescale of the (empty) accumulators, which makes the compiler peel that
// iteration (two copies of the chunk body, ~38 KB of code); 0: one copy, the rescale runs on empty accumulators too.
#ifndef MSPL_PEEL_FIRST
#define MSPL_PEEL_FIRST 1
#endif

template <int NCW, int P, int CH, int NSTAGE>
struct TmaCfg {
    static constexpr int kThreads = (NCW + 1) * 32;
    static constexpr int kTilePix = NCW * 32 * P;
    static constexpr int kStageFloats = 2 * CH * kTilePix;
    static constexpr size_t kRingBytes = sizeof(float) * (size_t)kStageFloats * NSTAGE;
    static size_t smem_bytes(int K) { return kRingBytes + 2 * NSTAGE * sizeof(uint64_t) + fuse_tally_smem_bytes(K) + 128; }
};

// (image, tile inside the image) of the tiles blockIdx.x, blockIdx.x + gridDim.x, ... without a 64-bit division per tile.
struct TileWalker {
    int64_t image, tile_in_image;
    int64_t step_images, step_tiles, tpi;
    MSPL_DEVINL TileWalker(int64_t first, int64_t stride, int64_t tiles_per_image)
        : image(first / tiles_per_image), tile_in_image(first % tiles_per_image), step_images(stride / tiles_per_image),
          step_tiles(stride % tiles_per_image), tpi(tiles_per_image) {}
    MSPL_DEVINL void next() {
        image += step_images;
        tile_in_image += step_tiles;
        if (tile_in_image >= tpi) { tile_in_image -= tpi; ++image; }
    }
};

// Producer warp of the TMA-staged kernels: walks this CTA's tiles in (tile, source, chunk) order and fills the ring.
template <int NCW, int P, int CH, int NSTAGE>
MSPL_DEVINL void tma_produce_tiles(const FuseParams& prm, float* ring, uint64_t* full, uint64_t* empty, int lane) {
    using Cfg = TmaCfg<NCW, P, CH, NSTAGE>;
    constexpr int TP = Cfg::kTilePix;
    const int S = prm.S;
    const int64_t hw = prm.hw;
    const int64_t tpi = (hw + TP - 1) / TP;
    const int64_t n_tiles = prm.n_img * tpi;
    const uint64_t policy = tma::evict_first_policy();
    int stage = 0;
    uint32_t phase = 0;
    TileWalker walk(blockIdx.x, gridDim.x, tpi);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, walk.next()) {
        const int64_t n = walk.image;
        const int64_t off = walk.tile_in_image * TP;
        const uint32_t row_bytes = (uint32_t)((hw - off < TP ? hw - off : TP) * sizeof(float));
        for (int s = 0; s < S; ++s) {
            const int C = prm.C[s];
            const float* pm = prm.main[s] + (n * C) * hw + off;
            const float* pa = prm.aux[s] + (n * C) * hw + off;
            for (int c0 = 0; c0 < C; c0 += CH) {
                const int cn = min(CH, C - c0);
                tma::mbar_wait(&empty[stage], phase ^ 1);          // all consumers released this slot
                float* dst = ring + (size_t)stage * Cfg::kStageFloats;
                if (cn < CH) {
                    // tail chunk: the class rows this source does not have are filled with kPadLogit HERE, so that the
                    // consumers' loads and math carry no predicates and exist only once in the instruction stream
                    const float4 pad = make_float4(kPadLogit, kPadLogit, kPadLogit, kPadLogit);
                    for (int j = cn; j < CH; ++j)
                        for (int i = lane; i < TP / 4; i += 32) {
                            reinterpret_cast<float4*>(dst + j * TP)[i] = pad;
                            reinterpret_cast<float4*>(dst + (CH + j) * TP)[i] = pad;
                        }
                    __threadfence_block();
                }
                __syncwarp();
                if (lane == 0) tma::mbar_arrive_expect_tx(&full[stage], 2 * cn * row_bytes);   // release: publishes the padding too
                for (int j = lane; j < 2 * cn; j += 32) {
                    const int head = j >= cn, c = head ? j - cn : j;
                    const float* src = (head ? pa : pm) + (int64_t)(c0 + c) * hw;
                    tma::bulk_g2s(dst + (head * CH + c) * TP, src, row_bytes, &full[stage], policy);
                }
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    }
}

template <int NCW, int P, int CH, int NSTAGE, int KT, bool GK, bool TOP2>
__global__ void __launch_bounds__((NCW + 1) * 32, 1) fuse_sources_tma_kernel(const __grid_constant__ FuseParams prm) {
    using Cfg = TmaCfg<NCW, P, CH, NSTAGE>;
    constexpr int TP = Cfg::kTilePix;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + Cfg::kRingBytes);
    uint64_t* empty = full + NSTAGE;
    uint32_t *s_hist, *s_cls;
    uint8_t* s_lut;
    tally_smem_init(prm, smem_raw + Cfg::kRingBytes + 2 * NSTAGE * sizeof(uint64_t), s_hist, s_cls, s_lut, Cfg::kThreads);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int S = prm.S;
    const int64_t hw = prm.hw;
    const int64_t tpi = (hw + TP - 1) / TP;            // tiles per image (the last one may be partial)
    const int64_t n_tiles = prm.n_img * tpi;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Tally<KT> tally;

    if (warp == NCW) {
        tma_produce_tiles<NCW, P, CH, NSTAGE>(prm, ring, full, empty, lane);
    } else {
        // ------------------------------- consumer warps -------------------------------
        const float fS = 1.0f / (float)S;
        int stage = 0;
        uint32_t phase = 0;
        const int px = (warp * 32 + lane) * P;          // this thread's first pixel inside the tile
        TileWalker walk(blockIdx.x, gridDim.x, tpi);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, walk.next()) {
            const int64_t n = walk.image;
            const int64_t off = walk.tile_in_image * TP + px;
            const bool active = off < hw;                 // partial last tile: the ring holds stale values past the image
            PixelFusion<P, KT, GK, TOP2> fus;
            fus.reset();
            for (int s = 0; s < S; ++s) {
                const int C = prm.C[s];
                SourceStats<P> st;
                st.reset();
                float zk[KT][P];
                reset_zk<KT, P>(zk);
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += CH) {
                    float m[CH][P], a[CH][P];
                    tma::mbar_wait(&full[stage], phase);
                    const float* src = ring + (size_t)stage * Cfg::kStageFloats + px;
#pragma unroll
                    for (int j = 0; j < CH; ++j) tma::lds<P>(src + j * TP, m[j]);
#pragma unroll
                    for (int j = 0; j < CH; ++j) tma::lds<P>(src + (CH + j) * TP, a[j]);
                    __syncwarp();
                    if (lane == 0) tma::mbar_arrive(&empty[stage]);    // values are in registers: hand the slot back
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    fold_chunk<P, CH, TOP2, GK, KT>(st, m, a, c0, MSPL_PEEL_FIRST ? c0 == 0 : false, s_lut + s * MSPL_MAX_SRC_CLASSES, zk);
                }
                float d[P];
                const int64_t goff = (n * C) * hw + (active ? off : 0);
                fus.add_source(st, zk, s_lut + s * MSPL_MAX_SRC_CLASSES, d, prm.main[s] + goff, prm.aux[s] + goff, C, hw);
                if (prm.kld[s] != nullptr && active) PixVec<P>::store(prm.kld[s] + n * hw + off, d);
            }
            int label[P];
            float conf[P], unc[P];
            fus.finish(prm, fS, label, conf, unc);
            if (active) {
                const int64_t o = n * hw + off;
                store_labels<P>(prm.label + o, label);
                if (prm.conf) PixVec<P>::store(prm.conf + o, conf);
                if (prm.unc) PixVec<P>::store(prm.unc + o, unc);
            }
            tally.template add<P>(prm, s_hist, label, conf, fus.marg, off, active);
        }
    }
    tally.flush(prm, s_hist, s_cls, Cfg::kThreads);
}

// ----------------------------------------------------------------------------------------------------------------------
// Labels-only variant: what the reference's generation loop actually keeps (uest_seg_multi_os.py:900-921 discards the KLD and
// never forms a confidence).  With no confidence, uncertainty or histogram of confidences requested there is nothing to
// exponentiate: per class and pixel the consumers do z = m + a/2 and a running first-argmax (4 instructions), so the kernel
// is HBM-bound with a wide margin even at reduced clocks.  Same ring, same producer.
// ----------------------------------------------------------------------------------------------------------------------
template <int NCW, int P, int CH, int NSTAGE, int KT>
__global__ void __launch_bounds__((NCW + 1) * 32, 1) fuse_labels_tma_kernel(const __grid_constant__ FuseParams prm) {
    using Cfg = TmaCfg<NCW, P, CH, NSTAGE>;
    constexpr int TP = Cfg::kTilePix;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + Cfg::kRingBytes);
    uint64_t* empty = full + NSTAGE;
    uint32_t *s_hist, *s_cls;
    uint8_t* s_lut;
    tally_smem_init(prm, smem_raw + Cfg::kRingBytes + 2 * NSTAGE * sizeof(uint64_t), s_hist, s_cls, s_lut, Cfg::kThreads);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int S = prm.S;
    const int64_t hw = prm.hw;
    const int64_t tpi = (hw + TP - 1) / TP;
    const int64_t n_tiles = prm.n_img * tpi;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Tally<KT> tally;
    if (warp == NCW) {
        tma_produce_tiles<NCW, P, CH, NSTAGE>(prm, ring, full, empty, lane);
    } else {
        int stage = 0;
        uint32_t phase = 0;
        const int px = (warp * 32 + lane) * P;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t n = tile / tpi;
            const int64_t off = (tile - n * tpi) * TP + px;
            const bool active = off < hw;
            uint32_t votes[P];
#pragma unroll
            for (int p = 0; p < P; ++p) votes[p] = 0;
            for (int s = 0; s < S; ++s) {
                const int C = prm.C[s];
                float best[P];
                int amax[P];
#pragma unroll
                for (int p = 0; p < P; ++p) { best[p] = -INFINITY; amax[p] = 0; }
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += CH) {
                    float m[CH][P], a[CH][P];
                    tma::mbar_wait(&full[stage], phase);
                    const float* src = ring + (size_t)stage * Cfg::kStageFloats + px;
#pragma unroll
                    for (int j = 0; j < CH; ++j) tma::lds<P>(src + j * TP, m[j]);
#pragma unroll
                    for (int j = 0; j < CH; ++j) tma::lds<P>(src + (CH + j) * TP, a[j]);
                    __syncwarp();
                    if (lane == 0) tma::mbar_arrive(&empty[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
#pragma unroll
                    for (int j = 0; j < CH; ++j)
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const float z = fmaf(0.5f, a[j][p], m[j][p]);     // as `pred + 0.5 * pred_aux` (:687); padded classes: -1.5e30
                            amax[p] = (z > best[p]) ? (c0 + j) : amax[p];    // strict >: first maximal index, as np.argmax (:904)
                            best[p] = fmaxf(best[p], z);
                        }
                }
#pragma unroll
                for (int p = 0; p < P; ++p) votes[p] += 1u << (4 * s_lut[s * MSPL_MAX_SRC_CLASSES + amax[p]]);
            }
            int label[P];
            float conf[P];
            bool marg[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                int bk = 0;
                uint32_t bc = votes[p] & 15u;
#pragma unroll
                for (int k = 1; k < KT; ++k) {        // merge_outputs: most votes, lowest class on ties (:713)
                    const uint32_t c = (votes[p] >> (4 * k)) & 15u;
                    if (c > bc) { bc = c; bk = k; }
                }
                label[p] = ((int)bc < prm.vote_t) ? prm.ignore : bk;     // (:716)
                conf[p] = 0.f;
                marg[p] = false;
            }
            if (active) store_labels<P>(prm.label + n * hw + off, label);
            tally.template add<P>(prm, s_hist, label, conf, marg, off, active);
        }
    }
    tally.flush(prm, s_hist, s_cls, Cfg::kThreads);
}

// ======================================================================================================================
// K1-lowres: the same fusion, reading the sources' logits at their native (pre-upsample) resolution and performing the
// network's final bilinear align_corners=True interpolation in the consumer warps (next-row component).
// HBM traffic drops from 8*sum(C) B/pixel to 4*sum(C)*(hm*wm + ha*wa)/(H*W) (1.25*sum(C) for the x2 / x4 heads of
// ESPDNetUE); the kernel becomes instruction-bound.
//   A tile is TP consecutive output pixels (row-major) of one image.  For every class of the chunk the producer copies
//   the block of source rows the tile's output rows interpolate from (whole rows, one bulk copy per class and head).
//   Arithmetic follows ATen's upsample_bilinear2d: src = dst * (in-1)/(out-1); i = (int)src; lambda = src - i;
//   val = h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11).
// ======================================================================================================================
// MS / AS: compile-time class strides (floats) of the main / aux blocks inside a stage, or 0 to take them from the
// geometry at run time.  With fixed strides every tap of every class is `LDS [tap_register + immediate]`: the per-class
// address arithmetic disappears from the interpolation loop.
template <int NCW, int P, int CH, int NSTAGE, int KT, bool GK, bool TOP2, int MS = 0, int AS = 0>
__global__ void __launch_bounds__((NCW + 1) * 32, 1) fuse_sources_lowres_kernel(const __grid_constant__ FuseParams prm) {
    constexpr int kThreads = (NCW + 1) * 32;
    constexpr int TP = NCW * 32 * P;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const LowresGeom& lr = prm.lr;
    const int main_stride = MS ? MS : lr.main_cls_stride, aux_stride = AS ? AS : lr.aux_cls_stride;
    const int aux_base = MS ? CH * MS : lr.aux_base;
    const int stage_floats = (MS && AS) ? CH * (MS + AS) : lr.stage_floats;
    float* ring = reinterpret_cast<float*>(smem_raw);
    const size_t ring_bytes = sizeof(float) * (size_t)stage_floats * NSTAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + ring_bytes);
    uint64_t* empty = full + NSTAGE;
    uint32_t *s_hist, *s_cls;
    uint8_t* s_lut;
    tally_smem_init(prm, smem_raw + ring_bytes + 2 * NSTAGE * sizeof(uint64_t), s_hist, s_cls, s_lut, kThreads);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            tma::mbar_init(&full[i], 1);
            tma::mbar_init(&empty[i], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int S = prm.S, H = lr.H, W = lr.W;
    const int64_t hw = prm.hw;
    const int64_t tpi = (hw + TP - 1) / TP;
    const int64_t n_tiles = prm.n_img * tpi;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Tally<KT> tally;

    if (warp == NCW) {
        // ------------------------------- producer warp -------------------------------
        const uint64_t policy = tma::evict_first_policy();
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t n = tile / tpi;
            const int64_t off = (tile - n * tpi) * TP;
            const int y_first = (int)(off / W);
            const int64_t last = (off + TP < hw ? off + TP : hw) - 1;
            const int y_last = (int)(last / W);
            for (int s = 0; s < S; ++s) {
                const int C = prm.C[s], hm = lr.hm[s], wm = lr.wm[s], ha = lr.ha[s], wa = lr.wa[s];
                const float rhm = lowres_scale(hm, H), rha = lowres_scale(ha, H);
                const int m0 = (int)(rhm * (float)y_first), m1 = min((int)(rhm * (float)y_last) + 1, hm - 1);
                const int a0 = (int)(rha * (float)y_first), a1 = min((int)(rha * (float)y_last) + 1, ha - 1);
                const uint32_t mbytes = (uint32_t)((m1 - m0 + 1) * wm * sizeof(float));
                const uint32_t abytes = (uint32_t)((a1 - a0 + 1) * wa * sizeof(float));
                const float* pm = prm.main[s] + ((n * C) * hm + m0) * wm;
                const float* pa = prm.aux[s] + ((n * C) * ha + a0) * wa;
                for (int c0 = 0; c0 < C; c0 += CH) {
                    const int cn = min(CH, C - c0);
                    tma::mbar_wait(&empty[stage], phase ^ 1);
                    float* dst = ring + (size_t)stage * stage_floats;
                    if (cn < CH) {      // pad the class blocks a tail chunk lacks (see fuse_sources_tma_kernel)
                        const float4 pad = make_float4(kPadLogit, kPadLogit, kPadLogit, kPadLogit);
                        for (int j = cn; j < CH; ++j) {
                            for (int i = lane; i < (int)(mbytes / 16); i += 32) reinterpret_cast<float4*>(dst + j * main_stride)[i] = pad;
                            for (int i = lane; i < (int)(abytes / 16); i += 32)
                                reinterpret_cast<float4*>(dst + aux_base + j * aux_stride)[i] = pad;
                        }
                        __threadfence_block();
                    }
                    __syncwarp();
                    if (lane == 0) tma::mbar_arrive_expect_tx(&full[stage], cn * (mbytes + abytes));
                    for (int j = lane; j < 2 * cn; j += 32) {
                        const int head = j >= cn, c = head ? j - cn : j;
                        if (head) tma::bulk_g2s(dst + aux_base + c * aux_stride, pa + (int64_t)(c0 + c) * ha * wa, abytes, &full[stage], policy);
                        else tma::bulk_g2s(dst + c * main_stride, pm + (int64_t)(c0 + c) * hm * wm, mbytes, &full[stage], policy);
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ------------------------------- consumer warps -------------------------------
        const float inv_s = 1.0f / (float)S;
        int stage = 0;
        uint32_t phase = 0;
        const int px = (warp * 32 + lane) * P;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t n = tile / tpi;
            const int64_t tile_off = (tile - n * tpi) * TP;
            const int64_t off = tile_off + px;
            const bool active = off < hw;
            const int y_first = (int)(tile_off / W);
            int yy[P], xx[P];
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int64_t o = active ? off + p : tile_off;      // idle lanes interpolate the tile's first pixel
                yy[p] = (int)(o / W);
                xx[p] = (int)(o - (int64_t)yy[p] * W);
            }
            PixelFusion<P, KT, GK, TOP2> fus;
            fus.reset();
            for (int s = 0; s < S; ++s) {
                const int C = prm.C[s], hm = lr.hm[s], wm = lr.wm[s], ha = lr.ha[s], wa = lr.wa[s];
                const float rhm = lowres_scale(hm, H), rwm = lowres_scale(wm, W);
                const float rha = lowres_scale(ha, H), rwa = lowres_scale(wa, W);
                const int m0 = (int)(rhm * (float)y_first), a0 = (int)(rha * (float)y_first);
                BilinearTap tm[P], ta[P];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    tm[p] = make_tap(yy[p], xx[p], hm, wm, rhm, rwm, m0);
                    ta[p] = make_tap(yy[p], xx[p], ha, wa, rha, rwa, a0);
                }
                SourceStats<P> st;
                st.reset();
                float zk[KT][P];
                reset_zk<KT, P>(zk);
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += CH) {
                    float m[CH][P], a[CH][P];
                    tma::mbar_wait(&full[stage], phase);
                    const float* src = ring + (size_t)stage * stage_floats;
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            m[j][p] = bilinear(src + j * main_stride, tm[p]);
                            a[j][p] = bilinear(src + aux_base + j * aux_stride, ta[p]);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) tma::mbar_arrive(&empty[stage]);
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    fold_chunk<P, CH, TOP2, GK, KT>(st, m, a, c0, c0 == 0, s_lut + s * MSPL_MAX_SRC_CLASSES, zk);
                }
                float d[P];
                // (the degenerate-pixel slow path would need the full-resolution logits this variant never materialises;
                //  mspl_fuse_sources_lowres documents the |logit| <= 64 contract instead)
                fus.add_source_lowres(st, zk, s_lut + s * MSPL_MAX_SRC_CLASSES, d);
                if (prm.kld[s] != nullptr && active) PixVec<P>::store(prm.kld[s] + n * hw + off, d);
            }
            int label[P];
            float conf[P], unc[P];
            fus.finish(prm, inv_s, label, conf, unc);
            if (active) {
                const int64_t o = n * hw + off;
                store_labels<P>(prm.label + o, label);
                if (prm.conf) PixVec<P>::store(prm.conf + o, conf);
                if (prm.unc) PixVec<P>::store(prm.unc + o, unc);
            }
            tally.template add<P>(prm, s_hist, label, conf, fus.marg, off, active);
        }
    }
    tally.flush(prm, s_hist, s_cls, kThreads);
}

}  // namespace mspl
