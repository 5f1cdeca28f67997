// Per-pixel statistics of one source's (main, aux) logit pair, accumulated chunk by chunk.
//
// For a pixel with class logits m_c (main head) and a_c (aux head), z_c = m_c + 0.5*a_c, the reference needs
//   softmax(z)                  uest_seg_multi_os.py:687-689
//   first-argmax_c              uest_seg_multi_os.py:904      (decided on z: softmax is monotone; ties -> lowest c)
//   KL(softmax(m)||softmax(a))  loss_fns/segmentation_loss.py:181-189
// All of them follow from the running quantities below, updated one chunk of CH classes at a time (online
// softmax: a chunk's maxima are folded in with ONE rescale per chunk, not one per class):
//   Mm, Sm = sum e^{m-Mm}     Ma, Sa = sum e^{a-Ma}     Mz, Sz = sum e^{z-Mz}
//   T  = sum e^{m-Mm} ((m-Mm) - (a-Ma))      z2 = second largest z      amax = first index of the largest z
// Then  KLD = T/Sm - log Sm + log Sa   (the reference's sum_c p1*(logp1 - logp2), with log_softmax written as
// (x - max) - log(sum) exactly as ATen does, so no large maxima are ever added back),
// max prob = 1/Sz,  top-2 margin = (1 - e^{z2-Mz})/Sz.
// When the whole source fits one chunk this IS the two-sweep (max, then sums) softmax.
#pragma once
#include "common.cuh"

namespace mspl {

// e^t for t <= 0 (t already max-subtracted, so the largest term is exactly e^0 = 1): FMUL + MUFU.EX2.
MSPL_DEVINL float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
MSPL_DEVINL float exp_neg(float t) { return ex2_approx(t * kLog2e); }
MSPL_DEVINL float exp_half_neg(float t) { return ex2_approx(t * (0.5f * kLog2e)); }   // e^{t/2}
// natural log through MUFU.LG2 (absolute error <= 2^-22.6 * |log2 x| * ln 2: ~5e-7 at worst for the ratios of softmax sums it
// is used on, inside the 2e-6 KLD floor), instead of the ~20-instruction branchy logf
MSPL_DEVINL float log_fast(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * 0.6931471805599453f;
}
// 1/x through MUFU.RCP alone (relative error <= 2^-23; the IEEE sequence adds a Newton step and a range check per call)
MSPL_DEVINL float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
MSPL_DEVINL float max3(float a, float b, float c) {        // FMNMX3 (sm_100+)
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
template <int N>
MSPL_DEVINL float max_of(const float (&v)[N], float init) {
    float r = init;
    int j = 0;
#pragma unroll
    for (; j + 1 < N; j += 2) r = max3(r, v[j], v[j + 1]);
    if (j < N) r = fmaxf(r, v[j]);
    return r;
}

// Running per-pixel statistics of one source.
//   Mm, Sm = sum e^{m-Mm}     Ma, Sa = sum e^{a-Ma}     T = sum e^{m-Mm} ((m-Mm) - (a-Ma))
//   Mz = max z, z2 = runner-up z, amax = first index of the largest z
//   Sz = sum e^{z - Rz} with the reference point Rz = Mm + Ma/2 (>= Mz): e^{z-Rz} = e^{m-Mm} * e^{(a-Ma)/2}, so the z-softmax
//        costs no exponential of its own (two MUFU.EX2 per class instead of three).
template <int P>
struct SourceStats {
    float Mm[P], Sm[P], Ma[P], Sa[P], Mz[P], Sz[P], T[P], z2[P];
    int amax[P];
    MSPL_DEVINL void reset() {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            // finite sentinel (not -inf): the rescale of an empty accumulator then evaluates to exactly 0 without NaNs, so
            // kernels may run it unconditionally on a source's first chunk
            Mm[p] = Ma[p] = Mz[p] = z2[p] = -1.0e30f;
            Sm[p] = Sa[p] = Sz[p] = T[p] = 0.f;
            amax[p] = 0;
        }
    }
};

// Value loaded in place of the classes a tail chunk does not have: hugely negative but finite, so that it never wins a
// max, its exponentials are exactly 0 and (m - Mm) - (a - Ma) stays finite (0) -- the math below needs no predicates.
constexpr float kPadLogit = -1.0e30f;

// Fold classes [c0, c0+CH) into the running stats.  m/a hold the chunk's logits; entries past the source's last class
// are kPadLogit (see load_chunk).  `first` (warp-uniform): nothing accumulated yet, skip the rescale.
// TOP2: also track the runner-up z (needed only for the near-tie report).
// GK: also track zk[k][p], the running max of z over the classes that `lut` maps to target class k.
template <int P, int CH, bool TOP2, bool GK, int K>
MSPL_DEVINL void fold_chunk(SourceStats<P>& st, const float (&m)[CH][P], const float (&a)[CH][P], int c0,
                            bool first, const uint8_t* __restrict__ lut, float (&zk)[K][P]) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        float z[CH], mv[CH], av[CH];
        float z1 = st.Mz[p], zr = st.z2[p];
        int i1 = st.amax[p];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            mv[j] = m[j][p];
            av[j] = a[j][p];
            // z exactly as the reference forms it: 0.5*a is exact in fp32, so the fused multiply-add
            // rounds once, just like `pred + 0.5 * pred_aux`.
            z[j] = fmaf(0.5f, av[j], mv[j]);
            if (TOP2) zr = fmaxf(zr, fminf(z1, z[j]));
            i1 = (z[j] > z1) ? (c0 + j) : i1;      // strict >: lowest index wins ties, as np.argmax
            z1 = fmaxf(z1, z[j]);
        }
        const float nMm = max_of<CH>(mv, st.Mm[p]), nMa = max_of<CH>(av, st.Ma[p]);
        float sm = 0.f, sa = 0.f, sz = 0.f, t = 0.f;
        if (!first) {   // rescale what earlier chunks accumulated to the new maxima
            const float dm = st.Mm[p] - nMm, da = st.Ma[p] - nMa;
            const float rm = exp_neg(dm), rh = exp_half_neg(da);
            sm = st.Sm[p] * rm;
            t = rm * fmaf(st.Sm[p], dm - da, st.T[p]);
            sa = st.Sa[p] * (rh * rh);
            sz = st.Sz[p] * (rm * rh);
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const float tm = mv[j] - nMm, ta = av[j] - nMa;
            const float em = exp_neg(tm), h = exp_half_neg(ta);
            sm += em;
            t = fmaf(em, tm - ta, t);
            sa = fmaf(h, h, sa);
            sz = fmaf(em, h, sz);
        }
        st.Mm[p] = nMm; st.Ma[p] = nMa; st.Mz[p] = z1; st.z2[p] = zr; st.amax[p] = i1;
        st.Sm[p] = sm; st.Sa[p] = sa; st.Sz[p] = sz; st.T[p] = t;
    }
    if (GK) {
        // class-major: the table entry is warp-uniform, so one uniform branch per class selects the target class whose
        // running max takes this class's z (padded classes read table slack and carry kPadLogit: harmless)
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int l = lut[c0 + j];
#pragma unroll
            for (int k = 1; k < K; ++k) {
                if (l == k) {
#pragma unroll
                    for (int p = 0; p < P; ++p) zk[k][p] = fmaxf(zk[k][p], fmaf(0.5f, a[j][p], m[j][p]));
                }
            }
        }
    }
}

// What the fusion needs from a finished source, per pixel.
struct SourceResult {
    float kld;      // KL(softmax(m) || softmax(a)) = T/Sm - log Sm + log Sa  (= sum_c p1 (logp1 - logp2), log_softmax taken as
                    // (x - max) - log(sum) like ATen, so no large maxima are ever added back)
    float rz;       // Rz = Mm + Ma/2, the reference point of Sz
    float inv_sz;   // 1 / Sz:  softmax(z)_c = e^{z_c - Rz} * inv_sz
    float pmax;     // probability of the argmax class = e^{Mz - Rz} * inv_sz
    bool degenerate;  // the two heads' maxima sit more than kMaxSharedExpGap logit units above Mz: caller takes the slow path
};

// Largest Rz - Mz for which the shared exponentials are trusted.  e^{z-Rz} is formed as a product of two MUFU.EX2 results
// whose arguments grow with the gap; an argument of magnitude 2^e carries an absolute rounding error of 2^(e-24), i.e. a
// relative error of ~0.7 * 2^(e-24) in the exponential.  Up to a gap of 16 (arguments < 32) the confidence stays within
// ~3e-6 of the exact softmax; at 64 -- where Sz would start to underflow -- it would be 3e-5 (seen by the seeded sweep in
// tests/test_gpu_fuzz.py on logits of standard deviation 40).  Network logits are O(10): the slow path is never taken on
// the benchmark's inputs, it exists so that pathological inputs still meet the 1e-5 tolerance.
constexpr float kMaxSharedExpGap = 16.f;
// Largest |Mz| for the same.  The reference takes the softmax of z AFTER rounding it to fp32 (`pred + 0.5*pred_aux`), the
// shared exponentials work on the unrounded sum (m-Mm) + (a-Ma)/2 and on a separately rounded Rz: the two differ by up to
// ~1.5 ulp(z) in the exponent: at most 7.6e-6 relative up to |Mz| = 48 (|Rz| <= 64), but 3e-5 at |z| = 256.  Beyond 48 the
// slow path recomputes the reference's own expression e^{z_c - Mz} from the rounded z.  (A limit of 32 put ~3 pixels in 10^5
// of the benchmark's N(0, 6.4^2) fused logits on the slow path and cost K1 1.6 %; 48 is 7.5 sigma away.)
#ifndef MSPL_MAX_EXP_LOGIT
#define MSPL_MAX_EXP_LOGIT 48.f
#endif
constexpr float kMaxSharedExpLogit = MSPL_MAX_EXP_LOGIT;

template <int P>
MSPL_DEVINL SourceResult finish_source(const SourceStats<P>& st, int p) {
    SourceResult r;
    const float inv_sm = rcp_fast(st.Sm[p]);
    r.kld = fmaf(st.T[p], inv_sm, log_fast(st.Sa[p] * inv_sm));
    r.rz = fmaf(0.5f, st.Ma[p], st.Mm[p]);
    r.inv_sz = rcp_fast(st.Sz[p]);
    r.pmax = fminf(exp_neg(st.Mz[p] - r.rz) * r.inv_sz, 1.0f);     // numerator and its term of Sz round separately: clamp
    r.degenerate = !(r.rz - st.Mz[p] <= kMaxSharedExpGap) || !(fabsf(st.Mz[p]) <= kMaxSharedExpLogit);     // also catches NaN
    return r;
}

// Slow path for a degenerate pixel: recompute 1/sum_c e^{z_c - Mz} directly from global memory (rare; divergent).
// Deliberately out of line and not unrolled: it must not bloat the hot loop's instruction footprint.
static __device__ __noinline__ float recompute_pmax(const float* __restrict__ pm, const float* __restrict__ pa, int C, int64_t hw, float Mz) {
    float s = 0.f;
#pragma unroll 1
    for (int c = 0; c < C; ++c) s += exp_neg(fmaf(0.5f, __ldg(pa + c * hw), __ldg(pm + c * hw)) - Mz);
    return __frcp_rn(s);
}

template <int P>
MSPL_DEVINL void fill_pad(float (&v)[P]) {
#pragma unroll
    for (int p = 0; p < P; ++p) v[p] = kPadLogit;
}

// Load P pixels of CH class planes of both heads (classes c0..c0+cn-1; cn warp-uniform); a tail chunk (cn < CH) is
// padded with kPadLogit.  Full chunks take the predicate-free path.
template <int P, int CH>
MSPL_DEVINL void load_chunk(const float* __restrict__ pm, const float* __restrict__ pa, int64_t hw, int cn,
                            float (&m)[CH][P], float (&a)[CH][P]) {
    if (cn == CH) {
#pragma unroll
        for (int j = 0; j < CH; ++j) PixVec<P>::load(pm + j * hw, m[j]);
#pragma unroll
        for (int j = 0; j < CH; ++j) PixVec<P>::load(pa + j * hw, a[j]);
    } else {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if (j < cn) { PixVec<P>::load(pm + j * hw, m[j]); PixVec<P>::load(pa + j * hw, a[j]); }
            else { fill_pad<P>(m[j]); fill_pad<P>(a[j]); }
        }
    }
}

}  // namespace mspl
