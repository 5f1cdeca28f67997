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
MSPL_DEVINL float exp_neg(float t) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t * kLog2e));
    return r;
}

template <int P>
struct SourceStats {
    float Mm[P], Sm[P], Ma[P], Sa[P], Mz[P], Sz[P], T[P], z2[P];
    int amax[P];
    MSPL_DEVINL void reset() {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            Mm[p] = Ma[p] = Mz[p] = z2[p] = -INFINITY;
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
        float z[CH];
        float cm = -INFINITY, ca = -INFINITY;
        float z1 = st.Mz[p], zr = st.z2[p];
        int i1 = st.amax[p];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            // z exactly as the reference forms it: 0.5*a is exact in fp32, so the fused multiply-add
            // rounds once, just like `pred + 0.5 * pred_aux`.
            z[j] = fmaf(0.5f, a[j][p], m[j][p]);
            cm = fmaxf(cm, m[j][p]);
            ca = fmaxf(ca, a[j][p]);
            if (TOP2) zr = fmaxf(zr, fminf(z1, z[j]));
            i1 = (z[j] > z1) ? (c0 + j) : i1;      // strict >: lowest index wins ties, as np.argmax
            z1 = fmaxf(z1, z[j]);
        }
        const float nMm = fmaxf(st.Mm[p], cm), nMa = fmaxf(st.Ma[p], ca);
        float sm = 0.f, sa = 0.f, sz = 0.f, t = 0.f;
        if (!first) {   // rescale what earlier chunks accumulated to the new maxima (one exp per stream)
            const float dm = st.Mm[p] - nMm, da = st.Ma[p] - nMa;
            const float rm = exp_neg(dm);
            sm = st.Sm[p] * rm;
            t = rm * fmaf(st.Sm[p], dm - da, st.T[p]);
            sa = st.Sa[p] * exp_neg(da);
            sz = st.Sz[p] * exp_neg(st.Mz[p] - z1);
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const float tm = m[j][p] - nMm, ta = a[j][p] - nMa;
            const float em = exp_neg(tm);
            sm += em;
            t = fmaf(em, tm - ta, t);
            sa += exp_neg(ta);
            sz += exp_neg(z[j] - z1);
        }
        if (GK) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int l = lut[c0 + j];          // warp-uniform; padded classes read table slack and carry kPadLogit
#pragma unroll
                for (int k = 1; k < K; ++k) zk[k][p] = fmaxf(zk[k][p], l == k ? z[j] : -INFINITY);
            }
        }
        st.Mm[p] = nMm; st.Ma[p] = nMa; st.Mz[p] = z1; st.z2[p] = zr; st.amax[p] = i1;
        st.Sm[p] = sm; st.Sa[p] = sa; st.Sz[p] = sz; st.T[p] = t;
    }
}

// KL(softmax(m) || softmax(a)) from the finished stats (IEEE division and accurate logf: the two log terms
// cancel to O(KLD), so their absolute error is what bounds the 1e-5 relative agreement).
template <int P>
MSPL_DEVINL float kld_of(const SourceStats<P>& st, int p) {
    return st.T[p] / st.Sm[p] - logf(st.Sm[p]) + logf(st.Sa[p]);
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
