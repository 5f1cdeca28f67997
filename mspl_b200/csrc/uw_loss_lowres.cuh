// K4-lowres: K4 with the network's final bilinear upsample -- and its backward -- fused in (next-row component).
// Included by uw_loss.cu (needs LossWorkspace, finish_loss, uw_ce_pixel).
//
// ESPDNetUE ends with F.interpolate(main, size, 'bilinear', align_corners=True) from H/2 x W/2 and the same for the aux head
// from H/4 x W/4 (model/segmentation/espdnet_ue.py:301-302).  Training then reads both full-resolution tensors in the loss
// (uest_seg_multi_os.py:1020-1023) and autograd writes two full-resolution gradients only for upsample_bilinear2d_backward to
// fold them back.  Here the loss takes the PRE-upsample tensors: a CTA owns a tile of TR output rows of one image, stages the
// source rows they interpolate from in shared memory, evaluates the K4 closed forms per output pixel on the interpolated
// logits, keeps the tile's per-pixel gradients in shared memory, and applies the transposed interpolation as a separable
// GATHER (columns, then rows; fixed summation order) before adding the tile's share to the low-resolution gradients.
// A low-resolution row receives shares from at most two tiles when TR covers its footprint (TR = 8 does for the x2 / x4
// heads), and two floating-point adds onto a zeroed element commute, so the gradients are reproducible bit for bit.
// The full-resolution logits and gradients never exist in HBM.
#pragma once

namespace mspl {

struct LowresGeom {
    int hm, wm, ha, wa, H, W;       // main / aux source sizes, output size
    int TR, nrm, nra;               // output rows per tile; staged source rows per tile (upper bounds)
    int tiles_per_img;
    float rhm, rwm, rha, rwa;       // ATen scales (in-1)/(out-1)
    float inv_w;                    // 1/W
};

constexpr int kLowresThreads = 512;

struct HeadTables {                 // shared-memory tables of one head
    int* ix;        // [W]     left source column of output column x
    float* lx;      // [W]     its lambda
    int* xstart;    // [win+1] first output column whose left source column is i (W when none; xstart[win] = W)
    int* iy;        // [TR]    upper source row of each output row of the current tile
    float* ly;      // [TR]
    float* wx;      // [2W]    column weights of the transposed interpolation: source column xs reads the output columns
                    //         [xstart[xs-1], xstart[xs+1]); its weights start at xstart[xs] + xstart[xs-1] (ranges of
                    //         neighbours overlap, each output column feeds at most two source columns, so 2W entries do)
    float* wy;      // [nr*TR] row weights of the current tile: wy[r*TR + yl] = weight of output row yl into source row r0 + r
};

// weight with which an output coordinate whose left/upper tap is `i` (lambda `lam`) reads source index `target`
MSPL_DEVINL float tap_weight(int i, float lam, int lim, int target) {
    const int i2 = i + (i < lim - 1 ? 1 : 0);
    return (i == target ? 1.0f - lam : 0.0f) + (i2 == target ? lam : 0.0f);
}

MSPL_DEVINL void split_index(int p, int width, float inv_width, int& row, int& col) {
    row = (int)(((float)p + 0.5f) * inv_width);
    col = p - row * width;
    if (col < 0) { --row; col += width; }
    if (col >= width) { ++row; col -= width; }
}

// transposed interpolation of one head: g[k][yl*W + x] (the tile's gradients) -> d_lr[k][r0 + r][xs] += ...
// Weights come from the tables (2 shared loads + 1 FMA per term); the summation order is fixed.
template <int K>
MSPL_DEVINL void lowres_gather(const float* __restrict__ g, int g_stride, float* __restrict__ T, const HeadTables& tb, int rows, int TR,
                               int W, int hin, int win, int r0, int nrows, float* __restrict__ d_img) {
    const float inv_win = 1.0f / (float)win;
    // columns: T[k][yl][xs] = sum_x wx(x -> xs) g[k][yl][x]
    for (int p = threadIdx.x; p < rows * win; p += kLowresThreads) {
        int yl, xs;
        split_index(p, win, inv_win, yl, xs);
        const int xa = tb.xstart[xs > 0 ? xs - 1 : 0], xb = tb.xstart[xs + 1 < win ? xs + 1 : win];
        const float* wq = tb.wx + tb.xstart[xs] + xa;             // weights of x = xa, xa+1, ... (layout: HeadTables::wx)
        float acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.f;
        for (int x = xa; x < xb; ++x) {
            const float wgt = wq[x - xa];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = fmaf(wgt, g[(size_t)k * g_stride + yl * W + x], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) T[(size_t)k * TR * win + p] = acc[k];
    }
    __syncthreads();
    // rows: d[k][r0 + r][xs] += sum_yl wy(yl -> r0 + r) T[k][yl][xs]
    for (int p = threadIdx.x; p < nrows * win; p += kLowresThreads) {
        int r, xs;
        split_index(p, win, inv_win, r, xs);
        float acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.f;
        for (int yl = 0; yl < rows; ++yl) {
            const float wgt = tb.wy[r * TR + yl];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = fmaf(wgt, T[(size_t)k * TR * win + yl * win + xs], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) atomicAdd(d_img + ((size_t)k * hin + r0) * win + p, acc[k]);
    }
    __syncthreads();
}

template <int K, bool BWD, typename TT>
__global__ void __launch_bounds__(kLowresThreads, 1) uw_ce_lowres_kernel(const float* __restrict__ main_lr, const float* __restrict__ aux_lr,
                                                                         const TT* __restrict__ target, const float* __restrict__ cw,
                                                                         int64_t n_img, const LowresGeom gm, float alpha, double inv_n,
                                                                         float gscale, float* __restrict__ out3, float* __restrict__ d_main,
                                                                         float* __restrict__ d_aux, LossWorkspace* ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int W = gm.W, TR = gm.TR, hm = gm.hm, wm = gm.wm, ha = gm.ha, wa = gm.wa;
    // ---- shared-memory carve-up (all 4-byte elements) ----
    float* s_g = reinterpret_cast<float*>(smem_raw);                        // [2][K][TR*W] tile gradients (BWD)
    const int g_stride = TR * W;
    float* s_src = s_g + (BWD ? 2 * K * g_stride : 0);                      // staged source rows; reused as T by the gather
    const int src_main = K * gm.nrm * wm, src_aux = K * gm.nra * wa;
    const int t_need = BWD ? K * TR * (wm > wa ? wm : wa) : 0;
    const int src_floats = src_main + src_aux > t_need ? src_main + src_aux : t_need;
    HeadTables tm, ta;
    tm.ix = reinterpret_cast<int*>(s_src + src_floats);
    tm.lx = reinterpret_cast<float*>(tm.ix + W);
    ta.ix = reinterpret_cast<int*>(tm.lx + W);
    ta.lx = reinterpret_cast<float*>(ta.ix + W);
    tm.xstart = reinterpret_cast<int*>(ta.lx + W);
    ta.xstart = tm.xstart + (wm + 1);
    tm.iy = ta.xstart + (wa + 1);
    tm.ly = reinterpret_cast<float*>(tm.iy + TR);
    ta.iy = reinterpret_cast<int*>(tm.ly + TR);
    ta.ly = reinterpret_cast<float*>(ta.iy + TR);
    float* s_w = ta.ly + TR;
    tm.wx = s_w + K;
    ta.wx = tm.wx + (BWD ? 2 * W : 0);
    tm.wy = ta.wx + (BWD ? 2 * W : 0);
    ta.wy = tm.wy + (BWD ? gm.nrm * TR : 0);

    // ---- per-CTA column tables (the same for every tile) ----
    if (threadIdx.x < K) s_w[threadIdx.x] = cw[threadIdx.x];
    for (int i = threadIdx.x; i <= wm; i += kLowresThreads) tm.xstart[i] = W;
    for (int i = threadIdx.x; i <= wa; i += kLowresThreads) ta.xstart[i] = W;
    for (int x = threadIdx.x; x < W; x += kLowresThreads) {
        const float fm = gm.rwm * (float)x, fa = gm.rwa * (float)x;
        tm.ix[x] = (int)fm; tm.lx[x] = fm - (float)(int)fm;
        ta.ix[x] = (int)fa; ta.lx[x] = fa - (float)(int)fa;
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += kLowresThreads) {
        if (x == 0 || tm.ix[x] != tm.ix[x - 1]) tm.xstart[tm.ix[x]] = x;
        if (x == 0 || ta.ix[x] != ta.ix[x - 1]) ta.xstart[ta.ix[x]] = x;
    }
    __syncthreads();
    if (BWD) {
        for (int xs = threadIdx.x; xs < wm; xs += kLowresThreads) {
            const int xa = tm.xstart[xs > 0 ? xs - 1 : 0], xb = tm.xstart[xs + 1 < wm ? xs + 1 : wm];
            float* wq = tm.wx + tm.xstart[xs] + xa;
            for (int x = xa; x < xb; ++x) wq[x - xa] = tap_weight(tm.ix[x], tm.lx[x], wm, xs);
        }
        for (int xs = threadIdx.x; xs < wa; xs += kLowresThreads) {
            const int xa = ta.xstart[xs > 0 ? xs - 1 : 0], xb = ta.xstart[xs + 1 < wa ? xs + 1 : wa];
            float* wq = ta.wx + ta.xstart[xs] + xa;
            for (int x = xa; x < xb; ++x) wq[x - xa] = tap_weight(ta.ix[x], ta.lx[x], wa, xs);
        }
    }
    __syncthreads();

    const float inv_nf = (float)inv_n;
    const int64_t n_tiles = n_img * gm.tiles_per_img;
    const size_t plane_m = (size_t)hm * wm, plane_a = (size_t)ha * wa;
    double acc_ce = 0, acc_d = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n = tile / gm.tiles_per_img;
        const int y0 = (int)(tile - n * gm.tiles_per_img) * TR;
        const int rows = gm.H - y0 < TR ? gm.H - y0 : TR;
        // source rows this tile interpolates from (uniform across the CTA)
        const int r0m = (int)(gm.rhm * (float)y0), r0a = (int)(gm.rha * (float)y0);
        int r1m = (int)(gm.rhm * (float)(y0 + rows - 1)) + 1, r1a = (int)(gm.rha * (float)(y0 + rows - 1)) + 1;
        r1m = r1m < hm - 1 ? r1m : hm - 1;
        r1a = r1a < ha - 1 ? r1a : ha - 1;
        const int nrm_t = r1m - r0m + 1, nra_t = r1a - r0a + 1;
        if (threadIdx.x < rows) {
            const float fm = gm.rhm * (float)(y0 + threadIdx.x), fa = gm.rha * (float)(y0 + threadIdx.x);
            tm.iy[threadIdx.x] = (int)fm; tm.ly[threadIdx.x] = fm - (float)(int)fm;
            ta.iy[threadIdx.x] = (int)fa; ta.ly[threadIdx.x] = fa - (float)(int)fa;
        }
        if (BWD) {      // row weights of this tile (same arithmetic as the taps above)
            for (int i = threadIdx.x; i < (nrm_t + nra_t) * TR; i += kLowresThreads) {
                const bool is_m = i < nrm_t * TR;
                const int j = is_m ? i : i - nrm_t * TR, r = j / TR, yl = j - r * TR;
                const float f = (is_m ? gm.rhm : gm.rha) * (float)(y0 + yl);
                const float wgt = yl < rows ? tap_weight((int)f, f - (float)(int)f, is_m ? hm : ha, (is_m ? r0m : r0a) + r) : 0.f;
                (is_m ? tm.wy : ta.wy)[j] = wgt;
            }
        }
        // ---- stage the source rows: per class one contiguous block of whole rows ----
        float* src_m = s_src;
        float* src_a = s_src + src_main;
        for (int k = 0; k < K; ++k) {
            const float* pm = main_lr + ((size_t)n * K + k) * plane_m + (size_t)r0m * wm;
            const float* pa = aux_lr + ((size_t)n * K + k) * plane_a + (size_t)r0a * wa;
            for (int i = threadIdx.x; i < nrm_t * wm; i += kLowresThreads) src_m[k * gm.nrm * wm + i] = __ldg(pm + i);
            for (int i = threadIdx.x; i < nra_t * wa; i += kLowresThreads) src_a[k * gm.nra * wa + i] = __ldg(pa + i);
        }
        __syncthreads();
        // ---- per output pixel: interpolate both heads, K4 closed forms ----
        const TT* tg = target + ((size_t)n * gm.H + y0) * W;
        for (int p = threadIdx.x; p < rows * W; p += kLowresThreads) {
            int yl, x;
            split_index(p, W, gm.inv_w, yl, x);
            BilinearTap bm, ba;
            bm.w1 = tm.lx[x]; bm.w0 = 1.0f - bm.w1; bm.h1 = tm.ly[yl]; bm.h0 = 1.0f - bm.h1;
            bm.dx = tm.ix[x] < wm - 1 ? 1 : 0; bm.dy = tm.iy[yl] < hm - 1 ? wm : 0;
            bm.o00 = (tm.iy[yl] - r0m) * wm + tm.ix[x];
            ba.w1 = ta.lx[x]; ba.w0 = 1.0f - ba.w1; ba.h1 = ta.ly[yl]; ba.h0 = 1.0f - ba.h1;
            ba.dx = ta.ix[x] < wa - 1 ? 1 : 0; ba.dy = ta.iy[yl] < ha - 1 ? wa : 0;
            ba.o00 = (ta.iy[yl] - r0a) * wa + ta.ix[x];
            float m[K], a[K], gmain[K], gaux[K], l, D;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                m[k] = bilinear(src_m + k * gm.nrm * wm, bm);
                a[k] = bilinear(src_a + k * gm.nra * wa, ba);
            }
            uw_ce_pixel<K, BWD>(m, a, (long long)__ldcs(tg + p), s_w, alpha, gscale, inv_nf, l, D, gmain, gaux);
            acc_ce += (double)l;
            acc_d += (double)D;
            if (BWD) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    s_g[k * g_stride + p] = gmain[k];
                    s_g[(K + k) * g_stride + p] = gaux[k];
                }
            }
        }
        __syncthreads();
        if (BWD) {
            lowres_gather<K>(s_g, g_stride, s_src, tm, rows, TR, W, hm, wm, r0m, nrm_t, d_main + (size_t)n * K * plane_m);
            lowres_gather<K>(s_g + K * g_stride, g_stride, s_src, ta, rows, TR, W, ha, wa, r0a, nra_t, d_aux + (size_t)n * K * plane_a);
        }
    }
    finish_loss<3, kLowresThreads>(acc_ce, acc_d, ws, inv_n, alpha, out3);
}

// bytes of dynamic shared memory the kernel needs for a geometry (mirrors the carve-up above)
inline size_t lowres_smem_bytes(const LowresGeom& g, int K, bool bwd) {
    const size_t gsz = bwd ? (size_t)2 * K * g.TR * g.W : 0;
    const size_t src = (size_t)K * ((size_t)g.nrm * g.wm + (size_t)g.nra * g.wa);
    const size_t tn = bwd ? (size_t)K * g.TR * (g.wm > g.wa ? g.wm : g.wa) : 0;
    const size_t tables = (size_t)4 * g.W + (g.wm + 1) + (g.wa + 1) + 4 * g.TR + K +
                          (bwd ? (size_t)4 * g.W + (size_t)(g.nrm + g.nra) * g.TR : 0);
    return 4 * (gsz + (src > tn ? src : tn) + tables);
}

}  // namespace mspl
