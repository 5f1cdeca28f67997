// Bilinear align_corners=True interpolation with ATen's upsample_bilinear2d arithmetic, shared by the fused-upsample
// kernels (K1-lowres in fuse_kernel.cuh, K4-lowres in uw_loss.cu): the networks' closing
// F.interpolate(..., mode='bilinear', align_corners=True) (model/segmentation/espdnet_ue.py:301-302).
//   src = dst * (in-1)/(out-1); i = (int)src; lambda = src - i;  val = h0*(w0*v00 + w1*v01) + h1*(w0*v10 + w1*v11)
#pragma once
#include "common.cuh"

namespace mspl {

struct BilinearTap {           // one output pixel's taps into one head, relative to the first staged source row
    int o00, dx, dy;           // offset of v00, +dx for the right column, +dy for the lower row
    float w0, w1, h0, h1;
};

MSPL_DEVINL float lowres_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

MSPL_DEVINL BilinearTap make_tap(int y, int x, int hin, int win, float rh, float rw, int first_row) {
    BilinearTap t;
    const float h1r = rh * (float)y, w1r = rw * (float)x;
    const int h1 = (int)h1r, w1 = (int)w1r;
    t.h1 = h1r - (float)h1; t.h0 = 1.0f - t.h1;
    t.w1 = w1r - (float)w1; t.w0 = 1.0f - t.w1;
    t.dy = (h1 < hin - 1) ? win : 0;
    t.dx = (w1 < win - 1) ? 1 : 0;
    t.o00 = (h1 - first_row) * win + w1;
    return t;
}

MSPL_DEVINL float bilinear(const float* __restrict__ s, const BilinearTap& t) {
    const float v00 = s[t.o00], v01 = s[t.o00 + t.dx], v10 = s[t.o00 + t.dy], v11 = s[t.o00 + t.dy + t.dx];
    return t.h0 * (t.w0 * v00 + t.w1 * v01) + t.h1 * (t.w0 * v10 + t.w1 * v11);
}

}  // namespace mspl
