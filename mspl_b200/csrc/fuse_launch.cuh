// Host-side launch helpers for the two K1 kernels (shared by libmspl_b200.so and tools/k1_sweep).
#pragma once
#include "fuse_kernel.cuh"

namespace mspl {

struct DeviceInfo {
    int sms = kNumSMs;
    bool ok = false;
};

inline DeviceInfo device_info() {
    DeviceInfo d;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return d;
    if (cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return d;
    d.ok = true;
    return d;
}

// Persistent grid: one wave of resident CTAs, each striding over the pixel tiles.
template <int P, int THREADS, typename Kern>
int launch_fuse_direct(Kern kern, const FuseParams& prm, cudaStream_t stream) {
    const size_t smem = fuse_tally_smem_bytes(prm.K);
    const DeviceInfo di = device_info();
    int per_sm = 0;
    if (!di.ok || cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return MSPL_ERR_CUDA;
    }
    const int64_t n_groups = prm.n_img * (prm.hw / P);
    const int64_t n_tiles = (n_groups + THREADS - 1) / THREADS;
    const int64_t cap = (int64_t)di.sms * per_sm;
    kern<<<(unsigned)(n_tiles < cap ? n_tiles : cap), THREADS, smem, stream>>>(prm);
    return launch_status();
}

// One CTA per SM (the shared-memory ring fills the SM), striding over the tiles.
template <typename Cfg, typename Kern>
int launch_fuse_tma(Kern kern, const FuseParams& prm, cudaStream_t stream) {
    const size_t smem = Cfg::smem_bytes(prm.K);
    const DeviceInfo di = device_info();
    if (!di.ok || cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return MSPL_ERR_CUDA;
    }
    const int64_t n_tiles = prm.n_img * ((prm.hw + Cfg::kTilePix - 1) / Cfg::kTilePix);
    kern<<<(unsigned)(n_tiles < di.sms ? n_tiles : di.sms), Cfg::kThreads, smem, stream>>>(prm);
    return launch_status();
}

// The bulk-copy path needs 16-byte aligned class rows (bases 16-byte aligned, pixels_per_image % 4 == 0).
template <typename Cfg>
bool tma_eligible(const FuseParams& prm) {
    if (prm.hw % 4 != 0) return false;
    for (int s = 0; s < prm.S; ++s)
        if (!aligned_to(prm.main[s], 16) || !aligned_to(prm.aux[s], 16) || (prm.kld[s] && !aligned_to(prm.kld[s], 16))) return false;
    return aligned_to(prm.label, 4) && (!prm.conf || aligned_to(prm.conf, 16)) && (!prm.unc || aligned_to(prm.unc, 16));
}

}  // namespace mspl
