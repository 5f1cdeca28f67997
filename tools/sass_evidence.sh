#!/bin/bash
# SASS evidence for profiles/: the mnemonics that show what the production kernels are made of (cuobjdump -sass on the built
# objects; runs without a GPU).  Usage: bash tools/sass_evidence.sh > profiles/r02_sass_evidence.txt
obj=mspl_b200/lib/obj
count() {   # object, mangled function, label
  cuobjdump -sass -fun "$2" "$obj/$1" 2>/dev/null > /tmp/_sass.txt
  printf "%s\n  %s\n  instructions: %s\n" "$3" "$2" "$(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' /tmp/_sass.txt)"
  for m in UBLKCP SYNCS MUFU.EX2 MUFU.RCP MUFU.LG2 "LDS.64" "LDS " FMNMX3 "FMNMX " FFMA2 FADD2 FMUL2 "FFMA " "FADD " "FMUL " "STG.E" ATOMS ATOMG REDUX "LDG.E" "STL" "LDL" "UCGABAR\|BAR.SYNC"; do
    printf "    %-9s %s\n" "$m" "$(grep -c "$m" /tmp/_sass.txt)"
  done
  grep -m1 "UBLKCP" /tmp/_sass.txt | sed 's/^\s*/    e.g. /; s/\s*\/\* 0x.*//'
  echo
}
echo "# cuobjdump -sass mnemonic counts (static instruction counts of the whole kernel, not per-pixel)."
echo "# UBLKCP = cp.async.bulk global->shared (the TMA bulk-copy unit), SYNCS = mbarrier ops, MUFU = special-function unit."
echo
count fuse_sources.o _ZN4mspl23fuse_sources_tma_kernelILi15ELi2ELi5ELi4ELi5ELb0EEEvNS_10FuseParamsE "K1 production kernel (vote 'all'): fuse_sources_tma_kernel<15 consumer warps, P=2, CH=5, 4 stages, K<=5, GK=false> -- FFMA2/FADD2/FMUL2 = packed f32x2 math"
count fuse_sources.o _ZN4mspl23fuse_sources_tma_kernelILi15ELi2ELi5ELi4ELi5ELb1EEEvNS_10FuseParamsE "K1 per-class-probability policies ('half'/int with confidence, 'prob'): same shape, GK=true"
count fuse_sources.o _ZN4mspl22fuse_labels_tma_kernelILi15ELi2ELi5ELi4ELi5EEEvNS_10FuseParamsE "K1 labels-only kernel (reference-exact output): fuse_labels_tma_kernel<15, 2, 5, 4, 5>"
count fuse_sources.o _ZN4mspl26fuse_sources_lowres_kernelILi15ELi2ELi5ELi4ELi5ELb0ELi768ELi384EEEvNS_10FuseParamsE "K1-lowres (fused final upsample), fixed class strides: fuse_sources_lowres_kernel<15, 2, 5, 4, 5, GK=false, 768, 384>"
count thresholds.o _ZN4mspl19cand_resolve_kernelEPKhPKfPKjPKyliiPNS_10RadixStateEPfPhSB_Py "threshold tail in one launch: cand_resolve_kernel (8-CTA cluster per class, histograms combined over distributed shared memory)"
count uw_loss.o _ZN4mspl18uw_ce_fused_kernelILi2ELi5ELb1ElLb0EEEvPKfS2_PKT2_S2_llfdfPfS6_S6_PNS_13LossWorkspaceEPy "K4 forward+backward, K=5, int64 targets: uw_ce_fused_kernel<P=2, K=5, BWD, int64, IOU=false>"
count uw_loss.o _ZN4mspl18uw_ce_fused_kernelILi2ELi5ELb1EhLb1EEEvPKfS2_PKT2_S2_llfdfPfS6_S6_PNS_13LossWorkspaceEPy "K4 forward+backward with the metric counts, uint8 targets: uw_ce_fused_kernel<2, 5, BWD, uint8, IOU=true>"
echo "# register / shared-memory use (cuobjdump -res-usage)"
for f in fuse_sources thresholds uw_loss; do cuobjdump -res-usage $obj/$f.o 2>/dev/null | grep -A1 -E "fuse_sources_tma_kernelILi15ELi2ELi5ELi4ELi5ELb[01]E|fuse_sources_lowres_kernelILi15ELi2ELi5ELi4ELi5ELb0ELi768|fuse_labels_tma_kernelILi15ELi2ELi5ELi4ELi5E|cand_resolve_kernel|uw_ce_fused_kernelILi2ELi5ELb1ElLb0|uw_ce_fused_kernelILi2ELi5ELb1EhLb1" | grep -v "^--" | sed 's/^ Function /  /; s/^  REG/      REG/'; done
