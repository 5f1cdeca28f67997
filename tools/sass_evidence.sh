#!/bin/bash
# SASS evidence for profiles/: the mnemonics that show what the production kernels are made of (cuobjdump -sass on the built
# objects; runs without a GPU).  Usage: bash tools/sass_evidence.sh > profiles/r01_sass_evidence.txt
obj=mspl_b200/lib/obj
count() {   # object, mangled function, label
  cuobjdump -sass -fun "$2" "$obj/$1" 2>/dev/null > /tmp/_sass.txt
  printf "%s\n  %s\n  instructions: %s\n" "$3" "$2" "$(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' /tmp/_sass.txt)"
  for m in UBLKCP SYNCS MUFU.EX2 MUFU.RCP MUFU.LG2 LDS.64 FMNMX3 "FMNMX " FFMA FADD "STG.E" ATOMS ATOMG REDUX "LDG.E" "STL" "LDL"; do
    printf "    %-9s %s\n" "$m" "$(grep -c "$m" /tmp/_sass.txt)"
  done
  grep -m1 "UBLKCP" /tmp/_sass.txt | sed 's/^\s*/    e.g. /; s/\s*\/\* 0x.*//'
  echo
}
echo "# cuobjdump -sass mnemonic counts (static instruction counts of the whole kernel, not per-pixel)."
echo "# UBLKCP = cp.async.bulk global->shared (the TMA bulk-copy unit), SYNCS = mbarrier ops, MUFU = special-function unit."
echo
count fuse_sources.o _ZN4mspl23fuse_sources_tma_kernelILi15ELi2ELi5ELi4ELi5ELb0ELb1EEEvNS_10FuseParamsE "K1 production kernel (vote 'all'): fuse_sources_tma_kernel<15 consumer warps, P=2, CH=5, 4 stages, K<=5, GK=false, TOP2=true>"
count fuse_sources.o _ZN4mspl23fuse_sources_tma_kernelILi19ELi2ELi5ELi3ELi5ELb1ELb1EEEvNS_10FuseParamsE "K1 per-class-probability kernel ('half'/int with confidence, 'prob'): <19, 2, 5, 3, 5, GK=true, TOP2=true>"
count fuse_sources.o _ZN4mspl22fuse_labels_tma_kernelILi15ELi2ELi5ELi4ELi5EEEvNS_10FuseParamsE "K1 labels-only kernel (reference-exact output): fuse_labels_tma_kernel<15, 2, 5, 4, 5>"
count uw_loss.o _ZN4mspl18uw_ce_fused_kernelILi2ELi5ELb1ElLb0EEEvPKfS2_PKT2_S2_llfdfPfS6_S6_PNS_13LossWorkspaceEPy "K4 forward+backward, K=5, int64 targets: uw_ce_fused_kernel<P=2, K=5, BWD, int64, IOU=false>"
count uw_loss.o _ZN4mspl18uw_ce_fused_kernelILi2ELi5ELb1EhLb1EEEvPKfS2_PKT2_S2_llfdfPfS6_S6_PNS_13LossWorkspaceEPy "K4 forward+backward with the metric counts, uint8 targets: uw_ce_fused_kernel<2, 5, BWD, uint8, IOU=true>"
echo "# register / shared-memory use (cuobjdump -res-usage)"
for f in fuse_sources uw_loss; do cuobjdump -res-usage $obj/$f.o 2>/dev/null | grep -A1 -E "fuse_sources_tma_kernelILi15ELi2ELi5ELi4ELi5ELb0ELb1|fuse_sources_tma_kernelILi19ELi2ELi5ELi3ELi5ELb1ELb1|fuse_labels_tma_kernelILi15ELi2ELi5ELi4ELi5E|uw_ce_fused_kernelILi2ELi5ELb1ElLb0|uw_ce_fused_kernelILi2ELi5ELb1EhLb1" | grep -v "^--" | sed 's/^ Function /  /; s/^  REG/      REG/'; done
