#!/bin/bash
# cuobjdump -sass of the hot kernels of libphet_b200.so (the instantiations config 4 / config 5 launch), gzipped
# under profiles/<round>_sass/.  usage: bash profiles/tools/dump_sass.sh r2
set -e
round=${1:-r2}
out=profiles/${round}_sass
mkdir -p $out
lib=phet_b200/libphet_b200.so
for f in \
  _ZN4phet8k_step1pIfLi16ELb1ELb0EEEvNS_13Step1PairArgsE:k_step1p_f32_nbw16_staged \
  _ZN4phet8k_step1pIfLi16ELb1ELb1EEEvNS_13Step1PairArgsE:k_step1p_f32_nbw16_big \
  _ZN4phet8k_step3pIfLi10ELi3EEEvNS_13Step3PairArgsE:k_step3p_f32_e10_p3 \
  _ZN4phet8k_step3sIfLi10ELb1ELb1EEEvNS_16Step3ScatterArgsE:k_step3s_f32_r10_ovf_device \
  _ZN4phet8k_step3sIfLi5ELb0ELb0EEEvNS_16Step3ScatterArgsE:k_step3s_f32_r5_table \
  _ZN4phet21k_normalize_transposeIffEEvPKT_llllPKiPKdS7_iPT0_:k_normalize_transpose_f32 \
  _ZN4phet13k_class_statsIfLb1EEEvNS_14ClassStatsArgsE:k_class_stats_f32_fast \
  _ZN4phet14k_kruskal_sortIfLi6ELb1EEEvNS_9AnovaArgsEi:k_kruskal_sort_f32_e6_staged \
  _ZN4phet15k_fy_apply_smemItEEvNS_6FyArgsE:k_fy_apply_smem_u16 \
  _ZN4phet12k_build_idxTEPKilllllPt:k_build_idxT ; do
  sym=${f%%:*}; name=${f##*:}
  cuobjdump -sass -fun "$sym" $lib 2>/dev/null | gzip -9 > $out/$name.sass.gz
  echo "$name $(zcat $out/$name.sass.gz | grep -cE "^\s+/\*[0-9a-f]{4,5}\*/") instructions"
done
