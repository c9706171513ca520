#!/bin/bash
# Builds libagrs_b200 variants that differ only in the pair GEMM's pipeline shape (raw ring R, lo ring L, splitter groups G) into
# build_variants/ (git-ignored, shipped to the GPU box) for tools/gpu_pair_variants.sh.  Usage: tools/build_pair_variants.sh "4 3 2" "5 2 2" ...
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /dev/null
mkdir -p build_variants
OBJ=autograd.rs_b200/build
for v in "$@"; do
  set -- $v; R=$1; L=$2; G=$3
  o=build_variants/pair_R${R}L${L}G${G}.o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr \
    -I include -I autograd.rs_b200/csrc -DAGRS_PAIR_RAW_STAGES=$R -DAGRS_PAIR_LO_STAGES=$L -DAGRS_SPLIT_GROUPS=$G \
    -c autograd.rs_b200/csrc/gemm_tcgen05_2cta.cu -o $o
  nvcc -shared -o build_variants/libagrs_b200_R${R}L${L}G${G}.so $(ls $OBJ/*.o | grep -v gemm_tcgen05_2cta.o) $o -gencode arch=compute_100a,code=sm_100a
  echo "built build_variants/libagrs_b200_R${R}L${L}G${G}.so"
done
