#!/bin/bash
# Builds libagrs_b200 variants that differ only in the pair GEMM's pipeline shape (raw ring R, lo ring L, splitter groups G) into
# build_variants/ (git-ignored, shipped to the GPU box) for tools/gpu_pair_variants.sh / tools/gpu_round2_first.sh.
# Usage: tools/build_pair_variants.sh "4 2 2" "4 2 1" "4 2 1 1" ...   (a 4th field of 1 adds -DAGRS_SPLIT_EARLY_CONVERT: name suffix E;
# 5th field: AGRS_SPLIT_F32X2 (packed FMUL2 / FADD2 split, suffix xN), 6th field: AGRS_RAW_EARLY_RELEASE (suffix rN), 7th / 8th: suspend-time hints in ns of the epilogue's / TMA producer's long waits (suffixes eN, tN), 9th: the same hint on EVERY other wait (splitters, MMA issuer; suffix aN); all default
# to what the source says when omitted)
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /dev/null
mkdir -p build_variants
OBJ=autograd.rs_b200/build
for v in "$@"; do
  set -- $v; R=$1; L=$2; G=$3; E=${4:-0}; X2=${5:-}; RE=${6:-}; EW=${7:-}; TW=${8:-}; AW=${9:-}
  tag=R${R}L${L}G${G}; extra="-DAGRS_SPLIT_EARLY_CONVERT=$E"
  if [ "$E" = "1" ]; then tag=${tag}E; fi
  if [ -n "$X2" ]; then tag=${tag}x${X2}; extra="$extra -DAGRS_SPLIT_F32X2=$X2"; fi
  if [ -n "$RE" ]; then tag=${tag}r${RE}; extra="$extra -DAGRS_RAW_EARLY_RELEASE=$RE"; fi
  if [ -n "$EW" ]; then tag=${tag}e${EW}; extra="$extra -DAGRS_EPI_WAIT_NS=$EW"; fi
  if [ -n "$TW" ]; then tag=${tag}t${TW}; extra="$extra -DAGRS_TMA_WAIT_NS=$TW"; fi
  if [ -n "$AW" ]; then tag=${tag}a${AW}; extra="$extra -DAGRS_WAIT_NS=$AW"; fi
  o=build_variants/pair_${tag}.o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr \
    -I include -I autograd.rs_b200/csrc -DAGRS_PAIR_RAW_STAGES=$R -DAGRS_PAIR_LO_STAGES=$L -DAGRS_SPLIT_GROUPS=$G $extra \
    -c autograd.rs_b200/csrc/gemm_tcgen05_2cta.cu -o $o
  nvcc -shared -o build_variants/libagrs_b200_${tag}.so $(ls $OBJ/*.o | grep -v gemm_tcgen05_2cta.o) $o -gencode arch=compute_100a,code=sm_100a
  echo "built build_variants/libagrs_b200_${tag}.so"
done
