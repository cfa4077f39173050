#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (troylee/nnet-asr = TNet v1.8)
# from the sources where they lie under /root/reference into oracle/_ref/.
#
#   oracle/_ref/TNet          CPU trainer   (src/TNet.cc + KaldiLib + TNetLib)   -> cpu baseline + oracle pin
#   oracle/_ref/TFeaCat       CPU forward-only tool (src/TFeaCat.cc, same libs)  -> golden for the TFeaCatCu drop-in
#   oracle/_ref/TNorm         CPU mean/variance estimator (src/TNorm.cc)         -> golden for the TNormCu drop-in
#   oracle/_ref/RefIoDump     oracle/ref_tools/io_dump.cc (ours) over the reference's FeatureRepository / LabelRepository
#                             -> golden for the drop-in's HTK / script-file / MLF readers
#   oracle/_ref/RefCacheDump  oracle/ref_tools/cache_dump.cc (ours) over the reference's CPU frame cache (TNetLib/Cache.cc)
#                             -> bit-exact golden for the oracle's cache (fill, leftovers, permutation, bunches, discards)
#   oracle/_ref/TNetCu        GPU trainer   (src/TNetCu.cc + CuBaseLib + CuTNetLib, legacy cuBLAS) -> golden on B200
#   oracle/_ref/TRbmCu, TRecurrentCu        same libs
#
# No reference source is copied into the repo.  Objects are compiled in a scratch
# dir under /tmp.  GotoBLAS2 (the reference's BLAS, src/Makefile:22-28) is not
# vendored upstream; OpenBLAS 0.3.15 from the python env stands in.
#
# Two upstream compile errors with a modern toolchain are patched on a /tmp copy
# of exactly two files:
#   (1) src/CuBaseLib/cukernels.cu:183 uses _sum_reduce before its definition (:278)
#       -> forward declaration added;
#   (2) src/CuTNetLib/cuClusterLinearity.cc:82-93 duplicated tail of Update()
#       (T. Lee's unfinished component, unused by any config) -> file left out of the
#       link and its factory case stubbed.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
WORK=${WORK:-/tmp/tnet_ref_build}
SP=$(python -c 'import site;print(site.getsitepackages()[0])')
OBLAS_DIR="$SP/opencv_python_headless.libs"
OBLAS=$(ls "$OBLAS_DIR"/libopenblasp-*.so | head -1)
CUDA=${CUDA:-/usr/local/cuda}

[ -d "$REF/src" ] || { echo "reference not present at $REF (GPU box?) - nothing to build"; exit 0; }
mkdir -p "$OUT" "$WORK/cpu" "$WORK/gpu"

CXXF="-std=gnu++98 -O2 -DHAVE_ATLAS -fpermissive -w -fPIC"
INC="-I$REF/src/KaldiLib -I$REF/src/TNetLib"

build_cpu() {
  local objs=()
  for f in "$REF"/src/KaldiLib/*.cc "$REF"/src/TNetLib/*.cc "$REF"/src/TNet.cc; do
    local o="$WORK/cpu/$(basename "${f%.cc}").o"
    if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then g++ $CXXF $INC -c "$f" -o "$o" & fi
    objs+=("$o")
  done
  wait
  g++ -o "$OUT/TNet" "${objs[@]}" "$OBLAS" -lpthread -Wl,--disable-new-dtags -Wl,-rpath,"$OBLAS_DIR"
  echo "built $OUT/TNet"
  # forward-only tool of the CPU library (src/TFeaCat.cc): golden vectors for the TFeaCatCu drop-in
  g++ $CXXF $INC -c "$REF/src/TFeaCat.cc" -o "$WORK/cpu/TFeaCat.o"
  local lobjs=()
  for o in "${objs[@]}"; do case "$o" in */TNet.o) ;; *) lobjs+=("$o");; esac; done
  g++ -o "$OUT/TFeaCat" "$WORK/cpu/TFeaCat.o" "${lobjs[@]}" "$OBLAS" -lpthread -Wl,--disable-new-dtags -Wl,-rpath,"$OBLAS_DIR"
  echo "built $OUT/TFeaCat"
  # global mean/variance estimator of the CPU library (src/TNorm.cc): golden vectors for the TNormCu drop-in
  g++ $CXXF $INC -c "$REF/src/TNorm.cc" -o "$WORK/cpu/TNorm.o"
  g++ -o "$OUT/TNorm" "$WORK/cpu/TNorm.o" "${lobjs[@]}" "$OBLAS" -lpthread -Wl,--disable-new-dtags -Wl,-rpath,"$OBLAS_DIR"
  echo "built $OUT/TNorm"
  # our dumper over the reference's own HTK / script-file / MLF readers (oracle/ref_tools/io_dump.cc): differential fixture for the
  # drop-in's front end (nnet-asr_b200/host/io.h)
  g++ $CXXF $INC -c "$HERE/ref_tools/io_dump.cc" -o "$WORK/cpu/io_dump.o"
  g++ -o "$OUT/RefIoDump" "$WORK/cpu/io_dump.o" "${lobjs[@]}" "$OBLAS" -lpthread -Wl,--disable-new-dtags -Wl,-rpath,"$OBLAS_DIR"
  echo "built $OUT/RefIoDump"
  # our driver over the reference's CPU frame cache (oracle/ref_tools/cache_dump.cc): bit-exact fixture for the oracle's cache
  g++ $CXXF $INC -c "$HERE/ref_tools/cache_dump.cc" -o "$WORK/cpu/cache_dump.o"
  g++ -o "$OUT/RefCacheDump" "$WORK/cpu/cache_dump.o" "${lobjs[@]}" "$OBLAS" -lpthread -Wl,--disable-new-dtags -Wl,-rpath,"$OBLAS_DIR"
  echo "built $OUT/RefCacheDump"
}

build_gpu() {
  [ -x "$CUDA/bin/nvcc" ] || { echo "no nvcc: skipping GPU reference"; return 0; }
  local G="$WORK/gpu"
  local CINC="$INC -I$REF/src/CuBaseLib -I$REF/src/CuTNetLib -I$CUDA/include"
  # patched copies (scratch only)
  { echo 'template<typename T> __device__ static T _sum_reduce(T buffer[]);'; cat "$REF/src/CuBaseLib/cukernels.cu"; } > "$G/cukernels_patched.cu"
  sed -e 's@#include "cukernels.h"@#include "'"$REF"'/src/CuBaseLib/cukernels.h"@' -i "$G/cukernels_patched.cu"
  # place the forward declaration after the include so MatrixDim etc. are known
  python - "$G/cukernels_patched.cu" <<'EOF'
import sys
p=sys.argv[1]; s=open(p).read().split('\n')
decl=s.pop(0)
i=[k for k,l in enumerate(s) if 'cukernels.h' in l][0]
s.insert(i+1,decl)
open(p,'w').write('\n'.join(s))
EOF
  # factory without the broken cluster-linearity component
  sed -e 's@#include "cuClusterLinearity.h"@@' \
      -e 's@case 17: pRet = new CuClusterLinearity(nInputs, nOutputs, mpTempBasisDir, pPred); break;@case 17: Error("clusterlinearity not built"); break;@' \
      "$REF/src/CuTNetLib/cuNetwork.cc" > "$G/cuNetwork_patched.cc"
  local NV="$CUDA/bin/nvcc -O2 -w -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC"
  $NV -I"$REF/src/CuBaseLib" -c "$G/cukernels_patched.cu" -o "$G/cukernels.o" &
  $NV -I"$REF/src/CuBaseLib" -c "$REF/src/CuBaseLib/curandkernels.cu" -o "$G/curandkernels.o" &
  local objs=("$G/cukernels.o" "$G/curandkernels.o")
  for f in "$REF"/src/CuBaseLib/*.cc "$REF"/src/CuTNetLib/*.cc; do
    case "$(basename "$f")" in cuClusterLinearity.cc|cuNetwork.cc) continue;; esac
    local o="$G/$(basename "${f%.cc}").o"
    g++ $CXXF $CINC -c "$f" -o "$o" &
    objs+=("$o")
  done
  g++ $CXXF $CINC -I"$REF/src/CuTNetLib" -c "$G/cuNetwork_patched.cc" -o "$G/cuNetwork.o" &
  objs+=("$G/cuNetwork.o")
  for m in TNetCu TRbmCu TRecurrentCu; do
    g++ $CXXF $CINC -c "$REF/src/$m.cc" -o "$G/$m.o" &
  done
  wait
  local kobjs=()
  for f in "$REF"/src/KaldiLib/*.cc; do kobjs+=("$WORK/cpu/$(basename "${f%.cc}").o"); done
  for m in TNetCu TRbmCu TRecurrentCu; do
    g++ -o "$OUT/$m" "$G/$m.o" "${objs[@]}" "${kobjs[@]}" "$OBLAS" \
        -L"$CUDA/lib64" -lcublas -lcudart -L"$CUDA/lib64/stubs" -lcuda -lpthread -Wl,--disable-new-dtags -Wl,-rpath,"$OBLAS_DIR" -Wl,-rpath,"$CUDA/lib64"
    echo "built $OUT/$m"
  done
}

build_cpu
build_gpu
