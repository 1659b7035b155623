#!/bin/bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference (scalable-arch/CAL_22-MPC) from the sources where they
# lie under $REF (default /root/reference) into oracle/_ref/ (git-ignored, travels with gpurun):
#   oracle/_ref/compressor    - the reference CLI (src/main.cpp), used via bin/run-style calls
#   oracle/_ref/libmpcref.so  - reference classes + oracle/ref_harness.cpp (per-line driver)
# The reference's own build system is NOT run.  Five third-party headers it needs are absent
# from the image (cxxopts, strutil, jsoncpp, libnpy, boost::hash); oracle/shims/ stands in
# for them (none does hot-path arithmetic).  fmt comes header-only from torch's include tree.
# g++ 13 needs three forced includes the sources rely on transitively.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)"; exit 0
fi
SP="$(python -c 'import site;print(site.getsitepackages()[0])')"
mkdir -p "$OUT"
FLAGS="-O3 -w -std=c++17 -include cstdint -include cstring -include algorithm -include memory -DFMT_HEADER_ONLY \
  -I$HERE/shims -I$SP/torch/include -I$SP/include/cudnn_frontend/thirdparty -I$REF/src"
SRCS="$REF/src/utils.cpp $REF/src/loader/*.cpp $REF/src/compressor/*.cpp $REF/src/compressor/VPCmodules/*.cpp"
g++ $FLAGS $REF/src/main.cpp $SRCS -o "$OUT/compressor" &
g++ $FLAGS -fPIC -shared "$HERE/ref_harness.cpp" $SRCS -o "$OUT/libmpcref.so" &
wait
ls -la "$OUT"
