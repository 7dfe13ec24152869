#!/bin/bash
# TEST INFRASTRUCTURE: compile the reference's OWN C-API tests for this path — cpp/tests/c_api/{pagerank,bfs,sssp,extract_paths,katz,hits,weakly_connected_components,eigenvector_centrality,degrees}_test.c,
# unmodified, from where they lie under $REF — against this repository's headers and link them with a libcugraph_c build
# (default: the CPU emulation build, so the binaries run in the GPU-less container; pass the real library on a GPU box
# that has the reference sources).  Outputs only into oracle/_ref/ (git-ignored).  No reference source is copied.
#   bash oracle/ref_ctests/build.sh [path/to/libcugraph_c*.so] [suffix of the binaries]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${REF:-/root/reference}"
LIB="${1:-$ROOT/cugraph_b200/lib/libcugraph_c_emu.so}"
SUFFIX="${2:-}"
OUT="$ROOT/oracle/_ref"
CUDA_INC="${CUDA_INC:-/usr/local/cuda/include}"
[ -d "$REF/cpp/tests/c_api" ] || { echo "reference sources not found under $REF"; exit 3; }
[ -f "$LIB" ] || { echo "library $LIB not built"; exit 4; }
mkdir -p "$OUT"
LIBDIR="$(dirname "$LIB")"; LIBNAME="$(basename "$LIB")"
for t in pagerank bfs sssp extract_paths katz hits weakly_connected_components eigenvector_centrality degrees; do
  gcc -std=gnu11 -O1 -w -I "$ROOT/include" -I "$HERE/include" -I "$CUDA_INC" \
      "$REF/cpp/tests/c_api/${t}_test.c" "$HERE/support.c" -o "$OUT/ref_${t}_test${SUFFIX}" \
      -L "$LIBDIR" -l:"$LIBNAME" -Wl,-rpath,'$ORIGIN/../../cugraph_b200/lib' -Wl,-rpath,"$LIBDIR" -lm
done
echo "built: $OUT/ref_{pagerank,bfs,sssp,extract_paths,katz,hits,weakly_connected_components,eigenvector_centrality,degrees}_test${SUFFIX} (against $LIB)"
