#!/usr/bin/env bash
# Source-level drop-in check: the UNMODIFIED command-line sources of bbc/vc2-reference (EncodeStream.cpp, DecodeStream.cpp
# and their parameter parsers) and its unmodified plumbing (Arrays, Picture, Frame, DataUnit, Utils, VLC - objects of
# oracle/_ref/obj) are linked with host/dropin/{WaveletTransform,Quantisation,Slices}.cpp - the replacement Library
# bodies over the CUDA C-ABI - instead of the reference's own three files.  Output: tests/dropin/_build/ (git-ignored,
# travels to the GPU box).  Build container only: needs /root/reference and oracle/_ref/obj (oracle/build_ref.sh).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${VC2_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_build"
if [ ! -d "$REF/src/Library" ]; then
  echo "dropin/build.sh: $REF not present (GPU box?) - keeping prebuilt tests/dropin/_build" >&2
  exit 0
fi
[ -f "$ROOT/oracle/_ref/obj/DataUnit.o" ] || bash "$ROOT/oracle/build_ref.sh"
mkdir -p "$OUT"
R="$REF/src"
F="-std=gnu++14 -O2 -w -fPIC -I$ROOT/oracle/boost_shim -I$R -I$R/Library -I$ROOT/include -I$ROOT/host/dropin"
for f in WaveletTransform Quantisation Slices; do
  g++ $F -c "$ROOT/host/dropin/$f.cpp" -o "$OUT/$f.o" &
done
wait
PLUMBING=$(ls "$ROOT"/oracle/_ref/obj/{Arrays,DataUnit,Frame,Picture,Utils,VLC}.o)
OURS="$OUT/WaveletTransform.o $OUT/Quantisation.o $OUT/Slices.o"
L="-L$ROOT/vc2_reference_b200 -lvc2b200 -Wl,-rpath,\$ORIGIN/../../../vc2_reference_b200"
g++ $F "$R/EncodeStream/EncodeStream.cpp" "$R/EncodeStream/EncodeParams.cpp" $OURS $PLUMBING $L -o "$OUT/EncodeStream" &
g++ $F "$R/DecodeStream/DecodeStream.cpp" "$R/DecodeStream/DecodeParams.cpp" $OURS $PLUMBING $L -o "$OUT/DecodeStream" &
wait
echo "built: $(ls "$OUT" | tr '\n' ' ')"
