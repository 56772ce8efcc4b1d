#!/usr/bin/env bash
# Build the UNMODIFIED bbc/vc2-reference sources, in place from /root/reference,
# into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
# Test infrastructure only: this is the parity oracle and the CPU baseline arm,
# never part of the product path.  Needs the Boost stand-in in oracle/boost_shim
# because system Boost is absent from this image.
#
#   oracle/_ref/EncodeStream, DecodeStream, DecodeFrame  - the reference command lines
#   oracle/_ref/libvc2ref.so                - reference Library + extern "C" taps (oracle/ref_taps.cpp)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${VC2_REFERENCE_ROOT:-/root/reference}"
OUT="${VC2_REF_OUT:-$HERE/_ref}"
OPT="${VC2_REF_OPT:--O2}"
if [ ! -d "$REF/src/Library" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
R="$REF/src"
F="-std=gnu++14 $OPT -w -fPIC -I$HERE/boost_shim -I$R -I$R/Library"
pids=()
for f in Arrays DataUnit Frame Picture Quantisation Slices Utils VLC WaveletTransform; do
  g++ $F -c "$R/Library/src/$f.cpp" -o "$OUT/obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
OBJS=$(ls "$OUT"/obj/{Arrays,DataUnit,Frame,Picture,Quantisation,Slices,Utils,VLC,WaveletTransform}.o)
g++ $F "$R/EncodeStream/EncodeStream.cpp" "$R/EncodeStream/EncodeParams.cpp" $OBJS -o "$OUT/EncodeStream" &
g++ $F "$R/DecodeStream/DecodeStream.cpp" "$R/DecodeStream/DecodeParams.cpp" $OBJS -o "$OUT/DecodeStream" &
g++ $F "$R/DecodeFrame/DecodeFrame.cpp" "$R/DecodeFrame/DecodeParams.cpp" $OBJS -o "$OUT/DecodeFrame" &
# libvc2ref.so also carries the CLI-level rate control (quantIndicesCBR lives in EncodeStream.cpp)
( g++ $F -Dmain=vc2ref_encodestream_main -c "$R/EncodeStream/EncodeStream.cpp" -o "$OUT/obj/EncodeStream_lib.o" &&
  g++ $F -c "$R/EncodeStream/EncodeParams.cpp" -o "$OUT/obj/EncodeParams_lib.o" &&
  g++ $F -shared "$HERE/ref_taps.cpp" "$OUT/obj/EncodeStream_lib.o" "$OUT/obj/EncodeParams_lib.o" $OBJS -o "$OUT/libvc2ref.so" ) &
wait
echo "built: $(ls "$OUT" | tr '\n' ' ')"
