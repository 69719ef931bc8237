#!/bin/bash
# Layout of K1's SASS: where the FADD groups, the shuffles and the atomics sit (is the rare path out of line?)
# usage: tools/sass_layout.sh [stream|strip] [f|13__nv_bfloat16]
kernel=${1:-stream}; elem=${2:-f}
cuobjdump -sass object_keypoints_b200/libokp.so | awk -v k="okp_peaks_${kernel}_kernelI${elem}" '/Function :/ {on = index($0, k) > 0} on' \
 | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+(@!?U?P[0-9T] )?\s*([A-Z0-9_.]+).*/\1 \3/' \
 | awk '{op=$2; sub(/\..*/,"",op); n++; if (op=="FADD") fadd++; if (op=="SHFL"||op=="ATOMS"||op=="LDS"||op=="VOTE"||op=="SYNCS"||op=="CALL"||op=="STL"||op=="LDG") {print n, $1, $2, "fadd_so_far="fadd}} END {print "total", n}'
