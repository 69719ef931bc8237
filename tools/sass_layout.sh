#!/bin/bash
# Layout of the strip kernel's SASS: where the FADD groups, the shuffles and the atomics sit (is the rare path out of line?)
cuobjdump -sass -fun _Z22okp_peaks_strip_kernel14CUtensorMap_stPKf12OkpStripPlanff15OkpDecodeTables ${1:-object_keypoints_b200/libokp.so} \
 | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+(@!?U?P[0-9T] )?\s*([A-Z0-9_.]+).*/\1 \3/' \
 | awk '{op=$2; sub(/\..*/,"",op); n++; if (op=="FADD") fadd++; if (op=="SHFL"||op=="ATOMS"||op=="LDS"||op=="VOTE"||op=="SYNCS"||op=="CALL"||op=="STL"||op=="LDG") {print n, $1, $2, "fadd_so_far="fadd}} END {print "total", n}'
