#!/bin/sh
# Static SASS opcode histogram of k_encode_chunks<3,false,true> between its barriers (no GPU needed).
# usage: tools/sass_regions.sh <libm1cu.so> [region: 0 = colour phase, 1 = block phase]
LIB=${1:-ec504_imageencoder_b200/libm1cu.so}; R=${2:-1}
F=$(mktemp)
cuobjdump -sass -fun '_Z15k_encode_chunksILi3ELb0ELb1EEv6M1Geom8M1NzKeysPKhPK8M1TablesPjS7_PsPi' "$LIB" > "$F"
echo "total static instructions: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' "$F")"
if [ "$R" = 0 ]; then A=1; else A=$(grep -n "BAR.SYNC" "$F" | sed -n "${R}p" | cut -d: -f1); fi
B=$(grep -n "BAR.SYNC" "$F" | sed -n "$((R+1))p" | cut -d: -f1)
awk -v a="$A" -v b="$B" 'NR>a && NR<b' "$F" | grep -o '^\s*/\*[0-9a-f]\{4\}\*/\s*\(@!\?U\?P[0-9T] \)\?[A-Z0-9_.]*' | awk '{print $NF}' | sort | uniq -c | sort -rn > "$F.h"
echo "region $R: $(awk '{s+=$1} END {print s}' "$F.h") static instructions"
head -${3:-24} "$F.h"
rm -f "$F" "$F.h"
