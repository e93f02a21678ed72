#!/bin/bash
# opcode histogram of one kernel (mangled-name regex) in an object file:  tools/sass_hist.sh <obj> <regex> [dump-file]
OBJ=$1; RE=$2; OUT=${3:-/tmp/k.sass}
FN=$(cuobjdump -sass $OBJ | grep -oE "Function : \S+" | awk '{print $3}' | grep -E "$RE" | head -1)
echo "kernel: $(echo $FN | c++filt | cut -c1-120)"
cuobjdump -sass -fun "$FN" $OBJ > $OUT
grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" $OUT | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${4:-40}
