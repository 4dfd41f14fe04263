#!/bin/bash
# usage: scripts/sass_fn.sh <mangled function name> [lib]  -> SASS of one kernel, one instruction per line (address + text)
LIB=${2:-spinwalk_b200/libspinwalk_b200.so}
cuobjdump -sass "$LIB" | awk -v fn="$1" '/Function :/{on=($3==fn)} on' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#^\s+/\*([0-9a-f]+)\*/\s+#\1 #; s#\s*/\* 0x[0-9a-f]+ \*/##; s#\s+;#;#'
