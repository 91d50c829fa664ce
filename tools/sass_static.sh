#!/bin/bash
# tools/sass_static.sh LIB.so KERNEL_SUBSTRING -- static SASS opcode histogram of one kernel (no GPU needed)
cuobjdump -sass "$1" | awk -v k="$2" '/Function :/{on=index($0,k)>0} on' | grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ )?[A-Z0-9_.]+" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-16}
