#!/bin/bash
# usage: tools/regs.sh file.o [name filter]   - registers / shared memory of every kernel in an object file, one line each
cuobjdump --dump-resource-usage "$1" 2>/dev/null | paste - - | grep -i "${2:-.}" | sed -E 's/.*Function ([^:]+):.*REG:([0-9]+).*SHARED:([0-9]+).*/\2 regs \3 smem \1/' | while read l; do set -- $l; echo "$1 $2 $3 $4 $(echo $5 | c++filt | sed -E 's/vc2::\(anonymous namespace\):://; s/\(vc2::.*//' )"; done
