#!/bin/bash
# usage: scripts/sass_kernel.sh <mangled-name-substring> [so]  -> /tmp/kernel.sass + opcode histogram
SO=${2:-gym_pomdp_b200/csrc/libpomdp_b200.so}
cuobjdump -sass $SO | awk -v pat="$1" '/Function : /{f = index($0, pat) > 0} f' > /tmp/kernel.sass
echo "lines: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/kernel.sass)"
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/kernel.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-25}
