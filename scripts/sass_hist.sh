#!/bin/bash
# SASS opcode histogram of one kernel: scripts/sass_hist.sh <object or .so> <mangled-name-substring>
# (static counts; used to see what a kernel's instruction mix is before spending GPU time)
set -e
obj=$1; pat=$2
fn=$(cuobjdump -elf "$obj" 2>/dev/null | grep -o "_Z[A-Za-z0-9_]*" | grep "$pat" | grep -v "_param_" | sort -u | head -1)
echo "kernel: $fn"
cuobjdump -sass -fun "$fn" "$obj" > /tmp/sass_hist.txt
echo "instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_hist.txt)"
grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_hist.txt | sed -E 's/^\s+\/\*[0-9a-f]{4}\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/[.;].*//' | sort | uniq -c | sort -rn | head -${3:-30}
