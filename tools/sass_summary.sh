#!/bin/bash
# SASS evidence of one kernel of libjxlb200.so: opcode histogram + every asynchronous-copy / barrier instruction with its address.
# Usage: tools/sass_summary.sh <kernel name substring> > profiles/sass_<kernel>.txt
set -e
k=$1
so=$(dirname "$0")/../jxl_coder_b200/libjxlb200.so
cuobjdump -sass "$so" | awk -v k="$k" '/Function :/{on=index($0,k)>0} on' > /tmp/sass_$k.txt
echo "# cuobjdump -sass jxl_coder_b200/libjxlb200.so, function matching '$k' ($(grep -c '^\s*/\*[0-9a-f]*\*/' /tmp/sass_$k.txt) instructions)"
cuobjdump -res-usage "$so" | grep -A1 "$k" | tail -1
echo "# opcode histogram"
grep -oE '^\s*/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+\s+)?[A-Z0-9_.]+' /tmp/sass_$k.txt | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -40
echo "# asynchronous copies, mbarrier operations, barriers, proxy fences"
grep -E 'UTMALDG|UTMASTG|UBLKCP|LDGSTS|SYNCS|BAR\.SYNC|FENCE|UTMACMDFLUSH|ARRIVE' /tmp/sass_$k.txt | sed 's/\s\+\/\* 0x[0-9a-f]* \*\///'
