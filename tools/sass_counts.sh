#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (tcgen05 MMAs, TMEM loads, TMA loads,
# packed fp32 math, mbarrier waits) in the built library.  Usage: tools/sass_counts.sh > profiles/rN_sass_counts.txt
LIB=${1:-fununifrac_b200/csrc/libfununifrac.so}
echo "# cuobjdump -sass $LIB ($(date -u +%Y-%m-%d)), built with: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo"
echo "# kernel | instructions | UTCIMMA(.2CTA) | UTCHMMA | LDTM | UTMALDG | UBLKCP | FADD2 | FFMA2 | SYNCS | UTCBAR | UTCATOMSWS"
cuobjdump -sass "$LIB" | awk '
/Function : /{ if (f != "") print_row(); f=$3; n=0; delete c }
/^[ \t]+\/\*[0-9a-f]+\*\//{ n++; op=$2; sub(/;$/,"",op); if (op ~ /^@/) op=$3; split(op, a, "."); base=a[1];
   if (op ~ /^UTCIMMA/) c["UTCIMMA"]++; if (op ~ /^UTCHMMA/) c["UTCHMMA"]++; if (op ~ /^LDTM/) c["LDTM"]++;
   if (op ~ /^UTMALDG/) c["UTMALDG"]++; if (op ~ /^UBLKCP/) c["UBLKCP"]++; if (op ~ /^FADD2/) c["FADD2"]++; if (op ~ /^FFMA2/) c["FFMA2"]++;
   if (op ~ /^SYNCS/) c["SYNCS"]++; if (op ~ /^UTCBAR/) c["UTCBAR"]++; if (op ~ /^UTCATOMSWS/) c["UTCATOMSWS"]++ }
function print_row() { printf "%s | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d\n", f, n, c["UTCIMMA"], c["UTCHMMA"], c["LDTM"], c["UTMALDG"], c["UBLKCP"], c["FADD2"], c["FFMA2"], c["SYNCS"], c["UTCBAR"], c["UTCATOMSWS"] }
END{ if (f != "") print_row() }' | c++filt | sed 's/fuf:://g; s/gram:://g' | sort
