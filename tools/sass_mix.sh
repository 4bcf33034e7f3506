#!/bin/bash
# SASS instruction mix of the hot kernels of psac_b200/libpsacb200.so (no GPU needed): cuobjdump -sass per kernel, opcode counts.
# Output: profiles/r2b_sass_mix.txt.  What to look for: ATOMS (ranking: one shared-memory atomic with return per key) / RED-style
# ATOMS without return (counting), LDG.E.128 / STG.E.128 (vector accesses), no UBLKCP / UTMALDG (no TMA in this library), no HMMA /
# UTCHMMA (no tensor-core work: everything is integer and HBM- or shared-memory-bound).
LIB=psac_b200/libpsacb200.so
OUT=${1:-profiles/r2b_sass_mix.txt}
: > "$OUT"
for pat in 'radix_scatter_seg_kernelINS_8ArraySrcIjjEEjLi512ELi16ELi2ELb0' 'radix_scatter_kernelINS_7TextSrcIjjEEjLi512ELi12ELb1ELi2ELb0' \
           'radix_scatter_kernelINS_6PosSrcIjjEEjLi512ELi16ELb0ELi2ELb0' 'heads_kernelIjjLi1ELb0' 'isa_scatter_kernelIjEE' \
           'suffix_tree_tile_kernelImjNS_11LocalSearchImEELb1' 'gsa_keys_kernelIjEE' 'wide_distinct_kernelIjEE'; do
  fn=$(cuobjdump -elf "$LIB" 2>/dev/null | grep -o "_ZN[A-Za-z0-9_]*${pat}[A-Za-z0-9_]*" | sort -u | head -1)
  [ -z "$fn" ] && { echo "== $pat: not found" >> "$OUT"; continue; }
  echo "== $(echo "$fn" | c++filt | cut -c1-150)" >> "$OUT"
  cuobjdump -sass -fun "$fn" "$LIB" > /tmp/sass_one.txt 2>/dev/null
  echo "   instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_one.txt)" >> "$OUT"
  grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" /tmp/sass_one.txt | awk '{print $2}' | grep -E "^(ATOMS|ATOMG|ATOM|RED|LDG|STG|LDS|STS|BAR|SHFL|MATCH|VOTE|UBLKCP|UTMALDG|HMMA|UTCHMMA|POPC|FLO|SHF|LDSM)" | sort | uniq -c | sort -rn | awk '{printf "   %6d %s\n", $1, $2}' >> "$OUT"
done
echo "== whole library: TMA / tensor-core opcodes" >> "$OUT"
cuobjdump -sass "$LIB" 2>/dev/null | grep -cE "UBLKCP|UTMALDG|UTMASTG|HMMA|UTCHMMA|IMMA" | awk '{print "   " $1 " (none expected: integer work bound by HBM / the shared-memory pipe)"}' >> "$OUT"
cat "$OUT"
