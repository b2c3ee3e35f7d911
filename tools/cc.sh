#!/bin/bash
# dev helper: compile ONE csrc file with the library's flags and print ptxas resource usage (demangled)
# usage: tools/cc.sh act_quant [grep-pattern]
f=$1; pat=${2:-.}
cd "$(dirname "$0")/../omniquant_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC,-ffp-contract=off \
  -Xptxas -v -c $f.cu -o /tmp/$f.o 2>&1 | c++filt | grep -E "error|warning|Function properties|Used|spill" | paste - - - | grep -E "$pat|error" | sed -e 's/ptxas info    : //g' -e 's/(anonymous namespace):://' | cut -c1-260
