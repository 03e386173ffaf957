#!/bin/bash
# build kernel variants (compile-time experiments) into proxtv_b200/variants/lib_<name>.so ; usage: tools/variants.sh name "flags" ...
set -e
cd "$(dirname "$0")/../proxtv_b200/csrc"
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  rm -f kernels_chunked.o
  make EXTRA="$flags" OUT=../variants/lib_$name.so >/dev/null
  echo "$name: $flags -> $(grep -A3 'Compiling entry function.*contigIdLb0ELi256ELi0' kernels_chunked.ptxas.log | grep Used)"
done
rm -f kernels_chunked.o
