#!/bin/bash
# build the kernel library of an older revision next to the current one: tools/build_rev.sh <rev> <name>  -> proxtv_b200/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
rev=$1; name=$2; tmp=build/rev_$name
rm -rf $tmp; mkdir -p $tmp/proxtv_b200/csrc $tmp/include proxtv_b200/variants
for f in $(git ls-tree --name-only $rev proxtv_b200/csrc/); do git show $rev:$f > $tmp/$f; done
git show $rev:include/proxtv_b200.h > $tmp/include/proxtv_b200.h
make -C $tmp/proxtv_b200/csrc -j8 OUT=../../../../proxtv_b200/variants/lib_$name.so > /dev/null
echo "built proxtv_b200/variants/lib_$name.so from $rev"
