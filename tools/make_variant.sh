#!/bin/bash
# tools/make_variant.sh NAME "EXTRA nvcc flags": builds a copy of the library under build/variants/NAME
set -e
name=$1; shift
dst=build/variants/$name
rm -rf $dst; mkdir -p $dst
cp -r gr-uwspr_b200 include $dst/
rm -rf $dst/gr-uwspr_b200/build $dst/gr-uwspr_b200/libuwspr_b200.so
make -s -C $dst/gr-uwspr_b200 EXTRA="$*" > /dev/null
grep -E "spill|Used" $dst/gr-uwspr_b200/build/fine.ptxas.log | head -2
