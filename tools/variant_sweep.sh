#!/bin/bash
# usage: tools/variant_sweep.sh "name1 name2" "444 4096"
for v in $1; do
  echo "== variant $v"
  SPF_B200_LIB=$PWD/build/libspf_$v.so tools/wave_sweep.sh "$2"
done
