#!/bin/bash
# usage: tools/ablate.sh "0 1 2 ..."  -- builds variants/libspf_ablN.so (-DSPF_ABL=N on top of $ABL_BASE flags) in parallel
for n in $1; do
  ( tools/build_variant.sh abl$n -DSPF_ABL=$n $ABL_BASE > /dev/null 2>&1; cp build/libspf_abl$n.so variants/ ) &
done
wait
ls variants | grep abl | tr '\n' ' '
