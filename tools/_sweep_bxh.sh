#!/bin/bash
# one alt build of the library per line: hybrid Gram kernel rate on the cfg4 block (+ fused scores unless GRAM_ONLY=1)
for n in $LIBS; do
  [ "$n" = main ] && n=""
  echo "== lib$n"
  GAUCHE_B200_LIB=$PWD/gauche_b200/csrc/libgauche_b200$n.so BX_CHECK_SKIP_EXACT=$SKIP BX_CHECK_GRAM_ONLY=$GRAM_ONLY BX_CHECK_NO_SIZES=1 timeout 200 python tools/bxh_check.py 2>&1 | grep -E "gram (bxh|mxf4x1)|sustained|fused|FAILED|EXACT|MISMATCH|rror"
done
