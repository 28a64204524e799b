for e in mxf4rs mxf4rs224; do timeout 120 python tools/umma_debug.py $e 2>&1 | tail -3; done
for e in mxf4x1 mxf4rs mxf4rs224; do echo "== $e"; timeout 120 python tools/profile_umma.py --engine $e --rows 16384 --train 100000 --launches 6 | tail -1; done
for m in 8 4 1; do echo "mxf4rs mode $m"; GAUCHE_B200_DEBUG_MODE=$m timeout 120 python tools/profile_umma.py --engine mxf4rs --rows 16384 --train 100000 --launches 4 | tail -1; done
