for st in tma lsu; do for m in 1032 1024 0; do echo "STORE=$st mode $m"; GAUCHE_B200_STORE=$st GAUCHE_B200_DEBUG_MODE=$m timeout 120 python tools/profile_umma.py --engine mxf4x1 --rows 16384 --train 100000 --launches 4 | tail -1; done; done
GAUCHE_B200_STORE=lsu timeout 120 python tools/umma_debug.py mxf4x1 | tail -1
GAUCHE_B200_STORE=lsu timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
