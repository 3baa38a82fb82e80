for d in 0 8 16 24; do echo "== MVF_K1T_DBG=$d"; MVF_K1T_DBG=$d timeout 100 python tools/k1t_debug.py prof 2>&1 | grep "MMA\|epilogue"; done
