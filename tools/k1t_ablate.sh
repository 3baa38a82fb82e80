for d in 0 1 2 6 7; do echo "== MVF_K1T_DBG=$d"; MVF_K1T_DBG=$d timeout 100 python tools/k1t_debug.py timing 2>&1 | grep "tensor_cores=True"; done
