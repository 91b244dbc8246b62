SPLATT_B200_MULTI_TIMING=1 timeout 900 python scripts/cpd_config5_shape.py 50000000 1 2 2>&1 | tail -14
