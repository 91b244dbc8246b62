timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "leaf_tiled or engine_device or sharded" 2>&1 | tail -5
for kt in -1 0 128 256 400 600 1000; do echo "KTILE=$kt"; KTILE=$kt timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 0 2>&1 | grep -E 'mode 0|stream 0'; done
echo L1 tile 96KB; SPLATT_B200_L1_TILE_KB=96 KTILE=0 timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 0 2>&1 | grep -E 'mode 0|stream 0'
