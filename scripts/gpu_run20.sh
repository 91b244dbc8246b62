for v in "0 0" "8 0" "2 4" "3 4"; do
  set -- $v
  echo "--- config 5 BATCH=$1 MINB=$2"
  SPLATT_B200_BATCH=$1 SPLATT_B200_MINB=$2 timeout 600 python scripts/quick_bench.py 1000000 200000000 64 3 0 zipf 2>&1 | grep "^mode"
done
