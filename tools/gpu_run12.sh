python -m pytest tests/test_scan_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in 0 5 7; do STRGPU_SCAN_VARIANT=$v STRGPU_MAX_STAGE=0 python tools/profile_scan.py | grep us_per; done
python tools/profile_scan.py | grep us_per
python tools/profile_scan.py --align 4 | grep us_per
