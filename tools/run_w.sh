timeout 100 python tools/lz_ck_probe.py 2>&1 | tail -1
for f in tests/golden/regress/lzwalk_*.bin; do b=$(basename $f | sed 's/_l2//'); cp $f /tmp/$b; timeout 60 python tools/gpu_lz_repro.py /tmp/$b spec=1 2>&1 | tail -1; done
timeout 300 python tools/gpu_lz_check.py 50 tma=1,spec=1,region=1 2>&1 | tail -6
