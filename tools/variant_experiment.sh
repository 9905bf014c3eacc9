# times every rsr_mjx_b200/librsrx_w*.so variant (each run under its own timeout: a variant may hang)
cp rsr_mjx_b200/librsrx.so /tmp/librsrx_orig.so
for f in rsr_mjx_b200/librsrx_w*.so; do cp $f rsr_mjx_b200/librsrx.so; echo "variant $f"; timeout 120 python tools/time_step_stationary.py ${1:-sf} ${2:-8192} 2>&1 | grep "N=" || echo "  (failed or timed out)"; done
cp /tmp/librsrx_orig.so rsr_mjx_b200/librsrx.so
