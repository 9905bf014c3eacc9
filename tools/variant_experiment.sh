cp rsr_mjx_b200/librsrx.so /tmp/librsrx_orig.so
echo "w8 (8 warps/SM)"; python tools/dev_gpu_check.py sf 8 1 2>&1 | grep "N=8192"
cp rsr_mjx_b200/librsrx_w6.so rsr_mjx_b200/librsrx.so; echo "w6 pad (6 warps/SM)"; RSRX_SMEM_PAD=70000 python tools/dev_gpu_check.py sf 8 1 2>&1 | grep "N=8192"
cp rsr_mjx_b200/librsrx_w4.so rsr_mjx_b200/librsrx.so; echo "w4 pad (4 warps/SM)"; RSRX_SMEM_PAD=120000 python tools/dev_gpu_check.py sf 8 1 2>&1 | grep "N=8192"
cp /tmp/librsrx_orig.so rsr_mjx_b200/librsrx.so
