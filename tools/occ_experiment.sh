for pad in 0 8000 20000 45000; do echo "pad $pad"; RSRX_SMEM_PAD=$pad python tools/dev_gpu_check.py sf 8 1 2>&1 | grep "N=8192"; done
