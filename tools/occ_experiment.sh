# ms per env.step at 8192 envs against the number of envs (warps) per CTA: RSRX_FORCE_WPB overrides launch_cfg's choice
for w in 7 10 14 16 19; do echo "envs per CTA $w"; RSRX_FORCE_WPB=$w timeout 90 python tools/time_step.py sf 8192 2>&1 | grep "N="; done
