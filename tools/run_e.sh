timeout 120 python tools/proj_only.py 1024 8192 8192 20 && timeout 300 ncu --set full --clock-control none -k regex:bt_umma_kernel -c 1 -o gpurun_out/r2_proj_tail -f python tools/proj_only.py 1024 8192 8192 3 2>&1 | tail -2
python tools/ncu_summary.py gpurun_out/r2_proj_tail.ncu-rep 2>&1 | grep -i "dram__bytes\|duration\|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed\|registers\|xbar\|lts__t_sector_hit" | head
timeout 120 python tools/proj_only.py 1024 2048 8192 20; timeout 120 python tools/proj_only.py 256 8192 8192 20; timeout 120 python tools/proj_only.py 4096 2048 2048 20
