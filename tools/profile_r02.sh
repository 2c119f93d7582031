#!/bin/bash
# Development tool: the ncu captures of round 2 (run on the GPU box through gpurun; reports land in gpurun_out/).
#   launch lists of one sweep (64 replicas in one lane / one replica, no graphs) + one --set full capture per hot
#   kernel family at 64 replicas, + the cluster kernel of DetHubbard (C5).  Read them here with tools/profiles_collect.sh.
set -u
export DQMC_LANES=1
NCU="ncu --clock-control none"
FULL="$NCU --set full --warp-sampling-interval 0 --import-source on -f"
DQMC_NO_GRAPHS=1 timeout 900 $NCU --metrics gpu__time_duration.sum -s 400 -c 3200 --csv --log-file gpurun_out/r02_launches.csv \
    python tools/profile_step.py --sweeps 1 > gpurun_out/r02_launches.log 2>&1
DQMC_NO_GRAPHS=1 timeout 900 $NCU --metrics gpu__time_duration.sum -s 577 -c 4800 --csv --log-file gpurun_out/r02_launches_r1.csv \
    python tools/profile_step.py --replicas 1 --sweeps 1 > gpurun_out/r02_launches_r1.log 2>&1
timeout 600 $FULL --profile-from-start off -k regex:update_window_kernel -s 2 -c 1 -o gpurun_out/prof_window_r02 python tools/profile_update_slice.py > gpurun_out/r02_ncu_window.log 2>&1
timeout 600 $FULL --profile-from-start off -k regex:update_gather_kernel -s 2 -c 1 -o gpurun_out/prof_gather_r02 python tools/profile_update_slice.py > gpurun_out/r02_ncu_gather.log 2>&1
timeout 600 $FULL --profile-from-start off -k regex:zgemm_rank_update2_kernel -s 2 -c 1 -o gpurun_out/prof_flush_r02 python tools/profile_update_slice.py > gpurun_out/r02_ncu_flush.log 2>&1
timeout 600 $FULL -k regex:cb_mult_bulk_kernel -s 30 -c 1 -o gpurun_out/prof_cb_r02 python tools/profile_step.py --sweeps 1 > gpurun_out/r02_ncu_cb.log 2>&1
timeout 600 $FULL -k regex:cb_mult_kernel -s 30 -c 1 -o gpurun_out/prof_cbrow_r02 python tools/profile_step.py --sweeps 1 > gpurun_out/r02_ncu_cbrow.log 2>&1
timeout 600 $FULL -k regex:zgemm_dmma2_kernel -s 6 -c 1 -o gpurun_out/prof_gemm_r02 python tools/profile_step.py --sweeps 1 > gpurun_out/r02_ncu_gemm.log 2>&1
timeout 600 $FULL -k regex:qr_panel2_kernel -s 20 -c 1 -o gpurun_out/prof_panel_r02 python tools/profile_step.py --sweeps 1 > gpurun_out/r02_ncu_panel.log 2>&1
timeout 600 $FULL -k regex:hub_update_slice_kernel -s 5 -c 1 -o gpurun_out/prof_hubbard_r02 python tools/time_configs.py C5 > gpurun_out/r02_ncu_hubbard.log 2>&1
ls -la gpurun_out/prof_*_r02.ncu-rep gpurun_out/r02_launches*.csv
