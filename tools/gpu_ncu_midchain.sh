#!/bin/bash
# ncu --set full of one MID-CHAIN whole-batch launch of the sector QR (site n=10, grid 12 x 888), the Theta contraction (cut l=10)
# and the energy sweep (saturated bonds), second warm-up step of tools/profile_step.py; outputs gpurun_out/prof_{sqr,theta,energy}_r02.ncu-rep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
APP="python tools/profile_step.py --chi 32 --batch 888 --steps 1 --warmup 2"
NCU="ncu --set full --clock-control none --import-source on --launch-count 1 -f"
$NCU -k regex:k_sector_qr --launch-skip 550 -o gpurun_out/prof_sqr_r02 $APP > gpurun_out/ncu_sqr_r02.log 2>&1; echo "sqr rc=$?"
$NCU -k regex:^k_theta$ --launch-skip 549 -o gpurun_out/prof_theta_r02 $APP > gpurun_out/ncu_theta_r02.log 2>&1; echo "theta rc=$?"
$NCU -k regex:^k_energy$ --launch-skip 4 -o gpurun_out/prof_energy_r02 $APP > gpurun_out/ncu_energy_r02.log 2>&1; echo "energy rc=$?"
ls -la gpurun_out/prof_sqr_r02.ncu-rep gpurun_out/prof_theta_r02.ncu-rep gpurun_out/prof_energy_r02.ncu-rep
