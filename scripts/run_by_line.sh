# Per-source-line stall tables of the two individual-level kernels (ncu --set full + SASS line info).
set -x
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -s 1 -c 1 -f -o gpurun_out/bl_i8 python bench.py --storage i8 --single-phase --genes 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-dense-arm --max-iter 4 --min-tol 0 > /dev/null 2>&1
python scripts/ncu_by_line.py gpurun_out/bl_i8.ncu-rep sushie_b200/_lib/libsushie_b200.so ibss_kernelILi3EaLb0ELi1E 32 > gpurun_out/r01c_ind_i8_by_line.txt 2>&1
ncu --set full --import-source on --clock-control none -k regex:ibss_kernel -s 1 -c 1 -f -o gpurun_out/bl_f64 python bench.py --storage f64 --single-phase --genes 296 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --max-iter 4 --min-tol 0 > /dev/null 2>&1
python scripts/ncu_by_line.py gpurun_out/bl_f64.ncu-rep sushie_b200/_lib/libsushie_b200.so ibss_kernelILi3EdLb0ELi1E 32 > gpurun_out/r01c_ind_f64_by_line.txt 2>&1
rm -f gpurun_out/*.ncu-rep
head -8 gpurun_out/r01c_ind_i8_by_line.txt | cut -c1-160
head -8 gpurun_out/r01c_ind_f64_by_line.txt | cut -c1-160
