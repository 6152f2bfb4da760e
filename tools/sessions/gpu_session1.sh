#!/bin/bash
# GPU session 1: gather probe, parity tests, baseline bench, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
{
for g in 0 32 64 128; do ./tools/gather_probe.bin $g 50 8 1024; done
./tools/gather_probe.bin 0 50 32 512
./tools/gather_probe.bin 0 30 128 512
./tools/gather_probe.bin 32 30 128 512
./tools/gather_probe.bin 0 50 160 256
} > gpurun_out/s1_probe.txt 2>&1
for g in 0 32; do
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --csv --log-file gpurun_out/s1_probe_ncu_g$g.csv ./tools/gather_probe.bin $g 50 8 256 > /dev/null 2>&1
done
ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --csv --log-file gpurun_out/s1_probe_ncu_128.csv ./tools/gather_probe.bin 0 30 128 128 > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/s1_pytest.txt
( time timeout 900 python bench.py ) > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -3 gpurun_out/s1_pytest.txt
cat gpurun_out/s1_probe.txt
