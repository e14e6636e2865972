#!/bin/bash
mkdir -p gpurun_out
timeout 400 ./lab_bin/take_lab2 100000000 100000000 3 2>&1 | tee gpurun_out/take_lab2_100m.txt
WMB=32 timeout 300 ./lab_bin/take_lab2 1000000000 125000000 3 2>&1 | tee gpurun_out/take_lab2_1b.txt
WMB=16 timeout 300 ./lab_bin/take_lab2 1000000000 125000000 3 2>&1 | tee -a gpurun_out/take_lab2_1b.txt
WMB=8 timeout 300 ./lab_bin/take_lab2 1000000000 125000000 3 2>&1 | tee -a gpurun_out/take_lab2_1b.txt
