#!/bin/bash
# last seconds of the round's GPU budget: the C driver on an R-MAT A A^T whose hub tile-rows take k_s1_heavy (and the heap sort of its long pair
# lists), checked by the driver's serial SPA (structure and values)
mkdir -p gpurun_out
timeout 12 ./driver/test_b200 -d 0 -aat 1 gen:rmat:13:16 16 16 > gpurun_out/r4p_driver_rmat13_aat.txt 2>&1; echo "driver exit $?" >> gpurun_out/r4p_driver_rmat13_aat.txt
tail -n 12 gpurun_out/r4p_driver_rmat13_aat.txt
