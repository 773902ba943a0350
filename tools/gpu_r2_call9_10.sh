#!/bin/bash
# calls 9 and 10 in one box
bash tools/gpu_r2_call9.sh
bash tools/gpu_r2_call10.sh
