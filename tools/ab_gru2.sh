#!/bin/bash
# A/B of the ConvGRU epilogue grouping on the box itself: rebuild with -DMRB_GRU2_GROUPS=1 / 2 and time both, interleaved.
for round in 1 2; do
  for G in 1 2; do
    MRIDC_B200_NVCC_EXTRA="-DMRB_GRU2_GROUPS=$G" python -m mridc_b200.build > /dev/null 2>&1
    echo "groups=$G round=$round: $(MRIDC_B200_NVCC_EXTRA="-DMRB_GRU2_GROUPS=$G" timeout 40 python tools/time_tc2.py 16 2>&1 | grep 'tc2 gru')"
  done
done
MRIDC_B200_NVCC_EXTRA="-DMRB_GRU2_GROUPS=2" timeout 90 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "tc2_gru" -x 2>&1 | tail -2
